"""Where does the end-to-end step (scene upload + render + read-back) spend its time?"""
import time, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ray_tracing_weekend_b200 as R
W, H, SPP = 1920, 1080, int(os.environ.get("SPP", "500"))
world, lights, cb = R.scenes.simple(20261018)
cam = cb.with_vfov(40.).with_aspect_ratio(W / H).with_max_depth(50).with_image_width(W).with_image_height(H).with_samples_per_pixel(SPP).build()
opts = R.RenderOptions()
def t():
    torch.cuda.synchronize(); return time.perf_counter()
for it in range(4):
    t0 = t(); sc = R.Scene(world, lights); t1 = t()
    _, rgb8, st = sc.render(cam, opts, want_sum=False, want_rgb8=True); t2 = t()
    sc.close(); t3 = t()
    print(f"iter {it}: scene_create {1e3*(t1-t0):.2f} ms  render {1e3*(t2-t1):.2f} ms (kernel {st['kernel_ms']:.2f}, device total {st['total_ms']:.2f})  destroy {1e3*(t3-t2):.2f} ms")
sc = R.Scene(world, lights)
for it in range(3):
    t1 = t(); _, rgb8, st = sc.render(cam, opts, want_sum=False, want_rgb8=True); t2 = t()
    print(f"resident scene: render {1e3*(t2-t1):.2f} ms (kernel {st['kernel_ms']:.2f}, device total {st['total_ms']:.2f})")
