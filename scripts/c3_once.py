"""One or more C3 frames (3840x2160, 1024 spp) through rtw_render: kernel time per frame.  usage: c3_once.py [reps]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ray_tracing_weekend_b200 as R
SEED = 20261018
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
arr = R.scenes.simple_arrays(SEED)
sc = R.Scene.from_arrays(arr["spheres"], arr["sphere_materials"], arr["planes"], arr["plane_materials"], arr["lights"])
cam = arr["cam"].with_vfov(40.).with_aspect_ratio(16 / 9).with_max_depth(50).with_image_width(3840).with_image_height(2160).with_samples_per_pixel(spp).build()
for k in range(reps):
    _, _, st = sc.render(cam, R.RenderOptions(seed=SEED), want_sum=False, want_rgb8=True)
    print(json.dumps(dict(rep=k, kernel_ms=round(st["kernel_ms"], 2), total_ms=round(st["total_ms"], 2), rays=st["rays"], launches=st["launches"])), flush=True)
sc.close()
