"""Per-warp timeline of the wavefront kernel (diagnostic build -DRTW_TIMELINE): when does the queue reach its background-only chunks,
when does it run dry, when does each warp exit, how many paths were in flight then.  usage: timeline_probe.py [spp ...]"""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("RTW_LIBRARY", os.path.join(ROOT, "ray_tracing_weekend_b200", "lib", "variants", "librtw_cuda_timeline.so"))
import numpy as np
import ray_tracing_weekend_b200 as R
SEED = 20261018
lib = C.CDLL(os.environ["RTW_LIBRARY"])
lib.rtw_debug_timeline.argtypes = [C.c_void_p, C.c_size_t]
world, lights, cb = R.scenes.simple(SEED)
sc = R.Scene(world, lights)
for spp in [int(a) for a in sys.argv[1:]] or [62, 8]:
    cam = cb.with_vfov(40.).with_aspect_ratio(16 / 9).with_max_depth(50).with_image_width(1920).with_image_height(1080).with_samples_per_pixel(spp).build()
    for _ in range(3):
        _, _, st = sc.render(cam, R.RenderOptions(seed=SEED), want_sum=False, want_rgb8=True)
    buf = np.zeros(148 * 32 * 8, dtype=np.uint64)
    rc = lib.rtw_debug_timeline(buf.ctypes.data, buf.size)
    t = buf.reshape(-1, 8)
    t = t[t[:, 0] > 0].astype(np.float64)
    t0 = t[:, 0].min()
    ms = lambda x: (x - t0) * 1e-6
    q = lambda x: [round(float(v), 3) for v in np.quantile(x, [0, 0.01, 0.5, 0.99, 1])]
    cheap = t[t[:, 1] > 0]
    print(json.dumps(dict(spp=spp, rc=rc, kernel_ms=round(st["kernel_ms"], 3), warps=len(t), start_ms=q(ms(t[:, 0])),
                          first_cheap_chunk_ms=q(ms(cheap[:, 1])) if len(cheap) else None, warps_seeing_cheap=len(cheap),
                          queue_dry_ms=q(ms(t[:, 2])), exit_ms=q(ms(t[:, 3])), drain_ms_per_warp=q((t[:, 3] - t[:, 2]) * 1e-6),
                          in_flight_at_dry=q(t[:, 4]), passes_after_dry=q(t[:, 5]), paths_per_warp=q(t[:, 6]))), flush=True)
sc.close()
