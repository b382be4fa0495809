"""cornell_box 1024x1024 / 256 spp: the three FP32 renderers of the general path (GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ray_tracing_weekend_b200 as rtw
world, lights, cb = rtw.scenes.cornell_box()
sc = rtw.Scene(world, lights)
cam = cb.with_vfov(40.).with_aspect_ratio(1.0).with_max_depth(50).with_image_width(1024).with_image_height(1024).with_samples_per_pixel(256).build()
for name, mode, flags in (("wavefront", rtw.RTW_WAVEFRONT, 0), ("pooled megakernel", rtw.RTW_MEGAKERNEL, 0), ("lane per pixel", rtw.RTW_MEGAKERNEL, rtw.RTW_FLAG_LANE_PER_PIXEL)):
    best = None
    for _ in range(3):
        _, _, st = sc.render(cam, rtw.RenderOptions(seed=1, mode=mode, flags=flags), want_sum=False)
        best = st if best is None or st["kernel_ms"] < best["kernel_ms"] else best
    print(name, "ms", round(best["kernel_ms"], 2), "Mrays/s", round(best["rays"] / best["kernel_ms"] / 1e3, 1), flush=True)
cam64 = cb.with_vfov(40.).with_aspect_ratio(1.0).with_max_depth(50).with_image_width(1024).with_image_height(1024).with_samples_per_pixel(32).build()
best = None
for _ in range(3):
    _, _, st = sc.render(cam64, rtw.RenderOptions(seed=1, precision=rtw.RTW_F64), want_sum=False)
    best = st if best is None or st["kernel_ms"] < best["kernel_ms"] else best
print("f64 (32 spp) ms", round(best["kernel_ms"], 2), "Mrays/s", round(best["rays"] / best["kernel_ms"] / 1e3, 1))
