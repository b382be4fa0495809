"""Throughput of the general FP32 renderers on cornell_box (GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ray_tracing_weekend_b200 as rtw
world, lights, cb = rtw.scenes.cornell_box()
sc = rtw.Scene(world, lights)
cam = cb.with_vfov(40.).with_aspect_ratio(1.0).with_max_depth(50).with_image_width(1024).with_image_height(1024).with_samples_per_pixel(256).build()
imgs = {}
for name, flags in (("pooled", 0), ("lane-per-pixel", rtw.RTW_FLAG_LANE_PER_PIXEL)):
    best = None
    for _ in range(3):
        img, _, st = sc.render(cam, rtw.RenderOptions(seed=1, precision=rtw.RTW_F32, flags=flags), want_rgb8=False)
        if best is None or st["kernel_ms"] < best["kernel_ms"]:
            best = st
    imgs[name] = img
    print(name, "ms", round(best["kernel_ms"], 2), "Mrays/s", round(best["rays"] / best["kernel_ms"] / 1e3, 1), "rays/path", best["rays"] / best["paths"])
a, b = imgs["pooled"], imgs["lane-per-pixel"]
ok = np.isfinite(a) & np.isfinite(b)
print("pooled vs lane: max rel diff of sums", float(np.max(np.abs(a[ok] - b[ok]) / np.maximum(1e-3, np.abs(b[ok])))), "nan mask equal", bool((np.isnan(a) == np.isnan(b)).all()))
