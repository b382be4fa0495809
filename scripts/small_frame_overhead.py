"""Fixed cost of a launch of the wavefront renderer: kernel time of the 1080p frame at several spp (GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ray_tracing_weekend_b200 as R
SEED = 20261018
world, lights, cb = R.scenes.simple(SEED)
sc = R.Scene(world, lights)
res = {}
for spp in (8, 16, 31, 32, 62, 63, 125, 250, 500):
    cam = cb.with_vfov(40.).with_aspect_ratio(16 / 9).with_max_depth(50).with_image_width(1920).with_image_height(1080).with_samples_per_pixel(spp).build()
    best = None
    for _ in range(4):
        _, _, st = sc.render(cam, R.RenderOptions(seed=SEED), want_sum=False, want_rgb8=True)
        best = st if best is None or st["kernel_ms"] < best["kernel_ms"] else best
    res[spp] = best["kernel_ms"]
    print(spp, "spp", round(best["kernel_ms"], 3), "ms", round(best["rays"] / best["kernel_ms"] / 1e3), "Mrays/s", flush=True)
slope = (res[500] - res[250]) / 250
print("ms per spp (250..500):", round(slope, 4), "-> fixed cost at 62 spp:", round(res[62] - 62 * slope, 3), "ms; at 8 spp:", round(res[8] - 8 * slope, 3))
