"""Where does the per-launch fixed cost of the wavefront kernel come from?  Kernel time of the 1080p frame at 4..62 spp for max_depth 1, 2, 5, 50."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ray_tracing_weekend_b200 as R
SEED = 20261018
world, lights, cb = R.scenes.simple(SEED)
sc = R.Scene(world, lights)
for depth in (1, 2, 5, 50):
    row = {}
    for spp in (4, 8, 16, 32, 62):
        cam = cb.with_vfov(40.).with_aspect_ratio(16 / 9).with_max_depth(depth).with_image_width(1920).with_image_height(1080).with_samples_per_pixel(spp).build()
        best = 1e9
        for _ in range(4):
            _, _, st = sc.render(cam, R.RenderOptions(seed=SEED), want_sum=False, want_rgb8=True)
            best = min(best, st["kernel_ms"])
        row[spp] = round(best, 3)
    slope = (row[62] - row[32]) / 30
    print(json.dumps(dict(depth=depth, ms=row, ms_per_spp=round(slope, 4), intercept=round(row[62] - 62 * slope, 3))), flush=True)
