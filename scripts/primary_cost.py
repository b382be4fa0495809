"""How much of the C2 frame is primary rays?  Renders the frame at max_depth = 1 (camera ray only), 2 and 50 and prints time and events."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ray_tracing_weekend_b200 as R
SEED = 20261018
arrays = R.scenes.simple_arrays(SEED)
sc = R.Scene.from_arrays(arrays["spheres"], arrays["sphere_materials"], arrays["planes"], arrays["plane_materials"], arrays["lights"])
for depth in (1, 2, 3, 50):
    cam = arrays["cam"].with_vfov(40.).with_aspect_ratio(16 / 9).with_max_depth(depth).with_image_width(1920).with_image_height(1080).with_samples_per_pixel(100).build()
    best = None
    for _ in range(3):
        _, _, st = sc.render(cam, R.RenderOptions(seed=SEED, mode=R.RTW_WAVEFRONT), want_sum=False, want_rgb8=True)
        best = st if best is None or st["kernel_ms"] < best["kernel_ms"] else best
    _, _, cnt = sc.render(cam, R.RenderOptions(seed=SEED, mode=R.RTW_WAVEFRONT, flags=R.RTW_FLAG_COUNT_EVENTS), want_sum=False, want_rgb8=False)
    print(json.dumps(dict(depth=depth, kernel_ms=best["kernel_ms"], rays=cnt["rays"], node_visits=cnt["node_visits"], sphere_tests=cnt["sphere_tests"],
                          light_tests=cnt["light_tests"], missed=cnt["missed"], lambertian=cnt["lambertian"])), flush=True)
sc.close()
