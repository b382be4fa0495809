"""Device-built LBVH (csrc/bvh_device.cu) vs host binned-SAH tree on BASELINE C4 (1 M spheres): scene-create time, render time,
node visits per ray, image agreement.  usage: lbvh_vs_sah.py [grid half-size n = 500] [spp = 16]   (GPU box)"""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ray_tracing_weekend_b200 as R
SEED = 20261018
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 16
A = R.scenes.simple_arrays(SEED, n)
cb = A["cam"].with_vfov(40.).with_aspect_ratio(16 / 9).with_max_depth(50).with_image_width(1920).with_image_height(1080).with_samples_per_pixel(spp)
if n >= 100:
    cb = cb.with_lookfrom((60., 30., 60.)).with_focus_dist(float(np.linalg.norm([60., 30., 60.])))
cam = cb.build()
imgs = {}
for name, mode in (("host-sah", R.RTW_BVH_HOST_SAH), ("device-lbvh", R.RTW_BVH_DEVICE_LBVH)):
    R.set_bvh_builder(mode)
    best_create = 1e9
    creates = []
    for _ in range(3):
        t0 = time.perf_counter()
        sc = R.Scene.from_arrays(A["spheres"], A["sphere_materials"], A["planes"], A["plane_materials"], A["lights"])
        creates.append(round(time.perf_counter() - t0, 4))
        best_create = min(best_create, creates[-1])
        info = sc.info()
        if _ < 2:
            sc.close()
    best = None
    for _ in range(2):
        img, _, st = sc.render(cam, R.RenderOptions(seed=SEED), want_rgb8=False)
        if best is None or st["kernel_ms"] < best["kernel_ms"]:
            best = st
    imgs[name] = img
    _, _, cnt = sc.render(cam, R.RenderOptions(seed=SEED, flags=R.RTW_FLAG_COUNT_EVENTS), want_sum=False, want_rgb8=False)
    print(name, "node visits / ray", round(cnt["node_visits"] / cnt["rays"], 2), "sphere tests / ray", round(cnt["sphere_tests"] / cnt["rays"], 2),
          "light tests / lambertian", round(cnt["light_tests"] / max(1, cnt["lambertian"]), 1), "counting kernel ms", round(cnt["kernel_ms"], 1))
    for mode, mname in ((R.RTW_MEGAKERNEL, "megakernel"),):
        _, _, st2 = sc.render(cam, R.RenderOptions(seed=SEED, mode=mode), want_sum=False, want_rgb8=False)
        print(name, mname, "ms", round(st2["kernel_ms"], 1))
    print(json.dumps(dict(builder=name, spheres=sc.n_spheres, scene_create_s=round(best_create, 4), scene_create_all_s=creates, info=info, kernel_ms=round(best["kernel_ms"], 2),
                          mrays_per_s=round(best["rays"] / best["kernel_ms"] * 1e-3, 1), rays=best["rays"])), flush=True)
    sc.close()
R.set_bvh_builder(R.RTW_BVH_AUTO)
a, b = imgs["host-sah"], imgs["device-lbvh"]
same = np.isclose(a, b, rtol=1e-6, atol=1e-6, equal_nan=True).all(axis=2)
print("FP32 images: identical pixels", float(same.mean()))
