"""FP32 C1 / 1024 spp (fix_nan) against the oracle fixture for a sweep of tmin: which FP32 tmin reproduces the f64 reference's
self-intersection statistics best?"""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ray_tracing_weekend_b200 as R
SEED = 20261018
g = np.load("tests/golden/c1_1024spp_oracle_rgb8.npz")
a, b = g["fix_a"], g["fix_b"]
psnr = lambda x, y: 10 * np.log10(255.0 ** 2 / ((x.astype(float) - y.astype(float)) ** 2).mean())
mae = lambda x, y: np.abs(x.astype(float) - y.astype(float)).mean()
world, lights, cb = R.scenes.simple(SEED)
sc = R.Scene(world, lights)
cam = cb.with_vfov(40.).with_aspect_ratio(400 / 225).with_max_depth(50).with_image_width(400).with_image_height(225).with_samples_per_pixel(1024).build()
print(json.dumps(dict(floor_psnr=psnr(a, b), floor_mae=mae(a, b), oracle_mean=float(a.mean()), oracle_rays_per_path=float(g["fix_a_rays_per_path"]))))
eps = 2.0 ** -23
for prec, name in ((R.RTW_F32, "f32"), (R.RTW_F64, "f64")):
    for k in ([-1.0, 4, 2, 1, 0.75, 0.5, 0.35, 0.25, 0.125, 0.0625, 0.0] if prec == R.RTW_F32 else [-1.0]):
        tmin = -1.0 if k < 0 else k * eps
        spp_cam = cam
        _, rgb8, st = sc.render(spp_cam, R.RenderOptions(seed=SEED + 2, precision=prec, tmin=tmin, flags=R.RTW_FLAG_FIX_NAN), want_sum=False)
        sph = (a != 255).any(axis=2)
        print(json.dumps(dict(precision=name, tmin_over_eps32=k, psnr=psnr(rgb8, a), mae=mae(rgb8, a), mean=float(rgb8.mean()), rays_per_path=st["rays"] / st["paths"],
                              mean_on_spheres=float(rgb8[sph].mean()), oracle_mean_on_spheres=float(a[sph].mean()), kernel_ms=st["kernel_ms"])), flush=True)
sc.close()
