"""Small renders of every kernel for compute-sanitizer (memcheck / racecheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ray_tracing_weekend_b200 as R
world, lights, cb = R.scenes.simple(20261018)
sc = R.Scene(world, lights)
cam = cb.with_vfov(40.).with_aspect_ratio(40 / 24).with_max_depth(50).with_image_width(40).with_image_height(24).with_samples_per_pixel(6).build()
for prec, mode, flags in ((R.RTW_F32, R.RTW_WAVEFRONT, 0), (R.RTW_F32, R.RTW_MEGAKERNEL, 0), (R.RTW_F32, R.RTW_MEGAKERNEL, R.RTW_FLAG_LANE_PER_PIXEL),
                          (R.RTW_F32, R.RTW_WAVEFRONT, R.RTW_FLAG_COUNT_EVENTS), (R.RTW_F64, R.RTW_MEGAKERNEL, 0)):
    img, rgb8, st = sc.render(cam, R.RenderOptions(precision=prec, mode=mode, flags=flags))
    print(prec, mode, flags, st["rays"], float(np.nan_to_num(img).sum()))
o = np.tile([10., 5., 10.], (256, 1)); d = np.random.default_rng(0).normal(size=(256, 3)) * 0.2 - o
print(sc.trace_batch(o, d)[0][:8], sc.trace_batch(o, d, precision=R.RTW_F64)[0][:8])
a = R.scenes.simple_arrays(5, 40)          # global-memory scene + light BVH
big = R.Scene.from_arrays(a["spheres"], a["sphere_materials"], a["planes"], a["plane_materials"], a["lights"])
for mode in (R.RTW_WAVEFRONT, R.RTW_MEGAKERNEL):
    img, _, st = big.render(cam, R.RenderOptions(mode=mode))
    print("big", mode, st["rays"])
print("done")
