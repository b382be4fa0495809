"""Diagnostics for the general-scene parity tests (GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ray_tracing_weekend_b200 as rtw
from oracle import pyoracle as oracle
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_gpu_general as T

world, lights = T._random_scene(rtw, np.random.default_rng(11))
scene = rtw.Scene(world, lights)
og = oracle.GScene(scene.desc.pod, scene.desc)
cb = rtw.CameraBuilder().with_lookfrom((0., 2., 16.)).with_lookat((0., 0., 0.)).with_focus_dist(16.)
cam = T._cam(cb)
o, d = T._rays(oracle, og, cam.pod, 3000, 7)
prim_o, t_o, _, _ = og.trace_batch(o, d)
prim_g, t_g = scene.trace_batch(o, d, precision=rtw.RTW_F64)
bad = np.nonzero(prim_o != prim_g)[0]
print("mismatches", len(bad), "t mismatches", (t_o != t_g).sum())
def desc(k):
    if k < 0: return "miss"
    e = world.items[k]
    tr = isinstance(e, rtw.Transformed)
    inst = e.instance if tr else e
    return f"{type(inst).__name__}{'+T' if tr else ''}"
from collections import Counter
print(Counter((desc(prim_o[i]), desc(prim_g[i])) for i in bad).most_common(20))
for i in bad[:8]:
    print(i, prim_o[i], prim_g[i], t_o[i], t_g[i], o[i], d[i])

# cornell f32 vs f64 means
scene2, og2, cb2 = T._build(rtw, oracle, "cornell_box")
cam2 = T._cam(cb2, 64, 64, 128, 50)
for prec, name in ((rtw.RTW_F64, "f64"), (rtw.RTW_F32, "f32")):
    for seed in (1, 2, 3):
        img, _, st = scene2.render(cam2, rtw.RenderOptions(seed=seed, precision=prec))
        a = img / 128
        ok = np.isfinite(a).all(axis=2)
        print(name, seed, "clamped mean", np.clip(a[ok], 0, 2).mean(), "finite", ok.mean(), "rays/path", st["rays"] / st["paths"], "ms", st["kernel_ms"])
for tmin in (1e-3,):
    for prec, name in ((rtw.RTW_F64, "f64"), (rtw.RTW_F32, "f32")):
        img, _, st = scene2.render(cam2, rtw.RenderOptions(seed=1, precision=prec, tmin=tmin))
        a = img / 128
        ok = np.isfinite(a).all(axis=2)
        print(name, "tmin", tmin, "clamped mean", np.clip(a[ok], 0, 2).mean(), "finite", ok.mean(), "rays/path", st["rays"] / st["paths"])

print("--- event counts per path, cornell 64x64x128")
for prec, name in ((rtw.RTW_F64, "f64"), (rtw.RTW_F32, "f32")):
    for tmin in (rtw.TMIN_REFERENCE, 1e-3):
        img, _, st = scene2.render(cam2, rtw.RenderOptions(seed=1, precision=prec, tmin=tmin, flags=rtw.RTW_FLAG_COUNT_EVENTS))
        n = st["paths"]
        print(name, tmin, {k: round(st[k] / n, 3) for k in ("rays", "lambertian", "metal", "dielectric", "absorbed", "missed", "depth_out", "light_tests", "node_visits")})
cam3 = T._cam(cb2, 1024, 1024, 64, 50)
for prec, name in ((rtw.RTW_F32, "f32"), (rtw.RTW_F64, "f64")):
    img, _, st = scene2.render(cam3, rtw.RenderOptions(seed=1, precision=prec))
    print("cornell 1024x1024x64", name, "ms", st["kernel_ms"], "Mrays/s", st["rays"] / st["kernel_ms"] / 1e3, "rays/path", st["rays"] / st["paths"])
