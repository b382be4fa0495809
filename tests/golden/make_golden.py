"""Generates the golden fixtures in this directory from the oracle (run from the repo root:
`python tests/golden/make_golden.py`).  The reference itself cannot run here (no Rust toolchain) and has
no golden vectors of its own, so these pin the ORACLE (and through it the CUDA path) against drift."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SEED = 20261018

desc = O.scene_simple(SEED)
sc = O.Scene(desc)
cam = O.camera_for(desc, 64, 36, 4, 50)
g = dict(seed=SEED, n_spheres=int(desc.spheres.shape[0]), n_lights=int(desc.lights.shape[0]),
         first_spheres=desc.spheres[:8].tolist(), bvh_stats=sc.bvh_stats(), rays={})
for mode, key in ((O.LIBM, "image_sum_libm"), (O.PORTABLE, "image_sum_portable")):
    img, _, cnt, _ = sc.render(cam, O.options(seed=SEED, math_mode=mode))
    g[key] = np.nan_to_num(img, nan=-1.0).reshape(-1)[::97].tolist()
    g["rays"][key] = cnt["rays"]
with open(os.path.join(HERE, "simple_seed20261018.json"), "w") as f:
    json.dump(g, f)

rng = np.random.default_rng(42)
n = 2000
i = rng.integers(0, 64, n); j = rng.integers(0, 36, n); s = rng.integers(0, 4, n)
o, d = O.get_rays(cam, O.options(seed=SEED), i, j, s)
prim, t, _ = sc.trace_batch(o, d)
hit = prim >= 0
p = o[hit] + d[hit] * t[hit][:, None]
k = rng.integers(0, len(p), n)
o = np.concatenate([o, p[k]]); d = np.concatenate([d, rng.normal(size=(n, 3))])
prim, t, _ = sc.trace_batch(o, d)
np.savez_compressed(os.path.join(HERE, "trace_batch_seed20261018.npz"), o=o, d=d, prim=prim, t=t)
print("golden written:", (prim >= 0).sum(), "hits of", len(prim))
