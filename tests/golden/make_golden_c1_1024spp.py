"""Golden fixture for north-star check (2): the ORACLE's converged render of BASELINE config C1 (scenes::simple seed 20261018,
400x225, depth 50) at 1024 spp, resolved to 8 bits like Colour::write_colour — two independent Philox seeds, with and without
the (non-reference) fix_nan mode.  Two seeds so that the test can take its tolerance from the fixture itself: the CUDA render
must be as close to seed A as seed B is.  ~4.5 CPU-minutes on 8 cores; run from the repo root:
    python tests/golden/make_golden_c1_1024spp.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as O  # noqa: E402

SEED, W, H, SPP, DEPTH = 20261018, 400, 225, 1024, 50
desc = O.scene_simple(SEED)
sc = O.Scene(desc)
cam = O.camera_for(desc, W, H, SPP, DEPTH)
out = {}
for fix in (True, False):
    for tag, seed in (("a", SEED), ("b", SEED + 1)):
        img, _, cnt, _ = sc.render(cam, O.options(seed=seed, rng_mode=O.W64, fix_nan=fix))
        out[f"{'fix' if fix else 'ref'}_{tag}"] = O.resolve(img, SPP)
        out[f"{'fix' if fix else 'ref'}_{tag}_rays_per_path"] = np.float64(cnt["rays"] / cnt["paths"])
# the same with a ROBUST tmin (1e-3): no self-intersection, so the image no longer depends on the rounding noise of hit points and an
# FP32 renderer can be held to the f64 noise floor
for fix in (True, False):
    for tag, seed in (("a", SEED), ("b", SEED + 1)):
        img, _, cnt, _ = sc.render(cam, O.options(seed=seed, tmin=1e-3, rng_mode=O.W64, fix_nan=fix))
        out[f"robust_{'fix' if fix else 'ref'}_{tag}"] = O.resolve(img, SPP)
        out[f"robust_{'fix' if fix else 'ref'}_{tag}_rays_per_path"] = np.float64(cnt["rays"] / cnt["paths"])
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "c1_1024spp_oracle_rgb8.npz"), **out)
