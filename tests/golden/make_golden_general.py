"""Golden fixture of the GENERAL-scene oracle (oracle/rtw_oracle_general.hpp) on the reference's scenes cornell_box, simple_light,
debugging_scene, simple_transform and checkered_spheres (scenes/src/lib.rs:123-153, 235-653).  Run from the repo root:
`python tests/golden/make_golden_general.py`.  The reference cannot run here (no Rust toolchain) and has no golden vectors
for this path, so this pins the ORACLE (and through it the CUDA path) against drift.  Scene descriptions come from the
product's host mirror (ray_tracing_weekend_b200.scenes, no GPU involved); all arithmetic is the oracle's, in PORTABLE
math mode (fixed IEEE sequences instead of libm sin/cos) so that the numbers do not depend on the libm build."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

HERE = os.path.dirname(os.path.abspath(__file__))
SEED = 20261018


def compute(O, R):
    out = {}
    for name in ("cornell_box", "simple_light", "debugging_scene", "simple_transform", "checkered_spheres"):
        gen = getattr(R.scenes, name)
        world, lights, cb = gen() if name in ("cornell_box", "checkered_spheres") else gen(SEED)
        d = R.SceneDescription(world, lights)
        g = O.GScene(d.pod, d)
        cam = cb.with_vfov(40.).with_aspect_ratio(1.0).with_image_width(24).with_image_height(24).with_samples_per_pixel(4).with_max_depth(12).build()
        img, _, cnt, _ = g.render(O.Camera.from_buffer_copy(cam.pod), O.options(seed=SEED, math_mode=O.PORTABLE, threads=1))
        out[name] = dict(n_world=d.n_world, n_lights=d.n_lights, rays=cnt["rays"], paths=cnt["paths"],
                         image_sum=np.nan_to_num(img, nan=-1.0).reshape(-1)[::7].tolist())
    # beyond the reference's scenes: get_plane_uv's rotated branch (plane.rs:48-54) and a CheckerTexture of CheckerTextures
    # (texture.rs:26-29, 46-55).  theta / cos / sin of a plane come from libm in every math mode (they are per-plane constants).
    checker = R.Lambertian(R.CheckerTexture.new_with_colours((0.9, 0.1, 0.1), (0.1, 0.1, 0.9), 0.37))
    nested = R.Lambertian(R.CheckerTexture(R.CheckerTexture.new_with_colours((1., 1., 0.), (0., 1., 1.), 0.05), R.NoiseTexture(3.0, SEED, 5), 0.2))
    world, lights = R.HittableList(), R.HittableList()
    for point, normal in (((0., -1., 0.), (0.2, 1., -0.4)), ((0., 4., 0.), (0.5, -0.7, 0.1)), ((-6., 0., 0.), (-1., 0., 0.))):
        world.add(R.Plane(point, normal, checker))
    world.add(R.Sphere((-1.8, 0.6, 1.), 0.9, nested))
    world.add(R.Quad((1.2, -0.5, 1.5), (1.5, 0., 0.3), (0., 1.5, 0.2), nested))
    lights.add(R.Sphere((2., 2., 1.), 0.4, R.DiffuseLight((3., 3., 3.)))); world.add(lights.items[0])
    d = R.SceneDescription(world, lights)
    g = O.GScene(d.pod, d)
    cam = (R.CameraBuilder().with_lookfrom((3., 1.5, 5.)).with_lookat((0., 1., 0.)).with_background((0.6, 0.7, 0.9)).with_vfov(40.).with_aspect_ratio(1.0)
           .with_image_width(24).with_image_height(24).with_samples_per_pixel(4).with_max_depth(12).build())
    img, _, cnt, _ = g.render(O.Camera.from_buffer_copy(cam.pod), O.options(seed=SEED, math_mode=O.PORTABLE, threads=1))
    out["tilted_planes_nested_checker"] = dict(n_world=d.n_world, n_lights=d.n_lights, rays=cnt["rays"], paths=cnt["paths"],
                                               image_sum=np.nan_to_num(img, nan=-1.0).reshape(-1)[::7].tolist())
    return out


if __name__ == "__main__":
    from oracle import pyoracle as O
    import ray_tracing_weekend_b200 as R
    g = compute(O, R)
    with open(os.path.join(HERE, "general_oracle.json"), "w") as f:
        json.dump(g, f)
    print({k: (v["rays"], v["paths"]) for k, v in g.items()})
