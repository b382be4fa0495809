"""The reference's other scenes on one GPU through the general path: one JSON line per scene and precision,
plus the oracle (CPU port of the reference) timed on a bounded sample of the same workload.
  cornell_box 1024x1024, 256 spp | simple_light 1920x1080, 256 spp | debugging_scene / simple_transform / checkered_spheres / perlin_spheres_lit 1920x1080, 64 spp
usage: general_configs.py [scene ...] [--once]     (--once: a single FP32 render of the first scene, for ncu)"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ray_tracing_weekend_b200 as R

SEED = 20261018
args = [a for a in sys.argv[1:] if not a.startswith("--")]
once = "--once" in sys.argv
CONFIGS = {"cornell_box": (1024, 1024, 256), "simple_light": (1920, 1080, 256), "debugging_scene": (1920, 1080, 64), "simple_transform": (1920, 1080, 64),
           # row f2 (textures): the reference's checkered_spheres, and its perlin_spheres with the noise sphere ALSO put into the lights list
           # (the reference's own lights list is empty there and its first diffuse bounce panics, hittable_list.rs:414-419)
           "checkered_spheres": (1920, 1080, 64), "perlin_spheres_lit": (1920, 1080, 64)}


def perlin_spheres_lit(seed):
    world, lights, cb = R.scenes.perlin_spheres(seed)
    lights.add(world.items[1])
    return world, lights, cb


which = args or list(CONFIGS)

for name in which:
    w, h, spp = CONFIGS[name]
    gen = perlin_spheres_lit if name == "perlin_spheres_lit" else getattr(R.scenes, name)
    world, lights, cb = gen() if name in ("cornell_box", "checkered_spheres") else gen(SEED)
    sc = R.Scene(world, lights)
    cam = cb.with_vfov(40.).with_aspect_ratio(w / h).with_max_depth(50).with_image_width(w).with_image_height(h).with_samples_per_pixel(spp).build()
    if once and os.environ.get("RTW_GENERAL_SMALL"):
        cam = cb.with_vfov(40.).with_aspect_ratio(w / h).with_max_depth(50).with_image_width(w // 2).with_image_height(h // 2).with_samples_per_pixel(64).build()
    if once:
        _, _, st = sc.render(cam, R.RenderOptions(seed=SEED, precision=R.RTW_F32), want_sum=False, want_rgb8=True)
        print(json.dumps(dict(scene=name, kernel_ms=st["kernel_ms"], rays=st["rays"])))
        break
    for prec, pname in ((R.RTW_F32, "f32"), (R.RTW_F64, "f64")):
        spp_p = spp if prec == R.RTW_F32 else max(8, spp // 8)          # the f64 path is ~8x slower: fewer samples, same metric
        camp = cb.with_vfov(40.).with_aspect_ratio(w / h).with_max_depth(50).with_image_width(w).with_image_height(h).with_samples_per_pixel(spp_p).build()
        best = None
        for _ in range(3):
            _, rgb8, st = sc.render(camp, R.RenderOptions(seed=SEED, precision=prec), want_sum=False, want_rgb8=True)
            if best is None or st["kernel_ms"] < best["kernel_ms"]:
                best = st
        _, _, cnt = sc.render(camp, R.RenderOptions(seed=SEED, precision=prec, flags=R.RTW_FLAG_COUNT_EVENTS), want_sum=False, want_rgb8=False)
        line = dict(scene=name, precision=pname, width=w, height=h, spp=spp_p, entries=sc.desc.n_world, lights=sc.desc.n_lights,
                    kernel_ms=round(best["kernel_ms"], 2), total_ms=round(best["total_ms"], 2),
                    mrays_per_s=round(best["rays"] / best["kernel_ms"] * 1e-3, 1), mpaths_per_s=round(best["paths"] / best["kernel_ms"] * 1e-3, 1),
                    rays_per_path=round(best["rays"] / best["paths"], 3), node_visits_per_ray=round(cnt["node_visits"] / cnt["rays"], 2),
                    black_pixel_fraction=round(float((rgb8 == 0).all(axis=2).mean()), 4))
        print(json.dumps(line), flush=True)
        if prec == R.RTW_F32:
            os.makedirs("gpurun_out", exist_ok=True)
            R.write_ppm(f"gpurun_out/{name}.ppm", rgb8[::4, ::4] if w > 1100 else rgb8[::2, ::2])
    # CPU port of the reference on the host cores, bounded sample: same scene / camera at reduced size
    from oracle import pyoracle as O
    og = O.GScene(sc.desc.pod, sc.desc)
    ws, hs, ss = w // 2, h // 2, 32             # bounded sample: a quarter of the pixels, 32 spp (seconds of CPU work)
    cams = cb.with_vfov(40.).with_aspect_ratio(w / h).with_max_depth(50).with_image_width(ws).with_image_height(hs).with_samples_per_pixel(ss).build()
    _, sec, cnt, _ = og.render(O.Camera.from_buffer_copy(cams.pod), O.options(seed=SEED))
    print(json.dumps(dict(scene=name, precision="cpu-oracle-f64", cores=O.hardware_threads(), sample=f"{ws}x{hs}, {ss} spp", seconds=round(sec, 3),
                          mrays_per_s=round(cnt["rays"] / sec * 1e-6, 2), mpaths_per_s=round(cnt["paths"] / sec * 1e-6, 2),
                          rays_per_path=round(cnt["rays"] / cnt["paths"], 3))), flush=True)
    sc.close()
