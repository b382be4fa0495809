"""A/B measurements of kernel variants in ONE GPU-box visit.

Each variant is (library file, environment): tuning builds come from ray_tracing_weekend_b200.build.build_variant() and run-time
knobs from the environment (RTW_SH_NODE_STRIDE, ...).  Every variant runs in its own process (the knobs are read once per
process), renders the same frame, and reports the best / median kernel time plus a SHA-1 of the resolved image so that
"bit-identical to the baseline" is checked, not assumed.

  python scripts/variant_bench.py [--config C2|C1|C4|C5|cornell_box|simple_light|debugging_scene|simple_transform|checkered_spheres] [--spp N] [--reps K] name[:lib][,ENV=val...] ...
  python scripts/variant_bench.py --child ...        (internal)
"""
import hashlib
import json
import os
import subprocess
import sys

GENERAL = {"cornell_box": (1024, 1024), "simple_light": (1920, 1080), "debugging_scene": (1920, 1080), "simple_transform": (1920, 1080),
           "checkered_spheres": (1920, 1080)}
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SEED = 20261018


def child(config, spp, reps, mode):
    import numpy as np
    import ray_tracing_weekend_b200 as R
    w, h = 1920, 1080
    cam_edit = None
    if config in GENERAL:                       # the reference's other scenes through the general path (scripts/general_configs.py sizes)
        w, h = GENERAL[config]
        gen = getattr(R.scenes, config)
        world, lights, cb = gen() if config in ("cornell_box", "checkered_spheres") else gen(SEED)
        sc = R.Scene(world, lights)
        cam = cb.with_vfov(40.).with_aspect_ratio(w / h).with_max_depth(50).with_image_width(w).with_image_height(h).with_samples_per_pixel(spp).build()
        return run(R, sc, cam, reps, mode)
    if config == "C1":
        arrays, w, h = R.scenes.simple_arrays(SEED), 400, 225
    elif config == "C5":
        arrays = R.scenes.simple_arrays(SEED, 11, 0.1, 0.2)
    elif config == "C4":
        arrays = R.scenes.simple_arrays(SEED, 500)
        cam_edit = lambda cb: cb.with_lookfrom((60., 30., 60.)).with_focus_dist(float(np.linalg.norm([60., 30., 60.])))
    else:
        arrays = R.scenes.simple_arrays(SEED)
    sc = R.Scene.from_arrays(arrays["spheres"], arrays["sphere_materials"], arrays["planes"], arrays["plane_materials"], arrays["lights"])
    cb = arrays["cam"].with_vfov(40.).with_aspect_ratio(w / h).with_max_depth(50).with_image_width(w).with_image_height(h).with_samples_per_pixel(spp)
    if cam_edit:
        cb = cam_edit(cb)
    cam = cb.build()
    return run(R, sc, cam, reps, mode)


def run(R, sc, cam, reps, mode):
    opts = R.RenderOptions(seed=SEED, mode=R.RTW_WAVEFRONT if mode == "wavefront" else R.RTW_MEGAKERNEL)
    times, st, rgb8 = [], None, None
    for k in range(reps + 1):
        _, rgb8, st = sc.render(cam, opts, want_sum=False, want_rgb8=True)
        if k:
            times.append(st["kernel_ms"])
    times.sort()
    out = dict(kernel_ms_best=times[0], kernel_ms_median=times[len(times) // 2], rays=st["rays"], paths=st["paths"],
               mrays_per_s=st["rays"] / times[0] * 1e-3, sha1=hashlib.sha1(rgb8.tobytes()).hexdigest()[:16])
    sc.close()
    print("RESULT " + json.dumps(out), flush=True)


def main():
    args = sys.argv[1:]
    config, spp, reps, mode = "C2", 100, 5, "wavefront"
    variants = []
    is_child = False
    while args:
        a = args.pop(0)
        if a == "--child": is_child = True
        elif a == "--config": config = args.pop(0)
        elif a == "--spp": spp = int(args.pop(0))
        elif a == "--reps": reps = int(args.pop(0))
        elif a == "--mode": mode = args.pop(0)
        else: variants.append(a)
    if is_child:
        return child(config, spp, reps, mode)
    base = None
    for v in variants or ["default"]:
        parts = v.split(",")
        name, _, lib = parts[0].partition(":")
        env = dict(os.environ)
        if lib:
            env["RTW_LIBRARY"] = os.path.join(ROOT, "ray_tracing_weekend_b200", "lib", "variants", f"librtw_cuda_{lib}.so")
        for kv in parts[1:]:
            k, _, val = kv.partition("=")
            env[k] = val
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", "--config", config, "--spp", str(spp), "--reps", str(reps),
                            "--mode", mode], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
        res = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
        if not res:
            print(json.dumps(dict(variant=v, config=config, error=r.stdout[-600:])), flush=True)
            continue
        out = json.loads(res[0][7:])
        if base is None:
            base = out
        out.update(variant=v, config=config, spp=spp, mode=mode, same_image_as_first=out["sha1"] == base["sha1"],
                   speedup_vs_first=base["kernel_ms_best"] / out["kernel_ms_best"])
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
