"""CPU tests of the product's host side: the C-ABI library loads and exports every declared symbol,
host helpers (camera builder, Philox, tile partition, scene generator) agree with the oracle, and compute
entry points fail loudly without a GPU.  No kernel is launched here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SEED = 20261018


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    return sorted(set(re.findall(r"RTW_API[^;(]*?\b(rtwh?_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(rtw):
    names = _declared("rtw.h") + _declared("rtw_host.h")
    assert len(names) >= 34
    lib = C.CDLL(rtw.library_path())
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ but not exported"
    from ray_tracing_weekend_b200 import _lib
    assert sorted(_lib.RTW_SYMBOLS + _lib.RTWH_SYMBOLS) == names
    out = subprocess.run(["nm", "-D", "--defined-only", rtw.library_path()], capture_output=True, text=True).stdout
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T " in l)
    assert exported == names, "the library exports exactly the declared C ABI"
    assert lib.rtw_abi_version() == 3


def test_rust_binding_declares_every_entry_point():
    """rust/cuda cannot be compiled here (no Rust toolchain), so its completeness is checked mechanically: the extern "C" block of
    lib.rs names exactly the functions include/rtw.h declares, and build.rs compiles exactly the units of build.py with its flags."""
    rs = open(os.path.join(ROOT, "rust", "cuda", "src", "lib.rs")).read()
    declared = sorted(set(re.findall(r"pub fn (rtw_[a-z0-9_]+)\s*\(", rs)))
    assert declared == _declared("rtw.h")
    from ray_tracing_weekend_b200 import build as B
    brs = open(os.path.join(ROOT, "rust", "cuda", "build.rs")).read()
    units = re.findall(r'\((?:csrc|host)\.join\("([a-z0-9_]+\.(?:cu|cpp))"\), &\[([^\]]*)\]\)', brs)
    assert [(u, [f.strip().strip('"') for f in fl.split(",") if f.strip()]) for u, fl in units] == [(os.path.basename(u), f) for u, f in B.UNITS]
    for flag in B.COMMON + B.ARCH:
        assert f'"{flag}"' in brs, flag
    assert '"-cudart", "static"' in brs and '"-ldl"' in brs
    # INTEGRATION.md quotes the same unit list
    integ = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    for u, _ in B.UNITS:
        assert os.path.basename(u) in integ, u


def test_library_contains_sm100a_kernels(rtw):
    out = subprocess.run(["cuobjdump", "-lelf", rtw.library_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_philox_host_matches_kat_and_oracle(rtw, oracle):
    assert [int(x) for x in rtw.philox4x32_10([0, 0, 0, 0], [0, 0])] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    rng = np.random.default_rng(0)
    for _ in range(50):
        ctr = rng.integers(0, 2 ** 32, 4, dtype=np.uint64).astype(np.uint32); key = rng.integers(0, 2 ** 32, 2, dtype=np.uint64).astype(np.uint32)
        assert np.array_equal(rtw.philox4x32_10(ctr, key), oracle.philox4x32_10(ctr, key))


def test_scene_generator_matches_oracle(rtw, oracle):
    for args in ((SEED, 11, 0.8, 0.95, 0), (7, 5, 0.1, 0.2, 1), (9, 3, 0.5, 0.6, 2)):
        world, lights, cb = rtw.scenes.simple(*args)
        d = oracle.scene_simple(*args)
        sp = np.array([[*s.center, s.radius] for s in world.list.spheres])
        assert np.array_equal(sp, d.spheres)
        assert np.array_equal(np.array([[*s.center, s.radius] for s in lights.spheres]).reshape(-1, 4), d.lights)
        mats = np.array([[s.material.kind, *s.material.colour, s.material.param] for s in world.list.spheres])
        assert np.array_equal(mats, d.materials_array()[d.sphere_mat])
        assert len(world.list.planes) == d.planes.shape[0]
    world, lights, _ = rtw.scenes.simple(SEED)
    kinds = np.bincount([s.material.kind for s in world.list.spheres], minlength=3)
    assert kinds[0] > 5 * kinds[2] and lights.len() == kinds[2]      # every glass sphere is mirrored in `lights` (lib.rs:203,217)


def test_camera_builder_matches_oracle(rtw, oracle):
    d = oracle.scene_simple(SEED)
    _, _, cb = rtw.scenes.simple(SEED)
    for (w, h) in ((400, 225), (1920, 1080), (3, 2)):
        cam = cb.with_vfov(40.).with_aspect_ratio(w / h).with_max_depth(50).with_image_width(w).with_image_height(h).with_samples_per_pixel(7).build()
        oc = oracle.camera_for(d, w, h, 7, 50)
        for f, g in (("center", "center"), ("pixel00_loc", "pixel00"), ("pixel_delta_u", "du"), ("pixel_delta_v", "dv"),
                     ("defocus_disk_u", "ddu"), ("defocus_disk_v", "ddv"), ("background", "background")):
            assert list(getattr(cam.pod, f)) == list(getattr(oc, g)), f
        assert (cam.pod.image_width, cam.pod.image_height, cam.pod.samples_per_pixel, cam.pod.max_depth) == (w, h, 7, 50)
    # the Option<> resolution table (camera.rs:130-156)
    assert (rtw.CameraBuilder().build().image_width, rtw.CameraBuilder().build().image_height) == (100, 100)
    c = rtw.CameraBuilder().with_aspect_ratio(16 / 9).build()
    assert (c.image_width, c.image_height) == (100, 56)
    c = rtw.CameraBuilder().with_aspect_ratio(2.0).with_image_height(30).build()
    assert (c.image_width, c.image_height) == (60, 30)
    c = rtw.CameraBuilder().with_image_width(33).build()
    assert (c.image_width, c.image_height) == (33, 33)


def test_tile_partition_helpers(rtw):
    from ray_tracing_weekend_b200 import dist as D
    for (w, h, world) in ((1920, 1080, 8), (400, 225, 3), (3, 2, 2), (16, 16, 4), (100, 70, 1)):
        assert rtw.tiles_total(w, h) == D.tiles_total(w, h)
        assert rtw.tiles_per_rank(w, h, world) == D.tiles_per_rank(w, h, world)
        seen = []
        for r in range(world):
            ids = D.local_tile_ids(w, h, r, world)
            assert len(ids) <= D.tiles_per_rank(w, h, world)
            assert all(D.tile_owner(k, world) == (r, n) for n, k in enumerate(ids))
            seen += ids
        assert sorted(seen) == list(range(D.tiles_total(w, h)))
        cover = np.zeros((h, w), dtype=int)
        for k in range(D.tiles_total(w, h)):
            i0, j0, i1, j1 = D.tile_rect(k, w, h)
            cover[j0:j1, i0:i1] += 1
        assert (cover == 1).all()


def test_argument_validation_needs_no_gpu(rtw):
    from ray_tracing_weekend_b200 import _lib
    L = rtw.load()
    # a Lambertian world with an empty lights list: rtw_scene_create (sphere path) refuses it — the reference panics on the first
    # light sample (hittable_list.rs:414-419); the general path accepts it and reports the panic per render (tests/test_gpu_general.py)
    world = rtw.HittableList(); world.add(rtw.Sphere((0, 0, 0), 1.0, rtw.Lambertian((0.5, 0.5, 0.5))))
    with pytest.raises(rtw.RtwError) as e:
        rtw.Scene(world, rtw.HittableList(), general=False)
    assert e.value.code == _lib.RTW_E_INVALID and "lights" in str(e.value)
    world = rtw.HittableList(); world.add(rtw.Sphere((0, 0, 0), -1.0, rtw.Metal((1, 1, 1), 0.0)))
    with pytest.raises(rtw.RtwError) as e:
        rtw.Scene(world, rtw.HittableList())
    assert e.value.code == _lib.RTW_E_INVALID
    out = C.c_void_p()
    assert L.rtw_scene_create(None, None, 1, None, None, 0, None, 0, None, 0, C.byref(out)) == _lib.RTW_E_INVALID
    assert L.rtw_camera_build(None, None) == _lib.RTW_E_INVALID
    n = C.c_size_t(0)
    assert L.rtw_scene_export_bvh(None, None, 0, C.byref(n), None, 0, C.byref(n)) == _lib.RTW_E_INVALID
    assert C.sizeof(C.c_double) * 6 + 6 * 4 == rtw.Scene.BVH_NODE_DTYPE.itemsize == 72        # rtw_bvh_node (include/rtw.h)
    with pytest.raises(TypeError):
        rtw.HittableList().add("quad")


def test_no_cpu_fallback(rtw, simple_scene):
    """Without a CUDA device every compute call must fail loudly (RTW_E_NO_DEVICE), never compute on the CPU."""
    from ray_tracing_weekend_b200 import _lib
    if rtw.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(rtw.RtwError) as e:
        rtw.Scene(simple_scene["world"], simple_scene["lights"])
    assert e.value.code == _lib.RTW_E_NO_DEVICE and "no CPU fallback" in str(e.value)
    cam = rtw.CameraBuilder().with_image_width(4).with_image_height(4).build()
    with pytest.raises(rtw.RtwError) as e:
        cam.get_rays([0], [0], [0])
    assert e.value.code == _lib.RTW_E_NO_DEVICE
    r = subprocess.run([os.path.join(os.path.dirname(rtw.library_path()), "rtw_bin"), "simple", "--width", "4", "--height", "4", "--spp", "1"],
                       capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under the package may import, link or execute it."""
    pkg = os.path.join(ROOT, "ray_tracing_weekend_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) in ("build", "lib", "__pycache__"):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "pyoracle" not in txt and "liboracle" not in txt and "rtw_oracle" not in txt, os.path.join(dirpath, f)


def test_cpp_host_writers_and_config(rtw, tmp_path):
    """The C++ mirror's P3 / P6 / PNG writers (bin/src/main.rs:89-104 writes P3, rows reversed) and its Config.toml reader
    (bin/src/config.rs: two of aspect_ratio / image_width / image_height, truncating casts) — compiled and run on the CPU."""
    lib_dir = os.path.dirname(rtw.library_path())
    exe = tmp_path / "writers_and_config"
    src = os.path.join(ROOT, "tests", "host_cpp", "writers_and_config.cpp")
    subprocess.run(["g++", "-O1", "-std=c++17", src, "-o", str(exe), "-L" + lib_dir, "-lrtw_cuda", "-Wl,-rpath," + lib_dir], check=True)
    cfg = tmp_path / "Config.toml"
    cfg.write_text("# [image]\n# image_width = 1200\n\n[image]\naspect_ratio = 1.5   # comment\nimage_width = 1_000\nsamples_per_pixel = 500\nmax_depth = 75\n")
    out = subprocess.run([str(exe), str(tmp_path), str(cfg)], capture_output=True, text=True, check=True).stdout.split()
    assert (float(out[0]), int(out[1]), int(out[2]), int(out[3]), int(out[4])) == (1.5, 1000, 666, 500, 75)     # (1000 / 1.5) as u32
    cfg.write_text("[image]\nimage_width = 400\nimage_height = 400\nsamples_per_pixel = 1000\nmax_depth = 50\n")  # the reference's own Config.toml
    out = subprocess.run([str(exe), str(tmp_path), str(cfg)], capture_output=True, text=True, check=True).stdout.split()
    assert (float(out[0]), int(out[1]), int(out[2]), int(out[3]), int(out[4])) == (1.0, 400, 400, 1000, 50)
    cfg.write_text("[image]\nimage_width = 400\nsamples_per_pixel = 10\nmax_depth = 5\n")
    assert "error" in subprocess.run([str(exe), str(tmp_path), str(cfg)], capture_output=True, text=True, check=True).stdout
    want = np.zeros((3, 5, 3), dtype=np.uint8)                  # top-down: file row 0 = render row 2
    for j in range(3):
        for i in range(5):
            want[2 - j, i] = (10 * i, 100 + j, i * j + 7)
    p3 = (tmp_path / "a.ppm").read_text().split()
    assert p3[:4] == ["P3", "5", "3", "255"] and np.array_equal(np.array(p3[4:], dtype=np.uint8).reshape(3, 5, 3), want)
    p6 = (tmp_path / "b.ppm").read_bytes()
    assert p6.startswith(b"P6\n5 3\n255\n") and np.array_equal(np.frombuffer(p6[len(b"P6\n5 3\n255\n"):], dtype=np.uint8).reshape(3, 5, 3), want)
    Image = pytest.importorskip("PIL.Image")
    assert np.array_equal(np.asarray(Image.open(tmp_path / "c.png").convert("RGB")), want)


@pytest.fixture(scope="module")
def dump_scenes_exe(rtw, tmp_path_factory):
    lib_dir = os.path.dirname(rtw.library_path())
    exe = tmp_path_factory.mktemp("dump_scenes") / "dump_scenes"
    src = os.path.join(ROOT, "tests", "host_cpp", "dump_scenes.cpp")
    subprocess.run(["g++", "-O1", "-std=c++17", src, "-o", str(exe), "-L" + lib_dir, "-lrtw_cuda", "-Wl,-rpath," + lib_dir], check=True)
    return exe


@pytest.mark.parametrize("name", ["simple_light", "cornell_box", "debugging_scene", "simple_transform", "checkered_spheres", "plane"])
def test_scene_mirrors_agree_general(rtw, tmp_path, dump_scenes_exe, name):
    """The scene constants live in hand-kept mirrors (C++ host/rtw_host.hpp for `rtw_bin`, Python scenes.py for the tests; the oracle's
    general scenes are built FROM the Python description, tests/test_gpu_general.py).  Every general scene of the CLI must come out
    of both generators as the same plain-data description, byte for byte: entity arrays, transforms, materials, textures, Perlin
    tables, the two entry lists, the is-BVH flags and the camera builder."""
    import ctypes as C
    exe = dump_scenes_exe
    out = tmp_path / (name + ".bin")
    subprocess.run([str(exe), name, str(out)], check=True)
    blob = out.read_bytes()
    gen = getattr(rtw.scenes, name)
    world, lights, cb = gen(SEED) if name in ("simple_light", "debugging_scene", "simple_transform") else gen()
    desc = rtw.SceneDescription(world, lights)
    pos = 0
    for key in ("spheres", "planes", "quads", "cuboids", "transforms", "materials", "textures", "perlins", "world", "lights"):
        n = int.from_bytes(blob[pos:pos + 8], "little"); pos += 8
        assert n == getattr(desc.pod, "n_" + key), (name, key, n)
        arr = desc._keep[key]
        size = C.sizeof(arr._type_) * n
        assert blob[pos:pos + size] == bytes(arr)[:size], (name, key)
        pos += size
    assert np.frombuffer(blob[pos:pos + 8], dtype=np.uint32).tolist() == [desc.pod.world_is_bvh, desc.pod.lights_is_bvh]
    pos += 8
    assert blob[pos:] == bytes(cb.pod), (name, "camera builder")
