"""GPU parity tests of the general-scene path (Quad, Triangle, Cuboid, Transformed<T>, DiffuseLight,
Isotropic, NoiseTexture, quad lights) through the C ABI (rtw_scene_create_general + the same render / batch entry points)
against the general oracle on the same scene description and the same Philox streams.

f64 path: bit-exact (hits, vertices, path radiance, whole images).  f32 path: statistically (mean radiance, PSNR)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SEED = 20261018
SCENES = ("cornell_box", "simple_light", "debugging_scene", "simple_transform", "checkered_spheres")


def _build(rtw, oracle, name):
    gen = getattr(rtw.scenes, name)
    world, lights, cb = gen() if name in ("cornell_box", "checkered_spheres") else gen(SEED)
    scene = rtw.Scene(world, lights)
    assert scene.general
    return scene, oracle.GScene(scene.desc.pod, scene.desc), cb


def _random_scene(rtw, rng, n=160):
    """A mixed bag: every entity kind, half of them Transformed, every material kind, a NoiseTexture."""
    noise = rtw.NoiseTexture(3.0, SEED, 2)
    mats = [rtw.Lambertian((0.7, 0.6, 0.5)), rtw.Metal((0.8, 0.8, 0.9), 0.1), rtw.Dialectric(1.5), rtw.DiffuseLight((2., 2., 2.)),
            rtw.Isotropic((0.4, 0.5, 0.6)), rtw.Lambertian(noise), rtw.INVISIBLE,
            rtw.Lambertian(rtw.CheckerTexture.new_with_colours((0.9, 0.1, 0.1), (0.1, 0.1, 0.9), 0.13)),
            rtw.DiffuseLight(rtw.CheckerTexture((1.5, 1.5, 0.5), rtw.NoiseTexture(2.0, SEED, 3), 0.21))]
    world, lights = rtw.HittableList(), rtw.HittableList()
    world.add(rtw.Plane((0., -6., 0.), (0., -1., 0.), mats[0]))
    world.add(rtw.Plane((0., 0., 0.), (0., 1., 0.), mats[7]))              # a checkered +y plane (uv = (x, z)), visible from below
    world.add(rtw.Plane((0., -5.5, 0.), (0.3, -1., 0.2), mats[7]))         # a tilted checkered plane: get_plane_uv's rotated branch
    for k in range(n):
        c = rng.uniform(-5, 5, 3)
        m = mats[int(rng.integers(0, len(mats)))]
        kind = k % 4
        if kind == 0:
            e = rtw.Sphere(tuple(c), float(rng.uniform(0.2, 0.8)), m)
        elif kind == 1:
            e = rtw.Quad(tuple(c), tuple(rng.normal(size=3)), tuple(rng.normal(size=3)), m)
        elif kind == 2:
            e = rtw.Triangle(tuple(c), tuple(rng.normal(size=3)), tuple(rng.normal(size=3)), m)
        else:
            e = rtw.Cuboid(tuple(c), tuple(c + rng.uniform(0.2, 1.2, 3)), m)
        if rng.random() < 0.5:
            # rotations only or rotation + tiny translation: a large translation sends the instance ray far away (the
            # reference adds it to the direction), which would leave nothing to compare
            e = e.transform(rtw.rotation(float(rng.uniform(-90, 90)), int(rng.integers(0, 3))))
            if rng.random() < 0.5:
                e = e.transform(rtw.Translation3(*rng.normal(size=3) * 0.05))
        world.add(e)
    lights.add(rtw.Quad((-1., 7., -1.), (2., 0., 0.), (0., 0., 2.), mats[3]))
    lights.add(rtw.Sphere((3., 6., 0.), 0.7, mats[3]))
    lights.add(rtw.Triangle((-4., 6., 2.), (1.5, 0., 0.), (0., 0., 1.5), mats[3]))
    lights.add(rtw.Cuboid((0., 0., 0.), (1., 1., 1.), mats[3]).transform(rtw.Translation3(0., 8., 0.)))
    for l in lights.items[:3]:
        world.add(l)
    return world, lights


def _rays(oracle, og, cam_pod, n, seed):
    rng = np.random.default_rng(seed)
    ocam = oracle.Camera.from_buffer_copy(cam_pod)
    i = rng.integers(0, ocam.width, n); j = rng.integers(0, ocam.height, n); s = rng.integers(0, 4, n)
    o, d = oracle.get_rays(ocam, oracle.options(seed=SEED), i, j, s)
    prim, t, p, _ = og.trace_batch(o, d)
    hit = prim >= 0
    if hit.sum() > 0:
        k = rng.integers(0, hit.sum(), n)
        o = np.concatenate([o, p[hit][k]]); d = np.concatenate([d, rng.normal(size=(n, 3))])
    return o, d


def _cam(cb, w=48, h=48, spp=4, depth=12):
    return cb.with_vfov(40.).with_aspect_ratio(w / h).with_image_width(w).with_image_height(h).with_samples_per_pixel(spp).with_max_depth(depth).build()


@pytest.mark.parametrize("name", SCENES + ("random",))
def test_general_hits_and_vertices_f64_bit_exact(rtw, oracle, name):
    if name == "random":
        world, lights = _random_scene(rtw, np.random.default_rng(11))
        scene = rtw.Scene(world, lights)
        og = oracle.GScene(scene.desc.pod, scene.desc)
        cb = rtw.CameraBuilder().with_lookfrom((0., 2., 16.)).with_lookat((0., 0., 0.)).with_focus_dist(16.)
    else:
        scene, og, cb = _build(rtw, oracle, name)
    try:
        cam = _cam(cb)
        o, d = _rays(oracle, og, cam.pod, 3000, 7)
        for tmin in (oracle.EPS, 1e-3):
            prim_o, t_o, _, _ = og.trace_batch(o, d, tmin=tmin)
            prim_g, t_g = scene.trace_batch(o, d, tmin=tmin, precision=rtw.RTW_F64)
            assert np.array_equal(prim_o, prim_g), f"{(prim_o != prim_g).sum()} id mismatches"
            assert np.array_equal(t_o, t_g)
        assert (prim_o >= 0).sum() > 200
        if name == "random":
            assert len(set(prim_o[prim_o >= 0])) > 60                 # many different entries are actually hit
        rng = np.random.default_rng(9)
        pixel = rng.integers(0, 90000, len(o)); sample = rng.integers(0, 100, len(o)); vertex = rng.integers(1, 51, len(o))
        ref = og.scatter_batch(o, d, pixel, sample, vertex, oracle.options(seed=SEED, math_mode=oracle.PORTABLE))
        got = scene.scatter_batch(o, d, pixel, sample, vertex, rtw.RenderOptions(seed=SEED, precision=rtw.RTW_F64))
        assert np.array_equal(ref["prim"], got["prim"]) and np.array_equal(ref["kind"], got["kind"])
        for k in ("t", "p", "normal", "dir", "weight"):
            assert np.array_equal(ref[k], got[k], equal_nan=True), k
    finally:
        scene.close()


@pytest.mark.parametrize("name", SCENES)
def test_general_image_f64_bit_exact(rtw, oracle, name):
    scene, og, cb = _build(rtw, oracle, name)
    try:
        cam = _cam(cb, 40, 30, 6, 20)
        ocam = oracle.Camera.from_buffer_copy(cam.pod)
        opts = oracle.options(seed=SEED, math_mode=oracle.PORTABLE)
        ref, _, cnt, _ = og.render(ocam, opts)
        got, rgb8, st = scene.render(cam, rtw.RenderOptions(seed=SEED, precision=rtw.RTW_F64, flags=rtw.RTW_FLAG_COUNT_EVENTS))
        assert np.array_equal(ref, got, equal_nan=True), f"{name}: {(~np.isclose(ref, got, equal_nan=True)).sum()} values differ"
        assert st["rays"] == cnt["rays"] and st["paths"] == cnt["paths"] == 40 * 30 * 6
        assert np.array_equal(rgb8, oracle.resolve(ref, 6))
        rng = np.random.default_rng(2)
        i = rng.integers(0, 40, 500); j = rng.integers(0, 30, 500); s = rng.integers(0, 6, 500)
        assert np.array_equal(og.path_radiance(ocam, opts, i, j, s),
                              scene.path_radiance(cam, rtw.RenderOptions(seed=SEED, precision=rtw.RTW_F64), i, j, s), equal_nan=True)
    finally:
        scene.close()


def test_general_path_equals_sphere_path_on_simple(rtw, oracle, simple_scene):
    """scenes::simple through rtw_scene_create_general == through rtw_scene_create (f64: bit-identical image)."""
    cam = (simple_scene["cb"].with_vfov(40.).with_aspect_ratio(16 / 9).with_image_width(64).with_image_height(36).with_samples_per_pixel(4)
           .with_max_depth(50).build())
    a = rtw.Scene(simple_scene["world"], simple_scene["lights"])
    b = rtw.Scene(simple_scene["world"], simple_scene["lights"], general=True)
    try:
        assert not a.general and b.general
        opts = rtw.RenderOptions(seed=SEED, precision=rtw.RTW_F64)
        ia, _, sa = a.render(cam, opts)
        ib, _, sb = b.render(cam, opts)
        assert np.array_equal(ia, ib, equal_nan=True) and sa["rays"] == sb["rays"]
    finally:
        a.close(); b.close()


def test_cornell_box_f32_image_statistics(rtw, oracle):
    """FP32 general path vs the f64 oracle with independent noise: same mean radiance, images agree to render noise."""
    scene, og, cb = _build(rtw, oracle, "cornell_box")
    try:
        w = h = 64
        spp = 128
        cam = _cam(cb, w, h, spp, 50)
        ocam = oracle.Camera.from_buffer_copy(cam.pod)
        ref, _, cnt, _ = og.render(ocam, oracle.options(seed=SEED + 1, math_mode=oracle.PORTABLE))
        got, _, st = scene.render(cam, rtw.RenderOptions(seed=SEED, precision=rtw.RTW_F32))
        ref64b, _, _ = scene.render(cam, rtw.RenderOptions(seed=SEED + 2, precision=rtw.RTW_F64))
        a, b, c = ref / spp, got / spp, ref64b / spp
        ok = np.isfinite(a).all(axis=2) & np.isfinite(b).all(axis=2) & np.isfinite(c).all(axis=2)
        assert ok.mean() > 0.9
        # fireflies (15x light through the glass sphere) dominate the raw mean: compare clamped radiance
        ca, cb_, cc = np.clip(a[ok], 0, 2), np.clip(b[ok], 0, 2), np.clip(c[ok], 0, 2)
        assert abs(cb_.mean() - ca.mean()) < 0.025 * ca.mean(), (ca.mean(), cb_.mean())
        mse_noise = ((ca - cc) ** 2).mean()            # two f64 renders with different seeds
        mse_f32 = ((ca - cb_) ** 2).mean()
        assert mse_f32 < 1.5 * mse_noise + 1e-4, (mse_f32, mse_noise)
        # rays per path: equal once tmin is above the rounding noise (5.264 vs 5.265 measured).  At the reference's tmin the
        # count of self-intersections on the r = 90 glass sphere depends on the rounding behaviour of the root formula
        # (f64 textbook quadratic: 8.2 rays / path, FP32 cancellation-free form: 11.7) while the radiance agrees (weights are 1).
        r64 = scene.render(cam, rtw.RenderOptions(seed=SEED, precision=rtw.RTW_F64, tmin=1e-3), want_sum=False, want_rgb8=False)[2]
        r32 = scene.render(cam, rtw.RenderOptions(seed=SEED, precision=rtw.RTW_F32, tmin=1e-3), want_sum=False, want_rgb8=False)[2]
        assert abs(r32["rays"] / r32["paths"] - r64["rays"] / r64["paths"]) < 0.01 * r64["rays"] / r64["paths"]
        assert abs(cnt["rays"] / cnt["paths"] - 8.2) < 0.3
    finally:
        scene.close()


def test_general_scene_validation(rtw):
    from ray_tracing_weekend_b200 import _lib
    white = rtw.Lambertian((0.5, 0.5, 0.5))
    world = rtw.HittableList(); world.add(rtw.Quad((0., 0., 0.), (1., 0., 0.), (0., 1., 0.), white))
    lights = rtw.HittableList(); lights.extend(rtw.Sphere((float(k), 5., 0.), 0.3, rtw.INVISIBLE) for k in range(6))
    with pytest.raises(rtw.RtwError) as e:                          # a BVH of > 5 lights: aux_random is broken in the reference
        rtw.Scene(world, rtw.BoundedVolumeHierarchy.from_list(lights))
    assert e.value.code == _lib.RTW_E_UNSUPPORTED
    # an empty world renders the background
    empty = rtw.Scene(rtw.HittableList(), lights, general=True)
    cam = rtw.CameraBuilder().with_image_width(5).with_image_height(3).with_samples_per_pixel(2).with_background((0.25, 0.5, 1.)).build()
    img, _, st = empty.render(cam, rtw.RenderOptions(seed=1, precision=rtw.RTW_F64))
    assert np.array_equal(img, np.broadcast_to(np.array([0.5, 1., 2.]), (3, 5, 3))) and st["rays"] == 30
    empty.close()


@pytest.mark.parametrize("name,cli", (("cornell_box", "cornell-box"), ("simple_light", "simple-light"), ("simple_transform", "simple-transform"),
                                      ("checkered_spheres", "checkered-spheres")))
def test_cpp_host_mirror_cli_renders_general_scenes(rtw, oracle, tmp_path, name, cli):
    """`rtw_bin cornell-box --backend cuda` (bin/src/main.rs flow through the C++ mirror of Quad / Cuboid / Transformed /
    DiffuseLight / NoiseTexture) writes the same PPM, byte for byte, as the Python mirror."""
    import os, subprocess
    exe = os.path.join(os.path.dirname(rtw.library_path()), "rtw_bin")
    out = tmp_path / "image.ppm"
    w, h, spp = 40, 30, 8
    r = subprocess.run([exe, cli, "--backend", "cuda", "--width", str(w), "--height", str(h), "--spp", str(spp), "--depth", "20",
                        "--seed", str(SEED), "--precision", "f64", "--out", str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    scene, og, cb = _build(rtw, oracle, name)
    try:
        cam = _cam(cb, w, h, spp, 20)
        _, rgb8, _ = scene.render(cam, rtw.RenderOptions(seed=SEED, precision=rtw.RTW_F64))
    finally:
        scene.close()
    ref = tmp_path / "ref.ppm"
    rtw.write_ppm(str(ref), rgb8)
    assert out.read_text() == ref.read_text()


@pytest.mark.parametrize("precision", ("f32", "f64"))
def test_general_render_independent_of_world_size(rtw, oracle, precision):
    """The multi-GPU building blocks on a general scene: tiles rendered as 1 rank and as 3 ranks give the same image bit for bit
    (streams are keyed by absolute pixel; FP32 accumulates in fixed point, f64 sums samples in order)."""
    import torch
    scene, og, cb = _build(rtw, oracle, "cornell_box")
    try:
        w, h, spp = 70, 50, 6
        cam = _cam(cb, w, h, spp, 30)
        prec = rtw.RTW_F32 if precision == "f32" else rtw.RTW_F64
        dtype = torch.float32 if precision == "f32" else torch.float64
        opts = rtw.RenderOptions(seed=SEED, precision=prec)
        one, _, _ = scene.render(cam, opts)
        world = 3
        tpr = rtw.tiles_per_rank(w, h, world)
        tiles = torch.zeros((world, tpr, 16, 16, 3), dtype=dtype, device="cuda")
        for r in range(world):
            scene.render_tiles_device(cam, opts, r, world, tiles[r].data_ptr())
        out = torch.zeros((h, w, 3), dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        rtw.untile_resolve_device(tiles.data_ptr(), prec, w, h, world, spp, out.data_ptr(), 0)
        torch.cuda.synchronize()
        assert np.array_equal(one, out.cpu().numpy(), equal_nan=True)
    finally:
        scene.close()


def test_checkered_spheres_texture_through_fix_nan(rtw, oracle):
    """checkered_spheres as the reference builds it poisons every sphere pixel (its light sits where the two spheres touch, so
    light-sampled paths end up inside the light: Sphere::pdf_value is NaN, sphere.rs:104-106).  With RTW_FLAG_FIX_NAN the checker
    pattern read through get_sphere_uv is visible: still bit-exact against the oracle's fix_nan mode, and both colours appear."""
    scene, og, cb = _build(rtw, oracle, "checkered_spheres")
    try:
        cam = _cam(cb, 60, 40, 8, 20)
        ocam = oracle.Camera.from_buffer_copy(cam.pod)
        ref, _, _, _ = og.render(ocam, oracle.options(seed=SEED, math_mode=oracle.PORTABLE, fix_nan=True))
        got, _, _ = scene.render(cam, rtw.RenderOptions(seed=SEED, precision=rtw.RTW_F64, flags=rtw.RTW_FLAG_FIX_NAN))
        assert np.array_equal(ref, got) and np.isfinite(got).all()
        plain, _, _ = scene.render(cam, rtw.RenderOptions(seed=SEED, precision=rtw.RTW_F64))
        assert np.isnan(plain).any(axis=2).mean() > 0.3
        cam_hi = _cam(cb, 60, 40, 128, 20)                            # FP32 vs f64 with independent noise: more samples
        f64_hi, _, _ = scene.render(cam_hi, rtw.RenderOptions(seed=SEED, precision=rtw.RTW_F64, flags=rtw.RTW_FLAG_FIX_NAN))
        f32_hi, _, _ = scene.render(cam_hi, rtw.RenderOptions(seed=SEED, precision=rtw.RTW_F32, flags=rtw.RTW_FLAG_FIX_NAN))
        sph = (f64_hi.sum(axis=2) < 0.98 * 128 * 3)                   # pixels that are not pure background
        assert sph.mean() > 0.3 and abs(f32_hi[sph].mean() - f64_hi[sph].mean()) < 0.05 * f64_hi[sph].mean()
        # first-vertex albedos: green-ish (0.2, 0.3, 0.1) and white (0.9, 0.9, 0.9) squares both occur
        ii, jj = np.meshgrid(np.arange(60), np.arange(40))
        o, d = cam.get_rays(ii.ravel(), jj.ravel(), np.zeros(2400), rtw.RenderOptions(seed=SEED, precision=rtw.RTW_F64))
        v = scene.scatter_batch(o, d, ii.ravel(), np.zeros(2400), np.ones(2400), rtw.RenderOptions(seed=SEED, precision=rtw.RTW_F64))
        w = v["weight"][(v["kind"] == 3) & (v["weight"].sum(axis=1) > 0)]
        ratio = w[:, 1] / w[:, 0]
        assert (np.abs(ratio - 1.0) < 1e-9).sum() > 50 and (np.abs(ratio - 1.5) < 1e-9).sum() > 50
    finally:
        scene.close()


REFERENCE_INTEGRATION_TESTS = {
    # integration-tests/src/lib.rs: scene generator + the CameraBuilder calls of each #[test] (they only assert "does not panic")
    "plane_test": ("plane", lambda cb, rtw: rtw.CameraBuilder().with_image_width(3).with_image_height(2).with_samples_per_pixel(10).with_max_depth(3)
                   .with_lookfrom((-13., 2., 3.)).with_lookat((0., 0., 0.)).with_vup((0., 1., 0.)).with_focus_dist(10.)),
    "small_light_test": ("simple_light", lambda cb, rtw: rtw.CameraBuilder().with_image_width(3).with_image_height(2).with_samples_per_pixel(10)
                         .with_max_depth(5).with_lookfrom((4., 2., 10.)).with_lookat((4., 2., -2.)).with_vup((0., 1., 0.)).with_focus_dist(4.)),
    "debugging_test": ("debugging_scene", lambda cb, rtw: cb.with_image_width(3).with_image_height(2).with_samples_per_pixel(50).with_max_depth(10)
                       .with_vfov(40.).with_lookat((0., 0., 0.)).with_lookfrom((0., 20., 0.))),
    "cornell_box_test": ("cornell_box", lambda cb, rtw: cb.with_image_width(3).with_image_height(2).with_samples_per_pixel(50).with_max_depth(10).with_vfov(40.)),
}


@pytest.mark.parametrize("name", sorted(REFERENCE_INTEGRATION_TESTS))
def test_reference_integration_tests(rtw, oracle, name):
    """The reference's own integration tests (plane_test, small_light_test, debugging_test, cornell_box_test; small_test is in
    test_gpu_parity.py) with their exact cameras: they must run without the panic the reference maps to an error here, and the
    f64 image equals the oracle's bit for bit."""
    scene_name, make_cam = REFERENCE_INTEGRATION_TESTS[name]
    gen = getattr(rtw.scenes, scene_name)
    world, lights, cb = gen(SEED) if scene_name in ("simple_light", "debugging_scene") else gen()
    cam = make_cam(cb, rtw).build()
    scene = rtw.Scene(world, lights)
    try:
        og = oracle.GScene(scene.desc.pod, scene.desc)
        ref, _, cnt, panicked = og.render(oracle.Camera.from_buffer_copy(cam.pod), oracle.options(seed=SEED, math_mode=oracle.PORTABLE))
        assert not panicked
        got, _, st = scene.render(cam, rtw.RenderOptions(seed=SEED, precision=rtw.RTW_F64))
        assert got.shape == (2, 3, 3) and np.array_equal(ref, got, equal_nan=True) and st["rays"] == cnt["rays"]
        scene.render(cam, rtw.RenderOptions(seed=SEED, precision=rtw.RTW_F32))
    finally:
        scene.close()


def test_empty_lights_list_maps_the_reference_panic(rtw, oracle):
    """scenes::plane seen from BELOW: Lambertian hits with an empty lights list.  The reference panics on the first path whose mixture
    pdf picks the lights (hittable_list.rs:414-419); the oracle reports it, the library returns RTW_E_INVALID from the render call."""
    from ray_tracing_weekend_b200 import _lib
    world, lights, _ = rtw.scenes.plane()
    cam = (rtw.CameraBuilder().with_image_width(8).with_image_height(6).with_samples_per_pixel(4).with_max_depth(5)
           .with_lookfrom((0., -5., 0.)).with_lookat((0., 0., 0.)).with_vup((1., 0., 0.)).with_focus_dist(5.).with_background((1., 1., 1.)).build())
    scene = rtw.Scene(world, lights)
    try:
        og = oracle.GScene(scene.desc.pod, scene.desc)
        _, _, _, panicked = og.render(oracle.Camera.from_buffer_copy(cam.pod), oracle.options(seed=SEED, math_mode=oracle.PORTABLE))
        assert panicked
        for prec in (rtw.RTW_F64, rtw.RTW_F32):
            with pytest.raises(rtw.RtwError) as e:
                scene.render(cam, rtw.RenderOptions(seed=SEED, precision=prec))
            assert e.value.code == _lib.RTW_E_INVALID and "panics" in str(e.value)
        # scenes::perlin_spheres as the reference builds it: Lambertian spheres in plain view, no lights -> the reference panics
        pw, pl, pcb = rtw.scenes.perlin_spheres(SEED)
        ps = rtw.Scene(pw, pl)
        pcam = pcb.with_image_width(12).with_image_height(12).with_samples_per_pixel(4).build()
        with pytest.raises(rtw.RtwError) as e:
            ps.render(pcam, rtw.RenderOptions(seed=SEED, precision=rtw.RTW_F32))
        assert e.value.code == _lib.RTW_E_INVALID
        ps.close()
        # the handle stays usable: the same scene from above renders (no Lambertian hit, no light sample)
        _, _, cb = rtw.scenes.plane()
        up = cb.with_image_width(4).with_image_height(4).with_samples_per_pixel(2).build()
        img, _, _ = scene.render(up, rtw.RenderOptions(seed=SEED, precision=rtw.RTW_F64))
        assert np.array_equal(img, np.full((4, 4, 3), 2.0))
    finally:
        scene.close()


def test_general_sample_partition_is_bit_identical(rtw, oracle):
    import torch
    from ray_tracing_weekend_b200 import dist as D
    scene, og, cb = _build(rtw, oracle, "cornell_box")
    try:
        w, h, spp = 40, 30, 7
        cam = _cam(cb, w, h, spp, 20)
        opts = rtw.RenderOptions(seed=SEED, precision=rtw.RTW_F32)
        one, one8, _ = scene.render(cam, opts)
        slots = D.tiles_total(w, h) * 256
        blocks = torch.zeros((2, D.accum_words(w, h)), dtype=torch.int64, device="cuda")
        for r in range(2):
            b, c = D.sample_range(spp, r, 2)
            scene.render_samples_device(cam, opts, b, c, blocks[r].data_ptr(), blocks[r].data_ptr() + 8 * 3 * slots)
        total = blocks.sum(dim=0)
        out = torch.zeros((h, w, 3), dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        rtw.resolve_accum_device(total.data_ptr(), total.data_ptr() + 8 * 3 * slots, w, h, spp, out.data_ptr(), 0)
        torch.cuda.synchronize()
        assert np.array_equal(one, out.cpu().numpy(), equal_nan=True)
        with pytest.raises(rtw.RtwError):                         # f64 sums samples in order: no sample partition
            scene.render_samples_device(cam, rtw.RenderOptions(seed=SEED, precision=rtw.RTW_F64), 0, 3, blocks[0].data_ptr(),
                                        blocks[0].data_ptr() + 8 * 3 * slots)
    finally:
        scene.close()


@pytest.mark.parametrize("name", ("cornell_box", "simple_light", "debugging_scene"))
def test_general_trace_f32_tolerance(rtw, oracle, name):
    """North-star check (1) for the FP32 general path: the same (f32-representable) rays against the f64 oracle — hit ids equal
    except near silhouettes / edges, t within 1e-5 relative (tmin well above FP32 noise so self-intersections stay out)."""
    scene, og, cb = _build(rtw, oracle, name)
    try:
        cam = _cam(cb, 64, 64, 4, 12)
        o, d = _rays(oracle, og, cam.pod, 4000, 3)
        o = o.astype(np.float32).astype(np.float64); d = d.astype(np.float32).astype(np.float64)
        tmin = 1e-2 if name == "cornell_box" else 1e-3               # cornell coordinates are ~555: FP32 noise of a hit point ~3e-5
        prim_o, t_o, _, _ = og.trace_batch(o, d, tmin=tmin)
        prim_g, t_g = scene.trace_batch(o, d, tmin=tmin, precision=rtw.RTW_F32)
        same = prim_o == prim_g
        assert same.mean() > 0.99, f"id mismatch fraction {1 - same.mean():.4%}"
        both = same & (prim_o >= 0)
        rel = np.abs(t_g[both] - t_o[both]) / np.abs(t_o[both])
        assert np.median(rel) < 2e-6 and np.quantile(rel, 0.99) < 1e-4, (np.median(rel), np.quantile(rel, 0.99), rel.max())
    finally:
        scene.close()


@pytest.mark.parametrize("name", ("cornell_box", "checkered_spheres", "random"))
def test_general_wavefront_is_bit_identical_to_megakernel(rtw, oracle, name):
    """The two FP32 renderers of the general path (warp-private wavefront / pooled megakernel) trace the same paths with the same
    arithmetic and accumulate in fixed point: identical images, identical ray counts."""
    if name == "random":
        world, lights = _random_scene(rtw, np.random.default_rng(11))
        scene = rtw.Scene(world, lights)
        cb = rtw.CameraBuilder().with_lookfrom((0., 2., 16.)).with_lookat((0., 0., 0.)).with_focus_dist(16.).with_background((0.6, 0.7, 0.9))
    else:
        scene, _, cb = _build(rtw, oracle, name)
    try:
        cam = _cam(cb, 96, 64, 24, 30)
        a, a8, sa = scene.render(cam, rtw.RenderOptions(seed=SEED, mode=rtw.RTW_WAVEFRONT))
        b, b8, sb = scene.render(cam, rtw.RenderOptions(seed=SEED, mode=rtw.RTW_MEGAKERNEL))
        assert sa["rays"] == sb["rays"] and sa["paths"] == sb["paths"] == 96 * 64 * 24
        assert np.array_equal(a, b, equal_nan=True) and np.array_equal(a8, b8)
        c, _, sc_ = scene.render(cam, rtw.RenderOptions(seed=SEED, mode=rtw.RTW_WAVEFRONT, flags=rtw.RTW_FLAG_COUNT_EVENTS))
        assert np.array_equal(a, c, equal_nan=True) and sc_["rays"] == sa["rays"]
    finally:
        scene.close()


def test_checkered_plane_of_any_orientation(rtw, oracle):
    """get_plane_uv (plane.rs:41-55): for a normal other than +y the hit point is rotated about (n x +y) onto +y and (u, v) is the
    fractional part of x and z.  The per-plane constants (theta, cos, sin, axis) come from the host's libm like the reference's per-hit
    calls; vertices (albedo = checker colour at the hit) and the image are bit-exact on the f64 path, a -y normal is refused (the
    reference's axis is 0 / 0 there and Plane::hit panics on the NaN).  Also here: a CheckerTexture whose even / odd are a finer
    CheckerTexture and a NoiseTexture (get_colour recurses, texture.rs:46-55), on a sphere and on a quad."""
    from ray_tracing_weekend_b200 import _lib
    checker = rtw.Lambertian(rtw.CheckerTexture.new_with_colours((0.9, 0.1, 0.1), (0.1, 0.1, 0.9), 0.37))
    light = rtw.DiffuseLight((3., 3., 3.))
    world, lights = rtw.HittableList(), rtw.HittableList()
    for point, normal in (((0., -1., 0.), (0.2, 1., -0.4)), ((0., 4., 0.), (0.5, -0.7, 0.1)), ((-6., 0., 0.), (-1., 0., 0.)),
                          ((0., 0., -7.), (0., 1e-9, -1.))):
        world.add(rtw.Plane(point, normal, checker))
    world.add(rtw.Sphere((0., 1., 0.), 0.8, rtw.Metal((0.8, 0.8, 0.8), 0.05)))
    # CheckerTexture's even / odd are textures themselves (texture.rs:26-29): a checker of (a finer checker, a noise texture)
    nested = rtw.Lambertian(rtw.CheckerTexture(rtw.CheckerTexture.new_with_colours((1., 1., 0.), (0., 1., 1.), 0.05),
                                               rtw.NoiseTexture(3.0, SEED, 5), 0.2))
    world.add(rtw.Sphere((-1.8, 0.6, 1.), 0.9, nested))
    world.add(rtw.Quad((1.2, -0.5, 1.5), (1.5, 0., 0.3), (0., 1.5, 0.2), nested))
    lights.add(rtw.Sphere((2., 2., 1.), 0.4, light)); world.add(lights.items[0])
    scene = rtw.Scene(world, lights)
    og = oracle.GScene(scene.desc.pod, scene.desc)
    try:
        cb = (rtw.CameraBuilder().with_lookfrom((3., 1.5, 5.)).with_lookat((0., 1., 0.)).with_background((0.6, 0.7, 0.9)))
        cam = _cam(cb, 40, 30, 4, 8)
        o, d = _rays(oracle, og, cam.pod, 1500, 5)
        prim_o, t_o, _, _ = og.trace_batch(o, d)
        prim_g, t_g = scene.trace_batch(o, d, precision=rtw.RTW_F64)
        assert np.array_equal(prim_o, prim_g) and np.array_equal(t_o, t_g)
        assert all((prim_o == k).sum() > 20 for k in range(4)), [(prim_o == k).sum() for k in range(6)]     # every plane is hit
        n = len(o)
        rng = np.random.default_rng(2)
        pixel, sample, vertex = rng.integers(0, 1200, n), rng.integers(0, 4, n), rng.integers(1, 8, n)
        opts = rtw.RenderOptions(seed=SEED, precision=rtw.RTW_F64)
        a = og.scatter_batch(o, d, pixel, sample, vertex, oracle.options(seed=SEED, math_mode=oracle.PORTABLE))
        b = scene.scatter_batch(o, d, pixel, sample, vertex, opts)
        assert np.array_equal(a["prim"], b["prim"]) and np.array_equal(a["kind"], b["kind"])
        for k in ("t", "p", "normal", "dir", "weight"):
            assert np.array_equal(a[k], b[k], equal_nan=True), k
        lamb = a["kind"] == 3                                          # V_DIFFUSE vertices on the planes carry the checker colour
        assert len(np.unique(np.round(a["weight"][lamb & (a["prim"] < 4)], 12), axis=0)) > 2
        assert (lamb & (a["prim"] == 5)).sum() > 20 and (lamb & (a["prim"] == 6)).sum() > 5              # the nested checkers are hit
        img_o, _, _, _ = og.render(oracle.Camera.from_buffer_copy(cam.pod), oracle.options(seed=SEED, math_mode=oracle.PORTABLE))
        img_g, _, _ = scene.render(cam, opts)
        assert np.array_equal(img_o, img_g, equal_nan=True)
        # FP32: both checker colours show up on every plane (the pattern is there), and the image is close to the f64 one
        # (at a tmin above the rounding noise: with the reference's tmin the self-intersection statistics of FP32 and f64 differ)
        img32, _, _ = scene.render(_cam(cb, 40, 30, 64, 8), rtw.RenderOptions(seed=SEED, precision=rtw.RTW_F32, tmin=1e-3))
        img64, _, _ = scene.render(_cam(cb, 40, 30, 64, 8), rtw.RenderOptions(seed=SEED + 1, precision=rtw.RTW_F64, tmin=1e-3))
        ok = np.isfinite(img32).all(axis=2) & np.isfinite(img64).all(axis=2)
        m32, m64 = np.clip(img32[ok] / 64, 0, 2).mean(), np.clip(img64[ok] / 64, 0, 2).mean()
        assert ok.mean() > 0.9 and abs(m32 - m64) < 0.05 * m64, (ok.mean(), m32, m64)
    finally:
        scene.close()
    bad = rtw.HittableList(); bad.add(rtw.Plane((0., 3., 0.), (0., -2., 0.), checker))
    with pytest.raises(rtw.RtwError) as e:
        rtw.Scene(bad, lights)
    assert e.value.code == _lib.RTW_E_INVALID and "-y" in str(e.value)
