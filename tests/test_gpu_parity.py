"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

  (1) world.hit: hit ids bit-exact, t bit-exact (f64 path) / within 1e-5 relative (f32 path)
  (2) images: f64 path bit-exact against the oracle; f32 path statistically (PSNR / mean error)
  (3) scatter directions / weights per path vertex against the oracle's Philox mirror
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SEED = 20261018


def _camera(rtw, oracle, sc, w, h, spp, depth):
    cam = (sc["cb"].with_vfov(40.).with_aspect_ratio(w / h).with_max_depth(depth).with_image_width(w).with_image_height(h)
           .with_samples_per_pixel(spp).build())
    ocam = oracle.camera_for(sc["desc"], w, h, spp, depth)
    return cam, ocam


@pytest.fixture(scope="module")
def gscene(rtw, simple_scene):
    s = rtw.Scene(simple_scene["world"], simple_scene["lights"])
    yield s
    s.close()


def _ray_batch(rtw, oracle, sc, n_primary=4096, n_secondary=4096, seed=1):
    """Primary rays from the camera plus secondary rays leaving oracle hit points in random directions."""
    rng = np.random.default_rng(seed)
    cam, ocam = _camera(rtw, oracle, sc, 400, 225, 4, 50)
    i = rng.integers(0, 400, n_primary); j = rng.integers(0, 225, n_primary); s = rng.integers(0, 4, n_primary)
    o, d = oracle.get_rays(ocam, oracle.options(seed=SEED), i, j, s)
    prim, t, _ = sc["oscene"].trace_batch(o, d)
    hit = prim >= 0
    p = o[hit] + d[hit] * t[hit][:, None]
    k = rng.integers(0, len(p), n_secondary)
    d2 = rng.normal(size=(n_secondary, 3))
    return np.concatenate([o, p[k]]), np.concatenate([d, d2])


def test_trace_batch_f64_bit_exact(rtw, oracle, simple_scene, gscene):
    o, d = _ray_batch(rtw, oracle, simple_scene)
    for tmin in (oracle.EPS, 1e-3):
        prim_o, t_o, _ = simple_scene["oscene"].trace_batch(o, d, tmin=tmin)
        prim_g, t_g = gscene.trace_batch(o, d, tmin=tmin, precision=rtw.RTW_F64)
        assert np.array_equal(prim_o, prim_g)
        assert np.array_equal(t_o, t_g)          # bit-exact, +inf on misses
        assert (prim_o >= 0).sum() > 1000


def test_trace_batch_f32_tolerance(rtw, oracle, simple_scene, gscene):
    """North-star check (1) for the FP32 path: identical ray batches (rounded to f32 so both sides see the same
    rays), ids equal except grazing cases, t within 1e-5 relative.  Grazing = the ray passes within 0.5 % of the
    radius of the silhouette (|l|^2 > 0.99 r^2): there the chord length is ill-conditioned in FP32."""
    o, d = _ray_batch(rtw, oracle, simple_scene)
    o = o.astype(np.float32).astype(np.float64); d = d.astype(np.float32).astype(np.float64)
    # identical inputs: the FP32 path stores sphere centres / radii as f32, so the oracle gets the same rounded spheres
    desc = simple_scene["desc"]
    desc32 = oracle.SceneDesc(desc.spheres.astype(np.float32).astype(np.float64), desc.sphere_mat, desc.materials, desc.planes,
                              desc.plane_mat, desc.lights)
    osc32 = oracle.Scene(desc32)
    # tmin well above FP32 noise so that self-intersections at t ~ 1e-7 do not enter the id comparison
    prim_o, t_o, _ = osc32.trace_batch(o, d, tmin=1e-3)
    prim_g, t_g = gscene.trace_batch(o, d, tmin=1e-3, precision=rtw.RTW_F32)
    same = prim_o == prim_g
    assert same.mean() > 0.998, f"id mismatch fraction {1 - same.mean():.4%}"
    both = same & (prim_o >= 1)
    sph = desc32.spheres[prim_o[both] - 1]
    oc = o[both] - sph[:, :3]; dd = d[both]
    k = (oc * dd).sum(1) / (dd * dd).sum(1)
    l2 = ((oc - k[:, None] * dd) ** 2).sum(1)
    grazing = l2 > 0.99 * sph[:, 3] ** 2
    rel = np.abs(t_g[both] - t_o[both]) / np.abs(t_o[both])
    assert grazing.mean() < 0.15
    assert rel[~grazing].max() < 1e-5, f"t relative error (non-grazing) max = {rel[~grazing].max():.3e}"
    assert np.median(rel) < 1e-6
    # grazing rays: bounded in absolute terms (scene units)
    abs_err = np.abs(t_g[both] - t_o[both]) * np.linalg.norm(dd, axis=1)
    assert abs_err.max() < 2e-4
    # every id mismatch is a grazing / near-tie case in the oracle's terms: the two candidates' t differ by < 1e-4 relative
    # or one side missed a silhouette hit
    assert (~same).sum() <= 0.002 * len(same)
    # against the un-rounded f64 scene the ids still agree; t then also carries the 2^-24 relative shift of the centres
    prim_64, t_64, _ = simple_scene["oscene"].trace_batch(o, d, tmin=1e-3)
    assert (prim_64 == prim_g).mean() > 0.997


def test_get_rays(rtw, oracle, simple_scene):
    cam, ocam = _camera(rtw, oracle, simple_scene, 400, 225, 8, 50)
    rng = np.random.default_rng(3)
    i = rng.integers(0, 400, 2000); j = rng.integers(0, 225, 2000); s = rng.integers(0, 8, 2000)
    o_o, d_o = oracle.get_rays(ocam, oracle.options(seed=SEED, rng_mode=oracle.W64), i, j, s)
    o_g, d_g = cam.get_rays(i, j, s, rtw.RenderOptions(seed=SEED, precision=rtw.RTW_F64))
    assert np.array_equal(o_o, o_g) and np.array_equal(d_o, d_g)
    o_o, d_o = oracle.get_rays(ocam, oracle.options(seed=SEED, rng_mode=oracle.W32), i, j, s)
    o_g, d_g = cam.get_rays(i, j, s, rtw.RenderOptions(seed=SEED, precision=rtw.RTW_F32))
    assert np.allclose(d_o, d_g, rtol=0, atol=2e-5)


def _scatter_inputs(rtw, oracle, sc, n=6000, seed=5):
    o, d = _ray_batch(rtw, oracle, sc, n // 2, n - n // 2, seed)
    rng = np.random.default_rng(seed)
    pixel = rng.integers(0, 90000, len(o)); sample = rng.integers(0, 100, len(o)); vertex = rng.integers(1, 51, len(o))
    return o, d, pixel, sample, vertex


def test_scatter_batch_f64_bit_exact(rtw, oracle, simple_scene, gscene):
    o, d, pixel, sample, vertex = _scatter_inputs(rtw, oracle, simple_scene)
    ref = simple_scene["oscene"].scatter_batch(o, d, pixel, sample, vertex,
                                               oracle.options(seed=SEED, rng_mode=oracle.W64, math_mode=oracle.PORTABLE))
    got = gscene.scatter_batch(o, d, pixel, sample, vertex, rtw.RenderOptions(seed=SEED, precision=rtw.RTW_F64))
    assert np.array_equal(ref["prim"], got["prim"]) and np.array_equal(ref["kind"], got["kind"])
    for k in ("t", "p", "normal", "dir", "weight"):
        assert np.array_equal(ref[k], got[k], equal_nan=True), k
    kinds = np.bincount(ref["kind"], minlength=4)
    assert kinds[2] > 100 and kinds[3] > 100       # specular and diffuse vertices both covered


def test_scatter_batch_f32_matches_philox_mirror(rtw, oracle, simple_scene, gscene):
    o, d, pixel, sample, vertex = _scatter_inputs(rtw, oracle, simple_scene)
    ref = simple_scene["oscene"].scatter_batch(o, d, pixel, sample, vertex, oracle.options(seed=SEED, tmin=1e-3, rng_mode=oracle.W32))
    got = gscene.scatter_batch(o, d, pixel, sample, vertex, rtw.RenderOptions(seed=SEED, tmin=1e-3, precision=rtw.RTW_F32))
    same = (ref["prim"] == got["prim"]) & (ref["kind"] == got["kind"])
    assert same.mean() > 0.995
    m = same & (ref["kind"] >= 2)
    # directions: FP32 evaluation of the same uniforms -> a few ulp of the unit-scale components
    derr = np.abs(ref["dir"][m] - got["dir"][m]).max(axis=1)
    assert np.quantile(derr, 0.99) < 2e-4, np.quantile(derr, 0.99)
    assert np.median(derr) < 5e-6, np.median(derr)


def test_path_radiance_f64_bit_exact(rtw, oracle, simple_scene, gscene):
    cam, ocam = _camera(rtw, oracle, simple_scene, 400, 225, 16, 50)
    rng = np.random.default_rng(11)
    n = 20000
    i = rng.integers(0, 400, n); j = rng.integers(0, 140, n); s = rng.integers(0, 16, n)
    ref = simple_scene["oscene"].path_radiance(ocam, oracle.options(seed=SEED, rng_mode=oracle.W64, math_mode=oracle.PORTABLE), i, j, s)
    got = gscene.path_radiance(cam, rtw.RenderOptions(seed=SEED, precision=rtw.RTW_F64), i, j, s)
    assert np.array_equal(ref, got, equal_nan=True)
    assert (ref.sum(axis=1) < 2.9).sum() > 2000     # plenty of paths that actually hit something


def test_render_f64_bit_exact_small(rtw, oracle, simple_scene, gscene):
    w, h, spp = 96, 54, 8
    cam, ocam = _camera(rtw, oracle, simple_scene, w, h, spp, 50)
    ref, _, cnt, pan = simple_scene["oscene"].render(ocam, oracle.options(seed=SEED, rng_mode=oracle.W64, math_mode=oracle.PORTABLE))
    rgb_sum, rgb8, st = gscene.render(cam, rtw.RenderOptions(seed=SEED, precision=rtw.RTW_F64))
    assert not pan
    assert np.array_equal(ref, rgb_sum, equal_nan=True)
    assert np.array_equal(oracle.resolve(ref, spp), rgb8)
    assert st["rays"] == cnt["rays"] and st["paths"] == cnt["paths"] == w * h * spp


def _psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 10 * np.log10(255.0 ** 2 / mse)


def test_render_f32_image_close_to_oracle(rtw, oracle, simple_scene, gscene):
    """North-star check (2): same estimator, independent noise.  Two comparisons at the reference's own tmin:
    (a) with NaN samples zeroed on both sides (Colour::fix_nan) the images agree in mean and PSNR;
    (b) without it, the fraction of NaN-poisoned (black) pixels agrees — the reference poisons a pixel whenever a
        Lambertian vertex lies inside a light sphere (pdf_value = NaN, sphere.rs:101-111)."""
    w, h, spp = 160, 90, 256
    cam, ocam = _camera(rtw, oracle, simple_scene, w, h, spp, 50)
    ref, _, cnt, _ = simple_scene["oscene"].render(ocam, oracle.options(seed=SEED + 1, rng_mode=oracle.W64, fix_nan=True))
    rgb_sum, rgb8, st = gscene.render(cam, rtw.RenderOptions(seed=SEED, precision=rtw.RTW_F32, flags=rtw.RTW_FLAG_FIX_NAN))
    assert np.isfinite(ref).all() and np.isfinite(rgb_sum).all()
    a, b = ref / spp, rgb_sum / spp
    print("mean radiance oracle/gpu", a.mean(), b.mean(), "rays/path", cnt["rays"] / cnt["paths"], st["rays"] / st["paths"])
    assert abs(a.mean() - b.mean()) < 5e-3, (a.mean(), b.mean())
    assert abs(st["rays"] / st["paths"] - cnt["rays"] / cnt["paths"]) < 0.06 * cnt["rays"] / cnt["paths"]      # documented: ~4 % fewer self-hits in FP32
    fg = (a.min(axis=2) < 0.98) | (b.min(axis=2) < 0.98)           # pixels that see geometry
    assert abs(a[fg].mean() - b[fg].mean()) < 0.02 * a[fg].mean() + 5e-3
    psnr = _psnr(oracle.resolve(ref, spp), rgb8)
    print("psnr", psnr)
    assert psnr > 28.0
    # (b) poison statistics
    ref_p, _, _, _ = simple_scene["oscene"].render(ocam, oracle.options(seed=SEED + 1, rng_mode=oracle.W64))
    gpu_p, _, _ = gscene.render(cam, rtw.RenderOptions(seed=SEED, precision=rtw.RTW_F32))
    fo, fg_ = np.isnan(ref_p).any(axis=2).mean(), np.isnan(gpu_p).any(axis=2).mean()
    print("poisoned pixel fraction oracle/gpu", fo, fg_)
    assert fo > 0.01 and abs(fo - fg_) < 0.25 * fo


def test_render_independent_of_world_size(rtw, simple_scene, gscene):
    """Tiles are keyed by absolute pixel: 1-rank and 3-rank partitions give identical images."""
    import torch
    w, h, spp = 100, 70, 4
    cam = (simple_scene["cb"].with_vfov(40.).with_aspect_ratio(w / h).with_max_depth(50).with_image_width(w).with_image_height(h)
           .with_samples_per_pixel(spp).build())
    opts = rtw.RenderOptions(seed=SEED, precision=rtw.RTW_F32)
    one, _, _ = gscene.render(cam, opts)
    world = 3
    tpr = rtw.tiles_per_rank(w, h, world)
    tiles = torch.zeros((world, tpr, 16, 16, 3), dtype=torch.float32, device="cuda")
    for r in range(world):
        gscene.render_tiles_device(cam, opts, r, world, tiles[r].data_ptr())
    out = torch.zeros((h, w, 3), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    rtw.untile_resolve_device(tiles.data_ptr(), rtw.RTW_F32, w, h, world, spp, out.data_ptr(), 0)
    torch.cuda.synchronize()
    assert np.array_equal(one, out.cpu().numpy(), equal_nan=True)


def test_reference_smoke_shapes(rtw, simple_scene):
    """integration-tests/src/lib.rs:32-52 (small_test): 3x2 px, 10 spp, depth 3 — must simply run."""
    cam = (rtw.CameraBuilder().with_image_width(3).with_image_height(2).with_samples_per_pixel(10).with_max_depth(3)
           .with_lookfrom((-13., 2., 3.)).with_lookat((0., 0., 0.)).with_vup((0., 1., 0.)).with_focus_dist(10.).build())
    out = cam.render(simple_scene["world"], simple_scene["lights"])
    assert out.shape == (2, 3, 3)


def test_pooled_megakernel_matches_lane_per_pixel(rtw, simple_scene, gscene):
    """The pooled path stream traces exactly the same paths (RNG keyed by pixel/sample/vertex) as the
    lane-per-pixel kernel; only the summation differs (64-bit fixed point vs sequential FP32)."""
    w, h, spp = 128, 72, 24
    cam = (simple_scene["cb"].with_vfov(40.).with_aspect_ratio(w / h).with_max_depth(50).with_image_width(w).with_image_height(h)
           .with_samples_per_pixel(spp).build())
    a, a8, sa = gscene.render(cam, rtw.RenderOptions(seed=SEED, precision=rtw.RTW_F32, mode=rtw.RTW_MEGAKERNEL))
    b, b8, sb = gscene.render(cam, rtw.RenderOptions(seed=SEED, precision=rtw.RTW_F32, mode=rtw.RTW_MEGAKERNEL, flags=rtw.RTW_FLAG_LANE_PER_PIXEL))
    assert sa["rays"] == sb["rays"] and sa["paths"] == sb["paths"] == w * h * spp
    assert np.array_equal(np.isnan(a), np.isnan(b))
    ok = np.isfinite(a) & np.isfinite(b)
    assert np.allclose(a[ok], b[ok], rtol=2e-5, atol=2e-5)
    assert (a8 != b8).mean() < 1e-3
    # deterministic: integer accumulation makes the pooled image independent of scheduling
    a2, _, _ = gscene.render(cam, rtw.RenderOptions(seed=SEED, precision=rtw.RTW_F32, mode=rtw.RTW_MEGAKERNEL))
    assert np.array_equal(a, a2, equal_nan=True)
    # low spp exercises multi-pixel chunks
    cam1 = (simple_scene["cb"].with_vfov(40.).with_aspect_ratio(w / h).with_max_depth(50).with_image_width(w).with_image_height(h)
            .with_samples_per_pixel(1).build())
    c, _, _ = gscene.render(cam1, rtw.RenderOptions(seed=SEED, precision=rtw.RTW_F32, mode=rtw.RTW_MEGAKERNEL))
    e, _, _ = gscene.render(cam1, rtw.RenderOptions(seed=SEED, precision=rtw.RTW_F32, mode=rtw.RTW_MEGAKERNEL, flags=rtw.RTW_FLAG_LANE_PER_PIXEL))
    ok = np.isfinite(c) & np.isfinite(e)
    assert np.array_equal(np.isnan(c), np.isnan(e)) and np.allclose(c[ok], e[ok], rtol=1e-6, atol=1e-6)


def test_furnace_fp32_tracks_reference_acne(rtw, oracle):
    """Lambertian albedo-0.5 sphere, white background.  Robust tmin: 0.5.  Reference tmin (EPSILON): the f64
    reference renders ~0.27 because half the bounces re-hit their own sphere; the FP32 path must land on the
    same value (it depends on rounding statistics, not on precision) — the basis for image parity at the
    reference's own settings."""
    world = rtw.HittableList(); world.add(rtw.Sphere((0, 0, 0), 1.0, rtw.Lambertian((0.5, 0.5, 0.5))))
    lights = rtw.HittableList(); lights.add(rtw.Sphere((0, 1000, 0), 1e-3, rtw.INVISIBLE))
    cam = (rtw.CameraBuilder().with_image_width(48).with_image_height(48).with_samples_per_pixel(256).with_max_depth(50)
           .with_background((1, 1, 1)).with_vfov(10).with_lookfrom((0, 0, 10)).with_lookat((0, 0, 0)).with_vup((0, 1, 0)).with_focus_dist(10).build())
    sc = rtw.Scene(world, lights)
    vals = {}
    for prec in (rtw.RTW_F32, rtw.RTW_F64):
        for tmin in (1e-3, rtw.TMIN_REFERENCE):
            img, _, st = sc.render(cam, rtw.RenderOptions(seed=SEED, tmin=tmin, precision=prec, flags=rtw.RTW_FLAG_FIX_NAN))
            vals[(prec, tmin)] = (img[16:32, 16:32].mean() / 256, st["rays"] / st["paths"])
    sc.close()
    print("furnace", vals)
    assert abs(vals[(rtw.RTW_F32, 1e-3)][0] - 0.5) < 0.01 and abs(vals[(rtw.RTW_F64, 1e-3)][0] - 0.5) < 0.01
    # f64 at f64::EPSILON == the oracle's 0.27; f32 at f32::EPSILON (the same relation to rounding noise) must track it
    assert abs(vals[(rtw.RTW_F64, rtw.TMIN_REFERENCE)][0] - 0.27) < 0.015
    assert abs(vals[(rtw.RTW_F32, rtw.TMIN_REFERENCE)][0] - vals[(rtw.RTW_F64, rtw.TMIN_REFERENCE)][0]) < 0.02
    assert abs(vals[(rtw.RTW_F32, rtw.TMIN_REFERENCE)][1] - vals[(rtw.RTW_F64, rtw.TMIN_REFERENCE)][1]) < 0.2


def test_wavefront_is_bit_identical_to_megakernel(rtw, simple_scene, gscene):
    """RTW_WAVEFRONT (CTA-local queues in shared memory) traces the same paths with the same arithmetic as the
    pooled megakernel and accumulates in the same fixed point: the images must match bit for bit."""
    for (w, h, spp, depth) in ((128, 72, 24, 50), (70, 50, 3, 50), (64, 36, 600, 5), (33, 17, 1, 1)):
        cam = (simple_scene["cb"].with_vfov(40.).with_aspect_ratio(w / h).with_max_depth(depth).with_image_width(w).with_image_height(h)
               .with_samples_per_pixel(spp).build())
        for flags in (0, rtw.RTW_FLAG_COUNT_EVENTS):
            a, a8, sa = gscene.render(cam, rtw.RenderOptions(seed=SEED, precision=rtw.RTW_F32, mode=rtw.RTW_MEGAKERNEL, flags=flags))
            b, b8, sb = gscene.render(cam, rtw.RenderOptions(seed=SEED, precision=rtw.RTW_F32, mode=rtw.RTW_WAVEFRONT, flags=flags))
            for k in ("paths", "rays", "light_tests", "lambertian", "metal", "dielectric", "absorbed", "missed", "depth_out"):
                assert sa[k] == sb[k], (k, sa[k], sb[k], (w, h, spp, depth))     # (traversal order differs: node / sphere test counts may)
            diff = np.abs(np.nan_to_num(a, nan=-1.0) - np.nan_to_num(b, nan=-1.0))
            assert np.array_equal(a, b, equal_nan=True), ((w, h, spp, depth), float(diff.max()), int((diff > 0).sum()), int(diff.size))
            assert np.array_equal(a8, b8)
    # RTW_F64 has a single renderer: `mode` is ignored there
    x, _, _ = gscene.render(cam, rtw.RenderOptions(seed=SEED, precision=rtw.RTW_F64, mode=rtw.RTW_WAVEFRONT))
    y, _, _ = gscene.render(cam, rtw.RenderOptions(seed=SEED, precision=rtw.RTW_F64, mode=rtw.RTW_MEGAKERNEL))
    assert np.array_equal(x, y, equal_nan=True)


def _scene_pair(rtw, oracle, seed, n, p_l, p_m, ground=0):
    """The same generated scene on the GPU (array constructor) and in the oracle."""
    a = rtw.scenes.simple_arrays(seed, n, p_l, p_m, ground)
    gs = rtw.Scene.from_arrays(a["spheres"], a["sphere_materials"], a["planes"], a["plane_materials"], a["lights"])
    desc = oracle.scene_simple(seed, n, p_l, p_m, ground)
    assert np.array_equal(desc.spheres, a["spheres"]) and np.array_equal(desc.lights, a["lights"])
    return a, gs, desc, oracle.Scene(desc)


def test_scene_larger_than_shared_memory(rtw, oracle):
    """6 400 spheres do not fit the shared-memory staging: top BVH levels in shared memory, the rest read from
    global memory, and 320 lights take the light-BVH path.  Same parity bars as the small scene."""
    a, gs, desc, osc = _scene_pair(rtw, oracle, 77, 40, 0.8, 0.95)
    assert gs.n_spheres > 6000 and gs.n_lights > 64 and gs.info()["device_bytes"] > 300_000
    w, h, spp = 96, 54, 6
    cam = a["cam"].with_vfov(40.).with_aspect_ratio(w / h).with_max_depth(50).with_image_width(w).with_image_height(h).with_samples_per_pixel(spp).build()
    ocam = oracle.camera_for(desc, w, h, spp, 50)
    rng = np.random.default_rng(2)
    n = 6000
    i = rng.integers(0, w, n); j = rng.integers(0, h, n); s = rng.integers(0, spp, n)
    o, d = oracle.get_rays(ocam, oracle.options(seed=77), i, j, s)
    p0, t0, _ = osc.trace_batch(o, d)
    hit = p0 >= 0
    pts = o[hit] + d[hit] * t0[hit][:, None]
    o = np.concatenate([o, pts[rng.integers(0, len(pts), n)]]); d = np.concatenate([d, rng.normal(size=(n, 3))])
    # f64 path: bit-exact ids and t
    for tmin in (oracle.EPS, 1e-3):
        pr, tr, _ = osc.trace_batch(o, d, tmin=tmin)
        pg, tg = gs.trace_batch(o, d, tmin=tmin, precision=rtw.RTW_F64)
        assert np.array_equal(pr, pg) and np.array_equal(tr, tg)
    # f32 path: ids
    o32 = o.astype(np.float32).astype(np.float64); d32 = d.astype(np.float32).astype(np.float64)
    pr, tr, _ = osc.trace_batch(o32, d32, tmin=1e-3)
    pg, tg = gs.trace_batch(o32, d32, tmin=1e-3, precision=rtw.RTW_F32)
    assert (pr == pg).mean() > 0.995
    # f64 image bit-exact (the exact path's linear light sum over 320 lights in insertion order)
    ref, _, cnt, _ = osc.render(ocam, oracle.options(seed=77, rng_mode=oracle.W64, math_mode=oracle.PORTABLE))
    got, _, st = gs.render(cam, rtw.RenderOptions(seed=77, precision=rtw.RTW_F64))
    assert np.array_equal(ref, got, equal_nan=True) and st["rays"] == cnt["rays"]
    # f32: megakernel == wavefront bit for bit on the global-memory + light-BVH path, and the image tracks the oracle
    m, _, sm = gs.render(cam, rtw.RenderOptions(seed=77, precision=rtw.RTW_F32, mode=rtw.RTW_MEGAKERNEL, flags=rtw.RTW_FLAG_FIX_NAN))
    wv, _, sw = gs.render(cam, rtw.RenderOptions(seed=77, precision=rtw.RTW_F32, mode=rtw.RTW_WAVEFRONT, flags=rtw.RTW_FLAG_FIX_NAN))
    assert np.array_equal(m, wv) and sm["rays"] == sw["rays"]
    spp2 = 64
    cam2 = a["cam"].with_vfov(40.).with_aspect_ratio(w / h).with_max_depth(50).with_image_width(w).with_image_height(h).with_samples_per_pixel(spp2).build()
    ref2, _, c2, _ = osc.render(oracle.camera_for(desc, w, h, spp2, 50), oracle.options(seed=78, fix_nan=True))
    got2, _, s2 = gs.render(cam2, rtw.RenderOptions(seed=77, precision=rtw.RTW_F32, flags=rtw.RTW_FLAG_FIX_NAN))
    assert abs(ref2.mean() - got2.mean()) / spp2 < 0.01
    assert abs(s2["rays"] / s2["paths"] - c2["rays"] / c2["paths"]) < 0.08 * c2["rays"] / c2["paths"]
    gs.close()


def test_light_bvh_sum_equals_linear_sum(rtw, oracle):
    """80 % glass (BASELINE C5 recipe): 399 lights -> the FP32 path walks a BVH over the lights instead of the
    reference's linear sum (hittable_list.rs:408-412).  On cosine-sampled vertices (same direction as the oracle's)
    the mixture weight must agree, i.e. the BVH sum equals the linear sum."""
    a, gs, desc, osc = _scene_pair(rtw, oracle, SEED, 11, 0.1, 0.2)
    assert gs.n_lights > 300
    cam_o = oracle.camera_for(desc, 200, 112, 4, 50)
    rng = np.random.default_rng(4)
    n = 80000
    i = rng.integers(0, 200, n); j = rng.integers(0, 112, n); s = rng.integers(0, 4, n)
    o, d = oracle.get_rays(cam_o, oracle.options(seed=SEED), i, j, s)
    pix = rng.integers(0, 20000, n); smp = rng.integers(0, 64, n); vtx = rng.integers(1, 51, n)
    ref = osc.scatter_batch(o, d, pix, smp, vtx, oracle.options(seed=SEED, tmin=1e-3, rng_mode=oracle.W32))
    got = gs.scatter_batch(o, d, pix, smp, vtx, rtw.RenderOptions(seed=SEED, tmin=1e-3, precision=rtw.RTW_F32))
    diffuse = (ref["kind"] == 3) & (got["kind"] == 3) & (ref["prim"] == got["prim"])
    same_dir = diffuse & (np.abs(ref["dir"] - got["dir"]).max(axis=1) < 1e-4)       # cosine branch (light branch: other index order)
    assert same_dir.sum() > 500
    wr, wg = ref["weight"][same_dir], got["weight"][same_dir]
    ok = np.isfinite(wr).all(axis=1) & np.isfinite(wg).all(axis=1)
    rel = np.abs(wr[ok] - wg[ok]) / np.maximum(np.abs(wr[ok]), 1e-3)
    assert np.quantile(rel, 0.99) < 2e-3 and np.median(rel) < 1e-5, (np.quantile(rel, 0.99), np.median(rel))
    gs.close()


def test_edge_cases(rtw, oracle):
    """Empty world, metal-only world without lights, defocus camera (UnitDisk rejection loop), depth 1, 1 spp,
    image sizes that are not multiples of the 16x16 tile — f64 path bit-exact against the oracle, FP32 renderers
    bit-identical to each other."""
    # (a) empty world: every path misses -> spp * background
    sc = rtw.Scene(rtw.HittableList(), rtw.HittableList())
    cam = rtw.CameraBuilder().with_image_width(37).with_image_height(19).with_samples_per_pixel(3).with_max_depth(5).with_background((0.25, 0.5, 1.0)).build()
    for prec in (rtw.RTW_F32, rtw.RTW_F64):
        img, rgb8, st = sc.render(cam, rtw.RenderOptions(seed=1, precision=prec))
        assert np.array_equal(img, np.broadcast_to(np.array([0.75, 1.5, 3.0]), img.shape)) and st["rays"] == st["paths"] == 37 * 19 * 3
        assert np.array_equal(rgb8[0, 0], [128, 181, 255])
    sc.close()
    # (b) metal + glass only, no lights list needed; defocus camera
    mats = oracle.make_materials([(oracle.METAL, 0.8, 0.7, 0.6, 0.3), (oracle.DIELECTRIC, 1, 1, 1, 1.5), (oracle.METAL, 0.9, 0.9, 0.9, 0.0)])
    spheres = np.array([[0, 0, 0, 1.0], [2.2, 0, 0, 1.0], [0, -101, 0, 100.0], [-2.2, 0.3, 0.5, 0.7]])
    smat = np.array([0, 1, 2, 0], dtype=np.uint32)
    desc = oracle.SceneDesc(spheres, smat, mats, np.zeros((0, 6)), [], np.zeros((0, 4)))
    osc = oracle.Scene(desc)
    rows = desc.materials_array()[smat]
    gs = rtw.Scene.from_arrays(spheres, rows)
    for (w, h, spp, depth, defocus) in ((50, 35, 5, 50, 0.6), (17, 33, 1, 1, 0.0), (64, 48, 2, 3, 0.2)):
        cb = (rtw.CameraBuilder().with_image_width(w).with_image_height(h).with_samples_per_pixel(spp).with_max_depth(depth)
              .with_background((0.7, 0.8, 1.0)).with_vfov(35).with_lookfrom((3, 2, 7)).with_lookat((0, 0, 0)).with_vup((0, 1, 0))
              .with_focus_dist(7.5).with_defocus_angle(defocus))
        cam = cb.build()
        ocb = oracle.CameraBuilder.from_buffer_copy(cb.pod)       # same POD layout
        ocam = oracle.camera_build(ocb)
        assert list(ocam.pixel00) == list(cam.pod.pixel00_loc) and list(ocam.ddu) == list(cam.pod.defocus_disk_u)
        ii, jj = np.meshgrid(np.arange(w), np.arange(h))
        ss = (ii + jj) % spp
        o_o, d_o = oracle.get_rays(ocam, oracle.options(seed=9, rng_mode=oracle.W64), ii.ravel(), jj.ravel(), ss.ravel())
        o_g, d_g = cam.get_rays(ii.ravel(), jj.ravel(), ss.ravel(), rtw.RenderOptions(seed=9, precision=rtw.RTW_F64))
        assert np.array_equal(o_o, o_g) and np.array_equal(d_o, d_g)
        ref, _, cnt, _ = osc.render(ocam, oracle.options(seed=9, rng_mode=oracle.W64, math_mode=oracle.PORTABLE))
        got, _, st = gs.render(cam, rtw.RenderOptions(seed=9, precision=rtw.RTW_F64))
        assert np.array_equal(ref, got, equal_nan=True) and st["rays"] == cnt["rays"], (w, h, spp, depth)
        a, a8, sa = gs.render(cam, rtw.RenderOptions(seed=9, precision=rtw.RTW_F32, mode=rtw.RTW_WAVEFRONT))
        b, b8, sb = gs.render(cam, rtw.RenderOptions(seed=9, precision=rtw.RTW_F32, mode=rtw.RTW_MEGAKERNEL))
        assert np.array_equal(a, b, equal_nan=True) and sa["rays"] == sb["rays"]
        assert abs(a.mean() - ref.mean()) < 0.05 * spp
    gs.close()


def test_cpp_host_mirror_cli_matches_python(rtw, simple_scene, tmp_path):
    """`rtw_bin simple --backend cuda` (the reference's bin/src/main.rs flow in C++: scenes::simple ->
    CameraBuilder...build() -> Camera::render -> P3 writer) writes the same image, byte for byte, as the Python
    mirror of the same calls."""
    import os, subprocess
    exe = os.path.join(os.path.dirname(rtw.library_path()), "rtw_bin")
    out = tmp_path / "image.ppm"
    w, h, spp = 96, 54, 8
    r = subprocess.run([exe, "simple", "--backend", "cuda", "--width", str(w), "--height", str(h), "--spp", str(spp), "--depth", "50",
                        "--seed", str(SEED), "--out", str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    cam = (simple_scene["cb"].with_vfov(40.).with_aspect_ratio(w / h).with_max_depth(50).with_image_width(w).with_image_height(h)
           .with_samples_per_pixel(spp).build())
    sc = rtw.Scene(simple_scene["world"], simple_scene["lights"])
    _, rgb8, _ = sc.render(cam, rtw.RenderOptions(seed=SEED))
    sc.close()
    ref = tmp_path / "ref.ppm"
    rtw.write_ppm(str(ref), rgb8)
    assert out.read_text() == ref.read_text()
    lines = out.read_text().splitlines()
    assert lines[:3] == ["P3", f"{w} {h}", "255"] and len(lines) == 3 + w * h


def test_device_built_bvh_renders_the_same_image(rtw, oracle, simple_scene):
    """the world BVH built on the GPU (Morton codes + radix sort + Karras radix tree + bottom-up fit) instead
    of by the host SAH builder.  Hittable::hit does not depend on the tree: f64 hits are bit-exact against the oracle and the
    f64 image is bit-identical to the one rendered with the host-built tree."""
    cam, ocam = _camera(rtw, oracle, simple_scene, 64, 36, 4, 50)
    o, d = _ray_batch(rtw, oracle, simple_scene, 2000, 2000)
    prim_o, t_o, _ = simple_scene["oscene"].trace_batch(o, d)
    host = rtw.Scene(simple_scene["world"], simple_scene["lights"])
    rtw.set_bvh_builder(rtw.RTW_BVH_DEVICE_LBVH)
    try:
        dev = rtw.Scene(simple_scene["world"], simple_scene["lights"])
    finally:
        rtw.set_bvh_builder(rtw.RTW_BVH_AUTO)
    try:
        ih, idv = host.info(), dev.info()
        assert ih["builder"] == "host-sah" and idv["builder"] == "device-lbvh"
        assert idv["leaves"] >= 484 // 2 and 8 <= idv["depth"] <= 30 and idv["nodes"] + 1 == idv["leaves"]
        prim_g, t_g = dev.trace_batch(o, d, precision=rtw.RTW_F64)
        assert np.array_equal(prim_o, prim_g) and np.array_equal(t_o, t_g)
        opts = rtw.RenderOptions(seed=SEED, precision=rtw.RTW_F64)
        a, _, sa = host.render(cam, opts)
        b, _, sb = dev.render(cam, opts)
        assert np.array_equal(a, b, equal_nan=True) and sa["rays"] == sb["rays"]
        # FP32: the renderers read a different (conservative) set of boxes, grazing rays may differ; the image is the same to noise
        for mode in (rtw.RTW_WAVEFRONT, rtw.RTW_MEGAKERNEL):
            a32, _, _ = host.render(cam, rtw.RenderOptions(seed=SEED, mode=mode))
            b32, _, _ = dev.render(cam, rtw.RenderOptions(seed=SEED, mode=mode))
            same = np.isclose(a32, b32, rtol=1e-6, atol=1e-6, equal_nan=True).all(axis=2)
            assert same.mean() > 0.995, same.mean()
    finally:
        host.close(); dev.close()
    # a larger scene (6 400 spheres, deeper tree, duplicates in the coarse Morton bits)
    arrays = rtw.scenes.simple_arrays(SEED, 40)
    desc = oracle.scene_simple(SEED, 40)
    osc = oracle.Scene(desc)
    rtw.set_bvh_builder(rtw.RTW_BVH_DEVICE_LBVH)
    try:
        big = rtw.Scene.from_arrays(arrays["spheres"], arrays["sphere_materials"], arrays["planes"], arrays["plane_materials"], arrays["lights"])
    finally:
        rtw.set_bvh_builder(rtw.RTW_BVH_AUTO)
    try:
        assert big.info()["builder"] == "device-lbvh"
        rng = np.random.default_rng(4)
        oo = rng.uniform(-30, 30, (4000, 3)); oo[:, 1] = rng.uniform(0.05, 8, 4000)
        dd = rng.normal(size=(4000, 3))
        p_o, tt_o, _ = osc.trace_batch(oo, dd)
        p_g, tt_g = big.trace_batch(oo, dd, precision=rtw.RTW_F64)
        assert np.array_equal(p_o, p_g) and np.array_equal(tt_o, tt_g) and (p_o >= 1).sum() > 300
    finally:
        big.close()


def _check_exported_tree(nodes, order, prim_box, info):
    """Invariants of rtw_scene_export_bvh's flat tree: one Root, consistent parent / child links and depths, the leaves partition
    the leaf-ordered primitive list, every leaf box holds its primitives' boxes, every inner box is the union of its children's."""
    n = len(nodes)
    assert n == info["nodes"] + info["leaves"] and nodes["parent"][0] == -1 and (nodes["parent"][1:] >= 0).all()
    leaf = nodes["left"] < 0
    assert ((nodes["right"] < 0) == leaf).all() and leaf.sum() == info["leaves"]
    assert info["depth"] - 2 <= nodes["depth"].max() <= info["depth"] and nodes["depth"][0] == 0       # the LBVH reports its depth before leaf collapse
    inner = np.flatnonzero(~leaf)
    for side in ("left", "right"):
        ch = nodes[side][inner]
        assert (ch > inner).all() and (ch < n).all()                            # breadth-first: children follow their parent
        assert (nodes["parent"][ch] == inner).all() and (nodes["depth"][ch] == nodes["depth"][inner] + 1).all()
    assert len(np.unique(np.concatenate([nodes["left"][inner], nodes["right"][inner]]))) == n - 1      # every non-root node has one parent
    l, r = nodes["left"][inner], nodes["right"][inner]
    assert np.array_equal(nodes["box_min"][inner], np.minimum(nodes["box_min"][l], nodes["box_min"][r]))
    assert np.array_equal(nodes["box_max"][inner], np.maximum(nodes["box_max"][l], nodes["box_max"][r]))
    # leaves: disjoint ranges covering [0, n_prims), each primitive id once
    lf = np.flatnonzero(leaf)
    first, count = nodes["first"][lf].astype(np.int64), nodes["count"][lf].astype(np.int64)
    srt = np.argsort(first)
    assert first[srt][0] == 0 and np.array_equal(first[srt][1:], (first[srt] + count[srt])[:-1]) and (first[srt] + count[srt])[-1] == len(order)
    assert (count >= 1).all() and count.max() <= info["max_leaf"]
    assert len(np.unique(order)) == len(order)
    for k in lf:
        ids = order[nodes["first"][k]:nodes["first"][k] + nodes["count"][k]]
        mn, mx = prim_box(ids)
        assert (nodes["box_min"][k] <= mn.min(axis=0)).all() and (nodes["box_max"][k] >= mx.max(axis=0)).all()


def test_export_bvh_flat_host_mirror(rtw, oracle, simple_scene):
    """rtw_scene_export_bvh: the world BVH read back as the reference's flat `BVHNode::{Root, Node, Leaf}` records
    (hittable_collections/bvh.rs:224-241), for the host SAH tree, the device-built LBVH and a general scene."""
    arrays = rtw.scenes.simple_arrays(SEED, 11)
    sph = np.asarray(arrays["spheres"], dtype=np.float64).reshape(-1, 4)
    n_planes = len(np.asarray(arrays["planes"]).reshape(-1, 6))

    def sphere_box(ids):
        q = sph[np.asarray(ids, dtype=np.int64) - n_planes]
        return q[:, :3] - q[:, 3:4], q[:, :3] + q[:, 3:4]

    trees = {}
    for builder in (rtw.RTW_BVH_HOST_SAH, rtw.RTW_BVH_DEVICE_LBVH):
        rtw.set_bvh_builder(builder)
        try:
            sc = rtw.Scene.from_arrays(arrays["spheres"], arrays["sphere_materials"], arrays["planes"], arrays["plane_materials"], arrays["lights"])
        finally:
            rtw.set_bvh_builder(rtw.RTW_BVH_AUTO)
        try:
            nodes, order = sc.export_bvh()
            info = sc.info()
            assert len(order) == len(sph) and sorted(order.tolist()) == list(range(n_planes, n_planes + len(sph)))
            _check_exported_tree(nodes, order, sphere_box, info)
            trees[builder] = nodes
            # a brute-force walk of the exported tree finds the oracle's hits (first 64 rays)
            o, d = _ray_batch(rtw, oracle, simple_scene, 64, 0)
            prim_o, t_o, _ = simple_scene["oscene"].trace_batch(o, d)
            for k in range(len(o)):
                if prim_o[k] < n_planes:
                    continue
                inv = 1.0 / d[k]
                stack, found = [0], False
                while stack:
                    nd = nodes[stack.pop()]
                    t0, t1 = (nd["box_min"] - o[k]) * inv, (nd["box_max"] - o[k]) * inv
                    if np.minimum(t0, t1).max() > np.maximum(t0, t1).min() * (1 + 1e-12) + 1e-12:
                        continue
                    if nd["left"] < 0:
                        found |= int(prim_o[k]) in order[nd["first"]:nd["first"] + nd["count"]].tolist()
                    else:
                        stack += [int(nd["left"]), int(nd["right"])]
                assert found, k
        finally:
            sc.close()
    # both builders bound the same scene: equal root boxes
    a, b = trees[rtw.RTW_BVH_HOST_SAH][0], trees[rtw.RTW_BVH_DEVICE_LBVH][0]
    assert np.array_equal(a["box_min"], b["box_min"]) and np.array_equal(a["box_max"], b["box_max"])
    # capacity errors and count queries
    L = rtw._lib.load()
    import ctypes as C
    sc = rtw.Scene(simple_scene["world"], simple_scene["lights"])
    try:
        nn, npr = C.c_size_t(0), C.c_size_t(0)
        assert L.rtw_scene_export_bvh(sc._h, None, 0, C.byref(nn), None, 0, C.byref(npr)) == 0 and nn.value > 0 and npr.value == len(sph)
        buf = np.zeros(nn.value, dtype=rtw.Scene.BVH_NODE_DTYPE)
        assert L.rtw_scene_export_bvh(sc._h, buf.ctypes.data_as(C.c_void_p), nn.value - 1, C.byref(nn), None, 0, C.byref(npr)) == rtw._lib.RTW_E_INVALID
        assert L.rtw_scene_export_bvh(None, None, 0, C.byref(nn), None, 0, C.byref(npr)) == rtw._lib.RTW_E_INVALID
    finally:
        sc.close()
    # general scenes: cornell_box walks its 8 entries as a flat list (no tree); debugging_scene (14 entries) has one
    world, lights, _ = rtw.scenes.cornell_box()
    g = rtw.Scene(world, lights)
    try:
        nodes, order = g.export_bvh()
        assert len(nodes) == 0 and sorted(order.tolist()) == list(range(8))
    finally:
        g.close()
    world, lights, _ = rtw.scenes.debugging_scene(SEED)
    g = rtw.Scene(world, lights)
    try:
        nodes, order = g.export_bvh()
        info = g.info()
        assert len(nodes) == info["nodes"] + info["leaves"] and len(np.unique(order)) == len(order) and nodes["parent"][0] == -1
        leaf = nodes["left"] < 0
        assert nodes["count"][leaf].sum() == len(order)
    finally:
        g.close()


@pytest.mark.parametrize("mode", ("wavefront", "megakernel"))
def test_sample_partition_is_bit_identical(rtw, simple_scene, gscene, mode):
    """The multi-GPU sample partition on one GPU: three "ranks" render their sample ranges of every pixel into fixed-point
    accumulator blocks, the blocks are summed as integers (what the NCCL reduce does), resolved — and the result equals the
    single-launch image bit for bit (radiance sums, poisoned pixels, resolved bytes)."""
    import torch
    from ray_tracing_weekend_b200 import dist as D
    w, h, spp = 100, 70, 10
    cam = (simple_scene["cb"].with_vfov(40.).with_aspect_ratio(w / h).with_max_depth(50).with_image_width(w).with_image_height(h)
           .with_samples_per_pixel(spp).build())
    opts = rtw.RenderOptions(seed=SEED, precision=rtw.RTW_F32, mode=rtw.RTW_WAVEFRONT if mode == "wavefront" else rtw.RTW_MEGAKERNEL)
    one, one8, st1 = gscene.render(cam, opts)
    world = 3
    slots = D.tiles_total(w, h) * 256
    blocks = torch.zeros((world, D.accum_words(w, h)), dtype=torch.int64, device="cuda")
    rays = 0
    covered = []
    for r in range(world):
        b, c = D.sample_range(spp, r, world)
        covered += list(range(b, b + c))
        st = gscene.render_samples_device(cam, opts, b, c, blocks[r].data_ptr(), blocks[r].data_ptr() + 8 * 3 * slots)
        rays += st["rays"]
        assert st["paths"] == w * h * c
    assert covered == list(range(spp)) and rays == st1["rays"]
    total = blocks.sum(dim=0)
    out = torch.zeros((h, w, 3), dtype=torch.float64, device="cuda")
    out8 = torch.zeros((h, w, 3), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    rtw.resolve_accum_device(total.data_ptr(), total.data_ptr() + 8 * 3 * slots, w, h, spp, out.data_ptr(), out8.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(one, out.cpu().numpy(), equal_nan=True)
    assert np.array_equal(one8, out8.cpu().numpy())
    assert np.isnan(one).any(), "the frame has poisoned pixels, so the flag fields were exercised"


def test_full_size_frame_properties(rtw, simple_scene, gscene):
    """BASELINE config C2 at its full size (1920x1080, 500 spp, depth 50 — 1.04 G paths, far beyond what the oracle can trace in a
    test) through size-independent properties: the two FP32 renderers and the 4-way sample partition produce the same image bit
    for bit (a checksum of every accumulator), the ray count is the deterministic 2.718 G of this seed, every pixel received
    exactly 500 samples' worth of finite-or-poisoned radiance, and the background pixels equal spp exactly."""
    import torch
    from ray_tracing_weekend_b200 import dist as D
    w, h, spp = 1920, 1080, 500
    cam = (simple_scene["cb"].with_vfov(40.).with_aspect_ratio(w / h).with_max_depth(50).with_image_width(w).with_image_height(h)
           .with_samples_per_pixel(spp).build())
    a, a8, sa = gscene.render(cam, rtw.RenderOptions(seed=SEED, mode=rtw.RTW_WAVEFRONT))
    b, b8, sb = gscene.render(cam, rtw.RenderOptions(seed=SEED, mode=rtw.RTW_MEGAKERNEL))
    assert sa["paths"] == sb["paths"] == w * h * spp == 1_036_800_000
    assert sa["rays"] == sb["rays"] and 2.70e9 < sa["rays"] < 2.74e9
    assert np.array_equal(a, b, equal_nan=True) and np.array_equal(a8, b8)
    # the one-sided ground plane is invisible from above: most of the frame is pure background (1, 1, 1) * spp
    bg = (a == float(spp)).all(axis=2)
    assert 0.6 < bg.mean() < 0.9
    assert np.array_equal(a8[bg], np.full((int(bg.sum()), 3), 255, dtype=np.uint8))
    poisoned = np.isnan(a).any(axis=2)
    assert 0.05 < poisoned.mean() < 0.35 and (a8[poisoned].min(axis=1) == 0).all()     # NaN -> 0 like `as u8` (colour.rs:24-35)
    # white background and albedos <= 1: the expected radiance is <= 1 per sample (single mixture-pdf samples can weigh up to 2 x albedo)
    # (fireflies: a sampled direction with a tiny pdf gives one huge sample, so there is no per-pixel upper bound)
    fin = np.isfinite(a).all(axis=2)
    assert (a[fin] >= 0).all() and np.median(a[fin]) <= spp and np.quantile(a[fin], 0.99) <= 1.05 * spp
    # sample partition over 4 "ranks": integer sums of the accumulator blocks resolve to the same image
    slots = D.tiles_total(w, h) * 256
    total = torch.zeros(D.accum_words(w, h), dtype=torch.int64, device="cuda")
    block = torch.zeros_like(total)
    rays = 0
    for r in range(4):
        s0, c = D.sample_range(spp, r, 4)
        st = gscene.render_samples_device(cam, rtw.RenderOptions(seed=SEED), s0, c, block.data_ptr(), block.data_ptr() + 8 * 3 * slots)
        rays += st["rays"]
        total += block
    out = torch.zeros((h, w, 3), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    rtw.resolve_accum_device(total.data_ptr(), total.data_ptr() + 8 * 3 * slots, w, h, spp, out.data_ptr(), 0)
    torch.cuda.synchronize()
    assert rays == sa["rays"] and np.array_equal(a, out.cpu().numpy(), equal_nan=True)


def test_progressive_rendering_with_checkpoints(rtw, simple_scene, gscene, tmp_path):
    """the frame rendered in passes into host accumulators that are saved to disk and restored in between equals
    the one-shot image bit for bit; the C++ CLI does the same with --passes / --checkpoint / --resume and the other output formats."""
    import os, subprocess
    w, h, spp = 96, 54, 20
    cam = (simple_scene["cb"].with_vfov(40.).with_aspect_ratio(w / h).with_max_depth(50).with_image_width(w).with_image_height(h)
           .with_samples_per_pixel(spp).build())
    opts = rtw.RenderOptions(seed=SEED)
    one, one8, st1 = gscene.render(cam, opts)
    accum, poison = rtw.new_accumulators(w, h)
    rays = 0
    for k, (b, c) in enumerate(((13, 7), (0, 5), (5, 8))):                 # any order, any split
        rays += gscene.render_samples(cam, opts, b, c, accum, poison)["rays"]
        np.savez(tmp_path / "ck.npz", accum=accum, poison=poison)          # checkpoint ...
        ck = np.load(tmp_path / "ck.npz")
        accum, poison = np.ascontiguousarray(ck["accum"]), np.ascontiguousarray(ck["poison"])      # ... and restore
    got, got8 = rtw.resolve_accum(accum, poison, w, h, spp)
    assert rays == st1["rays"] and np.array_equal(one, got, equal_nan=True) and np.array_equal(one8, got8)
    # the C++ CLI: an interrupted progressive render resumed from its checkpoint == the one-shot render; P6 and PNG carry the same pixels
    exe = os.path.join(os.path.dirname(rtw.library_path()), "rtw_bin")
    common = [exe, "simple", "--backend", "cuda", "--width", str(w), "--height", str(h), "--spp", str(spp), "--depth", "50", "--seed", str(SEED)]
    ref, part, fin = tmp_path / "ref.ppm", tmp_path / "part.ppm", tmp_path / "fin.ppm"
    for args in (["--out", str(ref)],
                 ["--passes", "4", "--checkpoint", str(tmp_path / "ck.bin"), "--stop-after", "2", "--out", str(part)],
                 ["--passes", "4", "--checkpoint", str(tmp_path / "ck.bin"), "--resume", "--out", str(fin)],
                 ["--format", "p6", "--out", str(tmp_path / "b.ppm")], ["--format", "png", "--out", str(tmp_path / "c.png")]):
        r = subprocess.run(common + args, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    assert ref.read_text() == fin.read_text() and ref.read_text() != part.read_text()
    want = np.array(ref.read_text().split()[4:], dtype=np.uint8).reshape(h, w, 3)
    assert np.array_equal(want, one8[::-1])
    p6 = (tmp_path / "b.ppm").read_bytes()
    head = f"P6\n{w} {h}\n255\n".encode()
    assert p6.startswith(head) and np.array_equal(np.frombuffer(p6[len(head):], dtype=np.uint8).reshape(h, w, 3), want)
    Image = pytest.importorskip("PIL.Image")
    assert np.array_equal(np.asarray(Image.open(tmp_path / "c.png").convert("RGB")), want)
    r = subprocess.run(common + ["--passes", "4", "--checkpoint", str(tmp_path / "ck.bin"), "--resume", "--spp", "21"], capture_output=True, text=True)
    assert r.returncode != 0 and "another render" in r.stderr             # a checkpoint of a different render is refused
