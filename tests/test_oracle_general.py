"""CPU tests of the general-scene oracle (oracle/rtw_oracle_general.hpp: Quad, Triangle, Cuboid, Transformed<T>,
DiffuseLight, Isotropic, NoiseTexture) and of the product's host helpers for those scenes.

Pinned by: the reference's OWN known-answer tests for Transformation::inverse (geometry/src/transformations.rs:131-160),
closed-form geometry, agreement with the sphere-only oracle on scenes::simple, and the committed golden fixture
tests/golden/general_oracle.json (made by tests/golden/make_golden_general.py)."""
import ctypes as C
import json
import math
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
SEED = 20261018


def test_reference_kat_inverse_times_itself_is_identity(oracle, rtw):
    """geometry/src/transformations.rs:131-149, verbatim: |id - I| < f64::EPSILON on the diagonal and the translation."""
    mat = [2., -1., 1., 1., 1., 1., 1., 1., 2.]
    for impl in ("oracle", "product"):
        if impl == "oracle":
            trans = oracle.transform_then(oracle.transform(rotation=mat), oracle.transform(translation=(0.5, 2., -1.)))
            inv = oracle.transform_inverse(trans)
            ident = oracle.transform_then(trans, inv)
            rot, tr = list(ident.rotation), list(ident.translation)
        else:
            trans = rtw.Transformation(rotation=tuple(mat)).then(rtw.Translation3(0.5, 2., -1.))
            ident = trans.then(trans.inverse())
            rot, tr = list(ident.rotation), list(ident.translation)
        for i in range(3):
            assert abs(rot[4 * i] - 1.) < 2.220446049250313e-16, impl
            assert abs(tr[i]) < 2.220446049250313e-16, impl


def test_reference_kat_inverse_of_identity_is_identity(oracle, rtw):
    """geometry/src/transformations.rs:151-160."""
    inv = oracle.transform_inverse(oracle.transform())
    assert list(inv.rotation) == [1., 0., 0., 0., 1., 0., 0., 0., 1.]
    assert rtw.Transformation().inverse().rotation == (1., 0., 0., 0., 1., 0., 0., 0., 1.)
    # a singular matrix has no inverse (matrix3.rs:14-17): "just say it's not hit"
    assert oracle.transform_inverse(oracle.transform(rotation=[1, 2, 3, 2, 4, 6, 0, 0, 1])) is None
    assert rtw.Transformation(rotation=(1., 2., 3., 2., 4., 6., 0., 0., 1.)).inverse() is None


def test_host_transform_helpers_match_oracle(oracle, rtw):
    rng = np.random.default_rng(3)
    for _ in range(40):
        a = rtw.Transformation(tuple(rng.normal(size=9)), tuple(rng.normal(size=3)))
        b = rtw.rotation(float(rng.uniform(-180, 180)), int(rng.integers(0, 3))).then(rtw.Translation3(*rng.normal(size=3)))
        oa, ob = oracle.transform(a.rotation, a.translation), oracle.transform(b.rotation, b.translation)
        ab, oab = a.then(b), oracle.transform_then(oa, ob)
        assert list(ab.rotation) == list(oab.rotation) and list(ab.translation) == list(oab.translation)
        ia, oia = a.inverse(), oracle.transform_inverse(oa)
        assert list(ia.rotation) == list(oia.rotation) and list(ia.translation) == list(oia.translation)
    for ax in range(3):
        r, o = rtw.rotation(37.5, ax), oracle.rotation(37.5, ax)
        assert list(r.rotation) == list(o.rotation)
    # rotation about Y (transformations.rs:51-55): [[c, 0, s], [0, 1, 0], [-s, 0, c]]
    r = rtw.rotation(90.0, rtw.Axis.Y).rotation
    assert abs(r[2] - 1.) < 1e-15 and abs(r[6] + 1.) < 1e-15 and r[4] == 1.


def test_perlin_tables_and_noise(oracle, rtw):
    t = oracle.perlin_generate(SEED, 0)
    p = rtw.NoiseTexture(4.0, SEED, 0).perlin()
    assert bytes(t) == bytes(p), "product host Perlin::new == oracle Perlin::new"
    for perm in (t.perm_x, t.perm_y, t.perm_z):
        assert sorted(perm) == list(range(256))
    v = np.array([list(r) for r in t.rand_vec])
    assert (np.sum(v * v, axis=1) < 1.).all() and abs(v.mean()) < 0.1          # UnitSphere: inside the unit ball
    assert bytes(oracle.perlin_generate(SEED, 1)) != bytes(t)
    # noise vanishes on the integer lattice (all eight weight vectors are lattice offsets, the (0,0,0) corner has weight 1)
    for q in ((0., 0., 0.), (3., -2., 7.), (-255., 256., 1000.)):
        assert oracle.perlin_turb(t, q, 0) == 0.
    rng = np.random.default_rng(1)
    vals = np.array([oracle.perlin_turb(t, rng.uniform(-50, 50, 3), 7) for _ in range(2000)])
    assert np.abs(vals).max() < 2. and vals.std() > 0.05
    # NoiseTexture's sine: the portable sequence is libm's sine to 1 ulp-ish over the range the texture feeds it
    for x in np.linspace(-300., 300., 20001):
        assert abs(oracle.sin_portable(x) - math.sin(x)) < 4e-16


def _quad_scene(rtw, oracle, items, lights=None):
    world = rtw.HittableList(); world.extend(items)
    li = rtw.HittableList(); li.extend(lights or [rtw.Sphere((0., 50., 0.), 1., rtw.INVISIBLE)])
    d = rtw.SceneDescription(world, li)
    return d, oracle.GScene(d.pod, d)


def test_quad_triangle_cuboid_closed_forms(oracle, rtw):
    white = rtw.Lambertian((0.5, 0.5, 0.5))
    d, g = _quad_scene(rtw, oracle, [rtw.Quad((0., 0., 0.), (2., 0., 0.), (0., 2., 0.), white),
                                     rtw.Triangle((10., 0., 0.), (2., 0., 0.), (0., 2., 0.), white),
                                     rtw.Cuboid((20., 0., 0.), (21., 1., 1.), white)])
    o = np.array([[0.5, 0.5, 5.], [1.9, 1.9, 5.], [2.1, 1., 5.], [10.5, 0.5, 5.], [11.9, 1.9, 5.], [20.5, 0.5, 5.], [20.5, 0.5, -5.], [20.5, 5., 0.5]])
    dr = np.array([[0., 0., -1.]] * 6 + [[0., 0., 1.], [0., -1., 0.]])
    prim, t, p, n = g.trace_batch(o, dr)
    assert list(prim) == [0, 0, -1, 1, -1, 2, 2, 2]                   # inside / inside / outside; triangle: u + v <= 1 only
    assert np.allclose(t[[0, 1, 3]], 5.) and np.allclose(t[5], 4.) and np.allclose(t[6], 5.) and np.allclose(t[7], 4.)
    assert np.allclose(n[0], [0, 0, 1]) and np.allclose(n[3], [0, 0, 2]), "the triangle's normal is n / (|n| / 2): length 2 (triangles.rs:44-45)"
    assert np.allclose(n[5], [0, 0, 1]) and np.allclose(n[6], [0, 0, -1]) and np.allclose(n[7], [0, 1, 0])
    # boxes: Quad::new pads the flat axis by 1e-4 once (aabox.rs:129-149)
    assert np.allclose(g.prim_box(0), [0, 0, -1e-4, 2, 2, 1e-4]) and np.allclose(g.prim_box(2), [20, 0, 0, 21, 1, 1], atol=2e-4)
    # Quad::pdf_value = distance^2 / (cos * area) (quadrilateral.rs:100-112), head-on from 5 away onto a 2x2 quad
    li = [rtw.Quad((0., 0., 0.), (2., 0., 0.), (0., 2., 0.), white)]
    d2, g2 = _quad_scene(rtw, oracle, [rtw.Sphere((1., 1., 5.), 0.5, white)], li)
    opts = oracle.options(seed=SEED)
    # a ray from below hits the sphere bottom at (1, 1, 4.5); lights.pdf_value there for the sampled direction is checked via the weight
    r = g2.scatter_batch(np.array([[1., 1., 0.1]]), np.array([[0., 0., 1.]]), [0], [0], [1], opts)
    assert r["kind"][0] == oracle.V_DIFFUSE and np.allclose(r["p"][0], [1, 1, 4.5])
    dirv = r["dir"][0]
    nd = dirv / np.linalg.norm(dirv)
    cos_n = max(0., float(np.dot(nd, r["normal"][0])))
    light = 0.
    if dirv[2] < 0:                                                   # towards the quad plane z = 0
        tq = -4.5 / dirv[2]; hit = np.array([1, 1, 4.5]) + tq * dirv
        if 0 <= hit[0] <= 2 and 0 <= hit[1] <= 2:
            light = (tq * tq * np.dot(dirv, dirv)) / (abs(dirv[2]) / np.linalg.norm(dirv) * 4.)
    want = 0.5 * (cos_n / math.pi) / (0.5 * light + 0.5 * cos_n / math.pi)
    assert np.allclose(r["weight"][0], want, rtol=1e-12)


def test_transformed_uses_the_reference_arithmetic(oracle, rtw):
    """Transformed<T>::hit (entities/transformations.rs:14-29): the ray DIRECTION also receives the inverse translation
    (transform_vector3d, geometry/src/transformations.rs:123-126), p is mapped back, the normal is not."""
    white = rtw.Lambertian((0.5, 0.5, 0.5))
    q = rtw.Quad((0., 0., 0.), (2., 0., 0.), (0., 2., 0.), white)
    d, g = _quad_scene(rtw, oracle, [q.transform(rtw.Translation3(0., 0., -3.))])
    # instance ray: origin (1, 1, 5) - (0,0,-3) = (1, 1, 8); direction (0,0,-1) + (0,0,3) = (0,0,2): it points AWAY from the quad
    prim, t, p, n = g.trace_batch(np.array([[1., 1., 5.]]), np.array([[0., 0., -1.]]))
    assert prim[0] == -1
    # a pure rotation has no translation term: behaves like a rigid transform
    d, g = _quad_scene(rtw, oracle, [q.transform(rtw.rotation(90., rtw.Axis.Y))])
    prim, t, p, n = g.trace_batch(np.array([[5., 1., -1.]]), np.array([[-1., 0., 0.]]))
    assert prim[0] == 0 and np.allclose(t[0], 5.) and np.allclose(p[0], [0., 1., -1.], atol=1e-12)
    assert np.allclose(np.abs(n[0]), [0, 0, 1]), "normal stays in instance space"
    # lights: a Transformed<T> keeps Hittable's defaults: pdf_value 0, random (1, 0, 0) (hittable.rs:175-181)
    li = [rtw.Cuboid((0., 0., 0.), (1., 1., 1.), rtw.DiffuseLight((1., 0., 0.))).transform(rtw.Translation3(0., 5., 0.))]
    d, g = _quad_scene(rtw, oracle, [rtw.Sphere((0., 0., 0.), 1., white)], li)
    hits = 0
    for s in range(64):
        r = g.scatter_batch(np.array([[0., 0., 5.]]), np.array([[0., 0., -1.]]), [0], [s], [1], oracle.options(seed=SEED))
        if np.array_equal(r["dir"][0], [1., 0., 0.]):
            hits += 1
    assert 16 <= hits <= 48                                            # half of the mixture draws


def test_general_oracle_equals_sphere_oracle_on_simple(oracle, rtw, simple_scene):
    """scenes::simple described as a general scene gives bit-identical hits, vertices and path radiance."""
    world, lights = simple_scene["world"], simple_scene["lights"]
    d = rtw.SceneDescription(world, lights)
    g = oracle.GScene(d.pod, d)
    s = simple_scene["oscene"]
    cam = oracle.camera_for(simple_scene["desc"], 64, 36, 4, 50)
    opts = oracle.options(seed=SEED, math_mode=oracle.PORTABLE)
    rng = np.random.default_rng(5)
    n = 600
    i, j, sm = rng.integers(0, 64, n), rng.integers(0, 36, n), rng.integers(0, 4, n)
    o, dr = oracle.get_rays(cam, opts, i, j, sm)
    p0, t0, _ = s.trace_batch(o, dr)
    p1, t1, _, _ = g.trace_batch(o, dr)
    assert np.array_equal(p0, p1) and np.array_equal(t0, t1) and (p0 >= 0).sum() > 100
    v0 = s.scatter_batch(o, dr, i, sm, np.ones(n), opts)
    v1 = g.scatter_batch(o, dr, i, sm, np.ones(n), opts)
    for k in ("prim", "t", "kind", "p", "normal", "dir", "weight"):
        assert np.array_equal(v0[k], v1[k], equal_nan=True), k
    r0 = s.path_radiance(cam, opts, i, j, sm)
    r1 = g.path_radiance(cam, opts, i, j, sm)
    assert np.array_equal(r0, r1, equal_nan=True)


def test_emission_and_isotropic(oracle, rtw):
    # a DiffuseLight sphere seen directly: radiance = its colour, one ray per path (material.rs:506-514, camera.rs:484-486)
    world = rtw.HittableList(); world.add(rtw.Sphere((0., 0., -5.), 2., rtw.DiffuseLight((3., 2., 1.))))
    li = rtw.HittableList(); li.add(rtw.Sphere((0., 0., -5.), 2., rtw.DiffuseLight((3., 2., 1.))))
    d = rtw.SceneDescription(world, li)
    g = oracle.GScene(d.pod, d)
    cb = rtw.CameraBuilder().with_image_width(8).with_image_height(8).with_samples_per_pixel(2).with_max_depth(5).with_vfov(20.)
    cam = oracle.Camera.from_buffer_copy(cb.build().pod)
    img, _, cnt, _ = g.render(cam, oracle.options(seed=SEED))
    assert np.array_equal(img[4, 4], [6., 4., 2.]) and cnt["rays"] == cnt["paths"]
    # Isotropic (material.rs:529-554): attenuation * (1/4pi) / (0.5 light + 0.5 / 4pi); a miss of the light gives attenuation * 2
    world = rtw.HittableList(); world.add(rtw.Sphere((0., 0., -5.), 2., rtw.Isotropic((0.25, 0.5, 1.))))
    li = rtw.HittableList(); li.add(rtw.Sphere((0., 100., 0.), 1., rtw.INVISIBLE))
    d = rtw.SceneDescription(world, li)
    g = oracle.GScene(d.pod, d)
    seen = set()
    for s in range(40):
        r = g.scatter_batch(np.array([[0., 0., 0.]]), np.array([[0., 0., -1.]]), [0], [s], [1], oracle.options(seed=SEED))
        assert r["kind"][0] == oracle.V_DIFFUSE
        w = r["weight"][0]
        if np.allclose(w, [0.5, 1., 2.]):
            seen.add("miss")                                           # SpherePdf direction that misses the light
        else:
            seen.add("light")
            assert (w < [0.5, 1., 2.]).all()
    assert "miss" in seen and "light" in seen


def test_portable_atan2_acos_and_checker(oracle, rtw):
    """Sphere::get_sphere_uv (sphere.rs:49-54) feeds CheckerTexture (texture.rs:46-55).  atan2 is the `libm` crate's (musl / msun)
    sequence, acos msun's: both within 1 ulp of this platform's libm, exact at the special points."""
    rng = np.random.default_rng(0)
    for _ in range(20000):
        y, x = rng.normal(size=2) * 10 ** rng.uniform(-3, 3)
        assert abs(oracle.atan2_msun(y, x) - math.atan2(y, x)) <= 2.3e-16 * abs(math.atan2(y, x))
    for x in np.concatenate([np.linspace(-1, 1, 20001), rng.uniform(-1, 1, 20000)]):
        assert abs(oracle.acos_msun(x) - math.acos(x)) <= 2.3e-16 * abs(math.acos(x))
    assert oracle.acos_msun(1.) == 0. and oracle.acos_msun(-1.) == math.pi and oracle.atan2_msun(0., -1.) == math.pi
    assert oracle.atan2_msun(1., 0.) == math.pi / 2 and math.isnan(oracle.acos_msun(1.5))
    # a checker on a quad: colour = even iff floor(u / scale) + floor(v / scale) is even, (u, v) the quad coordinates
    chk = rtw.Lambertian(rtw.CheckerTexture.new_with_colours((1., 0., 0.), (0., 0., 1.), 0.25))
    world = rtw.HittableList(); world.add(rtw.Quad((0., 0., 0.), (1., 0., 0.), (0., 1., 0.), chk))
    li = rtw.HittableList(); li.add(rtw.Sphere((0., 0., 50.), 1., rtw.INVISIBLE))
    d = rtw.SceneDescription(world, li)
    g = oracle.GScene(d.pod, d)
    uv = rng.uniform(0.01, 0.99, (400, 2))
    o = np.column_stack([uv, np.full(400, 3.)]); dr = np.tile([0., 0., -1.], (400, 1))
    seen = 0
    for k in range(400):
        # cosine-sampled vertices (weight = 2 * albedo when the light is missed) reveal the texture colour
        r = g.scatter_batch(o[k:k + 1], dr[k:k + 1], [k], [0], [1], oracle.options(seed=SEED))
        w = r["weight"][0]
        if np.allclose(w.sum(), 2.):
            even = (math.floor(uv[k, 0] / 0.25) + math.floor(uv[k, 1] / 0.25)) % 2 == 0
            assert np.allclose(w, [2., 0., 0.] if even else [0., 0., 2.])
            seen += 1
    assert seen > 100
    # on a sphere: u = atan2(-z, x) / tau, v = acos(y) / pi of the outward normal
    world = rtw.HittableList(); world.add(rtw.Sphere((0., 0., 0.), 2., chk))
    d = rtw.SceneDescription(world, li)
    g = oracle.GScene(d.pod, d)
    seen = 0
    for k in range(400):
        n = rng.normal(size=3); n /= np.linalg.norm(n)
        r = g.scatter_batch(np.array([n * 5.]), np.array([-n]), [k], [0], [1], oracle.options(seed=SEED))
        w = r["weight"][0]
        if np.allclose(w.sum(), 2.):
            nn = r["normal"][0]
            u, v = math.atan2(-nn[2], nn[0]) / math.tau, math.acos(nn[1]) / math.pi
            if min(abs((u / 0.25) % 1), abs((v / 0.25) % 1), 1 - abs((u / 0.25) % 1), 1 - abs((v / 0.25) % 1)) < 1e-6:
                continue
            even = (math.floor(u / 0.25) + math.floor(v / 0.25)) % 2 == 0
            assert np.allclose(w, [2., 0., 0.] if even else [0., 0., 2.])
            seen += 1
    assert seen > 100


def test_plane_uv_of_a_tilted_plane_and_nested_checker(oracle, rtw):
    """Plane::get_plane_uv (plane.rs:41-55) for a normal other than +y: Rodrigues' rotation of (p - point) about normal x (+y) by the
    angle between them, then the fractional parts of x and z — recomputed here in numpy and compared with the colour the oracle's
    checker picks.  Then a CheckerTexture whose `even` is a finer CheckerTexture (get_colour recurses, texture.rs:46-55)."""
    rng = np.random.default_rng(3)
    li = rtw.HittableList(); li.add(rtw.Sphere((0., 500., 0.), 1., rtw.INVISIBLE))
    scale = 0.3
    chk = rtw.Lambertian(rtw.CheckerTexture.new_with_colours((1., 0., 0.), (0., 0., 1.), scale))
    point, normal = np.array([0.5, -1., 0.25]), np.array([0.3, -0.8, 0.45])
    world = rtw.HittableList(); world.add(rtw.Plane(tuple(point), tuple(normal), chk))
    d = rtw.SceneDescription(world, li)
    g = oracle.GScene(d.pod, d)
    n = normal / np.linalg.norm(normal)
    up = np.array([0., 1., 0.])
    theta = math.atan2(np.linalg.norm(np.cross(n, up)), n @ up)
    k = np.cross(n, up) / np.linalg.norm(np.cross(n, up))
    seen = 0
    for s in range(600):
        o = point - 3. * n + rng.normal(size=3)                       # behind the one-sided plane: the ray must travel along +normal
        dr = n + 0.3 * rng.normal(size=3)
        r = g.scatter_batch(o[None], dr[None], [s], [0], [1], oracle.options(seed=SEED))
        if r["kind"][0] != oracle.V_DIFFUSE or not np.allclose(r["weight"][0].sum(), 2.):
            continue
        w = r["p"][0] - point
        rot = w * math.cos(theta) + np.cross(k, w) * math.sin(theta) + k * (k @ w) * (1. - math.cos(theta))
        assert abs(rot[1]) < 1e-9                                     # the plane has been rotated into y = 0
        u, v = math.fmod(rot[0], 1.), math.fmod(rot[2], 1.)           # f64::fract keeps the sign
        if min(abs((u / scale) % 1), abs((v / scale) % 1), 1 - abs((u / scale) % 1), 1 - abs((v / scale) % 1)) < 1e-6:
            continue
        even = (math.floor(u / scale) + math.floor(v / scale)) % 2 == 0
        assert np.allclose(r["weight"][0], [2., 0., 0.] if even else [0., 0., 2.])
        seen += 1
    assert seen > 100
    # nested: even = a finer checker (green / white), odd = blue
    fine = rtw.CheckerTexture.new_with_colours((0., 1., 0.), (1., 1., 1.), 0.05)
    nested = rtw.Lambertian(rtw.CheckerTexture(fine, (0., 0., 1.), 0.25))
    world = rtw.HittableList(); world.add(rtw.Quad((0., 0., 0.), (1., 0., 0.), (0., 1., 0.), nested))
    d = rtw.SceneDescription(world, li)
    assert d.pod.n_textures == 2                                      # the sub-texture precedes its parent in the table
    g = oracle.GScene(d.pod, d)
    uv = rng.uniform(0.01, 0.99, (500, 2))
    seen = set()
    for q in range(500):
        r = g.scatter_batch(np.array([[uv[q, 0], uv[q, 1], 3.]]), np.array([[0., 0., -1.]]), [q], [0], [1], oracle.options(seed=SEED))
        w = r["weight"][0]
        if r["kind"][0] != oracle.V_DIFFUSE or not np.allclose(w.max(), 2.):
            continue
        cell = lambda sc: (math.floor(uv[q, 0] / sc) + math.floor(uv[q, 1] / sc)) % 2 == 0
        if min(abs((uv[q] / 0.05) % 1).min(), (1 - abs((uv[q] / 0.05) % 1)).min()) < 1e-6:
            continue
        want = ([0., 2., 0.] if cell(0.05) else [2., 2., 2.]) if cell(0.25) else [0., 0., 2.]
        assert np.allclose(w, want), (uv[q], w, want)
        seen.add(tuple(want))
    assert len(seen) == 3


def test_general_golden_fixture(oracle, rtw):
    """Committed outputs of the general oracle on the reference's scenes (tests/golden/make_golden_general.py)."""
    with open(os.path.join(GOLDEN, "general_oracle.json")) as f:
        gold = json.load(f)
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_general", os.path.join(GOLDEN, "make_golden_general.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    compute = mod.compute
    got = compute(oracle, rtw)
    assert got == gold
