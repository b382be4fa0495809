"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes binding of oracle/_build/liboracle.so (the C++ f64 restatement of the reference hot path,
see rtw_oracle.hpp).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; the product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

W64, W32 = 0, 1
LIBM, PORTABLE = 0, 1
LAMBERTIAN, METAL, DIELECTRIC, INVISIBLE = 0, 1, 2, 3
V_MISS, V_ABSORB, V_SPECULAR, V_DIFFUSE = 0, 1, 2, 3
EPS = 2.220446049250313e-16


def _source_hash() -> str:
    import hashlib
    h = hashlib.sha256()
    for f in ("rtw_oracle.hpp", "rtw_oracle_general.hpp", "oracle_capi.cpp", "oracle_general_capi.cpp", "oracle_cli.cpp", "Makefile"):
        with open(os.path.join(_HERE, f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def build(force: bool = False) -> str:
    """Compile the restatement with g++ (make); returns the .so path.  Staleness is decided by source content
    (the built library travels inside repo snapshots whose mtimes mean nothing)."""
    import fcntl
    stamp = os.path.join(_HERE, "_build", ".source_hash")
    def stale():
        if not os.path.exists(_LIB_PATH) or not os.path.exists(stamp):
            return True
        with open(stamp) as f:
            return f.read().strip() != _source_hash()
    if force or stale():
        os.makedirs(os.path.join(_HERE, "_build"), exist_ok=True)
        with open(os.path.join(_HERE, "_build", ".build_lock"), "w") as lock:
            fcntl.flock(lock, fcntl.LOCK_EX)
            if force or stale():
                subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True)
                with open(stamp, "w") as f:
                    f.write(_source_hash())
    return _LIB_PATH


class Material(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("r", C.c_double), ("g", C.c_double), ("b", C.c_double), ("param", C.c_double)]


class Camera(C.Structure):
    _fields_ = [(n, C.c_double * 3) for n in ("center", "pixel00", "du", "dv", "ddu", "ddv", "background")] + [
        ("defocus_angle", C.c_double), ("width", C.c_uint32), ("height", C.c_uint32), ("spp", C.c_uint32),
        ("max_depth", C.c_uint32)]


class CameraBuilder(C.Structure):
    _fields_ = [("aspect_ratio", C.c_double), ("has_aspect", C.c_uint32), ("width", C.c_uint32), ("has_width", C.c_uint32),
                ("height", C.c_uint32), ("has_height", C.c_uint32), ("spp", C.c_uint32), ("max_depth", C.c_uint32),
                ("background", C.c_double * 3), ("vfov", C.c_double), ("lookfrom", C.c_double * 3),
                ("lookat", C.c_double * 3), ("vup", C.c_double * 3), ("defocus_angle", C.c_double),
                ("focus_dist", C.c_double)]


class Options(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("tmin", C.c_double), ("rng_mode", C.c_uint32), ("math_mode", C.c_uint32),
                ("faithful_bvh", C.c_uint32), ("threads", C.c_int32), ("fix_nan", C.c_uint32), ("reserved", C.c_uint32)]


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("rays", "paths", "box_tests", "box_builds", "node_visits", "sphere_tests",
                                          "plane_tests", "light_tests", "lambertian", "metal", "dielectric", "absorbed",
                                          "missed", "depth_out")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class BvhStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("nodes", "leaves", "depth", "max_leaf", "prims")]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_uniform.restype = C.c_double
        L.orc_uniform.argtypes = [C.c_uint64] + [C.c_uint32] * 6
        L.orc_uniform_index.restype = C.c_uint32
        L.orc_uniform_index.argtypes = [C.c_uint64] + [C.c_uint32] * 6
        L.orc_sincos.argtypes = [C.c_double, C.c_uint32, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.orc_desc_simple.restype = C.c_void_p
        L.orc_desc_simple.argtypes = [C.c_uint64, C.c_int32, C.c_double, C.c_double, C.c_int32]
        L.orc_desc_destroy.argtypes = [C.c_void_p]
        L.orc_desc_counts.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_desc_copy.argtypes = [C.c_void_p] * 8
        L.orc_scene_create.restype = C.c_void_p
        L.orc_scene_create.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p,
                                       C.c_void_p, C.c_uint64, C.c_void_p]
        L.orc_scene_destroy.argtypes = [C.c_void_p]
        L.orc_scene_bvh_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_trace_batch.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p,
                                      C.c_void_p, C.c_uint32, C.c_void_p]
        L.orc_scatter_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64] + [C.c_void_p] * 12
        L.orc_shade_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64] + [C.c_void_p] * 12
        L.orc_get_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64] + [C.c_void_p] * 5
        L.orc_render.restype = C.c_double
        L.orc_render.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        L.orc_path_radiance.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64] + [C.c_void_p] * 4
        L.orc_resolve.argtypes = [C.c_void_p, C.c_uint64, C.c_int32, C.c_void_p]
        L.orc_hardware_threads.restype = C.c_uint32
        L.orc_camera_build.argtypes = [C.c_void_p, C.c_void_p]
        # general scenes (rtw_oracle_general.hpp)
        L.orc_gscene_create.restype = C.c_void_p
        L.orc_gscene_create.argtypes = [C.c_void_p]
        L.orc_gscene_destroy.argtypes = [C.c_void_p]
        L.orc_gtrace_batch.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_double, C.c_double] + [C.c_void_p] * 4
        L.orc_gscatter_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64] + [C.c_void_p] * 13
        L.orc_gpath_radiance.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64] + [C.c_void_p] * 4
        L.orc_grender.restype = C.c_double
        L.orc_grender.argtypes = [C.c_void_p] * 6
        L.orc_perlin_generate.argtypes = [C.c_uint64, C.c_uint32, C.c_void_p]
        L.orc_perlin_turb.restype = C.c_double
        L.orc_perlin_turb.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
        L.orc_sin_portable.restype = C.c_double
        L.orc_sin_portable.argtypes = [C.c_double]
        L.orc_atan2_msun.restype = C.c_double
        L.orc_atan2_msun.argtypes = [C.c_double, C.c_double]
        L.orc_acos_msun.restype = C.c_double
        L.orc_acos_msun.argtypes = [C.c_double]
        L.orc_transform_then.argtypes = [C.c_void_p] * 3
        L.orc_transform_inverse.restype = C.c_int32
        L.orc_transform_inverse.argtypes = [C.c_void_p] * 2
        L.orc_rotation.argtypes = [C.c_double, C.c_int32, C.c_void_p]
        L.orc_gprim_box.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def philox4x32_10(ctr, key):
    ctr = np.asarray(ctr, dtype=np.uint32)
    key = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().orc_philox4x32_10(_p(ctr), _p(key), _p(out))
    return out


def uniform(seed, pixel, sample, vertex, k, rng_mode=W64, kind=0):
    return lib().orc_uniform(seed, pixel, sample, vertex, k, rng_mode, kind)


def uniform_index(seed, pixel, sample, vertex, k, n, rng_mode=W64):
    return lib().orc_uniform_index(seed, pixel, sample, vertex, k, rng_mode, n)


def sincos(phi, math_mode=LIBM):
    s, c = C.c_double(), C.c_double()
    lib().orc_sincos(phi, math_mode, C.byref(s), C.byref(c))
    return s.value, c.value


class SceneDesc:
    """Plain-data scene: what scenes::simple builds (scenes/src/lib.rs:155-233), as arrays."""

    def __init__(self, spheres, sphere_mat, materials, planes, plane_mat, lights, cam_builder=None):
        self.spheres = np.ascontiguousarray(spheres, dtype=np.float64).reshape(-1, 4)
        self.sphere_mat = np.ascontiguousarray(sphere_mat, dtype=np.uint32)
        self.materials = materials  # ctypes array of Material
        self.planes = np.ascontiguousarray(planes, dtype=np.float64).reshape(-1, 6)
        self.plane_mat = np.ascontiguousarray(plane_mat, dtype=np.uint32)
        self.lights = np.ascontiguousarray(lights, dtype=np.float64).reshape(-1, 4)
        self.cam_builder = cam_builder

    def materials_array(self):
        return np.array([[m.kind, m.r, m.g, m.b, m.param] for m in self.materials], dtype=np.float64).reshape(-1, 5)


def make_materials(rows):
    arr = (Material * len(rows))()
    for i, (kind, r, g, b, param) in enumerate(rows):
        arr[i] = Material(int(kind), float(r), float(g), float(b), float(param))
    return arr


def scene_simple(seed=20261018, n=11, p_lambertian=0.8, p_metal=0.95, ground=0) -> SceneDesc:
    L = lib()
    h = L.orc_desc_simple(seed, n, p_lambertian, p_metal, ground)
    try:
        cnt = np.zeros(4, dtype=np.uint64)
        L.orc_desc_counts(h, _p(cnt))
        ns, nm, npl, nl = (int(x) for x in cnt)
        spheres = np.zeros((ns, 4)); smat = np.zeros(ns, dtype=np.uint32)
        mats = (Material * nm)()
        planes = np.zeros((npl, 6)); pmat = np.zeros(npl, dtype=np.uint32)
        lights = np.zeros((nl, 4))
        cb = CameraBuilder()
        L.orc_desc_copy(h, _p(spheres), _p(smat), C.cast(mats, C.c_void_p), _p(planes), _p(pmat), _p(lights),
                        C.cast(C.pointer(cb), C.c_void_p))
    finally:
        L.orc_desc_destroy(h)
    return SceneDesc(spheres, smat, mats, planes, pmat, lights, cb)


def camera_build(cb: CameraBuilder) -> Camera:
    cam = Camera()
    lib().orc_camera_build(C.byref(cb), C.byref(cam))
    return cam


def camera_for(desc: SceneDesc, width, height, spp, max_depth, vfov=40.0) -> Camera:
    """What bin/src/main.rs:72-79 does with the scene's CameraBuilder."""
    cb = CameraBuilder.from_buffer_copy(desc.cam_builder)
    cb.vfov = vfov
    cb.aspect_ratio = width / height
    cb.has_aspect = 1
    cb.width, cb.has_width, cb.height, cb.has_height = width, 1, height, 1
    cb.spp, cb.max_depth = spp, max_depth
    return camera_build(cb)


class Scene:
    def __init__(self, desc: SceneDesc):
        self.desc = desc
        L = lib()
        self.h = L.orc_scene_create(len(desc.sphere_mat), _p(desc.spheres), _p(desc.sphere_mat), len(desc.materials),
                                    C.cast(desc.materials, C.c_void_p), len(desc.plane_mat), _p(desc.planes),
                                    _p(desc.plane_mat), len(desc.lights), _p(desc.lights))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_scene_destroy(self.h)
            self.h = None

    def bvh_stats(self):
        s = BvhStats()
        lib().orc_scene_bvh_stats(self.h, C.byref(s))
        return {n: int(getattr(s, n)) for n, _ in s._fields_}

    def trace_batch(self, o, d, tmin=EPS, tmax=float("inf"), faithful=False):
        o = np.ascontiguousarray(o, dtype=np.float64); d = np.ascontiguousarray(d, dtype=np.float64)
        n = o.shape[0]
        prim = np.zeros(n, dtype=np.int32); t = np.zeros(n)
        cnt = Counters()
        lib().orc_trace_batch(self.h, n, _p(o), _p(d), tmin, tmax, _p(prim), _p(t), int(faithful), C.byref(cnt))
        return prim, t, cnt.as_dict()

    def scatter_batch(self, o, d, pixel, sample, vertex, opts: Options):
        o = np.ascontiguousarray(o, dtype=np.float64); d = np.ascontiguousarray(d, dtype=np.float64)
        n = o.shape[0]
        pixel = np.ascontiguousarray(pixel, dtype=np.uint32); sample = np.ascontiguousarray(sample, dtype=np.uint32)
        vertex = np.ascontiguousarray(vertex, dtype=np.uint32)
        prim = np.zeros(n, dtype=np.int32); t = np.zeros(n); kind = np.zeros(n, dtype=np.uint32)
        p = np.zeros((n, 3)); normal = np.zeros((n, 3)); dr = np.zeros((n, 3)); w = np.zeros((n, 3))
        lib().orc_scatter_batch(self.h, C.byref(opts), n, _p(o), _p(d), _p(pixel), _p(sample), _p(vertex), _p(prim), _p(t),
                                _p(kind), _p(p), _p(normal), _p(dr), _p(w))
        return dict(prim=prim, t=t, kind=kind, p=p, normal=normal, dir=dr, weight=w)

    def shade_batch(self, d, p, normal, front_face, mat_kind, material, pixel, sample, vertex, opts: Options):
        """Material::scatter + mixture pdf on caller-supplied hit records (the counterpart of rtw_shade_batch)."""
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        u = lambda a: np.ascontiguousarray(a, dtype=np.uint32)
        d, p, normal, material = f(d), f(p), f(normal), f(material)
        front_face, mat_kind, pixel, sample, vertex = u(front_face), u(mat_kind), u(pixel), u(sample), u(vertex)
        n = d.shape[0]
        kind = np.zeros(n, dtype=np.uint32); dr = np.zeros((n, 3)); w = np.zeros((n, 3))
        lib().orc_shade_batch(self.h, C.byref(opts), n, _p(d), _p(p), _p(normal), _p(front_face), _p(mat_kind), _p(material),
                              _p(pixel), _p(sample), _p(vertex), _p(kind), _p(dr), _p(w))
        return dict(kind=kind, dir=dr, weight=w)

    def render(self, cam: Camera, opts: Options, row_begin=0, row_end=0xFFFFFFFF):
        img = np.zeros((cam.height, cam.width, 3))
        cnt = Counters(); pan = C.c_uint32(0)
        sec = lib().orc_render(self.h, C.byref(cam), C.byref(opts), _p(img), row_begin, row_end, C.byref(cnt), C.byref(pan))
        return img, sec, cnt.as_dict(), bool(pan.value)

    def path_radiance(self, cam: Camera, opts: Options, i, j, sample):
        i = np.ascontiguousarray(i, dtype=np.uint32); j = np.ascontiguousarray(j, dtype=np.uint32)
        sample = np.ascontiguousarray(sample, dtype=np.uint32)
        out = np.zeros((len(i), 3))
        lib().orc_path_radiance(self.h, C.byref(cam), C.byref(opts), len(i), _p(i), _p(j), _p(sample), _p(out))
        return out


def get_rays(cam: Camera, opts: Options, i, j, sample):
    i = np.ascontiguousarray(i, dtype=np.uint32); j = np.ascontiguousarray(j, dtype=np.uint32)
    sample = np.ascontiguousarray(sample, dtype=np.uint32)
    o = np.zeros((len(i), 3)); d = np.zeros((len(i), 3))
    lib().orc_get_rays(C.byref(cam), C.byref(opts), len(i), _p(i), _p(j), _p(sample), _p(o), _p(d))
    return o, d


def resolve(rgb_sum, spp):
    a = np.ascontiguousarray(rgb_sum, dtype=np.float64)
    out = np.zeros(a.shape, dtype=np.uint8)
    lib().orc_resolve(_p(a), a.size, spp, _p(out))
    return out


def options(seed=20261018, tmin=EPS, rng_mode=W64, math_mode=LIBM, faithful_bvh=False, threads=0, fix_nan=False) -> Options:
    return Options(seed, tmin, rng_mode, math_mode, int(faithful_bvh), threads, int(fix_nan), 0)


def hardware_threads():
    return int(lib().orc_hardware_threads())


# ---- general scenes (rtw_oracle_general.hpp) --------------------------------------------------------------------
DIFFUSE_LIGHT, ISOTROPIC = 4, 5


class GTransform(C.Structure):
    _fields_ = [("rotation", C.c_double * 9), ("translation", C.c_double * 3)]


class GPerlin(C.Structure):
    _fields_ = [("rand_vec", (C.c_double * 3) * 256), ("perm_x", C.c_uint8 * 256), ("perm_y", C.c_uint8 * 256), ("perm_z", C.c_uint8 * 256)]


def transform(rotation=(1, 0, 0, 0, 1, 0, 0, 0, 1), translation=(0, 0, 0)) -> GTransform:
    t = GTransform()
    t.rotation[:] = [float(x) for x in rotation]
    t.translation[:] = [float(x) for x in translation]
    return t


def transform_then(a: GTransform, b: GTransform) -> GTransform:
    out = GTransform()
    lib().orc_transform_then(C.byref(a), C.byref(b), C.byref(out))
    return out


def transform_inverse(a: GTransform):
    out = GTransform()
    return out if lib().orc_transform_inverse(C.byref(a), C.byref(out)) else None


def rotation(angle_degrees, axis) -> GTransform:
    out = GTransform()
    lib().orc_rotation(float(angle_degrees), int(axis), C.byref(out))
    return out


def perlin_generate(seed, index=0) -> GPerlin:
    out = GPerlin()
    lib().orc_perlin_generate(seed, index, C.byref(out))
    return out


def perlin_turb(tables: GPerlin, p, depth=7) -> float:
    """depth <= 0: Perlin::noise; else Perlin::turb."""
    p = np.ascontiguousarray(p, dtype=np.float64)
    return lib().orc_perlin_turb(C.byref(tables), _p(p), depth)


def sin_portable(x: float) -> float:
    return lib().orc_sin_portable(float(x))


def atan2_msun(y: float, x: float) -> float:
    return lib().orc_atan2_msun(float(y), float(x))


def acos_msun(x: float) -> float:
    return lib().orc_acos_msun(float(x))


class GScene:
    """General-scene oracle built from a scene description with the layout of rtw_scene_desc (include/rtw.h); `desc_pod` is
    any ctypes structure of that layout (the caller keeps the arrays it points into alive)."""

    def __init__(self, desc_pod, keepalive=None):
        self._keep = (desc_pod, keepalive)
        self.h = lib().orc_gscene_create(C.cast(C.pointer(desc_pod), C.c_void_p))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_gscene_destroy(self.h)
            self.h = None

    def prim_box(self, index, lights=False):
        out = np.zeros(6)
        lib().orc_gprim_box(self.h, int(lights), index, _p(out))
        return out

    def trace_batch(self, o, d, tmin=EPS, tmax=float("inf")):
        o = np.ascontiguousarray(o, dtype=np.float64); d = np.ascontiguousarray(d, dtype=np.float64)
        n = o.shape[0]
        prim = np.zeros(n, dtype=np.int32); t = np.zeros(n); p = np.zeros((n, 3)); normal = np.zeros((n, 3))
        lib().orc_gtrace_batch(self.h, n, _p(o), _p(d), tmin, tmax, _p(prim), _p(t), _p(p), _p(normal))
        return prim, t, p, normal

    def scatter_batch(self, o, d, pixel, sample, vertex, opts: Options):
        o = np.ascontiguousarray(o, dtype=np.float64); d = np.ascontiguousarray(d, dtype=np.float64)
        n = o.shape[0]
        pixel, sample, vertex = (np.ascontiguousarray(a, dtype=np.uint32) for a in (pixel, sample, vertex))
        prim = np.zeros(n, dtype=np.int32); t = np.zeros(n); kind = np.zeros(n, dtype=np.uint32)
        p, normal, dr, w, e = (np.zeros((n, 3)) for _ in range(5))
        lib().orc_gscatter_batch(self.h, C.byref(opts), n, _p(o), _p(d), _p(pixel), _p(sample), _p(vertex), _p(prim), _p(t),
                                 _p(kind), _p(p), _p(normal), _p(dr), _p(w), _p(e))
        return dict(prim=prim, t=t, kind=kind, p=p, normal=normal, dir=dr, weight=w, emitted=e)

    def path_radiance(self, cam: Camera, opts: Options, i, j, sample):
        i, j, sample = (np.ascontiguousarray(a, dtype=np.uint32) for a in (i, j, sample))
        out = np.zeros((len(i), 3))
        lib().orc_gpath_radiance(self.h, C.byref(cam), C.byref(opts), len(i), _p(i), _p(j), _p(sample), _p(out))
        return out

    def render(self, cam: Camera, opts: Options):
        img = np.zeros((cam.height, cam.width, 3))
        cnt = Counters(); pan = C.c_uint32(0)
        sec = lib().orc_grender(self.h, C.byref(cam), C.byref(opts), _p(img), C.byref(cnt), C.byref(pan))
        return img, sec, cnt.as_dict(), bool(pan.value)
