"""Host-side mirror of the reference's render interface over the C ABI (include/rtw.h).

Names follow the reference (N9199/ray_tracing_weekend) so tests read like its own:
  CameraBuilder().with_image_width(3)...build() -> Camera         shared/src/camera.rs:44-219
  Camera.render(world, lights) -> rows of sample sums               shared/src/camera.rs:295-297
  Sphere / Plane / Lambertian / Metal / Dialectric / INVISIBLE      shared/src/entities, material.rs
  HittableList.add, BoundedVolumeHierarchy.from_list                shared/src/hittable_collections
  scenes.simple(seed)                                               scenes/src/lib.rs:155-233
All arithmetic happens in librtw_cuda.so; nothing here computes pixels, and nothing falls back.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import (EPSILON, TMIN_REFERENCE, RTW_DIELECTRIC, RTW_DIFFUSE_LIGHT, RTW_F32, RTW_F64, RTW_FLAG_COUNT_EVENTS, RTW_FLAG_FIX_NAN,
                   RTW_INVISIBLE, RTW_ISOTROPIC, RTW_LAMBERTIAN, RTW_MEGAKERNEL, RTW_METAL, RTW_PRIM_CUBOID, RTW_PRIM_PLANE,
                   RTW_PRIM_QUAD, RTW_PRIM_SPHERE, RTW_PRIM_TRIANGLE, RTW_TEX_CHECKER, RTW_TEX_NOISE, RTW_WAVEFRONT, RtwError, rtw_camera,
                   rtw_camera_builder, rtw_cuboid, rtw_material, rtw_opts, rtw_perlin, rtw_plane, rtw_prim, rtw_quad,
                   rtw_scene_desc, rtw_sphere, rtw_stats, rtw_texture, rtw_transform)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


# ---- textures (shared/src/texture.rs) ------------------------------------------------------------------
@dataclass(frozen=True)
class NoiseTexture:                             # NoiseTexture::new(scale), texture.rs:57-102
    """Perlin tables are drawn from Philox stream (seed; 0x9E71A000 + index) — the reference uses the unseeded thread_rng."""
    scale: float
    seed: int = 20261018
    index: int = 0

    def perlin(self) -> rtw_perlin:
        out = rtw_perlin()
        _lib.load().rtw_perlin_generate(self.seed, self.index, C.byref(out))
        return out


@dataclass(frozen=True)
class CheckerTexture:                           # CheckerTexture::new / new_with_colours, texture.rs:24-55
    """even / odd: a colour triple (SolidColour) or a NoiseTexture."""
    even: object
    odd: object
    scale: float

    @staticmethod
    def new_with_colours(even, odd, scale) -> "CheckerTexture":
        return CheckerTexture(tuple(even), tuple(odd), float(scale))


# ---- materials (shared/src/material.rs) --------------------------------------------------------------
@dataclass(frozen=True)
class Material:
    kind: int
    colour: tuple = (0.0, 0.0, 0.0)
    param: float = 0.0
    texture: Optional[object] = None            # None = SolidColour(colour); NoiseTexture | CheckerTexture

    def pod(self, texture_index: int = 0) -> rtw_material:
        return rtw_material(self.kind, texture_index, float(self.colour[0]), float(self.colour[1]), float(self.colour[2]), float(self.param))


def _colour_or_texture(kind, arg):
    if isinstance(arg, (NoiseTexture, CheckerTexture)):
        return Material(kind, (0.0, 0.0, 0.0), 0.0, arg)
    return Material(kind, tuple(arg), 0.0)


def Lambertian(colour_or_texture) -> Material:     # Lambertian::new / new_with_colour, material.rs:331-351
    return _colour_or_texture(RTW_LAMBERTIAN, colour_or_texture)


def Metal(albedo, fuzz) -> Material:           # Metal::new, material.rs:401-405
    return Material(RTW_METAL, tuple(albedo), float(fuzz))


def Dialectric(index_of_refraction) -> Material:   # Dialectric::new, material.rs:443-448  (the reference's spelling)
    return Material(RTW_DIELECTRIC, (1.0, 1.0, 1.0), float(index_of_refraction))


def DiffuseLight(colour_or_texture) -> Material:   # DiffuseLight::new / new_with_colour, material.rs:498-504
    return _colour_or_texture(RTW_DIFFUSE_LIGHT, colour_or_texture)


def Isotropic(colour_or_texture) -> Material:      # Isotropic::new / new_with_colour, material.rs:521-527
    return _colour_or_texture(RTW_ISOTROPIC, colour_or_texture)


INVISIBLE = Material(RTW_INVISIBLE)             # INVISIBLE_PTR, material.rs:319-322


# ---- transformations (geometry/src/transformations.rs, default non-euclid build) ---------------------------
class Axis:
    X, Y, Z = 0, 1, 2


@dataclass(frozen=True)
class Transformation:                           # transformations.rs:96-136
    rotation: tuple = (1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0)     # row-major Matrix3
    translation: tuple = (0.0, 0.0, 0.0)

    def pod(self) -> rtw_transform:
        t = rtw_transform()
        t.rotation[:] = [float(x) for x in self.rotation]
        t.translation[:] = [float(x) for x in self.translation]
        return t

    @staticmethod
    def _from_pod(t: rtw_transform) -> "Transformation":
        return Transformation(tuple(t.rotation), tuple(t.translation))

    def then(self, other: "Transformation") -> "Transformation":
        a, b, out = self.pod(), other.pod(), rtw_transform()
        _lib.load().rtw_transform_then(C.byref(a), C.byref(b), C.byref(out))
        return Transformation._from_pod(out)

    def inverse(self) -> Optional["Transformation"]:
        a, out = self.pod(), rtw_transform()
        return Transformation._from_pod(out) if _lib.load().rtw_transform_inverse(C.byref(a), C.byref(out)) else None


def Translation3(x, y, z) -> Transformation:    # Translation3 = Vec3; From<Vec3> for Transformation, transformations.rs:87-94
    return Transformation(translation=(float(x), float(y), float(z)))


def rotation(angle_degrees, axis) -> Transformation:   # transformations.rs:38-64
    out = rtw_transform()
    _lib.load().rtw_rotation(float(angle_degrees), int(axis), C.byref(out))
    return Transformation._from_pod(out)


# ---- entities -------------------------------------------------------------------------------------------
class _Transformable:                           # Transformable::transform, transformations.rs:193-222
    def transform(self, t: Transformation) -> "Transformed":
        return Transformed(self, Transformation().then(t))


@dataclass(frozen=True)
class Sphere(_Transformable):                   # entities/sphere.rs:25-47
    center: tuple
    radius: float
    material: Material


@dataclass(frozen=True)
class Plane:                                    # entities/plane.rs:21-39
    point: tuple
    normal: tuple
    material: Material


@dataclass(frozen=True)
class Quad(_Transformable):                     # entities/quadrilateral.rs:23-56
    q: tuple
    u: tuple
    v: tuple
    material: Material


@dataclass(frozen=True)
class Triangle(_Transformable):                 # entities/triangles.rs:23-54
    q: tuple
    u: tuple
    v: tuple
    material: Material


@dataclass(frozen=True)
class Cuboid(_Transformable):                   # entities/cuboid.rs:21-50
    p: tuple
    q: tuple
    material: Material


@dataclass(frozen=True)
class Transformed:                              # Transformed<T>, transformations.rs:168-222
    instance: object
    transformation: Transformation

    def transform(self, t: Transformation) -> "Transformed":
        return Transformed(self.instance, self.transformation.then(t))

    @property
    def material(self):
        return self.instance.material


_ENTITY_KIND = {Sphere: RTW_PRIM_SPHERE, Plane: RTW_PRIM_PLANE, Quad: RTW_PRIM_QUAD, Triangle: RTW_PRIM_TRIANGLE, Cuboid: RTW_PRIM_CUBOID}


class HittableList:                             # hittable_collections/hittable_list.rs:247-294
    def __init__(self):
        self.items: List[object] = []

    def add(self, obj):
        inst = obj.instance if isinstance(obj, Transformed) else obj
        if type(inst) not in _ENTITY_KIND:
            raise TypeError(f"{type(obj).__name__} is outside the CUDA backend's scope (Sphere, Plane, Quad, Triangle, Cuboid, Transformed)")
        self.items.append(obj)

    def extend(self, objs):
        for o in objs:
            self.add(o)

    @property
    def spheres(self):
        return [o for o in self.items if isinstance(o, Sphere)]

    @property
    def planes(self):
        return [o for o in self.items if isinstance(o, Plane)]

    def is_simple(self) -> bool:
        """Spheres and planes with SolidColour Lambertian / Metal / Dialectric / Invisible materials only: the fast sphere path."""
        def plane_ok(o):        # Plane::get_aabbox puts an axis-aligned plane's box through the origin (plane.rs:78-107)
            n = np.asarray(o.normal, dtype=np.float64)
            n = n / np.sqrt((n * n).sum())
            return not any(abs(n[(a + 1) % 3]) < EPSILON and abs(n[(a + 2) % 3]) < EPSILON and o.point[a] != 0.0 for a in range(3))
        return all(isinstance(o, (Sphere, Plane)) and o.material.kind <= RTW_INVISIBLE and o.material.texture is None
                   and (isinstance(o, Sphere) or plane_ok(o)) for o in self.items)

    def len(self):
        return len(self.items)

    __len__ = len

    def is_empty(self):
        return self.len() == 0


class BoundedVolumeHierarchy:                   # hittable_collections/bvh.rs:106-143
    """Carries the primitives; the device BVH is built by rtw_scene_create."""

    def __init__(self, hlist: HittableList):
        self.list = hlist

    @classmethod
    def from_list(cls, hlist: HittableList):
        return cls(hlist)

    def len(self):
        return self.list.len()


class SceneDescription:
    """rtw_scene_desc plus the ctypes arrays it points into (kept alive here)."""

    def __init__(self, world, lights):
        w, l = _as_list(world), _as_list(lights)
        spheres, planes, quads, cuboids, transforms, materials, textures, perlins = [], [], [], [], [], [], [], []
        mat_index, tex_index = {}, {}

        def texture_ref(t) -> int:                 # 1-based index into textures[]
            if t not in tex_index:
                tx = rtw_texture()
                if isinstance(t, NoiseTexture):
                    perlins.append(t.perlin())
                    tx.kind, tx.perlin, tx.scale = RTW_TEX_NOISE, len(perlins) - 1, float(t.scale)
                else:
                    tx.kind, tx.scale = RTW_TEX_CHECKER, float(t.scale)
                    for name, sub in (("even", t.even), ("odd", t.odd)):
                        if isinstance(sub, (NoiseTexture, CheckerTexture)):      # sub-textures land in the table before their parent
                            setattr(tx, name, texture_ref(sub))
                        else:
                            getattr(tx, name + "_colour")[:] = [float(x) for x in sub]
                textures.append(tx)
                tex_index[t] = len(textures)
            return tex_index[t]

        def material_id(m: Material) -> int:
            if m not in mat_index:
                t = texture_ref(m.texture) if m.texture is not None else 0
                materials.append(m.pod(t))
                mat_index[m] = len(materials) - 1
            return mat_index[m]

        def entry(obj) -> rtw_prim:
            tr = -1
            if isinstance(obj, Transformed):
                transforms.append(obj.transformation.pod())
                tr = len(transforms) - 1
                obj = obj.instance
            kind = _ENTITY_KIND[type(obj)]
            f3 = lambda v: (C.c_double * 3)(*map(float, v))
            if kind == RTW_PRIM_SPHERE:
                spheres.append(rtw_sphere(*map(float, obj.center), float(obj.radius))); idx = len(spheres) - 1
            elif kind == RTW_PRIM_PLANE:
                planes.append(rtw_plane(*map(float, obj.point), *map(float, obj.normal))); idx = len(planes) - 1
            elif kind == RTW_PRIM_CUBOID:
                cuboids.append(rtw_cuboid(f3(obj.p), f3(obj.q))); idx = len(cuboids) - 1
            else:
                quads.append(rtw_quad(f3(obj.q), f3(obj.u), f3(obj.v))); idx = len(quads) - 1
            return rtw_prim(kind, idx, material_id(obj.material), tr)

        wl = [entry(o) for o in w.items]
        ll = [entry(o) for o in l.items]

        def arr(ctype, items):
            a = (ctype * max(1, len(items)))()
            for i, x in enumerate(items):
                a[i] = x
            return a

        self._keep = dict(spheres=arr(rtw_sphere, spheres), planes=arr(rtw_plane, planes), quads=arr(rtw_quad, quads),
                          cuboids=arr(rtw_cuboid, cuboids), transforms=arr(rtw_transform, transforms),
                          materials=arr(rtw_material, materials), textures=arr(rtw_texture, textures),
                          perlins=arr(rtw_perlin, perlins), world=arr(rtw_prim, wl), lights=arr(rtw_prim, ll))
        d = rtw_scene_desc()
        for name, items in (("spheres", spheres), ("planes", planes), ("quads", quads), ("cuboids", cuboids), ("transforms", transforms),
                            ("materials", materials), ("textures", textures), ("perlins", perlins), ("world", wl), ("lights", ll)):
            setattr(d, name, C.cast(self._keep[name], C.c_void_p))
            setattr(d, "n_" + name, len(items))
        d.world_is_bvh = 1 if isinstance(world, BoundedVolumeHierarchy) else 0
        d.lights_is_bvh = 1 if isinstance(lights, BoundedVolumeHierarchy) else 0
        self.pod = d
        self.upload_bytes = sum(C.sizeof(a) for a in self._keep.values())
        self.n_world, self.n_lights = len(wl), len(ll)


def _as_list(world) -> HittableList:
    if isinstance(world, BoundedVolumeHierarchy):
        return world.list
    if isinstance(world, HittableList):
        return world
    raise TypeError("world / lights must be a HittableList or a BoundedVolumeHierarchy")


@dataclass
class RenderOptions:
    seed: int = 20261018
    tmin: float = TMIN_REFERENCE               # camera.rs:473: EPSILON of the working precision
    precision: int = RTW_F32
    mode: int = RTW_WAVEFRONT                   # the faster FP32 renderer; RTW_MEGAKERNEL gives the same image
    flags: int = 0

    def pod(self) -> rtw_opts:
        return rtw_opts(self.seed, self.tmin, self.precision, self.mode, self.flags, 0)


class Scene:
    """rtw_scene handle: world (planes + spheres) and lights uploaded to the current CUDA device."""

    def __init__(self, world, lights, general: Optional[bool] = None):
        """general=None picks the sphere path (rtw_scene_create) when the scene allows it, else rtw_scene_create_general;
        True forces the general path."""
        w, l = _as_list(world), _as_list(lights)
        simple = (w.is_simple() and all(isinstance(o, Sphere) for o in l.items) and not isinstance(lights, BoundedVolumeHierarchy)
                  and (len(l.items) > 0 or not any(o.material.kind == RTW_LAMBERTIAN for o in w.items)))
        if general is None:
            general = not simple
        self.general = bool(general)
        if self.general:
            self.desc = SceneDescription(world, lights)
            self._h = C.c_void_p()
            _lib.check(_lib.load().rtw_scene_create_general(C.byref(self.desc.pod), C.byref(self._h)))
            self.n_spheres, self.n_planes, self.n_lights = len(w.spheres), len(w.planes), len(l.items)
            self.upload_bytes = self.desc.upload_bytes
            return
        ns, npl, nl = len(w.spheres), len(w.planes), len(l.spheres)
        mats = (rtw_material * max(1, ns + npl))()
        sph = (rtw_sphere * max(1, ns))()
        smat = np.arange(ns, dtype=np.uint32)
        pl = (rtw_plane * max(1, npl))()
        pmat = np.arange(ns, ns + npl, dtype=np.uint32)
        li = (rtw_sphere * max(1, nl))()
        for k, s in enumerate(w.spheres):
            sph[k] = rtw_sphere(*map(float, s.center), float(s.radius))
            mats[k] = s.material.pod()
        for k, p in enumerate(w.planes):
            pl[k] = rtw_plane(*map(float, p.point), *map(float, p.normal))
            mats[ns + k] = p.material.pod()
        for k, s in enumerate(l.spheres):
            li[k] = rtw_sphere(*map(float, s.center), float(s.radius))
        self._h = C.c_void_p()
        L = _lib.load()
        _lib.check(L.rtw_scene_create(C.cast(sph, C.c_void_p), _p(smat), ns, C.cast(pl, C.c_void_p), _p(pmat), npl,
                                      C.cast(mats, C.c_void_p), ns + npl, C.cast(li, C.c_void_p), nl, C.byref(self._h)))
        self.n_spheres, self.n_planes, self.n_lights = ns, npl, nl
        self.upload_bytes = ns * 32 + ns * 4 + npl * 48 + npl * 4 + (ns + npl) * 40 + nl * 32

    @classmethod
    def from_arrays(cls, spheres, sphere_materials, planes=None, plane_materials=None, lights=None):
        """Array form of the constructor for large scenes: spheres [n,4] f64 (cx,cy,cz,r), sphere_materials [n,5] f64
        (kind,r,g,b,param) one row per sphere, planes [m,6], plane_materials [m,5], lights [l,4]."""
        self = cls.__new__(cls)
        self.general = False
        spheres = np.ascontiguousarray(spheres, dtype=np.float64).reshape(-1, 4)
        sm = np.asarray(sphere_materials, dtype=np.float64).reshape(-1, 5)
        planes = np.ascontiguousarray(planes if planes is not None else np.zeros((0, 6)), dtype=np.float64).reshape(-1, 6)
        pm = np.asarray(plane_materials if plane_materials is not None else np.zeros((0, 5)), dtype=np.float64).reshape(-1, 5)
        lights = np.ascontiguousarray(lights if lights is not None else np.zeros((0, 4)), dtype=np.float64).reshape(-1, 4)
        ns, npl, nl = len(spheres), len(planes), len(lights)
        mats = np.zeros(ns + npl, dtype=np.dtype([("kind", "<u4"), ("reserved", "<u4"), ("r", "<f8"), ("g", "<f8"), ("b", "<f8"), ("param", "<f8")]))
        allm = np.concatenate([sm, pm]) if ns + npl else np.zeros((0, 5))
        mats["kind"] = allm[:, 0].astype(np.uint32); mats["r"] = allm[:, 1]; mats["g"] = allm[:, 2]; mats["b"] = allm[:, 3]; mats["param"] = allm[:, 4]
        smat = np.arange(ns, dtype=np.uint32); pmat = np.arange(ns, ns + npl, dtype=np.uint32)
        self._h = C.c_void_p()
        _lib.check(_lib.load().rtw_scene_create(_p(spheres), _p(smat), ns, _p(planes), _p(pmat), npl, _p(mats), ns + npl, _p(lights), nl,
                                                C.byref(self._h)))
        self.n_spheres, self.n_planes, self.n_lights = ns, npl, nl
        self.upload_bytes = ns * 32 + ns * 4 + npl * 48 + npl * 4 + (ns + npl) * 40 + nl * 32
        return self

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            _lib.load().rtw_scene_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:       # interpreter shutdown
            pass

    def info(self):
        out = np.zeros(5, dtype=np.uint64)
        _lib.check(_lib.load().rtw_scene_info(self._h, _p(out)))
        return dict(nodes=int(out[0]), leaves=int(out[1]), depth=int(out[2]), max_leaf=int(out[3]), device_bytes=int(out[4]),
                    builder={1: "host-sah", 2: "device-lbvh"}.get(int(_lib.load().rtw_scene_bvh_builder(self._h)), "?"))

    BVH_NODE_DTYPE = np.dtype([("box_min", np.float64, 3), ("box_max", np.float64, 3), ("parent", np.int32), ("left", np.int32),
                               ("right", np.int32), ("first", np.uint32), ("count", np.uint32), ("depth", np.uint32)], align=True)

    def export_bvh(self):
        """The world BVH as flat host records (rtw_bvh_node: the reference's `BVHNode::{Root, Node, Leaf}` sketch, bvh.rs:224-241)
        and the primitive ids in leaf order."""
        L = _lib.load()
        nn, npr = C.c_size_t(0), C.c_size_t(0)
        _lib.check(L.rtw_scene_export_bvh(self._h, None, 0, C.byref(nn), None, 0, C.byref(npr)))
        nodes = np.zeros(nn.value, dtype=self.BVH_NODE_DTYPE)
        order = np.zeros(npr.value, dtype=np.uint32)
        _lib.check(L.rtw_scene_export_bvh(self._h, _p(nodes) if nn.value else None, nn.value, C.byref(nn), _p(order) if npr.value else None,
                                          npr.value, C.byref(npr)))
        return nodes, order

    # Hittable::hit for a batch (hittable.rs:173)
    def trace_batch(self, o, d, tmin=EPSILON, tmax=float("inf"), precision=RTW_F32):
        o = np.ascontiguousarray(o, dtype=np.float64).reshape(-1, 3)
        d = np.ascontiguousarray(d, dtype=np.float64).reshape(-1, 3)
        n = o.shape[0]
        prim = np.full(n, -1, dtype=np.int32)
        t = np.full(n, np.inf)
        _lib.check(_lib.load().rtw_trace_batch(self._h, _p(o), _p(d), n, tmin, tmax, precision, _p(prim), _p(t)))
        return prim, t

    def scatter_batch(self, o, d, pixel, sample, vertex, opts: RenderOptions):
        o = np.ascontiguousarray(o, dtype=np.float64).reshape(-1, 3)
        d = np.ascontiguousarray(d, dtype=np.float64).reshape(-1, 3)
        n = o.shape[0]
        pixel, sample, vertex = (np.ascontiguousarray(a, dtype=np.uint32) for a in (pixel, sample, vertex))
        prim = np.zeros(n, dtype=np.int32); t = np.zeros(n); kind = np.zeros(n, dtype=np.uint32)
        p, normal, dr, w = (np.zeros((n, 3)) for _ in range(4))
        po = opts.pod()
        _lib.check(_lib.load().rtw_scatter_batch(self._h, C.byref(po), _p(o), _p(d), n, _p(pixel), _p(sample), _p(vertex),
                                                 _p(prim), _p(t), _p(kind), _p(p), _p(normal), _p(dr), _p(w)))
        return dict(prim=prim, t=t, kind=kind, p=p, normal=normal, dir=dr, weight=w)

    def shade_batch(self, d, p, normal, front_face, mat_kind, material, pixel, sample, vertex, opts: RenderOptions):
        """Material::scatter + the mixture-pdf sample on caller-supplied hit records (rtw_shade_batch): d = incoming direction,
        (p, normal, front_face) = the HitRecord, (mat_kind, material = albedo r, g, b, param) = the material that was hit."""
        f = lambda a, k: np.ascontiguousarray(a, dtype=np.float64).reshape(-1, k)
        d, p, normal, material = f(d, 3), f(p, 3), f(normal, 3), f(material, 4)
        front_face, mat_kind, pixel, sample, vertex = (np.ascontiguousarray(a, dtype=np.uint32) for a in (front_face, mat_kind, pixel, sample, vertex))
        n = d.shape[0]
        kind = np.zeros(n, dtype=np.uint32)
        dr, w = np.zeros((n, 3)), np.zeros((n, 3))
        po = opts.pod()
        _lib.check(_lib.load().rtw_shade_batch(self._h, C.byref(po), n, _p(d), _p(p), _p(normal), _p(front_face), _p(mat_kind), _p(material),
                                               _p(pixel), _p(sample), _p(vertex), _p(kind), _p(dr), _p(w)))
        return dict(kind=kind, dir=dr, weight=w)

    def path_radiance(self, camera: "Camera", opts: RenderOptions, i, j, sample):
        i, j, sample = (np.ascontiguousarray(a, dtype=np.uint32) for a in (i, j, sample))
        out = np.zeros((len(i), 3))
        po = opts.pod()
        _lib.check(_lib.load().rtw_path_radiance(self._h, C.byref(camera.pod), C.byref(po), _p(i), _p(j), _p(sample), len(i), _p(out)))
        return out

    def render(self, camera: "Camera", opts: Optional[RenderOptions] = None, want_sum=True, want_rgb8=True, out_rgb8: Optional[np.ndarray] = None):
        """rtw_render: host buffers out.  Returns (rgb_sum [h,w,3] f64 | None, rgb8 [h,w,3] u8 | None, stats dict).
        out_rgb8: a caller-owned [h,w,3] u8 array to receive the image (e.g. a view of pinned memory: the read-back is then a
        true asynchronous DMA instead of a staged pageable copy)."""
        opts = opts or RenderOptions()
        h, w = camera.pod.image_height, camera.pod.image_width
        rgb_sum = np.zeros((h, w, 3)) if want_sum else None
        rgb8 = np.zeros((h, w, 3), dtype=np.uint8) if want_rgb8 else None
        if out_rgb8 is not None:
            assert out_rgb8.dtype == np.uint8 and out_rgb8.shape == (h, w, 3) and out_rgb8.flags.c_contiguous
            rgb8 = out_rgb8
        st = rtw_stats()
        po = opts.pod()
        _lib.check(_lib.load().rtw_render(self._h, C.byref(camera.pod), C.byref(po), _p(rgb_sum), _p(rgb8), C.byref(st)))
        return rgb_sum, rgb8, st.as_dict()

    def render_multi(self, camera: "Camera", opts: Optional[RenderOptions] = None, n_gpus: int = 1, devices=None,
                     collective: int = _lib.RTW_COLLECTIVE_AUTO, want_sum=True, want_rgb8=True):
        """Camera::render on n_gpus devices of THIS process (rtw_render_multi): the scene is replicated, every GPU renders its share
        on its own stream, the frame's one collective (fused peer-memory reduce + resolve, or NCCL) runs inside the library."""
        opts = opts or RenderOptions()
        w, h = camera.image_width, camera.image_height
        rgb_sum = np.zeros((h, w, 3)) if want_sum else None
        rgb8 = np.zeros((h, w, 3), dtype=np.uint8) if want_rgb8 else None
        st = rtw_stats()
        po = opts.pod()
        devs = None if devices is None else np.ascontiguousarray(devices, dtype=np.int32)
        _lib.check(_lib.load().rtw_render_multi(self._h, C.byref(camera.pod), C.byref(po), int(n_gpus), _p(devs) if devs is not None else None,
                                                int(collective), _p(rgb_sum) if want_sum else None, _p(rgb8) if want_rgb8 else None, C.byref(st)))
        return rgb_sum, rgb8, st.as_dict()

    def render_rank(self, camera: "Camera", opts: RenderOptions, comm: "Comm", want_sum=False, want_rgb8=True):
        """This rank's part of one frame rendered by all ranks of `comm` (rtw_render_rank): the image arrives on rank 0."""
        w, h = camera.image_width, camera.image_height
        root = comm.rank == 0
        rgb_sum = np.zeros((h, w, 3)) if (want_sum and root) else None
        rgb8 = np.zeros((h, w, 3), dtype=np.uint8) if (want_rgb8 and root) else None
        st = rtw_stats()
        po = opts.pod()
        _lib.check(_lib.load().rtw_render_rank(self._h, C.byref(camera.pod), C.byref(po), comm._h, _p(rgb_sum) if rgb_sum is not None else None,
                                               _p(rgb8) if rgb8 is not None else None, C.byref(st)))
        return rgb_sum, rgb8, st.as_dict()

    def render_rank_device(self, camera: "Camera", opts: RenderOptions, comm: "Comm", d_rgb_sum_ptr: int = 0, d_rgb8_ptr: int = 0, stream: int = 0,
                           want_stats: bool = False):
        """rtw_render_rank_device: device outputs on rank 0, everything enqueued on `stream`; asynchronous unless want_stats."""
        st = rtw_stats()
        po = opts.pod()
        _lib.check(_lib.load().rtw_render_rank_device(self._h, C.byref(camera.pod), C.byref(po), comm._h, C.c_void_p(d_rgb_sum_ptr or None),
                                                      C.c_void_p(d_rgb8_ptr or None), C.c_void_p(stream or None), C.byref(st) if want_stats else None))
        return st.as_dict() if want_stats else None

    def sync(self):
        """rtw_scene_sync: wait for the scene's device, raise if a path did what makes the reference panic; returns the kernel time
        (ms) of the last render call."""
        ms = C.c_double(0.)
        _lib.check(_lib.load().rtw_scene_sync(self._h, C.byref(ms)))
        return ms.value

    def render_samples(self, camera: "Camera", opts: RenderOptions, sample_begin: int, sample_count: int, accum: np.ndarray, poison: np.ndarray):
        """rtw_render_samples: ADD samples [sample_begin, sample_begin + sample_count) of every pixel into the host accumulators
        (see new_accumulators); progressive rendering / checkpointing."""
        assert accum.dtype == np.uint64 and poison.dtype == np.uint32 and accum.flags.c_contiguous and poison.flags.c_contiguous
        st = rtw_stats()
        po = opts.pod()
        _lib.check(_lib.load().rtw_render_samples(self._h, C.byref(camera.pod), C.byref(po), sample_begin, sample_count, _p(accum), _p(poison),
                                                  C.byref(st)))
        return st.as_dict()

    def render_samples_device(self, camera: "Camera", opts: RenderOptions, sample_begin: int, sample_count: int, d_accum_ptr: int,
                              d_poison_ptr: int, stream: int = 0, want_stats=True):
        """rtw_render_samples_device: samples [sample_begin, sample_begin + sample_count) of every pixel into fixed-point accumulators."""
        st = rtw_stats()
        po = opts.pod()
        _lib.check(_lib.load().rtw_render_samples_device(self._h, C.byref(camera.pod), C.byref(po), sample_begin, sample_count,
                                                         C.c_void_p(d_accum_ptr), C.c_void_p(d_poison_ptr), C.c_void_p(stream),
                                                         C.byref(st) if want_stats else None))
        return st.as_dict() if want_stats else None

    def render_tiles_device(self, camera: "Camera", opts: RenderOptions, rank: int, world: int, d_tiles_ptr: int, stream: int = 0,
                            want_stats=True):
        st = rtw_stats()
        po = opts.pod()
        _lib.check(_lib.load().rtw_render_tiles_device(self._h, C.byref(camera.pod), C.byref(po), rank, world, C.c_void_p(d_tiles_ptr),
                                                       C.c_void_p(stream), C.byref(st) if want_stats else None))
        return st.as_dict() if want_stats else None


class Comm:
    """rtw_comm: the communicator of the one-process-per-GPU mode.  Rank 0 calls Comm.unique_id(), the 128 bytes travel out of band
    (torch.distributed, MPI, a file), every rank constructs Comm(id, rank, world) — a collective call."""

    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_uint8 * _lib.RTW_COMM_ID_BYTES)()
        _lib.check(_lib.load().rtw_comm_unique_id(buf))
        return bytes(buf)

    def __init__(self, unique_id: bytes, rank: int, world: int):
        assert len(unique_id) == _lib.RTW_COMM_ID_BYTES
        buf = (C.c_uint8 * _lib.RTW_COMM_ID_BYTES).from_buffer_copy(unique_id)
        self._h = C.c_void_p()
        _lib.check(_lib.load().rtw_comm_init_rank(buf, int(rank), int(world), C.byref(self._h)))
        self.rank, self.world = int(rank), int(world)

    def close(self):
        if self._h:
            _lib.load().rtw_comm_destroy(self._h)
            self._h = C.c_void_p()


def new_accumulators(width: int, height: int):
    """Zeroed host accumulators for Scene.render_samples: (accum [slots, 3] u64, poison [slots] u32)."""
    n = int(_lib.load().rtw_accum_slots(width, height))
    return np.zeros((n, 3), dtype=np.uint64), np.zeros(n, dtype=np.uint32)


def resolve_accum(accum: np.ndarray, poison: np.ndarray, width: int, height: int, spp: int, want_sum=True, want_rgb8=True):
    """rtw_resolve_accum: host accumulators -> (rgb_sum [h, w, 3] f64 | None, rgb8 [h, w, 3] u8 | None)."""
    rgb_sum = np.zeros((height, width, 3)) if want_sum else None
    rgb8 = np.zeros((height, width, 3), dtype=np.uint8) if want_rgb8 else None
    _lib.check(_lib.load().rtw_resolve_accum(_p(accum), _p(poison), width, height, spp, _p(rgb_sum), _p(rgb8)))
    return rgb_sum, rgb8


def resolve_accum_device(d_accum_ptr, d_poison_ptr, width, height, spp, d_rgb_sum_ptr=0, d_rgb8_ptr=0, stream=0):
    _lib.check(_lib.load().rtw_resolve_accum_device(C.c_void_p(d_accum_ptr), C.c_void_p(d_poison_ptr), width, height, spp,
                                                    C.c_void_p(d_rgb_sum_ptr) if d_rgb_sum_ptr else None,
                                                    C.c_void_p(d_rgb8_ptr) if d_rgb8_ptr else None, C.c_void_p(stream)))


def untile_resolve_device(d_tiles_all_ptr, precision, width, height, world, spp, d_rgb_sum_ptr=0, d_rgb8_ptr=0, stream=0):
    _lib.check(_lib.load().rtw_untile_resolve_device(C.c_void_p(d_tiles_all_ptr), precision, width, height, world, spp,
                                                     C.c_void_p(d_rgb_sum_ptr) if d_rgb_sum_ptr else None,
                                                     C.c_void_p(d_rgb8_ptr) if d_rgb8_ptr else None, C.c_void_p(stream)))


def tiles_per_rank(width, height, world):
    return int(_lib.load().rtw_tiles_per_rank(width, height, world))


def tiles_total(width, height):
    return int(_lib.load().rtw_tiles_total(width, height))


# ---- camera (shared/src/camera.rs) -------------------------------------------------------------------------
class Camera:
    def __init__(self, pod: rtw_camera):
        self.pod = pod

    @property
    def image_width(self):
        return self.pod.image_width

    @property
    def image_height(self):
        return self.pod.image_height

    def get_rays(self, i, j, sample, opts: Optional[RenderOptions] = None):
        """Camera::get_ray for a batch (camera.rs:274-293)."""
        opts = opts or RenderOptions()
        i, j, sample = (np.ascontiguousarray(a, dtype=np.uint32) for a in (i, j, sample))
        o = np.zeros((len(i), 3)); d = np.zeros((len(i), 3))
        po = opts.pod()
        _lib.check(_lib.load().rtw_get_rays(C.byref(self.pod), C.byref(po), _p(i), _p(j), _p(sample), len(i), _p(o), _p(d)))
        return o, d

    def render(self, world, lights, opts: Optional[RenderOptions] = None):
        """Camera::render (camera.rs:295-297): [height][width][3] f64 sample sums, row 0 = bottom row."""
        scene = Scene(world, lights)
        try:
            rgb_sum, _, _ = scene.render(self, opts, want_sum=True, want_rgb8=False)
        finally:
            scene.close()
        return rgb_sum


class CameraBuilder:                            # camera.rs:28-219
    def __init__(self, pod: Optional[rtw_camera_builder] = None):
        if pod is None:
            pod = rtw_camera_builder()
            pod.samples_per_pixel, pod.max_depth, pod.vfov, pod.focus_dist = 10, 10, 90.0, 10.0
            pod.lookat[2] = -1.0
            pod.vup[1] = 1.0
        self.pod = pod

    def _with(self, **kw):
        pod = rtw_camera_builder.from_buffer_copy(self.pod)
        for k, v in kw.items():
            if isinstance(v, (tuple, list, np.ndarray)):
                getattr(pod, k)[:] = [float(x) for x in v]
            else:
                setattr(pod, k, v)
        return CameraBuilder(pod)

    def with_aspect_ratio(self, a): return self._with(aspect_ratio=float(a), has_aspect_ratio=1)
    def with_image_width(self, w): return self._with(image_width=int(w), has_image_width=1)
    def with_image_height(self, h): return self._with(image_height=int(h), has_image_height=1)
    def with_samples_per_pixel(self, s): return self._with(samples_per_pixel=int(s))
    def with_max_depth(self, d): return self._with(max_depth=int(d))
    def with_background(self, c): return self._with(background=c)
    def with_vfov(self, v): return self._with(vfov=float(v))
    def with_lookfrom(self, p): return self._with(lookfrom=p)
    def with_lookat(self, p): return self._with(lookat=p)
    def with_vup(self, p): return self._with(vup=p)
    def with_defocus_angle(self, a): return self._with(defocus_angle=float(a))
    def with_focus_dist(self, f): return self._with(focus_dist=float(f))

    def build(self) -> Camera:
        cam = rtw_camera()
        _lib.check(_lib.load().rtw_camera_build(C.byref(self.pod), C.byref(cam)))
        return Camera(cam)


def philox4x32_10(ctr: Sequence[int], key: Sequence[int]) -> np.ndarray:
    ctr = np.asarray(ctr, dtype=np.uint32); key = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    _lib.load().rtw_philox4x32_10(_p(ctr), _p(key), _p(out))
    return out


def set_bvh_builder(mode: int):
    """rtw_set_bvh_builder: RTW_BVH_AUTO | RTW_BVH_HOST_SAH | RTW_BVH_DEVICE_LBVH for the scenes created afterwards."""
    _lib.check(_lib.load().rtw_set_bvh_builder(int(mode)))


def device_count() -> int:
    n = _lib.load().rtw_device_count()
    if n < 0:
        _lib.check(n)
    return n


def write_ppm(path, rgb8):
    """bin/src/main.rs:89-104: ASCII P3, rows reversed (row 0 of the render is the bottom row)."""
    h, w, _ = rgb8.shape
    with open(path, "w") as f:
        f.write(f"P3\n{w} {h}\n255\n")
        for j in range(h - 1, -1, -1):
            f.write("\n".join(f"{r} {g} {b}" for r, g, b in rgb8[j]) + "\n")
