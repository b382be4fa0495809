"""ctypes binding of librtw_cuda.so — the C ABI declared in include/rtw.h (+ include/rtw_host.h).

The library is the product; this module only marshals buffers.  There is no fallback of any kind:
if the shared library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

RTW_OK, RTW_E_INVALID, RTW_E_CUDA, RTW_E_NO_DEVICE, RTW_E_UNSUPPORTED, RTW_E_NOMEM = 0, -1, -2, -3, -4, -5
RTW_LAMBERTIAN, RTW_METAL, RTW_DIELECTRIC, RTW_INVISIBLE, RTW_DIFFUSE_LIGHT, RTW_ISOTROPIC = 0, 1, 2, 3, 4, 5
RTW_PRIM_SPHERE, RTW_PRIM_PLANE, RTW_PRIM_QUAD, RTW_PRIM_TRIANGLE, RTW_PRIM_CUBOID = 0, 1, 2, 3, 4
RTW_TEX_NOISE, RTW_TEX_CHECKER = 1, 2
RTW_BVH_AUTO, RTW_BVH_HOST_SAH, RTW_BVH_DEVICE_LBVH = 0, 1, 2
RTW_F32, RTW_F64 = 0, 1
RTW_MEGAKERNEL, RTW_WAVEFRONT = 0, 1
RTW_FLAG_FIX_NAN, RTW_FLAG_COUNT_EVENTS, RTW_FLAG_LANE_PER_PIXEL, RTW_FLAG_NO_CANDIDATES = 1, 2, 4, 8
RTW_COLLECTIVE_AUTO, RTW_COLLECTIVE_PEER, RTW_COLLECTIVE_NCCL = 0, 1, 2
RTW_COMM_ID_BYTES = 128
RTW_TILE_W = RTW_TILE_H = 16
EPSILON = 2.220446049250313e-16
TMIN_REFERENCE = -1.0      # rtw_opts.tmin: machine epsilon of the working precision (the reference's f64::EPSILON analogue)

# every symbol include/rtw.h and include/rtw_host.h declare
RTW_SYMBOLS = (
    "rtw_abi_version", "rtw_last_error", "rtw_camera_build", "rtw_philox4x32_10", "rtw_tiles_total", "rtw_tiles_per_rank",
    "rtw_device_count", "rtw_release_cached_memory", "rtw_scene_create", "rtw_scene_destroy", "rtw_scene_info", "rtw_render", "rtw_render_tiles_device",
    "rtw_untile_resolve_device", "rtw_trace_batch", "rtw_scatter_batch", "rtw_get_rays", "rtw_path_radiance",
    "rtw_render_samples_device", "rtw_resolve_accum_device", "rtw_accum_slots", "rtw_render_samples", "rtw_resolve_accum", "rtw_set_bvh_builder", "rtw_scene_bvh_builder", "rtw_scene_create_general", "rtw_transform_then", "rtw_transform_inverse", "rtw_rotation", "rtw_perlin_generate",
    "rtw_scene_export_bvh", "rtw_shade_batch", "rtw_render_multi", "rtw_comm_unique_id", "rtw_comm_init_rank", "rtw_comm_destroy",
    "rtw_comm_rank", "rtw_comm_world", "rtw_render_rank", "rtw_render_rank_device", "rtw_scene_sync",
)
RTWH_SYMBOLS = ("rtwh_scene_simple", "rtwh_scene_desc_destroy", "rtwh_scene_desc_counts", "rtwh_scene_desc_copy")


class RtwError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"rtw error {code}: {msg}")
        self.code = code


class rtw_material(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("texture", C.c_uint32), ("r", C.c_double), ("g", C.c_double), ("b", C.c_double),
                ("param", C.c_double)]


class rtw_quad(C.Structure):
    _fields_ = [("q", C.c_double * 3), ("u", C.c_double * 3), ("v", C.c_double * 3)]


class rtw_cuboid(C.Structure):
    _fields_ = [("p", C.c_double * 3), ("q", C.c_double * 3)]


class rtw_transform(C.Structure):
    _fields_ = [("rotation", C.c_double * 9), ("translation", C.c_double * 3)]


class rtw_prim(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("index", C.c_uint32), ("material", C.c_uint32), ("transform", C.c_int32)]


class rtw_texture(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("perlin", C.c_uint32), ("scale", C.c_double), ("even", C.c_uint32), ("odd", C.c_uint32),
                ("even_colour", C.c_double * 3), ("odd_colour", C.c_double * 3)]


class rtw_perlin(C.Structure):
    _fields_ = [("rand_vec", (C.c_double * 3) * 256), ("perm_x", C.c_uint8 * 256), ("perm_y", C.c_uint8 * 256), ("perm_z", C.c_uint8 * 256)]


class rtw_scene_desc(C.Structure):
    _fields_ = [("spheres", C.c_void_p), ("n_spheres", C.c_uint64), ("planes", C.c_void_p), ("n_planes", C.c_uint64),
                ("quads", C.c_void_p), ("n_quads", C.c_uint64), ("cuboids", C.c_void_p), ("n_cuboids", C.c_uint64),
                ("transforms", C.c_void_p), ("n_transforms", C.c_uint64), ("materials", C.c_void_p), ("n_materials", C.c_uint64),
                ("textures", C.c_void_p), ("n_textures", C.c_uint64), ("perlins", C.c_void_p), ("n_perlins", C.c_uint64),
                ("world", C.c_void_p), ("n_world", C.c_uint64), ("lights", C.c_void_p), ("n_lights", C.c_uint64),
                ("world_is_bvh", C.c_uint32), ("lights_is_bvh", C.c_uint32)]


class rtw_sphere(C.Structure):
    _fields_ = [("cx", C.c_double), ("cy", C.c_double), ("cz", C.c_double), ("r", C.c_double)]


class rtw_plane(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("px", "py", "pz", "nx", "ny", "nz")]


class rtw_camera(C.Structure):
    _fields_ = [(n, C.c_double * 3) for n in ("center", "pixel00_loc", "pixel_delta_u", "pixel_delta_v", "defocus_disk_u",
                                              "defocus_disk_v", "background")] + [
        ("defocus_angle", C.c_double), ("image_width", C.c_uint32), ("image_height", C.c_uint32),
        ("samples_per_pixel", C.c_uint32), ("max_depth", C.c_uint32)]


class rtw_camera_builder(C.Structure):
    _fields_ = [("aspect_ratio", C.c_double), ("has_aspect_ratio", C.c_uint32), ("image_width", C.c_uint32),
                ("has_image_width", C.c_uint32), ("image_height", C.c_uint32), ("has_image_height", C.c_uint32),
                ("samples_per_pixel", C.c_uint32), ("max_depth", C.c_uint32), ("background", C.c_double * 3),
                ("vfov", C.c_double), ("lookfrom", C.c_double * 3), ("lookat", C.c_double * 3), ("vup", C.c_double * 3),
                ("defocus_angle", C.c_double), ("focus_dist", C.c_double)]


class rtw_opts(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("tmin", C.c_double), ("precision", C.c_uint32), ("mode", C.c_uint32),
                ("flags", C.c_uint32), ("reserved", C.c_uint32)]


class rtw_stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("paths", "rays", "node_visits", "sphere_tests", "light_tests", "lambertian", "metal",
                                          "dielectric", "absorbed", "missed", "depth_out")] + [
        ("kernel_ms", C.c_double), ("total_ms", C.c_double), ("launches", C.c_uint32), ("reserved", C.c_uint32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_ if n != "reserved"}


_lib = None


def library_path() -> str:
    return _build.LIB_PATH


def load(build_if_missing: bool = True):
    """Load librtw_cuda.so (building it with nvcc first when absent/stale and allowed)."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    if os.environ.get("RTW_LIBRARY"):                      # tuning builds (scripts/): another build of the same library
        path = os.environ["RTW_LIBRARY"]
    elif build_if_missing and os.environ.get("RTW_NO_BUILD") != "1":
        path = _build.build_library()
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: run `python -m ray_tracing_weekend_b200.build` (needs nvcc). "
                          "There is no CPU fallback.")
    L = C.CDLL(path)
    vp, u32, u64, dbl, sz = C.c_void_p, C.c_uint32, C.c_uint64, C.c_double, C.c_size_t
    L.rtw_abi_version.restype = C.c_int
    L.rtw_last_error.restype = C.c_char_p
    L.rtw_camera_build.argtypes = [vp, vp]
    L.rtw_philox4x32_10.argtypes = [vp, vp, vp]
    L.rtw_philox4x32_10.restype = None
    L.rtw_tiles_total.argtypes = [u32, u32]; L.rtw_tiles_total.restype = u32
    L.rtw_tiles_per_rank.argtypes = [u32, u32, u32]; L.rtw_tiles_per_rank.restype = u32
    L.rtw_device_count.restype = C.c_int
    L.rtw_release_cached_memory.restype = C.c_int
    L.rtw_scene_create.argtypes = [vp, vp, sz, vp, vp, sz, vp, sz, vp, sz, vp]
    L.rtw_scene_destroy.argtypes = [vp]; L.rtw_scene_destroy.restype = None
    L.rtw_scene_info.argtypes = [vp, vp]
    L.rtw_scene_export_bvh.argtypes = [vp, vp, sz, vp, vp, sz, vp]
    L.rtw_render.argtypes = [vp, vp, vp, vp, vp, vp]
    L.rtw_render_tiles_device.argtypes = [vp, vp, vp, u32, u32, vp, vp, vp]
    L.rtw_untile_resolve_device.argtypes = [vp, u32, u32, u32, u32, u32, vp, vp, vp]
    L.rtw_render_samples_device.argtypes = [vp, vp, vp, u32, u32, vp, vp, vp, vp]
    L.rtw_resolve_accum_device.argtypes = [vp, vp, u32, u32, u32, vp, vp, vp]
    L.rtw_accum_slots.argtypes = [u32, u32]; L.rtw_accum_slots.restype = sz
    L.rtw_render_samples.argtypes = [vp, vp, vp, u32, u32, vp, vp, vp]
    L.rtw_resolve_accum.argtypes = [vp, vp, u32, u32, u32, vp, vp]
    L.rtw_trace_batch.argtypes = [vp, vp, vp, sz, dbl, dbl, u32, vp, vp]
    L.rtw_scatter_batch.argtypes = [vp, vp, vp, vp, sz] + [vp] * 10
    L.rtw_shade_batch.argtypes = [vp, vp, sz] + [vp] * 12
    L.rtw_render_multi.argtypes = [vp, vp, vp, C.c_int, vp, u32, vp, vp, vp]
    L.rtw_comm_unique_id.argtypes = [vp]
    L.rtw_comm_init_rank.argtypes = [vp, C.c_int, C.c_int, vp]
    L.rtw_comm_destroy.argtypes = [vp]; L.rtw_comm_destroy.restype = None
    L.rtw_comm_rank.argtypes = [vp]; L.rtw_comm_world.argtypes = [vp]
    L.rtw_render_rank.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    L.rtw_render_rank_device.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp]
    L.rtw_scene_sync.argtypes = [vp, vp]
    L.rtw_get_rays.argtypes = [vp, vp, vp, vp, vp, sz, vp, vp]
    L.rtw_path_radiance.argtypes = [vp, vp, vp, vp, vp, vp, sz, vp]
    L.rtw_scene_create_general.argtypes = [vp, vp]
    L.rtw_set_bvh_builder.argtypes = [C.c_int]
    L.rtw_scene_bvh_builder.argtypes = [vp]
    L.rtw_transform_then.argtypes = [vp, vp, vp]; L.rtw_transform_then.restype = None
    L.rtw_transform_inverse.argtypes = [vp, vp]; L.rtw_transform_inverse.restype = C.c_int
    L.rtw_rotation.argtypes = [dbl, C.c_int, vp]; L.rtw_rotation.restype = None
    L.rtw_perlin_generate.argtypes = [u64, u32, vp]; L.rtw_perlin_generate.restype = None
    L.rtwh_scene_simple.argtypes = [u64, C.c_int32, dbl, dbl, C.c_int32]; L.rtwh_scene_simple.restype = vp
    L.rtwh_scene_desc_destroy.argtypes = [vp]; L.rtwh_scene_desc_destroy.restype = None
    L.rtwh_scene_desc_counts.argtypes = [vp, vp]; L.rtwh_scene_desc_counts.restype = None
    L.rtwh_scene_desc_copy.argtypes = [vp] * 7; L.rtwh_scene_desc_copy.restype = None
    _lib = L
    return L


def check(rc: int):
    if rc != RTW_OK:
        raise RtwError(rc, load().rtw_last_error().decode("utf-8", "replace"))
