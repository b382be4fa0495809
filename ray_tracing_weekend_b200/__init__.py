"""ray_tracing_weekend_b200 — B200 (sm_100a) CUDA backend for the render path of
N9199/ray_tracing_weekend: Camera::render -> BVH traversal + Sphere::hit -> Material::scatter ->
spp accumulation -> gamma, behind the C ABI of include/rtw.h.

The CUDA library is the product; this package is the host-side mirror of the reference's interface
used by tests and bench.py.  Importing works without a GPU (nvcc builds the library), computing
does not: there is no CPU fallback."""
from . import scenes  # noqa: F401
from ._lib import (EPSILON, TMIN_REFERENCE, RTW_COLLECTIVE_AUTO, RTW_COLLECTIVE_NCCL, RTW_COLLECTIVE_PEER, RTW_BVH_AUTO, RTW_BVH_DEVICE_LBVH, RTW_BVH_HOST_SAH, RTW_F32, RTW_F64, RTW_FLAG_COUNT_EVENTS, RTW_FLAG_FIX_NAN, RTW_FLAG_LANE_PER_PIXEL, RTW_FLAG_NO_CANDIDATES, RTW_MEGAKERNEL, RTW_WAVEFRONT,  # noqa: F401
                   RtwError, library_path, load)
from .api import (INVISIBLE, Axis, BoundedVolumeHierarchy, Camera, CameraBuilder, CheckerTexture, Comm, Cuboid, Dialectric, DiffuseLight,  # noqa: F401
                  HittableList, Isotropic, Lambertian, Material, Metal, NoiseTexture, Plane, Quad, RenderOptions, Scene,
                  SceneDescription, Sphere, Transformation, Transformed, Translation3, Triangle, device_count, new_accumulators, philox4x32_10,
                  resolve_accum, resolve_accum_device, rotation, set_bvh_builder, tiles_per_rank, tiles_total, untile_resolve_device, write_ppm)
