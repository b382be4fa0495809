/* rtw.h — C ABI of the B200 (sm_100a) CUDA path-tracing backend for N9199/ray_tracing_weekend.
 *
 * This is the drop-in boundary for ONE path of the reference: Camera::render(world, lights)
 * (shared/src/camera.rs:295-297 -> render_internal :315-388 -> ray_colour_tail_call :460-522) and
 * the per-ray operations it is made of.  The reference has no FFI of its own; these are the entry
 * points a `cuda` crate of the Rust workspace binds (see INTEGRATION.md for the extern "C" block).
 *
 * Conventions: all structs are plain little-endian POD; the caller owns every buffer it passes and
 * the library keeps no pointer after a call returns; handles are opaque and owned by the library;
 * calls are blocking; a handle must not be used from two threads at once (distinct handles may).
 * Every function returns RTW_OK (0) or a negative RTW_E_* code and never aborts/unwinds;
 * rtw_last_error() gives the thread-local message of the last failure.
 * There is NO CPU fallback: without a CUDA device every compute call returns RTW_E_NO_DEVICE.
 */
#ifndef RTW_H
#define RTW_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTW_ABI_VERSION 3

#if defined(__GNUC__)
#define RTW_API __attribute__((visibility("default")))
#else
#define RTW_API
#endif

enum {
    RTW_OK = 0,
    RTW_E_INVALID = -1,      /* bad argument / inconsistent scene (the reference would panic) */
    RTW_E_CUDA = -2,         /* CUDA runtime error, text in rtw_last_error()                  */
    RTW_E_NO_DEVICE = -3,    /* no CUDA device visible                                        */
    RTW_E_UNSUPPORTED = -4,  /* valid in the reference but outside this backend's scope       */
    RTW_E_NOMEM = -5
};

/* DynMaterial flattened (shared/src/material.rs:251-325).  Textures: SolidColour only
 * (shared/src/texture.rs:15-22). */
enum {
    RTW_LAMBERTIAN = 0, /* Lambertian{texture}                   material.rs:327-376 */
    RTW_METAL = 1,      /* Metal{albedo=(r,g,b), fuzz=param}     material.rs:378-421 */
    RTW_DIELECTRIC = 2, /* Dialectric{index_of_refraction=param} material.rs:423-488 */
    RTW_INVISIBLE = 3,  /* Invisible: never scatters, emits 0    material.rs:319-325 */
    RTW_DIFFUSE_LIGHT = 4, /* DiffuseLight{texture}: emits, never scatters   material.rs:490-514   [general scenes only] */
    RTW_ISOTROPIC = 5      /* Isotropic{texture}: SpherePdf scatter          material.rs:516-554   [general scenes only] */
};
/* texture: 0 = SolidColour(r,g,b) (texture.rs:15-22); k > 0 = rtw_scene_desc.textures[k-1] (general scenes only). */
typedef struct { uint32_t kind; uint32_t texture; double r, g, b, param; } rtw_material;

/* Sphere{center, radius} (shared/src/entities/sphere.rs:25-30) */
typedef struct { double cx, cy, cz, r; } rtw_sphere;
/* Plane{point, normal}; normal is normalised by the library like Plane::new (entities/plane.rs:27-39).
 * One-sided: only rays with dir.normal > EPSILON hit (plane.rs:62-63). */
typedef struct { double px, py, pz, nx, ny, nz; } rtw_plane;

/* ---- general scenes: the other entities, Transformed<T>, emissive materials, NoiseTexture ---- */
/* Quad::new(q, u, v) (entities/quadrilateral.rs:36-56) and Triangle::new(q, u, v) (entities/triangles.rs:34-54). */
typedef struct { double q[3], u[3], v[3]; } rtw_quad;
/* Cuboid::new(p, q): six Quads (entities/cuboid.rs:26-50). */
typedef struct { double p[3], q[3]; } rtw_cuboid;
/* Transformation{rotation (row-major Matrix3), translation} of the default (non-euclid) build
 * (geometry/src/transformations.rs:96-100).  A Transformed<T> hits with the reference's arithmetic, including
 * transform_vector3d adding the translation to the ray DIRECTION (:123-126) and the normal staying in instance space
 * (entities/transformations.rs:14-29). */
typedef struct { double rotation[9]; double translation[3]; } rtw_transform;
enum { RTW_PRIM_SPHERE = 0, RTW_PRIM_PLANE = 1, RTW_PRIM_QUAD = 2, RTW_PRIM_TRIANGLE = 3, RTW_PRIM_CUBOID = 4 };
/* One entry of a HittableList: entity `index` of the array its kind names (quads[] for QUAD and TRIANGLE), material,
 * and transform = -1 or an index into transforms[] (the entry is then a Transformed<T>). */
typedef struct { uint32_t kind, index, material; int32_t transform; } rtw_prim;
enum { RTW_TEX_NOISE = 1,   /* NoiseTexture{noise: perlins[perlin], scale} (texture.rs:57-102) */
       RTW_TEX_CHECKER = 2 };/* CheckerTexture{inv_scale = 1 / scale, even, odd} (texture.rs:24-55): even / odd are 0 = SolidColour(even_colour /
                               odd_colour) or k > 0 = textures[k-1], a NoiseTexture or another CheckerTexture that comes EARLIER in the table.  Reads the hit's (u, v): Quad / Triangle /
                               Cuboid-face coordinates, Sphere::get_sphere_uv (sphere.rs:49-54), Plane::get_plane_uv (plane.rs:41-55: (x, z) for a +y
                               normal, else the fractional x / z of the point rotated onto +y; a checkered plane whose normal is exactly -y is
                               RTW_E_INVALID: the reference's rotation axis is 0 / 0 there and Plane::hit panics on the NaN). */
typedef struct { uint32_t kind, perlin; double scale; uint32_t even, odd; double even_colour[3], odd_colour[3]; } rtw_texture;
/* Perlin's tables (perlin.rs:13-19): rand_vec = 256 UnitSphere samples, three permutations of 0..255. */
typedef struct { double rand_vec[256][3]; uint8_t perm_x[256], perm_y[256], perm_z[256]; } rtw_perlin;
/* world / lights of any scene of scenes/src/lib.rs: two HittableLists over shared entity arrays.
 * world_is_bvh / lights_is_bvh record whether the reference wraps the list in BoundedVolumeHierarchy::from: for `world`
 * it does not change the result (argmin-t either way); for `lights` it changes the f64 rounding of pdf_value
 * (bvh.rs:67-76, 191-194) and is supported for at most 5 lights (one Leaf) — beyond that the reference's
 * aux_random (bvh.rs:78-93) indexes out of range. */
typedef struct {
    const rtw_sphere* spheres;       uint64_t n_spheres;
    const rtw_plane* planes;         uint64_t n_planes;
    const rtw_quad* quads;           uint64_t n_quads;
    const rtw_cuboid* cuboids;       uint64_t n_cuboids;
    const rtw_transform* transforms; uint64_t n_transforms;
    const rtw_material* materials;   uint64_t n_materials;
    const rtw_texture* textures;     uint64_t n_textures;
    const rtw_perlin* perlins;       uint64_t n_perlins;
    const rtw_prim* world;           uint64_t n_world;
    const rtw_prim* lights;          uint64_t n_lights;
    uint32_t world_is_bvh, lights_is_bvh;
} rtw_scene_desc;

/* The fields of Camera the render loop reads (shared/src/camera.rs:231-260). */
typedef struct {
    double center[3], pixel00_loc[3], pixel_delta_u[3], pixel_delta_v[3];
    double defocus_disk_u[3], defocus_disk_v[3], background[3];
    double defocus_angle;
    uint32_t image_width, image_height, samples_per_pixel, max_depth;
} rtw_camera;

/* CameraBuilder (shared/src/camera.rs:28-112); has_* mirror the Option<> fields. */
typedef struct {
    double aspect_ratio; uint32_t has_aspect_ratio;
    uint32_t image_width, has_image_width, image_height, has_image_height;
    uint32_t samples_per_pixel, max_depth;
    double background[3], vfov, lookfrom[3], lookat[3], vup[3], defocus_angle, focus_dist;
} rtw_camera_builder;

enum { RTW_F32 = 0,   /* fast path: FP32 arithmetic, 24-bit uniforms (stream layout W32)            */
       RTW_F64 = 1 }; /* reference-exact path: the reference's f64 operation order, no FMA
                         contraction, 53-bit uniforms (stream layout W64)                           */
enum { RTW_MEGAKERNEL = 0, RTW_WAVEFRONT = 1 };
#define RTW_TMIN_REFERENCE (-1.0)
enum { RTW_FLAG_FIX_NAN = 1u,       /* zero NaN components of a sample before accumulation (the dead
                                       ray_colour twin did, camera.rs:434); OFF = reference behaviour */
       RTW_FLAG_COUNT_EVENTS = 2u, /* fill the event counters of rtw_stats (slower)                 */
       RTW_FLAG_LANE_PER_PIXEL = 4u,/* RTW_F32 only, diagnostic: one lane per pixel instead of the pooled
                                       path stream (the baseline the pooled megakernel is measured against) */
       RTW_FLAG_NO_CANDIDATES = 8u };/* RTW_F32 wavefront, diagnostic: camera rays walk the BVH like every other ray instead of
                                       reading their pixel's candidate list (same image; the event counters then count the
                                       tree walk of every ray) */

typedef struct {
    uint64_t seed;        /* Philox4x32-10 key                                                      */
    double   tmin;        /* lower end of the world.hit range (camera.rs:473).  RTW_TMIN_REFERENCE (any negative
                             value) = the reference's choice, machine epsilon of the working precision: 2^-52 on
                             the RTW_F64 path (== f64::EPSILON), 2^-23 on the RTW_F32 path.  The reference's image
                             depends on how this compares with the rounding noise of hit points (self-intersection),
                             which is a relation in ulps, not in absolute terms. */
    uint32_t precision;   /* RTW_F32 | RTW_F64                                                      */
    uint32_t mode;        /* RTW_MEGAKERNEL | RTW_WAVEFRONT: which FP32 renderer (same image bit for bit; the
                             wavefront is the faster one).  Ignored by RTW_F64, which has a single renderer.    */
    uint32_t flags;
    uint32_t reserved;
} rtw_opts;

typedef struct {
    uint64_t paths;        /* ray_colour_call invocations (camera.rs:326)                           */
    uint64_t rays;         /* world.hit invocations (camera.rs:473): primary + every bounce         */
    uint64_t node_visits;  /* inner BVH nodes visited (two child-box tests each)   [COUNT_EVENTS]   */
    uint64_t sphere_tests; /* ray-sphere tests during world.hit                    [COUNT_EVENTS]   */
    uint64_t light_tests;  /* Sphere::hit calls made by lights.pdf_value           [COUNT_EVENTS]   */
    uint64_t lambertian, metal, dielectric; /* scatter calls by material           [COUNT_EVENTS]   */
    uint64_t absorbed, missed, depth_out;   /* path terminations by reason         [COUNT_EVENTS]   */
    double   kernel_ms;    /* CUDA-event time of the render kernels of this call                    */
    double   total_ms;     /* CUDA-event time of the whole call's device work incl. copies          */
    uint32_t launches;     /* kernels launched by this call                                         */
    uint32_t reserved;
} rtw_stats;

typedef struct rtw_scene rtw_scene;

/* ---- host-side helpers (no GPU needed) -------------------------------------------------------- */
RTW_API int         rtw_abi_version(void);
RTW_API const char* rtw_last_error(void);
/* CameraBuilder::build (shared/src/camera.rs:114-218), bit-identical f64 arithmetic. */
RTW_API int         rtw_camera_build(const rtw_camera_builder* builder, rtw_camera* out);
/* Philox4x32-10 block function; the device code runs the same rounds (KATs in tests/). */
RTW_API void        rtw_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* Image partition used by the tile entry points: tiles of RTW_TILE_W x RTW_TILE_H pixels; tile (tx, ty) has slot
 * k = ty * tiles_x + (tx + rot(ty)) % tiles_x with rot(ty) = ((ty * 0x9E3779B1u) >> 15) % tiles_x (rows rotated so no rank
 * owns whole tile columns); slot k belongs to rank k % world and is that rank's local tile k / world. */
#define RTW_TILE_W 16
#define RTW_TILE_H 16
RTW_API uint32_t    rtw_tiles_total(uint32_t width, uint32_t height);
RTW_API uint32_t    rtw_tiles_per_rank(uint32_t width, uint32_t height, uint32_t world); /* padded: same on every rank */

/* Transformation::then (geometry/src/transformations.rs:104-116): out = a.then(b), i.e. a applied first.  */
RTW_API void        rtw_transform_then(const rtw_transform* a, const rtw_transform* b, rtw_transform* out);
/* Transformation::inverse (:128-136 + matrix3.rs:12-30); returns 0 when the determinant is not a normal number. */
RTW_API int         rtw_transform_inverse(const rtw_transform* a, rtw_transform* out);
/* rotation(angle_degrees, axis 0|1|2) (:38-64). */
RTW_API void        rtw_rotation(double angle_degrees, int axis, rtw_transform* out);
/* Perlin::new (perlin.rs:46-57) with the tables drawn from Philox stream (seed; 0x9E71A000 + index, 0, 0) instead of
 * thread_rng. */
RTW_API void        rtw_perlin_generate(uint64_t seed, uint32_t index, rtw_perlin* out);

/* ---- device management ------------------------------------------------------------------------ */
RTW_API int rtw_device_count(void);           /* >= 0, or RTW_E_* */
/* Device buffers of destroyed scenes / finished renders are kept in a process-wide cache (<= 4 GiB) and reused;
 * this frees them. */
RTW_API int rtw_release_cached_memory(void);

/* ---- scene ------------------------------------------------------------------------------------ */
/* Replaces what scenes::simple hands to render: `world` = BoundedVolumeHierarchy::from(HittableList)
 * of planes + spheres (scenes/src/lib.rs:228), `lights` = HittableList of spheres (:229).
 * Copies the inputs, builds and flattens a BVH on the host and uploads it to the current device.
 * Primitive ids reported by the batch calls: planes first (0..n_planes-1), then spheres, each in
 * input order.  A scene with Lambertian materials and no lights is RTW_E_INVALID (the reference
 * panics on the first light sample, hittable_list.rs:414-419). */
RTW_API int  rtw_scene_create(const rtw_sphere* spheres, const uint32_t* sphere_material, size_t n_spheres,
                      const rtw_plane* planes, const uint32_t* plane_material, size_t n_planes,
                      const rtw_material* materials, size_t n_materials,
                      const rtw_sphere* lights, size_t n_lights,
                      rtw_scene** out);
/* General scenes: any world / lights pair the reference's scenes build from Sphere, Plane, Quad, Triangle, Cuboid,
 * Transformed<T>, with DiffuseLight / Isotropic materials and NoiseTexture (cornell_box, simple_light, debug,
 * simple_transform, perlin_spheres: scenes/src/lib.rs:40-89, 235-653).  Rendered by the general kernels (both
 * precisions; `mode` is ignored).  Primitive ids reported by the batch calls are indices into desc->world.
 * Transformed planes are RTW_E_UNSUPPORTED (their reference AABB is non-finite).
 * An EMPTY lights list next to scattering materials is accepted, as in the reference (scenes::plane, integration-tests plane_test):
 * the reference only panics when a path actually draws a light sample (hittable_list.rs:414-419); a render / batch call in which
 * that happens returns RTW_E_INVALID. */
RTW_API int  rtw_scene_create_general(const rtw_scene_desc* desc, rtw_scene** out);
/* Which builder makes the world BVH of the scenes created AFTER this call (process-wide).  The result of Hittable::hit does not
 * depend on the tree, so images are bit-identical either way; the builders trade build time against traversal cost.
 *   RTW_BVH_HOST_SAH:    binned SAH on one host thread (best trees; 1.0 s for 1 M spheres)
 *   RTW_BVH_DEVICE_LBVH: Morton codes + radix sort + Karras' radix tree + bottom-up fit, all on the GPU (rtw_scene_create only)
 *   RTW_BVH_AUTO:        device LBVH from 200 000 spheres, host SAH below (default)
 * A device-built tree deeper than the traversal stack allows (30 levels) silently falls back to the host builder. */
enum { RTW_BVH_AUTO = 0, RTW_BVH_HOST_SAH = 1, RTW_BVH_DEVICE_LBVH = 2 };
RTW_API int  rtw_set_bvh_builder(int mode);
RTW_API int  rtw_scene_bvh_builder(const rtw_scene* scene);   /* RTW_BVH_HOST_SAH or RTW_BVH_DEVICE_LBVH: the builder that made this scene's tree */
RTW_API void rtw_scene_destroy(rtw_scene* scene);
/* nodes, leaves, depth, max leaf size, bytes resident on the device */
RTW_API int  rtw_scene_info(const rtw_scene* scene, uint64_t out[5]);

/* The scene's world BVH as a flat host-side array: the layout the reference sketches for its own flat tree
 * (shared/src/hittable_collections/bvh.rs:224-241, `BVHNode::{Root, Node, Leaf}` with u32 child indices and parent links).
 * Node 0 is the Root (parent = -1); a node's children follow it in breadth-first order.
 *   inner node: left / right = child node indices, first = count = 0;  box = the union of its children's boxes (node_bbox)
 *   leaf:       left = right = -1; [first, first + count) is its range of prim_order (shape_index); box = the box around those entries
 * prim_order[k] is the primitive id (position in the world list, the id the batch calls report) of the k-th entry in leaf order.
 * Boxes are the f64 boxes the exact path traverses, whichever builder made the tree (host SAH or device LBVH: the nodes are read
 * back from the device).  Scenes whose bounded entries are walked as a flat list (general scenes with <= 8 of them) export 0 nodes
 * and the list order.  Call with nodes == NULL / prim_order == NULL to query the counts; RTW_E_INVALID if a capacity is too small. */
typedef struct rtw_bvh_node {
    double box_min[3], box_max[3];
    int32_t parent, left, right;
    uint32_t first, count;
    uint32_t depth;                 /* Root = 0 */
} rtw_bvh_node;
RTW_API int  rtw_scene_export_bvh(rtw_scene* scene, rtw_bvh_node* nodes, size_t node_capacity, size_t* n_nodes,
                                  uint32_t* prim_order, size_t prim_capacity, size_t* n_prims);

/* ---- Camera::render ---------------------------------------------------------------------------- */
/* Replaces Camera::render (shared/src/camera.rs:295-297).  Host buffers, both optional:
 *   rgb_sum: [height][width][3] f64, un-normalised sample sums == the Colour inside each
 *            SampledColour (camera.rs:381-387); row j = 0 is the BOTTOM row like the reference.
 *   rgb8:    [height][width][3] u8, resolved like Colour::write_colour (shared/src/colour.rs:15-36),
 *            same row order (bin/src/main.rs:94-98 reverses rows when writing the PPM). */
RTW_API int rtw_render(rtw_scene* scene, const rtw_camera* camera, const rtw_opts* opts,
               double* rgb_sum, uint8_t* rgb8, rtw_stats* stats);

/* Multi-GPU building blocks for callers that bring their own collective (rtw_render_multi / rtw_render_rank below do it inside
 * the library).
 * d_tiles is a DEVICE buffer of rtw_tiles_per_rank() * RTW_TILE_H * RTW_TILE_W * 3 elements
 * (float for RTW_F32, double for RTW_F64) that receives this rank's tiles; `stream` is a
 * cudaStream_t (0 = default stream). */
RTW_API int rtw_render_tiles_device(rtw_scene* scene, const rtw_camera* camera, const rtw_opts* opts,
                            uint32_t rank, uint32_t world, void* d_tiles, void* stream, rtw_stats* stats);
/* d_tiles_all: [world][tiles_per_rank][TILE_H][TILE_W][3] as gathered on the root; writes DEVICE
 * buffers d_rgb_sum ([h][w][3] f64, may be NULL) and d_rgb8 ([h][w][3] u8, may be NULL). */
RTW_API int rtw_untile_resolve_device(const void* d_tiles_all, uint32_t precision, uint32_t width, uint32_t height,
                              uint32_t world, uint32_t samples_per_pixel, double* d_rgb_sum, uint8_t* d_rgb8,
                              void* stream);

/* Sample partition (the better-balanced way to use several GPUs for the FP32 renderers): render samples
 * [sample_begin, sample_begin + sample_count) of EVERY pixel into the caller's fixed-point accumulators —
 *   d_accum:  [rtw_tiles_total() * RTW_TILE_H * RTW_TILE_W][3] u64, radiance sums in 2^-32 units, pixels in tile-slot order (world = 1)
 *   d_poison: [rtw_tiles_total() * RTW_TILE_H * RTW_TILE_W] u32, NaN / overflow flags, one 4-bit field each
 * (both are overwritten).  Integer sums commute and the flag fields do not carry into each other for up to 15 ranks, so SUMMING the
 * buffers of all ranks (one NCCL reduce) and resolving gives the single-GPU image bit for bit, whatever the number of ranks.  Every
 * rank sees every pixel, so the ranks' loads differ only by the rounding of samples_per_pixel / world.  RTW_F32 only (not with
 * RTW_FLAG_LANE_PER_PIXEL); paths are keyed by their absolute sample index. */
RTW_API int rtw_render_samples_device(rtw_scene* scene, const rtw_camera* camera, const rtw_opts* opts, uint32_t sample_begin,
                              uint32_t sample_count, void* d_accum, void* d_poison, void* stream, rtw_stats* stats);
/* (summed) accumulators -> DEVICE d_rgb_sum ([h][w][3] f64, may be NULL) and d_rgb8 ([h][w][3] u8, may be NULL); samples_per_pixel
 * is the camera's total. */
RTW_API int rtw_resolve_accum_device(const void* d_accum, const void* d_poison, uint32_t width, uint32_t height,
                             uint32_t samples_per_pixel, double* d_rgb_sum, uint8_t* d_rgb8, void* stream);

/* ---- N GPUs behind the C ABI ------------------------------------------------------------------------------------------
 * The seam is still Camera::render (camera.rs:295-297, called from bin/src/main.rs:82-86): one blocking call that returns the
 * frame.  Two ways to put N GPUs behind it, both with the frame's single collective INSIDE the library:
 *
 * (A) one process, N GPUs: rtw_render_multi.  The scene is replicated onto devices[0..n_gpus) (NULL: CUDA devices 0..n_gpus-1;
 *     devices[0] is the root and should be the device the scene was created on), every GPU renders its share on its own stream
 *     and the root returns the image.  RTW_F32 (fixed-point renderers): every GPU renders into a whole-image block of fixed-point
 *     accumulators — ALL samples of the pixels it owns when the frame's work queue is ordered (sphere scenes, wavefront renderer,
 *     pinhole camera: chunk c of the queue belongs to GPU (c + (c >> 4)) mod N), else samples [spp*g/N, spp*(g+1)/N) of every pixel
 *     (RTW_MULTI_PARTITION=samples in the environment forces that split) — and the collective is
 *       RTW_COLLECTIVE_PEER: one fused kernel per GPU over NVLink peer memory — GPU g sums pixel slots [S*g/N, S*(g+1)/N) of all N
 *                            accumulator blocks through peer loads, resolves them and stores the pixels into the root's image
 *                            (reduce-scatter + resolve + gather in one pass; needs peer access between all pairs)
 *       RTW_COLLECTIVE_NCCL: one ncclReduce (integer sum) of the blocks to the root (ncclCommInitAll) + resolve there
 *       RTW_COLLECTIVE_AUTO: PEER when every pair of devices has peer access, else NCCL.
 *     RTW_F64 (ordered per-pixel sums): tile partition, tile buffers copied to the root (cudaMemcpyPeerAsync), untile there.
 *     Integer sums commute and samples / tiles are keyed by absolute index, so the image equals rtw_render's bit for bit for any N.
 * (B) one process per GPU (how `torchrun` / `mpirun` launch): rtw_comm_unique_id on rank 0, the 128 bytes travel out of band,
 *     every rank calls rtw_comm_init_rank, then rtw_render_rank on every rank renders that rank's share, runs the one NCCL
 *     collective (ncclReduce of the accumulator blocks / grouped ncclSend+ncclRecv of the tile buffers) and resolves on rank 0,
 *     which alone receives the image.
 * NCCL is loaded at first use with dlopen (RTW_NCCL_LIBRARY, else libnccl.so.2 from the loader path); without it the NCCL
 * collective and the rank API return RTW_E_UNSUPPORTED and everything else works. */
enum { RTW_COLLECTIVE_AUTO = 0, RTW_COLLECTIVE_PEER = 1, RTW_COLLECTIVE_NCCL = 2 };
RTW_API int rtw_render_multi(rtw_scene* scene, const rtw_camera* camera, const rtw_opts* opts, int n_gpus, const int* devices,
                     uint32_t collective, double* rgb_sum, uint8_t* rgb8, rtw_stats* stats);
#define RTW_COMM_ID_BYTES 128
typedef struct rtw_comm rtw_comm;
RTW_API int  rtw_comm_unique_id(uint8_t id[RTW_COMM_ID_BYTES]);
RTW_API int  rtw_comm_init_rank(const uint8_t id[RTW_COMM_ID_BYTES], int rank, int world, rtw_comm** out);  /* collective; binds the current device */
RTW_API void rtw_comm_destroy(rtw_comm* comm);
RTW_API int  rtw_comm_rank(const rtw_comm* comm);
RTW_API int  rtw_comm_world(const rtw_comm* comm);
/* rgb_sum / rgb8: HOST buffers as in rtw_render, written on rank 0 only (other ranks may pass NULL).  stats: this rank's counters. */
RTW_API int  rtw_render_rank(rtw_scene* scene, const rtw_camera* camera, const rtw_opts* opts, rtw_comm* comm,
                     double* rgb_sum, uint8_t* rgb8, rtw_stats* stats);
/* The same with DEVICE outputs on rank 0 and everything enqueued on `stream` (cudaStream_t; 0 = default): returns without
 * synchronising when stats == NULL.  Scratch (accumulators / tile buffers) belongs to the scene handle. */
RTW_API int  rtw_render_rank_device(rtw_scene* scene, const rtw_camera* camera, const rtw_opts* opts, rtw_comm* comm,
                            double* d_rgb_sum, uint8_t* d_rgb8, void* stream, rtw_stats* stats);
/* Asynchronous contract of every *_device entry point called with stats == NULL: the call returns with work in flight on `stream`;
 * "the reference would have panicked here" (a path sampled an empty lights list, hittable_list.rs:414-419) is then reported by the
 * next synchronising call on the scene, or by this query, which synchronises the scene's device, returns RTW_E_INVALID if the flag
 * was raised since the last check (and clears it) and RTW_OK otherwise.  kernel_ms (may be NULL) receives the CUDA-event time of
 * the render kernels of the scene's last render call. */
RTW_API int  rtw_scene_sync(rtw_scene* scene, double* kernel_ms);

/* Progressive rendering and checkpointing with HOST buffers.  rtw_render_samples renders samples
 * [sample_begin, sample_begin + sample_count) of every pixel and ADDS them into the caller's accumulators
 *   accum:  [rtw_accum_slots(w, h)][3] u64 (2^-32 radiance units), poison: [rtw_accum_slots(w, h)] u32 (flags, combined with OR)
 * which the caller zeroes before the first pass and may save / restore between passes (a checkpoint is these two arrays plus the
 * next sample index).  Whatever the split into passes, rtw_resolve_accum yields the image of one rtw_render call bit for bit
 * (RTW_F32 renderers; samples are keyed by their absolute index). */
RTW_API size_t rtw_accum_slots(uint32_t width, uint32_t height);
RTW_API int rtw_render_samples(rtw_scene* scene, const rtw_camera* camera, const rtw_opts* opts, uint32_t sample_begin,
                       uint32_t sample_count, uint64_t* accum, uint32_t* poison, rtw_stats* stats);
RTW_API int rtw_resolve_accum(const uint64_t* accum, const uint32_t* poison, uint32_t width, uint32_t height,
                      uint32_t samples_per_pixel, double* rgb_sum, uint8_t* rgb8);

/* ---- per-ray operations (parity surface) ------------------------------------------------------- */
/* Hittable::hit of the world for a batch of rays (shared/src/hittable.rs:173; bvh.rs:163-188).
 * o, d: [n][3] f64 host arrays; prim_id: -1 = miss; t: +inf on a miss.  precision as in rtw_opts. */
RTW_API int rtw_trace_batch(rtw_scene* scene, const double* o, const double* d, size_t n, double tmin, double tmax,
                    uint32_t precision, int32_t* prim_id, double* t);
/* One path vertex for a batch of rays: world.hit(tmin..inf) + Material::scatter + for Lambertian the
 * MixturePdf sample and weight (camera.rs:473-521), drawing from stream (seed; pixel, sample, vertex).
 * kind: 0 miss, 1 absorbed, 2 specular (Reflect), 3 diffuse (Scatter).  Outputs [n] / [n][3] f64. */
RTW_API int rtw_scatter_batch(rtw_scene* scene, const rtw_opts* opts, const double* o, const double* d, size_t n,
                      const uint32_t* pixel, const uint32_t* sample, const uint32_t* vertex,
                      int32_t* prim_id, double* t, uint32_t* kind, double* p, double* normal,
                      double* dir, double* weight);
/* Material::scatter + the MixturePdf sample / weight (material.rs:357-488, camera.rs:504-521, pdf.rs:34-101) on CALLER-SUPPLIED hit
 * records: the sampling arithmetic alone, independent of how the hit point was traced (north-star check 3 on identical inputs).
 * Per record: d = direction of the incoming ray, p / normal / front_face = the HitRecord (hittable.rs:102-129), mat_kind and
 * material = (albedo r, g, b, param) of the material that was hit; stream (seed; pixel, sample, vertex).  Outputs as rtw_scatter_batch.
 * The scene supplies the lights list.  Sphere-path scenes only (RTW_E_UNSUPPORTED for general scenes: their shaders read textures
 * through the entity that was hit). */
RTW_API int rtw_shade_batch(rtw_scene* scene, const rtw_opts* opts, size_t n, const double* d, const double* p, const double* normal,
                    const uint32_t* front_face, const uint32_t* mat_kind, const double* material,
                    const uint32_t* pixel, const uint32_t* sample, const uint32_t* vertex,
                    uint32_t* kind, double* dir, double* weight);
/* Camera::get_ray (camera.rs:274-293) for a batch of (i, j, sample). */
RTW_API int rtw_get_rays(const rtw_camera* camera, const rtw_opts* opts, const uint32_t* i, const uint32_t* j,
                 const uint32_t* sample, size_t n, double* o, double* d);
/* Radiance of individual paths: ray_colour_call for (i, j, sample) (camera.rs:439-457). */
RTW_API int rtw_path_radiance(rtw_scene* scene, const rtw_camera* camera, const rtw_opts* opts, const uint32_t* i,
                      const uint32_t* j, const uint32_t* sample, size_t n, double* rgb);

#ifdef __cplusplus
}
#endif
#endif /* RTW_H */
