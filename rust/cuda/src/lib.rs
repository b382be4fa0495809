//! UNVERIFIED (never compiled here: no Rust toolchain in the build image).
//! `cuda` crate of the reference workspace: `extern "C"` binding of EVERY entry point of include/rtw.h (ABI version 3;
//! tests/test_abi_and_host.py::test_rust_binding_declares_every_entry_point keeps this list equal to the header's) plus the safe
//! wrappers `bin` calls for `--backend cuda [--gpus N]` instead of `Camera::render` (shared/src/camera.rs:295-297,
//! bin/src/main.rs:82-86).
//!
//! The reference hands `render` type-erased `&dyn Hittable`s whose fields are private, so the scene crosses the boundary as plain
//! data (`SceneDesc`, filled through the additive `Hittable::export` / `Material::describe` / `Camera::raw` accessors listed in
//! INTEGRATION.md section 3).  `SceneDesc` holds what ANY generator of scenes/src/lib.rs returns — spheres, planes, quads,
//! triangles, cuboids, `Transformed<T>`, every material and texture — and `Scene::new` picks `rtw_scene_create` (the sphere path)
//! when the scene is `scenes::simple`-shaped and `rtw_scene_create_general` otherwise, exactly like the C++ and Python mirrors.
pub mod philox;

use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_void};

// ---- plain-data structs of include/rtw.h ----------------------------------------------------------------------------------
#[repr(C)] #[derive(Clone, Copy, Default, Debug)] pub struct RtwMaterial { pub kind: u32, pub texture: u32, pub r: f64, pub g: f64, pub b: f64, pub param: f64 }
#[repr(C)] #[derive(Clone, Copy, Default, Debug)] pub struct RtwSphere { pub cx: f64, pub cy: f64, pub cz: f64, pub r: f64 }
#[repr(C)] #[derive(Clone, Copy, Default, Debug)] pub struct RtwPlane { pub px: f64, pub py: f64, pub pz: f64, pub nx: f64, pub ny: f64, pub nz: f64 }
#[repr(C)] #[derive(Clone, Copy, Default, Debug)] pub struct RtwQuad { pub q: [f64; 3], pub u: [f64; 3], pub v: [f64; 3] }
#[repr(C)] #[derive(Clone, Copy, Default, Debug)] pub struct RtwCuboid { pub p: [f64; 3], pub q: [f64; 3] }
#[repr(C)] #[derive(Clone, Copy, Debug)] pub struct RtwTransform { pub rotation: [f64; 9], pub translation: [f64; 3] }
#[repr(C)] #[derive(Clone, Copy, Default, Debug)] pub struct RtwPrim { pub kind: u32, pub index: u32, pub material: u32, pub transform: i32 }
#[repr(C)] #[derive(Clone, Copy, Default, Debug)]
pub struct RtwTexture { pub kind: u32, pub perlin: u32, pub scale: f64, pub even: u32, pub odd: u32, pub even_colour: [f64; 3], pub odd_colour: [f64; 3] }
#[repr(C)] #[derive(Clone, Copy)]
pub struct RtwPerlin { pub rand_vec: [[f64; 3]; 256], pub perm_x: [u8; 256], pub perm_y: [u8; 256], pub perm_z: [u8; 256] }
#[repr(C)]
pub struct RtwSceneDesc {
    pub spheres: *const RtwSphere, pub n_spheres: u64, pub planes: *const RtwPlane, pub n_planes: u64,
    pub quads: *const RtwQuad, pub n_quads: u64, pub cuboids: *const RtwCuboid, pub n_cuboids: u64,
    pub transforms: *const RtwTransform, pub n_transforms: u64, pub materials: *const RtwMaterial, pub n_materials: u64,
    pub textures: *const RtwTexture, pub n_textures: u64, pub perlins: *const RtwPerlin, pub n_perlins: u64,
    pub world: *const RtwPrim, pub n_world: u64, pub lights: *const RtwPrim, pub n_lights: u64,
    pub world_is_bvh: u32, pub lights_is_bvh: u32,
}
#[repr(C)] #[derive(Clone, Copy, Default, Debug)]
pub struct RtwCamera {
    pub center: [f64; 3], pub pixel00_loc: [f64; 3], pub pixel_delta_u: [f64; 3], pub pixel_delta_v: [f64; 3],
    pub defocus_disk_u: [f64; 3], pub defocus_disk_v: [f64; 3], pub background: [f64; 3], pub defocus_angle: f64,
    pub image_width: u32, pub image_height: u32, pub samples_per_pixel: u32, pub max_depth: u32,
}
#[repr(C)] #[derive(Clone, Copy, Default, Debug)]
pub struct RtwCameraBuilder {
    pub aspect_ratio: f64, pub has_aspect_ratio: u32, pub image_width: u32, pub has_image_width: u32, pub image_height: u32,
    pub has_image_height: u32, pub samples_per_pixel: u32, pub max_depth: u32, pub background: [f64; 3], pub vfov: f64,
    pub lookfrom: [f64; 3], pub lookat: [f64; 3], pub vup: [f64; 3], pub defocus_angle: f64, pub focus_dist: f64,
}
#[repr(C)] #[derive(Clone, Copy, Debug)]
pub struct RtwOpts { pub seed: u64, pub tmin: f64, pub precision: u32, pub mode: u32, pub flags: u32, pub reserved: u32 }
#[repr(C)] #[derive(Clone, Copy, Default, Debug)]
pub struct RtwStats {
    pub paths: u64, pub rays: u64, pub node_visits: u64, pub sphere_tests: u64, pub light_tests: u64, pub lambertian: u64,
    pub metal: u64, pub dielectric: u64, pub absorbed: u64, pub missed: u64, pub depth_out: u64,
    pub kernel_ms: f64, pub total_ms: f64, pub launches: u32, pub reserved: u32,
}
#[repr(C)] #[derive(Clone, Copy, Debug, Default)]
pub struct RtwBvhNode {
    pub box_min: [f64; 3], pub box_max: [f64; 3], pub parent: i32, pub left: i32, pub right: i32,
    pub first: u32, pub count: u32, pub depth: u32,
}

pub const RTW_ABI_VERSION: c_int = 3;
pub const RTW_OK: c_int = 0; pub const RTW_E_INVALID: c_int = -1; pub const RTW_E_CUDA: c_int = -2; pub const RTW_E_NO_DEVICE: c_int = -3;
pub const RTW_E_UNSUPPORTED: c_int = -4; pub const RTW_E_NOMEM: c_int = -5;
pub const RTW_F32: u32 = 0; pub const RTW_F64: u32 = 1;
pub const RTW_MEGAKERNEL: u32 = 0; pub const RTW_WAVEFRONT: u32 = 1;
pub const RTW_TMIN_REFERENCE: f64 = -1.0;
pub const RTW_FLAG_FIX_NAN: u32 = 1; pub const RTW_FLAG_COUNT_EVENTS: u32 = 2; pub const RTW_FLAG_LANE_PER_PIXEL: u32 = 4; pub const RTW_FLAG_NO_CANDIDATES: u32 = 8;
pub const RTW_LAMBERTIAN: u32 = 0; pub const RTW_METAL: u32 = 1; pub const RTW_DIELECTRIC: u32 = 2; pub const RTW_INVISIBLE: u32 = 3;
pub const RTW_DIFFUSE_LIGHT: u32 = 4; pub const RTW_ISOTROPIC: u32 = 5;
pub const RTW_PRIM_SPHERE: u32 = 0; pub const RTW_PRIM_PLANE: u32 = 1; pub const RTW_PRIM_QUAD: u32 = 2; pub const RTW_PRIM_TRIANGLE: u32 = 3;
pub const RTW_PRIM_CUBOID: u32 = 4; pub const RTW_TEX_NOISE: u32 = 1; pub const RTW_TEX_CHECKER: u32 = 2;
pub const RTW_BVH_AUTO: c_int = 0; pub const RTW_BVH_HOST_SAH: c_int = 1; pub const RTW_BVH_DEVICE_LBVH: c_int = 2;
pub const RTW_COLLECTIVE_AUTO: u32 = 0; pub const RTW_COLLECTIVE_PEER: u32 = 1; pub const RTW_COLLECTIVE_NCCL: u32 = 2;
pub const RTW_COMM_ID_BYTES: usize = 128;

// ---- every entry point of include/rtw.h, in header order ---------------------------------------------------------------------
#[link(name = "rtw_cuda")]
unsafe extern "C" {
    pub fn rtw_abi_version() -> c_int;
    pub fn rtw_last_error() -> *const c_char;
    pub fn rtw_camera_build(builder: *const RtwCameraBuilder, out: *mut RtwCamera) -> c_int;
    pub fn rtw_philox4x32_10(ctr: *const u32, key: *const u32, out: *mut u32);
    pub fn rtw_tiles_total(width: u32, height: u32) -> u32;
    pub fn rtw_tiles_per_rank(width: u32, height: u32, world: u32) -> u32;
    pub fn rtw_transform_then(a: *const RtwTransform, b: *const RtwTransform, out: *mut RtwTransform);
    pub fn rtw_transform_inverse(a: *const RtwTransform, out: *mut RtwTransform) -> c_int;
    pub fn rtw_rotation(angle_degrees: f64, axis: c_int, out: *mut RtwTransform);
    pub fn rtw_perlin_generate(seed: u64, index: u32, out: *mut RtwPerlin);
    pub fn rtw_device_count() -> c_int;
    pub fn rtw_release_cached_memory() -> c_int;
    pub fn rtw_scene_create(spheres: *const RtwSphere, sphere_material: *const u32, n_spheres: usize,
                            planes: *const RtwPlane, plane_material: *const u32, n_planes: usize,
                            materials: *const RtwMaterial, n_materials: usize,
                            lights: *const RtwSphere, n_lights: usize, out: *mut *mut c_void) -> c_int;
    pub fn rtw_scene_create_general(desc: *const RtwSceneDesc, out: *mut *mut c_void) -> c_int;
    pub fn rtw_set_bvh_builder(mode: c_int) -> c_int;
    pub fn rtw_scene_bvh_builder(scene: *const c_void) -> c_int;
    pub fn rtw_scene_destroy(scene: *mut c_void);
    pub fn rtw_scene_info(scene: *const c_void, out: *mut u64) -> c_int;
    pub fn rtw_scene_export_bvh(scene: *mut c_void, nodes: *mut RtwBvhNode, node_capacity: usize, n_nodes: *mut usize,
                                prim_order: *mut u32, prim_capacity: usize, n_prims: *mut usize) -> c_int;
    pub fn rtw_render(scene: *mut c_void, camera: *const RtwCamera, opts: *const RtwOpts, rgb_sum: *mut f64, rgb8: *mut u8,
                      stats: *mut RtwStats) -> c_int;
    pub fn rtw_render_tiles_device(scene: *mut c_void, camera: *const RtwCamera, opts: *const RtwOpts, rank: u32, world: u32,
                                   d_tiles: *mut c_void, stream: *mut c_void, stats: *mut RtwStats) -> c_int;
    pub fn rtw_untile_resolve_device(d_tiles_all: *const c_void, precision: u32, width: u32, height: u32, world: u32,
                                     samples_per_pixel: u32, d_rgb_sum: *mut f64, d_rgb8: *mut u8, stream: *mut c_void) -> c_int;
    pub fn rtw_render_samples_device(scene: *mut c_void, camera: *const RtwCamera, opts: *const RtwOpts, sample_begin: u32,
                                     sample_count: u32, d_accum: *mut c_void, d_poison: *mut c_void, stream: *mut c_void,
                                     stats: *mut RtwStats) -> c_int;
    pub fn rtw_resolve_accum_device(d_accum: *const c_void, d_poison: *const c_void, width: u32, height: u32, samples_per_pixel: u32,
                                    d_rgb_sum: *mut f64, d_rgb8: *mut u8, stream: *mut c_void) -> c_int;
    pub fn rtw_render_multi(scene: *mut c_void, camera: *const RtwCamera, opts: *const RtwOpts, n_gpus: c_int, devices: *const c_int,
                            collective: u32, rgb_sum: *mut f64, rgb8: *mut u8, stats: *mut RtwStats) -> c_int;
    pub fn rtw_comm_unique_id(id: *mut u8) -> c_int;
    pub fn rtw_comm_init_rank(id: *const u8, rank: c_int, world: c_int, out: *mut *mut c_void) -> c_int;
    pub fn rtw_comm_destroy(comm: *mut c_void);
    pub fn rtw_comm_rank(comm: *const c_void) -> c_int;
    pub fn rtw_comm_world(comm: *const c_void) -> c_int;
    pub fn rtw_render_rank(scene: *mut c_void, camera: *const RtwCamera, opts: *const RtwOpts, comm: *mut c_void, rgb_sum: *mut f64,
                           rgb8: *mut u8, stats: *mut RtwStats) -> c_int;
    pub fn rtw_render_rank_device(scene: *mut c_void, camera: *const RtwCamera, opts: *const RtwOpts, comm: *mut c_void,
                                  d_rgb_sum: *mut f64, d_rgb8: *mut u8, stream: *mut c_void, stats: *mut RtwStats) -> c_int;
    pub fn rtw_scene_sync(scene: *mut c_void, kernel_ms: *mut f64) -> c_int;
    pub fn rtw_accum_slots(width: u32, height: u32) -> usize;
    pub fn rtw_render_samples(scene: *mut c_void, camera: *const RtwCamera, opts: *const RtwOpts, sample_begin: u32, sample_count: u32,
                              accum: *mut u64, poison: *mut u32, stats: *mut RtwStats) -> c_int;
    pub fn rtw_resolve_accum(accum: *const u64, poison: *const u32, width: u32, height: u32, samples_per_pixel: u32, rgb_sum: *mut f64,
                             rgb8: *mut u8) -> c_int;
    pub fn rtw_trace_batch(scene: *mut c_void, o: *const f64, d: *const f64, n: usize, tmin: f64, tmax: f64, precision: u32,
                           prim_id: *mut i32, t: *mut f64) -> c_int;
    pub fn rtw_scatter_batch(scene: *mut c_void, opts: *const RtwOpts, o: *const f64, d: *const f64, n: usize, pixel: *const u32,
                             sample: *const u32, vertex: *const u32, prim_id: *mut i32, t: *mut f64, kind: *mut u32, p: *mut f64,
                             normal: *mut f64, dir: *mut f64, weight: *mut f64) -> c_int;
    pub fn rtw_shade_batch(scene: *mut c_void, opts: *const RtwOpts, n: usize, d: *const f64, p: *const f64, normal: *const f64,
                           front_face: *const u32, mat_kind: *const u32, material: *const f64, pixel: *const u32, sample: *const u32,
                           vertex: *const u32, kind: *mut u32, dir: *mut f64, weight: *mut f64) -> c_int;
    pub fn rtw_get_rays(camera: *const RtwCamera, opts: *const RtwOpts, i: *const u32, j: *const u32, sample: *const u32, n: usize,
                        o: *mut f64, d: *mut f64) -> c_int;
    pub fn rtw_path_radiance(scene: *mut c_void, camera: *const RtwCamera, opts: *const RtwOpts, i: *const u32, j: *const u32,
                             sample: *const u32, n: usize, rgb: *mut f64) -> c_int;
}

// ---- safe layer ----------------------------------------------------------------------------------------------------------------
#[derive(Debug)] pub struct CudaError(pub i32, pub String);
fn check(code: c_int) -> Result<(), CudaError> {
    if code == RTW_OK { return Ok(()); }
    let msg = unsafe { CStr::from_ptr(rtw_last_error()) }.to_string_lossy().into_owned();
    Err(CudaError(code, msg))
}

/// What any `SceneGenerator` of scenes/src/lib.rs returns, as plain data: two `HittableList`s (`world`, `lights`) of entries over
/// shared entity arrays.  Filled by `Hittable::export(&mut SceneDesc)`; `lights_is_bvh` records that the
/// lights were wrapped in `BoundedVolumeHierarchy::from` (it changes the f64 rounding of `pdf_value`, bvh.rs:67-76).
#[derive(Default)]
pub struct SceneDesc {
    pub spheres: Vec<RtwSphere>, pub planes: Vec<RtwPlane>, pub quads: Vec<RtwQuad>, pub cuboids: Vec<RtwCuboid>,
    pub transforms: Vec<RtwTransform>, pub materials: Vec<RtwMaterial>, pub textures: Vec<RtwTexture>, pub perlins: Vec<RtwPerlin>,
    pub world: Vec<RtwPrim>, pub lights: Vec<RtwPrim>,
    pub world_is_bvh: bool, pub lights_is_bvh: bool,
}

impl SceneDesc {
    /// `HittableList::is_simple` of the C++ / Python mirrors: spheres and planes through the origin (`Plane::get_aabbox` puts an
    /// axis-aligned plane's box through the origin, plane.rs:78-107), untransformed, SolidColour Lambertian / Metal / Dialectric /
    /// Invisible materials, sphere lights — the shape of `scenes::simple`, served by the shared-memory sphere kernels.
    pub fn is_simple(&self) -> bool {
        let entry_ok = |e: &RtwPrim| {
            let m = &self.materials[e.material as usize];
            e.transform < 0 && m.texture == 0 && m.kind <= RTW_INVISIBLE && match e.kind {
                RTW_PRIM_SPHERE => true,
                RTW_PRIM_PLANE => {
                    let p = &self.planes[e.index as usize];
                    let len = (p.nx * p.nx + p.ny * p.ny + p.nz * p.nz).sqrt();
                    let n = [p.nx / len, p.ny / len, p.nz / len];
                    let pt = [p.px, p.py, p.pz];
                    (0..3).all(|a| !(n[(a + 1) % 3].abs() < f64::EPSILON && n[(a + 2) % 3].abs() < f64::EPSILON && pt[a] != 0.0))
                }
                _ => false,
            }
        };
        self.world.iter().all(entry_ok) && !self.lights_is_bvh
            && self.lights.iter().all(|e| e.kind == RTW_PRIM_SPHERE && e.transform < 0)
    }
}

/// Owned scene handle on the current CUDA device.
pub struct Scene { handle: *mut c_void }
unsafe impl Send for Scene {}

impl Scene {
    pub fn new(desc: &SceneDesc) -> Result<Scene, CudaError> {
        let mut handle: *mut c_void = std::ptr::null_mut();
        if desc.is_simple() {
            // planes first, then spheres: the primitive ids the batch calls report (rtw.h)
            let spheres: Vec<RtwSphere> = desc.world.iter().filter(|e| e.kind == RTW_PRIM_SPHERE).map(|e| desc.spheres[e.index as usize]).collect();
            let sphere_material: Vec<u32> = desc.world.iter().filter(|e| e.kind == RTW_PRIM_SPHERE).map(|e| e.material).collect();
            let planes: Vec<RtwPlane> = desc.world.iter().filter(|e| e.kind == RTW_PRIM_PLANE).map(|e| desc.planes[e.index as usize]).collect();
            let plane_material: Vec<u32> = desc.world.iter().filter(|e| e.kind == RTW_PRIM_PLANE).map(|e| e.material).collect();
            let lights: Vec<RtwSphere> = desc.lights.iter().map(|e| desc.spheres[e.index as usize]).collect();
            check(unsafe {
                rtw_scene_create(spheres.as_ptr(), sphere_material.as_ptr(), spheres.len(), planes.as_ptr(), plane_material.as_ptr(), planes.len(),
                                 desc.materials.as_ptr(), desc.materials.len(), lights.as_ptr(), lights.len(), &mut handle)
            })?;
        } else {
            let d = RtwSceneDesc {
                spheres: desc.spheres.as_ptr(), n_spheres: desc.spheres.len() as u64, planes: desc.planes.as_ptr(), n_planes: desc.planes.len() as u64,
                quads: desc.quads.as_ptr(), n_quads: desc.quads.len() as u64, cuboids: desc.cuboids.as_ptr(), n_cuboids: desc.cuboids.len() as u64,
                transforms: desc.transforms.as_ptr(), n_transforms: desc.transforms.len() as u64,
                materials: desc.materials.as_ptr(), n_materials: desc.materials.len() as u64,
                textures: desc.textures.as_ptr(), n_textures: desc.textures.len() as u64, perlins: desc.perlins.as_ptr(), n_perlins: desc.perlins.len() as u64,
                world: desc.world.as_ptr(), n_world: desc.world.len() as u64, lights: desc.lights.as_ptr(), n_lights: desc.lights.len() as u64,
                world_is_bvh: desc.world_is_bvh as u32, lights_is_bvh: desc.lights_is_bvh as u32,
            };
            check(unsafe { rtw_scene_create_general(&d, &mut handle) })?;
        }
        Ok(Scene { handle })
    }
    pub fn raw(&self) -> *mut c_void { self.handle }
}
impl Drop for Scene { fn drop(&mut self) { unsafe { rtw_scene_destroy(self.handle) } } }

#[derive(Clone, Copy, Debug)]
pub struct RenderOptions { pub seed: u64, pub tmin: f64, pub precision: u32, pub mode: u32, pub flags: u32, pub n_gpus: i32, pub collective: u32 }
impl Default for RenderOptions {
    fn default() -> Self { RenderOptions { seed: 20261018, tmin: RTW_TMIN_REFERENCE, precision: RTW_F32, mode: RTW_WAVEFRONT, flags: 0, n_gpus: 1, collective: RTW_COLLECTIVE_AUTO } }
}

/// Drop-in for `Camera::render` (`--backend cuda [--gpus N]`): rows of un-normalised sample sums, row 0 = the bottom row
/// (camera.rs:179-188), to be wrapped as `SampledColour::from((Colour, spp))` (colour.rs:138-142) by the caller.  With
/// `n_gpus > 1` the frame is rendered by `rtw_render_multi` on CUDA devices 0..n_gpus of this process: same image bit for bit.
pub fn render_cuda(scene: &SceneDesc, camera: &RtwCamera, opt: &RenderOptions) -> Result<(Vec<Vec<[f64; 3]>>, RtwStats), CudaError> {
    let sc = Scene::new(scene)?;
    let (w, h) = (camera.image_width as usize, camera.image_height as usize);
    let mut sum = vec![0f64; w * h * 3];
    let mut stats = RtwStats::default();
    let opts = RtwOpts { seed: opt.seed, tmin: opt.tmin, precision: opt.precision, mode: opt.mode, flags: opt.flags, reserved: 0 };
    check(unsafe {
        if opt.n_gpus > 1 {
            rtw_render_multi(sc.raw(), camera, &opts, opt.n_gpus, std::ptr::null(), opt.collective, sum.as_mut_ptr(), std::ptr::null_mut(), &mut stats)
        } else {
            rtw_render(sc.raw(), camera, &opts, sum.as_mut_ptr(), std::ptr::null_mut(), &mut stats)
        }
    })?;
    let rows = (0..h).map(|j| (0..w).map(|i| { let k = (j * w + i) * 3; [sum[k], sum[k + 1], sum[k + 2]] }).collect()).collect();
    Ok((rows, stats))
}

/// One process per GPU (an MPI-style launcher): rank 0 makes the id, every rank joins, `render` returns the rows on rank 0.
pub struct Comm { handle: *mut c_void }
impl Comm {
    pub fn unique_id() -> Result<[u8; RTW_COMM_ID_BYTES], CudaError> {
        let mut id = [0u8; RTW_COMM_ID_BYTES];
        check(unsafe { rtw_comm_unique_id(id.as_mut_ptr()) })?;
        Ok(id)
    }
    pub fn init_rank(id: &[u8; RTW_COMM_ID_BYTES], rank: i32, world: i32) -> Result<Comm, CudaError> {
        let mut handle: *mut c_void = std::ptr::null_mut();
        check(unsafe { rtw_comm_init_rank(id.as_ptr(), rank, world, &mut handle) })?;
        Ok(Comm { handle })
    }
    pub fn rank(&self) -> i32 { unsafe { rtw_comm_rank(self.handle) } }
    pub fn world(&self) -> i32 { unsafe { rtw_comm_world(self.handle) } }
    pub fn render(&self, scene: &Scene, camera: &RtwCamera, opt: &RenderOptions) -> Result<(Option<Vec<f64>>, RtwStats), CudaError> {
        let npx = camera.image_width as usize * camera.image_height as usize;
        let mut sum = if self.rank() == 0 { Some(vec![0f64; npx * 3]) } else { None };
        let mut stats = RtwStats::default();
        let opts = RtwOpts { seed: opt.seed, tmin: opt.tmin, precision: opt.precision, mode: opt.mode, flags: opt.flags, reserved: 0 };
        let p = sum.as_mut().map(|v| v.as_mut_ptr()).unwrap_or(std::ptr::null_mut());
        check(unsafe { rtw_render_rank(scene.raw(), camera, &opts, self.handle, p, std::ptr::null_mut(), &mut stats) })?;
        Ok((sum, stats))
    }
}
impl Drop for Comm { fn drop(&mut self) { unsafe { rtw_comm_destroy(self.handle) } } }
