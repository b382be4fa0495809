//! UNVERIFIED (never compiled here: no Rust toolchain in the build image).
//! `cuda` crate of the reference workspace: thin `extern "C"` binding of include/rtw.h plus the safe
//! `render_cuda` that `bin` calls for `--backend cuda` instead of `Camera::render` (shared/src/camera.rs:295).
//!
//! The reference hands `render` type-erased `&dyn Hittable`s whose fields are private, so the scene is passed
//! as plain data.  The additive, non-breaking accessors this needs in the reference crates are listed in
//! INTEGRATION.md (`Sphere::{center,radius,material}`, `Material::describe`, `Camera::raw`).
pub mod philox;

use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)] #[derive(Clone, Copy, Default)] pub struct RtwMaterial { pub kind: u32, pub texture: u32, pub r: f64, pub g: f64, pub b: f64, pub param: f64 }
#[repr(C)] #[derive(Clone, Copy, Default)] pub struct RtwSphere { pub cx: f64, pub cy: f64, pub cz: f64, pub r: f64 }
#[repr(C)] #[derive(Clone, Copy, Default)] pub struct RtwPlane { pub px: f64, pub py: f64, pub pz: f64, pub nx: f64, pub ny: f64, pub nz: f64 }
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct RtwCamera {
    pub center: [f64; 3], pub pixel00_loc: [f64; 3], pub pixel_delta_u: [f64; 3], pub pixel_delta_v: [f64; 3],
    pub defocus_disk_u: [f64; 3], pub defocus_disk_v: [f64; 3], pub background: [f64; 3], pub defocus_angle: f64,
    pub image_width: u32, pub image_height: u32, pub samples_per_pixel: u32, pub max_depth: u32,
}
#[repr(C)] #[derive(Clone, Copy)]
pub struct RtwOpts { pub seed: u64, pub tmin: f64, pub precision: u32, pub mode: u32, pub flags: u32, pub reserved: u32 }
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct RtwStats {
    pub paths: u64, pub rays: u64, pub node_visits: u64, pub sphere_tests: u64, pub light_tests: u64, pub lambertian: u64,
    pub metal: u64, pub dielectric: u64, pub absorbed: u64, pub missed: u64, pub depth_out: u64,
    pub kernel_ms: f64, pub total_ms: f64, pub launches: u32, pub reserved: u32,
}
pub const RTW_F32: u32 = 0; pub const RTW_F64: u32 = 1;
pub const RTW_MEGAKERNEL: u32 = 0; pub const RTW_WAVEFRONT: u32 = 1;
pub const RTW_TMIN_REFERENCE: f64 = -1.0;
pub const RTW_LAMBERTIAN: u32 = 0; pub const RTW_METAL: u32 = 1; pub const RTW_DIELECTRIC: u32 = 2; pub const RTW_INVISIBLE: u32 = 3;
pub const RTW_DIFFUSE_LIGHT: u32 = 4; pub const RTW_ISOTROPIC: u32 = 5;

// ---- general scenes (ABI version 2): Quad, Triangle, Cuboid, Transformed<T>, NoiseTexture -----------------------
#[repr(C)] #[derive(Clone, Copy, Default)] pub struct RtwQuad { pub q: [f64; 3], pub u: [f64; 3], pub v: [f64; 3] }
#[repr(C)] #[derive(Clone, Copy, Default)] pub struct RtwCuboid { pub p: [f64; 3], pub q: [f64; 3] }
#[repr(C)] #[derive(Clone, Copy)] pub struct RtwTransform { pub rotation: [f64; 9], pub translation: [f64; 3] }
#[repr(C)] #[derive(Clone, Copy, Default)] pub struct RtwPrim { pub kind: u32, pub index: u32, pub material: u32, pub transform: i32 }
#[repr(C)] #[derive(Clone, Copy, Default)] pub struct RtwTexture { pub kind: u32, pub perlin: u32, pub scale: f64, pub even: u32, pub odd: u32, pub even_colour: [f64; 3], pub odd_colour: [f64; 3] }
#[repr(C)] #[derive(Clone, Copy)] pub struct RtwPerlin { pub rand_vec: [[f64; 3]; 256], pub perm_x: [u8; 256], pub perm_y: [u8; 256], pub perm_z: [u8; 256] }
pub const RTW_PRIM_SPHERE: u32 = 0; pub const RTW_PRIM_PLANE: u32 = 1; pub const RTW_PRIM_QUAD: u32 = 2; pub const RTW_PRIM_TRIANGLE: u32 = 3;
pub const RTW_PRIM_CUBOID: u32 = 4; pub const RTW_TEX_NOISE: u32 = 1; pub const RTW_TEX_CHECKER: u32 = 2;
#[repr(C)]
pub struct RtwSceneDesc {
    pub spheres: *const RtwSphere, pub n_spheres: u64, pub planes: *const RtwPlane, pub n_planes: u64,
    pub quads: *const RtwQuad, pub n_quads: u64, pub cuboids: *const RtwCuboid, pub n_cuboids: u64,
    pub transforms: *const RtwTransform, pub n_transforms: u64, pub materials: *const RtwMaterial, pub n_materials: u64,
    pub textures: *const RtwTexture, pub n_textures: u64, pub perlins: *const RtwPerlin, pub n_perlins: u64,
    pub world: *const RtwPrim, pub n_world: u64, pub lights: *const RtwPrim, pub n_lights: u64,
    pub world_is_bvh: u32, pub lights_is_bvh: u32,
}
#[link(name = "rtw_cuda")]
unsafe extern "C" {
    pub fn rtw_scene_create_general(desc: *const RtwSceneDesc, out: *mut *mut c_void) -> c_int;
    pub fn rtw_transform_then(a: *const RtwTransform, b: *const RtwTransform, out: *mut RtwTransform);
    pub fn rtw_transform_inverse(a: *const RtwTransform, out: *mut RtwTransform) -> c_int;
    pub fn rtw_rotation(angle_degrees: f64, axis: c_int, out: *mut RtwTransform);
    pub fn rtw_perlin_generate(seed: u64, index: u32, out: *mut RtwPerlin);
    pub fn rtw_accum_slots(width: u32, height: u32) -> usize;
    pub fn rtw_render_samples(scene: *mut c_void, camera: *const RtwCamera, opts: *const RtwOpts, sample_begin: u32, sample_count: u32,
                              accum: *mut u64, poison: *mut u32, stats: *mut RtwStats) -> c_int;
    pub fn rtw_resolve_accum(accum: *const u64, poison: *const u32, width: u32, height: u32, samples_per_pixel: u32, rgb_sum: *mut f64,
                             rgb8: *mut u8) -> c_int;
    pub fn rtw_set_bvh_builder(mode: c_int) -> c_int;           // 0 auto, 1 host SAH, 2 device LBVH
    pub fn rtw_scene_bvh_builder(scene: *const c_void) -> c_int;
    /// flat host mirror of the world BVH (the layout of `hittable_collections::bvh::flat::BVHNode`)
    pub fn rtw_scene_export_bvh(scene: *mut c_void, nodes: *mut RtwBvhNode, node_capacity: usize, n_nodes: *mut usize,
                                prim_order: *mut u32, prim_capacity: usize, n_prims: *mut usize) -> c_int;
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct RtwBvhNode {
    pub box_min: [f64; 3], pub box_max: [f64; 3],
    pub parent: i32, pub left: i32, pub right: i32,
    pub first: u32, pub count: u32, pub depth: u32,
}

#[link(name = "rtw_cuda")]
unsafe extern "C" {
    fn rtw_last_error() -> *const c_char;
    fn rtw_scene_create(spheres: *const RtwSphere, sphere_material: *const u32, n_spheres: usize,
                        planes: *const RtwPlane, plane_material: *const u32, n_planes: usize,
                        materials: *const RtwMaterial, n_materials: usize,
                        lights: *const RtwSphere, n_lights: usize, out: *mut *mut c_void) -> c_int;
    fn rtw_scene_destroy(scene: *mut c_void);
    fn rtw_render(scene: *mut c_void, camera: *const RtwCamera, opts: *const RtwOpts, rgb_sum: *mut f64, rgb8: *mut u8,
                  stats: *mut RtwStats) -> c_int;
}

/// Plain-data scene: what `scenes::simple` builds, seen through the additive accessors.
#[derive(Default)]
pub struct SceneDesc {
    pub spheres: Vec<RtwSphere>, pub sphere_material: Vec<u32>,
    pub planes: Vec<RtwPlane>, pub plane_material: Vec<u32>,
    pub materials: Vec<RtwMaterial>, pub lights: Vec<RtwSphere>,
}

#[derive(Debug)] pub struct CudaError(pub i32, pub String);

fn last_error(code: c_int) -> CudaError {
    let msg = unsafe { CStr::from_ptr(rtw_last_error()) }.to_string_lossy().into_owned();
    CudaError(code, msg)
}

/// Drop-in for `Camera::render`: rows of un-normalised sample sums, row 0 = bottom row (camera.rs:179-188),
/// to be wrapped as `SampledColour::from((Colour, spp))` (colour.rs:138-142) by the caller.
pub fn render_cuda(scene: &SceneDesc, camera: &RtwCamera, seed: u64, precision: u32, mode: u32)
                   -> Result<(Vec<Vec<[f64; 3]>>, RtwStats), CudaError> {
    let mut handle: *mut c_void = std::ptr::null_mut();
    let rc = unsafe {
        rtw_scene_create(scene.spheres.as_ptr(), scene.sphere_material.as_ptr(), scene.spheres.len(),
                         scene.planes.as_ptr(), scene.plane_material.as_ptr(), scene.planes.len(),
                         scene.materials.as_ptr(), scene.materials.len(), scene.lights.as_ptr(), scene.lights.len(), &mut handle)
    };
    if rc != 0 { return Err(last_error(rc)); }
    let (w, h) = (camera.image_width as usize, camera.image_height as usize);
    let mut sum = vec![0f64; w * h * 3];
    let mut stats = RtwStats::default();
    let opts = RtwOpts { seed, tmin: RTW_TMIN_REFERENCE, precision, mode, flags: 0, reserved: 0 };
    let rc = unsafe { rtw_render(handle, camera, &opts, sum.as_mut_ptr(), std::ptr::null_mut(), &mut stats) };
    unsafe { rtw_scene_destroy(handle) };
    if rc != 0 { return Err(last_error(rc)); }
    let rows = (0..h).map(|j| (0..w).map(|i| { let k = (j * w + i) * 3; [sum[k], sum[k + 1], sum[k + 2]] }).collect()).collect();
    Ok((rows, stats))
}
