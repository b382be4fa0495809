// rtw_host.hpp — host side of the CUDA backend, mirroring the reference's own interface for the render
// path (same type and method names, argument meaning and error behaviour) on top of the C ABI in
// include/rtw.h.  The reference is Rust; no Rust toolchain exists in this image, so the host layer is
// C++ here and the (unverified) Rust binding lives in rust/cuda/ — see INTEGRATION.md.
//
//   reference (Rust)                                   here (C++)
//   geometry::vec3::{Vec3, Point3}                     rtw_host::Vec3 / Point3
//   shared::colour::Colour, SampledColour              rtw_host::Colour, SampledColour
//   shared::material::{Lambertian, Metal, Dialectric}  rtw_host::Lambertian / Metal / Dialectric, INVISIBLE
//   shared::entities::{Sphere, Plane}                  rtw_host::Sphere / Plane
//   hittable_collections::hittable_list::HittableList  rtw_host::HittableList        (add, len)
//   hittable_collections::bvh::BoundedVolumeHierarchy  rtw_host::BoundedVolumeHierarchy::from(list)
//   shared::camera::{CameraBuilder, Camera}            rtw_host::CameraBuilder (with_*, build), Camera::render
//   scenes::simple                                     rtw_host::scenes::simple(seed)   (seeded: the reference's is not)
//   shared::entities::{Quad, Triangle, Cuboid}         rtw_host::Quad / Triangle / Cuboid
//   geometry::transformations::{Transformation,        rtw_host::Transformation, rotation(), Translation3, entity.transform(t)
//     Transformed<T>, rotation}, vec3::Translation3       -> Transformed<T>
//   shared::material::{DiffuseLight, Isotropic}        rtw_host::DiffuseLight / Isotropic
//   shared::texture::{SolidColour, NoiseTexture,       rtw_host::SolidColour / NoiseTexture (Perlin tables seeded) / CheckerTexture
//     CheckerTexture}
//   scenes::{simple_light, cornell_box, debugging_     rtw_host::scenes::{simple_light, cornell_box, debugging_scene, simple_transform,
//     scene, simple_transform, checkered_spheres}         checkered_spheres}
//
// Where the reference panics (unwrap / expect), this layer throws std::runtime_error.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <functional>
#include <memory>
#include <optional>
#include <sstream>
#include <stdexcept>
#include <string>
#include <tuple>
#include <type_traits>
#include <vector>

#include "../../include/rtw.h"

namespace rtw_host {

struct Vec3 {                                       // geometry/src/vec3/vec.rs:9-15
    double x = 0., y = 0., z = 0.;
    Vec3() = default;
    Vec3(double x_, double y_, double z_) : x(x_), y(y_), z(z_) {}
    Vec3 operator-(const Vec3& o) const { return {x - o.x, y - o.y, z - o.z}; }
    Vec3 operator*(const Vec3& o) const { return {x * o.x, y * o.y, z * o.z}; }
    double length() const { return std::sqrt(x * x + y * y + z * z); }
};
using Point3 = Vec3;

struct Colour {                                     // shared/src/colour.rs:6-8
    Vec3 v;
    Colour() = default;
    Colour(double r, double g, double b) : v(r, g, b) {}
};

// shared/src/colour.rs:136-148; Display == write_colour (:15-36)
struct SampledColour {
    Colour colour; int32_t samples = 1; uint8_t rgb[3] = {0, 0, 0};
    std::string to_string() const { return std::to_string(rgb[0]) + " " + std::to_string(rgb[1]) + " " + std::to_string(rgb[2]); }
};

// ---- textures (shared/src/texture.rs) ------------------------------------------------------------------
struct Texture;
using TexturePtr = std::shared_ptr<const Texture>;
struct Texture {
    uint32_t kind = 0;                              // 0 SolidColour, RTW_TEX_NOISE, RTW_TEX_CHECKER
    Colour colour;                                  // SolidColour (texture.rs:15-22)
    double scale = 1.; uint64_t seed = 20261018; uint32_t index = 0;    // NoiseTexture (:57-102): Perlin tables from Philox stream (seed; 0x9E71A000 + index)
    TexturePtr even, odd;                           // CheckerTexture (:24-55)
};
struct SolidColour { static TexturePtr new_(Colour c) { auto t = std::make_shared<Texture>(); t->colour = c; return t; } };
struct NoiseTexture {
    static TexturePtr new_(double scale, uint64_t seed = 20261018, uint32_t index = 0) {
        auto t = std::make_shared<Texture>(); t->kind = RTW_TEX_NOISE; t->scale = scale; t->seed = seed; t->index = index; return t;
    }
};
struct CheckerTexture {
    static TexturePtr new_(TexturePtr even, TexturePtr odd, double scale) {
        auto t = std::make_shared<Texture>(); t->kind = RTW_TEX_CHECKER; t->scale = scale; t->even = std::move(even); t->odd = std::move(odd); return t;
    }
    static TexturePtr new_with_colours(Colour even, Colour odd, double scale) { return new_(SolidColour::new_(even), SolidColour::new_(odd), scale); }
};

// ---- materials (shared/src/material.rs) -----------------------------------------------------------
struct Material { rtw_material pod{}; TexturePtr texture; };       // texture == nullptr (or a SolidColour texture): SolidColour(pod.r, pod.g, pod.b)
using MaterialPtr = std::shared_ptr<const Material>;
inline MaterialPtr make_material(uint32_t kind, Colour c, double param, TexturePtr tex = nullptr) {
    auto m = std::make_shared<Material>();
    if (tex && tex->kind == 0) { c = tex->colour; tex = nullptr; }
    m->pod.kind = kind; m->pod.r = c.v.x; m->pod.g = c.v.y; m->pod.b = c.v.z; m->pod.param = param; m->texture = std::move(tex);
    return m;
}
struct Lambertian {                                                                                                     // material.rs:331-355
    static MaterialPtr new_with_colour(Colour c) { return make_material(RTW_LAMBERTIAN, c, 0.); }
    static MaterialPtr new_(TexturePtr t) { return make_material(RTW_LAMBERTIAN, Colour(), 0., std::move(t)); }
};
struct Metal { static MaterialPtr new_(Colour albedo, double fuzz) { return make_material(RTW_METAL, albedo, fuzz); } }; // material.rs:401-405
struct Dialectric { static MaterialPtr new_(double ior) { return make_material(RTW_DIELECTRIC, Colour(1., 1., 1.), ior); } }; // material.rs:443-448
struct DiffuseLight {                                                                                                   // material.rs:498-504
    static MaterialPtr new_with_colour(Colour c) { return make_material(RTW_DIFFUSE_LIGHT, c, 0.); }
    static MaterialPtr new_(TexturePtr t) { return make_material(RTW_DIFFUSE_LIGHT, Colour(), 0., std::move(t)); }
};
struct Isotropic {                                                                                                      // material.rs:521-527
    static MaterialPtr new_with_colour(Colour c) { return make_material(RTW_ISOTROPIC, c, 0.); }
    static MaterialPtr new_(TexturePtr t) { return make_material(RTW_ISOTROPIC, Colour(), 0., std::move(t)); }
};
inline MaterialPtr invisible() { static MaterialPtr p = make_material(RTW_INVISIBLE, Colour(0., 0., 0.), 0.); return p; }  // material.rs:319-322

// ---- transformations (geometry/src/transformations.rs, the default non-euclid build) --------------------
enum class Axis : int { X = 0, Y = 1, Z = 2 };
struct Transformation {                             // :96-136; arithmetic in the library so every host language agrees bit for bit
    rtw_transform pod{{1., 0., 0., 0., 1., 0., 0., 0., 1.}, {0., 0., 0.}};
    Transformation then(const Transformation& t) const { Transformation o; rtw_transform_then(&pod, &t.pod, &o.pod); return o; }
    std::optional<Transformation> inverse() const { Transformation o; if (!rtw_transform_inverse(&pod, &o.pod)) return std::nullopt; return o; }
};
inline Transformation Translation3(double x, double y, double z) { Transformation t; t.pod.translation[0] = x; t.pod.translation[1] = y; t.pod.translation[2] = z; return t; }
inline Transformation rotation(double angle_degrees, Axis axis) { Transformation t; rtw_rotation(angle_degrees, (int)axis, &t.pod); return t; }

template <class T> struct Transformed {             // :168-222
    T instance; Transformation transformation;
    Transformed transform(const Transformation& t) const { return Transformed{instance, transformation.then(t)}; }
};
template <class Self> struct Transformable {        // Transformable::transform, :193-222
    Transformed<Self> transform(const Transformation& t) const { return Transformed<Self>{static_cast<const Self&>(*this), Transformation().then(t)}; }
};

// ---- entities --------------------------------------------------------------------------------------
struct Sphere : Transformable<Sphere> { Point3 center; double radius = 0.; MaterialPtr mat;      // entities/sphere.rs:25-47
    static Sphere new_(Point3 c, double r, MaterialPtr m) { Sphere s; s.center = c; s.radius = r; s.mat = std::move(m); return s; } };
struct Plane { Point3 point; Vec3 normal; MaterialPtr mat;                           // entities/plane.rs:21-39
    static Plane new_(Point3 p, Vec3 n, MaterialPtr m) { return Plane{p, n, std::move(m)}; } };
struct Quad : Transformable<Quad> { Point3 q; Vec3 u, v; MaterialPtr mat;            // entities/quadrilateral.rs:23-56
    static Quad new_(Point3 q, Vec3 u, Vec3 v, MaterialPtr m) { Quad s; s.q = q; s.u = u; s.v = v; s.mat = std::move(m); return s; } };
struct Triangle : Transformable<Triangle> { Point3 q; Vec3 u, v; MaterialPtr mat;    // entities/triangles.rs:23-54
    static Triangle new_(Point3 q, Vec3 u, Vec3 v, MaterialPtr m) { Triangle s; s.q = q; s.u = u; s.v = v; s.mat = std::move(m); return s; } };
struct Cuboid : Transformable<Cuboid> { Point3 p, q; MaterialPtr mat;                // entities/cuboid.rs:21-50
    static Cuboid new_(Point3 p, Point3 q, MaterialPtr m) { Cuboid s; s.p = p; s.q = q; s.mat = std::move(m); return s; } };

// One HittableList entry in insertion order (what rtw_prim + the entity arrays of rtw_scene_desc are filled from).
struct Entry {
    uint32_t kind = RTW_PRIM_SPHERE;
    double a[3] = {0, 0, 0}, b[3] = {0, 0, 0}, c[3] = {0, 0, 0}; double r = 0.;     // sphere: a = centre, r; plane: a = point, b = normal;
    MaterialPtr mat;                                                                 // quad / triangle: a = q, b = u, c = v; cuboid: a = p, b = q
    std::optional<Transformation> transform;
};

// hittable_collections/hittable_list.rs:247-294
class HittableList {
public:
    void add(const Sphere& s) { spheres_.push_back(s); push(RTW_PRIM_SPHERE, s.center, Vec3(), Vec3(), s.radius, s.mat, std::nullopt); }
    void add(const Plane& p) { planes_.push_back(p); push(RTW_PRIM_PLANE, p.point, p.normal, Vec3(), 0., p.mat, std::nullopt); }
    void add(const Quad& q) { push(RTW_PRIM_QUAD, q.q, q.u, q.v, 0., q.mat, std::nullopt); }
    void add(const Triangle& q) { push(RTW_PRIM_TRIANGLE, q.q, q.u, q.v, 0., q.mat, std::nullopt); }
    void add(const Cuboid& c) { push(RTW_PRIM_CUBOID, c.p, c.q, Vec3(), 0., c.mat, std::nullopt); }
    void add(const Transformed<Sphere>& t) { push(RTW_PRIM_SPHERE, t.instance.center, Vec3(), Vec3(), t.instance.radius, t.instance.mat, t.transformation); }
    void add(const Transformed<Quad>& t) { push(RTW_PRIM_QUAD, t.instance.q, t.instance.u, t.instance.v, 0., t.instance.mat, t.transformation); }
    void add(const Transformed<Triangle>& t) { push(RTW_PRIM_TRIANGLE, t.instance.q, t.instance.u, t.instance.v, 0., t.instance.mat, t.transformation); }
    void add(const Transformed<Cuboid>& t) { push(RTW_PRIM_CUBOID, t.instance.p, t.instance.q, Vec3(), 0., t.instance.mat, t.transformation); }
    size_t len() const { return entries_.size(); }
    bool is_empty() const { return len() == 0; }
    const std::vector<Sphere>& spheres() const { return spheres_; }
    const std::vector<Plane>& planes() const { return planes_; }
    const std::vector<Entry>& entries() const { return entries_; }
    // spheres + planes through the origin with SolidColour Lambertian / Metal / Dialectric / Invisible: the fast sphere path
    bool is_simple() const {
        for (const Entry& e : entries_) {
            if (e.kind > RTW_PRIM_PLANE || e.transform || !e.mat || e.mat->pod.kind > RTW_INVISIBLE || e.mat->texture) return false;
            if (e.kind == RTW_PRIM_PLANE) {
                double len = std::sqrt(e.b[0] * e.b[0] + e.b[1] * e.b[1] + e.b[2] * e.b[2]);
                for (int k = 0; k < 3; ++k)
                    if (std::fabs(e.b[(k + 1) % 3] / len) < 2.220446049250313e-16 && std::fabs(e.b[(k + 2) % 3] / len) < 2.220446049250313e-16 && e.a[k] != 0.) return false;
            }
        }
        return true;
    }
private:
    void push(uint32_t kind, Vec3 a, Vec3 b, Vec3 c, double r, MaterialPtr m, std::optional<Transformation> t) {
        Entry e; e.kind = kind; e.a[0] = a.x; e.a[1] = a.y; e.a[2] = a.z; e.b[0] = b.x; e.b[1] = b.y; e.b[2] = b.z; e.c[0] = c.x; e.c[1] = c.y; e.c[2] = c.z;
        e.r = r; e.mat = std::move(m); e.transform = std::move(t);
        entries_.push_back(std::move(e));
    }
    std::vector<Sphere> spheres_;
    std::vector<Plane> planes_;
    std::vector<Entry> entries_;
};

// hittable_collections/bvh.rs:106-143.  The host object only carries the primitives; the device BVH is
// built by rtw_scene_create (the hit result does not depend on the tree).
class BoundedVolumeHierarchy {
public:
    static BoundedVolumeHierarchy from(HittableList list) { BoundedVolumeHierarchy b; b.list_ = std::move(list); return b; }
    size_t len() const { return list_.len(); }
    const HittableList& list() const { return list_; }
private:
    HittableList list_;
};

// What Camera::render receives as `world` / `lights` (&dyn Hittable in the reference, camera.rs:295).
struct World {
    const HittableList* list;
    bool is_bvh;
    World(const HittableList& l) : list(&l), is_bvh(false) {}
    World(const BoundedVolumeHierarchy& b) : list(&b.list()), is_bvh(true) {}
};

// rtw_scene_desc + the arrays it points into, built from two lists (same layout rules as the Python mirror)
struct SceneDescription {
    std::vector<rtw_sphere> spheres; std::vector<rtw_plane> planes; std::vector<rtw_quad> quads; std::vector<rtw_cuboid> cuboids;
    std::vector<rtw_transform> transforms; std::vector<rtw_material> materials; std::vector<rtw_texture> textures;
    std::vector<rtw_perlin> perlins; std::vector<rtw_prim> world, lights;
    rtw_scene_desc pod{};
    SceneDescription(World w, World l) {
        std::vector<const Material*> seen_m; std::vector<std::pair<const Texture*, uint32_t>> seen_t;
        std::function<uint32_t(const TexturePtr&)> texture_ref = [&](const TexturePtr& t) -> uint32_t {      // 1-based index into textures[]
            for (const auto& k : seen_t) if (k.first == t.get()) return k.second;
            rtw_texture tx{};
            tx.kind = t->kind; tx.scale = t->scale;
            if (t->kind == RTW_TEX_NOISE) {
                rtw_perlin pn; rtw_perlin_generate(t->seed, t->index, &pn); perlins.push_back(pn);
                tx.perlin = (uint32_t)perlins.size() - 1;
            } else {
                const TexturePtr sub[2] = {t->even, t->odd};
                for (int k = 0; k < 2; ++k) {
                    if (!sub[k]) throw std::runtime_error("CheckerTexture without even / odd texture");
                    double* col = k == 0 ? tx.even_colour : tx.odd_colour;
                    if (sub[k]->kind == 0) { col[0] = sub[k]->colour.v.x; col[1] = sub[k]->colour.v.y; col[2] = sub[k]->colour.v.z; }
                    else (k == 0 ? tx.even : tx.odd) = texture_ref(sub[k]);
                }
            }
            textures.push_back(tx);
            seen_t.push_back({t.get(), (uint32_t)textures.size()});
            return (uint32_t)textures.size();
        };
        auto material_id = [&](const MaterialPtr& m) -> uint32_t {
            if (!m) throw std::runtime_error("primitive without material");
            for (size_t i = 0; i < seen_m.size(); ++i) if (seen_m[i] == m.get()) return (uint32_t)i;
            rtw_material pod = m->pod;
            pod.texture = m->texture ? texture_ref(m->texture) : 0u;
            materials.push_back(pod); seen_m.push_back(m.get());
            return (uint32_t)materials.size() - 1;
        };
        auto entry = [&](const Entry& e) {
            rtw_prim p{}; p.kind = e.kind; p.transform = -1; p.material = material_id(e.mat);
            if (e.transform) { transforms.push_back(e.transform->pod); p.transform = (int32_t)transforms.size() - 1; }
            switch (e.kind) {
                case RTW_PRIM_SPHERE: spheres.push_back({e.a[0], e.a[1], e.a[2], e.r}); p.index = (uint32_t)spheres.size() - 1; break;
                case RTW_PRIM_PLANE: planes.push_back({e.a[0], e.a[1], e.a[2], e.b[0], e.b[1], e.b[2]}); p.index = (uint32_t)planes.size() - 1; break;
                case RTW_PRIM_CUBOID: { rtw_cuboid c{{e.a[0], e.a[1], e.a[2]}, {e.b[0], e.b[1], e.b[2]}}; cuboids.push_back(c); p.index = (uint32_t)cuboids.size() - 1; break; }
                default: { rtw_quad q{{e.a[0], e.a[1], e.a[2]}, {e.b[0], e.b[1], e.b[2]}, {e.c[0], e.c[1], e.c[2]}}; quads.push_back(q); p.index = (uint32_t)quads.size() - 1; }
            }
            return p;
        };
        for (const Entry& e : w.list->entries()) world.push_back(entry(e));
        for (const Entry& e : l.list->entries()) lights.push_back(entry(e));
        pod.spheres = spheres.data(); pod.n_spheres = spheres.size(); pod.planes = planes.data(); pod.n_planes = planes.size();
        pod.quads = quads.data(); pod.n_quads = quads.size(); pod.cuboids = cuboids.data(); pod.n_cuboids = cuboids.size();
        pod.transforms = transforms.data(); pod.n_transforms = transforms.size(); pod.materials = materials.data(); pod.n_materials = materials.size();
        pod.textures = textures.data(); pod.n_textures = textures.size(); pod.perlins = perlins.data(); pod.n_perlins = perlins.size();
        pod.world = world.data(); pod.n_world = world.size(); pod.lights = lights.data(); pod.n_lights = lights.size();
        pod.world_is_bvh = w.is_bvh ? 1u : 0u; pod.lights_is_bvh = l.is_bvh ? 1u : 0u;
    }
    SceneDescription(const SceneDescription&) = delete;
    SceneDescription& operator=(const SceneDescription&) = delete;
};

enum class Precision : uint32_t { F32 = RTW_F32, F64 = RTW_F64 };
struct RenderOptions {
    uint64_t seed = 20261018;
    double tmin = RTW_TMIN_REFERENCE;               // camera.rs:473: machine epsilon of the working precision
    Precision precision = Precision::F32;
    uint32_t mode = RTW_WAVEFRONT, flags = 0;       // the faster FP32 renderer; RTW_MEGAKERNEL renders the same image
    int n_gpus = 1;                                 // > 1: rtw_render_multi on CUDA devices 0 .. n_gpus-1 of this process (same image bit for bit)
    uint32_t collective = RTW_COLLECTIVE_AUTO;      // how the N partial frames meet: fused peer-memory kernel or NCCL
};

class Camera;
// shared/src/camera.rs:28-219
class CameraBuilder {
public:
    CameraBuilder() {
        b_ = rtw_camera_builder{};
        b_.samples_per_pixel = 10; b_.max_depth = 10; b_.vfov = 90.;
        b_.lookat[2] = -1.; b_.vup[1] = 1.; b_.focus_dist = 10.;
    }
    CameraBuilder with_aspect_ratio(double a) const { auto c = *this; c.b_.aspect_ratio = a; c.b_.has_aspect_ratio = 1; return c; }
    CameraBuilder with_image_width(uint32_t w) const { auto c = *this; c.b_.image_width = w; c.b_.has_image_width = 1; return c; }
    CameraBuilder with_image_height(uint32_t h) const { auto c = *this; c.b_.image_height = h; c.b_.has_image_height = 1; return c; }
    CameraBuilder with_samples_per_pixel(uint16_t s) const { auto c = *this; c.b_.samples_per_pixel = s; return c; }
    CameraBuilder with_max_depth(uint32_t d) const { auto c = *this; c.b_.max_depth = d; return c; }
    CameraBuilder with_background(Colour k) const { auto c = *this; c.b_.background[0] = k.v.x; c.b_.background[1] = k.v.y; c.b_.background[2] = k.v.z; return c; }
    CameraBuilder with_vfov(double v) const { auto c = *this; c.b_.vfov = v; return c; }
    CameraBuilder with_lookfrom(Point3 p) const { auto c = *this; c.b_.lookfrom[0] = p.x; c.b_.lookfrom[1] = p.y; c.b_.lookfrom[2] = p.z; return c; }
    CameraBuilder with_lookat(Point3 p) const { auto c = *this; c.b_.lookat[0] = p.x; c.b_.lookat[1] = p.y; c.b_.lookat[2] = p.z; return c; }
    CameraBuilder with_vup(Vec3 p) const { auto c = *this; c.b_.vup[0] = p.x; c.b_.vup[1] = p.y; c.b_.vup[2] = p.z; return c; }
    CameraBuilder with_defocus_angle(double a) const { auto c = *this; c.b_.defocus_angle = a; return c; }
    CameraBuilder with_focus_dist(double f) const { auto c = *this; c.b_.focus_dist = f; return c; }
    Camera build() const;
    const rtw_camera_builder& pod() const { return b_; }
private:
    rtw_camera_builder b_;
};

// shared/src/camera.rs:227-312
class Camera {
public:
    explicit Camera(const rtw_camera& c) : c_(c) {}
    const rtw_camera& pod() const { return c_; }

    // Camera::render (camera.rs:295-297): out[j][i], j = 0 is the bottom row.
    // world + lights -> rtw_scene (the sphere path when the scene allows it, else the general path)
    static rtw_scene* make_scene(World world, World lights) {
        rtw_scene* scene = nullptr;
        bool simple = world.list->is_simple() && !lights.is_bvh && lights.list->spheres().size() == lights.list->len();
        if (simple && lights.list->is_empty())          // an empty lights list next to a Lambertian: only the general path accepts it
            for (const Entry& e : world.list->entries()) if (e.mat->pod.kind == RTW_LAMBERTIAN) simple = false;
        if (simple) {
            std::vector<rtw_sphere> spheres; std::vector<uint32_t> smat; std::vector<rtw_plane> planes; std::vector<uint32_t> pmat;
            std::vector<rtw_material> mats; std::vector<rtw_sphere> ls;
            auto mat_id = [&](const MaterialPtr& m) {
                if (!m) throw std::runtime_error("primitive without material");
                mats.push_back(m->pod);
                return (uint32_t)mats.size() - 1;
            };
            for (const Plane& p : world.list->planes()) { planes.push_back({p.point.x, p.point.y, p.point.z, p.normal.x, p.normal.y, p.normal.z}); pmat.push_back(mat_id(p.mat)); }
            for (const Sphere& s : world.list->spheres()) { spheres.push_back({s.center.x, s.center.y, s.center.z, s.radius}); smat.push_back(mat_id(s.mat)); }
            for (const Sphere& s : lights.list->spheres()) ls.push_back({s.center.x, s.center.y, s.center.z, s.radius});
            int rc = rtw_scene_create(spheres.data(), smat.data(), spheres.size(), planes.data(), pmat.data(), planes.size(), mats.data(),
                                      mats.size(), ls.data(), ls.size(), &scene);
            if (rc != RTW_OK) throw std::runtime_error(std::string("rtw_scene_create: ") + rtw_last_error());
        } else {
            SceneDescription d(world, lights);
            int rc = rtw_scene_create_general(&d.pod, &scene);
            if (rc != RTW_OK) throw std::runtime_error(std::string("rtw_scene_create_general: ") + rtw_last_error());
        }
        return scene;
    }
    static rtw_opts to_opts(const RenderOptions& opt) {
        rtw_opts o{}; o.seed = opt.seed; o.tmin = opt.tmin; o.precision = (uint32_t)opt.precision; o.mode = opt.mode; o.flags = opt.flags;
        return o;
    }

    std::vector<std::vector<SampledColour>> render(World world, World lights, const RenderOptions& opt = RenderOptions(),
                                                   rtw_stats* stats = nullptr) const {
        rtw_scene* scene = make_scene(world, lights);
        rtw_opts o = to_opts(opt);
        size_t npx = (size_t)c_.image_width * c_.image_height;
        std::vector<double> sum(npx * 3); std::vector<uint8_t> q(npx * 3);
        int rc = opt.n_gpus > 1 ? rtw_render_multi(scene, &c_, &o, opt.n_gpus, nullptr, opt.collective, sum.data(), q.data(), stats)
                                : rtw_render(scene, &c_, &o, sum.data(), q.data(), stats);
        rtw_scene_destroy(scene);
        if (rc != RTW_OK) throw std::runtime_error(std::string(opt.n_gpus > 1 ? "rtw_render_multi: " : "rtw_render: ") + rtw_last_error());
        return rows(sum, q);
    }

    // Progressive rendering with checkpoints: the samples are rendered in `passes` passes into 64-bit fixed-point
    // accumulators; after every pass the accumulators and the next sample index are written to `checkpoint` (if given), and a
    // render started with resume = true continues from that file.  The result equals render() bit for bit (FP32 renderers).
    std::vector<std::vector<SampledColour>> render_progressive(World world, World lights, const RenderOptions& opt, uint32_t passes,
                                                               const std::string& checkpoint = "", bool resume = false,
                                                               rtw_stats* stats = nullptr, uint32_t stop_after = 0) const {
        const uint32_t spp = c_.samples_per_pixel;
        const size_t n = rtw_accum_slots(c_.image_width, c_.image_height);
        std::vector<uint64_t> accum(n * 3, 0); std::vector<uint32_t> poison(n, 0);
        struct Header { char magic[8]; uint32_t width, height, spp, next_sample; uint64_t seed; } hd{};
        uint32_t next = 0;
        if (resume) {
            std::ifstream f(checkpoint, std::ios::binary);
            if (!f || !f.read(reinterpret_cast<char*>(&hd), sizeof(hd)) || std::string(hd.magic, 7) != "RTWCKPT" || hd.width != c_.image_width ||
                hd.height != c_.image_height || hd.spp != spp || hd.seed != opt.seed)
                throw std::runtime_error("checkpoint '" + checkpoint + "' is missing or belongs to another render");
            f.read(reinterpret_cast<char*>(accum.data()), (std::streamsize)(accum.size() * 8));
            f.read(reinterpret_cast<char*>(poison.data()), (std::streamsize)(poison.size() * 4));
            if (!f) throw std::runtime_error("checkpoint '" + checkpoint + "' is truncated");
            next = hd.next_sample;
        }
        rtw_scene* scene = make_scene(world, lights);
        rtw_opts o = to_opts(opt);
        rtw_stats total{};
        passes = passes ? passes : 1;
        const uint32_t per_pass = (spp + passes - 1) / passes;
        for (uint32_t done = 0; next < spp && (stop_after == 0 || done < stop_after); ++done) {     // stop_after: an "interrupted" run
            uint32_t count = std::min(per_pass, spp - next);
            rtw_stats st{};
            int rc = rtw_render_samples(scene, &c_, &o, next, count, accum.data(), poison.data(), &st);
            if (rc != RTW_OK) { rtw_scene_destroy(scene); throw std::runtime_error(std::string("rtw_render_samples: ") + rtw_last_error()); }
            total.paths += st.paths; total.rays += st.rays; total.kernel_ms += st.kernel_ms; total.total_ms += st.total_ms; total.launches += st.launches;
            next += count;
            if (!checkpoint.empty()) {
                std::memcpy(hd.magic, "RTWCKPT", 8); hd.width = c_.image_width; hd.height = c_.image_height; hd.spp = spp; hd.next_sample = next; hd.seed = opt.seed;
                std::ofstream f(checkpoint + ".tmp", std::ios::binary | std::ios::trunc);
                f.write(reinterpret_cast<const char*>(&hd), sizeof(hd));
                f.write(reinterpret_cast<const char*>(accum.data()), (std::streamsize)(accum.size() * 8));
                f.write(reinterpret_cast<const char*>(poison.data()), (std::streamsize)(poison.size() * 4));
                f.close();
                if (!f || std::rename((checkpoint + ".tmp").c_str(), checkpoint.c_str()) != 0) { rtw_scene_destroy(scene); throw std::runtime_error("cannot write checkpoint '" + checkpoint + "'"); }
            }
        }
        rtw_scene_destroy(scene);
        size_t npx = (size_t)c_.image_width * c_.image_height;
        std::vector<double> sum(npx * 3); std::vector<uint8_t> q(npx * 3);
        if (rtw_resolve_accum(accum.data(), poison.data(), c_.image_width, c_.image_height, spp, sum.data(), q.data()) != RTW_OK)
            throw std::runtime_error(std::string("rtw_resolve_accum: ") + rtw_last_error());
        if (stats) *stats = total;
        return rows(sum, q);
    }

private:
    std::vector<std::vector<SampledColour>> rows(const std::vector<double>& sum, const std::vector<uint8_t>& q) const {
        std::vector<std::vector<SampledColour>> out(c_.image_height, std::vector<SampledColour>(c_.image_width));
        for (uint32_t j = 0; j < c_.image_height; ++j)
            for (uint32_t i = 0; i < c_.image_width; ++i) {
                size_t k = ((size_t)j * c_.image_width + i) * 3;
                SampledColour& sc = out[j][i];
                sc.colour = Colour(sum[k], sum[k + 1], sum[k + 2]); sc.samples = (int32_t)c_.samples_per_pixel;
                sc.rgb[0] = q[k]; sc.rgb[1] = q[k + 1]; sc.rgb[2] = q[k + 2];
            }
        return out;
    }
    rtw_camera c_;
};
inline Camera CameraBuilder::build() const {
    rtw_camera c;
    if (rtw_camera_build(&b_, &c) != RTW_OK) throw std::runtime_error(rtw_last_error());
    return Camera(c);
}

// ---- bin/src/config.rs: the [image] table of Config.toml ------------------------------------------------
struct Image { double aspect_ratio; uint32_t image_width, image_height; uint16_t samples_per_pixel; uint8_t max_depth; };
// Config::get_image (config.rs:8-12, 52-99): aspect_ratio / image_width / image_height are optional, two of the three must be
// given; the missing one is derived with a truncating cast.  Throws where the reference unwraps a None / a parse error.
inline Image read_config(const std::string& path) {
    std::ifstream f(path);
    if (!f) throw std::runtime_error("cannot read '" + path + "'");
    std::optional<double> aspect; std::optional<uint32_t> w, h; std::optional<long> spp, depth;
    std::string line, table;
    while (std::getline(f, line)) {
        size_t hash = line.find('#');
        if (hash != std::string::npos) line.erase(hash);
        auto trim = [](std::string v) { size_t a = v.find_first_not_of(" \t\r"), b = v.find_last_not_of(" \t\r"); return a == std::string::npos ? std::string() : v.substr(a, b - a + 1); };
        line = trim(line);
        if (line.empty()) continue;
        if (line.front() == '[') { table = trim(line.substr(1, line.find(']') - 1)); continue; }
        size_t eq = line.find('=');
        if (eq == std::string::npos || table != "image") continue;
        std::string key = trim(line.substr(0, eq)), val = trim(line.substr(eq + 1));
        std::string digits;
        for (char c : val) if (c != '_') digits += c;            // TOML allows 1_000
        if (key == "aspect_ratio") aspect = std::stod(digits);
        else if (key == "image_width") w = (uint32_t)std::stoul(digits);
        else if (key == "image_height") h = (uint32_t)std::stoul(digits);
        else if (key == "samples_per_pixel") spp = std::stol(digits);
        else if (key == "max_depth") depth = std::stol(digits);
    }
    if (!spp || !depth || *spp < 0 || *spp > 65535 || *depth < 0 || *depth > 255)
        throw std::runtime_error("Config.toml: [image] needs samples_per_pixel (u16) and max_depth (u8)");
    Image im{};
    im.samples_per_pixel = (uint16_t)*spp; im.max_depth = (uint8_t)*depth;
    if (!aspect && h && w) { im.aspect_ratio = (double)*w / (double)*h; im.image_height = *h; im.image_width = *w; }
    else if (aspect && !h && w) { im.aspect_ratio = *aspect; im.image_height = (uint32_t)((double)*w / *aspect); im.image_width = *w; }
    else if (aspect && h && !w) { im.aspect_ratio = *aspect; im.image_height = *h; im.image_width = (uint32_t)((double)*h * *aspect); }
    else if (aspect && h && w) { im.aspect_ratio = *aspect; im.image_height = *h; im.image_width = *w; }
    else throw std::runtime_error("Config.toml: [image] needs two of aspect_ratio / image_width / image_height");
    return im;
}

// ---- image writers: the reference's ASCII P3 (bin/src/main.rs:89-104) plus binary P6 and PNG; rows top to bottom ------
inline void write_p3(const std::string& path, const std::vector<std::vector<SampledColour>>& img) {
    FILE* f = std::fopen(path.c_str(), "w");
    if (!f) throw std::runtime_error("cannot write '" + path + "'");
    std::fprintf(f, "P3\n%zu %zu\n255\n", img.empty() ? (size_t)0 : img[0].size(), img.size());
    for (size_t j = img.size(); j-- > 0;)
        for (const SampledColour& c : img[j]) std::fprintf(f, "%s\n", c.to_string().c_str());
    std::fclose(f);
}
inline std::vector<uint8_t> top_down_rgb(const std::vector<std::vector<SampledColour>>& img) {
    std::vector<uint8_t> px;
    for (size_t j = img.size(); j-- > 0;)
        for (const SampledColour& c : img[j]) { px.push_back(c.rgb[0]); px.push_back(c.rgb[1]); px.push_back(c.rgb[2]); }
    return px;
}
inline void write_p6(const std::string& path, const std::vector<std::vector<SampledColour>>& img) {
    std::ofstream f(path, std::ios::binary | std::ios::trunc);
    if (!f) throw std::runtime_error("cannot write '" + path + "'");
    f << "P6\n" << (img.empty() ? 0 : img[0].size()) << " " << img.size() << "\n255\n";
    std::vector<uint8_t> px = top_down_rgb(img);
    f.write(reinterpret_cast<const char*>(px.data()), (std::streamsize)px.size());
}
inline void write_png(const std::string& path, const std::vector<std::vector<SampledColour>>& img) {      // 8-bit RGB, stored (uncompressed) deflate blocks
    const uint32_t w = img.empty() ? 0 : (uint32_t)img[0].size(), h = (uint32_t)img.size();
    std::vector<uint8_t> px = top_down_rgb(img), raw;
    for (uint32_t j = 0; j < h; ++j) { raw.push_back(0); raw.insert(raw.end(), px.begin() + (size_t)j * w * 3, px.begin() + (size_t)(j + 1) * w * 3); }
    auto crc32 = [](const uint8_t* p, size_t n, uint32_t c = 0xffffffffu) { for (size_t i = 0; i < n; ++i) { c ^= p[i]; for (int k = 0; k < 8; ++k) c = (c >> 1) ^ (0xedb88320u & (0u - (c & 1u))); } return c; };
    auto be32 = [](std::vector<uint8_t>& v, uint32_t x) { v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x); };
    std::vector<uint8_t> z = {0x78, 0x01};
    uint32_t a = 1, b = 0;
    for (uint8_t c : raw) { a = (a + c) % 65521u; b = (b + a) % 65521u; }
    for (size_t off = 0; off < raw.size() || off == 0; off += 65535) {
        size_t len = std::min<size_t>(65535, raw.size() - off);
        z.push_back(off + len >= raw.size() ? 1 : 0);
        z.push_back(len & 255); z.push_back(len >> 8); z.push_back(~len & 255); z.push_back((~len >> 8) & 255);
        z.insert(z.end(), raw.begin() + off, raw.begin() + off + len);
        if (raw.empty()) break;
    }
    be32(z, (b << 16) | a);
    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    auto chunk = [&](const char* type, const std::vector<uint8_t>& data) {
        be32(out, (uint32_t)data.size());
        std::vector<uint8_t> td(type, type + 4);
        td.insert(td.end(), data.begin(), data.end());
        out.insert(out.end(), td.begin(), td.end());
        be32(out, crc32(td.data(), td.size()) ^ 0xffffffffu);
    };
    std::vector<uint8_t> ihdr;
    be32(ihdr, w); be32(ihdr, h); ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    chunk("IHDR", ihdr); chunk("IDAT", z); chunk("IEND", {});
    std::ofstream f(path, std::ios::binary | std::ios::trunc);
    if (!f) throw std::runtime_error("cannot write '" + path + "'");
    f.write(reinterpret_cast<const char*>(out.data()), (std::streamsize)out.size());
}

// ---- scenes ------------------------------------------------------------------------------------------
namespace scenes {

// Counter-based host RNG for scene construction: Philox4x32-10, key = seed, counter =
// (0x5CE9E000, 0, 0, block); draw k uses the 64-bit word x[2k] | x[2k+1] << 32.
class SceneRng {
public:
    explicit SceneRng(uint64_t seed) { key_[0] = (uint32_t)seed; key_[1] = (uint32_t)(seed >> 32); }
    uint64_t next_u64() {
        uint32_t block = k_ >> 1;
        if (block != cached_) { uint32_t ctr[4] = {0x5CE9E000u, 0u, 0u, block}; rtw_philox4x32_10(ctr, key_, buf_); cached_ = block; }
        uint32_t o = (k_ & 1) * 2; k_++;
        return (uint64_t)buf_[o] | ((uint64_t)buf_[o + 1] << 32);
    }
    double standard() { return (double)(next_u64() >> 11) * 0x1.0p-53; }                        // rand Standard
    double uniform_inclusive(double low, double high) {                                          // rand Uniform::new_inclusive
        const double max_rand = 1. - 2.220446049250313e-16;
        double scale = (high - low) / max_rand;
        while (scale * max_rand + low > high) scale = std::nextafter(scale, -INFINITY);
        return (double)(next_u64() >> 12) * 0x1.0p-52 * scale + low;
    }
private:
    uint32_t key_[2], buf_[4], k_ = 0, cached_ = 0xffffffffu;
};

struct Output { BoundedVolumeHierarchy world; HittableList lights; CameraBuilder cam; };
// the general scenes return either plain lists or BoundedVolumeHierarchy wrappers, like the reference's generators
struct GeneralOutput {
    HittableList world, lights; bool world_is_bvh = false, lights_is_bvh = false; CameraBuilder cam;
    BoundedVolumeHierarchy world_bvh, lights_bvh;
    World world_ref() const { return world_is_bvh ? World(world_bvh) : World(world); }
    World lights_ref() const { return lights_is_bvh ? World(lights_bvh) : World(lights); }
    void wrap(bool w, bool l) {
        world_is_bvh = w; lights_is_bvh = l;
        if (w) world_bvh = BoundedVolumeHierarchy::from(world);
        if (l) lights_bvh = BoundedVolumeHierarchy::from(lights);
    }
};

// scenes::simple (scenes/src/lib.rs:155-233) with an explicit seed, generalised by the grid half-size n,
// the material thresholds and the ground variant so the BASELINE stress configs share the recipe.
// The reference is (n = 11, p_lambertian = 0.8, p_metal = 0.95, ground = 0):
//   ground 0: one-sided Plane((0,0,0),(0,1,0)) with Lambertian(0.9)   (lib.rs:164-168)
//   ground 1: book-1 ground sphere (0,-1000,0) r = 1000, Lambertian(0.5)      [variant]
//   ground 2: no ground                                                       [variant]
inline Output simple(uint64_t seed, int n = 11, double p_lambertian = 0.8, double p_metal = 0.95, int ground = 0) {
    HittableList lights, world;
    MaterialPtr invisible_material = invisible();
    if (ground == 0) world.add(Plane::new_(Point3(0., 0., 0.), Vec3(0., 1., 0.), Lambertian::new_with_colour(Colour(0.9, 0.9, 0.9))));
    else if (ground == 1) world.add(Sphere::new_(Point3(0., -1000., 0.), 1000., Lambertian::new_with_colour(Colour(0.5, 0.5, 0.5))));
    MaterialPtr material1 = Dialectric::new_(1.5);
    SceneRng rng(seed);
    for (int a = -n; a < n; ++a) {
        for (int b = -n; b < n; ++b) {
            double choose_mat = rng.standard();
            double cx = (double)a + 0.9 * rng.standard();
            double cz = (double)b + 0.9 * rng.standard();
            Point3 center(cx, 0.2, cz);
            if ((center - Point3(4., 0.2, 0.)).length() > 0.9) {
                MaterialPtr mat;
                if (choose_mat < p_lambertian) {
                    Vec3 c1; c1.x = rng.standard(); c1.y = rng.standard(); c1.z = rng.standard();
                    Vec3 c2; c2.x = rng.standard(); c2.y = rng.standard(); c2.z = rng.standard();
                    Vec3 albedo = c1 * c2;
                    mat = Lambertian::new_with_colour(Colour(albedo.x, albedo.y, albedo.z));
                } else if (choose_mat < p_metal) {
                    double r = rng.uniform_inclusive(0.5, 1.), g = rng.uniform_inclusive(0.5, 1.), bl = rng.uniform_inclusive(0.5, 1.);
                    double fuzz = 1. - rng.uniform_inclusive(0.5, 1.);
                    mat = Metal::new_(Colour(r, g, bl), fuzz);
                } else {
                    lights.add(Sphere::new_(center, 0.2, invisible_material));
                    mat = material1;
                }
                world.add(Sphere::new_(center, 0.2, mat));
            }
        }
    }
    MaterialPtr material2 = Lambertian::new_with_colour(Colour(0.4, 0.2, 0.1));
    MaterialPtr material3 = Metal::new_(Colour(0.7, 0.6, 0.5), 0.);
    world.add(Sphere::new_(Point3(0., 1., 0.), 1., material1));
    world.add(Sphere::new_(Point3(-4., 1., 0.), 1., material2));
    world.add(Sphere::new_(Point3(4., 1., 0.), 1., material3));
    lights.add(Sphere::new_(Point3(0., 1., 0.), 1., invisible_material));
    Point3 lookfrom(10., 5., 10.), lookat(0., 0., 0.);
    CameraBuilder cam = CameraBuilder().with_lookfrom(lookfrom).with_lookat(lookat).with_focus_dist((lookfrom - lookat).length())
                            .with_vfov(40.).with_background(Colour(1., 1., 1.));
    return Output{BoundedVolumeHierarchy::from(std::move(world)), std::move(lights), cam};
}

// scenes::simple_light (scenes/src/lib.rs:235-290)
inline GeneralOutput simple_light(uint64_t seed) {
    GeneralOutput o;
    MaterialPtr pertext = Lambertian::new_(NoiseTexture::new_(4., seed));
    MaterialPtr difflight = DiffuseLight::new_with_colour(Colour(4., 4., 4.));
    o.world.add(Plane::new_(Point3(0., 0., 0.), Vec3(0., 1., 0.), pertext));
    o.world.add(Sphere::new_(Point3(0., 2., 0.), 2., pertext));
    o.world.add(Quad::new_(Point3(3., 1., -2.), Vec3(2., 0., 0.), Vec3(0., 2., 0.), difflight));
    o.lights.add(Quad::new_(Point3(3., 1., -2.), Vec3(2., 0., 0.), Vec3(0., 2., 0.), difflight));
    Point3 lookfrom(26., 3., 6.), lookat(0., 2., 0.);
    o.cam = CameraBuilder().with_lookfrom(lookfrom).with_lookat(lookat).with_focus_dist((lookfrom - lookat).length()).with_vfov(40.);
    return o;
}

// scenes::cornell_box (scenes/src/lib.rs:292-380)
inline GeneralOutput cornell_box() {
    GeneralOutput o;
    MaterialPtr red = Lambertian::new_with_colour(Colour(0.65, 0.05, 0.05)), white = Lambertian::new_with_colour(Colour(0.73, 0.73, 0.73)),
                green = Lambertian::new_with_colour(Colour(0.12, 0.45, 0.15)), light = DiffuseLight::new_with_colour(Colour(15., 15., 15.)),
                glass = Dialectric::new_(1.5);
    o.world.add(Quad::new_(Point3(555., 0., 0.), Vec3(0., 555., 0.), Vec3(0., 0., 555.), green));
    o.world.add(Quad::new_(Point3(0., 0., 0.), Vec3(0., 555., 0.), Vec3(0., 0., 555.), red));
    o.world.add(Quad::new_(Point3(0., 0., 0.), Vec3(555., 0., 0.), Vec3(0., 0., 555.), white));
    o.world.add(Quad::new_(Point3(0., 555., 0.), Vec3(555., 0., 0.), Vec3(0., 0., 555.), white));
    o.world.add(Quad::new_(Point3(0., 0., 555.), Vec3(0., 555., 0.), Vec3(555., 0., 0.), white));
    o.world.add(Cuboid::new_(Point3(), Point3(165., 330., 165.), white).transform(Translation3(265., 0., 295.)).transform(rotation(15., Axis::Y)));
    o.world.add(Sphere::new_(Point3(190., 90., 190.), 90., glass));
    o.world.add(Quad::new_(Point3(343., 554., 332.), Vec3(-130., 0., 0.), Vec3(0., 0., -105.), light));
    o.lights.add(Quad::new_(Point3(343., 554., 332.), Vec3(-130., 0., 0.), Vec3(0., 0., -105.), light));
    o.lights.add(Sphere::new_(Point3(190., 90., 190.), 90., glass));
    Point3 lookfrom(277.5, 277.5, -800.), lookat(277.5, 277.5, 0.);
    o.cam = CameraBuilder().with_lookfrom(lookfrom).with_lookat(lookat).with_vfov(40.).with_defocus_angle(0.).with_focus_dist((lookfrom - lookat).length());
    return o;
}

// scenes::checkered_spheres (scenes/src/lib.rs:123-153)
inline GeneralOutput checkered_spheres() {
    GeneralOutput o;
    MaterialPtr checker = Lambertian::new_(CheckerTexture::new_with_colours(Colour(0.2, 0.3, 0.1), Colour(0.9, 0.9, 0.9), 0.01));
    o.world.add(Sphere::new_(Point3(0., -10., 0.), 10., checker));
    o.world.add(Sphere::new_(Point3(0., 10., 0.), 10., checker));
    o.lights.add(Sphere::new_(Point3(0., 0., 0.), 0.1, checker));
    Point3 lookfrom(40., 1., 0.), lookat(0., 0., 0.);
    o.cam = CameraBuilder().with_lookfrom(lookfrom).with_lookat(lookat).with_focus_dist((lookfrom - lookat).length()).with_vfov(40.)
                .with_background(Colour(1., 1., 1.));
    return o;
}

// scenes::plane (scenes/src/lib.rs:91-121): an EMPTY lights list; the one-sided plane is invisible from this camera, so the reference
// never samples the lights (it would panic) — integration-tests plane_test renders it
inline GeneralOutput plane() {
    GeneralOutput o;
    o.world.add(Plane::new_(Point3(0., 0., 0.), Vec3(0., 1., 0.),
                            Lambertian::new_(CheckerTexture::new_with_colours(Colour(0.2, 0.3, 0.1), Colour(0.9, 0.9, 0.9), 0.32))));
    Point3 lookfrom(0., 30., 0.), lookat(0., 0., 0.);
    o.cam = CameraBuilder().with_lookfrom(lookfrom).with_lookat(lookat).with_focus_dist((lookfrom - lookat).length()).with_vfov(40.)
                .with_background(Colour(1., 1., 1.));
    return o;
}

inline void corner_walls(HittableList& world, const MaterialPtr& white) {       // the eight quads of lib.rs:403-448 / 527-572
    const double c[8][9] = {{6, 0, 6, 0, 2, 0, -2, 0, 0}, {6, 0, 6, 0, 2, 0, 0, 0, -2}, {-6, 0, 6, 0, 2, 0, 2, 0, 0}, {-6, 0, 6, 0, 2, 0, 0, 0, -2},
                            {-6, 0, -6, 0, 2, 0, 2, 0, 0}, {-6, 0, -6, 0, 2, 0, 0, 0, 2}, {6, 0, -6, 0, 2, 0, -2, 0, 0}, {6, 0, -6, 0, 2, 0, 0, 0, 2}};
    for (const auto& k : c) world.add(Quad::new_(Point3(k[0], k[1], k[2]), Vec3(k[3], k[4], k[5]), Vec3(k[6], k[7], k[8]), white));
}
inline CameraBuilder debug_camera() {
    return CameraBuilder().with_image_width(3).with_image_height(2).with_samples_per_pixel(10).with_max_depth(5)
        .with_lookfrom(Point3(0., 20., 0.)).with_lookat(Point3(0., 0., 0.)).with_focus_dist(4.);
}

// scenes::debugging_scene (scenes/src/lib.rs:382-505)
inline GeneralOutput debugging_scene(uint64_t seed) {
    GeneralOutput o;
    MaterialPtr pertext = Lambertian::new_(NoiseTexture::new_(4., seed));
    o.world.add(Plane::new_(Point3(0., 0., 0.), Vec3(0., 1., 0.), pertext));
    o.world.add(Sphere::new_(Point3(0., 2., 0.), 2., pertext));
    corner_walls(o.world, Lambertian::new_with_colour(Colour(0.75, 0.75, 0.75)));
    const double g[4][6] = {{5, 1, 5, 0.5, 0, 0.5}, {-5, 1, 5, 1, 0, 0}, {-5, 1, -5, 0, 1, 0}, {5, 1, -5, 0, 0, 1}};
    std::vector<Sphere> glow;
    for (const auto& k : g) glow.push_back(Sphere::new_(Point3(k[0], k[1], k[2]), 1., DiffuseLight::new_with_colour(Colour(k[3], k[4], k[5]))));
    for (const Sphere& s : glow) o.world.add(s);
    for (const Sphere& s : glow) o.lights.add(s);
    o.cam = debug_camera();
    o.wrap(true, true);
    return o;
}

// scenes::simple_transform (scenes/src/lib.rs:507-653)
inline GeneralOutput simple_transform(uint64_t seed) {
    GeneralOutput o;
    o.world.add(Plane::new_(Point3(0., 0., 0.), Vec3(0., 1., 0.), Lambertian::new_(NoiseTexture::new_(4., seed))));
    corner_walls(o.world, Lambertian::new_with_colour(Colour(0.75, 0.75, 0.75)));
    Cuboid original = Cuboid::new_(Point3(), Point3(1., 1., 1.), DiffuseLight::new_with_colour(Colour(1., 0., 0.)));
    Transformed<Cuboid> cubes[3] = {original.transform(Translation3(-0.5, 0., -0.5)), original.transform(Translation3(2., 0., 2.)),
                                    original.transform(Translation3(-3., 0., -3.)).transform(rotation(45., Axis::Y))};
    for (const auto& c : cubes) o.world.add(c);
    for (const auto& c : cubes) o.lights.add(c);
    o.cam = debug_camera();
    o.wrap(true, true);
    return o;
}

}  // namespace scenes
}  // namespace rtw_host
