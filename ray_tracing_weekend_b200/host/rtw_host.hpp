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
//
// Where the reference panics (unwrap / expect), this layer throws std::runtime_error.
#pragma once
#include <cmath>
#include <cstdint>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <tuple>
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

// ---- materials (shared/src/material.rs) -----------------------------------------------------------
struct Material { rtw_material pod{}; };
using MaterialPtr = std::shared_ptr<const Material>;
inline MaterialPtr make_material(uint32_t kind, Colour c, double param) {
    auto m = std::make_shared<Material>();
    m->pod.kind = kind; m->pod.r = c.v.x; m->pod.g = c.v.y; m->pod.b = c.v.z; m->pod.param = param;
    return m;
}
struct Lambertian { static MaterialPtr new_with_colour(Colour c) { return make_material(RTW_LAMBERTIAN, c, 0.); } };   // material.rs:345-355
struct Metal { static MaterialPtr new_(Colour albedo, double fuzz) { return make_material(RTW_METAL, albedo, fuzz); } }; // material.rs:401-405
struct Dialectric { static MaterialPtr new_(double ior) { return make_material(RTW_DIELECTRIC, Colour(1., 1., 1.), ior); } }; // material.rs:443-448
inline MaterialPtr invisible() { static MaterialPtr p = make_material(RTW_INVISIBLE, Colour(0., 0., 0.), 0.); return p; }  // material.rs:319-322

// ---- entities --------------------------------------------------------------------------------------
struct Sphere { Point3 center; double radius; MaterialPtr mat;                       // entities/sphere.rs:25-47
    static Sphere new_(Point3 c, double r, MaterialPtr m) { return Sphere{c, r, std::move(m)}; } };
struct Plane { Point3 point; Vec3 normal; MaterialPtr mat;                           // entities/plane.rs:21-39
    static Plane new_(Point3 p, Vec3 n, MaterialPtr m) { return Plane{p, n, std::move(m)}; } };

// hittable_collections/hittable_list.rs:247-294 — objects keep insertion order per type
class HittableList {
public:
    void add(const Sphere& s) { spheres_.push_back(s); }
    void add(const Plane& p) { planes_.push_back(p); }
    size_t len() const { return spheres_.size() + planes_.size(); }
    bool is_empty() const { return len() == 0; }
    const std::vector<Sphere>& spheres() const { return spheres_; }
    const std::vector<Plane>& planes() const { return planes_; }
private:
    std::vector<Sphere> spheres_;
    std::vector<Plane> planes_;
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
    World(const HittableList& l) : list(&l) {}
    World(const BoundedVolumeHierarchy& b) : list(&b.list()) {}
};

enum class Precision : uint32_t { F32 = RTW_F32, F64 = RTW_F64 };
struct RenderOptions {
    uint64_t seed = 20261018;
    double tmin = RTW_TMIN_REFERENCE;               // camera.rs:473: machine epsilon of the working precision
    Precision precision = Precision::F32;
    uint32_t mode = RTW_WAVEFRONT, flags = 0;       // the faster FP32 renderer; RTW_MEGAKERNEL renders the same image
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
    std::vector<std::vector<SampledColour>> render(World world, World lights, const RenderOptions& opt = RenderOptions(),
                                                   rtw_stats* stats = nullptr) const {
        std::vector<rtw_sphere> spheres; std::vector<uint32_t> smat; std::vector<rtw_plane> planes; std::vector<uint32_t> pmat;
        std::vector<rtw_material> mats; std::vector<rtw_sphere> ls;
        auto mat_id = [&](const MaterialPtr& m) {
            if (!m) throw std::runtime_error("primitive without material");
            mats.push_back(m->pod);
            return (uint32_t)mats.size() - 1;
        };
        for (const Plane& p : world.list->planes()) { planes.push_back({p.point.x, p.point.y, p.point.z, p.normal.x, p.normal.y, p.normal.z}); pmat.push_back(mat_id(p.mat)); }
        for (const Sphere& s : world.list->spheres()) { spheres.push_back({s.center.x, s.center.y, s.center.z, s.radius}); smat.push_back(mat_id(s.mat)); }
        if (!lights.list->planes().empty()) throw std::runtime_error("lights: only spheres are supported by the CUDA backend");
        for (const Sphere& s : lights.list->spheres()) ls.push_back({s.center.x, s.center.y, s.center.z, s.radius});
        rtw_scene* scene = nullptr;
        int rc = rtw_scene_create(spheres.data(), smat.data(), spheres.size(), planes.data(), pmat.data(), planes.size(), mats.data(),
                                  mats.size(), ls.data(), ls.size(), &scene);
        if (rc != RTW_OK) throw std::runtime_error(std::string("rtw_scene_create: ") + rtw_last_error());
        rtw_opts o{}; o.seed = opt.seed; o.tmin = opt.tmin; o.precision = (uint32_t)opt.precision; o.mode = opt.mode; o.flags = opt.flags;
        size_t npx = (size_t)c_.image_width * c_.image_height;
        std::vector<double> sum(npx * 3); std::vector<uint8_t> q(npx * 3);
        rc = rtw_render(scene, &c_, &o, sum.data(), q.data(), stats);
        rtw_scene_destroy(scene);
        if (rc != RTW_OK) throw std::runtime_error(std::string("rtw_render: ") + rtw_last_error());
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
private:
    rtw_camera c_;
};
inline Camera CameraBuilder::build() const {
    rtw_camera c;
    if (rtw_camera_build(&b_, &c) != RTW_OK) throw std::runtime_error(rtw_last_error());
    return Camera(c);
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

}  // namespace scenes
}  // namespace rtw_host
