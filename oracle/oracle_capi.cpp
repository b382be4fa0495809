// ORACLE — TEST INFRASTRUCTURE ONLY (see rtw_oracle.hpp).  Plain C entry points so tests/ and
// bench.py's cpu_baseline leg can drive the restatement through ctypes.
#include "rtw_oracle.hpp"

#include <chrono>

using namespace orc;

extern "C" {

struct orc_material { uint32_t kind; double r, g, b, param; };
struct orc_camera {
    double center[3], pixel00[3], du[3], dv[3], ddu[3], ddv[3], background[3];
    double defocus_angle;
    uint32_t width, height, spp, max_depth;
};
struct orc_camera_builder {
    double aspect_ratio; uint32_t has_aspect;
    uint32_t width, has_width, height, has_height;
    uint32_t spp, max_depth;
    double background[3], vfov, lookfrom[3], lookat[3], vup[3], defocus_angle, focus_dist;
};
struct orc_options { uint64_t seed; double tmin; uint32_t rng_mode, math_mode, faithful_bvh; int32_t threads; uint32_t fix_nan, reserved; };
struct orc_counters {
    uint64_t rays, paths, box_tests, box_builds, node_visits, sphere_tests, plane_tests, light_tests,
        lambertian, metal, dielectric, absorbed, missed, depth_out;
};
struct orc_bvh_stats { uint64_t nodes, leaves, depth, max_leaf, prims; };

static V3 v3(const double* p) { return {p[0], p[1], p[2]}; }
static void put(double* d, V3 v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; }

static Camera to_camera(const orc_camera* c) {
    Camera k;
    k.center = v3(c->center); k.pixel00_loc = v3(c->pixel00); k.pixel_delta_u = v3(c->du); k.pixel_delta_v = v3(c->dv);
    k.defocus_disk_u = v3(c->ddu); k.defocus_disk_v = v3(c->ddv); k.background = v3(c->background);
    k.defocus_angle = c->defocus_angle; k.image_width = c->width; k.image_height = c->height;
    k.samples_per_pixel = c->spp; k.max_depth = c->max_depth;
    return k;
}
static Options to_options(const orc_options* o) {
    Options k;
    k.seed = o->seed; k.tmin = o->tmin; k.rng_mode = o->rng_mode; k.math_mode = o->math_mode;
    k.faithful_bvh = o->faithful_bvh != 0; k.threads = o->threads; k.fix_nan = o->fix_nan != 0;
    return k;
}
static void from_counters(const Counters& c, orc_counters* o) {
    if (!o) return;
    *o = {c.rays, c.paths, c.box_tests, c.box_builds, c.node_visits, c.sphere_tests, c.plane_tests, c.light_tests,
          c.lambertian, c.metal, c.dielectric, c.absorbed, c.missed, c.depth_out};
}

void orc_philox4x32_10(const uint32_t* ctr, const uint32_t* key, uint32_t* out) { philox4x32_10(ctr, key, out); }

// k-th uniform of stream (seed; pixel, sample, vertex): kind 0 Standard, 1 Open01, 2 U[-0.5,0.5] jitter
double orc_uniform(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t vertex, uint32_t k, uint32_t rng_mode, uint32_t kind) {
    Stream s(seed, pixel, sample, vertex, rng_mode);
    s.k = k;
    return kind == 0 ? s.standard() : kind == 1 ? s.open01() : s.uniform_inclusive(-0.5, 0.5);
}
uint32_t orc_uniform_index(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t vertex, uint32_t k, uint32_t rng_mode, uint32_t n) {
    Stream s(seed, pixel, sample, vertex, rng_mode);
    s.k = k;
    return s.index(n);
}
void orc_sincos(double phi, uint32_t math_mode, double* s, double* c) { sincos_phi(phi, math_mode, s, c); }

void orc_camera_build(const orc_camera_builder* b, orc_camera* out) {
    CameraBuilder k;
    if (b->has_aspect) k.aspect_ratio = b->aspect_ratio;
    if (b->has_width) k.image_width = b->width;
    if (b->has_height) k.image_height = b->height;
    k.samples_per_pixel = b->spp; k.max_depth = b->max_depth; k.background = v3(b->background); k.vfov = b->vfov;
    k.lookfrom = v3(b->lookfrom); k.lookat = v3(b->lookat); k.vup = v3(b->vup);
    k.defocus_angle = b->defocus_angle; k.focus_dist = b->focus_dist;
    Camera c = camera_build(k);
    put(out->center, c.center); put(out->pixel00, c.pixel00_loc); put(out->du, c.pixel_delta_u); put(out->dv, c.pixel_delta_v);
    put(out->ddu, c.defocus_disk_u); put(out->ddv, c.defocus_disk_v); put(out->background, c.background);
    out->defocus_angle = c.defocus_angle; out->width = c.image_width; out->height = c.image_height;
    out->spp = c.samples_per_pixel; out->max_depth = c.max_depth;
}

// ---- scene description generator (scenes::simple restatement) --------------------------------
SceneDesc* orc_desc_simple(uint64_t seed, int32_t n, double p_lambertian, double p_metal, int32_t ground) {
    return new SceneDesc(scene_simple(seed, n, p_lambertian, p_metal, ground));
}
void orc_desc_destroy(SceneDesc* d) { delete d; }
void orc_desc_counts(const SceneDesc* d, uint64_t* out /*spheres, materials, planes, lights*/) {
    out[0] = d->sphere_mat.size(); out[1] = d->materials.size(); out[2] = d->plane_mat.size(); out[3] = d->lights.size() / 4;
}
void orc_desc_copy(const SceneDesc* d, double* spheres, uint32_t* sphere_mat, orc_material* mats, double* planes,
                   uint32_t* plane_mat, double* lights, orc_camera_builder* cam) {
    std::copy(d->spheres.begin(), d->spheres.end(), spheres);
    std::copy(d->sphere_mat.begin(), d->sphere_mat.end(), sphere_mat);
    for (size_t i = 0; i < d->materials.size(); ++i)
        mats[i] = {d->materials[i].kind, d->materials[i].albedo.x, d->materials[i].albedo.y, d->materials[i].albedo.z, d->materials[i].param};
    std::copy(d->planes.begin(), d->planes.end(), planes);
    std::copy(d->plane_mat.begin(), d->plane_mat.end(), plane_mat);
    std::copy(d->lights.begin(), d->lights.end(), lights);
    if (cam) {
        std::memset(cam, 0, sizeof(*cam));
        cam->spp = d->cam.samples_per_pixel; cam->max_depth = d->cam.max_depth;
        put(cam->background, d->cam.background); cam->vfov = d->cam.vfov;
        put(cam->lookfrom, d->cam.lookfrom); put(cam->lookat, d->cam.lookat); put(cam->vup, d->cam.vup);
        cam->defocus_angle = d->cam.defocus_angle; cam->focus_dist = d->cam.focus_dist;
    }
}

// ---- scene -------------------------------------------------------------------------------------
Scene* orc_scene_create(uint64_t ns, const double* spheres, const uint32_t* sphere_mat, uint64_t nm, const orc_material* mats,
                        uint64_t np, const double* planes, const uint32_t* plane_mat, uint64_t nl, const double* lights) {
    std::vector<Material> m(nm);
    for (uint64_t i = 0; i < nm; ++i) m[i] = {mats[i].kind, {mats[i].r, mats[i].g, mats[i].b}, mats[i].param};
    return scene_from_arrays(ns, spheres, sphere_mat, nm, m.data(), np, planes, plane_mat, nl, lights).release();
}
void orc_scene_destroy(Scene* s) { delete s; }
void orc_scene_bvh_stats(const Scene* s, orc_bvh_stats* o) {
    *o = {s->world->node_count(), s->world->leaf_count(), s->world->depth(), s->world->max_leaf(), s->world->len};
}

// world.hit for a batch of rays (parity check 1).  prim: planes first then spheres; -1 = miss.
void orc_trace_batch(const Scene* s, uint64_t n, const double* o, const double* d, double tmin, double tmax,
                     int32_t* prim, double* t, uint32_t faithful, orc_counters* counters) {
    Counters c;
    bool pan = false;
    for (uint64_t i = 0; i < n; ++i) {
        Ray r{v3(o + 3 * i), v3(d + 3 * i)};
        HitRecord rec;
        c.rays++;
        bool h = s->world->hit(r, tmin, tmax, &rec, &c, faithful != 0, &pan);
        prim[i] = h ? rec.prim : -1;
        t[i] = h ? rec.t : INF;
    }
    from_counters(c, counters);
}

// world.hit + Material::scatter (+ mixture pdf) for a batch of rays with explicit stream keys
// (parity check 3).  kind: VertexKind.  Outputs are written only as far as they are defined.
void orc_scatter_batch(const Scene* s, const orc_options* opt, uint64_t n, const double* o, const double* d,
                       const uint32_t* pixel, const uint32_t* sample, const uint32_t* vertex,
                       int32_t* prim, double* t, uint32_t* kind, double* p, double* normal, double* dir, double* weight) {
    Options op = to_options(opt);
    bool pan = false;
    for (uint64_t i = 0; i < n; ++i) {
        Ray r{v3(o + 3 * i), v3(d + 3 * i)};
        HitRecord rec;
        Vertex vx;
        if (!world_hit(*s, r, op.tmin, &rec, nullptr, false, &pan)) {
            prim[i] = -1; t[i] = INF; kind[i] = V_MISS;
            put(p + 3 * i, {0, 0, 0}); put(normal + 3 * i, {0, 0, 0}); put(dir + 3 * i, {0, 0, 0}); put(weight + 3 * i, {0, 0, 0});
            continue;
        }
        Stream rng(op.seed, pixel[i], sample[i], vertex[i], op.rng_mode);
        shade(*s, op, r, rec, rng, &vx, nullptr);
        prim[i] = rec.prim; t[i] = rec.t; kind[i] = vx.kind;
        put(p + 3 * i, rec.p); put(normal + 3 * i, rec.normal);
        put(dir + 3 * i, vx.kind >= V_SPECULAR ? vx.next.d : V3{0, 0, 0});
        put(weight + 3 * i, vx.weight);
    }
}

// Material::scatter + mixture pdf on caller-supplied hit records (parity check 3 on identical inputs): the counterpart of
// rtw_shade_batch.  mat_kind / material describe the material that was hit; the scene supplies the lights list.
void orc_shade_batch(const Scene* s, const orc_options* opt, uint64_t n, const double* d, const double* p, const double* normal,
                     const uint32_t* front_face, const uint32_t* mat_kind, const double* material, const uint32_t* pixel,
                     const uint32_t* sample, const uint32_t* vertex, uint32_t* kind, double* dir, double* weight) {
    Options op = to_options(opt);
    Scene local;                              // shade() looks the material up by index: one scratch material per record
    local.lights = s->lights;
    local.materials.resize(1);
    for (uint64_t i = 0; i < n; ++i) {
        Material m;
        m.kind = mat_kind[i]; m.albedo = v3(material + 4 * i); m.param = material[4 * i + 3];
        local.materials[0] = m;
        HitRecord rec;
        rec.p = v3(p + 3 * i); rec.normal = v3(normal + 3 * i); rec.t = 0.; rec.front_face = front_face[i] != 0; rec.prim = 0; rec.mat = 0;
        Ray r{rec.p, v3(d + 3 * i)};
        Stream rng(op.seed, pixel[i], sample[i], vertex[i], op.rng_mode);
        Vertex vx;
        shade(local, op, r, rec, rng, &vx, nullptr);
        kind[i] = vx.kind;
        put(dir + 3 * i, vx.kind >= V_SPECULAR ? vx.next.d : V3{0, 0, 0});
        put(weight + 3 * i, vx.weight);
    }
}

// primary rays for a list of (i, j, sample) (get_ray, camera.rs:274-293)
void orc_get_rays(const orc_camera* cam, const orc_options* opt, uint64_t n, const uint32_t* i, const uint32_t* j,
                  const uint32_t* sample, double* o, double* d) {
    Camera c = to_camera(cam);
    for (uint64_t k = 0; k < n; ++k) {
        Stream rng(opt->seed, j[k] * c.image_width + i[k], sample[k], 0, opt->rng_mode);
        Ray r = get_ray(c, i[k], j[k], rng);
        put(o + 3 * k, r.o); put(d + 3 * k, r.d);
    }
}

// Camera::render restricted to rows [row_begin, row_end).  rgb_sum: [h][w][3] f64, j = 0 bottom row.
// Returns wall seconds; *panicked is set if the reference would have panicked (plane uv check).
double orc_render(const Scene* s, const orc_camera* cam, const orc_options* opt, double* rgb_sum, uint32_t row_begin,
                  uint32_t row_end, orc_counters* counters, uint32_t* panicked) {
    Camera c = to_camera(cam);
    Options op = to_options(opt);
    Counters cnt;
    bool pan = false;
    auto t0 = std::chrono::steady_clock::now();
    render(*s, c, op, rgb_sum, &cnt, &pan, row_begin, row_end);
    auto t1 = std::chrono::steady_clock::now();
    from_counters(cnt, counters);
    if (panicked) *panicked = pan ? 1u : 0u;
    return std::chrono::duration<double>(t1 - t0).count();
}

// radiance of individual (pixel i, j, sample) paths — for path-level bit-exact checks
void orc_path_radiance(const Scene* s, const orc_camera* cam, const orc_options* opt, uint64_t n, const uint32_t* i,
                       const uint32_t* j, const uint32_t* sample, double* rgb) {
    Camera c = to_camera(cam);
    Options op = to_options(opt);
    bool pan = false;
    for (uint64_t k = 0; k < n; ++k) {
        uint32_t pixel = j[k] * c.image_width + i[k];
        Stream rng(op.seed, pixel, sample[k], 0, op.rng_mode);
        Ray r = get_ray(c, i[k], j[k], rng);
        V3 v = ray_colour(*s, c, op, r, pixel, sample[k], nullptr, &pan);
        if (op.fix_nan) { if (v.x != v.x) v.x = 0.; if (v.y != v.y) v.y = 0.; if (v.z != v.z) v.z = 0.; }
        put(rgb + 3 * k, v);
    }
}

// colour.rs:15-36 for a whole buffer
void orc_resolve(const double* rgb_sum, uint64_t n_values, int32_t spp, uint8_t* out) {
    for (uint64_t i = 0; i < n_values; ++i) out[i] = quantise(rgb_sum[i], spp);
}

uint32_t orc_hardware_threads() { return std::max(1u, std::thread::hardware_concurrency()); }

}  // extern "C"
