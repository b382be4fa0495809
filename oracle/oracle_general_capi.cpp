// ORACLE — TEST INFRASTRUCTURE ONLY (see rtw_oracle_general.hpp).  C entry points of the general-scene
// restatement.  orc_gdesc mirrors the layout of rtw_scene_desc (include/rtw.h) so that tests hand the very
// same arrays to the oracle and to the product.
#include "rtw_oracle_general.hpp"

#include <chrono>

using namespace orcg;

extern "C" {

struct orc_gmaterial { uint32_t kind, texture; double r, g, b, param; };
struct orc_gquad { double q[3], u[3], v[3]; };
struct orc_gcuboid { double p[3], q[3]; };
struct orc_gtransform { double rotation[9], translation[3]; };
struct orc_gprim { uint32_t kind, index, material; int32_t transform; };
struct orc_gtexture { uint32_t kind, perlin; double scale; uint32_t even, odd; double even_colour[3], odd_colour[3]; };
struct orc_gperlin { double rand_vec[256][3]; uint8_t perm_x[256], perm_y[256], perm_z[256]; };
struct orc_gdesc {
    const double* spheres; uint64_t n_spheres;          // [n][4]
    const double* planes; uint64_t n_planes;            // [n][6]
    const orc_gquad* quads; uint64_t n_quads;
    const orc_gcuboid* cuboids; uint64_t n_cuboids;
    const orc_gtransform* transforms; uint64_t n_transforms;
    const orc_gmaterial* materials; uint64_t n_materials;
    const orc_gtexture* textures; uint64_t n_textures;
    const orc_gperlin* perlins; uint64_t n_perlins;
    const orc_gprim* world; uint64_t n_world;
    const orc_gprim* lights; uint64_t n_lights;
    uint32_t world_is_bvh, lights_is_bvh;
};
// the structs of oracle_capi.cpp (same translation-unit-independent layouts)
struct orc_camera {
    double center[3], pixel00[3], du[3], dv[3], ddu[3], ddv[3], background[3];
    double defocus_angle;
    uint32_t width, height, spp, max_depth;
};
struct orc_options { uint64_t seed; double tmin; uint32_t rng_mode, math_mode, faithful_bvh; int32_t threads; uint32_t fix_nan, reserved; };
struct orc_counters {
    uint64_t rays, paths, box_tests, box_builds, node_visits, sphere_tests, plane_tests, light_tests,
        lambertian, metal, dielectric, absorbed, missed, depth_out;
};

static V3 gv3(const double* p) { return {p[0], p[1], p[2]}; }
static void gput(double* d, V3 v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; }
static Camera g_camera(const orc_camera* c) {
    Camera k;
    k.center = gv3(c->center); k.pixel00_loc = gv3(c->pixel00); k.pixel_delta_u = gv3(c->du); k.pixel_delta_v = gv3(c->dv);
    k.defocus_disk_u = gv3(c->ddu); k.defocus_disk_v = gv3(c->ddv); k.background = gv3(c->background);
    k.defocus_angle = c->defocus_angle; k.image_width = c->width; k.image_height = c->height;
    k.samples_per_pixel = c->spp; k.max_depth = c->max_depth;
    return k;
}
static Options g_options(const orc_options* o) {
    Options k;
    k.seed = o->seed; k.tmin = o->tmin; k.rng_mode = o->rng_mode; k.math_mode = o->math_mode;
    k.faithful_bvh = false; k.threads = o->threads; k.fix_nan = o->fix_nan != 0;
    return k;
}
static Transformation g_transform(const orc_gtransform& t) {
    Transformation k;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) k.rotation.m[i][j] = t.rotation[3 * i + j];
    k.translation = gv3(t.translation);
    return k;
}
static void g_put_transform(const Transformation& k, orc_gtransform* t) {
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) t->rotation[3 * i + j] = k.rotation.m[i][j];
    gput(t->translation, k.translation);
}

static Prim g_prim(const orc_gdesc* d, const orc_gprim& e, int32_t id) {
    Prim p;
    p.kind = e.kind; p.id = id;
    switch (e.kind) {
        case P_SPHERE: { const double* s = d->spheres + 4 * e.index; p.sphere = Sphere::make(gv3(s), s[3], e.material, id); break; }
        case P_PLANE: { const double* s = d->planes + 6 * e.index; p.plane = Plane::make(gv3(s), gv3(s + 3), e.material, id); break; }
        case P_CUBOID: { const orc_gcuboid& c = d->cuboids[e.index]; p.cuboid = Cuboid::make(gv3(c.p), gv3(c.q), e.material); break; }
        default: { const orc_gquad& q = d->quads[e.index]; p.quad = Quad::make(gv3(q.q), gv3(q.u), gv3(q.v), e.material, e.kind == P_TRIANGLE); break; }
    }
    if (e.transform >= 0) { p.transformed = true; p.tf = g_transform(d->transforms[e.transform]); }
    p.finalize();
    return p;
}

GScene* orc_gscene_create(const orc_gdesc* d) {
    auto* s = new GScene();
    for (uint64_t i = 0; i < d->n_materials; ++i) {
        const orc_gmaterial& m = d->materials[i];
        s->materials.push_back(GMaterial{m.kind, m.texture, {m.r, m.g, m.b}, m.param});
    }
    for (uint64_t i = 0; i < d->n_textures; ++i) {
        const orc_gtexture& t = d->textures[i];
        s->textures.push_back(Texture{t.kind, t.perlin, t.scale, t.even, t.odd, gv3(t.even_colour), gv3(t.odd_colour)});
    }
    for (uint64_t i = 0; i < d->n_perlins; ++i) {
        Perlin p;
        std::memcpy(p.rand_vec, d->perlins[i].rand_vec, sizeof(p.rand_vec));
        std::memcpy(p.perm_x, d->perlins[i].perm_x, 256); std::memcpy(p.perm_y, d->perlins[i].perm_y, 256); std::memcpy(p.perm_z, d->perlins[i].perm_z, 256);
        s->perlins.push_back(p);
    }
    for (uint64_t i = 0; i < d->n_world; ++i) s->world_list.add(g_prim(d, d->world[i], (int32_t)i));
    for (uint64_t i = 0; i < d->n_lights; ++i) s->lights.add(g_prim(d, d->lights[i], (int32_t)i));
    s->world_is_bvh = d->world_is_bvh != 0; s->lights_is_bvh = d->lights_is_bvh != 0;
    if (s->world_is_bvh) s->world_bvh = GBvh::from(s->world_list);
    return s;
}
void orc_gscene_destroy(GScene* s) { delete s; }

void orc_gtrace_batch(const GScene* s, uint64_t n, const double* o, const double* d, double tmin, double tmax, int32_t* prim, double* t,
                      double* p, double* normal) {
    bool pan = false;
    for (uint64_t i = 0; i < n; ++i) {
        Ray r{gv3(o + 3 * i), gv3(d + 3 * i)};
        GHit rec;
        bool h = s->world_is_bvh ? s->world_bvh->hit(r, tmin, tmax, &rec, nullptr, &pan) : s->world_list.hit(r, tmin, tmax, &rec, nullptr, &pan);
        prim[i] = h ? rec.prim : -1;
        t[i] = h ? rec.t : INF;
        if (p) gput(p + 3 * i, h ? rec.p : V3{0, 0, 0});
        if (normal) gput(normal + 3 * i, h ? rec.normal : V3{0, 0, 0});
    }
}

void orc_gscatter_batch(const GScene* s, const orc_options* opt, uint64_t n, const double* o, const double* d,
                        const uint32_t* pixel, const uint32_t* sample, const uint32_t* vertex,
                        int32_t* prim, double* t, uint32_t* kind, double* p, double* normal, double* dir, double* weight, double* emitted) {
    Options op = g_options(opt);
    bool pan = false;
    for (uint64_t i = 0; i < n; ++i) {
        Ray r{gv3(o + 3 * i), gv3(d + 3 * i)};
        GHit rec;
        GVertex vx;
        V3 z{0, 0, 0};
        if (!s->world_hit(r, op.tmin, &rec, nullptr, &pan)) {
            prim[i] = -1; t[i] = INF; kind[i] = V_MISS;
            gput(p + 3 * i, z); gput(normal + 3 * i, z); gput(dir + 3 * i, z); gput(weight + 3 * i, z);
            if (emitted) gput(emitted + 3 * i, z);
            continue;
        }
        Stream rng(op.seed, pixel[i], sample[i], vertex[i], op.rng_mode);
        gshade(*s, op, r, rec, rng, &vx, nullptr);
        prim[i] = rec.prim; t[i] = rec.t; kind[i] = vx.kind;
        gput(p + 3 * i, rec.p); gput(normal + 3 * i, rec.normal);
        gput(dir + 3 * i, vx.kind >= V_SPECULAR ? vx.next.d : z);
        gput(weight + 3 * i, vx.weight);
        if (emitted) gput(emitted + 3 * i, vx.emitted);
    }
}

void orc_gpath_radiance(const GScene* s, const orc_camera* cam, const orc_options* opt, uint64_t n, const uint32_t* i,
                        const uint32_t* j, const uint32_t* sample, double* rgb) {
    Camera c = g_camera(cam);
    Options op = g_options(opt);
    bool pan = false;
    for (uint64_t k = 0; k < n; ++k) {
        uint32_t pixel = j[k] * c.image_width + i[k];
        Stream rng(op.seed, pixel, sample[k], 0, op.rng_mode);
        Ray r = get_ray(c, i[k], j[k], rng);
        V3 v = gray_colour(*s, c, op, r, pixel, sample[k], nullptr, &pan);
        if (op.fix_nan) { if (v.x != v.x) v.x = 0.; if (v.y != v.y) v.y = 0.; if (v.z != v.z) v.z = 0.; }
        gput(rgb + 3 * k, v);
    }
}

double orc_grender(const GScene* s, const orc_camera* cam, const orc_options* opt, double* rgb_sum, orc_counters* counters, uint32_t* panicked) {
    Camera c = g_camera(cam);
    Options op = g_options(opt);
    Counters cnt;
    bool pan = false;
    auto t0 = std::chrono::steady_clock::now();
    grender(*s, c, op, rgb_sum, &cnt, &pan);
    auto t1 = std::chrono::steady_clock::now();
    if (counters)
        *counters = {cnt.rays, cnt.paths, cnt.box_tests, cnt.box_builds, cnt.node_visits, cnt.sphere_tests, cnt.plane_tests, cnt.light_tests,
                     cnt.lambertian, cnt.metal, cnt.dielectric, cnt.absorbed, cnt.missed, cnt.depth_out};
    if (panicked) *panicked = pan ? 1u : 0u;
    return std::chrono::duration<double>(t1 - t0).count();
}

// ---- small pieces exposed for known-answer tests -----------------------------------------------------
void orc_perlin_generate(uint64_t seed, uint32_t index, orc_gperlin* out) {
    Perlin p = Perlin::generate(seed, index);
    std::memcpy(out->rand_vec, p.rand_vec, sizeof(p.rand_vec));
    std::memcpy(out->perm_x, p.perm_x, 256); std::memcpy(out->perm_y, p.perm_y, 256); std::memcpy(out->perm_z, p.perm_z, 256);
}
double orc_perlin_turb(const orc_gperlin* t, const double* p, int32_t depth) {
    Perlin pn;
    std::memcpy(pn.rand_vec, t->rand_vec, sizeof(pn.rand_vec));
    std::memcpy(pn.perm_x, t->perm_x, 256); std::memcpy(pn.perm_y, t->perm_y, 256); std::memcpy(pn.perm_z, t->perm_z, 256);
    return depth <= 0 ? pn.noise(gv3(p)) : pn.turb(gv3(p), depth);
}
double orc_sin_portable(double x) { return sin_portable(x); }
double orc_atan2_msun(double y, double x) { return atan2_msun(y, x); }
double orc_acos_msun(double x) { return acos_msun(x); }
// Transformation::{apply, inverse} and rotation(): out = a.then(b); returns 0 if not invertible
void orc_transform_then(const orc_gtransform* a, const orc_gtransform* b, orc_gtransform* out) { g_put_transform(g_transform(*a).then(g_transform(*b)), out); }
int32_t orc_transform_inverse(const orc_gtransform* a, orc_gtransform* out) {
    Transformation inv;
    if (!g_transform(*a).inverse(&inv)) return 0;
    g_put_transform(inv, out);
    return 1;
}
void orc_rotation(double angle_deg, int32_t axis, orc_gtransform* out) { g_put_transform(rotation(angle_deg, axis, LIBM), out); }
// world-space box of a world / lights entry (Bounded::get_aabbox)
void orc_gprim_box(const GScene* s, uint32_t lights, uint32_t index, double* out6) {
    const GList& l = lights ? s->lights : s->world_list;
    l.for_each([&](const Prim& p) { if ((uint32_t)p.id == index) { gput(out6, p.box.mn); gput(out6 + 3, p.box.mx); } });
}

}  // extern "C"
