// capi.cu — the extern "C" surface declared in include/rtw.h: scene upload, Camera::render and the
// per-ray batch operations, on the current CUDA device.  No CPU fallback: every compute entry point
// fails with RTW_E_NO_DEVICE / RTW_E_CUDA when the GPU path is unavailable.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/rtw.h"
#include "bvh_build.hpp"
#include "rtw_launch.hpp"

using namespace rtw;

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CU(expr)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (expr);                                                                          \
        if (e_ != cudaSuccess) {                                                                          \
            int code_ = (e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver) ? RTW_E_NO_DEVICE  \
                        : (e_ == cudaErrorMemoryAllocation ? RTW_E_NOMEM : RTW_E_CUDA);                   \
            cudaGetLastError(); /* clear the non-sticky error so it cannot surface in a later call */      \
            return fail(code_, std::string(#expr) + ": " + cudaGetErrorString(e_));                       \
        }                                                                                                 \
    } while (0)

constexpr size_t kLightBvhThreshold = 64;   // more lights than this: BVH over the lights (FP32 path)

// Device allocations are recycled through a small process-wide cache: cudaMalloc / cudaFree of the tens of MB
// of per-render buffers (and cudaFree's implicit device synchronisation) otherwise dominate the end-to-end
// time of a scene_create -> render -> scene_destroy cycle.  rtw_release_cached_memory() empties it.
struct DevCache {
    std::mutex m;
    std::multimap<std::pair<int, size_t>, void*> blocks;
    size_t bytes = 0;
    static constexpr size_t kMaxBytes = size_t(4) << 30;
};
DevCache& dev_cache() { static DevCache c; return c; }
size_t round_alloc(size_t n) { return n <= (1u << 20) ? (n + 511) / 512 * 512 : (n + (1u << 20) - 1) / (1u << 20) * (1u << 20); }
cudaError_t cached_malloc(void** p, size_t n) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    n = round_alloc(n);
    {
        DevCache& c = dev_cache();
        std::lock_guard<std::mutex> g(c.m);
        auto it = c.blocks.find({dev, n});
        if (it != c.blocks.end()) { *p = it->second; c.blocks.erase(it); c.bytes -= n; return cudaSuccess; }
    }
    return cudaMalloc(p, n);
}
void cached_free(void* p, size_t n) {
    if (!p) return;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaFree(p); return; }
    n = round_alloc(n);
    DevCache& c = dev_cache();
    std::lock_guard<std::mutex> g(c.m);
    if (c.bytes + n > DevCache::kMaxBytes) { cudaFree(p); return; }
    c.blocks.insert({{dev, n}, p});
    c.bytes += n;
}

template <class P> struct DevBuf {
    P* p = nullptr; size_t n = 0;
    cudaError_t upload(const std::vector<P>& h) {
        release();
        n = h.size();
        if (!n) return cudaSuccess;
        cudaError_t e = cached_malloc(reinterpret_cast<void**>(&p), n * sizeof(P));
        if (e != cudaSuccess) { p = nullptr; n = 0; return e; }
        return cudaMemcpy(p, h.data(), n * sizeof(P), cudaMemcpyHostToDevice);
    }
    cudaError_t reserve(size_t count) {
        if (count <= n && p) return cudaSuccess;
        release();
        cudaError_t e = cached_malloc(reinterpret_cast<void**>(&p), count * sizeof(P));
        if (e != cudaSuccess) { p = nullptr; return e; }
        n = count;
        return cudaSuccess;
    }
    void release() { if (p) cached_free(p, n * sizeof(P)); p = nullptr; n = 0; }
    size_t bytes() const { return n * sizeof(P); }
};

float round_up(double v) { float f = (float)v; if ((double)f < v) f = std::nextafterf(f, INFINITY); return f; }

template <class T> void fill_node(Node<T>& n, const host::FlatNode& f);
template <> void fill_node<double>(Node<double>& n, const host::FlatNode& f) {
    for (int a = 0; a < 3; ++a) { n.la[a] = f.lbox.mn[a]; n.lb[a] = f.lbox.mx[a]; n.ra[a] = f.rbox.mn[a]; n.rb[a] = f.rbox.mx[a]; }
}
template <> void fill_node<float>(Node<float>& n, const host::FlatNode& f) {
    // (centre, half-extent) in FP32, conservative: the half-extent is rounded up from the f64 box around the
    // ROUNDED centre and padded by ~8 ulp of the box scale so FP32 slab arithmetic cannot cull a sphere its own
    // FP32 test would accept
    auto conv = [](const host::Box& b, float* c, float* h) {
        for (int a = 0; a < 3; ++a) {
            float cc = (float)(0.5 * (b.mn[a] + b.mx[a]));
            if (!std::isfinite(cc)) cc = 0.f;
            double e = std::fmax(b.mx[a] - (double)cc, (double)cc - b.mn[a]);
            float hh = round_up(e);
            hh += 1e-6f * std::fmax(1.f, std::fabs(cc) + hh);
            c[a] = cc; h[a] = hh;
        }
    };
    conv(f.lbox, n.la, n.lb);
    conv(f.rbox, n.ra, n.rb);
}

template <class T> struct SceneDev {
    DevBuf<Node<T>> nodes, light_nodes; DevBuf<Vec4T<T>> spheres, sphere_mat, lights; DevBuf<uint32_t> info; DevBuf<PlaneT<T>> planes;
    DevBuf<T> tiles;
    SceneView<T> view{};
    size_t bytes() const { return light_nodes.bytes() + nodes.bytes() + spheres.bytes() + sphere_mat.bytes() + lights.bytes() + info.bytes() + planes.bytes(); }
    void release() { light_nodes.release(); nodes.release(); spheres.release(); sphere_mat.release(); lights.release(); info.release(); planes.release(); tiles.release(); }
};

}  // namespace

struct rtw_scene {
    std::vector<rtw_sphere> spheres; std::vector<uint32_t> sphere_material;
    std::vector<rtw_plane> planes; std::vector<uint32_t> plane_material;
    std::vector<rtw_material> materials; std::vector<rtw_sphere> lights;
    host::Bvh bvh;
    SceneDev<float> f32; SceneDev<double> f64;
    int device = 0, sm_count = 0;
    unsigned int* d_work = nullptr; DeviceCounters* d_counters = nullptr;
    DevBuf<double> d_rgb_sum; DevBuf<uint8_t> d_rgb8;
    DevBuf<unsigned long long> d_accum; DevBuf<uint32_t> d_poison;   // pooled megakernel accumulators
    DevBuf<double> d_in0, d_in1, d_out0, d_out1, d_out2, d_out3, d_out4; DevBuf<uint32_t> d_u0, d_u1, d_u2, d_k; DevBuf<int32_t> d_prim;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    LaunchInfo last_launch;
    uint32_t last_launches = 1;
    uint32_t light_bvh_depth = 0;
};

namespace {

template <class T> int upload_scene(rtw_scene* s, SceneDev<T>& d) {
    const host::Bvh& b = s->bvh;
    std::vector<Node<T>> nodes(b.nodes.size());
    for (size_t i = 0; i < b.nodes.size(); ++i) {
        const host::FlatNode& f = b.nodes[i];
        Node<T> n{};
        fill_node<T>(n, f);
        n.left = f.left >= 0 ? f.left : encode_leaf(f.lfirst, f.lcount);
        n.right = f.right >= 0 ? f.right : encode_leaf(f.rfirst, f.rcount);
        nodes[i] = n;
    }
    size_t ns = s->spheres.size(), np = s->planes.size();
    std::vector<Vec4T<T>> sph(ns), mat(ns), lights(s->lights.size());
    std::vector<uint32_t> info(ns);
    for (size_t k = 0; k < ns; ++k) {
        uint32_t src = b.order[k];
        const rtw_sphere& q = s->spheres[src];
        const rtw_material& m = s->materials[s->sphere_material[src]];
        sph[k] = Vec4T<T>{(T)q.cx, (T)q.cy, (T)q.cz, (T)q.r};
        mat[k] = Vec4T<T>{(T)m.r, (T)m.g, (T)m.b, (T)m.param};
        info[k] = ((uint32_t)(np + src) << 2) | (m.kind & 3u);
    }
    // lights: insertion order; on the FP32 path a long list gets its own BVH (leaf order) so that
    // lights.pdf_value is not O(#lights) per diffuse bounce
    std::vector<Node<T>> light_nodes;
    const bool light_bvh = sizeof(T) == 4 && s->lights.size() > kLightBvhThreshold;
    if (light_bvh) {
        host::Builder lb;
        host::Bvh lbvh = lb.build(reinterpret_cast<const double*>(s->lights.data()), s->lights.size(), 4, kMaxTreeDepth);
        light_nodes.resize(lbvh.nodes.size());
        for (size_t i = 0; i < lbvh.nodes.size(); ++i) {
            const host::FlatNode& f = lbvh.nodes[i];
            Node<T> n{};
            fill_node<T>(n, f);
            n.left = f.left >= 0 ? f.left : encode_leaf(f.lfirst, f.lcount);
            n.right = f.right >= 0 ? f.right : encode_leaf(f.rfirst, f.rcount);
            light_nodes[i] = n;
        }
        for (size_t k = 0; k < s->lights.size(); ++k) {
            const rtw_sphere& q = s->lights[lbvh.order[k]];
            lights[k] = Vec4T<T>{(T)q.cx, (T)q.cy, (T)q.cz, (T)q.r};
        }
        s->light_bvh_depth = lbvh.depth;
    } else {
        for (size_t k = 0; k < s->lights.size(); ++k) {
            const rtw_sphere& q = s->lights[k];
            lights[k] = Vec4T<T>{(T)q.cx, (T)q.cy, (T)q.cz, (T)q.r};
        }
    }
    std::vector<PlaneT<T>> planes(np);
    for (size_t k = 0; k < np; ++k) {
        const rtw_plane& q = s->planes[k];
        const rtw_material& m = s->materials[s->plane_material[k]];
        double len = std::sqrt(q.nx * q.nx + q.ny * q.ny + q.nz * q.nz);     // Plane::new normalises (plane.rs:35)
        PlaneT<T> p{};
        p.point = V3<T>{(T)q.px, (T)q.py, (T)q.pz};
        p.normal = V3<T>{(T)(q.nx / len), (T)(q.ny / len), (T)(q.nz / len)};
        p.info = ((uint32_t)k << 2) | (m.kind & 3u);
        p.albedo[0] = (T)m.r; p.albedo[1] = (T)m.g; p.albedo[2] = (T)m.b; p.param = (T)m.param;
        planes[k] = p;
    }
    CU(d.nodes.upload(nodes)); CU(d.spheres.upload(sph)); CU(d.sphere_mat.upload(mat)); CU(d.info.upload(info));
    CU(d.lights.upload(lights)); CU(d.planes.upload(planes)); CU(d.light_nodes.upload(light_nodes));
    d.view.nodes = d.nodes.p; d.view.top_nodes = d.nodes.p; d.view.n_top = 0;
    d.view.spheres = d.spheres.p; d.view.sphere_mat = d.sphere_mat.p; d.view.sphere_info = d.info.p;
    d.view.planes = d.planes.p; d.view.lights = d.lights.p;
    d.view.n_nodes = (int32_t)nodes.size(); d.view.n_spheres = (int32_t)ns; d.view.n_planes = (int32_t)np;
    d.view.n_lights = (int32_t)lights.size();
    d.view.light_nodes = d.light_nodes.p; d.view.n_light_nodes = (int32_t)light_nodes.size();
    return RTW_OK;
}

// opts.tmin < 0 (RTW_TMIN_REFERENCE): machine epsilon of the working precision, the reference's f64::EPSILON analogue
template <class T> static inline T resolve_tmin(double tmin) {
    if (tmin < 0.) return sizeof(T) == 8 ? (T)2.220446049250313e-16 : (T)1.1920928955078125e-07;
    return (T)tmin;
}

template <class T> CameraT<T> to_camera(const rtw_camera* c) {
    CameraT<T> k{};
    auto v = [](const double* p) { return V3<T>{(T)p[0], (T)p[1], (T)p[2]}; };
    k.center = v(c->center); k.pixel00 = v(c->pixel00_loc); k.du = v(c->pixel_delta_u); k.dv = v(c->pixel_delta_v);
    k.ddu = v(c->defocus_disk_u); k.ddv = v(c->defocus_disk_v); k.background = v(c->background);
    k.defocus_angle = (T)c->defocus_angle;
    // rand 0.8.5 Uniform::new_inclusive(-0.5, 0.5): scale = (high - low) / (1 - eps), decreased until
    // low + scale * (1 - eps) <= high
    const double max_rand = 1. - 2.220446049250313e-16;
    double scale = 1. / max_rand;
    while (scale * max_rand + (-0.5) > 0.5) scale = std::nextafter(scale, -INFINITY);
    k.jitter_scale = (T)scale;
    k.width = c->image_width; k.height = c->image_height; k.spp = c->samples_per_pixel; k.max_depth = c->max_depth;
    return k;
}

int check_camera(const rtw_camera* c) {
    if (!c) return fail(RTW_E_INVALID, "camera is NULL");
    if (c->image_width == 0 || c->image_height == 0) return fail(RTW_E_INVALID, "empty image");
    if ((uint64_t)c->image_width * c->image_height >= (1ull << 32)) return fail(RTW_E_INVALID, "image too large");
    return RTW_OK;
}
int check_opts(const rtw_opts* o) {
    if (!o) return fail(RTW_E_INVALID, "opts is NULL");
    if (o->precision != RTW_F32 && o->precision != RTW_F64) return fail(RTW_E_INVALID, "opts.precision");
    if (o->mode != RTW_MEGAKERNEL && o->mode != RTW_WAVEFRONT) return fail(RTW_E_INVALID, "opts.mode");
    if (o->tmin != o->tmin) return fail(RTW_E_INVALID, "opts.tmin is NaN");
    return RTW_OK;
}

void read_stats(const DeviceCounters& c, rtw_stats* st) {
    st->paths = c.paths; st->rays = c.rays; st->node_visits = c.node_visits; st->sphere_tests = c.sphere_tests;
    st->light_tests = c.light_tests; st->lambertian = c.lambertian; st->metal = c.metal; st->dielectric = c.dielectric;
    st->absorbed = c.absorbed; st->missed = c.missed; st->depth_out = c.depth_out;
}

template <class T, class Launch>
int render_tiles_t(rtw_scene* s, SceneDev<T>& d, const rtw_camera* cam, const rtw_opts* o, uint32_t rank, uint32_t world, T* tiles,
                   cudaStream_t stream, Launch launch) {
    RenderParams<T> P{};
    P.scene = d.view; P.cam = to_camera<T>(cam); P.seed = o->seed; P.tmin = resolve_tmin<T>(o->tmin); P.flags = o->flags;
    P.rank = rank; P.world = world;
    P.tiles_x = (cam->image_width + kTileW - 1) / kTileW;
    P.tiles_total = rtw_tiles_total(cam->image_width, cam->image_height);
    P.n_local_tiles = rtw_tiles_per_rank(cam->image_width, cam->image_height, world);
    P.tiles = tiles; P.work_counter = s->d_work; P.counters = s->d_counters;
    P.stack_depth = std::min<uint32_t>(kStackDepth, std::max(s->bvh.depth, s->light_bvh_depth) + 2);
    CU(cudaMemsetAsync(s->d_work, 0, sizeof(unsigned int), stream));
    CU(cudaMemsetAsync(s->d_counters, 0, sizeof(DeviceCounters), stream));
    CU(launch(P, (o->flags & RTW_FLAG_COUNT_EVENTS) != 0, s->sm_count, stream, &s->last_launch));
    return RTW_OK;
}

}  // namespace

namespace {
template <class T> BatchParams<T> batch_params(rtw_scene* s, SceneDev<T>& d, size_t n) {
    BatchParams<T> P{};
    P.scene = d.view; P.n = n;
    P.o = s->d_in0.p; P.d = s->d_in1.p; P.a = s->d_u0.p; P.b = s->d_u1.p; P.c = s->d_u2.p;
    P.prim = s->d_prim.p; P.t = s->d_out0.p; P.kind = s->d_k.p;
    P.p = s->d_out1.p; P.normal = s->d_out2.p; P.dir = s->d_out3.p; P.weight = s->d_out4.p; P.rgb = s->d_out1.p;
    return P;
}
int reserve_batch(rtw_scene* s, size_t n) {
    CU(s->d_in0.reserve(3 * n)); CU(s->d_in1.reserve(3 * n)); CU(s->d_out0.reserve(n)); CU(s->d_out1.reserve(3 * n));
    CU(s->d_out2.reserve(3 * n)); CU(s->d_out3.reserve(3 * n)); CU(s->d_out4.reserve(3 * n));
    CU(s->d_u0.reserve(n)); CU(s->d_u1.reserve(n)); CU(s->d_u2.reserve(n)); CU(s->d_k.reserve(n)); CU(s->d_prim.reserve(n));
    return RTW_OK;
}
}  // namespace

namespace {
// get_rays needs no scene: a tiny pool of device scratch per call
struct Scratch {
    DevBuf<uint32_t> a, b, c; DevBuf<double> o, d;
    ~Scratch() { a.release(); b.release(); c.release(); o.release(); d.release(); }
};
}  // namespace

// ------------------------------------------------------------------------------------------------
extern "C" {

int rtw_abi_version(void) { return RTW_ABI_VERSION; }
const char* rtw_last_error(void) { return g_err.c_str(); }

void rtw_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], out);
}

uint32_t rtw_tiles_total(uint32_t width, uint32_t height) {
    return ((width + kTileW - 1) / kTileW) * ((height + kTileH - 1) / kTileH);
}
uint32_t rtw_tiles_per_rank(uint32_t width, uint32_t height, uint32_t world) {
    if (world == 0) return 0;
    return (rtw_tiles_total(width, height) + world - 1) / world;
}

// CameraBuilder::build (shared/src/camera.rs:114-218)
int rtw_camera_build(const rtw_camera_builder* b, rtw_camera* out) {
    if (!b || !out) return fail(RTW_E_INVALID, "NULL argument");
    struct D3 { double x, y, z; };
    auto ld = [](const double* p) { return D3{p[0], p[1], p[2]}; };
    auto sub = [](D3 a, D3 c) { return D3{a.x - c.x, a.y - c.y, a.z - c.z}; };
    auto add = [](D3 a, D3 c) { return D3{a.x + c.x, a.y + c.y, a.z + c.z}; };
    auto mul = [](D3 a, double s) { return D3{a.x * s, a.y * s, a.z * s}; };
    auto dv = [](D3 a, double s) { return D3{a.x / s, a.y / s, a.z / s}; };
    auto crs = [](D3 a, D3 c) { return D3{a.y * c.z - a.z * c.y, a.z * c.x - a.x * c.z, a.x * c.y - a.y * c.x}; };
    auto nrm = [&](D3 a) { return dv(a, std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z)); };
    auto st = [](double* p, D3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; };
    bool A = b->has_aspect_ratio != 0, H = b->has_image_height != 0, W = b->has_image_width != 0;
    double aspect; uint32_t h, w;
    if (!A && !H && !W) { aspect = 1.; h = 100; w = 100; }
    else if (!A && !H && W) { aspect = 1.; h = b->image_width; w = b->image_width; }
    else if (!A && H && !W) { aspect = 1.; h = b->image_height; w = b->image_height; }
    else if (A && !H && !W) { aspect = b->aspect_ratio; h = (uint32_t)std::round(100. / aspect); w = 100; }
    else if (!A && H && W) { aspect = (double)b->image_width / (double)b->image_height; h = b->image_height; w = b->image_width; }
    else if (A && !H && W) { aspect = b->aspect_ratio; h = (uint32_t)std::round((double)b->image_width / aspect); w = b->image_width; }
    else if (A && H && !W) { aspect = b->aspect_ratio; h = b->image_height; w = (uint32_t)std::round((double)b->image_height * aspect); }
    else { aspect = b->aspect_ratio; h = b->image_height; w = b->image_width; }
    const double PI = 3.14159265358979323846264338327950288;
    D3 lookfrom = ld(b->lookfrom), lookat = ld(b->lookat), vup = ld(b->vup);
    double theta = b->vfov * (PI / 180.);
    double hh = std::tan(theta / 2.);
    double viewport_height = 2. * hh * b->focus_dist;
    double viewport_width = viewport_height * aspect;
    D3 wv = sub(lookfrom, lookat);
    D3 cx = crs(vup, wv);
    if (std::fabs(cx.x) < 1e-8 && std::fabs(cx.y) < 1e-8 && std::fabs(cx.z) < 1e-8) wv = add(wv, D3{0.1, 0., 0.});
    wv = nrm(wv);
    D3 u = nrm(crs(vup, wv));
    D3 v = crs(wv, u);
    D3 viewport_u = mul(u, viewport_width), viewport_v = mul(v, viewport_height);
    D3 du = dv(viewport_u, (double)w), dvv = dv(viewport_v, (double)h);
    D3 corner = sub(sub(sub(lookfrom, mul(wv, b->focus_dist)), dv(viewport_u, 2.)), dv(viewport_v, 2.));
    D3 p00 = add(corner, dv(add(du, dvv), 2.));
    double defocus_radius = std::tan(b->defocus_angle / 2.) * b->focus_dist;
    std::memset(out, 0, sizeof(*out));
    st(out->center, lookfrom); st(out->pixel00_loc, p00); st(out->pixel_delta_u, du); st(out->pixel_delta_v, dvv);
    st(out->defocus_disk_u, mul(u, defocus_radius)); st(out->defocus_disk_v, mul(v, defocus_radius));
    st(out->background, ld(b->background));
    out->defocus_angle = b->defocus_angle;
    out->image_width = w; out->image_height = h; out->samples_per_pixel = b->samples_per_pixel; out->max_depth = b->max_depth;
    return RTW_OK;
}

int rtw_release_cached_memory(void) {
    DevCache& c = dev_cache();
    std::lock_guard<std::mutex> g(c.m);
    for (auto& kv : c.blocks) { cudaSetDevice(kv.first.first); cudaFree(kv.second); }
    c.blocks.clear(); c.bytes = 0;
    return RTW_OK;
}

int rtw_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) { cudaGetLastError(); return 0; }
    if (e != cudaSuccess) return fail(RTW_E_CUDA, cudaGetErrorString(e));
    return n;
}

int rtw_scene_create(const rtw_sphere* spheres, const uint32_t* sphere_material, size_t n_spheres,
                     const rtw_plane* planes, const uint32_t* plane_material, size_t n_planes,
                     const rtw_material* materials, size_t n_materials,
                     const rtw_sphere* lights, size_t n_lights, rtw_scene** out) {
    if (!out) return fail(RTW_E_INVALID, "out is NULL");
    *out = nullptr;
    if ((n_spheres && (!spheres || !sphere_material)) || (n_planes && (!planes || !plane_material)) ||
        (n_materials && !materials) || (n_lights && !lights))
        return fail(RTW_E_INVALID, "NULL array with non-zero count");
    if (n_spheres >= (1u << 27) || n_planes >= (1u << 16)) return fail(RTW_E_UNSUPPORTED, "too many primitives");
    bool any_lambertian = false;
    for (size_t i = 0; i < n_spheres; ++i) {
        if (sphere_material[i] >= n_materials) return fail(RTW_E_INVALID, "sphere material index out of range");
        if (!(spheres[i].r > 0.) || !std::isfinite(spheres[i].cx + spheres[i].cy + spheres[i].cz + spheres[i].r))
            return fail(RTW_E_INVALID, "sphere with non-finite centre or non-positive radius");
        any_lambertian |= materials[sphere_material[i]].kind == RTW_LAMBERTIAN;
    }
    for (size_t i = 0; i < n_planes; ++i) {
        if (plane_material[i] >= n_materials) return fail(RTW_E_INVALID, "plane material index out of range");
        any_lambertian |= materials[plane_material[i]].kind == RTW_LAMBERTIAN;
    }
    for (size_t i = 0; i < n_materials; ++i)
        if (materials[i].kind > RTW_INVISIBLE) return fail(RTW_E_UNSUPPORTED, "material kind outside Lambertian/Metal/Dialectric/Invisible");
    if (any_lambertian && n_lights == 0)
        return fail(RTW_E_INVALID, "Lambertian material with an empty lights list (the reference panics: HittableList shouldn't be empty)");
    int ndev = rtw_device_count();
    if (ndev < 0) return ndev;
    if (ndev == 0) return fail(RTW_E_NO_DEVICE, "no CUDA device: this backend has no CPU fallback");

    rtw_scene* s = new rtw_scene();
    s->spheres.assign(spheres, spheres + n_spheres); s->sphere_material.assign(sphere_material, sphere_material + n_spheres);
    s->planes.assign(planes, planes + n_planes); s->plane_material.assign(plane_material, plane_material + n_planes);
    s->materials.assign(materials, materials + n_materials); s->lights.assign(lights, lights + n_lights);
    host::Builder builder;
    int max_leaf = 2;                                     // measured best of 1..8 on `simple` (profiles/README.md); tuning knob for experiments: RTW_BVH_MAX_LEAF=1..8
    if (const char* e = std::getenv("RTW_BVH_MAX_LEAF")) { int v = std::atoi(e); if (v >= 1 && v <= 8) max_leaf = v; }
    s->bvh = builder.build(reinterpret_cast<const double*>(s->spheres.data()), n_spheres, max_leaf, kMaxTreeDepth);
    auto bail = [&](int code) { rtw_scene_destroy(s); return code; };
    cudaError_t e = cudaGetDevice(&s->device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&s->sm_count, cudaDevAttrMultiProcessorCount, s->device);
    if (e == cudaSuccess) e = cached_malloc(reinterpret_cast<void**>(&s->d_work), sizeof(unsigned int));
    if (e == cudaSuccess) e = cached_malloc(reinterpret_cast<void**>(&s->d_counters), sizeof(DeviceCounters));
    for (int i = 0; i < 4 && e == cudaSuccess; ++i) e = cudaEventCreate(&s->ev[i]);
    if (e != cudaSuccess) { fail(RTW_E_CUDA, cudaGetErrorString(e)); return bail(RTW_E_CUDA); }
    int rc = upload_scene<float>(s, s->f32);
    if (rc == RTW_OK) rc = upload_scene<double>(s, s->f64);
    if (rc != RTW_OK) return bail(rc);
    *out = s;
    return RTW_OK;
}

void rtw_scene_destroy(rtw_scene* s) {
    if (!s) return;
    s->f32.release(); s->f64.release();
    cached_free(s->d_work, sizeof(unsigned int));
    cached_free(s->d_counters, sizeof(DeviceCounters));
    s->d_rgb_sum.release(); s->d_rgb8.release(); s->d_accum.release(); s->d_poison.release();
    s->d_in0.release(); s->d_in1.release(); s->d_out0.release(); s->d_out1.release(); s->d_out2.release(); s->d_out3.release();
    s->d_out4.release(); s->d_u0.release(); s->d_u1.release(); s->d_u2.release(); s->d_k.release(); s->d_prim.release();
    for (auto& ev : s->ev) if (ev) cudaEventDestroy(ev);
    delete s;
}

int rtw_scene_info(const rtw_scene* s, uint64_t out[5]) {
    if (!s || !out) return fail(RTW_E_INVALID, "NULL argument");
    out[0] = s->bvh.nodes.size(); out[1] = s->bvh.leaves; out[2] = s->bvh.depth; out[3] = s->bvh.max_leaf;
    out[4] = s->f32.bytes() + s->f64.bytes();
    return RTW_OK;
}

int rtw_render_tiles_device(rtw_scene* s, const rtw_camera* cam, const rtw_opts* o, uint32_t rank, uint32_t world, void* d_tiles,
                            void* stream, rtw_stats* stats) {
    if (!s || !d_tiles) return fail(RTW_E_INVALID, "NULL argument");
    int rc = check_camera(cam); if (rc) return rc;
    rc = check_opts(o); if (rc) return rc;
    if (world == 0 || rank >= world) return fail(RTW_E_INVALID, "rank/world");
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t launches = 1;
    // RTW_F64 has one renderer (lane per pixel, samples summed in order); `mode` only selects among the FP32 renderers
    bool pooled = o->precision == RTW_F32 && (o->mode == RTW_WAVEFRONT || !(o->flags & RTW_FLAG_LANE_PER_PIXEL));
    if (pooled) {
        size_t n_slots = (size_t)rtw_tiles_per_rank(cam->image_width, cam->image_height, world) * kTileW * kTileH;
        CU(s->d_accum.reserve(n_slots * 3)); CU(s->d_poison.reserve(n_slots));
    }
    CU(cudaEventRecord(s->ev[0], st));
    if (pooled) {
        PoolParams Q{};
        Q.accum = s->d_accum.p; Q.poison = s->d_poison.p;
        Q.pixels_per_chunk = pool_pixels_per_chunk(cam->samples_per_pixel);
        uint32_t n_slots = rtw_tiles_per_rank(cam->image_width, cam->image_height, world) * kTileW * kTileH;
        Q.n_chunks = (n_slots + Q.pixels_per_chunk - 1) / Q.pixels_per_chunk;
        // the wavefront packs the remaining depth into 16 bits; deeper paths take the (bit-identical) megakernel
        const bool wavefront = o->mode == RTW_WAVEFRONT && cam->max_depth <= 0xffffu;
        const uint32_t bvh_depth = std::max(s->bvh.depth, s->light_bvh_depth);
        auto launch = [&](RenderParams<float> P, bool count, int sms, cudaStream_t str, LaunchInfo* info) {
            return wavefront ? launch_render_wavefront_f32(P, Q, bvh_depth, count, sms, str, info)
                             : launch_render_pool_f32(P, Q, count, sms, str, info);
        };
        rc = render_tiles_t<float>(s, s->f32, cam, o, rank, world, (float*)d_tiles, st, launch);
        launches = 2;
    } else if (o->precision == RTW_F32) rc = render_tiles_t<float>(s, s->f32, cam, o, rank, world, (float*)d_tiles, st, launch_render_f32);
    else rc = render_tiles_t<double>(s, s->f64, cam, o, rank, world, (double*)d_tiles, st, launch_render_f64);
    if (rc) return rc;
    s->last_launches = launches;
    CU(cudaEventRecord(s->ev[1], st));
    if (stats) {
        DeviceCounters c;
        CU(cudaMemcpyAsync(&c, s->d_counters, sizeof(c), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        std::memset(stats, 0, sizeof(*stats));
        read_stats(c, stats);
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, s->ev[0], s->ev[1]));
        stats->kernel_ms = ms; stats->total_ms = ms; stats->launches = launches;
    }
    return RTW_OK;
}

int rtw_untile_resolve_device(const void* d_tiles_all, uint32_t precision, uint32_t width, uint32_t height, uint32_t world,
                              uint32_t spp, double* d_rgb_sum, uint8_t* d_rgb8, void* stream) {
    if (!d_tiles_all || world == 0 || width == 0 || height == 0) return fail(RTW_E_INVALID, "bad argument");
    uint32_t tpr = rtw_tiles_per_rank(width, height, world);
    cudaStream_t st = (cudaStream_t)stream;
    if (precision == RTW_F32) CU(launch_untile_f32((const float*)d_tiles_all, width, height, world, tpr, spp, d_rgb_sum, d_rgb8, st));
    else if (precision == RTW_F64) CU(launch_untile_f64((const double*)d_tiles_all, width, height, world, tpr, spp, d_rgb_sum, d_rgb8, st));
    else return fail(RTW_E_INVALID, "precision");
    return RTW_OK;
}

int rtw_render(rtw_scene* s, const rtw_camera* cam, const rtw_opts* o, double* rgb_sum, uint8_t* rgb8, rtw_stats* stats) {
    if (!s) return fail(RTW_E_INVALID, "scene is NULL");
    int rc = check_camera(cam); if (rc) return rc;
    rc = check_opts(o); if (rc) return rc;
    size_t npx = (size_t)cam->image_width * cam->image_height;
    size_t tile_elems = (size_t)rtw_tiles_per_rank(cam->image_width, cam->image_height, 1) * kTileW * kTileH * 3;
    void* tiles;
    if (o->precision == RTW_F32) { CU(s->f32.tiles.reserve(tile_elems)); tiles = s->f32.tiles.p; }
    else { CU(s->f64.tiles.reserve(tile_elems)); tiles = s->f64.tiles.p; }
    if (rgb_sum) CU(s->d_rgb_sum.reserve(npx * 3));
    if (rgb8) CU(s->d_rgb8.reserve(npx * 3));
    CU(cudaEventRecord(s->ev[2], 0));
    rc = rtw_render_tiles_device(s, cam, o, 0, 1, tiles, nullptr, nullptr);
    if (rc) return rc;
    uint32_t launches = s->last_launches;
    if (rgb_sum || rgb8) {
        rc = rtw_untile_resolve_device(tiles, o->precision, cam->image_width, cam->image_height, 1, cam->samples_per_pixel,
                                       rgb_sum ? s->d_rgb_sum.p : nullptr, rgb8 ? s->d_rgb8.p : nullptr, nullptr);
        if (rc) return rc;
        launches++;
        if (rgb_sum) CU(cudaMemcpyAsync(rgb_sum, s->d_rgb_sum.p, npx * 3 * sizeof(double), cudaMemcpyDeviceToHost, 0));
        if (rgb8) CU(cudaMemcpyAsync(rgb8, s->d_rgb8.p, npx * 3, cudaMemcpyDeviceToHost, 0));
    }
    CU(cudaEventRecord(s->ev[3], 0));
    DeviceCounters c;
    CU(cudaMemcpy(&c, s->d_counters, sizeof(c), cudaMemcpyDeviceToHost));
    CU(cudaEventSynchronize(s->ev[3]));
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        read_stats(c, stats);
        float k = 0.f, t = 0.f;
        CU(cudaEventElapsedTime(&k, s->ev[0], s->ev[1]));
        CU(cudaEventElapsedTime(&t, s->ev[2], s->ev[3]));
        stats->kernel_ms = k; stats->total_ms = t; stats->launches = launches;
    }
    return RTW_OK;
}

// ---- batch entry points --------------------------------------------------------------------------

int rtw_trace_batch(rtw_scene* s, const double* o, const double* d, size_t n, double tmin, double tmax, uint32_t precision,
                    int32_t* prim_id, double* t) {
    if (!s || (n && (!o || !d || !prim_id || !t))) return fail(RTW_E_INVALID, "NULL argument");
    if (precision > RTW_F64) return fail(RTW_E_INVALID, "precision");
    if (n == 0) return RTW_OK;
    int rc = reserve_batch(s, n); if (rc) return rc;
    CU(cudaMemcpy(s->d_in0.p, o, 3 * n * sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(s->d_in1.p, d, 3 * n * sizeof(double), cudaMemcpyHostToDevice));
    if (precision == RTW_F32) {
        BatchParams<float> P = batch_params<float>(s, s->f32, n);
        P.tmin = (float)tmin; P.tmax = (float)tmax;
        CU(launch_trace_f32(P, 0));
    } else {
        BatchParams<double> P = batch_params<double>(s, s->f64, n);
        P.tmin = tmin; P.tmax = tmax;
        CU(launch_trace_f64(P, 0));
    }
    CU(cudaMemcpy(prim_id, s->d_prim.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(t, s->d_out0.p, n * sizeof(double), cudaMemcpyDeviceToHost));
    return RTW_OK;
}

int rtw_scatter_batch(rtw_scene* s, const rtw_opts* opts, const double* o, const double* d, size_t n, const uint32_t* pixel,
                      const uint32_t* sample, const uint32_t* vertex, int32_t* prim_id, double* t, uint32_t* kind, double* p,
                      double* normal, double* dir, double* weight) {
    if (!s || (n && (!o || !d || !pixel || !sample || !vertex || !prim_id || !t || !kind || !p || !normal || !dir || !weight)))
        return fail(RTW_E_INVALID, "NULL argument");
    int rc = check_opts(opts); if (rc) return rc;
    if (n == 0) return RTW_OK;
    rc = reserve_batch(s, n); if (rc) return rc;
    CU(cudaMemcpy(s->d_in0.p, o, 3 * n * sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(s->d_in1.p, d, 3 * n * sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(s->d_u0.p, pixel, n * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(s->d_u1.p, sample, n * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(s->d_u2.p, vertex, n * 4, cudaMemcpyHostToDevice));
    if (opts->precision == RTW_F32) {
        BatchParams<float> P = batch_params<float>(s, s->f32, n);
        P.seed = opts->seed; P.tmin = resolve_tmin<float>(opts->tmin); P.flags = opts->flags;
        CU(launch_scatter_f32(P, 0));
    } else {
        BatchParams<double> P = batch_params<double>(s, s->f64, n);
        P.seed = opts->seed; P.tmin = resolve_tmin<double>(opts->tmin); P.flags = opts->flags;
        CU(launch_scatter_f64(P, 0));
    }
    CU(cudaMemcpy(prim_id, s->d_prim.p, n * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(t, s->d_out0.p, n * 8, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(kind, s->d_k.p, n * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(p, s->d_out1.p, 3 * n * 8, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(normal, s->d_out2.p, 3 * n * 8, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(dir, s->d_out3.p, 3 * n * 8, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(weight, s->d_out4.p, 3 * n * 8, cudaMemcpyDeviceToHost));
    return RTW_OK;
}


int rtw_get_rays(const rtw_camera* cam, const rtw_opts* opts, const uint32_t* i, const uint32_t* j, const uint32_t* sample, size_t n,
                 double* o, double* d) {
    int rc = check_camera(cam); if (rc) return rc;
    rc = check_opts(opts); if (rc) return rc;
    if (n && (!i || !j || !sample || !o || !d)) return fail(RTW_E_INVALID, "NULL argument");
    if (n == 0) return RTW_OK;
    int ndev = rtw_device_count();
    if (ndev <= 0) return ndev < 0 ? ndev : fail(RTW_E_NO_DEVICE, "no CUDA device: this backend has no CPU fallback");
    Scratch sc;
    CU(sc.a.reserve(n)); CU(sc.b.reserve(n)); CU(sc.c.reserve(n)); CU(sc.o.reserve(3 * n)); CU(sc.d.reserve(3 * n));
    CU(cudaMemcpy(sc.a.p, i, n * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(sc.b.p, j, n * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(sc.c.p, sample, n * 4, cudaMemcpyHostToDevice));
    if (opts->precision == RTW_F32) {
        BatchParams<float> P{}; P.n = n; P.a = sc.a.p; P.b = sc.b.p; P.c = sc.c.p; P.cam = to_camera<float>(cam); P.seed = opts->seed;
        CU(launch_get_rays_f32(P, sc.o.p, sc.d.p, 0));
    } else {
        BatchParams<double> P{}; P.n = n; P.a = sc.a.p; P.b = sc.b.p; P.c = sc.c.p; P.cam = to_camera<double>(cam); P.seed = opts->seed;
        CU(launch_get_rays_f64(P, sc.o.p, sc.d.p, 0));
    }
    CU(cudaMemcpy(o, sc.o.p, 3 * n * 8, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(d, sc.d.p, 3 * n * 8, cudaMemcpyDeviceToHost));
    return RTW_OK;
}

int rtw_path_radiance(rtw_scene* s, const rtw_camera* cam, const rtw_opts* opts, const uint32_t* i, const uint32_t* j,
                      const uint32_t* sample, size_t n, double* rgb) {
    if (!s) return fail(RTW_E_INVALID, "scene is NULL");
    int rc = check_camera(cam); if (rc) return rc;
    rc = check_opts(opts); if (rc) return rc;
    if (n && (!i || !j || !sample || !rgb)) return fail(RTW_E_INVALID, "NULL argument");
    if (n == 0) return RTW_OK;
    rc = reserve_batch(s, n); if (rc) return rc;
    CU(cudaMemcpy(s->d_u0.p, i, n * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(s->d_u1.p, j, n * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(s->d_u2.p, sample, n * 4, cudaMemcpyHostToDevice));
    if (opts->precision == RTW_F32) {
        BatchParams<float> P = batch_params<float>(s, s->f32, n);
        P.cam = to_camera<float>(cam); P.seed = opts->seed; P.tmin = resolve_tmin<float>(opts->tmin); P.flags = opts->flags;
        CU(launch_path_radiance_f32(P, 0));
    } else {
        BatchParams<double> P = batch_params<double>(s, s->f64, n);
        P.cam = to_camera<double>(cam); P.seed = opts->seed; P.tmin = resolve_tmin<double>(opts->tmin); P.flags = opts->flags;
        CU(launch_path_radiance_f64(P, 0));
    }
    CU(cudaMemcpy(rgb, s->d_out1.p, 3 * n * 8, cudaMemcpyDeviceToHost));
    return RTW_OK;
}

}  // extern "C"
