// capi.cu — the extern "C" surface declared in include/rtw.h: scene upload, Camera::render and the
// per-ray batch operations, on the current CUDA device.  No CPU fallback: every compute entry point
// fails with RTW_E_NO_DEVICE / RTW_E_CUDA when the GPU path is unavailable.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/rtw.h"
#include "bvh_build.hpp"
#include "bvh_device.hpp"
#include "general_host.hpp"
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

constexpr size_t kFlatListLimit = 8;        // general scenes with at most this many bounded entries skip the BVH (linear, kind-sorted walk)
constexpr size_t kLightBvhThreshold = 64;   // more lights than this: BVH over the lights (FP32 path)
constexpr size_t kConnectThreshold = 2048;  // more lights than this: the wavefront runs the light walk as its own stage (CONNECT)

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
// Blocks are keyed by the device they were allocated on (recorded in the DevBuf, not read back from cudaGetDevice at free time:
// a scene may be destroyed while another GPU is current).  A block only enters the cache after its device has gone idle:
// cudaFree would have synchronised implicitly, and without that a kernel still running on one stream could see its buffer handed to
// another scene or stream (rtw_render_*_device with stats == NULL return with work in flight).
cudaError_t cached_malloc(void** p, size_t n, int* dev_out) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev_out) *dev_out = dev;
    n = round_alloc(n);
    {
        DevCache& c = dev_cache();
        std::lock_guard<std::mutex> g(c.m);
        auto it = c.blocks.find({dev, n});
        if (it != c.blocks.end()) { *p = it->second; c.blocks.erase(it); c.bytes -= n; return cudaSuccess; }
    }
    return cudaMalloc(p, n);
}
void cached_free(void* p, size_t n, int dev) {
    if (!p) return;
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess) { cudaGetLastError(); return; }
    if (cur != dev && cudaSetDevice(dev) != cudaSuccess) { cudaGetLastError(); return; }      // leak rather than free on the wrong device
    n = round_alloc(n);
    bool cached = false;
    if (cudaDeviceSynchronize() == cudaSuccess) {
        DevCache& c = dev_cache();
        std::lock_guard<std::mutex> g(c.m);
        if (c.bytes + n <= DevCache::kMaxBytes) { c.blocks.insert({{dev, n}, p}); c.bytes += n; cached = true; }
    } else cudaGetLastError();
    if (!cached) cudaFree(p);
    if (cur != dev) cudaSetDevice(cur);
}

template <class P> struct DevBuf {
    P* p = nullptr; size_t n = 0; int dev = 0;
    cudaError_t upload(const std::vector<P>& h) {
        release();
        n = h.size();
        if (!n) return cudaSuccess;
        cudaError_t e = cached_malloc(reinterpret_cast<void**>(&p), n * sizeof(P), &dev);
        if (e != cudaSuccess) { p = nullptr; n = 0; return e; }
        return cudaMemcpy(p, h.data(), n * sizeof(P), cudaMemcpyHostToDevice);
    }
    cudaError_t reserve(size_t count) {
        if (count <= n && p) return cudaSuccess;
        release();
        cudaError_t e = cached_malloc(reinterpret_cast<void**>(&p), count * sizeof(P), &dev);
        if (e != cudaSuccess) { p = nullptr; return e; }
        n = count;
        return cudaSuccess;
    }
    void release() { if (p) cached_free(p, n * sizeof(P), dev); p = nullptr; n = 0; }
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
    DevBuf<Node<T>> nodes; DevBuf<LNode> light_nodes; DevBuf<Vec4T<T>> spheres, sphere_mat, lights; DevBuf<uint32_t> info; DevBuf<PlaneT<T>> planes;
    DevBuf<T> tiles;
    DevBuf<unsigned char> nodes_staged;      // FP32, small scenes: the nodes again at stride kShNodeStridePadded (rtw_device.cuh)
    SceneView<T> view{};
    size_t bytes() const { return nodes_staged.bytes() + light_nodes.bytes() + nodes.bytes() + spheres.bytes() + sphere_mat.bytes() + lights.bytes() + info.bytes() + planes.bytes(); }
    void release() { nodes_staged.release(); light_nodes.release(); nodes.release(); spheres.release(); sphere_mat.release(); lights.release(); info.release(); planes.release(); tiles.release(); }
};

template <class T> struct SceneDevG {
    DevBuf<Node<T>> nodes; DevBuf<GPrim<T>> prims, unbounded, lights; DevBuf<Vec4T<T>> spheres; DevBuf<GPlane<T>> plane_geo;
    DevBuf<GQuad<T>> quads; DevBuf<GXform<T>> xforms; DevBuf<GMat<T>> mats; DevBuf<GTex<T>> textures; DevBuf<GPerlin<T>> perlins;
    SceneViewG<T> view{};
    size_t bytes() const {
        return nodes.bytes() + prims.bytes() + unbounded.bytes() + lights.bytes() + spheres.bytes() + plane_geo.bytes() + quads.bytes() +
               xforms.bytes() + mats.bytes() + textures.bytes() + perlins.bytes();
    }
    void release() {
        nodes.release(); prims.release(); unbounded.release(); lights.release(); spheres.release(); plane_geo.release(); quads.release();
        xforms.release(); mats.release(); textures.release(); perlins.release();
    }
};

// host copy of a general scene description (rtw_scene_desc owns nothing)
struct GeneralDesc {
    std::vector<rtw_sphere> spheres; std::vector<rtw_plane> planes; std::vector<rtw_quad> quads; std::vector<rtw_cuboid> cuboids;
    std::vector<rtw_transform> transforms; std::vector<rtw_material> materials; std::vector<rtw_texture> textures;
    std::vector<rtw_perlin> perlins; std::vector<rtw_prim> world, lights;
    bool lights_is_bvh = false;
};

}  // namespace

struct MultiReplica;                                 // capi_multi.inl: per-device state of rtw_render_multi
struct rtw_scene {
    std::vector<MultiReplica*> replicas;             // rtw_render_multi: replicas[0] is this handle's own device
    DevBuf<unsigned long long> d_block;              // rtw_render_rank: this rank's packed accumulators ([3 * slots] u64 + [slots] u32)
    DevBuf<unsigned char> d_gather, d_tiles_rank;    // tile partition: the buffers gathered on the root / this rank's tiles
    bool general = false;
    GeneralDesc gdesc;
    DevBuf<uint32_t> d_panic;                    // general scenes: "the reference would have panicked" flag written by the kernels
    SceneDevG<float> g32; SceneDevG<double> g64;
    std::vector<rtw_sphere> spheres; std::vector<uint32_t> sphere_material;
    std::vector<rtw_plane> planes; std::vector<uint32_t> plane_material;
    std::vector<rtw_material> materials; std::vector<rtw_sphere> lights;
    host::Bvh bvh;
    SceneDev<float> f32; SceneDev<double> f64;
    int device = 0, sm_count = 0;
    unsigned int* d_work = nullptr; DeviceCounters* d_counters = nullptr;
    DevBuf<double> d_rgb_sum; DevBuf<uint8_t> d_rgb8;
    DevBuf<unsigned long long> d_accum; DevBuf<uint32_t> d_poison;   // pooled megakernel accumulators
    DevBuf<uint4> d_cand;                                            // candidate lists of the camera rays, rebuilt by every render call
    DevBuf<unsigned char> d_cand_blocks;                             // ... and the per-block lists of their first level
    DevBuf<uint32_t> d_order;                                        // chunk order of the work queue (costly chunks first), rebuilt by every render call
    DevBuf<double> d_in2, d_in3; DevBuf<uint32_t> d_u3, d_u4;          // rtw_shade_batch inputs
    DevBuf<double> d_in0, d_in1, d_out0, d_out1, d_out2, d_out3, d_out4; DevBuf<uint32_t> d_u0, d_u1, d_u2, d_k; DevBuf<int32_t> d_prim;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaStream_t side_stream = nullptr;                              // render_background_kernel runs here, next to the wavefront kernel (set_background_side_stream)
    cudaEvent_t ev_side[2] = {nullptr, nullptr};
    LaunchInfo last_launch;
    uint32_t last_launches = 1;
    uint32_t light_bvh_depth = 0;
    int bvh_builder = RTW_BVH_HOST_SAH;          // which builder made the world BVH
    DeviceBvhInfo device_bvh;
};

namespace {
void multi_release(rtw_scene* s);

// w component of a light record: the radius on the exact path (Sphere::hit needs it), its FP32 square on the fast path, whose
// light test and cone sampling only ever use r^2 (one multiply less per light test: 30 G of them per 1080p / 500 spp frame)
inline double light_w(double r) { return r; }
inline float light_w(float r) { volatile float r2 = r * r; return r2; }

constexpr size_t kIsolationMaxSpheres = 4096;

template <class T> int upload_scene(rtw_scene* s, SceneDev<T>& d, bool world_on_device = false) {
    const host::Bvh& b = s->bvh;
    std::vector<Node<T>> nodes(world_on_device ? 0 : b.nodes.size());
    for (size_t i = 0; i < nodes.size(); ++i) {
        const host::FlatNode& f = b.nodes[i];
        Node<T> n{};
        fill_node<T>(n, f);
        n.left = f.left >= 0 ? f.left : encode_leaf(f.lfirst, f.lcount);
        n.right = f.right >= 0 ? f.right : encode_leaf(f.rfirst, f.rcount);
        nodes[i] = n;
    }
    size_t ns = s->spheres.size(), np = s->planes.size();
    const size_t ns_host = world_on_device ? 0 : ns;       // device-built: nodes and sorted sphere arrays are already in place
    std::vector<Vec4T<T>> sph(ns_host), mat(ns_host), lights(s->lights.size());
    std::vector<uint32_t> info(ns_host);
    for (size_t k = 0; k < ns_host; ++k) {
        uint32_t src = b.order[k];
        const rtw_sphere& q = s->spheres[src];
        const rtw_material& m = s->materials[s->sphere_material[src]];
        sph[k] = Vec4T<T>{(T)q.cx, (T)q.cy, (T)q.cz, (T)q.r};
        mat[k] = Vec4T<T>{(T)m.r, (T)m.g, (T)m.b, (T)m.param};
        info[k] = ((uint32_t)(np + src) << 2) | (m.kind & 3u);
    }
    // kSphereIsolated (see closest_prim_self): no other sphere's SURFACE comes near the ball of this one — a ray that leaves the sphere
    // and meets it again cannot meet anything else first.  Distance from c_s to the surface of o is | |c_s - c_o| - r_o |; the ball is
    // inflated by 0.1 % plus 1e-4 of the lengths involved, far beyond the FP32 error of a root.  O(n^2) on the host, small scenes only.
    static const bool iso_allowed = [] { const char* e = std::getenv("RTW_NO_SELF_HIT"); return !(e && std::atoi(e) == 1); }();   // A/B measurements
    if (sizeof(T) == 4 && iso_allowed && ns_host <= kIsolationMaxSpheres && ns + np < (1u << 29)) {
        for (size_t k = 0; k < ns_host; ++k) {
            const rtw_sphere& a = s->spheres[b.order[k]];
            bool isolated = a.r > 0. && std::isfinite(a.r);
            for (size_t j = 0; j < ns_host && isolated; ++j) {
                if (j == k) continue;
                const rtw_sphere& o = s->spheres[b.order[j]];
                const double dx = a.cx - o.cx, dy = a.cy - o.cy, dz = a.cz - o.cz, D = std::sqrt(dx * dx + dy * dy + dz * dz);
                const double R = a.r * 1.001 + 1e-4 * (D + std::fabs(o.r)) + 1e-6;
                if (!(std::fabs(D - std::fabs(o.r)) > R)) isolated = false;
            }
            if (isolated) info[k] |= kSphereIsolated;
        }
    }
    // lights: insertion order; on the FP32 path a long list gets its own BVH (leaf order) so that
    // lights.pdf_value is not O(#lights) per diffuse bounce
    std::vector<LNode> light_nodes;
    const bool light_bvh = sizeof(T) == 4 && s->lights.size() > kLightBvhThreshold;
    if (light_bvh) {
        host::Builder lb;
        // large light lists (L2-resident tree, latency-bound walk): one light per leaf, stored in the node itself; smaller ones keep
        // 4-light leaf ranges — half the nodes, and the whole tree stays in L1 (measured on C5, 399 lights: 20.2 vs 21.3 ms)
        const int light_leaf = s->lights.size() > kConnectThreshold ? 1 : 4;
        host::Bvh lbvh = lb.build(reinterpret_cast<const double*>(s->lights.data()), s->lights.size(), light_leaf, kMaxTreeDepth);
        // binned-SAH tree (inner nodes carrying both child boxes, breadth-first) -> depth-first list of the children, each with its
        // own box and a skip link: the stackless layout of LNode
        struct Emit {
            const host::Bvh& b; std::vector<LNode>& out; const std::vector<rtw_sphere>& src;
            void child(const host::Box& box, int32_t inner, uint32_t first, uint32_t count) {
                if (inner < 0 && count == 1) {                          // the light itself: (centre, r^2) exactly as in the light list
                    const rtw_sphere& q = src[b.order[first]];
                    LNode n{};
                    n.c[0] = (float)q.cx; n.c[1] = (float)q.cy; n.c[2] = (float)q.cz; n.hx = light_w((float)q.r);
                    n.leaf = kLNodeLight; n.skip = (int32_t)out.size() + 1;
                    out.push_back(n);
                    return;
                }
                Node<float> tmp{};
                host::FlatNode f{}; f.lbox = box; f.rbox = box;
                fill_node<float>(tmp, f);                               // conservative FP32 centre / half-extent, as for the world tree
                LNode n{};
                for (int a = 0; a < 3; ++a) n.c[a] = tmp.la[a];
                n.hx = tmp.lb[0]; n.hy = tmp.lb[1]; n.hz = tmp.lb[2];
                n.leaf = inner >= 0 ? kLNodeInner : (count == 0 ? kLNodeEmpty : (int32_t)((first << 4) | (count - 1)));
                const size_t me = out.size();
                out.push_back(n);
                if (inner >= 0) node(inner);
                out[me].skip = (int32_t)out.size();                     // patched to -1 for "past the end" below
            }
            void node(int32_t i) {
                const host::FlatNode& f = b.nodes[(size_t)i];
                child(f.lbox, f.left, f.lfirst, f.lcount);
                child(f.rbox, f.right, f.rfirst, f.rcount);
            }
        } emit{lbvh, light_nodes, s->lights};
        emit.node(0);
        for (LNode& n : light_nodes) if (n.skip >= (int32_t)light_nodes.size()) n.skip = -1;
        for (size_t k = 0; k < s->lights.size(); ++k) {
            const rtw_sphere& q = s->lights[lbvh.order[k]];
            lights[k] = Vec4T<T>{(T)q.cx, (T)q.cy, (T)q.cz, light_w((T)q.r)};
        }
        s->light_bvh_depth = 0;                                         // the walk is stackless: the light tree asks nothing of the traversal stacks
    } else {
        for (size_t k = 0; k < s->lights.size(); ++k) {
            const rtw_sphere& q = s->lights[k];
            lights[k] = Vec4T<T>{(T)q.cx, (T)q.cy, (T)q.cz, light_w((T)q.r)};
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
    if (!world_on_device) { CU(d.nodes.upload(nodes)); CU(d.spheres.upload(sph)); CU(d.sphere_mat.upload(mat)); CU(d.info.upload(info)); }
    CU(d.lights.upload(lights)); CU(d.planes.upload(planes)); CU(d.light_nodes.upload(light_nodes));
    d.view.nodes = d.nodes.p; d.view.top_nodes = d.nodes.p; d.view.n_top = 0;
    d.view.spheres = d.spheres.p; d.view.sphere_mat = d.sphere_mat.p; d.view.sphere_info = d.info.p;
    d.view.planes = d.planes.p; d.view.lights = d.lights.p;
    d.view.n_nodes = (int32_t)(world_on_device ? s->device_bvh.node_slots : nodes.size()); d.view.n_spheres = (int32_t)ns; d.view.n_planes = (int32_t)np;
    d.view.n_lights = (int32_t)lights.size();
    d.view.light_nodes = d.light_nodes.p; d.view.n_light_nodes = (int32_t)light_nodes.size();
    d.view.connect_stage = light_bvh && s->lights.size() > kConnectThreshold ? 1 : 0;
    d.view.nodes_staged = nullptr;
    if (sizeof(T) == 4 && d.view.n_nodes > 0 && (size_t)d.view.n_nodes * kShNodeStridePadded <= (64u << 10)) {
        // the copy the all-shared kernels stage: same 64-byte nodes, 80-byte stride (bank-conflict-free LDS.128)
        const size_t n = (size_t)d.view.n_nodes;
        CU(d.nodes_staged.reserve(n * kShNodeStridePadded + 16));
        CU(cudaMemset(d.nodes_staged.p, 0, n * kShNodeStridePadded + 16));
        CU(cudaMemcpy2D(d.nodes_staged.p, kShNodeStridePadded, d.nodes.p, sizeof(Node<T>), sizeof(Node<T>), n, cudaMemcpyDeviceToDevice));
        d.view.nodes_staged = d.nodes_staged.p;
    }
    return RTW_OK;
}

// ---- general scenes ------------------------------------------------------------------------------------------
// iteration order of a HittableList: buckets by TypeId (hittable_list.rs:270-294) — compiler-chosen in the reference,
// fixed here as [Plane, Sphere, Quad, Triangle, Cuboid, then the Transformed<...> of each]
uint32_t bucket_of(const rtw_prim& e) {
    uint32_t k = e.kind == RTW_PRIM_PLANE ? 0u : e.kind == RTW_PRIM_SPHERE ? 1u : e.kind;
    return k + (e.transform >= 0 ? 5u : 0u);
}

template <class T> int upload_general(rtw_scene* s, SceneDevG<T>& d) {
    const GeneralDesc& g = s->gdesc;
    auto v3 = [](host::D3 a) { return V3<T>{(T)a.x, (T)a.y, (T)a.z}; };
    std::vector<Vec4T<T>> spheres(g.spheres.size());
    for (size_t i = 0; i < g.spheres.size(); ++i) spheres[i] = Vec4T<T>{(T)g.spheres[i].cx, (T)g.spheres[i].cy, (T)g.spheres[i].cz, (T)g.spheres[i].r};
    std::vector<GPlane<T>> plane_geo(g.planes.size());
    for (size_t i = 0; i < g.planes.size(); ++i) {
        const rtw_plane& q = g.planes[i];
        double len = std::sqrt(q.nx * q.nx + q.ny * q.ny + q.nz * q.nz);       // Plane::new normalises (plane.rs:35)
        GPlane<T> pl{};
        pl.point = V3<T>{(T)q.px, (T)q.py, (T)q.pz};
        const host::D3 n{q.nx / len, q.ny / len, q.nz / len}, up{0., 1., 0.};
        pl.normal = v3(n);
        // get_plane_uv (plane.rs:41-47): theta = atan2(|n x V|, n . V) with V = +y, k = (n x V).normalize()
        const host::D3 c = host::cross3(n, up);
        const double clen = std::sqrt(c.x * c.x + c.y * c.y + c.z * c.z);
        const double theta = std::atan2(clen, host::dot3(n, up));
        pl.rotated = theta <= 2.220446049250313e-16 ? 0u : 1u;
        if (pl.rotated) { pl.k = v3(c / clen); pl.cos_theta = (T)std::cos(theta); pl.sin_theta = (T)std::sin(theta); }
        plane_geo[i] = pl;
    }
    std::vector<GXform<T>> xforms(g.transforms.size());
    std::vector<char> invertible(g.transforms.size(), 0);
    for (size_t i = 0; i < g.transforms.size(); ++i) {
        rtw_transform inv{};
        invertible[i] = host::transform_inverse(g.transforms[i], &inv) ? 1 : 0;
        GXform<T> x{};
        for (int k = 0; k < 9; ++k) { x.fwd[k] = (T)g.transforms[i].rotation[k]; x.inv[k] = (T)inv.rotation[k]; }
        for (int k = 0; k < 3; ++k) { x.ft[k] = (T)g.transforms[i].translation[k]; x.it[k] = (T)inv.translation[k]; }
        xforms[i] = x;
    }
    std::vector<GMat<T>> mats(g.materials.size());
    for (size_t i = 0; i < g.materials.size(); ++i) {
        const rtw_material& m = g.materials[i];
        GMat<T> o{};
        o.albedo[0] = (T)m.r; o.albedo[1] = (T)m.g; o.albedo[2] = (T)m.b; o.param = (T)m.param; o.kind = m.kind; o.texture = m.texture;
        mats[i] = o;
    }
    std::vector<GTex<T>> textures(g.textures.size());
    for (size_t i = 0; i < g.textures.size(); ++i) {
        const rtw_texture& t = g.textures[i];
        GTex<T> o{};
        o.kind = t.kind; o.perlin = t.perlin; o.even = t.even; o.odd = t.odd; o.scale = (T)t.scale;
        for (int a = 0; a < 3; ++a) { o.even_c[a] = (T)t.even_colour[a]; o.odd_c[a] = (T)t.odd_colour[a]; }
        textures[i] = o;
    }
    std::vector<GPerlin<T>> perlins(g.perlins.size());
    for (size_t i = 0; i < g.perlins.size(); ++i) {
        for (int k = 0; k < 256; ++k) for (int a = 0; a < 3; ++a) perlins[i].rand_vec[k][a] = (T)g.perlins[i].rand_vec[k][a];
        std::memcpy(perlins[i].perm_x, g.perlins[i].perm_x, 256); std::memcpy(perlins[i].perm_y, g.perlins[i].perm_y, 256);
        std::memcpy(perlins[i].perm_z, g.perlins[i].perm_z, 256);
    }
    std::vector<GQuad<T>> quads;
    auto push_quad = [&](const host::QuadH& q) {
        GQuad<T> o{v3(q.q), v3(q.u), v3(q.v), v3(q.w), v3(q.normal), (T)q.area};
        quads.push_back(o);
    };
    // one list entry -> device record + its world-space box (Bounded::get_aabbox)
    auto make_entry = [&](const rtw_prim& e, uint32_t id, GPrim<T>* out, host::Box* box, bool* bounded) {
        GPrim<T> p{};
        p.kind = e.kind; p.mat = e.material; p.id = id; p.xform = e.transform;
        host::Box ib{};
        *bounded = true;
        switch (e.kind) {
        case RTW_PRIM_SPHERE: {
            const rtw_sphere& q = g.spheres[e.index];
            p.first = e.index;
            ib.mn[0] = q.cx - q.r; ib.mn[1] = q.cy - q.r; ib.mn[2] = q.cz - q.r; ib.mx[0] = q.cx + q.r; ib.mx[1] = q.cy + q.r; ib.mx[2] = q.cz + q.r;
            break;
        }
        case RTW_PRIM_PLANE: {                                          // Plane::get_aabbox (plane.rs:78-107)
            p.first = e.index; *bounded = false;
            const rtw_plane& q = g.planes[e.index];
            double len = std::sqrt(q.nx * q.nx + q.ny * q.ny + q.nz * q.nz);
            double n[3] = {q.nx / len, q.ny / len, q.nz / len};
            const double eps = 2.220446049250313e-16, inf = std::numeric_limits<double>::infinity();
            for (int a = 0; a < 3; ++a) {
                bool flat = std::fabs(n[(a + 1) % 3]) < eps && std::fabs(n[(a + 2) % 3]) < eps;
                ib.mn[a] = flat ? 0. : -inf; ib.mx[a] = flat ? 0. : inf;
            }
            break;
        }
        case RTW_PRIM_CUBOID: {
            host::QuadH f[6];
            host::make_cuboid(host::ld3(g.cuboids[e.index].p), host::ld3(g.cuboids[e.index].q), f, &ib);
            p.first = (uint32_t)quads.size();
            for (auto& q : f) push_quad(q);
            break;
        }
        default: {
            const rtw_quad& q = g.quads[e.index];
            host::QuadH h = host::make_quad(host::ld3(q.q), host::ld3(q.u), host::ld3(q.v), e.kind == RTW_PRIM_TRIANGLE);
            p.first = (uint32_t)quads.size();
            push_quad(h);
            ib = h.box;
        }
        }
        if (*bounded && e.transform >= 0) ib = host::transformed_box(ib, g.transforms[e.transform]);
        *box = ib;
        for (int a = 0; a < 3; ++a) { p.box[a] = (T)ib.mn[a]; p.box[3 + a] = (T)ib.mx[a]; }
        *out = p;
    };
    std::vector<GPrim<T>> bounded, unbounded;
    std::vector<host::Box> boxes;
    for (size_t i = 0; i < g.world.size(); ++i) {
        const rtw_prim& e = g.world[i];
        GPrim<T> p; host::Box b; bool bd;
        make_entry(e, (uint32_t)i, &p, &b, &bd);
        if (e.transform >= 0 && !invertible[e.transform]) continue;     // "if there's no inverse just say it's not hit" (entities/transformations.rs:15-16)
        if (bd) { bounded.push_back(p); boxes.push_back(b); } else unbounded.push_back(p);
    }
    host::Builder builder;
    s->bvh = builder.build_boxes(boxes.data(), boxes.size(), 2, kMaxTreeDepth);
    std::vector<Node<T>> nodes(s->bvh.nodes.size());
    for (size_t i = 0; i < nodes.size(); ++i) {
        const host::FlatNode& f = s->bvh.nodes[i];
        Node<T> n{};
        fill_node<T>(n, f);
        n.left = f.left >= 0 ? f.left : encode_leaf(f.lfirst, f.lcount);
        n.right = f.right >= 0 ? f.right : encode_leaf(f.rfirst, f.rcount);
        nodes[i] = n;
    }
    std::vector<GPrim<T>> prims(bounded.size());
    const bool flat = bounded.size() <= kFlatListLimit;
    if (flat) {
        prims = bounded;
        std::stable_sort(prims.begin(), prims.end(), [](const GPrim<T>& a, const GPrim<T>& b) {
            return std::make_pair(a.kind, a.xform >= 0) < std::make_pair(b.kind, b.xform >= 0);
        });
    } else {
        for (size_t k = 0; k < bounded.size(); ++k) prims[k] = bounded[s->bvh.order[k]];
    }
    std::vector<size_t> lorder(g.lights.size());
    for (size_t i = 0; i < lorder.size(); ++i) lorder[i] = i;
    std::stable_sort(lorder.begin(), lorder.end(), [&](size_t a, size_t b) { return bucket_of(g.lights[a]) < bucket_of(g.lights[b]); });
    std::vector<GPrim<T>> lights(g.lights.size());
    for (size_t k = 0; k < lorder.size(); ++k) {
        host::Box b; bool bd;
        make_entry(g.lights[lorder[k]], (uint32_t)lorder[k], &lights[k], &b, &bd);
    }
    const int32_t n_lights = (int32_t)lights.size();
    if (lights.empty()) { GPrim<T> none{}; none.kind = P_NO_LIGHTS; none.xform = -1; lights.push_back(none); }    // see g_lights_random
    CU(d.nodes.upload(nodes)); CU(d.prims.upload(prims)); CU(d.unbounded.upload(unbounded)); CU(d.lights.upload(lights));
    CU(d.spheres.upload(spheres)); CU(d.plane_geo.upload(plane_geo)); CU(d.quads.upload(quads)); CU(d.xforms.upload(xforms));
    CU(d.mats.upload(mats)); CU(d.textures.upload(textures)); CU(d.perlins.upload(perlins));
    d.view.nodes = d.nodes.p; d.view.prims = d.prims.p; d.view.unbounded = d.unbounded.p; d.view.lights = d.lights.p;
    d.view.spheres = d.spheres.p; d.view.plane_geo = d.plane_geo.p; d.view.quads = d.quads.p; d.view.xforms = d.xforms.p;
    d.view.mats = d.mats.p; d.view.textures = d.textures.p; d.view.perlins = d.perlins.p;
    d.view.n_nodes = (int32_t)nodes.size(); d.view.n_prims = (int32_t)prims.size(); d.view.n_unbounded = (int32_t)unbounded.size();
    d.view.n_lights = n_lights; d.view.lights_is_bvh = g.lights_is_bvh ? 1u : 0u;
    d.view.flat = flat ? 1u : 0u;
    d.view.panic_flag = s->d_panic.p;
    d.view.has_xforms = 0;
    for (const GPrim<T>& p : prims) if (p.xform >= 0) d.view.has_xforms = 1;
    return RTW_OK;
}

// ---- device-side BVH construction (bvh_device.cu) ---------------------------------------------------------------
int g_bvh_builder = RTW_BVH_AUTO;
constexpr size_t kDeviceBuildThreshold = 200000;    // RTW_BVH_AUTO: spheres from which the LBVH's build time wins over the SAH tree's quality

// returns RTW_OK with *built = false when the tree came out deeper than the traversal stack allows (caller falls back to the host builder)
int build_world_on_device(rtw_scene* s, int max_leaf, bool* built) {
    *built = false;
    const size_t ns = s->spheres.size(), np = s->planes.size();
    std::vector<double> mats(4 * ns);
    std::vector<uint32_t> info(ns);
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (size_t k = 0; k < ns; ++k) {
        const rtw_sphere& q = s->spheres[k];
        const rtw_material& m = s->materials[s->sphere_material[k]];
        mats[4 * k] = m.r; mats[4 * k + 1] = m.g; mats[4 * k + 2] = m.b; mats[4 * k + 3] = m.param;
        info[k] = ((uint32_t)(np + k) << 2) | (m.kind & 3u);
        const double c[3] = {q.cx, q.cy, q.cz};
        for (int a = 0; a < 3; ++a) { lo[a] = std::fmin(lo[a], c[a]); hi[a] = std::fmax(hi[a], c[a]); }
    }
    DevBuf<double> d_sph, d_mat; DevBuf<uint32_t> d_info;
    auto cleanup = [&]() { d_sph.release(); d_mat.release(); d_info.release(); };
    cudaError_t e = d_sph.reserve(4 * ns);
    if (e == cudaSuccess) e = cudaMemcpy(d_sph.p, s->spheres.data(), ns * sizeof(rtw_sphere), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = d_mat.upload(mats);
    if (e == cudaSuccess) e = d_info.upload(info);
    if (e == cudaSuccess) e = s->f64.nodes.reserve(ns - 1);
    if (e == cudaSuccess) e = s->f32.nodes.reserve(ns - 1);
    if (e == cudaSuccess) e = s->f64.spheres.reserve(ns);
    if (e == cudaSuccess) e = s->f64.sphere_mat.reserve(ns);
    if (e == cudaSuccess) e = s->f32.spheres.reserve(ns);
    if (e == cudaSuccess) e = s->f32.sphere_mat.reserve(ns);
    if (e == cudaSuccess) e = s->f32.info.reserve(ns + 128);       // the staging copy reads the info section in 16-byte units
    if (e == cudaSuccess) e = s->f64.info.reserve(ns + 128);
    if (e == cudaSuccess)
        e = build_lbvh_device(d_sph.p, d_mat.p, d_info.p, ns, lo, hi, max_leaf, s->f64.nodes.p, s->f32.nodes.p, s->f64.spheres.p, s->f64.sphere_mat.p,
                              s->f32.spheres.p, s->f32.sphere_mat.p, s->f32.info.p, &s->device_bvh, 0);
    if (e == cudaSuccess) e = cudaMemcpy(s->f64.info.p, s->f32.info.p, ns * sizeof(uint32_t), cudaMemcpyDeviceToDevice);
    cleanup();
    if (e != cudaSuccess) { cudaGetLastError(); return fail(e == cudaErrorMemoryAllocation ? RTW_E_NOMEM : RTW_E_CUDA, std::string("device BVH build: ") + cudaGetErrorString(e)); }
    if (s->device_bvh.depth + 2 > (uint32_t)kStackDepth) return RTW_OK;      // too deep for the traversal stack
    s->bvh = host::Bvh();
    s->bvh.depth = s->device_bvh.depth; s->bvh.leaves = s->device_bvh.leaves; s->bvh.max_leaf = (uint32_t)max_leaf;
    s->bvh_builder = RTW_BVH_DEVICE_LBVH;
    *built = true;
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

// sample partition: this launch's samples of every pixel.  own_world > 1 (rtw_render_multi / rtw_render_rank): GPU own_rank of own_world
// renders its share of ONE frame — ALL samples of the pixels it owns when the work queue is ordered per frame (chunk_order_kernel deals
// the chunks out), else its share of every pixel's samples (render_device_impl decides and rewrites the range)
struct SampleRange { uint32_t begin = 0, count = 0; bool set = false; uint32_t own_rank = 0, own_world = 1; };

template <class T, class DEV, class Launch>
int render_tiles_t(rtw_scene* s, DEV& d, const rtw_camera* cam, const rtw_opts* o, uint32_t rank, uint32_t world, T* tiles,
                   cudaStream_t stream, Launch launch, SampleRange sr = SampleRange()) {
    RenderParams<T, decltype(d.view)> P{};
    P.scene = d.view; P.cam = to_camera<T>(cam); P.seed = o->seed; P.tmin = resolve_tmin<T>(o->tmin); P.flags = o->flags;
    if (sr.set) { P.cam.spp = sr.count; P.cam.sample_offset = sr.begin; }
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
template <class T, class DEV> BatchParams<T, decltype(DEV::view)> batch_params(rtw_scene* s, DEV& d, size_t n) {
    BatchParams<T, decltype(DEV::view)> P{};
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
// general scenes: did a path do what makes the reference panic?  (call after the stream has been synchronised)
int check_reference_panic(rtw_scene* s) {
    if (!s->general) return RTW_OK;
    uint32_t flag = 0;
    CU(cudaMemcpy(&flag, s->d_panic.p, sizeof(flag), cudaMemcpyDeviceToHost));
    if (!flag) return RTW_OK;
    CU(cudaMemset(s->d_panic.p, 0, sizeof(flag)));
    return fail(RTW_E_INVALID, "a path sampled an empty lights list: the reference panics here (HittableList shouldn't be empty, hittable_list.rs:414-419)");
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
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess) { cudaGetLastError(); cur = -1; }
    for (auto& kv : c.blocks) { cudaSetDevice(kv.first.first); cudaFree(kv.second); }
    if (cur >= 0) cudaSetDevice(cur);
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
        // Plane::get_aabbox (plane.rs:78-107) is the slab {axis = 0} for an axis-aligned normal whatever the plane's offset, and
        // bounded_hit tests it first: for a plane through the origin that equals Plane::hit's own range test (this path), for
        // an offset one it hides most of the plane (the general path applies the box)
        {
            const rtw_plane& q = planes[i];
            double len = std::sqrt(q.nx * q.nx + q.ny * q.ny + q.nz * q.nz);
            double n[3] = {q.nx / len, q.ny / len, q.nz / len}, pt[3] = {q.px, q.py, q.pz};
            for (int a = 0; a < 3; ++a)
                if (std::fabs(n[(a + 1) % 3]) < 2.220446049250313e-16 && std::fabs(n[(a + 2) % 3]) < 2.220446049250313e-16 && pt[a] != 0.)
                    return fail(RTW_E_UNSUPPORTED, "axis-aligned plane that does not pass through the origin: use rtw_scene_create_general");
        }
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
    const bool want_device = g_bvh_builder == RTW_BVH_DEVICE_LBVH || (g_bvh_builder == RTW_BVH_AUTO && n_spheres >= kDeviceBuildThreshold);
    auto bail = [&](int code) { rtw_scene_destroy(s); return code; };
    cudaError_t e = cudaGetDevice(&s->device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&s->sm_count, cudaDevAttrMultiProcessorCount, s->device);
    if (e == cudaSuccess) e = cached_malloc(reinterpret_cast<void**>(&s->d_work), sizeof(unsigned int), nullptr);
    if (e == cudaSuccess) e = cached_malloc(reinterpret_cast<void**>(&s->d_counters), sizeof(DeviceCounters), nullptr);
    for (int i = 0; i < 4 && e == cudaSuccess; ++i) e = cudaEventCreate(&s->ev[i]);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->side_stream, cudaStreamNonBlocking);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&s->ev_side[i], cudaEventDisableTiming);
    if (e != cudaSuccess) { fail(RTW_E_CUDA, cudaGetErrorString(e)); return bail(RTW_E_CUDA); }
    bool on_device = false;
    int rc = RTW_OK;
    if (want_device && n_spheres >= 2) rc = build_world_on_device(s, max_leaf, &on_device);
    if (rc != RTW_OK) return bail(rc);
    if (!on_device) s->bvh = builder.build(reinterpret_cast<const double*>(s->spheres.data()), n_spheres, max_leaf, kMaxTreeDepth);
    rc = upload_scene<float>(s, s->f32, on_device);
    if (rc == RTW_OK) rc = upload_scene<double>(s, s->f64, on_device);
    if (rc != RTW_OK) return bail(rc);
    *out = s;
    return RTW_OK;
}

int rtw_set_bvh_builder(int mode) {
    if (mode != RTW_BVH_AUTO && mode != RTW_BVH_HOST_SAH && mode != RTW_BVH_DEVICE_LBVH) return fail(RTW_E_INVALID, "bvh builder mode");
    g_bvh_builder = mode;
    return RTW_OK;
}
int rtw_scene_bvh_builder(const rtw_scene* s) { return s ? s->bvh_builder : fail(RTW_E_INVALID, "scene is NULL"); }

void rtw_transform_then(const rtw_transform* a, const rtw_transform* b, rtw_transform* out) { host::transform_then(*a, *b, out); }
int rtw_transform_inverse(const rtw_transform* a, rtw_transform* out) { return host::transform_inverse(*a, out) ? 1 : 0; }
void rtw_rotation(double angle_degrees, int axis, rtw_transform* out) { host::make_rotation(angle_degrees, axis, out); }
void rtw_perlin_generate(uint64_t seed, uint32_t index, rtw_perlin* out) { host::perlin_generate(seed, index, out); }

int rtw_scene_create_general(const rtw_scene_desc* d, rtw_scene** out) {
    if (!out) return fail(RTW_E_INVALID, "out is NULL");
    *out = nullptr;
    if (!d) return fail(RTW_E_INVALID, "desc is NULL");
    if ((d->n_spheres && !d->spheres) || (d->n_planes && !d->planes) || (d->n_quads && !d->quads) || (d->n_cuboids && !d->cuboids) ||
        (d->n_transforms && !d->transforms) || (d->n_materials && !d->materials) || (d->n_textures && !d->textures) ||
        (d->n_perlins && !d->perlins) || (d->n_world && !d->world) || (d->n_lights && !d->lights))
        return fail(RTW_E_INVALID, "NULL array with non-zero count");
    if (d->n_world >= (1u << 28) || d->n_lights >= (1u << 28)) return fail(RTW_E_UNSUPPORTED, "too many primitives");
    for (uint64_t i = 0; i < d->n_materials; ++i) {
        const rtw_material& m = d->materials[i];
        if (m.kind > RTW_ISOTROPIC) return fail(RTW_E_UNSUPPORTED, "unknown material kind");
        if (m.texture > d->n_textures) return fail(RTW_E_INVALID, "material texture index out of range");
        if (m.texture && m.kind != RTW_LAMBERTIAN && m.kind != RTW_DIFFUSE_LIGHT && m.kind != RTW_ISOTROPIC)
            return fail(RTW_E_INVALID, "only Lambertian, DiffuseLight and Isotropic carry a texture");
    }
    for (uint64_t i = 0; i < d->n_textures; ++i) {
        const rtw_texture& t = d->textures[i];
        if (t.kind == RTW_TEX_NOISE) {
            if (t.perlin >= d->n_perlins) return fail(RTW_E_INVALID, "texture perlin index out of range");
        } else if (t.kind == RTW_TEX_CHECKER) {
            for (uint32_t ref : {t.even, t.odd}) {
                if (ref > d->n_textures) return fail(RTW_E_INVALID, "checker sub-texture index out of range");
                // even / odd may be any texture, a CheckerTexture included (texture.rs:26-29); sub-textures precede their parent in
                // the table, which rules out cycles (the reference's Arc tree cannot have them either)
                if (ref && d->textures[ref - 1].kind == RTW_TEX_CHECKER && ref - 1 >= i)
                    return fail(RTW_E_INVALID, "a CheckerTexture's even / odd CheckerTexture must come earlier in the texture table");
            }
            if (!(t.scale != 0.)) return fail(RTW_E_INVALID, "checker scale is zero");
        } else return fail(RTW_E_UNSUPPORTED, "texture kind (NoiseTexture, CheckerTexture; SolidColour is texture 0)");
    }
    bool needs_lights = false;
    auto check = [&](const rtw_prim* list, uint64_t n, bool world) -> const char* {
        for (uint64_t i = 0; i < n; ++i) {
            const rtw_prim& e = list[i];
            uint64_t limit = e.kind == RTW_PRIM_SPHERE ? d->n_spheres : e.kind == RTW_PRIM_PLANE ? d->n_planes
                             : (e.kind == RTW_PRIM_QUAD || e.kind == RTW_PRIM_TRIANGLE) ? d->n_quads : e.kind == RTW_PRIM_CUBOID ? d->n_cuboids : 0;
            if (e.kind > RTW_PRIM_CUBOID) return "unknown primitive kind";
            if (e.index >= limit) return "primitive index out of range";
            if (e.material >= d->n_materials) return "primitive material index out of range";
            if (e.transform >= 0 && (uint64_t)e.transform >= d->n_transforms) return "primitive transform index out of range";
            if (e.transform < -1) return "primitive transform index out of range";
            if (e.kind == RTW_PRIM_PLANE && e.transform >= 0) return "transformed planes are not supported";
            if (e.kind == RTW_PRIM_PLANE && d->materials[e.material].texture && d->textures[d->materials[e.material].texture - 1].kind == RTW_TEX_CHECKER) {
                // get_plane_uv (plane.rs:41-55) rotates about k = (n x +y).normalize(): for a normal of exactly -y that is 0 / 0, (u, v)
                // is NaN and Plane::hit panics on the finiteness check (plane.rs:67-69)
                const rtw_plane& q = d->planes[e.index];
                if (q.nx == 0. && q.nz == 0. && q.ny < 0.) return "a CheckerTexture on a plane whose normal is -y: the reference panics (get_plane_uv is NaN)";
            }
            if (e.kind == RTW_PRIM_SPHERE) {
                const rtw_sphere& q = d->spheres[e.index];
                if (!(q.r > 0.) || !std::isfinite(q.cx + q.cy + q.cz + q.r)) return "sphere with non-finite centre or non-positive radius";
            }
            if (world && (d->materials[e.material].kind == RTW_LAMBERTIAN || d->materials[e.material].kind == RTW_ISOTROPIC)) needs_lights = true;
        }
        return nullptr;
    };
    if (const char* msg = check(d->world, d->n_world, true)) return fail(std::strstr(msg, "not supported") ? RTW_E_UNSUPPORTED : RTW_E_INVALID, msg);
    if (const char* msg = check(d->lights, d->n_lights, false)) return fail(std::strstr(msg, "not supported") ? RTW_E_UNSUPPORTED : RTW_E_INVALID, msg);
    (void)needs_lights;     // an empty lights list is accepted: the reference only panics when a path actually samples it (checked per render)
    if (d->lights_is_bvh && d->n_lights > 5)
        return fail(RTW_E_UNSUPPORTED, "a BoundedVolumeHierarchy of more than 5 lights (the reference's aux_random indexes out of range, bvh.rs:78-93)");
    int ndev = rtw_device_count();
    if (ndev < 0) return ndev;
    if (ndev == 0) return fail(RTW_E_NO_DEVICE, "no CUDA device: this backend has no CPU fallback");

    rtw_scene* s = new rtw_scene();
    s->general = true;
    GeneralDesc& g = s->gdesc;
    g.spheres.assign(d->spheres, d->spheres + d->n_spheres); g.planes.assign(d->planes, d->planes + d->n_planes);
    g.quads.assign(d->quads, d->quads + d->n_quads); g.cuboids.assign(d->cuboids, d->cuboids + d->n_cuboids);
    g.transforms.assign(d->transforms, d->transforms + d->n_transforms); g.materials.assign(d->materials, d->materials + d->n_materials);
    g.textures.assign(d->textures, d->textures + d->n_textures); g.perlins.assign(d->perlins, d->perlins + d->n_perlins);
    g.world.assign(d->world, d->world + d->n_world); g.lights.assign(d->lights, d->lights + d->n_lights);
    g.lights_is_bvh = d->lights_is_bvh != 0;
    auto bail = [&](int code) { rtw_scene_destroy(s); return code; };
    cudaError_t e = cudaGetDevice(&s->device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&s->sm_count, cudaDevAttrMultiProcessorCount, s->device);
    if (e == cudaSuccess) e = cached_malloc(reinterpret_cast<void**>(&s->d_work), sizeof(unsigned int), nullptr);
    if (e == cudaSuccess) e = cached_malloc(reinterpret_cast<void**>(&s->d_counters), sizeof(DeviceCounters), nullptr);
    for (int i = 0; i < 4 && e == cudaSuccess; ++i) e = cudaEventCreate(&s->ev[i]);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->side_stream, cudaStreamNonBlocking);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&s->ev_side[i], cudaEventDisableTiming);
    if (e != cudaSuccess) { fail(RTW_E_CUDA, cudaGetErrorString(e)); return bail(RTW_E_CUDA); }
    e = s->d_panic.reserve(1);
    if (e == cudaSuccess) e = cudaMemset(s->d_panic.p, 0, sizeof(uint32_t));
    if (e != cudaSuccess) { fail(RTW_E_CUDA, cudaGetErrorString(e)); return bail(RTW_E_CUDA); }
    int rc = upload_general<float>(s, s->g32);
    if (rc == RTW_OK) rc = upload_general<double>(s, s->g64);
    if (rc != RTW_OK) return bail(rc);
    *out = s;
    return RTW_OK;
}

void rtw_scene_destroy(rtw_scene* s) {
    if (!s) return;
    multi_release(s);
    s->d_block.release(); s->d_gather.release(); s->d_tiles_rank.release();
    s->f32.release(); s->f64.release(); s->g32.release(); s->g64.release(); s->d_panic.release();
    cached_free(s->d_work, sizeof(unsigned int), s->device);
    cached_free(s->d_counters, sizeof(DeviceCounters), s->device);
    s->d_rgb_sum.release(); s->d_rgb8.release(); s->d_accum.release(); s->d_poison.release(); s->d_cand.release(); s->d_cand_blocks.release(); s->d_order.release();
    s->d_in2.release(); s->d_in3.release(); s->d_u3.release(); s->d_u4.release();
    s->d_in0.release(); s->d_in1.release(); s->d_out0.release(); s->d_out1.release(); s->d_out2.release(); s->d_out3.release();
    s->d_out4.release(); s->d_u0.release(); s->d_u1.release(); s->d_u2.release(); s->d_k.release(); s->d_prim.release();
    for (auto& ev : s->ev) if (ev) cudaEventDestroy(ev);
    for (auto& ev : s->ev_side) if (ev) cudaEventDestroy(ev);
    if (s->side_stream) cudaStreamDestroy(s->side_stream);
    delete s;
}

int rtw_scene_info(const rtw_scene* s, uint64_t out[5]) {
    if (!s || !out) return fail(RTW_E_INVALID, "NULL argument");
    out[0] = s->bvh_builder == RTW_BVH_DEVICE_LBVH ? s->device_bvh.inner_nodes : s->bvh.nodes.size(); out[1] = s->bvh.leaves; out[2] = s->bvh.depth; out[3] = s->bvh.max_leaf;
    out[4] = s->f32.bytes() + s->f64.bytes() + s->g32.bytes() + s->g64.bytes();
    return RTW_OK;
}

static_assert(sizeof(rtw_bvh_node) == 72, "rtw_bvh_node layout (include/rtw.h, api.py BVH_NODE_DTYPE, rust/cuda RtwBvhNode)");
// The reference's flat-tree sketch (bvh.rs:224-241) as the host mirror format: Root / Node / Leaf records with parent links.
int rtw_scene_export_bvh(rtw_scene* s, rtw_bvh_node* out, size_t node_cap, size_t* n_nodes, uint32_t* prim_order, size_t prim_cap, size_t* n_prims) {
    if (!s || !n_nodes || !n_prims) return fail(RTW_E_INVALID, "NULL argument");
    cudaError_t e = cudaSetDevice(s->device);
    if (e != cudaSuccess) return fail(RTW_E_CUDA, cudaGetErrorString(e));
    // leaf order of the primitive ids, and the device nodes (f64: the exact path's (min, max) boxes)
    std::vector<uint32_t> order;
    std::vector<Node<double>> dn;
    bool flat = false;
    if (s->general) {
        std::vector<GPrim<double>> prims(s->g64.prims.n);
        if (!prims.empty()) e = cudaMemcpy(prims.data(), s->g64.prims.p, prims.size() * sizeof(GPrim<double>), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) return fail(RTW_E_CUDA, cudaGetErrorString(e));
        order.resize(prims.size());
        for (size_t k = 0; k < prims.size(); ++k) order[k] = prims[k].id;
        flat = s->g64.view.flat != 0;
        if (!flat) dn.resize((size_t)s->g64.view.n_nodes);
        if (!dn.empty()) e = cudaMemcpy(dn.data(), s->g64.nodes.p, dn.size() * sizeof(Node<double>), cudaMemcpyDeviceToHost);
    } else {
        order.resize((size_t)s->f64.view.n_spheres);
        if (!order.empty()) e = cudaMemcpy(order.data(), s->f64.info.p, order.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) return fail(RTW_E_CUDA, cudaGetErrorString(e));
        for (uint32_t& v : order) v >>= 2;                     // info = prim_id << 2 | material kind
        dn.resize((size_t)s->f64.view.n_nodes);
        if (!dn.empty()) e = cudaMemcpy(dn.data(), s->f64.nodes.p, dn.size() * sizeof(Node<double>), cudaMemcpyDeviceToHost);
    }
    if (e != cudaSuccess) return fail(RTW_E_CUDA, cudaGetErrorString(e));
    // breadth-first walk from the root; every child link of a device node becomes a record of its own
    std::vector<rtw_bvh_node> nodes;
    if (!dn.empty() && !order.empty()) {
        struct Item { int32_t link; int32_t self; };
        auto make = [&](const double* mn, const double* mx, int32_t parent, uint32_t depth) {
            rtw_bvh_node n{};
            for (int a = 0; a < 3; ++a) { n.box_min[a] = mn[a]; n.box_max[a] = mx[a]; }
            n.parent = parent; n.left = n.right = -1; n.depth = depth;
            nodes.push_back(n);
            return (int32_t)nodes.size() - 1;
        };
        const Node<double>& r = dn[0];
        double mn[3], mx[3];
        const bool l_ok = r.left != kEmptyLeaf, r_ok = r.right != kEmptyLeaf;
        for (int a = 0; a < 3; ++a) {
            mn[a] = l_ok && r_ok ? std::fmin(r.la[a], r.ra[a]) : (l_ok ? r.la[a] : r.ra[a]);
            mx[a] = l_ok && r_ok ? std::fmax(r.lb[a], r.rb[a]) : (l_ok ? r.lb[a] : r.rb[a]);
        }
        std::vector<Item> queue{{0, make(mn, mx, -1, 0)}};
        for (size_t head = 0; head < queue.size(); ++head) {
            const Item it = queue[head];
            if ((size_t)it.link >= dn.size()) return fail(RTW_E_CUDA, "corrupt BVH: child index out of range");
            const Node<double> nd = dn[(size_t)it.link];
            const int32_t links[2] = {nd.left, nd.right};
            const double* boxes[2][2] = {{nd.la, nd.lb}, {nd.ra, nd.rb}};
            for (int c = 0; c < 2; ++c) {
                if (links[c] == kEmptyLeaf) continue;
                const int32_t child = make(boxes[c][0], boxes[c][1], it.self, nodes[(size_t)it.self].depth + 1);
                (c == 0 ? nodes[(size_t)it.self].left : nodes[(size_t)it.self].right) = child;
                if (links[c] >= 0) queue.push_back({links[c], child});
                else {
                    const uint32_t enc = (uint32_t)~links[c];
                    nodes[(size_t)child].first = enc >> 4; nodes[(size_t)child].count = (enc & 15u) + 1u;
                }
            }
        }
    }
    *n_nodes = nodes.size(); *n_prims = order.size();
    if (out) {
        if (node_cap < nodes.size()) return fail(RTW_E_INVALID, "node_capacity is smaller than the tree");
        std::copy(nodes.begin(), nodes.end(), out);
    }
    if (prim_order) {
        if (prim_cap < order.size()) return fail(RTW_E_INVALID, "prim_capacity is smaller than the number of bounded entries");
        std::copy(order.begin(), order.end(), prim_order);
    }
    return RTW_OK;
}

}  // extern "C"

namespace {
// rtw_render_tiles_device (tile partition: ext_accum == nullptr) and rtw_render_samples_device (sample partition: the caller's
// fixed-point accumulators, every pixel, samples [sr.begin, sr.begin + sr.count))
int render_device_impl(rtw_scene* s, const rtw_camera* cam, const rtw_opts* o, uint32_t rank, uint32_t world, void* d_tiles,
                       void* stream, rtw_stats* stats, SampleRange sr, unsigned long long* ext_accum, uint32_t* ext_poison) {
    if (!s || (!d_tiles && !ext_accum)) return fail(RTW_E_INVALID, "NULL argument");
    int rc = check_camera(cam); if (rc) return rc;
    rc = check_opts(o); if (rc) return rc;
    if (world == 0 || rank >= world) return fail(RTW_E_INVALID, "rank/world");
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t launches = 1;
    // RTW_F64 has one renderer (lane per pixel, samples summed in order); `mode` only selects among the FP32 renderers
    bool pooled = o->precision == RTW_F32 && (o->mode == RTW_WAVEFRONT || !(o->flags & RTW_FLAG_LANE_PER_PIXEL));
    if (s->general && (o->flags & RTW_FLAG_LANE_PER_PIXEL)) pooled = false;      // `mode` does not apply to general scenes
    if (ext_accum && !pooled) return fail(RTW_E_UNSUPPORTED, "the sample partition needs a fixed-point FP32 renderer (RTW_F32 without RTW_FLAG_LANE_PER_PIXEL)");
    if (pooled && !ext_accum) {
        size_t n_slots = (size_t)rtw_tiles_per_rank(cam->image_width, cam->image_height, world) * kTileW * kTileH;
        CU(s->d_accum.reserve(n_slots * 3)); CU(s->d_poison.reserve(n_slots));
    }
    unsigned long long* accum_p = ext_accum ? ext_accum : s->d_accum.p;
    uint32_t* poison_p = ext_accum ? ext_poison : s->d_poison.p;
    {
        // diagnostic (scripts/timeline_probe.py, variant_bench.py): RTW_DEBUG_OWN="r,n" makes a plain single-GPU render behave like GPU r of n
        // of rtw_render_multi — it renders only the chunks that GPU would own (the image is partial; the timing is that GPU's)
        static const int dbg_own[2] = {[] { const char* e = std::getenv("RTW_DEBUG_OWN"); return e ? std::atoi(e) : 0; }(),
                                       [] { const char* e = std::getenv("RTW_DEBUG_OWN"); const char* c = e ? std::strchr(e, ',') : nullptr; return c ? std::atoi(c + 1) : 1; }()};
        if (dbg_own[1] > 1 && sr.own_world == 1 && !sr.set && world == 1) { sr.own_rank = (uint32_t)dbg_own[0] % (uint32_t)dbg_own[1]; sr.own_world = (uint32_t)dbg_own[1]; }
    }
    if (sr.own_world > 1) {
        // one frame over several GPUs: pixels if this call will order its work queue (same conditions as below), samples otherwise
        static const bool pixels_allowed = [] { const char* e = std::getenv("RTW_MULTI_PARTITION"); return !(e && std::string(e) == "samples"); }();
        static const bool cand_ok = [] { const char* e = std::getenv("RTW_NO_PRIMARY_CANDIDATES"); return !(e && std::atoi(e) == 1); }();
        static const bool order_ok = [] { const char* e = std::getenv("RTW_NO_CHUNK_ORDER"); return !(e && std::atoi(e) == 1); }();
        const uint32_t spp_all = cam->samples_per_pixel;
        const uint64_t chunks = ((uint64_t)rtw_tiles_per_rank(cam->image_width, cam->image_height, world) * kTileW * kTileH) / std::max<uint32_t>(1u, pool_pixels_per_chunk(spp_all)) + 1;
        const bool by_pixels = pixels_allowed && cand_ok && order_ok && !s->general && pooled && o->mode == RTW_WAVEFRONT && !(o->flags & RTW_FLAG_NO_CANDIDATES) &&
                               cam->max_depth <= 0xffffu && std::max(s->bvh.depth, s->light_bvh_depth) <= wavefront_max_bvh_depth() &&
                               cam->defocus_angle <= 2.220446049250313e-16 && s->f32.view.n_spheres > 0 && chunks < kChunkMask;
        if (by_pixels) { sr.begin = 0; sr.count = spp_all; sr.set = true; }
        else {
            const uint32_t b = (uint32_t)((uint64_t)spp_all * sr.own_rank / sr.own_world), e = (uint32_t)((uint64_t)spp_all * (sr.own_rank + 1) / sr.own_world);
            sr.begin = b; sr.count = e - b; sr.set = true; sr.own_rank = 0; sr.own_world = 1;
        }
    }
    const uint32_t spp_here = sr.set ? sr.count : cam->samples_per_pixel;
    CU(cudaEventRecord(s->ev[0], st));
    if (s->general) {
        // general scenes: FP32 = pooled path stream (or lane per pixel with RTW_FLAG_LANE_PER_PIXEL), f64 = lane per pixel
        if (pooled) {
            PoolParams Q{};
            Q.accum = accum_p; Q.poison = poison_p;
            Q.pixels_per_chunk = pool_pixels_per_chunk(spp_here);
            Q.sample_cap = pool_sample_cap(cam->samples_per_pixel);
            uint32_t n_slots = rtw_tiles_per_rank(cam->image_width, cam->image_height, world) * kTileW * kTileH;
            Q.n_chunks = (n_slots + Q.pixels_per_chunk - 1) / Q.pixels_per_chunk;
            Q.queue_cap = Q.n_chunks;
            const bool wavefront = o->mode == RTW_WAVEFRONT && cam->max_depth <= 0xffffu && s->bvh.depth + 2 <= 24;
            auto launch = [&](RenderParams<float, SceneViewG<float>> P, bool count, int sms, cudaStream_t str, LaunchInfo* info) {
                return wavefront ? launch_render_wavefront_general_f32(P, Q, s->bvh.depth, count, sms, str, info)
                                 : launch_render_pool_general_f32(P, Q, count, sms, str, info);
            };
            rc = render_tiles_t<float>(s, s->g32, cam, o, rank, world, (float*)d_tiles, st, launch, sr);
            launches = 2;
        } else if (o->precision == RTW_F32) rc = render_tiles_t<float>(s, s->g32, cam, o, rank, world, (float*)d_tiles, st, launch_render_general_f32);
        else rc = render_tiles_t<double>(s, s->g64, cam, o, rank, world, (double*)d_tiles, st, launch_render_general_f64);
    } else if (pooled) {
        PoolParams Q{};
        Q.accum = accum_p; Q.poison = poison_p;
        Q.pixels_per_chunk = pool_pixels_per_chunk(spp_here);
        Q.sample_cap = pool_sample_cap(cam->samples_per_pixel);
        uint32_t n_slots = rtw_tiles_per_rank(cam->image_width, cam->image_height, world) * kTileW * kTileH;
        Q.n_chunks = (n_slots + Q.pixels_per_chunk - 1) / Q.pixels_per_chunk;
        Q.queue_cap = Q.n_chunks;
        // the wavefront packs the remaining depth into 16 bits; deeper paths take the (bit-identical) megakernel
        // ... and a tree too deep for the wavefront's shared-memory stacks (a device-built LBVH can be) does too
        const uint32_t bvh_depth = std::max(s->bvh.depth, s->light_bvh_depth);
        const bool wavefront = o->mode == RTW_WAVEFRONT && cam->max_depth <= 0xffffu && bvh_depth <= wavefront_max_bvh_depth();
        // pinhole camera (get_ray's own predicate, camera.rs:285): the camera rays' closest hits come from per-pixel candidate lists
        // built here, once per frame, by a cone walk of the tree (primary_candidates_kernel)
        static const bool cand_allowed = [] { const char* e = std::getenv("RTW_NO_PRIMARY_CANDIDATES"); return !(e && std::atoi(e) == 1); }();
        uint4* cand = nullptr;
        // (wavefront only: in the pooled megakernel the extra branch of path_step costs more than the walks it saves — measured 56.5 ms
        // with the lists against 53.3 without on C2 / 100 spp; the images are identical either way)
        if (cand_allowed && !(o->flags & RTW_FLAG_NO_CANDIDATES) && wavefront && cam->defocus_angle <= 2.220446049250313e-16 && s->f32.view.n_spheres > 0) {
            CU(s->d_cand.reserve((size_t)cam->image_width * cam->image_height));
            CU(s->d_cand_blocks.reserve(primary_candidates_scratch_bytes(cam->image_width, cam->image_height)));
            CU(launch_primary_candidates_f32(s->f32.view, to_camera<float>(cam), s->d_cand_blocks.p, s->d_cand.p, st));
            launches++;
            cand = s->d_cand.p;
            launches++;
            // ... and they say which chunks of the path stream can meet a sphere: those go first (chunk_order_kernel)
            static const bool order_allowed = [] { const char* e = std::getenv("RTW_NO_CHUNK_ORDER"); return !(e && std::atoi(e) == 1); }();
            if (order_allowed && Q.n_chunks > 0 && Q.n_chunks < kChunkMask) {
                // ... the last costly chunks — two per warp — go out in pieces (kChunkSubs) when a chunk is long enough to be worth cutting
                // (RTW_SPLIT_CHUNKS_PER_WARP=0: none)
                static const long split_per_warp = [] { const char* e = std::getenv("RTW_SPLIT_CHUNKS_PER_WARP"); return e ? std::atol(e) : 2L; }();
                const uint64_t per_chunk = (uint64_t)Q.pixels_per_chunk * (spp_here ? spp_here : 1u);
                const uint64_t warps = 24u * (uint64_t)(s->sm_count > 0 ? s->sm_count : 148);      // (20 per SM without a light BVH: the sizes below are not that fine)
                const uint32_t split_chunks = per_chunk >= 32u * kChunkSubs && split_per_warp > 0
                                                  ? (uint32_t)std::min<uint64_t>((uint64_t)split_per_warp * warps, Q.n_chunks) : 0u;
                const uint32_t cap = Q.n_chunks + chunk_order_extra(split_chunks);
                CU(s->d_order.reserve(chunk_order_words(Q.n_chunks, cap)));
                CU(launch_chunk_order_f32(cand, s->f32.view, to_camera<float>(cam), rank, world, (cam->image_width + kTileW - 1) / kTileW,
                                          rtw_tiles_total(cam->image_width, cam->image_height), n_slots, Q.pixels_per_chunk, Q.n_chunks, cap,
                                          sr.own_rank, sr.own_world, s->d_order.p, st));
                Q.chunk_order = s->d_order.p;
                Q.queue_cap = cap;
                // ... and the background-only chunks beyond a tail of ~4 k paths per warp leave the wavefront's queue for a kernel of their own
                // (RTW_CHEAP_TAIL_PATHS=-1: they all stay in the queue)
                static const long tail_per_warp = [] { const char* e = std::getenv("RTW_CHEAP_TAIL_PATHS"); return e ? std::atol(e) : 4096L; }();
                const uint64_t tail = tail_per_warp < 0 ? 0xffffffffull : ((uint64_t)tail_per_warp * warps + per_chunk - 1) / per_chunk;
                CU(launch_chunk_split_f32(s->d_order.p, Q.n_chunks, cap, (uint32_t)std::min<uint64_t>(tail, 0xffffffffu), split_chunks, st));
                if (tail_per_warp >= 0) Q.queue_len = s->d_order.p + cap + 2 * (size_t)Q.n_chunks + 2;
                launches += 3;
            }
        }
        if (sr.own_world > 1 && !Q.chunk_order) return fail(RTW_E_UNSUPPORTED, "frame too large for the pixel split of a multi-GPU render (set RTW_MULTI_PARTITION=samples)");
        auto launch = [&](RenderParams<float> P, bool count, int sms, cudaStream_t str, LaunchInfo* info) {
            P.cand = cand;
            return wavefront ? launch_render_wavefront_f32(P, Q, bvh_depth, count, sms, str, info)
                             : launch_render_pool_f32(P, Q, count, sms, str, info);
        };
        // RTW_SIDE_STREAM=1: render_background_kernel on a second stream NEXT TO the wavefront kernel (its CTAs fit beside the wavefront's one CTA per
        // SM).  Measured and not the default: the frame's 5.4 ms of background work overlap, but the wavefront kernel loses more than that to the
        // guest in its instruction cache and issue slots — C2 131.0 -> 144.0 ms, one of eight GPUs' share 17.02 -> 18.40 (profiles/r2_side_stream_ab.jsonl)
        static const bool side_allowed = [] { const char* e = std::getenv("RTW_SIDE_STREAM"); return e && std::atoi(e) == 1; }();
        if (side_allowed) set_background_side_stream(s->side_stream, s->ev_side[0], s->ev_side[1]);
        rc = render_tiles_t<float>(s, s->f32, cam, o, rank, world, (float*)d_tiles, st, launch, sr);
        set_background_side_stream(nullptr, nullptr, nullptr);
        launches += 1;
    } else if (o->precision == RTW_F32) rc = render_tiles_t<float>(s, s->f32, cam, o, rank, world, (float*)d_tiles, st, launch_render_f32);
    else rc = render_tiles_t<double>(s, s->f64, cam, o, rank, world, (double*)d_tiles, st, launch_render_f64);
    if (rc) return rc;
    s->last_launches = launches;
    CU(cudaEventRecord(s->ev[1], st));
    if (stats) {
        DeviceCounters c;
        CU(cudaMemcpyAsync(&c, s->d_counters, sizeof(c), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        rc = check_reference_panic(s);
        if (rc) return rc;
        std::memset(stats, 0, sizeof(*stats));
        read_stats(c, stats);
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, s->ev[0], s->ev[1]));
        stats->kernel_ms = ms; stats->total_ms = ms; stats->launches = launches;
    }
    return RTW_OK;
}
}  // namespace

extern "C" {

int rtw_render_tiles_device(rtw_scene* s, const rtw_camera* cam, const rtw_opts* o, uint32_t rank, uint32_t world, void* d_tiles,
                            void* stream, rtw_stats* stats) {
    if (!d_tiles) return fail(RTW_E_INVALID, "NULL argument");
    return render_device_impl(s, cam, o, rank, world, d_tiles, stream, stats, SampleRange(), nullptr, nullptr);
}

int rtw_render_samples_device(rtw_scene* s, const rtw_camera* cam, const rtw_opts* o, uint32_t sample_begin, uint32_t sample_count,
                              void* d_accum, void* d_poison, void* stream, rtw_stats* stats) {
    if (!d_accum || !d_poison) return fail(RTW_E_INVALID, "NULL argument");
    if (cam && (uint64_t)sample_begin + sample_count > cam->samples_per_pixel) return fail(RTW_E_INVALID, "sample range outside the camera's samples_per_pixel");
    SampleRange sr; sr.begin = sample_begin; sr.count = sample_count; sr.set = true;
    return render_device_impl(s, cam, o, 0, 1, nullptr, stream, stats, sr, (unsigned long long*)d_accum, (uint32_t*)d_poison);
}

int rtw_resolve_accum_device(const void* d_accum, const void* d_poison, uint32_t width, uint32_t height, uint32_t spp, double* d_rgb_sum,
                             uint8_t* d_rgb8, void* stream) {
    if (!d_accum || !d_poison || width == 0 || height == 0) return fail(RTW_E_INVALID, "bad argument");
    CU(launch_resolve_accum_f32((const unsigned long long*)d_accum, (const uint32_t*)d_poison, width, height, spp, d_rgb_sum, d_rgb8, (cudaStream_t)stream));
    return RTW_OK;
}

size_t rtw_accum_slots(uint32_t width, uint32_t height) { return (size_t)rtw_tiles_total(width, height) * kTileW * kTileH; }

int rtw_render_samples(rtw_scene* s, const rtw_camera* cam, const rtw_opts* o, uint32_t sample_begin, uint32_t sample_count,
                       uint64_t* accum, uint32_t* poison, rtw_stats* stats) {
    if (!s || !accum || !poison) return fail(RTW_E_INVALID, "NULL argument");
    int rc = check_camera(cam); if (rc) return rc;
    const size_t n = rtw_accum_slots(cam->image_width, cam->image_height);
    CU(s->d_accum.reserve(n * 3)); CU(s->d_poison.reserve(n));
    rtw_stats st{};
    rc = rtw_render_samples_device(s, cam, o, sample_begin, sample_count, s->d_accum.p, s->d_poison.p, nullptr, &st);
    if (rc) return rc;
    std::vector<unsigned long long> a(n * 3);
    std::vector<uint32_t> p(n);
    CU(cudaMemcpy(a.data(), s->d_accum.p, n * 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(p.data(), s->d_poison.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < n * 3; ++i) accum[i] += a[i];          // integer sums commute: the order of the passes does not matter
    for (size_t i = 0; i < n; ++i) poison[i] |= p[i];
    if (stats) *stats = st;
    return RTW_OK;
}

int rtw_resolve_accum(const uint64_t* accum, const uint32_t* poison, uint32_t width, uint32_t height, uint32_t spp, double* rgb_sum,
                      uint8_t* rgb8) {
    if (!accum || !poison || width == 0 || height == 0) return fail(RTW_E_INVALID, "bad argument");
    int ndev = rtw_device_count();
    if (ndev <= 0) return ndev < 0 ? ndev : fail(RTW_E_NO_DEVICE, "no CUDA device: this backend has no CPU fallback");
    const size_t n = rtw_accum_slots(width, height), npx = (size_t)width * height;
    DevBuf<unsigned long long> da; DevBuf<uint32_t> dp; DevBuf<double> ds; DevBuf<uint8_t> d8;
    auto done = [&](int code) { da.release(); dp.release(); ds.release(); d8.release(); return code; };
    cudaError_t e = da.reserve(n * 3);
    if (e == cudaSuccess) e = dp.reserve(n);
    if (e == cudaSuccess && rgb_sum) e = ds.reserve(npx * 3);
    if (e == cudaSuccess && rgb8) e = d8.reserve(npx * 3);
    if (e == cudaSuccess) e = cudaMemcpy(da.p, accum, n * 3 * sizeof(unsigned long long), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(dp.p, poison, n * sizeof(uint32_t), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = launch_resolve_accum_f32(da.p, dp.p, width, height, spp, rgb_sum ? ds.p : nullptr, rgb8 ? d8.p : nullptr, 0);
    if (e == cudaSuccess && rgb_sum) e = cudaMemcpy(rgb_sum, ds.p, npx * 3 * sizeof(double), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && rgb8) e = cudaMemcpy(rgb8, d8.p, npx * 3, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { cudaGetLastError(); return done(fail(RTW_E_CUDA, std::string("rtw_resolve_accum: ") + cudaGetErrorString(e))); }
    return done(RTW_OK);
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
    rc = check_reference_panic(s);
    if (rc) return rc;
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
    if (s->general && precision == RTW_F32) {
        auto P = batch_params<float>(s, s->g32, n);
        P.tmin = (float)tmin; P.tmax = (float)tmax;
        CU(launch_trace_general_f32(P, 0));
    } else if (s->general) {
        auto P = batch_params<double>(s, s->g64, n);
        P.tmin = tmin; P.tmax = tmax;
        CU(launch_trace_general_f64(P, 0));
    } else if (precision == RTW_F32) {
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
    if (s->general && opts->precision == RTW_F32) {
        auto P = batch_params<float>(s, s->g32, n);
        P.seed = opts->seed; P.tmin = resolve_tmin<float>(opts->tmin); P.flags = opts->flags;
        CU(launch_scatter_general_f32(P, 0));
    } else if (s->general) {
        auto P = batch_params<double>(s, s->g64, n);
        P.seed = opts->seed; P.tmin = resolve_tmin<double>(opts->tmin); P.flags = opts->flags;
        CU(launch_scatter_general_f64(P, 0));
    } else if (opts->precision == RTW_F32) {
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
    return check_reference_panic(s);
}


int rtw_shade_batch(rtw_scene* s, const rtw_opts* opts, size_t n, const double* d, const double* p, const double* normal,
                    const uint32_t* front_face, const uint32_t* mat_kind, const double* material, const uint32_t* pixel,
                    const uint32_t* sample, const uint32_t* vertex, uint32_t* kind, double* dir, double* weight) {
    if (!s || (n && (!d || !p || !normal || !front_face || !mat_kind || !material || !pixel || !sample || !vertex || !kind || !dir || !weight)))
        return fail(RTW_E_INVALID, "NULL argument");
    int rc = check_opts(opts); if (rc) return rc;
    if (s->general) return fail(RTW_E_UNSUPPORTED, "rtw_shade_batch: sphere-path scenes only");
    for (size_t i = 0; i < n; ++i)
        if (mat_kind[i] > RTW_INVISIBLE) return fail(RTW_E_INVALID, "rtw_shade_batch: material kind");
    if (n == 0) return RTW_OK;
    CU(cudaSetDevice(s->device));
    rc = reserve_batch(s, n); if (rc) return rc;
    CU(s->d_in2.reserve(3 * n)); CU(s->d_in3.reserve(4 * n)); CU(s->d_u3.reserve(n)); CU(s->d_u4.reserve(n));
    CU(cudaMemcpy(s->d_in0.p, p, 3 * n * sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(s->d_in1.p, d, 3 * n * sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(s->d_in2.p, normal, 3 * n * sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(s->d_in3.p, material, 4 * n * sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(s->d_u0.p, pixel, n * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(s->d_u1.p, sample, n * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(s->d_u2.p, vertex, n * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(s->d_u3.p, front_face, n * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(s->d_u4.p, mat_kind, n * 4, cudaMemcpyHostToDevice));
    auto fill = [&](auto& P, const auto& view) {
        P.scene = view; P.seed = opts->seed; P.n = n;
        P.d = s->d_in1.p; P.p = s->d_in0.p; P.normal = s->d_in2.p; P.material = s->d_in3.p;
        P.front_face = s->d_u3.p; P.mat_kind = s->d_u4.p; P.pixel = s->d_u0.p; P.sample = s->d_u1.p; P.vertex = s->d_u2.p;
        P.kind = s->d_k.p; P.dir = s->d_out3.p; P.weight = s->d_out4.p;
    };
    if (opts->precision == RTW_F32) { ShadeParams<float> P{}; fill(P, s->f32.view); CU(launch_shade_f32(P, 0)); }
    else { ShadeParams<double> P{}; fill(P, s->f64.view); CU(launch_shade_f64(P, 0)); }
    CU(cudaMemcpy(kind, s->d_k.p, n * 4, cudaMemcpyDeviceToHost));
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
    if (s->general && opts->precision == RTW_F32) {
        auto P = batch_params<float>(s, s->g32, n);
        P.cam = to_camera<float>(cam); P.seed = opts->seed; P.tmin = resolve_tmin<float>(opts->tmin); P.flags = opts->flags;
        CU(launch_path_radiance_general_f32(P, 0));
    } else if (s->general) {
        auto P = batch_params<double>(s, s->g64, n);
        P.cam = to_camera<double>(cam); P.seed = opts->seed; P.tmin = resolve_tmin<double>(opts->tmin); P.flags = opts->flags;
        CU(launch_path_radiance_general_f64(P, 0));
    } else if (opts->precision == RTW_F32) {
        BatchParams<float> P = batch_params<float>(s, s->f32, n);
        P.cam = to_camera<float>(cam); P.seed = opts->seed; P.tmin = resolve_tmin<float>(opts->tmin); P.flags = opts->flags;
        CU(launch_path_radiance_f32(P, 0));
    } else {
        BatchParams<double> P = batch_params<double>(s, s->f64, n);
        P.cam = to_camera<double>(cam); P.seed = opts->seed; P.tmin = resolve_tmin<double>(opts->tmin); P.flags = opts->flags;
        CU(launch_path_radiance_f64(P, 0));
    }
    CU(cudaMemcpy(rgb, s->d_out1.p, 3 * n * 8, cudaMemcpyDeviceToHost));
    return check_reference_panic(s);
}

}  // extern "C"

#include "capi_multi.inl"
