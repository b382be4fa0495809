// rtw_device.cuh — device-side building blocks of the B200 path tracer: vectors, Philox4x32-10
// streams, the two arithmetic policies and the flattened scene views.
//
// Two instantiations of everything below exist:
//   Fast  (T = float,  EXACT = false): FP32, FMA contraction allowed, reciprocal-direction slab
//          tests, 24-bit uniforms (stream layout W32).  This is the throughput path.
//   Exact (T = double, EXACT = true ): the reference's f64 operation order (compiled with
//          -fmad=false so nothing is contracted), division-based slab tests, per-sphere AABB
//          pre-test, 53-bit uniforms (stream layout W64).  Bit-compatible with an f64 CPU run of the
//          same algorithm and stream.
// Reference citations are relative to the reference repository (N9199/ray_tracing_weekend).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace rtw {

#define RTW_HD __host__ __device__ __forceinline__
#define RTW_D __device__ __forceinline__

constexpr int kStackDepth = 32;      // traversal stack entries per thread (shared memory)
constexpr int kMaxTreeDepth = 31;    // enforced by the host builder
constexpr int kTileW = 16, kTileH = 16, kWarpTileW = 8, kWarpTileH = 4;
constexpr int kWarpTilesPerTile = (kTileW * kTileH) / 32;

// Tile slots.  Tiles are numbered row-major (tile row ty, column tx); slot k = ty * tiles_x + (tx + rot(ty)) % tiles_x
// rotates every tile row by a pseudo-random amount before slots are dealt round-robin to the ranks (slot k ->
// rank k % world, local tile k / world).  Without the rotation a rank owns whole tile COLUMNS whenever tiles_x is
// a multiple of world (1080p: 120 columns, 8 GPUs) and the ranks' loads differ by ~20 %.
RTW_HD uint32_t tile_row_rotation(uint32_t ty, uint32_t tiles_x) { return ((ty * 0x9E3779B1u) >> 15) % tiles_x; }
RTW_HD uint32_t tile_slot(uint32_t tx, uint32_t ty, uint32_t tiles_x) { return ty * tiles_x + (tx + tile_row_rotation(ty, tiles_x)) % tiles_x; }
RTW_HD void slot_tile(uint32_t slot, uint32_t tiles_x, uint32_t* tx, uint32_t* ty) {
    uint32_t y = slot / tiles_x, c = slot - y * tiles_x;
    *ty = y;
    *tx = (c + tiles_x - tile_row_rotation(y, tiles_x)) % tiles_x;
}

// ---------------------------------------------------------------------------------------------
template <class T> struct V3 { T x, y, z; };
template <class T> RTW_HD V3<T> mk(T x, T y, T z) { return V3<T>{x, y, z}; }
template <class T> RTW_HD V3<T> operator+(V3<T> a, V3<T> b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <class T> RTW_HD V3<T> operator-(V3<T> a, V3<T> b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <class T> RTW_HD V3<T> operator-(V3<T> a) { return {-a.x, -a.y, -a.z}; }
template <class T> RTW_HD V3<T> operator*(V3<T> a, T s) { return {a.x * s, a.y * s, a.z * s}; }
template <class T> RTW_HD V3<T> operator*(V3<T> a, V3<T> b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
template <class T> RTW_HD V3<T> operator/(V3<T> a, T s) { return {a.x / s, a.y / s, a.z / s}; }
template <class T> RTW_HD T dot(V3<T> a, V3<T> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }       // vec.rs:68-72
template <class T> RTW_HD V3<T> cross(V3<T> a, V3<T> b) {                                              // vec.rs:74-82
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// FP32 overloads with EXPLICIT fused multiply-adds.  Both kernel translation units are compiled with
// -fmad=false, so the compiler never decides where an FMA goes: every kernel (megakernel, wavefront, batch)
// evaluates a given expression with the same roundings, which makes their results bit-identical.
RTW_HD float dot(V3<float> a, V3<float> b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
RTW_HD V3<float> cross(V3<float> a, V3<float> b) {
    return {fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x))};
}
template <class T> RTW_HD T sqlen(V3<T> a) { return dot(a, a); }

template <class T> struct Ray { V3<T> o, d; };
template <class T> RTW_HD V3<T> at(const Ray<T>& r, T t) { return r.o + r.d * t; }                     // ray.rs:25-28
RTW_HD V3<float> at(const Ray<float>& r, float t) { return {fmaf(r.d.x, t, r.o.x), fmaf(r.d.y, t, r.o.y), fmaf(r.d.z, t, r.o.z)}; }

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11).  Same constants as curand_philox4x32_x.h.
RTW_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
#ifdef __CUDA_ARCH__
        uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
#else
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t h0 = (uint32_t)(p0 >> 32), l0 = (uint32_t)p0, h1 = (uint32_t)(p1 >> 32), l1 = (uint32_t)p1;
#endif
        uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Uniform stream of one (pixel, sample, vertex): counter = (pixel, sample, vertex, block), key = seed.
// The 32-bit outputs of consecutive blocks form one word sequence x[0], x[1], ...
//   W32 (fast):  uniform k <- x[k]                      24-bit
//   W64 (exact): uniform k <- x[2k] | x[2k+1] << 32     53-bit, rand 0.8 `Standard` / `Open01` semantics
#ifdef __CUDACC__
// One out-of-line copy for kernels whose code size matters more than the call: every refill site of a Stream otherwise inlines the
// ten rounds (~64 instructions; the general-scene renderers held ten copies = 10 KB of a kernel that stalls on instruction fetch).
static __device__ __noinline__ uint4 philox4x32_10_call(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    uint32_t o[4];
    philox4x32_10(c0, c1, c2, c3, k0, k1, o);
    return make_uint4(o[0], o[1], o[2], o[3]);
}
#endif
// ... which is every kernel by now: one copy instead of five took 1.4 % off the wavefront's C2 frame and 6.5 % off C5's once the kernel had
// been cut to the instruction cache's size (profiles/r2_code_size_combinations.jsonl); -DRTW_PHILOX_OUTLINE_ALL=false restores the inlined copies
#ifndef RTW_PHILOX_OUTLINE_ALL
#define RTW_PHILOX_OUTLINE_ALL true
#endif
template <bool EXACT> struct Stream {
    uint32_t k0, k1, pixel, sample, vertex;
    uint32_t k, blk, b0, b1, b2, b3;
    bool outline;       // a compile-time constant at every construction site: refill through philox4x32_10_call
    RTW_HD Stream(uint64_t seed, uint32_t pixel_, uint32_t sample_, uint32_t vertex_, bool outline_ = false)
        : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)), pixel(pixel_), sample(sample_), vertex(vertex_), k(0), blk(0xffffffffu),
          b0(0), b1(0), b2(0), b3(0), outline(outline_ || RTW_PHILOX_OUTLINE_ALL) {}
    RTW_HD uint32_t word32(uint32_t idx) {
        uint32_t block = idx >> 2;
        if (block != blk) {
            uint32_t o[4];
#ifdef __CUDA_ARCH__
            if (outline) { uint4 q = philox4x32_10_call(pixel, sample, vertex, block, k0, k1); o[0] = q.x; o[1] = q.y; o[2] = q.z; o[3] = q.w; }
            else
#endif
            philox4x32_10(pixel, sample, vertex, block, k0, k1, o);
            b0 = o[0]; b1 = o[1]; b2 = o[2]; b3 = o[3];
            blk = block;
        }
        uint32_t lo = (idx & 1) ? b1 : b0, hi = (idx & 1) ? b3 : b2;
        return (idx & 2) ? hi : lo;
    }
    RTW_HD uint64_t next64() { uint64_t lo = word32(2 * k), hi = word32(2 * k + 1); k++; return lo | (hi << 32); }
    RTW_HD uint32_t next32() { return word32(k++); }
};
// Standard: [0,1)
RTW_HD double standard(Stream<true>& s) { return (double)(s.next64() >> 11) * 0x1.0p-53; }
RTW_HD float standard(Stream<false>& s) { return (float)(s.next32() >> 8) * 0x1.0p-24f; }
// Open01: (0,1)
RTW_HD double open01(Stream<true>& s) { return (double)(s.next64() >> 12) * 0x1.0p-52 + 0x1.0p-53; }
RTW_HD float open01(Stream<false>& s) { return (float)(s.next32() >> 9) * 0x1.0p-23f + 0x1.0p-24f; }
// Uniform::new_inclusive(-0.5, 0.5) (camera.rs:275): rand's scale = 1/(1-eps) nudged down, see oracle.
RTW_HD double jitter(Stream<true>& s, double scale) { return (double)(s.next64() >> 12) * 0x1.0p-52 * scale + (-0.5); }
RTW_HD float jitter(Stream<false>& s, float) { return (float)(s.next32() >> 8) * 0x1.0p-24f * 1.0f + (-0.5f); }
// uniform index in [0, n)
RTW_HD uint32_t uindex(Stream<true>& s, uint32_t n) {
    uint64_t w = s.next64();
#ifdef __CUDA_ARCH__
    return (uint32_t)__umul64hi(w, (uint64_t)n);
#else
    return (uint32_t)(((unsigned __int128)w * n) >> 64);
#endif
}
RTW_HD uint32_t uindex(Stream<false>& s, uint32_t n) {
    uint32_t w = s.next32();
#ifdef __CUDA_ARCH__
    return __umulhi(w, n);
#else
    return (uint32_t)(((uint64_t)w * n) >> 32);
#endif
}

// 1-ulp reciprocal on the SFU (MUFU.RCP): the fast path's divisions
RTW_HD float frcp(float x) {
#ifdef __CUDA_ARCH__
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return 1.f / x;
#endif
}
RTW_HD double frcp(double x) { return 1. / x; }
// approximate square root on the SFU (MUFU.SQRT, ~1 ulp)
RTW_HD float fsqrt(float x) {
#ifdef __CUDA_ARCH__
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return sqrtf(x);
#endif
}

// ---------------------------------------------------------------------------------------------
// Arithmetic policies.
template <class T, bool EXACT> struct M;

template <> struct M<double, true> {
    static constexpr double PI = 3.14159265358979323846264338327950288;
    static constexpr double EPS = 2.220446049250313e-16;
    static RTW_HD double inf() { return __builtin_huge_val(); }
    static RTW_HD double max_(double a, double b) { return fmax(a, b); }    // f64::max ignores NaN
    static RTW_HD double min_(double a, double b) { return fmin(a, b); }
    static RTW_HD double sqrt_(double a) { return sqrt(a); }
    static RTW_HD V3<double> normalize(V3<double> a) { return a / sqrt(sqlen(a)); }                  // vec.rs:84-94
    static RTW_HD double div_pi(double a) { return a / PI; }
    static RTW_HD double div(double a, double b) { return a / b; }
    // sin/cos of phi = 2*PI*r: fixed IEEE operation sequence (Cody-Waite by pi/2 + fdlibm kernels),
    // no FMA — repeated bit for bit by any f64 implementation of the same sequence.
    static RTW_HD void sincos_2pi(double r, double* s, double* c) {
        double phi = 2. * PI * r;
        const double two_over_pi = 6.36619772367581382433e-01;
        const double pio2_1 = 1.57079632673412561417e+00, pio2_1t = 6.07710050650619224932e-11;
        int n = (int)(phi * two_over_pi + 0.5);
        double fn = (double)n;
        double y = (phi - fn * pio2_1) - fn * pio2_1t;
        double z = y * y;
        const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04,
                     S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
        const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05,
                     C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
        double ps = S1 + z * (S2 + z * (S3 + z * (S4 + z * (S5 + z * S6))));
        double pc = C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6))));
        double sy = y + (y * z) * ps;
        double cy = (1. - 0.5 * z) + (z * z) * pc;
        switch (n & 3) {
            case 0: *s = sy; *c = cy; break;
            case 1: *s = cy; *c = -sy; break;
            case 2: *s = -sy; *c = -cy; break;
            default: *s = -cy; *c = sy; break;
        }
    }
};

template <> struct M<float, false> {
    static constexpr float PI = 3.14159265358979323846f;
    static constexpr float EPS = 2.220446049250313e-16f;   // the reference's f64::EPSILON, representable in f32
    static RTW_HD float inf() { return __builtin_huge_valf(); }
    static RTW_HD float max_(float a, float b) { return fmaxf(a, b); }
    static RTW_HD float min_(float a, float b) { return fminf(a, b); }
    static RTW_HD float sqrt_(float a) { return fsqrt(a); }
    static RTW_HD V3<float> normalize(V3<float> a) {
#ifdef __CUDA_ARCH__
        return a * rsqrtf(sqlen(a));
#else
        return a / sqrtf(sqlen(a));
#endif
    }
    static RTW_HD float div_pi(float a) { return a * 0.318309886183790671538f; }
    // a / b of the fast path: one SFU reciprocal (the IEEE quotient is ~12 instructions and a slow-path call per site)
#ifdef RTW_IEEE_DIV
    static RTW_HD float div(float a, float b) { return a / b; }
#else
    static RTW_HD float div(float a, float b) { return a * frcp(b); }
#endif
    static RTW_HD void sincos_2pi(float r, float* s, float* c) {
#ifdef __CUDA_ARCH__
        sincospif(2.f * r, s, c);
#else
        *s = sinf(2.f * PI * r); *c = cosf(2.f * PI * r);
#endif
    }
};

// ---------------------------------------------------------------------------------------------
// Flattened scene.  Spheres are stored in BVH-leaf order ("sorted" index); info = prim_id << 2 | kind.
enum MatKind : uint32_t { LAMBERTIAN = 0, METAL = 1, DIELECTRIC = 2, INVISIBLE = 3 };
// bit 31 of a sphere's info word (FP32 scene copy only; set by the host for small scenes): the sphere is ISOLATED — no other sphere's
// surface comes near its ball — see closest_prim_self
constexpr uint32_t kSphereIsolated = 0x80000000u;

template <class T> struct Vec4T { T x, y, z, w; };
template <> struct __align__(16) Vec4T<float> { float x, y, z, w; };
template <> struct __align__(16) Vec4T<double> { double x, y, z, w; };

// Inner node with both child boxes.  child >= 0: inner node index; child < 0: leaf, ~child =
// first_sorted_sphere << 4 | (count - 1), count in 0..16 encoded as (count-1)&15 with a flag for 0
// Child boxes: (la, lb) / (ra, rb) are (min, max) on the f64 path and (centre, half-extent) on the FP32 path.
template <class T> struct __align__(16) Node {
    T la[3], lb[3], ra[3], rb[3];
    int32_t left, right;
    int32_t pad[2];
};
constexpr int32_t kEmptyLeaf = (int32_t)0x80000000;   // leaf with no spheres
constexpr int32_t kStop = (int32_t)0x80000001;        // bottom-of-stack code: traversal finished
RTW_HD int32_t encode_leaf(uint32_t first, uint32_t count) { return count == 0 ? kEmptyLeaf : ~(int32_t)((first << 4) | (count - 1)); }

// Node of the BVH over the LIGHTS (fast path, long light lists), laid out for a STACKLESS walk: nodes in depth-first order;
// `skip` is the node that follows this node's subtree (-1: the walk is over).  `leaf` says what the record holds:
//   kLNodeInner: a box (centre c, half-extent hx hy hz); if the ray crosses it the walk goes on with the next node (its first child)
//   kLNodeLight: ONE light sphere itself (centre c, hx = r^2): the leaf test needs no second, dependent load (ncu, first version
//                with 4-light leaf ranges: 39 % of the samples waiting for the lights' loads at 1.5 active lanes)
//   >= 0:        a box around the lights [leaf >> 4, +(leaf & 15) + 1) of the light list (kept for trees the builder cuts short)
//   kLNodeEmpty: nothing
// lights.pdf_value sums over ALL lights the ray crosses (hittable_list.rs:408-412), so the walk needs no ordering and no range
// shrinking — its whole state is one index plus the running sum, which is what lets the wavefront's CONNECT stage keep two walks
// per lane in flight and suspend / resume them.
struct __align__(32) LNode { float c[3], hx, hy, hz; int32_t skip, leaf; };
constexpr int32_t kLNodeInner = -1, kLNodeEmpty = -2, kLNodeLight = -3;
static_assert(sizeof(LNode) == 32, "two 16-byte loads per node");

template <class T> struct PlaneT { V3<T> point, normal; uint32_t info; uint32_t pad; T albedo[3]; T param; };

// Stride of a node in its SHARED-memory copy (all-shared scenes, SceneViewSh).  The four LDS.128 of a node visit read 16-byte
// chunk k of every lane's node; with the natural 64-byte stride chunk k of node n sits in bank group (4n + k) mod 8 — two of the
// eight 16-byte bank groups for all 32 lanes (ncu, round 1: 2.7-way conflicts on 10.8 G shared-load requests).  With an 80-byte
// stride the group is (5n + k) mod 8: n -> 5n mod 8 is a bijection, so the lanes' nodes spread over all eight groups.  The stride
// is a launch parameter (RenderParams::sh_node_stride: 80 when the padded copy fits next to the path slots, else 64).
constexpr uint32_t kShNodeStridePadded = 80;

template <class T> struct SceneView {
    const Node<T>* nodes;          // BFS order, node 0 = root (global memory)
    const void* nodes_staged;      // FP32: the same nodes at stride kShNodeStridePadded (source of the padded shared-memory copy), or NULL
    const Node<T>* top_nodes;      // the first n_top nodes again, possibly in shared memory
    const Vec4T<T>* spheres;       // sorted: (cx, cy, cz, r)
    const Vec4T<T>* sphere_mat;    // sorted: (albedo r, g, b, param)
    const uint32_t* sphere_info;   // sorted: prim_id << 2 | kind
    const PlaneT<T>* planes;
    const Vec4T<T>* lights;        // (cx, cy, cz, r) in the lights list's insertion order (leaf order when light_nodes is set)
    const LNode* light_nodes;      // optional stackless BVH over the lights (fast path, many lights); global memory
    int32_t n_nodes, n_top, n_spheres, n_planes, n_lights, n_light_nodes;
    int32_t connect_stage;         // the light tree is large: the wavefront renderer walks it in a stage of its own (CONNECT)
};

// Tag type: every section of the scene (nodes, spheres, materials, lights) sits in SHARED memory, so the
// accessors below can issue ld.shared.v4 instead of generic loads (ncu/SASS: the generic path compiled to
// 7 x LD.E.64 per node; this is 4 x LDS.128).
template <class T> struct SceneViewSh : SceneView<T> {
    uint32_t s_nodes = 0, s_spheres = 0, s_mat = 0, s_info = 0, s_lights = 0;   // 32-bit shared-window addresses
    uint32_t s_node_stride = 64;   // bytes between staged nodes (64, or kShNodeStridePadded)
    RTW_D void bind() {
        s_nodes = (uint32_t)__cvta_generic_to_shared(this->top_nodes);
        s_spheres = (uint32_t)__cvta_generic_to_shared(this->spheres);
        s_mat = (uint32_t)__cvta_generic_to_shared(this->sphere_mat);
        s_info = (uint32_t)__cvta_generic_to_shared(this->sphere_info);
        s_lights = (uint32_t)__cvta_generic_to_shared(this->lights);
    }
};
template <class T> RTW_D void bind_scene(SceneView<T>&, uint32_t) {}
template <class T> RTW_D void bind_scene(SceneViewSh<T>& sc, uint32_t node_stride) { sc.bind(); sc.s_node_stride = node_stride; }

RTW_D float4 lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
RTW_D uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
// 32 bytes with ONE request to the L1/TEX pipe (sm_100: ld.global.v8, SASS LDG.E.ENL2.256); p must be 32-byte aligned, read-only data
struct F8 { float4 a, b; };
RTW_D F8 ldg256(const void* p) {
    F8 r;
#ifndef RTW_NO_LDG256
    asm("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=f"(r.a.x), "=f"(r.a.y), "=f"(r.a.z), "=f"(r.a.w), "=f"(r.b.x), "=f"(r.b.y), "=f"(r.b.z), "=f"(r.b.w) : "l"(p));
#else
    r.a = __ldg(reinterpret_cast<const float4*>(p)); r.b = __ldg(reinterpret_cast<const float4*>(p) + 1);
#endif
    return r;
}
// 16 bytes of read-only global data.  Unlike __ldg (asm volatile in the CUDA headers) an unused result is dropped by the compiler, so a
// record can be fetched as a whole and only the chunks a code path reads turn into loads.
RTW_D float4 ldg128(const void* p) {
    float4 v;
    asm("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
RTW_D void unpack_node(float4 a, float4 b, float4 c, float4 d, Node<float>& nd) {
    nd.la[0] = a.x; nd.la[1] = a.y; nd.la[2] = a.z; nd.lb[0] = a.w;
    nd.lb[1] = b.x; nd.lb[2] = b.y; nd.ra[0] = b.z; nd.ra[1] = b.w;
    nd.ra[2] = c.x; nd.rb[0] = c.y; nd.rb[1] = c.z; nd.rb[2] = c.w;
    nd.left = __float_as_int(d.x); nd.right = __float_as_int(d.y);
}
// accessors: generic (any T) / FP32 vectorised / FP32 shared
template <class T> RTW_D void load_node(const SceneView<T>& sc, int32_t cur, Node<T>& nd) { nd = cur < sc.n_top ? sc.top_nodes[cur] : sc.nodes[cur]; }
RTW_D void load_node(const SceneView<float>& sc, int32_t cur, Node<float>& nd) {
#ifndef RTW_NO_LDG256
    if (cur < sc.n_top) {                                    // the BFS prefix a kernel has staged in shared memory (n_top > 0 only then)
        const uint32_t p = (uint32_t)__cvta_generic_to_shared(sc.top_nodes) + (uint32_t)cur * 64u;
        unpack_node(lds128(p), lds128(p + 16), lds128(p + 32), lds128(p + 48), nd);
    } else {                                                 // global memory: two 256-bit loads (cudaMalloc'ed, 64-byte records)
        const F8 lo = ldg256(sc.nodes + cur), hi = ldg256(reinterpret_cast<const char*>(sc.nodes + cur) + 32);
        unpack_node(lo.a, lo.b, hi.a, hi.b, nd);
    }
#else
    const float4* p = reinterpret_cast<const float4*>(cur < sc.n_top ? sc.top_nodes + cur : sc.nodes + cur);
    unpack_node(p[0], p[1], p[2], p[3], nd);
#endif
}
RTW_D void load_node(const SceneViewSh<float>& sc, int32_t cur, Node<float>& nd) {
    uint32_t p = sc.s_nodes + (uint32_t)cur * sc.s_node_stride;
    unpack_node(lds128(p), lds128(p + 16), lds128(p + 32), lds128(p + 48), nd);
}
template <class T> RTW_D Vec4T<T> load_sphere(const SceneView<T>& sc, int32_t i) { return sc.spheres[i]; }
template <class T> RTW_D Vec4T<T> load_sphere_mat(const SceneView<T>& sc, int32_t i) { return sc.sphere_mat[i]; }
template <class T> RTW_D uint32_t load_sphere_info(const SceneView<T>& sc, int32_t i) { return sc.sphere_info[i]; }
template <class T> RTW_D Vec4T<T> load_light(const SceneView<T>& sc, int32_t i) { return sc.lights[i]; }
RTW_D Vec4T<float> as_vec4(float4 v) { return Vec4T<float>{v.x, v.y, v.z, v.w}; }
RTW_D Vec4T<float> load_sphere(const SceneViewSh<float>& sc, int32_t i) { return as_vec4(lds128(sc.s_spheres + (uint32_t)i * 16u)); }
RTW_D Vec4T<float> load_sphere_mat(const SceneViewSh<float>& sc, int32_t i) { return as_vec4(lds128(sc.s_mat + (uint32_t)i * 16u)); }
RTW_D uint32_t load_sphere_info(const SceneViewSh<float>& sc, int32_t i) { return lds32(sc.s_info + (uint32_t)i * 4u); }
// with a light BVH the lights stay in global memory (read through L1), otherwise they are staged like the rest
RTW_D Vec4T<float> load_light(const SceneViewSh<float>& sc, int32_t i) {
    if (sc.n_light_nodes > 0) return as_vec4(__ldg(reinterpret_cast<const float4*>(sc.lights + i)));
    return as_vec4(lds128(sc.s_lights + (uint32_t)i * 16u));
}
// Compile-time knowledge of whether the scene has a light BVH (the wavefront is instantiated for either case, so that neither carries the
// other's code: the kernel sits at the instruction cache's capacity).  LN: 0 = flat light list, 1 = light BVH; a plain view decides at run time.
template <class B, int LN> struct LightMode : B { static constexpr int kLightMode = LN; };
template <class SC> struct light_mode { static constexpr int value = -1; };
template <class B, int LN> struct light_mode<LightMode<B, LN>> { static constexpr int value = LN; };
template <int LN> RTW_D Vec4T<float> load_light(const LightMode<SceneViewSh<float>, LN>& sc, int32_t i) {
    if constexpr (LN == 1) return as_vec4(__ldg(reinterpret_cast<const float4*>(sc.lights + i)));
    else return as_vec4(lds128(sc.s_lights + (uint32_t)i * 16u));
}

// ---------------------------------------------------------------------------------------------
// General scenes (rtw_general.cuh): one table of list entries (entity kind + index + optional Transformed<T>) under one BVH.
enum PrimKind : uint32_t { P_SPHERE = 0, P_PLANE = 1, P_QUAD = 2, P_TRIANGLE = 3, P_CUBOID = 4, P_NO_LIGHTS = 7 };
enum MatKindG : uint32_t { DIFFUSE_LIGHT = 4, ISOTROPIC = 5 };
// (general-scene records are 16-byte aligned and padded to whole 16-byte chunks: the FP32 kernels fetch them with 128-bit loads, g_rec)
template <class T> struct alignas(16) GQuad { V3<T> q, u, v, w, normal; T area; };       // Quad / Triangle / one Cuboid face (quadrilateral.rs:23-32)
// get_plane_uv (plane.rs:41-55) rotates (p - point) about k = normalize(normal x +y) by theta = angle(normal, +y) unless theta <= EPSILON;
// theta, cos, sin and k depend on the plane only and are evaluated on the host in f64 (the same libm calls the reference makes per hit)
template <class T> struct alignas(16) GPlane { V3<T> point, normal, k; T cos_theta, sin_theta; uint32_t rotated, pad; };
template <class T> struct alignas(16) GXform { T fwd[9], ft[3], inv[9], it[3]; };         // Transformation and its inverse (transformations.rs:96-136)
template <class T> struct alignas(16) GPrim {
    T box[6];                // the entry's own world-space box (min, max): bounded_hit's test on the exact path
    uint32_t kind, first;    // entity kind; index into spheres / plane_geo / quads (cuboid: first of its six quads)
    uint32_t mat, id;        // material index; position in the world list (the primitive id the batch calls report)
    int32_t xform;           // -1, or index into xforms: the entry is a Transformed<T>
    uint32_t pad;
};
template <class T> struct alignas(16) GMat { T albedo[3], param; uint32_t kind, texture; };                 // texture: 0 = SolidColour(albedo), k = textures[k-1]
enum TexKind : uint32_t { TEX_NOISE = 1, TEX_CHECKER = 2 };
template <class T> struct alignas(16) GTex { uint32_t kind, perlin, even, odd; T scale, even_c[3], odd_c[3]; };   // texture.rs:24-102
template <class T> struct GPerlin { T rand_vec[256][3]; uint8_t perm_x[256], perm_y[256], perm_z[256]; };
template <class T> struct SceneViewG {
    const Node<T>* nodes;          // BVH over `prims` (leaf ranges index it)
    const GPrim<T>* prims;         // bounded entries in BVH leaf order
    const GPrim<T>* unbounded;     // planes (infinite box): tested linearly
    const GPrim<T>* lights;        // the lights list in iteration order
    const Vec4T<T>* spheres; const GPlane<T>* plane_geo; const GQuad<T>* quads; const GXform<T>* xforms;
    const GMat<T>* mats; const GTex<T>* textures; const GPerlin<T>* perlins;
    int32_t n_nodes, n_prims, n_unbounded, n_lights;
    uint32_t* panic_flag;          // set when a path does what makes the reference panic (a light sample from an empty lights list)
    uint32_t lights_is_bvh, has_xforms, flat;   // flat: few entries, walked linearly (sorted by kind) instead of through the BVH
};
template <class SC> struct is_general { static constexpr bool value = false; };
template <class T> struct is_general<SceneViewG<T>> { static constexpr bool value = true; };

template <class T> struct CameraT {
    V3<T> center, pixel00, du, dv, ddu, ddv, background;
    T defocus_angle, jitter_scale;
    uint32_t width, height, spp, max_depth;
    uint32_t sample_offset;        // this launch renders samples [sample_offset, sample_offset + spp) of every pixel (sample partition)
};

struct DeviceCounters {   // u64 slots in global memory
    unsigned long long paths, rays, node_visits, sphere_tests, light_tests, lambertian, metal, dielectric, absorbed, missed, depth_out;
};

}  // namespace rtw
