// bvh_device.cu — device-side BVH construction for scenes where the host builder is the bottleneck
// (BASELINE config C4: 1 M spheres, 1.0 s of single-threaded binned SAH against a 1.1 s render).
//
// Replaces BoundedVolumeHierarchy::from (shared/src/hittable_collections/bvh.rs:106-143 + hittable_list.rs:318-379: three
// sorts per level on one core) by a linear BVH: 63-bit Morton codes of the sphere centres -> radix sort -> Karras' parallel
// radix-tree construction (one thread per inner node) -> bottom-up box fit (one thread per leaf, atomic arrival flags) ->
// emission in the renderers' node format (both child boxes per inner node; subtrees of <= max_leaf spheres collapse into
// leaf ranges of the Morton-sorted sphere array) -> breadth-first re-layout (stable sort by level) so that the top of the
// tree is the array prefix the kernels pin in shared memory.  Hittable::hit returns argmin-t over the primitives whose own tests
// pass, whatever the tree, so images are bit-identical to those rendered with the host-built tree
// (tests/test_gpu_parity.py::test_device_built_bvh_renders_the_same_image).
// Boxes are exact in f64 (c -/+ r, min / max); the FP32 nodes get the same outward rounding + padding as the host path.
#include <cub/device/device_radix_sort.cuh>

#include "bvh_device.hpp"

namespace rtw {
namespace {

__device__ __forceinline__ unsigned long long spread21(unsigned long long x) {     // 21 bits -> every third bit
    x &= 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__global__ void lbvh_morton_kernel(const double4* spheres, uint32_t n, double lox, double loy, double loz, double sx, double sy, double sz,
                                   unsigned long long* keys, uint32_t* idx) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double4 s = spheres[i];
    auto q = [](double v) { v = fmin(fmax(v, 0.), 2097151.); return (unsigned long long)v; };
    keys[i] = spread21(q((s.x - lox) * sx)) << 2 | spread21(q((s.y - loy) * sy)) << 1 | spread21(q((s.z - loz) * sz));
    idx[i] = i;
}

// common-prefix length of sorted keys i and j; equal keys are ordered by position (Karras 2012, section 4)
__device__ __forceinline__ int lbvh_delta(const unsigned long long* keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    unsigned long long a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz((unsigned)i ^ (unsigned)j);
    return __clzll((long long)(a ^ b));
}

// inner node i of the radix tree over n sorted keys: children (>= 0 inner node, < 0: ~leaf), covered range, parents
__global__ void lbvh_karras_kernel(const unsigned long long* keys, int n, int* left, int* right, int* first, int* last, int* inner_parent,
                                   int* leaf_parent) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int d = lbvh_delta(keys, n, i, i + 1) - lbvh_delta(keys, n, i, i - 1) >= 0 ? 1 : -1;
    int delta_min = lbvh_delta(keys, n, i, i - d);
    int lmax = 2;
    while (lbvh_delta(keys, n, i, i + lmax * d) > delta_min) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2)
        if (lbvh_delta(keys, n, i, i + (l + t) * d) > delta_min) l += t;
    int j = i + l * d;
    int delta_node = lbvh_delta(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) / 2;; t = (t + 1) / 2) {
        if (lbvh_delta(keys, n, i, i + (s + t) * d) > delta_node) s += t;
        if (t == 1) break;
    }
    int gamma = i + s * d + min(d, 0);
    int lo = min(i, j), hi = max(i, j);
    int lc = lo == gamma ? ~gamma : gamma, rc = hi == gamma + 1 ? ~(gamma + 1) : gamma + 1;
    left[i] = lc; right[i] = rc; first[i] = lo; last[i] = hi;
    if (lc >= 0) inner_parent[lc] = i; else leaf_parent[~lc] = i;
    if (rc >= 0) inner_parent[rc] = i; else leaf_parent[~rc] = i;
    if (i == 0) inner_parent[0] = -1;
}

// one thread per leaf: climb; the second thread to arrive at an inner node merges its children's boxes and goes on
__global__ void lbvh_fit_kernel(const double4* spheres, const uint32_t* idx, int n, const int* left, const int* right, const int* inner_parent,
                                const int* leaf_parent, double* nbox /* [n-1][6] */, unsigned int* arrived) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    auto child_box = [&](int c, double* b) {
        if (c < 0) {
            double4 s = spheres[idx[~c]];
            b[0] = s.x - s.w; b[1] = s.y - s.w; b[2] = s.z - s.w; b[3] = s.x + s.w; b[4] = s.y + s.w; b[5] = s.z + s.w;   // Sphere::new, sphere.rs:42-45
        } else {
#pragma unroll
            for (int a = 0; a < 6; ++a) b[a] = nbox[6 * (size_t)c + a];
        }
    };
    int cur = leaf_parent[k];
    while (cur >= 0) {
        __threadfence();
        if (atomicAdd(&arrived[cur], 1u) == 0u) break;                 // the sibling subtree is not finished yet
        __threadfence();
        double a[6], b[6];
        child_box(left[cur], a); child_box(right[cur], b);
#pragma unroll
        for (int x = 0; x < 3; ++x) { nbox[6 * (size_t)cur + x] = fmin(a[x], b[x]); nbox[6 * (size_t)cur + 3 + x] = fmax(a[3 + x], b[3 + x]); }
        cur = inner_parent[cur];
    }
}

// exact depth: one thread per leaf walks to the root (cheap: n * depth loads)
__global__ void lbvh_depth_kernel(int n, const int* inner_parent, const int* leaf_parent, unsigned int* max_depth) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    unsigned int depth = 1;
    for (int cur = leaf_parent[k]; cur >= 0; cur = inner_parent[cur]) depth++;
    atomicMax(max_depth, depth);
}

__device__ __forceinline__ float lbvh_round_up(double v) { float f = (float)v; if ((double)f < v) f = nextafterf(f, __int_as_float(0x7f800000)); return f; }
__device__ __forceinline__ void lbvh_conv(const double* b, float* c, float* h) {         // == fill_node<float> of capi.cu
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float cc = (float)(0.5 * (b[a] + b[3 + a]));
        if (!isfinite(cc)) cc = 0.f;
        double e = fmax(b[3 + a] - (double)cc, (double)cc - b[a]);
        float hh = lbvh_round_up(e);
        hh += 1e-6f * fmaxf(1.f, fabsf(cc) + hh);
        c[a] = cc; h[a] = hh;
    }
}

// inner node i -> Node<double> / Node<float>; a child covering <= max_leaf spheres becomes a leaf range
__global__ void lbvh_emit_kernel(const double4* spheres, const uint32_t* idx, int n, const int* left, const int* right, const int* first,
                                 const int* last, const double* nbox, int max_leaf, Node<double>* n64, Node<float>* n32, unsigned int* counts) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    Node<double> a{};
    Node<float> f{};
    if (i != 0 && last[i] - first[i] + 1 <= max_leaf) { n64[i] = a; n32[i] = f; return; }       // collapsed into its parent's leaf link
    auto child = [&](int c, double* box, int32_t* link) {
        if (c < 0) {
            double4 s = spheres[idx[~c]];
            box[0] = s.x - s.w; box[1] = s.y - s.w; box[2] = s.z - s.w; box[3] = s.x + s.w; box[4] = s.y + s.w; box[5] = s.z + s.w;
            *link = encode_leaf((uint32_t)~c, 1u);
            atomicAdd(&counts[1], 1u);
        } else {
            for (int x = 0; x < 6; ++x) box[x] = nbox[6 * (size_t)c + x];
            int cnt = last[c] - first[c] + 1;
            if (cnt <= max_leaf) { *link = encode_leaf((uint32_t)first[c], (uint32_t)cnt); atomicAdd(&counts[1], 1u); }
            else *link = c;
        }
    };
    double lb[6], rb[6];
    child(left[i], lb, &a.left); child(right[i], rb, &a.right);
    for (int x = 0; x < 3; ++x) { a.la[x] = lb[x]; a.lb[x] = lb[3 + x]; a.ra[x] = rb[x]; a.rb[x] = rb[3 + x]; }
    f.left = a.left; f.right = a.right;
    lbvh_conv(lb, f.la, f.lb); lbvh_conv(rb, f.ra, f.rb);
    n64[i] = a; n32[i] = f;
    atomicAdd(&counts[0], 1u);
}

// Breadth-first re-layout: the renderers pin the first K nodes (the top levels) in shared memory, which only pays when the
// array is level-ordered (C4 wavefront: 988 ms with Karras' order, where the top of the tree is scattered over the array).
// key = level of an emitted inner node (root = 1), 31 for the slots of collapsed subtrees (sorted to the end and dropped).
__global__ void lbvh_level_kernel(int n, const int* inner_parent, const int* first, const int* last, int max_leaf, uint32_t* level, uint32_t* id) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    uint32_t d = 0;
    for (int cur = i; cur >= 0; cur = inner_parent[cur]) d++;
    bool collapsed = i != 0 && last[i] - first[i] + 1 <= max_leaf;
    level[i] = collapsed ? 31u : min(d, 30u);
    id[i] = (uint32_t)i;
}
__global__ void lbvh_rank_kernel(int m, const uint32_t* sorted_id, uint32_t* new_id) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < m) new_id[sorted_id[r]] = (uint32_t)r;
}
__global__ void lbvh_relayout_kernel(int m_used, const uint32_t* sorted_id, const uint32_t* new_id, const Node<double>* in64, const Node<float>* in32,
                                     Node<double>* out64, Node<float>* out32) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m_used) return;
    uint32_t old = sorted_id[r];
    Node<double> a = in64[old];
    Node<float> f = in32[old];
    if (a.left >= 0) a.left = (int32_t)new_id[a.left];
    if (a.right >= 0) a.right = (int32_t)new_id[a.right];
    f.left = a.left; f.right = a.right;
    out64[r] = a; out32[r] = f;
}

__global__ void lbvh_gather_kernel(const double4* spheres, const double4* mats, const uint32_t* info, const uint32_t* idx, uint32_t n,
                                   Vec4T<double>* s64, Vec4T<double>* m64, Vec4T<float>* s32, Vec4T<float>* m32, uint32_t* info_sorted) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint32_t src = idx[k];
    double4 s = spheres[src], m = mats[src];
    s64[k] = Vec4T<double>{s.x, s.y, s.z, s.w}; m64[k] = Vec4T<double>{m.x, m.y, m.z, m.w};
    s32[k] = Vec4T<float>{(float)s.x, (float)s.y, (float)s.z, (float)s.w}; m32[k] = Vec4T<float>{(float)m.x, (float)m.y, (float)m.z, (float)m.w};
    info_sorted[k] = info[src];
}

struct Tmp {
    void* p = nullptr;
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
    ~Tmp() { if (p) cudaFree(p); }
};

}  // namespace

cudaError_t build_lbvh_device(const double* d_spheres, const double* d_mats, const uint32_t* d_info, size_t n_, const double lo[3], const double hi[3],
                              int max_leaf, Node<double>* nodes64, Node<float>* nodes32, Vec4T<double>* s64, Vec4T<double>* m64,
                              Vec4T<float>* s32, Vec4T<float>* m32, uint32_t* info_sorted, DeviceBvhInfo* out, cudaStream_t st) {
    const int n = (int)n_;
    if (n < 2) return cudaErrorInvalidValue;
    Tmp keys, keys2, idx, idx2, left, right, first, last, iparent, lparent, nbox, flags, sort_tmp, t64, t32, level, level2, nid, nid2, newid, sort_tmp2;
    cudaError_t e;
#define LB(x) do { e = (x); if (e != cudaSuccess) return e; } while (0)
    LB(keys.alloc(8 * (size_t)n)); LB(keys2.alloc(8 * (size_t)n)); LB(idx.alloc(4 * (size_t)n)); LB(idx2.alloc(4 * (size_t)n));
    LB(left.alloc(4 * (size_t)n)); LB(right.alloc(4 * (size_t)n)); LB(first.alloc(4 * (size_t)n)); LB(last.alloc(4 * (size_t)n));
    LB(iparent.alloc(4 * (size_t)n)); LB(lparent.alloc(4 * (size_t)n)); LB(nbox.alloc(48 * (size_t)n)); LB(flags.alloc(4 * (size_t)n + 16));
    LB(t64.alloc(sizeof(Node<double>) * (size_t)n)); LB(t32.alloc(sizeof(Node<float>) * (size_t)n));
    LB(level.alloc(4 * (size_t)n)); LB(level2.alloc(4 * (size_t)n)); LB(nid.alloc(4 * (size_t)n)); LB(nid2.alloc(4 * (size_t)n)); LB(newid.alloc(4 * (size_t)n));
    const int B = 256, G = (n + B - 1) / B;
    double s[3];
    for (int a = 0; a < 3; ++a) s[a] = hi[a] > lo[a] ? 2097152. / (hi[a] - lo[a]) : 0.;
    lbvh_morton_kernel<<<G, B, 0, st>>>((const double4*)d_spheres, (uint32_t)n, lo[0], lo[1], lo[2], s[0], s[1], s[2],
                                        (unsigned long long*)keys.p, (uint32_t*)idx.p);
    LB(cudaGetLastError());
    size_t tmp_bytes = 0;
    LB(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (const unsigned long long*)keys.p, (unsigned long long*)keys2.p, (const uint32_t*)idx.p,
                                       (uint32_t*)idx2.p, n, 0, 63, st));
    LB(sort_tmp.alloc(tmp_bytes));
    LB(cub::DeviceRadixSort::SortPairs(sort_tmp.p, tmp_bytes, (const unsigned long long*)keys.p, (unsigned long long*)keys2.p, (const uint32_t*)idx.p,
                                       (uint32_t*)idx2.p, n, 0, 63, st));
    lbvh_karras_kernel<<<G, B, 0, st>>>((const unsigned long long*)keys2.p, n, (int*)left.p, (int*)right.p, (int*)first.p, (int*)last.p,
                                        (int*)iparent.p, (int*)lparent.p);
    LB(cudaGetLastError());
    LB(cudaMemsetAsync(flags.p, 0, 4 * (size_t)n + 16, st));
    unsigned int* arrived = (unsigned int*)flags.p;
    unsigned int* scalars = arrived + n;          // [0] inner nodes emitted, [1] leaves, [2] depth
    lbvh_fit_kernel<<<G, B, 0, st>>>((const double4*)d_spheres, (const uint32_t*)idx2.p, n, (const int*)left.p, (const int*)right.p,
                                     (const int*)iparent.p, (const int*)lparent.p, (double*)nbox.p, arrived);
    LB(cudaGetLastError());
    lbvh_depth_kernel<<<G, B, 0, st>>>(n, (const int*)iparent.p, (const int*)lparent.p, scalars + 2);
    LB(cudaGetLastError());
    lbvh_emit_kernel<<<G, B, 0, st>>>((const double4*)d_spheres, (const uint32_t*)idx2.p, n, (const int*)left.p, (const int*)right.p,
                                      (const int*)first.p, (const int*)last.p, (const double*)nbox.p, max_leaf, (Node<double>*)t64.p, (Node<float>*)t32.p, scalars);
    LB(cudaGetLastError());
    // level-order the emitted nodes (stable sort by level keeps Karras' left-to-right order inside a level)
    lbvh_level_kernel<<<G, B, 0, st>>>(n, (const int*)iparent.p, (const int*)first.p, (const int*)last.p, max_leaf, (uint32_t*)level.p, (uint32_t*)nid.p);
    LB(cudaGetLastError());
    size_t tmp2 = 0;
    LB(cub::DeviceRadixSort::SortPairs(nullptr, tmp2, (const uint32_t*)level.p, (uint32_t*)level2.p, (const uint32_t*)nid.p, (uint32_t*)nid2.p, n - 1, 0, 5, st));
    LB(sort_tmp2.alloc(tmp2));
    LB(cub::DeviceRadixSort::SortPairs(sort_tmp2.p, tmp2, (const uint32_t*)level.p, (uint32_t*)level2.p, (const uint32_t*)nid.p, (uint32_t*)nid2.p, n - 1, 0, 5, st));
    lbvh_rank_kernel<<<G, B, 0, st>>>(n - 1, (const uint32_t*)nid2.p, (uint32_t*)newid.p);
    LB(cudaGetLastError());
    unsigned int h_inner = 0;
    LB(cudaMemcpyAsync(&h_inner, scalars, sizeof(h_inner), cudaMemcpyDeviceToHost, st));
    LB(cudaStreamSynchronize(st));
    lbvh_relayout_kernel<<<((int)h_inner + B - 1) / B, B, 0, st>>>((int)h_inner, (const uint32_t*)nid2.p, (const uint32_t*)newid.p, (const Node<double>*)t64.p,
                                                                  (const Node<float>*)t32.p, nodes64, nodes32);
    LB(cudaGetLastError());
    lbvh_gather_kernel<<<G, B, 0, st>>>((const double4*)d_spheres, (const double4*)d_mats, d_info, (const uint32_t*)idx2.p, (uint32_t)n, s64, m64, s32, m32,
                                        info_sorted);
    LB(cudaGetLastError());
    unsigned int h[4];
    LB(cudaMemcpyAsync(h, scalars, sizeof(h), cudaMemcpyDeviceToHost, st));
    LB(cudaStreamSynchronize(st));
#undef LB
    out->inner_nodes = h[0]; out->leaves = h[1]; out->depth = h[2]; out->node_slots = h[0];       // compact after the re-layout
    return cudaSuccess;
}

}  // namespace rtw
