// kernels_f32.cu — the fast (FP32) instantiation of every kernel.  Compiled with -fmad=false: fused
// multiply-adds are written explicitly (fmaf) in the hot code so that all kernels round identically.
#include <algorithm>
#include <cstdlib>
#include "rtw_launch.cuh"
#include "rtw_wavefront.cuh"

namespace rtw {
// wavefront launch shape: ONE CTA per SM (its warps never synchronise after the scene is staged), 24 warps,
// 96 path slots per warp -> ~216 KB of shared memory for `simple`
constexpr int kWfBlock = 768, kWfSlotsPerWarp = 96;
constexpr int kWfBlockFlat = 640, kWfSlotsPerWarpFlat = 112;       // scenes without a light BVH
}
namespace rtw {
RTW_DEFINE_LAUNCHERS(f32, float, false)
cudaError_t launch_render_pool_f32(RenderParams<float> P, PoolParams Q, bool count, int sm_count, cudaStream_t s, LaunchInfo* info) {
    return count ? launch_render_pool_impl<true>(P, Q, sm_count, s, info) : launch_render_pool_impl<false>(P, Q, sm_count, s, info);
}
cudaError_t launch_render_pool_general_f32(RenderParams<float, SceneViewG<float>> P, PoolParams Q, bool count, int sm_count, cudaStream_t s, LaunchInfo* info) {
    return count ? launch_render_pool_general_impl<true>(P, Q, sm_count, s, info) : launch_render_pool_general_impl<false>(P, Q, sm_count, s, info);
}
struct SideStream { cudaStream_t stream = nullptr; cudaEvent_t fork = nullptr, join = nullptr; };
static thread_local SideStream g_side;
void set_background_side_stream(cudaStream_t stream, cudaEvent_t fork, cudaEvent_t join) { g_side.stream = stream; g_side.fork = fork; g_side.join = join; }
template <bool COUNT, bool SH, int BLOCK, int NP, bool CONN, int LN>
cudaError_t launch_render_wavefront_sh(RenderParams<float> P, PoolParams Q, size_t smem, int sm_count, cudaStream_t s, LaunchInfo* info) {
    auto kernel = render_wavefront_kernel<COUNT, BLOCK, NP, SH, SceneView<float>, CONN, LN>;
    int grid = 0;
    cudaError_t e = persistent_grid(kernel, BLOCK, smem, sm_count, &grid, info);
    if (e != cudaSuccess) return e;
    kernel<<<grid, BLOCK, smem, s>>>(P, Q);
    return cudaGetLastError();
}
// LNS: what the all-shared kernel of this shape knows about the scene's lights at compile time (LightMode: 0 = flat list, 1 = light BVH)
template <bool COUNT, int BLOCK, int NP, bool CONN, int LNS = (CONN ? 1 : 0)>
cudaError_t launch_render_wavefront_shape(RenderParams<float> P, PoolParams Q, uint32_t bvh_depth, int sm_count, cudaStream_t s, LaunchInfo* info) {
    static_assert(!(CONN && LNS != 1), "the CONNECT stage walks the light BVH");
    if ((P.scene.n_light_nodes > 0) != (LNS == 1)) return cudaErrorInvalidConfiguration;
    // shared memory: per-thread traversal stacks + per-warp path slots; what is left (of 227 KB) stages the scene.
    // A walk pushes at most one entry per inner level below the stop code, so `bvh_depth` entries always suffice; two spare
    // entries are kept unless the tree is so deep that the stacks would not fit next to the path slots.
    const size_t limit = 226 * 1024;                      // 227 KB per CTA minus the static shared memory (mbarrier) and slack
    const size_t state = CONN ? wavefront_state_bytes_connect<BLOCK, NP>() : wavefront_state_bytes<BLOCK, NP>();
    const size_t max_entries = (limit - 1024 - state) / (sizeof(int32_t) * BLOCK);
    P.stack_depth = (uint32_t)std::min<size_t>(std::min<uint32_t>(kStackDepth, bvh_depth + 2), max_entries);
    if (P.stack_depth < bvh_depth) return cudaErrorInvalidConfiguration;
    const size_t fixed = sizeof(int32_t) * P.stack_depth * BLOCK + state;
    bool sh = false;
    plan_smem<float, false>(P, BLOCK, &sh, std::min<size_t>(kSmemSceneBudget, (limit - fixed) / 64 * 64));
    size_t scene = P.smem_nodes + (P.smem_spheres ? 2 * (size_t)P.smem_spheres + ((size_t)P.scene.n_spheres * 4 + 15) / 16 * 16 : 0) + P.smem_lights;
    size_t smem = fixed + scene;
    cudaError_t e = pool_clear(P, Q, s);
    if (e != cudaSuccess) return e;
    // render_background_kernel runs on the caller's second stream when there is one (set_background_side_stream): forked HERE — everything it reads
    // has been written, the accumulators are clear — but launched behind the wavefront kernel, so that its small CTAs fill what the wavefront's one
    // CTA per SM leaves free; joined before the finalize
    const SideStream side = g_side;
    const bool fork = Q.queue_len && side.stream && side.fork && side.join;
    if (fork) {
        e = cudaEventRecord(side.fork, s);
        if (e != cudaSuccess) return e;
    }
    // (all-shared scenes: one instantiation per kind of light list, LightMode)
    e = sh ? launch_render_wavefront_sh<COUNT, true, BLOCK, NP, CONN, LNS>(P, Q, smem, sm_count, s, info)
           : launch_render_wavefront_sh<COUNT, false, BLOCK, NP, CONN, -1>(P, Q, smem, sm_count, s, info);
    if (e != cudaSuccess) return e;
    if (Q.queue_len) {                                      // the background-only chunks the queue does not hold (chunk_split_kernel)
        if (fork) {
            e = cudaStreamWaitEvent(side.stream, side.fork, 0);
            if (e != cudaSuccess) return e;
        }
        e = launch_render_background_f32(P, Q, Q.chunk_order, sm_count, fork ? side.stream : s);
        if (e != cudaSuccess) return e;
        if (fork) {
            e = cudaEventRecord(side.join, side.stream);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(s, side.join, 0);
            if (e != cudaSuccess) return e;
        }
    }
    return pool_finalize(P, Q, s);
}
// scenes with a LARGE light BVH run the CONNECT variant (for short walks the stage costs more than it saves: C5, 399 lights, 29.8 ms
// with it against 21.3 without): 9 more bytes of state per path slot (suspended light walk), hence 84 slots per warp
constexpr int kWfSlotsPerWarpConnect = 84;
inline bool connect_stage_allowed() {
    static const bool ok = [] { const char* e = std::getenv("RTW_NO_CONNECT"); return !(e && std::atoi(e) == 1); }();   // A/B measurements
    return ok;
}
template <bool COUNT>
cudaError_t launch_render_wavefront_impl(RenderParams<float> P, PoolParams Q, uint32_t bvh_depth, int sm_count, cudaStream_t s, LaunchInfo* info) {
#ifdef RTW_WF_SWEEP
    // tuning build only (nvcc -DRTW_WF_SWEEP): launch shape from the environment
    const char* e = std::getenv("RTW_WF_SHAPE");
    int shape = e && P.scene.n_light_nodes == 0 ? std::atoi(e) : 0;
    if (shape == 1) return launch_render_wavefront_shape<COUNT, 896, 64, false>(P, Q, bvh_depth, sm_count, s, info);
    if (shape == 2) return launch_render_wavefront_shape<COUNT, 832, 80, false>(P, Q, bvh_depth, sm_count, s, info);
    if (shape == 3) return launch_render_wavefront_shape<COUNT, 640, 112, false>(P, Q, bvh_depth, sm_count, s, info);
    if (shape == 4) return launch_render_wavefront_shape<COUNT, 704, 96, false>(P, Q, bvh_depth, sm_count, s, info);
    if (shape == 5) return launch_render_wavefront_shape<COUNT, 768, 88, false>(P, Q, bvh_depth, sm_count, s, info);
    if (shape == 6) return launch_render_wavefront_shape<COUNT, 576, 128, false>(P, Q, bvh_depth, sm_count, s, info);
    if (shape == 7) return launch_render_wavefront_shape<COUNT, 512, 144, false>(P, Q, bvh_depth, sm_count, s, info);
    if (shape == 8) return launch_render_wavefront_shape<COUNT, 640, 104, false>(P, Q, bvh_depth, sm_count, s, info);
    if (shape == 9) return launch_render_wavefront_shape<COUNT, 448, 160, false>(P, Q, bvh_depth, sm_count, s, info);
#endif
    if (P.scene.n_light_nodes > 0 && P.scene.connect_stage && connect_stage_allowed()) {
#ifdef RTW_CONN_SWEEP
        const char* e = std::getenv("RTW_CONN_SHAPE");
        int shape = e ? std::atoi(e) : 0;
        if (shape == 1) return launch_render_wavefront_shape<COUNT, 512, 132, true>(P, Q, bvh_depth, sm_count, s, info);
        if (shape == 2) return launch_render_wavefront_shape<COUNT, 640, 104, true>(P, Q, bvh_depth, sm_count, s, info);
        if (shape == 3) return launch_render_wavefront_shape<COUNT, 384, 180, true>(P, Q, bvh_depth, sm_count, s, info);
#endif
        return launch_render_wavefront_shape<COUNT, kWfBlock, kWfSlotsPerWarpConnect, true>(P, Q, bvh_depth, sm_count, s, info);
    }
    // flat light list: 20 warps x 112 slots (swept again once the kernel fitted the instruction cache: 768x96 133.4 ms, 640x112 131.4, 576x128 134.4,
    // 640x104 134.1, 832x80 143.0 on C2; C1 +2.9 %); the light-BVH kernel stays at 24 x 96 (C5: 50.7 ms against 52.2 / 54.3 at 640x112 / 576x128)
    if (P.scene.n_light_nodes == 0) return launch_render_wavefront_shape<COUNT, kWfBlockFlat, kWfSlotsPerWarpFlat, false, 0>(P, Q, bvh_depth, sm_count, s, info);
    return launch_render_wavefront_shape<COUNT, kWfBlock, kWfSlotsPerWarp, false, 1>(P, Q, bvh_depth, sm_count, s, info);
}
// general scenes: the same warp-private wavefront and launch shape, scene tables in global memory.  Swept on cornell_box while the kernel
// was bound by instruction fetch: 512x96 460 ms, 512x128 450, 576x96 440, 640x96 408, 704x96 386, 768x96 364; again after its code
// had been cut from 146 to 57 KB: 768x96 259.0, 768x112 255.4, 640x128 276.4, 896x80 253.1 (72 registers, spills), 704x104 266.6, 832x88 256.5
constexpr int kWfBlockG = 768, kWfSlotsPerWarpG = 96;
template <bool COUNT, int BLOCK, int NP>
cudaError_t launch_render_wavefront_general_shape(RenderParams<float, SceneViewG<float>> P, PoolParams Q, uint32_t bvh_depth, int sm_count, cudaStream_t s, LaunchInfo* info) {
    const size_t limit = 226 * 1024, state = wavefront_state_bytes_general<BLOCK, NP>();
    const size_t max_entries = (limit - 1024 - state) / (sizeof(int32_t) * BLOCK);
    P.stack_depth = (uint32_t)std::min<size_t>(std::min<uint32_t>(kStackDepth, bvh_depth + 2), max_entries);
    if (P.stack_depth < bvh_depth) return cudaErrorInvalidConfiguration;
    const size_t smem = sizeof(int32_t) * P.stack_depth * BLOCK + state;
    cudaError_t e = pool_clear(P, Q, s);
    if (e != cudaSuccess) return e;
    auto kernel = render_wavefront_kernel<COUNT, BLOCK, NP, false, SceneViewG<float>>;
    int grid = 0;
    e = persistent_grid(kernel, BLOCK, smem, sm_count, &grid, info);
    if (e != cudaSuccess) return e;
    kernel<<<grid, BLOCK, smem, s>>>(P, Q);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return pool_finalize(P, Q, s);
}
template <bool COUNT>
cudaError_t launch_render_wavefront_general_impl(RenderParams<float, SceneViewG<float>> P, PoolParams Q, uint32_t bvh_depth, int sm_count, cudaStream_t s, LaunchInfo* info) {
#ifdef RTW_WF_SWEEP
    const char* e = std::getenv("RTW_WFG_SHAPE");
    int shape = e ? std::atoi(e) : 0;
    if (shape == 1) return launch_render_wavefront_general_shape<COUNT, 768, 112>(P, Q, bvh_depth, sm_count, s, info);
    if (shape == 2) return launch_render_wavefront_general_shape<COUNT, 640, 128>(P, Q, bvh_depth, sm_count, s, info);
    if (shape == 3) return launch_render_wavefront_general_shape<COUNT, 896, 80>(P, Q, bvh_depth, sm_count, s, info);
    if (shape == 4) return launch_render_wavefront_general_shape<COUNT, 704, 104>(P, Q, bvh_depth, sm_count, s, info);
    if (shape == 5) return launch_render_wavefront_general_shape<COUNT, 832, 88>(P, Q, bvh_depth, sm_count, s, info);
#endif
    return launch_render_wavefront_general_shape<COUNT, kWfBlockG, kWfSlotsPerWarpG>(P, Q, bvh_depth, sm_count, s, info);
}
cudaError_t launch_render_wavefront_general_f32(RenderParams<float, SceneViewG<float>> P, PoolParams Q, uint32_t bvh_depth, bool count, int sm_count,
                                                cudaStream_t s, LaunchInfo* info) {
    return count ? launch_render_wavefront_general_impl<true>(P, Q, bvh_depth, sm_count, s, info)
                 : launch_render_wavefront_general_impl<false>(P, Q, bvh_depth, sm_count, s, info);
}
cudaError_t launch_resolve_accum_f32(const unsigned long long* accum, const uint32_t* poison, uint32_t width, uint32_t height, uint32_t spp,
                                     double* rgb_sum, uint8_t* rgb8, cudaStream_t s) {
    return launch_resolve_accum(accum, poison, width, height, spp, rgb_sum, rgb8, s);
}
cudaError_t launch_peer_reduce_resolve_f32(const PeerBlocks& B, uint32_t slot_begin, uint32_t slot_end, uint32_t width, uint32_t height,
                                           uint32_t spp, double* rgb_sum, uint8_t* rgb8, cudaStream_t s) {
    if (slot_end <= slot_begin) return cudaSuccess;
    const uint32_t tiles_x = (width + kTileW - 1) / kTileW, tiles_total = tiles_x * ((height + kTileH - 1) / kTileH);
    peer_reduce_resolve_kernel<0><<<(slot_end - slot_begin + 255) / 256, 256, 0, s>>>(B, slot_begin, slot_end, width, height, tiles_x,
                                                                                       tiles_total, spp, rgb_sum, rgb8);
    return cudaGetLastError();
}
// deepest BVH the default wavefront shape can traverse: its per-thread stacks share the CTA's shared memory with the path slots
uint32_t wavefront_max_bvh_depth() {
    const size_t limit = 226 * 1024, state = wavefront_state_bytes<kWfBlock, kWfSlotsPerWarp>(), state_flat = wavefront_state_bytes<kWfBlockFlat, kWfSlotsPerWarpFlat>();
    size_t entries = std::min((limit - 1024 - state) / (sizeof(int32_t) * kWfBlock), (limit - 1024 - state_flat) / (sizeof(int32_t) * kWfBlockFlat));
    return (uint32_t)std::min<size_t>(entries, kStackDepth - 2);
}
cudaError_t launch_render_wavefront_f32(RenderParams<float> P, PoolParams Q, uint32_t bvh_depth, bool count, int sm_count, cudaStream_t s, LaunchInfo* info) {
    return count ? launch_render_wavefront_impl<true>(P, Q, bvh_depth, sm_count, s, info)
                 : launch_render_wavefront_impl<false>(P, Q, bvh_depth, sm_count, s, info);
}
size_t primary_candidates_scratch_bytes(uint32_t width, uint32_t height) {
    return (size_t)((width + kCandBlock - 1) / kCandBlock) * ((height + kCandBlock - 1) / kCandBlock) * sizeof(CandBlockList);
}
cudaError_t launch_primary_candidates_f32(const SceneView<float>& scene, const CameraT<float>& cam, void* scratch, uint4* cand, cudaStream_t s) {
    if (!cam.width || !cam.height) return cudaSuccess;
    dim3 grid((cam.width + kCandBlock - 1) / kCandBlock, (cam.height + kCandBlock - 1) / kCandBlock);
    const uint32_t n_blocks = grid.x * grid.y;
    CandBlockList* lists = reinterpret_cast<CandBlockList*>(scratch);
    block_candidates_kernel<0><<<(n_blocks + 127) / 128, 128, 0, s>>>(scene, cam, grid.x, n_blocks, lists);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    primary_candidates_kernel<0><<<grid, kCandBlock * kCandBlock, 0, s>>>(scene, cam, lists, cand);
    return cudaGetLastError();
}
cudaError_t launch_chunk_order_f32(const uint4* cand, const SceneView<float>& scene, const CameraT<float>& cam, uint32_t rank, uint32_t world,
                                   uint32_t tiles_x, uint32_t tiles_total, uint32_t n_slots, uint32_t pixels_per_chunk, uint32_t n_chunks, uint32_t cap,
                                   uint32_t own_rank, uint32_t own_world, uint32_t* order, cudaStream_t s) {
    if (!n_chunks) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(order + cap + 2 * (size_t)n_chunks, 0, 2 * sizeof(uint32_t), s);
    if (e != cudaSuccess) return e;
    chunk_order_kernel<0><<<(n_chunks + 255) / 256, 256, 0, s>>>(cand, scene, cam, rank, world, tiles_x, tiles_total, n_slots, pixels_per_chunk,
                                                                 n_chunks, cap, own_rank, own_world, order);
    return cudaGetLastError();
}
cudaError_t launch_chunk_split_f32(uint32_t* order, uint32_t n_chunks, uint32_t cap, uint32_t tail_chunks, uint32_t split_chunks, cudaStream_t s) {
    chunk_split_kernel<0><<<(cap + 255u) / 256u, 256, 0, s>>>(order, n_chunks, cap, tail_chunks, split_chunks);
    return cudaGetLastError();
}
size_t chunk_order_words(uint32_t n_chunks, uint32_t cap) { return order_words(n_chunks, cap); }
uint32_t chunk_order_extra(uint32_t split_chunks) { return order_extra(split_chunks); }
cudaError_t launch_render_background_f32(const RenderParams<float>& P, const PoolParams& Q, const uint32_t* order, int sm_count, cudaStream_t s) {
    static int per_sm = 0;                                  // resident CTAs per SM (64 registers: the slow path may spill, the loop does not)
    if (per_sm == 0) {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, render_background_kernel<0>, kBackgroundBlock, 0) != cudaSuccess || n < 1) n = 2;
        per_sm = n;
    }
    render_background_kernel<0><<<(sm_count > 0 ? sm_count : 148) * per_sm, kBackgroundBlock, 0, s>>>(P, Q, const_cast<uint32_t*>(order), Q.n_chunks);
    return cudaGetLastError();
}
// samples of this radiance or more set a pixel's overflow flag instead of being added: spp of them stay below 2^60 fixed-point units
// (2^28 radiance units) in total, whichever way they are split over lanes, launches and ranks
float pool_sample_cap(uint32_t spp_total) {
    const float cap = 268435456.f / (float)(spp_total ? spp_total : 1u);       // 2^28 / spp
    return cap < kFixedMax ? cap : kFixedMax;
}
// chunk = G pixel slots x spp paths; aim for >= 128 paths per queue transaction (512 in round 1: the smaller chunk costs nothing at
// 500 spp, where a chunk is one pixel anyway, and takes 10 % off an 8-spp frame: the warps run dry closer together)
uint32_t pool_pixels_per_chunk(uint32_t spp) {
    if (spp == 0) return 1;
    static const uint32_t target = [] { const char* e = std::getenv("RTW_CHUNK_PATHS"); int v = e ? std::atoi(e) : 0; return v >= 32 && v <= 65536 ? (uint32_t)v : 128u; }();
    uint32_t g = (target + spp - 1) / spp;
    return g < 1 ? 1 : (g > 256 ? 256 : g);
}
}
#ifdef RTW_TIMELINE
extern "C" __attribute__((visibility("default"))) int rtw_debug_timeline(unsigned long long* out, size_t n_words) {
    const size_t cap = sizeof(rtw::rtw_timeline) / sizeof(unsigned long long);
    return (int)cudaMemcpyFromSymbol(out, rtw::rtw_timeline, sizeof(unsigned long long) * (n_words < cap ? n_words : cap));
}
#endif
