// rtw_wavefront.cuh — the wavefront renderer (RTW_WAVEFRONT, fast path only).
//
// generate / extend / shade queues sorted by material and compacted with warp ballots — held in SHARED
// memory and private to each WARP:
//
//   every warp keeps NPW (96) paths in flight in its own slice of shared memory (SoA path state + one index
//   list per stage) and repeatedly runs the stage whose list is longest on up to 32 of its paths, one per lane:
//     GENERATE  free slots take the next paths of the warp's chunk of the pixel-major path stream
//               (Camera::get_ray, camera.rs:274-293)                                       -> extend list
//     EXTEND    closest hit (bvh.rs:163-188); miss -> the path ends (camera.rs:473-475), hit -> list of its material
//     SHADE     one list per material (Lambertian + light-pdf loop / Metal / Dialectric): all lanes run the same
//               Material::scatter (material.rs:357-488)                                    -> extend list
//
// Why: the megakernel runs trace + every material branch in lock step per lane; ncu shows 12.9 of 32 threads
// active, the expensive Lambertian branch (29-light pdf loop) executing with ~1/3 of the lanes.  Here a stage
// only runs when (nearly) 32 paths want it: with 96 slots per warp a simulation of the measured stage mix gives
// > 31 busy lanes per batch (64 slots: 26).
// Why warp-private and in shared memory: a global-memory wavefront moves ~130 B of path state per bounce through
// L2 / HBM (~350 GB per 1080p / 500 spp frame) and needs thousands of launches; a CTA-wide version of these queues
// (first attempt, profiles/) spent 21 % of its time in bar.sync and ran at 16 warps / SM.  Warp-private lists need
// no barrier and no atomics: every list operation is a ballot + popcount prefix, counts live in registers.
// The paths, their RNG streams and the arithmetic are exactly the pooled megakernel's and radiance is accumulated
// in the same 64-bit fixed point, so both renderers produce bit-identical images (tests/test_gpu_parity.py).
#pragma once
#include "rtw_kernels.cuh"

namespace rtw {

enum WfStage : int { WF_FREE = 0, WF_EXT = 1, WF_LAMB = 2, WF_METAL = 3, WF_DIEL = 4, WF_STAGES = 5, WF_CONN = 5 };

template <int NPW> struct WfWarp {
    float ox[NPW], oy[NPW], oz[NPW], dx[NPW], dy[NPW], dz[NPW], mx[NPW], my[NPW], mz[NPW], ht[NPW];
    uint32_t q[NPW], pix[NPW], smp[NPW], dep[NPW];  // dep = depth | res-is-NaN bits << 16
    int32_t hp[NPW];                                 // hit primitive: >= 0 sorted sphere, <= -2 plane
    uint8_t list[WF_STAGES][NPW];
};

// general scenes (rtw_general.cuh) also remember which quad of the winning entry was hit
template <int NPW> struct WfWarpG : WfWarp<NPW> { uint32_t hs[NPW]; };
// scenes with a light BVH (CONNECT stage): the suspended light walk of a path — the node to resume at and the running sum of
// lights.pdf_value — plus the stage's own list; S.ht carries the cosine term of the sampled direction through the stage
template <int NPW> struct WfWarpC : WfWarp<NPW> { int32_t lcur[NPW]; float lacc[NPW]; uint8_t conn[NPW]; };

// warp-synchronous push: every lane of the warp calls it; lanes with pred append `slot`
RTW_D void wf_push(uint8_t* list, uint32_t& count, bool pred, uint32_t slot, uint32_t lt_mask) {
    uint32_t m = __ballot_sync(0xffffffffu, pred);
    if (pred) list[count + __popc(m & lt_mask)] = (uint8_t)slot;
    count += __popc(m);
}

RTW_D V3<float> wf_res(uint32_t dep) {
    const float qnan = __int_as_float(0x7fc00000);
    return mk<float>((dep >> 16) & 1u ? qnan : 0.f, (dep >> 17) & 1u ? qnan : 0.f, (dep >> 18) & 1u ? qnan : 0.f);
}

// lane-private partial sums of the pixel the lane last finished a path of (same scheme as the pooled megakernel)
struct WfAcc {
    uint32_t q = 0xffffffffu, bad = 0;
    unsigned long long a0 = 0, a1 = 0, a2 = 0;
};
RTW_D void wf_acc_flush(const PoolParams& Q, WfAcc& A) {
    if (A.q != 0xffffffffu) {
        pool_flush(Q, A.q, A.a0, A.a1, A.a2);
        if (A.bad) atomicOr(Q.poison + A.q, A.bad);
    }
}
RTW_D void wf_finish(const PoolParams& Q, WfAcc& A, uint32_t q, V3<float> value, uint32_t flags) {
    if (flags & 1u) value = fix_nan(value);
    if (q != A.q) {
        wf_acc_flush(Q, A);
        A.q = q; A.a0 = A.a1 = A.a2 = 0ull; A.bad = 0;
    }
    A.a0 += pool_fixed(Q, value.x, 0, A.bad);
    A.a1 += pool_fixed(Q, value.y, 1, A.bad);
    A.a2 += pool_fixed(Q, value.z, 2, A.bad);
}

// The tail of one vertex of ray_colour_tail_call (camera.rs:484-521) for a path whose scatter has been evaluated: absorbed ->
// the path ends with mult * emitted + res; otherwise mult *= weight, depth -= 1, and the path either runs out of depth (ends with
// res) or goes back to EXTEND with its new ray.  Shared by SHADE and by CONNECT (which finishes Lambertian vertices).
// Returns 0 = continues (state stored, push to EXTEND), 1 = ended (*fin_value holds the sample).
template <bool COUNT, class WS>
RTW_D int wf_advance(WS& S, uint32_t slot, uint32_t kind, const Ray<float>& next, V3<float> w, V3<float> emitted, V3<float> mult, uint32_t dep,
                     V3<float>* fin_value, Tally& tl) {
    uint32_t depth = dep & 0xffffu;
    if (kind == V_ABSORB) {                                     // camera.rs:484-486
        *fin_value = mult * emitted + wf_res(dep);
        return 1;
    }
    if (kind == V_DIFFUSE) {                                    // res + mult * emitted (camera.rs:519): 0 or NaN per channel
        V3<float> rs = wf_res(dep) + mult * emitted;
        dep |= (rs.x != rs.x ? 1u << 16 : 0u) | (rs.y != rs.y ? 1u << 17 : 0u) | (rs.z != rs.z ? 1u << 18 : 0u);
    }
    mult = mult * w;
    depth -= 1;
    dep = (dep & 0xffff0000u) | depth;
    if (depth == 0) {                                           // camera.rs:470-472
        if (COUNT) tl.depth_out++;
        *fin_value = mk<float>(0, 0, 0) + wf_res(dep);
        return 1;
    }
    S.ox[slot] = next.o.x; S.oy[slot] = next.o.y; S.oz[slot] = next.o.z;
    S.dx[slot] = next.d.x; S.dy[slot] = next.d.y; S.dz[slot] = next.d.z;
    S.mx[slot] = mult.x; S.my[slot] = mult.y; S.mz[slot] = mult.z;
    S.dep[slot] = dep;
    return 0;
}

// CONNECT tuning: node visits between two finish / refill passes of a batch of light walks
#ifndef RTW_CONN_QUANTUM
#define RTW_CONN_QUANTUM 16
#endif
constexpr uint32_t kConnQuantum = RTW_CONN_QUANTUM;

#ifdef RTW_TIMELINE
// diagnostic build (scripts/timeline_probe.py): per warp, %globaltimer at kernel entry / first background-only chunk / queue dry / exit,
// the paths still in flight when the queue ran dry and the stage passes made after that
__device__ unsigned long long rtw_timeline[148 * 32 * 8];
RTW_D unsigned long long wf_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#endif

template <bool COUNT, int BLOCK, int NPW, bool SH, class SCENE = SceneView<float>, bool CONN = false, int LN = -1>
__global__ void __launch_bounds__(BLOCK, 1) render_wavefront_kernel(RenderParams<float, SCENE> P, PoolParams Q) {
    using T = float;
    constexpr bool EXACT = false;
    constexpr bool GEN = is_general<SCENE>::value;        // general scenes: entries of any kind, scene tables in global memory
    static_assert(NPW <= 255 && NPW >= 32, "slot indices are stored in one byte");
    static_assert(!(GEN && SH), "general scenes are read from global memory");
    static_assert(!(GEN && CONN), "the CONNECT stage serves the sphere path's light BVH");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: [stack][scene sections][one WfWarp per warp]
    const uint32_t stack_depth = P.stack_depth;
    int32_t* stack_base = reinterpret_cast<int32_t*>(smem_raw);
    unsigned char* cur_p = smem_raw + sizeof(int32_t) * stack_depth * BLOCK;
    using SC0 = typename std::conditional<SH, SceneViewSh<T>, SCENE>::type;
    using SC = typename std::conditional<(SH && LN >= 0), LightMode<SC0, LN>, SC0>::type;       // light BVH or not: known at compile time (LightMode)
    static_assert(!(CONN && LN == 0), "the CONNECT stage walks the light BVH");
    SC sc;
    if constexpr (GEN) sc = P.scene;
    else {
        SceneView<T> sc0 = P.scene;
        stage_scene(P, cur_p, sc0);                       // TMA bulk copies; the only block-wide wait of the kernel
        cur_p += P.smem_nodes + (P.smem_spheres ? 2u * P.smem_spheres + ((uint32_t)P.scene.n_spheres * 4u + 15u) / 16u * 16u : 0u) + P.smem_lights;
        static_cast<SceneView<T>&>(sc) = sc0;
        bind_scene(sc, P.sh_node_stride);
    }
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, lt_mask = (1u << lane) - 1u;
    using WS = typename std::conditional<GEN, WfWarpG<NPW>, typename std::conditional<CONN, WfWarpC<NPW>, WfWarp<NPW>>::type>::type;
    WS& S = reinterpret_cast<WS*>(cur_p)[warp];

    const CameraT<T>& cam = P.cam;
    int32_t* stack = stack_base + tid;
    const uint32_t n_slots = P.n_local_tiles * (kTileW * kTileH);
    const uint32_t spp = cam.spp, G = Q.pixels_per_chunk;
    uint32_t npaths = 0, nrays = 0;
    Tally tl;
    WfAcc acc;

    // warp-uniform state (registers): list lengths and the cursor into the path stream
    uint32_t n_free = NPW, n_ext = 0, n_lamb = 0, n_metal = 0, n_diel = 0, n_conn = 0;
    uint32_t chunk_next = 0, chunk_end = 0, chunk_q0 = 0;
    bool exhausted = (spp == 0 || cam.max_depth == 0);       // max_depth == 0: every path returns 0 (camera.rs:470-472)
    bool cheap_phase = false;                                // the queue has reached its background-only chunks (chunk_order_kernel)
    for (uint32_t i = lane; i < NPW; i += 32) S.list[WF_FREE][i] = (uint8_t)i;
#ifdef RTW_TIMELINE
    unsigned long long tl_t[4] = {wf_now(), 0ull, 0ull, 0ull};
    uint32_t tl_inflight = 0, tl_passes = 0;
#endif

    for (;;) {
        __syncwarp();
#ifdef RTW_TIMELINE
        if (cheap_phase && !tl_t[1]) tl_t[1] = wf_now();
        if (exhausted && chunk_next == chunk_end) { if (!tl_t[2]) { tl_t[2] = wf_now(); tl_inflight = NPW - n_free; } tl_passes++; }
#endif
        // ---- pick the stage with the longest list (free slots only count while the stream has paths left) ----
        uint32_t nf = (exhausted && chunk_next == chunk_end) ? 0u : n_free;
        // cheap tail of the stream: old paths first, whatever the length of their lists — new paths there end in GENERATE anyway, and
        // every pass spent on an old one now is a pass that does not keep the frame waiting after the stream has run dry
        if (cheap_phase && (n_ext | n_lamb | n_metal | n_diel | n_conn)) nf = 0u;
        int stage = WF_FREE;
        uint32_t best = nf;
        if (n_ext > best) { best = n_ext; stage = WF_EXT; }
        if (n_lamb > best) { best = n_lamb; stage = WF_LAMB; }
        if (n_metal > best) { best = n_metal; stage = WF_METAL; }
        if (n_diel > best) { best = n_diel; stage = WF_DIEL; }
        if (CONN && n_conn > best) { best = n_conn; stage = WF_CONN; }
        if (best == 0) break;                                 // nothing in flight and the stream is dry

        if (stage == WF_FREE) {
            // ---- GENERATE ----------------------------------------------------------------------------------
            if (chunk_next == chunk_end) {
                uint32_t c = 0;
                if (lane == 0) c = atomicAdd(P.work_counter, 1u);
                c = __shfl_sync(0xffffffffu, c, 0);
                if (c >= Q.queue_cap) { exhausted = true; continue; }
                uint32_t piece = 0;
                if (Q.chunk_order) {
                    c = __ldg(Q.chunk_order + c);
                    // the first of the chunks that belong to render_background_kernel: the queue ends here.  (A flag in the entry the warp
                    // fetches anyway: holding the queue length in a register instead cost this kernel 2 %, profiles/r2_background_kernel_ab.jsonl)
                    if (c & kChunkEnd) { exhausted = true; continue; }
                    cheap_phase = (c & kChunkCheap) != 0u;
                    piece = (c >> kChunkSubShift) & (2u * kChunkSubs - 1u);
                    c &= kChunkMask;
                }
                chunk_q0 = c * G;
                uint32_t npx = min(G, n_slots - chunk_q0);
                chunk_next = 0; chunk_end = npx * spp;
                if (piece) {                                  // one of the last costly chunks: this entry stands for an eighth of its paths
                    chunk_next = (piece - 1u) * chunk_end / kChunkSubs;
                    chunk_end = piece * chunk_end / kChunkSubs;
                }
                if (G == 1) {       // skip padding pixels (outside the image / padding tiles) as a whole
                    uint32_t tile = (chunk_q0 >> 8) * P.world + P.rank, in = chunk_q0 & 255u;
                    uint32_t ttx, tty;
                    slot_tile(tile, P.tiles_x, &ttx, &tty);
                    uint32_t i = ttx * kTileW + (in & 15u), j = tty * kTileH + (in >> 4);
                    if (!(tile < P.tiles_total && i < cam.width && j < cam.height)) chunk_end = chunk_next;
                }
                if (chunk_next == chunk_end) continue;
            }
            const uint32_t take = min(min(32u, n_free), chunk_end - chunk_next);
            n_free -= take;                                   // pop `take` slots from the end of the free list
            const bool active = lane < take;
            uint32_t slot = active ? S.list[WF_FREE][n_free + lane] : 0u;
            bool started = false;
            uint32_t first_kind = 0xffu;                      // camera ray answered from the pixel's candidate list: 0..2 material list
            if (active) {
                uint32_t r = chunk_next + lane;
                uint32_t pin = r / spp, sample = r - pin * spp + cam.sample_offset;
                uint32_t q = chunk_q0 + pin;
                uint32_t tile = (q >> 8) * P.world + P.rank, in = q & 255u;
                uint32_t ttx, tty;
                slot_tile(tile, P.tiles_x, &ttx, &tty);
                uint32_t i = ttx * kTileW + (in & 15u), j = tty * kTileH + (in >> 4);
                if (tile < P.tiles_total && i < cam.width && j < cam.height) {
                    uint32_t pixel = j * cam.width + i;
                    Stream<EXACT> rng(P.seed, pixel, sample, 0u, is_general<SC>::value && !EXACT);
                    Ray<T> ray = get_ray<T, EXACT>(cam, i, j, rng);
                    S.ox[slot] = ray.o.x; S.oy[slot] = ray.o.y; S.oz[slot] = ray.o.z;
                    S.dx[slot] = ray.d.x; S.dy[slot] = ray.d.y; S.dz[slot] = ray.d.z;
                    S.mx[slot] = 1.f; S.my[slot] = 1.f; S.mz[slot] = 1.f;
                    S.q[slot] = q; S.pix[slot] = pixel; S.smp[slot] = sample; S.dep[slot] = cam.max_depth;
                    started = true;
                    npaths++;
                    if constexpr (!GEN) {
                        if (P.cand) {
                            // the camera ray's closest hit from its pixel's candidate list (closest_prim_candidates): no EXTEND pass, and
                            // a path that sees only background (3 of 4 in `simple`) ends here without ever occupying its slot
                            const uint4 c = __ldg(P.cand + pixel);
                            if (c.x != kCandOverflow) {
                                nrays++;
                                T best_t; int32_t bestp;
                                if (closest_prim_candidates<COUNT>(sc, ray, P.tmin, M<T, EXACT>::inf(), c, &bestp, &best_t, tl)) {
                                    S.ht[slot] = best_t; S.hp[slot] = bestp;
                                    uint32_t k = bestp >= 0 ? (load_sphere_info(sc, bestp) & 3u) : (sc.planes[-2 - bestp].info & 3u);
                                    first_kind = k == LAMBERTIAN ? 0u : (k == METAL ? 1u : 2u);
                                } else {
                                    if (COUNT) tl.missed++;
                                    wf_finish(Q, acc, q, mk<T>(1.f, 1.f, 1.f) * cam.background + wf_res(cam.max_depth), P.flags);      // camera.rs:473-475
                                    first_kind = 3u;
                                }
                                started = false;
                            }
                        }
                    }
                }
            }
            chunk_next += take;
            __syncwarp();
            wf_push(S.list[WF_EXT], n_ext, started, slot, lt_mask);
            if constexpr (!GEN) {
                if (P.cand) {                                 // warp-uniform
                    wf_push(S.list[WF_LAMB], n_lamb, first_kind == 0u, slot, lt_mask);
                    wf_push(S.list[WF_METAL], n_metal, first_kind == 1u, slot, lt_mask);
                    wf_push(S.list[WF_DIEL], n_diel, first_kind == 2u, slot, lt_mask);
                }
            }
            wf_push(S.list[WF_FREE], n_free, active && !started && first_kind > 2u, slot, lt_mask);
        } else if (stage == WF_EXT) {
            // ---- EXTEND ------------------------------------------------------------------------------------
            const uint32_t n = min(32u, n_ext);
            n_ext -= n;
            const bool active = lane < n;
            uint32_t slot = active ? S.list[WF_EXT][n_ext + lane] : 0u;
            uint32_t kind = 0xffu;                            // 0..2 material list, 3 = miss (slot becomes free)
            // General scenes: a path that ends here hands its value to ONE wf_finish at the end of the stage (their kernels are
            // bound by instruction fetch, every inlined copy of the accumulate / flush code counts); the sphere kernels keep the
            // call in place (0.6 % faster that way).
            bool fin = false;
            V3<T> fin_value = mk<T>(0, 0, 0);
            if (active) {
                Ray<T> r{mk<T>(S.ox[slot], S.oy[slot], S.oz[slot]), mk<T>(S.dx[slot], S.dy[slot], S.dz[slot])};
                nrays++;
                // closest hit without the hit record: the winner's id and t are stored, SHADE builds the record
                T best_t; int32_t bestp;
                bool hit;
                if constexpr (GEN) {
                    uint32_t sub;
                    hit = g_closest_prim<T, EXACT, COUNT>(sc, r, P.tmin, M<T, EXACT>::inf(), &bestp, &sub, &best_t, stack, BLOCK, tl);
                    if (hit) {
                        S.ht[slot] = best_t; S.hp[slot] = bestp; S.hs[slot] = sub;
                        uint32_t k = sc.mats[g_entry<T>(sc, bestp)->mat].kind;
                        // DiffuseLight / Invisible never scatter (mult * emitted + res, camera.rs:484-486): they ride the first shade
                        // list, where g_shade returns V_ABSORB — a second inlined hit record here cost more in instruction fetch
                        // than the extra pass does
                        kind = k == METAL ? 1u : (k == DIELECTRIC ? 2u : 0u);
                    }
                } else {
                    hit = closest_prim<T, EXACT, COUNT, SC>(sc, r, P.tmin, M<T, EXACT>::inf(), &bestp, &best_t, stack, BLOCK, tl);
                    if (hit) {
                        S.ht[slot] = best_t; S.hp[slot] = bestp;
                        uint32_t k = bestp >= 0 ? (load_sphere_info(sc, bestp) & 3u) : (sc.planes[-2 - bestp].info & 3u);
                        kind = k == LAMBERTIAN ? 0u : (k == METAL ? 1u : 2u);
                    }
                }
                if (!hit) {
                    if (COUNT) tl.missed++;
                    V3<T> mult = mk<T>(S.mx[slot], S.my[slot], S.mz[slot]);
                    fin_value = mult * cam.background + wf_res(S.dep[slot]);                              // camera.rs:473-475
                    if constexpr (GEN) fin = true;
                    else wf_finish(Q, acc, S.q[slot], fin_value, P.flags);
                    kind = 3u;
                }
                if constexpr (GEN) { if (fin) wf_finish(Q, acc, S.q[slot], fin_value, P.flags); }
            }
            __syncwarp();
            wf_push(S.list[WF_LAMB], n_lamb, kind == 0u, slot, lt_mask);
            wf_push(S.list[WF_METAL], n_metal, kind == 1u, slot, lt_mask);
            wf_push(S.list[WF_DIEL], n_diel, kind == 2u, slot, lt_mask);
            wf_push(S.list[WF_FREE], n_free, kind == 3u, slot, lt_mask);
        } else if (CONN && stage == WF_CONN) {
            // ---- CONNECT: lights.pdf_value(dir) of Lambertian vertices (hittable_list.rs:408-412, pdf.rs:77-101) -----------------
            // 32 light-BVH walks side by side.  Walk lengths differ by orders of magnitude (a ray inside the slab that holds the
            // lights crosses hundreds of boxes, most rays a handful): a walk's state is (node, running sum), so when the batch has
            // thinned out the unfinished walks go back on the list and the next batch starts full again.
            if constexpr (CONN) {
            // Persistent batch, two walks per lane (the walk is bound by the latency of its dependent node loads: a second one in
            // flight hides half of it).  Every kConnQuantum steps the finished walks are completed — weight, mult, next stage — and
            // their lane slots refilled from the list IN PLACE, so a long walk stays on its lane while short ones come and go.  The
            // batch only lets go of unfinished walks (state: node index + running sum, back on the list) when the list is empty and
            // another stage has at least as many paths waiting as walks remain here.
            bool active[2] = {false, false}; uint32_t slot[2] = {0u, 0u}; int32_t cur[2] = {-1, -1}; float lacc[2] = {0.f, 0.f};
            V3<T> origin[2] = {mk<T>(0, 0, 0), mk<T>(0, 0, 0)}, nd[2] = {mk<T>(0, 0, 1), mk<T>(0, 0, 1)}; RayAux aux[2];
            ray_aux(Ray<float>{origin[0], nd[0]}, aux[0]); aux[1] = aux[0];
            for (;;) {
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    // ---- finish the walks that have ended ----
                    const bool done = active[k] && cur[k] < 0;
                    if (__any_sync(0xffffffffu, done)) {
                        bool to_ext = false, to_free = false;
                        const uint32_t sl = slot[k];
                        if (done) {
                            const int32_t hp = S.hp[sl];
                            V3<T> albedo;
                            if (hp >= 0) { Vec4T<T> m = load_sphere_mat(sc, hp); albedo = mk<T>(m.x, m.y, m.z); }
                            else { const PlaneT<T>& pl = sc.planes[-2 - hp]; albedo = mk<T>(pl.albedo[0], pl.albedo[1], pl.albedo[2]); }
                            const T cos_v = S.ht[sl];
                            const T light_v = lacc[k] * frcp((T)sc.n_lights);
                            V3<T> w = lambertian_weight<T, EXACT>(albedo, light_v, cos_v, cos_v);
                            Ray<T> next{origin[k], mk<T>(S.dx[sl], S.dy[sl], S.dz[sl])};
                            V3<T> fin_value;
                            if (wf_advance<COUNT>(S, sl, V_DIFFUSE, next, w, mk<T>(0, 0, 0), mk<T>(S.mx[sl], S.my[sl], S.mz[sl]), S.dep[sl], &fin_value, tl)) {
                                wf_finish(Q, acc, S.q[sl], fin_value, P.flags);
                                to_free = true;
                            } else to_ext = true;
                            active[k] = false;
                        }
                        __syncwarp();
                        wf_push(S.list[WF_EXT], n_ext, to_ext, sl, lt_mask);
                        wf_push(S.list[WF_FREE], n_free, to_free, sl, lt_mask);
                    }
                    // ---- refill idle lane slots from the list ----
                    const uint32_t idle = __ballot_sync(0xffffffffu, !active[k]);
                    const uint32_t take = min((uint32_t)__popc(idle), n_conn);
                    if (take) {
                        const uint32_t rank = __popc(idle & lt_mask);
                        n_conn -= take;
                        if (!active[k] && rank < take) {
                            const uint32_t sl = S.conn[n_conn + rank];
                            slot[k] = sl; active[k] = true;
                            origin[k] = mk<T>(S.ox[sl], S.oy[sl], S.oz[sl]);
                            nd[k] = M<T, EXACT>::normalize(mk<T>(S.dx[sl], S.dy[sl], S.dz[sl]));
                            cur[k] = S.lcur[sl]; lacc[k] = S.lacc[sl];
                            ray_aux(Ray<float>{origin[k], nd[k]}, aux[k]);
                        }
                    }
                }
                const uint32_t walking = __popc(__ballot_sync(0xffffffffu, cur[0] >= 0)) + __popc(__ballot_sync(0xffffffffu, cur[1] >= 0));
                if (walking == 0 && !__any_sync(0xffffffffu, active[0] || active[1])) break;
                if (walking != 0 && n_conn == 0) {
                    // nothing to refill with: yield if another stage has at least as many paths waiting as walks remain (free slots count
                    // exactly when the stage selection above would let GENERATE run — not on the queue's background-only tail, where
                    // old paths go first: yielding to a stage that is not going to run would bounce between here and there forever)
                    const uint32_t nf = ((exhausted && chunk_next == chunk_end) || cheap_phase) ? 0u : n_free;
                    const uint32_t other = max(max(n_ext, nf), max(n_lamb, max(n_metal, n_diel)));
                    if (other >= walking) {
                        // finished-but-not-yet-completed walks are completed by the next pass of the loop head; suspend the rest
                        bool again[2];
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            again[k] = active[k] && cur[k] >= 0;
                            if (again[k]) { S.lcur[slot[k]] = cur[k]; S.lacc[slot[k]] = lacc[k]; active[k] = false; cur[k] = -1; }
                        }
                        __syncwarp();
#pragma unroll
                        for (int k = 0; k < 2; ++k) wf_push(S.conn, n_conn, again[k], slot[k], lt_mask);
                        if (!__any_sync(0xffffffffu, active[0] || active[1])) break;
                        continue;                               // complete the walks that ended in the last quantum, then leave
                    }
                }
                if (walking != 0)
                    light_walk_pair<COUNT>(sc, origin[0], nd[0], aux[0], cur[0], lacc[0], origin[1], nd[1], aux[1], cur[1], lacc[1], kConnQuantum, tl);
            }
            }
        } else {
            // ---- SHADE: all lanes run the same material -------------------------------------------------------
            uint32_t cnt = stage == WF_LAMB ? n_lamb : (stage == WF_METAL ? n_metal : n_diel);
            const uint32_t n = min(32u, cnt);
            cnt -= n;
            if (stage == WF_LAMB) n_lamb = cnt; else if (stage == WF_METAL) n_metal = cnt; else n_diel = cnt;
            const bool active = lane < n;
            uint32_t slot = active ? S.list[stage][cnt + lane] : 0u;
            bool to_ext = false, to_free = false, to_conn = false;
            uint32_t again = 0xffu;                           // the next ray's closest hit is already known (closest_prim_self): 0..2 material list
            if (active) {
                Ray<T> r{mk<T>(S.ox[slot], S.oy[slot], S.oz[slot]), mk<T>(S.dx[slot], S.dy[slot], S.dz[slot])};
                Hit<T> h;
                if constexpr (GEN) g_hit_record<T, EXACT>(sc, r, g_entry<T>(sc, S.hp[slot]), S.hs[slot], S.ht[slot], &h);
                else hit_record<T, EXACT, SC>(sc, r, S.hp[slot], S.ht[slot], &h);
                uint32_t dep = S.dep[slot], depth = dep & 0xffffu;
                V3<T> mult = mk<T>(S.mx[slot], S.my[slot], S.mz[slot]);
                Stream<EXACT> rng(P.seed, S.pix[slot], S.smp[slot], cam.max_depth - depth + 1u, is_general<SC>::value && !EXACT);
                bool deferred = false;
                if constexpr (CONN) {
                    if (stage == WF_LAMB) {                     // warp-uniform: sample the direction here, the light term runs in CONNECT
                        if (COUNT) tl.lambertian++;
                        T cos_v, sp;
                        V3<T> dir = lambertian_sample<T, EXACT>(sc, h, rng, &cos_v, &sp);
                        S.ox[slot] = h.p.x; S.oy[slot] = h.p.y; S.oz[slot] = h.p.z;
                        S.dx[slot] = dir.x; S.dy[slot] = dir.y; S.dz[slot] = dir.z;
                        S.ht[slot] = sp;
                        S.lcur[slot] = light_walk_start(h.p, M<T, EXACT>::normalize(dir)); S.lacc[slot] = 0.f;
                        to_conn = true; deferred = true;
                    }
                }
                if (!deferred) {
                    Ray<T> next;
                    V3<T> w;
                    uint32_t kind = shade<T, EXACT, COUNT, SC>(sc, r, h, rng, &next, &w, tl);
                    V3<T> emitted = mk<T>(0, 0, 0);
                    if constexpr (GEN) emitted = g_emitted<T>(h);
                    V3<T> fin_value = mk<T>(0, 0, 0);
                    if (wf_advance<COUNT>(S, slot, kind, next, w, emitted, mult, dep, &fin_value, tl)) {
                        wf_finish(Q, acc, S.q[slot], fin_value, P.flags);
                        to_free = true;
                    } else {
                        to_ext = true;
                        if constexpr (!GEN) {
                            // the ray leaves an isolated sphere and meets it again: no EXTEND pass, straight back to a shade list
                            if (!(P.flags & 8u)) {            // RTW_FLAG_NO_CANDIDATES: every ray walks the tree
                                T best_t; int32_t bestp;
                                if (closest_prim_self<COUNT>(sc, next, S.hp[slot], P.tmin, M<T, EXACT>::inf(), &bestp, &best_t, tl)) {
                                    nrays++;
                                    S.ht[slot] = best_t; S.hp[slot] = bestp;
                                    uint32_t k = bestp >= 0 ? (load_sphere_info(sc, bestp) & 3u) : (sc.planes[-2 - bestp].info & 3u);
                                    again = k == LAMBERTIAN ? 0u : (k == METAL ? 1u : 2u);
                                    to_ext = false;
                                }
                            }
                        }
                    }
                }
            }
            __syncwarp();
            if constexpr (CONN) wf_push(S.conn, n_conn, to_conn, slot, lt_mask);
            wf_push(S.list[WF_EXT], n_ext, to_ext, slot, lt_mask);
            if constexpr (!GEN) {
                wf_push(S.list[WF_LAMB], n_lamb, again == 0u, slot, lt_mask);
                wf_push(S.list[WF_METAL], n_metal, again == 1u, slot, lt_mask);
                wf_push(S.list[WF_DIEL], n_diel, again == 2u, slot, lt_mask);
            }
            wf_push(S.list[WF_FREE], n_free, to_free, slot, lt_mask);
        }
    }
    wf_acc_flush(Q, acc);
#ifdef RTW_TIMELINE
    tl_t[3] = wf_now();
    if (lane == 0) {
        unsigned long long* o = rtw_timeline + ((size_t)blockIdx.x * 32 + warp) * 8;
        o[0] = tl_t[0]; o[1] = tl_t[1]; o[2] = tl_t[2]; o[3] = tl_t[3]; o[4] = tl_inflight; o[5] = tl_passes; o[6] = npaths; o[7] = nrays;
    }
#endif
    flush_counters<COUNT>(P.counters, npaths, nrays, tl);
}

template <int BLOCK, int NPW> size_t wavefront_state_bytes() { return sizeof(WfWarp<NPW>) * (BLOCK / 32); }
template <int BLOCK, int NPW> size_t wavefront_state_bytes_general() { return sizeof(WfWarpG<NPW>) * (BLOCK / 32); }
template <int BLOCK, int NPW> size_t wavefront_state_bytes_connect() { return sizeof(WfWarpC<NPW>) * (BLOCK / 32); }

}  // namespace rtw
