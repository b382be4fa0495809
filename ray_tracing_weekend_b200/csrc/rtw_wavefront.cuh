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

enum WfStage : int { WF_FREE = 0, WF_EXT = 1, WF_LAMB = 2, WF_METAL = 3, WF_DIEL = 4, WF_STAGES = 5 };

template <int NPW> struct WfWarp {
    float ox[NPW], oy[NPW], oz[NPW], dx[NPW], dy[NPW], dz[NPW], mx[NPW], my[NPW], mz[NPW], ht[NPW];
    uint32_t q[NPW], pix[NPW], smp[NPW], dep[NPW];  // dep = depth | res-is-NaN bits << 16
    int32_t hp[NPW];                                 // hit primitive: >= 0 sorted sphere, <= -2 plane
    uint8_t list[WF_STAGES][NPW];
};

// general scenes (rtw_general.cuh) also remember which quad of the winning entry was hit
template <int NPW> struct WfWarpG : WfWarp<NPW> { uint32_t hs[NPW]; };

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
    pool_add(A.a0, pool_fixed(value.x, 0, A.bad), 0, A.bad);
    pool_add(A.a1, pool_fixed(value.y, 1, A.bad), 1, A.bad);
    pool_add(A.a2, pool_fixed(value.z, 2, A.bad), 2, A.bad);
}

template <bool COUNT, int BLOCK, int NPW, bool SH, class SCENE = SceneView<float>>
__global__ void __launch_bounds__(BLOCK, 1) render_wavefront_kernel(RenderParams<float, SCENE> P, PoolParams Q) {
    using T = float;
    constexpr bool EXACT = false;
    constexpr bool GEN = is_general<SCENE>::value;        // general scenes: entries of any kind, scene tables in global memory
    static_assert(NPW <= 255 && NPW >= 32, "slot indices are stored in one byte");
    static_assert(!(GEN && SH), "general scenes are read from global memory");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: [stack][scene sections][one WfWarp per warp]
    const uint32_t stack_depth = P.stack_depth;
    int32_t* stack_base = reinterpret_cast<int32_t*>(smem_raw);
    unsigned char* cur_p = smem_raw + sizeof(int32_t) * stack_depth * BLOCK;
    using SC = typename std::conditional<SH, SceneViewSh<T>, SCENE>::type;
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
    using WS = typename std::conditional<GEN, WfWarpG<NPW>, WfWarp<NPW>>::type;
    WS& S = reinterpret_cast<WS*>(cur_p)[warp];

    const CameraT<T>& cam = P.cam;
    int32_t* stack = stack_base + tid;
    const uint32_t n_slots = P.n_local_tiles * (kTileW * kTileH);
    const uint32_t spp = cam.spp, G = Q.pixels_per_chunk;
    uint32_t npaths = 0, nrays = 0;
    Tally tl;
    WfAcc acc;

    // warp-uniform state (registers): list lengths and the cursor into the path stream
    uint32_t n_free = NPW, n_ext = 0, n_lamb = 0, n_metal = 0, n_diel = 0;
    uint32_t chunk_next = 0, chunk_end = 0, chunk_q0 = 0;
    bool exhausted = (spp == 0 || cam.max_depth == 0);       // max_depth == 0: every path returns 0 (camera.rs:470-472)
    for (uint32_t i = lane; i < NPW; i += 32) S.list[WF_FREE][i] = (uint8_t)i;

    for (;;) {
        __syncwarp();
        // ---- pick the stage with the longest list (free slots only count while the stream has paths left) ----
        uint32_t nf = (exhausted && chunk_next == chunk_end) ? 0u : n_free;
        int stage = WF_FREE;
        uint32_t best = nf;
        if (n_ext > best) { best = n_ext; stage = WF_EXT; }
        if (n_lamb > best) { best = n_lamb; stage = WF_LAMB; }
        if (n_metal > best) { best = n_metal; stage = WF_METAL; }
        if (n_diel > best) { best = n_diel; stage = WF_DIEL; }
        if (best == 0) break;                                 // nothing in flight and the stream is dry

        if (stage == WF_FREE) {
            // ---- GENERATE ----------------------------------------------------------------------------------
            if (chunk_next == chunk_end) {
                uint32_t c = 0;
                if (lane == 0) c = atomicAdd(P.work_counter, 1u);
                c = __shfl_sync(0xffffffffu, c, 0);
                if (c >= Q.n_chunks) { exhausted = true; continue; }
                chunk_q0 = c * G;
                uint32_t npx = min(G, n_slots - chunk_q0);
                chunk_next = 0; chunk_end = npx * spp;
                if (G == 1) {       // skip padding pixels (outside the image / padding tiles) as a whole
                    uint32_t tile = (chunk_q0 >> 8) * P.world + P.rank, in = chunk_q0 & 255u;
                    uint32_t ttx, tty;
                    slot_tile(tile, P.tiles_x, &ttx, &tty);
                    uint32_t i = ttx * kTileW + (in & 15u), j = tty * kTileH + (in >> 4);
                    if (!(tile < P.tiles_total && i < cam.width && j < cam.height)) chunk_end = 0;
                }
                if (chunk_next == chunk_end) continue;
            }
            const uint32_t take = min(min(32u, n_free), chunk_end - chunk_next);
            n_free -= take;                                   // pop `take` slots from the end of the free list
            const bool active = lane < take;
            uint32_t slot = active ? S.list[WF_FREE][n_free + lane] : 0u;
            bool started = false;
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
                }
            }
            chunk_next += take;
            __syncwarp();
            wf_push(S.list[WF_EXT], n_ext, started, slot, lt_mask);
            wf_push(S.list[WF_FREE], n_free, active && !started, slot, lt_mask);
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
                        uint32_t k = sc.mats[g_entry<T>(sc, bestp).mat].kind;
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
        } else {
            // ---- SHADE: all lanes run the same material -------------------------------------------------------
            uint32_t cnt = stage == WF_LAMB ? n_lamb : (stage == WF_METAL ? n_metal : n_diel);
            const uint32_t n = min(32u, cnt);
            cnt -= n;
            if (stage == WF_LAMB) n_lamb = cnt; else if (stage == WF_METAL) n_metal = cnt; else n_diel = cnt;
            const bool active = lane < n;
            uint32_t slot = active ? S.list[stage][cnt + lane] : 0u;
            bool to_ext = false, to_free = false;
            if (active) {
                Ray<T> r{mk<T>(S.ox[slot], S.oy[slot], S.oz[slot]), mk<T>(S.dx[slot], S.dy[slot], S.dz[slot])};
                Hit<T> h;
                if constexpr (GEN) g_hit_record<T, EXACT>(sc, r, g_entry<T>(sc, S.hp[slot]), S.hs[slot], S.ht[slot], &h);
                else hit_record<T, EXACT, SC>(sc, r, S.hp[slot], S.ht[slot], &h);
                uint32_t dep = S.dep[slot], depth = dep & 0xffffu;
                V3<T> mult = mk<T>(S.mx[slot], S.my[slot], S.mz[slot]);
                Stream<EXACT> rng(P.seed, S.pix[slot], S.smp[slot], cam.max_depth - depth + 1u, is_general<SC>::value && !EXACT);
                Ray<T> next;
                V3<T> w;
                uint32_t kind = shade<T, EXACT, COUNT, SC>(sc, r, h, rng, &next, &w, tl, stack, BLOCK);
                V3<T> emitted = mk<T>(0, 0, 0);
                if constexpr (GEN) emitted = g_emitted<T>(h);
                V3<T> fin_value = mk<T>(0, 0, 0);
                if (kind == V_ABSORB) {                                     // camera.rs:484-486
                    fin_value = mult * emitted + wf_res(dep);
                    if constexpr (!GEN) wf_finish(Q, acc, S.q[slot], fin_value, P.flags);
                    to_free = true;
                } else {
                    if (kind == V_DIFFUSE) {                                // res + mult * emitted (camera.rs:519): 0 or NaN per channel
                        V3<T> rs = wf_res(dep) + mult * emitted;
                        dep |= (rs.x != rs.x ? 1u << 16 : 0u) | (rs.y != rs.y ? 1u << 17 : 0u) | (rs.z != rs.z ? 1u << 18 : 0u);
                    }
                    mult = mult * w;
                    depth -= 1;
                    dep = (dep & 0xffff0000u) | depth;
                    if (depth == 0) {                                       // camera.rs:470-472
                        if (COUNT) tl.depth_out++;
                        fin_value = mk<T>(0, 0, 0) + wf_res(dep);
                        if constexpr (!GEN) wf_finish(Q, acc, S.q[slot], fin_value, P.flags);
                        to_free = true;
                    } else {
                        S.ox[slot] = next.o.x; S.oy[slot] = next.o.y; S.oz[slot] = next.o.z;
                        S.dx[slot] = next.d.x; S.dy[slot] = next.d.y; S.dz[slot] = next.d.z;
                        S.mx[slot] = mult.x; S.my[slot] = mult.y; S.mz[slot] = mult.z;
                        S.dep[slot] = dep;
                        to_ext = true;
                    }
                }
                if constexpr (GEN) { if (to_free) wf_finish(Q, acc, S.q[slot], fin_value, P.flags); }
            }
            __syncwarp();
            wf_push(S.list[WF_EXT], n_ext, to_ext, slot, lt_mask);
            wf_push(S.list[WF_FREE], n_free, to_free, slot, lt_mask);
        }
    }
    wf_acc_flush(Q, acc);
    flush_counters<COUNT>(P.counters, npaths, nrays, tl);
}

template <int BLOCK, int NPW> size_t wavefront_state_bytes() { return sizeof(WfWarp<NPW>) * (BLOCK / 32); }
template <int BLOCK, int NPW> size_t wavefront_state_bytes_general() { return sizeof(WfWarpG<NPW>) * (BLOCK / 32); }

}  // namespace rtw
