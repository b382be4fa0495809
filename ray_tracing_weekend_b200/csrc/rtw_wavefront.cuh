// rtw_wavefront.cuh — the wavefront renderer (RTW_WAVEFRONT, fast path only).
//
// generate / extend / shade queues, compacted with warp ballots and sorted by material — but CTA-local:
// every persistent CTA keeps NP paths in flight in SHARED memory (SoA path state + index lists per
// stage) and steps all of them one bounce per iteration:
//
//   GENERATE  free slots take the next paths of the CTA's chunk of the pixel-major path stream
//             (Camera::get_ray, camera.rs:274-293)                                   -> extend list
//   EXTEND    closest hit (bvh.rs:163-188): every lane pulls rays from the extend list one by one, so
//             a lane whose traversal ends early starts the next ray instead of idling
//             miss -> the path ends (camera.rs:473-475); hit -> list of its material
//   SHADE     one list per material (Lambertian + light-pdf loop / Metal / Dialectric), so a warp runs
//             one Material::scatter (material.rs:357-488)                             -> extend list
//
// Why CTA-local: a global-memory wavefront writes and re-reads ~130 B of path state per bounce
// (~380 GB per 1080p/500spp frame) and needs thousands of launches; the same queues in the 227 KB of
// shared memory cost neither.  The megakernel runs these stages per lane in lock step and leaves
// ~60 % of the lanes idle (ncu: 12.6 of 32 threads active); here lanes only ever execute a stage
// together with lanes that need the same stage.
// The paths, their RNG streams and the arithmetic are exactly the pooled megakernel's, and radiance is
// accumulated in the same 64-bit fixed point, so both renderers produce bit-identical images.
#pragma once
#include "rtw_kernels.cuh"

namespace rtw {

struct WfSegment { uint32_t q0, start, count, offset; };

template <int NP> struct WfLists {
    uint16_t free_[2][NP], ext[2][NP], lamb[NP], metal[NP], diel[NP];
    uint32_t n_free[2], n_ext[2], n_lamb, n_metal, n_diel, cursor;
    // path stream
    uint32_t chunk_next, chunk_end, chunk_q0, exhausted, n_gen, n_seg, done;
    WfSegment seg[4];
};

template <int NP> struct WfPaths {
    float ox[NP], oy[NP], oz[NP], dx[NP], dy[NP], dz[NP], mx[NP], my[NP], mz[NP], ht[NP];
    uint32_t q[NP], pix[NP], smp[NP], dep[NP];      // dep = depth | res-is-NaN bits << 16
    int32_t hp[NP];                                  // hit primitive: >= 0 sorted sphere, <= -2 plane
};

// warp-aggregated push of `idx` onto a shared-memory list by the lanes with pred set
RTW_D void wf_push(bool pred, uint16_t* list, uint32_t* count, uint32_t idx) {
    uint32_t m = __ballot_sync(__activemask(), pred);
    if (!pred) return;
    uint32_t lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(count, (uint32_t)__popc(m));
    base = __shfl_sync(m, base, leader);
    list[base + __popc(m & ((1u << lane) - 1u))] = (uint16_t)idx;
}

// a finished path: add its radiance to the pixel's fixed-point accumulators
RTW_D void wf_finish(const PoolParams& Q, uint32_t q, V3<float> value, uint32_t flags) {
    if (flags & 1u) value = fix_nan(value);
    uint32_t bad = 0;
    unsigned long long a0 = pool_fixed(value.x, 0, bad), a1 = pool_fixed(value.y, 1, bad), a2 = pool_fixed(value.z, 2, bad);
    pool_flush(Q, q, a0, a1, a2);
    if (bad) atomicOr(Q.poison + q, bad);
}

RTW_D V3<float> wf_res(uint32_t dep) {
    const float qnan = __int_as_float(0x7fc00000);
    return mk<float>((dep >> 16) & 1u ? qnan : 0.f, (dep >> 17) & 1u ? qnan : 0.f, (dep >> 18) & 1u ? qnan : 0.f);
}

template <bool COUNT, int BLOCK, int NP, bool SH>
__global__ void __launch_bounds__(BLOCK, 2) render_wavefront_kernel(RenderParams<float> P, PoolParams Q) {
    const uint32_t stack_depth = P.stack_depth;
    using T = float;
    constexpr bool EXACT = false;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: [stack][scene sections][paths][lists]
    int32_t* stack_base = reinterpret_cast<int32_t*>(smem_raw);
    SceneView<T> sc0 = P.scene;
    unsigned char* cur_p = smem_raw + sizeof(int32_t) * stack_depth * BLOCK;
    {
        if (P.smem_nodes) {
            uint4* dst = reinterpret_cast<uint4*>(cur_p);
            const uint4* src = reinterpret_cast<const uint4*>(P.scene.nodes);
            for (uint32_t i = threadIdx.x; i < P.smem_nodes / 16; i += BLOCK) dst[i] = src[i];
            sc0.top_nodes = reinterpret_cast<const Node<T>*>(cur_p);
            sc0.n_top = (int32_t)(P.smem_nodes / sizeof(Node<T>));
            cur_p += P.smem_nodes;
        }
        if (P.smem_spheres) {
            uint32_t n16 = P.smem_spheres / 16;
            uint4* dst = reinterpret_cast<uint4*>(cur_p);
            const uint4* src = reinterpret_cast<const uint4*>(P.scene.spheres);
            for (uint32_t i = threadIdx.x; i < n16; i += BLOCK) dst[i] = src[i];
            sc0.spheres = reinterpret_cast<const Vec4T<T>*>(cur_p);
            cur_p += P.smem_spheres;
            dst = reinterpret_cast<uint4*>(cur_p);
            src = reinterpret_cast<const uint4*>(P.scene.sphere_mat);
            for (uint32_t i = threadIdx.x; i < n16; i += BLOCK) dst[i] = src[i];
            sc0.sphere_mat = reinterpret_cast<const Vec4T<T>*>(cur_p);
            cur_p += P.smem_spheres;
            uint32_t* dsti = reinterpret_cast<uint32_t*>(cur_p);
            for (uint32_t i = threadIdx.x; i < (uint32_t)P.scene.n_spheres; i += BLOCK) dsti[i] = P.scene.sphere_info[i];
            sc0.sphere_info = dsti;
            cur_p += (P.scene.n_spheres * 4 + 15) / 16 * 16;
        }
        if (P.smem_lights) {
            uint4* dst = reinterpret_cast<uint4*>(cur_p);
            const uint4* src = reinterpret_cast<const uint4*>(P.scene.lights);
            for (uint32_t i = threadIdx.x; i < P.smem_lights / 16; i += BLOCK) dst[i] = src[i];
            sc0.lights = reinterpret_cast<const Vec4T<T>*>(cur_p);
            cur_p += P.smem_lights;
        }
    }
    using SC = typename std::conditional<SH, SceneViewSh<T>, SceneView<T>>::type;
    SC sc;
    static_cast<SceneView<T>&>(sc) = sc0;
    WfPaths<NP>& S = *reinterpret_cast<WfPaths<NP>*>(cur_p);
    WfLists<NP>& L = *reinterpret_cast<WfLists<NP>*>(cur_p + sizeof(WfPaths<NP>));

    const CameraT<T>& cam = P.cam;
    const uint32_t tid = threadIdx.x, lane = tid & 31, lt_mask = (1u << lane) - 1u;
    int32_t* stack = stack_base + tid;
    const uint32_t n_slots = P.n_local_tiles * (kTileW * kTileH);
    const uint32_t spp = cam.spp, G = Q.pixels_per_chunk;
    uint32_t npaths = 0, nrays = 0;
    Tally tl;

    for (uint32_t i = tid; i < NP; i += BLOCK) L.free_[0][i] = (uint16_t)i;
    if (tid == 0) {
        L.n_free[0] = NP; L.n_free[1] = 0; L.n_ext[0] = L.n_ext[1] = 0; L.n_lamb = L.n_metal = L.n_diel = 0; L.cursor = 0;
        L.chunk_next = L.chunk_end = L.chunk_q0 = 0; L.exhausted = (spp == 0 || cam.max_depth == 0) ? 1u : 0u; L.done = 0;
    }
    // max_depth == 0: every path returns 0 (camera.rs:470-472) — the accumulators are already zero
    uint32_t cb = 0;                                      // parity of the current list buffers
    for (;;) {
        __syncthreads();
        // ---- plan GENERATE: thread 0 maps the free slots onto segments of the path stream -----------------
        if (tid == 0) {
            uint32_t need = L.n_free[cb], nseg = 0, off = 0;
            while (need > 0 && nseg < 4) {
                if (L.chunk_next == L.chunk_end) {
                    if (L.exhausted) break;
                    uint32_t c = atomicAdd(P.work_counter, 1u);
                    if (c >= Q.n_chunks) { L.exhausted = 1; break; }
                    L.chunk_q0 = c * G;
                    uint32_t npx = min(G, n_slots - L.chunk_q0);
                    L.chunk_next = 0; L.chunk_end = npx * spp;
                    if (G == 1) {       // skip padding pixels as a whole
                        uint32_t tile = (L.chunk_q0 >> 8) * P.world + P.rank, in = L.chunk_q0 & 255u;
                        uint32_t ttx, tty;
                        slot_tile(tile, P.tiles_x, &ttx, &tty);
                        uint32_t i = ttx * kTileW + (in & 15u), j = tty * kTileH + (in >> 4);
                        if (!(tile < P.tiles_total && i < cam.width && j < cam.height)) L.chunk_end = 0;
                    }
                    continue;
                }
                uint32_t take = min(need, L.chunk_end - L.chunk_next);
                L.seg[nseg].q0 = L.chunk_q0; L.seg[nseg].start = L.chunk_next; L.seg[nseg].count = take; L.seg[nseg].offset = off;
                nseg++; off += take; L.chunk_next += take; need -= take;
            }
            L.n_seg = nseg; L.n_gen = off;
            L.cursor = 0;
        }
        __syncthreads();
        // ---- GENERATE ---------------------------------------------------------------------------------------
        {
            const uint32_t nfree = L.n_free[cb], ngen = L.n_gen, nseg = L.n_seg;
            for (uint32_t k0 = 0; k0 < nfree; k0 += BLOCK) {
                uint32_t k = k0 + tid;
                bool in_range = k < nfree;
                uint32_t slot = in_range ? L.free_[cb][k] : 0;
                bool started = false;
                if (in_range && k < ngen) {
                    uint32_t s = 0;
                    while (s + 1 < nseg && k >= L.seg[s + 1].offset) s++;
                    uint32_t r = L.seg[s].start + (k - L.seg[s].offset);
                    uint32_t pin = r / spp, sample = r - pin * spp;
                    uint32_t q = L.seg[s].q0 + pin;
                    uint32_t tile = (q >> 8) * P.world + P.rank, in = q & 255u;
                    uint32_t ttx, tty;
                        slot_tile(tile, P.tiles_x, &ttx, &tty);
                        uint32_t i = ttx * kTileW + (in & 15u), j = tty * kTileH + (in >> 4);
                    if (tile < P.tiles_total && i < cam.width && j < cam.height) {
                        uint32_t pixel = j * cam.width + i;
                        Stream<EXACT> rng(P.seed, pixel, sample, 0u);
                        Ray<T> ray = get_ray<T, EXACT>(cam, i, j, rng);
                        S.ox[slot] = ray.o.x; S.oy[slot] = ray.o.y; S.oz[slot] = ray.o.z;
                        S.dx[slot] = ray.d.x; S.dy[slot] = ray.d.y; S.dz[slot] = ray.d.z;
                        S.mx[slot] = 1.f; S.my[slot] = 1.f; S.mz[slot] = 1.f;
                        S.q[slot] = q; S.pix[slot] = pixel; S.smp[slot] = sample; S.dep[slot] = cam.max_depth;
                        started = true;
                        npaths++;
                    }
                }
                wf_push(started, L.ext[cb], &L.n_ext[cb], slot);
                wf_push(in_range && !started, L.free_[cb ^ 1], &L.n_free[cb ^ 1], slot);
            }
        }
        __syncthreads();
        if (L.n_ext[cb] == 0 && L.exhausted) break;      // nothing in flight and the stream is dry (uniform)
        // ---- EXTEND -----------------------------------------------------------------------------------------
        {
            const uint32_t n = L.n_ext[cb];
            const uint16_t* list = L.ext[cb];
            bool have = false;
            uint32_t slot = 0;
            Ray<T> r;
            RayAux aux;
            float inv_a = 0.f, best_t = 0.f;
            int32_t best = -1, cur = kStop;
            int sp = 0;
            const float tmin = P.tmin, tmax = M<T, EXACT>::inf();
            for (;;) {
                uint32_t want = __ballot_sync(0xffffffffu, !have);
                if (want) {
                    uint32_t base = 0;
                    if (lane == 0) base = atomicAdd(&L.cursor, (uint32_t)__popc(want));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (!have) {
                        uint32_t k = base + __popc(want & lt_mask);
                        if (k < n) {
                            slot = list[k];
                            r.o = mk<T>(S.ox[slot], S.oy[slot], S.oz[slot]);
                            r.d = mk<T>(S.dx[slot], S.dy[slot], S.dz[slot]);
                            have = true;
                            nrays++;
                            best = -1; best_t = tmax;
                            // planes: Plane::hit (entities/plane.rs:61-76), one-sided
                            for (int i = 0; i < sc.n_planes; ++i) {
                                const PlaneT<T>& pl = sc.planes[i];
                                T denom = dot(r.d, pl.normal);
                                if (!(denom > M<T, EXACT>::EPS)) continue;
                                T t = -dot(r.o - pl.point, pl.normal) * frcp(denom);
                                if (!(tmin <= t && t <= tmax)) continue;
                                if (best == -1 || t < best_t) { best_t = t; best = -2 - i; }
                            }
                            inv_a = frcp(sqlen(r.d));
                            ray_aux(r, aux);
                            stack[0] = kStop; sp = 1; cur = 0;
                        }
                    }
                }
                if (!__any_sync(0xffffffffu, have)) break;
                if (have) {
                    while (cur >= 0) {
                        Node<T> nd;
            load_node(sc, cur, nd);
                        if (COUNT) tl.node_visits++;
                        float tl_near, tr_near;
                        bool hl = box_hit_fast(nd.la, nd.lb, aux, tmin, best_t, &tl_near);
                        bool hr = box_hit_fast(nd.ra, nd.rb, aux, tmin, best_t, &tr_near);
                        int32_t l = nd.left, rr = nd.right;
                        if (hl && hr) {
                            bool swap = tr_near < tl_near;
                            stack[sp * BLOCK] = swap ? l : rr; sp++;
                            cur = swap ? rr : l;
                        } else if (hl) cur = l;
                        else if (hr) cur = rr;
                        else { sp--; cur = stack[sp * BLOCK]; }
                    }
                    if (cur == kStop) {
                        // traversal finished: miss ends the path (camera.rs:473-475), a hit is queued by material
                        have = false;
                        if (best == -1) {
                            if (COUNT) tl.missed++;
                            V3<T> mult = mk<T>(S.mx[slot], S.my[slot], S.mz[slot]);
                            wf_finish(Q, S.q[slot], mult * cam.background + wf_res(S.dep[slot]), P.flags);
                            uint32_t pos = atomicAdd(&L.n_free[cb ^ 1], 1u);
                            L.free_[cb ^ 1][pos] = (uint16_t)slot;
                        } else {
                            S.ht[slot] = best_t; S.hp[slot] = best;
                            uint32_t kind = best >= 0 ? (load_sphere_info(sc, best) & 3u) : (sc.planes[-2 - best].info & 3u);
                            uint16_t* dst = kind == LAMBERTIAN ? L.lamb : (kind == METAL ? L.metal : L.diel);
                            uint32_t* cnt = kind == LAMBERTIAN ? &L.n_lamb : (kind == METAL ? &L.n_metal : &L.n_diel);
                            uint32_t pos = atomicAdd(cnt, 1u);
                            dst[pos] = (uint16_t)slot;
                        }
                    } else {
                        if (cur != kEmptyLeaf) {
                            uint32_t enc = (uint32_t)~cur;
                            uint32_t first = enc >> 4, count = (enc & 15u) + 1u;
                            for (uint32_t i = first; i < first + count; ++i) {
                                Vec4T<T> s = load_sphere(sc, (int32_t)i);
                                if (COUNT) tl.sphere_tests++;
                                T t;
                                if (sphere_root_fast(s, r, inv_a, tmin, tmax, &t) && (best == -1 || t < best_t)) { best_t = t; best = (int32_t)i; }
                            }
                        }
                        sp--;
                        cur = stack[sp * BLOCK];
                    }
                }
            }
        }
        __syncthreads();
        // ---- SHADE: one list per material -------------------------------------------------------------------
#pragma unroll 1
        for (int m = 0; m < 3; ++m) {
            const uint16_t* list = m == 0 ? L.lamb : (m == 1 ? L.metal : L.diel);
            const uint32_t n = m == 0 ? L.n_lamb : (m == 1 ? L.n_metal : L.n_diel);
            for (uint32_t k0 = 0; k0 < n; k0 += BLOCK) {
                uint32_t k = k0 + tid;
                bool in_range = k < n;
                uint32_t slot = in_range ? list[k] : 0;
                bool to_ext = false, to_free = false;
                if (in_range) {
                    Ray<T> r{mk<T>(S.ox[slot], S.oy[slot], S.oz[slot]), mk<T>(S.dx[slot], S.dy[slot], S.dz[slot])};
                    Hit<T> h;
                    hit_record<T, EXACT, SC>(sc, r, S.hp[slot], S.ht[slot], &h);
                    uint32_t dep = S.dep[slot], depth = dep & 0xffffu;
                    V3<T> mult = mk<T>(S.mx[slot], S.my[slot], S.mz[slot]);
                    Stream<EXACT> rng(P.seed, S.pix[slot], S.smp[slot], cam.max_depth - depth + 1u);
                    Ray<T> next;
                    V3<T> w;
                    uint32_t kind = shade<T, EXACT, COUNT, SC>(sc, r, h, rng, &next, &w, tl, stack, BLOCK);
                    V3<T> emitted = mk<T>(0, 0, 0);
                    if (kind == V_ABSORB) {                                     // camera.rs:484-486
                        wf_finish(Q, S.q[slot], mult * emitted + wf_res(dep), P.flags);
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
                            wf_finish(Q, S.q[slot], mk<T>(0, 0, 0) + wf_res(dep), P.flags);
                            to_free = true;
                        } else {
                            S.ox[slot] = next.o.x; S.oy[slot] = next.o.y; S.oz[slot] = next.o.z;
                            S.dx[slot] = next.d.x; S.dy[slot] = next.d.y; S.dz[slot] = next.d.z;
                            S.mx[slot] = mult.x; S.my[slot] = mult.y; S.mz[slot] = mult.z;
                            S.dep[slot] = dep;
                            to_ext = true;
                        }
                    }
                }
                wf_push(to_ext, L.ext[cb ^ 1], &L.n_ext[cb ^ 1], slot);
                wf_push(to_free, L.free_[cb ^ 1], &L.n_free[cb ^ 1], slot);
            }
        }
        __syncthreads();
        if (tid == 0) { L.n_free[cb] = 0; L.n_ext[cb] = 0; L.n_lamb = L.n_metal = L.n_diel = 0; }
        cb ^= 1;
    }
    flush_counters<COUNT>(P.counters, npaths, nrays, tl);
}

template <int BLOCK, int NP> size_t wavefront_state_bytes() { return sizeof(WfPaths<NP>) + sizeof(WfLists<NP>); }

}  // namespace rtw
