// rtw_kernels.cuh — intersection, shading and the render / batch kernels, templated on the
// arithmetic policy (see rtw_device.cuh).  Included by kernels_f32.cu (fast) and kernels_f64.cu (exact).
#pragma once
#include <type_traits>
#include "rtw_device.cuh"

namespace rtw {

template <class T> struct Hit {
    V3<T> p, normal;       // HitRecord p / normal (shared/src/hittable.rs:102-129)
    T t;
    bool front_face;
    uint32_t info;         // prim_id << 2 | kind
    V3<T> albedo;
    T param;
    uint32_t gkind;        // general scenes only: the full material kind (DiffuseLight, Isotropic do not fit 2 bits)
};

struct Tally {             // per-thread event counters (only live when COUNT)
    uint32_t node_visits = 0, sphere_tests = 0, light_tests = 0, lambertian = 0, metal = 0, dielectric = 0,
             absorbed = 0, missed = 0, depth_out = 0;
};

// general scenes (rtw_general.cuh, included at the end of this file)
template <class T, bool EXACT, bool COUNT>
RTW_D bool g_closest_hit(const SceneViewG<T>& sc, const Ray<T>& r, T tmin, T tmax, Hit<T>* h, int32_t* stack, int stride, Tally& tl);
template <class T, bool EXACT, bool COUNT>
RTW_D uint32_t g_shade(const SceneViewG<T>& sc, const Ray<T>& r, const Hit<T>& h, Stream<EXACT>& rng, Ray<T>* next, V3<T>* weight, Tally& tl);
template <class T> RTW_D V3<T> g_emitted(const Hit<T>& h);

// ---------------------------------------------------------------------------------------------
// Ray / box.  Exact: AABoxHit for AABBox::hit (shared/src/hittable.rs:38-87) verbatim.
RTW_D bool box_hit_exact(const double* mn, const double* mx, const Ray<double>& r, double start, double end) {
    using Md = M<double, true>;
    double x_tmin = (mn[0] - r.o.x) / r.d.x, x_tmax = (mx[0] - r.o.x) / r.d.x;
    if (signbit(r.d.x)) { double s = x_tmin; x_tmin = x_tmax; x_tmax = s; }
    double tmin = x_tmin, tmax = x_tmax;
    double y_tmin = (mn[1] - r.o.y) / r.d.y, y_tmax = (mx[1] - r.o.y) / r.d.y;
    if (signbit(r.d.y)) { double s = y_tmin; y_tmin = y_tmax; y_tmax = s; }
    if (tmax < y_tmin || tmin > y_tmax) return false;
    tmin = Md::max_(tmin, y_tmin); tmax = Md::min_(tmax, y_tmax);
    double z_tmin = (mn[2] - r.o.z) / r.d.z, z_tmax = (mx[2] - r.o.z) / r.d.z;
    if (signbit(r.d.z)) { double s = z_tmin; z_tmin = z_tmax; z_tmax = s; }
    if (tmax < z_tmin || tmin > z_tmax) return false;
    tmin = Md::max_(tmin, z_tmin); tmax = Md::min_(tmax, z_tmax);
    return Md::max_(start, tmin) <= Md::min_(end, tmax);
}

// Fast: slabs from the box centre c and half-extent h.  Per axis m = (c - o) / d, e = h / |d|, the slab is
// [m - e, m + e]: three FMAs and NO per-axis min/max (ncu: the ALU pipe that executes FMNMX was the busiest
// pipe at 62 %, the FMA pipe at 34 %); entry / exit are one 3-input max / min each (FMNMX3).
// i = 1/d, oi = o/d, a = |1/d|.  NaNs (0 * inf) drop out of min/max: the axis becomes unconstrained — conservative.
struct RayAux { float ix, iy, iz, ox, oy, oz; };         // |1/d| is not kept: fabsf folds into the FMA's operand modifier
RTW_D float fmax3(float a, float b, float c) { float d; asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
RTW_D float fmin3(float a, float b, float c) { float d; asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
RTW_D void ray_aux(const Ray<float>& r, RayAux& a) {
    a.ix = frcp(r.d.x); a.iy = frcp(r.d.y); a.iz = frcp(r.d.z);
    a.ox = r.o.x * a.ix; a.oy = r.o.y * a.iy; a.oz = r.o.z * a.iz;
}
RTW_D bool box_hit_fast(const float* c, const float* h, const RayAux& a, float tmin, float tmax, float* tnear) {
    float mx = fmaf(c[0], a.ix, -a.ox), my = fmaf(c[1], a.iy, -a.oy), mz = fmaf(c[2], a.iz, -a.oz);
    float tn = fmax3(fmaf(-h[0], fabsf(a.ix), mx), fmaf(-h[1], fabsf(a.iy), my), fmaf(-h[2], fabsf(a.iz), mz));
    float tf = fmin3(fmaf(h[0], fabsf(a.ix), mx), fmaf(h[1], fabsf(a.iy), my), fmaf(h[2], fabsf(a.iz), mz));
    tn = fmaxf(tn, tmin);
    tf = fminf(tf, tmax);
    *tnear = tn;
    return tn <= tf;
}

// ---------------------------------------------------------------------------------------------
// Ray / sphere: Sphere::hit (shared/src/entities/sphere.rs:61-80).  Returns the accepted root.
// sphere_root: the reference's textbook quadratic, verbatim (used by the exact path in f64).
template <class T> RTW_D bool sphere_root(const Vec4T<T>& s, const Ray<T>& r, T a, T start, T end, T* t_out) {
    V3<T> oc = mk<T>(r.o.x - s.x, r.o.y - s.y, r.o.z - s.z);
    T half_b = dot(r.d, oc);
    T c = sqlen(oc) - s.w * s.w;
    T disc = half_b * half_b - a * c;
    if (!(disc > T(0))) return false;
    T sq = sqrt(disc);
    T root = (-half_b - sq) / a;
    if (!(start <= root && root <= end)) {
        root = (-half_b + sq) / a;
        if (!(start <= root && root <= end)) return false;
    }
    *t_out = root;
    return true;
}

// FP32 version of the same test.  hb^2 - a*c cancels catastrophically in FP32 for distant origins
// (|oc| >> r), so the discriminant is taken from the perpendicular residual l = oc - (hb/a) d
// (disc/a = r^2 - |l|^2, all terms of magnitude r), and the root that would cancel in -k -/+ s is
// obtained from the product of the roots (c/a)/q.  The sign of c = |oc|^2 - r^2 — which side of the
// surface the origin is on — decides the sign of the small root exactly as in the f64 reference.
// Root selection and the inclusive range test are the reference's.
RTW_D bool sphere_root_fast(const Vec4T<float>& s, const Ray<float>& r, float inv_a, float start, float end, float* t_out) {
    float ocx = r.o.x - s.x, ocy = r.o.y - s.y, ocz = r.o.z - s.z;
    float hb = fmaf(r.d.x, ocx, fmaf(r.d.y, ocy, r.d.z * ocz));
    float k = hb * inv_a;
    float lx = fmaf(-k, r.d.x, ocx), ly = fmaf(-k, r.d.y, ocy), lz = fmaf(-k, r.d.z, ocz);
    float r2 = s.w * s.w;
    float dq = fmaf(-lx, lx, fmaf(-ly, ly, fmaf(-lz, lz, r2)));   // disc / a
    if (!(dq > 0.f)) return false;
    float sq = fsqrt(dq * inv_a);                             // sqrt(disc) / a
    float c_a = fmaf(ocx, ocx, fmaf(ocy, ocy, fmaf(ocz, ocz, -r2))) * inv_a;
    float q = -(k + copysignf(sq, k));
    float other = c_a * frcp(q);
    float near_root = k < 0.f ? other : q, far_root = k < 0.f ? q : other;
    float root = near_root;
    if (!(start <= root && root <= end)) {
        root = far_root;
        if (!(start <= root && root <= end)) return false;
    }
    *t_out = root;
    return true;
}

// HitRecord::new (hittable.rs:102-129) for the winning primitive only.
// best >= 0: sorted sphere index; best <= -2: plane index -2 - best.
template <class T, bool EXACT, class SC>
RTW_D void hit_record(const SC& sc, const Ray<T>& r, int32_t best, T best_t, Hit<T>* h) {
    V3<T> outward;
    h->t = best_t;
    h->p = at(r, best_t);
    if (best >= 0) {
        Vec4T<T> s = load_sphere(sc, best);
        if constexpr (EXACT) outward = (h->p - mk<T>(s.x, s.y, s.z)) / s.w;            // sphere.rs:82-83
        else outward = (h->p - mk<T>(s.x, s.y, s.z)) * frcp(s.w);
        Vec4T<T> m = load_sphere_mat(sc, best);
        h->albedo = mk<T>(m.x, m.y, m.z); h->param = m.w; h->info = load_sphere_info(sc, best) & ~kSphereIsolated;
    } else {
        const PlaneT<T>& pl = sc.planes[-2 - best];
        outward = pl.normal;
        h->albedo = mk<T>(pl.albedo[0], pl.albedo[1], pl.albedo[2]); h->param = pl.param; h->info = pl.info;
    }
    h->front_face = dot(r.d, outward) < T(0);
    h->normal = h->front_face ? outward : -outward;
}

// ---------------------------------------------------------------------------------------------
// Closest hit: Hittable::hit of the world (bvh.rs:163-188 + hittable_list.rs:394-406), i.e.
// argmin-t over planes passing Plane::hit and spheres passing (own AABB test) && Sphere::hit.
// The reference visits both children with the un-shrunk range and keeps the first minimum; here the
// range is shrunk to the best t so far and children are visited near-first — the argmin is the same
// except for exact-t ties and the documented grazing cases.
// planes: Plane::hit (entities/plane.rs:61-76), one-sided; they have no finite box and are tested linearly
template <class T, bool EXACT, class SC>
RTW_D void closest_plane(const SC& sc, const Ray<T>& r, T tmin, T tmax, bool& found, T& best_t, int32_t& best) {
    using Mt = M<T, EXACT>;
    // not unrolled: the compiler's own choice (by 4, in each of the five inlined copies) made the wavefront kernel 3 144 instructions instead of
    // 2 552 for scenes that have one plane or none (-DRTW_PLANE_UNROLL=4 restores it)
#ifndef RTW_PLANE_UNROLL
#define RTW_PLANE_UNROLL 1
#endif
    constexpr int plane_unroll = RTW_PLANE_UNROLL;
#pragma unroll plane_unroll
    for (int i = 0; i < sc.n_planes; ++i) {
        const PlaneT<T>& pl = sc.planes[i];
        T denom = dot(r.d, pl.normal);
        if (!(denom > Mt::EPS)) continue;
        T t;
        if constexpr (EXACT) t = -dot(r.o - pl.point, pl.normal) / denom;
        else t = -dot(r.o - pl.point, pl.normal) * frcp(denom);
        if (!(tmin <= t && t <= tmax)) continue;
        if (!found || t < best_t) { found = true; best_t = t; best = -2 - i; }
    }
}

// Camera rays of a pinhole camera (defocus_angle <= EPSILON: every reference scene) all leave one point through one pixel's square:
// which spheres they can possibly hit is a property of the PIXEL, found once per frame by walking the BVH with the pixel's cone
// (primary_candidates_kernel) instead of once per sample with each ray.  The closest hit of a camera ray is then the argmin over
// its pixel's candidate list (<= 4 sorted sphere indices, kCandNone-terminated; kCandOverflow in .x: more than 4, walk the tree) and
// the planes — the same spheres pass the same sphere test, so the result is the traversal's; 75 % of the camera rays of `simple`
// have no candidate at all and end without touching the tree.
constexpr uint32_t kCandNone = 0xffffffffu, kCandOverflow = 0xfffffffeu;
template <bool COUNT, class SC>
RTW_D bool closest_prim_candidates(const SC& sc, const Ray<float>& r, float tmin, float tmax, uint4 cand, int32_t* best_out, float* t_out, Tally& tl) {
    bool found = false;
    float best_t = tmax;
    int32_t best = -1;
    closest_plane<float, false, SC>(sc, r, tmin, tmax, found, best_t, best);
    const float inv_a = frcp(sqlen(r.d));
#ifdef RTW_CAND_ROLL                       // tuning build (scripts/variant_bench.py): one inlined sphere test instead of four; slower, profiles/r2_code_size_combinations.jsonl
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
        const uint32_t id = k == 0 ? cand.x : (k == 1 ? cand.y : (k == 2 ? cand.z : cand.w));
        if (id == kCandNone) break;
        if (COUNT) tl.sphere_tests++;
        float t;
        if (sphere_root_fast(load_sphere(sc, (int32_t)id), r, inv_a, tmin, tmax, &t) && (!found || t < best_t)) { found = true; best_t = t; best = (int32_t)id; }
    }
#else
    const uint32_t ids[4] = {cand.x, cand.y, cand.z, cand.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (ids[k] == kCandNone) break;
        if (COUNT) tl.sphere_tests++;
        float t;
        if (sphere_root_fast(load_sphere(sc, (int32_t)ids[k]), r, inv_a, tmin, tmax, &t) && (!found || t < best_t)) { found = true; best_t = t; best = (int32_t)ids[k]; }
    }
#endif
    *best_out = best;
    *t_out = best_t;
    return found;
}

// A ray that LEAVES sphere `from` (the hit point of the previous vertex) and meets the same sphere again — self-intersection at the
// reference's tmin (about half the diffuse bounces), a refracted ray crossing its glass sphere, a zero-weight path
// bouncing on inside one — travels inside that sphere's (slightly inflated) ball from origin to hit point: both ends lie on its
// surface and a ball is convex.  If the sphere is ISOLATED (kSphereIsolated: no other sphere's surface comes near that ball, checked
// on the host with a margin far beyond FP32 rounding) no other sphere can be met first, so this root IS the tree walk's argmin: same
// function, same arguments, same bits.  Planes are unbounded and tested like in closest_prim.  Returns false when the walk is needed.
template <bool COUNT, class SC>
RTW_D bool closest_prim_self(const SC& sc, const Ray<float>& r, int32_t from, float tmin, float tmax, int32_t* best_out, float* t_out, Tally& tl) {
    if (from < 0 || !(load_sphere_info(sc, from) & kSphereIsolated)) return false;
    if (COUNT) tl.sphere_tests++;
    float t;
    if (!sphere_root_fast(load_sphere(sc, from), r, frcp(sqlen(r.d)), tmin, tmax, &t)) return false;
    bool found = false;
    float best_t = tmax;
    int32_t best = -1;
    closest_plane<float, false, SC>(sc, r, tmin, tmax, found, best_t, best);
    if (!found || t < best_t) { best_t = t; best = from; }
    *best_out = best;
    *t_out = best_t;
    return true;
}

template <class T, bool EXACT, bool COUNT, class SC>
RTW_D bool closest_prim(const SC& sc, const Ray<T>& r, T tmin, T tmax, int32_t* best_out, T* t_out, int32_t* stack, int stride, Tally& tl) {
    using Mt = M<T, EXACT>;
    bool found = false;
    T best_t = tmax;
    int32_t best = -1;      // >= 0 sorted sphere index; <= -2: plane ~index
    closest_plane<T, EXACT, SC>(sc, r, tmin, tmax, found, best_t, best);
    T a = sqlen(r.d);
    RayAux aux;
    float inv_a = 0.f;
    if constexpr (!EXACT) {
        inv_a = frcp(a);
        ray_aux(r, aux);
    }
    // while-while traversal: descend inner nodes until a leaf is reached, then test its spheres; a
    // stop code at the bottom of the stack ends the walk.
    stack[0] = kStop;
    int sp = 1;
    int32_t cur = 0;        // root inner node
    if constexpr (!EXACT) {
        // A ray with a non-finite origin or direction (NaN-poisoned paths keep bouncing) can hit no sphere:
        // the discriminant of sphere_root_fast is NaN or -inf for every sphere.  But its slab tests cull nothing either (NaNs drop
        // out of min / max), so it would visit EVERY node and sphere of the tree — 1.8 M tests per such ray on the 1 M-sphere scene,
        // up to 50 times per path, on a single lane.  Same result without the walk:
        const float probe = (r.o.x + r.o.y + r.o.z) + (r.d.x + r.d.y + r.d.z);
        if (!(fabsf(probe) <= 3.0e38f) && !(fabsf(r.o.x) <= 3.0e38f && fabsf(r.o.y) <= 3.0e38f && fabsf(r.o.z) <= 3.0e38f &&
                                             fabsf(r.d.x) <= 3.0e38f && fabsf(r.d.y) <= 3.0e38f && fabsf(r.d.z) <= 3.0e38f)) cur = kStop;
    }
    for (;;) {
        while (cur >= 0) {
            Node<T> nd;
            load_node(sc, cur, nd);
            if (COUNT) tl.node_visits++;
            bool hl, hr; T tl_near = T(0), tr_near = T(0);
            if constexpr (EXACT) {
                hl = box_hit_exact(nd.la, nd.lb, r, tmin, best_t);
                hr = box_hit_exact(nd.ra, nd.rb, r, tmin, best_t);
            } else {
                hl = box_hit_fast(nd.la, nd.lb, aux, tmin, best_t, &tl_near);
                hr = box_hit_fast(nd.ra, nd.rb, aux, tmin, best_t, &tr_near);
            }
            int32_t l = nd.left, rr = nd.right;
            if (hl && hr) {
                bool swap = !EXACT && tr_near < tl_near;
                int32_t near = swap ? rr : l, far = swap ? l : rr;
                stack[sp * stride] = far; sp++;
                cur = near;
            } else if (hl) {
                cur = l;
            } else if (hr) {
                cur = rr;
            } else {
                sp--;
                cur = stack[sp * stride];
            }
        }
        if (cur == kStop) break;
        if (cur != kEmptyLeaf) {
            uint32_t enc = (uint32_t)~cur;
            uint32_t first = enc >> 4, count = (enc & 15u) + 1u;
            for (uint32_t i = first; i < first + count; ++i) {
                Vec4T<T> s = load_sphere(sc, (int32_t)i);
                if constexpr (EXACT) {
                    // BoundedHittable::bounded_hit (hittable.rs:191-196): the sphere's own box first
                    // (Sphere::new's box, sphere.rs:42-45), with the un-shrunk range like the reference
                    double mn[3] = {s.x - s.w, s.y - s.w, s.z - s.w}, mx[3] = {s.x + s.w, s.y + s.w, s.z + s.w};
                    if (!box_hit_exact(mn, mx, r, tmin, tmax)) continue;
                }
                if (COUNT) tl.sphere_tests++;
                T t;
                bool hs;
                if constexpr (EXACT) hs = sphere_root<T>(s, r, a, tmin, tmax, &t);
                else hs = sphere_root_fast(s, r, inv_a, tmin, tmax, &t);
                if (hs && (!found || t < best_t)) { found = true; best_t = t; best = (int32_t)i; }
            }
        }
        sp--;
        cur = stack[sp * stride];
    }
    *best_out = best;
    *t_out = best_t;
    return found;
}

template <class T, bool EXACT, bool COUNT, class SC>
RTW_D bool closest_hit(const SC& sc, const Ray<T>& r, T tmin, T tmax, Hit<T>* h, int32_t* stack, int stride, Tally& tl) {
    if constexpr (is_general<SC>::value) return g_closest_hit<T, EXACT, COUNT>(sc, r, tmin, tmax, h, stack, stride, tl);
    else {
    int32_t best;
    T best_t;
    if (!closest_prim<T, EXACT, COUNT, SC>(sc, r, tmin, tmax, &best, &best_t, stack, stride, tl)) return false;
    hit_record<T, EXACT, SC>(sc, r, best, best_t, h);
    return true;
    }
}

// ---------------------------------------------------------------------------------------------
// geometry/src/onb.rs:8-35
template <class T, bool EXACT> struct Onb {
    V3<T> u, v, w;
    RTW_D explicit Onb(V3<T> n) {
        w = M<T, EXACT>::normalize(n);
        V3<T> a = fabs(w.x) > T(0.9) ? mk<T>(0, 1, 0) : mk<T>(1, 0, 0);
        v = M<T, EXACT>::normalize(cross(w, a));
        u = cross(w, v);
    }
    RTW_D V3<T> transform(V3<T> a) const {
        if constexpr (EXACT) return ((mk<T>(0, 0, 0) + u * a.x) + v * a.y) + w * a.z;
        else return mk<T>(fmaf(w.x, a.z, fmaf(v.x, a.y, u.x * a.x)), fmaf(w.y, a.z, fmaf(v.y, a.y, u.y * a.x)), fmaf(w.z, a.z, fmaf(v.z, a.y, u.z * a.x)));
    }
};

template <class T> RTW_D V3<T> reflect(V3<T> s, V3<T> o) { return s - (o * T(2)) * dot(s, o); }       // vec.rs:103-107
template <class T, bool EXACT> RTW_D V3<T> refract(V3<T> s, V3<T> o, T eta) {                         // vec.rs:109-116
    T cos_theta = M<T, EXACT>::min_(dot(s, -o), T(1));
    V3<T> perp = (s + o * cos_theta) * eta;
    V3<T> par = o * (-M<T, EXACT>::sqrt_(T(1) - sqlen(perp)));
    return perp + par;
}

// Sphere::pdf_value (sphere.rs:101-111) summed over the lights list (hittable_list.rs:408-412).
// Fast path: Sphere::hit(ray, 0..=inf) succeeds iff disc > 0 and the larger root is >= 0, i.e. iff
// disc > 0 && (hb <= 0 || c <= 0) — no square root or division until a light is actually hit.
// One light's term of the sum, fast path: unit direction nd; towards = nd.(c - o), perpendicular residual l = (c - o) - towards nd,
// hit iff |l|^2 < r^2 and (towards >= 0 or the origin is inside): 13 flop, no division or root on a miss.  s.w = r^2.
template <bool COUNT>
RTW_D void light_term(const Vec4T<float>& s, V3<float> origin, V3<float> nd, float& acc, Tally& tl) {
    using Mt = M<float, false>;
    if (COUNT) tl.light_tests++;
    float cx = s.x - origin.x, cy = s.y - origin.y, cz = s.z - origin.z;
    float towards = fmaf(nd.x, cx, fmaf(nd.y, cy, nd.z * cz));
    float lx = fmaf(-towards, nd.x, cx), ly = fmaf(-towards, nd.y, cy), lz = fmaf(-towards, nd.z, cz);
    float r2 = s.w;                                        // FP32 light records carry r^2 (upload_scene)
    if (fmaf(-lx, lx, fmaf(-ly, ly, fmaf(-lz, lz, r2))) > 0.f) {
        float distance_squared = fmaf(cx, cx, fmaf(cy, cy, cz * cz));
        if (towards >= 0.f || distance_squared <= r2) {
            float cos_theta_max = Mt::sqrt_(1.f - r2 * frcp(distance_squared));
            float solid_angle = 2.f * Mt::PI * (1.f - cos_theta_max);
            acc += frcp(solid_angle);
        }
    }
}

// Many lights (the reference's O(#lights) sum made sub-linear): stackless walk of the light BVH (LNode), evaluating only the
// lights whose box the ray (t in [0, inf)) crosses.  A light the ray misses contributes exactly 0 to the sum, so only the summation
// order changes — and that order (depth-first) is the same whoever runs the walk, in one go (megakernel, batch calls) or in
// pieces on different lanes (the wavefront's CONNECT stage), which keeps the renderers bit-identical.
// Runs at most max_steps node visits from `cur`; returns the node to resume at (-1: done).
// One node of the walk, already loaded (a, b = the record's two 16-byte halves): returns the next node.
template <bool COUNT, class SC>
RTW_D int32_t light_walk_step(const SC& sc, float4 a, float4 b, V3<float> origin, V3<float> nd, const RayAux& aux, int32_t cur, float& acc, Tally& tl) {
    const int32_t skip = __float_as_int(b.z), leaf = __float_as_int(b.w);
    if (leaf == kLNodeLight) {
        light_term<COUNT>(Vec4T<float>{a.x, a.y, a.z, a.w}, origin, nd, acc, tl);
        return skip;
    }
    const float c[3] = {a.x, a.y, a.z}, h[3] = {a.w, b.x, b.y};
    float tn;
    const bool hit = box_hit_fast(c, h, aux, 0.f, M<float, false>::inf(), &tn);
    if (hit && leaf >= 0) {
        const uint32_t first = (uint32_t)leaf >> 4, count = ((uint32_t)leaf & 15u) + 1u;
        for (uint32_t i = first; i < first + count; ++i) light_term<COUNT>(load_light(sc, (int32_t)i), origin, nd, acc, tl);
    }
    return (hit && leaf == kLNodeInner) ? cur + 1 : skip;
}
// Runs at most max_steps node visits from `cur`; returns the node to resume at (-1: done).
template <bool COUNT, class SC>
RTW_D int32_t light_walk(const SC& sc, V3<float> origin, V3<float> nd, const RayAux& aux, int32_t cur, float& acc, uint32_t max_steps, Tally& tl) {
    for (uint32_t step = 0; cur >= 0 && step < max_steps; ++step) {
        const F8 n = ldg256(sc.light_nodes + cur);
        cur = light_walk_step<COUNT>(sc, n.a, n.b, origin, nd, aux, cur, acc, tl);
    }
    return cur;
}
// Two walks in step: both nodes are requested before either is used, so a lane keeps two L2 round trips in flight (the walk is bound
// by the latency of these dependent loads: ncu, long_scoreboard 98 % on the first use of the node)
template <bool COUNT, class SC>
RTW_D void light_walk_pair(const SC& sc, V3<float> o0, V3<float> n0, const RayAux& x0, int32_t& c0, float& acc0,
                           V3<float> o1, V3<float> n1, const RayAux& x1, int32_t& c1, float& acc1, uint32_t max_steps, Tally& tl) {
    for (uint32_t step = 0; (c0 >= 0 || c1 >= 0) && step < max_steps; ++step) {
        // one 256-bit load per node (LDG.E.ENL2.256): the walk is bound by the L1/TEX pipe (ncu, C4: 76 % busy), which handles a
        // node's two 16-byte halves as two requests
        const F8 m0 = ldg256(sc.light_nodes + (c0 >= 0 ? c0 : 0)), m1 = ldg256(sc.light_nodes + (c1 >= 0 ? c1 : 0));
        if (c0 >= 0) c0 = light_walk_step<COUNT>(sc, m0.a, m0.b, o0, n0, x0, c0, acc0, tl);
        if (c1 >= 0) c1 = light_walk_step<COUNT>(sc, m1.a, m1.b, o1, n1, x1, c1, acc1, tl);
    }
}

// Where a light walk starts: at the root, or nowhere (-1) for a non-finite origin / direction — every light_term of such a ray
// is 0 (its discriminant is NaN), but its box tests cull nothing, so the walk would visit all the lights to add up zeros.
RTW_D int32_t light_walk_start(V3<float> origin, V3<float> nd) {
    const bool finite = fabsf(origin.x) <= 3.0e38f && fabsf(origin.y) <= 3.0e38f && fabsf(origin.z) <= 3.0e38f &&
                        fabsf(nd.x) <= 3.0e38f && fabsf(nd.y) <= 3.0e38f && fabsf(nd.z) <= 3.0e38f;
    return finite ? 0 : -1;
}

template <class T, bool EXACT, bool COUNT, class SC>
RTW_D T lights_pdf_value(const SC& sc, V3<T> origin, V3<T> dir, Tally& tl) {
    using Mt = M<T, EXACT>;
    T acc = T(0);
    if constexpr (EXACT) {
        T a = sqlen(dir);
        Ray<T> r{origin, dir};
        for (int i = 0; i < sc.n_lights; ++i) {
            Vec4T<T> s = load_light(sc, i);
            if (COUNT) tl.light_tests++;
            T t;
            T v = T(0);
            if (sphere_root<T>(s, r, a, T(0), Mt::inf(), &t)) {
                V3<T> cd = mk<T>(s.x - origin.x, s.y - origin.y, s.z - origin.z);
                T distance_squared = sqlen(cd);
                T cos_theta_max = Mt::sqrt_(T(1) - s.w * s.w / distance_squared);
                T solid_angle = T(2) * Mt::PI * (T(1) - cos_theta_max);
                v = T(1) / solid_angle;
            }
            acc = acc + v;
        }
        return acc / (T)sc.n_lights;
    } else {
        V3<T> nd = Mt::normalize(dir);
        constexpr int lm = light_mode<SC>::value;
        if (lm < 0 ? sc.n_light_nodes > 0 : lm == 1) {
            RayAux aux;
            ray_aux(Ray<float>{origin, nd}, aux);
            light_walk<COUNT>(sc, origin, nd, aux, light_walk_start(origin, nd), acc, 0xffffffffu, tl);
        } else if constexpr (lm == 0) {
            // (the wavefront without a light BVH) two passes: which lights does the ray's line cross — a bit per light, no branch — then the
            // terms of those lights, in the same order and from the same operands as light_term: the sum is bit for bit the one-pass loop's.
            // The one-pass loop enters its hit branch for one or two lanes at a time, ~12 times per warp and pass (ncu: 10 % of the kernel's
            // samples at 1.5 active lanes); here the lanes' hits are worked off together.  C2: 156.8 -> 154.4 ms (profiles/r2_light_defer_ab.jsonl)
            for (int base = 0; base < sc.n_lights; base += 32) {
                const int n = min(32, sc.n_lights - base);
                uint32_t m = 0;
#ifdef RTW_LIGHT_UNROLL                    // tuning build: unroll factor of the first pass (the compiler's own choice is 4); measured, ibid.
                constexpr int unroll = RTW_LIGHT_UNROLL;
#pragma unroll unroll
#endif
                for (int i = 0; i < n; ++i) {
                    const Vec4T<float> s = load_light(sc, base + i);
                    const float cx = s.x - origin.x, cy = s.y - origin.y, cz = s.z - origin.z;
                    const float towards = fmaf(nd.x, cx, fmaf(nd.y, cy, nd.z * cz));
                    const float lx = fmaf(-towards, nd.x, cx), ly = fmaf(-towards, nd.y, cy), lz = fmaf(-towards, nd.z, cz);
                    if (fmaf(-lx, lx, fmaf(-ly, ly, fmaf(-lz, lz, s.w))) > 0.f) m |= 1u << i;
                }
                if (COUNT) tl.light_tests += (uint32_t)n;
                while (m) {
                    const int i = __ffs((int)m) - 1;
                    m &= m - 1u;
                    const Vec4T<float> s = load_light(sc, base + i);
                    const float cx = s.x - origin.x, cy = s.y - origin.y, cz = s.z - origin.z;
                    const float towards = fmaf(nd.x, cx, fmaf(nd.y, cy, nd.z * cz));
                    const float distance_squared = fmaf(cx, cx, fmaf(cy, cy, cz * cz));
                    if (towards >= 0.f || distance_squared <= s.w) {
                        const float cos_theta_max = Mt::sqrt_(1.f - s.w * frcp(distance_squared));
                        const float solid_angle = 2.f * Mt::PI * (1.f - cos_theta_max);
                        acc += frcp(solid_angle);
                    }
                }
            }
        } else {
            for (int i = 0; i < sc.n_lights; ++i) light_term<COUNT>(load_light(sc, i), origin, nd, acc, tl);
        }
        return acc * frcp((T)sc.n_lights);
    }
}

// Sphere::random (sphere.rs:114-127)
template <class T, bool EXACT, bool W_IS_R2 = false>
RTW_D V3<T> sphere_random(const Vec4T<T>& s, V3<T> origin, Stream<EXACT>& rng) {
    using Mt = M<T, EXACT>;
    V3<T> direction = mk<T>(s.x - origin.x, s.y - origin.y, s.z - origin.z);
    T distance = Mt::sqrt_(sqlen(direction));
    Onb<T, EXACT> uvw(direction);
    T r1 = standard(rng);
    T r2 = standard(rng);
    const T radius_squared = W_IS_R2 ? s.w : s.w * s.w;    // the sphere path's FP32 light records carry r^2 (upload_scene)
    T z = T(1) + r1 * (Mt::sqrt_(T(1) - Mt::div(radius_squared, distance * distance)) - T(1));
    T sn, cs;
    Mt::sincos_2pi(r2, &sn, &cs);
    T x = cs * Mt::sqrt_(T(1) - z * z);
    T y = sn * Mt::sqrt_(T(1) - z * z);
    return uvw.transform(mk<T>(x, y, z));
}

enum VertexKind : uint32_t { V_MISS = 0, V_ABSORB = 1, V_SPECULAR = 2, V_DIFFUSE = 3 };

// The Lambertian branch of Material::scatter + ray_colour_tail_call in two halves, so that the wavefront can run the light term
// between them as a stage of its own (CONNECT): lambertian_sample draws the direction from the mixture pdf and evaluates the two
// cosine terms; lambertian_weight turns them and lights.pdf_value(dir) into the factor for `mult`.
template <class T, bool EXACT, class SC>
RTW_D V3<T> lambertian_sample(const SC& sc, const Hit<T>& h, Stream<EXACT>& rng, T* cos_v, T* scattering_pdf) {
    using Mt = M<T, EXACT>;
#ifndef RTW_LAMB_MERGE
#define RTW_LAMB_MERGE 0            // the light_mode the merged form is compiled for
#endif
    if constexpr (!EXACT && light_mode<SC>::value == RTW_LAMB_MERGE) {
        // (the wavefront without a light BVH) the two halves of the mixture share ONE copy of the frame, the sincos and the transform — the
        // kernel is bound by instruction fetch; each half computes what it computed before, from the same uniforms in the same order, so the
        // direction is bit for bit the two-copy form's.  C2 143.9 -> 133.0 ms; the light-BVH kernel is slower with it (C5 50.8 -> 52.3)
        V3<T> axis, dir;
        T phi, rad, z;
        if (standard(rng) < T(0.5)) {                       // Sphere::random (sphere.rs:114-127) of a light picked uniformly
            const uint32_t idx = uindex(rng, (uint32_t)sc.n_lights);
            const Vec4T<T> s = load_light(sc, (int32_t)idx);
            axis = mk<T>(s.x - h.p.x, s.y - h.p.y, s.z - h.p.z);
            const T distance = Mt::sqrt_(sqlen(axis));
            const T r1 = standard(rng);
            phi = standard(rng);
            z = T(1) + r1 * (Mt::sqrt_(T(1) - Mt::div(s.w, distance * distance)) - T(1));      // s.w = r^2 on the fast path
            rad = Mt::sqrt_(T(1) - z * z);
        } else {                                            // CosineWeightedHemisphere, utils.rs:146-161
            phi = standard(rng);
            const T r2 = standard(rng);
            axis = h.normal;
            rad = Mt::sqrt_(r2);
            z = Mt::sqrt_(T(1) - r2);
        }
        Onb<T, EXACT> uvw(axis);
        T sn, cs;
        Mt::sincos_2pi(phi, &sn, &cs);
        dir = uvw.transform(mk<T>(cs * rad, sn * rad, z));
        V3<T> nd = Mt::normalize(dir);
        *scattering_pdf = Mt::max_(Mt::div_pi(dot(h.normal, nd)), T(0));
        *cos_v = *scattering_pdf;
        return dir;
    }
    Onb<T, EXACT> uvw(h.normal);                            // CosinePdf::new, pdf.rs:39-43
    V3<T> dir;
    if (standard(rng) < T(0.5)) {                           // MixturePdf::generate, pdf.rs:94-100 (pdf1 = lights)
        uint32_t idx = uindex(rng, (uint32_t)sc.n_lights);
        dir = sphere_random<T, EXACT, !EXACT>(load_light(sc, (int32_t)idx), h.p, rng);
    } else {                                                // CosineWeightedHemisphere, utils.rs:146-161
        T r1 = standard(rng);
        T r2 = standard(rng);
        T sn, cs;
        Mt::sincos_2pi(r1, &sn, &cs);
        T x = cs * Mt::sqrt_(r2);
        T y = sn * Mt::sqrt_(r2);
        T z = Mt::sqrt_(T(1) - r2);
        dir = uvw.transform(mk<T>(x, y, z));
    }
    V3<T> nd = Mt::normalize(dir);
    *scattering_pdf = Mt::max_(Mt::div_pi(dot(h.normal, nd)), T(0));          // Lambertian::scattering_pdf, material.rs:372-375
    // CosinePdf::value, pdf.rs:46-49: the same cosine against uvw.w = normalize(h.normal).  The exact path evaluates it as the
    // reference does; the fast path uses the one value for both (h.normal is unit to FP32 precision), which also halves what a
    // suspended path has to carry through the CONNECT stage.
    if constexpr (EXACT) *cos_v = Mt::max_(Mt::div_pi(dot(nd, uvw.w)), T(0));
    else *cos_v = *scattering_pdf;
    return dir;
}
template <class T, bool EXACT>
RTW_D V3<T> lambertian_weight(V3<T> albedo, T light_v, T cos_v, T scattering_pdf) {
    T pdf_value = light_v * T(0.5) + cos_v * T(0.5);        // MixturePdf::value, pdf.rs:90-92
    if constexpr (EXACT) return (albedo * scattering_pdf) / pdf_value;       // camera.rs:518
    else return albedo * (scattering_pdf * frcp(pdf_value));
}

// Material::scatter (+ the Scatter branch of ray_colour_tail_call, camera.rs:484-521).
// Returns the vertex kind; on V_SPECULAR / V_DIFFUSE writes the next ray and the factor for `mult`.
// SPECULAR_ONLY: the caller has already dealt with Lambertian hits (g_shade borrows the Metal / Dielectric code from here).
template <class T, bool EXACT, bool COUNT, class SC, bool SPECULAR_ONLY = false>
RTW_D uint32_t shade(const SC& sc, const Ray<T>& r, const Hit<T>& h, Stream<EXACT>& rng, Ray<T>* next, V3<T>* weight, Tally& tl) {
    using Mt = M<T, EXACT>;
    if constexpr (is_general<SC>::value) return g_shade<T, EXACT, COUNT>(sc, r, h, rng, next, weight, tl);
    else {
    uint32_t kind = h.info & 3u;
    if (!SPECULAR_ONLY && kind == LAMBERTIAN) {                 // material.rs:357-376
        if (COUNT) tl.lambertian++;
        T cos_v, scattering_pdf;
        V3<T> dir = lambertian_sample<T, EXACT>(sc, h, rng, &cos_v, &scattering_pdf);
        T light_v = lights_pdf_value<T, EXACT, COUNT, SC>(sc, h.p, dir, tl);
        *next = Ray<T>{h.p, dir};
        *weight = lambertian_weight<T, EXACT>(h.albedo, light_v, cos_v, scattering_pdf);
        return V_DIFFUSE;
    }
    if (kind == METAL) {                                        // material.rs:407-421
        if (COUNT) tl.metal++;
        V3<T> reflected = reflect(Mt::normalize(r.d), h.normal);
        V3<T> ball;
        for (;;) {                                              // UnitSphere: uniform in the unit ball, utils.rs:99-122
            T a = T(2) * standard(rng) - T(1);
            T b = T(2) * standard(rng) - T(1);
            T c = T(2) * standard(rng) - T(1);
            ball = mk<T>(a, b, c);
            if (sqlen(ball) < T(1)) break;
        }
        V3<T> dir = reflected + ball * h.param;
        if (dot(dir, h.normal) > T(0)) { *next = Ray<T>{h.p, dir}; *weight = h.albedo; return V_SPECULAR; }
        if (COUNT) tl.absorbed++;
        return V_ABSORB;
    }
    if (kind == DIELECTRIC) {                                   // material.rs:457-488
        if (COUNT) tl.dielectric++;
        T ratio = h.front_face ? Mt::div(T(1), h.param) : h.param;
        V3<T> unit = Mt::normalize(r.d);
        T cos_theta = Mt::min_(dot(unit, -h.normal), T(1));
        T sin_theta = Mt::sqrt_(T(1) - cos_theta * cos_theta);
        bool do_reflect = ratio * sin_theta > T(1);
        if (!do_reflect) {
            T r0 = Mt::div(T(1) - ratio, T(1) + ratio);          // reflectance, material.rs:450-454
            r0 = r0 * r0;
            T om = T(1) - cos_theta;
            T p5 = ((om * om) * (om * om)) * om;
            do_reflect = (r0 + (T(1) - r0) * p5) > open01(rng);
        }
        V3<T> dir = do_reflect ? reflect(unit, h.normal) : refract<T, EXACT>(unit, h.normal, ratio);
        *next = Ray<T>{h.p, dir};
        *weight = mk<T>(1, 1, 1);
        return V_SPECULAR;
    }
    if (COUNT) tl.absorbed++;                                   // Invisible: Material defaults, material.rs:32-49
    return V_ABSORB;
    }
}

// Camera::get_ray (camera.rs:274-293)
template <class T, bool EXACT>
RTW_D Ray<T> get_ray(const CameraT<T>& cam, uint32_t i, uint32_t j, Stream<EXACT>& rng) {
    T ox = jitter(rng, cam.jitter_scale);
    T oy = jitter(rng, cam.jitter_scale);
    V3<T> pixel_sample = (cam.pixel00 + cam.du * ((T)i + ox)) + cam.dv * ((T)j + oy);
    V3<T> origin = cam.center;
    if (!(cam.defocus_angle <= M<T, EXACT>::EPS)) {
        T a, b;
        for (;;) {                                              // UnitDisk, utils.rs:124-144
            a = T(2) * standard(rng) - T(1);
            b = T(2) * standard(rng) - T(1);
            if (a * a + T(0) * T(0) + b * b < T(1)) break;
        }
        origin = (cam.center + cam.ddu * a) + cam.ddv * b;
    }
    return Ray<T>{origin, pixel_sample - origin};
}

// One step of ray_colour_tail_call (camera.rs:460-522) for a live path.  Returns true when the path
// finished and *value holds its radiance.
template <class T> struct PathState {
    Ray<T> r;
    V3<T> mult, res;
    uint32_t depth;
};

template <class T, bool EXACT, bool COUNT, class SC>
RTW_D bool path_step(const SC& sc, const CameraT<T>& cam, uint64_t seed, T tmin, uint32_t pixel, uint32_t sample,
                     PathState<T>& ps, V3<T>* value, int32_t* stack, int stride, uint32_t& nrays, Tally& tl, const uint4* cand = nullptr) {
    if (ps.depth == 0) {                                        // camera.rs:470-472
        if (COUNT) tl.depth_out++;
        *value = mk<T>(0, 0, 0) + ps.res;
        return true;
    }
    nrays++;
    Hit<T> h;
    bool hit, looked_up = false;
    if constexpr (!EXACT && !is_general<SC>::value) {
        if (cand && ps.depth == cam.max_depth) {                // camera ray: the pixel's candidate list instead of the tree
            const uint4 c = __ldg(cand + pixel);
            if (c.x != kCandOverflow) {
                int32_t best; T best_t;
                hit = closest_prim_candidates<COUNT>(sc, ps.r, tmin, M<T, EXACT>::inf(), c, &best, &best_t, tl);
                if (hit) hit_record<T, EXACT, SC>(sc, ps.r, best, best_t, &h);
                looked_up = true;
            }
        }
    }
    if (!looked_up) hit = closest_hit<T, EXACT, COUNT, SC>(sc, ps.r, tmin, M<T, EXACT>::inf(), &h, stack, stride, tl);
    if (!hit) {   // camera.rs:473-475
        if (COUNT) tl.missed++;
        *value = ps.mult * cam.background + ps.res;
        return true;
    }
    V3<T> emitted = mk<T>(0, 0, 0);                             // Material::emitted default, material.rs:42-44
    if constexpr (is_general<SC>::value) emitted = g_emitted<T>(h);
    Stream<EXACT> rng(seed, pixel, sample, cam.max_depth - ps.depth + 1u, is_general<SC>::value && !EXACT);
    Ray<T> next;
    V3<T> w;
    uint32_t kind = shade<T, EXACT, COUNT, SC>(sc, ps.r, h, rng, &next, &w, tl);
    if (kind == V_ABSORB) { *value = ps.mult * emitted + ps.res; return true; }       // camera.rs:484-486
    if (kind == V_DIFFUSE) ps.res = ps.res + ps.mult * emitted;                       // camera.rs:519
    ps.mult = ps.mult * w;
    ps.r = next;
    ps.depth -= 1;
    return false;
}

template <class T> RTW_D V3<T> fix_nan(V3<T> v) { return mk<T>(v.x != v.x ? T(0) : v.x, v.y != v.y ? T(0) : v.y, v.z != v.z ? T(0) : v.z); }

// ---------------------------------------------------------------------------------------------
template <class T, class SCENE = SceneView<T>> struct RenderParams {
    SCENE scene;
    CameraT<T> cam;
    uint64_t seed;
    T tmin;
    uint32_t flags;
    uint32_t rank, world, tiles_x, tiles_total, n_local_tiles;
    T* tiles;                       // [n_local_tiles][kTileH][kTileW][3]
    unsigned int* work_counter;
    DeviceCounters* counters;
    // shared-memory staging of the scene (fast path only): bytes of each section, 0 = keep in global
    uint32_t smem_nodes, smem_spheres, smem_lights;
    uint32_t sh_node_stride;        // all-shared scenes: stride of the staged nodes (64 = Node<float> as is, 80 = padded copy from scene.nodes_staged)
    uint32_t stack_depth;           // traversal stack entries per thread: BVH depth + 2, at most kStackDepth
    const uint4* cand;              // [height * width] candidate lists of the camera rays (primary_candidates_kernel), or NULL
};

RTW_D uint32_t warp_sum(uint32_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <bool COUNT>
RTW_D void flush_counters(DeviceCounters* c, uint32_t npaths, uint32_t nrays, const Tally& tl) {
    uint32_t lane = threadIdx.x & 31;
    uint32_t p = warp_sum(npaths), r = warp_sum(nrays);
    if (lane == 0) { atomicAdd(&c->paths, (unsigned long long)p); atomicAdd(&c->rays, (unsigned long long)r); }
    if (COUNT) {
        uint32_t v[9] = {tl.node_visits, tl.sphere_tests, tl.light_tests, tl.lambertian, tl.metal, tl.dielectric, tl.absorbed, tl.missed, tl.depth_out};
        unsigned long long* dst = &c->node_visits;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            uint32_t s = warp_sum(v[i]);
            if (lane == 0) atomicAdd(dst + i, (unsigned long long)s);
        }
    }
}

// Scene staging: the sections the host chose (plan_smem) are copied global -> shared by the TMA engine
// (cp.async.bulk, SASS UBLKCP) while the CTA waits on one mbarrier; afterwards `sc` points into shared memory.
RTW_D void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* mbar) {
    uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst), m = (uint32_t)__cvta_generic_to_shared(mbar);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(d), "l"(gmem_src), "r"(bytes), "r"(m) : "memory");
}
RTW_D void stage_scene(const RenderParams<float>& P, unsigned char* cur, SceneView<float>& sc) {
    __shared__ __align__(8) uint64_t mbar;
    const uint32_t info_bytes = ((uint32_t)P.scene.n_spheres * 4u + 15u) / 16u * 16u;
    unsigned char* p_nodes = cur;        cur += P.smem_nodes;
    unsigned char* p_spheres = cur;      cur += P.smem_spheres;
    unsigned char* p_mat = cur;          cur += P.smem_spheres;
    unsigned char* p_info = cur;         cur += P.smem_spheres ? info_bytes : 0u;
    unsigned char* p_lights = cur;
    const uint32_t total = P.smem_nodes + (P.smem_spheres ? 2u * P.smem_spheres + info_bytes : 0u) + P.smem_lights;
    if (total) {
        const uint32_t m = (uint32_t)__cvta_generic_to_shared(&mbar);
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(m) : "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(m), "r"(total) : "memory");
            if (P.smem_nodes) tma_load_1d(p_nodes, P.sh_node_stride == kShNodeStridePadded ? P.scene.nodes_staged : (const void*)P.scene.nodes, P.smem_nodes, &mbar);
            if (P.smem_spheres) {
                tma_load_1d(p_spheres, P.scene.spheres, P.smem_spheres, &mbar);
                tma_load_1d(p_mat, P.scene.sphere_mat, P.smem_spheres, &mbar);
                tma_load_1d(p_info, P.scene.sphere_info, info_bytes, &mbar);      // the allocation is padded to 512 B
            }
            if (P.smem_lights) tma_load_1d(p_lights, P.scene.lights, P.smem_lights, &mbar);
        }
        uint32_t done = 0;
        for (int spin = 0; spin < (1 << 22) && !done; ++spin)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(m) : "memory");
        if (!done) __trap();                                   // never hang the GPU on a staging bug
    }
    if (P.smem_nodes) { sc.top_nodes = reinterpret_cast<const Node<float>*>(p_nodes); sc.n_top = (int32_t)(P.smem_nodes / (P.sh_node_stride ? P.sh_node_stride : 64u)); }
    if (P.smem_spheres) {
        sc.spheres = reinterpret_cast<const Vec4T<float>*>(p_spheres);
        sc.sphere_mat = reinterpret_cast<const Vec4T<float>*>(p_mat);
        sc.sphere_info = reinterpret_cast<const uint32_t*>(p_info);
    }
    if (P.smem_lights) sc.lights = reinterpret_cast<const Vec4T<float>*>(p_lights);
}
template <class T> RTW_D void stage_scene(const RenderParams<T>&, unsigned char*, SceneView<T>&) {}   // exact path: global memory

// Megakernel: persistent CTAs; each warp pulls 8x4-pixel warp tiles from this GPU's work queue
// (an atomic counter); each lane owns one pixel and runs its spp paths back to back, regenerating a
// camera ray as soon as its previous path ends (render_internal + ray_colour_tail_call,
// camera.rs:315-388, 460-522).  Samples of a pixel are summed in sample order.
template <class T, bool EXACT, bool COUNT, int BLOCK, bool SH = false, class SCENE = SceneView<T>>
__global__ void __launch_bounds__(BLOCK) render_mega_kernel(RenderParams<T, SCENE> P) {
    static_assert(!(EXACT && SH), "the exact path reads the scene from global memory");
    static_assert(!(is_general<SCENE>::value && SH), "general scenes are read from global memory");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int32_t* stack_base = reinterpret_cast<int32_t*>(smem_raw);                  // [kStackDepth][BLOCK]
    using SC = typename std::conditional<SH, SceneViewSh<T>, SCENE>::type;
    SC scv;
    if constexpr (is_general<SCENE>::value) scv = P.scene;
    else {
        SceneView<T> sc = P.scene;
        if constexpr (!EXACT) stage_scene(P, smem_raw + sizeof(int32_t) * P.stack_depth * BLOCK, sc);
        static_cast<SceneView<T>&>(scv) = sc;
        bind_scene(scv, P.sh_node_stride);
    }
    const CameraT<T>& cam = P.cam;
    const uint32_t lane = threadIdx.x & 31;
    int32_t* stack = stack_base + threadIdx.x;
    const uint32_t n_warp_tiles = P.n_local_tiles * kWarpTilesPerTile;
    uint32_t npaths = 0, nrays = 0;
    Tally tl;
    for (;;) {
        uint32_t wt = 0;
        if (lane == 0) wt = atomicAdd(P.work_counter, 1u);
        wt = __shfl_sync(0xffffffffu, wt, 0);
        if (wt >= n_warp_tiles) break;
        uint32_t local_tile = wt / kWarpTilesPerTile, sub = wt % kWarpTilesPerTile;
        uint32_t tile = local_tile * P.world + P.rank;
        uint32_t lx = (sub % (kTileW / kWarpTileW)) * kWarpTileW + (lane % kWarpTileW);
        uint32_t ly = (sub / (kTileW / kWarpTileW)) * kWarpTileH + (lane / kWarpTileW);
        uint32_t ttx, tty;
        slot_tile(tile, P.tiles_x, &ttx, &tty);
        uint32_t i = ttx * kTileW + lx, j = tty * kTileH + ly;
        bool valid = tile < P.tiles_total && i < cam.width && j < cam.height;
        uint32_t pixel = j * cam.width + i;
        V3<T> acc = mk<T>(0, 0, 0);
        uint32_t sample = 0;
        bool alive = false;
        PathState<T> ps;
        ps.depth = 0;
        for (;;) {
            if (!alive && valid && sample < cam.spp) {
                Stream<EXACT> rng(P.seed, pixel, sample + cam.sample_offset, 0u, is_general<SC>::value && !EXACT);
                ps.r = get_ray<T, EXACT>(cam, i, j, rng);
                ps.mult = mk<T>(1, 1, 1); ps.res = mk<T>(0, 0, 0); ps.depth = cam.max_depth;
                alive = true;
                npaths++;
            }
            if (!__any_sync(0xffffffffu, alive)) break;
            if (alive) {
                V3<T> value;
                if (path_step<T, EXACT, COUNT, SC>(scv, cam, P.seed, P.tmin, pixel, sample + cam.sample_offset, ps, &value, stack, BLOCK, nrays, tl)) {
                    if (P.flags & 1u) value = fix_nan(value);
                    acc = acc + value;                          // fold(Colour::default(), +), camera.rs:335
                    alive = false;
                    sample++;
                }
            }
        }
        T* out = P.tiles + ((size_t)local_tile * (kTileW * kTileH) + (size_t)ly * kTileW + lx) * 3;
        out[0] = acc.x; out[1] = acc.y; out[2] = acc.z;         // zeros for padding pixels
    }
    flush_counters<COUNT>(P.counters, npaths, nrays, tl);
}

// ---------------------------------------------------------------------------------------------
// Pooled megakernel (fast path).  The lane-per-pixel kernel above leaves two thirds of the lanes idle
// (ncu: 10.6 of 32 threads active — a lane whose pixel shows only background finishes its spp paths
// long before a neighbour looking at glass).  Here lanes are decoupled from pixels: the image's paths
// form one pixel-major stream (all samples of a pixel are consecutive), warps pull chunks of it from
// a global queue and every idle lane takes the next path of its warp's chunk.  So the 32 lanes of a
// warp mostly trace samples of the SAME pixel (coherent primary rays, same first material) and no
// lane idles until the stream is dry.
// Radiance is accumulated in 64-bit fixed point (2^-32 units): integer adds commute, so the image is
// bit-identical no matter which lane, warp or GPU traces which path.  A lane keeps a private partial
// sum while it stays on one pixel and flushes it with one 64-bit reduction per channel when it moves on.
// NaN / overflow samples set per-pixel poison bits (a NaN sample poisons the pixel like in the
// reference, where the f64 sum becomes NaN and `as u8` maps it to 0).
constexpr float kFixedScale = 4294967296.f;            // 2^32
constexpr float kFixedMax = 1073741824.f;              // 2^30: upper bound of PoolParams::sample_cap

struct PoolParams {
    unsigned long long* accum;      // [n_local_tiles * 256][3]
    uint32_t* poison;               // [n_local_tiles * 256]: poison words (poison_nan / poison_inf below)
    uint32_t pixels_per_chunk;      // G: pixel slots per work chunk (1 when spp is large)
    uint32_t n_chunks;
    float sample_cap;               // samples of this radiance or more set the overflow flag instead of being added (pool_sample_cap)
    const uint32_t* chunk_order;    // NULL or the ordered queue (order_words(): layout at chunk_order_kernel): its k-th entry is a chunk (kChunkMask), flags
                                    // (kChunkCheap / kChunkEnd) and, for the last costly chunks, which eighth of the chunk's paths it stands for (kChunkSubShift)
    const uint32_t* queue_len;      // device word or NULL: the queue has been split (chunk_split_kernel) — entries flagged kChunkEnd belong to render_background_kernel
    uint32_t queue_cap;             // positions of the queue: n_chunks, or n_chunks + order_extra() when chunk_order is set
};
constexpr uint32_t kChunkCheap = 0x80000000u, kChunkEnd = 0x40000000u;
// The LAST costly chunks of the queue are handed out in kChunkSubs pieces (entry bits 26..29: 0 = the whole chunk, k = piece k - 1): at 500
// samples per pixel a chunk is one pixel's 500 paths, ~2 ms of a warp's time, and the warps would reach the end of the costly work up to one
// chunk apart.  (Cutting EVERY chunk costs 2-3 % — profiles/r2_sample_blocks_ab.jsonl; only the end of the queue needs the fine grain.)
constexpr uint32_t kChunkSubShift = 26u, kChunkSubs = 8u, kChunkMask = (1u << kChunkSubShift) - 1u;
// order buffer, in words: [0, cap) the queue | [cap, cap + n) background-only chunks | [cap + n, cap + 2n) costly chunks as chunk_order_kernel
// found them | 8 words, 5 used: their two counts (the kernel's cursors, zeroed by the host), the queue's length, the chunks of render_background_kernel, its cursor
RTW_HD uint32_t order_extra(uint32_t split_chunks) { return split_chunks * (kChunkSubs - 1u); }
RTW_HD size_t order_words(uint32_t n_chunks, uint32_t cap) { return (size_t)cap + 2u * (size_t)n_chunks + 8u; }

// poison word: six flags (NaN r/g/b, overflow r/g/b), each the low bit of its own 4-bit field, so that the words of up to 15
// ranks can be SUMMED by a reduce without one flag carrying into the next (a flag is set iff its field is non-zero)
RTW_HD uint32_t poison_nan(uint32_t channel) { return 1u << (4u * channel); }
RTW_HD uint32_t poison_inf(uint32_t channel) { return 1u << (12u + 4u * channel); }
RTW_HD bool poison_has_nan(uint32_t word, uint32_t channel) { return (word >> (4u * channel)) & 15u; }
RTW_HD bool poison_has_inf(uint32_t word, uint32_t channel) { return (word >> (12u + 4u * channel)) & 15u; }
// No silent wrap of the 64-bit accumulators, by construction instead of by checking every reduction: a pixel receives at most
// `spp` samples per frame over ALL ranks, and a sample of `sample_cap` = 2^28 / spp radiance units or more (PoolParams, at most
// 2^30) sets the channel's overflow flag instead of being added — so every sum, per lane, per pixel, per rank and after the ranks'
// accumulators have been added by an integer reduce, stays below 2^60.  A flagged pixel resolves to +inf -> 255, which is what the
// f64 sum of such samples resolves to as well unless spp is so large (> 2^18) that the cap drops below ~1000; the reductions stay
// fire-and-forget (RED, no returned value to wait for).
RTW_D void pool_flush(const PoolParams& Q, uint32_t q, unsigned long long a0, unsigned long long a1, unsigned long long a2) {
    if (a0) atomicAdd(Q.accum + 3 * (size_t)q + 0, a0);
    if (a1) atomicAdd(Q.accum + 3 * (size_t)q + 1, a1);
    if (a2) atomicAdd(Q.accum + 3 * (size_t)q + 2, a2);
}
RTW_D unsigned long long pool_fixed(const PoolParams& Q, float v, uint32_t channel, uint32_t& bad) {
    if (!(v < Q.sample_cap)) {                                 // NaN, +inf or too large to be summed safely
        bad |= (v != v) ? poison_nan(channel) : poison_inf(channel);
        return 0ull;
    }
    return __float2ull_rn(v * kFixedScale);                    // negative values clamp to 0
}

template <bool COUNT, int BLOCK, bool SH, class SCENE = SceneView<float>>
__global__ void __launch_bounds__(BLOCK, is_general<SCENE>::value ? 3 : 4) render_pool_kernel(RenderParams<float, SCENE> P, PoolParams Q) {
    using T = float;
    constexpr bool EXACT = false;
    static_assert(!(is_general<SCENE>::value && SH), "general scenes are read from global memory");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int32_t* stack_base = reinterpret_cast<int32_t*>(smem_raw);                  // [kStackDepth][BLOCK]
    using SC = typename std::conditional<SH, SceneViewSh<T>, SCENE>::type;
    SC scv;
    if constexpr (is_general<SCENE>::value) scv = P.scene;
    else {
        SceneView<T> sc = P.scene;
        stage_scene(P, smem_raw + sizeof(int32_t) * P.stack_depth * BLOCK, sc);
        static_cast<SceneView<T>&>(scv) = sc;
        bind_scene(scv, P.sh_node_stride);
    }
    const CameraT<T>& cam = P.cam;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;
    int32_t* stack = stack_base + threadIdx.x;
    const uint32_t n_slots = P.n_local_tiles * (kTileW * kTileH);
    const uint32_t spp = cam.spp, G = Q.pixels_per_chunk;
    uint32_t npaths = 0, nrays = 0;
    Tally tl;
    // warp-uniform chunk cursor
    uint32_t chunk_next = 0, chunk_end = 0, chunk_q0 = 0;
    bool exhausted = (spp == 0);
    // lane state
    bool alive = false;
    uint32_t q = 0xffffffffu, pixel = 0, sample = 0;
    uint32_t acc_q = 0xffffffffu, bad = 0;
    unsigned long long a0 = 0, a1 = 0, a2 = 0;
    PathState<T> ps;
    ps.depth = 0;
    for (;;) {
        uint32_t need = __ballot_sync(0xffffffffu, !alive);
        if (need) {
            if (chunk_next == chunk_end && !exhausted) {
                uint32_t c = 0;
                if (lane == 0) c = atomicAdd(P.work_counter, 1u);
                c = __shfl_sync(0xffffffffu, c, 0);
                if (c >= Q.n_chunks) {
                    exhausted = true;
                } else {
                    chunk_q0 = c * G;
                    uint32_t npx = min(G, n_slots - chunk_q0);
                    chunk_next = 0;
                    chunk_end = npx * spp;
                    if (G == 1) {       // skip padding pixels (outside the image / padding tiles) as a whole
                        uint32_t tile = (chunk_q0 >> 8) * P.world + P.rank, in = chunk_q0 & 255u;
                        uint32_t ttx, tty;
                        slot_tile(tile, P.tiles_x, &ttx, &tty);
                        uint32_t i = ttx * kTileW + (in & 15u), j = tty * kTileH + (in >> 4);
                        if (!(tile < P.tiles_total && i < cam.width && j < cam.height)) chunk_end = 0;
                    }
                }
            }
            if (!alive && chunk_next < chunk_end) {
                uint32_t r = chunk_next + __popc(need & lt_mask);
                if (r < chunk_end) {
                    uint32_t pin = r / spp;
                    sample = r - pin * spp + cam.sample_offset;
                    q = chunk_q0 + pin;
                    uint32_t tile = (q >> 8) * P.world + P.rank, in = q & 255u;
                    uint32_t ttx, tty;
                        slot_tile(tile, P.tiles_x, &ttx, &tty);
                        uint32_t i = ttx * kTileW + (in & 15u), j = tty * kTileH + (in >> 4);
                    if (tile < P.tiles_total && i < cam.width && j < cam.height) {
                        pixel = j * cam.width + i;
                        Stream<EXACT> rng(P.seed, pixel, sample, 0u, is_general<SC>::value && !EXACT);
                        ps.r = get_ray<T, EXACT>(cam, i, j, rng);
                        ps.mult = mk<T>(1, 1, 1); ps.res = mk<T>(0, 0, 0); ps.depth = cam.max_depth;
                        alive = true;
                        npaths++;
                    }
                }
            }
            chunk_next = min(chunk_end, chunk_next + (uint32_t)__popc(need));
        }
        if (!__any_sync(0xffffffffu, alive)) {
            if (exhausted) break;
            continue;
        }
        if (alive) {
            V3<T> value;
            if (path_step<T, EXACT, COUNT, SC>(scv, cam, P.seed, P.tmin, pixel, sample, ps, &value, stack, BLOCK, nrays, tl, P.cand)) {
                alive = false;
                if (P.flags & 1u) value = fix_nan(value);
                if (q != acc_q) {
                    if (acc_q != 0xffffffffu) {
                        pool_flush(Q, acc_q, a0, a1, a2);
                        if (bad) atomicOr(Q.poison + acc_q, bad);
                    }
                    acc_q = q; a0 = a1 = a2 = 0ull; bad = 0;
                }
                a0 += pool_fixed(Q, value.x, 0, bad);
                a1 += pool_fixed(Q, value.y, 1, bad);
                a2 += pool_fixed(Q, value.z, 2, bad);
            }
        }
    }
    if (acc_q != 0xffffffffu) {
        pool_flush(Q, acc_q, a0, a1, a2);
        if (bad) atomicOr(Q.poison + acc_q, bad);
    }
    flush_counters<COUNT>(P.counters, npaths, nrays, tl);
}

// fixed-point accumulators -> the float tile buffer of the ABI
template <int UNUSED = 0>
__global__ void pool_finalize_kernel(const unsigned long long* accum, const uint32_t* poison, float* tiles, uint32_t n_slots) {
    uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_slots) return;
    uint32_t bad = poison[q];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float v = (float)((double)accum[3 * (size_t)q + c] * (1.0 / 4294967296.0));
        if (poison_has_inf(bad, c)) v = __int_as_float(0x7f800000);
        if (poison_has_nan(bad, c)) v = __int_as_float(0x7fc00000);
        tiles[3 * (size_t)q + c] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// Candidate lists of the camera rays (see closest_prim_candidates).  One thread per pixel walks the FP32 tree with the pixel's
// cone: apex = the camera centre, axis = the direction through the pixel centre, half-angle alpha with tan(alpha) >= half the
// pixel's diagonal over the distance to it (the jitter is U[-0.5, 0.5]^2 pixels).  A sphere (centre c, radius R) — a primitive, or
// the bounding sphere of a node's box — can touch the cone only if it contains the apex, or lies ahead of it (s = axis.(c - o) >=
// -R) with its centre no farther from the axis than s tan(alpha) + R / cos(alpha).  The distance from the axis is |axis x (c - o)|:
// every term is of the size of R, nothing cancels, so the test is sound in FP32 at pixel cones of 1e-3 rad; R is inflated by 0.1 %
// plus 2e-5 of the distance to cover the rounding of the test itself.  Conservative tests only add candidates.
RTW_D bool cone_touches_sphere(V3<float> axis, float tan_a, float inv_cos_a, V3<float> v, float R) {
    const float d2 = dot(v, v);
    const float Ri = fmaf(R, 1.001f, 2e-5f * fsqrt(d2));
    if (d2 <= Ri * Ri) return true;
    const float s = dot(axis, v);
    if (s < -Ri) return false;
    const V3<float> w = cross(axis, v);
    const float lim = fmaf(fmaxf(s, 0.f), tan_a, Ri * inv_cos_a);
    return dot(w, w) <= lim * lim;
}
// Walks the tree with a cone and collects the sorted indices of the spheres it touches: up to CAP of them in ids[], returns their
// number (CAP + 1 = more than CAP).
template <int CAP>
RTW_D int cone_walk(const SceneView<float>& sc, V3<float> apex, V3<float> axis, float tan_a, float inv_cos_a, uint32_t* ids) {
    int n = 0;
    int32_t stack[kStackDepth];
    int sp = 0;
    int32_t cur = sc.n_spheres > 0 ? 0 : kStop;
    while (cur != kStop && n <= CAP) {
        if (cur >= 0) {
            Node<float> nd;
            const float4* p = reinterpret_cast<const float4*>(sc.nodes + cur);
            unpack_node(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3), nd);
            const V3<float> vl = mk<float>(nd.la[0], nd.la[1], nd.la[2]) - apex, vr = mk<float>(nd.ra[0], nd.ra[1], nd.ra[2]) - apex;
            const float Rl = fsqrt(nd.lb[0] * nd.lb[0] + nd.lb[1] * nd.lb[1] + nd.lb[2] * nd.lb[2]);
            const float Rr = fsqrt(nd.rb[0] * nd.rb[0] + nd.rb[1] * nd.rb[1] + nd.rb[2] * nd.rb[2]);
            const bool hl = nd.left != kEmptyLeaf && cone_touches_sphere(axis, tan_a, inv_cos_a, vl, Rl);
            const bool hr = nd.right != kEmptyLeaf && cone_touches_sphere(axis, tan_a, inv_cos_a, vr, Rr);
            if (hl && hr) { stack[sp++] = nd.right; cur = nd.left; }
            else if (hl) cur = nd.left;
            else if (hr) cur = nd.right;
            else cur = sp ? stack[--sp] : kStop;
        } else {
            const uint32_t enc = (uint32_t)~cur;
            const uint32_t first = enc >> 4, count = (enc & 15u) + 1u;
            for (uint32_t k = first; k < first + count; ++k) {
                const float4 s = __ldg(reinterpret_cast<const float4*>(sc.spheres + k));
                if (cone_touches_sphere(axis, tan_a, inv_cos_a, mk<float>(s.x, s.y, s.z) - apex, s.w)) {
                    if (n < CAP) ids[n] = k;
                    n++;
                }
            }
            cur = sp ? stack[--sp] : kStop;
        }
    }
    return n;
}
// Two levels, two launches.  (1) One THREAD per 8 x 8-pixel block walks the tree with the block's cone and writes up to 24 candidates to a
// scratch list.  (2) One thread per pixel filters its block's short list with its own cone; only a block with more than 24 candidates sends
// its pixels down the tree one by one.  (First version: one walk per pixel, 0.24 ms per 1080p frame.  Second: one CTA per block whose thread 0
// walked while 63 threads waited — at most 32 walks in flight per SM, 0.20 ms under ncu.  A walk is a chain of ~50 dependent node loads, so
// what it needs is many of them in flight: here 2 048 per SM.)
constexpr int kCandBlock = 8, kCandBlockCap = 24;
struct CandBlockList { uint32_t n; uint32_t ids[kCandBlockCap]; };          // n = kCandBlockCap + 1: more than fit
RTW_D float cand_half_pixel(const CameraT<float>& cam) { return 0.5f * (fsqrt(dot(cam.du, cam.du)) + fsqrt(dot(cam.dv, cam.dv))) * 1.001f; }   // >= half a pixel's diagonal
template <int UNUSED = 0>
__global__ void __launch_bounds__(128) block_candidates_kernel(SceneView<float> sc, CameraT<float> cam, uint32_t blocks_x, uint32_t n_blocks, CandBlockList* lists) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    const uint32_t bx = (b % blocks_x) * kCandBlock, by = (b / blocks_x) * kCandBlock;
    const float px = cand_half_pixel(cam);
    const V3<float> dc = (cam.pixel00 + cam.du * ((float)bx + 3.5f)) + cam.dv * ((float)by + 3.5f) - cam.center;
    const float len = fsqrt(dot(dc, dc)), hd = 8.f * px;           // pixel centres up to 3.5 away per axis + half a pixel of jitter
    uint32_t ids[kCandBlockCap];
    int n = kCandBlockCap + 1;
    if (len > 4.f * hd) {
        const float tan_a = hd * frcp(len - hd) * 1.001f;
        n = cone_walk<kCandBlockCap>(sc, cam.center, dc * frcp(len), tan_a, fsqrt(fmaf(tan_a, tan_a, 1.f)) * 1.0001f, ids);
    }
    CandBlockList& out = lists[b];
    out.n = (uint32_t)n;
    for (int k = 0; k < kCandBlockCap; ++k) if (k < n) out.ids[k] = ids[k];
}
template <int UNUSED = 0>
__global__ void __launch_bounds__(kCandBlock * kCandBlock) primary_candidates_kernel(SceneView<float> sc, CameraT<float> cam, const CandBlockList* lists, uint4* cand) {
    __shared__ CandBlockList bl;
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(lists + ((size_t)blockIdx.y * gridDim.x + blockIdx.x));
        if (threadIdx.x < sizeof(CandBlockList) / 4) reinterpret_cast<uint32_t*>(&bl)[threadIdx.x] = __ldg(src + threadIdx.x);
    }
    __syncthreads();
    const int block_n = (int)bl.n;
    const uint32_t* block_ids = bl.ids;
    const uint32_t bx = blockIdx.x * kCandBlock, by = blockIdx.y * kCandBlock;
    const float px = cand_half_pixel(cam);
    const uint32_t i = bx + threadIdx.x % kCandBlock, j = by + threadIdx.x / kCandBlock;
    if (i >= cam.width || j >= cam.height) return;
    const V3<float> dc = (cam.pixel00 + cam.du * (float)i) + cam.dv * (float)j - cam.center;
    const float len = fsqrt(dot(dc, dc));
    uint4 out = make_uint4(kCandNone, kCandNone, kCandNone, kCandNone);
    if (!(len > 4.f * px)) out.x = kCandOverflow;                       // degenerate camera: no cone, walk the tree
    else {
        const V3<float> axis = dc * frcp(len);
        const float tan_a = px * frcp(len - px) * 1.001f;
        const float inv_cos_a = fsqrt(fmaf(tan_a, tan_a, 1.f)) * 1.0001f;
        uint32_t ids[4] = {kCandNone, kCandNone, kCandNone, kCandNone};
        int n = 0;
        if (block_n <= kCandBlockCap) {
            for (int k = 0; k < block_n; ++k) {
                const float4 s = __ldg(reinterpret_cast<const float4*>(sc.spheres + block_ids[k]));
                if (cone_touches_sphere(axis, tan_a, inv_cos_a, mk<float>(s.x, s.y, s.z) - cam.center, s.w)) {
                    if (n < 4) ids[n] = block_ids[k];
                    n++;
                }
            }
        } else n = cone_walk<4>(sc, cam.center, axis, tan_a, inv_cos_a, ids);
        if (n > 4) out.x = kCandOverflow;
        else out = make_uint4(ids[0], ids[1], ids[2], ids[3]);
    }
    cand[j * cam.width + i] = out;
}

// Order of the work queue: longest processing time first.  A path that only sees background ends in GENERATE; a path that meets a
// sphere bounces on for up to max_depth vertices (zero-weight paths are followed to the end like in the reference), and with
// warp-private queues the frame is over when the LAST such path is — ~0.7 ms after the stream has run dry on `simple`, whatever the
// sample count (depth 1: none of it), i.e. 3 % of the 24 ms one of eight GPUs spends on its share.  The candidate lists already say
// which pixels can meet a sphere; a one-sided plane is met iff one of the pixel's corner rays meets it (the set of such directions is
// a half-space).  Chunks with such a pixel are dealt out first, the background-only chunks last — flagged kChunkCheap, on which a warp
// switches to "old paths first" — so the long paths end while the cheap tail of the stream still keeps every warp busy.
// Scheduling only: paths, RNG streams and the fixed-point sums do not depend on who traces what when.
// This kernel sorts the chunks into the two regions behind the queue (layout: order_words() above; the two cursors are zeroed by the host);
// chunk_split_kernel then writes the queue itself.
// own_world > 1 (several GPUs, one frame): this GPU only takes the chunks it OWNS — chunk c belongs to GPU (c + (c >> 4)) % own_world, a
// diagonal 1-chunk interleave — and renders ALL samples of their pixels.  Every GPU's accumulators then cover a disjoint set of pixels and
// the frame's one integer reduce puts them together; unlike a split of every pixel's samples, the per-pixel costs of a frame (candidate
// lookups, partial-sum flushes, queue transactions) are paid once, not once per GPU.
template <int UNUSED = 0>
__global__ void chunk_order_kernel(const uint4* cand, SceneView<float> sc, CameraT<float> cam, uint32_t rank, uint32_t world, uint32_t tiles_x,
                                   uint32_t tiles_total, uint32_t n_slots, uint32_t G, uint32_t n_chunks, uint32_t cap, uint32_t own_rank, uint32_t own_world,
                                   uint32_t* order) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chunks) return;
    if (own_world > 1u && (c + (c >> 4)) % own_world != own_rank) return;
    bool costly = false;
    const uint32_t q1 = min(n_slots, (c + 1u) * G);
    for (uint32_t q = c * G; q < q1 && !costly; ++q) {
        const uint32_t tile = (q >> 8) * world + rank, in = q & 255u;
        if (tile >= tiles_total) continue;
        uint32_t ttx, tty;
        slot_tile(tile, tiles_x, &ttx, &tty);
        const uint32_t i = ttx * kTileW + (in & 15u), j = tty * kTileH + (in >> 4);
        if (i >= cam.width || j >= cam.height) continue;
        if (__ldg(cand + (size_t)j * cam.width + i).x != kCandNone) { costly = true; break; }
        for (int k = 0; k < sc.n_planes && !costly; ++k) {
            const PlaneT<float>& pl = sc.planes[k];
            const float side = -dot(cam.center - pl.point, pl.normal);
#pragma unroll
            for (int corner = 0; corner < 4; ++corner) {
                const float fi = (float)i + ((corner & 1) ? 0.5f : -0.5f), fj = (float)j + ((corner & 2) ? 0.5f : -0.5f);
                const V3<float> d = (cam.pixel00 + cam.du * fi) + cam.dv * fj - cam.center;
                const float denom = dot(d, pl.normal);
                if (denom > 0.f && side >= 0.f) costly = true;
            }
        }
    }
    // ballot-aggregated cursors: one atomic per warp and class
    const uint32_t lane = threadIdx.x & 31u, active = __activemask();
    const uint32_t mc = __ballot_sync(active, costly), mh = active & ~mc;
    uint32_t base_c = 0, base_h = 0;
    const uint32_t leader = (uint32_t)__ffs(active) - 1u;
    uint32_t* cheap = order + cap;
    uint32_t* found = cheap + n_chunks;
    uint32_t* meta = found + n_chunks;
    if (lane == leader) {
        if (mc) base_c = atomicAdd(meta, (uint32_t)__popc(mc));
        if (mh) base_h = atomicAdd(meta + 1u, (uint32_t)__popc(mh));
    }
    base_c = __shfl_sync(active, base_c, (int)leader); base_h = __shfl_sync(active, base_h, (int)leader);
    const uint32_t lt = (1u << lane) - 1u;
    if (costly) found[base_c + (uint32_t)__popc(mc & lt)] = c;
    else cheap[base_h + (uint32_t)__popc(mh & lt)] = c | kChunkCheap;
}

// The background-only chunks need none of the wavefront's machinery — no path slots, no stage lists: their camera rays have an empty
// candidate list and no plane in reach, so each path is one ray, one plane test and the background colour.  chunk_split_kernel
// leaves the LAST `tail_chunks` of them in the wavefront's queue (that tail is what lets every warp finish its old paths while it still
// has new ones to generate, see chunk_order_kernel) and hands the rest to render_background_kernel: one warp per chunk, lanes
// striding over a pixel's samples — the same Philox stream, the same get_ray, the same plane test as GENERATE — and ONE set of
// reductions per pixel (n x the fixed-point background sample; integer arithmetic, so the sum is what n separate additions give).
// Should a plane be met after all (the classification has margins; this is the belt to its braces) the path is traced to its end
// right here with the megakernel's path_step.
// meta[0 / 1]: number of costly / background-only chunks (the cursors of chunk_order_kernel); [2]: queue length; [3]: chunks of
// render_background_kernel (the first that many entries of the background-only region).
// The queue, front to back: the costly chunks whole, the last `split_chunks` of them in kChunkSubs pieces each, `tail_chunks` background-only
// chunks, end marks.
template <int UNUSED = 0>
__global__ void chunk_split_kernel(uint32_t* order, uint32_t n_chunks, uint32_t cap, uint32_t tail_chunks, uint32_t split_chunks) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;             // one thread per queue position
    const uint32_t* cheap = order + cap;
    const uint32_t* found = cheap + n_chunks;
    uint32_t* meta = order + cap + 2u * (size_t)n_chunks;
    const uint32_t n_costly = meta[0], n_cheap = meta[1];
    const uint32_t n_split = min(n_costly, split_chunks), whole = n_costly - n_split, pieces = n_split * kChunkSubs;
    const uint32_t keep = min(n_cheap, tail_chunks), queue_len = whole + pieces + keep;
    if (p == 0) { meta[2] = queue_len; meta[3] = n_cheap - keep; meta[4] = 0u; }
    if (p >= cap) return;
    // every position behind the queue carries the end mark (the queue counter only grows: a warp that fetches one knows the queue has ended)
    uint32_t e = kChunkEnd;
    if (p < whole) e = found[p];
    else if (p < whole + pieces) e = found[whole + (p - whole) / kChunkSubs] | (((p - whole) % kChunkSubs + 1u) << kChunkSubShift);
    else if (p < queue_len) e = cheap[(n_cheap - keep) + (p - whole - pieces)];
    order[p] = e;
}

template <class SC>
__device__ __noinline__ void background_slow_path(const SC& sc, const RenderParams<float>& P, const PoolParams& Q, uint32_t q, uint32_t pixel, uint32_t i,
                                                  uint32_t j, uint32_t sample, int32_t* stack, int stride, uint32_t& nrays, Tally& tl) {
    Stream<false> rng(P.seed, pixel, sample, 0u);
    PathState<float> ps;
    ps.r = get_ray<float, false>(P.cam, i, j, rng);
    ps.mult = mk<float>(1, 1, 1); ps.res = mk<float>(0, 0, 0); ps.depth = P.cam.max_depth;
    V3<float> value;
    while (!path_step<float, false, false, SC>(sc, P.cam, P.seed, P.tmin, pixel, sample, ps, &value, stack, stride, nrays, tl)) {}
    if (P.flags & 1u) value = fix_nan(value);
    uint32_t bad = 0;
    pool_flush(Q, q, pool_fixed(Q, value.x, 0, bad), pool_fixed(Q, value.y, 1, bad), pool_fixed(Q, value.z, 2, bad));
    if (bad) atomicOr(Q.poison + q, bad);
}

// Launch shape: CTAs of 128 threads, 64 registers, no shared memory, chunks pulled from a cursor (+0.4 % on C2 over 256-thread CTAs striding over
// the chunks).  Small enough to sit NEXT TO the wavefront's one CTA per SM — which was tried (RTW_SIDE_STREAM=1, capi.cu) and costs the wavefront
// kernel more than the overlap returns.
constexpr int kBackgroundBlock = 128;
template <int UNUSED = 0>
__global__ void __launch_bounds__(kBackgroundBlock, 8) render_background_kernel(RenderParams<float> P, PoolParams Q, uint32_t* order, uint32_t n_chunks) {
    const uint32_t* cheap = order + Q.queue_cap;
    uint32_t* meta = order + Q.queue_cap + 2u * (size_t)n_chunks;
    const SceneView<float>& sc = P.scene;
    const CameraT<float>& cam = P.cam;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n_bg = __ldg(meta + 3);
    const uint32_t n_slots = P.n_local_tiles * (kTileW * kTileH), spp = cam.spp, G = Q.pixels_per_chunk;
    uint32_t npaths = 0, nrays = 0;
    Tally tl;
    V3<float> value = mk<float>(1.f, 1.f, 1.f) * cam.background + mk<float>(0.f, 0.f, 0.f);      // mult * background + res (camera.rs:473-475)
    if (P.flags & 1u) value = fix_nan(value);
    for (;;) {
        uint32_t k = 0;
        if (lane == 0) k = atomicAdd(meta + 4, 1u);
        k = __shfl_sync(0xffffffffu, k, 0);
        if (k >= n_bg) break;
        const uint32_t c = __ldg(cheap + k) & kChunkMask;
        const uint32_t q0 = c * G, npx = min(G, n_slots - q0);
        for (uint32_t pin = 0; pin < npx; ++pin) {
            const uint32_t q = q0 + pin;
            const uint32_t tile = (q >> 8) * P.world + P.rank, in = q & 255u;
            uint32_t ttx, tty;
            slot_tile(tile, P.tiles_x, &ttx, &tty);
            const uint32_t i = ttx * kTileW + (in & 15u), j = tty * kTileH + (in >> 4);
            if (!(tile < P.tiles_total && i < cam.width && j < cam.height)) continue;
            const uint32_t pixel = j * cam.width + i;
            uint32_t miss = 0;
            for (uint32_t s0 = lane; s0 < spp; s0 += 32u) {
                const uint32_t sample = s0 + cam.sample_offset;
                Stream<false> rng(P.seed, pixel, sample, 0u);
                const Ray<float> r = get_ray<float, false>(cam, i, j, rng);
                bool found = false;
                float best_t = M<float, false>::inf();
                int32_t best = -1;
                closest_plane<float, false, SceneView<float>>(sc, r, P.tmin, M<float, false>::inf(), found, best_t, best);
                npaths++;
                if (!found) { miss++; nrays++; }
                else {                                        // only this path walks a tree: its stack lives in local memory
                    int32_t stack_l[kStackDepth];
                    background_slow_path(sc, P, Q, q, pixel, i, j, sample, stack_l, 1, nrays, tl);
                }
            }
            miss = warp_sum(miss);
            if (lane == 0 && miss) {
                uint32_t bad = 0;
                const unsigned long long n = miss;
                pool_flush(Q, q, n * pool_fixed(Q, value.x, 0, bad), n * pool_fixed(Q, value.y, 1, bad), n * pool_fixed(Q, value.z, 2, bad));
                if (bad) atomicOr(Q.poison + q, bad);
                if (P.flags & 2u) tl.missed += miss;          // RTW_FLAG_COUNT_EVENTS
            }
        }
    }
    flush_counters<true>(P.counters, npaths, nrays, tl);
}

// ---------------------------------------------------------------------------------------------
// Batch kernels (one thread per ray) — the parity surface.
template <class T, class SCENE = SceneView<T>> struct BatchParams {
    SCENE scene;
    CameraT<T> cam;
    uint64_t seed;
    T tmin, tmax;
    uint32_t flags;
    size_t n;
    const double *o, *d;                     // [n][3]
    const uint32_t *a, *b, *c;               // (pixel, sample, vertex) or (i, j, sample)
    int32_t* prim; double* t; uint32_t* kind;
    double *p, *normal, *dir, *weight, *rgb;
};

template <class T> RTW_D V3<T> load3(const double* p, size_t i) { return mk<T>((T)p[3 * i], (T)p[3 * i + 1], (T)p[3 * i + 2]); }
template <class T> RTW_D void store3(double* p, size_t i, V3<T> v) { p[3 * i] = (double)v.x; p[3 * i + 1] = (double)v.y; p[3 * i + 2] = (double)v.z; }

template <class T, bool EXACT, int BLOCK, class SCENE = SceneView<T>>
__global__ void __launch_bounds__(BLOCK) trace_batch_kernel(BatchParams<T, SCENE> P) {
    __shared__ int32_t stack_s[kStackDepth * BLOCK];
    size_t idx = (size_t)blockIdx.x * BLOCK + threadIdx.x;
    if (idx >= P.n) return;
    Ray<T> r{load3<T>(P.o, idx), load3<T>(P.d, idx)};
    Hit<T> h;
    Tally tl;
    bool hit = closest_hit<T, EXACT, false, SCENE>(P.scene, r, P.tmin, P.tmax, &h, stack_s + threadIdx.x, BLOCK, tl);
    P.prim[idx] = hit ? (int32_t)(h.info >> 2) : -1;
    P.t[idx] = hit ? (double)h.t : __builtin_huge_val();
}

template <class T, bool EXACT, int BLOCK, class SCENE = SceneView<T>>
__global__ void __launch_bounds__(BLOCK) scatter_batch_kernel(BatchParams<T, SCENE> P) {
    __shared__ int32_t stack_s[kStackDepth * BLOCK];
    size_t idx = (size_t)blockIdx.x * BLOCK + threadIdx.x;
    if (idx >= P.n) return;
    Ray<T> r{load3<T>(P.o, idx), load3<T>(P.d, idx)};
    Hit<T> h;
    Tally tl;
    V3<T> zero = mk<T>(0, 0, 0);
    if (!closest_hit<T, EXACT, false, SCENE>(P.scene, r, P.tmin, M<T, EXACT>::inf(), &h, stack_s + threadIdx.x, BLOCK, tl)) {
        P.prim[idx] = -1; P.t[idx] = __builtin_huge_val(); P.kind[idx] = V_MISS;
        store3(P.p, idx, zero); store3(P.normal, idx, zero); store3(P.dir, idx, zero); store3(P.weight, idx, zero);
        return;
    }
    Stream<EXACT> rng(P.seed, P.a[idx], P.b[idx], P.c[idx]);
    Ray<T> next{zero, zero};
    V3<T> w = zero;
    uint32_t kind = shade<T, EXACT, false, SCENE>(P.scene, r, h, rng, &next, &w, tl);
    P.prim[idx] = (int32_t)(h.info >> 2); P.t[idx] = (double)h.t; P.kind[idx] = kind;
    store3(P.p, idx, h.p); store3(P.normal, idx, h.normal);
    store3(P.dir, idx, kind >= V_SPECULAR ? next.d : zero);
    store3(P.weight, idx, kind >= V_SPECULAR ? w : zero);
}

// Material::scatter + the mixture-pdf sample on CALLER-SUPPLIED hit records (rtw_shade_batch): the sampling arithmetic alone, with
// no dependence on how each side traced its hit point.  in: incoming direction, p, normal, front_face, material; stream keys.
template <class T> struct ShadeParams {
    SceneView<T> scene;
    uint64_t seed;
    size_t n;
    const double *d, *p, *normal, *material;     // [n][3] x3, [n][4] (albedo r, g, b, param)
    const uint32_t *front_face, *mat_kind, *pixel, *sample, *vertex;
    uint32_t* kind; double *dir, *weight;
};
template <class T, bool EXACT, int BLOCK>
__global__ void __launch_bounds__(BLOCK) shade_batch_kernel(ShadeParams<T> P) {
    size_t idx = (size_t)blockIdx.x * BLOCK + threadIdx.x;
    if (idx >= P.n) return;
    Hit<T> h;
    h.p = load3<T>(P.p, idx); h.normal = load3<T>(P.normal, idx); h.t = T(0);
    h.front_face = P.front_face[idx] != 0;
    h.info = P.mat_kind[idx] & 3u;
    h.albedo = mk<T>((T)P.material[4 * idx], (T)P.material[4 * idx + 1], (T)P.material[4 * idx + 2]); h.param = (T)P.material[4 * idx + 3];
    h.gkind = h.info;
    V3<T> zero = mk<T>(0, 0, 0);
    Ray<T> r{h.p, load3<T>(P.d, idx)};                       // Material::scatter reads only the direction of the incoming ray
    Stream<EXACT> rng(P.seed, P.pixel[idx], P.sample[idx], P.vertex[idx]);
    Ray<T> next{zero, zero};
    V3<T> w = zero;
    Tally tl;
    uint32_t kind = shade<T, EXACT, false, SceneView<T>>(P.scene, r, h, rng, &next, &w, tl);
    P.kind[idx] = kind;
    store3(P.dir, idx, kind >= V_SPECULAR ? next.d : zero);
    store3(P.weight, idx, kind >= V_SPECULAR ? w : zero);
}

template <class T, bool EXACT, int BLOCK>
__global__ void __launch_bounds__(BLOCK) get_rays_kernel(BatchParams<T> P, double* o, double* d) {
    size_t idx = (size_t)blockIdx.x * BLOCK + threadIdx.x;
    if (idx >= P.n) return;
    uint32_t i = P.a[idx], j = P.b[idx], s = P.c[idx];
    Stream<EXACT> rng(P.seed, j * P.cam.width + i, s, 0u);
    Ray<T> r = get_ray<T, EXACT>(P.cam, i, j, rng);
    store3(o, idx, r.o); store3(d, idx, r.d);
}

template <class T, bool EXACT, int BLOCK, class SCENE = SceneView<T>>
__global__ void __launch_bounds__(BLOCK) path_radiance_kernel(BatchParams<T, SCENE> P) {
    __shared__ int32_t stack_s[kStackDepth * BLOCK];
    size_t idx = (size_t)blockIdx.x * BLOCK + threadIdx.x;
    if (idx >= P.n) return;
    uint32_t i = P.a[idx], j = P.b[idx], s = P.c[idx];
    uint32_t pixel = j * P.cam.width + i;
    Stream<EXACT> rng(P.seed, pixel, s, 0u);
    PathState<T> ps;
    ps.r = get_ray<T, EXACT>(P.cam, i, j, rng);
    ps.mult = mk<T>(1, 1, 1); ps.res = mk<T>(0, 0, 0); ps.depth = P.cam.max_depth;
    V3<T> value;
    uint32_t nrays = 0;
    Tally tl;
    while (!path_step<T, EXACT, false, SCENE>(P.scene, P.cam, P.seed, P.tmin, pixel, s, ps, &value, stack_s + threadIdx.x, BLOCK, nrays, tl)) {}
    if (P.flags & 1u) value = fix_nan(value);
    store3(P.rgb, idx, value);
}

// ---------------------------------------------------------------------------------------------
// Untile + resolve: [world][tiles_per_rank][kTileH][kTileW][3] (T) -> rgb_sum f64 and / or rgb8.
// Colour::write_colour (shared/src/colour.rs:15-36): c/spp -> sqrt -> clamp(0,1) -> (256*x) as u8
// (`as u8` saturates and maps NaN to 0).  Done in f64 on the widened value in both precisions.
template <class T>
__global__ void untile_resolve_kernel(const T* tiles, uint32_t width, uint32_t height, uint32_t world, uint32_t tiles_per_rank,
                                      uint32_t tiles_x, uint32_t spp, double* rgb_sum, uint8_t* rgb8) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= width || j >= height) return;
    uint32_t tile = tile_slot(i / kTileW, j / kTileH, tiles_x);
    uint32_t rank = tile % world, local = tile / world;
    const T* src = tiles + (((size_t)rank * tiles_per_rank + local) * (kTileW * kTileH) + (size_t)(j % kTileH) * kTileW + (i % kTileW)) * 3;
    size_t dst = ((size_t)j * width + i) * 3;
    double scale = 1. / (double)(int32_t)spp;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        double v = (double)src[c];
        if (rgb_sum) rgb_sum[dst + c] = v;
        if (rgb8) {
            double g = sqrt(v * scale);
            if (g < 0.) g = 0.;
            if (g > 1.) g = 1.;
            double q = 256. * g;
            uint8_t b = (q != q) ? 0 : (q >= 255. ? 255 : (q <= 0. ? 0 : (uint8_t)q));
            rgb8[dst + c] = b;
        }
    }
}

// Sample partition: fixed-point accumulators (+ poison words) of the WHOLE image, e.g. the sum of several ranks' partial
// accumulators, -> rgb_sum / rgb8.  pool_finalize_kernel + untile_resolve_kernel in one pass, same arithmetic.
template <int UNUSED = 0>
__global__ void resolve_accum_kernel(const unsigned long long* accum, const uint32_t* poison, uint32_t width, uint32_t height, uint32_t tiles_x,
                                     uint32_t spp, double* rgb_sum, uint8_t* rgb8) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= width || j >= height) return;
    uint32_t tile = tile_slot(i / kTileW, j / kTileH, tiles_x);
    size_t q = (size_t)tile * (kTileW * kTileH) + (size_t)(j % kTileH) * kTileW + (i % kTileW);
    uint32_t bad = poison[q];
    size_t dst = ((size_t)j * width + i) * 3;
    double scale = 1. / (double)(int32_t)spp;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float f = (float)((double)accum[3 * q + c] * (1.0 / 4294967296.0));
        if (poison_has_inf(bad, c)) f = __int_as_float(0x7f800000);
        if (poison_has_nan(bad, c)) f = __int_as_float(0x7fc00000);
        double v = (double)f;
        if (rgb_sum) rgb_sum[dst + c] = v;
        if (rgb8) {
            double g = sqrt(v * scale);
            if (g < 0.) g = 0.;
            if (g > 1.) g = 1.;
            double qq = 256. * g;
            rgb8[dst + c] = (qq != qq) ? 0 : (qq >= 255. ? 255 : (qq <= 0. ? 0 : (uint8_t)qq));
        }
    }
}

// N GPUs in one process: the frame's collective and its resolve as ONE kernel per GPU over NVLink peer memory.  After the sample
// partition every GPU holds fixed-point accumulators of the WHOLE image for its share of the samples.  GPU g owns pixel slots
// [slot_begin, slot_end): it reads those slots from every GPU's accumulators through peer pointers (reduce-scatter), resolves them
// (same arithmetic as resolve_accum_kernel) and stores the pixels straight into the image buffers on the root GPU (gather).  No
// intermediate buffer, no second pass over the 58 MB blocks, and every link carries 1 / N of a block.
constexpr int kMaxPeers = 16;
struct PeerBlocks {
    const unsigned long long* accum[kMaxPeers];
    const uint32_t* poison[kMaxPeers];
    int n;
};
template <int UNUSED = 0>
__global__ void peer_reduce_resolve_kernel(PeerBlocks B, uint32_t slot_begin, uint32_t slot_end, uint32_t width, uint32_t height,
                                           uint32_t tiles_x, uint32_t tiles_total, uint32_t spp, double* rgb_sum, uint8_t* rgb8) {
    uint32_t q = slot_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= slot_end) return;
    uint32_t tile = q >> 8, in = q & 255u;
    if (tile >= tiles_total) return;
    uint32_t ttx, tty;
    slot_tile(tile, tiles_x, &ttx, &tty);
    uint32_t i = ttx * kTileW + (in & 15u), j = tty * kTileH + (in >> 4);
    if (i >= width || j >= height) return;
    unsigned long long a[3] = {0ull, 0ull, 0ull};
    uint32_t bad = 0;
    for (int r = 0; r < B.n; ++r) {
        const unsigned long long* src = B.accum[r] + 3 * (size_t)q;
        a[0] += src[0]; a[1] += src[1]; a[2] += src[2];
        bad |= B.poison[r][q];
    }
    size_t dst = ((size_t)j * width + i) * 3;
    double scale = 1. / (double)(int32_t)spp;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float f = (float)((double)a[c] * (1.0 / 4294967296.0));
        if (poison_has_inf(bad, c)) f = __int_as_float(0x7f800000);
        if (poison_has_nan(bad, c)) f = __int_as_float(0x7fc00000);
        double v = (double)f;
        if (rgb_sum) rgb_sum[dst + c] = v;
        if (rgb8) {
            double g = sqrt(v * scale);
            if (g < 0.) g = 0.;
            if (g > 1.) g = 1.;
            double qq = 256. * g;
            rgb8[dst + c] = (qq != qq) ? 0 : (qq >= 255. ? 255 : (qq <= 0. ? 0 : (uint8_t)qq));
        }
    }
}

}  // namespace rtw

#include "rtw_general.cuh"
