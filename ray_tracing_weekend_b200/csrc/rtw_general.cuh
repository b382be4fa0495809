// rtw_general.cuh — general scenes: Quad, Triangle, Cuboid, Transformed<T>, Plane and Sphere
// entries in one primitive table under one BVH, DiffuseLight / Isotropic materials, NoiseTexture (Perlin), lights
// lists that hold quads.  Included at the end of rtw_kernels.cuh; the render / batch kernels there are generic over
// the scene view and pick these functions up through closest_hit / shade / emitted_of.
// The sphere-only SceneView path (scenes::simple, the headline workload) does not go through this file.
#pragma once

namespace rtw {

// One scene record.  The exact path reads it in place.  The FP32 kernels fetch it with 128-bit non-coherent loads (LDG.E.128.CONSTANT):
// the records are read-only, 16-byte aligned and whole 16-byte chunks long (rtw_device.cuh), while member-by-member reads through a
// `const R&` compiled to one generic 32-bit load per float — 16 per tested quad, 12 per list entry (SASS of round 1: 247 scalar loads
// in the kernel) — in kernels that sit at the instruction-cache knee.
template <class T> struct GRec {
    template <class R> static RTW_D const R& get(const R* p) { return *p; }
};
#ifndef RTW_G_SCALAR_LOADS
template <> struct GRec<float> {
    template <class R> static RTW_D R get(const R* p) {
        static_assert(sizeof(R) % 16 == 0 && alignof(R) >= 16, "general-scene records are whole 16-byte chunks");
        union U { R r; float4 q[sizeof(R) / 16]; RTW_D U() {} } u;
        const float4* src = reinterpret_cast<const float4*>(p);
#pragma unroll
        for (int i = 0; i < (int)(sizeof(R) / 16); ++i) u.q[i] = ldg128(src + i);
        return u.r;
    }
};
#endif

// ---- Transformed<T> (entities/transformations.rs:14-29, geometry/src/transformations.rs:118-126) ----------------
template <class T> RTW_D V3<T> g_mat_vec(const T* m, V3<T> v) {                      // matrix3.rs:88-100: row . v
    return mk<T>(dot(mk<T>(m[0], m[1], m[2]), v), dot(mk<T>(m[3], m[4], m[5]), v), dot(mk<T>(m[6], m[7], m[8]), v));
}
template <class T> RTW_D Ray<T> g_instance_ray(const GXform<T>& X, const Ray<T>& r) {
    V3<T> it = mk<T>(X.it[0], X.it[1], X.it[2]);
    return Ray<T>{g_mat_vec<T>(X.inv, r.o) + it, g_mat_vec<T>(X.inv, r.d) + it};       // transform_vector3d adds the translation too
}

// Flat entities are where self-intersection ("acne") is decided: a ray leaving a quad re-hits it iff the rounding noise of
// dot(p - q, n) has the right sign and exceeds tmin * |denom|.  Those statistics depend on WHERE roundings happen, so the
// plane equation and the hit point are evaluated with the reference's unfused operation sequence in both precisions
// (dot = (x*x + y*y) + z*z, p = o + d*t) — the translation units are compiled with -fmad=false.
template <class T> RTW_D T g_dot(V3<T> a, V3<T> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <class T> RTW_D V3<T> g_at(const Ray<T>& r, T t) { return mk<T>(r.o.x + r.d.x * t, r.o.y + r.d.y * t, r.o.z + r.d.z * t); }

// AABBox::hit (hittable.rs:38-87) in the working precision, min / max form (handles the infinite slabs of a Plane's box)
template <class T, bool EXACT>
RTW_D bool g_box_hit(const T* mn, const T* mx, const Ray<T>& r, T start, T end) {
    using Mt = M<T, EXACT>;
    // FP32: one SFU reciprocal per axis instead of two IEEE divisions (six ~10-instruction sequences per call); same handling of
    // zero direction components ((mn - o) * inf = +-inf, NaN only where the division gives 0 / 0 as well)
    T ix, iy, iz;
    if constexpr (EXACT) { ix = iy = iz = T(0); } else { ix = frcp(r.d.x); iy = frcp(r.d.y); iz = frcp(r.d.z); }
    auto slab = [](T plane, T o, T d, T inv) { if constexpr (EXACT) return (plane - o) / d; else return (plane - o) * inv; };
    T x0 = slab(mn[0], r.o.x, r.d.x, ix), x1 = slab(mx[0], r.o.x, r.d.x, ix);
    if (signbit(r.d.x)) { T s = x0; x0 = x1; x1 = s; }
    T tmin = x0, tmax = x1;
    T y0 = slab(mn[1], r.o.y, r.d.y, iy), y1 = slab(mx[1], r.o.y, r.d.y, iy);
    if (signbit(r.d.y)) { T s = y0; y0 = y1; y1 = s; }
    if (tmax < y0 || tmin > y1) return false;
    tmin = Mt::max_(tmin, y0); tmax = Mt::min_(tmax, y1);
    T z0 = slab(mn[2], r.o.z, r.d.z, iz), z1 = slab(mx[2], r.o.z, r.d.z, iz);
    if (signbit(r.d.z)) { T s = z0; z0 = z1; z1 = s; }
    if (tmax < z0 || tmin > z1) return false;
    tmin = Mt::max_(tmin, z0); tmax = Mt::min_(tmax, z1);
    return Mt::max_(start, tmin) <= Mt::min_(end, tmax);
}

// ---- Quad::hit / Triangle::hit (quadrilateral.rs:79-98, triangles.rs:74-92) --------------------------------------
template <class T, bool EXACT>
RTW_D bool g_quad_hit(const GQuad<T>& Q, bool tri, const Ray<T>& r, T tmin, T tmax, T* t_out) {
    T denom = g_dot(r.d, Q.normal);
    if (!(fabs(denom) > M<T, EXACT>::EPS)) return false;
    // The IEEE division stays, also in FP32 (7 % of the kernel's instructions): with a 1-ulp SFU reciprocal the hit point lands a little
    // further from the plane, the self-intersection statistics of the NEXT ray change, and cornell_box comes out 3 % off the f64 mean
    // (and 12 % faster, for the wrong reason) — tried and reverted, test_cornell_box_f32_image_statistics catches it.
    T t = -(g_dot(r.o - Q.q, Q.normal) / denom);
    if (!(tmin <= t && t <= tmax)) return false;
    V3<T> pq = g_at(r, t) - Q.q;
    T a = dot(cross(pq, Q.v), Q.w), b = dot(cross(Q.u, pq), Q.w);                   // get_quad_uv
    bool inside = tri ? (T(0) <= a + b && a + b <= T(1)) : (T(0) <= a && a <= T(1) && T(0) <= b && b <= T(1));
    if (!inside) return false;
    *t_out = t;
    return true;
}

// Hittable::hit of one list entry in the entity's own space; *sub = the quad that was hit (cuboid face)
template <class T, bool EXACT, bool COUNT>
RTW_D bool g_prim_hit(const SceneViewG<T>& sc, const GPrim<T>& pr, const Ray<T>& r, T tmin, T tmax, T* t_out, uint32_t* sub, Tally& tl) {
    Ray<T> rr = r;
    if (pr.xform >= 0) { const auto& X = GRec<T>::get(sc.xforms + pr.xform); rr = g_instance_ray<T>(X, r); }
    *sub = pr.first;
    switch (pr.kind) {
    case P_SPHERE: {
        if (COUNT) tl.sphere_tests++;
        T a = sqlen(rr.d);
        if constexpr (EXACT) return sphere_root<T>(sc.spheres[pr.first], rr, a, tmin, tmax, t_out);
        else return sphere_root_fast(sc.spheres[pr.first], rr, frcp(a), tmin, tmax, t_out);
    }
    case P_PLANE: {                                                                     // plane.rs:61-76, one-sided
        const auto& pl = GRec<T>::get(sc.plane_geo + pr.first);
        T denom = g_dot(rr.d, pl.normal);
        if (!(denom > M<T, EXACT>::EPS)) return false;
        T t = -g_dot(rr.o - pl.point, pl.normal) / denom;
        if (!(tmin <= t && t <= tmax)) return false;
        *t_out = t;
        return true;
    }
    default: {
        // (Tried on FP32: the cuboid as three slab axes with the faces' exact t — ~170 instead of ~600 instructions per tested box, results
        // equal — but the extra 1.9 KB in the hot loop pushed the executed code back over the instruction cache: cornell_box 259 -> 286 ms.)
        // Cuboid (cuboid.rs:53-60): first minimum of the six faces; a Quad / Triangle is the same loop over one face — ONE inlined
        // copy of the quad test per call site instead of two (code size: these kernels wait on instruction fetch)
        const uint32_t faces = pr.kind == P_CUBOID ? 6u : 1u;
        const bool tri = pr.kind == P_TRIANGLE;
        bool any = false; T best = T(0);
#pragma unroll 1
        for (uint32_t f = 0; f < faces; ++f) {
            T t;
            const auto& Q = GRec<T>::get(sc.quads + pr.first + f);
            if (g_quad_hit<T, EXACT>(Q, tri, rr, tmin, tmax, &t) && (!any || t < best)) { any = true; best = t; *sub = pr.first + f; }
        }
        *t_out = best;
        return any;
    }
    }
}

// sin of a moderate argument as a fixed IEEE sequence (oracle: sin_portable) / FP32: sinf
RTW_D double g_sin(double x) {
    const double two_over_pi = 6.36619772367581382433e-01;
    const double pio2_1 = 1.57079632673412561417e+00, pio2_1t = 6.07710050650619224932e-11;
    double fn = floor(x * two_over_pi + 0.5);
    int n = (int)fn;
    double y = (x - fn * pio2_1) - fn * pio2_1t;
    double z = y * y;
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04,
                 S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05,
                 C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    double ps = S1 + z * (S2 + z * (S3 + z * (S4 + z * (S5 + z * S6))));
    double pc = C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6))));
    double sy = y + (y * z) * ps;
    double cy = (1. - 0.5 * z) + (z * z) * pc;
    switch (n & 3) { case 0: return sy; case 1: return cy; case 2: return -sy; default: return -cy; }
}
RTW_D float g_sin(float x) { return sinf(x); }

// Perlin::noise / turb (perlin.rs:59-110)
// f.rem_euclid(256.) as usize for an integer-valued f.  fmod by a power of two is f - 256 * trunc(f / 256) with every step exact
// (same bits as fmod, NaN for +-inf); spelled out because CUDA's inlined fmod is ~58 instructions and Perlin::noise needs it 24 times —
// the 49 copies were 45 KB of the 146 KB general kernels, whose ncu captures show them waiting on instruction fetch (i-cache hit 75 %).
template <class T> RTW_D int g_wrap256(T f) { T r = f - trunc(f * T(0.00390625)) * T(256); if (r < T(0)) r += T(256); return (int)r; }
template <class T> RTW_D T g_noise(const GPerlin<T>& pn, V3<T> p) {
    T fx = floor(p.x), fy = floor(p.y), fz = floor(p.z);
    T u = p.x - fx, v = p.y - fy, w = p.z - fz;
    T acc = T(0);
    const int px[2] = {pn.perm_x[g_wrap256(fx)], pn.perm_x[g_wrap256(fx + T(1))]};
    const int py[2] = {pn.perm_y[g_wrap256(fy)], pn.perm_y[g_wrap256(fy + T(1))]};
    const int pz[2] = {pn.perm_z[g_wrap256(fz)], pn.perm_z[g_wrap256(fz + T(1))]};
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                int idx = px[i] ^ py[j] ^ pz[k];
                V3<T> c = mk<T>(pn.rand_vec[idx][0], pn.rand_vec[idx][1], pn.rand_vec[idx][2]);
                T di = (T)i, dj = (T)j, dk = (T)k;
                V3<T> weight_v = mk<T>(u - di, v - dj, w - dk);
                T term = (di * u + (T(1) - di) * (T(1) - u)) * (dj * v + (T(1) - dj) * (T(1) - v)) * (dk * w + (T(1) - dk) * (T(1) - w)) * dot(c, weight_v);
                acc = acc + term;
            }
    return acc;
}
template <class T> RTW_D T g_turb(const GPerlin<T>& pn, V3<T> p, int depth) {
    T accum = T(0), weight = T(1);
    V3<T> tp = p;
    for (int i = 0; i < depth; ++i) { accum += weight * g_noise<T>(pn, tp); tp = tp * T(2); weight *= T(0.5); }
    return accum;
}
// atan2 / acos of Sphere::get_sphere_uv (sphere.rs:49-54).  f64: libm::atan2 of the reference is the `libm` crate's port of
// musl / msun (a fixed IEEE sequence, restated here); f64::acos is the platform libm's — msun's e_acos.c sequence stands in
// (oracle: atan2_msun / acos_msun).  FP32: CUDA's atan2f / acosf.
RTW_D double g_atan(double x) {
    const double atanhi[4] = {4.63647609000806093515e-01, 7.85398163397448278999e-01, 9.82793723247329054082e-01, 1.57079632679489655800e+00};
    const double atanlo[4] = {2.26987774529616870924e-17, 3.06161699786838301793e-17, 1.39033110312309984516e-17, 6.12323399573676603587e-17};
    const double aT[11] = {3.33333333333329318027e-01, -1.99999999998764832476e-01, 1.42857142725034663711e-01, -1.11111104054623557880e-01,
                           9.09088713343650656196e-02, -7.69187620504482999495e-02, 6.66107313738753120669e-02, -5.83357013379057348645e-02,
                           4.97687799461593236017e-02, -3.65315727442169155270e-02, 1.62858201153657823623e-02};
    uint32_t ix = (uint32_t)__double2hiint(x);
    const bool sign = (ix >> 31) != 0;
    ix &= 0x7fffffffu;
    int id;
    if (ix >= 0x44100000u) {
        if (x != x) return x;
        double z = atanhi[3] + 0x1p-120;
        return sign ? -z : z;
    }
    if (ix < 0x3fdc0000u) {
        if (ix < 0x3e400000u) return x;
        id = -1;
    } else {
        x = fabs(x);
        if (ix < 0x3ff30000u) {
            if (ix < 0x3fe60000u) { id = 0; x = (2.0 * x - 1.0) / (2.0 + x); }
            else { id = 1; x = (x - 1.0) / (x + 1.0); }
        } else {
            if (ix < 0x40038000u) { id = 2; x = (x - 1.5) / (1.0 + 1.5 * x); }
            else { id = 3; x = -1.0 / x; }
        }
    }
    double z = x * x, w = z * z;
    double s1 = z * (aT[0] + w * (aT[2] + w * (aT[4] + w * (aT[6] + w * (aT[8] + w * aT[10])))));
    double s2 = w * (aT[1] + w * (aT[3] + w * (aT[5] + w * (aT[7] + w * aT[9]))));
    if (id < 0) return x - x * (s1 + s2);
    z = atanhi[id] - (x * (s1 + s2) - atanlo[id] - x);
    return sign ? -z : z;
}
RTW_D double g_atan2(double y, double x) {
    const double pi = 3.1415926535897931160E+00, pi_lo = 1.2246467991473531772E-16;
    if (x != x || y != y) return x + y;
    uint32_t ix = (uint32_t)__double2hiint(x), lx = (uint32_t)__double2loint(x), iy = (uint32_t)__double2hiint(y), ly = (uint32_t)__double2loint(y);
    if (((ix - 0x3ff00000u) | lx) == 0) return g_atan(y);
    uint32_t m = ((iy >> 31) & 1u) | ((ix >> 30) & 2u);
    ix &= 0x7fffffffu; iy &= 0x7fffffffu;
    if ((iy | ly) == 0) { switch (m) { case 0: case 1: return y; case 2: return pi; default: return -pi; } }
    if ((ix | lx) == 0) return (m & 1u) ? -pi / 2 : pi / 2;
    if (ix == 0x7ff00000u) {
        if (iy == 0x7ff00000u) { switch (m) { case 0: return pi / 4; case 1: return -pi / 4; case 2: return 3 * pi / 4; default: return -3 * pi / 4; } }
        switch (m) { case 0: return 0.0; case 1: return -0.0; case 2: return pi; default: return -pi; }
    }
    if (ix + (64u << 20) < iy || iy == 0x7ff00000u) return (m & 1u) ? -pi / 2 : pi / 2;
    double z = ((m & 2u) && iy + (64u << 20) < ix) ? 0.0 : g_atan(fabs(y / x));
    switch (m) { case 0: return z; case 1: return -z; case 2: return pi - (z - pi_lo); default: return (z - pi_lo) - pi; }
}
RTW_D double g_acos(double x) {
    const double pio2_hi = 1.57079632679489655800e+00, pio2_lo = 6.12323399573676603587e-17;
    const double pS0 = 1.66666666666666657415e-01, pS1 = -3.25565818622400915405e-01, pS2 = 2.01212532134862925881e-01,
                 pS3 = -4.00555345006794114027e-02, pS4 = 7.91534994289814532176e-04, pS5 = 3.47933107596021167570e-05,
                 qS1 = -2.40339491173441421878e+00, qS2 = 2.02094576023350569471e+00, qS3 = -6.88283971605453293030e-01, qS4 = 7.70381505559019352791e-02;
    auto R = [&](double z) {
        double p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        double q = 1.0 + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        return p / q;
    };
    uint32_t hx = (uint32_t)__double2hiint(x), ix = hx & 0x7fffffffu;
    if (ix >= 0x3ff00000u) {
        if (((ix - 0x3ff00000u) | (uint32_t)__double2loint(x)) == 0) return (hx >> 31) ? 2 * pio2_hi + 0x1p-120 : 0.0;
        return 0.0 / (x - x);
    }
    if (ix < 0x3fe00000u) {
        if (ix <= 0x3c600000u) return pio2_hi + 0x1p-120;
        return pio2_hi - (x - (pio2_lo - x * R(x * x)));
    }
    if (hx >> 31) {
        double z = (1.0 + x) * 0.5, sq = sqrt(z), w = R(z) * sq - pio2_lo;
        return 2 * (pio2_hi - (sq + w));
    }
    double z = (1.0 - x) * 0.5, sq = sqrt(z);
    double df = __hiloint2double(__double2hiint(sq), 0);
    double c = (z - df * df) / (sq + df), w = R(z) * sq + c;
    return 2 * (df + w);
}
RTW_D float g_atan2(float y, float x) { return atan2f(y, x); }
RTW_D float g_acos(float x) { return acosf(x); }

// Sphere::get_sphere_uv (sphere.rs:49-54).  Out of line on purpose: only CheckerTexture hits reach it, and inlining the
// atan2 / acos sequences into the hit record cost every scene 13 % (FP32) to 37 % (f64) in registers and code size.
template <class T, bool EXACT> __device__ __noinline__ void g_sphere_uv(T nx, T ny, T nz, T* u, T* v) {
    *u = g_atan2(-nz, nx) / (T(2) * M<T, EXACT>::PI);
    *v = g_acos(ny) / M<T, EXACT>::PI;
}

// Texture::get_colour (texture.rs:15-22, 46-55, 90-102).  (u, v) is only read by a CheckerTexture.
template <class T> RTW_D V3<T> g_noise_colour(const SceneViewG<T>& sc, const GTex<T>& t, V3<T> point) {
    T arg = t.scale * point.z + g_turb<T>(sc.perlins[t.perlin], point, 7) * T(10);
    return mk<T>(T(0.5), T(0.5), T(0.5)) * (g_sin(arg) + T(1));
}
// Out of line: SolidColour is the common case, and two inlined copies of the 7-octave Perlin loops in every hit record
// slowed all general kernels down by 13 % (instruction footprint).
// CheckerTexture's even / odd are textures themselves (Arc<dyn Texture>, texture.rs:26-29): get_colour recurses with the same
// (u, v, point).  References point to EARLIER table entries (checked at scene creation), so the walk ends.  The reference's scenes
// nest one level (a checker of colours or noise), which g_texture_lookup handles itself; deeper trees continue here — kept in a
// function of its own because the register needs of g_texture_lookup shape ptxas's allocation in every kernel that calls it
// (with the loop inside it, cornell_box — which has no texture at all — rendered 7 % slower).
template <class T> __device__ __noinline__ V3<T> g_texture_nested(const SceneViewG<T>& sc, const GTex<T>* t, T u, T v, V3<T> point) {
    while (t->kind != TEX_NOISE) {
        T inv_scale = T(1) / t->scale;
        T cells = floor(u * inv_scale) + floor(v * inv_scale);
        bool is_even = cells - trunc(cells * T(0.5)) * T(2) == T(0);
        uint32_t ref = is_even ? t->even : t->odd;
        if (ref == 0) return is_even ? mk<T>(t->even_c[0], t->even_c[1], t->even_c[2]) : mk<T>(t->odd_c[0], t->odd_c[1], t->odd_c[2]);
        t = &sc.textures[ref - 1];
    }
    return g_noise_colour<T>(sc, *t, point);
}
template <class T> __device__ __noinline__ V3<T> g_texture_lookup(const SceneViewG<T>& sc, const GMat<T>& m, T u, T v, V3<T> point) {
    const GTex<T>& t = sc.textures[m.texture - 1];
    if (t.kind == TEX_NOISE) return g_noise_colour<T>(sc, t, point);
    T inv_scale = T(1) / t.scale;
    T cells = floor(u * inv_scale) + floor(v * inv_scale);
    bool is_even = cells - trunc(cells * T(0.5)) * T(2) == T(0);            // fmod(cells, 2) == 0, exactly (see g_wrap256)
    uint32_t ref = is_even ? t.even : t.odd;
    if (ref == 0) return is_even ? mk<T>(t.even_c[0], t.even_c[1], t.even_c[2]) : mk<T>(t.odd_c[0], t.odd_c[1], t.odd_c[2]);
    const GTex<T>& sub = sc.textures[ref - 1];
    if (sub.kind != TEX_NOISE) return g_texture_nested<T>(sc, &sub, u, v, point);
    return g_noise_colour<T>(sc, sub, point);
}
template <class T> RTW_D V3<T> g_texture(const SceneViewG<T>& sc, const GMat<T>& m, T u, T v, V3<T> point) {
    if (m.texture == 0) return mk<T>(m.albedo[0], m.albedo[1], m.albedo[2]);
    return g_texture_lookup<T>(sc, m, u, v, point);
}

// ---- closest hit over the planes + the BVH of bounded entries ----------------------------------------------------
// The winner as two words: code >= 0: index into sc.prims, code <= -2: unbounded entry -2 - code; sub = the quad that was hit.
template <class T> RTW_D const GPrim<T>* g_entry(const SceneViewG<T>& sc, int32_t code) { return code >= 0 ? sc.prims + code : sc.unbounded + (-2 - code); }

template <class T, bool EXACT, bool COUNT>
RTW_D bool g_closest_entry(const SceneViewG<T>& sc, const Ray<T>& r, T tmin, T tmax, const GPrim<T>** best_out, uint32_t* sub_out, T* t_out,
                           int32_t* stack, int stride, Tally& tl) {
    bool found = false;
    T best_t = tmax;
    const GPrim<T>* best = nullptr;
    uint32_t best_sub = 0;
    RayAux aux;
    if constexpr (!EXACT) ray_aux(r, aux);
    // ONE entry-test loop serves the three places entries come from — the unbounded list, the flat list, the BVH leaves — so the
    // entity tests are inlined once (the kernels stall on instruction fetch; three copies of every entity test did not help).
    // Entries are addressed in one index space: [0, n_unbounded) = sc.unbounded, n_unbounded + k = sc.prims[k].  Order of the
    // tests (it decides exact-t ties: first minimum): unbounded entries, then the list / the leaves in traversal order.
    const int nu = sc.n_unbounded;
    // a handful of entries (cornell_box: 8): every lane walks the whole list, sorted by kind on the host, in lockstep — no tree, no
    // per-lane leaf order, the kind switch is warp-uniform (ncu on the BVH version of this scene: leaf tests ran at 2-7 active
    // threads per warp)
    int i = 0, end = nu + (sc.flat ? sc.n_prims : 0);
    stack[0] = kStop;
    int sp = 1;
    int32_t cur = (sc.flat || sc.n_prims <= 0) ? kStop : 0;
    for (;;) {
#pragma unroll 1
        for (; i < end; ++i) {
            const bool unb = i < nu;
            const GPrim<T>* pp = unb ? sc.unbounded + i : sc.prims + (i - nu);
            const auto& pr = GRec<T>::get(pp);
            // The entry's own box, where it is part of the result.  Unbounded entries: bounded_hit (hittable.rs:191-196) with
            // Plane::get_aabbox (plane.rs:78-107) — an axis-aligned plane's box is the slab {axis = 0} whatever the plane's offset,
            // tested in both precisions.  Bounded entries: the exact path repeats bounded_hit with the un-shrunk range; the fast
            // path only for a Transformed<T>, which is hit with a DIFFERENT ray than its box (the instance ray's direction carries
            // the translation), so its world-space box is not only a culling aid.
            if constexpr (EXACT) {
                if (unb ? !g_box_hit<T, EXACT>(pr.box, pr.box + 3, r, tmin, tmax) : !box_hit_exact(pr.box, pr.box + 3, r, tmin, tmax)) continue;
            } else {
                if ((unb || pr.xform >= 0) && !g_box_hit<T, EXACT>(pr.box, pr.box + 3, r, tmin, tmax)) continue;
            }
            T t; uint32_t sub;
            if (g_prim_hit<T, EXACT, COUNT>(sc, pr, r, tmin, tmax, &t, &sub, tl) && (!found || t < best_t)) {
                found = true; best_t = t; best = pp; best_sub = sub;
            }
        }
        while (cur >= 0) {
            Node<T> nd = sc.nodes[cur];
            if (COUNT) tl.node_visits++;
            bool hl, hr; T tl_near = T(0), tr_near = T(0);
            // A Transformed<T> reports t along its INSTANCE ray, whose direction carries the inverse translation
            // (transformations.rs:123-126): that t says nothing about where the world ray meets the entry's box, so with
            // transforms in the scene the best t so far must not cull nodes (the reference never shrinks the range).
            const T far_limit = sc.has_xforms ? tmax : best_t;
            if constexpr (EXACT) {
                hl = box_hit_exact(nd.la, nd.lb, r, tmin, far_limit);
                hr = box_hit_exact(nd.ra, nd.rb, r, tmin, far_limit);
            } else {
                hl = box_hit_fast(nd.la, nd.lb, aux, tmin, far_limit, &tl_near);
                hr = box_hit_fast(nd.ra, nd.rb, aux, tmin, far_limit, &tr_near);
            }
            int32_t l = nd.left, rr = nd.right;
            if (hl && hr) {
                bool swap = !EXACT && tr_near < tl_near;
                stack[sp * stride] = swap ? l : rr; sp++;
                cur = swap ? rr : l;
            } else if (hl) cur = l;
            else if (hr) cur = rr;
            else { sp--; cur = stack[sp * stride]; }
        }
        if (cur == kStop) break;
        if (cur != kEmptyLeaf) {                                 // a leaf: its entries are the next range of the test loop
            uint32_t enc = (uint32_t)~cur;
            i = nu + (int)(enc >> 4); end = i + (int)(enc & 15u) + 1;
        }
        sp--;
        cur = stack[sp * stride];
    }
    if (!found) return false;
    *best_out = best;
    *sub_out = best_sub;
    *t_out = best_t;
    return true;
}
// the same with the winner as a code (what the wavefront stores in a path slot)
template <class T, bool EXACT, bool COUNT>
RTW_D bool g_closest_prim(const SceneViewG<T>& sc, const Ray<T>& r, T tmin, T tmax, int32_t* code_out, uint32_t* sub_out, T* t_out,
                          int32_t* stack, int stride, Tally& tl) {
    const GPrim<T>* best;
    if (!g_closest_entry<T, EXACT, COUNT>(sc, r, tmin, tmax, &best, sub_out, t_out, stack, stride, tl)) return false;
    *code_out = (best >= sc.prims && best < sc.prims + sc.n_prims) ? (int32_t)(best - sc.prims) : -2 - (int32_t)(best - sc.unbounded);
    return true;
}

// the normal alone: words 12..14 of the record (one 128-bit load in FP32)
template <class T> RTW_D V3<T> g_quad_normal(const GQuad<T>* Q) { return Q->normal; }
#ifndef RTW_G_SCALAR_LOADS
template <> RTW_D V3<float> g_quad_normal<float>(const GQuad<float>* Q) {
    const float4 v = ldg128(reinterpret_cast<const float4*>(Q) + 3);
    return mk<float>(v.x, v.y, v.z);
}
#endif
// (u, v) of a hit, read by CheckerTexture only — out of line like the texture lookup itself (code size, see g_texture_lookup)
template <class T, bool EXACT>
__device__ __noinline__ V3<T> g_hit_uv(const SceneViewG<T>& sc, const GPrim<T>& pr, uint32_t best_sub, V3<T> pi, V3<T> outward) {
    T uu = T(0), vv = T(0);
    T* u = &uu; T* v = &vv;
    if (pr.kind == P_SPHERE) g_sphere_uv<T, EXACT>(outward.x, outward.y, outward.z, u, v);               // of the outward normal, sphere.rs:83-84
    else if (pr.kind == P_PLANE) {                                         // get_plane_uv, plane.rs:41-55
        const auto& pl = GRec<T>::get(sc.plane_geo + pr.first);
        if (!pl.rotated) { *u = pi.x; *v = pi.z; }                         // the normal is +y
        else {                                                             // Rodrigues' rotation onto +y, then f64::fract of x and z
            V3<T> w = pi - pl.point;
            V3<T> rot = (w * pl.cos_theta + cross(pl.k, w) * pl.sin_theta) + (pl.k * dot(pl.k, w)) * (T(1) - pl.cos_theta);
            *u = rot.x - trunc(rot.x); *v = rot.z - trunc(rot.z);
        }
    } else {                                                               // get_quad_uv, quadrilateral.rs:58-63
        const auto& Q = GRec<T>::get(sc.quads + best_sub);
        V3<T> pq = pi - Q.q;
        *u = dot(cross(pq, Q.v), Q.w); *v = dot(cross(Q.u, pq), Q.w);
    }
    return mk<T>(uu, vv, T(0));                                            // by value: no address of a caller's local escapes
}

// HitRecord::new (hittable.rs:102-129) for the winner, in the entity's space, then p back to world space (transformations.rs:21-27)
template <class T, bool EXACT>
RTW_D void g_hit_record(const SceneViewG<T>& sc, const Ray<T>& r, const GPrim<T>* entry, uint32_t best_sub, T best_t, Hit<T>* h) {
    const auto& pr = GRec<T>::get(entry);
    Ray<T> rr = r;
    if (pr.xform >= 0) { const auto& X = GRec<T>::get(sc.xforms + pr.xform); rr = g_instance_ray<T>(X, r); }
    h->t = best_t;
    V3<T> p = pr.kind == P_SPHERE ? at(rr, best_t) : g_at(rr, best_t);
    V3<T> outward;
    if (pr.kind == P_SPHERE) {
        Vec4T<T> s = sc.spheres[pr.first];
        if constexpr (EXACT) outward = (p - mk<T>(s.x, s.y, s.z)) / s.w;
        else outward = (p - mk<T>(s.x, s.y, s.z)) * frcp(s.w);
    } else if (pr.kind == P_PLANE) outward = sc.plane_geo[pr.first].normal;
    else outward = g_quad_normal<T>(sc.quads + best_sub);
    h->front_face = dot(rr.d, outward) < T(0);
    h->normal = h->front_face ? outward : -outward;
    if (pr.xform >= 0) {
        const auto& X = GRec<T>::get(sc.xforms + pr.xform);
        p = g_mat_vec<T>(X.fwd, p) + mk<T>(X.ft[0], X.ft[1], X.ft[2]);
    }
    h->p = p;
    const auto& m = GRec<T>::get(sc.mats + pr.mat);
    h->gkind = m.kind;
    h->info = (pr.id << 2) | (m.kind & 3u);
    h->param = m.param;
    // Metal reads its own albedo; Lambertian / Isotropic / DiffuseLight read their texture at the hit point
    if (m.kind == LAMBERTIAN || m.kind >= DIFFUSE_LIGHT) {
        T u = T(0), v = T(0);
        if (m.texture && sc.textures[m.texture - 1].kind == TEX_CHECKER) {     // the hit's (u, v), in the entity's own space
            V3<T> uv = g_hit_uv<T, EXACT>(sc, pr, best_sub, g_at(rr, best_t), outward);
            u = uv.x; v = uv.y;
        }
        h->albedo = g_texture<T>(sc, m, u, v, p);
    } else h->albedo = mk<T>(m.albedo[0], m.albedo[1], m.albedo[2]);
}

template <class T, bool EXACT, bool COUNT>
RTW_D bool g_closest_hit(const SceneViewG<T>& sc, const Ray<T>& r, T tmin, T tmax, Hit<T>* h, int32_t* stack, int stride, Tally& tl) {
    const GPrim<T>* best; uint32_t sub; T t;
    if (!g_closest_entry<T, EXACT, COUNT>(sc, r, tmin, tmax, &best, &sub, &t, stack, stride, tl)) return false;
    g_hit_record<T, EXACT>(sc, r, best, sub, t, h);
    return true;
}

// ---- lights: HittableList::{pdf_value, random} (hittable_list.rs:408-420) over spheres, quads, triangles ---------
template <class T, bool EXACT, bool COUNT>
RTW_D T g_lights_pdf_value(const SceneViewG<T>& sc, V3<T> origin, V3<T> dir, Tally& tl) {
    using Mt = M<T, EXACT>;
    T acc = T(0);
    Ray<T> r{origin, dir};
    T a = sqlen(dir);
    for (int i = 0; i < sc.n_lights; ++i) {
        const auto& pr = GRec<T>::get(sc.lights + i);
        T v = T(0);
        if (pr.xform < 0) {                                        // Transformed<T>, Cuboid, Plane: Hittable default 0 (hittable.rs:175-177)
            if (pr.kind == P_SPHERE) {                             // sphere.rs:101-111
                if (COUNT) tl.light_tests++;
                Vec4T<T> s = sc.spheres[pr.first];
                T t;
                bool hit;
                if constexpr (EXACT) hit = sphere_root<T>(s, r, a, T(0), Mt::inf(), &t);
                else hit = sphere_root_fast(s, r, frcp(a), T(0), Mt::inf(), &t);
                if (hit) {
                    T distance_squared = sqlen(mk<T>(s.x - origin.x, s.y - origin.y, s.z - origin.z));
                    if constexpr (EXACT) {
                        T cos_theta_max = Mt::sqrt_(T(1) - s.w * s.w / distance_squared);
                        v = T(1) / (T(2) * Mt::PI * (T(1) - cos_theta_max));
                    } else {        // SFU reciprocals instead of IEEE divisions (~10 instructions each), like the sphere-only path
                        T cos_theta_max = Mt::sqrt_(T(1) - s.w * s.w * frcp(distance_squared));
                        v = frcp(T(2) * Mt::PI * (T(1) - cos_theta_max));
                    }
                }
            } else if (pr.kind == P_QUAD || pr.kind == P_TRIANGLE) {   // quadrilateral.rs:100-112
                if (COUNT) tl.light_tests++;
                const auto& Q = GRec<T>::get(sc.quads + pr.first);
                T t;
                if (g_quad_hit<T, EXACT>(Q, pr.kind == P_TRIANGLE, r, T(0), Mt::inf(), &t)) {
                    T distance_squared = t * t * a;
                    V3<T> n = dot(dir, Q.normal) < T(0) ? Q.normal : -Q.normal;
                    if constexpr (EXACT) {
                        T cosine = fabs(dot(dir, n) / Mt::sqrt_(a));
                        v = distance_squared / (cosine * Q.area);
                    } else {
                        T cosine = fabs(dot(dir, n) * frcp(Mt::sqrt_(a)));
                        v = distance_squared * frcp(cosine * Q.area);
                    }
                }
            }
        }
        acc = acc + v;
    }
    T len = (T)sc.n_lights;
    if (sc.lights_is_bvh) return ((acc / len) * len) / len;       // bvh.rs:67-76, 191-194 over one Leaf
    return acc / len;
}
template <class T, bool EXACT>
RTW_D V3<T> g_lights_random(const SceneViewG<T>& sc, V3<T> origin, Stream<EXACT>& rng) {
    const auto& pr = GRec<T>::get(sc.lights + uindex(rng, (uint32_t)sc.n_lights));
    if (pr.xform < 0) {
        if (pr.kind == P_SPHERE) return sphere_random<T, EXACT>(sc.spheres[pr.first], origin, rng);
        if (pr.kind == P_QUAD || pr.kind == P_TRIANGLE) {         // quadrilateral.rs:114-118, triangles.rs:108-117
            const auto& Q = GRec<T>::get(sc.quads + pr.first);
            T r1 = open01(rng), r2 = open01(rng);
            if (pr.kind == P_TRIANGLE && r1 + r2 > T(1)) { r1 = T(1) - r1; r2 = T(1) - r2; }
            return ((Q.q + Q.u * r1) + Q.v * r2) - origin;
        }
    }
    // An EMPTY lights list is uploaded as one entry of kind P_NO_LIGHTS (n_lights stays 0, so lights.pdf_value is 0 / 0 = NaN as in
    // the reference): drawing from it is the reference's panic (.expect("HittableList shouldn't be empty"), hittable_list.rs:414-419).
    // The flag makes the render call return RTW_E_INVALID; what this path computes afterwards is discarded with it.
    if (pr.kind == P_NO_LIGHTS) *sc.panic_flag = 1u;
    return mk<T>(1, 0, 0);                                        // Hittable::random default, hittable.rs:179-181
}

template <class T> RTW_D V3<T> g_emitted(const Hit<T>& h) {       // DiffuseLight::emitted (material.rs:506-514); others 0
    return h.gkind == DIFFUSE_LIGHT ? h.albedo : mk<T>(0, 0, 0);
}

// Material::scatter + the Scatter branch of ray_colour_tail_call (camera.rs:484-521)
template <class T, bool EXACT, bool COUNT>
RTW_D uint32_t g_shade(const SceneViewG<T>& sc, const Ray<T>& r, const Hit<T>& h, Stream<EXACT>& rng, Ray<T>* next, V3<T>* weight, Tally& tl) {
    using Mt = M<T, EXACT>;
    const uint32_t kind = h.gkind;
    if (kind == LAMBERTIAN || kind == ISOTROPIC) {
        const bool cosine = kind == LAMBERTIAN;
        if (COUNT && cosine) tl.lambertian++;
        Onb<T, EXACT> uvw(h.normal);
        V3<T> dir;
        if (standard(rng) < T(0.5)) dir = g_lights_random<T, EXACT>(sc, h.p, rng);      // MixturePdf::generate (pdf1 = lights)
        else if (cosine) {                                                              // CosineWeightedHemisphere, utils.rs:146-161
            T r1 = standard(rng), r2 = standard(rng);
            T sn, cs;
            Mt::sincos_2pi(r1, &sn, &cs);
            dir = uvw.transform(mk<T>(cs * Mt::sqrt_(r2), sn * Mt::sqrt_(r2), Mt::sqrt_(T(1) - r2)));
        } else {                                                                        // SpherePdf::generate = UnitSphere (pdf.rs:21-31)
            for (;;) {
                T a = T(2) * standard(rng) - T(1), b = T(2) * standard(rng) - T(1), c = T(2) * standard(rng) - T(1);
                dir = mk<T>(a, b, c);
                if (sqlen(dir) < T(1)) break;
            }
        }
        T light_v = g_lights_pdf_value<T, EXACT, COUNT>(sc, h.p, dir, tl);
        T own_v, scattering_pdf;
        if (cosine) {
            V3<T> nd = Mt::normalize(dir);
            own_v = Mt::max_(Mt::div_pi(dot(nd, uvw.w)), T(0));
            scattering_pdf = Mt::max_(Mt::div_pi(dot(h.normal, nd)), T(0));
        } else {
            own_v = T(1) / (T(4) * Mt::PI);
            scattering_pdf = T(1) / (T(4) * Mt::PI);
        }
        T pdf_value = light_v * T(0.5) + own_v * T(0.5);
        *next = Ray<T>{h.p, dir};
        if constexpr (EXACT) *weight = (h.albedo * scattering_pdf) / pdf_value;      // camera.rs:518
        else *weight = h.albedo * (scattering_pdf * frcp(pdf_value));
        return V_DIFFUSE;
    }
    if (kind == METAL || kind == DIELECTRIC) {
        // same code as the sphere-only scenes: shade() never touches the scene for these two kinds
        SceneView<T> none{};
        return shade<T, EXACT, COUNT, SceneView<T>, true>(none, r, h, rng, next, weight, tl);
    }
    if (COUNT) tl.absorbed++;                                      // DiffuseLight, Invisible: Material::scatter default None
    return V_ABSORB;
}

}  // namespace rtw
