// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product (see the header of rtw_oracle.hpp).
//
// General-scene extension of the CPU restatement: Quad, Triangle, Cuboid,
// Transformed<T>, DiffuseLight, Isotropic, NoiseTexture / Perlin, and lights lists that hold quads.
// Same conventions as rtw_oracle.hpp: C++17, f64, -ffp-contract=off, every function cites the reference
// file:line it follows.  PARITY UNPINNED, with one exception: the reference's own known-answer test for
// Transformation::inverse (geometry/src/transformations.rs:131-160) is repeated in tests/test_oracle_general.py.
//
// Documented choices where the reference leaves the behaviour open:
//   * HittableList buckets its objects by TypeId in TypeId ORDER (hittable_list.rs:270-294), which is
//     compiler-chosen.  The oracle fixes the order [Plane, Sphere, Quad, Triangle, Cuboid, then the
//     Transformed<...> of each].  It matters for exact-t ties, for the f64 summation order of
//     lights.pdf_value and for which list position a light index denotes.
//   * Perlin tables come from the unseeded thread_rng (perlin.rs:37-56); here from a Philox stream.
//   * NoiseTexture's f64::sin and Sphere uv's f64::acos are platform-libm calls whose bits are not pinned;
//     MathMode PORTABLE replaces them by fixed IEEE sequences (sin_portable, acos_msun).  Sphere uv's
//     atan2 is libm::atan2 (the `libm` crate, a port of musl / msun) in the reference and is restated
//     exactly (atan2_msun).  uv is evaluated only where a CheckerTexture reads it.
#pragma once
#include "rtw_oracle.hpp"

namespace orcg {
using namespace orc;

// ------------------------------------------------------------------------------------------------
// geometry/src/matrix3.rs, geometry/src/transformations.rs:30-128 (the default, non-euclid build)
struct Mat3 {
    double m[3][3] = {{1., 0., 0.}, {0., 1., 0.}, {0., 0., 1.}};                         // matrix3.rs:47-52
    V3 row(int i) const { return {m[i][0], m[i][1], m[i][2]}; }
    V3 col(int j) const { return {m[0][j], m[1][j], m[2][j]}; }
    double det() const {                                                                  // matrix3.rs:35-38
        const double a = m[0][0], b = m[0][1], c = m[0][2], d = m[1][0], e = m[1][1], f = m[1][2], g = m[2][0], h = m[2][1], i = m[2][2];
        return a * (e * i - f * h) + b * (f * g - d * i) + c * (d * h - e * g);
    }
    bool inverse(Mat3* out) const {                                                       // matrix3.rs:12-30
        double dt = det();
        if (!std::isnormal(dt)) return false;
        const double a = m[0][0], b = m[0][1], c = m[0][2], d = m[1][0], e = m[1][1], f = m[1][2], g = m[2][0], h = m[2][1], i = m[2][2];
        const double A = e * i - f * h, B = f * g - d * i, C = d * h - e * g;
        const double D = c * h - b * i, E = a * i - c * g, F = b * g - a * h;
        const double G = b * f - c * e, H = c * d - a * f, I = a * e - b * d;
        double r[3][3] = {{A / dt, D / dt, G / dt}, {B / dt, E / dt, H / dt}, {C / dt, F / dt, I / dt}};
        std::memcpy(out->m, r, sizeof(r));
        return true;
    }
};
inline V3 operator*(const Mat3& a, V3 v) { return {dot(a.row(0), v), dot(a.row(1), v), dot(a.row(2), v)}; }   // matrix3.rs:88-100
inline Mat3 operator*(const Mat3& a, const Mat3& b) {                                                        // matrix3.rs:67-86
    Mat3 o;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) o.m[i][j] = dot(a.row(i), b.col(j));
    return o;
}
struct Transformation {
    Mat3 rotation; V3 translation;
    Transformation apply(const Transformation& t) const {                                 // transformations.rs:104-111
        return {t.rotation * rotation, t.translation + t.rotation * translation};
    }
    Transformation then(const Transformation& t) const { return apply(t); }              // :113-116
    V3 transform_point3d(V3 p) const { return rotation * p + translation; }              // :118-121
    V3 transform_vector3d(V3 v) const { return rotation * v + translation; }             // :123-126 (adds the translation, as the reference does)
    bool inverse(Transformation* out) const {                                             // :128-136
        Mat3 r;
        if (!rotation.inverse(&r)) return false;
        out->rotation = r; out->translation = -(r * translation);
        return true;
    }
};
enum Axis { AX_X = 0, AX_Y = 1, AX_Z = 2 };
inline Transformation rotation(double angle_deg, int axis, uint32_t math) {               // transformations.rs:38-64
    double angle = angle_deg * (PI / 180.);
    double s, c;
    if (math == LIBM) { s = std::sin(angle); c = std::cos(angle); } else sincos_phi(angle, PORTABLE, &s, &c);
    Transformation t;
    if (axis == AX_X) { double r[3][3] = {{1., 0., 0.}, {0., c, -s}, {0., s, c}}; std::memcpy(t.rotation.m, r, sizeof(r)); }
    else if (axis == AX_Y) { double r[3][3] = {{c, 0., s}, {0., 1., 0.}, {-s, 0., c}}; std::memcpy(t.rotation.m, r, sizeof(r)); }
    else { double r[3][3] = {{c, -s, 0.}, {s, c, 0.}, {0., 0., 1.}}; std::memcpy(t.rotation.m, r, sizeof(r)); }
    return t;
}

// AABBox::from_points (aabox.rs:191-204): fold of enclose(point), each of which pads (aabox.rs:161-175)
inline AABB box_from_points(const V3* p, int n) {
    AABB b{p[0], p[0]};
    for (int i = 1; i < n; ++i) b = b.enclose(AABB{p[i], p[i]});
    return b;
}

// ------------------------------------------------------------------------------------------------
// sin for arbitrary moderate arguments (|x| < 2^20 * pi/2): round-to-nearest quadrant, two-constant
// Cody-Waite reduction, fdlibm kernels — the sequence sincos_phi uses, extended to negative / large x.
inline double sin_portable(double x) {
    const double two_over_pi = 6.36619772367581382433e-01;
    const double pio2_1 = 1.57079632673412561417e+00, pio2_1t = 6.07710050650619224932e-11;
    double fn = std::floor(x * two_over_pi + 0.5);
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

// atan2 as the `libm` crate evaluates it (Sphere::get_sphere_uv calls libm::atan2, sphere.rs:49-54; libm 0.2 is a port of
// musl's atan2.c / atan.c, i.e. FreeBSD msun's e_atan2.c / s_atan.c): a fixed sequence of IEEE operations.
inline uint32_t hi_word(double x) { uint64_t b; std::memcpy(&b, &x, 8); return (uint32_t)(b >> 32); }
inline uint32_t lo_word(double x) { uint64_t b; std::memcpy(&b, &x, 8); return (uint32_t)b; }
inline double atan_msun(double x) {
    static const double atanhi[4] = {4.63647609000806093515e-01, 7.85398163397448278999e-01, 9.82793723247329054082e-01, 1.57079632679489655800e+00};
    static const double atanlo[4] = {2.26987774529616870924e-17, 3.06161699786838301793e-17, 1.39033110312309984516e-17, 6.12323399573676603587e-17};
    static const double aT[11] = {3.33333333333329318027e-01, -1.99999999998764832476e-01, 1.42857142725034663711e-01, -1.11111104054623557880e-01,
                                  9.09088713343650656196e-02, -7.69187620504482999495e-02, 6.66107313738753120669e-02, -5.83357013379057348645e-02,
                                  4.97687799461593236017e-02, -3.65315727442169155270e-02, 1.62858201153657823623e-02};
    uint32_t ix = hi_word(x);
    const bool sign = (ix >> 31) != 0;
    ix &= 0x7fffffffu;
    int id;
    if (ix >= 0x44100000u) {                       // |x| >= 2^66
        if (x != x) return x;
        double z = atanhi[3] + 0x1p-120;
        return sign ? -z : z;
    }
    if (ix < 0x3fdc0000u) {                        // |x| < 0.4375
        if (ix < 0x3e400000u) return x;            // |x| < 2^-27
        id = -1;
    } else {
        x = std::fabs(x);
        if (ix < 0x3ff30000u) {                    // |x| < 1.1875
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
inline double atan2_msun(double y, double x) {
    const double pi = 3.1415926535897931160E+00, pi_lo = 1.2246467991473531772E-16;
    if (x != x || y != y) return x + y;
    uint32_t ix = hi_word(x), lx = lo_word(x), iy = hi_word(y), ly = lo_word(y);
    if (((ix - 0x3ff00000u) | lx) == 0) return atan_msun(y);       // x = 1.0
    uint32_t m = ((iy >> 31) & 1u) | ((ix >> 30) & 2u);            // 2 * sign(x) + sign(y)
    ix &= 0x7fffffffu; iy &= 0x7fffffffu;
    if ((iy | ly) == 0) { switch (m) { case 0: case 1: return y; case 2: return pi; default: return -pi; } }
    if ((ix | lx) == 0) return (m & 1u) ? -pi / 2 : pi / 2;
    if (ix == 0x7ff00000u) {
        if (iy == 0x7ff00000u) { switch (m) { case 0: return pi / 4; case 1: return -pi / 4; case 2: return 3 * pi / 4; default: return -3 * pi / 4; } }
        switch (m) { case 0: return 0.0; case 1: return -0.0; case 2: return pi; default: return -pi; }
    }
    if (ix + (64u << 20) < iy || iy == 0x7ff00000u) return (m & 1u) ? -pi / 2 : pi / 2;      // |y / x| > 2^64
    double z = ((m & 2u) && iy + (64u << 20) < ix) ? 0.0 : atan_msun(std::fabs(y / x));
    switch (m) { case 0: return z; case 1: return -z; case 2: return pi - (z - pi_lo); default: return (z - pi_lo) - pi; }
}
// f64::acos is the platform libm's acos in the reference; PORTABLE replaces it by msun's e_acos.c sequence
inline double acos_msun(double x) {
    const double pio2_hi = 1.57079632679489655800e+00, pio2_lo = 6.12323399573676603587e-17;
    const double pS0 = 1.66666666666666657415e-01, pS1 = -3.25565818622400915405e-01, pS2 = 2.01212532134862925881e-01,
                 pS3 = -4.00555345006794114027e-02, pS4 = 7.91534994289814532176e-04, pS5 = 3.47933107596021167570e-05,
                 qS1 = -2.40339491173441421878e+00, qS2 = 2.02094576023350569471e+00, qS3 = -6.88283971605453293030e-01, qS4 = 7.70381505559019352791e-02;
    auto R = [&](double z) {
        double p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        double q = 1.0 + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        return p / q;
    };
    uint32_t hx = hi_word(x), ix = hx & 0x7fffffffu;
    if (ix >= 0x3ff00000u) {                       // |x| >= 1 or NaN
        if (((ix - 0x3ff00000u) | lo_word(x)) == 0) return (hx >> 31) ? 2 * pio2_hi + 0x1p-120 : 0.0;
        return 0.0 / (x - x);
    }
    if (ix < 0x3fe00000u) {                        // |x| < 0.5
        if (ix <= 0x3c600000u) return pio2_hi + 0x1p-120;
        return pio2_hi - (x - (pio2_lo - x * R(x * x)));
    }
    if (hx >> 31) {                                // x < -0.5
        double z = (1.0 + x) * 0.5, sq = std::sqrt(z), w = R(z) * sq - pio2_lo;
        return 2 * (pio2_hi - (sq + w));
    }
    double z = (1.0 - x) * 0.5, sq = std::sqrt(z);
    uint64_t b; std::memcpy(&b, &sq, 8); b &= 0xffffffff00000000ull;
    double df; std::memcpy(&df, &b, 8);
    double c = (z - df * df) / (sq + df), w = R(z) * sq + c;
    return 2 * (df + w);
}

// shared/src/perlin.rs
struct Perlin {
    double rand_vec[256][3];
    uint8_t perm_x[256], perm_y[256], perm_z[256];
    // Perlin::new (perlin.rs:46-57), seeded: stream (pixel 0x9E71A000 + index, sample 0, vertex 0), W64.
    // rand_vec: 256 x UnitSphere (rejection triples, shuffle dropped); then perm_x, perm_y, perm_z, each a
    // Fisher-Yates pass j = Uniform::new(i, 256) (perlin.rs:37-44) drawn as i + index(256 - i).
    static Perlin generate(uint64_t seed, uint32_t index) {
        Perlin p;
        Stream rng(seed, 0x9E71A000u + index, 0, 0, W64);
        for (int i = 0; i < 256; ++i) {
            for (;;) {
                double a = 2. * rng.standard() - 1., b = 2. * rng.standard() - 1., c = 2. * rng.standard() - 1.;
                if (a * a + b * b + c * c < 1.) { p.rand_vec[i][0] = a; p.rand_vec[i][1] = b; p.rand_vec[i][2] = c; break; }
            }
        }
        uint8_t* perms[3] = {p.perm_x, p.perm_y, p.perm_z};
        for (uint8_t* pm : perms) {
            for (int i = 0; i < 256; ++i) pm[i] = (uint8_t)i;
            for (uint32_t i = 0; i < 255; ++i) { uint32_t j = i + rng.index(256 - i); std::swap(pm[i], pm[j]); }
        }
        return p;
    }
    static int wrap(double f) { double r = std::fmod(f, 256.); if (r < 0.) r += 256.; return (int)r; }   // f64::rem_euclid(256) as usize
    double noise(V3 p) const {                                                            // perlin.rs:59-83, 97-110
        double fx = std::floor(p.x), fy = std::floor(p.y), fz = std::floor(p.z);
        double u = p.x - fx, v = p.y - fy, w = p.z - fz;
        double acc = 0.;                                                                  // .sum() (a -0.0 start would only change an all -0.0 sum)
        for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) for (int k = 0; k < 2; ++k) {
            const double* c = rand_vec[perm_x[wrap(fx + (double)i)] ^ perm_y[wrap(fy + (double)j)] ^ perm_z[wrap(fz + (double)k)]];
            double di = (double)i, dj = (double)j, dk = (double)k;
            V3 weight_v{u - di, v - dj, w - dk};
            double term = (di * u + (1. - di) * (1. - u)) * (dj * v + (1. - dj) * (1. - v)) * (dk * w + (1. - dk) * (1. - w)) *
                          dot(V3{c[0], c[1], c[2]}, weight_v);
            acc = acc + term;
        }
        return acc;
    }
    double turb(V3 p, int depth) const {                                                  // perlin.rs:85-95
        double accum = 0.; V3 temp_p = p; double weight = 1.;
        for (int i = 0; i < depth; ++i) { accum += weight * noise(temp_p); temp_p = temp_p * 2.; weight *= 0.5; }
        return accum;
    }
};

enum TexKind : uint32_t { TEX_SOLID = 0, TEX_NOISE = 1, TEX_CHECKER = 2 };
// NoiseTexture{noise: perlins[perlin], scale} or CheckerTexture{inv_scale = 1 / scale, even, odd}: even / odd are texture
// references (0 = SolidColour(even_colour / odd_colour), k = textures[k-1]); texture.rs:24-55
struct Texture { uint32_t kind = TEX_SOLID; uint32_t perlin = 0; double scale = 1.; uint32_t even = 0, odd = 0; V3 even_colour, odd_colour; };
enum GMatKind : uint32_t { DIFFUSE_LIGHT = 4, ISOTROPIC = 5 };
struct GMaterial { uint32_t kind = LAMBERTIAN; uint32_t texture = 0; V3 albedo; double param = 0.; };   // texture: 0 = SolidColour(albedo), k = textures[k-1]

// ------------------------------------------------------------------------------------------------
struct GHit {
    V3 p, normal;
    V3 sphere_outward;        // spheres: the outward normal get_sphere_uv is evaluated on (sphere.rs:83-84); uv is computed on demand
    bool sphere_uv = false;
    double t = 0., u = 0., v = 0.;
    bool front_face = false;
    int32_t prim = -1;
    uint32_t mat = 0;
};
inline GHit make_ghit(const Ray& r, double t, V3 outward, double u, double v, uint32_t mat) {   // hittable.rs:102-129
    GHit h;
    h.sphere_uv = false;
    h.p = r.at(t);
    h.front_face = dot(r.d, outward) < 0.;
    h.normal = h.front_face ? outward : -outward;
    h.t = t; h.u = u; h.v = v; h.mat = mat;
    return h;
}

// shared/src/entities/quadrilateral.rs:23-118 and triangles.rs:23-120 (same structure; `tri` marks the differences)
struct Quad {
    V3 q, u, v, w, normal;
    double area = 0.;
    AABB box;
    uint32_t mat = 0;
    bool tri = false;
    static Quad make(V3 q, V3 u, V3 v, uint32_t mat, bool tri) {
        Quad s;
        s.q = q; s.u = u; s.v = v; s.mat = mat; s.tri = tri;
        V3 pts[5] = {q + (u + v) * 0.5, q, q + v, q + u, (q + u) + v};
        s.box = box_from_points(pts, tri ? 4 : 5);                                        // quadrilateral.rs:43 / triangles.rs:41
        V3 n = cross(u, v);
        s.w = n / square_length(n);
        s.area = tri ? length(n) / 2. : length(n);                                        // triangles.rs:44: the triangle's unit... normal = n / (|n|/2), length 2
        s.normal = n / s.area;
        return s;
    }
    bool hit(const Ray& r, double start, double end, GHit* out) const {                   // quadrilateral.rs:79-98 / triangles.rs:74-92
        double denom = dot(r.d, normal);
        if (!(std::fabs(denom) > EPS)) return false;
        double t = -(dot(r.o - q, normal) / denom);
        if (!(start <= t && t <= end)) return false;
        V3 point = r.at(t);
        double a = dot(cross(point - q, v), w);                                           // get_quad_uv, quadrilateral.rs:58-63
        double b = dot(cross(u, point - q), w);
        bool inside = tri ? (0. <= a + b && a + b <= 1.) : (0. <= a && a <= 1. && 0. <= b && b <= 1.);
        if (!inside) return false;
        if (out) *out = make_ghit(r, t, normal, a, b, mat);
        return true;
    }
    double pdf_value(V3 origin, V3 direction) const {                                     // quadrilateral.rs:100-112
        GHit rec;
        if (!hit(Ray{origin, direction}, 0., INF, &rec)) return 0.;
        double distance_squared = rec.t * rec.t * square_length(direction);
        double cosine = std::fabs(dot(direction, rec.normal) / length(direction));
        return distance_squared / (cosine * area);
    }
    V3 random(V3 origin, Stream& rng) const {                                             // quadrilateral.rs:114-118 / triangles.rs:108-117
        double r1 = rng.open01(), r2 = rng.open01();
        if (tri && r1 + r2 > 1.) { r1 = 1. - r1; r2 = 1. - r2; }
        V3 p = (q + u * r1) + v * r2;
        return p - origin;
    }
};

// shared/src/entities/cuboid.rs
struct Cuboid {
    Quad quads[6];
    static Cuboid make(V3 p, V3 q, uint32_t mat) {                                        // cuboid.rs:26-50
        V3 pts[2] = {p, q};
        AABB b = box_from_points(pts, 2);
        V3 min_p = b.mn, max_p = b.mx, delta = max_p - min_p;
        V3 dx{delta.x, 0., 0.}, dy{0., delta.y, 0.}, dz{0., 0., delta.z};
        Cuboid c;
        c.quads[0] = Quad::make(min_p, dx, dy, mat, false);
        c.quads[1] = Quad::make(min_p, dy, dz, mat, false);
        c.quads[2] = Quad::make(min_p, dx, dz, mat, false);
        c.quads[3] = Quad::make(max_p, -dx, -dy, mat, false);
        c.quads[4] = Quad::make(max_p, -dy, -dz, mat, false);
        c.quads[5] = Quad::make(max_p, -dx, -dz, mat, false);
        return c;
    }
    bool hit(const Ray& r, double start, double end, GHit* out) const {                   // cuboid.rs:53-60: min_by total_cmp, first minimum
        bool any = false; GHit best, rec;
        for (const Quad& q : quads)
            if (q.hit(r, start, end, &rec) && (!any || total_cmp(rec.t, best.t) < 0)) { best = rec; any = true; }
        if (any && out) *out = best;
        return any;
    }
    AABB get_aabbox() const {                                                             // cuboid.rs:62-73
        AABB b = quads[0].box;
        for (int i = 1; i < 6; ++i) b = b.enclose(quads[i].box);
        return b;
    }
};

enum PrimKind : uint32_t { P_SPHERE = 0, P_PLANE = 1, P_QUAD = 2, P_TRIANGLE = 3, P_CUBOID = 4 };

// One world / lights entry: an entity, optionally wrapped in Transformed<T>
// (geometry/src/transformations.rs:168-239, shared/src/entities/transformations.rs:10-30).
struct Prim {
    uint32_t kind = P_SPHERE;
    Sphere sphere; Plane plane; Quad quad; Cuboid cuboid;
    bool transformed = false;
    Transformation tf, inv; bool has_inv = false;
    AABB box;
    int32_t id = -1;
    uint32_t bucket() const { return (kind == P_PLANE ? 0u : kind == P_SPHERE ? 1u : kind) + (transformed ? 5u : 0u); }

    AABB instance_box() const {
        switch (kind) {
            case P_SPHERE: return sphere.box;
            case P_PLANE: return plane.get_aabbox();
            case P_CUBOID: return cuboid.get_aabbox();
            default: return quad.box;
        }
    }
    void finalize() {
        AABB ib = instance_box();
        if (!transformed) { box = ib; return; }
        has_inv = tf.inverse(&inv);
        V3 pts[8] = {{ib.mn.x, ib.mn.y, ib.mn.z}, {ib.mn.x, ib.mx.y, ib.mn.z}, {ib.mn.x, ib.mn.y, ib.mx.z}, {ib.mn.x, ib.mx.y, ib.mx.z},
                     {ib.mx.x, ib.mn.y, ib.mn.z}, {ib.mx.x, ib.mx.y, ib.mn.z}, {ib.mx.x, ib.mn.y, ib.mx.z}, {ib.mx.x, ib.mx.y, ib.mx.z}};   // aabox.rs:114-126
        for (V3& p : pts) p = tf.transform_point3d(p);
        box = box_from_points(pts, 8);                                                    // transformations.rs:224-233
    }
    bool instance_hit(const Ray& r, double start, double end, GHit* out, bool* panicked) const {
        switch (kind) {
            case P_SPHERE: {
                HitRecord h;
                if (!sphere.hit(r, start, end, &h)) return false;
                out->p = h.p; out->normal = h.normal; out->t = h.t; out->u = 0.; out->v = 0.; out->front_face = h.front_face; out->mat = h.mat;
                out->sphere_outward = (h.p - sphere.center) / sphere.radius; out->sphere_uv = true;
                return true;
            }
            case P_PLANE: {
                HitRecord h;
                if (!plane.hit(r, start, end, &h, panicked)) return false;
                out->p = h.p; out->normal = h.normal; out->t = h.t; out->front_face = h.front_face; out->mat = h.mat; out->sphere_uv = false;
                // get_plane_uv (plane.rs:41-55): theta = angle between the normal and +y; (x, z) of the point when it is 0
                {
                    V3 V{0., 1., 0.};
                    double theta = std::atan2(length(cross(plane.normal, V)), dot(plane.normal, V));
                    if (theta <= EPS) { out->u = h.p.x; out->v = h.p.z; }
                    else {
                        V3 k = normalize(cross(plane.normal, V));
                        V3 w = h.p - plane.point;
                        V3 rot = (w * std::cos(theta) + cross(k, w) * std::sin(theta)) + (k * dot(k, w)) * (1. - std::cos(theta));
                        double ip;
                        out->u = std::modf(rot.x, &ip); out->v = std::modf(rot.z, &ip);
                    }
                }
                return true;
            }
            case P_CUBOID: return cuboid.hit(r, start, end, out);
            default: return quad.hit(r, start, end, out);
        }
    }
    bool hit(const Ray& r, double start, double end, GHit* out, bool* panicked) const {
        bool ok;
        if (!transformed) ok = instance_hit(r, start, end, out, panicked);
        else {                                                                            // entities/transformations.rs:14-29
            if (!has_inv) return false;
            Ray rr{inv.transform_point3d(r.o), inv.transform_vector3d(r.d)};
            ok = instance_hit(rr, start, end, out, panicked);
            if (ok) out->p = tf.transform_point3d(out->p);
        }
        if (ok) out->prim = id;
        return ok;
    }
    double pdf_value(V3 origin, V3 direction, Counters* c) const {                        // Hittable default 0 (hittable.rs:175-177) unless overridden
        if (transformed) return 0.;
        if (kind == P_SPHERE) return sphere.pdf_value(origin, direction, c);
        if (kind == P_QUAD || kind == P_TRIANGLE) { if (c) c->light_tests++; return quad.pdf_value(origin, direction); }
        return 0.;
    }
};

// ------------------------------------------------------------------------------------------------
// HittableList / BoundedVolumeHierarchy over Prim, same structure as rtw_oracle.hpp's (which is specialised to
// planes + spheres); see the citations there.
constexpr int kBuckets = 10;
struct GList {
    std::vector<Prim> bucket[kBuckets]; std::optional<AABB> bucket_box[kBuckets];
    size_t len = 0;
    std::optional<AABB> aabbox;
    void add(const Prim& o) {
        aabbox = aabbox ? aabbox->enclose(o.box) : o.box;
        uint32_t b = o.bucket();
        bucket[b].push_back(o);
        bucket_box[b] = bucket_box[b] ? bucket_box[b]->enclose(o.box) : o.box;
        len++;
    }
    AABB get_aabbox() const { return aabbox ? *aabbox : AABB{}; }
    template <class F> void for_each(F f) const { for (int b = 0; b < kBuckets; ++b) for (const Prim& o : bucket[b]) f(o); }   // iter_hittable
    const Prim& nth(size_t i) const { for (int b = 0; b < kBuckets; ++b) { if (i < bucket[b].size()) return bucket[b][i]; i -= bucket[b].size(); } return bucket[0][0]; }
    bool hit(const Ray& r, double start, double end, GHit* best, Counters* c, bool* panicked) const {
        bool any = false; GHit rec;
        for (int b = 0; b < kBuckets; ++b) {
            if (bucket[b].empty() || !aabb_is_hit(*bucket_box[b], r, start, end, c)) continue;
            bool bany = false; GHit brec;
            for (const Prim& o : bucket[b]) {
                if (!aabb_is_hit(o.box, r, start, end, c)) continue;                      // bounded_hit, hittable.rs:191-196
                if (o.hit(r, start, end, &rec, panicked) && (!bany || rec.t < brec.t)) { brec = rec; bany = true; }
            }
            if (bany && (!any || brec.t < best->t)) { *best = brec; any = true; }
        }
        return any;
    }
    std::pair<GList, GList> split_by(int axis, double coord) const {
        GList left, right;
        auto push_box = [](GList& dst, const AABB& b) { dst.aabbox = dst.aabbox ? dst.aabbox->enclose(b) : b; };
        for (int b = 0; b < kBuckets; ++b) {
            GList l, r;
            for (size_t i = bucket[b].size(); i-- > 0;) (bucket[b][i].box.right_of(axis, coord) ? r : l).add(bucket[b][i]);
            if (!r.bucket[b].empty()) { right.len += r.bucket[b].size(); right.bucket[b] = r.bucket[b]; right.bucket_box[b] = r.bucket_box[b]; push_box(right, *r.bucket_box[b]); }
            if (!l.bucket[b].empty()) { left.len += l.bucket[b].size(); left.bucket[b] = l.bucket[b]; left.bucket_box[b] = l.bucket_box[b]; push_box(left, *l.bucket_box[b]); }
        }
        return {right, left};
    }
    void best_split_plane(int* axis_out, double* coord_out) const {
        size_t best0 = std::numeric_limits<size_t>::max();
        double best1 = INF; int best_axis = 0; double best_coord = 0.;
        std::vector<std::pair<double, double>> tmp;
        for (int axis = 0; axis < 3; ++axis) {
            tmp.clear();
            for_each([&](const Prim& o) { tmp.push_back({o.box.lo(axis), o.box.hi(axis)}); });
            std::stable_sort(tmp.begin(), tmp.end(), [](const auto& a, const auto& b) {
                int c = total_cmp(a.first, b.first);
                if (c == 0) c = total_cmp(a.second, b.second);
                return c < 0;
            });
            double median = tmp[tmp.size() / 2].first;
            size_t pp = 0;
            while (pp < tmp.size() && total_cmp(tmp[pp].first, median) < 0) pp++;
            double size = tmp.back().second - tmp.front().first;
            size_t cand0 = tmp.size() - 2 * pp;
            bool better = best0 > cand0 || (best0 == cand0 && (-best1) > (-size));
            if (better) { best0 = cand0; best1 = size; best_axis = axis; best_coord = median; }
        }
        *axis_out = best_axis; *coord_out = best_coord;
    }
    double pdf_value(V3 origin, V3 direction, Counters* c) const {                        // hittable_list.rs:408-412
        double acc = 0.;
        for_each([&](const Prim& o) { acc = acc + o.pdf_value(origin, direction, c); });
        return acc / (double)len;
    }
};
struct GBvh {
    bool leaf = true;
    GList list;
    std::unique_ptr<GBvh> left, right;
    size_t len = 0;
    AABB cached;
    static std::unique_ptr<GBvh> from(const GList& value) {
        auto n = std::make_unique<GBvh>();
        if (value.len <= 5) { n->list = value; n->len = value.len; n->cached = value.get_aabbox(); return n; }
        int axis; double coord;
        value.best_split_plane(&axis, &coord);
        auto pr = value.split_by(axis, coord);
        GList& l = pr.first; GList& r = pr.second;
        if (value.len == l.len) { n->list = l; n->len = l.len; n->cached = l.get_aabbox(); return n; }
        if (value.len == r.len) { n->list = r; n->len = r.len; n->cached = r.get_aabbox(); return n; }
        n->leaf = false;
        n->left = from(l); n->right = from(r);
        n->len = n->left->len + n->right->len;
        n->cached = n->left->cached.enclose(n->right->cached);
        return n;
    }
    bool hit(const Ray& r, double start, double end, GHit* out, Counters* c, bool* panicked) const {
        if (c) c->node_visits++;
        if (leaf) return list.hit(r, start, end, out, c, panicked);
        GHit a, b;
        bool ha = aabb_is_hit(left->cached, r, start, end, c) && left->hit(r, start, end, &a, c, panicked);
        bool hb = aabb_is_hit(right->cached, r, start, end, c) && right->hit(r, start, end, &b, c, panicked);
        if (!ha && !hb) return false;
        if (ha && hb) *out = (b.t < a.t) ? b : a;
        else *out = ha ? a : b;
        return true;
    }
    double aux_pdf_value(V3 o, V3 d, Counters* c) const {                                 // bvh.rs:67-76
        if (leaf) return list.pdf_value(o, d, c) * (double)list.len;
        return left->aux_pdf_value(o, d, c) + right->aux_pdf_value(o, d, c);
    }
};

struct GScene {
    std::vector<GMaterial> materials;
    std::vector<Texture> textures;
    std::vector<Perlin> perlins;
    GList world_list; std::unique_ptr<GBvh> world_bvh; bool world_is_bvh = false;
    GList lights; bool lights_is_bvh = false;      // a lights BVH is supported up to 5 lights (one Leaf); beyond that the
                                                   // reference's aux_random (bvh.rs:78-93) indexes out of range
    bool world_hit(const Ray& r, double tmin, GHit* rec, Counters* c, bool* panicked) const {
        if (c) c->rays++;
        return world_is_bvh ? world_bvh->hit(r, tmin, INF, rec, c, panicked) : world_list.hit(r, tmin, INF, rec, c, panicked);
    }
    double lights_pdf_value(V3 origin, V3 direction, Counters* c) const {
        if (!lights_is_bvh) return lights.pdf_value(origin, direction, c);
        return (lights.pdf_value(origin, direction, c) * (double)lights.len) / (double)lights.len;      // bvh.rs:191-194 over one Leaf
    }
    V3 lights_random(V3 origin, Stream& rng, uint32_t math, bool* panicked) const {       // hittable_list.rs:414-420 / bvh.rs:197-201
        if (lights.len == 0) {                                                            // .expect("HittableList shouldn't be empty")
            if (panicked) *panicked = true;
            return V3{1., 0., 0.};
        }
        const Prim& l = lights.nth(rng.index((uint32_t)lights.len));
        if (!l.transformed) {
            if (l.kind == P_SPHERE) return sphere_random(l.sphere, origin, rng, math);
            if (l.kind == P_QUAD || l.kind == P_TRIANGLE) return l.quad.random(origin, rng);
        }
        return V3{1., 0., 0.};                                                            // Hittable::random default, hittable.rs:179-181
    }
    V3 noise_colour(const Texture& t, V3 point, uint32_t math) const {                    // NoiseTexture::get_colour, texture.rs:90-102
        const Perlin& pn = perlins[t.perlin];
        double arg = t.scale * point.z + pn.turb(point, 7) * 10.;
        double s = math == LIBM ? std::sin(arg) : sin_portable(arg);
        return V3{0.5, 0.5, 0.5} * (s + 1.);
    }
    V3 texture_colour(const GMaterial& m, const GHit& rec, uint32_t math) const {
        if (m.texture == 0) return m.albedo;                                              // SolidColour, texture.rs:15-22
        const Texture* tp = &textures[m.texture - 1];
        if (tp->kind == TEX_NOISE) return noise_colour(*tp, rec.p, math);
        double u = rec.u, v = rec.v;                                                      // CheckerTexture::get_colour, texture.rs:46-55
        if (rec.sphere_uv) {                                                              // Sphere::get_sphere_uv, sphere.rs:49-54
            const double TAU = 6.28318530717958647692528676655900577;
            V3 n = rec.sphere_outward;
            u = atan2_msun(-n.z, n.x) / TAU;
            v = (math == LIBM ? std::acos(n.y) : acos_msun(n.y)) / PI;
        }
        // even / odd are textures themselves (Arc<dyn Texture>, texture.rs:26-29): get_colour recurses with the same (u, v, point)
        for (;;) {
            const Texture& t = *tp;
            double inv_scale = 1. / t.scale;
            bool is_even = std::fmod(std::floor(u * inv_scale) + std::floor(v * inv_scale), 2.) == 0.;
            uint32_t ref = is_even ? t.even : t.odd;
            if (ref == 0) return is_even ? t.even_colour : t.odd_colour;
            tp = &textures[ref - 1];
            if (tp->kind == TEX_NOISE) return noise_colour(*tp, rec.p, math);
        }
    }
};

struct GVertex { uint32_t kind = V_MISS; GHit rec; Ray next; V3 weight; V3 emitted; };

// Material::scatter / emitted + the Scatter branch of ray_colour_tail_call (camera.rs:478-521)
inline void gshade(const GScene& sc, const Options& opt, const Ray& r, const GHit& rec, Stream& rng, GVertex* vx, Counters* c, bool* panicked = nullptr) {
    const GMaterial& m = sc.materials[rec.mat];
    vx->rec = rec;
    vx->emitted = V3{0., 0., 0.};
    auto diffuse = [&](V3 attenuation, bool cosine) {
        Onb uvw(rec.normal);
        V3 dir;
        if (rng.standard() < 0.5) {                                                       // MixturePdf::generate, pdf.rs:94-100
            bool pan = false;
            dir = sc.lights_random(rec.p, rng, opt.math_mode, &pan);
            if (pan) {                        // the reference panics here; the oracle reports it and ends the path
                if (panicked) *panicked = true;
                vx->kind = V_ABSORB; vx->weight = V3{0., 0., 0.};
                return;
            }
        }
        else if (cosine) {                                                                // CosinePdf, utils.rs:146-161
            double r1 = rng.standard(), r2 = rng.standard();
            double sn, cs; sincos_phi(2. * PI * r1, opt.math_mode, &sn, &cs);
            dir = uvw.transform(V3{cs * std::sqrt(r2), sn * std::sqrt(r2), std::sqrt(1. - r2)});
        } else {                                                                          // SpherePdf::generate = UnitSphere, pdf.rs:21-31
            for (;;) {
                double a = 2. * rng.standard() - 1., b = 2. * rng.standard() - 1., d = 2. * rng.standard() - 1.;
                dir = V3{a, b, d};
                if (square_length(dir) < 1.) break;
            }
        }
        double light_v = sc.lights_pdf_value(rec.p, dir, c);
        double own_v = cosine ? rmax(dot(normalize(dir), uvw.w) / PI, 0.) : 1. / (4. * PI);
        double pdf_value = light_v * 0.5 + own_v * 0.5;
        double scattering_pdf = cosine ? rmax(dot(rec.normal, normalize(dir)) / PI, 0.) : 1. / (4. * PI);
        vx->kind = V_DIFFUSE; vx->next = Ray{rec.p, dir};
        vx->weight = (attenuation * scattering_pdf) / pdf_value;
    };
    switch (m.kind) {
    case LAMBERTIAN: if (c) c->lambertian++; diffuse(sc.texture_colour(m, rec, opt.math_mode), true); return;     // material.rs:357-376
    case ISOTROPIC: diffuse(sc.texture_colour(m, rec, opt.math_mode), false); return;                               // material.rs:529-554
    case DIFFUSE_LIGHT:                                                                                               // material.rs:506-514
        vx->emitted = sc.texture_colour(m, rec, opt.math_mode);
        vx->kind = V_ABSORB; vx->weight = V3{0., 0., 0.};
        return;
    case METAL: case DIELECTRIC: {
        // identical to rtw_oracle.hpp's shade(); restated on GHit
        HitRecord h; h.p = rec.p; h.normal = rec.normal; h.t = rec.t; h.front_face = rec.front_face; h.mat = 0;
        Scene tmp; tmp.materials.push_back(Material{m.kind, m.albedo, m.param});
        Vertex v;
        shade(tmp, opt, r, h, rng, &v, c);
        vx->kind = v.kind; vx->next = v.next; vx->weight = v.weight;
        return;
    }
    default:
        if (c) c->absorbed++;
        vx->kind = V_ABSORB; vx->weight = V3{0., 0., 0.};
        return;
    }
}

// camera.rs:460-522
inline V3 gray_colour(const GScene& sc, const Camera& cam, const Options& opt, Ray r, uint32_t pixel, uint32_t sample, Counters* c, bool* panicked) {
    V3 mult{1., 1., 1.}, res{0., 0., 0.};
    uint32_t depth = cam.max_depth, vertex = 1;
    if (c) c->paths++;
    for (;;) {
        if (depth == 0) { if (c) c->depth_out++; return V3{0., 0., 0.} + res; }
        GHit rec;
        if (!sc.world_hit(r, opt.tmin, &rec, c, panicked)) { if (c) c->missed++; return mult * cam.background + res; }
        Stream rng(opt.seed, pixel, sample, vertex, opt.rng_mode);
        GVertex vx;
        gshade(sc, opt, r, rec, rng, &vx, c, panicked);
        if (vx.kind == V_ABSORB) return mult * vx.emitted + res;
        if (vx.kind == V_DIFFUSE) res = res + mult * vx.emitted;
        mult = mult * vx.weight;
        r = vx.next;
        depth -= 1; vertex += 1;
    }
}

inline void grender(const GScene& sc, const Camera& cam, const Options& opt, double* rgb_sum, Counters* total, bool* panicked_out) {
    uint32_t W = cam.image_width, H = cam.image_height;
    int nt = opt.threads > 0 ? opt.threads : (int)std::max(1u, std::thread::hardware_concurrency());
    std::atomic<uint32_t> next_row{0};
    std::vector<Counters> cs(nt);
    std::vector<char> pan(nt, 0);
    auto worker = [&](int tid) {
        bool panicked = false;
        for (;;) {
            uint32_t j = next_row.fetch_add(1);
            if (j >= H) break;
            for (uint32_t i = 0; i < W; ++i) {
                uint32_t pixel = j * W + i;
                V3 acc{0., 0., 0.};
                for (uint32_t s = 0; s < cam.samples_per_pixel; ++s) {
                    Stream rng(opt.seed, pixel, s, 0, opt.rng_mode);
                    Ray r = get_ray(cam, i, j, rng);
                    V3 v = gray_colour(sc, cam, opt, r, pixel, s, &cs[tid], &panicked);
                    if (opt.fix_nan) { if (v.x != v.x) v.x = 0.; if (v.y != v.y) v.y = 0.; if (v.z != v.z) v.z = 0.; }
                    acc = acc + v;
                }
                rgb_sum[3 * (size_t)pixel] = acc.x; rgb_sum[3 * (size_t)pixel + 1] = acc.y; rgb_sum[3 * (size_t)pixel + 2] = acc.z;
            }
        }
        pan[tid] = panicked;
    };
    if (nt == 1) worker(0);
    else { std::vector<std::thread> th; for (int t = 0; t < nt; ++t) th.emplace_back(worker, t); for (auto& t : th) t.join(); }
    if (total) for (auto& c : cs) total->add(c);
    if (panicked_out) { *panicked_out = false; for (char p : pan) *panicked_out |= (p != 0); }
}

}  // namespace orcg
