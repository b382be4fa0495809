// general_host.hpp — host-side (f64) constructors of the general-scene entities: what Quad::new, Cuboid::new,
// Transformed<T>::get_aabbox, Transformation::{then, inverse}, rotation() and Perlin::new compute once per scene in
// the reference.  The device tables of rtw_general.cuh are filled from these.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#include "../../include/rtw.h"
#include "bvh_build.hpp"
#include "rtw_device.cuh"

namespace rtw {
namespace host {

struct D3 { double x, y, z; };
inline D3 operator+(D3 a, D3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline D3 operator-(D3 a, D3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline D3 operator-(D3 a) { return {-a.x, -a.y, -a.z}; }
inline D3 operator*(D3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline D3 operator/(D3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }
inline double dot3(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline D3 cross3(D3 a, D3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline D3 ld3(const double* p) { return {p[0], p[1], p[2]}; }

// AABBox::enclose + pad_to_minimum (geometry/src/aabox.rs:129-175): every enclose pads axes thinner than 1e-4
inline void pad_to_minimum(Box& b) {
    const double DELTA = 0.0001;
    for (int a = 0; a < 3; ++a)
        if (b.mx[a] - b.mn[a] < DELTA) { b.mn[a] -= DELTA; b.mx[a] += DELTA; }
}
inline Box enclose(Box b, const Box& o) {
    for (int a = 0; a < 3; ++a) { b.mn[a] = std::fmin(b.mn[a], o.mn[a]); b.mx[a] = std::fmax(b.mx[a], o.mx[a]); }
    pad_to_minimum(b);
    return b;
}
inline Box point_box(D3 p) { Box b; b.mn[0] = b.mx[0] = p.x; b.mn[1] = b.mx[1] = p.y; b.mn[2] = b.mx[2] = p.z; return b; }
inline Box box_from_points(const D3* p, int n) {                    // aabox.rs:191-204
    Box b = point_box(p[0]);
    for (int i = 1; i < n; ++i) b = enclose(b, point_box(p[i]));
    return b;
}

struct QuadH { D3 q, u, v, w, normal; double area; Box box; };
inline QuadH make_quad(D3 q, D3 u, D3 v, bool tri) {                // quadrilateral.rs:36-56 / triangles.rs:34-54
    QuadH s;
    s.q = q; s.u = u; s.v = v;
    D3 pts[5] = {q + (u + v) * 0.5, q, q + v, q + u, (q + u) + v};
    s.box = box_from_points(pts, tri ? 4 : 5);
    D3 n = cross3(u, v);
    s.w = n / dot3(n, n);
    double len = std::sqrt(dot3(n, n));
    s.area = tri ? len / 2. : len;
    s.normal = n / s.area;                                          // the triangle's "normal" has length 2, as in the reference
    return s;
}
inline void make_cuboid(D3 p, D3 q, QuadH out[6], Box* box) {       // cuboid.rs:26-50, 62-73
    D3 pts[2] = {p, q};
    Box b = box_from_points(pts, 2);
    D3 mn{b.mn[0], b.mn[1], b.mn[2]}, mx{b.mx[0], b.mx[1], b.mx[2]}, d = mx - mn;
    D3 dx{d.x, 0., 0.}, dy{0., d.y, 0.}, dz{0., 0., d.z};
    out[0] = make_quad(mn, dx, dy, false); out[1] = make_quad(mn, dy, dz, false); out[2] = make_quad(mn, dx, dz, false);
    out[3] = make_quad(mx, -dx, -dy, false); out[4] = make_quad(mx, -dy, -dz, false); out[5] = make_quad(mx, -dx, -dz, false);
    Box acc = out[0].box;
    for (int i = 1; i < 6; ++i) acc = enclose(acc, out[i].box);
    *box = acc;
}

// Matrix3 / Transformation (geometry/src/matrix3.rs, transformations.rs:96-136); rotation is row-major
inline D3 mat_vec(const double* m, D3 v) { return {dot3({m[0], m[1], m[2]}, v), dot3({m[3], m[4], m[5]}, v), dot3({m[6], m[7], m[8]}, v)}; }
inline void transform_then(const rtw_transform& a, const rtw_transform& b, rtw_transform* out) {   // a.apply(b): rotation = b.R * a.R, translation = b.t + b.R * a.t
    rtw_transform r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            r.rotation[3 * i + j] = dot3({b.rotation[3 * i], b.rotation[3 * i + 1], b.rotation[3 * i + 2]}, {a.rotation[j], a.rotation[3 + j], a.rotation[6 + j]});
    D3 t = ld3(b.translation) + mat_vec(b.rotation, ld3(a.translation));
    r.translation[0] = t.x; r.translation[1] = t.y; r.translation[2] = t.z;
    *out = r;
}
inline bool transform_inverse(const rtw_transform& t, rtw_transform* out) {
    const double a = t.rotation[0], b = t.rotation[1], c = t.rotation[2], d = t.rotation[3], e = t.rotation[4], f = t.rotation[5],
                 g = t.rotation[6], h = t.rotation[7], i = t.rotation[8];
    double det = a * (e * i - f * h) + b * (f * g - d * i) + c * (d * h - e * g);
    if (!std::isnormal(det)) return false;
    const double A = e * i - f * h, B = f * g - d * i, C = d * h - e * g;
    const double D = c * h - b * i, E = a * i - c * g, F = b * g - a * h;
    const double G = b * f - c * e, H = c * d - a * f, I = a * e - b * d;
    rtw_transform r;
    double m[9] = {A / det, D / det, G / det, B / det, E / det, H / det, C / det, F / det, I / det};
    std::memcpy(r.rotation, m, sizeof(m));
    D3 tr = -mat_vec(r.rotation, ld3(t.translation));
    r.translation[0] = tr.x; r.translation[1] = tr.y; r.translation[2] = tr.z;
    *out = r;
    return true;
}
inline void make_rotation(double angle_degrees, int axis, rtw_transform* out) {    // transformations.rs:38-64
    double angle = angle_degrees * (3.14159265358979323846264338327950288 / 180.);
    double c = std::cos(angle), s = std::sin(angle);
    rtw_transform r{};
    if (axis == 0) { double m[9] = {1., 0., 0., 0., c, -s, 0., s, c}; std::memcpy(r.rotation, m, sizeof(m)); }
    else if (axis == 1) { double m[9] = {c, 0., s, 0., 1., 0., -s, 0., c}; std::memcpy(r.rotation, m, sizeof(m)); }
    else { double m[9] = {c, -s, 0., s, c, 0., 0., 0., 1.}; std::memcpy(r.rotation, m, sizeof(m)); }
    *out = r;
}
// Transformed<T>::get_aabbox (transformations.rs:224-233): box of the eight transformed corners (aabox.rs:114-126)
inline Box transformed_box(const Box& ib, const rtw_transform& t) {
    D3 pts[8] = {{ib.mn[0], ib.mn[1], ib.mn[2]}, {ib.mn[0], ib.mx[1], ib.mn[2]}, {ib.mn[0], ib.mn[1], ib.mx[2]}, {ib.mn[0], ib.mx[1], ib.mx[2]},
                 {ib.mx[0], ib.mn[1], ib.mn[2]}, {ib.mx[0], ib.mx[1], ib.mn[2]}, {ib.mx[0], ib.mn[1], ib.mx[2]}, {ib.mx[0], ib.mx[1], ib.mx[2]}};
    for (D3& p : pts) p = mat_vec(t.rotation, p) + ld3(t.translation);
    return box_from_points(pts, 8);
}

// Perlin::new (perlin.rs:29-57), seeded: stream (seed; 0x9E71A000 + index, 0, 0), 53-bit uniforms.
inline void perlin_generate(uint64_t seed, uint32_t index, rtw_perlin* out) {
    Stream<true> rng(seed, 0x9E71A000u + index, 0u, 0u);
    for (int i = 0; i < 256; ++i) {
        for (;;) {                                                  // UnitSphere (utils.rs:99-122), shuffle dropped
            double a = 2. * standard(rng) - 1., b = 2. * standard(rng) - 1., c = 2. * standard(rng) - 1.;
            if (a * a + b * b + c * c < 1.) { out->rand_vec[i][0] = a; out->rand_vec[i][1] = b; out->rand_vec[i][2] = c; break; }
        }
    }
    uint8_t* perms[3] = {out->perm_x, out->perm_y, out->perm_z};
    for (uint8_t* pm : perms) {
        for (int i = 0; i < 256; ++i) pm[i] = (uint8_t)i;
        for (uint32_t i = 0; i < 255; ++i) {                        // j = Uniform::new(i, 256).sample (perlin.rs:37-44)
            uint32_t j = i + uindex(rng, 256u - i);
            uint8_t t = pm[i]; pm[i] = pm[j]; pm[j] = t;
        }
    }
}

}  // namespace host
}  // namespace rtw
