// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product.
//
// CPU restatement (C++17, f64, compile with -ffp-contract=off) of the hot path of
// N9199/ray_tracing_weekend: Camera::get_ray -> BoundedVolumeHierarchy::hit -> Sphere::hit ->
// Material::scatter (Lambertian / Metal / Dialectric) with the MixturePdf light sampling ->
// spp accumulation -> gamma / quantise.  Every function cites the reference file:line it follows
// (paths relative to /root/reference).
//
// PARITY UNPINNED: the reference ships no golden vectors, known-answer tests or fixtures for this
// path (its tests only assert "does not panic"), and it cannot be compiled here (needs nightly Rust
// + ~90 crates, no toolchain / network).  This oracle is therefore pinned only by (a) line-by-line
// review against the cited reference lines, (b) closed-form KATs (tests/), (c) Philox KATs that
// agree with /usr/local/cuda/include/curand_philox4x32_x.h constants.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this code.  The product (ray_tracing_weekend_b200/) never links, imports or calls it.
//
// Deliberate, documented deviations from the reference (all confined to the RNG, which the
// reference leaves unseeded and therefore unreproducible — camera.rs:346, scenes/src/lib.rs:172):
//   * random numbers come from a counter-based Philox4x32-10 stream keyed by (seed) and indexed by
//     (pixel, sample, path-vertex, draw-slot) instead of rand::SmallRng;  the *distributions* are the
//     reference's (rand 0.8.5 Standard / Open01 / Uniform::new_inclusive semantics restated below);
//   * UnitSphere's 3-element shuffle (utils.rs:115) is dropped: shuffling i.i.d. components is a
//     statistical no-op;
//   * lights.random's IteratorRandom::choose (hittable_list.rs:414-419) becomes one uniform index.
#pragma once
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <optional>
#include <thread>
#include <utility>
#include <vector>

namespace orc {

constexpr double PI = 3.14159265358979323846264338327950288;   // std::f64::consts::PI
constexpr double EPS = 2.220446049250313e-16;                    // f64::EPSILON = 2^-52
constexpr double INF = std::numeric_limits<double>::infinity();

// Rust f64::max / f64::min ignore a NaN operand == C fmax / fmin.
inline double rmax(double a, double b) { return std::fmax(a, b); }
inline double rmin(double a, double b) { return std::fmin(a, b); }

// f64::total_cmp (IEEE totalOrder): -NaN < -inf < ... < -0 < +0 < ... < +inf < +NaN.
inline int total_cmp(double a, double b) {
    int64_t x, y;
    std::memcpy(&x, &a, 8);
    std::memcpy(&y, &b, 8);
    x ^= (int64_t)((uint64_t)(x >> 63) >> 1);
    y ^= (int64_t)((uint64_t)(y >> 63) >> 1);
    return (x > y) - (x < y);
}

// ------------------------------------------------------------------------------------------------
// geometry/src/vec3/vec.rs:9-241
struct V3 {
    double x = 0., y = 0., z = 0.;
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }            // vec.rs:150-157
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }            // vec.rs:165-172
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }                                  // vec.rs:174-180
inline V3 operator*(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }               // vec.rs:209-215
inline V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }             // vec.rs:200-207
inline V3 operator/(V3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }               // vec.rs:217-223
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }               // vec.rs:68-72
inline V3 cross(V3 a, V3 b) {                                                             // vec.rs:74-82
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline double square_length(V3 a) { return dot(a, a); }                                   // vec.rs:62-66
inline double length(V3 a) { return std::sqrt(square_length(a)); }                        // vec.rs:56-60
inline V3 normalize(V3 a) { return a / length(a); }                                       // vec.rs:84-94
inline bool is_near_zero(V3 a) {                                                          // vec.rs:96-101
    const double e = 1e-8;
    return std::fabs(a.x) < e && std::fabs(a.y) < e && std::fabs(a.z) < e;
}
inline V3 reflect(V3 s, V3 o) { return s - (o * 2.) * dot(s, o); }                        // vec.rs:103-107
inline V3 refract(V3 s, V3 o, double etai_over_etat) {                                    // vec.rs:109-116
    double cos_theta = rmin(dot(s, -o), 1.);
    V3 r_out_perp = (s + o * cos_theta) * etai_over_etat;
    V3 r_out_parallel = o * (-std::sqrt(1. - square_length(r_out_perp)));
    return r_out_perp + r_out_parallel;
}

// shared/src/ray.rs:4-29
struct Ray {
    V3 o, d;
    V3 at(double t) const { return o + d * t; }
};

// geometry/src/onb.rs:8-35
struct Onb {
    V3 u, v, w;
    explicit Onb(V3 normal) {
        w = normalize(normal);
        V3 a = std::fabs(w.x) > 0.9 ? V3{0., 1., 0.} : V3{1., 0., 0.};
        v = normalize(cross(w, a));
        u = cross(w, v);
    }
    // (0..3).map(|i| self.0[i] * v[i]).sum()  with Sum = fold(Vec3::default(), +)   vec.rs:283-287
    V3 transform(V3 a) const { return ((V3{0., 0., 0.} + u * a.x) + v * a.y) + w * a.z; }
};

// ------------------------------------------------------------------------------------------------
// geometry/src/aabox.rs:86-190
struct AABB {
    V3 mn, mx;
    void pad_to_minimum() {                                                               // aabox.rs:129-149
        const double DELTA = 0.0001;
        double dx = mx.x - mn.x, dy = mx.y - mn.y, dz = mx.z - mn.z;
        if (dx < DELTA) { mn.x -= DELTA; mx.x += DELTA; }
        if (dy < DELTA) { mn.y -= DELTA; mx.y += DELTA; }
        if (dz < DELTA) { mn.z -= DELTA; mx.z += DELTA; }
    }
    AABB enclose(const AABB& b) const {                                                   // aabox.rs:161-175
        AABB r = *this;
        r.mn.x = rmin(r.mn.x, b.mn.x); r.mx.x = rmax(r.mx.x, b.mx.x);
        r.mn.y = rmin(r.mn.y, b.mn.y); r.mx.y = rmax(r.mx.y, b.mx.y);
        r.mn.z = rmin(r.mn.z, b.mn.z); r.mx.z = rmax(r.mx.z, b.mx.z);
        r.pad_to_minimum();
        return r;
    }
    double lo(int axis) const { return axis == 0 ? mn.x : axis == 1 ? mn.y : mn.z; }      // aabox.rs:151-159
    double hi(int axis) const { return axis == 0 ? mx.x : axis == 1 ? mx.y : mx.z; }
    bool right_of(int axis, double coord) const { return lo(axis) > coord; }              // aabox.rs:182-185
};

struct Counters {
    uint64_t rays = 0;          // world.hit calls (camera.rs:473)
    uint64_t paths = 0;         // ray_colour_call invocations (camera.rs:326)
    uint64_t box_tests = 0;     // AABBox::hit calls (hittable.rs:38)
    uint64_t box_builds = 0;    // Node get_aabbox recomputations (bvh.rs:145-151), faithful mode only
    uint64_t node_visits = 0;   // BVH::hit invocations (bvh.rs:163)
    uint64_t sphere_tests = 0;  // Sphere::hit calls from world.hit (sphere.rs:61)
    uint64_t plane_tests = 0;
    uint64_t light_tests = 0;   // Sphere::hit calls from pdf_value (sphere.rs:102)
    uint64_t lambertian = 0, metal = 0, dielectric = 0, absorbed = 0, missed = 0, depth_out = 0;
    void add(const Counters& o) {
        rays += o.rays; paths += o.paths; box_tests += o.box_tests; box_builds += o.box_builds;
        node_visits += o.node_visits; sphere_tests += o.sphere_tests; plane_tests += o.plane_tests;
        light_tests += o.light_tests; lambertian += o.lambertian; metal += o.metal;
        dielectric += o.dielectric; absorbed += o.absorbed; missed += o.missed; depth_out += o.depth_out;
    }
};

// shared/src/hittable.rs:38-87   AABoxHit for AABBox::hit  (returns is_some())
inline bool aabb_is_hit(const AABB& b, const Ray& r, double start, double end, Counters* c) {
    if (c) c->box_tests++;
    double x_tmin = (b.mn.x - r.o.x) / r.d.x;
    double x_tmax = (b.mx.x - r.o.x) / r.d.x;
    if (std::signbit(r.d.x)) std::swap(x_tmin, x_tmax);
    double tmin = x_tmin, tmax = x_tmax;
    double y_tmin = (b.mn.y - r.o.y) / r.d.y;
    double y_tmax = (b.mx.y - r.o.y) / r.d.y;
    if (std::signbit(r.d.y)) std::swap(y_tmin, y_tmax);
    if (tmax < y_tmin || tmin > y_tmax) return false;
    tmin = rmax(tmin, y_tmin);
    tmax = rmin(tmax, y_tmax);
    double z_tmin = (b.mn.z - r.o.z) / r.d.z;
    double z_tmax = (b.mx.z - r.o.z) / r.d.z;
    if (std::signbit(r.d.z)) std::swap(z_tmin, z_tmax);
    if (tmax < z_tmin || tmin > z_tmax) return false;
    tmin = rmax(tmin, z_tmin);
    tmax = rmin(tmax, z_tmax);
    return rmax(start, tmin) <= rmin(end, tmax);
}

// ------------------------------------------------------------------------------------------------
// shared/src/material.rs  — DynMaterial flattened to a POD (kind + SolidColour albedo + parameter)
enum MatKind : uint32_t { LAMBERTIAN = 0, METAL = 1, DIELECTRIC = 2, INVISIBLE = 3 };
struct Material {
    uint32_t kind = LAMBERTIAN;
    V3 albedo;       // Lambertian: SolidColour (texture.rs:15-22); Metal: albedo (material.rs:378-381)
    double param = 0.;  // Metal: fuzz; Dialectric: index_of_refraction (material.rs:423-425)
};

// shared/src/hittable.rs:90-129  HitRecord::new   (u, v are carried only for textures; SolidColour
// ignores them, so the oracle does not evaluate get_sphere_uv's atan2/acos — sphere.rs:49-54.)
struct HitRecord {
    V3 p, normal;
    double t = 0.;
    bool front_face = false;
    int32_t prim = -1;     // index into the world's primitive order (planes first, then spheres)
    uint32_t mat = 0;
};
inline HitRecord make_record(const Ray& r, double t, V3 outward_normal, int32_t prim, uint32_t mat) {
    HitRecord h;
    h.p = r.at(t);
    h.front_face = dot(r.d, outward_normal) < 0.;
    h.normal = h.front_face ? outward_normal : -outward_normal;
    h.t = t;
    h.prim = prim;
    h.mat = mat;
    return h;
}

// shared/src/entities/sphere.rs:25-127
struct Sphere {
    V3 center;
    double radius = 0.;
    uint32_t mat = 0;
    int32_t id = -1;
    AABB box;
    static Sphere make(V3 c, double r, uint32_t mat, int32_t id) {                        // sphere.rs:33-47
        Sphere s;
        s.center = c; s.radius = r; s.mat = mat; s.id = id;
        s.box = AABB{{c.x - r, c.y - r, c.z - r}, {c.x + r, c.y + r, c.z + r}};
        return s;
    }
    // sphere.rs:61-99
    bool hit(const Ray& r, double start, double end, HitRecord* out) const {
        V3 oc = r.o - center;
        double a = square_length(r.d);
        double half_b = dot(r.d, oc);
        double c = square_length(oc) - radius * radius;
        double discriminant = half_b * half_b - a * c;
        if (!(discriminant > 0.)) return false;
        double sq = std::sqrt(discriminant);
        double root = (-half_b - sq) / a;
        if (!(start <= root && root <= end)) {
            root = (-half_b + sq) / a;
            if (!(start <= root && root <= end)) return false;
        }
        if (out) {
            V3 p = r.at(root);
            V3 outward_normal = (p - center) / radius;
            *out = make_record(r, root, outward_normal, id, mat);
        }
        return true;
    }
    // sphere.rs:101-111
    double pdf_value(V3 origin, V3 direction, Counters* c) const {
        if (c) c->light_tests++;
        if (hit(Ray{origin, direction}, 0., INF, nullptr)) {
            double distance_squared = square_length(center - origin);
            double cos_theta_max = std::sqrt(1. - radius * radius / distance_squared);
            double solid_angle = 2. * PI * (1. - cos_theta_max);
            return 1. / solid_angle;
        }
        return 0.;
    }
};

// shared/src/entities/plane.rs:21-113
struct Plane {
    V3 point, normal;
    uint32_t mat = 0;
    int32_t id = -1;
    static Plane make(V3 p, V3 n, uint32_t mat, int32_t id) {                             // plane.rs:27-39
        Plane q; q.point = p; q.normal = normalize(n); q.mat = mat; q.id = id; return q;
    }
    // plane.rs:61-76.  get_plane_uv (:41-55) only feeds textures and a finiteness panic; the oracle
    // reports that panic condition through *panicked instead of aborting.
    bool hit(const Ray& r, double start, double end, HitRecord* out, bool* panicked) const {
        double denom = dot(r.d, normal);
        if (!(denom > EPS)) return false;
        double t = -dot(r.o - point, normal) / denom;
        V3 p = r.at(t);
        if (panicked && !(std::isfinite(p.x) && std::isfinite(p.z))) *panicked = true;
        if (!(start <= t && t <= end)) return false;                                      // range.contains(&t)
        if (out) *out = make_record(r, t, normal, id, mat);
        return true;
    }
    AABB get_aabbox() const {                                                             // plane.rs:78-107
        auto small = [](double v) { return std::fabs(v) < EPS; };
        AABB b;
        bool fx = small(normal.z) && small(normal.y);
        bool fy = small(normal.x) && small(normal.z);
        bool fz = small(normal.x) && small(normal.y);
        b.mn.x = fx ? 0. : -INF; b.mx.x = fx ? 0. : INF;
        b.mn.y = fy ? 0. : -INF; b.mx.y = fy ? 0. : INF;
        b.mn.z = fz ? 0. : -INF; b.mx.z = fz ? 0. : INF;
        return b;
    }
};

// ------------------------------------------------------------------------------------------------
// shared/src/hittable_collections/hittable_list.rs:247-420 (vector_based) + hittable_list/raw.rs.
// The reference buckets objects by TypeId; the order of TypeIds is compiler-chosen and only matters
// for exact-t ties, so the oracle fixes it as [Plane, Sphere].
struct HittableList {
    std::vector<Plane> planes;  std::optional<AABB> planes_box;     // RawHittableVec + cached_aabox (raw.rs:145-198)
    std::vector<Sphere> spheres; std::optional<AABB> spheres_box;
    size_t len = 0;
    std::optional<AABB> aabbox;

    void add(const Plane& o) {                                                            // hittable_list.rs:270-294, raw.rs:185-198
        AABB b = o.get_aabbox();
        aabbox = aabbox ? aabbox->enclose(b) : b;
        planes.push_back(o);
        planes_box = planes_box ? planes_box->enclose(b) : b;
        len++;
    }
    void add(const Sphere& o) {
        aabbox = aabbox ? aabbox->enclose(o.box) : o.box;
        spheres.push_back(o);
        spheres_box = spheres_box ? spheres_box->enclose(o.box) : o.box;
        len++;
    }
    AABB get_aabbox() const { return aabbox ? *aabbox : AABB{}; }                         // hittable_list.rs:422-425

    // hittable_list.rs:394-406 -> raw.rs:262-275 -> utils.rs:67-79 -> hittable.rs:191-196
    bool hit(const Ray& r, double start, double end, HitRecord* best, Counters* c, bool* panicked) const {
        bool any = false;
        HitRecord rec;
        if (!planes.empty() && aabb_is_hit(*planes_box, r, start, end, c)) {
            bool bany = false; HitRecord brec;
            for (const Plane& o : planes) {
                if (!aabb_is_hit(o.get_aabbox(), r, start, end, c)) continue;
                if (c) c->plane_tests++;
                if (o.hit(r, start, end, &rec, panicked) && (!bany || rec.t < brec.t)) { brec = rec; bany = true; }
            }
            if (bany && (!any || brec.t < best->t)) { *best = brec; any = true; }
        }
        if (!spheres.empty() && aabb_is_hit(*spheres_box, r, start, end, c)) {
            bool bany = false; HitRecord brec;
            for (const Sphere& o : spheres) {
                if (!aabb_is_hit(o.box, r, start, end, c)) continue;
                if (c) c->sphere_tests++;
                if (o.hit(r, start, end, &rec) && (!bany || rec.t < brec.t)) { brec = rec; bany = true; }
            }
            if (bany && (!any || brec.t < best->t)) { *best = brec; any = true; }
        }
        return any;
    }

    // raw.rs:84-104 (pop from the back; !right_of -> "left") + hittable_list.rs:296-316, which
    // RETURNS (right, left).  best_split binds that pair as (left, right) — kept as is.
    std::pair<HittableList, HittableList> split_by(int axis, double coord) const {
        HittableList left, right;
        auto push_box = [](HittableList& dst, const AABB& b) { dst.aabbox = dst.aabbox ? dst.aabbox->enclose(b) : b; };
        {   // Plane bucket
            HittableList l, r;
            for (size_t i = planes.size(); i-- > 0;) (planes[i].get_aabbox().right_of(axis, coord) ? r : l).add(planes[i]);
            if (!r.planes.empty()) { right.len += r.planes.size(); right.planes = r.planes; right.planes_box = r.planes_box; push_box(right, *r.planes_box); }
            if (!l.planes.empty()) { left.len += l.planes.size(); left.planes = l.planes; left.planes_box = l.planes_box; push_box(left, *l.planes_box); }
        }
        {   // Sphere bucket
            HittableList l, r;
            for (size_t i = spheres.size(); i-- > 0;) (spheres[i].box.right_of(axis, coord) ? r : l).add(spheres[i]);
            if (!r.spheres.empty()) { right.len += r.spheres.size(); right.spheres = r.spheres; right.spheres_box = r.spheres_box; push_box(right, *r.spheres_box); }
            if (!l.spheres.empty()) { left.len += l.spheres.size(); left.spheres = l.spheres; left.spheres_box = l.spheres_box; push_box(left, *l.spheres_box); }
        }
        return {right, left};
    }

    // hittable_list.rs:318-379
    struct Split { int axis = 0; double coord = 0.; };
    Split best_split_plane() const {
        size_t best0 = std::numeric_limits<size_t>::max();
        double best1 = INF; int best_axis = 0; double best_coord = 0.;
        std::vector<std::pair<double, double>> tmp;
        for (int axis = 0; axis < 3; ++axis) {
            tmp.clear();
            for (const Plane& o : planes) { AABB b = o.get_aabbox(); tmp.push_back({b.lo(axis), b.hi(axis)}); }
            for (const Sphere& o : spheres) tmp.push_back({o.box.lo(axis), o.box.hi(axis)});
            std::stable_sort(tmp.begin(), tmp.end(), [](const auto& a, const auto& b) {
                int c = total_cmp(a.first, b.first);
                if (c == 0) c = total_cmp(a.second, b.second);
                return c < 0;
            });
            double median = tmp[tmp.size() / 2].first;
            size_t partition_point = 0;
            while (partition_point < tmp.size() && total_cmp(tmp[partition_point].first, median) < 0) partition_point++;
            double bbox_axis_size = tmp.back().second - tmp.front().first;
            size_t cand0 = tmp.size() - 2 * partition_point;
            // (best.0, -best.1) > (cand0, -bbox_axis_size)  — lexicographic PartialOrd on (usize, f64)
            bool better = best0 > cand0 || (best0 == cand0 && (-best1) > (-bbox_axis_size));
            if (better) { best0 = cand0; best1 = bbox_axis_size; best_axis = axis; best_coord = median; }
        }
        return {best_axis, best_coord};
    }
};

// shared/src/hittable_collections/bvh.rs:26-34, 106-188 (plane_divided; the live BVH)
struct Bvh {
    bool leaf = true;
    HittableList list;                      // Leaf
    std::unique_ptr<Bvh> left, right;       // Node
    size_t len = 0;
    AABB cached;                            // == get_aabbox(); used when !faithful

    static std::unique_ptr<Bvh> from(const HittableList& value) {                         // bvh.rs:106-143
        auto n = std::make_unique<Bvh>();
        if (value.len <= 5) { n->leaf = true; n->list = value; n->len = value.len; n->cached = value.get_aabbox(); return n; }
        size_t len = value.len;
        HittableList::Split sp = value.best_split_plane();
        auto pr = value.split_by(sp.axis, sp.coord);
        HittableList& l = pr.first; HittableList& r = pr.second;   // (left, right) = self.split_by(plane)
        if (len == l.len) { n->leaf = true; n->list = l; n->len = l.len; n->cached = l.get_aabbox(); return n; }
        if (len == r.len) { n->leaf = true; n->list = r; n->len = r.len; n->cached = r.get_aabbox(); return n; }
        n->leaf = false;
        n->left = from(l);
        n->right = from(r);
        n->len = n->left->len + n->right->len;
        n->cached = n->left->cached.enclose(n->right->cached);
        return n;
    }
    AABB get_aabbox(Counters* c) const {                                                  // bvh.rs:145-151
        if (leaf) return list.get_aabbox();
        if (c) c->box_builds++;
        return left->get_aabbox(c).enclose(right->get_aabbox(c));
    }
    size_t depth() const { return leaf ? 1 : std::max(left->depth(), right->depth()) + 1; }      // bvh.rs:37-44
    size_t node_count() const { return leaf ? 1 : left->node_count() + right->node_count() + 1; } // bvh.rs:46-53
    size_t leaf_count() const { return leaf ? 1 : left->leaf_count() + right->leaf_count(); }
    size_t max_leaf() const { return leaf ? list.len : std::max(left->max_leaf(), right->max_leaf()); }

    // bvh.rs:163-188.  faithful: recompute child boxes recursively on every visit like the reference.
    bool hit(const Ray& r, double start, double end, HitRecord* out, Counters* c, bool faithful, bool* panicked) const {
        if (c) c->node_visits++;
        if (leaf) return list.hit(r, start, end, out, c, panicked);
        HitRecord a, b;
        bool ha = aabb_is_hit(faithful ? left->get_aabbox(c) : left->cached, r, start, end, c) &&
                  left->hit(r, start, end, &a, c, faithful, panicked);
        bool hb = aabb_is_hit(faithful ? right->get_aabbox(c) : right->cached, r, start, end, c) &&
                  right->hit(r, start, end, &b, c, faithful, panicked);
        if (!ha && !hb) return false;
        if (ha && hb) *out = (b.t < a.t) ? b : a;     // min_by keeps the first on ties
        else *out = ha ? a : b;
        return true;
    }
};

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011; constants as in /usr/local/cuda/include/curand_philox4x32_x.h:88-91)
inline void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int i = 0; i < 10; ++i) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

enum RngMode : uint32_t { W64 = 0, W32 = 1 };
enum MathMode : uint32_t { LIBM = 0, PORTABLE = 1 };

// Stream of uniforms for one (pixel, sample, vertex): counter = (pixel, sample, vertex, block),
// key = seed.  The 32-bit output words of consecutive blocks form one word sequence x[0], x[1], ...
//   W64: uniform k is built from the 64-bit word  x[2k] | x[2k+1] << 32   (53-bit, rand 0.8 semantics)
//   W32: uniform k is built from the 32-bit word  x[k]                    (24-bit, the f32 fast mode)
struct Stream {
    uint32_t key[2], pixel, sample, vertex, mode;
    uint32_t k = 0, cached_block = 0xffffffffu, buf[4];
    Stream(uint64_t seed, uint32_t pixel_, uint32_t sample_, uint32_t vertex_, uint32_t mode_)
        : pixel(pixel_), sample(sample_), vertex(vertex_), mode(mode_) {
        key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed >> 32);
    }
    uint32_t word32(uint32_t idx) {
        uint32_t block = idx >> 2;
        if (block != cached_block) {
            uint32_t ctr[4] = {pixel, sample, vertex, block};
            philox4x32_10(ctr, key, buf);
            cached_block = block;
        }
        return buf[idx & 3];
    }
    uint64_t next64() { uint64_t lo = word32(2 * k), hi = word32(2 * k + 1); k++; return lo | (hi << 32); }
    uint32_t next32() { return word32(k++); }
    // rand 0.8.5 `Standard` for f64: (next_u64() >> 11) as f64 * 2^-53  in [0,1)
    double standard() {
        if (mode == W64) return (double)(next64() >> 11) * 0x1.0p-53;
        return (double)(next32() >> 8) * 0x1.0p-24;
    }
    // rand 0.8.5 `Open01` for f64: 52 mantissa bits into [1,2) minus (1 - eps/2)  ==  (m + 0.5) * 2^-52
    double open01() {
        if (mode == W64) return (double)(next64() >> 12) * 0x1.0p-52 + 0x1.0p-53;
        return (double)(next32() >> 9) * 0x1.0p-23 + 0x1.0p-24;
    }
    // rand 0.8.5 `Uniform::new_inclusive(low, high)` for f64: scale = (high-low)/(1-eps), nudged down
    // until low + scale*(1-eps) <= high;  sample = (u52 in [0,1)) * scale + low.
    double uniform_inclusive(double low, double high) {
        if (mode == W64) {
            const double max_rand = 1. - EPS;
            double scale = (high - low) / max_rand;
            while (scale * max_rand + low > high) scale = std::nextafter(scale, -INF);
            return (double)(next64() >> 12) * 0x1.0p-52 * scale + low;
        }
        return (double)(next32() >> 8) * 0x1.0p-24 * (high - low) + low;
    }
    // uniform index in [0, n): widening multiply (replaces IteratorRandom::choose)
    uint32_t index(uint32_t n) {
        if (mode == W64) return (uint32_t)(((unsigned __int128)next64() * n) >> 64);
        return (uint32_t)(((uint64_t)next32() * n) >> 32);
    }
};

// sin/cos of phi in [0, 2*pi].  LIBM: what the reference's f64::sin/cos lower to on this platform.
// PORTABLE: fixed sequence of IEEE operations (Cody-Waite by pi/2 + fdlibm kernel polynomials, no FMA)
// that the f64 CUDA kernels repeat bit for bit.
inline void sincos_phi(double phi, uint32_t math, double* s, double* c) {
    if (math == LIBM) { *s = std::sin(phi); *c = std::cos(phi); return; }
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

// ------------------------------------------------------------------------------------------------
// shared/src/camera.rs:28-261
struct CameraBuilder {
    std::optional<double> aspect_ratio;
    std::optional<uint32_t> image_width, image_height;
    uint32_t samples_per_pixel = 10, max_depth = 10;
    V3 background{0., 0., 0.};
    double vfov = 90.;
    V3 lookfrom{0., 0., 0.}, lookat{0., 0., -1.}, vup{0., 1., 0.};
    double defocus_angle = 0., focus_dist = 10.;
};
struct Camera {
    uint32_t image_width = 0, image_height = 0, samples_per_pixel = 0, max_depth = 0;
    V3 background, center, pixel00_loc, pixel_delta_u, pixel_delta_v, defocus_disk_u, defocus_disk_v;
    double defocus_angle = 0.;
};
inline double rust_round(double v) { return std::round(v); }   // f64::round: half away from zero
inline Camera camera_build(const CameraBuilder& b) {                                      // camera.rs:114-218
    double aspect; uint32_t h, w;
    const auto &A = b.aspect_ratio; const auto &H = b.image_height; const auto &W = b.image_width;
    if (!A && !H && !W) { aspect = 1.; h = 100; w = 100; }
    else if (!A && !H && W) { aspect = 1.; h = *W; w = *W; }
    else if (!A && H && !W) { aspect = 1.; h = *H; w = *H; }
    else if (A && !H && !W) { aspect = *A; h = (uint32_t)rust_round(100. / *A); w = 100; }
    else if (!A && H && W) { aspect = (double)*W / (double)*H; h = *H; w = *W; }
    else if (A && !H && W) { aspect = *A; h = (uint32_t)rust_round((double)*W / *A); w = *W; }
    else if (A && H && !W) { aspect = *A; h = *H; w = (uint32_t)rust_round((double)*H * *A); }
    else { aspect = *A; h = *H; w = *W; }

    Camera c;
    V3 center = b.lookfrom;
    double theta = b.vfov * (PI / 180.);                       // f64::to_radians
    double hh = std::tan(theta / 2.);
    double viewport_height = 2. * hh * b.focus_dist;
    double viewport_width = viewport_height * aspect;
    V3 wv = b.lookfrom - b.lookat;
    if (is_near_zero(cross(b.vup, wv))) wv = wv + V3{0.1, 0., 0.};
    wv = normalize(wv);
    V3 u = normalize(cross(b.vup, wv));
    V3 v = cross(wv, u);
    V3 viewport_u = u * viewport_width;
    V3 viewport_v = v * viewport_height;
    c.pixel_delta_u = viewport_u / (double)w;
    c.pixel_delta_v = viewport_v / (double)h;
    V3 corner = ((center - (wv * b.focus_dist)) - viewport_u / 2.) - viewport_v / 2.;
    c.pixel00_loc = corner + (c.pixel_delta_u + c.pixel_delta_v) / 2.;
    double defocus_radius = std::tan(b.defocus_angle / 2.) * b.focus_dist;   // no deg->rad, as in the reference (:190)
    c.defocus_disk_u = u * defocus_radius;
    c.defocus_disk_v = v * defocus_radius;
    c.image_width = w; c.image_height = h;
    c.samples_per_pixel = b.samples_per_pixel; c.max_depth = b.max_depth;
    c.background = b.background; c.center = center; c.defocus_angle = b.defocus_angle;
    return c;
}

// ------------------------------------------------------------------------------------------------
struct Scene {
    std::vector<Material> materials;
    HittableList world_list;            // insertion order of scenes::simple (plane first)
    std::unique_ptr<Bvh> world;         // BoundedVolumeHierarchy::from(world)   scenes/lib.rs:228
    std::vector<Sphere> lights;         // plain HittableList of Invisible spheres   scenes/lib.rs:203,217,229
    void finalize() { world = Bvh::from(world_list); }
};

struct Options {
    uint64_t seed = 0;
    double tmin = EPS;                  // camera.rs:473
    uint32_t rng_mode = W64, math_mode = LIBM;
    bool faithful_bvh = false;          // true: recompute node boxes per visit like bvh.rs:145-151
    bool fix_nan = false;               // NOT reference behaviour: zero NaN components of a sample (Colour::fix_nan, colour.rs:52-58,
                                        // which only the dead recursive ray_colour applies, camera.rs:434)
    int threads = 0;
};

// shared/src/camera.rs:274-293
inline Ray get_ray(const Camera& cam, uint32_t i, uint32_t j, Stream& rng) {
    double ox = rng.uniform_inclusive(-0.5, 0.5);
    double oy = rng.uniform_inclusive(-0.5, 0.5);
    V3 pixel_sample = (cam.pixel00_loc + cam.pixel_delta_u * ((double)i + ox)) + cam.pixel_delta_v * ((double)j + oy);
    V3 origin = cam.center;
    if (!(cam.defocus_angle <= EPS)) {
        V3 p;                                                   // UnitDisk, utils.rs:124-144
        for (;;) {
            double a = 2. * rng.standard() - 1.;
            double b = 2. * rng.standard() - 1.;
            p = V3{a, 0., b};
            if (square_length(p) < 1.) break;
        }
        origin = (cam.center + cam.defocus_disk_u * p.x) + cam.defocus_disk_v * p.z;
    }
    return Ray{origin, pixel_sample - origin};
}

// One path vertex: world.hit + Material::scatter + (Lambertian) MixturePdf sampling.
enum VertexKind : uint32_t { V_MISS = 0, V_ABSORB = 1, V_SPECULAR = 2, V_DIFFUSE = 3 };
struct Vertex {
    uint32_t kind = V_MISS;
    HitRecord rec;
    Ray next;          // scattered / reflected ray
    V3 weight;         // factor applied to `mult`
};

inline bool world_hit(const Scene& sc, const Ray& r, double tmin, HitRecord* rec, Counters* c, bool faithful, bool* panicked) {
    if (c) c->rays++;
    return sc.world->hit(r, tmin, INF, rec, c, faithful, panicked);
}

// lights: HittableList::pdf_value (hittable_list.rs:408-412) and ::random (:414-420) over spheres
inline double lights_pdf_value(const Scene& sc, V3 origin, V3 direction, Counters* c) {
    double acc = 0.;
    for (const Sphere& s : sc.lights) acc = acc + s.pdf_value(origin, direction, c);
    return acc / (double)sc.lights.size();
}
inline V3 sphere_random(const Sphere& s, V3 origin, Stream& rng, uint32_t math) {         // sphere.rs:114-127
    V3 direction = s.center - origin;
    double distance = length(direction);
    Onb uvw(direction);
    double r1 = rng.standard();
    double r2 = rng.standard();
    double z = 1. + r1 * (std::sqrt(1. - s.radius * s.radius / (distance * distance)) - 1.);
    double phi = 2. * PI * r2;
    double sn, cs; sincos_phi(phi, math, &sn, &cs);
    double x = cs * std::sqrt(1. - z * z);
    double y = sn * std::sqrt(1. - z * z);
    return uvw.transform(V3{x, y, z});
}

// Material::scatter + the Scatter branch of ray_colour_tail_call (camera.rs:484-521)
inline void shade(const Scene& sc, const Options& opt, const Ray& r, const HitRecord& rec, Stream& rng,
                  Vertex* vx, Counters* c) {
    const Material& m = sc.materials[rec.mat];
    vx->rec = rec;
    switch (m.kind) {
    case LAMBERTIAN: {                                          // material.rs:357-376
        if (c) c->lambertian++;
        V3 attenuation = m.albedo;
        Onb uvw(rec.normal);                                    // CosinePdf::new, pdf.rs:39-43
        V3 dir;
        if (rng.standard() < 0.5) {                             // MixturePdf::generate, pdf.rs:94-100 (pdf1 = lights)
            uint32_t idx = rng.index((uint32_t)sc.lights.size());
            dir = sphere_random(sc.lights[idx], rec.p, rng, opt.math_mode);
        } else {                                                // CosineWeightedHemisphere, utils.rs:146-161
            double r1 = rng.standard();
            double r2 = rng.standard();
            double phi = 2. * PI * r1;
            double sn, cs; sincos_phi(phi, opt.math_mode, &sn, &cs);
            double x = cs * std::sqrt(r2);
            double y = sn * std::sqrt(r2);
            double z = std::sqrt(1. - r2);
            dir = uvw.transform(V3{x, y, z});
        }
        // MixturePdf::value, pdf.rs:90-92
        double light_v = lights_pdf_value(sc, rec.p, dir, c);
        double cos_v = rmax(dot(normalize(dir), uvw.w) / PI, 0.);   // CosinePdf::value, pdf.rs:46-49
        double pdf_value = light_v * 0.5 + cos_v * 0.5;
        // Lambertian::scattering_pdf, material.rs:372-375
        double scattering_pdf = rmax(dot(rec.normal, normalize(dir)) / PI, 0.);
        vx->kind = V_DIFFUSE;
        vx->next = Ray{rec.p, dir};
        vx->weight = (attenuation * scattering_pdf) / pdf_value;    // camera.rs:518
        return;
    }
    case METAL: {                                               // material.rs:407-421
        if (c) c->metal++;
        V3 reflected = reflect(normalize(r.d), rec.normal);
        V3 ball;                                                // UnitSphere (uniform in the unit ball), utils.rs:99-122
        for (;;) {
            double a = 2. * rng.standard() - 1.;
            double b = 2. * rng.standard() - 1.;
            double d = 2. * rng.standard() - 1.;
            ball = V3{a, b, d};
            if (square_length(ball) < 1.) break;
        }
        V3 dir = reflected + ball * m.param;
        if (dot(dir, rec.normal) > 0.) {
            vx->kind = V_SPECULAR; vx->next = Ray{rec.p, dir}; vx->weight = m.albedo;
        } else {
            if (c) c->absorbed++;
            vx->kind = V_ABSORB; vx->weight = V3{0., 0., 0.};
        }
        return;
    }
    case DIELECTRIC: {                                          // material.rs:457-488
        if (c) c->dielectric++;
        double ratio = rec.front_face ? 1. / m.param : m.param;
        V3 unit = normalize(r.d);
        double cos_theta = rmin(dot(unit, -rec.normal), 1.);
        double sin_theta = std::sqrt(1. - cos_theta * cos_theta);
        bool cannot_refract = ratio * sin_theta > 1.;
        bool do_reflect = cannot_refract;
        if (!do_reflect) {
            double r0 = (1. - ratio) / (1. + ratio);            // reflectance, material.rs:450-454
            r0 = r0 * r0;
            double om = 1. - cos_theta;
            double p5 = ((om * om) * (om * om)) * om;           // powi(5)
            do_reflect = (r0 + (1. - r0) * p5) > rng.open01();
        }
        V3 dir = do_reflect ? reflect(unit, rec.normal) : refract(unit, rec.normal, ratio);
        vx->kind = V_SPECULAR; vx->next = Ray{rec.p, dir}; vx->weight = V3{1., 1., 1.};
        return;
    }
    default:                                                    // Invisible: Material defaults, material.rs:32-49
        if (c) c->absorbed++;
        vx->kind = V_ABSORB; vx->weight = V3{0., 0., 0.};
        return;
    }
}

// shared/src/camera.rs:439-522 — the tail-recursive integrator written as its loop.
// Emission is identically 0 for Lambertian / Metal / Dialectric / Invisible (material.rs:42-44),
// so `res` stays 0; it is kept to mirror the state tuple.
inline V3 ray_colour(const Scene& sc, const Camera& cam, const Options& opt, Ray r, uint32_t pixel, uint32_t sample,
                     Counters* c, bool* panicked) {
    V3 mult{1., 1., 1.}, res{0., 0., 0.};
    uint32_t depth = cam.max_depth;
    uint32_t vertex = 1;
    if (c) c->paths++;
    for (;;) {
        if (depth == 0) { if (c) c->depth_out++; return V3{0., 0., 0.} + res; }
        HitRecord rec;
        if (!world_hit(sc, r, opt.tmin, &rec, c, opt.faithful_bvh, panicked)) {
            if (c) c->missed++;
            return mult * cam.background + res;
        }
        V3 emitted{0., 0., 0.};
        Stream rng(opt.seed, pixel, sample, vertex, opt.rng_mode);
        Vertex vx;
        shade(sc, opt, r, rec, rng, &vx, c);
        if (vx.kind == V_ABSORB) return mult * emitted + res;
        if (vx.kind == V_DIFFUSE) res = res + mult * emitted;
        mult = mult * vx.weight;
        r = vx.next;
        depth -= 1;
        vertex += 1;
    }
}

// shared/src/camera.rs:315-388 — per pixel: sum over spp, sequential in sample order.
inline void render(const Scene& sc, const Camera& cam, const Options& opt, double* rgb_sum /*[h][w][3], j=0 bottom*/,
                   Counters* total, bool* panicked_out, uint32_t row_begin = 0, uint32_t row_end = 0xffffffffu) {
    uint32_t W = cam.image_width, H = cam.image_height;
    row_end = std::min(row_end, H);
    int nt = opt.threads > 0 ? opt.threads : (int)std::max(1u, std::thread::hardware_concurrency());
    std::atomic<uint32_t> next_row{row_begin};
    std::vector<Counters> cs(nt);
    std::vector<char> pan(nt, 0);
    auto worker = [&](int tid) {
        Counters& c = cs[tid];
        bool panicked = false;
        for (;;) {
            uint32_t j = next_row.fetch_add(1);
            if (j >= row_end) break;
            for (uint32_t i = 0; i < W; ++i) {
                uint32_t pixel = j * W + i;
                V3 acc{0., 0., 0.};
                for (uint32_t s = 0; s < cam.samples_per_pixel; ++s) {
                    Stream rng(opt.seed, pixel, s, 0, opt.rng_mode);
                    Ray r = get_ray(cam, i, j, rng);
                    V3 v = ray_colour(sc, cam, opt, r, pixel, s, &c, &panicked);
                    if (opt.fix_nan) { if (v.x != v.x) v.x = 0.; if (v.y != v.y) v.y = 0.; if (v.z != v.z) v.z = 0.; }
                    acc = acc + v;
                }
                rgb_sum[3 * (size_t)pixel + 0] = acc.x;
                rgb_sum[3 * (size_t)pixel + 1] = acc.y;
                rgb_sum[3 * (size_t)pixel + 2] = acc.z;
            }
        }
        pan[tid] = panicked;
    };
    if (nt == 1) worker(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) th.emplace_back(worker, t);
        for (auto& t : th) t.join();
    }
    if (total) for (auto& c : cs) total->add(c);
    if (panicked_out) { *panicked_out = false; for (char p : pan) *panicked_out |= (p != 0); }
}

// shared/src/colour.rs:15-36: c/spp -> sqrt -> clamp(0,1) -> (256*x) as u8  (saturating, NaN -> 0)
inline uint8_t quantise(double sum, int32_t spp) {
    double scale = 1. / (double)spp;
    double v = std::sqrt(sum * scale);
    // f64::clamp(0,1): NaN stays NaN
    if (v < 0.) v = 0.;
    if (v > 1.) v = 1.;
    double q = 256. * v;
    if (std::isnan(q)) return 0;
    if (q >= 255.) return 255;
    if (q <= 0.) return 0;
    return (uint8_t)q;
}

// ------------------------------------------------------------------------------------------------
// scenes/src/lib.rs:155-233  scenes::simple, seeded.  Host draws: Philox stream (pixel = 0x5CE9E000,
// sample = 0, vertex = 0), W64 Standard.  Generalised by (n, p_lambertian, p_metal, ground) so the
// BASELINE stress configs use the same recipe: reference = (11, 0.8, 0.95, ground 0).
//   ground 0: the reference's one-sided Plane((0,0,0),(0,1,0)) with Lambertian(0.9)  (lib.rs:164-168)
//   ground 1: book-1 ground sphere (0,-1000,0) r=1000 Lambertian(0.5)   [variant, not the reference]
//   ground 2: none
struct SceneDesc {
    std::vector<double> spheres;        // n x 4
    std::vector<uint32_t> sphere_mat;
    std::vector<Material> materials;
    std::vector<double> planes;         // n x 6
    std::vector<uint32_t> plane_mat;
    std::vector<double> lights;         // n x 4
    CameraBuilder cam;
};
inline SceneDesc scene_simple(uint64_t seed, int n, double p_lambertian, double p_metal, int ground) {
    SceneDesc d;
    Stream rng(seed, 0x5CE9E000u, 0, 0, W64);
    auto add_mat = [&](uint32_t kind, V3 albedo, double param) { d.materials.push_back({kind, albedo, param}); return (uint32_t)d.materials.size() - 1; };
    auto add_sphere = [&](V3 c, double r, uint32_t m) { d.spheres.insert(d.spheres.end(), {c.x, c.y, c.z, r}); d.sphere_mat.push_back(m); };
    auto add_light = [&](V3 c, double r) { d.lights.insert(d.lights.end(), {c.x, c.y, c.z, r}); };
    if (ground == 0) {
        uint32_t gm = add_mat(LAMBERTIAN, {0.9, 0.9, 0.9}, 0.);
        d.planes.insert(d.planes.end(), {0., 0., 0., 0., 1., 0.});
        d.plane_mat.push_back(gm);
    } else if (ground == 1) {
        uint32_t gm = add_mat(LAMBERTIAN, {0.5, 0.5, 0.5}, 0.);
        add_sphere({0., -1000., 0.}, 1000., gm);
    }
    uint32_t material1 = add_mat(DIELECTRIC, {1., 1., 1.}, 1.5);
    for (int a = -n; a < n; ++a) {
        for (int b = -n; b < n; ++b) {
            double choose_mat = rng.standard();
            double cx = (double)a + 0.9 * rng.standard();
            double cz = (double)b + 0.9 * rng.standard();
            V3 center{cx, 0.2, cz};
            if (length(center - V3{4., 0.2, 0.}) > 0.9) {
                uint32_t mat;
                if (choose_mat < p_lambertian) {
                    V3 c1; c1.x = rng.standard(); c1.y = rng.standard(); c1.z = rng.standard();
                    V3 c2; c2.x = rng.standard(); c2.y = rng.standard(); c2.z = rng.standard();
                    mat = add_mat(LAMBERTIAN, c1 * c2, 0.);
                } else if (choose_mat < p_metal) {
                    V3 alb;                                     // random_f64_2 = Uniform::new_inclusive(0.5, 1.), utils.rs:93-97
                    alb.x = rng.uniform_inclusive(0.5, 1.); alb.y = rng.uniform_inclusive(0.5, 1.); alb.z = rng.uniform_inclusive(0.5, 1.);
                    double fuzz = 1. - rng.uniform_inclusive(0.5, 1.);
                    mat = add_mat(METAL, alb, fuzz);
                } else {
                    add_light(center, 0.2);
                    mat = material1;
                }
                add_sphere(center, 0.2, mat);
            }
        }
    }
    uint32_t material2 = add_mat(LAMBERTIAN, {0.4, 0.2, 0.1}, 0.);
    uint32_t material3 = add_mat(METAL, {0.7, 0.6, 0.5}, 0.);
    add_sphere({0., 1., 0.}, 1., material1);
    add_sphere({-4., 1., 0.}, 1., material2);
    add_sphere({4., 1., 0.}, 1., material3);
    add_light({0., 1., 0.}, 1.);
    d.cam.lookfrom = {10., 5., 10.};
    d.cam.lookat = {0., 0., 0.};
    d.cam.focus_dist = length(d.cam.lookfrom - d.cam.lookat);
    d.cam.vfov = 40.;
    d.cam.background = {1., 1., 1.};
    return d;
}

// world primitive ids: planes first (0..np-1), then spheres (np..np+ns-1), in insertion order.
inline std::unique_ptr<Scene> scene_from_arrays(size_t ns, const double* spheres, const uint32_t* sphere_mat,
                                                size_t nm, const Material* mats,
                                                size_t np, const double* planes, const uint32_t* plane_mat,
                                                size_t nl, const double* lights) {
    auto sc = std::make_unique<Scene>();
    sc->materials.assign(mats, mats + nm);
    for (size_t i = 0; i < np; ++i)
        sc->world_list.add(Plane::make({planes[6 * i], planes[6 * i + 1], planes[6 * i + 2]},
                                       {planes[6 * i + 3], planes[6 * i + 4], planes[6 * i + 5]}, plane_mat[i], (int32_t)i));
    for (size_t i = 0; i < ns; ++i)
        sc->world_list.add(Sphere::make({spheres[4 * i], spheres[4 * i + 1], spheres[4 * i + 2]}, spheres[4 * i + 3],
                                        sphere_mat[i], (int32_t)(np + i)));
    for (size_t i = 0; i < nl; ++i)
        sc->lights.push_back(Sphere::make({lights[4 * i], lights[4 * i + 1], lights[4 * i + 2]}, lights[4 * i + 3], 0, (int32_t)i));
    sc->finalize();
    return sc;
}

}  // namespace orc
