// bvh_build.hpp — host-side binned-SAH BVH2 builder over sphere boxes, flattened for the GPU.
//
// Replaces BoundedVolumeHierarchy::from (shared/src/hittable_collections/bvh.rs:106-143 +
// hittable_list.rs:318-379) for the device.  The topology is NOT the reference's median-start split:
// Hittable::hit returns argmin-t over the primitives whose own tests pass, which does not depend on
// the tree, so the builder is free to optimise for traversal cost.
// Output: inner nodes in breadth-first order (root = 0; the first K nodes are the top levels, which
// the kernels can pin in shared memory), each carrying both child boxes; leaves are ranges of the
// re-ordered sphere array.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <queue>
#include <vector>

namespace rtw {
namespace host {

struct Box {
    double mn[3], mx[3];
    void reset() { for (int a = 0; a < 3; ++a) { mn[a] = std::numeric_limits<double>::infinity(); mx[a] = -mn[a]; } }
    void grow(const Box& b) { for (int a = 0; a < 3; ++a) { mn[a] = std::min(mn[a], b.mn[a]); mx[a] = std::max(mx[a], b.mx[a]); } }
    void grow(const double* p) { for (int a = 0; a < 3; ++a) { mn[a] = std::min(mn[a], p[a]); mx[a] = std::max(mx[a], p[a]); } }
    double area() const {
        double dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
        if (!(dx >= 0 && dy >= 0 && dz >= 0)) return 0.;
        return 2. * (dx * dy + dx * dz + dy * dz);
    }
};

struct FlatNode {           // child >= 0: inner node; child < 0: leaf (first, count)
    Box lbox, rbox;
    int32_t left, right;    // inner index, or -1 for leaf with the range below
    uint32_t lfirst, lcount, rfirst, rcount;
};

struct Bvh {
    std::vector<FlatNode> nodes;        // BFS order
    std::vector<uint32_t> order;        // sorted position -> input sphere index
    uint32_t leaves = 0, depth = 0, max_leaf = 0;
};

class Builder {
public:
    // spheres: n x (cx, cy, cz, r).  Boxes are c -/+ r in f64 exactly as Sphere::new (sphere.rs:42-45).
    Bvh build(const double* spheres, size_t n, int max_leaf = 4, int max_depth = 31) {
        n_ = n; max_leaf_ = std::max(1, std::min(max_leaf, 16)); max_depth_ = max_depth;
        boxes_.resize(n); cent_.resize(3 * n); idx_.resize(n);
        for (size_t i = 0; i < n; ++i) {
            const double* s = spheres + 4 * i;
            for (int a = 0; a < 3; ++a) { boxes_[i].mn[a] = s[a] - s[3]; boxes_[i].mx[a] = s[a] + s[3]; cent_[3 * i + a] = s[a]; }
            idx_[i] = (uint32_t)i;
        }
        return finish();
    }
    // general primitives: n boxes (world-space Bounded::get_aabbox of each entry), centroid = box centre
    Bvh build_boxes(const Box* boxes, size_t n, int max_leaf = 4, int max_depth = 31) {
        n_ = n; max_leaf_ = std::max(1, std::min(max_leaf, 16)); max_depth_ = max_depth;
        boxes_.assign(boxes, boxes + n); cent_.resize(3 * n); idx_.resize(n);
        for (size_t i = 0; i < n; ++i) {
            for (int a = 0; a < 3; ++a) cent_[3 * i + a] = 0.5 * (boxes[i].mn[a] + boxes[i].mx[a]);
            idx_[i] = (uint32_t)i;
        }
        return finish();
    }

private:
    Bvh finish() {
        const size_t n = n_;
        tmp_.clear();
        depth_seen_ = 0;
        Bvh out;
        if (n == 0) {
            FlatNode f{}; f.lbox.reset(); f.rbox.reset(); f.left = f.right = -1; f.lfirst = f.lcount = f.rfirst = f.rcount = 0;
            for (int a = 0; a < 3; ++a) { f.lbox.mn[a] = f.lbox.mx[a] = f.rbox.mn[a] = f.rbox.mx[a] = 0.; }
            out.nodes.push_back(f);
            return out;
        }
        int root = build_range(0, (uint32_t)n, 1);
        // root must be an inner node: wrap a single leaf
        if (tmp_[root].count) {
            TNode w{}; w.box = tmp_[root].box; w.left = root; w.right = -1; w.count = 0; w.first = 0;
            tmp_.push_back(w); root = (int)tmp_.size() - 1;
        }
        // BFS flatten
        std::vector<int> bfs_index(tmp_.size(), -1);
        std::vector<int> queue; queue.push_back(root);
        for (size_t h = 0; h < queue.size(); ++h) {
            int t = queue[h];
            bfs_index[t] = (int)h;
            const TNode& tn = tmp_[t];
            if (tn.left >= 0 && tmp_[tn.left].count == 0) queue.push_back(tn.left);
            if (tn.right >= 0 && tmp_[tn.right].count == 0) queue.push_back(tn.right);
        }
        out.nodes.resize(queue.size());
        for (size_t h = 0; h < queue.size(); ++h) {
            const TNode& tn = tmp_[queue[h]];
            FlatNode f{};
            auto fill = [&](int child, Box& box, int32_t& link, uint32_t& first, uint32_t& count) {
                if (child < 0) { for (int a = 0; a < 3; ++a) box.mn[a] = box.mx[a] = 0.; link = -1; first = 0; count = 0; return; }
                const TNode& c = tmp_[child];
                box = c.box;
                if (c.count) { link = -1; first = c.first; count = c.count; out.leaves++; out.max_leaf = std::max(out.max_leaf, c.count); }
                else { link = bfs_index[child]; first = 0; count = 0; }
            };
            fill(tn.left, f.lbox, f.left, f.lfirst, f.lcount);
            fill(tn.right, f.rbox, f.right, f.rfirst, f.rcount);
            out.nodes[h] = f;
        }
        out.order = idx_;
        out.depth = depth_seen_;
        return out;
    }

    struct TNode { Box box; int left, right; uint32_t first, count; };
    size_t n_ = 0; int max_leaf_ = 4, max_depth_ = 31; uint32_t depth_seen_ = 0;
    std::vector<Box> boxes_; std::vector<double> cent_; std::vector<uint32_t> idx_; std::vector<TNode> tmp_;

    int make_leaf(uint32_t first, uint32_t count, const Box& box, uint32_t depth) {
        TNode t{}; t.box = box; t.left = t.right = -1; t.first = first; t.count = count;
        tmp_.push_back(t);
        depth_seen_ = std::max(depth_seen_, depth);
        return (int)tmp_.size() - 1;
    }

    int build_range(uint32_t first, uint32_t count, uint32_t depth) {
        Box box; box.reset(); Box cb; cb.reset();
        for (uint32_t i = first; i < first + count; ++i) { box.grow(boxes_[idx_[i]]); cb.grow(&cent_[3 * idx_[i]]); }
        if (count == 1) return make_leaf(first, count, box, depth);
        // depth budget: below this many remaining levels only balanced median splits are allowed
        uint32_t remaining = (uint32_t)max_depth_ > depth ? (uint32_t)max_depth_ - depth : 0;
        uint32_t need = 0; { uint32_t c = (count + max_leaf_ - 1) / max_leaf_; while ((1u << need) < c) need++; }
        bool force_median = need + 1 >= remaining;
        if (force_median && (int)count <= max_leaf_) return make_leaf(first, count, box, depth);
        constexpr int NB = 16;
        // cost model: sphere test = 1, inner node (two box tests + stack traffic) = 2.5
        const double c_node = 2.5;
        double best_cost = std::numeric_limits<double>::infinity(); int best_axis = -1, best_bin = -1;
        if (!force_median) {
            for (int a = 0; a < 3; ++a) {
                double lo = cb.mn[a], hi = cb.mx[a];
                if (!(hi > lo)) continue;
                Box bb[NB]; uint32_t bc[NB];
                for (int b = 0; b < NB; ++b) { bb[b].reset(); bc[b] = 0; }
                double k = NB / (hi - lo);
                for (uint32_t i = first; i < first + count; ++i) {
                    int b = std::min(NB - 1, std::max(0, (int)((cent_[3 * idx_[i] + a] - lo) * k)));
                    bb[b].grow(boxes_[idx_[i]]); bc[b]++;
                }
                double ra[NB]; uint32_t rc[NB]; Box acc; acc.reset(); uint32_t cnt = 0;
                for (int b = NB - 1; b > 0; --b) { acc.grow(bb[b]); cnt += bc[b]; ra[b] = acc.area(); rc[b] = cnt; }
                acc.reset(); cnt = 0;
                for (int b = 0; b < NB - 1; ++b) {
                    acc.grow(bb[b]); cnt += bc[b];
                    if (cnt == 0 || rc[b + 1] == 0) continue;
                    double cost = acc.area() * cnt + ra[b + 1] * rc[b + 1];
                    if (cost < best_cost) { best_cost = cost; best_axis = a; best_bin = b; }
                }
            }
        }
        double area = box.area();
        if (!force_median && (int)count <= max_leaf_) {
            double split_cost = best_axis >= 0 && area > 0 ? c_node + best_cost / area : std::numeric_limits<double>::infinity();
            if ((double)count <= split_cost) return make_leaf(first, count, box, depth);
        }
        uint32_t mid;
        if (best_axis >= 0) {
            double lo = cb.mn[best_axis], hi = cb.mx[best_axis], k = NB / (hi - lo);
            auto it = std::partition(idx_.begin() + first, idx_.begin() + first + count, [&](uint32_t id) {
                int b = std::min(NB - 1, std::max(0, (int)((cent_[3 * id + best_axis] - lo) * k)));
                return b <= best_bin;
            });
            mid = (uint32_t)(it - idx_.begin());
        } else {
            // median split on the widest centroid axis (also the depth-budget fallback)
            int a = 0; double w = -1;
            for (int k = 0; k < 3; ++k) if (cb.mx[k] - cb.mn[k] > w) { w = cb.mx[k] - cb.mn[k]; a = k; }
            mid = first + count / 2;
            std::nth_element(idx_.begin() + first, idx_.begin() + mid, idx_.begin() + first + count,
                             [&](uint32_t x, uint32_t y) { return cent_[3 * x + a] < cent_[3 * y + a]; });
            if ((int)count <= max_leaf_ && !(w > 0)) return make_leaf(first, count, box, depth);   // coincident centres
        }
        if (mid == first || mid == first + count) mid = first + count / 2;
        int l = build_range(first, mid - first, depth + 1);
        int r = build_range(mid, first + count - mid, depth + 1);
        TNode t{}; t.box = box; t.left = l; t.right = r; t.first = 0; t.count = 0;
        tmp_.push_back(t);
        return (int)tmp_.size() - 1;
    }
};

}  // namespace host
}  // namespace rtw
