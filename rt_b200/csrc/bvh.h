// bvh.h -- host-side BVH builder over the scene's spheres (binned SAH, two children per node, child boxes
// stored in the parent so one 64-byte node fetch decides both children).
//
// The reference has no acceleration structure (mg_ray_tracer.cpp:62-87 is an O(N) scan); BASELINE.json's
// north_star asks for one "only if the scene size calls for it" (C4: 100 001 spheres).  Traversal must return the
// *same* closest hit as the linear scan, bit for bit, so the device side (kernels.cuh: closest_hit_bvh) tests
// spheres with the same S4 arithmetic, breaks ties by (t, index) and culls boxes conservatively (BVH).
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <vector>

namespace rtcu_bvh {

// 64-byte node.  Child c (0 = left, 1 = right) has box [lo[c], hi[c]]; child[c] >= 0 is an inner node index,
// child[c] < 0 encodes a leaf: first primitive = ~child[c] in the reordered arrays, count[c] primitives (0 = empty).
struct Node {
    float x[4]; // l.lo.x, l.hi.x, r.lo.x, r.hi.x
    float y[4];
    float z[4];
    int32_t child[2];
    uint32_t count[2];
};
static_assert(sizeof(Node) == 64, "node must be 4 x float4");

struct Box {
    float lo[3], hi[3];
    void reset()
    {
        for (int k = 0; k < 3; k++) { lo[k] = FLT_MAX; hi[k] = -FLT_MAX; }
    }
    void grow(const Box& b)
    {
        for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], b.lo[k]); hi[k] = std::max(hi[k], b.hi[k]); }
    }
    void grow_point(const float* p)
    {
        for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], p[k]); hi[k] = std::max(hi[k], p[k]); }
    }
    float half_area() const
    {
        const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        return dx * dy + dy * dz + dz * dx;
    }
};

struct Result {
    std::vector<Node> nodes;       // nodes[0] is the root; top levels are in breadth-first order
    std::vector<uint32_t> order;   // order[k] = original sphere index of reordered primitive k
    uint32_t max_depth = 0;
};

constexpr int MAX_LEAF = 4;
constexpr int BINS = 16;

// spheres: n x {cx,cy,cz,radius}.  Boxes are rounded outward by one ulp-scale step so that c +- r computed in
// float still encloses the sphere.
inline Result build(const float* spheres, uint32_t n)
{
    Result out;
    std::vector<Box> boxes(n);
    std::vector<float> cent(3 * (size_t)n);
    for (uint32_t i = 0; i < n; i++)
    {
        const float* s = spheres + 4 * (size_t)i;
        const float r = std::fabs(s[3]);
        for (int k = 0; k < 3; k++)
        {
            const float pad = (std::fabs(s[k]) + r) * 2.4e-7f; // 4 ulp of the larger magnitude
            boxes[i].lo[k] = s[k] - r - pad;
            boxes[i].hi[k] = s[k] + r + pad;
            cent[3 * (size_t)i + k] = s[k];
        }
    }
    out.order.resize(n);
    for (uint32_t i = 0; i < n; i++) out.order[i] = i;

    struct Task { uint32_t begin, end; int32_t parent; int side; uint32_t depth; };
    // breadth-first construction: node indices of the top levels are contiguous (shared-memory cache on the device)
    std::vector<Task> queue;
    auto range_box = [&](uint32_t b, uint32_t e) {
        Box bb; bb.reset();
        for (uint32_t k = b; k < e; k++) bb.grow(boxes[out.order[k]]);
        return bb;
    };
    auto set_child = [&](int32_t parent, int side, const Box& bb, int32_t child, uint32_t count) {
        Node& p = out.nodes[parent];
        p.x[2 * side] = bb.lo[0]; p.x[2 * side + 1] = bb.hi[0];
        p.y[2 * side] = bb.lo[1]; p.y[2 * side + 1] = bb.hi[1];
        p.z[2 * side] = bb.lo[2]; p.z[2 * side + 1] = bb.hi[2];
        p.child[side] = child;
        p.count[side] = count;
    };
    const Box empty = { { FLT_MAX, FLT_MAX, FLT_MAX }, { -FLT_MAX, -FLT_MAX, -FLT_MAX } };

    // root is always an inner node; a scene of <= MAX_LEAF spheres becomes {leaf(all), empty}
    out.nodes.push_back(Node{});
    if (n <= (uint32_t)MAX_LEAF)
    {
        set_child(0, 0, n ? range_box(0, n) : empty, ~0, n);
        set_child(0, 1, empty, ~0, 0);
        out.max_depth = 1;
        return out;
    }
    queue.push_back({ 0, n, -1, 0, 1 });
    for (size_t qi = 0; qi < queue.size(); qi++)
    {
        const Task t = queue[qi];
        out.max_depth = std::max(out.max_depth, t.depth);
        const uint32_t count = t.end - t.begin;
        // this task becomes an inner node (the root task reuses node 0)
        int32_t self = 0;
        if (t.parent >= 0)
        {
            self = (int32_t)out.nodes.size();
            out.nodes.push_back(Node{});
            set_child(t.parent, t.side, range_box(t.begin, t.end), self, 0);
        }
        // choose the split: binned SAH over the centroid bounds, best of the three axes
        Box cb; cb.reset();
        for (uint32_t k = t.begin; k < t.end; k++) cb.grow_point(&cent[3 * (size_t)out.order[k]]);
        int best_axis = -1, best_bin = -1;
        float best_cost = FLT_MAX;
        for (int axis = 0; axis < 3; axis++)
        {
            const float ext = cb.hi[axis] - cb.lo[axis];
            if (!(ext > 0.0f)) continue;
            Box bin_box[BINS]; uint32_t bin_n[BINS] = {};
            for (auto& b : bin_box) b.reset();
            const float scale = BINS / ext;
            for (uint32_t k = t.begin; k < t.end; k++)
            {
                const uint32_t p = out.order[k];
                int b = (int)((cent[3 * (size_t)p + axis] - cb.lo[axis]) * scale);
                b = std::min(std::max(b, 0), BINS - 1);
                bin_box[b].grow(boxes[p]);
                bin_n[b]++;
            }
            float right_area[BINS]; uint32_t right_n[BINS];
            Box acc; acc.reset(); uint32_t cnt = 0;
            for (int b = BINS - 1; b > 0; b--)
            {
                acc.grow(bin_box[b]); cnt += bin_n[b];
                right_area[b] = cnt ? acc.half_area() : 0.0f; right_n[b] = cnt;
            }
            acc.reset(); cnt = 0;
            for (int b = 0; b < BINS - 1; b++)
            {
                acc.grow(bin_box[b]); cnt += bin_n[b];
                if (cnt == 0 || right_n[b + 1] == 0) continue;
                const float cost = acc.half_area() * cnt + right_area[b + 1] * right_n[b + 1];
                if (cost < best_cost) { best_cost = cost; best_axis = axis; best_bin = b; }
            }
        }
        uint32_t mid;
        if (best_axis >= 0)
        {
            const float lo = cb.lo[best_axis], scale = BINS / (cb.hi[best_axis] - cb.lo[best_axis]);
            auto it = std::partition(out.order.begin() + t.begin, out.order.begin() + t.end, [&](uint32_t p) {
                int b = (int)((cent[3 * (size_t)p + best_axis] - lo) * scale);
                b = std::min(std::max(b, 0), BINS - 1);
                return b <= best_bin;
            });
            mid = (uint32_t)(it - out.order.begin());
        }
        else
            mid = t.begin; // all centroids coincide
        if (mid == t.begin || mid == t.end)
        {
            // degenerate: split by index order in the middle (keeps leaves <= MAX_LEAF)
            mid = t.begin + count / 2;
            std::sort(out.order.begin() + t.begin, out.order.begin() + t.end);
        }
        const uint32_t ranges[2][2] = { { t.begin, mid }, { mid, t.end } };
        for (int side = 0; side < 2; side++)
        {
            const uint32_t b = ranges[side][0], e = ranges[side][1];
            if (e - b <= (uint32_t)MAX_LEAF)
            {
                std::sort(out.order.begin() + b, out.order.begin() + e); // ascending original index inside a leaf
                set_child(self, side, range_box(b, e), ~(int32_t)b, e - b);
            }
            else
                queue.push_back({ b, e, self, side, t.depth + 1 });
        }
    }
    return out;
}

} // namespace rtcu_bvh
