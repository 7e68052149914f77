// bvh.h -- host-side BVH builder over the scene's spheres (binned SAH, two children per node, child boxes
// stored in the parent so one 64-byte node fetch decides both children).
//
// The reference has no acceleration structure (mg_ray_tracer.cpp:62-87 is an O(N) scan); BASELINE.json's
// north_star asks for one "only if the scene size calls for it" (C4: 100 001 spheres).  Traversal must return the
// *same* closest hit as the linear scan, bit for bit, so the device side (kernels.cuh: closest_hit_bvh) tests
// spheres with the same S4 arithmetic, breaks ties by (t, index) and culls boxes conservatively (BVH).
#pragma once
#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <mutex>
#include <new>
#include <thread>
#include <vector>
#ifdef __linux__
#include <sched.h>
#endif

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

// Worker threads for one build.  They are started once and parked on a generation counter between parallel sections: a short
// spin (sections follow each other within microseconds in the lower levels), then a condition variable, so an idle or
// oversubscribed machine is not burnt.  `run(total, chunk, fn)` calls fn(i) for every i in [0, total), handing out `chunk`
// indices at a time, and returns when all calls have returned.
class Pool {
public:
    explicit Pool(unsigned workers)
    {
        try
        {
            threads_.reserve(workers);
            for (unsigned w = 0; w < workers; w++) threads_.emplace_back([this] { park(); });
        }
        catch (...)
        {
            // no more threads to be had (a process limit, no memory for a stack): work with the ones that started
        }
    }
    ~Pool()
    {
        quit_.store(true);
        publish();
        for (auto& th : threads_) th.join();
    }
    Pool(const Pool&) = delete;
    Pool& operator=(const Pool&) = delete;
    unsigned width() const { return (unsigned)threads_.size() + 1; }
    template <class F>
    void run(size_t total, size_t chunk, const F& fn)
    {
        if (threads_.empty() || total <= chunk)
        {
            for (size_t i = 0; i < total; i++) fn(i);
            return;
        }
        struct Thunk { static void call(const void* f, size_t i) { (*static_cast<const F*>(f))(i); } };
        fn_ = &fn; call_ = &Thunk::call; total_ = total; chunk_ = chunk;
        cursor_.store(0);
        parked_.store(0);
        publish();
        drain();
        // the workers are running (or about to): spin, and give the core away if it takes long
        for (unsigned spins = 0; parked_.load(std::memory_order_acquire) != threads_.size();)
        {
            relax();
            if (++spins > SPIN) { std::this_thread::yield(); spins = 0; }
        }
        if (failed_.exchange(false)) throw std::bad_alloc(); // an fn(i) threw on some thread (they only ever allocate)
    }

private:
#ifndef RTCU_BVH_POOL_SPIN
#define RTCU_BVH_POOL_SPIN (1u << 15) // ~1 ms of pause instructions: longer than any serial stretch of a build (a stress test shortens it)
#endif
    static constexpr unsigned SPIN = RTCU_BVH_POOL_SPIN;
    static void relax()
    {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#endif
    }
    void publish()
    {
        generation_.fetch_add(1); // seq_cst with the sleepers_ read below: a worker either sees the new generation or is seen here
        if (sleepers_.load() != 0)
        {
            { std::lock_guard<std::mutex> lock(mutex_); }
            wake_.notify_all();
        }
    }
    void drain() noexcept
    {
        for (size_t i = cursor_.fetch_add(chunk_); i < total_; i = cursor_.fetch_add(chunk_))
            try
            {
                for (size_t j = i, e = std::min(total_, i + chunk_); j < e && !failed_.load(std::memory_order_relaxed); j++) call_(fn_, j);
            }
            catch (...)
            {
                failed_.store(true); // the section is abandoned (remaining indices are skipped); run() reports it on the caller's thread
            }
    }
    void park()
    {
        for (uint64_t seen = 0;;)
        {
            unsigned spins = 0;
            while (generation_.load(std::memory_order_acquire) == seen && ++spins <= SPIN) relax();
            if (generation_.load() == seen)
            {
                std::unique_lock<std::mutex> lock(mutex_);
                sleepers_.fetch_add(1);
                wake_.wait(lock, [&] { return generation_.load() != seen; });
                sleepers_.fetch_sub(1);
            }
            seen++;
            if (quit_.load()) return;
            drain();
            parked_.fetch_add(1, std::memory_order_release);
        }
    }
    std::vector<std::thread> threads_;
    std::mutex mutex_;
    std::condition_variable wake_;
    std::atomic<uint64_t> generation_{ 0 };
    std::atomic<unsigned> sleepers_{ 0 };
    std::atomic<size_t> cursor_{ 0 };
    std::atomic<size_t> parked_{ 0 };
    std::atomic<bool> quit_{ false };
    std::atomic<bool> failed_{ false };
    const void* fn_ = nullptr;
    void (*call_)(const void*, size_t) = nullptr;
    size_t total_ = 0, chunk_ = 1;
};

// threads of a build: the machine's, at most 16 (the passes are memory-bound long before that); RTCU_BVH_THREADS overrides
inline unsigned thread_count()
{
    unsigned hw = std::thread::hardware_concurrency();
#ifdef __linux__
    cpu_set_t allowed; // a process pinned to fewer cores (taskset, a container's cpuset) gets that many
    if (sched_getaffinity(0, sizeof allowed, &allowed) == 0 && CPU_COUNT(&allowed) > 0) hw = std::min(hw ? hw : 1u, (unsigned)CPU_COUNT(&allowed));
#endif
    if (const char* e = std::getenv("RTCU_BVH_THREADS")) hw = (unsigned)std::max(1, std::atoi(e));
    return std::min(std::max(hw, 1u), 16u);
}

// spheres: n x {cx,cy,cz,radius}.  Boxes are rounded outward by one ulp-scale step so that c +- r computed in
// float still encloses the sphere.
//
// The result -- node order, primitive order, every box -- does not depend on the number of threads: boxes and bin counts are
// min / max / integer sums (exact in any order), a split decides set membership only, node indices are assigned serially in
// task order, and a leaf's primitives are sorted by original index (tests compare against RTCU_BVH_THREADS=1).
inline Result build(const float* spheres, uint32_t n, Pool& pool)
{
    Result out;
    const unsigned width = n >= 8192 ? pool.width() : 1u; // small scenes: one thread, the pool is not touched

    std::vector<Box> boxes(n);
    std::vector<float> cent(3 * (size_t)n);
    out.order.resize(n);
    pool.run(((size_t)n + 4095) / 4096, width > 1 ? 1 : SIZE_MAX, [&](size_t blk) {
        for (uint32_t i = (uint32_t)(blk * 4096), e = (uint32_t)std::min<size_t>(n, (blk + 1) * 4096); i < e; i++)
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
            out.order[i] = i;
        }
    });

    struct Task { uint32_t begin, end; int32_t parent; int side; uint32_t depth; };
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

    // Split rules beyond the 16-bin SAH, measured on the device in round 2 (profiles/README.md): ranges of <= sweep_below spheres
    // evaluate every split position exactly instead of binning (a 484-sphere scene is swept whole: C3 -4 %; a large scene only at
    // the bottom of the tree, where 64 is as good as 512 and costs no measurable build time); from 8192 spheres a side's cost counts
    // 4-sphere leaf blocks, ceil(n / 4), instead of spheres, which packs the leaves of a dense field (C4 -2 %; +1 % on C3, hence the
    // threshold).  RTCU_BVH_SWEEP=N and RTCU_BVH_LEAF_COST=0/1 override.
    uint32_t sweep_below = n <= 4096u ? 512u : 64u;
    bool leaf_block_cost = n >= 8192u;
    if (const char* e = std::getenv("RTCU_BVH_SWEEP")) sweep_below = (uint32_t)std::min(4096, std::max(0, std::atoi(e)));
    if (const char* e = std::getenv("RTCU_BVH_LEAF_COST")) leaf_block_cost = std::atoi(e) != 0;
    auto side_cost = [&](float half_area, uint32_t count) { return half_area * (float)(leaf_block_cost ? (count + MAX_LEAF - 1) / MAX_LEAF : count); };

    // binned SAH over the centroid bounds: per axis BINS boxes and counts, then the cheapest of the 3 x (BINS - 1) planes
    struct Bins {
        Box box[3][BINS]; uint32_t n[3][BINS];
        void reset()
        {
            for (int a = 0; a < 3; a++)
                for (int b = 0; b < BINS; b++) { box[a][b].reset(); n[a][b] = 0; }
        }
    };
    auto bin_of = [](float c, float lo, float scale) { return std::min(std::max((int)((c - lo) * scale), 0), BINS - 1); };
    auto fill_bins = [&](Bins& bins, const Box& cb, uint32_t b, uint32_t e) {
        for (int axis = 0; axis < 3; axis++)
        {
            const float ext = cb.hi[axis] - cb.lo[axis];
            if (!(ext > 0.0f)) continue;
            const float scale = BINS / ext;
            for (uint32_t k = b; k < e; k++)
            {
                const uint32_t p = out.order[k];
                const int bin = bin_of(cent[3 * (size_t)p + axis], cb.lo[axis], scale);
                bins.box[axis][bin].grow(boxes[p]);
                bins.n[axis][bin]++;
            }
        }
    };
    auto choose_split = [&](const Bins& bins, const Box& cb, int& best_axis, int& best_bin) {
        best_axis = -1; best_bin = -1;
        float best_cost = FLT_MAX;
        for (int axis = 0; axis < 3; axis++)
        {
            if (!(cb.hi[axis] - cb.lo[axis] > 0.0f)) continue;
            float right_area[BINS]; uint32_t right_n[BINS];
            Box acc; acc.reset(); uint32_t cnt = 0;
            for (int b = BINS - 1; b > 0; b--)
            {
                acc.grow(bins.box[axis][b]); cnt += bins.n[axis][b];
                right_area[b] = cnt ? acc.half_area() : 0.0f; right_n[b] = cnt;
            }
            acc.reset(); cnt = 0;
            for (int b = 0; b < BINS - 1; b++)
            {
                acc.grow(bins.box[axis][b]); cnt += bins.n[axis][b];
                if (cnt == 0 || right_n[b + 1] == 0) continue;
                const float cost = side_cost(acc.half_area(), cnt) + side_cost(right_area[b + 1], right_n[b + 1]);
                if (cost < best_cost) { best_cost = cost; best_axis = axis; best_bin = b; }
            }
        }
    };
    struct Split { Task child[2]; bool has[2]; };
    // after the split: a side of <= MAX_LEAF primitives becomes a leaf of this node, a larger one a task of the next level
    auto finish = [&](const Task& t, const int32_t self, uint32_t mid, Split& sp) {
        if (mid == t.begin || mid == t.end)
        {
            // degenerate (all centroids coincide, or every primitive on one side): split by index order in the middle
            mid = t.begin + (t.end - t.begin) / 2;
            std::sort(out.order.begin() + t.begin, out.order.begin() + t.end);
        }
        const uint32_t ranges[2][2] = { { t.begin, mid }, { mid, t.end } };
        for (int side = 0; side < 2; side++)
        {
            const uint32_t b = ranges[side][0], e = ranges[side][1];
            sp.has[side] = e - b > (uint32_t)MAX_LEAF;
            if (!sp.has[side])
            {
                std::sort(out.order.begin() + b, out.order.begin() + e); // ascending original index inside a leaf
                set_child(self, side, range_box(b, e), ~(int32_t)b, e - b);
            }
            else
                sp.child[side] = Task{ b, e, self, side, t.depth + 1 };
        }
    };

    // one task on one thread (most of the tree: a level has many tasks, which run side by side)
    auto process = [&](const Task& t, const int32_t self, Split& sp) {
        if (t.parent >= 0)
            set_child(t.parent, t.side, range_box(t.begin, t.end), self, 0);
        if (t.end - t.begin <= sweep_below)
        {
            // exact sweep: per axis, order by (centroid, index) -- a total order, so the result does not depend on the order the
            // range arrives in -- and evaluate all count - 1 split positions
            const uint32_t count = t.end - t.begin;
            std::vector<uint32_t> sorted(out.order.begin() + t.begin, out.order.begin() + t.end), best_order;
            std::vector<float> right_area(count);
            float best_cost = FLT_MAX;
            uint32_t best_left = 0;
            for (int axis = 0; axis < 3; axis++)
            {
                std::sort(sorted.begin(), sorted.end(), [&](uint32_t a, uint32_t b) {
                    const float ca = cent[3 * (size_t)a + axis], cb = cent[3 * (size_t)b + axis];
                    return ca < cb || (ca == cb && a < b);
                });
                Box acc; acc.reset();
                for (uint32_t k = count - 1; k > 0; k--) { acc.grow(boxes[sorted[k]]); right_area[k] = acc.half_area(); }
                acc.reset();
                for (uint32_t k = 0; k + 1 < count; k++)
                {
                    acc.grow(boxes[sorted[k]]);
                    const float cost = side_cost(acc.half_area(), k + 1) + side_cost(right_area[k + 1], count - k - 1);
                    if (cost < best_cost) { best_cost = cost; best_left = k + 1; best_order = sorted; }
                }
            }
            if (best_left) std::copy(best_order.begin(), best_order.end(), out.order.begin() + t.begin);
            finish(t, self, t.begin + best_left, sp);
            return;
        }
        Box cb; cb.reset();
        for (uint32_t k = t.begin; k < t.end; k++) cb.grow_point(&cent[3 * (size_t)out.order[k]]);
        Bins bins; bins.reset();
        fill_bins(bins, cb, t.begin, t.end);
        int best_axis, best_bin;
        choose_split(bins, cb, best_axis, best_bin);
        uint32_t mid = t.begin; // stays there when all centroids coincide
        if (best_axis >= 0)
        {
            const float lo = cb.lo[best_axis], scale = BINS / (cb.hi[best_axis] - cb.lo[best_axis]);
            auto it = std::partition(out.order.begin() + t.begin, out.order.begin() + t.end,
                                     [&](uint32_t p) { return bin_of(cent[3 * (size_t)p + best_axis], lo, scale) <= best_bin; });
            mid = (uint32_t)(it - out.order.begin());
        }
        finish(t, self, mid, sp);
    };

    // one task on all threads (the first levels, where there are fewer tasks than threads): the range is cut into slices, every
    // pass runs slice-parallel and the per-slice boxes / bins / counts are merged in slice order
    std::vector<uint32_t> scratch;
    auto process_wide = [&](const Task& t, const int32_t self, Split& sp) {
        const uint32_t count = t.end - t.begin;
        const uint32_t slices = std::min<uint32_t>(4 * width, (count + 1023) / 1024);
        auto slice = [&](size_t s) { return t.begin + (uint32_t)((uint64_t)count * s / slices); };
        std::vector<Box> part_box(slices), part_cb(slices);
        pool.run(slices, 1, [&](size_t s) {
            Box bb, cb; bb.reset(); cb.reset();
            for (uint32_t k = slice(s), e = slice(s + 1); k < e; k++)
            {
                const uint32_t p = out.order[k];
                bb.grow(boxes[p]); cb.grow_point(&cent[3 * (size_t)p]);
            }
            part_box[s] = bb; part_cb[s] = cb;
        });
        Box bb, cb; bb.reset(); cb.reset();
        for (uint32_t s = 0; s < slices; s++) { bb.grow(part_box[s]); cb.grow(part_cb[s]); }
        if (t.parent >= 0)
            set_child(t.parent, t.side, bb, self, 0);
        std::vector<Bins> part_bins(slices);
        pool.run(slices, 1, [&](size_t s) { part_bins[s].reset(); fill_bins(part_bins[s], cb, slice(s), slice(s + 1)); });
        Bins& bins = part_bins[0];
        for (uint32_t s = 1; s < slices; s++)
            for (int a = 0; a < 3; a++)
                for (int b = 0; b < BINS; b++) { bins.box[a][b].grow(part_bins[s].box[a][b]); bins.n[a][b] += part_bins[s].n[a][b]; }
        int best_axis, best_bin;
        choose_split(bins, cb, best_axis, best_bin);
        uint32_t mid = t.begin;
        if (best_axis >= 0)
        {
            // stable partition through a scratch copy: count per slice, exclusive scan, scatter, copy back
            const float lo = cb.lo[best_axis], scale = BINS / (cb.hi[best_axis] - cb.lo[best_axis]);
            auto left = [&](uint32_t p) { return bin_of(cent[3 * (size_t)p + best_axis], lo, scale) <= best_bin; };
            std::vector<uint32_t> n_left(slices + 1, 0);
            pool.run(slices, 1, [&](size_t s) {
                uint32_t c = 0;
                for (uint32_t k = slice(s), e = slice(s + 1); k < e; k++) c += left(out.order[k]) ? 1u : 0u;
                n_left[s + 1] = c;
            });
            for (uint32_t s = 0; s < slices; s++) n_left[s + 1] += n_left[s];
            mid = t.begin + n_left[slices];
            if (scratch.size() < n) scratch.resize(n);
            pool.run(slices, 1, [&](size_t s) {
                uint32_t l = t.begin + n_left[s], r = mid + (slice(s) - t.begin - n_left[s]);
                for (uint32_t k = slice(s), e = slice(s + 1); k < e; k++)
                {
                    const uint32_t p = out.order[k];
                    if (left(p)) scratch[l++] = p; else scratch[r++] = p;
                }
            });
            pool.run(slices, 1, [&](size_t s) { std::copy(scratch.begin() + slice(s), scratch.begin() + slice(s + 1), out.order.begin() + slice(s)); });
        }
        finish(t, self, mid, sp);
    };

    // Level-synchronous construction: the tasks of one level own disjoint primitive ranges and write disjoint fields (their own
    // node, one side of their parent).
    std::vector<Task> level{ Task{ 0, n, -1, 0, 1 } }, next_level;
    std::vector<Split> splits;
    std::vector<int32_t> self_of;
    while (!level.empty())
    {
        out.max_depth = std::max(out.max_depth, level[0].depth);
        // node indices in task order (the root task reuses node 0)
        self_of.assign(level.size(), 0);
        for (size_t i = 0; i < level.size(); i++)
            if (level[i].parent >= 0)
            {
                self_of[i] = (int32_t)out.nodes.size();
                out.nodes.push_back(Node{});
            }
        splits.assign(level.size(), Split{});
        uint32_t prims = 0;
        for (const Task& t : level) prims += t.end - t.begin;
        if (width == 1 || prims < 4096)
            for (size_t i = 0; i < level.size(); i++) process(level[i], self_of[i], splits[i]);
        else if (level.size() < 2 * (size_t)width)
            for (size_t i = 0; i < level.size(); i++)
                if (level[i].end - level[i].begin >= 8192) process_wide(level[i], self_of[i], splits[i]);
                else process(level[i], self_of[i], splits[i]);
        else
            pool.run(level.size(), std::min<size_t>(64, std::max<size_t>(1, level.size() / (16 * (size_t)width))),
                     [&](size_t i) { process(level[i], self_of[i], splits[i]); });
        next_level.clear();
        for (const Split& sp : splits)
            for (int side = 0; side < 2; side++)
                if (sp.has[side]) next_level.push_back(sp.child[side]);
        level.swap(next_level);
    }
    return out;
}

inline Result build(const float* spheres, uint32_t n)
{
    Pool pool(n >= 8192 ? thread_count() - 1 : 0);
    return build(spheres, n, pool);
}

} // namespace rtcu_bvh
