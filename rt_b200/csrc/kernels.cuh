// kernels.cuh -- sm_100a kernels of the path-tracing hot path.
//
//   k_render_mega<STAGE,FLAT,BVH>   register-resident paths: one thread owns one pixel and regenerates the next sample of
//                     the same pixel when a path ends (mg_ray_tracer.cpp:182-201 per pixel, :154-174 per path, iterated
//                     instead of recursed).  Linear scenes: primitives staged in shared memory as packed pairs, read as
//                     warp-uniform LDS.128, two spheres per FFMA2.  BVH scenes: closest_hit_bvh (exact, conservative culls).
//   k_render_stragglers<BVH>        second pass: one warp per pixel that exceeded its per-thread segment budget
//   k_resolve / k_reduce_resolve / k_reduce_resolve_rows   sum / spp, sqrt, clamp, pack (mg_ray_tracer.cpp:195-200); the
//                                   latter two fused with the cross-GPU sum over NVLink peer loads (one process / one per GPU)
//   k_intersect_batch / k_primary_rays / k_scatter_batch / k_philox_batch   step-wise parity kernels
//   k_fp32_peak                     FFMA / FFMA2 calibration stream for the roofline denominator
//   k_selftest_math                 exhaustive comparison of spec.cuh's cheaper sqrt / rcp / div with the IEEE intrinsics
// (wavefront.cuh: the queue-based pipeline; pool.cuh: the warp-local ray pool; both selectable, neither the default)
#pragma once
#include "spec.cuh"

namespace rtcu_dev {

struct SceneDev {
    const float4* spheres;   // {cx,cy,cz,r*r}
    const float4* pairs;     // pair_float4_count(n) float4 (padded with never-hit pairs): {cx0,cx1,cy0,cy1},{cz0,cz1,r2_0,r2_1} (packed-FP32 scan layout)
    const uint32_t* sphere_material;
    uint32_t n_spheres;
    const float4* planes;    // {nx,ny,nz,d}
    const uint32_t* plane_material;
    uint32_t n_planes;
    const MatRec* materials;
    uint32_t n_materials;
    // BVH over the spheres (built when the scene size calls for it: bvh.h builds a binary SAH tree, rtcu.cu collapses it to
    // 4-wide nodes).  8 float4 per node: children (0,1) as centre / half-extent {c0, c1, h0, h1} for x, y, z (interleaved
    // for the packed FP32 slab test; h rounded outward, -inf for an empty slot), the same for children (2,3), the four child
    // references as bit patterns, and the four H = h.x + h.y + h.z.
    // ref >= 0: inner node index; ref < 0: leaf = 0x80000000 | leaf number; leaf L is the 80-byte block leaf_blk[5L..5L+4]:
    // two packed pairs {cx0,cx1,cy0,cy1},{cz0,cz1,r2_0,r2_1} (padded with never-hit spheres) and, as bit patterns, the four
    // original sphere indices (0x7fffffff = padding) -- one address computation, five 16-byte loads issued together.
    const float4* bvh_nodes;
    const float4* leaf_blk;
    uint32_t n_bvh_nodes;
};

struct RenderParams {
    CameraConst cam;
    uint32_t width, height;
    uint32_t tile_x0, tile_y0, tile_x1, tile_y1;
    uint32_t sample_begin, sample_end;
    uint32_t max_bounces;
    uint32_t mode;
    PhiloxKeys rk;         // Philox round keys of the view's seed
    float spp_resolve;     // float(samples_per_pixel)
    int accumulate;
    float4* accum;         // width*height {sum_r,sum_g,sum_b,n}
    uint32_t* rgba8;       // nullable
    unsigned long long* counters; // [0] = segments
    uint32_t regen_threshold;     // regenerate ended lanes once this many lanes of the warp are idle (1 = immediately)
    // straggler hand-off: a thread that has traced `segment_budget` segments stops at the next sample boundary and
    // queues {pixel, next sample} for k_render_stragglers (one warp per queued pixel).  0 = no budget.
    uint32_t segment_budget;
    uint2* stragglers;            // queue storage, width*height entries
    unsigned int* straggler_count;
    // longest-tile-first scheduling (see launch_render): every frame records the segments of each tile in tile_cost; the next
    // frame of the same view takes its tiles from tile_order (tile ids by descending cost of the previous frame)
    uint32_t* tile_cost;          // += segments traced for the CTA's tile (nullable)
    const uint32_t* tile_order;   // CTA b renders tile tile_order[b] (nullable: CTA b renders tile b)
    int direct;                   // != 0: k_render_stragglers renders every pixel of the tile itself (no first pass, no queue)
    const struct BeamList* beam;  // non-null (direct mode): per 8x4 patch, the leaves its primary rays can hit (k_beam_lists)
    float* run_scratch;           // k_render_runs: per lane group, RUN x 3 x G partial sums (lanes' shares of the run's pixels)
};

#ifndef RTCU_PRIM_MISS
#define RTCU_PRIM_MISS 0xFFFFFFFFu
#define RTCU_PRIM_PLANE 0x80000000u
#endif

struct Hit { float t; uint32_t prim; }; // prim: sphere index | PLANE|index | MISS

// packed-scan layout sizes: pairs padded to an even count, plus two trailing never-hit pairs (prefetch target)
__host__ __device__ __forceinline__ uint32_t pair_count_padded(uint32_t n_spheres) { return ((n_spheres + 3u) >> 2) << 1; }
__host__ __device__ __forceinline__ uint32_t pair_float4_count(uint32_t n_spheres) { return 2u * (pair_count_padded(n_spheres) + 2u); }

// closest hit over planes then spheres with the reference's tie rules
// (mg_ray_tracer.cpp:35-102, :160-162).  s_pairs / s_pl may point to shared or global memory.
// NP > 0: the scene has at most 2 NP spheres (the reference's own scenes: basic.toml 3, dielectric.toml 7) and the pair array,
// padded with never-hit pairs, is swept as exactly NP pairs -- no loop counter, no remainder loop, no prefetch bookkeeping: the
// generic loop spends 38 instructions per segment on those (tools/ncu_phases.py on C2), as many as two pair tests.
template <int NP = 0>
__device__ __forceinline__ Hit closest_hit_linear(const float4* __restrict__ s_pairs, const uint32_t n_sph,
                                                  const float4* __restrict__ s_pl, const uint32_t n_pl, const Ray& r)
{
    // "nothing found yet" is a NaN: `!(best <= t)` then accepts the first candidate whatever its distance -- +inf included, which
    // the reference's `have && hit_dist <= t` (mg_ray_tracer.cpp:74) accepts too (a sphere whose r * r overflows) -- and the final
    // `distance >= 0` test reads it as a miss
    const float none = __int_as_float(0x7fc00000);
    float ts = none;
    int is = -1;
    // software pipeline: the next pair is loaded (warp-uniform LDS.128 x2) before the current one is tested, so the
    // shared-memory latency hides behind ~20 arithmetic instructions.  The array ends with never-hit pairs (r2 = -inf),
    // so the prefetch of the pair after the last one is always in bounds.  (Testing two pairs per iteration with a
    // two-pair prefetch was measured slower: 16 more live registers under the 64-register cap.)
    if (NP > 0)
    {
        // two pairs per iteration of a rolled loop: fully unrolled, ptxas issues all 2 NP loads up front and spills 200 bytes around
        // them under the 64-register cap (C2 2.65 -> 3.03 ms); one pair per iteration with a prefetch spills 108 bytes
#pragma unroll 1
        for (int j = 0; j < NP; j += 2)
        {
            sphere_pair_test(s_pairs[2 * j], s_pairs[2 * j + 1], j, r, ts, is);
            sphere_pair_test(s_pairs[2 * j + 2], s_pairs[2 * j + 3], j + 1, r, ts, is);
        }
    }
    else
    {
        const uint32_t n_pairs = (n_sph + 1u) >> 1;
        float4 A = s_pairs[0], B = s_pairs[1];
#pragma unroll 4
        for (uint32_t j = 0; j < n_pairs; j++)
        {
            const float4 An = s_pairs[2 * j + 2], Bn = s_pairs[2 * j + 3];
            sphere_pair_test(A, B, (int)j, r, ts, is);
            A = An;
            B = Bn;
        }
    }
    // A result counts as a hit when its distance is >= 0 (hit_result::operator bool, mg_ray_tracer.cpp:29-32): every accepted
    // distance is >= 0.001 except a NaN -- S4 overflowing to inf - inf on coordinates around 1e19 and beyond -- which the scan's
    // `best <= t` rule lets through and this rule then reports as a miss.  (is = -1 converts to RTCU_PRIM_MISS by itself.)
    Hit h;
    h.t = ts;
    h.prim = ts >= 0.0f ? (uint32_t)is : RTCU_PRIM_MISS;
    if (n_pl)
    {
        float tp = none;
        int ip = -1;
        for (uint32_t i = 0; i < n_pl; i++)
            plane_test(s_pl[i], (int)i, r, tp, ip);
        // select(spheres, planes): the sphere wins when a.distance <= b.distance (:95-102)
        if (ip >= 0 && tp >= 0.0f && !(h.prim != RTCU_PRIM_MISS && ts <= tp))
        {
            h.t = tp;
            h.prim = RTCU_PRIM_PLANE | (uint32_t)ip;
        }
    }
    if (h.prim == RTCU_PRIM_MISS)
        h.t = -1.0f;
    return h;
}

// hit normal (mg_ray_tracer.cpp:56-59, :84-86); the sphere centre is read back from the pair layout
__device__ __forceinline__ V3 hit_normal(const float4* __restrict__ s_pairs, const float4* __restrict__ s_pl, const Ray& r, const Hit h)
{
    if (h.prim & RTCU_PRIM_PLANE)
    {
        const float4 pl = s_pl[h.prim & 0x7FFFFFFFu];
        return v3(pl.x, pl.y, pl.z);
    }
    const float* base = reinterpret_cast<const float*>(s_pairs) + 8u * (h.prim >> 1) + (h.prim & 1u);
    return normalize3(v3_sub(ray_at(r.o, r.d, h.t), v3(base[0], base[2], base[4])));
}

// ---- BVH traversal with the linear scan's exact result ------------------------------------------------------
// The answer must equal closest_hit_linear bit for bit: the lexicographic minimum (t, index) over all spheres whose
// S4 test yields t >= 0.001.  Spheres are tested with the same S4 arithmetic, so only *culling* can change the result;
// it is made conservative against the rounding of S4 itself.  With u = 2^-24: e = c - o carries u|e| per component, e2 and
// a = e.d carry <= 5u e2 and <= 4u|e||d|, a^2 <= 8u e2, the fused e2 - a^2 one more u e2, so the computed discriminant
// differs from the geometric one by Delta <= (16u + 2|d.d - 1|) |e|^2 =: kappa^2 |e|^2.  A computed hit therefore means the
// line passes within sqrt(r^2 + Delta) <= r + kappa|e| of the centre, and the computed t is >= (entry of the ray into that
// inflated ball) - 4u|e|.  Every child box is therefore inflated by m = kappa * E, E = L1 distance from the ray origin to the
// box's farthest corner (>= |e| of any sphere inside), before the slab test and the `tmin <= best_t` ordering cull.
// (The tighter m = sqrt(r_min^2 + kappa^2 E^2) - r_min with the subtree's smallest radius was measured: 13-15 % fewer
// sphere tests but 2-7 % slower -- the extra per-node arithmetic costs more than the visits it saves.)
struct BvhStats { uint32_t nodes, tests; };

// two spheres of a leaf on the packed FP32 pipe (same arithmetic as sphere_pair_test, hence as the scan); candidates are
// accepted by (t, index) lexicographic order because leaves are visited in traversal order, not index order
// ANY_T (the rasterizer, rasterizer.cpp:48): no minimum distance, negative distances included.
template <bool ANY_T = false>
__device__ __forceinline__ void bvh_leaf_candidate(const float a, const float e2, const float r2, const float disc, const int index,
                                                   float& best_t, int& best_i)
{
    const float f = __fsqrt_rn(disc);
    const float t = (e2 < r2) ? __fadd_rn(a, f) : __fsub_rn(a, f);
    if ((ANY_T || !(t < 0.001f)) && (t < best_t || (t == best_t && index < best_i)))
    {
        best_t = t;
        best_i = index;
    }
}

template <bool ANY_T = false>
__device__ __forceinline__ void bvh_leaf_pair_test(const float4 A, const float4 B, const int i0, const int i1, const Ray& r, float& best_t, int& best_i)
{
    const float2 ex = __fadd2_rn(make_float2(A.x, A.y), make_float2(-r.o.x, -r.o.x));
    const float2 ey = __fadd2_rn(make_float2(A.z, A.w), make_float2(-r.o.y, -r.o.y));
    const float2 ez = __fadd2_rn(make_float2(B.x, B.y), make_float2(-r.o.z, -r.o.z));
    const float2 e2 = __ffma2_rn(ez, ez, __ffma2_rn(ey, ey, __fmul2_rn(ex, ex)));
    const float2 a = __ffma2_rn(ez, make_float2(r.d.z, r.d.z), __ffma2_rn(ey, make_float2(r.d.y, r.d.y), __fmul2_rn(ex, make_float2(r.d.x, r.d.x))));
    const float2 t1 = __ffma2_rn(make_float2(-a.x, -a.y), a, e2);
    const float2 disc = __fadd2_rn(make_float2(B.z, B.w), make_float2(-t1.x, -t1.y));
    if (!(disc.x < 0.0f) || !(disc.y < 0.0f))
    {
        if (!(disc.x < 0.0f))
            bvh_leaf_candidate<ANY_T>(a.x, e2.x, B.z, disc.x, i0, best_t, best_i);
        if (!(disc.y < 0.0f))
            bvh_leaf_candidate<ANY_T>(a.y, e2.y, B.w, disc.y, i1, best_t, best_i);
    }
}

// A leaf block pads its 1-4 spheres to two packed pairs, padding last (pack_bvh4 sorts a leaf by original index and fills up with
// never-hit spheres): when the third slot is padding the second pair holds nothing and its test can be skipped.  Every scene
// with one sphere far larger than the rest -- a ground sphere -- has such a leaf right under the root, and nearly every ray visits it
// (CPU proxy tools/bvh_visits.cpp: 46 % of the leaf visits of C3 / C5, 34 % of C4's).  RTCU_LEAF_SKIP_PAD: bit 0 = skip in the
// traversal, bit 1 = skip in the beam-list scan of the primary rays (where the branch is uniform over the lanes of a pixel).
#ifndef RTCU_LEAF_SKIP_PAD
#define RTCU_LEAF_SKIP_PAD 3 // measured (B200): C3 30.46 -> 29.78 ms, C4 80.45 -> 79.65 ms, a 32-sample slice of C5 32.85 -> 32.45 ms
#endif
__device__ __forceinline__ bool leaf_second_pair(const float4 idx, const bool skip_padding)
{
    return !skip_padding || __float_as_int(idx.z) != 0x7fffffff;
}

constexpr int BVH_STACK = 64; // entries; a visit pushes at most 3, so trees up to 20 levels of 4-wide nodes (4^20 leaves)

// traversal state of one ray; the node stack lives in (L1-cached) local memory of the calling kernel
struct Trav {
    uint32_t node;
    int sp;
    float best_t;
    int best_i;
    float ix, iy, iz; // 1 / d
    float kappa;      // margin factor, see above
    float madd;       // margin added on top of kappa * E: 0 for a ray, the origin spread for a pixel beam (beam_collect)
};

// returns false when the ray direction is too far from unit length for the conservative margins, or the origin so far out
// (beyond 2^62; NaN included) that S4 could overflow (caller falls back to the linear scan)
__device__ __forceinline__ bool trav_init(const Ray& r, Trav& tv)
{
    tv.node = 0;
    tv.sp = 0;
    tv.best_t = __int_as_float(0x7f800000);
    tv.best_i = 0x7fffffff;
    const float eps_d = fabsf(dot3(r.d, r.d) - 1.0f);
    // kappa and 1 / d only feed the slab tests, whose margins carry 1 % of slack: MUFU approximations (2 ulp; the square root as
    // x * rsqrt(x)) cost one instruction each where the IEEE forms expand to ten with a slow-path branch.  1.015 absorbs their
    // error; a zero or denormal component gives +-inf like the exact reciprocal (the axis then does not cull)
    const float k2 = 16.0f * 5.9604645e-8f + 2.0f * eps_d;
    tv.kappa = 1.015f * k2 * mufu_rsq(k2);
    tv.madd = 0.0f;
    tv.ix = mufu_rcp(r.d.x);
    tv.iy = mufu_rcp(r.d.y);
    tv.iz = mufu_rcp(r.d.z);
    return eps_d <= 1e-3f && fabsf(r.o.x) <= 0x1p62f && fabsf(r.o.y) <= 0x1p62f && fabsf(r.o.z) <= 0x1p62f; // (false for a NaN)
}

// slab test of two child boxes at once on the packed FP32 pipe.  cx/cy/cz = {c_a, c_b, h_a, h_b} per axis (centre and
// outward-rounded half-extent), H = h.x + h.y + h.z per child: E = |c - o|_1 + H bounds the L1 distance from the origin to the
// box's farthest corner, the inflated slab is tc -+ (h + m)|1/d| with tc = (c - o)/d and m = kappa E, and near / far come out
// already ordered: 4 FADD2 + 3 FMUL2 + 9 FFMA2 + 4 FADD + 4 FMNMX3 for two boxes.  Rounding is irrelevant here (far below the
// 1 % slack in kappa).  A zero direction component gives inf - inf = NaN on that axis, which min / max drop: the axis then
// does not cull (conservative).  An empty slot has h = H = -inf: near = +inf, far = -inf, never hit.
template <bool ANY_T>
__device__ __forceinline__ void slab_pair(const float4 cx, const float4 cy, const float4 cz, const float Ha, const float Hb, const Ray& r,
                                          const Trav& tv, float& tn_a, float& tn_b, bool& hit_a, bool& hit_b)
{
    const float2 dcx = __fadd2_rn(make_float2(cx.x, cx.y), make_float2(-r.o.x, -r.o.x));
    const float2 dcy = __fadd2_rn(make_float2(cy.x, cy.y), make_float2(-r.o.y, -r.o.y));
    const float2 dcz = __fadd2_rn(make_float2(cz.x, cz.y), make_float2(-r.o.z, -r.o.z));
    const float2 e3 = make_float2(fabsf(dcx.x) + fabsf(dcy.x) + fabsf(dcz.x), fabsf(dcx.y) + fabsf(dcy.y) + fabsf(dcz.y));
    const float2 m = __ffma2_rn(__fadd2_rn(e3, make_float2(Ha, Hb)), make_float2(tv.kappa, tv.kappa), make_float2(tv.madd, tv.madd));
    const float2 hx = __fadd2_rn(make_float2(cx.z, cx.w), m), hy = __fadd2_rn(make_float2(cy.z, cy.w), m), hz = __fadd2_rn(make_float2(cz.z, cz.w), m);
    const float2 tcx = __fmul2_rn(dcx, make_float2(tv.ix, tv.ix)), tcy = __fmul2_rn(dcy, make_float2(tv.iy, tv.iy)), tcz = __fmul2_rn(dcz, make_float2(tv.iz, tv.iz));
    const float2 ax = make_float2(fabsf(tv.ix), fabsf(tv.ix)), ay = make_float2(fabsf(tv.iy), fabsf(tv.iy)), az = make_float2(fabsf(tv.iz), fabsf(tv.iz));
    const float2 n_x = __ffma2_rn(make_float2(-hx.x, -hx.y), ax, tcx), f_x = __ffma2_rn(hx, ax, tcx);
    const float2 n_y = __ffma2_rn(make_float2(-hy.x, -hy.y), ay, tcy), f_y = __ffma2_rn(hy, ay, tcy);
    const float2 n_z = __ffma2_rn(make_float2(-hz.x, -hz.y), az, tcz), f_z = __ffma2_rn(hz, az, tcz);
    tn_a = fmaxf(fmaxf(n_x.x, n_y.x), n_z.x);
    tn_b = fmaxf(fmaxf(n_x.y, n_y.y), n_z.y);
    const float tf_a = fminf(fminf(f_x.x, f_y.x), f_z.x), tf_b = fminf(fminf(f_x.y, f_y.y), f_z.y);
    hit_a = tf_a >= (ANY_T ? tn_a : fmaxf(tn_a, 0.0f)) && tn_a <= tv.best_t;
    hit_b = tf_b >= (ANY_T ? tn_b : fmaxf(tn_b, 0.0f)) && tn_b <= tv.best_t;
}

// one visit of a 4-wide node: slab-test the four children, test leaf children immediately, continue with the nearest inner
// child, push the others (in slot order) / pop.  Returns true when the traversal is complete.  ANY_T: the whole line counts,
// not only t >= 0 (the margin argument above never uses the sign of t: a computed hit lies within the inflated ball, whose
// line interval lies within the inflated box).
template <bool ANY_T = false>
__device__ __forceinline__ bool trav_step(const SceneDev& sc, const Ray& r, Trav& tv, uint32_t* __restrict__ stack_ref,
                                          float* __restrict__ stack_t, BvhStats& st)
{
    st.nodes++;
    const float4* np = sc.bvh_nodes + 8u * tv.node;
    const float4 refs = __ldg(np + 6), hs = __ldg(np + 7);
    const uint32_t ref[4] = { __float_as_uint(refs.x), __float_as_uint(refs.y), __float_as_uint(refs.z), __float_as_uint(refs.w) };
    float tn[4];
    bool hit[4];
    slab_pair<ANY_T>(__ldg(np), __ldg(np + 1), __ldg(np + 2), hs.x, hs.y, r, tv, tn[0], tn[1], hit[0], hit[1]);
    slab_pair<ANY_T>(__ldg(np + 3), __ldg(np + 4), __ldg(np + 5), hs.z, hs.w, r, tv, tn[2], tn[3], hit[2], hit[3]);
    // leaves are tested immediately; of the inner children the nearest is descended, the others are pushed
    uint32_t next = 0xffffffffu;
    float next_t = 0.0f;
#pragma unroll
    for (int c = 0; c < 4; c++)
    {
        if (!hit[c]) continue;
        if (ref[c] & 0x80000000u)
        {
            const uint32_t leaf = ref[c] & 0x7fffffffu;
            const float4* lp = sc.leaf_blk + 5u * leaf; // leaf < 2^29 (checked at upload): 32-bit index arithmetic
            const float4 a0 = __ldg(lp), b0 = __ldg(lp + 1), a1 = __ldg(lp + 2), b1 = __ldg(lp + 3), idx = __ldg(lp + 4);
            bvh_leaf_pair_test<ANY_T>(a0, b0, __float_as_int(idx.x), __float_as_int(idx.y), r, tv.best_t, tv.best_i);
            if (leaf_second_pair(idx, (RTCU_LEAF_SKIP_PAD & 1) != 0)) // sphere test slots: a leaf holds 1-4 spheres
            {
                st.tests += 4;
                bvh_leaf_pair_test<ANY_T>(a1, b1, __float_as_int(idx.z), __float_as_int(idx.w), r, tv.best_t, tv.best_i);
            }
            else
                st.tests += 2;
        }
        else if (next == 0xffffffffu)
        {
            next = ref[c];
            next_t = tn[c];
        }
        else
        {
            // another inner child: keep the nearer as the one to descend, push the farther
            const bool swap = tn[c] < next_t;
            stack_ref[tv.sp] = swap ? next : ref[c];
            stack_t[tv.sp] = swap ? next_t : tn[c];
            tv.sp++;
            if (swap) { next = ref[c]; next_t = tn[c]; }
        }
    }
    if (next != 0xffffffffu && next_t <= tv.best_t)
    {
        tv.node = next;
        return false;
    }
    // pop, skipping entries that the current best already excludes
    while (tv.sp > 0)
    {
        tv.sp--;
        if (stack_t[tv.sp] <= tv.best_t)
        {
            tv.node = stack_ref[tv.sp];
            return false;
        }
    }
    return true;
}

// whole traversal for one ray; (best_t, best_i) = closest sphere hit or (inf, -1)
__device__ __forceinline__ bool closest_sphere_bvh(const SceneDev& sc, const Ray& r, float& best_t, int& best_i, BvhStats& st)
{
    Trav tv;
    if (!trav_init(r, tv))
        return false;
    uint32_t stack_ref[BVH_STACK];
    float stack_t[BVH_STACK];
    while (!trav_step(sc, r, tv, stack_ref, stack_t, st))
    {
    }
    best_t = tv.best_t;
    best_i = tv.best_i == 0x7fffffff ? -1 : tv.best_i;
    return true;
}

// ---- traversal 2: node visits and leaf tests as separate, warp-converged phases ("if-if") -------------------------------------
// The same tree, margins, culls and (t, index) rule as trav_step -- hence the same exact result -- with a different control
// structure, chosen because the kernels are issue-bound at 19-23 of 32 active lanes and a quarter of their instructions were
// branches:
//   * a leaf child is no longer tested inside the node visit that found it (up to four divergent sub-branches per visit): it is
//     ordered with the inner children -- the nearest hit child of any kind is processed next, the others go on the stack with
//     their entry distance -- so the pop-time cull `tn <= best_t` now also drops leaves that a nearer hit has made irrelevant;
//   * one loop iteration is { node visit, if the current item is an inner node } { leaf test, if it is (now) a leaf } { pop }:
//     every phase is entered by all lanes that need it at the same time, a lane may pass through all three in one iteration;
//   * the child bookkeeping is branch-free: the nearest child by three compare / selects, the other three by unconditional
//     8-byte stack stores whose stack-pointer increments are predicated.
// TOP > 0: the first TOP nodes of the breadth-first array (the top levels of the tree, which every ray visits) are read from
// the copy the CTA staged in shared memory.
struct TopNodes { const float4* nodes; uint32_t count; };

template <bool USE_TOP, bool ANY_T = false>
__device__ __forceinline__ bool closest_sphere_bvh2(const SceneDev& sc, const TopNodes top, const Ray& r, float& best_t, int& best_i, BvhStats& st)
{
    Trav tv;
    if (!trav_init(r, tv))
        return false;
    uint2 stack[BVH_STACK + 1]; // {entry distance (bits), child reference}; one spare slot for the unconditional stores
    int sp = 0;
    const uint32_t NONE = 0xffffffffu;
    const float inf = __int_as_float(0x7f800000);
    uint32_t cur = 0; // the root; (int)cur >= 0: inner node, < -1: leaf, == -1: nothing in hand
    for (;;)
    {
        if ((int)cur >= 0)
        {
            st.nodes++;
            float4 nd[8];
            if (USE_TOP && cur < top.count)
            {
                const float4* np = top.nodes + 8u * cur; // shared memory
#pragma unroll
                for (int k = 0; k < 8; k++) nd[k] = np[k];
            }
            else
            {
                const float4* np = sc.bvh_nodes + 8u * cur;
#pragma unroll
                for (int k = 0; k < 8; k++) nd[k] = __ldg(np + k);
            }
            const uint32_t ref[4] = { __float_as_uint(nd[6].x), __float_as_uint(nd[6].y), __float_as_uint(nd[6].z), __float_as_uint(nd[6].w) };
            float tn[4];
            bool hit[4];
            slab_pair<ANY_T>(nd[0], nd[1], nd[2], nd[7].x, nd[7].y, r, tv, tn[0], tn[1], hit[0], hit[1]);
            slab_pair<ANY_T>(nd[3], nd[4], nd[5], nd[7].z, nd[7].w, r, tv, tn[2], tn[3], hit[2], hit[3]);
            // the nearest hit child is processed next (ties: the lower slot) ...
            float nt = inf;
            uint32_t nref = NONE;
            int nslot = -1;
#pragma unroll
            for (int c = 0; c < 4; c++)
            {
                const bool nearer = hit[c] && (nslot < 0 || tn[c] < nt);
                nt = nearer ? tn[c] : nt;
                nref = nearer ? ref[c] : nref;
                nslot = nearer ? c : nslot;
            }
            // ... the other hit children go on the stack in slot order: store always, keep the slot only when it counts
#pragma unroll
            for (int c = 0; c < 4; c++)
            {
                stack[sp] = make_uint2(__float_as_uint(tn[c]), ref[c]);
                sp += (hit[c] && c != nslot) ? 1 : 0;
            }
            cur = nref;
        }
        if ((int)cur < -1)
        {
            const uint32_t leaf = cur & 0x7fffffffu;
            const float4* lp = sc.leaf_blk + 5u * leaf; // leaf < 2^29 (checked at upload): 32-bit index arithmetic
            const float4 a0 = __ldg(lp), b0 = __ldg(lp + 1), a1 = __ldg(lp + 2), b1 = __ldg(lp + 3), idx = __ldg(lp + 4);
            st.tests += 4; // sphere test slots (a leaf holds 1-4 spheres)
            bvh_leaf_pair_test<ANY_T>(a0, b0, __float_as_int(idx.x), __float_as_int(idx.y), r, tv.best_t, tv.best_i);
            bvh_leaf_pair_test<ANY_T>(a1, b1, __float_as_int(idx.z), __float_as_int(idx.w), r, tv.best_t, tv.best_i);
            cur = NONE;
        }
        if (cur == NONE)
        {
            // pop, skipping entries that the current best already excludes
            bool found = false;
            while (sp > 0)
            {
                const uint2 e = stack[--sp];
                if (__uint_as_float(e.x) <= tv.best_t)
                {
                    cur = e.y;
                    found = true;
                    break;
                }
            }
            if (!found)
                break;
        }
    }
    best_t = tv.best_t;
    best_i = tv.best_i == 0x7fffffff ? -1 : tv.best_i;
    return true;
}

// ---- patch beams: the primary rays of an 8x4-pixel patch share their candidate leaves ----------------------------------------
// More than half of all path segments are primary rays, and the reference's sample loop (mg_ray_tracer.cpp:187-194) sends all
// of a pixel's samples through the same pixel: their rays differ by at most the pixel's footprint, and a handful of neighbouring
// pixels by little more.  Before a frame of a BVH scene is traced, k_beam_lists therefore walks the tree ONCE per 8x4-pixel
// patch with the patch's centre ray and margins widened by the patch's footprint, WITHOUT the closest-hit cull, and records
// every leaf that any primary ray of the patch could hit (up to BEAM_MAX, sorted by entry distance).  A sample's primary ray
// then tests just those leaves -- S4, the (t, index) rule, exactly as a traversal would -- instead of descending from the
// root: the same result for a third of the instructions, and all lanes of a pixel run the same short loop.  Patches whose beam
// touches more leaves (grazing views over many spheres) keep the traversal.  (Measured: the median patch of C4 lists 4 leaves,
// 86 % of them at most 16; C3 / C5: median 2, none above 16.  One walk per patch serves 32 pixels x all their samples, so it
// pays even at C4's 64 samples per pixel, where a walk per pixel -- the first form of this -- did not.)
//
// Why the list is complete.  Let R0 = (o0, d0) be the centre ray and R' = (o', d') any primary ray of the patch, with
// |o' - o0| <= rho and |d' - d0| <= sigma (both maximal at the patch's corners: screen -> near / far points is affine when the
// viewport's perspective divide is constant, and the angle to d0 is quasi-convex over the far-minus-near quad; the corners are
// evaluated with the same arithmetic as the samples, plus slack for its rounding).  If S4 reports a hit of R' on a sphere
// (c, r) at t' >= 0, the line of R' passes within r + kappa'|e'| of c (see above), and R0(t') lies within rho + sigma t' of
// R'(t').  With E = the L1 distance from o0 to the far corner of a box around the sphere: |e'| <= E + rho and t' <= |e'| +
// r + kappa'|e'| <= 2.01 E + 1.01 rho, so R0 meets the box inflated by (kappa_b + 2.01 sigma) E + 2 rho at parameter t' --
// i.e. the slab test of the ordinary traversal with kappa := kappa_b + 2.01 sigma and an additive 2 rho passes for every
// ancestor box of the sphere, and its entry distance tn0 <= t' (which makes `tn0 > best_t` a valid reason to stop early).
// kappa_b is the margin factor of a ray that is unit length to within BEAM_EPS_D; a sample ray outside that (never seen:
// primary directions are normalised) traverses.  tests/test_bvh_replay.py replays this on the CPU against the oracle's scan.
#ifndef RTCU_BEAM_MAX
#define RTCU_BEAM_MAX 32 // capacity 8 / 16 / 24 / 32: C4 82.7 / 80.5 / 79.7 / 79.2 ms, C3 and C5 unchanged (no list above 16 there)
#endif
constexpr int BEAM_MAX = RTCU_BEAM_MAX;             // leaves per list
constexpr int BEAM_MAX_VISITS = 6 * BEAM_MAX;       // node visits after which a beam is given up
constexpr float BEAM_EPS_D = 16.0f * 5.9604645e-8f; // |d.d - 1| bound of the rays that may use a list
constexpr uint32_t BEAM_PATCH_W = 8, BEAM_PATCH_H = 4; // the patches k_render_stragglers enumerates in direct mode
struct BeamEntry { float tn; uint32_t leaf; };
struct BeamList { int n; int pad; BeamEntry e[BEAM_MAX]; }; // n < 0: no list, traverse

__device__ __forceinline__ void beam_collect(const float4* __restrict__ bvh_nodes, const Ray& r0, const float sigma, const float rho, BeamList* __restrict__ out)
{
    out->n = -1;
    Trav tv;
    if (!trav_init(r0, tv) || !(fabsf(dot3(r0.d, r0.d) - 1.0f) <= BEAM_EPS_D))
        return;
    tv.kappa = 1.01f * (1.01f * sqrtf(16.0f * 5.9604645e-8f + 2.0f * BEAM_EPS_D) + 2.01f * sigma);
    tv.madd = 2.02f * rho;
    uint32_t stack[BVH_STACK];
    BeamEntry list[BEAM_MAX];
    int sp = 0, n = 0, visits = 0;
    uint32_t node = 0;
    for (;;)
    {
        if (++visits > BEAM_MAX_VISITS) return;
        const float4* np = bvh_nodes + 8u * node;
        const float4 refs = __ldg(np + 6), hs = __ldg(np + 7);
        const uint32_t ref[4] = { __float_as_uint(refs.x), __float_as_uint(refs.y), __float_as_uint(refs.z), __float_as_uint(refs.w) };
        float tn[4];
        bool hit[4];
        slab_pair<false>(__ldg(np), __ldg(np + 1), __ldg(np + 2), hs.x, hs.y, r0, tv, tn[0], tn[1], hit[0], hit[1]); // (tv.best_t stays +inf: no cull)
        slab_pair<false>(__ldg(np + 3), __ldg(np + 4), __ldg(np + 5), hs.z, hs.w, r0, tv, tn[2], tn[3], hit[2], hit[3]);
        for (int c = 0; c < 4; c++)
        {
            if (!hit[c]) continue;
            if (ref[c] & 0x80000000u)
            {
                if (n == BEAM_MAX) return;
                int k = n++;
                for (; k > 0 && list[k - 1].tn > tn[c]; k--) // keep the list sorted by entry distance
                    list[k] = list[k - 1];
                list[k].tn = tn[c];
                list[k].leaf = ref[c] & 0x7fffffffu;
            }
            else
                stack[sp++] = ref[c];
        }
        if (sp == 0) break;
        node = stack[--sp];
    }
    for (int k = 0; k < n; k++) out->e[k] = list[k];
    out->n = n;
}

// closest sphere hit of a primary ray of the beam's patch: the (t, index) minimum over the listed leaves (read from global memory:
// the loads are uniform across the lanes of a pixel)
__device__ __forceinline__ void beam_closest_sphere(const SceneDev& sc, const BeamList* __restrict__ beam, const int n, const Ray& r, float& best_t, int& best_i,
                                                    BvhStats& st)
{
    best_t = __int_as_float(0x7f800000);
    best_i = 0x7fffffff;
    for (int k = 0; k < n; k++)
    {
        const uint2 e = __ldg(reinterpret_cast<const uint2*>(&beam->e[k]));
        if (__uint_as_float(e.x) > best_t) break; // no ray of the patch reaches this leaf (or a later one) before tn
        const float4* lp = sc.leaf_blk + 5u * e.y;
        const float4 a0 = __ldg(lp), b0 = __ldg(lp + 1), a1 = __ldg(lp + 2), b1 = __ldg(lp + 3), idx = __ldg(lp + 4);
        bvh_leaf_pair_test<false>(a0, b0, __float_as_int(idx.x), __float_as_int(idx.y), r, best_t, best_i);
        if (leaf_second_pair(idx, (RTCU_LEAF_SKIP_PAD & 2) != 0))
        {
            st.tests += 4;
            bvh_leaf_pair_test<false>(a1, b1, __float_as_int(idx.z), __float_as_int(idx.w), r, best_t, best_i);
        }
        else
            st.tests += 2;
    }
    best_i = best_i == 0x7fffffff ? -1 : best_i;
}

// sphere result of a traversal + the planes (linear) -> Hit, with the reference's select rule (mg_ray_tracer.cpp:95-102)
// The BVH kernels are issue-bound and their hot loop (generate, traverse, leaf tests, shade) is ~22 KB of SASS against a 32 KB
// L1.5 instruction cache (profiles/README.md: a variant whose loop grew to 40 KB spent 8x the cycles waiting for instructions).
// Code that a BVH scene rarely runs is therefore kept OUT of the loop body as real functions: the plane loop (the large scenes
// have no planes) and the scan that non-unit or far-away rays fall back to.
struct PlaneHit { float t; int i; };
__device__ __noinline__ PlaneHit closest_plane_cold(const float4* __restrict__ s_pl, const uint32_t n_pl, const Ray r)
{
    PlaneHit h;
    h.t = __int_as_float(0x7fc00000); // NaN = nothing found yet, see closest_hit_linear
    h.i = -1;
    for (uint32_t i = 0; i < n_pl; i++)
        plane_test(s_pl[i], (int)i, r, h.t, h.i);
    return h;
}
// (arguments by value: a reference to the kernel's SceneDev / RenderParams parameter would make the compiler keep a copy of the
// whole parameter in local memory and read it from there everywhere)
__device__ __noinline__ Hit closest_hit_scan_cold(const float4* __restrict__ pairs, const uint32_t n_spheres, const float4* __restrict__ s_pl, const uint32_t n_planes,
                                                  const Ray r)
{
    return closest_hit_linear(pairs, n_spheres, s_pl, n_planes, r);
}

__device__ __forceinline__ Hit combine_with_planes(const SceneDev& sc, const float4* __restrict__ s_pl, const Ray& r, const float ts, const int is)
{
    Hit h;
    h.t = ts;
    h.prim = is >= 0 ? (uint32_t)is : RTCU_PRIM_MISS; // (a traversal never yields a NaN distance: trav_init sends such rays to the scan)
    if (sc.n_planes)
    {
        const PlaneHit ph = closest_plane_cold(s_pl, sc.n_planes, r);
        const float tp = ph.t;
        const int ip = ph.i;
        if (ip >= 0 && tp >= 0.0f && !(is >= 0 && ts <= tp))
        {
            h.t = tp;
            h.prim = RTCU_PRIM_PLANE | (uint32_t)ip;
        }
    }
    if (h.prim == RTCU_PRIM_MISS)
        h.t = -1.0f;
    return h;
}

// closest hit: spheres through the BVH (or the global-memory linear scan when the ray is not unit length), planes linear
// TRAV: 0 = trav_step (leaf children tested inside the node visit), 1 = closest_sphere_bvh2 (deferred leaves, if-if phases)
//       2 = the same with the top levels of the tree read from shared memory (`top`)
template <int TRAV = 1>
__device__ __forceinline__ Hit closest_hit_bvh(const SceneDev& sc, const float4* __restrict__ s_pl, const Ray& r, BvhStats& st, const TopNodes top = TopNodes{ nullptr, 0u })
{
    float ts;
    int is;
    if (!(TRAV == 0 ? closest_sphere_bvh(sc, r, ts, is, st) : closest_sphere_bvh2<TRAV == 2>(sc, top, r, ts, is, st)))
    {
        st.tests += sc.n_spheres;
        return closest_hit_scan_cold(sc.pairs, sc.n_spheres, s_pl, sc.n_planes, r);
    }
    return combine_with_planes(sc, s_pl, r, ts, is);
}

// hit normal when primitives are not staged: centre from the global {cx,cy,cz,r*r} array
__device__ __forceinline__ V3 hit_normal_global(const SceneDev& sc, const float4* __restrict__ s_pl, const Ray& r, const Hit h)
{
    if (h.prim & RTCU_PRIM_PLANE)
    {
        const float4 pl = s_pl[h.prim & 0x7FFFFFFFu];
        return v3(pl.x, pl.y, pl.z);
    }
    const float4 sp = __ldg(sc.spheres + h.prim);
    return normalize3(v3_sub(ray_at(r.o, r.d, h.t), v3(sp.x, sp.y, sp.z)));
}

__device__ __forceinline__ uint32_t hit_material(const SceneDev& sc, const Hit h)
{
    return (h.prim & RTCU_PRIM_PLANE) ? __ldg(sc.plane_material + (h.prim & 0x7FFFFFFFu)) : __ldg(sc.sphere_material + h.prim);
}

__device__ __forceinline__ MatRec load_material(const SceneDev& sc, uint32_t m)
{
    const float4* p = reinterpret_cast<const float4*>(sc.materials + m);
    const float4 a = __ldg(p), b = __ldg(p + 1);
    MatRec r;
    r.att_r = a.x; r.att_g = a.y; r.att_b = a.z; r.roughness = a.w;
    r.ior = b.x; r.type = __float_as_uint(b.y); r.inv_ior = b.z; r.r0 = b.w;
    return r;
}

// jittered screen position of (pixel, sample): mg_ray_tracer.cpp:189
__device__ __forceinline__ Ray generate(const CameraConst& cam, const RngKey& key, uint32_t px, uint32_t py)
{
    float jx = 0.5f, jy = 0.5f;
    if (key.sample != 0u)
    {
        const uint4 b = rng_block(key, 0u, 0u);
        jx = u01(b.x);
        jy = u01(b.y);
    }
    return primary_ray(cam, __fadd_rn(__uint2float_rn(px), jx), __fadd_rn(__uint2float_rn(py), jy));
}

#ifndef RTCU_MEGA_TILE_H
#define RTCU_MEGA_TILE_H 8 // 16 -> 256 threads, 8 -> 128 threads
#endif
#ifndef RTCU_MEGA_MIN_BLOCKS
#define RTCU_MEGA_MIN_BLOCKS 8
#endif
constexpr int MEGA_TILE_W = 16, MEGA_TILE_H = RTCU_MEGA_TILE_H, MEGA_THREADS = MEGA_TILE_W * MEGA_TILE_H;

template <bool BVH>
__device__ __forceinline__ bool shade_segment(const SceneDev& sc, const RenderParams& p, const float4* __restrict__ s_sph,
                                              const float4* __restrict__ s_pl, const RngKey& key, Ray& ray, V3& thr, V3& sum, uint32_t& seg,
                                              const Hit h);

// One path segment (mg_ray_tracer.cpp:154-174, one level of the recursion): closest hit, then sky on a miss or one
// scatter event on a hit.  Returns true when the path ended (miss, absorbed, or bounce budget exhausted).
template <bool BVH, int TRAV = 1, int NP = 0>
__device__ __forceinline__ bool segment_step(const SceneDev& sc, const RenderParams& p, const float4* __restrict__ s_sph,
                                             const float4* __restrict__ s_pl, const RngKey& key, Ray& ray, V3& thr, V3& sum, uint32_t& seg,
                                             BvhStats& bst, const TopNodes top = TopNodes{ nullptr, 0u })
{
    const Hit h = BVH ? closest_hit_bvh<TRAV>(sc, s_pl, ray, bst, top) : closest_hit_linear<NP>(s_sph, sc.n_spheres, s_pl, sc.n_planes, ray);
    return shade_segment<BVH>(sc, p, s_sph, s_pl, key, ray, thr, sum, seg, h);
}

// the part of a segment after the closest hit is known: sky on a miss, one scatter event on a hit
template <bool BVH>
__device__ __forceinline__ bool shade_segment(const SceneDev& sc, const RenderParams& p, const float4* __restrict__ s_sph,
                                              const float4* __restrict__ s_pl, const RngKey& key, Ray& ray, V3& thr, V3& sum, uint32_t& seg,
                                              const Hit h)
{
    if (h.prim == RTCU_PRIM_MISS)
    {
        sum = v3_add(sum, v3_mul(thr, sky(ray.d))); // S12: iterative throughput
        return true;
    }
    const V3 n = BVH ? hit_normal_global(sc, s_pl, ray, h) : hit_normal(s_sph, s_pl, ray, h);
    const MatRec m = load_material(sc, hit_material(sc, h));
    const uint4 rnd = rng_block(key, seg + 1u, 0u);
    Ray next;
    const bool scattered = scatter(scatter_kind(p.mode, m.type), m, ray, h.t, n, key, seg + 1u, rnd, next);
    thr = v3_mul(thr, v3(m.att_r, m.att_g, m.att_b));
    ray = next;
    seg++;
    // absorbed (:173) or bounce budget exhausted (:157-158): radiance 0
    return !scattered || seg >= p.max_bounces;
}

// STAGE: primitives staged in dynamic shared memory (true) or read through L1 from global (false).
// FLAT: lanes regenerate inside a warp-vote loop (all lanes reconverge every segment; best when the O(N) sweep
// dominates).  !FLAT: plain per-thread loop, which the compiler nests as {generate; bounce until every lane's path
// ended} -- generate and shade then run at full lane occupancy, best for tiny N where they dominate.
// BVH: spheres are reached through the BVH (STAGE must be false; the structure lives in L1/L2).
template <bool STAGE, bool FLAT, bool BVH, int TRAV = 1>
__global__ void __launch_bounds__(MEGA_THREADS, RTCU_MEGA_MIN_BLOCKS) k_render_mega(const SceneDev sc, const RenderParams p)
{
    extern __shared__ float4 smem[];
    const float4* s_sph = sc.pairs;
    const float4* s_pl = sc.planes;
    if (STAGE)
    {
        const uint32_t n4 = pair_float4_count(sc.n_spheres);
        for (uint32_t i = threadIdx.x; i < n4; i += MEGA_THREADS)
            smem[i] = __ldg(sc.pairs + i);
        for (uint32_t i = threadIdx.x; i < sc.n_planes; i += MEGA_THREADS)
            smem[n4 + i] = __ldg(sc.planes + i);
        __syncthreads();
        s_sph = smem;
        s_pl = smem + n4;
    }

    // a warp covers an 8x4 pixel patch; the block a 16 x MEGA_TILE_H tile
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t tile_col = blockIdx.x, tile_row = blockIdx.y;
    const uint32_t tile_id = p.tile_order ? __ldg(p.tile_order + blockIdx.y * gridDim.x + blockIdx.x) : blockIdx.y * gridDim.x + blockIdx.x;
    if (p.tile_order)
    {
        tile_col = tile_id % gridDim.x;
        tile_row = tile_id / gridDim.x;
    }
    const uint32_t px = p.tile_x0 + tile_col * MEGA_TILE_W + (warp & 1u) * 8u + (lane & 7u);
    const uint32_t py = p.tile_y0 + tile_row * MEGA_TILE_H + (warp >> 1) * 4u + (lane >> 3);
    const bool in_tile = px < p.tile_x1 && py < p.tile_y1;

    // (measured on C3, 484 spheres: the plain loop keeps 23/32 lanes active in the sweep, the vote loop ~29/32;
    //  on C2, 7 spheres, the plain loop is ~10 % faster because generate/shade dominate)
    unsigned long long segs = 0;
    BvhStats bst;
    bst.nodes = 0;
    bst.tests = 0;
    RngKey key;
    key.ks = &p.rk;
    key.pixel = py * p.width + px;
    key.sample = p.sample_begin;
    V3 sum = v3(0.0f, 0.0f, 0.0f);
    V3 thr = v3(1.0f, 1.0f, 1.0f);
    uint32_t seg = 0;
    if (FLAT)
    {
        // lane state: `more` = the pixel still has samples to start, `live` = a path is in flight
        bool more = in_tile && p.sample_begin < p.sample_end;
        bool live = false;
        Ray ray;
        ray.o = v3(0.0f, 0.0f, 0.0f);
        ray.d = v3(0.0f, 0.0f, 1.0f);
        // the candidate leaves of the warp's 8x4 patch (k_beam_lists; the warp's patch IS one of its patches): n < 0 = traverse
        const BeamList* beam = nullptr;
        int beam_n = -1;
        if (BVH && p.beam && in_tile)
        {
            const uint32_t patches_x = (p.tile_x1 - p.tile_x0 + BEAM_PATCH_W - 1u) / BEAM_PATCH_W;
            beam = p.beam + ((py - p.tile_y0) / BEAM_PATCH_H) * patches_x + (px - p.tile_x0) / BEAM_PATCH_W;
            beam_n = __ldg(&beam->n);
        }
        for (;;)
        {
            // (1) regeneration, decided by a warp vote so every lane reconverges here each iteration: lanes whose path
            // ended start the next sample of their pixel once `regen_threshold` lanes are idle (or nothing is in flight)
            const unsigned idle = __ballot_sync(0xffffffffu, more && !live);
            const unsigned flying = __ballot_sync(0xffffffffu, live);
            if (!idle && !flying)
                break;
            bool fresh = false;
            if (idle && (!flying || __popc(idle) >= (int)p.regen_threshold))
            {
                if (more && !live)
                {
                    seg = 0;
                    thr = v3(1.0f, 1.0f, 1.0f);
                    ray = generate(p.cam, key, px, py);
                    live = true;
                    fresh = true;
                }
            }
            // (2) one path segment for every lane in flight -- two for a lane that has just started a sample in a patch with a
            // beam list: its primary hit comes from the list (pass 0), then it traverses with the others (pass 1).  One rolled
            // copy of the segment code serves both passes (instruction-cache footprint, see closest_plane_cold).
            const bool listed = BVH && fresh && beam_n >= 0 && fabsf(dot3(ray.d, ray.d) - 1.0f) <= BEAM_EPS_D;
#pragma unroll 1
            for (int pass = (BVH && __any_sync(0xffffffffu, listed)) ? 0 : 1; pass < 2; pass++)
            {
                if (pass == 0 ? listed : live)
                {
                    segs++;
                    Hit h;
                    if (BVH && pass == 0)
                    {
                        float ts;
                        int is;
                        beam_closest_sphere(sc, beam, beam_n, ray, ts, is, bst);
                        h = combine_with_planes(sc, s_pl, ray, ts, is);
                    }
                    else
                        h = BVH ? closest_hit_bvh<TRAV>(sc, s_pl, ray, bst) : closest_hit_linear(s_sph, sc.n_spheres, s_pl, sc.n_planes, ray);
                    if (shade_segment<BVH>(sc, p, s_sph, s_pl, key, ray, thr, sum, seg, h))
                    {
                        live = false;
                        key.sample++;
                        more = key.sample < p.sample_end;
                        if (more && p.segment_budget && segs >= p.segment_budget)
                        {
                            p.stragglers[atomicAdd(p.straggler_count, 1u)] = make_uint2(key.pixel, key.sample);
                            more = false;
                        }
                    }
                }
            }
        }
    }
    else if (in_tile && p.sample_begin < p.sample_end)
    {
        Ray ray = generate(p.cam, key, px, py);
        for (;;)
        {
            segs++;
            if (segment_step<BVH, TRAV>(sc, p, s_sph, s_pl, key, ray, thr, sum, seg, bst))
            {
                key.sample++;
                if (key.sample >= p.sample_end)
                    break;
                if (p.segment_budget && segs >= p.segment_budget)
                {
                    p.stragglers[atomicAdd(p.straggler_count, 1u)] = make_uint2(key.pixel, key.sample);
                    break;
                }
                seg = 0;
                thr = v3(1.0f, 1.0f, 1.0f);
                ray = generate(p.cam, key, px, py);
            }
        }
    }

    if (in_tile)
    {
        const size_t idx = (size_t)(py * p.width + px);
        // n = samples finished by this thread (key.sample has advanced past them); stragglers add the rest
        float4 acc = make_float4(sum.x, sum.y, sum.z, (float)(key.sample - p.sample_begin));
        if (p.accumulate)
        {
            const float4 old = p.accum[idx];
            acc.x = __fadd_rn(old.x, acc.x); acc.y = __fadd_rn(old.y, acc.y); acc.z = __fadd_rn(old.z, acc.z);
            acc.w = __fadd_rn(old.w, acc.w);
        }
        p.accum[idx] = acc;
        if (p.rgba8)
            p.rgba8[idx] = pack_pixel(acc.x, acc.y, acc.z, p.spp_resolve);
    }

    // exact counters: warp reduce, one atomic per warp
    unsigned long long nodes = bst.nodes, tests = bst.tests;
    for (int off = 16; off > 0; off >>= 1)
    {
        segs += __shfl_down_sync(0xffffffffu, segs, off);
        if (BVH)
        {
            nodes += __shfl_down_sync(0xffffffffu, nodes, off);
            tests += __shfl_down_sync(0xffffffffu, tests, off);
        }
    }
    if (lane == 0 && segs)
    {
        if (p.tile_cost)
            atomicAdd(p.tile_cost + tile_id, (uint32_t)segs);
        atomicAdd(p.counters, segs);
        if (BVH)
        {
            atomicAdd(p.counters + 1, nodes);
            atomicAdd(p.counters + 2, tests);
        }
    }
}

// Tile ids by descending cost: a stable counting sort over 64 linear cost classes (class 0 = the most expensive), the
// row-major order kept inside a class so that neighbouring CTAs still work on neighbouring tiles (BVH nodes and leaves stay
// warm in L1/L2).  One CTA of 32 warps; every warp owns a contiguous chunk of tiles, counts them per class, and after one
// block-wide exclusive scan over (class, warp) places them with match_any ranks -- deterministic.
constexpr int TILE_CLASSES = 64;
__device__ __forceinline__ uint32_t tile_class(uint32_t cost, uint32_t mx)
{
    return (uint32_t)(TILE_CLASSES - 1) - (uint32_t)((unsigned long long)min(cost, mx) * (TILE_CLASSES - 1) / mx);
}
__global__ void __launch_bounds__(1024) k_tile_order(const uint32_t* __restrict__ cost, uint32_t n, uint32_t* __restrict__ order)
{
    __shared__ uint32_t s_max;
    __shared__ uint32_t s_cnt[TILE_CLASSES * 32]; // [class][warp]
    __shared__ uint32_t s_part[32];
    const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    if (t == 0) s_max = 1u;
    for (uint32_t i = t; i < TILE_CLASSES * 32; i += 1024u) s_cnt[i] = 0u;
    __syncthreads();
    uint32_t m = 0;
    for (uint32_t i = t; i < n; i += 1024u) m = max(m, cost[i]);
    for (int off = 16; off > 0; off >>= 1) m = max(m, __shfl_down_sync(0xffffffffu, m, off));
    if (lane == 0) atomicMax(&s_max, m);
    __syncthreads();
    const uint32_t mx = s_max;
    const uint32_t chunk = ((n + 31u) / 32u + 31u) & ~31u; // tiles per warp, a multiple of 32
    const uint32_t lo = min(n, warp * chunk), hi = min(n, lo + chunk);
    for (uint32_t i = lo + lane; i < hi; i += 32u)
        atomicAdd(&s_cnt[tile_class(cost[i], mx) * 32u + warp], 1u);
    __syncthreads();
    // exclusive scan of the 2048 counters in (class, warp) order: two per thread
    const uint32_t a = s_cnt[2 * t], b = s_cnt[2 * t + 1];
    uint32_t v = a + b;
    for (int off = 1; off < 32; off <<= 1)
    {
        const uint32_t u = __shfl_up_sync(0xffffffffu, v, off);
        if (lane >= (uint32_t)off) v += u;
    }
    if (lane == 31) s_part[warp] = v;
    __syncthreads();
    if (warp == 0)
    {
        const uint32_t w = s_part[lane];
        uint32_t x = w;
        for (int off = 1; off < 32; off <<= 1)
        {
            const uint32_t u = __shfl_up_sync(0xffffffffu, x, off);
            if (lane >= (uint32_t)off) x += u;
        }
        s_part[lane] = x - w;
    }
    __syncthreads();
    const uint32_t excl = s_part[warp] + v - (a + b);
    s_cnt[2 * t] = excl;
    s_cnt[2 * t + 1] = excl + a;
    __syncthreads();
    for (uint32_t base = lo; base < hi; base += 32u)
    {
        const uint32_t i = base + lane;
        const bool on = i < hi;
        const uint32_t c = on ? tile_class(cost[i], mx) : 0xffffffffu;
        const unsigned peers = __match_any_sync(0xffffffffu, c);
        if (on)
        {
            const uint32_t at = s_cnt[c * 32u + warp] + __popc(peers & ((1u << lane) - 1u)); // rank in descending cost
            order[at] = i;
        }
        __syncwarp();
        if (on && lane == (uint32_t)(__ffs(peers) - 1))
            s_cnt[c * 32u + warp] += __popc(peers);
        __syncwarp();
    }
}

// Second pass of a render call: the few pixels whose paths are far longer than the frame average (a crevice between
// two bright spheres can take 20x the mean) would otherwise each keep one lane -- and its whole CTA slot -- busy long
// after the rest of the grid has drained.  Here one warp takes one queued pixel at a time (handed out dynamically) and its
// 32 lanes share the pixel's remaining samples, each lane claiming the next unclaimed one as soon as its path ends; the 32
// partial sums are combined in a fixed butterfly order and added to the pixel's partial result from the first pass --
// deterministic, since which lane traces which sample depends only on the path lengths.  Primitives are read from global memory (L1-resident for the scenes that reach this path).
// TRAV = 2: the CTA first copies the top TOP_NODES nodes of the tree (breadth-first order: its top levels) into shared memory.
constexpr uint32_t TOP_NODES = 21; // three levels of 4-wide nodes: 1 + 4 + 16, 2.7 KB
// BEAM (direct mode): primary rays test their patch's candidate leaves instead of traversing (k_beam_lists).
// One thread per 8x4-pixel patch of the tile (the patches k_render_stragglers enumerates in direct mode, row-major): the patch's
// footprint -- centre ray against the rays through its four corners -- and its candidate leaf list (beam_collect).
__global__ void __launch_bounds__(128) k_beam_lists(const SceneDev sc, const RenderParams p, BeamList* __restrict__ lists)
{
    const uint32_t tile_w = p.tile_x1 - p.tile_x0, tile_h = p.tile_y1 - p.tile_y0;
    const uint32_t patches_x = (tile_w + BEAM_PATCH_W - 1u) / BEAM_PATCH_W, patches_y = (tile_h + BEAM_PATCH_H - 1u) / BEAM_PATCH_H;
    const uint32_t patch = blockIdx.x * blockDim.x + threadIdx.x;
    if (patch >= patches_x * patches_y) return;
    const uint32_t x0 = p.tile_x0 + (patch % patches_x) * BEAM_PATCH_W, y0 = p.tile_y0 + (patch / patches_x) * BEAM_PATCH_H;
    const uint32_t x1 = min(x0 + BEAM_PATCH_W, p.tile_x1), y1 = min(y0 + BEAM_PATCH_H, p.tile_y1);
    const float fx0 = __uint2float_rn(x0), fy0 = __uint2float_rn(y0), fx1 = __uint2float_rn(x1), fy1 = __uint2float_rn(y1);
    const Ray rc = primary_ray(p.cam, 0.5f * (fx0 + fx1), 0.5f * (fy0 + fy1));
    float dd = 0.0f, oo = 0.0f;
#pragma unroll 1
    for (int k = 0; k < 4; k++)
    {
        const Ray rk = primary_ray(p.cam, (k & 1) ? fx1 : fx0, (k & 2) ? fy1 : fy0);
        dd = fmaxf(dd, fabsf(rk.d.x - rc.d.x) + fabsf(rk.d.y - rc.d.y) + fabsf(rk.d.z - rc.d.z));
        oo = fmaxf(oo, fabsf(rk.o.x - rc.o.x) + fabsf(rk.o.y - rc.o.y) + fabsf(rk.o.z - rc.o.z));
    }
    const float u16 = 16.0f * 5.9604645e-8f; // slack for the rounding of the ray arithmetic itself
    const float sigma = 1.01f * dd + u16, rho = 1.01f * oo + u16 * (fabsf(rc.o.x) + fabsf(rc.o.y) + fabsf(rc.o.z) + 1.0f);
    beam_collect(sc.bvh_nodes, rc, sigma, rho, lists + patch);
}

template <bool BVH, int G_LANES = 32, int TRAV = 1, bool BEAM = false, int MINB = 8, bool NESTED = false, int NP = 0>
__global__ void __launch_bounds__(128, MINB) k_render_stragglers(const SceneDev sc, const RenderParams p)
{
    __shared__ float4 s_top[TRAV == 2 ? 8 * TOP_NODES : 1];
    TopNodes top;
    top.nodes = s_top;
    top.count = 0;
    if (BVH && TRAV == 2)
    {
        top.count = min(TOP_NODES, sc.n_bvh_nodes);
        for (uint32_t i = threadIdx.x; i < 8u * top.count; i += blockDim.x) s_top[i] = __ldg(sc.bvh_nodes + i);
        __syncthreads();
    }
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t tile_w = p.tile_x1 - p.tile_x0, tile_h = p.tile_y1 - p.tile_y0;
    // G lanes share one pixel (32 in queue mode; 16 or 8 in direct mode, so that every lane gets >= 2 samples): a warp works
    // on 32 / G horizontally adjacent pixels at a time, lane group g on pixel g
    constexpr uint32_t G = G_LANES, ppw = 32u / G;
    const uint32_t grp = lane / G, gmask = (G == 32u ? 0xffffffffu : ((1u << G) - 1u) << (grp * G)), below = gmask & ((1u << lane) - 1u);
    const uint32_t count = p.direct ? ((tile_w + 7u) >> 3) * ((tile_h + 3u) >> 2) * G : *p.straggler_count;
    unsigned long long segs = 0;
    BvhStats bst;
    bst.nodes = 0;
    bst.tests = 0;
    for (;;)
    {
        // work items are handed out dynamically (their cost differs by orders of magnitude); counters[3] doubles as the cursor
        uint32_t item = 0;
        if (lane == 0) item = (uint32_t)atomicAdd(p.counters + 3, 1ull);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= count) break;
        uint2 w = make_uint2(0u, p.sample_end); // {pixel, first sample}; first sample = sample_end: nothing to do
        if (p.direct)
        {
            // pixels in 8x4 patches, patches row-major: consecutive warps work on neighbouring pixels (warm BVH nodes in L1)
            const uint32_t patches_x = (tile_w + 7u) >> 3, patch = item / G, in_patch = (item % G) * ppw + grp;
            const uint32_t qx = (patch % patches_x) * 8u + (in_patch & 7u), qy = (patch / patches_x) * 4u + (in_patch >> 3);
            if (qx < tile_w && qy < tile_h) // else: ragged edge of the patch grid
                w = make_uint2((p.tile_y0 + qy) * p.width + p.tile_x0 + qx, p.sample_begin);
        }
        else
            w = p.stragglers[item];
        const uint32_t px = w.x % p.width, py = w.x / p.width;
        RngKey key;
        key.ks = &p.rk;
        key.pixel = w.x;
        key.sample = 0;
        V3 sum = v3(0.0f, 0.0f, 0.0f);
        V3 thr = v3(1.0f, 1.0f, 1.0f);
        uint32_t seg = 0;
        Ray ray;
        ray.o = v3(0.0f, 0.0f, 0.0f);
        ray.d = v3(0.0f, 0.0f, 1.0f);
        // the candidate leaves of this pixel's 8x4 patch (k_beam_lists): n < 0 = the patch traverses
        const BeamList* beam = nullptr;
        int beam_n = -1;
        if (BVH && BEAM && p.direct && p.beam)
        {
            beam = p.beam + item / G; // direct mode: item / G is the patch
            beam_n = __ldg(&beam->n);
        }
        const bool use_beam = beam_n >= 0;
        // the lanes of a group share their pixel's remaining samples: a lane whose path ended takes the next unclaimed sample
        // at once (ballot rank in lane order -- deterministic), so lanes stay busy until the pixel runs out of samples
        uint32_t next = w.y;
        bool live = false;
        if (!BVH && NESTED)
        {
            // scan scenes with a handful of primitives: generate and shade dominate, and they run fullest when a lane simply loops
            // over its own share of the pixel's samples (lane l of the group: samples l, l + G, ...; cf. k_render_mega's plain loop)
            key.sample = w.y + (lane & (G - 1u));
            if (key.sample < p.sample_end)
            {
                ray = generate(p.cam, key, px, py);
                for (;;)
                {
                    segs++;
                    if (segment_step<BVH, TRAV, NP>(sc, p, sc.pairs, sc.planes, key, ray, thr, sum, seg, bst, top))
                    {
                        key.sample += G;
                        if (key.sample >= p.sample_end) break;
                        seg = 0;
                        thr = v3(1.0f, 1.0f, 1.0f);
                        ray = generate(p.cam, key, px, py);
                    }
                }
            }
            __syncwarp();
        }
        else
        for (;;)
        {
            const unsigned idle = __ballot_sync(0xffffffffu, !live) & gmask;
            const uint32_t mine = next + __popc(idle & below);
            const bool fresh = !live && mine < p.sample_end;
            if (fresh)
            {
                key.sample = mine;
                seg = 0;
                thr = v3(1.0f, 1.0f, 1.0f);
                ray = generate(p.cam, key, px, py);
                live = true;
            }
            next = min(p.sample_end, next + (uint32_t)__popc(idle));
            if (!__any_sync(0xffffffffu, live)) break;
            if (BVH && BEAM)
            {
                // Two passes over ONE copy of the segment code (the loop is kept rolled: instruction-cache footprint, see
                // closest_plane_cold).  Pass 0: the lanes that have just started a sample of a pixel with a beam list find their
                // primary hit in the list and shade it; pass 1: every lane in flight -- those lanes now with their first bounce --
                // traverses.  One iteration thus advances a fresh lane by two segments and the traversal only ever sees
                // secondary rays.
                const bool listed = use_beam && fresh && fabsf(dot3(ray.d, ray.d) - 1.0f) <= BEAM_EPS_D;
#pragma unroll 1
                for (int pass = __any_sync(0xffffffffu, listed) ? 0 : 1; pass < 2; pass++)
                {
                    if (pass == 0 ? listed : live)
                    {
                        segs++;
                        Hit h;
                        if (pass == 0)
                        {
                            float ts;
                            int is;
                            beam_closest_sphere(sc, beam, beam_n, ray, ts, is, bst);
                            h = combine_with_planes(sc, sc.planes, ray, ts, is);
                        }
                        else
                            h = closest_hit_bvh<TRAV>(sc, sc.planes, ray, bst, top);
                        if (shade_segment<BVH>(sc, p, sc.pairs, sc.planes, key, ray, thr, sum, seg, h)) live = false;
                    }
                }
            }
            else if (live)
            {
                segs++;
                if (segment_step<BVH, TRAV>(sc, p, sc.pairs, sc.planes, key, ray, thr, sum, seg, bst, top)) live = false;
            }
        }
        for (uint32_t off = G >> 1; off > 0; off >>= 1) // fixed butterfly inside the group
        {
            sum.x = __fadd_rn(sum.x, __shfl_xor_sync(0xffffffffu, sum.x, off));
            sum.y = __fadd_rn(sum.y, __shfl_xor_sync(0xffffffffu, sum.y, off));
            sum.z = __fadd_rn(sum.z, __shfl_xor_sync(0xffffffffu, sum.z, off));
        }
        if ((lane & (G - 1u)) == 0u && w.y < p.sample_end)
        {
            // queue mode: add onto the first pass' partial result; direct mode: store (or add when the call accumulates)
            float4 acc = (p.direct && !p.accumulate) ? make_float4(0.0f, 0.0f, 0.0f, 0.0f) : p.accum[w.x];
            acc.x = __fadd_rn(acc.x, sum.x); acc.y = __fadd_rn(acc.y, sum.y); acc.z = __fadd_rn(acc.z, sum.z);
            acc.w = __fadd_rn(acc.w, (float)(p.sample_end - w.y));
            p.accum[w.x] = acc;
            if (p.rgba8)
                p.rgba8[w.x] = pack_pixel(acc.x, acc.y, acc.z, p.spp_resolve);
        }
    }
    unsigned long long nodes = bst.nodes, tests = bst.tests;
    for (int off = 16; off > 0; off >>= 1)
    {
        segs += __shfl_down_sync(0xffffffffu, segs, off);
        nodes += __shfl_down_sync(0xffffffffu, nodes, off);
        tests += __shfl_down_sync(0xffffffffu, tests, off);
    }
    if (lane == 0 && segs)
    {
        atomicAdd(p.counters, segs);
        if (BVH)
        {
            atomicAdd(p.counters + 1, nodes);
            atomicAdd(p.counters + 2, tests);
        }
    }
}

// ---- direct mode without the end-of-pixel stall: lanes move on to the group's next pixel -------------------------------------
// k_render_stragglers in direct mode lets G lanes share ONE pixel's samples and ends the pixel with a reduction: when the pixel
// has no unclaimed sample left, its lanes idle until the slowest path is done.  With 4 samples per lane (C4: 64 spp, 16 lanes)
// that idling is 18 of 32 lanes active (ncu) -- the compaction tools/bvh_warp_sim.cpp priced at 9 %.  Here a warp takes a whole
// 8x4 patch and each of its 32 / G lane groups a RUN of G pixels of it; a lane whose path has ended claims the next unclaimed
// (pixel, sample) of the run in lexicographic order -- by ballot rank, so the assignment depends on path lengths only and is
// deterministic -- and simply carries on into the next pixel.  A lane visits the run's pixels in ascending order, so it adds
// to each pixel at most one partial sum: on leaving a pixel it parks that sum in a scratch slot (run pixel, lane), and when the
// run is done lane k adds up pixel k's slots in lane order (a fixed order: same paths, same segment counts, sums equal to the
// other kernels' up to fp32 order).  The stall now happens once per G pixels instead of once per pixel.
// Beams: the warp's patch is one beam list.  The scratch slots live in global memory (L2; a few accesses per lane and run).
// A warp item is GROUPS x RUN pixels of a patch (RUN <= G); shorter runs keep the item count high enough for the tail.
// MEASURED (B200, RTCU_BVH_RUNS=1), against k_render_stragglers direct mode: RUN 8: C3 33.1 vs 30.5 ms, C4 79.6 vs 80.5 ms, C5/8 241.5 vs
// 231.7 ms; RUN 4: 32.7 / 83.2 / 242.6; RUN 16 (a whole patch per warp): 38.9 / 84.0 / 248.8.  The stall it removes is real (C4) but at
// >= 256 samples per pixel there is little stall to remove and the claim bookkeeping, the lanes of a group straddling two pixels and
// the 172 B of spills (128 B there) cost more.  Kept as an experiment, not the default.
template <int G_LANES, bool BEAM, int RUN_PIXELS>
__global__ void __launch_bounds__(128, 8) k_render_runs(const SceneDev sc, const RenderParams p)
{
    // a warp item is GROUPS x RUN consecutive pixels of a patch (in the patch's row-major order); SUBS items make a patch
    constexpr uint32_t G = G_LANES, GROUPS = 32u / G, RUN = RUN_PIXELS, SUBS = 32u / (GROUPS * RUN);
    static_assert(RUN <= G && GROUPS * RUN * SUBS == 32u, "a run is at most G pixels (lane k sums pixel k) and runs tile the patch");
    const uint32_t lane = threadIdx.x & 31u, lg = lane & (G - 1u), grp = lane / G;
    const uint32_t gmask = ((G == 32u) ? 0xffffffffu : ((1u << G) - 1u)) << (grp * G), below = gmask & ((1u << lane) - 1u);
    const uint32_t tile_w = p.tile_x1 - p.tile_x0, tile_h = p.tile_y1 - p.tile_y0;
    const uint32_t patches_x = (tile_w + 7u) >> 3, count = patches_x * ((tile_h + 3u) >> 2) * SUBS;
    const uint32_t spp = p.sample_end - p.sample_begin; // >= 16 >= G: a claim wraps into the next pixel at most once
    float* const slots = p.run_scratch + ((size_t)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * GROUPS + grp) * (RUN * 3u * G);
    unsigned long long segs = 0;
    BvhStats bst;
    bst.nodes = 0;
    bst.tests = 0;
    const TopNodes top = TopNodes{ nullptr, 0u };
    for (;;)
    {
        uint32_t item = 0;
        if (lane == 0) item = (uint32_t)atomicAdd(p.counters + 3, 1ull);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= count) break;
        const uint32_t patch = item / SUBS, first = (item % SUBS) * GROUPS * RUN + grp * RUN; // the run's first pixel within the patch
        const uint32_t patch_x0 = (patch % patches_x) * 8u, patch_y0 = (patch / patches_x) * 4u;
        // the run's pixels that lie inside the tile (ragged patches at the right and bottom edge), as a bit mask over the run
        uint32_t valid = 0;
#pragma unroll
        for (uint32_t k = 0; k < RUN; k++)
        {
            const uint32_t in_patch = first + k;
            valid |= (patch_x0 + (in_patch & 7u) < tile_w && patch_y0 + (in_patch >> 3) < tile_h) ? (1u << k) : 0u;
        }
        const uint32_t n_valid = __popc(valid);
        for (uint32_t k = 0; k < RUN * 3u; k++) __stcg(slots + k * G + lg, 0.0f);
        const BeamList* beam = nullptr;
        int beam_n = -1;
        if (BEAM && p.beam)
        {
            beam = p.beam + patch;
            beam_n = __ldg(&beam->n);
        }
        const bool use_beam = beam_n >= 0;

        RngKey key;
        key.ks = &p.rk;
        key.pixel = 0;
        key.sample = 0;
        V3 sum = v3(0.0f, 0.0f, 0.0f);
        V3 thr = v3(1.0f, 1.0f, 1.0f);
        uint32_t seg = 0, px = 0, py = 0;
        int cur = -1; // the valid pixel of the run this lane is adding to
        Ray ray;
        ray.o = v3(0.0f, 0.0f, 0.0f);
        ray.d = v3(0.0f, 0.0f, 1.0f);
        uint32_t next_j = 0, next_s = 0; // the next unclaimed (valid pixel, sample) of the run; uniform within the group
        bool live = false;
        for (;;)
        {
            const unsigned idle = __ballot_sync(0xffffffffu, !live) & gmask;
            uint32_t j = next_j, smp = next_s + __popc(idle & below);
            if (smp >= spp) { smp -= spp; j++; }
            const bool fresh = !live && j < n_valid;
            if (fresh)
            {
                if ((int)j != cur)
                {
                    if (cur >= 0) // leaving a pixel: park this lane's share of it
                    {
                        __stcg(slots + (cur * 3u + 0u) * G + lg, sum.x);
                        __stcg(slots + (cur * 3u + 1u) * G + lg, sum.y);
                        __stcg(slots + (cur * 3u + 2u) * G + lg, sum.z);
                    }
                    cur = (int)j;
                    sum = v3(0.0f, 0.0f, 0.0f);
                    const uint32_t in_patch = first + __fns(valid, 0u, (int)j + 1); // the j-th valid pixel of the run
                    px = p.tile_x0 + patch_x0 + (in_patch & 7u);
                    py = p.tile_y0 + patch_y0 + (in_patch >> 3);
                    key.pixel = py * p.width + px;
                }
                key.sample = p.sample_begin + smp;
                seg = 0;
                thr = v3(1.0f, 1.0f, 1.0f);
                ray = generate(p.cam, key, px, py);
                live = true;
            }
            next_s += (uint32_t)__popc(idle);
            if (next_s >= spp) { next_s -= spp; next_j++; }
            if (next_j > n_valid) next_j = n_valid; // (all claimed: stay put)
            if (!__any_sync(0xffffffffu, live)) break;
            // pass 0: the lanes that have just started a sample find their primary hit in the patch's list; pass 1: every lane in
            // flight traverses (one rolled copy of the segment code, see k_render_stragglers)
            const bool listed = BEAM && use_beam && fresh && fabsf(dot3(ray.d, ray.d) - 1.0f) <= BEAM_EPS_D;
#pragma unroll 1
            for (int pass = (BEAM && __any_sync(0xffffffffu, listed)) ? 0 : 1; pass < 2; pass++)
            {
                if (pass == 0 ? listed : live)
                {
                    segs++;
                    Hit h;
                    if (BEAM && pass == 0)
                    {
                        float ts;
                        int is;
                        beam_closest_sphere(sc, beam, beam_n, ray, ts, is, bst);
                        h = combine_with_planes(sc, sc.planes, ray, ts, is);
                    }
                    else
                        h = closest_hit_bvh<0>(sc, sc.planes, ray, bst, top);
                    if (shade_segment<true>(sc, p, sc.pairs, sc.planes, key, ray, thr, sum, seg, h)) live = false;
                }
            }
        }
        if (cur >= 0)
        {
            __stcg(slots + (cur * 3u + 0u) * G + lg, sum.x);
            __stcg(slots + (cur * 3u + 1u) * G + lg, sum.y);
            __stcg(slots + (cur * 3u + 2u) * G + lg, sum.z);
        }
        __syncwarp();
        if (lg < n_valid) // lane k of the group adds up the run's k-th valid pixel, shares in lane order
        {
            float sx = 0.0f, sy = 0.0f, sz = 0.0f;
            for (uint32_t l = 0; l < G; l++)
            {
                sx = __fadd_rn(sx, __ldcg(slots + (lg * 3u + 0u) * G + l));
                sy = __fadd_rn(sy, __ldcg(slots + (lg * 3u + 1u) * G + l));
                sz = __fadd_rn(sz, __ldcg(slots + (lg * 3u + 2u) * G + l));
            }
            const uint32_t in_patch = first + __fns(valid, 0u, (int)lg + 1);
            const uint32_t pix = (p.tile_y0 + patch_y0 + (in_patch >> 3)) * p.width + p.tile_x0 + patch_x0 + (in_patch & 7u);
            float4 acc = p.accumulate ? p.accum[pix] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            acc.x = __fadd_rn(acc.x, sx); acc.y = __fadd_rn(acc.y, sy); acc.z = __fadd_rn(acc.z, sz);
            acc.w = __fadd_rn(acc.w, (float)spp);
            p.accum[pix] = acc;
            if (p.rgba8)
                p.rgba8[pix] = pack_pixel(acc.x, acc.y, acc.z, p.spp_resolve);
        }
        __syncwarp(); // the slots are zeroed again for the next patch only after they have been read
    }
    unsigned long long nodes = bst.nodes, tests = bst.tests;
    for (int off = 16; off > 0; off >>= 1)
    {
        segs += __shfl_down_sync(0xffffffffu, segs, off);
        nodes += __shfl_down_sync(0xffffffffu, nodes, off);
        tests += __shfl_down_sync(0xffffffffu, tests, off);
    }
    if (lane == 0 && segs)
    {
        atomicAdd(p.counters, segs);
        atomicAdd(p.counters + 1, nodes);
        atomicAdd(p.counters + 2, tests);
    }
}

// mg_ray_tracer.cpp:195-200 over a whole accumulation buffer
__global__ void k_resolve(const float4* __restrict__ accum, uint32_t n, float spp, uint32_t* __restrict__ rgba8)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
    {
        const float4 a = accum[i];
        rgba8[i] = pack_pixel(a.x, a.y, a.z, spp);
    }
}

// resolve fused with the cross-GPU sum: peers[g] are the other devices' accumulation buffers, read
// through NVLink peer mappings; the sum order is device 0, 1, 2, ... (deterministic).
struct PeerList { const float4* ptr[8]; int n; };
__global__ void k_reduce_resolve(float4* __restrict__ accum, const PeerList peers, uint32_t n, float spp, uint32_t* __restrict__ rgba8)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
    {
        float4 a = accum[i];
        for (int g = 0; g < peers.n; g++)
        {
            const float4 b = peers.ptr[g][i];
            a.x = __fadd_rn(a.x, b.x); a.y = __fadd_rn(a.y, b.y); a.z = __fadd_rn(a.z, b.z); a.w = __fadd_rn(a.w, b.w);
        }
        accum[i] = a;
        if (rgba8)
            rgba8[i] = pack_pixel(a.x, a.y, a.z, spp);
    }
}

// The same for one process per GPU: bufs[g] are all ranks' accumulation buffers (the rank's own plus the peers' buffers
// opened through CUDA IPC and read over NVLink), summed in rank order 0, 1, 2, ... for the pixels [first, first + count) of
// this rank's row band; the packed pixels go straight into the destination rank's image (a peer store when that is
// another GPU).  One kernel is the whole exchange step: no staging copy, no collective library on the data path.
__global__ void __launch_bounds__(256) k_reduce_resolve_rows(const PeerList bufs, uint32_t first, uint32_t count, float spp, uint32_t* __restrict__ rgba8)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < count)
    {
        const uint32_t i = first + k;
        float4 a = bufs.ptr[0][i];
        for (int g = 1; g < bufs.n; g++)
        {
            const float4 b = bufs.ptr[g][i];
            a.x = __fadd_rn(a.x, b.x); a.y = __fadd_rn(a.y, b.y); a.z = __fadd_rn(a.z, b.z); a.w = __fadd_rn(a.w, b.w);
        }
        rgba8[i] = pack_pixel(a.x, a.y, a.z, spp);
    }
}

// ---- the exchange step of a one-process-per-GPU frame as ONE launch per rank and no collective library --------------------
// Every rank owns one ExchangeFlags block in CUDA-IPC-shared device memory (zeroed at allocation).  Frame e (1, 2, 3, ...):
//   ready[g] == e   rank g's accumulation buffer of frame e is complete        (rank g stores it into EVERY rank's block)
//   done[g]  == e   rank g has stored its row band of frame e into the image   (rank g stores it into the DESTINATION's block)
// The kernel is stream-ordered after the rank's own trace kernels, so its first CTA to arrive publishes ready[rank] with
// st.release.sys over NVLink; all CTAs then poll their LOCAL block (ld.acquire.sys) until every rank has published, sum the
// buffers of their pixels through peer loads in rank order, resolve and store the packed pixels into the destination's image.
// The last CTA to finish publishes done[rank]; on the destination it also waits for the other ranks' done flags, so the end of
// the destination's kernel is the end of the frame.  The callers alternate between two accumulation buffers: rank r overwrites
// a buffer two frames after its peers read it, and a peer cannot publish ready[e+1] before its frame-e kernel has finished
// reading -- so no flag is needed to release the buffers.  Every wait gives up after `timeout` clock cycles and sets `error`
// (a rank that died must not hang the others' GPUs).
struct ExchangeFlags {
    unsigned int ready[8];
    unsigned int done[8];
    unsigned int ticket, finished; // arrival / completion counters of the local kernel (re-armed by its last CTA)
    unsigned int error;
    unsigned int pad[13];
};
static_assert(sizeof(ExchangeFlags) == 128, "ExchangeFlags is the 128-byte block rtcu_exchange_reduce_resolve documents");
struct FlagList { ExchangeFlags* ptr[8]; };

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p)
{
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// spins until *p has reached `epoch` (wrap-safe); false after `timeout` cycles
__device__ __forceinline__ bool wait_epoch(const unsigned int* p, unsigned int epoch, long long timeout)
{
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(p) - epoch) < 0)
    {
        if (clock64() - t0 > timeout) return false;
        __nanosleep(100);
    }
    return true;
}

__global__ void __launch_bounds__(256) k_exchange_reduce_resolve(const PeerList bufs, const FlagList flags, int rank, int dst, unsigned int epoch,
                                                                 uint32_t first, uint32_t count, float spp, uint32_t* __restrict__ rgba8, long long timeout)
{
    ExchangeFlags* mine = flags.ptr[rank];
    if (threadIdx.x == 0 && atomicAdd(&mine->ticket, 1u) == 0u)
    {
        __threadfence_system();
        for (int g = 0; g < bufs.n; g++) st_release_sys(&flags.ptr[g]->ready[rank], epoch);
    }
    if (threadIdx.x < (unsigned)bufs.n && !wait_epoch(&mine->ready[threadIdx.x], epoch, timeout)) atomicExch(&mine->error, 1u);
    __syncthreads();
    // (a grid of a few CTAs per SM striding over the band: the handshake, the fence and the completion count are per CTA)
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x)
    {
        const uint32_t i = first + k;
        float4 a = bufs.ptr[0][i];
        for (int g = 1; g < bufs.n; g++)
        {
            const float4 b = bufs.ptr[g][i];
            a.x = __fadd_rn(a.x, b.x); a.y = __fadd_rn(a.y, b.y); a.z = __fadd_rn(a.z, b.z); a.w = __fadd_rn(a.w, b.w);
        }
        rgba8[i] = pack_pixel(a.x, a.y, a.z, spp);
    }
    __syncthreads();
    if (threadIdx.x == 0)
        __threadfence_system(); // cumulative: the CTA's pixel stores (peer stores when the image lives on another GPU) before the count
    if (threadIdx.x == 0 && atomicAdd(&mine->finished, 1u) == gridDim.x - 1u)
    {
        __threadfence_system();
        mine->ticket = 0u;
        mine->finished = 0u;
        st_release_sys(&flags.ptr[dst]->done[rank], epoch);
        if (rank == dst)
            for (int g = 0; g < bufs.n; g++)
                if (!wait_epoch(&mine->done[g], epoch, timeout)) atomicExch(&mine->error, 1u);
    }
}

// ---- step-wise parity kernels ----------------------------------------------------------------------
template <bool STAGE, bool BVH, int TRAV = 1>
__global__ void __launch_bounds__(256) k_intersect_batch(const SceneDev sc, const float* __restrict__ o, const float* __restrict__ d,
                                                         uint32_t n, uint8_t* __restrict__ hit, uint32_t* __restrict__ prim,
                                                         float* __restrict__ t, float* __restrict__ nrm, unsigned long long* counters)
{
    extern __shared__ float4 smem[];
    const float4* s_sph = sc.pairs;
    const float4* s_pl = sc.planes;
    if (STAGE)
    {
        const uint32_t n4 = pair_float4_count(sc.n_spheres);
        for (uint32_t i = threadIdx.x; i < n4; i += blockDim.x)
            smem[i] = __ldg(sc.pairs + i);
        for (uint32_t i = threadIdx.x; i < sc.n_planes; i += blockDim.x)
            smem[n4 + i] = __ldg(sc.planes + i);
        __syncthreads();
        s_sph = smem;
        s_pl = smem + n4;
    }
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        Ray r;
        r.o = v3(o[3 * i], o[3 * i + 1], o[3 * i + 2]);
        r.d = v3(d[3 * i], d[3 * i + 1], d[3 * i + 2]);
        BvhStats bst;
        bst.nodes = 0;
        bst.tests = 0;
        const Hit h = BVH ? closest_hit_bvh<TRAV>(sc, s_pl, r, bst) : closest_hit_linear(s_sph, sc.n_spheres, s_pl, sc.n_planes, r);
        hit[i] = h.prim != RTCU_PRIM_MISS;
        prim[i] = h.prim;
        t[i] = h.t;
        if (nrm)
        {
            V3 nn = v3(0.0f, 0.0f, 0.0f);
            if (h.prim != RTCU_PRIM_MISS)
                nn = BVH ? hit_normal_global(sc, s_pl, r, h) : hit_normal(s_sph, s_pl, r, h);
            nrm[3 * i] = nn.x; nrm[3 * i + 1] = nn.y; nrm[3 * i + 2] = nn.z;
        }
        if (BVH)
        {
            atomicAdd(counters + 1, (unsigned long long)bst.nodes);
            atomicAdd(counters + 2, (unsigned long long)bst.tests);
        }
    }
}

__global__ void k_primary_rays(const CameraConst cam, uint32_t width, const PhiloxKeys rk, const uint32_t* __restrict__ px,
                               const uint32_t* __restrict__ py, const uint32_t* __restrict__ sample, uint32_t n,
                               float* __restrict__ o, float* __restrict__ d)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    RngKey k;
    k.ks = &rk; k.pixel = py[i] * width + px[i]; k.sample = sample[i];
    const Ray r = generate(cam, k, px[i], py[i]);
    o[3 * i] = r.o.x; o[3 * i + 1] = r.o.y; o[3 * i + 2] = r.o.z;
    d[3 * i] = r.d.x; d[3 * i + 1] = r.d.y; d[3 * i + 2] = r.d.z;
}

__global__ void k_scatter_batch(const SceneDev sc, uint32_t mode, const PhiloxKeys rk, uint32_t n, const uint32_t* __restrict__ material,
                                const float* __restrict__ o, const float* __restrict__ d, const float* __restrict__ t,
                                const float* __restrict__ nrm, const uint32_t* __restrict__ pixel, const uint32_t* __restrict__ sample,
                                const uint32_t* __restrict__ block, uint8_t* __restrict__ scattered, float* __restrict__ att,
                                float* __restrict__ o_out, float* __restrict__ d_out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Ray r;
    r.o = v3(o[3 * i], o[3 * i + 1], o[3 * i + 2]);
    r.d = v3(d[3 * i], d[3 * i + 1], d[3 * i + 2]);
    const V3 nn = v3(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]);
    RngKey k;
    k.ks = &rk; k.pixel = pixel[i]; k.sample = sample[i];
    const MatRec m = load_material(sc, material[i]);
    const uint4 rnd = rng_block(k, block[i], 0u);
    Ray out;
    out.d = v3(0.0f, 0.0f, 0.0f);
    const bool ok = scatter(scatter_kind(mode, m.type), m, r, t[i], nn, k, block[i], rnd, out);
    scattered[i] = ok;
    att[3 * i] = m.att_r; att[3 * i + 1] = m.att_g; att[3 * i + 2] = m.att_b;
    o_out[3 * i] = out.o.x; o_out[3 * i + 1] = out.o.y; o_out[3 * i + 2] = out.o.z;
    d_out[3 * i] = ok ? out.d.x : 0.0f; d_out[3 * i + 1] = ok ? out.d.y : 0.0f; d_out[3 * i + 2] = ok ? out.d.z : 0.0f;
}

__global__ void k_philox_batch(const uint4* __restrict__ ctr, uint32_t n, const PhiloxKeys rk, uint4* __restrict__ out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        out[i] = philox4x32(ctr[i], rk);
}


// ---- exhaustive check of the cheaper exact special functions of spec.cuh against the IEEE intrinsics ---------------
// counts[0], counts[1]: float patterns x (all 2^32) for which sqrt_then_rcp's s / inv differ from __fsqrt_rn(x) /
// __frcp_rn(__fsqrt_rn(x)) (two NaNs count as equal); counts[2]: (a, b) pairs with a in {0} u [2^-24, 2^24] (every float)
// and b from `divisors` for which div_by_const(a, b, RN(1/b)) differs from __fdiv_rn(a, b); counts[3]: pairs tested.
__device__ __forceinline__ bool same_float(float a, float b)
{
    return __float_as_uint(a) == __float_as_uint(b) || (a != a && b != b);
}
__global__ void __launch_bounds__(256) k_selftest_math(const float* __restrict__ divisors, uint32_t n_divisors, unsigned long long* __restrict__ counts)
{
    unsigned long long bad_s = 0, bad_inv = 0, bad_div = 0, pairs = 0;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < 0x100000000ull; i += stride)
    {
        const float x = __uint_as_float((uint32_t)i);
        float s, inv;
        sqrt_then_rcp(x, s, inv);
        const float rs = __fsqrt_rn(x);
        bad_s += !same_float(s, rs);
        bad_inv += !same_float(inv, __frcp_rn(rs));
        const bool in_domain = i == 0 || (i >= 0x33800000ull && i <= 0x4B800000ull); // 0, or 2^-24 .. 2^24
        if (in_domain)
            for (uint32_t k = 0; k < n_divisors; k++)
            {
                const float b = divisors[k];
                bad_div += !same_float(div_by_const(x, b, __frcp_rn(b)), __fdiv_rn(x, b));
                pairs++;
            }
    }
    atomicAdd(counts + 0, bad_s);
    atomicAdd(counts + 1, bad_inv);
    atomicAdd(counts + 2, bad_div);
    atomicAdd(counts + 3, pairs);
}

// ---- FP32 peak calibration: dependent-chain-free FFMA / FFMA2 streams, 16 independent accumulators per
// thread, no memory traffic.  Used by bench.py to report the roofline denominator at the clocks actually seen.
template <bool PACKED>
__global__ void __launch_bounds__(256) k_fp32_peak(float* __restrict__ out, int iters, float a, float b)
{
    float2 acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++)
        acc[i] = make_float2(threadIdx.x * 1e-3f + i, blockIdx.x * 1e-4f - i);
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int it = 0; it < iters; it++)
    {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int i = 0; i < 8; i++)
            {
                if (PACKED)
                    acc[i] = __ffma2_rn(acc[i], a2, b2);
                else
                {
                    acc[i].x = __fmaf_rn(acc[i].x, a, b);
                    acc[i].y = __fmaf_rn(acc[i].y, a, b);
                }
            }
    }
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; i++)
        s += acc[i].x + acc[i].y;
    if (s == 123.456f) // never true; keeps the chain alive
        out[0] = s;
}

} // namespace rtcu_dev
