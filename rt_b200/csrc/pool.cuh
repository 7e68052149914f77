// pool.cuh -- BVH scenes: a warp-local ray pool that decouples traversal from shading.
//
// Measured on k_render_mega<BVH> (C4, 100 001 spheres): primary rays keep 24.5/32 lanes active, but with bounces the mix of
// cheap rays (sky, far field) and expensive ones (a bounce that starts inside the sphere field and runs along it) drops the
// average to 15/32 -- every lane waits at the end of the traversal loop for the slowest ray of the warp before anybody
// shades.  Here a warp owns a pool of P path records in shared memory (ray, throughput, pixel slot, sample, segment):
//
//   refill   lanes without a ray pop the next READY record (warp ballot + rank, no atomics) and start its traversal
//   step     a few BVH node visits for every lane that holds a ray; finished traversals are appended to the PENDING queue
//   shade    once 32 records are pending (or nothing else can make progress) the whole warp shades one batch, one record per
//            lane: sky / scatter; a path that ends adds its radiance to the pixel's accumulator and the record is re-armed
//            with the next (pixel, sample) of the warp's work list; survivors go back to READY
//
// so cheap rays flow through the lanes while expensive ones keep traversing, and shading always runs on a full batch.  The
// closest hit of every ray is still the linear scan's exact answer (same trav_step), paths are identical (counter-based RNG),
// the result is deterministic; only the order in which a pixel's samples are summed differs from the sequential order.
#pragma once
#include "kernels.cuh"

namespace rtcu_dev {

#ifndef RTCU_POOL_K
#define RTCU_POOL_K 4 // BVH node visits per loop iteration
#endif
#ifndef RTCU_POOL_REFILL
#define RTCU_POOL_REFILL 16 // idle lanes that trigger a refill (swept 4..16 x K 1..4: 16 / 4 is the fastest on C3s / C4s)
#endif
#ifndef RTCU_POOL_BLOCKS
#define RTCU_POOL_BLOCKS 6
#endif
constexpr int POOL_P = 64;          // records per warp
constexpr int POOL_WARPS = 4;       // warps per CTA (128 threads, as the megakernel)

struct WarpPool {
    float ox[POOL_P], oy[POOL_P], oz[POOL_P], dx[POOL_P], dy[POOL_P], dz[POOL_P];
    float tr[POOL_P], tg[POOL_P], tb[POOL_P];
    float hit_t[POOL_P];
    uint32_t hit_prim[POOL_P];
    uint32_t sample[POOL_P];
    uint16_t seg[POOL_P];
    uint8_t slot[POOL_P];
    uint8_t ready_q[POOL_P], pend_q[POOL_P]; // ring buffers of record indices
    float acc_r[32], acc_g[32], acc_b[32];
};

__device__ __forceinline__ unsigned lanemask_lt()
{
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

__global__ void __launch_bounds__(32 * POOL_WARPS, RTCU_POOL_BLOCKS) k_render_pool(const SceneDev sc, const RenderParams p)
{
    __shared__ WarpPool pools[POOL_WARPS];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    WarpPool& wp = pools[warp];
    const float4* s_pl = sc.planes;

    // the warp's 8x4 pixel patch (same mapping as the megakernel); lane l owns pixel slot l for the final write
    const uint32_t px = p.tile_x0 + blockIdx.x * MEGA_TILE_W + (warp & 1u) * 8u + (lane & 7u);
    const uint32_t py = p.tile_y0 + blockIdx.y * MEGA_TILE_H + (warp >> 1) * 4u + (lane >> 3);
    const bool in_tile = px < p.tile_x1 && py < p.tile_y1;
    const unsigned valid = __ballot_sync(0xffffffffu, in_tile);
    const uint32_t n_valid = __popc(valid);
    const uint32_t n_samples = p.sample_end > p.sample_begin ? p.sample_end - p.sample_begin : 0u;
    const uint32_t total_work = n_valid * n_samples; // work item w: sample w / n_valid of the (w % n_valid)-th valid slot
    wp.acc_r[lane] = 0.0f;
    wp.acc_g[lane] = 0.0f;
    wp.acc_b[lane] = 0.0f;
    __syncwarp();

    // warp-uniform queue state, replicated in every lane
    uint32_t next_work = 0, ready_head = 0, ready_count = 0, pend_head = 0, pend_count = 0;
    const uint32_t base_x = p.tile_x0 + blockIdx.x * MEGA_TILE_W + (warp & 1u) * 8u, base_y = p.tile_y0 + blockIdx.y * MEGA_TILE_H + (warp >> 1) * 4u;

    // arms record `e` with work item `w` (a fresh primary ray) -- called by one lane per record
    auto arm = [&](uint32_t e, uint32_t w)
    {
        const uint32_t s = w / n_valid, k = w - s * n_valid;
        const uint32_t slot = __fns(valid, 0, (int)k + 1);
        const uint32_t x = base_x + (slot & 7u), y = base_y + (slot >> 3);
        RngKey key;
        key.ks = &p.rk;
        key.pixel = y * p.width + x;
        key.sample = p.sample_begin + s;
        const Ray r = generate(p.cam, key, x, y);
        wp.ox[e] = r.o.x; wp.oy[e] = r.o.y; wp.oz[e] = r.o.z;
        wp.dx[e] = r.d.x; wp.dy[e] = r.d.y; wp.dz[e] = r.d.z;
        wp.tr[e] = 1.0f; wp.tg[e] = 1.0f; wp.tb[e] = 1.0f;
        wp.sample[e] = key.sample;
        wp.seg[e] = 0;
        wp.slot[e] = (uint8_t)slot;
    };

    // initial fill: up to P records, 32 per round
    for (uint32_t r0 = 0; r0 < (uint32_t)POOL_P && next_work < total_work; r0 += 32u)
    {
        const uint32_t n = min(32u, min((uint32_t)POOL_P - r0, total_work - next_work));
        if (lane < n)
        {
            arm(r0 + lane, next_work + lane);
            wp.ready_q[(ready_head + ready_count + lane) % POOL_P] = (uint8_t)(r0 + lane);
        }
        next_work += n;
        ready_count += n;
        __syncwarp();
    }

    unsigned long long segs = 0;
    BvhStats bst;
    bst.nodes = 0;
    bst.tests = 0;
    bool has_ray = false;
    uint32_t entry = 0;
    Ray ray;
    ray.o = v3(0.0f, 0.0f, 0.0f);
    ray.d = v3(0.0f, 0.0f, 1.0f);
    Trav tv;
    tv.node = 0; tv.sp = 0; tv.best_t = 0.0f; tv.best_i = -1; tv.ix = tv.iy = tv.iz = 0.0f; tv.kappa = 0.0f; tv.madd = 0.0f;
    uint32_t stack_ref[BVH_STACK];
    float stack_t[BVH_STACK];

    for (;;)
    {
        // ---- refill: idle lanes take READY records -------------------------------------------------------------------
        const unsigned idle = __ballot_sync(0xffffffffu, !has_ray);
        bool finished = false; // this lane's traversal completed in this iteration
        // refills are batched: the refill path (6 LDS + trav_init) otherwise runs almost every iteration for ~2 lanes
        if (ready_count && (__popc(idle) >= RTCU_POOL_REFILL || idle == 0xffffffffu || (__popc(idle) && ready_count + pend_count < 32u)))
        {
            const uint32_t n_take = min((uint32_t)__popc(idle), ready_count);
            const uint32_t rank = __popc(idle & lanemask_lt());
            if (!has_ray && rank < n_take)
            {
                entry = wp.ready_q[(ready_head + rank) % POOL_P];
                ray.o = v3(wp.ox[entry], wp.oy[entry], wp.oz[entry]);
                ray.d = v3(wp.dx[entry], wp.dy[entry], wp.dz[entry]);
                has_ray = true;
                segs++;
                if (!trav_init(ray, tv))
                {
                    // direction too far from unit length for the conservative margins: scan all spheres now
                    float ts = __int_as_float(0x7f800000);
                    int is = -1;
                    const uint32_t n_pairs = (sc.n_spheres + 1u) >> 1;
                    for (uint32_t j = 0; j < n_pairs; j++)
                        sphere_pair_test(__ldg(sc.pairs + 2 * j), __ldg(sc.pairs + 2 * j + 1), (int)j, ray, ts, is);
                    bst.tests += sc.n_spheres;
                    tv.best_t = ts;
                    tv.best_i = is >= 0 ? is : 0x7fffffff;
                    finished = true;
                }
            }
            ready_head = (ready_head + n_take) % POOL_P;
            ready_count -= n_take;
        }

        // ---- one node visit for every lane that holds a ray -----------------------------------------------------------
        // (a few visits per iteration amortise the queue bookkeeping; a lane that finishes early idles for < POOL_K visits)
#pragma unroll 1
        for (int k = 0; k < RTCU_POOL_K; k++)
            if (has_ray && !finished)
                finished = trav_step(sc, ray, tv, stack_ref, stack_t, bst);

        // finished traversals -> PENDING (the sphere result travels in the record)
        const unsigned fin = __ballot_sync(0xffffffffu, has_ray && finished);
        if (fin)
        {
            if (has_ray && finished)
            {
                wp.hit_t[entry] = tv.best_t;
                wp.hit_prim[entry] = tv.best_i == 0x7fffffff ? RTCU_PRIM_MISS : (uint32_t)tv.best_i;
                wp.pend_q[(pend_head + pend_count + __popc(fin & lanemask_lt())) % POOL_P] = (uint8_t)entry;
                has_ray = false;
            }
            pend_count += __popc(fin);
            __syncwarp();
        }

        // ---- shade one batch when it is full, or when nothing else can make progress ------------------------------------
        const unsigned flying = __ballot_sync(0xffffffffu, has_ray);
        if (pend_count >= 32u || (pend_count && !ready_count && __popc(flying) <= 8))
        {
            const uint32_t n = min(32u, pend_count);
            bool ended = false, rearm = false;
            uint32_t e = 0, slot = 0;
            V3 rad = v3(0.0f, 0.0f, 0.0f);
            if (lane < n)
            {
                e = wp.pend_q[(pend_head + lane) % POOL_P];
                Ray r;
                r.o = v3(wp.ox[e], wp.oy[e], wp.oz[e]);
                r.d = v3(wp.dx[e], wp.dy[e], wp.dz[e]);
                V3 thr = v3(wp.tr[e], wp.tg[e], wp.tb[e]);
                uint32_t seg = wp.seg[e];
                slot = wp.slot[e];
                const float ts = wp.hit_t[e];
                const uint32_t prim = wp.hit_prim[e];
                const Hit h = combine_with_planes(sc, s_pl, r, ts, prim == RTCU_PRIM_MISS ? -1 : (int)prim);
                RngKey key;
                key.ks = &p.rk;
                key.pixel = (base_y + (slot >> 3)) * p.width + base_x + (slot & 7u);
                key.sample = wp.sample[e];
                ended = shade_segment<true>(sc, p, sc.pairs, s_pl, key, r, thr, rad, seg, h);
                if (!ended)
                {
                    wp.ox[e] = r.o.x; wp.oy[e] = r.o.y; wp.oz[e] = r.o.z;
                    wp.dx[e] = r.d.x; wp.dy[e] = r.d.y; wp.dz[e] = r.d.z;
                    wp.tr[e] = thr.x; wp.tg[e] = thr.y; wp.tb[e] = thr.z;
                    wp.seg[e] = (uint16_t)seg;
                }
            }
            // radiance of the paths that ended: per pixel slot, summed in lane order by the lowest lane of each group
            // (deterministic, no atomics: one read-modify-write per slot)
            const unsigned ended_mask = __ballot_sync(0xffffffffu, ended);
            if (ended_mask)
            {
                const unsigned group = __match_any_sync(0xffffffffu, ended ? slot : 0xffffffffu) & ended_mask;
                const bool leader = ended && lane == (uint32_t)(__ffs(group) - 1);
                // every lane walks the members of ITS group (shuffles need all lanes): at most a few iterations
                unsigned rest = ended ? group : 0u;
                float sr = 0.0f, sg = 0.0f, sb = 0.0f;
                while (__any_sync(0xffffffffu, rest != 0u))
                {
                    const int src = rest ? __ffs(rest) - 1 : 0;
                    const float vr = __shfl_sync(0xffffffffu, rad.x, src), vg = __shfl_sync(0xffffffffu, rad.y, src), vb = __shfl_sync(0xffffffffu, rad.z, src);
                    if (rest)
                    {
                        sr = __fadd_rn(sr, vr); sg = __fadd_rn(sg, vg); sb = __fadd_rn(sb, vb);
                        rest &= rest - 1u;
                    }
                }
                if (leader)
                {
                    wp.acc_r[slot] = __fadd_rn(wp.acc_r[slot], sr);
                    wp.acc_g[slot] = __fadd_rn(wp.acc_g[slot], sg);
                    wp.acc_b[slot] = __fadd_rn(wp.acc_b[slot], sb);
                }
                // re-arm the ended records with the next work items
                const uint32_t n_ended = __popc(ended_mask);
                const uint32_t n_new = min(n_ended, total_work - next_work);
                const uint32_t erank = __popc(ended_mask & lanemask_lt());
                rearm = ended && erank < n_new;
                if (rearm)
                    arm(e, next_work + erank);
                next_work += n_new;
            }
            // survivors and re-armed records -> READY
            const bool to_ready = (lane < n && !ended) || rearm;
            const unsigned rmask = __ballot_sync(0xffffffffu, to_ready);
            if (to_ready)
                wp.ready_q[(ready_head + ready_count + __popc(rmask & lanemask_lt())) % POOL_P] = (uint8_t)e;
            ready_count += __popc(rmask);
            pend_head = (pend_head + n) % POOL_P;
            pend_count -= n;
            __syncwarp();
        }
        else if (!flying && !ready_count && !pend_count)
            break;
    }

    if (in_tile)
    {
        const size_t idx = (size_t)(py * p.width + px);
        float4 acc = make_float4(wp.acc_r[lane], wp.acc_g[lane], wp.acc_b[lane], (float)n_samples);
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

} // namespace rtcu_dev
