// wavefront.cuh -- the wavefront pipeline of the path-tracing hot path (RTCU_PIPE_WAVEFRONT):
//
//   k_wf_generate    primary rays of one wave (all tile pixels x S consecutive samples) into the ray queue
//   k_wf_intersect   one thread per queued ray: closest hit (shared-memory sweep or BVH) -> hit queue, and the ray index is
//                    appended to the list of its scatter kind (miss / lambert / metal / dielectric): material sorting
//   k_wf_shade       walks the four lists back to back, so a warp shades one kind: sky on a miss, one scatter event on a
//                    hit; surviving paths are appended to the other ray queue (stream compaction: warp ballot + popc
//                    prefix, one atomic per warp)
//   k_wf_advance     swaps the queues' counters between bounces (device side, no host round trip)
//   k_wf_accumulate  adds the wave's per-sample radiances to the frame sum in sample order
//
// Queues are SoA float4 arrays in HBM: q_o = {origin, pixel}, q_d = {direction, slot << 16 | segment}, q_thr =
// {throughput}, q_hit = {t, primitive}.  Each path ends exactly once and then writes its radiance to rad[slot][pixel]
// (unique writer), and k_wf_accumulate adds the slots in ascending order, so the per-pixel sum has the same order as the
// megakernel's (and the reference's, mg_ray_tracer.cpp:187-194) whatever order the queues end up in.
#pragma once
#include "kernels.cuh"

namespace rtcu_dev {

struct WfQueues {
    float4* q_o[2];
    float4* q_d[2];
    float4* q_thr[2];
    uint2* q_hit;
    uint32_t* list[4];     // ray indices by kind
    float4* rad;           // [slot][tile pixel]
    uint32_t* counts;      // [0] n_in, [1] n_out, [2..5] kind counts
    uint32_t capacity;
};

struct WfWave {
    uint32_t sample0;      // first global sample index of the wave
    uint32_t n_slots;      // samples in this wave
    uint32_t tile_w, tile_h;
    int cur;               // which queue is the input
};

enum { WF_MISS = 0, WF_LAMBERT = 1, WF_METAL = 2, WF_DIELECTRIC = 3 };

__global__ void __launch_bounds__(256) k_wf_generate(const RenderParams p, const WfQueues q, const WfWave w)
{
    const uint32_t npix = w.tile_w * w.tile_h, n = npix * w.n_slots;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        // consecutive threads take the pixels of an 8x4 patch (same mapping as the megakernel) of one slot
        const uint32_t slot = i / npix, k = i - slot * npix;
        const uint32_t patches_x = (w.tile_w + 7u) >> 3;
        const uint32_t patch = k >> 5, in = k & 31u;
        uint32_t tx = (patch % patches_x) * 8u + (in & 7u), ty = (patch / patches_x) * 4u + (in >> 3);
        // tiles whose size is not a multiple of the patch: fall back to row-major for the whole wave
        if ((w.tile_w & 7u) || (w.tile_h & 3u))
        {
            tx = k % w.tile_w;
            ty = k / w.tile_w;
        }
        const uint32_t px = p.tile_x0 + tx, py = p.tile_y0 + ty;
        RngKey key;
        key.ks = &p.rk;
        key.pixel = py * p.width + px;
        key.sample = w.sample0 + slot;
        const Ray r = generate(p.cam, key, px, py);
        q.q_o[0][i] = make_float4(r.o.x, r.o.y, r.o.z, __uint_as_float(key.pixel));
        q.q_d[0][i] = make_float4(r.d.x, r.d.y, r.d.z, __uint_as_float(slot << 16));
        q.q_thr[0][i] = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
    {
        q.counts[0] = n;
        q.counts[1] = 0;
        q.counts[2] = q.counts[3] = q.counts[4] = q.counts[5] = 0;
    }
}

// appends `value` to list[kind] for every lane whose `kind` matches: one atomic per (warp, kind) actually present
__device__ __forceinline__ void wf_append_sorted(const WfQueues& q, int kind, uint32_t value, bool valid)
{
    const uint32_t lane = threadIdx.x & 31u;
#pragma unroll
    for (int k = 0; k < 4; k++)
    {
        const unsigned m = __ballot_sync(0xffffffffu, valid && kind == k);
        if (m)
        {
            uint32_t base = 0;
            if (lane == (uint32_t)(__ffs(m) - 1))
                base = atomicAdd(q.counts + 2 + k, (uint32_t)__popc(m));
            base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
            if (valid && kind == k)
                q.list[k][base + __popc(m & ((1u << lane) - 1u))] = value;
        }
    }
}

template <bool STAGE, bool BVH>
__global__ void __launch_bounds__(256) k_wf_intersect(const SceneDev sc, const RenderParams p, const WfQueues q, const int cur)
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
    const uint32_t n = q.counts[0];
    BvhStats bst;
    bst.nodes = 0;
    bst.tests = 0;
    // whole warps iterate together (the sorted append is a warp-collective)
    const uint32_t n_round = (n + 31u) & ~31u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x)
    {
        const bool valid = i < n;
        int kind = WF_MISS;
        if (valid)
        {
            const float4 o4 = q.q_o[cur][i], d4 = q.q_d[cur][i];
            Ray r;
            r.o = v3(o4.x, o4.y, o4.z);
            r.d = v3(d4.x, d4.y, d4.z);
            const Hit h = BVH ? closest_hit_bvh(sc, s_pl, r, bst) : closest_hit_linear(s_sph, sc.n_spheres, s_pl, sc.n_planes, r);
            q.q_hit[i] = make_uint2(__float_as_uint(h.t), h.prim);
            if (h.prim != RTCU_PRIM_MISS)
            {
                const uint32_t type = __ldg(&sc.materials[hit_material(sc, h)].type);
                kind = 1 + scatter_kind(p.mode, type);
            }
        }
        wf_append_sorted(q, kind, i, valid);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        atomicAdd(p.counters, (unsigned long long)n); // one segment per queued ray
    if (BVH)
    {
        unsigned long long nodes = bst.nodes, tests = bst.tests;
        for (int off = 16; off > 0; off >>= 1)
        {
            nodes += __shfl_down_sync(0xffffffffu, nodes, off);
            tests += __shfl_down_sync(0xffffffffu, tests, off);
        }
        if ((threadIdx.x & 31u) == 0 && nodes)
        {
            atomicAdd(p.counters + 1, nodes);
            atomicAdd(p.counters + 2, tests);
        }
    }
}

__global__ void __launch_bounds__(256) k_wf_shade(const SceneDev sc, const RenderParams p, const WfQueues q, const WfWave w)
{
    const uint32_t c0 = q.counts[2], c1 = q.counts[3], c2 = q.counts[4], c3 = q.counts[5];
    const uint32_t total = c0 + c1 + c2 + c3;
    const uint32_t npix = w.tile_w * w.tile_h;
    const int cur = w.cur, nxt = 1 - w.cur;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n_round = (total + 31u) & ~31u;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n_round; j += gridDim.x * blockDim.x)
    {
        bool survives = false;
        float4 o4 = make_float4(0, 0, 0, 0), d4 = o4, t4 = o4;
        if (j < total)
        {
            // the four kind lists back to back: a warp (almost always) shades a single kind
            const uint32_t i = j < c0 ? q.list[0][j] : (j < c0 + c1 ? q.list[1][j - c0] : (j < c0 + c1 + c2 ? q.list[2][j - c0 - c1] : q.list[3][j - c0 - c1 - c2]));
            o4 = q.q_o[cur][i];
            d4 = q.q_d[cur][i];
            t4 = q.q_thr[cur][i];
            const uint2 hh = q.q_hit[i];
            Hit h;
            h.t = __uint_as_float(hh.x);
            h.prim = hh.y;
            const uint32_t pixel = __float_as_uint(o4.w), meta = __float_as_uint(d4.w);
            const uint32_t slot = meta >> 16;
            uint32_t seg = meta & 0xffffu;
            RngKey key;
            key.ks = &p.rk;
            key.pixel = pixel;
            key.sample = w.sample0 + slot;
            Ray ray;
            ray.o = v3(o4.x, o4.y, o4.z);
            ray.d = v3(d4.x, d4.y, d4.z);
            V3 thr = v3(t4.x, t4.y, t4.z);
            V3 rad = v3(0.0f, 0.0f, 0.0f);
            // shade_segment adds thr * sky to `rad` on a miss (rad starts at 0: 0 + x == x exactly)
            const bool ended = shade_segment<true>(sc, p, sc.pairs, sc.planes, key, ray, thr, rad, seg, h);
            if (ended)
            {
                const uint32_t px = pixel % p.width - p.tile_x0, py = pixel / p.width - p.tile_y0;
                q.rad[(size_t)slot * npix + py * w.tile_w + px] = make_float4(rad.x, rad.y, rad.z, 1.0f);
            }
            else
            {
                survives = true;
                o4 = make_float4(ray.o.x, ray.o.y, ray.o.z, o4.w);
                d4 = make_float4(ray.d.x, ray.d.y, ray.d.z, __uint_as_float((slot << 16) | seg));
                t4 = make_float4(thr.x, thr.y, thr.z, 0.0f);
            }
        }
        // stream compaction of the survivors into the other queue
        const unsigned m = __ballot_sync(0xffffffffu, survives);
        if (m)
        {
            uint32_t base = 0;
            if (lane == (uint32_t)(__ffs(m) - 1))
                base = atomicAdd(q.counts + 1, (uint32_t)__popc(m));
            base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
            if (survives)
            {
                const uint32_t dst = base + __popc(m & ((1u << lane) - 1u));
                q.q_o[nxt][dst] = o4;
                q.q_d[nxt][dst] = d4;
                q.q_thr[nxt][dst] = t4;
            }
        }
    }
}

// between bounces: the output queue becomes the input queue, kind lists are emptied
__global__ void k_wf_advance(const WfQueues q)
{
    if (threadIdx.x == 0)
    {
        q.counts[0] = q.counts[1];
        q.counts[1] = 0;
        q.counts[2] = q.counts[3] = q.counts[4] = q.counts[5] = 0;
    }
}

// frame_sum[pixel] (+)= rad[0][pixel] + rad[1][pixel] + ... in ascending sample order; `first` starts the sum at 0
__global__ void __launch_bounds__(256) k_wf_accumulate(const RenderParams p, const WfQueues q, const WfWave w, float4* __restrict__ frame_sum, int first)
{
    const uint32_t npix = w.tile_w * w.tile_h;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < npix; k += gridDim.x * blockDim.x)
    {
        const uint32_t tx = k % w.tile_w, ty = k / w.tile_w;
        const size_t idx = (size_t)(p.tile_y0 + ty) * p.width + p.tile_x0 + tx;
        float4 acc = first ? make_float4(0.0f, 0.0f, 0.0f, 0.0f) : frame_sum[idx];
        for (uint32_t s = 0; s < w.n_slots; s++)
        {
            const float4 r = q.rad[(size_t)s * npix + k];
            acc.x = __fadd_rn(acc.x, r.x); acc.y = __fadd_rn(acc.y, r.y); acc.z = __fadd_rn(acc.z, r.z);
            acc.w = __fadd_rn(acc.w, 1.0f);
        }
        frame_sum[idx] = acc;
    }
}

// frame_sum -> accum (optionally on top of the previous contents) and the packed image, over the tile
__global__ void __launch_bounds__(256) k_wf_finish(const RenderParams p, const float4* __restrict__ frame_sum, uint32_t tile_w, uint32_t tile_h)
{
    const uint32_t npix = tile_w * tile_h;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < npix; k += gridDim.x * blockDim.x)
    {
        const uint32_t tx = k % tile_w, ty = k / tile_w;
        const size_t idx = (size_t)(p.tile_y0 + ty) * p.width + p.tile_x0 + tx;
        float4 acc = frame_sum[idx];
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
}

} // namespace rtcu_dev
