// rtcu.cu -- host side of the C ABI declared in include/rtcu.h (context, scene upload, launches).
// No CPU fallback: every compute entry point needs a CUDA device and says so when there is none.
#include "../../include/rtcu.h"
#include "kernels.cuh"
#include "wavefront.cuh"
#include "pool.cuh"
#include "raster.cuh"
#include "bvh.h"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cmath>
#include <cstring>
#include <exception>
#include <new>
#include <vector>

using namespace rtcu_dev;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

// C++ exceptions must not cross the C ABI: the entry points that allocate host memory (scene upload, the host BVH hooks) run
// their body under this guard and report a failed allocation / a failed thread start as an error code
template <class F>
int guarded(const char* what, F&& body) noexcept
{
    try
    {
        return body();
    }
    catch (const std::bad_alloc&)
    {
        return fail(RTCU_ERR_NOMEM, "%s: out of host memory", what);
    }
    catch (const std::exception& e)
    {
        return fail(RTCU_ERR_STATE, "%s: %s", what, e.what());
    }
}

#define CU(call)                                                                                                       \
    do                                                                                                                 \
    {                                                                                                                  \
        const cudaError_t e_ = (call);                                                                                 \
        if (e_ != cudaSuccess)                                                                                         \
            return fail(RTCU_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);    \
    } while (0)

constexpr size_t MAX_STAGE_BYTES = 200 * 1024; // dynamic shared memory budget for staged primitives
#ifndef RTCU_DEFAULT_TRAV
#define RTCU_DEFAULT_TRAV 0 // traversal variant of the direct-mode BVH kernel (kernels.cuh, closest_hit_bvh); RTCU_BVH_TRAV overrides
#endif

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0; // elements
    cudaError_t reserve(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        const cudaError_t e = cudaMalloc(&p, n * sizeof(T));
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

template <typename T>
struct PinnedBuf {
    T* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        const cudaError_t e = cudaMallocHost(&p, n * sizeof(T));
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release()
    {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

bool is_device_accessible_host(const void* ptr)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess)
    {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

} // namespace

// Experiment knobs (environment variables, DESIGN.md "Knobs"): read ONCE when the context is created and again only when the
// caller asks (rtcu_reload_env), never on the launch path.
struct Knobs {
    int regen_threshold = 0;   // RTCU_REGEN_THRESHOLD: 1..32, 0 = default
    int straggler_budget = -1; // RTCU_STRAGGLER_BUDGET: multiple of the call's samples, -1 = default
    int flat_loop = -1;        // RTCU_FLAT_LOOP: 0 / 1, -1 = by primitive count
    size_t wf_rays = 8u << 20; // RTCU_WF_RAYS: rays per wave of the wavefront pipeline
    bool pool = false;         // RTCU_BVH_KERNEL=pool
    bool direct = true;        // RTCU_BVH_DIRECT=0 disables the lanes-share-a-pixel BVH path
    int tile_order = -1;       // RTCU_TILE_ORDER: 0 row-major, 1 sorted, -1 measured per view
    bool zero_copy = true;     // RTCU_ZERO_COPY=0: always stage the image
    bool register_output = false; // rtcu_set_output_pinning / RTCU_REGISTER_OUTPUT=1: page-lock the caller's pageable image
    uint32_t bvh_threshold = 32; // RTCU_BVH_THRESHOLD
    int bvh_trav = -1;         // RTCU_BVH_TRAV: traversal variant of the BVH kernels (kernels.cuh, closest_hit_bvh): 0 = leaves tested
                               // inside the node visit (default, measured faster), 1 = deferred leaves, 2 = 1 + top levels in shared memory
    int bvh_lanes = 0;         // RTCU_BVH_LANES: lanes per pixel in direct mode (8 / 16 / 32), 0 = by sample count
    int scan_direct = -1;      // RTCU_SCAN_DIRECT: scan (non-BVH) scenes rendered with G lanes sharing a pixel: G = 2 / 4 / 8 / 16 always,
                               // 0 never, default: by frame size and sample count (scan_direct_lanes)
    bool zero_copy_direct = true; // RTCU_ZERO_COPY_DIRECT=0: no zero-copy output from the lanes-share-a-pixel kernels (their stores cross
                                  // PCIe 8 or 16 bytes at a time; measured, that still beats a copy after the frame: C1 e2e 0.483 -> 0.438 ms)
    bool scan_nested = true;   // RTCU_SCAN_NESTED=0: that kernel with samples claimed by ballot rank instead of a static split and a plain loop
    int bvh_run_pixels = 8;    // RTCU_BVH_RUN: pixels per lane-group run (4 / 8)
    bool bvh_runs = false;     // RTCU_BVH_RUNS=1: k_render_runs (lane groups walk runs of pixels) instead of k_render_stragglers'
                               // direct mode.  Measured, not default: C3 33.1 vs 30.5 ms, C4 79.6 vs 80.5 ms
    int bvh_minb = 8;          // RTCU_BVH_MINB: 6 / 7 / 8 CTAs per SM for the beam kernel (80 / 72 / 64 registers)
    int bvh_beam = -1;         // RTCU_BVH_BEAM=0: no patch beams (every primary ray traverses)
    bool scan_fixed_pairs = true; // RTCU_SCAN_FIXED_PAIRS=0: scan scenes of at most 8 spheres use the generic pair loop too
    int scan_minb = 0;            // RTCU_SCAN_MINB=6 / 8: CTAs per SM (80 / 64 registers) of the fixed-pair kernels, 0 = by pair count
    void load()
    {
        *this = Knobs{};
        if (const char* e = getenv("RTCU_REGEN_THRESHOLD")) { const int v = atoi(e); if (v >= 1 && v <= 32) regen_threshold = v; }
        if (const char* e = getenv("RTCU_STRAGGLER_BUDGET")) straggler_budget = atoi(e) < 0 ? 0 : atoi(e);
        if (const char* e = getenv("RTCU_FLAT_LOOP")) flat_loop = atoi(e) != 0;
        if (const char* e = getenv("RTCU_WF_RAYS")) wf_rays = strtoull(e, nullptr, 10);
        if (const char* e = getenv("RTCU_BVH_KERNEL")) pool = strcmp(e, "pool") == 0;
        if (const char* e = getenv("RTCU_BVH_DIRECT")) direct = e[0] != '0';
        if (const char* e = getenv("RTCU_TILE_ORDER")) tile_order = e[0] == '0' ? 0 : e[0] == '1' ? 1 : -1;
        if (const char* e = getenv("RTCU_ZERO_COPY")) zero_copy = e[0] != '0';
        if (const char* e = getenv("RTCU_REGISTER_OUTPUT")) register_output = e[0] == '1';
        if (const char* e = getenv("RTCU_BVH_THRESHOLD")) bvh_threshold = (uint32_t)strtoul(e, nullptr, 10);
        if (const char* e = getenv("RTCU_BVH_TRAV")) bvh_trav = atoi(e);
        if (const char* e = getenv("RTCU_BVH_LANES")) { const int v = atoi(e); if (v == 2 || v == 4 || v == 8 || v == 16 || v == 32) bvh_lanes = v; }
        if (const char* e = getenv("RTCU_BVH_RUNS")) bvh_runs = e[0] != '0';
        if (const char* e = getenv("RTCU_SCAN_NESTED")) scan_nested = e[0] != '0';
        if (const char* e = getenv("RTCU_ZERO_COPY_DIRECT")) zero_copy_direct = e[0] != '0';
        if (const char* e = getenv("RTCU_SCAN_DIRECT")) { const int v = atoi(e); scan_direct = (v == 2 || v == 4 || v == 8 || v == 16) ? v : (e[0] == '0' ? 0 : -1); }
        if (const char* e = getenv("RTCU_BVH_RUN")) { const int v = atoi(e); if (v == 4 || v == 8) bvh_run_pixels = v; }
        if (const char* e = getenv("RTCU_BVH_MINB")) { const int v = atoi(e); if (v >= 6 && v <= 8) bvh_minb = v; }
        if (const char* e = getenv("RTCU_BVH_BEAM")) bvh_beam = atoi(e) < 0 ? 0 : atoi(e);
        if (const char* e = getenv("RTCU_SCAN_FIXED_PAIRS")) scan_fixed_pairs = e[0] != '0';
        if (const char* e = getenv("RTCU_SCAN_MINB")) { const int v = atoi(e); scan_minb = (v == 6 || v == 8) ? v : 0; }
    }
};

// the caller's pageable image, page-locked and mapped once per (pointer, size): see ensure_registered
struct HostReg {
    void* ptr = nullptr;
    size_t bytes = 0;
    void* dev = nullptr; // device-side alias of ptr
    bool ok = false;     // false: registration failed for (ptr, bytes) -- do not retry every frame
};

struct rtcu_ctx {
    int device = 0;
    Knobs knobs;
    HostReg out_reg;
    // one render in flight per context: launches share the counters, the straggler queue and the tile-cost map, so a launch on
    // another stream first waits for the previous one (ev_launch is recorded after every launch)
    cudaEvent_t ev_launch = nullptr;
    cudaStream_t launch_stream = nullptr;
    bool launch_pending = false;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[6] = {};
    int sm_count = 0;

    // scene (device)
    DevBuf<unsigned char> scene_arena;       // every scene column and the BVH, each at a 256-byte aligned offset
    PinnedBuf<unsigned char> h_scene_arena;  // its staging copy: one H2D transfer per upload
    RasterScene raster = {};
    DevBuf<uint32_t> raster_prim;
    DevBuf<float> raster_depth;
    PinnedBuf<uint32_t> h_raster_prim;
    PinnedBuf<float> h_raster_depth;
    bool have_bvh = false;
    uint32_t bvh_depth = 0;
    float ms_bvh_build = 0.0f;
    SceneDev scene = {};
    bool have_scene = false;

    // frame buffers owned by the host entry points
    uint32_t accum_w = 0, accum_h = 0; // image size of the sums in `accum` (RTCU_FLAG_ACCUMULATE adds onto them)
    DevBuf<float4> accum;
    DevBuf<uint32_t> rgba8;
    PinnedBuf<uint32_t> h_rgba8;
    PinnedBuf<float> h_accum;
    DevBuf<unsigned long long> counters;
    DevBuf<uint32_t> tile_cost, tile_order; // longest-tile-first scheduling of the megakernel (launch_render)
    bool tile_hist_valid = false;           // tile_cost holds the per-tile segments of a frame of tile_hist_view
    rtcu_view tile_hist_view = {};
    cudaStream_t tile_hist_stream = nullptr; // ... recorded on this stream (the sort must be ordered after that frame)
    // which issue order is faster for this view is measured, not guessed: frame 1 row-major, frame 2 sorted, then the better
    int tile_phase = 0;                      // 0: time row-major, 1: time sorted, 2: decided
    int tile_pending = -1;                   // phase whose timing events are in flight
    int tile_sorted_wins = 0;
    float tile_ms[2] = { 0.0f, 0.0f };
    cudaEvent_t tile_ev[2] = {};
    DevBuf<float> run_scratch;           // k_render_runs: the lane groups' parked partial sums
    DevBuf<BeamList> beam_lists;         // per 8x4 patch of the last direct-mode frame: candidate leaves of its primary rays
    DevBuf<uint2> stragglers;            // straggler queue of the last render (width*height entries)
    DevBuf<unsigned int> straggler_count;
    PinnedBuf<unsigned long long> h_counters;

    // wavefront pipeline queues (allocated on first use)
    DevBuf<float4> wf_o[2], wf_d[2], wf_thr[2], wf_rad, wf_sum;
    DevBuf<uint2> wf_hit;
    DevBuf<uint32_t> wf_list[4], wf_counts;
    PinnedBuf<uint32_t> h_wf_counts;

    std::vector<cudaEvent_t> copy_events; // per-chunk completion events of staged device -> host copies

    // scratch for the batch entry points
    DevBuf<unsigned char> scratch;

    // buffers shared between the processes of a multi-GPU job through CUDA IPC: {pointer, owned (cudaMalloc here) or opened}
    std::vector<std::pair<void*, bool>> ipc;

    rtcu_stats stats = {};
    cudaStream_t last_stream = nullptr; // stream of the last rtcu_render_device call
    bool counters_pending = false;      // its counters have not been read back yet
};

namespace {

int make_params(const rtcu_ctx* ctx, const rtcu_view* v, RenderParams& p)
{
    if (!v) return fail(RTCU_ERR_INVALID, "view is null");
    if (v->width == 0 || v->height == 0) return fail(RTCU_ERR_INVALID, "empty image");
    if ((uint64_t)v->width * v->height > 0xFFFFFFFFull) return fail(RTCU_ERR_INVALID, "image too large");
    if (v->width > (1u << 23) || v->height > (1u << 23)) return fail(RTCU_ERR_INVALID, "image side above 2^23 pixels"); // pixel coordinates stay exact floats
    if (v->tile_x0 >= v->tile_x1 || v->tile_y0 >= v->tile_y1 || v->tile_x1 > v->width || v->tile_y1 > v->height)
        return fail(RTCU_ERR_INVALID, "bad tile [%u,%u)x[%u,%u) for %ux%u", v->tile_x0, v->tile_x1, v->tile_y0, v->tile_y1, v->width, v->height);
    if (v->sample_end < v->sample_begin) return fail(RTCU_ERR_INVALID, "bad sample range");
    if (v->samples_per_pixel == 0) return fail(RTCU_ERR_INVALID, "samples_per_pixel must be >= 1");
    if (v->max_bounces == 0) return fail(RTCU_ERR_INVALID, "max_bounces must be >= 1 (scene.cpp:532 clamps to [1,1000])");
    if (v->material_mode > RTCU_MODE_SM) return fail(RTCU_ERR_INVALID, "bad material_mode %u", v->material_mode);
    (void)ctx;
    memcpy(p.cam.m, v->inv_view_proj, sizeof p.cam.m);
    p.cam.w = (float)v->width;
    p.cam.h = (float)v->height;
    {
        const volatile float rw = 1.0f / p.cam.w, rh = 1.0f / p.cam.h; // IEEE division: RN(1/W), RN(1/H) for div_by_const
        p.cam.rw = rw;
        p.cam.rh = rh;
    }
    {
        // perspective-divide reciprocals as frame constants when h.w is pixel-independent (see CameraConst)
        const float* m = v->inv_view_proj;
        p.cam.w_const = m[3] == 0.0f && m[7] == 0.0f && std::isfinite(m[11]) && std::isfinite(m[15]);
        const volatile float b3 = m[15], f3 = b3 + m[11];
        p.cam.iwn = 1.0f / b3;
        p.cam.iwf = 1.0f / f3;
    }
    p.width = v->width; p.height = v->height;
    p.tile_x0 = v->tile_x0; p.tile_y0 = v->tile_y0; p.tile_x1 = v->tile_x1; p.tile_y1 = v->tile_y1;
    p.sample_begin = v->sample_begin; p.sample_end = v->sample_end;
    p.max_bounces = v->max_bounces;
    p.mode = v->material_mode;
    p.rk = philox_keys(make_uint2((uint32_t)v->seed, (uint32_t)(v->seed >> 32)));
    p.spp_resolve = (float)v->samples_per_pixel;
    p.accumulate = 0;
    p.accum = nullptr;
    p.rgba8 = nullptr;
    p.counters = nullptr;
    p.tile_cost = nullptr;
    p.tile_order = nullptr;
    p.direct = 0;
    p.beam = nullptr;
    p.run_scratch = nullptr;
    return RTCU_OK;
}

// lanes idle before ended paths are regenerated (see k_render_mega); RTCU_REGEN_THRESHOLD overrides for tuning runs
uint32_t regen_threshold_for(const rtcu_ctx* ctx) { return ctx->knobs.regen_threshold ? (uint32_t)ctx->knobs.regen_threshold : 1u; }

// per-thread segment budget before a pixel is handed to k_render_stragglers (0 disables; RTCU_STRAGGLER_BUDGET overrides
// with a multiple of the call's samples per pixel).  BVH scenes with many samples per pixel hand over early: a warp that
// shares ONE pixel's samples traverses more coherently than 32 neighbouring pixels do, and with >= 128 samples the 32 lanes
// stay busy (C3 scene, budget 1 vs 3: -4 % at 128 spp, -9 % at 512 spp; at 64 spp it is +4 %)
uint32_t straggler_budget_for(const rtcu_ctx* ctx, uint32_t n_samples, bool bvh)
{
    int mult = bvh && n_samples >= 128 ? 1 : 3;
    if (ctx->knobs.straggler_budget >= 0) mult = ctx->knobs.straggler_budget;
    if (mult <= 0 || n_samples < 2) return 0;
    return (uint32_t)mult * n_samples + 64u;
}

// loop structure of k_render_mega: the warp-vote (flat) loop pays off once the O(N) sweep dominates a segment
bool flat_loop_for(const rtcu_ctx* ctx, uint32_t n_prims)
{
    if (ctx->knobs.flat_loop >= 0) return ctx->knobs.flat_loop != 0;
    return n_prims >= 32;
}

size_t stage_bytes(const rtcu_ctx* ctx) { return ((size_t)pair_float4_count(ctx->scene.n_spheres) + ctx->scene.n_planes) * sizeof(float4); }

// Wavefront pipeline (RTCU_PIPE_WAVEFRONT): generate / intersect(+material sort) / shade(+compact) / advance per bounce
// over HBM queues, one wave = all tile pixels x S consecutive samples.  Results are bit-identical to the megakernel's
// first pass (same paths, per-pixel sums in sample order).
int launch_wavefront(rtcu_ctx* ctx, const rtcu_view* v, RenderParams p, bool use_bvh, cudaStream_t st)
{
    const uint32_t tw = v->tile_x1 - v->tile_x0, th = v->tile_y1 - v->tile_y0, npix = tw * th;
    const size_t target = ctx->knobs.wf_rays; // rays per wave
    uint32_t slots = (uint32_t)(target / npix);
    if (slots < 1) slots = 1;
    if (slots > 4096) slots = 4096;
    const uint32_t n_samples = v->sample_end - v->sample_begin;
    if (slots > n_samples && n_samples) slots = n_samples;
    const size_t cap = (size_t)npix * slots;
    if (cap > 0x7fffffffull) return fail(RTCU_ERR_INVALID, "wave too large");
    for (int i = 0; i < 2; i++)
    {
        CU(ctx->wf_o[i].reserve(cap)); CU(ctx->wf_d[i].reserve(cap)); CU(ctx->wf_thr[i].reserve(cap));
    }
    for (auto& l : ctx->wf_list) CU(l.reserve(cap));
    CU(ctx->wf_hit.reserve(cap));
    CU(ctx->wf_rad.reserve(cap));
    CU(ctx->wf_sum.reserve((size_t)v->width * v->height));
    CU(ctx->wf_counts.reserve(8));
    CU(ctx->h_wf_counts.reserve(8));
    WfQueues q;
    for (int i = 0; i < 2; i++) { q.q_o[i] = ctx->wf_o[i].p; q.q_d[i] = ctx->wf_d[i].p; q.q_thr[i] = ctx->wf_thr[i].p; }
    for (int k = 0; k < 4; k++) q.list[k] = ctx->wf_list[k].p;
    q.q_hit = ctx->wf_hit.p;
    q.rad = ctx->wf_rad.p;
    q.counts = ctx->wf_counts.p;
    q.capacity = (uint32_t)cap;

    const unsigned blocks = (unsigned)ctx->sm_count * 8;
    const size_t sb = stage_bytes(ctx);
    const bool stage = !use_bvh && sb <= MAX_STAGE_BYTES;
    uint32_t launches = 0;
    int first = 1;
    for (uint32_t s0 = v->sample_begin; s0 < v->sample_end; s0 += slots)
    {
        WfWave w;
        w.sample0 = s0;
        w.n_slots = v->sample_end - s0 < slots ? v->sample_end - s0 : slots;
        w.tile_w = tw;
        w.tile_h = th;
        w.cur = 0;
        k_wf_generate<<<blocks, 256, 0, st>>>(p, q, w);
        launches++;
        for (uint32_t b = 0; b < v->max_bounces; b++)
        {
            w.cur = (int)(b & 1u);
            if (use_bvh) k_wf_intersect<false, true><<<blocks, 256, 0, st>>>(ctx->scene, p, q, w.cur);
            else if (stage) k_wf_intersect<true, false><<<blocks, 256, sb, st>>>(ctx->scene, p, q, w.cur);
            else k_wf_intersect<false, false><<<blocks, 256, 0, st>>>(ctx->scene, p, q, w.cur);
            k_wf_shade<<<blocks, 256, 0, st>>>(ctx->scene, p, q, w);
            k_wf_advance<<<1, 32, 0, st>>>(q);
            launches += 3;
            CU(cudaGetLastError());
            if (b >= 1 && b + 1 < v->max_bounces)
            {
                // how many paths are still alive?  (one small synchronous read per bounce; paths die fast)
                CU(cudaMemcpyAsync(ctx->h_wf_counts.p, ctx->wf_counts.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
                CU(cudaStreamSynchronize(st));
                if (ctx->h_wf_counts.p[0] == 0) break;
            }
        }
        k_wf_accumulate<<<blocks, 256, 0, st>>>(p, q, w, ctx->wf_sum.p, first);
        launches++;
        first = 0;
    }
    if (first)
    {
        WfWave w;
        w.sample0 = v->sample_begin; w.n_slots = 0; w.tile_w = tw; w.tile_h = th; w.cur = 0;
        k_wf_accumulate<<<blocks, 256, 0, st>>>(p, q, w, ctx->wf_sum.p, 1);
        launches++;
    }
    k_wf_finish<<<blocks, 256, 0, st>>>(p, ctx->wf_sum.p, tw, th);
    launches++;
    CU(cudaGetLastError());
    ctx->stats.kernel_launches = launches;
    ctx->stats.pipeline = RTCU_PIPE_WAVEFRONT;
    return RTCU_OK;
}

// Scan (non-BVH) scenes: lanes per pixel of the lanes-share-a-pixel kernel (k_render_stragglers in direct mode), or 0 for the
// thread-per-pixel grid (k_render_mega).  The grid hands a 16x8 tile to a CTA for the whole frame and cannot balance tiles whose
// cost differs several-fold unless it has very many of them; pixel-sized work items taken from an atomic cursor can, and G lanes
// sharing a pixel shorten the longest item G-fold.  Measured on B200 (tools/sweep_scan_sizes.py, profiles/r2_scan_frame_sizes.txt;
// basic.toml and dielectric.toml with both scatter tables, 320x240 .. 3840x2160, 16 .. 256 samples): up to 1280x720 the shared
// kernel wins by 1.1x .. 4x with 8 lanes (16 from 64 samples); at 1920x1080 by 1 .. 25 % with 4 lanes (8 from 64 samples); at
// 3840x2160 the two are within 2 % and the grid -- sequential per-pixel sums, row-coalesced stores -- stays.  Below 16 samples
// (progressive refinement steps): 4 lanes from 8 samples, 2 from 4 -- up to 2.7x at 800x600, even at 1920x1080 (r2_scan_low_spp.txt).
// Thresholds are in pixels per resident thread (8 CTAs x 128 threads per SM), so they follow the SM count.
int scan_direct_lanes(const rtcu_ctx* ctx, const rtcu_view* v)
{
    const uint32_t n = v->sample_end - v->sample_begin;
    const int sd = ctx->knobs.scan_direct;
    if (sd >= 0) return (sd != 0 && n >= 2u * (uint32_t)sd) ? sd : 0;
    if (n < 4) return 0;
    const uint64_t pixels = (uint64_t)(v->tile_x1 - v->tile_x0) * (v->tile_y1 - v->tile_y0), threads = 1024ull * (uint64_t)ctx->sm_count;
    if (pixels >= 27 * threads)
    {
        // 3840x2160 and up: the thread-per-pixel grid (sequential sums, row-coalesced stores) ties with the generic shared-pixel kernel,
        // but not with the fixed-pair ones of scenes of at most 8 spheres (launch_render): basic.toml / dielectric.toml sm / mg at 4K,
        // grid vs 4 lanes: 9.33 / 10.10 / 9.44 vs 8.47 / 9.87 / 9.14 ms at 64 samples, 36.3 / 40.4 / 37.3 vs 33.1 / 38.9 / 36.0 at 256; at 16
        // samples 2 lanes: 2.40 / 2.56 / 2.46 vs 2.18 / 2.51 / 2.49
        const bool fixed_pairs = ctx->knobs.scan_fixed_pairs && ctx->knobs.scan_nested && ctx->scene.n_spheres <= 8;
        return fixed_pairs && n >= 16 ? (n >= 64 ? 4 : 2) : 0;
    }
    if (n < 16) return n >= 8 ? 4 : 2; // a lane gets at least two samples
    if (pixels >= 10 * threads) return n >= 64 ? 8 : 4;
    return n >= 64 ? 16 : 8;
}

// launches the trace kernels of one view on `st`; accum/rgba8 are device pointers
// whether launch_render will take the warp-per-pixel path (k_render_stragglers in direct mode) for this view
bool uses_direct_mode(const rtcu_ctx* ctx, const rtcu_view* v)
{
    const uint32_t accel = v->flags & 0xFu, pipe = v->flags & 0xF0u;
    const bool use_bvh = accel == RTCU_ACCEL_BVH || (accel == RTCU_ACCEL_AUTO && ctx->have_bvh && ctx->scene.n_spheres >= ctx->knobs.bvh_threshold);
    if (pipe == RTCU_PIPE_WAVEFRONT) return false;
    if (!(use_bvh && ctx->have_bvh)) return scan_direct_lanes(ctx, v) != 0;
    if (ctx->knobs.pool) return false;
    const uint32_t min_samples = ctx->knobs.bvh_lanes ? 2u * (uint32_t)ctx->knobs.bvh_lanes : 4u; // a lane gets at least two samples
    if (v->sample_end - v->sample_begin < min_samples) return false;
    return ctx->knobs.direct;
}

int launch_render_on(rtcu_ctx* ctx, const rtcu_view* v, float4* d_accum, uint32_t* d_rgba8, int accumulate, cudaStream_t st);

// launches the trace kernels of one view on `st` and marks the launch, so that the next one -- on whatever stream -- is ordered
// after it (the kernels share the context's counters, straggler queue and tile-cost map)
int launch_render(rtcu_ctx* ctx, const rtcu_view* v, float4* d_accum, uint32_t* d_rgba8, int accumulate, cudaStream_t st)
{
    const int rc = launch_render_on(ctx, v, d_accum, d_rgba8, accumulate, st);
    if (rc) return rc;
    CU(cudaEventRecord(ctx->ev_launch, st));
    ctx->launch_stream = st;
    ctx->launch_pending = true;
    return RTCU_OK;
}

int launch_render_on(rtcu_ctx* ctx, const rtcu_view* v, float4* d_accum, uint32_t* d_rgba8, int accumulate, cudaStream_t st)
{
    if (!ctx->have_scene) return fail(RTCU_ERR_STATE, "rtcu_upload_scene has not been called");
    RenderParams p;
    const int rc = make_params(ctx, v, p);
    if (rc) return rc;
    const uint32_t accel = v->flags & 0xFu, pipe = v->flags & 0xF0u;
    if (accel > RTCU_ACCEL_BVH) return fail(RTCU_ERR_INVALID, "bad accel selector %u", accel);
    if (accel == RTCU_ACCEL_BVH && !ctx->have_bvh) return fail(RTCU_ERR_STATE, "no BVH for this scene (no spheres, or 4-wide tree deeper than %d)", (BVH_STACK - 2) / 3);
    if (pipe > RTCU_PIPE_WAVEFRONT) return fail(RTCU_ERR_INVALID, "bad pipeline selector %u", pipe);
    const bool use_bvh = accel == RTCU_ACCEL_BVH || (accel == RTCU_ACCEL_AUTO && ctx->have_bvh && ctx->scene.n_spheres >= ctx->knobs.bvh_threshold);
    // one render in flight per context (rtcu.h): a launch on another stream than the previous one waits for it first
    if (ctx->launch_pending && ctx->launch_stream != st) CU(cudaStreamWaitEvent(st, ctx->ev_launch, 0));
    p.accum = d_accum;
    p.rgba8 = d_rgba8;
    p.accumulate = accumulate;
    p.counters = ctx->counters.p;
    p.regen_threshold = regen_threshold_for(ctx);
    // straggler hand-off (see k_render_stragglers): budget = 3x the samples of this call + 64 segments per pixel
    const uint32_t n_samples = v->sample_end - v->sample_begin;
    p.segment_budget = straggler_budget_for(ctx, n_samples, use_bvh);
    // the queue of the second pass: one entry per pixel of the tile at most -- and none at all where no thread can hand over (no
    // budget, or the frame takes the lanes-share-a-pixel path, which has no second pass)
    if (p.segment_budget && !uses_direct_mode(ctx, v))
        CU(ctx->stragglers.reserve((size_t)(v->tile_x1 - v->tile_x0) * (v->tile_y1 - v->tile_y0)));
    p.stragglers = ctx->stragglers.p;
    p.straggler_count = ctx->straggler_count.p;
    CU(cudaMemsetAsync(ctx->straggler_count.p, 0, sizeof(unsigned int), st));

    CU(cudaMemsetAsync(ctx->counters.p, 0, 8 * sizeof(unsigned long long), st));
    ctx->stats.accel = use_bvh ? RTCU_ACCEL_BVH : RTCU_ACCEL_LINEAR;
    ctx->stats.samples = (uint64_t)(v->tile_x1 - v->tile_x0) * (v->tile_y1 - v->tile_y0) * (v->sample_end - v->sample_begin);
    if (pipe == RTCU_PIPE_WAVEFRONT)
        return launch_wavefront(ctx, v, p, use_bvh, st); // PIPE_AUTO = megakernel: measured faster on every config
    const dim3 grid((v->tile_x1 - v->tile_x0 + MEGA_TILE_W - 1) / MEGA_TILE_W, (v->tile_y1 - v->tile_y0 + MEGA_TILE_H - 1) / MEGA_TILE_H);
    const size_t sb = stage_bytes(ctx);
    const bool flat = flat_loop_for(ctx, ctx->scene.n_spheres + ctx->scene.n_planes);
    // RTCU_BVH_KERNEL=pool selects the warp-local ray pool (pool.cuh): correct and deterministic, but measured 20-30 % slower
    // (DESIGN.md), so it is not the default.
    const bool pool = use_bvh && ctx->knobs.pool;
    if (pool) p.segment_budget = 0; // work is shared by the 32 lanes of a warp: no per-thread stragglers
    auto launch_mega = [&](const RenderParams& q) {
        if (pool) k_render_pool<<<grid, 32 * POOL_WARPS, 0, st>>>(ctx->scene, q);
        else if (use_bvh && ctx->knobs.bvh_trav != 1) k_render_mega<false, true, true, 0><<<grid, MEGA_THREADS, 0, st>>>(ctx->scene, q);
        else if (use_bvh) k_render_mega<false, true, true, 1><<<grid, MEGA_THREADS, 0, st>>>(ctx->scene, q); // flat loop (nested measures the same on C3/C4)
        else if (sb > MAX_STAGE_BYTES) k_render_mega<false, true, false><<<grid, MEGA_THREADS, 0, st>>>(ctx->scene, q);
        else if (flat) k_render_mega<true, true, false><<<grid, MEGA_THREADS, sb, st>>>(ctx->scene, q);
        else k_render_mega<true, false, false><<<grid, MEGA_THREADS, sb, st>>>(ctx->scene, q);
    };
    // Tile issue order.  CTAs are issued in index order and a tile's cost varies several-fold with what it shows, so the tiles
    // that happen to be issued last decide how long the tail of the grid is.  Every frame records the path segments of each
    // tile; later frames of the same view (same matrix, size, tile, sample count, depth and material table -- the interactive
    // app's progressive refinement, or a benchmark loop) may take the tiles in descending order of that cost (k_tile_order).
    // Measured, longest-first is worth -15 % on C1, -20 % on C2 with the mg table, -7 % on C3, but +2 % on C2 and +1..4 % on
    // C4, and neither a coarser classification nor interleaving expensive and cheap tiles predicts which -- so the choice
    // is measured too: the first frame of a view runs row-major, the second sorted, both timed with events on the caller's
    // stream, and the faster order is kept until the view or the scene changes.  It is a scheduling decision only: every
    // sample is traced every frame and the image is bit-identical whatever the order.  RTCU_TILE_ORDER=0: always
    // row-major; =1: always sorted once a cost map exists.
    // BVH scenes with at least 4 samples per call skip the thread-per-pixel kernel altogether: k_render_stragglers in direct
    // mode lets 16 (or 8 / 4 / 2: every lane gets at least two samples) lanes share ONE pixel's samples.  The samples of a pixel start from (almost) the same ray and their
    // first bounces from (almost) the same point, so the warp traverses far more coherently than 32 neighbouring pixels do,
    // and pixel-sized work items leave no grid tail (C4 -19 %, C3 -6 %; at 30 spp -18 % / -8 %).  The per-pixel sum is then a
    // fixed butterfly over the lane sums instead of the sequential sum (same paths, same segment count; fp32 summation
    // order only).  RTCU_BVH_DIRECT=0 disables.  Scan scenes: see scan_direct_lanes.
    if (uses_direct_mode(ctx, v))
    {
        p.direct = 1;
        p.segment_budget = 0;
        const unsigned long long n_items = (unsigned long long)((v->tile_x1 - v->tile_x0 + 7) / 8) * ((v->tile_y1 - v->tile_y0 + 3) / 4) * 32ull;
        if (n_items > 0xFFFFFFFFull) return fail(RTCU_ERR_INVALID, "tile too large");
        RenderParams q = p; // the kernel derives the item count from the tile (8x4 patches, ragged edges skipped)
        q.tile_cost = nullptr;
        // lanes per pixel: 16 (two pixels per warp) from 32 samples -- measured equal or better than 32 lanes on one pixel even
        // at 256 samples -- and 8 (four pixels per warp) below, so that every lane gets at least two samples
        const unsigned blocks = (unsigned)ctx->sm_count * 8;
        // Patch beams (kernels.cuh, k_beam_lists): one walk per 8x4-pixel patch before the frame replaces the primary rays'
        // traversals by a scan of the patch's candidate leaves.  Needs a viewport whose perspective divide is constant (every
        // camera::viewport, camera.hpp:122-137); RTCU_BVH_BEAM=0 disables.
        if (!use_bvh)
        {
            // Each lane loops over its own share of the pixel's samples (lane l: samples l, l + G, ...), the plain loop that is fastest for
            // a handful of primitives (RTCU_SCAN_NESTED=0: lanes claim samples by ballot rank as in the BVH kernels -- 5-13 % behind on
            // basic.toml and dielectric.toml, ahead only where a few paths are very long: dielectric.toml with the mg table below
            // 1280x720).  The primitives are read through L1: staging them in shared memory as k_render_mega does measures the same
            // (C1 0.405 vs 0.406 ms).
            // Scenes of at most 8 spheres (the reference's own: basic.toml 3, dielectric.toml 7) sweep a fixed number of packed pairs
            // (closest_hit_linear<NP>: the pair array is padded to an even pair count with never-hit pairs).  Measured (B200, kernel ms,
            // generic loop / fixed pairs at 64 registers / at 80 registers and 6 CTAs per SM): C1 0.356 / 0.336 / 0.322, C2 2.656 / 2.579 /
            // 2.674, C2 with the mg table 2.606 / 2.520 / 2.492 -- two pairs take 80 registers (104 bytes of spills at 64), four pairs 64.
            const uint32_t np = !ctx->knobs.scan_fixed_pairs ? 0u : ctx->scene.n_spheres <= 4 ? 2u : ctx->scene.n_spheres <= 8 ? 4u : 0u;
            const int scan_minb = ctx->knobs.scan_minb ? ctx->knobs.scan_minb : (np == 2u ? 6 : 8);
#define RTCU_LAUNCH_SCAN(G)                                                                                                         \
    do {                                                                                                                            \
        const unsigned b6 = (unsigned)ctx->sm_count * 6;                                                                            \
        if (!ctx->knobs.scan_nested) k_render_stragglers<false, G><<<blocks, 128, 0, st>>>(ctx->scene, q);                          \
        else if (np == 2u && scan_minb == 8) k_render_stragglers<false, G, 1, false, 8, true, 2><<<blocks, 128, 0, st>>>(ctx->scene, q); \
        else if (np == 2u) k_render_stragglers<false, G, 1, false, 6, true, 2><<<b6, 128, 0, st>>>(ctx->scene, q);                  \
        else if (np == 4u && scan_minb == 8) k_render_stragglers<false, G, 1, false, 8, true, 4><<<blocks, 128, 0, st>>>(ctx->scene, q); \
        else if (np == 4u) k_render_stragglers<false, G, 1, false, 6, true, 4><<<b6, 128, 0, st>>>(ctx->scene, q);                  \
        else k_render_stragglers<false, G, 1, false, 8, true><<<blocks, 128, 0, st>>>(ctx->scene, q);                               \
    } while (0)
            switch (scan_direct_lanes(ctx, v))
            {
            case 2: RTCU_LAUNCH_SCAN(2); break;
            case 4: RTCU_LAUNCH_SCAN(4); break;
            case 8: RTCU_LAUNCH_SCAN(8); break;
            default: RTCU_LAUNCH_SCAN(16); break;
            }
            CU(cudaGetLastError());
            ctx->tile_hist_valid = false;
            ctx->stats.kernel_launches = 1;
            ctx->stats.pipeline = RTCU_PIPE_MEGAKERNEL;
            return RTCU_OK;
        }
        const bool beam = ctx->knobs.bvh_beam != 0 && p.cam.w_const;
        // lanes per pixel: 16 (two pixels per warp) from 32 samples per call, 8 (four pixels per warp) below, so that every lane gets
        // at least two samples (RTCU_BVH_LANES to measure: with the beams 16 and 8 are within 2 % of each other, 4 is 3-7 % behind)
        // Below 16 samples (progressive refinement steps): 4 lanes from 8 samples, 2 lanes from 4 -- C3 / C4 at 800x600 .. 3840x2160:
        // 4-48 % faster than the thread-per-pixel kernel with its second pass (profiles/r2_bvh_low_spp.txt).
        const int lanes = ctx->knobs.bvh_lanes ? ctx->knobs.bvh_lanes : (n_samples >= 32 ? 16 : n_samples >= 16 ? 8 : n_samples >= 8 ? 4 : 2);
        const int trav = ctx->knobs.bvh_trav >= 0 && ctx->knobs.bvh_trav <= 2 ? ctx->knobs.bvh_trav : RTCU_DEFAULT_TRAV;
        // pixel beams (kernels.cuh, beam_collect): one walk per pixel replaces the primary rays' traversals; worth it once a lane
        // traces several samples of the pixel (RTCU_BVH_BEAM: samples per lane from which beams are used, 0 = never)
        q.beam = nullptr;
        if (beam)
        {
            const uint32_t n_patches = ((v->tile_x1 - v->tile_x0 + BEAM_PATCH_W - 1) / BEAM_PATCH_W) * ((v->tile_y1 - v->tile_y0 + BEAM_PATCH_H - 1) / BEAM_PATCH_H);
            CU(ctx->beam_lists.reserve(n_patches));
            k_beam_lists<<<(n_patches + 127) / 128, 128, 0, st>>>(ctx->scene, q, ctx->beam_lists.p);
            CU(cudaGetLastError());
            q.beam = ctx->beam_lists.p;
        }
        // runs (k_render_runs): a lane group renders G pixels of a patch in one go, its lanes moving on to the next pixel as soon
        // as the current one has no unclaimed sample.  RTCU_BVH_RUNS=1 only: the default is one pixel at a time (k_render_stragglers below)
        if (ctx->knobs.bvh_runs && trav == 0 && (lanes == 16 || lanes == 8))
        {
            const unsigned run_blocks = (unsigned)ctx->sm_count * 8;
            CU(ctx->run_scratch.reserve((size_t)run_blocks * 4 * 32 * 3 * (size_t)lanes));
            q.run_scratch = ctx->run_scratch.p;
            const int run = ctx->knobs.bvh_run_pixels;
#define RTCU_LAUNCH_RUNS(G, R) (beam ? k_render_runs<G, true, R><<<run_blocks, 128, 0, st>>>(ctx->scene, q) : k_render_runs<G, false, R><<<run_blocks, 128, 0, st>>>(ctx->scene, q))
            if (lanes == 16) { if (run == 4) RTCU_LAUNCH_RUNS(16, 4); else RTCU_LAUNCH_RUNS(16, 8); }
            else { if (run == 4) RTCU_LAUNCH_RUNS(8, 4); else RTCU_LAUNCH_RUNS(8, 8); }
            CU(cudaGetLastError());
            ctx->tile_hist_valid = false;
            ctx->stats.kernel_launches = beam ? 2 : 1;
            ctx->stats.pipeline = RTCU_PIPE_MEGAKERNEL;
            return RTCU_OK;
        }
        const int minb = ctx->knobs.bvh_minb; // experiment: CTAs per SM the beam kernel is compiled for (8 = 64 registers, 6 = 80)
#define RTCU_LAUNCH_DIRECT(G, T, B, M) k_render_stragglers<true, G, T, B, M><<<(unsigned)ctx->sm_count * M, 128, 0, st>>>(ctx->scene, q)
#define RTCU_LAUNCH_BEAM(G) RTCU_LAUNCH_DIRECT(G, 0, true, 8)
        if (trav == 1 && lanes == 16) RTCU_LAUNCH_DIRECT(16, 1, false, 8); // (the traversal experiments: 16 lanes, no beams)
        else if (trav == 2 && lanes == 16) RTCU_LAUNCH_DIRECT(16, 2, false, 8);
        else if (lanes == 32) { if (beam) RTCU_LAUNCH_BEAM(32); else RTCU_LAUNCH_DIRECT(32, 0, false, 8); }
        else if (lanes == 16)
        {
            // (the register-budget experiment, RTCU_BVH_MINB: 6 / 7 CTAs per SM = 80 / 72 registers, 16 lanes only)
            if (beam && minb == 6) RTCU_LAUNCH_DIRECT(16, 0, true, 6);
            else if (beam && minb == 7) RTCU_LAUNCH_DIRECT(16, 0, true, 7);
            else if (beam) RTCU_LAUNCH_BEAM(16);
            else RTCU_LAUNCH_DIRECT(16, 0, false, 8);
        }
        else if (lanes == 8) { if (beam) RTCU_LAUNCH_BEAM(8); else RTCU_LAUNCH_DIRECT(8, 0, false, 8); }
        else if (lanes == 4) { if (beam) RTCU_LAUNCH_BEAM(4); else RTCU_LAUNCH_DIRECT(4, 0, false, 8); }
        else { if (beam) RTCU_LAUNCH_BEAM(2); else RTCU_LAUNCH_DIRECT(2, 0, false, 8); }
        (void)blocks;
        CU(cudaGetLastError());
        ctx->tile_hist_valid = false;
        ctx->stats.kernel_launches = beam ? 2 : 1;
        ctx->stats.pipeline = RTCU_PIPE_MEGAKERNEL;
        return RTCU_OK;
    }
    // thread-per-pixel BVH frames (fewer than 16 samples per call: interactive refinement) use the patch beams too: a warp of
    // k_render_mega covers exactly one 8x4 patch
    if (use_bvh && !pool && ctx->knobs.bvh_beam != 0 && p.cam.w_const)
    {
        const uint32_t n_patches = ((v->tile_x1 - v->tile_x0 + BEAM_PATCH_W - 1) / BEAM_PATCH_W) * ((v->tile_y1 - v->tile_y0 + BEAM_PATCH_H - 1) / BEAM_PATCH_H);
        CU(ctx->beam_lists.reserve(n_patches));
        k_beam_lists<<<(n_patches + 127) / 128, 128, 0, st>>>(ctx->scene, p, ctx->beam_lists.p);
        CU(cudaGetLastError());
        p.beam = ctx->beam_lists.p;
    }
    const uint32_t n_tiles = grid.x * grid.y;
    const bool lpt = !pool && n_tiles >= 2u * 8u * (uint32_t)ctx->sm_count && ctx->knobs.tile_order != 0;
    ctx->stats.kernel_launches = 1;
    bool time_this_frame = false;
    if (lpt)
    {
        const rtcu_view& h = ctx->tile_hist_view;
        const bool same_view = ctx->tile_hist_valid && ctx->tile_hist_stream == st && memcmp(h.inv_view_proj, v->inv_view_proj, sizeof h.inv_view_proj) == 0 &&
                               h.width == v->width && h.height == v->height && h.tile_x0 == v->tile_x0 && h.tile_y0 == v->tile_y0 &&
                               h.tile_x1 == v->tile_x1 && h.tile_y1 == v->tile_y1 && h.sample_end - h.sample_begin == v->sample_end - v->sample_begin &&
                               h.max_bounces == v->max_bounces && h.material_mode == v->material_mode;
        if (!same_view)
        {
            ctx->tile_phase = 0;
            ctx->tile_pending = -1;
        }
        else if (ctx->tile_pending >= 0 && cudaEventQuery(ctx->tile_ev[1]) == cudaSuccess)
        {
            float ms = 0.0f;
            if (cudaEventElapsedTime(&ms, ctx->tile_ev[0], ctx->tile_ev[1]) == cudaSuccess)
            {
                ctx->tile_ms[ctx->tile_pending] = ms;
                if (ctx->tile_pending == 1) ctx->tile_sorted_wins = ctx->tile_ms[1] < ctx->tile_ms[0];
                ctx->tile_phase = ctx->tile_pending + 1;
            }
            ctx->tile_pending = -1;
        }
        cudaGetLastError(); // cudaEventQuery reports cudaErrorNotReady through the sticky-free error slot
        // a frame is a test frame (timed) when its phase has no measurement in flight; while one is in flight -- the caller
        // queues frames without synchronising -- frames run row-major, the order that is never a regression
        const bool forced = ctx->knobs.tile_order == 1;
        time_this_frame = !forced && ctx->tile_phase < 2 && ctx->tile_pending < 0 && (ctx->tile_phase == 0 || same_view);
        const bool sorted = same_view && (forced || (time_this_frame && ctx->tile_phase == 1) || (ctx->tile_phase == 2 && ctx->tile_sorted_wins));
        CU(ctx->tile_cost.reserve(n_tiles));
        CU(ctx->tile_order.reserve(n_tiles));
        if (time_this_frame)
        {
            for (auto& e : ctx->tile_ev)
                if (!e) CU(cudaEventCreate(&e));
            CU(cudaEventRecord(ctx->tile_ev[0], st));
        }
        if (sorted)
        {
            k_tile_order<<<1, 1024, 0, st>>>(ctx->tile_cost.p, n_tiles, ctx->tile_order.p);
            CU(cudaGetLastError());
            p.tile_order = ctx->tile_order.p;
            ctx->stats.kernel_launches = 2;
        }
        CU(cudaMemsetAsync(ctx->tile_cost.p, 0, n_tiles * sizeof(uint32_t), st));
        p.tile_cost = ctx->tile_cost.p;
        ctx->tile_hist_view = *v;
        ctx->tile_hist_stream = st;
        ctx->tile_hist_valid = true;
    }
    else
        ctx->tile_hist_valid = false;
    launch_mega(p);
    CU(cudaGetLastError());
    if (p.beam) ctx->stats.kernel_launches++;
    if (p.segment_budget)
    {
        // the accumulate flag only applies to the first pass: the second adds onto what the first wrote
        const unsigned blocks = (unsigned)ctx->sm_count * 8;
        if (use_bvh) k_render_stragglers<true, 32, 0><<<blocks, 128, 0, st>>>(ctx->scene, p);
        else k_render_stragglers<false><<<blocks, 128, 0, st>>>(ctx->scene, p);
        CU(cudaGetLastError());
        ctx->stats.kernel_launches++;
    }
    if (time_this_frame)
    {
        CU(cudaEventRecord(ctx->tile_ev[1], st));
        ctx->tile_pending = ctx->tile_phase;
    }
    ctx->stats.pipeline = RTCU_PIPE_MEGAKERNEL;
    ctx->stats.accel = use_bvh ? RTCU_ACCEL_BVH : RTCU_ACCEL_LINEAR;
    ctx->stats.samples = (uint64_t)(v->tile_x1 - v->tile_x0) * (v->tile_y1 - v->tile_y0) * (v->sample_end - v->sample_begin);
    return RTCU_OK;
}

int fetch_counters(rtcu_ctx* ctx, cudaStream_t st)
{
    CU(cudaMemcpyAsync(ctx->h_counters.p, ctx->counters.p, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    ctx->stats.segments = ctx->h_counters.p[0];
    if (ctx->stats.accel == RTCU_ACCEL_BVH)
    {
        ctx->stats.node_visits = ctx->h_counters.p[1];
        ctx->stats.sphere_tests = ctx->h_counters.p[2];
    }
    else
    {
        ctx->stats.sphere_tests = ctx->stats.segments * ctx->scene.n_spheres;
        ctx->stats.node_visits = 0;
    }
    return RTCU_OK;
}

// device -> caller's host buffer; direct when the buffer is pinned/registered, staged otherwise
template <typename T>
int copy_out(rtcu_ctx* ctx, T* dst, const T* d_src, size_t n, PinnedBuf<T>& staging)
{
    if (is_device_accessible_host(dst))
    {
        CU(cudaMemcpyAsync(dst, d_src, n * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    else
    {
        // pageable destination (the reference's image buffer, image.cpp:9-13): DMA into pinned staging in chunks and
        // copy chunk k to the caller while chunk k+1 is still in flight
        CU(staging.reserve(n));
        const size_t chunk = (size_t)(1u << 20) / sizeof(T); // 1 MiB
        const size_t n_chunks = (n + chunk - 1) / chunk;
        if (ctx->copy_events.size() < n_chunks)
        {
            const size_t old = ctx->copy_events.size();
            ctx->copy_events.resize(n_chunks, nullptr);
            for (size_t k = old; k < n_chunks; k++) CU(cudaEventCreateWithFlags(&ctx->copy_events[k], cudaEventDisableTiming));
        }
        for (size_t k = 0; k < n_chunks; k++)
        {
            const size_t off = k * chunk, cnt = n - off < chunk ? n - off : chunk;
            CU(cudaMemcpyAsync(staging.p + off, d_src + off, cnt * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaEventRecord(ctx->copy_events[k], ctx->stream));
        }
        for (size_t k = 0; k < n_chunks; k++)
        {
            const size_t off = k * chunk, cnt = n - off < chunk ? n - off : chunk;
            CU(cudaEventSynchronize(ctx->copy_events[k]));
            memcpy(dst + off, staging.p + off, cnt * sizeof(T));
        }
    }
    return RTCU_OK;
}

// ---- the caller's image as the device can reach it ---------------------------------------------------------------------
// The reference's image is pageable (image.cpp:9-13, muu::aligned_alloc) and lives as long as its back buffer: the same
// pointer arrives frame after frame.  It is page-locked and mapped ONCE per (pointer, size) -- cudaHostRegister costs about
// a frame, then every later frame is written by DMA (or by the kernels themselves, zero-copy) instead of being staged through
// a bounce buffer and a host memcpy.  A registration goes stale if the application frees the buffer and the allocator maps
// new pages at the same address; nothing tells the library, so every frame is verified: two sentinel pixels with alpha 0 (no
// packed pixel has that, colour.hpp:100-106) are written through the host pointer before the launch and must have been
// overwritten when the frame is complete.  If not, the registration is dropped and the frame is delivered by the staged copy.
uint32_t* output_alias(rtcu_ctx* ctx, uint32_t* out, size_t npix, bool* ours)
{
    *ours = false;
    HostReg& r = ctx->out_reg;
    const size_t bytes = npix * sizeof(uint32_t);
    if (r.ptr == out && r.bytes == bytes)
    {
        *ours = r.ok;
        return r.ok ? static_cast<uint32_t*>(r.dev) : nullptr;
    }
    // another image than the one registered here (a resize, another back buffer): drop that registration FIRST -- the new image
    // may overlap the old range, and a pointer inside a range this context registered must not be mistaken for caller-pinned
    // memory of the new size
    if (r.ptr && r.ok && cudaHostUnregister(r.ptr) != cudaSuccess) cudaGetLastError();
    r = HostReg{};
    if (is_device_accessible_host(out)) // pinned / registered / managed by the caller: the caller vouches for its extent
    {
        void* mapped = nullptr;
        if (cudaHostGetDevicePointer(&mapped, out, 0) == cudaSuccess && mapped) return static_cast<uint32_t*>(mapped);
        cudaGetLastError();
        return nullptr;
    }
    if (!ctx->knobs.register_output) return nullptr;
    r.ptr = out;
    r.bytes = bytes;
    if (cudaHostRegister(out, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped) != cudaSuccess)
    {
        cudaGetLastError();
        return nullptr; // remembered: not retried for this (pointer, size)
    }
    void* mapped = nullptr;
    if (cudaHostGetDevicePointer(&mapped, out, 0) != cudaSuccess || !mapped)
    {
        cudaGetLastError();
        cudaHostUnregister(out);
        return nullptr;
    }
    r.dev = mapped;
    r.ok = true;
    *ours = true;
    return static_cast<uint32_t*>(mapped);
}
void arm_sentinels(uint32_t* out, size_t npix)
{
    reinterpret_cast<volatile uint32_t*>(out)[0] = 0u;
    reinterpret_cast<volatile uint32_t*>(out)[npix - 1] = 0u;
}
bool sentinels_overwritten(const uint32_t* out, size_t npix)
{
    const volatile uint32_t* v = out;
    return (v[0] & 255u) == 255u && (v[npix - 1] & 255u) == 255u;
}
void drop_registration(rtcu_ctx* ctx)
{
    HostReg& r = ctx->out_reg;
    if (r.ptr && r.ok && cudaHostUnregister(r.ptr) != cudaSuccess) cudaGetLastError();
    r.ok = false; // (pointer, size) stay: not retried until the application hands over another buffer
    r.dev = nullptr;
}

// full frame, device -> the caller's image: one DMA when the image is page-locked (verified when the registration is ours),
// the staged copy otherwise
int deliver_image(rtcu_ctx* ctx, uint32_t* out, const uint32_t* d_src, size_t npix)
{
    bool ours = false;
    uint32_t* alias = output_alias(ctx, out, npix, &ours);
    if (alias)
    {
        if (ours) arm_sentinels(out, npix);
        CU(cudaMemcpyAsync(out, d_src, npix * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        if (!ours || sentinels_overwritten(out, npix)) return RTCU_OK;
        drop_registration(ctx);
    }
    return copy_out(ctx, out, d_src, npix, ctx->h_rgba8);
}

} // namespace

// ---- batch entry points: stage host arrays through one scratch allocation -----------------------
namespace {
struct Carver {
    unsigned char* base;
    size_t off = 0;
    template <typename T>
    T* take(size_t n)
    {
        off = (off + 255) & ~(size_t)255;
        T* p = reinterpret_cast<T*>(base + off);
        off += n * sizeof(T);
        return p;
    }
};
size_t padded(size_t bytes) { return ((bytes + 255) & ~(size_t)255) + 256; }
} // namespace

namespace {

// ---- device BVH: the host builder's binary tree (bvh.h) collapsed to 4-wide nodes ---------------------------------------
// A 4-wide node holds up to four child boxes: the two children of a binary node, the larger inner ones replaced by their own
// children until four are reached (greedy by surface area).  Halving the depth halves the node visits of a traversal, and a
// visit's fixed cost (loads, child ordering, stack) is paid once for four slab tests instead of twice for two each.
// Layout (8 float4 = 128 B per node): children (0,1) as centre / half-extent {c0,c1,h0,h1} for x, y, z, the same for children
// (2,3), then the four child references and the four H = h.x + h.y + h.z.  h is rounded outward so that [c - h, c + h]
// contains the builder's box; an empty slot has h = -inf and can never be hit.  Leaves: see SceneDev::leaf_blk.
struct Kid4 {
    float lo[3], hi[3];
    int32_t child;  // >= 0: binary inner node; < 0: leaf, first primitive = ~child
    uint32_t count; // primitives of a leaf (0 = empty)
};

Kid4 kid_of(const rtcu_bvh::Node& nd, int c)
{
    Kid4 k;
    k.lo[0] = nd.x[2 * c]; k.hi[0] = nd.x[2 * c + 1];
    k.lo[1] = nd.y[2 * c]; k.hi[1] = nd.y[2 * c + 1];
    k.lo[2] = nd.z[2 * c]; k.hi[2] = nd.z[2 * c + 1];
    k.child = nd.child[c];
    k.count = nd.count[c];
    return k;
}

void pack_bvh4(const rtcu_bvh::Result& bvh, const std::vector<float4>& sph, std::vector<float4>& nodes_dev, std::vector<float4>& leaf_blk,
               uint32_t& depth4, rtcu_bvh::Pool& pool)
{
    // Pass 1 (serial, structure only): breadth-first collapse.  A node's index in the 4-wide array is its position in the queue,
    // a leaf's block index the order in which the leaves are met.
    struct Wide {
        Kid4 kids[4];
        uint32_t ref[4]; // inner: index of the 4-wide child; leaf: 0x80000000 | block; empty: 0x80000000 (never hit)
        bool empty[4];
        int32_t binary;
        uint32_t depth;
    };
    std::vector<Wide> wide;
    wide.reserve(bvh.nodes.size());
    wide.push_back(Wide{});
    wide[0].binary = 0;
    wide[0].depth = 1;
    uint32_t n_leaves = 0;
    depth4 = 0;
    // Which binary nodes fold into a 4-wide node: the collapse of least SAH cost by dynamic programming (after Ylitie et al. 2017) --
    // C(e, i) = cheapest way to cover the subtree of child entry e with at most i slots of its parent's wide node: as one wide node
    // of its own (i = 1: area x node cost + the best distribution of ITS four slots over its two children), or dissolved into the
    // parent (its children share the i slots).  Measured against the greedy rule (replace the inner child of largest area by its two
    // children until four slots are filled; RTCU_BVH_COLLAPSE=greedy): C4 28 % fewer nodes, 95.8 -> 92.8 ms; C3 41.2 -> 41.0 ms.
    const char* collapse_env = getenv("RTCU_BVH_COLLAPSE");
    const bool sah_collapse = !(collapse_env && strcmp(collapse_env, "greedy") == 0);
    const size_t n_binary = bvh.nodes.size();
    std::vector<float> cost_entry, cost_split; // C[(2 * node + side) * 5 + i], D[node * 5 + j]
    std::vector<uint8_t> split_left;           // slots given to the left child by the best distribution D[node][j]
    if (sah_collapse)
    {
        constexpr float NODE_COST = 1.0f, LEAF_COST = 0.45f; // issue slots of a leaf visit relative to a node visit (~40 : 93)
        cost_entry.assign(n_binary * 10, 0.0f);
        cost_split.assign(n_binary * 5, 0.0f);
        split_left.assign(n_binary * 5, 1);
        for (size_t m = n_binary; m-- > 0;) // children have larger indices than their parent (level-order construction)
        {
            for (int side = 0; side < 2; side++)
            {
                const Kid4 e = kid_of(bvh.nodes[m], side);
                float* c = &cost_entry[(2 * m + side) * 5];
                const float dx = e.hi[0] - e.lo[0], dy = e.hi[1] - e.lo[1], dz = e.hi[2] - e.lo[2];
                const float area = (e.child < 0 && e.count == 0) ? 0.0f : dx * dy + dy * dz + dz * dx;
                if (e.child < 0)
                {
                    for (int i = 1; i <= 4; i++) c[i] = area * LEAF_COST;
                    continue;
                }
                c[1] = area * NODE_COST + cost_split[(size_t)e.child * 5 + 4];
                for (int i = 2; i <= 4; i++) c[i] = std::min(c[i - 1], cost_split[(size_t)e.child * 5 + i]);
            }
            for (int j = 2; j <= 4; j++)
            {
                float best = __builtin_inff();
                for (int k = 1; k < j; k++)
                {
                    const float v = cost_entry[(2 * m) * 5 + k] + cost_entry[(2 * m + 1) * 5 + (j - k)];
                    if (v < best) { best = v; split_left[m * 5 + j] = (uint8_t)k; }
                }
                cost_split[m * 5 + j] = best;
            }
        }
    }
    for (size_t qi = 0; qi < wide.size(); qi++)
    {
        Wide w = wide[qi];
        depth4 = std::max(depth4, w.depth);
        int nk = 2;
        if (sah_collapse)
        {
            // unfold the recorded decisions: (binary node, slots) pairs, left before right
            nk = 0;
            struct Todo { int32_t node; int slots; };
            Todo todo[8];
            int nt = 0;
            todo[nt++] = Todo{ w.binary, 4 };
            while (nt > 0)
            {
                const Todo t = todo[--nt];
                const int left = split_left[(size_t)t.node * 5 + t.slots];
                const int budget[2] = { left, t.slots - left };
                Todo defer[2];
                int nd = 0;
                for (int side = 0; side < 2; side++)
                {
                    const Kid4 e = kid_of(bvh.nodes[(size_t)t.node], side);
                    const float* c = &cost_entry[(2 * (size_t)t.node + side) * 5];
                    int i = budget[side];
                    while (i > 1 && c[i] == c[i - 1]) i--; // the smallest budget that reaches the optimum
                    if (e.child >= 0 && i >= 2) defer[nd++] = Todo{ e.child, i }; // dissolved: its children take the i slots
                    else w.kids[nk++] = e;
                }
                for (int k = nd - 1; k >= 0; k--) todo[nt++] = defer[k];
            }
        }
        else
        {
            w.kids[0] = kid_of(bvh.nodes[(size_t)w.binary], 0);
            w.kids[1] = kid_of(bvh.nodes[(size_t)w.binary], 1);
        }
        while (!sah_collapse && nk < 4)
        {
            int best = -1;
            float best_area = -1.0f;
            for (int j = 0; j < nk; j++)
            {
                if (w.kids[j].child < 0) continue;
                const float dx = w.kids[j].hi[0] - w.kids[j].lo[0], dy = w.kids[j].hi[1] - w.kids[j].lo[1], dz = w.kids[j].hi[2] - w.kids[j].lo[2];
                const float area = dx * dy + dy * dz + dz * dx;
                if (area > best_area) { best_area = area; best = j; }
            }
            if (best < 0) break;
            const rtcu_bvh::Node& nd = bvh.nodes[(size_t)w.kids[best].child];
            w.kids[best] = kid_of(nd, 0);
            w.kids[nk++] = kid_of(nd, 1);
        }
        for (int c = 0; c < 4; c++)
        {
            w.empty[c] = c >= nk || (w.kids[c].child < 0 && w.kids[c].count == 0) || !(w.kids[c].lo[0] <= w.kids[c].hi[0]);
            w.ref[c] = 0x80000000u; // an empty slot points at leaf 0 but is never hit
            if (w.empty[c]) continue;
            if (w.kids[c].child >= 0)
            {
                w.ref[c] = (uint32_t)wide.size();
                Wide child{};
                child.binary = w.kids[c].child;
                child.depth = w.depth + 1;
                wide.push_back(child);
            }
            else
                w.ref[c] = 0x80000000u | n_leaves++;
        }
        wide[qi] = w;
    }

    // Pass 2 (every node on its own, worker threads for large trees): boxes as centre / half-extent rounded outward, and a leaf as
    // one 80-byte block -- its 1-4 spheres as two packed pairs of the sweep's layout (padded with never-hit spheres), then their
    // four original indices as bit patterns
    const float ninf = -__builtin_inff();
    nodes_dev.assign(8 * wide.size(), make_float4(0.0f, 0.0f, 0.0f, 0.0f));
    leaf_blk.assign(5 * (size_t)n_leaves, make_float4(0.0f, 0.0f, 0.0f, 0.0f));
    auto fill = [&](size_t qi) {
        const Wide& w = wide[qi];
        float cen[4][3], half[4][3], hsum[4];
        for (int c = 0; c < 4; c++)
        {
            if (w.empty[c])
            {
                for (int k = 0; k < 3; k++) { cen[c][k] = 0.0f; half[c][k] = ninf; }
                hsum[c] = ninf;
                continue;
            }
            double hs = 0.0;
            for (int k = 0; k < 3; k++)
            {
                const float lo = w.kids[c].lo[k], hi = w.kids[c].hi[k];
                const float ce = (float)(0.5 * ((double)lo + (double)hi));
                const double need = std::max((double)hi - (double)ce, (double)ce - (double)lo);
                float h = (float)need;
                if ((double)h < need) h = nextafterf(h, __builtin_inff());
                cen[c][k] = ce;
                half[c][k] = h;
                hs += (double)h;
            }
            hsum[c] = (float)hs;
            if ((double)hsum[c] < hs) hsum[c] = nextafterf(hsum[c], __builtin_inff());
            if (w.kids[c].child >= 0) continue;
            const float4 never = make_float4(0.0f, 0.0f, 0.0f, ninf);
            float4 sp[4];
            uint32_t idx[4];
            for (uint32_t k = 0; k < 4; k++)
            {
                const bool real = k < w.kids[c].count;
                idx[k] = real ? bvh.order[(size_t)(~w.kids[c].child) + k] : 0x7fffffffu;
                sp[k] = real ? sph[idx[k]] : never;
            }
            float4* blk = &leaf_blk[5 * (size_t)(w.ref[c] & 0x7fffffffu)];
            for (int pr = 0; pr < 2; pr++)
            {
                blk[2 * pr] = make_float4(sp[2 * pr].x, sp[2 * pr + 1].x, sp[2 * pr].y, sp[2 * pr + 1].y);
                blk[2 * pr + 1] = make_float4(sp[2 * pr].z, sp[2 * pr + 1].z, sp[2 * pr].w, sp[2 * pr + 1].w);
            }
            memcpy(&blk[4], idx, sizeof(float4));
        }
        float4* out = &nodes_dev[8 * qi];
        for (int pr = 0; pr < 2; pr++)
            for (int k = 0; k < 3; k++)
                out[3 * pr + k] = make_float4(cen[2 * pr][k], cen[2 * pr + 1][k], half[2 * pr][k], half[2 * pr + 1][k]);
        memcpy(&out[6], w.ref, sizeof(float4));
        out[7] = make_float4(hsum[0], hsum[1], hsum[2], hsum[3]);
    };
    constexpr size_t BLOCK = 256;
    pool.run((wide.size() + BLOCK - 1) / BLOCK, wide.size() >= 4096 ? 1 : SIZE_MAX, [&](size_t blk) {
        for (size_t qi = blk * BLOCK, e = std::min(wide.size(), (blk + 1) * BLOCK); qi < e; qi++) fill(qi);
    });
    if (leaf_blk.empty()) // every reference needs a block to point at
        for (int k = 0; k < 5; k++) leaf_blk.push_back(make_float4(0.0f, 0.0f, ninf, ninf));
}

} // namespace

namespace {

// ---- rasterizer (rasterizer.cpp:22-88) ---------------------------------------------------------------------------------
int launch_rasterize(rtcu_ctx* ctx, const rtcu_view* v, uint32_t* d_rgba8, uint32_t* d_prim, float* d_depth, cudaStream_t st)
{
    if (!ctx->have_scene) return fail(RTCU_ERR_STATE, "rtcu_upload_scene has not been called");
    if (v->width == 0 || v->height == 0) return fail(RTCU_ERR_INVALID, "empty image");
    if ((uint64_t)v->width * v->height > 0xFFFFFFFFull) return fail(RTCU_ERR_INVALID, "image too large");
    if (v->tile_x0 >= v->tile_x1 || v->tile_y0 >= v->tile_y1 || v->tile_x1 > v->width || v->tile_y1 > v->height)
        return fail(RTCU_ERR_INVALID, "bad tile [%u,%u)x[%u,%u) for %ux%u", v->tile_x0, v->tile_x1, v->tile_y0, v->tile_y1, v->width, v->height);
    rtcu_view pv = *v; // the preview ignores the sampling fields; give make_params values it accepts
    pv.samples_per_pixel = pv.max_bounces = 1;
    pv.sample_begin = 0;
    pv.sample_end = 1;
    pv.material_mode = RTCU_MODE_MG;
    RenderParams rp;
    const int rc = make_params(ctx, &pv, rp);
    if (rc) return rc;
    RasterParams p;
    p.cam = rp.cam;
    p.width = v->width; p.height = v->height;
    p.tile_x0 = v->tile_x0; p.tile_y0 = v->tile_y0; p.tile_x1 = v->tile_x1; p.tile_y1 = v->tile_y1;
    p.rgba8 = d_rgba8;
    p.prim = d_prim;
    p.depth = d_depth;
    const dim3 grid((v->tile_x1 - v->tile_x0 + RASTER_TILE_W - 1) / RASTER_TILE_W, (v->tile_y1 - v->tile_y0 + RASTER_TILE_H - 1) / RASTER_TILE_H);
    const uint32_t accel = v->flags & 0xFu;
    if (accel > RTCU_ACCEL_BVH) return fail(RTCU_ERR_INVALID, "bad accel selector %u", accel);
    if (accel == RTCU_ACCEL_BVH && !ctx->have_bvh) return fail(RTCU_ERR_STATE, "no BVH for this scene (no spheres, or 4-wide tree deeper than %d)", (BVH_STACK - 2) / 3);
    const bool use_bvh = accel == RTCU_ACCEL_BVH || (accel == RTCU_ACCEL_AUTO && ctx->have_bvh && ctx->scene.n_spheres >= ctx->knobs.bvh_threshold);
    if (use_bvh) k_rasterize<true><<<grid, RASTER_TILE_W * RASTER_TILE_H, 0, st>>>(ctx->raster, ctx->scene, p);
    else k_rasterize<false><<<grid, RASTER_TILE_W * RASTER_TILE_H, 0, st>>>(ctx->raster, ctx->scene, p);
    CU(cudaGetLastError());
    ctx->stats.kernel_launches = 1;
    ctx->stats.pipeline = RTCU_PIPE_MEGAKERNEL;
    ctx->stats.accel = use_bvh ? RTCU_ACCEL_BVH : RTCU_ACCEL_LINEAR;
    ctx->stats.samples = (uint64_t)(v->tile_x1 - v->tile_x0) * (v->tile_y1 - v->tile_y0);
    ctx->stats.segments = ctx->stats.samples;
    ctx->stats.sphere_tests = ctx->stats.samples * ctx->raster.n_spheres;
    ctx->stats.node_visits = 0;
    ctx->counters_pending = false;
    return RTCU_OK;
}

// tile rows of a width*height device plane -> the caller's host plane (staged through `staging`)
template <typename T>
int copy_tile_out(rtcu_ctx* ctx, const rtcu_view* v, T* dst, const T* d_src, PinnedBuf<T>& staging)
{
    const size_t npix = (size_t)v->width * v->height;
    const bool full = v->tile_x0 == 0 && v->tile_y0 == 0 && v->tile_x1 == v->width && v->tile_y1 == v->height;
    if (full) return copy_out(ctx, dst, d_src, npix, staging);
    const size_t off = (size_t)v->tile_y0 * v->width, cnt = (size_t)(v->tile_y1 - v->tile_y0) * v->width;
    CU(staging.reserve(npix));
    CU(cudaMemcpyAsync(staging.p + off, d_src + off, cnt * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (uint32_t y = v->tile_y0; y < v->tile_y1; y++)
        memcpy(dst + (size_t)y * v->width + v->tile_x0, staging.p + (size_t)y * v->width + v->tile_x0, (size_t)(v->tile_x1 - v->tile_x0) * sizeof(T));
    return RTCU_OK;
}

} // namespace

extern "C" {

int rtcu_abi_version(void) { return RTCU_ABI_VERSION; }

const char* rtcu_last_error(void) { return g_err; }

uint32_t rtcu_bvh_threshold(void)
{
    Knobs k;
    k.load();
    return k.bvh_threshold; // 32: measured crossover (profiles/bvh_crossover_r1.jsonl): BVH is ahead from ~24 spheres, 20 % at 32, 2x at 64
}

int rtcu_reload_env(rtcu_ctx* ctx)
{
    if (!ctx) return fail(RTCU_ERR_INVALID, "null context");
    const bool pin = ctx->knobs.register_output;
    ctx->knobs.load();
    ctx->knobs.register_output = ctx->knobs.register_output || pin; // (a caller's rtcu_set_output_pinning survives)
    return RTCU_OK;
}

int rtcu_set_output_pinning(rtcu_ctx* ctx, int enable)
{
    if (!ctx) return fail(RTCU_ERR_INVALID, "null context");
    CU(cudaSetDevice(ctx->device));
    ctx->knobs.register_output = enable != 0;
    if (!enable)
    {
        CU(cudaStreamSynchronize(ctx->stream));
        drop_registration(ctx);
        ctx->out_reg = HostReg{};
    }
    return RTCU_OK;
}

int rtcu_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess)
    {
        cudaGetLastError();
        return 0;
    }
    return n;
}

rtcu_ctx* rtcu_create(int device)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
    {
        fail(RTCU_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        cudaGetLastError();
        return nullptr;
    }
    if (device < 0 || device >= n)
    {
        fail(RTCU_ERR_INVALID, "device %d out of range [0,%d)", device, n);
        return nullptr;
    }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
    {
        fail(RTCU_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
        return nullptr;
    }
    if (prop.major != 10)
    {
        fail(RTCU_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
        return nullptr;
    }
    rtcu_ctx* ctx = new (std::nothrow) rtcu_ctx;
    if (!ctx)
    {
        fail(RTCU_ERR_INVALID, "out of host memory");
        return nullptr;
    }
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    bool ok = cudaSetDevice(device) == cudaSuccess && cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; ok && i < 6; i++)
        ok = cudaEventCreate(&ctx->ev[i]) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&ctx->ev_launch, cudaEventDisableTiming) == cudaSuccess;
    ctx->knobs.load();
    ok = ok && ctx->counters.reserve(8) == cudaSuccess && ctx->h_counters.reserve(4) == cudaSuccess && ctx->straggler_count.reserve(1) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(k_render_mega<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MAX_STAGE_BYTES) == cudaSuccess
         && cudaFuncSetAttribute(k_render_mega<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MAX_STAGE_BYTES) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(k_wf_intersect<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MAX_STAGE_BYTES) == cudaSuccess
         && cudaFuncSetAttribute(k_intersect_batch<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MAX_STAGE_BYTES) == cudaSuccess;
    if (!ok)
    {
        fail(RTCU_ERR_CUDA, "context setup failed: %s", cudaGetErrorString(cudaGetLastError()));
        rtcu_destroy(ctx);
        return nullptr;
    }
    return ctx;
}

void rtcu_destroy(rtcu_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    ctx->scene_arena.release(); ctx->h_scene_arena.release();
    ctx->accum.release(); ctx->rgba8.release(); ctx->h_rgba8.release(); ctx->h_accum.release();
    for (int i = 0; i < 2; i++) { ctx->wf_o[i].release(); ctx->wf_d[i].release(); ctx->wf_thr[i].release(); }
    for (auto& l : ctx->wf_list) l.release();
    ctx->wf_rad.release(); ctx->wf_sum.release(); ctx->wf_hit.release(); ctx->wf_counts.release(); ctx->h_wf_counts.release();
    ctx->raster_prim.release(); ctx->raster_depth.release();
    ctx->h_raster_prim.release(); ctx->h_raster_depth.release();
    ctx->tile_cost.release(); ctx->tile_order.release(); ctx->beam_lists.release(); ctx->run_scratch.release();
    ctx->counters.release(); ctx->stragglers.release(); ctx->straggler_count.release(); ctx->h_counters.release(); ctx->scratch.release();
    for (auto& e : ctx->tile_ev)
        if (e) cudaEventDestroy(e);
    for (auto& b : ctx->ipc)
    {
        if (b.second) cudaFree(b.first);
        else cudaIpcCloseMemHandle(b.first);
    }
    for (auto& e : ctx->ev)
        if (e) cudaEventDestroy(e);
    if (ctx->ev_launch) cudaEventDestroy(ctx->ev_launch);
    if (ctx->out_reg.ptr && ctx->out_reg.ok && cudaHostUnregister(ctx->out_reg.ptr) != cudaSuccess) cudaGetLastError();
    for (auto& e : ctx->copy_events)
        if (e) cudaEventDestroy(e);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

namespace {
int upload_scene(rtcu_ctx* const* ctxs, uint32_t n_ctx, const rtcu_scene* s);
}

int rtcu_upload_scene(rtcu_ctx* ctx, const rtcu_scene* s)
{
    return guarded("rtcu_upload_scene", [&] { return upload_scene(&ctx, 1, s); });
}

int rtcu_upload_scene_multi(rtcu_ctx* const* ctxs, uint32_t n_ctx, const rtcu_scene* s)
{
    if (!ctxs || n_ctx == 0 || n_ctx > 8) return fail(RTCU_ERR_INVALID, "1..8 contexts");
    return guarded("rtcu_upload_scene_multi", [&] { return upload_scene(ctxs, n_ctx, s); });
}

namespace {
// validation, device columns and the BVH are made once; every context then receives the same arena from one pinned staging copy
int upload_scene(rtcu_ctx* const* ctxs, uint32_t n_ctx, const rtcu_scene* s)
{
    for (uint32_t g = 0; g < n_ctx; g++)
        if (!ctxs[g]) return fail(RTCU_ERR_INVALID, "null context");
    rtcu_ctx* ctx = ctxs[0];
    if (!ctx || !s) return fail(RTCU_ERR_INVALID, "null argument");
    if (s->n_materials == 0 || !s->materials) return fail(RTCU_ERR_INVALID, "scene has no materials (scene.cpp:565-566 always provides one)");
    if ((s->n_spheres && (!s->spheres || !s->sphere_material)) || (s->n_planes && (!s->planes || !s->plane_material)))
        return fail(RTCU_ERR_INVALID, "null primitive column");
    if (s->n_boxes && (!s->boxes || !s->box_material)) return fail(RTCU_ERR_INVALID, "null primitive column");
    if (s->n_spheres >= 0x40000000u || s->n_planes >= 0x40000000u || s->n_boxes >= 0x40000000u) return fail(RTCU_ERR_INVALID, "too many primitives");
    // scene.cpp:568-574 range-checks material indices at load; re-check at the boundary
    for (uint32_t i = 0; i < s->n_spheres; i++)
        if (s->sphere_material[i] >= s->n_materials) return fail(RTCU_ERR_INVALID, "sphere %u: material index %u out-of-range", i, s->sphere_material[i]);
    for (uint32_t i = 0; i < s->n_planes; i++)
        if (s->plane_material[i] >= s->n_materials) return fail(RTCU_ERR_INVALID, "plane %u: material index %u out-of-range", i, s->plane_material[i]);
    for (uint32_t i = 0; i < s->n_boxes; i++)
        if (s->box_material[i] >= s->n_materials) return fail(RTCU_ERR_INVALID, "box %u: material index %u out-of-range", i, s->box_material[i]);
    for (uint32_t i = 0; i < s->n_materials; i++)
        if (s->materials[i].type > RTCU_DIAMOND) return fail(RTCU_ERR_INVALID, "material %u: type %u is not a material_type", i, s->materials[i].type);

    CU(cudaSetDevice(ctx->device));
    std::vector<float4> sph(s->n_spheres);
    for (uint32_t i = 0; i < s->n_spheres; i++)
    {
        const float* p = s->spheres + 4 * (size_t)i;
        const float r = p[3];
        const volatile float r2 = r * r; // S4: r2 = r*r, one IEEE multiply
        sph[i] = make_float4(p[0], p[1], p[2], r2);
    }
    // packed-scan layout: pair j = spheres 2j, 2j+1 as {cx0,cx1,cy0,cy1},{cz0,cz1,r2_0,r2_1}; everything past the last
    // sphere (odd tail, padding to an even pair count, two prefetch pairs) is a never-hit entry: centre 0, r2 = -inf
    const float4 never = make_float4(0.0f, 0.0f, 0.0f, -__builtin_inff());
    std::vector<float4> pairs(pair_float4_count(s->n_spheres));
    for (size_t j = 0; 2 * j < pairs.size(); j++)
    {
        const float4 s0 = (2 * j < s->n_spheres) ? sph[2 * j] : never;
        const float4 s1 = (2 * j + 1 < s->n_spheres) ? sph[2 * j + 1] : never;
        pairs[2 * j] = make_float4(s0.x, s1.x, s0.y, s1.y);
        pairs[2 * j + 1] = make_float4(s0.z, s1.z, s0.w, s1.w);
    }
    std::vector<MatRec> mats(s->n_materials);
    for (uint32_t i = 0; i < s->n_materials; i++)
    {
        const rtcu_material& m = s->materials[i];
        MatRec r;
        const volatile float ar = m.albedo[0] * m.reflectivity, ag = m.albedo[1] * m.reflectivity, ab = m.albedo[2] * m.reflectivity;
        r.att_r = ar; r.att_g = ag; r.att_b = ab;
        r.roughness = m.roughness;
        r.ior = m.reflectivity;
        r.type = m.type;
        {
            // the per-material constants of the dielectric branch, each a single IEEE operation as in the source
            // (sm_ray_tracer.cpp:176-177, :205): 1/ior and r0 = ((1 - ior) / (1 + ior))^2
            const volatile float inv = 1.0f / m.reflectivity;
            const volatile float num = 1.0f - m.reflectivity, den = 1.0f + m.reflectivity;
            const volatile float q = num / den;
            const volatile float r0 = q * q;
            r.inv_ior = inv;
            r.r0 = r0;
        }
        mats[i] = r;
    }
    // rasterizer columns (rasterizer.cpp reads boxes.value() and materials.albedo(), which the path tracers never touch)
    std::vector<float4> boxes(2 * (size_t)s->n_boxes), albedo(s->n_materials);
    for (uint32_t i = 0; i < s->n_boxes; i++)
    {
        const float* b = s->boxes + 6 * (size_t)i;
        const volatile float lx = b[0] - b[3], ly = b[1] - b[4], lz = b[2] - b[5]; // S13: lo = c - e, hi = c + e
        const volatile float hx = b[0] + b[3], hy = b[1] + b[4], hz = b[2] + b[5];
        boxes[2 * i] = make_float4(lx, ly, lz, 0.0f);
        boxes[2 * i + 1] = make_float4(hx, hy, hz, 0.0f);
    }
    for (uint32_t i = 0; i < s->n_materials; i++)
        albedo[i] = make_float4(s->materials[i].albedo[0], s->materials[i].albedo[1], s->materials[i].albedo[2], s->materials[i].albedo[3]);
    // BVH: built whenever there are spheres (cheap for small scenes; lets ACCEL_BVH be requested explicitly for parity
    // tests); RTCU_ACCEL_AUTO uses it from rtcu_bvh_threshold() spheres up
    bool have_bvh = false;
    uint32_t bvh_depth = 0;
    float ms_bvh_build = 0.0f;
    std::vector<float4> nodes_dev, leaf_blk;
    // The traversal returns the scan's result only while S4 cannot overflow: with every centre coordinate, radius and ray origin
    // within 2^62, |e|^2, a^2 and r^2 stay finite, so no inf - inf = NaN can reach a comparison (the scan accepts a NaN distance where
    // the leaves' (t, index) rule rejects it).  A scene beyond that range -- the loader only rejects NaN / Inf, scene.cpp:95-99 --
    // keeps the reference's O(N) scan; rays whose origin leaves the range fall back to it one by one (trav_init).
    bool bvh_safe = true;
    for (uint32_t i = 0; i < s->n_spheres && bvh_safe; i++)
        for (int k = 0; k < 4; k++)
            bvh_safe = bvh_safe && std::fabs(s->spheres[4 * (size_t)i + k]) <= 0x1p62f; // false for NaN too
    if (s->n_spheres && bvh_safe)
    {
        const auto t0 = std::chrono::steady_clock::now();
        rtcu_bvh::Pool pool(s->n_spheres >= 8192 ? rtcu_bvh::thread_count() - 1 : 0); // one set of workers for build and pack
        const rtcu_bvh::Result bvh = rtcu_bvh::build(s->spheres, s->n_spheres, pool);
        ms_bvh_build = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
        bvh_depth = bvh.max_depth;
        uint32_t depth4 = 0;
        if (s->n_spheres < (1u << 29))
            pack_bvh4(bvh, sph, nodes_dev, leaf_blk, depth4, pool);
        // a visit pushes at most three children: the traversal stack needs 3 entries per level
        if (!nodes_dev.empty() && 3 * depth4 + 2 <= (uint32_t)BVH_STACK)
        {
            bvh_depth = depth4;
            have_bvh = true;
        }
        else
        {
            nodes_dev.clear();
            leaf_blk.clear();
        }
    }
    // One arena, one copy: every column is laid out at a 256-byte aligned offset of a pinned staging buffer and goes to the
    // device in a single transfer (a dozen small copies from pageable memory cost ~5 us each -- most of the upload for the
    // small scenes an interactive session re-sends).
    struct Seg { size_t off, bytes; const void* src; };
    std::vector<Seg> segs;
    size_t total = 0;
    auto add = [&](const void* src, size_t bytes) {
        const size_t off = (total + 255) & ~(size_t)255;
        total = off + (bytes ? bytes : 16); // an empty column still gets a valid address
        segs.push_back(Seg{ off, bytes, src });
        return off;
    };
    const size_t o_pairs = add(pairs.data(), pairs.size() * sizeof(float4));
    const size_t o_sph = add(sph.data(), sph.size() * sizeof(float4));
    const size_t o_sph_mat = add(s->sphere_material, (size_t)s->n_spheres * sizeof(uint32_t));
    const size_t o_planes = add(s->planes, (size_t)s->n_planes * sizeof(float4));
    const size_t o_plane_mat = add(s->plane_material, (size_t)s->n_planes * sizeof(uint32_t));
    const size_t o_mats = add(mats.data(), mats.size() * sizeof(MatRec));
    const size_t o_boxes = add(boxes.data(), boxes.size() * sizeof(float4));
    const size_t o_box_mat = add(s->box_material, (size_t)s->n_boxes * sizeof(uint32_t));
    const size_t o_albedo = add(albedo.data(), albedo.size() * sizeof(float4));
    const size_t o_nodes = add(nodes_dev.data(), nodes_dev.size() * sizeof(float4));
    const size_t o_leaves = add(leaf_blk.data(), leaf_blk.size() * sizeof(float4));
    // make sure no kernel of a previous frame still reads the old scene, nor a previous upload the staging buffer
    for (uint32_t g = 0; g < n_ctx; g++)
    {
        rtcu_ctx* c = ctxs[g];
        CU(cudaSetDevice(c->device));
        CU(cudaStreamSynchronize(c->stream));
        if (c->last_stream && c->last_stream != c->stream && cudaStreamSynchronize(c->last_stream) != cudaSuccess)
            cudaGetLastError(); // the caller's stream may be gone by now; nothing of ours can be running on it then
        CU(c->scene_arena.reserve(total));
    }
    CU(cudaSetDevice(ctx->device));
    CU(ctx->h_scene_arena.reserve(total)); // (pinned memory is portable: every device copies from the first context's staging)
    for (const Seg& g : segs)
        if (g.bytes) memcpy(ctx->h_scene_arena.p + g.off, g.src, g.bytes);
    for (uint32_t g = 0; g < n_ctx; g++)
    {
        CU(cudaSetDevice(ctxs[g]->device));
        CU(cudaMemcpyAsync(ctxs[g]->scene_arena.p, ctx->h_scene_arena.p, total, cudaMemcpyHostToDevice, ctxs[g]->stream));
    }
    for (uint32_t g = 0; g < n_ctx; g++)
    {
        rtcu_ctx* c = ctxs[g];
        CU(cudaSetDevice(c->device));
        CU(cudaStreamSynchronize(c->stream)); // renders on other streams (rtcu_render_device) must see the new scene
        c->tile_hist_valid = false; // per-tile costs belong to the scene they were measured on
        c->have_bvh = have_bvh;
        c->bvh_depth = bvh_depth;
        c->ms_bvh_build = ms_bvh_build;
        unsigned char* base = c->scene_arena.p;
        c->scene.bvh_nodes = c->have_bvh ? reinterpret_cast<const float4*>(base + o_nodes) : nullptr;
        c->scene.leaf_blk = c->have_bvh ? reinterpret_cast<const float4*>(base + o_leaves) : nullptr;
        c->scene.n_bvh_nodes = (uint32_t)(nodes_dev.size() / 8);
        c->scene.spheres = reinterpret_cast<const float4*>(base + o_sph);
        c->scene.pairs = reinterpret_cast<const float4*>(base + o_pairs);
        c->scene.sphere_material = reinterpret_cast<const uint32_t*>(base + o_sph_mat);
        c->scene.n_spheres = s->n_spheres;
        c->scene.planes = reinterpret_cast<const float4*>(base + o_planes);
        c->scene.plane_material = reinterpret_cast<const uint32_t*>(base + o_plane_mat);
        c->scene.n_planes = s->n_planes;
        c->scene.materials = reinterpret_cast<const MatRec*>(base + o_mats);
        c->scene.n_materials = s->n_materials;
        c->raster.spheres = c->scene.spheres;
        c->raster.pairs = c->scene.pairs;
        c->raster.sphere_material = c->scene.sphere_material;
        c->raster.n_spheres = s->n_spheres;
        c->raster.planes = c->scene.planes;
        c->raster.plane_material = c->scene.plane_material;
        c->raster.n_planes = s->n_planes;
        c->raster.boxes = reinterpret_cast<const float4*>(base + o_boxes);
        c->raster.box_material = reinterpret_cast<const uint32_t*>(base + o_box_mat);
        c->raster.n_boxes = s->n_boxes;
        c->raster.albedo = reinterpret_cast<const float4*>(base + o_albedo);
        c->have_scene = true;
    }
    return RTCU_OK;
}
} // namespace

int rtcu_rasterize_device(rtcu_ctx* ctx, const rtcu_view* view, uint32_t* d_rgba8, void* stream)
{
    if (!ctx || !view || !d_rgba8) return fail(RTCU_ERR_INVALID, "null argument");
    CU(cudaSetDevice(ctx->device));
    return launch_rasterize(ctx, view, d_rgba8, nullptr, nullptr, (cudaStream_t)stream);
}

int rtcu_rasterize(rtcu_ctx* ctx, const rtcu_view* view, uint32_t* rgba8_out, uint32_t* prim_out, float* depth_out)
{
    if (!ctx || !view || !rgba8_out) return fail(RTCU_ERR_INVALID, "null argument");
    CU(cudaSetDevice(ctx->device));
    const size_t npix = (size_t)view->width * view->height;
    if (npix == 0) return fail(RTCU_ERR_INVALID, "empty image");
    CU(ctx->rgba8.reserve(npix));
    if (prim_out) CU(ctx->raster_prim.reserve(npix));
    if (depth_out) CU(ctx->raster_depth.reserve(npix));
    const bool full = view->tile_x0 == 0 && view->tile_y0 == 0 && view->tile_x1 == view->width && view->tile_y1 == view->height;
    // zero-copy into a pinned / registered image, as rtcu_render does
    uint32_t* d_out = ctx->rgba8.p;
    bool zero_copy = false;
    bool verify = false;
    if (full && ctx->knobs.zero_copy)
    {
        if (uint32_t* alias = output_alias(ctx, rgba8_out, npix, &verify))
        {
            d_out = alias;
            zero_copy = true;
            if (verify) arm_sentinels(rgba8_out, npix);
        }
    }
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    int rc = launch_rasterize(ctx, view, d_out, prim_out ? ctx->raster_prim.p : nullptr, depth_out ? ctx->raster_depth.p : nullptr, ctx->stream);
    if (rc) return rc;
    CU(cudaEventRecord(ctx->ev[1], ctx->stream));
    if (zero_copy)
    {
        CU(cudaStreamSynchronize(ctx->stream));
        if (verify && !sentinels_overwritten(rgba8_out, npix))
        {
            // stale registration (see output_alias): drop it, draw the frame again into device memory and stage it out
            drop_registration(ctx);
            if ((rc = launch_rasterize(ctx, view, ctx->rgba8.p, nullptr, nullptr, ctx->stream))) return rc;
            if ((rc = copy_tile_out(ctx, view, rgba8_out, ctx->rgba8.p, ctx->h_rgba8))) return rc;
        }
    }
    else if ((rc = copy_tile_out(ctx, view, rgba8_out, ctx->rgba8.p, ctx->h_rgba8)))
        return rc;
    if (prim_out && (rc = copy_tile_out(ctx, view, prim_out, ctx->raster_prim.p, ctx->h_raster_prim))) return rc;
    if (depth_out && (rc = copy_tile_out(ctx, view, depth_out, ctx->raster_depth.p, ctx->h_raster_depth))) return rc;
    CU(cudaEventRecord(ctx->ev[2], ctx->stream));
    CU(cudaEventSynchronize(ctx->ev[2]));
    CU(cudaEventElapsedTime(&ctx->stats.ms_render, ctx->ev[0], ctx->ev[1]));
    CU(cudaEventElapsedTime(&ctx->stats.ms_d2h, ctx->ev[1], ctx->ev[2]));
    ctx->stats.ms_resolve = 0.0f;
    ctx->stats.ms_h2d = 0.0f;
    return RTCU_OK;
}

int rtcu_render_device(rtcu_ctx* ctx, const rtcu_view* view, float* d_accum, int accumulate, void* stream)
{
    if (!ctx || !view || !d_accum) return fail(RTCU_ERR_INVALID, "null argument");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream; // passed through: NULL is CUDA's default stream, as torch uses it
    ctx->last_stream = st;
    ctx->counters_pending = true;
    return launch_render(ctx, view, reinterpret_cast<float4*>(d_accum), nullptr, accumulate, st);
}

int rtcu_resolve_device(rtcu_ctx* ctx, const float* d_accum, uint32_t width, uint32_t height, uint32_t spp, uint32_t* d_rgba8, void* stream)
{
    if (!ctx || !d_accum || !d_rgba8 || !width || !height || !spp) return fail(RTCU_ERR_INVALID, "bad argument");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t n = width * height;
    k_resolve<<<(n + 255) / 256, 256, 0, st>>>(reinterpret_cast<const float4*>(d_accum), n, (float)spp, d_rgba8);
    CU(cudaGetLastError());
    return RTCU_OK;
}

int rtcu_sync(rtcu_ctx* ctx)
{
    if (!ctx) return fail(RTCU_ERR_INVALID, "null context");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    return RTCU_OK;
}

int rtcu_render(rtcu_ctx* ctx, const rtcu_view* view, uint32_t* rgba8_out, float* accum_out)
{
    if (!ctx || !view) return fail(RTCU_ERR_INVALID, "null argument");
    CU(cudaSetDevice(ctx->device));
    const size_t npix = (size_t)view->width * view->height;
    if (npix == 0) return fail(RTCU_ERR_INVALID, "empty image");
    // progressive refinement: this call's samples are added to the sums of the previous calls, which stay on the device
    const int accumulate = (view->flags & RTCU_FLAG_ACCUMULATE) ? 1 : 0;
    if (accumulate && (ctx->accum_w != view->width || ctx->accum_h != view->height || !ctx->accum.p))
        return fail(RTCU_ERR_STATE, "RTCU_FLAG_ACCUMULATE: the context holds no sums of a %ux%u image", view->width, view->height);
    CU(ctx->accum.reserve(npix));
    CU(ctx->rgba8.reserve(npix));
    ctx->accum_w = view->width;
    ctx->accum_h = view->height;
    const bool full = view->tile_x0 == 0 && view->tile_y0 == 0 && view->tile_x1 == view->width && view->tile_y1 == view->height;

    // Zero-copy output: when the caller's image is pinned / registered host memory and the whole frame is rendered, the
    // kernels store the packed pixels straight into it over PCIe as each pixel finishes, so the read-back overlaps the
    // frame instead of following it (RTCU_ZERO_COPY=0 disables; pageable buffers -- the reference's image.cpp -- are staged).
    // A pageable image (the reference's) is page-locked once per (pointer, size) and then treated the same way, with every frame
    // verified (see output_alias).
    uint32_t* d_out = rgba8_out ? ctx->rgba8.p : nullptr;
    bool zero_copy = false, verify = false;
    if (rgba8_out && full && ctx->knobs.zero_copy && (ctx->knobs.zero_copy_direct || !uses_direct_mode(ctx, view)))
    {
        if (uint32_t* alias = output_alias(ctx, rgba8_out, npix, &verify))
        {
            d_out = alias;
            zero_copy = true;
            if (verify) arm_sentinels(rgba8_out, npix);
        }
    }

    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    int rc = launch_render(ctx, view, ctx->accum.p, d_out, accumulate, ctx->stream);
    if (rc) return rc;
    CU(cudaEventRecord(ctx->ev[1], ctx->stream));

    // copy the tile rows back (whole image when the tile is the image)
    const size_t row0 = full ? 0 : view->tile_y0, rows = full ? view->height : view->tile_y1 - view->tile_y0;
    const size_t off = row0 * view->width, cnt = rows * view->width;
    if (rgba8_out && zero_copy)
    {
        // the kernels store the pixels straight into the caller's buffer: the one synchronisation in fetch_counters below
        // completes them together with the counters
    }
    else if (rgba8_out)
    {
        if (full)
            rc = deliver_image(ctx, rgba8_out, ctx->rgba8.p, npix);
        else
        {
            CU(ctx->h_rgba8.reserve(npix));
            CU(cudaMemcpyAsync(ctx->h_rgba8.p + off, ctx->rgba8.p + off, cnt * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
            for (uint32_t y = view->tile_y0; y < view->tile_y1; y++)
                memcpy(rgba8_out + (size_t)y * view->width + view->tile_x0, ctx->h_rgba8.p + (size_t)y * view->width + view->tile_x0,
                       (size_t)(view->tile_x1 - view->tile_x0) * sizeof(uint32_t));
        }
        if (rc) return rc;
    }
    if (accum_out)
    {
        if (full)
            rc = copy_out(ctx, accum_out, reinterpret_cast<const float*>(ctx->accum.p), npix * 4, ctx->h_accum);
        else
        {
            CU(ctx->h_accum.reserve(npix * 4));
            CU(cudaMemcpyAsync(ctx->h_accum.p + off * 4, reinterpret_cast<const float*>(ctx->accum.p) + off * 4, cnt * 4 * sizeof(float),
                               cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
            for (uint32_t y = view->tile_y0; y < view->tile_y1; y++)
                memcpy(accum_out + ((size_t)y * view->width + view->tile_x0) * 4, ctx->h_accum.p + ((size_t)y * view->width + view->tile_x0) * 4,
                       (size_t)(view->tile_x1 - view->tile_x0) * 4 * sizeof(float));
        }
        if (rc) return rc;
    }
    CU(cudaEventRecord(ctx->ev[2], ctx->stream));
    rc = fetch_counters(ctx, ctx->stream);
    if (rc) return rc;
    CU(cudaEventElapsedTime(&ctx->stats.ms_render, ctx->ev[0], ctx->ev[1]));
    CU(cudaEventElapsedTime(&ctx->stats.ms_d2h, ctx->ev[1], ctx->ev[2]));
    ctx->stats.ms_resolve = 0.0f;
    ctx->stats.ms_h2d = 0.0f;
    if (zero_copy && verify && !sentinels_overwritten(rgba8_out, npix))
    {
        // the registration was stale (the application re-allocated its image at the same address): the kernels wrote into pages
        // the application no longer sees.  Drop it and deliver this frame from the fp32 sums through the staged copy.
        drop_registration(ctx);
        k_resolve<<<(unsigned)((npix + 255) / 256), 256, 0, ctx->stream>>>(ctx->accum.p, (uint32_t)npix, (float)view->samples_per_pixel, ctx->rgba8.p);
        CU(cudaGetLastError());
        rc = copy_out(ctx, rgba8_out, ctx->rgba8.p, npix, ctx->h_rgba8);
        if (rc) return rc;
    }
    return RTCU_OK;
}

int rtcu_accum_download(rtcu_ctx* ctx, uint32_t width, uint32_t height, float* accum_out)
{
    if (!ctx || !accum_out) return fail(RTCU_ERR_INVALID, "null argument");
    if (!ctx->accum.p || ctx->accum_w != width || ctx->accum_h != height || !width || !height)
        return fail(RTCU_ERR_STATE, "the context holds no sums of a %ux%u image", width, height);
    CU(cudaSetDevice(ctx->device));
    return copy_out(ctx, accum_out, reinterpret_cast<const float*>(ctx->accum.p), (size_t)width * height * 4, ctx->h_accum);
}

int rtcu_accum_upload(rtcu_ctx* ctx, uint32_t width, uint32_t height, const float* accum_in)
{
    if (!ctx || !accum_in || !width || !height) return fail(RTCU_ERR_INVALID, "bad argument");
    if ((uint64_t)width * height > 0xFFFFFFFFull) return fail(RTCU_ERR_INVALID, "image too large");
    CU(cudaSetDevice(ctx->device));
    const size_t npix = (size_t)width * height;
    CU(cudaStreamSynchronize(ctx->stream));
    CU(ctx->accum.reserve(npix));
    CU(cudaMemcpyAsync(ctx->accum.p, accum_in, npix * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->accum_w = width;
    ctx->accum_h = height;
    return RTCU_OK;
}

int rtcu_render_multi(rtcu_ctx* const* ctxs, uint32_t n_ctx, const rtcu_view* view, uint32_t* rgba8_out, float* accum_out)
{
    if (!ctxs || n_ctx == 0 || n_ctx > 8 || !view) return fail(RTCU_ERR_INVALID, "bad argument");
    if (n_ctx == 1) return rtcu_render(ctxs[0], view, rgba8_out, accum_out);
    const size_t npix = (size_t)view->width * view->height;
    const uint32_t s0 = view->sample_begin, total = view->sample_end - view->sample_begin;
    // sample-range split: device g renders [s0 + g*total/G, s0 + (g+1)*total/G)
    for (uint32_t g = 0; g < n_ctx; g++)
    {
        rtcu_ctx* c = ctxs[g];
        if (!c) return fail(RTCU_ERR_INVALID, "null context %u", g);
        CU(cudaSetDevice(c->device));
        CU(c->accum.reserve(npix));
        c->accum_w = view->width;
        c->accum_h = view->height;
        CU(cudaMemsetAsync(c->accum.p, 0, npix * sizeof(float4), c->stream)); // pixels outside the tile must add 0
        rtcu_view v = *view;
        v.sample_begin = s0 + (uint32_t)((uint64_t)total * g / n_ctx);
        v.sample_end = s0 + (uint32_t)((uint64_t)total * (g + 1) / n_ctx);
        if (g == 0) CU(cudaEventRecord(c->ev[0], c->stream));
        c->stats.segments = c->stats.node_visits = c->stats.sphere_tests = 0; // a context with an empty share reports nothing
        c->stats.kernel_launches = 0;
        if (v.sample_end > v.sample_begin)
        {
            const int rc = launch_render(c, &v, c->accum.p, nullptr, 0, c->stream);
            if (rc) return rc;
        }
        else
            CU(cudaMemsetAsync(c->counters.p, 0, 8 * sizeof(unsigned long long), c->stream)); // not the counters of an earlier frame
        CU(cudaEventRecord(c->ev[3], c->stream));
    }
    // root waits for every peer, then sums their buffers through peer loads inside the resolve kernel
    rtcu_ctx* root = ctxs[0];
    CU(cudaSetDevice(root->device));
    PeerList peers;
    peers.n = 0;
    for (uint32_t g = 1; g < n_ctx; g++)
    {
        int can = 0;
        CU(cudaDeviceCanAccessPeer(&can, root->device, ctxs[g]->device));
        if (!can) return fail(RTCU_ERR_CUDA, "device %d cannot access peer %d", root->device, ctxs[g]->device);
        const cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[g]->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(RTCU_ERR_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
        cudaGetLastError();
        CU(cudaStreamWaitEvent(root->stream, ctxs[g]->ev[3], 0));
        peers.ptr[peers.n++] = ctxs[g]->accum.p;
    }
    CU(root->rgba8.reserve(npix));
    CU(cudaEventRecord(root->ev[1], root->stream));
    k_reduce_resolve<<<(unsigned)((npix + 255) / 256), 256, 0, root->stream>>>(root->accum.p, peers, (uint32_t)npix, (float)view->samples_per_pixel,
                                                                             rgba8_out ? root->rgba8.p : nullptr);
    CU(cudaGetLastError());
    CU(cudaEventRecord(root->ev[2], root->stream));
    int rc = RTCU_OK;
    if (rgba8_out) rc = deliver_image(root, rgba8_out, root->rgba8.p, npix);
    if (!rc && accum_out) rc = copy_out(root, accum_out, reinterpret_cast<const float*>(root->accum.p), npix * 4, root->h_accum);
    if (rc) return rc;
    uint64_t segs = 0, nodes = 0, tests = 0;
    uint32_t launches = 1; // the fused reduce + resolve
    const uint32_t accel = root->stats.accel;
    for (uint32_t g = 0; g < n_ctx; g++)
    {
        CU(cudaSetDevice(ctxs[g]->device));
        ctxs[g]->stats.accel = accel; // (an empty share launched nothing: count its zeroed counters the same way)
        rc = fetch_counters(ctxs[g], ctxs[g]->stream);
        if (rc) return rc;
        segs += ctxs[g]->stats.segments;
        nodes += ctxs[g]->stats.node_visits;
        tests += ctxs[g]->stats.sphere_tests;
        launches += ctxs[g]->stats.kernel_launches;
    }
    CU(cudaSetDevice(root->device));
    CU(cudaStreamSynchronize(root->stream));
    CU(cudaEventElapsedTime(&root->stats.ms_render, root->ev[0], root->ev[1]));
    CU(cudaEventElapsedTime(&root->stats.ms_resolve, root->ev[1], root->ev[2]));
    root->stats.segments = segs;
    root->stats.node_visits = nodes;   // counted per device (BVH); segments x spheres for the scan
    root->stats.sphere_tests = tests;
    root->stats.samples = (uint64_t)(view->tile_x1 - view->tile_x0) * (view->tile_y1 - view->tile_y0) * total;
    root->stats.kernel_launches = launches;
    return RTCU_OK;
}

// ---- one process per GPU: buffers shared through CUDA IPC, exchange fused into the resolve kernel -------------------------
static_assert(sizeof(cudaIpcMemHandle_t) == 64, "rtcu_ipc_* pass the handle as 64 opaque bytes");

int rtcu_ipc_alloc(rtcu_ctx* ctx, uint64_t bytes, void** d_ptr, unsigned char handle[64])
{
    if (!ctx || !bytes || !d_ptr || !handle) return fail(RTCU_ERR_INVALID, "null argument");
    CU(cudaSetDevice(ctx->device));
    void* p = nullptr;
    CU(cudaMalloc(&p, bytes));
    cudaIpcMemHandle_t h;
    const cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess)
    {
        cudaFree(p);
        return fail(RTCU_ERR_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
    }
    CU(cudaMemsetAsync(p, 0, bytes, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    memcpy(handle, &h, 64);
    ctx->ipc.emplace_back(p, true);
    *d_ptr = p;
    return RTCU_OK;
}

int rtcu_ipc_open(rtcu_ctx* ctx, const unsigned char handle[64], void** d_ptr)
{
    if (!ctx || !handle || !d_ptr) return fail(RTCU_ERR_INVALID, "null argument");
    CU(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    void* p = nullptr;
    CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    ctx->ipc.emplace_back(p, false);
    *d_ptr = p;
    return RTCU_OK;
}

int rtcu_ipc_release(rtcu_ctx* ctx, void* d_ptr)
{
    if (!ctx || !d_ptr) return fail(RTCU_ERR_INVALID, "null argument");
    CU(cudaSetDevice(ctx->device));
    for (size_t i = 0; i < ctx->ipc.size(); i++)
        if (ctx->ipc[i].first == d_ptr)
        {
            const bool owned = ctx->ipc[i].second;
            ctx->ipc.erase(ctx->ipc.begin() + (long)i);
            CU(cudaDeviceSynchronize());
            if (owned) CU(cudaFree(d_ptr));
            else CU(cudaIpcCloseMemHandle(d_ptr));
            return RTCU_OK;
        }
    return fail(RTCU_ERR_INVALID, "not a buffer of rtcu_ipc_alloc / rtcu_ipc_open");
}

int rtcu_reduce_resolve_rows(rtcu_ctx* ctx, const float* const* d_accums, uint32_t n_bufs, uint32_t width, uint32_t row0, uint32_t rows, uint32_t spp,
                             uint32_t* d_rgba8, void* stream)
{
    if (!ctx || !d_accums || !d_rgba8 || !width || !spp) return fail(RTCU_ERR_INVALID, "bad argument");
    if (n_bufs == 0 || n_bufs > 8) return fail(RTCU_ERR_INVALID, "1..8 buffers");
    if ((uint64_t)width * ((uint64_t)row0 + rows) > 0xFFFFFFFFull) return fail(RTCU_ERR_INVALID, "image too large");
    PeerList bufs;
    bufs.n = (int)n_bufs;
    for (uint32_t g = 0; g < n_bufs; g++)
    {
        if (!d_accums[g]) return fail(RTCU_ERR_INVALID, "buffer %u is null", g);
        bufs.ptr[g] = reinterpret_cast<const float4*>(d_accums[g]);
    }
    if (rows == 0) return RTCU_OK;
    CU(cudaSetDevice(ctx->device));
    const uint32_t count = width * rows;
    k_reduce_resolve_rows<<<(count + 255) / 256, 256, 0, (cudaStream_t)stream>>>(bufs, width * row0, count, (float)spp, d_rgba8);
    CU(cudaGetLastError());
    return RTCU_OK;
}

int rtcu_exchange_reduce_resolve(rtcu_ctx* ctx, const float* const* d_accums, void* const* d_flags, uint32_t n_ranks, uint32_t rank, uint32_t dst_rank,
                                 uint32_t epoch, uint32_t width, uint32_t row0, uint32_t rows, uint32_t spp, uint32_t* d_rgba8, void* stream)
{
    if (!ctx || !d_accums || !d_flags || !d_rgba8 || !width || !spp) return fail(RTCU_ERR_INVALID, "bad argument");
    if (n_ranks == 0 || n_ranks > 8 || rank >= n_ranks || dst_rank >= n_ranks) return fail(RTCU_ERR_INVALID, "1..8 ranks, rank and destination among them");
    if (epoch == 0) return fail(RTCU_ERR_INVALID, "frames are numbered from 1 (0 is the state of a fresh flag block)");
    if ((uint64_t)width * ((uint64_t)row0 + rows) > 0xFFFFFFFFull) return fail(RTCU_ERR_INVALID, "image too large");
    PeerList bufs;
    FlagList flags;
    bufs.n = (int)n_ranks;
    for (uint32_t g = 0; g < n_ranks; g++)
    {
        if (!d_accums[g] || !d_flags[g]) return fail(RTCU_ERR_INVALID, "buffer %u is null", g);
        bufs.ptr[g] = reinterpret_cast<const float4*>(d_accums[g]);
        flags.ptr[g] = reinterpret_cast<ExchangeFlags*>(d_flags[g]);
    }
    CU(cudaSetDevice(ctx->device));
    const uint32_t count = width * rows;
    // a rank without rows still takes part in the handshake; otherwise at most eight CTAs per SM, striding over the band
    const unsigned blocks = count ? std::min<unsigned>((count + 255) / 256, (unsigned)ctx->sm_count * 8u) : 1u;
    const long long timeout = 20ll * 2000000000ll;            // ~20 s of SM clock cycles
    k_exchange_reduce_resolve<<<blocks, 256, 0, (cudaStream_t)stream>>>(bufs, flags, (int)rank, (int)dst_rank, epoch, width * row0, count, (float)spp, d_rgba8, timeout);
    CU(cudaGetLastError());
    return RTCU_OK;
}

int rtcu_exchange_check(rtcu_ctx* ctx, const void* d_flags, void* stream)
{
    if (!ctx || !d_flags) return fail(RTCU_ERR_INVALID, "null argument");
    CU(cudaSetDevice(ctx->device));
    ExchangeFlags h;
    CU(cudaMemcpyAsync(&h, d_flags, sizeof h, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CU(cudaStreamSynchronize((cudaStream_t)stream));
    if (h.error) return fail(RTCU_ERR_STATE, "exchange: a rank did not arrive within the time limit (ready %u %u %u %u %u %u %u %u)", h.ready[0], h.ready[1],
                             h.ready[2], h.ready[3], h.ready[4], h.ready[5], h.ready[6], h.ready[7]);
    return RTCU_OK;
}

int rtcu_intersect_batch(rtcu_ctx* ctx, const float* o, const float* d, uint32_t n, uint8_t* hit, uint32_t* prim, float* t, float* normal,
                         uint32_t accel)
{
    if (!ctx || !o || !d || !hit || !prim || !t) return fail(RTCU_ERR_INVALID, "null argument");
    if (!ctx->have_scene) return fail(RTCU_ERR_STATE, "rtcu_upload_scene has not been called");
    if (accel > RTCU_ACCEL_BVH) return fail(RTCU_ERR_INVALID, "bad accel selector %u", accel);
    if (accel == RTCU_ACCEL_BVH && !ctx->have_bvh) return fail(RTCU_ERR_STATE, "no BVH for this scene");
    const bool use_bvh = accel == RTCU_ACCEL_BVH || (accel == RTCU_ACCEL_AUTO && ctx->have_bvh && ctx->scene.n_spheres >= ctx->knobs.bvh_threshold);
    if (n == 0) return RTCU_OK;
    CU(cudaSetDevice(ctx->device));
    const size_t need = 2 * padded((size_t)n * 12) + padded(n) + 2 * padded((size_t)n * 4) + padded((size_t)n * 12);
    CU(ctx->scratch.reserve(need));
    Carver c{ ctx->scratch.p };
    float* d_o = c.take<float>((size_t)n * 3);
    float* d_d = c.take<float>((size_t)n * 3);
    uint8_t* d_hit = c.take<uint8_t>(n);
    uint32_t* d_prim = c.take<uint32_t>(n);
    float* d_t = c.take<float>(n);
    float* d_n = c.take<float>((size_t)n * 3);
    cudaStream_t st = ctx->stream;
    if (ctx->launch_pending && ctx->launch_stream != st) CU(cudaStreamWaitEvent(st, ctx->ev_launch, 0)); // shares the counters with the render kernels
    CU(cudaMemcpyAsync(d_o, o, (size_t)n * 12, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_d, d, (size_t)n * 12, cudaMemcpyHostToDevice, st));
    const unsigned blocks = (unsigned)((n + 255) / 256 < (uint32_t)ctx->sm_count * 8 ? (n + 255) / 256 : ctx->sm_count * 8);
    const size_t sb = stage_bytes(ctx);
    CU(cudaMemsetAsync(ctx->counters.p, 0, 8 * sizeof(unsigned long long), st));
    CU(cudaEventRecord(ctx->ev[0], st));
    if (use_bvh && ctx->knobs.bvh_trav != 1)
        k_intersect_batch<false, true, 0><<<blocks, 256, 0, st>>>(ctx->scene, d_o, d_d, n, d_hit, d_prim, d_t, normal ? d_n : nullptr, ctx->counters.p);
    else if (use_bvh)
        k_intersect_batch<false, true, 1><<<blocks, 256, 0, st>>>(ctx->scene, d_o, d_d, n, d_hit, d_prim, d_t, normal ? d_n : nullptr, ctx->counters.p);
    else if (sb <= MAX_STAGE_BYTES)
        k_intersect_batch<true, false><<<blocks, 256, sb, st>>>(ctx->scene, d_o, d_d, n, d_hit, d_prim, d_t, normal ? d_n : nullptr, ctx->counters.p);
    else
        k_intersect_batch<false, false><<<blocks, 256, 0, st>>>(ctx->scene, d_o, d_d, n, d_hit, d_prim, d_t, normal ? d_n : nullptr, ctx->counters.p);
    CU(cudaGetLastError());
    CU(cudaEventRecord(ctx->ev[1], st));
    CU(cudaMemcpyAsync(hit, d_hit, n, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(prim, d_prim, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(t, d_t, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    if (normal) CU(cudaMemcpyAsync(normal, d_n, (size_t)n * 12, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaEventElapsedTime(&ctx->stats.ms_render, ctx->ev[0], ctx->ev[1]));
    ctx->stats.kernel_launches = 1;
    ctx->stats.accel = use_bvh ? RTCU_ACCEL_BVH : RTCU_ACCEL_LINEAR;
    const int rc = fetch_counters(ctx, st);
    if (rc) return rc;
    ctx->stats.segments = n;
    if (!use_bvh) ctx->stats.sphere_tests = (uint64_t)n * ctx->scene.n_spheres;
    return RTCU_OK;
}

int rtcu_primary_rays(rtcu_ctx* ctx, const rtcu_view* view, const uint32_t* px, const uint32_t* py, const uint32_t* sample, uint32_t n, float* o,
                      float* d)
{
    if (!ctx || !view || !px || !py || !sample || !o || !d) return fail(RTCU_ERR_INVALID, "null argument");
    if (n == 0) return RTCU_OK;
    CU(cudaSetDevice(ctx->device));
    RenderParams p;
    rtcu_view v = *view;
    if (v.max_bounces == 0) v.max_bounces = 1;
    const int rc = make_params(ctx, &v, p);
    if (rc) return rc;
    CU(ctx->scratch.reserve(3 * padded((size_t)n * 4) + 2 * padded((size_t)n * 12)));
    Carver c{ ctx->scratch.p };
    uint32_t* d_px = c.take<uint32_t>(n);
    uint32_t* d_py = c.take<uint32_t>(n);
    uint32_t* d_s = c.take<uint32_t>(n);
    float* d_o = c.take<float>((size_t)n * 3);
    float* d_d = c.take<float>((size_t)n * 3);
    cudaStream_t st = ctx->stream;
    CU(cudaMemcpyAsync(d_px, px, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_py, py, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_s, sample, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    k_primary_rays<<<(n + 255) / 256, 256, 0, st>>>(p.cam, p.width, p.rk, d_px, d_py, d_s, n, d_o, d_d);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(o, d_o, (size_t)n * 12, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(d, d_d, (size_t)n * 12, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return RTCU_OK;
}

int rtcu_scatter_batch(rtcu_ctx* ctx, uint32_t material_mode, uint64_t seed, uint32_t n, const uint32_t* material, const float* o, const float* d,
                       const float* t, const float* normal, const uint32_t* pixel, const uint32_t* sample, const uint32_t* block,
                       uint8_t* scattered, float* att, float* o_out, float* d_out)
{
    if (!ctx || !material || !o || !d || !t || !normal || !pixel || !sample || !block || !scattered || !att || !o_out || !d_out)
        return fail(RTCU_ERR_INVALID, "null argument");
    if (!ctx->have_scene) return fail(RTCU_ERR_STATE, "rtcu_upload_scene has not been called");
    if (material_mode > RTCU_MODE_SM) return fail(RTCU_ERR_INVALID, "bad material_mode");
    for (uint32_t i = 0; i < n; i++)
        if (material[i] >= ctx->scene.n_materials) return fail(RTCU_ERR_INVALID, "item %u: material index out-of-range", i);
    if (n == 0) return RTCU_OK;
    CU(cudaSetDevice(ctx->device));
    CU(ctx->scratch.reserve(5 * padded((size_t)n * 4) + 6 * padded((size_t)n * 12) + padded(n)));
    Carver c{ ctx->scratch.p };
    uint32_t* d_mat = c.take<uint32_t>(n);
    uint32_t* d_pix = c.take<uint32_t>(n);
    uint32_t* d_smp = c.take<uint32_t>(n);
    uint32_t* d_blk = c.take<uint32_t>(n);
    float* d_t = c.take<float>(n);
    float* d_o = c.take<float>((size_t)n * 3);
    float* d_d = c.take<float>((size_t)n * 3);
    float* d_n = c.take<float>((size_t)n * 3);
    float* d_att = c.take<float>((size_t)n * 3);
    float* d_oo = c.take<float>((size_t)n * 3);
    float* d_do = c.take<float>((size_t)n * 3);
    uint8_t* d_sc = c.take<uint8_t>(n);
    cudaStream_t st = ctx->stream;
    CU(cudaMemcpyAsync(d_mat, material, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_pix, pixel, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_smp, sample, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_blk, block, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_t, t, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_o, o, (size_t)n * 12, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_d, d, (size_t)n * 12, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_n, normal, (size_t)n * 12, cudaMemcpyHostToDevice, st));
    k_scatter_batch<<<(n + 255) / 256, 256, 0, st>>>(ctx->scene, material_mode, philox_keys(make_uint2((uint32_t)seed, (uint32_t)(seed >> 32))), n, d_mat, d_o, d_d,
                                                    d_t, d_n, d_pix, d_smp, d_blk, d_sc, d_att, d_oo, d_do);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(scattered, d_sc, n, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(att, d_att, (size_t)n * 12, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(o_out, d_oo, (size_t)n * 12, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(d_out, d_do, (size_t)n * 12, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return RTCU_OK;
}

int rtcu_philox_batch(rtcu_ctx* ctx, const uint32_t* ctr, uint32_t n, uint64_t key, uint32_t* out)
{
    if (!ctx || !ctr || !out) return fail(RTCU_ERR_INVALID, "null argument");
    if (n == 0) return RTCU_OK;
    CU(cudaSetDevice(ctx->device));
    CU(ctx->scratch.reserve(2 * padded((size_t)n * 16)));
    Carver c{ ctx->scratch.p };
    uint4* d_c = c.take<uint4>(n);
    uint4* d_o = c.take<uint4>(n);
    cudaStream_t st = ctx->stream;
    CU(cudaMemcpyAsync(d_c, ctr, (size_t)n * 16, cudaMemcpyHostToDevice, st));
    k_philox_batch<<<(n + 255) / 256, 256, 0, st>>>(d_c, n, philox_keys(make_uint2((uint32_t)key, (uint32_t)(key >> 32))), d_o);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, d_o, (size_t)n * 16, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return RTCU_OK;
}

int rtcu_measure_fp32_peak(rtcu_ctx* ctx, float* tflops_ffma, float* tflops_ffma2)
{
    if (!ctx || !tflops_ffma || !tflops_ffma2) return fail(RTCU_ERR_INVALID, "null argument");
    CU(cudaSetDevice(ctx->device));
    CU(ctx->scratch.reserve(256));
    const int iters = 4096, blocks = ctx->sm_count * 8;
    const double flop = (double)blocks * 256 * iters * 4 * 8 * 2 /*lanes of float2*/ * 2 /*fma*/;
    float* results[2] = { tflops_ffma, tflops_ffma2 };
    for (int variant = 0; variant < 2; variant++)
    {
        float best = 0.0f;
        for (int rep = 0; rep < 4; rep++) // rep 0 warms up
        {
            CU(cudaEventRecord(ctx->ev[4], ctx->stream));
            if (variant == 0)
                k_fp32_peak<false><<<blocks, 256, 0, ctx->stream>>>(reinterpret_cast<float*>(ctx->scratch.p), iters, 0.999f, 0.001f);
            else
                k_fp32_peak<true><<<blocks, 256, 0, ctx->stream>>>(reinterpret_cast<float*>(ctx->scratch.p), iters, 0.999f, 0.001f);
            CU(cudaGetLastError());
            CU(cudaEventRecord(ctx->ev[5], ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
            float ms = 0.0f;
            CU(cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]));
            const float tf = (float)(flop / (ms * 1e-3) / 1e12);
            if (rep > 0 && tf > best) best = tf;
        }
        *results[variant] = best;
    }
    return RTCU_OK;
}

int rtcu_selftest_math(rtcu_ctx* ctx, const float* divisors, uint32_t n_divisors, uint64_t counts[4])
{
    if (!ctx || !counts || (n_divisors && !divisors)) return fail(RTCU_ERR_INVALID, "null argument");
    if (n_divisors > 256) return fail(RTCU_ERR_INVALID, "at most 256 divisors");
    for (uint32_t i = 0; i < n_divisors; i++)
        if (!(divisors[i] >= 1.0f && divisors[i] <= 16777216.0f)) return fail(RTCU_ERR_INVALID, "divisor %u outside [1, 2^24]", i);
    CU(cudaSetDevice(ctx->device));
    CU(ctx->scratch.reserve(4 * sizeof(unsigned long long) + 256 * sizeof(float)));
    unsigned long long* d_counts = reinterpret_cast<unsigned long long*>(ctx->scratch.p);
    float* d_div = reinterpret_cast<float*>(ctx->scratch.p + 4 * sizeof(unsigned long long));
    CU(cudaMemsetAsync(d_counts, 0, 4 * sizeof(unsigned long long), ctx->stream));
    if (n_divisors) CU(cudaMemcpyAsync(d_div, divisors, n_divisors * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    k_selftest_math<<<ctx->sm_count * 16, 256, 0, ctx->stream>>>(d_div, n_divisors, d_counts);
    CU(cudaGetLastError());
    unsigned long long h[4];
    CU(cudaMemcpyAsync(h, d_counts, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < 4; i++) counts[i] = h[i];
    return RTCU_OK;
}

int rtcu_bvh_build_host(const float* spheres, uint32_t n, float* nodes_out, uint32_t* order_out, uint32_t max_nodes, uint32_t* n_nodes,
                        uint32_t* depth)
{
    if ((n && !spheres) || !n_nodes || !depth) return fail(RTCU_ERR_INVALID, "null argument");
    return guarded("rtcu_bvh_build_host", [&]() -> int {
        const rtcu_bvh::Result bvh = rtcu_bvh::build(spheres, n);
        *n_nodes = (uint32_t)bvh.nodes.size();
        *depth = bvh.max_depth;
        if (nodes_out)
        {
            if (bvh.nodes.size() > max_nodes) return fail(RTCU_ERR_INVALID, "nodes_out too small: need %zu", bvh.nodes.size());
            memcpy(nodes_out, bvh.nodes.data(), bvh.nodes.size() * sizeof(rtcu_bvh::Node));
        }
        if (order_out && n) memcpy(order_out, bvh.order.data(), n * sizeof(uint32_t));
        return RTCU_OK;
    });
}

int rtcu_bvh4_build_host(const float* spheres, uint32_t n, float* nodes_out, uint32_t max_nodes, float* leaves_out, uint32_t max_leaves,
                         uint32_t* n_nodes, uint32_t* n_leaves, uint32_t* depth)
{
    if ((n && !spheres) || !n_nodes || !n_leaves || !depth) return fail(RTCU_ERR_INVALID, "null argument");
    if (n == 0 || n >= (1u << 29)) return fail(RTCU_ERR_INVALID, "1 .. 2^29 - 1 spheres");
    return guarded("rtcu_bvh4_build_host", [&]() -> int {
        std::vector<float4> sph(n);
        for (uint32_t i = 0; i < n; i++)
        {
            const float* p = spheres + 4 * (size_t)i;
            const volatile float r2 = p[3] * p[3];
            sph[i] = make_float4(p[0], p[1], p[2], r2);
        }
        rtcu_bvh::Pool pool(n >= 8192 ? rtcu_bvh::thread_count() - 1 : 0);
        const rtcu_bvh::Result bvh = rtcu_bvh::build(spheres, n, pool);
        std::vector<float4> nodes, leaves;
        uint32_t d4 = 0;
        pack_bvh4(bvh, sph, nodes, leaves, d4, pool);
        *n_nodes = (uint32_t)(nodes.size() / 8);
        *n_leaves = (uint32_t)(leaves.size() / 5);
        *depth = d4;
        if (nodes_out)
        {
            if (*n_nodes > max_nodes) return fail(RTCU_ERR_INVALID, "nodes_out too small: need %u", *n_nodes);
            memcpy(nodes_out, nodes.data(), nodes.size() * sizeof(float4));
        }
        if (leaves_out)
        {
            if (*n_leaves > max_leaves) return fail(RTCU_ERR_INVALID, "leaves_out too small: need %u", *n_leaves);
            memcpy(leaves_out, leaves.data(), leaves.size() * sizeof(float4));
        }
        return RTCU_OK;
    });
}

int rtcu_get_stats(rtcu_ctx* ctx, rtcu_stats* out)
{
    if (!ctx || !out) return fail(RTCU_ERR_INVALID, "null argument");
    if (ctx->counters_pending)
    {
        CU(cudaSetDevice(ctx->device));
        const int rc = fetch_counters(ctx, ctx->last_stream);
        if (rc) return rc;
        ctx->counters_pending = false;
    }
    *out = ctx->stats;
    return RTCU_OK;
}

} // extern "C"
