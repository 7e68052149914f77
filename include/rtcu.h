/*
 * rtcu.h -- C ABI of the B200 (sm_100a) path-tracing library behind marzer/rt's renderer plugin.
 *
 * This is the drop-in boundary for the reference's hot path: everything
 * `rt::renderer_interface::render(const scene&, image_view&, muu::thread_pool&)`
 * (reference src/renderer.hpp:9-14) does for `mg_ray_tracer` (src/renderers/mg_ray_tracer.cpp:178-205)
 * and `sm_ray_tracer` (src/renderers/sm_ray_tracer.cpp:263-289) is reachable through these entry
 * points with plain pointers and sizes.  The reference-side binding is plugin/cuda_path_tracer.cpp
 * (see INTEGRATION.md); tests and bench.py bind the same symbols through ctypes.
 *
 * There is NO CPU fallback: every compute entry point returns RTCU_ERR_CUDA when no sm_100-class
 * device is usable.
 *
 * Conventions: every function returning int returns RTCU_OK (0) or a negative RTCU_ERR_*; the
 * message is available from rtcu_last_error() (thread-local).  Host pointers unless named d_*.
 */
#ifndef RTCU_H
#define RTCU_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTCU_ABI_VERSION 3

enum {
    RTCU_OK = 0,
    RTCU_ERR_INVALID = -1, /* bad argument (null pointer, empty tile, material index out of range) */
    RTCU_ERR_CUDA = -2,    /* CUDA runtime error, or no usable device                              */
    RTCU_ERR_STATE = -3,   /* call order (render before upload_scene)                              */
    RTCU_ERR_NOMEM = -4    /* a host allocation failed (C++ exceptions never cross this ABI)       */
};

/* rt::material_type values, reference src/common.hpp:105-115 (ABI of the materials table) */
enum {
    RTCU_LAMBERT = 0, RTCU_METAL = 1, RTCU_DIELECTRIC = 2, RTCU_AIR = 3,
    RTCU_VACUUM = 4, RTCU_WATER = 5, RTCU_ICE = 6, RTCU_DIAMOND = 7
};

/* which scatter_funcs table to apply:
 *   RTCU_MODE_MG  mg_ray_tracer.cpp:142-152  (metal -> metal, everything else lambert)
 *   RTCU_MODE_SM  sm_ray_tracer.cpp:221-236  (+ dielectric/air/vacuum/water/ice -> dielectric_scatter) */
enum { RTCU_MODE_MG = 0, RTCU_MODE_SM = 1 };

/* traversal selector for rtcu_view.flags / rtcu_intersect_batch */
enum {
    RTCU_ACCEL_AUTO = 0,   /* linear scan for small scenes, BVH above rtcu_bvh_threshold() spheres */
    RTCU_ACCEL_LINEAR = 1, /* the reference's O(N) scan (mg_ray_tracer.cpp:62-87)                  */
    RTCU_ACCEL_BVH = 2
};
/* pipeline selector, bits 4..7 of rtcu_view.flags */
enum {
    RTCU_PIPE_AUTO = 0 << 4,
    RTCU_PIPE_MEGAKERNEL = 1 << 4, /* register-resident paths, per-pixel sample regeneration */
    RTCU_PIPE_WAVEFRONT = 2 << 4   /* generate / intersect / shade / compact over HBM queues  */
};

/* bit 8 of rtcu_view.flags, rtcu_render only: progressive refinement.  The call's samples [sample_begin, sample_end) are ADDED to
 * the fp32 sums the context keeps on the device from its previous rtcu_render calls of the same image size, and rgba8_out is
 * resolved over samples_per_pixel (pass the total so far).  Because sample indices are global under the counter-based RNG, k
 * calls of n samples equal one call of k n samples up to fp32 summation order.  The sums never leave the device unless
 * accum_out / rtcu_accum_download asks for them. */
#define RTCU_FLAG_ACCUMULATE 0x100u

/* primitive ids reported by rtcu_intersect_batch / rtcu_rasterize: a sphere index, or one of these */
#define RTCU_PRIM_MISS  0xFFFFFFFFu
#define RTCU_PRIM_PLANE 0x80000000u /* | plane index */
#define RTCU_PRIM_BOX   0x40000000u /* | box index (rasterizer only) */

/* one row of rt::materials (reference src/soa.hpp:157-170) minus the `name` column */
typedef struct rtcu_material {
    uint32_t type;         /* materials.type()                                             */
    float    albedo[4];    /* materials.albedo(): rt::colour rgba                          */
    float    roughness;    /* materials.roughness()                                        */
    float    reflectivity; /* materials.reflectivity(); the IOR for dielectric-class types */
} rtcu_material;

/* a flattened rt::scene (reference src/scene.hpp:8-25).  The pointers are the reference's own
 * soagen columns: spheres.value() is muu::bounding_sphere<float>[] = {cx,cy,cz,radius} (soa.hpp:194),
 * planes.value() is muu::plane<float>[] = {nx,ny,nz,d}; boxes.value() is muu::bounding_box<float>[] =
 * {cx,cy,cz,ex,ey,ez} (soa.hpp:138-150).  Boxes are never hit by either ray tracer
 * (mg_ray_tracer.cpp:89-93): only rtcu_rasterize reads them, and they may be NULL / 0. */
typedef struct rtcu_scene {
    const float*         spheres;
    const uint32_t*      sphere_material; /* spheres.material() */
    uint32_t             n_spheres;
    const float*         planes;
    const uint32_t*      plane_material;
    uint32_t             n_planes;
    const rtcu_material* materials;
    uint32_t             n_materials;
    const float*         boxes;           /* ABI 2 */
    const uint32_t*      box_material;    /* boxes.material() */
    uint32_t             n_boxes;
} rtcu_scene;

/* one render call.  inv_view_proj is viewport::inverse_view_projection (reference src/camera.hpp:17,
 * :122-137), column-major (element (r,c) at [c*4+r]).  The call renders global sample indices
 * [sample_begin, sample_end) of the pixel rectangle [tile_x0,tile_x1) x [tile_y0,tile_y1); the
 * reference's single-call behaviour is sample range [0, samples_per_pixel) over the full image. */
typedef struct rtcu_view {
    float    inv_view_proj[16];
    uint32_t width, height;
    uint32_t samples_per_pixel; /* resolve divisor: scene::samples_per_pixel (scene.hpp:10) */
    uint32_t max_bounces;       /* scene::max_bounces (scene.hpp:11)                        */
    uint32_t sample_begin, sample_end;
    uint32_t tile_x0, tile_y0, tile_x1, tile_y1;
    uint64_t seed;
    uint32_t material_mode; /* RTCU_MODE_* */
    uint32_t flags;         /* RTCU_ACCEL_* | RTCU_PIPE_* */
} rtcu_view;

typedef struct rtcu_stats {
    uint64_t segments;       /* path segments traced by the last render call (exact)         */
    uint64_t samples;        /* pixel samples of the last render call                        */
    uint64_t sphere_tests;   /* ray-sphere tests (linear: segments*n_spheres; BVH: counted)  */
    uint64_t node_visits;    /* BVH nodes visited (0 for linear)                             */
    float    ms_render;      /* device time of the trace kernels, CUDA events                */
    float    ms_resolve;     /* device time of resolve/pack                                  */
    float    ms_h2d, ms_d2h; /* copies of the host entry points                              */
    uint32_t kernel_launches; /* kernels launched by the last call                           */
    uint32_t pipeline;        /* RTCU_PIPE_* actually used                                   */
    uint32_t accel;           /* RTCU_ACCEL_* actually used                                  */
    uint32_t reserved;
} rtcu_stats;

typedef struct rtcu_ctx rtcu_ctx;

/* ---- lifetime (one context per device; replaces `new T` in REGISTER_RENDERER, renderer.hpp:34-41) */
int         rtcu_abi_version(void);
int         rtcu_device_count(void);
rtcu_ctx*   rtcu_create(int device);   /* NULL on failure, see rtcu_last_error() */
void        rtcu_destroy(rtcu_ctx* ctx);
const char* rtcu_last_error(void);
uint32_t    rtcu_bvh_threshold(void);
/* The RTCU_* environment variables (experiment knobs) are read once in rtcu_create, never on the launch path;
 * this re-reads them for a context that is already alive (A/B tests). */
int         rtcu_reload_env(rtcu_ctx* ctx);

/* Threading and ordering.  A context serves ONE frame at a time: the render kernels of a context share its counters, its
 * straggler queue and its tile-cost map.  Calls on one context must come from one thread at a time (render() is called from the
 * UI thread only, window.cpp:213-217).  Launches may go to different streams (rtcu_render_device): the library orders each
 * launch after the previous one of the same context with an event, so two frames of one context never overlap on the device;
 * use one context per concurrently rendered frame. */

/* ---- scene upload: replaces the implicit `const scene&` argument of render (renderer.hpp:11).
 * Copies the columns to the device (and builds the BVH when the sphere count calls for it).  The
 * caller may free its buffers afterwards.  The build runs on worker threads that live for this call (the cores the process
 * may use, at most 16; RTCU_BVH_THREADS=n overrides); the tree does not depend on their number.  RTCU_ERR_NOMEM when a host
 * allocation fails. */
int rtcu_upload_scene(rtcu_ctx* ctx, const rtcu_scene* scene);
/* the same for the contexts of a single-process multi-GPU renderer (rtcu_render_multi): validated and built once, then copied to
 * every context's device */
int rtcu_upload_scene_multi(rtcu_ctx* const* ctxs, uint32_t n_ctx, const rtcu_scene* scene);

/* ---- render: replaces mg_ray_tracer::render / sm_ray_tracer::render.
 * rgba8_out (nullable): width*height uint32 in image_view layout (row 0 = top, src/image.hpp:150-159),
 *   packed as colour::operator uint32_t (colour.hpp:100-106); only the tile is written.
 * accum_out (nullable): width*height*4 floats {sum_r,sum_g,sum_b,n_samples}; only the tile is written.  The paths are the
 *   reference's (same segments per sample for the same seed); the per-pixel sum runs in ascending sample order as in
 *   mg_ray_tracer.cpp:187-194, except where several lanes share a pixel (from 4 samples per call: BVH scenes, and scenes
 *   without a BVH below 3840x2160; pixels handed to the second pass; multi-GPU sample splits): there it is a fixed tree of
 *   partial sums -- deterministic, equal to the sequential sum up to fp32 rounding, and independent of the tile only while the
 *   tile takes the same lane count as the frame (RTCU_SCAN_DIRECT=0 RTCU_BVH_DIRECT=0: always the sequential sum).
 * rgba8_out may be pinned or pageable.  A pinned image is written by the kernels themselves (zero-copy) or by one DMA; a pageable
 *   one (the reference's, image.cpp:9-13) is staged through a bounce buffer -- unless the caller has opted in with
 *   rtcu_set_output_pinning, see there. */
int rtcu_render(rtcu_ctx* ctx, const rtcu_view* view, uint32_t* rgba8_out, float* accum_out);
/* Opt-in for callers whose image outlives the frame, like the reference's back buffer (back_buffer.hpp: one rt::image per
 * window size, handed to render() frame after frame): with enable != 0 a pageable full-frame rgba8_out of rtcu_render /
 * rtcu_rasterize / rtcu_render_multi is page-locked once per (pointer, size) -- cudaHostRegister, undone when another image
 * arrives, on enable = 0 and in rtcu_destroy -- and from then on treated like pinned memory (no staging copy: C2's frame is
 * complete 0.9 ms earlier).  Every frame is verified to have landed in the caller's pages, so an image that was freed and
 * re-allocated at the same address costs one staged frame, never a wrong picture.  Off by default: a page-locked range that
 * the application has freed stays registered until the next frame notices, and other CUDA calls of the process that touch
 * re-used parts of it would fail -- acceptable for a renderer plugin that owns the process's CUDA use, not for a library
 * default. */
int rtcu_set_output_pinning(rtcu_ctx* ctx, int enable);

/* the fp32 sums {sum_r, sum_g, sum_b, n} the context holds from its last rtcu_render (width*height*4 floats): read them back
 * (checkpoint of a progressive render) or replace them (resume; the next RTCU_FLAG_ACCUMULATE call adds onto them) */
int rtcu_accum_download(rtcu_ctx* ctx, uint32_t width, uint32_t height, float* accum_out);
int rtcu_accum_upload(rtcu_ctx* ctx, uint32_t width, uint32_t height, const float* accum_in);

/* ---- preview: replaces rasterizer::render (reference src/renderers/rasterizer.cpp:22-88): one ray per pixel through the
 * pixel centre, nearest of planes, boxes, spheres (strict '<' in that order, no minimum distance), N.L shading against
 * the eye, no gamma.  Uses inv_view_proj, width, height, the tile and the RTCU_ACCEL_* bits of flags; every other field is
 * ignored.
 * rgba8_out as in rtcu_render.  prim_out / depth_out (nullable, width*height each): per pixel the sphere index,
 * RTCU_PRIM_PLANE|index, RTCU_PRIM_BOX|index or RTCU_PRIM_MISS, and the accepted distance -- the parity hooks. */
int rtcu_rasterize(rtcu_ctx* ctx, const rtcu_view* view, uint32_t* rgba8_out, uint32_t* prim_out, float* depth_out);
/* the same into DEVICE memory, asynchronously on `stream` (cudaStream_t as void*; NULL = legacy default stream) */
int rtcu_rasterize_device(rtcu_ctx* ctx, const rtcu_view* view, uint32_t* d_rgba8, void* stream);

/* ---- device-resident variants (no host copies) for multi-GPU composition and benchmarking.
 * d_accum: width*height float4 on ctx's device.  accumulate != 0 adds onto the existing contents.
 * stream: a cudaStream_t, passed through (NULL = CUDA's default stream, which is what torch's current
 * stream usually is); the call is asynchronous, rtcu_get_stats() synchronises that stream to read counters. */
int rtcu_render_device(rtcu_ctx* ctx, const rtcu_view* view, float* d_accum, int accumulate, void* stream);
/* resolve: divide by samples_per_pixel, sqrt, clamp, pack (mg_ray_tracer.cpp:195-200) */
int rtcu_resolve_device(rtcu_ctx* ctx, const float* d_accum, uint32_t width, uint32_t height,
                        uint32_t samples_per_pixel, uint32_t* d_rgba8, void* stream);
int rtcu_sync(rtcu_ctx* ctx);

/* ---- single-process multi-GPU render: sample ranges are split evenly over the contexts (one per
 * device), each device accumulates its range, ctxs[0] sums the peers' fp32 buffers through NVLink peer
 * loads inside its resolve kernel, and the host image is copied from ctxs[0]. */
int rtcu_render_multi(rtcu_ctx* const* ctxs, uint32_t n_ctx, const rtcu_view* view,
                      uint32_t* rgba8_out, float* accum_out);

/* ---- level-1 parity entry point: closest hit of n rays (o/d: n x {x,y,z}) against the uploaded
 * scene; mirrors test_planes/test_spheres/select (mg_ray_tracer.cpp:35-102).
 * hit[i] in {0,1}; prim[i] = sphere index | 0x80000000+plane index | 0xFFFFFFFF on miss;
 * t[i] = hit distance or -1; normal (nullable) n x {x,y,z}.  accel: RTCU_ACCEL_*. */
int rtcu_intersect_batch(rtcu_ctx* ctx, const float* o, const float* d, uint32_t n,
                         uint8_t* hit, uint32_t* prim, float* t, float* normal, uint32_t accel);

/* ---- step-wise parity entry points (device kernels run on n items, host buffers) */
/* primary rays for n (pixel x, pixel y, sample) triples: mg_ray_tracer.cpp:189-193 */
int rtcu_primary_rays(rtcu_ctx* ctx, const rtcu_view* view, const uint32_t* px, const uint32_t* py,
                      const uint32_t* sample, uint32_t n, float* o, float* d);
/* one scatter per item: mg_ray_tracer.cpp:109-140, sm_ray_tracer.cpp:181-219.
 * scattered[i] = 1/0; att/o_out/d_out n x 3. */
int rtcu_scatter_batch(rtcu_ctx* ctx, uint32_t material_mode, uint64_t seed, uint32_t n,
                       const uint32_t* material, const float* o, const float* d, const float* t,
                       const float* normal, const uint32_t* pixel, const uint32_t* sample,
                       const uint32_t* block, uint8_t* scattered, float* att, float* o_out, float* d_out);
/* blocks of the SPEC's random stream (Philox4x32, 7 rounds) for n counters (ctr n x 4, out n x 4) under one key */
int rtcu_philox_batch(rtcu_ctx* ctx, const uint32_t* ctr, uint32_t n, uint64_t key, uint32_t* out);

int rtcu_get_stats(rtcu_ctx* ctx, rtcu_stats* out);

/* ---- one process per GPU (torch.distributed / MPI style launches): the exchange step without a collective library.
 * rtcu_ipc_alloc allocates a zeroed device buffer and returns its 64-byte CUDA IPC handle; the other ranks of the node map
 * it with rtcu_ipc_open (NVLink peer access).  rtcu_reduce_resolve_rows then sums the n_bufs accumulation buffers in the
 * order given (use rank order for a deterministic result) over image rows [row0, row0 + rows) and writes the resolved
 * RGBA8 pixels (mg_ray_tracer.cpp:195-200) into d_rgba8 -- the full-image buffer of the destination rank, own or mapped --
 * in one kernel on `stream`.  The caller orders it after every rank's render (any barrier on the stream) and keeps the
 * buffers unchanged until every rank's reduce has finished.  rtcu_ipc_release frees / unmaps; rtcu_destroy does so too. */
int rtcu_ipc_alloc(rtcu_ctx* ctx, uint64_t bytes, void** d_ptr, unsigned char handle[64]);
int rtcu_ipc_open(rtcu_ctx* ctx, const unsigned char handle[64], void** d_ptr);
int rtcu_ipc_release(rtcu_ctx* ctx, void* d_ptr);
int rtcu_reduce_resolve_rows(rtcu_ctx* ctx, const float* const* d_accums, uint32_t n_bufs, uint32_t width, uint32_t row0, uint32_t rows,
                             uint32_t spp, uint32_t* d_rgba8, void* stream);
/* The whole exchange step as ONE launch per rank, with no barrier from a collective library around it.  d_flags[g] is rank g's
 * 128-byte flag block (rtcu_ipc_alloc(128), mapped by the peers like the buffers; zero = fresh).  Frame `epoch` (1, 2, 3, ... the
 * same sequence on every rank): the kernel publishes "rank's buffer complete" to every rank (st.release.sys over NVLink), waits
 * until all ranks have published (ld.acquire.sys on its own block), sums the n_ranks buffers over rows [row0, row0 + rows) in rank
 * order, resolves and stores into d_rgba8 (the destination rank's image, own or mapped), then publishes "band stored" to the
 * destination; the destination's kernel returns only when every band is stored.  The call must follow the rank's render on
 * `stream`.  Alternate between TWO accumulation buffers per rank from frame to frame: a buffer is then never overwritten while a
 * peer still reads it (kernels.cuh, k_exchange_reduce_resolve).  A rank that waits longer than ~20 s gives up and raises the
 * block's error word: rtcu_exchange_check synchronises `stream` and returns RTCU_ERR_STATE if that happened. */
int rtcu_exchange_reduce_resolve(rtcu_ctx* ctx, const float* const* d_accums, void* const* d_flags, uint32_t n_ranks, uint32_t rank,
                                 uint32_t dst_rank, uint32_t epoch, uint32_t width, uint32_t row0, uint32_t rows, uint32_t spp,
                                 uint32_t* d_rgba8, void* stream);
int rtcu_exchange_check(rtcu_ctx* ctx, const void* d_flags, void* stream);

/* exhaustive device-side check of the kernels' cheaper-but-exact square root / reciprocal / constant-divisor division
 * (rt_b200/csrc/spec.cuh) against the IEEE intrinsics: counts[0..1] = float patterns (of all 2^32) where sqrt / 1/sqrt
 * differ, counts[2] = differing quotients over every float a in {0} u [2^-24, 2^24] and every b in divisors (each in
 * [1, 2^24], at most 256), counts[3] = quotients tested.  All three must be 0. */
int rtcu_selftest_math(rtcu_ctx* ctx, const float* divisors, uint32_t n_divisors, uint64_t counts[4]);

/* ---- host-only: the BVH builder used by rtcu_upload_scene (binned SAH, two children per node), exposed so its
 * invariants can be checked without a device.  The reference has no acceleration structure (mg_ray_tracer.cpp:62-87
 * scans all spheres); traversal returns the scan's exact result.
 * nodes_out (nullable): max_nodes x 16 words {x[4], y[4], z[4] = l.lo,l.hi,r.lo,r.hi per axis; int32 child[2] (>= 0 inner
 * node, < 0 leaf starting at ~child in order_out); uint32 count[2]}.  order_out (nullable): n original sphere indices. */
int rtcu_bvh_build_host(const float* spheres, uint32_t n, float* nodes_out, uint32_t* order_out, uint32_t max_nodes,
                        uint32_t* n_nodes, uint32_t* depth);
/* the same tree as the device sees it: collapsed to 4-wide nodes (8 x float4 each: children (0,1) as {c0,c1,h0,h1} for x, y,
 * z -- centre and outward-rounded half-extent, -inf for an empty slot --, the same for children (2,3), the four child
 * references as bit patterns (>= 0 inner node, 0x80000000 | leaf), the four H = h.x+h.y+h.z) and 80-byte leaf blocks
 * (5 x float4: two packed sphere pairs {cx0,cx1,cy0,cy1},{cz0,cz1,r2_0,r2_1}, then four original indices as bit patterns,
 * 0x7fffffff = padding).  depth = levels of 4-wide nodes.  Pass NULL buffers first to learn the sizes. */
int rtcu_bvh4_build_host(const float* spheres, uint32_t n, float* nodes_out, uint32_t max_nodes, float* leaves_out, uint32_t max_leaves,
                         uint32_t* n_nodes, uint32_t* n_leaves, uint32_t* depth);

/* ---- roofline calibration: achieved non-tensor FP32 TFLOP/s of an FFMA stream and of a packed FFMA2
 * (fma.rn.f32x2) stream on ctx's device at the clocks it currently runs (no memory traffic). */
int rtcu_measure_fp32_peak(rtcu_ctx* ctx, float* tflops_ffma, float* tflops_ffma2);

#ifdef __cplusplus
}
#endif
#endif
