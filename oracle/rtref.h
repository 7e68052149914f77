/*
 * rtref.h -- CPU ORACLE for the marzer/rt path-tracing hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (rt_b200/, include/rtcu.h) never links, imports or calls it.
 *
 * PARITY STATUS: pinned against the reference's OWN renderer sources, unpinned against muu.  marzer/rt as a
 * whole cannot be built in this environment (its math library muu @06dbcecb, toml++, SDL2, imgui, argparse,
 * magic_enum and meson are absent) and it ships no tests, golden vectors or fixtures.
 * But oracle/Makefile `ref` compiles mg_ray_tracer.cpp, sm_ray_tracer.cpp, rasterizer.cpp and renderer.cpp
 * where they lie against oracle/ref_shim (a stand-in for the muu headers); this file reproduces the packed
 * images of that build bit for bit (tests/test_reference_build.py, tests/test_raster_oracle.py; fixtures
 * tests/golden/refbuild_*.npz, raster_*.npz).  What stays unpinned is muu's own arithmetic, restated from
 * its published algorithms as the numbered SPEC in rtref.c.  Further pins: hand-derived known-answer
 * vectors and the published Philox4x32 vectors (7 rounds = the SPEC's stream, and 10).
 *
 * The data structures deliberately mirror include/rtcu.h field by field so one harness can feed
 * both sides, but the two headers are independent files.
 */
#ifndef RTREF_H
#define RTREF_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* material_type enum values, reference src/common.hpp:105-115 */
enum {
    RTREF_LAMBERT = 0, RTREF_METAL = 1, RTREF_DIELECTRIC = 2, RTREF_AIR = 3,
    RTREF_VACUUM = 4, RTREF_WATER = 5, RTREF_ICE = 6, RTREF_DIAMOND = 7
};

/* scatter table selector: 0 = mg_ray_tracer.cpp:142-152, 1 = sm_ray_tracer.cpp:221-236 */
enum { RTREF_MODE_MG = 0, RTREF_MODE_SM = 1 };

typedef struct rtref_material {
    uint32_t type;        /* material_type, soa.hpp materials.type      */
    float    albedo[4];   /* rt::colour rgba, soa.hpp materials.albedo  */
    float    roughness;
    float    reflectivity; /* doubles as the IOR for dielectrics, scene.cpp:549-556 */
} rtref_material;

typedef struct rtref_scene {
    const float*          spheres;          /* n_spheres x {cx,cy,cz,radius} == spheres.value() */
    const uint32_t*       sphere_material;  /* spheres.material()                                */
    uint32_t              n_spheres;
    const float*          planes;           /* n_planes x {nx,ny,nz,d} == planes.value()         */
    const uint32_t*       plane_material;
    uint32_t              n_planes;
    const rtref_material* materials;
    uint32_t              n_materials;
    /* read by rtref_rasterize only; the path tracers never hit a box (mg_ray_tracer.cpp:89-93) */
    const float*          boxes;            /* n_boxes x {cx,cy,cz,ex,ey,ez} == boxes.value()    */
    const uint32_t*       box_material;
    uint32_t              n_boxes;
} rtref_scene;

typedef struct rtref_view {
    float    inv_view_proj[16]; /* viewport::inverse_view_projection, column-major m[c*4+r] */
    uint32_t width, height;     /* full image size                                           */
    uint32_t samples_per_pixel; /* divisor of the resolve step (scene::samples_per_pixel)    */
    uint32_t max_bounces;
    uint32_t sample_begin, sample_end; /* global sample indices rendered by this call       */
    uint32_t tile_x0, tile_y0, tile_x1, tile_y1; /* pixel rectangle [x0,x1) x [y0,y1)        */
    uint64_t seed;
    uint32_t material_mode;     /* RTREF_MODE_MG / RTREF_MODE_SM */
    uint32_t flags;             /* reserved, 0 */
} rtref_view;

/* --- RNG (replaces src/random.cpp on both sides) --------------------------------------- */
void  rtref_philox4x32_r(const uint32_t ctr[4], const uint32_t key[2], int rounds, uint32_t out[4]);
void  rtref_philox_stream(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]); /* the SPEC's stream: 7 rounds */
float rtref_u01(uint32_t x);                          /* (x >> 8) * 2^-24, in [0,1) */

/* --- level-1 parity entry point: closest hit over planes + spheres ---------------------- */
/* o/d: n x {x,y,z}.  hit[i] in {0,1}; prim[i] = sphere index, or 0x80000000|plane index,
 * 0xFFFFFFFF on miss; t[i] = distance (-1 on miss); nrm (nullable) n x {x,y,z}.          */
int rtref_intersect_batch(const rtref_scene* s, const float* o, const float* d, uint32_t n,
                          uint8_t* hit, uint32_t* prim, float* t, float* nrm);

/* --- unit entry points for step-wise parity --------------------------------------------- */
void rtref_primary_ray(const rtref_view* v, uint32_t px, uint32_t py, uint32_t sample,
                       float o[3], float d[3]);
/* the same for an explicit screen position (sx, sy) instead of pixel + sample jitter */
void rtref_screen_ray(const rtref_view* v, float sx, float sy, float o[3], float d[3]);
/* returns 1 if scattered, 0 if absorbed; block = philox block index (segment+1) */
int  rtref_scatter(const rtref_scene* s, uint32_t mode, uint32_t material,
                   const float o[3], const float d[3], float t, const float n[3],
                   uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t block,
                   float att[3], float o_out[3], float d_out[3]);
uint32_t rtref_pack_pixel(float sum_r, float sum_g, float sum_b, uint32_t spp);

/* --- the render call (mg_ray_tracer.cpp:178-205 / sm_ray_tracer.cpp:263-289) ------------ */
/* accum (nullable): width*height*4 floats {sum_r,sum_g,sum_b,n_samples}, written inside the
 * tile only.  rgba8 (nullable): width*height uint32, written inside the tile only, resolved
 * as sum / samples_per_pixel.  segments (nullable): number of path segments traced.
 * row_step > 1 renders only rows y0, y0+row_step, ... (bounded CPU-baseline samples).
 * threads <= 0 means "all online cores".                                                  */
int rtref_render(const rtref_scene* s, const rtref_view* v, uint32_t* rgba8, float* accum,
                 uint64_t* segments, int threads, uint32_t row_step);

/* per-sample radiance for one pixel (recursive trace, right-nested product); out[3]; returns
 * the number of segments of that path */
uint32_t rtref_trace_sample(const rtref_scene* s, const rtref_view* v, uint32_t px, uint32_t py,
                            uint32_t sample, float out[3]);

/* --- rasterizer.cpp:22-88: one ray per pixel through the pixel centre, nearest of planes, boxes, spheres (in that
 * order, strict '<', no minimum distance), N.L shading against the eye, no gamma.  Uses inv_view_proj, width, height and
 * the tile of the view; prim (nullable) receives per pixel the sphere index, 0x80000000|plane, 0x40000000|box or
 * 0xFFFFFFFF, depth (nullable) the accepted distance (max_dist + 1 on a miss).                                      */
int rtref_rasterize(const rtref_scene* s, const rtref_view* v, uint32_t* rgba8, uint32_t* prim, float* depth,
                    int threads, uint32_t row_step);
/* S13: ray against a centre/extents box; returns 1 and *t on a hit */
int rtref_ray_hits_box(const float o[3], const float d[3], const float box[6], float* t);

const char* rtref_build_flavour(void); /* "strict" or "fast" */

#ifdef __cplusplus
}
#endif
#endif
