// raster.cuh -- the preview renderer: reference src/renderers/rasterizer.cpp:22-88 as one CUDA kernel.
//
// One thread per pixel, one ray through the pixel centre, the nearest of planes, boxes and spheres in the reference's
// visiting order with its acceptance rule (`!hit || *hit >= dist` rejects: strict '<', first index wins ties, no
// minimum distance, so a sphere entirely behind the near plane is accepted with a negative distance exactly as the
// reference does), then N.L shading against the eye.  Arithmetic follows the numbered SPEC of spec.cuh (S1-S5, S7, S11)
// plus S13 (ray against box) below; results are bit-identical to the CPU checker's restatement and to the reference's own
// rasterizer.cpp compiled against the muu stand-in (tests/test_gpu_raster.py).
//
// The sphere sweep is the same packed-FP32 pair test as the path tracer's linear scan (FFMA2/FADD2/FMUL2, two spheres per
// step, uniform 16-byte loads); it is the only part whose cost grows with the scene, 11 packed + ~6 scalar instructions
// per pair and pixel.
#pragma once
#include "kernels.cuh"
#include <math_constants.h>

namespace rtcu_dev {

#ifndef RTCU_PRIM_BOX
#define RTCU_PRIM_BOX 0x40000000u
#endif

struct RasterScene {
    const float4* spheres;          // {cx,cy,cz,r*r}
    const float4* pairs;            // packed-scan layout, see SceneDev::pairs
    const uint32_t* sphere_material;
    uint32_t n_spheres;
    const float4* planes;           // {nx,ny,nz,d}
    const uint32_t* plane_material;
    uint32_t n_planes;
    const float4* boxes;            // 2 float4 per box: {lo.xyz,0},{hi.xyz,0}; lo = c - e, hi = c + e (one IEEE op each, host side)
    const uint32_t* box_material;
    uint32_t n_boxes;
    const float4* albedo;           // materials.albedo() rgba per material (MatRec only keeps albedo*reflectivity)
};

struct RasterParams {
    CameraConst cam;
    uint32_t width, height;
    uint32_t tile_x0, tile_y0, tile_x1, tile_y1;
    uint32_t* rgba8;
    uint32_t* prim;   // nullable
    float* depth;     // nullable
};

constexpr int RASTER_TILE_W = 32, RASTER_TILE_H = 8;

// S13: muu ray::hits(bounding_box) (rasterizer.cpp:47), Game-Physics-Cookbook slab form: per axis t1 = (lo - o)/d,
// t2 = (hi - o)/d by IEEE division; tmin = max of the per-axis minima, tmax = min of the maxima (fminf/fmaxf drop a NaN
// operand, as in C); tmax < 0 or tmin > tmax -> miss; origin inside (tmin < 0) -> tmax, else tmin.
__device__ __forceinline__ bool box_test(const float4 lo, const float4 hi, const Ray& r, float& t)
{
    const float t1x = __fdiv_rn(__fsub_rn(lo.x, r.o.x), r.d.x), t2x = __fdiv_rn(__fsub_rn(hi.x, r.o.x), r.d.x);
    const float t1y = __fdiv_rn(__fsub_rn(lo.y, r.o.y), r.d.y), t2y = __fdiv_rn(__fsub_rn(hi.y, r.o.y), r.d.y);
    const float t1z = __fdiv_rn(__fsub_rn(lo.z, r.o.z), r.d.z), t2z = __fdiv_rn(__fsub_rn(hi.z, r.o.z), r.d.z);
    float tmin = -CUDART_INF_F, tmax = CUDART_INF_F;
    tmin = fmaxf(tmin, fminf(t1x, t2x)); tmax = fminf(tmax, fmaxf(t1x, t2x));
    tmin = fmaxf(tmin, fminf(t1y, t2y)); tmax = fminf(tmax, fmaxf(t1y, t2y));
    tmin = fmaxf(tmin, fminf(t1z, t2z)); tmax = fminf(tmax, fmaxf(t1z, t2z));
    if (tmax < 0.0f || tmin > tmax) return false;
    t = (tmin < 0.0f) ? tmax : tmin;
    return true;
}

// rasterizer.cpp:47-52 for one sphere whose S4 terms are already known (disc >= 0)
__device__ __forceinline__ void raster_sphere_candidate(const float a, const float e2, const float r2, const float disc, const uint32_t index,
                                                        float& dist, uint32_t& prim)
{
    const float f = __fsqrt_rn(disc);
    const float t = (e2 < r2) ? __fadd_rn(a, f) : __fsub_rn(a, f);
    if (!(t >= dist))
    {
        dist = t;
        prim = index;
    }
}

__device__ __forceinline__ void raster_sphere_pair(const float4 A, const float4 B, const uint32_t pair, const Ray& r, float& dist, uint32_t& prim)
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
        if (!(disc.x < 0.0f)) raster_sphere_candidate(a.x, e2.x, B.z, disc.x, 2 * pair, dist, prim);
        if (!(disc.y < 0.0f)) raster_sphere_candidate(a.y, e2.y, B.w, disc.y, 2 * pair + 1, dist, prim);
    }
}

// BVH: the sphere loop becomes a traversal of the path tracer's tree (same exact-result construction, see kernels.cuh)
// started with best = (distance accepted so far, index -1): ties with a plane or box keep the earlier category, ties
// between spheres go to the lower index -- what the index-ordered loop with its strict '<' produces.
template <bool BVH>
__global__ void __launch_bounds__(RASTER_TILE_W* RASTER_TILE_H) k_rasterize(const RasterScene s, const SceneDev sc, const RasterParams p)
{
    const uint32_t x = p.tile_x0 + blockIdx.x * RASTER_TILE_W + (threadIdx.x % RASTER_TILE_W);
    const uint32_t y = p.tile_y0 + blockIdx.y * RASTER_TILE_H + (threadIdx.x / RASTER_TILE_W);
    if (x >= p.tile_x1 || y >= p.tile_y1) return;

    // :30-39
    V3 near_p, far_p;
    primary_points(p.cam, __fadd_rn((float)x, 0.5f), __fadd_rn((float)y, 0.5f), near_p, far_p);
    const V3 span = v3_sub(far_p, near_p);
    const float len2 = dot3(span, span);
    float max_dist, inv_len;
    sqrt_then_rcp(len2, max_dist, inv_len);                  // vec3::distance, and S2 on the same dot product
    float dist = __fadd_rn(max_dist, 1.0f);
    Ray r;
    r.o = near_p;
    r.d = v3_scale(span, inv_len);                           // vec3::direction

    uint32_t prim = RTCU_PRIM_MISS;
    int last_plane = -1;

    // hit_tests(scene.planes), :61
    for (uint32_t i = 0; i < s.n_planes; i++)
    {
        const float4 pl = __ldg(s.planes + i);
        const V3 n = v3(pl.x, pl.y, pl.z);
        const float nd = dot3(r.d, n);
        if (!(nd >= 0.0f))
        {
            const float t = __fdiv_rn(__fsub_rn(-pl.w, dot3(r.o, n)), nd);
            if (!(t < 0.0f) && !(t >= dist))
            {
                dist = t;
                prim = RTCU_PRIM_PLANE | i;
                last_plane = (int)i;
            }
        }
    }
    // hit_tests(scene.boxes), :62
    for (uint32_t i = 0; i < s.n_boxes; i++)
    {
        float t;
        if (box_test(__ldg(s.boxes + 2 * i), __ldg(s.boxes + 2 * i + 1), r, t) && !(t >= dist))
        {
            dist = t;
            prim = RTCU_PRIM_BOX | i;
        }
    }
    // hit_tests(scene.spheres), :63
    Trav tv;
    if (BVH && trav_init(r, tv))
    {
        tv.best_t = dist;
        tv.best_i = -1;
        uint32_t stack_ref[BVH_STACK];
        float stack_t[BVH_STACK];
        BvhStats st = { 0, 0 };
        while (!trav_step<true>(sc, r, tv, stack_ref, stack_t, st))
        {
        }
        if (tv.best_i >= 0)
        {
            dist = tv.best_t;
            prim = (uint32_t)tv.best_i;
        }
    }
    else
    {
        // packed pair sweep, the next pair in flight while this one is tested
        const uint32_t n_pairs = (s.n_spheres + 1u) >> 1;
        uint32_t sphere = RTCU_PRIM_MISS;
        float4 A = __ldg(s.pairs), B = __ldg(s.pairs + 1);
#pragma unroll 4
        for (uint32_t j = 0; j < n_pairs; j++)
        {
            const float4 An = __ldg(s.pairs + 2 * j + 2), Bn = __ldg(s.pairs + 2 * j + 3); // padded: always in bounds
            raster_sphere_pair(A, B, j, r, dist, sphere);
            A = An;
            B = Bn;
        }
        if (sphere != RTCU_PRIM_MISS) prim = sphere;
    }

    const size_t pixel = (size_t)y * p.width + x;
    if (p.prim) p.prim[pixel] = prim;
    if (p.depth) p.depth[pixel] = dist;

    float c[3];
    if (prim != RTCU_PRIM_MISS)
    {
        const V3 hit_pos = ray_at(r.o, r.d, dist); // :53
        V3 n = v3(0.0f, 1.0f, 0.0f);               // vec3::constants::up, :38
        uint32_t material;
        if (prim & RTCU_PRIM_PLANE)
        {
            const float4 pl = __ldg(s.planes + (prim & 0x3FFFFFFFu));
            n = v3(pl.x, pl.y, pl.z);              // :58
            material = __ldg(s.plane_material + (prim & 0x3FFFFFFFu));
        }
        else if (prim & RTCU_PRIM_BOX)
        {
            // :55-58 have no box branch: the normal stays what the last accepted plane left, else `up`
            if (last_plane >= 0)
            {
                const float4 pl = __ldg(s.planes + last_plane);
                n = v3(pl.x, pl.y, pl.z);
            }
            material = __ldg(s.box_material + (prim & 0x3FFFFFFFu));
        }
        else
        {
            const float4 sp = __ldg(s.spheres + prim);
            n = normalize3(v3_sub(hit_pos, v3(sp.x, sp.y, sp.z))); // :56
            material = __ldg(s.sphere_material + prim);
        }
        // :70-76 with lambert() of :14-20: min(0.25 + (l.n * albedo * 1.0f) * 0.75, 1), source expressions unfused (rule R)
        const V3 l = normalize3(v3_sub(near_p, hit_pos));
        const float k = dot3(l, n);
        const float4 al = __ldg(s.albedo + material);
        const float av[3] = { al.x, al.y, al.z };
#pragma unroll
        for (int i = 0; i < 3; i++)
        {
            const float lam = __fmul_rn(__fmul_rn(k, av[i]), 1.0f);
            const float v = __fadd_rn(0.25f, __fmul_rn(lam, 0.75f));
            c[i] = (v < 1.0f) ? v : 1.0f;
        }
    }
    else
    {
        // :65-66, :79-82: both sky colours are int-constructed and saturate to white (colour.hpp:64-83); lerp in S8 form
        const float a = __fdiv_rn((float)y, (float)(p.height - 1u));
        const float w = __fsub_rn(1.0f, a);
        c[0] = c[1] = c[2] = __fmaf_rn(1.0f, a, __fmul_rn(1.0f, w));
    }
    p.rgba8[pixel] = (to_byte(c[0]) << 24) | (to_byte(c[1]) << 16) | (to_byte(c[2]) << 8) | to_byte(1.0f);
}

} // namespace rtcu_dev
