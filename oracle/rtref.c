/*
 * rtref.c -- CPU ORACLE for the marzer/rt path-tracing hot path.  TEST INFRASTRUCTURE ONLY
 * (see rtref.h for who may load it).  Pinned bit for bit against the reference's own renderer
 * sources compiled against a muu stand-in (oracle/Makefile `ref`); muu's arithmetic itself is
 * restated (SPEC below) and UNPINNED -- see rtref.h.
 *
 * Every function cites the reference file:line it follows (paths relative to /root/reference).
 *
 * ------------------------------------------------------------------------------------------
 * ARITHMETIC SPEC (normative; the CUDA kernels implement the same numbered rules independently)
 *
 * All arithmetic is IEEE-754 binary32, round-to-nearest-even, no flush-to-zero, except S10's
 * Schlick polynomial which is binary64 as in the reference source.  fma(a,b,c) is a single
 * rounding.  The reference is built with -ffast-math -ffp-contract=fast (meson.build:153-160),
 * so its own contraction is compiler-chosen; this spec fixes one legal choice.  The strict
 * build of this file uses -ffp-contract=off so only the fmaf() written below fuses.
 *
 *  R   expressions written in the reference's own source (reflect, the metal and lambert sums, refract, the
 *      Schlick polynomial, attenuation products, the per-pixel sum) are evaluated exactly as written, left to
 *      right, one rounding per operator, NO contraction.  Only the muu primitives below (which the reference calls
 *      but does not contain) use fused multiply-adds; oracle/ref_shim implements them identically so that the
 *      reference's renderer sources, compiled against the stand-in, reproduce this file bit for bit.
 *  S1  dot3(a,b)      = fma(a.z,b.z, fma(a.y,b.y, a.x*b.x))
 *  S2  normalize(v)   = v * (1.0f / sqrtf(dot3(v,v)))        (one division, three multiplies)
 *  S3  at(o,d,t)      = fma(d, t, o) per component            (muu ray::at = origin + dir*t)
 *  S4  sphere hit     e = c - o; e2 = dot3(e,e); r2 = r*r; a = dot3(e,d);
 *                     disc = r2 - fma(-a, a, e2); disc < 0 -> miss; f = sqrtf(disc);
 *                     t = (e2 < r2) ? a + f : a - f           (muu ray::hits(bounding_sphere),
 *                     Game-Physics-Cookbook form, SURVEY 8a-3, UNVERIFIED against muu source)
 *  S5  plane hit      nd = dot3(d,n); nd >= 0 -> miss; t = (-pd - dot3(o,n)) / nd; t < 0 -> miss
 *  S6  closest        ascending index; candidate rejected if (t < 0.001f) or (have && best <= t)
 *                     (mg_ray_tracer.cpp:46-52, :73-79); sphere beats plane on ties (:95-102)
 *  S7  camera         q.x = px/W, q.y = py/H (IEEE div); ndc = (fma(2,q.x,-1), fma(-2,q.y,1), z);
 *                     col_k = column k of invVP; h = fma(col2, z, fma(col1, ndc.y, fma(col0, ndc.x, col3)));
 *                     p = h.xyz * (1.0f / h.w)                (camera.hpp:42-48)
 *  S8  sky            a = 0.5f*(d.y + 1.0f); c = fma(sky_k, a, 1.0f*(1.0f - a)), sky=(0.5,0.7,1.0)
 *                     (muu lerp = start*(1-alpha) + finish*alpha)  (mg_ray_tracer.cpp:163-164)
 *  S9  RNG            Philox4x32-7, key = (seed_lo, seed_hi), ctr = (pixel, sample, block, retry);
 *                     block 0 = pixel jitter (out[0], out[1]); block k+1 = scatter at the end of
 *                     segment k (lambert/metal out[0..2] = x,y,z; dielectric out[0]).
 *                     float = (x >> 8) * 2^-24.  retry increments only when a unit-vector draw is
 *                     exactly (0,0,0) (random.hpp:57-66).
 *  S10 schlick        r0 = ((1-ior)/(1+ior))^2 in fp32; x = (double)(1.0f - cos);
 *                     p = (float)((double)r0 + (double)(1.0f - r0) * ((x*x)*(x*x)*x))
 *                     (sm_ray_tracer.cpp:174-179; pow(double,5) restated as three multiplies)
 *  S11 resolve        c = sum / (float)spp; c = sqrtf(c); c = min(max(c,0),1);
 *                     byte = (uint32)(c * 255.99999f); rgba = r<<24 | g<<16 | b<<8 | 255
 *                     (mg_ray_tracer.cpp:195-200, colour.hpp:100-106)
 *  S12 radiance       recursive, right-nested: att_1 * (att_2 * (... * sky))  (mg_ray_tracer.cpp:171)
 *                     per pixel: sum += radiance in ascending sample order      (:187-194)
 *  S13 ray vs box     (rasterizer.cpp:47 only) lo = c - e, hi = c + e; per axis t1 = (lo-o)/d, t2 = (hi-o)/d;
 *                     tmin = max_k fminf(t1,t2), tmax = min_k fmaxf(t1,t2) (a NaN operand is dropped);
 *                     tmax < 0 or tmin > tmax -> miss; t = tmin < 0 ? tmax : tmin
 *                     (muu ray::hits(bounding_box), UNVERIFIED; Game-Physics-Cookbook form like S4/S5)
 * ------------------------------------------------------------------------------------------
 */
#define _GNU_SOURCE
#include "rtref.h"
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#ifdef RTREF_FAST
#define FMA(a, b, c) ((a) * (b) + (c)) /* fast build: let -ffp-contract=fast decide, as the reference does */
#else
#define FMA(a, b, c) fmaf((a), (b), (c))
#endif

typedef struct { float x, y, z; } v3;

static inline v3 v3_make(float x, float y, float z) { v3 r = { x, y, z }; return r; }
static inline v3 v3_add(v3 a, v3 b) { return v3_make(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 v3_sub(v3 a, v3 b) { return v3_make(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 v3_scale(v3 a, float s) { return v3_make(a.x * s, a.y * s, a.z * s); }
static inline v3 v3_neg(v3 a) { return v3_make(-a.x, -a.y, -a.z); }
/* S1 */
static inline float dot3(v3 a, v3 b) { return FMA(a.z, b.z, FMA(a.y, b.y, a.x * b.x)); }
/* S2: muu vector::normalize, called from mg_ray_tracer.cpp:85,:120,:133,:138,:193, random.hpp:64 */
static inline v3 normalize3(v3 v)
{
    const float inv = 1.0f / sqrtf(dot3(v, v));
    return v3_scale(v, inv);
}
/* S3: muu ray::at, mg_ray_tracer.cpp:85,:122,:139 */
static inline v3 ray_at(v3 o, v3 d, float t) { return v3_make(FMA(d.x, t, o.x), FMA(d.y, t, o.y), FMA(d.z, t, o.z)); }

typedef struct { v3 o, d; } ray_t;

/* ---- S9: Philox4x32-R (Salmon et al., SC'11; Random123 constants).  The stream of the SPEC uses R = 7 rounds, the fewest
 * the paper certifies as Crush-resistant (Random123's philox4x32_7; its default of 10 is a safety margin): the path tracer draws
 * one block per sample and one per scatter event, and the generator is 8-15 % of the device's instructions.  Both round
 * counts are pinned against Random123's kat_vectors (tests/test_oracle_kat.py). ----------------------- */
#define RTREF_PHILOX_ROUNDS 7
void rtref_philox4x32_r(const uint32_t ctr[4], const uint32_t key[2], int rounds, uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int round = 0; round < rounds; round++)
    {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
/* one block of the SPEC's stream */
void rtref_philox_stream(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    rtref_philox4x32_r(ctr, key, RTREF_PHILOX_ROUNDS, out);
}

float rtref_u01(uint32_t x) { return (float)(x >> 8) * 0x1.0p-24f; }

typedef struct { uint64_t seed; uint32_t pixel, sample; } rng_key;

static void rng_block(const rng_key* k, uint32_t block, uint32_t retry, float u[4])
{
    const uint32_t ctr[4] = { k->pixel, k->sample, block, retry };
    const uint32_t key[2] = { (uint32_t)k->seed, (uint32_t)(k->seed >> 32) };
    uint32_t out[4];
    rtref_philox_stream(ctr, key, out);
    for (int i = 0; i < 4; i++)
        u[i] = rtref_u01(out[i]);
}

/* random.hpp:57-66: normalize(U[0,1)^3), positive octant only, redraw iff exactly zero */
static v3 random_unit_vector(const rng_key* k, uint32_t block)
{
    for (uint32_t retry = 0;; retry++)
    {
        float u[4];
        rng_block(k, block, retry, u);
        if (u[0] == 0.0f && u[1] == 0.0f && u[2] == 0.0f)
            continue;
        return normalize3(v3_make(u[0], u[1], u[2]));
    }
}

/* ---- hit_result, mg_ray_tracer.cpp:22-33 -------------------------------------------------- */
typedef struct { float distance; v3 normal; uint32_t material; uint32_t prim; } hit_result;
#define MIN_HIT_DIST 0.001f /* mg_ray_tracer.cpp:20 */
#define PRIM_MISS 0xFFFFFFFFu
#define PRIM_PLANE 0x80000000u

static inline int hit_ok(const hit_result* h) { return h->distance >= 0.0f; } /* :29-32 */

/* S5: muu ray::hits(plane), called at mg_ray_tracer.cpp:46.  returns 1 and *t on hit */
static inline int ray_hits_plane(const ray_t* r, const float* pl, float* t)
{
    const v3 n = v3_make(pl[0], pl[1], pl[2]);
    const float nd = dot3(r->d, n);
    if (nd >= 0.0f)
        return 0;
    const float tt = (-pl[3] - dot3(r->o, n)) / nd;
    if (tt < 0.0f)
        return 0;
    *t = tt;
    return 1;
}

/* S4: muu ray::hits(bounding_sphere), called at mg_ray_tracer.cpp:73 */
static inline int ray_hits_sphere(const ray_t* r, const float* sp, float* t)
{
    const v3 e = v3_sub(v3_make(sp[0], sp[1], sp[2]), r->o);
    const float e2 = dot3(e, e);
    const float r2 = sp[3] * sp[3];
    const float a = dot3(e, r->d);
    const float disc = r2 - FMA(-a, a, e2);
    if (disc < 0.0f)
        return 0;
    const float f = sqrtf(disc);
    *t = (e2 < r2) ? a + f : a - f;
    return 1;
}

/* mg_ray_tracer.cpp:35-60 */
static hit_result test_planes(const rtref_scene* s, const ray_t* r)
{
    int have = 0;
    uint32_t hit_index = 0;
    float hit_dist = 0.0f;
    for (uint32_t i = 0; i < s->n_planes; i++)
    {
        float t;
        if (!ray_hits_plane(r, s->planes + 4 * (size_t)i, &t) || t < MIN_HIT_DIST || (have && hit_dist <= t))
            continue;
        have = 1;
        hit_index = i;
        hit_dist = t;
    }
    hit_result h = { -1.0f, { 0, 0, 0 }, 0, PRIM_MISS };
    if (!have)
        return h;
    const float* pl = s->planes + 4 * (size_t)hit_index;
    h.distance = hit_dist;
    h.normal = v3_make(pl[0], pl[1], pl[2]);
    h.material = s->plane_material[hit_index];
    h.prim = PRIM_PLANE | hit_index;
    return h;
}

/* mg_ray_tracer.cpp:62-87 */
static hit_result test_spheres(const rtref_scene* s, const ray_t* r)
{
    int have = 0;
    uint32_t hit_index = 0;
    float hit_dist = 0.0f;
    for (uint32_t i = 0; i < s->n_spheres; i++)
    {
        float t;
        if (!ray_hits_sphere(r, s->spheres + 4 * (size_t)i, &t) || t < MIN_HIT_DIST || (have && hit_dist <= t))
            continue;
        have = 1;
        hit_index = i;
        hit_dist = t;
    }
    hit_result h = { -1.0f, { 0, 0, 0 }, 0, PRIM_MISS };
    if (!have)
        return h;
    const float* sp = s->spheres + 4 * (size_t)hit_index;
    h.distance = hit_dist;
    /* vec3::direction(center, r.at(t)) = normalize(at - center), always outward (:85) */
    h.normal = normalize3(v3_sub(ray_at(r->o, r->d, hit_dist), v3_make(sp[0], sp[1], sp[2])));
    h.material = s->sphere_material[hit_index];
    h.prim = hit_index;
    return h;
}

/* mg_ray_tracer.cpp:89-93: boxes never hit */
static hit_result test_boxes(void)
{
    hit_result h = { -1.0f, { 0, 0, 0 }, 0, PRIM_MISS };
    return h;
}

/* mg_ray_tracer.cpp:95-102 */
static hit_result select_hit(hit_result a, hit_result b)
{
    if (!hit_ok(&a))
        return b;
    return (!hit_ok(&b) || a.distance <= b.distance) ? a : b;
}

/* mg_ray_tracer.cpp:160-162 */
static hit_result closest_hit(const rtref_scene* s, const ray_t* r)
{
    hit_result hit = test_planes(s, r);
    hit = select_hit(test_spheres(s, r), hit);
    hit = select_hit(test_boxes(), hit);
    return hit;
}

/* common.hpp:99-103 / sm_ray_tracer.cpp:156-159: v - 2*dot(v,n)*n */
static inline v3 reflect3(v3 v, v3 n)
{
    const float k = 2.0f * dot3(v, n);
    return v3_sub(v, v3_scale(n, k)); /* source expression, evaluated without contraction (SPEC rule R) */
}

/* attenuation = vec3{albedo * reflectivity}: mg_ray_tracer.cpp:115,:131, sm_ray_tracer.cpp:194 */
static inline v3 attenuation_of(const rtref_material* m)
{
    return v3_make(m->albedo[0] * m->reflectivity, m->albedo[1] * m->reflectivity, m->albedo[2] * m->reflectivity);
}

#define APPROX_ZERO_EPS 1e-5f /* muu vector::approx_zero default epsilon, UNVERIFIED; measure-zero branch */

/* mg_ray_tracer.cpp:109-123 */
static int lambert_scatter(const rtref_material* m, const ray_t* r, const hit_result* hit, const rng_key* k,
                           uint32_t block, v3* att, ray_t* out)
{
    *att = attenuation_of(m);
    v3 s = v3_add(hit->normal, random_unit_vector(k, block));
    if (fabsf(s.x) < APPROX_ZERO_EPS && fabsf(s.y) < APPROX_ZERO_EPS && fabsf(s.z) < APPROX_ZERO_EPS)
        s = hit->normal;
    s = normalize3(s);
    out->o = ray_at(r->o, r->d, hit->distance);
    out->d = s;
    return 1;
}

/* mg_ray_tracer.cpp:125-140 */
static int metal_scatter(const rtref_material* m, const ray_t* r, const hit_result* hit, const rng_key* k,
                         uint32_t block, v3* att, ray_t* out)
{
    *att = attenuation_of(m);
    const v3 refl = reflect3(normalize3(r->d), hit->normal);
    const v3 u = random_unit_vector(k, block);
    v3 s = v3_add(refl, v3_scale(u, m->roughness)); /* reflect(...) + roughness * random_unit_vector(), rule R */
    if (dot3(s, hit->normal) <= 0.0f)
        return 0;
    s = normalize3(s);
    out->o = ray_at(r->o, r->d, hit->distance);
    out->d = s;
    return 1;
}

/* sm_ray_tracer.cpp:161-172 */
static int refract3(v3 v, v3 n, float eta, v3* refracted)
{
    const float cos_i = -dot3(v, n);
    const float sin2_t = (eta * eta) * (1.0f - cos_i * cos_i); /* rule R: eta * eta * (1 - cos_i * cos_i) */
    if (sin2_t > 1.0f)
        return 0;
    const float cos_t = sqrtf(1.0f - sin2_t);
    const float k = eta * cos_i - cos_t;
    *refracted = v3_add(v3_scale(v, eta), v3_scale(n, k)); /* eta * v + (eta * cos_i - cos_t) * n */
    return 1;
}

/* S10, sm_ray_tracer.cpp:174-179 */
static float schlick(float cosine, float ref_idx)
{
    float r0 = (1.0f - ref_idx) / (1.0f + ref_idx);
    r0 = r0 * r0;
    const double x = (double)(1.0f - cosine);
    const double x2 = x * x;
    const double x5 = (x2 * x2) * x;
    const double w = (double)(1.0f - r0);
    const double prod = w * x5; /* kept as a separate statement: no fused multiply-add with the sum */
    return (float)((double)r0 + prod);
}

/* sm_ray_tracer.cpp:181-219 */
static int dielectric_scatter(const rtref_material* m, const ray_t* r, const hit_result* hit, const rng_key* k,
                              uint32_t block, v3* att, ray_t* out)
{
    const v3 reflected = reflect3(r->d, hit->normal);
    const float ior = m->reflectivity;
    *att = attenuation_of(m);

    const float dn = dot3(r->d, hit->normal);
    const float len = sqrtf(dot3(r->d, r->d));
    v3 outward;
    float ni_over_nt, cosine;
    if (dn > 0.0f)
    {
        outward = v3_neg(hit->normal);
        ni_over_nt = ior;
        cosine = (ior * dn) / len;
    }
    else
    {
        outward = hit->normal;
        ni_over_nt = 1.0f / ior;
        cosine = (-dn) / len;
    }

    v3 refracted = { 0, 0, 0 };
    float reflect_prob;
    if (refract3(r->d, outward, ni_over_nt, &refracted))
        reflect_prob = schlick(cosine, ior);
    else
        reflect_prob = 1.0f;

    float u[4];
    rng_block(k, block, 0, u);
    out->o = ray_at(r->o, r->d, hit->distance);
    out->d = (u[0] < reflect_prob) ? reflected : refracted;
    return 1;
}

typedef enum { SC_LAMBERT, SC_METAL, SC_DIELECTRIC } scatter_kind;

/* mg_ray_tracer.cpp:142-152 (mode 0) and sm_ray_tracer.cpp:221-236 (mode 1) */
static scatter_kind scatter_table(uint32_t mode, uint32_t type)
{
    if (type == RTREF_METAL)
        return SC_METAL;
    if (mode == RTREF_MODE_SM && type >= RTREF_DIELECTRIC && type <= RTREF_ICE)
        return SC_DIELECTRIC;
    return SC_LAMBERT; /* "default to lambert for unimplemented brdfs" -- includes diamond */
}

static int scatter(const rtref_scene* s, uint32_t mode, const ray_t* r, const hit_result* hit, const rng_key* k,
                   uint32_t block, v3* att, ray_t* out)
{
    const rtref_material* m = &s->materials[hit->material];
    switch (scatter_table(mode, m->type))
    {
        case SC_METAL: return metal_scatter(m, r, hit, k, block, att, out);
        case SC_DIELECTRIC: return dielectric_scatter(m, r, hit, k, block, att, out);
        default: return lambert_scatter(m, r, hit, k, block, att, out);
    }
}

/* S8, mg_ray_tracer.cpp:163-164 */
static inline v3 sky(v3 d)
{
    const float a = 0.5f * (d.y + 1.0f);
    const float w = 1.0f - a;
    return v3_make(FMA(0.5f, a, 1.0f * w), FMA(0.7f, a, 1.0f * w), FMA(1.0f, a, 1.0f * w));
}

/* mg_ray_tracer.cpp:154-174 (sm_ray_tracer.cpp:238-261); segment = recursion depth */
static v3 trace(const rtref_scene* s, uint32_t mode, ray_t r, unsigned max_bounces, const rng_key* k,
                uint32_t segment, uint32_t* n_segments)
{
    if (!(max_bounces--))
        return v3_make(0, 0, 0);
    (*n_segments)++;

    const hit_result hit = closest_hit(s, &r);
    if (!hit_ok(&hit))
        return sky(r.d);

    v3 att;
    ray_t next;
    if (scatter(s, mode, &r, &hit, k, segment + 1, &att, &next))
    {
        const v3 in = trace(s, mode, next, max_bounces, k, segment + 1, n_segments);
        return v3_make(att.x * in.x, att.y * in.y, att.z * in.z);
    }
    return v3_make(0, 0, 0);
}

/* S7: viewport::screen_to_world, camera.hpp:42-48 (muu matrix::transform_position: M*(p,1) / w) */
static v3 screen_to_world(const rtref_view* v, float sx, float sy, float depth)
{
    const float* m = v->inv_view_proj;
    const float qx = sx / (float)v->width;
    const float qy = sy / (float)v->height;
    const float nx = FMA(2.0f, qx, -1.0f);
    const float ny = FMA(-2.0f, qy, 1.0f);
    float h[4];
    for (int r = 0; r < 4; r++)
        h[r] = FMA(m[8 + r], depth, FMA(m[4 + r], ny, FMA(m[0 + r], nx, m[12 + r])));
    const float inv_w = 1.0f / h[3];
    return v3_make(h[0] * inv_w, h[1] * inv_w, h[2] * inv_w);
}

/* mg_ray_tracer.cpp:189-193 */
static ray_t primary_ray(const rtref_view* v, uint32_t px, uint32_t py, const rng_key* k)
{
    float jx = 0.5f, jy = 0.5f; /* sample 0 goes through the pixel centre (:189) */
    if (k->sample != 0)
    {
        float u[4];
        rng_block(k, 0, 0, u);
        jx = u[0];
        jy = u[1];
    }
    const float sx = (float)px + jx;
    const float sy = (float)py + jy;
    const v3 near_p = screen_to_world(v, sx, sy, 0.0f);
    const v3 far_p = screen_to_world(v, sx, sy, 1.0f);
    ray_t r;
    r.o = near_p;
    r.d = normalize3(v3_sub(far_p, near_p)); /* vec3::direction(near, far) */
    return r;
}

void rtref_primary_ray(const rtref_view* v, uint32_t px, uint32_t py, uint32_t sample, float o[3], float d[3])
{
    const rng_key k = { v->seed, py * v->width + px, sample };
    const ray_t r = primary_ray(v, px, py, &k);
    o[0] = r.o.x; o[1] = r.o.y; o[2] = r.o.z;
    d[0] = r.d.x; d[1] = r.d.y; d[2] = r.d.z;
}

/* the primary ray through an explicit screen position (what primary_ray computes once the jitter is known): lets a test place
 * rays on a pixel's corners and edges */
void rtref_screen_ray(const rtref_view* v, float sx, float sy, float o[3], float d[3])
{
    const v3 near_p = screen_to_world(v, sx, sy, 0.0f);
    const v3 far_p = screen_to_world(v, sx, sy, 1.0f);
    const v3 dir = normalize3(v3_sub(far_p, near_p));
    o[0] = near_p.x; o[1] = near_p.y; o[2] = near_p.z;
    d[0] = dir.x; d[1] = dir.y; d[2] = dir.z;
}

int rtref_scatter(const rtref_scene* s, uint32_t mode, uint32_t material, const float o[3], const float d[3], float t,
                  const float n[3], uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t block, float att[3],
                  float o_out[3], float d_out[3])
{
    const ray_t r = { { o[0], o[1], o[2] }, { d[0], d[1], d[2] } };
    const hit_result hit = { t, { n[0], n[1], n[2] }, material, 0 };
    const rng_key k = { seed, pixel, sample };
    v3 a = { 0, 0, 0 };
    ray_t out = { { 0, 0, 0 }, { 0, 0, 0 } };
    const int ok = scatter(s, mode, &r, &hit, &k, block, &a, &out);
    att[0] = a.x; att[1] = a.y; att[2] = a.z;
    o_out[0] = out.o.x; o_out[1] = out.o.y; o_out[2] = out.o.z;
    d_out[0] = out.d.x; d_out[1] = out.d.y; d_out[2] = out.d.z;
    return ok;
}

int rtref_intersect_batch(const rtref_scene* s, const float* o, const float* d, uint32_t n, uint8_t* hit,
                          uint32_t* prim, float* t, float* nrm)
{
    for (uint32_t i = 0; i < n; i++)
    {
        const ray_t r = { { o[3 * i], o[3 * i + 1], o[3 * i + 2] }, { d[3 * i], d[3 * i + 1], d[3 * i + 2] } };
        const hit_result h = closest_hit(s, &r);
        hit[i] = (uint8_t)hit_ok(&h);
        prim[i] = hit_ok(&h) ? h.prim : PRIM_MISS;
        t[i] = h.distance;
        if (nrm)
        {
            nrm[3 * i] = h.normal.x; nrm[3 * i + 1] = h.normal.y; nrm[3 * i + 2] = h.normal.z;
        }
    }
    return 0;
}

/* S11 */
static inline uint32_t to_byte(float c)
{
    c = fminf(fmaxf(c, 0.0f), 1.0f);
    return (uint32_t)(c * 255.99999f);
}

uint32_t rtref_pack_pixel(float sum_r, float sum_g, float sum_b, uint32_t spp)
{
    const float n = (float)spp;
    const float r = sqrtf(sum_r / n), g = sqrtf(sum_g / n), b = sqrtf(sum_b / n);
    return (to_byte(r) << 24) | (to_byte(g) << 16) | (to_byte(b) << 8) | to_byte(1.0f);
}

uint32_t rtref_trace_sample(const rtref_scene* s, const rtref_view* v, uint32_t px, uint32_t py, uint32_t sample,
                            float out[3])
{
    const rng_key k = { v->seed, py * v->width + px, sample };
    uint32_t nseg = 0;
    const v3 c = trace(s, v->material_mode, primary_ray(v, px, py, &k), v->max_bounces, &k, 0, &nseg);
    out[0] = c.x; out[1] = c.y; out[2] = c.z;
    return nseg;
}

/* ---- render: mg_ray_tracer.cpp:178-205; muu::thread_pool::for_range restated with pthreads ---- */
typedef struct {
    const rtref_scene* s;
    const rtref_view* v;
    uint32_t* rgba8;
    float* accum;
    uint32_t row_step;
    int tid, nthreads;
    uint64_t segments;
} job_t;

static void* render_rows(void* arg)
{
    job_t* j = (job_t*)arg;
    const rtref_view* v = j->v;
    uint64_t segs = 0;
    uint32_t row = 0;
    for (uint32_t y = v->tile_y0; y < v->tile_y1; y += j->row_step, row++)
    {
        if ((int)(row % (uint32_t)j->nthreads) != j->tid)
            continue;
        for (uint32_t x = v->tile_x0; x < v->tile_x1; x++)
        {
            const uint32_t pixel = y * v->width + x;
            v3 colour = { 0, 0, 0 };
            uint64_t pixel_segs = 0;
            for (uint32_t i = v->sample_begin; i < v->sample_end; i++)
            {
                const rng_key k = { v->seed, pixel, i };
                uint32_t nseg = 0;
                const v3 c = trace(j->s, v->material_mode, primary_ray(v, x, y, &k), v->max_bounces, &k, 0, &nseg);
                colour = v3_add(colour, c);
                pixel_segs += nseg;
            }
            segs += pixel_segs;
            if (j->accum)
            {
                float* a = j->accum + 4 * (size_t)pixel;
                a[0] = colour.x; a[1] = colour.y; a[2] = colour.z;
                /* flags bit 8 (oracle-only diagnostic): report the pixel's segment count instead of its sample count */
                a[3] = (v->flags & 0x100u) ? (float)pixel_segs : (float)(v->sample_end - v->sample_begin);
            }
            if (j->rgba8)
                j->rgba8[pixel] = rtref_pack_pixel(colour.x, colour.y, colour.z, v->samples_per_pixel);
        }
    }
    j->segments = segs;
    return NULL;
}

int rtref_render(const rtref_scene* s, const rtref_view* v, uint32_t* rgba8, float* accum, uint64_t* segments,
                 int threads, uint32_t row_step)
{
    if (!s || !v || v->tile_x1 > v->width || v->tile_y1 > v->height || v->sample_end < v->sample_begin)
        return -1;
    for (uint32_t i = 0; i < s->n_spheres; i++)
        if (s->sphere_material[i] >= s->n_materials)
            return -2;
    for (uint32_t i = 0; i < s->n_planes; i++)
        if (s->plane_material[i] >= s->n_materials)
            return -2;
    if (row_step == 0)
        row_step = 1;
    if (threads <= 0)
        threads = (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (threads < 1)
        threads = 1;
    if (threads > 1024)
        threads = 1024;

    job_t* jobs = (job_t*)calloc((size_t)threads, sizeof(job_t));
    pthread_t* tids = (pthread_t*)calloc((size_t)threads, sizeof(pthread_t));
    for (int t = 0; t < threads; t++)
    {
        job_t j = { s, v, rgba8, accum, row_step, t, threads, 0 };
        jobs[t] = j;
        if (t > 0)
            pthread_create(&tids[t], NULL, render_rows, &jobs[t]);
    }
    render_rows(&jobs[0]);
    uint64_t total = jobs[0].segments;
    for (int t = 1; t < threads; t++)
    {
        pthread_join(tids[t], NULL);
        total += jobs[t].segments;
    }
    if (segments)
        *segments = total;
    free(jobs);
    free(tids);
    return 0;
}

/* ==== rasterizer (rasterizer.cpp:22-88) ======================================================================== */

/* S13: muu ray::hits(bounding_box), called at rasterizer.cpp:47.  Game-Physics-Cookbook slab form like S4/S5
 * (UNVERIFIED against muu): lo = c - e, hi = c + e; per axis t1 = (lo - o)/d, t2 = (hi - o)/d (IEEE division, a zero
 * direction component gives +-inf or NaN; fminf/fmaxf drop a NaN operand); tmin = max of the per-axis minima,
 * tmax = min of the maxima; tmax < 0 or tmin > tmax -> miss; tmin < 0 (origin inside) -> tmax, else tmin.           */
int rtref_ray_hits_box(const float o[3], const float d[3], const float box[6], float* t)
{
    float tmin = -INFINITY, tmax = INFINITY;
    for (int k = 0; k < 3; k++)
    {
        const float lo = box[k] - box[3 + k], hi = box[k] + box[3 + k];
        const float t1 = (lo - o[k]) / d[k], t2 = (hi - o[k]) / d[k];
        tmin = fmaxf(tmin, fminf(t1, t2));
        tmax = fminf(tmax, fmaxf(t1, t2));
    }
    if (tmax < 0.0f || tmin > tmax)
        return 0;
    *t = (tmin < 0.0f) ? tmax : tmin;
    return 1;
}

#define PRIM_BOX 0x40000000u

/* the worker lambda, rasterizer.cpp:28-83 */
static uint32_t raster_pixel(const rtref_scene* s, const rtref_view* v, uint32_t px, uint32_t py, uint32_t* prim_out,
                             float* depth_out)
{
    const float sx = (float)px + 0.5f, sy = (float)py + 0.5f;
    const v3 near_p = screen_to_world(v, sx, sy, 0.0f); /* :31 */
    const v3 far_p = screen_to_world(v, sx, sy, 1.0f);  /* :32 */
    const v3 span = v3_sub(far_p, near_p);
    const float max_dist = sqrtf(dot3(span, span));     /* vec3::distance, :33 */
    float dist = max_dist + 1.0f;                       /* :35 */
    uint32_t prim = PRIM_MISS, material = 0;
    v3 hit_pos = { 0, 0, 0 };
    v3 hit_normal = { 0.0f, 1.0f, 0.0f };               /* vec3::constants::up, :38 */
    ray_t r;
    r.o = near_p;
    r.d = normalize3(span);                             /* :39 */

    /* hit_tests(scene.planes), :61 */
    for (uint32_t i = 0; i < s->n_planes; i++)
    {
        const float* pl = s->planes + 4 * (size_t)i;
        float t;
        if (!ray_hits_plane(&r, pl, &t) || t >= dist)   /* :47-49 */
            continue;
        dist = t;
        prim = PRIM_PLANE | i;
        material = s->plane_material[i];
        hit_pos = ray_at(r.o, r.d, t);
        hit_normal = v3_make(pl[0], pl[1], pl[2]);      /* :58 */
    }
    /* hit_tests(scene.boxes), :62 -- a box leaves hit_normal as it was (:55-58 have no box branch) */
    for (uint32_t i = 0; i < s->n_boxes; i++)
    {
        const float o[3] = { r.o.x, r.o.y, r.o.z }, d[3] = { r.d.x, r.d.y, r.d.z };
        float t;
        if (!rtref_ray_hits_box(o, d, s->boxes + 6 * (size_t)i, &t) || t >= dist)
            continue;
        dist = t;
        prim = PRIM_BOX | i;
        material = s->box_material[i];
        hit_pos = ray_at(r.o, r.d, t);
    }
    /* hit_tests(scene.spheres), :63 */
    for (uint32_t i = 0; i < s->n_spheres; i++)
    {
        const float* sp = s->spheres + 4 * (size_t)i;
        float t;
        if (!ray_hits_sphere(&r, sp, &t) || t >= dist)
            continue;
        dist = t;
        prim = i;
        material = s->sphere_material[i];
        hit_pos = ray_at(r.o, r.d, t);
        hit_normal = normalize3(v3_sub(hit_pos, v3_make(sp[0], sp[1], sp[2]))); /* :56 */
    }
    if (prim_out) *prim_out = prim;
    if (depth_out) *depth_out = dist;

    float c[3];
    if (prim != PRIM_MISS)
    {
        /* :70-76: min(0.25 + lambert(n, direction(hit_pos, near), albedo).rgb * 0.75, 1); lambert (:14-20) =
         * l.dot(n) * vec3{albedo} * intensity(1.0f) */
        const v3 l = normalize3(v3_sub(near_p, hit_pos));
        const float k = dot3(l, hit_normal);
        const float* albedo = s->materials[material].albedo;
        for (int i = 0; i < 3; i++)
        {
            const float lam = (k * albedo[i]) * 1.0f;
            const float x = 0.25f + lam * 0.75f; /* rule R: separate multiply and add */
            c[i] = (x < 1.0f) ? x : 1.0f;        /* vec3::min */
        }
    }
    else
    {
        /* :65-66, :79-82: colour{238,245,255} / colour{208,228,255} are int-constructed, so every channel saturates to
         * 1.0f (colour.hpp:64-83); the lerp (S8 form) of the two whites over y/(h-1) is kept as written */
        const float a = (float)py / (float)(v->height - 1u);
        const float w = 1.0f - a;
        for (int i = 0; i < 3; i++)
            c[i] = FMA(1.0f, a, 1.0f * w);
    }
    /* colour -> uint32_t (colour.hpp:100-106), no gamma on this path */
    return (to_byte(c[0]) << 24) | (to_byte(c[1]) << 16) | (to_byte(c[2]) << 8) | to_byte(1.0f);
}

typedef struct {
    const rtref_scene* s;
    const rtref_view* v;
    uint32_t* rgba8;
    uint32_t* prim;
    float* depth;
    uint32_t row_step;
    int tid, nthreads;
} raster_job_t;

static void* raster_rows(void* arg)
{
    raster_job_t* j = (raster_job_t*)arg;
    const rtref_view* v = j->v;
    uint32_t row = 0;
    for (uint32_t y = v->tile_y0; y < v->tile_y1; y += j->row_step, row++)
    {
        if ((int)(row % (uint32_t)j->nthreads) != j->tid)
            continue;
        for (uint32_t x = v->tile_x0; x < v->tile_x1; x++)
        {
            const size_t pixel = (size_t)y * v->width + x;
            uint32_t prim;
            float depth;
            const uint32_t c = raster_pixel(j->s, v, x, y, &prim, &depth);
            if (j->rgba8) j->rgba8[pixel] = c;
            if (j->prim) j->prim[pixel] = prim;
            if (j->depth) j->depth[pixel] = depth;
        }
    }
    return NULL;
}

int rtref_rasterize(const rtref_scene* s, const rtref_view* v, uint32_t* rgba8, uint32_t* prim, float* depth,
                    int threads, uint32_t row_step)
{
    if (!s || !v || v->tile_x1 > v->width || v->tile_y1 > v->height)
        return -1;
    for (uint32_t i = 0; i < s->n_spheres; i++)
        if (s->sphere_material[i] >= s->n_materials) return -2;
    for (uint32_t i = 0; i < s->n_planes; i++)
        if (s->plane_material[i] >= s->n_materials) return -2;
    for (uint32_t i = 0; i < s->n_boxes; i++)
        if (s->box_material[i] >= s->n_materials) return -2;
    if (row_step == 0) row_step = 1;
    if (threads <= 0) threads = (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (threads < 1) threads = 1;
    if (threads > 1024) threads = 1024;
    raster_job_t* jobs = (raster_job_t*)calloc((size_t)threads, sizeof(raster_job_t));
    pthread_t* tids = (pthread_t*)calloc((size_t)threads, sizeof(pthread_t));
    for (int t = 0; t < threads; t++)
    {
        const raster_job_t j = { s, v, rgba8, prim, depth, row_step, t, threads };
        jobs[t] = j;
        if (t > 0) pthread_create(&tids[t], NULL, raster_rows, &jobs[t]);
    }
    raster_rows(&jobs[0]);
    for (int t = 1; t < threads; t++)
        pthread_join(tids[t], NULL);
    free(jobs);
    free(tids);
    return 0;
}

const char* rtref_build_flavour(void)
{
#ifdef RTREF_FAST
    return "fast";
#else
    return "strict";
#endif
}
