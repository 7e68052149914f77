/* bvh_replay.c -- the device's BVH traversal (rt_b200/csrc/kernels.cuh: trav_init, slab_pair, trav_step, bvh_leaf_pair_test),
 * restated in C over the arrays rtcu_bvh4_build_host returns, so that the claim "traversal returns the linear scan's result bit
 * for bit" can be checked on the CPU against the oracle's scan -- for the default tree and for every experimental builder variant,
 * on far more rays than a GPU test has time for.  What is replayed: the conservative margins (kappa, E = |c - o|_1 + H), the
 * `tf >= max(tn, 0) && tn <= best_t` cull, leaves tested at once, nearest inner child first, the others pushed in slot order, the
 * pop-time cull, the (t, index) acceptance rule, S4 with explicit fmaf (compile with -ffp-contract=off).  What is not: the packed
 * FFMA2 instruction selection (the box arithmetic only has to be conservative; the leaf arithmetic is S4's, rounding for rounding).
 * Test infrastructure only (tests/test_bvh_replay.py builds it into a temporary directory). */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define STACK 64

static inline float dot3(const float a[3], const float b[3]) { return fmaf(a[2], b[2], fmaf(a[1], b[1], a[0] * b[0])); }

/* returns 0 when the direction is too far from unit length (the kernel then scans), else 1 with (t, index) of the closest hit
 * (index -1 = miss); visits / leaf_visits are added to the counters */
/* any_t: the preview renderer's variant (raster.cuh, rasterizer.cpp:41-52) -- the whole line counts, no minimum distance, and the
 * search starts from (start_t, start_i) = (distance of the closest plane / box so far, -1) */
int bvh_replay_ray(const float* nodes, const float* leaves, const float o[3], const float d[3], int any_t, float start_t, int32_t start_i,
                   float* t_out, int32_t* i_out, uint64_t* visits, uint64_t* leaf_visits, uint32_t* max_sp)
{
    const float eps_d = fabsf(dot3(d, d) - 1.0f);
    if (!(eps_d <= 1e-3f)) return 0;
    if (!(fabsf(o[0]) <= 0x1p62f && fabsf(o[1]) <= 0x1p62f && fabsf(o[2]) <= 0x1p62f)) return 0; /* S4 could overflow: the kernel scans */
    const float kappa = 1.01f * sqrtf(16.0f * 5.9604645e-8f + 2.0f * eps_d);
    const float inv[3] = { 1.0f / d[0], 1.0f / d[1], 1.0f / d[2] };
    float best_t = start_t;
    int32_t best_i = start_i;
    uint32_t stack_ref[STACK];
    float stack_t[STACK];
    uint32_t sp = 0, node = 0;
    for (;;)
    {
        const float* np = nodes + 32u * (size_t)node;
        uint32_t ref[4];
        memcpy(ref, np + 24, sizeof ref);
        const float* H = np + 28;
        float tn[4];
        int hit[4];
        (*visits)++;
        for (int c = 0; c < 4; c++)
        {
            const int pair = c >> 1, slot = c & 1;
            const float* ax = np + 12 * pair; /* x: {c0,c1,h0,h1}, y, z */
            float dc[3], tc[3];
            for (int k = 0; k < 3; k++) dc[k] = ax[4 * k + slot] - o[k];
            const float e3 = fabsf(dc[0]) + fabsf(dc[1]) + fabsf(dc[2]);
#ifdef BVH_REPLAY_NO_MARGIN /* what the rays of the test are for: without the margins the traversal misses grazing hits */
            const float m = 0.0f * (e3 + H[c]) * kappa;
#else
            const float m = (e3 + H[c]) * kappa;
#endif
            float near = -INFINITY, far = INFINITY;
            for (int k = 0; k < 3; k++)
            {
                const float h = ax[4 * k + 2 + slot] + m, a = fabsf(inv[k]);
                tc[k] = dc[k] * inv[k];
                near = fmaxf(near, fmaf(-h, a, tc[k])); /* fmaxf / fminf drop a NaN operand, as the device's do */
                far = fminf(far, fmaf(h, a, tc[k]));
            }
            tn[c] = near;
            hit[c] = far >= (any_t ? near : fmaxf(near, 0.0f)) && near <= best_t;
        }
        uint32_t next = 0xffffffffu;
        float next_t = 0.0f;
        for (int c = 0; c < 4; c++)
        {
            if (!hit[c]) continue;
            if (ref[c] & 0x80000000u)
            {
                const float* lp = leaves + 20u * (size_t)(ref[c] & 0x7fffffffu);
                int32_t idx[4];
                memcpy(idx, lp + 16, sizeof idx);
                (*leaf_visits)++;
                const int n_slots = idx[2] == 0x7fffffff ? 2 : 4; /* kernels.cuh leaf_second_pair: padding is last, a padded third slot = no second pair */
                for (int k = 0; k < n_slots; k++)
                {
                    const float* A = lp + 8 * (k >> 1);
                    const float* B = A + 4;
                    const int s = k & 1;
                    const float e[3] = { A[s] - o[0], A[2 + s] - o[1], B[s] - o[2] };
                    const float r2 = B[2 + s];
                    const float e2 = dot3(e, e), a = dot3(e, d);
                    const float disc = r2 - fmaf(-a, a, e2);
                    if (disc < 0.0f) continue;
                    const float f = sqrtf(disc);
                    const float t = (e2 < r2) ? a + f : a - f;
                    if ((any_t || !(t < 0.001f)) && (t < best_t || (t == best_t && idx[k] < best_i)))
                    {
                        best_t = t;
                        best_i = idx[k];
                    }
                }
            }
            else if (next == 0xffffffffu)
            {
                next = ref[c];
                next_t = tn[c];
            }
            else
            {
                const int swap = tn[c] < next_t;
                if (sp >= STACK) return -1;
                stack_ref[sp] = swap ? next : ref[c];
                stack_t[sp] = swap ? next_t : tn[c];
                sp++;
                if (sp > *max_sp) *max_sp = sp;
                if (swap) { next = ref[c]; next_t = tn[c]; }
            }
        }
        if (next != 0xffffffffu && next_t <= best_t) { node = next; continue; }
        int found = 0;
        while (sp > 0)
        {
            sp--;
            if (stack_t[sp] <= best_t) { node = stack_ref[sp]; found = 1; break; }
        }
        if (!found) break;
    }
    *t_out = best_t;
    *i_out = best_i == 0x7fffffff ? -1 : best_i;
    return 1;
}

/* ---- traversal 2 (kernels.cuh: closest_sphere_bvh2): leaf children are not tested inside the node visit but ordered with the
 * inner ones -- the nearest hit child of either kind is processed next, the others are pushed in slot order with their entry
 * distance, and the pop-time cull applies to leaves as well.  Same margins, same cull, same (t, index) rule. */
static void slab4(const float* np, const float o[3], const float inv[3], float kappa, float best_t, int any_t, float tn[4], int hit[4])
{
    const float* H = np + 28;
    for (int c = 0; c < 4; c++)
    {
        const int pair = c >> 1, slot = c & 1;
        const float* ax = np + 12 * pair;
        float dc[3];
        for (int k = 0; k < 3; k++) dc[k] = ax[4 * k + slot] - o[k];
        const float e3 = fabsf(dc[0]) + fabsf(dc[1]) + fabsf(dc[2]);
#ifdef BVH_REPLAY_NO_MARGIN
        const float m = 0.0f * (e3 + H[c]) * kappa;
#else
        const float m = (e3 + H[c]) * kappa;
#endif
        float near = -INFINITY, far = INFINITY;
        for (int k = 0; k < 3; k++)
        {
            const float h = ax[4 * k + 2 + slot] + m, a = fabsf(inv[k]), tc = dc[k] * inv[k];
            near = fmaxf(near, fmaf(-h, a, tc));
            far = fminf(far, fmaf(h, a, tc));
        }
        tn[c] = near;
        hit[c] = far >= (any_t ? near : fmaxf(near, 0.0f)) && near <= best_t;
    }
}

int bvh_replay_ray2(const float* nodes, const float* leaves, const float o[3], const float d[3], int any_t, float start_t, int32_t start_i,
                    float* t_out, int32_t* i_out, uint64_t* visits, uint64_t* leaf_visits, uint32_t* max_sp)
{
    const float eps_d = fabsf(dot3(d, d) - 1.0f);
    if (!(eps_d <= 1e-3f)) return 0;
    if (!(fabsf(o[0]) <= 0x1p62f && fabsf(o[1]) <= 0x1p62f && fabsf(o[2]) <= 0x1p62f)) return 0;
    const float kappa = 1.01f * sqrtf(16.0f * 5.9604645e-8f + 2.0f * eps_d);
    const float inv[3] = { 1.0f / d[0], 1.0f / d[1], 1.0f / d[2] };
    float best_t = start_t;
    int32_t best_i = start_i;
    uint32_t stack_ref[STACK + 1];
    float stack_t[STACK + 1];
    uint32_t sp = 0, cur = 0;
    const uint32_t NONE = 0xffffffffu;
    for (;;)
    {
        if ((int32_t)cur >= 0)
        {
            const float* np = nodes + 32u * (size_t)cur;
            uint32_t ref[4];
            memcpy(ref, np + 24, sizeof ref);
            float tn[4];
            int hit[4];
            (*visits)++;
            slab4(np, o, inv, kappa, best_t, any_t, tn, hit);
            float nt = INFINITY;
            uint32_t nref = NONE;
            int nslot = -1;
            for (int c = 0; c < 4; c++)
                if (hit[c] && (nslot < 0 || tn[c] < nt)) { nt = tn[c]; nref = ref[c]; nslot = c; }
#ifdef BVH_REPLAY_SORTED_PUSH /* experiment: push the other children farthest first, so that they pop nearest first */
            {
                int ord[4] = { 0, 1, 2, 3 };
                for (int i = 0; i < 4; i++)
                    for (int j = i + 1; j < 4; j++)
                        if (tn[ord[j]] > tn[ord[i]]) { const int x = ord[i]; ord[i] = ord[j]; ord[j] = x; }
                for (int i = 0; i < 4; i++)
                {
                    const int c = ord[i];
                    if (sp > STACK) return -1;
                    stack_t[sp] = tn[c];
                    stack_ref[sp] = ref[c];
                    sp += (hit[c] && c != nslot) ? 1 : 0;
                    if (sp > *max_sp) *max_sp = sp;
                }
            }
#else
            for (int c = 0; c < 4; c++)
            {
                if (sp > STACK) return -1;
                stack_t[sp] = tn[c];
                stack_ref[sp] = ref[c];
                sp += (hit[c] && c != nslot) ? 1 : 0;
                if (sp > *max_sp) *max_sp = sp;
            }
#endif
            cur = nref;
        }
        if ((int32_t)cur < -1)
        {
            const float* lp = leaves + 20u * (size_t)(cur & 0x7fffffffu);
            int32_t idx[4];
            memcpy(idx, lp + 16, sizeof idx);
            (*leaf_visits)++;
            for (int k = 0; k < 4; k++)
            {
                const float* A = lp + 8 * (k >> 1);
                const float* B = A + 4;
                const int s = k & 1;
                const float e[3] = { A[s] - o[0], A[2 + s] - o[1], B[s] - o[2] };
                const float r2 = B[2 + s];
                const float e2 = dot3(e, e), a = dot3(e, d);
                const float disc = r2 - fmaf(-a, a, e2);
                if (disc < 0.0f) continue;
                const float f = sqrtf(disc);
                const float t = (e2 < r2) ? a + f : a - f;
                if ((any_t || !(t < 0.001f)) && (t < best_t || (t == best_t && idx[k] < best_i)))
                {
                    best_t = t;
                    best_i = idx[k];
                }
            }
            cur = NONE;
        }
        if (cur == NONE)
        {
            int found = 0;
            while (sp > 0)
            {
                sp--;
                if (stack_t[sp] <= best_t) { cur = stack_ref[sp]; found = 1; break; }
            }
            if (!found) break;
        }
    }
    *t_out = best_t;
    *i_out = best_i == 0x7fffffff ? -1 : best_i;
    return 1;
}

/* ---- patch beams (kernels.cuh: k_beam_lists, beam_collect, beam_closest_sphere): ONE conservative walk per patch of pixels with
 * the patch's centre ray, margins widened by the patch's footprint and no closest-hit cull, collects the leaves any primary ray
 * of the patch could hit; a sample's primary ray then takes the (t, index) minimum over those leaves only.  rays: 5 x (o, d) =
 * the centre ray and the rays through the patch's four corners.  Returns the list length, or -1 when the patch keeps the
 * traversal (too many leaves or node visits, centre ray not unit length). */
#define BEAM_MAX 32
#define BEAM_MAX_VISITS 192
#define BEAM_EPS_D (16.0f * 5.9604645e-8f)
int bvh_replay_beam_collect(const float* nodes, const float* rays, float* list_tn, uint32_t* list_leaf)
{
    const float* o = rays;
    const float* d = rays + 3;
    float dd = 0.0f, oo = 0.0f;
    for (int k = 1; k < 5; k++)
    {
        const float* ok = rays + 6 * k;
        const float* dk = ok + 3;
        dd = fmaxf(dd, fabsf(dk[0] - d[0]) + fabsf(dk[1] - d[1]) + fabsf(dk[2] - d[2]));
        oo = fmaxf(oo, fabsf(ok[0] - o[0]) + fabsf(ok[1] - o[1]) + fabsf(ok[2] - o[2]));
    }
    const float u16 = 16.0f * 5.9604645e-8f;
    const float sigma = 1.01f * dd + u16, rho = 1.01f * oo + u16 * (fabsf(o[0]) + fabsf(o[1]) + fabsf(o[2]) + 1.0f);
    const float eps_d = fabsf(dot3(d, d) - 1.0f);
    if (!(eps_d <= 1e-3f) || !(eps_d <= BEAM_EPS_D)) return -1;
    if (!(fabsf(o[0]) <= 0x1p62f && fabsf(o[1]) <= 0x1p62f && fabsf(o[2]) <= 0x1p62f)) return -1;
#ifdef BVH_REPLAY_NO_BEAM_MARGIN /* what the footprint terms are for: the centre ray's own margins lose hits of the pixel's other rays */
    const float kappa = 1.01f * (1.01f * sqrtf(16.0f * 5.9604645e-8f + 2.0f * BEAM_EPS_D)), madd = 0.0f;
    (void)sigma; (void)rho;
#else
    const float kappa = 1.01f * (1.01f * sqrtf(16.0f * 5.9604645e-8f + 2.0f * BEAM_EPS_D) + 2.01f * sigma), madd = 2.02f * rho;
#endif
    const float inv[3] = { 1.0f / d[0], 1.0f / d[1], 1.0f / d[2] };
    uint32_t stack[STACK];
    int sp = 0, n = 0, visits = 0;
    uint32_t node = 0;
    for (;;)
    {
        if (++visits > BEAM_MAX_VISITS) return -1;
        const float* np = nodes + 32u * (size_t)node;
        uint32_t ref[4];
        memcpy(ref, np + 24, sizeof ref);
        const float* H = np + 28;
        for (int c = 0; c < 4; c++)
        {
            const int pair = c >> 1, slot = c & 1;
            const float* ax = np + 12 * pair;
            float dc[3];
            for (int k = 0; k < 3; k++) dc[k] = ax[4 * k + slot] - o[k];
            const float e3 = fabsf(dc[0]) + fabsf(dc[1]) + fabsf(dc[2]);
            const float m = fmaf(e3 + H[c], kappa, madd);
            float near = -INFINITY, far = INFINITY;
            for (int k = 0; k < 3; k++)
            {
                const float h = ax[4 * k + 2 + slot] + m, a = fabsf(inv[k]), tc = dc[k] * inv[k];
                near = fmaxf(near, fmaf(-h, a, tc));
                far = fminf(far, fmaf(h, a, tc));
            }
            if (!(far >= fmaxf(near, 0.0f))) continue; /* no closest-hit cull: best_t stays +inf */
            if (ref[c] & 0x80000000u)
            {
                if (n == BEAM_MAX) return -1;
                int k = n++;
                for (; k > 0 && list_tn[k - 1] > near; k--)
                {
                    list_tn[k] = list_tn[k - 1];
                    list_leaf[k] = list_leaf[k - 1];
                }
                list_tn[k] = near;
                list_leaf[k] = ref[c] & 0x7fffffffu;
            }
            else
            {
                if (sp >= STACK) return -1;
                stack[sp++] = ref[c];
            }
        }
        if (sp == 0) break;
        node = stack[--sp];
    }
    return n;
}

/* closest hit of n rays of the pixel over the listed leaves; usable[i] = 0 where the kernel would traverse instead (the ray is not
 * unit length to within BEAM_EPS_D) */
void bvh_replay_beam_closest(const float* leaves, const float* list_tn, const uint32_t* list_leaf, int n_list, const float* o, const float* d,
                             uint32_t n, uint8_t* hit, uint32_t* prim, float* t, uint8_t* usable)
{
    for (uint32_t i = 0; i < n; i++)
    {
        const float* oi = o + 3 * (size_t)i;
        const float* di = d + 3 * (size_t)i;
        usable[i] = fabsf(dot3(di, di) - 1.0f) <= BEAM_EPS_D;
        float best_t = INFINITY;
        int32_t best_i = 0x7fffffff;
        for (int k = 0; k < n_list; k++)
        {
            if (list_tn[k] > best_t) break;
            const float* lp = leaves + 20u * (size_t)list_leaf[k];
            int32_t idx[4];
            memcpy(idx, lp + 16, sizeof idx);
            const int n_slots = idx[2] == 0x7fffffff ? 2 : 4; /* as the device's list scan */
            for (int q = 0; q < n_slots; q++)
            {
                const float* A = lp + 8 * (q >> 1);
                const float* B = A + 4;
                const int s = q & 1;
                const float e[3] = { A[s] - oi[0], A[2 + s] - oi[1], B[s] - oi[2] };
                const float r2 = B[2 + s];
                const float e2 = dot3(e, e), a = dot3(e, di);
                const float disc = r2 - fmaf(-a, a, e2);
                if (disc < 0.0f) continue;
                const float f = sqrtf(disc);
                const float tt = (e2 < r2) ? a + f : a - f;
                if (!(tt < 0.001f) && (tt < best_t || (tt == best_t && idx[q] < best_i)))
                {
                    best_t = tt;
                    best_i = idx[q];
                }
            }
        }
        hit[i] = best_i != 0x7fffffff;
        prim[i] = hit[i] ? (uint32_t)best_i : 0xffffffffu;
        t[i] = hit[i] ? best_t : -1.0f;
    }
}

/* n rays; hit[i] = 1 / 0, prim, t as the scan reports them (t = -1 on a miss), skipped[i] = 1 where the kernel would scan instead.
 * Returns the deepest stack use, or -1 on overflow. */
int bvh_replay_batch(const float* nodes, const float* leaves, const float* o, const float* d, uint32_t n, uint8_t* hit, uint32_t* prim, float* t,
                     uint8_t* skipped, uint64_t counters[2], int any_t)
{
    uint32_t max_sp = 0;
    for (uint32_t i = 0; i < n; i++)
    {
        float tt = 0.0f;
        int32_t ii = -1;
#ifdef BVH_REPLAY_TRAV2 /* the path tracer's default traversal; the preview renderer (any_t) uses the first one */
        const int rc = bvh_replay_ray2(nodes, leaves, o + 3 * (size_t)i, d + 3 * (size_t)i, any_t, INFINITY, any_t ? -1 : 0x7fffffff, &tt, &ii,
                                       &counters[0], &counters[1], &max_sp);
#else
        const int rc = bvh_replay_ray(nodes, leaves, o + 3 * (size_t)i, d + 3 * (size_t)i, any_t, INFINITY, any_t ? -1 : 0x7fffffff, &tt, &ii,
                                      &counters[0], &counters[1], &max_sp);
#endif
        if (rc < 0) return -1;
        skipped[i] = rc == 0;
        hit[i] = rc == 1 && ii >= 0;
        prim[i] = hit[i] ? (uint32_t)ii : 0xffffffffu;
        t[i] = hit[i] ? tt : -1.0f;
    }
    return (int)max_sp;
}
