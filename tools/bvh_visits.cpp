// bvh_visits.cpp -- CPU proxy for the GPU traversal (kernels.cuh: trav_step): 4-wide node visits and leaf visits per path segment
// on a path-traced ray set (primary rays of a pinhole camera + the reference's lambert bounce), for the tree that
// rtcu_bvh4_build_host produces.  A tool for judging builder / collapse changes without a GPU: the traversal order, the
// (t, index) acceptance rule and the `tn <= best_t` culls are the kernel's; the conservative margins are left out, so the counts
// sit a few percent below rtcu_stats (C4: 7.4 here, 7.9 on the device; C3: 3.9 / 4.0).
//
//   g++ -O2 -Iinclude -o /tmp/bvh_visits tools/bvh_visits.cpp -Lrt_b200/lib -lrtcu -Wl,-rpath,$PWD/rt_b200/lib
//   python -c "from rt_b200 import synth; import numpy as np; np.ascontiguousarray(synth.grid_scene().spheres, np.float32).tofile('/tmp/c4.bin')"
//   /tmp/bvh_visits /tmp/c4.bin  0 6 8  0 -0.35 -1  3840 2160  8 10      # spheres  camera position  direction  W H  pixel step  depth
#include <rtcu.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <random>
#include <vector>
struct V { float x, y, z; };
static V operator+(V a, V b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
static V operator-(V a, V b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
static V operator*(V a, float s) { return { a.x * s, a.y * s, a.z * s }; }
static float dot(V a, V b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static V norm(V a) { return a * (1.0f / std::sqrt(dot(a, a))); }
static V cross(V a, V b) { return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }
struct Tree { std::vector<float> nodes, leaves; uint32_t nn, nl, depth; };
static uint64_t g_nodes, g_leaves, g_segs;
static bool closest(const Tree& T, V o, V d, float& bt, int& bi)
{
    const float inf = INFINITY;
    bt = inf; bi = 0x7fffffff;
    const float ix = 1 / d.x, iy = 1 / d.y, iz = 1 / d.z;
    uint32_t stack[128]; float stack_t[128]; int sp = 0;
    uint32_t node = 0;
    g_segs++;
    for (;;)
    {
        g_nodes++;
        const float* np = &T.nodes[32 * (size_t)node];
        uint32_t ref[4]; memcpy(ref, np + 24, 16);
        float tn[4]; bool hit[4];
        for (int c = 0; c < 4; c++)
        {
            const int pr = c / 2, sl = c % 2;
            const float* ax = np + 12 * pr;
            float n = -inf, f = inf;
            const float oo[3] = { o.x, o.y, o.z }, ii[3] = { ix, iy, iz };
            for (int k = 0; k < 3; k++)
            {
                const float cen = ax[4 * k + sl], h = ax[4 * k + 2 + sl];
                const float tc = (cen - oo[k]) * ii[k], hh = h * std::fabs(ii[k]);
                const float a = tc - hh, b = tc + hh;
                if (a == a) n = std::fmax(n, a);
                if (b == b) f = std::fmin(f, b);
            }
            tn[c] = n;
            hit[c] = f >= std::fmax(n, 0.0f) && n <= bt;
        }
        uint32_t next = 0xffffffffu; float next_t = 0;
        for (int c = 0; c < 4; c++)
        {
            if (!hit[c]) continue;
            if (ref[c] & 0x80000000u)
            {
                g_leaves++;
                const float* lp = &T.leaves[20 * (size_t)(ref[c] & 0x7fffffffu)];
                int idx[4]; memcpy(idx, lp + 16, 16);
                for (int k = 0; k < 4; k++)
                {
                    const float* a = lp + 8 * (k / 2); const float* b = a + 4; const int s = k % 2;
                    const V cc = { a[s], a[2 + s], b[s] }; const float r2 = b[2 + s];
                    const V e = cc - o; const float e2 = dot(e, e), aa = dot(e, d), disc = r2 - (e2 - aa * aa);
                    if (disc < 0) continue;
                    const float fq = std::sqrt(disc), t = e2 < r2 ? aa + fq : aa - fq;
                    if (t < 0.001f) continue;
                    if (t < bt || (t == bt && idx[k] < bi)) { bt = t; bi = idx[k]; }
                }
            }
            else if (next == 0xffffffffu) { next = ref[c]; next_t = tn[c]; }
            else
            {
                const bool sw = tn[c] < next_t;
                stack[sp] = sw ? next : ref[c]; stack_t[sp] = sw ? next_t : tn[c]; sp++;
                if (sw) { next = ref[c]; next_t = tn[c]; }
            }
        }
        if (next != 0xffffffffu && next_t <= bt) { node = next; continue; }
        bool found = false;
        while (sp > 0) { sp--; if (stack_t[sp] <= bt) { node = stack[sp]; found = true; break; } }
        if (!found) break;
    }
    return bi != 0x7fffffff;
}
int main(int argc, char** argv)
{
    const char* file = argv[1];
    const V cam_o = { (float)atof(argv[2]), (float)atof(argv[3]), (float)atof(argv[4]) }, cam_d = norm(V{ (float)atof(argv[5]), (float)atof(argv[6]), (float)atof(argv[7]) });
    const int W = atoi(argv[8]), H = atoi(argv[9]), step = atoi(argv[10]), depth = atoi(argv[11]);
    FILE* f = fopen(file, "rb"); fseek(f, 0, SEEK_END); const long bytes = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<float> sph(bytes / 4); if (fread(sph.data(), 1, bytes, f) != (size_t)bytes) return 1; fclose(f);
    const uint32_t n = (uint32_t)(sph.size() / 4);
    Tree T;
    if (rtcu_bvh4_build_host(sph.data(), n, nullptr, 0, nullptr, 0, &T.nn, &T.nl, &T.depth)) { puts(rtcu_last_error()); return 1; }
    T.nodes.resize(32 * (size_t)T.nn); T.leaves.resize(20 * (size_t)T.nl);
    rtcu_bvh4_build_host(sph.data(), n, T.nodes.data(), T.nn, T.leaves.data(), T.nl, &T.nn, &T.nl, &T.depth);
    const V right = norm(cross(cam_d, V{ 0, 1, 0 })), up = cross(right, cam_d);
    const float th = std::tan(0.5f * 0.78539816f), aspect = (float)W / H;
    std::mt19937 rng(1); std::uniform_real_distribution<float> U(0, 1);
    uint64_t prim_nodes = 0, prim_segs = 0;
    for (int y = step / 2; y < H; y += step)
        for (int x = step / 2; x < W; x += step)
        {
            V o = cam_o, d = norm(cam_d + right * ((2 * (x + 0.5f) / W - 1) * th * aspect) + up * ((1 - 2 * (y + 0.5f) / H) * th));
            for (int b = 0; b < depth; b++)
            {
                float t; int i;
                const uint64_t n0 = g_nodes;
                const bool h = closest(T, o, d, t, i);
                if (b == 0) { prim_nodes += g_nodes - n0; prim_segs++; }
                if (!h) break;
                const V p = o + d * t, c = { sph[4 * i], sph[4 * i + 1], sph[4 * i + 2] }, nrm = norm(p - c);
                V rv = { U(rng), U(rng), U(rng) };
                d = norm(nrm + norm(rv)); o = p;   // the reference's lambert scatter (positive-octant unit vector)
            }
        }
    printf("nodes4 %u leaves %u depth %u | segments %llu  node visits/seg %.3f  leaf visits/seg %.3f | primary: %.3f visits/seg\n", T.nn, T.nl, T.depth,
           (unsigned long long)g_segs, (double)g_nodes / g_segs, (double)g_leaves / g_segs, (double)prim_nodes / prim_segs);
}
