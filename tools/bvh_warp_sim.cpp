// bvh_warp_sim.cpp -- SIMT model of the BVH render kernels: how many lanes of a warp are busy while it traverses, under different
// ways of handing (pixel, sample) work to lanes.  A warp's traversal of one batch of segments costs max-over-lanes node visits
// (lanes that finish early idle until the slowest is done: `while (!trav_step())`); lanes busy = sum of visits / sum of maxima.
// Paths follow the reference's scatter rules (lambert / metal / dielectric with Schlick) on the real scenes, with a throw-away RNG:
// the statistics are the point, not the image.  The model reproduces the device's lane statistics (22-24 of 32 lanes busy here,
// 19-23 in the ncu captures of the lanes-share-a-pixel kernel) and answers what a GPU round would otherwise spend calls on:
//   * regrouping the rays of a 128-path pool into warps by primary / secondary, direction octant or bounce index before each
//     batch saves 3 % of the warp steps; sorting by the (unknowable) visit count itself only 8 %;
//   * compacting lanes whose pixel has run out of samples saves 9 %;
//   * the bound for any regrouping, with perfect knowledge over the whole frame, is 27 % (C4) / 31 % (C3).
//
//   g++ -O2 -Iinclude -o /tmp/bvh_warp_sim tools/bvh_warp_sim.cpp -Lrt_b200/lib -lrtcu -Wl,-rpath,$PWD/rt_b200/lib
//   inputs <base>.sph / .smat / .mat: float32 spheres (n x 4), uint32 material index per sphere, float32 (type, roughness, reflectivity) per material
//   /tmp/bvh_warp_sim /tmp/c4  0 6 8  0 -0.35 -1  3840 2160  64 10  60     # base  camera position  direction  W H  spp depth  8x4 patches sampled
#include <rtcu.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <random>
#include <string>
#include <vector>
struct V { float x, y, z; };
static V operator+(V a, V b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
static V operator-(V a, V b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
static V operator*(V a, float s) { return { a.x * s, a.y * s, a.z * s }; }
static float dot(V a, V b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static V norm(V a) { return a * (1.0f / std::sqrt(dot(a, a))); }
static V cross(V a, V b) { return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }
struct Tree { std::vector<float> nodes, leaves; uint32_t nn, nl, depth; };
static Tree T;
static std::vector<float> sph, mat; static std::vector<uint32_t> smat;
// returns node visits; bt/bi = closest hit
static int closest(V o, V d, float& bt, int& bi, int& leaf_visits)
{
    const float inf = INFINITY; bt = inf; bi = 0x7fffffff; leaf_visits = 0;
    const float ii[3] = { 1 / d.x, 1 / d.y, 1 / d.z }, oo[3] = { o.x, o.y, o.z };
    uint32_t stack[128]; float stack_t[128]; int sp = 0; uint32_t node = 0; int visits = 0;
    for (;;)
    {
        visits++;
        const float* np = &T.nodes[32 * (size_t)node];
        uint32_t ref[4]; memcpy(ref, np + 24, 16);
        float tn[4]; bool hit[4];
        for (int c = 0; c < 4; c++)
        {
            const int pr = c / 2, sl = c % 2; const float* ax = np + 12 * pr; float n = -inf, f = inf;
            for (int k = 0; k < 3; k++)
            {
                const float tc = (ax[4 * k + sl] - oo[k]) * ii[k], hh = ax[4 * k + 2 + sl] * std::fabs(ii[k]);
                const float a = tc - hh, b = tc + hh; if (a == a) n = std::fmax(n, a); if (b == b) f = std::fmin(f, b);
            }
            tn[c] = n; hit[c] = f >= std::fmax(n, 0.0f) && n <= bt;
        }
        uint32_t next = 0xffffffffu; float next_t = 0;
        for (int c = 0; c < 4; c++)
        {
            if (!hit[c]) continue;
            if (ref[c] & 0x80000000u)
            {
                leaf_visits++;
                const float* lp = &T.leaves[20 * (size_t)(ref[c] & 0x7fffffffu)]; int idx[4]; memcpy(idx, lp + 16, 16);
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
            else { const bool sw = tn[c] < next_t; stack[sp] = sw ? next : ref[c]; stack_t[sp] = sw ? next_t : tn[c]; sp++; if (sw) { next = ref[c]; next_t = tn[c]; } }
        }
        if (next != 0xffffffffu && next_t <= bt) { node = next; continue; }
        bool found = false;
        while (sp > 0) { sp--; if (stack_t[sp] <= bt) { node = stack[sp]; found = true; break; } }
        if (!found) break;
    }
    return visits;
}
// one path state
struct Path { V o, d; int bounce; bool alive; uint32_t rng; int px, py, sample; };
static float u01(uint32_t& s) { s = s * 1664525u + 1013904223u; return (s >> 8) * (1.0f / 16777216.0f); }
static V cam_o, cam_f, cam_r, cam_u; static int W, H, DEPTH; static float th, aspect;
static void start(Path& p, int px, int py, int sample)
{
    p.px = px; p.py = py; p.sample = sample; p.bounce = 0; p.alive = true;
    p.rng = (uint32_t)(px * 9781 + py * 6271 + sample * 26699) * 2654435761u + 12345u;
    const float jx = sample ? u01(p.rng) : 0.5f, jy = sample ? u01(p.rng) : 0.5f;
    p.o = cam_o; p.d = norm(cam_f + cam_r * ((2 * (px + jx) / W - 1) * th * aspect) + cam_u * ((1 - 2 * (py + jy) / H) * th));
}
// advance after a traversal: returns false when the path ended
static bool scatter(Path& p, float t, int i)
{
    if (i == 0x7fffffff) return false;
    if (++p.bounce >= DEPTH) return false;
    const V hp = p.o + p.d * t, c = { sph[4 * i], sph[4 * i + 1], sph[4 * i + 2] }, n = norm(hp - c);
    const float* m = &mat[3 * smat[i]]; const int type = (int)m[0];
    V rv = norm(V{ u01(p.rng), u01(p.rng), u01(p.rng) });
    if (type == 1) { const V v = norm(p.d); V s = (v - n * (2 * dot(v, n))) + rv * m[1]; if (dot(s, n) <= 0) return false; p.d = norm(s); }
    else if (type >= 2 && type <= 6)
    {
        const float ior = m[2]; const V refl = p.d - n * (2 * dot(p.d, n)); V on; float eta, cs; const float dn = dot(p.d, n), len = std::sqrt(dot(p.d, p.d));
        if (dn > 0) { on = n * -1.0f; eta = ior; cs = ior * dn / len; } else { on = n; eta = 1 / ior; cs = -dn / len; }
        const float ci = -dot(p.d, on), s2 = eta * eta * (1 - ci * ci); float prob = 1; V refr = refl;
        if (s2 <= 1) { const float ct = std::sqrt(1 - s2); refr = p.d * eta + on * (eta * ci - ct); float r0 = (1 - ior) / (1 + ior); r0 *= r0; prob = r0 + (1 - r0) * std::pow(1 - cs, 5.0f); }
        p.d = norm(u01(p.rng) < prob ? refl : refr);
    }
    else p.d = norm(n + rv);
    p.o = hp;
    return true;
}
struct Acc { double visits = 0, maxima = 0, batches = 0, segs = 0; };
// a warp of `lanes` lanes; each lane owns a queue position in `work` (pixel, sample) lists according to the policy
int main(int argc, char** argv)
{
    const std::string base = argv[1];
    cam_o = { (float)atof(argv[2]), (float)atof(argv[3]), (float)atof(argv[4]) }; cam_f = norm(V{ (float)atof(argv[5]), (float)atof(argv[6]), (float)atof(argv[7]) });
    W = atoi(argv[8]); H = atoi(argv[9]); const int SPP = atoi(argv[10]); DEPTH = atoi(argv[11]); const int tiles = atoi(argv[12]);
    auto rd = [&](const std::string& f, auto& v) { FILE* fp = fopen(f.c_str(), "rb"); fseek(fp, 0, SEEK_END); long b = ftell(fp); fseek(fp, 0, SEEK_SET); v.resize(b / sizeof(v[0])); if (fread(v.data(), 1, b, fp) != (size_t)b) exit(1); fclose(fp); };
    rd(base + ".sph", sph); rd(base + ".smat", smat); rd(base + ".mat", mat);
    const uint32_t n = sph.size() / 4;
    rtcu_bvh4_build_host(sph.data(), n, nullptr, 0, nullptr, 0, &T.nn, &T.nl, &T.depth);
    T.nodes.resize(32 * (size_t)T.nn); T.leaves.resize(20 * (size_t)T.nl);
    rtcu_bvh4_build_host(sph.data(), n, T.nodes.data(), T.nn, T.leaves.data(), T.nl, &T.nn, &T.nl, &T.depth);
    cam_r = norm(cross(cam_f, V{ 0, 1, 0 })); cam_u = cross(cam_r, cam_f); th = std::tan(0.5f * 0.78539816f); aspect = (float)W / H;
    std::mt19937 pick(7);
    // sample `tiles` 8x4 pixel patches spread over the image
    std::vector<std::pair<int, int>> patches;
    for (int k = 0; k < tiles; k++) patches.push_back({ (int)(pick() % (W / 8)) * 8, (int)(pick() % (H / 4)) * 4 });

    // policy A: thread per pixel -- lane l of the warp owns pixel l of the 8x4 patch and runs its samples in order (regenerating)
    // policy G: G lanes share a pixel (32/G pixels per warp, taken in patch order); a lane claims the pixel's next sample when its path ends
    // policy S: ideal regrouping bound -- all segments of the patch set, sorted by visit count, packed 32 at a time
    auto run_policy = [&](int G) {
        Acc acc; std::vector<int> all_visits;
        for (auto [x0, y0] : patches)
        {
            const int groups = 32 / G;                 // pixels processed concurrently by the warp
            for (int gp = 0; gp < 32; gp += groups)    // successive pixel sets of the patch
            {
                std::vector<Path> lane(32); std::vector<int> next_sample(groups, 0);
                auto claim = [&](int l) {
                    const int g = l / G, pix = gp + g, px = x0 + pix % 8, py = y0 + pix / 8;
                    if (next_sample[g] >= SPP) { lane[l].alive = false; return; }
                    start(lane[l], px, py, next_sample[g]++);
                };
                for (int l = 0; l < 32; l++) claim(l);
                for (;;)
                {
                    int mx = 0, sum = 0, live = 0;
                    for (int l = 0; l < 32; l++)
                    {
                        if (!lane[l].alive) continue;
                        float t; int i, lv; const int v = closest(lane[l].o, lane[l].d, t, i, lv) + lv; // leaf visits cost about a node visit each
                        all_visits.push_back(v);
                        sum += v; mx = std::max(mx, v); live++;
                        if (!scatter(lane[l], t, i)) claim(l);
                    }
                    if (!live) break;
                    acc.visits += sum; acc.maxima += mx; acc.batches++; acc.segs += live;
                }
            }
        }
        std::sort(all_visits.begin(), all_visits.end());
        double ideal_max = 0; for (size_t k = 31; k < all_visits.size(); k += 32) ideal_max += all_visits[k]; if (all_visits.size() % 32) ideal_max += all_visits.back();
        printf("G=%2d: segments %.0f  steps/seg %.2f  lanes busy in traversal %.1f/32  warp-steps per segment %.3f   (ideal regrouping: %.3f)\n", G, acc.segs, acc.visits / acc.segs,
               32.0 * acc.visits / (32.0 * acc.maxima), acc.maxima / acc.segs, ideal_max / all_visits.size());
    };
    for (int G : { 1, 8, 16, 32 }) run_policy(G);
    // policy R: a CTA-wide pool of 128 paths (G = 16: eight pixels in flight); before every batch the paths are regrouped into
    // warps by a key known before the traversal: KEY 0 = none (lane order), 1 = primary / secondary, 2 = + direction octant,
    // 3 = + bounce index, 9 = oracle (the visit count itself: what perfect knowledge inside the pool would give)
    for (int KEY : { 0, 1, 2, 3, 9 })
    {
        Acc acc; const int P = 128, G = 16;
        for (size_t pi = 0; pi + 3 < patches.size(); pi += 4)
        {
            // four patches -> 128 pixels; the pool works on 8 pixels at a time
            std::vector<std::pair<int,int>> pixels;
            for (int q = 0; q < 4; q++) for (int k = 0; k < 32; k++) pixels.push_back({ patches[pi + q].first + k % 8, patches[pi + q].second + k / 8 });
            for (size_t base = 0; base < pixels.size(); base += P / G)
            {
                std::vector<Path> lane(P); std::vector<int> next_sample(P / G, 0);
                auto claim = [&](int l) { const int g = l / G; if (next_sample[g] >= SPP) { lane[l].alive = false; return; } start(lane[l], pixels[base + g].first, pixels[base + g].second, next_sample[g]++); };
                for (int l = 0; l < P; l++) claim(l);
                for (;;)
                {
                    struct Item { int key, lane, v; }; std::vector<Item> items;
                    for (int l = 0; l < P; l++)
                    {
                        if (!lane[l].alive) continue;
                        float t; int i, lv; const int v = closest(lane[l].o, lane[l].d, t, i, lv) + lv;
                        const V d = lane[l].d; const int oct = (d.x < 0) | ((d.y < 0) << 1) | ((d.z < 0) << 2);
                        int key = 0;
                        if (KEY == 1) key = lane[l].bounce > 0; else if (KEY == 2) key = (lane[l].bounce > 0) * 8 + (lane[l].bounce > 0 ? oct : 0);
                        else if (KEY == 3) key = std::min(lane[l].bounce, 7) * 8 + (lane[l].bounce > 0 ? oct : 0); else if (KEY == 9) key = v;
                        items.push_back({ key, l, v });
                        if (!scatter(lane[l], t, i)) claim(l);
                    }
                    if (items.empty()) break;
                    if (KEY) std::stable_sort(items.begin(), items.end(), [](const Item& a, const Item& b) { return a.key < b.key; });
                    for (size_t k = 0; k < items.size(); k += 32)
                    {
                        int mx = 0; for (size_t j = k; j < std::min(items.size(), k + 32); j++) { mx = std::max(mx, items[j].v); acc.visits += items[j].v; }
                        acc.maxima += mx;
                    }
                    acc.segs += items.size();
                }
            }
        }
        printf("pool of 128, key %d: lanes busy %.1f/32  warp-steps per segment %.3f\n", KEY, acc.visits / acc.maxima, acc.maxima / acc.segs);
    }
}
