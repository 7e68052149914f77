// pool_stress.cpp -- the builder's worker pool (rt_b200/csrc/bvh.h) under the schedule it rarely sees: compiled with a spin of a few
// iterations, so the workers fall asleep on the condition variable between almost all sections and every hand-over goes through
// the sleep / wake protocol.  Builds many trees of varying size at varying thread counts and compares each with the one-thread
// tree; a lost wake-up would hang (the test has a timeout), a race would show as a different tree.
#include "../../rt_b200/csrc/bvh.h"
#include <cstdio>
#include <cstring>
#include <random>

static unsigned long long tree_hash(const rtcu_bvh::Result& r)
{
    unsigned long long h = 1469598103934665603ull;
    auto mix = [&](const void* p, size_t n) { const unsigned char* b = static_cast<const unsigned char*>(p); for (size_t i = 0; i < n; i++) h = (h ^ b[i]) * 1099511628211ull; };
    mix(r.nodes.data(), r.nodes.size() * sizeof(rtcu_bvh::Node));
    mix(r.order.data(), r.order.size() * sizeof(uint32_t));
    return h;
}

int main(int argc, char** argv)
{
    const int rounds = argc > 1 ? atoi(argv[1]) : 40;
    std::mt19937 g(99);
    std::uniform_real_distribution<float> u(0, 1);
    unsigned long long builds = 0;
    for (int round = 0; round < rounds; round++)
    {
        const uint32_t n = 8192 + g() % 30000;
        std::vector<float> s(4 * (size_t)n);
        for (uint32_t i = 0; i < n; i++) { s[4 * i] = 100 * u(g); s[4 * i + 1] = 5 * u(g); s[4 * i + 2] = 100 * u(g); s[4 * i + 3] = 0.05f + 0.3f * u(g); }
        unsigned long long want = 0;
        {
            rtcu_bvh::Pool one(0);
            want = tree_hash(rtcu_bvh::build(s.data(), n, one));
        }
        for (unsigned workers : { 1u, 2u, 5u, 11u })
        {
            rtcu_bvh::Pool pool(workers);
            for (int rep = 0; rep < 2; rep++, builds++) // the second build reuses workers that have been asleep
                if (tree_hash(rtcu_bvh::build(s.data(), n, pool)) != want) { printf("MISMATCH round %d workers %u\n", round, workers); return 1; }
        }
    }
    printf("ok %llu builds\n", builds);
    return 0;
}
