/* div_by_const_check.c -- spec.cuh's division of a pixel coordinate by the image side, q = a * rb, RN(a / b) = fma(rb, fma(-b, q, a), q)
 * with rb = RN(1 / b) rounded once on the host, checked on the CPU against IEEE division for EVERY image side up to `max_side` (a window
 * can be resized to any size; the device self-test covers all floats but only a handful of sides): all pixel centres x + 0.5, the pixel
 * corners x and x + 1, and `jitters` coordinates x + u with u = k * 2^-24 drawn from a 32-bit LCG (the RNG's float mapping, S9).
 * Returns the number of coordinates whose quotient differs; *checked = coordinates tried.  Build with -ffp-contract=off -fno-fast-math. */
#include <math.h>
#include <stdint.h>

static inline float div_by_const(float a, float b, float rb)
{
    const float q = a * rb;
    return fmaf(rb, fmaf(-b, q, a), q);
}

uint64_t div_by_const_check(uint32_t first_side, uint32_t max_side, uint32_t jitters, uint64_t* checked, float first_bad[2])
{
    uint64_t bad = 0, n = 0;
    uint32_t lcg = 12345u;
    for (uint32_t side = first_side; side <= max_side; side++)
    {
        const float b = (float)side;
        const volatile float rbv = 1.0f / b;
        const float rb = rbv;
        for (uint32_t x = 0; x < side; x++)
        {
            for (uint32_t k = 0; k < 3 + jitters; k++)
            {
                float a;
                if (k == 0) a = (float)x + 0.5f;
                else if (k == 1) a = (float)x;
                else if (k == 2) a = (float)x + 1.0f;
                else
                {
                    lcg = lcg * 1664525u + 1013904223u;
                    a = (float)x + (float)(lcg >> 8) * 5.9604645e-8f; /* one rounding, like __fadd_rn(px, u01) */
                }
                const volatile float want = a / b;
                n++;
                if (div_by_const(a, b, rb) != want && bad++ == 0) { first_bad[0] = a; first_bad[1] = b; }
            }
        }
    }
    *checked = n;
    return bad;
}
