/* tests/rt_math_host.c — TEST INFRASTRUCTURE: include/rt_math.h compiled for the host, array-wise, for tests/test_rt_math.py
 * (built by the test with gcc -O2 -ffp-contract=off, the flags of the oracle). */
#include <stdint.h>
#include "rt_math.h"

void rtm_eval(int fn, const float* a, const float* b, float* out, uint64_t n)
{
    for (uint64_t i = 0; i < n; ++i)
    {
        float x = a[i], y = b[i], r = 0.0f;
        switch (fn)
        {
        case 0: r = rt_sinf(x); break;
        case 1: r = rt_cosf(x); break;
        case 2: r = rt_tanf(x); break;
        case 3: r = rt_atan2f(x, y); break;
        case 4: r = rt_acosf(x); break;
        case 5: r = rt_powf(x, y); break;
        case 6: r = rt_fminf(x, y); break;
        case 7: r = rt_fmaxf(x, y); break;
        case 8: r = 1.0f / sqrtf(x); break;
        case 9: r = x / y; break;
        }
        out[i] = r;
    }
}
