/* TEST INFRASTRUCTURE: compares lane_slam_b200/csrc/libm_f32.cuh (compiled for the host) with the C library's
 * sinf / cosf / atanf / atan2f, bit for bit.  Built and run by tests/test_libm_f32.py. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "../../lane_slam_b200/csrc/libm_f32.cuh"

static long check_sincos(uint32_t stride)
{
    long bad = 0;
    const uint32_t hi = lmf_f2u(3.1415927f) + 2;
    for (uint64_t u = 0; u <= hi; u += stride) {
        for (int sg = 0; sg < 2; ++sg) {
            const float v = sg ? -lmf_u2f((uint32_t)u) : lmf_u2f((uint32_t)u);
            if (lmf_f2u(sinf(v)) != lmf_f2u(lmf_sinf(v))) ++bad;
            if (lmf_f2u(cosf(v)) != lmf_f2u(lmf_cosf(v))) ++bad;
        }
    }
    return bad;
}

static long check_atan(uint32_t stride, long pairs)
{
    long bad = 0;
    for (uint64_t u = 0; u < 0x7f800000ull; u += stride) {
        const float x = lmf_u2f((uint32_t)u);
        if (lmf_f2u(atanf(x)) != lmf_f2u(lmf_atanf(x))) ++bad;
        if (lmf_f2u(atanf(-x)) != lmf_f2u(lmf_atanf(-x))) ++bad;
    }
    uint64_t s = 88172645463325252ull;
    for (long i = 0; i < pairs; ++i) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        float y, x;
        if (i & 1) {
            y = lmf_u2f((uint32_t)s); x = lmf_u2f((uint32_t)(s >> 32));
            if (!(fabsf(y) < 1e30f) || !(fabsf(x) < 1e30f)) continue;
        } else {   /* segment-like: endpoint differences on a 1/4096 grid, |d| < 2048 */
            y = (float)((int32_t)(s & 0xffffff) - 0x800000) * (1.0f / 4096.0f);
            x = (float)((int32_t)((s >> 24) & 0xffffff) - 0x800000) * (1.0f / 4096.0f);
        }
        if (lmf_f2u(atan2f(y, x)) != lmf_f2u(lmf_atan2f(y, x))) ++bad;
    }
    /* axis cases */
    const float ax[6] = {0.0f, -0.0f, 1.0f, -1.0f, 37.25f, -1e-3f};
    for (int a = 0; a < 6; ++a)
        for (int b = 0; b < 6; ++b)
            if (lmf_f2u(atan2f(ax[a], ax[b])) != lmf_f2u(lmf_atan2f(ax[a], ax[b]))) ++bad;
    return bad;
}

int main(int argc, char **argv)
{
    const uint32_t stride = argc > 1 ? (uint32_t)atoi(argv[1]) : 1;
    const long pairs = argc > 2 ? atol(argv[2]) : 1500000000L;
    const long b1 = check_sincos(stride), b2 = check_atan(stride, pairs);
    printf("sincos_bad %ld atan_bad %ld\n", b1, b2);
    return (b1 || b2) ? 1 : 0;
}
