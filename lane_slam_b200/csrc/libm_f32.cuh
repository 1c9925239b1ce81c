// libm_f32.cuh -- float32 sinf / cosf / atan2f that are bit-identical to glibc 2.39's (x86-64) on the range the
// LBD path needs, so that the KeyLine angle and the support-region direction vector come out exactly as when the
// reference's line_descriptor C++ runs on a Linux host:
//   kl.angle = atan2(dy, dx)          LSDDetector_custom.cpp:190   (float overload -> glibc atan2f)
//   dL = (cos(direction), sin(..))    binary_descriptor_custom.cpp:1130-1131 (float overloads -> glibc cosf / sinf;
//                                     verified against the compiled reference, oracle/_ref: the double functions
//                                     rounded to float differ from it in the last bit for ~2 % of the lines)
// glibc's float functions are not correctly rounded, so a correctly rounded result is NOT a substitute.  These are
// restatements of the published algorithms:
//   sinf / cosf: the double-precision polynomial scheme glibc uses since 2.28 (fast reduction by pi/2 for |x| < 120,
//                odd / even minimax polynomials evaluated in double, one rounding to float);
//   atan2f / atanf: the fdlibm float scheme (four-interval argument reduction, odd/even split of an 11-term polynomial).
// tests/test_libm_f32.py compiles this header for the host (plain C arithmetic, no contraction) and compares it
// with the C library: every float in [-pi, pi] for sinf / cosf, a stride of all floats for atanf, random and
// segment-like pairs for atan2f (exhaustive runs: 0 mismatches over 2.2e9 / 4.3e9 / 1.3e9 inputs).
// Only finite inputs are supported (segment endpoints are finite); |x| >= 120 for sinf / cosf is not needed.
#pragma once
#include <stdint.h>
#include <string.h>

#ifdef __CUDA_ARCH__
#define LMF_FN __device__ __forceinline__
#define LMF_FMUL(a, b) __fmul_rn((a), (b))
#define LMF_FADD(a, b) __fadd_rn((a), (b))
#define LMF_FSUB(a, b) __fsub_rn((a), (b))
#define LMF_FDIV(a, b) __fdiv_rn((a), (b))
#define LMF_DMUL(a, b) __dmul_rn((a), (b))
#define LMF_DADD(a, b) __dadd_rn((a), (b))
#define LMF_D2F(a) __double2float_rn(a)
#define LMF_F2U(f) __float_as_uint(f)
#define LMF_U2F(u) __uint_as_float(u)
#else
#define LMF_FN static inline
#define LMF_FMUL(a, b) ((float)(a) * (float)(b))
#define LMF_FADD(a, b) ((float)(a) + (float)(b))
#define LMF_FSUB(a, b) ((float)(a) - (float)(b))
#define LMF_FDIV(a, b) ((float)(a) / (float)(b))
#define LMF_DMUL(a, b) ((double)(a) * (double)(b))
#define LMF_DADD(a, b) ((double)(a) + (double)(b))
#define LMF_D2F(a) ((float)(a))
static inline uint32_t lmf_f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float lmf_u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
#define LMF_F2U(f) lmf_f2u(f)
#define LMF_U2F(u) lmf_u2f(u)
#endif

// ---- sinf / cosf ---------------------------------------------------------------------------------------------------
// polynomial coefficients: cos(r) ~ c0 + c1 r^2 + .. + c4 r^8, sin(r) ~ r + s1 r^3 + s2 r^5 + s3 r^7 on [-pi/4, pi/4]
#define LMF_HPI_INV 0x1.45F306DC9C883p+23   /* 2/pi * 2^24 */
#define LMF_HPI 0x1.921FB54442D18p0         /* pi/2 */
#define LMF_C1 (-0x1.ffffffd0c621cp-2)
#define LMF_C2 0x1.55553e1068f19p-5
#define LMF_C3 (-0x1.6c087e89a359dp-10)
#define LMF_C4 0x1.99343027bf8c3p-16
#define LMF_S1 (-0x1.555545995a603p-3)
#define LMF_S2 0x1.1107605230bc4p-7
#define LMF_S3 (-0x1.994eb3774cf24p-13)

// sin polynomial (odd == 0) or cos polynomial (odd == 1) of the reduced argument; neg: cos coefficients negated
LMF_FN float lmf_sincos_poly(double x, double x2, int odd, int neg)
{
    if (!odd) {
        const double x3 = LMF_DMUL(x, x2);
        const double s1 = LMF_DADD(LMF_S2, LMF_DMUL(x2, LMF_S3));
        const double x7 = LMF_DMUL(x3, x2);
        const double s = LMF_DADD(x, LMF_DMUL(x3, LMF_S1));
        return LMF_D2F(LMF_DADD(s, LMF_DMUL(x7, s1)));
    }
    const double sg = neg ? -1.0 : 1.0;
    const double x4 = LMF_DMUL(x2, x2);
    const double c2 = LMF_DADD(sg * LMF_C3, LMF_DMUL(x2, sg * LMF_C4));
    const double c1 = LMF_DADD(sg * 1.0, LMF_DMUL(x2, sg * LMF_C1));
    const double x6 = LMF_DMUL(x4, x2);
    const double c = LMF_DADD(c1, LMF_DMUL(x4, sg * LMF_C2));
    return LMF_D2F(LMF_DADD(c, LMF_DMUL(x6, c2)));
}

LMF_FN uint32_t lmf_abstop12(float x) { return (LMF_F2U(x) >> 20) & 0x7ffu; }

// is_cos = 0: sinf(y); 1: cosf(y).  |y| < 120.
LMF_FN float lmf_sincosf(float y, int is_cos)
{
    double x = (double)y;
    if (lmf_abstop12(y) < 0x3f4u) {                          // |y| < pi/4 (compared on the top 12 bits, like glibc)
        if (lmf_abstop12(y) < 0x398u) return is_cos ? 1.0f : y;   // |y| < 2^-12
        return lmf_sincos_poly(x, LMF_DMUL(x, x), is_cos, 0);
    }
    const double r = LMF_DMUL(x, LMF_HPI_INV);
    const int n = ((int32_t)r + 0x800000) >> 24;              // quadrant, round to nearest
    x = LMF_DADD(x, -LMF_DMUL((double)n, LMF_HPI));
    const double s = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
    return lmf_sincos_poly(LMF_DMUL(x, s), LMF_DMUL(x, x), (n ^ is_cos) & 1, (n & 2) != 0);
}
LMF_FN float lmf_sinf(float y) { return lmf_sincosf(y, 0); }
LMF_FN float lmf_cosf(float y) { return lmf_sincosf(y, 1); }

// ---- atanf / atan2f ------------------------------------------------------------------------------------------------
LMF_FN float lmf_atanf(float x)
{
    const float hi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
    const float lo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
    const float a0 = 3.3333334327e-01f, a1 = -2.0000000298e-01f, a2 = 1.4285714924e-01f, a3 = -1.1111110449e-01f,
                a4 = 9.0908870101e-02f, a5 = -7.6918758452e-02f, a6 = 6.6610731184e-02f, a7 = -5.8335702866e-02f,
                a8 = 4.9768779427e-02f, a9 = -3.6531571299e-02f, a10 = 1.6285819933e-02f;
    const int32_t hx = (int32_t)LMF_F2U(x), ix = hx & 0x7fffffff;
    int id;
    if (ix >= 0x4c000000) {                                   // |x| >= 2^25
        const float r = LMF_FADD(hi[3], lo[3]);
        return hx > 0 ? r : -r;
    }
    if (ix < 0x3ee00000) {                                    // |x| < 0.4375
        if (ix < 0x31000000) return x;                        // |x| < 2^-29
        id = -1;
    } else {
        x = LMF_U2F((uint32_t)ix);
        if (ix < 0x3f980000) {                                // |x| < 1.1875
            if (ix < 0x3f300000) { id = 0; x = LMF_FDIV(LMF_FSUB(LMF_FMUL(2.0f, x), 1.0f), LMF_FADD(2.0f, x)); }
            else { id = 1; x = LMF_FDIV(LMF_FSUB(x, 1.0f), LMF_FADD(x, 1.0f)); }
        } else {
            if (ix < 0x401c0000) { id = 2; x = LMF_FDIV(LMF_FSUB(x, 1.5f), LMF_FADD(1.0f, LMF_FMUL(1.5f, x))); }   // |x| < 2.4375
            else { id = 3; x = LMF_FDIV(-1.0f, x); }
        }
    }
    const float z = LMF_FMUL(x, x), w = LMF_FMUL(z, z);
    float t = LMF_FADD(a8, LMF_FMUL(w, a10));
    t = LMF_FADD(a6, LMF_FMUL(w, t)); t = LMF_FADD(a4, LMF_FMUL(w, t)); t = LMF_FADD(a2, LMF_FMUL(w, t)); t = LMF_FADD(a0, LMF_FMUL(w, t));
    const float s1 = LMF_FMUL(z, t);
    float u = LMF_FADD(a7, LMF_FMUL(w, a9));
    u = LMF_FADD(a5, LMF_FMUL(w, u)); u = LMF_FADD(a3, LMF_FMUL(w, u)); u = LMF_FADD(a1, LMF_FMUL(w, u));
    const float s2 = LMF_FMUL(w, u);
    const float xs = LMF_FMUL(x, LMF_FADD(s1, s2));
    if (id < 0) return LMF_FSUB(x, xs);
    const float r = LMF_FSUB(hi[id], LMF_FSUB(LMF_FSUB(xs, lo[id]), x));
    return hx < 0 ? -r : r;
}

LMF_FN float lmf_atan2f(float y, float x)
{
    const float tiny = 1.0e-30f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
    const int32_t hx = (int32_t)LMF_F2U(x), ix = hx & 0x7fffffff, hy = (int32_t)LMF_F2U(y), iy = hy & 0x7fffffff;
    if (hx == 0x3f800000) return lmf_atanf(y);                // x == 1
    const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);        // 2 * sign(x) + sign(y)
    if (iy == 0) return m < 2 ? y : (m == 2 ? LMF_FADD(pi, tiny) : LMF_FSUB(-pi, tiny));
    if (ix == 0) return hy < 0 ? LMF_FSUB(-pi_o_2, tiny) : LMF_FADD(pi_o_2, tiny);
    const int k = (iy - ix) >> 23;
    float z;
    if (k > 60) z = LMF_FADD(pi_o_2, LMF_FMUL(0.5f, pi_lo));
    else if (hx < 0 && k < -60) z = 0.0f;
    else z = lmf_atanf(LMF_U2F(LMF_F2U(LMF_FDIV(y, x)) & 0x7fffffffu));
    switch (m) {
    case 0: return z;
    case 1: return LMF_U2F(LMF_F2U(z) ^ 0x80000000u);
    case 2: return LMF_FSUB(pi, LMF_FSUB(z, pi_lo));
    default: return LMF_FSUB(LMF_FSUB(z, pi_lo), pi);
    }
}
