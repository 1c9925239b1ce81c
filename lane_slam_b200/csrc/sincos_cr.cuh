// sincos_cr.cuh -- sin/cos of a double, evaluated in double-double and rounded once.
//
// Why: LSD's rectangle corners are c + l*(cos t, sin t) +- (w/2)*(-sin t, cos t).  A corner frequently
// coincides (mathematically) with a pixel centre, and rect_nfa then truncates that coordinate to an
// integer -- so the LAST BIT of cos/sin decides whether a pixel is counted.  The reference runs glibc's
// sin/cos (< 1 ulp, correctly rounded in all but ~1e-3 of cases); CUDA's sin/cos differ from them in
// the last bit far too often.  This version is correctly rounded except when the true value lies within
// ~2^-45 ulp of a rounding boundary.
//
// Method: t = k*pi/512 + r (double-double reduction), table of sin/cos(j*pi/512) in double-double
// (tools/gen_sincos_table.py), degree-7/6 Taylor kernels in double-double, angle addition, one rounding.
#pragma once
#include <math.h>

#if defined(__CUDA_ARCH__)
#define SCR_FN __device__ __forceinline__
#define SCR_FMA(a, b, c) __fma_rn((a), (b), (c))
#define SCR_TAB(j, i) c_sincos_tab[(j) * 4 + (i)]
#else
#define SCR_FN static inline
#define SCR_FMA(a, b, c) fma((a), (b), (c))
#define SCR_TAB(j, i) kSinCosTab[j][i]
#endif

namespace scr {
#include "sincos_tab.inc"

struct dd { double hi, lo; };

SCR_FN dd two_sum(double a, double b) { double s = a + b, bb = s - a; dd r = {s, (a - (s - bb)) + (b - bb)}; return r; }
SCR_FN dd quick_two_sum(double a, double b) { double s = a + b; dd r = {s, b - (s - a)}; return r; }
SCR_FN dd two_prod(double a, double b) { double p = a * b; dd r = {p, SCR_FMA(a, b, -p)}; return r; }
SCR_FN dd dd_add(dd a, dd b)
{
    dd s = two_sum(a.hi, b.hi), t = two_sum(a.lo, b.lo);
    s.lo += t.hi; s = quick_two_sum(s.hi, s.lo);
    s.lo += t.lo; return quick_two_sum(s.hi, s.lo);
}
SCR_FN dd dd_add_d(dd a, double b)
{
    dd s = two_sum(a.hi, b);
    s.lo += a.lo; return quick_two_sum(s.hi, s.lo);
}
SCR_FN dd dd_mul(dd a, dd b)
{
    dd p = two_prod(a.hi, b.hi);
    p.lo += a.hi * b.lo + a.lo * b.hi; return quick_two_sum(p.hi, p.lo);
}
SCR_FN dd dd_mul_d(dd a, double b)
{
    dd p = two_prod(a.hi, b);
    p.lo += a.lo * b; return quick_two_sum(p.hi, p.lo);
}
SCR_FN dd dd_neg(dd a) { dd r = {-a.hi, -a.lo}; return r; }
}  // namespace scr

#if defined(__CUDACC__)
__constant__ double c_sincos_tab[257 * 4];
#endif

// valid for 0 <= t < 64 (LSD angles are in [0, 3*pi]); returns false outside (caller falls back)
SCR_FN bool sincos_cr(double t, double *s_out, double *c_out)
{
    using namespace scr;
    if (!(t >= 0.0 && t < 64.0)) return false;
    double kd = rint(t * k512oPi);
    int k = (int)kd;
    // r = t - k*pi/512 in double-double
    dd kp = two_prod(kd, kPio512Hi);
    dd r = two_sum(t, -kp.hi);
    r = dd_add_d(r, -(kp.lo + kd * kPio512Lo));
    // Taylor kernels
    dd r2 = dd_mul(r, r);
    // sin r = r + r * ps,  ps = ((-1/5040 r2 + 1/120) r2 - 1/6) r2
    dd ps = dd_mul_d(r2, -1.0 / 5040.0);
    ps = dd_add_d(ps, 1.0 / 120.0);
    ps = dd_mul(ps, r2);
    { dd m16 = {kM16Hi, kM16Lo}; ps = dd_add(ps, m16); }
    ps = dd_mul(ps, r2);
    dd sr = dd_add(r, dd_mul(r, ps));
    // cos r = 1 + pc,  pc = (((1/40320 r2 - 1/720) r2 + 1/24) r2 - 1/2) r2
    dd pc = dd_mul_d(r2, 1.0 / 40320.0);
    pc = dd_add_d(pc, -1.0 / 720.0);
    pc = dd_mul(pc, r2);
    pc = dd_add_d(pc, 1.0 / 24.0);   // 1/24 in double: error 2^-58 * r2^2 ~ 2^-91
    pc = dd_mul(pc, r2);
    pc = dd_add_d(pc, -0.5);
    pc = dd_mul(pc, r2);
    dd cr = dd_add_d(pc, 1.0);
    // table
    int q = k >> 8, j = k & 255;
    dd S = {SCR_TAB(j, 0), SCR_TAB(j, 1)}, C = {SCR_TAB(j, 2), SCR_TAB(j, 3)};
    // sin(a + r), cos(a + r)
    dd st = dd_add(dd_mul(S, cr), dd_mul(C, sr));
    dd ct = dd_add(dd_mul(C, cr), dd_neg(dd_mul(S, sr)));
    switch (q & 3) {
    case 0: *s_out = st.hi; *c_out = ct.hi; break;
    case 1: *s_out = ct.hi; *c_out = -st.hi; break;
    case 2: *s_out = -st.hi; *c_out = -ct.hi; break;
    default: *s_out = -ct.hi; *c_out = st.hi; break;
    }
    return true;
}
