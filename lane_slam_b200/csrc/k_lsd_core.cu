// k_lsd_core.cu -- K8+K9: the LSD search on one colour image: pseudo-ordered seeds (1024 bins,
// descending, raster order inside a bin), region growing, rectangle fit, density refinement,
// NFA validation with LSD_REFINE_ADV rectangle improvement, segment emission.
//
// Replaces the second half of cv2.LineSegmentDetector.detect (line_detector_lsd.py:64-72; SURVEY.md A.6;
// NFA math == src/line_descriptor/include/line_descriptor/descriptor_custom.hpp:676-826).
//
// One warp owns one (frame, colour) image.  Region growing commits pixels strictly in the reference's
// order (FIFO over region points, 3x3 neighbours row-major, region angle updated after every accepted
// pixel), because the result depends on it; the warp evaluates the nine neighbours of a point in
// parallel and resolves acceptances in order.  Rectangle sums, the NFA pixel scan and the seed sort are
// warp-parallel (ballot / shuffle reductions, stable counting sort with match_any).
#include <cstdio>
#include <vector>
#include "common.cuh"
#include "sincos_cr.cuh"

namespace lsf {

#define FULL 0xffffffffu
constexpr double kPI = 3.14159265358979323846;
constexpr double k3_2PI = 3.0 * kPI / 2.0;
constexpr double k2PI = 2.0 * kPI;
constexpr double kDEG2RAD = kPI / 180.0;
constexpr double kLN10 = 2.30258509299404568402;

struct Rect { double x1, y1, x2, y2, width, x, y, theta, dx, dy, prec, p; };

struct Img {
    const LsdWord *words;
    LsdPix *pix;
    const u32 *nbr;
    u32 *reg;
    int W, H, swp, n, cap;
    double logNT;
};

__device__ __forceinline__ double wsum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ double wmin(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ double wmax(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
    return v;
}

// compact index of scaled pixel (x,y) or -1 when its level-line angle is undefined
__device__ __forceinline__ int lookup(const Img &im, int x, int y)
{
    uint2 wd = *reinterpret_cast<const uint2 *>(&im.words[(size_t)y * im.swp + (x >> 5)]);
    u32 b = x & 31;
    if (!((wd.x >> b) & 1u)) return -1;
    int idx = (int)(wd.y + __popc(wd.x & ((1u << b) - 1u)));
    return idx < im.n ? idx : -1;
}

__device__ __forceinline__ bool aligned_ang(double a, double theta, double prec)
{
    double n = fabs(theta - a);
    if (n > k3_2PI) n = fabs(n - k2PI);
    return n <= prec;
}

__device__ __forceinline__ double angle_diff_signed(double a, double b)
{
    double diff = a - b;
    while (diff <= -kPI) diff += k2PI;
    while (diff > kPI) diff -= k2PI;
    return diff;
}

// ---- NFA (all lanes compute the same value) -------------------------------------------------------
__host__ __device__ inline double log_gamma_d(double x)
{
    if (x > 15.0)
        return 0.918938533204673 + (x - 0.5) * log(x) - x + 0.5 * x * log(x * sinh(1 / x) + 1 / (810.0 * pow(x, 6.0)));
    const double q[7] = {75122.6331530, 80916.6278952, 36308.2951477, 8687.24529705, 1168.92649479, 83.8676043424,
                         2.50662827511};
    double a = (x + 0.5) * log(x + 5.5) - (x + 5.5), b = 0;
    for (int n = 0; n < 7; ++n) {
        a -= log(x + (double)n);
        b += q[n] * pow(x, (double)n);
    }
    return a + log(b);
}

__device__ bool double_equal_d(double a, double b)
{
    if (a == b) return true;
    double diff = fabs(a - b), aa = fabs(a), bb = fabs(b), m = aa > bb ? aa : bb;
    if (m < 2.2250738585072014e-308) m = 2.2250738585072014e-308;
    return (diff / m) <= 100.0 * 2.220446049250313e-16;
}

// log_gamma of the integers 0..LGTAB-1, evaluated on the host with the same formulas (and the host libm
// the reference itself runs on); rect_nfa only ever asks for integer arguments.
constexpr int LGTAB = 16384;
__device__ const double *g_lgtab;

__device__ __forceinline__ double lg_int(int v)
{
    return v < LGTAB ? g_lgtab[v] : log_gamma_d((double)v);
}

__device__ double nfa_d(int n, int k, double p, double logNT)
{
    if (n == 0 || k == 0) return -logNT;
    if (n == k) return -logNT - (double)n * log10(p);
    double p_term = p / (1 - p);
    double log1term = lg_int(n + 1) - lg_int(k + 1) - lg_int(n - k + 1) +
                      (double)k * log(p) + (double)(n - k) * log(1.0 - p);
    double term = exp(log1term);
    if (double_equal_d(term, 0)) {
        if (k > n * p) return -log1term / kLN10 - logNT;
        return -logNT;
    }
    double bin_tail = term;
    for (int i = k + 1; i <= n; ++i) {
        double bin_term = (double)(n - i + 1) / (double)i;
        double mult_term = bin_term * p_term;
        term *= mult_term;
        bin_tail += term;
        if (bin_term < 1) {
            double err = term * ((1 - pow(mult_term, (double)(n - i + 1))) / (1 - mult_term) - 1);
            if (err < 0.1 * fabs(-log10(bin_tail) - logNT) * bin_tail) break;
        }
    }
    return -log10(bin_tail) - logNT;
}

// ---- region growing -----------------------------------------------------------------------------------
// Grows from compact index `seed` with tolerance prec; fills im.reg[0..nreg) (acceptance order), marks
// pixels used, returns nreg and the final region angle.
//
// The FIFO is consumed three points per round: lanes 0-8 / 9-17 / 18-26 hold the 3x3 neighbourhoods (row
// major, centre idle) of three consecutive queue entries, fetched with one round trip (queue entry ->
// neighbour table -> pixel record).  Acceptances are then resolved strictly in the reference's order: group
// by group, neighbour by neighbour, the region angle updated after every accepted pixel.  A pixel accepted
// earlier in the round is dropped from the later groups (it is USED by then in the reference as well).
__device__ int grow(const Img &im, int seed, double prec, double &reg_angle_out)
{
    const int lane = threadIdx.x & 31;
    const int grp = lane / 9, slot = lane - 9 * grp;
    const bool nb_lane = grp < 3 && slot != 4;
    const int nslot = slot < 4 ? slot : slot - 1;
    int nreg = 1;
    double reg_angle = im.pix[seed].ang;
    float sumdx = (float)cos(reg_angle), sumdy = (float)sin(reg_angle);
    if (lane == 0) { im.reg[0] = (u32)seed; im.pix[seed].used = 1; }
    __syncwarp();
    for (int i = 0; i < nreg;) {
        const int navail = min(3, nreg - i);
        int idx = -1;
        double ang = 0;
        float c = 0, s = 0;
        if (nb_lane && grp < navail) {
            u32 p = im.reg[i + grp];
            u32 ni = im.nbr[(size_t)p * 8 + nslot];
            if (ni != LSD_NONE) {
                const uint4 *rp = reinterpret_cast<const uint4 *>(&im.pix[ni]);
                uint4 r0 = rp[0], r1 = rp[1];
                if (r1.z == 0) {   // not USED
                    idx = (int)ni;
                    ang = __hiloint2double((int)r0.y, (int)r0.x);
                    c = __uint_as_float(r0.z); s = __uint_as_float(r0.w);
                }
            }
        }
        for (int g = 0; g < navail; ++g) {
            u32 pending = __ballot_sync(FULL, idx >= 0) & (0x1ffu << (9 * g));
            while (pending) {
                bool al = idx >= 0 && aligned_ang(ang, reg_angle, prec);
                u32 m = __ballot_sync(FULL, al) & pending;
                if (!m) break;
                int k = __ffs(m) - 1;
                int idxk = __shfl_sync(FULL, idx, k);
                float ck = __shfl_sync(FULL, c, k), sk = __shfl_sync(FULL, s, k);
                if (lane == k) im.pix[idx].used = 1;
                if (idx == idxk) idx = -1;          // the accepted lane and its duplicates in later groups
                if (lane == 0) im.reg[nreg] = (u32)idxk;
                ++nreg;
                sumdx = __fadd_rn(sumdx, ck);
                sumdy = __fadd_rn(sumdy, sk);
                reg_angle = (double)fast_atan2_deg(sumdy, sumdx) * kDEG2RAD;
                pending &= ~((2u << k) - 1u);
            }
        }
        __syncwarp();  // used / reg writes of this round are visible to the next one
        i += navail;
    }
    reg_angle_out = reg_angle;
    return nreg;
}

// ---- reference-order accumulation --------------------------------------------------------------------
// The reference adds region points one by one in double precision; rounding (and through it theta, the
// rectangle corners and finally the NFA pixel counts) depends on that order.  Lanes compute the per-point
// terms in parallel, park them in shared memory, and every lane then adds them in point order.
struct Seq3 { double a[32], b[32], c[32]; };

__device__ __forceinline__ void seq_add3(Seq3 &sm, int cnt, double ta, double tb, double tc, double &A, double &B, double &C)
{
    const int lane = threadIdx.x & 31;
    __syncwarp();
    sm.a[lane] = ta; sm.b[lane] = tb; sm.c[lane] = tc;
    __syncwarp();
    for (int j = 0; j < cnt; ++j) { A += sm.a[j]; B += sm.b[j]; C += sm.c[j]; }
}

// ---- rectangle fit ------------------------------------------------------------------------------------
__device__ void region2rect(const Img &im, Seq3 &sm, int nreg, double reg_angle, double prec, double p, Rect &r)
{
    const int lane = threadIdx.x & 31;
    double x = 0, y = 0, sum = 0;
    for (int i0 = 0; i0 < nreg; i0 += 32) {
        int i = i0 + lane;
        double tx = 0, ty = 0, tw = 0;
        if (i < nreg) {
            u32 q = im.reg[i];
            u32 xy = im.pix[q].xy;
            tw = sqrt((double)im.pix[q].g2 / 4.0);
            tx = (double)(xy & 0xffffu) * tw;
            ty = (double)(xy >> 16) * tw;
        }
        seq_add3(sm, min(32, nreg - i0), tx, ty, tw, x, y, sum);
    }
    x /= sum; y /= sum;
    double Ixx = 0, Iyy = 0, Ixy = 0;
    for (int i0 = 0; i0 < nreg; i0 += 32) {
        int i = i0 + lane;
        double ta = 0, tb = 0, tc = 0;
        if (i < nreg) {
            u32 q = im.reg[i];
            u32 xy = im.pix[q].xy;
            double w = sqrt((double)im.pix[q].g2 / 4.0);
            double dx = (double)(xy & 0xffffu) - x, dy = (double)(xy >> 16) - y;
            ta = dy * dy * w; tb = dx * dx * w; tc = -(dx * dy * w);   // Ixy -= dx*dy*w
        }
        seq_add3(sm, min(32, nreg - i0), ta, tb, tc, Ixx, Iyy, Ixy);
    }
    double lambda = 0.5 * (Ixx + Iyy - sqrt((Ixx - Iyy) * (Ixx - Iyy) + 4.0 * Ixy * Ixy));
    double theta = (fabs(Ixx) > fabs(Iyy)) ? (double)fast_atan2_deg((float)(lambda - Ixx), (float)Ixy)
                                           : (double)fast_atan2_deg((float)Ixy, (float)(lambda - Iyy));
    theta *= kDEG2RAD;
    if (fabs(angle_diff_signed(theta, reg_angle)) > prec) theta += kPI;
    // the last bit of cos/sin decides pixel membership at rectangle corners: use the correctly rounded pair
    double dx, dy;
    if (!sincos_cr(theta, &dy, &dx)) { dx = cos(theta); dy = sin(theta); }
    // extents: min / max are order independent
    double lmin = 0, lmax = 0, wmn = 0, wmx = 0;
    for (int i = lane; i < nreg; i += 32) {
        u32 xy = im.pix[im.reg[i]].xy;
        double rdx = (double)(xy & 0xffffu) - x, rdy = (double)(xy >> 16) - y;
        double l = rdx * dx + rdy * dy, w = -rdx * dy + rdy * dx;
        lmax = fmax(lmax, l); lmin = fmin(lmin, l);
        wmx = fmax(wmx, w); wmn = fmin(wmn, w);
    }
    lmin = wmin(lmin); lmax = wmax(lmax); wmn = wmin(wmn); wmx = wmax(wmx);
    r.x1 = x + lmin * dx; r.y1 = y + lmin * dy; r.x2 = x + lmax * dx; r.y2 = y + lmax * dy;
    r.width = wmx - wmn; r.x = x; r.y = y; r.theta = theta; r.dx = dx; r.dy = dy; r.prec = prec; r.p = p;
    if (r.width < 1.0) r.width = 1.0;
}

__device__ __forceinline__ double rect_density(int nreg, const Rect &r)
{
    double dd = sqrt((r.x1 - r.x2) * (r.x1 - r.x2) + (r.y1 - r.y2) * (r.y1 - r.y2));
    return (double)nreg / (dd * r.width);
}

// ---- density refinement (refine + reduce_region_radius) ------------------------------------------------
__device__ bool refine(const Img &im, Seq3 &sm, int &nreg, double &reg_angle, double prec, double p, Rect &rec)
{
    const double density_th = 0.7;
    const int lane = threadIdx.x & 31;
    if (rect_density(nreg, rec) >= density_th) return true;
    const int seed = (int)im.reg[0];
    const u32 sxy = im.pix[seed].xy;
    const double xc = (double)(sxy & 0xffffu), yc = (double)(sxy >> 16);
    const double ang_c = im.pix[seed].ang;
    double sum = 0, s_sum = 0, cnt = 0;
    for (int i0 = 0; i0 < nreg; i0 += 32) {
        int i = i0 + lane;
        double ta = 0, tb = 0, tc = 0;
        if (i < nreg) {
            u32 q = im.reg[i];
            im.pix[q].used = 0;
            u32 xy = im.pix[q].xy;
            double ex = (double)(xy & 0xffffu) - xc, ey = (double)(xy >> 16) - yc;
            if (sqrt(ex * ex + ey * ey) < rec.width) {
                double ad = angle_diff_signed(im.pix[q].ang, ang_c);
                ta = ad; tb = ad * ad; tc = 1.0;
            }
        }
        seq_add3(sm, min(32, nreg - i0), ta, tb, tc, sum, s_sum, cnt);
    }
    __syncwarp();
    double mean = sum / cnt;
    double tau = 2.0 * sqrt((s_sum - 2.0 * mean * sum) / cnt + mean * mean);
    nreg = grow(im, seed, tau, reg_angle);
    if (nreg < 2) return false;
    region2rect(im, sm, nreg, reg_angle, prec, p, rec);
    double density = rect_density(nreg, rec);
    if (density < density_th) {
        double r1 = (xc - rec.x1) * (xc - rec.x1) + (yc - rec.y1) * (yc - rec.y1);
        double r2 = (xc - rec.x2) * (xc - rec.x2) + (yc - rec.y2) * (yc - rec.y2);
        double radSq = r1 > r2 ? r1 : r2;
        while (density < density_th) {
            radSq *= 0.75 * 0.75;
            // Drop points farther than the radius exactly like the reference's swap-with-last loop
            // (reg[i] <- reg[last], pop, re-test i): kept points before position K stay in place and the
            // j-th hole (ascending) receives the j-th kept point counted from the back.  The order matters
            // because the following rectangle sums are accumulated in region order.
            u32 *scratch = im.reg + im.cap;
            int K = 0;
            for (int i0 = 0; i0 < nreg; i0 += 32) {
                int i = i0 + lane;
                bool keep = false;
                if (i < nreg) {
                    u32 q = im.reg[i];
                    u32 xy = im.pix[q].xy;
                    double ex = xc - (double)(xy & 0xffffu), ey = yc - (double)(xy >> 16);
                    keep = !(ex * ex + ey * ey > radSq);
                    if (!keep) im.pix[q].used = 0;
                }
                K += __popc(__ballot_sync(FULL, keep));
            }
            // fillers: kept points at positions >= K, ranked from the back
            int before = 0;   // kept points in [0, i0)
            for (int i0 = 0; i0 < nreg; i0 += 32) {
                int i = i0 + lane;
                u32 q = 0;
                bool keep = false;
                if (i < nreg) {
                    q = im.reg[i];
                    u32 xy = im.pix[q].xy;
                    double ex = xc - (double)(xy & 0xffffu), ey = yc - (double)(xy >> 16);
                    keep = !(ex * ex + ey * ey > radSq);
                }
                u32 m = __ballot_sync(FULL, keep);
                int incl = before + __popc(m & ((2u << lane) - 1u));
                if (keep && i >= K) scratch[K - incl] = q;          // kept points after i = K - incl
                before += __popc(m);
            }
            __syncwarp();
            before = 0;
            for (int i0 = 0; i0 < K; i0 += 32) {
                int i = i0 + lane;
                bool keep = true;
                if (i < K) {
                    u32 xy = im.pix[im.reg[i]].xy;
                    double ex = xc - (double)(xy & 0xffffu), ey = yc - (double)(xy >> 16);
                    keep = !(ex * ex + ey * ey > radSq);
                }
                u32 m = __ballot_sync(FULL, keep);
                int excl = before + __popc(m & ((1u << lane) - 1u));  // kept points in [0, i)
                __syncwarp();
                if (i < K && !keep) im.reg[i] = scratch[i - excl];   // hole rank = holes before i
                before += __popc(m);
            }
            __syncwarp();
            int nout = K;
            nreg = nout;
            if (nreg < 2) return false;
            region2rect(im, sm, nreg, reg_angle, prec, p, rec);
            density = rect_density(nreg, rec);
        }
    }
    return true;
}

// double -> int like x86 cvttsd2si (what the reference's (int) casts compile to): truncation toward zero,
// INT_MIN for NaN / out-of-range values
__device__ __forceinline__ int to_int_x86(double v)
{
    if (!(v > -2147483649.0 && v < 2147483648.0)) return (int)0x80000000;
    return (int)v;
}

// ---- NFA of a rectangle: scan its pixels row by row (row-scan of the installed cv2 4.13) ---------------
__device__ __forceinline__ double slope_d(double ax, double ay, double bx, double by)
{
    return (ceil(by) != ceil(ay)) ? (bx - ax) / (by - ay) : 0.0;
}

__device__ __noinline__ double rect_nfa(const Img &im, const Rect &r)
{
    const int lane = threadIdx.x & 31;
    double hw = r.width / 2.0, dyhw = r.dy * hw, dxhw = r.dx * hw;
    double vx[4] = {r.x1 - dyhw, r.x2 - dyhw, r.x2 + dyhw, r.x1 + dyhw};
    double vy[4] = {r.y1 + dxhw, r.y2 + dxhw, r.y2 - dxhw, r.y1 - dxhw};
    int off = 0;
    for (int i = 1; i < 4; ++i)
        if (vy[i] == vy[off] ? (vx[i] < vx[off]) : (vy[i] < vy[off])) off = i;
    double ox[4], oy[4];
    for (int i = 0; i < 4; ++i) { ox[i] = vx[(i + off) & 3]; oy[i] = vy[(i + off) & 3]; }
    double fl = slope_d(ox[0], oy[0], ox[1], oy[1]), sl = slope_d(ox[1], oy[1], ox[2], oy[2]);
    double fr = slope_d(ox[0], oy[0], ox[3], oy[3]), sr = slope_d(ox[3], oy[3], ox[2], oy[2]);
    // ceil of huge/NaN doubles: clamp the row range to the image first
    double ysd = ceil(oy[0]), yed = ceil(oy[2]);
    if (!(ysd == ysd) || !(yed == yed)) return nfa_d(0, 0, r.p, im.logNT);
    int ys = ysd < 0 ? 0 : (ysd > (double)im.H ? im.H : (int)ysd);
    int ye = yed < -1 ? -1 : (yed > (double)(im.H - 1) ? im.H - 1 : (int)yed);
    double c1 = ceil(oy[1]), c3 = ceil(oy[3]);
    int tot = 0, alg = 0;
    for (int y0 = ys; y0 <= ye; y0 += 32) {
        int y = y0 + lane;
        int xa = 0, cnt = 0;
        if (y <= ye) {
            double yd = (double)y;
            double ll = (yd <= c1) ? ox[0] + (yd - oy[0]) * fl : ox[1] + (yd - oy[1]) * sl;
            double rl = (yd < c3) ? ox[0] + (yd - oy[0]) * fr : ox[3] + (yd - oy[3]) * sr;
            xa = max(to_int_x86(ceil(ll)), 0);
            int xb = min(to_int_x86(rl), im.W - 1);
            cnt = xb >= xa ? xb - xa + 1 : 0;
        }
        // inclusive scan of the span lengths
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int v = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += v;
        }
        int T = __shfl_sync(FULL, incl, 31);
        tot += T;
        for (int t = lane; t < ((T + 31) & ~31); t += 32) {
            // row of element t: first lane whose inclusive prefix exceeds t (binary search over lanes)
            int lo = 0;
#pragma unroll
            for (int st = 16; st > 0; st >>= 1) {
                int v = __shfl_sync(FULL, incl, lo + st - 1);
                if (v <= t) lo += st;
            }
            int r_incl = __shfl_sync(FULL, incl, lo), r_cnt = __shfl_sync(FULL, cnt, lo), r_xa = __shfl_sync(FULL, xa, lo);
            if (t < T) {
                int x = r_xa + (t - (r_incl - r_cnt));
                int idx = lookup(im, x, y0 + lo);
                if (idx >= 0 && aligned_ang(im.pix[idx].ang, r.theta, r.prec)) ++alg;
            }
        }
    }
    alg = __reduce_add_sync(FULL, alg);
    return nfa_d(tot, alg, r.p, im.logNT);
}

// LSD_REFINE_ADV rectangle improvement.  The reference's five trial phases (finer precision, narrower, one
// side in, other side in, finer precision again) run as one loop around a single rect_nfa call site: the
// NFA scan is large, and 21 inlined copies of it made the kernel instruction-cache bound (ncu: no_instruction).
__device__ double rect_improve(const Img &im, Rect &rec)
{
    const double log_eps = 0.0, delta = 0.5, delta_2 = 0.25;
    double log_nfa = 0;
    Rect r = rec;
    for (int step = 0; step <= 25; ++step) {
        const int phase = step == 0 ? 0 : 1 + (step - 1) / 5;
        if (step > 0 && (step - 1) % 5 == 0) {   // phase boundary: stop as soon as the rectangle is meaningful
            if (log_nfa > log_eps) break;
            r = rec;
        }
        bool eval = true;
        if (phase == 1) {
            r.p /= 2; r.prec = r.p * kPI;
        } else if (phase >= 2) {
            eval = (r.width - delta) >= 0.5;
            if (eval) {
                if (phase == 2) {
                    r.width -= delta;
                } else if (phase == 5) {
                    r.p /= 2; r.prec = r.p * kPI;
                } else {
                    const double sgn = phase == 3 ? 1.0 : -1.0;
                    r.x1 += sgn * -r.dy * delta_2; r.y1 += sgn * r.dx * delta_2;
                    r.x2 += sgn * -r.dy * delta_2; r.y2 += sgn * r.dx * delta_2;
                    r.width -= delta;
                }
            }
        }
        if (eval) {
            double v = rect_nfa(im, r);
            if (step == 0 || v > log_nfa) { log_nfa = v; rec = r; }
        }
    }
    return log_nfa;
}

// ---- kernel 1: seeds, region growing, rectangle fit, density refinement -> candidate rectangles ----------
// One warp per (frame, colour) image; strictly sequential in the reference's seed order because growing and
// refining change which pixels later seeds may use.  NFA validation does not touch that state, so it is
// deferred to kernel 2, where every candidate gets its own warp.
__device__ __forceinline__ void setup_img(Img &im, const Dims &d, int img, const LsdWord *lsdw, LsdPix *pix, const u32 *nbr,
                                          u32 *reg, const int *pixcount)
{
    im.n = pixcount[img];
    im.W = d.sw; im.H = d.sh; im.swp = d.swp;
    im.words = lsdw + (size_t)img * d.sh * d.swp;
    im.pix = pix + (size_t)img * d.pixcap;
    im.nbr = nbr ? nbr + (size_t)img * d.pixcap * 8 : nullptr;
    im.reg = reg ? reg + (size_t)img * 2 * d.pixcap : nullptr;   // second half: scratch of reduce_region_radius
    im.cap = d.pixcap;
    im.logNT = 5.0 * (log10((double)im.W) + log10((double)im.H)) / 2.0 + log10(11.0);
}

// ---- kernel 0: 8-neighbour table of the support pixels (fully parallel) ------------------------------------------
__global__ void __launch_bounds__(256) k_lsd_nbr(Dims d, const LsdWord *__restrict__ lsdw, const LsdPix *__restrict__ pix,
                                                const int *__restrict__ pixcount, u32 *__restrict__ nbr)
{
    const int img = blockIdx.x;
    Img im;
    setup_img(im, d, img, lsdw, const_cast<LsdPix *>(pix), nullptr, nullptr, pixcount);
    uint4 *out = reinterpret_cast<uint4 *>(nbr + (size_t)img * d.pixcap * 8);
    for (int i = threadIdx.x; i < im.n; i += 256) {
        u32 xy = im.pix[i].xy;
        int x = (int)(xy & 0xffffu), y = (int)(xy >> 16);
        u32 v[8];
        int k = 0;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                if (dx == 0 && dy == 0) continue;
                int xx = x + dx, yy = y + dy, r = -1;
                if (xx >= 0 && xx < im.W && yy >= 0 && yy < im.H) r = lookup(im, xx, yy);
                v[k++] = r >= 0 ? (u32)r : LSD_NONE;
            }
        out[2 * (size_t)i] = make_uint4(v[0], v[1], v[2], v[3]);
        out[2 * (size_t)i + 1] = make_uint4(v[4], v[5], v[6], v[7]);
    }
}

__global__ void __launch_bounds__(32) k_lsd_grow(Dims d, const LsdWord *__restrict__ lsdw, LsdPix *__restrict__ pix,
                                                const u32 *__restrict__ nbr, u32 *__restrict__ order,
                                                u32 *__restrict__ reg, const int *__restrict__ pixcount,
                                                const u32 *__restrict__ g2max, LsdCand *__restrict__ cand,
                                                int *__restrict__ candcount, uint2 *__restrict__ candlist, int *__restrict__ flags)
{
    __shared__ u32 hist[1024];
    __shared__ Seq3 sm;
    const int img = blockIdx.x, lane = threadIdx.x;
    Img im;
    setup_img(im, d, img, lsdw, pix, nbr, reg, pixcount);
    u32 *ord = order + (size_t)img * d.pixcap;
    const int n = im.n;
    if (n == 0) {
        if (lane == 0) candcount[img] = 0;
        return;
    }
    // ---- seed order: stable counting sort by bin = int(norm * 1023 / max_norm), descending ----
    const double max_grad = sqrt((double)g2max[img] / 4.0);
    const double bin_coef = max_grad > 0 ? 1023.0 / max_grad : 0.0;
    for (int i = lane; i < 1024; i += 32) hist[i] = 0;
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
        int bin = (int)(sqrt((double)im.pix[i].g2 / 4.0) * bin_coef);
        u32 key = (u32)(1023 - bin);
        im.reg[i] = key;  // reg doubles as key scratch during the sort
        atomicAdd(&hist[key], 1u);
    }
    __syncwarp();
    {
        // exclusive scan of 1024 counters: lane owns 32 consecutive entries
        u32 loc = 0;
        for (int j = 0; j < 32; ++j) loc += hist[lane * 32 + j];
        u32 incl = loc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u32 v = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += v;
        }
        u32 run = incl - loc;
        for (int j = 0; j < 32; ++j) { u32 c = hist[lane * 32 + j]; hist[lane * 32 + j] = run; run += c; }
    }
    __syncwarp();
    for (int i0 = 0; i0 < n; i0 += 32) {
        int i = i0 + lane;
        bool valid = i < n;
        u32 key = valid ? im.reg[i] : (2048u + (u32)lane);
        u32 m = __match_any_sync(FULL, key);
        int leader = __ffs(m) - 1;
        u32 off = 0;
        if (valid && lane == leader) { off = hist[key]; hist[key] = off + __popc(m); }
        off = __shfl_sync(FULL, off, leader);
        if (valid) ord[off + __popc(m & ((1u << lane) - 1u))] = (u32)i;
        __syncwarp();
    }
    __syncwarp();

    // ---- search ----
    const double prec = kPI * 22.5 / 180.0, p = 22.5 / 180.0;
    const int min_reg = (int)(-im.logNT / log10(p));
    int ncand = 0;
    LsdCand *out = cand + (size_t)img * d.segcap;
    for (int o0 = 0; o0 < n; o0 += 32) {
        int oi = o0 + lane;
        int ci = oi < n ? (int)ord[oi] : -1;
        int k = -1;
        while (true) {
            // seeds of this batch not yet visited and still unused *now* (refine may have released pixels)
            __syncwarp();
            u32 cnd = __ballot_sync(FULL, ci >= 0 && lane > k && im.pix[ci].used == 0);
            if (!cnd) break;
            k = __ffs(cnd) - 1;
            int seed = __shfl_sync(FULL, ci, k);
            double reg_angle;
            int nreg = grow(im, seed, prec, reg_angle);
            if (nreg < min_reg) continue;
            Rect rec;
            region2rect(im, sm, nreg, reg_angle, prec, p, rec);
            if (!refine(im, sm, nreg, reg_angle, prec, p, rec)) continue;
            if (ncand < d.segcap && lane == 0) {
                LsdCand c;
                c.x1 = rec.x1; c.y1 = rec.y1; c.x2 = rec.x2; c.y2 = rec.y2;
                c.width = rec.width; c.theta = rec.theta; c.dx = rec.dx; c.dy = rec.dy;
                out[ncand] = c;
                int slot = atomicAdd(&flags[3], 1);
                candlist[slot] = make_uint2((u32)img, (u32)ncand);
            }
            ++ncand;
        }
    }
    if (lane == 0) {
        if (ncand > d.segcap) { atomicMax(&flags[1], ncand); ncand = d.segcap; }
        candcount[img] = ncand;
    }
}

// ---- kernel 2: NFA validation with LSD_REFINE_ADV rectangle improvement, one warp per candidate ---------------
constexpr int VAL_WARPS = 4;

__global__ void __launch_bounds__(VAL_WARPS * 32, 6) k_lsd_validate(Dims d, const LsdWord *__restrict__ lsdw, const LsdPix *__restrict__ pix,
                                                                const int *__restrict__ pixcount, const LsdCand *__restrict__ cand,
                                                                const uint2 *__restrict__ candlist, const int *__restrict__ flags,
                                                                LsdSeg *__restrict__ candseg, u8 *__restrict__ candok)
{
    const int lane = threadIdx.x & 31;
    const int total = min(flags[3], d.n * 3 * d.segcap);
    const int nwarps = gridDim.x * VAL_WARPS;
    for (int t = blockIdx.x * VAL_WARPS + (threadIdx.x >> 5); t < total; t += nwarps) {
        uint2 e = candlist[t];
        const int img = (int)e.x, ci = (int)e.y;
        Img im;
        setup_img(im, d, img, lsdw, const_cast<LsdPix *>(pix), nullptr, nullptr, pixcount);
        const size_t o = (size_t)img * d.segcap + ci;
        LsdCand c = cand[o];
        Rect rec;
        rec.x1 = c.x1; rec.y1 = c.y1; rec.x2 = c.x2; rec.y2 = c.y2; rec.width = c.width; rec.theta = c.theta;
        rec.dx = c.dx; rec.dy = c.dy; rec.x = 0; rec.y = 0;
        rec.prec = kPI * 22.5 / 180.0; rec.p = 22.5 / 180.0;
        double log_nfa = rect_improve(im, rec);
        if (d.debug && lane == 0)
            printf("img %d cand %d nfa=%.17g p=%g w=%g (%.3f,%.3f)-(%.3f,%.3f)\n", img, ci, log_nfa, rec.p, rec.width, rec.x1, rec.y1,
                   rec.x2, rec.y2);
        if (lane == 0) {
            LsdSeg sg;
            sg.x1 = (float)((rec.x1 + 0.5) / 0.8); sg.y1 = (float)((rec.y1 + 0.5) / 0.8);
            sg.x2 = (float)((rec.x2 + 0.5) / 0.8); sg.y2 = (float)((rec.y2 + 0.5) / 0.8);
            candseg[o] = sg;
            candok[o] = log_nfa > 0.0 ? 1 : 0;
        }
    }
}

// ---- kernel 3: keep the validated candidates, in candidate (= acceptance) order ------------------------------------
__global__ void __launch_bounds__(128) k_lsd_emit(Dims d, const int *__restrict__ candcount, const LsdSeg *__restrict__ candseg,
                                                 const u8 *__restrict__ candok, LsdSeg *__restrict__ rawseg, int *__restrict__ segcount)
{
    const int lane = threadIdx.x & 31;
    const int img = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (img >= d.n * 3) return;
    const int nc = candcount[img];
    const size_t base = (size_t)img * d.segcap;
    int nout = 0;
    for (int i0 = 0; i0 < nc; i0 += 32) {
        int i = i0 + lane;
        bool ok = i < nc && candok[base + i];
        u32 m = __ballot_sync(FULL, ok);
        if (ok) rawseg[base + nout + __popc(m & ((1u << lane) - 1u))] = candseg[base + i];
        nout += __popc(m);
    }
    if (lane == 0) segcount[img] = nout;
}

void launch_lsd_core(const Dims &d, Buffers &b, cudaStream_t st)
{
    static bool tab_ready = false;
    static double *d_lgtab = nullptr;
    if (!tab_ready) {
        cudaMemcpyToSymbol(c_sincos_tab, scr::kSinCosTab, sizeof(scr::kSinCosTab));
        std::vector<double> tab(LGTAB);
        tab[0] = 0.0;
        for (int i = 1; i < LGTAB; ++i) tab[i] = log_gamma_d((double)i);
        cudaMalloc((void **)&d_lgtab, LGTAB * sizeof(double));
        cudaMemcpy(d_lgtab, tab.data(), LGTAB * sizeof(double), cudaMemcpyHostToDevice);
        cudaMemcpyToSymbol(g_lgtab, &d_lgtab, sizeof(d_lgtab));
        tab_ready = true;
    }
    k_lsd_nbr<<<d.n * 3, 256, 0, st>>>(d, b.lsdw, b.pix, b.pixcount, b.nbr);
    ++g_launches;
    k_lsd_grow<<<d.n * 3, 32, 0, st>>>(d, b.lsdw, b.pix, b.nbr, b.order, b.reg, b.pixcount, b.g2max, b.cand, b.candcount,
                                       b.candlist, b.flags);
    ++g_launches;
}

void launch_lsd_validate(const Dims &d, Buffers &b, cudaStream_t st)
{
    k_lsd_validate<<<148 * 6, VAL_WARPS * 32, 0, st>>>(d, b.lsdw, b.pix, b.pixcount, b.cand, b.candlist, b.flags, b.candseg, b.candok);
    ++g_launches;
    k_lsd_emit<<<(d.n * 3 + 3) / 4, 128, 0, st>>>(d, b.candcount, b.candseg, b.candok, b.rawseg, b.segcount);
    ++g_launches;
}

}  // namespace lsf
