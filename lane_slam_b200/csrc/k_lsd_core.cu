// k_lsd_core.cu -- K8+K9: the LSD search on one colour image: pseudo-ordered seeds (1024 bins,
// descending, raster order inside a bin), region growing, rectangle fit, density refinement,
// NFA validation with LSD_REFINE_ADV rectangle improvement, segment emission.
//
// Replaces the second half of cv2.LineSegmentDetector.detect (line_detector_lsd.py:64-72; SURVEY.md A.6;
// NFA math == src/line_descriptor/include/line_descriptor/descriptor_custom.hpp:676-826).
//
// The search inside one (frame, colour) image is order dependent (region growing commits pixels in FIFO
// order with the region angle updated after every accepted pixel; later seeds see the USED flags earlier
// regions left behind), so one warp owns one image and the cost is a chain of dependent steps.  The layout
// exists to make every step of that chain short:
//   k_lsd_index  (parallel, 256 threads per image): the seed order (stable counting sort) and one 160-byte
//                "fat" record per support pixel holding the compact index, angle, cos/sin and gradient of its
//                8 neighbours, so that consuming a queue entry needs ONE fetch instead of three dependent ones;
//   k_lsd_grow   (one warp per image): fat records are pulled into a shared-memory ring with cp.async at the
//                moment a pixel is accepted (and for upcoming seeds at batch load), USED flags live in a
//                shared-memory bitmap, four queue entries (32 neighbour visits) are resolved per round, and the
//                region angle is re-evaluated only when a visit is too close to the tolerance to decide
//                without it (bounded drift of the running sum) -- decisions are bit-identical to evaluating
//                it after every pixel;
//   k_lsd_validate (one warp per candidate): the NFA tests only read the image and run fully parallel.
// Rectangle sums are accumulated in region order through shared memory so every double rounding matches.
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
    const LsdPix *pix;      // thin records (angle, cos, sin, g2), raster order
    const u32 *pxy;         // (y << 16) | x of every support pixel
    const u32 *fat;         // [n][FAT_WORDS] neighbour records
    uint4 *reg;             // region list {idx, xy, g2, angle bits}; second half = scratch of reduce_region_radius
    u32 *used;              // USED bitmap (shared memory, or global when it does not fit)
    int W, H, swp, n, cap;
    double logNT;
};

__device__ __forceinline__ u32 smem_addr(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(void *smem, const void *g)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(smem)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

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

// USED bitmap.  SB: the bitmap is in shared memory (the address-space hint lets the compiler emit LDS / ATOMS)
template <bool SB>
__device__ __forceinline__ bool is_used(const Img &im, u32 idx)
{
    const u32 *p = &im.used[idx >> 5];
    if (SB) __builtin_assume(__isShared(p));
    return (*p >> (idx & 31)) & 1u;
}
template <bool SB>
__device__ __forceinline__ void set_used(const Img &im, u32 idx)
{
    u32 *p = &im.used[idx >> 5];
    if (SB) __builtin_assume(__isShared(p));
    atomicOr(p, 1u << (idx & 31));
}
template <bool SB>
__device__ __forceinline__ void clear_used(const Img &im, u32 idx)
{
    u32 *p = &im.used[idx >> 5];
    if (SB) __builtin_assume(__isShared(p));
    atomicAnd(p, ~(1u << (idx & 31)));
}

// the reference's isAligned distance: |theta - a|, folded once at 3*pi/2
__device__ __forceinline__ double ang_dist(double a, double theta)
{
    double n = fabs(theta - a);
    if (n > k3_2PI) n = fabs(n - k2PI);
    return n;
}
__device__ __forceinline__ bool aligned_ang(double a, double theta, double prec) { return ang_dist(a, theta) <= prec; }

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
// LSF_GROW_PROF=1 (env) -> d.debug & 2: per-image cycle / event counters of k_lsd_grow, printed by the host
struct Prof { long long t_total, t_grow, t_wait, t_rect, t_refine; int grows, rounds, accepts, refresh, cands; };
#define GPROF(stmt) do { stmt; } while (0)
constexpr int FAT_WORDS = 40;   // [idx x8][angle(deg) x8][cos x8][sin x8][g2 x8], neighbours row-major, centre skipped
constexpr int RING = 16;        // queue entries whose fat record can be resident at once
constexpr float kDEG2RADf = (float)kDEG2RAD, k3_2PIf = (float)k3_2PI, k2PIf = (float)k2PI;

struct Seq3 { double v[3][33]; };   // rows padded so that lanes 0..2 read different banks

struct GrowSm {
    __align__(16) u32 ring[RING][FAT_WORDS];      // fat records of queue positions q (slot q % RING)
    __align__(16) u32 seedrec[32][FAT_WORDS];     // fat records of the current batch of 32 seed candidates
    u32 qxy[RING];                                // xy of the queue entries in the ring
    u32 acc_idx[32];                              // pixels accepted by the current commit, in order
    float acc_c[32], acc_s[32];
    Seq3 seq;
};

// Grows from compact index `seed` with tolerance prec; fills im.reg[0..nreg) (acceptance order), marks
// pixels used, returns nreg and the final region angle.
//
// Round = up to four queue entries: lane (g, j) visits neighbour j (row major) of queue entry i+g, i.e. lane
// order == the reference's visiting order, and visits must be decided in that order: the reference
// recomputes reg_angle = fastAtan2(sumdy, sumdx) after every accepted pixel and tests the next visit against
// it.  A single warp running that chain is latency bound, so the chain is cut short:
//   * (theta0, |S0|) is a checkpoint of the exact angle.  After m more unit vectors were added the exact angle
//     differs from theta0 by at most B(m) = asin(m / |S0|) + 2 x (error of the fastAtan2 polynomial, 1.7e-4 rad
//     measured).  For every visit an upper bound m_ub of m at its turn is known (accepted so far + pending
//     visits before it), so a visit whose distance to theta0 lies outside [prec - B(m_ub), prec + B(m_ub)] is
//     decided from theta0 alone -- for all lanes at once, in float (the slack covers the float rounding).
//   * the first visit inside that band forces an exact step: new checkpoint at the current sums (if pixels
//     were accepted since the last one), the reference's double-precision test for every pending visit, and
//     the first aligned one is accepted.
// Both kinds of step give the decisions the reference makes.  Accepted pixels are committed in bulk: USED bits,
// region list entries, ring slots + cp.async of their fat records, and the float sums added in visit order.
template <bool SB>
__device__ __forceinline__ void commit(const Img &im, GrowSm &sm, u32 acc, int idx, u32 xy, u32 g2, float deg, float c, float s,
                                       int i, int &nreg, int &pf, float &sumdx, float &sumdy)
{
    const int lane = threadIdx.x & 31;
    const int cnt = __popc(acc), rank = __popc(acc & ((1u << lane) - 1u));
    const int room = pf == nreg ? max(0, min(cnt, i + RING - nreg)) : 0;   // positions that get a ring slot now
    __syncwarp();
    if ((acc >> lane) & 1u) {
        set_used<SB>(im, (u32)idx);
        im.reg[nreg + rank] = make_uint4((u32)idx, xy, g2, __float_as_uint(deg));
        sm.acc_idx[rank] = (u32)idx; sm.acc_c[rank] = c; sm.acc_s[rank] = s;
        if (rank < room) sm.qxy[(nreg + rank) % RING] = xy;
    }
    __syncwarp();
    for (int t = lane; t < room * (FAT_WORDS / 4); t += 32) {
        const int r = t / (FAT_WORDS / 4), part = t - r * (FAT_WORDS / 4);
        cp_async16(&sm.ring[(nreg + r) % RING][part * 4], im.fat + (size_t)sm.acc_idx[r] * FAT_WORDS + part * 4);
    }
    pf += room;
    for (int r = 0; r < cnt; ++r) {
        sumdx = __fadd_rn(sumdx, sm.acc_c[r]);
        sumdy = __fadd_rn(sumdy, sm.acc_s[r]);
    }
    nreg += cnt;
}

template <bool SB>
__device__ int grow(const Img &im, GrowSm &sm, Prof &pr, int seed, float seed_deg, u32 seed_g2, u32 seed_xy, float seed_c, float seed_s,
                    const u32 *seed_rec, double prec, double &reg_angle_out)
{
    const long long t_in = clock64();
    GPROF(++pr.grows);
    const int lane = threadIdx.x & 31;
    const u32 lt = (1u << lane) - 1u;
    const int g = lane >> 3, j = lane & 7;
    const int jj = j < 4 ? j : j + 1;                       // 3x3 position of neighbour j
    const int ndx = jj % 3 - 1, ndy = jj / 3 - 1;
    const bool lazy = prec <= 0.6;
    const float precf = (float)prec;
    int nreg = 1;
    double reg_angle = (double)seed_deg * kDEG2RAD;         // exact region angle whenever m == 0
    float sumdx = seed_c, sumdy = seed_s;                   // (float)cos(reg_angle), (float)sin(reg_angle)
    if (lane == 0) {
        im.reg[0] = make_uint4((u32)seed, seed_xy, seed_g2, __float_as_uint(seed_deg));
        set_used<SB>(im, (u32)seed);
        sm.qxy[0] = seed_xy;
    }
    if (!seed_rec && lane < FAT_WORDS / 4) cp_async16(&sm.ring[0][lane * 4], im.fat + (size_t)seed * FAT_WORDS + lane * 4);
    int pf = 1;                        // queue positions < pf have their record in flight / resident
    float th0f = (float)reg_angle;     // checkpoint angle (float copy) ...
    int m = 0, mmax = 0;               // ... pixels accepted since then, and how many the drift bound tolerates
    float kB = 1.21f;                  // B(m) = m * kB + slack;  kB = 1.21 / |S0|  (asin(x) <= 1.11 x for x <= 0.69)
    for (int i = 0; i < nreg;) {
        const int navail = min(4, nreg - i);
        // (rare) queue entries beyond the ring at acceptance time: fetch them now
        while (pf < nreg && pf < i + RING) {
            uint4 e = im.reg[pf];
            if (lane < FAT_WORDS / 4) cp_async16(&sm.ring[pf % RING][lane * 4], im.fat + (size_t)e.x * FAT_WORDS + lane * 4);
            if (lane == 0) sm.qxy[pf % RING] = e.y;
            ++pf;
        }
        {
            const long long tw = clock64();
            cp_async_wait_all();
            __syncwarp();
            GPROF(pr.t_wait += clock64() - tw; ++pr.rounds);
        }
        int idx = -1;
        float deg = 0.f, c = 0.f, s = 0.f;
        u32 g2 = 0, xy = 0;
        if (g < navail) {
            const int pos = i + g;
            const u32 *rec = (pos == 0 && seed_rec) ? seed_rec : sm.ring[pos % RING];
            u32 ni = rec[j];
            if (ni != LSD_NONE && !is_used<SB>(im, ni)) {
                idx = (int)ni;
                deg = __uint_as_float(rec[8 + j]); c = __uint_as_float(rec[16 + j]); s = __uint_as_float(rec[24 + j]);
                g2 = rec[32 + j];
                u32 cxy = sm.qxy[pos % RING];
                xy = (u32)((int)cxy + (ndy << 16) + ndx);
            }
        }
        u32 R = __ballot_sync(FULL, idx >= 0);          // undecided visits
        if (!R) { i += navail; continue; }
        const u32 peers = __match_any_sync(FULL, idx);  // visits of the same pixel (from several queue entries)
        u32 done = 0;                                   // visits accepted in this round
        const float angf = deg * kDEG2RADf;
        float d0f = fabsf(th0f - angf);
        if (d0f > k3_2PIf) d0f = fabsf(d0f - k2PIf);
        while (R) {
            // ---- bulk step: decide every visit that is clear of the tolerance band ----
            const bool in = (R >> lane) & 1u;
            const int mub = m + __popc(R & lt);
            const float B = (float)mub * kB + 0.002f;
            const bool okb = in && lazy && mub <= mmax;
            const u32 ma = __ballot_sync(FULL, okb && d0f <= precf - B);
            const u32 mn = __ballot_sync(FULL, okb && d0f >= precf + B);
            const u32 unsure = R & ~(ma | mn);
            const u32 below = unsure ? ((1u << (__ffs(unsure) - 1)) - 1u) : FULL;
            u32 acc = ma & below;
            if (acc) {
                acc = __ballot_sync(FULL, ((acc >> lane) & 1u) && (peers & acc & lt) == 0);   // first visit of a pixel wins
                commit<SB>(im, sm, acc, idx, xy, g2, deg, c, s, i, nreg, pf, sumdx, sumdy);
                m += __popc(acc);
                done |= acc;
                GPROF(pr.accepts += __popc(acc));
            }
            R &= ~below;
            R &= ~__ballot_sync(FULL, (peers & done) != 0);   // later visits of pixels accepted just now: USED
            if (!(R & unsure)) continue;                      // the unsure visit was such a duplicate (or none was)
            // ---- exact step: the reference's test at the current angle, up to the first accepted visit ----
            if (m > 0) {
                reg_angle = (double)fast_atan2_deg(sumdy, sumdx) * kDEG2RAD;
                th0f = (float)reg_angle; m = 0;
                const float S0 = sqrtf(sumdx * sumdx + sumdy * sumdy);
                kB = 1.21f / S0; mmax = (int)(0.69f * S0);
                d0f = fabsf(th0f - angf);
                if (d0f > k3_2PIf) d0f = fabsf(d0f - k2PIf);
                GPROF(++pr.refresh);
            }
            const u32 al = __ballot_sync(FULL, idx >= 0 && aligned_ang((double)deg * kDEG2RAD, reg_angle, prec)) & R;
            if (!al) break;                                   // nothing else in this round is aligned
            const int k = __ffs(al) - 1;
            commit<SB>(im, sm, 1u << k, idx, xy, g2, deg, c, s, i, nreg, pf, sumdx, sumdy);
            m = 1;
            done |= 1u << k;
            GPROF(++pr.accepts);
            R &= ~((2u << k) - 1u);
            R &= ~__ballot_sync(FULL, (peers & done) != 0);
            if (!lazy) {   // wide tolerance (refine's tau): keep the angle exact after every pixel
                reg_angle = (double)fast_atan2_deg(sumdy, sumdx) * kDEG2RAD;
                th0f = (float)reg_angle; m = 0;
            }
        }
        __syncwarp();  // used bits / qxy of this round are visible to the next one
        i += navail;
    }
    if (m > 0) reg_angle = (double)fast_atan2_deg(sumdy, sumdx) * kDEG2RAD;
    reg_angle_out = reg_angle;
    GPROF(pr.t_grow += clock64() - t_in);
    return nreg;
}

// ---- reference-order accumulation --------------------------------------------------------------------
// The reference adds region points one by one in double precision; rounding (and through it theta, the
// rectangle corners and finally the NFA pixel counts) depends on that order.  Lanes compute the per-point
// terms in parallel and park them in shared memory; lanes 0, 1, 2 then each add one of the three series in
// point order (one dependent chain per lane instead of three per lane).
__device__ __forceinline__ void seq_add3(Seq3 &sm, int cnt, double ta, double tb, double tc, double &acc)
{
    const int lane = threadIdx.x & 31;
    __syncwarp();
    sm.v[0][lane] = ta; sm.v[1][lane] = tb; sm.v[2][lane] = tc;
    __syncwarp();
    const double *src = sm.v[lane < 3 ? lane : 0];
    for (int q = 0; q < cnt; ++q) acc += src[q];
}
#define SEQ_RESULT(acc, A, B, C) do { A = __shfl_sync(FULL, acc, 0); B = __shfl_sync(FULL, acc, 1); C = __shfl_sync(FULL, acc, 2); } while (0)

__device__ __forceinline__ double entry_w(const uint4 &e) { return sqrt((double)e.z / 4.0); }   // modgrad

// ---- rectangle fit ------------------------------------------------------------------------------------
__device__ void region2rect(const Img &im, Seq3 &sm, int nreg, double reg_angle, double prec, double p, Rect &r)
{
    const int lane = threadIdx.x & 31;
    const uint4 zero = make_uint4(0, 0, 0, 0);
    double x, y, sum, acc = 0;
    uint4 nx = lane < nreg ? im.reg[lane] : zero;
    for (int i0 = 0; i0 < nreg; i0 += 32) {
        const uint4 e = nx;
        const int i = i0 + lane;
        if (i + 32 < nreg) nx = im.reg[i + 32];
        double tx = 0, ty = 0, tw = 0;
        if (i < nreg) {
            tw = entry_w(e);
            tx = (double)(e.y & 0xffffu) * tw;
            ty = (double)(e.y >> 16) * tw;
        }
        seq_add3(sm, min(32, nreg - i0), tx, ty, tw, acc);
    }
    SEQ_RESULT(acc, x, y, sum);
    x /= sum; y /= sum;
    double Ixx, Iyy, Ixy;
    acc = 0;
    nx = lane < nreg ? im.reg[lane] : zero;
    for (int i0 = 0; i0 < nreg; i0 += 32) {
        const uint4 e = nx;
        const int i = i0 + lane;
        if (i + 32 < nreg) nx = im.reg[i + 32];
        double ta = 0, tb = 0, tc = 0;
        if (i < nreg) {
            double w = entry_w(e);
            double dx = (double)(e.y & 0xffffu) - x, dy = (double)(e.y >> 16) - y;
            ta = dy * dy * w; tb = dx * dx * w; tc = -(dx * dy * w);   // Ixy -= dx*dy*w
        }
        seq_add3(sm, min(32, nreg - i0), ta, tb, tc, acc);
    }
    SEQ_RESULT(acc, Ixx, Iyy, Ixy);
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
        u32 xy = im.reg[i].y;
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
template <bool SB>
__device__ bool refine(const Img &im, GrowSm &gsm, Prof &pr, int &nreg, double &reg_angle, double prec, double p, Rect &rec,
                       float seed_c, float seed_s, const u32 *seed_rec)
{
    const double density_th = 0.7;
    const int lane = threadIdx.x & 31;
    Seq3 &sm = gsm.seq;
    if (rect_density(nreg, rec) >= density_th) return true;
    const uint4 se = im.reg[0];
    const int seed = (int)se.x;
    const double xc = (double)(se.y & 0xffffu), yc = (double)(se.y >> 16);
    const double ang_c = (double)__uint_as_float(se.w) * kDEG2RAD;
    double sum, s_sum, cnt, acc = 0;
    for (int i0 = 0; i0 < nreg; i0 += 32) {
        int i = i0 + lane;
        double ta = 0, tb = 0, tc = 0;
        if (i < nreg) {
            uint4 e = im.reg[i];
            clear_used<SB>(im, e.x);
            double ex = (double)(e.y & 0xffffu) - xc, ey = (double)(e.y >> 16) - yc;
            if (sqrt(ex * ex + ey * ey) < rec.width) {
                double ad = angle_diff_signed((double)__uint_as_float(e.w) * kDEG2RAD, ang_c);
                ta = ad; tb = ad * ad; tc = 1.0;
            }
        }
        seq_add3(sm, min(32, nreg - i0), ta, tb, tc, acc);
    }
    SEQ_RESULT(acc, sum, s_sum, cnt);
    __syncwarp();
    double mean = sum / cnt;
    double tau = 2.0 * sqrt((s_sum - 2.0 * mean * sum) / cnt + mean * mean);
    nreg = grow<SB>(im, gsm, pr, seed, __uint_as_float(se.w), se.z, se.y, seed_c, seed_s, seed_rec, tau, reg_angle);
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
            uint4 *scratch = im.reg + im.cap;
            int K = 0;
            for (int i0 = 0; i0 < nreg; i0 += 32) {
                int i = i0 + lane;
                bool keep = false;
                if (i < nreg) {
                    uint4 e = im.reg[i];
                    double ex = xc - (double)(e.y & 0xffffu), ey = yc - (double)(e.y >> 16);
                    keep = !(ex * ex + ey * ey > radSq);
                    if (!keep) clear_used<SB>(im, e.x);
                }
                K += __popc(__ballot_sync(FULL, keep));
            }
            // fillers: kept points at positions >= K, ranked from the back
            int before = 0;   // kept points in [0, i0)
            for (int i0 = 0; i0 < nreg; i0 += 32) {
                int i = i0 + lane;
                uint4 e = make_uint4(0, 0, 0, 0);
                bool keep = false;
                if (i < nreg) {
                    e = im.reg[i];
                    double ex = xc - (double)(e.y & 0xffffu), ey = yc - (double)(e.y >> 16);
                    keep = !(ex * ex + ey * ey > radSq);
                }
                u32 mk = __ballot_sync(FULL, keep);
                int incl = before + __popc(mk & ((2u << lane) - 1u));
                if (keep && i >= K) scratch[K - incl] = e;          // kept points after i = K - incl
                before += __popc(mk);
            }
            __syncwarp();
            before = 0;
            for (int i0 = 0; i0 < K; i0 += 32) {
                int i = i0 + lane;
                bool keep = true;
                if (i < K) {
                    u32 xy = im.reg[i].y;
                    double ex = xc - (double)(xy & 0xffffu), ey = yc - (double)(xy >> 16);
                    keep = !(ex * ex + ey * ey > radSq);
                }
                u32 mk = __ballot_sync(FULL, keep);
                int excl = before + __popc(mk & ((1u << lane) - 1u));  // kept points in [0, i)
                __syncwarp();
                if (i < K && !keep) im.reg[i] = scratch[i - excl];   // hole rank = holes before i
                before += __popc(mk);
            }
            __syncwarp();
            nreg = K;
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
                if (idx >= 0 && aligned_ang((double)im.pix[idx].deg * kDEG2RAD, r.theta, r.prec)) ++alg;
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

// ---- per-image view ---------------------------------------------------------------------------------------
__device__ __forceinline__ void setup_img(Img &im, const Dims &d, int img, const LsdWord *lsdw, const LsdPix *pix, const u32 *pxy,
                                          const u32 *fat, uint4 *reg, const int *pixcount)
{
    im.n = pixcount[img];
    im.W = d.sw; im.H = d.sh; im.swp = d.swp;
    im.words = lsdw + (size_t)img * d.sh * d.swp;
    im.pix = pix + (size_t)img * d.pixcap;
    im.pxy = pxy ? pxy + (size_t)img * d.pixcap : nullptr;
    im.fat = fat ? fat + (size_t)img * d.pixcap * FAT_WORDS : nullptr;
    im.reg = reg ? reg + (size_t)img * 2 * d.pixcap : nullptr;   // second half: scratch of reduce_region_radius
    im.used = nullptr;
    im.cap = d.pixcap;
    im.logNT = 5.0 * (log10((double)im.W) + log10((double)im.H)) / 2.0 + log10(11.0);
}

// ---- kernel 0: seed order + fat neighbour records (fully parallel, one CTA per image) ---------------------------
// warp 0: stable counting sort of the support pixels by bin = int(norm * 1023 / max_norm), descending (raster
//         order inside a bin) -> order[];  warps 1..7: one fat record per support pixel, 8 lanes per pixel.
__global__ void __launch_bounds__(256) k_lsd_index(Dims d, const LsdWord *__restrict__ lsdw, const LsdPix *__restrict__ pix,
                                                  const u32 *__restrict__ pxy, const int *__restrict__ pixcount,
                                                  const u32 *__restrict__ g2max, u32 *__restrict__ fat, u32 *__restrict__ order,
                                                  float2 *__restrict__ scs)
{
    __shared__ u32 hist[1024];
    const int img = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    Img im;
    setup_img(im, d, img, lsdw, pix, pxy, nullptr, nullptr, pixcount);
    const int n = im.n;
    if (n == 0) return;
    if (warp == 0) {
        u32 *ord = order + (size_t)img * d.pixcap;
        const double max_grad = sqrt((double)g2max[img] / 4.0);
        const double bin_coef = max_grad > 0 ? 1023.0 / max_grad : 0.0;
        for (int i = lane; i < 1024; i += 32) hist[i] = 0;
        __syncwarp();
        for (int i = lane; i < n; i += 32) {
            int bin = (int)(sqrt((double)im.pix[i].g2 / 4.0) * bin_coef);
            atomicAdd(&hist[1023 - bin], 1u);
        }
        __syncwarp();
        {
            // exclusive scan of 1024 counters: lane owns 32 consecutive entries
            u32 loc = 0;
            for (int q = 0; q < 32; ++q) loc += hist[lane * 32 + q];
            u32 incl = loc;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                u32 v = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += v;
            }
            u32 run = incl - loc;
            for (int q = 0; q < 32; ++q) { u32 c = hist[lane * 32 + q]; hist[lane * 32 + q] = run; run += c; }
        }
        __syncwarp();
        for (int i0 = 0; i0 < n; i0 += 32) {
            int i = i0 + lane;
            bool valid = i < n;
            u32 key = valid ? (u32)(1023 - (int)(sqrt((double)im.pix[i].g2 / 4.0) * bin_coef)) : (2048u + (u32)lane);
            u32 mm = __match_any_sync(FULL, key);
            int leader = __ffs(mm) - 1;
            u32 off = 0;
            if (valid && lane == leader) { off = hist[key]; hist[key] = off + __popc(mm); }
            off = __shfl_sync(FULL, off, leader);
            if (valid) ord[off + __popc(mm & ((1u << lane) - 1u))] = (u32)i;
            __syncwarp();
        }
    } else {
        const int j = lane & 7, jj = j < 4 ? j : j + 1;
        const int ndx = jj % 3 - 1, ndy = jj / 3 - 1;
        // what a region starts its sums with when this pixel is the seed: float(cos(angle)), float(sin(angle)) of the
        // DOUBLE angle (the per-pixel c/s are cosf/sinf of the float angle)
        float2 *ocs = scs + (size_t)img * d.pixcap;
        for (int p = threadIdx.x - 32; p < n; p += 224) {
            const double a = (double)im.pix[p].deg * kDEG2RAD;
            ocs[p] = make_float2((float)cos(a), (float)sin(a));
        }
        u32 *out = fat + (size_t)img * d.pixcap * FAT_WORDS;
        for (int p = (warp - 1) * 4 + (lane >> 3); p < n; p += 28) {
            u32 xy = im.pxy[p];
            int xx = (int)(xy & 0xffffu) + ndx, yy = (int)(xy >> 16) + ndy, r = -1;
            if (xx >= 0 && xx < im.W && yy >= 0 && yy < im.H) r = lookup(im, xx, yy);
            uint4 t = make_uint4(0, 0, 0, 0);
            if (r >= 0) t = *reinterpret_cast<const uint4 *>(&im.pix[r]);
            u32 *rec = out + (size_t)p * FAT_WORDS + j;
            rec[0] = r >= 0 ? (u32)r : LSD_NONE;
            rec[8] = t.x; rec[16] = t.y; rec[24] = t.z; rec[32] = t.w;
        }
    }
}

// ---- kernel 1: seeds, region growing, rectangle fit, density refinement -> candidate rectangles ----------
// One warp per (frame, colour) image; strictly sequential in the reference's seed order because growing and
// refining change which pixels later seeds may use.  NFA validation does not touch that state, so it is
// deferred to kernel 2, where every candidate gets its own warp.  Blocks are ordered colour-major (all white
// images first): white images hold the most support pixels and must not start last.
template <bool SB>
__global__ void __launch_bounds__(32) k_lsd_grow(Dims d, const LsdWord *__restrict__ lsdw, const LsdPix *__restrict__ pix,
                                                const u32 *__restrict__ pxy, const float2 *__restrict__ scs, const u32 *__restrict__ fat,
                                                const u32 *__restrict__ order, uint4 *__restrict__ reg,
                                                const int *__restrict__ pixcount, u32 *__restrict__ used_global,
                                                int used_words, LsdCand *__restrict__ cand, int *__restrict__ candcount,
                                                uint2 *__restrict__ candlist, int *__restrict__ flags, long long *__restrict__ prof)
{
    extern __shared__ __align__(16) u8 smraw[];
    Prof pr = {};
    const long long t_start = clock64();
    GrowSm &sm = *reinterpret_cast<GrowSm *>(smraw);
    const int lane = threadIdx.x;
    const int img = (blockIdx.x % d.n) * 3 + blockIdx.x / d.n;
    Img im;
    setup_img(im, d, img, lsdw, pix, pxy, fat, reg, pixcount);
    im.used = SB ? reinterpret_cast<u32 *>(smraw + sizeof(GrowSm)) : used_global + (size_t)img * used_words;
    const float2 *seedcs = scs + (size_t)img * d.pixcap;
    const u32 *ord = order + (size_t)img * d.pixcap;
    const int n = im.n;
    if (n == 0) {
        if (lane == 0) candcount[img] = 0;
        return;
    }
    for (int i = lane; i < (n + 31) / 32; i += 32) im.used[i] = 0;
    __syncwarp();

    const double prec = kPI * 22.5 / 180.0, p = 22.5 / 180.0;
    const int min_reg = (int)(-im.logNT / log10(p));
    int ncand = 0;
    LsdCand *out = cand + (size_t)img * d.segcap;
    int ci_next = lane < n ? (int)ord[lane] : -1;
    for (int o0 = 0; o0 < n; o0 += 32) {
        const int ci = ci_next;
        ci_next = o0 + 32 + lane < n ? (int)ord[o0 + 32 + lane] : -1;
        // this batch of 32 seed candidates: thin record, position, and (for the still unused ones) the fat record
        uint4 tp = make_uint4(0, 0, 0, 0);
        u32 cxy = 0;
        float2 ccs = make_float2(0.f, 0.f);
        bool want = ci >= 0 && !is_used<SB>(im, (u32)ci);
        if (want) {
            tp = *reinterpret_cast<const uint4 *>(&im.pix[ci]);
            cxy = im.pxy[ci];
            ccs = seedcs[ci];
            const u32 *src = im.fat + (size_t)ci * FAT_WORDS;
#pragma unroll
            for (int t = 0; t < FAT_WORDS / 4; ++t) cp_async16(&sm.seedrec[lane][t * 4], src + t * 4);
        }
        const u32 pfmask = __ballot_sync(FULL, want);
        int k = -1;
        while (true) {
            // seeds of this batch not yet visited and still unused *now*
            __syncwarp();
            u32 cnd = __ballot_sync(FULL, ci >= 0 && lane > k && !is_used<SB>(im, (u32)ci));
            if (!cnd) break;
            k = __ffs(cnd) - 1;
            const int seed = __shfl_sync(FULL, ci, k);
            const bool have = (pfmask >> k) & 1u;   // a seed released by an earlier refine was not pre-fetched
            u32 sdeg = __shfl_sync(FULL, tp.x, k), sg2 = __shfl_sync(FULL, tp.w, k), sxy = __shfl_sync(FULL, cxy, k);
            float sc = __shfl_sync(FULL, ccs.x, k), ss = __shfl_sync(FULL, ccs.y, k);
            if (!have) {
                sdeg = __float_as_uint(im.pix[seed].deg); sg2 = im.pix[seed].g2; sxy = im.pxy[seed];
                float2 t = seedcs[seed];
                sc = t.x; ss = t.y;
            }
            const u32 *srec = have ? sm.seedrec[k] : nullptr;
            double reg_angle;
            int nreg = grow<SB>(im, sm, pr, seed, __uint_as_float(sdeg), sg2, sxy, sc, ss, srec, prec, reg_angle);
            if (nreg < min_reg) continue;
            Rect rec;
            long long t0 = clock64();
            region2rect(im, sm.seq, nreg, reg_angle, prec, p, rec);
            long long t1 = clock64();
            bool okr = refine<SB>(im, sm, pr, nreg, reg_angle, prec, p, rec, sc, ss, srec);
            GPROF(pr.t_rect += t1 - t0; pr.t_refine += clock64() - t1; ++pr.cands);
            if (!okr) continue;
            if (ncand < d.segcap && lane == 0) {
                LsdCand c;
                c.x1 = rec.x1; c.y1 = rec.y1; c.x2 = rec.x2; c.y2 = rec.y2;
                c.width = rec.width; c.theta = rec.theta; c.dx = rec.dx; c.dy = rec.dy;
                out[ncand] = c;
                int slot = atomicAdd(&flags[3], 1);
                candlist[slot] = make_uint2((u32)(img + d.img0), (u32)ncand);
            }
            ++ncand;
        }
    }
    if (lane == 0) {
        if (ncand > d.segcap) { atomicMax(&flags[1], ncand); ncand = d.segcap; }
        candcount[img] = ncand;
        if (prof) {
            long long *o = prof + (size_t)img * 12;
            o[0] = clock64() - t_start; o[1] = pr.t_grow; o[2] = pr.t_wait; o[3] = pr.t_rect; o[4] = pr.t_refine;
            o[5] = pr.grows; o[6] = pr.rounds; o[7] = pr.accepts; o[8] = pr.refresh; o[9] = pr.cands; o[10] = n; o[11] = ncand;
        }
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
        setup_img(im, d, img, lsdw, pix, nullptr, nullptr, nullptr, pixcount);
        const size_t o = (size_t)img * d.segcap + ci;
        LsdCand c = cand[o];
        Rect rec;
        rec.x1 = c.x1; rec.y1 = c.y1; rec.x2 = c.x2; rec.y2 = c.y2; rec.width = c.width; rec.theta = c.theta;
        rec.dx = c.dx; rec.dy = c.dy; rec.x = 0; rec.y = 0;
        rec.prec = kPI * 22.5 / 180.0; rec.p = 22.5 / 180.0;
        double log_nfa = rect_improve(im, rec);
        if ((d.debug & 1) && lane == 0)
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
    k_lsd_index<<<d.n * 3, 256, 0, st>>>(d, b.lsdw, b.pix, b.pxy, b.pixcount, b.g2max, b.fat, b.order, b.scs);
    ++g_launches;
    // USED bitmap: shared memory when it fits (the normal case), the global fallback buffer otherwise
    const int used_words = (d.pixcap + 31) / 32;
    size_t smem = sizeof(GrowSm) + (size_t)used_words * 4;
    u32 *used_global = nullptr;
    if (smem > 200 * 1024) { smem = sizeof(GrowSm); used_global = b.usedbits; }
    static size_t attr = 0;
    if (!used_global && smem > attr) {
        cudaFuncSetAttribute(k_lsd_grow<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = smem;
    }
    long long *prof = nullptr;
    if (d.debug & 2) cudaMalloc((void **)&prof, (size_t)d.n * 3 * 12 * sizeof(long long));
    if (used_global)
        k_lsd_grow<false><<<d.n * 3, 32, smem, st>>>(d, b.lsdw, b.pix, b.pxy, b.scs, b.fat, b.order, b.reg, b.pixcount, used_global,
                                                     used_words, b.cand, b.candcount, b.candlist, b.flags, prof);
    else
        k_lsd_grow<true><<<d.n * 3, 32, smem, st>>>(d, b.lsdw, b.pix, b.pxy, b.scs, b.fat, b.order, b.reg, b.pixcount, used_global,
                                                    used_words, b.cand, b.candcount, b.candlist, b.flags, prof);
    ++g_launches;
    if (prof) {
        // developer aid (LSF_TRACE_LSD=2): where the slowest image and the average image spend their cycles
        std::vector<long long> h((size_t)d.n * 3 * 12);
        cudaStreamSynchronize(st);
        cudaMemcpy(h.data(), prof, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
        cudaFree(prof);
        int worst = 0;
        double avg[12] = {0};
        for (int i = 0; i < d.n * 3; ++i) {
            if (h[(size_t)i * 12] > h[(size_t)worst * 12]) worst = i;
            for (int q = 0; q < 12; ++q) avg[q] += (double)h[(size_t)i * 12 + q] / (d.n * 3);
        }
        const char *nm[12] = {"total", "grow", "wait", "rect", "refine", "grows", "rounds", "accepts", "refresh", "cands", "npix", "ncand"};
        fprintf(stderr, "[lsf grow prof] worst image %d:", worst);
        for (int q = 0; q < 12; ++q) fprintf(stderr, " %s=%lld", nm[q], h[(size_t)worst * 12 + q]);
        fprintf(stderr, "\n[lsf grow prof] mean:");
        for (int q = 0; q < 12; ++q) fprintf(stderr, " %s=%.0f", nm[q], avg[q]);
        fprintf(stderr, "\n");
    }
}

void launch_lsd_validate(const Dims &d, Buffers &b, cudaStream_t st)
{
    k_lsd_validate<<<148 * 6, VAL_WARPS * 32, 0, st>>>(d, b.lsdw, b.pix, b.pixcount, b.cand, b.candlist, b.flags, b.candseg, b.candok);
    ++g_launches;
    k_lsd_emit<<<(d.n * 3 + 3) / 4, 128, 0, st>>>(d, b.candcount, b.candseg, b.candok, b.rawseg, b.segcount);
    ++g_launches;
}

}  // namespace lsf
