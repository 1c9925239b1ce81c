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
//   k_lsd_index  (parallel, 256 threads per image): the seed order (stable counting sort), one 160-byte
//                "fat" record per support pixel holding the compact index, angle, cos/sin and gradient of its
//                8 neighbours, so that consuming a queue entry needs ONE fetch instead of three dependent ones,
//                and the 8-connected components of the support pixels: region growing never leaves a component,
//                so every component (grouped into at most 1024 tasks per image) is searched independently;
//   k_lsd_grow   (persistent warps pulling tasks): fat records are pulled into a shared-memory ring with cp.async at the
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
__device__ double g_lgtab[LGTAB];

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
#ifdef LSF_GROW_PROF
#define GPROF(...) __VA_ARGS__
#else
#define GPROF(...)
#endif
constexpr int FAT_WORDS = LSD_FAT_WORDS;   // 32 words = one 128-byte line: [idx x8][angle(deg) x8][cos x8][sin x8], neighbours row-major,
                                           // centre skipped (g2 of accepted pixels is fetched in bulk once a region is worth a rectangle)
constexpr int RING = 16;        // queue entries whose fat record can be resident at once
constexpr int SEED_SLOTS = 16;  // seed candidates of a batch of 32 whose fat record is fetched ahead (most are USED already)
constexpr float kDEG2RADf = (float)kDEG2RAD, k3_2PIf = (float)k3_2PI, k2PIf = (float)k2PI;

struct Seq3 { double v[3][33]; };   // rows padded so that lanes 0..2 read different banks

struct GrowSm {
    __align__(16) u32 ring[RING][FAT_WORDS];      // fat records of queue positions q (slot q % RING)
    __align__(16) u32 seedrec[SEED_SLOTS][FAT_WORDS];   // fat records of the still unused seed candidates of the current batch
    u32 qxy[RING];                                // xy of the queue entries in the ring
    float2 acc_cs[32];                            // cos / sin of the pixels accepted by the current commit, in order
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
    const int room = pf == nreg ? min(cnt, i + RING - nreg) : 0;   // positions that get a ring slot now (may be <= 0)
    __syncwarp();
    if ((acc >> lane) & 1u) {
        set_used<SB>(im, (u32)idx);
        im.reg[nreg + rank] = make_uint4((u32)idx, xy, g2, __float_as_uint(deg));
        sm.acc_cs[rank] = make_float2(c, s);
        if (rank < room) {
            const u32 slot = (u32)(nreg + rank) & (RING - 1);
            sm.qxy[slot] = xy;
            const u32 *src = im.fat + (size_t)idx * FAT_WORDS;
            u32 *dst = sm.ring[slot];
#pragma unroll
            for (int t = 0; t < FAT_WORDS / 4; ++t) cp_async16(dst + t * 4, src + t * 4);
        }
    }
    __syncwarp();
    if (room > 0) pf += room;
    for (int r = 0; r < cnt; ++r) {
        const float2 cs = sm.acc_cs[r];
        sumdx = __fadd_rn(sumdx, cs.x);
        sumdy = __fadd_rn(sumdy, cs.y);
    }
    nreg += cnt;
}

template <bool SB>
__device__ int grow(const Img &im, GrowSm &sm, Prof &pr, int seed, float seed_deg, u32 seed_g2, u32 seed_xy, float seed_c, float seed_s,
                    const u32 *seed_rec, double prec, double &reg_angle_out)
{
    GPROF(const long long t_in = clock64(); ++pr.grows);
    const int lane = threadIdx.x & 31;
    const u32 lt = (1u << lane) - 1u;
    const int g = lane >> 3, j = lane & 7;
    const int jj = j < 4 ? j : j + 1;                       // 3x3 position of neighbour j
    const int nd = ((jj / 3 - 1) << 16) + (jj % 3 - 1);     // xy offset of neighbour j
    const bool lazy = prec <= 0.6;
    const float precf = (float)prec;
    // drift per accepted pixel: it was aligned with the angle of its time, so it sits within prec + Bmax of the
    // checkpoint; its component perpendicular to S0 is at most sin(prec + Bmax), the parallel one is positive
    constexpr float Bmax = 0.2f;
    const float kfac = lazy ? 1.01f * sinf(precf + Bmax) : 0.f;
    int nreg = 1;
    double reg_angle = (double)seed_deg * kDEG2RAD;         // exact region angle whenever m == 0
    float sumdx = seed_c, sumdy = seed_s;                   // (float)cos(reg_angle), (float)sin(reg_angle)
    if (lane == 0) {
        im.reg[0] = make_uint4((u32)seed, seed_xy, seed_g2, __float_as_uint(seed_deg));
        set_used<SB>(im, (u32)seed);
        sm.qxy[0] = seed_xy;
    }
    if (seed_rec) {
        cp_async_wait_all();
        __syncwarp();
        if (lane < FAT_WORDS / 4) reinterpret_cast<uint4 *>(sm.ring[0])[lane] = reinterpret_cast<const uint4 *>(seed_rec)[lane];
    } else if (lane < FAT_WORDS / 4) {
        cp_async16(&sm.ring[0][lane * 4], im.fat + (size_t)seed * FAT_WORDS + lane * 4);
    }
    int pf = 1;                        // queue positions < pf have their record in flight / resident
    float th0f = (float)reg_angle;     // checkpoint angle (float copy) ...
    int m = 0;                         // ... pixels accepted since then
    float kB = kfac;                   // B(m) = m * kB + slack, kB = sin(prec + Bmax) / |S0|; |S0| = 1 at the seed
    for (int i = 0; i < nreg;) {
        const int navail = min(4, nreg - i);
        // (rare) queue entries beyond the ring at acceptance time: fetch them now
        while (pf < nreg && pf < i + RING) {
            uint4 e = im.reg[pf];
            if (lane < FAT_WORDS / 4) cp_async16(&sm.ring[pf & (RING - 1)][lane * 4], im.fat + (size_t)e.x * FAT_WORDS + lane * 4);
            if (lane == 0) sm.qxy[pf & (RING - 1)] = e.y;
            ++pf;
        }
        {
            GPROF(const long long tw = clock64());
            cp_async_wait_all();
            __syncwarp();
            GPROF(pr.t_wait += clock64() - tw; ++pr.rounds);
        }
        int idx = -1;
        float deg = 0.f, c = 0.f, s = 0.f;
        u32 g2 = 0, xy = 0;
        if (g < navail) {
            const u32 slot = (u32)(i + g) & (RING - 1);
            const u32 *rec = sm.ring[slot];
            u32 ni = rec[j];
            if (ni != LSD_NONE && !is_used<SB>(im, ni)) {
                idx = (int)ni;
                deg = __uint_as_float(rec[8 + j]); c = __uint_as_float(rec[16 + j]); s = __uint_as_float(rec[24 + j]);
                xy = sm.qxy[slot] + (u32)nd;
                // a visit that may be accepted: start pulling its fat record towards L2 now (the queue is short, the
                // cp.async issued at acceptance would otherwise pay the full DRAM latency one round later)
                const u32 *nf = im.fat + (size_t)ni * FAT_WORDS;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nf));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nf + 32));
            }
        }
        u32 R = __ballot_sync(FULL, idx >= 0);          // undecided visits
        if (!R) { i += navail; continue; }
        const u32 peers = __match_any_sync(FULL, idx);  // visits of the same pixel (from several queue entries)
        const float angf = deg * kDEG2RADf;
        float d0f = fabsf(th0f - angf);
        if (d0f > k3_2PIf) d0f = fabsf(d0f - k2PIf);
        while (R) {
            // ---- bulk step: decide every visit that is clear of the tolerance band ----
            if (lazy) {
                const bool in = (R >> lane) & 1u;
                // pass 1: loose bound (every pending visit before this one may be accepted) -> certainly rejected visits
                float B = (float)(m + __popc(R & lt)) * kB + 0.002f;
                const u32 mn1 = __ballot_sync(FULL, in && B <= Bmax && d0f >= precf + B);
                // pass 2: those cannot add to the drift
                B = (float)(m + __popc(R & ~mn1 & lt)) * kB + 0.002f;
                const bool okb = in && B <= Bmax;
                const u32 ma = __ballot_sync(FULL, okb && d0f <= precf - B);
                const u32 mn = __ballot_sync(FULL, okb && d0f >= precf + B) | mn1;
                const u32 unsure = R & ~(ma | mn);
                const u32 below = unsure ? ((unsure & (0u - unsure)) - 1u) : FULL;   // visits before the first unsure one
                u32 acc = ma & below;
                R &= ~below;
                if (acc) {
                    acc = __ballot_sync(FULL, ((acc >> lane) & 1u) && (peers & acc & lt) == 0);   // first visit of a pixel wins
                    commit<SB>(im, sm, acc, idx, xy, g2, deg, c, s, i, nreg, pf, sumdx, sumdy);
                    m += __popc(acc);
                    GPROF(pr.accepts += __popc(acc));
                    R &= ~__ballot_sync(FULL, (peers & acc) != 0);   // later visits of pixels accepted just now: USED
                }
                if (!R) break;
            }
            // ---- exact step: the reference's test at the current angle, up to the first accepted visit ----
            if (m > 0) {
                reg_angle = (double)fast_atan2_deg(sumdy, sumdx) * kDEG2RAD;
                th0f = (float)reg_angle; m = 0;
                kB = kfac * rsqrtf(sumdx * sumdx + sumdy * sumdy);
                d0f = fabsf(th0f - angf);
                if (d0f > k3_2PIf) d0f = fabsf(d0f - k2PIf);
                GPROF(++pr.refresh);
                if (lazy) continue;      // most visits are decidable from the new checkpoint
            }
            const u32 al = __ballot_sync(FULL, idx >= 0 && aligned_ang((double)deg * kDEG2RAD, reg_angle, prec)) & R;
            if (!al) break;                                   // nothing else in this round is aligned
            const u32 kbit = al & (0u - al);
            commit<SB>(im, sm, kbit, idx, xy, g2, deg, c, s, i, nreg, pf, sumdx, sumdy);
            m = 1;
            GPROF(++pr.accepts);
            R &= ~(kbit | (kbit - 1u));
            R &= ~__ballot_sync(FULL, (peers & kbit) != 0);
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

// g2 (squared gradient norm, the weight of a region point) is not part of the neighbour records: once a region is large
// enough to be fitted, the lanes fetch it for all its points at once (independent gathers, off the growing chain)
__device__ __forceinline__ void fill_g2(const Img &im, int nreg)
{
    const int lane = threadIdx.x & 31;
    for (int i = lane; i < nreg; i += 32) {
        const u32 idx = im.reg[i].x;
        im.reg[i].z = im.pix[idx].g2;
    }
    __syncwarp();
}

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
    fill_g2(im, nreg);
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
    alg = (int)__reduce_add_sync(FULL, (unsigned)alg);
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
    im.logNT = d.logNT;
}

// ---- kernel 0: seed order, fat neighbour records, connected components (fully parallel, one CTA per image) -------
// Region growing never leaves an 8-connected component of the support pixels, USED flags included, so the search
// of an image splits exactly into independent searches per component (seeds visited in the image's seed order
// restricted to the component; candidates merged back by the seed's position in the image order).  Components
// smaller than min_reg cannot produce a candidate and are dropped here.
//   phase 1  warp 0: stable counting sort of the support pixels by bin = int(norm * 1023 / max_norm), descending
//            (raster order inside a bin) -> order[];  warps 1..7: fat record of every pixel (8 lanes per pixel),
//            union-find merge with the W / NW / N / NE neighbours, seed cos/sin
//   phase 2  flatten labels (root = smallest index of the component), component sizes
//   phase 3  component slots; slot s belongs to task s % MAXC (a task is a union of whole components, so any number of
//            components is fine); tasks appended to the work list (>= BIG_COMP pixels from the front, the others from
//            the back: long chains start first)
//   phase 4  stable partition of order[] by task -> corder[] (+ cpos[] = position in order[])
#ifndef LSF_GROW_PER_SM_BUILD
#define LSF_GROW_PER_SM_BUILD 16
#endif
constexpr int GROW_PER_SM = LSF_GROW_PER_SM_BUILD;  // resident growing warps per SM, measured: 12 -> 5.3 ms, 16 -> 4.7 ms, 20 -> 5.8 ms, 28 (72 regs, spills) -> 6.0 ms
constexpr int MAXC = LSD_MAXC;   // tasks per image: component slot s goes to task s % MAXC (a task = a union of whole components)
constexpr int BIG_COMP = 768;
constexpr int SL_MAX = 8192;     // support pixels whose union-find labels fit in shared memory

// labels only ever decrease and always point into the same component (path halving included)
__device__ __forceinline__ u32 uf_find(u32 *label, u32 x)
{
    volatile u32 *vl = label;
    u32 p = vl[x];
    while (p != x) {
        u32 gp = vl[p];
        if (gp != p) atomicMin(&label[x], gp);
        x = p; p = gp;
    }
    return x;
}
__device__ __forceinline__ void uf_union(u32 *label, u32 a, u32 b)
{
    while (true) {
        a = uf_find(label, a); b = uf_find(label, b);
        if (a == b) return;
        if (a < b) { u32 t = a; a = b; b = t; }
        u32 old = atomicMin(&label[a], b);     // root a (the larger index) goes under b
        if (old == a) return;
        a = old;                                // a was re-parented meanwhile: merge its new parent with b
    }
}

// stable counting-sort scatter of one warp's contiguous segment [lo, hi) of a sequence: key(pos) in [0, nkeys),
// cursor[key] = first output slot for this warp's elements with that key (shared memory, private to the warp).
// KEY(pos, valid&) and EMIT(pos, dst) are lambdas; elements with valid == false are skipped.
template <typename KeyFn, typename EmitFn>
__device__ __forceinline__ void warp_stable_scatter(int lo, int hi, u32 *cursor, KeyFn key_of, EmitFn emit)
{
    const int lane = threadIdx.x & 31;
    for (int p0 = lo; p0 < hi; p0 += 32) {
        const int pos = p0 + lane;
        bool ok = pos < hi;
        u32 key = 0;
        if (ok) key = key_of(pos, ok);
        const u32 mk = ok ? key : (0x80000000u + (u32)lane);
        const u32 mm = __match_any_sync(FULL, mk);
        const int leader = __ffs(mm) - 1;
        u32 base = 0;
        if (ok && lane == leader) { base = cursor[key]; cursor[key] = base + __popc(mm); }
        base = __shfl_sync(FULL, base, leader);
        if (ok) emit(pos, base + __popc(mm & ((1u << lane) - 1u)));
        __syncwarp();
    }
}

__global__ void __launch_bounds__(256) k_lsd_index(Dims d, const LsdWord *__restrict__ lsdw, const LsdPix *__restrict__ pix,
                                                  const u32 *__restrict__ pxy, const int *__restrict__ pixcount,
                                                  const u32 *__restrict__ g2max, u32 *__restrict__ fat, u32 *__restrict__ order,
                                                  float2 *__restrict__ scs, u32 *__restrict__ label_, u32 *__restrict__ csize_,
                                                  u32 *__restrict__ coff_, u32 *__restrict__ corder_, u32 *__restrict__ cpos_,
                                                  uint2 *__restrict__ tasks, uint2 *__restrict__ worklist, int worklist_cap,
                                                  int *__restrict__ taskctr, int *__restrict__ candcount)
{
    // 32 KB used twice: union-find forest of the image (phases 1-2, when it fits; else the global label array),
    // then per-warp key counts / cursors of the two stable partitions (bins, then component slots)
    __shared__ u32 sbuf[8 * 1024];
    u32 (*cnt)[1024] = reinterpret_cast<u32 (*)[1024]>(sbuf);
    u32 *s_label = sbuf;
    __shared__ u32 s_wsum[8];
    __shared__ u32 s_ncomp;
    const int img = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    Img im;
    setup_img(im, d, img, lsdw, pix, pxy, nullptr, nullptr, pixcount);
    const int n = im.n;
    if (tid == 0) candcount[img] = 0;
    if (n == 0) return;
    const size_t ibase = (size_t)img * d.pixcap;
    u32 *label = label_ + ibase, *csize = csize_ + ibase, *coff = coff_ + ibase, *corder = corder_ + ibase, *cpos = cpos_ + ibase;
    u32 *ord = order + ibase;
    const int min_reg = d.min_reg;
    const double max_grad = sqrt((double)g2max[img] / 4.0);
    const double bin_coef = max_grad > 0 ? 1023.0 / max_grad : 0.0;
    // every warp owns a contiguous segment (multiple of 32) of any n-long sequence
    const int seg = (((n + 7) / 8) + 31) & ~31;
    const int p_lo = min(n, warp * seg), p_hi = min(n, p_lo + seg);
    u32 *lab = n <= SL_MAX ? s_label : label;
    for (int i = tid; i < n; i += 256) { lab[i] = (u32)i; csize[i] = 0; }
    if (tid == 0) s_ncomp = 0;
    __syncthreads();
    // ---- phase 1: fat records + unions; seed cos/sin ----
    for (int i = p_lo + lane; i < p_hi; i += 32) {
        const LsdPix t = im.pix[i];
        // what a region starts its sums with when this pixel is the seed: float(cos(angle)), float(sin(angle)) of the
        // DOUBLE angle (the per-pixel c/s are cosf/sinf of the float angle)
        const double a = (double)t.deg * kDEG2RAD;
        scs[ibase + i] = make_float2((float)cos(a), (float)sin(a));
    }
    {
        const int j = lane & 7, jj = j < 4 ? j : j + 1;
        const int ndx = jj % 3 - 1, ndy = jj / 3 - 1;
        u32 *out = fat + ibase * FAT_WORDS;
        constexpr int U = 4;      // pixels per thread in flight (three dependent loads each)
        for (int p0 = 0; p0 < n; p0 += 32 * U) {
            int pp[U], r[U];
            u32 xy[U];
            uint4 t[U];
#pragma unroll
            for (int u = 0; u < U; ++u) { pp[u] = p0 + 32 * u + (tid >> 3); xy[u] = pp[u] < n ? im.pxy[pp[u]] : 0u; }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                r[u] = -1;
                if (pp[u] < n) {
                    int xx = (int)(xy[u] & 0xffffu) + ndx, yy = (int)(xy[u] >> 16) + ndy;
                    if (xx >= 0 && xx < im.W && yy >= 0 && yy < im.H) r[u] = lookup(im, xx, yy);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                t[u] = make_uint4(0, 0, 0, 0);
                if (r[u] >= 0) t[u] = *reinterpret_cast<const uint4 *>(&im.pix[r[u]]);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (pp[u] < n) {
                    u32 *rec = out + (size_t)pp[u] * FAT_WORDS + j;
                    rec[0] = r[u] >= 0 ? (u32)r[u] : LSD_NONE;
                    rec[8] = t[u].x; rec[16] = t[u].y; rec[24] = t[u].z;
                }
            }
            // merge with the already visited neighbours.  NW-N, N-NE and W-NW are neighbours of each other (merged at
            // their own turn), so one merge with N is enough when N exists, at most two otherwise.
            // (one ballot per pixel batch instead of four shuffles: lanes 0..3 of a group hold NW, N, NE, W and merge themselves;
            // concurrent unions on the same pixel are fine, the forest is lock-free)
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const u32 have = (__ballot_sync(FULL, r[u] >= 0) >> (lane & 24)) & 0xfu;    // bit 0 NW, 1 N, 2 NE, 3 W of this group
                bool mine;
                if (have & 2u) mine = j == 1;                                        // N exists: it connects all the others
                else mine = (j == 2 && (have & 4u)) || (j == 0 && (have & 1u)) || (j == 3 && (have & 8u) && !(have & 1u));
                if (mine && pp[u] < n) uf_union(lab, (u32)pp[u], (u32)r[u]);
            }
        }
    }
    __syncthreads();
    // ---- phase 2: flatten labels, component sizes; then bin counts of the warp's pixels and their cursors ----
    for (int i0 = 0; i0 < n; i0 += 256) {
        const int i = i0 + tid;
        u32 r = LSD_NONE - (u32)lane;           // lanes past the end: distinct dummy keys
        if (i < n) { r = uf_find(lab, (u32)i); label[i] = r; }
        // consecutive pixels mostly share their root: one atomic per distinct root in the warp
        const u32 mm = __match_any_sync(FULL, r);
        if (i < n && lane == __ffs(mm) - 1) atomicAdd(&csize[r], (u32)__popc(mm));
    }
    __syncthreads();
    for (int i = tid; i < 8 * 1024; i += 256) sbuf[i] = 0;
    __syncthreads();
    for (int i = p_lo + lane; i < p_hi; i += 32) {
        int bin = (int)(sqrt((double)im.pix[i].g2 / 4.0) * bin_coef);
        atomicAdd(&cnt[warp][1023 - bin], 1u);
    }
    __syncthreads();
    {
        // exclusive scan over (key, warp); thread owns keys 4*tid .. 4*tid+3
        u32 sum = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            for (int w = 0; w < 8; ++w) sum += cnt[w][4 * tid + q];
        u32 incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u32 v = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        u32 run = incl - sum;
        for (int k = 0; k < warp; ++k) run += s_wsum[k];
#pragma unroll
        for (int q = 0; q < 4; ++q)
            for (int w = 0; w < 8; ++w) { u32 c = cnt[w][4 * tid + q]; cnt[w][4 * tid + q] = run; run += c; }
    }
    __syncthreads();
    // ---- phase 3: seed order (stable scatter of the warp's pixels); component slots -> tasks ----
    warp_stable_scatter(p_lo, p_hi, cnt[warp],
                        [&](int pos, bool &) { return (u32)(1023 - (int)(sqrt((double)im.pix[pos].g2 / 4.0) * bin_coef)); },
                        [&](int pos, u32 dst) { ord[dst] = (u32)pos; });
    __syncthreads();
    u32 *tsize = &cnt[0][0];                     // [MAXC] pixels per task (the bin cursors are dead now)
    for (int i = tid; i < MAXC; i += 256) tsize[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += 256) {
        if (label[i] == (u32)i) {
            const u32 sz = csize[i];
            if (sz >= (u32)min_reg) {
                const u32 slot = atomicAdd(&s_ncomp, 1u);   // slots in any order: components are independent of each other
                coff[i] = slot & (MAXC - 1);                // task of this component
                atomicAdd(&tsize[slot & (MAXC - 1)], sz);
            }
        }
    }
    __syncthreads();
    const u32 ntask = min(s_ncomp, (u32)MAXC);
    uint2 *tk = tasks + (size_t)img * MAXC;
    {
        // exclusive scan of the task sizes (thread owns tasks 4*tid .. 4*tid+3) -> first seed of every task
        u32 v[4], sum = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) { v[q] = tsize[4 * tid + q]; sum += v[q]; }
        u32 incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u32 t = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        u32 run = incl - sum;
        for (int k = 0; k < warp; ++k) run += s_wsum[k];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (4 * tid + q < (int)ntask) tk[4 * tid + q] = make_uint2(run, v[q]);
            run += v[q];
        }
    }
    __syncthreads();
    for (int t = tid; t < (int)ntask; t += 256) {
        const bool big = tk[t].y >= (u32)BIG_COMP;
        const int slot = big ? atomicAdd(&taskctr[0], 1) : worklist_cap - 1 - atomicAdd(&taskctr[1], 1);
        worklist[slot] = make_uint2((u32)img, (u32)t);
    }
    for (int i = tid; i < 8 * 1024; i += 256) sbuf[i] = 0;
    // ---- phase 4: seed list of every task, in image seed order (stable partition of order[]) ----
    // label[i] <- task of pixel i (LSD_NONE for the dropped small components)
    for (int i = tid; i < n; i += 256) {
        const u32 r = label[i];
        label[i] = csize[r] >= (u32)min_reg ? coff[r] : LSD_NONE;
    }
    __syncthreads();
    for (int pos = p_lo + lane; pos < p_hi; pos += 32) {
        const u32 sl = label[ord[pos]];
        if (sl != LSD_NONE) atomicAdd(&cnt[warp][sl], 1u);
    }
    __syncthreads();
    for (int t = tid; t < (int)ntask; t += 256) {
        u32 run = tk[t].x;
        for (int w = 0; w < 8; ++w) { const u32 c = cnt[w][t]; cnt[w][t] = run; run += c; }
    }
    __syncthreads();
    warp_stable_scatter(p_lo, p_hi, cnt[warp],
                        [&](int pos, bool &ok) { const u32 sl = label[ord[pos]]; ok = sl != LSD_NONE; return sl; },
                        [&](int pos, u32 dst) { corder[dst] = ord[pos]; cpos[dst] = (u32)pos; });
}

// ---- kernel 1: seeds, region growing, rectangle fit, density refinement -> candidate rectangles ----------
// Persistent single-warp blocks pull (image, component) tasks from the work list.  Inside a task everything is
// strictly sequential in the reference's seed order, because growing and refining change which pixels later
// seeds may use.  NFA validation does not touch that state, so it is deferred to kernel 2, where every
// candidate gets its own warp.
template <bool SB>
__global__ void __launch_bounds__(32, GROW_PER_SM) k_lsd_grow(Dims d, const LsdWord *__restrict__ lsdw, const LsdPix *__restrict__ pix,
                                                const u32 *__restrict__ pxy, const float2 *__restrict__ scs, const u32 *__restrict__ fat,
                                                const u32 *__restrict__ corder_, const u32 *__restrict__ cpos_,
                                                const uint2 *__restrict__ tasks, const uint2 *__restrict__ worklist, int worklist_cap,
                                                int *__restrict__ taskctr, uint4 *__restrict__ reg,
                                                const int *__restrict__ pixcount, u32 *__restrict__ used_global,
                                                int used_words, LsdCand *__restrict__ cand, u32 *__restrict__ candrank,
                                                int *__restrict__ candcount, uint2 *__restrict__ candlist, int *__restrict__ flags,
                                                long long *__restrict__ prof)
{
    extern __shared__ __align__(16) u8 smraw[];
    GrowSm &sm = *reinterpret_cast<GrowSm *>(smraw);
    const int lane = threadIdx.x;
    const double prec = kPI * 22.5 / 180.0, p = 22.5 / 180.0;
    while (true) {
        Prof pr = {};
        const long long t_start = clock64();
        int t = 0;
        if (lane == 0) t = atomicAdd(&taskctr[2], 1);
        t = __shfl_sync(FULL, t, 0);
        const int nbig = taskctr[0], nsmall = taskctr[1];
        if (t >= nbig + nsmall) break;
        const uint2 wl = worklist[t < nbig ? t : worklist_cap - 1 - (t - nbig)];
        const int img = (int)wl.x;
        const uint2 tk = tasks[(size_t)img * MAXC + wl.y];
        const int off = (int)tk.x, n = (int)tk.y;
        Img im;
        setup_img(im, d, img, lsdw, pix, pxy, fat, reg, pixcount);
        im.reg += off;                    // region list of this component (scratch half = +pixcap stays inside the image's buffer)
        im.used = SB ? reinterpret_cast<u32 *>(smraw + sizeof(GrowSm)) : used_global + (size_t)img * used_words;
        if (SB && im.n > used_words * 32) {       // the bitmap was sized from the previous batch: tell the host, which redoes the batch
            if (lane == 0) atomicMax(&flags[5], im.n);
            continue;
        }
        const float2 *seedcs = scs + (size_t)img * d.pixcap;
        const u32 *ord = corder_ + (size_t)img * d.pixcap + off;
        const u32 *opos = cpos_ + (size_t)img * d.pixcap + off;
        // private shared-memory bitmap: cleared per task.  The global fallback bitmap is shared by all tasks of an image
        // (they own disjoint components, so they never touch the same pixel) and was cleared once by the host before
        // the launch -- clearing it here would wipe the USED bits of tasks of the same image running on other blocks.
        __syncwarp();
        if (SB) for (int i = lane; i < (im.n + 31) / 32; i += 32) im.used[i] = 0;
        __syncwarp();
        const int min_reg = d.min_reg;
        LsdCand *out = cand + (size_t)img * d.segcap;
        u32 *orank = candrank + (size_t)img * d.segcap;
        int ci_next = lane < n ? (int)ord[lane] : -1;
        u32 pos_next = lane < n ? opos[lane] : 0u;
        for (int o0 = 0; o0 < n; o0 += 32) {
            const int ci = ci_next;
            const u32 cpos = pos_next;
            ci_next = o0 + 32 + lane < n ? (int)ord[o0 + 32 + lane] : -1;
            pos_next = o0 + 32 + lane < n ? opos[o0 + 32 + lane] : 0u;
            // this batch of 32 seed candidates: thin record, position, and (for the still unused ones) the fat record
            uint4 tp = make_uint4(0, 0, 0, 0);
            u32 cxy = 0;
            float2 ccs = make_float2(0.f, 0.f);
            bool want = ci >= 0 && !is_used<SB>(im, (u32)ci);
            const u32 wmask = __ballot_sync(FULL, want);
            const int wrank = __popc(wmask & ((1u << lane) - 1u));
            if (want) {
                tp = *reinterpret_cast<const uint4 *>(&im.pix[ci]);
                cxy = im.pxy[ci];
                ccs = seedcs[ci];
                if (wrank < SEED_SLOTS) {
                    const u32 *src = im.fat + (size_t)ci * FAT_WORDS;
#pragma unroll
                    for (int q = 0; q < FAT_WORDS / 4; ++q) cp_async16(&sm.seedrec[wrank][q * 4], src + q * 4);
                }
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
                const int krank = __popc(pfmask & ((1u << k) - 1u));
                u32 sdeg = __shfl_sync(FULL, tp.x, k), sg2 = __shfl_sync(FULL, tp.w, k), sxy = __shfl_sync(FULL, cxy, k);
                float sc = __shfl_sync(FULL, ccs.x, k), ss = __shfl_sync(FULL, ccs.y, k);
                const u32 srank = __shfl_sync(FULL, cpos, k);   // position of the seed in the image's seed order
                if (!have) {
                    sdeg = __float_as_uint(im.pix[seed].deg); sg2 = im.pix[seed].g2; sxy = im.pxy[seed];
                    float2 t2 = seedcs[seed];
                    sc = t2.x; ss = t2.y;
                }
                const u32 *srec = (have && krank < SEED_SLOTS) ? sm.seedrec[krank] : nullptr;
                double reg_angle;
                int nreg = grow<SB>(im, sm, pr, seed, __uint_as_float(sdeg), sg2, sxy, sc, ss, srec, prec, reg_angle);
                if (nreg < min_reg) continue;
                fill_g2(im, nreg);
                Rect rec;
                GPROF(long long t0 = clock64());
                region2rect(im, sm.seq, nreg, reg_angle, prec, p, rec);
                GPROF(long long t1 = clock64());
                bool okr = refine<SB>(im, sm, pr, nreg, reg_angle, prec, p, rec, sc, ss, srec);
                GPROF(pr.t_rect += t1 - t0; pr.t_refine += clock64() - t1; ++pr.cands);
                if (!okr) continue;
                if (lane == 0) {
                    const int slot = atomicAdd(&candcount[img], 1);
                    if (slot < d.segcap) {
                        LsdCand c;
                        c.x1 = rec.x1; c.y1 = rec.y1; c.x2 = rec.x2; c.y2 = rec.y2;
                        c.width = rec.width; c.theta = rec.theta; c.dx = rec.dx; c.dy = rec.dy;
                        out[slot] = c;
                        orank[slot] = srank;
                        const int gs = atomicAdd(&taskctr[3], 1);
                        candlist[gs] = make_uint2((u32)img, (u32)slot);
                    } else {
                        atomicMax(&flags[1], slot + 1);
                    }
                }
            }
        }
        if (lane == 0 && prof) {
            long long *o = prof + (size_t)t * 12;
            o[0] = clock64() - t_start; o[1] = pr.t_grow; o[2] = pr.t_wait; o[3] = pr.t_rect; o[4] = pr.t_refine;
            o[5] = pr.grows; o[6] = pr.rounds; o[7] = pr.accepts; o[8] = pr.refresh; o[9] = pr.cands; o[10] = n; o[11] = img;
        }
    }
}

// ---- kernel 2: NFA validation with LSD_REFINE_ADV rectangle improvement, one warp per candidate ---------------
constexpr int VAL_WARPS = 4;

__global__ void __launch_bounds__(VAL_WARPS * 32, 6) k_lsd_validate(Dims d, const LsdWord *__restrict__ lsdw, const LsdPix *__restrict__ pix,
                                                                   const int *__restrict__ pixcount, const LsdCand *__restrict__ cand,
                                                                const uint2 *__restrict__ candlist, int *__restrict__ taskctr,
                                                                LsdSeg *__restrict__ candseg, u8 *__restrict__ candok)
{
    const int lane = threadIdx.x & 31;
    const int total = min(taskctr[3], d.n * 3 * d.segcap);
    // candidates cost between one and 26 rectangle scans: warps pull them one at a time (a static split left 5 % of
    // the warps active on average, ncu)
    while (true) {
        int t = 0;
        if (lane == 0) t = atomicAdd(&taskctr[4], 1);
        t = __shfl_sync(FULL, t, 0);
        if (t >= total) break;
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

// ---- kernel 3: keep the validated candidates, ordered by the position of their seed in the image's seed order
// (= the reference's acceptance order; the component tasks of an image finish in arbitrary order) ---------------
__global__ void __launch_bounds__(128) k_lsd_emit(Dims d, const int *__restrict__ candcount, const LsdSeg *__restrict__ candseg,
                                                 const u8 *__restrict__ candok, const u32 *__restrict__ candrank,
                                                 LsdSeg *__restrict__ rawseg, int *__restrict__ segcount)
{
    const int img = blockIdx.x, tid = threadIdx.x;
    const int nc = min(candcount[img], d.segcap);
    const size_t base = (size_t)img * d.segcap;
    __shared__ int s_cnt;
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    int mine = 0;
    for (int i = tid; i < nc; i += 128) {
        if (!candok[base + i]) continue;
        const u32 r = candrank[base + i];
        int pos = 0;
        for (int q = 0; q < nc; ++q) pos += (candok[base + q] && candrank[base + q] < r) ? 1 : 0;
        rawseg[base + pos] = candseg[base + i];
        ++mine;
    }
    if (mine) atomicAdd(&s_cnt, mine);
    __syncthreads();
    if (tid == 0) segcount[img] = s_cnt;
}

void launch_lsd_core(const Dims &d, Buffers &b, cudaStream_t st, cudaEvent_t ev_indexed)
{
    static PerDevice tabs, attr;
    tabs.ensure(1, [] {
        cudaMemcpyToSymbol(c_sincos_tab, scr::kSinCosTab, sizeof(scr::kSinCosTab));
        std::vector<double> tab(LGTAB);
        tab[0] = 0.0;
        for (int i = 1; i < LGTAB; ++i) tab[i] = log_gamma_d((double)i);
        cudaMemcpyToSymbol(g_lgtab, tab.data(), LGTAB * sizeof(double));
    });
    const int nimg = d.n * 3, wl_cap = nimg * MAXC;
    k_lsd_index<<<nimg, 256, 0, st>>>(d, b.lsdw, b.pix, b.pxy, b.pixcount, b.g2max, b.fat, b.order, b.scs, b.label, b.csize, b.coff,
                                      b.corder, b.cpos, b.tasks, b.worklist, wl_cap, b.taskctr, b.candcount);
    ++g_launches;
    if (ev_indexed) cudaEventRecord(ev_indexed, st);      // timing split: seed order / records / components | growing
    // USED bitmap: shared memory when it fits (the normal case), the global fallback buffer otherwise
    int used_words = (d.pixcap + 31) / 32;
    if (d.grow_used_bits > 0 && d.grow_used_bits < d.pixcap && !getenv("LSF_FORCE_GLOBAL_USED")) used_words = (d.grow_used_bits + 31) / 32;
    size_t smem = sizeof(GrowSm) + (size_t)used_words * 4;
    u32 *used_global = nullptr;
    if (smem > 200 * 1024 || getenv("LSF_FORCE_GLOBAL_USED")) {
        used_words = (d.pixcap + 31) / 32;
        smem = sizeof(GrowSm); used_global = b.usedbits;
        cudaMemsetAsync(b.usedbits, 0, (size_t)nimg * used_words * sizeof(u32), st);
    }
    if (!used_global) attr.ensure(smem, [&] { cudaFuncSetAttribute(k_lsd_grow<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); });
    const int grid = 148 * (d.grow_per_sm > 0 && d.grow_per_sm < GROW_PER_SM ? d.grow_per_sm : GROW_PER_SM);   // persistent single-warp blocks
    long long *prof = nullptr;
    if (d.debug & 2) { cudaMalloc((void **)&prof, (size_t)wl_cap * 12 * sizeof(long long)); cudaMemsetAsync(prof, 0, (size_t)wl_cap * 12 * sizeof(long long), st); }
    if (used_global)
        k_lsd_grow<false><<<grid, 32, smem, st>>>(d, b.lsdw, b.pix, b.pxy, b.scs, b.fat, b.corder, b.cpos, b.tasks, b.worklist, wl_cap,
                                                  b.taskctr, b.reg, b.pixcount, used_global, used_words, b.cand, b.candrank, b.candcount,
                                                  b.candlist, b.flags, prof);
    else
        k_lsd_grow<true><<<grid, 32, smem, st>>>(d, b.lsdw, b.pix, b.pxy, b.scs, b.fat, b.corder, b.cpos, b.tasks, b.worklist, wl_cap,
                                                 b.taskctr, b.reg, b.pixcount, used_global, used_words, b.cand, b.candrank, b.candcount,
                                                 b.candlist, b.flags, prof);
    ++g_launches;
    if (prof) {
        // developer aid (LSF_TRACE_LSD=2): where the slowest image and the average image spend their cycles
        std::vector<long long> h((size_t)wl_cap * 12);
        cudaStreamSynchronize(st);
        cudaMemcpy(h.data(), prof, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
        cudaFree(prof);
        int worst = 0, nt = 0;
        double avg[12] = {0};
        for (int i = 0; i < wl_cap; ++i) {
            if (h[(size_t)i * 12] == 0) continue;
            ++nt;
            if (h[(size_t)i * 12] > h[(size_t)worst * 12]) worst = i;
            for (int q = 0; q < 12; ++q) avg[q] += (double)h[(size_t)i * 12 + q];
        }
        for (int q = 0; q < 12; ++q) avg[q] /= nt ? nt : 1;
        const char *nm[12] = {"total", "grow", "wait", "rect", "refine", "grows", "rounds", "accepts", "refresh", "cands", "npix", "img"};
        fprintf(stderr, "[lsf grow prof] %d tasks; worst task %d:", nt, worst);
        for (int q = 0; q < 12; ++q) fprintf(stderr, " %s=%lld", nm[q], h[(size_t)worst * 12 + q]);
        fprintf(stderr, "\n[lsf grow prof] mean:");
        for (int q = 0; q < 12; ++q) fprintf(stderr, " %s=%.0f", nm[q], avg[q]);
        fprintf(stderr, "\n");
    }
}

void launch_lsd_validate(const Dims &d, Buffers &b, cudaStream_t st)
{
    k_lsd_validate<<<148 * 6, VAL_WARPS * 32, 0, st>>>(d, b.lsdw, b.pix, b.pixcount, b.cand, b.candlist, b.taskctr, b.candseg, b.candok);
    ++g_launches;
    k_lsd_emit<<<d.n * 3, 128, 0, st>>>(d, b.candcount, b.candseg, b.candok, b.candrank, b.rawseg, b.segcount);
    ++g_launches;
}

}  // namespace lsf
