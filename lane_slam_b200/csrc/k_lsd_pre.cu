// k_lsd_pre.cu -- K6+K7(+K8 compaction): LSD pre-processing of one binary edge_color image:
// 7x7 fixed-point Gaussian (sigma 0.75; effective taps 4,56,136,56,4), 0.8x INTER_LINEAR_EXACT resize,
// 2x2 gradient, level-line angle (fastAtan2), "defined" test, raster-ordered compaction of support pixels.
//
// Replaces the first half of cv2.LineSegmentDetector.detect (line_detector_lsd.py:64-67):
// GaussianBlur + resize + ll_angle.  Input is a packed bit-plane (edge_color is
// 0/255), so the source costs N/8 bytes; the output is sparse: one {bits, base} word per 32 scaled
// pixels plus one 16-byte record (+ 4-byte position) per support pixel (0.3-3 % of the pixels).
//
// One CTA per (frame, colour) walks the image in bands of 8 scaled rows, so the running count of support
// pixels (= raster-order compact index) is known without a second pass.  Edges are sparse: per band the
// CTA derives which 32-pixel column words can be non-zero at each stage (from the OR of the band's source
// bit rows) and runs every stage only on those words; empty bands cost one load + one barrier.
#include <algorithm>
#include <cstdlib>
#include "common.cuh"

namespace lsf {

constexpr int PT = 256;   // threads
constexpr int NW = PT / 32;
constexpr int BR = 8;     // scaled rows per band (power of two: task decoding uses shifts)
static_assert(BR == 8, "task decoding below assumes BR == 8");
constexpr int HZR = 16;   // horizontal-blur rows held per band (source rows s0-2 .. s0+13)
constexpr int GR = 12;    // blurred rows per band (source rows s0 .. s0+11)

__device__ __forceinline__ int refl101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
    return i;
}

__device__ __forceinline__ int tapw(int j) { return (j == 0) ? 136 : (j == 1 || j == -1) ? 56 : 4; }

// compact the indices i in [0, n) with flag[i] != 0 into list[]; returns the count (call from one warp)
__device__ __forceinline__ int warp_compact(const u8 *flag, int n, u16 *list)
{
    const int lane = threadIdx.x & 31;
    int cnt = 0;
    for (int i0 = 0; i0 < n; i0 += 32) {
        int i = i0 + lane;
        bool f = i < n && flag[i];
        u32 m = __ballot_sync(0xffffffffu, f);
        if (f) list[cnt + __popc(m & ((1u << lane) - 1u))] = (u16)i;
        cnt += __popc(m);
    }
    return cnt;
}

__global__ void __launch_bounds__(PT) k_lsd_pre(Dims d, u32 g2_min, const u32 *__restrict__ planesB, LsdWord *__restrict__ lsdw,
                                               LsdPix *__restrict__ pix, u32 *__restrict__ pxy,
                                               int *__restrict__ pixcount, u32 *__restrict__ g2max, int *__restrict__ flags)
{
    extern __shared__ __align__(16) u8 smraw[];
    const int img = blockIdx.x, f = img / 3, c = img - f * 3;
    const int w = d.w, h = d.h, wp = d.wp, sw = d.sw, sh = d.sh, swp = d.swp;
    const u32 *src = planesB + ((size_t)f * PB_COUNT + PB_EC0 + c) * (size_t)h * wp;
    // smem carve-up
    u16 *hz = (u16 *)smraw;                            // [HZR][w]
    u8 *g = (u8 *)(hz + (size_t)HZR * w);              // [GR][w]
    u8 *sc = g + (size_t)GR * w;                       // [BR+1][sw]
    u32 *sbits = (u32 *)(sc + (((size_t)(BR + 1) * sw + 3) & ~(size_t)3));  // [HZR][wp]
    u32 *wbits = sbits + (size_t)HZR * wp;             // [BR*swp]
    u32 *wbase = wbits + (size_t)BR * swp;             // [BR*swp]
    u16 *listG = (u16 *)(wbase + (size_t)BR * swp);    // [wp]   source words needed by the blur stages
    u16 *listS = listG + wp;                           // [swp]  scaled words needed by the resize stage
    u16 *listD = listS + swp;                          // [swp]  scaled words that may hold support pixels
    u8 *actS = (u8 *)(listD + swp);                    // [wp]   source word has an edge bit in this band
    u8 *needG = actS + wp;                             // [wp]
    u8 *actD = needG + wp;                             // [swp]
    u8 *needS = actD + swp;                            // [swp]
    __shared__ u32 s_run, s_warp_tot[NW], s_gmax;
    __shared__ int s_nG, s_nS, s_nD;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    LsdWord *ow = lsdw + (size_t)img * sh * swp;
    LsdPix *opix = pix + (size_t)img * d.pixcap;
    u32 *opxy = pxy + (size_t)img * d.pixcap;
    if (tid == 0) { s_run = 0; s_gmax = 0; }
    u32 my_gmax = 0;
    __syncthreads();

    for (int ys0 = 0; ys0 < sh; ys0 += BR) {
        const int s0 = ys0 + (ys0 >> 2);           // first source row of the band
        const int hz_lo = s0 - 2;                  // source row held in hz slot 0
        const int band_rows = min(BR, sh - ys0);
        const int ntask = band_rows * swp;
        // ---- load the band's source bit rows; column activity ----
        for (int i = tid; i < wp; i += PT) { actS[i] = 0; needG[i] = 0; }
        for (int i = tid; i < swp; i += PT) { actD[i] = 0; needS[i] = 0; }
        __syncthreads();
        int nz = 0;
        for (int i = tid; i < HZR * wp; i += PT) {
            int r = i / wp, xw = i - r * wp, yy = hz_lo + r;
            u32 v = (yy >= 0 && yy < h) ? src[(size_t)yy * wp + xw] : 0u;
            sbits[i] = v;
            if (v) { actS[xw] = 1; nz = 1; }
        }
        nz = __syncthreads_or(nz);
        if (!nz) {
            u32 run = s_run;
            for (int i = tid; i < ntask; i += PT) ow[(size_t)ys0 * swp + i] = LsdWord{0u, run};
            continue;
        }
        // scaled word xsw can be non-zero only if a source bit lies in columns [40 xsw - 2, 40 xsw + 43]
        for (int xsw = tid; xsw < swp; xsw += PT) {
            int wlo = max(0, (40 * xsw - 2) >> 5), whi = min(wp - 1, (40 * xsw + 43) >> 5);
            int a = 0;
            for (int k = wlo; k <= whi; ++k) a |= actS[k];
            if (a) {
                actD[xsw] = 1;
                needS[xsw] = 1;
                if (xsw + 1 < swp) needS[xsw + 1] = 1;   // right neighbour of lane 31
            }
        }
        __syncthreads();
        for (int xsw = tid; xsw < swp; xsw += PT) {
            if (needS[xsw]) {
                int wlo = (40 * xsw) >> 5, whi = min(wp - 1, (40 * xsw + 40) >> 5);
                for (int k = wlo; k <= whi; ++k) needG[k] = 1;
            }
        }
        __syncthreads();
        if (warp == 0) { int nG = warp_compact(needG, wp, listG); if (lane == 0) s_nG = nG; }
        if (warp == 1) { int nS = warp_compact(needS, swp, listS); if (lane == 0) s_nS = nS; }
        if (warp == 2) { int nD = warp_compact(actD, swp, listD); if (lane == 0) s_nD = nD; }
        for (int i = tid; i < ntask; i += PT) { wbits[i] = 0; wbase[i] = 0; }
        __syncthreads();
        const int nG = s_nG, nS = s_nS, nD = s_nD;

        // ---- horizontal blur on bits (taps 4,56,136,56,4; reflect-101), value/255 in Q8 ----
        for (int t = warp; t < HZR * nG; t += NW) {
            int r = t & (HZR - 1), xw = listG[t >> 4], x = xw * 32 + lane;   // HZR == 16
            const u32 *row = sbits + r * wp;
            int acc = 0;
            if (xw > 0 && xw * 32 + 33 < w) {
                // interior word: 36-bit window = 2 bits of the left word | this word | 2 bits of the right word
                unsigned long long win = ((unsigned long long)row[xw] << 2) | (row[xw - 1] >> 30) |
                                         ((unsigned long long)(row[xw + 1] & 3u) << 34);
                u32 v = (u32)(win >> lane) & 31u;   // bit j <-> column x - 2 + j
                acc = 4 * (int)((v & 1u) + ((v >> 4) & 1u)) + 56 * (int)(((v >> 1) & 1u) + ((v >> 3) & 1u)) + 136 * (int)((v >> 2) & 1u);
            } else if (x < w) {
#pragma unroll
                for (int j = -2; j <= 2; ++j) {
                    int xx = refl101(x + j, w);
                    acc += tapw(j) * (int)((row[xx >> 5] >> (xx & 31)) & 1u);
                }
            }
            if (x < w) hz[r * w + x] = (u16)acc;
        }
        __syncthreads();
        // ---- vertical blur -> g (u8) for source rows s0 .. s0+GR-1 (clamped to h-1) ----
        for (int t = warp; t < GR * nG; t += NW) {
            int q = t / GR, r = t - q * GR, x = listG[q] * 32 + lane;       // GR is a compile-time constant
            if (x < w) {
                int gy = min(s0 + r, h - 1);
                int acc = 0;
                if (gy >= 2 && gy + 2 < h) {
                    const u16 *hp = hz + (gy - hz_lo) * w + x;
                    acc = 4 * ((int)hp[-2 * w] + (int)hp[2 * w]) + 56 * ((int)hp[-w] + (int)hp[w]) + 136 * (int)hp[0];
                } else {
#pragma unroll
                    for (int j = -2; j <= 2; ++j) {
                        int yy = refl101(gy + j, h);
                        acc += tapw(j) * (int)hz[(yy - hz_lo) * w + x];
                    }
                }
                g[r * w + x] = (u8)((acc * 255 + 32768) >> 16);
            }
        }
        __syncthreads();
        // ---- 0.8x bilinear (INTER_LINEAR_EXACT) -> scaled rows ys0 .. ys0+BR ----
        for (int t = warp; t < (BR + 1) * nS; t += NW) {
            int q = t / (BR + 1), rs = t - q * (BR + 1), xs = listS[q] * 32 + lane, ys = ys0 + rs;
            if (xs < sw) {
                u8 val = 0;
                if (ys < sh) {
                    int sy = ys + (ys >> 2), sx = xs + (xs >> 2);
                    int ay = 32 + 64 * (ys & 3), ax = 32 + 64 * (xs & 3);
                    int r0 = min(sy, h - 1) - s0, r1 = min(sy + 1, h - 1) - s0;
                    int x0 = min(sx, w - 1), x1 = min(sx + 1, w - 1);
                    int h0 = g[r0 * w + x0] * (256 - ax) + g[r0 * w + x1] * ax;
                    int h1 = g[r1 * w + x0] * (256 - ax) + g[r1 * w + x1] * ax;
                    val = (u8)((h0 * (256 - ay) + h1 * ay + 32768) >> 16);
                }
                sc[rs * sw + xs] = val;
            }
        }
        __syncthreads();
        // ---- gradient / defined bits on the active words ----
        for (int t = warp; t < BR * nD; t += NW) {
            int rs = t & (BR - 1), xw = listD[t >> 3], xs = xw * 32 + lane, ys = ys0 + rs;   // BR == 8
            if (rs >= band_rows) continue;                                                   // warp-uniform
            bool def = false;
            if (xs < sw - 1 && ys < sh - 1) {
                const u8 *r0 = sc + rs * sw + xs, *r1 = r0 + sw;
                int DA = (int)r1[1] - (int)r0[0], BC = (int)r0[1] - (int)r1[0];
                int gx = DA + BC, gy = DA - BC;
                u32 g2 = (u32)(gx * gx + gy * gy);
                def = g2 >= g2_min;
            }
            u32 bits = __ballot_sync(0xffffffffu, def);
            if (lane == 0) { wbits[rs * swp + xw] = bits; wbase[rs * swp + xw] = __popc(bits); }
        }
        __syncthreads();
        // ---- exclusive scan of the per-word counts over the band ----
        {
            u32 run = s_run;
            int per = (ntask + PT - 1) / PT;
            int b0 = tid * per, b1 = min(ntask, b0 + per);
            u32 sum = 0;
            for (int i = b0; i < b1; ++i) sum += wbase[i];
            u32 incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                u32 v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            if (lane == 31) s_warp_tot[warp] = incl;
            __syncthreads();
            u32 woff = 0;
            for (int k = 0; k < warp; ++k) woff += s_warp_tot[k];
            u32 excl = run + woff + incl - sum;
            for (int i = b0; i < b1; ++i) { u32 cnt = wbase[i]; wbase[i] = excl; excl += cnt; }
            __syncthreads();
            if (tid == PT - 1) s_run = excl;   // last thread ends at the band total
        }
        __syncthreads();
        // ---- words of the band, then the compact records of the active words ----
        for (int i = tid; i < ntask; i += PT) ow[(size_t)ys0 * swp + i] = LsdWord{wbits[i], wbase[i]};
        for (int t = warp; t < BR * nD; t += NW) {
            int rs = t & (BR - 1), xw = listD[t >> 3], xs = xw * 32 + lane, ys = ys0 + rs;
            if (rs >= band_rows) continue;
            u32 bits = wbits[rs * swp + xw], base = wbase[rs * swp + xw];
            if ((bits >> lane) & 1u) {
                u32 idx = base + __popc(bits & ((1u << lane) - 1u));
                const u8 *r0 = sc + rs * sw + xs, *r1 = r0 + sw;
                int DA = (int)r1[1] - (int)r0[0], BC = (int)r0[1] - (int)r1[0];
                int gx = DA + BC, gy = DA - BC;
                u32 g2 = (u32)(gx * gx + gy * gy);
                my_gmax = max(my_gmax, g2);
                if (idx < (u32)d.pixcap) {
                    float a = fast_atan2_deg((float)gx, (float)-gy);
                    double ar = (double)a * (3.14159265358979323846 / 180.0);
                    float af = (float)ar;
                    LsdPix p;
                    p.deg = a;
                    p.c = (float)cos((double)af);   // cosf(float(angle)), correctly rounded
                    p.s = (float)sin((double)af);
                    p.g2 = g2;
                    opix[idx] = p;
                    opxy[idx] = ((u32)ys << 16) | (u32)xs;
                }
            }
        }
        __syncthreads();
    }
    atomicMax(&s_gmax, my_gmax);
    __syncthreads();
    if (tid == 0) {
        u32 n = s_run;
        if (n > (u32)d.pixcap) { atomicMax(&flags[0], (int)n); n = d.pixcap; }
        pixcount[img] = (int)n;
        g2max[img] = s_gmax;
    }
}

// =====================================================================================================================
// v2: no block-level synchronisation.  One WARP owns a task = (colour image, band of 8 scaled rows, 32 scaled columns)
// and runs every stage for it in a 2.3 KB private shared-memory patch: the source bits it depends on (16 rows x 46
// columns) -> horizontal 5-tap (42 columns) -> vertical 5-tap (12 rows) -> 0.8x bilinear (9 x 33) -> gradient bits.
// Adjacent tasks recompute the few columns / rows they share instead of exchanging them.  Most tasks see no edge bit
// at all and leave after one load.  The raster-order compact index needs the support-pixel count of every earlier
// word, so the work is split in three launches:
//   k_lsd_pre_a  all tasks: defined bits + count per word; tasks with support pixels are appended to an active list
//   k_lsd_pre_b  per image: exclusive scan of the word counts (-> LsdWord.base, pixcount)
//   k_lsd_pre_c  active tasks only: from the 324-byte scaled patch pass a kept, write the 16-byte records + positions
//                and the per-image maximum gradient
// Arithmetic is op-for-op that of k_lsd_pre (v1, kept for LSF_PRE_V1=1).
constexpr int PCW = 42;                                    // source columns of a task: 40 xsw .. 40 xsw + 41
constexpr int PATCH_WORDS = (BR + 1) * 36 / 4;          // the scaled patch of a task: 9 rows x 36 bytes
struct PreSm {
    unsigned long long win[16];                            // per source row: bit k <-> column 40 xsw - 2 + k (46 bits)
    u16 hz[16][PCW + 2];
    u8 g[12][PCW + 6];
    __align__(4) u8 sc[BR + 1][36];
};

// fills sm.sc for task (b, xsw); `rows` = the 16 source bit rows s0-2 .. s0+13 of the band ([16][wp], zero outside the
// frame).  Returns false when no source bit can reach the task.
__device__ __forceinline__ bool pre_patch(const Dims &d, const u32 *rows, int b, int xsw, PreSm &sm)
{
    const int lane = threadIdx.x & 31;
    const int w = d.w, h = d.h, wp = d.wp, sw = d.sw, sh = d.sh;
    const int ys0 = b * BR, s0 = ys0 + (ys0 >> 2), c0 = 40 * xsw;
    // ---- source bits: lane r holds row s0 - 2 + r; bit k <-> column c0 - 2 + k, reflect-101 at the frame borders ----
    unsigned long long win = 0;
    if (lane < 16) {
        const u32 *row = rows + lane * wp;
        if (c0 >= 2) {
            const int j0 = (c0 - 2) >> 5, sft = (c0 - 2) & 31;
            const u32 lo = row[j0], mid = j0 + 1 < wp ? row[j0 + 1] : 0u, hi = j0 + 2 < wp ? row[j0 + 2] : 0u;
            win = (((unsigned long long)mid << 32) | lo) >> sft;
            if (sft) win |= (unsigned long long)hi << (64 - sft);
            win &= (1ull << 46) - 1ull;
            const int nin = w - (c0 - 2);                  // window columns inside the frame (bits >= nin lie right of it)
            if (nin < 46) {
                win &= (1ull << nin) - 1ull;               // (plane words carry no bits beyond w, but be explicit)
                if (win) {
                    // right border: column w-1+j mirrors w-1-j, which is inside this window (c0 + 43 >= w)
                    // (only the two columns right of the frame are ever used by the 5-tap blur of column w-1)
                    for (int k = nin; k < min(46, nin + 2); ++k) win |= ((win >> (2 * (nin - 1) - k)) & 1ull) << k;
                }
            }
        } else {
            // left border (c0 == 0): columns -2, -1 mirror columns 2, 1
            const unsigned long long in = ((unsigned long long)(wp > 1 ? row[1] : 0u) << 32) | row[0];
            win = ((in << 2) | ((in >> 2) & 1ull) | (((in >> 1) & 1ull) << 1)) & ((1ull << 46) - 1ull);
            const int nin = w + 2;                         // tiny frames: right border inside the same window
            if (nin < 46) {
                win &= (1ull << nin) - 1ull;
                for (int k = nin; k < min(46, nin + 2); ++k) win |= ((win >> (2 * (nin - 1) - k)) & 1ull) << k;
            }
        }
    }
    if (!__ballot_sync(0xffffffffu, win != 0ull)) return false;
    if (lane < 16) sm.win[lane] = win;
    __syncwarp();
    // ---- horizontal blur on bits (taps 4,56,136,56,4), value/255 in Q8 ----
    for (int v = lane; v < 16 * PCW; v += 32) {
        const int r = v / PCW, k = v - r * PCW;
        const u32 q = (u32)(sm.win[r] >> k) & 31u;         // bit j <-> column c0 + k - 2 + j
        sm.hz[r][k] = (u16)(4 * (int)((q & 1u) + ((q >> 4) & 1u)) + 56 * (int)(((q >> 1) & 1u) + ((q >> 3) & 1u)) + 136 * (int)((q >> 2) & 1u));
    }
    __syncwarp();
    // ---- vertical blur -> g (u8) for source rows s0 .. s0+11 (clamped to h-1), reflect-101 in y ----
    for (int v = lane; v < 12 * PCW; v += 32) {
        const int r = v / PCW, k = v - r * PCW;
        const int gy = min(s0 + r, h - 1);
        int acc;
        if (gy >= 2 && gy + 2 < h) {
            const int q = gy - (s0 - 2);
            acc = 4 * ((int)sm.hz[q - 2][k] + (int)sm.hz[q + 2][k]) + 56 * ((int)sm.hz[q - 1][k] + (int)sm.hz[q + 1][k]) + 136 * (int)sm.hz[q][k];
        } else {
            acc = 0;
#pragma unroll
            for (int j = -2; j <= 2; ++j) acc += tapw(j) * (int)sm.hz[refl101(gy + j, h) - (s0 - 2)][k];
        }
        sm.g[r][k] = (u8)((acc * 255 + 32768) >> 16);
    }
    __syncwarp();
    // ---- 0.8x bilinear (INTER_LINEAR_EXACT) -> scaled rows ys0 .. ys0+8, columns 32 xsw .. 32 xsw + 32 ----
    for (int v = lane; v < (BR + 1) * 33; v += 32) {
        const int rs = v / 33, kx = v - rs * 33;
        const int ys = ys0 + rs, xs = 32 * xsw + kx;
        u8 val = 0;
        if (xs < sw && ys < sh) {
            const int sy = ys + (ys >> 2), sx = xs + (xs >> 2);
            const int ay = 32 + 64 * (ys & 3), ax = 32 + 64 * (xs & 3);
            const int r0 = min(sy, h - 1) - s0, r1 = min(sy + 1, h - 1) - s0;
            const int x0 = min(sx, w - 1) - c0, x1 = min(sx + 1, w - 1) - c0;
            const int h0 = sm.g[r0][x0] * (256 - ax) + sm.g[r0][x1] * ax;
            const int h1 = sm.g[r1][x0] * (256 - ax) + sm.g[r1][x1] * ax;
            val = (u8)((h0 * (256 - ay) + h1 * ay + 32768) >> 16);
        }
        sm.sc[rs][kx] = val;
    }
    __syncwarp();
    return true;
}

// loads the 16 source bit rows of band b into `rows` ([16][wp], zero outside the frame); returns true if any bit is set
__device__ __forceinline__ bool pre_load_band(const Dims &d, const u32 *__restrict__ src, int b, u32 *rows)
{
    const int lane = threadIdx.x & 31, wp = d.wp, h = d.h;
    const int ys0 = b * BR, s0 = ys0 + (ys0 >> 2);
    u32 nz = 0;
    for (int i = lane; i < 16 * wp; i += 32) {
        const int r = i / wp, xw = i - r * wp, yy = s0 - 2 + r;
        const u32 v = (yy >= 0 && yy < h) ? src[(size_t)yy * wp + xw] : 0u;
        rows[i] = v;
        nz |= v;
    }
    __syncwarp();
    return __ballot_sync(0xffffffffu, nz != 0) != 0;
}

__global__ void __launch_bounds__(PT) k_lsd_pre_a(Dims d, u32 g2_min, int nbands, const u32 *__restrict__ planesB, LsdWord *__restrict__ lsdw,
                                                 u32 *__restrict__ active, int *__restrict__ nactive, u32 *__restrict__ patches)
{
    extern __shared__ __align__(16) u8 pre_dyn[];       // per warp: PreSm + 16 x wp words of source bits
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t per_warp = sizeof(PreSm) + (size_t)16 * d.wp * 4;
    PreSm &sm = *reinterpret_cast<PreSm *>(pre_dyn + warp * per_warp);
    u32 *rows = reinterpret_cast<u32 *>(pre_dyn + warp * per_warp + sizeof(PreSm));
    const int swp = d.swp, sw = d.sw, sh = d.sh;
    const int nrow = d.n * 3 * nbands;                   // (image, band) pairs
    for (int t = blockIdx.x * NW + warp; t < nrow; t += gridDim.x * NW) {
        const int img = t / nbands, b = t - img * nbands;
        const int f = img / 3, c = img - f * 3, ys0 = b * BR;
        const u32 *src = planesB + ((size_t)f * PB_COUNT + PB_EC0 + c) * (size_t)d.h * d.wp;
        LsdWord *ow = lsdw + (size_t)img * sh * swp;
        __syncwarp();
        if (!pre_load_band(d, src, b, rows)) {
            // most bands hold no edge at all
            for (int i = lane; i < BR * swp; i += 32) {
                const int rs = i / swp;
                if (ys0 + rs < sh) ow[(size_t)ys0 * swp + i] = LsdWord{0u, 0u};
            }
            continue;
        }
        for (int xsw = 0; xsw < swp; ++xsw) {
            const bool any = pre_patch(d, rows, b, xsw, sm);
            u32 mybits = 0;                                 // lane rs keeps the word of row ys0 + rs
            if (any) {
                for (int rs = 0; rs < BR; ++rs) {
                    const int ys = ys0 + rs, xs = 32 * xsw + lane;
                    bool def = false;
                    if (xs < sw - 1 && ys < sh - 1) {
                        const u8 *r0 = &sm.sc[rs][lane], *r1 = &sm.sc[rs + 1][lane];
                        const int DA = (int)r1[1] - (int)r0[0], BC = (int)r0[1] - (int)r1[0];
                        const int gx = DA + BC, gy = DA - BC;
                        def = (u32)(gx * gx + gy * gy) >= g2_min;
                    }
                    const u32 bits = __ballot_sync(0xffffffffu, def);
                    if (lane == rs) mybits = bits;
                }
            }
            if (lane < BR && ys0 + lane < sh) ow[(size_t)(ys0 + lane) * swp + xsw] = LsdWord{mybits, (u32)__popc(mybits)};
            if (__ballot_sync(0xffffffffu, mybits != 0)) {
                // active task: keep its scaled patch (9 x 36 bytes) for the record pass
                int slot = 0;
                if (lane == 0) { slot = atomicAdd(nactive, 1); active[slot] = (u32)(t * swp + xsw); }
                slot = __shfl_sync(0xffffffffu, slot, 0);
                u32 *dst = patches + (size_t)slot * PATCH_WORDS;
                const u32 *ps = reinterpret_cast<const u32 *>(&sm.sc[0][0]);
                for (int i = lane; i < PATCH_WORDS; i += 32) dst[i] = ps[i];
            }
            __syncwarp();
        }
    }
}

// per image: exclusive scan of the word counts in raster order -> LsdWord.base; pixel count (clamped), max gradient reset
__global__ void __launch_bounds__(PT) k_lsd_pre_b(Dims d, LsdWord *__restrict__ lsdw, int *__restrict__ pixcount, u32 *__restrict__ g2max,
                                                 int *__restrict__ flags)
{
    __shared__ u32 s_w[NW];
    __shared__ u32 s_run;
    const int img = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    LsdWord *ow = lsdw + (size_t)img * d.sh * d.swp;
    const int nw = d.sh * d.swp;
    if (tid == 0) s_run = 0;
    __syncthreads();
    for (int i0 = 0; i0 < nw; i0 += PT) {
        const int i = i0 + tid;
        const u32 c = i < nw ? ow[i].base : 0u;
        u32 incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u32 v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        u32 base = s_run;
        for (int k = 0; k < warp; ++k) base += s_w[k];
        if (i < nw) ow[i].base = base + incl - c;
        __syncthreads();
        if (tid == PT - 1) s_run = base + incl;
        __syncthreads();
    }
    if (tid == 0) {
        u32 n = s_run;
        if (n > (u32)d.pixcap) { atomicMax(&flags[0], (int)n); n = d.pixcap; }
        pixcount[img] = (int)n;
        g2max[img] = 0;
    }
}

__global__ void __launch_bounds__(PT) k_lsd_pre_c(Dims d, int nbands, const LsdWord *__restrict__ lsdw, const u32 *__restrict__ active,
                                                 const int *__restrict__ nactive, const u32 *__restrict__ patches,
                                                 LsdPix *__restrict__ pix, u32 *__restrict__ pxy, u32 *__restrict__ g2max)
{
    __shared__ __align__(4) u8 s_sc[NW][BR + 1][36];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u8 (*sc)[36] = s_sc[warp];
    const int swp = d.swp, sh = d.sh;
    const int per_img = nbands * swp, na = *nactive;
    for (int a = blockIdx.x * NW + warp; a < na; a += gridDim.x * NW) {
        const int t = (int)active[a];
        const int img = t / per_img, rem = t - img * per_img, b = rem / swp, xsw = rem - b * swp;
        const int ys0 = b * BR;
        const LsdWord *ow = lsdw + (size_t)img * sh * swp;
        LsdPix *opix = pix + (size_t)img * d.pixcap;
        u32 *opxy = pxy + (size_t)img * d.pixcap;
        __syncwarp();
        {
            const u32 *ps = patches + (size_t)a * PATCH_WORDS;
            u32 *dw = reinterpret_cast<u32 *>(&sc[0][0]);
            for (int i = lane; i < PATCH_WORDS; i += 32) dw[i] = ps[i];
        }
        __syncwarp();
        u32 gmax = 0;
        for (int rs = 0; rs < BR && ys0 + rs < sh; ++rs) {
            const int ys = ys0 + rs, xs = 32 * xsw + lane;
            const LsdWord wd = ow[(size_t)ys * swp + xsw];
            if ((wd.bits >> lane) & 1u) {
                const u32 idx = wd.base + __popc(wd.bits & ((1u << lane) - 1u));
                const u8 *r0 = &sc[rs][lane], *r1 = &sc[rs + 1][lane];
                const int DA = (int)r1[1] - (int)r0[0], BC = (int)r0[1] - (int)r1[0];
                const int gx = DA + BC, gy = DA - BC;
                const u32 g2 = (u32)(gx * gx + gy * gy);
                gmax = max(gmax, g2);
                if (idx < (u32)d.pixcap) {
                    const float ang = fast_atan2_deg((float)gx, (float)-gy);
                    const double ar = (double)ang * (3.14159265358979323846 / 180.0);
                    const float af = (float)ar;
                    LsdPix p;
                    p.deg = ang;
                    p.c = (float)cos((double)af);           // cosf / sinf of float(angle), correctly rounded
                    p.s = (float)sin((double)af);
                    p.g2 = g2;
                    opix[idx] = p;
                    opxy[idx] = ((u32)ys << 16) | (u32)xs;
                }
            }
        }
        gmax = __reduce_max_sync(0xffffffffu, gmax);
        if (lane == 0 && gmax) atomicMax(&g2max[img], gmax);
    }
}

static size_t lsd_pre_smem(const Dims &d)
{
    size_t s = (size_t)HZR * d.w * 2 + (size_t)GR * d.w + (((size_t)(BR + 1) * d.sw + 3) & ~(size_t)3) +
               (size_t)HZR * d.wp * 4 + (size_t)BR * d.swp * 8 + (size_t)(d.wp + 2 * d.swp) * 2 + (size_t)(2 * d.wp + 2 * d.swp);
    return s + 16;
}

void launch_lsd_pre(const Dims &d, const u32 *planesB, Buffers &b, cudaStream_t st)
{
    size_t smem = lsd_pre_smem(d);
    static PerDevice attr, attr2;
    // smallest integer g2 with sqrt(g2/4.0) > rho, rho = 2/sin(22.5 deg)   (ll_angle threshold)
    const double rho = 2.0 / sin(3.14159265358979323846 * 22.5 / 180.0);
    u32 g2_min = 0;
    while (!(sqrt((double)g2_min / 4.0) > rho)) ++g2_min;
    const size_t smem2 = NW * (sizeof(PreSm) + (size_t)16 * d.wp * 4);
    if (getenv("LSF_PRE_V1") || smem2 > 200 * 1024) {     // v1: frames wider than ~12 000 pixels (and A/B runs)
        attr.ensure(smem, [&] { cudaFuncSetAttribute(k_lsd_pre, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); });
        k_lsd_pre<<<d.n * 3, PT, smem, st>>>(d, g2_min, planesB, b.lsdw, b.pix, b.pxy, b.pixcount, b.g2max, b.flags);
        ++g_launches;
        return;
    }
    const int nbands = (d.sh + BR - 1) / BR;
    const long long nrow = (long long)d.n * 3 * nbands;
    const int grid_a = (int)std::min<long long>((nrow + NW - 1) / NW, 148 * 8);
    attr2.ensure(smem2, [&] { cudaFuncSetAttribute(k_lsd_pre_a, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2); });
    k_lsd_pre_a<<<grid_a, PT, smem2, st>>>(d, g2_min, nbands, planesB, b.lsdw, b.preact, b.prectr, b.prepatch);
    k_lsd_pre_b<<<d.n * 3, PT, 0, st>>>(d, b.lsdw, b.pixcount, b.g2max, b.flags);
    k_lsd_pre_c<<<148 * 8, PT, 0, st>>>(d, nbands, b.lsdw, b.preact, b.prectr, b.prepatch, b.pix, b.pxy, b.g2max);
    g_launches += 3;
}

}  // namespace lsf
