// k_lsd_pre.cu -- K6+K7(+K8 compaction): LSD pre-processing of one binary edge_color image:
// 7x7 fixed-point Gaussian (sigma 0.75), 0.8x INTER_LINEAR_EXACT resize, 2x2 gradient, level-line
// angle (fastAtan2), "defined" test, raster-ordered compaction of the support pixels.
//
// Replaces the first half of cv2.LineSegmentDetector.detect (line_detector_lsd.py:64-67):
// GaussianBlur + resize + ll_angle.  Input is a packed bit-plane (edge_color is
// 0/255), so the source costs N/8 bytes; the output is sparse: one {bits, base} word per 32 scaled
// pixels plus one 16-byte record per support pixel (0.3-3 % of the pixels).
//
// One CTA per (frame, colour); the CTA walks the image in bands of 8 scaled rows, so the running
// count of support pixels (= raster-order compact index) is known without a second pass.
#include "common.cuh"

namespace lsf {

constexpr int PT = 256;   // threads
constexpr int BR = 8;     // scaled rows per band
constexpr int HZR = 16;   // horizontal-blur rows held per band (source rows s0-2 .. s0+13)
constexpr int GR = 12;    // blurred rows per band (source rows s0 .. s0+11)

__device__ __forceinline__ int refl101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
    return i;
}

__global__ void __launch_bounds__(PT) k_lsd_pre(Dims d, u32 g2_min, const u32 *__restrict__ planesB, LsdWord *__restrict__ lsdw,
                                               LsdPix *__restrict__ pix, u32 *__restrict__ pixxy, u8 *__restrict__ used,
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
    __shared__ u32 s_run, s_warp_tot[PT / 32];
    __shared__ u32 s_gmax;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    LsdWord *ow = lsdw + (size_t)img * sh * swp;
    LsdPix *opix = pix + (size_t)img * d.pixcap;
    u32 *oxy = pixxy + (size_t)img * d.pixcap;
    u8 *oused = used + (size_t)img * d.pixcap;
    if (tid == 0) { s_run = 0; s_gmax = 0; }
    u32 my_gmax = 0;
    __syncthreads();

    for (int ys0 = 0; ys0 < sh; ys0 += BR) {
        const int s0 = ys0 + (ys0 >> 2);           // first source row of the band
        const int hz_lo = s0 - 2;                  // source row held in hz slot 0
        // ---- load source bit rows; detect an empty band ----
        int nz = 0;
        for (int i = tid; i < HZR * wp; i += PT) {
            int r = i / wp, yy = hz_lo + r;
            u32 v = (yy >= 0 && yy < h) ? src[(size_t)yy * wp + (i - r * wp)] : 0u;
            sbits[i] = v;
            nz |= v != 0;
        }
        nz = __syncthreads_or(nz);
        const int band_rows = min(BR, sh - ys0);
        if (!nz) {
            u32 run = s_run;
            for (int i = tid; i < band_rows * swp; i += PT) ow[(size_t)ys0 * swp + i] = LsdWord{0u, run};
            continue;
        }
        // ---- horizontal blur taps [4,56,136,56,4] on bits (reflect-101), value/255 in Q8 ----
        for (int i = tid; i < HZR * w; i += PT) {
            int r = i / w, x = i - r * w;
            const u32 *row = sbits + r * wp;
            int acc = 0;
#pragma unroll
            for (int j = -2; j <= 2; ++j) {
                int xx = refl101(x + j, w);
                int q = (j == 0) ? 136 : (j == 1 || j == -1) ? 56 : 4;
                acc += q * (int)((row[xx >> 5] >> (xx & 31)) & 1u);
            }
            hz[i] = (u16)acc;
        }
        __syncthreads();
        // ---- vertical blur -> g (u8) for source rows s0 .. s0+GR-1 (clamped to h-1) ----
        for (int i = tid; i < GR * w; i += PT) {
            int r = i / w, x = i - r * w;
            int gy = min(s0 + r, h - 1);
            int acc = 0;
#pragma unroll
            for (int j = -2; j <= 2; ++j) {
                int yy = refl101(gy + j, h);
                int q = (j == 0) ? 136 : (j == 1 || j == -1) ? 56 : 4;
                acc += q * (int)hz[(yy - hz_lo) * w + x];
            }
            g[i] = (u8)((acc * 255 + 32768) >> 16);
        }
        __syncthreads();
        // ---- 0.8x bilinear (INTER_LINEAR_EXACT) -> scaled rows ys0 .. ys0+BR ----
        for (int i = tid; i < (BR + 1) * sw; i += PT) {
            int rs = i / sw, xs = i - rs * sw, ys = ys0 + rs;
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
            sc[i] = val;
        }
        __syncthreads();
        // ---- gradient / defined bits: warp task = (row, word) ----
        const int ntask = band_rows * swp;
        for (int t = warp; t < ntask; t += PT / 32) {
            int rs = t / swp, xw = t - rs * swp, xs = xw * 32 + lane, ys = ys0 + rs;
            bool def = false;
            if (xs < sw - 1 && ys < sh - 1) {
                const u8 *r0 = sc + rs * sw + xs, *r1 = r0 + sw;
                int DA = (int)r1[1] - (int)r0[0], BC = (int)r0[1] - (int)r1[0];
                int gx = DA + BC, gy = DA - BC;
                u32 g2 = (u32)(gx * gx + gy * gy);
                def = g2 >= g2_min;
            }
            u32 bits = __ballot_sync(0xffffffffu, def);
            if (lane == 0) { wbits[t] = bits; wbase[t] = __popc(bits); }
        }
        __syncthreads();
        // ---- exclusive scan of the per-word counts over the band (ntask <= a few hundred) ----
        {
            u32 run = s_run;
            // each thread owns a contiguous chunk
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
        // ---- write words + compact records ----
        for (int t = warp; t < ntask; t += PT / 32) {
            int rs = t / swp, xw = t - rs * swp, xs = xw * 32 + lane, ys = ys0 + rs;
            u32 bits = wbits[t], base = wbase[t];
            if (lane == 0) ow[(size_t)ys * swp + xw] = LsdWord{bits, base};
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
                    p.ang_deg = a;
                    p.c = (float)cos((double)af);   // cosf(float(angle)), correctly rounded
                    p.s = (float)sin((double)af);
                    p.g2 = g2;
                    opix[idx] = p;
                    oxy[idx] = ((u32)ys << 16) | (u32)xs;
                    oused[idx] = 0;
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

static size_t lsd_pre_smem(const Dims &d)
{
    size_t s = (size_t)HZR * d.w * 2 + (size_t)GR * d.w + (((size_t)(BR + 1) * d.sw + 3) & ~(size_t)3) +
               (size_t)HZR * d.wp * 4 + (size_t)BR * d.swp * 8;
    return s + 16;
}

void launch_lsd_pre(const Dims &d, const u32 *planesB, Buffers &b, cudaStream_t st)
{
    size_t smem = lsd_pre_smem(d);
    static size_t attr = 0;
    if (smem > attr) {
        cudaFuncSetAttribute(k_lsd_pre, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = smem;
    }
    // smallest integer g2 with sqrt(g2/4.0) > rho, rho = 2/sin(22.5 deg)   (ll_angle threshold)
    const double rho = 2.0 / sin(3.14159265358979323846 * 22.5 / 180.0);
    u32 g2_min = 0;
    while (!(sqrt((double)g2_min / 4.0) > rho)) ++g2_min;
    k_lsd_pre<<<d.n * 3, PT, smem, st>>>(d, g2_min, planesB, b.lsdw, b.pix, b.pixxy, b.used, b.pixcount, b.g2max, b.flags);
    ++g_launches;
}

}  // namespace lsf
