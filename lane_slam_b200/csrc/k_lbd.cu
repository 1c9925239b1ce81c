// k_lbd.cu -- K11: LBD line descriptors.
//   (a) dense: gray -> GaussianBlur 5x5 sigma 1 (Q8 taps 14,62,104,62,14) -> Sobel 3x3 (CV_16S), reflect-101
//       borders; written as interleaved (dx,dy) int16 pairs so the per-line gather is one 4-byte load.
//       Replaces BinaryDescriptor::computeGaussianPyramid / computeSobel
//       (src/line_descriptor/src/binary_descriptor_custom.cpp:350-398).
//   (b) per line: KeyLine fill (LSDDetector_custom.cpp:76-102,176-197) + computeLBD
//       (binary_descriptor_custom.cpp:1026-1372, weights :217-259) + binaryConversion over the 32
//       band pairs (:401-412, :74-107, :645-667) -> 32 bytes.
// One warp per line: lane h (and h+32) walks row h of the 63-row support region sequentially, so every
// float accumulation happens in the reference's order (bit-identical sums).
#include "common.cuh"
#include "libm_f32.cuh"

namespace lsf {

constexpr int GT_W = 64, GT_H = 32;

__device__ __forceinline__ int refl101b(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
    return i;
}

__global__ void __launch_bounds__(256) k_gray_sobel(Dims d, const u8 *__restrict__ gray, short2 *__restrict__ dxy)
{
    __shared__ u8 g[(GT_H + 6) * (GT_W + 8)];       // gray tile with 3-px halo (row stride 72)
    __shared__ u16 hz[(GT_H + 6) * (GT_W + 2)];     // horizontal pass on (GT_W+2) columns
    __shared__ u8 bl[(GT_H + 2) * (GT_W + 4)];      // blurred tile with 1-px halo (row stride 68)
    const int tid = threadIdx.x, x0 = blockIdx.x * GT_W, y0 = blockIdx.y * GT_H, f = blockIdx.z;
    const u8 *src = gray + (size_t)f * d.h * d.w;
    for (int i = tid; i < (GT_H + 6) * (GT_W + 6); i += 256) {
        int r = i / (GT_W + 6), k = i - r * (GT_W + 6);
        int yy = refl101b(y0 - 3 + r, d.h), xx = refl101b(x0 - 3 + k, d.w);
        g[r * (GT_W + 8) + k] = src[(size_t)yy * d.w + xx];
    }
    __syncthreads();
    for (int i = tid; i < (GT_H + 6) * (GT_W + 2); i += 256) {
        int r = i / (GT_W + 2), k = i - r * (GT_W + 2);   // column k <-> image x0-1+k <-> g column k+2
        const u8 *p = g + r * (GT_W + 8) + k;
        hz[i] = (u16)(14 * p[0] + 62 * p[1] + 104 * p[2] + 62 * p[3] + 14 * p[4]);
    }
    __syncthreads();
    for (int i = tid; i < (GT_H + 2) * (GT_W + 2); i += 256) {
        int r = i / (GT_W + 2), k = i - r * (GT_W + 2);   // row r <-> image y0-1+r <-> hz row r+2
        const u16 *p = hz + r * (GT_W + 2) + k;
        int acc = 14 * p[0] + 62 * p[GT_W + 2] + 104 * p[2 * (GT_W + 2)] + 62 * p[3 * (GT_W + 2)] + 14 * p[4 * (GT_W + 2)];
        bl[r * (GT_W + 4) + k] = (u8)((acc + 32768) >> 16);
    }
    __syncthreads();
    for (int i = tid; i < GT_H * GT_W; i += 256) {
        int ty = i / GT_W, tx = i - ty * GT_W;
        int iy = y0 + ty, ix = x0 + tx;
        if (iy >= d.h || ix >= d.w) continue;
        const u8 *p = bl + (ty + 1) * (GT_W + 4) + tx + 1;
        const int S = GT_W + 4;
        int dx = (p[-S + 1] + 2 * p[1] + p[S + 1]) - (p[-S - 1] + 2 * p[-1] + p[S - 1]);
        int dy = (p[S - 1] + 2 * p[S] + p[S + 1]) - (p[-S - 1] + 2 * p[-S] + p[-S + 1]);
        dxy[((size_t)f * d.h + iy) * d.w + ix] = make_short2((short)dx, (short)dy);
    }
}

// ---- fast path (w % 4 == 0): no shared memory, no barriers -----------------------------------------------------
// A thread owns 4 adjacent columns and marches down GS_R output rows with everything in registers: per input row
// it loads three aligned words (columns x-4 .. x+7, reflect-101 applied to the loaded words at the frame edges),
// forms the six horizontal 5-tap sums it needs (columns x-1 .. x+4), keeps the last five rows of those for the
// vertical 5-tap (-> blurred bytes) and the last three blurred rows for the 3x3 Sobel; one 16-byte store per row.
// The gray image extended by reflect-101 is symmetric about the border rows / columns and the Gaussian kernel is
// symmetric, so blurring the extended image reproduces the reflect-101 border of the blurred image that Sobel
// wants (blurred[-1] == blurred[1]) without a special case.
constexpr int GS_R = 24;

__device__ __forceinline__ int refl101i(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * (n - 1) - i : i); }

__global__ void __launch_bounds__(128) k_gray_sobel4(Dims d, const u8 *__restrict__ gray, short2 *__restrict__ dxy)
{
    const int x = (blockIdx.x * 32 + threadIdx.x) * 4;
    const int y0 = (blockIdx.y * 4 + threadIdx.y) * GS_R;
    const int f = blockIdx.z, w = d.w, h = d.h;
    if (x >= w || y0 >= h) return;
    const u8 *src = gray + (size_t)f * h * w;
    short2 *dst = dxy + (size_t)f * h * w;
    int hz[5][6];      // horizontal sums of the last five input rows, columns x-1 .. x+4
    int bl[3][6];      // blurred values of the last three blurred rows, columns x-1 .. x+4
#pragma unroll
    for (int t = 0; t < GS_R + 6; ++t) {
        // input row of this step; blurred row t-4 <-> image row y0 - 3 + (t - 2) ... see below
        const int gy = refl101i(y0 - 3 + t, h);
        const u32 *rp = reinterpret_cast<const u32 *>(src + (size_t)gy * w + x);
        u32 w1 = rp[0];
        u32 w0 = x > 0 ? rp[-1] : 0u, w2 = x + 4 < w ? rp[1] : 0u;
        if (x == 0) w0 = __byte_perm(w1, w2, 0x1234);           // columns 4,3,2,1
        if (x + 4 >= w) w2 = __byte_perm(w0, w1, 0x3456);       // columns w-2, w-3, w-4, w-5
        // bytes c[0..11] = columns x-4 .. x+7; sum k needs c[k+1 .. k+5]
        int c[12];
#pragma unroll
        for (int k = 0; k < 4; ++k) { c[k] = (w0 >> (8 * k)) & 255; c[4 + k] = (w1 >> (8 * k)) & 255; c[8 + k] = (w2 >> (8 * k)) & 255; }
#pragma unroll
        for (int k = 0; k < 6; ++k) hz[t % 5][k] = 14 * (c[k + 1] + c[k + 5]) + 62 * (c[k + 2] + c[k + 4]) + 104 * c[k + 3];
        if (t >= 4) {
            // blurred row centred on the input row of step t-2, i.e. image row y0 - 1 + (t - 4)
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                int acc = 14 * (hz[(t - 4) % 5][k] + hz[t % 5][k]) + 62 * (hz[(t - 3) % 5][k] + hz[(t - 1) % 5][k]) + 104 * hz[(t - 2) % 5][k];
                bl[(t - 4) % 3][k] = (acc + 32768) >> 16;
            }
        }
        if (t >= 6) {
            const int y = y0 + (t - 6);            // output row: blurred rows y-1, y, y+1 = steps t-2, t-1, t
            if (y < h) {
                const int *a = bl[(t - 6) % 3], *m = bl[(t - 5) % 3], *b = bl[(t - 4) % 3];
                int col[6], dif[6];
#pragma unroll
                for (int k = 0; k < 6; ++k) { col[k] = a[k] + 2 * m[k] + b[k]; dif[k] = b[k] - a[k]; }
                uint4 o;
                u32 *op = &o.x;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    int dx = col[k + 2] - col[k], dy = dif[k] + 2 * dif[k + 1] + dif[k + 2];
                    op[k] = ((u32)dx & 0xffffu) | ((u32)dy << 16);
                }
                *reinterpret_cast<uint4 *>(dst + (size_t)y * w + x) = o;
            }
        }
    }
}

void launch_gray_sobel(const Dims &d, const u8 *gray, short *dx, short *, cudaStream_t st)
{
    if ((d.w & 3) == 0 && d.w >= 8 && d.h >= 8) {
        dim3 grid((d.w + 127) / 128, (d.h + 4 * GS_R - 1) / (4 * GS_R), d.n), block(32, 4);
        k_gray_sobel4<<<grid, block, 0, st>>>(d, gray, reinterpret_cast<short2 *>(dx));
    } else {
        dim3 grid((d.w + GT_W - 1) / GT_W, (d.h + GT_H - 1) / GT_H, d.n);
        k_gray_sobel<<<grid, 256, 0, st>>>(d, gray, reinterpret_cast<short2 *>(dx));
    }
    ++g_launches;
}

// ---- per-line LBD ---------------------------------------------------------------------------------------
__constant__ float c_gaussL[21];
__constant__ float c_gaussG[63];
__constant__ unsigned char c_comb[32][2];
static PerDevice g_lbd_once;

static void ensure_lbd_tables()
{
    g_lbd_once.ensure(1, [] {
    float gl[21], gg[63];
    // integer divisions are intentional (binary_descriptor_custom.cpp:227-258)
    double u = (7 * 3 - 1) / 2, sigma = (7 * 2 + 1) / 2, inv = -1 / (2 * sigma * sigma);
    for (int i = 0; i < 21; ++i) { double dis = i - u; gl[i] = (float)exp(dis * dis * inv); }
    u = (9 * 7 - 1) / 2; sigma = u; inv = -1 / (2 * sigma * sigma);
    for (int i = 0; i < 63; ++i) { double dis = i - u; gg[i] = (float)exp(dis * dis * inv); }
    static const unsigned char comb[32][2] = {
        {0, 1}, {0, 2}, {0, 3}, {0, 4}, {0, 5}, {0, 6}, {1, 2}, {1, 3}, {1, 4}, {1, 5}, {1, 6}, {2, 3}, {2, 4}, {2, 5}, {2, 6}, {2, 7},
        {2, 8}, {3, 4}, {3, 5}, {3, 6}, {3, 7}, {3, 8}, {4, 5}, {4, 6}, {4, 7}, {4, 8}, {5, 6}, {5, 7}, {5, 8}, {6, 7}, {6, 8}, {7, 8}};
    cudaMemcpyToSymbol(c_gaussL, gl, sizeof(gl));
    cudaMemcpyToSymbol(c_gaussG, gg, sizeof(gg));
    cudaMemcpyToSymbol(c_comb, comb, sizeof(comb));
    });
}

constexpr int LBD_WARPS = 4;

__global__ void __launch_bounds__(LBD_WARPS * 32, 8) k_lbd(Dims d, const float *__restrict__ lines, const int *__restrict__ frame_of_seg,
                                                       int nseg_cap, const int *__restrict__ seg_lo_dev,
                                                       const int *__restrict__ seg_hi_dev, int *__restrict__ cursor,
                                                       const short2 *__restrict__ dxy, u8 *__restrict__ desc)
{
    __shared__ float rows[LBD_WARPS][63][4];   // per-row sums scaled by the global Gaussian weight
    __shared__ float dvec[LBD_WARPS][72];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int seg_lo = seg_lo_dev ? *seg_lo_dev : 0;
    const int nseg = min(*seg_hi_dev, nseg_cap);
    const int W = d.w, H = d.h;
    // lines differ a lot in length: warps pull them one at a time
    while (true) {
        int sidx = 0;
        if (lane == 0) sidx = seg_lo + atomicAdd(cursor, 1);
        sidx = __shfl_sync(0xffffffffu, sidx, 0);
        if (sidx >= nseg) break;
        // ---- KeyLine fill (octave 0) ----
        float4 ln = reinterpret_cast<const float4 *>(lines)[sidx];
        float e0 = ln.x, e1 = ln.y, e2 = ln.z, e3 = ln.w;
        if (e0 < 0) e0 = 0;
        if (e0 >= W) e0 = (float)W - 1.0f;
        if (e2 < 0) e2 = 0;
        if (e2 >= W) e2 = (float)W - 1.0f;
        if (e1 < 0) e1 = 0;
        if (e1 >= H) e1 = (float)H - 1.0f;
        if (e3 < 0) e3 = 0;
        if (e3 >= H) e3 = (float)H - 1.0f;
        const float direction = lmf_atan2f(__fsub_rn(e3, e1), __fsub_rn(e2, e0));   // == glibc atan2f (libm_f32.cuh)
        int px0 = __float2int_rn(e0), py0 = __float2int_rn(e1), px1 = __float2int_rn(e2), py1 = __float2int_rn(e3);
        int len = max(abs(px1 - px0), abs(py1 - py0)) + 1;  // LineIterator(8-connected).count
        const short2 *img = dxy + (size_t)frame_of_seg[sidx] * H * W;
        // ---- support region walk ----
        const short halfHeight = 31;
        const short halfWidth = (short)((len - 1) / 2);
        const float mx = (float)(0.5 * (double)(e0 + e2)), my = (float)(0.5 * (double)(e1 + e3));
        const float dL0 = lmf_cosf(direction), dL1 = lmf_sinf(direction);           // == glibc cosf / sinf
        const float dO0 = -dL1, dO1 = dL0;
        float s0x = __fadd_rn(__fadd_rn(__fmul_rn(-dL0, (float)halfWidth), __fmul_rn(dL1, (float)halfHeight)), mx);
        float s0y = __fadd_rn(__fsub_rn(__fmul_rn(-dL1, (float)halfWidth), __fmul_rn(dL0, (float)halfHeight)), my);
        for (int rnd = 0; rnd < 2; ++rnd) {
            int hID = lane + 32 * rnd;
            if (hID < 63) {
                // the reference advances the row origin iteratively: replicate the float32 recurrence
                float ox = s0x, oy = s0y;
                for (int t = 0; t < hID; ++t) { ox = __fsub_rn(ox, dL1); oy = __fadd_rn(oy, dL0); }
                float sx = ox, sy = oy;
                float pL = 0, nL = 0, pO = 0, nO = 0;
                // four gathers in flight per lane; the float sums are still added in walking order
                int wID = 0;
                for (; wID + 8 <= len; wID += 8) {
                    short2 gq[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        int tx = (int)(short)__float2int_rn(roundf(sx));
                        int ty = (int)(short)__float2int_rn(roundf(sy));
                        int xc = tx < 0 ? 0 : (tx > W - 1 ? W - 1 : tx);
                        int yc = ty < 0 ? 0 : (ty > H - 1 ? H - 1 : ty);
                        gq[u] = img[(size_t)yc * W + xc];
                        sx = __fadd_rn(sx, dL0);
                        sy = __fadd_rn(sy, dL1);
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        float gDL = __fadd_rn(__fmul_rn((float)gq[u].x, dL0), __fmul_rn((float)gq[u].y, dL1));
                        float gDO = __fadd_rn(__fmul_rn((float)gq[u].x, dO0), __fmul_rn((float)gq[u].y, dO1));
                        if (gDL > 0) pL = __fadd_rn(pL, gDL); else nL = __fsub_rn(nL, gDL);
                        if (gDO > 0) pO = __fadd_rn(pO, gDO); else nO = __fsub_rn(nO, gDO);
                    }
                }
                for (; wID < len; ++wID) {
                    int tx = (int)(short)__float2int_rn(roundf(sx));
                    int ty = (int)(short)__float2int_rn(roundf(sy));
                    int xc = tx < 0 ? 0 : (tx > W - 1 ? W - 1 : tx);
                    int yc = ty < 0 ? 0 : (ty > H - 1 ? H - 1 : ty);
                    short2 gq = img[(size_t)yc * W + xc];
                    float gDL = __fadd_rn(__fmul_rn((float)gq.x, dL0), __fmul_rn((float)gq.y, dL1));
                    float gDO = __fadd_rn(__fmul_rn((float)gq.x, dO0), __fmul_rn((float)gq.y, dO1));
                    if (gDL > 0) pL = __fadd_rn(pL, gDL); else nL = __fsub_rn(nL, gDL);
                    if (gDO > 0) pO = __fadd_rn(pO, gDO); else nO = __fsub_rn(nO, gDO);
                    sx = __fadd_rn(sx, dL0);
                    sy = __fadd_rn(sy, dL1);
                }
                float cg = c_gaussG[hID];
                rows[warp][hID][0] = __fmul_rn(cg, pL);
                rows[warp][hID][1] = __fmul_rn(cg, nL);
                rows[warp][hID][2] = __fmul_rn(cg, pO);
                rows[warp][hID][3] = __fmul_rn(cg, nO);
            }
        }
        __syncwarp();
        // ---- band accumulation: item = (band, quantity); rows visited in ascending hID like the reference ----
        for (int item = lane; item < 72; item += 32) {
            int b = item >> 3, q = item & 7;       // q 0..3: sums (pgdL, ngdL, pgdO, ngdO); 4..7: squares
            int comp = q & 3;
            bool sq = q >= 4;
            float acc = 0;
            int h0 = max(0, 7 * (b - 1)), h1 = min(63, 7 * (b + 2));
            for (int hID = h0; hID < h1; ++hID) {
                int hb = hID / 7;
                float coef = hb == b ? c_gaussL[hID % 7 + 7] : (hb == b + 1 ? c_gaussL[hID % 7 + 14] : c_gaussL[hID % 7]);
                float v = rows[warp][hID][comp];
                if (sq) acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(coef, coef), __fmul_rn(v, v)));
                else acc = __fadd_rn(acc, __fmul_rn(coef, v));
            }
            dvec[warp][item] = acc;   // raw band sums; layout [b][0..3 sums, 4..7 squares]
        }
        __syncwarp();
        // ---- mean / std, normalisation, clamp, renormalise: scalar float chain on every lane (identical) ----
        float des[72];
        {
            const float invN2 = (float)(1.0 / (7 * 2.0)), invN3 = (float)(1.0 / (7 * 3.0));
            for (int b = 0; b < 9; ++b) {
                float invN = (b == 0 || b == 8) ? invN2 : invN3;
                for (int j = 0; j < 4; ++j) {
                    float temp = __fmul_rn(dvec[warp][8 * b + j], invN);
                    des[8 * b + j] = temp;
                    des[8 * b + 4 + j] = sqrtf(__fsub_rn(__fmul_rn(dvec[warp][8 * b + 4 + j], invN), __fmul_rn(temp, temp)));
                }
            }
            float tempM = 0, tempS = 0;
            for (int b = 0; b < 9; ++b) {
                for (int j = 0; j < 4; ++j) tempM = __fadd_rn(tempM, __fmul_rn(des[8 * b + j], des[8 * b + j]));
                for (int j = 4; j < 8; ++j) tempS = __fadd_rn(tempS, __fmul_rn(des[8 * b + j], des[8 * b + j]));
            }
            tempM = __fdiv_rn(1.f, sqrtf(tempM));
            tempS = __fdiv_rn(1.f, sqrtf(tempS));
            for (int b = 0; b < 9; ++b) {
                for (int j = 0; j < 4; ++j) des[8 * b + j] = __fmul_rn(des[8 * b + j], tempM);
                for (int j = 4; j < 8; ++j) des[8 * b + j] = __fmul_rn(des[8 * b + j], tempS);
            }
            for (int i = 0; i < 72; ++i)
                if ((double)des[i] > 0.4) des[i] = (float)0.4;
            float temp = 0;
            for (int i = 0; i < 72; ++i) temp = __fadd_rn(temp, __fmul_rn(des[i], des[i]));
            temp = __fdiv_rn(1.f, sqrtf(temp));
            for (int i = 0; i < 72; ++i) des[i] = __fmul_rn(des[i], temp);
        }
        // ---- binary conversion: lane c compares band pair c ----
        {
            const int c0 = c_comb[lane][0], c1 = c_comb[lane][1];
            unsigned r = 0;
            // publish the final 72 floats through shared memory (lane-dependent indexing)
            __syncwarp();
            for (int i = lane; i < 72; i += 32) dvec[warp][i] = des[i];
            __syncwarp();
            for (int i = 0; i < 8; ++i)
                if (dvec[warp][8 * c0 + i] > dvec[warp][8 * c1 + i]) r += 1u << i;
            desc[(size_t)sidx * 32 + lane] = (u8)r;
        }
        __syncwarp();
    }
}

void launch_lbd(const Dims &d, const float *lines, const int *frame_of_seg, int nseg_cap, const int *seg_lo_dev,
                const int *seg_hi_dev, const short *dx, const short *, u8 *desc, int *cursor, cudaStream_t st)
{
    ensure_lbd_tables();
    int grid = 148 * 8;
    if (nseg_cap < grid * LBD_WARPS) grid = (nseg_cap + LBD_WARPS - 1) / LBD_WARPS;
    if (grid < 1) grid = 1;
    k_lbd<<<grid, LBD_WARPS * 32, 0, st>>>(d, lines, frame_of_seg, nseg_cap, seg_lo_dev, seg_hi_dev, cursor, reinterpret_cast<const short2 *>(dx), desc);
    ++g_launches;
}

}  // namespace lsf
