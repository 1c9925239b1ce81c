// k_color_canny.cu -- K1+K2+K3: resize/crop/colour-correct, BGR->HSV colour masks, BGR->GRAY,
// 3-channel Canny gradient + non-maximum suppression.  One pass over the frame: 3N bytes in,
// 5 bit-planes (0.625N bytes) + N bytes gray out.
//
// Replaces (reference, relative to /root/reference):
//   src/line_detector/src/line_detector_node.py:163-175   resize INTER_NEAREST, crop, AntiInstagram, convertScaleAbs
//   src/line_detector/include/line_detector/line_detector_lsd.py:137-139  cvtColor BGR2GRAY / BGR2HSV, Canny(bgr)
//   .../line_detector_lsd.py:40-47  inRange (x4)
// Arithmetic = OpenCV 4.13 fixed-point models.
//
// Tile 128x32 pixels + halo, staged into shared memory by one TMA 3-D box load of uint32 elements
// ([frame][row][word], 104 words x 36 rows, zero fill outside the frame) when there is no resize; a gather
// loader otherwise.  Each thread handles runs of 4 pixels with 32-bit shared-memory loads.
#include <algorithm>
#include "common.cuh"

namespace lsf {

constexpr int TW = 128, TH = 32, HALO = 2;
constexpr int XOFF = 16;                 // bytes of left padding: TMA needs a 16-byte aligned inner start (measured)
constexpr int ROWB = 416;                // bytes per tile row = 104 uint32: columns -5 .. 132 of the tile
constexpr int BOX_Y = TH + 2 * HALO;     // 36
constexpr int NRUN = TW / 4 + 2;         // 34 runs of 4 columns covering tile columns -4 .. 131
constexpr int MAGW = 4 * NRUN;           // 136 u16 per magnitude row; column c lives at c + 4
constexpr int NT = 256;

__constant__ int c_sdiv[256];
__constant__ int c_hdiv[256];
static PerDevice g_tabs_once;

// sdiv / hdiv for the marching kernel (the tiled kernel reads the per-context table, build_color_tables)
static void ensure_tables()
{
    g_tabs_once.ensure(1, [] {
    int sdiv[256], hdiv[256];
    sdiv[0] = hdiv[0] = 0;
    for (int i = 1; i < 256; ++i) {
        sdiv[i] = (int)lrint((255 << 12) / (1.0 * i));
        hdiv[i] = (int)lrint((180 << 12) / (6.0 * i));
    }
    cudaMemcpyToSymbol(c_sdiv, sdiv, sizeof(sdiv));
    cudaMemcpyToSymbol(c_hdiv, hdiv, sizeof(hdiv));
    });
}

__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }

// byte k (0..19) of a 5-word window
#define WB(W, k) (int)(((W)[(k) >> 2] >> (8 * ((k) & 3))) & 0xffu)

// HSV tables of a context, built once on the host: [sdiv 256 x i32][hdiv 256 x i32][lutH 256][lutS 256][lutV 256]
// (lut bit i: value inside colour range i = white, yellow, red1, red2)
void build_color_tables(const ColorParams &cp, u8 *dst /* COLOR_TABLE_BYTES */)
{
    int *sdiv = reinterpret_cast<int *>(dst), *hdiv = sdiv + 256;
    u8 *lh = dst + 2048, *ls = lh + 256, *lv = ls + 256;
    sdiv[0] = hdiv[0] = 0;
    for (int i = 1; i < 256; ++i) {
        sdiv[i] = (int)lrint((255 << 12) / (1.0 * i));
        hdiv[i] = (int)lrint((180 << 12) / (6.0 * i));
    }
    for (int v = 0; v < 256; ++v) {
        int mh = 0, ms = 0, mv = 0;
        for (int i = 0; i < 4; ++i) {
            mh |= (v >= cp.lo[i][0] && v <= cp.hi[i][0]) << i;
            ms |= (v >= cp.lo[i][1] && v <= cp.hi[i][1]) << i;
            mv |= (v >= cp.lo[i][2] && v <= cp.hi[i][2]) << i;
        }
        lh[v] = (u8)mh; ls[v] = (u8)ms; lv[v] = (u8)mv;
    }
}

template <bool USE_TMA>
__global__ void __launch_bounds__(NT) k_color_canny(const __grid_constant__ CUtensorMap tmap, Dims d, ColorParams cp,
                                                    const u8 *__restrict__ src, const u8 *__restrict__ tables,
                                                    u32 *__restrict__ planesA, u8 *__restrict__ gray)
{
    __shared__ __align__(128) u8 tile[BOX_Y * ROWB];
    __shared__ __align__(16) u16 mag[(TH + 2) * MAGW];
    __shared__ __align__(16) u32 dxy[TH * TW];        // (dx & 0xffff) | (dy << 16) of the selected channel
    __shared__ int s_sdiv[256], s_hdiv[256];
    __shared__ u8 s_lutH[256], s_lutS[256], s_lutV[256];   // bit i: value inside colour range i (white, yellow, red1, red2)
    __shared__ __align__(8) u64 bar;

    const int tid = threadIdx.x;
    const int tx0 = blockIdx.x * TW, ty0 = blockIdx.y * TH, f = blockIdx.z;
    if (USE_TMA) {
        const u32 bar_a = smem_u32(&bar);
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_a), "r"(1));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(ROWB * BOX_Y)
                         : "memory");
            int c0 = (tx0 * 3 - XOFF) / 4, c1 = ty0 - HALO + d.top, c2 = f + d.f0;   // uint32 elements
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                ::"r"(smem_u32(tile)), "l"(&tmap), "r"(c0), "r"(c1), "r"(c2), "r"(bar_a)
                : "memory");
        }
        {   // HSV tables while the tile is in flight
            const int *ti = reinterpret_cast<const int *>(tables);
            s_sdiv[tid] = ti[tid];
            s_hdiv[tid] = ti[256 + tid];
            s_lutH[tid] = tables[2048 + tid]; s_lutS[tid] = tables[2304 + tid]; s_lutV[tid] = tables[2560 + tid];
        }
        u32 done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(bar_a), "r"(0)
                : "memory");
        }
    } else {
        {
            const int *ti = reinterpret_cast<const int *>(tables);
            s_sdiv[tid] = ti[tid];
            s_hdiv[tid] = ti[256 + tid];
            s_lutH[tid] = tables[2048 + tid]; s_lutS[tid] = tables[2304 + tid]; s_lutV[tid] = tables[2560 + tid];
        }
        const u8 *fsrc = src + (size_t)f * d.src_frame;
        for (int i = tid; i < BOX_Y * ROWB; i += NT) {
            int r = i / ROWB, k = i - r * ROWB;
            int yy = ty0 - HALO + r, xb = tx0 * 3 - XOFF + k;
            u8 v = 0;
            if (yy >= 0 && yy < d.h && xb >= 0 && xb < d.w * 3) {
                int px = xb / 3, c = xb - px * 3;
                int sy = nearest_src(yy + d.top, d.src_h, d.dh), sx = nearest_src(px, d.src_w, d.dw);
                v = fsrc[(size_t)sy * d.src_pitch + (size_t)sx * 3 + c];
            }
            tile[i] = v;
        }
        __syncthreads();
    }
    if (!d.identity_color) {
        if (USE_TMA) __syncthreads();
        for (int i = tid; i < BOX_Y * ROWB; i += NT) {
            int c = ((i % ROWB) + 2) % 3;  // tile row starts at byte 3*tx0 - 16: channel = (k - 16) mod 3
            tile[i] = color_correct(tile[i], cp.ai_scale[c], cp.ai_shift[c]);
        }
    }
    __syncthreads();

    // ---- phase 1: per-channel Sobel, L1 magnitude, first maximal channel.  Task = (row, run of 4 columns) ----
    for (int t = tid; t < (TH + 2) * NRUN; t += NT) {
        const int r = t / NRUN, q = t - r * NRUN;      // tile row -1 + r, tile columns 4q-4 .. 4q-1
        const int iy = ty0 - 1 + r, ix0 = tx0 + 4 * q - 4;
        int best[4] = {0, 0, 0, 0}, bdx[4] = {0, 0, 0, 0}, bdy[4] = {0, 0, 0, 0};
        if (iy >= 0 && iy < d.h && ix0 + 3 >= 0 && ix0 < d.w) {
            // BORDER_REPLICATE in y: clamp the neighbour rows to the image
            const int rm = max(iy - 1, 0) - (ty0 - HALO), r0 = iy - (ty0 - HALO), rp = min(iy + 1, d.h - 1) - (ty0 - HALO);
            const bool wal = (d.w & 3) == 0;       // runs line up with the image borders
            if ((ix0 >= 1 && ix0 + 4 <= d.w - 1) || (wal && ix0 >= 0 && ix0 + 4 <= d.w)) {
                // fast path: 5 words per row hold columns ix0-1 .. ix0+4 (window bytes 1 .. 18).  The first / last run of an
                // image row gets its outside column (BORDER_REPLICATE) by one byte permute per row.
                const u32 *w0 = reinterpret_cast<const u32 *>(tile + rm * ROWB) + 3 * q;
                const u32 *w1 = reinterpret_cast<const u32 *>(tile + r0 * ROWB) + 3 * q;
                const u32 *w2 = reinterpret_cast<const u32 *>(tile + rp * ROWB) + 3 * q;
                u32 T[5], M[5], B[5];
#pragma unroll
                for (int i = 0; i < 5; ++i) { T[i] = w0[i]; M[i] = w1[i]; B[i] = w2[i]; }
                if (ix0 == 0) {                    // column -1 := column 0 (bytes 1..3 := bytes 4..6)
                    T[0] = __byte_perm(T[0], T[1], 0x6540); M[0] = __byte_perm(M[0], M[1], 0x6540); B[0] = __byte_perm(B[0], B[1], 0x6540);
                }
                if (ix0 + 4 == d.w) {              // column w := column w-1 (bytes 16..18 := bytes 13..15)
                    T[4] = __byte_perm(T[3], T[4], 0x7321); M[4] = __byte_perm(M[3], M[4], 0x7321); B[4] = __byte_perm(B[3], B[4], 0x7321);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) best[i] = -1;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    int V[6], D[6];
#pragma unroll
                    for (int j = 0; j < 6; ++j) {
                        int tt = WB(T, 1 + 3 * j + c), mm = WB(M, 1 + 3 * j + c), bb = WB(B, 1 + 3 * j + c);
                        V[j] = tt + 2 * mm + bb;
                        D[j] = bb - tt;
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        int dx = V[i + 2] - V[i], dy = D[i] + 2 * D[i + 1] + D[i + 2];
                        int nrm = abs(dx) + abs(dy);
                        if (nrm > best[i]) { best[i] = nrm; bdx[i] = dx; bdy[i] = dy; }
                    }
                }
            } else {
                // image border: clamp every neighbour column
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    int ix = ix0 + i;
                    if (ix < 0 || ix >= d.w) continue;
                    int cm = (max(ix - 1, 0) - tx0) * 3 + XOFF, c0 = (ix - tx0) * 3 + XOFF, cq = (min(ix + 1, d.w - 1) - tx0) * 3 + XOFF;
                    const u8 *p0 = tile + rm * ROWB, *p1 = tile + r0 * ROWB, *p2 = tile + rp * ROWB;
                    best[i] = -1;
                    for (int c = 0; c < 3; ++c) {
                        int a00 = p0[cm + c], a01 = p0[c0 + c], a02 = p0[cq + c];
                        int a10 = p1[cm + c], a12 = p1[cq + c];
                        int a20 = p2[cm + c], a21 = p2[c0 + c], a22 = p2[cq + c];
                        int dx = (a02 + 2 * a12 + a22) - (a00 + 2 * a10 + a20);
                        int dy = (a20 + 2 * a21 + a22) - (a00 + 2 * a01 + a02);
                        int nrm = abs(dx) + abs(dy);
                        if (nrm > best[i]) { best[i] = nrm; bdx[i] = dx; bdy[i] = dy; }
                    }
                }
            }
        }
        // magnitudes outside the image stay 0
        uint2 mg = make_uint2((u32)best[0] | ((u32)best[1] << 16), (u32)best[2] | ((u32)best[3] << 16));
        *reinterpret_cast<uint2 *>(&mag[r * MAGW + 4 * q]) = mg;
        if (r >= 1 && r <= TH && q >= 1 && q <= TW / 4) {
            uint4 v;
            v.x = ((u32)bdx[0] & 0xffffu) | ((u32)bdy[0] << 16); v.y = ((u32)bdx[1] & 0xffffu) | ((u32)bdy[1] << 16);
            v.z = ((u32)bdx[2] & 0xffffu) | ((u32)bdy[2] << 16); v.w = ((u32)bdx[3] & 0xffffu) | ((u32)bdy[3] << 16);
            *reinterpret_cast<uint4 *>(&dxy[(r - 1) * TW + 4 * (q - 1)]) = v;
        }
    }
    __syncthreads();

    // ---- phase 2: NMS + thresholds, HSV colour masks, gray.  Warp = one row, lane = 4 consecutive pixels ----
    const int warp = tid >> 5, lane = tid & 31;
    for (int ty = warp; ty < TH; ty += NT / 32) {
        const int iy = ty0 + ty;
        if (iy >= d.h) break;  // warp-uniform
        u32 nib[PA_COUNT] = {0, 0, 0, 0, 0};
        u32 gpack = 0;
        const uint4 dv = *reinterpret_cast<const uint4 *>(&dxy[ty * TW + 4 * lane]);
        const u32 dvs[4] = {dv.x, dv.y, dv.z, dv.w};
        const u32 *pw = reinterpret_cast<const u32 *>(tile + (ty + HALO) * ROWB + XOFF) + 3 * lane;
        const u32 P[3] = {pw[0], pw[1], pw[2]};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int tx = 4 * lane + i, ix = tx0 + tx;
            if (ix >= d.w) continue;
            const u16 *m = mag + (ty + 1) * MAGW + tx + 4;
            int c = m[0];
            if (c > cp.canny_lo) {
                int xs = (int)(short)(dvs[i] & 0xffffu), ys = (int)dvs[i] >> 16;
                int ax = abs(xs);
                long long ay = (long long)abs(ys) << 15;
                long long t22 = (long long)ax * 13573, t67 = t22 + ((long long)ax << 16);
                bool ismax;
                if (ay < t22) ismax = c > m[-1] && c >= m[1];
                else if (ay > t67) ismax = c > m[-MAGW] && c >= m[MAGW];
                else {
                    int s = ((xs ^ ys) < 0) ? -1 : 1;
                    ismax = c > m[-MAGW - s] && c > m[MAGW + s];
                }
                if (ismax) {
                    nib[PA_CAND] |= 1u << i;
                    if (c > cp.canny_hi) nib[PA_STRONG] |= 1u << i;
                }
            }
            const int b = WB(P, 3 * i), g = WB(P, 3 * i + 1), r = WB(P, 3 * i + 2);
            // colour ranges: test V first, then S, then H -- most pixels fail on V (or S) and skip the hue maths
            int v = max(b, max(g, r));
            u32 in = s_lutV[v];
            if (in) {
                int mn = min(b, min(g, r)), diff = v - mn;
                int s = (diff * s_sdiv[v] + 2048) >> 12;
                in &= s_lutS[s];
                if (in) {
                    int hh = (v == r) ? (g - b) : (v == g) ? (b - r + 2 * diff) : (r - g + 4 * diff);
                    hh = (hh * s_hdiv[diff] + 2048) >> 12;
                    if (hh < 0) hh += 180;
                    in &= s_lutH[hh];
                    nib[PA_RAW_W] |= (in & 1u) << i;
                    nib[PA_RAW_Y] |= ((in >> 1) & 1u) << i;
                    nib[PA_RAW_R] |= (((in >> 2) | (in >> 3)) & 1u) << i;
                }
            }
            gpack |= (u32)((b * 3735 + g * 19235 + r * 9798 + 16384) >> 15) << (8 * i);
        }
        if (gray) {
            const int ix = tx0 + 4 * lane;
            u8 *gp = gray + ((size_t)f * d.h + iy) * d.w + ix;
            if (ix + 3 < d.w && (d.w & 3) == 0) *reinterpret_cast<u32 *>(gp) = gpack;
            else
                for (int i = 0; i < 4 && ix + i < d.w; ++i) gp[i] = (u8)(gpack >> (8 * i));
        }
        // 8 lanes x 4 bits -> one 32-bit plane word (OR butterfly inside each group of 8 lanes)
#pragma unroll
        for (int pl = 0; pl < PA_COUNT; ++pl) {
            u32 v = nib[pl] << (4 * (lane & 7));
            v |= __shfl_xor_sync(0xffffffffu, v, 1);
            v |= __shfl_xor_sync(0xffffffffu, v, 2);
            v |= __shfl_xor_sync(0xffffffffu, v, 4);
            nib[pl] = v;
        }
        if ((lane & 7) < PA_COUNT) {
            const int pl = lane & 7, xw = (tx0 >> 5) + (lane >> 3);
            u32 val = pl == 0 ? nib[0] : pl == 1 ? nib[1] : pl == 2 ? nib[2] : pl == 3 ? nib[3] : nib[4];
            if (xw < d.wp) planesA[(((size_t)f * PA_COUNT + pl) * d.h + iy) * d.wp + xw] = val;
        }
    }
}

// ---- v3 ------------------------------------------------------------------------------------------------------------------
// Same tile, same outputs, bit-identical results, ~4x fewer instructions per pixel on lane frames than the kernel above
// (ncu r02a: 198 thread-instructions per pixel, 49 % in phase 1, 46 % in phase 2; the kernel is issue bound):
//   phase 1  task = (row, run of 8 columns).  (a) FLAT TEST on bytes: VABSDIFF4 of the run's 3 x 10-pixel window against its
//            centre pixel, thresholded with one masked add and OR-ed into an accumulator (4 instructions per word).  If
//            every byte lies within +-T of the centre pixel's channel value, the per-channel range of the window is
//            R <= 2T and no Sobel L1 magnitude inside it can exceed 6R (the weights of |dx| + |dy| for any sign pattern sum
//            to +6 / -6), so with T = canny_lo / 12 none of the 8 pixels can be a Canny candidate: their magnitudes are
//            stored as 0 -- a candidate neighbour p compares its own c_p > canny_lo >= m_q against them, so the outcome of
//            every NMS comparison is unchanged.  (b) the other runs go to a shared-memory work list and are processed
//            densely: bytes widened to 16-bit lanes (PRMT), vertical sums / differences, |dx| by max - min, dy with one
//            funnel per register, channel maximum -- all in 16x2 integer SIMD (VIADD.16x2, VIMNMX.[SU]16x2).  Only the
//            magnitude is kept; (c) runs that touch the left / right image border take a scalar path (BORDER_REPLICATE).
//   phase 2  task = (row, 8 columns).  Quick rejects on whole words: no magnitude above canny_lo -> no NMS work; no byte
//            >= the smallest V threshold of any colour -> no HSV work.  Candidates (a few % of the pixels) recompute the
//            signed gradient of the first maximal channel from the tile.  Gray by IDP.2A.  The five 8-bit masks of four
//            neighbouring lanes become plane words with a 4 x 4 byte transpose (2 shuffles).
constexpr int NRUN8 = 17;                // runs per tile row: run q = tile columns 8q-1 .. 8q+6
constexpr int MAGW3 = 136;               // u16 per magnitude row; tile column c at index c + 1
constexpr int NTASK1 = (TH + 2) * NRUN8; // 578

__device__ __forceinline__ u32 vmaxu2(u32 a, u32 b) { return __vmaxu2(a, b); }
__device__ __forceinline__ u32 vminu2(u32 a, u32 b) { return __vminu2(a, b); }

// Sobel L1 magnitude (maximum over the channels) + signed gradient of the first maximal channel at tile pixel (row r of the
// tile buffer, image column ix), BORDER_REPLICATE in x through clamped columns; rows rm / rp are already clamped.
__device__ __forceinline__ int sobel_scalar(const u8 *tile, int rm, int r0, int rp, int ix, int tx0, int w, int &bdx, int &bdy)
{
    const int cm = (max(ix - 1, 0) - tx0) * 3 + XOFF, c0 = (ix - tx0) * 3 + XOFF, cq = (min(ix + 1, w - 1) - tx0) * 3 + XOFF;
    const u8 *p0 = tile + rm * ROWB, *p1 = tile + r0 * ROWB, *p2 = tile + rp * ROWB;
    int best = -1;
    bdx = bdy = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int a00 = p0[cm + c], a01 = p0[c0 + c], a02 = p0[cq + c];
        const int a10 = p1[cm + c], a12 = p1[cq + c];
        const int a20 = p2[cm + c], a21 = p2[c0 + c], a22 = p2[cq + c];
        const int dx = (a02 + 2 * a12 + a22) - (a00 + 2 * a10 + a20);
        const int dy = (a20 + 2 * a21 + a22) - (a00 + 2 * a01 + a02);
        const int nrm = abs(dx) + abs(dy);
        if (nrm > best) { best = nrm; bdx = dx; bdy = dy; }
    }
    return best;
}

template <bool USE_TMA>
__global__ void __launch_bounds__(NT, 3) k_color_canny3(const __grid_constant__ CUtensorMap tmap, Dims d, ColorParams cp,
                                                       const u8 *__restrict__ src, const u8 *__restrict__ tables,
                                                       u32 *__restrict__ planesA, u8 *__restrict__ gray, int vmin)
{
    __shared__ __align__(128) u8 tile[BOX_Y * ROWB + 32];
    __shared__ __align__(16) u16 mag[(TH + 2) * MAGW3];
    __shared__ int s_sdiv[256], s_hdiv[256];
    __shared__ u8 s_lutH[256], s_lutS[256], s_lutV[256];
    __shared__ u16 s_work[NTASK1];
    __shared__ u16 s_cand[TH * TW];          // pixels above canny_lo: (row << 7) | column
    __shared__ u32 s_cbits[2 * TH * 4];      // NMS survivors / strong survivors as bit rows of the tile
    __shared__ int s_nwork, s_ncand;
    __shared__ __align__(8) u64 bar;

    const int tid = threadIdx.x, lane = tid & 31;
    const int tx0 = blockIdx.x * TW, ty0 = blockIdx.y * TH, f = blockIdx.z;
    if (tid == 0) { s_nwork = 0; s_ncand = 0; }
    s_cbits[tid] = 0;                         // 2 * TH * 4 == NT words
    if (USE_TMA) {
        const u32 bar_a = smem_u32(&bar);
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_a), "r"(1));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(ROWB * BOX_Y) : "memory");
            int c0 = (tx0 * 3 - XOFF) / 4, c1 = ty0 - HALO + d.top, c2 = f + d.f0;
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                ::"r"(smem_u32(tile)), "l"(&tmap), "r"(c0), "r"(c1), "r"(c2), "r"(bar_a)
                : "memory");
        }
        {
            const int *ti = reinterpret_cast<const int *>(tables);
            s_sdiv[tid] = ti[tid];
            s_hdiv[tid] = ti[256 + tid];
            s_lutH[tid] = tables[2048 + tid]; s_lutS[tid] = tables[2304 + tid]; s_lutV[tid] = tables[2560 + tid];
        }
        u32 done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(bar_a), "r"(0)
                : "memory");
        }
    } else {
        {
            const int *ti = reinterpret_cast<const int *>(tables);
            s_sdiv[tid] = ti[tid];
            s_hdiv[tid] = ti[256 + tid];
            s_lutH[tid] = tables[2048 + tid]; s_lutS[tid] = tables[2304 + tid]; s_lutV[tid] = tables[2560 + tid];
        }
        const u8 *fsrc = src + (size_t)f * d.src_frame;
        for (int i = tid; i < BOX_Y * ROWB; i += NT) {
            int r = i / ROWB, k = i - r * ROWB;
            int yy = ty0 - HALO + r, xb = tx0 * 3 - XOFF + k;
            u8 v = 0;
            if (yy >= 0 && yy < d.h && xb >= 0 && xb < d.w * 3) {
                int px = xb / 3, c = xb - px * 3;
                int sy = nearest_src(yy + d.top, d.src_h, d.dh), sx = nearest_src(px, d.src_w, d.dw);
                v = fsrc[(size_t)sy * d.src_pitch + (size_t)sx * 3 + c];
            }
            tile[i] = v;
        }
        __syncthreads();
    }
    if (!d.identity_color) {
        if (USE_TMA) __syncthreads();
        for (int i = tid; i < BOX_Y * ROWB; i += NT) {
            int c = ((i % ROWB) + 2) % 3;
            tile[i] = color_correct(tile[i], cp.ai_scale[c], cp.ai_shift[c]);
        }
    }
    __syncthreads();

    const int lo = cp.canny_lo, hi = cp.canny_hi;
    // ---- phase 1a: flat test / border runs / work list ----
    {
        const int T = min(lo / 12, 127);
        const u32 K = 0x01010101u * (u32)(0x7f - T);
        for (int t = tid; t < NTASK1; t += NT) {
            const int r = t / NRUN8, q = t - r * NRUN8;          // magnitude row r <-> tile row -1 + r; columns 8q-1 .. 8q+6
            const int iy = ty0 - 1 + r, c0 = tx0 + 8 * q - 1;    // image row, first image column of the run
            uint4 zero = make_uint4(0, 0, 0, 0);
            uint4 *mrow = reinterpret_cast<uint4 *>(&mag[r * MAGW3 + 8 * q]);
            if (iy < 0 || iy >= d.h || c0 > d.w - 1 || c0 + 7 < 0) { *mrow = zero; continue; }
            const int rm = max(iy - 1, 0) - (ty0 - HALO), r0 = iy - (ty0 - HALO), rp = min(iy + 1, d.h - 1) - (ty0 - HALO);
            if (c0 == -1 || c0 == d.w - 1) {      // the two border runs of an image row whose width is a multiple of 8: SIMD with a patch
                s_work[atomicAdd(&s_nwork, 1)] = (u16)t;
                continue;
            }
            if (c0 - 1 < 0 || c0 + 8 > d.w - 1) {
                // any other window that leaves the image on the left or right: scalar, replicated columns, zeros outside
                u32 m[4] = {0, 0, 0, 0};
                for (int j = 0; j < 8; ++j) {
                    const int ix = c0 + j;
                    if (ix < 0 || ix >= d.w) continue;
                    int bx, by;
                    const int v = sobel_scalar(tile, rm, r0, rp, ix, tx0, d.w, bx, by);
                    m[j >> 1] |= (u32)v << (16 * (j & 1));
                }
                *mrow = make_uint4(m[0], m[1], m[2], m[3]);
                continue;
            }
            // window = bytes 24q+10 .. 24q+39 of the three rows = words 6q+2 .. 6q+9 (first two bytes of the first word excluded)
            const u32 *w0 = reinterpret_cast<const u32 *>(tile + rm * ROWB) + 6 * q + 2;
            const u32 *w1 = reinterpret_cast<const u32 *>(tile + r0 * ROWB) + 6 * q + 2;
            const u32 *w2 = reinterpret_cast<const u32 *>(tile + rp * ROWB) + 6 * q + 2;
            // centre pixel (output pixel 3) of the middle row: bytes 24q+22 .. 24q+24 = word 3 bytes 2,3 + word 4 byte 0
            const u32 refw = __byte_perm(w1[3], w1[4], 0x4432);
            // byte i of word k has channel (1 + k + i) % 3
            const u32 P[3] = {__byte_perm(refw, 0, 0x1021), __byte_perm(refw, 0, 0x2102), __byte_perm(refw, 0, 0x0210)};
            u32 acc = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const u32 pk = P[k % 3];
                u32 da = __vabsdiffu4(w0[k], pk), db = __vabsdiffu4(w1[k], pk), dc = __vabsdiffu4(w2[k], pk);
                if (k == 0) { da &= 0xffff0000u; db &= 0xffff0000u; dc &= 0xffff0000u; }
                acc |= ((da & 0x7f7f7f7fu) + K) | da;
                acc |= ((db & 0x7f7f7f7fu) + K) | db;
                acc |= ((dc & 0x7f7f7f7fu) + K) | dc;
            }
            if ((acc & 0x80808080u) == 0) { *mrow = zero; continue; }
            s_work[atomicAdd(&s_nwork, 1)] = (u16)t;
        }
    }
    __syncthreads();
    // ---- phase 1b: 16x2 SIMD Sobel of the listed runs ----
    {
        const int nwork = s_nwork;
        for (int wi = tid; wi < nwork; wi += NT) {
            const int t = s_work[wi];
            const int r = t / NRUN8, q = t - r * NRUN8;
            const int iy = ty0 - 1 + r;
            const int rm = max(iy - 1, 0) - (ty0 - HALO), r0 = iy - (ty0 - HALO), rp = min(iy + 1, d.h - 1) - (ty0 - HALO);
            const uint2 *w0 = reinterpret_cast<const uint2 *>(tile + rm * ROWB) + 3 * q + 1;
            const uint2 *w1 = reinterpret_cast<const uint2 *>(tile + r0 * ROWB) + 3 * q + 1;
            const uint2 *w2 = reinterpret_cast<const uint2 *>(tile + rp * ROWB) + 3 * q + 1;
            // halfword stream h = 0 .. 31 of a row = bytes 24q+8 .. 24q+39; stream pixel p (0 .. 9, output pixel j = p - 1)
            // channel c sits at h = 2 + 3p + c.  V = T + 2M + B, D' = B - T - 1 (B + ~T), registers k = h / 2.
            // BORDER_REPLICATE for the run that starts one column left of the image (stream pixel 1 := pixel 2, i.e. bytes 1..3
            // of word 1 := bytes 0..2 of word 2) and for the run whose first column is the last image column (pixel 2 := pixel 1)
            const int c0 = tx0 + 8 * q - 1;
            const bool lb = c0 == -1, rb = c0 == d.w - 1;
            u32 V[16], D[16];
#pragma unroll
            for (int k2 = 0; k2 < 4; ++k2) {
                uint2 a = w0[k2], m = w1[k2], b = w2[k2];
                if (k2 == 0 && (lb || rb)) {
                    const uint2 a1 = w0[1], m1 = w1[1], b1 = w2[1];
                    if (lb) { a.y = __byte_perm(a.y, a1.x, 0x6540); m.y = __byte_perm(m.y, m1.x, 0x6540); b.y = __byte_perm(b.y, b1.x, 0x6540); }
                }
                if (k2 == 1 && rb) {
                    const uint2 a1 = w0[0], m1 = w1[0], b1 = w2[0];
                    a.x = __byte_perm(a.x, a1.y, 0x3765); m.x = __byte_perm(m.x, m1.y, 0x3765); b.x = __byte_perm(b.x, b1.y, 0x3765);
                }
                const u32 aw[2] = {a.x, a.y}, mw[2] = {m.x, m.y}, bw[2] = {b.x, b.y};
#pragma unroll
                for (int u = 0; u < 2; ++u) {
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        const u32 sel = hf ? 0x4342u : 0x4140u;
                        const u32 tt = __byte_perm(aw[u], 0, sel), mm = __byte_perm(mw[u], 0, sel), bb = __byte_perm(bw[u], 0, sel);
                        const int k = 4 * k2 + 2 * u + hf;
                        V[k] = __vadd2(__vadd2(tt, bb), __vadd2(mm, mm));
                        D[k] = __vadd2(bb, ~tt);
                    }
                }
            }
            // g = h - 3: |dx|[g] = |V[g+6] - V[g]|, dy[g] = D[g] + 2 D[g+3] + D[g+6]; registers k' = 1 .. 12 hold g = 2 .. 25,
            // i.e. output pixel j channel c at g = 2 + 3j + c
            u32 M[13];
#pragma unroll
            for (int k = 1; k <= 12; ++k) {
                const u32 adx = vmaxu2(V[k + 3], V[k]) - vminu2(V[k + 3], V[k]);            // lanes never borrow: max >= min
                const u32 mid = __byte_perm(D[k + 1], D[k + 2], 0x5432);                     // D[g+3]
                const u32 y = __vadd2(__vadd2(__vadd2(D[k], D[k + 3]), __vadd2(mid, mid)), 0x00040004u);   // dy (the -1 biases add up to -4)
                const u32 ady = __vmaxs2(y, __vadd2(~y, 0x00010001u));
                M[k] = adx + ady;                                                            // <= 2040 per lane
            }
            // maximum over the three channels of every pixel: pixels 2i, 2i+1 live in registers 3i+1 .. 3i+3
            u32 o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const u32 a = M[3 * i + 1], b = M[3 * i + 2], c = M[3 * i + 3];
                const u32 ev = vmaxu2(vmaxu2(a, a >> 16), b);            // low lane: max(a.lo, a.hi, b.lo)
                const u32 od = vmaxu2(vmaxu2(c, c >> 16), b >> 16);      // low lane: max(c.lo, c.hi, b.hi)
                o[i] = __byte_perm(ev, od, 0x5410);
            }
            if (lb) o[0] &= 0xffff0000u;                                   // column -1 is outside the image
            if (rb) { o[0] &= 0x0000ffffu; o[1] = o[2] = o[3] = 0; }       // only column w-1 is inside
            *reinterpret_cast<uint4 *>(&mag[r * MAGW3 + 8 * q]) = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
    __syncthreads();

    // ---- phase 2a: candidate list (pixels above canny_lo), HSV colour masks, gray; task = (row, 8 columns) ----
    constexpr int NT2 = TH * (TW / 8) / NT;          // tasks per thread (2)
    u32 mWYR[NT2];
#pragma unroll
    for (int it = 0; it < NT2; ++it) {
        const int t = tid + it * NT;
        const int ty = t >> 4, g = t & 15;
        const int iy = ty0 + ty, ix0 = tx0 + 8 * g;
        u32 mW = 0, mY = 0, mR = 0;
        const int nvalid = iy < d.h ? min(8, d.w - ix0) : 0;       // <= 0: nothing of this task is inside the image
        if (nvalid > 0) {
            const int trow = ty + HALO;
            // -- pixels above the low threshold go to the candidate list (a few % of the image; handled densely in 2b) --
            const u16 *mc = mag + (ty + 1) * MAGW3 + 8 * g;          // mc[0] = column 8g-1, mc[1 .. 8] = this task's columns
            const uint4 mv = *reinterpret_cast<const uint4 *>(mc);
            const u32 m8 = mc[8];
            u32 mx = vmaxu2(vmaxu2(mv.x, mv.y), vmaxu2(mv.z, mv.w));
            mx = max(max(mx & 0xffffu, mx >> 16), m8);
            if ((int)mx > lo) {
                const u32 mm[5] = {mv.x, mv.y, mv.z, mv.w, m8};
                u32 cmask = 0;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int c = (int)((mm[(i + 1) >> 1] >> (16 * ((i + 1) & 1))) & 0xffffu);
                    if (c > lo && i < nvalid) cmask |= 1u << i;
                }
                if (cmask) {
                    int pos = atomicAdd(&s_ncand, __popc(cmask));
                    while (cmask) {
                        const int i = __ffs(cmask) - 1;
                        cmask &= cmask - 1;
                        s_cand[pos++] = (u16)((ty << 7) | (8 * g + i));
                    }
                }
            }
            // -- colours + gray: 8 pixels = 24 bytes = words 6g+4 .. 6g+9 of the tile row --
            const uint2 *pw = reinterpret_cast<const uint2 *>(tile + trow * ROWB) + 3 * g + 2;
            const uint2 q0 = pw[0], q1 = pw[1], q2 = pw[2];
            const u32 Wd[6] = {q0.x, q0.y, q1.x, q1.y, q2.x, q2.y};
            // any byte >= vmin (the smallest V threshold of all colour ranges)?  v = max(b, g, r) >= vmin needs one.
            u32 anyv;
            {
                const u32 Kv = 0x01010101u * (u32)(0x80 - min(vmin, 128));
                u32 a = 0;
#pragma unroll
                for (int k = 0; k < 6; ++k) a |= ((Wd[k] & 0x7f7f7f7fu) + Kv) | Wd[k];
                anyv = vmin <= 128 ? (a & 0x80808080u) : 1u;
            }
            u32 gl = 0, gh = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                // pixel i = bytes 3i .. 3i+2 -> px = [b, g, r, 0]
                const int bo = 3 * i, wk = bo >> 2, sh = bo & 3;
                const u32 px = sh == 0 ? (Wd[wk] & 0x00ffffffu) : __byte_perm(Wd[wk], Wd[wk < 5 ? wk + 1 : 5], sh == 1 ? 0x4321 : sh == 2 ? 0x4432 : 0x4543) & 0x00ffffffu;
                // gray = (b 3735 + g 19235 + r 9798 + 16384) >> 15
                const u32 gy = __dp2a_hi(9798u, px, __dp2a_lo(3735u | (19235u << 16), px, 16384u)) >> 15;
                if (i < 4) gl |= gy << (8 * i); else gh |= gy << (8 * (i - 4));
                if (anyv && i < nvalid) {
                    const int b = px & 0xff, gg = (px >> 8) & 0xff, rr = px >> 16;
                    const int v = max(b, max(gg, rr));
                    u32 in = s_lutV[v];
                    if (in) {
                        const int mn = min(b, min(gg, rr)), diff = v - mn;
                        const int sv = (diff * s_sdiv[v] + 2048) >> 12;
                        in &= s_lutS[sv];
                        if (in) {
                            int hh = (v == rr) ? (gg - b) : (v == gg) ? (b - rr + 2 * diff) : (rr - gg + 4 * diff);
                            hh = (hh * s_hdiv[diff] + 2048) >> 12;
                            if (hh < 0) hh += 180;
                            in &= s_lutH[hh];
                            mW |= (in & 1u) << i;
                            mY |= ((in >> 1) & 1u) << i;
                            mR |= (((in >> 2) | (in >> 3)) & 1u) << i;
                        }
                    }
                }
            }
            if (gray) {
                u8 *gp = gray + ((size_t)f * d.h + iy) * d.w + ix0;
                if (nvalid == 8 && (d.w & 7) == 0) *reinterpret_cast<uint2 *>(gp) = make_uint2(gl, gh);
                else
                    for (int i = 0; i < nvalid; ++i) gp[i] = (u8)((i < 4 ? gl >> (8 * i) : gh >> (8 * (i - 4))) & 0xff);
            }
        }
        mWYR[it] = mW | (mY << 8) | (mR << 16);
    }
    __syncthreads();
    // ---- phase 2b: the candidates, one per thread: signed gradient of the first maximal channel, direction, NMS ----
    {
        const int ncand = s_ncand;
        for (int ci = tid; ci < ncand; ci += NT) {
            const int e = s_cand[ci], ty = e >> 7, tx = e & 127;
            const int iy = ty0 + ty, ix = tx0 + tx;
            const int rm = max(iy - 1, 0) - (ty0 - HALO), rp = min(iy + 1, d.h - 1) - (ty0 - HALO);
            const u16 *m = mag + (ty + 1) * MAGW3 + tx + 1;
            const int c = m[0];
            int xs, ys;
            sobel_scalar(tile, rm, ty + HALO, rp, ix, tx0, d.w, xs, ys);
            const int ax = abs(xs);
            const long long ay = (long long)abs(ys) << 15;
            const long long t22 = (long long)ax * 13573, t67 = t22 + ((long long)ax << 16);
            bool ismax;
            if (ay < t22) ismax = c > m[-1] && c >= m[1];
            else if (ay > t67) ismax = c > m[-MAGW3] && c >= m[MAGW3];
            else {
                const int sg = ((xs ^ ys) < 0) ? -1 : 1;
                ismax = c > m[-MAGW3 - sg] && c > m[MAGW3 + sg];
            }
            if (ismax) {
                atomicOr(&s_cbits[ty * 4 + (tx >> 5)], 1u << (tx & 31));
                if (c > hi) atomicOr(&s_cbits[TH * 4 + ty * 4 + (tx >> 5)], 1u << (tx & 31));
            }
        }
    }
    __syncthreads();
    // ---- phase 2c: plane words.  W / Y / R: 4 lanes x 8 bits by a 4 x 4 byte transpose; cand / strong: from shared memory ----
#pragma unroll
    for (int it = 0; it < NT2; ++it) {
        const int t = tid + it * NT;
        const int ty = t >> 4, g = t & 15;
        const int iy = ty0 + ty;
        u32 A = mWYR[it];
        {
            const u32 o1 = __shfl_xor_sync(0xffffffffu, A, 1);
            A = (lane & 1) ? __byte_perm(A, o1, 0x3715) : __byte_perm(A, o1, 0x6240);
            const u32 o2 = __shfl_xor_sync(0xffffffffu, A, 2);
            A = (lane & 2) ? __byte_perm(A, o2, 0x3276) : __byte_perm(A, o2, 0x5410);
        }
        const int xw = (tx0 >> 5) + (g >> 2);
        if (iy < d.h && xw < d.wp) {
            const int pl = lane & 3;      // lane j of the group holds plane j (W, Y, R); lane 3 writes the candidate plane
            const u32 val = pl < 3 ? A : s_cbits[ty * 4 + (g >> 2)];
            planesA[(((size_t)f * PA_COUNT + pl) * d.h + iy) * d.wp + xw] = val;
            if (pl == 0) planesA[(((size_t)f * PA_COUNT + PA_STRONG) * d.h + iy) * d.wp + xw] = s_cbits[TH * 4 + ty * 4 + (g >> 2)];
        }
    }
}

// ---- experimental: register-marching kernel (LSF_MARCH=1) -------------------------------------------------------------
// Bit-identical alternative to k_color_canny for frames without resize / colour transform and w % 4 == 0.  Measured on
// B200 (1000 x 640x480): 2.78 ms vs 2.62 ms for the tiled kernel -- 27 % fewer instructions (198 vs 272 per pixel) but
// 110-123 registers hold the occupancy at 16 warps/SM, so it is NOT the default.  A thread owns
// 4 adjacent columns and marches down CM_R output rows; nothing goes through shared memory except the HSV tables.
// Per input row it loads 7 aligned words (columns x-2 .. x+5, replicate border synthesised with byte permutes),
// widens the 24 channel bytes to 16-bit lanes (2 per register, interleaved B,G,R as in memory) and runs the Sobel
// arithmetic on the packed lanes with the 16x2 integer SIMD instructions of sm_100 (VIADD.16x2, VIMNMX.S16x2):
//   vertical smoothing   V = row[r-1] + 2 row[r] + row[r+1]        (two packed adds: pair sums of consecutive rows)
//   horizontal smoothing T = lane[j-3] + 2 lane[j] + lane[j+3]     (same channel of the neighbouring pixels = lanes +-3)
//   |dx| = |V[j+3] - V[j-3]|,  |dy| = |T(r+1) - T(r-1)|,  L1 magnitude, maximum over the 3 channels per pixel
// Signed gradients, the winning channel and the NMS direction are only evaluated for the few pixels above the low
// threshold.  Results are bit-identical to k_color_canny.
constexpr int CM_R = 24;                  // output rows per thread (CM_R + 4 steps; the step loop is NOT unrolled: an
constexpr int CM_WARPS = 4;               // unrolled body overflowed the instruction cache, ncu: no_instruction stalls)

// |a - b| + 1024 in both 16-bit lanes, for lanes a, b <= 1023:  D = a + (1023 - b) = a - b + 1023;
// max(D + 1, 2047 - D) = 1024 + |a - b|   (LOP3, VIADD.16x2, LOP3, VIADDMNMX.S16x2)
__device__ __forceinline__ u32 vabsdiff2b(u32 a, u32 b)
{
    const u32 D = __vadd2(a, b ^ 0x03ff03ffu);
    return __viaddmax_s16x2(D, 0x00010001u, D ^ 0x07ff07ffu);
}
__device__ __forceinline__ int lane16(const u32 *r, int j) { return (j & 1) ? (int)(r[j >> 1] >> 16) : (int)(r[j >> 1] & 0xffffu); }

__global__ void __launch_bounds__(CM_WARPS * 32) k_color_canny_march(Dims d, ColorParams cp, const u8 *__restrict__ src,
                                                                    u32 *__restrict__ planesA, u8 *__restrict__ gray)
{
    __shared__ int s_sdiv[256], s_hdiv[256];
    __shared__ u8 s_lutH[256], s_lutS[256], s_lutV[256];   // bit i: value inside colour range i (white, yellow, red1, red2)
    const int lane = threadIdx.x, tid = threadIdx.y * 32 + lane;
    for (int i = tid; i < 256; i += CM_WARPS * 32) {
        s_sdiv[i] = c_sdiv[i];
        s_hdiv[i] = c_hdiv[i];
        int mh = 0, ms = 0, mv = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            mh |= (i >= cp.lo[q][0] && i <= cp.hi[q][0]) << q;
            ms |= (i >= cp.lo[q][1] && i <= cp.hi[q][1]) << q;
            mv |= (i >= cp.lo[q][2] && i <= cp.hi[q][2]) << q;
        }
        s_lutH[i] = (u8)mh; s_lutS[i] = (u8)ms; s_lutV[i] = (u8)mv;
    }
    __syncthreads();
    const int w = d.w, h = d.h, f = blockIdx.z;
    const int x = (blockIdx.x * 32 + lane) * 4;
    const int y0 = (blockIdx.y * CM_WARPS + threadIdx.y) * CM_R;
    if (y0 >= h) return;                                   // warp-uniform
    const bool active = x < w;
    const bool left = x == 0, right = x + 4 >= w;
    const u8 *fsrc = src + (size_t)f * d.src_frame;
    const int lo = cp.canny_lo, hi = cp.canny_hi;

    u32 E1[12], P0[12], T0[9], T1[9];      // previous row (16-bit lanes), pair sum of the two rows before, T of the two rows before
    int M0[6], M1[6];                      // magnitudes of the two previous gradient rows, columns x-1 .. x+4
    u32 dirq = 0;                          // direction codes of the middle magnitude row: 2 bits per own pixel
    u32 maskq0 = 0, maskq1 = 0, grayq0 = 0, grayq1 = 0;   // colour-mask nibbles / gray of the rows loaded two / one steps ago
#pragma unroll
    for (int k = 0; k < 12; ++k) { E1[k] = 0; P0[k] = 0; }
#pragma unroll
    for (int k = 0; k < 9; ++k) { T0[k] = 0; T1[k] = 0; }
#pragma unroll
    for (int k = 0; k < 6; ++k) { M0[k] = 0; M1[k] = 0; }

    {
        // the 7 words of a row; words 0,1 / 5,6 lie outside the frame for the first / last thread of a row
        auto load_row = [&](int ry, u32 *Wn) {
#pragma unroll
            for (int q = 0; q < 7; ++q) Wn[q] = 0;
            if (active) {
                const int cy = min(max(ry, 0), h - 1);
                const u32 *rp = reinterpret_cast<const u32 *>(fsrc + (size_t)(cy + d.top) * d.src_pitch) + (3 * x) / 4 - 2;
                Wn[2] = __ldg(rp + 2); Wn[3] = __ldg(rp + 3); Wn[4] = __ldg(rp + 4);
                if (!left) { Wn[0] = __ldg(rp); Wn[1] = __ldg(rp + 1); }
                if (!right) { Wn[5] = __ldg(rp + 5); Wn[6] = __ldg(rp + 6); }
            }
        };
        u32 Wn[7];
        load_row(y0 - 2, Wn);
#pragma unroll 1
        for (int t = 0; t < CM_R + 4; ++t) {
            const int ry = y0 - 2 + t;                     // input row of this step
            // ---- this row's words (loaded one step ahead), replicate border, next row's loads in flight ----
            u32 W[7];
#pragma unroll
            for (int q = 0; q < 7; ++q) W[q] = Wn[q];
            if (left) { W[0] = __byte_perm(W[2], 0, 0x1000); W[1] = __byte_perm(W[2], 0, 0x2102); }
            if (right) { W[5] = __byte_perm(W[4], 0, 0x1321); W[6] = __byte_perm(W[4], 0, 0x0032); }
            load_row(ry + 1, Wn);
            u32 E2[12];
#pragma unroll
            for (int m = 0; m < 6; ++m) { E2[2 * m] = __byte_perm(W[m], 0, 0x4342); E2[2 * m + 1] = __byte_perm(W[m + 1], 0, 0x4140); }
            // ---- colour masks + gray of this row (output two steps later) ----
            u32 msk = 0, gpk = 0;
            if (active && t >= 2 && t < CM_R + 2 && ry < h) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int b = lane16(E2, 6 + 3 * i), g = lane16(E2, 7 + 3 * i), r = lane16(E2, 8 + 3 * i);
                    int v = max(b, max(g, r));
                    u32 in = s_lutV[v];
                    if (in) {
                        int mn = min(b, min(g, r)), diff = v - mn;
                        int sv = (diff * s_sdiv[v] + 2048) >> 12;
                        in &= s_lutS[sv];
                        if (in) {
                            int hh = (v == r) ? (g - b) : (v == g) ? (b - r + 2 * diff) : (r - g + 4 * diff);
                            hh = (hh * s_hdiv[diff] + 2048) >> 12;
                            if (hh < 0) hh += 180;
                            in &= s_lutH[hh];
                            msk |= ((in & 1u) | ((in & 2u) << 3) | ((((in >> 2) | (in >> 3)) & 1u) << 8)) << i;   // W: bits 0-3, Y: 4-7, R: 8-11
                        }
                    }
                    gpk |= (u32)((b * 3735 + g * 19235 + r * 9798 + 16384) >> 15) << (8 * i);
                }
            }
            // ---- horizontal smoothing of this row: T[k] = lanes (2k+1, 2k+2), k = 1..9 ----
            u32 T2[9];
#pragma unroll
            for (int k = 1; k <= 9; ++k) {
                const u32 mid = __funnelshift_r(E2[k], E2[k + 1], 16);            // lanes (2k+1, 2k+2)
                T2[k - 1] = __vadd2(__vadd2(E2[k - 1], E2[k + 2]), __vadd2(mid, mid));   // lanes (2k-2,2k-1) + (2k+4,2k+5) + 2 mid
            }
            // ---- gradient row r = ry - 1: |dx| from V, |dy| from T of the rows around it ----
            u32 P1[12];
#pragma unroll
            for (int k = 0; k < 12; ++k) P1[k] = __vadd2(E1[k], E2[k]);
            int Mn[6] = {0, 0, 0, 0, 0, 0};
            u32 dirn = 0;
            const int gr = ry - 1;
            if (t >= 2 && gr >= 0 && gr < h) {              // warp-uniform
                u32 V[12], Mg[9];
#pragma unroll
                for (int k = 0; k < 12; ++k) V[k] = __vadd2(P0[k], P1[k]);
#pragma unroll
                for (int k = 1; k <= 9; ++k) Mg[k - 1] = __vadd2(vabsdiff2b(V[k + 2], V[k - 1]), vabsdiff2b(T2[k - 1], T0[k - 1]));   // + 2048 per lane
                // Mg[k-1] lanes (2k+1, 2k+2): column c (0..5 <-> x-1..x+4), channel ch sits at stream lane 3(c+1)+ch
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                    const int j = 3 * (c + 1) - 3;          // index into the lane stream of Mg (lane 3 -> 0)
                    const int m0 = lane16(Mg, j), m1 = lane16(Mg, j + 1), m2 = lane16(Mg, j + 2);
                    int mx = max(m0, max(m1, m2)) - 2048;
                    if ((c == 0 && left) || (c == 5 && right) || !active) mx = 0;   // outside the frame
                    Mn[c] = mx;
                    if (c >= 1 && c <= 4 && mx > lo) {
                        // rare: signed gradient of the first maximal channel -> NMS direction code
                        const int ch = m0 - 2048 == mx ? 0 : (m1 - 2048 == mx ? 1 : 2);
                        const int js = 3 * (c + 1) + ch;   // stream lane of that channel
                        int xs, ys;
                        // lanes js+-3 of V, lane js of T2 / T0 (T arrays start at stream lane 3)
                        if (ch == 0) { xs = lane16(V, 3 * (c + 1) + 3) - lane16(V, 3 * (c + 1) - 3); ys = lane16(T2, 3 * (c + 1) - 3) - lane16(T0, 3 * (c + 1) - 3); }
                        else if (ch == 1) { xs = lane16(V, 3 * (c + 1) + 4) - lane16(V, 3 * (c + 1) - 2); ys = lane16(T2, 3 * (c + 1) - 2) - lane16(T0, 3 * (c + 1) - 2); }
                        else { xs = lane16(V, 3 * (c + 1) + 5) - lane16(V, 3 * (c + 1) - 1); ys = lane16(T2, 3 * (c + 1) - 1) - lane16(T0, 3 * (c + 1) - 1); }
                        (void)js;
                        const int ax = abs(xs);
                        const long long ay = (long long)abs(ys) << 15;
                        const long long t22 = (long long)ax * 13573, t67 = t22 + ((long long)ax << 16);
                        u32 code = ay < t22 ? 0u : (ay > t67 ? 1u : (((xs ^ ys) < 0) ? 3u : 2u));
                        dirn |= code << (2 * (c - 1));
                    }
                }
            }
            // ---- NMS + thresholds of output row y = ry - 2 (needs magnitude rows y-1, y, y+1) ----
            const int y = ry - 2;
            if (t >= 4 && y < h) {                          // warp-uniform
                const int *mu = M0, *mm = M1, *md = Mn;   // rows y-1, y, y+1
                u32 cand = 0, strong = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int c = mm[i + 1];
                    if (c > lo) {
                        const u32 code = (dirq >> (2 * i)) & 3u;
                        bool ismax;
                        if (code == 0) ismax = c > mm[i] && c >= mm[i + 2];
                        else if (code == 1) ismax = c > mu[i + 1] && c >= md[i + 1];
                        else if (code == 2) ismax = c > mu[i] && c > md[i + 2];         // s = +1: (y-1, x-1), (y+1, x+1)
                        else ismax = c > mu[i + 2] && c > md[i];                        // s = -1: (y-1, x+1), (y+1, x-1)
                        if (ismax) { cand |= 1u << i; if (c > hi) strong |= 1u << i; }
                    }
                }
                // ---- output: 5 plane nibbles -> words (8 lanes x 4 bits), gray ----
                const u32 mq = maskq0;                      // masks of the row loaded two steps ago (= row y)
                u32 nib[PA_COUNT] = {mq & 15u, (mq >> 4) & 15u, (mq >> 8) & 15u, cand, strong};
#pragma unroll
                for (int pl = 0; pl < PA_COUNT; ++pl) {
                    u32 v = nib[pl] << (4 * (lane & 7));
                    if (__any_sync(0xffffffffu, v != 0)) {
                        v |= __shfl_xor_sync(0xffffffffu, v, 1);
                        v |= __shfl_xor_sync(0xffffffffu, v, 2);
                        v |= __shfl_xor_sync(0xffffffffu, v, 4);
                    }
                    nib[pl] = v;
                }
                if ((lane & 7) < PA_COUNT) {
                    const int pl = lane & 7, xw = blockIdx.x * 4 + (lane >> 3);
                    u32 val = pl == 0 ? nib[0] : pl == 1 ? nib[1] : pl == 2 ? nib[2] : pl == 3 ? nib[3] : nib[4];
                    if (xw < d.wp) planesA[(((size_t)f * PA_COUNT + pl) * h + y) * d.wp + xw] = val;
                }
                if (gray && active) *reinterpret_cast<u32 *>(gray + ((size_t)f * h + y) * w + x) = grayq0;
            }
            // ---- rotate the windows ----
#pragma unroll
            for (int k = 0; k < 12; ++k) { P0[k] = P1[k]; E1[k] = E2[k]; }
#pragma unroll
            for (int k = 0; k < 9; ++k) { T0[k] = T1[k]; T1[k] = T2[k]; }
#pragma unroll
            for (int k = 0; k < 6; ++k) { M0[k] = M1[k]; M1[k] = Mn[k]; }
            dirq = dirn;
            maskq0 = maskq1; maskq1 = msk; grayq0 = grayq1; grayq1 = gpk;
        }
    }
}

void launch_color_canny(const Dims &d, const ColorParams &cp, const u8 *src, const TmaDesc &tma, const u8 *tables, u32 *planesA,
                        u8 *gray, cudaStream_t st)
{
    ensure_tables();
    const bool fast = d.identity_geom && d.identity_color && (d.w & 3) == 0 && d.w >= 8 && d.h >= 4 &&
                      (d.src_pitch & 3) == 0 && (d.src_frame & 3) == 0 && (((uintptr_t)src) & 3) == 0 && getenv("LSF_MARCH");
    if (fast) {
        dim3 grid((d.w + 127) / 128, (d.h + CM_WARPS * CM_R - 1) / (CM_WARPS * CM_R), d.n), block(32, CM_WARPS);
        k_color_canny_march<<<grid, block, 0, st>>>(d, cp, src, planesA, gray);
        ++g_launches;
        return;
    }
    dim3 grid((d.w + TW - 1) / TW, (d.h + TH - 1) / TH, d.n);
    static const bool use_v2 = getenv("LSF_CC_V2") != nullptr;      // the round-1 kernel, for A/B runs
    if (use_v2) {
        if (tma.valid)
            k_color_canny<true><<<grid, NT, 0, st>>>(tma.map, d, cp, src, tables, planesA, gray);
        else
            k_color_canny<false><<<grid, NT, 0, st>>>(tma.map, d, cp, src, tables, planesA, gray);
    } else {
        int vmin = 256;          // smallest V lower bound over the four colour ranges: below it no pixel has a colour
        for (int i = 0; i < 4; ++i) vmin = std::min(vmin, std::max(0, cp.lo[i][2]));
        if (tma.valid)
            k_color_canny3<true><<<grid, NT, 0, st>>>(tma.map, d, cp, src, tables, planesA, gray, vmin);
        else
            k_color_canny3<false><<<grid, NT, 0, st>>>(tma.map, d, cp, src, tables, planesA, gray, vmin);
    }
    ++g_launches;
}

// ---- parity taps ------------------------------------------------------------------------------------
__global__ void k_unpack_plane(const u32 *__restrict__ plane, int h, int w, int wp, u8 *__restrict__ dst)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= h * w) return;
    int y = i / w, x = i - y * w;
    dst[i] = ((plane[(size_t)y * wp + (x >> 5)] >> (x & 31)) & 1) ? 255 : 0;
}

void launch_unpack_plane(const u32 *plane, int h, int w, int wp, u8 *dst, cudaStream_t st)
{
    k_unpack_plane<<<(h * w + 255) / 256, 256, 0, st>>>(plane, h, w, wp, dst);
    ++g_launches;
}

__global__ void k_labels_tap(const u32 *__restrict__ pa, int h, int w, int wp, u8 *__restrict__ dst)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= h * w) return;
    int y = i / w, x = i - y * w;
    size_t o = (size_t)y * wp + (x >> 5), ps = (size_t)h * wp;
    int sh = x & 31;
    u32 v = ((pa[o] >> sh) & 1) | (((pa[o + ps] >> sh) & 1) << 1) | (((pa[o + 2 * ps] >> sh) & 1) << 2);
    u32 cand = (pa[o + 3 * ps] >> sh) & 1, strong = (pa[o + 4 * ps] >> sh) & 1;
    v |= (strong ? 2u : cand ? 1u : 0u) << 4;
    dst[i] = (u8)v;
}

void launch_labels_tap(const u32 *planesA_frame, int h, int w, int wp, u8 *dst, cudaStream_t st)
{
    k_labels_tap<<<(h * w + 255) / 256, 256, 0, st>>>(planesA_frame, h, w, wp, dst);
    ++g_launches;
}

__global__ void k_image_tap(Dims d, ColorParams cp, const u8 *__restrict__ fsrc, u8 *__restrict__ dst)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.h * d.w * 3) return;
    int c = i % 3, p = i / 3, y = p / d.w, x = p - y * d.w;
    int sy = nearest_src(y + d.top, d.src_h, d.dh), sx = nearest_src(x, d.src_w, d.dw);
    u8 v = fsrc[(size_t)sy * d.src_pitch + (size_t)sx * 3 + c];
    dst[i] = d.identity_color ? v : color_correct(v, cp.ai_scale[c], cp.ai_shift[c]);
}

void launch_image_tap(const Dims &d, const ColorParams &cp, const u8 *src_frame, u8 *dst, cudaStream_t st)
{
    k_image_tap<<<(d.h * d.w * 3 + 255) / 256, 256, 0, st>>>(d, cp, src_frame, dst);
    ++g_launches;
}

}  // namespace lsf
