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
// Tile 128x16 pixels + halo, staged into shared memory by one TMA 3-D box load of uint32 elements
// ([frame][row][word], 104 words x 20 rows, zero fill outside the frame) when there is no resize; a gather
// loader otherwise.  Each thread handles runs of 4 pixels with 32-bit shared-memory loads.
#include "common.cuh"

namespace lsf {

constexpr int TW = 128, TH = 16, HALO = 2;
constexpr int XOFF = 16;                 // bytes of left padding: TMA needs a 16-byte aligned inner start (measured)
constexpr int ROWB = 416;                // bytes per tile row = 104 uint32: columns -5 .. 132 of the tile
constexpr int ROWW = ROWB / 4;
constexpr int BOX_Y = TH + 2 * HALO;     // 20
constexpr int NRUN = TW / 4 + 2;         // 34 runs of 4 columns covering tile columns -4 .. 131
constexpr int MAGW = 4 * NRUN;           // 136 u16 per magnitude row; column c lives at c + 4
constexpr int NT = 256;

__constant__ int c_sdiv[256];
__constant__ int c_hdiv[256];
static bool g_tabs_ready = false;

static void ensure_tables()
{
    if (g_tabs_ready) return;
    int sdiv[256], hdiv[256];
    sdiv[0] = hdiv[0] = 0;
    for (int i = 1; i < 256; ++i) {
        sdiv[i] = (int)lrint((255 << 12) / (1.0 * i));
        hdiv[i] = (int)lrint((180 << 12) / (6.0 * i));
    }
    cudaMemcpyToSymbol(c_sdiv, sdiv, sizeof(sdiv));
    cudaMemcpyToSymbol(c_hdiv, hdiv, sizeof(hdiv));
    g_tabs_ready = true;
}

__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }

// byte k (0..19) of a 5-word window
#define WB(W, k) (int)(((W)[(k) >> 2] >> (8 * ((k) & 3))) & 0xffu)

template <bool USE_TMA>
__global__ void __launch_bounds__(NT) k_color_canny(const __grid_constant__ CUtensorMap tmap, Dims d, ColorParams cp,
                                                    const u8 *__restrict__ src, u32 *__restrict__ planesA,
                                                    u8 *__restrict__ gray)
{
    __shared__ __align__(128) u8 tile[BOX_Y * ROWB];
    __shared__ __align__(16) u16 mag[(TH + 2) * MAGW];
    __shared__ __align__(16) u32 dxy[TH * TW];        // (dx & 0xffff) | (dy << 16) of the selected channel
    __shared__ int s_sdiv[256], s_hdiv[256];
    __shared__ u8 s_lutH[256], s_lutS[256], s_lutV[256];   // bit i: value inside colour range i (white, yellow, red1, red2)
    __shared__ __align__(8) u64 bar;

    const int tid = threadIdx.x;
    const int tx0 = blockIdx.x * TW, ty0 = blockIdx.y * TH, f = blockIdx.z;
    {
        s_sdiv[tid] = c_sdiv[tid];
        s_hdiv[tid] = c_hdiv[tid];
        int mh = 0, ms = 0, mv = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            mh |= (tid >= cp.lo[i][0] && tid <= cp.hi[i][0]) << i;
            ms |= (tid >= cp.lo[i][1] && tid <= cp.hi[i][1]) << i;
            mv |= (tid >= cp.lo[i][2] && tid <= cp.hi[i][2]) << i;
        }
        s_lutH[tid] = (u8)mh; s_lutS[tid] = (u8)ms; s_lutV[tid] = (u8)mv;
    }

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
            if (ix0 >= 1 && ix0 + 4 <= d.w - 1) {
                // fast path: columns ix0-1 .. ix0+4 all inside the image -> 5 words per row
                const u32 *w0 = reinterpret_cast<const u32 *>(tile + rm * ROWB) + 3 * q;
                const u32 *w1 = reinterpret_cast<const u32 *>(tile + r0 * ROWB) + 3 * q;
                const u32 *w2 = reinterpret_cast<const u32 *>(tile + rp * ROWB) + 3 * q;
                u32 T[5], M[5], B[5];
#pragma unroll
                for (int i = 0; i < 5; ++i) { T[i] = w0[i]; M[i] = w1[i]; B[i] = w2[i]; }
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

void launch_color_canny(const Dims &d, const ColorParams &cp, const u8 *src, const TmaDesc &tma, u32 *planesA, u8 *gray,
                        cudaStream_t st)
{
    ensure_tables();
    dim3 grid((d.w + TW - 1) / TW, (d.h + TH - 1) / TH, d.n);
    if (tma.valid)
        k_color_canny<true><<<grid, NT, 0, st>>>(tma.map, d, cp, src, planesA, gray);
    else
        k_color_canny<false><<<grid, NT, 0, st>>>(tma.map, d, cp, src, planesA, gray);
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
