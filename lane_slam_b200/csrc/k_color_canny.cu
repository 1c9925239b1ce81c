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
// Tile 64x32 pixels + 2-pixel halo, staged into shared memory by one TMA 3-D box load
// ([frame][row][byte], zero fill outside the frame) when there is no resize; a gather loader otherwise.
#include "common.cuh"

namespace lsf {

constexpr int TW = 64, TH = 32, HALO = 2;
constexpr int XOFF = 16;                 // bytes of left padding: TMA needs a 16-byte aligned inner start (measured)
constexpr int BOX_X = 224;               // bytes per tile row: 16 + (64+2)*3 = 214, padded to a multiple of 16
constexpr int BOX_Y = TH + 2 * HALO;     // 36
constexpr int MAGW = TW + 4;             // 68: (TW+2) columns + pad
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

template <bool USE_TMA>
__global__ void __launch_bounds__(NT) k_color_canny(const __grid_constant__ CUtensorMap tmap, Dims d, ColorParams cp,
                                                    const u8 *__restrict__ src, u32 *__restrict__ planesA,
                                                    u8 *__restrict__ gray)
{
    __shared__ __align__(128) u8 tile[BOX_Y * BOX_X];
    __shared__ u16 mag[(TH + 2) * MAGW];
    __shared__ short2 dxy[TH * TW];
    __shared__ int s_sdiv[256], s_hdiv[256];
    __shared__ __align__(8) u64 bar;

    const int tid = threadIdx.x;
    const int tx0 = blockIdx.x * TW, ty0 = blockIdx.y * TH, f = blockIdx.z;
    s_sdiv[tid] = c_sdiv[tid];
    s_hdiv[tid] = c_hdiv[tid];

    if (USE_TMA) {
        const u32 bar_a = smem_u32(&bar);
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_a), "r"(1));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(BOX_X * BOX_Y)
                         : "memory");
            int c0 = tx0 * 3 - XOFF, c1 = ty0 - HALO + d.top, c2 = f;
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                ::"r"(smem_u32(tile)), "l"(&tmap), "r"(c0), "r"(c1), "r"(c2), "r"(bar_a)
                : "memory");
        }
        // all threads wait for the transaction bytes (phase parity 0)
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
        for (int i = tid; i < BOX_Y * BOX_X; i += NT) {
            int r = i / BOX_X, k = i - r * BOX_X;
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
        for (int i = tid; i < BOX_Y * BOX_X; i += NT) {
            int c = ((i % BOX_X) + 2) % 3;  // tile row starts at byte 3*tx0 - 16: channel = (k - 16) mod 3
            tile[i] = color_correct(tile[i], cp.ai_scale[c], cp.ai_shift[c]);
        }
        __syncthreads();
    }

    // ---- phase 1: per-channel Sobel, L1 magnitude, first maximal channel; (TH+2) x (TW+2) positions ----
    for (int p = tid; p < (TH + 2) * (TW + 2); p += NT) {
        int ty = p / (TW + 2) - 1, tx = p - (ty + 1) * (TW + 2) - 1;
        int iy = ty0 + ty, ix = tx0 + tx;
        int best = 0, bdx = 0, bdy = 0;
        if (iy >= 0 && iy < d.h && ix >= 0 && ix < d.w) {
            // BORDER_REPLICATE: clamp neighbour coordinates to the image, then address the tile
            int rm = (max(iy - 1, 0) - (ty0 - HALO)) * BOX_X, r0 = (iy - (ty0 - HALO)) * BOX_X,
                rp = (min(iy + 1, d.h - 1) - (ty0 - HALO)) * BOX_X;
            int cm = (max(ix - 1, 0) - tx0) * 3 + XOFF, c0 = (ix - tx0) * 3 + XOFF, cq = (min(ix + 1, d.w - 1) - tx0) * 3 + XOFF;
            best = -1;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                int a00 = tile[rm + cm + c], a01 = tile[rm + c0 + c], a02 = tile[rm + cq + c];
                int a10 = tile[r0 + cm + c], a12 = tile[r0 + cq + c];
                int a20 = tile[rp + cm + c], a21 = tile[rp + c0 + c], a22 = tile[rp + cq + c];
                int dx = (a02 + 2 * a12 + a22) - (a00 + 2 * a10 + a20);
                int dy = (a20 + 2 * a21 + a22) - (a00 + 2 * a01 + a02);
                int nrm = abs(dx) + abs(dy);
                if (nrm > best) { best = nrm; bdx = dx; bdy = dy; }
            }
        }
        mag[(ty + 1) * MAGW + tx + 1] = (u16)best;
        if (ty >= 0 && ty < TH && tx >= 0 && tx < TW) dxy[ty * TW + tx] = make_short2((short)bdx, (short)bdy);
    }
    __syncthreads();

    // ---- phase 2: NMS + thresholds, HSV masks, gray; one warp = 32 consecutive pixels of one row ----
    const int warp = tid >> 5, lane = tid & 31;
    const int half = warp & 1;
    const int tx = half * 32 + lane, ix = tx0 + tx;
    const int xw = ix >> 5;
    for (int k = 0; k < TH / 4; ++k) {
        int ty = (warp >> 1) + 4 * k, iy = ty0 + ty;
        if (iy >= d.h) break;  // warp-uniform
        bool inimg = ix < d.w;
        int cand = 0, strong = 0, mw = 0, my = 0, mr = 0;
        if (inimg) {
            const u16 *m = mag + (ty + 1) * MAGW + tx + 1;
            int c = m[0];
            if (c > cp.canny_lo) {
                short2 g = dxy[ty * TW + tx];
                int xs = g.x, ys = g.y;
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
                cand = ismax;
                strong = ismax && c > cp.canny_hi;
            }
            const u8 *px = tile + (ty + HALO) * BOX_X + tx * 3 + XOFF;
            int b = px[0], g = px[1], r = px[2];
            int v = max(b, max(g, r)), mn = min(b, min(g, r)), diff = v - mn;
            int s = (diff * s_sdiv[v] + 2048) >> 12;
            int hh = (v == r) ? (g - b) : (v == g) ? (b - r + 2 * diff) : (r - g + 4 * diff);
            hh = (hh * s_hdiv[diff] + 2048) >> 12;
            if (hh < 0) hh += 180;
#define INR(i) (hh >= cp.lo[i][0] && hh <= cp.hi[i][0] && s >= cp.lo[i][1] && s <= cp.hi[i][1] && v >= cp.lo[i][2] && v <= cp.hi[i][2])
            mw = INR(0);
            my = INR(1);
            mr = INR(2) || INR(3);
#undef INR
            if (gray) gray[((size_t)f * d.h + iy) * d.w + ix] = (u8)((b * 3735 + g * 19235 + r * 9798 + 16384) >> 15);
        }
        u32 w0 = __ballot_sync(0xffffffffu, mw), w1 = __ballot_sync(0xffffffffu, my), w2 = __ballot_sync(0xffffffffu, mr);
        u32 w3 = __ballot_sync(0xffffffffu, cand), w4 = __ballot_sync(0xffffffffu, strong);
        if (lane < PA_COUNT && xw < d.wp) {
            u32 val = lane == 0 ? w0 : lane == 1 ? w1 : lane == 2 ? w2 : lane == 3 ? w3 : w4;
            planesA[(((size_t)f * PA_COUNT + lane) * d.h + iy) * d.wp + xw] = val;
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
