// k_hysteresis.cu -- K4+K5: Canny hysteresis (8-connected growth of strong pixels through
// candidates) and, on the result, 3x3 cross dilation of the three colour masks AND edges.
//
// Replaces: the hysteresis stage inside cv2.Canny (line_detector_lsd.py:60-62) and
//           cv2.dilate(bw, ellipse 3x3 = cross) + cv2.bitwise_and(bw, edges) (line_detector_lsd.py:52-56).
//
// Everything is bit-parallel on packed planes (32 pixels per word).  One CTA owns one frame; the
// candidate/strong planes live in shared memory in row strips.  A frame whose planes fit one strip
// (any frame up to ~640x600) converges entirely in shared memory; larger frames sweep their strips
// down and up until no strip changes.  The fixed point is unique, so the result is bit-exact.
#include "common.cuh"

namespace lsf {

constexpr int HT = 512;                      // threads per CTA
constexpr size_t HYST_SMEM = 160 * 1024;     // bytes of plane storage per CTA

// OR of a row word with its left/right neighbours (bit i-1, i, i+1), carrying across words
__device__ __forceinline__ u32 h3(const u32 *row, int xw, int wp)
{
    u32 m = row[xw];
    u32 l = xw > 0 ? row[xw - 1] : 0u, r = xw + 1 < wp ? row[xw + 1] : 0u;
    return m | (m << 1) | (l >> 31) | (m >> 1) | (r << 31);
}

// grow seeds through runs of consecutive candidate bits inside one word (both directions)
__device__ __forceinline__ u32 fill_word(u32 seeds, u32 cand)
{
    u32 x = seeds & cand, p = cand;
    x |= p & (x << 1); p &= p << 1;
    x |= p & (x << 2); p &= p << 2;
    x |= p & (x << 4); p &= p << 4;
    x |= p & (x << 8); p &= p << 8;
    x |= p & (x << 16);
    p = cand;
    x |= p & (x >> 1); p &= p >> 1;
    x |= p & (x >> 2); p &= p >> 2;
    x |= p & (x >> 4); p &= p >> 4;
    x |= p & (x >> 8); p &= p >> 8;
    x |= p & (x >> 16);
    return x;
}

// Converge rows [1, rows] of the strip held in smem (row 0 and rows+1 are halo rows, read only).
// Returns (block-uniform) whether any word changed.
__device__ int converge_strip(const u32 *cand, volatile u32 *strong, int rows, int wp)
{
    int any = 0;
    const int nwords = rows * wp;
    while (true) {
        int changed = 0;
        for (int i = threadIdx.x; i < nwords; i += HT) {
            int r = i / wp + 1, xw = i - (r - 1) * wp;
            u32 c = cand[r * wp + xw];
            if (!c) continue;
            u32 s = strong[r * wp + xw];
            if ((s & c) == c) continue;
            const u32 *up = (const u32 *)strong + (r - 1) * wp, *md = (const u32 *)strong + r * wp,
                      *dn = (const u32 *)strong + (r + 1) * wp;
            u32 nb = h3(up, xw, wp) | h3(md, xw, wp) | h3(dn, xw, wp);
            u32 ns = fill_word(s | (c & nb), c) | s;
            if (ns != s) { strong[r * wp + xw] = ns; changed = 1; }
        }
        if (!__syncthreads_or(changed)) break;
        any = 1;
    }
    return any;
}

__global__ void __launch_bounds__(HT) k_hysteresis(Dims d, int dilate, int strip_rows, const u32 *__restrict__ planesA,
                                                  u32 *__restrict__ planesB)
{
    extern __shared__ u32 sm[];
    const int f = blockIdx.x, wp = d.wp, h = d.h;
    const size_t ps = (size_t)h * wp;
    const u32 *pa = planesA + (size_t)f * PA_COUNT * ps;
    u32 *pb = planesB + (size_t)f * PB_COUNT * ps;
    const u32 *g_cand = pa + PA_CAND * ps, *g_strong = pa + PA_STRONG * ps;
    u32 *g_edge = pb + PB_EDGE * ps;
    u32 *cand = sm, *strong = sm + (size_t)(strip_rows + 2) * wp;
    const int nstrips = (h + strip_rows - 1) / strip_rows;

    if (nstrips == 1) {
        for (int i = threadIdx.x; i < (h + 2) * wp; i += HT) {
            int r = i / wp;
            bool in = r >= 1 && r <= h;
            cand[i] = in ? g_cand[i - wp] : 0u;
            strong[i] = in ? g_strong[i - wp] : 0u;
        }
        __syncthreads();
        converge_strip(cand, strong, h, wp);
        for (int i = threadIdx.x; i < h * wp; i += HT) g_edge[i] = strong[i + wp];
    } else {
        for (int i = threadIdx.x; i < h * wp; i += HT) g_edge[i] = g_strong[i];
        __syncthreads();
        int dir = 1;
        while (true) {
            int changed_any = 0;
            for (int k = 0; k < nstrips; ++k) {
                int s = dir > 0 ? k : nstrips - 1 - k;
                int r0 = s * strip_rows, r1 = min(h, r0 + strip_rows), rows = r1 - r0;
                for (int i = threadIdx.x; i < (rows + 2) * wp; i += HT) {
                    int r = r0 - 1 + i / wp;
                    bool in = r >= 0 && r < h;
                    size_t go = (size_t)r * wp + (i % wp);
                    cand[i] = in ? g_cand[go] : 0u;
                    strong[i] = in ? g_edge[go] : 0u;
                }
                __syncthreads();
                if (converge_strip(cand, strong, rows, wp)) {
                    changed_any = 1;
                    for (int i = threadIdx.x; i < rows * wp; i += HT) g_edge[(size_t)r0 * wp + i] = strong[i + wp];
                }
                __syncthreads();
            }
            if (!changed_any) break;
            dir = -dir;
        }
    }
    __syncthreads();

    // ---- dilate (3x3 cross) the raw colour masks, AND with the final edges ----
    const u32 lastmask = (d.w & 31) ? ((1u << (d.w & 31)) - 1u) : 0xffffffffu;
    for (int i = threadIdx.x; i < 3 * h * wp; i += HT) {
        int c = i / (h * wp), j = i - c * (h * wp);
        int y = j / wp, xw = j - y * wp;
        const u32 *raw = pa + (size_t)(PA_RAW_W + c) * ps;
        u32 m = raw[j], bw = m;
        if (dilate >= 3) {
            u32 l = xw > 0 ? raw[j - 1] : 0u, r = xw + 1 < wp ? raw[j + 1] : 0u;
            bw |= (m << 1) | (l >> 31) | (m >> 1) | (r << 31);
            if (y > 0) bw |= raw[j - wp];
            if (y + 1 < h) bw |= raw[j + wp];
            if (xw == wp - 1) bw &= lastmask;
        }
        pb[(size_t)(PB_BW0 + c) * ps + j] = bw;
        pb[(size_t)(PB_EC0 + c) * ps + j] = bw & g_edge[j];
    }
}

void launch_hysteresis(const Dims &d, int dilate, const u32 *planesA, u32 *planesB, cudaStream_t st)
{
    static PerDevice once;
    once.ensure(1, [] { cudaFuncSetAttribute(k_hysteresis, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HYST_SMEM); });
    // rows per strip so that two planes (+2 halo rows each) fit; a whole frame if possible
    size_t row_bytes = (size_t)d.wp * 4 * 2;
    int max_rows = (int)(HYST_SMEM / row_bytes) - 2;
    int strip_rows = d.h <= max_rows ? d.h : max_rows;
    size_t smem = (size_t)(strip_rows + 2) * row_bytes;
    k_hysteresis<<<d.n, HT, smem, st>>>(d, dilate, strip_rows, planesA, planesB);
    ++g_launches;
}

}  // namespace lsf
