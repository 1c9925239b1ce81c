// k_jpeg.cu -- baseline JPEG decode on the GPU, bit-identical to cv2.imdecode (libjpeg-turbo defaults): the step before the
// line path (duckietown_utils/jpg.py:21-31, line_detector_node.py:155).  A frame crosses PCIe as ~75 KB of JPEG instead of
// 921.6 KB of BGR.  Arithmetic in jpeg_core.cuh (shared with the host-side simulation the CPU tests run).
//   k_jpeg_huff   one CTA per image: stuffing removal, self-synchronising parallel Huffman decode (subsequence per thread),
//                 prefix sum of the coefficient slots, coefficient write, DC prediction
//   k_jpeg_idct   one thread per 8x8 block: dequantise + jpeg_idct_islow -> component planes
//   k_jpeg_color  fancy upsampling + YCbCr -> BGR, 4 pixels per thread, into the frame buffer the front end reads
// and lsf_front_end_batch_jpeg (the C ABI entry): header parsing on the host (a few hundred bytes per file), one H2D copy of
// the concatenated files, the three kernels, then the ordinary batch pipeline on the decoded frames (device resident).
#include <vector>

#include "ctx.cuh"
#include "jpeg_core.cuh"

namespace lsf {

#ifndef LSF_JT
#define LSF_JT 512
#endif
constexpr int JT = LSF_JT;               // threads (= subsequences) per image

struct JpegItem {
    u32 ent_off, ent_len;                // entropy-coded bytes inside the blob
    u32 tabs, pad;                       // index of the Huffman table set
};

struct JpegState {
    u8 *blob; size_t blob_cap;
    u32 *clean; size_t clean_words;      // per image: stuffing-free big-endian words
    int16_t *coef; size_t coef_per_img;  // per image: nblocks * 64
    int16_t *dcdiff;                     // per image: nblocks DC differences (decode order)
    u8 *plane; size_t plane_per_img;     // per image: Y, Cb, Cr planes back to back (padded to whole MCUs)
    JpegItem *items, *h_items;
    u16 *qtabs, *h_qtabs;                // [n][3][64]
    jd::Tabs *tabsets, *h_tabsets; int tabs_cap;
    int *status, *h_status;
    int n_cap, W, H;
};

__device__ __forceinline__ int block_excl_scan(int v, int *sm, int &total)
{
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int x = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += x;
    }
    __syncthreads();
    if (lane == 31) sm[warp] = incl;
    __syncthreads();
    int off = 0, tot = 0;
    for (int k = 0; k < JT / 32; ++k) { const int w = sm[k]; if (k < warp) off += w; tot += w; }
    total = tot;
    return off + incl - v;
}

__global__ void __launch_bounds__(JT) k_jpeg_huff(jd::Image g, const u8 *__restrict__ blob, const JpegItem *__restrict__ items,
                                                 const jd::Tabs *__restrict__ tabsets, u32 *__restrict__ clean_all, size_t clean_words,
                                                 int16_t *__restrict__ coef_all, size_t coef_per_img, int16_t *__restrict__ dcdiff_all,
                                                 int *__restrict__ status)
{
    __shared__ jd::Tabs tabs;
    __shared__ jd::Span E[2][JT];
    __shared__ int s_scan[JT / 32];
    __shared__ u8 s_chg[JT];
    const int t = threadIdx.x, img = blockIdx.x;
    const JpegItem it = items[img];
    {
        const u32 *src = reinterpret_cast<const u32 *>(tabsets + it.tabs);
        u32 *dst = reinterpret_cast<u32 *>(&tabs);
        for (int i = t; i < (int)(sizeof(jd::Tabs) / 4); i += JT) dst[i] = src[i];
    }
    // ---- phase 0: remove the stuffed zero bytes (a 0x00 that follows a 0xFF), write big-endian words ----
    const u8 *raw = blob + it.ent_off;
    const u32 len = it.ent_len;
    u32 *words = clean_all + (size_t)img * clean_words;
    u8 *cbytes = reinterpret_cast<u8 *>(words);
    const u32 C = ((len + JT - 1) / JT + 3) & ~3u;
    const u32 lo = min(len, (u32)t * C), hi = min(len, lo + C);
    int removed = 0;
    for (u32 j = lo; j < hi; ++j) removed += (raw[j] == 0 && j > 0 && raw[j - 1] == 0xFF) ? 1 : 0;
    int total_removed;
    const int before = block_excl_scan(removed, s_scan, total_removed);
    {
        u32 k = lo - (u32)before;
        for (u32 j = lo; j < hi; ++j) {
            const u8 b = raw[j];
            if (b == 0 && j > 0 && raw[j - 1] == 0xFF) continue;
            cbytes[k ^ 3u] = b;
            ++k;
        }
    }
    const u32 nb = len - (u32)total_removed;           // stuffing-free bytes
    if (t < 12) { const u32 k = nb + t; if ((size_t)k < clean_words * 4) cbytes[k ^ 3u] = 0; }   // zero tail: decode_span reads two words ahead
    __syncthreads();
    // ---- phase A: every thread decodes its subsequence from a guessed state (own first bit, DC of slot 0) ----
    const u32 S = max(16u, ((nb + JT - 1) / JT + 3) & ~3u);
    const int nsub = (int)((nb + S - 1) / S);
    const u32 total_bits = nb * 8;
    const u32 limit = min(total_bits, (u32)(t + 1) * S * 8);
    jd::Span mine; mine.pos = (u32)t * S * 8; mine.s = 0; mine.adv = 0;
    if (t < nsub) jd::decode_span(words, tabs, g.slot_dc, g.slot_ac, g.bpm, mine, limit, (int16_t *)nullptr, 0u, 0u);
    E[0][t] = mine;
    // ---- phase B: re-decode from the left neighbour's end state until nothing changes ----
    int cur = 0, rounds = 0;
    bool dirty = t >= 1 && t < nsub;
    while (true) {
        __syncthreads();
        int changed = 0;
        if (dirty) {
            jd::Span st; st.pos = E[cur][t - 1].pos; st.s = E[cur][t - 1].s; st.adv = 0;
            jd::decode_span(words, tabs, g.slot_dc, g.slot_ac, g.bpm, st, limit, (int16_t *)nullptr, 0u, 0u);
            changed = (st.pos != mine.pos || st.s != mine.s) ? 1 : 0;
            mine = st;
        }
        E[cur ^ 1][t] = mine;
        s_chg[t] = (u8)changed;
        const int any = __syncthreads_or(changed);
        dirty = t >= 1 && t < nsub && s_chg[t - 1];
        cur ^= 1;
        if (!any) break;
        if (++rounds > JT + 2) { if (t == 0) status[img] = 1; break; }
    }
    __syncthreads();
    // ---- phase C: first coefficient slot of every subsequence ----
    int tot_adv;
    const u32 ustart = (u32)block_excl_scan(t < nsub ? (int)mine.adv : 0, s_scan, tot_adv);
    // ---- phase D: decode once more from the true state, writing the coefficients (DC entries hold differences) ----
    const u32 nblocks = (u32)g.mcux * g.mcuy * g.bpm;
    int16_t *coef = coef_all + (size_t)img * coef_per_img;
    if (t < nsub) {
        jd::Span st;
        if (t == 0) { st.pos = 0; st.s = 0; } else { st.pos = E[cur][t - 1].pos; st.s = E[cur][t - 1].s; }
        st.adv = 0;
        jd::decode_span(words, tabs, g.slot_dc, g.slot_ac, g.bpm, st, limit, coef, ustart, nblocks, dcdiff_all + (size_t)img * (coef_per_img / 64));
    }
    __syncthreads();
    // ---- DC prediction: running sum of the differences per component, in decode order ----
    {
        const int nm = g.mcux * g.mcuy, per = (nm + JT - 1) / JT;
        const int m0 = min(nm, t * per), m1 = min(nm, m0 + per);
        const int16_t *dd = dcdiff_all + (size_t)img * (coef_per_img / 64);       // DC differences, one per block, decode order
        int sum[3] = {0, 0, 0};
        for (int m = m0; m < m1; ++m)
            for (int s = 0; s < g.bpm; ++s) sum[g.slot_comp[s]] += dd[m * g.bpm + s];
        int base[3], dummy;
        for (int c = 0; c < 3; ++c) base[c] = block_excl_scan(sum[c], s_scan, dummy);
        for (int m = m0; m < m1; ++m)
            for (int s = 0; s < g.bpm; ++s) {
                const int c = g.slot_comp[s];
                base[c] += dd[m * g.bpm + s];
                coef[((size_t)m * g.bpm + s) * 64] = (int16_t)base[c];
            }
    }
}

// plane layout of one image: component c at plane_off[c], row stride bw[c] * 8
__device__ __forceinline__ size_t plane_off(const jd::Image &g, int c)
{
    size_t o = 0;
    for (int k = 0; k < c; ++k) o += (size_t)g.bw[k] * g.bh[k] * 64;
    return o;
}

// 8 threads per block: thread j runs the column pass on column j, the 8 x 8 intermediate goes through shared memory (rows padded
// to 9 words), then the row pass on row j and one 8-byte store.  256 threads = 32 blocks per CTA.
__global__ void __launch_bounds__(256) k_jpeg_idct(jd::Image g, int n, const int16_t *__restrict__ coef_all, size_t coef_per_img,
                                                  const u16 *__restrict__ qtabs, u8 *__restrict__ plane_all, size_t plane_per_img)
{
    __shared__ int ws[32][8][9];
    const int nblocks = g.mcux * g.mcuy * g.bpm;
    const int j = threadIdx.x & 7, lb = threadIdx.x >> 3;
    const int img = blockIdx.y;
    int tt = blockIdx.x * 32 + lb;
    const bool ok = tt < nblocks;
    int c = 0, by = 0, bx = 0;
    const int16_t *cf = coef_all;
    if (ok) {
        // enumerate blocks plane by plane (neighbouring groups write neighbouring blocks of a plane row)
        while (c < g.ncomp - 1 && tt >= g.bw[c] * g.bh[c]) { tt -= g.bw[c] * g.bh[c]; ++c; }
        by = tt / g.bw[c]; bx = tt - by * g.bw[c];
        int slot0 = 0;
        for (int k = 0; k < c; ++k) slot0 += g.hs[k] * g.vs[k];
        const int mcu = (by / g.vs[c]) * g.mcux + (bx / g.hs[c]);
        const int slot = slot0 + (by % g.vs[c]) * g.hs[c] + (bx % g.hs[c]);
        cf = coef_all + (size_t)img * coef_per_img + ((size_t)mcu * g.bpm + slot) * 64;
        const u16 *q = qtabs + ((size_t)img * 3 + c) * 64;
        int o[8];
        jd::idct_1d(cf[j] * __ldg(q + j), cf[8 + j] * __ldg(q + 8 + j), cf[16 + j] * __ldg(q + 16 + j), cf[24 + j] * __ldg(q + 24 + j),
                    cf[32 + j] * __ldg(q + 32 + j), cf[40 + j] * __ldg(q + 40 + j), cf[48 + j] * __ldg(q + 48 + j), cf[56 + j] * __ldg(q + 56 + j), o);
#pragma unroll
        for (int r = 0; r < 8; ++r) ws[lb][r][j] = jd::descale(o[r], 11);
    }
    __syncwarp();
    if (ok) {
        const int *w = ws[lb][j];
        int o[8];
        jd::idct_1d(w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7], o);
        u32 lo = 0, hi = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            lo |= (u32)jd::range_limit(jd::descale(o[k], 18)) << (8 * k);
            hi |= (u32)jd::range_limit(jd::descale(o[4 + k], 18)) << (8 * k);
        }
        const int stride = g.bw[c] * 8;
        u8 *dst = plane_all + (size_t)img * plane_per_img + plane_off(g, c) + ((size_t)by * 8 + j) * stride + bx * 8;
        *reinterpret_cast<uint2 *>(dst) = make_uint2(lo, hi);
    }
}

__global__ void __launch_bounds__(256) k_jpeg_color(jd::Image g, int n, const u8 *__restrict__ plane_all, size_t plane_per_img,
                                                   u8 *__restrict__ bgr, size_t frame_bytes)
{
    const int W = g.W, H = g.H, W4 = (W + 3) / 4;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)n * H * W4) return;
    const int img = (int)(gid / ((long long)H * W4));
    const int rem = (int)(gid - (long long)img * H * W4);
    const int y = rem / W4, x0 = (rem - y * W4) * 4;
    const u8 *pl = plane_all + (size_t)img * plane_per_img;
    const u8 *py = pl, *pcb = pl + plane_off(g, 1), *pcr = pl + plane_off(g, 2);
    __align__(16) u8 out[12];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int x = min(x0 + i, W - 1);
        const int yy = py[(size_t)y * g.bw[0] * 8 + x];
        if (g.ncomp == 1) { out[3 * i] = out[3 * i + 1] = out[3 * i + 2] = (u8)yy; continue; }
        const int hs1 = g.hmax / g.hs[1], vs1 = g.vmax / g.vs[1];
        const int dw = (W * g.hs[1] + g.hmax - 1) / g.hmax, dh = (H * g.vs[1] + g.vmax - 1) / g.vmax;
        const int cb = jd::chroma_at(pcb, g.bw[1] * 8, dw, dh, hs1, vs1, x, y);
        const int cr = jd::chroma_at(pcr, g.bw[2] * 8, dw, dh, hs1, vs1, x, y);
        jd::ycc_to_bgr(yy, cb, cr, out + 3 * i);
    }
    u8 *dst = bgr + (size_t)img * frame_bytes + ((size_t)y * W + x0) * 3;
    if (x0 + 3 < W && (W & 3) == 0) {
        u32 *d32 = reinterpret_cast<u32 *>(dst);
        d32[0] = reinterpret_cast<const u32 *>(out)[0]; d32[1] = reinterpret_cast<const u32 *>(out)[1]; d32[2] = reinterpret_cast<const u32 *>(out)[2];
    } else {
        for (int i = 0; i < 4 && x0 + i < W; ++i) { dst[3 * i] = out[3 * i]; dst[3 * i + 1] = out[3 * i + 1]; dst[3 * i + 2] = out[3 * i + 2]; }
    }
}

// 4:2:0, 8 pixels per thread (W % 8 == 0): the six vertical sums 3 * near + far of each chroma plane are formed once and shared
// by the eight pixels; aligned word loads for luma and chroma, three 8-byte stores.
__global__ void __launch_bounds__(256) k_jpeg_color_420(jd::Image g, int n, const u8 *__restrict__ plane_all, size_t plane_per_img,
                                                       u8 *__restrict__ bgr, size_t frame_bytes)
{
    const int W = g.W, H = g.H, W8 = W / 8;
    const int img = blockIdx.y;
    const int rem = blockIdx.x * blockDim.x + threadIdx.x;
    if (rem >= H * W8) return;
    const int y = rem / W8, x0 = (rem - y * W8) * 8;
    const u8 *pl = plane_all + (size_t)img * plane_per_img;
    const int ys = g.bw[0] * 8, cs = g.bw[1] * 8;
    const int dw = (W + 1) / 2, dh = (H + 1) / 2;
    const int sy = y >> 1, sx = x0 >> 1;
    int ny = (y & 1) ? sy + 1 : sy - 1;
    ny = ny < 0 ? 0 : (ny > dh - 1 ? dh - 1 : ny);
    const uint2 yv = *reinterpret_cast<const uint2 *>(pl + (size_t)y * ys + x0);
    int col[2][6];       // vertical sums of chroma columns sx-1 .. sx+4 (edge columns are only used inside the image)
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const u8 *p = pl + plane_off(g, 1 + c);
        const u8 *r0 = p + (size_t)sy * cs + sx, *r1 = p + (size_t)ny * cs + sx;
        const u32 a = *reinterpret_cast<const u32 *>(r0), b = *reinterpret_cast<const u32 *>(r1);
#pragma unroll
        for (int k = 0; k < 4; ++k) col[c][1 + k] = (int)((a >> (8 * k)) & 0xff) * 3 + (int)((b >> (8 * k)) & 0xff);
        col[c][0] = sx > 0 ? r0[-1] * 3 + r1[-1] : 0;
        col[c][5] = sx + 4 < dw ? r0[4] * 3 + r1[4] : 0;
    }
    __align__(8) u8 out[24];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int yy = (int)(((i < 4 ? yv.x : yv.y) >> (8 * (i & 3))) & 0xff);
        const int k = 1 + (i >> 1);
        int cc[2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int cur = col[c][k];
            if (i & 1) cc[c] = (sx + (i >> 1) + 1 > dw - 1) ? (cur * 4 + 7) >> 4 : (cur * 3 + col[c][k + 1] + 7) >> 4;
            else cc[c] = (sx + (i >> 1) == 0) ? (cur * 4 + 8) >> 4 : (cur * 3 + col[c][k - 1] + 8) >> 4;
        }
        jd::ycc_to_bgr(yy, cc[0], cc[1], out + 3 * i);
    }
    uint2 *dst = reinterpret_cast<uint2 *>(bgr + (size_t)img * frame_bytes + ((size_t)y * W + x0) * 3);
    dst[0] = reinterpret_cast<const uint2 *>(out)[0]; dst[1] = reinterpret_cast<const uint2 *>(out)[1]; dst[2] = reinterpret_cast<const uint2 *>(out)[2];
}

}  // namespace lsf

using namespace lsf;

void jpeg_destroy(lsf_ctx *ctx)
{
    JpegState *j = (JpegState *)ctx->jpeg;
    if (!j) return;
    for (void *p : {(void *)j->blob, (void *)j->clean, (void *)j->coef, (void *)j->dcdiff, (void *)j->plane, (void *)j->items, (void *)j->qtabs, (void *)j->tabsets,
                    (void *)j->status})
        if (p) cudaFree(p);
    for (void *p : {(void *)j->h_items, (void *)j->h_qtabs, (void *)j->h_tabsets, (void *)j->h_status}) if (p) cudaFreeHost(p);
    delete j;
    ctx->jpeg = nullptr;
}

static int jpeg_reserve(lsf_ctx *ctx, const jd::Image &g, int n, size_t blob_bytes)
{
    JpegState *j = (JpegState *)ctx->jpeg;
    if (!j) {
        j = new JpegState();
        memset(j, 0, sizeof(*j));
        ctx->jpeg = j;
    }
    const size_t nblocks = (size_t)g.mcux * g.mcuy * g.bpm;
    size_t plane = 0;
    for (int c = 0; c < g.ncomp; ++c) plane += (size_t)g.bw[c] * g.bh[c] * 64;
    if (n > j->n_cap || g.W != j->W || g.H != j->H || nblocks * 64 > j->coef_per_img || plane > j->plane_per_img) {
        for (void *p : {(void *)j->coef, (void *)j->dcdiff, (void *)j->plane, (void *)j->items, (void *)j->qtabs, (void *)j->status}) if (p) cudaFree(p);
        for (void *p : {(void *)j->h_items, (void *)j->h_qtabs, (void *)j->h_status}) if (p) cudaFreeHost(p);
        j->coef = nullptr; j->dcdiff = nullptr; j->plane = nullptr; j->items = nullptr; j->qtabs = nullptr; j->status = nullptr;
        j->h_items = nullptr; j->h_qtabs = nullptr; j->h_status = nullptr;
        const int cap = std::max(n, ctx->max_batch);
        j->coef_per_img = nblocks * 64; j->plane_per_img = (plane + 15) & ~(size_t)15;
        CK(cudaMalloc((void **)&j->coef, (size_t)cap * j->coef_per_img * sizeof(int16_t)));
        CK(cudaMalloc((void **)&j->dcdiff, (size_t)cap * (j->coef_per_img / 64) * sizeof(int16_t)));
        CK(cudaMalloc((void **)&j->plane, (size_t)cap * j->plane_per_img));
        CK(cudaMalloc((void **)&j->items, (size_t)cap * sizeof(JpegItem)));
        CK(cudaMalloc((void **)&j->qtabs, (size_t)cap * 3 * 64 * sizeof(u16)));
        CK(cudaMalloc((void **)&j->status, (size_t)cap * sizeof(int)));
        CK(cudaMallocHost((void **)&j->h_items, (size_t)cap * sizeof(JpegItem)));
        CK(cudaMallocHost((void **)&j->h_qtabs, (size_t)cap * 3 * 64 * sizeof(u16)));
        CK(cudaMallocHost((void **)&j->h_status, (size_t)cap * sizeof(int)));
        j->n_cap = cap; j->W = g.W; j->H = g.H;
    }
    if (blob_bytes + 64 > j->blob_cap) {
        if (j->blob) cudaFree(j->blob);
        j->blob = nullptr; j->blob_cap = 0;
        const size_t cap = std::max(blob_bytes + 64, (size_t)ctx->max_batch * g.W * g.H / 4 + 4096);
        CK(cudaMalloc((void **)&j->blob, cap));
        j->blob_cap = cap;
    }
    return LSF_OK;
}

extern "C" int lsf_front_end_batch_jpeg(lsf_ctx *ctx, const uint8_t *blob, const int64_t *offsets, int n, int stages, int k, lsf_segments *out)
{
    if (!ctx) return LSF_E_ARG;
    if (!blob || !offsets || !out || n <= 0) return fail(ctx, LSF_E_ARG, "lsf_front_end_batch_jpeg: null argument or n <= 0");
    if (n > ctx->max_batch) return fail(ctx, LSF_E_CAPACITY, "lsf_front_end_batch_jpeg: n exceeds max_batch " + std::to_string(ctx->max_batch));
    ENTER(ctx);
    // ---- headers (host): geometry of the first file is the batch geometry; Huffman table sets are deduplicated ----
    jd::Image g, gi;
    std::vector<jd::Tabs> sets;
    std::vector<uint64_t> set_hash;
    std::vector<JpegItem> items(n);
    std::vector<jd::Image> imgs;      // only the quantisation tables differ per image
    imgs.reserve(n);
    size_t max_ent = 0;
    for (int i = 0; i < n; ++i) {
        const uint8_t *d = blob + offsets[i];
        const size_t len = (size_t)(offsets[i + 1] - offsets[i]);
        uint64_t h = 0;
        int rc = jd::parse(d, len, gi, nullptr, &h);
        if (rc == 0 && gi.restart) rc = -2;
        if (rc) return fail(ctx, LSF_E_ARG, "lsf_front_end_batch_jpeg: frame " + std::to_string(i) + (rc == -2 ?
                            " is not a baseline Huffman JPEG this decoder covers (8-bit, 1 or 3 components, sampling <= 2, no restart "
                            "intervals): decode it on the host and pass BGR to lsf_front_end_batch" : " is malformed"));
        if (i == 0) g = gi;
        if (gi.W != g.W || gi.H != g.H || gi.ncomp != g.ncomp || gi.bpm != g.bpm || memcmp(gi.hs, g.hs, sizeof(g.hs)) || memcmp(gi.vs, g.vs, sizeof(g.vs)) ||
            memcmp(gi.slot_dc, g.slot_dc, sizeof(g.slot_dc)) || memcmp(gi.slot_ac, g.slot_ac, sizeof(g.slot_ac)))
            return fail(ctx, LSF_E_ARG, "lsf_front_end_batch_jpeg: frame " + std::to_string(i) + " differs from frame 0 in size / sampling / table selection");
        int si = -1;
        for (size_t q = 0; q < set_hash.size(); ++q) if (set_hash[q] == h) si = (int)q;
        if (si < 0) {
            jd::Tabs t;
            jd::parse(d, len, gi, &t, nullptr);
            sets.push_back(t); set_hash.push_back(h);
            si = (int)sets.size() - 1;
        }
        items[i].ent_off = (u32)(offsets[i] - offsets[0]) + gi.ent_off; items[i].ent_len = gi.ent_len; items[i].tabs = (u32)si; items[i].pad = 0;
        max_ent = std::max(max_ent, (size_t)gi.ent_len);
        imgs.push_back(gi);
    }
    if (g.H > ctx->max_src_h || g.W > ctx->max_src_w) return fail(ctx, LSF_E_CAPACITY, "lsf_front_end_batch_jpeg: frame larger than max_src_h x max_src_w");
    const size_t blob_bytes = (size_t)(offsets[n] - offsets[0]);
    if (blob_bytes >= ((size_t)1 << 32)) return fail(ctx, LSF_E_CAPACITY, "lsf_front_end_batch_jpeg: more than 4 GB of JPEG data in one batch");
    int rc = jpeg_reserve(ctx, g, n, blob_bytes);
    if (rc) return rc;
    JpegState *j = (JpegState *)ctx->jpeg;
    const size_t cw = (max_ent + 3) / 4 + 4;
    if (cw * (size_t)n > j->clean_words) {
        if (j->clean) cudaFree(j->clean);
        j->clean = nullptr; j->clean_words = 0;
        const size_t want = std::max(cw * (size_t)n, (size_t)ctx->max_batch * ((size_t)g.W * g.H / 16 + 16));
        CK(cudaMalloc((void **)&j->clean, want * 4));
        j->clean_words = want;
    }
    if ((int)sets.size() > j->tabs_cap) {
        if (j->tabsets) cudaFree(j->tabsets);
        if (j->h_tabsets) cudaFreeHost(j->h_tabsets);
        j->tabsets = nullptr; j->h_tabsets = nullptr;
        const int cap = std::max<int>((int)sets.size(), 8);
        CK(cudaMalloc((void **)&j->tabsets, (size_t)cap * sizeof(jd::Tabs)));
        CK(cudaMallocHost((void **)&j->h_tabsets, (size_t)cap * sizeof(jd::Tabs)));
        j->tabs_cap = cap;
    }
    memcpy(j->h_items, items.data(), (size_t)n * sizeof(JpegItem));
    memcpy(j->h_tabsets, sets.data(), sets.size() * sizeof(jd::Tabs));
    for (int i = 0; i < n; ++i) memcpy(j->h_qtabs + (size_t)i * 192, imgs[i].q, 192 * sizeof(u16));
    ctx->n_events = 0;
    mark(ctx, "start");
    cudaStream_t st = ctx->st;
    // the frame buffer the decoded frames go to: the ctx's first staging buffer (nothing may be staged in it)
    if (ctx->staged[0].valid || ctx->staged[1].valid) { CK(cudaStreamSynchronize(ctx->copy_st)); ctx->staged[0].valid = ctx->staged[1].valid = false; }
    u8 *frames = ctx->stage_buf[0];
    const size_t frame_bytes = (size_t)g.W * g.H * 3;
    CK(cudaMemcpyAsync(j->blob, blob + offsets[0], blob_bytes, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(j->items, j->h_items, (size_t)n * sizeof(JpegItem), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(j->tabsets, j->h_tabsets, sets.size() * sizeof(jd::Tabs), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(j->qtabs, j->h_qtabs, (size_t)n * 192 * sizeof(u16), cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(j->coef, 0, (size_t)n * j->coef_per_img * sizeof(int16_t), st));
    CK(cudaMemsetAsync(j->status, 0, (size_t)n * sizeof(int), st));
    mark(ctx, "jpeg_h2d");
    const size_t per_img_words = j->clean_words / (size_t)std::max(n, 1);
    k_jpeg_huff<<<n, JT, 0, st>>>(g, j->blob, j->items, j->tabsets, j->clean, per_img_words, j->coef, j->coef_per_img, j->dcdiff, j->status);
    if (getenv("LSF_JPEG_SPLIT_TIMING")) mark(ctx, "jpeg_huffman");
    const long long nblk = (long long)n * g.mcux * g.mcuy * g.bpm;
    k_jpeg_idct<<<dim3((unsigned)((nblk / n + 31) / 32), n), 256, 0, st>>>(g, n, j->coef, j->coef_per_img, j->qtabs, j->plane, j->plane_per_img);
    if (getenv("LSF_JPEG_SPLIT_TIMING")) mark(ctx, "jpeg_idct");
    if (g.ncomp == 3 && g.hs[0] == 2 && g.vs[0] == 2 && (g.W & 7) == 0 && (frame_bytes & 7) == 0) {
        const int npx8 = g.H * (g.W / 8);
        k_jpeg_color_420<<<dim3((unsigned)((npx8 + 255) / 256), n), 256, 0, st>>>(g, n, j->plane, j->plane_per_img, frames, frame_bytes);
    } else {
        const long long npx4 = (long long)n * g.H * ((g.W + 3) / 4);
        k_jpeg_color<<<(unsigned)((npx4 + 255) / 256), 256, 0, st>>>(g, n, j->plane, j->plane_per_img, frames, frame_bytes);
    }
    g_launches += 3;
    mark(ctx, "jpeg_decode");
    ctx->events_keep = true;
    CK(cudaMemcpyAsync(j->h_status, j->status, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaGetLastError());
    ctx->jpeg_last_bytes = (long long)blob_bytes;
    rc = lsf_front_end_batch(ctx, frames, n, g.H, g.W, (size_t)g.W * 3, LSF_MEM_DEVICE, stages, k, out);
    for (int i = 0; i < n; ++i)
        if (j->h_status[i]) return fail(ctx, LSF_E_INTERNAL, "lsf_front_end_batch_jpeg: the Huffman decode of frame " + std::to_string(i) + " did not converge");
    return rc;
}
