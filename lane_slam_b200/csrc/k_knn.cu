// k_knn.cu -- K13: exact brute-force Hamming kNN over packed 256-bit descriptors (carry-save adders + __popc).
//
// Replaces BinaryDescriptorMatcher::knnMatch (src/line_descriptor/src/binary_descriptor_matcher.cpp:258-335,
// Mihasher::query :635-753) and the distance kernel match() (src/line_descriptor/src/bitops_custom.hpp:83-96).
// Contract: ascending distance; neighbours farther than max_dist are not reported (-1).  Ties inside one distance come
// in the reference's own order (default) or by ascending train index (tie_order = 1, what cv2.BFMatcher does).
// The reference's order is the discovery order of Mihasher(256, 32) (:634-753): one-byte substrings, radius s = 0, 1, ..
// per substring, substrings k = 0 .. 31 in turn, flip patterns in increasing integer order, bucket members in train order
// -> rows of equal distance are ordered by ( s* = min_k popc(q[k]^t[k]), k* = first byte reaching s*, q[k*]^t[k*], index ).
// The hash tables themselves are a CPU-side accelerator of the same exact search and are not reproduced; the key is
// evaluated only for the few pairs that can still enter a query's top-k list.  Pinned against the compiled reference
// (tests/golden/lbd_reference.npz).
//
// Each thread keeps one query (8 x 32-bit words) in registers; map descriptors are staged through shared
// memory and broadcast to the warp.  The map is split across gridDim.y so the grid fills the 148 SMs;
// a second kernel merges the per-split candidates.
#include "common.cuh"

namespace lsf {

constexpr int KT = 256;        // threads = queries per CTA
constexpr int KTILE = 256;     // map descriptors per shared-memory tile (8 KB)

// Hamming distance of two 256-bit codes.  Eight POPC per pair saturate the XU pipe (ncu: 92 % of peak, POPC issues at a
// quarter of the integer ALU rate), so the eight XOR words first go through a carry-save adder tree (LOP3 xor3 / majority
// on the ALU pipe) that leaves four words of weight 1, 2, 4, 4: four POPC instead of eight.
__device__ __forceinline__ u32 xor3(u32 a, u32 b, u32 c) { return a ^ b ^ c; }
__device__ __forceinline__ u32 maj3(u32 a, u32 b, u32 c) { return (a & b) | (c & (a | b)); }
__device__ __forceinline__ int hamming256(const uint4 &qa, const uint4 &qb, const uint4 &a, const uint4 &b)
{
    const u32 x0 = qa.x ^ a.x, x1 = qa.y ^ a.y, x2 = qa.z ^ a.z, x3 = qa.w ^ a.w;
    const u32 x4 = qb.x ^ b.x, x5 = qb.y ^ b.y, x6 = qb.z ^ b.z, x7 = qb.w ^ b.w;
    const u32 s1 = xor3(x0, x1, x2), c1 = maj3(x0, x1, x2);
    const u32 s2 = xor3(x3, x4, x5), c2 = maj3(x3, x4, x5);
    const u32 s3 = xor3(s1, s2, x6), c3 = maj3(s1, s2, x6);
    const u32 ones = s3 ^ x7, c4 = s3 & x7;
    const u32 t = xor3(c1, c2, c3), f1 = maj3(c1, c2, c3);
    const u32 twos = t ^ c4, f2 = t & c4;
    return __popc(ones) + 2 * __popc(twos) + 4 * (__popc(f1) + __popc(f2));
}

// sort key of a candidate: (distance << 17) | discovery key (s* 4 bits | k* 5 bits | xor byte 8 bits); 0 in index order.
// The discovery key of (query of lane X, train row) is evaluated by the whole warp: all lanes hold the same train row at
// the same time, lane L takes byte L of the XOR (the query words of lane X are broadcast with eight shuffles), and one
// warp-wide minimum of (popcount << 13 | byte index << 8 | byte) is the key.  ~25 warp instructions instead of ~200 per
// candidate evaluated by a single lane while the other 31 wait.
constexpr int KEY_SHIFT = 17;
__device__ __forceinline__ u32 sel8(const uint4 &a, const uint4 &b, int w)
{
    const u32 lo = (w & 2) ? ((w & 1) ? a.w : a.z) : ((w & 1) ? a.y : a.x);
    const u32 hi = (w & 2) ? ((w & 1) ? b.w : b.z) : ((w & 1) ? b.y : b.x);
    return (w & 4) ? hi : lo;
}
// enter: this lane's candidate may enter its list.  Returns this lane's key (valid where enter is set).  Warp-uniform call.
__device__ __forceinline__ int mih_keys_warp(bool enter, const uint4 &qa, const uint4 &qb, const uint4 &a, const uint4 &b)
{
    const int lane = threadIdx.x & 31;
    unsigned em = __ballot_sync(0xffffffffu, enter);
    int mykey = 0;
    if (!em) return 0;
    const u32 rw = sel8(a, b, lane >> 2);             // the train row's word that holds byte `lane`
    while (em) {
        const int X = __ffs(em) - 1;
        em &= em - 1;
        uint4 xa, xb;
        xa.x = __shfl_sync(0xffffffffu, qa.x, X); xa.y = __shfl_sync(0xffffffffu, qa.y, X);
        xa.z = __shfl_sync(0xffffffffu, qa.z, X); xa.w = __shfl_sync(0xffffffffu, qa.w, X);
        xb.x = __shfl_sync(0xffffffffu, qb.x, X); xb.y = __shfl_sync(0xffffffffu, qb.y, X);
        xb.z = __shfl_sync(0xffffffffu, qb.z, X); xb.w = __shfl_sync(0xffffffffu, qb.w, X);
        const u32 byte = ((sel8(xa, xb, lane >> 2) ^ rw) >> (8 * (lane & 3))) & 0xffu;
        const u32 packed = ((u32)__popc(byte) << 13) | ((u32)lane << 8) | byte;
        const u32 kx = __reduce_min_sync(0xffffffffu, packed);
        if (lane == X) mykey = (int)kx;
    }
    return mykey;
}

template <int K>
__device__ __forceinline__ void knn_insert(int (&bd)[K], int (&bi)[K], int dd, int idx)
{
    // caller guarantees dd < bd[K-1] (strict: an equal key keeps the earlier, smaller index)
#pragma unroll
    for (int j = K - 1; j > 0; --j) {
        if (bd[j - 1] > dd) { bd[j] = bd[j - 1]; bi[j] = bi[j - 1]; }
        else if (bd[j] > dd) { bd[j] = dd; bi[j] = idx; }
    }
    if (bd[0] > dd) { bd[0] = dd; bi[0] = idx; }
}

constexpr int QPT = 1;         // queries per thread (2 was measured slower: 0.51 vs 0.42 ms at C4, although the one-query
                               // kernel is MIO-throttled on the tile broadcasts)

template <int K, bool REFORDER>
__global__ void __launch_bounds__(KT) k_knn_partial(const uint4 *__restrict__ q, int nq_cap, const int *__restrict__ nq_dev,
                                                   const uint4 *__restrict__ m, int nm, int max_dist, int chunk,
                                                   int *__restrict__ pidx, int *__restrict__ pdist)
{
    __shared__ uint4 tile[KTILE * 2];
    const int nq = nq_dev ? min(*nq_dev, nq_cap) : nq_cap;
    if (blockIdx.x * KT * QPT >= nq) return;
    const int split = blockIdx.y, nsplit = gridDim.y;
    const int m0 = split * chunk, m1 = min(nm, m0 + chunk);
    int qi[QPT];
    uint4 qa[QPT], qb[QPT];
    int bd[QPT][K], bi[QPT][K];
#pragma unroll
    for (int u = 0; u < QPT; ++u) {
        qi[u] = (blockIdx.x * QPT + u) * KT + threadIdx.x;
        qa[u] = make_uint4(0, 0, 0, 0); qb[u] = qa[u];
        if (qi[u] < nq) { qa[u] = q[2 * (size_t)qi[u]]; qb[u] = q[2 * (size_t)qi[u] + 1]; }
#pragma unroll
        for (int j = 0; j < K; ++j) { bd[u][j] = 0x7fffffff; bi[u][j] = -1; }
    }
    int wd[QPT];                       // distance a candidate must not exceed to enter the list (kept in a register)
#pragma unroll
    for (int u = 0; u < QPT; ++u) wd[u] = max_dist;
    for (int t0 = m0; t0 < m1; t0 += KTILE) {
        const int nt = min(KTILE, m1 - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < nt * 2; i += KT) tile[i] = m[2 * (size_t)t0 + i];
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < nt; ++j) {
            const uint4 a = tile[2 * j], b = tile[2 * j + 1];
#pragma unroll
            for (int u = 0; u < QPT; ++u) {
                const int dd = hamming256(qa[u], qb[u], a, b);
                const bool enter = dd <= wd[u];       // rare: may enter the list
                const int dkey = REFORDER ? mih_keys_warp(enter, qa[u], qb[u], a, b) : 0;
                if (enter) {
                    const int key = (dd << KEY_SHIFT) | dkey;
                    if (key < bd[u][K - 1]) {
                        knn_insert<K>(bd[u], bi[u], key, t0 + j);
                        // index order: an equal distance never displaces an earlier row, so the bar can be one lower
                        wd[u] = min(max_dist, (bd[u][K - 1] >> KEY_SHIFT) - (REFORDER ? 0 : 1));
                    }
                }
            }
        }
    }
#pragma unroll
    for (int u = 0; u < QPT; ++u) {
        if (qi[u] < nq) {
            size_t o = ((size_t)qi[u] * nsplit + split) * K;
#pragma unroll
            for (int j = 0; j < K; ++j) { pidx[o + j] = bi[u][j]; pdist[o + j] = bd[u][j]; }
        }
    }
}

// merge nsplit*K candidates per query (already ascending inside a split; splits ascend in train index); pdist holds sort keys
__global__ void k_knn_merge(int nq_cap, const int *__restrict__ nq_dev, int nsplit, int K, int k, const int *__restrict__ pidx,
                            const int *__restrict__ pdist, int *__restrict__ idx, int *__restrict__ dist)
{
    const int nq = nq_dev ? min(*nq_dev, nq_cap) : nq_cap;
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    const int *ci = pidx + (size_t)qi * nsplit * K, *cd = pdist + (size_t)qi * nsplit * K;
    int lastd = -1, lasti = -1;
    for (int r = 0; r < k; ++r) {
        int bestd = 0x7fffffff, besti = -1;
        for (int c = 0; c < nsplit * K; ++c) {
            int dd = cd[c], ii = ci[c];
            if (ii < 0) continue;
            // strictly after the previous pick in (dist, idx) order
            if (dd < lastd || (dd == lastd && ii <= lasti)) continue;
            if (dd < bestd || (dd == bestd && ii < besti)) { bestd = dd; besti = ii; }
        }
        idx[(size_t)qi * k + r] = besti;
        dist[(size_t)qi * k + r] = besti >= 0 ? (bestd >> KEY_SHIFT) : -1;
        if (besti < 0) {
            for (int r2 = r + 1; r2 < k; ++r2) { idx[(size_t)qi * k + r2] = -1; dist[(size_t)qi * k + r2] = -1; }
            break;
        }
        lastd = bestd; lasti = besti;
    }
}

// ---- frame-to-frame association: queries = segments of frame f, train = segments of frame f-1 ----------
// (with one GPU and an epoch of one frame the map snapshot is the previous frame.)  One CTA
// per frame; frame 0 matches the carry (last frame of the previous batch).  trainIdx is the index inside
// the previous frame's segment list.
template <int K>
__global__ void __launch_bounds__(128) k_knn_prev(const uint4 *__restrict__ desc, const int *__restrict__ frame_off, int f_begin, int k, int max_dist, int tie_order,
                                                 const uint4 *__restrict__ carry, int carry_n, int *__restrict__ idx,
                                                 int *__restrict__ dist)
{
    __shared__ uint4 tile[128 * 2];
    const int f = f_begin + blockIdx.x;
    const int q0 = frame_off[f], q1 = frame_off[f + 1];
    const uint4 *tp; int nt;
    if (f > 0) { int t0 = frame_off[f - 1]; tp = desc + 2 * (size_t)t0; nt = q0 - t0; }
    else { tp = carry; nt = carry_n; }
    for (int qb = q0; qb < q1; qb += 128) {
        const int qi = qb + threadIdx.x;
        uint4 qa = make_uint4(0, 0, 0, 0), qbv = qa;
        if (qi < q1) { qa = desc[2 * (size_t)qi]; qbv = desc[2 * (size_t)qi + 1]; }
        int bd[K], bi[K];
#pragma unroll
        for (int j = 0; j < K; ++j) { bd[j] = 0x7fffffff; bi[j] = -1; }
        for (int t0 = 0; t0 < nt; t0 += 128) {
            const int m = min(128, nt - t0);
            __syncthreads();
            for (int i = threadIdx.x; i < m * 2; i += 128) tile[i] = tp[2 * (size_t)t0 + i];
            __syncthreads();
            for (int j = 0; j < m; ++j) {
                uint4 a = tile[2 * j], b = tile[2 * j + 1];
                const int dd = hamming256(qa, qbv, a, b);
                const bool enter = dd <= (bd[K - 1] >> KEY_SHIFT) && dd <= max_dist;
                const int dkey = tie_order ? 0 : mih_keys_warp(enter, qa, qbv, a, b);
                if (enter) {
                    const int key = (dd << KEY_SHIFT) | dkey;
                    if (key < bd[K - 1]) knn_insert<K>(bd, bi, key, t0 + j);
                }
            }
        }
        if (qi < q1) {
            for (int j = 0; j < k; ++j) {
                bool ok = j < K && bi[j < K ? j : 0] >= 0;
                idx[(size_t)qi * k + j] = ok ? bi[j] : -1;
                dist[(size_t)qi * k + j] = ok ? (bd[j] >> KEY_SHIFT) : -1;
            }
        }
    }
}

void launch_knn_prev(const u8 *desc, const int *frame_off, int f_begin, int n, int k, int max_dist, int tie_order, const u8 *carry,
                     int carry_n, int *idx, int *dist, cudaStream_t st)
{
    const uint4 *d4 = (const uint4 *)desc, *c4 = (const uint4 *)carry;
    if (k <= 1) k_knn_prev<1><<<n, 128, 0, st>>>(d4, frame_off, f_begin, k, max_dist, tie_order, c4, carry_n, idx, dist);
    else if (k <= 2) k_knn_prev<2><<<n, 128, 0, st>>>(d4, frame_off, f_begin, k, max_dist, tie_order, c4, carry_n, idx, dist);
    else if (k <= 4) k_knn_prev<4><<<n, 128, 0, st>>>(d4, frame_off, f_begin, k, max_dist, tie_order, c4, carry_n, idx, dist);
    else k_knn_prev<8><<<n, 128, 0, st>>>(d4, frame_off, f_begin, k, max_dist, tie_order, c4, carry_n, idx, dist);
    ++g_launches;
}

static int knn_K(int k) { return k <= 1 ? 1 : k <= 2 ? 2 : k <= 4 ? 4 : 8; }

static int knn_nsplit(int nq, int nm)
{
    int qtiles = (nq + KT * QPT - 1) / (KT * QPT);
    if (qtiles < 1) qtiles = 1;
    int ns = (148 * 4 + qtiles - 1) / qtiles;
    int maxs = (nm + KTILE * 2 - 1) / (KTILE * 2);
    if (ns > maxs) ns = maxs;
    if (ns < 1) ns = 1;
    return ns;
}

size_t knn_scratch_bytes(int nq, int nm, int k)
{
    return (size_t)nq * knn_nsplit(nq, nm) * knn_K(k) * 2 * sizeof(int);
}

void launch_knn(const u8 *q, int nq_cap, const int *nq_dev, const u8 *m, int nm, int k, int max_dist, int tie_order, int *idx, int *dist,
                void *scratch, size_t, cudaStream_t st)
{
    if (nq_cap <= 0) return;
    const int K = knn_K(k), nsplit = knn_nsplit(nq_cap, nm);
    int chunk = (nm + nsplit - 1) / nsplit;
    chunk = ((chunk + KTILE - 1) / KTILE) * KTILE;
    int *pidx = (int *)scratch, *pdist = pidx + (size_t)nq_cap * nsplit * K;
    dim3 grid((nq_cap + KT * QPT - 1) / (KT * QPT), nsplit);
    const uint4 *q4 = (const uint4 *)q, *m4 = (const uint4 *)m;
    switch (K) {
    case 1: (tie_order ? k_knn_partial<1, false> : k_knn_partial<1, true>)<<<grid, KT, 0, st>>>(q4, nq_cap, nq_dev, m4, nm, max_dist, chunk, pidx, pdist); break;
    case 2: (tie_order ? k_knn_partial<2, false> : k_knn_partial<2, true>)<<<grid, KT, 0, st>>>(q4, nq_cap, nq_dev, m4, nm, max_dist, chunk, pidx, pdist); break;
    case 4: (tie_order ? k_knn_partial<4, false> : k_knn_partial<4, true>)<<<grid, KT, 0, st>>>(q4, nq_cap, nq_dev, m4, nm, max_dist, chunk, pidx, pdist); break;
    default: (tie_order ? k_knn_partial<8, false> : k_knn_partial<8, true>)<<<grid, KT, 0, st>>>(q4, nq_cap, nq_dev, m4, nm, max_dist, chunk, pidx, pdist); break;
    }
    ++g_launches;
    k_knn_merge<<<(nq_cap + 127) / 128, 128, 0, st>>>(nq_cap, nq_dev, nsplit, K, k, pidx, pdist, idx, dist);
    ++g_launches;
}

}  // namespace lsf
