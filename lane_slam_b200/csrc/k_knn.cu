// k_knn.cu -- K13: exact brute-force Hamming kNN over packed 256-bit descriptors (__popc).
//
// Replaces BinaryDescriptorMatcher::knnMatch (src/line_descriptor/src/binary_descriptor_matcher.cpp:258-335,
// Mihasher::query :635-753) and the distance kernel match() (src/line_descriptor/src/bitops_custom.hpp:83-96).
// Contract: ascending distance, ties by ascending train index, neighbours farther than
// max_dist are not reported (-1).  The multi-index hash is a CPU-side accelerator of the same exact search
// and is not reproduced.
//
// Each thread keeps one query (8 x 32-bit words) in registers; map descriptors are staged through shared
// memory and broadcast to the warp.  The map is split across gridDim.y so the grid fills the 148 SMs;
// a second kernel merges the per-split candidates.
#include "common.cuh"

namespace lsf {

constexpr int KT = 256;        // threads = queries per CTA
constexpr int KTILE = 256;     // map descriptors per shared-memory tile (8 KB)

template <int K>
__device__ __forceinline__ void knn_insert(int (&bd)[K], int (&bi)[K], int dd, int idx)
{
    // caller guarantees dd < bd[K-1] (strict: an equal distance keeps the earlier, smaller index)
#pragma unroll
    for (int j = K - 1; j > 0; --j) {
        if (bd[j - 1] > dd) { bd[j] = bd[j - 1]; bi[j] = bi[j - 1]; }
        else if (bd[j] > dd) { bd[j] = dd; bi[j] = idx; }
    }
    if (bd[0] > dd) { bd[0] = dd; bi[0] = idx; }
}

template <int K>
__global__ void __launch_bounds__(KT) k_knn_partial(const uint4 *__restrict__ q, int nq_cap, const int *__restrict__ nq_dev,
                                                   const uint4 *__restrict__ m, int nm, int max_dist, int chunk,
                                                   int *__restrict__ pidx, int *__restrict__ pdist)
{
    __shared__ uint4 tile[KTILE * 2];
    const int nq = nq_dev ? min(*nq_dev, nq_cap) : nq_cap;
    const int qi = blockIdx.x * KT + threadIdx.x;
    if (blockIdx.x * KT >= nq) return;
    const int split = blockIdx.y, nsplit = gridDim.y;
    const int m0 = split * chunk, m1 = min(nm, m0 + chunk);
    uint4 qa = make_uint4(0, 0, 0, 0), qb = qa;
    if (qi < nq) { qa = q[2 * (size_t)qi]; qb = q[2 * (size_t)qi + 1]; }
    int bd[K], bi[K];
#pragma unroll
    for (int j = 0; j < K; ++j) { bd[j] = 0x7fffffff; bi[j] = -1; }
    for (int t0 = m0; t0 < m1; t0 += KTILE) {
        const int nt = min(KTILE, m1 - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < nt * 2; i += KT) tile[i] = m[2 * (size_t)t0 + i];
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < nt; ++j) {
            uint4 a = tile[2 * j], b = tile[2 * j + 1];
            int dd = __popc(qa.x ^ a.x) + __popc(qa.y ^ a.y) + __popc(qa.z ^ a.z) + __popc(qa.w ^ a.w) +
                     __popc(qb.x ^ b.x) + __popc(qb.y ^ b.y) + __popc(qb.z ^ b.z) + __popc(qb.w ^ b.w);
            if (dd < bd[K - 1] && dd <= max_dist) knn_insert<K>(bd, bi, dd, t0 + j);
        }
    }
    if (qi < nq) {
        size_t o = ((size_t)qi * nsplit + split) * K;
#pragma unroll
        for (int j = 0; j < K; ++j) { pidx[o + j] = bi[j]; pdist[o + j] = bd[j]; }
    }
}

// merge nsplit*K candidates per query (already ascending inside a split; splits ascend in train index)
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
        dist[(size_t)qi * k + r] = besti >= 0 ? bestd : -1;
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
__global__ void __launch_bounds__(128) k_knn_prev(const uint4 *__restrict__ desc, const int *__restrict__ frame_off, int f_begin, int k, int max_dist,
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
                int dd = __popc(qa.x ^ a.x) + __popc(qa.y ^ a.y) + __popc(qa.z ^ a.z) + __popc(qa.w ^ a.w) +
                         __popc(qbv.x ^ b.x) + __popc(qbv.y ^ b.y) + __popc(qbv.z ^ b.z) + __popc(qbv.w ^ b.w);
                if (dd < bd[K - 1] && dd <= max_dist) knn_insert<K>(bd, bi, dd, t0 + j);
            }
        }
        if (qi < q1) {
            for (int j = 0; j < k; ++j) {
                bool ok = j < K && bi[j < K ? j : 0] >= 0;
                idx[(size_t)qi * k + j] = ok ? bi[j] : -1;
                dist[(size_t)qi * k + j] = ok ? bd[j] : -1;
            }
        }
    }
}

void launch_knn_prev(const u8 *desc, const int *frame_off, int f_begin, int n, int k, int max_dist, const u8 *carry, int carry_n,
                     int *idx, int *dist, cudaStream_t st)
{
    const uint4 *d4 = (const uint4 *)desc, *c4 = (const uint4 *)carry;
    if (k <= 1) k_knn_prev<1><<<n, 128, 0, st>>>(d4, frame_off, f_begin, k, max_dist, c4, carry_n, idx, dist);
    else if (k <= 2) k_knn_prev<2><<<n, 128, 0, st>>>(d4, frame_off, f_begin, k, max_dist, c4, carry_n, idx, dist);
    else if (k <= 4) k_knn_prev<4><<<n, 128, 0, st>>>(d4, frame_off, f_begin, k, max_dist, c4, carry_n, idx, dist);
    else k_knn_prev<8><<<n, 128, 0, st>>>(d4, frame_off, f_begin, k, max_dist, c4, carry_n, idx, dist);
    ++g_launches;
}

static int knn_K(int k) { return k <= 1 ? 1 : k <= 2 ? 2 : k <= 4 ? 4 : 8; }

static int knn_nsplit(int nq, int nm)
{
    int qtiles = (nq + KT - 1) / KT;
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

void launch_knn(const u8 *q, int nq_cap, const int *nq_dev, const u8 *m, int nm, int k, int max_dist, int *idx, int *dist,
                void *scratch, size_t, cudaStream_t st)
{
    if (nq_cap <= 0) return;
    const int K = knn_K(k), nsplit = knn_nsplit(nq_cap, nm);
    int chunk = (nm + nsplit - 1) / nsplit;
    chunk = ((chunk + KTILE - 1) / KTILE) * KTILE;
    int *pidx = (int *)scratch, *pdist = pidx + (size_t)nq_cap * nsplit * K;
    dim3 grid((nq_cap + KT - 1) / KT, nsplit);
    const uint4 *q4 = (const uint4 *)q, *m4 = (const uint4 *)m;
    switch (K) {
    case 1: k_knn_partial<1><<<grid, KT, 0, st>>>(q4, nq_cap, nq_dev, m4, nm, max_dist, chunk, pidx, pdist); break;
    case 2: k_knn_partial<2><<<grid, KT, 0, st>>>(q4, nq_cap, nq_dev, m4, nm, max_dist, chunk, pidx, pdist); break;
    case 4: k_knn_partial<4><<<grid, KT, 0, st>>>(q4, nq_cap, nq_dev, m4, nm, max_dist, chunk, pidx, pdist); break;
    default: k_knn_partial<8><<<grid, KT, 0, st>>>(q4, nq_cap, nq_dev, m4, nm, max_dist, chunk, pidx, pdist); break;
    }
    ++g_launches;
    k_knn_merge<<<(nq_cap + 127) / 128, 128, 0, st>>>(nq_cap, nq_dev, nsplit, K, k, pidx, pdist, idx, dist);
    ++g_launches;
}

}  // namespace lsf
