// k_lane_filter.cu -- the histogram lane filter on the ground segments of a batch: per frame
//   predict(dt, v, w)   LaneFilterHistogram.predict   src/lane_filter/include/lane_filter/lane_filter.py:47-72
//   update(segments)    .update / .generate_measurement_likelihood                         :75-102 (votes: k_lane_votes)
//   getEstimate / getMax                                                                    :104-112
// as the node does for every SegmentList (src/lane_filter/src/lane_filter_node.py:53-65).  The belief of frame t depends on
// the belief of frame t-1, so the batch is a chain: ONE CTA walks the frames, its threads own the histogram cells
// (23 x 30 by default).  Every float64 operation is done in the order numpy / scipy.ndimage do it, so beliefs and estimates
// are bit-identical to the reference class:
//   * process model: sources visited in raster order, p_belief[target] += belief[source]  -> every target cell adds its
//     sources in raster order;
//   * scipy.ndimage.gaussian_filter(mode='constant'): axis 0 then axis 1, symmetric correlate1d:
//     tmp = x[l] w[0]; for j = r .. 1: tmp += (x[l - j] + x[l + j]) w[j];
//   * np.sum over the (contiguous) histogram: numpy's pairwise summation -- blocks of <= 128 elements summed with 8
//     strided accumulators, combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), halves split at a multiple of 8; the host passes
//     the leaf table and the combine order for the cell count;
//   * transcendental inputs (sin(phi) of the grid, Gaussian mask weights, the prior) come from the host: numpy / scipy there
//     are what the reference calls.
#include "common.cuh"

namespace lsf {

constexpr int LF_THREADS = 1024;

struct LfSm {
    double *bel, *a, *b, *leafr;     // [ncell] x3, [nleaf][8]
    int *tgt;                        // [ncell]
};

// numpy pairwise sum of x[0 .. ncell) with the host-made plan; result broadcast to all threads
__device__ double lf_pairwise(const double *x, const LaneFilterPlan &pl, double *leafr, double *leafsum)
{
    const int t = threadIdx.x;
    __syncthreads();
    if (t < pl.nleaf * 8) {
        const int leaf = t >> 3, j = t & 7, off = pl.leaf_off[leaf], len = pl.leaf_len[leaf];
        if (len >= 8) {
            double r = x[off + j];
            for (int i = 8; i < len - (len % 8); i += 8) r = __dadd_rn(r, x[off + i + j]);
            leafr[t] = r;
        }
    }
    __syncthreads();
    if (t < pl.nleaf) {
        const int off = pl.leaf_off[t], len = pl.leaf_len[t];
        double res;
        if (len < 8) {
            res = 0.0;
            for (int i = 0; i < len; ++i) res = __dadd_rn(res, x[off + i]);
        } else {
            const double *r = leafr + 8 * t;
            res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])), __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
            for (int i = len - (len % 8); i < len; ++i) res = __dadd_rn(res, x[off + i]);
        }
        leafsum[t] = res;
    }
    __syncthreads();
    if (t == 0)
        for (int k = 0; k < pl.ncomb; ++k) leafsum[pl.comb_a[k]] = __dadd_rn(leafsum[pl.comb_a[k]], leafsum[pl.comb_b[k]]);
    __syncthreads();
    const double s = leafsum[0];
    __syncthreads();
    return s;
}

__global__ void __launch_bounds__(LF_THREADS) k_lane_filter(LaneFilterPlan pl, int n_frames, int use_propagation,
                                                           const double *__restrict__ dt_v_w, const int *__restrict__ hist,
                                                           const double *__restrict__ d_grid, const double *__restrict__ phi_grid,
                                                           const double *__restrict__ sin_phi, const double *__restrict__ w_d,
                                                           const double *__restrict__ w_phi, double *__restrict__ belief,
                                                           double *__restrict__ est)
{
    extern __shared__ __align__(16) unsigned char lf_raw[];
    const int nc = pl.nd * pl.nphi, t = threadIdx.x;
    double *bel = reinterpret_cast<double *>(lf_raw), *pa = bel + nc, *pb = pa + nc, *leafr = pb + nc, *leafsum = leafr + 8 * LF_MAX_LEAVES;
    int *tgt = reinterpret_cast<int *>(leafsum + LF_MAX_LEAVES);
    __shared__ double s_best;
    __shared__ int s_besti;
    const bool cell = t < nc;
    const int ci = cell ? t / pl.nphi : 0, cj = cell ? t - ci * pl.nphi : 0;
    if (cell) bel[t] = belief[t];
    __syncthreads();
    for (int f = 0; f < n_frames; ++f) {
        if (use_propagation) {
            const double dt = dt_v_w[3 * f], v = dt_v_w[3 * f + 1], w = dt_v_w[3 * f + 2];
            const double vdt = __dmul_rn(v, dt), wdt = __dmul_rn(w, dt);
            if (cell) {
                int tg = -1;
                if (bel[t] > 0) {
                    const double d_t = __dadd_rn(d_grid[t], __dmul_rn(vdt, sin_phi[t])), phi_t = __dadd_rn(phi_grid[t], wdt);
                    if (!(d_t > pl.d_max || d_t < pl.d_min || phi_t < pl.phi_min || phi_t > pl.phi_max)) {
                        const int in = (int)floor(__ddiv_rn(__dsub_rn(d_t, pl.d_min), pl.delta_d));
                        const int jn = (int)floor(__ddiv_rn(__dsub_rn(phi_t, pl.phi_min), pl.delta_phi));
                        if (in >= 0 && in < pl.nd && jn >= 0 && jn < pl.nphi) tg = in * pl.nphi + jn;   // (on the upper edge numpy raises IndexError)
                    }
                }
                tgt[t] = tg;
            }
            __syncthreads();
            if (cell) {
                double acc = 0.0;
                for (int s = 0; s < nc; ++s)
                    if (tgt[s] == t) acc = __dadd_rn(acc, bel[s]);
                pa[t] = acc;
            }
            __syncthreads();
            if (cell) {     // axis 0 (d)
                double tmp = __dmul_rn(pa[t], w_d[0]);
                for (int j = pl.r_d; j >= 1; --j) {
                    const double lo = ci - j >= 0 ? pa[t - j * pl.nphi] : 0.0, hi = ci + j < pl.nd ? pa[t + j * pl.nphi] : 0.0;
                    tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn(lo, hi), w_d[j]));
                }
                pb[t] = tmp;
            }
            __syncthreads();
            if (cell) {     // axis 1 (phi)
                double tmp = __dmul_rn(pb[t], w_phi[0]);
                for (int j = pl.r_phi; j >= 1; --j) {
                    const double lo = cj - j >= 0 ? pb[t - j] : 0.0, hi = cj + j < pl.nphi ? pb[t + j] : 0.0;
                    tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn(lo, hi), w_phi[j]));
                }
                pa[t] = tmp;
            }
            const double tot = lf_pairwise(pa, pl, leafr, leafsum);
            if (tot != 0.0 && cell) bel[t] = __ddiv_rn(pa[t], tot);
            __syncthreads();
        }
        // update with the votes of frame f
        {
            const int h = cell ? hist[(size_t)f * nc + t] : 0;
            const int votes = __syncthreads_count(h != 0) ? 1 : 0;          // any vote at all
            if (votes) {
                if (cell) pb[t] = (double)h;
                const double nv = lf_pairwise(pb, pl, leafr, leafsum);       // np.sum of the counts (exact)
                if (cell) { pb[t] = __ddiv_rn(pb[t], nv); pa[t] = __dmul_rn(bel[t], pb[t]); }
                const double tot = lf_pairwise(pa, pl, leafr, leafsum);
                if (cell) bel[t] = tot == 0.0 ? pb[t] : __ddiv_rn(pa[t], tot);
                __syncthreads();
            }
        }
        // estimate: first maximum in raster order
        if (t == 0) { s_best = -1.0; s_besti = 0; }
        __syncthreads();
        {
            double v = cell ? bel[t] : -1.0;
            int idx = t;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, v, o);
                const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
                if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
            }
            if ((t & 31) == 0) pa[nc > 32 ? (t >> 5) : 0] = v, tgt[t >> 5] = idx;      // pa / tgt are free here
        }
        __syncthreads();
        if (t == 0) {
            double v = -1.0; int idx = 0;
            for (int wv = 0; wv < LF_THREADS / 32; ++wv)
                if (pa[wv] > v || (pa[wv] == v && tgt[wv] < idx)) { v = pa[wv]; idx = tgt[wv]; }
            const int bi = idx / pl.nphi, bj = idx - bi * pl.nphi;
            est[3 * f] = __dadd_rn(pl.d_min, __dmul_rn(__dadd_rn((double)bi, 0.5), pl.delta_d));
            est[3 * f + 1] = __dadd_rn(pl.phi_min, __dmul_rn(__dadd_rn((double)bj, 0.5), pl.delta_phi));
            est[3 * f + 2] = v;
        }
        __syncthreads();
    }
    if (cell) belief[t] = bel[t];
}

size_t lane_filter_smem(int ncell) { return (size_t)(3 * ncell + 9 * LF_MAX_LEAVES) * sizeof(double) + (size_t)(ncell > 64 ? ncell : 64) * sizeof(int); }

void launch_lane_filter(const LaneFilterPlan &pl, int n_frames, int use_propagation, const double *dt_v_w, const int *hist,
                        const double *d_grid, const double *phi_grid, const double *sin_phi, const double *w_d, const double *w_phi,
                        double *belief, double *est, cudaStream_t st)
{
    const int nc = pl.nd * pl.nphi;
    const size_t smem = lane_filter_smem(nc);
    static PerDevice attr;
    if (smem > 48 * 1024) attr.ensure(smem, [&] { cudaFuncSetAttribute(k_lane_filter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); });
    k_lane_filter<<<1, LF_THREADS, smem, st>>>(pl, n_frames, use_propagation, dt_v_w, hist, d_grid, phi_grid, sin_phi, w_d, w_phi, belief, est);
    ++g_launches;
}

}  // namespace lsf
