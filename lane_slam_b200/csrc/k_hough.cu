// k_hough.cu -- the alternative detector LineDetectorHSV (src/line_detector/include/line_detector/line_detector1.py:11-137, selected by
// eight of the ten shipped line_detector_node YAML files) on the GPU.  setImage and the colour filter are the LSD detector's and
// are already on the device after a batch (bit-planes: dilated colour masks bw_c, edge_color_c = bw_c & Canny edges); this file
// adds what differs: cv2.HoughLinesP on every edge_color image (hough_core.cuh, bit-identical to OpenCV's progressive
// probabilistic Hough transform) and _findNormal / _correctPixelOrdering / toSegmentMsg in the reference's float64 arithmetic.
//   k_hough_p          one warp per (frame, colour): collect the edge points in raster order, then the sequential transform with
//                      the 180 accumulator updates of every vote dealt to the lanes
//   k_hough_segments   one thread per line: normal, endpoint order, centre, normalised pixels
// and lsf_hough_batch (the C ABI entry): runs on the maps of the LAST lsf_front_end_batch / _jpeg call of the ctx.
//
// STATUS (round 2): hough_core.cuh is verified against cv2.HoughLinesP and against the reference's own class on the host
// (tests/test_hough_core.py).  The kernels below were written after the round's GPU budget was spent: their first run on a GPU is
// the test tests/test_gpu_parity.py::test_hough_detector_isolated, which runs them in a process of its own.
#include <algorithm>
#include <vector>

#include "ctx.cuh"
#include "hough_core.cuh"

namespace lsf {

constexpr int HOUGH_WARPS_PER_BLOCK = 4;

struct HoughState {
    int h, w, n_cap, max_lines, nwarps;
    float *trig;                 // [180][2]
    int32_t *accum;              // [nwarps][180 * numrho]
    u8 *mask;                    // [nwarps][h * w]
    u32 *nzloc;                  // [nwarps][h * w]
    int32_t *raw;                // [n_cap * 3][max_lines][4] lines in the order HoughLinesP returns them
    int *count;                  // [n_cap * 3] lines per (frame, colour); [n_cap * 3] = task counter, [n_cap * 3 + 1] = overflow flag
    int *offset;                 // [n_cap * 3 + 1] first output row of every (frame, colour)
    // output rows (frame order, white / yellow / red inside a frame)
    u8 *o_color; float *o_lines; double *o_normals; float *o_centers; float *o_pixn; float *o_nf32; double *o_ground; u8 *o_keep;
    int *h_count;                // pinned [n_cap * 3 + 2]
    size_t rows_cap;
};

__global__ void __launch_bounds__(HOUGH_WARPS_PER_BLOCK * 32) k_hough_p(int h, int w, int wp, int n_tasks, const u32 *__restrict__ planesB,
                                                                        int threshold, int line_length, int line_gap,
                                                                        const float *__restrict__ trig, int32_t *__restrict__ accum_all,
                                                                        u8 *__restrict__ mask_all, u32 *__restrict__ nzloc_all,
                                                                        int32_t *__restrict__ raw_all, int max_lines, int *__restrict__ count)
{
    const int lane = threadIdx.x & 31;
    const int wid = blockIdx.x * HOUGH_WARPS_PER_BLOCK + (threadIdx.x >> 5);
    const int numrho = (w + h) * 2 + 1;
    const size_t acc_sz = (size_t)hp::NUMANGLE * numrho, npix = (size_t)h * w, ps = (size_t)h * wp;
    int32_t *accum = accum_all + (size_t)wid * acc_sz;
    u8 *mask = mask_all + (size_t)wid * npix;
    u32 *nz = nzloc_all + (size_t)wid * npix;
    int *task_ctr = count + n_tasks, *overflow = count + n_tasks + 1;
    for (;;) {
        int task = 0;
        if (lane == 0) task = atomicAdd(task_ctr, 1);
        task = __shfl_sync(0xffffffffu, task, 0);
        if (task >= n_tasks) break;
        const int img = task / 3, c = task - img * 3;
        const u32 *plane = planesB + ((size_t)img * PB_COUNT + PB_EC0 + c) * ps;
        const int cnt = hp::collect(plane, h, w, wp, accum, acc_sz, mask, nz);
        hp::Task t;
        t.width = w; t.height = h; t.threshold = threshold; t.line_length = line_length; t.line_gap = line_gap;
        t.numrho = numrho; t.trig = trig; t.accum = accum; t.mask = mask; t.nzloc = nz; t.count = cnt;
        t.lines = raw_all + (size_t)task * max_lines * 4; t.max_lines = max_lines;
        const int nl = hp::hough_lines_p(t);
        if (lane == 0) {
            count[task] = min(nl, max_lines);
            if (nl > max_lines) atomicMax(overflow, nl);
        }
        __syncwarp();
    }
}

struct PlaneMask {
    const u32 *p; int wp;
    __host__ __device__ bool operator()(int y, int x) const { return (p[(size_t)y * wp + (x >> 5)] >> (x & 31)) & 1u; }
};

__global__ void __launch_bounds__(256) k_hough_segments(int h, int w, int wp, int n_tasks, const u32 *__restrict__ planesB,
                                                       const int32_t *__restrict__ raw_all, int max_lines, const int *__restrict__ count,
                                                       const int *__restrict__ offset, int top_cutoff, double inv_w, double inv_h,
                                                       u8 *__restrict__ o_color, float *__restrict__ o_lines, double *__restrict__ o_normals,
                                                       float *__restrict__ o_centers, float *__restrict__ o_pixn, float *__restrict__ o_nf32)
{
    const int task = blockIdx.x;
    if (task >= n_tasks) return;
    const int nl = count[task];
    const size_t ps = (size_t)h * wp;
    const int img = task / 3, c = task - img * 3;
    const PlaneMask bw = {planesB + ((size_t)img * PB_COUNT + PB_BW0 + c) * ps, wp};
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < nl; i += gridDim.y * blockDim.x) {
        const int32_t *r = raw_all + ((size_t)task * max_lines + i) * 4;
        int32_t line[4];
        double nrm[2], ctr[2];
        hp::find_normal(r[0], r[1], r[2], r[3], h, w, bw, line, nrm, ctr);
        float pixn[4], nf[2];
        hp::normalized_fields(line, nrm, top_cutoff, inv_w, inv_h, pixn, nf);
        const size_t o = (size_t)offset[task] + i;
        o_color[o] = (u8)c;
        for (int k = 0; k < 4; ++k) { o_lines[o * 4 + k] = (float)line[k]; o_pixn[o * 4 + k] = pixn[k]; }
        o_normals[o * 2] = nrm[0]; o_normals[o * 2 + 1] = nrm[1];
        o_centers[o * 2] = (float)ctr[0]; o_centers[o * 2 + 1] = (float)ctr[1];
        o_nf32[o * 2] = nf[0]; o_nf32[o * 2 + 1] = nf[1];
    }
}

}  // namespace lsf

using namespace lsf;

void hough_destroy(lsf_ctx *ctx)
{
    HoughState *s = (HoughState *)ctx->hough;
    if (!s) return;
    for (void *p : {(void *)s->trig, (void *)s->accum, (void *)s->mask, (void *)s->nzloc, (void *)s->raw, (void *)s->count, (void *)s->offset,
                    (void *)s->o_color, (void *)s->o_lines, (void *)s->o_normals, (void *)s->o_centers, (void *)s->o_pixn, (void *)s->o_nf32,
                    (void *)s->o_ground, (void *)s->o_keep})
        if (p) cudaFree(p);
    if (s->h_count) cudaFreeHost(s->h_count);
    delete s;
    ctx->hough = nullptr;
}

static int hough_reserve(lsf_ctx *ctx, int n, int h, int w)
{
    HoughState *s = (HoughState *)ctx->hough;
    if (s && s->h == h && s->w == w && s->n_cap >= n) return LSF_OK;
    hough_destroy(ctx);
    s = new HoughState();
    memset(s, 0, sizeof(*s));
    ctx->hough = s;
    int sms = 148;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
    s->h = h; s->w = w; s->n_cap = std::max(n, ctx->max_batch);
    s->nwarps = sms * HOUGH_WARPS_PER_BLOCK;
    s->max_lines = std::max(4096, h * w / 64);
    const size_t acc_sz = (size_t)hp::NUMANGLE * hp::numrho(w, h), npix = (size_t)h * w, tasks = (size_t)s->n_cap * 3;
    s->rows_cap = tasks * s->max_lines;
    CK(cudaMalloc((void **)&s->trig, hp::NUMANGLE * 2 * sizeof(float)));
    CK(cudaMalloc((void **)&s->accum, (size_t)s->nwarps * acc_sz * sizeof(int32_t)));
    CK(cudaMalloc((void **)&s->mask, (size_t)s->nwarps * npix));
    CK(cudaMalloc((void **)&s->nzloc, (size_t)s->nwarps * npix * sizeof(u32)));
    CK(cudaMalloc((void **)&s->raw, s->rows_cap * 4 * sizeof(int32_t)));
    CK(cudaMalloc((void **)&s->count, (tasks + 2) * sizeof(int)));
    CK(cudaMalloc((void **)&s->offset, (tasks + 1) * sizeof(int)));
    CK(cudaMalloc((void **)&s->o_color, s->rows_cap));
    CK(cudaMalloc((void **)&s->o_lines, s->rows_cap * 4 * sizeof(float)));
    CK(cudaMalloc((void **)&s->o_normals, s->rows_cap * 2 * sizeof(double)));
    CK(cudaMalloc((void **)&s->o_centers, s->rows_cap * 2 * sizeof(float)));
    CK(cudaMalloc((void **)&s->o_pixn, s->rows_cap * 4 * sizeof(float)));
    CK(cudaMalloc((void **)&s->o_nf32, s->rows_cap * 2 * sizeof(float)));
    CK(cudaMalloc((void **)&s->o_ground, s->rows_cap * 4 * sizeof(double)));
    CK(cudaMalloc((void **)&s->o_keep, s->rows_cap));
    CK(cudaMallocHost((void **)&s->h_count, (tasks + 2) * sizeof(int)));
    float trig[hp::NUMANGLE * 2];
    hp::make_trig(trig);                       // the host's libm, like OpenCV's own table
    CK(cudaMemcpy(s->trig, trig, sizeof(trig), cudaMemcpyHostToDevice));
    return LSF_OK;
}

// LineDetectorHSV.detectLines for the three colours of every frame of the ctx's LAST batch (its colour masks and edges are still
// on the device), then -- with project_ground != 0 -- ground projection and line sanity of the segments, like LSF_STAGE_GROUND.
// hough_threshold / hough_min_line_length / hough_max_line_gap: the YAML keys (line_detector1.py:64).  Rows: frame by frame, white /
// yellow / red inside a frame, HoughLinesP's order inside a colour; lines_px holds the (integer) endpoints after ordering.
extern "C" int lsf_hough_batch(lsf_ctx *ctx, int hough_threshold, int hough_min_line_length, int hough_max_line_gap, int project_ground,
                               lsf_segments *out)
{
    if (!ctx) return LSF_E_ARG;
    if (!out) return fail(ctx, LSF_E_ARG, "lsf_hough_batch: null argument");
    if (!ctx->have_batch) return fail(ctx, LSF_E_ARG, "lsf_hough_batch: run lsf_front_end_batch (LSF_STAGE_DETECT) on the frames first");
    if (hough_threshold < 1 || hough_min_line_length < 1 || hough_max_line_gap < 0)
        return fail(ctx, LSF_E_CONFIG, "lsf_hough_batch: hough_threshold >= 1, hough_min_line_length >= 1, hough_max_line_gap >= 0");
    ENTER(ctx);
    const Dims d = ctx->d;
    const int n = d.n, tasks = n * 3;
    if (d.h > 65535 || d.w > 65535) return fail(ctx, LSF_E_CAPACITY, "lsf_hough_batch: image larger than 65535 pixels a side");
    int rc = hough_reserve(ctx, n, d.h, d.w);
    if (rc) return rc;
    HoughState *s = (HoughState *)ctx->hough;
    cudaStream_t st = ctx->st;
    ctx->n_events = 0;
    mark(ctx, "start");
    CK(cudaMemsetAsync(s->count, 0, (size_t)(tasks + 2) * sizeof(int), st));
    k_hough_p<<<s->nwarps / HOUGH_WARPS_PER_BLOCK, HOUGH_WARPS_PER_BLOCK * 32, 0, st>>>(d.h, d.w, d.wp, tasks, ctx->b.planesB, hough_threshold,
                                                                                   hough_min_line_length, hough_max_line_gap, s->trig, s->accum,
                                                                                   s->mask, s->nzloc, s->raw, s->max_lines, s->count);
    ++g_launches;
    mark(ctx, "hough_p");
    CK(cudaMemcpyAsync(s->h_count, s->count, (size_t)(tasks + 2) * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    if (s->h_count[tasks + 1])
        return fail(ctx, LSF_E_CAPACITY, "lsf_hough_batch: a colour image holds " + std::to_string(s->h_count[tasks + 1]) + " lines, more than " +
                                             std::to_string(s->max_lines));
    std::vector<int> off(tasks + 1, 0);
    for (int t = 0; t < tasks; ++t) off[t + 1] = off[t] + s->h_count[t];
    const int S = off[tasks];
    out->n_frames = n; out->n_segments = S;
    if (S > out->capacity) return fail(ctx, LSF_E_CAPACITY, "lsf_segments.capacity too small: need " + std::to_string(S));
    CK(cudaMemcpyAsync(s->offset, off.data(), (size_t)(tasks + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
    if (S > 0) {
        k_hough_segments<<<dim3(tasks, 4), 256, 0, st>>>(d.h, d.w, d.wp, tasks, ctx->b.planesB, s->raw, s->max_lines, s->count, s->offset, d.top,
                                                        1.0 / (double)d.dw, 1.0 / (double)d.dh, s->o_color, s->o_lines, s->o_normals, s->o_centers,
                                                        s->o_pixn, s->o_nf32);
        ++g_launches;
        mark(ctx, "hough_segments");
        if (project_ground) {
            launch_project_filter(ctx->cam, s->o_pixn, s->o_color, S, s->o_ground, s->o_keep, st);
            mark(ctx, "ground");
        }
    }
    const cudaMemcpyKind kind = out_kind(out->mem);
    std::vector<int> fo(n + 1);
    for (int f = 0; f <= n; ++f) fo[f] = off[f * 3];
    if (out->mem == LSF_MEM_DEVICE) {
        if (out->counts) CK(cudaMemcpyAsync(out->counts, s->count, (size_t)tasks * sizeof(int), cudaMemcpyDeviceToDevice, st));
        if (out->frame_offset) CK(cudaMemcpyAsync(out->frame_offset, fo.data(), (size_t)(n + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
    } else {
        if (out->counts) memcpy(out->counts, s->h_count, (size_t)tasks * sizeof(int));
        if (out->frame_offset) memcpy(out->frame_offset, fo.data(), (size_t)(n + 1) * sizeof(int));
    }
    if (S > 0) {
        const size_t r = S;
#define COPY(dst, srcp, bytes) do { if (dst) CK(cudaMemcpyAsync(dst, srcp, (bytes), kind, st)); } while (0)
        COPY(out->color, s->o_color, r);
        COPY(out->lines_px, s->o_lines, r * 16);
        COPY(out->normals, s->o_normals, r * 16);
        COPY(out->centers, s->o_centers, r * 8);
        COPY(out->pixels_normalized, s->o_pixn, r * 16);
        COPY(out->normal_f32, s->o_nf32, r * 8);
        if (project_ground) { COPY(out->ground, s->o_ground, r * 32); COPY(out->keep, s->o_keep, r); }
#undef COPY
    }
    mark(ctx, "d2h");
    CK(cudaStreamSynchronize(st));       // also keeps fo / off alive until the copies that read them are done
    CK(cudaGetLastError());
    return LSF_OK;
}
