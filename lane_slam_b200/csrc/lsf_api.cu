// lsf_api.cu -- the C ABI (include/lsf.h): context, device scratch, batch pipeline, copies.
#include "ctx.cuh"

namespace lsf {
static thread_local long long t_launch_sink = 0;
thread_local long long *t_launches = &t_launch_sink;
}
using namespace lsf;

std::string g_create_error;

// ---- defaults (the reference's YAML files) --------------------------------------------------------------
extern "C" int lsf_default_config(lsf_config *c)
{
    if (!c) return LSF_E_ARG;
    memset(c, 0, sizeof(*c));
    c->img_h = 120; c->img_w = 160; c->top_cutoff = 40;  // line_detector_node/default.yaml:1-2
    const int lo[4][3] = {{0, 0, 150}, {25, 140, 100}, {0, 140, 100}, {165, 140, 100}};
    const int hi[4][3] = {{180, 60, 255}, {45, 255, 255}, {15, 255, 255}, {180, 255, 255}};
    memcpy(c->hsv_lo, lo, sizeof(lo)); memcpy(c->hsv_hi, hi, sizeof(hi));
    c->dilation_kernel_size = 3; c->canny_lo = 80; c->canny_hi = 200;
    for (int i = 0; i < 3; ++i) { c->ai_scale[i] = 1.f; c->ai_shift[i] = 0.f; }
    const double K[9] = {307.7379294605756, 0, 329.692367951685, 0, 314.9827773443905, 244.4605588877848, 0, 0, 1};
    const double D[5] = {-0.2565888993516047, 0.04481160508242147, -0.00505275149956019, 0.001308569367976665, 0};
    const double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    const double P[12] = {210.1107940673828, 0, 327.2577820024981, 0, 0, 253.8408660888672, 239.9969353923052, 0, 0, 0, 1, 0};
    const double Hg[9] = {-4.89775e-05, -0.0002150858, -0.1818273, 0.00099274, 1.202336e-06, -0.3280241,
                          -0.0004281805, -0.007185673, 1};
    memcpy(c->K, K, sizeof(K)); memcpy(c->D, D, sizeof(D)); memcpy(c->R, R, sizeof(R)); memcpy(c->P, P, sizeof(P));
    memcpy(c->Hgnd, Hg, sizeof(Hg));
    c->cam_w = 640; c->cam_h = 480;
    c->lanewidth = 0.23; c->linewidth_white = 0.05; c->linewidth_yellow = 0.025;
    c->d_min = -0.15; c->d_max = 0.3; c->phi_min = -1.5; c->phi_max = 1.5;
    c->max_batch = 1; c->max_src_h = 480; c->max_src_w = 640;
    return LSF_OK;
}


extern "C" int lsf_create(const lsf_config *cfg, lsf_ctx **out)
{
    if (!cfg || !out) return fail(nullptr, LSF_E_ARG, "lsf_create: null argument");
    *out = nullptr;
    if (cfg->img_h < 16 || cfg->img_w < 16 || cfg->top_cutoff < 0 || cfg->img_h - cfg->top_cutoff < 8)
        return fail(nullptr, LSF_E_CONFIG, "lsf_create: img_size / top_cutoff out of range (need >= 16x16 and >= 8 rows after the cut)");
    if (cfg->img_w > 32768 || cfg->img_h > 32768) return fail(nullptr, LSF_E_CONFIG, "lsf_create: img_size too large (max 32768)");
    if (cfg->dilation_kernel_size != 3 && cfg->dilation_kernel_size != 1)
        return fail(nullptr, LSF_E_CONFIG, "lsf_create: dilation_kernel_size must be 3 (cross) or 1; other ellipse sizes are not implemented");
    if (cfg->canny_lo > cfg->canny_hi) return fail(nullptr, LSF_E_CONFIG, "lsf_create: canny_thresholds must be [low, high]");
    if (cfg->cam_w <= 0 || cfg->cam_h <= 0) return fail(nullptr, LSF_E_CONFIG, "lsf_create: camera size missing");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(nullptr, LSF_E_CUDA, "lsf_create: no CUDA device (this library has no CPU fallback)");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, LSF_E_CONFIG, "lsf_create: bad device ordinal");
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, cfg->device);
    if (prop.major != 10) {
        char buf[256];
        snprintf(buf, sizeof(buf), "lsf_create: device %d is sm_%d%d; this library is built for sm_100a only", cfg->device, prop.major, prop.minor);
        return fail(nullptr, LSF_E_CUDA, buf);
    }
    lsf_ctx *ctx = new lsf_ctx();
    ctx->cfg = *cfg;
    ctx->device = cfg->device;
    ctx->have_batch = false;
    ctx->map = nullptr; ctx->map_n = 0; ctx->map_cap = 0;
    ctx->map_ground = nullptr; ctx->map_color = nullptr; ctx->map_frame = nullptr; ctx->pose_dev = nullptr; ctx->pose_cap = 0;
    memset(&ctx->ex, 0, sizeof(ctx->ex));
    ctx->lane_filter = nullptr;
    ctx->jpeg = nullptr; ctx->jpeg_last_bytes = 0; ctx->events_keep = false;
    ctx->hough = nullptr;
    ctx->ex.world = 0;
    ctx->knn_scratch = nullptr; ctx->knn_scratch_cap = 0;
    ctx->tap_tmp = nullptr; ctx->tap_cap = 0;
    ctx->seg_in = nullptr; ctx->seg_in_cap = 0;
    ctx->tma.valid = 0; ctx->tma_src = nullptr;
    ctx->n_events = 0;
    ctx->copy_st = nullptr; ctx->ev_begin = nullptr;
    ctx->kept_rec = nullptr; ctx->last_S = 0; ctx->last_stages = 0;
    ctx->stage_buf[0] = ctx->stage_buf[1] = nullptr; ctx->stage_seq = 0;
    for (int i = 0; i < 2; ++i) { ctx->staged[i].valid = false; ctx->staged[i].ev = nullptr; }
    for (int i = 0; i < 8; ++i) ctx->aux[i] = nullptr;
    ctx->carry = nullptr; ctx->carry_n = 0; ctx->carry_cap = 0; ctx->h_small = nullptr; ctx->st = nullptr;
    memset(&ctx->b, 0, sizeof(ctx->b));
    auto bail = [&](int code, const std::string &msg) { g_create_error = msg; lsf_destroy(ctx); return code; };
#define CKC(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return bail(LSF_E_CUDA, std::string(#call " failed: ") + cudaGetErrorString(e_)); \
    } while (0)
    CKC(cudaSetDevice(ctx->device));
    lsf::t_launches = &ctx->launches;
    CKC(cudaStreamCreateWithFlags(&ctx->st, cudaStreamNonBlocking));
    CKC(cudaStreamCreateWithFlags(&ctx->copy_st, cudaStreamNonBlocking));
    for (int i = 0; i < 8; ++i) CKC(cudaStreamCreateWithFlags(&ctx->aux[i], cudaStreamNonBlocking));
    CKC(cudaEventCreateWithFlags(&ctx->ev_begin, cudaEventDisableTiming));
    ctx->max_batch = cfg->max_batch > 0 ? cfg->max_batch : 1;
    ctx->max_src_h = cfg->max_src_h > 0 ? cfg->max_src_h : cfg->img_h;
    ctx->max_src_w = cfg->max_src_w > 0 ? cfg->max_src_w : cfg->img_w;
    ctx->h = cfg->img_h - cfg->top_cutoff; ctx->w = cfg->img_w;
    ctx->wp = (ctx->w + 31) / 32;
    ctx->sw = (int)lrint(ctx->w * 0.8); ctx->sh = (int)lrint(ctx->h * 0.8);
    ctx->swp = (ctx->sw + 31) / 32;
    const size_t Np = (size_t)ctx->sh * ctx->sw;
    // default: every scaled pixel may be a support pixel for small batches; an eighth of them for large ones (support
    // pixels are 0.3-3 % of the image on lane frames; the USED bitmap of this size lives in shared memory)
    ctx->pixcap = cfg->max_pixels_per_color > 0 ? cfg->max_pixels_per_color
                                                : (int)(ctx->max_batch <= 128 ? Np : std::max<size_t>(4096, Np / 8));
    if ((size_t)ctx->pixcap > Np) ctx->pixcap = (int)Np;
    ctx->pixcap_auto = cfg->max_pixels_per_color <= 0;
    ctx->segcap = cfg->max_segments_per_color > 0 ? cfg->max_segments_per_color
                                                  : (int)std::max<size_t>(512, (size_t)ctx->h * ctx->w / 256);
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 3; ++j) { ctx->cp.lo[i][j] = cfg->hsv_lo[i][j]; ctx->cp.hi[i][j] = cfg->hsv_hi[i][j]; }
    ctx->cp.canny_lo = cfg->canny_lo; ctx->cp.canny_hi = cfg->canny_hi;
    for (int i = 0; i < 3; ++i) { ctx->cp.ai_scale[i] = cfg->ai_scale[i]; ctx->cp.ai_shift[i] = cfg->ai_shift[i]; }
    memcpy(ctx->cam.K, cfg->K, sizeof(cfg->K)); memcpy(ctx->cam.D, cfg->D, sizeof(cfg->D));
    memcpy(ctx->cam.R, cfg->R, sizeof(cfg->R)); memcpy(ctx->cam.P, cfg->P, sizeof(cfg->P));
    memcpy(ctx->cam.Hg, cfg->Hgnd, sizeof(cfg->Hgnd));
    ctx->cam.cam_w = cfg->cam_w; ctx->cam.cam_h = cfg->cam_h;
    ctx->cam.lanewidth = cfg->lanewidth; ctx->cam.lw_white = cfg->linewidth_white; ctx->cam.lw_yellow = cfg->linewidth_yellow;
    ctx->cam.d_min = cfg->d_min; ctx->cam.d_max = cfg->d_max; ctx->cam.phi_min = cfg->phi_min; ctx->cam.phi_max = cfg->phi_max;

    const size_t n = ctx->max_batch, N = (size_t)ctx->h * ctx->w, ps = (size_t)ctx->h * ctx->wp;
    Buffers &b = ctx->b;
    CKC(dalloc(&b.src, n * ctx->max_src_h * ctx->max_src_w * 3));
    ctx->stage_buf[0] = b.src;
    for (int i = 0; i < 2; ++i) CKC(cudaEventCreateWithFlags(&ctx->staged[i].ev, cudaEventDisableTiming));
    CKC(dalloc(&b.ctab, COLOR_TABLE_BYTES));
    {
        u8 tab[COLOR_TABLE_BYTES];
        build_color_tables(ctx->cp, tab);
        CKC(cudaMemcpy(b.ctab, tab, COLOR_TABLE_BYTES, cudaMemcpyHostToDevice));
    }
    CKC(dalloc(&b.planesA, n * PA_COUNT * ps));
    CKC(dalloc(&b.planesB, n * PB_COUNT * ps));
    CKC(dalloc(&b.gray, n * N));
    CKC(dalloc(&b.dx, n * N * 2));
    b.dy = nullptr;
    CKC(dalloc(&b.lsdw, n * 3 * (size_t)ctx->sh * ctx->swp));
    CKC(dalloc(&b.preact, n * 3 * (size_t)((ctx->sh + 7) / 8) * ctx->swp));
    CKC(dalloc(&b.prepatch, n * 3 * (size_t)((ctx->sh + 7) / 8) * ctx->swp * 81));
    CKC(dalloc(&b.prectr, 64));
    CKC(dalloc(&b.pix, n * 3 * ctx->pixcap));
    CKC(dalloc(&b.pxy, n * 3 * ctx->pixcap));
    CKC(dalloc(&b.fat, n * 3 * ctx->pixcap * LSD_FAT_WORDS));
    CKC(dalloc(&b.scs, n * 3 * ctx->pixcap));
    CKC(dalloc(&b.usedbits, n * 3 * (size_t)((ctx->pixcap + 31) / 32)));
    CKC(dalloc(&b.order, n * 3 * ctx->pixcap));
    CKC(dalloc(&b.label, n * 3 * ctx->pixcap)); CKC(dalloc(&b.csize, n * 3 * ctx->pixcap)); CKC(dalloc(&b.coff, n * 3 * ctx->pixcap));
    CKC(dalloc(&b.corder, n * 3 * ctx->pixcap)); CKC(dalloc(&b.cpos, n * 3 * ctx->pixcap));
    CKC(dalloc(&b.tasks, n * 3 * LSD_MAXC)); CKC(dalloc(&b.worklist, n * 3 * LSD_MAXC)); CKC(dalloc(&b.taskctr, 64 * 8));
    CKC(dalloc(&b.candrank, n * 3 * ctx->segcap));
    CKC(dalloc(&b.reg, n * 3 * ctx->pixcap * 2));
    CKC(dalloc(&b.pixcount, n * 3));
    CKC(dalloc(&b.g2max, n * 3));
    CKC(dalloc(&b.rawseg, n * 3 * ctx->segcap));
    CKC(dalloc(&b.cand, n * 3 * ctx->segcap));
    CKC(dalloc(&b.candcount, n * 3));
    CKC(dalloc(&b.candlist, n * 3 * ctx->segcap));
    CKC(dalloc(&b.candseg, n * 3 * ctx->segcap));
    CKC(dalloc(&b.candok, n * 3 * ctx->segcap));
    CKC(dalloc(&b.segcount, n * 3));
    CKC(dalloc(&b.frame_off, (n + 1) + n * 3));
    b.imgoff = b.frame_off + (n + 1);
    CKC(dalloc(&b.flags, 8));
    b.outcap = (int)std::min<size_t>(n * 3 * ctx->segcap, (size_t)1 << 30);
    const size_t oc = b.outcap;
    CKC(dalloc(&b.o_color, oc)); CKC(dalloc(&b.o_lines, oc * 4)); CKC(dalloc(&b.o_normals, oc * 2));
    CKC(dalloc(&b.o_centers, oc * 2)); CKC(dalloc(&b.o_pixn, oc * 4)); CKC(dalloc(&b.o_nf32, oc * 2));
    CKC(dalloc(&b.o_ground, oc * 4)); CKC(dalloc(&b.o_keep, oc)); CKC(dalloc(&b.o_desc, oc * 32));
    CKC(dalloc(&b.o_frame, oc));
    b.o_midx = nullptr; b.o_mdist = nullptr;
    CKC(cudaMallocHost((void **)&ctx->h_small, (n * 3 + n + 1 + 8 + n * 3) * sizeof(int)));
    CKC(cudaMemsetAsync(b.flags, 0, 8 * sizeof(int), ctx->st));
    CKC(cudaStreamSynchronize(ctx->st));
    ctx->launches = 0;
    ctx->last_src_valid = false;
    ctx->grow_bits_hint = 0;
    *out = ctx;
    return LSF_OK;
#undef CKC
}


extern "C" void lsf_destroy(lsf_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    Buffers &b = ctx->b;
    void *ptrs[] = {ctx->stage_buf[0], b.ctab, b.planesA, b.planesB, b.gray, b.dx, b.lsdw, b.preact, b.prepatch, b.prectr, b.pix, b.pxy, b.fat, b.scs, b.usedbits, b.order, b.label, b.csize, b.coff, b.corder, b.cpos, b.tasks, b.worklist, b.taskctr, b.candrank, b.reg, b.pixcount,
                    b.g2max, b.rawseg, b.cand, b.candcount, b.candlist, b.candseg, b.candok, b.segcount, b.frame_off, b.flags, b.o_color, b.o_lines, b.o_normals, b.o_centers,
                    b.o_pixn, b.o_nf32, b.o_ground, b.o_keep, b.o_desc, b.o_frame, b.o_midx, b.o_mdist, ctx->map,
                    ctx->knn_scratch, ctx->tap_tmp, ctx->seg_in, ctx->carry};
    for (void *p : ptrs) if (p) cudaFree(p);
    if (ctx->stage_buf[1]) cudaFree(ctx->stage_buf[1]);
    if (ctx->kept_rec) cudaFree(ctx->kept_rec);
    exchange_destroy(ctx);
    lane_filter_destroy(ctx);
    jpeg_destroy(ctx);
    hough_destroy(ctx);
    for (void *p : {(void *)ctx->map_ground, (void *)ctx->map_color, (void *)ctx->map_frame, (void *)ctx->pose_dev}) if (p) cudaFree(p);
    for (int i = 0; i < 2; ++i) if (ctx->staged[i].ev) cudaEventDestroy(ctx->staged[i].ev);
    if (ctx->h_small) cudaFreeHost(ctx->h_small);
    for (auto &e : ctx->events) cudaEventDestroy(e.ev);
    for (auto &e : ctx->ev_copy) cudaEventDestroy(e);
    for (auto &e : ctx->ev_done) cudaEventDestroy(e);
    for (auto &e : ctx->ev_off) cudaEventDestroy(e);
    for (auto &e : ctx->ev_lbd) cudaEventDestroy(e);
    if (ctx->ev_begin) cudaEventDestroy(ctx->ev_begin);
    if (ctx->copy_st) cudaStreamDestroy(ctx->copy_st);
    for (int i = 0; i < 8; ++i) if (ctx->aux[i]) cudaStreamDestroy(ctx->aux[i]);
    if (ctx->st) cudaStreamDestroy(ctx->st);
    delete ctx;
}

extern "C" const char *lsf_last_error(const lsf_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

extern "C" int lsf_set_color_transform(lsf_ctx *ctx, const float scale[3], const float shift[3])
{
    if (!ctx || !scale || !shift) return LSF_E_ARG;
    // the six values change together: a batch that starts meanwhile sees either the old or the new transform, never a mix
    std::lock_guard<std::mutex> lk(ctx->ai_mu);
    for (int i = 0; i < 3; ++i) { ctx->cfg.ai_scale[i] = scale[i]; ctx->cfg.ai_shift[i] = shift[i]; }
    return LSF_OK;
}

extern "C" int lsf_set_chunk_frames(lsf_ctx *ctx, int chunk_frames)
{
    if (!ctx) return LSF_E_ARG;
    ctx->cfg.chunk_frames = chunk_frames;
    return LSF_OK;
}

extern "C" int lsf_set_tie_order(lsf_ctx *ctx, int tie_order)
{
    if (!ctx || (tie_order != LSF_TIES_REFERENCE && tie_order != LSF_TIES_INDEX)) return LSF_E_ARG;
    ctx->cfg.tie_order = tie_order;
    return LSF_OK;
}

extern "C" int lsf_capacities(const lsf_ctx *ctx, int *max_segments_per_color, int *max_pixels_per_color, int *max_output_rows)
{
    if (!ctx) return LSF_E_ARG;
    if (max_segments_per_color) *max_segments_per_color = ctx->segcap;
    if (max_pixels_per_color) *max_pixels_per_color = ctx->pixcap;
    if (max_output_rows) *max_output_rows = ctx->b.outcap;
    return LSF_OK;
}

extern "C" int lsf_image_dims(const lsf_ctx *ctx, int *h, int *w, int *lsd_h, int *lsd_w)
{
    if (!ctx) return LSF_E_ARG;
    if (h) *h = ctx->h;
    if (w) *w = ctx->w;
    if (lsd_h) *lsd_h = ctx->sh;
    if (lsd_w) *lsd_w = ctx->sw;
    return LSF_OK;
}

static bool g_debug_sync = getenv("LSF_DEBUG_SYNC") != nullptr;

// ---- TMA descriptor for the input frames: [frame][row][byte] uint8, box 224 x 36 x 1 (inner start 16-byte aligned) ---------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static void make_tma(lsf_ctx *ctx, const u8 *src, int n, int sh, int sw, size_t pitch)
{
    if (ctx->tma.valid && ctx->tma_src == src && ctx->tma_n == n && ctx->tma_h == sh && ctx->tma_w == sw && ctx->tma_pitch == pitch) return;
    ctx->tma.valid = 0;
    if (getenv("LSF_NO_TMA")) return;
    if (sw < 128 || sh < 40 || ((sw * 3) % 4) != 0 || (pitch % 16) != 0 || (((uintptr_t)src) % 16) != 0 || ((pitch * sh) % 16) != 0) return;
    static PFN_encodeTiled enc = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            enc = (PFN_encodeTiled)fn;
    }
    if (!enc) return;
    // uint32 elements: the box may be 104 words (416 bytes) wide; 8-bit elements cap the box at 256 bytes
    cuuint64_t gdim[3] = {(cuuint64_t)sw * 3 / 4, (cuuint64_t)sh, (cuuint64_t)n};
    cuuint64_t gstr[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * sh};
    cuuint32_t box[3] = {104, 36, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&ctx->tma.map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, (void *)src, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (g_debug_sync) fprintf(stderr, "[lsf] cuTensorMapEncodeTiled -> %d\n", (int)r);
    if (r != CUDA_SUCCESS) return;
    ctx->tma.valid = 1;
    ctx->tma_src = src; ctx->tma_n = n; ctx->tma_h = sh; ctx->tma_w = sw; ctx->tma_pitch = pitch;
}

void mark(lsf_ctx *ctx, const char *name)
{
    if (g_debug_sync) {
        cudaError_t e = cudaStreamSynchronize(ctx->st);
        if (e == cudaSuccess) e = cudaGetLastError();
        fprintf(stderr, "[lsf] stage %-20s %s\n", name, e == cudaSuccess ? "ok" : cudaGetErrorString(e));
    }
    if (ctx->n_events == (int)ctx->events.size()) {
        StageTime s; s.name = name;
        cudaEventCreate(&s.ev);
        ctx->events.push_back(s);
    }
    ctx->events[ctx->n_events].name = name;
    cudaEventRecord(ctx->events[ctx->n_events].ev, ctx->st);
    ++ctx->n_events;
}

extern "C" int lsf_last_timings(lsf_ctx *ctx, const char **names, float *ms, int cap)
{
    if (!ctx) return 0;
    int n = 0;
    for (int i = 0; i + 1 < ctx->n_events && n < cap; ++i) {
        float t = 0;
        cudaEventElapsedTime(&t, ctx->events[i].ev, ctx->events[i + 1].ev);
        names[n] = ctx->events[i + 1].name;
        ms[n] = t;
        ++n;
    }
    return n;
}

extern "C" long long lsf_launch_count(const lsf_ctx *ctx) { return ctx ? ctx->launches : 0; }
extern "C" void *lsf_stream(lsf_ctx *ctx) { return ctx ? (void *)ctx->st : nullptr; }
extern "C" const char *lsf_version(void) { return "lsf 0.1 (sm_100a)"; }

// host -> device copy of `rows` rows of `row_bytes` bytes: one linear copy when the host rows are contiguous
// every buffer whose size follows max_pixels_per_color, reallocated for a larger capacity (the batch is over: nothing reads them)
static int regrow_pixcap(lsf_ctx *ctx, int need)
{
    Buffers &b = ctx->b;
    const size_t n = ctx->max_batch, Np = (size_t)ctx->sh * ctx->sw;
    const size_t cap = std::min(Np, (((size_t)need + (size_t)need / 4 + 1023) / 1024) * 1024);
    CK(cudaDeviceSynchronize());
    for (void *p : {(void *)b.pix, (void *)b.pxy, (void *)b.fat, (void *)b.scs, (void *)b.usedbits, (void *)b.order, (void *)b.label, (void *)b.csize,
                    (void *)b.coff, (void *)b.corder, (void *)b.cpos, (void *)b.reg})
        if (p) cudaFree(p);
    b.pix = nullptr; b.pxy = nullptr; b.fat = nullptr; b.scs = nullptr; b.usedbits = nullptr; b.order = nullptr; b.label = nullptr; b.csize = nullptr;
    b.coff = nullptr; b.corder = nullptr; b.cpos = nullptr; b.reg = nullptr;
    ctx->pixcap = (int)cap;
    CK(dalloc(&b.pix, n * 3 * cap));
    CK(dalloc(&b.pxy, n * 3 * cap));
    CK(dalloc(&b.fat, n * 3 * cap * LSD_FAT_WORDS));
    CK(dalloc(&b.scs, n * 3 * cap));
    CK(dalloc(&b.usedbits, n * 3 * ((cap + 31) / 32)));
    CK(dalloc(&b.order, n * 3 * cap));
    CK(dalloc(&b.label, n * 3 * cap)); CK(dalloc(&b.csize, n * 3 * cap)); CK(dalloc(&b.coff, n * 3 * cap));
    CK(dalloc(&b.corder, n * 3 * cap)); CK(dalloc(&b.cpos, n * 3 * cap));
    CK(dalloc(&b.reg, n * 3 * cap * 2));
    return LSF_OK;
}

static cudaError_t h2d_rows(void *dst, const void *src, size_t src_pitch, size_t row_bytes, size_t rows, cudaStream_t st)
{
    if (src_pitch == row_bytes) return cudaMemcpyAsync(dst, src, row_bytes * rows, cudaMemcpyHostToDevice, st);
    return cudaMemcpy2DAsync(dst, row_bytes, src, src_pitch, row_bytes, rows, cudaMemcpyHostToDevice, st);
}

extern "C" int lsf_prefetch_batch(lsf_ctx *ctx, const uint8_t *bgr, int n, int src_h, int src_w, size_t pitch)
{
    if (!ctx) return LSF_E_ARG;
    if (!bgr || n <= 0 || n > ctx->max_batch || src_h <= 0 || src_w <= 0 || src_h > ctx->max_src_h || src_w > ctx->max_src_w ||
        pitch < (size_t)src_w * 3)
        return fail(ctx, LSF_E_ARG, "lsf_prefetch_batch: bad frames / geometry");
    ENTER(ctx);
    if (!ctx->stage_buf[1]) CK(cudaMalloc((void **)&ctx->stage_buf[1], (size_t)ctx->max_batch * ctx->max_src_h * ctx->max_src_w * 3));
    // the staging buffer that is not holding an unconsumed batch; the oldest one if both do
    int slot = !ctx->staged[0].valid ? 0 : !ctx->staged[1].valid ? 1 : (ctx->staged[0].seq < ctx->staged[1].seq ? 0 : 1);
    lsf_ctx::Staged &sg = ctx->staged[slot];
    if (ctx->last_src == ctx->stage_buf[slot]) ctx->last_src_valid = false;   // the frames of the last batch are overwritten
    CK(h2d_rows(ctx->stage_buf[slot], bgr, pitch, (size_t)src_w * 3, (size_t)src_h * n, ctx->copy_st));
    CK(cudaEventRecord(sg.ev, ctx->copy_st));
    sg.host = bgr; sg.n = n; sg.h = src_h; sg.w = src_w; sg.pitch = pitch; sg.valid = true; sg.seq = ++ctx->stage_seq;
    return LSF_OK;
}

extern "C" int lsf_cancel_prefetch(lsf_ctx *ctx)
{
    if (!ctx) return LSF_E_ARG;
    ENTER(ctx);
    CK(cudaStreamSynchronize(ctx->copy_st));
    ctx->staged[0].valid = ctx->staged[1].valid = false;
    return LSF_OK;
}

cudaMemcpyKind out_kind(int mem) { return mem == LSF_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost; }

int ensure_knn(lsf_ctx *ctx, int nq, int nm, int k)
{
    size_t need = knn_scratch_bytes(nq, nm, k);
    if (need > ctx->knn_scratch_cap) {
        if (ctx->knn_scratch) cudaFree(ctx->knn_scratch);
        ctx->knn_scratch = nullptr; ctx->knn_scratch_cap = 0;
        CK(cudaMalloc(&ctx->knn_scratch, need));
        ctx->knn_scratch_cap = need;
    }
    return LSF_OK;
}

// ---- the batch pipeline ---------------------------------------------------------------------------------------
extern "C" int lsf_front_end_batch(lsf_ctx *ctx, const uint8_t *bgr, int n, int src_h, int src_w, size_t pitch, int mem_kind,
                                   int stages, int k, lsf_segments *out)
{
    if (!ctx) return LSF_E_ARG;
    if (!bgr || !out || n <= 0) return fail(ctx, LSF_E_ARG, "lsf_front_end_batch: null frames / output or n <= 0");
    if (n > ctx->max_batch) return fail(ctx, LSF_E_CAPACITY, "lsf_front_end_batch: n exceeds max_batch " + std::to_string(ctx->max_batch));
    if (src_h <= 0 || src_w <= 0 || src_h > ctx->max_src_h || src_w > ctx->max_src_w)
        return fail(ctx, LSF_E_CAPACITY, "lsf_front_end_batch: frame larger than max_src_h x max_src_w");
    if (pitch < (size_t)src_w * 3) return fail(ctx, LSF_E_ARG, "lsf_front_end_batch: pitch smaller than a row");
    if (!(stages & LSF_STAGE_DETECT)) return fail(ctx, LSF_E_ARG, "lsf_front_end_batch: stages must include LSF_STAGE_DETECT");
    const int any_match = stages & (LSF_STAGE_MATCH | LSF_STAGE_MATCH_PREV);
    if (any_match && !(stages & LSF_STAGE_DESCRIBE))
        return fail(ctx, LSF_E_ARG, "lsf_front_end_batch: matching needs LSF_STAGE_DESCRIBE");
    if ((stages & LSF_STAGE_MATCH) && (stages & LSF_STAGE_MATCH_PREV))
        return fail(ctx, LSF_E_ARG, "lsf_front_end_batch: choose one of LSF_STAGE_MATCH / LSF_STAGE_MATCH_PREV");
    if (any_match && (k < 1 || k > 8)) return fail(ctx, LSF_E_ARG, "lsf_front_end_batch: k must be 1..8");
    ENTER(ctx);
    Buffers &b = ctx->b;
    Dims d;
    d.n = n; d.src_h = src_h; d.src_w = src_w;
    d.dh = ctx->cfg.img_h; d.dw = ctx->cfg.img_w; d.top = ctx->cfg.top_cutoff;
    d.h = ctx->h; d.w = ctx->w; d.wp = ctx->wp; d.sh = ctx->sh; d.sw = ctx->sw; d.swp = ctx->swp;
    d.pixcap = ctx->pixcap; d.segcap = ctx->segcap;
    d.logNT = 5.0 * (log10((double)d.sw) + log10((double)d.sh)) / 2.0 + log10(11.0);
    d.min_reg = (int)(-d.logNT / log10(22.5 / 180.0));
    d.identity_geom = (d.dh == src_h && d.dw == src_w);
    d.debug = getenv("LSF_TRACE_LSD") ? atoi(getenv("LSF_TRACE_LSD")) : 0;
    {   // snapshot of the colour transform for this batch (lsf_set_color_transform may be called from another thread)
        std::lock_guard<std::mutex> lk(ctx->ai_mu);
        for (int i = 0; i < 3; ++i) { ctx->cp.ai_scale[i] = ctx->cfg.ai_scale[i]; ctx->cp.ai_shift[i] = ctx->cfg.ai_shift[i]; }
    }
    d.identity_color = 1;
    for (int i = 0; i < 3; ++i) if (ctx->cp.ai_scale[i] != 1.f || ctx->cp.ai_shift[i] != 0.f) d.identity_color = 0;

    d.f0 = 0; d.grow_per_sm = 0;
    d.grow_used_bits = ctx->grow_bits_hint;
    if (!ctx->events_keep) { ctx->n_events = 0; mark(ctx, "start"); }
    ctx->events_keep = false;
    int chunk = ctx->cfg.chunk_frames;
    if (getenv("LSF_CHUNK_FRAMES")) chunk = atoi(getenv("LSF_CHUNK_FRAMES"));
    const bool host_in = mem_kind != LSF_MEM_DEVICE;
    // automatic: host frames -> 8 chunks (the copy of chunk c+1 hides behind the kernels of chunk c);
    // device frames -> 2 chunks (only the latency-bound LSD search of one half overlaps the dense kernels of the other)
    if (chunk == 0) chunk = n >= 64 ? (host_in ? std::max(16, (n + 7) / 8) : (n + 1) / 2) : n;
    const bool auto_chunk = ctx->cfg.chunk_frames == 0 && !getenv("LSF_CHUNK_FRAMES");
    const u8 *src;
    bool staged_hit = false;
    if (!host_in) {
        src = bgr; d.src_pitch = pitch; d.src_frame = pitch * src_h;
    } else {
        d.src_pitch = (size_t)src_w * 3; d.src_frame = d.src_pitch * src_h;
        // frames staged ahead by lsf_prefetch_batch (oldest matching one)?
        int hit = -1;
        for (int i = 0; i < 2; ++i) {
            const lsf_ctx::Staged &sg = ctx->staged[i];
            if (sg.valid && sg.host == bgr && sg.n == n && sg.h == src_h && sg.w == src_w && sg.pitch == pitch &&
                (hit < 0 || sg.seq < ctx->staged[hit].seq))
                hit = i;
        }
        if (hit >= 0) {
            staged_hit = true;
            if (auto_chunk && n >= 64) chunk = (n + 1) / 2;     // the frames are already on the device
            b.src = ctx->stage_buf[hit];
            ctx->staged[hit].valid = false;
            CK(cudaStreamWaitEvent(ctx->st, ctx->staged[hit].ev, 0));
        } else {
            // A host batch that was not staged: whatever is still staged belongs to a replay that was abandoned (a staged
            // entry is matched by pointer + geometry only, so it must not survive to meet a reused buffer with new
            // contents).  Drop it, then copy now.
            if (ctx->staged[0].valid || ctx->staged[1].valid) {
                CK(cudaStreamSynchronize(ctx->copy_st));
                ctx->staged[0].valid = ctx->staged[1].valid = false;
            }
            b.src = ctx->stage_buf[0];
        }
        src = b.src;
    }
    if (chunk < 0 || chunk > n || (d.debug & 2)) chunk = n;
    if ((n + chunk - 1) / chunk > 64) chunk = (n + 63) / 64;
    const int nchunks = (n + chunk - 1) / chunk;
    ctx->last_src = src; ctx->last_src_valid = true; ctx->d = d; ctx->have_batch = true;
    if (d.identity_geom) make_tma(ctx, src, n, src_h, src_w, d.src_pitch); else ctx->tma.valid = 0;
    CK(cudaMemsetAsync(b.flags, 0, 8 * sizeof(int), ctx->st));
    CK(cudaMemsetAsync(b.taskctr, 0, 64 * 8 * sizeof(int), ctx->st));
    CK(cudaMemsetAsync(b.prectr, 0, 64 * sizeof(int), ctx->st));
    const bool describe = (stages & LSF_STAGE_DESCRIBE) != 0, match_prev = (stages & LSF_STAGE_MATCH_PREV) != 0;
    if (match_prev) {
        if (!b.o_midx) { CK(dalloc(&b.o_midx, (size_t)b.outcap * 8)); CK(dalloc(&b.o_mdist, (size_t)b.outcap * 8)); }
        if (!ctx->carry) { ctx->carry_cap = 3 * ctx->segcap; CK(cudaMalloc((void **)&ctx->carry, (size_t)ctx->carry_cap * 32)); }
    }
    const size_t ps = (size_t)d.h * d.wp, N = (size_t)d.h * d.w;
    const bool piped = nchunks > 1;
    const int grow_piped = getenv("LSF_GROW_PIPED") ? atoi(getenv("LSF_GROW_PIPED")) : 10;   // measured: 16 -> 10.56 ms, 10 -> 10.10, 8 -> 10.16, 6 -> 10.29
    if (piped) {
        // Chunk pipeline: chunk c is copied on the copy stream while earlier chunks compute on the aux streams
        // (round robin).  The region-growing kernel is a chain of dependent steps that leaves issue slots free,
        // so it overlaps with the dense kernels of other chunks.  Two scalars chain the chunks: the output row
        // where a chunk starts (seg_offsets) and the previous frame's descriptors (frame-to-frame matching).
        while ((int)ctx->ev_copy.size() < nchunks) {
            cudaEvent_t e[4];
            for (int i = 0; i < 4; ++i) CK(cudaEventCreateWithFlags(&e[i], cudaEventDisableTiming));
            ctx->ev_copy.push_back(e[0]); ctx->ev_done.push_back(e[1]); ctx->ev_off.push_back(e[2]); ctx->ev_lbd.push_back(e[3]);
        }
        CK(cudaEventRecord(ctx->ev_begin, ctx->st));     // after the resets (and everything of the previous call)
        CK(cudaStreamWaitEvent(ctx->copy_st, ctx->ev_begin, 0));
        for (int i = 0; i < 8; ++i) CK(cudaStreamWaitEvent(ctx->aux[i], ctx->ev_begin, 0));
    }
    for (int c = 0; c < nchunks; ++c) {
        const int f0 = c * chunk, nc = std::min(chunk, n - f0);
        cudaStream_t cs = piped ? ctx->aux[c & 7] : ctx->st;
#define MARK(name) do { if (!piped) mark(ctx, name); } while (0)
        if (host_in && !staged_hit) {
            cudaStream_t hs = piped ? ctx->copy_st : ctx->st;
            CK(h2d_rows(b.src + (size_t)f0 * d.src_frame, bgr + (size_t)f0 * pitch * src_h, pitch, (size_t)src_w * 3,
                        (size_t)src_h * nc, hs));
            if (piped) {
                CK(cudaEventRecord(ctx->ev_copy[c], hs));
                CK(cudaStreamWaitEvent(cs, ctx->ev_copy[c], 0));
            }
            MARK("h2d");
        }
        Dims dc = d;
        dc.n = nc; dc.f0 = f0;
        dc.grow_per_sm = piped ? grow_piped : ctx->cfg.grow_warps_per_sm;
        if (getenv("LSF_GROW_PER_SM")) dc.grow_per_sm = atoi(getenv("LSF_GROW_PER_SM"));
        Buffers bc = b;
        const size_t i0 = (size_t)3 * f0;
        bc.planesA += (size_t)f0 * PA_COUNT * ps; bc.planesB += (size_t)f0 * PB_COUNT * ps;
        bc.gray += (size_t)f0 * N;
        bc.preact += i0 * (size_t)((d.sh + 7) / 8) * d.swp; bc.prectr += c;
        bc.prepatch += i0 * (size_t)((d.sh + 7) / 8) * d.swp * 81;
        bc.lsdw += i0 * d.sh * d.swp; bc.pix += i0 * d.pixcap; bc.pxy += i0 * d.pixcap; bc.scs += i0 * d.pixcap;
        bc.fat += i0 * d.pixcap * LSD_FAT_WORDS; bc.order += i0 * d.pixcap; bc.reg += i0 * 2 * d.pixcap;
        bc.usedbits += i0 * (size_t)((d.pixcap + 31) / 32);
        bc.pixcount += i0; bc.g2max += i0; bc.cand += i0 * d.segcap; bc.candcount += i0;
        bc.label += i0 * d.pixcap; bc.csize += i0 * d.pixcap; bc.coff += i0 * d.pixcap; bc.corder += i0 * d.pixcap;
        bc.cpos += i0 * d.pixcap; bc.tasks += i0 * LSD_MAXC; bc.worklist += i0 * LSD_MAXC; bc.taskctr += 8 * c;
        bc.candrank += i0 * d.segcap; bc.candlist += i0 * d.segcap; bc.candseg += i0 * d.segcap; bc.candok += i0 * d.segcap;
        bc.rawseg += i0 * d.segcap; bc.segcount += i0; bc.imgoff += i0; bc.frame_off += f0;
        launch_color_canny(dc, ctx->cp, src + (size_t)f0 * d.src_frame, ctx->tma, b.ctab, bc.planesA, bc.gray, cs);
        MARK("color_canny");
        launch_hysteresis(dc, ctx->cfg.dilation_kernel_size, bc.planesA, bc.planesB, cs);
        MARK("hysteresis_dilate");
        launch_lsd_pre(dc, bc.planesB, bc, cs);
        MARK("lsd_pre");
        if (!piped) {      // single stream: time the indexing kernel and the growing kernel separately
            if (ctx->n_events == (int)ctx->events.size()) { StageTime s; s.name = "lsd_index"; cudaEventCreate(&s.ev); ctx->events.push_back(s); }
            ctx->events[ctx->n_events].name = "lsd_index";
            launch_lsd_core(dc, bc, cs, ctx->events[ctx->n_events].ev);
            ++ctx->n_events;
        } else {
            launch_lsd_core(dc, bc, cs);
        }
        MARK("lsd_grow");
        if (describe) {
            launch_gray_sobel(dc, bc.gray, b.dx + (size_t)f0 * N * 2, nullptr, cs);
            MARK("gray_sobel");
        }
        launch_lsd_validate(dc, bc, cs);
        MARK("lsd_validate");
        if (piped && c > 0) CK(cudaStreamWaitEvent(cs, ctx->ev_off[c - 1], 0));   // this chunk's first output row
        launch_seg_offsets(dc, bc, cs);
        if (piped) CK(cudaEventRecord(ctx->ev_off[c], cs));
        launch_segments(dc, ctx->cam, bc, (stages & LSF_STAGE_GROUND) ? 1 : 0, cs);
        MARK("segments");
        if (describe) {
            launch_lbd(d, b.o_lines, b.o_frame, b.outcap, b.frame_off + f0, b.frame_off + f0 + nc, b.dx, nullptr, b.o_desc, bc.taskctr + 5, cs);
            if (piped) CK(cudaEventRecord(ctx->ev_lbd[c], cs));
            MARK("lbd");
        }
        if (match_prev) {
            if (piped && c > 0) CK(cudaStreamWaitEvent(cs, ctx->ev_lbd[c - 1], 0));   // descriptors of frame f0 - 1
            launch_knn_prev(b.o_desc, b.frame_off, f0, nc, k, LSF_MATCH_RADIUS, ctx->cfg.tie_order, ctx->carry, ctx->carry_n, b.o_midx, b.o_mdist, cs);
            MARK("knn_prev");
        }
        if (piped) {
            CK(cudaEventRecord(ctx->ev_done[c], cs));
            CK(cudaStreamWaitEvent(ctx->st, ctx->ev_done[c], 0));
        }
#undef MARK
    }
    if (piped) mark(ctx, "chunks(h2d .. knn_prev)");
    // small results first: counts, offsets, flags
    int *hs = ctx->h_small;
    CK(cudaMemcpyAsync(hs, b.segcount, (size_t)n * 3 * sizeof(int), cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaMemcpyAsync(hs + n * 3, b.frame_off, (size_t)(n + 1) * sizeof(int), cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaMemcpyAsync(hs + n * 3 + n + 1, b.flags, 8 * sizeof(int), cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaMemcpyAsync(hs + n * 3 + n + 1 + 8, b.pixcount, (size_t)n * 3 * sizeof(int), cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    CK(cudaGetLastError());
    const int *flags = hs + n * 3 + n + 1;
    {   // size the USED bitmap of the next batch's growing warps from what this batch held (+50 %, at least 4096 pixels)
        int maxn = 0;
        for (int i = 0; i < n * 3; ++i) maxn = std::max(maxn, hs[n * 3 + n + 1 + 8 + i]);
        const int had = ctx->grow_bits_hint;
        ctx->grow_bits_hint = std::min(ctx->pixcap, std::max(4096, ((maxn + maxn / 2 + 1023) / 1024) * 1024));
        if (flags[5] && had > 0) {
            // an image outgrew the bitmap sized from the previous batch (scene change): redo this batch with the new size
            ctx->grow_bits_hint = std::min(ctx->pixcap, std::max(ctx->grow_bits_hint, ((flags[5] + flags[5] / 2 + 1023) / 1024) * 1024));
            return lsf_front_end_batch(ctx, bgr, n, src_h, src_w, pitch, mem_kind, stages, k, out);
        }
    }
    if (flags[0] && ctx->pixcap_auto && (size_t)ctx->pixcap < (size_t)ctx->sh * ctx->sw) {
        // the default capacity (an eighth of the scaled image for large batches) was a guess: size it from this batch's count
        // (+25 %) and run the batch again.  A capacity the caller set is a contract and stays an error.
        int rc = regrow_pixcap(ctx, flags[0]);
        if (rc) return rc;
        return lsf_front_end_batch(ctx, bgr, n, src_h, src_w, pitch, mem_kind, stages, k, out);
    }
    if (flags[0]) return fail(ctx, LSF_E_CAPACITY, "LSD support pixels of one colour image = " + std::to_string(flags[0]) +
                                                       " exceed max_pixels_per_color = " + std::to_string(ctx->pixcap));
    if (flags[1]) return fail(ctx, LSF_E_CAPACITY, "segments of one colour image = " + std::to_string(flags[1]) +
                                                       " exceed max_segments_per_color = " + std::to_string(ctx->segcap));
    if (flags[2]) return fail(ctx, LSF_E_CAPACITY, "segments of the batch = " + std::to_string(flags[2]) +
                                                       " exceed the output row capacity " + std::to_string(b.outcap));
    const int S = hs[n * 3 + n];
    out->n_frames = n; out->n_segments = S;
    ctx->last_S = S; ctx->last_stages = stages;
    if (S > out->capacity) return fail(ctx, LSF_E_CAPACITY, "lsf_segments.capacity too small: need " + std::to_string(S));

    // matching against the ctx map (needs S on the host only for the grid size; queries are device resident)
    if ((stages & LSF_STAGE_MATCH) && S > 0) {
        if (ctx->map_n <= 0) return fail(ctx, LSF_E_ARG, "LSF_STAGE_MATCH: the map is empty (lsf_map_add first)");
        int rc = ensure_knn(ctx, S, ctx->map_n, k);
        if (rc) return rc;
        if (!b.o_midx) { CK(dalloc(&b.o_midx, (size_t)b.outcap * 8)); CK(dalloc(&b.o_mdist, (size_t)b.outcap * 8)); }
        launch_knn(b.o_desc, S, nullptr, ctx->map, ctx->map_n, k, LSF_MATCH_RADIUS, ctx->cfg.tie_order, b.o_midx, b.o_mdist, ctx->knn_scratch, ctx->knn_scratch_cap, ctx->st);
        mark(ctx, "knn");
    }
    if (stages & LSF_STAGE_MATCH_PREV) {
        // keep the last frame's descriptors for the next batch
        int l0 = hs[n * 3 + n - 1], l1 = hs[n * 3 + n];
        ctx->carry_n = l1 - l0;
        if (ctx->carry_n > 0) CK(cudaMemcpyAsync(ctx->carry, b.o_desc + (size_t)l0 * 32, (size_t)ctx->carry_n * 32, cudaMemcpyDeviceToDevice, ctx->st));
    }
    const cudaMemcpyKind kind = out_kind(out->mem);
    if (out->counts) {
        if (out->mem == LSF_MEM_DEVICE) CK(cudaMemcpyAsync(out->counts, b.segcount, (size_t)n * 3 * sizeof(int), kind, ctx->st));
        else memcpy(out->counts, hs, (size_t)n * 3 * sizeof(int));
    }
    if (out->frame_offset) {
        if (out->mem == LSF_MEM_DEVICE) CK(cudaMemcpyAsync(out->frame_offset, b.frame_off, (size_t)(n + 1) * sizeof(int), kind, ctx->st));
        else memcpy(out->frame_offset, hs + n * 3, (size_t)(n + 1) * sizeof(int));
    }
    if (S > 0) {
        const size_t s = S;
#define COPY(dst, srcp, bytes) do { if (dst) CK(cudaMemcpyAsync(dst, srcp, (bytes), kind, ctx->st)); } while (0)
        COPY(out->color, b.o_color, s);
        COPY(out->lines_px, b.o_lines, s * 16);
        COPY(out->normals, b.o_normals, s * 16);
        COPY(out->centers, b.o_centers, s * 8);
        COPY(out->pixels_normalized, b.o_pixn, s * 16);
        COPY(out->normal_f32, b.o_nf32, s * 8);
        if (stages & LSF_STAGE_GROUND) { COPY(out->ground, b.o_ground, s * 32); COPY(out->keep, b.o_keep, s); }
        if (stages & LSF_STAGE_DESCRIBE) COPY(out->desc, b.o_desc, s * 32);
        if (any_match) { COPY(out->match_idx, b.o_midx, s * k * 4); COPY(out->match_dist, b.o_mdist, s * k * 4); }
#undef COPY
    }
    mark(ctx, "d2h");
    CK(cudaStreamSynchronize(ctx->st));
    CK(cudaGetLastError());
    return LSF_OK;
}

extern "C" int lsf_detect_batch(lsf_ctx *ctx, const uint8_t *bgr, int n, int src_h, int src_w, size_t pitch, int mem_kind,
                                lsf_segments *out)
{
    return lsf_front_end_batch(ctx, bgr, n, src_h, src_w, pitch, mem_kind, LSF_STAGE_DETECT, 0, out);
}

int stage_in(lsf_ctx *ctx, size_t bytes)
{
    if (bytes > ctx->seg_in_cap) {
        if (ctx->seg_in) cudaFree(ctx->seg_in);
        ctx->seg_in = nullptr; ctx->seg_in_cap = 0;
        CK(cudaMalloc((void **)&ctx->seg_in, bytes));
        ctx->seg_in_cap = bytes;
    }
    return LSF_OK;
}

extern "C" int lsf_describe_batch(lsf_ctx *ctx, lsf_segments *segs)
{
    if (!ctx || !segs) return LSF_E_ARG;
    if (!ctx->have_batch) return fail(ctx, LSF_E_ARG, "lsf_describe_batch: no frames resident (call lsf_detect_batch first)");
    if (!segs->lines_px || !segs->frame_offset || !segs->desc) return fail(ctx, LSF_E_ARG, "lsf_describe_batch: lines_px, frame_offset and desc are required");
    if (segs->n_frames != ctx->d.n) return fail(ctx, LSF_E_ARG, "lsf_describe_batch: n_frames differs from the resident batch");
    ENTER(ctx);
    const int S = segs->n_segments, n = ctx->d.n;
    if (S <= 0) return LSF_OK;
    // layout of the staging buffer: lines f32[S][4] | frame_of_seg i32[S] | nseg i32 | desc u8[S][32]
    size_t off_frame = (size_t)S * 16, off_n = off_frame + (size_t)S * 4, off_desc = (off_n + 4 + 15) & ~(size_t)15;
    int rc = stage_in(ctx, off_desc + (size_t)S * 32);
    if (rc) return rc;
    std::vector<int> fo(n + 1);
    if (segs->mem == LSF_MEM_DEVICE) {
        CK(cudaMemcpy(fo.data(), segs->frame_offset, (n + 1) * sizeof(int), cudaMemcpyDeviceToHost));
        CK(cudaMemcpyAsync(ctx->seg_in, segs->lines_px, (size_t)S * 16, cudaMemcpyDeviceToDevice, ctx->st));
    } else {
        memcpy(fo.data(), segs->frame_offset, (n + 1) * sizeof(int));
        CK(cudaMemcpyAsync(ctx->seg_in, segs->lines_px, (size_t)S * 16, cudaMemcpyHostToDevice, ctx->st));
    }
    if (fo[n] != S) return fail(ctx, LSF_E_ARG, "lsf_describe_batch: frame_offset[n_frames] != n_segments");
    std::vector<int> fos(S + 1);
    for (int f = 0; f < n; ++f) {
        if (fo[f] > fo[f + 1] || fo[f] < 0) return fail(ctx, LSF_E_ARG, "lsf_describe_batch: frame_offset not monotone");
        for (int i = fo[f]; i < fo[f + 1]; ++i) fos[i] = f;
    }
    fos[S] = S;
    CK(cudaMemcpyAsync(ctx->seg_in + off_frame, fos.data(), (size_t)(S + 1) * 4, cudaMemcpyHostToDevice, ctx->st));
    launch_gray_sobel(ctx->d, ctx->b.gray, ctx->b.dx, nullptr, ctx->st);
    CK(cudaMemsetAsync(ctx->b.flags + 3, 0, sizeof(int), ctx->st));      // work cursor of the LBD kernel
    launch_lbd(ctx->d, (const float *)ctx->seg_in, (const int *)(ctx->seg_in + off_frame), S, nullptr, (const int *)(ctx->seg_in + off_n),
               ctx->b.dx, nullptr, ctx->seg_in + off_desc, ctx->b.flags + 3, ctx->st);
    CK(cudaMemcpyAsync(segs->desc, ctx->seg_in + off_desc, (size_t)S * 32, out_kind(segs->mem), ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    CK(cudaGetLastError());
    return LSF_OK;
}

extern "C" int lsf_project_filter_batch(lsf_ctx *ctx, const float *pixn, const uint8_t *color, int n_seg, int mem_kind,
                                        double *ground, uint8_t *keep)
{
    if (!ctx) return LSF_E_ARG;
    if (n_seg < 0 || (n_seg > 0 && (!pixn || !color || !ground || !keep))) return fail(ctx, LSF_E_ARG, "lsf_project_filter_batch: null argument");
    if (n_seg == 0) return LSF_OK;
    ENTER(ctx);
    const size_t S = n_seg;
    if (mem_kind == LSF_MEM_DEVICE) {
        launch_project_filter(ctx->cam, pixn, color, n_seg, ground, keep, ctx->st);
    } else {
        size_t off_col = S * 16, off_g = (off_col + S + 15) & ~(size_t)15, off_k = off_g + S * 32;
        int rc = stage_in(ctx, off_k + S);
        if (rc) return rc;
        CK(cudaMemcpyAsync(ctx->seg_in, pixn, S * 16, cudaMemcpyHostToDevice, ctx->st));
        CK(cudaMemcpyAsync(ctx->seg_in + off_col, color, S, cudaMemcpyHostToDevice, ctx->st));
        launch_project_filter(ctx->cam, (const float *)ctx->seg_in, ctx->seg_in + off_col, n_seg, (double *)(ctx->seg_in + off_g),
                              ctx->seg_in + off_k, ctx->st);
        CK(cudaMemcpyAsync(ground, ctx->seg_in + off_g, S * 32, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaMemcpyAsync(keep, ctx->seg_in + off_k, S, cudaMemcpyDeviceToHost, ctx->st));
    }
    CK(cudaStreamSynchronize(ctx->st));
    CK(cudaGetLastError());
    return LSF_OK;
}

extern "C" int lsf_knn_hamming(lsf_ctx *ctx, const uint8_t *query, int nq, const uint8_t *train, int nm, int k, int max_dist,
                               int mem_kind, int32_t *idx, int32_t *dist)
{
    if (!ctx) return LSF_E_ARG;
    if (nq < 0 || nm < 0 || k < 1 || k > 8) return fail(ctx, LSF_E_ARG, "lsf_knn_hamming: bad sizes (k must be 1..8)");
    if (nq == 0) return LSF_OK;
    if (!query || !idx || !dist || (nm > 0 && !train)) return fail(ctx, LSF_E_ARG, "lsf_knn_hamming: null argument");
    ENTER(ctx);
    ctx->n_events = 0;
    const size_t Q = nq, M = nm;
    const u8 *dq, *dm; int *di, *dd;
    if (mem_kind == LSF_MEM_DEVICE) {
        dq = query; dm = train; di = idx; dd = dist;
    } else {
        size_t off_m = Q * 32, off_i = off_m + M * 32, off_d = off_i + Q * k * 4;
        int rc = stage_in(ctx, off_d + Q * k * 4);
        if (rc) return rc;
        CK(cudaMemcpyAsync(ctx->seg_in, query, Q * 32, cudaMemcpyHostToDevice, ctx->st));
        if (M) CK(cudaMemcpyAsync(ctx->seg_in + off_m, train, M * 32, cudaMemcpyHostToDevice, ctx->st));
        dq = ctx->seg_in; dm = ctx->seg_in + off_m; di = (int *)(ctx->seg_in + off_i); dd = (int *)(ctx->seg_in + off_d);
    }
    int rc = ensure_knn(ctx, nq, nm > 0 ? nm : 1, k);
    if (rc) return rc;
    mark(ctx, "start");
    launch_knn(dq, nq, nullptr, dm, nm, k, max_dist, ctx->cfg.tie_order, di, dd, ctx->knn_scratch, ctx->knn_scratch_cap, ctx->st);
    mark(ctx, "knn");
    if (mem_kind != LSF_MEM_DEVICE) {
        CK(cudaMemcpyAsync(idx, di, Q * k * 4, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaMemcpyAsync(dist, dd, Q * k * 4, cudaMemcpyDeviceToHost, ctx->st));
    }
    CK(cudaStreamSynchronize(ctx->st));
    CK(cudaGetLastError());
    return LSF_OK;
}

extern "C" int lsf_pack_kept_records(lsf_ctx *ctx, int frame_base, void **records, int *n_records)
{
    if (!ctx || !records || !n_records) return LSF_E_ARG;
    if (!ctx->have_batch || !(ctx->last_stages & LSF_STAGE_GROUND))
        return fail(ctx, LSF_E_ARG, "lsf_pack_kept_records: the last batch must have run LSF_STAGE_GROUND");
    ENTER(ctx);
    if (!ctx->kept_rec) CK(cudaMalloc((void **)&ctx->kept_rec, (size_t)ctx->b.outcap * 72 + 16));
    if (!(ctx->last_stages & LSF_STAGE_DESCRIBE)) CK(cudaMemsetAsync(ctx->b.o_desc, 0, (size_t)ctx->last_S * 32, ctx->st));
    int *cnt = ctx->b.flags + 4;     // its own counter (flags[2] reports output-row overflow of the batch)
    launch_pack_kept(ctx->last_S, frame_base, ctx->b, ctx->kept_rec, ctx->b.outcap, cnt, ctx->st);
    CK(cudaMemcpyAsync(ctx->h_small, cnt, sizeof(int), cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    CK(cudaGetLastError());
    *records = ctx->kept_rec;
    *n_records = ctx->h_small[0];
    return LSF_OK;
}

extern "C" int lsf_lane_votes(lsf_ctx *ctx, double delta_d, double delta_phi, int nd, int nphi, int mem_kind, int32_t *hist)
{
    if (!ctx || !hist) return LSF_E_ARG;
    if (!ctx->have_batch || !(ctx->last_stages & LSF_STAGE_GROUND))
        return fail(ctx, LSF_E_ARG, "lsf_lane_votes: the last batch must have run LSF_STAGE_GROUND");
    if (!(delta_d > 0) || !(delta_phi > 0) || nd <= 0 || nphi <= 0 || (long long)nd * nphi > (1 << 20))
        return fail(ctx, LSF_E_ARG, "lsf_lane_votes: bad histogram geometry");
    ENTER(ctx);
    const size_t cells = (size_t)ctx->d.n * nd * nphi;
    int *dh = hist;
    if (mem_kind != LSF_MEM_DEVICE) {
        int rc = stage_in(ctx, cells * sizeof(int));
        if (rc) return rc;
        dh = (int *)ctx->seg_in;
    }
    CK(cudaMemsetAsync(dh, 0, cells * sizeof(int), ctx->st));
    launch_lane_votes(ctx->cam, ctx->last_S, delta_d, delta_phi, nd, nphi, ctx->b, dh, ctx->st);
    if (mem_kind != LSF_MEM_DEVICE) CK(cudaMemcpyAsync(hist, dh, cells * sizeof(int), cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    CK(cudaGetLastError());
    return LSF_OK;
}

extern "C" int lsf_map_clear(lsf_ctx *ctx)
{
    if (!ctx) return LSF_E_ARG;
    ctx->map_n = 0;
    return LSF_OK;
}

extern "C" int lsf_map_size(lsf_ctx *ctx) { return ctx ? ctx->map_n : LSF_E_ARG; }

extern "C" int lsf_reset_sequence(lsf_ctx *ctx)
{
    if (!ctx) return LSF_E_ARG;
    ctx->carry_n = 0;
    return LSF_OK;
}

// ---- parity taps --------------------------------------------------------------------------------------------
extern "C" int lsf_get_tap(lsf_ctx *ctx, int tap, int frame, void *dst, size_t dst_bytes)
{
    if (!ctx || !dst) return LSF_E_ARG;
    if (!ctx->have_batch) return fail(ctx, LSF_E_ARG, "lsf_get_tap: no batch processed yet");
    const Dims &d = ctx->d;
    if (frame < 0 || frame >= d.n) return fail(ctx, LSF_E_ARG, "lsf_get_tap: frame out of range");
    ENTER(ctx);
    const size_t N = (size_t)d.h * d.w, ps = (size_t)d.h * d.wp;
    size_t need = tap == LSF_TAP_IMAGE ? N * 3 : (tap == LSF_TAP_DX || tap == LSF_TAP_DY) ? N * 2 : N;
    if (dst_bytes < need) return fail(ctx, LSF_E_CAPACITY, "lsf_get_tap: need " + std::to_string(need) + " bytes");
    if (need > ctx->tap_cap) {
        if (ctx->tap_tmp) cudaFree(ctx->tap_tmp);
        ctx->tap_tmp = nullptr; ctx->tap_cap = 0;
        CK(cudaMalloc((void **)&ctx->tap_tmp, N * 4));
        ctx->tap_cap = N * 4;
    }
    const u32 *pa = ctx->b.planesA + (size_t)frame * PA_COUNT * ps;
    const u32 *pb = ctx->b.planesB + (size_t)frame * PB_COUNT * ps;
    const void *srcp = ctx->tap_tmp;
    switch (tap) {
    case LSF_TAP_IMAGE:
        if (!ctx->last_src_valid) return fail(ctx, LSF_E_ARG, "lsf_get_tap: the frames of the last batch were overwritten by a later lsf_prefetch_batch");
        launch_image_tap(d, ctx->cp, ctx->last_src + (size_t)frame * d.src_frame, ctx->tap_tmp, ctx->st); break;
    case LSF_TAP_LABELS: launch_labels_tap(pa, d.h, d.w, d.wp, ctx->tap_tmp, ctx->st); break;
    case LSF_TAP_EDGES: launch_unpack_plane(pb + PB_EDGE * ps, d.h, d.w, d.wp, ctx->tap_tmp, ctx->st); break;
    case LSF_TAP_BW_WHITE: case LSF_TAP_BW_YELLOW: case LSF_TAP_BW_RED:
        launch_unpack_plane(pb + (PB_BW0 + tap - LSF_TAP_BW_WHITE) * ps, d.h, d.w, d.wp, ctx->tap_tmp, ctx->st); break;
    case LSF_TAP_EC_WHITE: case LSF_TAP_EC_YELLOW: case LSF_TAP_EC_RED:
        launch_unpack_plane(pb + (PB_EC0 + tap - LSF_TAP_EC_WHITE) * ps, d.h, d.w, d.wp, ctx->tap_tmp, ctx->st); break;
    case LSF_TAP_GRAY: srcp = ctx->b.gray + (size_t)frame * N; break;
    case LSF_TAP_DX: case LSF_TAP_DY: {
        // de-interleave on the host
        std::vector<short> tmp(N * 2);
        CK(cudaMemcpyAsync(tmp.data(), ctx->b.dx + (size_t)frame * N * 2, N * 4, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));
        short *o = (short *)dst;
        for (size_t i = 0; i < N; ++i) o[i] = tmp[2 * i + (tap == LSF_TAP_DY ? 1 : 0)];
        return LSF_OK;
    }
    default: return fail(ctx, LSF_E_ARG, "lsf_get_tap: unknown tap");
    }
    CK(cudaMemcpyAsync(dst, srcp, need, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    CK(cudaGetLastError());
    return LSF_OK;
}
