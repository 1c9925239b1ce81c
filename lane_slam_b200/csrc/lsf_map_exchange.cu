// lsf_map_exchange.cu -- the consumer side of the path behind the C ABI (include/lsf.h):
//   * the device-resident map of accumulated lines (segments in the map frame + colour + frame id + descriptor):
//     what show_map + TF keep (src/show_map/src/show_map.py:28-43, src/odometry/src/odometry.py:110-120) and what
//     BinaryDescriptorMatcher::add/train hold for matching (binary_descriptor_matcher.cpp:55-105);
//   * the multi-GPU exchange step: ONE ncclAllGather of fixed-capacity slots on a side stream, ordered
//     after the packing kernel by an event and overlapped with the next batch; no host round trip in between;
//   * the diff-drive odometry integrator (odometry.py:66-120), host arithmetic like the reference's.
// NCCL is bound at run time (dlopen of libnccl.so.2, e.g. the one torch ships): the library itself links only cudart.
#include <dlfcn.h>

#include "ctx.cuh"

// ---- map -------------------------------------------------------------------------------------------------------------
// grow the map arrays (descriptors, segments, colour, frame id) to hold `need` lines
static int map_reserve(lsf_ctx *ctx, int need)
{
    if (need <= ctx->map_cap) return LSF_OK;
    int ncap = std::max(std::max(ctx->map_cap * 2, need), 4096);
    u8 *nd = nullptr, *nc = nullptr; double *ng = nullptr; int *nf = nullptr;
    CK(cudaMalloc((void **)&nd, (size_t)ncap * 32)); CK(cudaMalloc((void **)&ng, (size_t)ncap * 32));
    CK(cudaMalloc((void **)&nc, (size_t)ncap)); CK(cudaMalloc((void **)&nf, (size_t)ncap * 4));
    if (ctx->map_n) {
        CK(cudaMemcpyAsync(nd, ctx->map, (size_t)ctx->map_n * 32, cudaMemcpyDeviceToDevice, ctx->st));
        CK(cudaMemcpyAsync(ng, ctx->map_ground, (size_t)ctx->map_n * 32, cudaMemcpyDeviceToDevice, ctx->st));
        CK(cudaMemcpyAsync(nc, ctx->map_color, (size_t)ctx->map_n, cudaMemcpyDeviceToDevice, ctx->st));
        CK(cudaMemcpyAsync(nf, ctx->map_frame, (size_t)ctx->map_n * 4, cudaMemcpyDeviceToDevice, ctx->st));
    }
    CK(cudaStreamSynchronize(ctx->st));
    for (void *p : {(void *)ctx->map, (void *)ctx->map_ground, (void *)ctx->map_color, (void *)ctx->map_frame}) if (p) cudaFree(p);
    ctx->map = nd; ctx->map_ground = ng; ctx->map_color = nc; ctx->map_frame = nf; ctx->map_cap = ncap;
    return LSF_OK;
}

extern "C" int lsf_map_add(lsf_ctx *ctx, const uint8_t *desc, int n, int mem_kind)
{
    if (!ctx || n < 0 || (n > 0 && !desc)) return LSF_E_ARG;
    if (n == 0) return LSF_OK;
    ENTER(ctx);
    int rc = map_reserve(ctx, ctx->map_n + n);
    if (rc) return rc;
    CK(cudaMemcpyAsync(ctx->map + (size_t)ctx->map_n * 32, desc, (size_t)n * 32,
                       mem_kind == LSF_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->st));
    CK(cudaMemsetAsync(ctx->map_ground + (size_t)ctx->map_n * 4, 0, (size_t)n * 32, ctx->st));
    CK(cudaMemsetAsync(ctx->map_color + ctx->map_n, 0xff, (size_t)n, ctx->st));
    CK(cudaMemsetAsync(ctx->map_frame + ctx->map_n, 0xff, (size_t)n * 4, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    ctx->map_n += n;
    return LSF_OK;
}

extern "C" int lsf_map_append_records(lsf_ctx *ctx, const void *records, int n, int mem_kind, const double *poses, int pose_frame_base,
                                      int n_poses)
{
    if (!ctx || n < 0 || (n > 0 && !records) || n_poses < 0 || (n_poses > 0 && !poses)) return LSF_E_ARG;
    if (n == 0) return LSF_OK;
    ENTER(ctx);
    int rc = map_reserve(ctx, ctx->map_n + n);
    if (rc) return rc;
    const u8 *rec = (const u8 *)records;
    if (mem_kind != LSF_MEM_DEVICE) {
        rc = stage_in(ctx, (size_t)n * 72);
        if (rc) return rc;
        CK(cudaMemcpyAsync(ctx->seg_in, records, (size_t)n * 72, cudaMemcpyHostToDevice, ctx->st));
        rec = ctx->seg_in;
    }
    if (n_poses > 0) {
        // {x, y, cos(theta), sin(theta)}: the trigonometry is done here, with the host libm the reference itself runs on
        std::vector<double> p4((size_t)n_poses * 4);
        for (int i = 0; i < n_poses; ++i) {
            p4[4 * i] = poses[3 * i]; p4[4 * i + 1] = poses[3 * i + 1];
            p4[4 * i + 2] = cos(poses[3 * i + 2]); p4[4 * i + 3] = sin(poses[3 * i + 2]);
        }
        if (n_poses > ctx->pose_cap) {
            if (ctx->pose_dev) cudaFree(ctx->pose_dev);
            ctx->pose_dev = nullptr; ctx->pose_cap = 0;
            CK(cudaMalloc((void **)&ctx->pose_dev, (size_t)n_poses * 32));
            ctx->pose_cap = n_poses;
        }
        CK(cudaMemcpyAsync(ctx->pose_dev, p4.data(), (size_t)n_poses * 32, cudaMemcpyHostToDevice, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));       // p4 goes out of scope
    }
    launch_map_append(rec, n, n_poses > 0 ? ctx->pose_dev : nullptr, pose_frame_base, n_poses, ctx->map_n, ctx->map_ground, ctx->map_color,
                      ctx->map_frame, ctx->map, ctx->st);
    CK(cudaStreamSynchronize(ctx->st));
    CK(cudaGetLastError());
    ctx->map_n += n;
    return LSF_OK;
}

extern "C" int lsf_map_append(lsf_ctx *ctx, const double *poses, int frame_base)
{
    if (!ctx) return LSF_E_ARG;
    void *rec = nullptr; int n = 0;
    int rc = lsf_pack_kept_records(ctx, frame_base, &rec, &n);
    if (rc) return rc;
    return lsf_map_append_records(ctx, rec, n, LSF_MEM_DEVICE, poses, frame_base, poses ? ctx->d.n : 0);
}

extern "C" int lsf_map_read(lsf_ctx *ctx, int first, int count, double *ground, uint8_t *color, int32_t *frame, uint8_t *desc)
{
    if (!ctx || first < 0 || count < 0 || first + count > ctx->map_n) return LSF_E_ARG;
    if (count == 0) return LSF_OK;
    ENTER(ctx);
    if (ground) CK(cudaMemcpyAsync(ground, ctx->map_ground + (size_t)first * 4, (size_t)count * 32, cudaMemcpyDeviceToHost, ctx->st));
    if (color) CK(cudaMemcpyAsync(color, ctx->map_color + first, (size_t)count, cudaMemcpyDeviceToHost, ctx->st));
    if (frame) CK(cudaMemcpyAsync(frame, ctx->map_frame + first, (size_t)count * 4, cudaMemcpyDeviceToHost, ctx->st));
    if (desc) CK(cudaMemcpyAsync(desc, ctx->map + (size_t)first * 32, (size_t)count * 32, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    return LSF_OK;
}

// k-NN of the descriptors of the last batch against the map, after the fact (the epoch replay appends the previous
// epoch's gathered lines while this batch is being detected, then matches)
extern "C" int lsf_match_batch(lsf_ctx *ctx, int k, int mem_kind, int32_t *match_idx, int32_t *match_dist)
{
    if (!ctx || !match_idx || !match_dist) return LSF_E_ARG;
    if (k < 1 || k > 8) return fail(ctx, LSF_E_ARG, "lsf_match_batch: k must be 1..8");
    if (!ctx->have_batch || !(ctx->last_stages & LSF_STAGE_DESCRIBE))
        return fail(ctx, LSF_E_ARG, "lsf_match_batch: the last batch must have run LSF_STAGE_DESCRIBE");
    const int S = ctx->last_S;
    if (S == 0) return LSF_OK;
    ENTER(ctx);
    Buffers &b = ctx->b;
    if (!b.o_midx) { CK(dalloc(&b.o_midx, (size_t)b.outcap * 8)); CK(dalloc(&b.o_mdist, (size_t)b.outcap * 8)); }
    if (ctx->map_n <= 0) {
        CK(cudaMemsetAsync(b.o_midx, 0xff, (size_t)S * k * 4, ctx->st));
        CK(cudaMemsetAsync(b.o_mdist, 0xff, (size_t)S * k * 4, ctx->st));
    } else {
        int rc = ensure_knn(ctx, S, ctx->map_n, k);
        if (rc) return rc;
        ctx->n_events = 0;
        mark(ctx, "start");
        launch_knn(b.o_desc, S, nullptr, ctx->map, ctx->map_n, k, LSF_MATCH_RADIUS, ctx->cfg.tie_order, b.o_midx, b.o_mdist, ctx->knn_scratch,
                   ctx->knn_scratch_cap, ctx->st);
        mark(ctx, "knn");
    }
    const cudaMemcpyKind kind = out_kind(mem_kind);
    CK(cudaMemcpyAsync(match_idx, b.o_midx, (size_t)S * k * 4, kind, ctx->st));
    CK(cudaMemcpyAsync(match_dist, b.o_mdist, (size_t)S * k * 4, kind, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    CK(cudaGetLastError());
    return LSF_OK;
}

// ---- NCCL, bound at run time -------------------------------------------------------------------------------------------
namespace {
typedef struct { char internal[128]; } nccl_unique_id;            // == ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128)
typedef int (*pfn_ncclGetUniqueId)(nccl_unique_id *);
typedef int (*pfn_ncclCommInitRank)(void **comm, int nranks, nccl_unique_id id, int rank);
typedef int (*pfn_ncclCommDestroy)(void *comm);
typedef int (*pfn_ncclAllGather)(const void *send, void *recv, size_t count, int dtype, void *comm, cudaStream_t st);
typedef const char *(*pfn_ncclGetErrorString)(int);
struct Nccl {
    void *so = nullptr;
    pfn_ncclGetUniqueId GetUniqueId = nullptr;
    pfn_ncclCommInitRank CommInitRank = nullptr;
    pfn_ncclCommDestroy CommDestroy = nullptr;
    pfn_ncclAllGather AllGather = nullptr;
    pfn_ncclGetErrorString GetErrorString = nullptr;
    std::string err;
};
Nccl g_nccl;
std::mutex g_nccl_mu;
constexpr int NCCL_UINT8 = 1;     // ncclUint8

bool nccl_load()
{
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.AllGather) return true;
    const char *names[] = {getenv("LSF_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
        if (!nm || !*nm) continue;
        g_nccl.so = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.so) break;
    }
    if (!g_nccl.so) { g_nccl.err = std::string("cannot load libnccl.so.2 (set LSF_NCCL_LIB): ") + (dlerror() ? dlerror() : ""); return false; }
    g_nccl.GetUniqueId = (pfn_ncclGetUniqueId)dlsym(g_nccl.so, "ncclGetUniqueId");
    g_nccl.CommInitRank = (pfn_ncclCommInitRank)dlsym(g_nccl.so, "ncclCommInitRank");
    g_nccl.CommDestroy = (pfn_ncclCommDestroy)dlsym(g_nccl.so, "ncclCommDestroy");
    g_nccl.GetErrorString = (pfn_ncclGetErrorString)dlsym(g_nccl.so, "ncclGetErrorString");
    pfn_ncclAllGather ag = (pfn_ncclAllGather)dlsym(g_nccl.so, "ncclAllGather");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.GetErrorString || !ag) {
        g_nccl.err = "libnccl misses a required symbol";
        return false;
    }
    g_nccl.AllGather = ag;
    return true;
}

int nccl_fail(lsf_ctx *ctx, const char *what, int rc)
{
    return fail(ctx, LSF_E_NCCL, std::string(what) + " failed: " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"));
}
}  // namespace

extern "C" int lsf_nccl_unique_id(void *id128)
{
    if (!id128) return LSF_E_ARG;
    if (!nccl_load()) { g_create_error = g_nccl.err; return LSF_E_NCCL; }
    nccl_unique_id id;
    int rc = g_nccl.GetUniqueId(&id);
    if (rc) { g_create_error = std::string("ncclGetUniqueId failed: ") + g_nccl.GetErrorString(rc); return LSF_E_NCCL; }
    memcpy(id128, &id, 128);
    return LSF_OK;
}

void exchange_destroy(lsf_ctx *ctx)
{
    lsf_ctx::Exchange &x = ctx->ex;
    if (x.st) cudaStreamSynchronize(x.st);
    if (x.comm && x.own_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(x.comm);
    for (int i = 0; i < 2; ++i) {
        for (void *p : {(void *)x.send[i], (void *)x.recv[i], (void *)x.gathered[i], (void *)x.meta[i]}) if (p) cudaFree(p);
        if (x.ev_packed[i]) cudaEventDestroy(x.ev_packed[i]);
        if (x.ev_done[i]) cudaEventDestroy(x.ev_done[i]);
    }
    if (x.h_meta) cudaFreeHost(x.h_meta);
    if (x.st) cudaStreamDestroy(x.st);
    memset(&x, 0, sizeof(x));
}

extern "C" int lsf_exchange_init(lsf_ctx *ctx, const void *unique_id128, int rank, int world, int max_records)
{
    if (!ctx) return LSF_E_ARG;
    if (world < 1 || world > 64 || rank < 0 || rank >= world || (world > 1 && !unique_id128))
        return fail(ctx, LSF_E_ARG, "lsf_exchange_init: bad rank / world (1..64) or missing unique id");
    ENTER(ctx);
    exchange_destroy(ctx);
    lsf_ctx::Exchange &x = ctx->ex;
    x.rank = rank; x.world = world;
    x.cap = max_records > 0 ? max_records : 64 * ctx->max_batch;          // kept segments are ~10-30 per frame
    x.slot_bytes = (16 + (size_t)x.cap * 72 + 15) & ~(size_t)15;
    CK(cudaStreamCreateWithFlags(&x.st, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        CK(cudaMalloc((void **)&x.send[i], x.slot_bytes));
        CK(cudaMalloc((void **)&x.recv[i], x.slot_bytes * world));
        CK(cudaMalloc((void **)&x.gathered[i], (size_t)x.cap * world * 72 + 16));
        CK(cudaMalloc((void **)&x.meta[i], (world + 2) * sizeof(int)));
        CK(cudaEventCreateWithFlags(&x.ev_packed[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&x.ev_done[i], cudaEventDisableTiming));
    }
    CK(cudaMallocHost((void **)&x.h_meta, 2 * (world + 2) * sizeof(int)));
    if (world > 1) {
        if (!nccl_load()) return fail(ctx, LSF_E_NCCL, g_nccl.err);
        nccl_unique_id id;
        memcpy(&id, unique_id128, 128);
        int rc = g_nccl.CommInitRank(&x.comm, world, id, rank);
        if (rc) { x.comm = nullptr; return nccl_fail(ctx, "ncclCommInitRank", rc); }
        x.own_comm = true;
    }
    return LSF_OK;
}

// Start the exchange of the last batch's kept segments.  Packing runs on the ctx stream (it reads the batch's output
// rows, which the next batch overwrites); the all-gather and the compaction run on the exchange stream behind an event,
// so they overlap whatever the caller launches next on the ctx.
extern "C" int lsf_allgather_segments(lsf_ctx *ctx, void *comm, int frame_base)
{
    if (!ctx) return LSF_E_ARG;
    lsf_ctx::Exchange &x = ctx->ex;
    if (x.world < 1) return fail(ctx, LSF_E_ARG, "lsf_allgather_segments: call lsf_exchange_init first");
    if (!ctx->have_batch || !(ctx->last_stages & LSF_STAGE_GROUND))
        return fail(ctx, LSF_E_ARG, "lsf_allgather_segments: the last batch must have run LSF_STAGE_GROUND");
    ENTER(ctx);
    const int p = x.parity;
    if (x.pending[p]) return fail(ctx, LSF_E_ARG, "lsf_allgather_segments: two exchanges are already in flight (lsf_exchange_wait first)");
    void *use_comm = comm ? comm : x.comm;
    if (x.world > 1 && !use_comm) return fail(ctx, LSF_E_NCCL, "lsf_allgather_segments: no communicator");
    if (!(ctx->last_stages & LSF_STAGE_DESCRIBE)) CK(cudaMemsetAsync(ctx->b.o_desc, 0, (size_t)ctx->last_S * 32, ctx->st));
    // slot = { int count; pad; records }: the count travels with the records, no separate collective
    launch_pack_kept(ctx->last_S, frame_base, ctx->b, x.send[p] + 16, x.cap, reinterpret_cast<int *>(x.send[p]), ctx->st);
    CK(cudaEventRecord(x.ev_packed[p], ctx->st));
    CK(cudaStreamWaitEvent(x.st, x.ev_packed[p], 0));
    if (x.world > 1) {
        int rc = g_nccl.AllGather(x.send[p], x.recv[p], x.slot_bytes, NCCL_UINT8, use_comm, x.st);
        if (rc) return nccl_fail(ctx, "ncclAllGather", rc);
    } else {
        CK(cudaMemcpyAsync(x.recv[p], x.send[p], x.slot_bytes, cudaMemcpyDeviceToDevice, x.st));
    }
    launch_gather_compact(x.recv[p], x.slot_bytes, x.world, x.cap, x.gathered[p], x.meta[p], x.st);
    CK(cudaMemcpyAsync(x.h_meta + p * (x.world + 2), x.meta[p], (x.world + 2) * sizeof(int), cudaMemcpyDeviceToHost, x.st));
    CK(cudaEventRecord(x.ev_done[p], x.st));
    x.pending[p] = true;
    x.parity ^= 1;
    return LSF_OK;
}

// Finish the OLDEST exchange in flight: records of all ranks in rank order (device memory, valid until two more
// exchanges were started), total and per-rank counts.
extern "C" int lsf_exchange_wait(lsf_ctx *ctx, void **records, int *n_total, int *counts, int counts_cap)
{
    if (!ctx || !records || !n_total) return LSF_E_ARG;
    lsf_ctx::Exchange &x = ctx->ex;
    if (x.world < 1) return fail(ctx, LSF_E_ARG, "lsf_exchange_wait: call lsf_exchange_init first");
    ENTER(ctx);
    int p = x.parity;                       // the older of the two slots is the one the next start would reuse
    if (!x.pending[p]) p ^= 1;
    if (!x.pending[p]) return fail(ctx, LSF_E_ARG, "lsf_exchange_wait: no exchange in flight");
    CK(cudaEventSynchronize(x.ev_done[p]));
    CK(cudaStreamWaitEvent(ctx->st, x.ev_done[p], 0));    // later work on the ctx stream may read the gathered records
    x.pending[p] = false;
    const int *m = x.h_meta + p * (x.world + 2);
    if (m[x.world + 1])
        return fail(ctx, LSF_E_CAPACITY, "exchange: a rank kept " + std::to_string(m[x.world + 1]) + " segments, more than max_records = " +
                                             std::to_string(x.cap));
    *records = x.gathered[p];
    *n_total = m[x.world];
    if (counts) for (int r = 0; r < x.world && r < counts_cap; ++r) counts[r] = m[r];
    return LSF_OK;
}

// ---- odometry (host arithmetic, like the reference node) -----------------------------------------------------------------
// OdometryNode.getPose + drive (src/odometry/src/odometry.py:66-120): dt from the stamps' nanosecond fields, the step is
// applied only when 0 < dt < 0.3; equal wheel speeds -> straight along (cos, -sin); otherwise rotation about the centre of
// curvature with wheel base l = 0.5.  Same operation order as the Python/numpy code, double precision, host libm.
extern "C" int lsf_odometry_init(lsf_odometry *st, double stamp_nsecs)
{
    if (!st) return LSF_E_ARG;
    st->x = 0.0; st->y = 0.0; st->theta = 0.0;
    st->last_t = stamp_nsecs / 1e9;
    st->dt = 0.1;
    return LSF_OK;
}

extern "C" int lsf_odometry_step(lsf_odometry *st, double stamp_nsecs, double vel_left, double vel_right)
{
    if (!st) return LSF_E_ARG;
    st->dt = stamp_nsecs / 1e9 - st->last_t;
    st->last_t = stamp_nsecs / 1e9;
    if (!((st->dt > 0.0) && (st->dt < 0.3))) return 0;          // the reference skips the update (and the TF broadcast)
    const double Vl = vel_left, Vr = vel_right, l = 0.5;
    const double ang = st->theta;
    if (Vl == Vr) {
        const double k = st->dt * Vl;                            // self.dt * Vl * get_dir_vec(angle): left to right
        st->x = st->x + k * cos(ang);
        st->y = st->y + k * (-sin(ang));
        return 1;
    }
    const double w = (Vr - Vl) / l;
    const double r = (l * (Vl + Vr)) / (2 * (Vl - Vr));
    const double rot = w * st->dt;
    const double rvx = sin(ang), rvy = cos(ang);                 // get_right_vec
    const double px = st->x, py = st->y;
    const double cx = px + r * rvx, cy = py + r * rvy;
    const double dx = px - cx, dy = py - cy;                     // rotate_point
    const double ndx = dx * cos(rot) + dy * sin(rot);
    const double ndy = dy * cos(rot) - dx * sin(rot);
    st->x = cx + ndx; st->y = cy + ndy;
    st->theta = ang + rot;
    return 1;
}

// ---- histogram lane filter ------------------------------------------------------------------------------
// numpy's pairwise summation of n elements as a plan: leaves (<= 128 elements) and the order their sums are combined in
static void pairwise_plan(int off, int n, LaneFilterPlan &pl, int &result_leaf)
{
    if (n <= 128) {
        result_leaf = pl.nleaf;
        pl.leaf_off[pl.nleaf] = (short)off; pl.leaf_len[pl.nleaf] = (short)n;
        ++pl.nleaf;
        return;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    int a = 0, b = 0;
    pairwise_plan(off, n2, pl, a);
    pairwise_plan(off + n2, n - n2, pl, b);
    pl.comb_a[pl.ncomb] = (unsigned char)a; pl.comb_b[pl.ncomb] = (unsigned char)b;
    ++pl.ncomb;
    result_leaf = a;
}

struct LaneFilterState {
    LaneFilterPlan pl;
    double *d_grid, *phi_grid, *sin_phi, *w_d, *w_phi, *belief, *belief0, *dvw, *est;
    int *hist;
    int frames_cap;
};

static void lane_filter_free(lsf_ctx *ctx)
{
    LaneFilterState *s = (LaneFilterState *)ctx->lane_filter;
    if (!s) return;
    for (void *p : {(void *)s->d_grid, (void *)s->phi_grid, (void *)s->sin_phi, (void *)s->w_d, (void *)s->w_phi, (void *)s->belief, (void *)s->belief0,
                    (void *)s->dvw, (void *)s->est, (void *)s->hist})
        if (p) cudaFree(p);
    delete s;
    ctx->lane_filter = nullptr;
}

void lane_filter_destroy(lsf_ctx *ctx) { lane_filter_free(ctx); }

extern "C" int lsf_lane_filter_init(lsf_ctx *ctx, const lsf_lane_filter_config *c)
{
    if (!ctx || !c) return LSF_E_ARG;
    const long long nc = (long long)c->nd * c->nphi;
    if (c->nd < 1 || c->nphi < 1 || nc > 1024 || c->r_d < 0 || c->r_phi < 0 || c->r_d > 64 || c->r_phi > 64 || !(c->delta_d > 0) || !(c->delta_phi > 0) ||
        !c->d_grid || !c->phi_grid || !c->sin_phi || !c->w_d || !c->w_phi || !c->belief0)
        return fail(ctx, LSF_E_CONFIG, "lsf_lane_filter_init: bad histogram geometry (at most 1024 cells) or missing table");
    ENTER(ctx);
    lane_filter_free(ctx);
    LaneFilterState *s = new LaneFilterState();
    memset(s, 0, sizeof(*s));
    ctx->lane_filter = s;
    LaneFilterPlan &pl = s->pl;
    pl.nd = c->nd; pl.nphi = c->nphi; pl.r_d = c->r_d; pl.r_phi = c->r_phi;
    pl.d_min = c->d_min; pl.d_max = c->d_max; pl.phi_min = c->phi_min; pl.phi_max = c->phi_max; pl.delta_d = c->delta_d; pl.delta_phi = c->delta_phi;
    int root = 0;
    pairwise_plan(0, (int)nc, pl, root);
    auto up = [&](double **dst, const double *src, size_t n) -> cudaError_t {
        cudaError_t e = cudaMalloc((void **)dst, n * sizeof(double));
        return e != cudaSuccess ? e : cudaMemcpyAsync(*dst, src, n * sizeof(double), cudaMemcpyHostToDevice, ctx->st);
    };
    CK(up(&s->d_grid, c->d_grid, nc)); CK(up(&s->phi_grid, c->phi_grid, nc)); CK(up(&s->sin_phi, c->sin_phi, nc));
    CK(up(&s->w_d, c->w_d, c->r_d + 1)); CK(up(&s->w_phi, c->w_phi, c->r_phi + 1));
    CK(up(&s->belief0, c->belief0, nc)); CK(up(&s->belief, c->belief0, nc));
    CK(cudaStreamSynchronize(ctx->st));
    return LSF_OK;
}

extern "C" int lsf_lane_filter_reset(lsf_ctx *ctx)
{
    if (!ctx || !ctx->lane_filter) return LSF_E_ARG;
    ENTER(ctx);
    LaneFilterState *s = (LaneFilterState *)ctx->lane_filter;
    CK(cudaMemcpyAsync(s->belief, s->belief0, (size_t)s->pl.nd * s->pl.nphi * sizeof(double), cudaMemcpyDeviceToDevice, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    return LSF_OK;
}

extern "C" int lsf_lane_filter_batch(lsf_ctx *ctx, const double *dt_v_w, int use_propagation, double *estimates)
{
    if (!ctx || !estimates || (use_propagation && !dt_v_w)) return LSF_E_ARG;
    LaneFilterState *s = (LaneFilterState *)ctx->lane_filter;
    if (!s) return fail(ctx, LSF_E_ARG, "lsf_lane_filter_batch: call lsf_lane_filter_init first");
    if (!ctx->have_batch || !(ctx->last_stages & LSF_STAGE_GROUND))
        return fail(ctx, LSF_E_ARG, "lsf_lane_filter_batch: the last batch must have run LSF_STAGE_GROUND");
    ENTER(ctx);
    const int n = ctx->d.n, nc = s->pl.nd * s->pl.nphi;
    if (n > s->frames_cap) {
        for (void *p : {(void *)s->dvw, (void *)s->est, (void *)s->hist}) if (p) cudaFree(p);
        s->dvw = s->est = nullptr; s->hist = nullptr; s->frames_cap = 0;
        CK(cudaMalloc((void **)&s->dvw, (size_t)n * 3 * sizeof(double))); CK(cudaMalloc((void **)&s->est, (size_t)n * 3 * sizeof(double)));
        CK(cudaMalloc((void **)&s->hist, (size_t)n * nc * sizeof(int)));
        s->frames_cap = n;
    }
    if (use_propagation) CK(cudaMemcpyAsync(s->dvw, dt_v_w, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, ctx->st));
    CK(cudaMemsetAsync(s->hist, 0, (size_t)n * nc * sizeof(int), ctx->st));
    launch_lane_votes(ctx->cam, ctx->last_S, s->pl.delta_d, s->pl.delta_phi, s->pl.nd, s->pl.nphi, ctx->b, s->hist, ctx->st);
    launch_lane_filter(s->pl, n, use_propagation ? 1 : 0, s->dvw, s->hist, s->d_grid, s->phi_grid, s->sin_phi, s->w_d, s->w_phi, s->belief, s->est,
                       ctx->st);
    CK(cudaMemcpyAsync(estimates, s->est, (size_t)n * 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    CK(cudaGetLastError());
    return LSF_OK;
}

extern "C" int lsf_lane_filter_belief(lsf_ctx *ctx, double *belief)
{
    if (!ctx || !belief || !ctx->lane_filter) return LSF_E_ARG;
    ENTER(ctx);
    LaneFilterState *s = (LaneFilterState *)ctx->lane_filter;
    CK(cudaMemcpyAsync(belief, s->belief, (size_t)s->pl.nd * s->pl.nphi * sizeof(double), cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    return LSF_OK;
}
