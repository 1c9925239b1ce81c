// ctx.cuh -- the context behind the C ABI (include/lsf.h), shared by the lsf_*.cu translation units.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

extern std::string g_create_error;
using namespace lsf;

struct StageTime { const char *name; cudaEvent_t ev; };

struct lsf_ctx {
    lsf_config cfg;
    int device;
    cudaStream_t st;
    Buffers b;
    Dims d;            // geometry of the last batch
    ColorParams cp;
    CamParams cam;
    int max_batch, max_src_h, max_src_w;
    int h, w, wp, sh, sw, swp, pixcap, segcap;
    bool pixcap_auto;                    // max_pixels_per_color was left to the library: it grows on demand
    bool have_batch;
    const u8 *last_src;   // device pointer of the last batch's frames
    // map of descriptors
    u8 *map;              // [map_cap][32] descriptors of the accumulated map lines (what LSF_STAGE_MATCH reads)
    double *map_ground;   // [map_cap][4] segments in the map frame
    u8 *map_color;        // [map_cap]
    int *map_frame;       // [map_cap] global frame id the line was seen in (-1: added through lsf_map_add)
    int map_n, map_cap;
    void *lane_filter;    // LaneFilterState (lsf_map_exchange.cu), or NULL
    void *jpeg;           // JpegState (k_jpeg.cu), or NULL
    void *hough;          // HoughState (k_hough.cu), or NULL
    bool events_keep;     // lsf_front_end_batch_jpeg: the batch continues the timing events of the decode stage
    long long jpeg_last_bytes;   // compressed bytes copied to the device by the last lsf_front_end_batch_jpeg
    double *pose_dev; int pose_cap;   // {x, y, cos, sin} per frame, staging of lsf_map_append_records
    // exchange step (lsf_exchange_init / lsf_allgather_segments / lsf_exchange_wait)
    struct Exchange {
        void *comm; bool own_comm; int rank, world, cap;      // cap: records per rank and exchange
        size_t slot_bytes;
        u8 *send[2], *recv[2], *gathered[2]; int *meta[2];    // double buffered: the gather of step i overlaps step i+1
        int *h_meta;                                          // pinned [2][world + 2]
        cudaStream_t st; cudaEvent_t ev_packed[2], ev_done[2];
        int parity; bool pending[2];
    } ex;
    void *knn_scratch;
    size_t knn_scratch_cap;
    u8 *carry;            // descriptors of the last frame of the previous batch
    int carry_n, carry_cap;
    // pinned host staging
    int *h_small;         // [n*3 + n+1 + 4]
    u8 *tap_tmp;
    size_t tap_cap;
    u8 *seg_in;           // staging for describe/project inputs given in host memory
    size_t seg_in_cap;
    // TMA
    TmaDesc tma;
    const u8 *tma_src; int tma_n, tma_h, tma_w; size_t tma_pitch;
    u8 *kept_rec;         // packed exchange records of the last batch (lsf_pack_kept_records)
    int last_S, last_stages;
    // staged input: two staging buffers; a prefetch (lsf_prefetch_batch) fills one while the other is being processed
    struct Staged { const u8 *host; int n, h, w; size_t pitch; cudaEvent_t ev; bool valid; unsigned long long seq; };
    u8 *stage_buf[2];
    Staged staged[2];
    unsigned long long stage_seq;
    // chunk pipeline: copy stream, compute streams, per-chunk events
    cudaStream_t copy_st;
    cudaStream_t aux[8];
    std::vector<cudaEvent_t> ev_copy, ev_done, ev_off, ev_lbd;
    cudaEvent_t ev_begin;
    // timing
    std::vector<StageTime> events;
    int n_events;
    long long launches;          // kernels launched on behalf of this ctx
    std::mutex ai_mu;            // guards cfg.ai_scale / ai_shift (lsf_set_color_transform may run concurrently with a batch)
    int grow_bits_hint;          // see Dims::grow_used_bits; 0 until a batch has been counted
    bool last_src_valid;         // LSF_TAP_IMAGE: the frames of the last batch are still where last_src points
    std::string err;
};

#define CK(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) {                                                                       \
            char buf_[512];                                                                            \
            snprintf(buf_, sizeof(buf_), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            ctx->err = buf_;                                                                           \
            return LSF_E_CUDA;                                                                         \
        }                                                                                              \
    } while (0)

// entry of every call that touches the device: select the ctx's device, count launches on the ctx
#define ENTER(ctx) do { CK(cudaSetDevice((ctx)->device)); lsf::t_launches = &(ctx)->launches; } while (0)

inline int fail(lsf_ctx *ctx, int code, const std::string &msg)
{
    if (ctx) ctx->err = msg; else g_create_error = msg;
    return code;
}


// helpers defined in lsf_api.cu
int stage_in(lsf_ctx *ctx, size_t bytes);              // ctx->seg_in scratch of at least `bytes`
int ensure_knn(lsf_ctx *ctx, int nq, int nm, int k);   // ctx->knn_scratch for a (nq x nm, k) search
void mark(lsf_ctx *ctx, const char *name);              // timing event on ctx->st
void lane_filter_destroy(lsf_ctx *ctx);                // lsf_map_exchange.cu
void exchange_destroy(lsf_ctx *ctx);
void jpeg_destroy(lsf_ctx *ctx);                       // k_jpeg.cu
void hough_destroy(lsf_ctx *ctx);                      // k_hough.cu
cudaMemcpyKind out_kind(int mem);
template <typename T> inline cudaError_t dalloc(T **p, size_t count) { return cudaMalloc((void **)p, (count ? count : 1) * sizeof(T)); }
