// common.cuh -- shared declarations of the lsf CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <stddef.h>
#include <mutex>

#include "../../include/lsf.h"

namespace lsf {

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;

// ---- bit-plane indices -------------------------------------------------------------------------
// planes A: written by the colour+Canny kernel
enum { PA_RAW_W = 0, PA_RAW_Y = 1, PA_RAW_R = 2, PA_CAND = 3, PA_STRONG = 4, PA_COUNT = 5 };
// planes B: written by the hysteresis/dilate kernel
enum { PB_EDGE = 0, PB_BW0 = 1, PB_EC0 = 4, PB_COUNT = 7 };

// ---- geometry of one batch ---------------------------------------------------------------------
struct Dims {
    int n;                 // frames
    int src_h, src_w;      // input frame
    size_t src_pitch;      // bytes per input row
    size_t src_frame;      // bytes per input frame
    int dh, dw, top;       // resize target (img_size) and top_cutoff
    int h, w;              // processed image: h = dh - top, w = dw
    int wp;                // 32-bit words per bit-plane row
    int sh, sw, swp;       // LSD scaled image and its words per row
    int pixcap;            // LSD support-pixel capacity per colour image
    int segcap;            // segment capacity per colour image
    int identity_geom;     // 1: no resize (dh==src_h, dw==src_w)
    int identity_color;    // 1: AntiInstagram scale==1, shift==0
    int debug;             // LSF_TRACE_LSD: bit 0 device printf of every LSD candidate, bit 1 grow cycle counters
    double logNT;          // LSD: 5 (log10 sw + log10 sh) / 2 + log10 11, evaluated on the host (the reference's libm)
    int min_reg;           // LSD: int(-logNT / log10(22.5 / 180)), smallest region worth a rectangle
    int grow_per_sm;       // persistent growing warps per SM for this launch (fewer when chunks overlap: leaves registers
                           // and issue slots to the dense kernels of the other chunk)
    int grow_used_bits;    // size of the shared-memory USED bitmap of a growing warp, in support pixels (0 = pixcap): the largest
                           // per-image support-pixel count of the previous batch plus head-room; an image that exceeds it is
                           // reported through flags[5] and the batch is redone with the full size
    int f0;                // first frame of this launch inside the batch (TMA coordinate, frame ids of the output rows;
                           // every per-frame / per-image buffer is pre-offset to the chunk)
};

// one entry per 32 scaled pixels: defined-angle bits + number of defined pixels before this word
struct LsdWord { u32 bits; u32 base; };

// per LSD support pixel (raster order inside one colour image), 16 bytes.  The level-line angle in radians
// is double(deg) * pi/180 exactly as the reference computes it from the float fastAtan2 result.
struct __align__(16) LsdPix {
    float deg;      // fastAtan2 level-line angle in degrees
    float c, s;     // cosf / sinf of float(angle in radians)
    u32 g2;         // gx^2 + gy^2 (norm = sqrt(g2/4))
};
constexpr u32 LSD_NONE = 0xffffffffu;
constexpr int LSD_FAT_WORDS = 32;   // words per neighbour record of a support pixel (k_lsd_core.cu)
constexpr int LSD_MAXC = 1024;   // growing tasks per colour image (component slot s -> task s % LSD_MAXC)

// parameters handed to kernels by value
struct ColorParams {
    int lo[4][3], hi[4][3];
    int canny_lo, canny_hi;
    float ai_scale[3], ai_shift[3];
};

struct CamParams {
    double K[9], D[5], R[9], P[12], Hg[9];
    int cam_w, cam_h;
    double lanewidth, lw_white, lw_yellow, d_min, d_max, phi_min, phi_max;
};

// raw LSD output per accepted segment (before normals / ordering)
struct LsdSeg { float x1, y1, x2, y2; };

// rectangle candidate handed from the growing kernel to the NFA validation kernel
struct LsdCand { double x1, y1, x2, y2, width, theta, dx, dy; };

// ---- device buffers of a context ------------------------------------------------------------------
struct Buffers {
    u8 *src;            // staged input frames (when the caller passes host memory)
    u8 *ctab;           // HSV tables of the context (build_color_tables)
    u32 *planesA;       // [n][PA_COUNT][h][wp]
    u32 *planesB;       // [n][PB_COUNT][h][wp]
    u8 *gray;           // [n][h][w]
    short *dx, *dy;     // [n][h][w]  (descriptor path)
    LsdWord *lsdw;      // [n*3][sh][swp]
    u32 *preact;        // [n*3][ceil(sh/8)][swp] active (image, band, word) tasks of the LSD pre-pass
    u32 *prepatch;      // [n*3*ceil(sh/8)*swp][81] scaled 9 x 36 byte patch of every active task (same slot as in preact)
    int *prectr;        // [64]           per pipeline chunk: number of active tasks
    LsdPix *pix;        // [n*3][pixcap]
    u32 *pxy;           // [n*3][pixcap]  (y << 16) | x in the scaled image
    float2 *scs;        // [n*3][pixcap]  float(cos), float(sin) of the double angle (sums of a region seeded here)
    u32 *fat;           // [n*3][pixcap][LSD_FAT_WORDS] per pixel: index, angle, cos, sin of its 8 neighbours (LSD_NONE = undefined)
    u32 *order;         // [n*3][pixcap]  seed order (compact indices)
    u32 *label, *csize, *coff;   // [n*3][pixcap] connected components: root label, size (at roots), seed-list cursor (at roots)
    u32 *corder, *cpos; // [n*3][pixcap]  seed order partitioned by component; position of each entry in order[]
    uint2 *tasks;       // [n*3][LSD_MAXC] {offset into corder, size} of every task (union of components with >= min_reg pixels)
    uint2 *worklist;    // [n*3*LSD_MAXC] (image, task) work list of the growing kernel: big tasks from the front, small from the back
    int *taskctr;       // [64][8]        per pipeline chunk: big tasks, small tasks, grow cursor, candidates, validate cursor,
                        //                LBD cursor
    u32 *candrank;      // [n*3][segcap]  position of the candidate's seed in order[] (restores the acceptance order)
    uint4 *reg;         // [n*3][2*pixcap] region point list {idx, xy, g2, angle bits} + scratch
    u32 *usedbits;      // [n*3][ceil(pixcap/32)] USED bitmap, only when it does not fit in shared memory
    int *pixcount;      // [n*3]
    u32 *g2max;         // [n*3]
    LsdCand *cand;      // [n*3][segcap]  candidates after refine, in seed order
    int *candcount;     // [n*3]
    uint2 *candlist;    // [n*3*segcap]   flat (image, candidate) work list for the validation kernel
    LsdSeg *candseg;    // [n*3][segcap]  validated endpoints
    u8 *candok;         // [n*3][segcap]  1 iff NFA accepted
    LsdSeg *rawseg;     // [n*3][segcap]
    int *segcount;      // [n*3]
    int *frame_off;     // [n+1]          first output row of every frame (global: chunks chain through it)
    int *imgoff;        // [n*3]          first output row of every colour image
    int *flags;         // [8] 0 pix overflow, 1 seg/candidate overflow, 2 out capacity, 3 LBD cursor (lsf_describe_batch), 4 pack counter,
                        //     5 support pixels of an image that did not fit the USED bitmap of the growing kernel
    // compacted per-segment outputs (capacity outcap)
    int outcap;
    u8 *o_color; float *o_lines; double *o_normals; float *o_centers; float *o_pixn; float *o_nf32;
    double *o_ground; u8 *o_keep; u8 *o_desc; int *o_frame; int *o_midx; int *o_mdist;
};

// ---- kernel launchers (one per .cu file) ------------------------------------------------------------
struct TmaDesc { CUtensorMap map; int valid; };

constexpr int COLOR_TABLE_BYTES = 2048 + 768;
void build_color_tables(const ColorParams &cp, u8 *dst);
void launch_color_canny(const Dims &d, const ColorParams &cp, const u8 *src, const TmaDesc &tma, const u8 *tables, u32 *planesA,
                        u8 *gray, cudaStream_t st);
void launch_hysteresis(const Dims &d, int dilate, const u32 *planesA, u32 *planesB, cudaStream_t st);
void launch_lsd_pre(const Dims &d, const u32 *planesB, Buffers &b, cudaStream_t st);
void launch_lsd_core(const Dims &d, Buffers &b, cudaStream_t st, cudaEvent_t ev_indexed = nullptr);      // seeds + region growing + refine
void launch_lsd_validate(const Dims &d, Buffers &b, cudaStream_t st);  // NFA validation + emit
void launch_seg_offsets(const Dims &d, Buffers &b, cudaStream_t st);
void launch_segments(const Dims &d, const CamParams &cam, Buffers &b, int do_ground, cudaStream_t st);
void launch_lane_votes(const CamParams &cam, int nseg, double delta_d, double delta_phi, int nd, int nphi, const Buffers &b, int *hist,
                       cudaStream_t st);
// lane filter (k_lane_filter.cu): histogram geometry + numpy's pairwise-summation plan for nd * nphi cells (host-made)
constexpr int LF_MAX_LEAVES = 32;
struct LaneFilterPlan {
    int nd, nphi, r_d, r_phi;
    double d_min, d_max, phi_min, phi_max, delta_d, delta_phi;
    int nleaf, ncomb;
    short leaf_off[LF_MAX_LEAVES], leaf_len[LF_MAX_LEAVES];
    unsigned char comb_a[LF_MAX_LEAVES], comb_b[LF_MAX_LEAVES];
};
size_t lane_filter_smem(int ncell);
void launch_lane_filter(const LaneFilterPlan &pl, int n_frames, int use_propagation, const double *dt_v_w, const int *hist,
                        const double *d_grid, const double *phi_grid, const double *sin_phi, const double *w_d, const double *w_phi,
                        double *belief, double *est, cudaStream_t st);
void launch_pack_kept(int nseg, int frame_base, const Buffers &b, u8 *rec, int cap, int *count, cudaStream_t st);
void launch_gather_compact(const u8 *slots, size_t slot_bytes, int world, int cap, u8 *out, int *meta, cudaStream_t st);
void launch_map_append(const u8 *rec, int n, const double *pose4, int pose_base, int n_pose, int map_n, double *m_ground, u8 *m_color,
                       int *m_frame, u8 *m_desc, cudaStream_t st);
void launch_gray_sobel(const Dims &d, const u8 *gray, short *dx, short *dy, cudaStream_t st);
void launch_lbd(const Dims &d, const float *lines, const int *frame_of_seg, int nseg_cap, const int *seg_lo_dev,
                const int *seg_hi_dev, const short *dx, const short *dy, u8 *desc, int *cursor, cudaStream_t st);
void launch_project_filter(const CamParams &cam, const float *pixn, const u8 *color, int nseg, double *ground, u8 *keep,
                           cudaStream_t st);
void launch_knn(const u8 *q, int nq_cap, const int *nq_dev, const u8 *m, int nm, int k, int max_dist, int tie_order, int *idx, int *dist,
                void *scratch, size_t scratch_bytes, cudaStream_t st);
size_t knn_scratch_bytes(int nq, int nm, int k);
void launch_knn_prev(const u8 *desc, const int *frame_off, int f_begin, int n, int k, int max_dist, int tie_order, const u8 *carry,
                     int carry_n, int *idx, int *dist, cudaStream_t st);
void launch_unpack_plane(const u32 *plane, int h, int w, int wp, u8 *dst, cudaStream_t st);
void launch_labels_tap(const u32 *planesA_frame, int h, int w, int wp, u8 *dst, cudaStream_t st);
void launch_image_tap(const Dims &d, const ColorParams &cp, const u8 *src_frame, u8 *dst, cudaStream_t st);

// Kernels launched by this library: every API entry points t_launches at its ctx's counter (several contexts, on the same
// or on different devices, may be driven from different threads).
extern thread_local long long *t_launches;
#define g_launches (*::lsf::t_launches)

// One-time initialisation PER DEVICE (constant tables, >48 KB shared-memory opt-ins): __constant__ / __device__ symbols and
// function attributes belong to the device's context, so a flag per device ordinal -- not per process -- guards them.
constexpr int LSF_MAX_DEVICES = 64;
struct PerDevice {
    std::mutex mu;
    size_t level[LSF_MAX_DEVICES] = {};
    // runs f() when `want` exceeds what was set up on the current device so far (want = 1 for plain once-only tables)
    template <typename F> void ensure(size_t want, F f)
    {
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= LSF_MAX_DEVICES) dev = 0;
        std::lock_guard<std::mutex> lk(mu);
        if (level[dev] >= want) return;
        f();
        level[dev] = want;
    }
};

// ---- small device helpers -----------------------------------------------------------------------------
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// nearest-resize source index: min(floor(d * src/dst), src-1)  (cv2.resize INTER_NEAREST)
__device__ __forceinline__ int nearest_src(int d, int src, int dst)
{
    if (src == dst) return d;
    int s = (int)floor((double)d * ((double)src / (double)dst));
    return s > src - 1 ? src - 1 : s;
}

// AntiInstagram scale/shift (float32) followed by cv2.convertScaleAbs: sat_u8(rint(|v*scale + shift|))
__device__ __forceinline__ u8 color_correct(u8 v, float sc, float sf)
{
    float t = __fadd_rn(__fmul_rn((float)v, sc), sf);
    int r = __float2int_rn(fabsf(t));
    return (u8)(r > 255 ? 255 : r);
}

// OpenCV fastAtan2 (degrees), float32 polynomial without FMA contraction
__device__ __forceinline__ float fast_atan2_deg(float y, float x)
{
    const float k = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * k, p3 = -0.3258083974640975f * k;
    const float p5 = 0.1555786518463281f * k, p7 = -0.04432655554792128f * k;
    const float eps = 2.2204460492503131e-16f;
    float ax = fabsf(x), ay = fabsf(y), a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

}  // namespace lsf
