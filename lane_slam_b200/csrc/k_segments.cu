// k_segments.cu -- K10+K12: per-segment epilogue, one thread per segment, fused in one pass:
//   normals + endpoint ordering      line_detector_lsd.py:74-125 (_findNormal, _checkBounds, _correctPixelOrdering)
//   normalisation + float32 wire     line_detector_node.py:194-205, 251-265 (Vector2D.msg float32)
//   ground projection                GroundProjection.py:38-48 (vector2pixel), :64-78 (pixel2ground), SURVEY.md A.7
//   line_sanity filter               line_sanity_node.py:48-117
// plus the scan that turns per-(frame,colour) counts into the frame-ordered output layout
// (white, yellow, red; LSD order inside a colour: line_detector_node.py:197-205).
#include "common.cuh"

namespace lsf {

// ---- exclusive scan of segcount[n*3] -> imgoff[n*3], frame_off[n+1] (single block) ----
// A pipeline chunk continues where the previous one ended: frame_off points at the chunk's first frame and,
// unless this is the first chunk, frame_off[0] already holds the previous chunk's end.
__global__ void __launch_bounds__(1024) k_seg_offsets(int nimg, int first_chunk, const int *__restrict__ segcount,
                                                     int *__restrict__ imgoff, int *__restrict__ frame_off, int outcap,
                                                     int *__restrict__ flags)
{
    __shared__ int wtot[32];
    __shared__ int carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = first_chunk ? 0 : frame_off[0];
    __syncthreads();
    for (int base = 0; base < nimg; base += 1024) {
        int i = base + tid;
        int v = i < nimg ? segcount[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wtot[warp] = incl;
        __syncthreads();
        int woff = 0;
        for (int k = 0; k < warp; ++k) woff += wtot[k];
        int excl = carry + woff + incl - v;
        if (i < nimg) {
            imgoff[i] = excl;
            if (i % 3 == 0) frame_off[i / 3] = excl;
        }
        __syncthreads();
        if (tid == 1023) carry = excl + v;
        __syncthreads();
    }
    if (tid == 0) {
        frame_off[nimg / 3] = carry;
        if (carry > outcap) atomicMax(&flags[2], carry);
    }
}

// ---- ground projection + sanity (float64, as numpy / cv2.undistortPoints) ----
__device__ __forceinline__ void undistort_point(const CamParams &cam, double u, double v, double &ou, double &ov)
{
    const double fx = cam.K[0], fy = cam.K[4], cx = cam.K[2], cy = cam.K[5];
    const double k1 = cam.D[0], k2 = cam.D[1], p1 = cam.D[2], p2 = cam.D[3], k3 = cam.D[4];
    double x = (u - cx) / fx, y = (v - cy) / fy;
    const double x0 = x, y0 = y;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        double r2 = x * x + y * y;
        double icdist = 1.0 / (1 + ((k3 * r2 + k2) * r2 + k1) * r2);
        double dX = 2 * p1 * x * y + p2 * (r2 + 2 * x * x);
        double dY = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y;
        x = (x0 - dX) * icdist;
        y = (y0 - dY) * icdist;
    }
    double xx = cam.R[0] * x + cam.R[1] * y + cam.R[2], yy = cam.R[3] * x + cam.R[4] * y + cam.R[5];
    double ww = 1. / (cam.R[6] * x + cam.R[7] * y + cam.R[8]);
    x = xx * ww; y = yy * ww;
    ou = x * cam.P[0] + cam.P[2];
    ov = y * cam.P[5] + cam.P[6];
}

__device__ __forceinline__ void vec2ground(const CamParams &cam, float vx, float vy, double &gx, double &gy)
{
    double u = cam.cam_w * (double)vx, v = cam.cam_h * (double)vy;
    if (u < 0) u = 0;
    if (u > cam.cam_w - 1) u = cam.cam_w - 1;
    if (v < 0) v = 0;
    if (v > cam.cam_h - 1) v = 0;  // sic: GroundProjection.py:47
    double ru, rv;
    undistort_point(cam, u, v, ru, rv);
    const double *H = cam.Hg;
    double g0 = H[0] * ru + H[1] * rv + H[2], g1 = H[3] * ru + H[4] * rv + H[5], g2 = H[6] * ru + H[7] * rv + H[8];
    gx = g0 / g2; gy = g1 / g2;
}

// (d_i, phi_i) of a ground segment: fancyFilters (line_sanity_node.py:75-117) == LaneFilterHistogram.generateVote
// (src/lane_filter/include/lane_filter/lane_filter.py:123-155).  Returns false for colours other than white / yellow.
__device__ __forceinline__ bool lane_vote(const CamParams &cam, double p1x, double p1y, double p2x, double p2y, int colour,
                                          double &d_i, double &phi_i)
{
    if (colour != LSF_WHITE && colour != LSF_YELLOW) return false;
    double dx = p2x - p1x, dy = p2y - p1y, nrm = sqrt(dx * dx + dy * dy);
    double tx = dx / nrm, ty = dy / nrm, nx = -ty, ny = tx;
    double d1 = nx * p1x + ny * p1y, d2 = nx * p2x + ny * p2y;
    d_i = (d1 + d2) / 2; phi_i = asin(ty);
    if (colour == LSF_WHITE) {
        if (p1x > p2x) d_i = d_i - cam.lw_white;
        else { d_i = -d_i; phi_i = -phi_i; }
        d_i = d_i - cam.lanewidth / 2;
    } else {
        if (p2x > p1x) { d_i = d_i - cam.lw_yellow; phi_i = -phi_i; }
        else d_i = -d_i;
        d_i = cam.lanewidth / 2 - d_i;
    }
    return true;
}

__device__ __forceinline__ u8 sanity_keep(const CamParams &cam, double p1x, double p1y, double p2x, double p2y, int colour)
{
    if (p1x < 0 || p2x < 0) return 0;
    double d_i, phi_i;
    if (!lane_vote(cam, p1x, p1y, p2x, p2y, colour, d_i, phi_i)) return 0;
    if (d_i > cam.d_max || d_i < cam.d_min || phi_i < cam.phi_min || phi_i > cam.phi_max) return 0;
    return 1;
}

// ---- lane-filter measurement votes: histogram of (d_i, phi_i) of the kept segments of every frame ----
// Replaces the vote loop of LaneFilterHistogram.generate_measurement_likelihood (lane_filter.py:82-102): its colour /
// x >= 0 / range tests are exactly the line_sanity keep mask.  One thread per segment, atomic counts.
__global__ void k_lane_votes(CamParams cam, int nseg, double delta_d, double delta_phi, int nd, int nphi,
                             const u8 *__restrict__ keep, const int *__restrict__ frame, const u8 *__restrict__ color,
                             const double *__restrict__ ground, int *__restrict__ hist)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nseg || !keep[i]) return;
    const double2 a = reinterpret_cast<const double2 *>(ground)[2 * (size_t)i], b = reinterpret_cast<const double2 *>(ground)[2 * (size_t)i + 1];
    double d_i, phi_i;
    if (!lane_vote(cam, a.x, a.y, b.x, b.y, color[i], d_i, phi_i)) return;
    const int bi = (int)floor((d_i - cam.d_min) / delta_d), bj = (int)floor((phi_i - cam.phi_min) / delta_phi);
    if (bi < 0 || bi >= nd || bj < 0 || bj >= nphi) return;   // a vote exactly on the upper edge (the reference raises IndexError)
    atomicAdd(&hist[((size_t)frame[i] * nd + bi) * nphi + bj], 1);
}

void launch_lane_votes(const CamParams &cam, int nseg, double delta_d, double delta_phi, int nd, int nphi, const Buffers &b, int *hist,
                       cudaStream_t st)
{
    if (nseg <= 0) return;
    k_lane_votes<<<(nseg + 127) / 128, 128, 0, st>>>(cam, nseg, delta_d, delta_phi, nd, nphi, b.o_keep, b.o_frame, b.o_color, b.o_ground, hist);
    ++g_launches;
}

__global__ void __launch_bounds__(128) k_segments(Dims d, CamParams cam, int do_ground, const LsdSeg *__restrict__ rawseg,
                                                 const int *__restrict__ segcount, const int *__restrict__ imgoff,
                                                 const u32 *__restrict__ planesB, int outcap, u8 *__restrict__ o_color,
                                                 float *__restrict__ o_lines, double *__restrict__ o_normals,
                                                 float *__restrict__ o_centers, float *__restrict__ o_pixn,
                                                 float *__restrict__ o_nf32, double *__restrict__ o_ground,
                                                 u8 *__restrict__ o_keep, int *__restrict__ o_frame)
{
    const int img = blockIdx.y, f = img / 3, c = img - f * 3;
    const int cnt = segcount[img];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cnt) return;
    const int o = imgoff[img] + i;
    if (o >= outcap) return;
    LsdSeg s = rawseg[(size_t)img * d.segcap + i];
    float x1 = s.x1, y1 = s.y1, x2 = s.x2, y2 = s.y2;
    // float32 numpy arithmetic (no FMA): length, unit normal candidates, centre
    float ddx = x1 - x2, ddy = y1 - y2;
    float len = sqrtf(__fadd_rn(__fmul_rn(ddx, ddx), __fmul_rn(ddy, ddy)));
    float dx = __fdiv_rn(y2 - y1, len), dy = __fdiv_rn(x1 - x2, len);
    float cx = __fdiv_rn(x1 + x2, 2.f), cy = __fdiv_rn(y1 + y2, 2.f);
    int x3 = clampi((int)__fsub_rn(cx, __fmul_rn(3.f, dx)), 0, d.w - 1), y3 = clampi((int)__fsub_rn(cy, __fmul_rn(3.f, dy)), 0, d.h - 1);
    int x4 = clampi((int)__fadd_rn(cx, __fmul_rn(3.f, dx)), 0, d.w - 1), y4 = clampi((int)__fadd_rn(cy, __fmul_rn(3.f, dy)), 0, d.h - 1);
    const u32 *bw = planesB + ((size_t)f * PB_COUNT + PB_BW0 + c) * (size_t)d.h * d.wp;
    int b3 = (bw[(size_t)y3 * d.wp + (x3 >> 5)] >> (x3 & 31)) & 1, b4 = (bw[(size_t)y4 * d.wp + (x4 >> 5)] >> (x4 & 31)) & 1;
    int sign = (b3 && !b4) ? 1 : -1;
    double nx = (double)dx * (double)sign, ny = (double)dy * (double)sign;
    double flag = __dsub_rn(__dmul_rn((double)(x2 - x1), ny), __dmul_rn((double)(y2 - y1), nx));
    if (flag > 0) { float t = x1; x1 = x2; x2 = t; t = y1; y1 = y2; y2 = t; }
    o_color[o] = (u8)c;
    o_frame[o] = f + d.f0;
    reinterpret_cast<float4 *>(o_lines)[o] = make_float4(x1, y1, x2, y2);
    reinterpret_cast<double2 *>(o_normals)[o] = make_double2(nx, ny);
    reinterpret_cast<float2 *>(o_centers)[o] = make_float2(cx, cy);
    reinterpret_cast<float2 *>(o_nf32)[o] = make_float2((float)nx, (float)ny);
    // (lines + [0, cut, 0, cut]) * [1/W, 1/H, 1/W, 1/H] in float64, stored as float32 (Vector2D)
    double rw = 1.0 / (double)d.dw, rh = 1.0 / (double)d.dh, cut = (double)d.top;
    float p0 = (float)__dmul_rn((double)x1, rw), p1 = (float)__dmul_rn(__dadd_rn((double)y1, cut), rh);
    float p2 = (float)__dmul_rn((double)x2, rw), p3 = (float)__dmul_rn(__dadd_rn((double)y2, cut), rh);
    reinterpret_cast<float4 *>(o_pixn)[o] = make_float4(p0, p1, p2, p3);
    if (do_ground) {
        double g[4];
        vec2ground(cam, p0, p1, g[0], g[1]);
        vec2ground(cam, p2, p3, g[2], g[3]);
        reinterpret_cast<double2 *>(o_ground)[2 * o] = make_double2(g[0], g[1]);
        reinterpret_cast<double2 *>(o_ground)[2 * o + 1] = make_double2(g[2], g[3]);
        o_keep[o] = sanity_keep(cam, g[0], g[1], g[2], g[3], c);
    }
}

void launch_seg_offsets(const Dims &d, Buffers &b, cudaStream_t st)
{
    k_seg_offsets<<<1, 1024, 0, st>>>(d.n * 3, d.f0 == 0, b.segcount, b.imgoff, b.frame_off, b.outcap, b.flags);
    ++g_launches;
}

void launch_segments(const Dims &d, const CamParams &cam, Buffers &b, int do_ground, cudaStream_t st)
{
    dim3 grid((d.segcap + 127) / 128, d.n * 3);
    k_segments<<<grid, 128, 0, st>>>(d, cam, do_ground, b.rawseg, b.segcount, b.imgoff, b.planesB, b.outcap, b.o_color, b.o_lines,
                                     b.o_normals, b.o_centers, b.o_pixn, b.o_nf32, b.o_ground, b.o_keep, b.o_frame);
    ++g_launches;
}

// ---- kept segments -> 72-byte exchange records (single CTA: scan of the keep flags, then scatter) ----
__global__ void __launch_bounds__(1024) k_pack_kept(int nseg, int frame_base, const u8 *__restrict__ keep, const int *__restrict__ frame,
                                                   const u8 *__restrict__ color, const double *__restrict__ ground,
                                                   const u8 *__restrict__ desc, u8 *__restrict__ rec, int cap, int *__restrict__ count)
{
    __shared__ int wtot[32];
    __shared__ int carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nseg; base += 1024) {
        const int i = base + tid;
        const int v = (i < nseg && keep[i]) ? 1 : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wtot[warp] = incl;
        __syncthreads();
        int woff = 0;
        for (int k = 0; k < warp; ++k) woff += wtot[k];
        const int excl = carry + woff + incl - v;
        if (v && excl < cap) {      // records past the capacity are dropped; *count still reports how many there were
            u8 *r = rec + (size_t)excl * 72;
            *reinterpret_cast<int *>(r) = frame[i] + frame_base;
            *reinterpret_cast<u32 *>(r + 4) = (u32)color[i];
            const double2 *g = reinterpret_cast<const double2 *>(ground) + 2 * (size_t)i;
            const double2 g0 = g[0], g1 = g[1];
            double *rg = reinterpret_cast<double *>(r + 8);          // records are 72 bytes: 8-byte aligned only
            rg[0] = g0.x; rg[1] = g0.y; rg[2] = g1.x; rg[3] = g1.y;
            const uint4 *dq = reinterpret_cast<const uint4 *>(desc) + 2 * (size_t)i;
            *reinterpret_cast<uint2 *>(r + 40) = make_uint2(dq[0].x, dq[0].y);
            *reinterpret_cast<uint2 *>(r + 48) = make_uint2(dq[0].z, dq[0].w);
            *reinterpret_cast<uint2 *>(r + 56) = make_uint2(dq[1].x, dq[1].y);
            *reinterpret_cast<uint2 *>(r + 64) = make_uint2(dq[1].z, dq[1].w);
        }
        __syncthreads();
        if (tid == 1023) carry = excl + v;
        __syncthreads();
    }
    if (tid == 0) *count = carry;
}

void launch_pack_kept(int nseg, int frame_base, const Buffers &b, u8 *rec, int cap, int *count, cudaStream_t st)
{
    k_pack_kept<<<1, 1024, 0, st>>>(nseg, frame_base, b.o_keep, b.o_frame, b.o_color, b.o_ground, b.o_desc, rec, cap, count);
    ++g_launches;
}

// ---- exchange step, after the all-gather: [world] slots of { int count; pad to 16 B; cap x 72-byte records } ->
// one contiguous record array in rank order (= global frame order for contiguous shards) + counts[world], total ----
__global__ void __launch_bounds__(256) k_gather_compact(const u8 *__restrict__ slots, size_t slot_bytes, int world, int cap,
                                                       u8 *__restrict__ out, int *__restrict__ meta /* [world + 2]: counts, total, overflow */)
{
    __shared__ int s_off[65];
    if (threadIdx.x == 0) {
        int run = 0, over = 0;
        for (int r = 0; r < world; ++r) {
            int c = *reinterpret_cast<const int *>(slots + (size_t)r * slot_bytes);
            if (c > cap) { over = max(over, c); c = cap; }
            s_off[r] = run;
            run += c;
        }
        s_off[world] = run;
        if (blockIdx.x == 0) {
            for (int r = 0; r < world; ++r) meta[r] = s_off[r + 1] - s_off[r];
            meta[world] = run; meta[world + 1] = over;
        }
    }
    __syncthreads();
    // 72-byte records = 18 words; grid-stride over all words of all ranks
    for (int r = 0; r < world; ++r) {
        const u32 *src = reinterpret_cast<const u32 *>(slots + (size_t)r * slot_bytes + 16);
        u32 *dst = reinterpret_cast<u32 *>(out + (size_t)s_off[r] * 72);
        const size_t nw = (size_t)(s_off[r + 1] - s_off[r]) * 18;
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nw; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
    }
}

void launch_gather_compact(const u8 *slots, size_t slot_bytes, int world, int cap, u8 *out, int *meta, cudaStream_t st)
{
    k_gather_compact<<<148, 256, 0, st>>>(slots, slot_bytes, world, cap, out, meta);
    ++g_launches;
}

// ---- map accumulation (what show_map + TF do, src/show_map/src/show_map.py:28-43, src/odometry/src/odometry.py:110-120):
// every record's ground segment, expressed in the robot frame ("duck") of ITS frame, is moved to the map frame with that
// frame's pose  p_map = R(theta) p + (x, y)  and appended to the device-resident map together with colour, frame id and
// descriptor.  cos / sin of theta come from the host (pose[f] = {x, y, cos, sin}): the host's libm is the reference's. ----
__global__ void __launch_bounds__(256) k_map_append(const u8 *__restrict__ rec, int n, const double4 *__restrict__ pose, int pose_base,
                                                   int n_pose, int map_n, double *__restrict__ m_ground, u8 *__restrict__ m_color,
                                                   int *__restrict__ m_frame, u8 *__restrict__ m_desc)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u8 *r = rec + (size_t)i * 72;
    const int frame = *reinterpret_cast<const int *>(r);
    const double *g = reinterpret_cast<const double *>(r + 8);
    double x1 = g[0], y1 = g[1], x2 = g[2], y2 = g[3];
    const int pf = frame - pose_base;
    if (pose && pf >= 0 && pf < n_pose) {
        const double4 p = pose[pf];
        const double ax = __dadd_rn(__dsub_rn(__dmul_rn(p.z, x1), __dmul_rn(p.w, y1)), p.x);
        const double ay = __dadd_rn(__dadd_rn(__dmul_rn(p.w, x1), __dmul_rn(p.z, y1)), p.y);
        const double bx = __dadd_rn(__dsub_rn(__dmul_rn(p.z, x2), __dmul_rn(p.w, y2)), p.x);
        const double by = __dadd_rn(__dadd_rn(__dmul_rn(p.w, x2), __dmul_rn(p.z, y2)), p.y);
        x1 = ax; y1 = ay; x2 = bx; y2 = by;
    }
    const size_t o = (size_t)map_n + i;
    m_ground[4 * o] = x1; m_ground[4 * o + 1] = y1; m_ground[4 * o + 2] = x2; m_ground[4 * o + 3] = y2;
    m_color[o] = r[4];
    m_frame[o] = frame;
    const uint2 *d = reinterpret_cast<const uint2 *>(r + 40);
    uint2 *md = reinterpret_cast<uint2 *>(m_desc + o * 32);
    md[0] = d[0]; md[1] = d[1]; md[2] = d[2]; md[3] = d[3];
}

void launch_map_append(const u8 *rec, int n, const double *pose4, int pose_base, int n_pose, int map_n, double *m_ground, u8 *m_color,
                       int *m_frame, u8 *m_desc, cudaStream_t st)
{
    if (n <= 0) return;
    k_map_append<<<(n + 255) / 256, 256, 0, st>>>(rec, n, reinterpret_cast<const double4 *>(pose4), pose_base, n_pose, map_n, m_ground,
                                                   m_color, m_frame, m_desc);
    ++g_launches;
}

// ---- standalone ground projection + sanity over S segments (lsf_project_filter_batch) ----
__global__ void k_project_filter(CamParams cam, const float *__restrict__ pixn, const u8 *__restrict__ color, int nseg,
                                 double *__restrict__ ground, u8 *__restrict__ keep)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nseg) return;
    float4 p = reinterpret_cast<const float4 *>(pixn)[i];
    double g[4];
    vec2ground(cam, p.x, p.y, g[0], g[1]);
    vec2ground(cam, p.z, p.w, g[2], g[3]);
    ground[4 * (size_t)i] = g[0]; ground[4 * (size_t)i + 1] = g[1];
    ground[4 * (size_t)i + 2] = g[2]; ground[4 * (size_t)i + 3] = g[3];
    keep[i] = sanity_keep(cam, g[0], g[1], g[2], g[3], color[i]);
}

void launch_project_filter(const CamParams &cam, const float *pixn, const u8 *color, int nseg, double *ground, u8 *keep,
                           cudaStream_t st)
{
    if (nseg <= 0) return;
    k_project_filter<<<(nseg + 127) / 128, 128, 0, st>>>(cam, pixn, color, nseg, ground, keep);
    ++g_launches;
}

}  // namespace lsf
