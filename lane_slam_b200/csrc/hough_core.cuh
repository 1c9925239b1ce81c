// hough_core.cuh -- cv2.HoughLinesP as LineDetectorHSV calls it (src/line_detector/include/line_detector/line_detector1.py:63-69:
// rho = 1, theta = pi/180, threshold / minLineLength / maxLineGap from the YAML), i.e. OpenCV's progressive probabilistic Hough
// transform (imgproc/hough.cpp, HoughLinesProbabilistic; OpenCV 4.13 is what the reference's Python imports here).  The source of
// OpenCV is not part of the reference repository: the published algorithm is restated and pinned against cv2.HoughLinesP itself
// (tests/test_hough_core.py, bit-exact line lists, same order).
//
// The algorithm is sequential by construction -- edge points are drawn in the order of OpenCV's RNG, every draw votes into the
// accumulator, a winning direction is walked pixel by pixel and its points are retired -- so ONE WARP runs one (image, colour)
// task: the 180 accumulator updates of a vote are dealt to the lanes (angle n belongs to lane n % 32, so no two lanes ever touch
// the same cell), everything scalar (the draw, the mask, the walk) is decided by lane 0 and broadcast.  Compiled for the host the
// same code runs with one lane: that build is what the CPU test compares with cv2.
#pragma once
#include <stdint.h>
#include <math.h>

#ifdef __CUDACC__
#define HP_FN __host__ __device__ __forceinline__
#else
#define HP_FN static inline
#endif

namespace hp {

constexpr int NUMANGLE = 180;            // cvRound(CV_PI / (float)(pi / 180))

struct Task {
    int width, height;
    int threshold, line_length, line_gap;   // cvRound of the Python arguments
    int numrho;                              // cvRound((width + height) * 2 + 1)
    const float *trig;                       // [NUMANGLE][2]: (float)cos(n * theta), (float)sin(n * theta), made on the host (its libm)
    int32_t *accum;                          // [NUMANGLE][numrho], zero on entry
    uint8_t *mask;                           // [height][width]: 1 = edge point not yet retired (filled by collect())
    uint32_t *nzloc;                         // [count]: y << 16 | x of the edge points in raster order
    int count;
    int32_t *lines;                          // out: [max_lines][4] = x1, y1, x2, y2
    int max_lines;
};

// Warp primitives: the device's own; one lane on the host; or -- HP_EMULATE_WARP, test infrastructure (oracle/csrc/hough_check.cpp) --
// 32 host threads that meet at a barrier for every shuffle, so that the lane protocol below (who computes, who broadcasts, which
// lane owns which accumulator cell, that every lane reaches every shuffle) runs on a CPU, under a race detector if wanted.
#if defined(__CUDA_ARCH__)
#define HP_LANE ((int)(threadIdx.x & 31))
#define HP_NLANES 32
#define HP_BCAST(x) __shfl_sync(0xffffffffu, (x), 0)
#define HP_SHFL_IDX(x, src) __shfl_sync(0xffffffffu, (x), (src))
#define HP_SHFL_UP(x, o) __shfl_up_sync(0xffffffffu, (x), (o))
#define HP_SHFL_XOR(x, o) __shfl_xor_sync(0xffffffffu, (x), (o))
#define HP_SYNC() __syncwarp()
#define HP_POPC(x) __popc(x)
HP_FN int cv_round(float v) { return __float2int_rn(v); }
#elif defined(HP_EMULATE_WARP)
int hp_emu_lane();                                   // provided by the emulation harness
int hp_emu_shfl(int value, int src_lane);            // every lane calls it; returns the value lane src_lane passed
void hp_emu_sync();
#define HP_LANE hp_emu_lane()
#define HP_NLANES 32
#define HP_BCAST(x) hp_emu_shfl((int)(x), 0)
#define HP_SHFL_IDX(x, src) hp_emu_shfl((int)(x), (src))
#define HP_SHFL_UP(x, o) hp_emu_shfl((int)(x), hp_emu_lane() >= (o) ? hp_emu_lane() - (o) : hp_emu_lane())
#define HP_SHFL_XOR(x, o) hp_emu_shfl((int)(x), hp_emu_lane() ^ (o))
#define HP_SYNC() hp_emu_sync()
#define HP_POPC(x) __builtin_popcount(x)
HP_FN int cv_round(float v) { return (int)lrintf(v); }
#else
#define HP_LANE 0
#define HP_NLANES 1
#define HP_BCAST(x) (x)
#define HP_SHFL_IDX(x, src) (x)
#define HP_SHFL_UP(x, o) (x)
#define HP_SHFL_XOR(x, o) (x)
#define HP_SYNC() do { } while (0)
#define HP_POPC(x) __builtin_popcount(x)
HP_FN int cv_round(float v) { return (int)lrintf(v); }      // cvRound(float) = cvtss2si: to nearest, ties to even
#endif

// host: the tables and sizes HoughLinesProbabilistic derives from (rho, theta) = (1, (float)(pi / 180))
inline int numrho(int width, int height) { return (width + height) * 2 + 1; }
inline void make_trig(float *trig /* [NUMANGLE][2] */)
{
    const float theta = (float)(3.14159265358979323846 / 180.0);
    for (int n = 0; n < NUMANGLE; ++n) {
        trig[n * 2] = (float)(cos((double)n * theta) * 1.0f);
        trig[n * 2 + 1] = (float)(sin((double)n * theta) * 1.0f);
    }
}

// cv::RNG (multiply-with-carry), RNG::uniform(int a, int b)
HP_FN uint32_t rng_next(uint64_t &state)
{
    state = (uint64_t)(uint32_t)state * 4164903690u + (uint32_t)(state >> 32);
    return (uint32_t)state;
}

// votes of one point: accum[n][r(n)] += delta for the lane's angles; returns through best_val / best_n the lane's largest cell after
// the update (first n wins among equals, like the sequential scan)
HP_FN void vote(const Task &t, int i, int j, int delta, int &best_val, int &best_n)
{
    for (int n = HP_LANE; n < NUMANGLE; n += HP_NLANES) {
        int r = cv_round((float)j * t.trig[n * 2] + (float)i * t.trig[n * 2 + 1]);
        r += (t.numrho - 1) / 2;
        const int val = (t.accum[(size_t)n * t.numrho + r] += delta);
        if (best_val < val) { best_val = val; best_n = n; }
    }
}

// Stage 1 of HoughLinesProbabilistic for one task: zero the accumulator, fill the byte mask and list the non-zero points of the
// bit-plane (bit x & 31 of word [y][x >> 5]) in raster order.  Lane l takes word base + l of every 32-word step; the offsets are a
// warp prefix sum of the population counts.  Returns the number of points (the same value in every lane).
HP_FN int collect(const uint32_t *plane, int h, int w, int wp, int32_t *accum, size_t acc_sz, uint8_t *mask, uint32_t *nz)
{
    const int lane = HP_LANE;
    for (size_t k = (size_t)lane; k < acc_sz; k += HP_NLANES) accum[k] = 0;
    int cnt = 0;
    const int nwords = h * wp;
    for (int base = 0; base < nwords; base += HP_NLANES) {
        const int wi = base + lane;
        uint32_t bits = 0;
        int y = 0, x0 = 0, valid = 0;
        if (wi < nwords) {
            y = wi / wp; x0 = (wi - y * wp) * 32;
            valid = w - x0 < 32 ? w - x0 : 32;          // pixels of this word inside the row
            bits = plane[wi];
            if (valid < 32) bits &= valid > 0 ? ((1u << valid) - 1u) : 0u;
        }
        const int mine = HP_POPC(bits);
        int incl = mine;
        for (int o = 1; o < HP_NLANES; o <<= 1) { const int v = HP_SHFL_UP(incl, o); if (lane >= o) incl += v; }
        int at = cnt + incl - mine;
        for (int b = 0; b < valid; ++b) {
            const uint32_t on = (bits >> b) & 1u;
            mask[(size_t)y * w + x0 + b] = (uint8_t)on;
            if (on) nz[at++] = ((uint32_t)y << 16) | (uint32_t)(x0 + b);
        }
        cnt += HP_SHFL_IDX(incl, HP_NLANES - 1);
    }
    HP_SYNC();                                          // accumulator zeroed and lists written by all lanes, read by others next
    return cnt;
}

// Returns the number of lines found (all lanes return the same value).  lane 0 owns mask / nzloc / lines.
HP_FN int hough_lines_p(const Task &t)
{
    uint64_t rng = 0xffffffffffffffffull;            // RNG rng((uint64)-1)
    const int width = t.width, height = t.height;
    const int shift = 16;
    int nlines = 0;
    const bool lead = HP_LANE == 0;
    for (int count = t.count; count > 0; --count) {
        // choose random point out of the remaining ones, "remove" it by overriding it with the last element
        int i = 0, j = 0, alive = 0;
        const uint32_t draw = rng_next(rng);         // every lane keeps the same generator state
        if (lead) {
            const int idx = (int)(draw % (uint32_t)count);       // uniform(0, count); count == 0 never gets here
            const uint32_t p = t.nzloc[idx];
            t.nzloc[idx] = t.nzloc[count - 1];
            i = (int)(p >> 16); j = (int)(p & 0xffffu);
            alive = t.mask[(size_t)i * width + j];   // 0: it has been excluded already (belongs to some other line)
        }
        i = HP_BCAST(i); j = HP_BCAST(j); alive = HP_BCAST(alive);
        if (!alive) continue;
        // update accumulator, find the most probable line
        int max_val = t.threshold - 1, max_n = 0;
        vote(t, i, j, +1, max_val, max_n);
        for (int o = HP_NLANES / 2; o > 0; o >>= 1) {   // largest value, smallest angle among equals = the sequential scan's answer
            const int v = HP_SHFL_XOR(max_val, o), n = HP_SHFL_XOR(max_n, o);
            if (v > max_val || (v == max_val && n < max_n)) { max_val = v; max_n = n; }
        }
        // if it is too "weak" candidate, continue with another point
        if (max_val < t.threshold) continue;
        // from the current point walk in each direction along the found line and extract the line segment
        const float a = -t.trig[max_n * 2 + 1], b = t.trig[max_n * 2];
        int x0 = j, y0 = i, dx0, dy0, xflag;
        if (fabsf(a) > fabsf(b)) {
            xflag = 1;
            dx0 = a > 0 ? 1 : -1;
            dy0 = cv_round(b * (float)(1 << shift) / fabsf(a));
            y0 = (y0 << shift) + (1 << (shift - 1));
        } else {
            xflag = 0;
            dy0 = b > 0 ? 1 : -1;
            dx0 = cv_round(a * (float)(1 << shift) / fabsf(b));
            x0 = (x0 << shift) + (1 << (shift - 1));
        }
        int end_x[2] = {0, 0}, end_y[2] = {0, 0};
        if (lead) {
            for (int k = 0; k < 2; ++k) {
                int gap = 0, x = x0, y = y0, dx = dx0, dy = dy0;
                if (k > 0) { dx = -dx; dy = -dy; }
                // walk along the line using fixed-point arithmetic, stop at the image border or in case of too big gap
                for (;; x += dx, y += dy) {
                    int i1, j1;
                    if (xflag) { j1 = x; i1 = y >> shift; } else { j1 = x >> shift; i1 = y; }
                    if (j1 < 0 || j1 >= width || i1 < 0 || i1 >= height) break;
                    // for each non-zero point: update line end, reset the gap
                    if (t.mask[(size_t)i1 * width + j1]) { gap = 0; end_y[k] = i1; end_x[k] = j1; }
                    else if (++gap > t.line_gap) break;
                }
            }
        }
        for (int k = 0; k < 2; ++k) { end_x[k] = HP_BCAST(end_x[k]); end_y[k] = HP_BCAST(end_y[k]); }
        const int adx = end_x[1] - end_x[0], ady = end_y[1] - end_y[0];
        const bool good_line = (adx < 0 ? -adx : adx) >= t.line_length || (ady < 0 ? -ady : ady) >= t.line_length;
        for (int k = 0; k < 2; ++k) {
            int x = x0, y = y0, dx = dx0, dy = dy0;
            if (k > 0) { dx = -dx; dy = -dy; }
            // walk again: retire the points of the segment (and take their votes back when the line is kept)
            for (;; x += dx, y += dy) {
                int i1, j1;
                if (xflag) { j1 = x; i1 = y >> shift; } else { j1 = x >> shift; i1 = y; }
                int m = 0;
                if (lead) {
                    uint8_t *md = t.mask + (size_t)i1 * width + j1;
                    m = *md;
                    *md = 0;
                }
                m = HP_BCAST(m);
                if (m && good_line) {
                    int dummy_v = 0x7fffffff, dummy_n = 0;
                    vote(t, i1, j1, -1, dummy_v, dummy_n);
                }
                if (i1 == end_y[k] && j1 == end_x[k]) break;
            }
        }
        if (good_line) {
            if (lead && nlines < t.max_lines) {
                int32_t *o = t.lines + (size_t)nlines * 4;
                o[0] = end_x[0]; o[1] = end_y[0]; o[2] = end_x[1]; o[3] = end_y[1];
            }
            ++nlines;
        }
    }
    HP_SYNC();
    return nlines;
}

// LineDetectorHSV._findNormal + _correctPixelOrdering (line_detector1.py:71-119) for ONE line of cv2.HoughLinesP, in numpy's dtypes:
// the lines are int32, so the squared length is an integer sum, everything after it float64; probe pixels are truncated toward
// zero and clamped (_checkBounds); the normal points from the coloured side to the other; the endpoints are swapped when
// (p2 - p1) x normal > 0.  bw_at(y, x) = the dilated colour mask.  min_line_length >= 1 keeps the length non-zero.
template <typename BW>
HP_FN void find_normal(int x1, int y1, int x2, int y2, int h, int w, const BW &bw_at, int32_t *line /* 4, ordered */, double *normal /* 2 */,
                       double *center /* 2 */)
{
    const long long ss = (long long)((x1 - x2) * (x1 - x2)) + (long long)((y1 - y2) * (y1 - y2));
    const double length = sqrt((double)ss);             // numpy: int64 ** 0.5 (== sqrt for every integer, checked up to 2e6)
    const double dx = (double)(y2 - y1) / length, dy = (double)(x1 - x2) / length;
    const double cx = (double)(x1 + x2) / 2.0, cy = (double)(y1 + y2) / 2.0;
    long long x3 = (long long)(cx - 3.0 * dx), y3 = (long long)(cy - 3.0 * dy), x4 = (long long)(cx + 3.0 * dx), y4 = (long long)(cy + 3.0 * dy);
    x3 = x3 < 0 ? 0 : (x3 >= w ? w - 1 : x3); x4 = x4 < 0 ? 0 : (x4 >= w ? w - 1 : x4);
    y3 = y3 < 0 ? 0 : (y3 >= h ? h - 1 : y3); y4 = y4 < 0 ? 0 : (y4 >= h ? h - 1 : y4);
    const double sign = (bw_at((int)y3, (int)x3) && !bw_at((int)y4, (int)x4)) ? 1.0 : -1.0;
    const double nx = dx * sign, ny = dy * sign;
    const bool swap = ((double)(x2 - x1) * ny - (double)(y2 - y1) * nx) > 0;
    line[0] = swap ? x2 : x1; line[1] = swap ? y2 : y1; line[2] = swap ? x1 : x2; line[3] = swap ? y1 : y2;
    normal[0] = nx; normal[1] = ny;
    center[0] = cx; center[1] = cy;
}

// line_detector_node.py toSegmentMsg (:251-265) for int32 lines: (lines + (0, cut, 0, cut)) * (1/W, 1/H, 1/W, 1/H) in float64, the
// Vector2D fields are float32 on the wire.  inv_w / inv_h = 1. / img_size[1], 1. / img_size[0] (the full image, before the cut).
HP_FN void normalized_fields(const int32_t *line, const double *normal, int top_cutoff, double inv_w, double inv_h, float *pixn /* 4 */,
                             float *nrm /* 2 */)
{
    pixn[0] = (float)((double)line[0] * inv_w);
    pixn[1] = (float)((double)(line[1] + top_cutoff) * inv_h);
    pixn[2] = (float)((double)line[2] * inv_w);
    pixn[3] = (float)((double)(line[3] + top_cutoff) * inv_h);
    nrm[0] = (float)normal[0]; nrm[1] = (float)normal[1];
}

}  // namespace hp
