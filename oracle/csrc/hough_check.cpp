// TEST INFRASTRUCTURE: lane_slam_b200/csrc/hough_core.cuh (the product's progressive probabilistic Hough transform) compiled for
// the host -- one lane -- so that the CPU tests can compare it with cv2.HoughLinesP.  Built and used by tests/test_hough_core.py.
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../lane_slam_b200/csrc/hough_core.cuh"

extern "C" __attribute__((visibility("default")))
int hpc_hough_lines_p(const uint8_t *edge, int height, int width, int threshold, int line_length, int line_gap, int32_t *lines, int max_lines)
{
    std::vector<float> trig(hp::NUMANGLE * 2);
    hp::make_trig(trig.data());
    hp::Task t;
    t.width = width; t.height = height; t.threshold = threshold; t.line_length = line_length; t.line_gap = line_gap;
    t.numrho = hp::numrho(width, height);
    std::vector<int32_t> accum((size_t)hp::NUMANGLE * t.numrho, 0);
    std::vector<uint8_t> mask((size_t)width * height, 0);
    std::vector<uint32_t> nz;
    for (int y = 0; y < height; ++y)
        for (int x = 0; x < width; ++x)
            if (edge[(size_t)y * width + x]) { mask[(size_t)y * width + x] = 1; nz.push_back(((uint32_t)y << 16) | (uint32_t)x); }
    t.trig = trig.data(); t.accum = accum.data(); t.mask = mask.data(); t.nzloc = nz.data(); t.count = (int)nz.size();
    t.lines = lines; t.max_lines = max_lines;
    return hp::hough_lines_p(t);
}

struct ByteMask {
    const uint8_t *p; int w;
    bool operator()(int y, int x) const { return p[(size_t)y * w + x] != 0; }
};

// the lines of hpc_hough_lines_p -> ordered lines, normals, centres, wire fields (what k_hough_segments computes per line)
extern "C" __attribute__((visibility("default")))
void hpc_find_normals(const int32_t *lines, int n, const uint8_t *bw, int h, int w, int top_cutoff, int img_h, int img_w, int32_t *ordered,
                      double *normals, double *centers, float *pixn, float *nrm)
{
    const ByteMask m = {bw, w};
    for (int i = 0; i < n; ++i) {
        hp::find_normal(lines[4 * i], lines[4 * i + 1], lines[4 * i + 2], lines[4 * i + 3], h, w, m, ordered + 4 * i, normals + 2 * i, centers + 2 * i);
        hp::normalized_fields(ordered + 4 * i, normals + 2 * i, top_cutoff, 1.0 / (double)img_w, 1.0 / (double)img_h, pixn + 4 * i, nrm + 2 * i);
    }
}
