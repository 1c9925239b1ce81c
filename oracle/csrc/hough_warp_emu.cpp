// TEST INFRASTRUCTURE: the warp protocol of lane_slam_b200/csrc/hough_core.cuh (what k_hough_p runs per task: collect() and
// hough_lines_p() with 32 lanes) executed by 32 host threads that meet at a barrier for every shuffle / __syncwarp.  A lane that
// skips a shuffle the others reach deadlocks here like it would be undefined on the device; cells of the accumulator, the mask and
// the point list are shared memory here as there.  Built and used by tests/test_hough_core.py (optionally with -fsanitize=thread).
#define HP_EMULATE_WARP
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../lane_slam_b200/csrc/hough_core.cuh"

static thread_local int t_lane;
static pthread_barrier_t g_bar;
static int g_slot[32];

namespace hp {
int hp_emu_lane() { return t_lane; }
int hp_emu_shfl(int value, int src_lane)
{
    g_slot[t_lane] = value;
    pthread_barrier_wait(&g_bar);
    const int r = g_slot[src_lane & 31];
    pthread_barrier_wait(&g_bar);
    return r;
}
void hp_emu_sync() { pthread_barrier_wait(&g_bar); }
}  // namespace hp

struct Job {
    const uint32_t *plane; int h, w, wp, th, ml, mg;
    int32_t *accum; size_t acc_sz; uint8_t *mask; uint32_t *nz; const float *trig; int32_t *lines; int max_lines;
    int lane, result, count;
};

static void *lane_main(void *p)
{
    Job *j = (Job *)p;
    t_lane = j->lane;
    const int cnt = hp::collect(j->plane, j->h, j->w, j->wp, j->accum, j->acc_sz, j->mask, j->nz);
    hp::Task t;
    t.width = j->w; t.height = j->h; t.threshold = j->th; t.line_length = j->ml; t.line_gap = j->mg;
    t.numrho = hp::numrho(j->w, j->h); t.trig = j->trig; t.accum = j->accum; t.mask = j->mask; t.nzloc = j->nz; t.count = cnt;
    t.lines = j->lines; t.max_lines = j->max_lines;
    j->count = cnt;
    j->result = hp::hough_lines_p(t);
    return nullptr;
}

// edge: [h][w] bytes.  The bit-plane handed to the lanes has garbage in the bits beyond w, the scratch buffers start as garbage.
extern "C" __attribute__((visibility("default")))
int hpe_hough_lines_p(const uint8_t *edge, int h, int w, int th, int ml, int mg, int32_t *lines, int max_lines, int *count_out)
{
    const int wp = (w + 31) / 32;
    std::vector<uint32_t> plane((size_t)h * wp, 0);
    for (int y = 0; y < h; ++y) {
        for (int x = 0; x < w; ++x)
            if (edge[(size_t)y * w + x]) plane[(size_t)y * wp + (x >> 5)] |= 1u << (x & 31);
        if (w & 31) plane[(size_t)y * wp + wp - 1] |= ~0u << (w & 31);
    }
    const size_t acc_sz = (size_t)hp::NUMANGLE * hp::numrho(w, h);
    std::vector<int32_t> accum(acc_sz, 0x55555555);
    std::vector<uint8_t> mask((size_t)h * w, 0xCC);
    std::vector<uint32_t> nz((size_t)h * w, 0xDEADBEEF);
    std::vector<float> trig(hp::NUMANGLE * 2);
    hp::make_trig(trig.data());
    pthread_barrier_init(&g_bar, nullptr, 32);
    Job jobs[32];
    pthread_t th_[32];
    for (int l = 0; l < 32; ++l) {
        jobs[l] = Job{plane.data(), h, w, wp, th, ml, mg, accum.data(), acc_sz, mask.data(), nz.data(), trig.data(), lines, max_lines, l, 0, 0};
        pthread_create(&th_[l], nullptr, lane_main, &jobs[l]);
    }
    for (int l = 0; l < 32; ++l) pthread_join(th_[l], nullptr);
    pthread_barrier_destroy(&g_bar);
    for (int l = 1; l < 32; ++l)
        if (jobs[l].result != jobs[0].result || jobs[l].count != jobs[0].count) return -1;       // the lanes must agree
    if (count_out) *count_out = jobs[0].count;
    return jobs[0].result;
}
