// TEST INFRASTRUCTURE: see cuda_threads_emu.h
#include "cuda_threads_emu.h"

namespace cuemu {
pthread_barrier_t block_bar, warp_bar[MAX_WARPS];
int warp_slot[MAX_WARPS][32];
int n_threads;
}  // namespace cuemu
thread_local cuemu::Idx threadIdx, blockIdx;
cuemu::Idx blockDim, gridDim;

namespace {
struct Start { int tid, block, block_y; void (*body)(void *); void *arg; };
void *thread_main(void *p)
{
    Start *s = (Start *)p;
    threadIdx = cuemu::Idx{(unsigned)s->tid, 0, 0};
    blockIdx = cuemu::Idx{(unsigned)s->block, (unsigned)s->block_y, 0};
    s->body(s->arg);
    return nullptr;
}
}  // namespace

void cuemu_run_block(int nthreads, int block, void (*body)(void *), void *arg, int block_y)
{
    cuemu::n_threads = nthreads;
    blockDim = cuemu::Idx{(unsigned)nthreads, 1, 1};
    pthread_barrier_init(&cuemu::block_bar, nullptr, nthreads);
    for (int w = 0; w < nthreads / 32; ++w) pthread_barrier_init(&cuemu::warp_bar[w], nullptr, 32);
    pthread_t *th = new pthread_t[nthreads];
    Start *st = new Start[nthreads];
    pthread_attr_t at;
    pthread_attr_init(&at);
    pthread_attr_setstacksize(&at, 1 << 20);
    for (int t = 0; t < nthreads; ++t) {
        st[t] = Start{t, block, block_y, body, arg};
        if (pthread_create(&th[t], &at, thread_main, &st[t])) { fprintf(stderr, "cuemu: pthread_create failed\n"); abort(); }
    }
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], nullptr);
    pthread_barrier_destroy(&cuemu::block_bar);
    for (int w = 0; w < nthreads / 32; ++w) pthread_barrier_destroy(&cuemu::warp_bar[w]);
    delete[] th;
    delete[] st;
}
