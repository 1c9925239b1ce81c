// TEST INFRASTRUCTURE: enough of the CUDA execution model to run ONE thread block of a kernel -- its real source text -- on host
// threads: threadIdx / blockIdx, __shared__ (function-local statics), __syncthreads / __syncwarp (pthread barriers), the warp
// shuffles / ballot / reduce (a slot per lane between two warp barriers: a lane that does not reach a full-mask shuffle hangs the
// run, as it would be undefined on the device), atomics.  Every thread is a pthread and the barriers are the only
// synchronisation, so ThreadSanitizer reports exactly the shared-memory races a missing __syncthreads leaves (what
// compute-sanitizer --tool racecheck reports on the device).  Used by oracle/csrc/jpeg_huff_emu.cpp.
#pragma once
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) alignas(n)

namespace cuemu {
constexpr int MAX_WARPS = 32;
struct Idx { unsigned x, y, z; };
extern pthread_barrier_t block_bar, warp_bar[MAX_WARPS];
extern int warp_slot[MAX_WARPS][32];
extern int n_threads;
}  // namespace cuemu
extern thread_local cuemu::Idx threadIdx, blockIdx;
extern cuemu::Idx blockDim, gridDim;

static inline void __syncthreads() { pthread_barrier_wait(&cuemu::block_bar); }
static inline void __syncwarp(unsigned = 0xffffffffu) { pthread_barrier_wait(&cuemu::warp_bar[threadIdx.x >> 5]); }

static inline int cuemu_exchange(int v, int src_lane)
{
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    cuemu::warp_slot[w][l] = v;
    pthread_barrier_wait(&cuemu::warp_bar[w]);
    const int r = cuemu::warp_slot[w][src_lane & 31];
    pthread_barrier_wait(&cuemu::warp_bar[w]);
    return r;
}
static inline int __shfl_sync(unsigned, int v, int src) { return cuemu_exchange(v, src); }
static inline unsigned __shfl_sync(unsigned, unsigned v, int src) { return (unsigned)cuemu_exchange((int)v, src); }
static inline int __shfl_up_sync(unsigned, int v, unsigned d) { const int l = threadIdx.x & 31; return cuemu_exchange(v, l >= (int)d ? l - (int)d : l); }
static inline unsigned __shfl_up_sync(unsigned m, unsigned v, unsigned d) { return (unsigned)__shfl_up_sync(m, (int)v, d); }
static inline int __shfl_xor_sync(unsigned, int v, int x) { return cuemu_exchange(v, (threadIdx.x & 31) ^ x); }
static inline unsigned __ballot_sync(unsigned, int pred)
{
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    cuemu::warp_slot[w][l] = pred ? 1 : 0;
    pthread_barrier_wait(&cuemu::warp_bar[w]);
    unsigned r = 0;
    for (int k = 0; k < 32; ++k) r |= (unsigned)cuemu::warp_slot[w][k] << k;
    pthread_barrier_wait(&cuemu::warp_bar[w]);
    return r;
}
static inline int __reduce_add_sync(unsigned, int v)
{
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    cuemu::warp_slot[w][l] = v;
    pthread_barrier_wait(&cuemu::warp_bar[w]);
    int r = 0;
    for (int k = 0; k < 32; ++k) r += cuemu::warp_slot[w][k];
    pthread_barrier_wait(&cuemu::warp_bar[w]);
    return r;
}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline unsigned atomicOr(unsigned *p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
static inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline int atomicMax(int *p, int v)
{
    int o = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (o < v && !__atomic_compare_exchange_n(p, &o, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) { }
    return o;
}
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
// shared-memory "addresses" of the kernels' inline-PTX loads: the low 32 bits of the host address of a static object
static inline uint32_t __cvta_generic_to_shared(const void *p) { return (uint32_t)(uintptr_t)p; }

static inline int __float2int_rn(float v) { return (int)lrintf(v); }
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
template <typename T> static inline T __ldg(const T *p) { return *p; }

// run `body(arg)` as one block of nthreads threads (a multiple of 32) with blockIdx = (block, block_y); gridDim is the caller's to set
void cuemu_run_block(int nthreads, int block, void (*body)(void *), void *arg, int block_y = 0);
