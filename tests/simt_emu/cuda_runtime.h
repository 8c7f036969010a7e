// TEST INFRASTRUCTURE ONLY -- a host-side stand-in for <cuda_runtime.h> so that the PRODUCT's device code
// (steppingstone_b200/csrc/ss_device.cuh + ss_kernels.cu, unmodified apart from the kernel-launch syntax, which
// tests/simt_emu/build_emu.py rewrites textually) can be compiled with g++ and executed lane by lane on the CPU.
//
// Purpose: run the kernels' arithmetic against the oracle in the CPU test-suite (tests/test_simt_emu.py) and let a
// kernel change be checked for correctness without a GPU.  It is NOT a CPU path of the product: nothing under
// steppingstone_b200/ can load the library built from this (tests/simt_emu/_build/libss_emu.so, own ctypes binding in
// tests/simt_emu/binding.py), it is ~10^4 times slower than the GPU, and bench.py never sees it.
//
// Execution model: one CUDA thread = one fiber (ucontext); a block's fibers are scheduled round-robin on one OS thread;
// every warp collective (__shfl*_sync, __ballot_sync, __any/__all_sync, __match_any_sync, __syncwarp) and __syncthreads
// is a rendez-vous: a fiber deposits its operand and yields until all live lanes of its warp (threads of its block)
// have arrived.  Between collectives a fiber runs alone, so a shared-memory exchange that is not separated by a
// collective shows up as a wrong result here (on the GPU it would be a race).  Blocks run one after another.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

#define SS_SIMT_EMU 1
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __grid_constant__
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static

using std::max;
using std::min;

struct alignas(16) float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{ x, y, z, w }; }
struct dim3 { unsigned x = 1, y = 1, z = 1; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };

// ------------------------------------------------------------------------------------------------ runtime API (host memory)
typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyHostToHost };
enum { cudaStreamNonBlocking = 1 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize };
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated CUDA error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
template <typename T> static inline cudaError_t cudaMalloc(T** p, size_t n) { *p = (T*)aligned_alloc(256, (n + 255) / 256 * 256); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <typename T> static inline cudaError_t cudaMallocHost(T** p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t)(uintptr_t)1; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }

// ------------------------------------------------------------------------------------------------ fibers
namespace simt {

struct Fiber {
    ucontext_t ctx;           // portable fallback
    void* sp = nullptr;       // x86-64: saved stack pointer of the hand-written switch (no sigprocmask syscall per switch)
    char* stack = nullptr;
    dim3 tid;
    int lane = 0, warp = 0;
    unsigned op = 0;          // collectives executed so far (parity selects the exchange buffer)
    bool done = false;
};
struct WarpCtx {
    uint64_t slot[2][32];
    int arrived[2] = { 0, 0 }, readers[2] = { 0, 0 };
    int live = 32;
};
struct Block {
    std::vector<Fiber> fibers;
    std::vector<WarpCtx> warps;
    ucontext_t sched;
    void* sched_sp = nullptr;
    Fiber* cur = nullptr;
    int nthreads = 0, live = 0, bar_arrived = 0;
    unsigned bar_gen = 0;
    dim3 bid, bdim, gdim;
    std::function<void()> body;
};
extern Block* g_block;
extern unsigned char* g_dynsmem;
extern unsigned long long g_collectives;   // warp collectives executed (per lane), a cheap work counter for kernel variants

void yield();             // back to the block's scheduler (emu_runtime.cpp)

// rendez-vous of the live lanes of the calling fiber's warp: deposit `bits`, wait for the others, then `read` sees every
// lane's deposit in slot[]; buffers alternate by collective parity, so a lane that is already depositing for the next
// collective cannot overwrite what a slower lane still has to read
template <typename R>
static inline auto warp_collective(uint64_t bits, R read) {
    Fiber* f = g_block->cur;
    WarpCtx& w = g_block->warps[f->warp];
    const int p = f->op & 1;
    f->op++;
    g_collectives++;
    w.slot[p][f->lane] = bits;
    w.arrived[p]++;
    while (w.arrived[p] < w.live) yield();
    auto r = read(w.slot[p], f->lane);
    if (++w.readers[p] == w.live) { w.arrived[p] = 0; w.readers[p] = 0; }
    else {
        // the buffer is recycled two collectives later; nobody can get there before every lane has read (see above)
    }
    return r;
}
template <typename T> static inline uint64_t to_bits(T v) { uint64_t b = 0; static_assert(sizeof(T) <= 8, ""); memcpy(&b, &v, sizeof(T)); return b; }
template <typename T> static inline T from_bits(uint64_t b) { T v; memcpy(&v, &b, sizeof(T)); return v; }

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);

}  // namespace simt

#define threadIdx (simt::g_block->cur->tid)
#define blockIdx (simt::g_block->bid)
#define blockDim (simt::g_block->bdim)
#define gridDim (simt::g_block->gdim)

template <typename T> static inline T __shfl_sync(unsigned, T v, int src, int = 32) {
    return simt::warp_collective(simt::to_bits(v), [&](const uint64_t* s, int) { return simt::from_bits<T>(s[src & 31]); });
}
template <typename T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) {
    return simt::warp_collective(simt::to_bits(v), [&](const uint64_t* s, int l) { return simt::from_bits<T>(s[l >= (int)d ? l - (int)d : l]); });
}
template <typename T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) {
    return simt::warp_collective(simt::to_bits(v), [&](const uint64_t* s, int l) { return simt::from_bits<T>(s[l + (int)d < 32 ? l + (int)d : l]); });
}
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) {
    return simt::warp_collective(simt::to_bits(v), [&](const uint64_t* s, int l) { return simt::from_bits<T>(s[(l ^ m) & 31]); });
}
static inline unsigned __ballot_sync(unsigned, int pred) {
    return simt::warp_collective((uint64_t)(pred != 0), [&](const uint64_t* s, int) { unsigned m = 0; for (int i = 0; i < 32; i++) m |= (unsigned)(s[i] & 1) << i; return m; });
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, pred) == 0xffffffffu; }
template <typename T> static inline unsigned __match_any_sync(unsigned, T v) {
    return simt::warp_collective(simt::to_bits(v), [&](const uint64_t* s, int l) { unsigned m = 0; for (int i = 0; i < 32; i++) m |= (unsigned)(s[i] == s[l]) << i; return m; });
}
static inline void __syncwarp(unsigned = 0xffffffffu) { simt::warp_collective(0, [](const uint64_t*, int) { return 0; }); }
static inline void __syncthreads() {
    simt::Block* b = simt::g_block;
    const unsigned gen = b->bar_gen;
    if (++b->bar_arrived == b->live) { b->bar_arrived = 0; b->bar_gen++; }
    else while (b->bar_gen == gen) simt::yield();
}

static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline double rsqrt(double x) { return 1.0 / sqrt(x); }
template <typename T> static inline T atomicAdd(T* p, T v) { T old = *p; *p = old + v; return old; }   // fibers never pre-empt each other
template <typename T> static inline T atomicOr(T* p, T v) { T old = *p; *p = old | v; return old; }
static inline void __stcs(float4* p, float4 v) { *p = v; }
template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }
