// cuda_emu.h -- TEST INFRASTRUCTURE ONLY: a tiny CUDA-on-CPU shim.
//
// The build container has nvcc but no GPU.  To debug the *logic* of the kernels in
// colorvideovdp_b200/csrc (tile/halo indexing, padding rules, reductions) without spending GPU time,
// tests/emu/build_emu.sh compiles the very same .cu sources with g++ against this header into
// tests/emu/libcvvdp_b200_emu.so.  Every CUDA thread is a ucontext fiber; __syncthreads() and the warp
// shuffles are cooperative yield points, blocks run one after another.  It is a mock device for the
// `-m "not gpu"` tests; the product package never loads it (it only ever loads the nvcc-built library
// and refuses to run without a CUDA device).
#pragma once
#ifndef CVVDP_EMU
#error "cuda_emu.h is only for the CPU emulation build"
#endif

#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define __constant__ static
#define __grid_constant__

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3_emu { unsigned x, y, z; };
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct int2 { int x, y; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }

namespace emu {
enum State { RUNNABLE = 0, WAIT_BLOCK = 1, WAIT_WARP = 2, DONE = 3 };
struct Fiber {
    ucontext_t ctx;
    State state;
    uint3_emu tid;
};
struct Sched {
    ucontext_t main_ctx;
    std::vector<Fiber> fibers;
    std::vector<char *> stacks;
    Fiber *cur = nullptr;
    uint3_emu blockIdx_{0, 0, 0};
    dim3 blockDim_, gridDim_;
    unsigned char *dyn_smem = nullptr;
    size_t dyn_smem_cap = 0;
    std::function<void()> body;
    uint64_t shfl_scratch[32 * 64];  // [warp][lane]
};
inline Sched &S() {
    static Sched s;
    return s;
}
static const size_t kStack = 256 * 1024;

inline void yield_to_main() { swapcontext(&S().cur->ctx, &S().main_ctx); }
inline void fiber_entry() {
    S().body();
    S().cur->state = DONE;
    yield_to_main();
}
inline void syncthreads() {
    S().cur->state = WAIT_BLOCK;
    yield_to_main();
}
inline void yield_runnable() {  // spin-wait helper (mbarrier emulation): stay runnable, let others run
    S().cur->state = RUNNABLE;
    yield_to_main();
}
inline void syncwarp() {
    S().cur->state = WAIT_WARP;
    yield_to_main();
}

inline void run_block(unsigned nthreads) {
    Sched &s = S();
    if (s.fibers.size() < nthreads) s.fibers.resize(nthreads);
    while (s.stacks.size() < nthreads) s.stacks.push_back((char *)malloc(kStack));
    for (unsigned t = 0; t < nthreads; ++t) {
        Fiber &f = s.fibers[t];
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = s.stacks[t];
        f.ctx.uc_stack.ss_size = kStack;
        f.ctx.uc_link = &s.main_ctx;
        makecontext(&f.ctx, (void (*)())fiber_entry, 0);
        f.state = RUNNABLE;
        f.tid.x = t % s.blockDim_.x;
        f.tid.y = (t / s.blockDim_.x) % s.blockDim_.y;
        f.tid.z = t / (s.blockDim_.x * s.blockDim_.y);
    }
    const unsigned nwarps = (nthreads + 31) / 32;
    for (;;) {
        bool progressed = false;
        for (unsigned t = 0; t < nthreads; ++t) {
            Fiber &f = s.fibers[t];
            if (f.state != RUNNABLE) continue;
            s.cur = &f;
            swapcontext(&s.main_ctx, &f.ctx);
            progressed = true;
        }
        // release warps whose live lanes all wait at a warp sync
        bool released = false;
        for (unsigned w = 0; w < nwarps; ++w) {
            unsigned lo = w * 32, hi = std::min(nthreads, lo + 32);
            bool all = true, any = false;
            for (unsigned t = lo; t < hi; ++t) {
                if (s.fibers[t].state == WAIT_WARP) any = true;
                else if (s.fibers[t].state != DONE) all = false;
            }
            if (any && all) {
                for (unsigned t = lo; t < hi; ++t)
                    if (s.fibers[t].state == WAIT_WARP) s.fibers[t].state = RUNNABLE;
                released = true;
            }
        }
        if (released) continue;
        bool all = true, any = false, alive = false;
        for (unsigned t = 0; t < nthreads; ++t) {
            State st = s.fibers[t].state;
            if (st != DONE) alive = true;
            if (st == WAIT_BLOCK) any = true;
            else if (st != DONE) all = false;
        }
        if (!alive) break;
        if (any && all) {
            for (unsigned t = 0; t < nthreads; ++t)
                if (s.fibers[t].state == WAIT_BLOCK) s.fibers[t].state = RUNNABLE;
            continue;
        }
        if (!progressed) {
            fprintf(stderr, "cuda_emu: deadlock (divergent barrier / partial-warp shuffle)\n");
            abort();
        }
    }
}

template <typename Fn>
inline void launch(dim3 grid, dim3 block, size_t smem, Fn fn) {
    Sched &s = S();
    s.gridDim_ = grid;
    s.blockDim_ = block;
    if (smem > s.dyn_smem_cap) {
        free(s.dyn_smem);
        s.dyn_smem = (unsigned char *)aligned_alloc(1024, (smem + 1023) / 1024 * 1024);
        s.dyn_smem_cap = smem;
    }
    s.body = fn;
    unsigned nthreads = block.x * block.y * block.z;
    for (unsigned z = 0; z < grid.z; ++z)
        for (unsigned y = 0; y < grid.y; ++y)
            for (unsigned x = 0; x < grid.x; ++x) {
                s.blockIdx_ = uint3_emu{x, y, z};
                if (smem) memset(s.dyn_smem, 0xCD, smem);  // poison: catch reads of unwritten smem
                run_block(nthreads);
            }
}

template <typename T>
inline T shfl_generic(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shfl payload");
    Sched &s = S();
    unsigned lin = s.cur->tid.x + s.cur->tid.y * s.blockDim_.x + s.cur->tid.z * s.blockDim_.x * s.blockDim_.y;
    unsigned w = lin / 32, lane = lin % 32;
    uint64_t bits = 0;
    memcpy(&bits, &v, sizeof(T));
    s.shfl_scratch[w * 32 + lane] = bits;
    syncwarp();
    uint64_t r = s.shfl_scratch[w * 32 + (src_lane & 31)];
    syncwarp();
    T out;
    memcpy(&out, &r, sizeof(T));
    return out;
}
inline unsigned lane_id() {
    Sched &s = S();
    return (s.cur->tid.x + s.cur->tid.y * s.blockDim_.x + s.cur->tid.z * s.blockDim_.x * s.blockDim_.y) % 32;
}
}  // namespace emu

#define threadIdx (emu::S().cur->tid)
#define blockIdx (emu::S().blockIdx_)
#define blockDim (emu::S().blockDim_)
#define gridDim (emu::S().gridDim_)
#define __syncthreads() emu::syncthreads()
#define __syncwarp(...) emu::syncwarp()
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int m) { return emu::shfl_generic(v, (int)emu::lane_id() ^ m); }
template <typename T> inline T __shfl_down_sync(unsigned, T v, int d) {
    int l = (int)emu::lane_id() + d;
    return emu::shfl_generic(v, l > 31 ? (int)emu::lane_id() : l);
}
template <typename T> inline T __shfl_sync(unsigned, T v, int l) { return emu::shfl_generic(v, l); }
template <typename T> inline T __ldg(const T *p) { return *p; }
template <typename T> inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
template <typename T> inline T atomicMin(T *p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <typename T> inline T atomicMax(T *p, T v) { T o = *p; if (v > o) *p = v; return o; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }

static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float __saturatef(float a) { return a < 0 ? 0.f : (a > 1 ? 1.f : a); }
static inline float exp10f_emu(float a) { return powf(10.f, a); }
static inline float rsqrtf(float a) { return 1.0f / sqrtf(a); }

// ---- runtime API ----
typedef int cudaError_t;
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16, cudaDevAttrMaxSharedMemoryPerBlockOptin = 97 };
static inline const char *cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaMalloc(void **p, size_t n) {
    *p = aligned_alloc(256, (n + 255) / 256 * 256 + 256);
    if (*p) memset(*p, 0xCD, n);  // poison
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
template <typename T> static inline cudaError_t cudaMalloc(T **p, size_t n) { return cudaMalloc((void **)p, n); }
static inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void **p, size_t n) { *p = malloc(n); return cudaSuccess; }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = 0) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (void *)1; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = (void *)1; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
template <typename T> static inline cudaError_t cudaFuncSetAttribute(T, cudaFuncAttribute, int) { return cudaSuccess; }
struct cudaFuncAttributes { size_t sharedSizeBytes = 0; };
template <typename T> static inline cudaError_t cudaFuncGetAttributes(cudaFuncAttributes *a, T) { *a = cudaFuncAttributes(); return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr a, int) {
    *v = (a == cudaDevAttrMultiProcessorCount) ? 4 : 232448;
    return cudaSuccess;
}
static inline cudaError_t cudaMemGetInfo(size_t *f, size_t *t) { *f = *t = (size_t)8 << 30; return cudaSuccess; }
