// cvvdp_common.cuh -- platform switch (nvcc / CPU emulation for tests), small device helpers.
#pragma once

#ifdef CVVDP_EMU
#include "cuda_emu.h"  // tests/emu: mock CUDA runtime for the no-GPU logic tests, never shipped
#define CVVDP_DYN_SMEM(name) unsigned char *name = emu::S().dyn_smem
#define CVVDP_LAUNCH(kfn, grid, block, smem, stream, ...) \
    emu::launch(grid, block, smem, [&]() { kfn(__VA_ARGS__); })
#else
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#define CVVDP_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define CVVDP_LAUNCH(kfn, grid, block, smem, stream, ...) kfn<<<grid, block, smem, stream>>>(__VA_ARGS__)
#endif

#include <stdint.h>

#include "../../include/cvvdp_b200.h"

namespace cvvdp {

// ---- fast transcendental helpers (MUFU on the GPU, libm in the emulation build) --------------
__device__ __forceinline__ float f_lg2(float x) {
#ifdef __CUDA_ARCH__
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return log2f(x);
#endif
}
__device__ __forceinline__ float f_ex2(float x) {
#ifdef __CUDA_ARCH__
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return exp2f(x);
#endif
}
__device__ __forceinline__ float f_rcp(float x) {
#ifdef __CUDA_ARCH__
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / x;
#endif
}
// x^p for x >= 0 (x == 0 -> 0 for p > 0)
__device__ __forceinline__ float f_pow(float x, float p) { return f_ex2(p * f_lg2(x)); }

__device__ __forceinline__ float half_bits_to_float(unsigned short h) {
#if defined(__CUDA_ARCH__)
    return __half2float(__ushort_as_half(h));
#else
    unsigned sign = (h >> 15) & 1u, ex = (h >> 10) & 0x1Fu, man = h & 0x3FFu;
    float v;
    if (ex == 0) v = ldexpf((float)man, -24);
    else if (ex == 31) v = man ? NAN : INFINITY;
    else v = ldexpf((float)(man | 0x400u), (int)ex - 25);
    return sign ? -v : v;
#endif
}
__device__ __forceinline__ unsigned short float_to_half_bits(float f) {
#if defined(__CUDA_ARCH__)
    return __half_as_ushort(__float2half_rn(f));
#else
    uint32_t x;
    memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    int32_t ex = (int32_t)((x >> 23) & 0xFF) - 127 + 15;
    uint32_t man = x & 0x7FFFFFu;
    if (((x >> 23) & 0xFF) == 0xFF) return (unsigned short)(sign | 0x7C00u | (man ? 0x200u : 0));
    if (ex >= 31) return (unsigned short)(sign | 0x7C00u);
    if (ex <= 0) {
        if (ex < -10) return (unsigned short)sign;
        man |= 0x800000u;
        int shift = 14 - ex;
        uint32_t hm = man >> shift, rem = man & ((1u << shift) - 1), half = 1u << (shift - 1);
        if (rem > half || (rem == half && (hm & 1))) hm++;
        return (unsigned short)(sign | hm);
    }
    uint32_t hm = man >> 13, rem = man & 0x1FFFu;
    uint32_t out = sign | ((uint32_t)ex << 10) | hm;
    if (rem > 0x1000u || (rem == 0x1000u && (hm & 1))) out++;
    return (unsigned short)out;
#endif
}

// ---- Ampere-style asynchronous global->shared copies (LDGSTS), 16 bytes each ---------------------
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
#ifdef __CUDA_ARCH__
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src) : "memory");
#else
    memcpy(smem_dst, gmem_src, 16);
#endif
}
__device__ __forceinline__ void cp_async_commit() {
#ifdef __CUDA_ARCH__
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
__device__ __forceinline__ void cp_async_wait_all() {
#ifdef __CUDA_ARCH__
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
}

// ---- float4 arithmetic ------------------------------------------------------------------------
__device__ __forceinline__ float4 f4(float v) { return make_float4(v, v, v, v); }
__device__ __forceinline__ float4 operator+(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 operator-(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
__device__ __forceinline__ float4 operator*(float s, float4 a) { return make_float4(s * a.x, s * a.y, s * a.z, s * a.w); }
__device__ __forceinline__ float4 fma4(float s, float4 a, float4 acc) {
    return make_float4(fmaf(s, a.x, acc.x), fmaf(s, a.y, acc.y), fmaf(s, a.z, acc.z), fmaf(s, a.w, acc.w));
}

// ---- kernel argument structures ----------------------------------------------------------------
struct ClipView {
    const void *data;
    long long s[5];  // element strides B, C, F, H, W
    int frame0, n_frames;
    int ring;  // > 0: frames live in a ring of that many slots (frame f at slot f % ring); 0: linear view
};
__device__ __forceinline__ int frame_slot(const ClipView &cv, int f) { return cv.ring > 0 ? f % cv.ring : f - cv.frame0; }

struct DisplayDev {
    int eotf;
    float gamma;  // EOTF_GAMMA exponent, or HLG system gamma
    float Ypeak, Yblack, Yrefl, exposure;
    float lin_lo;  // max(0.005, Yblack) (display_model.py:349)
    float M[9];    // RGB -> DKLd65, fp32, row-major
};

}  // namespace cvvdp
