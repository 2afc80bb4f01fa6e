// cvvdp_common.cuh -- platform switch (nvcc / CPU emulation for tests), small device helpers.
#pragma once

#ifdef CVVDP_EMU
#include "cuda_emu.h"  // tests/emu: mock CUDA runtime for the no-GPU logic tests, never shipped
#define CVVDP_DYN_SMEM(name) unsigned char *name = emu::S().dyn_smem
#define CVVDP_LAUNCH(kfn, grid, block, smem, stream, ...) \
    emu::launch(grid, block, smem, [&]() { kfn(__VA_ARGS__); })
#else
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#define CVVDP_DYN_SMEM(name) extern __shared__ __align__(1024) unsigned char name[]
#define CVVDP_LAUNCH(kfn, grid, block, smem, stream, ...) kfn<<<grid, block, smem, stream>>>(__VA_ARGS__)
#endif

#include <stdint.h>

#include <type_traits>

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
template <int N>
__device__ __forceinline__ void cp_async_wait_n() {  // at most N of this thread's groups still in flight
#ifdef __CUDA_ARCH__
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
#endif
}
__device__ __forceinline__ void cp_async_wait_all() {
#ifdef __CUDA_ARCH__
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
}

// ---- TMA (cp.async.bulk.tensor) + mbarrier ---------------------------------------------------------
// A 3-D tiled tensor map over an array of float4 planes viewed as fp32 [plane][row][4*w].
#ifdef CVVDP_EMU
struct TensorMap3D {
    const float *base;
    int dim[3];          // 4*w, h, planes
    long long stride[3]; // in floats
    int box[3];
};
#else
typedef CUtensorMap TensorMap3D;
#endif

#ifdef CVVDP_EMU
// mock device: the copy is synchronous; *bar counts completed phases
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int) { *bar = 0; }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *, unsigned) {}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    while (((*(volatile unsigned long long *)bar) & 1ull) == parity) emu::yield_runnable();
}
__device__ __forceinline__ void fence_proxy_async() {}
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const TensorMap3D *map, int c0, int c1, int c2,
                                            unsigned long long *) {
    float *dst = (float *)smem_dst;
    for (int z = 0; z < map->box[2]; ++z)
        for (int y = 0; y < map->box[1]; ++y)
            for (int x = 0; x < map->box[0]; ++x) {
                const int gx = c0 + x, gy = c1 + y, gz = c2 + z;
                const bool in = gx >= 0 && gx < map->dim[0] && gy >= 0 && gy < map->dim[1] && gz >= 0 && gz < map->dim[2];
                dst[((long long)z * map->box[1] + y) * map->box[0] + x] =
                    in ? map->base[gz * map->stride[2] + gy * map->stride[1] + gx * map->stride[0]] : 0.f;
            }
}
__device__ __forceinline__ void mbar_emu_complete(unsigned long long *bar) { *bar += 1; }
#else
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
    unsigned ok;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// Tile load: box at element coordinates (c0, c1, c2) = (4*x, row, plane); out-of-bounds elements are
// zero-filled; completion is signalled on `bar` with the full box byte count.
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const TensorMap3D *map, int c0, int c1, int c2,
                                            unsigned long long *bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            (unsigned)__cvta_generic_to_shared(smem_dst)),
        "l"(map), "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void mbar_emu_complete(unsigned long long *) {}
#endif

// ---- 1-D bulk copies (cp.async.bulk, the TMA engine without a tensor map): 16-byte aligned source,
// destination and size; completion is signalled on an mbarrier like the tensor copies -------------------
#ifdef CVVDP_EMU
__device__ __forceinline__ void bulk_copy_g2s(void *smem_dst, const void *gmem_src, unsigned bytes, unsigned long long *) {
    memcpy(smem_dst, gmem_src, bytes);
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *) {}
#else
__device__ __forceinline__ void bulk_copy_g2s(void *smem_dst, const void *gmem_src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
#endif

// ---- float4 arithmetic ------------------------------------------------------------------------
// On sm_100 the four lanes of a pixel are processed as two packed fp32x2 operations (FFMA2 / FADD2 /
// FMUL2, PTX fma.rn.f32x2): same IEEE results per lane, half the issue slots -- these kernels are
// bound by instruction issue, not by the FP32 lanes.
__device__ __forceinline__ float4 f4(float v) { return make_float4(v, v, v, v); }
// packed pairs (one FFMA2 / FMUL2 / FADD2 each on sm_100)
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
#if defined(__CUDA_ARCH__) && !defined(CVVDP_NO_F32X2)
    return __ffma2_rn(a, b, c);
#else
    return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
#if defined(__CUDA_ARCH__) && !defined(CVVDP_NO_F32X2)
    return __fmul2_rn(a, b);
#else
    return make_float2(a.x * b.x, a.y * b.y);
#endif
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
#if defined(__CUDA_ARCH__) && !defined(CVVDP_NO_F32X2)
    return __fadd2_rn(a, b);
#else
    return make_float2(a.x + b.x, a.y + b.y);
#endif
}
__device__ __forceinline__ float2 bc2(float v) { return make_float2(v, v); }
#if defined(__CUDA_ARCH__) && !defined(CVVDP_NO_F32X2)
__device__ __forceinline__ float4 operator+(float4 a, float4 b) {
    const float2 lo = __fadd2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y));
    const float2 hi = __fadd2_rn(make_float2(a.z, a.w), make_float2(b.z, b.w));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ float4 operator-(float4 a, float4 b) {
    const float2 lo = __fadd2_rn(make_float2(a.x, a.y), make_float2(-b.x, -b.y));
    const float2 hi = __fadd2_rn(make_float2(a.z, a.w), make_float2(-b.z, -b.w));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ float4 operator*(float s, float4 a) {
    const float2 ss = make_float2(s, s);
    const float2 lo = __fmul2_rn(ss, make_float2(a.x, a.y)), hi = __fmul2_rn(ss, make_float2(a.z, a.w));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ float4 fma4(float s, float4 a, float4 acc) {
    const float2 ss = make_float2(s, s);
    const float2 lo = __ffma2_rn(ss, make_float2(a.x, a.y), make_float2(acc.x, acc.y));
    const float2 hi = __ffma2_rn(ss, make_float2(a.z, a.w), make_float2(acc.z, acc.w));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}
#else
__device__ __forceinline__ float4 operator+(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 operator-(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
__device__ __forceinline__ float4 operator*(float s, float4 a) { return make_float4(s * a.x, s * a.y, s * a.z, s * a.w); }
__device__ __forceinline__ float4 fma4(float s, float4 a, float4 acc) {
    return make_float4(fmaf(s, a.x, acc.x), fmaf(s, a.y, acc.y), fmaf(s, a.z, acc.z), fmaf(s, a.w, acc.w));
}
#endif

// ---- kernel argument structures ----------------------------------------------------------------
struct ClipView {
    const void *data;
    long long s[5];  // element strides B, C, F, H, W
    int frame0, n_frames;
    int ring;  // > 0: frames live in a ring of that many slots (frame f at slot f % ring); 0: linear view
};
__device__ __forceinline__ int frame_slot(const ClipView &cv, int f) { return cv.ring > 0 ? f % cv.ring : f - cv.frame0; }

struct YuvDev {   // planar YUV description (see cvvdp_b200_yuv); chroma == 0: not YUV
    int chroma;
    int W, H;
    float yw, yo, cw, co;          // limited-range unpack: Y' = clip(yw*Y - yo, 0, 1), C' = clip(cw*C - co, -.5, .5)
    float m_rv, m_gu, m_gv, m_bu;  // YCbCr -> RGB
};

struct DisplayDev {
    int eotf;
    float gamma;  // EOTF_GAMMA exponent, or HLG system gamma
    float Ypeak, Yblack, Yrefl, exposure;
    float lin_lo;  // max(0.005, Yblack) (display_model.py:349)
    float M[9];    // RGB -> DKLd65, fp32, row-major
};

}  // namespace cvvdp
