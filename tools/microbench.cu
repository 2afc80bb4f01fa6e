// microbench.cu -- issue / pipe model of one B200 SM sub-partition for the instruction mixes of the band and
// temporal kernels: packed FFMA2 against scalar FFMA, and what can issue in the shadow of a packed instruction
// (ALU, MUFU, LDS.128).  One CTA per SM, W warps per scheduler; reports cycles per loop iteration per scheduler.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/microbench tools/microbench.cu && /tmp/microbench
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 2048

// packed operands as 64-bit registers so that the compiler has no reason to shuffle register pairs around
__device__ __forceinline__ void ffma2(unsigned long long &a, unsigned long long b, unsigned long long c) {
    asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a) : "l"(b), "l"(c));
}
__device__ __forceinline__ void fadd2(unsigned long long &a, unsigned long long c) {
    asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a) : "l"(c));
}
__device__ __forceinline__ unsigned long long pack(float x, float y) {
    return ((unsigned long long)__float_as_uint(y) << 32) | __float_as_uint(x);
}

// mode 0: 8 FFMA2 (independent chains)            mode 1: 16 FFMA (independent chains, same flops)
// mode 2: 8 FFMA2 + 8 LOP3/IADD3                  mode 3: 8 FFMA2 + 2 MUFU.EX2
// mode 4: 8 FFMA2 + 2 LDS.128                     mode 5: 8 FFMA2 + 8 FMNMX (alu pipe)
// mode 6: 16 FFMA + 8 LOP3/IADD3                  mode 7: 8 FFMA2 + 8 ALU + 2 MUFU + 2 LDS.128
// mode 8: 8 FADD2                                 mode 9: 2 MUFU only      mode 10: 2 LDS.128 only   mode 11: 8 ALU only
template <int MODE>
__global__ void __launch_bounds__(512, 1) k_mix(float *out, long long *cycles, float seed) {
    __shared__ float4 sm[1024];
    const int tid = threadIdx.x;
    sm[tid] = make_float4(seed, seed, seed, seed);
    sm[tid + 512] = make_float4(seed, seed, seed, seed);
    __syncthreads();
    unsigned long long a[8];
    float s[16];
    unsigned u[8];
    float m0 = seed + tid * 1e-3f, m1 = seed - tid * 1e-3f;
    float4 l0 = make_float4(0, 0, 0, 0), l1 = l0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a[i] = pack(seed + i, seed - i);
        u[i] = tid * 7 + i;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) s[i] = seed + i;
    const unsigned long long b = pack(1.0001f, 0.9999f), c = pack(seed, -seed);
    const float bx = 1.0001f, cx = seed;
    int idx = tid;
    unsigned lacc = 0u;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 0 || MODE == 2 || MODE == 3 || MODE == 4 || MODE == 5 || MODE == 7) {
#pragma unroll
            for (int i = 0; i < 8; ++i) ffma2(a[i], b, c);
        }
        if (MODE == 8) {
#pragma unroll
            for (int i = 0; i < 8; ++i) fadd2(a[i], c);
        }
        if (MODE == 1 || MODE == 6) {
#pragma unroll
            for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(s[i]) : "f"(bx), "f"(cx));
        }
        if (MODE == 2 || MODE == 6 || MODE == 7 || MODE == 11) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {  // 8 x (SHF + LOP3): two alu-pipe instructions each
                unsigned t;
                asm volatile("shr.u32 %0, %1, 3;" : "=r"(t) : "r"(u[i]));
                asm volatile("xor.b32 %0, %0, %1;" : "+r"(u[i]) : "r"(t));
            }
        }
        if (MODE == 5) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("min.f32 %0, %0, %1;" : "+f"(s[i]) : "f"(s[8 + i]));
        }
        if (MODE == 3 || MODE == 7 || MODE == 9) {
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(m0));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(m1));
        }
        if (MODE == 4 || MODE == 7 || MODE == 10) {
            const unsigned sa = (unsigned)__cvta_generic_to_shared(&sm[idx & 511]);
            float4 x, y;
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(sa) : "memory");
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+8192];" : "=f"(y.x), "=f"(y.y), "=f"(y.z), "=f"(y.w) : "r"(sa) : "memory");
            lacc ^= __float_as_uint(x.x) ^ __float_as_uint(y.w);  // one LOP3 consumes both loads
            idx += 32;
        }
    }
    const long long t1 = clock64();
    float r = m0 + m1 + (float)lacc + l0.x + l1.x;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += __uint_as_float((unsigned)a[i]) + __uint_as_float((unsigned)(a[i] >> 32)) + (float)u[i];
#pragma unroll
    for (int i = 0; i < 16; ++i) r += s[i];
    out[blockIdx.x * blockDim.x + tid] = r;
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, float *out, long long *cyc) {
    for (int wps = 1; wps <= 4; ++wps) {  // warps per scheduler
        const int threads = wps * 4 * 32;
        k_mix<MODE><<<148, threads>>>(out, cyc, 1.0f);
        cudaDeviceSynchronize();
        k_mix<MODE><<<148, threads>>>(out, cyc, 1.0f);
        cudaDeviceSynchronize();
        long long h[148];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0; i < 148; ++i) avg += (double)h[i];
        avg /= 148.0;
        printf("%-44s warps/sched %d: %7.2f cycles per iteration per scheduler (%6.2f per warp-iteration)\n", name, wps,
               avg / ITERS, avg / ITERS / wps);
    }
}

int main() {
    float *out;
    long long *cyc;
    cudaMalloc(&out, 148 * 512 * sizeof(float));
    cudaMalloc(&cyc, 148 * sizeof(long long));
    run<0>("8 FFMA2", out, cyc);
    run<8>("8 FADD2", out, cyc);
    run<1>("16 FFMA", out, cyc);
    run<11>("8x(SHF+LOP3)", out, cyc);
    run<9>("2 MUFU.EX2", out, cyc);
    run<10>("2 LDS.128", out, cyc);
    run<2>("8 FFMA2 + 8x(SHF+LOP3)", out, cyc);
    run<6>("16 FFMA + 8x(SHF+LOP3)", out, cyc);
    run<5>("8 FFMA2 + 8 FMNMX", out, cyc);
    run<3>("8 FFMA2 + 2 MUFU.EX2", out, cyc);
    run<4>("8 FFMA2 + 2 LDS.128", out, cyc);
    run<7>("8 FFMA2 + 16 ALU + 2 MUFU + 2 LDS.128", out, cyc);
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
