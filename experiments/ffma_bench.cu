// Issue rate of FFMA (3-register) against the packed FFMA2 on one SM sub-partition: W warps per scheduler, 16 independent
// accumulators per thread (no dependency stalls).  Prints cycles per warp-instruction per scheduler.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
template <int MODE>
__global__ void k(float* out, int reps, long long* cyc, float s) {
    float a[16];
    unsigned long long p[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) { a[j] = threadIdx.x * 0.001f + j; p[j] = (unsigned long long)__float_as_uint(a[j]) * 0x100000001ull; }
    const float m = s, c = 1e-3f;
    const unsigned long long m2 = (unsigned long long)__float_as_uint(m) * 0x100000001ull, c2 = (unsigned long long)__float_as_uint(c) * 0x100000001ull;
    __syncthreads();
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (MODE == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[j]) : "f"(m), "f"(c));
            else p[j] = fma2(p[j], m2, c2);
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    float acc = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) acc += a[j] + __uint_as_float((unsigned)p[j]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    const int reps = 2000;
    for (int warps = 4; warps <= 32; warps *= 2) {
        for (int mode = 0; mode < 2; ++mode) {
            if (mode == 0) k<0><<<148, warps * 32>>>(out, reps, cyc, 0.999f); else k<1><<<148, warps * 32>>>(out, reps, cyc, 0.999f);
            cudaDeviceSynchronize();
            long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
            const double per = avg / ((double)reps * 16 * (warps / 4.0));
            printf("%2d warps/SM (%d per scheduler)  %-6s %5.2f cycles per warp-instruction per scheduler  = %5.1f FMA/clk/SM\n", warps, warps / 4,
                   mode ? "FFMA2" : "FFMA", per, (mode ? 64.0 : 32.0) / per * 4);
        }
    }
    return 0;
}
