// Streaming-rate microbenchmark for the load paths the kernels can use (B200, sm_100a):
//   mode 0: TMA 2-D boxes {inner px, R rows} from a [rows, HW] fp32 matrix, 128B swizzle (inner = 32)
//   mode 1: same, no swizzle, inner = 32..256 px
//   mode 2: 1-D bulk copies of contiguous chunks
//   mode 3: plain LDG.128 streaming (each thread sums what it loads), no smem
// Each CTA runs a ring of S stages; one thread issues, one thread recycles (no data touch), so the number
// measured is what the copy engine + memory system deliver.   usage: membench mode inner rows stages ctas_per_sm
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include "../dlwp_benchmark_b200/csrc/tc_common.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

struct P { int mode, inner, rows, S; long HW, nrows; long items_per_row_block, nitems; const float* src; float* sink; };

__global__ void __launch_bounds__(128) k_tma(const __grid_constant__ CUtensorMap tm, const P p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t item_bytes = (uint32_t)p.inner * p.rows * 4;
    uint64_t* full = (uint64_t*)(base + (size_t)p.S * item_bytes);
    uint64_t* empty = full + p.S;
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.S; ++s) { tc::mbar_init(full + s, 1); tc::mbar_init(empty + s, 1); }
        tc::fence_barrier_init();
    }
    __syncthreads();
    const long per = (p.nitems + gridDim.x - 1) / gridDim.x;
    const long i0 = blockIdx.x * per, i1 = min(p.nitems, i0 + per);
    if (threadIdx.x == 0) {
        uint32_t s = 0, ph = 0;
        for (long it = i0; it < i1; ++it) {
            tc::mbar_wait(empty + s, ph ^ 1);
            tc::mbar_expect_tx(full + s, item_bytes);
            const long rb = it / p.items_per_row_block, c = it % p.items_per_row_block;
            if (p.mode == 2) {
                tc::bulk_load_1d(base + (size_t)s * item_bytes, (const uint8_t*)p.src + (size_t)it * item_bytes, item_bytes, full + s);
            } else {
                tc::tma_load_2d(base + (size_t)s * item_bytes, &tm, (int)(c * p.inner), (int)(rb * p.rows), full + s);
            }
            if (++s == (uint32_t)p.S) { s = 0; ph ^= 1; }
        }
    } else if (threadIdx.x == 32) {
        uint32_t s = 0, ph = 0;
        for (long it = i0; it < i1; ++it) {
            tc::mbar_wait(full + s, ph);
            tc::mbar_arrive(empty + s);
            if (++s == (uint32_t)p.S) { s = 0; ph ^= 1; }
        }
    }
}

__global__ void __launch_bounds__(256) k_ldg(const float4* __restrict__ src, float* sink, long n4) {
    float acc = 0.f;
    const long stride = (long)gridDim.x * blockDim.x;
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n4; i += 4 * stride) {
        float4 a = __ldg(src + i), b = __ldg(src + i + stride), c = __ldg(src + i + 2 * stride), d = __ldg(src + i + 3 * stride);
        acc += a.x + b.y + c.z + d.w;
    }
    for (; i < n4; i += stride) acc += __ldg(src + i).x;
    if (acc == 1.2345f) *sink = acc;
}

int main(int argc, char** argv) {
    P p;
    p.mode = argc > 1 ? atoi(argv[1]) : 0;
    p.inner = argc > 2 ? atoi(argv[2]) : 32;
    p.rows = argc > 3 ? atoi(argv[3]) : 64;
    p.S = argc > 4 ? atoi(argv[4]) : 4;
    const int cps = argc > 5 ? atoi(argv[5]) : 1;
    p.HW = 4096; p.nrows = 64L * 64 * 8;                       // 8 x [64*64 rows, 4096 px] = 537 MB
    const size_t bytes = (size_t)p.HW * p.nrows * 4;
    float* d; float* sink;
    CK(cudaMalloc(&d, bytes)); CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(d, 0, bytes));
    p.src = d; p.sink = sink;
    CUtensorMap tm; memset(&tm, 0, sizeof(tm));
    const uint32_t item_bytes = (uint32_t)p.inner * p.rows * 4;
    if (p.mode <= 1) {
        if (sb200_make_tmap_2d_f32(&tm, d, p.HW, p.nrows, p.HW * 4, p.inner, p.rows, p.mode == 0 ? 1 : 0)) { printf("tmap failed\n"); return 1; }
        p.items_per_row_block = p.HW / p.inner;
        p.nitems = p.items_per_row_block * (p.nrows / p.rows);
    } else {
        p.items_per_row_block = 1;
        p.nitems = bytes / item_bytes;
    }
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0));
        if (p.mode == 3) {
            k_ldg<<<148 * 8, 256>>>((const float4*)d, sink, (long)(bytes / 16));
        } else {
            const size_t smem = 1024 + (size_t)p.S * item_bytes + 16 * p.S + 64;
            CK(cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_tma<<<148 * cps, 128, smem>>>(tm, p);
        }
        CK(cudaGetLastError());
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    printf("mode %d inner %3d rows %3d S %2d cps %d item %6u B : %8.1f GB/s (%.3f ms)\n", p.mode, p.inner, p.rows, p.S, cps,
           item_bytes, bytes / best * 1e-6, best);
    return 0;
}
