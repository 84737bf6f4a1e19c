// Bring-up scaffold: built only with `make BRINGUP=1` (not part of the shipped library or of include/spectral_b200.h).
// Self-test of the tcgen05 building blocks (descriptor conventions), used by tests/ and when
// bringing up a new operand layout.  D[128, N] = A[128, K] * B[N, K]^T with selectable smem layouts.
#include "common.cuh"
#include "tc_common.cuh"

__global__ void __launch_bounds__(128, 1)
tc_selftest_kernel(const float* __restrict__ Ag, const float* __restrict__ Bg, float* __restrict__ D, int N, int K,
                   int a_layout_in, int use_mask, uint32_t* __restrict__ info) {
    const int a_layout = a_layout_in & 15, b_sw32 = a_layout_in >> 4;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment by OFFSET from the __shared__ array, so every derived pointer keeps the shared address
    // space (a uintptr_t round-trip turns all later accesses into generic LD/ST through L1TEX)
    uint8_t* base = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t a_bytes = (uint32_t)((K + 31) / 32) * 128u * 128u;   // K-major rows are 128 B wide even when K < 32
    uint8_t* As = base;
    uint8_t* Bs = As + ((a_bytes + 1023) & ~1023u);
    const uint32_t b_chunk = (uint32_t)N * 128;
    uint8_t* tail = Bs + (((uint32_t)((K + 31) / 32) * b_chunk + 1023) & ~1023u);
    uint64_t* mma_bar = reinterpret_cast<uint64_t*>(tail);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int idx = tid; idx < 128 * K; idx += 128) {
        const int m = idx / K, k = idx % K;
        const float v = Ag[idx];
        uint32_t off;
        if (a_layout == 0) off = (uint32_t)(k >> 5) * (128u * 128u) + tc::sw128_kmajor_off(m, k & 31);
        else if (a_layout == 3) off = tc::sw32_kmajor_off(m, k, 128u * 32u);
        else off = tc::sw128b32_mnmajor_off(m, k, (uint32_t)K * 128u);
        *reinterpret_cast<float*>(As + off) = v;
    }
    for (int idx = tid; idx < N * K; idx += 128) {
        const int n = idx / K, k = idx % K;
        if (b_sw32) *reinterpret_cast<float*>(Bs + tc::sw32_kmajor_off(n, k, (uint32_t)N * 32u)) = Bg[idx];
        else *reinterpret_cast<float*>(Bs + (uint32_t)(k >> 5) * b_chunk + tc::sw128_kmajor_off(n, k & 31)) = Bg[idx];
    }
    if (tid == 0) {
        tc::mbar_init(mma_bar, 1);
        tc::fence_barrier_init();
    }
    uint32_t cols = 32;
    while (cols < (uint32_t)N) cols <<= 1;
    if (warp == 0) {
        __syncwarp();
        tc::tmem_alloc(tmem_slot, cols);
        tc::tmem_relinquish();
    }
    tc::fence_proxy_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem_d = *tmem_slot;
    if (tid == 0) {
        info[0] = tmem_d;
        const uint32_t idesc = tc::make_idesc_tf32(128, N, (a_layout == 1 || a_layout == 2), 0);
        info[1] = idesc;
        for (int ks = 0; ks < K / 8; ++ks) {
            uint64_t ad;
            if (a_layout == 0)
                ad = tc::make_smem_desc(tc::smem_u32(As) + (uint32_t)(ks >> 2) * (128u * 128u) + (uint32_t)(ks & 3) * 32, 16, 1024, tc::LAYOUT_SW128);
            else if (a_layout == 3)
                ad = tc::make_smem_desc(tc::smem_u32(As) + (uint32_t)ks * (128u * 32u), 16, 256, tc::LAYOUT_SW32);
            else if (a_layout == 1)
                ad = tc::make_smem_desc(tc::smem_u32(As) + (uint32_t)ks * 1024, (uint32_t)K * 128, 512, tc::LAYOUT_SW128_BASE32B);
            else
                ad = tc::make_smem_desc(tc::smem_u32(As) + (uint32_t)ks * 1024, 512, (uint32_t)K * 128, tc::LAYOUT_SW128_BASE32B);
            const uint64_t bd = b_sw32 ? tc::make_smem_desc(tc::smem_u32(Bs) + (uint32_t)ks * ((uint32_t)N * 32u), 16, 256, tc::LAYOUT_SW32)
                                       : tc::make_smem_desc(tc::smem_u32(Bs) + (uint32_t)(ks >> 2) * b_chunk + (uint32_t)(ks & 3) * 32, 16, 1024, tc::LAYOUT_SW128);
            if (ks == 0) { info[2] = (uint32_t)ad; info[3] = (uint32_t)(ad >> 32); info[4] = (uint32_t)bd; info[5] = (uint32_t)(bd >> 32); }
            if (use_mask) {
                uint32_t z = 0, acc = ks > 0;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                             ::"r"(tmem_d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc), "r"(z) : "memory");
            } else {
                tc::umma_tf32(tmem_d, ad, bd, idesc, ks > 0 ? 1u : 0u);
            }
        }
        tc::umma_commit(mma_bar);
    }
    tc::mbar_wait(mma_bar, 0);
    tc::tc_fence_after_sync();
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t r[32];
        tc::tmem_ld_32x32b_x32(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (c0 + j < N) D[(warp * 32 + lane) * N + c0 + j] = __uint_as_float(r[j]);
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_d, cols);
}

extern "C" int sb200_tc_selftest(const float* A, const float* B, float* D, int N, int K, int a_layout, int use_mask,
                                 uint32_t* info, void* stream) {
    SB_REQUIRE(A && B && D && info, "tc_selftest: NULL argument");
    SB_REQUIRE(N % 16 == 0 && N >= 16 && N <= 256 && K % 8 == 0 && K >= 8 && K <= 128, "tc_selftest: bad N/K");
    const size_t smem = 1024 + (size_t)((K + 31) / 32) * 128 * 128 +
                        (((size_t)((K + 31) / 32) * N * 128 + 1023) & ~(size_t)1023) + 64;
    SB_CHECK_CUDA(cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sb_launch(tc_selftest_kernel, 1, 128, smem, (cudaStream_t)stream, A, B, D, N, K, a_layout, use_mask, info);
    SB_LAUNCH_CHECK();
    return 0;
}
