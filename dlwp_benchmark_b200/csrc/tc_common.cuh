// Blackwell (sm_100a) primitives used by the tensor-core kernels: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 (TMEM alloc / mma / commit / ld) and the shared-memory / instruction descriptors.
// Descriptor bit layouts follow the PTX ISA "tcgen05 matrix descriptors" (same fields CUTLASS's
// cute/arch/mma_sm100_desc.hpp documents); nothing here depends on CUTLASS.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------- mbarrier ----------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// generic-proxy writes to shared memory -> visible to the async proxy (TMA / tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
#ifdef SB200_TEST_WAIT
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
#else
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
#endif
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps (launch failure) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) __trap();
    }
}

// Whole-warp wait: ONE lane polls the barrier, the others park at the warp barrier (fewer try_wait requests on the
// shared-memory pipe; measured neutral on the kernels of this library, kept because it is never worse).
// __syncwarp() orders memory among the participating lanes, so data published before the barrier completed is
// visible to all of them afterwards.
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity) {
    if ((threadIdx.x & 31) == 0) mbar_wait(bar, parity);
    __syncwarp();
}

// ---------------- TMA ----------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// plain 1-D bulk copy global -> shared (no tensor map)
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ---------------- TMEM ----------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {   // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {      // same warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], kind::tf32, issued by ONE thread
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier once all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns: thread i of the warp gets lane (base_lane + i), columns c..c+31
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------- descriptors ----------------
// shared-memory matrix descriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48),
// base_offset [49,52), lbo_mode [52], layout_type [61,64) (0 none, 2 = 128B swizzle, 4 = 64B, 6 = 32B)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(layout & 7) << 61;
    return d;
}
// The same descriptor as two 32-bit halves, so that an issue loop only has to add to the low word:
//   lo = start>>4 | (LBO>>4)<<16        hi = SBO>>4 | version(1)<<14 | layout<<29
__host__ __device__ __forceinline__ constexpr uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
    return ((saddr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
__host__ __device__ __forceinline__ constexpr uint32_t desc_hi(uint32_t sbo_bytes, uint32_t layout) {
    return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | ((layout & 7u) << 29);
}
// One lane of a converged warp (elect.sync).  Guarding the single-thread roles (TMA producer, MMA issuer) with this
// instead of `lane == 0` lets ptxas see that exactly one thread is active, so instructions that take uniform-register
// operands (UTCHMMA, UTMALDG, ..) are emitted once instead of inside an ELECT / BRA.U.ANY serialisation loop each.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
    return pred != 0;
}
// D[tmem] (+)= A * B with descriptors given as (lo, hi) halves
__device__ __forceinline__ void umma_tf32_lh(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                             uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
constexpr uint32_t LAYOUT_NONE = 0, LAYOUT_SW128_BASE32B = 1, LAYOUT_SW128 = 2, LAYOUT_SW64 = 4, LAYOUT_SW32 = 6;

// instruction descriptor, kind::tf32, fp32 accumulate: c_format=F32 [4,6), a/b format TF32(2) [7,10)/[10,13),
// a_major [15], b_major [16] (0 = K-major, 1 = MN-major), N>>3 [17,23), M>>4 [24,29)
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// byte offset of element (row, k) inside a K-major 128B-swizzled tile whose rows hold 32 fp32 (128 B):
// 8-row atoms of 1024 B; 16-byte chunk index XOR (row % 8).  Tile base must be 1024-byte aligned.
__device__ __forceinline__ uint32_t sw128_kmajor_off(int row, int k /* 0..31 */) {
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((k >> 2) ^ (row & 7)) & 7) << 4) + ((k & 3) << 2));
}

// byte offset of element (row, k) inside a K-major 32B-swizzled tile: one tile per k-step of 8 fp32
// (rows of 32 B, 8-row atoms of 256 B, 16-byte chunk index XOR ((row >> 2) & 1)); consecutive k-steps
// are `kstep_stride` bytes apart.  Descriptor: LAYOUT_SW32, SBO = 256.  Tile base 256-byte aligned.
__device__ __forceinline__ uint32_t sw32_kmajor_off(int row, int k, uint32_t kstep_stride) {
    return (uint32_t)(k >> 3) * kstep_stride + (uint32_t)row * 32u + (uint32_t)(((((k & 7) >> 2) ^ ((row >> 2) & 1)) & 1) << 4) +
           (uint32_t)((k & 3) << 2);
}

// byte offset of element (mn, k) inside an MN-major tile of 32-bit elements in the SWIZZLE_128B_BASE32B layout
// (the only MN-major layout tcgen05 accepts for tf32; TMA: CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B):
// rows = k (128 B = 32 mn elements each), 4-row atoms of 512 B, 32-byte chunk index XOR (k % 4);
// 32-element mn blocks are `mn_block_stride` bytes apart.  Tile base must be 512-byte aligned.
__device__ __forceinline__ uint32_t sw128b32_mnmajor_off(int mn, int k, uint32_t mn_block_stride) {
    return (uint32_t)(mn >> 5) * mn_block_stride + (uint32_t)k * 128u + (uint32_t)(((((mn & 31) >> 3) ^ (k & 3)) & 3) << 5) +
           (uint32_t)((mn & 7) << 2);
}

// 3xTF32 operand split  v = hi + lo:  hi = v rounded to NEAREST tf32 (10 explicit mantissa bits), lo = (v - hi) rounded
// to nearest tf32 (v - hi is exact in fp32).  The tensor core reads a 32-bit operand as tf32 by DROPPING its low 13 bits,
// so whatever is not rounded here is truncated there.  Measured in round 2 on whole models (rel-L2 of the cfg3 / cfg5
// outputs against the fp64 oracle; the FFMA path gives 2e-6, plain fp32 torch 1.3e-6):
//     hi truncated, lo truncated by the hardware   1.4e-5 / 1.8e-5   (every product biased by about -2^-21: lo >= 0 always)
//     hi rounded,   lo truncated by the hardware   > 1e-5 at cfg3     (truncation always shrinks |lo|: still a bias)
//     hi rounded,   lo rounded                      7.8e-6            (symmetric errors, <= 2^-23 relative per product)
// Integer rounding (add half an ulp of tf32, clear 13 bits: round-half-away) runs at full ALU rate; cvt.rna.tf32.f32
// is a conversion-pipe instruction and measured 4 % slower on the cfg2 step.
__device__ __forceinline__ float tf32_rna(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u); }
__device__ __forceinline__ float tf32_trunc(float v) { return tf32_rna(v); }          // "hi" part (historic name)
__device__ __forceinline__ float tf32_cut(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }   // what the tensor core reads
__device__ __forceinline__ float tf32_lo(float v, float hi) { return tf32_rna(v - hi); }
__device__ __forceinline__ float4 tf32_lo4(float4 v, float4 h) {
    return make_float4(tf32_lo(v.x, h.x), tf32_lo(v.y, h.y), tf32_lo(v.z, h.z), tf32_lo(v.w, h.w));
}

}  // namespace tc

// host side: build a 2-D tiled tensor map (fp32) through the driver entry point (no -lcuda needed)
int sb200_make_tmap_2d_f32(CUtensorMap* out, const void* base, uint64_t dim0, uint64_t dim1, uint64_t stride1_bytes,
                           uint32_t box0, uint32_t box1, int swizzle /* 0 none, 1 = 128B, 2 = 128B with 32B atoms */);
