// tcgen05 (5th-gen tensor core) version of the fused [row synthesis +] pointwise channel mix + epilogue.
//
//   D[p, n] = sum_m A[b, m, p] * Wp[n, m]                   p = 128 pixels of one sample (UMMA M = 128)
//                                                           n = all output channels        (UMMA N <= 256)
// A tile:  TMA (cp.async.bulk.tensor.2d, 128B swizzle) straight from the NCHW activation viewed as
//          [B*M, H*W]: four boxes {32 px, KC channels} with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B = the MN-major
//          SWIZZLE_128B_BASE32B operand layout (the only MN-major layout tcgen05 accepts for 32-bit types).
// B tile:  the (tiny) weight matrix, written by the CTA's threads in the K-major SWIZZLE_128B layout.
// D:       fp32 accumulator in TMEM; epilogue reads it with tcgen05.ld (lane = pixel), adds bias,
//          applies GELU / GELU' and stores coalesced along the pixel dimension.
//
// Precision: kind::tf32 keeps 10 mantissa bits, so the fp32 parity path splits both operands
// (x = hi + lo, hi = the tf32-representable part) and issues three MMAs per k-step
// (hi*hi + lo*hi + hi*lo): error ~2^-21, inside the 1e-5 parity bar.  PASSES = 1 is the plain TF32
// path (parity ~1e-3, the north_star's "tensor-core path" tolerance).
#include <mutex>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

// ---------------------------------------------------------------------------------------------
// host helpers
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;
static std::once_flag g_encode_once;

static void load_encode() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
        g_encode = (PFN_encodeTiled)fn;
    else
        cudaGetLastError();
}

int sb200_make_tmap_2d_f32(CUtensorMap* out, const void* base, uint64_t dim0, uint64_t dim1, uint64_t stride1_bytes,
                           uint32_t box0, uint32_t box1, int swizzle) {
    std::call_once(g_encode_once, load_encode);
    SB_REQUIRE(g_encode != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    {
        // cuTensorMapEncodeTiled is a DRIVER call and needs a current context.  A thread that has made no runtime call
        // yet (the autograd engine's worker when torch's caching allocator served every allocation) has none:
        // cudaSetDevice on the current device binds the primary context (CUDA 12) and is legal during graph capture.
        int dev = 0;
        SB_CHECK_CUDA(cudaGetDevice(&dev));
        SB_CHECK_CUDA(cudaSetDevice(dev));
    }
    cuuint64_t gdim[2] = {dim0, dim1};
    cuuint64_t gstr[1] = {stride1_bytes};
    cuuint32_t box[2] = {box0, box1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE,
                          swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_128B
                                       : (swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_NONE),
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with %d (dims %llu x %llu, box %u x %u)", (int)r,
               (unsigned long long)dim0, (unsigned long long)dim1, box0, box1);
    return 0;
}

#define g_tc_mode (sb_tc_mode())

// ---------------------------------------------------------------------------------------------
// kernel: persistent, warp-specialised
//   warp 0      TMA producer (one elected lane)  activation K-chunks (32 channels x 128 px) -> smem ring
//   warp 1      MMA issuer (one elected lane)    tcgen05.mma into a double-buffered TMEM accumulator
//   warps 4-7   stagers: split the landed activation chunks of the coming tiles into tf32 hi / lo parts, in their own loop
//   warps 8-23  drainers: stage the Phi operand of the NEXT tile (spectral term; or generate the A operand, ASRC == 1),
//               then the epilogue of the CURRENT tile: TMEM -> registers -> bias / GELU / GELU' -> global
// so the MMAs and the split of tile t+1 execute while tile t is being written out; TMA runs ahead by the ring depth.
// (Until round 2 the same 16 warps did split and epilogue one after the other; see "Warp roles" below.)
//
// Spectral term (row synthesis):  D[p, n] += sum_kk E[p, kk] * Phi[b, n, y(p), kk]
//   E   : constant [128 px][K2pad] operand (plan table; block-diagonal over the R image rows of a tile),
//         K-major SWIZZLE_32B, resident in smem as hi/lo
//   Phi : per tile [N][K2pad] (column-synthesised spectrum of the R rows of this tile, rotated to the
//         tile's x offset), staged by the workers as K-major SWIZZLE_32B hi/lo, double-buffered
// ---------------------------------------------------------------------------------------------
constexpr int TP_PX = 128;
// Warp roles (warpgroup aligned, registers re-divided with setmaxnreg after the common setup):
//   warps 0-3   TMA producer, MMA issuer, two idle warps                      (40 registers)
//   warps 4-7   stagers: split the TMA-fed A tiles of the coming tiles        (56 registers)
//   warps 8-23  drainers: epilogue, 4 TMEM lane quarters x 4 column parts     (96 registers)
// With an on-chip generated A operand (ASRC == 1: one GELU per element) the drainers stage as well and the stagers idle.
constexpr int TP_STAGE_WARPS = 4;
constexpr int TP_DRAIN_WARPS = 16;
constexpr int TP_THREADS = 32 * (4 + TP_STAGE_WARPS + TP_DRAIN_WARPS);
constexpr int TP_WTHREADS = 32 * TP_STAGE_WARPS;   // threads that stage (ASRC == 0)
// setmaxnreg only moves registers inside the CTA's own allocation (768 threads x 80 at launch = 61440):
// 128 x 40 + 128 x 56 + 512 x 96 = 61440.  Asking for more than the decs release blocks forever.
constexpr int TP_REGS_CTRL = 40, TP_REGS_STAGE = 56, TP_REGS_DRAIN = 96;
static_assert(128 * TP_REGS_CTRL + 32 * TP_STAGE_WARPS * TP_REGS_STAGE + 32 * TP_DRAIN_WARPS * TP_REGS_DRAIN <= TP_THREADS * 80,
              "register budget of the warp roles");

#ifdef SB200_BRINGUP
#define TP_TRACE(role, i, ev) do { if (p.trace && blockIdx.x == 0 && (i) < 32) p.trace[((role) * 32 + (i)) * 4 + (ev)] = clock64(); } while (0)
#else
#define TP_TRACE(role, i, ev) do { } while (0)
#endif
struct TcPwParams {
    const float* Wp; int64_t w_sn, w_sm;
    const float* bias; const float* zprev;
    float* z_out; float* y_out;
    int B, M, N, KC, nkc, stages;
    int64_t HW;
    int64_t ntiles;
    int mode, apply_act;
    uint32_t idesc, idesc_spec, tmem_cols;
    int tpr_log2;            // log2(drainer threads cooperating on one channel while staging Phi)
    int nlo;                 // 3-pass, TMA-fed A: the lo halves of the split live in their own nlo-slot ring behind the raw
                             // slots (0: every stage is [hi | lo]); a raw slot then costs KC*512 bytes instead of twice that,
                             // so twice as many loads are in flight for the same shared memory
    long long* trace;        // bring-up builds only (SB200_TP_TRACE): clock64 event log of CTA 0
    int dbg;                 // bring-up builds only (SB200_TP_DBG): 1 GELU -> identity, 2 epilogue drains nothing, 4 no MMAs
    int corr;                // 3-pass mode, 4N <= 512: the lo*hi + hi*lo corrections accumulate in their own TMEM columns (behind the
                             // N main columns of each buffer), so the large hi*hi terms go through a chain of K/8 truncating
                             // tensor-core accumulations instead of 3K/8; the epilogue adds the two
    // spectral term (Phi == NULL: none)
    const float2* Phi; const float* E; const float2* rot;
    int H, W, Mx, R, V, K2, K2pad;
    int bias_mma;            // bias rides in the spare K column of the synthesis operands (E row K2 = 1, Phi col K2 = bias)
    // fused MLP head (EPI 5 / 6): second (1-output) layer folded into the epilogue
    const float* w2;         // [N] weights of the output channel
    const float* b2;         // [1] or NULL
    const float* gy;         // EPI 6: [B, HW] gradient of the head output
    float* colsum_ws;        // EPI 6: [grid*4][2][256] per-warp-row partial column sums (gb1 | gw2) + [grid*4] partial sum(gy)
};

// EPI: 0 fwd linear | 1 fwd GELU, also write pre-activation z | 2 fwd GELU | 3 bwd * GELU'(zprev) | 4 bwd plain
//      5 fused head forward:  y[b,p] = sum_n w2[n] GELU(D[p,n] + b1[n]) + b2     (the N-channel hidden never leaves the SM)
//      6 fused head backward: recomputed D = z1;  gz1[b,n,p] = w2[n] gy[b,p] GELU'(z1) is written (y_out) and the pixel
//        reductions gb1[n] = sum gz1, gw2[n] = sum gy GELU(z1), gb2 = sum gy leave through colsum_ws (N == 256)
//      7 fused lifting-tail backward (1 input channel):  gz1[n,p] = D[p,n] GELU'(w1[n] x[p] + b1[n]) is never stored;
//        only gb1[n] = sum_p gz1 and gw1[n] = sum_p gz1 x[p] leave, through colsum_ws (w2 = w1, bias = b1, gy = x)

// sums of 16 per-lane values over the 32 lanes of a warp, transposing while reducing: 16 SHFL instead of 80.
// Returns, in every lane, the total of column ((lane>>4)&1)*8 + ((lane>>3)&1)*4 + ((lane>>2)&1)*2 + ((lane>>1)&1).
__device__ __forceinline__ float warp_colsum16(const float (&v)[16], int lane) {
    float a[8], b[4], c[2];
    const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float keep = h16 ? v[i + 8] : v[i], send = h16 ? v[i] : v[i + 8];
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float keep = h8 ? a[i + 4] : a[i], send = h8 ? a[i] : a[i + 4];
        b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float keep = h4 ? b[i + 2] : b[i], send = h4 ? b[i] : b[i + 2];
        c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    const float keep = h2 ? c[1] : c[0], send = h2 ? c[0] : c[1];
    float d = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    return d;
}
// ASRC: 0 = activation tiles arrive by TMA | 1 = generated on chip: A[m, px] = GELU(w1[m] x[b, px] + b1[m]), the hidden
//       layer of a 1-input-channel lifting MLP (w1 = p.w2, b1 = p.b2, x = p.gy): the 256-channel tensor never exists in HBM
// SPEC: the spectral term exists (p.Phi != NULL); compiled out otherwise so that the drainers of the plain pointwise
//       kernels carry no Phi staging state
template <int PASSES, int EPI, int ASRC = 0, bool SPEC = false>
__global__ void __launch_bounds__(TP_THREADS, 1)
tc_pointwise_kernel(const __grid_constant__ CUtensorMap tmapA, const TcPwParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment by OFFSET from the __shared__ array, so every derived pointer keeps the shared address
    // space (a uintptr_t round-trip turns all later accesses into generic LD/ST through L1TEX)
    uint8_t* base = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    const int KC = p.KC, nkc = p.nkc, S = p.stages;
    const bool spectral = SPEC;
    const uint32_t a_bytes = (uint32_t)KC * 512;                       // 4 boxes x KC rows x 128 B
    const int NLO = (PASSES == 3 && ASRC == 0) ? p.nlo : 0;
    const uint32_t a_stage_bytes = a_bytes * ((PASSES == 3 && NLO == 0) ? 2 : 1);    // [hi | lo], or raw -> hi alone (lo ring)
    const int kchunks = (p.M + 31) / 32;                               // 32-wide K chunks of the resident weight tile
    const uint32_t b_chunk_bytes = (uint32_t)p.N * 128;                // N rows x 128 B
    const uint32_t b_bytes = nkc ? (((uint32_t)kchunks * b_chunk_bytes + 1023) & ~1023u) : 0u;
    const int ksteps2 = spectral ? p.K2pad / 8 : 0;
    const uint32_t e_bytes = (uint32_t)ksteps2 * 4096;                 // [kstep][128 rows][32 B]
    const uint32_t phi_kstep = (uint32_t)p.N * 32;
    const uint32_t phi_bytes = ((uint32_t)ksteps2 * phi_kstep + 1023) & ~1023u;
    uint8_t* B_hi = base;
    uint8_t* B_lo = B_hi + b_bytes;
    uint8_t* E_hi = B_lo + (PASSES == 3 ? b_bytes : 0);
    uint8_t* E_lo = E_hi + e_bytes;
    uint8_t* Phi_s = E_lo + (PASSES == 3 ? e_bytes : 0);              // [2 buffers][hi | lo]
    const uint32_t phi_buf_bytes = phi_bytes * (PASSES == 3 ? 2 : 1);
    uint8_t* A_st = Phi_s + 2 * phi_buf_bytes;
    uint8_t* LO_st = A_st + (uint32_t)S * a_stage_bytes;              // [NLO] lo ring
    uint8_t* tail = LO_st + (uint32_t)NLO * a_bytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);           // [S]  TMA landed
    uint64_t* split_bar = full_bar + S;                                // [S]  hi/lo split done (workers -> MMA)
    uint64_t* empty_bar = split_bar + S;                               // [S]  MMAs that read the stage are done
    uint64_t* tfull_bar = empty_bar + S;                               // [2]  accumulator complete
    uint64_t* tempty_bar = tfull_bar + 2;                              // [2]  accumulator drained by the epilogue
    uint64_t* phi_bar = tempty_bar + 2;                                // [2]  Phi operand staged
    uint64_t* lo_empty = phi_bar + 2;                                  // [NLO <= 4] MMAs that read the lo slot are done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(lo_empty + 4);
    float* bias_s = reinterpret_cast<float*>(tail + 512);              // [256] bias for the epilogue (zeros when unused); barriers use < 512 B
    float2* rot_s = reinterpret_cast<float2*>(bias_s + 256);           // [V][Mx] tile phase table
    float* w2_s = reinterpret_cast<float*>(rot_s + 256);               // [256]       EPI 5 / 6 (launcher adds the bytes)
    float* red_s = w2_s + 256;                                         // [2][4][128] EPI 5 cross-warp partial sums

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t tiles_per_b = (uint32_t)((p.HW + TP_PX - 1) / TP_PX);
    const bool bias_epi = p.bias != nullptr && !p.bias_mma;

    // ---- one-time setup ----
    if (tid == 0) {
        if (nkc) tc::tma_prefetch_desc(&tmapA);
        for (int s = 0; s < S; ++s) {
            tc::mbar_init(full_bar + s, 1);
            tc::mbar_init(split_bar + s, ASRC == 1 ? TP_DRAIN_WARPS : TP_STAGE_WARPS);
            tc::mbar_init(empty_bar + s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            tc::mbar_init(tfull_bar + a, 1);
            tc::mbar_init(tempty_bar + a, TP_DRAIN_WARPS);
            tc::mbar_init(phi_bar + a, TP_DRAIN_WARPS);
        }
        for (int l = 0; l < 4; ++l) tc::mbar_init(lo_empty + l, 1);
        tc::fence_barrier_init();
    }
    if (warp == 0) {
        __syncwarp();
        tc::tmem_alloc(tmem_slot, p.tmem_cols);
        tc::tmem_relinquish();
    }
    if (spectral) {
        // resident synthesis operand E[kk][px] -> K-major 32B-swizzled hi / lo; with bias_mma the spare row K2 is all ones
        // (plan tables are immutable: staged BEFORE the grid-dependency wait, overlapping the previous kernel's tail)
        for (int idx = tid; idx < p.K2pad * 128; idx += TP_THREADS) {
            const int k = idx >> 7, px = idx & 127;
            const float v = (p.bias_mma && k == p.K2) ? 1.0f : __ldg(p.E + idx);
            const float hi = tc::tf32_trunc(v);
            const uint32_t off = tc::sw32_kmajor_off(px, k, 4096u);
            *reinterpret_cast<float*>(E_hi + off) = hi;
            if (PASSES == 3) *reinterpret_cast<float*>(E_lo + off) = tc::tf32_lo(v, hi);
        }
        for (int idx = tid; idx < p.V * p.Mx; idx += TP_THREADS) rot_s[idx] = __ldg(p.rot + idx);
        // zero both Phi buffers once: the K padding columns are never written again
        for (int idx = tid; idx < (int)(2 * phi_buf_bytes / 16); idx += TP_THREADS)
            reinterpret_cast<float4*>(Phi_s)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // resident weight tile: Wp[n, m] -> K-major 128B-swizzled rows, split into tf32 hi / lo
    if (nkc) {
        for (int idx = tid; idx < p.N * p.M; idx += TP_THREADS) {
            const int n = idx / p.M, k = idx % p.M;
            const float w = __ldg(p.Wp + (int64_t)n * p.w_sn + (int64_t)k * p.w_sm);
            const float hi = tc::tf32_trunc(w);
            const uint32_t off = (uint32_t)(k >> 5) * b_chunk_bytes + tc::sw128_kmajor_off(n, k & 31);
            *reinterpret_cast<float*>(B_hi + off) = hi;
            if (PASSES == 3) *reinterpret_cast<float*>(B_lo + off) = tc::tf32_lo(w, hi);
        }
    }
    for (int idx = tid; idx < 256; idx += TP_THREADS) bias_s[idx] = (bias_epi && idx < p.N) ? __ldg(p.bias + idx) : 0.f;
    if (EPI == 5 || EPI == 6 || EPI == 7)
        for (int idx = tid; idx < 256; idx += TP_THREADS) w2_s[idx] = idx < p.N ? __ldg(p.w2 + idx) : 0.f;
    if (ASRC == 1)
        for (int idx = tid; idx < 256; idx += TP_THREADS) {
            w2_s[idx] = idx < p.M ? __ldg(p.w2 + idx) : 0.f;
            red_s[idx] = idx < p.M ? __ldg(p.b2 + idx) : 0.f;
        }
    if (spectral && p.bias_mma) {
        __syncthreads();
        // Phi column K2 carries the bias (constant over tiles) in both buffers
        for (int idx = tid; idx < 2 * p.N; idx += TP_THREADS) {
            const int a = idx >= p.N, n = idx - a * p.N;
            const float v = __ldg(p.bias + n);
            const float hi = tc::tf32_trunc(v);
            const uint32_t off = tc::sw32_kmajor_off(n, p.K2, phi_kstep);
            *reinterpret_cast<float*>(Phi_s + (uint32_t)a * phi_buf_bytes + off) = hi;
            if (PASSES == 3) *reinterpret_cast<float*>(Phi_s + (uint32_t)a * phi_buf_bytes + phi_bytes + off) = tc::tf32_lo(v, hi);
        }
    }
    tc::fence_proxy_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    const uint32_t first = blockIdx.x, stride = gridDim.x, ntiles = (uint32_t)p.ntiles;
    const uint32_t my_tiles = first < ntiles ? (ntiles - first + stride - 1) / stride : 0;

    // Each role branch opens with the setmaxnreg of its warpgroup(s): registers move from the control warps and the
    // stagers to the drainers (every warp of a warpgroup executes the same instruction).
    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TP_REGS_CTRL));
    if (warp == 0) {
        // ================= TMA producer =================
        if (tc::elect_one() && nkc && ASRC == 0) {
            uint32_t s = 0, ph = 0;                                 // ring position / phase
            int gc = 0; (void)gc;
            for (uint32_t it = 0; it < my_tiles; ++it) {
                const uint32_t tile = first + it * stride;
                const uint32_t b = tile / tiles_per_b;
                const int p_base = (int)((tile - b * tiles_per_b) * TP_PX);
                for (int kc = 0; kc < nkc; ++kc) {
                    TP_TRACE(0, gc, 0);
                    tc::mbar_wait(empty_bar + s, ph ^ 1);
                    TP_TRACE(0, gc, 1);
                    uint8_t* dst = A_st + s * a_stage_bytes;
                    tc::mbar_expect_tx(full_bar + s, a_bytes);
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        tc::tma_load_2d(dst + (uint32_t)i * KC * 128, &tmapA, p_base + 32 * i, (int)b * p.M + kc * KC,
                                        full_bar + s);
                    TP_TRACE(0, gc, 2); ++gc;
                    if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        // One thread issues every tcgen05.mma of the CTA, so the loop bodies are kept to a handful of
        // integer instructions: descriptor high words are loop constants, low words advance by adds.
        if (tc::elect_one()) {
            uint32_t s = 0, ph = 0, lo_l = 0;
            const uint32_t a_hi32 = tc::desc_hi(512, tc::LAYOUT_SW128_BASE32B);
            const uint32_t b_hi32 = tc::desc_hi(1024, tc::LAYOUT_SW128);
            const uint32_t s_hi32 = tc::desc_hi(256, tc::LAYOUT_SW32);
            const uint32_t a_lbo = (uint32_t)KC * 128;
            const uint32_t bh_base = tc::desc_lo(tc::smem_u32(B_hi), 16), bl_base = tc::desc_lo(tc::smem_u32(B_lo), 16);
            const uint32_t eh_base = tc::desc_lo(tc::smem_u32(E_hi), 16), el_base = tc::desc_lo(tc::smem_u32(E_lo), 16);
            const uint32_t phi_step = phi_kstep >> 4;
            const int ksteps = KC / 8;
            for (uint32_t it = 0; it < my_tiles; ++it) {
                const uint32_t a = it & 1;
                const uint32_t tround = it >> 1;
                TP_TRACE(3, it, 0);
                tc::mbar_wait(tempty_bar + a, (tround & 1) ^ 1);
                TP_TRACE(3, it, 1);
                tc::tc_fence_after_sync();
                const uint32_t bufw = (uint32_t)p.N * (p.corr ? 2u : 1u);
                const uint32_t tmem_d = tmem_base + (uint32_t)a * bufw;
                const uint32_t tmem_c = p.corr ? tmem_d + (uint32_t)p.N : tmem_d;   // correction accumulator
                uint32_t started = 0, started_c = p.corr ? 0u : 1u;
                if (spectral) {
                    tc::mbar_wait(phi_bar + a, tround & 1);
                    tc::tc_fence_after_sync();
                    uint32_t fh = tc::desc_lo(tc::smem_u32(Phi_s + (uint32_t)a * phi_buf_bytes), 16);
                    uint32_t fl = fh + (phi_bytes >> 4);
                    uint32_t eh = eh_base, el = el_base;
                    for (int ks = 0; ks < ksteps2; ++ks) {
                        tc::umma_tf32_lh(tmem_d, eh, s_hi32, fh, s_hi32, p.idesc_spec, started);
                        started = 1;
                        if (PASSES == 3) {
                            tc::umma_tf32_lh(tmem_c, el, s_hi32, fh, s_hi32, p.idesc_spec, started_c);
                            tc::umma_tf32_lh(tmem_c, eh, s_hi32, fl, s_hi32, p.idesc_spec, 1u);
                            started_c = 1;
                        }
                        eh += 4096 >> 4; el += 4096 >> 4; fh += phi_step; fl += phi_step;
                    }
                }
                for (int kc = 0; kc < nkc; ++kc) {
                    TP_TRACE(1, it * nkc + kc, 0);
                    tc::mbar_wait(((PASSES == 3 || ASRC == 1) ? split_bar : full_bar) + s, ph);
                    TP_TRACE(1, it * nkc + kc, 1);
                    tc::tc_fence_after_sync();
                    uint32_t ah = tc::desc_lo(tc::smem_u32(A_st + s * a_stage_bytes), a_lbo);
                    uint32_t al = NLO ? tc::desc_lo(tc::smem_u32(LO_st + lo_l * a_bytes), a_lbo) : ah + (a_bytes >> 4);
                    // weight tile: K chunk (kc*KC)/32, 32-byte k-steps inside the 128-byte swizzled rows
                    const int kg0 = kc * KC;
                    uint32_t boff = ((uint32_t)(kg0 >> 5) * b_chunk_bytes + (uint32_t)((kg0 & 31) >> 3) * 32) >> 4;
                    for (int ks = 0; ks < ksteps; ++ks) {
#ifdef SB200_BRINGUP
                        if (p.dbg & 4) break;
#endif
                        tc::umma_tf32_lh(tmem_d, ah, a_hi32, bh_base + boff, b_hi32, p.idesc, started);
                        started = 1;
                        if (PASSES == 3) {
                            tc::umma_tf32_lh(tmem_c, al, a_hi32, bh_base + boff, b_hi32, p.idesc, started_c);
                            tc::umma_tf32_lh(tmem_c, ah, a_hi32, bl_base + boff, b_hi32, p.idesc, 1u);
                            started_c = 1;
                        }
                        ah += 1024 >> 4; al += 1024 >> 4; boff += 32 >> 4;   // KC <= 32: stays inside one 32-wide K chunk
                    }
                    tc::umma_commit(empty_bar + s);            // stage reusable once these MMAs have read it
                    TP_TRACE(1, it * nkc + kc, 2);
                    if (NLO) {
                        tc::umma_commit(lo_empty + lo_l);
                        if (++lo_l == (uint32_t)NLO) lo_l = 0;
                    }
                    if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
                }
                tc::umma_commit(tfull_bar + a);                // accumulator of this tile complete
            }
        }
    }
    } else {
        // ================= stagers (tile t+1, t+2, ..) and drainers (tile t): two independent loops =================
        // The staging of the coming tiles runs under the epilogue of tile t instead of in front of it.
        const bool stager = warp < 4 + TP_STAGE_WARPS;
        const int wk = warp - 4 - TP_STAGE_WARPS;              // drainers: 0..15
        // The stagers split the TMA-fed A tiles (3-pass).  The Phi operand of the spectral term and the generated A operand
        // (ASRC == 1) are staged by the drainers in front of their epilogue: both are per-thread register pipelines.
        const int wtid = tid - 32 * (4 + TP_STAGE_WARPS);      // drainers: 0..511
        const int stid = tid - 128;                            // stagers: 0..127
        const int quarter = warp & 3;                          // TMEM lane quarter this warp may access
        const int cpart = wk >> 2;                             // drainers: which quarter of the columns this warp drains
        const int ncol_part = ((p.N + 3) / 4 + 3) & ~3;        // columns per part (multiple of 4)
        const int c_begin = cpart * ncol_part;
        const int c_end = min(p.N, c_begin + ncol_part);
        uint32_t sp_s = 0, sp_ph = 0;                          // split-pass ring position / phase
        uint32_t lo_l = 0, lo_ph = 0;                          // lo ring position / phase
        // Phi staging geometry (tile-invariant, drainer threads): tpr threads cooperate on output channel n_st
        const int tpr = 1 << p.tpr_log2;
        const int n_st = wtid >> p.tpr_log2, sub = wtid & (tpr - 1);
        const int per_n = p.R * p.Mx;
        const bool st_active = SPEC && n_st < p.N;
        const int64_t HW = p.HW;
        const uint64_t hw_bytes = (uint64_t)p.HW * 4;

        // Division-free tile cursors: tile(it) = first + it * stride -> (sample b, tile t inside the sample).  The loop
        // below keeps three of them (epilogue tile it, staged tile it + 1, prefetched tile it + 2) and advances by adds.
        struct Cur { uint32_t b, t; };
        const uint32_t adv_q = stride / tiles_per_b, adv_r = stride - adv_q * tiles_per_b;
        auto advance = [&](Cur c) {
            c.b += adv_q; c.t += adv_r;
            if (c.t >= tiles_per_b) { c.t -= tiles_per_b; ++c.b; }
            return c;
        };
        Cur cur_e; cur_e.b = first / tiles_per_b; cur_e.t = first - cur_e.b * tiles_per_b;
        Cur cur_p = cur_e, cur_f = cur_e;                   // set properly before the loop

        // Phi elements of the first staging round are fetched one tile ahead (registers), so that their global-load
        // latency is hidden behind the epilogue of the previous tile
        float2 fpre[4];
        // tile -> (sample b, first image row y0, 128-px segment v of the row); V == 1: a tile is R whole rows
        const uint32_t Vseg = (uint32_t)p.V, Rrows = (uint32_t)p.R;
        const bool v_pow2 = (Vseg & (Vseg - 1)) == 0;
        const uint32_t v_log2 = 31u - (uint32_t)__clz((int)Vseg);
        auto phi_src = [&](Cur c, uint32_t& v) -> const float2* {
            uint32_t y0;
            if (Vseg == 1) { y0 = c.t * Rrows; v = 0; }
            else { y0 = v_pow2 ? (c.t >> v_log2) : (c.t / Vseg); v = c.t - y0 * Vseg; }
            return p.Phi + ((size_t)(c.b * (uint32_t)p.N + (uint32_t)n_st) * (uint32_t)p.H + y0) * (uint32_t)p.Mx;   // R*Mx contiguous complex
        };
        // tile-invariant staging slots of this thread: element rem = sub + u * tpr of its channel goes to byte offset
        // st_off[u] (kk = 2*rem: k-step rem/4, 16-byte half (rem/2)&1 XOR swizzle bit, 8-byte slot rem&1)
        const bool one_round = per_n <= 4 * tpr;
        uint32_t st_off[4];
        bool st_ok[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int rem = sub + u * tpr;
            st_ok[u] = st_active && rem < per_n;
            st_off[u] = (uint32_t)n_st * 32u + (uint32_t)(rem >> 2) * phi_kstep +
                        (uint32_t)(((((rem >> 1) & 1) ^ ((n_st >> 2) & 1)) << 4) | ((rem & 1) << 3));
        }
        auto prefetch_phi = [&](Cur c) {
            if (SPEC && st_active) {
                uint32_t v;
                const float2* src = phi_src(c, v);
#pragma unroll
                for (int u = 0; u < 4; ++u) fpre[u] = st_ok[u] ? __ldg(src + sub + u * tpr) : make_float2(0.f, 0.f);
            }
        };
        auto prepare_tile = [&](uint32_t it, Cur c) {
            if (SPEC) {
                uint8_t* pbuf = Phi_s + (it & 1) * phi_buf_bytes;
                if (st_active && one_round) {
                    // every element of this thread is already in registers (fpre): rotate (V > 1), split, store
                    const float2* rt = rot_s;
                    if (Vseg > 1) {
                        uint32_t v;
                        (void)phi_src(c, v);
                        rt = rot_s + v * p.Mx;                                                     // V > 1 implies R == 1: kx = rem
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (st_ok[u]) {
                            float2 g = fpre[u];
                            if (Vseg > 1) {
                                const float2 cr = rt[sub + u * tpr];
                                g = make_float2(fpre[u].x * cr.x - fpre[u].y * cr.y, fpre[u].x * cr.y + fpre[u].y * cr.x);
                            }
                            const float2 h = make_float2(tc::tf32_trunc(g.x), tc::tf32_trunc(g.y));
                            *reinterpret_cast<float2*>(pbuf + st_off[u]) = h;
                            if (PASSES == 3) *reinterpret_cast<float2*>(pbuf + phi_bytes + st_off[u]) = make_float2(tc::tf32_lo(g.x, h.x), tc::tf32_lo(g.y, h.y));
                        }
                    }
                } else if (st_active) {
                    uint8_t* ph = pbuf + (uint32_t)n_st * 32u;
                    uint32_t v;
                    const float2* src = phi_src(c, v);
                    const float2* rt = rot_s + v * p.Mx;                                           // V > 1 implies R == 1: kx = rem
                    for (int r0 = sub; r0 < per_n; r0 += 4 * tpr) {
                        float2 f[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int rem = r0 + u * tpr;
                            if (r0 == sub) f[u] = fpre[u];
                            else f[u] = rem < per_n ? __ldg(src + rem) : make_float2(0.f, 0.f);
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int rem = r0 + u * tpr;
                            if (rem < per_n) {
                                float2 g = f[u];
                                if (p.V > 1) {
                                    const float2 cr = rt[rem];
                                    g = make_float2(f[u].x * cr.x - f[u].y * cr.y, f[u].x * cr.y + f[u].y * cr.x);
                                }
                                // kk = 2*rem: k-step rem/4, 16-byte half (rem/2)&1 XOR swizzle bit, 8-byte slot rem&1
                                const uint32_t off = (uint32_t)(rem >> 2) * phi_kstep +
                                                     (uint32_t)(((((rem >> 1) & 1) ^ ((n_st >> 2) & 1)) << 4) | ((rem & 1) << 3));
                                const float2 h = make_float2(tc::tf32_trunc(g.x), tc::tf32_trunc(g.y));
                                *reinterpret_cast<float2*>(ph + off) = h;
                                if (PASSES == 3) *reinterpret_cast<float2*>(ph + phi_bytes + off) = make_float2(tc::tf32_lo(g.x, h.x), tc::tf32_lo(g.y, h.y));
                            }
                        }
                    }
                }
                tc::fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(phi_bar + (it & 1));
            }
            if (ASRC == 1) {
                // generate the operand: thread -> (channel row k_local, 8-pixel chunk c8) of every 32-channel K chunk
                const uint32_t b = c.b;
                const int64_t px = (int64_t)c.t * TP_PX + (wtid & 15) * 8;
                float xv[8];
                {
                    const float* xs = p.gy + (int64_t)b * p.HW + px;
                    if (px + 8 <= p.HW) {
                        const float4 x0 = __ldg(reinterpret_cast<const float4*>(xs)), x1 = __ldg(reinterpret_cast<const float4*>(xs) + 1);
                        xv[0] = x0.x; xv[1] = x0.y; xv[2] = x0.z; xv[3] = x0.w; xv[4] = x1.x; xv[5] = x1.y; xv[6] = x1.z; xv[7] = x1.w;
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) xv[j] = px + j < p.HW ? __ldg(xs + j) : 0.f;
                    }
                }
                const int c8 = wtid & 15;
                for (int kc = 0; kc < nkc; ++kc) {
                    tc::mbar_wait(empty_bar + sp_s, sp_ph ^ 1);          // MMAs that read this stage are done
                    if ((wtid >> 4) < KC) {
                        const int k_local = wtid >> 4;                   // 512 staging threads: one channel row each
                        const uint32_t goff = tc::sw128b32_mnmajor_off(c8 * 8, k_local, (uint32_t)KC * 128u);
                        const int m = kc * KC + k_local;
                        const float w = w2_s[m], bb = red_s[m];
                        float hv[8], lv[8];
#pragma unroll
                        for (int j = 0; j < 8; j += 2) {
                            float g0 = fmaf(w, xv[j], bb), g1 = fmaf(w, xv[j + 1], bb);
                            gelu2(g0, g1);
                            hv[j] = tc::tf32_trunc(g0); hv[j + 1] = tc::tf32_trunc(g1);
                            lv[j] = tc::tf32_lo(g0, hv[j]); lv[j + 1] = tc::tf32_lo(g1, hv[j + 1]);
                        }
                        float4* dh = reinterpret_cast<float4*>(A_st + sp_s * a_stage_bytes + goff);
                        dh[0] = make_float4(hv[0], hv[1], hv[2], hv[3]);
                        dh[1] = make_float4(hv[4], hv[5], hv[6], hv[7]);
                        if (PASSES == 3) {
                            float4* dl = reinterpret_cast<float4*>(A_st + sp_s * a_stage_bytes + a_bytes + goff);
                            dl[0] = make_float4(lv[0], lv[1], lv[2], lv[3]);
                            dl[1] = make_float4(lv[4], lv[5], lv[6], lv[7]);
                        }
                    }
                    tc::fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(split_bar + sp_s);
                    if (++sp_s == (uint32_t)S) { sp_s = 0; sp_ph ^= 1; }
                }
            }
        };

        if (stager) {
            asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TP_REGS_STAGE));
            if (ASRC == 0 && PASSES == 3) {
                // raw fp32 tile -> tf32 hi (in place) and lo, chunk by chunk as the TMA loads land
                for (uint32_t g = 0; g < my_tiles * (uint32_t)nkc; ++g) {
                    if (stid == 0) TP_TRACE(2, g, 0);
                    tc::mbar_wait(full_bar + sp_s, sp_ph);
                    if (stid == 0) TP_TRACE(2, g, 1);
                    float4* ah = reinterpret_cast<float4*>(A_st + sp_s * a_stage_bytes);
                    float4* al = reinterpret_cast<float4*>(A_st + sp_s * a_stage_bytes + a_bytes);
                    if (NLO) {
                        tc::mbar_wait(lo_empty + lo_l, lo_ph ^ 1);       // MMAs of the chunk that used this lo slot are done
                        al = reinterpret_cast<float4*>(LO_st + lo_l * a_bytes);
                        if (++lo_l == (uint32_t)NLO) { lo_l = 0; lo_ph ^= 1; }
                    }
                    for (int idx = stid; idx < (int)(a_bytes / 16); idx += TP_WTHREADS) {
                        const float4 v = ah[idx];
                        const float4 h = make_float4(tc::tf32_trunc(v.x), tc::tf32_trunc(v.y), tc::tf32_trunc(v.z), tc::tf32_trunc(v.w));
                        ah[idx] = h;
                        al[idx] = tc::tf32_lo4(v, h);
                    }
                    tc::fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(split_bar + sp_s);
                    if (stid == 0) TP_TRACE(2, g, 2);
                    if (++sp_s == (uint32_t)S) { sp_s = 0; sp_ph ^= 1; }
                }
            }
        } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(TP_REGS_DRAIN));
        float hsum[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};    // EPI 6 / 7: column sums (x4 | x4) and sum(gy)
        if ((SPEC || ASRC == 1) && my_tiles > 0) {
            prefetch_phi(cur_e);
            prepare_tile(0, cur_e);
            cur_p = advance(cur_e);
            if (my_tiles > 1) prefetch_phi(cur_p);
            cur_f = advance(cur_p);
        }
        for (uint32_t it = 0; it < my_tiles; ++it) {
            // cur_e = tile it (epilogue), cur_p = tile it + 1 (staged now), cur_f = tile it + 2 (prefetched now)
            const uint32_t b = cur_e.b;
            const uint32_t p_base = cur_e.t * TP_PX;
            const int64_t pp = (int64_t)p_base + quarter * 32 + lane;
            const bool in_range = pp < HW;
            float zp[16];
            if (EPI == 3) {
                // GELU' inputs of the first column chunk: in flight while the next tile is prepared
                const float* zsrc = p.zprev + ((int64_t)b * p.N + c_begin) * HW + pp;
                const int nv = min(16, c_end - c_begin);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    zp[j] = (in_range && j < nv) ? __ldg(zsrc) : 0.f;
                    zsrc += HW;
                }
            }
            if (SPEC || ASRC == 1) {
                if (it + 1 < my_tiles) {
                    prepare_tile(it + 1, cur_p);
                    if (it + 2 < my_tiles) prefetch_phi(cur_f);
                }
                cur_e = cur_p; cur_p = cur_f; cur_f = advance(cur_f);
            } else {
                cur_e = advance(cur_e);
            }
            const uint32_t a = it & 1;
            const uint32_t tround = it >> 1;
            tc::mbar_wait(tfull_bar + a, tround & 1);
            if (wk == 0 && lane == 0) TP_TRACE(3, it, 2);
            tc::tc_fence_after_sync();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + a * (uint32_t)p.N * (p.corr ? 2u : 1u);
            if constexpr (EPI == 5) {
                // ---- fused head forward: this warp's columns -> one partial dot product per pixel ----
                float part = 0.f, part1 = 0.f, part2 = 0.f, part3 = 0.f;      // four chains instead of one 64-deep FMA chain
                for (int c0 = c_begin; c0 < c_end; c0 += 16) {
#ifdef SB200_BRINGUP
                    if (p.dbg & 2) break;
#endif
                    uint32_t r[16];
                    tc::tmem_ld_32x32b_x16(taddr + (uint32_t)c0, r);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        const float4 bq = *reinterpret_cast<const float4*>(bias_s + c0 + j);
                        const float4 wq = *reinterpret_cast<const float4*>(w2_s + c0 + j);
                        float g0 = __uint_as_float(r[j + 0]) + bq.x, g1 = __uint_as_float(r[j + 1]) + bq.y;
                        float g2 = __uint_as_float(r[j + 2]) + bq.z, g3 = __uint_as_float(r[j + 3]) + bq.w;
#ifdef SB200_BRINGUP
                        if (!(p.dbg & 1))
#endif
                        {
                            gelu2(g0, g1);
                            gelu2(g2, g3);
                        }
                        part = fmaf(wq.x, g0, part);
                        part1 = fmaf(wq.y, g1, part1);
                        part2 = fmaf(wq.z, g2, part2);
                        part3 = fmaf(wq.w, g3, part3);
                    }
                }
                part = (part + part1) + (part2 + part3);
                tc::tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(tempty_bar + a);
                if (wk == 0 && lane == 0) TP_TRACE(3, it, 3);
                // the four warps sharing this TMEM lane quarter hold the four column parts of the same 32 pixels
                float* red = red_s + (it & 1) * 512;
                red[cpart * 128 + quarter * 32 + lane] = part;
                asm volatile("bar.sync %0, 128;" ::"r"(2 + quarter) : "memory");
                if (cpart == 0 && in_range) {
                    const int q = quarter * 32 + lane;
                    const float y = (red[q] + red[128 + q]) + (red[256 + q] + red[384 + q]) + (p.b2 ? __ldg(p.b2) : 0.f);
                    p.y_out[(int64_t)b * HW + pp] = y;
                }
                continue;
            }
            if constexpr (EPI == 6) {
                // ---- fused head backward, stage 1 (N == 256: four 16-column chunks per warp) ----
                const float gyv = in_range ? __ldg(p.gy + (int64_t)b * HW + pp) : 0.f;
                if (cpart == 0) hsum[8] += gyv;
#pragma unroll
                for (int ci = 0; ci < 4; ++ci) {
                    const int c0 = c_begin + 16 * ci;
                    uint32_t r[16];
                    tc::tmem_ld_32x32b_x16(taddr + (uint32_t)c0, r);
                    tc::tmem_ld_wait();
                    float s0[16], s1[16];
#pragma unroll
                    for (int j = 0; j < 16; j += 2) {
                        const float za = __uint_as_float(r[j]) + bias_s[c0 + j], zb = __uint_as_float(r[j + 1]) + bias_s[c0 + j + 1];
                        float2 cdf, e;
                        gelu_core2(za, zb, cdf, e);
                        const float gpa = fmaf(za * 0.39894228040143267794f, e.x, cdf.x);
                        const float gpb = fmaf(zb * 0.39894228040143267794f, e.y, cdf.y);
                        s0[j] = w2_s[c0 + j] * gyv * gpa;           // gz1 (0 outside the image: gyv = 0)
                        s0[j + 1] = w2_s[c0 + j + 1] * gyv * gpb;
                        s1[j] = gyv * (za * cdf.x);                  // gy * GELU(z1)
                        s1[j + 1] = gyv * (zb * cdf.y);
                    }
                    if (in_range) {
                        uint64_t ya = reinterpret_cast<uint64_t>(p.y_out + ((int64_t)b * p.N + c0) * HW + pp);
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            asm volatile("st.global.f32 [%0], %1;" ::"l"(ya), "f"(s0[j]) : "memory");
                            ya += hw_bytes;
                        }
                    }
                    hsum[ci] += warp_colsum16(s0, lane);
                    hsum[4 + ci] += warp_colsum16(s1, lane);
                }
                tc::tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(tempty_bar + a);
                continue;
            }
            if constexpr (EPI == 7) {
                const float xv = in_range ? __ldg(p.gy + (int64_t)b * HW + pp) : 0.f;
#pragma unroll
                for (int ci = 0; ci < 4; ++ci) {
                    const int c0 = c_begin + 16 * ci;
                    uint32_t r[16];
                    tc::tmem_ld_32x32b_x16(taddr + (uint32_t)c0, r);
                    tc::tmem_ld_wait();
                    float s0[16], s1[16];
#pragma unroll
                    for (int j = 0; j < 16; j += 2) {
                        float ga, gb;
                        gelu_grad2(fmaf(w2_s[c0 + j], xv, bias_s[c0 + j]), fmaf(w2_s[c0 + j + 1], xv, bias_s[c0 + j + 1]), ga, gb);
                        const float gza = in_range ? __uint_as_float(r[j]) * ga : 0.f;
                        const float gzb = in_range ? __uint_as_float(r[j + 1]) * gb : 0.f;
                        s0[j] = gza; s0[j + 1] = gzb;
                        s1[j] = gza * xv; s1[j + 1] = gzb * xv;
                    }
                    hsum[ci] += warp_colsum16(s0, lane);
                    hsum[4 + ci] += warp_colsum16(s1, lane);
                }
                tc::tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(tempty_bar + a);
                continue;
            }
            for (int c0 = c_begin; c0 < c_end; c0 += 16) {
#ifdef SB200_BRINGUP
                if (p.dbg & 2) break;
#endif
                const int nv = min(16, c_end - c0);
                const int64_t off0 = ((int64_t)b * p.N + c0) * HW + pp;
                uint32_t r[16];
                if (EPI == 3 && c0 != c_begin) {
                    const float* zsrc = p.zprev + off0;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        zp[j] = (in_range && j < nv) ? __ldg(zsrc) : 0.f;
                        zsrc += HW;
                    }
                }
                tc::tmem_ld_32x32b_x16(taddr + (uint32_t)c0, r);
                tc::tmem_ld_wait();
                if (PASSES == 3 && p.corr) {
                    uint32_t rc[16];
                    tc::tmem_ld_32x32b_x16(taddr + (uint32_t)(p.N + c0), rc);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(rc[j]));
                }
                if (bias_epi) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        const float4 bq = *reinterpret_cast<const float4*>(bias_s + c0 + j);   // c0 % 4 == 0; bias_s zero-padded
                        r[j + 0] = __float_as_uint(__uint_as_float(r[j + 0]) + bq.x);
                        r[j + 1] = __float_as_uint(__uint_as_float(r[j + 1]) + bq.y);
                        r[j + 2] = __float_as_uint(__uint_as_float(r[j + 2]) + bq.z);
                        r[j + 3] = __float_as_uint(__uint_as_float(r[j + 3]) + bq.w);
                    }
                }
                if (in_range && nv == 16) {
                    // fast path (whole chunk valid): straight-line code, 16 independent GELU chains in flight
                    float v[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
                    if (EPI == 1) {
                        uint64_t za = reinterpret_cast<uint64_t>(p.z_out + off0);
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            asm volatile("st.global.f32 [%0], %1;" ::"l"(za), "f"(v[j]) : "memory");
                            za += hw_bytes;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 16; j += 2) {
#ifdef SB200_BRINGUP
                        if (p.dbg & 1) continue;
#endif
                        if (EPI == 1 || EPI == 2) gelu2(v[j], v[j + 1]);
                        if (EPI == 3) {
                            float ga, gb;
                            gelu_grad2(zp[j], zp[j + 1], ga, gb);
                            v[j] *= ga; v[j + 1] *= gb;
                        }
                    }
                    uint64_t ya = reinterpret_cast<uint64_t>(p.y_out + off0);
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        asm volatile("st.global.f32 [%0], %1;" ::"l"(ya), "f"(v[j]) : "memory");
                        ya += hw_bytes;
                    }
                } else if (in_range) {
                    float* ydst = p.y_out + off0;
                    float* zdst = (EPI == 1) ? p.z_out + off0 : nullptr;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        if (j < nv) {
                            float v = __uint_as_float(r[j]);
                            if (EPI == 1) { *zdst = v; zdst += HW; }
                            if (EPI == 1 || EPI == 2) v = gelu_f(v);
                            if (EPI == 3) v *= gelu_grad_f(zp[j]);
                            *ydst = v;
                            ydst += HW;
                        }
                    }
                }
            }
            tc::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(tempty_bar + a);
            if (wk == 0 && lane == 0) TP_TRACE(3, it, 3);
        }
        if (EPI == 6 || EPI == 7) {
            // flush the per-warp column sums: row = (CTA, lane quarter); each cpart owns 64 of the 256 columns
            const int col = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
            float* row = p.colsum_ws + (size_t)(blockIdx.x * 4 + quarter) * 512;
            if (!(lane & 1)) {
    #pragma unroll
                for (int ci = 0; ci < 4; ++ci) {
                    row[cpart * 64 + 16 * ci + col] = hsum[ci];
                    row[256 + cpart * 64 + 16 * ci + col] = hsum[4 + ci];
                }
            }
            if (cpart == 0) {
                float g = hsum[8];
    #pragma unroll
                for (int o = 16; o > 0; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
                if (lane == 0) p.colsum_ws[(size_t)gridDim.x * 4 * 512 + blockIdx.x * 4 + quarter] = g;
            }
        }
        }   // drainers
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, p.tmem_cols);
}

#ifdef SB200_BRINGUP
static long long* tp_trace_buf(cudaStream_t st) {
    static long long* dev = nullptr;
    if (!dev) cudaMalloc(&dev, 4 * 32 * 4 * sizeof(long long));
    cudaMemsetAsync(dev, 0, 4 * 32 * 4 * sizeof(long long), st);
    return dev;
}
static void tp_trace_dump(const TcPwParams& p, int epi, cudaStream_t st) {
    long long h[4 * 32 * 4];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, p.trace, sizeof(h), cudaMemcpyDeviceToHost);
    const long long t0 = h[2];
    printf("tc_pointwise trace: epi=%d M=%d N=%d spectral=%d KC=%d nkc=%d stages=%d nlo=%d\n", epi, p.M, p.N, p.Phi != nullptr, p.KC, p.nkc, p.stages, p.nlo);
    const char* names[4] = {"producer(wait0,wait1,issued)", "mma chunk(wait0,wait1,committed)", "split(wait0,landed,done)", "tile(mma tempty0,tempty1; epi wake,done)"};
    for (int r = 0; r < 4; ++r) {
        printf("%s\n", names[r]);
        for (int i = 0; i < (r == 3 ? 8 : 16); ++i) {
            printf("  %2d:", i);
            for (int e = 0; e < 4; ++e) printf(" %7lld", h[(r * 32 + i) * 4 + e] ? h[(r * 32 + i) * 4 + e] - t0 : -1);
            printf("\n");
        }
    }
    fflush(stdout);
}
#endif

// out[0][n] = gb1, out[1][n] = gw2 (n < 256), gb2 = sum(gy): fixed-order sums over the (CTA, quarter) rows.
// 32 consecutive outputs x 8 row lanes per block (coalesced partial reads), shared-memory tree at the end.
__global__ void __launch_bounds__(256) head_colsum_reduce_kernel(const float* __restrict__ ws, int rows, float* __restrict__ gb1,
                                                                 float* __restrict__ gw2, float* __restrict__ gb2, int N) {
    __shared__ float part[8][33];
    const int o = threadIdx.x & 31, rl = threadIdx.x >> 5;
    if (blockIdx.x == 16) {                              // gb2: the tail of the workspace
        float g = 0.f;
        for (int r = threadIdx.x; r < rows; r += 256) g += ws[(size_t)rows * 512 + r];
        part[rl][o] = g;
        __syncthreads();
        if (threadIdx.x == 0 && gb2) {
            float t = 0.f;
            for (int a = 0; a < 8; ++a)
                for (int c = 0; c < 32; ++c) t += part[a][c];
            *gb2 = t;
        }
        return;
    }
    const int idx = blockIdx.x * 32 + o;                 // 0..511
    float acc = 0.f;
    for (int r = rl; r < rows; r += 8) acc += ws[(size_t)r * 512 + idx];
    part[rl][o] = acc;
    __syncthreads();
    if (rl == 0) {
#pragma unroll
        for (int a = 1; a < 8; ++a) acc += part[a][o];
        if (idx < 256) { if (idx < N) gb1[idx] = acc; }
        else if (idx - 256 < N) gw2[idx - 256] = acc;
    }
}

// ---------------------------------------------------------------------------------------------
// dispatch (called from sb200_rowidft_pointwise)
// ---------------------------------------------------------------------------------------------

int sb200_tc_rowidft_pointwise(sb200_plan_t plan, int pass, const PwParams& q, cudaStream_t st, int* handled) {
    *handled = 0;
    if (g_tc_mode == 0) return 0;
    const int64_t HW = (int64_t)q.H * q.W;
    const int M = q.M, N = q.N;
    const bool has_pw = q.Wp != nullptr, has_spec = q.Phi != nullptr;
    if (has_pw && (M % 8 != 0 || !(M <= 32 || M % 32 == 0))) return 0;
    if (N % 16 != 0 || N < 16 || N > 256) return 0;
    if (HW % 4 != 0) return 0;
    if (has_pw && ((reinterpret_cast<uintptr_t>(q.A) & 15) != 0 || (int64_t)q.B * M >= (1LL << 31))) return 0;
    const sb200_tc_tables* tt = (const sb200_tc_tables*)plan->tc;
    if (has_spec && (tt == nullptr || (reinterpret_cast<uintptr_t>(q.Phi) & 7) != 0 || tt->V * q.Mx > 256)) return 0;
    const int passes = g_tc_mode;

    TcPwParams p = {};
    memset(&p, 0, sizeof(p));
    p.Wp = q.Wp; p.w_sn = q.w_sn; p.w_sm = q.w_sm; p.bias = q.bias; p.zprev = q.zprev;
    p.z_out = q.z_out; p.y_out = q.y_out; p.B = q.B; p.M = M; p.N = N; p.HW = HW;
    p.KC = has_pw ? (M < 32 ? M : 32) : 8;
    if (const int kc = sb_env_int("SB200_TP_KC", 0)) if (has_pw && (kc == 8 || kc == 16) && M % kc == 0 && kc < p.KC) p.KC = kc;
    p.nkc = has_pw ? M / p.KC : 0;
    p.mode = q.mode; p.apply_act = q.apply_act;
    p.idesc = tc::make_idesc_tf32(128, N, /*A MN-major*/ 1, /*B K-major*/ 0);
    p.idesc_spec = tc::make_idesc_tf32(128, N, 0, 0);
    p.corr = (passes == 3 && 4 * N <= 512) ? 1 : 0;
    p.dbg = sb_env_int("SB200_TP_DBG", 0);
    uint32_t cols = 32;
    while (cols < (uint32_t)(2 * N * (p.corr ? 2 : 1))) cols <<= 1;
    p.tmem_cols = cols;
    p.ntiles = (HW + TP_PX - 1) / TP_PX * q.B;
    p.H = q.H; p.W = q.W; p.Mx = q.Mx;
    if (has_spec) {
        p.Phi = q.Phi; p.E = tt->E[pass]; p.rot = tt->rot;
        p.R = tt->R; p.V = tt->V; p.K2 = tt->K2; p.K2pad = tt->K2pad;
        p.bias_mma = (passes == 3 && q.bias != nullptr && tt->K2 < tt->K2pad) ? 1 : 0;
    }

    const size_t mult = passes == 3 ? 2 : 1;
    const size_t b_bytes = has_pw ? ((((size_t)((M + 31) / 32) * N * 128 + 1023) & ~(size_t)1023) * mult) : 0;
    const size_t e_bytes = has_spec ? (size_t)(p.K2pad / 8) * 4096 * mult : 0;
    const size_t phi_bytes = has_spec ? ((((size_t)(p.K2pad / 8) * N * 32 + 1023) & ~(size_t)1023) * mult * 2) : 0;
    const size_t fixed0 = 1024 + b_bytes + e_bytes + phi_bytes + 512 + 1024 + 2048;   // + barriers, bias_s, rot_s
    // 3-pass: lo ring of two slots, the stages are raw slots alone.  When fewer than two tiles of 32-channel slots fit
    // (large resident weight tile), 16-channel slots keep the same bytes in flight at a finer release granularity.
    p.nlo = (passes == 3 && has_pw) ? sb_env_int("SB200_TP_NLO", 0) : 0;
    if (p.nlo && M % 16 == 0 && p.KC == 32 && (224 * 1024 - (long long)fixed0 - 2 * 16384) / 16384 < 2 * (M / 32)) {
        p.KC = 16;
        p.nkc = M / 16;
    }
    const size_t a_stage = (size_t)p.KC * 512 * (p.nlo ? 1 : mult);
    const size_t fixed = fixed0 + (size_t)p.nlo * p.KC * 512;
    int stages = has_pw ? (p.nlo ? 8 : 6) : 1;
    static const size_t ring_budget = []() {                       // bytes the ring may grow to (SB200_TP_SMEM_KB: bring-up builds)
        const int kb = sb_env_int("SB200_TP_SMEM_KB", 0);
        return (size_t)((kb >= 64 && kb <= 227) ? kb : 224) * 1024;
    }();
    while (stages > 2 && fixed + stages * a_stage > ring_budget) --stages;
    if (fixed + stages * a_stage > 227 * 1024) {
        stages = 1;
        if (fixed + stages * a_stage > 227 * 1024) return 0;           // does not fit: CUDA-core kernel
    }
    p.stages = stages;
    const size_t smem = fixed + stages * a_stage;

    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    if (has_pw)
        if (int rc = sb200_make_tmap_2d_f32(&tmap, q.A, (uint64_t)HW, (uint64_t)q.B * M, (uint64_t)HW * 4, 32, (uint32_t)p.KC, 2))
            return rc;
    const int g_num_sms = sb200_num_sms();
    const unsigned grid = (unsigned)(p.ntiles < g_num_sms ? p.ntiles : g_num_sms);
    int tl = 0;
    while ((1 << (tl + 1)) * N <= 32 * TP_DRAIN_WARPS) ++tl;
    p.tpr_log2 = tl;
    int epi;
    if (q.mode == 0) epi = q.apply_act ? (q.z_out ? 1 : 2) : 0;
    else epi = q.zprev ? 3 : 4;
    if (q.mode == 0 && !q.apply_act && q.z_out) return 0;              // (never requested) keep the generic kernel
#define TP_LAUNCH_S(PS, EP, SP)                                                                                        \
    do {                                                                                                               \
        SB_CHECK_CUDA(cudaFuncSetAttribute(tc_pointwise_kernel<PS, EP, 0, SP>,                                        \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                   \
        sb_launch(tc_pointwise_kernel<PS, EP, 0, SP>, grid, TP_THREADS, smem, st, tmap, p);                            \
    } while (0)
#define TP_LAUNCH(PS, EP)                                                                                              \
    do {                                                                                                               \
        if (has_spec) TP_LAUNCH_S(PS, EP, true); else TP_LAUNCH_S(PS, EP, false);                                      \
    } while (0)
#define TP_LAUNCH_EPI(PS)                                                                                              \
    switch (epi) {                                                                                                     \
        case 0: TP_LAUNCH(PS, 0); break;                                                                               \
        case 1: TP_LAUNCH(PS, 1); break;                                                                               \
        case 2: TP_LAUNCH(PS, 2); break;                                                                               \
        case 3: TP_LAUNCH(PS, 3); break;                                                                               \
        default: TP_LAUNCH(PS, 4); break;                                                                              \
    }
#ifdef SB200_BRINGUP
    const int want_trace = sb_env_int("SB200_TP_TRACE", 0);
    const bool trace_this = want_trace && sb_env_int("SB200_TP_TRACE_EPI", -1) == epi;
    if (trace_this) p.trace = tp_trace_buf(st);
#endif
    if (passes == 3) { TP_LAUNCH_EPI(3) } else { TP_LAUNCH_EPI(1) }
#undef TP_LAUNCH_EPI
#undef TP_LAUNCH
#undef TP_LAUNCH_S
    SB_LAUNCH_CHECK();
#ifdef SB200_BRINGUP
    if (trace_this) {
        static int calls = 0;
        if (++calls == want_trace) tp_trace_dump(p, epi, st);
    }
#endif
    *handled = 1;
    return 0;
}


// ---------------------------------------------------------------------------------------------
// fused pointwise MLP head (projection of the FNO: C -> N=256 -> 1 channel)
// ---------------------------------------------------------------------------------------------
static int tp_head_launch(int epi, const float* h, const float* W1, int64_t w_sn, int64_t w_sm, const float* b1,
                          const float* w2, const float* b2, const float* gy, float* out, float* ws, int B, int M, int N,
                          int64_t HW, cudaStream_t st, unsigned* grid_out) {
    SB_REQUIRE(g_tc_mode != 0, "mlp_head: the fused head runs on the tcgen05 path (tc mode 1 or 3)");
    SB_REQUIRE(M % 8 == 0 && (M <= 32 || M % 32 == 0), "mlp_head: in-channels %d not supported (multiple of 8; of 32 above 32)", M);
    SB_REQUIRE(N == 256, "mlp_head: hidden width must be 256 (got %d)", N);
    SB_REQUIRE(HW % 4 == 0 && (reinterpret_cast<uintptr_t>(h) & 15) == 0 && (int64_t)B * M < (1LL << 31), "mlp_head: layout");
    const int passes = g_tc_mode;
    TcPwParams p = {};
    memset(&p, 0, sizeof(p));
    p.Wp = W1; p.w_sn = w_sn; p.w_sm = w_sm; p.bias = b1; p.y_out = out; p.B = B; p.M = M; p.N = N; p.HW = HW;
    p.KC = M < 32 ? M : 32;
    if (const int kc = sb_env_int("SB200_TP_KC", 0)) if ((kc == 8 || kc == 16) && M % kc == 0 && kc < p.KC) p.KC = kc;
    p.nkc = M / p.KC;
    p.idesc = tc::make_idesc_tf32(128, N, 1, 0);
    p.idesc_spec = tc::make_idesc_tf32(128, N, 0, 0);
    p.tmem_cols = 512;
    p.ntiles = (HW + TP_PX - 1) / TP_PX * B;
    p.w2 = w2; p.b2 = b2; p.gy = gy; p.colsum_ws = ws;
    p.dbg = sb_env_int("SB200_TP_DBG", 0);
    const size_t mult = passes == 3 ? 2 : 1;
    const size_t b_bytes = ((((size_t)((M + 31) / 32) * N * 128 + 1023) & ~(size_t)1023) * mult);
    const size_t fixed0 = 1024 + b_bytes + 512 + 1024 + 2048 + 1024 + 4096;     // + barriers, bias_s, rot_s, w2_s, red_s
    // 3-pass: two-slot lo ring + raw slots; the 256-row weight tile leaves ~85 KB, so the slots are 16 channels deep
    // (two tiles of the activation stream in flight instead of one: the loop was bound by the TMA round trip)
    p.nlo = passes == 3 ? sb_env_int("SB200_TP_NLO", 0) : 0;
    if (p.nlo && M % 16 == 0 && p.KC == 32 && (224 * 1024 - (long long)fixed0 - 2 * 16384) / 16384 < 2 * (M / 32)) {
        p.KC = 16;
        p.nkc = M / 16;
    }
    const size_t a_stage = (size_t)p.KC * 512 * (p.nlo ? 1 : mult);
    const size_t fixed = fixed0 + (size_t)p.nlo * p.KC * 512;
    int stages = p.nlo ? 8 : 6;
    int stages_max = sb_env_int("SB200_TP_STAGES", 8);
    if (stages > stages_max) stages = stages_max;
    while (stages > 2 && fixed + stages * a_stage > (size_t)sb_env_int("SB200_TP_HEAD_KB", 224) * 1024) --stages;
    SB_REQUIRE(fixed + stages * a_stage <= 227 * 1024, "mlp_head: shared memory does not fit (M=%d)", M);
    p.stages = stages;
    const size_t smem = fixed + stages * a_stage;
    CUtensorMap tmap;
    if (int rc = sb200_make_tmap_2d_f32(&tmap, h, (uint64_t)HW, (uint64_t)B * M, (uint64_t)HW * 4, 32, (uint32_t)p.KC, 2)) return rc;
    const int g_num_sms = sb200_num_sms();
    const unsigned grid = (unsigned)(p.ntiles < g_num_sms ? p.ntiles : g_num_sms);
    *grid_out = grid;
#define TP_HEAD(PS, EP)                                                                                                \
    do {                                                                                                               \
        SB_CHECK_CUDA(cudaFuncSetAttribute(tc_pointwise_kernel<PS, EP>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                           (int)smem));                                                                \
        sb_launch(tc_pointwise_kernel<PS, EP>, grid, TP_THREADS, smem, st, tmap, p);                                          \
    } while (0)
#ifdef SB200_BRINGUP
    static long long* trace_dev = nullptr;
    const int want_trace = epi == 5 ? sb_env_int("SB200_TP_TRACE", 0) : 0;
    if (want_trace) {
        if (!trace_dev) cudaMalloc(&trace_dev, 4 * 32 * 4 * sizeof(long long));
        cudaMemsetAsync(trace_dev, 0, 4 * 32 * 4 * sizeof(long long), st);
        p.trace = trace_dev;
    }
#endif
    if (epi == 5)      { if (passes == 3) TP_HEAD(3, 5); else TP_HEAD(1, 5); }
    else if (epi == 6) { if (passes == 3) TP_HEAD(3, 6); else TP_HEAD(1, 6); }
    else               { if (passes == 3) TP_HEAD(3, 7); else TP_HEAD(1, 7); }
#undef TP_HEAD
    SB_LAUNCH_CHECK();
#ifdef SB200_BRINGUP
    if (want_trace) {
        static int calls = 0;
        if (++calls == want_trace) {
            long long h[4 * 32 * 4];
            cudaStreamSynchronize(st);
            cudaMemcpy(h, p.trace, sizeof(h), cudaMemcpyDeviceToHost);
            const long long t0 = h[2];
            printf("head trace: KC=%d nkc=%d stages=%d nlo=%d\n", p.KC, p.nkc, p.stages, p.nlo);
            const char* names[4] = {"producer(wait0,wait1,issued)", "mma chunk(wait0,wait1,committed)", "split(wait0,landed,done)", "tile(mma tempty0,tempty1; epi wake,done)"};
            for (int r = 0; r < 4; ++r) {
                printf("%s\n", names[r]);
                for (int i = 0; i < (r == 3 ? 8 : 24); ++i) {
                    printf("  %2d:", i);
                    for (int e = 0; e < 4; ++e) printf(" %7lld", h[(r * 32 + i) * 4 + e] ? h[(r * 32 + i) * 4 + e] - t0 : -1);
                    printf("\n");
                }
            }
            fflush(stdout);
        }
    }
#endif
    return 0;
}

static int tp_sms() {
    return sb200_num_sms();
}

extern "C" int sb200_mlp_head_fwd(const float* h, const float* W1, const float* b1, const float* w2, const float* b2,
                                  float* y, int B, int M, int N, int64_t HW, void* stream, int tc_mode) {
    SbModeScope _mode(tc_mode);
    SB_REQUIRE(h && W1 && b1 && w2 && y, "mlp_head_fwd: NULL argument");
    if (B <= 0) return 0;
    unsigned grid;
    return tp_head_launch(5, h, W1, M, 1, b1, w2, b2, nullptr, y, nullptr, B, M, N, HW, (cudaStream_t)stream, &grid);
}

extern "C" int64_t sb200_mlp_head_bwd_workspace(void) { return (int64_t)tp_sms() * 4 * 513; }

extern "C" int sb200_mlp_head_bwd(const float* h, const float* W1, const float* b1, const float* w2, const float* gy,
                                  float* gz1, float* gb1, float* gw2, float* gb2, float* workspace, int B, int M, int N,
                                  int64_t HW, void* stream, int tc_mode) {
    SbModeScope _mode(tc_mode);
    SB_REQUIRE(h && W1 && b1 && w2 && gy && gz1 && gb1 && gw2 && workspace, "mlp_head_bwd: NULL argument");
    if (B <= 0) return 0;
    unsigned grid;
    if (int rc = tp_head_launch(6, h, W1, M, 1, b1, w2, nullptr, gy, gz1, workspace, B, M, N, HW, (cudaStream_t)stream, &grid)) return rc;
    sb_launch(head_colsum_reduce_kernel, 17, 256, 0, (cudaStream_t)stream, workspace, (int)grid * 4, gb1, gw2, gb2, N);
    SB_LAUNCH_CHECK();
    return 0;
}

// Lifting tail backward for a 1-input-channel lifting MLP  x -> gelu(w1 x + b1) [256] -> W2 [C,256]:
//   gz1[b,n,p] = (sum_c W2[c,n] g[b,c,p]) * gelu'(w1[n] x[b,p] + b1[n])  stays on chip; only
//   gb1[n] = sum gz1 and gw1[n] = sum gz1 x leave.   g [B,C,HW], W2 [C,256] (row-major), x [B,HW].
extern "C" int sb200_lift_tail_bwd(const float* g, const float* W2, const float* w1, const float* b1, const float* x,
                                   float* gw1, float* gb1, float* workspace, int B, int C, int N, int64_t HW,
                                   void* stream, int tc_mode) {
    SbModeScope _mode(tc_mode);
    SB_REQUIRE(g && W2 && w1 && b1 && x && gw1 && gb1 && workspace, "lift_tail_bwd: NULL argument");
    if (B <= 0) return 0;
    unsigned grid;
    // out channel n of the transposed product reads W2[c, n]: stride 1 over n, N over the contraction index c
    if (int rc = tp_head_launch(7, g, W2, 1, N, b1, w1, nullptr, x, nullptr, workspace, B, C, N, HW, (cudaStream_t)stream, &grid))
        return rc;
    sb_launch(head_colsum_reduce_kernel, 17, 256, 0, (cudaStream_t)stream, workspace, (int)grid * 4, gb1, gw1, nullptr, N);
    SB_LAUNCH_CHECK();
    return 0;
}

// Lifting MLP forward for ONE input channel (neuralop FNO.lifting = MLP(1 -> 256 -> C)):
//      y[b,c,p] = sum_n W2[c,n] gelu(w1[n] x[b,p] + b1[n]) + b2[c]
// the hidden operand is generated tile by tile in shared memory (ASRC = 1), so it never reaches HBM.
extern "C" int sb200_lift_fwd(const float* x, const float* w1, const float* b1, const float* W2, const float* b2, float* y,
                              int B, int N, int C, int64_t HW, void* stream, int tc_mode) {
    SbModeScope _mode(tc_mode);
    SB_REQUIRE(x && w1 && b1 && W2 && y, "lift_fwd: NULL argument");
    SB_REQUIRE(g_tc_mode != 0, "lift_fwd: runs on the tcgen05 path (tc mode 1 or 3)");
    SB_REQUIRE(N == 256, "lift_fwd: hidden width must be 256 (got %d)", N);
    SB_REQUIRE(C % 16 == 0 && C >= 16 && C <= 256, "lift_fwd: output channels %d not supported", C);
    SB_REQUIRE(HW % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0, "lift_fwd: layout");
    if (B <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int passes = g_tc_mode;
    TcPwParams p = {};
    memset(&p, 0, sizeof(p));
    p.Wp = W2; p.w_sn = N; p.w_sm = 1; p.bias = b2; p.y_out = y; p.B = B; p.M = N; p.N = C; p.HW = HW;
    p.KC = 32; p.nkc = N / 32;
    p.idesc = tc::make_idesc_tf32(128, C, 1, 0);
    p.idesc_spec = tc::make_idesc_tf32(128, C, 0, 0);
    p.corr = (passes == 3 && 4 * C <= 512) ? 1 : 0;
    p.dbg = sb_env_int("SB200_TP_DBG", 0);
    uint32_t cols = 32;
    while (cols < (uint32_t)(2 * C * (p.corr ? 2 : 1))) cols <<= 1;
    p.tmem_cols = cols;
    p.ntiles = (HW + TP_PX - 1) / TP_PX * B;
    p.w2 = w1; p.b2 = b1; p.gy = x;
    const size_t mult = passes == 3 ? 2 : 1;
    const size_t a_stage = (size_t)p.KC * 512 * mult;
    const size_t b_bytes = ((((size_t)(N / 32) * C * 128 + 1023) & ~(size_t)1023) * mult);
    const size_t fixed = 1024 + b_bytes + 512 + 1024 + 2048 + 1024 + 4096;
    int stages = 6;
    while (stages > 2 && fixed + stages * a_stage > 208 * 1024) --stages;
    SB_REQUIRE(fixed + stages * a_stage <= 227 * 1024, "lift_fwd: shared memory does not fit (C=%d)", C);
    p.stages = stages;
    const size_t smem = fixed + stages * a_stage;
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    const int sms = tp_sms();
    const unsigned grid = (unsigned)(p.ntiles < sms ? p.ntiles : sms);
    if (passes == 3) {
        SB_CHECK_CUDA(cudaFuncSetAttribute(tc_pointwise_kernel<3, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sb_launch(tc_pointwise_kernel<3, 0, 1>, grid, TP_THREADS, smem, st, tmap, p);
    } else {
        SB_CHECK_CUDA(cudaFuncSetAttribute(tc_pointwise_kernel<1, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sb_launch(tc_pointwise_kernel<1, 0, 1>, grid, TP_THREADS, smem, st, tmap, p);
    }
    SB_LAUNCH_CHECK();
    return 0;
}
