// tcgen05 truncated real DFT along W (the row stage of the analysis for grids the fused small-grid kernel
// does not cover: W = 128, 256, ...):
//
//      T[row, kx] = sum_x x[row, x] * rowF[x, kx]          rows = nimg * H,  kx < Mx  (complex, interleaved)
//
// as ONE dense GEMM against the precomputed twiddle matrix (SURVEY.md 8a / north_star (a)):
//      D[128 rows, N] = A[128 rows, K = W] * B[N, K]^T,    N = 2*Mx rounded up to 16, n = 2*kx + {0: re, 1: im}
//   A : the image rows as they lie in memory (the contraction index x is the contiguous one), TMA boxes
//       {32 x, 128 rows} with the 128B swizzle = the K-major SW128 UMMA operand directly; a ring of K chunks
//   B : the twiddles, resident in shared memory for the whole kernel (K-major SW128, tf32 hi / lo)
//   D : fp32 in TMEM, double buffered, so the epilogue of tile t overlaps the MMAs of tile t + 1
// replaces torch.fft.rfftn's row pass + the kx slice of neuralop SpectralConv.forward (and, with the pass-1
// tables, the adjoint of irfftn's row pass).  3xTF32 operand split as in tc_pointwise.cu (parity <= 1e-5);
// sb200_set_tc_mode(1) runs the single TF32 pass, mode 0 keeps the FFMA kernel (rowdft_fwd_kernel).
// The same kernel also runs the column stage of the analysis (MODE 2) when the row stage leaves its result in the
// planar, transposed layout T'[img][n = 2*kx + c][y] (MODE 1): then the column DFT is again "rows x K" against a
// resident twiddle matrix, with rows = (img, kx, re|im), K = y and B[n' = 2*ky + c'][y] = (Re, Im) colF[ky][y]:
//      D[(img,kx,0)][(ky,r)] = sum_y a*tr   D[(img,kx,0)][(ky,i)] = sum_y b*tr        (colF = a + i b, T = tr + i ti)
//      D[(img,kx,1)][(ky,r)] = sum_y a*ti   D[(img,kx,1)][(ky,i)] = sum_y b*ti
//      Xh[img][ky][kx] = (D00 - D11) + i (D10 + D01): the two rows are adjacent TMEM lanes, one SHFL in the epilogue.
// The intermediate never takes the interleaved [.., Mx, 2] form, its stores are full 128-byte lines, and both
// stages of the large-grid analysis run on the tensor cores.
//   warp 0     TMA producer (one lane)
//   warp 1     MMA issuer (one lane)
//   warps 2-17 workers: tf32 hi/lo split of the landed chunks of tile t + 1, then epilogue of tile t
#include "common.cuh"
#include "tc_common.cuh"


namespace {

constexpr int TR_ROWS = 128;
constexpr int TR_WORKER_WARPS = 16;
constexpr int TR_THREADS = 32 * (2 + TR_WORKER_WARPS);
constexpr int TR_WTHREADS = 32 * TR_WORKER_WARPS;
constexpr uint32_t TR_A_BYTES = TR_ROWS * 128;          // one K chunk: 128 rows x 32 fp32
constexpr uint32_t TR_NLO = 2;                          // buffers for the tf32 "lo" part of a chunk

struct TcRdParams {
    const float2* tab;       // MODE 0/1: rowF [K = W][Mx];  MODE 2: colF [My][K = H]
    float* out;              // MODE 0: T [rows][Mx] complex | MODE 1: T' [img][2*Mx][H] | MODE 2: Xh [img][My][Mx] complex
    int64_t rows;            // GEMM rows: images * H (MODE 0/1), images * 2*Mx (MODE 2)
    uint32_t ntiles;
    int K, Mx, My, H, N, nvalid, nkc, stages;     // nvalid = 2*Mx (MODE 0/1) or 2*My (MODE 2) meaningful accumulator columns
    uint32_t idesc, tmem_cols;
};

// MODE 0: row DFT, interleaved complex output (public sb200_rowdft_fwd)
// MODE 1: row DFT, planar transposed output T' (first half of sb200_analysis)
// MODE 2: column DFT from T' (second half of sb200_analysis)
template <int PASSES, int MODE>
__global__ void __launch_bounds__(TR_THREADS, 1)
tc_rowdft_kernel(const __grid_constant__ CUtensorMap tmapX, const TcRdParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* base = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    const int S = p.stages, nkc = p.nkc, N = p.N;
    const uint32_t b_chunk = (uint32_t)N * 128;                       // N rows x 128 B per K chunk (N % 8 == 0)
    const uint32_t b_bytes = (uint32_t)nkc * b_chunk;
    // The ring of TMA destinations (S stages; the landed fp32 chunk is truncated in place to its tf32 "hi" part) is
    // decoupled from the tf32 "lo" parts, which only live from the split to the MMAs of the same chunk: TR_NLO buffers
    // are enough, and the shared memory saved buys ring depth = bytes in flight (what the strided-row TMA stream needs).
    const uint32_t stage_bytes = TR_A_BYTES;
    uint8_t* B_hi = base;
    uint8_t* B_lo = B_hi + b_bytes;
    uint8_t* A_st = B_lo + (PASSES == 3 ? b_bytes : 0);
    uint8_t* A_lo = A_st + (uint32_t)S * stage_bytes;                 // [TR_NLO][16 KB]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(A_lo + (PASSES == 3 ? TR_NLO * TR_A_BYTES : 0));
    uint64_t* split_bar = full_bar + S;
    uint64_t* empty_bar = split_bar + S;
    uint64_t* tfull_bar = empty_bar + S;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        tc::tma_prefetch_desc(&tmapX);
        for (int s = 0; s < S; ++s) {
            tc::mbar_init(full_bar + s, 1);
            tc::mbar_init(split_bar + s, TR_WORKER_WARPS);
            tc::mbar_init(empty_bar + s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            tc::mbar_init(tfull_bar + a, 1);
            tc::mbar_init(tempty_bar + a, TR_WORKER_WARPS);
        }
        tc::fence_barrier_init();
    }
    if (warp == 0) {
        __syncwarp();
        tc::tmem_alloc(tmem_slot, p.tmem_cols);
        tc::tmem_relinquish();
    }
    // resident twiddle operand B[n][k]; rows n >= nvalid are zero
    //   MODE 0/1: B[n][x] = rowF[x][n >> 1].{x, y}        MODE 2: B[n][y] = colF[n >> 1][y].{x, y}
    for (int idx = tid; idx < N * p.K; idx += TR_THREADS) {
        int x, n;
        if (MODE == 2) { n = idx / p.K; x = idx - n * p.K; }           // k fastest: a warp reads along one table row
        else { x = idx / N; n = idx - x * N; }                         // n fastest: likewise
        float v = 0.f;
        if (n < p.nvalid) {
            const float2 t = MODE == 2 ? __ldg(p.tab + (size_t)(n >> 1) * p.K + x) : __ldg(p.tab + (size_t)x * p.Mx + (n >> 1));
            v = (n & 1) ? t.y : t.x;
        }
        const float hi = tc::tf32_trunc(v);
        const uint32_t off = (uint32_t)(x >> 5) * b_chunk + tc::sw128_kmajor_off(n, x & 31);
        *reinterpret_cast<float*>(B_hi + off) = hi;
        if (PASSES == 3) *reinterpret_cast<float*>(B_lo + off) = tc::tf32_lo(v, hi);
    }
    tc::fence_proxy_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    const uint32_t first = blockIdx.x, stride = gridDim.x, ntiles = p.ntiles;
    const uint32_t my_tiles = first < ntiles ? (ntiles - first + stride - 1) / stride : 0;

    if (warp == 0) {
        if (tc::elect_one()) {
            uint32_t s = 0, ph = 0;
            for (uint32_t it = 0; it < my_tiles; ++it) {
                const int row0 = (int)((first + it * stride) * TR_ROWS);
                for (int kc = 0; kc < nkc; ++kc) {
                    tc::mbar_wait(empty_bar + s, ph ^ 1);
                    tc::mbar_expect_tx(full_bar + s, TR_A_BYTES);
                    tc::tma_load_2d(A_st + s * stage_bytes, &tmapX, kc * 32, row0, full_bar + s);
                    if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (tc::elect_one()) {
            const uint32_t hi32 = tc::desc_hi(1024, tc::LAYOUT_SW128);
            const uint32_t bh_base = tc::desc_lo(tc::smem_u32(B_hi), 16), bl_base = tc::desc_lo(tc::smem_u32(B_lo), 16);
            uint32_t s = 0, ph = 0, lo = 0;
            for (uint32_t it = 0; it < my_tiles; ++it) {
                const uint32_t a = it & 1, tround = it >> 1;
                tc::mbar_wait(tempty_bar + a, (tround & 1) ^ 1);
                tc::tc_fence_after_sync();
                const uint32_t tmem_d = tmem_base + a * (uint32_t)N;
                uint32_t started = 0;
                uint32_t boff = 0;
                for (int kc = 0; kc < nkc; ++kc) {
                    tc::mbar_wait((PASSES == 3 ? split_bar : full_bar) + s, ph);
                    tc::tc_fence_after_sync();
                    uint32_t ah = tc::desc_lo(tc::smem_u32(A_st + s * stage_bytes), 16);
                    uint32_t al = tc::desc_lo(tc::smem_u32(A_lo + lo * TR_A_BYTES), 16);
                    uint32_t bo = boff;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        tc::umma_tf32_lh(tmem_d, ah, hi32, bh_base + bo, hi32, p.idesc, started);
                        started = 1;
                        if (PASSES == 3) {
                            tc::umma_tf32_lh(tmem_d, al, hi32, bh_base + bo, hi32, p.idesc, 1u);
                            tc::umma_tf32_lh(tmem_d, ah, hi32, bl_base + bo, hi32, p.idesc, 1u);
                        }
                        ah += 32 >> 4; al += 32 >> 4; bo += 32 >> 4;
                    }
                    tc::umma_commit(empty_bar + s);                    // frees the stage AND the lo buffer of this chunk
                    boff += b_chunk >> 4;
                    if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
                    if (++lo == TR_NLO) lo = 0;
                }
                tc::umma_commit(tfull_bar + a);
            }
        }
    } else {
        const int wk = warp - 2, wtid = tid - 64;
        const int quarter = warp & 3;                       // TMEM lane quarter this warp may access
        const int cpart = wk >> 2;                          // 16-column slice of the accumulator it drains
        const bool drains = cpart * 16 < p.nvalid;
        const int c0 = cpart * 16;                          // first accumulator column of this warp
        const int npair = min(8, (p.nvalid - c0) >> 1);     // complex values (MODE 0) / ky values (MODE 2) in the slice
        uint32_t sp_s = 0, sp_ph = 0, sp_lo = 0;            // chunk g: stage, phase parity, lo buffer
        uint32_t lag_s = 0, lag_ph = 0, g = 0;             // chunk g - TR_NLO, the previous user of that lo buffer
        auto split_tile = [&]() {
            if (PASSES == 3) {
                for (int kc = 0; kc < nkc; ++kc) {
                    if (g >= TR_NLO) {
                        // lo[sp_lo] was last read by the MMAs of chunk g - TR_NLO; their commit is that chunk's empty phase.
                        // (That barrier cannot have moved a further phase on: its next use is chunk g - TR_NLO + S > g.)
                        tc::mbar_wait(empty_bar + lag_s, lag_ph);
                        if (++lag_s == (uint32_t)S) { lag_s = 0; lag_ph ^= 1; }
                    }
                    ++g;
                    tc::mbar_wait(full_bar + sp_s, sp_ph);
                    float4* ah = reinterpret_cast<float4*>(A_st + sp_s * stage_bytes);
                    float4* al = reinterpret_cast<float4*>(A_lo + sp_lo * TR_A_BYTES);
#pragma unroll
                    for (int j = 0; j < (int)(TR_A_BYTES / 16) / TR_WTHREADS; ++j) {
                        const int idx = wtid + j * TR_WTHREADS;
                        const float4 v = ah[idx];
                        const float4 h = make_float4(tc::tf32_trunc(v.x), tc::tf32_trunc(v.y), tc::tf32_trunc(v.z), tc::tf32_trunc(v.w));
                        ah[idx] = h;
                        al[idx] = tc::tf32_lo4(v, h);
                    }
                    tc::fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(split_bar + sp_s);
                    if (++sp_s == (uint32_t)S) { sp_s = 0; sp_ph ^= 1; }
                    if (++sp_lo == TR_NLO) sp_lo = 0;
                }
            }
        };
        if (my_tiles > 0) split_tile();
        for (uint32_t it = 0; it < my_tiles; ++it) {
            if (it + 1 < my_tiles) split_tile();
            const uint32_t a = it & 1, tround = it >> 1;
            tc::mbar_wait(tfull_bar + a, tround & 1);
            tc::tc_fence_after_sync();
            if (drains) {
                uint32_t r[16];
                tc::tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(quarter * 32) << 16) + a * (uint32_t)N + (uint32_t)(cpart * 16), r);
                tc::tmem_ld_wait();
                const int64_t row = (int64_t)(first + it * stride) * TR_ROWS + quarter * 32 + lane;
                if (MODE == 0) {
                    if (row < p.rows) {
                        float2* dst = reinterpret_cast<float2*>(p.out) + row * p.Mx + (c0 >> 1);
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (j < npair) dst[j] = make_float2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
                    }
                } else if (MODE == 1) {
                    // T'[img][n][y]: the 32 lanes of a warp are 32 consecutive y of one image (32 | H): full-line stores
                    if (row < p.rows) {
                        const int64_t img = row / p.H;
                        const int y = (int)(row - img * p.H);
                        float* dst = p.out + (img * p.nvalid + c0) * p.H + y;
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (c0 + j < p.nvalid) dst[(int64_t)j * p.H] = __uint_as_float(r[j]);
                    }
                } else {
                    // rows (img, kx, 0) and (img, kx, 1) are lanes 2l and 2l + 1 (2*Mx and the tile base are even)
                    const int64_t img = row / (2 * p.Mx);
                    const int n = (int)(row - img * (2 * p.Mx));
                    const bool writer = !(lane & 1) && row < p.rows;
                    float2* dst = reinterpret_cast<float2*>(p.out) + (img * p.My + (c0 >> 1)) * p.Mx + (n >> 1);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float own_r = __uint_as_float(r[2 * j]), own_i = __uint_as_float(r[2 * j + 1]);
                        const float oth_r = __shfl_xor_sync(0xffffffffu, own_r, 1), oth_i = __shfl_xor_sync(0xffffffffu, own_i, 1);
                        if (writer && j < npair) dst[(int64_t)j * p.Mx] = make_float2(own_r - oth_i, oth_r + own_i);
                    }
                }
            }
            tc::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(tempty_bar + a);
        }
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, p.tmem_cols);
}

}  // namespace

// geometry both stages need: K chunks of 32 fp32, twiddle operand + at least a 2-stage ring inside 227 KB
static bool tr_fits(int K, int nvalid, int passes, int* N_out, int* stages_out, size_t* smem_out) {
    if (K % 32 != 0 || K < 64 || K > 2048 || nvalid < 2 || (nvalid & 1)) return false;
    const int N = (nvalid + 15) / 16 * 16;
    if (N > 64) return false;
    const size_t mult = passes == 3 ? 2 : 1;
    const size_t b_bytes = (size_t)(K / 32) * N * 128 * mult;
    const size_t a_stage = (size_t)TR_A_BYTES;
    const size_t fixed = 1024 + b_bytes + (passes == 3 ? TR_NLO * TR_A_BYTES : 0) + 512;
    int stages = 8;
    while (stages > 2 && fixed + stages * a_stage > 224 * 1024) --stages;
    if (fixed + stages * a_stage > 227 * 1024) return false;
    *N_out = N; *stages_out = stages; *smem_out = fixed + stages * a_stage;
    return true;
}

template <int MODE>
static int tr_launch(const float* in, const float2* tab, float* out, int64_t rows, int K, int nvalid, int Mx, int My, int H,
                     int passes, cudaStream_t st) {
    TcRdParams p;
    size_t smem = 0;
    if (!tr_fits(K, nvalid, passes, &p.N, &p.stages, &smem)) {
        sb200_set_error("tc_rowdft: internal: geometry K=%d nvalid=%d not supported", K, nvalid);
        return 2;
    }
    p.tab = tab; p.out = out; p.rows = rows;
    p.ntiles = (uint32_t)((rows + TR_ROWS - 1) / TR_ROWS);
    p.K = K; p.Mx = Mx; p.My = My; p.H = H; p.nvalid = nvalid; p.nkc = K / 32;
    p.idesc = tc::make_idesc_tf32(128, p.N, 0, 0);
    uint32_t cols = 32;
    while (cols < (uint32_t)(2 * p.N)) cols <<= 1;
    p.tmem_cols = cols;
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    if (int rc = sb200_make_tmap_2d_f32(&tmap, in, (uint64_t)K, (uint64_t)rows, (uint64_t)K * 4, 32, TR_ROWS, 1)) return rc;
    const unsigned nsm = (unsigned)sb200_num_sms();
    const unsigned grid = p.ntiles < nsm ? p.ntiles : nsm;
    if (passes == 3) {
        SB_CHECK_CUDA(cudaFuncSetAttribute(tc_rowdft_kernel<3, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sb_launch(tc_rowdft_kernel<3, MODE>, grid, TR_THREADS, smem, st, tmap, p);
    } else {
        SB_CHECK_CUDA(cudaFuncSetAttribute(tc_rowdft_kernel<1, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sb_launch(tc_rowdft_kernel<1, MODE>, grid, TR_THREADS, smem, st, tmap, p);
    }
    SB_LAUNCH_CHECK();
    return 0;
}

static bool tr_enabled() {
    static const bool disabled = sb_env_flag("SB200_TC_ROWDFT_OFF");      // experiments: force the FFMA kernels
    return !disabled && sb_tc_mode() != 0;
}

// public row stage (sb200_rowdft_fwd): x [rows][W] -> T [rows][Mx] interleaved complex
int sb200_tc_rowdft_fwd(sb200_plan_t plan, int pass, const float* x, float* T, int64_t rows, cudaStream_t st,
                        int* handled) {
    *handled = 0;
    if (!tr_enabled()) return 0;
    const int passes = sb_tc_mode();
    int N, stages;
    size_t smem;
    if (!tr_fits(plan->W, 2 * plan->Mx, passes, &N, &stages, &smem)) return 0;
    if ((reinterpret_cast<uintptr_t>(x) & 15) != 0 || (reinterpret_cast<uintptr_t>(T) & 7) != 0) return 0;
    if (rows >= (1LL << 31) - TR_ROWS) return 0;
    if (int rc = tr_launch<0>(x, plan->rowF[pass], T, rows, plan->W, 2 * plan->Mx, plan->Mx, plan->My, plan->H, passes, st))
        return rc;
    *handled = 1;
    return 0;
}

// both stages of the analysis on the tensor cores: x [nimg][H][W] -> scratch T' [nimg][2*Mx][H] -> Xh [nimg][My][Mx];
// scratch holds sb200_analysis_scratch() floats (the same 2*nimg*H*Mx as the interleaved intermediate)
int sb200_tc_analysis(sb200_plan_t plan, int pass, const float* x, float* Xh, int64_t nimg, float* scratch, cudaStream_t st,
                      int* handled) {
    *handled = 0;
    if (!tr_enabled() || scratch == nullptr) return 0;
    static const bool col_off = sb_env_flag("SB200_TC_COLDFT_OFF");       // experiments: tensor-core row stage only
    if (col_off) return 0;
    const int passes = sb_tc_mode();
    const int H = plan->H, W = plan->W, Mx = plan->Mx, My = plan->My;
    int N, stages;
    size_t smem;
    if (!tr_fits(W, 2 * Mx, passes, &N, &stages, &smem) || !tr_fits(H, 2 * My, passes, &N, &stages, &smem)) return 0;
    if ((reinterpret_cast<uintptr_t>(x) & 15) != 0 || (reinterpret_cast<uintptr_t>(scratch) & 15) != 0 ||
        (reinterpret_cast<uintptr_t>(Xh) & 7) != 0)
        return 0;
    if (nimg * H >= (1LL << 31) - TR_ROWS || nimg * 2 * Mx >= (1LL << 31) - TR_ROWS) return 0;
    if (int rc = tr_launch<1>(x, plan->rowF[pass], scratch, nimg * H, W, 2 * Mx, Mx, My, H, passes, st)) return rc;
    if (int rc = tr_launch<2>(scratch, plan->colF[pass], Xh, nimg * 2 * Mx, H, 2 * My, Mx, My, H, passes, st)) return rc;
    *handled = 1;
    return 0;
}
