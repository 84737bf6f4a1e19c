// Truncated 2-D analysis (rfftn -> fftshift -> slice of neuralop SpectralConv.forward, and with the pass-1 tables the
// adjoint of irfftn) for small grids (H dividing 128, W a multiple of 32 up to 128) in ONE kernel:
//
//   row stage     T[128 rows, (kx,c)] = X[128 rows, W] . Bw[(kx,c), W]^T       tcgen05 GEMM against the twiddle matrix
//                 (north_star (a)); the image rows as they lie in HBM are the K-major A operand: TMA {32 x, 128 rows},
//                 128B swizzle; 3xTF32 with the twiddle operand stacked [hi | lo] along N, so that a k-step costs two
//                 MMAs  A_hi x [B_hi | B_lo]  and  A_lo x B_hi  instead of three (the tile's A operand is read twice, not
//                 three times, and the single issuing thread has a third fewer instructions to get out)
//   transpose     four warps read the accumulator from TMEM (lane = image row), add the hi*lo columns and store T to a
//                 double-buffered shared-memory tile [row][kx]
//   column stage  Xh[g][ky][kx] = sum_y colF[ky][y] T[g][y][kx]   exact fp32 FFMA2 from shared memory by five warps
//                 (74 k real MACs per tile against 295 k in the row stage: too small and too skinny -- N = G*2*Mx = 36 --
//                 for the tensor core: a chained second tcgen05 GEMM was built and measured first)
//
// T never goes to HBM: traffic = x once + Xh once.  PASSES = 1 is the single-pass TF32 mode.
//   warp 0 TMA producer | warp 1 MMA issuer | warps 2-5 tf32 split of landed X chunks | warps 6-9 transpose |
//   warps 10-14 column stage
#include "common.cuh"
#include "tc_common.cuh"


namespace {

constexpr int AT_ROWS = 128;
constexpr int AT_SPLIT_WARPS = 4, AT_TW_WARPS = 4, AT_COL_WARPS = 5;
constexpr int AT_COL_THREADS = 32 * AT_COL_WARPS;
constexpr int AT_THREADS = 32 * (2 + AT_SPLIT_WARPS + AT_TW_WARPS + AT_COL_WARPS);
constexpr uint32_t AT_X_BYTES = AT_ROWS * 128;     // one K chunk of the X tile (hi part; the lo part has its own ring slot)

struct AtParams {
    const float2* rowF;      // [W][Mx]
    const float2* colF;      // [My][H]
    float2* Xh;              // [nimg][My][Mx]
    int64_t nimg;
    uint32_t ntiles;
    int H, W, Mx, My, G;
    int NR;                  // 2*Mx rounded up to 16
    int nkr;                 // K chunks of 32: W/32
    int stages;
    uint32_t idesc_full, idesc_half, tmem_cols;
    int debug;               // bring-up builds only: 1 skip column math, 2 skip T writes, 4 skip MMAs, 8 skip the split math
};

template <int PASSES>
__global__ void __launch_bounds__(AT_THREADS, 1)
analysis_tc_kernel(const __grid_constant__ CUtensorMap tmapX, const AtParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* base = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    const int S = p.stages, NR = p.NR, nkr = p.nkr, H = p.H, Mx = p.Mx, My = p.My;
    const int NB = PASSES == 3 ? 2 * NR : NR;            // rows of the stacked twiddle operand [hi | lo]
    const int Myp = (My + 1) & ~1;
    const uint32_t bw_chunk = (uint32_t)NB * 128, bw_bytes = ((uint32_t)nkr * bw_chunk + 1023u) & ~1023u;
    const uint32_t slot_bytes = AT_X_BYTES * (PASSES == 3 ? 2 : 1);     // [hi | lo] of one chunk
    uint8_t* Bw = base;
    uint8_t* X_st = Bw + bw_bytes;
    float2* Tbuf = reinterpret_cast<float2*>(X_st + (uint32_t)S * slot_bytes);    // [2][128][Mx]
    float2* ctw = Tbuf + 2 * AT_ROWS * Mx;                                         // [H][Myp]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(ctw + (size_t)H * Myp);
    uint64_t* split_bar = full_bar + S;
    uint64_t* empty_bar = split_bar + S;
    uint64_t* dr_full = empty_bar + S;      // [2] accumulator complete
    uint64_t* dr_free = dr_full + 2;        // [2] transposers have read it
    uint64_t* t_full = dr_free + 2;         // [2] T tile written
    uint64_t* t_empty = t_full + 2;         // [2] T tile consumed by the column warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        tc::tma_prefetch_desc(&tmapX);
        for (int s = 0; s < S; ++s) {
            tc::mbar_init(full_bar + s, 1);
            tc::mbar_init(split_bar + s, AT_SPLIT_WARPS);
            tc::mbar_init(empty_bar + s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            tc::mbar_init(dr_full + a, 1);
            tc::mbar_init(dr_free + a, AT_TW_WARPS);
            tc::mbar_init(t_full + a, AT_TW_WARPS);
            tc::mbar_init(t_empty + a, AT_COL_WARPS);
        }
        tc::fence_barrier_init();
    }
    if (warp == 0) {
        __syncwarp();
        tc::tmem_alloc(tmem_slot, p.tmem_cols);
        tc::tmem_relinquish();
    }
    // resident operands (plan tables are immutable)
    //   Bw[n][x]: rows [0, NR) = hi(rowF[x][n >> 1].{x, y}), rows [NR, 2NR) = its lo part; K-major SW128 chunks of 32 x
    for (int idx = tid; idx < NR * p.W; idx += AT_THREADS) {
        const int x = idx / NR, n = idx - x * NR;
        float v = 0.f;
        if (n < 2 * Mx) {
            const float2 t = __ldg(p.rowF + (size_t)x * Mx + (n >> 1));
            v = (n & 1) ? t.y : t.x;
        }
        const float hi = tc::tf32_rna(v);
        const uint32_t cb = (uint32_t)(x >> 5) * bw_chunk;
        *reinterpret_cast<float*>(Bw + cb + tc::sw128_kmajor_off(n, x & 31)) = hi;
        if (PASSES == 3) *reinterpret_cast<float*>(Bw + cb + tc::sw128_kmajor_off(NR + n, x & 31)) = tc::tf32_lo(v, hi);
    }
    for (int idx = tid; idx < H * Myp; idx += AT_THREADS) {
        const int y = idx / Myp, k = idx - y * Myp;
        ctw[idx] = k < My ? __ldg(p.colF + (size_t)k * H + y) : make_float2(0.f, 0.f);
    }
    tc::fence_proxy_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    const uint32_t first = blockIdx.x, stride = gridDim.x, ntiles = p.ntiles;
    const uint32_t my_tiles = first < ntiles ? (ntiles - first + stride - 1) / stride : 0;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            uint32_t s = 0, ph = 0;
            for (uint32_t it = 0; it < my_tiles; ++it) {
                const int row0 = (int)((first + it * stride) * AT_ROWS);
                for (int kc = 0; kc < nkr; ++kc) {
                    tc::mbar_wait(empty_bar + s, ph ^ 1);
                    tc::mbar_expect_tx(full_bar + s, AT_X_BYTES);
                    tc::tma_load_2d(X_st + s * slot_bytes, &tmapX, kc * 32, row0, full_bar + s);
                    if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            const uint32_t hi32 = tc::desc_hi(1024, tc::LAYOUT_SW128);
            const uint32_t bw0 = tc::desc_lo(tc::smem_u32(Bw), 16);
            uint32_t s = 0, ph = 0;
            for (uint32_t t = 0; t < my_tiles; ++t) {
                const uint32_t a = t & 1;
                tc::mbar_wait(dr_free + a, ((t >> 1) & 1) ^ 1);
                tc::tc_fence_after_sync();
                const uint32_t d = tmem_base + a * (uint32_t)NB;
                uint32_t started = 0, bo = 0;
                for (int kc = 0; kc < nkr; ++kc) {
                    tc::mbar_wait((PASSES == 3 ? split_bar : full_bar) + s, ph);
                    tc::tc_fence_after_sync();
                    uint32_t ah = tc::desc_lo(tc::smem_u32(X_st + s * slot_bytes), 16);
                    uint32_t al = ah + (AT_X_BYTES >> 4);
                    uint32_t b = bw0 + bo;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        if (p.debug & 4) break;
                        // D[:, 0:NR] += A_hi B_hi (+ A_lo B_hi);  D[:, NR:2NR] += A_hi B_lo
                        tc::umma_tf32_lh(d, ah, hi32, b, hi32, p.idesc_full, started);
                        if (PASSES == 3) tc::umma_tf32_lh(d, al, hi32, b, hi32, p.idesc_half, 1u);
                        started = 1;
                        ah += 2; al += 2; b += 2;
                    }
                    tc::umma_commit(empty_bar + s);
                    bo += bw_chunk >> 4;
                    if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
                }
                tc::umma_commit(dr_full + a);
            }
        }
    } else if (warp < 2 + AT_SPLIT_WARPS) {
        // ================= X chunk split (hi in place, lo into the second half of the ring slot) =================
        if (PASSES == 3) {
            const int wtid = tid - 64;
            uint32_t sp_s = 0, sp_ph = 0;
            const uint32_t nchunks = my_tiles * (uint32_t)nkr;
            for (uint32_t c = 0; c < nchunks; ++c) {
                tc::mbar_wait_warp(full_bar + sp_s, sp_ph);
                float4* ah = reinterpret_cast<float4*>(X_st + sp_s * slot_bytes);
                float4* al = reinterpret_cast<float4*>(X_st + sp_s * slot_bytes + AT_X_BYTES);
#pragma unroll
                for (int j = 0; j < (int)(AT_X_BYTES / 16) / (32 * AT_SPLIT_WARPS); ++j) {
                    if (p.debug & 8) break;
                    const int idx = wtid + j * 32 * AT_SPLIT_WARPS;
                    const float4 v = ah[idx];
                    const float4 h = make_float4(tc::tf32_rna(v.x), tc::tf32_rna(v.y), tc::tf32_rna(v.z), tc::tf32_rna(v.w));
                    ah[idx] = h;
                    al[idx] = tc::tf32_lo4(v, h);
                }
                tc::fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(split_bar + sp_s);
                if (++sp_s == (uint32_t)S) { sp_s = 0; sp_ph ^= 1; }
            }
        }
    } else if (warp < 2 + AT_SPLIT_WARPS + AT_TW_WARPS) {
        // ================= transpose: TMEM accumulator -> T tile in shared memory =================
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        for (uint32_t t = 0; t < my_tiles; ++t) {
            const uint32_t a = t & 1, par = (t >> 1) & 1;
            tc::mbar_wait_warp(dr_full + a, par);
            tc::tc_fence_after_sync();
            float v[64];
#pragma unroll
            for (int c0 = 0; c0 < 64; c0 += 16) {
                if (c0 < NR) {
                    uint32_t r[16];
                    tc::tmem_ld_32x32b_x16(taddr + a * (uint32_t)NB + (uint32_t)c0, r);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[c0 + j] = __uint_as_float(r[j]);
                    if (PASSES == 3) {
                        tc::tmem_ld_32x32b_x16(taddr + a * (uint32_t)NB + (uint32_t)(NR + c0), r);
                        tc::tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[c0 + j] += __uint_as_float(r[j]);
                    }
                }
            }
            tc::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(dr_free + a);
            tc::mbar_wait_warp(t_empty + a, par ^ 1);          // the column warps are done with this T slot
            float2* Tb = Tbuf + (size_t)a * AT_ROWS * Mx + (size_t)row * Mx;
#pragma unroll
            for (int k = 0; k < 32; ++k)
                if (k < Mx && !(p.debug & 2)) Tb[k] = make_float2(v[2 * k], v[2 * k + 1]);
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(t_full + a);
        }
    } else {
        // ================= column stage (exact fp32 FFMA2): two ky per item =================
        const int ct = tid - 32 * (2 + AT_SPLIT_WARPS + AT_TW_WARPS);
        const int kyp_n = Myp / 2;
        const int items = p.G * kyp_n * Mx;
        for (uint32_t t = 0; t < my_tiles; ++t) {
            const uint32_t a = t & 1, par = (t >> 1) & 1;
            tc::mbar_wait_warp(t_full + a, par);
            const float2* Tb = Tbuf + (size_t)a * AT_ROWS * Mx;
            const int64_t img0 = (int64_t)(first + t * stride) * p.G;
            for (int item = ct; item < items && !(p.debug & 1); item += AT_COL_THREADS) {
                const int kx = item % Mx;
                const int rest = item / Mx;
                const int kyp = rest % kyp_n, g = rest / kyp_n;
                const float2* tp = Tb + (size_t)g * H * Mx + kx;
                const float4* twp = reinterpret_cast<const float4*>(ctw + kyp * 2);
                float2 acc0 = make_float2(0.f, 0.f), acc1 = acc0, acc2 = acc0, acc3 = acc0;   // two independent chains per ky
                const int wstep = Myp >> 1;
#pragma unroll 4
                for (int y = 0; y < H; y += 2) {
                    const float2 tv0 = tp[0], tv1 = tp[Mx];
                    const float4 w0 = twp[0], w1 = twp[wstep];
                    cmac2(acc0, make_float2(w0.x, w0.y), tv0);
                    cmac2(acc1, make_float2(w0.z, w0.w), tv0);
                    cmac2(acc2, make_float2(w1.x, w1.y), tv1);
                    cmac2(acc3, make_float2(w1.z, w1.w), tv1);
                    tp += 2 * Mx;
                    twp += 2 * wstep;
                }
                const int64_t img = img0 + g;
                if (img < p.nimg) {
                    const int ky = kyp * 2;
                    p.Xh[(img * My + ky) * Mx + kx] = make_float2(acc0.x + acc2.x, acc0.y + acc2.y);
                    if (ky + 1 < My) p.Xh[(img * My + ky + 1) * Mx + kx] = make_float2(acc1.x + acc3.x, acc1.y + acc3.y);
                }
            }
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(t_empty + a);
        }
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, p.tmem_cols);
}

struct AtGeom { int G, NR, stages; uint32_t tmem_cols; size_t smem; };

bool at_geometry(const sb200_plan_s* pl, int passes, AtGeom* g) {
    const int H = pl->H, W = pl->W, Mx = pl->Mx, My = pl->My;
    if (H < 8 || AT_ROWS % H != 0 || (H & 1)) return false;
    if (W % 32 != 0 || W > 128) return false;
    if (2 * Mx > 64 || Mx > 32) return false;
    g->G = AT_ROWS / H;
    g->NR = (2 * Mx + 15) / 16 * 16;
    const int NB = passes == 3 ? 2 * g->NR : g->NR;
    uint32_t cols = 32;
    while (cols < (uint32_t)(2 * NB)) cols <<= 1;
    g->tmem_cols = cols;
    const int Myp = (My + 1) & ~1;
    const size_t bw = (((size_t)(W / 32) * NB * 128) + 1023) & ~(size_t)1023;
    const size_t slot = (size_t)AT_X_BYTES * (passes == 3 ? 2 : 1);
    const size_t fixed = 1024 + bw + (size_t)2 * AT_ROWS * Mx * 8 + (size_t)H * Myp * 8 + 512;
    int stages = 6;
    while (stages > 2 && fixed + stages * slot > 220 * 1024) --stages;
    if (fixed + stages * slot > 227 * 1024) return false;
    g->stages = stages;
    g->smem = fixed + stages * slot;
    return true;
}

}  // namespace

bool sb200_analysis_tc_supported(sb200_plan_t plan) {
    const int mode = sb_tc_mode();
    if (mode == 0) return false;
    AtGeom g;
    return at_geometry(plan, mode == 1 ? 1 : 3, &g);
}

int sb200_analysis_tc(sb200_plan_t plan, int pass, const float* x, float* Xh, int64_t nimg, cudaStream_t st, int* handled) {
    *handled = 0;
    const int mode = sb_tc_mode();
    if (mode == 0) return 0;
    const int passes = mode == 1 ? 1 : 3;
    AtGeom g;
    if (!at_geometry(plan, passes, &g)) return 0;
    if ((reinterpret_cast<uintptr_t>(x) & 15) != 0 || (reinterpret_cast<uintptr_t>(Xh) & 7) != 0) return 0;
    const int64_t rows = nimg * plan->H;
    if (rows >= (1LL << 31) - AT_ROWS) return 0;
    AtParams p;
    memset(&p, 0, sizeof(p));
    p.rowF = plan->rowF[pass]; p.colF = plan->colF[pass]; p.Xh = reinterpret_cast<float2*>(Xh);
    p.nimg = nimg; p.ntiles = (uint32_t)((rows + AT_ROWS - 1) / AT_ROWS);
    p.H = plan->H; p.W = plan->W; p.Mx = plan->Mx; p.My = plan->My; p.G = g.G; p.NR = g.NR;
    p.nkr = plan->W / 32; p.stages = g.stages;
    p.idesc_full = tc::make_idesc_tf32(128, passes == 3 ? 2 * g.NR : g.NR, 0, 0);
    p.idesc_half = tc::make_idesc_tf32(128, g.NR, 0, 0);
    p.tmem_cols = g.tmem_cols;
    p.debug = sb_env_int("SB200_AT_DEBUG", 0);
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    if (int rc = sb200_make_tmap_2d_f32(&tmap, x, (uint64_t)plan->W, (uint64_t)rows, (uint64_t)plan->W * 4, 32, AT_ROWS, 1)) return rc;
    const unsigned nsm = (unsigned)sb200_num_sms();
    const unsigned grid = p.ntiles < nsm ? p.ntiles : nsm;
    if (passes == 3) {
        SB_CHECK_CUDA(cudaFuncSetAttribute(analysis_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
        sb_launch(analysis_tc_kernel<3>, grid, AT_THREADS, g.smem, st, tmap, p);
    } else {
        SB_CHECK_CUDA(cudaFuncSetAttribute(analysis_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
        sb_launch(analysis_tc_kernel<1>, grid, AT_THREADS, g.smem, st, tmap, p);
    }
    SB_LAUNCH_CHECK();
    *handled = 1;
    return 0;
}
