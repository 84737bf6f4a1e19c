// General fp32 GEMM on the tcgen05 tensor cores (3xTF32 split for fp32 parity, or one TF32 pass):
//
//      D[m, n] = sum_k A(m, k) * B(n, k)      (+ bias[n], GELU / GELU'(aux) / residual in the epilogue)
//
// Either operand may be K-major (memory [rows][K], the contraction index contiguous: activations [tokens][C],
// nn.Linear weights [out][in]) or MN-major (memory [K][rows]: the same arrays used transposed -- data gradients
// need W^T, weight gradients contract over the token axis of two [tokens][C] arrays).  Both arrive by TMA exactly as
// they lie in HBM: K-major tiles as {32 k, rows} boxes with the 128B swizzle (UMMA SWIZZLE_128B, K-major), MN-major
// tiles as {32 rows, 32 k} boxes with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B (UMMA SWIZZLE_128B_BASE32B, the only
// MN-major layout tcgen05 accepts for 32-bit operands).  No operand is ever transposed or copied in HBM.
//
// This is the FourCastNet token path of the reference (src/dlwpbench/models/fourcastnet/fourcastnet.py:42-57 ``Mlp``
// = Linear -> GELU -> Linear, :283-293 PatchEmbed, :257 head) and the backward of those layers.
//
//   persistent CTAs over work units (m-tile 128, n-tile BN <= 128, k-split);
//   warp 0 TMA producer | warp 1 MMA issuer | warps 2-9: tf32 hi/lo split (+ optional GELU transform) of the landed
//   chunks | warps 10-17: epilogue (overlaps the next unit's MMAs through the double-buffered accumulators)
//   TMEM: two accumulator buffers; in the 3-pass mode each buffer holds TWO accumulators -- hi*hi products in
//   one, the 2^-11 times smaller lo*hi + hi*lo corrections in the other -- so the chain of (truncating) tensor-core
//   accumulations that the large terms go through is K/8 long instead of 3K/8; the epilogue adds the two.
//   split-K (weight gradients: K = tokens): partial tiles go to a workspace, a second kernel reduces them in a
//   fixed order (deterministic).
#include "common.cuh"
#include "tc_common.cuh"


namespace {

constexpr int TG_SPLIT_WARPS = 8;              // operand split (+ transform) only
constexpr int TG_EPI_WARPS = 8;                // epilogue only: two per TMEM lane quarter
constexpr int TG_THREADS = 32 * (2 + TG_SPLIT_WARPS + TG_EPI_WARPS);
constexpr int TG_SPLIT_THREADS = 32 * TG_SPLIT_WARPS;
constexpr uint32_t TG_A_BYTES = 128 * 128;     // one K chunk of the A tile: 128 rows x 32 fp32 (either major)
constexpr uint32_t TG_NLO = 2;                 // buffers for the tf32 "lo" parts (live from the split to the MMAs)

struct TgParams {
    int M, N, K;
    int BN;                       // n-tile width (multiple of 16; of 32 when B is MN-major)
    int a_mn, b_mn;               // 1 = MN-major operand
    int mtiles, ntiles, nsplit;
    int kc_total;                 // K chunks of 32 (last one zero-filled by TMA)
    int kc_per_split;
    uint32_t nunits;
    int stages;
    int nlo;                      // depth of the ring of tf32 "lo" buffers: 2 (decoupled from the stages), or `stages` with b_pre
    int b_pre;                    // the B operand arrives pre-split: hi plane through tmapB, lo plane through tmapBlo (weights that
                                  // every m-tile re-reads are split ONCE by a tiny kernel instead of once per tile by the CTAs)
    int R;                        // rotating "main" accumulators per TMEM buffer (chain of truncating accumulations = K/(8R))
    // direct epilogue (nsplit == 1)
    float* D; int64_t ldd;
    const float* bias;            // [N] or NULL
    int act;                      // 0 none | 1 GELU | 2 multiply by GELU'(aux[m][n]) | 3 ReLU | 4 soft-shrink(lam) |
                                  // 5 multiply by (aux > 0) | 6 multiply by (aux != 0)   (ReLU / soft-shrink backward masks)
    const float* aux; int64_t ld_aux;
    const float* resid; int64_t ld_res; int res_rows;   // residual row = m % res_rows when res_rows > 0 (pos_embed broadcast)
    float* zout; int64_t ld_z;    // pre-activation store or NULL
    // batch of independent GEMMs of identical geometry (block-diagonal layers): batch bi shifts the TMA coordinates of the
    // operands and the epilogue pointers; nothing is copied or gathered
    int nbatch; uint32_t units_per_batch;
    int a_off;                    // A: added to the dim-0 coordinate (K-major: k, MN-major: m) per batch
    int b_off0, b_off1;           // B: added to the dim-0 / dim-1 coordinates per batch
    int64_t d_off, bias_off;      // elements added to D / zout / aux / resid pointers, and to bias, per batch
    float lam;                    // soft-shrink threshold (act 4)
    int a_xform, b_xform;         // 1: the operand is GELU(what lies in HBM), applied in the split pass (h = GELU(z) never stored)
    int vec_ok;                   // every epilogue pointer / leading dimension allows 16-byte accesses
    // split-K epilogue (nsplit > 1): ws[split][M][N]
    float* ws;
    uint32_t idesc, tmem_cols;
    long long* trace;             // bring-up builds only: clock64 stamps of CTA 0 [role][chunk or unit][event]
};
#ifdef SB200_BRINGUP
#define TG_TRACE(role, i, ev) do { if (p.trace && blockIdx.x == 0 && (i) < 32) p.trace[((role) * 32 + (i)) * 4 + (ev)] = clock64(); } while (0)
#else
#define TG_TRACE(role, i, ev) do { } while (0)
#endif

template <int PASSES>
__global__ void __launch_bounds__(TG_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB,
               const __grid_constant__ CUtensorMap tmapBlo, const TgParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* base = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    const int S = p.stages, BN = p.BN;
    const uint32_t b_bytes = (uint32_t)BN * 128;
    const uint32_t stage_bytes = TG_A_BYTES + b_bytes;                 // [A chunk | B chunk]
    uint8_t* St = base;
    uint8_t* Lo = St + (uint32_t)S * stage_bytes;                      // [nlo][stage_bytes]
    const uint32_t NLO = (uint32_t)p.nlo;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(Lo + (PASSES == 3 ? NLO * stage_bytes : 0));
    uint64_t* split_bar = full_bar + S;
    uint64_t* empty_bar = split_bar + S;
    uint64_t* tfull_bar = empty_bar + S;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        tc::tma_prefetch_desc(&tmapA);
        tc::tma_prefetch_desc(&tmapB);
        for (int s = 0; s < S; ++s) {
            tc::mbar_init(full_bar + s, 1);
            tc::mbar_init(split_bar + s, TG_SPLIT_WARPS);
            tc::mbar_init(empty_bar + s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            tc::mbar_init(tfull_bar + a, 1);
            tc::mbar_init(tempty_bar + a, TG_EPI_WARPS);
        }
        tc::fence_barrier_init();
    }
    if (warp == 0) {
        __syncwarp();
        tc::tmem_alloc(tmem_slot, p.tmem_cols);
        tc::tmem_relinquish();
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    const uint32_t first = blockIdx.x, stride = gridDim.x, nunits = p.nunits;
    const uint32_t my_units = first < nunits ? (nunits - first + stride - 1) / stride : 0;
    const uint32_t R = (uint32_t)p.R;
    const uint32_t acc_cols = (uint32_t)BN * (R + (PASSES == 3 ? 1u : 0u));   // TMEM columns of one buffer: [main x R | corr]

    // unit -> (m-tile, n-tile, k-split); n-tiles fastest so that concurrently running CTAs share the A rows in L2
    auto decode = [&](uint32_t u, int& m0, int& n0, int& kc0, int& kcn, uint32_t& split) {
        u %= p.units_per_batch;
        split = u % (uint32_t)p.nsplit;
        const uint32_t t = u / (uint32_t)p.nsplit;
        n0 = (int)(t % (uint32_t)p.ntiles) * BN;
        m0 = (int)(t / (uint32_t)p.ntiles) * 128;
        kc0 = (int)split * p.kc_per_split;
        kcn = min(p.kc_per_split, p.kc_total - kc0);
    };

    if (warp == 0) {
        // ================= TMA producer =================
        if (tc::elect_one()) {
            uint32_t s = 0, ph = 0, gc = 0;
            for (uint32_t it = 0; it < my_units; ++it) {
                int m0, n0, kc0, kcn; uint32_t split;
                decode(first + it * stride, m0, n0, kc0, kcn, split);
                const int bi = (int)((first + it * stride) / p.units_per_batch);
                // batch shifts: K-major operands carry k in dim 0, MN-major operands carry the m / n index in dim 0
                const int ak = p.a_mn ? 0 : bi * p.a_off, am = p.a_mn ? bi * p.a_off : 0;
                const int bk = p.b_mn ? bi * p.b_off1 : bi * p.b_off0, bn = p.b_mn ? bi * p.b_off0 : bi * p.b_off1;
                for (int kc = 0; kc < kcn; ++kc) {
                    const int k0 = (kc0 + kc) * 32;
                    TG_TRACE(0, gc, 0);
                    tc::mbar_wait(empty_bar + s, ph ^ 1);
                    TG_TRACE(0, gc, 1);
                    uint8_t* dA = St + s * stage_bytes;
                    uint8_t* dB = dA + TG_A_BYTES;
                    tc::mbar_expect_tx(full_bar + s, stage_bytes + (p.b_pre ? b_bytes : 0u));
                    if (p.a_mn) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) tc::tma_load_2d(dA + i * 4096, &tmapA, am + m0 + 32 * i, k0, full_bar + s);
                    } else {
                        tc::tma_load_2d(dA, &tmapA, ak + k0, m0, full_bar + s);
                    }
                    if (p.b_mn) {
                        for (int j = 0; j < BN / 32; ++j) tc::tma_load_2d(dB + j * 4096, &tmapB, bn + n0 + 32 * j, bk + k0, full_bar + s);
                    } else {
                        tc::tma_load_2d(dB, &tmapB, bk + k0, bn + n0, full_bar + s);
                    }
                    if (p.b_pre) {                          // lo plane of B straight into the lo slot of this stage (nlo == S)
                        uint8_t* dL = Lo + s * stage_bytes + TG_A_BYTES;
                        if (p.b_mn) {
                            for (int j = 0; j < BN / 32; ++j) tc::tma_load_2d(dL + j * 4096, &tmapBlo, n0 + 32 * j, k0, full_bar + s);
                        } else {
                            tc::tma_load_2d(dL, &tmapBlo, k0, n0, full_bar + s);
                        }
                    }
                    TG_TRACE(0, gc, 2); ++gc;
                    if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (tc::elect_one()) {
            // K-major tile: 8-row atoms of 1024 B (SBO), k-step = +32 B.  MN-major tile: 32-row (mn) blocks 4096 B apart
            // (LBO), 4-k-row atoms of 512 B (SBO), k-step (8 k rows) = +1024 B.
            const uint32_t a_hi32 = p.a_mn ? tc::desc_hi(512, tc::LAYOUT_SW128_BASE32B) : tc::desc_hi(1024, tc::LAYOUT_SW128);
            const uint32_t b_hi32 = p.b_mn ? tc::desc_hi(512, tc::LAYOUT_SW128_BASE32B) : tc::desc_hi(1024, tc::LAYOUT_SW128);
            const uint32_t a_lbo = p.a_mn ? 4096u : 16u, b_lbo = p.b_mn ? 4096u : 16u;
            const uint32_t a_step = (p.a_mn ? 1024u : 32u) >> 4, b_step = (p.b_mn ? 1024u : 32u) >> 4;
            uint32_t s = 0, ph = 0, lo = 0, gc = 0;
            for (uint32_t it = 0; it < my_units; ++it) {
                int m0, n0, kc0, kcn; uint32_t split;
                decode(first + it * stride, m0, n0, kc0, kcn, split);
                const uint32_t a = it & 1, tround = it >> 1;
                TG_TRACE(3, it, 0);
                tc::mbar_wait(tempty_bar + a, (tround & 1) ^ 1);
                TG_TRACE(3, it, 1);
                tc::tc_fence_after_sync();
                const uint32_t d_buf = tmem_base + a * acc_cols;
                const uint32_t d_corr = d_buf + R * (uint32_t)BN;
                uint32_t started_c = 0, ra = 0;
                for (int kc = 0; kc < kcn; ++kc) {
                    const uint32_t d_main = d_buf + ra * (uint32_t)BN;
                    uint32_t started = (uint32_t)kc >= R ? 1u : 0u;
                    TG_TRACE(1, gc, 0);
                    tc::mbar_wait(((PASSES == 3 || p.a_xform || p.b_xform) ? split_bar : full_bar) + s, ph);
                    TG_TRACE(1, gc, 1);
                    tc::tc_fence_after_sync();
                    const uint32_t sa = tc::smem_u32(St + s * stage_bytes);
                    const uint32_t la = tc::smem_u32(Lo + lo * stage_bytes);
                    uint32_t ah = tc::desc_lo(sa, a_lbo), bh = tc::desc_lo(sa + TG_A_BYTES, b_lbo);
                    uint32_t al = tc::desc_lo(la, a_lbo), bl = tc::desc_lo(la + TG_A_BYTES, b_lbo);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        tc::umma_tf32_lh(d_main, ah, a_hi32, bh, b_hi32, p.idesc, started);
                        if (PASSES == 3) {
                            tc::umma_tf32_lh(d_corr, al, a_hi32, bh, b_hi32, p.idesc, started_c);
                            tc::umma_tf32_lh(d_corr, ah, a_hi32, bl, b_hi32, p.idesc, 1u);
                        }
                        started = 1; started_c = 1;
                        ah += a_step; al += a_step; bh += b_step; bl += b_step;
                    }
                    tc::umma_commit(empty_bar + s);                    // frees the stage AND the lo buffer of this chunk
                    TG_TRACE(1, gc, 2); ++gc;
                    if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
                    if (++lo == NLO) lo = 0;
                    if (++ra == R) ra = 0;
                }
                tc::umma_commit(tfull_bar + a);
            }
        }
    } else if (warp < 2 + TG_SPLIT_WARPS) {
        // ================= splitter warps: tf32 hi/lo split (+ optional GELU of an operand) of every landed chunk =====
        // They never touch the epilogue, so the MMA pipeline is fed continuously while the epilogue warps drain the
        // previous unit's accumulators.
        if (PASSES == 3 || p.a_xform || p.b_xform) {
            const int wtid = tid - 64;
            uint32_t sp_s = 0, sp_ph = 0, sp_lo = 0;
            uint32_t lag_s = 0, lag_ph = 0, g = 0;
            const int nA4 = (int)(TG_A_BYTES / 16), n4 = (int)(stage_bytes / 16);
            for (uint32_t it = 0; it < my_units; ++it) {
                int m0, n0, kc0, kcn; uint32_t split;
                decode(first + it * stride, m0, n0, kc0, kcn, split);
                for (int kc = 0; kc < kcn; ++kc) {
                    if (wtid == 0) TG_TRACE(2, g, 0);
                    if (PASSES == 3 && !p.b_pre && g >= NLO) {
                        // lo[sp_lo] was last read by the MMAs of chunk g - TG_NLO: their commit is that chunk's empty phase
                        tc::mbar_wait(empty_bar + lag_s, lag_ph);
                        if (++lag_s == (uint32_t)S) { lag_s = 0; lag_ph ^= 1; }
                    }
                    if (wtid == 0) TG_TRACE(2, g, 1);
                    tc::mbar_wait(full_bar + sp_s, sp_ph);
                    if (wtid == 0) TG_TRACE(2, g, 2);
                    float4* ah = reinterpret_cast<float4*>(St + sp_s * stage_bytes);
                    float4* al = reinterpret_cast<float4*>(Lo + sp_lo * stage_bytes);
                    const int nsp = p.b_pre ? nA4 : n4;
                    for (int idx = wtid; idx < nsp; idx += TG_SPLIT_THREADS) {
                        float4 v = ah[idx];
                        const bool xf = idx < nA4 ? p.a_xform : p.b_xform;
                        if (xf) {                                      // operand = GELU(stored pre-activation): never in HBM
                            // (scalar form on purpose: with the packed gelu4 here the cfg4-width AFNO backward GEMMs of this same
                            //  kernel lose their correction terms -- tests/test_fourcastnet_gpu.py::test_cfg4_width_vs_oracle,
                            //  1e-3 instead of 5e-7 -- although this branch is not taken there; not understood, see NOTES.md)
                            v = make_float4(gelu_f(v.x), gelu_f(v.y), gelu_f(v.z), gelu_f(v.w));
                            if (PASSES != 3 || p.b_pre) ah[idx] = v;
                        }
                        if (PASSES == 3) {
                            // hi is NOT written back: the tensor core reads the fp32 word as tf32 by dropping its low 13 bits,
                            // i.e. hi = trunc(v); lo = RN_tf32(v - trunc(v)) makes hi + lo exact to 2^-22 |v| (the residual is
                            // rounded, so nothing is left for the hardware to truncate).  Saves a third of the split's
                            // shared-memory writes -- these kernels are bound by shared-memory bandwidth.
                            // (measured: faster where only A is split in the CTA, slower in the split-K weight-gradient GEMMs
                            // that split both operands -- those keep the rounded hi written back)
                            if (p.b_pre) {
                                const float4 h = make_float4(tc::tf32_cut(v.x), tc::tf32_cut(v.y), tc::tf32_cut(v.z), tc::tf32_cut(v.w));
                                al[idx] = tc::tf32_lo4(v, h);
                            } else {
                                const float4 h = make_float4(tc::tf32_rna(v.x), tc::tf32_rna(v.y), tc::tf32_rna(v.z), tc::tf32_rna(v.w));
                                ah[idx] = h;
                                al[idx] = tc::tf32_lo4(v, h);
                            }
                        }
                    }
                    tc::fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(split_bar + sp_s);
                    if (wtid == 0) TG_TRACE(2, g, 3);
                    ++g;
                    if (++sp_s == (uint32_t)S) { sp_s = 0; sp_ph ^= 1; }
                    if (++sp_lo == NLO) sp_lo = 0;
                }
            }
        }
    } else {
        // ================= epilogue warps: TMEM -> registers -> fused epilogue -> global =================
        const int wk = warp - 2 - TG_SPLIT_WARPS;          // 0..7
        const int quarter = warp & 3;                       // TMEM lane quarter this warp may access
        const int cpart = wk >> 2;                          // column half it drains
        const int ncol_part = (((BN + 1) / 2) + 15) & ~15;  // multiple of 16
        const int c_begin = cpart * ncol_part;
        const int c_end = min(BN, c_begin + ncol_part);
        for (uint32_t it = 0; it < my_units; ++it) {
            int m0, n0, kc0, kcn; uint32_t split;
            decode(first + it * stride, m0, n0, kc0, kcn, split);
            const uint32_t a = it & 1, tround = it >> 1;
            tc::mbar_wait(tfull_bar + a, tround & 1);
            if (wk == 0 && lane == 0) TG_TRACE(3, it, 2);
            tc::tc_fence_after_sync();
            const int m = m0 + quarter * 32 + lane;
            const bool row_ok = m < p.M;
            const int64_t bi = (int64_t)((first + it * stride) / p.units_per_batch);
            const int64_t eoff = bi * p.d_off;                 // D / zout / aux / resid shift of this batch
            const float* bias_b = p.bias ? p.bias + bi * p.bias_off : nullptr;
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + a * acc_cols;
            const uint32_t nacc = min(R, (uint32_t)kcn);     // main accumulators this unit actually wrote
            for (int c0 = c_begin; c0 < c_end; c0 += 16) {
                uint32_t r[16];
                float v[16];
                if (PASSES == 3) {
                    tc::tmem_ld_32x32b_x16(taddr + R * (uint32_t)BN + (uint32_t)c0, r);     // small corrections first
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = 0.f;
                }
                for (uint32_t q = 0; q < nacc; ++q) {
                    tc::tmem_ld_32x32b_x16(taddr + q * (uint32_t)BN + (uint32_t)c0, r);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] += __uint_as_float(r[j]);
                }
                const int n = n0 + c0;
                if (!row_ok || n >= p.N) continue;
                const bool full16 = n + 16 <= p.N;
                if (p.nsplit > 1) {
                    float* dst = p.ws + (((int64_t)split * p.nbatch + bi) * p.M + m) * p.N + n;
                    if (full16 && (p.N & 3) == 0) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) if (n + j < p.N) dst[j] = v[j];
                    }
                    continue;
                }
                const bool vec = full16 && p.vec_ok;          // every pointer / leading dimension is 16-byte aligned
                if (p.bias) {
                    if (vec) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            const float4 b = __ldg(reinterpret_cast<const float4*>(bias_b + n + j));
                            v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) if (n + j < p.N) v[j] += __ldg(bias_b + n + j);
                    }
                }
                if (p.zout) {
                    float* zd = p.zout + eoff + (int64_t)m * p.ld_z + n;
                    if (vec) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(zd + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) if (n + j < p.N) zd[j] = v[j];
                    }
                }
                if (p.act == 1) {
#pragma unroll
                    for (int j = 0; j < 16; j += 2) gelu2(v[j], v[j + 1]);
                } else if (p.act == 3) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
                } else if (p.act == 4) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = v[j] > p.lam ? v[j] - p.lam : (v[j] < -p.lam ? v[j] + p.lam : 0.f);
                } else if (p.act == 5 || p.act == 6) {
                    const float* ax = p.aux + eoff + (int64_t)m * p.ld_aux + n;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        if (n + j < p.N) {
                            const float q = __ldg(ax + j);
                            v[j] = (p.act == 5 ? q > 0.f : q != 0.f) ? v[j] : 0.f;
                        }
                    }
                } else if (p.act == 2) {
                    const float* ax = p.aux + eoff + (int64_t)m * p.ld_aux + n;
                    if (vec) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            const float4 z4 = __ldg(reinterpret_cast<const float4*>(ax + j));
                            const float4 gg = gelu_grad4(z4);
                            v[j] *= gg.x; v[j + 1] *= gg.y; v[j + 2] *= gg.z; v[j + 3] *= gg.w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) if (n + j < p.N) v[j] *= gelu_grad_f(__ldg(ax + j));
                    }
                }
                if (p.resid) {
                    const float* rs = p.resid + eoff + (int64_t)(p.res_rows > 0 ? m % p.res_rows : m) * p.ld_res + n;
                    if (vec) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            const float4 q4 = __ldg(reinterpret_cast<const float4*>(rs + j));
                            v[j] += q4.x; v[j + 1] += q4.y; v[j + 2] += q4.z; v[j + 3] += q4.w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) if (n + j < p.N) v[j] += __ldg(rs + j);
                    }
                }
                if (p.D) {
                    float* dst = p.D + eoff + (int64_t)m * p.ldd + n;
                    if (vec) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) if (n + j < p.N) dst[j] = v[j];
                    }
                }
            }
            tc::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(tempty_bar + a);
            if (wk == 0 && lane == 0) TG_TRACE(3, it, 3);
        }
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, p.tmem_cols);
}

// out[m][n] = sum_s ws[s][m][n]  (fixed order: deterministic)
__global__ void __launch_bounds__(256) tg_splitk_reduce_kernel(const float* __restrict__ ws, float* __restrict__ out, int64_t ldd,
                                                               int M, int N, int nsplit, int nbatch, int64_t d_off) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int64_t per = (int64_t)M * N;
    if (i >= per * nbatch) return;
    float s = 0.f;
    for (int k = 0; k < nsplit; ++k) s += ws[(int64_t)k * per * nbatch + i];
    const int64_t bi = i / per, r = i - bi * per;
    out[bi * d_off + (r / N) * ldd + (r % N)] = s;
}

// exact-fp32 CUDA-core GEMM (tc mode 0 and shapes / alignments the tensor-core kernel does not take): 64 x 64 tile,
// 4 x 4 outputs per thread, generic strides for both majors; same epilogue as the tensor-core kernel.
struct FgParams {
    const float* A; int64_t a_sm, a_sk;        // element (m,k) at A[m*a_sm + k*a_sk]
    const float* B; int64_t b_sn, b_sk;
    int M, N, K;
    float* D; int64_t ldd;
    const float* bias; int act;
    const float* aux; int64_t ld_aux;
    const float* resid; int64_t ld_res; int res_rows;
    float* zout; int64_t ld_z;
    int a_xform, b_xform;
    float lam;
};
__global__ void __launch_bounds__(256) ffma_gemm_kernel(const FgParams p) {
    __shared__ float As[16][65], Bs[16][65];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < p.K; k0 += 16) {
        for (int idx = threadIdx.x; idx < 64 * 16; idx += 256) {
            // consecutive threads walk the contiguous axis of each operand
            int r, k;
            if (p.a_sk == 1) { k = idx & 15; r = idx >> 4; } else { r = idx & 63; k = idx >> 6; }
            const int m = m0 + r, kk = k0 + k;
            float av = (m < p.M && kk < p.K) ? __ldg(p.A + (int64_t)m * p.a_sm + (int64_t)kk * p.a_sk) : 0.f;
            if (p.a_xform) av = gelu_f(av);
            As[k][r] = (m < p.M && kk < p.K) ? av : 0.f;
            if (p.b_sk == 1) { k = idx & 15; r = idx >> 4; } else { r = idx & 63; k = idx >> 6; }
            const int n = n0 + r, kb = k0 + k;
            float bv = (n < p.N && kb < p.K) ? __ldg(p.B + (int64_t)n * p.b_sn + (int64_t)kb * p.b_sk) : 0.f;
            if (p.b_xform) bv = gelu_f(bv);
            Bs[k][r] = (n < p.N && kb < p.K) ? bv : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = As[k][ty * 4 + i]; b[i] = Bs[k][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= p.M) continue;
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= p.N) continue;
            float v = acc[i][j];
            if (p.bias) v += __ldg(p.bias + n);
            if (p.zout) p.zout[(int64_t)m * p.ld_z + n] = v;
            if (p.act == 1) v = gelu_f(v);
            else if (p.act == 2) v *= gelu_grad_f(__ldg(p.aux + (int64_t)m * p.ld_aux + n));
            else if (p.act == 3) v = fmaxf(v, 0.f);
            else if (p.act == 4) v = v > p.lam ? v - p.lam : (v < -p.lam ? v + p.lam : 0.f);
            else if (p.act == 5) v = __ldg(p.aux + (int64_t)m * p.ld_aux + n) > 0.f ? v : 0.f;
            else if (p.act == 6) v = __ldg(p.aux + (int64_t)m * p.ld_aux + n) != 0.f ? v : 0.f;
            if (p.resid) v += __ldg(p.resid + (int64_t)(p.res_rows > 0 ? m % p.res_rows : m) * p.ld_res + n);
            if (p.D) p.D[(int64_t)m * p.ldd + n] = v;
        }
    }
}

// what the tensor-core kernel needs from an operand: 16-byte aligned base and row stride, at least one full box row
bool tg_operand_ok(const float* ptr, int64_t ld, int mn_major, int rows, int K) {
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ld & 3) != 0) return false;
    (void)mn_major; (void)rows; (void)K;
    return true;
}

struct TgGeom { int BN, R, ntiles, mtiles, nsplit, kc_total, kc_per_split, stages, nlo, b_pre; size_t smem; };

bool tg_geometry(int M, int N, int K, int b_mn, int passes, int want_split, int b_xform, TgGeom* g) {
    if (M <= 0 || N <= 0 || K <= 0) return false;
    g->mtiles = (M + 127) / 128;
    g->kc_total = (K + 31) / 32;
    // The tensor core's accumulation truncates, so the error of an accumulator grows with the number of MMAs chained
    // into it (measured: 2.4e-6 at K = 1024 in one chain).  Policy: at most ~12 K-chunks (48 MMAs) per main accumulator;
    // long contractions use 64-wide n-tiles, which leaves TMEM room for 3 rotating main accumulators per buffer.
    int BN = N >= 128 ? 128 : (N + 15) / 16 * 16;
    int nsplit = 1;
    int kc_unit = g->kc_total;
    if (want_split) {
        const int tiles = ((N + 127) / 128) * g->mtiles;
        const int sms = sb200_num_sms();
        nsplit = (sms + tiles - 1) / tiles;                      // fill the machine ...
        if (nsplit > g->kc_total / 8) nsplit = g->kc_total / 8;  // ... keeping at least 8 chunks per split
        if (nsplit < 1) nsplit = 1;
        kc_unit = (g->kc_total + nsplit - 1) / nsplit;
    }
    if (kc_unit > 12 && BN > 64) BN = 64;
    if (b_mn) BN = (BN + 31) / 32 * 32;
    if (BN > 128) BN = 128;
    g->BN = BN;
    g->ntiles = (N + BN - 1) / BN;
    int R = 512 / (2 * BN) - (passes == 3 ? 1 : 0);
    if (R > 4) R = 4;
    if (R < 1) R = 1;
    g->R = R;
    if (want_split && kc_unit > 12 * R) {                        // still too long: more splits (more partial traffic, less error)
        nsplit = (g->kc_total + 12 * R - 1) / (12 * R);
    }
    g->kc_per_split = (g->kc_total + nsplit - 1) / nsplit;
    g->nsplit = (g->kc_total + g->kc_per_split - 1) / g->kc_per_split;
    const size_t stage = TG_A_BYTES + (size_t)BN * 128;
    // pre-split B when it is re-read by many m-tiles (weights of the forward / data-gradient GEMMs)
    g->b_pre = (passes == 3 && !want_split && !b_xform && g->mtiles >= 8) ? 1 : 0;
    if (g->b_pre) {
        int stages = 4;                                          // every stage owns a lo slot
        while (stages > 2 && 1024 + 512 + (size_t)stages * 2 * stage > 220 * 1024) --stages;
        if (1024 + 512 + (size_t)stages * 2 * stage > 227 * 1024) { g->b_pre = 0; }
        else { g->stages = stages; g->nlo = stages; g->smem = 1024 + 512 + (size_t)stages * 2 * stage; return true; }
    }
    const size_t fixed = 1024 + (passes == 3 ? TG_NLO * stage : 0) + 512;
    int stages = 6;
    while (stages > 2 && fixed + stages * stage > 220 * 1024) --stages;
    if (fixed + stages * stage > 227 * 1024) return false;
    g->stages = stages;
    g->nlo = TG_NLO;
    g->smem = fixed + stages * stage;
    return true;
}

}  // namespace

// B -> [hi plane | lo plane] (same layout twice)
__global__ void __launch_bounds__(256) tg_presplit_kernel(const float* __restrict__ b, float* __restrict__ out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const float v = __ldg(b + i);
    const float h = tc::tf32_rna(v);
    out[i] = h;
    out[n + i] = tc::tf32_lo(v, h);
}

namespace {

struct TgCall {
    const float* A; int64_t lda; int a_mn; int64_t a_ext0, a_ext1;     // tensor-map extents: dim 0 (contiguous), dim 1
    const float* B; int64_t ldb; int b_mn; int64_t b_ext0, b_ext1;
    float* D; int64_t ldd;
    int M, N, K;
    const float* bias; int act; const float* aux; int64_t ld_aux;
    const float* resid; int64_t ld_res; int res_rows;
    float* zout; int64_t ld_z;
    int a_xform, b_xform, split_k;
    int nbatch, a_off, b_off0, b_off1; int64_t d_off, bias_off;
    float lam;
};

int64_t tg_workspace(int M, int N, int K, int b_mn, int split_k, int nbatch) {
    if (sb_tc_mode() == 0) return 0;
    const int passes = sb_tc_mode() == 1 ? 1 : 3;
    TgGeom g;
    if (!split_k) {
        if (passes != 3 || nbatch > 1) return 0;
        if (!tg_geometry(M, N, K, b_mn, 3, 0, 0, &g) || !g.b_pre) return 0;
        return 2 * (int64_t)N * K;                           // pre-split copy of B (hi plane, lo plane)
    }
    if (!tg_geometry(M, N, K, b_mn, passes, 1, 0, &g)) return 0;
    return g.nsplit > 1 ? (int64_t)g.nsplit * nbatch * M * N : 0;
}

int tg_run(const TgCall& c, float* workspace, cudaStream_t st) {
    SB_REQUIRE(c.A && c.B && (c.D || c.zout), "gemm: NULL operand");
    SB_REQUIRE(c.D || !c.split_k, "gemm: split-K needs D");
    SB_REQUIRE(c.M > 0 && c.N > 0 && c.K > 0 && c.nbatch > 0, "gemm: non-positive size");
    SB_REQUIRE(c.act >= 0 && c.act <= 6, "gemm: act must be 0..6");
    SB_REQUIRE(!(c.act == 2 || c.act == 5 || c.act == 6) || c.aux, "gemm: this activation needs aux");
    SB_REQUIRE(!c.split_k || (!c.bias && c.act == 0 && !c.resid && !c.zout), "gemm: split-K has no fused epilogue");
    const int M = c.M, N = c.N, K = c.K;
    const int mode = sb_tc_mode();
    TgGeom g;
    const int passes = mode == 1 ? 1 : 3;
    bool tc_ok = mode != 0 && tg_operand_ok(c.A, c.lda, c.a_mn, M, K) && tg_operand_ok(c.B, c.ldb, c.b_mn, N, K) &&
                 tg_geometry(M, N, K, c.b_mn, passes, c.split_k, c.b_xform, &g);
    if (tc_ok && c.nbatch > 1 && ((c.a_off | c.b_off0 | c.b_off1) & 3)) tc_ok = false;     // TMA coordinates, keep 16-byte granules
    if (tc_ok && g.nsplit > 1 && workspace == nullptr) tc_ok = false;
    if (tc_ok && g.b_pre && (c.nbatch > 1 || workspace == nullptr || (c.ldb != (c.b_mn ? N : K)))) {
        // no pre-split copy (batched, no room, or a strided B): per-tile split instead
        g.b_pre = 0;
        const size_t stage = TG_A_BYTES + (size_t)g.BN * 128;
        const size_t fixed = 1024 + TG_NLO * stage + 512;
        int stages = 6;
        while (stages > 2 && fixed + stages * stage > 220 * 1024) --stages;
        g.stages = stages; g.nlo = TG_NLO; g.smem = fixed + stages * stage;
    }
    if (!tc_ok) {
        // exact-fp32 CUDA-core kernel, one launch per batch
        for (int bi = 0; bi < c.nbatch; ++bi) {
            FgParams f;
            f.A = c.A + (int64_t)bi * c.a_off; f.a_sm = c.a_mn ? 1 : c.lda; f.a_sk = c.a_mn ? c.lda : 1;
            f.B = c.B + (int64_t)bi * ((int64_t)c.b_off1 * c.ldb + c.b_off0); f.b_sn = c.b_mn ? 1 : c.ldb; f.b_sk = c.b_mn ? c.ldb : 1;
            f.M = M; f.N = N; f.K = K; f.D = c.D ? c.D + bi * c.d_off : nullptr; f.ldd = c.ldd;
            f.bias = c.bias ? c.bias + bi * c.bias_off : nullptr; f.act = c.act;
            f.aux = c.aux ? c.aux + bi * c.d_off : nullptr; f.ld_aux = c.ld_aux;
            f.resid = c.resid ? c.resid + bi * c.d_off : nullptr; f.ld_res = c.ld_res; f.res_rows = c.res_rows;
            f.zout = c.zout ? c.zout + bi * c.d_off : nullptr; f.ld_z = c.ld_z;
            f.a_xform = c.a_xform; f.b_xform = c.b_xform; f.lam = c.lam;
            dim3 grid((N + 63) / 64, (M + 63) / 64);
            SB_REQUIRE(grid.y <= 65535, "gemm: M too large for the CUDA-core kernel");
            sb_launch(ffma_gemm_kernel, grid, 256, 0, st, f);
            SB_LAUNCH_CHECK();
        }
        return 0;
    }
    TgParams p;
    memset(&p, 0, sizeof(p));
    p.M = M; p.N = N; p.K = K; p.BN = g.BN; p.a_mn = c.a_mn; p.b_mn = c.b_mn;
    p.mtiles = g.mtiles; p.ntiles = g.ntiles; p.nsplit = g.nsplit; p.kc_total = g.kc_total; p.kc_per_split = g.kc_per_split;
    p.units_per_batch = (uint32_t)g.mtiles * (uint32_t)g.ntiles * (uint32_t)g.nsplit;
    p.nunits = p.units_per_batch * (uint32_t)c.nbatch;
    p.nbatch = c.nbatch; p.a_off = c.a_off; p.b_off0 = c.b_off0; p.b_off1 = c.b_off1; p.d_off = c.d_off; p.bias_off = c.bias_off;
    p.lam = c.lam;
    p.stages = g.stages; p.R = g.R; p.nlo = g.nlo; p.b_pre = g.b_pre;
    p.D = c.D; p.ldd = c.ldd; p.bias = c.bias; p.act = c.act; p.aux = c.aux; p.ld_aux = c.ld_aux; p.resid = c.resid; p.ld_res = c.ld_res;
    p.res_rows = c.res_rows; p.a_xform = c.a_xform; p.b_xform = c.b_xform;
    {
        auto al = [](const void* q, int64_t ld) { return q == nullptr || ((reinterpret_cast<uintptr_t>(q) & 15) == 0 && (ld & 3) == 0); };
        p.vec_ok = al(c.D, c.ldd) && al(c.bias, 0) && al(c.aux, c.ld_aux) && al(c.resid, c.ld_res) && al(c.zout, c.ld_z) &&
                   (c.d_off & 3) == 0 && (c.bias_off & 3) == 0;
    }
    p.zout = c.zout; p.ld_z = c.ld_z; p.ws = workspace;
    p.idesc = tc::make_idesc_tf32(128, g.BN, c.a_mn, c.b_mn);
    uint32_t cols = 32;
    while (cols < (uint32_t)(2 * g.BN * (g.R + (passes == 3 ? 1 : 0)))) cols <<= 1;
    p.tmem_cols = cols;
    CUtensorMap tmA, tmB, tmBlo;
    memset(&tmA, 0, sizeof(tmA));
    memset(&tmB, 0, sizeof(tmB));
    memset(&tmBlo, 0, sizeof(tmBlo));
    const float* Bhi = c.B;
    if (g.b_pre) {
        const int64_t nb = (int64_t)N * K;
        sb_launch(tg_presplit_kernel, (unsigned)ceil_div64(nb, 256), 256, 0, st, c.B, workspace, nb);
        SB_LAUNCH_CHECK();
        Bhi = workspace;
    }
    // K-major: dims {K extent, rows}, box {32, rows of the tile}.  MN-major: dims {rows extent, K extent}, box {32, 32}.
    if (c.a_mn) { if (int rc = sb200_make_tmap_2d_f32(&tmA, c.A, (uint64_t)c.a_ext0, (uint64_t)c.a_ext1, (uint64_t)c.lda * 4, 32, 32, 2)) return rc; }
    else        { if (int rc = sb200_make_tmap_2d_f32(&tmA, c.A, (uint64_t)c.a_ext0, (uint64_t)c.a_ext1, (uint64_t)c.lda * 4, 32, 128, 1)) return rc; }
    for (int pl = 0; pl < (g.b_pre ? 2 : 1); ++pl) {
        CUtensorMap* tm = pl ? &tmBlo : &tmB;
        const float* bp = Bhi + (pl ? (int64_t)N * K : 0);
        if (c.b_mn) { if (int rc = sb200_make_tmap_2d_f32(tm, bp, (uint64_t)c.b_ext0, (uint64_t)c.b_ext1, (uint64_t)c.ldb * 4, 32, 32, 2)) return rc; }
        else        { if (int rc = sb200_make_tmap_2d_f32(tm, bp, (uint64_t)c.b_ext0, (uint64_t)c.b_ext1, (uint64_t)c.ldb * 4, 32, (uint32_t)g.BN, 1)) return rc; }
    }
    if (!g.b_pre) tmBlo = tmB;
    const unsigned nsm = (unsigned)sb200_num_sms();
    const unsigned grid = p.nunits < nsm ? p.nunits : nsm;
#ifdef SB200_BRINGUP
    static long long* trace_dev = nullptr;
    const int want_trace = sb_env_int("SB200_TG_TRACE", 0);
    if (want_trace) {
        if (!trace_dev) cudaMalloc(&trace_dev, 4 * 32 * 4 * sizeof(long long));
        cudaMemsetAsync(trace_dev, 0, 4 * 32 * 4 * sizeof(long long), st);
        p.trace = trace_dev;
    }
#endif
    if (passes == 3) {
        SB_CHECK_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
        sb_launch(tc_gemm_kernel<3>, grid, TG_THREADS, g.smem, st, tmA, tmB, tmBlo, p);
    } else {
        SB_CHECK_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
        sb_launch(tc_gemm_kernel<1>, grid, TG_THREADS, g.smem, st, tmA, tmB, tmBlo, p);
    }
    SB_LAUNCH_CHECK();
#ifdef SB200_BRINGUP
    if (want_trace) {
        static int calls = 0;
        if (++calls == want_trace) {
            long long h[4 * 32 * 4];
            cudaStreamSynchronize(st);
            cudaMemcpy(h, p.trace, sizeof(h), cudaMemcpyDeviceToHost);
            long long t0 = h[(0 * 32 + 0) * 4 + 0];
            const char* names[4] = {"producer(wait0,wait1,issued)", "mma(wait0,wait1,committed)", "split(start,lag ok,full ok,done)", "unit(mma tempty0,tempty1; epi wake,done)"};
            for (int r = 0; r < 4; ++r) {
                printf("%s\n", names[r]);
                for (int i = 0; i < (r == 3 ? 4 : 26); ++i) {
                    printf("  %2d:", i);
                    for (int e = 0; e < 4; ++e) printf(" %7lld", h[(r * 32 + i) * 4 + e] ? h[(r * 32 + i) * 4 + e] - t0 : -1);
                    printf("\n");
                }
            }
        }
    }
#endif
    if (g.nsplit > 1) {
        sb_launch(tg_splitk_reduce_kernel, (unsigned)ceil_div64((int64_t)M * N * c.nbatch, 256), 256, 0, st, (const float*)workspace,
                  c.D, c.ldd, M, N, g.nsplit, c.nbatch, c.d_off);
        SB_LAUNCH_CHECK();
    }
    return 0;
}

}  // namespace

extern "C" int64_t sb200_gemm_workspace(int M, int N, int K, int b_mn, int split_k, int tc_mode) {
    SbModeScope _mode(tc_mode);
    return tg_workspace(M, N, K, b_mn, split_k, 1);
}

extern "C" int sb200_gemm(const float* A, int64_t lda, int a_mn, const float* B, int64_t ldb, int b_mn, float* D, int64_t ldd,
                          int M, int N, int K, const float* bias, int act, const float* aux, int64_t ld_aux,
                          const float* resid, int64_t ld_res, int res_rows, float* zout, int64_t ld_z, int a_xform,
                          int b_xform, int split_k, float* workspace, void* stream, int tc_mode) {
    SbModeScope _mode(tc_mode);
    SB_REQUIRE(act >= 0 && act <= 2, "gemm: act must be 0, 1 or 2");
    TgCall c;
    memset(&c, 0, sizeof(c));
    c.A = A; c.lda = lda; c.a_mn = a_mn; c.a_ext0 = a_mn ? M : K; c.a_ext1 = a_mn ? K : M;
    c.B = B; c.ldb = ldb; c.b_mn = b_mn; c.b_ext0 = b_mn ? N : K; c.b_ext1 = b_mn ? K : N;
    c.D = D; c.ldd = ldd; c.M = M; c.N = N; c.K = K; c.bias = bias; c.act = act; c.aux = aux; c.ld_aux = ld_aux;
    c.resid = resid; c.ld_res = ld_res; c.res_rows = res_rows; c.zout = zout; c.ld_z = ld_z;
    c.a_xform = a_xform; c.b_xform = b_xform; c.split_k = split_k; c.nbatch = 1;
    return tg_run(c, workspace, (cudaStream_t)stream);
}

extern "C" int64_t sb200_gemm_batched_workspace(const sb200_gemm_desc* d, int tc_mode) {
    SbModeScope _mode(tc_mode);
    if (!d) return 0;
    return tg_workspace(d->M, d->N, d->K, d->b_mn, d->split_k, d->nbatch);
}

extern "C" int sb200_gemm_batched(const sb200_gemm_desc* d, const float* A, const float* B, float* D, const float* bias,
                                  const float* aux, float* workspace, void* stream, int tc_mode) {
    SbModeScope _mode(tc_mode);
    SB_REQUIRE(d != nullptr, "gemm_batched: NULL descriptor");
    TgCall c;
    memset(&c, 0, sizeof(c));
    c.A = A; c.lda = d->lda; c.a_mn = d->a_mn; c.a_ext0 = d->a_ext0; c.a_ext1 = d->a_ext1;
    c.B = B; c.ldb = d->ldb; c.b_mn = d->b_mn; c.b_ext0 = d->b_ext0; c.b_ext1 = d->b_ext1;
    c.D = D; c.ldd = d->ldd; c.M = d->M; c.N = d->N; c.K = d->K; c.bias = bias; c.act = d->act; c.aux = aux; c.ld_aux = d->ldd;
    c.split_k = d->split_k; c.nbatch = d->nbatch; c.a_off = d->a_off; c.b_off0 = d->b_off0; c.b_off1 = d->b_off1;
    c.d_off = d->d_off; c.bias_off = d->bias_off; c.lam = d->lam;
    return tg_run(c, workspace, (cudaStream_t)stream);
}
