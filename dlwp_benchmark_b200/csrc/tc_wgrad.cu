// tcgen05 weight gradient of the pointwise (1x1) channel mix:
//       gW[o, i] = sum_{b,p} g[b, o, p] * x[b, i, p]
// Both operands are K-major as they lie in memory (the contraction index, the pixel, is the contiguous
// one), so TMA boxes {32 px, C channels} with the plain 128B swizzle are the UMMA operands directly.
//   A rows = channels of the tensor with more channels ("P"), 128 per M-block,
//   B rows = channels of the other tensor ("Q", UMMA N),  K = pixels, 8 per tcgen05.mma (tf32).
// Each persistent CTA owns a contiguous range of pixel chunks, accumulates its partial product in TMEM
// over the whole range and writes it to the workspace once; a second kernel reduces the partials
// in a fixed order (deterministic).  3xTF32 operand split as in tc_pointwise.cu.
#include "common.cuh"
#include "tc_common.cuh"


constexpr int TW_WORKER_WARPS = 16;
constexpr int TW_THREADS = 32 * (2 + TW_WORKER_WARPS);

#ifdef SB200_BRINGUP
#define TW_TRACE(role, i, ev) do { if (p.trace && blockIdx.x == 0 && (i) < 16) p.trace[((role) * 16 + (i)) * 4 + (ev)] = clock64(); } while (0)
#else
#define TW_TRACE(role, i, ev) do { } while (0)
#endif
struct TcWgParams {
    long long* trace;        // bring-up builds (SB200_WG_TRACE): clock64 log of CTA 0
    int Pc, Qc;              // channels of the A-side / B-side tensors
    int Qn;                  // UMMA N: Qc, or Qc + 16 when a constant ones-row is appended (bias gradient = sum of the A side)
    int mblocks;             // ceil(Pc / 128)
    int CH;                  // 32-pixel chunks per pipeline item
    int stages;
    int debug;               // bring-up switches (SB200_WG_DEBUG): 1 = skip the MMAs (measures the TMA ring alone)
    int R;                   // rotating TMEM accumulators (shortens each tensor-core accumulation chain)
    int64_t HW;
    int64_t items_per_b, nitems;
    float* ws;               // [grid][mblocks*128][Qc]
    uint32_t idesc, tmem_cols;
    // ASRC = 1: the A side is generated on chip, P[n, px] = GELU(w1[n] x[b, px] + b1[n]) (hidden layer of a 1-input-channel
    // lifting MLP), instead of being read from HBM
    const float* gen_x; const float* gen_w1; const float* gen_b1;
};

template <int PASSES, int ASRC = 0>
__global__ void __launch_bounds__(TW_THREADS, 1)
tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmapP, const __grid_constant__ CUtensorMap tmapQ, const TcWgParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment by OFFSET from the __shared__ array, so every derived pointer keeps the shared address
    // space (a uintptr_t round-trip turns all later accesses into generic LD/ST through L1TEX)
    uint8_t* base = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    const int S = p.stages, CH = p.CH;
    const uint32_t p_chunk = (uint32_t)p.Pc * 128, q_chunk = (uint32_t)p.Qc * 128;
    const uint32_t ones_bytes = (uint32_t)(p.Qn - p.Qc) * 128;         // constant rows [1,1,...; 0...] behind the Q rows
    const uint32_t chunk_bytes = p_chunk + q_chunk + ones_bytes;       // [P rows | Q rows | ones], 128 B per row
    const uint32_t tma_bytes = (uint32_t)CH * (ASRC ? q_chunk : p_chunk + q_chunk);
    const uint32_t raw_bytes = (uint32_t)CH * chunk_bytes;
    const uint32_t stage_bytes = raw_bytes * (PASSES == 3 ? 2 : 1);    // [hi | lo]
    uint8_t* St = base;
    uint8_t* tail = St + (uint32_t)S * stage_bytes + 16384;            // slack: a short last M-block reads past its rows
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
    uint64_t* split_bar = full_bar + S;
    uint64_t* empty_bar = split_bar + S;
    uint64_t* done_bar = empty_bar + S;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        tc::tma_prefetch_desc(&tmapP);
        tc::tma_prefetch_desc(&tmapQ);
        for (int s = 0; s < S; ++s) {
            tc::mbar_init(full_bar + s, 1);
            tc::mbar_init(split_bar + s, TW_WORKER_WARPS);
            tc::mbar_init(empty_bar + s, 1);
        }
        tc::mbar_init(done_bar, 1);
        tc::fence_barrier_init();
    }
    if (warp == 0) {
        __syncwarp();
        tc::tmem_alloc(tmem_slot, p.tmem_cols);
        tc::tmem_relinquish();
    }
    if (ones_bytes) {
        // row Qc of the B operand is all ones (hi part; its lo part and the other padding rows are zero) in every
        // chunk of every stage: D[:, Qc] = sum over pixels of the A side = the bias gradient.  TMA never writes here
        // and the split pass maps (1, 0) to itself.
        const int per_blk = (int)(ones_bytes / 4);
        const int nblk = S * CH * (PASSES == 3 ? 2 : 1);
        for (int idx = tid; idx < nblk * per_blk; idx += TW_THREADS) {
            const int blk = idx / per_blk, e = idx % per_blk;
            const int half = blk / (S * CH), sc = blk % (S * CH);
            uint8_t* dst = St + (uint32_t)(sc / CH) * stage_bytes + (uint32_t)half * raw_bytes + (uint32_t)(sc % CH) * chunk_bytes +
                           p_chunk + q_chunk;
            reinterpret_cast<float*>(dst)[e] = (half == 0 && e < 32) ? 1.0f : 0.0f;
        }
        tc::fence_proxy_async_smem();
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    // items of this CTA: blockIdx.x, blockIdx.x + grid, ... (interleaved: at any moment the CTAs of the grid work on
    // neighbouring 32-pixel chunks of the same channel rows, i.e. on the same DRAM pages, instead of 148 regions of their own)
    const int64_t i0 = blockIdx.x, istride = gridDim.x;
    const int64_t my_items = i0 < p.nitems ? (p.nitems - i0 + istride - 1) / istride : 0;

    if (warp == 0) {
        if (tc::elect_one()) {
            uint32_t s = 0, ph = 0;
            const uint32_t items_per_b = (uint32_t)p.items_per_b;
            for (uint32_t q = 0; q < (uint32_t)my_items; ++q) {
                const uint32_t item = (uint32_t)(i0 + (int64_t)q * istride);
                const uint32_t b = item / items_per_b;
                const int px0 = (int)((item - b * items_per_b) * 32u * (uint32_t)CH);
                tc::mbar_wait(empty_bar + s, ph ^ 1);
                uint8_t* dst = St + s * stage_bytes;
                tc::mbar_expect_tx(full_bar + s, tma_bytes);
                for (int j = 0; j < CH; ++j) {
                    if (ASRC == 0)
                        tc::tma_load_2d(dst + (uint32_t)j * chunk_bytes, &tmapP, px0 + 32 * j, (int)b * p.Pc, full_bar + s);
                    tc::tma_load_2d(dst + (uint32_t)j * chunk_bytes + p_chunk, &tmapQ, px0 + 32 * j, (int)b * p.Qc, full_bar + s);
                }
                if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // single-thread MMA issue: constant descriptor high word, low words advance by adds
        if (tc::elect_one()) {
            const uint32_t hi32 = tc::desc_hi(1024, tc::LAYOUT_SW128);
            const uint32_t lo_delta = raw_bytes >> 4;
            const uint32_t mb_step = (128u * 128u) >> 4;
            uint32_t s = 0, ph = 0, ra = 0;
            for (uint32_t q = 0; q < (uint32_t)my_items; ++q) {
                TW_TRACE(1, q, 0);
                tc::mbar_wait(((PASSES == 3 || ASRC == 1) ? split_bar : full_bar) + s, ph);
                TW_TRACE(1, q, 1);
                tc::tc_fence_after_sync();
                const uint32_t base_lo = tc::desc_lo(tc::smem_u32(St + s * stage_bytes), 16);
                const uint32_t dacc = tmem_base + ra * (uint32_t)(p.mblocks * p.Qn);
                const uint32_t first_acc = q >= (uint32_t)p.R ? 1u : 0u;
                if (p.debug & 1) {
                    tc::mbar_arrive(empty_bar + s);
                    if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
                    continue;
                }
                for (int j = 0; j < CH; ++j) {
                    uint32_t pj = base_lo + (((uint32_t)j * chunk_bytes) >> 4);
                    uint32_t qj = pj + (p_chunk >> 4);
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint32_t acc = (first_acc | (uint32_t)(j > 0) | (uint32_t)(ks > 0));
                        uint32_t ph_ = pj, d = dacc;
                        for (int mb = 0; mb < p.mblocks; ++mb) {
                            tc::umma_tf32_lh(d, ph_, hi32, qj, hi32, p.idesc, acc);
                            if (PASSES == 3) {
                                tc::umma_tf32_lh(d, ph_ + lo_delta, hi32, qj, hi32, p.idesc, 1u);
                                tc::umma_tf32_lh(d, ph_, hi32, qj + lo_delta, hi32, p.idesc, 1u);
                            }
                            ph_ += mb_step; d += (uint32_t)p.Qn;
                        }
                        pj += 32 >> 4; qj += 32 >> 4;
                    }
                }
                tc::umma_commit(empty_bar + s);
                TW_TRACE(1, q, 2);
                if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
                if (++ra == (uint32_t)p.R) ra = 0;
            }
            tc::umma_commit(done_bar);
        }
    } else {
        const int wk = warp - 2;
        const int wtid = tid - 64;
        if (ASRC == 1) {
            // thread -> 16-byte piece (4 px) `piece` of rows (wtid >> 3) + 64 i of every 32-px chunk
            const int piece = wtid & 7, row0 = wtid >> 3;
            const uint32_t items_per_b = (uint32_t)p.items_per_b;
            // the rows of this thread (row0 + 64 i, Pc <= 256) never change: their w1 / b1 live in registers, and the row loop
            // below is unrolled so that the four GELU chains of a chunk overlap (it was a serial loop with two global loads
            // in front of every chain: 5600 cycles per 32-pixel item against ~1500 of MUFU + issue time)
            float wv[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int row = row0 + 64 * i;
                wv[i] = row < p.Pc ? __ldg(p.gen_w1 + row) : 0.f;
                bv[i] = row < p.Pc ? __ldg(p.gen_b1 + row) : 0.f;
            }
            uint32_t s = 0, ph = 0;
            for (uint32_t q = 0; q < (uint32_t)my_items; ++q) {
                const uint32_t item = (uint32_t)(i0 + (int64_t)q * istride);
                const uint32_t b = item / items_per_b;
                const int64_t px0 = (int64_t)(item - b * items_per_b) * 32 * CH;
                if (wtid == 0) TW_TRACE(0, q, 0);
                tc::mbar_wait(empty_bar + s, ph ^ 1);                 // MMAs that read this stage are done
                if (wtid == 0) TW_TRACE(0, q, 1);
                uint8_t* sb = St + s * stage_bytes;
                for (int j = 0; j < CH; ++j) {
                    const int64_t px = px0 + 32 * j + piece * 4;
                    float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (px + 4 <= p.HW) xv = __ldg(reinterpret_cast<const float4*>(p.gen_x + (int64_t)b * p.HW + px));
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int row = row0 + 64 * i;
                        if (row >= p.Pc) break;
                        const float w = wv[i], bb = bv[i];
                        float4 g = gelu4(make_float4(fmaf(w, xv.x, bb), fmaf(w, xv.y, bb), fmaf(w, xv.z, bb), fmaf(w, xv.w, bb)));
                        if (px + 4 > p.HW) g = make_float4(0.f, 0.f, 0.f, 0.f);     // pixels outside the image contribute nothing
                        const float4 h = make_float4(tc::tf32_trunc(g.x), tc::tf32_trunc(g.y), tc::tf32_trunc(g.z), tc::tf32_trunc(g.w));
                        const uint32_t off = (uint32_t)j * chunk_bytes + (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((piece ^ (row & 7)) << 4));
                        *reinterpret_cast<float4*>(sb + off) = h;
                        if (PASSES == 3)
                            *reinterpret_cast<float4*>(sb + raw_bytes + off) = tc::tf32_lo4(g, h);
                    }
                }
                if (wtid == 0) TW_TRACE(0, q, 2);
                tc::mbar_wait(full_bar + s, ph);                       // Q side landed
                if (wtid == 0) TW_TRACE(0, q, 3);
                if (PASSES == 3) {
                    for (int j = 0; j < CH; ++j) {
                        float4* ah = reinterpret_cast<float4*>(sb + (uint32_t)j * chunk_bytes + p_chunk);
                        float4* al = reinterpret_cast<float4*>(sb + raw_bytes + (uint32_t)j * chunk_bytes + p_chunk);
                        for (int idx = wtid; idx < (int)(q_chunk / 16); idx += 32 * TW_WORKER_WARPS) {
                            const float4 v = ah[idx];
                            const float4 h = make_float4(tc::tf32_trunc(v.x), tc::tf32_trunc(v.y), tc::tf32_trunc(v.z), tc::tf32_trunc(v.w));
                            ah[idx] = h;
                            al[idx] = tc::tf32_lo4(v, h);
                        }
                    }
                }
                tc::fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(split_bar + s);
                if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
            }
        } else if (PASSES == 3) {
            uint32_t s = 0, ph = 0;
            for (uint32_t q = 0; q < (uint32_t)my_items; ++q) {
                tc::mbar_wait(full_bar + s, ph);
                float4* ah = reinterpret_cast<float4*>(St + s * stage_bytes);
                float4* al = reinterpret_cast<float4*>(St + s * stage_bytes + raw_bytes);
                for (int idx = wtid; idx < (int)(raw_bytes / 16); idx += 32 * TW_WORKER_WARPS) {
                    const float4 v = ah[idx];
                    const float4 h = make_float4(tc::tf32_trunc(v.x), tc::tf32_trunc(v.y), tc::tf32_trunc(v.z), tc::tf32_trunc(v.w));
                    ah[idx] = h;
                    al[idx] = tc::tf32_lo4(v, h);
                }
                tc::fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(split_bar + s);
                if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
            }
        }
        // ---- epilogue: partial product -> workspace (zeros when this CTA had no work) ----
        if (my_items > 0) {
            tc::mbar_wait(done_bar, 0);
            tc::tc_fence_after_sync();
        }
        const int quarter = warp & 3;
        const int cpart = wk >> 2;                                   // 4 column parts
        const int ncol_part = ((p.Qn + 3) / 4 + 3) & ~3;
        const int c_begin = cpart * ncol_part;
        const int c_end = min(p.Qn, c_begin + ncol_part);
        float* wsc = p.ws + (int64_t)blockIdx.x * p.mblocks * 128 * p.Qn;
        const int nacc = (int)(my_items < p.R ? my_items : p.R);
        for (int mb = 0; mb < p.mblocks; ++mb) {
            const int row = mb * 128 + quarter * 32 + lane;
            for (int c0 = c_begin; c0 < c_end; c0 += 16) {
                float r[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) r[j] = 0.f;
                for (int ra = 0; ra < nacc; ++ra) {                   // fp32 sum of the rotating accumulators
                    uint32_t t[16];
                    tc::tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(quarter * 32) << 16) +
                                               (uint32_t)((ra * p.mblocks + mb) * p.Qn + c0), t);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) r[j] += __uint_as_float(t[j]);
                }
                if (row < p.Pc) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (c0 + j < c_end) wsc[(int64_t)row * p.Qn + c0 + j] = r[j];
                }
            }
        }
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, p.tmem_cols);
}

// out[(transpose ? q*Pc + pch : pch*Qc + q)] = sum_c ws[c][pch][q]   (ws rows padded to mblocks*128, Qn columns);
// column Qc (when Qn > Qc) is the ones-row product = the bias gradient of the A side.
// one warp per output element: lanes stride over the partials, fixed-order shuffle reduction (deterministic)
__global__ void __launch_bounds__(256)
tc_wgrad_reduce_kernel(const float* __restrict__ ws, float* __restrict__ out, float* __restrict__ gbias, int Pc, int Qc,
                       int Qn, int prow_pad, int nparts, int transpose) {
    const int64_t e = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const int cols = Qn > Qc ? Qc + 1 : Qc;
    if (e >= (int64_t)Pc * cols) return;
    const int pch = (int)(e / cols), q = (int)(e % cols);
    const int64_t part = (int64_t)prow_pad * Qn;
    const float* src = ws + (int64_t)pch * Qn + q;
    float s = 0.f;
    for (int c = lane; c < nparts; c += 32) s += __ldg(src + c * part);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) {
        if (q == Qc) { if (gbias) gbias[pch] = s; }
        else out[transpose ? (int64_t)q * Pc + pch : (int64_t)pch * Qc + q] = s;
    }
}

// per-channel sums: out[c] = sum_{b,p} g[b,c,p]   (one warp per (b,c) row, then a fixed-order sum over b)
__global__ void __launch_bounds__(256)
channel_rowsum_kernel(const float* __restrict__ g, float* __restrict__ part, int64_t rows, int64_t HW) {
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float4* src = reinterpret_cast<const float4*>(g + row * HW);
    float s = 0.f;
    for (int64_t i = lane; i < HW / 4; i += 32) {
        const float4 v = __ldg(src + i);
        s += (v.x + v.y) + (v.z + v.w);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) part[row] = s;
}
__global__ void __launch_bounds__(256)
channel_sum_final_kernel(const float* __restrict__ part, float* __restrict__ out, int B, int C) {
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= C) return;
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += part[(int64_t)b * C + c];
    out[c] = s;
}

static int tw_num_sms() { return sb200_num_sms(); }

// shape eligibility + geometry, shared by the workspace query and the launcher
static bool tw_geometry(int Cout, int Cin, int64_t HW, int* Pc, int* Qc, int* transpose) {
    if (HW % 4 != 0) return false;
    int P = Cout, Q = Cin, tr = 0;
    if (Cin > Cout) { P = Cin; Q = Cout; tr = 1; }
    if (Q % 16 != 0 || Q < 16 || Q > 256) return false;
    if (P % 8 != 0 || P > 256) return false;
    if (P % 128 != 0 && P > 128) return false;
    const int mblocks = (P + 127) / 128;
    if (mblocks * Q > 512) return false;
    *Pc = P; *Qc = Q; *transpose = tr;
    return true;
}

int64_t sb200_tc_wgrad_workspace(int B, int Cout, int Cin, int64_t HW) {
    int Pc, Qc, tr;
    if (sb_tc_mode() == 0 || !tw_geometry(Cout, Cin, HW, &Pc, &Qc, &tr)) return 0;
    const int mblocks = (Pc + 127) / 128;
    return (int64_t)tw_num_sms() * mblocks * 128 * (Qc + 16) + (int64_t)B * Cout;
}

static int tw_launch(const float* g, const float* x, float* gW, float* gbias, int B, int Cout, int Cin, int64_t HW,
                     float* workspace, cudaStream_t st, int* handled, const float* gen_x, const float* gen_w1,
                     const float* gen_b1) {
    *handled = 0;
    const bool gen = gen_x != nullptr;
    const int passes = sb_tc_mode();
    int Pc, Qc, tr;
    if (passes == 0 || !tw_geometry(Cout, Cin, HW, &Pc, &Qc, &tr)) return 0;
    if ((reinterpret_cast<uintptr_t>(g) & 15) || (!gen && (reinterpret_cast<uintptr_t>(x) & 15))) return 0;
    if (gen && (tr != 1 || (reinterpret_cast<uintptr_t>(gen_x) & 15))) return 0;      // the generated tensor is the wide (A) side
    if ((int64_t)B * Pc >= (1LL << 31)) return 0;
    const float* Pt = tr ? x : g;
    const float* Qt = tr ? g : x;

    TcWgParams p;
    p.trace = nullptr;
#ifdef SB200_BRINGUP
    static long long* trace_dev = nullptr;
    const int want_trace = gen ? sb_env_int("SB200_WG_TRACE", 0) : 0;
    if (want_trace) {
        if (!trace_dev) cudaMalloc(&trace_dev, 2 * 16 * 4 * sizeof(long long));
        cudaMemsetAsync(trace_dev, 0, 2 * 16 * 4 * sizeof(long long), st);
        p.trace = trace_dev;
    }
#endif
    p.gen_x = gen_x; p.gen_w1 = gen_w1; p.gen_b1 = gen_b1;
    p.Pc = Pc; p.Qc = Qc; p.mblocks = (Pc + 127) / 128; p.HW = HW;
    // bias gradient = sum over pixels of g: free from the MMA when g is the A side (ones-row appended to the B side)
    const bool bias_mma = passes == 3 && gbias != nullptr && tr == 0 && Qc + 16 <= 256 && p.mblocks * (Qc + 16) <= 512;
    p.Qn = bias_mma ? Qc + 16 : Qc;
    const size_t chunk = (size_t)(Pc + p.Qn) * 128;
    static const int env_ch = sb_env_int("SB200_WG_CH", 0);       // bring-up builds only (compiled out otherwise)
    static const int env_st = sb_env_int("SB200_WG_STAGES", 0);
    static const int env_dbg = sb_env_int("SB200_WG_DEBUG", 0);
    p.debug = env_dbg;
    p.CH = env_ch > 0 ? env_ch : 1;
    const size_t stage = chunk * p.CH * (passes == 3 ? 2 : 1);
    int stages = env_st > 0 ? env_st : 8;
    while (stages > 1 && 1024 + stages * stage + 16384 + 256 > 210 * 1024) --stages;
    if (1024 + stages * stage + 16384 + 256 > 227 * 1024) return 0;
    p.stages = stages;
    const size_t smem = 1024 + stages * stage + 16384 + 256;
    p.items_per_b = (HW + 32 * p.CH - 1) / (32 * p.CH);
    p.nitems = p.items_per_b * B;
    if (p.nitems >= (1LL << 31)) return 0;
    p.idesc = tc::make_idesc_tf32(128, p.Qn, 0, 0);
    p.R = 512 / (p.mblocks * p.Qn);
    if (p.R > 8) p.R = 8;
    if (p.R < 1) p.R = 1;
    uint32_t cols = 32;
    while (cols < (uint32_t)(p.R * p.mblocks * p.Qn)) cols <<= 1;
    p.tmem_cols = cols;
    p.ws = workspace;
    const int sms = tw_num_sms();
    const unsigned grid = (unsigned)(p.nitems < sms ? p.nitems : sms);

    CUtensorMap tmP, tmQ;
    if (int rc = sb200_make_tmap_2d_f32(&tmQ, Qt, (uint64_t)HW, (uint64_t)B * Qc, (uint64_t)HW * 4, 32, (uint32_t)Qc, 1)) return rc;
    if (gen) {
        tmP = tmQ;     // unused by the kernel
        if (passes == 3) {
            SB_CHECK_CUDA(cudaFuncSetAttribute(tc_wgrad_kernel<3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            sb_launch(tc_wgrad_kernel<3, 1>, grid, TW_THREADS, smem, st, tmP, tmQ, p);
        } else {
            SB_CHECK_CUDA(cudaFuncSetAttribute(tc_wgrad_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            sb_launch(tc_wgrad_kernel<1, 1>, grid, TW_THREADS, smem, st, tmP, tmQ, p);
        }
    } else if (int rc = sb200_make_tmap_2d_f32(&tmP, Pt, (uint64_t)HW, (uint64_t)B * Pc, (uint64_t)HW * 4, 32, (uint32_t)Pc, 1)) {
        return rc;
    } else if (passes == 3) {
        SB_CHECK_CUDA(cudaFuncSetAttribute(tc_wgrad_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sb_launch(tc_wgrad_kernel<3>, grid, TW_THREADS, smem, st, tmP, tmQ, p);
    } else {
        SB_CHECK_CUDA(cudaFuncSetAttribute(tc_wgrad_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sb_launch(tc_wgrad_kernel<1>, grid, TW_THREADS, smem, st, tmP, tmQ, p);
    }
    SB_LAUNCH_CHECK();
#ifdef SB200_BRINGUP
    if (want_trace) {
        static int calls = 0;
        if (++calls == want_trace) {
            long long h[2 * 16 * 4];
            cudaStreamSynchronize(st);
            cudaMemcpy(h, p.trace, sizeof(h), cudaMemcpyDeviceToHost);
            const long long t0 = h[0];
            printf("tc_wgrad (generated A) trace, CTA 0, stages=%d CH=%d: workers(start, stage free, generated, Q landed) | mma(wait0, split ok, committed)\n", p.stages, p.CH);
            for (int i = 0; i < 14; ++i) {
                printf("  %2d:", i);
                for (int r = 0; r < 2; ++r) {
                    for (int e = 0; e < (r ? 3 : 4); ++e) printf(" %7lld", h[(r * 16 + i) * 4 + e] ? h[(r * 16 + i) * 4 + e] - t0 : -1);
                    printf("   |");
                }
                printf("\n");
            }
            fflush(stdout);
        }
    }
#endif
    const int64_t E = (int64_t)Pc * (bias_mma ? Qc + 1 : Qc);
    sb_launch(tc_wgrad_reduce_kernel, (unsigned)ceil_div64(E, 8), 256, 0, st, workspace, gW, bias_mma ? gbias : nullptr, Pc, Qc, p.Qn,
                                                                       p.mblocks * 128, (int)grid, tr);
    SB_LAUNCH_CHECK();
    if (gbias && !bias_mma) {
        float* part = workspace + (int64_t)sms * p.mblocks * 128 * (Qc + 16);
        const int64_t rows = (int64_t)B * Cout;
        sb_launch(channel_rowsum_kernel, (unsigned)ceil_div64(rows, 8), 256, 0, st, g, part, rows, HW);
        SB_LAUNCH_CHECK();
        sb_launch(channel_sum_final_kernel, (unsigned)((Cout + 255) / 256), 256, 0, st, part, gbias, B, Cout);
        SB_LAUNCH_CHECK();
    }
    *handled = 1;
    return 0;
}

int sb200_tc_pointwise_wgrad(const float* g, const float* x, float* gW, float* gbias, int B, int Cout, int Cin,
                             int64_t HW, float* workspace, cudaStream_t st, int* handled) {
    return tw_launch(g, x, gW, gbias, B, Cout, Cin, HW, workspace, st, handled, nullptr, nullptr, nullptr);
}

// Weight gradient of the second lifting layer for a 1-input-channel lifting MLP:
//      gW2[c, n] = sum_{b,p} g[b,c,p] gelu(w1[n] x[b,p] + b1[n]),   gb2[c] = sum_{b,p} g[b,c,p]
// the hidden activations are regenerated on chip (ASRC = 1) instead of being re-read from HBM.
// workspace: sb200_pointwise_wgrad_workspace(B, C, 256, HW) floats.
extern "C" int sb200_lift_wgrad(const float* g, const float* x, const float* w1, const float* b1, float* gW2, float* gb2,
                                float* workspace, int B, int C, int N, int64_t HW, void* stream, int tc_mode) {
    SbModeScope _mode(tc_mode);
    SB_REQUIRE(g && x && w1 && b1 && gW2 && workspace, "lift_wgrad: NULL argument");
    SB_REQUIRE(sb_tc_mode() != 0, "lift_wgrad: runs on the tcgen05 path (tc mode 1 or 3)");
    SB_REQUIRE(N > C, "lift_wgrad: the hidden width (%d) must exceed the output channels (%d)", N, C);
    if (B <= 0) return 0;
    int handled = 0;
    if (int rc = tw_launch(g, nullptr, gW2, gb2, B, C, N, HW, workspace, (cudaStream_t)stream, &handled, x, w1, b1)) return rc;
    SB_REQUIRE(handled, "lift_wgrad: shape not covered by the tcgen05 kernel (C=%d, N=%d, HW=%lld)", C, N, (long long)HW);
    return 0;
}
