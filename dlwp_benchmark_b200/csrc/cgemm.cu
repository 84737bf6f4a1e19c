// Strided complex GEMM for the Tucker-factorised spectral weights (TFNO):
//
//      C[m, n] = sum_k opA(A[m, k]) * opB(B[k, n])        (complex64, interleaved re/im)
//
// m and k may each be a two-level composite index (m -> (m / M2, m % M2) with two strides), which is what
// a mode product / factor gradient of a 4-way tensor looks like without ever permuting the tensor:
//
//   mode product      T'[.., n, ..] = sum_k U[n, k] T[.., k, ..]          (rows = all other modes, composite)
//   core/input grad   gT[.., k, ..] = sum_n conj(U[n, k]) gT'[.., n, ..]
//   factor grad       gU[n, k]      = sum_rest gT'[.., n, ..] conj(T[.., k, ..])   (reduction index composite, split-K)
//
// replaces tltorch's Tucker reconstruction / tensorly einsum chain behind neuralop's
// `SpectralConv(factorization="Tucker")` (reached from src/dlwpbench/models/fno/fno.py:136-146) and its
// autograd backward.  fp32 FFMA2 arithmetic (packed pairs over n), register tile TM x TN per thread,
// planar re/im shared-memory tiles so that the inner loop is LDS.128 + broadcast FFMA2.
#include "common.cuh"

namespace {

constexpr int CG_THREADS = 256;
constexpr int CG_BK = 16;

constexpr int CG_MAXG = 8;     // problems of identical geometry per launch (the layers of one FNO)

struct CgParams {
    const float2* A[CG_MAXG];
    const float2* B[CG_MAXG];
    float2* C[CG_MAXG];   // final outputs, or (every entry) the partial workspace [g][s][M][N] when partial != 0
    int M, N, K;
    int M2, K2;           // inner extents of the composite m / k index (1 = plain index)
    long long sAm1, sAm2, sAk1, sAk2;
    long long sBk1, sBk2, sBn;
    long long sCm1, sCm2, sCn;
    int splits, kchunk;   // grid.z = groups*splits; split s reduces k in [s*kchunk, min(K, (s+1)*kchunk))
    int conjA, conjB;
    int partial;
    int a_kfast, b_kfast; // which index the loader walks fastest (the unit-stride one)
};

__device__ __forceinline__ long long cg_off2(int idx, int inner, long long s1, long long s2) {
    if (inner == 1) return (long long)idx * s2;
    return (long long)(idx / inner) * s1 + (long long)(idx % inner) * s2;
}

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(CG_THREADS) cgemm_kernel(const CgParams p) {
    static_assert((BM / TM) * (BN / TN) == CG_THREADS, "thread tiling must cover the block tile");
    constexpr int LA = BM * CG_BK / CG_THREADS, LB = BN * CG_BK / CG_THREADS;
    static_assert(LA >= 1 && LB >= 1, "tile too small");
    constexpr int PA = BM + 4, PB = BN + 4;
    __shared__ __align__(16) float As_re[CG_BK][PA], As_im[CG_BK][PA];
    __shared__ __align__(16) float Bs_re[CG_BK][PB], Bs_im[CG_BK][PB];

    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int g = blockIdx.z / p.splits, s = blockIdx.z % p.splits;
    const float2* __restrict__ Ag = p.A[g];
    const float2* __restrict__ Bg = p.B[g];
    const int kbeg = s * p.kchunk;
    const int kend = min(p.K, kbeg + p.kchunk);

    // ---- loader geometry: element e = tid + 256*j of the [BM x BK] (or [BK x BN]) tile ----
    int a_m[LA], a_k[LA], b_n[LB], b_k[LB];
    long long a_off[LA], b_off[LB];     // offset of the index that does not move with the k loop
#pragma unroll
    for (int j = 0; j < LA; ++j) {
        const int e = tid + CG_THREADS * j;
        if (p.a_kfast) { a_k[j] = e % CG_BK; a_m[j] = e / CG_BK; }
        else           { a_m[j] = e % BM;    a_k[j] = e / BM; }
        const int m = m0 + a_m[j];
        a_off[j] = m < p.M ? cg_off2(m, p.M2, p.sAm1, p.sAm2) : -1;
    }
#pragma unroll
    for (int j = 0; j < LB; ++j) {
        const int e = tid + CG_THREADS * j;
        if (p.b_kfast) { b_k[j] = e % CG_BK; b_n[j] = e / CG_BK; }
        else           { b_n[j] = e % BN;    b_k[j] = e / BN; }
        const int n = n0 + b_n[j];
        b_off[j] = n < p.N ? (long long)n * p.sBn : -1;
    }

    const int tx = tid % (BM / TM), ty = tid / (BM / TM);
    float2 acc_re[TM][(TN + 1) / 2], acc_im[TM][(TN + 1) / 2];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < (TN + 1) / 2; ++j) { acc_re[i][j] = make_float2(0.f, 0.f); acc_im[i][j] = make_float2(0.f, 0.f); }

    float2 ra[LA], rb[LB];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int j = 0; j < LA; ++j) {
            const int k = k0 + a_k[j];
            float2 v = make_float2(0.f, 0.f);
            if (a_off[j] >= 0 && k < kend) v = __ldg(Ag + a_off[j] + cg_off2(k, p.K2, p.sAk1, p.sAk2));
            ra[j] = v;
        }
#pragma unroll
        for (int j = 0; j < LB; ++j) {
            const int k = k0 + b_k[j];
            float2 v = make_float2(0.f, 0.f);
            if (b_off[j] >= 0 && k < kend) v = __ldg(Bg + b_off[j] + cg_off2(k, p.K2, p.sBk1, p.sBk2));
            rb[j] = v;
        }
    };
    const float sa = p.conjA ? -1.f : 1.f, sb = p.conjB ? -1.f : 1.f;

    fetch(kbeg);
    for (int k0 = kbeg; k0 < kend; k0 += CG_BK) {
        __syncthreads();
#pragma unroll
        for (int j = 0; j < LA; ++j) { As_re[a_k[j]][a_m[j]] = ra[j].x; As_im[a_k[j]][a_m[j]] = sa * ra[j].y; }
#pragma unroll
        for (int j = 0; j < LB; ++j) { Bs_re[b_k[j]][b_n[j]] = rb[j].x; Bs_im[b_k[j]][b_n[j]] = sb * rb[j].y; }
        __syncthreads();
        if (k0 + CG_BK < kend) fetch(k0 + CG_BK);
#pragma unroll
        for (int kk = 0; kk < CG_BK; ++kk) {
            float ar[TM], ai[TM], br[TN + 1], bi[TN + 1];
            if constexpr (TM == 4) {
                const float4 v = *reinterpret_cast<const float4*>(&As_re[kk][tx * 4]);
                const float4 w = *reinterpret_cast<const float4*>(&As_im[kk][tx * 4]);
                ar[0] = v.x; ar[1] = v.y; ar[2] = v.z; ar[3] = v.w;
                ai[0] = w.x; ai[1] = w.y; ai[2] = w.z; ai[3] = w.w;
            } else {
#pragma unroll
                for (int i = 0; i < TM; ++i) { ar[i] = As_re[kk][tx * TM + i]; ai[i] = As_im[kk][tx * TM + i]; }
            }
            if constexpr (TN == 4) {
                const float4 v = *reinterpret_cast<const float4*>(&Bs_re[kk][ty * 4]);
                const float4 w = *reinterpret_cast<const float4*>(&Bs_im[kk][ty * 4]);
                br[0] = v.x; br[1] = v.y; br[2] = v.z; br[3] = v.w;
                bi[0] = w.x; bi[1] = w.y; bi[2] = w.z; bi[3] = w.w;
            } else {
#pragma unroll
                for (int j = 0; j < TN; ++j) { br[j] = Bs_re[kk][ty * TN + j]; bi[j] = Bs_im[kk][ty * TN + j]; }
                if (TN & 1) { br[TN] = 0.f; bi[TN] = 0.f; }
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < (TN + 1) / 2; ++j) {
                    const float2 brp = make_float2(br[2 * j], br[2 * j + 1]);
                    const float2 bip = make_float2(bi[2 * j], bi[2 * j + 1]);
                    acc_re[i][j] = ffma2(make_float2(ar[i], ar[i]), brp, acc_re[i][j]);
                    acc_re[i][j] = ffma2(make_float2(-ai[i], -ai[i]), bip, acc_re[i][j]);
                    acc_im[i][j] = ffma2(make_float2(ar[i], ar[i]), bip, acc_im[i][j]);
                    acc_im[i][j] = ffma2(make_float2(ai[i], ai[i]), brp, acc_im[i][j]);
                }
        }
    }

    // ---- epilogue ----
    float2* __restrict__ Cg = p.C[g];
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + tx * TM + i;
        if (m >= p.M) continue;
        const long long rowoff = p.partial ? ((long long)blockIdx.z * p.M + m) * p.N : cg_off2(m, p.M2, p.sCm1, p.sCm2);
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = n0 + ty * TN + j;
            if (n >= p.N) continue;
            const float re = (j & 1) ? acc_re[i][j / 2].y : acc_re[i][j / 2].x;
            const float im = (j & 1) ? acc_im[i][j / 2].y : acc_im[i][j / 2].x;
            Cg[rowoff + (p.partial ? (long long)n : (long long)n * p.sCn)] = make_float2(re, im);
        }
    }
}

// Skinny products (N <= 32, K <= 64, many rows): the small spatial modes of a Tucker tensor.  One thread owns TWO
// rows and all N outputs of each in registers; opB(B) sits in shared memory as (re, im, -im, re) so a complex MAC is
// two FFMA2 with the thread's A element as the broadcast scalar and one warp-uniform LDS.128.
constexpr int CS_THREADS = 128;
constexpr int CS_MAXK = 64;

template <int NP>
__global__ void __launch_bounds__(CS_THREADS) cskinny_kernel(const CgParams p) {
    __shared__ __align__(16) float4 Bs[CS_MAXK * NP];
    const int g = blockIdx.y;
    const float2* __restrict__ Ag = p.A[g];
    const float2* __restrict__ Bg = p.B[g];
    float2* __restrict__ Cg = p.C[g];
    const float sb = p.conjB ? -1.f : 1.f, sa = p.conjA ? -1.f : 1.f;
    for (int idx = threadIdx.x; idx < p.K * NP; idx += CS_THREADS) {
        const int k = idx / NP, n = idx % NP;
        float2 u = make_float2(0.f, 0.f);
        if (n < p.N) u = __ldg(Bg + (long long)k * p.sBk2 + (long long)n * p.sBn);
        u.y *= sb;
        Bs[idx] = make_float4(u.x, u.y, -u.y, u.x);
    }
    __syncthreads();
    const int mbase = blockIdx.x * (2 * CS_THREADS) + threadIdx.x;
    int m[2] = {mbase, mbase + CS_THREADS};
    long long aoff[2];
    bool ok[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        ok[r] = m[r] < p.M;
        aoff[r] = ok[r] ? cg_off2(m[r], p.M2, p.sAm1, p.sAm2) : 0;
    }
    float2 acc[2][NP];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int n = 0; n < NP; ++n) acc[r][n] = make_float2(0.f, 0.f);
    // k in chunks of 8: the loads of a chunk are issued back to back and the NEXT chunk is in flight during the math
    constexpr int KC = 8;
    float2 cur[2][KC], nxt[2][KC];
    auto fetch = [&](float2 (&dst)[2][KC], int k0) {
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int kk = 0; kk < KC; ++kk) {
                const int k = k0 + kk;
                float2 v = make_float2(0.f, 0.f);
                if (ok[r] && k < p.K) v = __ldg(Ag + aoff[r] + (long long)k * p.sAk2);
                v.y *= sa;
                dst[r][kk] = v;
            }
    };
    fetch(cur, 0);
    for (int k0 = 0; k0 < p.K; k0 += KC) {
        if (k0 + KC < p.K) fetch(nxt, k0 + KC);
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) {
            if (k0 + kk < p.K) {
                const float4* bk = Bs + (k0 + kk) * NP;
#pragma unroll
                for (int n = 0; n < NP; ++n) {
                    const float4 u = bk[n];
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        acc[r][n] = ffma2(make_float2(cur[r][kk].x, cur[r][kk].x), make_float2(u.x, u.y), acc[r][n]);
                        acc[r][n] = ffma2(make_float2(cur[r][kk].y, cur[r][kk].y), make_float2(u.z, u.w), acc[r][n]);
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int kk = 0; kk < KC; ++kk) cur[r][kk] = nxt[r][kk];
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        if (!ok[r]) continue;
        float2* dst = Cg + cg_off2(m[r], p.M2, p.sCm1, p.sCm2);
#pragma unroll
        for (int n = 0; n < NP; ++n)
            if (n < p.N) dst[(long long)n * p.sCn] = acc[r][n];
    }
}

// C_g[m, n] = sum_s ws[g][s][m][n]   (deterministic split-K reduction; output through the C strides).
// 32 consecutive outputs x 8 split lanes per block: coalesced partial reads, shared-memory tree at the end.
struct CgOut { float2* C[CG_MAXG]; };

__global__ void __launch_bounds__(256) cgemm_reduce_kernel(const float2* __restrict__ ws, const CgOut out, int M, int N,
                                                           int splits, int M2, long long sCm1, long long sCm2,
                                                           long long sCn) {
    __shared__ float2 part[8][32];
    const int o = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const long long MN = (long long)M * N;
    const long long idx = (long long)blockIdx.x * 32 + o;
    const int g = blockIdx.y;
    float2 a = make_float2(0.f, 0.f);
    if (idx < MN) {
        const float2* w = ws + (long long)g * splits * MN + idx;
        for (int s = sl; s < splits; s += 8) {
            const float2 v = w[(long long)s * MN];
            a.x += v.x; a.y += v.y;
        }
    }
    part[sl][o] = a;
    __syncthreads();
    if (sl == 0 && idx < MN) {
#pragma unroll
        for (int t = 1; t < 8; ++t) { a.x += part[t][o].x; a.y += part[t][o].y; }
        const int m = (int)(idx / N), n = (int)(idx % N);
        out.C[g][cg_off2(m, M2, sCm1, sCm2) + (long long)n * sCn] = a;
    }
}

int g_cg_sms = 0;

struct CgShape { int bm, bn; };

CgShape cg_pick(int M, int N) {
    if (M <= 16 && N <= 16) return {16, 16};
    if (M <= 32 && N <= 32) return {32, 32};
    if (N <= 16) return {256, 16};
    if (N <= 32) return {128, 32};
    return {64, 64};
}

int cg_splits(const sb200_cgemm_desc* d, int ngroups, int* kchunk) {
    const CgShape t = cg_pick(d->M, d->N);
    const long long tiles = (long long)((d->M + t.bm - 1) / t.bm) * ((d->N + t.bn - 1) / t.bn) * ngroups;
    int splits = 1;
    if (g_cg_sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&g_cg_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
            g_cg_sms = 148;
    }
    const long long want = 2LL * g_cg_sms;
    if (tiles < want && d->K > 4 * CG_BK) {
        const long long by_tiles = (want + tiles - 1) / tiles;
        const long long by_k = (d->K + 4 * CG_BK - 1) / (4 * CG_BK);
        splits = (int)(by_tiles < by_k ? by_tiles : by_k);
        if (splits < 1) splits = 1;
    }
    int kc = (d->K + splits - 1) / splits;
    kc = (kc + CG_BK - 1) / CG_BK * CG_BK;
    splits = (d->K + kc - 1) / kc;
    *kchunk = kc;
    return splits;
}

}  // namespace

extern "C" int64_t sb200_cgemm_workspace(const sb200_cgemm_desc* d, int ngroups) {
    if (!d || d->M <= 0 || d->N <= 0 || d->K <= 0 || ngroups <= 0) return 0;
    int kc = 0;
    if (d->N <= 32 && d->K <= CS_MAXK && d->K2 == 1 && d->M >= 2048) return 0;     // skinny kernel: no split
    const int splits = cg_splits(d, ngroups, &kc);
    return splits > 1 ? 2LL * ngroups * splits * d->M * d->N : 0;
}

extern "C" int sb200_cgemm_grouped(const sb200_cgemm_desc* d, int ngroups, const float* const* A, const float* const* B,
                                   float* const* C, float* workspace, void* stream) {
    SB_REQUIRE(d && A && B && C, "cgemm: NULL argument");
    SB_REQUIRE(ngroups >= 1 && ngroups <= CG_MAXG, "cgemm: 1..%d problems per call (got %d)", CG_MAXG, ngroups);
    SB_REQUIRE(d->M > 0 && d->N > 0 && d->K > 0, "cgemm: empty problem (M=%d N=%d K=%d)", d->M, d->N, d->K);
    SB_REQUIRE(d->M2 >= 1 && d->K2 >= 1, "cgemm: composite extents must be >= 1");
    cudaStream_t st = (cudaStream_t)stream;
    CgParams p;
    p.M = d->M; p.N = d->N; p.K = d->K; p.M2 = d->M2; p.K2 = d->K2;
    p.sAm1 = d->sAm1; p.sAm2 = d->sAm2; p.sAk1 = d->sAk1; p.sAk2 = d->sAk2;
    p.sBk1 = d->sBk1; p.sBk2 = d->sBk2; p.sBn = d->sBn;
    p.sCm1 = d->sCm1; p.sCm2 = d->sCm2; p.sCn = d->sCn;
    p.conjA = d->conjA; p.conjB = d->conjB;
    p.a_kfast = (d->sAk2 == 1 && d->sAm2 != 1) ? 1 : 0;
    p.b_kfast = (d->sBk2 == 1 && d->sBn != 1) ? 1 : 0;
    static const bool skinny_off = sb_env_flag("SB200_CGEMM_NOSKINNY");     // experiments: tiled kernel for every shape
    const bool skinny = d->N <= 32 && d->K <= CS_MAXK && d->K2 == 1 && d->M >= 2048 && !skinny_off;
    p.splits = skinny ? 1 : cg_splits(d, ngroups, &p.kchunk);
    if (skinny) p.kchunk = d->K;
    p.partial = p.splits > 1;
    SB_REQUIRE(!p.partial || workspace, "cgemm: this shape needs sb200_cgemm_workspace() floats of workspace");
    CgOut out;
    for (int g = 0; g < CG_MAXG; ++g) {
        const int gg = g < ngroups ? g : 0;
        SB_REQUIRE(A[gg] && B[gg] && C[gg], "cgemm: NULL operand in group %d", gg);
        p.A[g] = reinterpret_cast<const float2*>(A[gg]);
        p.B[g] = reinterpret_cast<const float2*>(B[gg]);
        out.C[g] = reinterpret_cast<float2*>(C[gg]);
        p.C[g] = p.partial ? reinterpret_cast<float2*>(workspace) : out.C[g];
    }
    if (skinny) {
        dim3 sgrid((unsigned)((d->M + 2 * CS_THREADS - 1) / (2 * CS_THREADS)), (unsigned)ngroups);
        if (d->N <= 8) sb_launch(cskinny_kernel<8>, sgrid, CS_THREADS, 0, st, p);
        else if (d->N <= 16) sb_launch(cskinny_kernel<16>, sgrid, CS_THREADS, 0, st, p);
        else sb_launch(cskinny_kernel<32>, sgrid, CS_THREADS, 0, st, p);
        SB_LAUNCH_CHECK();
        return 0;
    }
    const CgShape t = cg_pick(d->M, d->N);
    dim3 grid((unsigned)((d->M + t.bm - 1) / t.bm), (unsigned)((d->N + t.bn - 1) / t.bn), (unsigned)(p.splits * ngroups));
    SB_REQUIRE(grid.y <= 65535u && grid.z <= 65535u, "cgemm: grid too large");
    if (t.bm == 16) sb_launch(cgemm_kernel<16, 16, 1, 1>, grid, CG_THREADS, 0, st, p);
    else if (t.bm == 32) sb_launch(cgemm_kernel<32, 32, 2, 2>, grid, CG_THREADS, 0, st, p);
    else if (t.bm == 256) sb_launch(cgemm_kernel<256, 16, 4, 4>, grid, CG_THREADS, 0, st, p);
    else if (t.bm == 128) sb_launch(cgemm_kernel<128, 32, 4, 4>, grid, CG_THREADS, 0, st, p);
    else sb_launch(cgemm_kernel<64, 64, 4, 4>, grid, CG_THREADS, 0, st, p);
    SB_LAUNCH_CHECK();
    if (p.partial) {
        const long long MN = (long long)d->M * d->N;
        dim3 rgrid((unsigned)((MN + 31) / 32), (unsigned)ngroups);
        sb_launch(cgemm_reduce_kernel, rgrid, 256, 0, st, reinterpret_cast<const float2*>(workspace), out, d->M, d->N, p.splits,
                                                   d->M2, d->sCm1, d->sCm2, d->sCn);
        SB_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int sb200_cgemm(const sb200_cgemm_desc* d, const float* A, const float* B, float* C, float* workspace,
                           void* stream) {
    return sb200_cgemm_grouped(d, 1, &A, &B, &C, workspace, stream);
}
