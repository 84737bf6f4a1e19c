// Channels-first (FNO / TFNO SpectralConv) transform stages, fp32 CUDA-core versions.
//
//   rowdft_fwd : x[rows,W]        -> T[rows,Mx]      truncated real DFT along W
//   coldft_fwd : T[img,H,Mx]      -> Xh[img,My,Mx]   truncated complex DFT along H
//   coldft_inv : Yh[img,My,Mx]    -> Phi[img,H,Mx]   zero-padded inverse complex DFT along H
//   modes_gemm : per-mode complex channel contraction (and its two adjoints)
//
// Math spec: SURVEY.md section 8c (restating neuralop SpectralConv.forward, fftshift era).
// These are exact-fp32 (FFMA) kernels: they are the 1e-5 parity path and serve every
// shape; the tcgen05 versions in tc_*.cu take over the HBM-heavy row stages on aligned shapes.
#include "common.cuh"

// ======================================================================================
// rowdft_fwd
// ======================================================================================
constexpr int RD_ROWS = 64;
constexpr int RD_XC = 64;

template <int KPT>
__global__ void __launch_bounds__(256)
rowdft_fwd_kernel(const float* __restrict__ x, const float2* __restrict__ tab, float2* __restrict__ T,
                  int64_t R, int W, int Mx, int vec_ok) {
    __shared__ float xs[RD_ROWS][RD_XC + 1];
    __shared__ float2 ts[RD_XC][4 * KPT];
    const int tid = threadIdx.x;
    const int r = tid & 63, g = tid >> 6;
    const int64_t row0 = (int64_t)blockIdx.x * RD_ROWS;

    for (int cb = 0; cb < Mx; cb += 4 * KPT) {
        float2 acc[KPT];
#pragma unroll
        for (int j = 0; j < KPT; ++j) acc[j] = make_float2(0.f, 0.f);

        for (int x0 = 0; x0 < W; x0 += RD_XC) {
            if (vec_ok) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int rr = (tid >> 4) + 16 * i;
                    const int cc = (tid & 15) * 4;
                    const int64_t row = row0 + rr;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (row < R && x0 + cc < W)   // W % 4 == 0 => the whole float4 is in range
                        v = __ldg(reinterpret_cast<const float4*>(x + row * W + x0 + cc));
                    xs[rr][cc + 0] = v.x; xs[rr][cc + 1] = v.y; xs[rr][cc + 2] = v.z; xs[rr][cc + 3] = v.w;
                }
            } else {
                for (int idx = tid; idx < RD_ROWS * RD_XC; idx += 256) {
                    const int rr = idx >> 6, cc = idx & 63;
                    const int64_t row = row0 + rr;
                    float v = 0.f;
                    if (row < R && x0 + cc < W) v = __ldg(x + row * W + x0 + cc);
                    xs[rr][cc] = v;
                }
            }
            for (int idx = tid; idx < RD_XC * 4 * KPT; idx += 256) {
                const int xx = idx / (4 * KPT), j = idx % (4 * KPT);
                float2 v = make_float2(0.f, 0.f);
                if (x0 + xx < W && cb + j < Mx) v = __ldg(tab + (int64_t)(x0 + xx) * Mx + cb + j);
                ts[xx][j] = v;
            }
            __syncthreads();
#pragma unroll 8
            for (int xx = 0; xx < RD_XC; ++xx) {
                const float a = xs[r][xx];
#pragma unroll
                for (int j = 0; j < KPT; ++j) {
                    const float2 t = ts[xx][g * KPT + j];
                    acc[j].x = fmaf(a, t.x, acc[j].x);
                    acc[j].y = fmaf(a, t.y, acc[j].y);
                }
            }
            __syncthreads();
        }
        const int64_t row = row0 + r;
        if (row < R) {
#pragma unroll
            for (int j = 0; j < KPT; ++j) {
                const int kx = cb + g * KPT + j;
                if (kx < Mx) T[row * Mx + kx] = acc[j];
            }
        }
    }
}

int sb200_tc_rowdft_fwd(sb200_plan_t plan, int pass, const float* x, float* T, int64_t rows, cudaStream_t st,
                        int* handled);   // tc_rowdft.cu

extern "C" int sb200_rowdft_fwd(sb200_plan_t p, int pass, const float* x, float* T, int64_t rows, void* stream) {
    SB_REQUIRE(p && x && T, "rowdft_fwd: NULL argument");
    SB_REQUIRE(pass == 0 || pass == 1, "rowdft_fwd: pass must be 0 or 1");
    if (rows <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    int handled = 0;
    if (int rc = sb200_tc_rowdft_fwd(p, pass, x, T, rows, st, &handled)) return rc;
    if (handled) return 0;
    const int W = p->W, Mx = p->Mx;
    const int vec_ok = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    int kpt = (Mx + 3) / 4;
    if (kpt > 8) kpt = 8;
    dim3 grid((unsigned)ceil_div64(rows, RD_ROWS)), block(256);
    const float2* tab = p->rowF[pass];
    float2* To = reinterpret_cast<float2*>(T);
#define RD_CASE(K) case K: rowdft_fwd_kernel<K><<<grid, block, 0, st>>>(x, tab, To, rows, W, Mx, vec_ok); break;
    switch (kpt) {
        RD_CASE(1) RD_CASE(2) RD_CASE(3) RD_CASE(4) RD_CASE(5) RD_CASE(6) RD_CASE(7) RD_CASE(8)
        default: SB_REQUIRE(false, "rowdft_fwd: internal kpt=%d", kpt);
    }
#undef RD_CASE
    SB_LAUNCH_CHECK();
    return 0;
}

// ======================================================================================
// coldft_fwd / coldft_inv
// ======================================================================================
constexpr int CF_HC = 16;

__global__ void __launch_bounds__(1024)
coldft_fwd_kernel(const float2* __restrict__ T, const float2* __restrict__ CF, float2* __restrict__ Xh,
                  int64_t nimg, int H, int My, int Mx, int IPB, int KGB) {
    extern __shared__ float2 sm[];
    float2* Ts = sm;                                  // [IPB][HC][Mx]
    float2* Cs = sm + (size_t)IPB * CF_HC * Mx;       // [KGB*4][HC+1]
    const int tid = threadIdx.x;
    const int64_t img0 = (int64_t)blockIdx.x * IPB;
    const int kyb = blockIdx.y * KGB * 4;
    const int tpi = KGB * Mx;
    const int img_l = tid / tpi;
    const int rem = tid % tpi;
    const int kgl = rem / Mx, kx = rem % Mx;
    const bool active = (img_l < IPB) && (img0 + img_l < nimg);
    float2 acc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = make_float2(0.f, 0.f);

    for (int y0 = 0; y0 < H; y0 += CF_HC) {
        const int per_img = CF_HC * Mx;
        for (int idx = tid; idx < IPB * per_img; idx += blockDim.x) {
            const int il = idx / per_img, rest = idx % per_img;
            const int yy = rest / Mx;
            float2 v = make_float2(0.f, 0.f);
            if (img0 + il < nimg && y0 + yy < H) v = __ldg(T + ((img0 + il) * H + y0) * Mx + rest);
            Ts[idx] = v;
        }
        for (int idx = tid; idx < KGB * 4 * CF_HC; idx += blockDim.x) {
            const int k = idx / CF_HC, yy = idx % CF_HC;
            const int ky = kyb + k;
            float2 v = make_float2(0.f, 0.f);
            if (ky < My && y0 + yy < H) v = __ldg(CF + (int64_t)ky * H + y0 + yy);
            Cs[k * (CF_HC + 1) + yy] = v;
        }
        __syncthreads();
        if (active) {
#pragma unroll
            for (int yy = 0; yy < CF_HC; ++yy) {
                const float2 t = Ts[(img_l * CF_HC + yy) * Mx + kx];
#pragma unroll
                for (int j = 0; j < 4; ++j) cmac(acc[j], Cs[(kgl * 4 + j) * (CF_HC + 1) + yy], t);
            }
        }
        __syncthreads();
    }
    if (active) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int ky = kyb + kgl * 4 + j;
            if (ky < My) Xh[((img0 + img_l) * My + ky) * Mx + kx] = acc[j];
        }
    }
}

__global__ void __launch_bounds__(1024)
coldft_inv_kernel(const float2* __restrict__ Yh, const float2* __restrict__ CI, float2* __restrict__ Phi,
                  int64_t nimg, int H, int My, int Mx, int IPB, int YGB) {
    extern __shared__ float2 sm[];
    float2* Ys = sm;                                  // [IPB][My][Mx]
    float2* Cs = sm + (size_t)IPB * My * Mx;          // [YGB*4][My+1]
    const int tid = threadIdx.x;
    const int64_t img0 = (int64_t)blockIdx.x * IPB;
    const int yb = blockIdx.y * YGB * 4;
    const int tpi = YGB * Mx;
    const int img_l = tid / tpi;
    const int rem = tid % tpi;
    const int ygl = rem / Mx, kx = rem % Mx;
    const bool active = (img_l < IPB) && (img0 + img_l < nimg);

    const int per_img = My * Mx;
    for (int idx = tid; idx < IPB * per_img; idx += blockDim.x) {
        const int il = idx / per_img;
        float2 v = make_float2(0.f, 0.f);
        if (img0 + il < nimg) v = __ldg(Yh + img0 * per_img + idx);
        Ys[idx] = v;
    }
    for (int idx = tid; idx < YGB * 4 * My; idx += blockDim.x) {
        const int k = idx / My, ky = idx % My;
        const int y = yb + k;
        float2 v = make_float2(0.f, 0.f);
        if (y < H) v = __ldg(CI + (int64_t)y * My + ky);
        Cs[k * (My + 1) + ky] = v;
    }
    __syncthreads();
    if (!active) return;
    float2 acc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = make_float2(0.f, 0.f);
#pragma unroll 4
    for (int ky = 0; ky < My; ++ky) {
        const float2 t = Ys[(img_l * My + ky) * Mx + kx];
#pragma unroll
        for (int j = 0; j < 4; ++j) cmac(acc[j], Cs[(ygl * 4 + j) * (My + 1) + ky], t);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int y = yb + ygl * 4 + j;
        if (y < H) Phi[((img0 + img_l) * H + y) * Mx + kx] = acc[j];
    }
}

static int col_launch_cfg(int outer_groups, int Mx, int* GB, int* IPB, int* threads) {
    int gb = 256 / Mx;
    if (gb < 1) gb = 1;
    if (gb > outer_groups) gb = outer_groups;
    int ipb = 256 / (gb * Mx);
    if (ipb < 1) ipb = 1;
    int th = ipb * gb * Mx;
    th = (th + 31) / 32 * 32;
    *GB = gb; *IPB = ipb; *threads = th;
    return th <= 1024 ? 0 : 1;
}

extern "C" int sb200_coldft_fwd(sb200_plan_t p, int pass, const float* T, float* Xh, int64_t nimg, void* stream) {
    SB_REQUIRE(p && T && Xh, "coldft_fwd: NULL argument");
    SB_REQUIRE(pass == 0 || pass == 1, "coldft_fwd: pass must be 0 or 1");
    if (nimg <= 0) return 0;
    const int H = p->H, My = p->My, Mx = p->Mx;
    int KGB, IPB, threads;
    SB_REQUIRE(col_launch_cfg((My + 3) / 4, Mx, &KGB, &IPB, &threads) == 0, "coldft_fwd: Mx=%d too large", Mx);
    if ((int64_t)IPB > nimg) IPB = (int)nimg;
    const size_t smem = ((size_t)IPB * CF_HC * Mx + (size_t)KGB * 4 * (CF_HC + 1)) * sizeof(float2);
    SB_REQUIRE(smem <= 160 * 1024, "coldft_fwd: shared memory %zu too large", smem);
    if (smem > 48 * 1024)
        SB_CHECK_CUDA(cudaFuncSetAttribute(coldft_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)ceil_div64(nimg, IPB), (unsigned)(((My + 3) / 4 + KGB - 1) / KGB));
    coldft_fwd_kernel<<<grid, threads, smem, (cudaStream_t)stream>>>(
        reinterpret_cast<const float2*>(T), p->colF[pass], reinterpret_cast<float2*>(Xh), nimg, H, My, Mx, IPB, KGB);
    SB_LAUNCH_CHECK();
    return 0;
}

extern "C" int sb200_coldft_inv(sb200_plan_t p, int pass, const float* Yh, float* Phi, int64_t nimg, void* stream) {
    SB_REQUIRE(p && Yh && Phi, "coldft_inv: NULL argument");
    SB_REQUIRE(pass == 0 || pass == 1, "coldft_inv: pass must be 0 or 1");
    if (nimg <= 0) return 0;
    const int H = p->H, My = p->My, Mx = p->Mx;
    int YGB, IPB, threads;
    SB_REQUIRE(col_launch_cfg((H + 3) / 4, Mx, &YGB, &IPB, &threads) == 0, "coldft_inv: Mx=%d too large", Mx);
    if ((int64_t)IPB > nimg) IPB = (int)nimg;
    const size_t smem = ((size_t)IPB * My * Mx + (size_t)YGB * 4 * (My + 1)) * sizeof(float2);
    SB_REQUIRE(smem <= 160 * 1024, "coldft_inv: shared memory %zu too large (My*Mx=%d)", smem, My * Mx);
    if (smem > 48 * 1024)
        SB_CHECK_CUDA(cudaFuncSetAttribute(coldft_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)ceil_div64(nimg, IPB), (unsigned)(((H + 3) / 4 + YGB - 1) / YGB));
    coldft_inv_kernel<<<grid, threads, smem, (cudaStream_t)stream>>>(
        reinterpret_cast<const float2*>(Yh), p->colI[pass], reinterpret_cast<float2*>(Phi), nimg, H, My, Mx, IPB, YGB);
    SB_LAUNCH_CHECK();
    return 0;
}

// ======================================================================================
// analysis: x -> Xh in one call (fused single kernel when the grid allows, else row + column stages)
// ======================================================================================
bool sb200_analysis_fused_supported(sb200_plan_t plan);                                              // analysis_fused.cu
int sb200_analysis_fused(sb200_plan_t plan, int pass, const float* x, float* Xh, int64_t nimg, cudaStream_t st, int* handled);

extern "C" int64_t sb200_analysis_scratch(sb200_plan_t p, int64_t nimg) {
    if (!p || nimg <= 0) return 0;
    if (sb200_analysis_fused_supported(p)) return 0;
    return nimg * p->H * p->Mx * 2;
}

extern "C" int sb200_analysis(sb200_plan_t p, int pass, const float* x, float* Xh, int64_t nimg, float* scratch, void* stream) {
    SB_REQUIRE(p && x && Xh, "analysis: NULL argument");
    SB_REQUIRE(pass == 0 || pass == 1, "analysis: pass must be 0 or 1");
    if (nimg <= 0) return 0;
    int handled = 0;
    if (int rc = sb200_analysis_fused(p, pass, x, Xh, nimg, (cudaStream_t)stream, &handled)) return rc;
    if (handled) return 0;
    SB_REQUIRE(scratch != nullptr, "analysis: this grid needs sb200_analysis_scratch() floats of scratch");
    if (int rc = sb200_rowdft_fwd(p, pass, x, scratch, nimg * p->H, stream)) return rc;
    return sb200_coldft_fwd(p, pass, scratch, Xh, nimg, stream);
}

// ======================================================================================
// modes_gemm: out[p,q,k] = sum_r opA(A[r,p,k]) * opB(B[r,q,k])
// ======================================================================================
constexpr int MG_RC = 8;

template <int KT>
__global__ void __launch_bounds__(256)
modes_gemm_kernel(const float2* __restrict__ A, int64_t sAr, int64_t sAp,
                  const float2* __restrict__ B, int64_t sBr, int64_t sBq,
                  float2* __restrict__ out, int64_t sOp, int64_t sOq,
                  int P, int Q, int R, int K, int conjA, int conjB) {
    constexpr int KB = 4 * KT;
    __shared__ __align__(16) float2 As[MG_RC][32][KB];
    __shared__ __align__(16) float2 Bs[MG_RC][32][KB];
    const int tid = threadIdx.x;
    const int tk = tid & 3, tq = (tid >> 2) & 7, tp = tid >> 5;
    const int k0 = blockIdx.x * KB, p0 = blockIdx.y * 32, q0 = blockIdx.z * 32;

    float2 acc[4][4][KT];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int t = 0; t < KT; ++t) acc[i][j][t] = make_float2(0.f, 0.f);

    for (int r0 = 0; r0 < R; r0 += MG_RC) {
        for (int idx = tid; idx < MG_RC * 32 * KB; idx += 256) {
            const int kk = idx % KB;
            const int pp = (idx / KB) % 32;
            const int rr = idx / (KB * 32);
            const int r = r0 + rr, k = k0 + kk;
            float2 va = make_float2(0.f, 0.f), vb = make_float2(0.f, 0.f);
            if (r < R && k < K) {
                if (p0 + pp < P) {
                    va = __ldg(A + (int64_t)r * sAr + (int64_t)(p0 + pp) * sAp + k);
                    if (conjA) va.y = -va.y;
                }
                if (q0 + pp < Q) {
                    vb = __ldg(B + (int64_t)r * sBr + (int64_t)(q0 + pp) * sBq + k);
                    if (conjB) vb.y = -vb.y;
                }
            }
            As[rr][pp][kk] = va;
            Bs[rr][pp][kk] = vb;
        }
        __syncthreads();
#pragma unroll
        for (int rr = 0; rr < MG_RC; ++rr) {
            float2 a[4][KT], b[4][KT];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int t = 0; t < KT; ++t) {
                    a[i][t] = As[rr][tp + 8 * i][tk * KT + t];
                    b[i][t] = Bs[rr][tq + 8 * i][tk * KT + t];
                }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int t = 0; t < KT; ++t) cmac(acc[i][j][t], a[i][t], b[j][t]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int p = p0 + tp + 8 * i;
        if (p >= P) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int q = q0 + tq + 8 * j;
            if (q >= Q) continue;
#pragma unroll
            for (int t = 0; t < KT; ++t) {
                const int k = k0 + tk * KT + t;
                if (k < K) out[(int64_t)p * sOp + (int64_t)q * sOq + k] = acc[i][j][t];
            }
        }
    }
}

extern "C" int sb200_modes_gemm(const float* A, int64_t sAr, int64_t sAp, const float* B, int64_t sBr, int64_t sBq,
                                float* out, int64_t sOp, int64_t sOq, int P, int Q, int R, int K, int conj_flags,
                                void* stream) {
    SB_REQUIRE(A && B && out, "modes_gemm: NULL argument");
    SB_REQUIRE(P >= 0 && Q >= 0 && R >= 0 && K >= 0, "modes_gemm: negative extent");
    if (P == 0 || Q == 0 || K == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int pb = (P + 31) / 32, qb = (Q + 31) / 32;
    SB_REQUIRE(pb <= 65535 && qb <= 65535, "modes_gemm: P/Q too large");
    const int conjA = conj_flags & 1, conjB = (conj_flags >> 1) & 1;
    const float2* A2 = reinterpret_cast<const float2*>(A);
    const float2* B2 = reinterpret_cast<const float2*>(B);
    float2* O2 = reinterpret_cast<float2*>(out);
    // prefer 8 modes per block; fall back to 4 when that would leave SMs idle
    const int64_t blocks8 = (int64_t)((K + 7) / 8) * pb * qb;
    if (blocks8 >= 2 * 148) {
        dim3 grid((K + 7) / 8, pb, qb);
        modes_gemm_kernel<2><<<grid, 256, 0, st>>>(A2, sAr, sAp, B2, sBr, sBq, O2, sOp, sOq, P, Q, R, K, conjA, conjB);
    } else {
        dim3 grid((K + 3) / 4, pb, qb);
        modes_gemm_kernel<1><<<grid, 256, 0, st>>>(A2, sAr, sAp, B2, sBr, sBq, O2, sOp, sOq, P, Q, R, K, conjA, conjB);
    }
    SB_LAUNCH_CHECK();
    return 0;
}
