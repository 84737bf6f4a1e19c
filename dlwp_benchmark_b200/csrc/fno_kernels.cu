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

extern "C" int sb200_rowdft_fwd(sb200_plan_t p, int pass, const float* x, float* T, int64_t rows, void* stream, int tc_mode) {
    SbModeScope _mode(tc_mode);
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
#define RD_CASE(K) case K: sb_launch(rowdft_fwd_kernel<K>, grid, block, 0, st, x, tab, To, rows, W, Mx, vec_ok); break;
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

// v2 of the zero-padded inverse column transform: one thread owns one (image, kx) pair and a segment of output rows;
// its My input modes live in registers (plus the (-im, re) rotated copy, so that a complex MAC is two FFMA2 with a
// warp-uniform twiddle broadcast from shared memory).  Reads and writes are kx-contiguous per image.
// fold != 0 (H even): rows y and y + H/2 share their twiddles up to (-1)^ky, so one pass over the modes gives both:
// out[y] = E + O, out[y + H/2] = fold * (E - O) with E / O the sums over the even- / odd-indexed modes (fold = +1 when the
// first retained frequency is even, -1 when odd).  The block then owns yseg rows of the first half and their partners.
template <int MYP>
__global__ void __launch_bounds__(128)
coldft_inv2_kernel(const float2* __restrict__ Yh, const float2* __restrict__ CI, float2* __restrict__ Phi,
                   int64_t nitems, int H, int My, int Mx, int yseg, float fold) {
    extern __shared__ __align__(16) float2 tw[];          // [yseg][MYP] twiddles of this block's rows
    const int y0 = blockIdx.y * yseg;
    const int Hr = fold != 0.f ? H / 2 : H;               // rows that are computed
    const int ny = min(yseg, Hr - y0);
    for (int idx = threadIdx.x; idx < yseg * MYP; idx += blockDim.x) {
        const int yy = idx / MYP, ky = idx % MYP;
        tw[idx] = (yy < ny && ky < My) ? __ldg(CI + (int64_t)(y0 + yy) * My + ky) : make_float2(0.f, 0.f);
    }
    const int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = item < nitems;
    const int64_t img = active ? item / Mx : 0;
    const int kx = active ? (int)(item % Mx) : 0;
    float2 v[MYP], vr[MYP];
    const float2* src = Yh + img * My * Mx + kx;
#pragma unroll
    for (int ky = 0; ky < MYP; ++ky) {
        v[ky] = (active && ky < My) ? __ldg(src + (int64_t)ky * Mx) : make_float2(0.f, 0.f);
        vr[ky] = make_float2(-v[ky].y, v[ky].x);
    }
    __syncthreads();
    if (!active) return;
    float2* dst = Phi + (img * H + y0) * Mx + kx;
    for (int yy = 0; yy < ny; ++yy) {
        const float4* t4 = reinterpret_cast<const float4*>(tw + yy * MYP);
        float2 a0 = make_float2(0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
#pragma unroll
        for (int h = 0; h < MYP / 2; ++h) {
            const float4 t = t4[h];                        // twiddles of ky = 2h, 2h+1
            a0 = ffma2(make_float2(t.x, t.x), v[2 * h], a0);
            a1 = ffma2(make_float2(t.y, t.y), vr[2 * h], a1);
            a2 = ffma2(make_float2(t.z, t.z), v[2 * h + 1], a2);
            a3 = ffma2(make_float2(t.w, t.w), vr[2 * h + 1], a3);
        }
        const float2 e = make_float2(a0.x + a1.x, a0.y + a1.y), o = make_float2(a2.x + a3.x, a2.y + a3.y);
        dst[(int64_t)yy * Mx] = make_float2(e.x + o.x, e.y + o.y);
        if (fold != 0.f) dst[(int64_t)(yy + Hr) * Mx] = make_float2(fold * (e.x - o.x), fold * (e.y - o.y));
    }
}

template <int MYP>
static int coldft_inv2_launch(const float2* Yh, const float2* CI, float2* Phi, int64_t nimg, int H, int My, int Mx,
                              int ky0, cudaStream_t st) {
    const int64_t nitems = nimg * Mx;
    const float fold = (H % 2 == 0) ? ((((ky0 % 2) + 2) % 2) ? -1.f : 1.f) : 0.f;
    const int Hr = fold != 0.f ? H / 2 : H;
    // enough row segments to give every SM sub-partition several warps
    int segs = 1;
    while (segs < 8 && (nitems / 32) * segs < 4 * 4 * 148 && Hr / (segs * 2) >= 4) segs *= 2;
    const int yseg = (Hr + segs - 1) / segs;
    dim3 grid((unsigned)ceil_div64(nitems, 128), (unsigned)((Hr + yseg - 1) / yseg));
    const size_t smem = (size_t)yseg * MYP * sizeof(float2);
    if (smem > 48 * 1024)
        SB_CHECK_CUDA(cudaFuncSetAttribute(coldft_inv2_kernel<MYP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sb_launch(coldft_inv2_kernel<MYP>, grid, 128, smem, st, Yh, CI, Phi, nitems, H, My, Mx, yseg, fold);
    SB_LAUNCH_CHECK();
    return 0;
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
    sb_launch(coldft_fwd_kernel, grid, threads, smem, (cudaStream_t)stream, 
        reinterpret_cast<const float2*>(T), p->colF[pass], reinterpret_cast<float2*>(Xh), nimg, H, My, Mx, IPB, KGB);
    SB_LAUNCH_CHECK();
    return 0;
}

extern "C" int sb200_coldft_inv(sb200_plan_t p, int pass, const float* Yh, float* Phi, int64_t nimg, void* stream) {
    SB_REQUIRE(p && Yh && Phi, "coldft_inv: NULL argument");
    SB_REQUIRE(pass == 0 || pass == 1, "coldft_inv: pass must be 0 or 1");
    if (nimg <= 0) return 0;
    const int H = p->H, My = p->My, Mx = p->Mx;
    static const bool inv_v1 = sb_env_flag("SB200_COLDFT_INV_V1");          // experiments: first-generation kernel
    if (My <= 32 && !inv_v1) {
        const float2* Y2 = reinterpret_cast<const float2*>(Yh);
        float2* P2 = reinterpret_cast<float2*>(Phi);
        cudaStream_t st = (cudaStream_t)stream;
        if (My <= 8) return coldft_inv2_launch<8>(Y2, p->colI[pass], P2, nimg, H, My, Mx, p->ky0, st);
        if (My <= 16) return coldft_inv2_launch<16>(Y2, p->colI[pass], P2, nimg, H, My, Mx, p->ky0, st);
        if (My <= 24) return coldft_inv2_launch<24>(Y2, p->colI[pass], P2, nimg, H, My, Mx, p->ky0, st);
        return coldft_inv2_launch<32>(Y2, p->colI[pass], P2, nimg, H, My, Mx, p->ky0, st);
    }
    int YGB, IPB, threads;
    SB_REQUIRE(col_launch_cfg((H + 3) / 4, Mx, &YGB, &IPB, &threads) == 0, "coldft_inv: Mx=%d too large", Mx);
    if ((int64_t)IPB > nimg) IPB = (int)nimg;
    const size_t smem = ((size_t)IPB * My * Mx + (size_t)YGB * 4 * (My + 1)) * sizeof(float2);
    SB_REQUIRE(smem <= 160 * 1024, "coldft_inv: shared memory %zu too large (My*Mx=%d)", smem, My * Mx);
    if (smem > 48 * 1024)
        SB_CHECK_CUDA(cudaFuncSetAttribute(coldft_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)ceil_div64(nimg, IPB), (unsigned)(((H + 3) / 4 + YGB - 1) / YGB));
    sb_launch(coldft_inv_kernel, grid, threads, smem, (cudaStream_t)stream, 
        reinterpret_cast<const float2*>(Yh), p->colI[pass], reinterpret_cast<float2*>(Phi), nimg, H, My, Mx, IPB, YGB);
    SB_LAUNCH_CHECK();
    return 0;
}

// ======================================================================================
// analysis: x -> Xh in one call (fused single kernel when the grid allows, else row + column stages)
// ======================================================================================
bool sb200_analysis_fused_supported(sb200_plan_t plan);                                              // analysis_fused.cu
int sb200_analysis_fused(sb200_plan_t plan, int pass, const float* x, float* Xh, int64_t nimg, cudaStream_t st, int* handled);

int sb200_tc_analysis(sb200_plan_t plan, int pass, const float* x, float* Xh, int64_t nimg, float* scratch, cudaStream_t st,
                      int* handled);   // tc_rowdft.cu

// which path serves grids that both the fused FFMA kernel and the tensor-core two-stage path cover
// (SB200_ANALYSIS_PREFER_TC=1; bring-up builds only)
static bool analysis_prefers_tc() {
    static const bool v = sb_env_flag("SB200_ANALYSIS_PREFER_TC");
    return v;
}

extern "C" int64_t sb200_analysis_scratch(sb200_plan_t p, int64_t nimg) {
    if (!p || nimg <= 0) return 0;
    if (sb200_analysis_fused_supported(p) && !analysis_prefers_tc()) return 0;
    return nimg * p->H * p->Mx * 2;
}

extern "C" int sb200_analysis(sb200_plan_t p, int pass, const float* x, float* Xh, int64_t nimg, float* scratch, void* stream, int tc_mode) {
    SbModeScope _mode(tc_mode);
    SB_REQUIRE(p && x && Xh, "analysis: NULL argument");
    SB_REQUIRE(pass == 0 || pass == 1, "analysis: pass must be 0 or 1");
    if (nimg <= 0) return 0;
    int handled = 0;
    if (analysis_prefers_tc() && scratch != nullptr) {
        if (int rc = sb200_tc_analysis(p, pass, x, Xh, nimg, scratch, (cudaStream_t)stream, &handled)) return rc;
        if (handled) return 0;
    }
    if (int rc = sb200_analysis_fused(p, pass, x, Xh, nimg, (cudaStream_t)stream, &handled)) return rc;
    if (handled) return 0;
    SB_REQUIRE(scratch != nullptr, "analysis: this grid needs sb200_analysis_scratch() floats of scratch");
    if (int rc = sb200_tc_analysis(p, pass, x, Xh, nimg, scratch, (cudaStream_t)stream, &handled)) return rc;   // tc_rowdft.cu
    if (handled) return 0;
    if (int rc = sb200_rowdft_fwd(p, pass, x, scratch, nimg * p->H, stream, sb_tc_mode())) return rc;
    return sb200_coldft_fwd(p, pass, scratch, Xh, nimg, stream);
}

// ======================================================================================
// modes_gemm: out[p,q,k] = sum_r opA(A[r,p,k]) * opB(B[r,q,k])
// ======================================================================================
constexpr int MG_RC = 8;

// v2: CTA tile = 32 p x 32 q x 4 modes, 8 warps; a warp owns ONE mode (so the shared-memory operand reads are
// warp-broadcast LDS.128) and a thread owns p = tp + 8i, q = tq + 8j (i, j < 4).  Operands are staged planar
// (re / im) with the p (q) axis permuted, pos(p) = 4*(p & 7) + (p >> 3), so that a thread's four p are one
// LDS.128; the complex MAC is four broadcast FFMA2 over q pairs; the next r-slab is prefetched into registers
// during the math; the output goes through shared memory so that global stores are k-contiguous full sectors.
// Every shared-memory access pattern below is bank-conflict free (see the lane maps).
constexpr int MG2_PS = 40;                    // floats per (r, k) row: 32 positions + pad (bank = 8*k + pos)
constexpr int MG2_OS = 40;                    // float2 per (k, q) output row
constexpr int MG2_OPLANE = 32 * MG2_OS + 4;   // float2 per k plane of the output stage

__global__ void __launch_bounds__(256)
modes_gemm2_kernel(const float2* __restrict__ A, int64_t sAr, int64_t sAp,
                   const float2* __restrict__ B, int64_t sBr, int64_t sBq,
                   float2* __restrict__ out, int64_t sOp, int64_t sOq,
                   int P, int Q, int R, int K, int conjA, int conjB) {
    constexpr int STAGE_FLOATS = 4 * MG_RC * 4 * MG2_PS, OUT_FLOATS = 2 * 4 * MG2_OPLANE;
    __shared__ __align__(16) float smem[STAGE_FLOATS > OUT_FLOATS ? STAGE_FLOATS : OUT_FLOATS];
    float (*As_re)[4][MG2_PS] = reinterpret_cast<float (*)[4][MG2_PS]>(smem);
    float (*As_im)[4][MG2_PS] = reinterpret_cast<float (*)[4][MG2_PS]>(smem + MG_RC * 4 * MG2_PS);
    float (*Bs_re)[4][MG2_PS] = reinterpret_cast<float (*)[4][MG2_PS]>(smem + 2 * MG_RC * 4 * MG2_PS);
    float (*Bs_im)[4][MG2_PS] = reinterpret_cast<float (*)[4][MG2_PS]>(smem + 3 * MG_RC * 4 * MG2_PS);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int kk = warp & 3;                            // mode of this warp
    const int tp = lane & 7;                            // p = tp + 8i
    const int tq = (lane >> 3) + 4 * (warp >> 2);       // q = tq + 8j
    const int k0 = blockIdx.x * 4, p0 = blockIdx.y * 32, q0 = blockIdx.z * 32;

    // loader lane map: k = lane & 3 (one 32-byte sector per (r, p)), p = i + 8c with i = ((lane >> 2) & 1) + 2*(warp & 3),
    // c = lane >> 3; r = (warp >> 2) + 2j.  Staging position 4i + c -> bank 8k + 4i + c: 32 distinct banks per warp.
    const int l_k = lane & 3, l_i = ((lane >> 2) & 1) + 2 * (warp & 3), l_c = lane >> 3, l_r = warp >> 2;
    const int l_p = l_i + 8 * l_c, l_pos = 4 * l_i + l_c;
    const bool ka = k0 + l_k < K;
    const bool pa = ka && (p0 + l_p < P), qa = ka && (q0 + l_p < Q);
    const float2* Ap = A + (int64_t)(p0 + l_p) * sAp + k0 + l_k;
    const float2* Bp = B + (int64_t)(q0 + l_p) * sBq + k0 + l_k;
    const float sa = conjA ? -1.f : 1.f, sb = conjB ? -1.f : 1.f;

    float2 ra[4], rb[4];
    auto fetch = [&](int r0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = r0 + l_r + 2 * j;
            ra[j] = (pa && r < R) ? __ldg(Ap + (int64_t)r * sAr) : make_float2(0.f, 0.f);
            rb[j] = (qa && r < R) ? __ldg(Bp + (int64_t)r * sBr) : make_float2(0.f, 0.f);
        }
    };
    float2 acc_re[4][2], acc_im[4][2];                  // [i][q pair]: q = tq + 8*(2*pair), tq + 8*(2*pair + 1)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) { acc_re[i][j] = make_float2(0.f, 0.f); acc_im[i][j] = make_float2(0.f, 0.f); }

    fetch(0);
    for (int r0 = 0; r0 < R; r0 += MG_RC) {
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int rr = l_r + 2 * j;
            As_re[rr][l_k][l_pos] = ra[j].x; As_im[rr][l_k][l_pos] = sa * ra[j].y;
            Bs_re[rr][l_k][l_pos] = rb[j].x; Bs_im[rr][l_k][l_pos] = sb * rb[j].y;
        }
        __syncthreads();
        if (r0 + MG_RC < R) fetch(r0 + MG_RC);
#pragma unroll
        for (int rr = 0; rr < MG_RC; ++rr) {
            const float4 ar = *reinterpret_cast<const float4*>(&As_re[rr][kk][4 * tp]);
            const float4 ai = *reinterpret_cast<const float4*>(&As_im[rr][kk][4 * tp]);
            const float4 br = *reinterpret_cast<const float4*>(&Bs_re[rr][kk][4 * tq]);
            const float4 bi = *reinterpret_cast<const float4*>(&Bs_im[rr][kk][4 * tq]);
            const float a_r[4] = {ar.x, ar.y, ar.z, ar.w}, a_i[4] = {ai.x, ai.y, ai.z, ai.w};
            const float2 brp[2] = {make_float2(br.x, br.y), make_float2(br.z, br.w)};
            const float2 bip[2] = {make_float2(bi.x, bi.y), make_float2(bi.z, bi.w)};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    acc_re[i][j] = ffma2(make_float2(a_r[i], a_r[i]), brp[j], acc_re[i][j]);
                    acc_re[i][j] = ffma2(make_float2(-a_i[i], -a_i[i]), bip[j], acc_re[i][j]);
                    acc_im[i][j] = ffma2(make_float2(a_r[i], a_r[i]), bip[j], acc_im[i][j]);
                    acc_im[i][j] = ffma2(make_float2(a_i[i], a_i[i]), brp[j], acc_im[i][j]);
                }
        }
    }
    // ---- epilogue: Os[k][q][p] through shared memory, then k-contiguous (full-sector) global stores ----
    __syncthreads();
    float2* Os = reinterpret_cast<float2*>(smem);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float re = (j & 1) ? acc_re[i][j >> 1].y : acc_re[i][j >> 1].x;
            const float im = (j & 1) ? acc_im[i][j >> 1].y : acc_im[i][j >> 1].x;
            Os[kk * MG2_OPLANE + (tq + 8 * j) * MG2_OS + tp + 8 * i] = make_float2(re, im);
        }
    __syncthreads();
#pragma unroll 4
    for (int e = tid; e < 32 * 32 * 4; e += 256) {
        const int k = e & 3, pp = (e >> 2) & 31, q = e >> 7;
        if (k0 + k < K && p0 + pp < P && q0 + q < Q)
            out[(int64_t)(p0 + pp) * sOp + (int64_t)(q0 + q) * sOq + k0 + k] = Os[k * MG2_OPLANE + q * MG2_OS + pp];
    }
}

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
    static const bool gemm_v1 = sb_env_flag("SB200_MODES_GEMM_V1");          // experiments: first-generation kernel
    if (!gemm_v1) {
        dim3 grid((K + 3) / 4, pb, qb);
        sb_launch(modes_gemm2_kernel, grid, 256, 0, st, A2, sAr, sAp, B2, sBr, sBq, O2, sOp, sOq, P, Q, R, K, conjA, conjB);
        SB_LAUNCH_CHECK();
        return 0;
    }
    // prefer 8 modes per block; fall back to 4 when that would leave SMs idle
    const int64_t blocks8 = (int64_t)((K + 7) / 8) * pb * qb;
    if (blocks8 >= 2 * 148) {
        dim3 grid((K + 7) / 8, pb, qb);
        sb_launch(modes_gemm_kernel<2>, grid, 256, 0, st, A2, sAr, sAp, B2, sBr, sBq, O2, sOp, sOq, P, Q, R, K, conjA, conjB);
    } else {
        dim3 grid((K + 3) / 4, pb, qb);
        sb_launch(modes_gemm_kernel<1>, grid, 256, 0, st, A2, sAr, sAp, B2, sBr, sBq, O2, sOp, sOq, P, Q, R, K, conjA, conjB);
    }
    SB_LAUNCH_CHECK();
    return 0;
}
