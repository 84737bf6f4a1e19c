// Channels-last kernels for FourCastNet's AFNO2D filter
// (reference: src/nsbench/models/fourcastnet/fourcastnet.py:77-126).
//
//   cl_rowdft_fwd   x[rows,W,C]            -> T[rows,Mx,C]   complex   (truncated real DFT along w)
//   cl_coldft       in[B,I,Mx,C] complex   -> out[B,J,Mx,C]            (complex DFT along h, fwd or inv)
//   cl_rowidft_res  Phi[rows,Mx,C] complex -> y[rows,W,C] = synth + resid
//   blocklinear     block-diagonal complex linear layer (+bias, +ReLU / softshrink), its data-gradient
//                   (conjugate-transposed weights, activation mask applied on load) and weight gradient
//
// fp32 CUDA-core kernels; the channel index is the coalesced dimension everywhere.
#include "common.cuh"

// ======================================================================================
// cl_rowdft_fwd
// ======================================================================================
constexpr int CLR_WC = 32;   // w-chunk staged in smem

template <int KG, int CPT>
__global__ void __launch_bounds__(128)
cl_rowdft_fwd_kernel(const float* __restrict__ x, const float2* __restrict__ tab /*[W][Mx]*/,
                     float2* __restrict__ T, int W, int Mx, int C) {
    __shared__ float2 ts[CLR_WC][KG];
    const int64_t row = blockIdx.x;
    const int c0 = (blockIdx.y * 128 + threadIdx.x) * CPT;
    const bool active = c0 < C;
    const float* xr = x + row * (int64_t)W * C;
    for (int kb = 0; kb < Mx; kb += KG) {
        float2 acc[KG][CPT];
#pragma unroll
        for (int k = 0; k < KG; ++k)
#pragma unroll
            for (int c = 0; c < CPT; ++c) acc[k][c] = make_float2(0.f, 0.f);
        // real input: samples n and W - n share the cosine and negate the sine, so only e = x[n] + x[W-n] (real part)
        // and o = x[n] - x[W-n] (imaginary part) enter, n <= W/2: half the multiply-adds
        const int nh = W / 2;                                    // last folded sample
        for (int w0 = 0; w0 <= nh; w0 += CLR_WC) {
            __syncthreads();
            for (int idx = threadIdx.x; idx < CLR_WC * KG; idx += 128) {
                const int ww = idx / KG, k = idx % KG;
                float2 v = make_float2(0.f, 0.f);
                if (w0 + ww <= nh && kb + k < Mx) v = __ldg(tab + (int64_t)(w0 + ww) * Mx + kb + k);
                ts[ww][k] = v;
            }
            __syncthreads();
            if (active) {
                const int wn = min(CLR_WC, nh + 1 - w0);
                for (int ww = 0; ww < wn; ++ww) {
                    const int n = w0 + ww;
                    const bool paired = n > 0 && 2 * n != W;     // x[0] and (W even) x[W/2] have no partner
                    float e[CPT], o[CPT];
                    if (CPT == 2) {
                        const float2 v = __ldg(reinterpret_cast<const float2*>(xr + (int64_t)n * C + c0));
                        float2 u = make_float2(0.f, 0.f);
                        if (paired) u = __ldg(reinterpret_cast<const float2*>(xr + (int64_t)(W - n) * C + c0));
                        e[0] = v.x + u.x; e[CPT - 1] = v.y + u.y;
                        o[0] = paired ? v.x - u.x : 0.f; o[CPT - 1] = paired ? v.y - u.y : 0.f;
                    } else {
                        const float v = __ldg(xr + (int64_t)n * C + c0);
                        const float u = paired ? __ldg(xr + (int64_t)(W - n) * C + c0) : 0.f;
                        e[0] = v + u; o[0] = paired ? v - u : 0.f;
                    }
#pragma unroll
                    for (int k = 0; k < KG; ++k) {
                        const float2 t = ts[ww][k];
#pragma unroll
                        for (int c = 0; c < CPT; ++c) {
                            acc[k][c].x = fmaf(e[c], t.x, acc[k][c].x);
                            acc[k][c].y = fmaf(o[c], t.y, acc[k][c].y);
                        }
                    }
                }
            }
        }
        if (active) {
#pragma unroll
            for (int k = 0; k < KG; ++k) {
                if (kb + k >= Mx) break;
                float2* o = T + (row * Mx + kb + k) * (int64_t)C + c0;
                if (CPT == 2) *reinterpret_cast<float4*>(o) = make_float4(acc[k][0].x, acc[k][0].y, acc[k][CPT - 1].x, acc[k][CPT - 1].y);
                else o[0] = acc[k][0];
            }
        }
    }
}

template <int KG, int CPT>
static int launch_cl_rowdft(const float* x, const float2* tab, float2* T, int64_t rows, int W, int Mx, int C,
                            cudaStream_t st) {
    dim3 grid((unsigned)rows, (unsigned)((C / CPT + 127) / 128));
    sb_launch(cl_rowdft_fwd_kernel<KG, CPT>, grid, 128, 0, st, x, tab, T, W, Mx, C);
    SB_LAUNCH_CHECK();
    return 0;
}

extern "C" int sb200_cl_rowdft_fwd(sb200_plan_t p, int pass, const float* x, float* T, int64_t rows, int C, void* stream) {
    SB_REQUIRE(p && x && T, "cl_rowdft_fwd: NULL argument");
    SB_REQUIRE(pass == 0 || pass == 1, "cl_rowdft_fwd: pass must be 0 or 1");
    SB_REQUIRE(rows < (1LL << 31), "cl_rowdft_fwd: too many rows");
    if (rows <= 0 || C <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    float2* To = reinterpret_cast<float2*>(T);
    const float2* tab = p->rowF[pass];
    const int W = p->W, Mx = p->Mx;
    if (C % 2 == 0) {
        if (Mx <= 5) return launch_cl_rowdft<5, 2>(x, tab, To, rows, W, Mx, C, st);
        if (Mx <= 9) return launch_cl_rowdft<9, 2>(x, tab, To, rows, W, Mx, C, st);
        return launch_cl_rowdft<17, 2>(x, tab, To, rows, W, Mx, C, st);
    }
    if (Mx <= 9) return launch_cl_rowdft<9, 1>(x, tab, To, rows, W, Mx, C, st);
    return launch_cl_rowdft<17, 1>(x, tab, To, rows, W, Mx, C, st);
}

// ======================================================================================
// cl_rowidft_res:  y[row][w][c] = sum_kx Phi[row][kx][c] (.) RI[kx][w]  (+ resid)
// ======================================================================================
template <int KG, int CPT>
__global__ void __launch_bounds__(128)
cl_rowidft_res_kernel(const float2* __restrict__ Phi, const float2* __restrict__ tab /*[Mx][W]*/,
                      const float* __restrict__ resid, const float* __restrict__ resid2, float* __restrict__ y, int W, int Mx,
                      int C, int accumulate) {
    extern __shared__ float2 tsm[];     // [KG][W]
    const int64_t row = blockIdx.x;
    const int c0 = (blockIdx.y * 128 + threadIdx.x) * CPT;
    const bool active = c0 < C;
    const int kb = blockIdx.z * KG;
    for (int idx = threadIdx.x; idx < KG * W; idx += 128) {
        const int k = idx / W, w = idx % W;
        float2 v = make_float2(0.f, 0.f);
        if (kb + k < Mx) v = __ldg(tab + (int64_t)(kb + k) * W + w);
        tsm[idx] = v;
    }
    __syncthreads();
    if (!active) return;
    float2 ph[KG][CPT];
#pragma unroll
    for (int k = 0; k < KG; ++k) {
        if (kb + k < Mx) {
            const float2* s = Phi + (row * Mx + kb + k) * (int64_t)C + c0;
            if (CPT == 2) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(s));
                ph[k][0] = make_float2(v.x, v.y); ph[k][CPT - 1] = make_float2(v.z, v.w);
            } else {
                ph[k][0] = __ldg(s);
            }
        } else {
#pragma unroll
            for (int c = 0; c < CPT; ++c) ph[k][c] = make_float2(0.f, 0.f);
        }
    }
    // real output: y[w] = A + B and y[W - w] = A - B with A = sum Re(Phi) cos-part, B = sum Im(Phi) sin-part (the cosine is
    // even in w, the sine odd), so one pass over the modes gives two samples
    auto emit = [&](int w, const float (&v)[CPT]) {
        float o[CPT];
        const int64_t off = (row * W + w) * (int64_t)C + c0;
#pragma unroll
        for (int c = 0; c < CPT; ++c) o[c] = v[c];
        if (resid != nullptr && blockIdx.z == 0) {
            if (CPT == 2) { const float2 r = __ldg(reinterpret_cast<const float2*>(resid + off)); o[0] += r.x; o[CPT - 1] += r.y; }
            else o[0] += __ldg(resid + off);
        }
        if (resid2 != nullptr && blockIdx.z == 0) {      // second skip connection of the FourCastNet block (double_skip)
            if (CPT == 2) { const float2 r = __ldg(reinterpret_cast<const float2*>(resid2 + off)); o[0] += r.x; o[CPT - 1] += r.y; }
            else o[0] += __ldg(resid2 + off);
        }
        if (CPT == 2) *reinterpret_cast<float2*>(y + off) = make_float2(o[0], o[CPT - 1]);
        else y[off] = o[0];
    };
    for (int w = 0; w <= W / 2; ++w) {
        float A[CPT], Bv[CPT];
#pragma unroll
        for (int c = 0; c < CPT; ++c) { A[c] = 0.f; Bv[c] = 0.f; }
#pragma unroll
        for (int k = 0; k < KG; ++k) {
            const float2 t = tsm[k * W + w];
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                A[c] = fmaf(ph[k][c].x, t.x, A[c]);
                Bv[c] = fmaf(ph[k][c].y, t.y, Bv[c]);
            }
        }
        float v[CPT];
#pragma unroll
        for (int c = 0; c < CPT; ++c) v[c] = A[c] + Bv[c];
        emit(w, v);
        if (w > 0 && 2 * w != W) {
#pragma unroll
            for (int c = 0; c < CPT; ++c) v[c] = A[c] - Bv[c];
            emit(W - w, v);
        }
    }
}

template <int KG, int CPT>
static int launch_cl_rowidft(const float2* Phi, const float2* tab, const float* resid, const float* resid2, float* y,
                             int64_t rows, int W, int Mx, int C, cudaStream_t st) {
    const size_t smem = (size_t)KG * W * sizeof(float2);
    SB_REQUIRE(smem <= 160 * 1024, "cl_rowidft_res: W=%d too large", W);
    if (smem > 48 * 1024)
        SB_CHECK_CUDA(cudaFuncSetAttribute(cl_rowidft_res_kernel<KG, CPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)rows, (unsigned)((C / CPT + 127) / 128), 1);
    sb_launch(cl_rowidft_res_kernel<KG, CPT>, grid, 128, smem, st, Phi, tab, resid, resid2, y, W, Mx, C, 0);
    SB_LAUNCH_CHECK();
    return 0;
}

extern "C" int sb200_cl_rowidft_res2(sb200_plan_t p, int pass, const float* Phi, const float* resid, const float* resid2,
                                     float* y, int64_t rows, int C, void* stream);
extern "C" int sb200_cl_rowidft_res(sb200_plan_t p, int pass, const float* Phi, const float* resid, float* y,
                                    int64_t rows, int C, void* stream) {
    return sb200_cl_rowidft_res2(p, pass, Phi, resid, nullptr, y, rows, C, stream);
}
extern "C" int sb200_cl_rowidft_res2(sb200_plan_t p, int pass, const float* Phi, const float* resid, const float* resid2,
                                     float* y, int64_t rows, int C, void* stream) {
    SB_REQUIRE(p && Phi && y, "cl_rowidft_res: NULL argument");
    SB_REQUIRE(pass == 0 || pass == 1, "cl_rowidft_res: pass must be 0 or 1");
    SB_REQUIRE(rows < (1LL << 31), "cl_rowidft_res: too many rows");
    if (rows <= 0 || C <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const float2* Ph = reinterpret_cast<const float2*>(Phi);
    const float2* tab = p->rowI[pass];
    const int W = p->W, Mx = p->Mx;
    SB_REQUIRE(Mx <= 33, "cl_rowidft_res: Mx=%d > 33 retained columns is not implemented", Mx);
    if (C % 2 == 0 && Mx <= 17) {
        if (Mx <= 5) return launch_cl_rowidft<5, 2>(Ph, tab, resid, resid2, y, rows, W, Mx, C, st);
        if (Mx <= 9) return launch_cl_rowidft<9, 2>(Ph, tab, resid, resid2, y, rows, W, Mx, C, st);
        return launch_cl_rowidft<17, 2>(Ph, tab, resid, resid2, y, rows, W, Mx, C, st);
    }
    if (Mx <= 9) return launch_cl_rowidft<9, 1>(Ph, tab, resid, resid2, y, rows, W, Mx, C, st);
    if (Mx <= 17) return launch_cl_rowidft<17, 1>(Ph, tab, resid, resid2, y, rows, W, Mx, C, st);
    return launch_cl_rowidft<33, 1>(Ph, tab, resid, resid2, y, rows, W, Mx, C, st);
}

// ======================================================================================
// cl_coldft: out[b][j][kx][c] = sum_i tab[j][i] * in[b][i][kx][c]   (complex), j<J, i<I
// ======================================================================================
constexpr int CLC_JG = 16;
constexpr int CLC_IC = 32;

__global__ void __launch_bounds__(128)
cl_coldft_kernel(const float2* __restrict__ in, const float2* __restrict__ tab /*[J][I]*/, float2* __restrict__ out,
                 int I, int J, int Mx, int C) {
    __shared__ float2 ts[CLC_JG][CLC_IC + 1];
    const int b = blockIdx.x / Mx, kx = blockIdx.x % Mx;
    const int c = blockIdx.y * 128 + threadIdx.x;
    const int j0 = blockIdx.z * CLC_JG;
    const bool active = c < C;
    float2 acc[CLC_JG];
#pragma unroll
    for (int j = 0; j < CLC_JG; ++j) acc[j] = make_float2(0.f, 0.f);
    for (int i0 = 0; i0 < I; i0 += CLC_IC) {
        __syncthreads();
        for (int idx = threadIdx.x; idx < CLC_JG * CLC_IC; idx += 128) {
            const int j = idx / CLC_IC, ii = idx % CLC_IC;
            float2 v = make_float2(0.f, 0.f);
            if (j0 + j < J && i0 + ii < I) v = __ldg(tab + (int64_t)(j0 + j) * I + i0 + ii);
            ts[j][ii] = v;
        }
        __syncthreads();
        if (active) {
            const int in_ = min(CLC_IC, I - i0);
            for (int ii = 0; ii < in_; ++ii) {
                const float2 t = __ldg(in + (((int64_t)b * I + i0 + ii) * Mx + kx) * C + c);
#pragma unroll
                for (int j = 0; j < CLC_JG; ++j) cmac(acc[j], ts[j][ii], t);
            }
        }
    }
    if (active) {
#pragma unroll
        for (int j = 0; j < CLC_JG; ++j)
            if (j0 + j < J) out[(((int64_t)b * J + j0 + j) * Mx + kx) * C + c] = acc[j];
    }
}

// Radix-2 folds of the column transform over a grid height H that is even (tab[j][i] = s exp(-+ 2 pi i ky_j y / H)):
//   FOLD_IN  (analysis, I = H):  samples y and y + H/2 share their twiddles up to (-1)^ky, so the even frequencies see
//            in[y] + in[y + H/2] and the odd ones in[y] - in[y + H/2] over half the samples
//   FOLD_OUT (synthesis, J = H): out[y] = E + O and out[y + H/2] = +-(E - O) with E / O the sums over the even- / odd-indexed
//            retained frequencies: one pass over the modes gives two rows
// `first_odd`: the first retained frequency ky0 is odd (the parity then alternates with the index).
template <int FOLD_IN>
__global__ void __launch_bounds__(128)
cl_coldft_fold_kernel(const float2* __restrict__ in, const float2* __restrict__ tab /*[J][I]*/, float2* __restrict__ out,
                      int I, int J, int Mx, int C, int first_odd) {
    __shared__ float2 ts[CLC_JG][CLC_IC + 1];
    const int b = blockIdx.x / Mx, kx = blockIdx.x % Mx;
    const int c = blockIdx.y * 128 + threadIdx.x;
    const bool active = c < C;
    float2 acc[CLC_JG];
#pragma unroll
    for (int j = 0; j < CLC_JG; ++j) acc[j] = make_float2(0.f, 0.f);
    if (FOLD_IN) {
        const int j0 = blockIdx.z * CLC_JG;
        const int Ih = I / 2;
        for (int i0 = 0; i0 < Ih; i0 += CLC_IC) {
            __syncthreads();
            for (int idx = threadIdx.x; idx < CLC_JG * CLC_IC; idx += 128) {
                const int j = idx / CLC_IC, ii = idx % CLC_IC;
                float2 v = make_float2(0.f, 0.f);
                if (j0 + j < J && i0 + ii < Ih) v = __ldg(tab + (int64_t)(j0 + j) * I + i0 + ii);
                ts[j][ii] = v;
            }
            __syncthreads();
            if (active) {
                const int in_ = min(CLC_IC, Ih - i0);
                for (int ii = 0; ii < in_; ++ii) {
                    const float2 t1 = __ldg(in + (((int64_t)b * I + i0 + ii) * Mx + kx) * C + c);
                    const float2 t2 = __ldg(in + (((int64_t)b * I + i0 + ii + Ih) * Mx + kx) * C + c);
                    const float2 sm = make_float2(t1.x + t2.x, t1.y + t2.y), df = make_float2(t1.x - t2.x, t1.y - t2.y);
                    const float2 ve = first_odd ? df : sm, vo = first_odd ? sm : df;      // for even / odd output index
#pragma unroll
                    for (int j = 0; j < CLC_JG; ++j) cmac(acc[j], ts[j][ii], (j & 1) ? vo : ve);
                }
            }
        }
        if (active) {
#pragma unroll
            for (int j = 0; j < CLC_JG; ++j)
                if (j0 + j < J) out[(((int64_t)b * J + j0 + j) * Mx + kx) * C + c] = acc[j];
        }
    } else {
        // acc[2r] / acc[2r + 1]: sums over the even- / odd-indexed inputs for output row j0 + r (r < CLC_JG / 2) of the first half
        constexpr int RH = CLC_JG / 2;
        const int j0 = blockIdx.z * RH;
        const int Jh = J / 2;
        for (int i0 = 0; i0 < I; i0 += CLC_IC) {
            __syncthreads();
            for (int idx = threadIdx.x; idx < RH * CLC_IC; idx += 128) {
                const int j = idx / CLC_IC, ii = idx % CLC_IC;
                float2 v = make_float2(0.f, 0.f);
                if (j0 + j < Jh && i0 + ii < I) v = __ldg(tab + (int64_t)(j0 + j) * I + i0 + ii);
                ts[j][ii] = v;
            }
            __syncthreads();
            if (active) {
                const int in_ = min(CLC_IC, I - i0);               // CLC_IC is even: the index parity is that of ii
                int ii = 0;
                for (; ii + 1 < in_; ii += 2) {
                    const float2 ta = __ldg(in + (((int64_t)b * I + i0 + ii) * Mx + kx) * C + c);
                    const float2 tb = __ldg(in + (((int64_t)b * I + i0 + ii + 1) * Mx + kx) * C + c);
#pragma unroll
                    for (int r = 0; r < RH; ++r) {
                        cmac(acc[2 * r], ts[r][ii], ta);
                        cmac(acc[2 * r + 1], ts[r][ii + 1], tb);
                    }
                }
                if (ii < in_) {
                    const float2 ta = __ldg(in + (((int64_t)b * I + i0 + ii) * Mx + kx) * C + c);
#pragma unroll
                    for (int r = 0; r < RH; ++r) cmac(acc[2 * r], ts[r][ii], ta);
                }
            }
        }
        if (active) {
            const float fs = first_odd ? -1.f : 1.f;
#pragma unroll
            for (int r = 0; r < RH; ++r) {
                if (j0 + r < Jh) {
                    const float2 e = acc[2 * r], o = acc[2 * r + 1];
                    out[(((int64_t)b * J + j0 + r) * Mx + kx) * C + c] = make_float2(e.x + o.x, e.y + o.y);
                    out[(((int64_t)b * J + j0 + r + Jh) * Mx + kx) * C + c] = make_float2(fs * (e.x - o.x), fs * (e.y - o.y));
                }
            }
        }
    }
}

// fold: 0 none, 1 inputs (I = H even), 2 outputs (J = H even)
static int cl_coldft(const float2* in, const float2* tab, float2* out, int B, int I, int J, int Mx, int C, int fold, int first_odd,
                     cudaStream_t st) {
    SB_REQUIRE((int64_t)B * Mx < (1LL << 31), "cl_coldft: too many columns");
    if (fold == 1) {
        dim3 grid((unsigned)(B * Mx), (unsigned)((C + 127) / 128), (unsigned)((J + CLC_JG - 1) / CLC_JG));
        sb_launch(cl_coldft_fold_kernel<1>, grid, 128, 0, st, in, tab, out, I, J, Mx, C, first_odd);
    } else if (fold == 2) {
        dim3 grid((unsigned)(B * Mx), (unsigned)((C + 127) / 128), (unsigned)((J / 2 + CLC_JG / 2 - 1) / (CLC_JG / 2)));
        sb_launch(cl_coldft_fold_kernel<0>, grid, 128, 0, st, in, tab, out, I, J, Mx, C, first_odd);
    } else {
        dim3 grid((unsigned)(B * Mx), (unsigned)((C + 127) / 128), (unsigned)((J + CLC_JG - 1) / CLC_JG));
        sb_launch(cl_coldft_kernel, grid, 128, 0, st, in, tab, out, I, J, Mx, C);
    }
    SB_LAUNCH_CHECK();
    return 0;
}

extern "C" int sb200_cl_coldft_fwd(sb200_plan_t p, int pass, const float* T, float* Xh, int B, int C, void* stream) {
    SB_REQUIRE(p && T && Xh, "cl_coldft_fwd: NULL argument");
    SB_REQUIRE(pass == 0 || pass == 1, "cl_coldft_fwd: pass must be 0 or 1");
    if (B <= 0 || C <= 0) return 0;
    return cl_coldft(reinterpret_cast<const float2*>(T), p->colF[pass], reinterpret_cast<float2*>(Xh), B, p->H, p->My,
                     p->Mx, C, p->H % 2 == 0 ? 1 : 0, ((p->ky0 % 2) + 2) % 2, (cudaStream_t)stream);
}

extern "C" int sb200_cl_coldft_inv(sb200_plan_t p, int pass, const float* Yh, float* Phi, int B, int C, void* stream) {
    SB_REQUIRE(p && Yh && Phi, "cl_coldft_inv: NULL argument");
    SB_REQUIRE(pass == 0 || pass == 1, "cl_coldft_inv: pass must be 0 or 1");
    if (B <= 0 || C <= 0) return 0;
    return cl_coldft(reinterpret_cast<const float2*>(Yh), p->colI[pass], reinterpret_cast<float2*>(Phi), B, p->My, p->H,
                     p->Mx, C, p->H % 2 == 0 ? 2 : 0, ((p->ky0 % 2) + 2) % 2, (cudaStream_t)stream);
}

// ======================================================================================
// block-diagonal complex linear layer
// ======================================================================================
__device__ __forceinline__ float act_mask(float v, int kind) {   // derivative of the activation given its OUTPUT
    return kind == 1 ? (v > 0.f ? 1.f : 0.f) : (kind == 2 ? (v != 0.f ? 1.f : 0.f) : 1.f);
}
__device__ __forceinline__ float apply_act(float v, int kind, float lam) {
    if (kind == 1) return v > 0.f ? v : 0.f;
    if (kind == 2) return v > lam ? v - lam : (v < -lam ? v + lam : 0.f);
    return v;
}

struct BlParams {
    const float2* in;        // [ntok][nb*Ni]
    const float2* mask_src;  // same shape as `in` or NULL; in *= act'(mask_src) on load
    int mask_kind;
    const float* Wr; const float* Wi;     // (n, i, o) at n*sWn + i*sWi + o*sWo
    int64_t sWn, sWi, sWo;
    int conjW;
    const float* br; const float* bi;     // [nb][No] or NULL
    float2* out;             // [ntok][nb*No]
    int act; float lam;
    int64_t ntok; int nb, Ni, No;
};

constexpr int BL_TOK = 64, BL_KC = 16, BL_OT = 32;

__global__ void __launch_bounds__(256)
blocklinear_kernel(const BlParams p) {
    __shared__ float2 Xs[BL_TOK][BL_KC + 1];
    __shared__ __align__(16) float Wrs[BL_KC][BL_OT];
    __shared__ __align__(16) float Wis[BL_KC][BL_OT];
    const int tid = threadIdx.x;
    const int tl = tid & 63, og = tid >> 6;
    const int64_t t0 = (int64_t)blockIdx.x * BL_TOK;
    const int n = blockIdx.y;
    const int o0 = blockIdx.z * BL_OT;
    const int64_t in_ld = (int64_t)p.nb * p.Ni, out_ld = (int64_t)p.nb * p.No;
    float2 acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = make_float2(0.f, 0.f);

    for (int i0 = 0; i0 < p.Ni; i0 += BL_KC) {
        __syncthreads();
        for (int idx = tid; idx < BL_TOK * BL_KC; idx += 256) {
            const int tt = idx / BL_KC, ii = idx % BL_KC;
            float2 v = make_float2(0.f, 0.f);
            if (t0 + tt < p.ntok && i0 + ii < p.Ni) {
                const int64_t off = (t0 + tt) * in_ld + (int64_t)n * p.Ni + i0 + ii;
                v = __ldg(p.in + off);
                if (p.mask_src) {
                    const float2 m = __ldg(p.mask_src + off);
                    v.x *= act_mask(m.x, p.mask_kind);
                    v.y *= act_mask(m.y, p.mask_kind);
                }
            }
            Xs[tt][ii] = v;
        }
        for (int idx = tid; idx < BL_KC * BL_OT; idx += 256) {
            const int ii = idx / BL_OT, oo = idx % BL_OT;
            float wr = 0.f, wi = 0.f;
            if (i0 + ii < p.Ni && o0 + oo < p.No) {
                const int64_t off = (int64_t)n * p.sWn + (int64_t)(i0 + ii) * p.sWi + (int64_t)(o0 + oo) * p.sWo;
                wr = __ldg(p.Wr + off);
                wi = __ldg(p.Wi + off);
                if (p.conjW) wi = -wi;
            }
            Wrs[ii][oo] = wr;
            Wis[ii][oo] = wi;
        }
        __syncthreads();
#pragma unroll
        for (int ii = 0; ii < BL_KC; ++ii) {
            const float2 x = Xs[tl][ii];
            const float4 r0 = *reinterpret_cast<const float4*>(&Wrs[ii][og * 8]);
            const float4 r1 = *reinterpret_cast<const float4*>(&Wrs[ii][og * 8 + 4]);
            const float4 q0 = *reinterpret_cast<const float4*>(&Wis[ii][og * 8]);
            const float4 q1 = *reinterpret_cast<const float4*>(&Wis[ii][og * 8 + 4]);
            const float wr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
            const float wi[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) cmac(acc[j], x, make_float2(wr[j], wi[j]));
        }
    }
    if (t0 + tl >= p.ntok) return;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int o = o0 + og * 8 + j;
        if (o >= p.No) continue;
        float2 v = acc[j];
        if (p.br) { v.x += __ldg(p.br + (int64_t)n * p.No + o); v.y += __ldg(p.bi + (int64_t)n * p.No + o); }
        v.x = apply_act(v.x, p.act, p.lam);
        v.y = apply_act(v.y, p.act, p.lam);
        p.out[(t0 + tl) * out_ld + (int64_t)n * p.No + o] = v;
    }
}

static int launch_blocklinear(const BlParams& p, cudaStream_t st) {
    if (p.ntok <= 0) return 0;
    SB_REQUIRE(ceil_div64(p.ntok, BL_TOK) < (1LL << 31), "blocklinear: too many tokens");
    dim3 grid((unsigned)ceil_div64(p.ntok, BL_TOK), (unsigned)p.nb, (unsigned)((p.No + BL_OT - 1) / BL_OT));
    sb_launch(blocklinear_kernel, grid, 256, 0, st, p);
    SB_LAUNCH_CHECK();
    return 0;
}

// forward: out = act(in (*) W + b);   w [2,nb,Ni,No] planar (index 0 = real, 1 = imag), b [2,nb,No]
extern "C" int sb200_afno_blocklinear_fwd(const float* in, const float* w, const float* b, float* out, int64_t ntok,
                                          int nb, int Ni, int No, int act, float lam, void* stream) {
    SB_REQUIRE(in && w && out, "afno_blocklinear_fwd: NULL argument");
    BlParams p;
    p.in = reinterpret_cast<const float2*>(in); p.mask_src = nullptr; p.mask_kind = 0;
    p.Wr = w; p.Wi = w + (int64_t)nb * Ni * No;
    p.sWn = (int64_t)Ni * No; p.sWi = No; p.sWo = 1; p.conjW = 0;
    p.br = b; p.bi = b ? b + (int64_t)nb * No : nullptr;
    p.out = reinterpret_cast<float2*>(out); p.act = act; p.lam = lam;
    p.ntok = ntok; p.nb = nb; p.Ni = Ni; p.No = No;
    return launch_blocklinear(p, (cudaStream_t)stream);
}

// data gradient: gin[t,n,i] = sum_o (gout[t,n,o] * act'(fwd_out[t,n,o])) (*) conj(W[n,i,o])
extern "C" int sb200_afno_blocklinear_dgrad(const float* gout, const float* fwd_out, int mask_kind, const float* w,
                                            float* gin, int64_t ntok, int nb, int Ni, int No, void* stream) {
    SB_REQUIRE(gout && w && gin, "afno_blocklinear_dgrad: NULL argument");
    BlParams p;
    p.in = reinterpret_cast<const float2*>(gout);
    p.mask_src = reinterpret_cast<const float2*>(fwd_out); p.mask_kind = mask_kind;
    p.Wr = w; p.Wi = w + (int64_t)nb * Ni * No;
    p.sWn = (int64_t)Ni * No; p.sWi = 1; p.sWo = No; p.conjW = 1;     // contraction index is o, output index is i
    p.br = nullptr; p.bi = nullptr;
    p.out = reinterpret_cast<float2*>(gin); p.act = 0; p.lam = 0.f;
    p.ntok = ntok; p.nb = nb; p.Ni = No; p.No = Ni;
    return launch_blocklinear(p, (cudaStream_t)stream);
}

// ---- weight gradient: gW[n,i,o] = sum_t conj(a[t,n,i]) * g'[t,n,o];  gb[n,o] = sum_t g'[t,n,o] ----
constexpr int BW_T = 16;

__global__ void __launch_bounds__(256)
blocklinear_wgrad_kernel(const float2* __restrict__ a, const float2* __restrict__ g, const float2* __restrict__ mask_src,
                         int mask_kind, float* __restrict__ ws, int64_t ntok, int64_t chunk_tok, int nb, int Ni, int No,
                         int o_tiles) {
    __shared__ __align__(16) float2 As[BW_T][32];
    __shared__ __align__(16) float2 Gs[BW_T][32];
    const int tid = threadIdx.x;
    const int ti = tid >> 4, to = tid & 15;          // 16 x 16 threads, 2 i x 2 o each
    const int chunk = blockIdx.x, n = blockIdx.y;
    const int i0 = (blockIdx.z / o_tiles) * 32, o0 = (blockIdx.z % o_tiles) * 32;
    const int64_t tc0 = (int64_t)chunk * chunk_tok;
    int64_t tc1 = tc0 + chunk_tok;
    if (tc1 > ntok) tc1 = ntok;
    const int64_t a_ld = (int64_t)nb * Ni, g_ld = (int64_t)nb * No;
    float2 acc[2][2];
    float2 accb[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        accb[i] = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 2; ++j) acc[i][j] = make_float2(0.f, 0.f);
    }
    for (int64_t t0 = tc0; t0 < tc1; t0 += BW_T) {
        __syncthreads();
        for (int idx = tid; idx < BW_T * 32; idx += 256) {
            const int tt = idx >> 5, cc = idx & 31;
            float2 va = make_float2(0.f, 0.f), vg = va;
            if (t0 + tt < tc1) {
                if (i0 + cc < Ni) {
                    va = __ldg(a + (t0 + tt) * a_ld + (int64_t)n * Ni + i0 + cc);
                    va.y = -va.y;
                }
                if (o0 + cc < No) {
                    const int64_t off = (t0 + tt) * g_ld + (int64_t)n * No + o0 + cc;
                    vg = __ldg(g + off);
                    if (mask_src) {
                        const float2 m = __ldg(mask_src + off);
                        vg.x *= act_mask(m.x, mask_kind);
                        vg.y *= act_mask(m.y, mask_kind);
                    }
                }
            }
            As[tt][cc] = va;
            Gs[tt][cc] = vg;
        }
        __syncthreads();
#pragma unroll
        for (int tt = 0; tt < BW_T; ++tt) {
            const float4 av = *reinterpret_cast<const float4*>(&As[tt][ti * 2]);
            const float4 gv = *reinterpret_cast<const float4*>(&Gs[tt][to * 2]);
            const float2 a2[2] = {make_float2(av.x, av.y), make_float2(av.z, av.w)};
            const float2 g2[2] = {make_float2(gv.x, gv.y), make_float2(gv.z, gv.w)};
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) cmac(acc[i][j], a2[i], g2[j]);
            accb[0].x += g2[0].x; accb[0].y += g2[0].y;
            accb[1].x += g2[1].x; accb[1].y += g2[1].y;
        }
    }
    // workspace layout per chunk: [2][nb][Ni][No] planar weights followed by [2][nb][No] planar bias
    const int64_t wsz = (int64_t)nb * Ni * No, bsz = (int64_t)nb * No;
    float* wsc = ws + (int64_t)chunk * (2 * wsz + 2 * bsz);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int ii = i0 + ti * 2 + i;
        if (ii >= Ni) continue;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int oo = o0 + to * 2 + j;
            if (oo >= No) continue;
            const int64_t off = ((int64_t)n * Ni + ii) * No + oo;
            wsc[off] = acc[i][j].x;
            wsc[wsz + off] = acc[i][j].y;
        }
    }
    if (ti == 0 && i0 == 0) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int oo = o0 + to * 2 + j;
            if (oo >= No) continue;
            wsc[2 * wsz + (int64_t)n * No + oo] = accb[j].x;
            wsc[2 * wsz + bsz + (int64_t)n * No + oo] = accb[j].y;
        }
    }
}

__global__ void __launch_bounds__(256)
chunk_reduce_kernel(const float* __restrict__ ws, float* __restrict__ out0, int64_t n0, float* __restrict__ out1,
                    int64_t n1, int nchunks) {
    const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int64_t E = n0 + n1;
    if (e >= E) return;
    float s = 0.f;
#pragma unroll 8
    for (int c = 0; c < nchunks; ++c) s += __ldg(ws + (int64_t)c * E + e);
    if (e < n0) out0[e] = s;
    else out1[e - n0] = s;
}

static void bw_chunking(int64_t ntok, int nb, int64_t* chunk_tok, int* nchunks) {
    int64_t target = (2 * 148 + nb - 1) / nb;
    if (target < 1) target = 1;
    int64_t ct = (ntok + target - 1) / target;
    ct = (ct + BW_T - 1) / BW_T * BW_T;
    if (ct < BW_T) ct = BW_T;
    *chunk_tok = ct;
    *nchunks = (int)((ntok + ct - 1) / ct);
}

extern "C" int64_t sb200_afno_blocklinear_wgrad_workspace(int64_t ntok, int nb, int Ni, int No) {
    int64_t ct; int nc;
    bw_chunking(ntok, nb, &ct, &nc);
    return (int64_t)nc * (2 * (int64_t)nb * Ni * No + 2 * (int64_t)nb * No);
}

// gw [2,nb,Ni,No], gb [2,nb,No] (planar, same layout as the parameters)
extern "C" int sb200_afno_blocklinear_wgrad(const float* a, const float* gout, const float* fwd_out, int mask_kind,
                                            float* gw, float* gb, int64_t ntok, int nb, int Ni, int No,
                                            float* workspace, void* stream) {
    SB_REQUIRE(a && gout && gw && gb && workspace, "afno_blocklinear_wgrad: NULL argument");
    if (ntok <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    int64_t ct; int nc;
    bw_chunking(ntok, nb, &ct, &nc);
    const int i_tiles = (Ni + 31) / 32, o_tiles = (No + 31) / 32;
    dim3 grid(nc, nb, i_tiles * o_tiles);
    sb_launch(blocklinear_wgrad_kernel, grid, 256, 0, st, reinterpret_cast<const float2*>(a),
                                                   reinterpret_cast<const float2*>(gout),
                                                   reinterpret_cast<const float2*>(fwd_out), mask_kind, workspace, ntok,
                                                   ct, nb, Ni, No, o_tiles);
    SB_LAUNCH_CHECK();
    const int64_t n0 = 2 * (int64_t)nb * Ni * No, n1 = 2 * (int64_t)nb * No;
    sb_launch(chunk_reduce_kernel, (unsigned)ceil_div64(n0 + n1, 256), 256, 0, st, workspace, gw, n0, gb, n1, nc);
    SB_LAUNCH_CHECK();
    return 0;
}

// ======================================================================================
// Real embedding of the complex block weights for the tensor-core path (tc_gemm.cu, sb200_gemm_batched):
//   out[(j,c'')] = sum_{(i,c)} in[(i,c)] * E[(j,c'')][(i,c)]  reproduces  out = in * (wr + i wi)  on interleaved (re, im) data
// ======================================================================================
__global__ void __launch_bounds__(256) afno_embed_kernel(const float* __restrict__ w, float* __restrict__ E, int nb, int Ni, int No) {
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;      // over E: [nb][No][2][Ni][2]
    const int64_t total = (int64_t)nb * No * 2 * Ni * 2;
    if (idx >= total) return;
    const int c = (int)(idx & 1);
    const int i = (int)((idx >> 1) % Ni);
    const int cc = (int)((idx / (2 * Ni)) & 1);
    const int j = (int)((idx / (4 * Ni)) % No);
    const int b = (int)(idx / ((int64_t)4 * Ni * No));
    const int64_t plane = (int64_t)nb * Ni * No;
    const int64_t wi_ = ((int64_t)b * Ni + i) * No + j;
    const float wr = __ldg(w + wi_), wim = __ldg(w + plane + wi_);
    E[idx] = cc == 0 ? (c == 0 ? wr : -wim) : (c == 0 ? wim : wr);
}
__global__ void __launch_bounds__(256) afno_unembed_kernel(const float* __restrict__ gE, float* __restrict__ gw, int nb, int Ni, int No) {
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;      // over gw planes: [nb][Ni][No]
    const int64_t plane = (int64_t)nb * Ni * No;
    if (idx >= plane) return;
    const int j = (int)(idx % No);
    const int i = (int)((idx / No) % Ni);
    const int b = (int)(idx / ((int64_t)Ni * No));
    const float* e = gE + (((int64_t)b * No + j) * 2) * Ni * 2;      // E[b][j][cc][i][c]
    const float e00 = __ldg(e + (0 * Ni + i) * 2 + 0), e01 = __ldg(e + (0 * Ni + i) * 2 + 1);
    const float e10 = __ldg(e + (1 * Ni + i) * 2 + 0), e11 = __ldg(e + (1 * Ni + i) * 2 + 1);
    gw[idx] = e00 + e11;
    gw[plane + idx] = e10 - e01;
}
__global__ void __launch_bounds__(256) mask_mul_kernel(const float* __restrict__ g, const float* __restrict__ src, float* __restrict__ out,
                                                       int64_t n, int kind) {
    const int64_t i = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
    if (i + 3 < n) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(g + i)), s = __ldg(reinterpret_cast<const float4*>(src + i));
        float4 o;
        o.x = (kind == 1 ? s.x > 0.f : s.x != 0.f) ? a.x : 0.f;
        o.y = (kind == 1 ? s.y > 0.f : s.y != 0.f) ? a.y : 0.f;
        o.z = (kind == 1 ? s.z > 0.f : s.z != 0.f) ? a.z : 0.f;
        o.w = (kind == 1 ? s.w > 0.f : s.w != 0.f) ? a.w : 0.f;
        *reinterpret_cast<float4*>(out + i) = o;
    } else {
        for (int64_t k = i; k < n; ++k) out[k] = (kind == 1 ? src[k] > 0.f : src[k] != 0.f) ? g[k] : 0.f;
    }
}
extern "C" int sb200_afno_embed(const float* w, float* E, int nb, int Ni, int No, void* stream) {
    SB_REQUIRE(w && E && nb > 0 && Ni > 0 && No > 0, "afno_embed: bad argument");
    const int64_t total = (int64_t)nb * No * 2 * Ni * 2;
    sb_launch(afno_embed_kernel, (unsigned)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream, w, E, nb, Ni, No);
    SB_LAUNCH_CHECK();
    return 0;
}
extern "C" int sb200_afno_unembed(const float* gE, float* gw, int nb, int Ni, int No, void* stream) {
    SB_REQUIRE(gE && gw && nb > 0 && Ni > 0 && No > 0, "afno_unembed: bad argument");
    const int64_t plane = (int64_t)nb * Ni * No;
    sb_launch(afno_unembed_kernel, (unsigned)ceil_div64(plane, 256), 256, 0, (cudaStream_t)stream, gE, gw, nb, Ni, No);
    SB_LAUNCH_CHECK();
    return 0;
}
extern "C" int sb200_mask_mul(const float* g, const float* src, float* out, int64_t n, int kind, void* stream) {
    SB_REQUIRE(g && src && out && n > 0 && (kind == 1 || kind == 2), "mask_mul: bad argument");
    SB_REQUIRE(((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
               "mask_mul: pointers must be 16-byte aligned");
    sb_launch(mask_mul_kernel, (unsigned)ceil_div64(n, 1024), 256, 0, (cudaStream_t)stream, g, src, out, n, kind);
    SB_LAUNCH_CHECK();
    return 0;
}
