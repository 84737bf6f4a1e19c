// Channels-last token kernels around the FourCastNet block (reference: src/dlwpbench/models/fourcastnet/fourcastnet.py
// :156-193 ``Block`` = LN -> AFNO2D -> +res -> LN -> Mlp -> +res; src/nsbench/.../fourcastnet.py:129-165):
//   LayerNorm over the channel axis, forward and backward (HBM-bound: one pass over the tokens each),
//   column sums [T, N] -> [N] (bias gradients of the token Linear layers), batch sums [B, n] -> [n] (pos_embed gradient).
// All reductions run in a fixed order (deterministic).
#include "common.cuh"

namespace {

constexpr int LN_MAXV = 8;            // float4 per lane kept in registers: C <= 32 * 4 * 8 = 1024

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// one warp per token; y = (x - mean) * rstd * gamma + beta   (biased variance, like torch.nn.LayerNorm)
__global__ void __launch_bounds__(256) ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, float* __restrict__ y,
                                                     float* __restrict__ mean_out, float* __restrict__ rstd_out, int64_t T, int C,
                                                     float eps) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = (int64_t)gridDim.x * 8;
    const int nv = C / 128, rem = C - nv * 128;          // full 128-channel groups (+ a tail group of rem channels)
    for (int64_t t = warp0; t < T; t += nwarps) {
        const float* xs = x + t * C;
        float4 v[LN_MAXV];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < LN_MAXV; ++i) {
            const int c = i * 128 + lane * 4;
            v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < nv || (i == nv && lane * 4 < rem)) {
                v[i] = __ldg(reinterpret_cast<const float4*>(xs + c));
                s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
            }
        }
        const float mean = warp_sum(s) / (float)C;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < LN_MAXV; ++i) {
            if (i < nv || (i == nv && lane * 4 < rem)) {
                const float a = v[i].x - mean, b = v[i].y - mean, c2 = v[i].z - mean, d = v[i].w - mean;
                q += (a * a + b * b) + (c2 * c2 + d * d);
            }
        }
        const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
        if (lane == 0) { mean_out[t] = mean; rstd_out[t] = rstd; }
        float* ys = y + t * C;
#pragma unroll
        for (int i = 0; i < LN_MAXV; ++i) {
            const int c = i * 128 + lane * 4;
            if (i < nv || (i == nv && lane * 4 < rem)) {
                const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c)), b = __ldg(reinterpret_cast<const float4*>(beta + c));
                float4 o;
                o.x = fmaf((v[i].x - mean) * rstd, g.x, b.x);
                o.y = fmaf((v[i].y - mean) * rstd, g.y, b.y);
                o.z = fmaf((v[i].z - mean) * rstd, g.z, b.z);
                o.w = fmaf((v[i].w - mean) * rstd, g.w, b.w);
                *reinterpret_cast<float4*>(ys + c) = o;
            }
        }
    }
}

// dx = rstd * (g - mean_c(g) - xhat * mean_c(g * xhat)),  g = dy * gamma;  (+ dres: a gradient that bypasses the LN --
// the residual branch of the block -- added in the same pass);  per-block partial sums of dgamma = dy * xhat, dbeta = dy
__global__ void __launch_bounds__(256) ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                     const float* __restrict__ gamma, const float* __restrict__ mean_in,
                                                     const float* __restrict__ rstd_in, const float* __restrict__ dres,
                                                     float* __restrict__ dx, float* __restrict__ part, int64_t T, int C) {
    extern __shared__ float sm[];                        // [8 warps][2][C]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int nv = C / 128, rem = C - nv * 128;
    // contiguous token range of this block, warps interleaved inside it (fixed assignment: deterministic)
    const int64_t per = (T + gridDim.x - 1) / gridDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * per, t1 = min(T, t0 + per);
    float4 ag[LN_MAXV], ab[LN_MAXV];
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) { ag[i] = make_float4(0.f, 0.f, 0.f, 0.f); ab[i] = ag[i]; }
    for (int64_t t = t0 + w; t < t1; t += 8) {
        const float mean = __ldg(mean_in + t), rstd = __ldg(rstd_in + t);
        float4 xh[LN_MAXV], g[LN_MAXV];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < LN_MAXV; ++i) {
            const int c = i * 128 + lane * 4;
            if (i < nv || (i == nv && lane * 4 < rem)) {
                const float4 xv = __ldg(reinterpret_cast<const float4*>(x + t * C + c));
                const float4 dv = __ldg(reinterpret_cast<const float4*>(dy + t * C + c));
                const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma + c));
                xh[i] = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
                g[i] = make_float4(dv.x * gm.x, dv.y * gm.y, dv.z * gm.z, dv.w * gm.w);
                s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
                s2 += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
                ag[i].x = fmaf(dv.x, xh[i].x, ag[i].x); ag[i].y = fmaf(dv.y, xh[i].y, ag[i].y);
                ag[i].z = fmaf(dv.z, xh[i].z, ag[i].z); ag[i].w = fmaf(dv.w, xh[i].w, ag[i].w);
                ab[i].x += dv.x; ab[i].y += dv.y; ab[i].z += dv.z; ab[i].w += dv.w;
            }
        }
        const float m1 = warp_sum(s1) / (float)C, m2 = warp_sum(s2) / (float)C;
#pragma unroll
        for (int i = 0; i < LN_MAXV; ++i) {
            const int c = i * 128 + lane * 4;
            if (i < nv || (i == nv && lane * 4 < rem)) {
                float4 o;
                o.x = rstd * (g[i].x - m1 - xh[i].x * m2);
                o.y = rstd * (g[i].y - m1 - xh[i].y * m2);
                o.z = rstd * (g[i].z - m1 - xh[i].z * m2);
                o.w = rstd * (g[i].w - m1 - xh[i].w * m2);
                if (dres) {
                    const float4 r = __ldg(reinterpret_cast<const float4*>(dres + t * C + c));
                    o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
                }
                *reinterpret_cast<float4*>(dx + t * C + c) = o;
            }
        }
    }
    // block partials: warps -> shared -> one row of `part` ([grid][2][C])
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
        const int c = i * 128 + lane * 4;
        if (i < nv || (i == nv && lane * 4 < rem)) {
            *reinterpret_cast<float4*>(sm + (w * 2 + 0) * C + c) = ag[i];
            *reinterpret_cast<float4*>(sm + (w * 2 + 1) * C + c) = ab[i];
        }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < 2 * C; idx += 256) {
        float s = 0.f;
        for (int k = 0; k < 8; ++k) s += sm[k * 2 * C + idx];
        part[(int64_t)blockIdx.x * 2 * C + idx] = s;
    }
}

// out[j] = sum_r part[r][j], j < n  (rows added in order)
__global__ void __launch_bounds__(256) rows_reduce_kernel(const float* __restrict__ part, float* __restrict__ out0,
                                                          float* __restrict__ out1, int rows, int n_each) {
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (j >= 2 * n_each) return;
    float s = 0.f;
    for (int r = 0; r < rows; ++r) s += part[(int64_t)r * 2 * n_each + j];
    if (j < n_each) out0[j] = s;
    else if (out1) out1[j - n_each] = s;
}

// column sums of a row-major [T, N] matrix: block b sums its contiguous row range -> part[b][N]
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ a, int64_t lda, float* __restrict__ part,
                                                             int64_t T, int N) {
    const int64_t per = (T + gridDim.x - 1) / gridDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * per, t1 = min(T, t0 + per);
    for (int n = threadIdx.x; n < N; n += 256) {
        float s = 0.f;
        for (int64_t t = t0; t < t1; ++t) s += __ldg(a + t * lda + n);
        part[(int64_t)blockIdx.x * N + n] = s;
    }
}
__global__ void __launch_bounds__(256) colsum_final_kernel(const float* __restrict__ part, float* __restrict__ out, int rows, int N) {
    const int n = blockIdx.x * 256 + threadIdx.x;
    if (n >= N) return;
    float s = 0.f;
    for (int r = 0; r < rows; ++r) s += part[(int64_t)r * N + n];
    out[n] = s;
}

// out[i] = sum_b a[b][i]
__global__ void __launch_bounds__(256) batch_sum_kernel(const float* __restrict__ a, float* __restrict__ out, int B, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += __ldg(a + (int64_t)b * n + i);
    out[i] = s;
}

int ln_grid(int64_t T) {
    const int64_t want = (T + 7) / 8;
    const int cap = sb200_num_sms() * 8;
    return (int)(want < cap ? want : cap);
}

}  // namespace

extern "C" int sb200_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd,
                                   int64_t T, int C, float eps, void* stream) {
    SB_REQUIRE(x && gamma && beta && y && mean && rstd, "layernorm_fwd: NULL argument");
    SB_REQUIRE(C % 4 == 0 && C >= 4 && C <= 128 * LN_MAXV, "layernorm_fwd: C=%d must be a multiple of 4 and <= %d", C, 128 * LN_MAXV);
    if (T <= 0) return 0;
    sb_launch(ln_fwd_kernel, (unsigned)ln_grid(T), 256, 0, (cudaStream_t)stream, x, gamma, beta, y, mean, rstd, T, C, eps);
    SB_LAUNCH_CHECK();
    return 0;
}

extern "C" int64_t sb200_layernorm_bwd_workspace(int64_t T, int C) { return (int64_t)ln_grid(T) * 2 * C; }

extern "C" int sb200_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* mean, const float* rstd,
                                   const float* dres, float* dx, float* dgamma, float* dbeta, float* workspace, int64_t T,
                                   int C, void* stream) {
    SB_REQUIRE(dy && x && gamma && mean && rstd && dx && dgamma && dbeta && workspace, "layernorm_bwd: NULL argument");
    SB_REQUIRE(C % 4 == 0 && C >= 4 && C <= 128 * LN_MAXV, "layernorm_bwd: C=%d must be a multiple of 4 and <= %d", C, 128 * LN_MAXV);
    if (T <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = ln_grid(T);
    const size_t smem = (size_t)8 * 2 * C * sizeof(float);
    SB_CHECK_CUDA(cudaFuncSetAttribute(ln_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sb_launch(ln_bwd_kernel, (unsigned)grid, 256, smem, st, dy, x, gamma, mean, rstd, dres, dx, workspace, T, C);
    SB_LAUNCH_CHECK();
    sb_launch(rows_reduce_kernel, (unsigned)((2 * C + 255) / 256), 256, 0, st, (const float*)workspace, dgamma, dbeta, grid, C);
    SB_LAUNCH_CHECK();
    return 0;
}

extern "C" int64_t sb200_colsum_workspace(int64_t T, int N) {
    int64_t blocks = (T + 255) / 256;
    const int cap = sb200_num_sms() * 4;
    if (blocks > cap) blocks = cap;
    return blocks * N;
}

extern "C" int sb200_colsum(const float* a, int64_t lda, float* out, int64_t T, int N, float* workspace, void* stream) {
    SB_REQUIRE(a && out && workspace, "colsum: NULL argument");
    SB_REQUIRE(T > 0 && N > 0, "colsum: non-positive size");
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = (int)(sb200_colsum_workspace(T, N) / N);
    sb_launch(colsum_partial_kernel, (unsigned)blocks, 256, 0, st, a, lda, workspace, T, N);
    SB_LAUNCH_CHECK();
    sb_launch(colsum_final_kernel, (unsigned)((N + 255) / 256), 256, 0, st, (const float*)workspace, out, blocks, N);
    SB_LAUNCH_CHECK();
    return 0;
}

extern "C" int sb200_batch_sum(const float* a, float* out, int B, int64_t n, void* stream) {
    SB_REQUIRE(a && out && B > 0 && n > 0, "batch_sum: bad argument");
    sb_launch(batch_sum_kernel, (unsigned)ceil_div64(n, 256), 256, 0, (cudaStream_t)stream, a, out, B, n);
    SB_LAUNCH_CHECK();
    return 0;
}
