// Channels-last token kernels around the FourCastNet block (reference: src/dlwpbench/models/fourcastnet/fourcastnet.py
// :156-193 ``Block`` = LN -> AFNO2D -> +res -> LN -> Mlp -> +res; src/nsbench/.../fourcastnet.py:129-165):
//   LayerNorm over the channel axis, forward and backward (HBM-bound: one pass over the tokens each),
//   column sums [T, N] -> [N] (bias gradients of the token Linear layers), batch sums [B, n] -> [n] (pos_embed gradient).
// All reductions run in a fixed order (deterministic).
#include "common.cuh"

namespace {

constexpr int LN_MAXC = 1024;         // C <= 32 lanes * 4 floats * 8 register slots

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// one warp per token; y = (x - mean) * rstd * gamma + beta   (biased variance, like torch.nn.LayerNorm)
// LN_MAXV = float4 register slots per lane (ceil(C / 128)): a template parameter so that narrow C does not pay for 8
template <int LN_MAXV>
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
template <int LN_MAXV>
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

// out[j] = sum_r part[r][j], j < ncol (columns [0, n_each) -> out0, the rest -> out1): 32 columns x 8 row lanes per block (coalesced 128-byte row reads, 4 loads
// in flight per thread), fixed-order tree at the end (deterministic)
__global__ void __launch_bounds__(256) rows_reduce_kernel(const float* __restrict__ part, float* __restrict__ out0,
                                                          float* __restrict__ out1, int rows, int ncol, int n_each) {
    __shared__ float red[8][33];
    const int c = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + c;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (j < ncol) {
        int r = rl;
        for (; r + 24 < rows; r += 32) {
            s0 += part[(int64_t)r * ncol + j];
            s1 += part[(int64_t)(r + 8) * ncol + j];
            s2 += part[(int64_t)(r + 16) * ncol + j];
            s3 += part[(int64_t)(r + 24) * ncol + j];
        }
        for (; r < rows; r += 8) s0 += part[(int64_t)r * ncol + j];
    }
    red[rl][c] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (rl == 0 && j < ncol) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += red[k][c];
        if (j < n_each) out0[j] = t;
        else if (out1) out1[j - n_each] = t;
    }
}

// column sums of a row-major [T, N] matrix: block (bx, by) sums its contiguous row range of 128 columns (float4 per lane,
// 8 row lanes, 4 rows in flight) -> part[bx][N]
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ a, int64_t lda, float* __restrict__ part,
                                                             int64_t T, int N, int vec) {
    __shared__ float4 red[8][33];
    const int64_t per = (T + gridDim.x - 1) / gridDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * per, t1 = min(T, t0 + per);
    const int c = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int n = blockIdx.y * 128 + c * 4;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < N) {
        for (int64_t t = t0 + rl; t < t1; t += 8) {
            const float* src = a + t * lda + n;
            if (vec) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(src));
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            } else {
                s.x += __ldg(src);
                if (n + 1 < N) s.y += __ldg(src + 1);
                if (n + 2 < N) s.z += __ldg(src + 2);
                if (n + 3 < N) s.w += __ldg(src + 3);
            }
        }
    }
    red[rl][c] = s;
    __syncthreads();
    if (rl == 0 && n < N) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 8; ++k) { t.x += red[k][c].x; t.y += red[k][c].y; t.z += red[k][c].z; t.w += red[k][c].w; }
        float* dst = part + (int64_t)blockIdx.x * N + n;
        dst[0] = t.x;
        if (n + 1 < N) dst[1] = t.y;
        if (n + 2 < N) dst[2] = t.z;
        if (n + 3 < N) dst[3] = t.w;
    }
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
    const int cap = sb200_num_sms() * 4;
    return (int)(want < cap ? want : cap);
}
int ln_bwd_grid(int64_t T) {              // fewer, longer blocks: every block leaves one row of partial sums to reduce
    const int64_t want = (T + 63) / 64;
    const int cap = sb200_num_sms() * 2;
    return (int)(want < cap ? want : cap);
}

#define LN_DISPATCH(C, CALL)                                   \
    do {                                                       \
        const int _nv = ((C) + 127) / 128;                     \
        if (_nv <= 1) { constexpr int NV = 1; CALL; }          \
        else if (_nv <= 2) { constexpr int NV = 2; CALL; }     \
        else if (_nv <= 3) { constexpr int NV = 3; CALL; }     \
        else if (_nv <= 4) { constexpr int NV = 4; CALL; }     \
        else if (_nv <= 6) { constexpr int NV = 6; CALL; }     \
        else { constexpr int NV = 8; CALL; }                   \
    } while (0)

}  // namespace

extern "C" int sb200_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd,
                                   int64_t T, int C, float eps, void* stream) {
    SB_REQUIRE(x && gamma && beta && y && mean && rstd, "layernorm_fwd: NULL argument");
    SB_REQUIRE(C % 4 == 0 && C >= 4 && C <= LN_MAXC, "layernorm_fwd: C=%d must be a multiple of 4 and <= %d", C, LN_MAXC);
    if (T <= 0) return 0;
    LN_DISPATCH(C, sb_launch(ln_fwd_kernel<NV>, (unsigned)ln_grid(T), 256, 0, (cudaStream_t)stream, x, gamma, beta, y, mean, rstd,
                             T, C, eps));
    SB_LAUNCH_CHECK();
    return 0;
}

extern "C" int64_t sb200_layernorm_bwd_workspace(int64_t T, int C) { return (int64_t)ln_bwd_grid(T) * 2 * C; }

extern "C" int sb200_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* mean, const float* rstd,
                                   const float* dres, float* dx, float* dgamma, float* dbeta, float* workspace, int64_t T,
                                   int C, void* stream) {
    SB_REQUIRE(dy && x && gamma && mean && rstd && dx && dgamma && dbeta && workspace, "layernorm_bwd: NULL argument");
    SB_REQUIRE(C % 4 == 0 && C >= 4 && C <= LN_MAXC, "layernorm_bwd: C=%d must be a multiple of 4 and <= %d", C, LN_MAXC);
    if (T <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = ln_bwd_grid(T);
    const size_t smem = (size_t)8 * 2 * C * sizeof(float);
    LN_DISPATCH(C, {
        SB_CHECK_CUDA(cudaFuncSetAttribute(ln_bwd_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sb_launch(ln_bwd_kernel<NV>, (unsigned)grid, 256, smem, st, dy, x, gamma, mean, rstd, dres, dx, workspace, T, C);
    });
    SB_LAUNCH_CHECK();
    sb_launch(rows_reduce_kernel, (unsigned)((2 * C + 31) / 32), 256, 0, st, (const float*)workspace, dgamma, dbeta, grid, 2 * C, C);
    SB_LAUNCH_CHECK();
    return 0;
}

static int colsum_blocks(int64_t T, int N) {
    int64_t blocks = (T + 63) / 64;
    const int ny = (N + 127) / 128;
    int cap = sb200_num_sms() * 4 / ny;
    if (cap < 16) cap = 16;
    if (blocks > cap) blocks = cap;
    return (int)blocks;
}
extern "C" int64_t sb200_colsum_workspace(int64_t T, int N) { return (int64_t)colsum_blocks(T, N) * N; }

extern "C" int sb200_colsum(const float* a, int64_t lda, float* out, int64_t T, int N, float* workspace, void* stream) {
    SB_REQUIRE(a && out && workspace, "colsum: NULL argument");
    SB_REQUIRE(T > 0 && N > 0, "colsum: non-positive size");
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = colsum_blocks(T, N);
    const int vec = ((reinterpret_cast<uintptr_t>(a) & 15) == 0 && (lda & 3) == 0 && (N & 3) == 0) ? 1 : 0;
    sb_launch(colsum_partial_kernel, dim3((unsigned)blocks, (unsigned)((N + 127) / 128)), 256, 0, st, a, lda, workspace, T, N, vec);
    SB_LAUNCH_CHECK();
    // rows_reduce with n_each = N and no second output: columns [0, N) of a [blocks][N] table
    sb_launch(rows_reduce_kernel, (unsigned)((N + 31) / 32), 256, 0, st, (const float*)workspace, out, (float*)nullptr, blocks,
              N, N);
    SB_LAUNCH_CHECK();
    return 0;
}

extern "C" int sb200_batch_sum(const float* a, float* out, int B, int64_t n, void* stream) {
    SB_REQUIRE(a && out && B > 0 && n > 0, "batch_sum: bad argument");
    sb_launch(batch_sum_kernel, (unsigned)ceil_div64(n, 256), 256, 0, (cudaStream_t)stream, a, out, B, n);
    SB_LAUNCH_CHECK();
    return 0;
}
