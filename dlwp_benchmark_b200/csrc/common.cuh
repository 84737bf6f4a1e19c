// Shared helpers for the spectral_b200 kernel library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "spectral_b200.h"

struct sb200_plan_s {
    int H, W, ky0, My, Mx;
    double scale_fwd, scale_inv;
    int device;
    // table sets indexed by pass (0 = forward-pass transforms, 1 = their adjoints)
    float2* rowF[2];   // [W][Mx]   (a*cos, -a*sin)
    float2* colF[2];   // [My][H]   s*exp(-i th)
    float2* colI[2];   // [H][My]   s*exp(+i th)
    float2* rowI[2];   // [Mx][W]   (a*cos, -a*sin)
    // planar / padded copies for the tensor-core kernels are appended by tc_plan.cu
    void* tc;          // opaque (sb200_tc_tables*) or NULL
};

void sb200_set_error(const char* fmt, ...);

#define SB_CHECK_CUDA(expr)                                                              \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            sb200_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,           \
                            cudaGetErrorString(_e));                                     \
            return 1;                                                                    \
        }                                                                                \
    } while (0)

#define SB_REQUIRE(cond, ...)                                                            \
    do {                                                                                 \
        if (!(cond)) {                                                                   \
            sb200_set_error(__VA_ARGS__);                                                \
            return 2;                                                                    \
        }                                                                                \
    } while (0)

#define SB_LAUNCH_CHECK()                                                                \
    do {                                                                                 \
        cudaError_t _e = cudaGetLastError();                                             \
        if (_e != cudaSuccess) {                                                         \
            sb200_set_error("%s:%d: kernel launch failed: %s", __FILE__, __LINE__,       \
                            cudaGetErrorString(_e));                                     \
            return 3;                                                                    \
        }                                                                                \
    } while (0)

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// exact (erf) GELU and its derivative -- matches torch.nn.functional.gelu(approximate='none')
__device__ __forceinline__ float gelu_f(float z) {
    return 0.5f * z * (1.0f + erff(z * 0.70710678118654752440f));
}
__device__ __forceinline__ float gelu_grad_f(float z) {
    const float cdf = 0.5f * (1.0f + erff(z * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * expf(-0.5f * z * z);
    return cdf + z * pdf;
}

__device__ __forceinline__ void cmac(float2& acc, const float2 a, const float2 b) {
    acc.x = fmaf(a.x, b.x, acc.x);
    acc.x = fmaf(-a.y, b.y, acc.x);
    acc.y = fmaf(a.x, b.y, acc.y);
    acc.y = fmaf(a.y, b.x, acc.y);
}

// parameters of the fused row-synthesis + pointwise kernel (pointwise.cu / tc_*.cu)
struct PwParams {
    const float2* Phi;   // [B,N,H,Mx] complex
    const float2* RI;    // [Mx][W]
    const float* A;      // [B,M,H,W] or NULL
    const float* Wp;     // (n,m) at n*w_sn + m*w_sm, or NULL
    int64_t w_sn, w_sm;
    const float* bias;   // [N] or NULL
    const float* zprev;  // [B,N,H,W] or NULL (mode 1)
    float* z_out;        // [B,N,H,W] or NULL (mode 0)
    float* y_out;        // [B,N,H,W]
    int B, M, N, H, W, Mx;
    int mode, apply_act;
    int nrows_max;       // rows of Phi staged per tile
};
