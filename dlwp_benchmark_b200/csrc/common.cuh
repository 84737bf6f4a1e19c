// Shared helpers for the spectral_b200 kernel library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

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

// tables for the tcgen05 row-synthesis (tc_plan.cu); NULL in the plan when the grid is not tileable
struct sb200_tc_tables {
    int R, V, K2, K2pad;
    float* E[2];      // [K2pad][128]
    float2* rot;      // [V][Mx]
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

// every kernel launch of the library is followed by exactly one SB_LAUNCH_CHECK: it also counts launches
// (sb200_kernel_launches(), reported by bench.py as gpu_launches)
extern unsigned long long g_sb200_launches;

#define SB_LAUNCH_CHECK()                                                                \
    do {                                                                                 \
        __atomic_fetch_add(&g_sb200_launches, 1ULL, __ATOMIC_RELAXED);                   \
        cudaError_t _e = cudaGetLastError();                                             \
        if (_e != cudaSuccess) {                                                         \
            sb200_set_error("%s:%d: kernel launch failed: %s", __FILE__, __LINE__,       \
                            cudaGetErrorString(_e));                                     \
            return 3;                                                                    \
        }                                                                                \
    } while (0)

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Precision mode of the tensor-core stages (0 = CUDA-core fp32 kernels only, 1 = single-pass TF32, 3 = 3xTF32 parity
// mode).  It is an ARGUMENT of every entry point that has tensor-core kernels behind it: the library keeps no mode of
// its own.  Inside a call the value travels in a thread-local that the entry point sets and restores (SbModeScope), so
// internal helpers need no extra parameter and concurrent calls from different threads cannot see each other's mode.
extern thread_local int sb200_tl_mode;
struct SbModeScope {
    int prev;
    explicit SbModeScope(int m) : prev(sb200_tl_mode) { sb200_tl_mode = (m == 0 || m == 1 || m == 3) ? m : 3; }
    ~SbModeScope() { sb200_tl_mode = prev; }
};
static inline int sb_tc_mode() { return sb200_tl_mode; }

// SM count of the CURRENT device (per-device cache; the library is used from one process per GPU, but nothing
// stops a caller from driving several devices from one process).  Defined in plan.cu.
int sb200_num_sms();

// Experiment switches are compiled out of the shipped library: they exist only in bring-up builds
// (make BRINGUP=1 -> -DSB200_BRINGUP); a production build always takes the default.
static inline int sb_env_int(const char* name, int dflt) {
#ifdef SB200_BRINGUP
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
#else
    (void)name;
    return dflt;
#endif
}
static inline bool sb_env_flag(const char* name) { return sb_env_int(name, 0) != 0; }

// Every kernel of the library is launched through sb_launch() (one place to add launch attributes).
// Programmatic dependent launch was tried here (session 5) and removed: the kernels read their inputs through the
// non-coherent path (__ldg / const __restrict__ -> LDG.E.CONSTANT), which is only legal for data that is read-only for
// the whole lifetime of the grid; with PDL a grid's lifetime starts before its producer has finished, the L1
// invalidation of its launch happens too early, and stale lines were observed (wrong results in 2 parity tests).
// It also measured 8 % slower on the cfg2 train step (early-resident CTAs of the next kernels crowd the SMs).
#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
static inline void sb_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cfg.attrs = nullptr;
    cfg.numAttrs = 0;
    (void)cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);   // errors surface in SB_LAUNCH_CHECK
}
#endif

// GELU(z) = z * Phi(z) (the exact erf form of torch.nn.functional.gelu(approximate='none')) and its
// derivative Phi(z) + z * phi(z).  erf is evaluated with the Abramowitz-Stegun 7.1.26 rational form
// (|abs error| <= 1.5e-7 in exact arithmetic, ~2.5e-7 in fp32) built on MUFU rcp / ex2: about 15
// instructions instead of ~45 for erff + expf, which matters because every activation element of the
// FNO passes through one of these in a fused epilogue.  Phi for z < 0 is formed without cancellation.
// Both share exp(-z^2/2), so the derivative costs three more instructions than the value.
__device__ __forceinline__ float sb_rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float sb_ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void gelu_core(float z, float& cdf, float& e) {
    const float x = fabsf(z) * 0.70710678118654752440f;
    const float t = sb_rcp_approx(fmaf(0.3275911f, x, 1.0f));
    float poly = fmaf(t, 1.061405429f, -1.453152027f);
    poly = fmaf(poly, t, 1.421413741f);
    poly = fmaf(poly, t, -0.284496736f);
    poly = fmaf(poly, t, 0.254829592f);
    poly *= t;
    e = sb_ex2_approx(-1.4426950408889634f * x * x);   // exp(-z^2 / 2)
    const float half_tail = 0.5f * poly * e;            // 0.5 * erfc(|z| / sqrt 2)
    cdf = z < 0.f ? half_tail : 1.0f - half_tail;
}
__device__ __forceinline__ float gelu_f(float z) {
    float cdf, e;
    gelu_core(z, cdf, e);
    return z * cdf;
}
__device__ __forceinline__ float gelu_grad_f(float z) {
    float cdf, e;
    gelu_core(z, cdf, e);
    return fmaf(z * 0.39894228040143267794f, e, cdf);
}

// Packed fp32x2 FMA (Blackwell FFMA2): d = a * b + c on both halves with ONE issue slot.  ptxas folds operand
// broadcasts (x, x), half swaps and the (-lo, +hi) sign pattern into FFMA2 operand modifiers, so a complex
// multiply-accumulate costs two instructions instead of four.
__device__ __forceinline__ float2 ffma2(const float2 a, const float2 b, const float2 c) {
    unsigned long long ra, rb, rc, rd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    float2 d;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
    return d;
}
// acc += a * b (complex), two FFMA2
__device__ __forceinline__ void cmac2(float2& acc, const float2 a, const float2 b) {
    acc = ffma2(make_float2(a.x, a.x), b, acc);
    acc = ffma2(make_float2(a.y, a.y), make_float2(-b.y, b.x), acc);
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
