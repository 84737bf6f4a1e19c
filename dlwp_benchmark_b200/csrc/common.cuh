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
// Two elements at a time on the packed fp32x2 pipe (FFMA2 / FMUL2): a three-register FFMA issues every second
// cycle per scheduler, the packed form does two elements in the same slot, so the pair costs ~11 fma-pipe slots
// instead of ~26 and the evaluation becomes bound by its four MUFU ops.  Same A-S 7.1.26 form as gelu_core (the
// constants are folded: 0.3275911 / sqrt 2, -log2(e) / 2, the 0.5 of erfc into the coefficients).
__device__ __forceinline__ unsigned long long sb_pk(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float2 sb_upk(unsigned long long v) {
    float2 d;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(v));
    return d;
}
__device__ __forceinline__ unsigned long long sb_fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long sb_mul2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
#define SB_K2(c) sb_pk((c), (c))
__device__ __forceinline__ void gelu_core2(float z0, float z1, float2& cdf, float2& e) {
    const unsigned long long den = sb_fma2(sb_pk(fabsf(z0), fabsf(z1)), SB_K2(0.23164188861846924f), SB_K2(1.0f));
    const float2 d = sb_upk(den);
    const unsigned long long t = sb_pk(sb_rcp_approx(d.x), sb_rcp_approx(d.y));
    unsigned long long poly = sb_fma2(t, SB_K2(0.5f * 1.061405429f), SB_K2(0.5f * -1.453152027f));
    poly = sb_fma2(poly, t, SB_K2(0.5f * 1.421413741f));
    poly = sb_fma2(poly, t, SB_K2(0.5f * -0.284496736f));
    poly = sb_fma2(poly, t, SB_K2(0.5f * 0.254829592f));
    poly = sb_mul2(poly, t);
    const unsigned long long zz = sb_pk(z0, z1);
    const float2 arg = sb_upk(sb_mul2(sb_mul2(zz, SB_K2(-0.72134752044448170368f)), zz));
    e = make_float2(sb_ex2_approx(arg.x), sb_ex2_approx(arg.y));          // exp(-z^2 / 2)
    const unsigned long long ht = sb_mul2(poly, sb_pk(e.x, e.y));          // 0.5 * erfc(|z| / sqrt 2)
    const float2 h = sb_upk(ht), o = sb_upk(sb_fma2(ht, SB_K2(-1.0f), SB_K2(1.0f)));
    cdf = make_float2(z0 < 0.f ? h.x : o.x, z1 < 0.f ? h.y : o.y);
}
// a <- GELU(a), b <- GELU(b)
__device__ __forceinline__ void gelu2(float& a, float& b) {
    float2 cdf, e;
    gelu_core2(a, b, cdf, e);
    const float2 r = sb_upk(sb_mul2(sb_pk(a, b), sb_pk(cdf.x, cdf.y)));
    a = r.x; b = r.y;
}
// ga <- GELU'(za), gb <- GELU'(zb)
__device__ __forceinline__ void gelu_grad2(float za, float zb, float& ga, float& gb) {
    float2 cdf, e;
    gelu_core2(za, zb, cdf, e);
    const unsigned long long zs = sb_mul2(sb_pk(za, zb), SB_K2(0.39894228040143267794f));
    const float2 r = sb_upk(sb_fma2(zs, sb_pk(e.x, e.y), sb_pk(cdf.x, cdf.y)));
    ga = r.x; gb = r.y;
}
__device__ __forceinline__ float4 gelu4(float4 v) {
    gelu2(v.x, v.y);
    gelu2(v.z, v.w);
    return v;
}
__device__ __forceinline__ float4 gelu_grad4(const float4 z) {
    float4 g;
    gelu_grad2(z.x, z.y, g.x, g.y);
    gelu_grad2(z.z, z.w, g.z, g.w);
    return g;
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
