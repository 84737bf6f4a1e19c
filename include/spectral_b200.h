/*
 * spectral_b200 -- C ABI of the B200 (sm_100a) spectral-convolution kernel library.
 *
 * This is the drop-in boundary for the FNO-family hot path of amazon-science/dlwp-benchmark.
 * The reference has no FFI of its own (it is pure PyTorch); what it calls on this path are
 * library primitives.  Each entry point below names the reference call it replaces
 * (paths relative to the reference root; "neuralop" = neuraloperator @05c01c3 pinned by
 * README.md:34-35, reached from src/nsbench/models/fno/fno.py:19-27 and
 * src/dlwpbench/models/fno/fno.py:38-47,136-146).
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller (PyTorch: tensor.data_ptr()),
 *    contiguous, fp32; complex data is interleaved (re,im) exactly like torch.complex64 /
 *    torch.view_as_real
 *  - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream);
 *    the library never synchronises, never allocates activation memory and never switches
 *    device; scratch is passed in by the caller
 *  - every function returns 0 on success, non-zero on failure; sb200_last_error() returns a
 *    thread-local message.  There is NO CPU fallback: without a CUDA device every compute
 *    entry point fails.
 *  - plans hold only immutable twiddle tables (device memory) for one (H, W, ky0, My, Mx,
 *    scales) combination and may be shared by concurrent callers.
 */
#ifndef SPECTRAL_B200_H
#define SPECTRAL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sb200_plan_s* sb200_plan_t;

/* library / build identification: returns 10000*major + 100*minor + patch */
int sb200_version(void);
/* number of CUDA kernels this library has launched in this process (monotonic; for launch accounting) */
int64_t sb200_kernel_launches(void);
/* thread-local description of the last failure (never NULL) */
const char* sb200_last_error(void);
/* compute capability major*10+minor of the current device, or <0 when no device */
int sb200_device_arch(void);

/* Precision mode of the GEMM-shaped stages -- the trailing `tc_mode` argument of every entry point that has
 * tensor-core kernels behind it:  0 = CUDA-core (exact fp32 FFMA) kernels only, 1 = tcgen05 kind::tf32 single pass
 * (parity <= 1e-2 over a 20-step rollout), 3 = tcgen05 3xTF32 split (parity <= 1e-5; any other value means 3).
 * Shapes the tcgen05 kernels do not cover always run on the CUDA-core kernels.  The library keeps NO mode of its own:
 * the only mutable library state is the plan cache, a launch counter and a thread-local error string. */


/* ---- plans ---------------------------------------------------------------------------
 * Retained block: rows ky = (ky0 + j) mod H, j in [0,My); cols kx in [0,Mx).
 * FNO (neuralop SpectralConv.forward, fftshift-era slicing): ky0 = lo - H/2, norm "forward"
 *   => scale_fwd = 1/(H*W), scale_inv = 1.
 * AFNO2D (src/nsbench/models/fourcastnet/fourcastnet.py:84,92-95,123): ky0 = r0,
 *   norm "ortho" => scale_fwd = scale_inv = 1/sqrt(h*w).
 * Tables are built in double precision on the host and uploaded to the current device. */
int sb200_plan_create(sb200_plan_t* plan, int H, int W, int ky0, int My, int Mx,
                      double scale_fwd, double scale_inv);
int sb200_plan_destroy(sb200_plan_t plan);

/* `pass` selects the table set: 0 = the transforms of the forward pass (analysis un-weighted
 * * scale_fwd, synthesis with Hermitian weights 1/2/1 * scale_inv); 1 = their adjoints used
 * by the backward pass (analysis of the output gradient WITH Hermitian weights * scale_inv,
 * synthesis with weights 1 * scale_fwd). */

/* ---- channels-first (FNO) stages -------------------------------------------------------
 * replaces torch.fft.rfftn + fftshift + slice (neuralop SpectralConv.forward) */
/* x [rows, W] real  ->  T [rows, Mx] complex   (rows = B*C*H).  With tc mode 1 / 3 and W a multiple of 32 (>= 64) this
 * is a tcgen05 GEMM against the twiddle matrix (tc_rowdft.cu), else the fp32 FFMA kernel. */
int sb200_rowdft_fwd(sb200_plan_t plan, int pass, const float* x, float* T, int64_t rows, void* stream, int tc_mode);
/* T [nimg, H, Mx] complex -> Xh [nimg, My, Mx] complex */
int sb200_coldft_fwd(sb200_plan_t plan, int pass, const float* T, float* Xh, int64_t nimg, void* stream);
/* Yh [nimg, My, Mx] complex -> Phi [nimg, H, Mx] complex
 * replaces zeros + scatter + fftshift + the H-axis half of torch.fft.irfftn */
int sb200_coldft_inv(sb200_plan_t plan, int pass, const float* Yh, float* Phi, int64_t nimg, void* stream);

/* x [nimg, H, W] real -> Xh [nimg, My, Mx] complex: the two stages above in one call.  On small grids
 * (W in {32, 64}, H dividing 256) a single fused kernel keeps the row-transformed spectrum on chip and
 * `scratch` may be NULL; otherwise `scratch` must hold sb200_analysis_scratch(plan, nimg) floats (16-byte aligned)
 * and, for W and H multiples of 32 (>= 64) with 2*Mx, 2*My <= 64, both stages run as tcgen05 GEMMs with a planar
 * transposed intermediate in `scratch` (layout private to the library).
 * replaces torch.fft.rfftn + fftshift + slice (pass 0) / the adjoint of irfftn (pass 1). */
int64_t sb200_analysis_scratch(sb200_plan_t plan, int64_t nimg);
int sb200_analysis(sb200_plan_t plan, int pass, const float* x, float* Xh, int64_t nimg, float* scratch, void* stream, int tc_mode);

/* Batched-over-modes complex contraction
 *      out[p,q,k] = sum_r opA(A[r,p,k]) * opB(B[r,q,k]),   k contiguous,
 * with element strides (in complex elements) for r/p/q.  conj flags: bit0 = conj A, bit1 = conj B.
 * replaces torch.einsum("bixy,ioxy->boxy") (neuralop _contract_dense) and its two autograd
 * products (grad wrt input spectrum, grad wrt weight). */
int sb200_modes_gemm(const float* A, int64_t sAr, int64_t sAp,
                     const float* B, int64_t sBr, int64_t sBq,
                     float* out, int64_t sOp, int64_t sOq,
                     int P, int Q, int R, int K, int conj_flags, void* stream);

/* Fused pointwise MLP head (the FNO projection  C -> 256 -> 1,  neuralop FNO.projection = MLP(n_layers=2),
 * reached from src/nsbench/models/fno/fno.py:19-27):
 *      y[b,p] = sum_n w2[n] * gelu( sum_m W1[n,m] h[b,m,p] + b1[n] ) + b2
 * on the tcgen05 path; the 256-channel hidden tensor never reaches HBM.  h [B,M,HW], W1 [256,M], y [B,HW].
 * Restrictions (else non-zero return, caller uses the two-kernel path): N == 256, M % 8 == 0 (M % 32 == 0
 * above 32), HW % 4 == 0, tc mode != 0. */
int sb200_mlp_head_fwd(const float* h, const float* W1, const float* b1, const float* w2, const float* b2,
                       float* y, int B, int M, int N, int64_t HW, void* stream, int tc_mode);
/* Backward, stage 1: recomputes z1 = W1 h + b1 on the tensor cores and writes
 *      gz1[b,n,p] = w2[n] gy[b,p] gelu'(z1[b,n,p])          (input of the W1 weight / data gradient kernels)
 *      gb1[n] = sum_{b,p} gz1,   gw2[n] = sum_{b,p} gy gelu(z1),   gb2 = sum gy   (gb2 may be NULL)
 * workspace: sb200_mlp_head_bwd_workspace() floats. */
int64_t sb200_mlp_head_bwd_workspace(void);
int sb200_mlp_head_bwd(const float* h, const float* W1, const float* b1, const float* w2, const float* gy,
                       float* gz1, float* gb1, float* gw2, float* gb2, float* workspace, int B, int M, int N,
                       int64_t HW, void* stream, int tc_mode);

/* Backward of the first lifting layer when the model has ONE input channel (neuralop FNO.lifting =
 * MLP(in_channels=1, hidden=256, out=C); src/nsbench/configs/model/fno.yaml): with g = gradient wrt the lifting
 * output [B,C,HW],
 *      gz1[b,n,p] = (sum_c W2[c,n] g[b,c,p]) * gelu'(w1[n] x[b,p] + b1[n])      (recomputed from x, never stored)
 *      gb1[n] = sum_{b,p} gz1,     gw1[n] = sum_{b,p} gz1 x[b,p]
 * Same restrictions and workspace as sb200_mlp_head_bwd (N == 256, C % 8 == 0, ...). */
int sb200_lift_tail_bwd(const float* g, const float* W2, const float* w1, const float* b1, const float* x,
                        float* gw1, float* gb1, float* workspace, int B, int C, int N, int64_t HW, void* stream, int tc_mode);

/* Lifting MLP of a ONE-input-channel model, forward:   y[b,c,p] = sum_n W2[c,n] gelu(w1[n] x[b,p] + b1[n]) + b2[c]
 * (neuralop FNO.lifting = MLP(1 -> 256 -> C)).  The 256-channel hidden operand is generated tile by tile in
 * shared memory and consumed by tcgen05.mma: it never exists in HBM.  x [B,HW], W2 [C,256], y [B,C,HW];
 * N == 256, C % 16 == 0, HW % 4 == 0, tc mode != 0. */
int sb200_lift_fwd(const float* x, const float* w1, const float* b1, const float* W2, const float* b2, float* y,
                   int B, int N, int C, int64_t HW, void* stream, int tc_mode);
/* ... and the weight gradient of its second layer with the hidden activations regenerated on chip:
 *      gW2[c,n] = sum_{b,p} g[b,c,p] gelu(w1[n] x[b,p] + b1[n]),   gb2[c] = sum_{b,p} g[b,c,p]   (gb2 may be NULL)
 * workspace: sb200_pointwise_wgrad_workspace(B, C, N, HW) floats. */
int sb200_lift_wgrad(const float* g, const float* x, const float* w1, const float* b1, float* gW2, float* gb2,
                     float* workspace, int B, int C, int N, int64_t HW, void* stream, int tc_mode);

/* Strided complex GEMM   C[m,n] = sum_k opA(A[m,k]) * opB(B[k,n])   (complex64, strides in complex elements).
 * m and k may be two-level composite indices: m -> (m / M2, m % M2) addressed with (s?m1, s?m2), likewise k
 * with K2 (M2 = 1 / K2 = 1: plain index using the *2 stride).  This expresses every mode product, core /
 * input gradient and factor gradient of a Tucker tensor without permuting it.  When the output tile count is
 * small and K long the reduction is split (deterministically) and needs sb200_cgemm_workspace() floats.
 * replaces tltorch TuckerTensor reconstruction / tensorly's einsum chain behind neuralop
 * SpectralConv(factorization="Tucker") (src/dlwpbench/models/fno/fno.py:136-146) and its autograd backward. */
typedef struct sb200_cgemm_desc {
    int32_t M, N, K, M2, K2, conjA, conjB, reserved;
    int64_t sAm1, sAm2, sAk1, sAk2;
    int64_t sBk1, sBk2, sBn;
    int64_t sCm1, sCm2, sCn;
} sb200_cgemm_desc;
int64_t sb200_cgemm_workspace(const sb200_cgemm_desc* desc, int ngroups);
int sb200_cgemm(const sb200_cgemm_desc* desc, const float* A, const float* B, float* C, float* workspace, void* stream);
/* the same product for `ngroups` (1..8) operand triples of identical geometry in ONE launch (the layers of an
 * FNO); A/B/C are HOST arrays of device pointers; workspace sized by sb200_cgemm_workspace(desc, ngroups) */
int sb200_cgemm_grouped(const sb200_cgemm_desc* desc, int ngroups, const float* const* A, const float* const* B,
                        float* const* C, float* workspace, void* stream);

/* Fused row synthesis + pointwise (1x1) channel mix + epilogue.
 *   acc[b,n,y,x] = sum_kx ( Phi[b,n,y,kx].re * RI[kx][x].x + Phi.im * RI[kx][x].y )
 *                + sum_m  Wp[n,m] * A[b,m,y,x]            (skipped when Wp == NULL)
 *                + bias[n]                                  (skipped when bias == NULL)
 *   mode 0 (forward):  z_out = acc (if non-NULL);  y_out = gelu(acc) if apply_act else acc
 *   mode 1 (backward): y_out = acc * gelu'(zprev) if zprev != NULL else acc
 * Wp element (n,m) is read at Wp[n*w_sn + m*w_sm] (so the transpose is a stride swap).
 * replaces the W-axis half of irfftn + bias add + fno_skips Conv2d(1x1) + add + F.gelu
 * (neuralop SpectralConv.forward / FNOBlocks.forward_with_postactivation) and, in mode 1,
 * the corresponding autograd nodes. */
int sb200_rowidft_pointwise(sb200_plan_t plan, int pass, const float* Phi,
                            const float* A, const float* Wp, int64_t w_sn, int64_t w_sm,
                            const float* bias, const float* zprev,
                            float* z_out, float* y_out,
                            int B, int M, int N, int mode, int apply_act, void* stream, int tc_mode);

/* Weight / bias gradient of the pointwise (1x1) channel mix:
 *   gW[o,i] = sum_{b,p} g[b,o,p] * x[b,i,p]      gbias[o] = sum_{b,p} g[b,o,p]   (gbias may be NULL)
 * `workspace` must hold sb200_pointwise_wgrad_workspace(...) floats.  Deterministic (two-phase).
 * replaces the Conv2d weight-grad and bias-sum autograd nodes. */
int64_t sb200_pointwise_wgrad_workspace(int B, int Cout, int Cin, int64_t HW, int tc_mode);
int sb200_pointwise_wgrad(const float* g, const float* x, float* gW, float* gbias,
                          int B, int Cout, int Cin, int64_t HW, float* workspace, void* stream, int tc_mode);

/* Pointwise layer with few output channels (N <= 8), e.g. the FNO projection's last 1x1 conv
 * (neuralop MLP.fcs[1], 256 -> out_channels):  acc[b,n,p] = sum_m Wp[n,m] A[b,m,p] + bias[n];
 * z_out = acc (if non-NULL), y_out = gelu(acc) if apply_act else acc.  A is read once. */
int sb200_pointwise_small_n(const float* A, const float* Wp, const float* bias, float* z_out, float* y_out,
                            int B, int M, int N, int64_t HW, int apply_act, void* stream);
/* Weight gradient of a 1x1 conv when one side has few channels (S <= 16):
 *   out_dot[s,l] = sum_{b,p} small[b,s,p] * big[b,l,p]   (written as [l,s] when transpose != 0)
 *   out_small[s] = sum small,  out_big[l] = sum big       (either may be NULL)
 * small [B,S,HW], big [B,L,HW].  Deterministic two-phase reduction. */
int64_t sb200_wgrad_small_workspace(int B, int S, int L, int64_t HW);
int sb200_wgrad_small(const float* small, const float* big, float* out_dot, float* out_small, float* out_big,
                      int B, int S, int L, int64_t HW, int transpose, float* workspace, void* stream);

/* ---- channels-last (FourCastNet AFNO2D) stages ------------------------------------------
 * reference: AFNO2D.forward, src/nsbench/models/fourcastnet/fourcastnet.py:77-126
 * (identical copy src/dlwpbench/models/fourcastnet/fourcastnet.py:78-127).
 * x is [B,h,w,C]; complex tensors are [...,C,2].  Plans: sb200_plan_create(h, w, r0, r1-r0, kc,
 * 1/sqrt(hw), 1/sqrt(hw)).                                                                  */
/* x [rows,W,C] -> T [rows,Mx,C] complex (rows = B*h); replaces the w-axis half of rfft2 (:84) */
int sb200_cl_rowdft_fwd(sb200_plan_t plan, int pass, const float* x, float* T, int64_t rows, int C, void* stream);
/* T [B,H,Mx,C] -> Xh [B,My,Mx,C]; replaces the h-axis half of rfft2 + the row/col slicing (:92-95) */
int sb200_cl_coldft_fwd(sb200_plan_t plan, int pass, const float* T, float* Xh, int B, int C, void* stream);
/* Yh [B,My,Mx,C] -> Phi [B,H,Mx,C]; replaces zeros + scatter (:87-90) + h-axis half of irfft2 (:123) */
int sb200_cl_coldft_inv(sb200_plan_t plan, int pass, const float* Yh, float* Phi, int B, int C, void* stream);
/* Phi [rows,Mx,C] (+ resid [rows,W,C] or NULL) -> y [rows,W,C]; replaces the w-axis half of irfft2
 * and the residual add `x + bias` (:126) */
int sb200_cl_rowidft_res2(sb200_plan_t plan, int pass, const float* Phi, const float* resid, const float* resid2,
                          float* y, int64_t rows, int C, void* stream);   /* + a second residual (Block double_skip, :186-188) */
int sb200_cl_rowidft_res(sb200_plan_t plan, int pass, const float* Phi, const float* resid, float* y,
                         int64_t rows, int C, void* stream);
/* Block-diagonal complex linear layer: out[t,n,o] = act( sum_i in[t,n,i] * W[n,i,o] + b[n,o] ),
 * w [2,nb,Ni,No] and b [2,nb,No] planar (index 0 = real part, 1 = imaginary part) exactly like the
 * reference parameters w1/b1/w2/b2 (:72-75); act: 0 none, 1 ReLU on re and im separately (:95-105),
 * 2 softshrink(lam) on re and im (:121).  replaces the 8 einsums + relu/softshrink. */
int sb200_afno_blocklinear_fwd(const float* in, const float* w, const float* b, float* out, int64_t ntok,
                               int nb, int Ni, int No, int act, float lam, void* stream);
/* gin[t,n,i] = sum_o (gout[t,n,o] * act'(fwd_out[t,n,o])) * conj(W[n,i,o]); mask_kind as `act`;
 * fwd_out may be NULL (no activation) */
int sb200_afno_blocklinear_dgrad(const float* gout, const float* fwd_out, int mask_kind, const float* w,
                                 float* gin, int64_t ntok, int nb, int Ni, int No, void* stream);
/* gw[.,n,i,o] = sum_t conj(a[t,n,i]) * (gout*act')[t,n,o];  gb[.,n,o] = sum_t (gout*act')[t,n,o] */
int64_t sb200_afno_blocklinear_wgrad_workspace(int64_t ntok, int nb, int Ni, int No);
int sb200_afno_blocklinear_wgrad(const float* a, const float* gout, const float* fwd_out, int mask_kind,
                                 float* gw, float* gb, int64_t ntok, int nb, int Ni, int No,
                                 float* workspace, void* stream);

/* Element-wise helpers used between layers: y = gelu(z);  gz = gy * gelu'(z) */
int sb200_gelu_fwd(const float* z, float* y, int64_t n, void* stream);
int sb200_gelu_bwd(const float* gy, const float* z, float* gz, int64_t n, void* stream);

/* ---- channels-last token path (FourCastNet block remainder: src/dlwpbench/models/fourcastnet/fourcastnet.py:42-57
 * Mlp, :156-193 Block, :283-293 PatchEmbed / head; src/nsbench/models/fourcastnet/fourcastnet.py:129-165) ----------
 *
 * General fp32 GEMM on the tcgen05 tensor cores:  D[m][n] = sum_k A(m,k) B(n,k)  (row-major D, leading dim ldd).
 *   a_mn = 0: A is K-major, element (m,k) at A[m*lda + k]  (activations [tokens][C], nn.Linear weights [out][in]);
 *   a_mn = 1: A is MN-major, element (m,k) at A[k*lda + m] (the same arrays used transposed); likewise b_mn / ldb.
 *   Epilogue (split_k == 0): + bias[n]; zout[m][n] = pre-activation (optional); act 1 = GELU, act 2 = multiply by
 *   GELU'(aux[m][n]); + resid[(res_rows > 0 ? m % res_rows : m)][n]; replaces nn.Linear / F.gelu / the residual adds.
 *   a_xform / b_xform = 1: the operand is GELU(what lies in memory), applied on chip while the tile is split into its
 *   tf32 parts -- the activated hidden tensor of an MLP (h = GELU(z)) is never written to HBM; D may be NULL when only
 *   the pre-activation `zout` is wanted.
 *   split_k != 0: the K range is split over CTAs (weight gradients: K = tokens), partials in `workspace`
 *   (sb200_gemm_workspace floats) are reduced in a fixed order; no fused epilogue.
 *   Bases and leading dimensions must be 16-byte multiples for the tensor-core path; anything else (and tc mode 0)
 *   runs the exact-fp32 CUDA-core kernel. */
int64_t sb200_gemm_workspace(int M, int N, int K, int b_mn, int split_k, int tc_mode);
int sb200_gemm(const float* A, int64_t lda, int a_mn, const float* B, int64_t ldb, int b_mn, float* D, int64_t ldd,
               int M, int N, int K, const float* bias, int act, const float* aux, int64_t ld_aux,
               const float* resid, int64_t ld_res, int res_rows, float* zout, int64_t ld_z, int a_xform, int b_xform,
               int split_k, float* workspace, void* stream, int tc_mode);

/* A batch of independent GEMMs of identical geometry in ONE launch (block-diagonal layers: the AFNO2D block MLP of
 * src/nsbench/models/fourcastnet/fourcastnet.py:95-121 as nb real-embedded complex GEMMs).  Batch bi shifts the operand
 * coordinates and the output pointers; nothing is gathered or copied:
 *   A: element (m,k) of batch bi at A[m*lda + k + bi*a_off] (a_mn = 0) or A[k*lda + m + bi*a_off] (a_mn = 1);
 *   B: element (n,k) at B[(n + bi*b_off1)*ldb + k + bi*b_off0] (b_mn = 0) or B[(k + bi*b_off1)*ldb + n + bi*b_off0] (b_mn = 1);
 *   D, aux (leading dimension ldd): + bi*d_off;  bias: + bi*bias_off.
 * a_ext0/a_ext1, b_ext0/b_ext1: full extents of the arrays the operands are windows of (dim 0 = contiguous axis).
 * act: 0 none | 3 ReLU | 4 soft-shrink(lam) | 5 multiply by (aux > 0) | 6 multiply by (aux != 0). */
typedef struct sb200_gemm_desc {
    int32_t M, N, K, nbatch;
    int32_t a_mn, b_mn, act, split_k;
    int32_t a_off, b_off0, b_off1, reserved;
    int64_t lda, ldb, ldd;
    int64_t a_ext0, a_ext1, b_ext0, b_ext1;
    int64_t d_off, bias_off;
    float lam; float reserved2;
} sb200_gemm_desc;
int64_t sb200_gemm_batched_workspace(const sb200_gemm_desc* d, int tc_mode);
int sb200_gemm_batched(const sb200_gemm_desc* d, const float* A, const float* B, float* D, const float* bias,
                       const float* aux, float* workspace, void* stream, int tc_mode);

/* AFNO2D block weights w [2,nb,Ni,No] (index 0 real, 1 imaginary) <-> the real embedding E [nb][2*No][2*Ni] of the complex
 * matrices (E[(j,0)][(i,0)] = wr, E[(j,0)][(i,1)] = -wi, E[(j,1)][(i,0)] = wi, E[(j,1)][(i,1)] = wr), and the adjoint map
 * for gradients; out = g * mask(src) with kind 1: src > 0 (ReLU), 2: src != 0 (soft-shrink). */
int sb200_afno_embed(const float* w, float* E, int nb, int Ni, int No, void* stream);
int sb200_afno_unembed(const float* gE, float* gw, int nb, int Ni, int No, void* stream);
int sb200_mask_mul(const float* g, const float* src, float* out, int64_t n, int kind, void* stream);

/* LayerNorm over the last (channel) axis of x [T, C] (biased variance, eps inside the sqrt: torch.nn.LayerNorm as
 * built at fourcastnet.py:236 norm_layer = partial(nn.LayerNorm, eps=1e-6)); mean / rstd [T] are saved for the backward.
 * Backward: dx (+ dres, a gradient that bypasses the norm, added in the same pass), dgamma, dbeta.
 * workspace: sb200_layernorm_bwd_workspace(T, C) floats. */
int sb200_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd,
                        int64_t T, int C, float eps, void* stream);
int64_t sb200_layernorm_bwd_workspace(int64_t T, int C);
int sb200_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* mean, const float* rstd,
                        const float* dres, float* dx, float* dgamma, float* dbeta, float* workspace, int64_t T, int C,
                        void* stream);

/* out[n] = sum_t a[t*lda + n] (bias gradient of a token Linear); workspace: sb200_colsum_workspace(T, N) floats */
int64_t sb200_colsum_workspace(int64_t T, int N);
int sb200_colsum(const float* a, int64_t lda, float* out, int64_t T, int N, float* workspace, void* stream);
/* out[i] = sum_b a[b*n + i] (pos_embed gradient) */
int sb200_batch_sum(const float* a, float* out, int B, int64_t n, void* stream);

/* out[c] = sum_{b,p} g[b,c,p] for g [B,C,HW]: the bias gradient of a SpectralConv that has no skip convolution
 * (with a skip the sum comes out of sb200_pointwise_wgrad).  workspace: sb200_channel_sum_workspace(B,C) floats. */
int64_t sb200_channel_sum_workspace(int B, int C);
int sb200_channel_sum(const float* g, float* out, int B, int C, int64_t HW, float* workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SPECTRAL_B200_H */
