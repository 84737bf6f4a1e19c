// Fused truncated 2-D analysis for small grids:  x[img, H, W] (real)  ->  Xh[img, My, Mx] (complex)
//
//   T[y, kx]   = sum_x  x[y, x] * rowF[x, kx]            (real -> complex, along W)
//   Xh[ky, kx] = sum_y  colF[ky, y] * T[y, kx]            (complex, along H)
//
// replaces torch.fft.rfftn + fftshift + slice of neuralop SpectralConv.forward (and, with the pass-1
// tables, the adjoint of irfftn in its backward) without ever writing the row-transformed spectrum T
// to HBM: a persistent CTA streams groups of 256 image rows (G = 256 / H whole images) through a
// two-deep TMA ring, four "row" warps turn a group into T in shared memory and five "column" warps
// finish the previous group, so the only HBM traffic is x once (+ the tiny Xh).
//
// Row stage: one thread owns two image rows and folds the sum twice: the real-input symmetry
//   Re T[k] = sum_{n=0}^{W/2} (x[n] + x[W-n]) cos,   Im T[k] = -sum_{n=1}^{W/2-1} (x[n] - x[W-n]) sin
// and then n <-> W/2 - n, whose twiddles differ by (-1)^k (even and odd frequencies accumulate different folded
// samples), so a row costs W/4 + 1 twiddle records instead of W; twiddles are warp-uniform (broadcast LDS.128), image
// rows are read from the 128B-swizzled TMA tile with conflict-free LDS.128.  The column stage folds y <-> y + H/2 the
// same way (round 2: 34.5 -> 26.9 us at cfg2 shapes; the stages were bound by FFMA2 issue + shared-memory wavefronts).  Exact fp32 FFMA arithmetic (no tensor cores:
// with N = 2*Mx <= 34 output columns a 3xTF32 MMA would be shared-memory-bandwidth bound).
#include "common.cuh"
#include "tc_common.cuh"

constexpr int AF_ROWS = 256;                 // image rows per group
#ifndef AF_RTHREADS_V
#define AF_RTHREADS_V 128      // 2 image rows per row-stage thread: every broadcast twiddle LDS.128 feeds 20 FFMA2 instead of 10
#endif                         // (measured in round 2 at cfg2 shapes: 34.5 us against 37.7 us with 256 threads x 1 row)
constexpr int AF_RTHREADS = AF_RTHREADS_V;   // row-stage threads
constexpr int AF_RPT = AF_ROWS / AF_RTHREADS; // image rows per row-stage thread
#ifndef AF_CTHREADS_V
#define AF_CTHREADS_V 160      // column-stage threads; an item is (image, kx, AF_KYT frequencies): 144 items per group at cfg2 shapes
#endif
#ifndef AF_KYT_V
#define AF_KYT_V 4             // measured at cfg2 shapes (B200, isolated, us): KYT/threads 4/160 26.9, 2/288 26.3, 8/96 28.4;
#endif                         // 256 row threads x 1 row: +2 us.  Before the two folds: 34.5
constexpr int AF_KYT = AF_KYT_V;
constexpr int AF_CTHREADS = AF_CTHREADS_V;   // column-stage threads
constexpr int AF_THREADS = AF_RTHREADS + AF_CTHREADS;

struct AfParams {
    const float2* rowF;      // [W][MX]
    const float2* colF;      // [My][H]
    float2* Xh;              // [nimg][My][MX]
    int64_t nimg;
    int H, W, My, G;
    int ngroups;
    float fold_sign;         // column stage radix-2 fold (H even): T[y] and T[y + H/2] share their twiddles up to (-1)^ky, so the
                             // even frequencies see T[y] + T[y + H/2] and the odd ones T[y] - T[y + H/2] over half the rows.
                             // +1: the first frequency of every group of four is even, -1: odd, 0: no fold (H odd)
    int nxb, ntb;            // depth of the x ring (TMA destinations) and of the T ring (row stage -> column stage)
    long long* trace;        // bring-up builds (SB200_AF_TRACE): clock64 log of CTA 0, [role][group][event]
};
#ifdef SB200_BRINGUP
#define AF_TRACE(role, i, ev) do { if (p.trace && blockIdx.x == 0 && (i) < 16) p.trace[((role) * 16 + (i)) * 4 + (ev)] = clock64(); } while (0)
#else
#define AF_TRACE(role, i, ev) do { } while (0)
#endif

__device__ __forceinline__ float4 af_lds128(const uint8_t* p) { return *reinterpret_cast<const float4*>(p); }

template <int MX>
__global__ void __launch_bounds__(AF_THREADS, 1)
analysis_fused_kernel(const __grid_constant__ CUtensorMap tmapX, const AfParams p) {
    // even / odd frequency split of the row stage (see below): NE even and NO odd frequencies, as fp32x2 pairs
    constexpr int NE = (MX + 1) / 2, NO = MX / 2, PE = (NE + 1) / 2, PO = (NO + 1) / 2, NP = PE + PO;
    constexpr int TWS = 4 * NP;                          // floats per row-twiddle record [Ce | Co | Se | So]
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment by OFFSET from the __shared__ array, so every derived pointer keeps the shared address
    // space (a uintptr_t round-trip turns all later accesses into generic LD/ST through L1TEX)
    uint8_t* base = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    const int H = p.H, W = p.W, My = p.My;
    const int Myp = (My + 3) & ~3;
    const uint32_t half_bytes = AF_ROWS * 128;           // one 32-float column block of a group
    const uint32_t xbuf_bytes = (uint32_t)(W / 32) * half_bytes;
    const int NXB = p.nxb, NTB = p.ntb;
    uint8_t* xbuf = base;                                                         // [NXB][W/32][256][128 B]
    float2* Tbuf = reinterpret_cast<float2*>(xbuf + (uint32_t)NXB * xbuf_bytes);  // [NTB][256][MX]
    float* rtw = reinterpret_cast<float*>(Tbuf + NTB * AF_ROWS * MX);            // [W/2+1][TWS]
    float2* ctw = reinterpret_cast<float2*>(rtw + (W / 2 + 1) * TWS);            // [H][Myp]
    uint64_t* bars = reinterpret_cast<uint64_t*>(ctw + (size_t)H * Myp);
    uint64_t* xfull = bars;          // [NXB <= 4] TMA landed
    uint64_t* tfull = bars + 4;      // [NTB <= 2] T written by the row warps
    uint64_t* tempty = bars + 6;     // [NTB <= 2] T consumed by the column warps

    const int tid = threadIdx.x;
    if (tid == 0) {
        tc::tma_prefetch_desc(&tmapX);
        for (int i = 0; i < 4; ++i) tc::mbar_init(xfull + i, 1);
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(tfull + i, AF_RTHREADS);
            tc::mbar_init(tempty + i, AF_CTHREADS);
        }
        tc::fence_barrier_init();
    }
    const int first = blockIdx.x, stride = gridDim.x;
    const int my_groups = first < p.ngroups ? (p.ngroups - first + stride - 1) / stride : 0;
    const int nhalf = W / 32;
    auto issue = [&](int i, int b) {    // TMA loads of local group i into x buffer b (one thread)
        const int row0 = (first + i * stride) * AF_ROWS;
        tc::mbar_expect_tx(xfull + b, xbuf_bytes);
        for (int h = 0; h < nhalf; ++h)
            tc::tma_load_2d(xbuf + (uint32_t)b * xbuf_bytes + (uint32_t)h * half_bytes, &tmapX, 32 * h, row0, xfull + b);
    };

    // the first groups are requested before the tables are staged: their flight time covers the prologue
    if (tid == 0)
        for (int i = 0; i < NXB && i < my_groups; ++i) issue(i, i);
    // twiddle tables -> shared memory
    for (int idx = tid; idx < (W / 4 + 1) * TWS; idx += AF_THREADS) {
        const int m = idx / TWS, j = idx % TWS;
        // record of folded sample m: cosines of the even frequencies (2*PE floats), of the odd ones (2*PO), then the sines
        int i = j, k = -1;
        bool sine = false;
        if (i < 2 * PE) { if (i < NE) k = 2 * i; }
        else if ((i -= 2 * PE) < 2 * PO) { if (i < NO) k = 2 * i + 1; }
        else if ((i -= 2 * PO) < 2 * PE) { sine = true; if (i < NE) k = 2 * i; }
        else { i -= 2 * PE; sine = true; if (i < NO) k = 2 * i + 1; }
        float v = 0.f;
        if (k >= 0) v = sine ? __ldg(&p.rowF[(size_t)m * MX + k].y) : __ldg(&p.rowF[(size_t)m * MX + k].x);
        rtw[idx] = v;
    }
    for (int idx = tid; idx < H * Myp; idx += AF_THREADS) {
        const int y = idx / Myp, k = idx % Myp;
        ctw[idx] = k < My ? __ldg(p.colF + (size_t)k * H + y) : make_float2(0.f, 0.f);
    }
    __syncthreads();

    if (tid < AF_RTHREADS) {
        // =========================== row stage ===========================
        const int nch = W / 4;
        int b = 0, tb = 0;                  // x / T ring slots of group i
        uint32_t xpar = 0, tpar = 0;        // their phase parities
        for (int i = 0; i < my_groups; ++i) {
            if (tid == 0) AF_TRACE(0, i, 0);
            tc::mbar_wait(xfull + b, xpar);
            if (tid == 0) AF_TRACE(0, i, 1);
            const uint8_t* xb = xbuf + (uint32_t)b * xbuf_bytes;
            // Two folds halve the multiply-adds twice.  Real input: x[n] and x[W-n] share cos and negate sin, so only
            // e[n] = x[n] + x[W-n], o[n] = x[n] - x[W-n], n <= W/2, enter.  Then n and W/2 - n share their twiddles up to
            // (-1)^k: the even frequencies see e[m] + e[W/2-m] (o[m] - o[W/2-m]), the odd ones e[m] - e[W/2-m]
            // (o[m] + o[W/2-m]), m <= W/4 (m = W/4 is its own partner).  Accumulators: [even pairs | odd pairs].
            float2 are[AF_RPT][NP], aim[AF_RPT][NP];
            float cA[AF_RPT], cB[AF_RPT];                           // x[W - 4c], x[W/2 - 4c] carried between chunks
#pragma unroll
            for (int q = 0; q < AF_RPT; ++q) { cA[q] = 0.f; cB[q] = 0.f; }
#pragma unroll
            for (int q = 0; q < AF_RPT; ++q)
#pragma unroll
                for (int k = 0; k < NP; ++k) { are[q][k] = make_float2(0.f, 0.f); aim[q][k] = make_float2(0.f, 0.f); }
            auto mac = [&](const float4* twp, const float (&ee)[AF_RPT], const float (&eo)[AF_RPT], const float (&oe)[AF_RPT],
                           const float (&od)[AF_RPT]) {
                float2 tw[TWS / 2];
#pragma unroll
                for (int v = 0; v < TWS / 4; ++v) {
                    const float4 t4 = twp[v];
                    tw[2 * v] = make_float2(t4.x, t4.y);
                    tw[2 * v + 1] = make_float2(t4.z, t4.w);
                }
#pragma unroll
                for (int q = 0; q < AF_RPT; ++q) {
#pragma unroll
                    for (int k = 0; k < PE; ++k) {
                        are[q][k] = ffma2(make_float2(ee[q], ee[q]), tw[k], are[q][k]);
                        aim[q][k] = ffma2(make_float2(oe[q], oe[q]), tw[NP + k], aim[q][k]);
                    }
#pragma unroll
                    for (int k = 0; k < PO; ++k) {
                        are[q][PE + k] = ffma2(make_float2(eo[q], eo[q]), tw[PE + k], are[q][PE + k]);
                        aim[q][PE + k] = ffma2(make_float2(od[q], od[q]), tw[NP + PE + k], aim[q][PE + k]);
                    }
                }
            };
#pragma unroll 1
            for (int c = 0; c < nch / 4; ++c) {
                float ee[4][AF_RPT], eo[4][AF_RPT], oe[4][AF_RPT], od[4][AF_RPT];
#pragma unroll
                for (int q = 0; q < AF_RPT; ++q) {
                    const int r = tid + q * AF_RTHREADS;
                    const uint8_t* rowp = xb + (uint32_t)r * 128;
                    auto chunk = [&](int ch) {
                        return af_lds128(rowp + (uint32_t)(ch >> 3) * half_bytes + (uint32_t)(((ch & 7) ^ (r & 7)) << 4));
                    };
                    const float4 xa = chunk(c), xz = chunk(nch - 1 - c);                 // x[m..],      x[W-m-4 .. W-m-1]
                    const float4 xc = chunk(nch / 2 - 1 - c), xd = chunk(nch / 2 + c);   // x[W/2-m-4..], x[W/2+m ..]
                    float e1[4], o1[4], e2[4], o2[4];
                    // m = 4c + j pairs with W - m: the carry, then xz.w, xz.z, xz.y
                    e1[0] = c == 0 ? xa.x : xa.x + cA[q];  o1[0] = c == 0 ? 0.f : xa.x - cA[q];
                    e1[1] = xa.y + xz.w; o1[1] = xa.y - xz.w;
                    e1[2] = xa.z + xz.z; o1[2] = xa.z - xz.z;
                    e1[3] = xa.w + xz.y; o1[3] = xa.w - xz.y;
                    // W/2 - m pairs with W/2 + m: the carry (x[W/2] with itself at m = 0), then xc.w, xc.z, xc.y
                    e2[0] = c == 0 ? xd.x : cB[q] + xd.x;  o2[0] = c == 0 ? 0.f : cB[q] - xd.x;
                    e2[1] = xc.w + xd.y; o2[1] = xc.w - xd.y;
                    e2[2] = xc.z + xd.z; o2[2] = xc.z - xd.z;
                    e2[3] = xc.y + xd.w; o2[3] = xc.y - xd.w;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        ee[j][q] = e1[j] + e2[j]; eo[j][q] = e1[j] - e2[j];
                        oe[j][q] = o1[j] - o2[j]; od[j][q] = o1[j] + o2[j];
                    }
                    cA[q] = xz.x; cB[q] = xc.x;
                }
                const float4* twp = reinterpret_cast<const float4*>(rtw + 4 * c * TWS);
#pragma unroll
                for (int j = 0; j < 4; ++j) mac(twp + j * (TWS / 4), ee[j], eo[j], oe[j], od[j]);
            }
            {   // m = W/4 is its own partner: e = x[W/4] + x[3W/4], o = x[W/4] - x[3W/4] go to both parities
                float e16[AF_RPT], o16[AF_RPT];
#pragma unroll
                for (int q = 0; q < AF_RPT; ++q) { e16[q] = cB[q] + cA[q]; o16[q] = cB[q] - cA[q]; }
                mac(reinterpret_cast<const float4*>(rtw + (W / 4) * TWS), e16, e16, o16, o16);
            }
            if (tid == 0) AF_TRACE(0, i, 2);
            tc::mbar_wait(tempty + tb, tpar ^ 1);        // the column warps are done with this T slot
            if (tid == 0) AF_TRACE(0, i, 3);
            float2* Tb = Tbuf + (size_t)tb * AF_ROWS * MX;
#pragma unroll
            for (int q = 0; q < AF_RPT; ++q)
#pragma unroll
                for (int k = 0; k < MX; ++k) {
                    const int pi = (k & 1) ? PE + (k >> 2) : (k >> 2);            // pair of frequency k inside its parity class
                    const float re = ((k >> 1) & 1) ? are[q][pi].y : are[q][pi].x;
                    const float im = ((k >> 1) & 1) ? aim[q][pi].y : aim[q][pi].x;
                    Tb[(tid + q * AF_RTHREADS) * MX + k] = make_float2(re, im);
                }
            tc::mbar_arrive(tfull + tb);
            // every row thread is done reading xbuf[b]: refill it with group i + NXB
            asm volatile("bar.sync 1, %0;" ::"n"(AF_RTHREADS) : "memory");
            if (tid == 0 && i + NXB < my_groups) issue(i + NXB, b);
            if (++b == NXB) { b = 0; xpar ^= 1; }
            if (++tb == NTB) { tb = 0; tpar ^= 1; }
        }
    } else {
        // =========================== column stage ===========================
        const int ct = tid - AF_RTHREADS;
        const int kyq_n = Myp / 4;
        const int items = p.G * kyq_n * MX;
        int tb = 0;
        uint32_t tpar = 0;
        for (int i = 0; i < my_groups; ++i) {
            if (ct == 0) AF_TRACE(1, i, 0);
            tc::mbar_wait(tfull + tb, tpar);
            if (ct == 0) AF_TRACE(1, i, 1);
            const float2* Tb = Tbuf + (size_t)tb * AF_ROWS * MX;
            const int64_t img0 = (int64_t)(first + i * stride) * p.G;
            if (p.fold_sign != 0.f) {
                // folded: item = (image g, kx, AF_KYT consecutive frequencies): T[y] +- T[y + H/2] against their twiddles
                const float2 sg = make_float2(p.fold_sign, p.fold_sign);
                const float2 ng = make_float2(-p.fold_sign, -p.fold_sign);
                const int wstep = Myp >> 1;
                const int kyb_n = (Myp + AF_KYT - 1) / AF_KYT;
                const int items2 = p.G * kyb_n * MX;
                for (int item = ct; item < items2; item += AF_CTHREADS) {
                    const int kx = item % MX;
                    const int rest = item / MX;
                    const int kyb = rest % kyb_n, g = rest / kyb_n;
                    const float2* tp = Tb + (size_t)g * H * MX + kx;
                    const float2* tq = tp + (size_t)(H / 2) * MX;
                    const float4* twp = reinterpret_cast<const float4*>(ctw + kyb * AF_KYT);
                    float2 acc[AF_KYT];
#pragma unroll
                    for (int j = 0; j < AF_KYT; ++j) acc[j] = make_float2(0.f, 0.f);
#pragma unroll 2
                    for (int y = 0; y < H / 2; ++y) {
                        const float2 t1 = *tp, t2 = *tq;
                        const float2 ta = ffma2(t2, sg, t1), tb2 = ffma2(t2, ng, t1);
#pragma unroll
                        for (int j = 0; j < AF_KYT; j += 2) {
                            if (kyb * AF_KYT + j < Myp) {           // (Myp is a multiple of 4: pairs never straddle it)
                                const float4 w01 = twp[j >> 1];
                                cmac2(acc[j], make_float2(w01.x, w01.y), ta);
                                cmac2(acc[j + 1], make_float2(w01.z, w01.w), tb2);
                            }
                        }
                        tp += MX; tq += MX;
                        twp += wstep;
                    }
                    const int64_t img = img0 + g;
                    if (img < p.nimg) {
#pragma unroll
                        for (int j = 0; j < AF_KYT; ++j) {
                            const int ky = kyb * AF_KYT + j;
                            if (ky < My) p.Xh[(img * My + ky) * MX + kx] = acc[j];
                        }
                    }
                }
            } else
            for (int item = ct; item < items; item += AF_CTHREADS) {
                const int kx = item % MX;
                const int rest = item / MX;
                const int kyq = rest % kyq_n, g = rest / kyq_n;
                const float2* Tg = Tb + (size_t)g * H * MX + kx;
                const float4* twp = reinterpret_cast<const float4*>(ctw + kyq * 4);
                float2 acc[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[j] = make_float2(0.f, 0.f);
                const float2* tp = Tg;
                const int wstep = Myp >> 1;
#pragma unroll 4
                for (int y = 0; y < H; ++y) {
                    const float2 tv = *tp;
                    const float4 w01 = twp[0];
                    const float4 w23 = twp[1];
                    cmac2(acc[0], make_float2(w01.x, w01.y), tv);
                    cmac2(acc[1], make_float2(w01.z, w01.w), tv);
                    cmac2(acc[2], make_float2(w23.x, w23.y), tv);
                    cmac2(acc[3], make_float2(w23.z, w23.w), tv);
                    tp += MX;
                    twp += wstep;
                }
                const int64_t img = img0 + g;
                if (img < p.nimg) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int ky = kyq * 4 + j;
                        if (ky < My) p.Xh[(img * My + ky) * MX + kx] = acc[j];
                    }
                }
            }
            tc::mbar_arrive(tempty + tb);
            if (ct == 0) AF_TRACE(1, i, 2);
            if (++tb == NTB) { tb = 0; tpar ^= 1; }
        }
    }
}


static size_t af_smem_bytes(const sb200_plan_s* p, int nxb, int ntb) {
    const int Myp = (p->My + 3) & ~3;
    const int TWS = (2 * p->Mx + 3) / 4 * 4;
    return 1024 + (size_t)nxb * (p->W / 32) * AF_ROWS * 128 + (size_t)ntb * AF_ROWS * p->Mx * 8 +
           (size_t)(p->W / 2 + 1) * TWS * 4 + (size_t)p->H * Myp * 8 + 64;
}

// ring depths (x ring = TMA destinations, T ring = row stage -> column stage).  Measured on B200 at cfg2 shapes
// (session 5, L2-flushed): {2,2} 51.4 us, {3,1} 53.3 us, {2,1} 52.3 us -- the kernel is bound by the latency of its
// FFMA stages, not by bytes in flight, so the second T slot is worth more than a third x buffer.
// SB200_AF_RING=<nxb><ntb> overrides (experiments).
static void af_ring(const sb200_plan_s* p, int* nxb, int* ntb) {
    static const int env = sb_env_int("SB200_AF_RING", 0);
    if (env >= 11 && env / 10 <= 4 && env % 10 >= 1 && env % 10 <= 2 && af_smem_bytes(p, env / 10, env % 10) <= 227 * 1024) {
        *nxb = env / 10; *ntb = env % 10;
        return;
    }
    const int cand[][2] = {{2, 2}, {2, 1}, {1, 1}};
    for (const auto& c : cand)
        if (af_smem_bytes(p, c[0], c[1]) <= 227 * 1024) { *nxb = c[0]; *ntb = c[1]; return; }
    *nxb = 0; *ntb = 0;
}

// geometry check shared by the scratch query and the launcher
static bool af_supported(const sb200_plan_s* p) {
    const int H = p->H, W = p->W, Mx = p->Mx;
    if (W != 32 && W != 64) return false;
    if (H < 8 || AF_ROWS % H != 0) return false;
    if (!(Mx == 5 || Mx == 7 || Mx == 9 || Mx == 13 || Mx == 17)) return false;
    if (Mx > W / 2 + 1) return false;
    int nxb, ntb;
    af_ring(p, &nxb, &ntb);
    return nxb > 0;
}

bool sb200_analysis_fused_supported(sb200_plan_t plan) { return af_supported(plan); }

int sb200_analysis_fused(sb200_plan_t plan, int pass, const float* x, float* Xh, int64_t nimg, cudaStream_t st, int* handled) {
    *handled = 0;
    if (!af_supported(plan) || (reinterpret_cast<uintptr_t>(x) & 15) != 0) return 0;
    const int H = plan->H, W = plan->W, My = plan->My, Mx = plan->Mx;
    const int64_t rows = nimg * H;
    if (rows >= (1LL << 31)) return 0;
    AfParams p;
    p.rowF = plan->rowF[pass]; p.colF = plan->colF[pass]; p.Xh = reinterpret_cast<float2*>(Xh);
    p.nimg = nimg; p.H = H; p.W = W; p.My = My; p.G = AF_ROWS / H;
    p.ngroups = (int)((rows + AF_ROWS - 1) / AF_ROWS);
    af_ring(plan, &p.nxb, &p.ntb);
    p.fold_sign = (H % 2 == 0) ? ((((plan->ky0 % 2) + 2) % 2) ? -1.f : 1.f) : 0.f;
    p.trace = nullptr;
#ifdef SB200_BRINGUP
    static long long* trace_dev = nullptr;
    const int want_trace = sb_env_int("SB200_AF_TRACE", 0);
    if (want_trace) {
        if (!trace_dev) cudaMalloc(&trace_dev, 2 * 16 * 4 * sizeof(long long));
        cudaMemsetAsync(trace_dev, 0, 2 * 16 * 4 * sizeof(long long), st);
        p.trace = trace_dev;
    }
#endif
    const size_t smem = af_smem_bytes(plan, p.nxb, p.ntb);
    CUtensorMap tmap;
    if (int rc = sb200_make_tmap_2d_f32(&tmap, x, (uint64_t)W, (uint64_t)rows, (uint64_t)W * 4, 32, AF_ROWS, 1)) return rc;
    const int nsm = sb200_num_sms();
    const unsigned grid = (unsigned)(p.ngroups < nsm ? p.ngroups : nsm);
#define AF_LAUNCH(MXV)                                                                                                  \
    case MXV:                                                                                                           \
        SB_CHECK_CUDA(cudaFuncSetAttribute(analysis_fused_kernel<MXV>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                           (int)smem));                                                                  \
        sb_launch(analysis_fused_kernel<MXV>, grid, AF_THREADS, smem, st, tmap, p);                                             \
        break;
    switch (Mx) {
        AF_LAUNCH(5) AF_LAUNCH(7) AF_LAUNCH(9) AF_LAUNCH(13) AF_LAUNCH(17)
        default: return 0;
    }
#undef AF_LAUNCH
    SB_LAUNCH_CHECK();
#ifdef SB200_BRINGUP
    if (want_trace) {
        static int calls = 0;
        if (++calls == want_trace) {
            long long h[2 * 16 * 4];
            cudaStreamSynchronize(st);
            cudaMemcpy(h, p.trace, sizeof(h), cudaMemcpyDeviceToHost);
            const long long t0 = h[0];
            printf("analysis_fused trace (CTA 0): row(wait x0, x landed, math done, T slot free) | col(wait T0, T ready, done)\n");
            for (int i = 0; i < 8; ++i) {
                printf("  %2d:", i);
                for (int r = 0; r < 2; ++r) {
                    for (int e = 0; e < (r ? 3 : 4); ++e) printf(" %7lld", h[(r * 16 + i) * 4 + e] ? h[(r * 16 + i) * 4 + e] - t0 : -1);
                    printf("   |");
                }
                printf("\n");
            }
            fflush(stdout);
        }
    }
#endif
    *handled = 1;
    return 0;
}
