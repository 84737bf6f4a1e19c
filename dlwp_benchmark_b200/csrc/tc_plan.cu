// Plan-owned tables for the tensor-core kernels.
//   E[pass]  : [K2pad][128] fp32, the row-synthesis operand for one 128-pixel tile
//              (W >= 128: one image row segment starting at x = 0; W < 128: R = 128/W whole rows,
//               block-diagonal over the rows).  k = r*2*Mx + 2*kx + {0: re, 1: im}.
//   rot      : [V][Mx] (cos, sin)(2 pi kx (128 v) / W): per-mode phase that moves a tile starting at
//              x0 = 128 v back to x = 0 (applied to the Phi operand while it is staged).
#include <math.h>

#include <vector>

#include "common.cuh"

void sb200_tc_tables_destroy(sb200_plan_s* p);

static double herm_w(int kx, int W) {
    if (kx == 0) return 1.0;
    if ((W % 2 == 0) && kx == W / 2) return 1.0;
    return 2.0;
}

int sb200_tc_tables_create(sb200_plan_s* p) {
    p->tc = nullptr;
    const int W = p->W, H = p->H, Mx = p->Mx;
    int R, V;
    if (W >= 128 && W % 128 == 0) { R = 1; V = W / 128; }
    else if (W < 128 && 128 % W == 0) { R = 128 / W; V = 1; }
    else return 0;                                   // tile geometry not supported: CUDA-core kernels only
    if (((int64_t)H * W) % 128 != 0) return 0;
    const int K2 = R * 2 * Mx;
    const int K2pad = (K2 + 7) / 8 * 8;
    if (K2pad > 64) return 0;
    sb200_tc_tables* t = (sb200_tc_tables*)calloc(1, sizeof(sb200_tc_tables));
    if (!t) { sb200_set_error("tc tables: out of host memory"); return 1; }
    t->R = R; t->V = V; t->K2 = K2; t->K2pad = K2pad;
    const double two_pi = 6.283185307179586476925286766559;
    for (int pass = 0; pass < 2; ++pass) {
        std::vector<float> E((size_t)K2pad * 128, 0.f);
        for (int px = 0; px < 128; ++px) {
            const int r = (W < 128) ? px / W : 0;
            const int x = (W < 128) ? px % W : px;
            for (int kx = 0; kx < Mx; ++kx) {
                const long rr = ((long)kx * x) % W;
                const double th = two_pi * (double)rr / (double)W;
                const double a = (pass == 0) ? herm_w(kx, W) * p->scale_inv : 1.0;
                E[(size_t)(r * 2 * Mx + 2 * kx) * 128 + px] = (float)(a * cos(th));
                E[(size_t)(r * 2 * Mx + 2 * kx + 1) * 128 + px] = (float)(-a * sin(th));
            }
        }
        if (cudaMalloc((void**)&t->E[pass], E.size() * sizeof(float)) != cudaSuccess ||
            cudaMemcpy(t->E[pass], E.data(), E.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
            sb200_set_error("tc tables: device allocation failed");
            cudaGetLastError();
            p->tc = t;
            sb200_tc_tables_destroy(p);
            return 1;
        }
    }
    std::vector<float2> rot((size_t)V * Mx);
    for (int v = 0; v < V; ++v)
        for (int kx = 0; kx < Mx; ++kx) {
            const long rr = ((long)kx * 128 * v) % W;
            const double th = two_pi * (double)rr / (double)W;
            rot[(size_t)v * Mx + kx] = make_float2((float)cos(th), (float)sin(th));
        }
    if (cudaMalloc((void**)&t->rot, rot.size() * sizeof(float2)) != cudaSuccess ||
        cudaMemcpy(t->rot, rot.data(), rot.size() * sizeof(float2), cudaMemcpyHostToDevice) != cudaSuccess) {
        sb200_set_error("tc tables: device allocation failed");
        cudaGetLastError();
        p->tc = t;
        sb200_tc_tables_destroy(p);
        return 1;
    }
    p->tc = t;
    return 0;
}

void sb200_tc_tables_destroy(sb200_plan_s* p) {
    sb200_tc_tables* t = (sb200_tc_tables*)p->tc;
    if (!t) return;
    for (int pass = 0; pass < 2; ++pass)
        if (t->E[pass]) cudaFree(t->E[pass]);
    if (t->rot) cudaFree(t->rot);
    free(t);
    p->tc = nullptr;
}
