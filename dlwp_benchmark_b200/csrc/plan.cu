// Plans (immutable twiddle tables), error reporting and library identification.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>

#include <vector>

#include "common.cuh"

static thread_local char g_err[512] = "";
thread_local int sb200_tl_mode = 3;

void sb200_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* sb200_last_error(void) { return g_err; }
unsigned long long g_sb200_launches = 0;
extern "C" int64_t sb200_kernel_launches(void) { return (int64_t)__atomic_load_n(&g_sb200_launches, __ATOMIC_RELAXED); }
extern "C" int sb200_version(void) { return 10000 * 0 + 100 * 1 + 0; }

extern "C" int sb200_device_arch(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        cudaGetLastError();
        sb200_set_error("no CUDA device available");
        return -1;
    }
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
        cudaGetLastError();
        sb200_set_error("cudaGetDeviceProperties failed");
        return -1;
    }
    return p.major * 10 + p.minor;
}

// SM count of the current device, cached per device ordinal
int sb200_num_sms() {
    static int cache[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        cudaGetLastError();
        return 148;
    }
    if (dev >= 0 && dev < 64) {
        const int c = __atomic_load_n(&cache[dev], __ATOMIC_RELAXED);
        if (c > 0) return c;
    }
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return 148;
    }
    if (dev >= 0 && dev < 64) __atomic_store_n(&cache[dev], n, __ATOMIC_RELAXED);
    return n;
}

static double hermitian_weight(int kx, int W) {
    if (kx == 0) return 1.0;
    if ((W % 2 == 0) && kx == W / 2) return 1.0;
    return 2.0;
}

static int upload(float2** dst, const std::vector<float2>& src) {
    SB_CHECK_CUDA(cudaMalloc((void**)dst, src.size() * sizeof(float2)));
    SB_CHECK_CUDA(cudaMemcpy(*dst, src.data(), src.size() * sizeof(float2), cudaMemcpyHostToDevice));
    return 0;
}

int sb200_tc_tables_create(sb200_plan_s* p);   // tc_plan.cu
void sb200_tc_tables_destroy(sb200_plan_s* p);

extern "C" int sb200_plan_create(sb200_plan_t* out, int H, int W, int ky0, int My, int Mx,
                                 double scale_fwd, double scale_inv) {
    SB_REQUIRE(out != nullptr, "plan_create: NULL output pointer");
    SB_REQUIRE(H > 0 && W > 0 && My > 0 && Mx > 0, "plan_create: non-positive size");
    SB_REQUIRE(My <= H, "plan_create: My=%d exceeds H=%d", My, H);
    SB_REQUIRE(Mx <= W / 2 + 1, "plan_create: Mx=%d exceeds W/2+1=%d", Mx, W / 2 + 1);
    int dev = 0;
    SB_CHECK_CUDA(cudaGetDevice(&dev));
    sb200_plan_s* p = (sb200_plan_s*)calloc(1, sizeof(sb200_plan_s));
    SB_REQUIRE(p != nullptr, "plan_create: out of host memory");
    p->H = H; p->W = W; p->ky0 = ky0; p->My = My; p->Mx = Mx;
    p->scale_fwd = scale_fwd; p->scale_inv = scale_inv; p->device = dev;

    const double two_pi = 6.283185307179586476925286766559;
    for (int pass = 0; pass < 2; ++pass) {
        std::vector<float2> rowF((size_t)W * Mx), rowI((size_t)Mx * W);
        std::vector<float2> colF((size_t)My * H), colI((size_t)H * My);
        for (int x = 0; x < W; ++x)
            for (int k = 0; k < Mx; ++k) {
                const long r = ((long)k * x) % W;              // exact argument reduction
                const double th = two_pi * (double)r / (double)W;
                const double c = cos(th), s = sin(th);
                const double aF = (pass == 0) ? 1.0 : hermitian_weight(k, W) * scale_inv;
                const double aI = (pass == 0) ? hermitian_weight(k, W) * scale_inv : 1.0;
                rowF[(size_t)x * Mx + k] = make_float2((float)(aF * c), (float)(-aF * s));
                rowI[(size_t)k * W + x] = make_float2((float)(aI * c), (float)(-aI * s));
            }
        for (int j = 0; j < My; ++j) {
            long ky = ((long)ky0 + j) % H;
            if (ky < 0) ky += H;
            for (int y = 0; y < H; ++y) {
                const long r = (ky * y) % H;
                const double th = two_pi * (double)r / (double)H;
                const double c = cos(th), s = sin(th);
                const double sF = (pass == 0) ? scale_fwd : 1.0;
                const double sI = (pass == 0) ? 1.0 : scale_fwd;
                colF[(size_t)j * H + y] = make_float2((float)(sF * c), (float)(-sF * s));
                colI[(size_t)y * My + j] = make_float2((float)(sI * c), (float)(sI * s));
            }
        }
        if (upload(&p->rowF[pass], rowF) || upload(&p->rowI[pass], rowI) ||
            upload(&p->colF[pass], colF) || upload(&p->colI[pass], colI)) {
            sb200_plan_destroy(p);
            return 1;
        }
    }
    if (sb200_tc_tables_create(p)) {
        sb200_plan_destroy(p);
        return 1;
    }
    *out = p;
    return 0;
}

extern "C" int sb200_plan_destroy(sb200_plan_t p) {
    if (!p) return 0;
    sb200_tc_tables_destroy(p);
    for (int pass = 0; pass < 2; ++pass) {
        if (p->rowF[pass]) cudaFree(p->rowF[pass]);
        if (p->rowI[pass]) cudaFree(p->rowI[pass]);
        if (p->colF[pass]) cudaFree(p->colF[pass]);
        if (p->colI[pass]) cudaFree(p->colI[pass]);
    }
    free(p);
    return 0;
}
