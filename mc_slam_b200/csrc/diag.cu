// diag.cu -- diagnostic entry points (include/vilba_diag.h): single kernels of the path behind a plain C call, so
// that tests / tools can check them against numpy and time them alone.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/vilba_diag.h"
#include "kernels.h"

namespace vilba {
}  // namespace vilba

using namespace vilba;

extern "C" int vilba_diag_dense_solve(int32_t device, int32_t n, const double* S, const double* b, int32_t variant,
                                      int32_t cluster, int32_t n_windows, int32_t reps, double* x_out, int32_t* fail_out,
                                      double* avg_us) {
    if (n < 1 || !S || !b || n_windows < 1 || n_windows > kMaxBatch || reps < 1 || cluster < 1) return VILBA_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return VILBA_ERR_NO_DEVICE;
    const int ld = (n + 3) & ~3;
    const size_t mat = (size_t)ld * n;
    const size_t scr = std::max<size_t>(chol_la_scratch_doubles(n), chol_big_scratch_doubles(n));
    // per window: S | bs | x | Lfac | cminv | cdinv, then pristine S | b shared by all windows
    const size_t per_win = 2 * mat + 3 * (size_t)ld + scr + 64;  // the last 64: debug counters
    double* dev = nullptr;
    LmState* lm = nullptr;
    DevWindow* dwp = nullptr;
    cudaStream_t s = nullptr, side = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr, ef = nullptr, ej = nullptr;
    int status = VILBA_ERR_CUDA;
    std::vector<double> hS(mat, 0.0);
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) hS[(size_t)j * ld + i] = S[(size_t)j * n + i];
    std::vector<DevWindow> dw(n_windows);
    std::vector<LmState> hlm(n_windows);
    float ms_total = 0.f;
    do {
        if (cudaMalloc(&dev, sizeof(double) * (per_win * n_windows + mat + ld)) != cudaSuccess) break;
        if (cudaMalloc(&lm, sizeof(LmState) * n_windows) != cudaSuccess) break;
        if (cudaMalloc(&dwp, sizeof(DevWindow) * n_windows) != cudaSuccess) break;
        if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) break;
        if (cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking) != cudaSuccess) break;
        if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) break;
        if (cudaEventCreateWithFlags(&ef, cudaEventDisableTiming) != cudaSuccess) break;
        if (cudaEventCreateWithFlags(&ej, cudaEventDisableTiming) != cudaSuccess) break;
        double* S0 = dev + per_win * n_windows;
        double* b0 = S0 + mat;
        if (cudaMemcpy(S0, hS.data(), sizeof(double) * mat, cudaMemcpyHostToDevice) != cudaSuccess) break;
        if (cudaMemcpy(b0, b, sizeof(double) * n, cudaMemcpyHostToDevice) != cudaSuccess) break;
        std::memset(hlm.data(), 0, sizeof(LmState) * n_windows);
        for (int i = 0; i < n_windows; ++i) {
            DevWindow& w = dw[i];
            std::memset(&w, 0, sizeof(w));
            double* p = dev + per_win * i;
            w.n = n, w.lds = ld, w.n_free = n / 15;
            w.S = p, p += mat;
            w.bs = p, p += ld;
            w.x = p, p += ld;
            w.Lfac = p, p += mat;
            w.cminv = p, p += scr;
            w.cdinv = p, p += ld;
            w.dbg = reinterpret_cast<long long*>(p);
            w.S_w = w.S, w.bs_w = w.bs;
            w.lm = lm + i;
            hlm[i].phase = PH_TRIAL;
        }
        if (cudaMemset(dev, 0, sizeof(double) * (per_win * n_windows)) != cudaSuccess) break;
        if (cudaMemcpy(lm, hlm.data(), sizeof(LmState) * n_windows, cudaMemcpyHostToDevice) != cudaSuccess) break;
        if (cudaMemcpy(dwp, dw.data(), sizeof(DevWindow) * n_windows, cudaMemcpyHostToDevice) != cudaSuccess) break;
        LaunchDims d;
        std::memset(&d, 0, sizeof(d));
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) break;
        d.sm_count = prop.multiProcessorCount;
        d.n_windows = n_windows;
        d.chol_cluster = cluster;
        cudaError_t e = cudaSuccess;
        if (variant == 1) {
            if (!chol_la_fits(n, cluster)) { status = VILBA_ERR_ARG; break; }
            e = configure_chol_la();
        } else if (variant == 2) {
            d.chol_big_tiles = (n + 63) / 64;
            e = configure_chol_big(n);
        } else {
            status = VILBA_ERR_ARG;
            break;
        }
        if (e != cudaSuccess) break;
        bool ok = true;
        for (int r = 0; r < reps && ok; ++r) {
            for (int i = 0; i < n_windows && ok; ++i) {
                ok = cudaMemcpyAsync(dw[i].S, S0, sizeof(double) * mat, cudaMemcpyDeviceToDevice, s) == cudaSuccess &&
                     cudaMemcpyAsync(dw[i].bs, b0, sizeof(double) * n, cudaMemcpyDeviceToDevice, s) == cudaSuccess;
            }
            if (!ok) break;
            cudaEventRecord(e0, s);
            if (variant == 1) e = launch_chol_la(s, dwp, n_windows, cluster, n);
            else e = launch_chol_big(s, side, ef, ej, dwp, d);
            cudaEventRecord(e1, s);
            if (e != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess) { ok = false; break; }
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            ms_total += ms;
        }
        if (!ok) {
            std::fprintf(stderr, "vilba_diag_dense_solve: %s\n", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        if (x_out && cudaMemcpy(x_out, dw[0].x, sizeof(double) * n, cudaMemcpyDeviceToHost) != cudaSuccess) break;
        if (cudaMemcpy(hlm.data(), lm, sizeof(LmState) * n_windows, cudaMemcpyDeviceToHost) != cudaSuccess) break;
        if (fail_out) *fail_out = hlm[0].chol_fail;
        if (std::getenv("VILBA_DEBUG_COUNTERS")) {  // phase timers of a -DVILBA_LA_TIMING build (cycles of the last launch)
            long long h[40];
            if (cudaMemcpy(h, dw[0].dbg, sizeof(h), cudaMemcpyDeviceToHost) == cudaSuccess) {
                static const char* who[3] = {"fg warp 0", "fg warp 1", "worker 0"};
                for (int g = 0; g < 3; ++g) {
                    std::fprintf(stderr, "[vilba dbg] n=%d cluster=%d %s:", n, cluster, who[g]);
                    for (int i = 0; i < 12; ++i) std::fprintf(stderr, " %lld", h[12 * g + i]);
                    std::fprintf(stderr, "\n");
                }

            }
        }
        if (avg_us) *avg_us = 1e3 * ms_total / reps;
        status = VILBA_OK;
    } while (false);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (ef) cudaEventDestroy(ef);
    if (ej) cudaEventDestroy(ej);
    if (s) cudaStreamDestroy(s);
    if (side) cudaStreamDestroy(side);
    if (dwp) cudaFree(dwp);
    if (lm) cudaFree(lm);
    if (dev) cudaFree(dev);
    return status;
}

extern "C" int vilba_diag_dense_supported(int32_t n, int32_t variant, int32_t cluster) {
    if (n < 1 || cluster < 1) return 0;
    if (variant == 1) return chol_la_fits(n, cluster) && cluster <= 16;
    return variant == 2;
}
