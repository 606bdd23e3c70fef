// Micro-benchmark + numerical check of the factor-group building blocks (mc_slam_b200/csrc/chol_fg.cuh) against the
// one-warp versions they replace.  Build (no GPU needed) and run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -cudart shared -o tools/ubench_fg tools/ubench_fg.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../mc_slam_b200/csrc/chol_fg.cuh"

using namespace vilba;

__global__ void k_old_factor(const double* A, double* Dt, double* d, long long* cyc) {
    __shared__ double Wsm[16 + 4 * 32];
    const int lane = threadIdx.x;
    double drow[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) drow[c] = (c <= lane) ? A[lane * 32 + c] : 0.0;
    __syncwarp();
    const long long t0 = clock64();
    warp_ldlt_mb4<32>(drow, lane, Wsm);
    const long long t1 = clock64();
#pragma unroll
    for (int c = 0; c < 32; ++c) {
        Dt[c * 32 + lane] = (c < lane) ? drow[c] : 0.0;
        if (c == lane) d[lane] = drow[c];
    }
    if (lane == 0) cyc[0] = t1 - t0;
}

__global__ void k_fg4_factor(const double* A, double* Dt_out, double* d, long long* cyc) {
    __shared__ double Dt[32 * 32];
    __shared__ double scratch[kFgScratch];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double a[2][4];
#pragma unroll
    for (int sl = 0; sl < 2; ++sl)
#pragma unroll
        for (int i = 0; i < 4; ++i) a[sl][i] = A[lane * 32 + 4 * (w + 4 * sl) + i];
    __syncthreads();
    const long long t0 = clock64();
    const FgPivots p = fg4_factor(a, lane, w, Dt, scratch, 2);
    const long long t1 = clock64();
    __syncthreads();
    for (int i = threadIdx.x; i < 1024; i += 128) Dt_out[i] = Dt[i];
    if (w == 0) d[lane] = p.d;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

// 32 rows: old = one thread per row (one warp), new = 2 threads per row (two warps)
__global__ void k_old_rowsolve(const double* R, const double* Dt_in, double* X, long long* cyc) {
    __shared__ double Dt[32 * 32];
    const int lane = threadIdx.x;
    for (int i = lane; i < 1024; i += 32) Dt[i] = Dt_in[i];
    double xr[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) xr[c] = R[lane * 32 + c];
    __syncwarp();
    const long long t0 = clock64();
#pragma unroll
    for (int kk = 0; kk < 31; ++kk)
#pragma unroll
        for (int c = kk + 1; c < 32; ++c) xr[c] -= xr[kk] * Dt[kk * 32 + c];
    const long long t1 = clock64();
#pragma unroll
    for (int c = 0; c < 32; ++c) X[lane * 32 + c] = xr[c];
    if (lane == 0) cyc[0] = t1 - t0;
}

__global__ void k_rowsolve_halves(const double* R, const double* Dt_in, double* X, long long* cyc) {
    __shared__ double Dt[32 * 32];
    __shared__ double scr[32 * 17];
    const int lane = threadIdx.x;
    for (int i = lane; i < 1024; i += blockDim.x) Dt[i] = Dt_in[i];
    double lo[16], hi[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) lo[c] = R[lane * 32 + c], hi[c] = R[lane * 32 + 16 + c];
    __syncwarp();
    const long long t0 = clock64();
    rowsolve_lo(lo, Dt);
#pragma unroll
    for (int c = 0; c < 16; ++c) scr[lane * 17 + c] = lo[c];
    rowsolve_hi(hi, scr + lane * 17, Dt);
    const long long t1 = clock64();
#pragma unroll
    for (int c = 0; c < 16; ++c) X[lane * 32 + c] = lo[c], X[lane * 32 + 16 + c] = hi[c];
    if (lane == 0) cyc[0] = t1 - t0;
}

int main() {
    const int n = 32;
    std::vector<double> A(n * n), L(n * n, 0.0), d(n), R(n * n);
    srand(3);
    std::vector<double> B(n * n);
    for (auto& v : B) v = rand() / (double)RAND_MAX - 0.5;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double s = (i == j) ? 0.5 : 0.0;
            for (int k = 0; k < n; ++k) s += B[i * n + k] * B[j * n + k];
            A[i * n + j] = s;
        }
    for (auto& v : R) v = rand() / (double)RAND_MAX - 0.5;
    // CPU LDL^T
    std::vector<double> W = A;
    for (int j = 0; j < n; ++j) {
        d[j] = W[j * n + j];
        for (int i = j + 1; i < n; ++i) L[i * n + j] = W[i * n + j] / d[j];
        for (int i = j + 1; i < n; ++i)
            for (int c = j + 1; c <= i; ++c) W[i * n + c] -= L[i * n + j] * d[j] * L[c * n + j];
    }
    double *dA, *dDt, *dd, *dR, *dX;
    long long* dc;
    cudaMalloc(&dA, 8 * n * n), cudaMalloc(&dDt, 8 * n * n), cudaMalloc(&dd, 8 * n), cudaMalloc(&dR, 8 * n * n), cudaMalloc(&dX, 8 * n * n);
    cudaMalloc(&dc, 64);
    cudaMemcpy(dA, A.data(), 8 * n * n, cudaMemcpyHostToDevice);
    cudaMemcpy(dR, R.data(), 8 * n * n, cudaMemcpyHostToDevice);
    std::vector<double> hDt(n * n), hd(n), hX(n * n);
    long long cyc = 0;
    auto check_factor = [&](const char* name) {
        cudaMemcpy(hDt.data(), dDt, 8 * n * n, cudaMemcpyDeviceToHost);
        cudaMemcpy(hd.data(), dd, 8 * n, cudaMemcpyDeviceToHost);
        cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost);
        double el = 0, ed = 0;
        for (int k = 0; k < n; ++k) {
            ed = fmax(ed, fabs(hd[k] - d[k]) / fabs(d[k]));
            for (int c = 0; c < n; ++c) el = fmax(el, fabs(hDt[k * n + c] - (c > k ? L[c * n + k] : 0.0)));
        }
        printf("%-28s %8lld cycles   max |dL| %.2e  max rel |dd| %.2e  (%s)\n", name, cyc, el, ed, cudaGetErrorString(cudaGetLastError()));
        if (el > 1e-9) {
            int shown = 0;
            for (int k = 0; k < n && shown < 12; ++k)
                for (int c = k + 1; c < n && shown < 12; ++c)
                    if (fabs(hDt[k * n + c] - L[c * n + k]) > 1e-9) printf("   l(%d,%d) = %g, expected %g\n", c, k, hDt[k * n + c], L[c * n + k]), ++shown;
            for (int k = 0; k < n; ++k) if (fabs(hd[k] - d[k]) > 1e-9 * fabs(d[k])) { printf("   first bad d: d[%d] = %g, expected %g\n", k, hd[k], d[k]); break; }
        }
    };
    for (int rep = 0; rep < 3; ++rep) {
        k_old_factor<<<1, 32>>>(dA, dDt, dd, dc);
        cudaDeviceSynchronize();
        check_factor("warp_ldlt_mb4<32> (1 warp)");
        cudaMemset(dDt, 0, 8 * n * n);
        k_fg4_factor<<<1, 128>>>(dA, dDt, dd, dc);
        cudaDeviceSynchronize();
        check_factor("fg4_factor (4 warps)");
    }
    // row solve reference
    std::vector<double> Xr = R;
    for (int r = 0; r < n; ++r)
        for (int k = 0; k < n; ++k)
            for (int c = k + 1; c < n; ++c) Xr[r * n + c] -= Xr[r * n + k] * L[c * n + k];
    auto check_solve = [&](const char* name) {
        cudaMemcpy(hX.data(), dX, 8 * n * n, cudaMemcpyDeviceToHost);
        cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost);
        double e = 0;
        for (int i = 0; i < n * n; ++i) e = fmax(e, fabs(hX[i] - Xr[i]));
        printf("%-28s %8lld cycles   max |dX| %.2e  (%s)\n", name, cyc, e, cudaGetErrorString(cudaGetLastError()));
    };
    // the solves read the factor the GPU produced (dDt of the last run)
    for (int rep = 0; rep < 3; ++rep) {
        k_old_rowsolve<<<1, 32>>>(dR, dDt, dX, dc);
        cudaDeviceSynchronize();
        check_solve("row solve, thread per row");
        cudaMemset(dX, 0, 8 * n * n);
        k_rowsolve_halves<<<1, 32>>>(dR, dDt, dX, dc);
        cudaDeviceSynchronize();
        check_solve("row solve in two halves");
    }
    return 0;
}
