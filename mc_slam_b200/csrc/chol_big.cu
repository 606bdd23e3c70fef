// chol_big.cu -- reduced camera system of LARGE windows (n > 640, e.g. BASELINE config 4: n = 1485): blocked
// right-looking Cholesky over the whole GPU instead of one thread-block cluster.
//
// Replaces LinearSolverEigen::solve (g2o/solvers/linear_solver_eigen.h:94-124: SimplicialLDLT factor + two
// triangular solves) for the dense SPD matrix S the Schur step leaves in DevWindow::S (lower triangle,
// column-major: A(i,j), i >= j, at S[j * lds + i]).  Per 64-column step k:
//     potrf   : one CTA factors the diagonal tile L_kk                                    (sequential part)
//     trsm    : one THREAD per row of the panel solves x L_kk^T = a in registers; the right-hand side b_s rides as
//               one more row, so the forward substitution y = L^-1 b_s needs no pass of its own
//     update  : one CTA per 64x64 tile of the trailing matrix, C_IJ -= P_I P_J^T (register-tiled FP64 GEMM),
//               the diagonal-tile CTAs also update the right-hand side
// followed by one CTA that back-substitutes L^T x = y.  A non-positive pivot raises LmState::chol_fail (the LM
// controller then rejects the trial, optimization_algorithm_levenberg.cpp:126-127) and is replaced by 1 so
// that everything stays finite.  Every kernel returns at once unless its window is in the TRIAL phase.
#include <algorithm>

#include "lba_common.cuh"

namespace vilba {

constexpr int kBT = 64;        // tile edge
constexpr int kBTP = kBT + 1;  // padded shared-memory stride

__device__ __forceinline__ int big_tiles(int n) { return (n + kBT - 1) / kBT; }

__global__ void __launch_bounds__(256) bigchol_potrf_kernel(const DevWindow* __restrict__ wp, int k) {
    const DevWindow w = wp[blockIdx.y];
    if (w.lm->phase != PH_TRIAL) return;
    const int n = w.n, ld = w.lds;
    if (k >= big_tiles(n)) return;
    const int k0 = k * kBT, nb = min(kBT, n - k0);
    __shared__ double T[kBT][kBTP];
    __shared__ double s_inv;
    __shared__ int s_fail;
    double* A = w.S;
    const int tid = threadIdx.x;
    if (tid == 0) s_fail = 0;
    for (int idx = tid; idx < kBT * kBT; idx += blockDim.x) {
        const int c = idx / kBT, i = idx - kBT * c;  // consecutive threads walk down a column: coalesced
        T[i][c] = (i < nb && c < nb && i >= c) ? A[(size_t)(k0 + c) * ld + k0 + i] : 0.0;
    }
    __syncthreads();
    for (int j = 0; j < nb; ++j) {
        if (tid == 0) {
            double d = T[j][j];
            if (!(d > 0.0)) {
                s_fail = 1;
                d = 1.0;
            }
            d = sqrt(d);
            T[j][j] = d;
            s_inv = 1.0 / d;
        }
        __syncthreads();
        if (tid > j && tid < nb) T[tid][j] *= s_inv;
        __syncthreads();
        // trailing update of the tile: T(i,c) -= L(i,j) L(c,j), j < c <= i
        const int rem = nb - j - 1;
        for (int idx = tid; idx < rem * rem; idx += blockDim.x) {
            const int ci = idx / rem, ii = idx - rem * ci;
            const int c = j + 1 + ci, i = j + 1 + ii;
            if (i >= c) T[i][c] -= T[i][j] * T[c][j];
        }
        __syncthreads();
    }
    for (int idx = tid; idx < kBT * kBT; idx += blockDim.x) {
        const int c = idx / kBT, i = idx - kBT * c;
        if (i < nb && c < nb && i >= c) A[(size_t)(k0 + c) * ld + k0 + i] = T[i][c];
    }
    if (tid == 0 && s_fail) w.lm->chol_fail = 1;
}

// rows [k0 + 64, n) of the panel and the rhs row (index n): x L_kk^T = a, one thread per row
__global__ void __launch_bounds__(128) bigchol_trsm_kernel(const DevWindow* __restrict__ wp, int k) {
    const DevWindow w = wp[blockIdx.y];
    if (w.lm->phase != PH_TRIAL) return;
    const int n = w.n, ld = w.lds;
    if (k >= big_tiles(n)) return;
    const int k0 = k * kBT, nb = min(kBT, n - k0);
    const int row0 = k0 + kBT;  // first panel row (may be >= n: then only the rhs row is left)
    const int rows = max(0, n - row0) + 1;  // + rhs
    if ((int)(blockIdx.x * blockDim.x) >= rows) return;
    __shared__ double L[kBT][kBTP];  // L_kk padded with the identity
    __shared__ double dinv[kBT];
    double* A = w.S;
    const int tid = threadIdx.x;
    for (int idx = tid; idx < kBT * kBT; idx += blockDim.x) {
        const int c = idx / kBT, i = idx - kBT * c;
        double v = (i == c) ? 1.0 : 0.0;
        if (i < nb && c < nb && i >= c) v = A[(size_t)(k0 + c) * ld + k0 + i];
        L[i][c] = v;
    }
    __syncthreads();
    if (tid < kBT) dinv[tid] = 1.0 / L[tid][tid];
    __syncthreads();
    const int r = blockIdx.x * blockDim.x + tid;
    if (r >= rows) return;
    const bool is_rhs = (r == rows - 1);
    const int gi = row0 + r;  // global row (unused for the rhs)
    double x[kBT];
#pragma unroll
    for (int c = 0; c < kBT; ++c) {
        double v = 0.0;
        if (c < nb) v = is_rhs ? w.bs[k0 + c] : A[(size_t)(k0 + c) * ld + gi];
        x[c] = v;
    }
#pragma unroll
    for (int c = 0; c < kBT; ++c) {
        double s0 = x[c], s1 = 0.0;
#pragma unroll
        for (int m = 0; m + 1 < c; m += 2) {
            s0 = fma(-x[m], L[c][m], s0);
            s1 = fma(-x[m + 1], L[c][m + 1], s1);
        }
        if (c & 1) s0 = fma(-x[c - 1], L[c][c - 1], s0);
        x[c] = (s0 + s1) * dinv[c];
    }
    if (is_rhs) {
#pragma unroll
        for (int c = 0; c < kBT; ++c)
            if (c < nb) w.x[k0 + c] = x[c];  // y_k = forward-substituted right-hand side
    } else {
#pragma unroll
        for (int c = 0; c < kBT; ++c)
            if (c < nb) A[(size_t)(k0 + c) * ld + gi] = x[c];
    }
}

// trailing tiles (I >= J > k): C_IJ -= P_I P_J^T ; diagonal tiles also b_J -= P_J y_k
__global__ void __launch_bounds__(256) bigchol_update_kernel(const DevWindow* __restrict__ wp, int k) {
    const DevWindow w = wp[blockIdx.y];
    if (w.lm->phase != PH_TRIAL) return;
    const int n = w.n, ld = w.lds;
    const int ntile = big_tiles(n);
    const int T = ntile - k - 1;  // tiles left below / right of tile k
    if (T <= 0) return;
    const int npair = T * (T + 1) / 2;
    extern __shared__ double upd_sm[];
    double (*PI)[kBTP] = reinterpret_cast<double (*)[kBTP]>(upd_sm);
    double (*PJ)[kBTP] = reinterpret_cast<double (*)[kBTP]>(upd_sm + kBT * kBTP);
    double* A = w.S;
    const int k0 = k * kBT;  // tile k is complete here (T > 0): 64 columns
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    for (int pr = blockIdx.x; pr < npair; pr += gridDim.x) {
        // pr = I' (I' + 1) / 2 + J', 0 <= J' <= I' < T
        int Ip = (int)((sqrtf(8.0f * (float)pr + 1.0f) - 1.0f) * 0.5f);
        while (Ip * (Ip + 1) / 2 > pr) --Ip;
        while ((Ip + 1) * (Ip + 2) / 2 <= pr) ++Ip;
        const int Jp = pr - Ip * (Ip + 1) / 2;
        const int I0 = (k + 1 + Ip) * kBT, J0 = (k + 1 + Jp) * kBT;
        const int ni = min(kBT, n - I0), nj = min(kBT, n - J0);
        __syncthreads();
        for (int idx = tid; idx < kBT * kBT; idx += blockDim.x) {
            const int c = idx / kBT, i = idx - kBT * c;
            PI[i][c] = (i < ni) ? A[(size_t)(k0 + c) * ld + I0 + i] : 0.0;
            PJ[i][c] = (i < nj) ? A[(size_t)(k0 + c) * ld + J0 + i] : 0.0;
        }
        __syncthreads();
        double acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
#pragma unroll 8
        for (int c = 0; c < kBT; ++c) {
            double ra[4], rb[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) ra[a] = PI[tx + 16 * a][c];  // lanes along the rows of C: coalesced stores
#pragma unroll
            for (int b = 0; b < 4; ++b) rb[b] = PJ[ty + 16 * b][c];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = fma(ra[a], rb[b], acc[a][b]);
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int j = ty + 16 * b;
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const int i = tx + 16 * a;
                if (i < ni && j < nj && I0 + i >= J0 + j) A[(size_t)(J0 + j) * ld + I0 + i] -= acc[a][b];
            }
        }
        if (Ip == Jp && tid < nj) {  // rhs rows of this diagonal tile
            double s = 0.0;
#pragma unroll 8
            for (int c = 0; c < kBT; ++c) s = fma(PJ[tid][c], w.x[k0 + c], s);
            w.bs[J0 + tid] -= s;
        }
    }
}

// L^T x = y, one CTA
__global__ void __launch_bounds__(1024) bigchol_backsub_kernel(const DevWindow* __restrict__ wp) {
    const DevWindow w = wp[blockIdx.y];
    if (w.lm->phase != PH_TRIAL) return;
    const int n = w.n, ld = w.lds;
    extern __shared__ double sm[];
    double* xs = sm;                     // n   solution (the part below the current tile is final)
    double* Lt = xs + ((n + 1) & ~1);    // 64 x 65 diagonal tile
    double* sv = Lt + kBT * kBTP;        // 64
    const double* A = w.S;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    const int ntile = big_tiles(n);
    for (int k = ntile - 1; k >= 0; --k) {
        const int k0 = k * kBT, nb = min(kBT, n - k0), below = k0 + kBT;
        for (int idx = tid; idx < kBT * kBT; idx += blockDim.x) {
            const int c = idx / kBT, i = idx - kBT * c;
            Lt[i * kBTP + c] = (i < nb && c < nb && i >= c) ? A[(size_t)(k0 + c) * ld + k0 + i] : (i == c ? 1.0 : 0.0);
        }
        // s_c = y_c - sum_{i >= below} L(i, k0 + c) x_i : one warp per column, lanes along the (contiguous) rows
        for (int c = warp; c < nb; c += nwarp) {
            const double* col = A + (size_t)(k0 + c) * ld;
            double s = 0.0;
            for (int i = below + lane; i < n; i += 32) s = fma(col[i], xs[i], s);
            s = warp_sum(s);
            if (lane == 0) sv[c] = w.x[k0 + c] - s;
        }
        __syncthreads();
        if (warp == 0) {  // 64 x 64 triangular solve with L_kk^T: lanes own columns c and c + 32
            double s0 = lane < nb ? sv[lane] : 0.0, s1 = lane + 32 < nb ? sv[lane + 32] : 0.0;
            for (int m = nb - 1; m >= 0; --m) {
                const double sm_ = __shfl_sync(0xffffffffu, m < 32 ? s0 : s1, m & 31);
                const double xm = sm_ / Lt[m * kBTP + m];
                if (lane == (m & 31)) {
                    if (m < 32) s0 = xm; else s1 = xm;
                }
                if (lane < m) s0 = fma(-Lt[m * kBTP + lane], xm, s0);
                if (lane + 32 < m) s1 = fma(-Lt[m * kBTP + lane + 32], xm, s1);
            }
            if (lane < nb) xs[k0 + lane] = s0;
            if (lane + 32 < nb) xs[k0 + lane + 32] = s1;
        }
        __syncthreads();
    }
    for (int i = tid; i < n; i += blockDim.x) w.x[i] = xs[i];
}

size_t chol_big_backsub_smem(int n_cap) { return sizeof(double) * ((size_t)((n_cap + 1) & ~1) + kBT * kBTP + kBT); }

constexpr size_t kUpdateSmem = sizeof(double) * 2 * kBT * kBTP;

cudaError_t configure_chol_big(int n_cap) {
    cudaError_t e = cudaFuncSetAttribute(bigchol_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kUpdateSmem);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(bigchol_backsub_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chol_big_backsub_smem(n_cap));
}

cudaError_t launch_chol_big(cudaStream_t s, const DevWindow* wp, const LaunchDims& d) {
    const int ntile = d.chol_big_tiles;
    for (int k = 0; k < ntile; ++k) {
        bigchol_potrf_kernel<<<dim3(1, d.n_windows), 256, 0, s>>>(wp, k);
        const int rows = (ntile - k - 1) * kBT + 1;
        bigchol_trsm_kernel<<<dim3((rows + 127) / 128, d.n_windows), 128, 0, s>>>(wp, k);
        const int T = ntile - k - 1;
        if (T > 0) {
            const int npair = T * (T + 1) / 2;
            bigchol_update_kernel<<<dim3(std::min(npair, 4 * d.sm_count), d.n_windows), 256, kUpdateSmem, s>>>(wp, k);
        }
    }
    bigchol_backsub_kernel<<<dim3(1, d.n_windows), 1024, chol_big_backsub_smem(d.chol_big_tiles * kBT), s>>>(wp);
    return cudaGetLastError();
}

}  // namespace vilba
