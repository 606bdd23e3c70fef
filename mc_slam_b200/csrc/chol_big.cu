// chol_big.cu -- reduced camera system of LARGE windows (n > 480, e.g. BASELINE config 4: n = 1485): blocked
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
#include <cstdlib>

#include "lba_common.cuh"

namespace vilba {

constexpr int kBT = 64;        // tile edge
constexpr int kBTP = kBT + 1;  // padded shared-memory stride
constexpr int kXS = kBT + 8;   // stride of the row-per-column buffer of trsm (2-way conflicts at most)

__device__ __forceinline__ int big_tiles(int n) { return (n + kBT - 1) / kBT; }

// The kernels keep their loops rolled on purpose: fully unrolled register-resident variants of potrf / trsm were
// measured 3x slower -- 4 k instructions of straight-line code executed once by two warps are bound by
// instruction fetch.

// diagonal tile: 256 threads, 4 per row; the pivot's reciprocal square root is computed redundantly by every
// thread (no publication step): two barriers per column
template <int TPR>  // threads per row of the tile
__global__ void __launch_bounds__(64 * TPR) bigchol_potrf_kernel(const DevWindow* __restrict__ wp, int k) {
    const DevWindow w = wp[blockIdx.y];
    if (w.lm->phase != PH_TRIAL) return;
    const int n = w.n, ld = w.lds;
    if (k >= big_tiles(n)) return;
    const int k0 = k * kBT, nb = min(kBT, n - k0);
    __shared__ double T[kBT][kBTP];
    double* A = w.S;
    // a warp = 32 consecutive rows with the same column phase q: T[row][c] is conflict-free, T[c][j] a broadcast
    const int tid = threadIdx.x, row = (tid & 31) + 32 * ((tid >> 5) & 1), q = tid >> 6;
#pragma unroll 4
    for (int idx = tid; idx < kBT * kBT; idx += 64 * TPR) {
        const int c = idx / kBT, i = idx - kBT * c;  // consecutive threads walk down a column: coalesced
        T[i][c] = (i < nb && c < nb && i >= c) ? A[(size_t)(k0 + c) * ld + k0 + i] : 0.0;
    }
    __shared__ double s_rs[2];  // 1 / sqrt(pivot) of the current column, published one column ahead
    __shared__ int s_bad;
    if (tid == 0) {
        s_bad = 0;
        double d0 = nb > 0 ? A[(size_t)k0 * ld + k0] : 1.0;
        if (!(d0 > 0.0)) {
            s_bad = 1;
            d0 = 1.0;
            T[0][0] = 1.0;  // (this thread loaded T[0][0] itself)
        }
        s_rs[0] = rsqrt(d0);
    }
    double l_prev = 0.0;  // L(row, j - 1): stored one column late, when nobody reads the old column any more
    for (int j = 0; j < nb; ++j) {
        __syncthreads();  // the updates of column j - 1 are complete, s_rs[j & 1] is published
        if (j > 0 && row >= j - 1 && q == 0) T[row][j - 1] = l_prev;
        const double inv = s_rs[j & 1];
        // L(j,j) = sqrt(d) = d / sqrt(d), L(i,j) = a(i,j) / sqrt(d); L(c,j) of the other rows is recomputed from a(c,j)
        const double lij = T[row][j] * inv;
        if (row > j) {
            // the owner of the next pivot finishes it first and publishes its reciprocal square root, so that the
            // rsqrt chain runs beside the other threads' updates instead of in front of everybody's next column
            if (row == j + 1 && q == 0 && j + 1 < nb) {
                double dn = fma(-lij, T[j + 1][j] * inv, T[j + 1][j + 1]);
                T[j + 1][j + 1] = dn;
                if (!(dn > 0.0)) {
                    s_bad = 1;
                    dn = 1.0;
                    T[j + 1][j + 1] = 1.0;
                }
                s_rs[(j + 1) & 1] = rsqrt(dn);
            } else {
                for (int c0 = j + 1 + q; c0 <= row; c0 += 4 * TPR) {  // chunks of 4: loads first, then stores
                    double tv[4], cv[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int c = c0 + TPR * i;
                        tv[i] = c <= row ? T[row][c] : 0.0;
                        cv[i] = c <= row ? T[c][j] : 0.0;
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int c = c0 + TPR * i;
                        if (c <= row) T[row][c] = fma(-lij, cv[i] * inv, tv[i]);
                    }
                }
            }
        }
        l_prev = lij;
    }
    __syncthreads();
    if (nb > 0 && row >= nb - 1 && row < kBT && q == 0) T[row][nb - 1] = l_prev;
    __syncthreads();
#pragma unroll 4
    for (int idx = tid; idx < kBT * kBT; idx += 64 * TPR) {
        const int c = idx / kBT, i = idx - kBT * c;
        if (i < nb && c < nb && i >= c) A[(size_t)(k0 + c) * ld + k0 + i] = T[i][c];
    }
    if (tid < nb) w.cdinv[k0 + tid] = 1.0 / T[tid][tid];  // for the triangular solves
    if (tid == 0 && s_bad) w.lm->chol_fail = 1;
}

// rows [k0 + 64, n) of the panel and the rhs row (index n): x L_kk^T = a.  64 rows per CTA, 4 threads per row
// (one warp = 8 rows), the rows live in shared memory; right-looking: once x_c is final the rest of the row is
// updated, split over the 4 threads of the row, ordered by __syncwarp
__global__ void __launch_bounds__(256) bigchol_trsm_kernel(const DevWindow* __restrict__ wp, int k) {
    const DevWindow w = wp[blockIdx.y];
    if (w.lm->phase != PH_TRIAL) return;
    const int n = w.n, ld = w.lds;
    if (k >= big_tiles(n)) return;
    const int k0 = k * kBT, nb = min(kBT, n - k0);
    const int row0 = k0 + kBT;              // first panel row (may be >= n: then only the rhs row is left)
    const int rows = max(0, n - row0) + 1;  // + rhs
    const int r_base = blockIdx.x * kBT;
    if (r_base >= rows) return;
    extern __shared__ double trsm_sm[];
    double (*L)[kBTP] = reinterpret_cast<double (*)[kBTP]>(trsm_sm);              // L_kk padded with the identity
    double (*X)[kXS] = reinterpret_cast<double (*)[kXS]>(trsm_sm + kBT * kBTP);  // X[c][row]
    double* dinv = trsm_sm + kBT * kBTP + kBT * kXS;
    double* A = w.S;
    const int tid = threadIdx.x;
#pragma unroll 4
    for (int idx = tid; idx < kBT * kBT; idx += 256) {
        const int c = idx / kBT, i = idx - kBT * c;
        double v = (i == c) ? 1.0 : 0.0;
        if (i < nb && c < nb && i >= c) v = A[(size_t)(k0 + c) * ld + k0 + i];
        L[i][c] = v;
        const int r = r_base + i;  // row i of this CTA, column c
        double xv = 0.0;
        if (r < rows && c < nb) xv = (r == rows - 1) ? w.bs[k0 + c] : A[(size_t)(k0 + c) * ld + row0 + r];
        X[c][i] = xv;
    }
    if (tid < kBT) dinv[tid] = tid < nb ? w.cdinv[k0 + tid] : 1.0;
    __syncthreads();
    // one warp = 8 rows x 4 column phases, phase-major: the 8 lanes of a phase read 64 contiguous bytes
    const int lane = tid & 31, rl = (tid >> 5) * 8 + (lane & 7), q = lane >> 3;
    for (int c = 0; c < nb; ++c) {
        const double xc = X[c][rl] * dinv[c];
        __syncwarp();
        if (q == 0) X[c][rl] = xc;
        // chunks of 8 entries: all loads, then all stores
        for (int m0 = c + 1 + q; m0 < nb; m0 += 32) {
            double xv[8], lv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int m = m0 + 4 * i;
                xv[i] = m < nb ? X[m][rl] : 0.0;
                lv[i] = m < nb ? L[m][c] : 0.0;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int m = m0 + 4 * i;
                if (m < nb) X[m][rl] = fma(-xc, lv[i], xv[i]);
            }
        }
        __syncwarp();
    }
    __syncthreads();
#pragma unroll 4
    for (int idx = tid; idx < kBT * kBT; idx += 256) {
        const int c = idx / kBT, i = idx - kBT * c;
        const int r = r_base + i;
        if (r < rows && c < nb) {
            if (r == rows - 1)
                w.x[k0 + c] = X[c][i];  // y_k = forward-substituted right-hand side
            else
                A[(size_t)(k0 + c) * ld + row0 + r] = X[c][i];
        }
    }
}

// trailing tiles (I >= J > k): C_IJ -= P_I P_J^T ; diagonal tiles also b_J -= P_J y_k
// part 0: every tile; part 1: only the tiles of block column k + 1 (what the next potrf / trsm need: "look-ahead");
// part 2: the other tiles (they run on a side stream beside the next step's potrf and trsm)
__global__ void __launch_bounds__(256) bigchol_update_kernel(const DevWindow* __restrict__ wp, int k, int part) {
    const DevWindow w = wp[blockIdx.y];
    if (w.lm->phase != PH_TRIAL) return;
    const int n = w.n, ld = w.lds;
    const int ntile = big_tiles(n);
    const int T = ntile - k - 1;  // tiles left below / right of tile k
    if (T <= 0) return;
    const int npair = part == 0 ? T * (T + 1) / 2 : (part == 1 ? T : (T - 1) * T / 2);
    extern __shared__ double upd_sm[];
    double (*PI)[kBTP] = reinterpret_cast<double (*)[kBTP]>(upd_sm);
    double (*PJ)[kBTP] = reinterpret_cast<double (*)[kBTP]>(upd_sm + kBT * kBTP);
    double* A = w.S;
    const int k0 = k * kBT;  // tile k is complete here (T > 0): 64 columns
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    for (int pr = blockIdx.x; pr < npair; pr += gridDim.x) {
        // pr = I' (I' + 1) / 2 + J', 0 <= J' <= I' < T   (part 1: J' = 0; part 2: the triangle without its first column)
        int Ip = (int)((sqrtf(8.0f * (float)pr + 1.0f) - 1.0f) * 0.5f);
        while (Ip * (Ip + 1) / 2 > pr) --Ip;
        while ((Ip + 1) * (Ip + 2) / 2 <= pr) ++Ip;
        int Jp = pr - Ip * (Ip + 1) / 2;
        if (part == 1) Ip = pr, Jp = 0;
        if (part == 2) Ip += 1, Jp += 1;
        const int I0 = (k + 1 + Ip) * kBT, J0 = (k + 1 + Jp) * kBT;
        const int ni = min(kBT, n - I0), nj = min(kBT, n - J0);
        __syncthreads();
#pragma unroll 8
        for (int idx = tid; idx < kBT * kBT; idx += 256) {
            const int c = idx / kBT, i = idx - kBT * c;
            PI[i][c] = (i < ni) ? A[(size_t)(k0 + c) * ld + I0 + i] : 0.0;
            PJ[i][c] = (i < nj) ? A[(size_t)(k0 + c) * ld + J0 + i] : 0.0;
        }
        // the tile of C, fetched while the products are formed
        double cv[4][4];
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const int i = tx + 16 * a, j = ty + 16 * b;
                cv[a][b] = (i < ni && j < nj && I0 + i >= J0 + j) ? A[(size_t)(J0 + j) * ld + I0 + i] : 0.0;
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
                if (i < ni && j < nj && I0 + i >= J0 + j) A[(size_t)(J0 + j) * ld + I0 + i] = cv[a][b] - acc[a][b];
            }
        }
        if (Ip == Jp && tid < nj) {  // rhs rows of this diagonal tile
            double s0 = 0.0, s1 = 0.0;
#pragma unroll 8
            for (int c = 0; c < kBT; c += 2) {
                s0 = fma(PJ[tid][c], w.x[k0 + c], s0);
                s1 = fma(PJ[tid][c + 1], w.x[k0 + c + 1], s1);
            }
            w.bs[J0 + tid] -= s0 + s1;
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
#pragma unroll 4
        for (int idx = tid; idx < kBT * kBT; idx += 1024) {
            const int c = idx / kBT, i = idx - kBT * c;
            Lt[i * kBTP + c] = (i < nb && c < nb && i >= c) ? A[(size_t)(k0 + c) * ld + k0 + i] : (i == c ? 1.0 : 0.0);
        }
        // s_c = y_c - sum_{i >= below} L(i, k0 + c) x_i : one warp per column, lanes along the (contiguous) rows
        for (int c = warp; c < nb; c += nwarp) {
            const double* col = A + (size_t)(k0 + c) * ld;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            int i = below + lane;
            for (; i + 96 < n; i += 128) {
                const double a0 = col[i], a1 = col[i + 32], a2 = col[i + 64], a3 = col[i + 96];
                s0 = fma(a0, xs[i], s0), s1 = fma(a1, xs[i + 32], s1), s2 = fma(a2, xs[i + 64], s2), s3 = fma(a3, xs[i + 96], s3);
            }
            for (; i < n; i += 32) s0 = fma(col[i], xs[i], s0);
            const double s = warp_sum((s0 + s1) + (s2 + s3));
            if (lane == 0) sv[c] = w.x[k0 + c] - s;
        }
        __syncthreads();
        if (warp == 0) {  // 64 x 64 triangular solve with L_kk^T: lanes own columns c and c + 32
            double s0 = lane < nb ? sv[lane] : 0.0, s1 = lane + 32 < nb ? sv[lane + 32] : 0.0;
            const double di0 = lane < nb ? w.cdinv[k0 + lane] : 1.0, di1 = lane + 32 < nb ? w.cdinv[k0 + lane + 32] : 1.0;
            for (int m = nb - 1; m >= 0; --m) {
                const double xm = __shfl_sync(0xffffffffu, m < 32 ? s0 * di0 : s1 * di1, m & 31);
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
constexpr size_t kTrsmSmem = sizeof(double) * (kBT * kBTP + kBT * kXS + kBT);

cudaError_t configure_chol_big(int n_cap) {
    cudaError_t e = opt_in_max_smem(bigchol_update_kernel);
    if (e != cudaSuccess) return e;
    e = opt_in_max_smem(bigchol_trsm_kernel);
    if (e != cudaSuccess) return e;
    return opt_in_max_smem(bigchol_backsub_kernel);
}

cudaError_t launch_chol_big(cudaStream_t s, cudaStream_t side, cudaEvent_t ev_trsm, cudaEvent_t ev_rest, const DevWindow* wp,
                            const LaunchDims& d) {
    const int ntile = d.chol_big_tiles;
    static const int tpr = std::getenv("VILBA_POTRF_TPR") ? std::atoi(std::getenv("VILBA_POTRF_TPR")) : 8;
    static const bool lookahead = !(std::getenv("VILBA_CHOL_LOOKAHEAD") && std::atoi(std::getenv("VILBA_CHOL_LOOKAHEAD")) == 0);
    cudaError_t e;
    bool rest_pending = false;
    for (int k = 0; k < ntile; ++k) {
        if (tpr == 2)
            bigchol_potrf_kernel<2><<<dim3(1, d.n_windows), 128, 0, s>>>(wp, k);
        else if (tpr == 16)
            bigchol_potrf_kernel<16><<<dim3(1, d.n_windows), 1024, 0, s>>>(wp, k);
        else if (tpr == 4)
            bigchol_potrf_kernel<4><<<dim3(1, d.n_windows), 256, 0, s>>>(wp, k);
        else
            bigchol_potrf_kernel<8><<<dim3(1, d.n_windows), 512, 0, s>>>(wp, k);
        const int rows = (ntile - k - 1) * kBT + 1;
        bigchol_trsm_kernel<<<dim3((rows + 63) / 64, d.n_windows), 256, kTrsmSmem, s>>>(wp, k);
        const int T = ntile - k - 1;
        if (T <= 0) continue;
        if (!lookahead) {
            bigchol_update_kernel<<<dim3(std::min(T * (T + 1) / 2, 4 * d.sm_count), d.n_windows), 256, kUpdateSmem, s>>>(wp, k, 0);
            continue;
        }
        // look-ahead: block column k + 1 on the main stream (the next potrf / trsm wait only for it), the rest of
        // the trailing matrix on the side stream, beside them
        if (rest_pending) {  // this step's tiles were last written by the previous step's rest
            if ((e = cudaStreamWaitEvent(s, ev_rest, 0)) != cudaSuccess) return e;
            rest_pending = false;
        }
        if (T > 1) {
            if ((e = cudaEventRecord(ev_trsm, s)) != cudaSuccess) return e;
            if ((e = cudaStreamWaitEvent(side, ev_trsm, 0)) != cudaSuccess) return e;
        }
        bigchol_update_kernel<<<dim3(T, d.n_windows), 256, kUpdateSmem, s>>>(wp, k, 1);
        if (T > 1) {
            bigchol_update_kernel<<<dim3(std::min((T - 1) * T / 2, 4 * d.sm_count), d.n_windows), 256, kUpdateSmem, side>>>(wp, k, 2);
            if ((e = cudaEventRecord(ev_rest, side)) != cudaSuccess) return e;
            rest_pending = true;
        }
    }
    if (rest_pending && (e = cudaStreamWaitEvent(s, ev_rest, 0)) != cudaSuccess) return e;
    bigchol_backsub_kernel<<<dim3(1, d.n_windows), 1024, chol_big_backsub_smem(d.chol_big_tiles * kBT), s>>>(wp);
    return cudaGetLastError();
}

}  // namespace vilba
