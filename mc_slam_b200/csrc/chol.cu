// chol.cu -- K4b: dense FP64 Cholesky factor + solve of the reduced camera system S x = b_s
// (n = 15 * #free key-frames: 135 / 285 / 1485).  Replaces LinearSolverEigen::solve
// (Thirdparty/g2o/g2o/solvers/linear_solver_eigen.h:94-124; SimplicialLDLT on an effectively dense
// matrix).  A non-positive pivot raises LmState::chol_fail => the LM controller rejects the trial,
// like the `ok2 == false` path of optimization_algorithm_levenberg.cpp:126-127.
//
// Latency-oriented design for one thread-block cluster (1..16 CTAs on neighbouring SMs):
//   * the matrix stays in global memory (L2-resident, <= 17.6 MB); the upper triangle of the row-major
//     S is addressed as the lower triangle of a column-major matrix, L(i,j) at S[j*n+i];
//   * right-looking, NB=16 columns per step.  EVERY CTA redundantly factors the 16x16 diagonal block
//     (one warp, rows in registers, broadcasts by shuffle) and solves the whole panel (one row per
//     thread), so the panel never has to be exchanged between CTAs;
//   * the rank-16 trailing update is split over the cluster in 4x4 register tiles whose old values are
//     prefetched from L2 before the panel is ready; ONE hardware cluster barrier per step publishes them;
//   * the right-hand side rides along as an extra matrix row, so the forward substitution costs
//     nothing; CTA 0 back-substitutes with per-block warp solves.
#include <cooperative_groups.h>

#include "kernels.h"
#include "vmath.cuh"

namespace cg = cooperative_groups;

namespace vilba {

namespace {

constexpr int NB = kCholNB;       // 16
constexpr int LDP = NB + 1;       // padded shared-memory row stride (bank-conflict free for column walks)

// Cholesky of a jb x jb block held one row per lane (row r in lane r, entries a[0..r]).
// Returns false if a pivot is not positive.
__device__ __forceinline__ bool warp_potrf16(double (&a)[NB], int lane, int jb, double& my_inv) {
    bool ok = true;
    my_inv = 0.0;
#pragma unroll
    for (int j = 0; j < NB; ++j) {
        if (j < jb) {
            const double ajj = __shfl_sync(0xffffffffu, a[j], j);
            if (!(ajj > 0.0)) ok = false;
            const double inv = 1.0 / sqrt(ajj);
            if (lane == j) my_inv = inv;  // 1 / L(j,j)
            if (lane == j)
                a[j] = ajj * inv;
            else if (lane > j)
                a[j] *= inv;
#pragma unroll
            for (int c = j + 1; c < NB; ++c) {
                const double lcj = __shfl_sync(0xffffffffu, a[j], c);  // L(c,j)
                if (c < jb && lane >= c) a[c] -= a[j] * lcj;
            }
        }
    }
    return ok;
}

}  // namespace

__global__ void __launch_bounds__(kCholThreads) chol_cluster_kernel(DevWindow w) {
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = (int)cluster.block_rank();
    const int csize = (int)cluster.num_blocks();
    extern __shared__ double smem[];
    const int n = w.n;
    double* D = smem;                         // NB x LDP   factored diagonal block L11
    double* Pn = smem + NB * LDP;             // (n + 1) x LDP  panel L21 (+ rhs row)
    double* xs = Pn + (size_t)(n + 1) * LDP;  // n  solution during back substitution
    double* Dinv = xs + n;                    // NB  reciprocals of the diagonal of L11
    double* A = w.S;       // working matrix: column block k is only READ in step k, the trailing part is updated
    double* Lf = w.Lfac;   // factor output (same addressing); written by CTA 0, never read inside the loop
    double* y = w.bs;      // rhs row, updated like a matrix row
    double* yf = w.x;      // forward-substituted rhs (L^-1 b), later overwritten by the solution
    __shared__ int s_fail;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int lane = tid & 31;
    if (tid == 0) s_fail = 0;
    __syncthreads();

    for (int j0 = 0; j0 < n; j0 += NB) {
        const int jb = min(NB, n - j0);
        const int rows_below = n - j0 - jb;
        const int m_rows = rows_below + 1;  // + the rhs row
        // ---- (1) loads: diagonal block rows into warp 0's registers, panel rows into registers ----
        double drow[NB];
        if (tid < 32) {
#pragma unroll
            for (int c = 0; c < NB; ++c)
                drow[c] = (lane < jb && c <= lane && c < jb) ? A[(size_t)(j0 + c) * n + j0 + lane] : 0.0;
        }
        // first panel row of this thread (prefetched before the diagonal block is ready)
        double xr0[NB];
        {
            const int rr = tid;
            const bool is_rhs = (rr == rows_below);
            const int gi = j0 + jb + rr;
#pragma unroll
            for (int c = 0; c < NB; ++c)
                xr0[c] = (rr < m_rows && c < jb) ? (is_rhs ? y[j0 + c] : A[(size_t)(j0 + c) * n + gi]) : 0.0;
        }
        // ---- (2) prefetch the old values of this thread's first trailing tile ----
        const int tr = (m_rows + 3) >> 2, tc = (rows_below + 3) >> 2;
        const int ntiles = tr * tc;
        const int gthreads = csize * nt;
        const int t_first = crank * nt + tid;
        double old[4][4];
        {
            const int t = t_first;
            const int tj = t / tr, ti = t - tj * tr;
            if (t < ntiles && ti >= tj) {
#pragma unroll
                for (int b = 0; b < 4; ++b)
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        const int r = 4 * ti + a, c = 4 * tj + b;
                        double v = 0.0;
                        if (r < m_rows && c < rows_below && c <= r)
                            v = (r == rows_below) ? y[j0 + jb + c] : A[(size_t)(j0 + jb + c) * n + (j0 + jb + r)];
                        old[a][b] = v;
                    }
            }
        }
        // ---- (3) factor the diagonal block (warp 0), publish to shared ----
        if (tid < 32) {
            double my_inv;
            const bool ok = warp_potrf16(drow, lane, jb, my_inv);
            if (!ok && lane == 0) s_fail = 1;
            if (lane < NB) Dinv[lane] = my_inv;
            if (lane < NB) {
#pragma unroll
                for (int c = 0; c < NB; ++c) D[lane * LDP + c] = drow[c];
            }
            if (crank == 0 && lane < jb) {
#pragma unroll
                for (int c = 0; c < NB; ++c)
                    if (c <= lane && c < jb) Lf[(size_t)(j0 + c) * n + j0 + lane] = drow[c];
            }
        }
        __syncthreads();
        // ---- (4) panel: X L11^T = A21, one row per thread (every CTA solves all rows) ----
        for (int rr = tid; rr < m_rows; rr += nt) {
            double xr[NB];
            const bool is_rhs = (rr == rows_below);
            const int gi = j0 + jb + rr;
            if (rr == tid) {
#pragma unroll
                for (int c = 0; c < NB; ++c) xr[c] = xr0[c];
            } else {
#pragma unroll
                for (int c = 0; c < NB; ++c)
                    xr[c] = (c < jb) ? (is_rhs ? y[j0 + c] : A[(size_t)(j0 + c) * n + gi]) : 0.0;
            }
#pragma unroll
            for (int c = 0; c < NB; ++c) {
                if (c < jb) {
                    double s = xr[c];
#pragma unroll
                    for (int k = 0; k < c; ++k) s -= xr[k] * D[c * LDP + k];
                    xr[c] = s * Dinv[c];
                }
            }
#pragma unroll
            for (int c = 0; c < NB; ++c) Pn[(size_t)rr * LDP + c] = xr[c];
            if (crank == 0) {
#pragma unroll
                for (int c = 0; c < NB; ++c)
                    if (c < jb) {
                        if (is_rhs)
                            yf[j0 + c] = xr[c];
                        else
                            Lf[(size_t)(j0 + c) * n + gi] = xr[c];
                    }
            }
        }
        __syncthreads();
        // ---- (5) trailing update A22 -= P P^T (and rhs -= P_rhs P^T), tiles split over the cluster ----
        for (int t = t_first; t < ntiles; t += gthreads) {
            const int tj = t / tr, ti = t - tj * tr;
            if (ti < tj) continue;
            const int r0 = 4 * ti, c0 = 4 * tj;
            double acc[4][4];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
            const double* pr = Pn + (size_t)min(r0, m_rows - 1) * LDP;
            const double* pc = Pn + (size_t)min(c0, m_rows - 1) * LDP;
#pragma unroll
            for (int k = 0; k < NB; ++k) {
                double vr[4], vc[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    vr[a] = (r0 + a < m_rows) ? pr[a * LDP + k] : 0.0;
                    vc[a] = (c0 + a < rows_below) ? pc[a * LDP + k] : 0.0;
                }
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) acc[a][b] += vr[a] * vc[b];
            }
            const bool pre = (t == t_first);
#pragma unroll
            for (int b = 0; b < 4; ++b)
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const int r = r0 + a, c = c0 + b;
                    if (r >= m_rows || c >= rows_below || c > r) continue;
                    double* p = (r == rows_below) ? &y[j0 + jb + c] : &A[(size_t)(j0 + jb + c) * n + (j0 + jb + r)];
                    const double o = pre ? old[a][b] : *p;
                    *p = o - acc[a][b];
                }
        }
        // ---- (6) publish the updated trailing matrix to the whole cluster ----
        if (csize > 1)
            cluster.sync();
        else
            __syncthreads();
    }

    // ---- back substitution L^T x = y on CTA 0 ----
    if (crank != 0) return;
    __syncthreads();  // CTA 0's own writes to yf / Lf are ordered by the block barrier
    for (int i = tid; i < n; i += nt) xs[i] = yf[i];
    __syncthreads();
    const int nblk = (n + NB - 1) / NB;
    for (int kb = nblk - 1; kb >= 0; --kb) {
        const int j0 = kb * NB;
        const int jb = min(NB, n - j0);
        // prefetch this thread's slice of block row kb of L for the update below: L(j0+t, c), c < j0
        // (one column c per thread)
        if (tid < 32) {
            // solve L11^T x_b = y_b : row r of L11 in lane r
            double lrow[NB];
#pragma unroll
            for (int c = 0; c < NB; ++c) lrow[c] = (lane < jb && c <= lane) ? Lf[(size_t)(j0 + c) * n + j0 + lane] : 0.0;
            double xv = (lane < jb) ? xs[j0 + lane] : 0.0;
#pragma unroll
            for (int j = NB - 1; j >= 0; --j) {
                if (j < jb) {
                    // x_j = y_j / L(j,j); then y_c -= L(j,c) x_j for c < j   (L(j,c) lives in lane j, entry c)
                    const double ljj = __shfl_sync(0xffffffffu, lrow[j], j);
                    const double xj = __shfl_sync(0xffffffffu, xv, j) / ljj;
                    if (lane == j) xv = xj;
                    // lane c needs L(j,c): held by lane j at index c -> dynamic index; use a shuffle per c
#pragma unroll
                    for (int c = 0; c < NB; ++c) {
                        const double ljc = __shfl_sync(0xffffffffu, lrow[c], j);
                        if (lane == c && c < j) xv -= ljc * xj;
                    }
                }
            }
            if (lane < jb) xs[j0 + lane] = xv;
        }
        __syncthreads();
        // y_c -= sum_t L(j0+t, c) x(j0+t) for all c < j0
        for (int c = tid; c < j0; c += nt) {
            const double* col = Lf + (size_t)c * n + j0;
            double s = 0.0;
#pragma unroll
            for (int t = 0; t < NB; ++t)
                if (t < jb) s += col[t] * xs[j0 + t];
            xs[c] -= s;
        }
        __syncthreads();
    }
    for (int i = tid; i < n; i += nt) w.x[i] = xs[i];
    if (tid == 0) w.lm->chol_fail = s_fail;
}

static size_t chol_cluster_smem(int n) {
    return sizeof(double) * ((size_t)NB * LDP + (size_t)(n + 1) * LDP + (size_t)n + NB);
}

cudaError_t launch_chol_cluster(cudaStream_t s, const DevWindow& w, int cluster_size) {
    const size_t sm = chol_cluster_smem(w.n);
    static size_t configured = 0;
    static bool nonportable = false;
    cudaError_t e;
    if (sm > configured) {
        e = cudaFuncSetAttribute(chol_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e != cudaSuccess) return e;
        configured = sm;
    }
    if (cluster_size > 8 && !nonportable) {
        e = cudaFuncSetAttribute(chol_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return e;
        nonportable = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cluster_size, 1, 1);
    cfg.blockDim = dim3(kCholThreads, 1, 1);
    cfg.dynamicSmemBytes = sm;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster_size;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, chol_cluster_kernel, w);
}

}  // namespace vilba
