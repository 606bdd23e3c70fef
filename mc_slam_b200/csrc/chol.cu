// chol.cu -- K4b: dense FP64 Cholesky factor + solve of the reduced camera system S x = b_s
// (n = 15 * #free key-frames: 135 / 285 / 1485).  Replaces LinearSolverEigen::solve
// (Thirdparty/g2o/g2o/solvers/linear_solver_eigen.h:94-124; SimplicialLDLT on an effectively dense
// matrix).  A non-positive pivot raises LmState::chol_fail => the LM controller rejects the trial,
// like the `ok2 == false` path of optimization_algorithm_levenberg.cpp:126-127.
//
// The factorisation is latency-bound (n sequential pivots), so everything is organised around the
// serial chain, for one thread-block cluster (1..16 CTAs on neighbouring SMs):
//   * the matrix stays in global memory (L2-resident, <= 17.6 MB); the upper triangle of the row-major
//     S is addressed as the lower triangle of a column-major matrix, L(i,j) at S[j*n+i];
//   * right-looking, NB=16 columns per step.  EVERY CTA redundantly factors the 16x16 diagonal block
//     and solves the whole panel (one row per thread), so nothing but the trailing matrix is exchanged;
//   * the diagonal block is factored by one warp in square-root-free LDL^T form with 4x4 micro-blocks:
//     the 4x4 pivot block is broadcast through shared memory once and factored redundantly by every
//     lane (4 reciprocals on the chain, no per-column communication); square roots are taken once per
//     block, off the chain;
//   * the rank-16 trailing update is split over the cluster in 4x4 register tiles (rows/columns strided
//     so that panel reads are bank-conflict free and global accesses coalesce); old values are
//     prefetched from L2 before the panel is ready; ONE hardware cluster barrier per step publishes them;
//   * the right-hand side rides along as an extra matrix row, so the forward substitution is free;
//     CTA 0 back-substitutes: per block one warp holds L11^T in registers (one multiply + one shuffle
//     + one fma on the chain per unknown), the other threads apply the block row with prefetched data.
#include <cooperative_groups.h>

#include "chol_common.cuh"
#include "vmath.cuh"

namespace cg = cooperative_groups;

namespace vilba {

namespace {


// Back substitution L^T x = y: the unknowns live in the registers of ONE warp (lane owns c = lane + 32 i);
// per unknown the serial chain is one multiply, one shuffle and one fma.  The 32 rows of L that a
// block of 32 unknowns needs are staged in shared memory by the other warps (double buffered, coalesced
// reads of the row-major factor) so the chain never waits for L2.
template <int SLOTS>
__device__ __forceinline__ void backsub_staged(const double* __restrict__ Lr, int ld, const double* __restrict__ rdiag,
                                               const double* __restrict__ yf, double* __restrict__ xout, int n,
                                               double* stage /* 2 x 16 x (32*SLOTS+1) */, int tid, int nt) {
    constexpr int LDB = 32 * SLOTS + 1;
    constexpr int HB = 16;  // rows per staged half-block
    const int lane = tid & 31;
    const int nslots = (n + 31) / 32;
    const int nhalf = 2 * nslots;  // half-blocks, processed from the last one down
    auto load_half = [&](int h, double* buf, int t0, int tn) {
        const int ncols = 32 * (h / 2 + 1);
        for (int i = t0; i < HB * ncols; i += tn) {
            const int r = i / ncols, c = i - r * ncols;
            const int j = HB * h + r;
            buf[r * LDB + c] = (j < n && c < j) ? Lr[(size_t)j * ld + c] : 0.0;
        }
    };
    load_half(nhalf - 1, stage, tid, nt);
    __syncthreads();
    double yv[SLOTS];
    if (tid < 32) {
#pragma unroll
        for (int i = 0; i < SLOTS; ++i) yv[i] = (lane + 32 * i < n) ? yf[lane + 32 * i] : 0.0;
    }
#pragma unroll
    for (int slot = SLOTS - 1; slot >= 0; --slot) {
        if (slot >= nslots) continue;
        const double rd = (tid < 32 && lane + 32 * slot < n) ? rdiag[lane + 32 * slot] : 0.0;
#pragma unroll
        for (int hh = 1; hh >= 0; --hh) {
            const int h = 2 * slot + hh;
            double* buf = stage + ((nhalf - 1 - h) & 1) * HB * LDB;
            double* nxt = stage + ((nhalf - h) & 1) * HB * LDB;
            if (tid >= 32) {
                if (h > 0) load_half(h - 1, nxt, tid - 32, nt - 32);
            } else {
#pragma unroll 4
                for (int r = HB - 1; r >= 0; --r) {
                    const int jj = HB * hh + r;
                    const double* row = buf + r * LDB;
                    double lv[SLOTS];
#pragma unroll
                    for (int i = 0; i < SLOTS; ++i) lv[i] = (i <= slot) ? row[lane + 32 * i] : 0.0;  // 0 for c >= j
                    const double xj = __shfl_sync(0xffffffffu, yv[slot] * rd, jj);
                    if (lane == jj) yv[slot] = xj;
#pragma unroll
                    for (int i = 0; i < SLOTS; ++i)
                        if (i <= slot) yv[i] -= lv[i] * xj;
                }
            }
            __syncthreads();
        }
    }
    if (tid < 32) {
#pragma unroll
        for (int i = 0; i < SLOTS; ++i)
            if (lane + 32 * i < n) xout[lane + 32 * i] = yv[i];
    }
}

}  // namespace

template <int NB>
__global__ void __launch_bounds__(NB == 32 ? 384 : kCholThreads) chol_cluster_kernel(const DevWindow* __restrict__ wp) {
    constexpr int LDP = NB + 1;  // padded shared-memory row stride of the panel (conflict-free column walks)
    const DevWindow w = wp[blockIdx.y];  // one window per grid row
    if (w.lm->phase != PH_TRIAL) return;  // uniform over the cluster: nobody reaches a cluster barrier
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = (int)cluster.block_rank();
    const int csize = (int)cluster.num_blocks();
    extern __shared__ double smem[];
    const int n = w.n;
    const int ld = w.lds;  // leading dimension of S / Lfac (n rounded up to a multiple of 4: aligned 16-byte row groups)
    double* Dt = smem;                        // NB x NB   unit-lower L11, transposed: Dt[k*NB + c] = l(c,k)
    double* Pn = Dt + NB * NB;                // (n + 8) x LDP  panel X = A21 L11^-T D^-1/2 (+ rhs row + zero pad)
    double* xs = Pn + (size_t)(n + 8) * LDP;  // n   solution during back substitution
    double* Dsq = xs + n;                     // NB  sqrt(d)
    double* Dis = Dsq + NB;                   // NB  1/sqrt(d)
    double* Wsm = Dis + NB;                   // 16 + 4 NB  warp-private scratch of the diagonal factorisation
    double* Stage = Wsm + 16 + 4 * NB;                 // 2 x 16 x 321  row stage of the back substitution (n <= 320)
    double* A = w.S;       // working matrix: column block k is only READ in step k, the trailing part is updated
    double* Lf = w.Lfac;   // factor output (same addressing); written by CTA 0, never read inside the loop
    double* y = w.bs;      // rhs row, updated like a matrix row
    double* yf = w.x;      // forward-substituted rhs (L^-1 b), later overwritten by the solution
    double* rdiag = w.cdinv;  // 1 / L(j,j), written by CTA 0 for the back substitution
    __shared__ int s_fail;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int lane = tid & 31;
    if (tid == 0) s_fail = 0;
    __syncthreads();
#ifdef VILBA_CHOL_TIMING
    long long tph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tq = clock64();
#define TPH(i) { long long tn_ = clock64(); tph[i] += tn_ - tq; tq = tn_; }
#else
#define TPH(i)
#endif

    // warp 0 is the dedicated factor warp; warps 1.. are the panel/tile workers
    const bool is_factor_warp = tid < 32;
    const int wt = tid - 32, nwt = nt - 32;  // worker thread index / count
    for (int j0 = 0; j0 < n; j0 += NB) {
        const int jb = min(NB, n - j0);
        const int rows_below = n - j0 - jb;
        const int m_rows = rows_below + 1;       // + the rhs row
        const int trp = (m_rows + 3) >> 2;       // 4-row groups of the panel
        // lower-triangular 4x4 tiles over (m_rows) x (rows_below): tile (ti,tj), tj <= ti
        const int ntiles = trp * (trp + 1) / 2;
        const int gthreads = csize * nwt;
        const int t_first = crank * nwt + wt;
        double xr[NB];  // worker: panel row being solved (the first one is prefetched before the block barrier)
        if (is_factor_warp) {
            // ---- factor warp: (1) load the diagonal block (identity padded), (3) factor, publish ----
            double drow[NB];
#pragma unroll
            for (int c = 0; c < NB; ++c) {
                double v = (c == lane) ? 1.0 : 0.0;
                if (lane < jb && c <= lane) v = A[(size_t)(j0 + c) * ld + j0 + lane];
                drow[c] = v;
            }
            TPH(0)
#ifdef VILBA_CHOL_TIMING
            const bool ok = (w.dbg_flags & 4) ? true : warp_ldlt_mb4<NB>(drow, lane, Wsm);
#else
            const bool ok = warp_ldlt_mb4<NB>(drow, lane, Wsm);
#endif
            if (!ok && lane == 0) s_fail = 1;
            double dr = 1.0;
#pragma unroll
            for (int c = 0; c < NB; ++c)
                if (c == lane) dr = drow[c];
            const double sq = sqrt(dr), isq = 1.0 / sq;
            if (lane < NB) {
                Dsq[lane] = sq;
                Dis[lane] = isq;
#pragma unroll
                for (int c = 0; c < NB; ++c) Dt[c * NB + lane] = (c < lane) ? drow[c] : 0.0;  // l(lane,c)
            }
            __syncwarp();
            if (crank == 0 && lane < jb) {  // Cholesky factor block: L(r,c) = l(r,c) sqrt(d_c), L(r,r) = sqrt(d_r)
#pragma unroll
                for (int c = 0; c < NB; ++c)
                    if (c <= lane) Lf[(size_t)(j0 + lane) * ld + j0 + c] = (c == lane) ? sq : drow[c] * Dsq[c];
                rdiag[j0 + lane] = isq;
            }
            TPH(1)
        } else {
            // ---- workers: (1) first panel row of the thread, (2) old values of its first tile ----
            {
                const int rr = wt;
                const bool is_rhs = (rr == rows_below);
                const int gi = j0 + jb + rr;
#pragma unroll
                for (int c = 0; c < NB; ++c)
                    xr[c] = (rr < m_rows && c < jb) ? (is_rhs ? y[j0 + c] : A[(size_t)(j0 + c) * ld + gi]) : 0.0;
            }
        }
        __syncthreads();
        TPH(2)
        // ---- (4) panel: Z L11^T = A21 with unit-lower L11 (right-looking inside the row: one fma on
        //      the serial chain per column), then X = Z D^-1/2.  Every CTA solves all rows.  Row r is
        //      stored at position (r & 3) * trp + (r >> 2) so that the 4 rows of a tile are read
        //      bank-conflict free; the rows of the last (partial) group are zero. ----
        if (!is_factor_warp) {
            for (int rr = wt; rr < 4 * trp; rr += nwt) {
                const bool is_rhs = (rr == rows_below);
                const int gi = j0 + jb + rr;
                if (rr != wt) {
#pragma unroll
                    for (int c = 0; c < NB; ++c)
                        xr[c] = (rr < m_rows && c < jb) ? (is_rhs ? y[j0 + c] : A[(size_t)(j0 + c) * ld + gi]) : 0.0;
                }
#ifdef VILBA_CHOL_TIMING
                if (!(w.dbg_flags & 2))
#endif
#pragma unroll
                for (int k = 0; k < NB - 1; ++k) {
#pragma unroll
                    for (int c = k + 1; c < NB; ++c) xr[c] -= xr[k] * Dt[k * NB + c];  // l(c,k)
                }
                double* prow = Pn + (size_t)((rr & 3) * trp + (rr >> 2)) * LDP;
#pragma unroll
                for (int c = 0; c < NB; ++c) {
                    xr[c] *= Dis[c];
                    prow[c] = xr[c];
                }
                // factor output (row-major copy of L and L^-1 b): rows dealt round-robin to the CTAs of the
                // cluster -- every CTA holds the whole panel -- as 16-byte stores (j0 and ld are multiples of 4)
                if (rr < m_rows && (rr % csize) == crank) {
                    if (is_rhs) {
#pragma unroll
                        for (int c = 0; c < NB; ++c)
                            if (c < jb) yf[j0 + c] = xr[c];
                    } else {
                        double2* dst = reinterpret_cast<double2*>(&Lf[(size_t)gi * ld + j0]);
#pragma unroll
                        for (int c = 0; c < NB; c += 2)
                            if (c < jb) dst[c >> 1] = make_double2(xr[c], xr[c + 1]);  // columns >= jb are zero padding
                    }
                }
            }
        }
        TPH(3)
        __syncthreads();
        TPH(4)
        // ---- (5) trailing update A22 -= X X^T (and rhs -= x_rhs X^T), lower tiles split over the cluster ----
#ifdef VILBA_CHOL_TIMING
        if (!is_factor_warp && !(w.dbg_flags & 1)) {
#else
        if (!is_factor_warp) {
#endif
            for (int t = t_first; t < ntiles; t += gthreads) {
                int ti, tj;
                decode_tile(t, trp, ti, tj);
                const bool interior = (ti > tj && 4 * ti + 3 < rows_below);
                double acc[4][4];
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
                const double* pr = Pn + (size_t)ti * LDP;
                const double* pc = Pn + (size_t)tj * LDP;
                const int sr = trp * LDP;
#pragma unroll
                for (int k = 0; k < NB; ++k) {
                    const double vr0 = pr[k], vr1 = pr[sr + k], vr2 = pr[2 * sr + k], vr3 = pr[3 * sr + k];
                    const double vc0 = pc[k], vc1 = pc[sr + k], vc2 = pc[2 * sr + k], vc3 = pc[3 * sr + k];
                    acc[0][0] += vr0 * vc0, acc[0][1] += vr0 * vc1, acc[0][2] += vr0 * vc2, acc[0][3] += vr0 * vc3;
                    acc[1][0] += vr1 * vc0, acc[1][1] += vr1 * vc1, acc[1][2] += vr1 * vc2, acc[1][3] += vr1 * vc3;
                    acc[2][0] += vr2 * vc0, acc[2][1] += vr2 * vc1, acc[2][2] += vr2 * vc2, acc[2][3] += vr2 * vc3;
                    acc[3][0] += vr3 * vc0, acc[3][1] += vr3 * vc1, acc[3][2] += vr3 * vc2, acc[3][3] += vr3 * vc3;
                }
                if (interior) {
                    // interior tile: the 4 rows of a column are 32 contiguous, 16-byte aligned bytes
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        double2* p = reinterpret_cast<double2*>(&A[(size_t)(j0 + jb + 4 * tj + b) * ld + (j0 + jb + 4 * ti)]);
                        const double2 lo = p[0], hi = p[1];
                        p[0] = make_double2(lo.x - acc[0][b], lo.y - acc[1][b]);
                        p[1] = make_double2(hi.x - acc[2][b], hi.y - acc[3][b]);
                    }
                } else {
#pragma unroll
                    for (int b = 0; b < 4; ++b)
#pragma unroll
                        for (int a = 0; a < 4; ++a) {
                            const int r = 4 * ti + a, c = 4 * tj + b;
                            if (r >= m_rows || c >= rows_below || c > r) continue;
                            double* p =
                                (r == rows_below) ? &y[j0 + jb + c] : &A[(size_t)(j0 + jb + c) * ld + (j0 + jb + r)];
                            *p -= acc[a][b];
                        }
                }
            }
        }
        TPH(5)
        // ---- (6) publish the updated trailing matrix to the whole cluster ----
#ifdef VILBA_CHOL_TIMING
        if (w.dbg_flags & 8)
            __syncthreads();
        else
#endif
        if (csize > 1)
            cluster.sync();
        else
            __syncthreads();
        TPH(6)
    }

    // ---- back substitution L^T x = y on CTA 0 ----
    if (crank != 0) return;
#ifdef VILBA_CHOL_TIMING
    if (w.dbg_flags & 16) return;
#endif
    __syncthreads();  // CTA 0's own writes to yf / Lf / rdiag are ordered by the block barrier
    double* stage = Pn;  // the panel area is free now
    if (n <= 160 && w.chol_stage) {
        backsub_staged<5>(Lf, ld, rdiag, yf, xs, n, Stage, tid, nt);
    } else if (n <= 320 && w.chol_stage) {
        backsub_staged<10>(Lf, ld, rdiag, yf, xs, n, Stage, tid, nt);
    } else {
        // large systems: the unknowns live in shared memory (same algorithm, one row per step, one warp)
        if (tid < 32) {
            for (int i = lane; i < n; i += 32) xs[i] = yf[i];
            __syncwarp();
            for (int j = n - 1; j >= 0; --j) {
                const double xj = xs[j] * rdiag[j];
                __syncwarp();
                if (lane == 0) xs[j] = xj;
                const double* row = Lf + (size_t)j * ld;
                for (int c = lane; c < j; c += 32) xs[c] -= row[c] * xj;
                __syncwarp();
            }
        }
    }
    (void)stage;
    __syncthreads();
    for (int i = tid; i < n; i += nt) w.x[i] = xs[i];
    if (tid == 0) w.lm->chol_fail = s_fail;
#ifdef VILBA_CHOL_TIMING
    TPH(7)
    if (tid == 0 && w.dbg) {
        for (int i = 0; i < 8; ++i) atomicAdd((unsigned long long*)&w.dbg[i], (unsigned long long)tph[i]);
        atomicAdd((unsigned long long*)&w.dbg[8], 1ull);
    }
#endif
}

bool chol_has_stage(int n_cap) { return n_cap <= 640; }
int chol_block_size(int n_cap) { return n_cap <= 640 ? 32 : 16; }  // 32 columns per step while the panel fits in shared memory

size_t chol_smem_bytes(int n) {
    const size_t nb = (size_t)chol_block_size(n);
    const size_t stage = chol_has_stage(n) ? (size_t)2 * 16 * 321 : 0;
    return sizeof(double) * (nb * nb + (size_t)(n + 8) * (nb + 1) + (size_t)n + 2 * nb + 16 + 4 * nb + stage);
}

cudaError_t configure_chol(const LaunchDims& d) {
    cudaError_t e = opt_in_max_smem(chol_cluster_kernel<16>);
    if (e != cudaSuccess) return e;
    e = opt_in_max_smem(chol_cluster_kernel<32>);
    if (e != cudaSuccess) return e;
    if (d.chol_cluster > 8) {
        e = cudaFuncSetAttribute(chol_cluster_kernel<16>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(chol_cluster_kernel<32>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    }
    return e;
}

cudaError_t launch_chol_cluster(cudaStream_t s, const DevWindow* wp, const LaunchDims& d) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(d.chol_cluster, d.n_windows, 1);
    cfg.blockDim = dim3(d.chol_nb == 32 ? 384 : kCholThreads, 1, 1);  // 32-column steps need > 128 registers per thread
    cfg.dynamicSmemBytes = d.smem_chol;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = d.chol_cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (d.chol_nb == 32) return cudaLaunchKernelEx(&cfg, chol_cluster_kernel<32>, wp);
    return cudaLaunchKernelEx(&cfg, chol_cluster_kernel<16>, wp);
}

}  // namespace vilba
