// chol_la.cu -- K4b for small and medium reduced camera systems (n <= ~500: BASELINE configs 1, 3, 5):
// dense FP64 LDL^T factor + solve of S x = b_s on ONE thread-block cluster, with the trailing matrix resident in
// (distributed) shared memory and a look-ahead factor group.  Replaces LinearSolverEigen::solve
// (Thirdparty/g2o/g2o/solvers/linear_solver_eigen.h:94-124: SimplicialLDLT factorize + solve on an effectively
// dense matrix) with the same failure rule: negative pivots are factored through, only a zero (or non-finite)
// pivot raises LmState::chol_fail and makes the LM controller reject the trial
// (optimization_algorithm_levenberg.cpp:126-127).
//
// Why this shape.  The factorisation is a chain of n dependent pivots; 7.7 MFLOP at n = 285 are nothing for 8 SMs,
// so everything is organised around the chain and around NOT touching L2 on it:
//   * the lower triangle lives as 4x4 tiles in the shared memory of the cluster's CTAs, every tile owned by one
//     worker thread for the whole factorisation (no read-modify-write of the trailing matrix through L2, no
//     synchronisation on tiles).  Only the NEXT block column (32 columns) is published to global memory per step;
//   * right-looking, 32 columns per step.  The panel rows (rhs riding as one more row = free forward substitution)
//     are dealt round-robin to the CTAs of the cluster, one row per thread, written to global memory (they are the
//     rows of the factor the back substitution needs anyway) and fetched back by every CTA into its shared-memory
//     panel after a cluster barrier: a row solve streams the whole diagonal factor from shared memory for ONE row
//     (8 bytes x 32 lanes per fma), so solving all rows in every CTA is bound by the 128 B/clk shared-memory port
//     (measured 13 k cycles per step), the exchange costs ~2 k;
//   * look-ahead: warps 0..3 of every CTA are a "factor group" that owns no tiles.  While the workers of step k
//     solve their panel rows and update their tiles, the factor group solves the 32 panel rows of the next diagonal
//     block itself (4 threads per row), applies the step-k update to that block (published one step early into a
//     scratch buffer, i.e. with the updates of steps < k) and factors it (chol_fg.cuh: square-root free 4x4
//     micro-block LDL^T spread over the four warps).  The 32-pivot chain therefore runs BESIDE the panel / update
//     work of the previous step instead of in front of it;
//   * back substitution with explicitly inverted diagonal blocks (inverted by an otherwise idle worker warp while
//     the panel rows are being solved): per block 16 warps reduce the rows below the block (prefetched one block
//     ahead), one warp applies the 32 x 32 inverse; no per-unknown chain.
// Signs: S = L |D|^(1/2) Sigma |D|^(1/2) L^T with Sigma = diag(sign d).  The panel stores X = Z |D|^(-1/2) Sigma,
// the update subtracts sum_k sigma_k x_r[k] x_c[k]; since the rhs is processed exactly like a matrix row, the
// forward-substituted row equals Sigma (L |D|^(1/2))^-1 b and the back substitution needs no sign at all.
#include <cooperative_groups.h>

#include "chol_fg.cuh"

namespace cg = cooperative_groups;

namespace vilba {

namespace {

constexpr int kLaThreads = 512;
constexpr int kLaFgWarps = 4;                      // factor group: warps 0..3 (one per SM sub-partition)
constexpr int kLaFg = 32 * kLaFgWarps;
constexpr int kLaWorkers = kLaThreads - kLaFg;     // 384 = 12 warps
constexpr int kLaWorkerWarps = kLaWorkers / 32;
constexpr int kNB = 32;                            // columns per step
constexpr int kLDP = kNB + 1;                      // row stride of the panel in shared memory
constexpr int kXDS = kNB + 2;                      // row stride of the factor group's transposed panel rows (16-byte rows)

__device__ __forceinline__ void bar_workers() { asm volatile("bar.sync 1, %0;" ::"n"(kLaWorkers) : "memory"); }
__device__ __forceinline__ void bar_fg() { asm volatile("bar.sync 2, %0;" ::"n"(kLaFg) : "memory"); }
// mid-step barrier of a one-CTA "cluster": the workers wait for everybody, the factor group only signals
__device__ __forceinline__ void bar_mid_wait() { asm volatile("bar.sync 4, %0;" ::"n"(kLaThreads) : "memory"); }
__device__ __forceinline__ void bar_mid_arrive() { asm volatile("bar.arrive 4, %0;" ::"n"(kLaThreads) : "memory"); }

struct CholJob {
    int n, ld;
    double* A;     // S, lower triangle column-major: A(i,j), i >= j, at A[j * ld + i]; block columns get overwritten
    double* y;     // right-hand side (n), overwritten
    double* x;     // out: solution (n); holds the forward-substituted rhs in between
    double* Lf;    // n * ld: rows of the factor below the diagonal blocks, row-major (for the back substitution)
    double* scr;   // scratch, chol_la_scratch_doubles(n): 2 look-ahead diagonal blocks, the diagonal factor blocks, |d|^-1/2
    int* fail;     // out: 1 if a pivot was zero / not finite
    int nt;        // tiles per worker thread
    long long* dbg;  // optional phase timers (-DVILBA_LA_TIMING builds only)
};

#ifdef VILBA_LA_TIMING
#define LA_T0() long long tq_ = clock64(); long long tph_[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}
#define LA_T(i) { const long long tn_ = clock64(); tph_[i] += tn_ - tq_; tq_ = tn_; }
#define LA_T0B() const long long tb_ = clock64()
#else
#define LA_T0()
#define LA_T(i)
#define LA_T0B()
#endif

struct LaLayout {
    // offsets in doubles; everything but the tiles sits at a compile-time offset (addresses become immediates)
    static constexpr int dt = 0;                           // [2][NB*NB] unit-lower factor of the diagonal block, transposed
    static constexpr int dis = dt + 2 * kNB * kNB;         // [2][NB] signed |d|^-1/2
    static constexpr int dab = dis + 2 * kNB;              // [2][NB] |d|^-1/2
    static constexpr int sg = dab + 2 * kNB;               // [2][NB] sign of d
    static constexpr int fgs = sg + 2 * kNB;               // scratch of fg4_factor
    static constexpr int xdt = fgs + kFgScratch;           // [NB][kXDS] panel rows of the next diagonal block, transposed
    static constexpr int fgin = xdt + kNB * kXDS;          // [2][NB*NB] pushed by the tile owners of the whole cluster:
                                                           //   [0] rows of the next diagonal block in the current block column,
                                                           //   [1] the diagonal block after the next (lower triangle), both [c][r]
    static constexpr int pn = fgin + 2 * kNB * kNB;        // panel
    int tl, total;
    __host__ __device__ LaLayout(int n, int nt, int cluster) {
        // panel rows; the area doubles as scratch of the row solves (one padded row per solving thread: this CTA's share
        // of the panel rows + the 32 rows of the identity)
        const int prow = ((n > kNB ? n - kNB : 0) + 1 + 3) & ~3;
        const int jobs = ((n > kNB ? n - kNB : 0) + cluster) / cluster + kNB;
        const int srow = jobs < kLaWorkers ? jobs : kLaWorkers;
        const int rows = prow > srow ? prow : srow;
        const int work = rows * kLDP + nt * 16 * kLaWorkers;   // panel + tiles ...
        const int back = backsub_smem_doubles(n, kLaThreads);          // ... reused by the back substitution
        tl = pn + rows * kLDP;
        total = pn + (work > back ? work : back);
    }
};
static_assert(LaLayout::fgs % 2 == 0 && LaLayout::xdt % 2 == 0 && LaLayout::fgin % 2 == 0, "16-byte alignment of the vector accesses");

__device__ __forceinline__ void chol_la_body(const CholJob& J, double* smem) {
    constexpr int NB = kNB, LDP = kLDP;
    constexpr int TPB = NB / 4;  // tile columns per block column
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = (int)cluster.block_rank();
    const int csize = (int)cluster.num_blocks();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool is_fg = warp < kLaFgWarps;
    const int wt = tid - kLaFg, ww = warp - kLaFgWarps;
    const int n = J.n, ld = J.ld;
    const int TR = (n + 4) >> 2;  // tile rows, the rhs row (index n) included
    const int TC = (n + 3) >> 2;  // tile columns
    const int ntiles = TC * (2 * TR - TC + 1) / 2;
    const int nblk = (n + NB - 1) / NB;
    double* const Dt = smem + LaLayout::dt;
    double* const Dis = smem + LaLayout::dis;
    double* const Dab = smem + LaLayout::dab;
    double* const Sg = smem + LaLayout::sg;
    double* const Fgs = smem + LaLayout::fgs;
    double* const XdT = smem + LaLayout::xdt;
    double* const FgIn = smem + LaLayout::fgin;
    double* const Pn = smem + LaLayout::pn;
    double* const Tl = smem + LaLayout(n, J.nt, csize).tl;
    double* A = J.A;
    double* y = J.y;
    double* yf = J.x;
    double* Minv_g = J.scr;                               // [nblk][NB*NB] inverse of every diagonal factor block, row-major
    __shared__ int s_fail;
    __shared__ int s_neg[2];
    if (tid == 0) s_fail = 0, s_neg[0] = 0, s_neg[1] = 0;
    __syncthreads();
    LA_T0();

    // factor group: factor the diagonal block `blk` held in a[2][4] (chol_fg.cuh layout), publish it as buffer `buf`
    auto fg_factor_publish = [&](double (&a)[2][4], int buf) {
        const FgPivots p = fg4_factor(a, lane, warp, Dt + buf * NB * NB, Fgs, 3);
        if (!p.ok) s_fail = 1;
        if (warp == 0) {
            const double sgn = p.d < 0.0 ? -1.0 : 1.0;
            const double isq = 1.0 / sqrt(fabs(p.d));
            const unsigned anyneg = __ballot_sync(0xffffffffu, p.d < 0.0);
            Dis[buf * NB + lane] = isq * sgn;
            Dab[buf * NB + lane] = isq;
            Sg[buf * NB + lane] = sgn;
            if (lane == 0) s_neg[buf] = anyneg != 0;
        }
    };
    // global tile index of slot i of this worker thread: warps' worth of 32 consecutive tiles dealt round-robin to
    // the CTAs of the cluster, then to the worker warps
    auto tile_index = [&](int i) { return 32 * (crank + csize * (ww + kLaWorkerWarps * i)) + lane; };
    // what the factor groups read at the start of a step travels through distributed shared memory, not L2: the owner of a
    // tile stores it straight into the [c][r] input buffer of EVERY CTA of the cluster (visible behind the cluster barrier)
    auto push_tile = [&](double* buf0, int rl0, int cl0, const double (&t)[4][4]) {
        for (int q = 0; q < csize; ++q) {
            double* p = cluster.map_shared_rank(buf0, q);
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                double2* d = reinterpret_cast<double2*>(p + (cl0 + b) * NB + rl0);
                d[0] = make_double2(t[0][b], t[1][b]);
                d[1] = make_double2(t[2][b], t[3][b]);
            }
        }
    };

    // ---------------------------------------------------------------------------------------------
    // prologue: workers fetch their tiles (and publish the look-ahead copy of diagonal block 1), the factor
    // group factors diagonal block 0
    // ---------------------------------------------------------------------------------------------
    if (!is_fg) {
        for (int i = 0; i < J.nt; ++i) {
            const int t = tile_index(i);
            if (t >= ntiles) break;
            int ti, tj;
            decode_tile(t, TR, ti, tj);
            double* slot = Tl + (size_t)i * 16 * kLaWorkers + wt;
            double t4[4][4];
#pragma unroll
            for (int b = 0; b < 4; ++b)
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const int r = 4 * ti + a, c = 4 * tj + b;
                    double v = 0.0;
                    if (c < n && c <= r && r <= n) v = (r == n) ? y[c] : A[(size_t)c * ld + r];
                    slot[(a + 4 * b) * kLaWorkers] = v;
                    t4[a][b] = v;
                }
            if (ti >= TPB && ti < 2 * TPB) {  // inputs of the factor groups in step 0
                if (tj < TPB) push_tile(FgIn, 4 * ti - NB, 4 * tj, t4);
                else push_tile(FgIn + NB * NB, 4 * ti - NB, 4 * tj - NB, t4);
            }
        }
    } else {
        const int jb = min(NB, n);
        double a[2][4];
#pragma unroll
        for (int sl = 0; sl < 2; ++sl)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int c = 4 * (warp + 4 * sl) + i;
                a[sl][i] = (lane < jb && c <= lane) ? A[(size_t)c * ld + lane] : (c == lane ? 1.0 : 0.0);
            }
        fg_factor_publish(a, 0);
    }
    LA_T(is_fg ? 0 : 6)
    if (csize > 1) cluster.sync(); else __syncthreads();
    LA_T(is_fg ? 5 : 11)

    for (int k = 0; k < nblk; ++k) {
        const int j0 = k * NB;
        const int jb = min(NB, n - j0);
        const int r0 = j0 + jb;              // first row / column of the trailing matrix
        const int rows_below = n - r0;
        const int m_rows = rows_below + 1;   // + the rhs row
        const int trp = (m_rows + 3) >> 2;   // 4-row groups of the panel
        const int tcol0 = r0 >> 2;
        const int buf = k & 1;
        const double* const Dtk = Dt + buf * NB * NB;
        const double* Disk = Dis + buf * NB;
        if (!is_fg) {
            // ---- panel: Z L11^T = A21 with unit-lower L11, X = Z |D|^-1/2 Sigma.  (A) this CTA's share of the rows (round
            //      robin over the cluster), one row per thread, straight to global memory: rows of the factor for the back
            //      substitution / L^-1 b.  Meanwhile the last worker warp of CTA 0 inverts the diagonal block ----
            {
                // rows crank, crank + csize, ... of the panel, then (CTA 0) the 32 rows of the identity: a row e_j solved
                // against L11 is column j of L11^-1, i.e. the inverse of the diagonal block for the back substitution
                const int n_mine = (m_rows - crank + csize - 1) / csize;
                const int n_jobs = n_mine + (crank == 0 ? NB : 0);
                double* const scr = Pn + (size_t)wt * LDP;  // the panel area is idle until the fetch below
                for (int job = wt; job < n_jobs; job += kLaWorkers) {
                    const bool inv = job >= n_mine;   // a row of the identity
                    const int rr = crank + csize * job;
                    const bool is_rhs = !inv && (rr == rows_below);
                    const double* src = is_rhs ? y + j0 : A + (size_t)j0 * ld + (r0 + rr);
                    const long long stride = is_rhs ? 1 : ld;
                    double lo[16], hi[16];
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        lo[c] = (!inv && c < jb) ? __ldcg(src + c * stride) : ((inv && c == job - n_mine) ? 1.0 : 0.0);
                        hi[c] = (!inv && 16 + c < jb) ? __ldcg(src + (16 + c) * stride) : ((inv && 16 + c == job - n_mine) ? 1.0 : 0.0);
                    }
#pragma unroll
                    for (int c = 0; c < 16; ++c) scr[16 + c] = hi[c];  // parked until the first half is done
                    rowsolve_lo(lo, Dtk);
                    double* gout = inv ? Minv_g + (size_t)k * NB * NB + (job - n_mine)
                                       : (is_rhs ? yf + j0 : J.Lf + (size_t)(r0 + rr) * ld + j0);
                    const double* scale = inv ? Dab + buf * NB : Disk;
                    const long long ostride = inv ? NB : 1;  // Minv[c][j] = |d_c|^-1/2 (L11^-1)[c][j], row-major
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        scr[c] = lo[c];
                        if (inv || c < jb) gout[c * ostride] = lo[c] * scale[c];
                    }
#pragma unroll
                    for (int c = 0; c < 16; ++c) hi[c] = scr[16 + c];
                    rowsolve_hi(hi, scr, Dtk);
#pragma unroll
                    for (int c = 0; c < 16; ++c)
                        if (inv || 16 + c < jb) gout[(16 + c) * ostride] = hi[c] * scale[16 + c];
                }
            }
            LA_T(7)
            if (csize > 1) {
                cluster.barrier_arrive();
                cluster.barrier_wait();
            } else {
                bar_mid_wait();
            }
            LA_T(8)
            // ---- (B) every CTA fetches the whole panel into shared memory.  Row r is stored at position
            //      (r & 3) * trp + (r >> 2), so that the 4 rows of a tile are read conflict free; the rows of the last
            //      (partial) group are zero ----
            if (rows_below > 0) {
                for (int rr = wt; rr < 4 * trp; rr += kLaWorkers) {
                    double* prow = Pn + (size_t)((rr & 3) * trp + (rr >> 2)) * LDP;
                    if (rr < m_rows) {
                        const double2* g = reinterpret_cast<const double2*>(rr == rows_below ? yf + j0 : J.Lf + (size_t)(r0 + rr) * ld + j0);
                        double2 v[NB / 2];
#pragma unroll
                        for (int c = 0; c < NB / 2; ++c) v[c] = __ldcg(g + c);
#pragma unroll
                        for (int c = 0; c < NB / 2; ++c) prow[2 * c] = v[c].x, prow[2 * c + 1] = v[c].y;
                    } else {
#pragma unroll
                        for (int c = 0; c < NB; ++c) prow[c] = 0.0;
                    }
                }
            }
            bar_workers();
            LA_T(9)
            // ---- trailing update of the owned tiles, A22 -= X Sigma X^T, and publication of the next block column ----
            if (rows_below > 0) {
                const bool neg = s_neg[buf] != 0;
                const double* Sgk = Sg + buf * NB;
                const int sr = trp * LDP;
                for (int i = 0; i < J.nt; ++i) {
                    const int t = tile_index(i);
                    if (t >= ntiles) break;
                    int ti, tj;
                    decode_tile(t, TR, ti, tj);
                    if (tj < tcol0) continue;  // finished columns
                    double* slot = Tl + (size_t)i * 16 * kLaWorkers + wt;
                    double acc[4][4];
#pragma unroll
                    for (int b = 0; b < 4; ++b)
#pragma unroll
                        for (int a = 0; a < 4; ++a) acc[a][b] = slot[(a + 4 * b) * kLaWorkers];
                    const double* pr = Pn + (size_t)(ti - tcol0) * LDP;
                    const double* pc = Pn + (size_t)(tj - tcol0) * LDP;
                    if (!neg) {
#pragma unroll
                        for (int kk = 0; kk < NB; ++kk) {
                            const double vr0 = pr[kk], vr1 = pr[sr + kk], vr2 = pr[2 * sr + kk], vr3 = pr[3 * sr + kk];
                            const double vc0 = pc[kk], vc1 = pc[sr + kk], vc2 = pc[2 * sr + kk], vc3 = pc[3 * sr + kk];
                            acc[0][0] -= vr0 * vc0, acc[0][1] -= vr0 * vc1, acc[0][2] -= vr0 * vc2, acc[0][3] -= vr0 * vc3;
                            acc[1][0] -= vr1 * vc0, acc[1][1] -= vr1 * vc1, acc[1][2] -= vr1 * vc2, acc[1][3] -= vr1 * vc3;
                            acc[2][0] -= vr2 * vc0, acc[2][1] -= vr2 * vc1, acc[2][2] -= vr2 * vc2, acc[2][3] -= vr2 * vc3;
                            acc[3][0] -= vr3 * vc0, acc[3][1] -= vr3 * vc1, acc[3][2] -= vr3 * vc2, acc[3][3] -= vr3 * vc3;
                        }
                    } else {  // a negative pivot in this block column (indefinite S): carry the signs
#pragma unroll 4
                        for (int kk = 0; kk < NB; ++kk) {
                            const double sk = Sgk[kk];
                            const double vr0 = pr[kk] * sk, vr1 = pr[sr + kk] * sk, vr2 = pr[2 * sr + kk] * sk, vr3 = pr[3 * sr + kk] * sk;
                            const double vc0 = pc[kk], vc1 = pc[sr + kk], vc2 = pc[2 * sr + kk], vc3 = pc[3 * sr + kk];
                            acc[0][0] -= vr0 * vc0, acc[0][1] -= vr0 * vc1, acc[0][2] -= vr0 * vc2, acc[0][3] -= vr0 * vc3;
                            acc[1][0] -= vr1 * vc0, acc[1][1] -= vr1 * vc1, acc[1][2] -= vr1 * vc2, acc[1][3] -= vr1 * vc3;
                            acc[2][0] -= vr2 * vc0, acc[2][1] -= vr2 * vc1, acc[2][2] -= vr2 * vc2, acc[2][3] -= vr2 * vc3;
                            acc[3][0] -= vr3 * vc0, acc[3][1] -= vr3 * vc1, acc[3][2] -= vr3 * vc2, acc[3][3] -= vr3 * vc3;
                        }
                    }
#pragma unroll
                    for (int b = 0; b < 4; ++b)
#pragma unroll
                        for (int a = 0; a < 4; ++a) slot[(a + 4 * b) * kLaWorkers] = acc[a][b];
                    // publish: block column k+1 below its diagonal block -> S / rhs (read by every CTA's panel solve of
                    // step k+1); diagonal block k+2 -> look-ahead scratch (read by the factor groups during step k+1)
                    if (tj < tcol0 + TPB) {
                        if (ti >= tcol0 + TPB && 4 * ti + 3 < n) {  // interior: 4 rows of a column = 32 aligned bytes
#pragma unroll
                            for (int b = 0; b < 4; ++b) {
                                double2* p = reinterpret_cast<double2*>(&A[(size_t)(4 * tj + b) * ld + 4 * ti]);
                                p[0] = make_double2(acc[0][b], acc[1][b]);
                                p[1] = make_double2(acc[2][b], acc[3][b]);
                            }
                        } else {
#pragma unroll
                            for (int b = 0; b < 4; ++b)
#pragma unroll
                                for (int a = 0; a < 4; ++a) {
                                    const int r = 4 * ti + a, c = 4 * tj + b;
                                    if (c >= n || r > n) continue;
                                    if (r == n) y[c] = acc[a][b];
                                    else if (r >= r0 + NB) A[(size_t)c * ld + r] = acc[a][b];
                                }
                        }
                    }
                    // inputs of the factor groups in step k+1: the rows of diagonal block k+2 in block column k+1, and
                    // diagonal block k+2 itself (so far updated by the steps <= k)
                    if (ti >= tcol0 + TPB && ti < tcol0 + 2 * TPB && tj < tcol0 + 2 * TPB) {
                        // (ONE buffer: the factor groups read it into registers before they arrive at the mid-step
                        // barrier, the workers get here behind that barrier)
                        if (tj < tcol0 + TPB) push_tile(FgIn, 4 * ti - (r0 + NB), 4 * tj - r0, acc);
                        else push_tile(FgIn + NB * NB, 4 * ti - (r0 + NB), 4 * tj - (r0 + NB), acc);
                    }
                }
            }
            LA_T(10)
        } else if (rows_below > 0) {
            // ---- factor group: look-ahead factorisation of diagonal block k+1 (jb == NB here) ----
            const int jbn = min(NB, rows_below);
            // (1) the jbn panel rows of the next diagonal block, one thread per row (warp 0), into XdT[c][row].  The inputs
            //     were pushed into FgIn by the tile owners in the previous step
            if (warp == 0) {
                double lo[16], hi[16];
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    lo[c] = (lane < jbn) ? FgIn[c * NB + lane] : 0.0;
                    hi[c] = (lane < jbn) ? FgIn[(16 + c) * NB + lane] : 0.0;
                }
                double* const scr = Fgs + lane * 17;  // fg4_factor's scratch is idle here (608 doubles >= 32 * 17)
                rowsolve_lo(lo, Dtk);
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    scr[c] = lo[c];
                    XdT[c * kXDS + lane] = lo[c] * Disk[c];
                }
                rowsolve_hi(hi, scr, Dtk);
#pragma unroll
                for (int c = 0; c < 16; ++c) XdT[(16 + c) * kXDS + lane] = hi[c] * Disk[16 + c];
            }
            LA_T(1)
            bar_fg();
            LA_T(2)
            // (2) D' = D(k+1, with the updates of the steps < k) - Xd Sigma Xd^T in the register layout of fg4_factor:
            //     lane = row, warp w = columns [4w, 4w+4) and [16+4w, 16+4w+4).  With D in registers FgIn is free: arrive at
            //     the mid-step barrier (the workers overwrite FgIn behind it; they get there after their own row solves,
            //     i.e. not before this point: nobody waits)
            double a[2][4];
#pragma unroll
            for (int sl = 0; sl < 2; ++sl)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int c = 4 * (warp + 4 * sl) + i;
                    a[sl][i] = (lane < jbn && c <= lane) ? FgIn[NB * NB + c * NB + lane] : (c == lane ? 1.0 : 0.0);
                }
            if (csize > 1) cluster.barrier_arrive(); else bar_mid_arrive();
            {
                const bool neg = s_neg[buf] != 0;
                const double* Sgk = Sg + buf * NB;
                double acc[2][4] = {{0.0, 0.0, 0.0, 0.0}, {0.0, 0.0, 0.0, 0.0}};
#pragma unroll 8
                for (int kk = 0; kk < NB; ++kk) {
                    double own = XdT[kk * kXDS + lane];
                    if (neg) own *= Sgk[kk];
                    const double2* cp0 = reinterpret_cast<const double2*>(XdT + kk * kXDS + 4 * warp);
                    const double2* cp1 = reinterpret_cast<const double2*>(XdT + kk * kXDS + 16 + 4 * warp);
                    const double2 c0a = cp0[0], c0b = cp0[1], c1a = cp1[0], c1b = cp1[1];
                    acc[0][0] = fma(own, c0a.x, acc[0][0]), acc[0][1] = fma(own, c0a.y, acc[0][1]);
                    acc[0][2] = fma(own, c0b.x, acc[0][2]), acc[0][3] = fma(own, c0b.y, acc[0][3]);
                    acc[1][0] = fma(own, c1a.x, acc[1][0]), acc[1][1] = fma(own, c1a.y, acc[1][1]);
                    acc[1][2] = fma(own, c1b.x, acc[1][2]), acc[1][3] = fma(own, c1b.y, acc[1][3]);
                }
#pragma unroll
                for (int sl = 0; sl < 2; ++sl)
#pragma unroll
                    for (int i = 0; i < 4; ++i) a[sl][i] -= acc[sl][i];
            }
            LA_T(3)
            // (3) factor it
            fg_factor_publish(a, buf ^ 1);
            LA_T(4)
            if (csize > 1) cluster.barrier_wait();  // (the workers passed it long ago)
        } else if (csize > 1) {  // last step: nothing to factor, but the mid-step barrier counts every thread
            cluster.barrier_arrive();
            cluster.barrier_wait();
        } else {
            bar_mid_arrive();
        }
        if (csize > 1) cluster.sync(); else __syncthreads();
        LA_T(is_fg ? 5 : 11)
    }
#ifdef VILBA_LA_TIMING
    // [0] prologue factor, [1] panel rows of the next diagonal block, [2] wait, [3] diagonal update, [4] factor,
    // [5] cluster barrier; workers: [6] prologue, [7] own panel rows, [8] exchange barrier, [9] panel fetch, [10] tiles,
    // [11] cluster barrier
    if (J.dbg && crank == 0 && (tid == 0 || tid == 32 || tid == kLaFg)) {
        long long* o = J.dbg + (tid == 0 ? 0 : tid == 32 ? 12 : 24);
        for (int i = 0; i < 12; ++i) o[i] = tph_[i];
    }
#endif

    // ---------------------------------------------------------------------------------------------
    // back substitution L~^T x = yf on CTA 0 (L~ = L |D|^1/2), block by block from the last one:
    //   x_k = Minv_k^T (yf_k - sum over the rows r below the block of L~(r, block)^T x_r).
    // The 16 warps split the rows below (lane = column of the block, 256-byte coalesced row segments, fetched one block
    // ahead so that L2 latency is off the chain), their partial sums meet in shared memory, warp 0 applies the inverse.
    // ---------------------------------------------------------------------------------------------
    if (crank != 0) return;
    LA_T0B();
    backsub_blocks<kLaThreads>(n, ld, J.Lf, Minv_g, yf, J.x, Pn);
    if (tid == 0) *J.fail = s_fail;
#ifdef VILBA_LA_TIMING
    if (J.dbg && tid == 0) J.dbg[11] = clock64() - tb_;  // back substitution
#endif
}

__global__ void __launch_bounds__(kLaThreads, 1) chol_la_kernel(const DevWindow* __restrict__ wp, int nt) {
    const DevWindow* w = wp + blockIdx.y;  // one window per grid row
    if (w->lm->phase != PH_TRIAL) return;  // uniform over the cluster: nobody reaches a cluster barrier
    extern __shared__ double smem[];
    CholJob J;
    J.n = w->n, J.ld = w->lds;
    J.A = w->S, J.y = w->bs, J.x = w->x, J.Lf = w->Lfac, J.scr = w->cminv;
    J.fail = &w->lm->chol_fail;
    J.nt = nt;
    J.dbg = w->dbg;
    chol_la_body(J, smem);
}

}  // namespace

// tiles per worker thread for a system of dimension n on a cluster of `cluster` CTAs
int chol_la_tiles_per_thread(int n, int cluster) {
    const int TR = (n + 4) >> 2, TC = (n + 3) >> 2;
    const int ntiles = TC * (2 * TR - TC + 1) / 2;
    const int per_round = 32 * kLaWorkerWarps * cluster;
    return (ntiles + per_round - 1) / per_round;
}
size_t chol_la_smem_bytes(int n, int cluster) {
    return sizeof(double) * (size_t)LaLayout(n, chol_la_tiles_per_thread(n, cluster), cluster).total;
}
size_t chol_la_scratch_doubles(int n) { return (size_t)((n + kNB - 1) / kNB) * kNB * kNB; }
bool chol_la_fits(int n, int cluster) { return n >= 1 && chol_la_smem_bytes(n, cluster) <= 227 * 1024 - 64; }

cudaError_t configure_chol_la() {
    cudaError_t e = opt_in_max_smem(chol_la_kernel);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(chol_la_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
}

cudaError_t launch_chol_la(cudaStream_t s, const DevWindow* wp, int n_windows, int cluster, int n_cap) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cluster, n_windows, 1);
    cfg.blockDim = dim3(kLaThreads, 1, 1);
    cfg.dynamicSmemBytes = chol_la_smem_bytes(n_cap, cluster);
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, chol_la_kernel, wp, chol_la_tiles_per_thread(n_cap, cluster));
}

}  // namespace vilba
