// chol_big.cu -- reduced camera system of LARGE windows (n > 480, e.g. BASELINE config 4: n = 1485): blocked
// right-looking LDL^T over the whole GPU instead of one thread-block cluster.
//
// Replaces LinearSolverEigen::solve (g2o/solvers/linear_solver_eigen.h:94-124: SimplicialLDLT factor + two
// triangular solves) for the dense symmetric matrix S the Schur step leaves in DevWindow::S (lower triangle,
// column-major: A(i,j), i >= j, at S[j * lds + i]).  Same failure rule as SimplicialLDLT and chol_la.cu: negative
// pivots are factored through (S = M Sigma M^T, Sigma = diag(sign d)), only a zero / non-finite pivot raises
// LmState::chol_fail (the LM controller then rejects the trial, optimization_algorithm_levenberg.cpp:126-127).
// Per 64-column step k:
//     diag    : one CTA (a 4-warp factor group, chol_fg.cuh) factors the 64 x 64 diagonal tile as two 32 x 32 blocks:
//               LDL^T of block A, the 32 rows of block B against it, LDL^T of the updated block B; it leaves the two
//               unit-lower factors, the scalings / signs and the inverse of block A in a per-step scratch record
//     panel   : one THREAD per row of the panel (rows below the tile + the right-hand side as one more row = free
//               forward substitution): x1 = y1 M_A^-T Sigma_A, y2 -= x1 Sigma_A X21^T, x2 = y2 M_B^-T Sigma_B in
//               registers, 16 values at a time (chol_fg.cuh); the rows go back to S (for the update) and, row-major,
//               to Lfac (for the back substitution); one more CTA inverts block B meanwhile
//     update  : one CTA per 64x64 trailing tile, C_IJ -= P_I Sigma P_J^T (register-tiled FP64 GEMM), the diagonal-tile
//               CTAs also update the right-hand side; block column k+1 first (look-ahead), the rest on a side stream
// followed by one CTA that back-substitutes with the inverted 32 x 32 diagonal blocks (backsub_blocks).  Every kernel
// returns at once unless its window is in the TRIAL phase.
#include <algorithm>
#include <cstdlib>

#include "chol_fg.cuh"
#include "lba_common.cuh"

namespace vilba {

constexpr int kBT = 64;        // tile edge
constexpr int kBTP = kBT + 1;  // padded shared-memory stride

__device__ __forceinline__ int big_tiles(int n) { return (n + kBT - 1) / kBT; }

// ---- per-step scratch record (in DevWindow::cminv, behind the inverses of the 32-wide diagonal blocks) ----
constexpr int kRecDtA = 0;                 // [32*32] unit-lower factor of block A, transposed: Dt[k*32 + c] = l(c,k)
constexpr int kRecDtB = 1024;              // [32*32] the same for block B
constexpr int kRecX21 = 2048;              // [32][34] rows of block B against block A, transposed: X21T[k*34 + r] = X21(r,k)
constexpr int kRecDisA = 2048 + 32 * 34;   // [32] signed |d|^-1/2 of block A, then: dis B, dab A, dab B, sg A, sg B
constexpr int kRecNeg = kRecDisA + 6 * 32; // [2] a negative pivot in block A / B
constexpr int kRecDoubles = kRecNeg + 8;
constexpr int kXS34 = 34;

__host__ __device__ inline size_t big_minv_doubles(int n) { return (size_t)((n + 31) / 32) * 1024; }
size_t chol_big_scratch_doubles(int n) { return big_minv_doubles(n) + (size_t)((n + kBT - 1) / kBT) * kRecDoubles; }

__device__ __forceinline__ void bar_grp(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

// One warp: inverse of a 32 x 32 diagonal factor block M = Lu |D|^1/2 (unit-lower Lu given transposed in Dt) into
// out[r * 32 + c], row-major: a row e_j of the identity solved against Lu^T is column j of Lu^-1.
__device__ __forceinline__ void invert_block(const double* __restrict__ Dt, const double* __restrict__ dab, double* __restrict__ out,
                                             double* scr /* 32 * 17 */, int lane) {
    double lo[16], hi[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) lo[c] = (c == lane) ? 1.0 : 0.0, hi[c] = (16 + c == lane) ? 1.0 : 0.0;
    rowsolve_lo(lo, Dt);
    double* my = scr + lane * 17;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        my[c] = lo[c];
        out[c * 32 + lane] = lo[c] * dab[c];
    }
    rowsolve_hi(hi, my, Dt);
#pragma unroll
    for (int c = 0; c < 16; ++c) out[(16 + c) * 32 + lane] = hi[c] * dab[16 + c];
}

// diagonal tile k: 128 threads = one factor group
__global__ void __launch_bounds__(128) bigchol_diag_kernel(const DevWindow* __restrict__ wp, int k) {
    const DevWindow* w = wp + blockIdx.y;
    if (w->lm->phase != PH_TRIAL) return;
    const int n = w->n, ld = w->lds;
    if (k >= big_tiles(n)) return;
    const int k0 = k * kBT, nb = min(kBT, n - k0);
    const int jbA = min(32, nb), jbB = nb - jbA;
    __shared__ __align__(16) double DtA[1024], DtB[1024], XdT[32 * kXS34], Fgs[kFgScratch], Scr[32 * 17], Sc[6 * 32];
    __shared__ int s_neg[2], s_fail;
    double* const A = w->S;
    double* rec = w->cminv + big_minv_doubles(n) + (size_t)k * kRecDoubles;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_fail = 0;
    double* disA = Sc, *disB = Sc + 32, *dabA = Sc + 64, *dabB = Sc + 96, *sgA = Sc + 128, *sgB = Sc + 160;
    auto publish = [&](const FgPivots& p, double* dis, double* dab, double* sg, int which) {
        if (!p.ok) s_fail = 1;
        if (warp == 0) {
            const double sgn = p.d < 0.0 ? -1.0 : 1.0;
            const double isq = 1.0 / sqrt(fabs(p.d));
            const unsigned anyneg = __ballot_sync(0xffffffffu, p.d < 0.0);
            dis[lane] = isq * sgn, dab[lane] = isq, sg[lane] = sgn;
            if (lane == 0) s_neg[which] = anyneg != 0;
        }
    };
    // ---- block A ----
    double a[2][4];
#pragma unroll
    for (int sl = 0; sl < 2; ++sl)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = 4 * (warp + 4 * sl) + i;
            a[sl][i] = (lane < jbA && c <= lane) ? A[(size_t)(k0 + c) * ld + k0 + lane] : (c == lane ? 1.0 : 0.0);
        }
    publish(fg4_factor(a, lane, warp, DtA, Fgs, 1), disA, dabA, sgA, 0);
    bar_grp(2);
    // ---- rows of block B against block A (warp 0), inverse of block A for the back substitution (warp 1) ----
    if (warp == 0) {
        double lo[16], hi[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            lo[c] = (lane < jbB) ? A[(size_t)(k0 + c) * ld + k0 + 32 + lane] : 0.0;
            hi[c] = (lane < jbB) ? A[(size_t)(k0 + 16 + c) * ld + k0 + 32 + lane] : 0.0;
        }
        double* const my = Fgs + lane * 17;  // fg4_factor's scratch is idle here
        rowsolve_lo(lo, DtA);
#pragma unroll
        for (int c = 0; c < 16; ++c) my[c] = lo[c], XdT[c * kXS34 + lane] = lo[c] * disA[c];
        rowsolve_hi(hi, my, DtA);
#pragma unroll
        for (int c = 0; c < 16; ++c) XdT[(16 + c) * kXS34 + lane] = hi[c] * disA[16 + c];
    } else if (warp == 1) {
        invert_block(DtA, dabA, w->cminv + (size_t)(2 * k) * 1024, Scr, lane);
    }
    bar_grp(2);
    // ---- block B, updated with the rows just solved: D' = D - X21 Sigma_A X21^T (lane = row, warp = column phase) ----
#pragma unroll
    for (int sl = 0; sl < 2; ++sl)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = 4 * (warp + 4 * sl) + i;
            a[sl][i] = (lane < jbB && c <= lane) ? A[(size_t)(k0 + 32 + c) * ld + k0 + 32 + lane] : (c == lane ? 1.0 : 0.0);
        }
    {
        double acc[2][4] = {{0.0, 0.0, 0.0, 0.0}, {0.0, 0.0, 0.0, 0.0}};
#pragma unroll 8
        for (int kk = 0; kk < 32; ++kk) {
            const double own = XdT[kk * kXS34 + lane] * sgA[kk];
            const double2* cp0 = reinterpret_cast<const double2*>(XdT + kk * kXS34 + 4 * warp);
            const double2* cp1 = reinterpret_cast<const double2*>(XdT + kk * kXS34 + 16 + 4 * warp);
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
    publish(fg4_factor(a, lane, warp, DtB, Fgs, 1), disB, dabB, sgB, 1);
    bar_grp(2);
    // ---- the record for the panel kernel, the tile's own rows of the factor for the back substitution ----
#pragma unroll
    for (int i = 0; i < 8; ++i) rec[kRecDtA + tid + 128 * i] = DtA[tid + 128 * i], rec[kRecDtB + tid + 128 * i] = DtB[tid + 128 * i];
    for (int i = tid; i < 32 * kXS34; i += 128) rec[kRecX21 + i] = XdT[i];
    for (int i = tid; i < 6 * 32; i += 128) rec[kRecDisA + i] = Sc[i];
    if (tid < 2) rec[kRecNeg + tid] = (double)s_neg[tid];
    if (tid < nb) w->cdinv[k0 + tid] = Sc[128 + tid];  // signs of this tile's pivots (update kernel)
    for (int i = tid; i < 32 * 32; i += 128) {  // Lf(k0 + 32 + r, k0 + c) = X21(r, c), row-major
        const int r = i >> 5, c = i & 31;
        if (r < jbB) w->Lfac[(size_t)(k0 + 32 + r) * ld + k0 + c] = XdT[c * kXS34 + r];
    }
    if (tid == 0 && s_fail) w->lm->chol_fail = 1;
}

// rows [k0 + 64, n) of the panel and the rhs row: one thread per row, 64 rows per CTA; the last CTA of the grid inverts
// block B of the diagonal tile instead (for the back substitution)
constexpr int kPanelThreads = 64;
__global__ void __launch_bounds__(kPanelThreads) bigchol_panel_kernel(const DevWindow* __restrict__ wp, int k) {
    const DevWindow* w = wp + blockIdx.y;
    if (w->lm->phase != PH_TRIAL) return;
    const int n = w->n, ld = w->lds;
    if (k >= big_tiles(n)) return;
    const int k0 = k * kBT, nb = min(kBT, n - k0);
    const int row0 = k0 + kBT;              // first panel row (may be >= n: then only the rhs row is left)
    const int rows = max(0, n - row0) + 1;  // + rhs
    __shared__ __align__(16) double Rec[kRecDoubles];
    __shared__ double Scr[kPanelThreads * 33];
    const double* rec = w->cminv + big_minv_doubles(n) + (size_t)k * kRecDoubles;
    const int tid = threadIdx.x;
    for (int i = tid; i < kRecDoubles; i += kPanelThreads) Rec[i] = rec[i];
    __syncthreads();
    const double* DtA = Rec + kRecDtA, *DtB = Rec + kRecDtB, *X21T = Rec + kRecX21;
    const double* disA = Rec + kRecDisA, *disB = disA + 32, *dabB = disA + 96, *sgA = disA + 128;
    if (blockIdx.x == gridDim.x - 1) {
        if (tid < 32) invert_block(DtB, dabB, w->cminv + (size_t)(2 * k + 1) * 1024, Scr, tid);
        return;
    }
    const int r = blockIdx.x * kPanelThreads + tid;
    if (r >= rows) return;
    const bool is_rhs = (r == rows - 1);
    double* const A = w->S;
    double* my = Scr + tid * 33;
    auto src = [&](int c) { return is_rhs ? w->bs[k0 + c] : A[(size_t)(k0 + c) * ld + row0 + r]; };
    auto dst = [&](int c, double v) {
        if (c >= nb) return;
        if (is_rhs) {
            w->x[k0 + c] = v;  // forward-substituted right-hand side
        } else {
            A[(size_t)(k0 + c) * ld + row0 + r] = v;        // panel for the trailing update (column-major)
            w->Lfac[(size_t)(row0 + r) * ld + k0 + c] = v;  // row of the factor for the back substitution (row-major)
        }
    };
    double lo[16], hi[16];
    // ---- x1 = y1 M_A^-T Sigma_A ----
#pragma unroll
    for (int c = 0; c < 16; ++c) lo[c] = (c < nb) ? src(c) : 0.0, hi[c] = (16 + c < nb) ? src(16 + c) : 0.0;
    rowsolve_lo(lo, DtA);
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        my[c] = lo[c];
        dst(c, lo[c] * disA[c]);
    }
    rowsolve_hi(hi, my, DtA);
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        const double v = hi[c] * disA[16 + c];
        dst(16 + c, v);
        my[16 + c] = v * sgA[16 + c];  // x1 Sigma_A for the product below
    }
#pragma unroll
    for (int c = 0; c < 16; ++c) my[c] = my[c] * disA[c] * sgA[c];
    if (nb <= 32) return;
    // ---- y2 -= (x1 Sigma_A) X21^T, 16 columns at a time; x2 = y2 M_B^-T Sigma_B ----
#pragma unroll
    for (int c = 0; c < 16; ++c) lo[c] = (32 + c < nb) ? src(32 + c) : 0.0, hi[c] = (48 + c < nb) ? src(48 + c) : 0.0;
#pragma unroll 4
    for (int kk = 0; kk < 32; ++kk) {
        const double xk = my[kk];
        const double2* p = reinterpret_cast<const double2*>(X21T + kk * kXS34);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const double2 v = p[c];
            lo[2 * c] = fma(-xk, v.x, lo[2 * c]), lo[2 * c + 1] = fma(-xk, v.y, lo[2 * c + 1]);
        }
    }
#pragma unroll 4
    for (int kk = 0; kk < 32; ++kk) {
        const double xk = my[kk];
        const double2* p = reinterpret_cast<const double2*>(X21T + kk * kXS34 + 16);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const double2 v = p[c];
            hi[2 * c] = fma(-xk, v.x, hi[2 * c]), hi[2 * c + 1] = fma(-xk, v.y, hi[2 * c + 1]);
        }
    }
    rowsolve_lo(lo, DtB);
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        my[c] = lo[c];
        dst(32 + c, lo[c] * disB[c]);
    }
    rowsolve_hi(hi, my, DtB);
#pragma unroll
    for (int c = 0; c < 16; ++c) dst(48 + c, hi[c] * disB[16 + c]);
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
    // S = M Sigma M^T: a negative pivot in this block column (indefinite S) puts its sign on the J operand
    const double* rec = w.cminv + big_minv_doubles(n) + (size_t)k * kRecDoubles;
    const bool neg = rec[kRecNeg] != 0.0 || rec[kRecNeg + 1] != 0.0;
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
            PJ[i][c] = (i < nj) ? A[(size_t)(k0 + c) * ld + J0 + i] * (neg ? w.cdinv[k0 + c] : 1.0) : 0.0;
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

// M^T x = y with the inverted 32-wide diagonal blocks, one CTA (chol_fg.cuh)
constexpr int kBackThreads = 512;
__global__ void __launch_bounds__(kBackThreads) bigchol_backsub_kernel(const DevWindow* __restrict__ wp) {
    const DevWindow* w = wp + blockIdx.y;
    if (w->lm->phase != PH_TRIAL) return;
    extern __shared__ double back_sm[];
    backsub_blocks<kBackThreads>(w->n, w->lds, w->Lfac, w->cminv, w->x, w->x, back_sm);
}

size_t chol_big_backsub_smem(int n_cap) { return sizeof(double) * (size_t)backsub_smem_doubles(n_cap, kBackThreads); }

constexpr size_t kUpdateSmem = sizeof(double) * 2 * kBT * kBTP;

cudaError_t configure_chol_big(int n_cap) {
    (void)n_cap;
    cudaError_t e = opt_in_max_smem(bigchol_update_kernel);
    if (e != cudaSuccess) return e;
    return opt_in_max_smem(bigchol_backsub_kernel);
}

cudaError_t launch_chol_big(cudaStream_t s, cudaStream_t side, cudaEvent_t ev_trsm, cudaEvent_t ev_rest, const DevWindow* wp,
                            const LaunchDims& d) {
    const int ntile = d.chol_big_tiles;
    static const bool lookahead = !(std::getenv("VILBA_CHOL_LOOKAHEAD") && std::atoi(std::getenv("VILBA_CHOL_LOOKAHEAD")) == 0);
    cudaError_t e;
    bool rest_pending = false;
    for (int k = 0; k < ntile; ++k) {
        bigchol_diag_kernel<<<dim3(1, d.n_windows), 128, 0, s>>>(wp, k);
        const int rows = (ntile - k - 1) * kBT + 1;
        bigchol_panel_kernel<<<dim3((rows + kPanelThreads - 1) / kPanelThreads + 1, d.n_windows), kPanelThreads, 0, s>>>(wp, k);
        const int T = ntile - k - 1;
        if (T <= 0) continue;
        if (!lookahead) {
            bigchol_update_kernel<<<dim3(std::min(T * (T + 1) / 2, 4 * d.sm_count), d.n_windows), 256, kUpdateSmem, s>>>(wp, k, 0);
            continue;
        }
        // look-ahead: block column k + 1 on the main stream (the next diag / panel kernels wait only for it), the rest of
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
    bigchol_backsub_kernel<<<dim3(1, d.n_windows), kBackThreads, chol_big_backsub_smem(d.chol_big_tiles * kBT), s>>>(wp);
    return cudaGetLastError();
}

}  // namespace vilba
