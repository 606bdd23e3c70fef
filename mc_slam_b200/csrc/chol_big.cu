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

// Code size matters in these kernels: every launch starts with a cold instruction cache and the unrolled register
// kernels (fg4_factor ~1.5 k instructions, a row solve ~0.9 k) are executed once or twice per launch by a few warps, so
// each of them is instantiated ONCE per kernel and reached through a loop (second pass: warm) or by several warps at once.

// diagonal tile k: 128 threads = one factor group; pass 0 = block A, pass 1 = block B
__global__ void __launch_bounds__(128) bigchol_diag_kernel(const DevWindow* __restrict__ wp, int k) {
    const DevWindow* w = wp + blockIdx.y;
    if (w->lm->phase != PH_TRIAL) return;
    const int n = w->n, ld = w->lds;
    if (k >= big_tiles(n)) return;
    const int k0 = k * kBT, nb = min(kBT, n - k0);
    const int jbA = min(32, nb), jbB = nb - jbA;
    __shared__ __align__(16) double Dt[2][1024], XdT[32 * kXS34], Fgs[kFgScratch], Scr[64 * 17], Sc[6 * 32];
    __shared__ int s_neg[2], s_fail;
    double* const A = w->S;
    double* rec = w->cminv + big_minv_doubles(n) + (size_t)k * kRecDoubles;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_fail = 0;
    double* const dis = Sc, *const dab = Sc + 64, *const sg = Sc + 128;  // [2][32] each: block A, block B
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        const int jb = pass ? jbB : jbA, o = k0 + 32 * pass;
        // ---- the block (pass 1: updated with the rows solved in pass 0, D' = D - X21 Sigma_A X21^T); lane = row,
        //      warp = column phase ----
        double a[2][4];
#pragma unroll
        for (int sl = 0; sl < 2; ++sl)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int c = 4 * (warp + 4 * sl) + i;
                a[sl][i] = (lane < jb && c <= lane) ? A[(size_t)(o + c) * ld + o + lane] : (c == lane ? 1.0 : 0.0);
            }
        if (pass) {
            double acc[2][4] = {{0.0, 0.0, 0.0, 0.0}, {0.0, 0.0, 0.0, 0.0}};
#pragma unroll 4
            for (int kk = 0; kk < 32; ++kk) {
                const double own = XdT[kk * kXS34 + lane] * sg[kk];
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
        const FgPivots p = fg4_factor(a, lane, warp, Dt[pass], Fgs, 1);
        if (!p.ok) s_fail = 1;
        if (warp == 0) {
            const double sgn = p.d < 0.0 ? -1.0 : 1.0;
            const double isq = 1.0 / sqrt(fabs(p.d));
            const unsigned anyneg = __ballot_sync(0xffffffffu, p.d < 0.0);
            dis[32 * pass + lane] = isq * sgn, dab[32 * pass + lane] = isq, sg[32 * pass + lane] = sgn;
            if (lane == 0) s_neg[pass] = anyneg != 0;
        }
        bar_grp(2);
        // ---- row solves against this block: pass 0: warp 0 = the rows of block B (-> XdT), warp 1 = the rows of the
        //      identity (-> inverse of block A for the back substitution).  The inverse of block B is left to the panel
        //      kernel: nothing in this kernel or the next waits for it ----
        if (pass == 0 && warp < 2) {
            const bool inv = warp == 1;
            double lo[16], hi[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                lo[c] = inv ? (c == lane ? 1.0 : 0.0) : ((lane < jbB) ? A[(size_t)(k0 + c) * ld + k0 + 32 + lane] : 0.0);
                hi[c] = inv ? (16 + c == lane ? 1.0 : 0.0) : ((lane < jbB) ? A[(size_t)(k0 + 16 + c) * ld + k0 + 32 + lane] : 0.0);
            }
            double* const my = Scr + tid * 17;
            const double* scale = inv ? dab : dis;
            double* out = inv ? w->cminv + (size_t)(2 * k) * 1024 + lane : XdT + lane;
            const int ostride = inv ? 32 : kXS34;  // Minv[c][j], row-major / XdT[c][row]
            rowsolve_lo(lo, Dt[0]);
#pragma unroll
            for (int c = 0; c < 16; ++c) my[c] = lo[c], out[c * ostride] = lo[c] * scale[c];
            rowsolve_hi(hi, my, Dt[0]);
#pragma unroll
            for (int c = 0; c < 16; ++c) out[(16 + c) * ostride] = hi[c] * scale[16 + c];
        }
        bar_grp(2);
    }
    // ---- the record for the panel kernel, the tile's own rows of the factor for the back substitution ----
#pragma unroll
    for (int i = 0; i < 8; ++i) rec[kRecDtA + tid + 128 * i] = Dt[0][tid + 128 * i], rec[kRecDtB + tid + 128 * i] = Dt[1][tid + 128 * i];
    for (int i = tid; i < 32 * kXS34; i += 128) rec[kRecX21 + i] = XdT[i];
    for (int i = tid; i < 6 * 32; i += 128) rec[kRecDisA + i] = Sc[i];
    if (tid < 2) rec[kRecNeg + tid] = (double)s_neg[tid];
    if (tid < nb) w->cdinv[k0 + tid] = sg[tid];  // signs of this tile's pivots (update kernel)
    for (int i = tid; i < 32 * 32; i += 128) {  // Lf(k0 + 32 + r, k0 + c) = X21(r, c), row-major
        const int r = i >> 5, c = i & 31;
        if (r < jbB) w->Lfac[(size_t)(k0 + 32 + r) * ld + k0 + c] = XdT[c * kXS34 + r];
    }
    if (tid == 0 && s_fail) w->lm->chol_fail = 1;
}

// rows [k0 + 64, n) of the panel and the rhs row: one thread per row, 64 rows per CTA; the last CTA of the grid solves
// the rows of the identity against block B instead (= its inverse, for the back substitution)
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
    const double* X21T = Rec + kRecX21;
    const double* dis = Rec + kRecDisA, *dab = dis + 64, *sg = dis + 128;  // [2][32] each (record layout: dis A, dis B, dab A, ...)
    const bool inv = blockIdx.x == gridDim.x - 1;
    const int r = blockIdx.x * kPanelThreads + tid;
    if (inv ? tid >= 32 : r >= rows) return;
    const bool is_rhs = !inv && (r == rows - 1);
    double* const A = w->S;
    double* my = Scr + tid * 33;
    auto src = [&](int c) { return c >= nb ? 0.0 : (is_rhs ? w->bs[k0 + c] : A[(size_t)(k0 + c) * ld + row0 + r]); };
    auto dst = [&](int c, double v) {
        if (c >= nb) return;
        if (is_rhs) {
            w->x[k0 + c] = v;  // forward-substituted right-hand side
        } else {
            A[(size_t)(k0 + c) * ld + row0 + r] = v;        // panel for the trailing update (column-major)
            w->Lfac[(size_t)(row0 + r) * ld + k0 + c] = v;  // row of the factor for the back substitution (row-major)
        }
    };
    // half 0: x1 = y1 M_A^-T Sigma_A;  half 1: y2 -= (x1 Sigma_A) X21^T, x2 = y2 M_B^-T Sigma_B   (one copy of the solve)
#pragma unroll 1
    for (int half = inv ? 1 : 0; half < 2; ++half) {
        if (half == 1 && nb <= 32) break;
        const double* Dt = Rec + (half ? kRecDtB : kRecDtA);
        double lo[16], hi[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            lo[c] = inv ? (c == tid ? 1.0 : 0.0) : src(32 * half + c);
            hi[c] = inv ? (16 + c == tid ? 1.0 : 0.0) : src(32 * half + 16 + c);
        }
        if (half == 1 && !inv) {
#pragma unroll 2
            for (int kk = 0; kk < 32; ++kk) {
                const double xk = my[kk];
                const double2* p = reinterpret_cast<const double2*>(X21T + kk * kXS34);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const double2 v = p[c], u = p[8 + c];
                    lo[2 * c] = fma(-xk, v.x, lo[2 * c]), lo[2 * c + 1] = fma(-xk, v.y, lo[2 * c + 1]);
                    hi[2 * c] = fma(-xk, u.x, hi[2 * c]), hi[2 * c + 1] = fma(-xk, u.y, hi[2 * c + 1]);
                }
            }
        }
        const double* scale = (inv ? dab : dis) + 32 * half;
        const double* sgh = sg + 32 * half;
        double x1s[16];
        rowsolve_lo(lo, Dt);
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            my[c] = lo[c];
            const double v = lo[c] * scale[c];
            x1s[c] = v * sgh[c];
            if (inv) w->cminv[(size_t)(2 * k + 1) * 1024 + c * 32 + tid] = v; else dst(32 * half + c, v);
        }
        rowsolve_hi(hi, my, Dt);
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const double v = hi[c] * scale[16 + c];
            if (inv) w->cminv[(size_t)(2 * k + 1) * 1024 + (16 + c) * 32 + tid] = v; else dst(32 * half + 16 + c, v);
            my[16 + c] = v * sgh[16 + c];  // x1 Sigma_A for the product of the second half
        }
#pragma unroll
        for (int c = 0; c < 16; ++c) my[c] = x1s[c];
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
