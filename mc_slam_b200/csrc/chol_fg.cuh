// chol_fg.cuh -- building blocks of the look-ahead cluster factorisation (chol_la.cu), written for LATENCY: the
// chain of dependent pivots is what bounds a small dense factorisation, so every block keeps its state in a handful
// of registers per thread (room for the compiler to run shared-memory loads ahead of the FP64 chain) and spreads the
// independent work over the four warps of a "factor group" (one warp per SM sub-partition).
//
// Thread layout of a 32 x 32 block inside a group of 128 threads: lane = row r, warp w owns the 4-column micro-blocks
// w and w + 4, i.e. columns [4w, 4w+4) and [16+4w, 16+4w+4); a thread holds those 8 entries of its row in a[2][4].
#pragma once
#include "chol_common.cuh"

namespace vilba {

constexpr int kFgNB = 32;

// scratch of fg4_factor in doubles: per-warp pivot blocks, double-buffered u / l publication, the pivots
constexpr int kFgScratch = 4 * 16 + 2 * (4 * 32 + 32 * 4) + 32;

struct FgPivots {
    double d;      // pivot of row `lane` (valid in every warp)
    bool ok;       // no zero / non-finite pivot among the pivot blocks THIS warp owned
};

// LDL^T of a 32 x 32 block distributed as described above (lower triangle; a[slot][i] = A(r, 4 (w + 4 slot) + i),
// entries above the diagonal are ignored and end up holding garbage).  Writes the unit-lower factor transposed,
// Dt[k * 32 + c] = l(c, k) for c > k and 0 for c <= k, and returns d_r.  One 128-thread named barrier (`bar_id`) per
// 4 columns; the 4x4 pivot blocks are factored redundantly by every lane of the owning warp (no per-column
// communication).
// One half of the eight pivot blocks (slot 0: micro-blocks 0..3, slot 1: micro-blocks 4..7; the slot is a compile-time
// constant because it indexes registers).  Fully unrolled: rolled, the four steps of a half measured 35 % slower (the
// scheduler no longer overlaps the tail of one pivot block with the head of the next).
template <int SLOT>
__device__ __forceinline__ void fg4_factor_half(double (&a)[2][4], int lane, int w, double* __restrict__ Dt,
                                                double* __restrict__ scratch, int bar_id, bool& ok) {
    double* Bc = scratch + 16 * w;             // this warp's 4 x 4 pivot block
    double* Ub = scratch + 64;                 // [2][4][32]  u(r, j..j+3), structure of arrays
    double* Lq = scratch + 64 + 2 * 128;       // [2][32][4]  l(c, j..j+3)
    double* dsh = scratch + 64 + 4 * 128;      // [32] pivots
    // trailing update of one of this thread's two micro-blocks with the u / l published for the current pivot block
    auto update_slot = [&](double (&as)[4], int c0, const double* ub, const double* lq) {
        const double u0 = ub[lane], u1 = ub[32 + lane], u2 = ub[64 + lane], u3 = ub[96 + lane];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const double2 la = reinterpret_cast<const double2*>(lq + 4 * (c0 + i))[0];
            const double2 lb = reinterpret_cast<const double2*>(lq + 4 * (c0 + i))[1];
            as[i] -= (u0 * la.x + u1 * la.y) + (u2 * lb.x + u3 * lb.y);
        }
    };
#pragma unroll
    for (int s4 = 0; s4 < 4; ++s4) {
        const int s = 4 * SLOT + s4;
        const int j = 4 * s;
        double* ub = Ub + (s & 1) * 128;
        double* lq = Lq + (s & 1) * 128;
        if (w == s4) {
            // ---- owner warp: factor the pivot block, scale its rows, publish ----
            if (lane >= j && lane < j + 4) {
                double* p = Bc + 4 * (lane - j);
                reinterpret_cast<double2*>(p)[0] = make_double2(a[SLOT][0], a[SLOT][1]);
                reinterpret_cast<double2*>(p)[1] = make_double2(a[SLOT][2], a[SLOT][3]);
            }
            __syncwarp();
            const double b00 = Bc[0], b10 = Bc[4], b20 = Bc[8], b30 = Bc[12];
            double b11 = Bc[5], b21 = Bc[9], b31 = Bc[13], b22 = Bc[10], b32 = Bc[14], b33 = Bc[15];
            const double d0 = b00, r0 = fast_rcp(d0);
            const double l10 = b10 * r0, l20 = b20 * r0, l30 = b30 * r0;
            b11 -= l10 * b10, b21 -= l20 * b10, b31 -= l30 * b10;
            b22 -= l20 * b20, b32 -= l30 * b20, b33 -= l30 * b30;
            const double d1 = b11, r1 = fast_rcp(d1);
            const double l21 = b21 * r1, l31 = b31 * r1;
            b22 -= l21 * b21, b32 -= l31 * b21, b33 -= l31 * b31;
            const double d2 = b22, r2 = fast_rcp(d2);
            const double l32 = b32 * r2;
            b33 -= l32 * b32;
            const double d3 = b33;
            const double r3 = fast_rcp(d3);
            if (pivot_bad(d0) || pivot_bad(d1) || pivot_bad(d2) || pivot_bad(d3)) ok = false;
            // u = A(r, j..j+3) Lb^-T (unscaled), l = u D^-1.  Branch free: for the rows of the pivot block itself the
            // same formulas give l(r, k) for k < r - j and garbage for k >= r - j, which nobody reads (the stores into Dt
            // mask it; Ub / Lq of rows <= j+3 only ever feed entries above the diagonal, which are ignored)
            const double u0 = a[SLOT][0];
            const double u1 = fma(-u0, l10, a[SLOT][1]);
            const double u2 = fma(-u1, l21, fma(-u0, l20, a[SLOT][2]));
            const double u3 = fma(-u2, l32, fma(-u1, l31, fma(-u0, l30, a[SLOT][3])));
            const double q0 = u0 * r0, q1 = u1 * r1, q2 = u2 * r2, q3 = u3 * r3;
            ub[lane] = u0, ub[32 + lane] = u1, ub[64 + lane] = u2, ub[96 + lane] = u3;
            reinterpret_cast<double2*>(lq + 4 * lane)[0] = make_double2(q0, q1);
            reinterpret_cast<double2*>(lq + 4 * lane)[1] = make_double2(q2, q3);
            if (lane == 0) {
                reinterpret_cast<double2*>(dsh + j)[0] = make_double2(d0, d1);
                reinterpret_cast<double2*>(dsh + j)[1] = make_double2(d2, d3);
            }
            // the factor itself: l(lane, j + i) -> Dt[(j + i) * 32 + lane]; rows above the column get 0
            Dt[(j + 0) * 32 + lane] = (lane > j + 0) ? q0 : 0.0;
            Dt[(j + 1) * 32 + lane] = (lane > j + 1) ? q1 : 0.0;
            Dt[(j + 2) * 32 + lane] = (lane > j + 2) ? q2 : 0.0;
            Dt[(j + 3) * 32 + lane] = (lane > j + 3) ? q3 : 0.0;
        }
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        // ---- trailing update of this warp's columns right of the pivot block (micro-block of slot sl is w + 4 sl):
        //      the next owner's pivot block first ----
        if (SLOT == 0) {
            if (w > s4) update_slot(a[0], 4 * w, ub, lq);
            update_slot(a[1], 4 * (w + 4), ub, lq);
        } else {
            if (w > s4) update_slot(a[1], 4 * (w + 4), ub, lq);
        }
    }
}

// LDL^T of a 32 x 32 block distributed as described above (lower triangle; a[slot][i] = A(r, 4 (w + 4 slot) + i),
// entries above the diagonal are ignored and end up holding garbage).  Writes the unit-lower factor transposed,
// Dt[k * 32 + c] = l(c, k) for c > k and 0 for c <= k, and returns d_r.  One 128-thread named barrier (`bar_id`) per
// 4 columns; the 4x4 pivot blocks are factored redundantly by every lane of the owning warp (no per-column
// communication).
__device__ __forceinline__ FgPivots fg4_factor(double (&a)[2][4], int lane, int w, double* __restrict__ Dt,
                                               double* __restrict__ scratch, int bar_id) {
    FgPivots out;
    out.ok = true;
    fg4_factor_half<0>(a, lane, w, Dt, scratch, bar_id, out.ok);
    fg4_factor_half<1>(a, lane, w, Dt, scratch, bar_id, out.ok);
    out.d = scratch[64 + 4 * 128 + lane];  // behind the barrier of the last pivot block
    return out;
}

// One panel row per thread, in two halves of 16 columns: z L11^T = a with unit-lower L11 given as Dt[k*32 + c] = l(c,k).
// rowsolve_lo solves columns 0..15 in place; rowsolve_hi solves columns 16..31 given the (unscaled) solution of the
// first half in memory.  Only 16 values are live at any time, so inside the big kernel (128-register cap, ~45
// registers of loop state) the register file still has room for the loads of the factor to run ahead of the fma chain
// -- with a whole 32-entry row per thread ptxas serialises every shared-memory load behind the previous pair of fma
// (measured 9 k cycles per row instead of 2 k).
__device__ __forceinline__ void rowsolve_lo(double (&lo)[16], const double* __restrict__ Dt) {
#pragma unroll
    for (int k = 0; k < 15; ++k) {
#pragma unroll
        for (int c = k + 1; c < 16; ++c) lo[c] = fma(-lo[k], Dt[k * 32 + c], lo[c]);
    }
}
__device__ __forceinline__ void rowsolve_hi(double (&hi)[16], const double* __restrict__ xlo, const double* __restrict__ Dt) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const double xk = xlo[k];
#pragma unroll
        for (int c = 0; c < 16; ++c) hi[c] = fma(-xk, Dt[k * 32 + 16 + c], hi[c]);
    }
#pragma unroll
    for (int k = 0; k < 15; ++k) {
#pragma unroll
        for (int c = k + 1; c < 16; ++c) hi[c] = fma(-hi[k], Dt[(16 + k) * 32 + 16 + c], hi[c]);
    }
}

// Back substitution M^T x = yf for a factor stored as (a) the rows below the 32-wide diagonal blocks, row-major in Lf
// (Lf[r * ld + c] = M(r, c)), and (b) the inverse of every diagonal block (Minv[blk][r * 32 + c], row-major), block by
// block from the last one:  x_k = Minv_k^T (yf_k - sum over the rows r below the block of M(r, block)^T x_r).
// One CTA of NT threads: its warps split the rows below (lane = column of the block: 256-byte coalesced row segments,
// the first 16 rows per warp fetched one block ahead so that L2 latency is off the chain), their partial sums meet in
// shared memory, warp 0 applies the inverse -- no per-unknown chain.  smem: backsub_smem_doubles(n, NT) doubles.
__host__ __device__ constexpr int backsub_smem_doubles(int n, int nt) { return 2 * 32 * 32 + (nt / 32 + 1) * 32 + n + 32; }

template <int NT>
__device__ __forceinline__ void backsub_blocks(int n, int ld, const double* __restrict__ Lf, const double* __restrict__ Minv_g,
                                               const double* __restrict__ yf, double* __restrict__ xout, double* smem) {
    constexpr int NB = 32, NW = NT / 32, MV = (NB * NB / 2 + NT - 1) / NT;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nblk = (n + NB - 1) / NB;
    double* Mi = smem;                  // [2][NB*NB] inverse of the current / next diagonal block
    double* part = Mi + 2 * NB * NB;    // [NW][NB] partial sums
    double* ts = part + NW * NB;        // [NB]
    double* xs = ts + NB;               // [n + NB]
    auto fetch_rows = [&](int blk, double (&v)[16]) {
        const int c = blk * NB + lane;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int r = (blk + 1) * NB + warp + NW * i;
            v[i] = (r < n && c < n) ? __ldcg(Lf + (size_t)r * ld + c) : 0.0;
        }
    };
    double cur[16], nxt[16];
    double2 mnext[MV];
    fetch_rows(nblk - 1, cur);
    {
        const double2* m = reinterpret_cast<const double2*>(Minv_g + (size_t)(nblk - 1) * NB * NB);
        double2* dst = reinterpret_cast<double2*>(Mi + ((nblk - 1) & 1) * NB * NB);
        for (int i = tid; i < NB * NB / 2; i += NT) dst[i] = __ldcg(m + i);
    }
    __syncthreads();
    for (int blk = nblk - 1; blk >= 0; --blk) {
        const int j0 = blk * NB, jb = min(NB, n - j0);
        if (blk > 0) {  // next block's rows and inverse: in flight during this block's reduction
            fetch_rows(blk - 1, nxt);
            const double2* m = reinterpret_cast<const double2*>(Minv_g + (size_t)(blk - 1) * NB * NB);
#pragma unroll
            for (int i = 0; i < MV; ++i)
                if (tid + NT * i < NB * NB / 2) mnext[i] = __ldcg(m + tid + NT * i);
        }
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
            const int r = j0 + NB + warp + NW * i;
            s0 = fma(cur[i], (r < n) ? xs[r] : 0.0, s0);
            s1 = fma(cur[i + 1], (r + NW < n) ? xs[r + NW] : 0.0, s1);
        }
        if (j0 + lane < n) {  // systems with more than 16 NW rows below a block: the rest straight from L2, 16 in flight
            for (int r = j0 + NB + warp + NW * 16; r < n; r += 16 * NW) {
                double l[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) l[i] = (r + i * NW < n) ? __ldcg(Lf + (size_t)(r + i * NW) * ld + j0 + lane) : 0.0;
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    s0 = fma(l[i], (r + i * NW < n) ? xs[r + i * NW] : 0.0, s0);
                    s1 = fma(l[i + 1], (r + (i + 1) * NW < n) ? xs[r + (i + 1) * NW] : 0.0, s1);
                }
            }
        }
        part[warp * NB + lane] = s0 + s1;
        __syncthreads();
        if (warp == 0) {
            double t = (lane < jb) ? __ldcg(yf + j0 + lane) : 0.0;
#pragma unroll
            for (int w2 = 0; w2 < NW; ++w2) t -= part[w2 * NB + lane];
            ts[lane] = t;
            __syncwarp();
            const double* m = Mi + (blk & 1) * NB * NB;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
            for (int r = 0; r < NB; r += 4) {
                a0 = fma(m[r * NB + lane], ts[r], a0);
                a1 = fma(m[(r + 1) * NB + lane], ts[r + 1], a1);
                a2 = fma(m[(r + 2) * NB + lane], ts[r + 2], a2);
                a3 = fma(m[(r + 3) * NB + lane], ts[r + 3], a3);
            }
            xs[j0 + lane] = (lane < jb) ? (a0 + a1) + (a2 + a3) : 0.0;
        }
        if (blk > 0) {
            double2* dst = reinterpret_cast<double2*>(Mi + ((blk - 1) & 1) * NB * NB);
#pragma unroll
            for (int i = 0; i < MV; ++i)
                if (tid + NT * i < NB * NB / 2) dst[tid + NT * i] = mnext[i];
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 16; ++i) cur[i] = nxt[i];
    }
    for (int i = tid; i < n; i += NT) xout[i] = xs[i];
}

}  // namespace vilba
