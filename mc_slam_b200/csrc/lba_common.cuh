// lba_common.cuh -- device helpers shared by the local-BA kernels: shared-memory staging of the
// key-frame states and the residual functions of the three edge types.
#pragma once
#include <cfloat>

#include "kernels.h"
#include "vmath.cuh"

namespace vilba {

// ------------------------------------------------------------------------------------------------
// shared-memory stage of the key-frame states
// ------------------------------------------------------------------------------------------------
struct KfSmem {
    double* cam;   // K * 12 : R_cw (9, row-major) | P_wb (3)
    double* full;  // K * 22 : NavState
    int* blk;      // K
};

__device__ __forceinline__ KfSmem kf_smem_carve(double* base, int K) {
    KfSmem s;
    s.cam = base;
    s.full = base + 12 * (size_t)K;
    s.blk = reinterpret_cast<int*>(s.full + 22 * (size_t)K);
    return s;
}
__host__ __device__ static inline size_t kf_smem_bytes(int K) { return sizeof(double) * 34 * (size_t)K + sizeof(int) * (size_t)K + 16; }

// Loads (and for APPLY first updates: VertexNavStatePVR/Bias::oplusImpl, g2otypes.h:505-546,
// NavState.cpp:81-109) the key-frame states.  CTA 0 publishes the updated states to the trial buffer.
template <bool APPLY>
__device__ __forceinline__ void kf_stage(const DevWindow& w, const KfSmem& s, int cur) {
    const double* src = (cur ? w.kf_state[1] : w.kf_state[0]);
    double* dst = (cur ? w.kf_state[0] : w.kf_state[1]);
    const M3 Rcb = ldm3(w.Rcb);
    for (int k = threadIdx.x; k < w.K; k += blockDim.x) {
        const double* p = src + 22 * (size_t)k;
        double st[22];
#pragma unroll
        for (int i = 0; i < 22; ++i) st[i] = p[i];
        const int blk = w.kf_block[k];
        if (APPLY && blk >= 0) {
            const double* x = w.x + 15 * (size_t)blk;
            st[0] += x[0], st[1] += x[1], st[2] += x[2];
            st[3] += x[3], st[4] += x[4], st[5] += x[5];
            Q4 q = so3_mul(Q4{st[6], st[7], st[8], st[9]}, so3_exp(v3(x[6], x[7], x[8])));
            st[6] = q.w, st[7] = q.x, st[8] = q.y, st[9] = q.z;
#pragma unroll
            for (int i = 0; i < 6; ++i) st[16 + i] += x[9 + i];
        }
        if (APPLY && blockIdx.x == 0) {
#pragma unroll
            for (int i = 0; i < 22; ++i) dst[22 * (size_t)k + i] = st[i];
        }
#pragma unroll
        for (int i = 0; i < 22; ++i) s.full[22 * k + i] = st[i];
        const M3 Rwb = q_to_matrix(Q4{st[6], st[7], st[8], st[9]});
        const M3 Rcw = Rcb * transpose(Rwb);
        stm3(s.cam + 12 * k, Rcw);
        s.cam[12 * k + 9] = st[0], s.cam[12 * k + 10] = st[1], s.cam[12 * k + 11] = st[2];
        s.blk[k] = blk;
    }
}

// ------------------------------------------------------------------------------------------------
// edge math
// ------------------------------------------------------------------------------------------------
struct MonoObs {
    float u, v, is2;
    int kf;
    bool culled, robust;
};
__device__ __forceinline__ MonoObs load_obs(const int4* obs, int e) {
    const int4 r = obs[e];
    MonoObs o;
    o.u = __int_as_float(r.x);
    o.v = __int_as_float(r.y);
    o.is2 = __int_as_float(r.z);
    o.kf = r.w & OBS_KF_MASK;
    o.culled = (r.w & OBS_CULLED) != 0;
    o.robust = (r.w & OBS_ROBUST) != 0;
    return o;
}

// RobustKernelHuber::robustify (robust_kernel_impl.cpp:78-91); delta^2 lives in a float member of the reference's kernel
// (robust_kernel_impl.h:84), so it is rounded to single precision here too
__device__ __forceinline__ void huber(double e, double delta, double& rho0, double& rho1) {
    const double dsqr = (double)(float)(delta * delta);
    if (e <= dsqr) {
        rho0 = e;
        rho1 = 1.0;
    } else {
        const double sqrte = sqrt(e);
        rho0 = 2 * sqrte * delta - dsqr;
        rho1 = delta / sqrte;
    }
}

// EdgeNavStatePVRPointXYZ::computeError (g2otypes.h:636-684)
__device__ __forceinline__ void mono_error(const DevWindow& w, const double* cam, V3 Pw, const MonoObs& o,
                                           double& e0, double& e1, V3& Paux, V3& Pc) {
    const M3 Rcw = ldm3(cam);
    Paux = Rcw * (Pw - ld3(cam + 9));
    Pc = Paux + ld3(w.tcb);
    // one IEEE reciprocal instead of the reference's two divisions (x / z, y / z): the FP64 pipe is the limiter
    // of the per-edge kernels and a division costs ~20 instructions; the result moves by <= 1 ulp
    const double iz = 1.0 / Pc.z;
    const double px = Pc.x * iz, py = Pc.y * iz;
    e0 = (double)o.u - (px * w.fx + w.cx);
    e1 = (double)o.v - (py * w.fy + w.cy);
}

// EdgeNavStatePVR::computeError (g2otypes.cpp:529-585); returns [rP, rV, rPhi]
__device__ __forceinline__ void pvr_error(const DevWindow& w, const double* si, const double* sj,
                                          const double* M, V3& rP, V3& rV, V3& rPhi) {
    const V3 Pi = ld3(si), Vi = ld3(si + 3), Pj = ld3(sj), Vj = ld3(sj + 3);
    const Q4 Ri = q_normalized(Q4{si[6], si[7], si[8], si[9]});  // Get_R() copies => renormalises
    const Q4 Rj = q_normalized(Q4{sj[6], sj[7], sj[8], sj[9]});
    const V3 dbg = ld3(si + 16), dba = ld3(si + 19);
    const V3 g = ld3(w.g);
    const double T = M[VILBA_PI_DT], T2 = T * T;
    const Q4 RiT = so3_inverse(Ri);
    rP = q_rotate(RiT, Pj - Pi - Vi * T - (0.5 * g) * T2) -
         (ld3(M + VILBA_PI_DP) + ldm3(M + VILBA_PI_JPG) * dbg + ldm3(M + VILBA_PI_JPA) * dba);
    rV = q_rotate(RiT, Vj - Vi - g * T) -
         (ld3(M + VILBA_PI_DV) + ldm3(M + VILBA_PI_JVG) * dbg + ldm3(M + VILBA_PI_JVA) * dba);
    const Q4 dRij = q_normalized(q_from_matrix(ldm3(M + VILBA_PI_DR)));
    const Q4 dR_dbg = so3_exp(ldm3(M + VILBA_PI_JRG) * dbg);
    const Q4 rR = so3_mul(so3_mul(so3_inverse(so3_mul(dRij, dR_dbg)), RiT), Rj);
    rPhi = so3_log(rR);
}

// EdgeNavStateBias::computeError (g2otypes.cpp:703-722)
__device__ __forceinline__ void bias_error(const double* si, const double* sj, V3& rg, V3& ra) {
    rg = (ld3(sj + 10) + ld3(sj + 16)) - (ld3(si + 10) + ld3(si + 16));
    ra = (ld3(sj + 13) + ld3(sj + 19)) - (ld3(si + 13) + ld3(si + 19));
}

__device__ __forceinline__ double quad9(const double* info, const double* e) {  // e^T (info e)
    double s = 0.0;
#pragma unroll
    for (int r = 0; r < 9; ++r) {
        double t = 0.0;
#pragma unroll
        for (int c = 0; c < 9; ++c) t += info[9 * r + c] * e[c];
        s += e[r] * t;
    }
    return s;
}

__device__ __forceinline__ void atomic_max_nonneg(unsigned long long* addr, double v) {
    atomicMax(addr, (unsigned long long)__double_as_longlong(v));  // order-preserving for v >= 0
}

// Start of an LM iteration behind the linearisation (one CTA of 256 threads): computeLambdaInit at iteration 0
// (optimization_algorithm_levenberg.cpp:166-180: tau * max |diag H|), the per-iteration bookkeeping of solve() (:61-80),
// phase LINEARIZE -> TRIAL.  Runs at the tail of assemble_hpp (last CTA to arrive), or as its own kernel behind the
// reduction of diag H when the window is sharded.
__device__ __forceinline__ void lm_iter_begin_cta(const DevWindow& w, LmState* s) {
    __shared__ double red[8];
    const int iteration = s->iter;
    double m = 0.0;
    if (iteration == 0) {
        if (w.sharded)  // diag(H_pp) summed over the ranks | every rank's max |diag H_ll| (RED_DIAG)
            for (int d = threadIdx.x; d < w.n + w.shard_world; d += blockDim.x) m = fmax(m, fabs(w.diag_red[d]));
        else
            for (int d = threadIdx.x; d < w.n; d += blockDim.x) m = fmax(m, fabs(__ldcg(w.Hpp + (size_t)d * w.n + d)));
    }
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        if (iteration == 0) {  // computeLambdaInit (optimization_algorithm_levenberg.cpp:166-180)
            double mx = w.sharded ? 0.0 : __longlong_as_double((long long)s->maxdiag_bits);
            for (int i = 0; i < (int)(blockDim.x >> 5); ++i) mx = fmax(mx, red[i]);
            s->lambda = w.lm_tau * mx;
            s->ni = 2.0;
            s->n_bad = 0;
        }
        s->ini_chi = s->current_chi;
        s->lambda_first = s->lambda;
        s->qmax = 0;
        s->iter_result = -1;
        s->maxdiag_bits = 0ull;
        s->chi_acc = 0.0;
        s->scale_acc = 0.0;
        s->phase = PH_TRIAL;
    }
}

// ---- bulk asynchronous copies global -> shared memory (TMA, cp.async.bulk) completing on an mbarrier ----
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE%=;\n"
        "bra LAB_WAIT%=;\n"
        "DONE%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

}  // namespace vilba
