// vilba_api.cu -- host C++ controller + the extern "C" boundary declared in include/vilba.h.
//
// The host side mirrors the control flow of Optimizer::LocalBundleAdjustmentNavState phases C..E
// (src/Optimizer.cpp:2643-2701) and of SparseOptimizer::optimize / OptimizationAlgorithmLevenberg::solve
// (g2o/core/sparse_optimizer.cpp:354-419, optimization_algorithm_levenberg.cpp:61-164); all arithmetic
// runs in the CUDA kernels of lba_kernels.cu / preint.cu.  There is no CPU fallback: without a usable
// CUDA device vilba_create() returns NULL.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "kernels.h"

using namespace vilba;

namespace {

struct Arena {
    char* base = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (base) cudaFree(base);
        base = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + (1u << 20);
        cudaError_t e = cudaMalloc(&base, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (base) cudaFree(base);
        base = nullptr;
        cap = 0;
    }
};

struct Pinned {
    char* base = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (base) cudaFreeHost(base);
        base = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + (1u << 16);
        cudaError_t e = cudaMallocHost(&base, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (base) cudaFreeHost(base);
        base = nullptr;
        cap = 0;
    }
};

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// offsets of one window inside the device arena
struct Layout {
    // input section (one H2D copy)
    size_t kf_state0, pts0, imu_preint, obs0, pt_obs_begin, kf_block, imu_i, imu_j, blk_edge_i, blk_edge_j, blk_kf,
        pair_a, pair_b, input_end;
    size_t edge_pt, pair_begin, pair_ea, pair_eb, pt_mask;
    // work section
    size_t kf_state[2], pts[2], imu_info, imu_err, obs, obs_chi2, Hpp, bp, Hll, bl, W, lin_partial, imu_slot, mono_sum, Y, S, Lfac, cminv, cdinv, bs, x, lm, dbg, n_culled,
        outlier, total;
};

Layout make_layout(int K, int NI, int P, int E, int n, int n_free, int n_pairs, size_t n_triples, int lin_ctas) {
    Layout L;
    size_t o = 0;
    auto take = [&](size_t bytes) {
        size_t at = o;
        o = align_up(o + bytes);
        return at;
    };
    L.kf_state0 = take(sizeof(double) * 22 * (size_t)K);
    L.pts0 = take(sizeof(double) * 3 * (size_t)P);
    L.imu_preint = take(sizeof(double) * 142 * (size_t)NI);
    L.obs0 = take(sizeof(int4) * (size_t)E);
    L.pt_obs_begin = take(sizeof(int) * ((size_t)P + 1));
    L.kf_block = take(sizeof(int) * (size_t)K);
    L.imu_i = take(sizeof(int) * (size_t)NI);
    L.imu_j = take(sizeof(int) * (size_t)NI);
    L.blk_edge_i = take(sizeof(int) * (size_t)n_free);
    L.blk_edge_j = take(sizeof(int) * (size_t)n_free);
    L.blk_kf = take(sizeof(int) * (size_t)n_free);
    L.pair_a = take(sizeof(int) * (size_t)n_pairs);
    L.pair_b = take(sizeof(int) * (size_t)n_pairs);
    L.input_end = o;
    L.edge_pt = take(sizeof(int) * (size_t)E);
    L.pair_begin = take(sizeof(int) * ((size_t)n_pairs + 1));
    L.pair_ea = take(sizeof(int) * n_triples);
    L.pair_eb = take(sizeof(int) * n_triples);
    L.pt_mask = take(sizeof(unsigned long long) * 8 * (size_t)P);
    for (int b = 0; b < 2; ++b) L.kf_state[b] = take(sizeof(double) * 22 * (size_t)K);
    for (int b = 0; b < 2; ++b) L.pts[b] = take(sizeof(double) * 3 * (size_t)P);
    L.imu_info = take(sizeof(double) * 81 * (size_t)NI);
    L.imu_err = take(sizeof(double) * 15 * (size_t)NI);
    L.obs = take(sizeof(int4) * (size_t)E);
    L.obs_chi2 = take(sizeof(double) * (size_t)E);
    L.Hpp = take(sizeof(double) * (size_t)n * n);
    L.bp = take(sizeof(double) * (size_t)n);
    L.Hll = take(sizeof(double) * 6 * (size_t)P);
    L.bl = take(sizeof(double) * 3 * (size_t)P);
    L.W = take(sizeof(double) * 18 * (size_t)E);
    L.lin_partial = take(sizeof(double) * 27 * (size_t)n_free * lin_ctas);
    L.imu_slot = take(sizeof(double) * 930 * (size_t)NI);
    L.mono_sum = take(sizeof(double) * 27 * (size_t)n_free);
    L.Y = take(sizeof(double) * 24 * (size_t)E);
    const size_t lds = ((size_t)n + 3) & ~(size_t)3;
    L.S = take(sizeof(double) * lds * n);
    L.Lfac = take(sizeof(double) * lds * n);
    L.cminv = take(sizeof(double) * 256 * ((size_t)n / 16 + 2));
    L.cdinv = take(sizeof(double) * (size_t)n);
    L.bs = take(sizeof(double) * (size_t)n);
    L.x = take(sizeof(double) * (size_t)n);
    L.lm = take(sizeof(LmState));
    L.dbg = take(sizeof(long long) * 16);
    L.n_culled = take(sizeof(int) * 4);
    L.outlier = take((size_t)E);
    L.total = o;
    return L;
}

}  // namespace

struct vilba_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;  // side stream: IMU-edge linearisation runs beside the mono edges
    cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_fork = nullptr, ev_join = nullptr;
    vilba_params prm;
    Arena arena, preint_arena;
    Pinned pinned, pinned_small;
    std::string err;
    int sm_count = 148;
    // resident window
    bool has_window = false;
    Layout L;
    DevWindow dw;             // host copy
    DevWindow* dwp = nullptr; // device copy the kernels read (fixed address => graph-capturable launches)
    LaunchDims dims;
    int cap_K = 0, cap_nf = 0, cap_n = 0;  // capacities the shared-memory sizes / the graph were built for
    cudaGraphExec_t slot_graph = nullptr;
    bool use_graph = true;   // env VILBA_GRAPH=0 launches the slot kernels one by one
    int win_E = 0, win_NI = 0, win_P = 0, win_K = 0;
    int last_cur = 0;        // estimate buffer that holds the result of the last solve
    // stats
    vilba_stats stats;
    bool profiling = false;
    std::vector<vilba_ctx*> lanes;  // sub-contexts of vilba_local_ba_batch: independent windows run concurrently
    int n_lanes = 8;                // env VILBA_BATCH_LANES
    std::vector<cudaEvent_t> probes;  // 6 events per profiled slot
    size_t probes_used = 0;
    double dbg_ms[4] = {0, 0, 0, 0};
};

namespace {

bool fail(vilba_ctx* c, cudaError_t e, const char* what) {
    if (e == cudaSuccess) return false;
    c->err = std::string(what) + ": " + cudaGetErrorString(e);
    return true;
}

#define CK(call, what)                          \
    do {                                        \
        if (fail(ctx, (call), what)) return VILBA_ERR_CUDA; \
    } while (0)

int check_window(const vilba_window* w) {
    if (!w || w->n_kf <= 0 || w->n_imu < 0 || w->n_pts < 0 || w->n_obs < 0) return VILBA_ERR_ARG;
    if (w->n_kf > (OBS_KF_MASK) || w->n_kf > kMaxKF) return VILBA_ERR_ARG;
    if (!w->kf_state || !w->kf_flags) return VILBA_ERR_ARG;
    if (w->n_imu && (!w->imu_kf_i || !w->imu_kf_j || !w->imu_preint)) return VILBA_ERR_ARG;
    if (w->n_pts && (!w->pt_xyz || !w->pt_obs_begin)) return VILBA_ERR_ARG;
    if (w->n_obs && (!w->obs_kf || !w->obs_uv || !w->obs_inv_sigma2)) return VILBA_ERR_ARG;
    int n_free = 0;
    for (int k = 0; k < w->n_kf; ++k) {
        const bool fixed = (w->kf_flags[k] & VILBA_KF_FIXED) != 0;
        if (!fixed && !(w->kf_flags[k] & VILBA_KF_HAS_BIAS)) return VILBA_ERR_ARG;
        n_free += fixed ? 0 : 1;
    }
    if (n_free == 0) return VILBA_ERR_ARG;
    for (int e = 0; e < w->n_imu; ++e) {
        const int i = w->imu_kf_i[e], j = w->imu_kf_j[e];
        if (i < 0 || j < 0 || i >= w->n_kf || j >= w->n_kf) return VILBA_ERR_ARG;
        if (!(w->kf_flags[i] & VILBA_KF_HAS_BIAS) || !(w->kf_flags[j] & VILBA_KF_HAS_BIAS)) return VILBA_ERR_ARG;
    }
    if (w->n_pts && (w->pt_obs_begin[0] != 0 || w->pt_obs_begin[w->n_pts] != w->n_obs)) return VILBA_ERR_ARG;
    for (int e = 0; e < w->n_obs; ++e)
        if (w->obs_kf[e] < 0 || w->obs_kf[e] >= w->n_kf) return VILBA_ERR_ARG;
    // observations of a point ordered by key-frame (MapPoint::GetObservations order), each key-frame once
    for (int p = 0; p < w->n_pts; ++p) {
        if (w->pt_obs_begin[p + 1] < w->pt_obs_begin[p]) return VILBA_ERR_ARG;
        for (int e = w->pt_obs_begin[p] + 1; e < w->pt_obs_begin[p + 1]; ++e)
            if (w->obs_kf[e] <= w->obs_kf[e - 1]) return VILBA_ERR_ARG;
    }
    return VILBA_OK;
}

// profiling: six events per slot (see launch_slot); drained after a stream synchronize
cudaEvent_t* probe_take(vilba_ctx* ctx) {
    if (!ctx->profiling) return nullptr;
    if (ctx->probes_used + 8 > ctx->probes.size()) {
        for (int i = 0; i < 8; ++i) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            ctx->probes.push_back(e);
        }
    }
    cudaEvent_t* p = &ctx->probes[ctx->probes_used];
    ctx->probes_used += 8;
    return p;
}
void probe_drain(vilba_ctx* ctx) {
    for (size_t i = 0; i + 8 <= ctx->probes_used; i += 8) {
        float lin = 0, sch = 0, chol = 0, t = 0;
        cudaEvent_t* p = &ctx->probes[i];
        if (cudaEventElapsedTime(&lin, p[0], p[1]) != cudaSuccess) continue;
        cudaEventElapsedTime(&sch, p[2], p[3]);
        cudaEventElapsedTime(&chol, p[3], p[4]);
        // a slot whose group returned early (nothing to do in that phase) takes a few microseconds
        if (lin > 0.008f) {
            ctx->stats.linearize_ms += lin, ctx->stats.linearize_launches++;
            if (cudaEventElapsedTime(&t, p[0], p[7]) == cudaSuccess) ctx->dbg_ms[0] += t;  // mono linearize alone
        }
        if (chol > 0.008f) {
            ctx->stats.schur_ms += sch, ctx->stats.schur_launches++;
            ctx->stats.solve_ms += chol, ctx->stats.solve_launches++;
            if (cudaEventElapsedTime(&t, p[2], p[6]) == cudaSuccess) ctx->dbg_ms[1] += t;  // schur_prep alone
            if (cudaEventElapsedTime(&t, p[4], p[5]) == cudaSuccess) ctx->dbg_ms[2] += t;  // update_eval alone
        }
    }
    ctx->probes_used = 0;
}

// ------------------------------------------------------------------------------------------------
// flatten + upload: phase A/B of the reference function become "pack into pinned memory, one H2D"
// ------------------------------------------------------------------------------------------------
int upload_window(vilba_ctx* ctx, const vilba_window* w) {
    const auto t_begin = std::chrono::steady_clock::now();
    int st = check_window(w);
    if (st != VILBA_OK) {
        ctx->err = "invalid window";
        return st;
    }
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    const int K = w->n_kf, NI = w->n_imu, P = w->n_pts, E = w->n_obs;
    std::vector<int> kf_block(K, -1);
    int n_free = 0;
    for (int k = 0; k < K; ++k)
        if (!(w->kf_flags[k] & VILBA_KF_FIXED)) kf_block[k] = n_free++;
    const int n = 15 * n_free;
    // v2 accumulation structures: IMU edge of every block, map point of every edge, and per key-frame
    // block pair (a <= b) the list of (edge_a, edge_b) that share a map point
    std::vector<int> blk_edge_i(n_free, -1), blk_edge_j(n_free, -1);
    for (int e = 0; e < NI; ++e) {
        const int bi = kf_block[w->imu_kf_i[e]], bj = kf_block[w->imu_kf_j[e]];
        if ((bi >= 0 && blk_edge_i[bi] >= 0) || (bj >= 0 && blk_edge_j[bj] >= 0) || (bi >= 0 && bi == bj)) {
            ctx->err = "a key-frame may start / end at most one IMU edge";
            return VILBA_ERR_ARG;
        }
        if (bi >= 0) blk_edge_i[bi] = e;
        if (bj >= 0) blk_edge_j[bj] = e;
    }
    const int n_pairs = n_free * (n_free + 1) / 2;
    auto pair_index = [n_free](int a, int b) { return a * n_free - a * (a - 1) / 2 + (b - a); };
    std::vector<int> pair_a(n_pairs), pair_b(n_pairs), blk_kf(n_free);
    for (int a = 0; a < n_free; ++a)
        for (int b = a; b < n_free; ++b) pair_a[pair_index(a, b)] = a, pair_b[pair_index(a, b)] = b;
    for (int k = 0; k < K; ++k)
        if (kf_block[k] >= 0) blk_kf[kf_block[k]] = k;
    // the (edge_a, edge_b) lists themselves are built on the device (pairs.cu); capacity = sum m (m + 1) / 2
    size_t n_triples = 0;
    for (int p = 0; p < P; ++p) {
        const size_t m = (size_t)(w->pt_obs_begin[p + 1] - w->pt_obs_begin[p]);
        n_triples += m * (m + 1) / 2;
    }
    const int lin_ctas = ctx->dims.point_grid;
    const Layout L = make_layout(K, NI, P, E, n, n_free, n_pairs, n_triples, lin_ctas);
    CK(ctx->arena.reserve(L.total), "cudaMalloc(arena)");
    CK(ctx->pinned.reserve(L.input_end), "cudaMallocHost(staging)");
    char* h = ctx->pinned.base;
    std::memcpy(h + L.kf_state0, w->kf_state, sizeof(double) * 22 * (size_t)K);
    if (P) std::memcpy(h + L.pts0, w->pt_xyz, sizeof(double) * 3 * (size_t)P);
    if (NI) std::memcpy(h + L.imu_preint, w->imu_preint, sizeof(double) * 142 * (size_t)NI);
    int4* ho = reinterpret_cast<int4*>(h + L.obs0);
    for (int e = 0; e < E; ++e) {
        int4 r;
        std::memcpy(&r.x, &w->obs_uv[2 * (size_t)e], 4);
        std::memcpy(&r.y, &w->obs_uv[2 * (size_t)e + 1], 4);
        std::memcpy(&r.z, &w->obs_inv_sigma2[e], 4);
        r.w = w->obs_kf[e] | OBS_ROBUST;  // every mono edge starts with its Huber kernel (Optimizer.cpp:2622-2624)
        ho[e] = r;
    }
    if (P) std::memcpy(h + L.pt_obs_begin, w->pt_obs_begin, sizeof(int) * ((size_t)P + 1));
    else std::memset(h + L.pt_obs_begin, 0, sizeof(int));
    std::memcpy(h + L.kf_block, kf_block.data(), sizeof(int) * (size_t)K);
    if (NI) {
        std::memcpy(h + L.imu_i, w->imu_kf_i, sizeof(int) * (size_t)NI);
        std::memcpy(h + L.imu_j, w->imu_kf_j, sizeof(int) * (size_t)NI);
    }
    std::memcpy(h + L.blk_edge_i, blk_edge_i.data(), sizeof(int) * (size_t)n_free);
    std::memcpy(h + L.blk_edge_j, blk_edge_j.data(), sizeof(int) * (size_t)n_free);
    std::memcpy(h + L.blk_kf, blk_kf.data(), sizeof(int) * (size_t)n_free);
    std::memcpy(h + L.pair_a, pair_a.data(), sizeof(int) * (size_t)n_pairs);
    std::memcpy(h + L.pair_b, pair_b.data(), sizeof(int) * (size_t)n_pairs);
    char* d = ctx->arena.base;
    const auto t_packed = std::chrono::steady_clock::now();
    CK(cudaMemcpyAsync(d, h, L.input_end, cudaMemcpyHostToDevice, ctx->stream), "H2D window");
    if (std::getenv("VILBA_DEBUG_COUNTERS"))
        std::fprintf(stderr, "[vilba dbg] flatten+pair lists %.3f ms, %zu bytes H2D, %zu list entries\n",
                     std::chrono::duration<double, std::milli>(t_packed - t_begin).count(), L.input_end, n_triples);

    DevWindow& dw = ctx->dw;
    std::memset(&dw, 0, sizeof(dw));
    dw.K = K, dw.NI = NI, dw.P = P, dw.E = E, dw.n_free = n_free, dw.n = n;
    dw.lds = (n + 3) & ~3;
    for (int b = 0; b < 2; ++b) {
        dw.kf_state[b] = reinterpret_cast<double*>(d + L.kf_state[b]);
        dw.pts[b] = reinterpret_cast<double*>(d + L.pts[b]);
    }
    dw.kf_block = reinterpret_cast<const int*>(d + L.kf_block);
    dw.imu_i = reinterpret_cast<const int*>(d + L.imu_i);
    dw.imu_j = reinterpret_cast<const int*>(d + L.imu_j);
    dw.imu_preint = reinterpret_cast<const double*>(d + L.imu_preint);
    dw.imu_info = reinterpret_cast<double*>(d + L.imu_info);
    dw.imu_err = reinterpret_cast<double*>(d + L.imu_err);
    dw.pt_obs_begin = reinterpret_cast<const int*>(d + L.pt_obs_begin);
    dw.obs = reinterpret_cast<int4*>(d + L.obs);
    dw.obs_chi2 = reinterpret_cast<double*>(d + L.obs_chi2);
    dw.Hpp = reinterpret_cast<double*>(d + L.Hpp);
    dw.bp = reinterpret_cast<double*>(d + L.bp);
    dw.Hll = reinterpret_cast<double*>(d + L.Hll);
    dw.bl = reinterpret_cast<double*>(d + L.bl);
    dw.W = reinterpret_cast<double*>(d + L.W);
    dw.lin_partial = reinterpret_cast<double*>(d + L.lin_partial);
    dw.imu_slot = reinterpret_cast<double*>(d + L.imu_slot);
    dw.mono_sum = reinterpret_cast<double*>(d + L.mono_sum);
    dw.Y = reinterpret_cast<double*>(d + L.Y);
    dw.blk_edge_i = reinterpret_cast<const int*>(d + L.blk_edge_i);
    dw.blk_edge_j = reinterpret_cast<const int*>(d + L.blk_edge_j);
    dw.edge_pt = reinterpret_cast<const int*>(d + L.edge_pt);
    dw.n_pairs = n_pairs;
    dw.pair_a = reinterpret_cast<const int*>(d + L.pair_a);
    dw.pair_b = reinterpret_cast<const int*>(d + L.pair_b);
    dw.pair_begin = reinterpret_cast<const int*>(d + L.pair_begin);
    dw.pair_ea = reinterpret_cast<const int*>(d + L.pair_ea);
    dw.pair_eb = reinterpret_cast<const int*>(d + L.pair_eb);
    dw.blk_kf = reinterpret_cast<const int*>(d + L.blk_kf);
    dw.edge_pt_rw = reinterpret_cast<int*>(d + L.edge_pt);
    dw.pair_begin_rw = reinterpret_cast<int*>(d + L.pair_begin);
    dw.pair_ea_rw = reinterpret_cast<int*>(d + L.pair_ea);
    dw.pair_eb_rw = reinterpret_cast<int*>(d + L.pair_eb);
    dw.pt_mask = reinterpret_cast<unsigned long long*>(d + L.pt_mask);
    dw.S = reinterpret_cast<double*>(d + L.S);
    dw.Lfac = reinterpret_cast<double*>(d + L.Lfac);
    dw.cminv = reinterpret_cast<double*>(d + L.cminv);
    dw.cdinv = reinterpret_cast<double*>(d + L.cdinv);
    dw.bs = reinterpret_cast<double*>(d + L.bs);
    dw.x = reinterpret_cast<double*>(d + L.x);
    dw.lm = reinterpret_cast<LmState*>(d + L.lm);
    dw.dbg = reinterpret_cast<long long*>(d + L.dbg);
    if (const char* e = std::getenv("VILBA_CHOL_ABLATE")) dw.dbg_flags = std::atoi(e);
    dw.fx = w->fx, dw.fy = w->fy, dw.cx = w->cx, dw.cy = w->cy;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) dw.Rcb[3 * r + c] = w->Rbc[3 * c + r];  // Rcb = Rbc^T
    for (int r = 0; r < 3; ++r) {  // tcb = -(Rcb * Pbc)
        double s = 0.0;
        for (int c = 0; c < 3; ++c) s += dw.Rcb[3 * r + c] * w->Pbc[c];
        dw.tcb[r] = -s;
    }
    for (int i = 0; i < 3; ++i) dw.g[i] = w->gravity[i];
    const vilba_params& p = ctx->prm;
    dw.huber_mono = p.huber_mono, dw.huber_pvr = p.huber_pvr, dw.huber_bias = p.huber_bias;
    dw.chi2_gate = p.chi2_gate;
    dw.inv_gyr_rw2 = 1.0 / p.gyr_bias_rw2, dw.inv_acc_rw2 = 1.0 / p.acc_bias_rw2;
    dw.lm_tau = p.lm_tau, dw.lm_good_lo = p.lm_good_lo, dw.lm_good_hi = p.lm_good_hi;
    dw.max_trials = p.max_trials;

    ctx->L = L;
    ctx->win_E = E, ctx->win_NI = NI, ctx->win_P = P, ctx->win_K = K;
    // shared-memory capacities (and the captured graph) grow monotonically
    if (K > ctx->cap_K || n_free > ctx->cap_nf || n > ctx->cap_n) {
        ctx->cap_K = std::max(ctx->cap_K, (K + 31) / 32 * 32);
        ctx->cap_nf = std::max(ctx->cap_nf, (n_free + 7) / 8 * 8);
        ctx->cap_n = std::max(ctx->cap_n, 15 * ctx->cap_nf);
        ctx->dims.smem_point = point_smem_bytes(ctx->cap_K);
        ctx->dims.smem_lin = linearize_smem_bytes(ctx->cap_K, ctx->cap_nf);
        ctx->dims.smem_chol = chol_smem_bytes(ctx->cap_n);
        ctx->dims.chol_nb = chol_block_size(ctx->cap_n);
        if (const char* e = std::getenv("VILBA_CHOL_NB")) ctx->dims.chol_nb = (std::atoi(e) == 16) ? 16 : ctx->dims.chol_nb;
        if (ctx->dims.smem_lin > 227 * 1024 || ctx->dims.smem_chol > 227 * 1024) {
            ctx->err = "window too large for the shared-memory stages";
            return VILBA_ERR_ARG;
        }
        CK(configure_kernels(ctx->dims), "cudaFuncSetAttribute");
        if (ctx->slot_graph) {
            cudaGraphExecDestroy(ctx->slot_graph);
            ctx->slot_graph = nullptr;
        }
    }
    dw.chol_stage = chol_has_stage(ctx->cap_n) ? 1 : 0;
    // device copy of the descriptor (staged behind the inputs in the pinned buffer)
    CK(ctx->pinned_small.reserve(sizeof(DevWindow) + sizeof(LmState) + 256), "cudaMallocHost(desc)");
    std::memcpy(ctx->pinned_small.base, &dw, sizeof(DevWindow));
    CK(cudaMemcpyAsync(ctx->dwp, ctx->pinned_small.base, sizeof(DevWindow), cudaMemcpyHostToDevice, ctx->stream),
       "H2D descriptor");
    // the list builder reads the working copy of the observation table
    if (E)
        CK(cudaMemcpyAsync(d + L.obs, d + L.obs0, sizeof(int4) * (size_t)E, cudaMemcpyDeviceToDevice, ctx->stream),
           "obs copy");
    CK(launch_build_pair_lists(ctx->stream, ctx->dwp, ctx->dims), "pair lists");
    ctx->stats.kernel_launches += 4;
    ctx->has_window = true;
    return VILBA_OK;
}

// restart from the uploaded initial state (device-to-device)
int reset_window(vilba_ctx* ctx) {
    const Layout& L = ctx->L;
    char* d = ctx->arena.base;
    cudaStream_t s = ctx->stream;
    CK(cudaMemcpyAsync(d + L.kf_state[0], d + L.kf_state0, sizeof(double) * 22 * (size_t)ctx->win_K,
                       cudaMemcpyDeviceToDevice, s), "reset kf");
    CK(cudaMemcpyAsync(d + L.kf_state[1], d + L.kf_state0, sizeof(double) * 22 * (size_t)ctx->win_K,
                       cudaMemcpyDeviceToDevice, s), "reset kf");
    if (ctx->win_P) {
        CK(cudaMemcpyAsync(d + L.pts[0], d + L.pts0, sizeof(double) * 3 * (size_t)ctx->win_P,
                           cudaMemcpyDeviceToDevice, s), "reset pts");
        CK(cudaMemcpyAsync(d + L.pts[1], d + L.pts0, sizeof(double) * 3 * (size_t)ctx->win_P,
                           cudaMemcpyDeviceToDevice, s), "reset pts");
    }
    if (ctx->win_E) {
        CK(cudaMemcpyAsync(d + L.obs, d + L.obs0, sizeof(int4) * (size_t)ctx->win_E, cudaMemcpyDeviceToDevice, s),
           "reset obs");
        CK(cudaMemsetAsync(d + L.obs_chi2, 0, sizeof(double) * (size_t)ctx->win_E, s), "reset chi2");
    }
    CK(cudaMemsetAsync(d + L.lm, 0, sizeof(LmState), s), "reset lm");
    if (std::getenv("VILBA_DEBUG_COUNTERS") && ctx->stats.solve_launches > 0) {
        std::fprintf(stderr, "[vilba dbg] per launch: linearize_v2 %.1f us, schur_prep %.1f us, update_eval %.1f us\n",
                     1e3 * ctx->dbg_ms[0] / std::max<long long>(1, ctx->stats.linearize_launches),
                     1e3 * ctx->dbg_ms[1] / ctx->stats.solve_launches, 1e3 * ctx->dbg_ms[2] / ctx->stats.solve_launches);
    }
    if (std::getenv("VILBA_DEBUG_COUNTERS")) {
        long long h[16];
        if (cudaMemcpy(h, d + L.dbg, sizeof(h), cudaMemcpyDeviceToHost) == cudaSuccess && h[8] > 0) {
            std::fprintf(stderr, "[vilba dbg] chol calls=%lld avg cycles/phase:", h[8]);
            for (int i = 0; i < 8; ++i) std::fprintf(stderr, " %lld", h[i] / h[8]);
            std::fprintf(stderr, "\n");
        }
        CK(cudaMemsetAsync(d + L.dbg, 0, sizeof(long long) * 16, s), "reset dbg");
    }
    return VILBA_OK;
}

// D2H of the controller state; waits for the stream while polling the caller's stop flag
int read_lm(vilba_ctx* ctx, LmState* out, const volatile uint8_t* stop_flag) {
    char* hp = ctx->pinned_small.base + sizeof(DevWindow) + 64;
    CK(cudaMemcpyAsync(hp, ctx->dw.lm, sizeof(LmState), cudaMemcpyDeviceToHost, ctx->stream), "D2H lm");
    if (stop_flag) {
        bool sent = false;
        while (cudaStreamQuery(ctx->stream) == cudaErrorNotReady) {
            if (!sent && *stop_flag) {  // mirror of `bool* pbStopFlag` (sparse_optimizer.h:188)
                static const int one = 1;
                cudaMemcpyAsync(&ctx->dw.lm->stop, &one, sizeof(int), cudaMemcpyHostToDevice, ctx->stream2);
                sent = true;
            }
        }
    }
    CK(cudaStreamSynchronize(ctx->stream), "sync");
    std::memcpy(out, hp, sizeof(LmState));
    probe_drain(ctx);
    return VILBA_OK;
}

bool stop_requested(const volatile uint8_t* f) { return f && *f; }

int ensure_graph(vilba_ctx* ctx) {
    if (!ctx->use_graph || ctx->slot_graph) return VILBA_OK;
    cudaGraph_t g = nullptr;
    CK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal), "begin capture");
    cudaError_t e = launch_slot(ctx->stream, ctx->stream2, ctx->ev_fork, ctx->ev_join, ctx->dwp, ctx->dims, nullptr);
    cudaError_t e2 = cudaStreamEndCapture(ctx->stream, &g);
    if (fail(ctx, e, "capture slot") || fail(ctx, e2, "end capture")) return VILBA_ERR_CUDA;
    e = cudaGraphInstantiate(&ctx->slot_graph, g, 0);
    cudaGraphDestroy(g);
    if (fail(ctx, e, "graph instantiate")) return VILBA_ERR_CUDA;
    return VILBA_OK;
}

// SparseOptimizer::optimize(iterations) with OptimizationAlgorithmLevenberg (sparse_optimizer.cpp:354-419).
// The loop itself runs on the device: the host enqueues slots (CUDA graph launches) without
// synchronising and looks at the controller state once they have drained.
int run_stage(vilba_ctx* ctx, int stage, int iterations, vilba_result* out, const volatile uint8_t* stop_flag,
              LmState* lm) {
    cudaStream_t s = ctx->stream;
    vilba_stats& stt = ctx->stats;
    // computeActiveErrors + activeRobustChi2 at the first iteration; later iterations inherit currentChi
    // of the accepted trial (identical by construction: the errors are those of the accepted state)
    CK(launch_eval_initial(s, ctx->dwp, ctx->dims), "eval");
    CK(launch_stage_begin(s, ctx->dwp, stage, iterations), "stage_begin");
    stt.kernel_launches += 2;
    const bool graph = ctx->use_graph && !ctx->profiling;
    if (graph) {
        int r = ensure_graph(ctx);
        if (r != VILBA_OK) return r;
    }
    const int first_trace = lm->n_trace;
    int slots = iterations + 1;  // one trial per iteration when every step is accepted, plus one spare
    for (int round = 0; round < 64; ++round) {
        for (int i = 0; i < slots; ++i) {
            if (graph)
                CK(cudaGraphLaunch(ctx->slot_graph, s), "graph launch");
            else
                CK(launch_slot(s, ctx->stream2, ctx->ev_fork, ctx->ev_join, ctx->dwp, ctx->dims, probe_take(ctx)), "slot");
            stt.kernel_launches += kKernelsPerSlot;
        }
        int r = read_lm(ctx, lm, stop_flag);
        if (r != VILBA_OK) return r;
        if (lm->phase == PH_DONE) break;
        slots = 2;  // rejected trials used up the slack: keep going
    }
    for (int i = first_trace; i < lm->n_trace && out->n_trace < VILBA_MAX_TRACE; ++i) {
        const IterRec& t = lm->trace[i];
        vilba_iter_record& rec = out->trace[out->n_trace++];
        rec.stage = t.stage, rec.iteration = t.iteration, rec.trials = t.trials, rec.result = t.result;
        rec.n_active_edges = t.n_active, rec.accepted = t.accepted;
        rec.chi2_initial = t.chi0, rec.chi2_final = t.chi1, rec.lambda = t.lambda, rec.lambda_first_trial = t.lambda_first;
        stt.lm_iterations++;
        stt.lm_trials += t.trials;
        stt.edges_linearized += t.n_active;
    }
    return VILBA_OK;
}

int solve_resident(vilba_ctx* ctx, vilba_result* out, const volatile uint8_t* stop_flag) {
    if (!ctx->has_window) {
        ctx->err = "no window uploaded";
        return VILBA_ERR_ARG;
    }
    out->n_trace = 0;
    out->stage2_ran = 0;
    out->n_outliers_stage1 = 0;
    out->solve_ms = 0.0;
    out->status = VILBA_OK;
    if (stop_requested(stop_flag)) {  // Optimizer.cpp:2643-2645
        out->status = VILBA_ABORTED;
        return VILBA_ABORTED;
    }
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    cudaStream_t s = ctx->stream;
    int r = reset_window(ctx);
    if (r != VILBA_OK) return r;
    CK(cudaEventRecord(ctx->ev_a, s), "event");
    CK(launch_imu_prepare(s, ctx->dwp), "imu_prepare");
    ctx->stats.kernel_launches += 1;
    LmState lm;
    std::memset(&lm, 0, sizeof(lm));
    r = run_stage(ctx, 1, ctx->prm.iters_stage1, out, stop_flag, &lm);
    if (r != VILBA_OK) return r;
    if (!stop_requested(stop_flag) && !lm.stop) {  // bDoMore (Optimizer.cpp:2650-2656)
        CK(launch_cull(s, ctx->dwp, ctx->dims), "cull");
        ctx->stats.kernel_launches += 1;
        r = run_stage(ctx, 2, ctx->prm.iters_stage2, out, stop_flag, &lm);
        if (r != VILBA_OK) return r;
        out->n_outliers_stage1 = lm.n_culled;
        out->stage2_ran = 1;
    }
    uint8_t* outl = reinterpret_cast<uint8_t*>(ctx->arena.base + ctx->L.outlier);
    CK(launch_final_flags(s, ctx->dwp, ctx->dims, outl), "final_flags");
    ctx->stats.kernel_launches += 1;
    CK(cudaEventRecord(ctx->ev_b, s), "event");
    CK(cudaStreamSynchronize(s), "sync");
    probe_drain(ctx);
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b), "elapsed");
    out->solve_ms = ms;
    ctx->last_cur = lm.cur;
    return VILBA_OK;
}

int download_window(vilba_ctx* ctx, vilba_result* out) {
    if (!ctx->has_window) return VILBA_ERR_ARG;
    struct { int cur; } lm = {ctx->last_cur};
    const Layout& L = ctx->L;
    char* d = ctx->arena.base;
    cudaStream_t s = ctx->stream;
    if (out->kf_state)
        CK(cudaMemcpyAsync(out->kf_state, d + L.kf_state[lm.cur], sizeof(double) * 22 * (size_t)ctx->win_K,
                           cudaMemcpyDeviceToHost, s), "D2H kf");
    if (out->pt_xyz && ctx->win_P)
        CK(cudaMemcpyAsync(out->pt_xyz, d + L.pts[lm.cur], sizeof(double) * 3 * (size_t)ctx->win_P,
                           cudaMemcpyDeviceToHost, s), "D2H pts");
    if (out->obs_outlier && ctx->win_E)
        CK(cudaMemcpyAsync(out->obs_outlier, d + L.outlier, (size_t)ctx->win_E, cudaMemcpyDeviceToHost, s),
           "D2H outlier");
    if (out->obs_chi2 && ctx->win_E)
        CK(cudaMemcpyAsync(out->obs_chi2, d + L.obs_chi2, sizeof(double) * (size_t)ctx->win_E,
                           cudaMemcpyDeviceToHost, s), "D2H chi2");
    CK(cudaStreamSynchronize(s), "sync");
    return VILBA_OK;
}

}  // namespace

// ================================================================================================
// extern "C"
// ================================================================================================
extern "C" {

void vilba_default_params(vilba_params* p) {
    std::memset(p, 0, sizeof(*p));
    p->iters_stage1 = 5;
    p->iters_stage2 = 10;
    p->max_trials = 10;
    p->huber_mono = (double)(float)std::sqrt(5.991);
    p->huber_pvr = (double)(float)std::sqrt(100 * 21.666);
    p->huber_bias = (double)(float)std::sqrt(100 * 16.812);
    p->chi2_gate = 5.991;
    p->lm_tau = 1e-5;
    p->lm_good_lo = 1. / 3.;
    p->lm_good_hi = 2. / 3.;
    p->gyr_bias_rw2 = 2.0e-5 * 2.0e-5;
    p->acc_bias_rw2 = 5.0e-3 * 5.0e-3;
    p->gyr_meas_cov = 1.7e-4 * 1.7e-4 / 0.005;
    p->acc_meas_cov = 2.0e-3 * 2.0e-3 / 0.005 * 100;
}

const char* vilba_version(void) { return "vilba 0.1 (sm_100a, FP64)"; }

vilba_ctx* vilba_create(int device, const vilba_params* params) {
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0 || device < 0 || device >= n_dev) {
        std::fprintf(stderr, "vilba_create: no usable CUDA device %d (found %d); there is no CPU fallback\n", device,
                     n_dev);
        return nullptr;
    }
    if (cudaSetDevice(device) != cudaSuccess) return nullptr;
    vilba_ctx* ctx = new vilba_ctx();
    ctx->device = device;
    if (params)
        ctx->prm = *params;
    else
        vilba_default_params(&ctx->prm);
    std::memset(&ctx->stats, 0, sizeof(ctx->stats));
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
    ctx->dims.sm_count = ctx->sm_count;
    ctx->dims.point_grid = kPointGridPerSM * ctx->sm_count;
    ctx->dims.chol_cluster = 8;
    ctx->dims.chol_nb = 32;
    ctx->dims.smem_point = ctx->dims.smem_lin = ctx->dims.smem_chol = 0;
    if (const char* e = std::getenv("VILBA_CHOL_CLUSTER")) ctx->dims.chol_cluster = std::max(1, std::atoi(e));
    if (const char* e = std::getenv("VILBA_GRAPH")) ctx->use_graph = std::atoi(e) != 0;
    if (const char* e = std::getenv("VILBA_BATCH_LANES")) ctx->n_lanes = std::max(1, std::atoi(e));
    if (cudaMalloc(&ctx->dwp, sizeof(DevWindow)) != cudaSuccess) {
        delete ctx;
        return nullptr;
    }
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreate(&ctx->ev_a) != cudaSuccess || cudaEventCreate(&ctx->ev_b) != cudaSuccess) {
        std::fprintf(stderr, "vilba_create: %s\n", cudaGetErrorString(cudaGetLastError()));
        delete ctx;
        return nullptr;
    }
    return ctx;
}

void vilba_destroy(vilba_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (vilba_ctx* sub : ctx->lanes) vilba_destroy(sub);
    ctx->lanes.clear();
    cudaSetDevice(ctx->device);
    for (cudaEvent_t e : ctx->probes) cudaEventDestroy(e);
    if (ctx->slot_graph) cudaGraphExecDestroy(ctx->slot_graph);
    if (ctx->dwp) cudaFree(ctx->dwp);
    ctx->arena.release();
    ctx->preint_arena.release();
    ctx->pinned.release();
    ctx->pinned_small.release();
    if (ctx->ev_a) cudaEventDestroy(ctx->ev_a);
    if (ctx->ev_b) cudaEventDestroy(ctx->ev_b);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* vilba_last_error(const vilba_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int vilba_window_upload(vilba_ctx* ctx, const vilba_window* win) {
    if (!ctx) return VILBA_ERR_ARG;
    return upload_window(ctx, win);
}

int vilba_window_solve_resident(vilba_ctx* ctx, vilba_result* out) {
    if (!ctx || !out) return VILBA_ERR_ARG;
    int r = solve_resident(ctx, out, nullptr);
    out->status = r;
    return r;
}

int vilba_window_download(vilba_ctx* ctx, vilba_result* out) {
    if (!ctx || !out) return VILBA_ERR_ARG;
    return download_window(ctx, out);
}

int vilba_local_ba(vilba_ctx* ctx, const vilba_window* win, vilba_result* out, const volatile uint8_t* stop_flag) {
    if (!ctx || !out) return VILBA_ERR_ARG;
    out->n_trace = 0;
    out->stage2_ran = 0;
    out->n_outliers_stage1 = 0;
    out->solve_ms = 0.0;
    if (stop_flag && *stop_flag) {  // silent early return, nothing written (Optimizer.cpp:2643-2645)
        out->status = VILBA_ABORTED;
        return VILBA_ABORTED;
    }
    int r = upload_window(ctx, win);
    if (r == VILBA_OK) r = solve_resident(ctx, out, stop_flag);
    if (r == VILBA_OK) r = download_window(ctx, out);
    out->status = r;
    return r;
}

// Independent windows (BASELINE config 5): a pool of sub-contexts ("lanes"), each with its own streams, arena
// and CUDA graph, driven by one host thread per lane.  A single window leaves most of the GPU idle during its
// latency-bound phases (the 8-CTA Cholesky cluster above all), so concurrent windows fill the machine.
int vilba_local_ba_batch(vilba_ctx* ctx, int32_t n_windows, const vilba_window* win, vilba_result* out) {
    if (!ctx || n_windows < 0 || (n_windows && (!win || !out))) return VILBA_ERR_ARG;
    if (n_windows == 0) return VILBA_OK;
    const int lanes = std::max(1, std::min(ctx->n_lanes, (int)n_windows));
    while ((int)ctx->lanes.size() < lanes) {
        vilba_ctx* sub = vilba_create(ctx->device, &ctx->prm);
        if (!sub) {
            ctx->err = "could not create a batch lane";
            return VILBA_ERR_CUDA;
        }
        ctx->lanes.push_back(sub);
    }
    std::vector<int> status(lanes, VILBA_OK);
    std::vector<std::thread> th;
    for (int l = 0; l < lanes; ++l)
        th.emplace_back([&, l]() {
            for (int i = l; i < n_windows; i += lanes) {
                const int r = vilba_local_ba(ctx->lanes[l], &win[i], &out[i], nullptr);
                if (r < 0) status[l] = r;
            }
        });
    for (auto& t : th) t.join();
    int worst = VILBA_OK;
    for (int l = 0; l < lanes; ++l) {
        const vilba_stats& s = ctx->lanes[l]->stats;
        ctx->stats.kernel_launches += s.kernel_launches;
        ctx->stats.lm_iterations += s.lm_iterations;
        ctx->stats.lm_trials += s.lm_trials;
        ctx->stats.edges_linearized += s.edges_linearized;
        std::memset(&ctx->lanes[l]->stats, 0, sizeof(vilba_stats));
        if (status[l] < 0) {
            worst = status[l];
            ctx->err = ctx->lanes[l]->err;
        }
    }
    return worst;
}

int vilba_preintegrate_batch_dev(vilba_ctx* ctx, int32_t n_pairs, int32_t n_samples, const int32_t* sample_begin_dev,
                                 const double* gyro_dev, const double* acc_dev, const double* dt_dev,
                                 const double* bg_dev, const double* ba_dev, double* out_dev) {
    if (!ctx || n_pairs < 0) return VILBA_ERR_ARG;
    (void)n_samples;
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    CK(launch_preint_batch(ctx->stream, n_pairs, sample_begin_dev, gyro_dev, acc_dev, dt_dev, bg_dev, ba_dev, out_dev,
                           ctx->prm.gyr_meas_cov, ctx->prm.acc_meas_cov, 8), "preint_batch");
    ctx->stats.kernel_launches += n_pairs > 0 ? 1 : 0;
    return VILBA_OK;
}

int vilba_preintegrate_batch(vilba_ctx* ctx, int32_t n_pairs, const int32_t* sample_begin, const double* gyro,
                             const double* acc, const double* dt, const double* bg, const double* ba, double* out) {
    if (!ctx || n_pairs < 0) return VILBA_ERR_ARG;
    if (n_pairs == 0) return VILBA_OK;
    if (!sample_begin || !gyro || !acc || !dt || !bg || !ba || !out) return VILBA_ERR_ARG;
    for (int p = 0; p < n_pairs; ++p)
        if (sample_begin[p + 1] < sample_begin[p]) return VILBA_ERR_ARG;
    if (sample_begin[0] != 0) return VILBA_ERR_ARG;
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    const size_t ns = (size_t)sample_begin[n_pairs];
    size_t o = 0;
    auto take = [&](size_t bytes) {
        size_t at = o;
        o = align_up(o + bytes);
        return at;
    };
    const size_t o_sb = take(sizeof(int) * ((size_t)n_pairs + 1));
    const size_t o_g = take(sizeof(double) * 3 * ns), o_a = take(sizeof(double) * 3 * ns), o_t = take(sizeof(double) * ns);
    const size_t o_bg = take(sizeof(double) * 3 * (size_t)n_pairs), o_ba = take(sizeof(double) * 3 * (size_t)n_pairs);
    const size_t in_end = o;
    const size_t o_out = take(sizeof(double) * 142 * (size_t)n_pairs);
    CK(ctx->preint_arena.reserve(o), "cudaMalloc(preint)");
    CK(ctx->pinned.reserve(in_end > sizeof(double) * 142 * (size_t)n_pairs ? in_end : sizeof(double) * 142 * (size_t)n_pairs),
       "cudaMallocHost(preint)");
    char* h = ctx->pinned.base;
    std::memcpy(h + o_sb, sample_begin, sizeof(int) * ((size_t)n_pairs + 1));
    std::memcpy(h + o_g, gyro, sizeof(double) * 3 * ns);
    std::memcpy(h + o_a, acc, sizeof(double) * 3 * ns);
    std::memcpy(h + o_t, dt, sizeof(double) * ns);
    std::memcpy(h + o_bg, bg, sizeof(double) * 3 * (size_t)n_pairs);
    std::memcpy(h + o_ba, ba, sizeof(double) * 3 * (size_t)n_pairs);
    char* d = ctx->preint_arena.base;
    CK(cudaMemcpyAsync(d, h, in_end, cudaMemcpyHostToDevice, ctx->stream), "H2D preint");
    int r = vilba_preintegrate_batch_dev(ctx, n_pairs, (int)ns, reinterpret_cast<const int*>(d + o_sb),
                                         reinterpret_cast<const double*>(d + o_g), reinterpret_cast<const double*>(d + o_a),
                                         reinterpret_cast<const double*>(d + o_t), reinterpret_cast<const double*>(d + o_bg),
                                         reinterpret_cast<const double*>(d + o_ba), reinterpret_cast<double*>(d + o_out));
    if (r != VILBA_OK) return r;
    CK(cudaMemcpyAsync(h, d + o_out, sizeof(double) * 142 * (size_t)n_pairs, cudaMemcpyDeviceToHost, ctx->stream),
       "D2H preint");
    CK(cudaStreamSynchronize(ctx->stream), "sync");
    std::memcpy(out, h, sizeof(double) * 142 * (size_t)n_pairs);
    return VILBA_OK;
}

void vilba_get_stats(const vilba_ctx* ctx, vilba_stats* s) {
    if (ctx && s) *s = ctx->stats;
}
void vilba_reset_stats(vilba_ctx* ctx) {
    if (ctx) std::memset(&ctx->stats, 0, sizeof(ctx->stats));
}
void vilba_set_profiling(vilba_ctx* ctx, int on) {
    if (ctx) ctx->profiling = on != 0;
}

}  // extern "C"
