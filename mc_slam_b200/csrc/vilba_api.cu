// vilba_api.cu -- host C++ controller + the extern "C" boundary declared in include/vilba.h.
//
// The host side mirrors the control flow of Optimizer::LocalBundleAdjustmentNavState phases C..E
// (src/Optimizer.cpp:2643-2701) and of SparseOptimizer::optimize / OptimizationAlgorithmLevenberg::solve
// (g2o/core/sparse_optimizer.cpp:354-419, optimization_algorithm_levenberg.cpp:61-164); all arithmetic
// runs in the CUDA kernels of lba_kernels.cu / lba_v2.cu / chol_la.cu / chol_big.cu / preint.cu.  There is no CPU fallback:
// without a usable CUDA device vilba_create() returns NULL.
//
// A context holds a *batch* of 1..kMaxBatch independent windows.  Every kernel is launched once for the
// whole batch (grid.y = window), each window carries its own device-resident LM controller, and a window
// that has finished simply returns early from the remaining launches.  A single window is a batch of one.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <mutex>
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

// offsets of one window inside the three regions of the device arena
struct Layout {
    // input region (all windows contiguous => one H2D copy per batch); offsets relative to in_base
    size_t kf_state0, pts0, imu_preint, obs0, pt_obs_begin, kf_block, imu_i, imu_j, blk_edge_i, blk_edge_j, blk_kf,
        pair_a, pair_b, in_bytes;
    // output region (all windows contiguous => one D2H copy per batch); offsets relative to out_base
    size_t o_kf, o_pts, o_outlier, out_bytes;
    size_t chi_bytes;  // per-edge chi2 lives in a second region behind the outputs of all windows: copied only on request
    // work region; offsets relative to wk_base
    size_t edge_pt, pair_begin, pair_ea, pair_eb, pt_mask;
    size_t kf_state[2], pts[2], imu_info, imu_err, obs, obs_chi2, Hpp, bp, Hll, bl, W, lin_partial, imu_slot, mono_sum, Y,
        schur_partial, ts_rec, ts_hdr, S, Lfac, cminv, cdinv, bs, x, dbg, outlier, chi_partial, chi_counter, diag_red, S_part, hpp_span, s_span, wk_bytes;
};

struct WinMeta {
    int K = 0, NI = 0, P = 0, E = 0, n_free = 0, n = 0, n_pairs = 0;
    size_t n_triples = 0;
    int tile_edges[3] = {0, 0, 0};  // most edges in a tile of 32 / 16 / 8 map points
    Layout L;
    size_t in_base = 0, out_base = 0, chi_base = 0, wk_base = 0;
};

Layout make_layout(const WinMeta& m, int lin_ctas, int sp_ctas, int tile_pts, bool sharded) {
    const bool gather = sp_ctas == 0;  // the pair lists and Y = W D^-1 only exist for the Schur gather
    const size_t K = m.K, NI = m.NI, P = m.P, E = m.E, n = m.n, n_free = m.n_free, n_pairs = m.n_pairs;
    Layout L;
    size_t o = 0;
    auto take = [&](size_t bytes) {
        size_t at = o;
        o = align_up(o + bytes);
        return at;
    };
    L.kf_state0 = take(sizeof(double) * 22 * K);
    L.pts0 = take(sizeof(double) * 3 * P);
    L.imu_preint = take(sizeof(double) * 142 * NI);
    L.obs0 = take(sizeof(int4) * E);
    L.pt_obs_begin = take(sizeof(int) * (P + 1));
    L.kf_block = take(sizeof(int) * K);
    L.imu_i = take(sizeof(int) * NI);
    L.imu_j = take(sizeof(int) * NI);
    L.blk_edge_i = take(sizeof(int) * n_free);
    L.blk_edge_j = take(sizeof(int) * n_free);
    L.blk_kf = take(sizeof(int) * n_free);
    L.pair_a = take(sizeof(int) * n_pairs);
    L.pair_b = take(sizeof(int) * n_pairs);
    L.in_bytes = o;
    o = 0;
    L.o_kf = take(sizeof(double) * 22 * K);
    L.o_pts = take(sizeof(double) * 3 * P);
    L.o_outlier = take(E);
    L.out_bytes = o;
    o = 0;
    take(sizeof(double) * E);
    L.chi_bytes = o;
    o = 0;
    L.edge_pt = take(sizeof(int) * E);
    L.pair_begin = take(sizeof(int) * (n_pairs + 1));
    L.pair_ea = take(gather ? sizeof(int) * m.n_triples : 0);
    L.pair_eb = take(gather ? sizeof(int) * m.n_triples : 0);
    L.pt_mask = take(sizeof(unsigned long long) * 8 * P);
    for (int b = 0; b < 2; ++b) L.kf_state[b] = take(sizeof(double) * 22 * K);
    for (int b = 0; b < 2; ++b) L.pts[b] = take(sizeof(double) * 3 * P);
    L.imu_info = take(sizeof(double) * 81 * NI);
    L.imu_err = take(sizeof(double) * 15 * NI);
    L.obs = take(sizeof(int4) * E);
    L.obs_chi2 = take(sizeof(double) * E);
    L.Hpp = take(sizeof(double) * n * n);
    L.bp = take(sizeof(double) * n);  // directly behind H_pp (same reason)
    L.hpp_span = L.bp + sizeof(double) * n - L.Hpp;
    L.Hll = take(sizeof(double) * 6 * P);
    L.bl = take(sizeof(double) * 3 * P);
    L.W = take(sizeof(double) * 18 * E);
    L.lin_partial = take(sizeof(double) * 27 * n_free * (size_t)lin_ctas);
    L.imu_slot = take(sizeof(double) * 930 * NI);
    L.mono_sum = take(sizeof(double) * 27 * n_free);
    L.Y = take(gather ? sizeof(double) * 24 * E : 0);
    L.schur_partial = take(sizeof(double) * std::max(schur_partial_doubles((int)n_free), schur_pair_partial_doubles((int)n_free)) * (size_t)sp_ctas);
    L.ts_rec = take(gather ? 0 : sizeof(double) * schur_tile_rec_doubles((int)P) + 256);
    L.ts_hdr = take(gather ? 0 : sizeof(unsigned) * schur_tile_hdr_words((int)P, tile_pts));
    const size_t lds = (n + 3) & ~(size_t)3;
    L.S = take(sizeof(double) * lds * n);
    L.bs = take(sizeof(double) * n);  // directly behind S: one allreduce covers S | b_s of a sharded window
    L.s_span = L.bs + sizeof(double) * n - L.S;
    L.Lfac = take(sizeof(double) * lds * n);
    L.cminv = take(sizeof(double) * std::max<size_t>(chol_la_scratch_doubles((int)n), chol_big_scratch_doubles((int)n)));
    L.cdinv = take(sizeof(double) * n);
    L.x = take(sizeof(double) * n);
    L.dbg = take(sizeof(long long) * 16);
    L.outlier = take(E);
    L.chi_partial = take(sizeof(double) * 2 * (size_t)lin_ctas);
    L.chi_counter = take(sizeof(unsigned) * 4);
    L.diag_red = take(sharded ? sizeof(double) * (n + 64) : 0);  // diag(H_pp) | per-rank max |diag H_ll| (one small allreduce)
    L.S_part = take(sharded ? L.s_span : 0);                     // send buffer of the allreduce of S | b_s
    L.wk_bytes = o;
    return L;
}

struct GraphEntry {
    LaunchDims dims;
    cudaGraphExec_t exec;
};

}  // namespace

struct vilba_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;  // side stream: IMU-edge linearisation runs beside the mono edges
    cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_fork = nullptr, ev_join = nullptr;
    vilba_params prm;
    Arena arena, preint_arena;
    Pinned pinned, pinned_out, pinned_small, pinned_preint;  // (pre-integration has its own staging: an upload's H2D may still read `pinned`)
    std::string err;
    int sm_count = 148;
    int max_batch = kMaxBatch;
    // resident batch
    int n_win = 0;
    std::vector<WinMeta> meta;
    std::vector<DevWindow> dw;  // host copies
    DevWindow* dwp = nullptr;   // device array the kernels read (fixed address => graph-capturable launches)
    size_t in_total = 0, out_total = 0, out_small = 0, lm_base = 0, out_region = 0;  // arena: [in | out | chi2 | lm array | work]
    LaunchDims dims;
    int chol_cluster = 8;
    int preint_group = 0;            // env VILBA_PREINT_GROUP: 0 = scan kernel, lanes per pair chosen from the average interval length;
                                     // -8/-16/-32 = scan kernel with that many lanes per pair; 8/16/32 = the sequential
                                     // entry-parallel kernel
    int chol_la_mode = 1;            // env VILBA_CHOL_LA: 0 never, 1 automatic, 2 always (when the tiles fit)
    int chol_big_above = 480;        // env VILBA_CHOL_BIG_ABOVE: reduced systems larger than this use chol_big.cu (the cluster
                                     // kernel's panel + row stage fit in shared memory up to n = 508)
    bool schur_gather_only = false;  // env VILBA_SCHUR=gather (ablation)
    int sp_grid_cap = 74;            // env VILBA_SP_GRID: point subsets per window of the tile-scan Schur kernel
    int sp_pair_lanes = 0;           // env VILBA_SP_PAIR=1: lane-per-pair variant of the tile-scan Schur kernel
    int sp_sets = 0;                 // env VILBA_SP_SETS: block-pair subsets (0 = automatic)
    int sp_mma = 0;                  // env VILBA_SP_MMA=1: tile kernel with one warp per hit on the FP64 MMA (lba_v2.cu)
    int tile_edges[3] = {0, 0, 0};   // most edges in a tile of 32 / 16 / 8 map points over the current batch
    int cap_K = 0, cap_nf = 0, cap_n = 0;  // capacities the shared-memory sizes were configured for
    std::vector<GraphEntry> graphs;        // one captured LM slot per launch geometry
    bool use_graph = true;                 // env VILBA_GRAPH=0 launches the slot kernels one by one
    // stats
    vilba_stats stats;
    bool profiling = false;
    // point-sharded window over several GPUs (one process per GPU): NCCL is loaded at run time
    void* nccl_lib = nullptr;
    ncclComm_t comm = nullptr;
    int comm_rank = 0, comm_world = 1;
    bool comm_graph = true;   // the sharded slot (kernels + allreduces) is captured in a CUDA graph too; env VILBA_COMM_GRAPH=0 disables
    ncclResult_t (*p_ncclCommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*p_ncclCommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*p_ncclAllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*p_ncclGetErrorString)(ncclResult_t) = nullptr;
    SlotComm slot_comm;
    // a large batch is split over a few sub-contexts ("lanes": own streams, arena, graphs, one host thread each)
    // whose batched solves run concurrently: the kernels are latency-bound, a second and third stream fill the SMs
    int batch_total_hint = 0;            // lane: windows of the whole batch (all lanes), for the Cholesky cluster size
    int split = 0;                       // lanes the resident batch is split over (0: it lives in this context)
    std::vector<int> split_first;        // first window of every lane (+ end)
    cudaEvent_t start_after = nullptr;   // lane: event of the parent stream the solve starts after
    cudaEvent_t ev_done = nullptr;       // lane: recorded behind the last kernel of a solve
    std::vector<vilba_ctx*> lanes;  // sub-contexts the windows of a large batch are split over
    std::mutex upload_mx;           // lanes flatten + enqueue their H2D one after the other, each with all host threads
    bool upload_exclusive = false;  // (set on a lane while it holds the parent's upload_mx)
    int n_lanes = 4;                // env VILBA_BATCH_LANES
    std::vector<cudaEvent_t> probes;  // 8 events per profiled slot
    size_t probes_used = 0;
    double dbg_ms[4] = {0, 0, 0, 0};
};

namespace {

bool fail(vilba_ctx* c, cudaError_t e, const char* what) {
    if (e == cudaSuccess) return false;
    const LaunchDims& d = c->dims;
    char geo[256];
    std::snprintf(geo, sizeof(geo), " [windows %d, grids %d/%d/%d/%d/%d, schur %d warps x %d sets x %d, chol %d (big %d), smem %zu/%zu/%zu]",
                  d.n_windows, d.point_grid, d.imu_grid, d.gather_grid, d.reduce_grid, d.assemble_grid, d.sp_warps, d.sp_sets, d.sp_grid,
                  d.chol_cluster, d.chol_big_tiles, d.smem_point, d.smem_lin, d.smem_sp);
    c->err = std::string(what) + ": " + cudaGetErrorString(e) + geo;
    return true;
}

#define CK(call, what)                          \
    do {                                        \
        if (fail(ctx, (call), what)) return VILBA_ERR_CUDA; \
    } while (0)

int check_window(const vilba_window* w) {
    if (!w || w->n_kf <= 0 || w->n_imu < 0 || w->n_pts < 0 || w->n_obs < 0) return VILBA_ERR_ARG;
    if (w->n_kf > (OBS_KF_MASK) || w->n_kf > kMaxKF) return VILBA_ERR_ARG - 200;  // (reported as "too many key-frames")
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
    if (!w->n_pts && w->n_obs) return VILBA_ERR_ARG;
    // one pass over the observations: key-frame indices in range, and the observations of a point ordered by
    // key-frame (MapPoint::GetObservations order), each key-frame once
    const int32_t* kf = w->obs_kf;
    const int K = w->n_kf;
    for (int p = 0; p < w->n_pts; ++p) {
        const int e0 = w->pt_obs_begin[p], e1 = w->pt_obs_begin[p + 1];
        if (e1 < e0 || e0 < 0 || e1 > w->n_obs) return VILBA_ERR_ARG;
        int prev = -1;
        for (int e = e0; e < e1; ++e) {
            const int k = kf[e];
            if (k <= prev || k >= K) return VILBA_ERR_ARG;  // also rejects k < 0
            prev = k;
        }
    }
    return VILBA_OK;
}

// profiling: eight events per slot (see launch_slot); drained after a stream synchronize
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
        if (chol > 0.012f) {
            ctx->stats.schur_ms += sch, ctx->stats.schur_launches++;
            ctx->stats.solve_ms += chol, ctx->stats.solve_launches++;
            if (cudaEventElapsedTime(&t, p[2], p[6]) == cudaSuccess) ctx->dbg_ms[1] += t;  // schur_prep alone
            if (cudaEventElapsedTime(&t, p[4], p[5]) == cudaSuccess) {  // update_eval alone
                ctx->dbg_ms[2] += t;
                ctx->stats.update_ms += t, ctx->stats.update_launches++;
            }
        }
    }
    ctx->probes_used = 0;
}

// launch geometry of a batch: the per-window grids shrink as the batch grows so that one launch is about
// one resident wave (2 CTAs of the per-point kernels per SM) whatever the number of windows
LaunchDims choose_dims(const vilba_ctx* ctx, int n_win, int max_ni, int max_K, int max_nf) {
    LaunchDims d = ctx->dims;
    const int sm = ctx->sm_count;
    d.sm_count = sm;
    d.n_windows = n_win;
    d.point_grid = std::max(4, kPointGridPerSM * sm / n_win);
    d.imu_grid = n_win == 1 ? kImuGrid : std::max(1, std::min(kImuGrid, max_ni));
    d.gather_grid = std::max(2, 2 * sm / n_win);
    d.reduce_grid = std::max(2, sm / n_win);
    d.assemble_grid = std::max(4, 8 * sm / n_win);
    // linearize_v2 keeps one private [P,Phi] accumulator per warp and free key-frame in shared memory: windows with
    // many free key-frames (a global BA) run it with fewer warps per CTA.  Sized by the current batch.
    {
        const int K32 = std::max(32, (max_K + 31) / 32 * 32), nf8 = std::max(8, (max_nf + 7) / 8 * 8);
        d.lin_threads = kPointThreads;
        while (d.lin_threads > 32 && linearize_smem_bytes(K32, nf8, d.lin_threads) > 227 * 1024) d.lin_threads /= 2;
        d.smem_lin = linearize_smem_bytes(K32, nf8, d.lin_threads);
    }
    // tile-scan Schur kernel for windows of <= 32 key-frames, else the gather over pair lists
    d.sp_warps = d.sp_sets = d.sp_grid = d.sp_tile_pts = 0;
    d.smem_sp = 0;
    // (the Schur and Cholesky variants follow the CURRENT batch, not the capacities: one large window must not push
    // every later small one onto the paths for large windows)
    if (!ctx->schur_gather_only && max_K > 0 && schur_tile_fits(max_K, max_nf)) {
        const int max_pairs = max_nf * (max_nf + 1) / 2;
        const int max_groups = (max_pairs + 1) / 2;  // a lane group owns two block pairs (a close and a distant one)
        // CTAs of about 7 warps (35 lane groups): measured best for 19 free key-frames when lanes of a batch overlap
        // (profiles/r2_schur_variants.md); a single window keeps the finer split, its CTAs have the machine to themselves
        d.sp_sets = ctx->sp_sets > 0 ? ctx->sp_sets : (n_win > 1 ? std::max(1, (max_groups + 34) / 35) : (max_pairs >= 40 ? 4 : 1));
        d.sp_warps = std::min(16, std::max(1, ((max_groups + d.sp_sets - 1) / d.sp_sets + 4) / 5));
        d.sp_sets = (max_groups + 5 * d.sp_warps - 1) / (5 * d.sp_warps);  // every group must have its lanes
        d.sp_grid = std::max(1, std::min(ctx->sp_grid_cap, 2 * sm / (d.sp_sets * n_win)));
        d.sp_tile_pts = std::max(4, std::min(32, (int)(46 * 1024 / (144 * (size_t)std::min(32, max_K)))));
        d.sp_pair_lanes = 0;
        if (ctx->sp_pair_lanes && schur_pair_lanes(max_nf) <= 384) {
            d.sp_pair_lanes = 1;
            d.sp_sets = 1;
            d.sp_warps = (schur_pair_lanes(max_nf) + 31) / 32;
            d.sp_grid = std::max(1, std::min(ctx->sp_grid_cap, 2 * sm / n_win));
        }
        if (const char* e = std::getenv("VILBA_SP_PSUB")) d.sp_grid = std::max(1, std::atoi(e));
        if (const char* e = std::getenv("VILBA_SP_TILE")) d.sp_tile_pts = std::max(2, std::min(d.sp_tile_pts, std::atoi(e)));
        d.smem_sp = schur_tile_smem_bytes(max_K, d.sp_tile_pts);
        d.sp_mma = 0, d.sp_tile_edges = 0;
        if (ctx->sp_mma && !d.sp_pair_lanes && ctx->tile_edges[0] > 0) {
            // tensor-pipe kernel: the largest tile (32, 16 or 8 map points) whose buffers leave room for two CTAs per SM;
            // the warps that cover all block pairs are cut into `sets` CTAs of at most 16
            const int tp[3] = {32, 16, 8};
            int pick = 0;
            while (pick < 2 && schur_mma_smem_bytes(ctx->tile_edges[pick], tp[pick]) > 110 * 1024) ++pick;
            if (const char* e = std::getenv("VILBA_SP_TILE")) pick = std::atoi(e) >= 32 ? 0 : (std::atoi(e) >= 16 ? 1 : 2);
            if (schur_mma_smem_bytes(ctx->tile_edges[pick], tp[pick]) <= 227 * 1024) {
                const int warps_all = schur_mma_units(max_nf);
                int sets = ctx->sp_sets > 0 ? ctx->sp_sets : std::max(1, (warps_all + 15) / 16);
                sets = std::max(sets, (warps_all + 15) / 16);
                d.sp_mma = 1;
                d.sp_sets = sets;
                d.sp_warps = (warps_all + sets - 1) / sets;
                d.sp_tile_pts = tp[pick];
                d.sp_tile_edges = ctx->tile_edges[pick];
                d.sp_grid = std::max(1, std::min(ctx->sp_grid_cap, 2 * sm / (d.sp_sets * n_win)));
                if (const char* e = std::getenv("VILBA_SP_PSUB")) d.sp_grid = std::max(1, std::atoi(e));
                d.smem_sp = schur_mma_smem_bytes(d.sp_tile_edges, d.sp_tile_pts);
            }
        }
    }
    // Reduced system.  chol_la.cu (one cluster per window, the trailing matrix in the shared memory of the cluster) takes
    // every system that fits (n <= 345 with 8 CTAs; n = 285 needs 4, n = 135 one): the full cluster for one window
    // (latency), the smallest cluster that fits for a batch (the windows of all lanes share the SMs).  Larger systems, or
    // env VILBA_CHOL_LA=0: the multi-kernel blocked factorisation of chol_big.cu.  Both are LDL^T with the failure rule of
    // Eigen's SimplicialLDLT (a zero or non-finite pivot fails the trial, a negative one does not).
    const int n_cur = 15 * max_nf;
    d.chol_n = n_cur;
    d.chol_la = 0, d.chol_big_tiles = 0;
    d.chol_cluster = ctx->chol_cluster;
    if (max_nf > 0) {
        int c = n_win == 1 ? ctx->chol_cluster : 1;
        while (c < 8 && !chol_la_fits(n_cur, c)) c *= 2;
        if (ctx->chol_la_mode != 0 && n_cur <= ctx->chol_big_above && chol_la_fits(n_cur, c)) {
            d.chol_la = 1;
            d.chol_cluster = c;
        } else {
            d.chol_big_tiles = (n_cur + 63) / 64;
        }
    }
    return d;
}

void drop_graphs(vilba_ctx* ctx) {
    for (GraphEntry& g : ctx->graphs) cudaGraphExecDestroy(g.exec);
    ctx->graphs.clear();
}

// flatten one window into its slice of the pinned staging buffer (phase A/B of the reference function:
// gather + graph build, Optimizer.cpp:2329-2639, become "pack + one H2D")
void pack_window(const vilba_window* w, const WinMeta& m, char* h) {
    const Layout& L = m.L;
    const int K = m.K, NI = m.NI, P = m.P, E = m.E, n_free = m.n_free, n_pairs = m.n_pairs;
    std::memcpy(h + L.kf_state0, w->kf_state, sizeof(double) * 22 * (size_t)K);
    if (P) std::memcpy(h + L.pts0, w->pt_xyz, sizeof(double) * 3 * (size_t)P);
    if (NI) std::memcpy(h + L.imu_preint, w->imu_preint, sizeof(double) * 142 * (size_t)NI);
    // the observations travel as the caller's three arrays (plain copies: the host threads of 8 processes x 4 lanes share
    // the cores); reset_kernel interleaves them into the 16-byte edge records at the start of every solve
    if (E) {
        char* ho = h + L.obs0;
        std::memcpy(ho, w->obs_uv, 8 * (size_t)E);
        std::memcpy(ho + 8 * (size_t)E, w->obs_inv_sigma2, 4 * (size_t)E);
        std::memcpy(ho + 12 * (size_t)E, w->obs_kf, 4 * (size_t)E);
    }
    if (P) std::memcpy(h + L.pt_obs_begin, w->pt_obs_begin, sizeof(int) * ((size_t)P + 1));
    else std::memset(h + L.pt_obs_begin, 0, sizeof(int));
    int* kf_block = reinterpret_cast<int*>(h + L.kf_block);
    int* blk_kf = reinterpret_cast<int*>(h + L.blk_kf);
    int nf = 0;
    for (int k = 0; k < K; ++k) {
        kf_block[k] = (w->kf_flags[k] & VILBA_KF_FIXED) ? -1 : nf++;
        if (kf_block[k] >= 0) blk_kf[kf_block[k]] = k;
    }
    if (NI) {
        std::memcpy(h + L.imu_i, w->imu_kf_i, sizeof(int) * (size_t)NI);
        std::memcpy(h + L.imu_j, w->imu_kf_j, sizeof(int) * (size_t)NI);
    }
    // IMU edge of every block (as key-frame i / as key-frame j) and the key-frame block pairs (a <= b)
    int* blk_edge_i = reinterpret_cast<int*>(h + L.blk_edge_i);
    int* blk_edge_j = reinterpret_cast<int*>(h + L.blk_edge_j);
    for (int b = 0; b < n_free; ++b) blk_edge_i[b] = blk_edge_j[b] = -1;
    for (int e = 0; e < NI; ++e) {
        const int bi = kf_block[w->imu_kf_i[e]], bj = kf_block[w->imu_kf_j[e]];
        if (bi >= 0) blk_edge_i[bi] = e;
        if (bj >= 0) blk_edge_j[bj] = e;
    }
    int* pair_a = reinterpret_cast<int*>(h + L.pair_a);
    int* pair_b = reinterpret_cast<int*>(h + L.pair_b);
    int t = 0;
    for (int a = 0; a < n_free; ++a)
        for (int b = a; b < n_free; ++b, ++t) pair_a[t] = a, pair_b[t] = b;
    (void)n_pairs;
}

void fill_dev_window(const vilba_ctx* ctx, const vilba_window* w, const WinMeta& m, char* d, LmState* lm, DevWindow& dw) {
    const Layout& L = m.L;
    char* in = d + m.in_base;
    char* out = d + m.out_base;
    char* wk = d + m.wk_base;
    std::memset(&dw, 0, sizeof(dw));
    dw.K = m.K, dw.NI = m.NI, dw.P = m.P, dw.E = m.E, dw.n_free = m.n_free, dw.n = m.n;
    dw.lds = (m.n + 3) & ~3;
    for (int b = 0; b < 2; ++b) {
        dw.kf_state[b] = reinterpret_cast<double*>(wk + L.kf_state[b]);
        dw.pts[b] = reinterpret_cast<double*>(wk + L.pts[b]);
    }
    dw.kf_state0 = reinterpret_cast<const double*>(in + L.kf_state0);
    dw.pts0 = reinterpret_cast<const double*>(in + L.pts0);
    dw.obs0 = in + L.obs0;
    // every mono edge starts with its Huber kernel (Optimizer.cpp:2622-2624) unless bRobust is false (:1590-1595)
    dw.obs_flags0 = (ctx->prm.mode & VILBA_MODE_MONO_NOT_ROBUST) ? 0 : OBS_ROBUST;
    dw.outlier = reinterpret_cast<uint8_t*>(wk + L.outlier);
    dw.out_kf_state = reinterpret_cast<double*>(out + L.o_kf);
    dw.out_pts = reinterpret_cast<double*>(out + L.o_pts);
    dw.out_chi2 = reinterpret_cast<double*>(d + m.chi_base);
    dw.out_outlier = reinterpret_cast<uint8_t*>(out + L.o_outlier);
    dw.kf_block = reinterpret_cast<const int*>(in + L.kf_block);
    dw.imu_i = reinterpret_cast<const int*>(in + L.imu_i);
    dw.imu_j = reinterpret_cast<const int*>(in + L.imu_j);
    dw.imu_preint = reinterpret_cast<const double*>(in + L.imu_preint);
    dw.imu_info = reinterpret_cast<double*>(wk + L.imu_info);
    dw.imu_err = reinterpret_cast<double*>(wk + L.imu_err);
    dw.pt_obs_begin = reinterpret_cast<const int*>(in + L.pt_obs_begin);
    dw.obs = reinterpret_cast<int4*>(wk + L.obs);
    dw.obs_chi2 = reinterpret_cast<double*>(wk + L.obs_chi2);
    dw.Hpp = reinterpret_cast<double*>(wk + L.Hpp);
    dw.bp = reinterpret_cast<double*>(wk + L.bp);
    const bool sharded = ctx->comm != nullptr;
    dw.Hpp_w = dw.Hpp;  // (sharded: this rank's partial sums stay on this rank)
    dw.bp_w = dw.bp;
    dw.shard_owner = (!sharded || ctx->comm_rank == 0) ? 1 : 0;
    dw.sharded = sharded ? 1 : 0;
    dw.shard_rank = ctx->comm_rank, dw.shard_world = ctx->comm_world;
    dw.diag_red = reinterpret_cast<double*>(wk + L.diag_red);
    dw.Hll = reinterpret_cast<double*>(wk + L.Hll);
    dw.bl = reinterpret_cast<double*>(wk + L.bl);
    dw.W = reinterpret_cast<double*>(wk + L.W);
    dw.lin_partial = reinterpret_cast<double*>(wk + L.lin_partial);
    dw.imu_slot = reinterpret_cast<double*>(wk + L.imu_slot);
    dw.mono_sum = reinterpret_cast<double*>(wk + L.mono_sum);
    dw.Y = reinterpret_cast<double*>(wk + L.Y);
    dw.schur_partial = reinterpret_cast<double*>(wk + L.schur_partial);
    dw.ts_rec = reinterpret_cast<double*>(wk + L.ts_rec);
    dw.ts_hdr = reinterpret_cast<unsigned*>(wk + L.ts_hdr);
    dw.blk_edge_i = reinterpret_cast<const int*>(in + L.blk_edge_i);
    dw.blk_edge_j = reinterpret_cast<const int*>(in + L.blk_edge_j);
    dw.edge_pt = reinterpret_cast<const int*>(wk + L.edge_pt);
    dw.n_pairs = m.n_pairs;
    dw.pair_a = reinterpret_cast<const int*>(in + L.pair_a);
    dw.pair_b = reinterpret_cast<const int*>(in + L.pair_b);
    dw.pair_begin = reinterpret_cast<const int*>(wk + L.pair_begin);
    dw.pair_ea = reinterpret_cast<const int*>(wk + L.pair_ea);
    dw.pair_eb = reinterpret_cast<const int*>(wk + L.pair_eb);
    dw.blk_kf = reinterpret_cast<const int*>(in + L.blk_kf);
    dw.edge_pt_rw = reinterpret_cast<int*>(wk + L.edge_pt);
    dw.pair_begin_rw = reinterpret_cast<int*>(wk + L.pair_begin);
    dw.pair_ea_rw = reinterpret_cast<int*>(wk + L.pair_ea);
    dw.pair_eb_rw = reinterpret_cast<int*>(wk + L.pair_eb);
    dw.pt_mask = reinterpret_cast<unsigned long long*>(wk + L.pt_mask);
    dw.S = reinterpret_cast<double*>(wk + L.S);
    dw.Lfac = reinterpret_cast<double*>(wk + L.Lfac);
    dw.cminv = reinterpret_cast<double*>(wk + L.cminv);
    dw.cdinv = reinterpret_cast<double*>(wk + L.cdinv);
    dw.bs = reinterpret_cast<double*>(wk + L.bs);
    dw.S_w = sharded ? reinterpret_cast<double*>(wk + L.S_part) : dw.S;
    dw.bs_w = sharded ? reinterpret_cast<double*>(wk + L.S_part + (L.bs - L.S)) : dw.bs;
    dw.x = reinterpret_cast<double*>(wk + L.x);
    dw.lm = lm;
    dw.dbg = reinterpret_cast<long long*>(wk + L.dbg);
    dw.chi_partial = reinterpret_cast<double*>(wk + L.chi_partial);
    dw.chi_counter = reinterpret_cast<unsigned*>(wk + L.chi_counter);
    if (const char* e = std::getenv("VILBA_TS_ABLATE")) dw.dbg_flags = std::atoi(e);
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
}

// Host threads one context may use for flatten / scatter.  Several processes share the host (one per GPU under torchrun:
// LOCAL_WORLD_SIZE) and a split batch runs one thread per lane on top, so the budget is cores / (processes * lanes):
// 8 processes x 4 lanes x 8 flatten threads on a 32-core host is what cost the 8-GPU end-to-end run 9 % in round 1.
int host_thread_budget(int lanes) {
    static const int fixed = std::getenv("VILBA_HOST_THREADS") ? std::max(1, std::atoi(std::getenv("VILBA_HOST_THREADS"))) : 0;
    if (fixed) return fixed;
    static const int procs = std::getenv("LOCAL_WORLD_SIZE") ? std::max(1, std::atoi(std::getenv("LOCAL_WORLD_SIZE"))) : 1;
    const int hw = (int)std::max(1u, std::thread::hardware_concurrency());
    return std::max(1, std::min(8, hw / (procs * std::max(1, lanes))));
}

template <class F>
void parallel_for(int n, int max_threads, F&& f) {
    const int nt = std::max(1, std::min(n, max_threads));
    if (nt == 1) {
        for (int i = 0; i < n; ++i) f(i);
        return;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t)
        th.emplace_back([&, t]() {
            for (int i = t; i < n; i += nt) f(i);
        });
    for (auto& x : th) x.join();
}

// ------------------------------------------------------------------------------------------------
// point-sharded window: the three exchanges of one LM slot, all on the library stream: a small one after
// the linearisation (what computeLambdaInit needs), S | b_s after the Schur step, chi2 | scale after the update.  H_pp
// and b_p never cross the wire: every rank folds its own partial sums into its partial S.  The reductions are enqueued
// unconditionally (the device-side controller skips the kernels); S | b_s goes from a send buffer only this rank's
// kernels write, the two small ones are in place behind kernels that rewrite their inputs only in the phase that
// uses them, so repeating them in a slot whose phase does not need them is harmless.
// ------------------------------------------------------------------------------------------------
void* open_nccl() {
    static void* lib = nullptr;
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);  // the copy torch already loaded, if any
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    return lib;
}

cudaError_t slot_reduce(void* self, int which, cudaStream_t s) {
    vilba_ctx* ctx = static_cast<vilba_ctx*>(self);
    const DevWindow& dw = ctx->dw[0];
    const Layout& L = ctx->meta[0].L;
    ncclResult_t r = ncclSuccess;
    switch (which) {
        case RED_DIAG:  // diag(H_pp) summed | every rank's max |diag H_ll| in its own slot (zeros elsewhere): one sum
            r = ctx->p_ncclAllReduce(dw.diag_red, dw.diag_red, (size_t)dw.n + (size_t)ctx->comm_world, ncclDouble, ncclSum, ctx->comm, s);
            break;
        case RED_S:
            r = ctx->p_ncclAllReduce(dw.S_w, dw.S, L.s_span / sizeof(double), ncclDouble, ncclSum, ctx->comm, s);
            break;
        case RED_CHI:  // chi_acc and scale_acc are adjacent
            r = ctx->p_ncclAllReduce(&dw.lm->chi_acc, &dw.lm->chi_acc, 2, ncclDouble, ncclSum, ctx->comm, s);
            break;
    }
    if (r != ncclSuccess) {
        ctx->err = std::string("ncclAllReduce: ") + (ctx->p_ncclGetErrorString ? ctx->p_ncclGetErrorString(r) : "?");
        return cudaErrorUnknown;
    }
    return cudaSuccess;
}

// ------------------------------------------------------------------------------------------------
// flatten + upload a batch of windows
// ------------------------------------------------------------------------------------------------
int upload_batch(vilba_ctx* ctx, int n_win, const vilba_window* wins) {
    const auto t_begin = std::chrono::steady_clock::now();
    ctx->n_win = 0;
    if (n_win <= 0 || n_win > ctx->max_batch || !wins) {
        ctx->err = "invalid batch size";
        return VILBA_ERR_ARG;
    }
    if (ctx->comm && n_win != 1) {
        ctx->err = "a sharded context solves one window at a time";
        return VILBA_ERR_ARG;
    }
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    // a captured sharded slot also holds the allreduces of slot_reduce(): their addresses and counts follow this window's
    // layout (P, E, n_free, arena base), not only the launch geometry the graph cache is keyed on -- re-capture per window
    if (ctx->comm) drop_graphs(ctx);
    std::vector<WinMeta>& meta = ctx->meta;
    meta.assign(n_win, WinMeta());
    std::vector<int> status(n_win, VILBA_OK);
    const int host_threads = host_thread_budget(ctx->batch_total_hint > n_win && !ctx->upload_exclusive ? ctx->n_lanes : 1);
    parallel_for(n_win, host_threads, [&](int i) {
        const vilba_window* w = &wins[i];
        int st = check_window(w);
        if (st != VILBA_OK) {
            status[i] = st;
            return;
        }
        WinMeta& m = meta[i];
        m.K = w->n_kf, m.NI = w->n_imu, m.P = w->n_pts, m.E = w->n_obs;
        std::vector<int> kf_block(m.K, -1);
        for (int k = 0; k < m.K; ++k)
            if (!(w->kf_flags[k] & VILBA_KF_FIXED)) kf_block[k] = m.n_free++;
        m.n = 15 * m.n_free;
        m.n_pairs = m.n_free * (m.n_free + 1) / 2;
        std::vector<char> is_i(m.n_free, 0), is_j(m.n_free, 0);
        for (int e = 0; e < m.NI; ++e) {
            const int bi = kf_block[w->imu_kf_i[e]], bj = kf_block[w->imu_kf_j[e]];
            if ((bi >= 0 && is_i[bi]) || (bj >= 0 && is_j[bj]) || (bi >= 0 && bi == bj)) {
                status[i] = VILBA_ERR_ARG - 100;  // a key-frame may start / end at most one IMU edge
                return;
            }
            if (bi >= 0) is_i[bi] = 1;
            if (bj >= 0) is_j[bj] = 1;
        }
        // the (edge_a, edge_b) lists of the Schur gather are built on the device (pairs.cu); capacity = sum m (m + 1) / 2
        for (int p = 0; p < m.P; ++p) {
            const size_t mm = (size_t)(w->pt_obs_begin[p + 1] - w->pt_obs_begin[p]);
            m.n_triples += mm * (mm + 1) / 2;
        }
        // most edges in a tile of 32 / 16 / 8 consecutive map points (sizes the tiles of the column-slot Schur kernel)
        for (int p = 0; p < m.P; p += 8) {
            const int* ob = w->pt_obs_begin;
            m.tile_edges[2] = std::max(m.tile_edges[2], ob[std::min(p + 8, m.P)] - ob[p]);
            if (p % 16 == 0) m.tile_edges[1] = std::max(m.tile_edges[1], ob[std::min(p + 16, m.P)] - ob[p]);
            if (p % 32 == 0) m.tile_edges[0] = std::max(m.tile_edges[0], ob[std::min(p + 32, m.P)] - ob[p]);
        }
    });
    int max_K = 0, max_nf = 0, max_ni = 0;
    for (int i = 0; i < n_win; ++i) {
        if (status[i] != VILBA_OK) {
            ctx->err = status[i] == VILBA_ERR_ARG - 100   ? "a key-frame may start / end at most one IMU edge"
                       : status[i] == VILBA_ERR_ARG - 200 ? "more than VILBA_MAX_KEYFRAMES (256) key-frames in one window"
                                                          : "invalid window";
            return VILBA_ERR_ARG;
        }
        max_K = std::max(max_K, meta[i].K), max_nf = std::max(max_nf, meta[i].n_free), max_ni = std::max(max_ni, meta[i].NI);
    }
    for (int k = 0; k < 3; ++k) {
        int te = 1;
        for (int i = 0; i < n_win; ++i) te = std::max(te, meta[i].tile_edges[k]);
        ctx->tile_edges[k] = (te + 15) / 16 * 16;  // rounded: windows of similar shape share a captured graph
    }
    // shared-memory capacities grow monotonically; the captured graphs depend on them
    if (max_K > ctx->cap_K || max_nf > ctx->cap_nf) {
        ctx->cap_K = std::max(ctx->cap_K, (max_K + 31) / 32 * 32);
        ctx->cap_nf = std::max(ctx->cap_nf, (max_nf + 7) / 8 * 8);
        ctx->cap_n = std::max(ctx->cap_n, 15 * ctx->cap_nf);
        ctx->dims.smem_point = point_smem_bytes(ctx->cap_K);
        ctx->dims.chol_cluster = ctx->chol_cluster;
        CK(configure_kernels(ctx->dims), "cudaFuncSetAttribute");
        drop_graphs(ctx);
    }
    ctx->dims = choose_dims(ctx, n_win, max_ni, max_K, max_nf);
    // arena: [inputs of all windows | outputs of all windows | LmState array | work regions]
    size_t in_o = 0, out_o = 0, chi_o = 0, wk_o = 0;
    for (int i = 0; i < n_win; ++i) {
        WinMeta& m = meta[i];
        m.L = make_layout(m, ctx->dims.point_grid, ctx->dims.sp_grid, std::max(1, ctx->dims.sp_tile_pts), ctx->comm != nullptr);
        m.in_base = in_o, in_o += m.L.in_bytes;
        m.out_base = out_o, out_o += m.L.out_bytes;
        m.chi_base = chi_o, chi_o += m.L.chi_bytes;
        m.wk_base = wk_o, wk_o += m.L.wk_bytes;
    }
    const size_t lm_bytes = align_up(sizeof(LmState) * (size_t)n_win);
    ctx->in_total = in_o, ctx->out_small = out_o, ctx->out_total = out_o + chi_o;
    ctx->out_region = in_o;
    ctx->lm_base = in_o + ctx->out_total;
    const size_t wk_region = ctx->lm_base + lm_bytes;
    for (int i = 0; i < n_win; ++i)
        meta[i].out_base += ctx->out_region, meta[i].chi_base += ctx->out_region + out_o, meta[i].wk_base += wk_region;
    CK(ctx->arena.reserve(wk_region + wk_o), "cudaMalloc(arena)");
    CK(ctx->pinned.reserve(in_o), "cudaMallocHost(staging)");
    CK(ctx->pinned_out.reserve(ctx->out_total + lm_bytes), "cudaMallocHost(results)");
    CK(ctx->pinned_small.reserve(sizeof(DevWindow) * (size_t)n_win + 256), "cudaMallocHost(desc)");
    char* h = ctx->pinned.base;
    char* d = ctx->arena.base;
    parallel_for(n_win, host_threads, [&](int i) { pack_window(&wins[i], meta[i], h + meta[i].in_base); });
    const auto t_packed = std::chrono::steady_clock::now();
    CK(cudaMemcpyAsync(d, h, in_o, cudaMemcpyHostToDevice, ctx->stream), "H2D windows");
    if (std::getenv("VILBA_DEBUG_COUNTERS"))
        std::fprintf(stderr, "[vilba dbg] flatten %d windows %.3f ms, %zu bytes H2D\n", n_win,
                     std::chrono::duration<double, std::milli>(t_packed - t_begin).count(), in_o);
    ctx->dw.resize(n_win);
    LmState* lm0 = reinterpret_cast<LmState*>(d + ctx->lm_base);
    for (int i = 0; i < n_win; ++i) {
        fill_dev_window(ctx, &wins[i], meta[i], d, lm0 + i, ctx->dw[i]);
    }
    std::memcpy(ctx->pinned_small.base, ctx->dw.data(), sizeof(DevWindow) * (size_t)n_win);
    CK(cudaMemcpyAsync(ctx->dwp, ctx->pinned_small.base, sizeof(DevWindow) * (size_t)n_win, cudaMemcpyHostToDevice,
                       ctx->stream), "H2D descriptors");
    // working copies of the estimates / observation table, then the Schur pair lists (they read the table)
    CK(launch_reset(ctx->stream, ctx->dwp, ctx->dims), "reset");
    if (ctx->dims.sp_warps == 0) {
        CK(launch_build_pair_lists(ctx->stream, ctx->dwp, ctx->dims), "pair lists");
        ctx->stats.kernel_launches += 4;
    }
    ctx->stats.kernel_launches += 1;
    ctx->n_win = n_win;
    return VILBA_OK;
}

// D2H of the controller states; waits for the stream while polling the caller's stop flag
int read_lm(vilba_ctx* ctx, std::vector<LmState>& lm, const volatile uint8_t* stop_flag) {
    char* hp = ctx->pinned_out.base + ctx->out_total;
    const size_t bytes = sizeof(LmState) * (size_t)ctx->n_win;
    CK(cudaMemcpyAsync(hp, ctx->arena.base + ctx->lm_base, bytes, cudaMemcpyDeviceToHost, ctx->stream), "D2H lm");
    if (stop_flag && !ctx->comm) {  // a sharded solve cannot be interrupted half way: the ranks would diverge
        bool sent = false;
        while (cudaStreamQuery(ctx->stream) == cudaErrorNotReady) {
            if (!sent && *stop_flag) {  // mirror of `bool* pbStopFlag` (sparse_optimizer.h:188)
                static const int one = 1;
                for (int i = 0; i < ctx->n_win; ++i)
                    cudaMemcpyAsync(&ctx->dw[i].lm->stop, &one, sizeof(int), cudaMemcpyHostToDevice, ctx->stream2);
                sent = true;
            }
        }
    }
    CK(cudaStreamSynchronize(ctx->stream), "sync");
    lm.resize(ctx->n_win);
    std::memcpy(lm.data(), hp, bytes);
    probe_drain(ctx);
    return VILBA_OK;
}

bool stop_requested(const volatile uint8_t* f) { return f && *f; }

// field-wise (the struct has padding bytes, memcmp would compare them)
bool same_dims(const LaunchDims& a, const LaunchDims& b) {
    return a.sm_count == b.sm_count && a.n_windows == b.n_windows && a.point_grid == b.point_grid && a.imu_grid == b.imu_grid &&
           a.gather_grid == b.gather_grid && a.reduce_grid == b.reduce_grid && a.assemble_grid == b.assemble_grid &&
           a.sp_warps == b.sp_warps && a.sp_sets == b.sp_sets && a.sp_grid == b.sp_grid && a.sp_tile_pts == b.sp_tile_pts &&
           a.sp_pair_lanes == b.sp_pair_lanes && a.sp_mma == b.sp_mma && a.sp_tile_edges == b.sp_tile_edges && a.chol_cluster == b.chol_cluster && a.chol_big_tiles == b.chol_big_tiles &&
           a.chol_la == b.chol_la && a.chol_n == b.chol_n && a.smem_point == b.smem_point && a.smem_lin == b.smem_lin && a.lin_threads == b.lin_threads &&
           a.smem_sp == b.smem_sp;
}

int ensure_graph(vilba_ctx* ctx, cudaGraphExec_t* out) {
    for (const GraphEntry& g : ctx->graphs)
        if (same_dims(g.dims, ctx->dims)) {
            *out = g.exec;
            return VILBA_OK;
        }
    cudaGraph_t g = nullptr;
    CK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal), "begin capture");
    cudaError_t e = launch_slot(ctx->stream, ctx->stream2, ctx->ev_fork, ctx->ev_join, ctx->dwp, ctx->dims, nullptr,
                                ctx->comm ? &ctx->slot_comm : nullptr);  // NCCL collectives are capturable
    cudaError_t e2 = cudaStreamEndCapture(ctx->stream, &g);
    if (fail(ctx, e, "capture slot") || fail(ctx, e2, "end capture")) return VILBA_ERR_CUDA;
    GraphEntry ge;
    std::memset(&ge, 0, sizeof(ge));
    ge.dims = ctx->dims;
    e = cudaGraphInstantiate(&ge.exec, g, 0);
    cudaGraphDestroy(g);
    if (fail(ctx, e, "graph instantiate")) return VILBA_ERR_CUDA;
    ctx->graphs.push_back(ge);
    *out = ge.exec;
    return VILBA_OK;
}

// SparseOptimizer::optimize(iterations) with OptimizationAlgorithmLevenberg (sparse_optimizer.cpp:354-419).
// The loop itself runs on the device: the host enqueues slots (CUDA graph launches) without
// synchronising and looks at the controller states once they have drained.
int run_stage(vilba_ctx* ctx, int stage, int iterations, vilba_result* out, const volatile uint8_t* stop_flag,
              std::vector<LmState>& lm) {
    cudaStream_t s = ctx->stream;
    vilba_stats& stt = ctx->stats;
    // computeActiveErrors + activeRobustChi2 at the first iteration; later iterations inherit currentChi
    // of the accepted trial (identical by construction: the errors are those of the accepted state)
    CK(launch_eval_initial(s, ctx->dwp, ctx->dims), "eval");
    const SlotComm* comm = ctx->comm ? &ctx->slot_comm : nullptr;
    if (comm && slot_reduce(ctx, RED_CHI, s) != cudaSuccess) return VILBA_ERR_COMM;
    CK(launch_stage_begin(s, ctx->dwp, ctx->dims, stage, iterations), "stage_begin");
    stt.kernel_launches += 2;
    const bool graph = ctx->use_graph && !ctx->profiling && (!comm || ctx->comm_graph);
    cudaGraphExec_t exec = nullptr;
    if (graph) {
        int r = ensure_graph(ctx, &exec);
        if (r != VILBA_OK) return r;
    }
    std::vector<int> first_trace(ctx->n_win);
    for (int i = 0; i < ctx->n_win; ++i) first_trace[i] = lm[i].n_trace;
    // one trial per iteration when every step is accepted.  A single window gets one blind spare slot (a rejected trial then
    // costs no host round trip; an unused slot is ~30 us of early-exit launches); a batch gets none: idle slots of one lane
    // cost the other lanes' kernels their SMs (measured: 0 / 1 / 2 / 3 spares = 41.4 / 41.1 / 41.1 / 40.8 k LM iterations/s),
    // and while a lane waits for the host the others run
    static const int spare_env = std::getenv("VILBA_SPARE_SLOTS") ? std::max(0, std::atoi(std::getenv("VILBA_SPARE_SLOTS"))) : -1;
    int slots = iterations + (spare_env >= 0 ? spare_env : (ctx->n_win > 1 ? 0 : 1));
    // every slot runs at least one trial of every unfinished window, so (iterations + 1) * max_trials slots always suffice
    const int max_rounds = 2 + ((iterations + 1) * std::max(1, ctx->prm.max_trials)) / 2;
    bool done = false;
    for (int round = 0; round < max_rounds && !done; ++round) {
        for (int i = 0; i < slots; ++i) {
            if (graph)
                CK(cudaGraphLaunch(exec, s), "graph launch");
            else
                CK(launch_slot(s, ctx->stream2, ctx->ev_fork, ctx->ev_join, ctx->dwp, ctx->dims, probe_take(ctx), comm), "slot");
            stt.kernel_launches += kernels_per_slot(ctx->dims, comm != nullptr);
        }
        int r = read_lm(ctx, lm, stop_flag);
        if (r != VILBA_OK) return r;
        done = true;
        for (int i = 0; i < ctx->n_win; ++i) done = done && lm[i].phase == PH_DONE;
        slots = 2;  // rejected trials used up the slack: keep going
    }
    if (!done) {
        ctx->err = "LM stage did not finish within iterations * max_trials slots";
        return VILBA_ERR_CUDA;
    }
    for (int wdx = 0; wdx < ctx->n_win; ++wdx) {
        vilba_result& o = out[wdx];
        for (int i = first_trace[wdx]; i < lm[wdx].n_trace && o.n_trace < VILBA_MAX_TRACE; ++i) {
            const IterRec& t = lm[wdx].trace[i];
            vilba_iter_record& rec = o.trace[o.n_trace++];
            rec.stage = t.stage, rec.iteration = t.iteration, rec.trials = t.trials, rec.result = t.result;
            rec.n_active_edges = t.n_active, rec.accepted = t.accepted;
            rec.chi2_initial = t.chi0, rec.chi2_final = t.chi1, rec.lambda = t.lambda, rec.lambda_first_trial = t.lambda_first;
            stt.lm_iterations++;
            stt.lm_trials += t.trials;
            stt.edges_linearized += t.n_active;
        }
    }
    return VILBA_OK;
}

void debug_counters(vilba_ctx* ctx) {
    if (!std::getenv("VILBA_DEBUG_COUNTERS") || ctx->n_win < 1) return;
    if (ctx->stats.solve_launches > 0)
        std::fprintf(stderr, "[vilba dbg] per launch: linearize_v2 %.1f us, schur_prep %.1f us, update_eval %.1f us\n",
                     1e3 * ctx->dbg_ms[0] / std::max<long long>(1, ctx->stats.linearize_launches),
                     1e3 * ctx->dbg_ms[1] / ctx->stats.solve_launches, 1e3 * ctx->dbg_ms[2] / ctx->stats.solve_launches);
    long long h[16];
    if (cudaMemcpy(h, ctx->dw[0].dbg, sizeof(h), cudaMemcpyDeviceToHost) == cudaSuccess && h[8] > 0) {
        std::fprintf(stderr, "[vilba dbg] chol calls=%lld avg cycles/phase:", h[8]);
        for (int i = 0; i < 8; ++i) std::fprintf(stderr, " %lld", h[i] / h[8]);
        std::fprintf(stderr, "\n");
    }
    cudaMemset(ctx->dw[0].dbg, 0, sizeof(long long) * 16);
}

// phases C..E for every window of the resident batch; out[i] receives status / trace / timing
int solve_batch(vilba_ctx* ctx, vilba_result* out, const volatile uint8_t* stop_flag) {
    if (ctx->n_win <= 0) {
        ctx->err = "no window uploaded";
        return VILBA_ERR_ARG;
    }
    const int nw = ctx->n_win;
    for (int i = 0; i < nw; ++i) {
        out[i].n_trace = 0;
        out[i].stage2_ran = 0;
        out[i].n_outliers_stage1 = 0;
        out[i].solve_ms = 0.0;
        out[i].status = VILBA_OK;
    }
    // VILBA_MODE_SINGLE_STAGE (GlobalBundleAdjustmentNavState): one optimize(), which g2o ends after zero iterations
    // when the flag is already up (sparse_optimizer.cpp:376), and the estimates are written back all the same
    const bool single_stage = (ctx->prm.mode & VILBA_MODE_SINGLE_STAGE) != 0;
    if (!single_stage && stop_requested(stop_flag)) {  // Optimizer.cpp:2643-2645
        for (int i = 0; i < nw; ++i) out[i].status = VILBA_ABORTED;
        return VILBA_ABORTED;
    }
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    cudaStream_t s = ctx->stream;
    debug_counters(ctx);
    if (ctx->start_after) CK(cudaStreamWaitEvent(s, ctx->start_after, 0), "wait start");
    CK(launch_reset(s, ctx->dwp, ctx->dims), "reset");  // every solve restarts from the uploaded state
    CK(cudaEventRecord(ctx->ev_a, s), "event");
    CK(launch_imu_prepare(s, ctx->dwp, ctx->dims), "imu_prepare");
    ctx->stats.kernel_launches += 2;
    std::vector<LmState> lm(nw);
    std::memset(lm.data(), 0, sizeof(LmState) * (size_t)nw);
    int r = VILBA_OK;
    if (!stop_requested(stop_flag) && ctx->prm.iters_stage1 > 0) r = run_stage(ctx, 1, ctx->prm.iters_stage1, out, stop_flag, lm);
    if (r != VILBA_OK) return r;
    bool stopped = stop_requested(stop_flag);
    for (int i = 0; i < nw; ++i) stopped = stopped || lm[i].stop;
    if (!stopped && !single_stage) {  // bDoMore (Optimizer.cpp:2650-2656)
        CK(launch_cull(s, ctx->dwp, ctx->dims), "cull");
        ctx->stats.kernel_launches += 1;
        r = run_stage(ctx, 2, ctx->prm.iters_stage2, out, stop_flag, lm);
        if (r != VILBA_OK) return r;
        for (int i = 0; i < nw; ++i) {
            out[i].n_outliers_stage1 = lm[i].n_culled;
            out[i].stage2_ran = 1;
        }
    }
    CK(launch_final_flags(s, ctx->dwp, ctx->dims), "final_flags");
    ctx->stats.kernel_launches += 1;
    CK(cudaEventRecord(ctx->ev_b, s), "event");
    if (ctx->ev_done) CK(cudaEventRecord(ctx->ev_done, s), "event");
    CK(cudaStreamSynchronize(s), "sync");
    probe_drain(ctx);
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b), "elapsed");
    for (int i = 0; i < nw; ++i) out[i].solve_ms = ms;  // device time of the whole batch
    return VILBA_OK;
}

// pack the results of every window on the device, one D2H, scatter into the caller's arrays
int download_batch(vilba_ctx* ctx, vilba_result* out) {
    if (ctx->n_win <= 0) return VILBA_ERR_ARG;
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");  // lanes run on their own host threads
    cudaStream_t s = ctx->stream;
    CK(launch_export(s, ctx->dwp, ctx->dims), "export");
    ctx->stats.kernel_launches += 1;
    char* hp = ctx->pinned_out.base;
    // states, points and outlier flags of all windows always; the per-edge chi2 (2/3 of the bytes, nothing the reference's
    // function hands back) only if some result asks for it
    bool want_chi2 = false;
    for (int i = 0; i < ctx->n_win; ++i) want_chi2 = want_chi2 || out[i].obs_chi2 != nullptr;
    CK(cudaMemcpyAsync(hp, ctx->arena.base + ctx->out_region, want_chi2 ? ctx->out_total : ctx->out_small, cudaMemcpyDeviceToHost, s),
       "D2H results");
    CK(cudaStreamSynchronize(s), "sync");
    parallel_for(ctx->n_win, host_thread_budget(ctx->batch_total_hint > ctx->n_win ? ctx->n_lanes : 1), [&](int i) {
        const WinMeta& m = ctx->meta[i];
        const char* src = hp + (m.out_base - ctx->out_region);
        vilba_result& o = out[i];
        if (o.kf_state) std::memcpy(o.kf_state, src + m.L.o_kf, sizeof(double) * 22 * (size_t)m.K);
        if (o.pt_xyz && m.P) std::memcpy(o.pt_xyz, src + m.L.o_pts, sizeof(double) * 3 * (size_t)m.P);
        if (o.obs_chi2 && m.E) std::memcpy(o.obs_chi2, hp + (m.chi_base - ctx->out_region), sizeof(double) * (size_t)m.E);
        if (o.obs_outlier && m.E) std::memcpy(o.obs_outlier, src + m.L.o_outlier, (size_t)m.E);
    });
    return VILBA_OK;
}

void add_stats(vilba_stats& a, const vilba_stats& b) {
    a.kernel_launches += b.kernel_launches;
    a.lm_iterations += b.lm_iterations;
    a.lm_trials += b.lm_trials;
    a.edges_linearized += b.edges_linearized;
    a.linearize_ms += b.linearize_ms, a.linearize_launches += b.linearize_launches;
    a.schur_ms += b.schur_ms, a.schur_launches += b.schur_launches;
    a.solve_ms += b.solve_ms, a.solve_launches += b.solve_launches;
    a.update_ms += b.update_ms, a.update_launches += b.update_launches;
    a.preint_ms += b.preint_ms, a.preint_launches += b.preint_launches;
}

// ---- a batch split over lanes ---------------------------------------------------------------------
int ensure_lanes(vilba_ctx* ctx, int lanes) {
    while ((int)ctx->lanes.size() < lanes) {
        vilba_ctx* sub = vilba_create(ctx->device, &ctx->prm);
        if (!sub) {
            ctx->err = "could not create a batch lane";
            return VILBA_ERR_CUDA;
        }
        sub->max_batch = ctx->max_batch;
        if (cudaEventCreateWithFlags(&sub->ev_done, cudaEventDisableTiming) != cudaSuccess) return VILBA_ERR_CUDA;
        ctx->lanes.push_back(sub);
    }
    return VILBA_OK;
}

// how many lanes a batch of n windows is split over, and where the cuts are
int plan_split(const vilba_ctx* ctx, int n, std::vector<int>& first) {
    int lanes = 1;
    if (!ctx->comm && ctx->n_lanes > 1 && n >= 16) lanes = std::min(ctx->n_lanes, n / 8);
    lanes = std::max(lanes, (n + ctx->max_batch - 1) / ctx->max_batch);
    first.assign(lanes + 1, 0);
    for (int l = 0; l <= lanes; ++l) first[l] = (int)((long long)n * l / lanes);
    return lanes;
}

template <class F>
int for_each_lane(vilba_ctx* ctx, int lanes, bool concurrent, F&& f) {
    std::vector<int> status(lanes, VILBA_OK);
    if (concurrent && lanes > 1) {
        std::vector<std::thread> th;
        for (int l = 0; l < lanes; ++l) th.emplace_back([&, l]() { status[l] = f(l); });
        for (auto& t : th) t.join();
    } else {
        for (int l = 0; l < lanes; ++l) status[l] = f(l);
    }
    int worst = VILBA_OK;
    for (int l = 0; l < lanes; ++l) {
        add_stats(ctx->stats, ctx->lanes[l]->stats);
        std::memset(&ctx->lanes[l]->stats, 0, sizeof(vilba_stats));
        if (status[l] != VILBA_OK && worst >= 0) {
            worst = status[l];
            ctx->err = ctx->lanes[l]->err;
        }
    }
    return worst;
}

int upload_split(vilba_ctx* ctx, int n, const vilba_window* wins) {
    ctx->split = 0;
    ctx->n_win = 0;
    if (n <= 0 || !wins) return VILBA_ERR_ARG;
    const int lanes = plan_split(ctx, n, ctx->split_first);
    if (lanes == 1) return upload_batch(ctx, n, wins);
    int r = ensure_lanes(ctx, lanes);
    if (r != VILBA_OK) return r;
    r = for_each_lane(ctx, lanes, true, [&](int l) {
        ctx->lanes[l]->batch_total_hint = n;
        return upload_batch(ctx->lanes[l], ctx->split_first[l + 1] - ctx->split_first[l], wins + ctx->split_first[l]);
    });
    if (r == VILBA_OK) ctx->split = lanes;
    return r;
}

int solve_split(vilba_ctx* ctx, vilba_result* out) {
    if (!ctx->split) return solve_batch(ctx, out, nullptr);
    const int lanes = ctx->split;
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    CK(cudaEventRecord(ctx->ev_a, ctx->stream), "event");
    for (int l = 0; l < lanes; ++l) {
        ctx->lanes[l]->start_after = ctx->ev_a;
        ctx->lanes[l]->profiling = ctx->profiling;
    }
    // (a profiling pass keeps the lanes concurrent: the per-kernel-group times are those of the timed configuration,
    // contention between the lanes included)
    int r = for_each_lane(ctx, lanes, true, [&](int l) {
        return solve_batch(ctx->lanes[l], out + ctx->split_first[l], nullptr);
    });
    for (int l = 0; l < lanes; ++l) {
        ctx->lanes[l]->start_after = nullptr;
        if (r == VILBA_OK) CK(cudaStreamWaitEvent(ctx->stream, ctx->lanes[l]->ev_done, 0), "wait lane");
    }
    if (r != VILBA_OK) return r;
    CK(cudaEventRecord(ctx->ev_b, ctx->stream), "event");
    CK(cudaStreamSynchronize(ctx->stream), "sync");
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b), "elapsed");
    const int n = ctx->split_first[lanes];
    for (int i = 0; i < n; ++i) out[i].solve_ms = ms;  // device time from the first kernel of any lane to the last
    return VILBA_OK;
}

int download_split(vilba_ctx* ctx, vilba_result* out) {
    if (!ctx->split) return download_batch(ctx, out);
    return for_each_lane(ctx, ctx->split, true, [&](int l) { return download_batch(ctx->lanes[l], out + ctx->split_first[l]); });
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

const char* vilba_version(void) { return "vilba 0.2 (sm_100a, FP64)"; }

int vilba_max_batch(void) { return kMaxBatch; }

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
    std::memset(&ctx->dims, 0, sizeof(ctx->dims));
    if (const char* e = std::getenv("VILBA_CHOL_CLUSTER")) ctx->chol_cluster = std::max(1, std::atoi(e));
    if (const char* e = std::getenv("VILBA_GRAPH")) ctx->use_graph = std::atoi(e) != 0;
    if (const char* e = std::getenv("VILBA_CHOL_BIG_ABOVE")) ctx->chol_big_above = std::atoi(e);
    if (const char* e = std::getenv("VILBA_CHOL_LA")) ctx->chol_la_mode = std::atoi(e);
    if (const char* e = std::getenv("VILBA_PREINT_GROUP")) ctx->preint_group = std::atoi(e);
    if (const char* e = std::getenv("VILBA_SCHUR")) ctx->schur_gather_only = std::strcmp(e, "gather") == 0;
    if (const char* e = std::getenv("VILBA_SP_GRID")) ctx->sp_grid_cap = std::max(1, std::atoi(e));
    if (const char* e = std::getenv("VILBA_SP_SETS")) ctx->sp_sets = std::max(1, std::atoi(e));
    if (const char* e = std::getenv("VILBA_SP_MMA")) ctx->sp_mma = std::atoi(e) != 0;
    if (const char* e = std::getenv("VILBA_SP_PAIR")) ctx->sp_pair_lanes = std::atoi(e);
    if (const char* e = std::getenv("VILBA_COMM_GRAPH")) ctx->comm_graph = std::atoi(e) != 0;
    if (const char* e = std::getenv("VILBA_BATCH_LANES")) ctx->n_lanes = std::max(1, std::atoi(e));
    if (const char* e = std::getenv("VILBA_MAX_BATCH")) ctx->max_batch = std::max(1, std::min(kMaxBatch, std::atoi(e)));
    ctx->dims = choose_dims(ctx, 1, 0, 0, 0);
    if (cudaMalloc(&ctx->dwp, sizeof(DevWindow) * kMaxBatch) != cudaSuccess) {
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
    drop_graphs(ctx);
    if (ctx->comm && ctx->p_ncclCommDestroy) ctx->p_ncclCommDestroy(ctx->comm);
    if (ctx->dwp) cudaFree(ctx->dwp);
    ctx->arena.release();
    ctx->preint_arena.release();
    ctx->pinned.release();
    ctx->pinned_out.release();
    ctx->pinned_small.release();
    ctx->pinned_preint.release();
    if (ctx->ev_done) cudaEventDestroy(ctx->ev_done);
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
    ctx->split = 0;
    return upload_batch(ctx, 1, win);
}

int vilba_window_solve_resident(vilba_ctx* ctx, vilba_result* out) {
    if (!ctx || !out) return VILBA_ERR_ARG;
    if (ctx->n_win != 1) {
        ctx->err = "no single window uploaded";
        return VILBA_ERR_ARG;
    }
    int r = solve_batch(ctx, out, nullptr);
    out->status = r;
    return r;
}

int vilba_window_download(vilba_ctx* ctx, vilba_result* out) {
    if (!ctx || !out || ctx->n_win != 1) return VILBA_ERR_ARG;
    return download_batch(ctx, out);
}

int vilba_batch_upload(vilba_ctx* ctx, int32_t n_windows, const vilba_window* win) {
    if (!ctx) return VILBA_ERR_ARG;
    return upload_split(ctx, n_windows, win);
}

int vilba_batch_groups(const vilba_ctx* ctx) { return ctx ? std::max(1, ctx->split) : 0; }

static int resident_count(const vilba_ctx* ctx) { return ctx->split ? ctx->split_first[ctx->split] : ctx->n_win; }

int vilba_batch_solve_resident(vilba_ctx* ctx, int32_t n_windows, vilba_result* out) {
    if (!ctx || !out || n_windows != resident_count(ctx)) return VILBA_ERR_ARG;
    int r = solve_split(ctx, out);
    for (int i = 0; i < n_windows; ++i) out[i].status = r;
    return r;
}

int vilba_batch_download(vilba_ctx* ctx, int32_t n_windows, vilba_result* out) {
    if (!ctx || !out || n_windows != resident_count(ctx)) return VILBA_ERR_ARG;
    return download_split(ctx, out);
}

// ---- one large window sharded by map point over the GPUs of a node (BASELINE config 4) ----
int vilba_comm_unique_id(void* out128) {
    if (!out128 || sizeof(ncclUniqueId) != 128) return VILBA_ERR_ARG;
    void* lib = open_nccl();
    if (!lib) return VILBA_ERR_COMM;
    auto get = reinterpret_cast<ncclResult_t (*)(ncclUniqueId*)>(dlsym(lib, "ncclGetUniqueId"));
    if (!get || get(static_cast<ncclUniqueId*>(out128)) != ncclSuccess) return VILBA_ERR_COMM;
    return VILBA_OK;
}

int vilba_comm_init(vilba_ctx* ctx, const void* unique_id128, int32_t rank, int32_t world) {
    if (!ctx || !unique_id128 || world < 1 || world > 64 || rank < 0 || rank >= world || ctx->comm) return VILBA_ERR_ARG;
    void* lib = open_nccl();
    if (!lib) {
        ctx->err = "libnccl.so.2 not found";
        return VILBA_ERR_COMM;
    }
    ctx->nccl_lib = lib;
    ctx->p_ncclCommInitRank = reinterpret_cast<decltype(ctx->p_ncclCommInitRank)>(dlsym(lib, "ncclCommInitRank"));
    ctx->p_ncclCommDestroy = reinterpret_cast<decltype(ctx->p_ncclCommDestroy)>(dlsym(lib, "ncclCommDestroy"));
    ctx->p_ncclAllReduce = reinterpret_cast<decltype(ctx->p_ncclAllReduce)>(dlsym(lib, "ncclAllReduce"));
    ctx->p_ncclGetErrorString = reinterpret_cast<decltype(ctx->p_ncclGetErrorString)>(dlsym(lib, "ncclGetErrorString"));
    if (!ctx->p_ncclCommInitRank || !ctx->p_ncclAllReduce || !ctx->p_ncclCommDestroy) {
        ctx->err = "NCCL symbols missing";
        return VILBA_ERR_COMM;
    }
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    ncclUniqueId id;
    std::memcpy(&id, unique_id128, sizeof(id));
    const ncclResult_t r = ctx->p_ncclCommInitRank(&ctx->comm, world, id, rank);
    if (r != ncclSuccess) {
        ctx->comm = nullptr;
        ctx->err = std::string("ncclCommInitRank: ") + (ctx->p_ncclGetErrorString ? ctx->p_ncclGetErrorString(r) : "?");
        return VILBA_ERR_COMM;
    }
    drop_graphs(ctx);  // slots captured without the collectives
    ctx->comm_rank = rank, ctx->comm_world = world;
    ctx->slot_comm.self = ctx;
    ctx->slot_comm.reduce = slot_reduce;
    ctx->n_win = 0;  // layouts depend on the mode
    return VILBA_OK;
}

int vilba_shard_points(const vilba_window* win, int32_t rank, int32_t world, int32_t* p_begin, int32_t* p_end) {
    if (!win || world < 1 || rank < 0 || rank >= world || !p_begin || !p_end) return VILBA_ERR_ARG;
    // contiguous point ranges balanced by edge count: rank r owns the points whose first edge lies in
    // [E r / world, E (r + 1) / world)
    auto cut = [&](int r) {
        if (r <= 0) return 0;
        if (r >= world) return (int)win->n_pts;
        const long long target = (long long)win->n_obs * r / world;
        int lo = 0, hi = win->n_pts;
        while (lo < hi) {
            const int mid = (lo + hi) / 2;
            if (win->pt_obs_begin[mid] < target) lo = mid + 1; else hi = mid;
        }
        return lo;
    };
    *p_begin = win->n_pts ? cut(rank) : 0;
    *p_end = win->n_pts ? cut(rank + 1) : 0;
    return VILBA_OK;
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
    ctx->split = 0;
    int r = upload_batch(ctx, 1, win);
    if (r == VILBA_OK) r = solve_batch(ctx, out, stop_flag);
    if (r == VILBA_OK) r = download_batch(ctx, out);
    out->status = r;
    return r;
}

void vilba_global_ba_params(const vilba_params* base, int32_t n_iterations, int32_t robust, vilba_params* p) {
    if (base) *p = *base;
    else vilba_default_params(p);
    p->mode = VILBA_MODE_SINGLE_STAGE | (robust ? 0 : VILBA_MODE_MONO_NOT_ROBUST);
    p->iters_stage1 = n_iterations;  // optimizer.optimize(nIterations)  Optimizer.cpp:1624
    p->iters_stage2 = 0;
    if (robust) {
        p->huber_pvr = (double)(float)std::sqrt(21.666);   // const float thHuberNavStatePVR   Optimizer.cpp:1438
        p->huber_bias = (double)(float)std::sqrt(16.812);  // const float thHuberNavStateBias  Optimizer.cpp:1439
        p->huber_mono = (double)(float)std::sqrt(5.99);    // const float thHuber2D            Optimizer.cpp:1541
    } else {
        // no kernel on the IMU edges: with an infinite threshold rho' == 1 exactly, and g2o's robust branch
        // (base_multi_edge.hpp:36-48, base_binary_edge.hpp:61-76) then multiplies by 1.0 -- bit-identical
        p->huber_pvr = p->huber_bias = INFINITY;
    }
}

// Entry 1b: Optimizer::GlobalBundleAdjustmentNavState (Optimizer.cpp:1392-1668), phases "optimize" only
int vilba_global_ba(vilba_ctx* ctx, const vilba_window* win, int32_t n_iterations, int32_t robust, vilba_result* out,
                    const volatile uint8_t* stop_flag) {
    if (!ctx || !out || n_iterations < 0) return VILBA_ERR_ARG;
    if (ctx->comm) {
        ctx->err = "vilba_global_ba on a sharded context is not supported";
        return VILBA_ERR_ARG;
    }
    out->n_trace = 0;
    out->stage2_ran = 0;
    out->n_outliers_stage1 = 0;
    out->solve_ms = 0.0;
    const vilba_params saved = ctx->prm;
    vilba_global_ba_params(&saved, n_iterations, robust, &ctx->prm);
    ctx->split = 0;
    int r = upload_batch(ctx, 1, win);
    if (r == VILBA_OK) r = solve_batch(ctx, out, stop_flag);
    if (r == VILBA_OK) r = download_batch(ctx, out);
    ctx->prm = saved;
    out->status = r;
    return r;
}

// Independent windows (BASELINE config 5).  16 or more windows are split over up to 4 lanes (sub-contexts with their
// own streams, arena and host thread) that upload, solve and download concurrently; every lane is ONE batched solve
// (each kernel launched once for all its windows).  More than max_batch * lanes windows take several rounds.
int vilba_local_ba_batch(vilba_ctx* ctx, int32_t n_windows, const vilba_window* win, vilba_result* out) {
    if (!ctx || n_windows < 0 || (n_windows && (!win || !out))) return VILBA_ERR_ARG;
    if (n_windows == 0) return VILBA_OK;
    // at most max_batch * lanes windows are resident at a time; within such a round the lanes run concurrently
    // (upload, solve and download of one lane overlap the others')
    const int round = ctx->max_batch * std::max(1, ctx->n_lanes);
    int worst = VILBA_OK;
    for (int b = 0; b < n_windows; b += round) {
        const int nb = std::min(round, (int)n_windows - b);
        std::vector<int> first;
        const int lanes = plan_split(ctx, nb, first);
        int r;
        if (lanes == 1) {
            r = upload_batch(ctx, nb, win + b);
            if (r == VILBA_OK) r = solve_batch(ctx, out + b, nullptr);
            if (r == VILBA_OK) r = download_batch(ctx, out + b);
        } else {
            r = ensure_lanes(ctx, lanes);
            ctx->split = 0;
            if (r == VILBA_OK)
                r = for_each_lane(ctx, lanes, true, [&](int l) {
                    vilba_ctx* c = ctx->lanes[l];
                    c->batch_total_hint = nb;
                    const int f = b + first[l], n = first[l + 1] - first[l];
                    // Staggered start: the lanes flatten their windows one after the other, each with the whole host thread
                    // budget, so the first lane's kernels run while the others still pack (all lanes packing side by side
                    // with a quarter of the threads each kept the GPU idle until the last one was done)
                    int q;
                    {
                        std::lock_guard<std::mutex> lock(ctx->upload_mx);
                        c->upload_exclusive = true;
                        q = upload_batch(c, n, win + f);
                        c->upload_exclusive = false;
                    }
                    if (q == VILBA_OK) q = solve_batch(c, out + f, nullptr);
                    if (q == VILBA_OK) q = download_batch(c, out + f);
                    return q;
                });
        }
        for (int i = 0; i < nb; ++i) out[b + i].status = r;
        if (r != VILBA_OK && worst >= 0) worst = r;
    }
    return worst;
}

int vilba_preintegrate_batch_dev(vilba_ctx* ctx, int32_t n_pairs, int32_t n_samples, const int32_t* sample_begin_dev,
                                 const double* gyro_dev, const double* acc_dev, const double* dt_dev,
                                 const double* bg_dev, const double* ba_dev, double* out_dev) {
    if (!ctx || n_pairs < 0) return VILBA_ERR_ARG;
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    if (ctx->profiling) CK(cudaEventRecord(ctx->ev_a, ctx->stream), "event");
    int group = ctx->preint_group;
    if (group == 0 && n_pairs > 0) {  // tiles of G samples: short intervals waste fewer lanes with small groups
        const int avg = n_samples / n_pairs;
        group = avg <= 64 ? -8 : (avg <= 160 ? -16 : -32);
    }
    CK(launch_preint_batch(ctx->stream, n_pairs, sample_begin_dev, gyro_dev, acc_dev, dt_dev, bg_dev, ba_dev, out_dev,
                           ctx->prm.gyr_meas_cov, ctx->prm.acc_meas_cov, group), "preint_batch");
    ctx->stats.kernel_launches += n_pairs > 0 ? 1 : 0;
    if (ctx->profiling) {  // device time of the kernel alone, CUDA events on the launching stream
        CK(cudaEventRecord(ctx->ev_b, ctx->stream), "event");
        CK(cudaEventSynchronize(ctx->ev_b), "sync");
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b), "elapsed");
        ctx->stats.preint_ms += ms, ctx->stats.preint_launches++;
    }
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
    CK(ctx->pinned_preint.reserve(in_end > sizeof(double) * 142 * (size_t)n_pairs ? in_end : sizeof(double) * 142 * (size_t)n_pairs),
       "cudaMallocHost(preint)");
    char* h = ctx->pinned_preint.base;
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

// include/vilba_diag.h: the normal equations of the first LM trial, straight from the device buffers
int vilba_diag_first_trial(vilba_ctx* ctx, const vilba_window* win, double* lambda_out, double* Hpp, double* bp, double* Hll,
                           double* bl, double* W, double* S, double* bs) {
    if (!ctx || !win || ctx->comm) return VILBA_ERR_ARG;
    ctx->split = 0;
    int r = upload_batch(ctx, 1, win);
    if (r != VILBA_OK) return r;
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    cudaStream_t s = ctx->stream;
    CK(launch_reset(s, ctx->dwp, ctx->dims), "reset");
    CK(launch_imu_prepare(s, ctx->dwp, ctx->dims), "imu_prepare");
    CK(launch_eval_initial(s, ctx->dwp, ctx->dims), "eval");
    CK(launch_stage_begin(s, ctx->dwp, ctx->dims, 1, 1), "stage_begin");
    LaunchDims d = ctx->dims;
    d.dbg_stop_after_schur = 1;
    CK(launch_slot(s, ctx->stream2, ctx->ev_fork, ctx->ev_join, ctx->dwp, d, nullptr, nullptr), "slot");
    CK(cudaStreamSynchronize(s), "sync");
    const DevWindow& dw = ctx->dw[0];
    const WinMeta& m = ctx->meta[0];
    const size_t n = (size_t)m.n;
    LmState lm;
    CK(cudaMemcpy(&lm, dw.lm, sizeof(lm), cudaMemcpyDeviceToHost), "D2H lm");
    if (lambda_out) *lambda_out = lm.lambda_first;
    auto get = [&](double* dst, const double* src, size_t count) {
        return !dst || !count || cudaMemcpy(dst, src, sizeof(double) * count, cudaMemcpyDeviceToHost) == cudaSuccess;
    };
    bool ok = get(Hpp, dw.Hpp, n * n) && get(bp, dw.bp, n) && get(Hll, dw.Hll, 6 * (size_t)m.P) && get(bl, dw.bl, 3 * (size_t)m.P) &&
              get(W, dw.W, 18 * (size_t)m.E) && get(bs, dw.bs, n);
    if (ok && S && n)
        ok = cudaMemcpy2D(S, sizeof(double) * n, dw.S, sizeof(double) * (size_t)dw.lds, sizeof(double) * n, n,
                          cudaMemcpyDeviceToHost) == cudaSuccess;
    if (!ok) {
        ctx->err = "vilba_diag_first_trial: copy failed";
        return VILBA_ERR_CUDA;
    }
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
