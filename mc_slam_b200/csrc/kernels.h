// kernels.h -- device-side data layout and kernel launchers of the VI local-BA path (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vilba.h"

namespace vilba {

constexpr int kPreintThreads = 256;
constexpr int kPointThreads = 256;  // 8 warps per CTA, one warp per map point
constexpr int kCholThreads = 512;
constexpr int kMaxKF = 256;          // key-frames per window supported by the shared-memory stage
constexpr int kPointGridPerSM = 2;   // CTAs per SM of the grid-stride per-point kernels (fixed grid => graph-capturable)
constexpr int kImuGrid = 64;         // CTAs (one warp each) of the per-IMU-edge kernels, grid-stride
constexpr int kMaxBatch = 64;        // windows per batched launch (grid.y)

// obs record: 16 bytes, one vector load per mono edge.
//   x = bits of float u, y = bits of float v, z = bits of float invSigma2,
//   w = key-frame index (bits 0..23) | OBS_CULLED | OBS_ROBUST
constexpr int OBS_KF_MASK = 0x00ffffff;
constexpr int OBS_CULLED = 1 << 24;  // e->setLevel(1)            (Optimizer.cpp:2667-2670)
constexpr int OBS_ROBUST = 1 << 25;  // edge still has its Huber  (Optimizer.cpp:2672)

// Device-resident Levenberg-Marquardt controller state
// (members of OptimizationAlgorithmLevenberg, optimization_algorithm_levenberg.cpp:41-60)
enum LmPhase { PH_LINEARIZE = 0, PH_TRIAL = 1, PH_DONE = 2 };

struct IterRec {  // one outer LM iteration, written by the device-side controller
    int stage, iteration, trials, result, n_active, accepted;
    double chi0, chi1, lambda, lambda_first;
};

struct LmState {
    double lambda;        // _currentLambda
    double ni;            // _ni
    double current_chi;   // currentChi
    double ini_chi;       // iniChi
    double temp_chi;      // tempChi
    double rho;
    double lambda_first;  // lambda used by the first trial of this iteration (trace)
    double chi_acc;       // accumulator: robust chi2 of the state being evaluated
    double scale_acc;     // accumulator: sum_j x_j (lambda x_j + b_j) over landmarks
    unsigned long long maxdiag_bits;  // max |diag(H_ll)| as ordered bits (non-negative doubles)
    int cur;              // which of the two estimate buffers holds the current state
    int chol_fail;        // non-positive pivot in the reduced system
    int qmax;             // _levenbergIterations
    int n_bad;            // _nBad
    int accepted;         // last trial accepted
    int iter_result;      // -1 running, 0 OK, 1 Terminate
    int stop;             // force-stop flag mirrored from the host
    // device-side control flow: the host enqueues "slots" (linearise-if-needed + one LM trial + decide)
    // without synchronising; every kernel looks at `phase` and returns early when it has nothing to do
    int phase;            // LmPhase
    int iter;             // outer iteration inside the current optimize() call
    int max_iters;        // iterations of the current optimize() call
    int stage;            // 1 or 2
    int n_active;         // active edges of the stage (mono + 2 per IMU pair)
    int n_culled;         // edges moved to level 1 by the cull
    int n_trace;
    IterRec trace[VILBA_MAX_TRACE];
};

struct DevWindow {
    int K, NI, P, E, n_free, n;  // n = 15 * n_free
    int lds;                     // leading dimension of S and Lfac: n rounded up to a multiple of 4
    // estimates, double buffered (index LmState::cur = accepted state, 1-cur = trial state)
    double* kf_state[2];  // K * 22
    double* pts[2];       // P * 3
    const double* kf_state0;  // uploaded initial estimates (every solve restarts from them)
    const double* pts0;
    const char* obs0;        // E x 16 B as uploaded: uv (2 f32 per edge) | inv sigma^2 (f32) | key-frame index (i32)
    int obs_flags0;          // flag bits every edge record starts a solve with (OBS_ROBUST or 0)
    uint8_t* outlier;     // E: final outlier flags
    // packed results of this window inside the batch's output region (one D2H for the whole batch)
    double* out_kf_state;
    double* out_pts;
    double* out_chi2;
    uint8_t* out_outlier;
    const int* kf_block;  // K: block index among free key-frames, -1 if fixed
    // imu edges
    const int* imu_i;
    const int* imu_j;
    const double* imu_preint;  // NI * 142
    double* imu_info;          // NI * 81   inverse of cov_P_V_Phi
    double* imu_err;           // NI * 15   cached _error of (PVR 9, Bias 6)
    // mono edges, CSR by point
    const int* pt_obs_begin;  // P + 1
    int4* obs;                // E records
    double* obs_chi2;         // E: e->chi2() of the last evaluation that had the edge active
    // normal equations
    double* Hpp;  // n * n   upper block triangle + full diagonal blocks
    double* bp;   // n
    // where the assemble / Schur kernels WRITE: the same buffers for a whole window; this rank's partial sums
    // (send buffers of the allreduce into Hpp|bp and S|bs) when the window is point-sharded over several GPUs
    double* Hpp_w;
    double* bp_w;
    double* S_w;
    double* bs_w;
    int shard_owner;  // 1: this rank adds the terms that exist once per window (IMU edges, lambda I in S); whole window: 1
    int sharded;      // 1: point-sharded over several GPUs: Hpp / bp hold THIS RANK's partial sums (they never cross the wire:
                      // every rank folds its own partial H_pp into its partial S), diag_red carries what lambda's start needs
    int shard_rank, shard_world;
    double* diag_red; // n + shard_world: diag(H_pp) (summed over the ranks) | max |diag H_ll| of every rank
    double* Hll;  // P * 6   (xx,xy,xz,yy,yz,zz)
    double* bl;   // P * 3
    double* W;    // E * 18  H_pl block of the edge, 6x3 rows [P,Phi]
    double* S;    // n * n   reduced camera system (upper triangle used)
    double* Lfac; // n * n   Cholesky factor of S (same addressing as S)
    double* cminv; // (n/16 + 2) * 256   inverse of every 16x16 diagonal factor block (look-ahead variant)
    double* cdinv; // n      1 / L(j,j)
    double* bs;   // n
    double* x;    // n       pose increment
    // v2 (atomic-free) accumulation
    double* lin_partial;     // lin_ctas * n_free * 27   per-CTA partial pose blocks of the mono edges
    double* imu_slot;        // NI * 930                 30x30 block + 30 rhs of every IMU edge pair
    double* mono_sum;        // n_free * 27              fixed-order sum of the CTA partials
    double* Y;               // E * 24                   per trial: W D^-1 (18) and W (D^-1 b_l) (6)   (gather mode)
    double* ts_rec;          // P * 12                   per trial, per point: D^-1, D^-1 b_l, observer masks (tile-scan mode)
    unsigned* ts_hdr;        // tiles * 32               per trial, per tile: points observed by every key-frame
    double* schur_partial;   // sp_grid * (n_pairs * 36 + n_free * 6)   partial sums of sum_l W D^-1 W^T (tile-scan mode)
    const int* blk_edge_i;   // n_free: IMU edge in which the block is key-frame i, or -1
    const int* blk_edge_j;   // n_free: IMU edge in which the block is key-frame j, or -1
    const int* edge_pt;      // E: map point of every mono edge           (built on the device, pairs.cu)
    int n_pairs;             // n_free (n_free + 1) / 2 key-frame block pairs (a <= b)
    const int* pair_a;       // n_pairs
    const int* pair_b;       // n_pairs
    const int* blk_kf;       // n_free: key-frame index of every free block
    const int* pair_begin;   // n_pairs + 1                                (built on the device)
    const int* pair_ea;      // (edge of block a, edge of block b) sharing a map point   (built on the device)
    const int* pair_eb;
    int* edge_pt_rw;         // writable aliases used by the list builder
    int* pair_begin_rw;
    int* pair_ea_rw;
    int* pair_eb_rw;
    unsigned long long* pt_mask;  // P * 8: 256-bit key-frame masks per map point (all observers | free observers)
    LmState* lm;
    double* chi_partial;     // 2 * point_grid: per-CTA partial sums of chi2 and the landmark part of the gain scale
    unsigned* chi_counter;   // CTAs of update_eval that have delivered their partial
    long long* dbg;  // optional debug counters (16 x int64), may be null
    int dbg_flags;   // timing-ablation switches of the Schur tile kernels (env VILBA_TS_ABLATE; only honoured by -DVILBA_TS_ABLATE builds)
    // calibration (g2otypes.h:686-705)
    double fx, fy, cx, cy;
    double Rcb[9];  // Rbc^T
    double tcb[3];  // -Rcb * Pbc
    double g[3];
    // parameters
    double huber_mono, huber_pvr, huber_bias, chi2_gate;
    double inv_gyr_rw2, inv_acc_rw2;
    double lm_tau, lm_good_lo, lm_good_hi;
    int max_trials;
};

// Opt a kernel in to the largest dynamic shared-memory size the device allows.  The attribute is per FUNCTION and
// process-wide, not per context: a context configured for small windows must not lower the limit under a context that
// handles large ones, so every kernel simply gets the maximum (the carve-out still follows the size of each launch).
template <class F>
inline cudaError_t opt_in_max_smem(F func) {
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, func);
    if (e != cudaSuccess) return e;
    int dev = 0, optin = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev)) != cudaSuccess) return e;
    return cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes);
}

// ---- K1 ------------------------------------------------------------------------------------------
size_t preint_smem_bytes(int group);
cudaError_t launch_preint_batch(cudaStream_t stream, int n_pairs, const int* sample_begin, const double* gyro,
                                const double* acc, const double* dt, const double* bg, const double* ba,
                                double* out, double gyr_cov, double acc_cov, int group);

// point-sharded window: collectives the host controller inserts between the kernels of a slot
enum SlotReduce { RED_DIAG = 0, RED_S = 2, RED_CHI = 3 };
struct SlotComm {
    void* self;
    cudaError_t (*reduce)(void* self, int which, cudaStream_t s);
};

// ---- local BA ------------------------------------------------------------------------------------
// All kernels take a pointer to the device-resident DevWindow, so that their launch parameters do not
// change between windows and one LM "slot" can be captured once in a CUDA graph.
struct LaunchDims {
    int sm_count;
    int n_windows;        // grid.y: windows solved by one launch
    int point_grid;       // grid.x of the per-point kernels (per window)
    int imu_grid;         // grid.x of the per-IMU-edge kernels
    int gather_grid;      // grid.x of the Schur gather (CTAs of 1024 threads looping over block pairs)
    int reduce_grid;      // grid.x of reduce_partials
    int assemble_grid;    // grid.x of assemble_hpp
    int sp_warps;         // warps per CTA of the tile-scan Schur kernel; 0 => gather over pair lists instead
    int sp_sets;          // the key-frame block pairs are cut into this many contiguous subsets (CTAs) ...
    int sp_grid;          // ... and the points into this many subsets (= partial sums per window)
    int sp_tile_pts;      // map points per shared-memory tile
    int sp_pair_lanes;    // 1: one lane per block pair (36 accumulators), one CTA covers all pairs of the window
    int sp_mma;           // 1: tile kernel with one warp per hit on the FP64 MMA (operands Z = W G, D^-1 = G G^T); an option, not the default
    int sp_tile_edges;    // ... its shared-memory tile holds this many edges (the largest tile of the batch)
    int chol_cluster;     // CTAs of the Cholesky cluster
    int chol_la;          // 1: look-ahead cluster kernel with the trailing matrix in shared memory (chol_la.cu)
    int chol_n;           // largest reduced system of the batch (sizes the shared memory of chol_la)
    int chol_big_tiles;   // > 0: multi-kernel blocked LDL^T (chol_big.cu) with this many 64-column steps (systems too large for chol_la)
    size_t smem_point;    // dynamic shared memory of update_eval / flags
    size_t smem_lin;      // ... of linearize_v2
    int lin_threads;      // threads per CTA of linearize_v2: fewer warps (= fewer private accumulators) for many free key-frames
    size_t smem_sp;       // ... of schur_tile
    int dbg_stop_after_schur;  // diagnostics only (vilba_diag_first_trial): the slot ends behind the Schur step, S | b_s intact
};
size_t point_smem_bytes(int K);
size_t linearize_smem_bytes(int K, int n_free, int threads);
bool schur_tile_fits(int K, int n_free);   // the tile-scan Schur kernel handles windows of <= 32 key-frames
size_t schur_tile_smem_bytes(int max_K, int tile_pts);
size_t schur_tile_rec_doubles(int P);
int schur_mma_units(int n_free);                  // warps of the tensor-pipe kernel that cover all block pairs
size_t schur_mma_smem_bytes(int tile_edges, int tile_pts);
int schur_pair_lanes(int n_free);                 // lanes (= threads) of the lane-per-pair tile kernel
size_t schur_pair_partial_doubles(int n_free);    // size of one of its partial sums
size_t schur_tile_hdr_words(int P, int tile_pts);
size_t schur_partial_doubles(int n_free);
cudaError_t configure_chol_big(int n_cap);
size_t chol_big_scratch_doubles(int n);  // DevWindow::cminv must hold this many doubles when chol_big.cu is used
// look-ahead cluster kernel with the trailing matrix in shared memory (chol_la.cu)
int chol_la_tiles_per_thread(int n, int cluster);
size_t chol_la_smem_bytes(int n, int cluster);
size_t chol_la_scratch_doubles(int n);   // DevWindow::cminv must hold this many doubles
bool chol_la_fits(int n, int cluster);
cudaError_t configure_chol_la();
cudaError_t launch_chol_la(cudaStream_t s, const DevWindow* wp, int n_windows, int cluster, int n_cap);
cudaError_t launch_chol_big(cudaStream_t s, cudaStream_t side, cudaEvent_t ev_trsm, cudaEvent_t ev_rest, const DevWindow* wp,
                            const LaunchDims& d);
cudaError_t configure_kernels(const LaunchDims& d);  // cudaFuncSetAttribute for the large-smem kernels

cudaError_t launch_imu_prepare(cudaStream_t s, const DevWindow* wp, const LaunchDims& d);
cudaError_t launch_reset(cudaStream_t s, const DevWindow* wp, const LaunchDims& d);
cudaError_t launch_export(cudaStream_t s, const DevWindow* wp, const LaunchDims& d);
cudaError_t launch_eval_initial(cudaStream_t s, const DevWindow* wp, const LaunchDims& d);  // eval at the current state
cudaError_t launch_stage_begin(cudaStream_t s, const DevWindow* wp, const LaunchDims& d, int stage, int max_iters);
// one slot = [linearize (mono | imu on `side`) -> reduce -> assemble -> iter_begin] if phase == LINEARIZE,
//            [schur prep -> gather -> cholesky -> update+eval -> decide] if phase == TRIAL
// `probe` (8 timing events, or NULL; [6] after schur_prep, [7] after linearize_v2): [0,1] linearize+reduce+assemble, [2,3] Schur prep+gather, [3,4] Cholesky,
// [4,5] update+eval
cudaError_t launch_slot(cudaStream_t s, cudaStream_t side, cudaEvent_t fork, cudaEvent_t join, const DevWindow* wp,
                        const LaunchDims& d, cudaEvent_t* probe, const SlotComm* comm = nullptr);
cudaError_t launch_build_pair_lists(cudaStream_t s, const DevWindow* wp, const LaunchDims& d);
cudaError_t launch_cull(cudaStream_t s, const DevWindow* wp, const LaunchDims& d);
cudaError_t launch_final_flags(cudaStream_t s, const DevWindow* wp, const LaunchDims& d);
// kernels one slot launches (for vilba_stats::kernel_launches): 5 common (+ 4 when sharded: iter_begin, decide, diag, scale) + Schur
// (2 gather / 3 tile scan) + Cholesky
inline int kernels_per_slot(const LaunchDims& d, bool sharded = false) {
    const int schur = d.sp_warps > 0 ? 3 : 2;
    const int chol = d.chol_big_tiles > 0 ? 3 * d.chol_big_tiles : 1;  // diag + panel + update per step (last update replaced by the back substitution)
    return 5 + (sharded ? 4 : 0) + schur + chol;
}

}  // namespace vilba
