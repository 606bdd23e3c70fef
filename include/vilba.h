/*
 * vilba.h -- C ABI of the B200-native visual-inertial local bundle adjustment
 * (VI local BA) and batched IMU pre-integration path.
 *
 * This is the drop-in boundary for ONE hot path of mc275/MC_SLAM.  The
 * reference has no FFI layer: the path is entered through two C++ symbols,
 *
 *   Optimizer::LocalBundleAdjustmentNavState(KeyFrame*, const std::list<KeyFrame*>&,
 *                                            bool* pbStopFlag, Map*, cv::Mat& gw, LocalMapping*)
 *       -- include/Optimizer.h:44-46, src/Optimizer.cpp:2320-2771
 *   IMUPreintegrator::update(const Vector3d& omega, const Vector3d& acc, const double& dt)
 *       -- src/IMU/IMUPreintegrator.h:27-33, src/IMU/IMUPreintegrator.cpp:63-112
 *          (driven by KeyFrame::ComputePreInt, src/KeyFrame.cpp:195-252)
 *
 * so the entry points below are what a maintainer's C++ shim for those two
 * symbols binds (see INTEGRATION.md and shim/).  Everything is plain pointers
 * and sizes; no torch / CUDA types appear in a signature.  All pointers are
 * HOST pointers unless the name ends in `_dev`.
 *
 * The same structs are consumed by the CPU oracle (oracle/vilba_oracle.h) so
 * that parity tests feed both sides identical bytes.
 */
#ifndef VILBA_H
#define VILBA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------------------------------------
 * Sizes of the flat records
 * ------------------------------------------------------------------------------------------- */

/* One NavState (src/IMU/NavState.h:126-138) flattened to 22 doubles:
 *   [0:3) P   [3:6) V   [6:10) unit quaternion of R stored (w,x,y,z)
 *   [10:13) bias_gyr   [13:16) bias_acc   [16:19) delta_bias_gyr   [19:22) delta_bias_acc      */
#define VILBA_NS_DOUBLES 22

/* One IMUPreintegrator (src/IMU/IMUPreintegrator.h:177-196) flattened to 142 doubles, row-major:
 *   [0:3) delta_P  [3:6) delta_V  [6:15) delta_R (3x3)
 *   [15:24) J_P_Biasg  [24:33) J_P_Biasa  [33:42) J_V_Biasg  [42:51) J_V_Biasa  [51:60) J_R_Biasg
 *   [60:141) cov_P_V_Phi (9x9, order P,V,Phi)   [141] delta_time                                 */
#define VILBA_PREINT_DOUBLES 142
#define VILBA_PI_DP 0
#define VILBA_PI_DV 3
#define VILBA_PI_DR 6
#define VILBA_PI_JPG 15
#define VILBA_PI_JPA 24
#define VILBA_PI_JVG 33
#define VILBA_PI_JVA 42
#define VILBA_PI_JRG 51
#define VILBA_PI_COV 60
#define VILBA_PI_DT 141

/* Key-frames per window, free and fixed together: their states are staged in shared memory by every per-point
 * kernel (the reference has no limit; its local windows hold 10-20 local and a few dozen fixed key-frames, a global
 * BA over a longer map has to be cut or sharded).  Larger windows are rejected with VILBA_ERR_ARG. */
#define VILBA_MAX_KEYFRAMES 256

/* key-frame flags */
#define VILBA_KF_FIXED 1u      /* VertexNavStatePVR::setFixed(true)   (src/Optimizer.cpp:2454-2467)        */
#define VILBA_KF_HAS_BIAS 2u   /* a VertexNavStateBias exists for it  (src/Optimizer.cpp:2435-2443,2470-2478) */

/* vilba_params.mode: which caller's optimisation schedule the solve follows */
#define VILBA_MODE_SINGLE_STAGE 1u    /* one optimize(iters_stage1): no cull, no second stage, the stop flag does
                                         not abort before the solve (GlobalBundleAdjustmentNavState,
                                         src/Optimizer.cpp:1621-1624)                                            */
#define VILBA_MODE_MONO_NOT_ROBUST 2u /* mono edges are built without a Huber kernel (bRobust == false,
                                         src/Optimizer.cpp:1590-1595)                                            */

/* status codes */
#define VILBA_OK 0
#define VILBA_ABORTED 1          /* *stop_flag was set before optimisation: nothing written (Optimizer.cpp:2643-2645) */
#define VILBA_ERR_ARG (-1)
#define VILBA_ERR_CUDA (-2)
#define VILBA_ERR_NO_DEVICE (-3)
#define VILBA_ERR_COMM (-4)

/* ---------------------------------------------------------------------------------------------
 * Parameters.  Defaults (vilba_default_params) are the literals of the reference.
 * ------------------------------------------------------------------------------------------- */
typedef struct vilba_params {
    int32_t iters_stage1;      /* 5   optimizer.optimize(5)   src/Optimizer.cpp:2648                   */
    int32_t iters_stage2;      /* 10  optimizer.optimize(10)  src/Optimizer.cpp:2676                   */
    int32_t max_trials;        /* 10  _maxTrialsAfterFailure  optimization_algorithm_levenberg.cpp:51  */
    int32_t mode;              /* VILBA_MODE_* flags; 0 = LocalBundleAdjustmentNavState                   */
    double huber_mono;         /* (double)(float)sqrt(5.991)        Optimizer.cpp:2580,2624            */
    double huber_pvr;          /* (double)(float)sqrt(100*21.666)   Optimizer.cpp:2487                 */
    double huber_bias;         /* (double)(float)sqrt(100*16.812)   Optimizer.cpp:2488                 */
    double chi2_gate;          /* 5.991 (double literal)            Optimizer.cpp:2667,2694            */
    double lm_tau;             /* 1e-5  optimization_algorithm_levenberg.cpp:45                        */
    double lm_good_lo;         /* 1/3   :47                                                            */
    double lm_good_hi;         /* 2/3   :46                                                            */
    double gyr_bias_rw2;       /* (2e-5)^2   src/IMU/imudata.cpp:25                                    */
    double acc_bias_rw2;       /* (5e-3)^2   src/IMU/imudata.cpp:26                                    */
    double gyr_meas_cov;       /* 1.7e-4^2/0.005       (diagonal value) src/IMU/imudata.cpp:28-29      */
    double acc_meas_cov;       /* 2.0e-3^2/0.005*100   (diagonal value) src/IMU/imudata.cpp:30-31      */
} vilba_params;

void vilba_default_params(vilba_params* p);

/* ---------------------------------------------------------------------------------------------
 * One local-BA window, struct-of-arrays.  This is phase A+B of the reference function
 * (gather + graph build, src/Optimizer.cpp:2329-2639) flattened; the shim produces it from
 * KeyFrame/MapPoint objects, the synthetic generator produces it directly.
 *
 * Ordering contract (mirrors g2o's deterministic ordering, sparse_optimizer.cpp:166-190,482-487):
 *   - free key-frames appear in kf_* in increasing KeyFrame::mnId; their position among the free
 *     ones is their block index in the reduced camera system (15 scalars each: P,V,Phi,dbg,dba);
 *   - imu edges in the order of lLocalKeyFrames (Optimizer.cpp:2494-2541);
 *   - points in lLocalMapPoints order, and the mono observations of point p are
 *     obs[pt_obs_begin[p] .. pt_obs_begin[p+1]) ordered by KeyFrame id (MapPoint.h:28).
 * ------------------------------------------------------------------------------------------- */
typedef struct vilba_window {
    /* key-frames */
    int32_t n_kf;                  /* <= VILBA_MAX_KEYFRAMES (fixed ones included)                         */
    int32_t n_imu;                 /* number of (EdgeNavStatePVR, EdgeNavStateBias) pairs                  */
    int32_t n_pts;
    int32_t n_obs;                 /* number of EdgeNavStatePVRPointXYZ                                    */
    const double* kf_state;        /* n_kf * 22  (VILBA_NS_DOUBLES)                                        */
    const uint8_t* kf_flags;       /* n_kf       VILBA_KF_*                                                */
    const int64_t* kf_id;          /* n_kf       KeyFrame::mnId (informational; may be NULL)               */
    /* imu edges: EdgeNavStatePVR(PVR_i, PVR_j, Bias_i) + EdgeNavStateBias(Bias_i, Bias_j) */
    const int32_t* imu_kf_i;       /* n_imu      index into kf_* of pKF0 = pKF1->GetPrevKeyFrame()         */
    const int32_t* imu_kf_j;       /* n_imu      index into kf_* of pKF1                                   */
    const double* imu_preint;      /* n_imu * 142  pKF1->GetIMUPreInt()                                    */
    /* map points and mono observations (CSR by point) */
    const double* pt_xyz;          /* n_pts * 3  float-valued doubles (Converter::toVector3d)              */
    const int32_t* pt_obs_begin;   /* n_pts + 1                                                            */
    const int32_t* obs_kf;         /* n_obs      index into kf_*                                           */
    const float* obs_uv;           /* n_obs * 2  kpUn.pt.{x,y}                                             */
    const float* obs_inv_sigma2;   /* n_obs      pKFi->mvInvLevelSigma2[kpUn.octave]                       */
    /* calibration */
    double fx, fy, cx, cy;         /* float-valued (KeyFrame::fx.. are float)                              */
    double Rbc[9];                 /* row-major, ConfigParam::GetEigTbc().topLeftCorner(3,3)               */
    double Pbc[3];
    double gravity[3];             /* Converter::toVector3d(gw): float-valued                              */
} vilba_window;

/* Per outer-iteration record (what g2o's verbose line would print, sparse_optimizer.cpp:399-411,
 * plus what the parity tests compare). */
typedef struct vilba_iter_record {
    int32_t stage;                 /* 1 or 2                                                               */
    int32_t iteration;             /* index inside its optimize() call                                     */
    int32_t trials;                /* _levenbergIterations                                                 */
    int32_t result;                /* 0 OK, 1 Terminate                                                    */
    int32_t n_active_edges;        /* active mono + imu pvr + imu bias edges                               */
    int32_t accepted;              /* 1 if the last trial was accepted                                     */
    double chi2_initial;           /* iniChi: activeRobustChi2() at iteration start                        */
    double chi2_final;             /* currentChi at iteration end                                          */
    double lambda;                 /* _currentLambda at iteration end                                      */
    double lambda_first_trial;     /* _currentLambda used by the first trial                               */
} vilba_iter_record;

#define VILBA_MAX_TRACE 64

typedef struct vilba_result {
    /* caller-allocated output arrays (same shapes as the inputs) */
    double* kf_state;              /* n_kf * 22  : P,V,R,dbg,dba updated for free KFs, others copied        */
    double* pt_xyz;                /* n_pts * 3  : optimised positions (double; the shim rounds to float)   */
    uint8_t* obs_outlier;          /* n_obs : 1 if chi2 > gate or depth <= 0 at the end (Optimizer.cpp:2694) */
    double* obs_chi2;              /* n_obs : e->chi2() as the reference's final loop reads it (may be NULL) */
    /* filled by the library */
    int32_t status;                /* VILBA_OK / VILBA_ABORTED / <0                                         */
    int32_t stage2_ran;            /* 0 if the stop flag interrupted after stage 1 (Optimizer.cpp:2650-2656) */
    int32_t n_trace;
    int32_t n_outliers_stage1;     /* edges moved to level 1 by the cull (Optimizer.cpp:2667-2670)          */
    vilba_iter_record trace[VILBA_MAX_TRACE];
    double solve_ms;               /* device time of the solve phases C..E (CUDA events), or CPU time (oracle) */
} vilba_result;

/* ---------------------------------------------------------------------------------------------
 * Context
 * ------------------------------------------------------------------------------------------- */
typedef struct vilba_ctx vilba_ctx;

/* Creates a context on CUDA device `device` (owning stream, scratch and graphs).
 * Returns NULL if no CUDA device / the kernels cannot be loaded: there is NO CPU fallback. */
vilba_ctx* vilba_create(int device, const vilba_params* params /* NULL = defaults */);
void vilba_destroy(vilba_ctx* ctx);
const char* vilba_last_error(const vilba_ctx* ctx);
const char* vilba_version(void);

/* ---------------------------------------------------------------------------------------------
 * Entry 1: local BA.  Replaces phases C..E of Optimizer::LocalBundleAdjustmentNavState
 * (src/Optimizer.cpp:2643-2701): optimize(5) robust, cull, optimize(10), outlier flags.
 * `stop_flag` may be NULL; it is the reference's `bool* pbStopFlag` (polled between LM trials,
 * optimization_algorithm_levenberg.cpp:149, sparse_optimizer.cpp:376).
 * ------------------------------------------------------------------------------------------- */
int vilba_local_ba(vilba_ctx* ctx, const vilba_window* win, vilba_result* out,
                   const volatile uint8_t* stop_flag);

/* ---------------------------------------------------------------------------------------------
 * Entry 1b: global BA.  Replaces the optimisation of Optimizer::GlobalBundleAdjustmentNavState
 * (src/Optimizer.cpp:1392-1668): the same vertices and edges as the local BA, built over
 * the whole map (`win` = every good key-frame, the one with mnId 0 flagged VILBA_KF_FIXED | VILBA_KF_HAS_BIAS, and
 * every map point with at least one observation), then ONE optimize(n_iterations) -- no cull, no second stage.
 * `robust` is the reference's `bRobust`: when non-zero every edge carries a Huber kernel with the thresholds of
 * Optimizer.cpp:1438-1439,1541 (sqrt(21.666), sqrt(16.812), sqrt(5.99), each rounded to float); when zero no edge
 * has one.  The context's own parameters are used for everything else and are left unchanged.
 * A stop flag that is already set makes g2o run zero iterations (sparse_optimizer.cpp:376): the call returns
 * VILBA_OK with the input estimates, as the reference writes them back.  out->obs_outlier / obs_chi2 are filled
 * from the final errors with the context's chi2 gate; the reference does not read them on this path.
 * ------------------------------------------------------------------------------------------- */
int vilba_global_ba(vilba_ctx* ctx, const vilba_window* win, int32_t n_iterations, int32_t robust, vilba_result* out,
                    const volatile uint8_t* stop_flag);
/* the parameters vilba_global_ba solves with, derived from `base` (NULL = defaults) */
void vilba_global_ba_params(const vilba_params* base, int32_t n_iterations, int32_t robust, vilba_params* p);

/* Many independent windows in one call (BASELINE config 5).  The windows are solved in batched launches
 * (a few concurrent lanes of up to vilba_max_batch() windows each); out[i] corresponds to win[i]. */
int vilba_local_ba_batch(vilba_ctx* ctx, int32_t n_windows, const vilba_window* win, vilba_result* out);

/* Device-resident variant used to time the solve without host<->device copies:
 * upload once, solve many times (each solve restarts from the uploaded initial state), download. */
int vilba_window_upload(vilba_ctx* ctx, const vilba_window* win);
int vilba_window_solve_resident(vilba_ctx* ctx, vilba_result* out /* only status/trace/solve_ms filled */);
int vilba_window_download(vilba_ctx* ctx, vilba_result* out);

/* The same for a resident batch of independent windows: every kernel is launched once per lane for all the
 * lane's windows (one grid row per window) and each window runs its own device-side LM controller; batches of
 * 16 or more windows are split over up to 4 lanes (env VILBA_BATCH_LANES) that run concurrently.
 * out[i].solve_ms is the device time of the whole batch (first kernel of any lane to the last). */
int vilba_max_batch(void);
/* lanes (concurrent sub-batches, each with its own streams) the resident batch is split over: a launch of any
 * kernel covers n_windows / vilba_batch_groups() windows */
int vilba_batch_groups(const vilba_ctx* ctx);
int vilba_batch_upload(vilba_ctx* ctx, int32_t n_windows, const vilba_window* win);
int vilba_batch_solve_resident(vilba_ctx* ctx, int32_t n_windows, vilba_result* out);
int vilba_batch_download(vilba_ctx* ctx, int32_t n_windows, vilba_result* out);

/* ---------------------------------------------------------------------------------------------
 * One LARGE window sharded by map point over the GPUs of a node (BASELINE config 4; one process per GPU).
 * Every rank holds all key-frames and IMU edges and a contiguous range of the map points (with all their
 * observations).  Per LM trial the ranks exchange, with ncclAllReduce over NVLink: H_pp | b_p (sum) and
 * max |diag H_ll| (max) after the linearisation, the partial reduced camera systems S | b_s (sum) after the
 * Schur step, and chi2 | gain-scale (sum) after the update; the reduced system is solved redundantly and the
 * LM decisions are taken from the reduced (bit-identical) values, so the ranks never diverge.
 *
 *   rank 0:  vilba_comm_unique_id(id)  -> broadcast the 128 bytes to the other ranks (MPI, torch.distributed, ...)
 *   all:     vilba_comm_init(ctx, id, rank, world)            (collective, once per context)
 *   all:     vilba_shard_points(win, rank, world, &p0, &p1)   (which points this rank owns)
 *   all:     vilba_local_ba(ctx, sub_window, out, NULL)       (collective; sub_window = all key-frames / IMU edges
 *                                                              + the points [p0, p1) and their observations)
 * Every rank returns the same key-frame states and ITS points / outlier flags; trace[i].n_active_edges counts the
 * rank's own edges (the IMU edges on rank 0) and sums to the window's count over the ranks.  The stop flag is only
 * honoured before the solve starts.  NCCL (libnccl.so.2) is loaded at run time by vilba_comm_init.
 * ------------------------------------------------------------------------------------------- */
int vilba_comm_unique_id(void* out128);
int vilba_comm_init(vilba_ctx* ctx, const void* unique_id128, int32_t rank, int32_t world);
int vilba_shard_points(const vilba_window* win, int32_t rank, int32_t world, int32_t* p_begin, int32_t* p_end);

/* ---------------------------------------------------------------------------------------------
 * Entry 2: batched IMU pre-integration.  Replaces the IMUPreintegrator::reset()+update() loop of
 * KeyFrame::ComputePreInt (src/KeyFrame.cpp:210-249) for n_pairs key-frame pairs at once.
 * Pair p integrates samples [sample_begin[p], sample_begin[p+1]); each sample is one update() call
 * with omega = gyro - bg[p], acc = acc - ba[p] and its own dt (the caller lays out the leading
 * partial interval as an extra sample exactly as ComputePreInt does).
 * gyro/acc: 3 doubles per sample (x,y,z interleaved).  out: n_pairs * 142 doubles.
 * ------------------------------------------------------------------------------------------- */
int vilba_preintegrate_batch(vilba_ctx* ctx, int32_t n_pairs, const int32_t* sample_begin,
                             const double* gyro, const double* acc, const double* dt,
                             const double* bg, const double* ba, double* out);

/* Device-resident variant (pointers are device pointers on the ctx's device, stream-ordered on the
 * ctx stream; used by bench.py for the HBM-resident number). */
int vilba_preintegrate_batch_dev(vilba_ctx* ctx, int32_t n_pairs, int32_t n_samples,
                                 const int32_t* sample_begin_dev, const double* gyro_dev,
                                 const double* acc_dev, const double* dt_dev, const double* bg_dev,
                                 const double* ba_dev, double* out_dev);

/* ---------------------------------------------------------------------------------------------
 * Wire / on-disk format of a window (host code, no GPU needed).  The reference has no serialised form of a
 * window: it re-gathers it from KeyFrame / MapPoint objects on every call (src/Optimizer.cpp:2329-2402).  The
 * blob is the struct-of-arrays window above made self-describing (256-byte header: magic "VILBAWIN", version,
 * endianness tag, counts, calibration, payload length and FNV-1a 64 checksum; then the arrays in declaration
 * order, each padded to 8 bytes), so that windows captured from a patched reference build can be stored, sent to
 * another process / GPU and replayed through vilba_local_ba.  Layout: mc_slam_b200/csrc/window_blob.cu.
 *   vilba_window_blob_size   : bytes vilba_window_serialize will write (0 for an invalid window)
 *   vilba_window_serialize   : writes the blob into buf (capacity >= blob size); *written = its size
 *   vilba_window_deserialize : zero-copy view: *out points INTO buf (8-byte aligned, must outlive the view);
 *                              rejects wrong magic / version / endianness, truncation, checksum mismatch and
 *                              out-of-range indices with VILBA_ERR_ARG
 * ------------------------------------------------------------------------------------------- */
size_t vilba_window_blob_size(const vilba_window* win);
int vilba_window_serialize(const vilba_window* win, void* buf, size_t capacity, size_t* written);
int vilba_window_deserialize(const void* buf, size_t len, vilba_window* out);

/* ---------------------------------------------------------------------------------------------
 * Introspection for the bench harness
 * ------------------------------------------------------------------------------------------- */
typedef struct vilba_stats {
    int64_t kernel_launches;       /* kernels launched by this ctx since creation / last reset            */
    int64_t lm_iterations;         /* outer LM iterations executed                                         */
    int64_t lm_trials;             /* inner trials executed                                                */
    int64_t edges_linearized;      /* active edges summed over buildSystem() calls                         */
    double linearize_ms;           /* accumulated device time of the linearise+accumulate kernel           */
    int64_t linearize_launches;
    double schur_ms;
    int64_t schur_launches;
    double solve_ms;               /* reduced-system factor + solve                                        */
    int64_t solve_launches;
    double update_ms;              /* oplus + landmark back-substitution + residual evaluation of a trial  */
    int64_t update_launches;
    double preint_ms;              /* pre-integration kernel (vilba_preintegrate_batch[_dev]), profiling on */
    int64_t preint_launches;
} vilba_stats;

void vilba_get_stats(const vilba_ctx* ctx, vilba_stats* s);
void vilba_reset_stats(vilba_ctx* ctx);
/* when on, per-kernel-group CUDA-event timing is collected into vilba_stats: the slot kernels are launched one by one
 * instead of through the CUDA graph (events in between); the lanes of a split batch still run concurrently, i.e. in the
 * configuration that is timed without profiling */
void vilba_set_profiling(vilba_ctx* ctx, int on);

#ifdef __cplusplus
}
#endif
#endif /* VILBA_H */
