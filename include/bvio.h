/*
 * bvio.h -- C-ABI of the B200-native sliding-window VIO solver and anticipated
 * feature selector.  This is the drop-in boundary behind two call sites of
 * plusk01/Anticipated-VINS-Mono (all citations relative to the reference tree):
 *
 *   Estimator::optimization()   vins_estimator/src/estimator.cpp:661-994
 *   FeatureSelector::select()   vins_estimator/src/feature_selector.cpp:74-202
 *
 * Everything is plain C: POD structs, raw pointers and sizes.  All pointers are
 * HOST pointers owned by the caller; the library copies to/from HBM itself.
 * All floating point is IEEE double.  No function throws; every entry point
 * returns 0 (BVIO_OK) or a negative bvio_status.  A context is single-caller
 * (the reference runs this path under Estimator's m_estimator lock,
 * vins_estimator/src/estimator_node.cpp:222-377).
 *
 * There is NO CPU fallback behind this header: if no CUDA device is usable
 * bvio_create() fails with BVIO_ERR_CUDA.
 */
#ifndef BVIO_H_
#define BVIO_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BVIO_ABI_VERSION 2

typedef enum {
  BVIO_OK = 0,
  BVIO_ERR_INVALID = -1,    /* bad argument / inconsistent sizes            */
  BVIO_ERR_CUDA = -2,       /* CUDA runtime error (see bvio_last_error)     */
  BVIO_ERR_NUMERIC = -3,    /* non-finite cost or state                      */
  BVIO_ERR_UNSUPPORTED = -4,/* option not implemented on the device path     */
  BVIO_ERR_NCCL = -5
} bvio_status;

/* ------------------------------------------------------------------------ */
/*  Sliding-window bundle adjustment (Estimator::optimization)               */
/* ------------------------------------------------------------------------ */

/* One preintegrated IMU constraint between frames i-1 and i.
 * Replaces the members of IntegrationBase read by IMUFactor::Evaluate
 * (vins_estimator/src/factor/integration_base.h:195-203,
 *  vins_estimator/src/factor/imu_factor.h:19-179).
 * jacobian/covariance are 15x15 ROW-major in the reference's state order
 * O_P=0,O_R=3,O_V=6,O_BA=9,O_BG=12 (vins_estimator/src/parameters.h:58-65). */
typedef struct {
  double delta_p[3];
  double delta_q[4];      /* x y z w */
  double delta_v[3];
  double lin_ba[3];       /* linearized_ba */
  double lin_bg[3];       /* linearized_bg */
  double sum_dt;          /* factor skipped when > 10 (estimator.cpp:705)   */
  double jacobian[225];
  double covariance[225];
} bvio_preint;

/* Kinds of parameter block a marginalization prior can keep
 * (vins_estimator/src/estimator.cpp:823-921: Pose[i], SpeedBias[i],
 *  Ex_Pose[0], Td[0]). */
enum { BVIO_BLK_POSE = 0, BVIO_BLK_SPEEDBIAS = 1, BVIO_BLK_EXPOSE = 2, BVIO_BLK_TD = 3 };

/* Linearized prior  r = lin_res + lin_jac * dx   left by the previous
 * marginalization: MarginalizationInfo::{linearized_jacobians,
 * linearized_residuals, keep_block_size, keep_block_idx, keep_block_data}
 * (vins_estimator/src/factor/marginalization_factor.h:64-69) evaluated as in
 * MarginalizationFactor::Evaluate (marginalization_factor.cpp:333-381). */
typedef struct {
  int32_t n;                   /* residual count = sum of LOCAL block sizes  */
  int32_t nblocks;
  const int32_t* block_kind;   /* [nblocks] BVIO_BLK_*                        */
  const int32_t* block_frame;  /* [nblocks] frame index (POSE/SPEEDBIAS)      */
  const int32_t* block_idx;    /* [nblocks] first column in lin_jac (= keep_block_idx - m) */
  const double* x0;            /* linearization points, GLOBAL sizes 7/9/7/1 concatenated in block order */
  const double* lin_jac;       /* n x n, COLUMN-major (Eigen::MatrixXd)       */
  const double* lin_res;       /* n                                            */
} bvio_prior;

/* The window = Estimator's Ceres parameter arrays after vector2double()
 * (vins_estimator/src/estimator.cpp:477-519; estimator.h:109-115) plus the
 * measurements optimization() walks (estimator.cpp:702-755).
 * Landmarks are FeatureManager entries that pass
 * used_num >= 2 && start_frame < WINDOW_SIZE-2 (estimator.cpp:712-718), in
 * list order; observation 0 of each landmark is the anchor observation
 * (feature_per_frame[0]), observation k>0 yields one ProjectionFactor. */
typedef struct {
  int32_t K;                    /* keyframes = WINDOW_SIZE+1                   */
  double* para_pose;            /* [K][7]  px py pz qx qy qz qw    (in/out)    */
  double* para_speed_bias;      /* [K][9]  v ba bg                 (in/out)    */
  double* para_ex_pose;         /* [7]     tic, qic xyzw           (in/out)    */
  double* para_td;              /* [1]                             (in/out)    */
  int32_t L;                    /* landmarks                                   */
  double* inv_depth;            /* [L] para_Feature: inverse depth in anchor frame (in/out) */
  const int32_t* lm_obs_offset; /* [L+1] CSR into obs_* arrays                 */
  const int32_t* obs_frame;     /* [n_obs] frame of each observation, ascending per landmark */
  const double* obs_xy;         /* [n_obs][2] normalized-plane point (z == 1)  */
  const double* obs_vel;        /* [n_obs][2] or NULL (estimate_td only)       */
  const double* obs_td;         /* [n_obs]    or NULL (estimate_td only)       */
  const double* obs_row;        /* [n_obs]    or NULL (estimate_td only)       */
  const bvio_preint* preint;    /* [K]; entry 0 unused; entry j links j-1 -> j */
  const bvio_prior* prior;      /* NULL when there is no last_marginalization_info */
  /* ---- ABI v2: relocalization factors (estimator.cpp:760-792), optional ----- */
  /* When relocalization_info is set the reference adds the pose block relo_Pose (free, PoseLocalParameterization) and,
   * for every window landmark matched in the loop-closure frame, one ProjectionFactor(pts_i, match_point) between
   * Pose[start_frame] and relo_Pose with the same Cauchy loss.  n_relo = 0: none.  Not combined with estimate_td
   * (the reference uses the plain ProjectionFactor here even when ESTIMATE_TD is on): BVIO_ERR_UNSUPPORTED.
   * bvio_marginalize ignores these fields, as the reference's marginalization does. */
  int32_t n_relo;
  double* relo_pose;            /* [7] relo_Pose, px py pz qx qy qz qw   (in/out); may be NULL when n_relo == 0 */
  const int32_t* relo_lm;       /* [n_relo] landmark index, strictly ascending  */
  const double* relo_xy;        /* [n_relo][2] match_points x, y (z = 1)        */
} bvio_window;

enum { BVIO_STRATEGY_LM = 0, BVIO_STRATEGY_DOGLEG = 1 };

/* Solver options = the globals optimization() reads (parameters.cpp:45-143)
 * plus Ceres' Solver::Options defaults that the reference leaves untouched
 * (estimator.cpp:794-806).  bvio_default_opts() fills the EuRoC values. */
typedef struct {
  int32_t max_iters;            /* NUM_ITERATIONS (8)                          */
  double  max_time_s;           /* options.max_solver_time_in_seconds (SOLVER_TIME, or 4/5 of it before MARGIN_OLD,
                                 * estimator.cpp:799-806); 0 = no cut.  Checked on the DEVICE clock after every
                                 * iteration: time since this solve's first kernel (H2D copy excluded); a solve cut
                                 * short ends with BVIO_TERM_TIME */
  int32_t estimate_extrinsic;   /* ESTIMATE_EXTRINSIC != 0 frees para_ex_pose  */
  int32_t estimate_td;          /* ESTIMATE_TD                                 */
  double  focal_length;         /* FOCAL_LENGTH 460: sqrt_info = focal/1.5 I2  */
  double  cauchy_a;             /* CauchyLoss(1.0)                             */
  double  G[3];                 /* (0,0,g_norm)                                */
  double  TR, ROW;              /* rolling shutter (td factor only)            */
  double  function_tolerance;   /* 1e-6  */
  double  gradient_tolerance;   /* 1e-10 */
  double  parameter_tolerance;  /* 1e-8  */
  double  initial_radius;       /* 1e4   */
  double  min_relative_decrease;/* 1e-3  */
  int32_t strategy;             /* BVIO_STRATEGY_LM (default) or _DOGLEG       */
  int32_t jacobi_scaling;       /* 1                                            */
} bvio_opts;

enum {
  BVIO_TERM_MAX_ITERS = 0,
  BVIO_TERM_FUNCTION_TOL = 1,
  BVIO_TERM_GRADIENT_TOL = 2,
  BVIO_TERM_PARAMETER_TOL = 3,
  BVIO_TERM_FAILURE = 4,        /* non-finite cost / radius underflow          */
  BVIO_TERM_TIME = 5
};

typedef struct {
  int32_t iterations;           /* linear solves attempted (accepted+rejected) */
  int32_t num_accepted;
  int32_t num_rejected;
  int32_t termination;          /* BVIO_TERM_*                                  */
  double  initial_cost;
  double  final_cost;
  double  final_radius;
  double  final_gradient_max;
  double  device_ms;            /* CUDA-event time of the device work          */
} bvio_summary;

typedef struct bvio_ctx bvio_ctx;

/* Context: one CUDA device, one stream set, scratch in HBM.  Single caller. */
int  bvio_create(int device, bvio_ctx** out);
void bvio_destroy(bvio_ctx* ctx);
const char* bvio_last_error(const bvio_ctx* ctx);
int  bvio_abi_version(void);
void bvio_default_opts(bvio_opts* opts);

/* Replaces ceres::Solve(options,&problem,&summary) at estimator.cpp:809 with
 * the problem optimization() builds at estimator.cpp:663-755.  Overwrites
 * para_* and inv_depth with the solution (what Ceres leaves in the parameter
 * arrays before double2vector(), estimator.cpp:814). */
int bvio_optimize(bvio_ctx* ctx, bvio_window* window, const bvio_opts* opts,
                  bvio_summary* summary);

/* Same, for B independent windows in one launch sequence (throughput mode:
 * replay, multi-agent, multi-hypothesis).  windows[b] and summaries[b] as above. */
int bvio_optimize_batch(bvio_ctx* ctx, bvio_window* windows, int32_t B,
                        const bvio_opts* opts, bvio_summary* summaries);

/* Device-resident batch: upload once, solve many times (benchmarks, streaming).
 * bvio_batch_solve resets every window to its uploaded initial state first. */
typedef struct bvio_batch bvio_batch;
int  bvio_batch_upload(bvio_ctx* ctx, const bvio_window* windows, int32_t B,
                       const bvio_opts* opts, bvio_batch** out);
int  bvio_batch_solve(bvio_ctx* ctx, bvio_batch* batch);               /* async on ctx stream */
int  bvio_batch_download(bvio_ctx* ctx, bvio_batch* batch, bvio_window* windows,
                         bvio_summary* summaries);                      /* syncs */
void bvio_batch_free(bvio_ctx* ctx, bvio_batch* batch);
/* Per-kernel device time of one solve, for roofline reporting: out_ms = {linearize, solve, cost,
 * sum} in ms summed over all passes, out_launches = launches of each kernel.  Synchronous. */
int  bvio_batch_solve_timed(bvio_ctx* ctx, bvio_batch* batch, double out_ms[4], int32_t out_launches[3]);
/* stream the batch runs on (cudaStream_t as void*), for event timing */
void* bvio_stream(bvio_ctx* ctx);
/* kernels launched by this context since creation (for gpu_launches) */
int64_t bvio_launch_count(const bvio_ctx* ctx);

/* Replaces the marginalization tail of optimization(), estimator.cpp:816-991:
 * flag 0 = MARGIN_OLD (drop Pose[0], SpeedBias[0] and landmarks anchored at
 * frame 0), flag 1 = MARGIN_SECOND_NEW (drop Pose[K-2]).  The window holds the
 * post-solve state.  Output blocks already carry the shifted frame indices
 * (addr_shift, estimator.cpp:904-916 / 962-984).  out->* point into
 * caller-provided storage: cap_n >= 15*K + 7 and cap_blocks >= 2*K + 2 always suffice.
 * With opts->estimate_td the dropped frame's factors are ProjectionTdFactors and para_Td is a kept block
 * (estimator.cpp:863-871).  (J, r) come from the eigen-decomposition the reference uses
 * (marginalization_factor.cpp:268-291); the environment variable BVIO_MARG_CHOLESKY=1 selects a pivoted Cholesky
 * factor of the same quadratic form instead (about half the latency). */
typedef struct {
  int32_t n, nblocks;
  int32_t* block_kind; int32_t* block_frame; int32_t* block_idx;
  double* x0; double* lin_jac; double* lin_res;
  int32_t cap_n, cap_blocks;    /* capacities of the arrays above             */
} bvio_prior_out;
int bvio_marginalize(bvio_ctx* ctx, const bvio_window* window, const bvio_opts* opts,
                     int32_t flag, bvio_prior_out* out);

/* The same, split: _begin enqueues the marginalization on the context's second stream and returns at once; _end waits
 * for it and fills `out` (whose arrays must stay valid in between).  The new prior is first needed by the NEXT frame's
 * optimization(), so the caller can run FeatureSelector::select() -- bvio_select on the context's main stream -- or wait
 * for the next image while it computes.  One marginalization in flight per context.  *job is NULL when there was nothing
 * to compute (out->n <= 0); bvio_marginalize_end(ctx, NULL) is a no-op. */
typedef struct bvio_marg_job bvio_marg_job;
int bvio_marginalize_begin(bvio_ctx* ctx, const bvio_window* window, const bvio_opts* opts, int32_t flag,
                           bvio_prior_out* out, bvio_marg_job** job);
int bvio_marginalize_end(bvio_ctx* ctx, bvio_marg_job* job);

/* The step right before optimization() in Estimator::solveOdometry (estimator.cpp:471):
 * FeatureManager::triangulate (feature_manager.cpp:202-257).  For every landmark of the window, the DLT depth in its
 * anchor camera frame from all its observations (right singular vector of the (2 n_obs) x 4 system, svd_V[2]/svd_V[3]);
 * values below 0.1 are replaced by init_depth (INIT_DEPTH = 5.0, parameters.cpp:3).  Uses para_pose, para_ex_pose and
 * the observation CSR only; depth_out[L].  The caller keeps deciding which landmarks need it (estimated_depth <= 0). */
int bvio_triangulate(bvio_ctx* ctx, const bvio_window* window, double init_depth, double* depth_out);

/* IMU preintegration between two frames: IntegrationBase::{push_back, propagate, midPointIntegration} and -- called
 * again with new linearization biases -- repropagate (factor/integration_base.h:30-158), which in the reference runs in
 * Estimator::processIMU (estimator.cpp:86-117).  Sample 0 of a segment is (acc_0, gyr_0) of the constructor (its dt is
 * ignored); samples 1..n_samples-1 are the push_back() calls.  Noise densities ACC_N, GYR_N, ACC_W, GYR_W
 * (parameters.cpp:69-72) fill the 18x18 noise matrix (:21-27). */
typedef struct {
  int32_t n_samples;
  const double* dt;             /* [n_samples]                                   */
  const double* acc;            /* [n_samples][3]                                */
  const double* gyr;            /* [n_samples][3]                                */
  double lin_ba[3], lin_bg[3];  /* linearized_ba / linearized_bg                 */
} bvio_imu_segment;
int bvio_preintegrate(bvio_ctx* ctx, const bvio_imu_segment* segments, int32_t n_segments, double acc_n, double gyr_n,
                      double acc_w, double gyr_w, bvio_preint* out /* [n_segments] */);

/* ------------------------------------------------------------------------ */
/*  Anticipated feature selection (FeatureSelector::select)                  */
/* ------------------------------------------------------------------------ */

/* Pinhole + radtan model used by PinholeCamera::spaceToPlane
 * (camera_model/src/camera_models/PinholeCamera.cc:520-542, 672-688). */
typedef struct {
  double fx, fy, cx, cy, k1, k2, p1, p2;
  int32_t width, height;
} bvio_camera;

/* Omega_PRIOR for bvio_select_in.omega_prior, from the back end's own information instead of the reference's I9
 * (addOmegaPrior, feature_selector.cpp:602-609; named as future work in support_files/report/paper/anticipation.tex
 * :146-152): the 9 x 9 information (row-major; position, velocity, accelerometer bias -- the selector's state order,
 * state_defs.h:15) that the window's prior, IMU and projection factors hold on its NEWEST frame x_k, every other state
 * and all landmarks marginalized out, linearized at the window's current state.  Opt-in: the reference behaviour is
 * omega_prior = NULL.  BVIO_ERR_NUMERIC when the window leaves the other states undetermined. */
int bvio_window_omega_prior(bvio_ctx* ctx, const bvio_window* window, const bvio_opts* opts, double* omega9 /* [81] */);

/* HorizonGenerator::imu (utility/horizon_generator.cpp:25-70), the step before select() in IMU-horizon mode: x_k and the
 * IMU-propagated x_{k+1} open the horizon, then x_{k+1} is propagated nr_imu steps of delta_imu per future frame with the
 * latest accelerometer / gyro sample held constant (bias ba0 of x_k, gravity (0,0,-9.80665), state_defs.h:37-41).
 * Quaternions x y z w, not re-normalised (the reference multiplies by Utility::deltaQ's unnormalised increment).
 * Fills horizon_pos[H+1][3] / horizon_quat[H+1][4] in the layout bvio_select_in takes. */
int bvio_horizon_imu(bvio_ctx* ctx, int32_t H, const double pos0[3], const double quat0[4], const double ba0[3],
                     const double pos1[3], const double quat1[4], const double vel1[3], const double acc[3],
                     const double gyr[3], int32_t nr_imu, double delta_imu, double* horizon_pos, double* horizon_quat);

/* Inputs of the numerical part of select(): everything
 * calcInfoFromRobotMotion / calcInfoFromFeatures / selectInformativeFeatures
 * read (feature_selector.cpp:239-728).  Horizon states are
 * state_horizon_t[0..H] (utility/state_defs.h:18-19), produced by
 * HorizonGenerator::{imu,groundTruth} (utility/horizon_generator.cpp:25-123)
 * which stays host-side input preparation. */
typedef struct {
  int32_t H;                    /* HORIZON (reference compile-time 13)         */
  const double* horizon_pos;    /* [H+1][3]  P_WB of x_k, x_k+1 .. x_k+H       */
  const double* horizon_quat;   /* [H+1][4]  Q_WB  x y z w                     */
  double q_ic[4];               /* x y z w                                     */
  double t_ic[3];
  bvio_camera cam;
  int32_t nr_imu;               /* nrImuMeasurements                           */
  double delta_imu;
  double acc_var;               /* accVarDTime_  (reference passes ACC_N)      */
  double acc_bias_var;          /* accBiasVarDTime_ (reference passes ACC_W)   */
  /* new features (image_new): ascending feature id = std::map order           */
  int32_t N;
  const int32_t* cand_id;       /* [N]                                          */
  const double* cand_xy;        /* [N][2] calibrated pixel (x,y,1)             */
  const double* cand_prob;      /* [N]    fPROB                                 */
  /* already tracked features present in this image (subset)                  */
  int32_t U;
  const int32_t* used_id;       /* [U]                                          */
  const double* used_xy;        /* [U][2]                                       */
  /* depth cloud = initKDTree() output: landmarks projected into frame k+1   */
  int32_t C;
  const double* cloud_xy;       /* [C][2]                                       */
  const double* cloud_depth;    /* [C]  estimated_depth of that landmark       */
  int32_t kappa;                /* max(0, maxFeatures - |subset|)               */
  /* ---- ABI v2: optional (NULL = the v1 behaviour) --------------------------- */
  /* state_k1_ of the reference: the IMU-propagated x_k+1 set by setNextStateFromImuPropagation
   * (feature_selector.cpp:56-70).  calcInfoFromFeatures back-projects every feature with THIS state (:247-266) and
   * builds the k+1 information block from it (:321-326), while state_kkH[1] = horizon[1] only enters the IMU
   * information.  In IMU-horizon mode the two coincide (horizon_generator.cpp:35); in ground-truth-horizon mode
   * (use_ground_truth_hgen: 1, config/euroc/euroc_config.yaml:88; horizon_generator.cpp:106-117) they differ.
   * NULL: horizon[1] is used for both.                                         */
  const double* state_k1_pos;   /* [3]  P_WB  or NULL                           */
  const double* state_k1_quat;  /* [4]  Q_WB x y z w  or NULL                   */
  /* Omega_PRIOR on x_k: 9 x 9 row-major in the selector's state order (position, velocity, accelerometer bias;
   * state_defs.h:25-29), added to the top-left block instead of the reference's I9 (addOmegaPrior,
   * feature_selector.cpp:602-609).  NULL = I9 (the reference).  bvio_prior_omega9() extracts it from a
   * marginalization prior (the report's stated future work, support_files/report/paper/anticipation.tex:146-152). */
  const double* omega_prior;    /* [81] or NULL                                  */
} bvio_select_in;

typedef struct {
  int32_t n_selected;
  int32_t n_candidates_valid;   /* candidates that survive numVisible > 1      */
  int64_t candidates_scored;    /* (candidate, round) log-det gains evaluated  */
  double  final_logdet;         /* logdet(Omega + OmegaS) after the last round */
  double  min_margin;           /* min over rounds of best - second best value */
  double  device_ms;
  /* ---- ABI v2: how the selection ran -------------------------------------- */
  int32_t transport;            /* 0 one GPU (also: sharded call that the planner kept on one GPU),
                                 * 1 fused: all rounds in one kernel, winner records over peer memory (NVLink),
                                 * 2 one kernel per round with an ncclAllGather in between                    */
  int32_t world;                /* ranks that shared the candidates                                          */
  int32_t grid;                 /* CTAs of the scoring kernel                                                */
  int32_t cpw;                  /* candidates per warp of the persistent kernel (0: one kernel per round)    */
  double  round_score_us;       /* persistent kernel, mean per greedy round as seen by CTA 0:                */
  double  round_barrier_us;     /*   scoring + merge + update, wait at the grid barrier,                     */
  double  round_exchange_us;    /*   wait for the peers' records (0 on one GPU)                              */
} bvio_select_summary;

/* Replaces calcInfoFromRobotMotion + addOmegaPrior + 2x calcInfoFromFeatures +
 * selectInformativeFeatures (feature_selector.cpp:139-170).  out_ids receives
 * at most kappa feature ids in SELECTION ORDER (the reference's blacklist,
 * feature_selector.cpp:685); out_values (nullable) the winning log-det of each
 * round. */
int bvio_select(bvio_ctx* ctx, const bvio_select_in* in, int32_t* out_ids,
                double* out_values, bvio_select_summary* summary);

/* Multi-GPU selection: one process per GPU, candidates sharded by contiguous
 * id blocks, one exchange of the winners' information blocks per greedy round.
 * uid is the 128-byte ncclUniqueId created by bvio_nccl_unique_id() on rank 0
 * and distributed by the caller (torch.distributed / MPI / files). */
int bvio_nccl_unique_id(void* uid128);
int bvio_comm_init(bvio_ctx* ctx, const void* uid128, int32_t rank, int32_t world);
int bvio_select_sharded(bvio_ctx* ctx, const bvio_select_in* in, int32_t* out_ids,
                        double* out_values, bvio_select_summary* summary);

/* Device-resident selector problem (benchmarks): upload, run (async), fetch.  */
typedef struct bvio_selprob bvio_selprob;
int  bvio_select_upload(bvio_ctx* ctx, const bvio_select_in* in, bvio_selprob** out);
/* mode 0: this GPU alone (also on a context that has a communicator); 1: shared with the communicator's ranks when the
 * planner finds that sharding pays (what bvio_select_upload / bvio_select_sharded do); 2: shared, always */
int  bvio_select_upload_mode(bvio_ctx* ctx, const bvio_select_in* in, int32_t mode, bvio_selprob** out);
int  bvio_select_run(bvio_ctx* ctx, bvio_selprob* prob);
int  bvio_select_fetch(bvio_ctx* ctx, bvio_selprob* prob, int32_t* out_ids,
                       double* out_values, bvio_select_summary* summary);
void bvio_select_free(bvio_ctx* ctx, bvio_selprob* prob);

/* ------------------------------------------------------------------------ */
/*  Fine-grained device entry points (parity tests of individual rows)      */
/* ------------------------------------------------------------------------ */

/* One linearization of the window at its current state: writes the reduced
 * (Schur-complemented, undamped) system  S[np*np] row-major, g[np], the
 * per-landmark h[L], b[L] and the total cost 0.5*sum rho.  np = 15*K (+6 if
 * estimate_extrinsic, +1 if estimate_td). */
int bvio_debug_linearize(bvio_ctx* ctx, const bvio_window* window, const bvio_opts* opts,
                         double* S, double* g, double* h, double* b, double* cost);

/* Per-candidate compact information blocks: C[N][3H*3H] (row-major, the
 * position rows/cols of frames k+1..k+H of Delta_ell) and valid[N]. */
int bvio_debug_build_delta(bvio_ctx* ctx, const bvio_select_in* in, double* C, int32_t* valid,
                           double* omega /* [9(H+1)]^2 row-major, Omega_kkH incl. prior */);

#ifdef __cplusplus
}
#endif
#endif /* BVIO_H_ */
