// ba.h -- device-side data layout of a batch of sliding windows in HBM, shared by the host
// API (bvio_api.cu) and the BA kernels (ba_kernels.cu).
//
// State layout (DESIGN.md "Data layout in HBM"): reduced system dimension np = 15*K, frame-major
// 15-blocks [dp(3) dtheta(3) dv(3) dba(3) dbg(3)].  Visual factors only touch the 6 pose dims of
// each frame, so the landmark-parallel Schur kernel produces a compact 6K x 6K block matrix.
#pragma once
#include "common.cuh"

namespace bvio {

constexpr int IMU_REC = 288;      // doubles per compact IMU record (287 used, padded to 16 B multiple)
// compact IMU record offsets (doubles)
constexpr int IR_DP = 0, IR_DQ = 3, IR_DV = 7, IR_BA = 10, IR_BG = 13, IR_DT = 16, IR_DPDBA = 17, IR_DPDBG = 26,
              IR_DQDBG = 35, IR_DVDBA = 44, IR_DVDBG = 53, IR_SQ = 62;  // sqrt_info 15x15 row-major at 62..286
constexpr int PREINT_DOUBLES = 467;   // sizeof(bvio_preint)/8
constexpr int IMU_OUT = 496;          // per IMU factor: J^T J lower packed (465) + J^T r (30) + cost (1)

constexpr int PRIOR_MAXB = 32;    // max kept parameter blocks in a prior
constexpr int BA_THREADS = 256;   // linearize / cost kernels
// ba_solve runs 256 threads x 2 CTAs per SM (throughput) or 512 x 1 (latency, BaBatch::solve_wide)

struct BaCtrl {                   // per-window LM state, lives in HBM
  double cost;                    // cost at X[cur]
  double cand_cost;
  double radius, decrease_factor;
  double model_pose;              // pose part of the model cost change
  double gmax;                    // max |gradient| at X[cur]
  double initial_cost;
  double rho;
  int cur;                        // which of the two state buffers is current
  int done, termination;
  int iterations, accepted, rejected;
  int first;                      // 1 until the first linearization has fixed the Jacobi scaling
  int solve_ok;                   // a step was computed this pass (Cholesky succeeded)
  int invalid_run;
  int stepped;                    // this pass computed a step that the cost kernel must judge
  unsigned int ticket;            // CTAs of the cost / dogleg kernel that finished (last one decides)
  int pad;
  // BVIO_STRATEGY_DOGLEG (Ceres DoglegStrategy, TRADITIONAL_DOGLEG)
  double mu;                      // Gauss-Newton regularisation (min_mu 1e-8, x10 on failure, /5 on success)
  double dsum[6];                 // pose parts of the dogleg dot products (ba_solve -> ba_dogleg)
  double ca, cb;                  // step = ca * (scaled gradient direction t) + cb * (Gauss-Newton step)
  double step_norm;               // |step| in Ceres' diagonally scaled space (drives the radius update)
  unsigned long long t0_ns;       // %globaltimer at the reset of this solve (max_solver_time_in_seconds cut)
  unsigned long long stamps[6];   // BVIO_DEBUG: phase time stamps of the last ba_solve pass (assembly, vectors, factor, ...)
};

// per-(window,tile) output record of ba_linearize, in doubles:
//   [0 .. NPb*36)   S blocks (frame p >= q), 6x6 row-major each: visual H_pp minus the Schur term
//   then 6K: reduced gradient, 6K: unreduced gradient, 6K: diag of visual H_pp (no Schur term)
//   then 4 scalars: cost, max |b_l|, pad, pad
__host__ __device__ inline int tile_rec_doubles(int K) { return (K * (K + 1) / 2) * 36 + 18 * K + 4; }
// per-(window,tile) output record of ba_cost: cost, model_lm, step2, x2
constexpr int COST_REC = 4;
// per-(window,tile) output record of ba_dogleg: |g_y|^2, t^T H t, |gn_y|^2, g^T n, n^T D n, t^T D n (landmark parts)
constexpr int DOG_REC = 8;

struct BaBatch {                  // all pointers are device pointers
  int B, K, np, T;                // windows, keyframes, reduced dimension, landmark tiles per window (cost / dogleg kernels)
  int TL;                         // landmark tiles per window of the linearization (tile records); <= T
  int total_L, total_obs, nmax;   // nmax = max prior dimension (row stride of the prior arrays)
  int chunk_l;                    // landmarks per chunk in ba_linearize (bounded by shared memory)
  int undamped;                   // debug: linearize without LM damping (bvio_debug_linearize)
  // solver options
  int max_iters, jacobi_scaling;
  int strategy;                   // BVIO_STRATEGY_LM / BVIO_STRATEGY_DOGLEG
  int solve_wide;                 // fewer windows than SMs: ba_solve with 512 threads per window
  int use_mma;                    // linearize with the FP64 tensor-core (DMMA) kernel (extrinsics fixed)
  int use_ws;                     // ... its warp-specialised variant (one 512-thread CTA per SM; throughput mode)
  int est_ex;                     // estimate_extrinsic: +6 dims, extrinsic block = pseudo-frame K of the visual layout
  int est_td;                     // estimate_td: +1 dim (after the extrinsic block), pseudo-frame K + est_ex
  double sqrt_info, cauchy_a, G[3];
  double function_tolerance, gradient_tolerance, parameter_tolerance, initial_radius, min_relative_decrease;
  double max_time_s;              // options.max_solver_time_in_seconds (estimator.cpp:799-806); 0 = no cut
  // structure
  const int* lm_base;             // [B+1] first global landmark of each window
  const int* lm_off;              // [total_L+1] CSR: global observation offsets
  const int* obs_frame;           // [total_obs]
  const double2* obs_xy;          // [total_obs]
  // state, double buffered
  double* pose[2];                // [B][K][7]
  double* sb[2];                  // [B][K][9]
  const double* ex;               // [B][7] uploaded extrinsic pose
  double* exs[2];                 // [B][7] extrinsic state, double buffered (both = ex when it is constant)
  double* ex_out;                 // [B][7]
  const double* td0;              // [B] uploaded camera-IMU time offset
  double* tds[2];                 // [B] td state, double buffered
  double* td_out;                 // [B]
  const double2* obs_vel;         // [total_obs] feature velocity on the normalized plane (estimate_td only)
  const double* obs_shift;        // [total_obs] -td_obs + TR/ROW (row - ROW/2) (estimate_td only)
  double* invd[2];                // [total_L]
  const double* pose0; const double* sb0; const double* invd0;   // uploaded initial state (for reset)
  double* pose_out; double* sb_out; double* invd_out;            // final state gathered from X[cur]
  // IMU
  const double* preint_raw;       // [B][K][467] as uploaded (bvio_preint)
  double* imu;                    // [B][K][IMU_REC] compact records (entry 0 unused); sum_dt > 10 => skipped
  double* imu_out;                // [B][K][IMU_OUT]
  // prior
  const int* pr_n; const int* pr_nb;       // [B]
  const int* pr_kind; const int* pr_frame; const int* pr_idx;   // [B][PRIOR_MAXB]
  const double* pr_x0;            // [B][PRIOR_MAXB*9]
  const double* pr_jac;           // [B][nmax*nmax] column-major n x n
  const double* pr_res;           // [B][nmax]
  double* pr_H;                   // [B][nmax*nmax] J^T J (symmetric)
  int* pr_map;                    // [B][nmax] prior column -> reduced state index (-1: constant block)
  double* pr_out;                 // [B][nmax+1] J^T r at X[cur], then 0.5 |r|^2
  double* S0;                     // [B][(np+1)(np+2)/2] latency mode only (else null): the prior's J^T J already scattered
                                  // into the packed lower reduced layout -- ba_solve starts from it instead of from zero
  int* pr_inv;                    // [B][np] reduced state index -> prior column (-1: none)
  // linearization products
  double* h; double* b; double* sl2;   // [total_L]
  double* w;                      // [total_obs][6]
  double* wex;                    // [total_L][est_ex + est_td][6] extra-block parts of the landmark's coupling row
  double* tile_out;               // [B][T][tile_rec_doubles(K)]
  double* cost_out;               // [B][T+1][COST_REC]
  double* delta_p;                // [B][np]  LM step / Gauss-Newton step (dogleg)
  double* dog_t;                  // [B][np]  dogleg: scaled gradient direction t = scale^2 g / D^2 (pose part)
  double* dog_l;                  // [total_L][2] dogleg: landmark parts (t_l, Gauss-Newton dlambda_l)
  double* dog_out;                // [B][T][DOG_REC] per-tile partial dot products
  double* scale_p;                // [B][np] Jacobi scaling of the pose/speed-bias columns
  double* solve_vec;              // [B][5][np] ba_solve's gradient / diagonal / damping vectors
  double* dbg_S; double* dbg_g;   // [B][np*np], [B][np] (debug linearize only, else null)
  BaCtrl* ctrl;                   // [B]
};

// launchers (ba_kernels.cu); all asynchronous on `st`; return number of kernels launched
int ba_launch_prepare(const BaBatch& bt, cudaStream_t st);
int ba_launch_reset(const BaBatch& bt, cudaStream_t st);
int ba_launch_iteration(const BaBatch& bt, cudaStream_t st, bool with_step, cudaEvent_t* ev = nullptr);  // ev[4]: before/after each kernel
int ba_launch_finish(const BaBatch& bt, cudaStream_t st);
size_t ba_linearize_smem_bytes(int K, int chunk_l, int n_extra_blocks);
int ba_pick_chunk(int K, int n_extra_blocks);
size_t ba_solve_smem_bytes(int np);
size_t ba_linearize_mma_smem_bytes(int K);
size_t ba_marginalize_smem_bytes(int K, int nmax, int n);
int ba_launch_marginalize(const BaBatch& bt, int flag, int m, int n, int n0, const int* dropidx, const int* keepidx, double* A,
                          double* b, double* part, double* out_jac, double* out_res, int* status, int method, cudaStream_t st);
size_t ba_marginalize_part_doubles(int K, int G);   // scratch of the factor kernel's partial sums
int ba_marginalize_groups(int n0);                  // CTAs of the factor kernel for n0 landmarks anchored at the dropped frame
int ba_launch_marginal9(const double* S, int np, int frame, double* out81, int* status, cudaStream_t st);
int ba_configure(void);   // cudaFuncSetAttribute for the big-smem kernels; returns cudaError_t
void ba_ws_prof_dump(void);   // development builds (make WSPROF=1) print the warp-specialised linearize kernel's phase cycles; otherwise a no-op

}  // namespace bvio
