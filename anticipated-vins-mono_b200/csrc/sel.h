// sel.h -- device-side layout of one anticipated-feature-selection problem
// (FeatureSelector::select, vins_estimator/src/feature_selector.cpp:74-202), shared by sel_api.cu and
// sel_kernels.cu.
//
// Compact formulation (DESIGN.md "Selector"): Delta_ell is zero outside the 3H position rows/cols of the
// future frames k+1..k+H (feature_selector.cpp:335-359), so with the D = 9(H+1) variables split into
// positions p (T = 3H) and the rest o,
//     logdet(Omega + Omega_S + p Delta_ell) = logdet(M_oo) + logdet(R + p C_ell),
//     R = S0 + sum_selected p C,   S0 = M_pp - M_po M_oo^-1 M_op   (constant over the greedy rounds),
// where C_ell is the T x T position block of Delta_ell.  Each candidate costs one T x T Cholesky per
// round (one warp) instead of the reference's D x D Eigen LLT (utility.h:143-167).
#pragma once
#include "common.cuh"
#include "../../include/bvio.h"

namespace bvio {

constexpr int SEL_WARPS = 8;          // warps per CTA in the build / round kernels
constexpr int SEL_REC_HDR = 4;        // winner record header: value, second, candidate index, prob
constexpr int SEL_MAX_WORLD = 8;      // ranks of the fused (peer-memory) exchange: one NVSwitch domain
constexpr size_t SEL_MBOX_BYTES = 4096;   // per-rank mailbox: [2 parities][world][4] doubles, then [2][world] u64 epoch flags
constexpr size_t SEL_MBOX_FLAG_OFF = 2048;

struct SelCtrl {
  unsigned long long scored;          // (candidate, round) log-dets evaluated (all ranks after finalize)
  double min_margin;                  // min over rounds of best - second best
  double logdet_oo;                   // logdet(M_oo), constant
  double final_logdet;
  int n_selected, round, n_valid, pad;
  unsigned int ticket, pad2;
  int peer_timeout, pad3;             // fused exchange: a peer's record did not arrive within the time limit
  // where a greedy round of the persistent kernel spends its time (thread 0 of CTA 0, summed over the rounds, ns)
  unsigned long long t_score, t_barrier, t_exchange;
};

struct SelProb {
  int H, T, TT, D, Do;                // T = 3H, TT = T(T+1)/2, D = 9(H+1), Do = D - T
  int N, U, C, kappa, nr_imu;
  int c0, c1;                         // candidates scored by this rank: [c0, c1)
  int b0, b1;                         // candidates whose information blocks this rank builds: [c0, c1), or all of
                                      // them in fused mode (the build is 1 % of a select; replicating it means a
                                      // round's exchange is a 32-byte record instead of a 3.7 KB block)
  int fused;                          // multi-GPU rounds inside the persistent kernel, exchange over peer memory
  unsigned long long epoch_base;      // flags carry epoch_base + round + 1: never reset, never ambiguous
  double* mbox; unsigned long long* mflag;                     // this rank's mailbox (peers write into it)
  double* peer_mbox[SEL_MAX_WORLD]; unsigned long long* peer_flag[SEL_MAX_WORLD];   // every rank's mailbox, peer-mapped
  int rank, world;
  int grid_round;                     // CTAs of the round kernel
  int grid_persist, cpw;              // persistent single-kernel path: CTAs, candidates per warp (0 = not used)
  int c_smem;                         // persistent path: the warps' information blocks are staged in shared memory (else
                                      // they are read from L2 / HBM every round: large candidate sets)
  double delta_imu, acc_var, acc_bias_var;
  double q_ic[4], t_ic[3];
  double k1_pos[3], k1_quat[4];       // state_k1_ (feature_selector.cpp:247-250): = horizon[1] unless the caller gave one
  int has_prior;                      // omega_prior given: added instead of I9 (feature_selector.cpp:602-609)
  double omega_prior[81];
  bvio_camera cam;
  const double* hpos;                 // [H+1][3]
  const double* hquat;                // [H+1][4] xyzw
  const double2* cand_xy; const double* cand_prob;       // [N]
  const double2* used_xy;             // [U]
  const double2* cloud_xy; const double* cloud_depth;    // [C]
  double* Cc;                         // [N][TT] packed-lower position blocks of Delta_ell (local range filled)
  double* Cu;                         // [U][TT]
  int* valid;                         // [N] numVisible > 1 (local range)
  int* valid_u;                       // [U]
  int* taken;                         // [N] selected in an earlier round
  double* depth;                      // [N+U] NN depth used (debug)
  double* pair;                       // [H][4*81] per-pair Omega, A, A^T Omega, A^T Omega A
  double* omega;                      // [D*D] Omega_kkH incl. prior (row-major), before the used features
  double* R;                          // [TT] packed lower: S0 + selected
  double* blk_best;                   // [grid_round][4] value, second, idx, count
  double* rec_send;                   // [SEL_REC_HDR + TT] this rank's winner record
  double* rec_all;                    // [world][SEL_REC_HDR + TT] gathered records
  int* out_idx; double* out_val;      // [kappa] candidate INDEX (host maps to ids) and winning log-det
  SelCtrl* ctrl;
};

int sel_configure(void);
int sel_launch_reset(const SelProb& sp, cudaStream_t st);
int sel_launch_build(const SelProb& sp, cudaStream_t st);      // build_delta + omega/Schur
int sel_launch_round(const SelProb& sp, cudaStream_t st);      // score + local winner (+ apply when world == 1)
int sel_plan_persist(SelProb& sp, int sm_count);             // 1 when the persistent path fits
int sel_launch_persist(const SelProb& sp, cudaStream_t st);   // all kappa rounds in one kernel (single GPU)
int sel_launch_apply(const SelProb& sp, cudaStream_t st);      // multi-GPU: pick among gathered records
int sel_launch_final(const SelProb& sp, cudaStream_t st);
int sel_launch_expand(const SelProb& sp, double* Cfull, cudaStream_t st);   // debug: packed -> dense [N][T*T]
size_t sel_omega_smem_bytes(int H);

}  // namespace bvio
