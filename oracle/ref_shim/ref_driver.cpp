// ref_driver.cpp -- C entry points around the REFERENCE's own factor sources, compiled from where they lie under
// /root/reference (see oracle/Makefile target `ref`; output oracle/_ref/libvins_ref.so).  TEST INFRASTRUCTURE ONLY:
// it pins the oracle's restatement (oracle/ba_oracle.cpp) to the arithmetic of the reference's code.
//
// What runs reference code here:  ProjectionFactor::Evaluate, ProjectionTdFactor::Evaluate, IMUFactor::Evaluate,
// IntegrationBase::{push_back, propagate, midPointIntegration, repropagate, evaluate}, PoseLocalParameterization::Plus,
// ResidualBlockInfo::Evaluate (loss corrector), MarginalizationInfo::{addResidualBlockInfo, preMarginalize, marginalize,
// getParameterBlocks}, MarginalizationFactor::Evaluate, Utility::{deltaQ, Qleft, Qright, skewSymmetric, logdet}.
// What is ours: the Eigen / Ceres / ROS stand-in headers next to this file (the libraries are not installed here), and
// the glue below that feeds the factors -- for marginalization it restates which residual blocks
// Estimator::optimization() hands to MarginalizationInfo and with which drop sets (estimator.cpp:816-991), because
// estimator.cpp itself needs the whole ROS / OpenCV / Ceres stack.
#include <cstring>
#include <iostream>
#include <map>
#include <vector>
using namespace std;   // integration_base.h prints with unqualified cout / endl

#include "factor/imu_factor.h"
#include "factor/marginalization_factor.h"
#include "factor/pose_local_parameterization.h"
#include "factor/projection_factor.h"
#include "factor/projection_td_factor.h"

#include "utility/horizon_generator.h"
#include "feature_manager.h"
#include "feature_selector.h"

#include "../../include/bvio.h"

// globals of parameters.cpp that the factor code reads
double ACC_N, ACC_W, GYR_N, GYR_W;
Eigen::Vector3d G;
double TR, ROW, COL, TD;
double INIT_DEPTH = 5.0, MIN_PARALLAX = 10.0 / 460.0;
int ESTIMATE_EXTRINSIC = 0, ESTIMATE_TD = 0, NUM_ITERATIONS = 8;
double SOLVER_TIME = 0.04;
std::vector<Eigen::Matrix3d> RIC;
std::vector<Eigen::Vector3d> TIC;

namespace {
Eigen::Vector3d v3(const double* p) { return Eigen::Vector3d(p[0], p[1], p[2]); }

void load(IntegrationBase& ib, const bvio_preint* pre) {
  ib.delta_p = v3(pre->delta_p);
  ib.delta_q = Eigen::Quaterniond(pre->delta_q[3], pre->delta_q[0], pre->delta_q[1], pre->delta_q[2]);
  ib.delta_v = v3(pre->delta_v);
  ib.sum_dt = pre->sum_dt;
  for (int i = 0; i < 15; ++i) for (int j = 0; j < 15; ++j) {
    ib.jacobian(i, j) = pre->jacobian[i * 15 + j];
    ib.covariance(i, j) = pre->covariance[i * 15 + j];
  }
}
void store(const IntegrationBase& ib, bvio_preint* pre) {
  for (int i = 0; i < 3; ++i) { pre->delta_p[i] = ib.delta_p(i); pre->delta_v[i] = ib.delta_v(i); pre->lin_ba[i] = ib.linearized_ba(i); pre->lin_bg[i] = ib.linearized_bg(i); }
  pre->delta_q[0] = ib.delta_q.x(); pre->delta_q[1] = ib.delta_q.y(); pre->delta_q[2] = ib.delta_q.z(); pre->delta_q[3] = ib.delta_q.w();
  pre->sum_dt = ib.sum_dt;
  for (int i = 0; i < 15; ++i) for (int j = 0; j < 15; ++j) {
    pre->jacobian[i * 15 + j] = ib.jacobian(i, j);
    pre->covariance[i * 15 + j] = ib.covariance(i, j);
  }
}

// The window copied into Estimator-like parameter arrays (estimator.h:109-115) with stable addresses.
struct Params {
  int K, L;
  std::vector<double> pose, sb, feat;
  double ex[7], td[1];
  explicit Params(const bvio_window* w) : K(w->K), L(w->L), pose(w->para_pose, w->para_pose + 7 * w->K),
      sb(w->para_speed_bias, w->para_speed_bias + 9 * w->K), feat(w->inv_depth, w->inv_depth + w->L) {
    memcpy(ex, w->para_ex_pose, sizeof ex); td[0] = w->para_td ? w->para_td[0] : 0.0;
  }
  double* Pose(int i) { return &pose[7 * i]; }
  double* SpeedBias(int i) { return &sb[9 * i]; }
  double* Feature(int l) { return &feat[l]; }
  double* block(int kind, int frame) {
    return kind == BVIO_BLK_POSE ? Pose(frame) : kind == BVIO_BLK_SPEEDBIAS ? SpeedBias(frame) : kind == BVIO_BLK_EXPOSE ? ex : td;
  }
  bool identify(const double* p, int* kind, int* frame) {
    for (int i = 0; i < K; ++i) {
      if (p == Pose(i)) { *kind = BVIO_BLK_POSE; *frame = i; return true; }
      if (p == SpeedBias(i)) { *kind = BVIO_BLK_SPEEDBIAS; *frame = i; return true; }
    }
    if (p == ex) { *kind = BVIO_BLK_EXPOSE; *frame = 0; return true; }
    if (p == td) { *kind = BVIO_BLK_TD; *frame = 0; return true; }
    return false;
  }
};

int global_size(int kind) { return kind == BVIO_BLK_POSE || kind == BVIO_BLK_EXPOSE ? 7 : kind == BVIO_BLK_SPEEDBIAS ? 9 : 1; }

// last_marginalization_info + last_marginalization_parameter_blocks rebuilt from the C-ABI's prior
MarginalizationInfo* info_from_prior(const bvio_prior* pr, Params& P, std::vector<double*>* blocks, std::vector<double>* x0_store) {
  MarginalizationInfo* info = new MarginalizationInfo();
  info->n = pr->n; info->m = 0;
  int tot = 0; for (int b = 0; b < pr->nblocks; ++b) tot += global_size(pr->block_kind[b]);
  x0_store->assign(pr->x0, pr->x0 + tot);
  int off = 0;
  for (int b = 0; b < pr->nblocks; ++b) {
    int gs = global_size(pr->block_kind[b]);
    info->keep_block_size.push_back(gs);
    info->keep_block_idx.push_back(pr->block_idx[b]);
    info->keep_block_data.push_back(x0_store->data() + off);
    blocks->push_back(P.block(pr->block_kind[b], pr->block_frame[b]));
    off += gs;
  }
  info->linearized_jacobians.resize(pr->n, pr->n);
  info->linearized_residuals.resize(pr->n);
  for (int i = 0; i < pr->n; ++i) { info->linearized_residuals(i) = pr->lin_res[i]; for (int j = 0; j < pr->n; ++j) info->linearized_jacobians(i, j) = pr->lin_jac[(size_t)j * pr->n + i]; }
  return info;
}
}  // namespace

extern "C" {

void ref_projection_factor(const double pts_i[3], const double pts_j[3], const double pose_i[7], const double pose_j[7],
                           const double ex_pose[7], double inv_dep, double sqrt_info, double res[2], double* jac_i,
                           double* jac_j, double* jac_ex, double* jac_f) {
  ProjectionFactor::sqrt_info = sqrt_info * Eigen::Matrix2d::Identity();
  ProjectionFactor f(v3(pts_i), v3(pts_j));
  const double* params[4] = {pose_i, pose_j, ex_pose, &inv_dep};
  double* jac[4] = {jac_i, jac_j, jac_ex, jac_f};
  f.Evaluate(params, res, jac);
}

void ref_projection_td_factor(const double pts_i[3], const double pts_j[3], const double vel_i[2], const double vel_j[2],
                              double td_i, double td_j, double row_i, double row_j, double TR_, double ROW_,
                              const double pose_i[7], const double pose_j[7], const double ex_pose[7], double inv_dep,
                              double td, double sqrt_info, double res[2], double* jac_i, double* jac_j, double* jac_ex,
                              double* jac_f, double* jac_td) {
  TR = TR_; ROW = ROW_;
  ProjectionTdFactor::sqrt_info = sqrt_info * Eigen::Matrix2d::Identity();
  ProjectionTdFactor f(v3(pts_i), v3(pts_j), Eigen::Vector2d(vel_i[0], vel_i[1]), Eigen::Vector2d(vel_j[0], vel_j[1]), td_i, td_j, row_i, row_j);
  const double* params[5] = {pose_i, pose_j, ex_pose, &inv_dep, &td};
  double* jac[5] = {jac_i, jac_j, jac_ex, jac_f, jac_td};
  f.Evaluate(params, res, jac);
}

void ref_imu_factor(const bvio_preint* pre, const double G_[3], const double pose_i[7], const double sb_i[9],
                    const double pose_j[7], const double sb_j[9], double res[15], double* jac_pi, double* jac_sbi,
                    double* jac_pj, double* jac_sbj) {
  G = v3(G_);
  IntegrationBase ib(Eigen::Vector3d::Zero(), Eigen::Vector3d::Zero(), v3(pre->lin_ba), v3(pre->lin_bg));
  load(ib, pre);
  IMUFactor f(&ib);
  const double* params[4] = {pose_i, sb_i, pose_j, sb_j};
  double* jac[4] = {jac_pi, jac_sbi, jac_pj, jac_sbj};
  f.Evaluate(params, res, jac);
}

void ref_preint_propagate(bvio_preint* pre, double dt, const double acc_0[3], const double gyr_0[3], const double acc_1[3],
                          const double gyr_1[3], double acc_n, double gyr_n, double acc_w, double gyr_w) {
  ACC_N = acc_n; GYR_N = gyr_n; ACC_W = acc_w; GYR_W = gyr_w;
  IntegrationBase ib(v3(acc_0), v3(gyr_0), v3(pre->lin_ba), v3(pre->lin_bg));
  load(ib, pre);
  ib.propagate(dt, v3(acc_1), v3(gyr_1));
  store(ib, pre);
}

// n push_backs from scratch: acc / gyr hold n + 1 samples (sample 0 starts the interval); then, when new_ba != NULL,
// IntegrationBase::repropagate(new_ba, new_bg)
void ref_preintegrate(int n, const double* dt, const double* acc, const double* gyr, const double ba[3], const double bg[3],
                      double acc_n, double gyr_n, double acc_w, double gyr_w, const double* new_ba, const double* new_bg,
                      bvio_preint* out) {
  ACC_N = acc_n; GYR_N = gyr_n; ACC_W = acc_w; GYR_W = gyr_w;
  IntegrationBase ib(v3(acc), v3(gyr), v3(ba), v3(bg));
  for (int i = 0; i < n; ++i) ib.push_back(dt[i], v3(acc + 3 * (i + 1)), v3(gyr + 3 * (i + 1)));
  if (new_ba) ib.repropagate(v3(new_ba), v3(new_bg));
  store(ib, out);
}

void ref_pose_plus(const double x[7], const double delta[6], double out[7]) {
  ceres::LocalParameterization* p = new PoseLocalParameterization();
  p->Plus(x, delta, out);
  delete p;
}

double ref_logdet(const double* M, int n) {
  Eigen::MatrixXd A(n, n);
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) A(i, j) = M[(size_t)i * n + j];
  return Utility::logdet(A, true);
}

// ResidualBlockInfo::Evaluate on one ProjectionFactor with CauchyLoss(a): loss-corrected residual and Jacobians
void ref_projection_block_corrected(const double pts_i[3], const double pts_j[3], const double pose_i[7], const double pose_j[7],
                                    const double ex_pose[7], double inv_dep, double sqrt_info, double cauchy_a, double res[2],
                                    double jac_i[14], double jac_j[14], double jac_ex[14], double jac_f[2]) {
  ProjectionFactor::sqrt_info = sqrt_info * Eigen::Matrix2d::Identity();
  double pi[7], pj[7], ex[7], f[1] = {inv_dep};
  memcpy(pi, pose_i, sizeof pi); memcpy(pj, pose_j, sizeof pj); memcpy(ex, ex_pose, sizeof ex);
  ceres::CauchyLoss loss(cauchy_a);
  ResidualBlockInfo rb(new ProjectionFactor(v3(pts_i), v3(pts_j)), &loss, std::vector<double*>{pi, pj, ex, f}, std::vector<int>{0, 3});
  rb.Evaluate();
  for (int i = 0; i < 2; ++i) res[i] = rb.residuals(i);
  double* outs[4] = {jac_i, jac_j, jac_ex, jac_f};
  for (int b = 0; b < 4; ++b) for (int i = 0; i < rb.jacobians[b].rows(); ++i) for (int j = 0; j < rb.jacobians[b].cols(); ++j)
    outs[b][i * rb.jacobians[b].cols() + j] = rb.jacobians[b](i, j);
  delete rb.cost_function; delete[] rb.raw_jacobians;
}

// MarginalizationFactor::Evaluate of the window's prior at the window's state: residual [n] and the Jacobian in local
// coordinates, dense row-major [n][n] (columns = the prior's own block_idx layout)
int ref_prior_eval(const bvio_prior* pr, const bvio_window* w, double* res, double* jac) {
  Params P(w);
  std::vector<double*> blocks; std::vector<double> x0;
  MarginalizationInfo* info = info_from_prior(pr, P, &blocks, &x0);
  MarginalizationFactor f(info);
  int n = pr->n;
  std::vector<std::vector<double>> J(blocks.size());
  std::vector<double*> jp;
  for (size_t b = 0; b < blocks.size(); ++b) { J[b].assign((size_t)n * info->keep_block_size[b], 0.0); jp.push_back(J[b].data()); }
  f.Evaluate(blocks.data(), res, jp.data());
  if (jac) {
    std::fill(jac, jac + (size_t)n * n, 0.0);
    for (size_t b = 0; b < blocks.size(); ++b) {
      int gs = info->keep_block_size[b], ls = info->localSize(gs), idx = info->keep_block_idx[b];
      for (int i = 0; i < n; ++i) for (int j = 0; j < ls; ++j) jac[(size_t)i * n + idx + j] = J[b][(size_t)i * gs + j];
      for (int i = 0; i < n; ++i) for (int j = ls; j < gs; ++j) if (J[b][(size_t)i * gs + j] != 0.0) return -2;   // 7th column must be zero
    }
  }
  info->keep_block_data.clear();
  return 0;
}

// Estimator::optimization()'s marginalization step (estimator.cpp:816-991) with the reference's MarginalizationInfo.
int ref_marginalize(const bvio_window* w, const bvio_opts* o, int flag, bvio_prior_out* out) {
  Params P(w);
  const int K = P.K, WS = K - 1;
  G = v3(o->G); TR = o->TR; ROW = o->ROW;
  ProjectionFactor::sqrt_info = o->focal_length / 1.5 * Eigen::Matrix2d::Identity();
  ProjectionTdFactor::sqrt_info = o->focal_length / 1.5 * Eigen::Matrix2d::Identity();
  ceres::LossFunction* loss_function = new ceres::CauchyLoss(o->cauchy_a);
  std::vector<double*> last_blocks; std::vector<double> x0;
  MarginalizationInfo* last = w->prior ? info_from_prior(w->prior, P, &last_blocks, &x0) : nullptr;
  std::vector<IntegrationBase*> keep_alive;
  MarginalizationInfo* info = new MarginalizationInfo();
  std::unordered_map<long, double*> addr_shift;
  if (flag == 0) {
    if (last) {
      std::vector<int> drop_set;
      for (int i = 0; i < (int)last_blocks.size(); ++i) if (last_blocks[i] == P.Pose(0) || last_blocks[i] == P.SpeedBias(0)) drop_set.push_back(i);
      info->addResidualBlockInfo(new ResidualBlockInfo(new MarginalizationFactor(last), NULL, last_blocks, drop_set));
    }
    if (w->preint[1].sum_dt < 10.0) {
      IntegrationBase* ib = new IntegrationBase(Eigen::Vector3d::Zero(), Eigen::Vector3d::Zero(), v3(w->preint[1].lin_ba), v3(w->preint[1].lin_bg));
      load(*ib, &w->preint[1]); keep_alive.push_back(ib);
      info->addResidualBlockInfo(new ResidualBlockInfo(new IMUFactor(ib), NULL,
          std::vector<double*>{P.Pose(0), P.SpeedBias(0), P.Pose(1), P.SpeedBias(1)}, std::vector<int>{0, 1}));
    }
    for (int l = 0; l < P.L; ++l) {
      int o0 = w->lm_obs_offset[l], o1 = w->lm_obs_offset[l + 1];
      int imu_i = w->obs_frame[o0];
      if (imu_i != 0) continue;
      Eigen::Vector3d pts_i(w->obs_xy[2 * o0], w->obs_xy[2 * o0 + 1], 1.0);
      for (int k = o0 + 1; k < o1; ++k) {
        int imu_j = w->obs_frame[k];
        Eigen::Vector3d pts_j(w->obs_xy[2 * k], w->obs_xy[2 * k + 1], 1.0);
        if (o->estimate_td) {
          ProjectionTdFactor* f_td = new ProjectionTdFactor(pts_i, pts_j, Eigen::Vector2d(w->obs_vel[2 * o0], w->obs_vel[2 * o0 + 1]),
              Eigen::Vector2d(w->obs_vel[2 * k], w->obs_vel[2 * k + 1]), w->obs_td[o0], w->obs_td[k], w->obs_row[o0], w->obs_row[k]);
          info->addResidualBlockInfo(new ResidualBlockInfo(f_td, loss_function,
              std::vector<double*>{P.Pose(imu_i), P.Pose(imu_j), P.ex, P.Feature(l), P.td}, std::vector<int>{0, 3}));
        } else {
          info->addResidualBlockInfo(new ResidualBlockInfo(new ProjectionFactor(pts_i, pts_j), loss_function,
              std::vector<double*>{P.Pose(imu_i), P.Pose(imu_j), P.ex, P.Feature(l)}, std::vector<int>{0, 3}));
        }
      }
    }
    for (int i = 1; i <= WS; ++i) {
      addr_shift[reinterpret_cast<long>(P.Pose(i))] = P.Pose(i - 1);
      addr_shift[reinterpret_cast<long>(P.SpeedBias(i))] = P.SpeedBias(i - 1);
    }
  } else {
    bool has = false;
    for (double* b : last_blocks) has |= (b == P.Pose(WS - 1));
    if (!last || !has) { out->n = -1; out->nblocks = 0; return 0; }
    std::vector<int> drop_set;
    for (int i = 0; i < (int)last_blocks.size(); ++i) {
      if (last_blocks[i] == P.SpeedBias(WS - 1)) return -3;                 // ROS_ASSERT, estimator.cpp:937
      if (last_blocks[i] == P.Pose(WS - 1)) drop_set.push_back(i);
    }
    info->addResidualBlockInfo(new ResidualBlockInfo(new MarginalizationFactor(last), NULL, last_blocks, drop_set));
    for (int i = 0; i <= WS; ++i) {
      if (i == WS - 1) continue;
      int to = i == WS ? i - 1 : i;
      addr_shift[reinterpret_cast<long>(P.Pose(i))] = P.Pose(to);
      addr_shift[reinterpret_cast<long>(P.SpeedBias(i))] = P.SpeedBias(to);
    }
  }
  addr_shift[reinterpret_cast<long>(P.ex)] = P.ex;
  if (o->estimate_td) addr_shift[reinterpret_cast<long>(P.td)] = P.td;
  info->preMarginalize();
  info->marginalize();
  std::vector<double*> parameter_blocks = info->getParameterBlocks(addr_shift);

  int n = info->n, nb = (int)parameter_blocks.size();
  if (n > out->cap_n || nb > out->cap_blocks) return -4;
  out->n = n; out->nblocks = nb;
  int off = 0;
  for (int b = 0; b < nb; ++b) {
    int kind, frame;
    if (!P.identify(parameter_blocks[b], &kind, &frame)) return -5;
    out->block_kind[b] = kind; out->block_frame[b] = frame; out->block_idx[b] = info->keep_block_idx[b] - info->m;
    int gs = info->keep_block_size[b];
    for (int i = 0; i < gs; ++i) out->x0[off + i] = info->keep_block_data[b][i];
    off += gs;
  }
  for (int i = 0; i < n; ++i) { out->lin_res[i] = info->linearized_residuals(i); for (int j = 0; j < n; ++j) out->lin_jac[(size_t)j * n + i] = info->linearized_jacobians(i, j); }
  return 0;   // the MarginalizationInfo objects are leaked on purpose (their destructor frees with mismatched delete)
}

// HorizonGenerator::imu (utility/horizon_generator.cpp:25-70).  The reference's horizon length is the compile-time
// constant HORIZON (state_defs.h:8).
int ref_horizon_length(void) { return HORIZON; }

void ref_horizon_imu(int H, const double pos0[3], const double quat0[4], const double ba0[3], const double pos1[3],
                     const double quat1[4], const double vel1[3], const double acc[3], const double gyr[3], int nr_imu,
                     double delta_imu, double* horizon_pos, double* horizon_quat) {
  if (H != HORIZON) std::abort();
  HorizonGenerator hg{ros::NodeHandle()};
  state_t s0, s1;
  s0.first.setZero(); s1.first.setZero();
  s0.first.segment<3>(xPOS) = v3(pos0); s0.first.segment<3>(xB_A) = v3(ba0);
  s0.second = Eigen::Quaterniond(quat0[3], quat0[0], quat0[1], quat0[2]);
  s1.first.segment<3>(xPOS) = v3(pos1); s1.first.segment<3>(xVEL) = v3(vel1); s1.first.segment<3>(xB_A) = v3(ba0);
  s1.second = Eigen::Quaterniond(quat1[3], quat1[0], quat1[1], quat1[2]);
  state_horizon_t hz = hg.imu(s0, s1, v3(acc), v3(gyr), nr_imu, delta_imu);
  for (int h = 0; h <= HORIZON; ++h) {
    for (int i = 0; i < 3; ++i) horizon_pos[3 * h + i] = hz[h].first(xPOS + i);
    horizon_quat[4 * h + 0] = hz[h].second.x(); horizon_quat[4 * h + 1] = hz[h].second.y();
    horizon_quat[4 * h + 2] = hz[h].second.z(); horizon_quat[4 * h + 3] = hz[h].second.w();
  }
}

// HorizonGenerator ground-truth mode (horizon_generator.cpp:74-123, 169-210): the constructor loads the EuRoC-style
// csv named by the "gt_data_csv" parameter; every call advances the generator's private seek cursor.
void* ref_horizon_gt_open(const char* csv_path) {
  ros::NodeHandle nh;
  nh.setParam("gt_data_csv", csv_path);
  return new HorizonGenerator(nh);
}
void ref_horizon_gt(void* handle, double timestamp0, const double pos0[3], const double quat0[4], double delta_frame,
                    double* horizon_pos, double* horizon_quat) {
  HorizonGenerator* hg = static_cast<HorizonGenerator*>(handle);
  state_t s0;
  s0.first.setZero();
  s0.first(xTIMESTAMP) = timestamp0;
  s0.first.segment<3>(xPOS) = v3(pos0);
  s0.second = Eigen::Quaterniond(quat0[3], quat0[0], quat0[1], quat0[2]);
  state_horizon_t hz = hg->groundTruth(s0, s0, delta_frame);
  for (int h = 0; h <= HORIZON; ++h) {
    for (int i = 0; i < 3; ++i) horizon_pos[3 * h + i] = hz[h].first(xPOS + i);
    horizon_quat[4 * h + 0] = hz[h].second.x(); horizon_quat[4 * h + 1] = hz[h].second.y();
    horizon_quat[4 * h + 2] = hz[h].second.z(); horizon_quat[4 * h + 3] = hz[h].second.w();
  }
}
void ref_horizon_gt_close(void* handle) { delete static_cast<HorizonGenerator*>(handle); }

// ---- FeatureManager (feature_manager.cpp) -----------------------------------------------------------------------------------
struct RefFM {
  Eigen::Matrix3d Rs[WINDOW_SIZE + 1];
  Eigen::Vector3d Ps[WINDOW_SIZE + 1];
  Eigen::Matrix3d ric[NUM_OF_CAM];
  Eigen::Vector3d tic[NUM_OF_CAM];
  FeatureManager fm;
  RefFM() : fm(Rs) {}
};
int ref_window_size(void) { return WINDOW_SIZE; }
void* ref_fm_create(double init_depth, double min_parallax) { INIT_DEPTH = init_depth; MIN_PARALLAX = min_parallax; return new RefFM(); }
void ref_fm_destroy(void* h) { delete static_cast<RefFM*>(h); }
// addFeatureCheckParallax: pts [n][7] = x y z u v vx vy; returns 1 when the second-newest frame is a keyframe (MARGIN_OLD)
int ref_fm_add_frame(void* h, int frame_count, int n, const int32_t* ids, const double* pts, double td) {
  map<int, vector<pair<int, Eigen::Matrix<double, 7, 1>>>> image;
  for (int i = 0; i < n; ++i) {
    Eigen::Matrix<double, 7, 1> v;
    for (int k = 0; k < 7; ++k) v(k) = pts[7 * i + k];
    image[ids[i]].emplace_back(0, v);
  }
  return static_cast<RefFM*>(h)->fm.addFeatureCheckParallax(frame_count, image, td) ? 1 : 0;
}
int ref_fm_last_track_num(void* h) { return static_cast<RefFM*>(h)->fm.last_track_num; }
// poses [n_frames][7] (p, q xyzw) of the window's frames, extrinsic [7]
void ref_fm_set_poses(void* h, int n_frames, const double* poses, const double* ex) {
  RefFM* r = static_cast<RefFM*>(h);
  for (int i = 0; i < n_frames && i <= WINDOW_SIZE; ++i) {
    r->Ps[i] = v3(poses + 7 * i);
    r->Rs[i] = Eigen::Quaterniond(poses[7 * i + 6], poses[7 * i + 3], poses[7 * i + 4], poses[7 * i + 5]).toRotationMatrix();
  }
  r->tic[0] = v3(ex);
  r->ric[0] = Eigen::Quaterniond(ex[6], ex[3], ex[4], ex[5]).toRotationMatrix();
  r->fm.setRic(r->ric);
}
void ref_fm_triangulate(void* h) { RefFM* r = static_cast<RefFM*>(h); r->fm.triangulate(r->Ps, r->tic, r->ric); }
int ref_fm_feature_count(void* h) { return static_cast<RefFM*>(h)->fm.getFeatureCount(); }
void ref_fm_get_depth_vector(void* h, double* out) {
  Eigen::VectorXd d = static_cast<RefFM*>(h)->fm.getDepthVector();
  for (int i = 0; i < d.size(); ++i) out[i] = d(i);
}
void ref_fm_set_depth(void* h, int n, const double* x) {
  Eigen::VectorXd v(n);
  for (int i = 0; i < n; ++i) v(i) = x[i];
  static_cast<RefFM*>(h)->fm.setDepth(v);
}
void ref_fm_remove_failures(void* h) { static_cast<RefFM*>(h)->fm.removeFailures(); }
// slideWindowOld with shift_depth (estimator.cpp:1089-1106): camera poses from the IMU poses of the dropped and the new frame 0
void ref_fm_remove_back_shift_depth(void* h, const double* pose_marg, const double* pose_new) {
  RefFM* r = static_cast<RefFM*>(h);
  Eigen::Matrix3d back_R0 = Eigen::Quaterniond(pose_marg[6], pose_marg[3], pose_marg[4], pose_marg[5]).toRotationMatrix();
  Eigen::Matrix3d Rs0 = Eigen::Quaterniond(pose_new[6], pose_new[3], pose_new[4], pose_new[5]).toRotationMatrix();
  Eigen::Matrix3d R0 = back_R0 * r->ric[0], R1 = Rs0 * r->ric[0];
  Eigen::Vector3d P0 = v3(pose_marg) + back_R0 * r->tic[0], P1 = v3(pose_new) + Rs0 * r->tic[0];
  r->fm.removeBackShiftDepth(R0, P0, R1, P1);
}
void ref_fm_remove_front(void* h, int frame_count) { static_cast<RefFM*>(h)->fm.removeFront(frame_count); }
// dump: per feature id, start_frame, number of observations, estimated_depth, solve_flag; returns the count
int ref_fm_dump(void* h, int cap, int32_t* ids, int32_t* start, int32_t* nobs, double* depth) {
  int k = 0;
  for (auto& f : static_cast<RefFM*>(h)->fm.feature) {
    if (k < cap) { ids[k] = f.feature_id; start[k] = f.start_frame; nobs[k] = (int)f.feature_per_frame.size(); depth[k] = f.estimated_depth; }
    ++k;
  }
  return k;
}

}  // extern "C"

// ---- FeatureSelector (feature_selector.cpp) -----------------------------------------------------------------------------------
// estimator.cpp / initial_ex_rotation.cpp are not compiled (they need the full ROS / OpenCV / Ceres stack); the selector
// only reads data members of Estimator, so the two constructors its type needs are defined here, empty.
InitialEXRotation::InitialEXRotation() {}
// the initializer (initial/*.cpp: OpenCV-based SfM and alignment) is out of scope and never reached by the driver
GlobalSFM::GlobalSFM() {}
bool GlobalSFM::construct(int, Quaterniond*, Vector3d*, int, const Matrix3d, const Vector3d, vector<SFMFeature>&, map<int, Vector3d>&) { std::abort(); }
bool InitialEXRotation::CalibrationExRotation(vector<pair<Vector3d, Vector3d>>, Quaterniond, Matrix3d&) { std::abort(); }
bool VisualIMUAlignment(map<double, ImageFrame>&, Vector3d*, Vector3d&, VectorXd&) { std::abort(); }
bool MotionEstimator::solveRelativeRT(const vector<pair<Vector3d, Vector3d>>&, Matrix3d&, Vector3d&) { std::abort(); }

namespace {
struct RefSel {
  Estimator est;
  std::unique_ptr<FeatureSelector> sel;
};
image_t make_image(int n, const int32_t* ids, const double* xy, const double* prob) {
  image_t image;
  for (int i = 0; i < n; ++i) {
    Eigen::Matrix<double, fSIZE, 1> v;
    v.setZero();
    v(0) = xy[2 * i]; v(1) = xy[2 * i + 1]; v(2) = 1.0; v(fPROB) = prob ? prob[i] : 1.0;
    image[ids[i]].emplace_back(0, v);
  }
  return image;
}
}  // namespace

extern "C" {

void* ref_sel_create_gt(const bvio_camera* cam, const double q_ic[4], const double t_ic[3], double acc_var, double acc_bias_var,
                        int max_features, int init_thresh, const char* gt_csv);
void* ref_sel_create(const bvio_camera* cam, const double q_ic[4], const double t_ic[3], double acc_var, double acc_bias_var,
                     int max_features, int init_thresh) {
  return ref_sel_create_gt(cam, q_ic, t_ic, acc_var, acc_bias_var, max_features, init_thresh, nullptr);
}
// gt_csv != NULL: ground-truth horizon mode (USE_GT, the "gt_data_csv" parameter of the node)
void* ref_sel_create_gt(const bvio_camera* cam, const double q_ic[4], const double t_ic[3], double acc_var, double acc_bias_var,
                        int max_features, int init_thresh, const char* gt_csv) {
  camodocal::PinholeParams& c = camodocal::CameraFactory::registered();
  c.fx = cam->fx; c.fy = cam->fy; c.cx = cam->cx; c.cy = cam->cy; c.k1 = cam->k1; c.k2 = cam->k2; c.p1 = cam->p1; c.p2 = cam->p2;
  c.width = cam->width; c.height = cam->height;
  RefSel* r = new RefSel();
  r->est.ric[0] = Eigen::Quaterniond(q_ic[3], q_ic[0], q_ic[1], q_ic[2]).toRotationMatrix();
  r->est.tic[0] = v3(t_ic);
  r->est.solver_flag = Estimator::INITIAL;
  ros::NodeHandle nh;
  if (gt_csv) nh.setParam("gt_data_csv", gt_csv);
  r->sel.reset(new FeatureSelector(nh, r->est, "unused.yaml"));
  r->sel->setParameters(acc_var, acc_bias_var, true, max_features, init_thresh, gt_csv != nullptr);
  return r;
}
void ref_sel_destroy(void* h) { delete static_cast<RefSel*>(h); }

// back-end state the selector reads: window poses [WINDOW_SIZE+1][7] (p, q xyzw), velocity and accelerometer bias of the
// newest frame, and the landmarks of FeatureManager (anchor observation, depth, solve_flag) for initKDTree
void ref_sel_set_backend(void* h, const double* poses, const double* vel_last, const double* ba_last, int n_lm,
                         const int32_t* lm_id, const int32_t* lm_start, const int32_t* lm_nobs, const double* lm_xy,
                         const double* lm_depth, const int32_t* lm_solve_flag) {
  RefSel* r = static_cast<RefSel*>(h);
  for (int i = 0; i <= WINDOW_SIZE; ++i) {
    r->est.Ps[i] = v3(poses + 7 * i);
    r->est.Rs[i] = Eigen::Quaterniond(poses[7 * i + 6], poses[7 * i + 3], poses[7 * i + 4], poses[7 * i + 5]).toRotationMatrix();
  }
  r->est.Vs[WINDOW_SIZE] = v3(vel_last);
  r->est.Bas[WINDOW_SIZE] = v3(ba_last);
  r->est.f_manager.feature.clear();
  for (int l = 0; l < n_lm; ++l) {
    FeaturePerId f(lm_id[l], lm_start[l]);
    Eigen::Matrix<double, 7, 1> pt;
    pt.setZero(); pt(0) = lm_xy[2 * l]; pt(1) = lm_xy[2 * l + 1]; pt(2) = 1.0;
    for (int k = 0; k < lm_nobs[l]; ++k) f.feature_per_frame.push_back(FeaturePerFrame(pt, 0.0));
    f.estimated_depth = lm_depth[l];
    f.solve_flag = lm_solve_flag[l];
    r->est.f_manager.feature.push_back(f);
  }
}

// One FeatureSelector::select().  stamp = (sec, nsec) of the image header; the state of frame k+1 as handed to
// setNextStateFromImuPropagation.  image = n features (id, x, y, prob).  Returns the number of newly selected ids
// (written to out_selected in selection order); *n_tracked = size of the tracked list after the call; the ids left in
// `image` (what the back end receives) are written to out_image (capacity n), their count to *n_image.
int ref_sel_select(void* h, int initialized, unsigned stamp_sec, unsigned stamp_nsec, const double P1[3], const double Q1[4],
                   const double V1[3], const double a1[3], const double w1[3], const double Ba1[3], int nr_imu, int n,
                   const int32_t* ids, const double* xy, const double* prob, int32_t* out_selected, int32_t* n_tracked,
                   int32_t* out_image, int32_t* n_image) {
  RefSel* r = static_cast<RefSel*>(h);
  r->est.solver_flag = initialized ? Estimator::NON_LINEAR : Estimator::INITIAL;
  std_msgs::Header header;
  header.stamp.sec = stamp_sec; header.stamp.nsec = stamp_nsec;
  r->sel->setNextStateFromImuPropagation(header.stamp.toSec(), v3(P1), Eigen::Quaterniond(Q1[3], Q1[0], Q1[1], Q1[2]), v3(V1),
                                         v3(a1), v3(w1), v3(Ba1));
  image_t image = make_image(n, ids, xy, prob);
  auto res = r->sel->select(image, header, nr_imu);
  for (size_t i = 0; i < res.second.size(); ++i) out_selected[i] = res.second[i];
  *n_tracked = (int)res.first.size();
  int k = 0;
  for (const auto& f : image) out_image[k++] = f.first;
  *n_image = k;
  return (int)res.second.size();
}

}  // extern "C"


// ---- Estimator::optimization() around an injected solve (estimator.cpp:477-610, 661-994) -----------------------------------------
namespace {
struct EstHook {
  Estimator* est = nullptr;
  const double *inj_pose = nullptr, *inj_sb = nullptr, *inj_ex = nullptr, *inj_feat = nullptr, *inj_td = nullptr;
  double *entry_pose = nullptr, *entry_sb = nullptr, *entry_ex = nullptr, *entry_feat = nullptr;
  double entry_cost = 0;
  int counts[8] = {0};    // prior, imu, projection, projection_td, other, ex constant, parameter blocks, residual blocks
  int options[4] = {0};   // max_num_iterations, DOGLEG?, DENSE_SCHUR?, _
  double max_time = 0;
  int L = 0;
  std::vector<double> H, g;   // normal equations of the whole problem at entry, local coordinates
  int dim = 0;
  bool session = false;
  ceres::Problem* live = nullptr;      // valid only while the solve callback runs
  void (*callback)(void) = nullptr;    // external trust-region driver (tests): called instead of injecting a solution
  const double* inj_relo = nullptr;    // relocalization: solution for relo_Pose
  int n_relo_blocks = 0;               // ProjectionFactors attached to relo_Pose in the recorded problem
} g_hook;
void (*g_solve_callback)(void) = nullptr;
}  // namespace
// When set (adapters/vins/interpose.cpp does, in libvins_bvio.so), ref_estimator_optimization() hands the Estimator it
// built to this function instead of installing the injecting solve hook: whoever set it answers ceres::Solve.
extern "C" { void (*ref_on_estimator_created)(void* estimator) = nullptr; }
namespace {
// relocalization inputs for the next ref_estimator_optimization() call (what setReloFrame() leaves in the Estimator)
struct PendingRelo { bool on = false; std::vector<Eigen::Vector3d> match; double pose[7]; int local_index = 0; double out_pose[7]; double rel_t[3]; double rel_yaw = 0; } g_relo;

void est_solve_hook(const ceres::Solver::Options& o, ceres::Problem* pb, ceres::Solver::Summary*) {
  EstHook& h = g_hook;
  Estimator& e = *h.est;
  if (h.session) {                      // a live Estimator fed frame by frame (ref_est_*): only the external solve
    h.L = e.f_manager.getFeatureCount();
    h.live = pb;
    if (g_solve_callback) g_solve_callback();
    h.live = nullptr;
    return;
  }
  memcpy(h.entry_pose, e.para_Pose, sizeof(double) * 7 * (WINDOW_SIZE + 1));
  memcpy(h.entry_sb, e.para_SpeedBias, sizeof(double) * 9 * (WINDOW_SIZE + 1));
  memcpy(h.entry_ex, e.para_Ex_Pose, sizeof(double) * 7);
  for (int l = 0; l < h.L; ++l) h.entry_feat[l] = e.para_Feature[l][0];
  // the objective Ceres would minimise, through the reference's own cost functions: 1/2 sum rho(|r|^2)
  double cost = 0;
  for (auto& rb : pb->residual_blocks) {
    std::vector<double> r(rb.cost->num_residuals());
    rb.cost->Evaluate(rb.blocks.data(), r.data(), nullptr);
    double s = 0; for (double x : r) s += x * x;
    if (rb.loss) { double rho[3]; rb.loss->Evaluate(s, rho); cost += 0.5 * rho[0]; } else cost += 0.5 * s;
    if (dynamic_cast<MarginalizationFactor*>(rb.cost)) h.counts[0]++;
    else if (dynamic_cast<IMUFactor*>(rb.cost)) h.counts[1]++;
    else if (dynamic_cast<ProjectionFactor*>(rb.cost)) h.counts[2]++;
    else if (dynamic_cast<ProjectionTdFactor*>(rb.cost)) h.counts[3]++;
    else h.counts[4]++;
  }
  h.entry_cost = cost;
  // Gauss-Newton normal equations J^T J, J^T r of the whole problem at entry, every block evaluated and loss-corrected
  // by the reference's own ResidualBlockInfo::Evaluate.  Local column layout: frame-major [pose 6 | speed-bias 9] x K,
  // extrinsic 6, td 1, then one column per feature.
  {
    // relocalization (estimator.cpp:760-792): the extra pose block relo_Pose takes 6 more columns after the features
    const int K1 = WINDOW_SIZE + 1, npar = 15 * K1 + 7, dim = npar + h.L + (e.relocalization_info ? 6 : 0);
    auto col_of = [&](double* p, int* size) {
      for (int i = 0; i < K1; ++i) { if (p == e.para_Pose[i]) { *size = 6; return 15 * i; } if (p == e.para_SpeedBias[i]) { *size = 9; return 15 * i + 6; } }
      if (p == e.para_Ex_Pose[0]) { *size = 6; return 15 * K1; }
      if (p == e.para_Td[0]) { *size = 1; return 15 * K1 + 6; }
      if (p == e.relo_Pose) { *size = 6; return npar + h.L; }
      for (int l = 0; l < h.L; ++l) if (p == e.para_Feature[l]) { *size = 1; return npar + l; }
      *size = 0; return -1;
    };
    h.n_relo_blocks = 0;
    for (auto& rb : pb->residual_blocks) for (double* p : rb.blocks) if (p == e.relo_Pose) h.n_relo_blocks++;
    h.dim = dim; h.H.assign((size_t)dim * dim, 0.0); h.g.assign(dim, 0.0);
    for (auto& rb : pb->residual_blocks) {
      ResidualBlockInfo info(rb.cost, rb.loss, rb.blocks, std::vector<int>{});
      info.Evaluate();
      const int nr = rb.cost->num_residuals(), nb = (int)rb.blocks.size();
      std::vector<int> c0(nb), sz(nb);
      for (int b = 0; b < nb; ++b) c0[b] = col_of(rb.blocks[b], &sz[b]);
      for (int a = 0; a < nb; ++a) {
        if (c0[a] < 0) continue;
        for (int i = 0; i < sz[a]; ++i) {
          double gi = 0; for (int r = 0; r < nr; ++r) gi += info.jacobians[a](r, i) * info.residuals(r);
          h.g[c0[a] + i] += gi;
          for (int b = 0; b < nb; ++b) {
            if (c0[b] < 0) continue;
            for (int j = 0; j < sz[b]; ++j) {
              double v = 0; for (int r = 0; r < nr; ++r) v += info.jacobians[a](r, i) * info.jacobians[b](r, j);
              h.H[(size_t)(c0[a] + i) * dim + c0[b] + j] += v;
            }
          }
        }
      }
      delete[] info.raw_jacobians;
    }
  }
  for (auto& b : pb->parameter_blocks) if (b.ptr == e.para_Ex_Pose[0] && b.constant) h.counts[5] = 1;
  h.counts[6] = (int)pb->parameter_blocks.size(); h.counts[7] = (int)pb->residual_blocks.size();
  h.options[0] = o.max_num_iterations; h.options[1] = o.trust_region_strategy_type == ceres::DOGLEG;
  h.options[2] = o.linear_solver_type == ceres::DENSE_SCHUR; h.max_time = o.max_solver_time_in_seconds;
  if (g_solve_callback) {              // an external driver iterates on the live problem through ref_live_*
    h.live = pb;
    g_solve_callback();
    h.live = nullptr;
    return;
  }
  // "the solve": overwrite the parameter blocks with the solution computed elsewhere
  memcpy(e.para_Pose, h.inj_pose, sizeof(double) * 7 * (WINDOW_SIZE + 1));
  memcpy(e.para_SpeedBias, h.inj_sb, sizeof(double) * 9 * (WINDOW_SIZE + 1));
  memcpy(e.para_Ex_Pose, h.inj_ex, sizeof(double) * 7);
  for (int l = 0; l < h.L; ++l) e.para_Feature[l][0] = h.inj_feat[l];
  if (h.inj_td) e.para_Td[0][0] = h.inj_td[0];
  if (h.inj_relo) memcpy(e.relo_Pose, h.inj_relo, sizeof(double) * 7);
}
}  // namespace

extern "C" {

// ---- the live problem, for a trust-region loop driven from outside while Estimator::optimization() waits in "Solve" ----------
// Local column layout as in est_solve_hook: [pose 6 | speed-bias 9] x (WINDOW_SIZE+1), extrinsic 6, td 1, one per feature.
// Global state layout: para_Pose 7 x K1, para_SpeedBias 9 x K1, para_Ex_Pose 7, para_Td 1, para_Feature L.
void ref_set_solve_callback(void (*cb)(void)) { g_solve_callback = cb; }
int ref_live_dims(int32_t* n_residuals, int32_t* n_local, int32_t* n_global) {
  if (!g_hook.live) return -1;
  int nr = 0; for (auto& rb : g_hook.live->residual_blocks) nr += rb.cost->num_residuals();
  const int K1 = WINDOW_SIZE + 1, relo = g_hook.est->relocalization_info ? 1 : 0;   // relo_Pose: last 6 local / 7 global entries
  *n_residuals = nr; *n_local = 15 * K1 + 7 + g_hook.L + 6 * relo; *n_global = 16 * K1 + 8 + g_hook.L + 7 * relo;
  return 0;
}
void ref_live_get_state(double* x) {
  Estimator& e = *g_hook.est; const int K1 = WINDOW_SIZE + 1;
  memcpy(x, e.para_Pose, sizeof(double) * 7 * K1); memcpy(x + 7 * K1, e.para_SpeedBias, sizeof(double) * 9 * K1);
  memcpy(x + 16 * K1, e.para_Ex_Pose, sizeof(double) * 7); x[16 * K1 + 7] = e.para_Td[0][0];
  for (int l = 0; l < g_hook.L; ++l) x[16 * K1 + 8 + l] = e.para_Feature[l][0];
  if (e.relocalization_info) memcpy(x + 16 * K1 + 8 + g_hook.L, e.relo_Pose, sizeof(double) * 7);
}
void ref_live_set_state(const double* x) {
  Estimator& e = *g_hook.est; const int K1 = WINDOW_SIZE + 1;
  memcpy(e.para_Pose, x, sizeof(double) * 7 * K1); memcpy(e.para_SpeedBias, x + 7 * K1, sizeof(double) * 9 * K1);
  memcpy(e.para_Ex_Pose, x + 16 * K1, sizeof(double) * 7); e.para_Td[0][0] = x[16 * K1 + 7];
  for (int l = 0; l < g_hook.L; ++l) e.para_Feature[l][0] = x[16 * K1 + 8 + l];
  if (e.relocalization_info) memcpy(e.relo_Pose, x + 16 * K1 + 8 + g_hook.L, sizeof(double) * 7);
}
// x (+) delta through the problem's own LocalParameterization objects (PoseLocalParameterization::Plus), plain addition elsewhere
void ref_live_plus(const double* x, const double* delta, double* out) {
  const int K1 = WINDOW_SIZE + 1, relo = g_hook.est->relocalization_info ? 1 : 0, ng = 16 * K1 + 8 + g_hook.L + 7 * relo;
  for (int i = 0; i < ng; ++i) out[i] = x[i];
  ceres::LocalParameterization* lp = nullptr;
  for (auto& b : g_hook.live->parameter_blocks) if (b.local) { lp = b.local; break; }
  for (int i = 0; i < K1; ++i) {
    lp->Plus(x + 7 * i, delta + 15 * i, out + 7 * i);
    for (int a = 0; a < 9; ++a) out[7 * K1 + 9 * i + a] = x[7 * K1 + 9 * i + a] + delta[15 * i + 6 + a];
  }
  lp->Plus(x + 16 * K1, delta + 15 * K1, out + 16 * K1);
  out[16 * K1 + 7] = x[16 * K1 + 7] + delta[15 * K1 + 6];
  for (int l = 0; l < g_hook.L; ++l) out[16 * K1 + 8 + l] = x[16 * K1 + 8 + l] + delta[15 * K1 + 7 + l];
  if (relo) lp->Plus(x + 16 * K1 + 8 + g_hook.L, delta + 15 * K1 + 7 + g_hook.L, out + 16 * K1 + 8 + g_hook.L);
}
// residuals r [n_residuals] and Jacobian J [n_residuals][n_local] (row-major), loss-corrected by the reference's
// ResidualBlockInfo::Evaluate, at the state currently in the parameter arrays; *cost = 1/2 sum rho(|r_uncorrected|^2)
int ref_live_evaluate(double* J, double* r, double* cost) {
  if (!g_hook.live) return -1;
  Estimator& e = *g_hook.est;
  const int K1 = WINDOW_SIZE + 1, npar = 15 * K1 + 7, nl = npar + g_hook.L + (e.relocalization_info ? 6 : 0);
  auto col_of = [&](double* p, int* size) {
    for (int i = 0; i < K1; ++i) { if (p == e.para_Pose[i]) { *size = 6; return 15 * i; } if (p == e.para_SpeedBias[i]) { *size = 9; return 15 * i + 6; } }
    if (p == e.para_Ex_Pose[0]) { *size = 6; return 15 * K1; }
    if (p == e.para_Td[0]) { *size = 1; return 15 * K1 + 6; }
    if (p == e.relo_Pose) { *size = 6; return npar + g_hook.L; }
    for (int l = 0; l < g_hook.L; ++l) if (p == e.para_Feature[l]) { *size = 1; return npar + l; }
    *size = 0; return -1;
  };
  int row = 0; double c = 0;
  for (auto& rb : g_hook.live->residual_blocks) {
    const int nr = rb.cost->num_residuals();
    std::vector<double> raw(nr);
    rb.cost->Evaluate(rb.blocks.data(), raw.data(), nullptr);
    double s = 0; for (double x : raw) s += x * x;
    if (rb.loss) { double rho[3]; rb.loss->Evaluate(s, rho); c += 0.5 * rho[0]; } else c += 0.5 * s;
    ResidualBlockInfo info(rb.cost, rb.loss, rb.blocks, std::vector<int>{});
    info.Evaluate();
    for (int k = 0; k < nr; ++k) { r[row + k] = info.residuals(k); for (int j = 0; j < nl; ++j) J[(size_t)(row + k) * nl + j] = 0.0; }
    for (int b = 0; b < (int)rb.blocks.size(); ++b) {
      int sz, c0 = col_of(rb.blocks[b], &sz);
      if (c0 < 0) return -2;
      for (int k = 0; k < nr; ++k) for (int j = 0; j < sz; ++j) J[(size_t)(row + k) * nl + c0 + j] += info.jacobians[b](k, j);
    }
    delete[] info.raw_jacobians;
    row += nr;
  }
  *cost = c;
  return 0;
}

// Relocalization for the NEXT ref_estimator_optimization() call: n matches (landmark index = feature id there, ascending;
// x, y on the normalized plane of the loop-closure frame), the initial relo_Pose and relo_frame_local_index.
void ref_estimator_set_relo(int n, const int32_t* lm, const double* xy, const double* relo_pose, int local_index) {
  g_relo.on = true; g_relo.match.clear(); g_relo.local_index = local_index;
  for (int k = 0; k < n; ++k) g_relo.match.push_back(Eigen::Vector3d(xy[2 * k], xy[2 * k + 1], (double)lm[k]));
  memcpy(g_relo.pose, relo_pose, sizeof g_relo.pose);
}
// after that call: relo_Pose as the Estimator holds it, relo_relative_t / relo_relative_yaw computed by double2vector
// (estimator.cpp:589-607), and how many residual blocks the recorded problem attached to relo_Pose
int ref_estimator_get_relo(double* relo_pose, double* rel_t, double* rel_yaw) {
  memcpy(relo_pose, g_relo.out_pose, sizeof g_relo.out_pose);
  for (int a = 0; a < 3; ++a) rel_t[a] = g_relo.rel_t[a];
  *rel_yaw = g_relo.rel_yaw;
  return g_hook.n_relo_blocks;
}

// Normal equations of the problem the last ref_estimator_optimization() call handed to Ceres (see est_solve_hook)
int ref_estimator_last_normal(double* H, double* g, int cap_dim) {
  if (g_hook.dim > cap_dim) return -g_hook.dim;
  std::copy(g_hook.H.begin(), g_hook.H.end(), H);
  std::copy(g_hook.g.begin(), g_hook.g.end(), g);
  return g_hook.dim;
}

// Runs the reference's Estimator::optimization() on a window (K must be WINDOW_SIZE + 1) with ceres::Solve replaced by
// "write `solved` into the parameter blocks".  entry_* = what vector2double() produced; scal = {objective at entry,
// max_solver_time}; counts / options as filled by the hook; post_* = Ps / Rs (row-major 3x3) / Vs / Bas / Bgs, extrinsic
// (tic, ric row-major), td, per-landmark estimated_depth after double2vector(); prior = the new
// last_marginalization_info (n = -1 when the reference kept the old one).
int ref_estimator_optimization(const bvio_window* w, const bvio_opts* o, int flag, const bvio_window* solved, double* entry_pose,
                               double* entry_sb, double* entry_ex, double* entry_feat, double* scal, int32_t* counts, int32_t* options,
                               double* post_P, double* post_R, double* post_V, double* post_Ba, double* post_Bg, double* post_ex,
                               double* post_td, double* post_depth, bvio_prior_out* prior) {
  if (w->K != WINDOW_SIZE + 1 || w->L > NUM_OF_F) return -1;
  const int K = w->K, L = w->L;
  ESTIMATE_EXTRINSIC = o->estimate_extrinsic; ESTIMATE_TD = o->estimate_td; NUM_ITERATIONS = o->max_iters;
  SOLVER_TIME = o->max_time_s; TD = w->para_td ? w->para_td[0] : 0.0; TR = o->TR; ROW = o->ROW; G = v3(o->G);
  void* mem = calloc(1, sizeof(Estimator));            // the reference's Estimator lives in zero-initialised static storage
  Estimator* e = new (mem) Estimator();
  ProjectionFactor::sqrt_info = o->focal_length / 1.5 * Eigen::Matrix2d::Identity();
  ProjectionTdFactor::sqrt_info = o->focal_length / 1.5 * Eigen::Matrix2d::Identity();
  for (int i = 0; i < K; ++i) {
    const double* p = w->para_pose + 7 * i; const double* s = w->para_speed_bias + 9 * i;
    e->Ps[i] = v3(p); e->Rs[i] = Eigen::Quaterniond(p[6], p[3], p[4], p[5]).toRotationMatrix();
    e->Vs[i] = v3(s); e->Bas[i] = v3(s + 3); e->Bgs[i] = v3(s + 6);
    if (i >= 1) {
      e->pre_integrations[i] = new IntegrationBase(Eigen::Vector3d::Zero(), Eigen::Vector3d::Zero(), v3(w->preint[i].lin_ba), v3(w->preint[i].lin_bg));
      load(*e->pre_integrations[i], &w->preint[i]);
    }
  }
  e->tic[0] = v3(w->para_ex_pose);
  e->ric[0] = Eigen::Quaterniond(w->para_ex_pose[6], w->para_ex_pose[3], w->para_ex_pose[4], w->para_ex_pose[5]).toRotationMatrix();
  e->f_manager.setRic(e->ric);
  e->td = TD;
  e->frame_count = WINDOW_SIZE; e->solver_flag = Estimator::NON_LINEAR;
  e->marginalization_flag = flag == 0 ? Estimator::MARGIN_OLD : Estimator::MARGIN_SECOND_NEW;
  for (int l = 0; l < L; ++l) {
    int o0 = w->lm_obs_offset[l], o1 = w->lm_obs_offset[l + 1];
    FeaturePerId f(l, w->obs_frame[o0]);
    for (int k = o0; k < o1; ++k) {
      Eigen::Matrix<double, 7, 1> pt;
      pt.setZero(); pt(0) = w->obs_xy[2 * k]; pt(1) = w->obs_xy[2 * k + 1]; pt(2) = 1.0;
      if (w->obs_vel) { pt(4) = w->obs_row[k]; pt(5) = w->obs_vel[2 * k]; pt(6) = w->obs_vel[2 * k + 1]; }
      f.feature_per_frame.push_back(FeaturePerFrame(pt, w->obs_td ? w->obs_td[k] : 0.0));
    }
    f.estimated_depth = 1.0 / w->inv_depth[l];
    f.solve_flag = 1;
    e->f_manager.feature.push_back(f);
  }
  std::vector<double> x0;
  if (w->prior) {
    const bvio_prior* pr = w->prior;
    MarginalizationInfo* info = new MarginalizationInfo();
    info->n = pr->n; info->m = 0;
    int tot = 0; for (int b = 0; b < pr->nblocks; ++b) tot += global_size(pr->block_kind[b]);
    x0.assign(pr->x0, pr->x0 + tot);
    int off = 0;
    for (int b = 0; b < pr->nblocks; ++b) {
      int kind = pr->block_kind[b], fr = pr->block_frame[b], gs = global_size(kind);
      info->keep_block_size.push_back(gs); info->keep_block_idx.push_back(pr->block_idx[b]); info->keep_block_data.push_back(x0.data() + off);
      e->last_marginalization_parameter_blocks.push_back(kind == BVIO_BLK_POSE ? e->para_Pose[fr] : kind == BVIO_BLK_SPEEDBIAS ? e->para_SpeedBias[fr]
                                                         : kind == BVIO_BLK_EXPOSE ? e->para_Ex_Pose[0] : e->para_Td[0]);
      off += gs;
    }
    info->linearized_jacobians.resize(pr->n, pr->n); info->linearized_residuals.resize(pr->n);
    for (int i = 0; i < pr->n; ++i) { info->linearized_residuals(i) = pr->lin_res[i]; for (int j = 0; j < pr->n; ++j) info->linearized_jacobians(i, j) = pr->lin_jac[(size_t)j * pr->n + i]; }
    e->last_marginalization_info = info;
  }
  if (g_relo.on) {                       // the state setReloFrame() (estimator.cpp:1109-1127) leaves behind
    e->relocalization_info = 1; e->relo_frame_local_index = g_relo.local_index; e->match_points = g_relo.match;
    memcpy(e->relo_Pose, g_relo.pose, sizeof g_relo.pose);
    e->prev_relo_t = Eigen::Vector3d(0, 0, 0); e->prev_relo_r = Eigen::Matrix3d::Identity();
  }
  MarginalizationInfo* before = e->last_marginalization_info;
  g_hook = EstHook();
  g_hook.est = e; g_hook.L = L;
  if (g_relo.on && solved->relo_pose) g_hook.inj_relo = solved->relo_pose;
  g_hook.inj_pose = solved->para_pose; g_hook.inj_sb = solved->para_speed_bias; g_hook.inj_ex = solved->para_ex_pose;
  g_hook.inj_feat = solved->inv_depth; g_hook.inj_td = o->estimate_td ? solved->para_td : nullptr;
  g_hook.entry_pose = entry_pose; g_hook.entry_sb = entry_sb; g_hook.entry_ex = entry_ex; g_hook.entry_feat = entry_feat;
  if (ref_on_estimator_created) ref_on_estimator_created(e);
  else ceres::solve_hook() = est_solve_hook;
  e->optimization();
  ceres::solve_hook() = nullptr;
  if (g_relo.on) {
    memcpy(g_relo.out_pose, e->relo_Pose, sizeof g_relo.out_pose);
    for (int a = 0; a < 3; ++a) g_relo.rel_t[a] = e->relo_relative_t(a);
    g_relo.rel_yaw = e->relo_relative_yaw;
    g_relo.on = false;
  }
  scal[0] = g_hook.entry_cost; scal[1] = g_hook.max_time;
  for (int i = 0; i < 8; ++i) counts[i] = g_hook.counts[i];
  for (int i = 0; i < 4; ++i) options[i] = g_hook.options[i];
  for (int i = 0; i < K; ++i) {
    for (int a = 0; a < 3; ++a) { post_P[3 * i + a] = e->Ps[i](a); post_V[3 * i + a] = e->Vs[i](a); post_Ba[3 * i + a] = e->Bas[i](a); post_Bg[3 * i + a] = e->Bgs[i](a); }
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) post_R[9 * i + 3 * a + b] = e->Rs[i](a, b);
  }
  for (int a = 0; a < 3; ++a) post_ex[a] = e->tic[0](a);
  for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) post_ex[3 + 3 * a + b] = e->ric[0](a, b);
  post_td[0] = e->td;
  { int l = 0; for (auto& f : e->f_manager.feature) post_depth[l++] = f.estimated_depth; }
  // the new prior
  MarginalizationInfo* info = e->last_marginalization_info;
  if (info == before || info == nullptr) { prior->n = -1; prior->nblocks = 0; return 0; }
  int n = info->n, nb = (int)e->last_marginalization_parameter_blocks.size();
  if (n > prior->cap_n || nb > prior->cap_blocks) return -4;
  prior->n = n; prior->nblocks = nb;
  int off = 0;
  for (int b = 0; b < nb; ++b) {
    double* p = e->last_marginalization_parameter_blocks[b];
    int kind = -1, frame = 0;
    for (int i = 0; i < K; ++i) { if (p == e->para_Pose[i]) { kind = BVIO_BLK_POSE; frame = i; } if (p == e->para_SpeedBias[i]) { kind = BVIO_BLK_SPEEDBIAS; frame = i; } }
    if (p == e->para_Ex_Pose[0]) kind = BVIO_BLK_EXPOSE;
    if (p == e->para_Td[0]) kind = BVIO_BLK_TD;
    if (kind < 0) return -5;
    prior->block_kind[b] = kind; prior->block_frame[b] = frame; prior->block_idx[b] = info->keep_block_idx[b] - info->m;
    int gs = info->keep_block_size[b];
    for (int i = 0; i < gs; ++i) prior->x0[off + i] = info->keep_block_data[b][i];
    off += gs;
  }
  for (int i = 0; i < n; ++i) { prior->lin_res[i] = info->linearized_residuals(i); for (int j = 0; j < n; ++j) prior->lin_jac[(size_t)j * n + i] = info->linearized_jacobians(i, j); }
  return 0;    // the Estimator and its MarginalizationInfo objects are leaked on purpose (see ref_marginalize)
}

}  // extern "C"


// ---- Estimator::slideWindow() (estimator.cpp:996-1107) --------------------------------------------------------------------------
extern "C" {

// Builds an Estimator holding a full window and calls slideWindow().  Interval j (1..WINDOW_SIZE) links frame j-1 to j:
// imu_n[j] samples (dt, acc, gyr) after its start sample imu_start[j] = (acc0, gyr0), linearization biases imu_lin[j] =
// (ba, bg) as they were when the interval was opened; all intervals concatenated in
// imu_dt / imu_acc / imu_gyr.  Features as CSR like bvio_window plus a depth per feature (<= 0: not triangulated).
// Outputs: the shifted states, the packed preintegration of every interval that survives (index j-1 for MARGIN_OLD,
// merged into WINDOW_SIZE-1 for MARGIN_SECOND_NEW), and the FeatureManager dump.
int ref_estimator_slide(int flag, const double* poses, const double* sb, const double* ex, const int32_t* imu_n,
                        const double* imu_start, const double* imu_lin, const double* imu_dt, const double* imu_acc, const double* imu_gyr,
                        double acc_n, double gyr_n, double acc_w, double gyr_w, int n_feat, const int32_t* f_id,
                        const int32_t* f_off, const int32_t* f_start, const double* f_xy, const double* f_depth,
                        double* out_poses, double* out_sb, bvio_preint* out_pre, int32_t* out_sum,
                        int cap, int32_t* d_id, int32_t* d_start, int32_t* d_nobs, double* d_depth) {
  const int K = WINDOW_SIZE + 1;
  ACC_N = acc_n; GYR_N = gyr_n; ACC_W = acc_w; GYR_W = gyr_w; TD = 0;
  void* mem = calloc(1, sizeof(Estimator));
  Estimator* e = new (mem) Estimator();
  e->tic[0] = v3(ex);
  e->ric[0] = Eigen::Quaterniond(ex[6], ex[3], ex[4], ex[5]).toRotationMatrix();
  e->f_manager.setRic(e->ric);
  int off = 0;
  for (int i = 0; i < K; ++i) {
    const double* p = poses + 7 * i; const double* s = sb + 9 * i;
    e->Ps[i] = v3(p); e->Rs[i] = Eigen::Quaterniond(p[6], p[3], p[4], p[5]).toRotationMatrix();
    e->Vs[i] = v3(s); e->Bas[i] = v3(s + 3); e->Bgs[i] = v3(s + 6);
    e->Headers[i].stamp = ros::Time(100.0 + 0.1 * i);
    ImageFrame fr(map<int, vector<pair<int, Eigen::Matrix<double, 7, 1>>>>(), e->Headers[i].stamp.toSec());
    fr.pre_integration = new IntegrationBase(Eigen::Vector3d::Zero(), Eigen::Vector3d::Zero(), Eigen::Vector3d::Zero(), Eigen::Vector3d::Zero());
    e->all_image_frame.insert(make_pair(e->Headers[i].stamp.toSec(), fr));
    if (i >= 1) {
      e->pre_integrations[i] = new IntegrationBase(v3(imu_start + 6 * i), v3(imu_start + 6 * i + 3), v3(imu_lin + 6 * i), v3(imu_lin + 6 * i + 3));
      for (int k = 0; k < imu_n[i]; ++k, ++off) {
        e->pre_integrations[i]->push_back(imu_dt[off], v3(imu_acc + 3 * off), v3(imu_gyr + 3 * off));
        e->dt_buf[i].push_back(imu_dt[off]);
        e->linear_acceleration_buf[i].push_back(v3(imu_acc + 3 * off));
        e->angular_velocity_buf[i].push_back(v3(imu_gyr + 3 * off));
        e->acc_0 = v3(imu_acc + 3 * off); e->gyr_0 = v3(imu_gyr + 3 * off);
      }
    }
  }
  for (int l = 0; l < n_feat; ++l) {
    FeaturePerId f(f_id[l], f_start[l]);
    for (int k = f_off[l]; k < f_off[l + 1]; ++k) {
      Eigen::Matrix<double, 7, 1> pt; pt.setZero(); pt(0) = f_xy[2 * k]; pt(1) = f_xy[2 * k + 1]; pt(2) = 1.0;
      f.feature_per_frame.push_back(FeaturePerFrame(pt, 0.0));
    }
    f.estimated_depth = f_depth[l];
    e->f_manager.feature.push_back(f);
  }
  e->frame_count = WINDOW_SIZE; e->solver_flag = Estimator::NON_LINEAR;
  e->marginalization_flag = flag == 0 ? Estimator::MARGIN_OLD : Estimator::MARGIN_SECOND_NEW;
  e->slideWindow();
  for (int i = 0; i < K; ++i) {
    Eigen::Quaterniond q(e->Rs[i]);
    for (int a = 0; a < 3; ++a) { out_poses[7 * i + a] = e->Ps[i](a); out_sb[9 * i + a] = e->Vs[i](a); out_sb[9 * i + 3 + a] = e->Bas[i](a); out_sb[9 * i + 6 + a] = e->Bgs[i](a); }
    out_poses[7 * i + 3] = q.x(); out_poses[7 * i + 4] = q.y(); out_poses[7 * i + 5] = q.z(); out_poses[7 * i + 6] = q.w();
    memset(&out_pre[i], 0, sizeof(bvio_preint));
    if (i >= 1 && e->pre_integrations[i]) store(*e->pre_integrations[i], &out_pre[i]);
  }
  out_sum[0] = e->sum_of_back; out_sum[1] = e->sum_of_front;
  int k = 0;
  for (auto& f : e->f_manager.feature) {
    if (k < cap) { d_id[k] = f.feature_id; d_start[k] = f.start_frame; d_nobs[k] = (int)f.feature_per_frame.size(); d_depth[k] = f.estimated_depth; }
    ++k;
  }
  return k;
}

}  // extern "C"


// ---- Estimator::processIMU() (estimator.cpp:86-119): the state prediction of the incoming frame --------------------------------
extern "C" {
// State of the newest frame (pose7, sb9) + the start sample (acc_0, gyr_0) + n samples (dt, acc, gyr) -> predicted state
// after feeding the samples through processIMU with frame_count = 1; out_R is Rs[j] row-major (the reference never
// re-orthonormalises it inside a frame), out_pre the preintegration it accumulated.
void ref_estimator_process_imu(const double* pose, const double* sb, const double* start, int n, const double* dt, const double* acc,
                               const double* gyr, const double* G_, double acc_n, double gyr_n, double acc_w, double gyr_w,
                               double* out_P, double* out_R, double* out_V, bvio_preint* out_pre) {
  ACC_N = acc_n; GYR_N = gyr_n; ACC_W = acc_w; GYR_W = gyr_w; TD = 0;
  void* mem = calloc(1, sizeof(Estimator));
  Estimator* e = new (mem) Estimator();
  e->g = v3(G_);
  e->frame_count = 1;
  e->Ps[1] = v3(pose); e->Rs[1] = Eigen::Quaterniond(pose[6], pose[3], pose[4], pose[5]).toRotationMatrix();
  e->Vs[1] = v3(sb); e->Bas[1] = v3(sb + 3); e->Bgs[1] = v3(sb + 6);
  e->first_imu = true; e->acc_0 = v3(start); e->gyr_0 = v3(start + 3);
  e->tmp_pre_integration = new IntegrationBase(e->acc_0, e->gyr_0, e->Bas[1], e->Bgs[1]);
  for (int k = 0; k < n; ++k) e->processIMU(dt[k], v3(acc + 3 * k), v3(gyr + 3 * k));
  for (int a = 0; a < 3; ++a) { out_P[a] = e->Ps[1](a); out_V[a] = e->Vs[1](a); for (int b = 0; b < 3; ++b) out_R[3 * a + b] = e->Rs[1](a, b); }
  store(*e->pre_integrations[1], out_pre);
}
}  // extern "C"


// ---- a live Estimator fed frame by frame: processIMU / processImage (estimator.cpp:86-186) with the external solve ---------------
extern "C" {
void* ref_est_create(const double* ex, const double* G_, double focal, int max_iters, double acc_n, double gyr_n, double acc_w,
                     double gyr_w, double init_depth, double min_parallax) {
  ACC_N = acc_n; GYR_N = gyr_n; ACC_W = acc_w; GYR_W = gyr_w; TD = 0; TR = 0; ROW = 480;
  INIT_DEPTH = init_depth; MIN_PARALLAX = min_parallax;
  ESTIMATE_EXTRINSIC = 0; ESTIMATE_TD = 0; NUM_ITERATIONS = max_iters; SOLVER_TIME = 0.04;
  G = v3(G_);
  RIC.assign(1, Eigen::Quaterniond(ex[6], ex[3], ex[4], ex[5]).toRotationMatrix());
  TIC.assign(1, v3(ex));
  void* mem = calloc(1, sizeof(Estimator));
  Estimator* e = new (mem) Estimator();
  e->setParameter();
  e->g = G;
  ProjectionFactor::sqrt_info = focal / 1.5 * Eigen::Matrix2d::Identity();
  g_hook = EstHook();
  g_hook.est = e; g_hook.session = true;
  ceres::solve_hook() = est_solve_hook;
  return e;
}
void ref_est_release(void*) { g_hook = EstHook(); ceres::solve_hook() = nullptr; }     // the Estimator itself is leaked (see above)
void ref_est_set_state(void* h, int i, const double* pose, const double* sb) {
  Estimator* e = static_cast<Estimator*>(h);
  e->Ps[i] = v3(pose); e->Rs[i] = Eigen::Quaterniond(pose[6], pose[3], pose[4], pose[5]).toRotationMatrix();
  e->Vs[i] = v3(sb); e->Bas[i] = v3(sb + 3); e->Bgs[i] = v3(sb + 6);
}
void ref_est_set_bias(void* h, int i, const double* ba, const double* bg) {
  Estimator* e = static_cast<Estimator*>(h);
  e->Bas[i] = v3(ba); e->Bgs[i] = v3(bg);
}
void ref_est_get_states(void* h, double* poses, double* sb) {
  Estimator* e = static_cast<Estimator*>(h);
  for (int i = 0; i <= WINDOW_SIZE; ++i) {
    Eigen::Quaterniond q(e->Rs[i]);
    for (int a = 0; a < 3; ++a) { poses[7 * i + a] = e->Ps[i](a); sb[9 * i + a] = e->Vs[i](a); sb[9 * i + 3 + a] = e->Bas[i](a); sb[9 * i + 6 + a] = e->Bgs[i](a); }
    poses[7 * i + 3] = q.x(); poses[7 * i + 4] = q.y(); poses[7 * i + 5] = q.z(); poses[7 * i + 6] = q.w();
  }
}
void ref_est_set_nonlinear(void* h) { static_cast<Estimator*>(h)->solver_flag = Estimator::NON_LINEAR; }
int ref_est_frame_count(void* h) { return static_cast<Estimator*>(h)->frame_count; }
int ref_est_prior_size(void* h) { Estimator* e = static_cast<Estimator*>(h); return e->last_marginalization_info ? e->last_marginalization_info->n : -1; }
void ref_est_process_imu(void* h, int n, const double* dt, const double* acc, const double* gyr) {
  Estimator* e = static_cast<Estimator*>(h);
  for (int k = 0; k < n; ++k) e->processIMU(dt[k], v3(acc + 3 * k), v3(gyr + 3 * k));
}
// -> marginalization_flag the reference decided for this frame (0 MARGIN_OLD, 1 MARGIN_SECOND_NEW)
int ref_est_process_image(void* h, double stamp, int n, const int32_t* ids, const double* xy) {
  Estimator* e = static_cast<Estimator*>(h);
  map<int, vector<pair<int, Eigen::Matrix<double, 7, 1>>>> image;
  for (int i = 0; i < n; ++i) {
    Eigen::Matrix<double, 7, 1> v; v.setZero(); v(0) = xy[2 * i]; v(1) = xy[2 * i + 1]; v(2) = 1.0;
    image[ids[i]].emplace_back(0, v);
  }
  std_msgs::Header header; header.stamp = ros::Time(stamp);
  e->processImage(image, header);
  return (int)e->marginalization_flag;
}
int ref_est_dump_features(void* h, int cap, int32_t* ids, int32_t* start, int32_t* nobs, double* depth) {
  int k = 0;
  for (auto& f : static_cast<Estimator*>(h)->f_manager.feature) {
    if (k < cap) { ids[k] = f.feature_id; start[k] = f.start_frame; nobs[k] = (int)f.feature_per_frame.size(); depth[k] = f.estimated_depth; }
    ++k;
  }
  return k;
}
}  // extern "C"
