// bvio_adapter.h -- reference-side adapter: the code a maintainer of plusk01/Anticipated-VINS-Mono adds under
// vins_estimator/src/ to put libbvio.so (include/bvio.h) behind Estimator::optimization() and FeatureSelector::select().
// It is written against the reference's own classes (Estimator, FeatureManager, IntegrationBase, MarginalizationInfo,
// FeatureSelector's types) and compiles both in the reference's catkin build (real Eigen / Ceres / camodocal) and here
// against the stand-in headers of oracle/ref_shim (define BVIO_ADAPTER_SHIM), where tests drive the reference's unmodified
// Estimator / FeatureSelector through it on a B200 (adapters/vins/interpose.cpp, tests/test_zz_gpu_adapter.py).
//
// Call sites it replaces (all in the reference tree, vins_estimator/src):
//   solve()        ceres::Solve(options, &problem, &summary)                       estimator.cpp:809
//                  for the problem built at estimator.cpp:663-792 (vector2double() at :692 has already run; the solution
//                  lands in the para_* arrays exactly where Ceres leaves it, so double2vector() at :814 follows unchanged)
//   marginalize()  marginalization_info->preMarginalize(); ->marginalize();        estimator.cpp:897-901, :952-957
//                  for the residual blocks added at :822-893 (MARGIN_OLD) / :933-947 (MARGIN_SECOND_NEW); fills the
//                  MarginalizationInfo so that getParameterBlocks(addr_shift) (:903-914) and the next frame's
//                  MarginalizationFactor work unchanged
//   select()       calcInfoFromRobotMotion + addOmegaPrior + 2 x calcInfoFromFeatures + selectInformativeFeatures
//                                                                                  feature_selector.cpp:139-171
#pragma once
#include <vector>

#include "estimator.h"
#include "feature_selector.h"
#include "bvio.h"

namespace bvio_adapter {

// Owns every array a bvio_window points into, except the state: para_pose / para_speed_bias / para_ex_pose / para_td /
// inv_depth alias the Estimator's own Ceres parameter arrays (estimator.h:109-115), so the solve is in place.
struct Window {
  std::vector<int32_t> lm_off, obs_frame;
  std::vector<double> obs_xy, obs_vel, obs_td, obs_row;
  std::vector<bvio_preint> preint;
  std::vector<int32_t> prior_kind, prior_frame, prior_idx;
  std::vector<double> prior_x0, prior_jac, prior_res;
  std::vector<int32_t> relo_lm;
  std::vector<double> relo_xy;
  bvio_prior prior;
  bvio_window w;
};

// IntegrationBase -> bvio_preint (the members IMUFactor::Evaluate reads, integration_base.h:195-203)
void pack_preint(const IntegrationBase& pre, bvio_preint* out);

// last_marginalization_info + last_marginalization_parameter_blocks -> bvio_prior.  false: there is no prior.
bool pack_prior(const Estimator& e, Window* win);

// Landmarks and observations in the order optimization() walks them (estimator.cpp:710-755): CSR arrays of bvio_window
// (+ velocity / td / row when ESTIMATE_TD) and the relocalization matches (estimator.cpp:760-792).  Returns the number of
// landmarks = feature_index + 1 = f_manager.getFeatureCount().
int fill_structure(Estimator& e, Window* win);

// Everything: state pointers into the Estimator, structure, preintegrations, prior.
void fill_window(Estimator& e, Window* win);

// The globals optimization() reads (parameters.cpp) and the Solver::Options it sets (estimator.cpp:794-806).
bvio_opts make_opts(const Estimator& e);

// In place of ceres::Solve at estimator.cpp:809.  Returns a bvio_status.
int solve(bvio_ctx* ctx, Estimator& e, bvio_summary* summary);

// In place of preMarginalize() + marginalize() at estimator.cpp:897-901 (flag 0, MARGIN_OLD) and :952-957 (flag 1,
// MARGIN_SECOND_NEW): `info` has received its addResidualBlockInfo() calls; afterwards it holds m, n, parameter_block_idx,
// parameter_block_data, linearized_jacobians, linearized_residuals like the reference's own procedure leaves them (block
// order: frame-major instead of the reference's address-hash order; the quadratic form is the same).
int marginalize(bvio_ctx* ctx, Estimator& e, MarginalizationInfo* info);

// bvio_prior_out -> MarginalizationInfo (used by marginalize(); exposed for tests)
void install_prior(Estimator& e, MarginalizationInfo* info, const bvio_prior_out& out, int flag);

// Owns the arrays of a bvio_select_in.
struct SelectInputs {
  std::vector<double> horizon_pos, horizon_quat, cand_xy, cand_prob, used_xy, cloud_xy, cloud_depth;
  std::vector<int32_t> cand_id, used_id;
  double k1_pos[3], k1_quat[4];
  bvio_select_in in;
};

// Inputs of the numerical part of select() (feature_selector.cpp:139-171): the horizon of generateFutureHorizon(),
// the IMU-propagated state_k1_ (back-projection of the features, :247-266), extrinsics, camera, IMU parameters, the new
// features (candidates) and the tracked subset, the depth cloud of initKDTree() (:380-421) and kappa (:162).
void fill_select_in(const Estimator& e, const state_horizon_t& state_kkH, const state_t& state_k1,
                    const Eigen::Quaterniond& q_IC, const Eigen::Vector3d& t_IC, const bvio_camera& cam, int nr_imu,
                    double delta_imu, double acc_var, double acc_bias_var, const image_t& image_new, const image_t& subset,
                    int kappa, SelectInputs* out);

// In place of selectInformativeFeatures (feature_selector.cpp:170-171): ids in selection order.
int select(bvio_ctx* ctx, const SelectInputs& in, std::vector<int>* selected, bvio_select_summary* summary);

}  // namespace bvio_adapter
