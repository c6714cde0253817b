// bvio_adapter.cpp -- see bvio_adapter.h.  Everything here is bookkeeping over the reference's own data structures: no
// arithmetic of the hot path runs on the host (libbvio.so has no CPU path).
#include "bvio_adapter.h"

#include <algorithm>
#include <cstring>
#include <unordered_map>

namespace bvio_adapter {

namespace {
const int K = WINDOW_SIZE + 1;

// which Ceres parameter block of the Estimator an address is (estimator.h:109-115)
bool identify(const Estimator& e, const double* p, int* kind, int* frame) {
  for (int i = 0; i < K; ++i) {
    if (p == e.para_Pose[i]) { *kind = BVIO_BLK_POSE; *frame = i; return true; }
    if (p == e.para_SpeedBias[i]) { *kind = BVIO_BLK_SPEEDBIAS; *frame = i; return true; }
  }
  if (p == e.para_Ex_Pose[0]) { *kind = BVIO_BLK_EXPOSE; *frame = 0; return true; }
  if (p == e.para_Td[0]) { *kind = BVIO_BLK_TD; *frame = 0; return true; }
  return false;
}
double* block_addr(Estimator& e, int kind, int frame) {
  return kind == BVIO_BLK_POSE ? e.para_Pose[frame] : kind == BVIO_BLK_SPEEDBIAS ? e.para_SpeedBias[frame]
         : kind == BVIO_BLK_EXPOSE ? e.para_Ex_Pose[0] : e.para_Td[0];
}
int global_size(int kind) { return kind == BVIO_BLK_SPEEDBIAS ? 9 : kind == BVIO_BLK_TD ? 1 : 7; }
int local_size(int kind) { return kind == BVIO_BLK_SPEEDBIAS ? 9 : kind == BVIO_BLK_TD ? 1 : 6; }
}  // namespace

void pack_preint(const IntegrationBase& pre, bvio_preint* out) {
  for (int i = 0; i < 3; ++i) {
    out->delta_p[i] = pre.delta_p(i); out->delta_v[i] = pre.delta_v(i);
    out->lin_ba[i] = pre.linearized_ba(i); out->lin_bg[i] = pre.linearized_bg(i);
  }
  out->delta_q[0] = pre.delta_q.x(); out->delta_q[1] = pre.delta_q.y(); out->delta_q[2] = pre.delta_q.z(); out->delta_q[3] = pre.delta_q.w();
  out->sum_dt = pre.sum_dt;
  for (int i = 0; i < 15; ++i)
    for (int j = 0; j < 15; ++j) {
      out->jacobian[i * 15 + j] = pre.jacobian(i, j);
      out->covariance[i * 15 + j] = pre.covariance(i, j);
    }
}

bool pack_prior(const Estimator& e, Window* win) {
  const MarginalizationInfo* info = e.last_marginalization_info;
  if (!info) return false;
  const std::vector<double*>& blocks = e.last_marginalization_parameter_blocks;
  const int nb = (int)blocks.size(), n = info->n;
  win->prior_kind.resize(nb); win->prior_frame.resize(nb); win->prior_idx.resize(nb);
  win->prior_x0.clear();
  for (int b = 0; b < nb; ++b) {
    int kind = 0, frame = 0;
    identify(e, blocks[b], &kind, &frame);
    win->prior_kind[b] = kind; win->prior_frame[b] = frame;
    win->prior_idx[b] = info->keep_block_idx[b] - info->m;            // MarginalizationFactor::Evaluate, marginalization_factor.cpp:349
    const int gs = info->keep_block_size[b];
    win->prior_x0.insert(win->prior_x0.end(), info->keep_block_data[b], info->keep_block_data[b] + gs);
  }
  win->prior_jac.resize((size_t)n * n); win->prior_res.resize(n);
  for (int i = 0; i < n; ++i) {
    win->prior_res[i] = info->linearized_residuals(i);
    for (int j = 0; j < n; ++j) win->prior_jac[(size_t)j * n + i] = info->linearized_jacobians(i, j);   // column-major
  }
  bvio_prior& p = win->prior;
  p.n = n; p.nblocks = nb;
  p.block_kind = win->prior_kind.data(); p.block_frame = win->prior_frame.data(); p.block_idx = win->prior_idx.data();
  p.x0 = win->prior_x0.data(); p.lin_jac = win->prior_jac.data(); p.lin_res = win->prior_res.data();
  return true;
}

int fill_structure(Estimator& e, Window* win) {
  win->lm_off.assign(1, 0);
  win->obs_frame.clear(); win->obs_xy.clear(); win->obs_vel.clear(); win->obs_td.clear(); win->obs_row.clear();
  win->relo_lm.clear(); win->relo_xy.clear();
  int feature_index = -1, retrive_feature_index = 0;
  for (auto& it_per_id : e.f_manager.feature) {
    it_per_id.used_num = it_per_id.feature_per_frame.size();
    if (!(it_per_id.used_num >= 2 && it_per_id.start_frame < WINDOW_SIZE - 2)) continue;   // estimator.cpp:712-718
    ++feature_index;
    int frame = it_per_id.start_frame;
    for (auto& it_per_frame : it_per_id.feature_per_frame) {
      win->obs_frame.push_back(frame++);
      win->obs_xy.push_back(it_per_frame.point.x()); win->obs_xy.push_back(it_per_frame.point.y());
      if (ESTIMATE_TD) {                                                                    // estimator.cpp:732-740
        win->obs_vel.push_back(it_per_frame.velocity.x()); win->obs_vel.push_back(it_per_frame.velocity.y());
        win->obs_td.push_back(it_per_frame.cur_td);
        win->obs_row.push_back(it_per_frame.uv.y());
      }
    }
    win->lm_off.push_back((int32_t)win->obs_frame.size());
    // relocalization matches, walked exactly like estimator.cpp:768-790
    if (e.relocalization_info && it_per_id.start_frame <= e.relo_frame_local_index) {
      while (retrive_feature_index < (int)e.match_points.size() && (int)e.match_points[retrive_feature_index].z() < it_per_id.feature_id)
        retrive_feature_index++;
      if (retrive_feature_index < (int)e.match_points.size() && (int)e.match_points[retrive_feature_index].z() == it_per_id.feature_id) {
        win->relo_lm.push_back(feature_index);
        win->relo_xy.push_back(e.match_points[retrive_feature_index].x());
        win->relo_xy.push_back(e.match_points[retrive_feature_index].y());
        retrive_feature_index++;
      }
    }
  }
  return feature_index + 1;
}

void fill_window(Estimator& e, Window* win) {
  const int L = fill_structure(e, win);
  win->preint.resize(K);
  std::memset(win->preint.data(), 0, sizeof(bvio_preint) * K);
  for (int j = 1; j < K; ++j) pack_preint(*e.pre_integrations[j], &win->preint[j]);
  bvio_window& w = win->w;
  std::memset(&w, 0, sizeof w);
  w.K = K;
  w.para_pose = &e.para_Pose[0][0]; w.para_speed_bias = &e.para_SpeedBias[0][0];
  w.para_ex_pose = e.para_Ex_Pose[0]; w.para_td = e.para_Td[0];
  w.L = L; w.inv_depth = &e.para_Feature[0][0];
  w.lm_obs_offset = win->lm_off.data(); w.obs_frame = win->obs_frame.data(); w.obs_xy = win->obs_xy.data();
  if (ESTIMATE_TD) { w.obs_vel = win->obs_vel.data(); w.obs_td = win->obs_td.data(); w.obs_row = win->obs_row.data(); }
  w.preint = win->preint.data();
  w.prior = pack_prior(e, win) ? &win->prior : nullptr;
  w.n_relo = (int32_t)win->relo_lm.size();
  if (w.n_relo > 0) { w.relo_pose = e.relo_Pose; w.relo_lm = win->relo_lm.data(); w.relo_xy = win->relo_xy.data(); }
}

bvio_opts make_opts(const Estimator& e) {
  bvio_opts o;
  bvio_default_opts(&o);
  o.max_iters = NUM_ITERATIONS;                                            // estimator.cpp:799
  // estimator.cpp:803-806
  o.max_time_s = e.marginalization_flag == Estimator::MARGIN_OLD ? SOLVER_TIME * 4.0 / 5.0 : SOLVER_TIME;
  o.estimate_extrinsic = ESTIMATE_EXTRINSIC ? 1 : 0;                       // estimator.cpp:677-683
  o.estimate_td = ESTIMATE_TD ? 1 : 0;                                     // estimator.cpp:685-689
  o.focal_length = FOCAL_LENGTH;                                           // ProjectionFactor::sqrt_info, estimator.cpp:17
  o.cauchy_a = 1.0;                                                        // estimator.cpp:666
  for (int i = 0; i < 3; ++i) o.G[i] = G(i);
  o.TR = TR; o.ROW = ROW;
  o.strategy = BVIO_STRATEGY_DOGLEG;                                       // estimator.cpp:798
  return o;
}

int solve(bvio_ctx* ctx, Estimator& e, bvio_summary* summary) {
  Window win;
  fill_window(e, &win);
  bvio_opts o = make_opts(e);
  return bvio_optimize(ctx, &win.w, &o, summary);
}

void install_prior(Estimator& e, MarginalizationInfo* info, const bvio_prior_out& out, int flag) {
  // dropped blocks: everything addResidualBlockInfo() registered with index 0 (marginalization_factor.cpp:101-105);
  // m = their total local size, in any order (nothing reads the dropped columns afterwards)
  int m = 0;
  for (auto& it : info->parameter_block_idx) { it.second = m; m += info->localSize(info->parameter_block_size[it.first]); }
  info->m = m;
  info->n = out.n;
  for (int b = 0; b < out.nblocks; ++b) {
    // bvio_marginalize reports frames after the window shift; the MarginalizationInfo is keyed by the present addresses
    int frame = out.block_frame[b];
    if (out.block_kind[b] == BVIO_BLK_POSE || out.block_kind[b] == BVIO_BLK_SPEEDBIAS)
      frame = flag == 0 ? frame + 1 : (frame == K - 2 ? K - 1 : frame);
    const long addr = reinterpret_cast<long>(block_addr(e, out.block_kind[b], frame));
    info->parameter_block_idx[addr] = m + out.block_idx[b];
  }
  info->linearized_jacobians.resize(out.n, out.n);
  info->linearized_residuals.resize(out.n);
  for (int i = 0; i < out.n; ++i) {
    info->linearized_residuals(i) = out.lin_res[i];
    for (int j = 0; j < out.n; ++j) info->linearized_jacobians(i, j) = out.lin_jac[(size_t)j * out.n + i];
  }
}

int marginalize(bvio_ctx* ctx, Estimator& e, MarginalizationInfo* info) {
  // preMarginalize()'s bookkeeping half (marginalization_factor.cpp:117-128): freeze the linearization point of every
  // block the factors touch.  Its numerical half (it->Evaluate()) runs on the device.
  for (auto it : info->factors) {
    it->raw_jacobians = nullptr;                      // never evaluated on the host; ~MarginalizationInfo delete[]s it
    std::vector<int> block_sizes = it->cost_function->parameter_block_sizes();
    for (int i = 0; i < (int)block_sizes.size(); ++i) {
      const long addr = reinterpret_cast<long>(it->parameter_blocks[i]);
      if (info->parameter_block_data.find(addr) == info->parameter_block_data.end()) {
        double* data = new double[block_sizes[i]];
        std::memcpy(data, it->parameter_blocks[i], sizeof(double) * block_sizes[i]);
        info->parameter_block_data[addr] = data;
      }
    }
  }
  // MARGIN_OLD drops Pose[0] (estimator.cpp:822-893), MARGIN_SECOND_NEW drops Pose[WINDOW_SIZE-1] (:933-947)
  const int flag = info->parameter_block_idx.count(reinterpret_cast<long>(e.para_Pose[0])) ? 0 : 1;
  Window win;
  fill_window(e, &win);
  bvio_opts o = make_opts(e);
  const int cap_n = 15 * K + 16, cap_b = 2 * K + 2;
  std::vector<int32_t> bk(cap_b), bf(cap_b), bi(cap_b);
  std::vector<double> x0(9 * cap_b), jac((size_t)cap_n * cap_n), res(cap_n);
  bvio_prior_out out;
  out.block_kind = bk.data(); out.block_frame = bf.data(); out.block_idx = bi.data();
  out.x0 = x0.data(); out.lin_jac = jac.data(); out.lin_res = res.data();
  out.cap_n = cap_n; out.cap_blocks = cap_b; out.n = 0; out.nblocks = 0;
  int rc = bvio_marginalize(ctx, &win.w, &o, flag, &out);
  if (rc != BVIO_OK) return rc;
  if (out.n < 0) return BVIO_ERR_INVALID;             // the caller checks for Pose[WINDOW_SIZE-1] before it gets here (:926-927)
  install_prior(e, info, out, flag);
  return BVIO_OK;
}

void fill_select_in(const Estimator& e, const state_horizon_t& state_kkH, const state_t& state_k1,
                    const Eigen::Quaterniond& q_IC, const Eigen::Vector3d& t_IC, const bvio_camera& cam, int nr_imu,
                    double delta_imu, double acc_var, double acc_bias_var, const image_t& image_new, const image_t& subset,
                    int kappa, SelectInputs* out) {
  bvio_select_in& in = out->in;
  std::memset(&in, 0, sizeof in);
  in.H = HORIZON;
  out->horizon_pos.resize(3 * (HORIZON + 1)); out->horizon_quat.resize(4 * (HORIZON + 1));
  for (int h = 0; h <= HORIZON; ++h) {
    for (int a = 0; a < 3; ++a) out->horizon_pos[3 * h + a] = state_kkH[h].first(xPOS + a);
    const Eigen::Quaterniond& q = state_kkH[h].second;
    out->horizon_quat[4 * h] = q.x(); out->horizon_quat[4 * h + 1] = q.y(); out->horizon_quat[4 * h + 2] = q.z(); out->horizon_quat[4 * h + 3] = q.w();
  }
  in.horizon_pos = out->horizon_pos.data(); in.horizon_quat = out->horizon_quat.data();
  // state_k1_: equals state_kkH[1] in IMU-horizon mode, differs in ground-truth mode (horizon_generator.cpp:106-117)
  for (int a = 0; a < 3; ++a) out->k1_pos[a] = state_k1.first(xPOS + a);
  out->k1_quat[0] = state_k1.second.x(); out->k1_quat[1] = state_k1.second.y(); out->k1_quat[2] = state_k1.second.z(); out->k1_quat[3] = state_k1.second.w();
  in.state_k1_pos = out->k1_pos; in.state_k1_quat = out->k1_quat;
  in.q_ic[0] = q_IC.x(); in.q_ic[1] = q_IC.y(); in.q_ic[2] = q_IC.z(); in.q_ic[3] = q_IC.w();
  for (int a = 0; a < 3; ++a) in.t_ic[a] = t_IC(a);
  in.cam = cam;
  in.nr_imu = nr_imu; in.delta_imu = delta_imu; in.acc_var = acc_var; in.acc_bias_var = acc_bias_var;
  // candidates = image_new, tracked = subset; std::map order = ascending feature id
  out->cand_id.clear(); out->cand_xy.clear(); out->cand_prob.clear(); out->used_id.clear(); out->used_xy.clear();
  for (const auto& f : image_new) {
    out->cand_id.push_back(f.first);
    out->cand_xy.push_back(f.second[0].second(0)); out->cand_xy.push_back(f.second[0].second(1));
    out->cand_prob.push_back(f.second[0].second(fPROB));
  }
  for (const auto& f : subset) {
    out->used_id.push_back(f.first);
    out->used_xy.push_back(f.second[0].second(0)); out->used_xy.push_back(f.second[0].second(1));
  }
  in.N = (int32_t)out->cand_id.size(); in.cand_id = out->cand_id.data(); in.cand_xy = out->cand_xy.data(); in.cand_prob = out->cand_prob.data();
  in.U = (int32_t)out->used_id.size(); in.used_id = out->used_id.data(); in.used_xy = out->used_xy.data();
  // depth cloud: initKDTree()'s dataset (feature_selector.cpp:396-421)
  out->cloud_xy.clear(); out->cloud_depth.clear();
  const Eigen::Quaterniond q1_inv = state_k1.second.inverse(), qic_inv = q_IC.inverse();
  const Eigen::Vector3d P1(state_k1.first(xPOS), state_k1.first(xPOS + 1), state_k1.first(xPOS + 2));
  for (const auto& it_per_id : e.f_manager.feature) {
    const int used_num = it_per_id.feature_per_frame.size();
    if (!(used_num >= 2 && it_per_id.start_frame < WINDOW_SIZE - 2)) continue;
    if (it_per_id.start_frame > WINDOW_SIZE * 3.0 / 4.0 || it_per_id.solve_flag != 1) continue;
    const int imu_i = it_per_id.start_frame;
    Eigen::Vector3d pts_i = it_per_id.feature_per_frame[0].point * it_per_id.estimated_depth;
    Eigen::Vector3d w_pts_i = e.Rs[imu_i] * (e.ric[0] * pts_i + e.tic[0]) + e.Ps[imu_i];
    Eigen::Vector3d p_IL_k1 = q1_inv * (w_pts_i - P1);
    Eigen::Vector3d p_CL_k1 = qic_inv * (p_IL_k1 - t_IC);
    out->cloud_xy.push_back(p_CL_k1(0) / p_CL_k1(2)); out->cloud_xy.push_back(p_CL_k1(1) / p_CL_k1(2));
    out->cloud_depth.push_back(it_per_id.estimated_depth);
  }
  in.C = (int32_t)out->cloud_depth.size(); in.cloud_xy = out->cloud_xy.data(); in.cloud_depth = out->cloud_depth.data();
  in.kappa = kappa;
}

int select(bvio_ctx* ctx, const SelectInputs& in, std::vector<int>* selected, bvio_select_summary* summary) {
  std::vector<int32_t> ids(std::max(1, in.in.kappa));
  bvio_select_summary s;
  std::memset(&s, 0, sizeof s);
  int rc = bvio_select(ctx, &in.in, ids.data(), nullptr, &s);
  if (rc != BVIO_OK) return rc;
  selected->assign(ids.begin(), ids.begin() + s.n_selected);
  if (summary) *summary = s;
  return BVIO_OK;
}

}  // namespace bvio_adapter
