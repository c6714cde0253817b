// oracle/sel_oracle.cpp -- CPU restatement of FeatureSelector's numerical path.
// TEST INFRASTRUCTURE ONLY.  Pinned against the reference's own compiled sources (oracle/_ref, tests/test_reference_pin.py)
// except the Ceres trust-region loop, which is unpinned by reference code (see oracle/oracle.h).
//
// Literal restatement (dense 9(H+1) x 9(H+1) matrices, Hadamard upper bounds in a
// std::map<double,int,greater>, lazy greedy with Cholesky log-det) of
//   calcInfoFromRobotMotion / createLinearImuMatrices / addOmegaPrior
//                                   vins_estimator/src/feature_selector.cpp:463-609
//   calcInfoFromFeatures / inFOV    vins_estimator/src/feature_selector.cpp:239-376
//   findNNDepth                     vins_estimator/src/feature_selector.cpp:437-459
//   selectInformativeFeatures / sortedlogDetUB
//                                   vins_estimator/src/feature_selector.cpp:613-728
//   Utility::logdet                 vins_estimator/src/utility/utility.h:143-167
//   PinholeCamera::spaceToPlane / distortion
//                                   camera_model/src/camera_models/PinholeCamera.cc:520-542,672-688
// H is a runtime parameter here (reference: compile-time HORIZON = 13, state_defs.h:8).
#include "oracle.h"
#include "linalg.hpp"
#include <map>
#include <functional>
#include <limits>
#include <vector>

using namespace orc;

namespace {
inline V3 v3(const double* p) { return {p[0], p[1], p[2]}; }
inline Q4 q4(const double* p) { return {p[0], p[1], p[2], p[3]}; }

void linear_imu(Q4 Qi, Q4 Qj, int nr, double dImu, double accVar, double biasVar, double* omega9, double* ablk9,
                double* cov9) {
  M3 Nij = zero3(), Mij = zero3();
  double CCt_11 = 0, CCt_12 = 0;
  for (int i = 0; i < nr; ++i) {
    Q4 q = qslerp(Qi, i / static_cast<double>(nr), Qj);
    double jkh = (nr - i - 0.5);
    M3 R = qmat(q);
    Nij = Nij + jkh * R;
    Mij = Mij + R;
    CCt_11 += jkh * jkh;
    CCt_12 += jkh;
  }
  const double d2 = dImu * dImu, d3 = d2 * dImu, d4 = d3 * dImu;
  double cov[81] = {0};
  for (int i = 0; i < 3; i++) {
    cov[i * 9 + i] = 1.0 * nr * CCt_11 * d4 * accVar;
    cov[i * 9 + 3 + i] = 1.0 * CCt_12 * d3 * accVar;
    cov[(3 + i) * 9 + i] = cov[i * 9 + 3 + i];
    cov[(3 + i) * 9 + 3 + i] = 1.0 * nr * d2 * accVar;
    cov[(6 + i) * 9 + 6 + i] = 1.0 * nr * biasVar;
  }
  Nij = d2 * Nij;
  Mij = dImu * Mij;
  double A[81] = {0};
  for (int i = 0; i < 9; i++) A[i * 9 + i] = -1.0;
  for (int i = 0; i < 3; i++) A[i * 9 + 3 + i] = -1.0 * nr * dImu;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) { A[i * 9 + 6 + j] = Nij[i][j]; A[(3 + i) * 9 + 6 + j] = Mij[i][j]; }
  if (cov9) std::memcpy(cov9, cov, sizeof cov);
  if (ablk9) std::memcpy(ablk9, A, sizeof A);
  if (omega9) lu_inverse(cov, 9, omega9);
}

void omega_imu(const bvio_select_in* in, double* Om) {
  int H = in->H, D = 9 * (H + 1);
  std::fill(Om, Om + (size_t)D * D, 0.0);
  for (int h = 1; h <= H; ++h) {
    double W[81], A[81];
    linear_imu(q4(in->horizon_quat + 4 * (h - 1)), q4(in->horizon_quat + 4 * h), in->nr_imu, in->delta_imu,
               in->acc_var, in->acc_bias_var, W, A, nullptr);
    double AtW[81], AtWA[81];
    for (int i = 0; i < 9; i++)
      for (int j = 0; j < 9; j++) {
        double s = 0;
        for (int k = 0; k < 9; k++) s += A[k * 9 + i] * W[k * 9 + j];
        AtW[i * 9 + j] = s;
      }
    for (int i = 0; i < 9; i++)
      for (int j = 0; j < 9; j++) {
        double s = 0;
        for (int k = 0; k < 9; k++) s += AtW[i * 9 + k] * A[k * 9 + j];
        AtWA[i * 9 + j] = s;
      }
    int r0 = (h - 1) * 9, r1 = h * 9;
    for (int i = 0; i < 9; i++)
      for (int j = 0; j < 9; j++) {
        Om[(size_t)(r0 + i) * D + r0 + j] += AtWA[i * 9 + j];
        Om[(size_t)(r0 + i) * D + r1 + j] += AtW[i * 9 + j];
        Om[(size_t)(r1 + i) * D + r0 + j] += AtW[j * 9 + i];
        Om[(size_t)(r1 + i) * D + r1 + j] += W[i * 9 + j];
      }
  }
  // addOmegaPrior (feature_selector.cpp:602-609): I9 in the reference; a caller-supplied 9 x 9 block when given (ABI v2)
  if (in->omega_prior) {
    for (int i = 0; i < 9; i++)
      for (int j = 0; j < 9; j++) Om[(size_t)i * D + j] += in->omega_prior[i * 9 + j];
  } else {
    for (int i = 0; i < 9; i++) Om[(size_t)i * D + i] += 1.0;
  }
}

void space_to_plane(const bvio_camera& c, V3 P, double px[2]) {
  double mx = P.x / P.z, my = P.y / P.z;
  double mx2 = mx * mx, my2 = my * my, mxy = mx * my, rho2 = mx2 + my2;
  double rad = c.k1 * rho2 + c.k2 * rho2 * rho2;
  double dx = mx * rad + 2.0 * c.p1 * mxy + c.p2 * (rho2 + 2.0 * mx2);
  double dy = my * rad + 2.0 * c.p2 * mxy + c.p1 * (rho2 + 2.0 * my2);
  px[0] = c.fx * (mx + dx) + c.cx;
  px[1] = c.fy * (my + dy) + c.cy;
}
bool in_fov(const bvio_camera& c, const double px[2]) {
  int u = (int)std::round(px[0]), v = (int)std::round(px[1]);
  return (0 <= u && u < c.width) && (0 <= v && v < c.height);
}

// exact 1-NN (squared L2). nanoflann keeps the first-found minimum under strict '<'
// (nanoflann.hpp:175-199); for distinct distances every exact search agrees.
double nn_depth(const bvio_select_in* in, double x, double y) {
  if (in->C == 0) return 1.0;
  int best = 0;
  double bd = std::numeric_limits<double>::max();
  for (int i = 0; i < in->C; i++) {
    double dx = in->cloud_xy[2 * i] - x, dy = in->cloud_xy[2 * i + 1] - y;
    double d = dx * dx + dy * dy;
    if (d < bd) { bd = d; best = i; }
  }
  return in->cloud_depth[best];
}


// one feature: returns false when numVisible == 1. Ch = H blocks (index h-1).
bool feature_blocks(const bvio_select_in* in, double fx, double fy, std::vector<M3>& Ch, M3& W, double* depth) {
  int H = in->H;
  Q4 q_IC = q4(in->q_ic);
  V3 t_IC = v3(in->t_ic);
  // state_k1_ of the reference (feature_selector.cpp:247-250): differs from state_kkH[1] in ground-truth horizon mode
  V3 P1 = v3(in->state_k1_pos ? in->state_k1_pos : in->horizon_pos + 3);
  Q4 Q1 = q4(in->state_k1_quat ? in->state_k1_quat : in->horizon_quat + 4);
  V3 t_WC_k1 = P1 + qrot(Q1, t_IC);
  Q4 q_WC_k1 = qmul(Q1, q_IC);
  V3 feature{fx, fy, 1.0};
  double d = nn_depth(in, fx, fy);
  if (depth) *depth = d;
  feature = normalized(feature) * d;
  V3 pell = t_WC_k1 + qrot(q_WC_k1, feature);
  int numVisible = 1;
  Ch.assign(H, zero3());
  M3 EtE = zero3();
  for (int h = 2; h <= H; ++h) {
    Q4 Qh = q4(in->horizon_quat + 4 * h);
    V3 t_WC_h = v3(in->horizon_pos + 3 * h) + qrot(Qh, t_IC);
    Q4 q_WC_h = qmul(Qh, q_IC);
    V3 uell = normalized(qrot(qinv(q_WC_h), pell - t_WC_h));
    double px[2];
    space_to_plane(in->cam, uell, px);
    if (!in_fov(in->cam, px)) continue;
    M3 Bh = skew(uell) * qmat(qinv(qmul(q_WC_h, q_IC)));
    Ch[h - 1] = transpose(Bh) * Bh;
    EtE = EtE + Ch[h - 1];
    ++numVisible;
  }
  if (numVisible == 1) return false;
  M3 Bh = skew(normalized(feature)) * qmat(qinv(qmul(q_WC_k1, q_IC)));
  Ch[0] = transpose(Bh) * Bh;
  EtE = EtE + Ch[0];
  W = inverse3(EtE);
  return true;
}

}  // namespace

extern "C" {

void oracle_linear_imu_matrices(const double qi[4], const double qj[4], int nr_imu, double delta_imu, double acc_var,
                                double acc_bias_var, double* omega9, double* ablk9, double* cov9) {
  linear_imu(q4(qi), q4(qj), nr_imu, delta_imu, acc_var, acc_bias_var, omega9, ablk9, cov9);
}

void oracle_omega_imu(const bvio_select_in* in, double* omega) { omega_imu(in, omega); }

double oracle_logdet(const double* M, int n) {
  std::vector<double> l(M, M + (size_t)n * n);
  // Eigen LLT does not fail on a non-positive pivot, it takes sqrt of it (-> NaN)
  for (int j = 0; j < n; j++) {
    double d = l[(size_t)j * n + j];
    for (int k = 0; k < j; k++) d -= l[(size_t)j * n + k] * l[(size_t)j * n + k];
    d = std::sqrt(d);
    l[(size_t)j * n + j] = d;
    for (int i = j + 1; i < n; i++) {
      double s = l[(size_t)i * n + j];
      for (int k = 0; k < j; k++) s -= l[(size_t)i * n + k] * l[(size_t)j * n + k];
      l[(size_t)i * n + j] = s / d;
    }
  }
  double ld = 0;
  for (int i = 0; i < n; i++) ld += std::log(l[(size_t)i * n + i]);
  return 2 * ld;
}

void oracle_build_delta(const bvio_select_in* in, int which, double* delta, double* C, int32_t* valid, double* depth) {
  int H = in->H, D = 9 * (H + 1), T = 3 * H;
  int n = which ? in->U : in->N;
  const double* xy = which ? in->used_xy : in->cand_xy;
  std::vector<M3> Ch;
  M3 W;
  for (int f = 0; f < n; f++) {
    double dep;
    bool ok = feature_blocks(in, xy[2 * f], xy[2 * f + 1], Ch, W, &dep);
    if (depth) depth[f] = dep;
    valid[f] = ok;
    double* Df = delta ? delta + (size_t)f * D * D : nullptr;
    double* Cf = C ? C + (size_t)f * T * T : nullptr;
    if (Df) std::fill(Df, Df + (size_t)D * D, 0.0);
    if (Cf) std::fill(Cf, Cf + (size_t)T * T, 0.0);
    if (!ok) continue;
    for (int j = 1; j <= H; ++j)
      for (int i = j; i <= H; ++i) {
        M3 Dij = Ch[i - 1] * W * transpose(Ch[j - 1]);
        for (int a = 0; a < 3; a++)
          for (int b = 0; b < 3; b++) {
            double lo = (i == j) ? Ch[i - 1][a][b] - Dij[a][b] : -Dij[a][b];
            if (Df) {
              Df[(size_t)(9 * i + a) * D + 9 * j + b] = lo;
              if (i != j) Df[(size_t)(9 * j + b) * D + 9 * i + a] = lo;
            }
            if (Cf) {
              Cf[(size_t)(3 * (i - 1) + a) * T + 3 * (j - 1) + b] = lo;
              if (i != j) Cf[(size_t)(3 * (j - 1) + b) * T + 3 * (i - 1) + a] = lo;
            }
          }
      }
  }
}

int oracle_select(const bvio_select_in* in, int32_t* out_ids, double* out_values, bvio_select_summary* sum) {
  int H = in->H, D = 9 * (H + 1), N = in->N;
  size_t DD = (size_t)D * D;
  std::vector<double> Omega(DD);
  omega_imu(in, Omega.data());
  // Delta_ells (std::map<int, omega_horizon_t>) -- key order = ascending id
  std::vector<double> Delta((size_t)N * DD);
  std::vector<int32_t> valid(N);
  oracle_build_delta(in, 0, Delta.data(), nullptr, valid.data(), nullptr);
  std::map<int, int> by_id;  // id -> index, valid only
  int nvalid = 0;
  for (int f = 0; f < N; f++)
    if (valid[f]) { by_id[in->cand_id[f]] = f; nvalid++; }
  // Omega += Delta_used (feature_selector.cpp:620-623)
  if (in->U > 0) {
    std::vector<double> Du((size_t)in->U * DD);
    std::vector<int32_t> vu(in->U);
    oracle_build_delta(in, 1, Du.data(), nullptr, vu.data(), nullptr);
    std::map<int, int> used_by_id;
    for (int f = 0; f < in->U; f++) if (vu[f]) used_by_id[in->used_id[f]] = f;
    for (auto& kv : used_by_id)
      for (size_t k = 0; k < DD; k++) Omega[k] += Du[(size_t)kv.second * DD + k];
  }
  std::vector<int> blacklist;
  std::vector<double> OmegaS(DD, 0.0), M(DD), A(DD);
  int64_t scored = 0;
  double min_margin = std::numeric_limits<double>::infinity();
  int nsel = 0;
  for (int it = 0; it < in->kappa; ++it) {
    // sortedlogDetUB
    std::map<double, int, std::greater<double>> UBs;
    for (size_t k = 0; k < DD; k++) M[k] = Omega[k] + OmegaS[k];
    for (auto& kv : by_id) {
      int fid = kv.first, f = kv.second;
      if (std::find(blacklist.begin(), blacklist.end(), fid) != blacklist.end()) continue;
      double p = in->cand_prob[f];
      const double* Df = &Delta[(size_t)f * DD];
      double ub = 0;
      for (int d = 0; d < D; d++) ub += std::log(M[(size_t)d * D + d] + p * Df[(size_t)d * D + d]);
      UBs[ub] = fid;
    }
    double fMax = -1.0, second = -std::numeric_limits<double>::infinity();
    int lMax = -1;
    for (const auto& fpair : UBs) {
      int fid = fpair.second;
      double ub = fpair.first;
      if (ub < fMax) break;
      int f = by_id.at(fid);
      double p = in->cand_prob[f];
      const double* Df = &Delta[(size_t)f * DD];
      for (size_t k = 0; k < DD; k++) A[k] = Omega[k] + OmegaS[k] + p * Df[k];
      double fValue = oracle_logdet(A.data(), D);
      scored++;
      if (fValue > fMax) { second = fMax; fMax = fValue; lMax = fid; }
      else if (fValue > second) second = fValue;
    }
    if (lMax > -1) {
      int f = by_id.at(lMax);
      double p = in->cand_prob[f];
      const double* Df = &Delta[(size_t)f * DD];
      for (size_t k = 0; k < DD; k++) OmegaS[k] += p * Df[k];
      blacklist.push_back(lMax);
      if (out_ids) out_ids[nsel] = lMax;
      if (out_values) out_values[nsel] = fMax;
      nsel++;
      if (second > -1.0) min_margin = std::min(min_margin, fMax - second);
    }
  }
  if (sum) {
    sum->n_selected = nsel;
    sum->n_candidates_valid = nvalid;
    sum->candidates_scored = scored;
    for (size_t k = 0; k < DD; k++) M[k] = Omega[k] + OmegaS[k];
    sum->final_logdet = oracle_logdet(M.data(), D);
    sum->min_margin = min_margin;
    sum->device_ms = 0;
    sum->transport = 0; sum->world = 1; sum->grid = 0; sum->cpw = 0;
    sum->round_score_us = sum->round_barrier_us = sum->round_exchange_us = 0;
  }
  return BVIO_OK;
}

}  // extern "C"


// oracle_select with the candidates back-projected from an explicit x_{k+1} (the reference's IMU-propagated state_k1_)
// instead of horizon[1]: what FeatureSelector::select does in ground-truth horizon mode.
extern "C" int oracle_select_k1(const bvio_select_in* in, const double k1_pos[3], const double k1_quat[4], int32_t* out_ids,
                                double* out_values, bvio_select_summary* summary) {
  bvio_select_in in2 = *in;
  in2.state_k1_pos = k1_pos; in2.state_k1_quat = k1_quat;
  return oracle_select(&in2, out_ids, out_values, summary);
}
