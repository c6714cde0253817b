// interpose.cpp -- TEST GLUE (not part of what a maintainer ships): puts bvio_adapter behind the reference's UNMODIFIED
// Estimator::optimization() and FeatureSelector::select() at link time, so that the drop-in can be proven with the
// reference's own classes driving libbvio.so on a B200.  Built into oracle/_ref/libvins_bvio.so (oracle/Makefile target
// `ref_bvio`) together with the reference's sources compiled from where they lie.
//
// How the three call sites are reached without editing the reference:
//   * ceres::Solve (estimator.cpp:809): the stand-in ceres.h forwards Solve() to ceres::solve_hook(); the hook below calls
//     bvio_adapter::solve().  (In the reference's real build the maintainer replaces that one line instead.)
//   * MarginalizationInfo::preMarginalize / marginalize (estimator.cpp:897-901, 952-957): marginalization_factor.cpp is
//     compiled with -DpreMarginalize=preMarginalize_reference -Dmarginalize=marginalize_reference, which renames the
//     reference's two definitions; the definitions below take their place (same class, same header).
//   * FeatureSelector::{calcInfoFromRobotMotion, calcInfoFromFeatures, selectInformativeFeatures}
//     (feature_selector.cpp:139-171): called from select() inside the same translation unit, so their definitions in
//     feature_selector.o are made weak (objcopy -W) and the strong versions below win at link time; they capture the
//     horizon, skip the host-side information matrices and hand the selection to bvio_adapter::select().
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "bvio_adapter.h"

namespace {
bvio_ctx* g_ctx = nullptr;
Estimator* g_est = nullptr;
int g_counts[4] = {0, 0, 0, 0};            // solves, marginalizations, selections routed to libbvio; last bvio_status
bvio_summary g_last_summary;
bvio_select_summary g_last_select;
struct Captured { state_horizon_t state_kkH; int nr_imu; double delta_imu; bool valid = false; } g_cap;

void die(const char* where, int rc) {
  std::fprintf(stderr, "[bvio glue] %s failed: %d (%s)\n", where, rc, g_ctx ? bvio_last_error(g_ctx) : "no context");
  std::abort();
}

// what optimization() handed to "Ceres" must be the problem the adapter hands to libbvio (estimator.cpp:694-792)
void check_problem(const ceres::Problem* pb, const bvio_adapter::Window& win) {
  const bvio_window& w = win.w;
  int n_imu = 0;
  for (int j = 1; j < w.K; ++j) n_imu += !(w.preint[j].sum_dt > 10.0);
  const int n_proj = (w.L ? w.lm_obs_offset[w.L] : 0) - w.L;
  const int want = (w.prior ? 1 : 0) + n_imu + n_proj + w.n_relo;
  if ((int)pb->residual_blocks.size() != want) {
    std::fprintf(stderr, "[bvio glue] problem mismatch: reference built %d residual blocks, adapter window holds %d\n",
                 (int)pb->residual_blocks.size(), want);
    std::abort();
  }
}

void glue_solve_hook(const ceres::Solver::Options& o, ceres::Problem* pb, ceres::Solver::Summary* s) {
  if (!g_ctx || !g_est) die("solve (glue not attached)", BVIO_ERR_INVALID);
  bvio_adapter::Window win;
  bvio_adapter::fill_window(*g_est, &win);
  check_problem(pb, win);
  bvio_opts opts = bvio_adapter::make_opts(*g_est);
  if (opts.max_iters != o.max_num_iterations || (o.trust_region_strategy_type == ceres::DOGLEG) != (opts.strategy == BVIO_STRATEGY_DOGLEG))
    die("solve (options mismatch)", BVIO_ERR_INVALID);
  if (std::getenv("BVIO_GLUE_NO_TIME_CUT")) opts.max_time_s = 0.0;
  int rc = bvio_optimize(g_ctx, &win.w, &opts, &g_last_summary);
  g_counts[0]++; g_counts[3] = rc;
  if (rc != BVIO_OK) die("bvio_optimize", rc);
  if (s) s->iterations.assign(g_last_summary.iterations, 0);
}
}  // namespace

// ---- link-time replacements of the reference's member functions ----------------------------------------------------------------
void MarginalizationInfo::preMarginalize() {}        // both halves happen in marginalize() below (bvio_adapter::marginalize)

void MarginalizationInfo::marginalize() {
  if (!g_ctx || !g_est) die("marginalize (glue not attached)", BVIO_ERR_INVALID);
  int rc = bvio_adapter::marginalize(g_ctx, *g_est, this);
  g_counts[1]++; g_counts[3] = rc;
  if (rc != BVIO_OK) die("bvio_marginalize", rc);
}

omega_horizon_t FeatureSelector::calcInfoFromRobotMotion(const state_horizon_t& x_kkH, double nrImuMeasurements, double deltaImu) {
  g_cap.state_kkH = x_kkH; g_cap.nr_imu = (int)nrImuMeasurements; g_cap.delta_imu = deltaImu; g_cap.valid = true;
  return omega_horizon_t::Zero();                     // Omega_kkH is built on the device (sel_omega_kernel)
}

std::map<int, omega_horizon_t> FeatureSelector::calcInfoFromFeatures(const image_t&, const state_horizon_t&) {
  return {};                                          // Delta_ell: built on the device (sel_build_kernel)
}

std::vector<int> FeatureSelector::selectInformativeFeatures(image_t& subset, const image_t& image, int kappa, const omega_horizon_t&,
                                                            const std::map<int, omega_horizon_t>&, const std::map<int, omega_horizon_t>&) {
  if (!g_ctx || !g_cap.valid) die("select (glue not attached)", BVIO_ERR_INVALID);
  bvio_camera cam;
  {
    const camodocal::PinholeParams& c = static_cast<camodocal::PinholeCameraShim*>(m_camera_.get())->params();
    cam.fx = c.fx; cam.fy = c.fy; cam.cx = c.cx; cam.cy = c.cy; cam.k1 = c.k1; cam.k2 = c.k2; cam.p1 = c.p1; cam.p2 = c.p2;
    cam.width = c.width; cam.height = c.height;
  }
  bvio_adapter::SelectInputs in;
  bvio_adapter::fill_select_in(estimator_, g_cap.state_kkH, state_k1_, q_IC_, t_IC_, cam, g_cap.nr_imu, g_cap.delta_imu,
                               accVarDTime_, accBiasVarDTime_, image, subset, kappa, &in);
  std::vector<int> ids;
  int rc = bvio_adapter::select(g_ctx, in, &ids, &g_last_select);
  g_counts[2]++; g_counts[3] = rc;
  if (rc != BVIO_OK) die("bvio_select", rc);
  for (int id : ids) subset[id] = image.at(id);       // feature_selector.cpp:677
  g_cap.valid = false;
  return ids;
}

// ---- C entry points for the tests ------------------------------------------------------------------------------------------------
extern "C" {
extern void (*ref_on_estimator_created)(void* estimator);   // ref_driver.cpp
void bvio_glue_attach(void* estimator);
// route the Estimator that the next ref_estimator_optimization() calls build (ref_driver.cpp) through libbvio
void bvio_glue_capture_created(int on) { ref_on_estimator_created = on ? bvio_glue_attach : nullptr; }
int bvio_glue_enable(int device) {
  if (g_ctx) return 0;
  return bvio_create(device, &g_ctx);
}
// attach to a live Estimator created by ref_est_create (ref_driver.cpp): from now on its optimization() solves and
// marginalizes through libbvio
void bvio_glue_attach(void* estimator) {
  g_est = static_cast<Estimator*>(estimator);
  ceres::solve_hook() = glue_solve_hook;
}
void bvio_glue_disable(void) {
  g_est = nullptr;
  ceres::solve_hook() = nullptr;
  if (g_ctx) { bvio_destroy(g_ctx); g_ctx = nullptr; }
}
void bvio_glue_counts(int32_t out[4]) { for (int i = 0; i < 4; ++i) out[i] = g_counts[i]; }
void bvio_glue_last_summary(bvio_summary* s, bvio_select_summary* ss) { if (s) *s = g_last_summary; if (ss) *ss = g_last_select; }
long long bvio_glue_launches(void) { return g_ctx ? bvio_launch_count(g_ctx) : 0; }
}
