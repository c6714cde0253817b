/*
 * oracle/oracle.h -- CPU restatement of the reference hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load liboracle.so.  The
 * product library (libbvio.so) never includes, links or calls anything here.
 *
 * PARITY: the reference ships no tests, golden vectors or fixtures for this
 * path, and Eigen / Ceres / ROS / OpenCV are not installed in this image.
 * The reference's own sources for the path are nevertheless compiled here,
 * unmodified and from where they lie under /root/reference, against stand-in
 * Eigen / Ceres / ROS headers (oracle/ref_shim/, `make -C oracle ref` ->
 * oracle/_ref/libvins_ref.so), and tests/test_reference_pin.py checks this
 * restatement against them row by row: projection_factor.cpp,
 * projection_td_factor.cpp, imu_factor.h + integration_base.h,
 * pose_local_parameterization.cpp, marginalization_factor.cpp (loss corrector,
 * marginalize, prior factor), utility.h, feature_manager.cpp,
 * utility/horizon_generator.cpp and feature_selector.cpp (select() end to end,
 * with the vendored nanoflann).
 * estimator.cpp is compiled too: Estimator::optimization() runs with a
 * recording ceres::Problem and with ceres::Solve handing control to the test
 * (vector2double, the problem census / objective / normal equations,
 * double2vector and the marginalization glue are the reference's own).
 * STILL UNPINNED BY REFERENCE CODE: Ceres itself ("tested with 1.14.0",
 * feature_tracker/README.md:7), an un-vendored dependency.  The control logic
 * of its trust-region loop (LM / traditional dogleg, Jacobi scaling, step
 * acceptance, radius update, termination) is restated from its published
 * algorithm twice -- here with Schur elimination, and densely in numpy
 * (tests/np_ref.py), where it drives the reference's live problem -- and the
 * two agree; neither has been checked against a real Ceres build.  The stand-in headers
 * replace Eigen's LLT / inverse / SelfAdjointEigenSolver / JacobiSVD with plain
 * textbook versions, and camodocal's PinholeCamera (needs OpenCV) with a
 * restatement of its spaceToPlane: agreement there is to rounding error or by
 * invariant (J^T J, selected ids), not bit for bit.
 * Further pins: analytic Jacobians vs finite differences with the convention of
 * ProjectionFactor::check (projection_factor.cpp:123-225), the known-answer
 * case of support_files/scripts/createMatricesLinearImuFactor.m, the algebraic
 * identities listed in SURVEY.md section 8c.
 *
 * Same structs as include/bvio.h so that the parity tests feed identical
 * inputs to both sides.
 */
#ifndef BVIO_ORACLE_H_
#define BVIO_ORACLE_H_
#include "../include/bvio.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- factor level (rows a2, a3, a4, a6) ---------------------------------- */
/* ProjectionFactor::Evaluate, projection_factor.cpp:21-121.
 * jac_* row-major 2x7, 2x7, 2x7, 2x1 (any may be NULL). sqrt_info scalar*I2. */
void oracle_projection_factor(const double pts_i[3], const double pts_j[3], const double pose_i[7],
                              const double pose_j[7], const double ex_pose[7], double inv_dep,
                              double sqrt_info, double res[2], double* jac_i, double* jac_j,
                              double* jac_ex, double* jac_f);
/* a2': ProjectionTdFactor::Evaluate (projection_td_factor.cpp:34-141); jacobians 2x7 row-major, jac_f / jac_td 2x1 */
void oracle_projection_td_factor(const double pts_i[3], const double pts_j[3], const double vel_i[2],
                                 const double vel_j[2], double td_i, double td_j, double row_i, double row_j, double TR,
                                 double ROW, const double pose_i[7], const double pose_j[7], const double ex_pose[7],
                                 double inv_dep, double td, double sqrt_info, double res[2], double* jac_i, double* jac_j,
                                 double* jac_ex, double* jac_f, double* jac_td);

/* IMUFactor::Evaluate, imu_factor.h:19-179.  jac row-major 15x7,15x9,15x7,15x9. */
void oracle_imu_factor(const bvio_preint* pre, const double G[3], const double pose_i[7],
                       const double sb_i[9], const double pose_j[7], const double sb_j[9],
                       double res[15], double* jac_pi, double* jac_sbi, double* jac_pj, double* jac_sbj);
/* sqrt_info = LLT(covariance^-1).matrixL().transpose(), imu_factor.h:64 (row-major 15x15) */
void oracle_imu_sqrt_info(const double cov[225], double sqrt_info[225]);
/* MarginalizationFactor::Evaluate residual part, marginalization_factor.cpp:333-365 */
void oracle_prior_residual(const bvio_prior* prior, const bvio_window* w, double* res /*[n]*/, double* dx /*[n]*/);
/* IntegrationBase::propagate, integration_base.h:127-158: one push_back on a bvio_preint
 * whose acc_0/gyr_0 are passed explicitly. */
void oracle_preint_propagate(bvio_preint* pre, double dt, const double acc_0[3], const double gyr_0[3],
                             const double acc_1[3], const double gyr_1[3], double acc_n, double gyr_n,
                             double acc_w, double gyr_w);

/* ---- solver level (rows a1, a5, a7, a8, a9) ------------------------------ */
/* total cost 0.5*sum rho(s) (visual) + 0.5*|r|^2 (imu, prior) at the window's state */
double oracle_cost(const bvio_window* w, const bvio_opts* opts);
/* undamped reduced system at the window's state; layouts as bvio_debug_linearize */
int oracle_linearize(const bvio_window* w, const bvio_opts* opts, double* S, double* g, double* h,
                     double* b, double* cost);
/* Full solve. opts->strategy selects LM (what the device runs) or DOGLEG (what the
 * reference configures, estimator.cpp:798).  Overwrites para_* / inv_depth. */
int oracle_optimize(bvio_window* w, const bvio_opts* opts, bvio_summary* summary);
/* Estimator::double2vector gauge re-anchoring, estimator.cpp:521-555: pose0/R0 are the
 * pre-solve Ps[0], Rs[0] (as para_pose row 0 before the solve). Rewrites para_pose,
 * para_speed_bias velocity in place. */
void oracle_double2vector(const double pre_pose0[7], int K, double* para_pose, double* para_speed_bias);
/* Marginalization, estimator.cpp:816-991 + marginalization_factor.cpp:89-319 */
int oracle_marginalize(const bvio_window* w, const bvio_opts* opts, int flag, bvio_prior_out* out);

/* ---- selector (rows a10-a15) -------------------------------------------- */
/* Omega_kkH incl. addOmegaPrior: [9(H+1)]^2 row-major (feature_selector.cpp:463-609) */
void oracle_omega_imu(const bvio_select_in* in, double* omega);
/* createLinearImuMatrices for one pair: Omega (9x9 = covImu^-1) and Ablk, row-major */
void oracle_linear_imu_matrices(const double qi[4], const double qj[4], int nr_imu, double delta_imu,
                                double acc_var, double acc_bias_var, double* omega9, double* ablk9,
                                double* cov9);
/* calcInfoFromFeatures for the candidates (or the used set when which=1):
 * dense Delta_ell [n][D*D] (D = 9(H+1)) when delta != NULL, compact C [n][3H*3H]
 * when C != NULL, valid[n] (0 when numVisible == 1), depth[n] the NN depth used. */
void oracle_build_delta(const bvio_select_in* in, int which, double* delta, double* C, int32_t* valid,
                        double* depth);
/* Literal lazy-greedy select (feature_selector.cpp:613-728) on dense D x D matrices. */
int oracle_select(const bvio_select_in* in, int32_t* out_ids, double* out_values,
                  bvio_select_summary* summary);
/* oracle_select with the candidates back-projected from an explicit x_{k+1} (state_k1_, feature_selector.cpp:247-250)
 * instead of horizon[1]: FeatureSelector::select in ground-truth horizon mode.  Not yet in the C-ABI. */
int oracle_window_omega_prior(const bvio_window* w, const bvio_opts* opts, double* omega9);
int oracle_select_k1(const bvio_select_in* in, const double k1_pos[3], const double k1_quat[4], int32_t* out_ids,
                     double* out_values, bvio_select_summary* summary);
/* Utility::logdet(M, true), utility.h:143-167 */
double oracle_logdet(const double* M, int n);

/* f2: FeatureManager::triangulate (feature_manager.cpp:202-257): DLT depth of every landmark, depth_out[L] */
int oracle_triangulate(const bvio_window* w, double init_depth, double* depth_out);

/* f3: HorizonGenerator::imu (utility/horizon_generator.cpp:25-70) */
void oracle_horizon_imu(int H, const double pos0[3], const double quat0[4], const double ba0[3], const double pos1[3],
                        const double quat1[4], const double vel1[3], const double acc[3], const double gyr[3], int nr_imu,
                        double delta_imu, double* horizon_pos, double* horizon_quat);

#ifdef __cplusplus
}
#endif
#endif
