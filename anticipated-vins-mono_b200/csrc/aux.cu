// aux.cu -- the steps either side of the hot path, on device behind the same C-ABI:
//   bvio_triangulate   FeatureManager::triangulate (vins_estimator/src/feature_manager.cpp:202-257): DLT depth of
//                      every landmark from all its observations, the step right before optimization() in
//                      Estimator::solveOdometry (estimator.cpp:471)
// One warp per landmark: lane = row of the (2 n_obs) x 4 DLT matrix (n_obs <= 16 -> 32 rows, exactly one warp); the
// right singular vector of the smallest singular value comes from a one-sided (Hestenes) Jacobi SVD whose column dot
// products are warp reductions -- the same accuracy class as the Eigen::JacobiSVD the reference calls (:243), no
// normal equations.
#include "ba.h"
#include "ctx.h"
#include <string>

using namespace bvio;

namespace {

__global__ void __launch_bounds__(256) tri_kernel(int L, int K, const double* __restrict__ pose, const double* __restrict__ ex,
                                                  const int* __restrict__ lm_off, const int* __restrict__ obs_frame,
                                                  const double2* __restrict__ obs_xy, double init_depth, double* __restrict__ depth) {
  const int lane = threadIdx.x & 31, l = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (l >= L) return;
  const int o0 = lm_off[l], n = lm_off[l + 1] - o0;
  double Ric[9];
  qmat(q4{ex[3], ex[4], ex[5], ex[6]}, Ric);
  const d3 tic{ex[0], ex[1], ex[2]};
  // anchor camera frame: R0 = Rs[i] ric, t0 = Ps[i] + Rs[i] tic  (:220-221)
  const int fi = obs_frame[o0];
  const double* pi = pose + (size_t)fi * 7;
  double Ri[9], R0[9];
  qmat(q4{pi[3], pi[4], pi[5], pi[6]}, Ri);
  mm3(Ri, Ric, R0);
  const d3 t0 = d3{pi[0], pi[1], pi[2]} + mv3(Ri, tic);
  // this lane's row: observation lane/2, row lane&1
  double a[4] = {0, 0, 0, 0};
  const int ko = lane >> 1;
  if (ko < n) {
    const int fj = obs_frame[o0 + ko];
    const double* pj = pose + (size_t)fj * 7;
    double Rj[9], R1[9], R[9];
    qmat(q4{pj[3], pj[4], pj[5], pj[6]}, Rj);
    mm3(Rj, Ric, R1);
    const d3 t1 = d3{pj[0], pj[1], pj[2]} + mv3(Rj, tic);
    const d3 t = mtv3(R0, t1 - t0);              // R0^T (t1 - t0)
    mtm3(R0, R1, R);                             // R0^T R1
    // P = [R^T | -R^T t]  (:233-235): row r of P = (column r of R, -column r of R . t)
    double P[3][4];
#pragma unroll
    for (int r = 0; r < 3; r++) {
      P[r][0] = R[0 * 3 + r]; P[r][1] = R[1 * 3 + r]; P[r][2] = R[2 * 3 + r];
      P[r][3] = -(R[0 * 3 + r] * t.x + R[1 * 3 + r] * t.y + R[2 * 3 + r] * t.z);
    }
    const double2 xy = obs_xy[o0 + ko];
    const double nrm = sqrt(xy.x * xy.x + xy.y * xy.y + 1.0);
    const double f0 = xy.x / nrm, f1 = xy.y / nrm, f2 = 1.0 / nrm;       // point.normalized()  (:236)
    const int rr = lane & 1;
    const double fr = rr ? f1 : f0;
#pragma unroll
    for (int c = 0; c < 4; c++) a[c] = fr * P[2][c] - f2 * P[rr][c];     // (:237-238)
  }
  // one-sided Jacobi: rotate column pairs until they are mutually orthogonal; V accumulates the rotations
  double V[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
  for (int sweep = 0; sweep < 30; sweep++) {
    double off = 0;
#pragma unroll
    for (int p = 0; p < 3; p++)
#pragma unroll
      for (int q = p + 1; q < 4; q++) {
        const double al = warp_sum(a[p] * a[p]), be = warp_sum(a[q] * a[q]), ga = warp_sum(a[p] * a[q]);
        if (ga == 0.0) continue;
        off = fmax(off, fabs(ga) / sqrt(al * be + 1e-300));
        const double zeta = (be - al) / (2.0 * ga);
        const double tt = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + tt * tt), s = c * tt;
        const double ap = a[p], aq = a[q];
        a[p] = c * ap - s * aq; a[q] = s * ap + c * aq;
#pragma unroll
        for (int r = 0; r < 4; r++) { const double vp = V[r][p], vq = V[r][q]; V[r][p] = c * vp - s * vq; V[r][q] = s * vp + c * vq; }
      }
    if (off <= 1e-15) break;
  }
  // smallest singular value = smallest column norm
  double best = warp_sum(a[0] * a[0]);
  int bc = 0;
#pragma unroll
  for (int c = 1; c < 4; c++) { const double nn = warp_sum(a[c] * a[c]); if (nn < best) { best = nn; bc = c; } }
  if (lane == 0) {
    double v2 = V[2][0], v3 = V[3][0];
#pragma unroll
    for (int c = 1; c < 4; c++) if (bc == c) { v2 = V[2][c]; v3 = V[3][c]; }
    double d = v2 / v3;                                                  // svd_V[2] / svd_V[3]  (:244)
    if (!(d >= 0.1)) d = init_depth;                                     // (:251-254); NaN falls back too
    depth[l] = d;
  }
}

}  // namespace

extern "C" int bvio_triangulate(bvio_ctx* ctx, const bvio_window* w, double init_depth, double* depth_out) {
  if (!ctx || !w || !depth_out) return fail(ctx, BVIO_ERR_INVALID, "null argument");
  if (w->K < 1 || w->L < 0 || !w->para_pose || !w->para_ex_pose) return fail(ctx, BVIO_ERR_INVALID, "null state arrays");
  if (w->L == 0) return BVIO_OK;
  if (!w->lm_obs_offset || !w->obs_frame || !w->obs_xy) return fail(ctx, BVIO_ERR_INVALID, "null landmark arrays");
  const int L = w->L, K = w->K, nobs = w->lm_obs_offset[L];
  for (int l = 0; l < L; l++) {
    const int n = w->lm_obs_offset[l + 1] - w->lm_obs_offset[l];
    if (n < 1 || n > BVIO_KMAX) return fail(ctx, BVIO_ERR_INVALID, "landmark needs 1..16 observations");
  }
  for (int k = 0; k < nobs; k++)
    if (w->obs_frame[k] < 0 || w->obs_frame[k] >= K) return fail(ctx, BVIO_ERR_INVALID, "obs_frame out of range");
  cudaSetDevice(ctx->device);
  Carver cv;
  const size_t D = sizeof(double), I = sizeof(int);
  size_t o_pose = cv.take((size_t)K * 7 * D), o_ex = cv.take(7 * D), o_off = cv.take((size_t)(L + 1) * I);
  size_t o_fr = cv.take((size_t)nobs * I), o_xy = cv.take((size_t)nobs * 2 * D);
  const size_t in_bytes = cv.off;
  size_t o_out = cv.take((size_t)L * D);
  Slab slab;
  bool from_cache = false;
  cudaError_t e = slab_acquire(ctx->sel_cache, ctx->sel_cache_busy, cv.off, cv.off, slab, from_cache);
  if (e != cudaSuccess) return fail(ctx, BVIO_ERR_CUDA, std::string("triangulate alloc: ") + cudaGetErrorString(e));
  memcpy(slab.h + o_pose, w->para_pose, (size_t)K * 7 * D);
  memcpy(slab.h + o_ex, w->para_ex_pose, 7 * D);
  memcpy(slab.h + o_off, w->lm_obs_offset, (size_t)(L + 1) * I);
  memcpy(slab.h + o_fr, w->obs_frame, (size_t)nobs * I);
  memcpy(slab.h + o_xy, w->obs_xy, (size_t)nobs * 2 * D);
  e = cudaMemcpyAsync(slab.d, slab.h, in_bytes, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) {
    tri_kernel<<<(L + 7) / 8, 256, 0, ctx->stream>>>(L, K, (const double*)(slab.d + o_pose), (const double*)(slab.d + o_ex),
                                                    (const int*)(slab.d + o_off), (const int*)(slab.d + o_fr),
                                                    (const double2*)(slab.d + o_xy), init_depth, (double*)(slab.d + o_out));
    ctx->launches += 1;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(slab.h + o_out, slab.d + o_out, (size_t)L * D, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e == cudaSuccess) memcpy(depth_out, slab.h + o_out, (size_t)L * D);
  slab_release(slab, ctx->sel_cache_busy, from_cache);
  if (e != cudaSuccess) return fail(ctx, BVIO_ERR_CUDA, std::string("triangulate: ") + cudaGetErrorString(e));
  return BVIO_OK;
}
