// aux.cu -- the steps either side of the hot path, on device behind the same C-ABI:
//   bvio_triangulate   FeatureManager::triangulate (vins_estimator/src/feature_manager.cpp:202-257): DLT depth of
//                      every landmark from all its observations, the step right before optimization() in
//                      Estimator::solveOdometry (estimator.cpp:471)
//   bvio_preintegrate  IntegrationBase::{push_back, propagate, midPointIntegration, repropagate}
//                      (vins_estimator/src/factor/integration_base.h:30-158): the IMU factors' inputs
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

// ---------------------------------------------------------------------------------------------
// IntegrationBase::{push_back, propagate, midPointIntegration} (integration_base.h:30-158): one CTA per frame
// interval.  The state (delta_p/q/v) is advanced by thread 0; the 15x15 products jacobian = F jacobian and
// covariance = F cov F^T + V noise V^T (:142-143) are spread over the CTA, one output entry per thread.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void put33(double* M, int ld, int r0, int c0, const double* m, double s) {
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) M[(r0 + i) * ld + c0 + j] = s * m[i * 3 + j];
}

__global__ void __launch_bounds__(256) preint_kernel(int nseg, const int* __restrict__ seg_off, const double* __restrict__ dts,
                                                     const double* __restrict__ acc, const double* __restrict__ gyr,
                                                     const double* __restrict__ lin_b /*[nseg][6]*/, double acc_n, double gyr_n,
                                                     double acc_w, double gyr_w, double* __restrict__ out /*[nseg][467]*/) {
  __shared__ double F[15 * 15], V[15 * 18], J[225], Cv[225], T[225], st[16], noise[18];
  const int sg = blockIdx.x, tid = threadIdx.x;
  if (sg >= nseg) return;
  const int s0 = seg_off[sg], ns = seg_off[sg + 1] - s0 - 1;     // ns propagation steps from ns + 1 samples
  const double* ba = lin_b + sg * 6;
  const double* bg = ba + 3;
  for (int i = tid; i < 225; i += blockDim.x) { J[i] = (i / 15 == i % 15) ? 1.0 : 0.0; Cv[i] = 0.0; }
  if (tid < 18) noise[tid] = (tid < 3 || (tid >= 6 && tid < 9)) ? acc_n * acc_n : (tid < 12 ? gyr_n * gyr_n : (tid < 15 ? acc_w * acc_w : gyr_w * gyr_w));
  if (tid == 0) { for (int i = 0; i < 16; i++) st[i] = 0.0; st[6] = 1.0; }   // dp(0..2) dq xyzw(3..6) dv(7..9) sum_dt(10)
  __syncthreads();
  for (int k = 0; k < ns; k++) {
    for (int i = tid; i < 225; i += blockDim.x) F[i] = 0.0;
    for (int i = tid; i < 270; i += blockDim.x) V[i] = 0.0;
    __syncthreads();
    if (tid == 0) {
      const double dt = dts[s0 + k + 1];
      const d3 a0{acc[3 * (s0 + k)] - ba[0], acc[3 * (s0 + k) + 1] - ba[1], acc[3 * (s0 + k) + 2] - ba[2]};
      const d3 a1{acc[3 * (s0 + k + 1)] - ba[0], acc[3 * (s0 + k + 1) + 1] - ba[1], acc[3 * (s0 + k + 1) + 2] - ba[2]};
      const d3 wx{0.5 * (gyr[3 * (s0 + k)] + gyr[3 * (s0 + k + 1)]) - bg[0], 0.5 * (gyr[3 * (s0 + k) + 1] + gyr[3 * (s0 + k + 1) + 1]) - bg[1],
                  0.5 * (gyr[3 * (s0 + k) + 2] + gyr[3 * (s0 + k + 1) + 2]) - bg[2]};
      const q4 dq{st[3], st[4], st[5], st[6]};
      const d3 dp{st[0], st[1], st[2]}, dv{st[7], st[8], st[9]};
      // midPointIntegration (:64-88)
      const d3 un_acc_0 = qrot(dq, a0);
      const q4 rq = qmul(dq, q4{wx.x * dt / 2, wx.y * dt / 2, wx.z * dt / 2, 1.0});
      const d3 un_acc_1 = qrot(rq, a1);
      const d3 un_acc = 0.5 * (un_acc_0 + un_acc_1);
      const d3 rp = dp + dt * dv + (0.5 * dt * dt) * un_acc;
      const d3 rv = dv + dt * un_acc;
      double Rq[9], Rr[9], Rw[9], Ra0[9], Ra1[9], I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, m1[9], m2[9], ImW[9];
      qmat(dq, Rq); qmat(rq, Rr);
      const double sk[3][9] = {{0, -wx.z, wx.y, wx.z, 0, -wx.x, -wx.y, wx.x, 0}, {0, -a0.z, a0.y, a0.z, 0, -a0.x, -a0.y, a0.x, 0},
                               {0, -a1.z, a1.y, a1.z, 0, -a1.x, -a1.y, a1.x, 0}};
      for (int i = 0; i < 9; i++) { Rw[i] = sk[0][i]; Ra0[i] = sk[1][i]; Ra1[i] = sk[2][i]; ImW[i] = I3[i] - dt * Rw[i]; }
      // F (:93-112)
      double RqA0[9], RrA1[9], RrA1ImW[9], sum[9];
      mm3(Rq, Ra0, RqA0); mm3(Rr, Ra1, RrA1); mm3(RrA1, ImW, RrA1ImW);
      put33(F, 15, 0, 0, I3, 1.0);
      for (int i = 0; i < 9; i++) sum[i] = (-0.25 * dt * dt) * RqA0[i] + (-0.25 * dt * dt) * RrA1ImW[i];
      put33(F, 15, 0, 3, sum, 1.0);
      put33(F, 15, 0, 6, I3, dt);
      for (int i = 0; i < 9; i++) m1[i] = Rq[i] + Rr[i];
      put33(F, 15, 0, 9, m1, -0.25 * dt * dt);
      put33(F, 15, 0, 12, RrA1, -0.25 * dt * dt * -dt);
      put33(F, 15, 3, 3, ImW, 1.0);
      put33(F, 15, 3, 12, I3, -dt);
      for (int i = 0; i < 9; i++) sum[i] = (-0.5 * dt) * RqA0[i] + (-0.5 * dt) * RrA1ImW[i];
      put33(F, 15, 6, 3, sum, 1.0);
      put33(F, 15, 6, 6, I3, 1.0);
      put33(F, 15, 6, 9, m1, -0.5 * dt);
      put33(F, 15, 6, 12, RrA1, -0.5 * dt * -dt);
      put33(F, 15, 9, 9, I3, 1.0);
      put33(F, 15, 12, 12, I3, 1.0);
      // V (:115-127)
      for (int i = 0; i < 9; i++) m2[i] = -RrA1[i];
      put33(V, 18, 0, 0, Rq, 0.25 * dt * dt);
      put33(V, 18, 0, 3, m2, 0.25 * dt * dt * 0.5 * dt);
      put33(V, 18, 0, 6, Rr, 0.25 * dt * dt);
      put33(V, 18, 0, 9, m2, 0.25 * dt * dt * 0.5 * dt);
      put33(V, 18, 3, 3, I3, 0.5 * dt);
      put33(V, 18, 3, 9, I3, 0.5 * dt);
      put33(V, 18, 6, 0, Rq, 0.5 * dt);
      put33(V, 18, 6, 3, m2, 0.5 * dt * 0.5 * dt);
      put33(V, 18, 6, 6, Rr, 0.5 * dt);
      put33(V, 18, 6, 9, m2, 0.5 * dt * 0.5 * dt);
      put33(V, 18, 9, 12, I3, dt);
      put33(V, 18, 12, 15, I3, dt);
      const q4 qn = qnormalized(rq);
      st[0] = rp.x; st[1] = rp.y; st[2] = rp.z; st[3] = qn.x; st[4] = qn.y; st[5] = qn.z; st[6] = qn.w;
      st[7] = rv.x; st[8] = rv.y; st[9] = rv.z; st[10] += dt;
    }
    __syncthreads();
    double jn = 0, tn = 0;
    const int i = tid / 15, j = tid - 15 * i;
    if (tid < 225) {
      for (int q = 0; q < 15; q++) { jn += F[i * 15 + q] * J[q * 15 + j]; tn += F[i * 15 + q] * Cv[q * 15 + j]; }
    }
    __syncthreads();
    if (tid < 225) { J[tid] = jn; T[tid] = tn; }
    __syncthreads();
    if (tid < 225) {
      double cn = 0;
      for (int q = 0; q < 15; q++) cn += T[i * 15 + q] * F[j * 15 + q];
      for (int q = 0; q < 18; q++) cn += V[i * 18 + q] * noise[q] * V[j * 18 + q];
      Cv[tid] = cn;
    }
    __syncthreads();
  }
  double* o = out + (size_t)sg * PREINT_DOUBLES;
  if (tid < 3) { o[tid] = st[tid]; o[7 + tid] = st[7 + tid]; o[10 + tid] = ba[tid]; o[13 + tid] = bg[tid]; }
  if (tid < 4) o[3 + tid] = st[3 + tid];
  if (tid == 0) o[16] = st[10];
  for (int e = tid; e < 225; e += blockDim.x) { o[17 + e] = J[e]; o[17 + 225 + e] = Cv[e]; }
}

}  // namespace

extern "C" int bvio_preintegrate(bvio_ctx* ctx, const bvio_imu_segment* segs, int32_t n_segs, double acc_n, double gyr_n,
                                 double acc_w, double gyr_w, bvio_preint* out) {
  if (!ctx || !segs || n_segs < 1 || !out) return fail(ctx, BVIO_ERR_INVALID, "null argument");
  static_assert(sizeof(bvio_preint) == PREINT_DOUBLES * sizeof(double), "bvio_preint layout");
  int total = 0;
  for (int s = 0; s < n_segs; s++) {
    if (segs[s].n_samples < 1 || !segs[s].dt || !segs[s].acc || !segs[s].gyr) return fail(ctx, BVIO_ERR_INVALID, "bad IMU segment");
    total += segs[s].n_samples;
  }
  cudaSetDevice(ctx->device);
  Carver cv;
  const size_t D = sizeof(double), I = sizeof(int);
  size_t o_off = cv.take((size_t)(n_segs + 1) * I), o_dt = cv.take((size_t)total * D), o_acc = cv.take((size_t)total * 3 * D);
  size_t o_gyr = cv.take((size_t)total * 3 * D), o_b = cv.take((size_t)n_segs * 6 * D);
  const size_t in_bytes = cv.off;
  size_t o_out = cv.take((size_t)n_segs * PREINT_DOUBLES * D);
  Slab slab;
  bool from_cache = false;
  cudaError_t e = slab_acquire(ctx->sel_cache, ctx->sel_cache_busy, cv.off, cv.off, slab, from_cache);
  if (e != cudaSuccess) return fail(ctx, BVIO_ERR_CUDA, std::string("preintegrate alloc: ") + cudaGetErrorString(e));
  int* h_off = (int*)(slab.h + o_off);
  int run = 0;
  for (int s = 0; s < n_segs; s++) {
    h_off[s] = run;
    memcpy(slab.h + o_dt + (size_t)run * D, segs[s].dt, (size_t)segs[s].n_samples * D);
    memcpy(slab.h + o_acc + (size_t)run * 3 * D, segs[s].acc, (size_t)segs[s].n_samples * 3 * D);
    memcpy(slab.h + o_gyr + (size_t)run * 3 * D, segs[s].gyr, (size_t)segs[s].n_samples * 3 * D);
    memcpy(slab.h + o_b + (size_t)s * 6 * D, segs[s].lin_ba, 3 * D);
    memcpy(slab.h + o_b + (size_t)s * 6 * D + 3 * D, segs[s].lin_bg, 3 * D);
    run += segs[s].n_samples;
  }
  h_off[n_segs] = run;
  e = cudaMemcpyAsync(slab.d, slab.h, in_bytes, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) {
    preint_kernel<<<n_segs, 256, 0, ctx->stream>>>(n_segs, (const int*)(slab.d + o_off), (const double*)(slab.d + o_dt),
                                                   (const double*)(slab.d + o_acc), (const double*)(slab.d + o_gyr),
                                                   (const double*)(slab.d + o_b), acc_n, gyr_n, acc_w, gyr_w, (double*)(slab.d + o_out));
    ctx->launches += 1;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(slab.h + o_out, slab.d + o_out, (size_t)n_segs * PREINT_DOUBLES * D, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e == cudaSuccess) memcpy(out, slab.h + o_out, (size_t)n_segs * PREINT_DOUBLES * D);
  slab_release(slab, ctx->sel_cache_busy, from_cache);
  if (e != cudaSuccess) return fail(ctx, BVIO_ERR_CUDA, std::string("preintegrate: ") + cudaGetErrorString(e));
  return BVIO_OK;
}

// ---------------------------------------------------------------------------------------------
// HorizonGenerator::imu (utility/horizon_generator.cpp:25-70): constant-acceleration, constant-rate propagation of
// x_{k+1} over the horizon.  A recurrence of (H-1) * nr_imu dependent steps: one thread.
// ---------------------------------------------------------------------------------------------
namespace {
__global__ void horizon_imu_kernel(int H, const double* __restrict__ in /*pos0 quat0 ba0 pos1 quat1 vel1 a w: 3 4 3 3 4 3 3 3*/,
                                   int nr_imu, double delta, double* __restrict__ hpos, double* __restrict__ hquat) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const d3 grav{0.0, 0.0, -9.80665};                                     // state_defs.h:37-41
  const d3 ba{in[7], in[8], in[9]};
  d3 p{in[10], in[11], in[12]}, v{in[17], in[18], in[19]};
  q4 q{in[13], in[14], in[15], in[16]};
  const d3 a{in[20] - ba.x, in[21] - ba.y, in[22] - ba.z};
  const q4 qimu{in[23] * delta / 2, in[24] * delta / 2, in[25] * delta / 2, 1.0};   // Utility::deltaQ(w * deltaImu), not normalised
  for (int i = 0; i < 3; i++) { hpos[i] = in[i]; hpos[3 + i] = in[10 + i]; }
  for (int i = 0; i < 4; i++) { hquat[i] = in[3 + i]; hquat[4 + i] = in[13 + i]; }
  for (int h = 2; h <= H; h++) {
    for (int i = 0; i < nr_imu; i++) {
      q = qmul(q, qimu);                                                 // :52
      const d3 qa = qrot(q, a);                                          // Eigen quaternion * vector
      v = v + delta * (grav + qa);                                       // :58
      p = p + delta * v + (0.5 * delta * delta) * grav + (0.5 * delta * delta) * qa;   // :61
    }
    hpos[3 * h] = p.x; hpos[3 * h + 1] = p.y; hpos[3 * h + 2] = p.z;
    hquat[4 * h] = q.x; hquat[4 * h + 1] = q.y; hquat[4 * h + 2] = q.z; hquat[4 * h + 3] = q.w;
  }
}
}  // namespace

extern "C" int bvio_horizon_imu(bvio_ctx* ctx, int32_t H, const double* pos0, const double* quat0, const double* ba0,
                                const double* pos1, const double* quat1, const double* vel1, const double* acc,
                                const double* gyr, int32_t nr_imu, double delta_imu, double* horizon_pos,
                                double* horizon_quat) {
  if (!ctx || !pos0 || !quat0 || !ba0 || !pos1 || !quat1 || !vel1 || !acc || !gyr || !horizon_pos || !horizon_quat)
    return fail(ctx, BVIO_ERR_INVALID, "null argument");
  if (H < 1 || H > BVIO_HMAX || nr_imu < 0) return fail(ctx, BVIO_ERR_INVALID, "H out of range [1,16] / nr_imu < 0");
  cudaSetDevice(ctx->device);
  Carver cv;
  const size_t D = sizeof(double);
  size_t o_in = cv.take(26 * D);
  const size_t in_bytes = cv.off;
  size_t o_pos = cv.take((size_t)(H + 1) * 3 * D), o_quat = cv.take((size_t)(H + 1) * 4 * D);
  Slab slab;
  bool from_cache = false;
  cudaError_t e = slab_acquire(ctx->sel_cache, ctx->sel_cache_busy, cv.off, cv.off, slab, from_cache);
  if (e != cudaSuccess) return fail(ctx, BVIO_ERR_CUDA, std::string("horizon alloc: ") + cudaGetErrorString(e));
  double* h = (double*)(slab.h + o_in);
  memcpy(h, pos0, 3 * D); memcpy(h + 3, quat0, 4 * D); memcpy(h + 7, ba0, 3 * D); memcpy(h + 10, pos1, 3 * D);
  memcpy(h + 13, quat1, 4 * D); memcpy(h + 17, vel1, 3 * D); memcpy(h + 20, acc, 3 * D); memcpy(h + 23, gyr, 3 * D);
  e = cudaMemcpyAsync(slab.d, slab.h, in_bytes, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) {
    horizon_imu_kernel<<<1, 32, 0, ctx->stream>>>(H, (const double*)(slab.d + o_in), nr_imu, delta_imu, (double*)(slab.d + o_pos),
                                                 (double*)(slab.d + o_quat));
    ctx->launches += 1;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(slab.h + o_pos, slab.d + o_pos, cv.off - o_pos, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e == cudaSuccess) { memcpy(horizon_pos, slab.h + o_pos, (size_t)(H + 1) * 3 * D); memcpy(horizon_quat, slab.h + o_quat, (size_t)(H + 1) * 4 * D); }
  slab_release(slab, ctx->sel_cache_busy, from_cache);
  if (e != cudaSuccess) return fail(ctx, BVIO_ERR_CUDA, std::string("horizon_imu: ") + cudaGetErrorString(e));
  return BVIO_OK;
}

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
