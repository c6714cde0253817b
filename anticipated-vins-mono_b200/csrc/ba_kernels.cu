// ba_kernels.cu -- sliding-window bundle adjustment on sm_100a, IEEE double throughout (the blocks are 2x6 / 6x6 /
// 15x15 and irregular: no tcgen05 / TMEM; the one dense sub-problem, Gram products over a chunk's landmarks, runs on the
// FP64 tensor-core path mma.sync.m8n8k4.f64).  Replaces what ceres::Solve does for the problem
// Estimator::optimization() builds (vins_estimator/src/estimator.cpp:661-809):
//
//   ba_prepare        once per upload: IMU sqrt-information (imu_factor.h:64), prior J^T J
//   ba_linearize_mma  ProjectionFactor::Evaluate (projection_factor.cpp:21-121) + Cauchy corrector
//                     (marginalization_factor.cpp:37-68), one factor per thread; per-landmark J^T J / J^T r and Schur
//                     elimination of the inverse depth as DMMA Gram products; one extra CTA per window evaluates the IMU
//                     factors (imu_factor.h:19-179) and the marginalization prior (marginalization_factor.cpp:333-381)
//   ba_linearize<XB>  the same with free extrinsics and / or td (ProjectionTdFactor, projection_td_factor.cpp:34-141)
//                     as pseudo-frames of the visual layout
//   ba_solve          one CTA per window: assemble the reduced system in shared memory, LM damping (Ceres
//                     LevenbergMarquardtStrategy) or the mu-regularised Gauss-Newton step of DoglegStrategy, blocked
//                     Cholesky, back substitution
//   ba_dogleg         traditional dogleg: landmark parts of the dot products, Gauss-Newton / Cauchy / interpolated step
//   ba_cost           back-substitute the inverse depths, PoseLocalParameterization::Plus
//                     (pose_local_parameterization.cpp:3-18), residual-only cost at the candidate, and the
//                     trust-region accept / reject decision (last CTA of the window)
//   ba_marginalize    MarginalizationInfo::{preMarginalize, marginalize} (marginalization_factor.cpp:109-297)
//
// All reductions run in a fixed order: results are bit-reproducible run to run.
#include <cstdio>
#include <mutex>
#include "ba.h"
#include <float.h>

namespace bvio {

// local block-pair table: pair index -> (a, b), a >= b, for up to BVIO_KMAX observations/frames
__constant__ unsigned char c_triA[BVIO_KMAX * (BVIO_KMAX + 1) / 2];
__constant__ unsigned char c_triB[BVIO_KMAX * (BVIO_KMAX + 1) / 2];
// packed-lower decode for 30x30 IMU blocks
__constant__ unsigned char c_tri30A[465];
__constant__ unsigned char c_tri30B[465];

constexpr int FR = 12;            // per-frame staged state: R (9, row-major) + P (3)
constexpr int STG = 29;           // per-factor staging stride (28 used; odd => conflict-free)
constexpr int STGX = 41;          // same with the 2x6 extrinsic Jacobian at 28..39 (estimate_extrinsic)
constexpr int ACS = 37;           // padded stride of one 6x6 accumulator block (odd => lanes hit distinct banks)

__device__ __forceinline__ int tri(int i, int j) { return i * (i + 1) / 2 + j; }   // i >= j

// landmarks of tile t of nt: the linearization cuts a window into bt.TL tiles (tile records), the cost / dogleg kernels
// into bt.T (their own partial sums)
__device__ __forceinline__ void tile_range(const BaBatch& bt, int w, int t, int& l0, int& l1, int nt) {
  int base = bt.lm_base[w], Lw = bt.lm_base[w + 1] - base;
  int tl = (Lw + nt - 1) / nt;
  l0 = base + min(Lw, t * tl);
  l1 = base + min(Lw, (t + 1) * tl);
}

// stage R,P of the K frames of window w (state buffer `buf`) and the extrinsics into shared memory
__device__ __forceinline__ void stage_frames(const BaBatch& bt, int w, const double* pose, const double* ex,
                                             double* sFr, double* sEx) {
  for (int k = threadIdx.x; k <= bt.K; k += blockDim.x) {
    const double* p = (k < bt.K) ? pose + (size_t)(w * bt.K + k) * 7 : ex + (size_t)w * 7;
    double* o = (k < bt.K) ? sFr + k * FR : sEx;
    qmat(q4{p[3], p[4], p[5], p[6]}, o);
    o[9] = p[0]; o[10] = p[1]; o[11] = p[2];
  }
}

// visual layout (6 dims per frame, then one 6-slot per extra block) <-> reduced layout (15 per frame, then the
// extrinsic 6, then td 1)
__device__ __forceinline__ int vis2red(const BaBatch& bt, int p, int r) {
  if (p < bt.K) return 15 * p + r;
  if (bt.est_ex && p == bt.K) return 15 * bt.K + r;
  return r == 0 ? 15 * bt.K + 6 * bt.est_ex : -1;           // td lives in column 0 of its slot
}
__device__ __forceinline__ int red2vis(const BaBatch& bt, int i) {
  if (i < 15 * bt.K) { const int p = i / 15, r = i - 15 * p; return r < 6 ? 6 * p + r : -1; }
  const int j = i - 15 * bt.K;
  if (bt.est_ex && j < 6) return 6 * bt.K + j;
  return 6 * (bt.K + bt.est_ex);
}

struct ProjGeom {   // what both the residual-only and the full evaluation need
  d3 pimu_i, pimu_j, pcj;
};

// residual of one ProjectionFactor (projection_factor.cpp:35-49) with staged rotation matrices
__device__ __forceinline__ ProjGeom proj_geom(const double* Fi, const double* Fj, const double* Ex, double xi,
                                              double yi, double lam) {
  ProjGeom g;
  double il = 1.0 / lam;
  d3 pci{xi * il, yi * il, il};
  d3 tic{Ex[9], Ex[10], Ex[11]};
  g.pimu_i = mv3(Ex, pci) + tic;
  d3 pw = mv3(Fi, g.pimu_i) + d3{Fi[9], Fi[10], Fi[11]};
  g.pimu_j = mtv3(Fj, pw - d3{Fj[9], Fj[10], Fj[11]});
  g.pcj = mtv3(Ex, g.pimu_j - tic);
  return g;
}

// Diagonal regularisation of one column with Jacobi scale^2 = s2 and clamped scaled diagonal `cl`:
//   LM      (LevenbergMarquardtStrategy): D^2 = cl / radius          -> unscaled cl / (radius s2)
//   dogleg  (DoglegStrategy GN step)    : D^2 = mu * cl              -> unscaled mu cl / s2
// One division per call: the strategy-dependent factor (mu or 1/radius) is formed once per thread.
__device__ __forceinline__ double damp_factor(double radius, double mu, int dogleg) { return dogleg ? mu : 1.0 / radius; }
__device__ __forceinline__ double damp_term(double cl, double s2, double dfac) { return dfac * cl / s2; }

__device__ __forceinline__ void cauchy(double a, double s, double& rho0, double& rho1) {
  double bb = a * a, c = 1.0 / bb;
  double sum = 1.0 + s * c, inv = 1.0 / sum;
  rho0 = bb * log(sum);
  rho1 = fmax(DBL_MIN, inv);
}

// ---------------------------------------------------------------------------------------------
// IMU factor: raw residual (15) and, optionally, raw Jacobian (15 x 30, columns [pose_i sb_i pose_j sb_j]
// in local 6/9/6/9 coordinates) before whitening.  One thread per factor.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void put3(double* J, int r0, int c0, const double* m, double s) {
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) J[(r0 + i) * 30 + c0 + j] = s * m[i * 3 + j];
}
__device__ __forceinline__ void skew3(d3 v, double* m) {
  m[0] = 0; m[1] = -v.z; m[2] = v.y; m[3] = v.z; m[4] = 0; m[5] = -v.x; m[6] = -v.y; m[7] = v.x; m[8] = 0;
}
// (Qleft(a) * Qright(b)).bottomRightCorner<3,3>()  (utility.h:49-67)
__device__ __forceinline__ void qleft_qright_br(q4 a, q4 b, double* out) {
  double la[9], rb[9], sa[9], sb[9];
  skew3(d3{a.x, a.y, a.z}, sa);
  skew3(d3{b.x, b.y, b.z}, sb);
#pragma unroll
  for (int i = 0; i < 9; i++) { la[i] = sa[i]; rb[i] = -sb[i]; }
  la[0] += a.w; la[4] += a.w; la[8] += a.w;
  rb[0] += b.w; rb[4] += b.w; rb[8] += b.w;
  mm3(la, rb, out);
  const double av[3] = {a.x, a.y, a.z}, bv[3] = {b.x, b.y, b.z};
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) out[i * 3 + j] -= av[i] * bv[j];
}
__device__ __forceinline__ void qleft_br(q4 a, double* out) {
  skew3(d3{a.x, a.y, a.z}, out);
  out[0] += a.w; out[4] += a.w; out[8] += a.w;
}

__device__ void imu_raw(const double* rec, const double* G, const double* pi, const double* sbi, const double* pj,
                        const double* sbj, double* r /*[15]*/, double* J /*[15*30] zeroed, or null*/) {
  d3 Pi{pi[0], pi[1], pi[2]}, Pj{pj[0], pj[1], pj[2]};
  q4 Qi{pi[3], pi[4], pi[5], pi[6]}, Qj{pj[3], pj[4], pj[5], pj[6]};
  d3 Vi{sbi[0], sbi[1], sbi[2]}, Bai{sbi[3], sbi[4], sbi[5]}, Bgi{sbi[6], sbi[7], sbi[8]};
  d3 Vj{sbj[0], sbj[1], sbj[2]}, Baj{sbj[3], sbj[4], sbj[5]}, Bgj{sbj[6], sbj[7], sbj[8]};
  d3 g{G[0], G[1], G[2]};
  double dt = rec[IR_DT];
  d3 dba = Bai - d3{rec[IR_BA], rec[IR_BA + 1], rec[IR_BA + 2]};
  d3 dbg = Bgi - d3{rec[IR_BG], rec[IR_BG + 1], rec[IR_BG + 2]};
  q4 dq{rec[IR_DQ], rec[IR_DQ + 1], rec[IR_DQ + 2], rec[IR_DQ + 3]};
  // IntegrationBase::evaluate (integration_base.h:160-186)
  d3 th = mv3(rec + IR_DQDBG, dbg);
  q4 cdq = qmul(dq, q4{th.x / 2, th.y / 2, th.z / 2, 1.0});
  d3 cdv = d3{rec[IR_DV], rec[IR_DV + 1], rec[IR_DV + 2]} + mv3(rec + IR_DVDBA, dba) + mv3(rec + IR_DVDBG, dbg);
  d3 cdp = d3{rec[IR_DP], rec[IR_DP + 1], rec[IR_DP + 2]} + mv3(rec + IR_DPDBA, dba) + mv3(rec + IR_DPDBG, dbg);
  q4 Qii = qinv(Qi);
  d3 tp = qrot(Qii, (0.5 * dt * dt) * g + Pj - Pi - dt * Vi);
  d3 tv = qrot(Qii, dt * g + Vj - Vi);
  d3 rp = tp - cdp, rv = tv - cdv;
  q4 qe = qmul(qinv(cdq), qmul(Qii, Qj));
  r[0] = rp.x; r[1] = rp.y; r[2] = rp.z;
  r[3] = 2 * qe.x; r[4] = 2 * qe.y; r[5] = 2 * qe.z;
  r[6] = rv.x; r[7] = rv.y; r[8] = rv.z;
  r[9] = Baj.x - Bai.x; r[10] = Baj.y - Bai.y; r[11] = Baj.z - Bai.z;
  r[12] = Bgj.x - Bgi.x; r[13] = Bgj.y - Bgi.y; r[14] = Bgj.z - Bgi.z;
  if (!J) return;
  double RiT[9], m[9], m2[9];
  qmat(Qii, RiT);
  const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  // pose_i: columns 0..5
  put3(J, 0, 0, RiT, -1.0);
  skew3(tp, m); put3(J, 0, 3, m, 1.0);
  qleft_qright_br(qmul(qinv(Qj), Qi), cdq, m); put3(J, 3, 3, m, -1.0);
  skew3(tv, m); put3(J, 6, 3, m, 1.0);
  // speed-bias_i: columns 6..14
  put3(J, 0, 6, RiT, -dt);
  put3(J, 0, 9, rec + IR_DPDBA, -1.0);
  put3(J, 0, 12, rec + IR_DPDBG, -1.0);
  qleft_br(qmul(qmul(qinv(Qj), Qi), dq), m); mm3(m, rec + IR_DQDBG, m2); put3(J, 3, 12, m2, -1.0);
  put3(J, 6, 6, RiT, -1.0);
  put3(J, 6, 9, rec + IR_DVDBA, -1.0);
  put3(J, 6, 12, rec + IR_DVDBG, -1.0);
  put3(J, 9, 9, I3, -1.0);
  put3(J, 12, 12, I3, -1.0);
  // pose_j: columns 15..20
  put3(J, 0, 15, RiT, 1.0);
  qleft_br(qmul(qmul(qinv(cdq), Qii), Qj), m); put3(J, 3, 18, m, 1.0);
  // speed-bias_j: columns 21..29
  put3(J, 6, 21, RiT, 1.0);
  put3(J, 9, 24, I3, 1.0);
  put3(J, 12, 27, I3, 1.0);
}

// MarginalizationFactor::Evaluate residual: dx per kept block, r = r0 + J dx. Block-cooperative.
// sdx, sr: shared [nmax]. Returns nothing; caller syncs.
__device__ void prior_residual(const BaBatch& bt, int w, const double* pose, const double* sb, const double* ex,
                               const double* td, double* sdx, double* sr) {
  int n = bt.pr_n[w], nb = bt.pr_nb[w];
  for (int bidx = threadIdx.x; bidx < nb; bidx += blockDim.x) {
    int kind = bt.pr_kind[w * PRIOR_MAXB + bidx], fr = bt.pr_frame[w * PRIOR_MAXB + bidx];
    int idx = bt.pr_idx[w * PRIOR_MAXB + bidx];
    const double* x0 = bt.pr_x0 + (size_t)(w * PRIOR_MAXB + bidx) * 9;
    const double* x = kind == 0 ? pose + (size_t)(w * bt.K + fr) * 7
                      : kind == 1 ? sb + (size_t)(w * bt.K + fr) * 9 : ex + (size_t)w * 7;
    if (kind == 1) {
      for (int i = 0; i < 9; i++) sdx[idx + i] = x[i] - x0[i];
    } else if (kind == 3) {
      sdx[idx] = td[w] - x0[0];
    } else {
      for (int i = 0; i < 3; i++) sdx[idx + i] = x[i] - x0[i];
      q4 dq = qmul(qinv(q4{x0[3], x0[4], x0[5], x0[6]}), q4{x[3], x[4], x[5], x[6]});
      double s = (dq.w >= 0) ? 2.0 : -2.0;
      sdx[idx + 3] = s * dq.x; sdx[idx + 4] = s * dq.y; sdx[idx + 5] = s * dq.z;
    }
  }
  __syncthreads();
  const double* Jc = bt.pr_jac + (size_t)w * bt.nmax * bt.nmax;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double s = bt.pr_res[(size_t)w * bt.nmax + i];
    for (int k = 0; k < n; k++) s += Jc[(size_t)k * n + i] * sdx[k];
    sr[i] = s;
  }
  __syncthreads();
}

// =============================================================================================
// prepare
// =============================================================================================
__global__ void __launch_bounds__(BA_THREADS) ba_prepare_kernel(BaBatch bt) {
  const int w = blockIdx.x, K = bt.K;
  __shared__ double sU[8][15 * 16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // ---- compact IMU records; sqrt_info = LLT(cov^-1).matrixL().transpose() (imu_factor.h:64), computed
  //      as the inverse of the upper "reverse Cholesky" factor of cov (cov = U~ U~^T  =>
  //      cov^-1 = U~^-T U~^-1, and U~^-1 is upper triangular with positive diagonal = L^T by uniqueness).
  for (int j = 1 + warp; j < K; j += 8) {
    const double* raw = bt.preint_raw + (size_t)(w * K + j) * PREINT_DOUBLES;
    double* rec = bt.imu + (size_t)(w * K + j) * IMU_REC;
    for (int i = lane; i < 17; i += 32) rec[i] = raw[i];   // delta_p,q,v, lin_ba, lin_bg, sum_dt
    const double* Jb = raw + 17;
    for (int i = lane; i < 45; i += 32) {
      int blk = i / 9, e = i - blk * 9, r = e / 3, c = e - r * 3;
      const int r0[5] = {0, 0, 3, 6, 6}, c0[5] = {9, 12, 12, 9, 12};
      rec[IR_DPDBA + i] = Jb[(r0[blk] + r) * 15 + c0[blk] + c];
    }
    const double* cov = raw + 17 + 225;
    double* U = sU[warp];   // 15 x 16 (padded)
    for (int i = lane; i < 225; i += 32) U[(i / 15) * 16 + (i % 15)] = cov[i];
    __syncwarp();
    for (int c = 14; c >= 0; c--) {
      double d = U[c * 16 + c];
      for (int k = c + 1; k < 15; k++) d -= U[c * 16 + k] * U[c * 16 + k];
      d = sqrt(d);
      double v = 0;
      if (lane < c) {
        v = U[lane * 16 + c];
        for (int k = c + 1; k < 15; k++) v -= U[lane * 16 + k] * U[c * 16 + k];
        v /= d;
      }
      __syncwarp();
      if (lane < c) U[lane * 16 + c] = v;
      if (lane == c) U[c * 16 + c] = d;
      __syncwarp();
    }
    // X = U~^-1 (upper): lane = column
    if (lane < 15) {
      int c = lane;
      double x[15];
#pragma unroll
      for (int i = 0; i < 15; i++) x[i] = 0;
#pragma unroll
      for (int i = 14; i >= 0; i--) {
        if (i == c) x[i] = 1.0 / U[i * 16 + i];
        else if (i < c) {
          double s = 0;
#pragma unroll
          for (int k = i + 1; k < 15; k++) if (k <= c) s += U[i * 16 + k] * x[k];
          x[i] = -s / U[i * 16 + i];
        }
      }
#pragma unroll
      for (int i = 0; i < 15; i++) rec[IR_SQ + i * 15 + c] = x[i];
    }
    __syncwarp();
  }
  // ---- prior: column map and H = J^T J
  int n = bt.pr_n[w], nb = bt.pr_nb[w];
  int* map = bt.pr_map + (size_t)w * bt.nmax;
  for (int i = threadIdx.x; i < bt.nmax; i += blockDim.x) map[i] = -1;
  __syncthreads();
  for (int bidx = threadIdx.x; bidx < nb; bidx += blockDim.x) {
    int kind = bt.pr_kind[w * PRIOR_MAXB + bidx], fr = bt.pr_frame[w * PRIOR_MAXB + bidx];
    int idx = bt.pr_idx[w * PRIOR_MAXB + bidx];
    int loc = kind == 1 ? 9 : (kind == 3 ? 1 : 6);
    int base = kind == 0 ? 15 * fr : (kind == 1 ? 15 * fr + 6 : ((kind == 2 && bt.est_ex) ? 15 * K
               : ((kind == 3 && bt.est_td) ? 15 * K + 6 * bt.est_ex : -1)));   // constant blocks drop out
    for (int i = 0; i < loc; i++) map[idx + i] = base < 0 ? -1 : base + i;
  }
  const double* Jc = bt.pr_jac + (size_t)w * bt.nmax * bt.nmax;
  double* H = bt.pr_H + (size_t)w * bt.nmax * bt.nmax;
  // H = J^T J through shared-memory tiles: 16 rows (k) of all n columns at a time (coalesced reads of the column-major J),
  // eight outputs per thread per pass -- the prior's Hessian is on the critical path of every one-window call
  {
    constexpr int TK = 16, ACC = 8;
    __shared__ double sJt[TK * 97];                        // [TK][n] for n <= 96; wider priors take the direct loop
    if (n <= 96) {
      const int npass = (n * n + ACC * (int)blockDim.x - 1) / (ACC * (int)blockDim.x);
      for (int pass = 0; pass < npass; pass++) {
        double acc[ACC];
        int ea[ACC], eb[ACC];
#pragma unroll
        for (int u = 0; u < ACC; u++) {
          const int e = (pass * ACC + u) * blockDim.x + threadIdx.x;
          acc[u] = 0.0;
          ea[u] = e < n * n ? e / n : -1;
          eb[u] = e < n * n ? e - (e / n) * n : 0;
        }
        for (int k0 = 0; k0 < n; k0 += TK) {
          __syncthreads();
          for (int i = threadIdx.x; i < TK * n; i += blockDim.x) {
            const int c = i / TK, kk = i - c * TK;
            sJt[kk * 97 + c] = (k0 + kk < n) ? Jc[(size_t)c * n + k0 + kk] : 0.0;
          }
          __syncthreads();
#pragma unroll
          for (int u = 0; u < ACC; u++) {
            if (ea[u] < 0) continue;
#pragma unroll
            for (int kk = 0; kk < TK; kk++) acc[u] = fma(sJt[kk * 97 + ea[u]], sJt[kk * 97 + eb[u]], acc[u]);
          }
        }
#pragma unroll
        for (int u = 0; u < ACC; u++) if (ea[u] >= 0) H[(size_t)ea[u] * bt.nmax + eb[u]] = acc[u];
      }
    } else {
      for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
        int a = e / n, b2 = e - a * n;
        double s = 0;
        for (int k = 0; k < n; k++) s += Jc[(size_t)a * n + k] * Jc[(size_t)b2 * n + k];
        H[(size_t)a * bt.nmax + b2] = s;
      }
    }
  }
  __syncthreads();
  // reduced index -> prior column, and (latency mode) the prior Hessian scattered once into ba_solve's packed layout
  int* inv = bt.pr_inv + (size_t)w * bt.np;
  for (int i = threadIdx.x; i < bt.np; i += blockDim.x) inv[i] = -1;
  __syncthreads();
  for (int a = threadIdx.x; a < n; a += blockDim.x) if (map[a] >= 0) inv[map[a]] = a;
  if (bt.S0) {
    const int N1 = bt.np + 1;
    double* S0 = bt.S0 + (size_t)w * (N1 * (N1 + 1) / 2);
    for (int i = threadIdx.x; i < N1 * (N1 + 1) / 2; i += blockDim.x) S0[i] = 0.0;
    __syncthreads();
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
      int a = e / n, b2 = e - a * n;
      int ra = map[a], rb = map[b2];
      if (ra < 0 || rb < 0 || ra < rb) continue;
      S0[tri(ra, rb)] = H[(size_t)a * bt.nmax + b2];
    }
  }
}

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

__global__ void ba_reset_kernel(BaBatch bt) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (size_t k = i; k < (size_t)bt.B * bt.K * 7; k += stride) bt.pose[0][k] = bt.pose0[k];
  for (size_t k = i; k < (size_t)bt.B * bt.K * 9; k += stride) bt.sb[0][k] = bt.sb0[k];
  for (size_t k = i; k < (size_t)bt.total_L; k += stride) bt.invd[0][k] = bt.invd0[k];
  for (size_t k = i; k < (size_t)bt.B * 7; k += stride) { bt.exs[0][k] = bt.ex[k]; bt.exs[1][k] = bt.ex[k]; }
  for (size_t k = i; k < (size_t)bt.B; k += stride) { bt.tds[0][k] = bt.td0[k]; bt.tds[1][k] = bt.td0[k]; }
  for (size_t k = i; k < (size_t)bt.B; k += stride) {
    BaCtrl c;
    c.cost = 0; c.cand_cost = 0; c.radius = bt.initial_radius; c.decrease_factor = 2.0;
    c.model_pose = 0; c.gmax = 0; c.initial_cost = 0; c.rho = 0;
    c.cur = 0; c.done = 0; c.termination = 0 /*MAX_ITERS*/; c.iterations = 0; c.accepted = 0; c.rejected = 0;
    c.first = 1; c.solve_ok = 0; c.invalid_run = 0; c.stepped = 0; c.ticket = 0; c.pad = 0;
    c.mu = 1e-8; c.ca = 0; c.cb = 1; c.step_norm = 0;
    for (int q = 0; q < 6; q++) c.dsum[q] = 0;
    c.t0_ns = global_ns();
    bt.ctrl[k] = c;
  }
}

__global__ void ba_finish_kernel(BaBatch bt) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  const size_t np7 = (size_t)bt.K * 7, np9 = (size_t)bt.K * 9;
  for (size_t k = i; k < (size_t)bt.B * np7; k += stride) bt.pose_out[k] = bt.pose[bt.ctrl[k / np7].cur][k];
  for (size_t k = i; k < (size_t)bt.B * np9; k += stride) bt.sb_out[k] = bt.sb[bt.ctrl[k / np9].cur][k];
  for (size_t k = i; k < (size_t)bt.B * 7; k += stride) bt.ex_out[k] = bt.exs[bt.ctrl[k / 7].cur][k];
  for (size_t k = i; k < (size_t)bt.B; k += stride) bt.td_out[k] = bt.tds[bt.ctrl[k].cur][k];
  for (int w = blockIdx.x; w < bt.B; w += gridDim.x) {
    int cur = bt.ctrl[w].cur;
    for (int l = bt.lm_base[w] + threadIdx.x; l < bt.lm_base[w + 1]; l += blockDim.x) bt.invd_out[l] = bt.invd[cur][l];
  }
}

// =============================================================================================
// linearize: visual tiles (blockIdx.x < T) + IMU/prior CTA (blockIdx.x == T)
// =============================================================================================
// f0 .. f1: which IMU factors (0-based: factor f links frame f to f + 1) this CTA evaluates; do_prior: also the prior.
// One CTA does everything for large batches; with fewer windows than SMs (latency mode) every factor gets its own CTA.
__device__ void imu_prior_linearize(const BaBatch& bt, int w, int cur, double* sm, int f0, int f1, bool do_prior, bool compact = false) {
  // compact: shared memory holds only the factors f0 .. f1-1 (slot = f - f0) instead of all K - 1
  const int K = bt.K, nf = compact ? f1 - f0 : K - 1, fo = compact ? f0 : 0;
  double* sJ = sm - (size_t)fo * 450;   // [nf][450] raw Jacobians (indexed by the factor number)
  double* sJ2 = sJ + nf * 450;          // [nf][450] whitened
  double* sR = sm + (size_t)nf * 900 - (size_t)fo * 15;   // [nf][15] raw
  double* sR2 = sR + nf * 15;           // [nf][15] whitened
  double* sdx = sm + (size_t)nf * 930;  // [nmax]
  double* spr = sdx + bt.nmax;          // [nmax]
  const double* pose = bt.pose[cur];
  const double* sb = bt.sb[cur];
  const int nfl = f1 - f0;
  for (int i = threadIdx.x; i < nfl * 450; i += blockDim.x) sJ[f0 * 450 + i] = 0.0;
  __syncthreads();
  if ((int)threadIdx.x < nfl) {
    int f = f0 + threadIdx.x, j = f + 1;
    const double* rec = bt.imu + (size_t)(w * K + j) * IMU_REC;
    if (!(rec[IR_DT] > 10.0))
      imu_raw(rec, bt.G, pose + (size_t)(w * K + j - 1) * 7, sb + (size_t)(w * K + j - 1) * 9,
              pose + (size_t)(w * K + j) * 7, sb + (size_t)(w * K + j) * 9, sR + f * 15, sJ + f * 450);
    else
      for (int i = 0; i < 15; i++) sR[f * 15 + i] = 0.0;
  }
  __syncthreads();
  // whiten: J2 = SI * J, r2 = SI * r (SI upper triangular)
  for (int e = threadIdx.x; e < nfl * 465; e += blockDim.x) {
    int f = f0 + e / 465, o = e - (e / 465) * 465;
    const double* SI = bt.imu + (size_t)(w * K + f + 1) * IMU_REC + IR_SQ;
    int i = o / 31, c = o - i * 31;
    double s = 0;
    if (c < 30) {
      for (int k = i; k < 15; k++) s += SI[i * 15 + k] * sJ[f * 450 + k * 30 + c];
      sJ2[f * 450 + i * 30 + c] = s;
    } else {
      for (int k = i; k < 15; k++) s += SI[i * 15 + k] * sR[f * 15 + k];
      sR2[f * 15 + i] = s;
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < nfl * IMU_OUT; e += blockDim.x) {
    int f = f0 + e / IMU_OUT, o = e - (e / IMU_OUT) * IMU_OUT;
    const double* J = sJ2 + f * 450;
    const double* r = sR2 + f * 15;
    double s = 0;
    if (o < 465) {
      int a = c_tri30A[o], b2 = c_tri30B[o];
      for (int k = 0; k < 15; k++) s += J[k * 30 + a] * J[k * 30 + b2];
    } else if (o < 495) {
      int a = o - 465;
      for (int k = 0; k < 15; k++) s += J[k * 30 + a] * r[k];
    } else {
      for (int k = 0; k < 15; k++) s += r[k] * r[k];
      s *= 0.5;
    }
    bt.imu_out[(size_t)(w * K + f + 1) * IMU_OUT + o] = s;
  }
  if (!do_prior) return;
  // prior
  int n = bt.pr_n[w];
  double* po = bt.pr_out + (size_t)w * (bt.nmax + 1);
  if (n > 0) {
    prior_residual(bt, w, pose, sb, bt.exs[cur], bt.tds[cur], sdx, spr);
    const double* Jc = bt.pr_jac + (size_t)w * bt.nmax * bt.nmax;
    for (int a = threadIdx.x; a < n; a += blockDim.x) {
      double s = 0;
      for (int k = 0; k < n; k++) s += Jc[(size_t)a * n + k] * spr[k];
      po[a] = s;
    }
    if (threadIdx.x == 0) {
      double c = 0;
      for (int k = 0; k < n; k++) c += spr[k] * spr[k];
      po[bt.nmax] = 0.5 * c;
    }
  } else if (threadIdx.x == 0) {
    po[bt.nmax] = 0.0;
  }
}

// Visual part, chunked: the CTA walks its landmark tile in chunks of CL landmarks.
//   A1  thread = (landmark, factor) slot: residual + Jacobians + Cauchy -> 28 doubles per factor in smem
//   A2  thread = (landmark, output): anchor reductions A^T A, A^T c, A^T r, h, b
//   A3  thread = landmark: LM damping of the eliminated depth, h/b/w to HBM
//   B   thread = owner of fixed 1x6 strips of the 6K x 6K block matrix, accumulated in REGISTERS over the
//       chunk's landmarks (fixed order => deterministic; no shared-memory accumulators, no atomics)
// XB extra parameter blocks ride as pseudo-frames K .. K+XB-1 of the visual layout: the extrinsic pose (estimate_extrinsic,
// estimator.cpp:672-683; Jacobian projection_factor.cpp:97-106) and/or the camera-IMU time offset td (estimate_td,
// estimator.cpp:732-740; ProjectionTdFactor, projection_td_factor.cpp:34-141; a 2x1 Jacobian carried in column 0 of
// a 2x6 slot).  Every landmark "observes" pseudo-frame x with w_x = sum_f X_f^T c_f, and its J^T J terms are
// (x,x) = sum_f X_f^T X_f, (x,anchor) = sum_f X_f^T A_f, (x,frame of f) = X_f^T B_f, (td,ex) = sum_f T_f^T E_f.
template <int NS, int XB>
__global__ void __launch_bounds__(BA_THREADS, 2) ba_linearize_kernel(BaBatch bt) {
  extern __shared__ double sm[];
  const int w = blockIdx.y, t = blockIdx.x;
  BaCtrl* ctrl = bt.ctrl + w;
  if (ctrl->done) return;
  const int cur = ctrl->cur;
  if (t >= bt.TL) {   // IMU / prior CTAs: one for all, or (latency mode) the prior CTA followed by one CTA per IMU factor
    if (gridDim.x == (unsigned)bt.TL + 1) imu_prior_linearize(bt, w, cur, sm, 0, bt.K - 1, true);
    else if (t == bt.TL) imu_prior_linearize(bt, w, cur, sm, 0, 0, true);
    else imu_prior_linearize(bt, w, cur, sm, t - bt.TL - 1, t - bt.TL, false);
    return;
  }

  constexpr int SG = STG + 12 * XB;                   // per-factor staging stride (29 / 41 / 53: odd)
  constexpr int NRED = 35 + 69 * XB + (XB == 2 ? 36 : 0);   // per-landmark reductions over the factors
  constexpr bool EX = XB > 0;
  const int K = bt.K, KE = K + XB, K6 = 6 * KE, NPb = KE * (KE + 1) / 2, FS = K - 1, CL = bt.chunk_l;
  const int x_ex = bt.est_ex ? 0 : -1, x_td = bt.est_td ? bt.est_ex : -1;   // pseudo-frame index of each extra block
  const int tid = threadIdx.x;
  double* sFr = sm;                                  // [K*FR]
  double* sEx = sFr + K * FR;                        // [FR]
  double* sRed = sEx + FR;                           // [16]
  double* sFac = sRed + 16;                          // [CL*FS*SG]
  double* sW = sFac + (size_t)CL * FS * SG;          // [CL*K6]
  double* sAtA = sW + (size_t)CL * K6;               // [CL*36]
  double* sgA = sAtA + (size_t)CL * 36;              // [CL*6]
  double* sSc = sgA + (size_t)CL * 6;                // [CL*4] inv_hd, b, h, -
  double* sEtE = sSc + (size_t)CL * 4;               // [XB][CL*36] sum X^T X, then [XB][CL*36] sum X^T A,
  double* sEtA = sEtE + (size_t)XB * CL * 36;        // [XB][CL*6] sum X^T r, and (XB == 2) [CL*36] sum X_1^T X_0
  double* sgE = sEtA + (size_t)XB * CL * 36;
  double* sXX = sgE + (size_t)XB * CL * 6;
  int* sMeta = reinterpret_cast<int*>(sXX + (XB == 2 ? (size_t)CL * 36 : 0));  // [CL*2] o0, n
  signed char* sOidx = reinterpret_cast<signed char*>(sMeta + CL * 2);  // [CL*KE] frame -> observation index

  // Units this thread owns: a unit is half a 6x6 block (3 rows x 6 columns = 18 register accumulators).
  // Block order: the KE diagonal blocks first, then the off-diagonal blocks (p > q) sorted by column q, so
  // that "q is this landmark's anchor frame" is (nearly) warp-uniform.  The last warp owns the three
  // gradient vectors instead of units.
  const int NUNIT = NPb * 2, TS = (NUNIT + NS - 1) / NS;         // TS <= BA_THREADS - 32 by choice of NS
  int s_p[NS], s_q[NS], s_r[NS];
  double acc[NS][3][6];
#pragma unroll
  for (int u = 0; u < NS; u++) {
    const int uidx = tid + TS * u;
    s_p[u] = -1; s_q[u] = -1; s_r[u] = 0;
    if (tid < TS && uidx < NUNIT) {
      const int blk = uidx >> 1;
      s_r[u] = 3 * (uidx & 1);
      if (blk < KE) { s_p[u] = blk; s_q[u] = blk; }
      else {
        int rem = blk - KE, q = 0;
        while (rem >= KE - 1 - q) { rem -= KE - 1 - q; q++; }
        s_q[u] = q; s_p[u] = q + 1 + rem;
      }
    }
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int c = 0; c < 6; c++) acc[u][i][c] = 0.0;
  }
  const bool gwarp = tid >= BA_THREADS - 32;          // gradient elements lane, lane+32, lane+64 (< 6KE <= 96)
  const int glane = tid - (BA_THREADS - 32);
  double g_red[3] = {0, 0, 0}, g_bp[3] = {0, 0, 0}, g_dg[3] = {0, 0, 0};
  double cost_t = 0, gmax_t = 0;

  stage_frames(bt, w, bt.pose[cur], bt.exs[cur], sFr, sEx);
  const double dfac = damp_factor(ctrl->radius, ctrl->mu, bt.strategy);
  const int first = ctrl->first;
  const double* invd = bt.invd[cur];
  int l0, l1;
  tile_range(bt, w, t, l0, l1, bt.TL);

  for (int lb = l0; lb < l1; lb += CL) {
    const int nl = min(CL, l1 - lb);
    __syncthreads();                                 // previous chunk fully consumed (and frames staged)
    for (int i = tid; i < nl * KE; i += BA_THREADS) sOidx[i] = (EX && (i % KE) >= K) ? (signed char)(i % KE) : (signed char)-1;
    __syncthreads();
    // ---- A1: factor evaluation
    {
      const int lc = tid / FS, f = tid - lc * FS;
      if (lc < nl) {
        const int l = lb + lc;
        const int o0 = bt.lm_off[l], n = bt.lm_off[l + 1] - o0;
        const int fi = bt.obs_frame[o0];
        if (f == 0) { sMeta[lc * 2] = o0; sMeta[lc * 2 + 1] = n; sOidx[lc * KE + fi] = 0; }
        if (f < n - 1) {
          const double lam = invd[l];
          double2 pi = bt.obs_xy[o0], pj = bt.obs_xy[o0 + 1 + f];
          double2 vi = {0, 0}, vj = {0, 0};
          if (XB > 0 && x_td >= 0) {   // pts_td = pts - (td - td_obs + TR/ROW row) velocity (projection_td_factor.cpp:51-52)
            vi = bt.obs_vel[o0]; vj = bt.obs_vel[o0 + 1 + f];
            const double td = bt.tds[cur][w], si_ = td + bt.obs_shift[o0], sj_ = td + bt.obs_shift[o0 + 1 + f];
            pi.x -= si_ * vi.x; pi.y -= si_ * vi.y; pj.x -= sj_ * vj.x; pj.y -= sj_ * vj.y;
          }
          const int fj = bt.obs_frame[o0 + 1 + f];
          sOidx[lc * KE + fj] = (signed char)(f + 1);
          const double* Fi = sFr + fi * FR;
          const double* Fj = sFr + fj * FR;
          ProjGeom g = proj_geom(Fi, Fj, sEx, pi.x, pi.y, lam);
          const double inv = 1.0 / g.pcj.z, si = bt.sqrt_info;
          const double r0 = si * (g.pcj.x * inv - pj.x), r1 = si * (g.pcj.y * inv - pj.y);
          const double red[2][3] = {{si * inv, 0.0, -si * g.pcj.x * inv * inv}, {0.0, si * inv, -si * g.pcj.y * inv * inv}};
          double Gm[2][3], Q[2][3];
#pragma unroll
          for (int a = 0; a < 2; a++)
#pragma unroll
            for (int c = 0; c < 3; c++)
              Gm[a][c] = red[a][0] * sEx[c * 3 + 0] + red[a][1] * sEx[c * 3 + 1] + red[a][2] * sEx[c * 3 + 2];
#pragma unroll
          for (int a = 0; a < 2; a++)
#pragma unroll
            for (int c = 0; c < 3; c++) Q[a][c] = Gm[a][0] * Fj[c * 3 + 0] + Gm[a][1] * Fj[c * 3 + 1] + Gm[a][2] * Fj[c * 3 + 2];
          double rho0, rho1;
          cauchy(bt.cauchy_a, r0 * r0 + r1 * r1, rho0, rho1);
          cost_t += 0.5 * rho0;
          const double sr = sqrt(rho1);
          double* st = sFac + (size_t)(lc * FS + f) * SG;
          const d3 dimu = g.pimu_i - d3{sEx[9], sEx[10], sEx[11]};
          double cc[2];
#pragma unroll
          for (int a = 0; a < 2; a++) {
            const d3 u = mtv3(Fi, d3{Q[a][0], Q[a][1], Q[a][2]});              // Ri^T Q[a]^T
            const d3 jr = cross3(g.pimu_i, u);                                 // -(Q Ri [pts_imu_i]x) row
            const d3 jjr = cross3(d3{Gm[a][0], Gm[a][1], Gm[a][2]}, g.pimu_j); // (Gm [pts_imu_j]x) row
            st[a * 6 + 0] = sr * Q[a][0]; st[a * 6 + 1] = sr * Q[a][1]; st[a * 6 + 2] = sr * Q[a][2];
            st[a * 6 + 3] = sr * jr.x; st[a * 6 + 4] = sr * jr.y; st[a * 6 + 5] = sr * jr.z;
            st[12 + a * 6 + 0] = -sr * Q[a][0]; st[12 + a * 6 + 1] = -sr * Q[a][1]; st[12 + a * 6 + 2] = -sr * Q[a][2];
            st[12 + a * 6 + 3] = sr * jjr.x; st[12 + a * 6 + 4] = sr * jjr.y; st[12 + a * 6 + 5] = sr * jjr.z;
            cc[a] = sr * (-dot3(u, dimu) / lam);
          }
          st[24] = cc[0]; st[25] = cc[1]; st[26] = sr * r0; st[27] = sr * r1;
          if (XB > 0) {
#pragma unroll
            for (int k = 28; k < SG - 1; k++) st[k] = 0.0;
          }
          if (XB > 0 && x_td >= 0) {
            // td Jacobian (projection_td_factor.cpp:131-136): -Q Ri ric velocity_i / lambda + sqrt_info velocity_j
            const d3 rv = mv3(sEx, d3{vi.x, vi.y, 0.0});
#pragma unroll
            for (int a = 0; a < 2; a++) {
              const d3 u = mtv3(Fi, d3{Q[a][0], Q[a][1], Q[a][2]});
              st[28 + 12 * x_td + 6 * a] = sr * (-dot3(u, rv) / lam + si * (a == 0 ? vj.x : vj.y));
            }
          }
          if (XB > 0 && x_ex >= 0) {
            // extrinsic Jacobian: reduce * [ ric^T (Rj^T Ri - I) | -tmp_r [pci]x + [tmp_r pci]x + [tvec]x ]
            // tmp_r = ric^T Rj^T Ri ric ; tvec = ric^T (Rj^T (Ri tic + Pi - Pj) - tic)  (projection_factor.cpp:100-105)
            const d3 tic{sEx[9], sEx[10], sEx[11]};
            const d3 pci = (1.0 / lam) * d3{pi.x, pi.y, 1.0};
            double RjTRi[9], T1[9], tmp_r[9];
            mtm3(Fj, Fi, RjTRi);
            mtm3(sEx, RjTRi, T1);
            mm3(T1, sEx, tmp_r);
            const d3 inner = mtv3(Fj, mv3(Fi, tic) + d3{Fi[9], Fi[10], Fi[11]} - d3{Fj[9], Fj[10], Fj[11]}) - tic;
            const d3 tvec = mtv3(sEx, inner);
            const d3 trp = mv3(tmp_r, pci);
#pragma unroll
            for (int a = 0; a < 2; a++) {
              const double rd[3] = {red[a][0], red[a][1], red[a][2]};
#pragma unroll
              for (int c = 0; c < 3; c++) {
                double sacc = 0;
#pragma unroll
                for (int k = 0; k < 3; k++) sacc += rd[k] * (T1[k * 3 + c] - sEx[c * 3 + k]);
                st[28 + a * 6 + c] = sr * sacc;
              }
              const d3 rowv{rd[0], rd[1], rd[2]};
              const d3 e1 = cross3(mtv3(tmp_r, rowv), pci);
              const d3 e2 = cross3(rowv, trp);
              const d3 e3 = cross3(rowv, tvec);
              st[28 + a * 6 + 3] = sr * (-e1.x + e2.x + e3.x);
              st[28 + a * 6 + 4] = sr * (-e1.y + e2.y + e3.y);
              st[28 + a * 6 + 5] = sr * (-e1.z + e2.z + e3.z);
            }
          }
          double* wo = sW + (size_t)lc * K6 + (f + 1) * 6;
          double* wg = bt.w + (size_t)(o0 + 1 + f) * 6;
#pragma unroll
          for (int k = 0; k < 6; k++) { const double v = st[12 + k] * cc[0] + st[18 + k] * cc[1]; wo[k] = v; wg[k] = v; }
        }
      }
    }
    __syncthreads();
    // ---- A2: reductions over the landmark's factors, all of the form sum_f st[p]*st[q] + st[p2]*st[q2]:
    //      A^T A lower triangle 21, A^T c 6, A^T r 6, h, b;  EX: E^T E lower 21, E^T A 36, E^T c 6, E^T r 6
    for (int task = tid; task < nl * NRED; task += BA_THREADS) {
      const int lc = task / NRED, o = task - lc * NRED;
      int p, q, p2, q2;
      if (o < 21) { p = c_triA[o]; q = c_triB[o]; p2 = p + 6; q2 = q + 6; }
      else if (o < 27) { p = o - 21; q = 24; p2 = p + 6; q2 = 25; }
      else if (o < 33) { p = o - 27; q = 26; p2 = p + 6; q2 = 27; }
      else if (o == 33) { p = 24; q = 24; p2 = 25; q2 = 25; }
      else if (o == 34) { p = 24; q = 26; p2 = 25; q2 = 27; }
      else if (o < 35 + 69 * XB) {
        const int x = (o - 35) / 69, r = (o - 35) - 69 * x, bx = 28 + 12 * x;
        if (r < 21) { p = bx + c_triA[r]; q = bx + c_triB[r]; p2 = p + 6; q2 = q + 6; }                 // X^T X
        else if (r < 57) { const int i = (r - 21) / 6; p = bx + i; q = (r - 21) - 6 * i; p2 = p + 6; q2 = q + 6; }   // X^T A
        else if (r < 63) { p = bx + (r - 57); q = 24; p2 = p + 6; q2 = 25; }                             // X^T c
        else { p = bx + (r - 63); q = 26; p2 = p + 6; q2 = 27; }                                         // X^T r
      } else { const int i = (o - 35 - 138) / 6; p = 40 + i; q = 28 + (o - 35 - 138) - 6 * i; p2 = p + 6; q2 = q + 6; }   // X_1^T X_0
      const int nf = sMeta[lc * 2 + 1] - 1;
      const double* st = sFac + (size_t)lc * FS * SG;
      double s0 = 0, s1 = 0;
      int f = 0;
      for (; f + 1 < nf; f += 2, st += 2 * SG) {
        s0 += st[p] * st[q] + st[p2] * st[q2];
        s1 += st[SG + p] * st[SG + q] + st[SG + p2] * st[SG + q2];
      }
      if (f < nf) s0 += st[p] * st[q] + st[p2] * st[q2];
      const double sv = s0 + s1;
      if (o < 21) { sAtA[lc * 36 + p * 6 + q] = sv; sAtA[lc * 36 + q * 6 + p] = sv; }
      else if (o < 27) { sW[(size_t)lc * K6 + (o - 21)] = sv; bt.w[(size_t)sMeta[lc * 2] * 6 + (o - 21)] = sv; }
      else if (o < 33) sgA[lc * 6 + (o - 27)] = sv;
      else if (o == 33) sSc[lc * 4 + 2] = sv;
      else if (o == 34) sSc[lc * 4 + 1] = sv;
      else if (o < 35 + 69 * XB) {
        const int x = (o - 35) / 69, r = (o - 35) - 69 * x, bx = 28 + 12 * x;
        if (r < 21) { const int a = p - bx, b = q - bx; sEtE[(x * CL + lc) * 36 + a * 6 + b] = sv; sEtE[(x * CL + lc) * 36 + b * 6 + a] = sv; }
        else if (r < 57) sEtA[(x * CL + lc) * 36 + (r - 21)] = sv;
        else if (r < 63) { sW[(size_t)lc * K6 + 6 * (K + x) + (r - 57)] = sv; bt.wex[((size_t)(lb + lc) * XB + x) * 6 + (r - 57)] = sv; }
        else sgE[(x * CL + lc) * 6 + (r - 63)] = sv;
      } else sXX[lc * 36 + (o - 35 - 138)] = sv;
    }
    __syncthreads();
    // ---- A3: damping of the eliminated block (Ceres LevenbergMarquardtStrategy, Jacobi scaling), outputs
    if (tid < nl) {
      const int l = lb + tid;
      const double h = sSc[tid * 4 + 2], b = sSc[tid * 4 + 1];
      double sl2 = 1.0;
      if (bt.jacobi_scaling) {
        if (first) { const double q = 1.0 / (1.0 + sqrt(h)); sl2 = q * q; }
        else sl2 = bt.sl2[l];
      }
      const double ddl = damp_term(fmin(fmax(sl2 * h, 1e-6), 1e32), sl2, dfac);
      double inv_hd = 1.0 / (h + ddl);
      if (bt.undamped) inv_hd = (h > 0) ? 1.0 / h : 0.0;
      sSc[tid * 4] = inv_hd;
      gmax_t = fmax(gmax_t, fabs(b));
      bt.h[l] = h; bt.b[l] = b;
      if (first) bt.sl2[l] = sl2;
    }
    __syncthreads();
    // ---- B: register accumulation of sum_l (J^T J - w w^T / (h + d))
    for (int lc = 0; lc < nl; lc++) {
      const double inv = sSc[lc * 4], bl = sSc[lc * 4 + 1];
      const signed char* oi = sOidx + lc * KE;
      const double* W = sW + (size_t)lc * K6;
      const double* F = sFac + (size_t)lc * FS * SG;
#pragma unroll
      for (int u = 0; u < NS; u++) {
        if (s_p[u] < 0) continue;
        const int a = oi[s_p[u]], b = oi[s_q[u]];
        if (a < 0 || b < 0) continue;
        const int r0 = s_r[u];
        const double* wa = W + a * 6 + r0;
        const double* wb = W + b * 6;
        double war[3], wbv[6];
#pragma unroll
        for (int i = 0; i < 3; i++) war[i] = -inv * wa[i];
#pragma unroll
        for (int c = 0; c < 6; c++) wbv[c] = wb[c];
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
          for (int c = 0; c < 6; c++) acc[u][i][c] = fma(war[i], wbv[c], acc[u][i][c]);
        if (EX && a >= K) {
          const int x = a - K;
          if (b >= K || b == 0) {
            // sum_f X^T X (diagonal), sum_f X_1^T X_0 (td against extrinsics), sum_f X^T A (against the anchor)
            const double* at = (b == a ? sEtE + (x * CL + lc) * 36 : (b >= K ? sXX + lc * 36 : sEtA + (x * CL + lc) * 36)) + r0 * 6;
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
              for (int c = 0; c < 6; c++) acc[u][i][c] += at[i * 6 + c];
          } else {
            const double* st = F + (b - 1) * SG;                             // X_f^T B_f, f = b - 1
            const int bx = 28 + 12 * x;
            double x0[3], x1[3], y0[6], y1[6];
#pragma unroll
            for (int i = 0; i < 3; i++) { x0[i] = st[bx + r0 + i]; x1[i] = st[bx + 6 + r0 + i]; }
#pragma unroll
            for (int c = 0; c < 6; c++) { y0[c] = st[12 + c]; y1[c] = st[18 + c]; }
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
              for (int c = 0; c < 6; c++) acc[u][i][c] += x0[i] * y0[c] + x1[i] * y1[c];
          }
        } else if (b == 0 && a == 0) {
          const double* at = sAtA + lc * 36 + r0 * 6;      // sum_f A_f^T A_f
#pragma unroll
          for (int i = 0; i < 3; i++)
#pragma unroll
            for (int c = 0; c < 6; c++) acc[u][i][c] += at[i * 6 + c];
        } else if (b == 0 || a == b) {
          // (a,0) -> B_a^T A_a ; (a,a) -> B_a^T B_a
          const double* st = F + (a - 1) * SG;
          const double* sy = (b == 0) ? st : st + 12;
          double x0[3], x1[3], y0[6], y1[6];
#pragma unroll
          for (int i = 0; i < 3; i++) { x0[i] = st[12 + r0 + i]; x1[i] = st[18 + r0 + i]; }
#pragma unroll
          for (int c = 0; c < 6; c++) { y0[c] = sy[c]; y1[c] = sy[6 + c]; }
#pragma unroll
          for (int i = 0; i < 3; i++)
#pragma unroll
            for (int c = 0; c < 6; c++) acc[u][i][c] += x0[i] * y0[c] + x1[i] * y1[c];
        }
      }
      if (gwarp) {
#pragma unroll
        for (int v = 0; v < 3; v++) {
          const int gi = glane + 32 * v;
          if (gi >= K6) continue;
          const int gp = gi / 6, gr = gi - gp * 6;
          const int a = oi[gp];
          if (a < 0) continue;
          double gv, dv;
          if (EX && a >= K) { gv = sgE[((a - K) * CL + lc) * 6 + gr]; dv = sEtE[((a - K) * CL + lc) * 36 + gr * 7]; }
          else if (a == 0) { gv = sgA[lc * 6 + gr]; dv = sAtA[lc * 36 + gr * 7]; }
          else {
            const double* st = F + (a - 1) * SG;
            gv = st[12 + gr] * st[26] + st[18 + gr] * st[27];
            dv = st[12 + gr] * st[12 + gr] + st[18 + gr] * st[18 + gr];
          }
          g_bp[v] += gv;
          g_red[v] += gv - W[a * 6 + gr] * bl * inv;
          g_dg[v] += dv;
        }
      }
    }
  }
  // ---- one tile record to HBM
  double* out = bt.tile_out + (size_t)(w * bt.TL + t) * tile_rec_doubles(KE);
#pragma unroll
  for (int u = 0; u < NS; u++) {
    if (s_p[u] < 0) continue;
    double* o = out + (size_t)tri(s_p[u], s_q[u]) * 36 + s_r[u] * 6;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int c = 0; c < 6; c++) o[i * 6 + c] = acc[u][i][c];
  }
  if (gwarp) {
#pragma unroll
    for (int v = 0; v < 3; v++) {
      const int gi = glane + 32 * v;
      if (gi < K6) { out[NPb * 36 + gi] = g_red[v]; out[NPb * 36 + K6 + gi] = g_bp[v]; out[NPb * 36 + 2 * K6 + gi] = g_dg[v]; }
    }
  }
  __syncthreads();
  const double c = block_sum(cost_t, sRed);
  const double gm = block_max(gmax_t, sRed + 8);
  if (tid == 0) {
    const int REC = NPb * 36 + 3 * K6;
    out[REC] = c; out[REC + 1] = gm; out[REC + 2] = 0; out[REC + 3] = 0;
  }
}

// =============================================================================================
// linearize, FP64 tensor-core variant (default when the extrinsics are fixed).
//
// Same factor evaluation (A1) as above, but (i) a chunk is a run of landmarks holding <= 256 FACTORS, one
// factor per thread, so no lane idles on a short track, and (ii) every accumulation into the 6K x 6K block
// matrix is a small dense Gram product issued as mma.sync.m8n8k4.f64 (DMMA: the same IEEE double FMAs as
// DFMA at 1/8 of the issue slots -- this kernel was issue-bound, not FP64-pipe-bound):
//   P1   S -= W~^T W~        W~ [landmark][6K+1]: sqrt(1/(h+d)) w_l scattered to frame positions, zero where
//                            unobserved, last column beta_l = b_l sqrt(1/(h+d)) => row 6K of the product is the
//                            Schur correction of the gradient.  Lower 8x8 tiles spread over the 8 warps.
//   P2a  (p,p) += X_p^T X_p  X_p = rows [B_f | r_f] of the factors observed in frame p (column 6 => sum B^T r)
//   P2b  (p,q) += B^T A      over the landmarks anchored at q that are seen in p
//   AtA  (q,q) += Y^T Y      Y = rows [A_f | r_f] of all factors anchored at q, split-K over the warps
// P1/P2a live in registers across chunks, P2b/AtA go to a shared-memory copy of the tile record each chunk
// (single owner per element, fixed order => bit-reproducible).  Landmarks are expected grouped by anchor
// frame (FeatureManager's list order); any order is correct, grouped is fast (the q loop has 1-2 trips).
// =============================================================================================
constexpr int MM_NF = BA_THREADS;   // factor slots per chunk
constexpr int MM_CL = 40;           // landmarks per chunk (multiple of 4)

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__host__ __device__ inline int mm_ntile(int K) { return (6 * K + 1 + 7) / 8; }
// row stride of W~ in doubles, = 4 or 12 (mod 16): the DMMA fragment loads (lane (g, tq) reads row tq, column g of a tile)
// then hit 16 distinct 8-byte banks per half-warp (a stride of 8 mod 16 put rows tq and tq + 2 on the same banks: 2-way)
__host__ __device__ inline int mm_wstride(int K) { return 8 * mm_ntile(K) + 4; }

template <int TM, int WSC>   // WSC: compile-time row stride of W~ (0 = runtime) so that P1's loads take immediates
__global__ void __launch_bounds__(BA_THREADS, 2) ba_linearize_mma_kernel(BaBatch bt) {
  extern __shared__ double sm[];
  const int w = blockIdx.y, t = blockIdx.x;
  BaCtrl* ctrl = bt.ctrl + w;
  if (ctrl->done) return;
  const int cur = ctrl->cur;
  if (t >= bt.TL) {   // IMU / prior CTAs: one for all, or (latency mode) the prior CTA followed by one CTA per IMU factor
    if (gridDim.x == (unsigned)bt.TL + 1) imu_prior_linearize(bt, w, cur, sm, 0, bt.K - 1, true);
    else if (t == bt.TL) imu_prior_linearize(bt, w, cur, sm, 0, 0, true);
    else imu_prior_linearize(bt, w, cur, sm, t - bt.TL - 1, t - bt.TL, false);
    return;
  }

  const int K = bt.K, K6 = 6 * K, NPb = K * (K + 1) / 2, NT = mm_ntile(K), WS = WSC ? WSC : mm_wstride(K);
  const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5, g = lane >> 2, tq = lane & 3;
  double* sFr = sm;                                  // [K*FR]
  double* sEx = sFr + K * FR;                        // [FR]
  double* sRed = sEx + FR;                           // [16]
  double* sFac = sRed + 16;                          // [MM_NF*STG] A(12) B(12) c(2) r(2) per factor
  double* sW = sFac + MM_NF * STG;                   // [MM_CL*WS]
  double* sSc = sW + MM_CL * WS;                     // [MM_CL*4] 1/(h+d), b, -, -
  double* sAcc = sSc + MM_CL * 4;                    // [NPb*36] the tile record's blocks
  double* sGb = sAcc + NPb * 36;                     // [K6] sum J^T r (unreduced gradient)
  double* sGr = sGb + K6;                            // [K6] Schur correction of the gradient
  double* sDg = sGr + K6;                            // [K6] diag of the unreduced H_pp
  double* sPart = sDg + K6;                          // [8*64] split-K partial tiles of AtA
  int* sFirst = reinterpret_cast<int*>(sPart + 8 * 64);   // [MM_CL] first factor slot of the landmark
  int* sNobs = sFirst + MM_CL;                       // [MM_CL]
  int* sO0 = sNobs + MM_CL;                          // [MM_CL] first observation (global index)
  int* sAnc = sO0 + MM_CL;                           // [MM_CL] anchor frame
  int* sQlo = sAnc + MM_CL;                          // [BVIO_KMAX] first / one-past-last landmark of the chunk anchored
  int* sQhi = sQlo + BVIO_KMAX;                      // [BVIO_KMAX] at frame q (landmarks arrive grouped by anchor)
  short* sSlot = reinterpret_cast<short*>(sQhi + BVIO_KMAX);      // [(MM_CL+4)*K] frame -> factor slot, -1 unobserved, -2 anchor
  unsigned char* sFl = reinterpret_cast<unsigned char*>(sSlot + (MM_CL + 4) * K);   // [MM_NF] slot -> landmark
  unsigned char* sFp = sFl + MM_NF;                  // [MM_NF] slot -> frame

  // P1 tiles of this warp: lower-triangular tile index wp + 8 s
  const int ntile = NT * (NT + 1) / 2;
  int ti[TM], tj[TM];
  double accW[TM][2];
#pragma unroll
  for (int s = 0; s < TM; s++) {
    int idx = wp + 8 * s, i = 0;
    ti[s] = -1; tj[s] = 0;
    if (idx < ntile) {
      while ((i + 1) * (i + 2) / 2 <= idx) i++;
      ti[s] = 8 * i; tj[s] = 8 * (idx - i * (i + 1) / 2);   // column offsets of the tile's row / column fragment
    }
    accW[s][0] = accW[s][1] = 0.0;
  }
  // P2a: frames wp and wp + 8
  double accD[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};   // two independent chains per tile
  double cost_t = 0, gmax_t = 0;

  for (int i = tid; i < NPb * 36 + 3 * K6; i += BA_THREADS) sAcc[i] = 0.0;
  stage_frames(bt, w, bt.pose[cur], bt.exs[cur], sFr, sEx);
  const double dfac = damp_factor(ctrl->radius, ctrl->mu, bt.strategy);
  const int first = ctrl->first;
  const double cauchy_b = bt.cauchy_a * bt.cauchy_a, cauchy_c = 1.0 / cauchy_b;
  const double* invd = bt.invd[cur];
  int l0, l1;
  tile_range(bt, w, t, l0, l1, bt.TL);

  for (int lb = l0; lb < l1;) {
    // ---- chunk extent: as many landmarks as fit MM_NF factor slots (and MM_CL rows)
    const int obase = bt.lm_off[lb];
    int pred = 0;
    if (tid >= 1 && tid <= MM_CL && lb + tid <= l1) pred = (bt.lm_off[lb + tid] - obase - tid) <= MM_NF;
    const int nl = __syncthreads_count(pred);        // also: previous chunk fully consumed, frames staged
    const int nfac = bt.lm_off[lb + nl] - obase - nl;
    const int nl4 = (nl + 3) & ~3;
    for (int i = tid; i < nl4 * WS; i += BA_THREADS) sW[i] = 0.0;
    for (int i = tid; i < (nl + 4) * K; i += BA_THREADS) sSlot[i] = -1;   // four pad rows: the pair loops need no bound check
    if (tid < BVIO_KMAX) { sQlo[tid] = MM_CL; sQhi[tid] = 0; }
    __syncthreads();
    // ---- A0: slot tables
    if (tid < nl) {
      const int l = lb + tid;
      const int o0 = bt.lm_off[l], n = bt.lm_off[l + 1] - o0, fs = (o0 - obase) - tid;
      const int fi = bt.obs_frame[o0];
      sFirst[tid] = fs; sNobs[tid] = n; sO0[tid] = o0; sAnc[tid] = fi;
      sSlot[tid * K + fi] = -2;
      for (int k = 1; k < n; k++) {
        const int fj = bt.obs_frame[o0 + k], slot = fs + k - 1;
        sSlot[tid * K + fj] = (short)slot;
        sFl[slot] = (unsigned char)tid; sFp[slot] = (unsigned char)fj;
      }
      atomicMin(&sQlo[fi], tid);
      atomicMax(&sQhi[fi], tid + 1);
    }
    __syncthreads();
    // ---- A1: factor evaluation, one factor per thread
    if (tid < nfac) {
      const int lc = sFl[tid], fj = sFp[tid], l = lb + lc;
      const int o0 = sO0[lc], fi = sAnc[lc], ko = o0 + 1 + (tid - sFirst[lc]);
      const double lam = invd[l];
      const double2 pi = bt.obs_xy[o0], pj = bt.obs_xy[ko];
      const double* Fi = sFr + fi * FR;
      const double* Fj = sFr + fj * FR;
      ProjGeom gm = proj_geom(Fi, Fj, sEx, pi.x, pi.y, lam);
      const double inv = 1.0 / gm.pcj.z, si = bt.sqrt_info;
      const double r0 = si * (gm.pcj.x * inv - pj.x), r1 = si * (gm.pcj.y * inv - pj.y);
      const double red[2][3] = {{si * inv, 0.0, -si * gm.pcj.x * inv * inv}, {0.0, si * inv, -si * gm.pcj.y * inv * inv}};
      double Gm[2][3], Q[2][3];
#pragma unroll
      for (int a = 0; a < 2; a++)
#pragma unroll
        for (int c = 0; c < 3; c++)
          Gm[a][c] = red[a][0] * sEx[c * 3 + 0] + red[a][1] * sEx[c * 3 + 1] + red[a][2] * sEx[c * 3 + 2];
#pragma unroll
      for (int a = 0; a < 2; a++)
#pragma unroll
        for (int c = 0; c < 3; c++) Q[a][c] = Gm[a][0] * Fj[c * 3 + 0] + Gm[a][1] * Fj[c * 3 + 1] + Gm[a][2] * Fj[c * 3 + 2];
      // Cauchy loss: rho' always; rho = a^2 log(1 + s / a^2) only feeds the cost, which ba_solve reads from the
      // linearization only the first time (later it is the accepted candidate's, from ba_cost)
      const double csum = 1.0 + (r0 * r0 + r1 * r1) * cauchy_c;
      if (first) cost_t += 0.5 * cauchy_b * log(csum);
      const double sr = sqrt(fmax(DBL_MIN, 1.0 / csum));
      double* st = sFac + (size_t)tid * STG;
      const d3 dimu = gm.pimu_i - d3{sEx[9], sEx[10], sEx[11]};
      double cc[2];
#pragma unroll
      for (int a = 0; a < 2; a++) {
        const d3 u = mtv3(Fi, d3{Q[a][0], Q[a][1], Q[a][2]});              // Ri^T Q[a]^T
        const d3 jr = cross3(gm.pimu_i, u);                                // -(Q Ri [pts_imu_i]x) row
        const d3 jjr = cross3(d3{Gm[a][0], Gm[a][1], Gm[a][2]}, gm.pimu_j); // (Gm [pts_imu_j]x) row
        st[a * 6 + 0] = sr * Q[a][0]; st[a * 6 + 1] = sr * Q[a][1]; st[a * 6 + 2] = sr * Q[a][2];
        st[a * 6 + 3] = sr * jr.x; st[a * 6 + 4] = sr * jr.y; st[a * 6 + 5] = sr * jr.z;
        st[12 + a * 6 + 0] = -sr * Q[a][0]; st[12 + a * 6 + 1] = -sr * Q[a][1]; st[12 + a * 6 + 2] = -sr * Q[a][2];
        st[12 + a * 6 + 3] = sr * jjr.x; st[12 + a * 6 + 4] = sr * jjr.y; st[12 + a * 6 + 5] = sr * jjr.z;
        cc[a] = sr * (-dot3(u, dimu) / lam);
      }
      st[24] = cc[0]; st[25] = cc[1]; st[26] = sr * r0; st[27] = sr * r1;
      double* wo = sW + lc * WS + 6 * fj;
      double* wg = bt.w + (size_t)ko * 6;
#pragma unroll
      for (int k = 0; k < 6; k++) { const double v = st[12 + k] * cc[0] + st[18 + k] * cc[1]; wo[k] = v; wg[k] = v; }
    }
    __syncthreads();
    // ---- A2: two threads per landmark sum over its factors: A^T c (the anchor's w, 3 components each), and
    //      h = sum c^T c with the damping of the eliminated depth (Ceres LevenbergMarquardtStrategy / dogleg mu,
    //      Jacobi scaling) on one, b = sum c^T r on the other
    if (tid < 2 * nl) {
      const int lc = tid >> 1, half = tid & 1, o3 = 3 * half, l = lb + lc;
      const int nf = sNobs[lc] - 1;
      const double* st = sFac + (size_t)sFirst[lc] * STG;
      double a0 = 0, a1 = 0, a2 = 0, hb = 0;
      for (int f = 0; f < nf; f++, st += STG) {
        const double c0 = st[24], c1 = st[25];
        a0 += st[o3] * c0 + st[o3 + 6] * c1;
        a1 += st[o3 + 1] * c0 + st[o3 + 7] * c1;
        a2 += st[o3 + 2] * c0 + st[o3 + 8] * c1;
        hb += half ? c0 * st[26] + c1 * st[27] : c0 * c0 + c1 * c1;
      }
      double* wo = sW + lc * WS + 6 * sAnc[lc] + o3;
      double* wg = bt.w + (size_t)sO0[lc] * 6 + o3;
      wo[0] = a0; wo[1] = a1; wo[2] = a2; wg[0] = a0; wg[1] = a1; wg[2] = a2;
      if (half) {
        sW[lc * WS + K6] = hb;                       // column 6K of W: b_l => row 6K of P1 = Schur gradient term
        gmax_t = fmax(gmax_t, fabs(hb));
        bt.b[l] = hb;
      } else {
        const double h = hb;
        double sl2 = 1.0;
        if (bt.jacobi_scaling) {
          if (first) { const double q = 1.0 / (1.0 + sqrt(h)); sl2 = q * q; }
          else sl2 = bt.sl2[l];
        }
        const double ddl = damp_term(fmin(fmax(sl2 * h, 1e-6), 1e32), sl2, dfac);
        double inv_hd = 1.0 / (h + ddl);
        if (bt.undamped) inv_hd = (h > 0) ? 1.0 / h : 0.0;
        sSc[lc * 4] = inv_hd;
        bt.h[l] = h;
        if (first) bt.sl2[l] = sl2;
      }
    }
    __syncthreads();
    // ---- AtA(q) partials first (their reduction overlaps the other products)
    int nq = 0;
    for (int q = 0; q < K; q++) {
      const int lq0 = sQlo[q], lq1 = sQhi[q];
      if (lq1 <= lq0) continue;
      if (nq++ > 0) __syncthreads();                 // sPart reused
      {
        // factor slots of the landmarks anchored at q are contiguous: [f0, f1)
        const int f0 = sFirst[lq0], f1 = sFirst[lq1 - 1] + sNobs[lq1 - 1] - 1;
        const int goff = (g < 6 ? 6 * (tq & 1) + g : 26 + (tq & 1));
        const unsigned span = (unsigned)(f1 - f0);
        double c0 = 0, c1 = 0, e0 = 0, e1 = 0;
        const double* px = sFac + goff;
        for (int sb = (f0 & ~1) + 2 * wp; sb < f1; sb += 32) {   // warp-uniform trip count (mma.sync); 2 chains
          const int s0 = sb + (tq >> 1), s1 = s0 + 16;
          double x = 0.0, y = 0.0;
          if ((unsigned)(s0 - f0) < span && g < 7) x = px[(size_t)s0 * STG];
          if ((unsigned)(s1 - f0) < span && g < 7) y = px[(size_t)s1 * STG];
          dmma884(c0, c1, x, x);
          dmma884(e0, e1, y, y);
        }
        sPart[wp * 64 + g * 8 + 2 * tq] = c0 + e0; sPart[wp * 64 + g * 8 + 2 * tq + 1] = c1 + e1;
      }
      __syncthreads();
      if (tid < 64) {
        const int m = tid >> 3, n = tid & 7;
        double sacc = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) sacc += sPart[k * 64 + tid];
        if (m < 6 && n < 6) {
          sAcc[tri(q, q) * 36 + m * 6 + n] += sacc;
          if (m == n) sDg[6 * q + m] += sacc;
        } else if (m == 6 && n < 6) sGb[6 * q + n] += sacc;
      }
      // ---- P2b: blocks (p, q), p > q, one warp per p (warps taken from the top: P2a loads the low warps)
      for (int p = q + 1 + (7 - wp); p < K; p += 8) {
        double d0 = 0, d1 = 0, e0 = 0, e1 = 0;
        const short* sl = sSlot + (lq0 + (tq >> 1)) * K + p;
        const double* pf = sFac + 6 * (tq & 1) + g;
        for (int lc = lq0 + (tq >> 1); lc - (tq >> 1) < lq1; lc += 4, sl += 4 * K) {
          double xa = 0.0, xb = 0.0, ya = 0.0, yb = 0.0;
          const int s0 = sl[0], s1 = sl[2 * K];
          if (s0 >= 0 && g < 6 && lc < lq1) { const double* st = pf + (size_t)s0 * STG; xa = st[12]; xb = st[0]; }
          if (s1 >= 0 && g < 6 && lc + 2 < lq1) { const double* st = pf + (size_t)s1 * STG; ya = st[12]; yb = st[0]; }
          dmma884(d0, d1, xa, xb);
          dmma884(e0, e1, ya, yb);
        }
        if (g < 6 && tq < 3) {
          double* o = sAcc + tri(p, q) * 36 + g * 6 + 2 * tq;
          o[0] += d0 + e0; o[1] += d1 + e1;
        }
      }
    }
    // ---- P1: S -= W^T diag(1/(h+d)) W (register tiles; the row scaling rides on the A fragment)
    {
      const double* wb = sW + tq * WS + g;
#pragma unroll
      for (int kk = 0; kk < MM_CL / 4; kk++) {
        if (4 * kk < nl4) {
          const double inv = (4 * kk + tq < nl) ? sSc[(4 * kk + tq) * 4] : 0.0;
          const double* wr = wb + kk * 4 * WS;
#pragma unroll
          for (int s = 0; s < TM; s++) {
            if (ti[s] < 0) continue;
            dmma884(accW[s][0], accW[s][1], wr[ti[s]] * inv, wr[tj[s]]);
          }
        }
      }
    }
    // ---- P2a: diagonal blocks (p, p) from the factors seen in frame p (non-anchor side)
#pragma unroll
    for (int u = 0; u < 2; u++) {
      const int p = wp + 8 * u;
      if (p >= K) continue;
      const short* sl = sSlot + (tq >> 1) * K + p;
      const double* px = sFac + (g < 6 ? 12 + 6 * (tq & 1) + g : 26 + (tq & 1));
      for (int lc = 0; lc < nl; lc += 4, sl += 4 * K) {
        double x = 0.0, y = 0.0;
        const int s0 = sl[0], s1 = sl[2 * K];           // rows nl .. nl+3 are padding (-1)
        if (s0 >= 0 && g < 7) x = px[(size_t)s0 * STG];
        if (s1 >= 0 && g < 7) y = px[(size_t)s1 * STG];
        dmma884(accD[u][0], accD[u][1], x, x);
        dmma884(accD[u][2], accD[u][3], y, y);
      }
    }
    lb += nl;
  }
  // ---- fold the register tiles into the shared record (single owner per element in each phase)
  __syncthreads();
#pragma unroll
  for (int u = 0; u < 2; u++) {
    const int p = wp + 8 * u;
    if (p >= K) continue;
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const int m = g, n = 2 * tq + e;
      const double v = accD[u][e] + accD[u][2 + e];
      if (m < 6 && n < 6) {
        sAcc[tri(p, p) * 36 + m * 6 + n] += v;
        if (m == n) sDg[6 * p + m] += v;
      } else if (m == 6 && n < 6) sGb[6 * p + n] += v;
    }
  }
  __syncthreads();
#pragma unroll
  for (int s = 0; s < TM; s++) {
    if (ti[s] < 0) continue;
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const int r = ti[s] + g, c = tj[s] + 2 * tq + e;
      if (c >= K6 || c > r) continue;
      if (r < K6) sAcc[tri(r / 6, c / 6) * 36 + (r % 6) * 6 + (c % 6)] -= accW[s][e];
      else if (r == K6) sGr[c] = accW[s][e];
    }
  }
  __syncthreads();
  // ---- one tile record to HBM
  double* out = bt.tile_out + (size_t)(w * bt.TL + t) * tile_rec_doubles(K);
  for (int i = tid; i < NPb * 36; i += BA_THREADS) out[i] = sAcc[i];
  for (int i = tid; i < K6; i += BA_THREADS) {
    out[NPb * 36 + i] = sGb[i] - sGr[i];
    out[NPb * 36 + K6 + i] = sGb[i];
    out[NPb * 36 + 2 * K6 + i] = sDg[i];
  }
  const double c = block_sum(cost_t, sRed);
  const double gmx = block_max(gmax_t, sRed + 8);
  if (tid == 0) {
    const int REC = NPb * 36 + 3 * K6;
    out[REC] = c; out[REC + 1] = gmx; out[REC + 2] = 0; out[REC + 3] = 0;
  }
}

size_t ba_linearize_mma_smem_bytes(int K) {
  const int NPb = K * (K + 1) / 2, WS = mm_wstride(K);
  size_t d = (size_t)(K + 1) * FR + 16 + (size_t)MM_NF * STG + (size_t)MM_CL * WS + MM_CL * 4 + (size_t)NPb * 36 + 18 * K + 8 * 64;
  size_t bytes = d * sizeof(double) + (size_t)(4 * MM_CL + 2 * BVIO_KMAX) * sizeof(int) + (size_t)((MM_CL + 4) * K) * sizeof(short) + 2 * MM_NF;
  return (bytes + 15) & ~size_t(15);
}

// IMU factors and marginalization prior of every window, one CTA per window (thread f evaluates the raw Jacobian of
// factor f, all threads whiten and form J^T J; then the prior): the companion launch of ba_linearize_ws_kernel, whose CTAs
// fill an SM each and carry the visual tiles only.  (One CTA per (window, factor) was measured 2x slower: the raw
// Jacobian is one thread's work either way, and eleven times the CTAs only queue behind each other.)
constexpr int IMU_THREADS = 256;
__global__ void __launch_bounds__(IMU_THREADS, 2) ba_imu_prior_kernel(BaBatch bt) {
  extern __shared__ double sm[];
  const int w = blockIdx.x;
  const BaCtrl* ctrl = bt.ctrl + w;
  if (ctrl->done) return;
  imu_prior_linearize(bt, w, ctrl->cur, sm, 0, bt.K - 1, true);
}

// =============================================================================================
// Warp-specialised variant (throughput mode): the same stages as ba_linearize_mma_kernel, but the factor evaluation
// (A0-A2: FP64 pipe, register-hungry) and the Gram products (DMMA: tensor pipe, accumulators in registers) run in
// DIFFERENT warps of one 512-thread CTA per SM and overlap chunk by chunk through double-buffered shared memory:
//   warps 0-7   producers: chunk c -> buffer c & 1 (slot tables, factor records, W~ rows, 1/(h+d)), h / b / w to HBM
//   warps 8-15  consumers: AtA, P2b, P1, P2a over buffer c & 1; tile record to HBM at the end
// Hand-over with named barriers (bar.arrive by the side that is done, bar.sync by the side that waits): FULL[b] / EMPTY[b],
// plus one barrier id per role for its internal phases.  The producers no longer carry the 36 accumulator registers
// through the evaluation, the consumers never wait for an evaluation they do not depend on.
// =============================================================================================
constexpr int WS_THREADS = 512, WS_ROLE = 256;
#ifdef BVIO_WS_PROF   // development build only (make WSPROF=1): cycles per role / phase, summed over CTAs, printed at bvio_destroy
__device__ unsigned long long g_ws_prof[16 + 16 * 16];   // [16 + 16 * (8 * role + warp) + phase] per-warp phase cycles
#define WSP_DECL() __shared__ unsigned long long s_wsp[16 * 16]
#define WSP_INIT() do { for (int i_ = threadIdx.x; i_ < 256; i_ += blockDim.x) s_wsp[i_] = 0; } while (0)
#define WSP_T0() long long wsp_t = clock64()
// accumulated per CTA in shared memory (a global atomic per probe would queue up in the LSU and distort what follows)
#define WSP_ADD(i) do { const long long n_ = clock64(); if (lane == 0) s_wsp[16 * (8 * role + wp) + (i)] += (unsigned long long)(n_ - wsp_t); wsp_t = clock64(); } while (0)
#define WSP_FLUSH() do { if (lane == 0) for (int i_ = 0; i_ < 16; i_++) atomicAdd(&g_ws_prof[16 + 16 * (8 * role + wp) + i_], s_wsp[16 * (8 * role + wp) + i_]); } while (0)
#define WSP_CHUNK() atomicAdd(&g_ws_prof[15], 1ull)
#else
#define WSP_CHUNK() do {} while (0)
#define WSP_DECL() do {} while (0)
#define WSP_INIT() do {} while (0)
#define WSP_FLUSH() do {} while (0)
#define WSP_T0() do {} while (0)
#define WSP_ADD(i) do {} while (0)
#endif
enum { BAR_PROD = 1, BAR_CONS = 2, BAR_FULL0 = 3, BAR_EMPTY0 = 5 };
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

struct WsBuf {            // one chunk's staging area
  double *fac, *w, *sc;
  int *first, *nobs, *o0, *anc, *qlo, *qhi, *qlist, *meta;   // qlist: anchors present in the chunk, q | lq0 << 8 | lq1 << 16
  short* slot;
};
// fac is preceded by one all-zero record (slot -1): masked fragment loads read it instead of branching
__host__ __device__ inline size_t ws_buf_doubles(int WS) { return (size_t)(MM_NF + 1) * STG + 1 + (size_t)MM_CL * WS + MM_CL * 4; }
__host__ __device__ inline size_t ws_buf_ints(void) { return 4 * (MM_CL + 2) + 3 * BVIO_KMAX + 8; }   // first / o0 hold MM_CL + 1 entries
__host__ __device__ inline size_t ws_buf_shorts(int K) { return (size_t)((MM_CL + 8) * K + 1) & ~size_t(1); }

template <int TM, int WSC>
__global__ void __launch_bounds__(WS_THREADS, 1) ba_linearize_ws_kernel(BaBatch bt) {
  extern __shared__ double sm[];
  const int w = blockIdx.y, t = blockIdx.x;
  BaCtrl* ctrl = bt.ctrl + w;
  if (ctrl->done) return;
  const int cur = ctrl->cur;              // (the IMU factors and the prior: ba_imu_prior_kernel, launched beside this one)
  const int K = bt.K, K6 = 6 * K, NPb = K * (K + 1) / 2, NT = mm_ntile(K), WS = WSC ? WSC : mm_wstride(K);
  const int role = threadIdx.x >> 8, tid = threadIdx.x & (WS_ROLE - 1), lane = tid & 31, wp = tid >> 5, g = lane >> 2, tq = lane & 3;
  // ---- shared memory: common part, then the two chunk buffers (doubles, ints, shorts, bytes)
  double* sFr = sm;                                  // [K*FR]
  double* sEx = sFr + K * FR;                        // [FR]
  double* sRed = sEx + FR;                           // [32]
  double* sAcc = sRed + 32;                          // [NPb*36]
  double* sGb = sAcc + NPb * 36;                     // [K6]
  double* sGr = sGb + K6;                            // [K6]
  double* sDg = sGr + K6;                            // [K6]
  double* sAtA = sDg + K6;                           // [K][8 warps][28] lower triangle of the split-K AtA partials
  double* sPiece = sAtA + K * 8 * 28;                // [MM_CL][2 pieces][8] per-landmark sums of the current chunk (producers)
  double* dbl = sPiece + MM_CL * 16;
  int* ints = reinterpret_cast<int*>(dbl + 2 * ws_buf_doubles(WS));
  short* shorts = reinterpret_cast<short*>(ints + 2 * ws_buf_ints());
  auto buffer = [&](int b) {
    WsBuf u;
    u.fac = dbl + b * ws_buf_doubles(WS) + STG; u.w = u.fac + MM_NF * STG + 1; u.sc = u.w + MM_CL * WS;
    u.first = ints + b * ws_buf_ints(); u.nobs = u.first + MM_CL + 2; u.o0 = u.nobs + MM_CL + 2; u.anc = u.o0 + MM_CL + 2;
    u.qlo = u.anc + MM_CL + 2; u.qhi = u.qlo + BVIO_KMAX; u.qlist = u.qhi + BVIO_KMAX; u.meta = u.qlist + BVIO_KMAX;
    u.slot = shorts + b * ws_buf_shorts(K);
    return u;
  };

  WSP_DECL();
  WSP_INIT();
  for (int i = threadIdx.x; i < NPb * 36 + 3 * K6 + K * 8 * 28; i += WS_THREADS) sAcc[i] = 0.0;
  if (threadIdx.x < 2 * STG) dbl[(threadIdx.x / STG) * ws_buf_doubles(WS) + threadIdx.x % STG] = 0.0;   // the zero records
  stage_frames(bt, w, bt.pose[cur], bt.exs[cur], sFr, sEx);
  int l0, l1;
  tile_range(bt, w, t, l0, l1, bt.TL);
  __syncthreads();                                   // the last CTA-wide barrier: the roles part here

  if (role == 0) {
    // =========================================== producers ===========================================
    const double dfac = damp_factor(ctrl->radius, ctrl->mu, bt.strategy);
    const int first = ctrl->first;
    const double* invd = bt.invd[cur];
    const double cauchy_b = bt.cauchy_a * bt.cauchy_a, cauchy_c = 1.0 / cauchy_b;
    double cost_t = 0, gmax_t = 0;
    int c = 0;
    WSP_T0();
    // chunk extents are computed one chunk ahead (every warp on its own): lane j holds lm_off[lb + j + 1] and [lb + j + 33]
    int obase = l0 < l1 ? bt.lm_off[l0] : 0, v1 = 0, v2 = 0;
    if (l0 + lane + 1 <= l1) v1 = bt.lm_off[l0 + lane + 1];
    if (lane + 33 <= MM_CL && l0 + lane + 33 <= l1) v2 = bt.lm_off[l0 + lane + 33];
    for (int lb = l0; lb < l1; c++) {
      const int b = c & 1;
      const WsBuf u = buffer(b);
      // chunk extent: as many landmarks as fit MM_NF factor slots (and MM_CL rows)
      const int j1 = lane + 1, j2 = lane + 33;
      const bool p1 = lb + j1 <= l1 && (v1 - obase - j1) <= MM_NF;
      const bool p2 = j2 <= MM_CL && lb + j2 <= l1 && (v2 - obase - j2) <= MM_NF;
      const int nl = __popc(__ballot_sync(0xffffffffu, p1)) + __popc(__ballot_sync(0xffffffffu, p2));
      const int onext = nl <= 32 ? __shfl_sync(0xffffffffu, v1, (nl - 1) & 31) : __shfl_sync(0xffffffffu, v2, (nl - 33) & 31);
      const int nfac = onext - obase - nl;
      const int nl4 = (nl + 3) & ~3;
      // first observation of landmarks lane and lane + 32 of this chunk (for the staging below), then prefetch the next extent
      const int up1 = __shfl_up_sync(0xffffffffu, v1, 1), up2 = __shfl_up_sync(0xffffffffu, v2, 1), v1last = __shfl_sync(0xffffffffu, v1, 31);
      const int oa = lane == 0 ? obase : up1, ob = lane == 0 ? v1last : up2;
      const int na = v1 - oa, nb = v2 - ob;            // observations of landmarks lane / lane + 32
      const int cur_obase = obase;
      {
        const int lbn = lb + nl;
        obase = onext; v1 = 0; v2 = 0;
        if (lbn + lane + 1 <= l1) v1 = bt.lm_off[lbn + lane + 1];
        if (lane + 33 <= MM_CL && lbn + lane + 33 <= l1) v2 = bt.lm_off[lbn + lane + 33];
      }
      WSP_ADD(0);
      if (c >= 2) bar_sync(BAR_EMPTY0 + b, WS_THREADS);          // the consumers are done with this buffer
      WSP_ADD(1);
      for (int i = tid; i < nl4 * WS / 2; i += WS_ROLE) reinterpret_cast<double2*>(u.w)[i] = double2{0.0, 0.0};
      for (int i = tid; i < (nl + 8) * K; i += WS_ROLE) u.slot[i] = -1;   // 8 pad rows: the pair loops need no bound check
      if (tid < BVIO_KMAX) { u.qlo[tid] = MM_CL; u.qhi[tid] = 0; }
      if (wp == 0) {                                   // per-landmark tables: first observation, first factor slot, track length
        if (lane <= nl) { u.o0[lane] = oa; u.first[lane] = oa - cur_obase - lane; if (lane < nl) { u.nobs[lane] = na; u.anc[lane] = 0; } }
        if (lane + 32 <= nl) { u.o0[lane + 32] = ob; u.first[lane + 32] = ob - cur_obase - lane - 32; if (lane + 32 < nl) { u.nobs[lane + 32] = nb; u.anc[lane + 32] = 0; } }
        if (lane == 0) u.meta[0] = nl;
      }
      bar_sync(BAR_PROD, WS_ROLE);
      WSP_ADD(2);
      WSP_ADD(3);
      // ---- A1: factor evaluation, one factor per thread
      double red8[8] = {0, 0, 0, 0, 0, 0, 0, 0};      // this factor's share of its landmark's sums: A^T c (6), c^T c, c^T r
      int lcs = -1, fss = 0;                           // landmark / its first factor slot (-1: idle thread)
      if (tid < nfac) {
        int lc = 0;                                    // the landmark of factor slot tid: the last one with first[lc] <= tid
        for (int hi = nl; hi - lc > 1;) { const int mid = (lc + hi) >> 1; if (u.first[mid] <= tid) lc = mid; else hi = mid; }
        const int l = lb + lc, fs = u.first[lc];
        const int o0 = u.o0[lc], ko = o0 + 1 + (tid - fs);
        const int fi = bt.obs_frame[o0], fj = bt.obs_frame[ko];
        const double lam = invd[l];
        const double2 pi = bt.obs_xy[o0], pj = bt.obs_xy[ko];
        u.slot[lc * K + fj] = (short)tid;
        if (tid == fs) { u.anc[lc] = fi; atomicMin(&u.qlo[fi], lc); atomicMax(&u.qhi[fi], lc + 1); }
        const double* Fi = sFr + fi * FR;
        const double* Fj = sFr + fj * FR;
        ProjGeom gm = proj_geom(Fi, Fj, sEx, pi.x, pi.y, lam);
        const double inv = 1.0 / gm.pcj.z, si = bt.sqrt_info;
        const double r0 = si * (gm.pcj.x * inv - pj.x), r1 = si * (gm.pcj.y * inv - pj.y);
        const double red[2][3] = {{si * inv, 0.0, -si * gm.pcj.x * inv * inv}, {0.0, si * inv, -si * gm.pcj.y * inv * inv}};
        double Gm[2][3], Q[2][3];
#pragma unroll
        for (int a = 0; a < 2; a++)
#pragma unroll
          for (int cc = 0; cc < 3; cc++)
            Gm[a][cc] = red[a][0] * sEx[cc * 3 + 0] + red[a][1] * sEx[cc * 3 + 1] + red[a][2] * sEx[cc * 3 + 2];
#pragma unroll
        for (int a = 0; a < 2; a++)
#pragma unroll
          for (int cc = 0; cc < 3; cc++) Q[a][cc] = Gm[a][0] * Fj[cc * 3 + 0] + Gm[a][1] * Fj[cc * 3 + 1] + Gm[a][2] * Fj[cc * 3 + 2];
        // Cauchy loss: rho' = 1 / (1 + s / a^2) always; rho = a^2 log(1 + s / a^2) only feeds the cost, which the solve
        // reads from the linearization only the first time (later it is the accepted candidate's, from ba_cost)
        const double s2 = r0 * r0 + r1 * r1, csum = 1.0 + s2 * cauchy_c;
        if (first) cost_t += 0.5 * cauchy_b * log(csum);
        const double sr = sqrt(fmax(DBL_MIN, 1.0 / csum));
        double* st = u.fac + (size_t)tid * STG;
        const d3 dimu = gm.pimu_i - d3{sEx[9], sEx[10], sEx[11]};
        double cc2[2], Av[2][6], Bv[2][6];
#pragma unroll
        for (int a = 0; a < 2; a++) {
          const d3 uu = mtv3(Fi, d3{Q[a][0], Q[a][1], Q[a][2]});              // Ri^T Q[a]^T
          const d3 jr = cross3(gm.pimu_i, uu);                               // -(Q Ri [pts_imu_i]x) row
          const d3 jjr = cross3(d3{Gm[a][0], Gm[a][1], Gm[a][2]}, gm.pimu_j); // (Gm [pts_imu_j]x) row
          Av[a][0] = sr * Q[a][0]; Av[a][1] = sr * Q[a][1]; Av[a][2] = sr * Q[a][2];
          Av[a][3] = sr * jr.x; Av[a][4] = sr * jr.y; Av[a][5] = sr * jr.z;
#pragma unroll
          for (int k = 0; k < 6; k++) st[a * 6 + k] = Av[a][k];
          Bv[a][0] = -sr * Q[a][0]; Bv[a][1] = -sr * Q[a][1]; Bv[a][2] = -sr * Q[a][2];
          Bv[a][3] = sr * jjr.x; Bv[a][4] = sr * jjr.y; Bv[a][5] = sr * jjr.z;
#pragma unroll
          for (int k = 0; k < 6; k++) st[12 + a * 6 + k] = Bv[a][k];
          cc2[a] = sr * (-dot3(uu, dimu) / lam);   // (a true division: bit-compatible with the latency-mode kernel)
        }
        st[24] = cc2[0]; st[25] = cc2[1]; st[26] = sr * r0; st[27] = sr * r1;
#pragma unroll
        for (int k = 0; k < 6; k++) red8[k] = Av[0][k] * cc2[0] + Av[1][k] * cc2[1];
        red8[6] = cc2[0] * cc2[0] + cc2[1] * cc2[1];
        red8[7] = cc2[0] * (sr * r0) + cc2[1] * (sr * r1);
        lcs = lc; fss = fs;
        // 16-byte stores: rows of W~ and of bt.w start at even double offsets (WS, 6 fj and 6 ko are even)
        double2* wo = reinterpret_cast<double2*>(u.w + lc * WS + 6 * fj);
        double2* wg = reinterpret_cast<double2*>(bt.w + (size_t)ko * 6);
#pragma unroll
        for (int k = 0; k < 3; k++) {
          const double2 v{Bv[0][2 * k] * cc2[0] + Bv[1][2 * k] * cc2[1], Bv[0][2 * k + 1] * cc2[0] + Bv[1][2 * k + 1] * cc2[1]};
          wo[k] = v; wg[k] = v;
        }
      }
      // per-landmark sums: the factors of a landmark sit in consecutive threads, so a segmented scan over the warp
      // (shuffles, fixed order) leaves the sum of each run in its last thread; a landmark that straddles two warps
      // (<= 15 factors: never more) leaves two pieces, added in slot order by A2
#pragma unroll
      for (int off = 1; off < 16; off <<= 1) {
        const int lo = __shfl_up_sync(0xffffffffu, lcs, off);
        const bool take = lane >= off && lo == lcs && lcs >= 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
          const double o = __shfl_up_sync(0xffffffffu, red8[k], off);
          if (take) red8[k] += o;
        }
      }
      {
        const int ln = __shfl_down_sync(0xffffffffu, lcs, 1);
        if (lcs >= 0 && (lane == 31 || ln != lcs)) {
          double2* o = reinterpret_cast<double2*>(sPiece + (2 * lcs + (fss < (tid & ~31) ? 1 : 0)) * 8);
#pragma unroll
          for (int k = 0; k < 4; k++) o[k] = double2{red8[2 * k], red8[2 * k + 1]};
        }
      }
      bar_sync(BAR_PROD, WS_ROLE);
      WSP_ADD(4);
      // ---- A2: one thread per landmark adds the pieces: A^T c (the anchor's w), h = sum c^T c with the damping of the
      //      eliminated depth (Ceres LevenbergMarquardtStrategy / dogleg mu, Jacobi scaling), b = sum c^T r
      if (wp == 7 && lane < K) {                       // anchors present in this chunk, for the consumers' loops
        const int a = u.qlo[lane], e = u.qhi[lane];
        const unsigned present = __ballot_sync((1u << K) - 1u, e > a);
        if (e > a) u.qlist[__popc(present & ((1u << lane) - 1u))] = lane | (a << 8) | (e << 16);
        if (lane == 0) u.meta[1] = __popc(present);
      }
      if (tid < nl) {
        const int lc = tid, l = lb + lc, fs = u.first[lc], nf = u.nobs[lc] - 1;
        double sl2 = 1.0;                              // Jacobi scale^2 of the depth column: fixed at the first linearization
        if (bt.jacobi_scaling && !first) sl2 = bt.sl2[l];
        double t8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (nf > 0) {
          const double2* p0 = reinterpret_cast<const double2*>(sPiece + 2 * lc * 8);
#pragma unroll
          for (int k = 0; k < 4; k++) { const double2 v = p0[k]; t8[2 * k] = v.x; t8[2 * k + 1] = v.y; }
          if ((fs >> 5) != ((fs + nf - 1) >> 5)) {     // the run crosses a warp boundary: second piece
#pragma unroll
            for (int k = 0; k < 4; k++) { const double2 v = p0[4 + k]; t8[2 * k] += v.x; t8[2 * k + 1] += v.y; }
          }
        }
        double2* wo = reinterpret_cast<double2*>(u.w + lc * WS + 6 * u.anc[lc]);
        double2* wg = reinterpret_cast<double2*>(bt.w + (size_t)u.o0[lc] * 6);
#pragma unroll
        for (int k = 0; k < 3; k++) { const double2 v{t8[2 * k], t8[2 * k + 1]}; wo[k] = v; wg[k] = v; }
        const double h = t8[6], vb = t8[7];
        u.w[lc * WS + K6] = vb;                        // column 6K of W: b_l => row 6K of P1 = Schur gradient term
        gmax_t = fmax(gmax_t, fabs(vb));
        bt.b[l] = vb;
        if (bt.jacobi_scaling && first) { const double q = 1.0 / (1.0 + sqrt(h)); sl2 = q * q; }
        const double ddl = damp_term(fmin(fmax(sl2 * h, 1e-6), 1e32), sl2, dfac);
        double inv_hd = 1.0 / (h + ddl);
        if (bt.undamped) inv_hd = (h > 0) ? 1.0 / h : 0.0;
        u.sc[lc * 4] = inv_hd;
        bt.h[l] = h;
        if (first) bt.sl2[l] = sl2;
      }
      WSP_ADD(5);
      __threadfence_block();
      bar_arrive(BAR_FULL0 + b, WS_THREADS);           // chunk c is ready
      WSP_ADD(6);
      lb += nl;
    }
    // drain: leave no barrier half-complete (and the consumers' last reads behind us)
    if (c >= 2) bar_sync(BAR_EMPTY0 + (c & 1), WS_THREADS);
    if (c >= 1) bar_sync(BAR_EMPTY0 + ((c + 1) & 1), WS_THREADS);
    // cost and gradient-max partials of this tile
    cost_t = warp_sum(cost_t); gmax_t = warp_max(gmax_t);
    if (lane == 0) { sRed[wp] = cost_t; sRed[8 + wp] = gmax_t; }
    bar_sync(BAR_PROD, WS_ROLE);
    if (tid == 0) {
      double cs = 0, gm = sRed[8];
      for (int i = 0; i < 8; i++) { cs += sRed[i]; gm = fmax(gm, sRed[8 + i]); }
      double* out = bt.tile_out + (size_t)(w * bt.TL + t) * tile_rec_doubles(K);
      const int REC = NPb * 36 + 3 * K6;
      out[REC] = cs; out[REC + 1] = gm; out[REC + 2] = 0; out[REC + 3] = 0;
    }
    WSP_FLUSH();
    return;
  }

  // =========================================== consumers ===========================================
  const int ntile = NT * (NT + 1) / 2;
  int ti[TM], tj[TM];
  double accW[TM][2];
#pragma unroll
  for (int s = 0; s < TM; s++) {
    int idx = wp + 8 * s, i = 0;
    ti[s] = -1; tj[s] = 0;
    if (idx < ntile) {
      while ((i + 1) * (i + 2) / 2 <= idx) i++;
      ti[s] = 8 * i; tj[s] = 8 * (idx - i * (i + 1) / 2);
    }
    accW[s][0] = accW[s][1] = 0.0;
  }
  double accD[2][8] = {{0, 0, 0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0, 0, 0}};
  // P2a frames go to warps 0 .. npw-1 (two each): with K <= 12 the two top warps take none -- they carry the P2b blocks
  // (p, q) next to the diagonal, present for every anchor q, while the low warps' P2b blocks only exist for small q
  const int npw = K <= 12 ? 6 : 8;
  int curq = -1;                                       // anchor frame the AtA partials belong to
  double ata[4] = {0, 0, 0, 0};
  // this warp's split-K partial of AtA(q) goes to its own shared-memory slot (no barrier); summed over the warps at the end
  auto flush_ata = [&](int q) {
    double* o = sAtA + (q * 8 + wp) * 28 + g * (g + 1) / 2 + 2 * tq;
    if (g < 7 && 2 * tq <= g) o[0] += ata[0] + ata[2];
    if (g < 7 && 2 * tq + 1 <= g) o[1] += ata[1] + ata[3];
    ata[0] = ata[1] = ata[2] = ata[3] = 0.0;
  };
  // masked lanes of the fragment loads (rows 6 / 7 of a tile) read the zero record through a zero stride
  const int mul7 = g < 7 ? STG : 0, mul6 = g < 6 ? STG : 0;
  int c = 0;
  WSP_T0();
  for (int lb = l0; lb < l1; c++) {
    const int b = c & 1;
    const WsBuf u = buffer(b);
    bar_sync(BAR_FULL0 + b, WS_THREADS);
    WSP_ADD(8);
    const int nl = u.meta[0], nl4 = (nl + 3) & ~3;
    // ---- AtA(q): split-K partials stay in registers across chunks (landmarks arrive grouped by anchor, so q rarely
    //      changes)
    const int nq = u.meta[1];
    for (int iq = 0; iq < nq; iq++) {
      const int qe = u.qlist[iq], q = qe & 255, lq0 = (qe >> 8) & 255, lq1 = qe >> 16;
      WSP_ADD(14);
      if (q != curq) {
        if (curq >= 0) flush_ata(curq);
        curq = q;
      }
      WSP_ADD(13);
      {
        // factor slots of the landmarks anchored at q are contiguous: [f0, f1)
        const int f0 = u.first[lq0], f1 = u.first[lq1 - 1] + u.nobs[lq1 - 1] - 1;
        const unsigned span = (unsigned)(f1 - f0);
        const double* px = u.fac + (g < 6 ? 6 * (tq & 1) + g : g == 6 ? 26 + (tq & 1) : -STG);
        for (int sb = (f0 & ~1) + 2 * wp; sb < f1; sb += 32) {   // warp-uniform trip count (mma.sync); 2 chains
          const int s0 = sb + (tq >> 1), s1 = s0 + 16;
          const int i0 = (unsigned)(s0 - f0) < span ? s0 : -1, i1 = (unsigned)(s1 - f0) < span ? s1 : -1;
          const double x = px[i0 * mul7], y = px[i1 * mul7];
          dmma884(ata[0], ata[1], x, x);
          dmma884(ata[2], ata[3], y, y);
        }
      }
      WSP_ADD(9);
      // ---- P2b: blocks (p, q), p > q, one warp per p; four independent chains
      for (int p = q + 1 + (7 - wp); p < K; p += 8) {
        double d[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const short* sl = u.slot + (lq0 + (tq >> 1)) * K + p;
        const double* pf = u.fac + (g < 6 ? 6 * (tq & 1) + g : -STG);
        for (int lc = lq0 + (tq >> 1); lc - (tq >> 1) < lq1; lc += 8, sl += 8 * K) {
          double xa[4], xb[4];
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const int sj = lc + 2 * j < lq1 ? (int)sl[2 * j * K] : -1;
            const double* st = pf + sj * mul6;
            xa[j] = st[12]; xb[j] = st[0];
          }
#pragma unroll
          for (int j = 0; j < 4; j++) dmma884(d[2 * j], d[2 * j + 1], xa[j], xb[j]);
        }
        if (g < 6 && tq < 3) {
          double* o = sAcc + tri(p, q) * 36 + g * 6 + 2 * tq;
          o[0] += (d[0] + d[2]) + (d[4] + d[6]); o[1] += (d[1] + d[3]) + (d[5] + d[7]);
        }
      }
      WSP_ADD(12);
    }
    WSP_ADD(7);
    // ---- P1: S -= W^T diag(1/(h+d)) W (register tiles; the row scaling rides on the A fragment)
    {
      const double* wb = u.w + tq * WS + g;
#pragma unroll
      for (int kk = 0; kk < MM_CL / 4; kk++) {
        if (4 * kk < nl4) {
          const double inv = (4 * kk + tq < nl) ? u.sc[(4 * kk + tq) * 4] : 0.0;
          const double* wr = wb + kk * 4 * WS;
#pragma unroll
          for (int s = 0; s < TM; s++) {
            if (ti[s] < 0) continue;
            dmma884(accW[s][0], accW[s][1], wr[ti[s]] * inv, wr[tj[s]]);
          }
        }
      }
        }
    WSP_ADD(10);
    // ---- P2a: diagonal blocks (p, p) from the factors seen in frame p (non-anchor side); four independent chains
#pragma unroll
    for (int uu = 0; uu < 2; uu++) {
      const int p = wp + npw * uu;
      if (wp >= npw || p >= K) continue;
      const short* sl = u.slot + (tq >> 1) * K + p;
      const double* px = u.fac + (g < 6 ? 12 + 6 * (tq & 1) + g : g == 6 ? 26 + (tq & 1) : -STG);
      for (int lc = 0; lc < nl; lc += 8, sl += 8 * K) {
        double x[4];
#pragma unroll
        for (int j = 0; j < 4; j++) x[j] = px[(int)sl[2 * j * K] * mul7];   // unobserved / padding rows (nl .. nl+7): slot -1 = zeros
#pragma unroll
        for (int j = 0; j < 4; j++) dmma884(accD[uu][2 * j], accD[uu][2 * j + 1], x[j], x[j]);
      }
    }
    bar_arrive(BAR_EMPTY0 + b, WS_THREADS);            // buffer b may be overwritten
    WSP_ADD(11);
    if (tid == 0) { WSP_CHUNK(); }
    lb += nl;
  }
  // ---- fold the register tiles into the shared record (single owner per element in each phase)
  if (curq >= 0) flush_ata(curq);
  bar_sync(BAR_CONS, WS_ROLE);
  for (int i = tid; i < K * 28; i += WS_ROLE) {        // AtA(q): the eight warps' partials in a fixed order
    const int q = i / 28, e = i - 28 * q, m = c_triA[e], n = c_triB[e];
    const double* pp = sAtA + q * 8 * 28 + e;
    double sacc = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) sacc += pp[k * 28];
    if (m < 6) {
      sAcc[tri(q, q) * 36 + m * 6 + n] += sacc;
      if (m != n) sAcc[tri(q, q) * 36 + n * 6 + m] += sacc;
      else sDg[6 * q + m] += sacc;
    } else if (n < 6) sGb[6 * q + n] += sacc;
  }
  bar_sync(BAR_CONS, WS_ROLE);
#pragma unroll
  for (int uu = 0; uu < 2; uu++) {
    const int p = wp + npw * uu;
    if (wp >= npw || p >= K) continue;
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const int m = g, n = 2 * tq + e;
      const double v = (accD[uu][e] + accD[uu][2 + e]) + (accD[uu][4 + e] + accD[uu][6 + e]);
      if (m < 6 && n < 6) {
        sAcc[tri(p, p) * 36 + m * 6 + n] += v;
        if (m == n) sDg[6 * p + m] += v;
      } else if (m == 6 && n < 6) sGb[6 * p + n] += v;
    }
  }
  bar_sync(BAR_CONS, WS_ROLE);
  auto fold_tile = [&](int r0, int c0, const double* acc) {
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const int r = r0 + g, cc = c0 + 2 * tq + e;
      if (cc >= K6 || cc > r) continue;
      if (r < K6) sAcc[tri(r / 6, cc / 6) * 36 + (r % 6) * 6 + (cc % 6)] -= acc[e];
      else if (r == K6) sGr[cc] = acc[e];
    }
  };
#pragma unroll
  for (int s = 0; s < TM; s++)
    if (ti[s] >= 0) fold_tile(ti[s], tj[s], accW[s]);
  bar_sync(BAR_CONS, WS_ROLE);
  // ---- one tile record to HBM (cost / gradient-max slots: the producers')
  double* out = bt.tile_out + (size_t)(w * bt.TL + t) * tile_rec_doubles(K);
  for (int i = tid; i < NPb * 36; i += WS_ROLE) out[i] = sAcc[i];
  for (int i = tid; i < K6; i += WS_ROLE) {
    out[NPb * 36 + i] = sGb[i] - sGr[i];
    out[NPb * 36 + K6 + i] = sGb[i];
    out[NPb * 36 + 2 * K6 + i] = sDg[i];
  }
  WSP_FLUSH();
}

void ba_ws_prof_dump(void) {
#ifdef BVIO_WS_PROF
  static unsigned long long h[16 + 256];
  if (cudaMemcpyFromSymbol(h, g_ws_prof, sizeof h) != cudaSuccess || !h[15]) return;
  const char* nm[16] = {"extent", "waitEMPTY", "zero", "-", "A1", "A2", "arrive", "q tail", "waitFULL", "AtA", "P1", "P2a", "P2b", "flush", "read qlist", ""};
  fprintf(stderr, "[ws prof] %llu chunks; cycles per chunk per warp\n", h[15]);
  for (int r = 0; r < 2; r++)
    for (int i = 0; i < 15; i++) {
      if (!nm[i][0] || (r == 0) != (i < 8)) continue;
      fprintf(stderr, "[ws prof] %s %-10s", r ? "C" : "P", nm[i]);
      for (int wq = 0; wq < 8; wq++) fprintf(stderr, " %6.0f", (double)h[16 + 16 * (8 * r + wq) + i] / (double)h[15]);
      fprintf(stderr, "\n");
    }
#endif
}
size_t ba_linearize_ws_smem_bytes(int K) {
  const int NPb = K * (K + 1) / 2, WS = mm_wstride(K);
  size_t d = (size_t)(K + 1) * FR + 32 + (size_t)NPb * 36 + 18 * K + (size_t)K * 8 * 28 + MM_CL * 16 + 2 * ws_buf_doubles(WS);
  size_t bytes = d * sizeof(double) + 2 * ws_buf_ints() * sizeof(int) + 2 * ws_buf_shorts(K) * sizeof(short);
  return (bytes + 15) & ~size_t(15);
}

// =============================================================================================
// solve: one CTA per window
// =============================================================================================
// EX: the reduced system carries extra dimensions after the K 15-blocks (6 extrinsic and/or 1 td: np = 15K + 6 / + 1 /
// + 7); the last Cholesky panel is then partial (the panel width is a compile-time 15 otherwise).
// NTHR: 256 threads x 2 CTAs per SM for throughput (large batches), 512 x 1 for latency (fewer windows than SMs).
template <bool EX, int NTHR>
__global__ void __launch_bounds__(NTHR, NTHR <= 256 ? 2 : 1) ba_solve_kernel(BaBatch bt, int with_step) {
  extern __shared__ double sm[];
  const int w = blockIdx.x, K = bt.K, np = bt.np, KE = K + bt.est_ex + bt.est_td, K6 = 6 * KE, NPb = KE * (KE + 1) / 2;
  const int NB = (np + 14) / 15;                   // diagonal panels of the blocked Cholesky
  const int tid = threadIdx.x, nthr = blockDim.x;
  BaCtrl* ctrl = bt.ctrl + w;
  if (ctrl->done) return;
  const int N1 = np + 1;                          // augmented dimension
  double* S = sm;                                 // packed lower (N1)(N1+1)/2
  // O(np)-touched vectors live in HBM/L2 so that the packed matrix alone decides the shared-memory footprint
  // (two CTAs per SM for np = 165: one window's barrier waits hide behind the other's work)
  double* bp = bt.solve_vec + (size_t)w * 5 * np;  // [np] unreduced gradient
  double* gr = bp + np;                           // [np] reduced gradient
  double* dH = gr + np;                           // [np] diag of the undamped, unreduced H_pp
  double* ddp = dH + np;                          // [np]
  double* tv = ddp + np;                          // [np] dogleg: scaled gradient direction t
  double* yv = S + (size_t)N1 * (N1 + 1) / 2;     // [np]
  double* red = yv + np;                          // [32]
  double* dinv = red + 32;                        // [np] reciprocal diagonal of L
  __shared__ int s_fail;
  const int REC = NPb * 36 + 3 * K6;
  const int TREC = tile_rec_doubles(KE);
  const double* tiles = bt.tile_out + (size_t)w * bt.TL * TREC;
  const double* imo = bt.imu_out + (size_t)w * K * IMU_OUT;
  const int n = bt.pr_n[w];
  const int* map = bt.pr_map + (size_t)w * bt.nmax;
  const double* pH = bt.pr_H + (size_t)w * bt.nmax * bt.nmax;
  const double* po = bt.pr_out + (size_t)w * (bt.nmax + 1);

  if (tid == 0) ctrl->stamps[0] = global_ns();
  const double* S0 = bt.S0 ? bt.S0 + (size_t)w * (N1 * (N1 + 1) / 2) : nullptr;
  if (S0) for (int i = tid; i < N1 * (N1 + 1) / 2; i += nthr) S[i] = S0[i];     // prior already in place
  else for (int i = tid; i < N1 * (N1 + 1) / 2; i += nthr) S[i] = 0.0;
  if (tid == 0) s_fail = 0;
  __syncthreads();
  // visual blocks
  for (int idx = tid; idx < NPb * 36; idx += nthr) {
    double s = 0;
    for (int t = 0; t < bt.TL; t++) s += tiles[(size_t)t * TREC + idx];
    int blk = idx / 36, rc = idx - blk * 36, r = rc / 6, c = rc - r * 6;
    int p = c_triA[blk], q = c_triB[blk];
    if (p == q && c > r) continue;
    const int ri = vis2red(bt, p, r), ci = vis2red(bt, q, c);
    if (ri < 0 || ci < 0) continue;
    S[tri(ri, ci)] += s;
  }
  __syncthreads();
  // IMU blocks: odd then even factors (consecutive factors overlap on one frame block)
  for (int par = 0; par < 2; par++) {
    for (int e = tid; e < (K - 1) * 465; e += nthr) {
      int f = e / 465, o = e - f * 465, j = f + 1;
      if ((j & 1) != par) continue;
      int a = c_tri30A[o], b2 = c_tri30B[o];
      S[tri(15 * (j - 1) + a, 15 * (j - 1) + b2)] += imo[(size_t)j * IMU_OUT + o];
    }
    __syncthreads();
  }
  // prior (already in S0 in latency mode)
  if (!S0)
    for (int e = tid; e < n * n; e += nthr) {
      int a = e / n, b2 = e - a * n;
      int ra = map[a], rb = map[b2];
      if (ra < 0 || rb < 0 || ra < rb) continue;
      S[tri(ra, rb)] += pH[(size_t)a * bt.nmax + b2];
    }
  const int* inv = bt.pr_inv + (size_t)w * np;
  // vectors
  for (int i = tid; i < np; i += nthr) {
    int p = i / 15, r = i - p * 15;
    double g1 = 0, g2 = 0, d = 0;
    const int v = red2vis(bt, i);
    if (v >= 0) {
      for (int t = 0; t < bt.TL; t++) {
        const double* rec = tiles + (size_t)t * TREC + NPb * 36;
        g2 += rec[v]; g1 += rec[K6 + v]; d += rec[2 * K6 + v];
      }
    }
    if (p + 1 < K) { const double* o = imo + (size_t)(p + 1) * IMU_OUT; g1 += o[465 + r]; g2 += o[465 + r]; d += o[tri(r, r)]; }
    if (p >= 1 && p < K) { const double* o = imo + (size_t)p * IMU_OUT; g1 += o[465 + 15 + r]; g2 += o[465 + 15 + r]; d += o[tri(15 + r, 15 + r)]; }
    { const int a = inv[i]; if (a >= 0 && a < n) { g1 += po[a]; g2 += po[a]; d += pH[(size_t)a * bt.nmax + a]; } }
    bp[i] = g1; gr[i] = g2; dH[i] = d;
  }
  __syncthreads();
  if (tid == 0) ctrl->stamps[1] = global_ns();
  // gradient max-norm, cost on the first pass
  double m = 0;
  for (int i = tid; i < np; i += nthr) m = fmax(m, fabs(bp[i]));
  for (int t = tid; t < bt.TL; t += nthr) m = fmax(m, tiles[(size_t)t * TREC + REC + 1]);
  m = block_max(m, red);
  const int first = ctrl->first;
  const double radius = ctrl->radius;
  __syncthreads();
  if (first || bt.undamped) {
    for (int i = tid; i < np; i += nthr) bt.scale_p[(size_t)w * np + i] = bt.jacobi_scaling ? 1.0 / (1.0 + sqrt(dH[i])) : 1.0;
  }
  if (tid == 0) {
    ctrl->gmax = m;
    if (first) {
      double c = 0;
      for (int t = 0; t < bt.TL; t++) c += tiles[(size_t)t * TREC + REC];
      for (int j = 1; j < K; j++) c += imo[(size_t)j * IMU_OUT + 495];
      c += po[bt.nmax];
      ctrl->cost = c; ctrl->initial_cost = c; ctrl->first = 0;
    }
  }
  if (bt.undamped) {   // debug path: dump the undamped reduced system
    __syncthreads();
    if (bt.dbg_S) {
      double* dS = bt.dbg_S + (size_t)w * np * np;
      for (int e = tid; e < np * np; e += nthr) {
        int i = e / np, j = e - i * np;
        dS[e] = (i >= j) ? S[tri(i, j)] : S[tri(j, i)];
      }
      for (int i = tid; i < np; i += nthr) bt.dbg_g[(size_t)w * np + i] = gr[i];
    }
    return;
  }
  // termination tests that need the fresh gradient (oracle_optimize order)
  int stop = 0;
  if (m <= bt.gradient_tolerance) stop = 2;            // BVIO_TERM_GRADIENT_TOL
  else if (ctrl->iterations >= bt.max_iters || !with_step) stop = -1;   // MAX_ITERS (termination already 0)
  if (stop) {
    __syncthreads();
    if (tid == 0) { ctrl->done = 1; if (stop > 0) ctrl->termination = stop; ctrl->stepped = 0; }
    return;
  }
  const int dogleg = bt.strategy;
  const double dfac = damp_factor(radius, ctrl->mu, dogleg);
  auto scale2 = [&](int i) {
    const double sp = first ? (bt.jacobi_scaling ? 1.0 / (1.0 + sqrt(dH[i])) : 1.0) : bt.scale_p[(size_t)w * np + i];
    return sp * sp;
  };
  // dogleg (Ceres DoglegStrategy::ComputeGradient / ComputeCauchyPoint): t = scale^2 g / D^2 is the
  // scaled steepest-descent direction in x space; pose parts of |g_y|^2 = g^T t and t^T H t (S still holds
  // the undamped reduced matrix, the landmark kernel adds the rest of H)
  double dg_gsq = 0, dg_tSt = 0;
  if (dogleg) {
    for (int i = tid; i < np; i += nthr) {
      const double sp2 = scale2(i);
      tv[i] = sp2 * bp[i] / fmin(fmax(sp2 * dH[i], 1e-6), 1e32);
    }
    __syncthreads();
    double a = 0, q = 0;
    for (int i = tid; i < np; i += nthr) {
      double sacc = 0;
      const double* row = S + tri(i, 0);
      for (int j = 0; j < i; j++) sacc += row[j] * tv[j];
      q += tv[i] * (2.0 * sacc + row[i] * tv[i]);
      a += bp[i] * tv[i];
    }
    dg_gsq = block_sum(a, red);
    dg_tSt = block_sum(q, red);
    __syncthreads();
  }
  // damping + augmented row
  for (int i = tid; i < np; i += nthr) {
    double sp2 = scale2(i);
    double d = damp_term(fmin(fmax(sp2 * dH[i], 1e-6), 1e32), sp2, dfac);
    ddp[i] = d;
    S[tri(i, i)] += d;
    S[tri(np, i)] = -gr[i];
  }
  __syncthreads();
  if (tid == 0) ctrl->stamps[2] = global_ns();
  // blocked Cholesky (15-wide panels) of the augmented matrix: the last row becomes y = L^-1 (-g)
  for (int kb = 0; kb < NB; kb++) {
    const int j0 = 15 * kb;
    const int bw = EX ? min(15, np - j0) : 15;
    if (tid < 32) {
      // bw x bw diagonal block in registers: lane i holds row i, column j broadcast by shuffles
      double a[15];
#pragma unroll
      for (int k = 0; k < 15; k++) a[k] = (tid < bw && k <= tid) ? S[tri(j0 + tid, j0 + k)] : 0.0;
      bool bad = false;
#pragma unroll
      for (int j = 0; j < 15; j++) {
        double pj = __shfl_sync(0xffffffffu, a[j], j);
        if (EX && j >= bw) pj = 1.0;
        bad |= !(pj > 0.0);
        const double inv = rsqrt(pj);
        if (tid == j && j < bw) dinv[j0 + j] = inv;   // 1 / L_jj for the panel and the back substitution
        a[j] *= inv;
#pragma unroll
        for (int k = j + 1; k < 15; k++) {
          const double lkj = __shfl_sync(0xffffffffu, a[j], k);
          a[k] = fma(-a[j], lkj, a[k]);
        }
      }
      if (tid < bw) {
#pragma unroll
        for (int k = 0; k < 15; k++) if (k <= tid) S[tri(j0 + tid, j0 + k)] = a[k];
      }
      if (bad && tid == 0) s_fail = 1;
    }
    __syncthreads();
    if (s_fail) break;
    // panel: rows below the diagonal block (incl. the augmented row)
    for (int i = j0 + bw + tid; i <= np; i += nthr) {
      // right-looking: as soon as x_q is known it is eliminated from every later column -- a dependency chain of two
      // operations per column instead of a running dot product
      double x[15];
      double* row = S + tri(i, j0);
#pragma unroll
      for (int c = 0; c < 15; c++) x[c] = (EX && c >= bw) ? 0.0 : row[c];
#pragma unroll
      for (int q = 0; q < 15; q++) {
        if (EX && q >= bw) break;
        x[q] *= dinv[j0 + q];
#pragma unroll
        for (int c = q + 1; c < 15; c++)
          if (!EX || c < bw) x[c] = fma(-x[q], S[tri(j0 + c, j0 + q)], x[c]);
      }
#pragma unroll
      for (int c = 0; c < 15; c++) if (!EX || c < bw) row[c] = x[c];
    }
    __syncthreads();
    // trailing update
    const int r0 = j0 + bw, mrows = np + 1 - r0;
    // 4 x 4 register tiles of the lower triangle: 120 shared loads feed 240 FMAs
    const int mt = (mrows + 3) >> 2, ntile = mt * (mt + 1) / 2;
    for (int tix = tid; tix < ntile; tix += nthr) {
      int ti = (int)((sqrtf(8.0f * (float)tix + 1.0f) - 1.0f) * 0.5f);
      while (ti * (ti + 1) / 2 > tix) ti--;
      while ((ti + 1) * (ti + 2) / 2 <= tix) ti++;
      const int tk = tix - ti * (ti + 1) / 2;
      const int i0 = r0 + 4 * ti, k0 = r0 + 4 * tk;
      const double* ri[4];
      const double* rk[4];
#pragma unroll
      for (int a = 0; a < 4; a++) { ri[a] = S + tri(min(i0 + a, np), j0); rk[a] = S + tri(min(k0 + a, np), j0); }
      double acc[4][4];
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b] = 0.0;
#pragma unroll
      for (int c = 0; c < 15; c++) {
        if (EX && c >= bw) break;
        double av[4], bv[4];
#pragma unroll
        for (int a = 0; a < 4; a++) { av[a] = ri[a][c]; bv[a] = rk[a][c]; }
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
          for (int b = 0; b < 4; b++) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
      }
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) {
          const int i = i0 + a, k = k0 + b;
          if (i <= np && k <= i) S[tri(i, k)] -= acc[a][b];
        }
    }
    __syncthreads();
  }
  if (s_fail) {
    if (tid == 0) { ctrl->solve_ok = 0; ctrl->stepped = 1; ctrl->iterations++; }
    return;
  }
  if (tid == 0) ctrl->stamps[3] = global_ns();
  // back substitution L^T dp = y (warp 0)
  for (int i = tid; i < np; i += nthr) yv[i] = S[tri(np, i)];
  __syncthreads();
  // L^T x = y, column oriented, in ONE warp and without a barrier: lane l keeps y_i for i = l, l + 32, ... in
  // registers; for k = np-1 .. 0: x_k = y_k / L_kk is broadcast by a shuffle and every lane eliminates it from its
  // rows above (row k of the packed L is contiguous: coalesced shared-memory reads that do not depend on the chain).
  if (tid < 32) {
    constexpr int YS = 8;                          // 8 * 32 = 256 >= np (np <= 226)
    double y[YS];
#pragma unroll
    for (int u = 0; u < YS; u++) y[u] = (tid + 32 * u < np) ? yv[tid + 32 * u] : 0.0;
#pragma unroll
    for (int u = YS - 1; u >= 0; u--) {
      if (32 * u >= np) continue;
      for (int kk = min(31, np - 1 - 32 * u); kk >= 0; kk--) {
        const int k = 32 * u + kk;
        const double xk = __shfl_sync(0xffffffffu, y[u] * dinv[k], kk);
        if (tid == kk) y[u] = xk;
        const double* Lk = S + tri(k, 0);
#pragma unroll
        for (int v = 0; v <= u; v++) {
          const int i = tid + 32 * v;
          if (i < k) y[v] = fma(-Lk[i], xk, y[v]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < YS; u++) if (tid + 32 * u < np) yv[tid + 32 * u] = y[u];
  }
  __syncthreads();
  if (tid == 0) ctrl->stamps[4] = global_ns();
  // step + pose part of the model cost change:  -g^T d - 1/2 d^T H d = -1/2 g^T d + 1/2 d^T D d
  double mp = 0, s_bgn = 0, s_gn2 = 0, s_nDn = 0, s_tDn = 0;
  for (int i = tid; i < np; i += nthr) {
    double d = yv[i];
    bt.delta_p[(size_t)w * np + i] = d;
    mp += -0.5 * bp[i] * d + 0.5 * ddp[i] * d * d;
    if (dogleg) {
      const double sp2 = scale2(i), cl = fmin(fmax(sp2 * dH[i], 1e-6), 1e32);
      bt.dog_t[(size_t)w * np + i] = tv[i];
      s_bgn += bp[i] * d; s_gn2 += cl * d * d / sp2; s_nDn += ddp[i] * d * d; s_tDn += ddp[i] * tv[i] * d;
    }
  }
  mp = block_sum(mp, red);
  if (dogleg) {
    s_bgn = block_sum(s_bgn, red); s_gn2 = block_sum(s_gn2, red);
    s_nDn = block_sum(s_nDn, red); s_tDn = block_sum(s_tDn, red);
    if (tid == 0) {
      ctrl->dsum[0] = dg_gsq; ctrl->dsum[1] = dg_tSt; ctrl->dsum[2] = s_gn2;
      ctrl->dsum[3] = s_bgn; ctrl->dsum[4] = s_nDn; ctrl->dsum[5] = s_tDn;
    }
  }
  if (tid == 0) { ctrl->model_pose = mp; ctrl->solve_ok = 1; ctrl->stepped = 1; ctrl->iterations++; ctrl->stamps[5] = global_ns(); }
}

// =============================================================================================
// cost at the candidate + accept/reject
// =============================================================================================
__device__ void decide_step(const BaBatch& bt, int w, double cv, double ml, double s2, double x2);
// Ceres checks max_solver_time_in_seconds after every iteration (TrustRegionMinimizer::
// FinalizeIterationAndCheckIfMinimizerCanContinue); the reference sets it to SOLVER_TIME, or 4/5 of it when the oldest
// frame is about to be marginalized (estimator.cpp:799-806).  Here the clock is the device's: time since the reset
// kernel of this solve (the H2D copy before it is not counted).
__device__ void decide(const BaBatch& bt, int w, double cv, double ml, double s2, double x2) {
  decide_step(bt, w, cv, ml, s2, x2);
  BaCtrl* c = bt.ctrl + w;
  if (bt.max_time_s > 0.0 && !c->done && (double)(global_ns() - c->t0_ns) * 1e-9 >= bt.max_time_s) {
    c->done = 1; c->termination = 5;   // BVIO_TERM_TIME
  }
}
__device__ void decide_step(const BaBatch& bt, int w, double cv, double ml, double s2, double x2) {
  BaCtrl* c = bt.ctrl + w;
  c->stepped = 0;
  c->ticket = 0;
  bool valid = c->solve_ok != 0;
  double model = c->model_pose + ml;
  if (valid && !(model > 0)) valid = false;
  const int dogleg = bt.strategy;
  if (!valid) {
    c->rejected++;
    if (++c->invalid_run >= 5) { c->done = 1; c->termination = 4; return; }
    if (dogleg) c->mu *= 10.0;      // DoglegStrategy::StepIsInvalid: mu_ *= mu_increase_factor_
    else { c->radius /= c->decrease_factor; c->decrease_factor *= 2; }
    if (c->radius < 1e-32) { c->done = 1; c->termination = 4; }
    return;
  }
  c->invalid_run = 0;
  c->cand_cost = cv;
  double step_norm = sqrt(s2), x_norm = sqrt(x2);
  if (step_norm <= bt.parameter_tolerance * (x_norm + bt.parameter_tolerance)) { c->done = 1; c->termination = 3; return; }
  double cost_change = c->cost - cv;
  if (fabs(cost_change) <= bt.function_tolerance * c->cost) { c->done = 1; c->termination = 1; return; }
  double rho = cost_change / model;
  c->rho = rho;
  if (isfinite(cv) && rho > bt.min_relative_decrease) {
    c->accepted++;
    c->cur ^= 1;
    c->cost = cv;
    if (dogleg) {                   // DoglegStrategy::StepAccepted
      if (rho < 0.25) c->radius *= 0.5;
      if (rho > 0.75) c->radius = fmax(c->radius, 3.0 * c->step_norm);
      c->mu = fmax(1e-8, 2.0 * c->mu / 10.0);
    } else {
      double t = 2.0 * rho - 1.0;
      c->radius = fmin(1e16, c->radius / fmax(1.0 / 3.0, 1.0 - t * t * t));
      c->decrease_factor = 2.0;
    }
  } else {
    c->rejected++;
    if (dogleg) c->radius *= 0.5;   // DoglegStrategy::StepRejected
    else { c->radius /= c->decrease_factor; c->decrease_factor *= 2; }
    if (c->radius < 1e-32) { c->done = 1; c->termination = 4; }
  }
}

// Traditional dogleg (Ceres DoglegStrategy::ComputeStep, TRADITIONAL_DOGLEG), landmark part + combination.
// ba_solve left the Gauss-Newton step n_p (mu-regularised) and the scaled gradient direction t_p; here 16 lanes
// per landmark back-substitute n_l = -(b + w.n_p)/(h + d), form t_l = s^2 b / D^2 and the landmark parts of
//   |g_y|^2 = g.t,  t^T H t,  |n_y|^2,  g.n,  n^T D n,  t^T D n          (D = mu * Jacobi-scaled diagonal)
// with  t^T H t = t_p^T S t_p + sum_l [(w.t_p)^2/(h+d) + 2 t_l (w.t_p) + h t_l^2]   (S = reduced matrix).
// The last CTA of the window (ticket) adds the pose parts, picks the dogleg coefficients
//   step = ca * t + cb * n      (Gauss-Newton / Cauchy / interpolated, in Ceres' scaled y space)
// and the model decrease, using H n = -g - D n:  n^T H n = -g.n - n^T D n,  t^T H n = -g.t - t^T D n.
__global__ void __launch_bounds__(BA_THREADS) ba_dogleg_kernel(BaBatch bt) {
  extern __shared__ double sm[];
  const int w = blockIdx.y, t = blockIdx.x, np = bt.np;
  BaCtrl* ctrl = bt.ctrl + w;
  if (ctrl->done || !ctrl->stepped || !ctrl->solve_ok) return;
  double* sn = sm;             // [np] Gauss-Newton step (pose part)
  double* st = sn + np;        // [np] t (pose part)
  double* red = st + np;       // [32 * DOG_REC]
  __shared__ int s_last;
  for (int i = threadIdx.x; i < np; i += blockDim.x) {
    sn[i] = bt.delta_p[(size_t)w * np + i];
    st[i] = bt.dog_t[(size_t)w * np + i];
  }
  __syncthreads();
  // 8 lanes per landmark, lane = observation (a second trip for tracks longer than 8): 4 landmarks per warp
  const int grp = threadIdx.x >> 3, l16 = threadIdx.x & 7, NG = blockDim.x >> 3;
  const unsigned gmask = 0xffffffffu;   // the four groups of a warp always iterate together (warp-uniform trip count)
  const double mu = ctrl->mu;
  int l0, l1;
  tile_range(bt, w, t, l0, l1, bt.T);
  double a[6] = {0, 0, 0, 0, 0, 0};
  for (int lw = l0 + (grp & ~3); lw < l1; lw += NG) {
    const bool act = lw + (grp & 3) < l1;
    const int l = act ? lw + (grp & 3) : l1 - 1;      // idle groups shadow a valid landmark, results dropped
    const int o0 = bt.lm_off[l], n = bt.lm_off[l + 1] - o0;
    double wn = 0, wt = 0;
#pragma unroll
    for (int trip = 0; trip < 2; trip++) {
      const int ko = l16 + 8 * trip;
      if (ko < n) {
        const int fr = bt.obs_frame[o0 + ko];
        const double2* wp = reinterpret_cast<const double2*>(bt.w + (size_t)(o0 + ko) * 6);   // 48-byte rows: three 16-byte loads
#pragma unroll
        for (int k = 0; k < 3; k++) {
          const double2 v = wp[k];
          wn += v.x * sn[15 * fr + 2 * k]; wt += v.x * st[15 * fr + 2 * k];
          wn += v.y * sn[15 * fr + 2 * k + 1]; wt += v.y * st[15 * fr + 2 * k + 1];
        }
      }
    }
    if (l16 == 0 && (bt.est_ex | bt.est_td)) {      // extra blocks: extrinsic (6) and / or td (column 0 of its slot)
      const int XB = bt.est_ex + bt.est_td;
      const double* wp = bt.wex + (size_t)l * XB * 6;
      if (bt.est_ex) {
#pragma unroll
        for (int k = 0; k < 6; k++) { wn += wp[k] * sn[15 * bt.K + k]; wt += wp[k] * st[15 * bt.K + k]; }
      }
      if (bt.est_td) { const int o = 15 * bt.K + 6 * bt.est_ex; wn += wp[6 * bt.est_ex] * sn[o]; wt += wp[6 * bt.est_ex] * st[o]; }
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) { wn += __shfl_xor_sync(gmask, wn, o, 8); wt += __shfl_xor_sync(gmask, wt, o, 8); }
    if (l16 == 0 && act) {
      const double h = bt.h[l], b = bt.b[l];
      const double sl2 = bt.jacobi_scaling ? bt.sl2[l] : 1.0;
      const double cl = fmin(fmax(sl2 * h, 1e-6), 1e32);
      const double ddl = mu * cl / sl2, hd = h + ddl;
      const double tl = sl2 * b / cl;
      const double nl = -(b + wn) / hd;
      bt.dog_l[(size_t)2 * l] = tl; bt.dog_l[(size_t)2 * l + 1] = nl;
      a[0] += b * tl;
      a[1] += wt * wt / hd + 2.0 * tl * wt + h * tl * tl;
      a[2] += cl * nl * nl / sl2;
      a[3] += b * nl;
      a[4] += ddl * nl * nl;
      a[5] += ddl * tl * nl;
    }
  }
  if (l16 == 0) {
#pragma unroll
    for (int k = 0; k < 6; k++) red[grp * DOG_REC + k] = a[k];
  }
  __syncthreads();
  double* out = bt.dog_out + (size_t)(w * bt.T + t) * DOG_REC;
  if (threadIdx.x < 6) {
    double s = 0;
    for (int g = 0; g < NG; g++) s += red[g * DOG_REC + threadIdx.x];
    out[threadIdx.x] = s;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned old = atomicAdd(&ctrl->ticket, 1u);
    s_last = (old == (unsigned)(bt.T - 1));
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    double v[6];
    for (int k = 0; k < 6; k++) v[k] = ctrl->dsum[k];
    const double* rec = bt.dog_out + (size_t)w * bt.T * DOG_REC;
    for (int q = 0; q < bt.T; q++)
      for (int k = 0; k < 6; k++) v[k] += __ldcg(rec + q * DOG_REC + k);
    const double gsq = v[0], tHt = v[1], gn2 = v[2], gdn = v[3], nDn = v[4], tDn = v[5];
    const double radius = ctrl->radius;
    const double alpha = gsq / tHt;                       // Cauchy step length: |g|^2 / |J g|^2 (y space)
    const double g_norm = sqrt(gsq), gn_norm = sqrt(gn2);
    double ca, cb, sn2;
    if (gn_norm <= radius) { ca = 0; cb = 1; sn2 = gn2; }
    else if (g_norm * alpha >= radius) { ca = -(radius / g_norm); cb = 0; sn2 = radius * radius; }
    else {
      // y-space dot product gradient . gn equals g . n in x space
      const double b_dot_a = -alpha * gdn;
      const double a_sq = alpha * alpha * gsq;
      const double bma_sq = a_sq - 2 * b_dot_a + gn2;
      const double c = b_dot_a - a_sq;
      const double d = sqrt(c * c + bma_sq * (radius * radius - a_sq));
      const double beta = (c <= 0) ? (d - c) / bma_sq : (radius * radius - a_sq) / (d + c);
      ca = -alpha * (1.0 - beta); cb = beta;
      sn2 = ca * ca * gsq + 2.0 * ca * cb * gdn + cb * cb * gn2;
    }
    const double tHn = -gsq - tDn, nHn = -gdn - nDn;
    const double quad = ca * ca * tHt + 2.0 * ca * cb * tHn + cb * cb * nHn;
    ctrl->ca = ca; ctrl->cb = cb;
    ctrl->step_norm = sqrt(fmax(sn2, 0.0));
    ctrl->model_pose = -(ca * gsq + cb * gdn) - 0.5 * quad;   // the whole model decrease (cost kernel adds 0)
    ctrl->ticket = 0;
  }
}

// MINB: CTAs per SM the register allocation aims at -- 4 (64 registers, some spills) for throughput batches, where the
// kernel is latency-bound and twice the warps hide it (0.40 -> 0.30 ms per pass of 592 windows); 2 (128 registers, no
// spills) when there are fewer CTAs than SMs anyway (latency mode)
template <int MINB>
__global__ void __launch_bounds__(BA_THREADS, MINB) ba_cost_kernel(BaBatch bt) {
  extern __shared__ double sm[];
  const int w = blockIdx.y, t = blockIdx.x, K = bt.K, np = bt.np;
  BaCtrl* ctrl = bt.ctrl + w;
  if (ctrl->done || !ctrl->stepped) return;
  if (!ctrl->solve_ok) {
    if (t == 0 && threadIdx.x == 0) decide(bt, w, 0, 0, 0, 0);
    return;
  }
  const int cur = ctrl->cur, nxt = cur ^ 1;
  double* sdp = sm;                 // [np]
  double* sPose = sdp + np;         // [(K+1)*7] candidate poses, then the (candidate) extrinsic pose
  double* sFr = sPose + (K + 1) * 7 + 1;   // [K*FR]; sPose[(K+1)*7] = candidate td
  double* sEx = sFr + K * FR;       // [FR]
  double* red = sEx + FR;           // [32*4 + 32]
  double* extra = red + 160;        // IMU/prior CTA scratch
  __shared__ int s_last;
  const int dogleg = bt.strategy;
  const double ca = ctrl->ca, cb = ctrl->cb, mu = ctrl->mu;
  for (int i = threadIdx.x; i < np; i += blockDim.x) {
    double d = bt.delta_p[(size_t)w * np + i];
    if (dogleg) d = ca * bt.dog_t[(size_t)w * np + i] + cb * d;
    sdp[i] = d;
  }
  __syncthreads();
  for (int k = threadIdx.x; k <= K; k += blockDim.x) {
    if (k < K) pose_plus(bt.pose[cur] + (size_t)(w * K + k) * 7, sdp + 15 * k, sPose + k * 7);
    else if (bt.est_ex) pose_plus(bt.exs[cur] + (size_t)w * 7, sdp + 15 * K, sPose + K * 7);
    else for (int i = 0; i < 7; i++) sPose[K * 7 + i] = bt.exs[cur][(size_t)w * 7 + i];
    if (k == K) sPose[(K + 1) * 7] = bt.tds[cur][w] + (bt.est_td ? sdp[15 * K + 6 * bt.est_ex] : 0.0);
  }
  __syncthreads();
  double* co = bt.cost_out + (size_t)(w * (bt.T + 1) + t) * COST_REC;
  if (t < bt.T) {
    for (int k = threadIdx.x; k <= K; k += blockDim.x) {
      const double* p = sPose + k * 7;
      double* o = (k < K) ? sFr + k * FR : sEx;
      qmat(q4{p[3], p[4], p[5], p[6]}, o);
      o[9] = p[0]; o[10] = p[1]; o[11] = p[2];
    }
    __syncthreads();
    // 8 lanes per landmark, lane = observation (a second trip for tracks longer than 8): 4 landmarks per warp
    const int grp = threadIdx.x >> 3, l16 = threadIdx.x & 7, NG = blockDim.x >> 3;
    const unsigned gmask = 0xffffffffu;   // the four groups of a warp always iterate together (warp-uniform trip count)
    const double dfac = damp_factor(ctrl->radius, mu, dogleg);
    int l0, l1;
    tile_range(bt, w, t, l0, l1, bt.T);
    double a_cost = 0, a_model = 0, a_s2 = 0, a_x2 = 0;
    for (int lw = l0 + (grp & ~3); lw < l1; lw += NG) {
      const bool act = lw + (grp & 3) < l1;
      const int l = act ? lw + (grp & 3) : l1 - 1;    // idle groups shadow a valid landmark, results dropped
      const int o0 = bt.lm_off[l], n = bt.lm_off[l + 1] - o0;
      const double lam = bt.invd[cur][l], h = bt.h[l], b = bt.b[l];
      const int fi = bt.obs_frame[o0];
      int fr0 = 0, fr1 = 0;
      double part = 0;
      if (l16 < n) fr0 = bt.obs_frame[o0 + l16];
      if (l16 + 8 < n) fr1 = bt.obs_frame[o0 + l16 + 8];
      if (!dogleg) {
        if (l16 < n) {
          const double2* wp = reinterpret_cast<const double2*>(bt.w + (size_t)(o0 + l16) * 6);   // 48-byte rows: three 16-byte loads
          const double* d = sdp + 15 * fr0;
#pragma unroll
          for (int k = 0; k < 3; k++) { const double2 v = wp[k]; part += v.x * d[2 * k]; part += v.y * d[2 * k + 1]; }
        }
        if (l16 + 8 < n) {
          const double2* wp = reinterpret_cast<const double2*>(bt.w + (size_t)(o0 + l16 + 8) * 6);
          const double* d = sdp + 15 * fr1;
#pragma unroll
          for (int k = 0; k < 3; k++) { const double2 v = wp[k]; part += v.x * d[2 * k]; part += v.y * d[2 * k + 1]; }
        }
        if (l16 == 0 && (bt.est_ex | bt.est_td)) {
          const double* we = bt.wex + (size_t)l * (bt.est_ex + bt.est_td) * 6;
          if (bt.est_ex) {
#pragma unroll
            for (int k = 0; k < 6; k++) part += we[k] * sdp[15 * K + k];
          }
          if (bt.est_td) part += we[6 * bt.est_ex] * sdp[15 * K + 6 * bt.est_ex];
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) part += __shfl_xor_sync(gmask, part, o, 8);
      }
      double sl2 = bt.jacobi_scaling ? bt.sl2[l] : 1.0;
      double ddl = damp_term(fmin(fmax(sl2 * h, 1e-6), 1e32), sl2, dfac);
      double dl = -(b + part) / (h + ddl);
      if (dogleg) dl = ca * bt.dog_l[(size_t)2 * l] + cb * bt.dog_l[(size_t)2 * l + 1];
      double lamc = lam + dl;
      double fc = 0;
#pragma unroll
      for (int trip = 0; trip < 2; trip++) {
        const int ko = l16 + 8 * trip, myfr = trip ? fr1 : fr0;
        if (ko >= 1 && ko < n) {
          double2 pi = bt.obs_xy[o0], pj = bt.obs_xy[o0 + ko];
          if (bt.est_td) {
            const double tdc = sPose[(K + 1) * 7], si_ = tdc + bt.obs_shift[o0], sj_ = tdc + bt.obs_shift[o0 + ko];
            const double2 vi = bt.obs_vel[o0], vj = bt.obs_vel[o0 + ko];
            pi.x -= si_ * vi.x; pi.y -= si_ * vi.y; pj.x -= sj_ * vj.x; pj.y -= sj_ * vj.y;
          }
          ProjGeom g = proj_geom(sFr + fi * FR, sFr + myfr * FR, sEx, pi.x, pi.y, lamc);
          double inv = 1.0 / g.pcj.z;
          double r0 = bt.sqrt_info * (g.pcj.x * inv - pj.x), r1 = bt.sqrt_info * (g.pcj.y * inv - pj.y);
          double rho0, rho1;
          cauchy(bt.cauchy_a, r0 * r0 + r1 * r1, rho0, rho1);
          fc += 0.5 * rho0;
        }
      }
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) fc += __shfl_xor_sync(gmask, fc, o, 8);
      if (l16 == 0 && act) {
        bt.invd[nxt][l] = lamc;
        a_cost += fc;
        if (!dogleg) a_model += -0.5 * b * dl + 0.5 * ddl * dl * dl;
        a_s2 += dl * dl;
        a_x2 += lam * lam;
      }
    }
    if (l16 == 0) { red[grp * 4] = a_cost; red[grp * 4 + 1] = a_model; red[grp * 4 + 2] = a_s2; red[grp * 4 + 3] = a_x2; }
    __syncthreads();
    if (threadIdx.x < 4) {
      double s = 0;
      for (int g = 0; g < NG; g++) s += red[g * 4 + threadIdx.x];
      co[threadIdx.x] = s;
    }
  } else {
    // candidate poses / speed-biases to the other buffer, IMU + prior cost, pose part of the norms
    double* sSb = extra;                   // [K*9]
    double* sR = sSb + K * 9;              // [(K-1)*15]
    double* sdx = sR + (K - 1) * 15;       // [nmax]
    double* spr = sdx + bt.nmax;           // [nmax]
    double* pn = bt.pose[nxt] + (size_t)w * K * 7;
    double* sn = bt.sb[nxt] + (size_t)w * K * 9;
    const double* pc = bt.pose[cur] + (size_t)w * K * 7;
    const double* sc = bt.sb[cur] + (size_t)w * K * 9;
    double s2 = 0, x2 = 0;
    for (int i = threadIdx.x; i < K * 7; i += blockDim.x) {
      double v = sPose[i], o = pc[i];
      pn[i] = v; s2 += (o - v) * (o - v); x2 += o * o;
    }
    if (bt.est_ex && threadIdx.x < 7) {
      double v = sPose[K * 7 + threadIdx.x], o = bt.exs[cur][(size_t)w * 7 + threadIdx.x];
      bt.exs[nxt][(size_t)w * 7 + threadIdx.x] = v; s2 += (o - v) * (o - v); x2 += o * o;
    }
    if (bt.est_td && threadIdx.x == 32) {
      const double v = sPose[(K + 1) * 7], o = bt.tds[cur][w];
      bt.tds[nxt][w] = v; s2 += (o - v) * (o - v); x2 += o * o;
    }
    for (int i = threadIdx.x; i < K * 9; i += blockDim.x) {
      int k = i / 9, r = i - k * 9;
      double o = sc[i], v = o + sdp[15 * k + 6 + r];
      sSb[i] = v; sn[i] = v; s2 += (o - v) * (o - v); x2 += o * o;
    }
    s2 = block_sum(s2, red);
    x2 = block_sum(x2, red + 32);
    __syncthreads();
    if (threadIdx.x < K - 1) {
      int j = threadIdx.x + 1;
      const double* rec = bt.imu + (size_t)(w * K + j) * IMU_REC;
      double* r = sR + threadIdx.x * 15;
      if (!(rec[IR_DT] > 10.0)) {
        double raw[15];
        imu_raw(rec, bt.G, sPose + (j - 1) * 7, sSb + (j - 1) * 9, sPose + j * 7, sSb + j * 9, raw, nullptr);
        const double* SI = rec + IR_SQ;
        for (int i = 0; i < 15; i++) {
          double s = 0;
          for (int k = i; k < 15; k++) s += SI[i * 15 + k] * raw[k];
          r[i] = s;
        }
      } else {
        for (int i = 0; i < 15; i++) r[i] = 0;
      }
    }
    __syncthreads();
    double c = 0;
    for (int i = threadIdx.x; i < (K - 1) * 15; i += blockDim.x) c += 0.5 * sR[i] * sR[i];
    int n = bt.pr_n[w];
    if (n > 0) {
      // prior_residual reads the state from global memory laid out [B][K][..]; the candidate was just
      // written to the nxt buffers by this CTA
      __threadfence_block();
      __syncthreads();
      prior_residual(bt, w, bt.pose[nxt], bt.sb[nxt], bt.exs[nxt], bt.tds[nxt], sdx, spr);
      for (int i = threadIdx.x; i < n; i += blockDim.x) c += 0.5 * spr[i] * spr[i];
    }
    c = block_sum(c, red);
    if (threadIdx.x == 0) { co[0] = c; co[1] = 0; co[2] = s2; co[3] = x2; }
  }
  // last CTA of this window decides
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned old = atomicAdd(&ctrl->ticket, 1u);
    s_last = (old == (unsigned)bt.T);
  }
  __syncthreads();
  if (s_last && threadIdx.x < 32) {
    // warp 0 sums the T+1 partial records with a fixed tree (T <= 32), then lane 0 decides
    __threadfence();
    const double2* co2 = reinterpret_cast<const double2*>(bt.cost_out + (size_t)w * (bt.T + 1) * COST_REC);
    double v[4] = {0, 0, 0, 0};
    for (int q = threadIdx.x; q <= bt.T; q += 32) {
      const double2 a = __ldcg(co2 + 2 * q), b = __ldcg(co2 + 2 * q + 1);
      v[0] += a.x; v[1] += a.y; v[2] += b.x; v[3] += b.y;
    }
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] = warp_sum(v[k]);
    if (threadIdx.x == 0) decide(bt, w, v[0], v[1], v[2], v[3]);
  }
}

// =============================================================================================
// launchers
// =============================================================================================
static size_t imu_prior_smem(int K, int nmax) { return sizeof(double) * ((size_t)(K - 1) * (900 + 30) + 2 * nmax); }

size_t ba_linearize_smem_bytes(int K, int CL, int XB) {
  const int FS = K - 1, KE = K + XB;
  size_t d = (size_t)(K + 1) * FR + 16 + (size_t)CL * (FS * (STG + 12 * XB) + 6 * KE + 36 + 6 + 4 + 78 * XB + (XB == 2 ? 36 : 0));
  size_t bytes = d * sizeof(double) + (size_t)CL * 2 * sizeof(int) + (size_t)CL * KE;
  return (bytes + 15) & ~size_t(15);
}
// landmarks per chunk: one (landmark, factor) slot per thread, capped so that two CTAs fit an SM
int ba_pick_chunk(int K, int XB) {
  int CL = BA_THREADS / (K - 1);
  if (CL > 64) CL = 64;
  while (CL > 1 && ba_linearize_smem_bytes(K, CL, XB) > 100 * 1024) CL--;
  return CL;
}
size_t ba_solve_smem_bytes(int np) {
  int N1 = np + 1;
  return sizeof(double) * ((size_t)N1 * (N1 + 1) / 2 + 2 * np + 32);
}
static size_t cost_smem(int K, int nmax) {
  return sizeof(double) * ((size_t)15 * K + 8 + (K + 1) * 7 + (K + 1) * FR + 160 + K * 9 + (K - 1) * 15 + 2 * nmax);
}

cudaError_t ba_configure_marginalize(void);   // below, next to the kernels
__global__ void ba_marginal9_kernel(const double* S, int np, int frame, double* out81, int* status);
// __constant__ tables and the dynamic-shared-memory opt-ins are PER DEVICE: remember which devices of this process
// have been set up (bvio_create may be called for several GPUs in one process)
static std::mutex g_cfg_mutex;
static unsigned long long g_cfg_devices[4] = {0, 0, 0, 0};
int ba_configure(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return (int)cudaGetLastError();
  std::lock_guard<std::mutex> lock(g_cfg_mutex);
  const bool known = dev >= 0 && dev < 256 && ((g_cfg_devices[dev >> 6] >> (dev & 63)) & 1ull);
  if (!known) {
    unsigned char ta[BVIO_KMAX * (BVIO_KMAX + 1) / 2], tb[BVIO_KMAX * (BVIO_KMAX + 1) / 2];
    int e = 0;
    for (int a = 0; a < BVIO_KMAX; a++) for (int b = 0; b <= a; b++) { ta[e] = a; tb[e] = b; e++; }
    unsigned char ua[465], ub[465];
    e = 0;
    for (int a = 0; a < 30; a++) for (int b = 0; b <= a; b++) { ua[e] = a; ub[e] = b; e++; }
    cudaError_t err;
    if ((err = cudaMemcpyToSymbol(c_triA, ta, sizeof ta)) != cudaSuccess) return err;
    if ((err = cudaMemcpyToSymbol(c_triB, tb, sizeof tb)) != cudaSuccess) return err;
    if ((err = cudaMemcpyToSymbol(c_tri30A, ua, sizeof ua)) != cudaSuccess) return err;
    if ((err = cudaMemcpyToSymbol(c_tri30B, ub, sizeof ub)) != cudaSuccess) return err;
#define BVIO_LIN_ATTR(NS, XB) \
    if ((err = cudaFuncSetAttribute(ba_linearize_kernel<NS, XB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024)) != cudaSuccess) return err;
    BVIO_LIN_ATTR(1, 0) BVIO_LIN_ATTR(2, 0) BVIO_LIN_ATTR(3, 0) BVIO_LIN_ATTR(4, 0)
    BVIO_LIN_ATTR(1, 1) BVIO_LIN_ATTR(2, 1) BVIO_LIN_ATTR(3, 1) BVIO_LIN_ATTR(4, 1)
    BVIO_LIN_ATTR(1, 2) BVIO_LIN_ATTR(2, 2) BVIO_LIN_ATTR(3, 2) BVIO_LIN_ATTR(4, 2)
#undef BVIO_LIN_ATTR
    if ((err = cudaFuncSetAttribute(ba_linearize_mma_kernel<6, 76>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)) != cudaSuccess) return err;
    if ((err = cudaFuncSetAttribute(ba_linearize_mma_kernel<6, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)) != cudaSuccess) return err;
    if ((err = cudaFuncSetAttribute(ba_linearize_ws_kernel<6, 76>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024)) != cudaSuccess) return err;
    if ((err = cudaFuncSetAttribute(ba_imu_prior_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)) != cudaSuccess) return err;
    if ((err = cudaFuncSetAttribute(ba_linearize_ws_kernel<6, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024)) != cudaSuccess) return err;
    if ((err = cudaFuncSetAttribute(ba_linearize_mma_kernel<10, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)) != cudaSuccess) return err;
    if ((err = cudaFuncSetAttribute(ba_solve_kernel<false, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024)) != cudaSuccess) return err;
    if ((err = cudaFuncSetAttribute(ba_solve_kernel<true, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024)) != cudaSuccess) return err;
    if ((err = cudaFuncSetAttribute(ba_solve_kernel<false, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024)) != cudaSuccess) return err;
    if ((err = cudaFuncSetAttribute(ba_solve_kernel<true, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024)) != cudaSuccess) return err;
    if ((err = cudaFuncSetAttribute(ba_cost_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024)) != cudaSuccess) return err;
    if ((err = cudaFuncSetAttribute(ba_cost_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024)) != cudaSuccess) return err;
    if ((err = ba_configure_marginalize()) != cudaSuccess) return err;
    if ((err = cudaFuncSetAttribute(ba_marginal9_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024)) != cudaSuccess) return err;
    if (dev >= 0 && dev < 256) g_cfg_devices[dev >> 6] |= 1ull << (dev & 63);
  }
  return 0;
}

int ba_launch_prepare(const BaBatch& bt, cudaStream_t st) {
  ba_prepare_kernel<<<bt.B, BA_THREADS, 0, st>>>(bt);
  return 1;
}
int ba_launch_reset(const BaBatch& bt, cudaStream_t st) {
  size_t work = (size_t)bt.B * bt.K * 9;
  if ((size_t)bt.total_L > work) work = bt.total_L;
  int blocks = (int)((work + 255) / 256);
  if (blocks > 1184) blocks = 1184;
  if (blocks < 1) blocks = 1;
  ba_reset_kernel<<<blocks, 256, 0, st>>>(bt);
  return 1;
}
int ba_launch_iteration(const BaBatch& bt, cudaStream_t st, bool with_step, cudaEvent_t* ev) {
  const int XB = bt.est_ex + bt.est_td;
  size_t s1 = ba_linearize_smem_bytes(bt.K, bt.chunk_l, XB), s1b = imu_prior_smem(bt.K, bt.nmax);
  if (s1b > s1) s1 = s1b;
  if (ev) cudaEventRecord(ev[0], st);
  const int KE = bt.K + XB;
  const int nstrip = (KE * (KE + 1) / 2) * 2;   // half-block units
  // latency mode (fewer windows than SMs): every IMU factor in its own CTA
  const dim3 grid(bt.TL + 1 + (bt.solve_wide ? bt.K - 1 : 0), bt.B);
  int nlin = 1;                                  // launches of the linearization
#define BVIO_LIN(NS) \
  { if (XB == 2) ba_linearize_kernel<NS, 2><<<grid, BA_THREADS, s1, st>>>(bt); \
    else if (XB == 1) ba_linearize_kernel<NS, 1><<<grid, BA_THREADS, s1, st>>>(bt); \
    else ba_linearize_kernel<NS, 0><<<grid, BA_THREADS, s1, st>>>(bt); }
  if (bt.use_mma) {
    size_t sm1 = ba_linearize_mma_smem_bytes(bt.K);
    if (s1b > sm1) sm1 = s1b;
    const int NT = mm_ntile(bt.K);
    const size_t smw = ba_linearize_ws_smem_bytes(bt.K);
    // throughput mode: producers / consumers overlapped in one 512-thread CTA per SM
    if (bt.use_ws && NT * (NT + 1) / 2 <= 48 && smw <= 224 * 1024) {
      const dim3 gv(bt.TL, bt.B);
      // (the IMU / prior kernel on a second stream beside the tiles was measured: no gain, the tiles' four waves end together)
      ba_imu_prior_kernel<<<bt.B, IMU_THREADS, s1b, st>>>(bt);
      if (mm_wstride(bt.K) == 76) ba_linearize_ws_kernel<6, 76><<<gv, WS_THREADS, smw, st>>>(bt);
      else ba_linearize_ws_kernel<6, 0><<<gv, WS_THREADS, smw, st>>>(bt);
      nlin = 2;
    }
    else if (NT * (NT + 1) / 2 <= 48 && mm_wstride(bt.K) == 76) ba_linearize_mma_kernel<6, 76><<<grid, BA_THREADS, sm1, st>>>(bt);
    else if (NT * (NT + 1) / 2 <= 48) ba_linearize_mma_kernel<6, 0><<<grid, BA_THREADS, sm1, st>>>(bt);
    else ba_linearize_mma_kernel<10, 0><<<grid, BA_THREADS, sm1, st>>>(bt);
  }
  else if (nstrip <= BA_THREADS - 32) BVIO_LIN(1)
  else if (nstrip <= 2 * (BA_THREADS - 32)) BVIO_LIN(2)
  else if (nstrip <= 3 * (BA_THREADS - 32)) BVIO_LIN(3)
  else BVIO_LIN(4)
#undef BVIO_LIN
  if (ev) cudaEventRecord(ev[1], st);
  const size_t ssm = ba_solve_smem_bytes(bt.np);
  if (bt.solve_wide) {
    if (XB) ba_solve_kernel<true, 512><<<bt.B, 512, ssm, st>>>(bt, with_step ? 1 : 0);
    else ba_solve_kernel<false, 512><<<bt.B, 512, ssm, st>>>(bt, with_step ? 1 : 0);
  } else {
    if (XB) ba_solve_kernel<true, 256><<<bt.B, 256, ssm, st>>>(bt, with_step ? 1 : 0);
    else ba_solve_kernel<false, 256><<<bt.B, 256, ssm, st>>>(bt, with_step ? 1 : 0);
  }
  if (!with_step || bt.undamped) { if (ev) { cudaEventRecord(ev[2], st); cudaEventRecord(ev[3], st); } return 1 + nlin; }
  int nk = 2 + nlin;
  if (bt.strategy) {   // dogleg combination; its time is booked with the reduced solve
    ba_dogleg_kernel<<<dim3(bt.T, bt.B), BA_THREADS, sizeof(double) * (2 * bt.np + 32 * DOG_REC), st>>>(bt);
    nk++;
  }
  if (ev) cudaEventRecord(ev[2], st);
  if (bt.solve_wide) ba_cost_kernel<2><<<dim3(bt.T + 1, bt.B), BA_THREADS, cost_smem(bt.K, bt.nmax), st>>>(bt);
  else ba_cost_kernel<4><<<dim3(bt.T + 1, bt.B), BA_THREADS, cost_smem(bt.K, bt.nmax), st>>>(bt);
  if (ev) cudaEventRecord(ev[3], st);
  return nk;
}
int ba_launch_finish(const BaBatch& bt, cudaStream_t st) {
  int blocks = bt.B < 1184 ? bt.B : 1184;
  ba_finish_kernel<<<blocks, 256, 0, st>>>(bt);
  return 1;
}


// =============================================================================================
// Omega_PRIOR for the selector (the report's future work: support_files/report/paper/anticipation.tex
// :146-152): the information the WHOLE window -- prior, IMU and visual factors, landmarks eliminated -- holds on
// x_k = (position, velocity, accelerometer bias) of one frame, i.e. the Schur complement of the undamped reduced matrix
// S onto those nine dimensions.  With the nine put last, the Cholesky factor of the permuted S has L_aa L_aa^T = that
// complement.  One CTA; S comes from a debug linearization (dbg_S, row-major np x np).
// =============================================================================================
__global__ void __launch_bounds__(256) ba_marginal9_kernel(const double* S, int np, int frame, double* out81, int* status) {
  extern __shared__ double sm[];
  const int tid = threadIdx.x, nt = blockDim.x;
  double* Lp = sm;                                          // packed lower np(np+1)/2
  int* perm = reinterpret_cast<int*>(Lp + (size_t)np * (np + 1) / 2);   // [np] new position -> reduced index
  __shared__ int s_bad;
  if (tid == 0) {
    const int base = 15 * frame;
    const int sel[9] = {base, base + 1, base + 2, base + 6, base + 7, base + 8, base + 9, base + 10, base + 11};
    int k = 0;
    for (int i = 0; i < np; i++) {
      bool is_a = false;
      for (int q = 0; q < 9; q++) is_a |= (sel[q] == i);
      if (!is_a) perm[k++] = i;
    }
    for (int q = 0; q < 9; q++) perm[k++] = sel[q];
    s_bad = 0;
  }
  __syncthreads();
  for (int e = tid; e < np * (np + 1) / 2; e += nt) {
    int i = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
    while (i * (i + 1) / 2 > e) i--;
    while ((i + 1) * (i + 2) / 2 <= e) i++;
    const int j = e - i * (i + 1) / 2;
    Lp[e] = 0.5 * (S[(size_t)perm[i] * np + perm[j]] + S[(size_t)perm[j] * np + perm[i]]);
  }
  __syncthreads();
  const int nb = np - 9;                                    // eliminate the first nb variables
  for (int k = 0; k < nb; k++) {
    const double piv = Lp[tri(k, k)];
    if (piv == 0.0) continue;                               // a variable nothing constrains or couples to (the unused speed-bias
                                                            // slot of a relocalization frame): its row and column are zero
    if (!(piv > 0.0)) { if (tid == 0) s_bad = 1; break; }
    const double inv = rsqrt(piv);
    __syncthreads();
    for (int i = k + tid; i < np; i += nt) Lp[tri(i, k)] *= inv;          // (the diagonal becomes sqrt(piv))
    __syncthreads();
    const int rem = np - k - 1;
    for (int e = tid; e < rem * (rem + 1) / 2; e += nt) {
      int i = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
      while (i * (i + 1) / 2 > e) i--;
      while ((i + 1) * (i + 2) / 2 <= e) i++;
      const int j = e - i * (i + 1) / 2;
      Lp[tri(k + 1 + i, k + 1 + j)] -= Lp[tri(k + 1 + i, k)] * Lp[tri(k + 1 + j, k)];
    }
    __syncthreads();
  }
  __syncthreads();
  if (tid < 81) {
    const int a = tid / 9, b2 = tid - a * 9;
    const int i = nb + (a >= b2 ? a : b2), j = nb + (a >= b2 ? b2 : a);
    out81[tid] = s_bad ? NAN : Lp[tri(i, j)];                // what is left in the trailing 9 x 9 block IS the complement
  }
  if (tid == 0) *status = s_bad;
}
int ba_launch_marginal9(const double* S, int np, int frame, double* out81, int* status, cudaStream_t st) {
  const size_t smem = sizeof(double) * ((size_t)np * (np + 1) / 2) + sizeof(int) * np + 16;
  ba_marginal9_kernel<<<1, 256, smem, st>>>(S, np, frame, out81, status);
  return 1;
}

// =============================================================================================
// marginalization (row a9): MarginalizationInfo::{preMarginalize, marginalize}
// (marginalization_factor.cpp:109-297) for the factor sets Estimator::optimization() builds at
// estimator.cpp:816-991.  One CTA.  The scalar inverse-depth blocks are eliminated analytically
// (they are mutually independent), the <= 15 remaining dropped dimensions by the reference's eigen
// pseudo-inverse (eps = 1e-8), and the kept system is factored back to (J, r) by a parallel
// cyclic-Jacobi eigen-decomposition in shared memory (stands in for Eigen::SelfAdjointEigenSolver).
// =============================================================================================
struct MargArgs {
  int flag;             // 0 MARGIN_OLD, 1 MARGIN_SECOND_NEW
  int M;                // 15K + 7: frame-major 15-blocks, then the extrinsic block, then td
  int m, n, ne;         // dropped / kept dimensions, ne = n rounded up to even
  int n0;               // landmarks anchored at frame 0 (the first n0 of the window in device order)
  int G;                // CTAs of the factor kernel that hold visual partial sums (the IMU CTA comes after them)
  const int* dropidx;   // [m] indices into the M layout
  const int* keepidx;   // [n]
  double* A;            // [M*M] scratch
  double* b;            // [M]
  double* part;         // [G][VD*VD + VD] visual partial sums, then [930] IMU J^T J (30x30) and J^T r (30)
  double* out_jac;      // [n*n] column-major linearized_jacobians
  double* out_res;      // [n]
  int* status;          // [4] sweeps used (eigen) / rank (cholesky), pivot range; then [8] u64 phase time stamps (ns)
  int method;           // 0: eigen-decomposition like the reference, 1: pivoted Cholesky (same J^T J, J^T r)
};

constexpr int MARG_THREADS = 1024;  // the elimination / eigen-decomposition CTA
constexpr int MARG_FT = 256;        // threads of a factor CTA
constexpr int MARG_NB = 4;          // landmarks a factor CTA evaluates together
constexpr int MF = 42;              // per-factor staging: A(12) B(12) E(12) c(2) r(2) td(2)

// column `d` (visual layout: 6 dims per frame, then 6 extrinsic dims, then td) of factor f's 2 x . Jacobian
__device__ __forceinline__ void marg_col(const double* st, int fj, int K, int d, double& j0, double& j1) {
  j0 = 0; j1 = 0;
  if (d < 6) { j0 = st[d]; j1 = st[6 + d]; }
  else if (d < 6 * K) { if (d >= 6 * fj && d < 6 * fj + 6) { j0 = st[12 + d - 6 * fj]; j1 = st[18 + d - 6 * fj]; } }
  else if (d < 6 * K + 6) { j0 = st[24 + d - 6 * K]; j1 = st[30 + d - 6 * K]; }
  else { j0 = st[40]; j1 = st[41]; }
}
// LOCAL column a of a factor: 0..5 pose of frame 0, 6..11 pose of frame fj, 12..17 extrinsics, 18 td
__device__ __forceinline__ void marg_lcol(const double* st, int a, double& j0, double& j1) {
  if (a < 18) { const int blk = a / 6, r = a - 6 * blk; j0 = st[12 * blk + r]; j1 = st[12 * blk + 6 + r]; }
  else { j0 = st[40]; j1 = st[41]; }
}
__device__ __forceinline__ int marg_gcol(int a, int fj, int K) {
  return a < 6 ? a : (a < 12 ? 6 * fj + a - 6 : 6 * K + a - 12);     // a == 18 -> 6K + 6
}

// residual + Jacobians of one visual factor anchored at frame 0, loss-corrected: 42 doubles
__device__ void marg_eval_factor(const BaBatch& bt, int w, int l, int o0, int f, const double* sFr, const double* sEx, double* st, int* fj_out) {
  const int fj = bt.obs_frame[o0 + 1 + f];
  double2 pi = bt.obs_xy[o0], pj = bt.obs_xy[o0 + 1 + f];
  double2 vi = {0, 0}, vj = {0, 0};
  if (bt.est_td) {   // ProjectionTdFactor: time-shifted points (projection_td_factor.cpp:51-52)
    vi = bt.obs_vel[o0]; vj = bt.obs_vel[o0 + 1 + f];
    const double td = bt.td0[w], si_ = td + bt.obs_shift[o0], sj_ = td + bt.obs_shift[o0 + 1 + f];
    pi.x -= si_ * vi.x; pi.y -= si_ * vi.y; pj.x -= sj_ * vj.x; pj.y -= sj_ * vj.y;
  }
  const double lam = bt.invd0[l];
  const double* Fi = sFr;
  const double* Fj = sFr + fj * FR;
  ProjGeom g = proj_geom(Fi, Fj, sEx, pi.x, pi.y, lam);
  const double inv = 1.0 / g.pcj.z, si = bt.sqrt_info;
  const double r0 = si * (g.pcj.x * inv - pj.x), r1 = si * (g.pcj.y * inv - pj.y);
  const double red[2][3] = {{si * inv, 0.0, -si * g.pcj.x * inv * inv}, {0.0, si * inv, -si * g.pcj.y * inv * inv}};
  double rho0, rho1;
  cauchy(bt.cauchy_a, r0 * r0 + r1 * r1, rho0, rho1);
  const double sr = sqrt(rho1);
  const d3 tic{sEx[9], sEx[10], sEx[11]};
  const d3 pci = (1.0 / lam) * d3{pi.x, pi.y, 1.0};
  // tmp_r = ric^T Rj^T Ri ric ; tvec = ric^T (Rj^T (Ri tic + Pi - Pj) - tic)   (projection_factor.cpp:100-105)
  double RjTRi[9], T1[9], tmp_r[9];
  mtm3(Fj, Fi, RjTRi);
  mtm3(sEx, RjTRi, T1);               // ric^T Rj^T Ri
  mm3(T1, sEx, tmp_r);
  const d3 inner = mtv3(Fj, mv3(Fi, tic) + d3{Fi[9], Fi[10], Fi[11]} - d3{Fj[9], Fj[10], Fj[11]}) - tic;
  const d3 tvec = mtv3(sEx, inner);
  const d3 trp = mv3(tmp_r, pci);
  for (int a = 0; a < 2; a++) {
    double Gm[3], Q[3];
    for (int c = 0; c < 3; c++) Gm[c] = red[a][0] * sEx[c * 3 + 0] + red[a][1] * sEx[c * 3 + 1] + red[a][2] * sEx[c * 3 + 2];
    for (int c = 0; c < 3; c++) Q[c] = Gm[0] * Fj[c * 3 + 0] + Gm[1] * Fj[c * 3 + 1] + Gm[2] * Fj[c * 3 + 2];
    const d3 u = mtv3(Fi, d3{Q[0], Q[1], Q[2]});
    const d3 jr = cross3(g.pimu_i, u);
    const d3 jjr = cross3(d3{Gm[0], Gm[1], Gm[2]}, g.pimu_j);
    st[a * 6 + 0] = sr * Q[0]; st[a * 6 + 1] = sr * Q[1]; st[a * 6 + 2] = sr * Q[2];
    st[a * 6 + 3] = sr * jr.x; st[a * 6 + 4] = sr * jr.y; st[a * 6 + 5] = sr * jr.z;
    st[12 + a * 6 + 0] = -sr * Q[0]; st[12 + a * 6 + 1] = -sr * Q[1]; st[12 + a * 6 + 2] = -sr * Q[2];
    st[12 + a * 6 + 3] = sr * jjr.x; st[12 + a * 6 + 4] = sr * jjr.y; st[12 + a * 6 + 5] = sr * jjr.z;
    // extrinsic block: reduce * [ ric^T (Rj^T Ri - I) | -tmp_r [pci]x + [tmp_r pci]x + [tvec]x ]
    const double rd[3] = {red[a][0], red[a][1], red[a][2]};
    double el[3];
    for (int c = 0; c < 3; c++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += rd[k] * (T1[k * 3 + c] - sEx[c * 3 + k]);   // ric^T Rj^T Ri - ric^T
      el[c] = s;
    }
    // row^T [v]x = (row x v)^T ;  row^T (-tmp_r [pci]x) = -((tmp_r^T row) x pci)^T
    const d3 rowv{rd[0], rd[1], rd[2]};
    const d3 t1 = mtv3(tmp_r, rowv);
    const d3 e1 = cross3(t1, pci);
    const d3 e2 = cross3(rowv, trp);
    const d3 e3 = cross3(rowv, tvec);
    st[24 + a * 6 + 0] = sr * el[0]; st[24 + a * 6 + 1] = sr * el[1]; st[24 + a * 6 + 2] = sr * el[2];
    st[24 + a * 6 + 3] = sr * (-e1.x + e2.x + e3.x); st[24 + a * 6 + 4] = sr * (-e1.y + e2.y + e3.y);
    st[24 + a * 6 + 5] = sr * (-e1.z + e2.z + e3.z);
    st[36 + a] = sr * (-dot3(u, g.pimu_i - tic) / lam);
    // td Jacobian (projection_td_factor.cpp:131-136); zero when td is not estimated
    st[40 + a] = bt.est_td ? sr * (-dot3(u, mv3(sEx, d3{vi.x, vi.y, 0.0})) / lam + si * (a == 0 ? vj.x : vj.y)) : 0.0;
  }
  st[38] = sr * r0; st[39] = sr * r1;
  *fj_out = fj;
}

// ---- kernel 1: the factors of the dropped frame, landmark-parallel over a few CTAs ----------------------------------------------
// CTA g < G: landmarks g*NB, g*NB + G*NB, ... of the n0 anchored at frame 0 (estimator.cpp:852-893): factor evaluation, analytic
// elimination of the landmark's inverse depth (its 1x1 block of the reference's pseudo-inverse, eps = 1e-8), accumulation
// of  sum_f J_f^T J_f - w w^T / h  and  sum_f J_f^T r_f - w b / h  in shared memory, one partial sum per CTA to HBM.
// CTA G: the IMU factor 0 -> 1 (estimator.cpp:841-850): raw Jacobian by one thread, whitening and J^T J by all.
__global__ void __launch_bounds__(MARG_FT) ba_marg_factors_kernel(BaBatch bt, MargArgs ma) {
  extern __shared__ double sm[];
  const int tid = threadIdx.x, nt = blockDim.x, K = bt.K, w = 0;
  const int VD = 6 * K + 6 + bt.est_td;
  if ((int)blockIdx.x == ma.G) {
    double* sJ = sm;            // [450] raw
    double* sJ2 = sJ + 450;     // [450] whitened
    double* sR = sJ2 + 450;     // [15] raw, [15] whitened
    double* out = ma.part + (size_t)ma.G * (VD * VD + VD);
    const double* rec = bt.imu + (size_t)(w * K + 1) * IMU_REC;
    const bool use = rec[IR_DT] < 10.0;
    for (int i = tid; i < 450; i += nt) sJ[i] = 0.0;
    __syncthreads();
    if (use && tid == 0) imu_raw(rec, bt.G, bt.pose0, bt.sb0, bt.pose0 + 7, bt.sb0 + 9, sR, sJ);
    __syncthreads();
    if (use) {
      const double* SI = rec + IR_SQ;
      for (int e = tid; e < 465; e += nt) {
        int i = e / 31, c = e - i * 31;
        double s = 0;
        if (c < 30) { for (int k = i; k < 15; k++) s += SI[i * 15 + k] * sJ[k * 30 + c]; sJ2[i * 30 + c] = s; }
        else { for (int k = i; k < 15; k++) s += SI[i * 15 + k] * sR[k]; sR[15 + i] = s; }
      }
    }
    __syncthreads();
    for (int e = tid; e < 930; e += nt) {
      double s = 0;
      if (use) {
        if (e < 900) { int a = e / 30, c = e - a * 30; for (int k = 0; k < 15; k++) s += sJ2[k * 30 + a] * sJ2[k * 30 + c]; }
        else { int a = e - 900; for (int k = 0; k < 15; k++) s += sJ2[k * 30 + a] * sR[15 + k]; }
      }
      out[e] = s;
    }
    return;
  }
  double* sFr = sm;                                   // [K*FR]
  double* sEx = sFr + K * FR;                         // [FR]
  double* sF = sEx + FR;                              // [NB][16][MF]
  double* sWv = sF + MARG_NB * 16 * MF;               // [NB][VD] w
  double* sG = sWv + MARG_NB * VD;                    // [NB][VD] J^T r
  double* sS = sG + MARG_NB * VD;                     // [NB][2] 1/h (thresholded), b
  double* As = sS + 2 * MARG_NB;                      // [VD*VD]
  double* gs = As + VD * VD;                          // [VD]
  int* sFj = reinterpret_cast<int*>(gs + VD);         // [NB][16] frame of factor f
  int* sFof = sFj + MARG_NB * 16;                     // [NB][16] frame -> factor index (-1: none)
  int* sNf = sFof + MARG_NB * 16;                     // [NB] factors of the landmark
  stage_frames(bt, w, bt.pose0, bt.ex, sFr, sEx);
  for (int e = tid; e < VD * VD + VD; e += nt) As[e] = 0.0;
  const int L0 = bt.lm_base[0];
  const int NC = 18 + bt.est_td;                      // local columns of a factor
  for (int base = blockIdx.x * MARG_NB; base < ma.n0; base += ma.G * MARG_NB) {
    const int nb = min(MARG_NB, ma.n0 - base);
    __syncthreads();
    if (tid < MARG_NB * 16) sFof[tid] = -1;
    __syncthreads();
    if (tid < MARG_NB * 16) {
      const int lb = tid >> 4, f = tid & 15;
      if (lb < nb) {
        const int l = L0 + base + lb, o0 = bt.lm_off[l], nfac = bt.lm_off[l + 1] - o0 - 1;
        if (f == 0) sNf[lb] = nfac;
        if (f < nfac) {
          int fj;
          marg_eval_factor(bt, w, l, o0, f, sFr, sEx, sF + (lb * 16 + f) * MF, &fj);
          sFj[lb * 16 + f] = fj;
          sFof[lb * 16 + fj] = f;
        }
      }
    }
    __syncthreads();
    // w = sum_f J_f^T c_f, J^T r, h, b of every landmark of the batch
    for (int it = tid; it < nb * (VD + 1); it += nt) {
      const int lb = it / (VD + 1), d = it - lb * (VD + 1), nfac = sNf[lb];
      const double* F = sF + lb * 16 * MF;
      if (d < VD) {
        double wsum = 0, gsum = 0;
        for (int f = 0; f < nfac; f++) {
          double j0, j1;
          const double* st = F + f * MF;
          marg_col(st, sFj[lb * 16 + f], K, d, j0, j1);
          wsum += j0 * st[36] + j1 * st[37];
          gsum += j0 * st[38] + j1 * st[39];
        }
        sWv[lb * VD + d] = wsum; sG[lb * VD + d] = gsum;
      } else {
        double h = 0, bb = 0;
        for (int f = 0; f < nfac; f++) { const double* st = F + f * MF; h += st[36] * st[36] + st[37] * st[37]; bb += st[36] * st[38] + st[37] * st[39]; }
        sS[2 * lb] = h > 1e-8 ? 1.0 / h : 0.0;       // eps of the reference's pseudo-inverse
        sS[2 * lb + 1] = bb;
      }
    }
    __syncthreads();
    // dense part: the Schur term of the eliminated depths (single owner per entry)
    for (int e = tid; e < VD * VD; e += nt) {
      const int d1 = e / VD, d2 = e - d1 * VD;
      double s = 0;
      for (int lb = 0; lb < nb; lb++) s += sWv[lb * VD + d1] * sWv[lb * VD + d2] * sS[2 * lb];
      As[e] -= s;
    }
    for (int d = tid; d < VD; d += nt) {
      double s = 0;
      for (int lb = 0; lb < nb; lb++) s += sG[lb * VD + d] - sWv[lb * VD + d] * sS[2 * lb + 1] * sS[2 * lb];
      gs[d] += s;
    }
    // sparse part: sum_f J_f^T J_f.  (local a, local b) with both columns shared by all factors (frame 0, extrinsics, td):
    // one owner sums over every factor; with a column of frame fj involved: one owner per (frame, a, b)
    for (int it = tid; it < NC * NC; it += nt) {
      const int a = it / NC, b2 = it - a * NC;
      if ((a >= 6 && a < 12) || (b2 >= 6 && b2 < 12)) continue;
      double s = 0;
      for (int lb = 0; lb < nb; lb++)
        for (int f = 0; f < sNf[lb]; f++) {
          double a0, a1, b0, b1;
          const double* st = sF + (lb * 16 + f) * MF;
          marg_lcol(st, a, a0, a1); marg_lcol(st, b2, b0, b1);
          s += a0 * b0 + a1 * b1;
        }
      As[marg_gcol(a, 0, K) * VD + marg_gcol(b2, 0, K)] += s;
    }
    for (int it = tid; it < (K - 1) * NC * NC; it += nt) {
      const int p = 1 + it / (NC * NC), ab = it - (p - 1) * NC * NC, a = ab / NC, b2 = ab - a * NC;
      if (!((a >= 6 && a < 12) || (b2 >= 6 && b2 < 12))) continue;
      double s = 0;
      for (int lb = 0; lb < nb; lb++) {
        const int f = sFof[lb * 16 + p];
        if (f < 0) continue;
        double a0, a1, b0, b1;
        const double* st = sF + (lb * 16 + f) * MF;
        marg_lcol(st, a, a0, a1); marg_lcol(st, b2, b0, b1);
        s += a0 * b0 + a1 * b1;
      }
      As[marg_gcol(a, p, K) * VD + marg_gcol(b2, p, K)] += s;
    }
  }
  __syncthreads();
  double* out = ma.part + (size_t)blockIdx.x * (VD * VD + VD);
  for (int e = tid; e < VD * VD + VD; e += nt) out[e] = As[e];
}

// Jacobi rotation that annihilates a_pq
__device__ __forceinline__ void jacobi_cs(double app, double aqq, double apq, double& c, double& s) {
  c = 1.0; s = 0.0;
  if (apq == 0.0) return;
  // with d = aqq - app, a2 = 2 apq, h = hypot(d, a2):  cos(2 theta) = |d| / h,  c = sqrt((1 + |d|/h) / 2),
  // s = sign(d) a2 / (2 h c) -- the smaller rotation.  Two reciprocal square roots, no division: this chain of dependent
  // FP64 operations is what every parallel-ordering step waits for.  c^2 + s^2 = 1 to rounding.
  const double d = aqq - app, a2 = 2.0 * apq;
  const double r = rsqrt(d * d + a2 * a2);         // 1 / h
  const double u = fma(0.5 * fabs(d), r, 0.5);     // (1 + |d| / h) / 2  in [1/2, 1]
  const double rc = rsqrt(u);
  c = u * rc;
  s = (d >= 0 ? 0.5 : -0.5) * a2 * r * rc;
}

// in-warp Jacobi eigen-decomposition of a symmetric me x me matrix (me even, <= 16) in shared memory, parallel (round-robin)
// ordering: me/2 disjoint rotations per step, no CTA barrier.  Am, Vm: [16*17].  Eigenvalues end up on the diagonal of Am.
__device__ void warp_jacobi16(double* Am, double* Vm, int me, int lane) {
  constexpr int LD = 17;
  const int half = me / 2;
  __shared__ int s_pq[16];
  __shared__ double s_cs[16];
  for (int sweep = 0; sweep < 30 && me >= 2; sweep++) {
    double off = 0, dg = 0;
    for (int e = lane; e < me * me; e += 32) {
      const int i = e / me, j = e - i * me;
      const double v = Am[i * LD + j] * Am[i * LD + j];
      if (i == j) dg += v; else off += v;
    }
    off = warp_sum(off); dg = warp_sum(dg);
    if (off <= 1e-30 * (dg + 1e-300)) break;
    for (int step = 0; step < me - 1; step++) {
      if (lane < half) {
        int p, q;
        if (lane == 0) { p = me - 1; q = step; }
        else { p = (step + lane) % (me - 1); q = (step - lane + (me - 1)) % (me - 1); }
        if (p > q) { int t2 = p; p = q; q = t2; }
        double c, s;
        jacobi_cs(Am[p * LD + p], Am[q * LD + q], Am[p * LD + q], c, s);
        s_pq[2 * lane] = p; s_pq[2 * lane + 1] = q; s_cs[2 * lane] = c; s_cs[2 * lane + 1] = s;
      }
      __syncwarp();
      for (int e = lane; e < half * half; e += 32) {
        const int I = e / half, Jp = e - I * half;
        const int p = s_pq[2 * I], q = s_pq[2 * I + 1], r = s_pq[2 * Jp], t2 = s_pq[2 * Jp + 1];
        const double ci = s_cs[2 * I], si = s_cs[2 * I + 1], cj = s_cs[2 * Jp], sj = s_cs[2 * Jp + 1];
        const double apr = Am[p * LD + r], apt = Am[p * LD + t2], aqr = Am[q * LD + r], aqt = Am[q * LD + t2];
        const double bpr = ci * apr - si * aqr, bpt = ci * apt - si * aqt;
        const double bqr = si * apr + ci * aqr, bqt = si * apt + ci * aqt;
        Am[p * LD + r] = cj * bpr - sj * bpt; Am[p * LD + t2] = sj * bpr + cj * bpt;
        Am[q * LD + r] = cj * bqr - sj * bqt; Am[q * LD + t2] = sj * bqr + cj * bqt;
      }
      for (int e = lane; e < half * me; e += 32) {
        const int pr = e / me, k = e - pr * me;
        const int p = s_pq[2 * pr], q = s_pq[2 * pr + 1];
        const double c = s_cs[2 * pr], s2 = s_cs[2 * pr + 1];
        const double vkp = Vm[k * LD + p], vkq = Vm[k * LD + q];
        Vm[k * LD + p] = c * vkp - s2 * vkq; Vm[k * LD + q] = s2 * vkp + c * vkq;
      }
      __syncwarp();
    }
  }
}

// ---- kernel 2: assembly, elimination of the dropped block, eigen-decomposition of the kept system (one CTA) -------------------
__global__ void __launch_bounds__(MARG_THREADS, 1) ba_marginalize_kernel(BaBatch bt, MargArgs ma) {
  extern __shared__ double sm[];
  const int tid = threadIdx.x, nt = blockDim.x, K = bt.K, M = ma.M, w = 0;
  const int VD = 6 * K + 6 + bt.est_td;     // visual layout dimension (frames, extrinsics, td when it is estimated)
  double* A = ma.A;
  double* bv = ma.b;
  unsigned long long* phase = reinterpret_cast<unsigned long long*>(ma.status + 4);   // BVIO_DEBUG: where the time goes
  if (tid == 0) phase[0] = global_ns();
  for (int i = tid; i < M * M; i += nt) A[i] = 0.0;
  for (int i = tid; i < M; i += nt) bv[i] = 0.0;
  __syncthreads();
  auto vmap = [&](int d) { return d < 6 * K ? 15 * (d / 6) + d % 6 : 15 * K + (d - 6 * K); };   // td: 15K + 6

  if (ma.flag == 0) {
    // ---- partial sums of the factor kernel, in a fixed order
    const size_t PS = (size_t)VD * VD + VD;
    for (int e = tid; e < VD * VD; e += nt) {
      double s = 0;
      for (int g = 0; g < ma.G; g++) s += ma.part[g * PS + e];
      const int d1 = e / VD, d2 = e - d1 * VD;
      A[(size_t)vmap(d1) * M + vmap(d2)] = s;
    }
    for (int d = tid; d < VD; d += nt) {
      double s = 0;
      for (int g = 0; g < ma.G; g++) s += ma.part[g * PS + VD * VD + d];
      bv[vmap(d)] = s;
    }
    __syncthreads();
    const double* imu = ma.part + ma.G * PS;         // 30x30 J^T J, then J^T r of the IMU factor 0 -> 1
    for (int e = tid; e < 930; e += nt) {
      if (e < 900) { const int a = e / 30, c = e - a * 30; A[(size_t)a * M + c] += imu[e]; }
      else bv[e - 900] += imu[e];
    }
    __syncthreads();
  }
  if (tid == 0) phase[1] = global_ns();
  // ---- prior (estimator.cpp:822-839 / 932-948): J^T J comes from ba_prepare_kernel (pr_H), J^T r from the residual here
  {
    const int np_ = bt.pr_n[w];
    if (np_ > 0) {
      double* sdx = sm;
      double* spr = sdx + bt.nmax;
      int* pmap = reinterpret_cast<int*>(spr + bt.nmax);   // prior column -> M layout (incl. extrinsics)
      const int nb = bt.pr_nb[w];
      for (int i = tid; i < np_; i += nt) pmap[i] = -1;
      __syncthreads();
      for (int bidx = tid; bidx < nb; bidx += nt) {
        int kind = bt.pr_kind[w * PRIOR_MAXB + bidx], fr = bt.pr_frame[w * PRIOR_MAXB + bidx], idx = bt.pr_idx[w * PRIOR_MAXB + bidx];
        int loc = kind == 1 ? 9 : (kind == 3 ? 1 : 6);
        int base = kind == 0 ? 15 * fr : (kind == 1 ? 15 * fr + 6 : (kind == 2 ? 15 * K : 15 * K + 6));
        for (int i = 0; i < loc; i++) pmap[idx + i] = base + i;
      }
      __syncthreads();
      prior_residual(bt, w, bt.pose0, bt.sb0, bt.ex, bt.td0, sdx, spr);
      const double* Jc = bt.pr_jac + (size_t)w * bt.nmax * bt.nmax;
      const double* Hp = bt.pr_H + (size_t)w * bt.nmax * bt.nmax;
      for (int e = tid; e < np_ * np_; e += nt) {
        int a = e / np_, c = e - a * np_;
        if (pmap[a] < 0 || pmap[c] < 0) continue;
        A[(size_t)pmap[a] * M + pmap[c]] += Hp[(size_t)a * bt.nmax + c];
      }
      // J^T r: one warp per column
      for (int a = tid >> 5; a < np_; a += nt >> 5) {
        if (pmap[a] < 0) continue;
        double s = 0;
        for (int k = tid & 31; k < np_; k += 32) s += Jc[(size_t)a * np_ + k] * spr[k];
        s = warp_sum(s);
        if ((tid & 31) == 0) bv[pmap[a]] += s;
      }
    }
    __syncthreads();
  }
  if (tid == 0) phase[2] = global_ns();
  // ---- eliminate the m dropped dimensions: Amm pseudo-inverse by Jacobi eigen-decomposition (warp 0)
  const int m = ma.m, n = ma.n, ne = ma.ne;
  double* Am = sm;                    // [16*17]
  double* Vm = Am + 272;              // [16*17]
  double* Ai = Vm + 272;              // [16*16] pseudo-inverse
  double* Tm = Ai + 256;              // [n*16]
  const int me = (m + 1) & ~1;
  for (int e = tid; e < 256; e += nt) {
    int i = e >> 4, j = e & 15;
    Am[i * 17 + j] = (i < m && j < m) ? 0.5 * (A[(size_t)ma.dropidx[i] * M + ma.dropidx[j]] + A[(size_t)ma.dropidx[j] * M + ma.dropidx[i]]) : 0.0;
    Vm[i * 17 + j] = (i == j) ? 1.0 : 0.0;
  }
  __syncthreads();
  if (tid < 32) warp_jacobi16(Am, Vm, me, tid);
  __syncthreads();
  if (tid == 0) phase[3] = global_ns();
  for (int e = tid; e < 256; e += nt) {
    int i = e >> 4, j = e & 15;
    double s = 0;
    if (i < m && j < m)
      for (int k = 0; k < m; k++) { double lam = Am[k * 17 + k]; if (lam > 1e-8) s += Vm[i * 17 + k] * Vm[j * 17 + k] / lam; }
    Ai[e] = s;
  }
  __syncthreads();
  for (int e = tid; e < n * 16; e += nt) {
    int i = e >> 4, j = e & 15;
    double s = 0;
    if (j < m) for (int k = 0; k < m; k++) s += A[(size_t)ma.keepidx[i] * M + ma.dropidx[k]] * Ai[k * 16 + j];
    Tm[e] = s;
  }
  __syncthreads();
  // ---- kept system Ar = Arr - Arm Amm^+ Amr (symmetrised), br; eigen-decomposition by parallel Jacobi
  const int ld = ne | 1;                      // odd row stride: column sweeps hit distinct banks
  double* Ar = Tm + (size_t)n * 16;           // [ne*ld]
  double* Vr = Ar + (size_t)ne * ld;          // [ne*ld]
  double* br = Vr + (size_t)ne * ld;          // [ne]
  double* cs = br + ne;                       // [ne] (c,s) per pair
  int* pp = reinterpret_cast<int*>(cs + ne);  // [ne] pairs
  double* red = reinterpret_cast<double*>(pp + ne + (ne & 1));   // [32]
  for (int e = tid; e < ne * ne; e += nt) {
    int i = e / ne, j = e - i * ne;
    double s = 0;
    if (i < n && j < n) {
      s = A[(size_t)ma.keepidx[i] * M + ma.keepidx[j]];
      for (int k = 0; k < m; k++) s -= Tm[i * 16 + k] * A[(size_t)ma.dropidx[k] * M + ma.keepidx[j]];
    }
    Ar[i * ld + j] = s;
    Vr[i * ne + j] = (i == j) ? 1.0 : 0.0;        // V is kept TRANSPOSED with row stride ne: Vr[c * ne + k] = V(k, c)
  }
  for (int i = tid; i < ne; i += nt) {
    double s = 0;
    if (i < n) {
      s = bv[ma.keepidx[i]];
      for (int k = 0; k < m; k++) s -= Tm[i * 16 + k] * bv[ma.dropidx[k]];
    }
    br[i] = s;
  }
  __syncthreads();
  for (int e = tid; e < ne * ne; e += nt) {       // symmetrise (two passes: read pairs, then write)
    int i = e / ne, j = e - i * ne;
    if (i > j) { double v = 0.5 * (Ar[i * ld + j] + Ar[j * ld + i]); Ar[i * ld + j] = v; Ar[j * ld + i] = v; }
  }
  __syncthreads();
  if (tid == 0) { phase[4] = global_ns(); phase[7] = (unsigned long long)clock64(); }
  if (ma.method == 1) {
    // ---- opt-in (BVIO_MARG_CHOLESKY=1; the default below is the reference's eigen-decomposition):
    //      (J, r) by diagonally pivoted Cholesky.  The reference factors A = V S V^T and keeps J = sqrt(S) V^T,
    //      r = sqrt(S)^-1 V^T b (:283-291); every use of the prior -- r0 + J dx in MarginalizationFactor::Evaluate,
    //      its cost, J^T J and J^T r in the next marginalization (:295-296) -- is invariant under J -> Q J,
    //      r -> Q r with Q orthogonal, so any factor with J^T J = A, J^T r = b is the same prior.  Pivoting on the
    //      largest remaining diagonal is rank revealing: it stops when what is left is below the reference's
    //      eps = 1e-8 (marginalization_factor.h:70), the counterpart of its eigenvalue threshold.  Difference to the
    //      eigen route: when A is rank deficient (gauge directions) that one also projects b onto range(A); here the
    //      rounding-level component of b along the null directions (~1e-4 |b|, no curvature behind it) stays in J^T r.
    int* perm = pp;                            // [ne]
    double* col = cs;                          // [ne] current column of L (pivot order)
    __shared__ int s_piv;
    __shared__ double s_d;
    for (int e = tid; e < n * n; e += nt) ma.out_jac[e] = 0.0;
    for (int i = tid; i < n; i += nt) { perm[i] = i; ma.out_res[i] = 0.0; }
    __syncthreads();
    int rank = 0;
    double minpiv = 1e300, maxpiv = 0;
    for (int k = 0; k < n; k++) {
      if (tid < 32) {
        double best = -1.0;
        int bi = n;
        for (int i = k + tid; i < n; i += 32) {
          const double v = Ar[perm[i] * ld + perm[i]];
          if (v > best) { best = v; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const double ov = __shfl_xor_sync(0xffffffffu, best, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (tid == 0) {
          s_piv = (best > 1e-8) ? bi : -1;
          if (best > 1e-8) { const int t2 = perm[k]; perm[k] = perm[bi]; perm[bi] = t2; s_d = sqrt(best); }
        }
      }
      __syncthreads();
      if (s_piv < 0) break;
      rank = k + 1;
      const int pk = perm[k];
      const double d = s_d, rk = br[pk] / d;
      minpiv = fmin(minpiv, d * d); maxpiv = fmax(maxpiv, d * d);
      for (int i = k + tid; i < n; i += nt) {
        const double lik = (i == k) ? d : Ar[perm[i] * ld + pk] / d;
        col[i] = lik;
        ma.out_jac[(size_t)perm[i] * n + k] = lik;         // J(k, perm[i]) = L(i, k), column-major
      }
      if (tid == 0) ma.out_res[k] = rk;
      __syncthreads();
      const int rem = n - k - 1;
      for (int e = tid; e < rem * rem; e += nt) {
        const int i = k + 1 + e / rem, j = k + 1 + e % rem;
        Ar[perm[i] * ld + perm[j]] -= col[i] * col[j];
      }
      for (int i = k + 1 + tid; i < n; i += nt) br[perm[i]] -= col[i] * rk;
      __syncthreads();
    }
    if (tid == 0) { ma.status[0] = rank; ma.status[1] = (int)(100 * log10(minpiv)); ma.status[2] = (int)(100 * log10(maxpiv + 1e-300)); phase[5] = phase[6] = global_ns(); }
    return;
  }
  // Parallel (round-robin) two-sided Jacobi.  Work is assigned once: thread t < half (half+1)/2 owns the pair-of-pairs
  // (I >= J) -- its 2x2 block of A and, by symmetry, the mirrored block -- and up to three (pair, row) items of V; per
  // step: `half` threads pick the rotations, one barrier, everybody applies A <- J^T A J and V <- V J in place
  // (single owner per element), one barrier.  Only the lower triangle of A is kept up to date (element (i, j) lives at
  // row max(i,j), column min(i,j)): half the shared-memory stores of a full symmetric update.
  int sweeps = 0;
  const int half = ne / 2, nblk = half * (half + 1) / 2;
  int2* pq = reinterpret_cast<int2*>(pp);       // [half] (p, q)
  double2* rot = reinterpret_cast<double2*>(cs);   // [half] (c, s)
  int bI = -1, bJ = -1;
  const bool blk_more = nblk > nt;              // more pair-of-pairs than threads (n > ~88): the rest in a loop
  if (tid < nblk) {
    bI = (int)((sqrtf(8.0f * (float)tid + 1.0f) - 1.0f) * 0.5f);
    while (bI * (bI + 1) / 2 > tid) bI--;
    while ((bI + 1) * (bI + 2) / 2 <= tid) bI++;
    bJ = tid - bI * (bI + 1) / 2;
  }
  // V <- V J: an item is (pair, two consecutive rows); with V stored transposed the two columns of a pair are two
  // contiguous runs, read and written as double2 (conflict-free, half the shared-memory instructions)
  constexpr int VI = 2;
  int vpr[VI], vk[VI];                          // this thread's (pair, row-pair) items of V
#pragma unroll
  for (int u = 0; u < VI; u++) {
    const int e = tid + u * nt;
    vpr[u] = e < half * half ? e / half : -1;
    vk[u] = e < half * half ? 2 * (e - (e / half) * half) : 0;
  }
  const bool v_more = half * half > VI * nt;
  for (; sweeps < 40 && ne >= 2; sweeps++) {
    double off = 0, dg = 0;
    for (int e = tid; e < ne * ne; e += nt) {
      int i = e / ne, j = e - i * ne;
      if (j > i) continue;
      double v = Ar[i * ld + j] * Ar[i * ld + j];
      if (i == j) dg += v; else off += 2.0 * v;
    }
    off = block_sum(off, red);
    dg = block_sum(dg, red);
    __syncthreads();
    if (off <= 1e-30 * (dg + 1e-300)) break;   // off-diagonal Frobenius norm below 1e-15 ||A||: converged in double
    for (int step = 0; step < ne - 1; step++) {
      // round-robin pairing: ne-1 is fixed, the others rotate
      if (tid < half) {
        int p, q;
        if (tid == 0) { p = ne - 1; q = step; }
        else {
          p = step + tid; if (p >= ne - 1) p -= ne - 1;
          q = step - tid; if (q < 0) q += ne - 1;
        }
        if (p > q) { int t2 = p; p = q; q = t2; }
        double c, s;
        jacobi_cs(Ar[p * ld + p], Ar[q * ld + q], Ar[q * ld + p], c, s);
        pq[tid] = make_int2(p, q); rot[tid] = make_double2(c, s);
      }
      __syncthreads();
      for (int e = tid, I = bI, Jb = bJ; e < nblk; e += nt) {
        if (e != tid) {                              // only when there are more pair-of-pairs than threads (n > ~88)
          I = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
          while (I * (I + 1) / 2 > e) I--;
          while ((I + 1) * (I + 2) / 2 <= e) I++;
          Jb = e - I * (I + 1) / 2;
        }
        const int2 a = pq[I], b2 = pq[Jb];
        const double2 ri = rot[I], rj = rot[Jb];
        const int p = a.x, q = a.y, r = b2.x, t2 = b2.y;
        // lower-triangle addresses; on the diagonal block (I == J: r = p, t2 = q) both (p,q) and (q,p) name one element
        const int ipr = p >= r ? p * ld + r : r * ld + p, ipt = p >= t2 ? p * ld + t2 : t2 * ld + p;
        const int iqr = q >= r ? q * ld + r : r * ld + q, iqt = q >= t2 ? q * ld + t2 : t2 * ld + q;
        const double apr = Ar[ipr], apt = Ar[ipt], aqr = Ar[iqr], aqt = Ar[iqt];
        // rows: [p'; q'] = [c -s; s c] [p; q]
        const double bpr = ri.x * apr - ri.y * aqr, bpt = ri.x * apt - ri.y * aqt;
        const double bqr = ri.y * apr + ri.x * aqr, bqt = ri.y * apt + ri.x * aqt;
        // columns: [r' t'] = [r t] [c s; -s c]
        const double npr = rj.x * bpr - rj.y * bpt, npt = rj.y * bpr + rj.x * bpt, nqr = rj.x * bqr - rj.y * bqt, nqt = rj.y * bqr + rj.x * bqt;
        Ar[ipr] = npr; Ar[iqt] = nqt; Ar[iqr] = nqr;
        if (I != Jb) Ar[ipt] = npt;                  // (diagonal block: ipt == iqr, the same element, npt == nqr up to rounding)
        if (!blk_more) break;
      }
#pragma unroll
      for (int u = 0; u < VI; u++) {
        if (vpr[u] < 0) continue;
        const int2 a = pq[vpr[u]];
        const double2 r = rot[vpr[u]];
        double2* cp = reinterpret_cast<double2*>(Vr + a.x * ne + vk[u]);
        double2* cq = reinterpret_cast<double2*>(Vr + a.y * ne + vk[u]);
        const double2 vp = *cp, vq = *cq;
        *cp = make_double2(r.x * vp.x - r.y * vq.x, r.x * vp.y - r.y * vq.y);
        *cq = make_double2(r.y * vp.x + r.x * vq.x, r.y * vp.y + r.x * vq.y);
      }
      if (v_more)
        for (int e = tid + VI * nt; e < half * half; e += nt) {
          const int pr = e / half, k = 2 * (e - pr * half);
          const int2 a = pq[pr];
          const double2 r = rot[pr];
          double2* cp = reinterpret_cast<double2*>(Vr + a.x * ne + k);
          double2* cq = reinterpret_cast<double2*>(Vr + a.y * ne + k);
          const double2 vp = *cp, vq = *cq;
          *cp = make_double2(r.x * vp.x - r.y * vq.x, r.x * vp.y - r.y * vq.y);
          *cq = make_double2(r.y * vp.x + r.x * vq.x, r.y * vp.y + r.x * vq.y);
        }
      __syncthreads();
    }
  }
  if (tid == 0) { phase[5] = global_ns(); phase[7] = (unsigned long long)clock64() - phase[7]; }
  // ---- linearized_jacobians = sqrt(S) V^T, linearized_residuals = sqrt(S^-1) V^T b  (:283-291)
  for (int e = tid; e < n * n; e += nt) {
    int i = e / n, k = e - i * n;       // column-major J(k, i) at [i*n + k]
    double lam = Ar[k * ld + k];
    ma.out_jac[e] = (lam > 1e-8 ? sqrt(lam) : 0.0) * Vr[k * ne + i];
  }
  for (int k = tid; k < n; k += nt) {
    double lam = Ar[k * ld + k], s = 0;
    for (int i = 0; i < n; i++) s += Vr[k * ne + i] * br[i];
    ma.out_res[k] = (lam > 1e-8 ? sqrt(1.0 / lam) : 0.0) * s;
  }
  if (tid == 0) { ma.status[0] = sweeps; phase[6] = global_ns(); }
}

cudaError_t ba_configure_marginalize(void) {
  cudaError_t e = cudaFuncSetAttribute(ba_marginalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(ba_marg_factors_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  return e;
}
size_t ba_marginalize_smem_bytes(int K, int nmax, int n) {
  int ne = (n + 1) & ~1;
  size_t c = 2 * (size_t)nmax + nmax / 2 + 2;                                             // prior phase
  size_t d = 800 + (size_t)n * 16 + 2 * (size_t)ne * (ne + 1) + 2 * ne + ne / 2 + 2 + 32 + 8;   // elimination + Jacobi
  return (c > d ? c : d) * sizeof(double);
}
static size_t marg_factors_smem_bytes(int K) {
  const int VD = 6 * K + 7;
  size_t d = (size_t)(K + 1) * FR + MARG_NB * 16 * MF + 2 * MARG_NB * VD + 2 * MARG_NB + (size_t)VD * VD + VD;
  size_t imu = 930 + 32;
  return (d > imu ? d : imu) * sizeof(double) + sizeof(int) * (2 * MARG_NB * 16 + MARG_NB + 4);
}
size_t ba_marginalize_part_doubles(int K, int G) {
  const int VD = 6 * K + 7;
  return (size_t)G * ((size_t)VD * VD + VD) + 930;
}
int ba_marginalize_groups(int n0) { int g = (n0 + MARG_NB - 1) / MARG_NB; return g < 1 ? 1 : (g > 24 ? 24 : g); }

int ba_launch_marginalize(const BaBatch& bt, int flag, int m, int n, int n0, const int* dropidx, const int* keepidx, double* A,
                          double* b, double* part, double* out_jac, double* out_res, int* status, int method, cudaStream_t st) {
  MargArgs ma;
  ma.method = method;
  ma.flag = flag; ma.M = 15 * bt.K + 7; ma.m = m; ma.n = n; ma.ne = (n + 1) & ~1;
  ma.n0 = n0; ma.G = ba_marginalize_groups(n0); ma.part = part;
  ma.dropidx = dropidx; ma.keepidx = keepidx; ma.A = A; ma.b = b; ma.out_jac = out_jac; ma.out_res = out_res; ma.status = status;
  int launches = 1;
  if (flag == 0) {
    ba_marg_factors_kernel<<<ma.G + 1, MARG_FT, marg_factors_smem_bytes(bt.K), st>>>(bt, ma);
    launches++;
  }
  ba_marginalize_kernel<<<1, MARG_THREADS, ba_marginalize_smem_bytes(bt.K, bt.nmax, n), st>>>(bt, ma);
  return launches;
}

}  // namespace bvio
