// sel_kernels.cu -- anticipated feature selection on sm_100a (FP64).  Replaces the numerical part of
// FeatureSelector::select (vins_estimator/src/feature_selector.cpp:139-170):
//
//   sel_build   calcInfoFromFeatures + findNNDepth + inFOV (feature_selector.cpp:239-376, 437-459) and
//               PinholeCamera::spaceToPlane (camera_model/src/camera_models/PinholeCamera.cc:520-542):
//               one warp per candidate, lane = horizon frame; writes the packed T x T position block
//   sel_omega   calcInfoFromRobotMotion + createLinearImuMatrices + addOmegaPrior
//               (feature_selector.cpp:463-609), adds the tracked features' blocks (:620-623) and
//               Schur-complements the non-position variables once (see sel.h)
//   sel_round   one greedy round of selectInformativeFeatures (feature_selector.cpp:633-682): one warp per
//               remaining candidate computes logdet(R + p C_ell) by an in-warp Cholesky (the value the
//               reference gets from Utility::logdet, utility.h:143-167); the last CTA picks the arg-max
//               and applies Omega_S += p Delta (:674)
//   sel_apply   multi-GPU: pick among the gathered per-rank winner records, identical on every rank
#include "sel.h"
#include <float.h>
#include <mutex>

namespace bvio {

__device__ __forceinline__ int tri(int i, int j) { return i * (i + 1) / 2 + j; }   // i >= j
__device__ __forceinline__ d3 normalized3(d3 a) {
  double n = sqrt(dot3(a, a));
  return {a.x / n, a.y / n, a.z / n};
}
__device__ __forceinline__ void tri_decode(int e, int& i, int& j) {
  i = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
  while (i * (i + 1) / 2 > e) i--;
  while ((i + 1) * (i + 2) / 2 <= e) i++;
  j = e - i * (i + 1) / 2;
}

// =============================================================================================
__global__ void sel_reset_kernel(SelProb sp) {
  int i = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
  for (int k = i; k < sp.N; k += stride) { sp.taken[k] = 0; sp.valid[k] = 0; }
  for (int k = i; k < sp.kappa; k += stride) { sp.out_idx[k] = -1; sp.out_val[k] = 0.0; }
  if (i == 0) {
    SelCtrl c;
    c.scored = 0; c.min_margin = INFINITY; c.logdet_oo = 0; c.final_logdet = 0;
    c.n_selected = 0; c.round = 0; c.n_valid = 0; c.pad = 0; c.ticket = 0; c.pad2 = 0; c.peer_timeout = 0; c.pad3 = 0;
    c.t_score = 0; c.t_barrier = 0; c.t_exchange = 0;
    *sp.ctrl = c;
  }
}

// =============================================================================================
// build: one warp per feature (local candidates, then the tracked features)
// =============================================================================================
__global__ void __launch_bounds__(32 * SEL_WARPS) sel_build_kernel(SelProb sp) {
  __shared__ double sC[SEL_WARPS][BVIO_HMAX * 9];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nloc = sp.b1 - sp.b0, H = sp.H;
  const int f = blockIdx.x * SEL_WARPS + warp;
  if (f >= nloc + sp.U) return;
  const bool is_used = f >= nloc;
  const int idx = is_used ? f - nloc : sp.b0 + f;
  const double2 xy = is_used ? sp.used_xy[idx] : sp.cand_xy[idx];
  // findNNDepth (feature_selector.cpp:437-459): exact 1-NN, first-found minimum
  double bd = DBL_MAX;
  int bi = 0x7fffffff;
  for (int i = lane; i < sp.C; i += 32) {
    double dx = sp.cloud_xy[i].x - xy.x, dy = sp.cloud_xy[i].y - xy.y;
    double d = dx * dx + dy * dy;
    if (d < bd) { bd = d; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double od = __shfl_xor_sync(0xffffffffu, bd, o);
    int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
  }
  if (bi == 0x7fffffff) bi = 0;
  const double depth = sp.C > 0 ? sp.cloud_depth[bi] : 1.0;

  const q4 qic{sp.q_ic[0], sp.q_ic[1], sp.q_ic[2], sp.q_ic[3]};
  const d3 tic{sp.t_ic[0], sp.t_ic[1], sp.t_ic[2]};
  const d3 P1{sp.k1_pos[0], sp.k1_pos[1], sp.k1_pos[2]};
  const q4 Q1{sp.k1_quat[0], sp.k1_quat[1], sp.k1_quat[2], sp.k1_quat[3]};
  const d3 t_wc1 = P1 + qrot(Q1, tic);
  const q4 q_wc1 = qmul(Q1, qic);
  d3 feat = normalized3(d3{xy.x, xy.y, 1.0});
  feat = depth * feat;
  const d3 pell = t_wc1 + qrot(q_wc1, feat);

  const int h = lane + 1;
  bool vis = false;
  double Cm[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (h <= H) {
    d3 u;
    q4 qwc;
    if (h == 1) {
      u = normalized3(feat); qwc = q_wc1; vis = true;
    } else {
      const q4 Qh{sp.hquat[4 * h], sp.hquat[4 * h + 1], sp.hquat[4 * h + 2], sp.hquat[4 * h + 3]};
      const d3 t_wch = d3{sp.hpos[3 * h], sp.hpos[3 * h + 1], sp.hpos[3 * h + 2]} + qrot(Qh, tic);
      qwc = qmul(Qh, qic);
      u = normalized3(qrot(qinv(qwc), pell - t_wch));
      // PinholeCamera::spaceToPlane with radtan distortion, then inFOV (round to int pixel)
      const bvio_camera& c = sp.cam;
      double mx = u.x / u.z, my = u.y / u.z;
      double mx2 = mx * mx, my2 = my * my, mxy = mx * my, rho2 = mx2 + my2;
      double rad = c.k1 * rho2 + c.k2 * rho2 * rho2;
      double ddx = mx * rad + 2.0 * c.p1 * mxy + c.p2 * (rho2 + 2.0 * mx2);
      double ddy = my * rad + 2.0 * c.p2 * mxy + c.p1 * (rho2 + 2.0 * my2);
      double pu = c.fx * (mx + ddx) + c.cx, pv = c.fy * (my + ddy) + c.cy;
      double ru = round(pu), rv = round(pv);
      vis = (0.0 <= ru && ru < (double)c.width) && (0.0 <= rv && rv < (double)c.height);
    }
    if (vis) {
      // B_h = [u]x * ((q_WC_h * q_IC)^-1).R  -- q_IC applied twice, as the reference does (:304, :321)
      double Rm[9], Bm[9];
      qmat(qinv(qmul(qwc, qic)), Rm);
      const double sk[9] = {0, -u.z, u.y, u.z, 0, -u.x, -u.y, u.x, 0};
      mm3(sk, Rm, Bm);
      mtm3(Bm, Bm, Cm);
    }
  }
  const unsigned vm = __ballot_sync(0xffffffffu, vis && h >= 2);
  const bool ok = (1 + __popc(vm)) > 1;
  if (lane == 0) {
    if (is_used) sp.valid_u[idx] = ok; else { sp.valid[idx] = ok; if (ok) atomicAdd(&sp.ctrl->n_valid, 1); }
    sp.depth[is_used ? sp.N + idx : idx] = depth;
  }
  if (!ok) return;
  double* myC = sC[warp];
  if (h <= H) {
#pragma unroll
    for (int k = 0; k < 9; k++) myC[(h - 1) * 9 + k] = Cm[k];
  }
  __syncwarp();
  // EtE in the reference's summation order: h = 2..H, then the k+1 block (:307-326)
  double E[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int hh = 2; hh <= H; hh++)
#pragma unroll
    for (int k = 0; k < 9; k++) E[k] += myC[(hh - 1) * 9 + k];
#pragma unroll
  for (int k = 0; k < 9; k++) E[k] += myC[k];
  // Eigen fixed-size 3x3 inverse: cofactors / determinant
  double W[9];
  {
    double c00 = E[4] * E[8] - E[5] * E[7], c01 = E[2] * E[7] - E[1] * E[8], c02 = E[1] * E[5] - E[2] * E[4];
    double c10 = E[5] * E[6] - E[3] * E[8], c11 = E[0] * E[8] - E[2] * E[6], c12 = E[2] * E[3] - E[0] * E[5];
    double c20 = E[3] * E[7] - E[4] * E[6], c21 = E[1] * E[6] - E[0] * E[7], c22 = E[0] * E[4] - E[1] * E[3];
    double det = E[0] * c00 + E[1] * c10 + E[2] * c20, id = 1.0 / det;
    W[0] = id * c00; W[1] = id * c01; W[2] = id * c02; W[3] = id * c10; W[4] = id * c11; W[5] = id * c12;
    W[6] = id * c20; W[7] = id * c21; W[8] = id * c22;
  }
  double* out = (is_used ? sp.Cu : sp.Cc) + (size_t)idx * sp.TT;
  const int npairs = H * (H + 1) / 2;
  for (int pr = lane; pr < npairs; pr += 32) {
    int i, j;
    tri_decode(pr, i, j);   // 0-based frames i >= j  (reference's i,j = 1..H)
    double CW[9], Dij[9];
    mm3(myC + i * 9, W, CW);
    mmt3(CW, myC + j * 9, Dij);
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
      for (int b = 0; b < 3; b++) {
        if (i == j && b > a) continue;
        double lo = (i == j) ? myC[i * 9 + a * 3 + b] - Dij[a * 3 + b] : -Dij[a * 3 + b];
        out[tri(3 * i + a, 3 * j + b)] = lo;
      }
  }
}

// =============================================================================================
// omega: one CTA
// =============================================================================================
__global__ void __launch_bounds__(256) sel_omega_kernel(SelProb sp) {
  extern __shared__ double sm[];
  const int H = sp.H, D = sp.D, T = sp.T, Do = sp.Do, tid = threadIdx.x, nt = blockDim.x;
  double* M = sm;
  int* oi = reinterpret_cast<int*>(M + (size_t)D * D);
  int* pi = oi + Do;
  // ---- createLinearImuMatrices per consecutive pair (feature_selector.cpp:531-598)
  if (tid < H) {
    const int h = tid + 1;
    const q4 Qi{sp.hquat[4 * (h - 1)], sp.hquat[4 * (h - 1) + 1], sp.hquat[4 * (h - 1) + 2], sp.hquat[4 * (h - 1) + 3]};
    const q4 Qj{sp.hquat[4 * h], sp.hquat[4 * h + 1], sp.hquat[4 * h + 2], sp.hquat[4 * h + 3]};
    const int nr = sp.nr_imu;
    double Nij[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, Mij[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    double c11 = 0, c12 = 0;
    for (int i = 0; i < nr; ++i) {
      q4 q = qslerp(Qi, i / (double)nr, Qj);
      double jkh = nr - i - 0.5, Rm[9];
      qmat(q, Rm);
#pragma unroll
      for (int k = 0; k < 9; k++) { Nij[k] += jkh * Rm[k]; Mij[k] += Rm[k]; }
      c11 += jkh * jkh;
      c12 += jkh;
    }
    const double dI = sp.delta_imu, d2 = dI * dI, d3v = d2 * dI, d4 = d3v * dI;
    const double a = 1.0 * nr * c11 * d4 * sp.acc_var, b = 1.0 * c12 * d3v * sp.acc_var;
    const double c = 1.0 * nr * d2 * sp.acc_var, e = 1.0 * nr * sp.acc_bias_var;
    // covImu = [[aI bI 0],[bI cI 0],[0 0 eI]]: inverse in closed form
    const double det = a * c - b * b;
    double* Wm = sp.pair + (size_t)(h - 1) * 324;
    double* Am = Wm + 81;
    for (int k = 0; k < 162; k++) Wm[k] = 0.0;
    for (int k = 0; k < 3; k++) {
      Wm[k * 9 + k] = c / det; Wm[k * 9 + 3 + k] = -b / det; Wm[(3 + k) * 9 + k] = -b / det;
      Wm[(3 + k) * 9 + 3 + k] = a / det; Wm[(6 + k) * 9 + 6 + k] = 1.0 / e;
    }
    for (int k = 0; k < 9; k++) Am[k * 9 + k] = -1.0;
    for (int k = 0; k < 3; k++) Am[k * 9 + 3 + k] = -1.0 * nr * dI;
    for (int r = 0; r < 3; r++)
      for (int cc = 0; cc < 3; cc++) { Am[r * 9 + 6 + cc] = d2 * Nij[r * 3 + cc]; Am[(3 + r) * 9 + 6 + cc] = dI * Mij[r * 3 + cc]; }
  }
  __syncthreads();
  for (int e = tid; e < H * 81; e += nt) {
    int hh = e / 81, rc = e - hh * 81, r = rc / 9, c = rc - r * 9;
    const double* Wm = sp.pair + (size_t)hh * 324;
    const double* Am = Wm + 81;
    double s = 0;
    for (int k = 0; k < 9; k++) s += Am[k * 9 + r] * Wm[k * 9 + c];
    sp.pair[(size_t)hh * 324 + 162 + rc] = s;   // A^T W
  }
  __syncthreads();
  for (int e = tid; e < H * 81; e += nt) {
    int hh = e / 81, rc = e - hh * 81, r = rc / 9, c = rc - r * 9;
    const double* Am = sp.pair + (size_t)hh * 324 + 81;
    const double* AtW = sp.pair + (size_t)hh * 324 + 162;
    double s = 0;
    for (int k = 0; k < 9; k++) s += AtW[r * 9 + k] * Am[k * 9 + c];
    sp.pair[(size_t)hh * 324 + 243 + rc] = s;   // A^T W A
  }
  __syncthreads();
  // ---- block-tridiagonal Omega_kkH (feature_selector.cpp:507-523) + I9 prior (:605-608)
  for (int e = tid; e < D * D; e += nt) {
    int i = e / D, j = e - i * D, bi = i / 9, bj = j / 9, a = i - bi * 9, b = j - bj * 9;
    double v = 0;
    if (bi == bj) {
      if (bi >= 1) v += sp.pair[(size_t)(bi - 1) * 324 + a * 9 + b];
      if (bi < H) v += sp.pair[(size_t)bi * 324 + 243 + a * 9 + b];
      if (bi == 0) v += sp.has_prior ? sp.omega_prior[a * 9 + b] : (a == b ? 1.0 : 0.0);
    } else if (bj == bi + 1) {
      v = sp.pair[(size_t)(bj - 1) * 324 + 162 + a * 9 + b];
    } else if (bi == bj + 1) {
      v = sp.pair[(size_t)(bi - 1) * 324 + 162 + b * 9 + a];
    }
    M[e] = v;
    sp.omega[e] = v;
  }
  if (tid == 0) {
    int no = 0, npos = 0;
    for (int i = 0; i < D; i++) {
      int blk = i / 9, a = i - blk * 9;
      if (blk >= 1 && a < 3) pi[npos++] = i; else oi[no++] = i;
    }
  }
  __syncthreads();
  // ---- Omega += Delta_used (feature_selector.cpp:620-623)
  for (int e = tid; e < T * T; e += nt) {
    int t1 = e / T, t2 = e - t1 * T;
    int lo = t1 >= t2 ? tri(t1, t2) : tri(t2, t1);
    double v = M[pi[t1] * D + pi[t2]];
    for (int u = 0; u < sp.U; u++)
      if (sp.valid_u[u]) v += sp.Cu[(size_t)u * sp.TT + lo];
    M[pi[t1] * D + pi[t2]] = v;
  }
  __syncthreads();
  // ---- Cholesky of M_oo (in place, through the index map)
  for (int k = 0; k < Do; k++) {
    const int ok = oi[k];
    const double d = sqrt(M[ok * D + ok]);
    __syncthreads();
    if (tid == 0) M[ok * D + ok] = d;
    for (int i = k + 1 + tid; i < Do; i += nt) M[oi[i] * D + ok] /= d;
    __syncthreads();
    const int m = Do - 1 - k;
    for (int e = tid; e < m * m; e += nt) {
      int ii = e / m, jj = e - ii * m;
      if (jj > ii) continue;
      int ri = oi[k + 1 + ii], rj = oi[k + 1 + jj];
      M[ri * D + rj] -= M[ri * D + ok] * M[rj * D + ok];
    }
    __syncthreads();
  }
  // ---- Y = L^-1 M_op, stored over M_op
  if (tid < T) {
    const int pc = pi[tid];
    for (int i = 0; i < Do; i++) {
      const int ri = oi[i];
      double s = M[ri * D + pc];
      for (int k = 0; k < i; k++) s -= M[ri * D + oi[k]] * M[oi[k] * D + pc];
      M[ri * D + pc] = s / M[ri * D + ri];
    }
  }
  __syncthreads();
  // ---- R = S0 = M_pp - Y^T Y ; logdet(M_oo)
  for (int e = tid; e < T * T; e += nt) {
    int t1 = e / T, t2 = e - t1 * T;
    if (t2 > t1) continue;
    double s = M[pi[t1] * D + pi[t2]];
    for (int i = 0; i < Do; i++) s -= M[oi[i] * D + pi[t1]] * M[oi[i] * D + pi[t2]];
    sp.R[tri(t1, t2)] = s;
  }
  if (tid == 0) {
    double ld = 0;
    for (int i = 0; i < Do; i++) ld += log(M[oi[i] * D + oi[i]]);
    sp.ctrl->logdet_oo = 2.0 * ld;
  }
}

// =============================================================================================
// in-warp Cholesky of a packed-lower T x T matrix: returns sum(log diag(L)) or NaN.
// Register-resident: lane i holds row i (and i+32 when T > 32); column j is broadcast with warp
// shuffles, so the whole factorization runs without shared-memory round trips.  T is a compile-time
// constant so that every register index is static.
// =============================================================================================
// T > 32 (HORIZON 11..16; the reference compiles HORIZON = 13, T = 39): lane i keeps row i of the leading 32 x 32 block
// as above, and lane k additionally keeps column k of the E = T - 32 trailing rows.  While column j of L11 is being
// finalised the trailing rows ride along (L21 = A21 L11^-T: E extra broadcasts per column), the E x E Schur complement
// A22 - L21 L21^T is E(E+1)/2 warp reductions, and its small Cholesky runs replicated in every lane.  ~50 live doubles
// per lane instead of the 2 x T of the row-per-lane scheme (which spilled).
template <int T, bool ADD>
__device__ __forceinline__ double warp_chol_logdet_wide(const double* A, int lane, const double* C, double p) {
  constexpr int E = T - 32;
  double a[32], b[E];
  {
    const double* row = A + tri(lane, 0);
    const double* crow = ADD ? C + tri(lane, 0) : nullptr;
#pragma unroll
    for (int k = 0; k < 32; k++) a[k] = (k <= lane) ? (ADD ? fma(p, crow[k], row[k]) : row[k]) : 0.0;
#pragma unroll
    for (int r = 0; r < E; r++) b[r] = ADD ? fma(p, C[tri(32 + r, lane)], A[tri(32 + r, lane)]) : A[tri(32 + r, lane)];
  }
  double mypiv = 1.0;
  bool bad = false;
#pragma unroll
  for (int j = 0; j < 32; j++) {
    const double pj = __shfl_sync(0xffffffffu, a[j], j);
    bad |= !(pj > 0.0);
    if (lane == j) mypiv = pj;
    const double inv = rsqrt(pj);
    a[j] *= inv;                                   // lane i: L11[i][j]
#pragma unroll
    for (int k = j + 1; k < 32; k++) {
      const double lkj = __shfl_sync(0xffffffffu, a[j], k);
      a[k] = fma(-a[j], lkj, a[k]);
    }
#pragma unroll
    for (int r = 0; r < E; r++) {
      const double l21 = __shfl_sync(0xffffffffu, b[r] * inv, j);      // L21[r][j], final
      if (lane == j) b[r] = l21;
      else if (lane > j) b[r] = fma(-l21, a[j], b[r]);                   // A21[r][lane] -= L21[r][j] L11[lane][j]
    }
  }
  // Schur complement of the trailing block, replicated in every lane, then its Cholesky
  double s22[E][E];
#pragma unroll
  for (int r = 0; r < E; r++)
#pragma unroll
    for (int c = 0; c <= r; c++) {
      const double a22 = ADD ? fma(p, C[tri(32 + r, 32 + c)], A[tri(32 + r, 32 + c)]) : A[tri(32 + r, 32 + c)];
      s22[r][c] = a22 - warp_sum(b[r] * b[c]);
    }
  double l = warp_sum(log(mypiv));
#pragma unroll
  for (int j = 0; j < E; j++) {
    const double pj = s22[j][j];
    bad |= !(pj > 0.0);
    l += log(pj);
    const double inv = rsqrt(pj);
#pragma unroll
    for (int r = j + 1; r < E; r++) s22[r][j] *= inv;
#pragma unroll
    for (int r = j + 1; r < E; r++)
#pragma unroll
      for (int c = j + 1; c <= r; c++) s22[r][c] = fma(-s22[r][j], s22[c][j], s22[r][c]);
  }
  return bad ? NAN : 0.5 * l;
}

template <int T, bool ADD = false>
__device__ __forceinline__ double warp_chol_logdet(const double* A, int lane, const double* C = nullptr, double p = 0.0) {
  if constexpr (T > 32) return warp_chol_logdet_wide<T, ADD>(A, lane, C, p);
  constexpr int R = (T + 31) / 32;
  double a[R][T];
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int i = lane + 32 * r;
    const double* row = A + tri(i < T ? i : 0, 0);
    const double* crow = ADD ? C + tri(i < T ? i : 0, 0) : nullptr;
#pragma unroll
    for (int k = 0; k < T; k++) a[r][k] = (i < T && k <= i) ? (ADD ? fma(p, crow[k], row[k]) : row[k]) : 0.0;
  }
  double mypiv[R];
#pragma unroll
  for (int r = 0; r < R; r++) mypiv[r] = 1.0;
  bool bad = false;
#pragma unroll
  for (int j = 0; j < T; j++) {
    const double pj = __shfl_sync(0xffffffffu, a[j / 32][j], j % 32);
    bad |= !(pj > 0.0);
    if (lane == (j % 32)) mypiv[j / 32] = pj;
    const double inv = rsqrt(pj);
#pragma unroll
    for (int r = 0; r < R; r++) a[r][j] *= inv;
#pragma unroll
    for (int k = j + 1; k < T; k++) {
      const double lkj = __shfl_sync(0xffffffffu, a[k / 32][j], k % 32);
#pragma unroll
      for (int r = 0; r < R; r++) a[r][k] = fma(-a[r][j], lkj, a[r][k]);
    }
  }
  double l = 0;
#pragma unroll
  for (int r = 0; r < R; r++)
    if (lane + 32 * r < T) l += log(mypiv[r]);
  l = warp_sum(l);
  return bad ? NAN : 0.5 * l;
}

__device__ __forceinline__ void merge_best(double& best, double& second, int& idx, double ob, double os, int oidx) {
  // (best, idx) ordered by value.  An exact tie is the reference's upper-bound collision: bit-identical candidates
  // share one slot of the `UBs[ub] = feature_id` map, only the later-iterated (larger) id is evaluated in that round
  // (feature_selector.cpp:697,724) -- so the larger index takes the slot and the twin does not count as runner-up.
  if (ob > best) {
    second = fmax(fmax(best, second), os);
    best = ob; idx = oidx;
  } else if (ob == best && oidx >= 0 && idx >= 0) {
    idx = max(idx, oidx);
    second = fmax(second, os);
  } else {
    second = fmax(second, fmax(ob, os));
  }
}

// Omega_S += p Delta (feature_selector.cpp:671-681) on the compact state; all threads of one CTA
__device__ void apply_winner(const SelProb& sp, double best, double second, int idx, double p, const double* C,
                             unsigned long long cnt) {
  SelCtrl* c = sp.ctrl;
  if (idx >= 0) {
    for (int e = threadIdx.x; e < sp.TT; e += blockDim.x) sp.R[e] += p * C[e];
  }
  if (threadIdx.x == 0) {
    if (idx >= 0) {
      sp.taken[idx] = 1;
      sp.out_idx[c->n_selected] = idx;
      sp.out_val[c->n_selected] = best;
      c->n_selected++;
      if (second > -1.0) c->min_margin = fmin(c->min_margin, best - second);
    }
    c->scored += cnt;
    c->round++;
    c->ticket = 0;
  }
}

template <int H>
__global__ void __launch_bounds__(32 * SEL_WARPS) sel_round_kernel(SelProb sp) {
  extern __shared__ double sm[];
  const int T = sp.T, TT = sp.TT, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* sR = sm;
  double* A = sR + TT + (size_t)warp * TT;
  double* sred = sm + (size_t)(1 + SEL_WARPS) * TT;   // [SEL_WARPS*4] then [4] broadcast
  __shared__ int s_last;
  for (int e = threadIdx.x; e < TT; e += blockDim.x) sR[e] = sp.R[e];
  __syncthreads();
  const double ld_oo = sp.ctrl->logdet_oo;
  double best = -1.0, second = -INFINITY;   // fMax starts at -1 (feature_selector.cpp:639)
  int bidx = -1;
  double cnt = 0;
  for (int i = sp.c0 + blockIdx.x * SEL_WARPS + warp; i < sp.c1; i += gridDim.x * SEL_WARPS) {
    if (!sp.valid[i] || sp.taken[i]) continue;
    const double p = sp.cand_prob[i];
    const double* Ci = sp.Cc + (size_t)i * TT;
    for (int e = lane; e < TT; e += 32) A[e] = sR[e] + p * Ci[e];
    __syncwarp();
    const double ld = warp_chol_logdet<3 * H>(A, lane);
    const double val = ld_oo + 2.0 * ld;
    cnt += 1;
    if (val > best) { second = best; best = val; bidx = i; }
    else if (val == best && bidx >= 0) bidx = i;      // exact tie = UB collision: the larger index takes the slot (merge_best)
    else if (val > second) second = val;
    __syncwarp();
  }
  if (lane == 0) { sred[warp * 4] = best; sred[warp * 4 + 1] = second; sred[warp * 4 + 2] = (double)bidx; sred[warp * 4 + 3] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double b = -1.0, s = -INFINITY, c = 0;
    int ix = -1;
    for (int q = 0; q < SEL_WARPS; q++) { merge_best(b, s, ix, sred[q * 4], sred[q * 4 + 1], (int)sred[q * 4 + 2]); c += sred[q * 4 + 3]; }
    double* bb = sp.blk_best + (size_t)blockIdx.x * 4;
    bb[0] = b; bb[1] = s; bb[2] = (double)ix; bb[3] = c;
    __threadfence();
    unsigned old = atomicAdd(&sp.ctrl->ticket, 1u);
    s_last = (old == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double* bc = sred + SEL_WARPS * 4;
  {
    // parallel merge of the per-CTA records: (best, idx) has a total order, so the tree is exact
    double b = -1.0, s = -INFINITY, c = 0;
    int ix = -1;
    for (unsigned q = threadIdx.x; q < gridDim.x; q += blockDim.x) {
      const double2 r0 = __ldcg(reinterpret_cast<const double2*>(sp.blk_best) + 2 * q);
      const double2 r1 = __ldcg(reinterpret_cast<const double2*>(sp.blk_best) + 2 * q + 1);
      merge_best(b, s, ix, r0.x, r0.y, (int)r1.x);
      c += r1.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, b, o), os = __shfl_xor_sync(0xffffffffu, s, o);
      const int oi = __shfl_xor_sync(0xffffffffu, ix, o);
      c += __shfl_xor_sync(0xffffffffu, c, o);
      merge_best(b, s, ix, ob, os, oi);
    }
    __syncthreads();   // sred[0 .. SEL_WARPS*4) is free again
    if (lane == 0) { sred[warp * 4] = b; sred[warp * 4 + 1] = s; sred[warp * 4 + 2] = (double)ix; sred[warp * 4 + 3] = c; }
    __syncthreads();
    if (threadIdx.x == 0) {
      b = -1.0; s = -INFINITY; c = 0; ix = -1;
      for (int q = 0; q < SEL_WARPS; q++) { merge_best(b, s, ix, sred[q * 4], sred[q * 4 + 1], (int)sred[q * 4 + 2]); c += sred[q * 4 + 3]; }
      bc[0] = b; bc[1] = s; bc[2] = (double)ix; bc[3] = c;
    }
  }
  __syncthreads();
  const double b = bc[0], s = bc[1];
  const int ix = (int)bc[2];
  const unsigned long long c = (unsigned long long)bc[3];
  if (sp.world == 1) {
    apply_winner(sp, b, s, ix, ix >= 0 ? sp.cand_prob[ix] : 0.0, ix >= 0 ? sp.Cc + (size_t)ix * TT : nullptr, c);
  } else {
    // this rank's winner record for the exchange
    double* r = sp.rec_send;
    if (threadIdx.x == 0) {
      r[0] = b; r[1] = s; r[2] = (double)ix; r[3] = ix >= 0 ? sp.cand_prob[ix] : 0.0;
      sp.ctrl->scored += c;
      sp.ctrl->ticket = 0;
    }
    for (int e = threadIdx.x; e < TT; e += blockDim.x) r[SEL_REC_HDR + e] = ix >= 0 ? sp.Cc[(size_t)ix * TT + e] : 0.0;
  }
}

// ---------------------------------------------------------------------------------------------
// Single-GPU fast path: ALL kappa greedy rounds in one cooperative kernel.  Each warp keeps its
// candidates' packed C_ell in shared memory for the whole selection, every CTA keeps its own copy of R;
// per round: score -> per-CTA record -> ONE grid-wide sync -> every CTA merges the records (same total
// order => same winner everywhere) and applies R += p C_best to its copy.  Records are double-buffered
// by round parity, so one barrier per round is enough.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    while (*reinterpret_cast<volatile unsigned int*>(counter) < target) { }
    __threadfence();
  }
  __syncthreads();
}

// system-scope accesses for the peer-memory exchange (the mailbox of rank r is ordinary device memory of GPU r that
// the other ranks have mapped through CUDA IPC; stores travel over NVLink / NVSwitch)
__device__ __forceinline__ void st_relaxed_sys(double* p, double v) {
  asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

template <int H>
__global__ void __launch_bounds__(32 * SEL_WARPS, 2) sel_persist_kernel(SelProb sp) {
  extern __shared__ double sm[];
  constexpr int T = 3 * H, TT = T * (T + 1) / 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, cpw = sp.cpw;
  double* sR = sm;                                         // [TT]
  const int csm = sp.c_smem;
  const int slots = csm ? cpw : 1;                         // c_smem: all of the warp's blocks stay staged; else one at a time
  double* sC = sR + TT + (size_t)warp * slots * TT;        // [slots][TT]
  double* sred = sm + (size_t)(1 + SEL_WARPS * slots) * TT;   // [SEL_WARPS*4] + [4]
  const int gw = blockIdx.x * SEL_WARPS + warp, nw = gridDim.x * SEL_WARPS;
  // candidates of this warp: index, prob, alive (valid and not yet taken), one per slot k
  for (int k = 0; k < cpw && csm; k++) {
    const int i = sp.c0 + gw + k * nw;
    if (i < sp.c1 && sp.valid[i])
      for (int e = lane; e < TT; e += 32) sC[(size_t)k * TT + e] = sp.Cc[(size_t)i * TT + e];
  }
  for (int e = threadIdx.x; e < TT; e += blockDim.x) sR[e] = sp.R[e];
  unsigned long long alive = 0;                            // bit k: slot k still a candidate (cpw <= 64)
  for (int k = 0; k < cpw; k++) {
    const int i = sp.c0 + gw + k * nw;
    if (i < sp.c1 && sp.valid[i]) alive |= 1ull << k;
  }
  __syncthreads();
  const double ld_oo = sp.ctrl->logdet_oo;
  int n_selected = 0;
  double min_margin = INFINITY;
  unsigned long long scored = 0;
  const bool timer = blockIdx.x == 0 && threadIdx.x == 0;
  unsigned long long t_score = 0, t_barrier = 0, t_exchange = 0, tq = timer ? global_ns() : 0;
  for (int round = 0; round < sp.kappa; round++) {
    double best = -1.0, second = -INFINITY, cnt = 0;
    int bidx = -1;
    for (int k = 0; k < cpw; k++) {
      if (!((alive >> k) & 1ull)) continue;
      const int i = sp.c0 + gw + k * nw;
      const double* Ck = sC + (size_t)(csm ? k : 0) * TT;
      if (!csm) {                                            // coalesced copy from L2 / HBM into the warp's staging slot
        __syncwarp();
        const double* src = sp.Cc + (size_t)i * TT;
        for (int e = lane; e < TT; e += 32) sC[e] = __ldg(src + e);
        __syncwarp();
      }
      const double ld = warp_chol_logdet<T, true>(sR, lane, Ck, sp.cand_prob[i]);
      const double val = ld_oo + 2.0 * ld;
      cnt += 1;
      if (val > best) { second = best; best = val; bidx = i; }
      else if (val == best && bidx >= 0) bidx = i;    // exact tie = UB collision: the larger index takes the slot (merge_best)
      else if (val > second) second = val;
    }
    if (lane == 0) { sred[warp * 4] = best; sred[warp * 4 + 1] = second; sred[warp * 4 + 2] = (double)bidx; sred[warp * 4 + 3] = cnt; }
    __syncthreads();
    double* rec = sp.blk_best + (size_t)(round & 1) * gridDim.x * 4;
    if (threadIdx.x == 0) {
      double b = -1.0, s = -INFINITY, c = 0;
      int ix = -1;
      for (int q = 0; q < SEL_WARPS; q++) { merge_best(b, s, ix, sred[q * 4], sred[q * 4 + 1], (int)sred[q * 4 + 2]); c += sred[q * 4 + 3]; }
      double* bb = rec + (size_t)blockIdx.x * 4;
      __stcg(bb, b); __stcg(bb + 1, s); __stcg(bb + 2, (double)ix); __stcg(bb + 3, c);
    }
    if (timer) { const unsigned long long t = global_ns(); t_score += t - tq; tq = t; }
    grid_barrier(&sp.ctrl->ticket, (unsigned)(round + 1) * gridDim.x);
    if (timer) { const unsigned long long t = global_ns(); t_barrier += t - tq; tq = t; }
    // merge all CTA records (identical on every CTA)
    {
      double b = -1.0, s = -INFINITY, c = 0;
      int ix = -1;
      for (unsigned q = threadIdx.x; q < gridDim.x; q += blockDim.x) {
        const double2 r0 = __ldcg(reinterpret_cast<const double2*>(rec) + 2 * q);
        const double2 r1 = __ldcg(reinterpret_cast<const double2*>(rec) + 2 * q + 1);
        merge_best(b, s, ix, r0.x, r0.y, (int)r1.x);
        c += r1.y;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, b, o), os = __shfl_xor_sync(0xffffffffu, s, o);
        const int oi = __shfl_xor_sync(0xffffffffu, ix, o);
        c += __shfl_xor_sync(0xffffffffu, c, o);
        merge_best(b, s, ix, ob, os, oi);
      }
      if (lane == 0) { sred[warp * 4] = b; sred[warp * 4 + 1] = s; sred[warp * 4 + 2] = (double)ix; sred[warp * 4 + 3] = c; }
      __syncthreads();
      if (threadIdx.x == 0) {
        b = -1.0; s = -INFINITY; c = 0; ix = -1;
        for (int q = 0; q < SEL_WARPS; q++) { merge_best(b, s, ix, sred[q * 4], sred[q * 4 + 1], (int)sred[q * 4 + 2]); c += sred[q * 4 + 3]; }
        double* bc = sred + SEL_WARPS * 4;
        bc[0] = b; bc[1] = s; bc[2] = (double)ix; bc[3] = c;
      }
      __syncthreads();
    }
    if (sp.world > 1) {
      // ---- fused exchange: CTA 0 stores this rank's 32-byte winner record into every rank's mailbox
      //      (peer stores over NVLink), then every CTA of every rank waits for the world's records of this round in
      //      its OWN mailbox and merges them in rank order -> the same winner everywhere, no host, no NCCL call
      double* bcw = sred + SEL_WARPS * 4;
      // epochs run on across selections (epoch_base += kappa per run) so that consecutive exchanges always
      // alternate parity: a rank can be at most one exchange ahead of a peer's reads, two slots are enough
      const unsigned long long ep = sp.epoch_base + (unsigned long long)round + 1ull;
      const int par = (int)(ep & 1ull), world = sp.world;
      if (blockIdx.x == 0 && threadIdx.x < world) {
        const int r = threadIdx.x;
        double* dst = sp.peer_mbox[r] + ((size_t)par * world + sp.rank) * 4;
        st_relaxed_sys(dst, bcw[0]); st_relaxed_sys(dst + 1, bcw[1]); st_relaxed_sys(dst + 2, bcw[2]); st_relaxed_sys(dst + 3, bcw[3]);
        st_release_sys(sp.peer_flag[r] + (size_t)par * world + sp.rank, ep);
      }
      __syncthreads();                                     // bcw consumed before it is overwritten below
      double* sx = sred;                                   // [world][4] (world <= 8 = SEL_WARPS)
      if (threadIdx.x < world) {
        const int r = threadIdx.x;
        const unsigned long long* fl = sp.mflag + (size_t)par * world + r;
        const unsigned long long t0 = global_ns();
        bool ok = true;
        while (ld_acquire_sys(fl) < ep) {
          if (global_ns() - t0 > 3000000000ull) { ok = false; break; }     // 3 s: a peer died; give up, do not hang
        }
        const double* src = sp.mbox + ((size_t)par * world + r) * 4;
        sx[r * 4] = ok ? ld_relaxed_sys(src) : -1.0; sx[r * 4 + 1] = ok ? ld_relaxed_sys(src + 1) : -INFINITY;
        sx[r * 4 + 2] = ok ? ld_relaxed_sys(src + 2) : -1.0; sx[r * 4 + 3] = ok ? ld_relaxed_sys(src + 3) : 0.0;
        if (!ok) sp.ctrl->peer_timeout = 1;
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        double b = -1.0, s = -INFINITY, c = 0;
        int ix = -1;
        for (int r = 0; r < world; r++) { merge_best(b, s, ix, sx[r * 4], sx[r * 4 + 1], (int)sx[r * 4 + 2]); c += sx[r * 4 + 3]; }
        bcw[0] = b; bcw[1] = s; bcw[2] = (double)ix; bcw[3] = c;
      }
      __syncthreads();
      if (timer) { const unsigned long long t = global_ns(); t_exchange += t - tq; tq = t; }
    }
    const double* bc = sred + SEL_WARPS * 4;
    const double wb = bc[0], ws = bc[1];
    const int wix = (int)bc[2];
    scored += (unsigned long long)bc[3];
    if (wix >= 0) {
      // Omega_S += p Delta (feature_selector.cpp:674) on this CTA's copy of R
      const double p = sp.cand_prob[wix];
      const double* Cw = sp.Cc + (size_t)wix * TT;
      for (int e = threadIdx.x; e < TT; e += blockDim.x) sR[e] += p * __ldg(Cw + e);
      // the owner warp retires the candidate
      const int rel = wix - sp.c0 - gw;
      if (rel >= 0 && rel % nw == 0 && rel / nw < cpw) alive &= ~(1ull << (rel / nw));
      if (blockIdx.x == 0 && threadIdx.x == 0) { sp.taken[wix] = 1; sp.out_idx[n_selected] = wix; sp.out_val[n_selected] = wb; }
      n_selected++;
      if (ws > -1.0) min_margin = fmin(min_margin, wb - ws);
    }
    __syncthreads();
  }
  if (blockIdx.x == 0) {
    for (int e = threadIdx.x; e < TT; e += blockDim.x) sp.R[e] = sR[e];
    if (threadIdx.x == 0) {
      SelCtrl* c = sp.ctrl;
      c->n_selected = n_selected; c->min_margin = min_margin; c->scored = scored; c->round = sp.kappa;
      c->t_score = t_score; c->t_barrier = t_barrier; c->t_exchange = t_exchange;   // (merge + apply are booked with the next score)
    }
  }
}

// multi-GPU: every rank holds the same gathered records and applies the same update
__global__ void __launch_bounds__(256) sel_apply_kernel(SelProb sp) {
  __shared__ double bc[4];
  const int RS = SEL_REC_HDR + sp.TT;
  if (threadIdx.x == 0) {
    double b = -1.0, s = -INFINITY;
    int ix = -1, who = -1;
    for (int r = 0; r < sp.world; r++) {
      const double* rec = sp.rec_all + (size_t)r * RS;
      int before = ix;
      merge_best(b, s, ix, rec[0], rec[1], (int)rec[2]);
      if (ix != before) who = r;
    }
    bc[0] = b; bc[1] = s; bc[2] = (double)ix; bc[3] = (double)who;
  }
  __syncthreads();
  const int ix = (int)bc[2], who = (int)bc[3];
  const double* rec = sp.rec_all + (size_t)(who < 0 ? 0 : who) * RS;
  apply_winner(sp, bc[0], bc[1], ix, rec[3], rec + SEL_REC_HDR, 0ull);
}

template <int H>
__global__ void __launch_bounds__(32) sel_final_kernel(SelProb sp) {
  extern __shared__ double sm[];
  for (int e = threadIdx.x; e < sp.TT; e += 32) sm[e] = sp.R[e];
  __syncwarp();
  double ld = warp_chol_logdet<3 * H>(sm, threadIdx.x);
  if (threadIdx.x == 0) sp.ctrl->final_logdet = sp.ctrl->logdet_oo + 2.0 * ld;
}

__global__ void sel_expand_kernel(SelProb sp, double* Cfull) {
  const int T = sp.T;
  const size_t total = (size_t)sp.N * T * T;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    int f = (int)(e / (T * T)), rc = (int)(e - (size_t)f * T * T), r = rc / T, c = rc - r * T;
    double v = 0;
    if (f >= sp.b0 && f < sp.b1 && sp.valid[f]) v = sp.Cc[(size_t)f * sp.TT + (r >= c ? tri(r, c) : tri(c, r))];
    Cfull[e] = v;
  }
}

// multi-GPU finalize: dir 0 packs (scored, n_valid) for the all-reduce, dir 1 writes the sums back
__global__ void sel_counts_kernel(SelProb sp, unsigned long long* buf, int dir) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (dir == 0) { buf[0] = sp.ctrl->scored; buf[1] = (unsigned long long)sp.ctrl->n_valid; }
  else { sp.ctrl->scored = buf[0]; sp.ctrl->n_valid = (int)buf[1]; }
}
int sel_launch_counts(const SelProb& sp, unsigned long long* buf, int dir, cudaStream_t st) {
  sel_counts_kernel<<<1, 32, 0, st>>>(sp, buf, dir);
  return 1;
}

// =============================================================================================
size_t sel_omega_smem_bytes(int H) {
  int D = 9 * (H + 1);
  return sizeof(double) * (size_t)D * D + sizeof(int) * (size_t)D + 16;
}
static size_t round_smem(int TT) { return sizeof(double) * ((size_t)(1 + SEL_WARPS) * TT + SEL_WARPS * 4 + 4); }

#define BVIO_SEL_FOR_EACH_H(M) M(1) M(2) M(3) M(4) M(5) M(6) M(7) M(8) M(9) M(10) M(11) M(12) M(13) M(14) M(15) M(16)
// the shared-memory opt-ins are per device (several contexts on different GPUs may live in one process)
static std::mutex g_sel_cfg_mutex;
static unsigned long long g_sel_cfg_devices[4] = {0, 0, 0, 0};
int sel_configure(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return (int)cudaGetLastError();
  std::lock_guard<std::mutex> lock(g_sel_cfg_mutex);
  if (dev >= 0 && dev < 256 && ((g_sel_cfg_devices[dev >> 6] >> (dev & 63)) & 1ull)) return 0;
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(sel_omega_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)) != cudaSuccess) return e;
#define BVIO_SEL_ATTR(HH) \
  if ((e = cudaFuncSetAttribute(sel_round_kernel<HH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)) != cudaSuccess) return e;
  BVIO_SEL_FOR_EACH_H(BVIO_SEL_ATTR)
#undef BVIO_SEL_ATTR
#define BVIO_SEL_ATTR(HH) \
  if ((e = cudaFuncSetAttribute(sel_persist_kernel<HH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)) != cudaSuccess) return e;
  BVIO_SEL_FOR_EACH_H(BVIO_SEL_ATTR)
#undef BVIO_SEL_ATTR
  if (dev >= 0 && dev < 256) g_sel_cfg_devices[dev >> 6] |= 1ull << (dev & 63);
  return 0;
}
int sel_launch_reset(const SelProb& sp, cudaStream_t st) {
  int blocks = (sp.N + 255) / 256;
  if (blocks < 1) blocks = 1;
  sel_reset_kernel<<<blocks, 256, 0, st>>>(sp);
  return 1;
}
int sel_launch_build(const SelProb& sp, cudaStream_t st) {
  int nf = (sp.b1 - sp.b0) + sp.U, n = 0;
  if (nf > 0) { sel_build_kernel<<<(nf + SEL_WARPS - 1) / SEL_WARPS, 32 * SEL_WARPS, 0, st>>>(sp); n++; }
  sel_omega_kernel<<<1, 256, sel_omega_smem_bytes(sp.H), st>>>(sp);
  return n + 1;
}
int sel_launch_round(const SelProb& sp, cudaStream_t st) {
  switch (sp.H) {
#define BVIO_SEL_CASE(HH) case HH: sel_round_kernel<HH><<<sp.grid_round, 32 * SEL_WARPS, round_smem(sp.TT), st>>>(sp); break;
    BVIO_SEL_FOR_EACH_H(BVIO_SEL_CASE)
#undef BVIO_SEL_CASE
    default: break;
  }
  return 1;
}
static size_t persist_smem(int TT, int cpw, int c_smem) {
  return sizeof(double) * ((size_t)(1 + SEL_WARPS * (c_smem ? cpw : 1)) * TT + SEL_WARPS * 4 + 4);
}

// Picks the grid of the persistent kernel: every CTA must be co-resident (the kernel spins on a grid
// barrier).  Small candidate sets keep their information blocks in shared memory (cpw <= 3 at H = 10); larger ones
// (up to 64 candidates per warp: ~150 000 candidates per GPU) read them from L2 / HBM every round.
// Returns 0 when the problem does not fit (the caller then runs one kernel per round).
int sel_plan_persist(SelProb& sp, int sm_count) {
  const int nloc = sp.c1 - sp.c0;
  sp.grid_persist = 0; sp.cpw = 0; sp.c_smem = 1;
  if ((sp.world != 1 && !sp.fused) || nloc <= 0 || sp.kappa <= 0) return 0;
  int per_sm = 0;
  cudaError_t e = cudaSuccess;
  int want = (nloc + SEL_WARPS - 1) / SEL_WARPS;
  for (int pass = 0; pass < 2; pass++) {
    const int c_smem = pass == 0;
    for (int cpw = 1; cpw <= (c_smem ? 8 : 64); cpw++) {
      size_t smem = persist_smem(sp.TT, cpw, c_smem);
      if (smem > 100 * 1024) break;
      switch (sp.H) {
#define BVIO_SEL_CASE(HH) case HH: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sel_persist_kernel<HH>, 32 * SEL_WARPS, smem); break;
        BVIO_SEL_FOR_EACH_H(BVIO_SEL_CASE)
#undef BVIO_SEL_CASE
        default: return 0;
      }
      if (e != cudaSuccess || per_sm < 1) return 0;
      int cap = per_sm * sm_count;
      int need = (want + cpw - 1) / cpw;
      if (need <= cap) { sp.grid_persist = need; sp.cpw = cpw; sp.c_smem = c_smem; return 1; }
      if (!c_smem) {
        // with the blocks in L2 the grid is the whole machine: pick the smallest cpw that covers the candidates
        int cpw_need = (want + cap - 1) / cap;
        if (cpw_need > 64) return 0;
        cpw = cpw_need - 1;   // loop increment lands on it
      }
    }
  }
  return 0;
}
int sel_launch_persist(const SelProb& sp, cudaStream_t st) {
  size_t smem = persist_smem(sp.TT, sp.cpw, sp.c_smem);
  switch (sp.H) {
  // cooperative launch: the driver guarantees (or refuses) co-residency of all CTAs, which the grid
  // barrier inside the kernel relies on
#define BVIO_SEL_CASE(HH) case HH: { SelProb arg = sp; void* args[] = {&arg}; \
    cudaLaunchCooperativeKernel((const void*)sel_persist_kernel<HH>, dim3(sp.grid_persist), dim3(32 * SEL_WARPS), args, smem, st); } break;
    BVIO_SEL_FOR_EACH_H(BVIO_SEL_CASE)
#undef BVIO_SEL_CASE
    default: break;
  }
  return 1;
}
int sel_launch_apply(const SelProb& sp, cudaStream_t st) {
  sel_apply_kernel<<<1, 256, 0, st>>>(sp);
  return 1;
}
int sel_launch_final(const SelProb& sp, cudaStream_t st) {
  switch (sp.H) {
#define BVIO_SEL_CASE(HH) case HH: sel_final_kernel<HH><<<1, 32, sizeof(double) * sp.TT, st>>>(sp); break;
    BVIO_SEL_FOR_EACH_H(BVIO_SEL_CASE)
#undef BVIO_SEL_CASE
    default: break;
  }
  return 1;
}
int sel_launch_expand(const SelProb& sp, double* Cfull, cudaStream_t st) {
  sel_expand_kernel<<<296, 256, 0, st>>>(sp, Cfull);
  return 1;
}

}  // namespace bvio
