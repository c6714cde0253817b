// common.cuh -- small FP64 device helpers shared by the BA and selector kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define BVIO_KMAX 16          // max keyframes in a window (reference: WINDOW_SIZE+1 = 11)
#define BVIO_HMAX 16          // max selector horizon (reference: HORIZON = 13)

namespace bvio {

struct d3 { double x, y, z; };
__device__ __forceinline__ d3 operator+(d3 a, d3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ d3 operator-(d3 a, d3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ d3 operator*(double s, d3 a) { return {s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ double dot3(d3 a, d3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ d3 cross3(d3 a, d3 b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// row-major 3x3 in a flat array
__device__ __forceinline__ d3 mv3(const double* m, d3 v) {
  return {m[0] * v.x + m[1] * v.y + m[2] * v.z, m[3] * v.x + m[4] * v.y + m[5] * v.z,
          m[6] * v.x + m[7] * v.y + m[8] * v.z};
}
__device__ __forceinline__ d3 mtv3(const double* m, d3 v) {  // m^T v
  return {m[0] * v.x + m[3] * v.y + m[6] * v.z, m[1] * v.x + m[4] * v.y + m[7] * v.z,
          m[2] * v.x + m[5] * v.y + m[8] * v.z};
}
__device__ __forceinline__ void mm3(const double* a, const double* b, double* c) {
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) c[i * 3 + j] = a[i * 3] * b[j] + a[i * 3 + 1] * b[3 + j] + a[i * 3 + 2] * b[6 + j];
}
__device__ __forceinline__ void mtm3(const double* a, const double* b, double* c) {  // a^T b
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) c[i * 3 + j] = a[i] * b[j] + a[3 + i] * b[3 + j] + a[6 + i] * b[6 + j];
}
__device__ __forceinline__ void mmt3(const double* a, const double* b, double* c) {  // a b^T
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++)
      c[i * 3 + j] = a[i * 3] * b[j * 3] + a[i * 3 + 1] * b[j * 3 + 1] + a[i * 3 + 2] * b[j * 3 + 2];
}

struct q4 { double x, y, z, w; };
__device__ __forceinline__ q4 qmul(q4 a, q4 b) {
  return {a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y, a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
          a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x, a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z};
}
__device__ __forceinline__ q4 qinv(q4 q) {
  double n2 = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
  return {-q.x / n2, -q.y / n2, -q.z / n2, q.w / n2};
}
__device__ __forceinline__ q4 qnormalized(q4 q) {
  double n = sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  return {q.x / n, q.y / n, q.z / n, q.w / n};
}
__device__ __forceinline__ d3 qrot(q4 q, d3 v) {  // Eigen _transformVector
  d3 u{q.x, q.y, q.z};
  d3 uv = cross3(u, v);
  uv = uv + uv;
  return v + q.w * uv + cross3(u, uv);
}
__device__ __forceinline__ void qmat(q4 q, double* r) {  // Eigen toRotationMatrix, row-major
  double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  r[0] = 1 - (tyy + tzz); r[1] = txy - twz; r[2] = txz + twy;
  r[3] = txy + twz; r[4] = 1 - (txx + tzz); r[5] = tyz - twx;
  r[6] = txz - twy; r[7] = tyz + twx; r[8] = 1 - (txx + tyy);
}
__device__ __forceinline__ q4 qslerp(q4 a, double t, q4 b) {  // Eigen slerp
  const double one = 1.0 - 2.220446049250313e-16;
  double d = a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
  double ad = fabs(d), s0, s1;
  if (ad >= one) { s0 = 1.0 - t; s1 = t; }
  else {
    double th = acos(ad), st = sin(th);
    s0 = sin((1.0 - t) * th) / st;
    s1 = sin(t * th) / st;
  }
  if (d < 0) s1 = -s1;
  return {s0 * a.x + s1 * b.x, s0 * a.y + s1 * b.y, s0 * a.z + s1 * b.z, s0 * a.w + s1 * b.w};
}
// PoseLocalParameterization::Plus (pose_local_parameterization.cpp:3-18)
__device__ __forceinline__ void pose_plus(const double* x, const double* d, double* out) {
  out[0] = x[0] + d[0]; out[1] = x[1] + d[1]; out[2] = x[2] + d[2];
  q4 q = qnormalized(qmul(q4{x[3], x[4], x[5], x[6]}, q4{d[3] / 2, d[4] / 2, d[5] / 2, 1.0}));
  out[3] = q.x; out[4] = q.y; out[5] = q.z; out[6] = q.w;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// deterministic block reduction (fixed order); red must hold blockDim/32 doubles
__device__ __forceinline__ double block_sum(double v, double* red) {
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  double s = 0;
  for (int i = 0; i < nw; i++) s += red[i];
  return s;
}
__device__ __forceinline__ double block_max(double v, double* red) {
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  double s = red[0];
  for (int i = 1; i < nw; i++) s = fmax(s, red[i]);
  return s;
}

}  // namespace bvio
