// fp64_lat.cu -- issue-to-issue cycles of dependent and independent DFMA / DMMA (mma.sync.m8n8k4.f64) / shared-memory
// loads on the GPU this runs on, for one SM: warps per CTA x independent chains per warp.  These set the latency floors of
// the BA kernels (pivot chains, Gram-product chains).
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_lat fp64_lat.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NC>
__global__ void dmma_chain(double* out, long long* cyc, int iters) {
  double c[NC][2];
  for (int i = 0; i < NC; i++) c[i][0] = c[i][1] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-6, b = 1.0 - threadIdx.x * 1e-6;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NC; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < NC; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int NC>
__global__ void dfma_chain(double* out, long long* cyc, int iters) {
  double c[NC];
  for (int i = 0; i < NC; i++) c[i] = threadIdx.x * 1e-3 + i;
  double b = 1.0000001, d = 1e-9;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NC; i++) c[i] = fma(c[i], b, d);
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < NC; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void lds_chain(double* out, long long* cyc, int iters) {
  __shared__ int nxt[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) nxt[i] = (i + 33) & 1023;
  __syncthreads();
  int p = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) p = nxt[p];
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = p;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  double* out; long long* cyc; long long h;
  cudaMalloc(&out, sizeof(double) * 1024); cudaMalloc(&cyc, 64);
  const int iters = 4096;
  printf("{");
#define RUN(name, kern, threads, per_iter) \
  kern<<<1, threads>>>(out, cyc, iters); kern<<<1, threads>>>(out, cyc, iters); cudaDeviceSynchronize(); \
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("\"%s\": %.1f, ", name, (double)h / iters / (per_iter));
  RUN("dmma_1warp_1chain", dmma_chain<1>, 32, 1)
  RUN("dmma_1warp_2chains", dmma_chain<2>, 32, 2)
  RUN("dmma_1warp_4chains", dmma_chain<4>, 32, 4)
  RUN("dmma_1warp_8chains", dmma_chain<8>, 32, 8)
  RUN("dmma_4warps_1chain", dmma_chain<1>, 128, 1)
  RUN("dmma_4warps_4chains", dmma_chain<4>, 128, 4)
  RUN("dmma_8warps_1chain", dmma_chain<1>, 256, 1)
  RUN("dmma_8warps_2chains", dmma_chain<2>, 256, 2)
  RUN("dmma_8warps_6chains", dmma_chain<6>, 256, 6)
  RUN("dmma_16warps_4chains", dmma_chain<4>, 512, 4)
  RUN("dfma_1warp_1chain", dfma_chain<1>, 32, 1)
  RUN("dfma_1warp_4chains", dfma_chain<4>, 32, 4)
  RUN("dfma_8warps_1chain", dfma_chain<1>, 256, 1)
  RUN("dfma_8warps_4chains", dfma_chain<4>, 256, 4)
  RUN("dfma_16warps_8chains", dfma_chain<8>, 512, 8)
  RUN("lds_dependent", lds_chain, 32, 1)
  printf("\"unit\": \"cycles per instruction per warp (issue-to-issue)\"}\n");
  return 0;
}
