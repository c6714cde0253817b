// fp64_peak.cu -- measured FP64 ceilings of the GPU this runs on: DFMA (vector pipe) and
// mma.sync.m8n8k4.f64 (DMMA).  Used as the compute roofline denominator beside the HBM one.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double* out, int iters) {
  double a[8], b = 1.0000001, c = 1e-9;
  for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = fma(a[i], b, c);
  }
  double s = 0;
  for (int i = 0; i < 8; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dmma_kernel(double* out, int iters) {
  double c[4][2];
  for (int i = 0; i < 4; i++) c[i][0] = c[i][1] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-6, b = 1.0 - threadIdx.x * 1e-6;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
  for (int i = 0; i < 4; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int blocks = p.multiProcessorCount * 4, threads = 512, iters = 20000;
  double* out;
  cudaMalloc(&out, sizeof(double) * blocks * threads);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  double best_fma = 0, best_mma = 0;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0); dfma_kernel<<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    double tf = 2.0 * 8 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
    if (tf > best_fma) best_fma = tf;
    cudaEventRecord(e0); dmma_kernel<<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    double tm = 2.0 * 256 * 4 * iters * (double)blocks * (threads / 32) / (ms * 1e-3) / 1e12;
    if (tm > best_mma) best_mma = tm;
  }
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"dfma_tflops\": %.3f, \"dmma_tflops\": %.3f}\n", p.name, p.multiProcessorCount, best_fma, best_mma);
  return 0;
}
