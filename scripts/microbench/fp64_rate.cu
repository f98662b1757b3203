// Microbenchmark: fp64 FMA (DFMA) and DMMA (mma.sync m8n8k4 f64) throughput on B200.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma_kernel(double* out, int iters) {
  double a[16];
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3 + i;
  double x = 1.0000001, y = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fma(a[i], x, y);
  }
  double s = 0; for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void ffma_kernel(float* out, int iters) {
  float a[16];
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3f + i;
  float x = 1.0000001f, y = 1e-9f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], x, y);
  }
  float s = 0; for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void dmma_kernel(double* out, int iters) {
  double c[8][2];
  for (int i = 0; i < 8; ++i) { c[i][0] = 0; c[i][1] = 0; }
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0; for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int warps = 4; warps <= 32; warps *= 2) {
    int threads = 32 * warps; if (threads > 1024) threads = 1024;
    int blocks = sms * (warps > 32 ? 2 : 1);
    int iters = 20000; float ms;
    dfma_kernel<<<blocks, threads>>>(out, 100);
    cudaEventRecord(e0); dfma_kernel<<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    double fl = 2.0 * 16 * iters * (double)blocks * threads;
    printf("DFMA  warps/SM=%2d: %.3f ms  %.2f TFLOP/s  (%.1f FMA/clk/SM at 1.9GHz)\n", warps, ms, fl / ms * 1e-9, fl / 2 / (ms * 1e-3) / sms / 1.9e9);
    ffma_kernel<<<blocks, threads>>>((float*)out, 100);
    cudaEventRecord(e0); ffma_kernel<<<blocks, threads>>>((float*)out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("FFMA  warps/SM=%2d: %.3f ms  %.2f TFLOP/s\n", warps, ms, fl / ms * 1e-9);
    dmma_kernel<<<blocks, threads>>>(out, 100);
    cudaEventRecord(e0); dmma_kernel<<<blocks, threads>>>(out, iters / 4); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    double flm = 2.0 * 8 * 8 * 4 * 8 * (iters / 4) * (double)blocks * warps;
    printf("DMMA  warps/SM=%2d: %.3f ms  %.2f TFLOP/s\n", warps, ms, flm / ms * 1e-9);
  }
  return 0;
}
