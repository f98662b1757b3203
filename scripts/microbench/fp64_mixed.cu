// Do DFMA (vector fp64) and DMMA (fp64 mma.sync m8n8k4) share one pipe on B200, or do they add up?
#include <cstdio>
#include <cuda_runtime.h>
template <int NF, int NM>
__global__ void mixed(double* out, int iters) {
  double a[8]; double c[4][2];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  for (int i = 0; i < 4; ++i) { c[i][0] = 0; c[i][1] = 0; }
  double x = 1.0000001, y = 1e-9, ma = threadIdx.x * 1e-3, mb = 1.0 + threadIdx.x * 1e-4;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (NF) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fma(a[i], x, y);
      }
      if (NM) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c[r][0]), "+d"(c[r][1]) : "d"(ma), "d"(mb));
      }
    }
  }
  double s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NF, int NM> void run(const char* name) {
  double* out; cudaMalloc(&out, 8 * 148 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int iters = 20000, blocks = 148, threads = 512; float ms;
  mixed<NF, NM><<<blocks, threads>>>(out, 100);
  cudaEventRecord(e0); mixed<NF, NM><<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  cudaEventElapsedTime(&ms, e0, e1);
  double ffl = NF ? 2.0 * 32 * iters * (double)blocks * threads : 0;          // 4 x 8 DFMA per iteration per thread
  double mfl = NM ? 2.0 * 256 * 4 * iters * (double)blocks * (threads / 32) : 0;  // 4 DMMA per iteration per warp
  printf("%-22s %.3f ms  DFMA %.1f TF + DMMA %.1f TF = %.1f TFLOP/s\n", name, ms, ffl / ms * 1e-9, mfl / ms * 1e-9, (ffl + mfl) / ms * 1e-9);
  cudaFree(out);
}
int main() { run<1, 0>("DFMA only"); run<0, 1>("DMMA only"); run<1, 1>("DFMA + DMMA interleaved"); return 0; }
