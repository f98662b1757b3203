// DFMA latency / ILP requirement on B200: one CTA per SM, W warps, C independent chains per thread.
#include <cstdio>
#include <cuda_runtime.h>
template <int C>
__global__ void chains(double* out, int iters, long long* cyc) {
  double a[C];
  for (int i = 0; i < C; ++i) a[i] = threadIdx.x * 1e-3 + i;
  const double x = 1.0000001, y = 1e-9;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < C; ++i) a[i] = fma(a[i], x, y);
  }
  long long t1 = clock64();
  double s = 0; for (int i = 0; i < C; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (blockIdx.x == 0 && threadIdx.x == 0) *cyc = t1 - t0;
}
template <int C> void run(int warps_per_smsp) {
  int sms = 148; double* out; cudaMalloc(&out, 8 * 148 * 1024); long long* cyc; cudaMallocManaged(&cyc, 8);
  int threads = 32 * 4 * warps_per_smsp; int iters = 2000;
  chains<C><<<sms, threads>>>(out, 10, cyc); cudaDeviceSynchronize();
  chains<C><<<sms, threads>>>(out, iters, cyc); cudaDeviceSynchronize();
  double per = (double)*cyc / (iters * 8.0 * C);   // cycles per DFMA per warp
  double util = warps_per_smsp * 2.0 / (per * 1.0);  // fraction of the 2-cycle issue rate used per SMSP... per-warp DFMA rate * warps * 2 cycles
  printf("chains=%2d warps/SMSP=%d: %.2f cycles per DFMA per warp -> pipe utilisation %.0f%%\n", C, warps_per_smsp, per, 100.0 * warps_per_smsp * 2.0 / per / 1.0 / (1.0) > 100 ? 100.0 : 100.0 * warps_per_smsp * 2.0 / per);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int w = 1; w <= 4; ++w) { run<1>(w); run<2>(w); run<3>(w); run<4>(w); run<6>(w); run<8>(w); run<12>(w); }
  return 0;
}
