// Does the DFMA rate on B200 depend on how many DISTINCT 64-bit register operands an instruction
// reads?  fp64_rate.cu measured a[i] = fma(a[i], x, y) (x, y loop invariant: operand reuse cache),
// the EM kernels issue acc = fma(cw, psi, acc) with three different registers every time.
//   V0: a[i] = fma(a[i], x, y)        one new 64-bit register per instruction
//   V1: a[i] = fma(b[i], x, a[i])     two
//   V2: a[i] = fma(b[i], c[i], a[i])  three, all different for consecutive instructions
//   V3: a[i] = fma(b[i], c[j], a[i])  three, c[j] shared by 4 consecutive instructions (reuse possible)
//   V4: DMUL d[i] = b[i] * c[i] ; a[i] += d[i]   (mul + add pairs)
#include <cstdio>
#include <cuda_runtime.h>
template <int V>
__global__ void __launch_bounds__(512, 1) k(double* out, const double* in, long long* cyc, int iters) {
  double a[8], b[8], c[8];
  for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 1e-3 + i; b[i] = in[i] + threadIdx.x * 1e-9; c[i] = in[8 + i] + threadIdx.x * 1e-9; }
  const double x = in[16], y = in[17];
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (V == 0) a[i] = fma(a[i], x, y);
        if (V == 1) a[i] = fma(b[i], x, a[i]);
        if (V == 2) a[i] = fma(b[i], c[(i + r) & 7], a[i]);
        if (V == 3) a[i] = fma(b[i], c[(i >> 2) + 2 * r], a[i]);
        if (V == 4) { if (i & 1) a[i] = b[i] * c[(i + r) & 7]; else a[i] += a[i + 1]; }
      }
    }
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int V> void run(const char* name, int threads) {
  double *out, *in; long long* cyc;
  cudaMalloc(&out, 8 * 148 * 512); cudaMalloc(&in, 8 * 32); cudaMalloc(&cyc, 8 * 148);
  double h[18]; for (int i = 0; i < 18; ++i) h[i] = 1.0 + 1e-7 * i; h[17] = 1e-9;
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  const int iters = 20000;
  k<V><<<148, threads>>>(out, in, cyc, 100);
  k<V><<<148, threads>>>(out, in, cyc, iters);
  cudaDeviceSynchronize();
  long long hc[148]; cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += hc[i]; c /= 148;
  printf("%-58s %2d warps/SM  %.1f fp64 lane-ops/clk/SM\n", name, threads / 32, (double)iters * 32 * threads / c);
  cudaFree(out); cudaFree(in); cudaFree(cyc);
}
int main() {
  for (int th : {512, 256, 128}) {
    run<0>("V0 fma(a, x, y): 1 new register operand", th);
    run<1>("V1 fma(b, x, a): 2 register operands + 1 invariant", th);
    run<2>("V2 fma(b, c, a): 3 different register operands", th);
    run<3>("V3 fma(b, c', a): 3 operands, c' shared by 4 in a row", th);
    run<4>("V4 dmul / dadd pairs", th);
  }
  return 0;
}
