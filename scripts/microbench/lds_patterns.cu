// Shared-memory load cost on B200 for the access patterns of the EM kernels: how many cycles of the
// SM's shared-memory pipe does one warp-wide LDS take when (a) every lane reads the same 16 bytes
// (coefficient broadcast), (b) the two half-warps read two different 16-byte words, (c) the lanes
// read 32 consecutive 16-byte words (observation rows, lane = frame), (d) the lanes pick one of the
// eight 16-byte words of ONE 128-byte row (observation row, lane = slot), (e) 8-byte variants, and
// what is left of the DFMA rate when such loads are interleaved with an fp64 stream.
// Output: cycles per LDS instruction per SM (16 resident warps, throughput bound).
#include <cstdio>
#include <cuda_runtime.h>

// the loaded words are folded with integer XORs (ALU pipe), so the fp64 pipe is free for the DFMA stream
__device__ __forceinline__ void lds128(unsigned& x, unsigned& y, unsigned addr) {
  unsigned a, b, c, d;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr));
  x ^= a ^ c; y ^= b ^ d;
}
__device__ __forceinline__ void lds64(unsigned& x, unsigned& y, unsigned addr) {
  unsigned a, b;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(addr));
  x ^= a; y ^= b;
}

// MODE: 0 uniform 128, 1 two addresses 128 (half-warps), 2 consecutive 128 (512 B), 3 eight words of one
// row 128, 4 uniform 64, 5 consecutive 64 (256 B), 6 eight distinct 8-byte words stride 24 B (64-bit),
// 7 four rows x eight words 128 (lane = frame & 3 rows: 4 x 128 B), 8 uniform 128 with 4 distinct rows
template <int MODE, int NFMA>
__global__ void __launch_bounds__(512, 1) k(double* out, long long* cyc, int iters) {
  extern __shared__ __align__(128) unsigned char sm[];
  double* s = reinterpret_cast<double*>(sm);
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) s[i] = 1.0 + i * 1e-6;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned base = (unsigned)__cvta_generic_to_shared(sm) + warp * 2048;
  unsigned addr;
  if (MODE == 0) addr = base;
  else if (MODE == 1) addr = base + (lane >> 4) * 144;
  else if (MODE == 2) addr = base + lane * 16;
  else if (MODE == 3) addr = base + ((lane * 5) & 7) * 16;
  else if (MODE == 4) addr = base;
  else if (MODE == 5) addr = base + lane * 8;
  else if (MODE == 6) addr = base + (lane & 7) * 24;
  else if (MODE == 7) addr = base + (lane & 3) * 128 + (((lane >> 2) ^ (lane & 3)) & 7) * 16;
  else addr = base + (lane >> 3) * 144;
  unsigned a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  double f[8];
  for (int i = 0; i < 8; ++i) f[i] = lane * 1e-3 + i;
  const double xm = 1.0000001, ym = 1e-9;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (MODE == 4 || MODE == 5 || MODE == 6) { if (r & 1) lds64(a0, a1, addr + r * 256); else lds64(a2, a3, addr + r * 256); }
      else { if (r & 1) lds128(a0, a1, addr + r * 256); else lds128(a2, a3, addr + r * 256); }
#pragma unroll
      for (int j = 0; j < NFMA; ++j) f[(r * NFMA + j) & 7] = fma(f[(r * NFMA + j) & 7], xm, ym);
    }
  }
  long long t1 = clock64();
  __syncthreads();
  double sum = (double)(a0 ^ a1 ^ a2 ^ a3);
  for (int i = 0; i < 8; ++i) sum += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE, int NFMA>
void run(const char* name) {
  double* out; long long* cyc;
  cudaMalloc(&out, 8 * 148 * 512); cudaMalloc(&cyc, 8 * 148);
  const int iters = 4000, blocks = 148, threads = 512;
  cudaFuncSetAttribute(k<MODE, NFMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  k<MODE, NFMA><<<blocks, threads, 65536>>>(out, cyc, 10);
  k<MODE, NFMA><<<blocks, threads, 65536>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  const double nlds = (double)iters * 8 * (threads / 32);  // LDS instructions per SM
  const double nfma = nlds * NFMA;
  printf("%-44s NFMA=%d  %.2f cycles per LDS per SM", name, NFMA, c / nlds);
  if (NFMA) printf("   DFMA %.1f lanes/clk/SM", nfma * 32 / c);
  printf("\n");
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0, 0>("LDS.128 uniform (1 word)");
  run<1, 0>("LDS.128 two words (half-warps)");
  run<8, 0>("LDS.128 four words (quarter-warps)");
  run<2, 0>("LDS.128 32 consecutive words (512 B)");
  run<3, 0>("LDS.128 8 words of one 128 B row");
  run<7, 0>("LDS.128 4 rows x 8 words");
  run<4, 0>("LDS.64 uniform");
  run<5, 0>("LDS.64 32 consecutive (256 B)");
  run<6, 0>("LDS.64 8 words stride 24 B");
  run<0, 2>("LDS.128 uniform + 2 DFMA");
  run<0, 4>("LDS.128 uniform + 4 DFMA");
  run<0, 8>("LDS.128 uniform + 8 DFMA");
  run<2, 4>("LDS.128 consecutive + 4 DFMA");
  run<2, 8>("LDS.128 consecutive + 8 DFMA");
  run<3, 4>("LDS.128 8-of-row + 4 DFMA");
  run<3, 8>("LDS.128 8-of-row + 8 DFMA");
  run<4, 4>("LDS.64 uniform + 4 DFMA");
  return 0;
}
