// C-ABI entry points for the batched small-matrix linear algebra (see include/pbb.h).
#include "common.cuh"
#include "heig.cuh"

namespace pbb {

// one warp per matrix
__global__ void heig_batched_kernel(const double2* __restrict__ a, int n, int D, double* __restrict__ w,
                                    double2* __restrict__ v, int* status, int warps) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x * warps + warp;
  if (m >= n) return;
  const size_t per = (jacobi_smem_bytes(D) + 15) & ~(size_t)15;
  double2* A = reinterpret_cast<double2*>(smem_raw + per * warp);
  double2* V = A + D * D;
  double* rot = reinterpret_cast<double*>(V + D * D);
  const double2* __restrict__ am = a + (size_t)m * D * D;
  bool bad = false;
  // Hermitian part of the input, (A + A^H) / 2: LAPACK reads one triangle only
  for (int i = lane; i < D * D; i += 32) {
    const int r = i / D, c = i - r * D;
    const double2 x = am[r * D + c], y = am[c * D + r];
    const double2 h = make_double2(0.5 * (x.x + y.x), r == c ? 0.0 : 0.5 * (x.y - y.y));
    bad |= !isfinite(h.x) || !isfinite(h.y);
    A[i] = h;
  }
  __syncwarp();
  const int sweeps = warp_jacobi(A, V, rot, D, lane);
  if ((__any_sync(0xffffffffu, bad) || sweeps > kJacobiMaxSweeps) && lane == 0 && status) atomicMax(status, m + 1);
  for (int x = lane; x < D; x += 32) {
    const int r = eig_rank(A, D, x);
    w[(size_t)m * D + r] = A[x * D + x].x;
    for (int d = 0; d < D; ++d) v[(size_t)m * D * D + d * D + r] = V[d * D + x];
  }
}

}  // namespace pbb

using namespace pbb;

extern "C" int pbb_heig_batched(const void* a, int n, int D, double* w, void* v, int* status, void* stream) {
  PBB_CHECK_ARG(a != nullptr, 1, "a is null");
  PBB_CHECK_ARG(n > 0, 2, "n must be positive");
  PBB_CHECK_ARG(D > 0 && D <= 64, 3, "need 0 < D <= 64");
  PBB_CHECK_ARG(w != nullptr, 4, "w is null");
  PBB_CHECK_ARG(v != nullptr, 5, "v is null");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t per = (jacobi_smem_bytes(D) + 15) & ~(size_t)15;
  int warps = (int)((size_t)(96 * 1024) / per);
  if (warps > 4) warps = 4;
  if (warps < 1) warps = 1;
  PBB_CUDA(cudaFuncSetAttribute(heig_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  heig_batched_kernel<<<(n + warps - 1) / warps, 32 * warps, per * warps, st>>>(
      reinterpret_cast<const double2*>(a), n, D, w, reinterpret_cast<double2*>(v), status, warps);
  PBB_CUDA(cudaGetLastError());
  return 0;
}
