// C-ABI entry points for the beamforming side: batched small-matrix linear algebra,
// PSD estimation and beamformer application (see include/pbb.h).
#include "common.cuh"
#include <cstring>
#include "em_args.cuh"
#include "heig.cuh"
#include "linalg_kernels.cuh"
#include "prof.cuh"

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
  const int sweeps = warp_jacobi_any(A, V, rot, D, lane);
  if ((__any_sync(0xffffffffu, bad) || sweeps > kJacobiMaxSweeps) && lane == 0 && status) atomicMax(status, m + 1);
  for (int x = lane; x < D; x += 32) {
    const int r = eig_rank(A, D, x);
    w[(size_t)m * D + r] = A[x * D + x].x;
    for (int d = 0; d < D; ++d) v[(size_t)m * D * D + d * D + r] = V[d * D + x];
  }
}

// out[j] = sum_i in[i * C + j] (fixed order)
__global__ void colsum_kernel(const double* __restrict__ in, double* __restrict__ out, int rows, int C) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= C) return;
  double s = 0.0;
  for (int i = 0; i < rows; ++i) s += in[(size_t)i * C + j];
  out[j] = s;
}

static int warps_for(size_t per_warp_bytes) {
  int w = (int)((size_t)(96 * 1024) / per_warp_bytes);
  if (w > 4) w = 4;
  if (w < 1) w = 1;
  return w;
}

}  // namespace pbb

using namespace pbb;

extern "C" {

int pbb_heig_batched(const void* a, int n, int D, double* w, void* v, int* status, void* stream) {
  PBB_CHECK_ARG(a != nullptr, 1, "a is null");
  PBB_CHECK_ARG(n > 0, 2, "n must be positive");
  PBB_CHECK_ARG(D > 0 && D <= 64, 3, "need 0 < D <= 64");
  PBB_CHECK_ARG(w != nullptr, 4, "w is null");
  PBB_CHECK_ARG(v != nullptr, 5, "v is null");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t per = (jacobi_smem_bytes(D) + 15) & ~(size_t)15;
  const int warps = warps_for(per);
  PBB_CUDA(cudaFuncSetAttribute(heig_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  LaunchScope ls("heig_batched_kernel", st);
  heig_batched_kernel<<<(n + warps - 1) / warps, 32 * warps, per * warps, st>>>(
      reinterpret_cast<const double2*>(a), n, D, w, reinterpret_cast<double2*>(v), status, warps);
  PBB_CUDA(cudaGetLastError());
  return 0;
}

int pbb_gev_batched(const void* a, const void* b, int n, int D, void* w, int* status, void* stream) {
  PBB_CHECK_ARG(a != nullptr, 1, "target PSD is null");
  PBB_CHECK_ARG(b != nullptr, 2, "noise PSD is null");
  PBB_CHECK_ARG(n > 0, 3, "n must be positive");
  PBB_CHECK_ARG(D > 0 && D <= 64, 4, "need 0 < D <= 64");
  PBB_CHECK_ARG(w != nullptr, 5, "w is null");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t per = ((size_t)3 * D * D * sizeof(double2) + (size_t)((D + 1) / 2) * 6 * sizeof(double) + 15) &
                     ~(size_t)15;
  const int warps = warps_for(per);
  PBB_CUDA(cudaFuncSetAttribute(gev_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  LaunchScope ls("gev_kernel", st);
  gev_kernel<<<(n + warps - 1) / warps, 32 * warps, per * warps, st>>>(
      reinterpret_cast<const double2*>(a), reinterpret_cast<const double2*>(b), n, D,
      reinterpret_cast<double2*>(w), status, warps);
  PBB_CUDA(cudaGetLastError());
  return 0;
}

int pbb_solve_batched(const void* a, const void* b, int n, int D, int R, int hermitize, void* x, int* status,
                      void* stream) {
  PBB_CHECK_ARG(a != nullptr, 1, "a is null");
  PBB_CHECK_ARG(b != nullptr, 2, "b is null");
  PBB_CHECK_ARG(n > 0, 3, "n must be positive");
  PBB_CHECK_ARG(D > 0 && D <= 64, 4, "need 0 < D <= 64");
  PBB_CHECK_ARG(R > 0 && R <= 64, 5, "need 0 < R <= 64");
  PBB_CHECK_ARG(x != nullptr, 7, "x is null");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t per = solve_smem_per_warp(D, R);
  const int warps = warps_for(per);
  PBB_CUDA(cudaFuncSetAttribute(solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  LaunchScope ls("solve_kernel", st);
  solve_kernel<<<(n + warps - 1) / warps, 32 * warps, per * warps, st>>>(
      reinterpret_cast<const double2*>(a), reinterpret_cast<const double2*>(b), n, D, R, hermitize,
      reinterpret_cast<double2*>(x), status, warps);
  PBB_CUDA(cudaGetLastError());
  return 0;
}

int pbb_mvdr(const void* atf, const void* noise_psd, int n, int D, void* w, void* scratch, int* status,
             void* stream) {
  PBB_CHECK_ARG(atf != nullptr, 1, "atf is null");
  PBB_CHECK_ARG(noise_psd != nullptr, 2, "noise PSD is null");
  PBB_CHECK_ARG(w != nullptr, 5, "w is null");
  PBB_CHECK_ARG(scratch != nullptr, 6, "scratch (n * D complex128) is null");
  int r = pbb_solve_batched(noise_psd, atf, n, D, 1, 1, scratch, status, stream);
  if (r) return r;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  LaunchScope ls("mvdr_scale_kernel", st);
  mvdr_scale_kernel<<<(n + 127) / 128, 128, 0, st>>>(reinterpret_cast<const double2*>(atf),
                                                      reinterpret_cast<const double2*>(scratch), n, D,
                                                      reinterpret_cast<double2*>(w));
  PBB_CUDA(cudaGetLastError());
  return 0;
}

int pbb_souden(const void* phi, const void* target_psd, const void* noise_psd, int n, int D, double eps, void* mat,
               void* num, void* den, void* num_sum, void* den_sum, void* stream) {
  PBB_CHECK_ARG(phi && target_psd && noise_psd, 1, "input is null");
  PBB_CHECK_ARG(n > 0 && D > 0 && D <= 64, 4, "bad shape");
  PBB_CHECK_ARG(mat && num && den && num_sum && den_sum, 7, "output is null");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  {
    LaunchScope ls("souden_kernel", st);
    souden_kernel<<<n, 64, 0, st>>>(reinterpret_cast<const double2*>(phi),
                                    reinterpret_cast<const double2*>(target_psd),
                                    reinterpret_cast<const double2*>(noise_psd), n, D, eps,
                                    reinterpret_cast<double2*>(mat), reinterpret_cast<double2*>(num),
                                    reinterpret_cast<double2*>(den));
    PBB_CUDA(cudaGetLastError());
  }
  LaunchScope ls("colsum_kernel", st);
  colsum_kernel<<<1, 128, 0, st>>>(reinterpret_cast<const double*>(num), reinterpret_cast<double*>(num_sum), n, 2 * D);
  colsum_kernel<<<1, 128, 0, st>>>(reinterpret_cast<const double*>(den), reinterpret_cast<double*>(den_sum), n, 2 * D);
  PBB_CUDA(cudaGetLastError());
  return 0;
}

int pbb_blind_analytic_normalization(const void* vector, const void* noise_psd, int n, int D, void* out,
                                     void* stream) {
  PBB_CHECK_ARG(vector && noise_psd, 1, "input is null");
  PBB_CHECK_ARG(n > 0 && D > 0, 3, "bad shape");
  PBB_CHECK_ARG(out != nullptr, 5, "out is null");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  LaunchScope ls("ban_kernel", st);
  ban_kernel<<<(n + 127) / 128, 128, 0, st>>>(reinterpret_cast<const double2*>(vector),
                                              reinterpret_cast<const double2*>(noise_psd), n, D,
                                              reinterpret_cast<double2*>(out));
  PBB_CUDA(cudaGetLastError());
  return 0;
}

static int apply_bf_launch(const void* vector, const void* mix, int dtype, int B, int F, int D, int T, void* out,
                           void* stream) {
  PBB_CHECK_ARG(vector && mix, 1, "input is null");
  PBB_CHECK_ARG(dtype == PBB_C64 || dtype == PBB_C128, 3, "bad dtype");
  PBB_CHECK_ARG(B > 0 && B <= 65535 && F > 0 && F <= 65535 && D > 0 && T > 0, 4, "bad shape");
  PBB_CHECK_ARG(out != nullptr, 7, "out is null");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid((T + 255) / 256, F, B);
  LaunchScope ls("apply_bf_kernel", st);
  if (dtype == PBB_C128)
    apply_bf_kernel<double2><<<grid, 256, 0, st>>>(reinterpret_cast<const double2*>(vector),
                                                    reinterpret_cast<const double2*>(mix), F, D, T,
                                                    reinterpret_cast<double2*>(out));
  else
    apply_bf_kernel<float2><<<grid, 256, 0, st>>>(reinterpret_cast<const double2*>(vector),
                                                   reinterpret_cast<const float2*>(mix), F, D, T,
                                                   reinterpret_cast<double2*>(out));
  PBB_CUDA(cudaGetLastError());
  return 0;
}

int pbb_apply_beamforming_vector(const void* vector, const void* mix, int dtype, int F, int D, int T, void* out,
                                 void* stream) {
  return apply_bf_launch(vector, mix, dtype, 1, F, D, T, out, stream);
}

int pbb_apply_beamforming_vector_shared(const void* vector, const void* mix, int dtype, int B, int F, int D, int T,
                                        void* out, void* stream) {
  return apply_bf_launch(vector, mix, dtype, B, F, D, T, out, stream);
}

int pbb_rank_one_estimate(const void* vector, const void* covariance, int n, int D, void* out, void* stream) {
  PBB_CHECK_ARG(vector && covariance, 1, "input is null");
  PBB_CHECK_ARG(n > 0 && D > 0 && D <= 64, 3, "bad shape");
  PBB_CHECK_ARG(out != nullptr, 5, "out is null");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  LaunchScope ls("rank_one_kernel", st);
  rank_one_kernel<<<n, 64, 0, st>>>(reinterpret_cast<const double2*>(vector),
                                    reinterpret_cast<const double2*>(covariance), n, D,
                                    reinterpret_cast<double2*>(out));
  PBB_CUDA(cudaGetLastError());
  return 0;
}

int pbb_matvec_batched(const void* matrix, const void* vector, int n, int D, void* out, void* stream) {
  PBB_CHECK_ARG(matrix && vector, 1, "input is null");
  PBB_CHECK_ARG(n > 0 && D > 0 && D <= 64, 3, "bad shape");
  PBB_CHECK_ARG(out != nullptr, 5, "out is null");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  LaunchScope ls("matvec_kernel", st);
  matvec_kernel<<<(n + 127) / 128, 128, 0, st>>>(reinterpret_cast<const double2*>(matrix),
                                                 reinterpret_cast<const double2*>(vector), n, D,
                                                 reinterpret_cast<double2*>(out));
  PBB_CUDA(cudaGetLastError());
  return 0;
}

size_t pbb_psd_workspace_bytes(int F, int T, int D, int K) {
  if (F <= 0 || T <= 0 || D <= 0 || K <= 0) return 0;
  return (size_t)F * ((T + 31) / 32) * K * ((size_t)D * D + 1) * sizeof(double) + 256;
}

int pbb_power_spectral_density(const void* observation, int dtype, int F, int D, int T, const double* mask, int K,
                               int normalize, void* psd, void* workspace, size_t workspace_bytes, void* stream) {
  PBB_CHECK_ARG(observation != nullptr, 1, "observation is null");
  PBB_CHECK_ARG(dtype == PBB_C64 || dtype == PBB_C128, 2, "bad dtype");
  PBB_CHECK_ARG(F > 0 && D > 0 && D < 35 && T > 0, 3, "bad shape");
  PBB_CHECK_ARG(K > 0 && K < kMaxK, 7, "need 0 < K < 20");
  PBB_CHECK_ARG(psd != nullptr, 9, "psd is null");
  PBB_CHECK_ARG(workspace != nullptr && workspace_bytes >= pbb_psd_workspace_bytes(F, T, D, K), 10,
                "workspace too small (pbb_psd_workspace_bytes)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  EmArgs a;
  memset(&a, 0, sizeof(a));
  a.z = observation; a.zs = T; a.F = F; a.T = T; a.D = D; a.K = K;
  a.mode = kModeM; a.aff_in = mask; a.q_in = nullptr;
  a.part = reinterpret_cast<double*>(workspace);
  const int nch = launch_em(a, dtype, 0, st);
  if (nch <= 0) return nch ? nch : 1;
  const int scale = mask == nullptr ? 2 : (normalize ? 1 : 0);
  LaunchScope ls("psd_finalize_kernel", st);
  psd_finalize_kernel<<<dim3(F, K), 64, 0, st>>>(a.part, nch, F, K, D, T, scale, reinterpret_cast<double2*>(psd));
  PBB_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
