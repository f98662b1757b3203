// Frequency permutation alignment on the device (pb_bss/permutation_alignment.py):
// DHTVPermutationAlignment.calculate_mapping (:295-355) with the 'cos' similarity
// (_ScoreMatrix.multiply, :404-410) and the greedy assignment (:525-553), and
// apply_mapping (:54-104).  See include/pbb.h.
#include "common.cuh"
#include "prof.cuh"

namespace pbb {

constexpr int kDhtvMaxK = 9;       // the reference asserts K < 10 (permutation_alignment.py:200)
constexpr int kDhtvThreads = 1024;

__device__ inline double block_sum(double v, double* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double s = 0.0;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
  return s;
}

// features = mask / max(||mask||_T, tiny) per (k, f)   (:309-310, _parameterized_vector_norm :358-377)
__global__ void dhtv_normalize_kernel(const double* __restrict__ mask, double* __restrict__ feat, int KF, int T) {
  const int row = blockIdx.x;
  if (row >= KF) return;
  __shared__ double red[32];
  const double* __restrict__ m = mask + (size_t)row * T;
  double s = 0.0;
  for (int t = threadIdx.x; t < T; t += blockDim.x) s += m[t] * m[t];
  const double n = sqrt(block_sum(s, red));
  const double d = fmax(n, kTiny);
  for (int t = threadIdx.x; t < T; t += blockDim.x) feat[(size_t)row * T + t] = m[t] / d;
}

// The whole alignment plan in one single-CTA kernel: the work per iteration is a
// few hundred kflop, the algorithm is a chain of ~60 dependent iterations, so
// latency (block barriers, L2 round trips) is what matters, not parallel width.
__global__ void __launch_bounds__(kDhtvThreads) dhtv_kernel(double* __restrict__ feat, double* __restrict__ cent,
                                                            const int* __restrict__ plan, int nplan, int K, int F,
                                                            int T, long long* __restrict__ mapping) {
  __shared__ double red[32];
  __shared__ double cnorm[kDhtvMaxK];
  __shared__ int changed;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  for (int i = tid; i < K * F; i += blockDim.x) mapping[i] = i / F;  // mapping[k][f] = k
  __syncthreads();
  for (int p = 0; p < nplan; ++p) {
    const int iters = plan[3 * p], start = plan[3 * p + 1], end = plan[3 * p + 2];
    for (int it = 0; it < iters; ++it) {
      // (a) centroid over the segment's bins, L2-normalised over time (:334-340)
      for (int i = tid; i < K * T; i += blockDim.x) {
        const int k = i / T, t = i - k * T;
        double s = 0.0;
        for (int f = start; f < end; ++f) s += feat[((size_t)k * F + f) * T + t];
        cent[i] = s / (double)(end - start);
      }
      if (tid == 0) changed = 0;
      __syncthreads();
      for (int k = 0; k < K; ++k) {
        double s = 0.0;
        for (int t = tid; t < T; t += blockDim.x) { const double c = cent[k * T + t]; s += c * c; }
        const double n = sqrt(block_sum(s, red));
        if (tid == 0) cnorm[k] = fmax(n, kTiny);
      }
      __syncthreads();
      for (int i = tid; i < K * T; i += blockDim.x) cent[i] = cent[i] / cnorm[i / T];
      __syncthreads();
      // (b) one warp per bin: K x K scores, greedy assignment, permute (:342-350)
      for (int f = start + warp; f < end; f += nwarp) {
        double score[kDhtvMaxK * kDhtvMaxK];
        for (int kr = 0; kr < K; ++kr)
          for (int km = 0; km < K; ++km) {
            double s = 0.0;
            for (int t = lane; t < T; t += 32) s += feat[((size_t)km * F + f) * T + t] * cent[kr * T + t];
            score[kr * K + km] = warp_sum(s);  // identical in every lane
          }
        int perm[kDhtvMaxK];
        bool ident = true;
        for (int r = 0; r < K; ++r) {
          // first maximum of the row-major flattened matrix (np.argmax), then blank its row and column
          int bi = 0, bj = 0;
          double best = -INFINITY;
          bool found = false;
          for (int i = 0; i < K; ++i)
            for (int j = 0; j < K; ++j) {
              const double v = score[i * K + j];
              if (!found || v > best) { best = v; bi = i; bj = j; found = true; }
            }
          for (int j = 0; j < K; ++j) score[bi * K + j] = -INFINITY;
          for (int i = 0; i < K; ++i) score[i * K + bj] = -INFINITY;
          perm[bi] = bj;
        }
        for (int k = 0; k < K; ++k) ident = ident && perm[k] == k;
        if (!ident) {
          for (int t = lane; t < T; t += 32) {
            double v[kDhtvMaxK];
            for (int k = 0; k < K; ++k) v[k] = feat[((size_t)k * F + f) * T + t];
            for (int k = 0; k < K; ++k) feat[((size_t)k * F + f) * T + t] = v[perm[k]];
          }
          if (lane == 0) {
            long long mv[kDhtvMaxK];
            for (int k = 0; k < K; ++k) mv[k] = mapping[(size_t)k * F + f];
            for (int k = 0; k < K; ++k) mapping[(size_t)k * F + f] = mv[perm[k]];
            changed = 1;
          }
        }
      }
      __syncthreads();
      const int ch = changed;
      __syncthreads();
      if (!ch) break;  // nothing_changed (:352-353)
    }
  }
}

__global__ void apply_mapping_kernel(const double* __restrict__ mask, const long long* __restrict__ mapping, int K,
                                     int F, int T, double* __restrict__ out) {
  const int kf = blockIdx.x;  // k * F + f
  const int f = kf % F;
  const long long src = mapping[kf];
  const double* __restrict__ s = mask + ((size_t)src * F + f) * T;
  double* __restrict__ o = out + (size_t)kf * T;
  for (int t = threadIdx.x; t < T; t += blockDim.x) o[t] = s[t];
}

}  // namespace pbb

using namespace pbb;

extern "C" {

int pbb_dhtv_mapping(const double* mask, int K, int F, int T, const int* plan, int nplan, double* features,
                     double* centroid, long long* mapping, void* stream) {
  PBB_CHECK_ARG(mask != nullptr, 1, "mask is null");
  PBB_CHECK_ARG(K > 0 && K <= kDhtvMaxK, 2, "need 0 < K < 10 (permutation_alignment.py:200)");
  PBB_CHECK_ARG(F > 0, 3, "F must be positive");
  PBB_CHECK_ARG(T > 0, 4, "T must be positive");
  PBB_CHECK_ARG(plan != nullptr && nplan > 0, 5, "alignment plan is empty");
  PBB_CHECK_ARG(features && centroid, 7, "scratch (features (K,F,T), centroid (K,T)) is null");
  PBB_CHECK_ARG(mapping != nullptr, 9, "mapping is null");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  {
    LaunchScope ls("dhtv_normalize_kernel", st);
    dhtv_normalize_kernel<<<K * F, 128, 0, st>>>(mask, features, K * F, T);
    PBB_CUDA(cudaGetLastError());
  }
  LaunchScope ls("dhtv_kernel", st);
  dhtv_kernel<<<1, kDhtvThreads, 0, st>>>(features, centroid, plan, nplan, K, F, T, mapping);
  PBB_CUDA(cudaGetLastError());
  return 0;
}

int pbb_apply_mapping(const double* mask, const long long* mapping, int K, int F, int T, double* out,
                      void* stream) {
  PBB_CHECK_ARG(mask != nullptr, 1, "mask is null");
  PBB_CHECK_ARG(mapping != nullptr, 2, "mapping is null");
  PBB_CHECK_ARG(K > 0 && K < 20 && F > 0 && T > 0, 3, "bad shape (K < 20, permutation_alignment.py:102)");
  PBB_CHECK_ARG(out != nullptr, 6, "out is null");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  LaunchScope ls("apply_mapping_kernel", st);
  apply_mapping_kernel<<<K * F, 128, 0, st>>>(mask, mapping, K, F, T, out);
  PBB_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
