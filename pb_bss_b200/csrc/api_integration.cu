// C-ABI entry points for the integrated spatial + spectral mixture model (pb_bss/distribution/gcacgmm.py: the cACG
// of the multi-channel observation combined with a Gaussian over per-(bin, frame) embedding vectors) -- see
// include/pbb.h.  The spatial part reuses the cACGMM kernels (quadratic form, M-step, eigendecomposition); this
// unit adds the small HBM-streaming pieces around them:
//   pbb_cacg_log_pdf              -D log q - sum log lambda                      (cacg.py:198-201)
//   pbb_gaussian_log_pdf          diagonal / spherical Gaussian over (F, T, E)   (gaussian.py:57-135)
//   pbb_gaussian_fit              weighted mean + variance, two passes           (gaussian.py:155-193)
//   pbb_log_pdf_to_affiliation    softmax * weight, clip (mixture_model_utils.py:7-55), optionally after the
//                                 per-bin search over the K! pairings of spatial and spectral classes (:58-130)
//   pbb_class_weight              L1-normalised sums of the masked affiliations  (gcacgmm.py:283-291)
// All reductions run in a fixed order (bit-reproducible).
#include <cstring>

#include "common.cuh"
#include "em_args.cuh"
#include "prof.cuh"

namespace pbb {

constexpr int kMaxKInt = kMaxK;
constexpr int kIntMaxK = 6;    // K! permutations are enumerated per bin
constexpr int kIntMaxE = 64;   // embedding dimension held in registers / shared memory

__device__ inline double block_sum_256(double v, double* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double s = 0.0;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
  return s;
}

__global__ void cacg_log_pdf_kernel(const double* __restrict__ q, const double* __restrict__ eigenvalues, int F, int K,
                                    int T, int D, double* __restrict__ out) {
  const int fk = blockIdx.y;
  double ld = 0.0;
  for (int d = 0; d < D; ++d) ld += log(eigenvalues[(size_t)fk * D + d]);
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const double qq = fmax(fabs(q[(size_t)fk * T + t]), kTiny);
  out[(size_t)fk * T + t] = -(double)D * log(qq) - ld;
}

// out[f][k][t] = -E/2 log(2 pi) + log_det[k] - 1/2 sum_e (pc[k][e] (x[f][t][e] - mean[k][e]))^2           (spherical)
// diagonal != 0: the reference's DiagonalGaussian.log_pdf contracts its (K, E) precision_cholesky with the einsum
// '...dD,...nD->...nd' (gaussian.py:79-87), i.e. white[k][n][d] = sum_e pc[d][e] (x[n][e] - mean[k][e]) with d running
// over the CLASSES, and log_pdf = ... - 1/2 sum_d white^2.  A drop-in has to return what the reference returns, so this
// branch evaluates exactly that expression.
__global__ void gaussian_log_pdf_kernel(const double* __restrict__ x, const double* __restrict__ mean,
                                        const double* __restrict__ pc, const double* __restrict__ log_det, int F, int T,
                                        int E, int K, int diagonal, double* __restrict__ out) {
  extern __shared__ double sm[];  // mean [K][E], pc [K][E]
  for (int i = threadIdx.x; i < 2 * K * E; i += blockDim.x) sm[i] = i < K * E ? mean[i] : pc[i - K * E];
  __syncthreads();
  const int f = blockIdx.y, t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const double* __restrict__ xr = x + ((size_t)f * T + t) * E;
  double xe[kIntMaxE];
  for (int e = 0; e < E; ++e) xe[e] = xr[e];
  const double c0 = -0.5 * (double)E * log(2.0 * 3.14159265358979323846);
  for (int k = 0; k < K; ++k) {
    double s = 0.0;
    if (diagonal == 2) {
      // von Mises-Fisher (von_mises_fisher.py:66-81): concentration * <mean, x / max(||x||, tiny)> - log_norm;
      // pc[k][0] = concentration, log_det[k] = log_norm
      double dot = 0.0, n2 = 0.0;
      for (int e = 0; e < E; ++e) { dot += sm[k * E + e] * xe[e]; n2 += xe[e] * xe[e]; }
      out[((size_t)f * K + k) * T + t] = sm[K * E + k * E] * (dot / fmax(sqrt(n2), kTiny)) - log_det[k];
      continue;
    }
    if (diagonal) {
      for (int d = 0; d < K; ++d) {
        double w = 0.0;
        for (int e = 0; e < E; ++e) w += sm[K * E + d * E + e] * (xe[e] - sm[k * E + e]);
        s += w * w;
      }
    } else {
      for (int e = 0; e < E; ++e) {
        const double w = sm[K * E + k * E + e] * (xe[e] - sm[k * E + e]);
        s += w * w;
      }
    }
    out[((size_t)f * K + k) * T + t] = c0 + log_det[k] - 0.5 * s;
  }
}

// pass 0: partial[f][k][0..E) = sum_t w x, [E] = sum_t w ; pass 1: partial[f][k][0..E) = sum_t w (x - mean)^2
__global__ void gaussian_fit_partial_kernel(const double* __restrict__ x, const double* __restrict__ w,
                                            const double* __restrict__ mean, int F, int T, int E, int K, int pass,
                                            double* __restrict__ partial) {
  __shared__ double red[8];
  const int f = blockIdx.x, k = blockIdx.y;
  const double* __restrict__ wr = w + ((size_t)f * K + k) * T;
  for (int e = 0; e <= E; ++e) {
    if (pass == 1 && e == E) break;
    double s = 0.0;
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
      const double ww = wr[t];
      if (e == E) s += ww;
      else {
        const double xv = x[((size_t)f * T + t) * E + e];
        if (pass == 0) s += ww * xv;
        else { const double d = xv - mean[k * E + e]; s += ww * d * d; }
      }
    }
    s = block_sum_256(s, red);
    if (threadIdx.x == 0) partial[((size_t)f * K + k) * (E + 1) + e] = s;
  }
}
// mean[k][e] = sum_f partial / max(denominator, tiny); denominator kept in denom[k]
__global__ void gaussian_fit_mean_kernel(const double* __restrict__ partial, int F, int E, int K,
                                         double* __restrict__ mean, double* __restrict__ denom) {
  const int k = blockIdx.x, e = threadIdx.x;
  if (e > E) return;
  double s = 0.0;
  for (int f = 0; f < F; ++f) s += partial[((size_t)f * K + k) * (E + 1) + e];
  __shared__ double den;
  if (e == E) { den = fmax(s, kTiny); denom[k] = den; }
  __syncthreads();
  if (e < E) mean[k * E + e] = s / den;
}
// covariance: diagonal (K, E) or spherical (K)
__global__ void gaussian_fit_cov_kernel(const double* __restrict__ partial, const double* __restrict__ denom, int F, int E,
                                        int K, int spherical, double* __restrict__ cov) {
  const int k = blockIdx.x, e = threadIdx.x;
  __shared__ double v[kIntMaxE];
  if (e < E) {
    double s = 0.0;
    for (int f = 0; f < F; ++f) s += partial[((size_t)f * K + k) * (E + 1) + e];
    v[e] = s;
    if (!spherical) cov[k * E + e] = s / denom[k];
  }
  __syncthreads();
  if (spherical && e == 0) {
    double s = 0.0;
    for (int i = 0; i < E; ++i) s += v[i];
    cov[k] = s / (denom[k] * (double)E);
  }
}

__device__ __forceinline__ double weight_of(const double* __restrict__ w, int mode, int f, int k, int t, int K, int T) {
  switch (mode) {
    case PBB_WEIGHT_CONST: return 1.0 / K;
    case PBB_WEIGHT_TIED_TIME: return w[(size_t)k * T + t];
    case PBB_WEIGHT_TIED: return w[k];
    default: return w[(size_t)f * K + k];
  }
}

// One CTA per bin.  lp[k][t] = sa * a[f][perm(k)][t] + sb * b[f][k][t]; perm = identity, or (inline_pa) the first of
// itertools.permutations(range(K)) that maximises sum_{k,t} softmax_k(lp) * lp (mixture_model_utils.py:93-115).
__global__ void __launch_bounds__(256) log_pdf_to_affiliation_kernel(
    const double* __restrict__ a, const double* __restrict__ b, double sa, double sb, const double* __restrict__ weight,
    int weight_mode, const uint8_t* __restrict__ activity, double eps, int inline_pa, int F, int K, int T,
    double* __restrict__ out, int* __restrict__ chosen) {
  __shared__ double red[8];
  __shared__ int best_perm[kIntMaxK];
  const int f = blockIdx.x, tid = threadIdx.x;
  const double* __restrict__ af = a + (size_t)f * K * T;
  const double* __restrict__ bf = b ? b + (size_t)f * K * T : nullptr;
  int perm[kIntMaxK];
  for (int k = 0; k < K; ++k) perm[k] = k;
  if (inline_pa && bf != nullptr) {
    int cand[kIntMaxK];
    for (int k = 0; k < K; ++k) cand[k] = k;
    double best = -INFINITY;
    bool have = false;
    while (true) {
      double aux = 0.0;
      for (int t = tid; t < T; t += blockDim.x) {
        double lp[kIntMaxK], m = -INFINITY;
        for (int k = 0; k < K; ++k) { lp[k] = sa * af[(size_t)cand[k] * T + t] + sb * bf[(size_t)k * T + t]; m = fmax(m, lp[k]); }
        double den = 0.0, ex[kIntMaxK];
        for (int k = 0; k < K; ++k) { ex[k] = exp(lp[k] - m); den += ex[k]; }
        den = fmax(den, kTiny);
        for (int k = 0; k < K; ++k) aux += (ex[k] / den) * lp[k];
      }
      aux = block_sum_256(aux, red);
      if (!have || aux > best) {
        best = aux; have = true;
        for (int k = 0; k < K; ++k) perm[k] = cand[k];
      }
      int i = K - 2;  // next lexicographic permutation
      while (i >= 0 && cand[i] > cand[i + 1]) --i;
      if (i < 0) break;
      int j = K - 1;
      while (cand[j] < cand[i]) --j;
      { const int tmp = cand[i]; cand[i] = cand[j]; cand[j] = tmp; }
      for (int x = i + 1, y = K - 1; x < y; ++x, --y) { const int tmp = cand[x]; cand[x] = cand[y]; cand[y] = tmp; }
    }
    if (tid == 0 && chosen != nullptr)
      for (int k = 0; k < K; ++k) chosen[(size_t)f * K + k] = perm[k];
  }
  (void)best_perm;
  for (int t = tid; t < T; t += blockDim.x) {
    double lp[kIntMaxK], m = -INFINITY;
    for (int k = 0; k < K; ++k) {
      lp[k] = sa * af[(size_t)perm[k] * T + t] + (bf ? sb * bf[(size_t)k * T + t] : 0.0);
      m = fmax(m, lp[k]);
    }
    double g[kIntMaxK], den = 0.0;
    for (int k = 0; k < K; ++k) {
      g[k] = exp(lp[k] - m) * weight_of(weight, weight_mode, f, k, t, K, T);
      if (activity != nullptr && !activity[((size_t)f * K + k) * T + t]) g[k] = 0.0;
      den += g[k];
    }
    den = fmax(den, kTiny);
    for (int k = 0; k < K; ++k) {
      double v = g[k] / den;
      if (eps != 0.0) v = fmin(fmax(v, eps), 1.0 - eps);
      out[((size_t)f * K + k) * T + t] = v;
    }
  }
}

// per-bin class weights: w[f][k] = sum_t m[f][k][t] / sum_k sum_t m[f][k][t]   (gcacgmm.py:286-291, axis (-1,))
__global__ void class_weight_kernel(const double* __restrict__ m, int F, int K, int T, double* __restrict__ w) {
  __shared__ double red[8];
  __shared__ double s[kMaxKInt];
  const int f = blockIdx.x;
  for (int k = 0; k < K; ++k) {
    double v = 0.0;
    for (int t = threadIdx.x; t < T; t += blockDim.x) v += m[((size_t)f * K + k) * T + t];
    v = block_sum_256(v, red);
    if (threadIdx.x == 0) s[k] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int k = 0; k < K; ++k) tot += s[k];
    for (int k = 0; k < K; ++k) w[(size_t)f * K + k] = s[k] / tot;
  }
}

}  // namespace pbb

using namespace pbb;

extern "C" {

int pbb_cacg_log_pdf(const double* quadratic, const double* eigenvalues, int F, int K, int T, int D, double* log_pdf,
                     void* stream) {
  PBB_CHECK_ARG(quadratic && eigenvalues, 1, "input is null");
  PBB_CHECK_ARG(F > 0 && K > 0 && T > 0 && D > 0, 3, "bad shape");
  PBB_CHECK_ARG(log_pdf != nullptr, 7, "output is null");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  LaunchScope ls("cacg_log_pdf_kernel", st);
  cacg_log_pdf_kernel<<<dim3((T + 127) / 128, F * K), 128, 0, st>>>(quadratic, eigenvalues, F, K, T, D, log_pdf);
  PBB_CUDA(cudaGetLastError());
  return 0;
}

int pbb_gaussian_log_pdf(const double* embedding, const double* mean, const double* precision_cholesky,
                         const double* log_det, int F, int T, int E, int K, int diagonal, double* log_pdf,
                         void* stream) {
  PBB_CHECK_ARG(embedding && mean && precision_cholesky && log_det, 1, "input is null");
  PBB_CHECK_ARG(F > 0 && T > 0, 5, "bad shape");
  PBB_CHECK_ARG(E > 0 && E <= kIntMaxE, 7, "need 0 < E <= 64");
  PBB_CHECK_ARG(K > 0 && K < kMaxK, 8, "bad K");
  PBB_CHECK_ARG(log_pdf != nullptr, 10, "output is null");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  LaunchScope ls("gaussian_log_pdf_kernel", st);
  gaussian_log_pdf_kernel<<<dim3((T + 127) / 128, F), 128, (size_t)2 * K * E * sizeof(double), st>>>(
      embedding, mean, precision_cholesky, log_det, F, T, E, K, diagonal, log_pdf);
  PBB_CUDA(cudaGetLastError());
  return 0;
}

size_t pbb_gaussian_fit_scratch_doubles(int F, int E, int K) { return (size_t)F * K * (E + 1) + K; }

int pbb_gaussian_fit(const double* embedding, const double* weight, int F, int T, int E, int K, int spherical,
                     double* mean, double* covariance, double* scratch, void* stream) {
  PBB_CHECK_ARG(embedding && weight, 1, "input is null");
  PBB_CHECK_ARG(F > 0 && T > 0, 3, "bad shape");
  PBB_CHECK_ARG(E > 0 && E <= kIntMaxE, 5, "need 0 < E <= 64");
  PBB_CHECK_ARG(K > 0 && K < kMaxK, 6, "bad K");
  PBB_CHECK_ARG(mean && covariance && scratch, 8, "output / scratch is null (pbb_gaussian_fit_scratch_doubles)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  double* partial = scratch;
  double* denom = scratch + (size_t)F * K * (E + 1);
  LaunchScope ls("gaussian_fit_kernels", st);
  gaussian_fit_partial_kernel<<<dim3(F, K), 256, 0, st>>>(embedding, weight, mean, F, T, E, K, 0, partial);
  gaussian_fit_mean_kernel<<<K, kIntMaxE + 1, 0, st>>>(partial, F, E, K, mean, denom);
  gaussian_fit_partial_kernel<<<dim3(F, K), 256, 0, st>>>(embedding, weight, mean, F, T, E, K, 1, partial);
  gaussian_fit_cov_kernel<<<K, kIntMaxE, 0, st>>>(partial, denom, F, E, K, spherical, covariance);
  PBB_CUDA(cudaGetLastError());
  return 0;
}

int pbb_log_pdf_to_affiliation(const double* log_pdf_a, const double* log_pdf_b, double scale_a, double scale_b,
                               const double* weight, int weight_mode, const uint8_t* activity, double affiliation_eps,
                               int inline_pa, int F, int K, int T, double* affiliation, int* permutation,
                               void* stream) {
  PBB_CHECK_ARG(log_pdf_a != nullptr, 1, "log pdf is null");
  PBB_CHECK_ARG(weight != nullptr || weight_mode == PBB_WEIGHT_CONST, 5, "weight is null");
  PBB_CHECK_ARG(weight_mode >= 0 && weight_mode <= PBB_WEIGHT_TIED, 6, "bad weight_mode");
  PBB_CHECK_ARG(F > 0 && T > 0, 10, "bad shape");
  PBB_CHECK_ARG(K > 0 && K <= kIntMaxK, 11, "need 0 < K <= 6 (K! pairings per bin)");
  PBB_CHECK_ARG(!inline_pa || log_pdf_b != nullptr, 9, "the inline alignment pairs TWO log pdfs");
  PBB_CHECK_ARG(affiliation != nullptr, 13, "output is null");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  LaunchScope ls("log_pdf_to_affiliation_kernel", st);
  log_pdf_to_affiliation_kernel<<<F, 256, 0, st>>>(log_pdf_a, log_pdf_b, scale_a, scale_b, weight, weight_mode, activity,
                                                   affiliation_eps, inline_pa, F, K, T, affiliation, permutation);
  PBB_CUDA(cudaGetLastError());
  return 0;
}

int pbb_class_weight(const double* masked_affiliation, int F, int K, int T, double* weight, void* stream) {
  PBB_CHECK_ARG(masked_affiliation != nullptr, 1, "input is null");
  PBB_CHECK_ARG(F > 0 && K > 0 && K < kMaxK && T > 0, 2, "bad shape");
  PBB_CHECK_ARG(weight != nullptr, 5, "output is null");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  LaunchScope ls("class_weight_kernel", st);
  class_weight_kernel<<<F, 256, 0, st>>>(masked_affiliation, F, K, T, weight);
  PBB_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
