// Persistent cACGMM EM kernel: all EM iterations of a fit in ONE launch.
//
// Work unit ("task") = one EM iteration of one frequency bin: E-step with the
// bin's current model fused with the M-step accumulation over all T frames,
// then the model update of that bin.  Tasks are handed out in iteration-major
// order by an atomic ticket counter; bins are independent, so the only
// dependency is task (bin, it) -> (bin, it + 1), tracked by a per-bin
// release/acquire flag in L2.  Because a task only ever waits for a LOWER
// ticket, and tickets are only held by running CTAs, the schedule cannot
// deadlock, needs no grid-wide barrier and balances itself to within one task
// (F = 513 bins do not divide 148 SMs; 51300 tasks do).
//
// Inside a task the CTA is D/2 warps; warp g owns slot group g (common.cuh) and
// lane = frame.  The observation rows of the bin stream from L2 into a 2-stage
// shared-memory ring with 1-D TMA bulk copies (cp.async.bulk + mbarrier
// complete_tx), prefetched one 128-frame chunk ahead -- across task boundaries,
// so the next bin's first chunk arrives while this bin's model is updated.
//
// Model update per (bin, class): the E-step only needs B^{-1} and log det B up
// to a common scale (log_pdf = -D log q - log det is invariant to B -> s B,
// and gamma/q rescales the next scatter matrix by the same s), so intermediate
// iterations invert the trace-normalised scatter matrix by Gauss-Jordan
// elimination (HPD: no pivoting) instead of an eigendecomposition.  That is
// exact unless the reference would floor an eigenvalue
// (complex_angular_central_gaussian.py:111-126); the bound
// lambda_min / lambda_max >= 1 / (tr(A) tr(A^{-1})) > floor proves it would
// not.  If the bound fails (cond > ~1e9) the warp falls back to the Jacobi
// eigensolver with the reference's normalise-and-floor semantics.  The last
// iteration always leaves the raw scatter sums for cacg_update_kernel, which
// produces the reference-exact eigenvectors / eigenvalues / weights.
#pragma once
#include "common.cuh"
#include "em_kernels.cuh"
#include "heig.cuh"

namespace pbb {

struct PersistArgs {
  const void* z;   // (F, D, zs) unit-norm observation, rows zero padded to zs
  int zs;          // row stride in frames, multiple of 32
  int F, T;
  int iterations;  // EM iterations in this launch
  int first_is_m;  // iteration 0 is an M-step from aff_in with q = 1 (cacgmm.py:206-228)
  int user_model;  // iteration 0 uses a user supplied model: log-domain softmax
  int softmax_fast;
  const double* aff_in;     // (F, K, T)
  const double* saliency;   // (F, T) or null      [FULL]
  const uint8_t* activity;  // (F, K, T) or null   [FULL]
  double aff_eps;
  double eigenvalue_floor;
  int covariance_norm;
  int weight_mode;
  double* coef;  // (F, K, NS)   model state, updated in place
  double* ld;    // (F, K)
  double* w;     // (F, K)
  double* ew;    // (F, K)
  double* part;  // (F, K, NS + 1) raw scatter sums of the last iteration
  int* flags;    // (F) number of model updates published for the bin
  int* ticket;   // (1)
  int* status;
};

// ---- PTX helpers --------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// reciprocal without the slow-path branches of 1.0 / x: hardware seed + 2 Newton
// steps; exact to ~1 ulp for normal, finite x (here 1e-300 < x < 1e300).
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
}

__device__ __forceinline__ double2 lds_cplx(const double2* p) { return *p; }
__device__ __forceinline__ double2 lds_cplx(const float2* p) {
  const float2 v = *p;
  return make_double2((double)v.x, (double)v.y);
}

constexpr int kStageFrames = 128;  // frames per ring stage (4 steps of 32)
constexpr int kStages = 2;

template <int D, int K, typename CT>
struct PersistSmem {
  static constexpr int NS = D * D;
  static constexpr int M = D / 2;
  CT zbuf[kStages][D][kStageFrames];
  double2 A[K][NS];     // scatter matrix / its inverse
  double2 V[K][NS];     // eigenvectors (Jacobi fallback only)
  double coef[K][NS];   // E-step form of the bin's model
  double xq[2][M][K][32];
  double S[K][NS + 1];  // scatter sums + sum of gamma
  double rot[K][((D + 1) / 2) * 6];
  double lam[K][D];
  double ld[K], w[K], ew[K];
  uint64_t full[kStages];
  int tab[NS];
  int tick[2];
};

// One EM step of 32 frames for slot group g.  FULL adds saliency / activity /
// the log-domain softmax.
template <int D, int K, typename CT, bool FULL>
struct PersistStep {
  static constexpr int NSG = group_shape(D).nsg, NLOC = group_shape(D).nloc, NS = D * D;
  static constexpr int NFULL = group_shape(D).nfull, HALF = group_shape(D).half;

  __device__ static __forceinline__ void psi_of(const double2 (&x)[NLOC], double (&psi)[NSG]) {
    psi[0] = x[0].x * x[0].x + x[0].y * x[0].y;
    psi[1] = x[1].x * x[1].x + x[1].y * x[1].y;
    psi[2] = x[0].x * x[1].x + x[0].y * x[1].y;
    psi[3] = x[0].x * x[1].y - x[0].y * x[1].x;
#pragma unroll
    for (int j = 0; j < NFULL; ++j) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const double2 u = x[q / 2], v = x[2 + 2 * j + (q % 2)];
        psi[4 + 8 * j + 2 * q] = u.x * v.x + u.y * v.y;
        psi[4 + 8 * j + 2 * q + 1] = u.x * v.y - u.y * v.x;
      }
    }
    if (HALF) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const double2 u = x[h], v = x[2 + 2 * NFULL + h];
        psi[4 + 8 * NFULL + 2 * h] = u.x * v.x + u.y * v.y;
        psi[4 + 8 * NFULL + 2 * h + 1] = u.x * v.y - u.y * v.x;
      }
    }
  }
};

template <int D, int K>
__device__ __forceinline__ void persist_softmax(const double (&q)[K], const double* __restrict__ ld,
                                                const double* __restrict__ w, const double* __restrict__ ew,
                                                bool fast, double eps, double (&gam)[K], double (&invq)[K]) {
  double a[K];
  if (fast) {
    double qmin = q[0];
#pragma unroll
    for (int k = 1; k < K; ++k) qmin = fmin(qmin, q[k]);
#pragma unroll
    for (int k = 0; k < K; ++k) {
      invq[k] = fast_rcp(q[k]);
      a[k] = ew[k] * ipow<D>(qmin * invq[k]);
    }
  } else {
    double lp[K];
    double m = -INFINITY;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      lp[k] = -(double)D * log(q[k]) - ld[k];
      m = fmax(m, lp[k]);
      invq[k] = 1.0 / q[k];
    }
#pragma unroll
    for (int k = 0; k < K; ++k) a[k] = exp(lp[k] - m) * w[k];
  }
  double den = a[0];
#pragma unroll
  for (int k = 1; k < K; ++k) den += a[k];
  const double inv = fast_rcp(fmax(den, kTiny));
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double g = a[k] * inv;
    if (eps != 0.0) g = fmin(fmax(g, eps), 1.0 - eps);
    gam[k] = g;
  }
}

// In-place inverse of a Hermitian positive definite D x D matrix in shared
// memory by Gauss-Jordan elimination without pivoting (one warp).  Returns
// log det; *ok is false if a pivot is not positive / finite.
template <int D>
__device__ __forceinline__ double warp_hpd_inverse(double2* __restrict__ A, int lane, bool* ok) {
  constexpr int NS = D * D;
  constexpr int PER = (NS + 31) / 32;
  double ldet = 0.0;
  bool good = true;
#pragma unroll 1
  for (int j = 0; j < D; ++j) {
    const double p = A[j * D + j].x;
    good = good && (p > 0.0) && isfinite(p);
    const double ip = 1.0 / p;
    ldet += log(p);
    double2 nv[PER];
#pragma unroll
    for (int r = 0; r < PER; ++r) {
      const int idx = lane + 32 * r;
      if (idx < NS) {
        const int i = idx / D, k = idx - i * D;
        const double2 aik = A[idx], aij = A[i * D + j], ajk = A[j * D + k];
        double2 v;
        if (i == j && k == j) v = make_double2(ip, 0.0);
        else if (i == j) v = make_double2(ajk.x * ip, ajk.y * ip);
        else if (k == j) v = make_double2(-aij.x * ip, -aij.y * ip);
        else {
          const double tr = (aij.x * ajk.x - aij.y * ajk.y) * ip, ti = (aij.x * ajk.y + aij.y * ajk.x) * ip;
          v = make_double2(aik.x - tr, aik.y - ti);
        }
        nv[r] = v;
      }
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < PER; ++r) {
      const int idx = lane + 32 * r;
      if (idx < NS) A[idx] = nv[r];
    }
    __syncwarp();
  }
  *ok = good;
  return ldet;
}

template <int D, int K, typename CT, bool FULL>
__global__ void __launch_bounds__(32 * (D / 2), (D == 8 ? 3 : (D == 6 ? 4 : 6)))
em_persistent_kernel(const PersistArgs a) {
  using SM = PersistSmem<D, K, CT>;
  using ST = PersistStep<D, K, CT, FULL>;
  constexpr int NS = D * D, M = D / 2, NSG = ST::NSG, NLOC = ST::NLOC;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SM& sm = *reinterpret_cast<SM*>(smem_raw);
  const int tid = threadIdx.x, g = tid >> 5, lane = tid & 31;
  const int F = a.F, T = a.T, zs = a.zs;
  const int total = a.iterations * F;
  const int nchunks = (zs + kStageFrames - 1) / kStageFrames;
  const CT* __restrict__ zbase = reinterpret_cast<const CT*>(a.z);

  for (int s = tid; s < NS; s += blockDim.x) sm.tab[s] = slot_pack(D, s);
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&sm.full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    sm.tick[0] = atomicAdd(a.ticket, 1);
  }
  __syncthreads();
  int cur = sm.tick[0];
  unsigned chunk_cnt = 0;  // chunks consumed so far by this CTA (ring position)

  auto issue_chunk = [&](int bin, int c, unsigned n) {  // thread 0 only
    const int st = n & 1u;
    const int t0 = c * kStageFrames;
    const int nf = min(kStageFrames, zs - t0);
    const uint32_t bytes = (uint32_t)nf * sizeof(CT);
    mbar_expect_tx(&sm.full[st], bytes * D);
#pragma unroll
    for (int d = 0; d < D; ++d)
      bulk_g2s(&sm.zbuf[st][d][0], zbase + ((size_t)bin * D + d) * zs + t0, bytes, &sm.full[st]);
  };
  if (tid == 0 && cur < total) issue_chunk(cur % F, 0, 0);

  // per-warp constants: channel offsets of the group's local channels
  int choff[NLOC];
#pragma unroll
  for (int l = 0; l < NLOC; ++l) choff[l] = group_channel(D, g, l) * kStageFrames;

  while (cur < total) {
    const int it = cur / F, bin = cur - it * F;
    const bool mstep_only = a.first_is_m && it == 0;
    const bool last_it = it == a.iterations - 1;
    if (tid == 0) {
      sm.tick[1] = atomicAdd(a.ticket, 1);  // the task after this one (prefetch target)
      if (!mstep_only) {
        while (ld_acquire_gpu(a.flags + bin) < it) __nanosleep(40);
      }
    }
    __syncthreads();
    const int nxt = sm.tick[1];
    if (!mstep_only) {
      const double* __restrict__ cf = a.coef + (size_t)bin * K * NS;
      for (int i = tid; i < K * NS; i += blockDim.x) (&sm.coef[0][0])[i] = __ldcg(cf + i);
      if (tid < K) {
        sm.ew[tid] = __ldcg(a.ew + (size_t)bin * K + tid);
        sm.ld[tid] = __ldcg(a.ld + (size_t)bin * K + tid);
        sm.w[tid] = __ldcg(a.w + (size_t)bin * K + tid);
      }
    }
    const bool fast = FULL ? (a.softmax_fast && !(a.user_model && it == 0)) : true;

    double acc[K * NSG];
#pragma unroll
    for (int i = 0; i < K * NSG; ++i) acc[i] = 0.0;
    double sg[K];
#pragma unroll
    for (int k = 0; k < K; ++k) sg[k] = 0.0;
    int buf = 0;

#pragma unroll 1
    for (int c = 0; c < nchunks; ++c) {
      __syncthreads();  // previous chunk fully consumed (and model staged): its stage may be refilled
      if (tid == 0) {
        if (c + 1 < nchunks) issue_chunk(bin, c + 1, chunk_cnt + 1);
        else if (nxt < total) issue_chunk(nxt % F, 0, chunk_cnt + 1);
      }
      const int st = chunk_cnt & 1u;
      mbar_wait(&sm.full[st], (chunk_cnt >> 1) & 1u);
      ++chunk_cnt;
      const CT* __restrict__ zb = &sm.zbuf[st][0][0];
      const int t_chunk = c * kStageFrames;
      const int nsteps = (min(kStageFrames, zs - t_chunk)) >> 5;
#pragma unroll 1
      for (int j = 0; j < nsteps; ++j) {
        const int t = t_chunk + j * 32 + lane;
        const bool valid = t < T;
        double2 x[NLOC];
#pragma unroll
        for (int l = 0; l < NLOC; ++l) x[l] = lds_cplx(zb + choff[l] + j * 32 + lane);
        double psi[NSG];
        ST::psi_of(x, psi);
        double gam[K], invq[K];
        if (!mstep_only) {
          const double* __restrict__ cg = &sm.coef[0][g * NSG];
#pragma unroll
          for (int k = 0; k < K; ++k) {
            double pq = 0.0;
#pragma unroll
            for (int i = 0; i < NSG; i += 2) {
              const double2 cc = *reinterpret_cast<const double2*>(cg + k * NS + i);
              pq = fma(cc.x, psi[i], pq);
              pq = fma(cc.y, psi[i + 1], pq);
            }
            sm.xq[buf][g][k][lane] = pq;
          }
          __syncthreads();
          double q[K];
#pragma unroll
          for (int k = 0; k < K; ++k) {
            double v = sm.xq[buf][0][k][lane];
#pragma unroll
            for (int gg = 1; gg < M; ++gg) v += sm.xq[buf][gg][k][lane];
            q[k] = fmax(fabs(v), 10.0 * kTiny);
          }
          buf ^= 1;
          if (FULL && a.activity != nullptr) {
            // masked classes get zero posterior mass (mixture_model_utils.py:39-41):
            // fold the mask into the class weights of this frame
            double ewm[K], wm[K];
#pragma unroll
            for (int k = 0; k < K; ++k) {
              const bool on = a.activity[((size_t)bin * K + k) * T + (valid ? t : 0)] != 0;
              ewm[k] = on ? sm.ew[k] : 0.0;
              wm[k] = on ? sm.w[k] : 0.0;
            }
            persist_softmax<D, K>(q, sm.ld, wm, ewm, fast, a.aff_eps, gam, invq);
          } else {
            persist_softmax<D, K>(q, sm.ld, sm.w, sm.ew, fast, a.aff_eps, gam, invq);
          }
        } else {
#pragma unroll
          for (int k = 0; k < K; ++k) {
            gam[k] = a.aff_in[((size_t)bin * K + k) * T + (valid ? t : 0)];
            invq[k] = 1.0;
          }
        }
        double sal = 1.0;
        if (FULL && a.saliency != nullptr) sal = a.saliency[(size_t)bin * T + (valid ? t : 0)];
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const double gs = valid ? (FULL ? gam[k] * sal : gam[k]) : 0.0;
          const double cw = gs * invq[k];
          sg[k] += gs;
#pragma unroll
          for (int i = 0; i < NSG; ++i) acc[k * NSG + i] = fma(cw, psi[i], acc[k * NSG + i]);
        }
      }
    }

    // ---- reduce the 32 frames of each warp; group g owns slots [g*NSG, (g+1)*NSG) ----
    warp_reduce_halving<K * NSG>(acc, lane);
    {
      int lo, hi;
      reduce_range<K * NSG>(lane, lo, hi);
#pragma unroll
      for (int j = 0; j < HalvingSizes<K * NSG>::n5; ++j) {
        const int idx = lo + j;
        if (idx < hi) {
          const int k = idx / NSG, i = idx - k * NSG;
          sm.S[k][g * NSG + i] = acc[j];
        }
      }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const double v = warp_sum(sg[k]);
      if (g == 0 && lane == 0) sm.S[k][NS] = v;
    }
    __syncthreads();

    if (last_it) {
      // leave the raw sums for cacg_update_kernel (reference-exact eigendecomposition)
      double* __restrict__ po = a.part + (size_t)bin * K * (NS + 1);
      for (int i = tid; i < K * (NS + 1); i += blockDim.x) po[i] = (&sm.S[0][0])[i];
    } else {
      // ---- model update, one warp per class -----------------------------------------
      for (int k = g; k < K; k += M) {
        double2* A = sm.A[k];
        double* Ad = reinterpret_cast<double*>(A);
        const double scale = (double)D / fmax(sm.S[k][NS], kTiny);
        bool bad = false;
        auto build = [&]() {
          for (int s = lane; s < NS; s += 32) {
            const int pk = sm.tab[s];
            const int d = pk & 255, e = (pk >> 8) & 255, kind = pk >> 16;
            const double v = sm.S[k][s] * scale;
            bad |= !isfinite(v);
            if (kind == 0) { Ad[2 * (d * D + d)] = v; Ad[2 * (d * D + d) + 1] = 0.0; }
            else if (kind == 1) { Ad[2 * (d * D + e)] = v; Ad[2 * (e * D + d)] = v; }
            else { Ad[2 * (d * D + e) + 1] = -v; Ad[2 * (e * D + d) + 1] = v; }
          }
          __syncwarp();
        };
        build();
        // trace-normalise to tr = D: keeps all classes on a comparable scale
        double tr = 0.0;
        for (int d = lane; d < D; d += 32) tr += A[d * D + d].x;
        tr = warp_sum(tr);
        const double tn = (double)D / fmax(tr, kTiny);
        for (int i = lane; i < NS; i += 32) { A[i].x *= tn; A[i].y *= tn; }
        __syncwarp();
        bool ok;
        double ldk = warp_hpd_inverse<D>(A, lane, &ok);
        double tinv = 0.0;
        for (int d = lane; d < D; d += 32) tinv += A[d * D + d].x;
        tinv = warp_sum(tinv);
        // lambda_min / lambda_max >= 1 / (tr(A) tr(A^-1)) = 1 / (D tinv)
        const bool no_floor = ok && isfinite(tinv) && ((double)D * tinv * a.eigenvalue_floor < 0.5);
        double* __restrict__ co = a.coef + ((size_t)bin * K + k) * NS;
        if (__any_sync(0xffffffffu, bad)) {
          if (lane == 0) atomicMax(a.status, bin + 1);
        }
        if (no_floor) {
          for (int s = lane; s < NS; s += 32) {
            const int pk = sm.tab[s];
            const int d = pk & 255, e = (pk >> 8) & 255, kind = pk >> 16;
            const double2 u = A[d * D + e], v = A[e * D + d];
            co[s] = kind == 0 ? u.x : (kind == 1 ? (u.x + v.x) : -(u.y - v.y));
          }
        } else {
          // reference semantics: eigendecomposition, normalise, floor (cacg.py:95-126)
          build();
          if (a.covariance_norm == PBB_NORM_TRACE) {
            for (int i = lane; i < NS; i += 32) { A[i].x *= tn / D; A[i].y *= tn / D; }
            __syncwarp();
          }
          warp_jacobi(A, sm.V[k], sm.rot[k], D, lane);
          double lmax = -INFINITY;
          for (int d = lane; d < D; d += 32) lmax = fmax(lmax, A[d * D + d].x);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) lmax = fmax(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
          for (int d = lane; d < D; d += 32) {
            double l = A[d * D + d].x;
            if (a.covariance_norm == PBB_NORM_EIGENVALUE) l = fmax(l / fmax(lmax, kTiny), a.eigenvalue_floor);
            else l = fmax(l, lmax * a.eigenvalue_floor);
            if (!isfinite(l)) atomicMax(a.status, bin + 1);
            sm.lam[k][d] = l;
          }
          __syncwarp();
          ldk = model_from_eig_warp(sm.V[k], sm.lam[k], sm.tab, D, lane, co);
        }
        if (lane == 0) sm.ld[k] = ldk;
      }
      __threadfence();
      __syncthreads();
      if (tid < K) {
        const int k = tid;
        double wk;
        if (a.weight_mode == PBB_WEIGHT_CONST) wk = 1.0 / K;
        else if (!(FULL && a.saliency != nullptr)) wk = sm.S[k][NS] / (double)T;
        else {
          double n1 = 0.0;
          for (int j = 0; j < K; ++j) n1 += fabs(sm.S[j][NS]);
          wk = sm.S[k][NS] / (n1 == 0.0 ? 1e-10 : n1);
        }
        double ldmin = sm.ld[0];
        for (int j = 1; j < K; ++j) ldmin = fmin(ldmin, sm.ld[j]);
        a.w[(size_t)bin * K + k] = wk;
        a.ld[(size_t)bin * K + k] = sm.ld[k];
        a.ew[(size_t)bin * K + k] = wk * exp(ldmin - sm.ld[k]);
        __threadfence();
      }
      __syncthreads();
      if (tid == 0) st_release_gpu(a.flags + bin, it + 1);
    }
    cur = nxt;
  }
}

}  // namespace pbb
