// Persistent cACGMM EM kernel, "quad-lane" variant (lean path: no saliency, no
// activity mask, product-form softmax).  Same task flow as em_persistent.cuh
// (ticket-scheduled (bin, iteration) tasks, per-bin release/acquire flags,
// Gauss-Jordan model update, raw sums for the final eigendecomposition), but a
// different decomposition INSIDE the CTA:
//
//   * the 4 warps of a CTA split the bin's FRAMES; warp w streams its own
//     quarter of every observation row through a private 2-stage TMA ring;
//   * inside a warp, lane = (slot group g, frame j): the D/2 slot groups of one
//     frame sit in lanes j, j+FS, j+2FS, ... so the partial quadratic forms are
//     combined with two shuffle-xor steps instead of shared memory + a CTA
//     barrier per step.
//
// A task therefore needs two CTA barriers (before / after the model update)
// instead of ~20, warps drift freely against each other inside the EM pass,
// and every warp still runs one identical instruction stream.
#pragma once
#include "common.cuh"
#include "em_kernels.cuh"
#include "em_persistent.cuh"
#include "heig.cuh"

namespace pbb {

constexpr int kQuadWarps = 4;        // warps per CTA = frame quarters of a bin
constexpr int kQuadStageFrames = 32; // frames per private ring stage

template <int D>
struct QuadDims {
  static constexpr int M = D / 2;
  static constexpr int MP = M <= 1 ? 1 : (M <= 2 ? 2 : (M <= 4 ? 4 : 8));  // groups padded to a power of two
  static constexpr int FS = 32 / MP;                                       // frames per warp sub-step
  static constexpr int NSG = group_shape(D).nsg;
  static constexpr int NSGP = NSG + 2;  // padded group stride of the coefficient table (bank spread)
};

template <int D, int K, typename CT>
struct QuadSmem {
  static constexpr int NS = D * D;
  static constexpr int M = D / 2;
  static constexpr int ROWS = stage_rows(D);
  CT zbuf[kQuadWarps][kStages][ROWS][kQuadStageFrames];
  double2 A[K][NS];
  double2 V[K][NS];
  double coef[K][M][QuadDims<D>::NSGP];
  double Sp[kQuadWarps][K][NS + 1];  // per-warp partial scatter sums
  double S[K][NS + 1];
  double rot[K][((D + 1) / 2) * 6];
  double lam[K][D];
  double ld[K], ew[K];
  uint64_t full[kQuadWarps][kStages];
  int tab[NS];
  int tick[2];
};

// Sum N per-lane values over the FS lanes that share a slot group (lane bits
// below log2(FS)) with the halving butterfly; afterwards lane j holds the
// totals of indices [lo, hi), at most ceil(N / FS) of them.
template <int N, int FS>
__device__ __forceinline__ void group_reduce_halving(double (&v)[N], int lane, int& lo, int& hi) {
  lo = 0; hi = N;
  int live = N;
#pragma unroll
  for (int off = FS / 2; off >= 1; off >>= 1) {
    const int nh = (live + 1) / 2;
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      if (i < nh) {
        const double a = v[i];
        const double b = (i + nh < live) ? v[i + nh] : 0.0;
        const double keep = upper ? b : a;
        const double send = upper ? a : b;
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
      }
    }
    if (upper) lo += nh; else hi = min(hi, lo + nh);
    live = nh;
  }
  if (hi < lo) hi = lo;
}

// CREG: keep the group's K*NSG model coefficients in registers for the whole task (no
// shared-memory loads in the E-step); needs the 255-register budget, i.e. 2 CTAs per SM.
template <int D, int K, typename CT, int FPL, bool CREG>
__global__ void __launch_bounds__(32 * kQuadWarps, (FPL == 2 || CREG) ? 2 : 3) em_quad_kernel(const PersistArgs a) {
  using SM = QuadSmem<D, K, CT>;
  using Q = QuadDims<D>;
  using G = GroupDims<D>;
  constexpr int NS = D * D, M = Q::M, MP = Q::MP, FS = Q::FS, NSG = Q::NSG, NSGP = Q::NSGP, NLOC = G::NLOC;
  constexpr int FPS = FS * FPL;  // frames per warp step
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SM& sm = *reinterpret_cast<SM*>(smem_raw);
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int g = lane / FS, j = lane % FS;   // slot group, frame within the sub-step
  const int gc = g < M ? g : M - 1;         // ghost lanes (D=6) read a valid group and contribute nothing
  const bool real = g < M;
  const int F = a.F, T = a.T, zs = a.zs;
  const int total = a.iterations * F;
  // warp w owns frames [w * fpw, (w + 1) * fpw) of every row
  const int fpw = ((zs / 32 + kQuadWarps - 1) / kQuadWarps) * 32;
  const int t_lo = min(w * fpw, zs), t_hi = min(t_lo + fpw, zs);
  const int nchunks = (t_hi - t_lo + kQuadStageFrames - 1) / kQuadStageFrames;  // warp-private chunks per task
  const CT* __restrict__ zbase = reinterpret_cast<const CT*>(a.z);

  for (int s = tid; s < NS; s += blockDim.x) sm.tab[s] = slot_pack(D, s);
  if (lane == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&sm.full[w][s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid == 0) sm.tick[0] = atomicAdd(a.ticket, 1);
  __syncthreads();
  int cur = sm.tick[0];
  unsigned chunk_cnt = 0;  // chunks consumed by THIS WARP (ring position)

  auto issue_chunk = [&](int bin, int c, unsigned n) {  // whole warp: lane r copies staged row r
    const int st = n & 1u;
    const int t0 = t_lo + c * kQuadStageFrames;
    const int nf = min(kQuadStageFrames, t_hi - t0);
    const uint32_t bytes = (uint32_t)nf * sizeof(CT);
    if (lane == 0) mbar_expect_tx(&sm.full[w][st], bytes * SM::ROWS);
    if (lane < SM::ROWS)
      bulk_g2s(&sm.zbuf[w][st][lane][0], zbase + ((size_t)bin * D + row_channel(D, lane)) * zs + t0, bytes,
               &sm.full[w][st]);
  };
  if (cur < total && nchunks > 0) issue_chunk(cur % F, 0, 0);

  while (cur < total) {
    const int it = cur / F, bin = cur - it * F;
    const bool mstep_only = a.first_is_m && it == 0;
    const bool last_it = it == a.iterations - 1;
    if (tid == 0) {
      sm.tick[1] = atomicAdd(a.ticket, 1);
      if (!mstep_only) {
        while (ld_acquire_gpu(a.flags + bin) < it) __nanosleep(40);
      }
    }
    __syncthreads();  // (B1) ticket + flag; previous task's shared state is dead
    const int nxt = sm.tick[1];
    if (!mstep_only) {
      const double* __restrict__ cf = a.coef + (size_t)bin * K * NS;
      for (int i = tid; i < K * NS; i += blockDim.x) {
        const int k = i / NS, s = i - k * NS;
        sm.coef[k][s / NSG][s % NSG] = __ldcg(cf + i);
      }
      if (tid < K) sm.ew[tid] = __ldcg(a.ew + (size_t)bin * K + tid);
    }
    __syncthreads();  // (B2) model staged

    double acc[K * NSG];
#pragma unroll
    for (int i = 0; i < K * NSG; ++i) acc[i] = 0.0;
    double sg[K];
#pragma unroll
    for (int k = 0; k < K; ++k) sg[k] = 0.0;
    const double* __restrict__ cg = &sm.coef[0][gc][0];
    const double eps = a.aff_eps;
    double cf[CREG ? K * NSG : 1];
    if constexpr (CREG) {
      if (!mstep_only) {
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
          for (int i = 0; i < NSG; ++i) cf[k * NSG + i] = cg[k * (M * NSGP) + i];
      }
    }

#pragma unroll 1
    for (int c = 0; c < nchunks; ++c) {
      __syncwarp();  // every lane is done with the stage that gets refilled now
      if (c + 1 < nchunks) issue_chunk(bin, c + 1, chunk_cnt + 1);
      else if (nxt < total) issue_chunk(nxt % F, 0, chunk_cnt + 1);
      const int st = chunk_cnt & 1u;
      mbar_wait(&sm.full[w][st], (chunk_cnt >> 1) & 1u);
      ++chunk_cnt;
      const int t_chunk = t_lo + c * kQuadStageFrames;
      const int nfr = min(kQuadStageFrames, t_hi - t_chunk);
      const CT* __restrict__ zrow = &sm.zbuf[w][st][2 * gc][0] + j;
#pragma unroll 1
      for (int f0 = 0; f0 < nfr; f0 += FPS, zrow += FPS) {
        double psi[FPL][NSG];
#pragma unroll
        for (int u = 0; u < FPL; ++u) {
          double2 x[NLOC];
#pragma unroll
          for (int l = 0; l < NLOC; ++l) x[l] = lds_cplx(zrow + l * kQuadStageFrames + u * FS);
          group_psi<D>(x, psi[u]);
        }
        double gam[FPL][K], cw[FPL][K];
        if (!mstep_only) {
          double p0[FPL][K], p1[FPL][K];
#pragma unroll
          for (int u = 0; u < FPL; ++u)
#pragma unroll
            for (int k = 0; k < K; ++k) { p0[u][k] = 0.0; p1[u][k] = 0.0; }
#pragma unroll
          for (int i = 0; i < NSG; i += 2) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
              double2 cc;
              if constexpr (CREG) cc = make_double2(cf[k * NSG + i], cf[k * NSG + i + 1]);
              else cc = *reinterpret_cast<const double2*>(cg + k * (M * NSGP) + i);
#pragma unroll
              for (int u = 0; u < FPL; ++u) {
                p0[u][k] = fma(cc.x, psi[u][i], p0[u][k]);
                p1[u][k] = fma(cc.y, psi[u][i + 1], p1[u][k]);
              }
            }
          }
#pragma unroll
          for (int u = 0; u < FPL; ++u) {
            double q[K];
#pragma unroll
            for (int k = 0; k < K; ++k) {
              double v = p0[u][k] + p1[u][k];
              if (MP != M) v = real ? v : 0.0;
#pragma unroll
              for (int off = FS; off < 32; off <<= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
              q[k] = fabs(v);
            }
            softmax_product<D, K>(q, sm.ew, eps, gam[u], cw[u]);
          }
        } else {
#pragma unroll
          for (int u = 0; u < FPL; ++u) {
            const int t = t_chunk + f0 + u * FS + j;
            const bool valid = t < T;
#pragma unroll
            for (int k = 0; k < K; ++k) {
              const double gk = valid ? a.aff_in[((size_t)bin * K + k) * T + t] : 0.0;
              gam[u][k] = gk;
              cw[u][k] = gk;
            }
          }
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
#pragma unroll
          for (int u = 0; u < FPL; ++u) {
            sg[k] += gam[u][k];
#pragma unroll
            for (int i = 0; i < NSG; ++i) acc[k * NSG + i] = fma(cw[u][k], psi[u][i], acc[k * NSG + i]);
          }
        }
      }
    }

    // ---- per-warp partial sums -> shared memory --------------------------------------
    {
      int lo, hi;
      group_reduce_halving<K * NSG, FS>(acc, lane, lo, hi);
      constexpr int PER = (K * NSG + FS - 1) / FS;
#pragma unroll
      for (int r = 0; r < PER; ++r) {
        const int idx = lo + r;
        if (idx < hi && real) {
          const int k = idx / NSG, i = idx - k * NSG;
          sm.Sp[w][k][g * NSG + i] = acc[r];
        }
      }
#pragma unroll
      for (int k = 0; k < K; ++k) {
        double v = sg[k];
#pragma unroll
        for (int off = FS / 2; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) sm.Sp[w][k][NS] = v;
      }
    }
    __syncthreads();  // (B3) all partials written
    for (int i = tid; i < K * (NS + 1); i += blockDim.x) {
      double v = (&sm.Sp[0][0][0])[i];
#pragma unroll
      for (int ww = 1; ww < kQuadWarps; ++ww) v += (&sm.Sp[ww][0][0])[i];
      (&sm.S[0][0])[i] = v;
    }
    __syncthreads();  // (B4)
    if (!mstep_only && zs > T && tid < K) {
      // the zs - T zero-padded frames of every row behaved like zero observations:
      // remove their posterior mass from the sum of gamma
      double q1[K], gp[K], cp[K];
#pragma unroll
      for (int k = 0; k < K; ++k) q1[k] = 0.0;
      softmax_product<D, K>(q1, sm.ew, eps, gp, cp);
      double corr = 0.0;
#pragma unroll
      for (int k = 0; k < K; ++k) corr = (k == tid) ? gp[k] : corr;
      sm.S[tid][NS] -= (double)(zs - T) * corr;
    }
    __syncthreads();  // (B5)

    if (last_it) {
      double* __restrict__ po = a.part + (size_t)bin * K * (NS + 1);
      for (int i = tid; i < K * (NS + 1); i += blockDim.x) po[i] = (&sm.S[0][0])[i];
    } else {
      for (int k = w; k < K; k += kQuadWarps) {
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
        double tr = 0.0;
        for (int d = lane; d < D; d += 32) tr += A[d * D + d].x;
        tr = warp_sum(tr);
        const double tn = (double)D / fmax(tr, kTiny);
        for (int i = lane; i < NS; i += 32) { A[i].x *= tn; A[i].y *= tn; }
        __syncwarp();
        bool ok;
        const double det = warp_hpd_inverse<D>(A, lane, &ok);
        double ldk = log(det);
        double tinv = 0.0;
        for (int d = lane; d < D; d += 32) tinv += A[d * D + d].x;
        tinv = warp_sum(tinv);
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
      __syncthreads();  // (B6)
      if (tid < K) {
        const int k = tid;
        const double wk = a.weight_mode == PBB_WEIGHT_CONST ? 1.0 / K : sm.S[k][NS] / (double)T;
        double ldmin = sm.ld[0];
        for (int jj = 1; jj < K; ++jj) ldmin = fmin(ldmin, sm.ld[jj]);
        a.w[(size_t)bin * K + k] = wk;
        a.ld[(size_t)bin * K + k] = sm.ld[k];
        a.ew[(size_t)bin * K + k] = wk * exp(ldmin - sm.ld[k]);
      }
      __syncthreads();  // (B7) then one cumulative gpu-scope release
      if (tid == 0) st_release_gpu(a.flags + bin, it + 1);
    }
    cur = nxt;
  }
}

}  // namespace pbb
