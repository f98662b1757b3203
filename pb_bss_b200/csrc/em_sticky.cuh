// "Sticky bins": the cACGMM fit of FEW bins (D = 8, lean variant), one thread-block cluster per bin for ALL iterations.
//
// em_ws_kernel hands (bin, iteration) tasks to whichever CTA is free; the model and the dependency flag of a bin
// travel through L2 between iterations.  With fewer bins than CTA slots (a rank's shard of a bin-sharded utterance,
// short STFTs) that chain -- sweep, partial sums, update, publish, poll, model load: ~15 us even with the frame
// split -- is all there is, and most SMs idle.  Here a cluster of S = 1, 2 or 4 CTAs keeps ONE bin from the first to
// the last iteration:
//   * CTA p of the cluster holds the ring stages [p n / S, (p + 1) n / S) of the bin's staged observation in shared
//     memory for the whole fit (at most kWsStages of them: one TMA bulk copy each, at the start);
//   * per iteration its four EM warps sweep those frames (the hot loop of em_ws_kernel, unchanged) and leave the
//     scatter sums in shared memory -> cluster barrier -> the update warps of CTA 0 add the S partial sums in rank
//     order through DSMEM (bit-reproducible), update the model (cacg_update_class, unchanged) and store it into
//     every CTA's model buffer through DSMEM -> cluster barrier -> every CTA's EM warps start the next sweep.
// No tickets, no flags, no partial sums through L2, no re-streaming of the observation; the E / M arithmetic, the
// update and therefore the results are those of em_ws_kernel with the frame split (same summation order over the
// parts).  The host uses it when F * S CTAs are co-resident-sized (F * S <= 2 x SMs) and a part fits the ring
// (n / S <= kWsStages); otherwise em_ws_kernel runs.
#pragma once
#include <cooperative_groups.h>

#include "em_ws.cuh"

namespace pbb {

#ifndef PBB_STICKY_EM_REGS
// Register budget of the EM warps (the helpers get 256 - this), multiple of 8.  200 / 56 instead of em_ws_kernel's
// 208 / 48: the update warps spill less, and a spill reload behind a cluster barrier (which invalidates the L1) is
// an L2 round trip.  scripts/sticky_regs_ab.py: F = 65: 1.045 -> 1.016 ms, F = 17: 0.949 -> 0.927 ms, same results.
#define PBB_STICKY_EM_REGS 200
#endif

__device__ __forceinline__ void sticky_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int K, typename CT>
__global__ void __launch_bounds__(256, 2) em_sticky_kernel(const PersistArgs a) {
  constexpr int D = 8, MODEL = 0;
  using SM = WsSmem<D, K, CT>;
  using G = GroupDims<D>;
  constexpr int NS = D * D, M = D / 2, NSG = G::NSG;
  constexpr int NU = K < 3 ? K : 3;  // updater warps
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SM& sm = *reinterpret_cast<SM*>(smem_raw);
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int S = (int)cluster.num_blocks(), part = (int)cluster.block_rank();
  const int bin = blockIdx.x / S;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int T = a.T, zs = a.zs;
  const int nchunks = (zs + kStageFrames - 1) / kStageFrames;
  const int c0 = part * nchunks / S, c1 = (part + 1) * nchunks / S;  // this CTA's ring stages (c1 - c0 <= kWsStages)
  constexpr uint32_t kStageBytes = (uint32_t)(SM::ROWS * kStageFrames * sizeof(CT));

  for (int s = tid; s < NS; s += blockDim.x) sm.tab[s] = slot_pack(D, s);
  if (tid == 0) {
    for (int s = 0; s < kWsStages; ++s) mbar_init(&sm.full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    // the part's observation: resident for the whole fit
    const CT* __restrict__ zbase = reinterpret_cast<const CT*>(a.z);
    for (int c = c0; c < c1; ++c) {
      mbar_expect_tx(&sm.full[c - c0], kStageBytes);
      bulk_g2s(&sm.zbuf[c - c0][0][0], zbase + ((size_t)bin * nchunks + c) * (SM::ROWS * kStageFrames), kStageBytes,
               &sm.full[c - c0]);
    }
  }

  if (warp < M) {
    // =============================== EM warps ===============================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(PBB_STICKY_EM_REGS));
    const int g = warp;
    int buf = 0;
#pragma unroll 1
    for (int it = 0; it < a.iterations; ++it) {
      const bool mstep_only = a.first_is_m && it == 0;
      double acc[K * NSG];
#pragma unroll
      for (int i = 0; i < K * NSG; ++i) acc[i] = 0.0;
      double sg[K];
#pragma unroll
      for (int k = 0; k < K; ++k) sg[k] = 0.0;
#pragma unroll 1
      for (int c = c0; c < c1; ++c) {
        const int st = c - c0;
        mbar_wait(&sm.full[st], 0u);  // completes once; later iterations pass straight through
        const int t_chunk = c * kStageFrames;
        const int nsteps = (min(kStageFrames, zs - t_chunk)) >> 5;
        if (!mstep_only) {
          lean_chunk2_split<D, K, CT>(sm, 0, g, st, nsteps >> 1, lane, a.aff_eps, acc, sg);
          if (nsteps & 1) {
            // odd tail step: every group evaluates all 32 frames; only group 0 counts them
            double sgt[K];
#pragma unroll
            for (int k = 0; k < K; ++k) sgt[k] = 0.0;
            buf = 0;
            lean_chunk<D, K, CT, MODEL, true>(sm, 0, g, st, 1, lane, buf, a.aff_eps, acc, sgt, nsteps - 1);
#pragma unroll
            for (int k = 0; k < K; ++k) sg[k] += g == 0 ? sgt[k] : 0.0;
          }
        } else {
          general_chunk<D, K, CT, false>(a, sm, g, bin, st, t_chunk, nsteps, lane, buf, true, true, acc, sg);
        }
      }
      if (!mstep_only && zs > T && c1 == nchunks) {
        // the zs - T padded frames of every row behaved like zero observations: take them out in one place
        double q1[K], gp[K], cp[K];
#pragma unroll
        for (int k = 0; k < K; ++k) q1[k] = 0.0;
        softmax_product<D, K>(q1, sm.ew[0], a.aff_eps, gp, cp);
        const int npad_lane = (g == 0 && lane >= 32 - (zs - T)) ? 1 : 0;
#pragma unroll
        for (int k = 0; k < K; ++k) sg[k] -= npad_lane ? gp[k] : 0.0;
      }
      if (mstep_only && g != 0) {  // the M-step-only pass counts gamma in every group: keep group 0's
#pragma unroll
        for (int k = 0; k < K; ++k) sg[k] = 0.0;
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
            sm.S[0][k][g * NSG + i] = acc[j];
          }
        }
      }
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const double v = warp_sum(sg[k]);
        if (lane == 0) sm.sgp[0][g][k] = v;
      }
      sticky_cluster_sync();  // (1) every part's sums are complete
      sticky_cluster_sync();  // (2) CTA 0 has stored the new model into every CTA's shared memory
      if (it + 1 < a.iterations) {
        // weights and ew from the raw scalars (sum of gamma, log det), as the producer warp of em_ws_kernel does
        if (tid < K) {
          const double ldk = sm.ld[tid];
          double ldmin = ldk;
#pragma unroll
          for (int j = 0; j < K; ++j) ldmin = fmin(ldmin, sm.ld[j]);
          const double sgam = sm.S[1][tid][NS];
          const double wk = a.weight_mode == PBB_WEIGHT_CONST ? 1.0 / K : sgam / (double)T;
          sm.ew[0][tid] = wk * exp(ldmin - ldk);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * M) : "memory");
      }
    }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(256 - PBB_STICKY_EM_REGS));
    const int u = warp - M - 1;  // updater index (warp M only keeps the barriers company)
#pragma unroll 1
    for (int it = 0; it < a.iterations; ++it) {
      sticky_cluster_sync();  // (1)
      if (part == 0 && u >= 0 && u < NU) {
        const bool last_it = it == a.iterations - 1;
        // sums over the parts in rank order (fixed: the result does not depend on timing); sum of gamma = the four
        // groups' shares of every part
        for (int k = u; k < K; k += NU) {
          for (int i = lane; i < NS + 1; i += 32) {
            double v = 0.0;
            for (int p = 0; p < S; ++p) {
              if (i < NS) {
                v += *cluster.map_shared_rank(&sm.S[0][k][i], p);
              } else {
                const double* sp = cluster.map_shared_rank(&sm.sgp[0][0][k], p);
                v += (sp[0] + sp[K]) + (sp[2 * K] + sp[3 * K]);
              }
            }
            sm.S[1][k][i] = v;
          }
        }
        __syncwarp();
        if (last_it) {
          // leave the raw sums for cacg_update_kernel (reference-exact eigendecomposition)
          double* __restrict__ po = a.part + (size_t)bin * K * (NS + 1);
          for (int k = u; k < K; k += NU)
            for (int i = lane; i < NS + 1; i += 32) po[k * (NS + 1) + i] = sm.S[1][k][i];
        } else {
          for (int k = u; k < K; k += NU) {
            // the class model goes to a staging row in this CTA's shared memory, then into every CTA's model buffer
            // through DSMEM: no L2 round trip between the update and the next sweep
            cacg_update_class<D, false>(a, bin, k, K, lane, sm.A[k], sm.V[k], sm.lam[k], sm.S[1][k], sm.tab, &sm.ld[k],
                                        nullptr, &sm.coef[1][k][0]);
            __syncwarp();
            const double c0v = sm.coef[1][k][lane], c1v = sm.coef[1][k][lane + 32];
            const double ldk = sm.ld[k], sgam = sm.S[1][k][NS];
            for (int p = 0; p < S; ++p) {
              double* rc = cluster.map_shared_rank(&sm.coef[0][k][0], p);
              rc[lane] = c0v;
              rc[lane + 32] = c1v;
              if (lane == 0 && p != 0) {
                *cluster.map_shared_rank(&sm.ld[k], p) = ldk;
                *cluster.map_shared_rank(&sm.S[1][k][NS], p) = sgam;
              }
            }
          }
        }
      }
      sticky_cluster_sync();  // (2)
    }
  }
}

}  // namespace pbb
