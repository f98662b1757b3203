// "Lane = slot" persistent cACGMM EM kernel (D = 8 microphones, lean variant) -- round 2 hot path.
//
// Same task model, numerics and global protocol as em_ws.cuh (task = one EM iteration of one bin,
// atomic tickets, per-bin release/acquire flags, Gauss-Jordan model update with the Jacobi
// fallback, raw scatter sums of the last iteration left for cacg_update_kernel), but the arithmetic
// is laid out differently, so that the hot loop has NO barrier between warps, a third less
// shared-memory traffic and four compute warps per scheduler instead of two:
//
//   * One CTA per SM works on ONE task at a time with 16 compute warps; a warp owns a UNIT of 32
//     frames (T = 500 -> 16 units -> one per warp).  Task latency is a quarter of em_ws's and only
//     148 tasks are in flight, so the (bin, it) -> (bin, it + 1) dependency has 3.5 task times of
//     slack at F = 513 instead of 1.7.
//   * E phase of a unit (lane = frame): the two half-warps split the 36 Hermitian slot pairs of
//     psi = z z^H between them (same instruction stream, different channels: see ls_chan /
//     ls_entry_pair), each lane covers frames f and f + 16 of the unit, computes its psi entries on
//     the fly from the 8 channels (never stored) and accumulates the K quadratic forms with the
//     coefficients broadcast from shared memory.  One shuffle exchanges the half-warps' partial
//     sums; every lane then owns the complete q of ONE frame, evaluates the posterior
//     (softmax_product) and writes gamma / q for its frame to the warp's private buffer.
//   * M phase of the same unit (lane = slot pair): lane l < 28 owns the complex off-diagonal entry
//     (d, e) of the K scatter matrices, lanes 28..31 the diagonals 0..3; the warp walks the 32
//     frames, every lane re-reads its two channels of the frame from the (bank-swizzled) row,
//     rebuilds its psi entry (4 ops) and adds it into 2K accumulators with gamma / q broadcast from
//     the buffer.  The accumulators ARE the scatter sums (no replication over lanes, no butterfly
//     reduction at the end of the task); diagonals 4..7 take one extra pass per 8 frames.
//   * psi is computed twice (E and M), 640 + softmax fp64 operations per frame instead of 512, but
//     nothing is exchanged between warps: the 16 partial scatter sums of a task are added by the
//     update warps in a fixed order.
//
// Helper warpgroups (register budgets via setmaxnreg): warp 16 = producer (tickets, flags, model ->
// shared memory in the E-phase order, 1-D TMA bulk copies into an 8-stage ring = this task + the
// next one); warps 17..19 and 21..23 = two sets of update warps (one class each) that alternate
// tasks, so an update may take two task times.
//
// Staged observation layout (normalize_staged_kernel / stream_load_kernel, layout 1): frame-major,
// z[f][chunk][frame][slot] with slot = channel ^ swizzle(frame): a frame is one 128-byte row, so the
// M phase (all lanes the same frame) and the E phase (lane = frame, stride 128 B) are both free of
// bank conflicts, and a ring stage (128 frames) is one contiguous 16 KB block = one UBLKCP.
#pragma once
#include "em_ws.cuh"

namespace pbb {

constexpr int kLsWarps = 16;       // compute warps
constexpr int kLsHelpers = 8;      // helper warps (2 warpgroups)
constexpr int kLsThreads = 32 * (kLsWarps + kLsHelpers);
constexpr int kLsStages = 8;       // ring stages of kStageFrames frames
constexpr int kLsUnit = 32;        // frames per unit
#ifndef PBB_LS_REGS
#define PBB_LS_REGS 104
#endif
constexpr int kLsRegs = PBB_LS_REGS;   // 512 x 104 + 256 x 32 = 61440 = 768 x 80
constexpr int kLsHelperRegs = 32;

// ---- E-phase slot structure ---------------------------------------------------------------------
// Local channel labels of a half-warp: L[0..3] = its own channel set, L[4..7] = the other set
// (rotated by one for the upper half, which makes the 4 x 4 cross block tile exactly once).
__host__ __device__ constexpr int ls_chan(int h, int r) {
  return h == 0 ? r : (r < 4 ? 4 + r : ((r - 4 + 1) & 3));
}
// entry 0, 1: diagonals (L0, L1), (L2, L3); entries 2..7: pairs inside L[0..3]; entries 8..15:
// cross pairs (L[i], L[4 + j]) with j - i in {0, 1} mod 4
struct LsLoc { int u, v; };
__host__ __device__ constexpr LsLoc ls_entry_pair(int e) {
  if (e < 8) {
    const int w = e - 2;
    const int u = w < 3 ? 0 : (w < 5 ? 1 : 2);
    const int v = w < 3 ? w + 1 : (w < 5 ? w - 1 : 3);
    return {u, v};
  }
  const int c = e - 8, i = c >> 1, j = (i + (c & 1)) & 3;
  return {i, 4 + j};
}
// slot (common.cuh order) of |z_c|^2, and of Re psi of the unordered channel pair {a, b}
__host__ __device__ constexpr int ls_diag_slot(int c) {
  for (int s = 0; s < 64; ++s) {
    const SlotInfo si = slot_info(8, s);
    if (si.kind == 0 && si.d == c) return s;
  }
  return -1;
}
__host__ __device__ constexpr int ls_pair_slot(int a, int b, bool* reversed) {
  for (int s = 0; s < 64; ++s) {
    const SlotInfo si = slot_info(8, s);
    if (si.kind != 1) continue;
    if (si.d == a && si.e == b) { *reversed = false; return s; }
    if (si.d == b && si.e == a) { *reversed = true; return s; }
  }
  return -1;
}
// p-th complex slot pair in slot order: psi(d, e) = conj(z_d) z_e, real part in slot s, imaginary in s + 1
struct LsPair { int d, e, s; };
__host__ __device__ constexpr LsPair ls_pair(int p) {
  int n = 0;
  for (int s = 0; s < 64; ++s) {
    const SlotInfo si = slot_info(8, s);
    if (si.kind == 1) {
      if (n == p) return {si.d, si.e, s};
      ++n;
    }
  }
  return {-1, -1, -1};
}

// byte offset swizzle of a frame row: slot = channel ^ ls_swz(frame)
template <typename CT>
__host__ __device__ constexpr int ls_swz(int frame) {
  return sizeof(CT) == 16 ? (frame & 7) : ((frame >> 1) & 7);
}

template <int K, typename CT>
struct LsSmem {
  static constexpr int D = 8, NS = 64;
  static constexpr int NU = K < 3 ? K : 3;
  static constexpr int HALF = 16 * K + 1;  // double2 entries per half-warp model (+1: bank shift between the halves)
  CT zbuf[kLsStages][kStageFrames][D];
  double2 coefE[2][2 * HALF];            // [model buffer][half][entry][class]
  double cw[kLsWarps][kLsUnit * K];      // gamma / q of the warp's unit, frame-major
  double Spart[kLsWarps][K][NS + 1];     // per-warp scatter sums + sum of gamma of the finished task
  double S[2][K][NS + 1];                // per update set: the summed scatter sums
  double2 A[2][K][NS];
  double2 V[2][K][NS];
  double lam[2][K][D];
  double ld[2][K];
  alignas(16) double ew[2][4];
  int tab[NS];
  int tabE[NS];
  int desc[2][4];
  int sdesc[2][4];
  uint64_t full[kLsStages], empty[kLsStages];
  uint64_t model_full[2], model_empty[2];
  uint64_t s_full[2], s_empty;
};

__device__ __forceinline__ double2 ls_lds(const unsigned char* p, double2*) {
  return *reinterpret_cast<const double2*>(p);
}
__device__ __forceinline__ double2 ls_lds(const unsigned char* p, float2*) {
  const float2 v = *reinterpret_cast<const float2*>(p);
  return make_double2((double)v.x, (double)v.y);
}

// ---- E phase of one unit -------------------------------------------------------------------------
// zu: first row of the unit in the ring stage; ce: this half-warp's model; returns the lane's frame
// posterior in gam / cw (lane l owns frame l of the unit)
template <int K, typename CT>
__device__ __forceinline__ void ls_e_phase(const unsigned char* __restrict__ zu, const double2* __restrict__ ce,
                                           const double* __restrict__ ew, int lane, double eps,
                                           double (&gam)[K], double (&cw)[K]) {
  constexpr int RB = 8 * (int)sizeof(CT);
  const int f = lane & 15, h = lane >> 4;
  const int sw = ls_swz<CT>(f);  // frames f and f + 16 share it
  const unsigned char* rowA = zu + f * RB;
  const unsigned char* rowB = rowA + 16 * RB;
  double2 xa[8], xb[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int ch = h == 0 ? r : (r < 4 ? 4 + r : ((r - 4 + 1) & 3));
    const int off = (ch ^ sw) * (int)sizeof(CT);
    xa[r] = ls_lds(rowA + off, (CT*)nullptr);
    xb[r] = ls_lds(rowB + off, (CT*)nullptr);
  }
  double qa[K], qb[K];
  // diagonal entries: (|x0|^2, |x1|^2), (|x2|^2, |x3|^2)
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const double a0 = fma(xa[2 * e].x, xa[2 * e].x, xa[2 * e].y * xa[2 * e].y);
    const double a1 = fma(xa[2 * e + 1].x, xa[2 * e + 1].x, xa[2 * e + 1].y * xa[2 * e + 1].y);
    const double b0 = fma(xb[2 * e].x, xb[2 * e].x, xb[2 * e].y * xb[2 * e].y);
    const double b1 = fma(xb[2 * e + 1].x, xb[2 * e + 1].x, xb[2 * e + 1].y * xb[2 * e + 1].y);
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const double2 c = ce[e * K + k];
      if (e == 0) {
        qa[k] = c.x * a0;
        qb[k] = c.x * b0;
      } else {
        qa[k] = fma(c.x, a0, qa[k]);
        qb[k] = fma(c.x, b0, qb[k]);
      }
      qa[k] = fma(c.y, a1, qa[k]);
      qb[k] = fma(c.y, b1, qb[k]);
    }
  }
  static_for<14>([&](auto ic) {
    constexpr int e = 2 + decltype(ic)::value;
    constexpr LsLoc pr = ls_entry_pair(e);
    const double2 ua = xa[pr.u], va = xa[pr.v], ub = xb[pr.u], vb = xb[pr.v];
    const double ra = fma(ua.x, va.x, ua.y * va.y);
    const double ia = fma(ua.x, va.y, -(ua.y * va.x));
    const double rb = fma(ub.x, vb.x, ub.y * vb.y);
    const double ib = fma(ub.x, vb.y, -(ub.y * vb.x));
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const double2 c = ce[e * K + k];
      qa[k] = fma(c.x, ra, qa[k]);
      qb[k] = fma(c.x, rb, qb[k]);
      qa[k] = fma(c.y, ia, qa[k]);
      qb[k] = fma(c.y, ib, qb[k]);
    }
  });
  // lower half keeps frame f, upper half frame f + 16: swap the partial sums the other half needs
  double q[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const double send = h ? qa[k] : qb[k];
    const double mine = h ? qb[k] : qa[k];
    q[k] = fabs(mine + __shfl_xor_sync(0xffffffffu, send, 16));
  }
  softmax_product<8, K>(q, ew, eps, gam, cw);
}

// ---- M phase of one unit -------------------------------------------------------------------------
// lane l < 28: complex slot pair (dl, el); lanes 28..31: diagonal l - 28 (dl = el).  od / oe: byte
// offsets of the lane's two channels inside an unswizzled row.  acc[2k], acc[2k+1]: real / imaginary
// part of class k; acc8[k]: diagonals 4..7 (lane = (frame & 7) + 8 (diagonal - 4), summed over the
// frame lanes at the end of the task).
template <int K, typename CT>
__device__ __forceinline__ void ls_m_phase(const unsigned char* __restrict__ zu, const double* __restrict__ cwb,
                                           int lane, int od, int oe, double (&acc)[2 * K], double (&acc8)[K]) {
  constexpr int RB = 8 * (int)sizeof(CT);
  constexpr int CS = (int)sizeof(CT);
  const int jx = lane & 7, dq = lane >> 3;
#pragma unroll 1
  for (int g8 = 0; g8 < kLsUnit / 8; ++g8) {
    const unsigned char* __restrict__ zg = zu + g8 * 8 * RB;
    const double* __restrict__ cg = cwb + g8 * 8 * K;
    const int swb = sizeof(CT) == 16 ? 0 : ((g8 & 1) << 2);
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      double cwr[4 * K];
#pragma unroll
      for (int i = 0; i < 2 * K; ++i) {
        const double2 v = *reinterpret_cast<const double2*>(cg + hh * 4 * K + 2 * i);
        cwr[2 * i] = v.x;
        cwr[2 * i + 1] = v.y;
      }
      double2 a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int j = hh * 4 + i;
        const int sw = (sizeof(CT) == 16 ? j : (swb | (j >> 1))) * CS;
        a[i] = ls_lds(zg + j * RB + (od ^ sw), (CT*)nullptr);
        b[i] = ls_lds(zg + j * RB + (oe ^ sw), (CT*)nullptr);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const double pr = fma(a[i].x, b[i].x, a[i].y * b[i].y);
        const double pi = fma(a[i].x, b[i].y, -(a[i].y * b[i].x));
#pragma unroll
        for (int k = 0; k < K; ++k) {
          acc[2 * k] = fma(cwr[i * K + k], pr, acc[2 * k]);
          acc[2 * k + 1] = fma(cwr[i * K + k], pi, acc[2 * k + 1]);
        }
      }
    }
    // diagonals 4..7 of these 8 frames: lane = (frame jx, diagonal 4 + dq)
    {
      const int sw = (sizeof(CT) == 16 ? jx : (swb | (jx >> 1)));
      const double2 a = ls_lds(zg + jx * RB + (((4 + dq) ^ sw) * CS), (CT*)nullptr);
      const double pr = fma(a.x, a.x, a.y * a.y);
#pragma unroll
      for (int k = 0; k < K; ++k) acc8[k] = fma(cg[jx * K + k], pr, acc8[k]);
    }
  }
}

template <int K, typename CT>
__global__ void __launch_bounds__(kLsThreads, 1) em_ls_kernel(const PersistArgs a) {
  constexpr int D = 8, NS = 64;
  using SM = LsSmem<K, CT>;
  constexpr int NU = SM::NU;
  constexpr int RB = 8 * (int)sizeof(CT);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SM& sm = *reinterpret_cast<SM*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int F = a.F, T = a.T, zs = a.zs;
  const int total = a.iterations * F;
  const int nchunks = (zs + kStageFrames - 1) / kStageFrames;
  constexpr uint32_t kStageBytes = (uint32_t)(kStageFrames * RB);

  for (int s = tid; s < NS; s += blockDim.x) sm.tab[s] = slot_pack(D, s);
  if (tid < 64) {
    // E-phase position of every slot: half h, entry e, part -> (slot, sign)
    const int h = tid >> 5, e = (tid >> 1) & 15, part = tid & 1;
    int s;
    bool neg = false;
    if (e < 2) {
      s = ls_diag_slot(ls_chan(h, 2 * e + part));
    } else {
      const LsLoc pr = ls_entry_pair(e);
      bool rev = false;
      s = ls_pair_slot(ls_chan(h, pr.u), ls_chan(h, pr.v), &rev) + part;
      neg = part && rev;
    }
    sm.tabE[s] = tid | (neg ? 256 : 0);
  }
  if (tid == 0) {
    for (int s = 0; s < kLsStages; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 4); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&sm.model_full[s], 1);
      mbar_init(&sm.model_empty[s], kLsWarps);
      mbar_init(&sm.s_full[s], kLsWarps);
    }
    mbar_init(&sm.s_empty, NU);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp < kLsWarps) {
    // =============================== compute warps ===============================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kLsRegs));
    // M-phase identity of the lane
    int od, oe, s_re;
    {
      int dl, el;
      if (lane < 28) {
        const LsPair pr = ls_pair(lane);
        dl = pr.d; el = pr.e; s_re = pr.s;
      } else {
        dl = el = lane - 28;
        s_re = ls_diag_slot(lane - 28);
      }
      od = dl * (int)sizeof(CT);
      oe = el * (int)sizeof(CT);
    }
    const int s_d8 = ls_diag_slot(4 + (lane >> 3));
    const int nunits_all = nchunks * (kStageFrames / kLsUnit);
    double* __restrict__ cwb = sm.cw[warp];
#ifdef PBB_PHASE_TIMING
    long long _tp = clock64();
#undef PBB_PH
#define PBB_PH(i) do { if (warp == 0 && lane == 0) { long long _t = clock64(); atomicAdd(&a.phase[i], (unsigned long long)(_t - _tp)); _tp = _t; } } while (0)
#endif
#pragma unroll 1
    for (unsigned n = 0;; ++n) {
      const int mb = n & 1;
      mbar_wait(&sm.model_full[mb], (n >> 1) & 1u);
      PBB_PH(0);  // wait for the staged model
      const int bin = sm.desc[mb][0], it = sm.desc[mb][1];
      if (bin < 0) {
        // no more tasks: tell both update sets (the other set's next task would have been n + 1)
        if (n > 0) mbar_wait(&sm.s_empty, (n - 1) & 1u);
        if (lane == 0) {
          if (warp == 0) { sm.sdesc[0][0] = -1; sm.sdesc[1][0] = -1; }
          mbar_arrive(&sm.s_full[mb]);
          mbar_arrive(&sm.s_full[mb ^ 1]);
        }
        break;
      }
      const bool mstep_only = a.first_is_m && it == 0;
      double acc[2 * K], acc8[K], sg[K];
#pragma unroll
      for (int i = 0; i < 2 * K; ++i) acc[i] = 0.0;
#pragma unroll
      for (int k = 0; k < K; ++k) { acc8[k] = 0.0; sg[k] = 0.0; }
      const double2* __restrict__ ce = sm.coefE[mb] + (lane >> 4) * SM::HALF;
      const unsigned gbase = n * (unsigned)nchunks;
#pragma unroll 1
      for (int u = warp; u < nunits_all; u += kLsWarps) {
        const unsigned g = gbase + (unsigned)(u >> 2);
        const int st = g % kLsStages;
        const int t0 = u * kLsUnit;
        if (t0 < T) {
          mbar_wait(&sm.full[st], (g / kLsStages) & 1u);
          PBB_PH(2);  // TMA wait
          const unsigned char* __restrict__ zu =
              reinterpret_cast<const unsigned char*>(&sm.zbuf[st][0][0]) + (u & 3) * kLsUnit * RB;
          double gam[K], cw[K];
          const bool valid = t0 + lane < T;
          if (!mstep_only) {
            ls_e_phase<K, CT>(zu, ce, sm.ew[mb], lane, a.aff_eps, gam, cw);
#pragma unroll
            for (int k = 0; k < K; ++k) {
              gam[k] = valid ? gam[k] : 0.0;
              cw[k] = valid ? cw[k] : 0.0;
            }
          } else {
            const int tc = valid ? t0 + lane : 0;
#pragma unroll
            for (int k = 0; k < K; ++k) {
              const double v = __ldcg(a.aff_in + ((size_t)bin * K + k) * T + tc);
              gam[k] = valid ? v : 0.0;
              cw[k] = gam[k];
            }
          }
          __syncwarp();  // the previous unit's M phase is done with the buffer
#pragma unroll
          for (int k = 0; k < K; ++k) {
            sg[k] += gam[k];
            cwb[lane * K + k] = cw[k];
          }
          __syncwarp();
          PBB_PH(3);  // E phase
          ls_m_phase<K, CT>(zu, cwb, lane, od, oe, acc, acc8);
          PBB_PH(4);  // M phase
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[st]);
      }
      if (lane == 0) mbar_arrive(&sm.model_empty[mb]);  // done with this task's model

      // ---- hand the warp's partial scatter sums to the update warps ----
#pragma unroll
      for (int k = 0; k < K; ++k) {
        acc8[k] += __shfl_xor_sync(0xffffffffu, acc8[k], 1);
        acc8[k] += __shfl_xor_sync(0xffffffffu, acc8[k], 2);
        acc8[k] += __shfl_xor_sync(0xffffffffu, acc8[k], 4);
        sg[k] = warp_sum(sg[k]);
      }
      if (n > 0) mbar_wait(&sm.s_empty, (n - 1) & 1u);  // the previous task's partial sums were consumed
      PBB_PH(5);  // wait for the partial-sum buffer
#pragma unroll
      for (int k = 0; k < K; ++k) {
        double* __restrict__ sp = sm.Spart[warp][k];
        sp[s_re] = acc[2 * k];
        if (lane < 28) sp[s_re + 1] = acc[2 * k + 1];
        if ((lane & 7) == 0) sp[s_d8] = acc8[k];
        if (lane == 0) sp[NS] = sg[k];
      }
      if (warp == 0 && lane == 0) { sm.sdesc[mb][0] = bin; sm.sdesc[mb][1] = it; }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.s_full[mb]);
      PBB_PH(6);  // hand-over
    }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kLsHelperRegs));
    const int hw = warp - kLsWarps;
    if (hw == 0) {
      // =============================== producer ===============================
      const CT* __restrict__ zbase = reinterpret_cast<const CT*>(a.z);
      unsigned chunk_cnt = 0;
#pragma unroll 1
      for (unsigned n = 0;; ++n) {
        const int mb = n & 1;
        int t = 0;
        if (lane == 0) t = atomicAdd(a.ticket, 1);
        t = __shfl_sync(0xffffffffu, t, 0);
        int bin = -1, it = 0;
        if (t < total) {
          if (a.order != nullptr) {
            const int v = __ldcg(a.order + t);
            bin = v & 0xffff;
            it = v >> 16;
          } else {
            decode_ticket(t, F, a.iterations, a.wave_c, bin, it);
          }
        }
        const bool mstep_only = a.first_is_m && it == 0;
        const bool late_z = mstep_only && a.wait_load;  // streamed upload: the bin may not have arrived yet
        int issued = 0;
        auto issue_chunks = [&](int upto, bool blocking) {
          while (issued < upto) {
            const int st = chunk_cnt % kLsStages;
            const uint32_t par = ((chunk_cnt / kLsStages) & 1u) ^ 1u;
            if (blocking) {
              mbar_wait_relaxed(&sm.empty[st], par, 100);
            } else {
              uint32_t done;
              asm volatile(
                  "{\n"
                  ".reg .pred p;\n"
                  "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                  "selp.u32 %0, 1, 0, p;\n"
                  "}\n"
                  : "=r"(done)
                  : "r"(smem_u32(&sm.empty[st])), "r"(par)
                  : "memory");
              if (!done) break;
            }
            mbar_expect_tx(&sm.full[st], kStageBytes);
            bulk_g2s(&sm.zbuf[st][0][0], zbase + ((size_t)bin * nchunks + issued) * (kStageFrames * 8), kStageBytes,
                     &sm.full[st]);
            ++chunk_cnt;
            ++issued;
          }
        };
        if (bin >= 0 && lane == 0) {
          if (!late_z) issue_chunks(nchunks, false);
          if (mstep_only) {
            if (a.wait_load) while (ld_acquire_gpu(a.flags + bin) < 0) __nanosleep(200);
          } else {
            while (ld_acquire_gpu(a.flags + bin) < it) {
              issue_chunks(nchunks, false);
              __nanosleep(40);
            }
          }
          if (late_z) asm volatile("fence.proxy.async;" ::: "memory");
        }
        __syncwarp();
        mbar_wait_relaxed(&sm.model_empty[mb], ((n >> 1) & 1u) ^ 1u, 100);  // compute warps are done with task n - 2
        if (bin >= 0 && !mstep_only) {
          const double* __restrict__ cf = a.coef + (size_t)bin * K * NS;
          double* __restrict__ dst = reinterpret_cast<double*>(sm.coefE[mb]);
          for (int i = lane; i < K * NS; i += 32) {
            const int k = i >> 6, s = i & 63;
            const int te = sm.tabE[s];
            const int h = (te >> 5) & 1, e = (te >> 1) & 15, part = te & 1;
            const double v = __ldcg(cf + i);
            dst[((h * SM::HALF + e * K + k) << 1) + part] = (te & 256) ? -v : v;
          }
          if (lane < K) {
            const double ldk = __ldcg(a.ld + (size_t)bin * 4 + lane);
            double ldmin = ldk;
#pragma unroll
            for (int j = 0; j < K; ++j) ldmin = fmin(ldmin, __ldcg(a.ld + (size_t)bin * 4 + j));
            const double sgam = __ldcg(a.ew + (size_t)bin * 4 + lane);
            const double wk = a.weight_mode == PBB_WEIGHT_CONST ? 1.0 / K : sgam / (double)T;
            sm.ew[mb][lane] = wk * exp(ldmin - ldk);
          }
        }
        if (lane == 0) { sm.desc[mb][0] = bin; sm.desc[mb][1] = it; }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.model_full[mb]);
        if (bin < 0) break;
        if (lane == 0) issue_chunks(nchunks, true);
        __syncwarp();
      }
    } else if (((hw - 1) & 3) < NU) {
      // =============================== update warps ===============================
      const int set = (hw - 1) >> 2, u = (hw - 1) & 3;
#pragma unroll 1
      for (unsigned n = set;; n += 2) {
        mbar_wait_relaxed(&sm.s_full[set], (n >> 1) & 1u, 100);
        const int bin = sm.sdesc[set][0], it = sm.sdesc[set][1];
        if (bin < 0) break;
        const bool last_it = it == a.iterations - 1;
        // partial sums of the 16 compute warps, fixed order
        for (int k = u; k < K; k += NU) {
          for (int s = lane; s < NS + 1; s += 32) {
            double v = sm.Spart[0][k][s];
#pragma unroll
            for (int w = 1; w < kLsWarps; ++w) v += sm.Spart[w][k][s];
            sm.S[set][k][s] = v;
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.s_empty);
        if (last_it) {
          asm volatile("bar.sync %0, %1;" ::"r"(2 + set), "n"(NU * 32) : "memory");
          double* __restrict__ po = a.part + (size_t)bin * K * (NS + 1);
          for (int i = u * 32 + lane; i < K * (NS + 1); i += NU * 32) po[i] = (&sm.S[set][0][0])[i];
        } else {
          for (int k = u; k < K; k += NU)
            cacg_update_class<D, false>(a, bin, k, K, lane, sm.A[set][k], sm.V[set][k], sm.lam[set][k], sm.S[set][k],
                                        sm.tab, &sm.ld[set][k]);
          __syncwarp();
          asm volatile("bar.sync %0, %1;" ::"r"(2 + set), "n"(NU * 32) : "memory");
          if (u == 0 && lane == 0) st_release_gpu(a.flags + bin, it + 1);
        }
      }
    }
  }
}

}  // namespace pbb
