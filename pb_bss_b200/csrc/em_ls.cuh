// "Lane = slot" persistent cACGMM EM kernel (D = 8 microphones, lean variant) -- round 2 hot path.
//
// Same task model, numerics and global protocol as em_ws.cuh (task = one EM iteration of one bin,
// atomic tickets, per-bin release/acquire flags, Gauss-Jordan model update with the Jacobi
// fallback, raw scatter sums of the last iteration left for cacg_update_kernel), but the arithmetic
// is laid out differently, so that the hot loop has NO barrier between warps, a third less
// shared-memory traffic and four compute warps per scheduler instead of two:
//
//   * A CTA (two per SM) works on one task at a time with 8 compute warps; a warp owns UNITS of 32
//     frames (T = 500 -> 16 units -> two per warp).  The two CTAs of an SM are never in step, so the
//     task boundary of one (hand-over, next model, ring wait) overlaps the arithmetic of the other.
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
// Operand order matters: a DFMA that reads three different 64-bit registers issues every 3 cycles on
// B200, one that finds an operand in the reuse cache every 2 (scripts/microbench/fp64_operands.cu),
// so the FMA blocks walk the classes in a snake that shares one operand between neighbours.
//
// Helper warpgroup (register budgets via setmaxnreg): warp 8 = producer (tickets, flags, model ->
// shared memory in the E-phase order, 1-D TMA bulk copies into the ring: this task's chunks + the
// first of the next task); warps 9..11 = update warps (one class each), which work on task n while
// the compute warps already run task n + 1.
//
// Staged observation layout (normalize_staged_kernel / stream_load_kernel, layout 1): frame-major,
// z[f][chunk][frame][slot] with slot = channel ^ swizzle(frame): a frame is one 128-byte row, so the
// M phase (all lanes the same frame) and the E phase (lane = frame, stride 128 B) are both free of
// bank conflicts, and a ring stage (128 frames) is one contiguous 16 KB block = one UBLKCP.
#pragma once
#include "em_ws.cuh"

namespace pbb {

constexpr int kLsWarps = 8;        // compute warps (2 warpgroups)
constexpr int kLsHelpers = 4;      // helper warps (1 warpgroup)
constexpr int kLsThreads = 32 * (kLsWarps + kLsHelpers);
constexpr int kLsUnit = 32;        // frames per unit
#ifndef PBB_LS_REGS
#define PBB_LS_REGS 96
#endif
constexpr int kLsRegs = PBB_LS_REGS;   // 2 CTAs / SM: 256 x 96 + 128 x 48 = 30720 = 384 x 80
constexpr int kLsHelperRegs = (384 * 80 - 256 * PBB_LS_REGS) / 128;
constexpr int kLsSmemBudget = 113 * 1024;  // per CTA, two CTAs per SM
#ifndef PBB_LS_LEAD
#define PBB_LS_LEAD 0
#endif
#ifndef PBB_LS_HELPERS_FIRST
#define PBB_LS_HELPERS_FIRST 1
#endif

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
struct LsSmemRest {
  static constexpr int D = 8, NS = 64;
  static constexpr int HALF = 16 * K + 1;  // double2 entries per half-warp model (+1: bank shift between the halves)
  struct alignas(16) Model {
    double2 c[2 * HALF];                 // [half][entry][class]
    double ew[4];                        // w_k exp(ld_min - ld_k)
  };
  static_assert(sizeof(Model) == ls_model_doubles(K) * sizeof(double), "model block layout");
  Model model[2];                        // current / next task, filled by one TMA bulk copy each
  double cw[kLsWarps][2 * kLsUnit * K];  // gamma / q of the warp's unit, frame-major, then gamma
  double Spart[kLsWarps][K][NS + 1];     // per-warp scatter sums + sum of gamma of the finished task
  double S[K][NS + 1];                   // update warps: the summed scatter sums
  double2 A[K][NS];
  double2 V[K][NS];
  double lam[K][D];
  double ld[2][K];                       // log det of the classes, by task parity (read by the publishing warp)
  double sgam[2][K];                     // sum of gamma, by task parity
  int tab[NS];
  int tabE[NS];
  int desc[2][4];
  int sdesc[4];
  uint64_t full[8], empty[8];
  uint64_t model_full[2], model_empty[2];
  uint64_t s_full, s_empty;
};
// ring stages that fit next to the rest (this task's chunks + a head start on the next task's)
template <int K, typename CT>
constexpr int ls_stages() {
  const int n = (kLsSmemBudget - (int)sizeof(LsSmemRest<K, CT>)) / (int)(kStageFrames * 8 * sizeof(CT));
  return n > 8 ? 8 : n;
}
template <int K, typename CT>
struct LsSmem : LsSmemRest<K, CT> {
  static constexpr int NU = K < 3 ? K : 3;
  static constexpr int STAGES = ls_stages<K, CT>();
  CT zbuf[STAGES][kStageFrames][8];
};

__device__ __forceinline__ double2 ls_lds(const unsigned char* p, double2*) {
  return *reinterpret_cast<const double2*>(p);
}
__device__ __forceinline__ double2 ls_lds(const unsigned char* p, float2*) {
  const float2 v = *reinterpret_cast<const float2*>(p);
  return make_double2((double)v.x, (double)v.y);
}

// q_k += c_k.x * r + c_k.y * i for two frames (a, b), classes walked in a snake so that neighbouring
// DFMAs share the coefficient or the psi operand (operand reuse cache)
template <int K>
__device__ __forceinline__ void ls_q_fma(const double2 (&c)[K], double ra, double ia, double rb, double ib,
                                         double (&qa)[K], double (&qb)[K]) {
#pragma unroll
  for (int k = 0; k < K; ++k) {
    if ((k & 1) == 0) {
      qa[k] = fma(c[k].x, ra, qa[k]);
      qb[k] = fma(c[k].x, rb, qb[k]);
    } else {
      qb[k] = fma(c[k].x, rb, qb[k]);
      qa[k] = fma(c[k].x, ra, qa[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < K; ++k) {
    if ((k & 1) == 0) {
      qa[k] = fma(c[k].y, ia, qa[k]);
      qb[k] = fma(c[k].y, ib, qb[k]);
    } else {
      qb[k] = fma(c[k].y, ib, qb[k]);
      qa[k] = fma(c[k].y, ia, qa[k]);
    }
  }
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
  auto chan_off = [&](int r) {
    const int ch = h == 0 ? r : (r < 4 ? 4 + r : ((r - 4 + 1) & 3));
    return (ch ^ sw) * (int)sizeof(CT);
  };
  double2 xa[8], xb[8];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int off = chan_off(r);
    xa[r] = ls_lds(rowA + off, (CT*)nullptr);
    xb[r] = ls_lds(rowB + off, (CT*)nullptr);
  }
  double qa[K], qb[K];
#pragma unroll
  for (int k = 0; k < K; ++k) { qa[k] = 0.0; qb[k] = 0.0; }
  // diagonal entries: (|x0|^2, |x1|^2), (|x2|^2, |x3|^2)
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const double a0 = fma(xa[2 * e].x, xa[2 * e].x, xa[2 * e].y * xa[2 * e].y);
    const double a1 = fma(xa[2 * e + 1].x, xa[2 * e + 1].x, xa[2 * e + 1].y * xa[2 * e + 1].y);
    const double b0 = fma(xb[2 * e].x, xb[2 * e].x, xb[2 * e].y * xb[2 * e].y);
    const double b1 = fma(xb[2 * e + 1].x, xb[2 * e + 1].x, xb[2 * e + 1].y * xb[2 * e + 1].y);
    double2 c[K];
#pragma unroll
    for (int k = 0; k < K; ++k) c[k] = ce[e * K + k];
    ls_q_fma<K>(c, a0, a1, b0, b1, qa, qb);
  }
  // pairs; the cross entries run 15, 8, 9, ..., 14 so that each channel of the other set is needed by two
  // consecutive entries only and is loaded right before them
  static_for<14>([&](auto ic) {
    constexpr int i = decltype(ic)::value;
    constexpr int e = i < 6 ? 2 + i : (i == 6 ? 15 : 8 + (i - 7));
    constexpr LsLoc pr = ls_entry_pair(e);
    if constexpr (i >= 6 && ((i - 6) & 1) == 0) {
      const int off = chan_off(pr.v);
      xa[pr.v] = ls_lds(rowA + off, (CT*)nullptr);
      xb[pr.v] = ls_lds(rowB + off, (CT*)nullptr);
    }
    const double2 ua = xa[pr.u], va = xa[pr.v], ub = xb[pr.u], vb = xb[pr.v];
    const double ta = ua.y * va.y, tb = ub.y * vb.y;
    const double sa = ua.y * va.x, sb = ub.y * vb.x;
    const double ra = fma(ua.x, va.x, ta);
    const double ia = fma(ua.x, va.y, -sa);
    const double rb = fma(ub.x, vb.x, tb);
    const double ib = fma(ub.x, vb.y, -sb);
    double2 c[K];
#pragma unroll
    for (int k = 0; k < K; ++k) c[k] = ce[e * K + k];
    ls_q_fma<K>(c, ra, ia, rb, ib, qa, qb);
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
// frame lanes at the end of the task); accg[k]: sum of gamma (every lane group keeps the same copy).
// The loads of frame pair p + 1 are in flight while pair p is accumulated (two register sets).
template <int K, typename CT>
struct LsPairRegs {
  double2 a[2], b[2];
  double c[2 * K];
};
template <int K, typename CT>
__device__ __forceinline__ void ls_m_load(LsPairRegs<K, CT>& r, const unsigned char* __restrict__ zg,
                                          const double* __restrict__ cg, int i0, int swb, int od, int oe) {
  constexpr int RB = 8 * (int)sizeof(CT);
  constexpr int CS = (int)sizeof(CT);
#pragma unroll
  for (int i = 0; i < K; ++i) {
    const double2 v = *reinterpret_cast<const double2*>(cg + i0 * K + 2 * i);
    r.c[2 * i] = v.x;
    r.c[2 * i + 1] = v.y;
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int j = i0 + i;
    const int sw = (sizeof(CT) == 16 ? j : (swb | (j >> 1))) * CS;
    r.a[i] = ls_lds(zg + j * RB + (od ^ sw), (CT*)nullptr);
    r.b[i] = ls_lds(zg + j * RB + (oe ^ sw), (CT*)nullptr);
  }
}
template <int K, typename CT>
__device__ __forceinline__ void ls_m_fma(const LsPairRegs<K, CT>& r, double (&acc)[2 * K]) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const double t = r.a[i].y * r.b[i].y, u = r.a[i].y * r.b[i].x;
    const double pr = fma(r.a[i].x, r.b[i].x, t);
    const double pi = fma(r.a[i].x, r.b[i].y, -u);
    // snake over the classes: neighbours share gamma / q or the psi entry (operand reuse cache)
#pragma unroll
    for (int k = 0; k < K; ++k) {
      if ((k & 1) == 0) {
        acc[2 * k] = fma(r.c[i * K + k], pr, acc[2 * k]);
        acc[2 * k + 1] = fma(r.c[i * K + k], pi, acc[2 * k + 1]);
      } else {
        acc[2 * k + 1] = fma(r.c[i * K + k], pi, acc[2 * k + 1]);
        acc[2 * k] = fma(r.c[i * K + k], pr, acc[2 * k]);
      }
    }
  }
}
template <int K, typename CT>
__device__ __forceinline__ void ls_m_phase(const unsigned char* __restrict__ zu, const double* __restrict__ cwb,
                                           int lane, int od, int oe, double (&acc)[2 * K], double (&acc8)[K],
                                           double (&accg)[K]) {
  constexpr int RB = 8 * (int)sizeof(CT);
  constexpr int CS = (int)sizeof(CT);
  constexpr int NG = kLsUnit / 8;
  const int jx = lane & 7, dq = lane >> 3;
  const double* __restrict__ gmb = cwb + kLsUnit * K;  // gamma of the unit's frames
  LsPairRegs<K, CT> r0, r1;
  ls_m_load<K, CT>(r0, zu, cwb, 0, 0, od, oe);
#pragma unroll 1
  for (int g8 = 0; g8 < NG; ++g8) {
    const unsigned char* __restrict__ zg = zu + g8 * 8 * RB;
    const double* __restrict__ cg = cwb + g8 * 8 * K;
    const int swb = sizeof(CT) == 16 ? 0 : ((g8 & 1) << 2);
    // diagonals 4..7 and the sum of gamma of these 8 frames: lane = (frame jx, diagonal 4 + dq); loads first
    const int swx = (sizeof(CT) == 16 ? jx : (swb | (jx >> 1)));
    const double2 ax = ls_lds(zg + jx * RB + (((4 + dq) ^ swx) * CS), (CT*)nullptr);
    double cx[K], gx[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      cx[k] = cg[jx * K + k];
      gx[k] = gmb[(g8 * 8 + jx) * K + k];
    }
    ls_m_load<K, CT>(r1, zg, cg, 2, swb, od, oe);
    ls_m_fma<K, CT>(r0, acc);
    ls_m_load<K, CT>(r0, zg, cg, 4, swb, od, oe);
    ls_m_fma<K, CT>(r1, acc);
    ls_m_load<K, CT>(r1, zg, cg, 6, swb, od, oe);
    ls_m_fma<K, CT>(r0, acc);
    {
      // first pair of the next group (the last group re-reads its own: nothing beyond the unit is touched)
      const int gn = g8 + 1 < NG ? g8 + 1 : g8;
      const int swn = sizeof(CT) == 16 ? 0 : ((gn & 1) << 2);
      ls_m_load<K, CT>(r0, zu + gn * 8 * RB, cwb + gn * 8 * K, 0, swn, od, oe);
    }
    ls_m_fma<K, CT>(r1, acc);
    const double px = fma(ax.x, ax.x, ax.y * ax.y);
#pragma unroll
    for (int k = 0; k < K; ++k) {
      acc8[k] = fma(cx[k], px, acc8[k]);
      accg[k] += gx[k];
    }
  }
}

// ---- fast model update of one (bin, class) by one warp ----------------------------------------------
// The E-step only needs B^-1 and log det B up to a common scale per class (em_persistent.cuh), so the
// scatter matrix is scaled by an exact power of two (trace in [1, 2): no division, no rounding) and
// inverted by Gauss-Jordan elimination held in REGISTERS: lane l owns entries (l >> 3, l & 7) and
// (4 + (l >> 3), l & 7); pivot row, pivot column and pivot travel by warp shuffles, so a pivot step is one
// round of independent shuffles plus ~6 dependent fp64 operations (the reciprocal of the pivot overlaps the
// products).  Returns false when the reference would floor an eigenvalue (or anything is not finite):
// the caller then runs cacg_update_class, which has the reference's eigendecomposition semantics.
// em[r]: slot of the lane's entry r (bit 8: diagonal, bit 9: this lane stores the coefficients, bit 10: the
// entry is psi itself, not its conjugate).
template <int K>
__device__ __forceinline__ bool ls_update_fast(const PersistArgs& a, int bin, int k, int lane,
                                               const double* __restrict__ Sk, const int (&em)[2],
                                               const int* __restrict__ tabE, double* __restrict__ ld_out) {
  constexpr int NS = 64;
  double trs;
  {
    double d[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) d[c] = Sk[ls_diag_slot(c)];
    trs = ((d[0] + d[1]) + (d[2] + d[3])) + ((d[4] + d[5]) + (d[6] + d[7]));
  }
  const int ex = (__double2hiint(trs) >> 20) & 0xfff;  // sign + biased exponent
  bool ok = ex > 0 && ex < 0x7fe;                      // positive, normal, finite
  const double f = __hiloint2double((2046 - (ex & 0x7ff)) << 20, 0);
  double2 m[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int sl = em[r] & 255;
    const double re = Sk[sl] * f;
    double im = 0.0;
    if (!(em[r] & 256)) {
      im = Sk[sl + 1] * f;
      im = (em[r] & 1024) ? im : -im;
    }
    ok = ok && isfinite(re) && isfinite(im);
    m[r] = make_double2(re, im);
  }
  double det = 1.0, pmin = 1.0;
  const int row = lane >> 3, col = lane & 7;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    constexpr unsigned FULLM = 0xffffffffu;
    const int rj = j >> 2;
    const int src_row = (j & 3) * 8 + col;
    const int src_col = (lane & ~7) + j;
    const double2 pr = make_double2(__shfl_sync(FULLM, m[rj].x, src_row), __shfl_sync(FULLM, m[rj].y, src_row));
    const double2 c0 = make_double2(__shfl_sync(FULLM, m[0].x, src_col), __shfl_sync(FULLM, m[0].y, src_col));
    const double2 c1 = make_double2(__shfl_sync(FULLM, m[1].x, src_col), __shfl_sync(FULLM, m[1].y, src_col));
    const double p = __shfl_sync(FULLM, m[rj].x, (j & 3) * 8 + j);
    pmin = fmin(pmin, p);
    det *= p;
    const double ip = fast_rcp(p);
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const double2 aij = r ? c1 : c0;
      const double tx = fma(aij.x, pr.x, -(aij.y * pr.y));
      const double ty = fma(aij.x, pr.y, aij.y * pr.x);
      double2 v = make_double2(fma(-tx, ip, m[r].x), fma(-ty, ip, m[r].y));
      const bool in_row = (r == rj) && (row == (j & 3));
      if (col == j) v = make_double2(-aij.x * ip, -aij.y * ip);
      if (in_row) v = make_double2(pr.x * ip, pr.y * ip);
      if (in_row && col == j) v = make_double2(ip, 0.0);
      m[r] = v;
    }
  }
  // lambda_min / lambda_max >= 1 / (tr(A) tr(A^-1)) > floor: the reference would not floor anything
  double tinv = (row == col ? m[0].x : 0.0) + (row + 4 == col ? m[1].x : 0.0);
  tinv = warp_sum(tinv);
  // every pivot positive (HPD) and nothing overflowed: det and tr(A^-1) are finite then
  ok = ok && (pmin > 0.0) && (det < 1e300);
  const bool no_floor = ok && isfinite(tinv) && (trs * f * tinv * a.eigenvalue_floor < 0.5);
  if (!__all_sync(0xffffffffu, no_floor)) return false;
  double* __restrict__ blk = a.coef + (size_t)bin * ls_model_doubles(K);
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int sl = em[r] & 255;
    if (em[r] & 256) {
      ls_store_coef(blk, tabE, K, k, sl, m[r].x);
    } else if (em[r] & 512) {
      // slot (d, e) holds 2 Re / -2 Im of B^-1[d][e]; this lane's entry is B^-1[d][e] or its transpose
      const double im = (em[r] & 1024) ? -m[r].y : m[r].y;
      ls_store_coef(blk, tabE, K, k, sl, 2.0 * m[r].x);
      ls_store_coef(blk, tabE, K, k, sl + 1, -2.0 * im);
    }
  }
  const double ldk = log(det);
  if (lane == 0) {
    *ld_out = ldk;
    a.ld[(size_t)bin * 4 + k] = ldk;
    a.ew[(size_t)bin * 4 + k] = Sk[NS];
  }
  return true;
}

template <int K, typename CT>
__global__ void __launch_bounds__(kLsThreads, 2) em_ls_kernel(const PersistArgs a) {
  constexpr int D = 8, NS = 64;
  using SM = LsSmem<K, CT>;
  constexpr int NU = SM::NU, STAGES = SM::STAGES;
  static_assert(STAGES >= 2, "ring too small");
  static_assert(kLsWarps == 8, "update warps add eight partial sums");
  constexpr int RB = 8 * (int)sizeof(CT);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SM& sm = *reinterpret_cast<SM*>(smem_raw);
  // The helper warpgroup takes the LOWEST warp ids: the warp scheduler prefers older / lower warps, and the
  // latency-bound helper streams (producer, update chain) must not queue behind the compute warps.
  const int tid = threadIdx.x, lane = tid & 31;
#if PBB_LS_HELPERS_FIRST
  const int warp = ((tid >> 5) + kLsWarps) % (kLsWarps + kLsHelpers);  // hardware warps 0..3 -> roles 8..11
#else
  const int warp = tid >> 5;
#endif
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
    for (int s = 0; s < STAGES; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 4); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&sm.model_full[s], 1);
      mbar_init(&sm.model_empty[s], kLsWarps);
    }
    mbar_init(&sm.s_full, kLsWarps);
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
        // no more tasks: tell the update warps
        if (n > 0) mbar_wait(&sm.s_empty, (n - 1) & 1u);
        if (lane == 0) {
          if (warp == 0) sm.sdesc[0] = -1;
          mbar_arrive(&sm.s_full);
        }
        break;
      }
      const bool mstep_only = a.first_is_m && it == 0;
      double acc[2 * K], acc8[K], accg[K];
#pragma unroll
      for (int i = 0; i < 2 * K; ++i) acc[i] = 0.0;
#pragma unroll
      for (int k = 0; k < K; ++k) { acc8[k] = 0.0; accg[k] = 0.0; }
      const double2* __restrict__ ce = sm.model[mb].c + (lane >> 4) * SM::HALF;
      const unsigned gbase = n * (unsigned)nchunks;
#pragma unroll 1
      for (int u = warp; u < nunits_all; u += kLsWarps) {
        const unsigned g = gbase + (unsigned)(u >> 2);
        const int st = g % STAGES;
        const int t0 = u * kLsUnit;
        if (t0 < T) {
          mbar_wait(&sm.full[st], (g / STAGES) & 1u);
          PBB_PH(2);  // TMA wait
          const unsigned char* __restrict__ zu =
              reinterpret_cast<const unsigned char*>(&sm.zbuf[st][0][0]) + (u & 3) * kLsUnit * RB;
          double gam[K], cw[K];
          const bool valid = t0 + lane < T;
          if (!mstep_only) {
            ls_e_phase<K, CT>(zu, ce, sm.model[mb].ew, lane, a.aff_eps, gam, cw);
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
            cwb[lane * K + k] = cw[k];
            cwb[(kLsUnit + lane) * K + k] = gam[k];
          }
          __syncwarp();
          PBB_PH(3);  // E phase
          ls_m_phase<K, CT>(zu, cwb, lane, od, oe, acc, acc8, accg);
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
        accg[k] += __shfl_xor_sync(0xffffffffu, accg[k], 1);
        accg[k] += __shfl_xor_sync(0xffffffffu, accg[k], 2);
        accg[k] += __shfl_xor_sync(0xffffffffu, accg[k], 4);
      }
      if (n > 0) mbar_wait(&sm.s_empty, (n - 1) & 1u);  // the previous task's partial sums were consumed
      PBB_PH(5);  // reduce + wait for the partial-sum buffer
#pragma unroll
      for (int k = 0; k < K; ++k) {
        double* __restrict__ sp = sm.Spart[warp][k];
        sp[s_re] = acc[2 * k];
        if (lane < 28) sp[s_re + 1] = acc[2 * k + 1];
        if ((lane & 7) == 0) sp[s_d8] = acc8[k];
        if (lane == 0) sp[NS] = accg[k];
      }
      if (warp == 0 && lane == 0) { sm.sdesc[0] = bin; sm.sdesc[1] = it; sm.sdesc[2] = sm.desc[mb][2]; }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.s_full);
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
        // bins with an all-zero frame keep the reference's own normalisation (em_persistent.cuh): the flag
        // travels with the task so that the update warps do not wait for it
        int dead_bin = 0;
        if (bin >= 0 && a.dead != nullptr && !late_z) dead_bin = __ldcg(a.dead + bin);
        int issued = 0;
        auto issue_chunks = [&](int upto, bool blocking) {
          while (issued < upto) {
            const int st = chunk_cnt % STAGES;
            const uint32_t par = ((chunk_cnt / STAGES) & 1u) ^ 1u;
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
#ifdef PBB_PHASE_TIMING
        long long _pt = clock64();
#define PBB_PP(i) do { if (lane == 0) { long long _t = clock64(); atomicAdd(&a.phase[i], (unsigned long long)(_t - _pt)); _pt = _t; } } while (0)
#else
#define PBB_PP(i) do { } while (0)
#endif
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
        PBB_PP(12);  // ticket taken -> dependency published
        mbar_wait_relaxed(&sm.model_empty[mb], ((n >> 1) & 1u) ^ 1u, 100);  // compute warps are done with task n - 2
        PBB_PP(13);  // wait for the model buffer
        if (late_z && a.dead != nullptr) dead_bin = __ldcg(a.dead + bin);  // written before the arrival flag
        if (bin >= 0 && !mstep_only) {
          // the bin's model block (written by another CTA's update warps, published through the flag): every
          // lane copies 16-byte pieces, all loads in flight before the first store
          constexpr int kPieces = ls_model_doubles(K) / 2;
          constexpr int kRounds = (kPieces + 31) / 32;
          const double2* __restrict__ src = reinterpret_cast<const double2*>(a.coef + (size_t)bin * ls_model_doubles(K));
          double2* __restrict__ dst = reinterpret_cast<double2*>(&sm.model[mb]);
          double2 v[kRounds];
#pragma unroll
          for (int r = 0; r < kRounds; ++r)
            if (lane + 32 * r < kPieces) v[r] = __ldcg(src + lane + 32 * r);
#pragma unroll
          for (int r = 0; r < kRounds; ++r)
            if (lane + 32 * r < kPieces) dst[lane + 32 * r] = v[r];
        }
        if (lane == 0) {
          sm.desc[mb][0] = bin;
          sm.desc[mb][1] = it;
          sm.desc[mb][2] = dead_bin;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.model_full[mb]);
        if (bin < 0) break;
        PBB_PP(14);  // model copy issued
        if (lane == 0) {
          issue_chunks(nchunks, true);
#if PBB_LS_LEAD
          // Optionally hold the next ticket back until the first chunk of this task has been consumed
          // (half-way through for a four-chunk task).
          const unsigned x = chunk_cnt - (unsigned)nchunks;
          mbar_wait_relaxed(&sm.empty[x % STAGES], (x / STAGES) & 1u, 100);
#endif
        }
        __syncwarp();
        PBB_PP(15);  // ring refill of this task
      }
    } else if (hw - 1 < NU) {
      // =============================== update warps ===============================
      const int u = hw - 1;
      // entries of the register Gauss-Jordan this lane owns: (lane >> 3 [+ 4], lane & 7)
      int em[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int i = (lane >> 3) + 4 * r, c = lane & 7;
        if (i == c) {
          em[r] = ls_diag_slot(i) | 256;
        } else {
          bool rev = false;
          const int sl = ls_pair_slot(i, c, &rev);
          // slot (d, e): scatter entry [d][e] = conj(psi(d, e)), entry [e][d] = psi(d, e)  (common.cuh)
          em[r] = sl | (rev ? 1024 : 512);
        }
      }
#ifdef PBB_PHASE_TIMING
      long long _tp = clock64();
#undef PBB_PH
#define PBB_PH(i) do { if (u == 0 && lane == 0) { long long _t = clock64(); atomicAdd(&a.phase[i], (unsigned long long)(_t - _tp)); _tp = _t; } } while (0)
#endif
#pragma unroll 1
      for (unsigned n = 0;; ++n) {
        mbar_wait_relaxed(&sm.s_full, n & 1u, 100);
        PBB_PH(7);  // updater idle
        const int bin = sm.sdesc[0], it = sm.sdesc[1], dead_bin = sm.sdesc[2];
        if (bin < 0) break;
        const bool last_it = it == a.iterations - 1;
        const int par = n & 1;
        // partial sums of the compute warps, fixed order
        for (int k = u; k < K; k += NU) {
          for (int s = lane; s < NS + 1; s += 32) {
            double v[kLsWarps];
#pragma unroll
            for (int w = 0; w < kLsWarps; ++w) v[w] = sm.Spart[w][k][s];
            sm.S[k][s] = ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.s_empty);
        if (last_it) {
          asm volatile("bar.sync 2, %0;" ::"n"(NU * 32) : "memory");
          double* __restrict__ po = a.part + (size_t)bin * K * (NS + 1);
          for (int i = u * 32 + lane; i < K * (NS + 1); i += NU * 32) po[i] = (&sm.S[0][0])[i];
          // S is rewritten by the next task's sum: everybody must be done reading
          asm volatile("bar.sync 2, %0;" ::"n"(NU * 32) : "memory");
        } else {
          const bool plain = a.covariance_norm != PBB_NORM_NONE && dead_bin == 0;
          for (int k = u; k < K; k += NU) {
            PBB_PH(8);
            if (lane == 0) sm.sgam[par][k] = sm.S[k][NS];
            if (!(plain && ls_update_fast<K>(a, bin, k, lane, sm.S[k], em, sm.tabE, &sm.ld[par][k])))
              cacg_update_class<D, false, true>(a, bin, k, K, lane, sm.A[k], sm.V[k], sm.lam[k], sm.S[k], sm.tab,
                                                &sm.ld[par][k], sm.tabE);
            PBB_PH(9);
          }
          __syncwarp();
          asm volatile("bar.sync 2, %0;" ::"n"(NU * 32) : "memory");
          if (u == 0) {
            // weights and ew of the published model (sum of gamma, log det of every class)
            // (ld / sgam of this task parity are not rewritten before the barrier of the task after next)
            if (lane < K) {
              double ldmin = sm.ld[par][0];
#pragma unroll
              for (int j = 1; j < K; ++j) ldmin = fmin(ldmin, sm.ld[par][j]);
              const double wk = a.weight_mode == PBB_WEIGHT_CONST ? 1.0 / K : sm.sgam[par][lane] / (double)T;
              a.coef[(size_t)bin * ls_model_doubles(K) + 4 * SM::HALF + lane] = wk * exp(ldmin - sm.ld[par][lane]);
            }
            __syncwarp();
            if (lane == 0) st_release_gpu(a.flags + bin, it + 1);
          }
        }
        PBB_PH(1);  // updater busy
      }
    }
  }
}

}  // namespace pbb
