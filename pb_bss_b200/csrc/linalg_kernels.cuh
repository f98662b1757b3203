// Batched small-matrix kernels of the beamforming side (pb_bss/extraction):
// Hermitian eigendecomposition, generalised Hermitian eigenproblem (GEV),
// linear solves (MVDR, Souden), blind analytic normalisation, PSD assembly and
// beamformer application.  One warp per D x D problem, matrices in shared memory.
#pragma once
#include "common.cuh"
#include "heig.cuh"

namespace pbb {

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cmulc(double2 a, double2 b) {  // a * conj(b)
  return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__device__ __forceinline__ double2 cdiv(double2 a, double2 b) {
  const double d = b.x * b.x + b.y * b.y;
  return make_double2((a.x * b.x + a.y * b.y) / d, (a.y * b.x - a.x * b.y) / d);
}

// ---- Cholesky B = L L^H, lower triangle in place (warp) ------------------------
// returns false if B is not positive definite (LAPACK zpotrf INFO > 0, which is
// what zhegvd reports as INFO = N + i, get_gev_vector.pyx:130-147)
__device__ inline bool warp_cholesky(double2* __restrict__ B, int D, int lane) {
  bool ok = true;
  for (int j = 0; j < D; ++j) {
    const double djj = B[j * D + j].x;
    ok = ok && (djj > 0.0) && isfinite(djj);
    const double ljj = sqrt(fmax(djj, kTiny));
    __syncwarp();
    for (int i = j + lane; i < D; i += 32) {
      if (i == j) B[j * D + j] = make_double2(ljj, 0.0);
      else { const double2 v = B[i * D + j]; B[i * D + j] = make_double2(v.x / ljj, v.y / ljj); }
    }
    __syncwarp();
    // trailing update: B[i][k] -= L[i][j] conj(L[k][j]) for j < k <= i
    const int n = D - j - 1;
    for (int idx = lane; idx < n * n; idx += 32) {
      const int i = j + 1 + idx / n, k = j + 1 + idx % n;
      if (k <= i) {
        const double2 p = cmulc(B[i * D + j], B[k * D + j]);
        B[i * D + k].x -= p.x;
        B[i * D + k].y -= p.y;
      }
    }
    __syncwarp();
  }
  return ok;
}

// X <- L^{-1} X for lower-triangular L (forward substitution, all columns of X in parallel)
__device__ inline void warp_trsm_lower(const double2* __restrict__ L, double2* __restrict__ X, int D, int ncol,
                                       int lane) {
  for (int c = lane; c < ncol; c += 32) {
    for (int i = 0; i < D; ++i) {
      double2 s = X[i * ncol + c];
      for (int k = 0; k < i; ++k) {
        const double2 p = cmul(L[i * D + k], X[k * ncol + c]);
        s.x -= p.x; s.y -= p.y;
      }
      const double d = L[i * D + i].x;
      X[i * ncol + c] = make_double2(s.x / d, s.y / d);
    }
  }
  __syncwarp();
}

// X <- L^{-H} X (backward substitution with the conjugate transpose of L)
__device__ inline void warp_trsm_lower_h(const double2* __restrict__ L, double2* __restrict__ X, int D, int ncol,
                                         int lane) {
  for (int c = lane; c < ncol; c += 32) {
    for (int i = D - 1; i >= 0; --i) {
      double2 s = X[i * ncol + c];
      for (int k = i + 1; k < D; ++k) {
        const double2 lk = L[k * D + i];  // (L^H)[i][k] = conj(L[k][i])
        const double2 p = cmul(make_double2(lk.x, -lk.y), X[k * ncol + c]);
        s.x -= p.x; s.y -= p.y;
      }
      const double d = L[i * D + i].x;
      X[i * ncol + c] = make_double2(s.x / d, s.y / d);
    }
  }
  __syncwarp();
}

// ---- generalised Hermitian eigenproblem: top eigenvector ---------------------
// scipy.linalg.eigh(a, b) / LAPACK zhegvd ITYPE=1 (beamformer.py:367-411,
// get_gev_vector.pyx:124-150): B = L L^H, C = L^{-1} A L^{-H}, C y = lambda y,
// w = L^{-H} y, so w^H B w = 1.  out (n, D): eigenvector of the LARGEST eigenvalue.
__global__ void gev_kernel(const double2* __restrict__ a, const double2* __restrict__ b, int n, int D,
                           double2* __restrict__ out, int* status, int warps) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x * warps + warp;
  if (m >= n) return;
  const size_t per = ((size_t)3 * D * D * sizeof(double2) + (size_t)((D + 1) / 2) * 6 * sizeof(double) + 15) &
                     ~(size_t)15;
  double2* C = reinterpret_cast<double2*>(smem_raw + per * warp);  // A -> C (Jacobi input)
  double2* V = C + D * D;
  double2* L = V + D * D;
  double* rot = reinterpret_cast<double*>(L + D * D);
  const double2* __restrict__ am = a + (size_t)m * D * D;
  const double2* __restrict__ bm = b + (size_t)m * D * D;
  bool bad = false;
  // zhegvd reads the lower triangles (UPLO = 'L'); use the Hermitian parts
  for (int i = lane; i < D * D; i += 32) {
    const int r = i / D, c = i - r * D;
    const double2 x = am[r * D + c], y = am[c * D + r];
    const double2 p = bm[r * D + c], q = bm[c * D + r];
    C[i] = make_double2(0.5 * (x.x + y.x), r == c ? 0.0 : 0.5 * (x.y - y.y));
    L[i] = make_double2(0.5 * (p.x + q.x), r == c ? 0.0 : 0.5 * (p.y - q.y));
    bad |= !isfinite(C[i].x) || !isfinite(C[i].y) || !isfinite(L[i].x) || !isfinite(L[i].y);
  }
  __syncwarp();
  const bool pd = warp_cholesky(L, D, lane);
  // C <- L^{-1} C L^{-H}:  first C <- L^{-1} C (columns), then C <- (L^{-1} C^H)^H
  warp_trsm_lower(L, C, D, D, lane);
  for (int i = lane; i < D * D; i += 32) {  // conjugate transpose into V (scratch)
    const int r = i / D, c = i - r * D;
    const double2 v = C[c * D + r];
    V[i] = make_double2(v.x, -v.y);
  }
  __syncwarp();
  warp_trsm_lower(L, V, D, D, lane);
  for (int i = lane; i < D * D; i += 32) {  // hermitise the result back into C
    const int r = i / D, c = i - r * D;
    const double2 u = V[c * D + r], v = V[r * D + c];  // conj(V^T) and V agree up to rounding
    C[i] = make_double2(0.5 * (u.x + v.x), r == c ? 0.0 : 0.5 * (-u.y + v.y));
  }
  __syncwarp();
  const int sweeps = warp_jacobi_any(C, V, rot, D, lane);
  int best = 0;
  double lmax = C[0].x;
  for (int d = 1; d < D; ++d) {
    const double l = C[d * D + d].x;
    if (l >= lmax) { lmax = l; best = d; }
  }
  // w = L^{-H} y
  double2* yv = C;                                // reuse C's first column block as the rhs (D x 1)
  __syncwarp();
  for (int d = lane; d < D; d += 32) yv[d] = V[d * D + best];
  __syncwarp();
  warp_trsm_lower_h(L, yv, D, 1, lane);
  for (int d = lane; d < D; d += 32) out[(size_t)m * D + d] = yv[d];
  if ((!pd || __any_sync(0xffffffffu, bad) || sweeps > kJacobiMaxSweeps) && lane == 0 && status)
    atomicMax(status, m + 1);
}

// shared memory of one warp of solve_kernel: A, X and -- when the minimum-norm fallback is available (D <= kLstsqMaxD)
// -- a second matrix, the eigenvectors, a right-hand-side scratch and the Jacobi rotations
constexpr int kLstsqMaxD = 40;
__host__ __device__ inline size_t solve_smem_per_warp(int D, int R) {
  size_t b = (size_t)(D * D + D * R) * sizeof(double2);
  if (D <= kLstsqMaxD) b += (size_t)(2 * D * D + D * R) * sizeof(double2) + (size_t)((D + 1) / 2) * 6 * sizeof(double);
  return (b + 15) & ~(size_t)15;
}

// ---- general complex solve A X = B with partial pivoting (np.linalg.solve / zgesv) ----
// A (n, D, D), B (n, D, R) -> X (n, D, R).  hermitize: use (A + A^H) / 2 (beamformer.py:246-248).
// An exactly singular system (zero pivot: LinAlgError in the reference) takes the reference's fallback,
// np.linalg.lstsq (beamformer.py:251-256, math/solve.py:95-114): the minimum-norm solution X = A^+ B.  A Hermitian A
// (the PSD matrices of this path) is pseudo-inverted through its eigendecomposition, eigenvalues below
// eps * D * max|lambda| count as zero like LAPACK's rcond; a non-Hermitian A through A^H A (singular values below
// ~1e-7 of the largest count as zero there).  An
// all-zero A gives X = 0 (test_beamformer.py:211-376).  Non-finite input propagates as NaN, as it does in LAPACK.
// status (may be null) is only set when the fallback is unavailable (D > kLstsqMaxD).
__global__ void solve_kernel(const double2* __restrict__ a, const double2* __restrict__ b, int n, int D, int R,
                             int hermitize, double2* __restrict__ x, int* status, int warps) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x * warps + warp;
  if (m >= n) return;
  const size_t per = solve_smem_per_warp(D, R);
  double2* A = reinterpret_cast<double2*>(smem_raw + per * warp);
  double2* X = A + D * D;
  const double2* __restrict__ am = a + (size_t)m * D * D;
  for (int i = lane; i < D * D; i += 32) {
    const int r = i / D, c = i - r * D;
    const double2 u = am[i];
    if (hermitize) {
      const double2 v = am[c * D + r];
      A[i] = make_double2(0.5 * (u.x + v.x), 0.5 * (u.y - v.y));
    } else {
      A[i] = u;
    }
  }
  for (int i = lane; i < D * R; i += 32) X[i] = b[(size_t)m * D * R + i];
  __syncwarp();
  bool singular = false, nonfinite = false;
  for (int j = 0; j < D; ++j) {
    // pivot search (every lane redundantly: D is tiny)
    int piv = j;
    double best = -1.0;
    for (int i = j; i < D; ++i) {
      const double2 v = A[i * D + j];
      const double mag = fabs(v.x) + fabs(v.y);  // LAPACK izamax uses |re| + |im|
      if (mag > best) { best = mag; piv = i; }
    }
    if (!isfinite(best)) { nonfinite = true; break; }
    if (!(best > 0.0)) { singular = true; break; }
    if (piv != j) {
      for (int c = lane; c < D; c += 32) { const double2 t = A[j * D + c]; A[j * D + c] = A[piv * D + c]; A[piv * D + c] = t; }
      for (int c = lane; c < R; c += 32) { const double2 t = X[j * R + c]; X[j * R + c] = X[piv * R + c]; X[piv * R + c] = t; }
    }
    __syncwarp();
    const double2 p = A[j * D + j];
    // eliminate below
    for (int idx = lane; idx < (D - j - 1) * (D - j - 1 + R); idx += 32) {
      const int w = D - j - 1 + R;
      const int i = j + 1 + idx / w, cc = idx % w;
      const double2 f = cdiv(A[i * D + j], p);
      if (cc < D - j - 1) {
        const int c = j + 1 + cc;
        const double2 q = cmul(f, A[j * D + c]);
        A[i * D + c].x -= q.x; A[i * D + c].y -= q.y;
      } else {
        const int c = cc - (D - j - 1);
        const double2 q = cmul(f, X[j * R + c]);
        X[i * R + c].x -= q.x; X[i * R + c].y -= q.y;
      }
    }
    __syncwarp();
  }
  if (nonfinite) {
    for (int i = lane; i < D * R; i += 32) X[i] = make_double2(NAN, NAN);
    __syncwarp();
  } else if (!singular) {
    for (int c = lane; c < R; c += 32) {  // back substitution, columns in parallel
      for (int i = D - 1; i >= 0; --i) {
        double2 s = X[i * R + c];
        for (int k = i + 1; k < D; ++k) {
          const double2 q = cmul(A[i * D + k], X[k * R + c]);
          s.x -= q.x; s.y -= q.y;
        }
        X[i * R + c] = cdiv(s, A[i * D + i]);
      }
    }
    __syncwarp();
  } else if (D > kLstsqMaxD) {
    if (lane == 0 && status) atomicMax(status, m + 1);
  } else {
    // ---- minimum-norm least squares (np.linalg.lstsq) ----
    double2* G = X + D * R;          // matrix to diagonalise
    double2* V = G + D * D;          // its eigenvectors
    double2* Y = V + D * D;          // right-hand side in the eigenbasis
    double* rot = reinterpret_cast<double*>(Y + D * R);
    double asym = 0.0, amax = 0.0;
    for (int i = lane; i < D * D; i += 32) {
      const int r = i / D, c = i - r * D;
      double2 u = am[i];
      const double2 v = am[c * D + r];
      if (hermitize) u = make_double2(0.5 * (u.x + v.x), 0.5 * (u.y - v.y));
      A[i] = u;
      asym = fmax(asym, hermitize ? 0.0 : fabs(u.x - v.x) + fabs(u.y + v.y));
      amax = fmax(amax, fabs(u.x) + fabs(u.y));
    }
    for (int i = lane; i < D * R; i += 32) X[i] = b[(size_t)m * D * R + i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      asym = fmax(asym, __shfl_xor_sync(0xffffffffu, asym, o));
      amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    }
    __syncwarp();
    const bool herm = asym <= 1e-14 * amax;
    // G = A (Hermitian) or A^H A; Y0 = B or A^H B
    for (int i = lane; i < D * D; i += 32) {
      const int r = i / D, c = i - r * D;
      double2 g = A[i];
      if (!herm) {
        g = make_double2(0.0, 0.0);
        for (int k = 0; k < D; ++k) {
          const double2 q = cmulc(A[k * D + c], A[k * D + r]);  // conj(A[k][r]) * A[k][c]
          g.x += q.x; g.y += q.y;
        }
      }
      G[i] = g;
    }
    for (int i = lane; i < D * R; i += 32) {
      const int r = i / R, c = i - r * R;
      double2 y = X[i];
      if (!herm) {
        y = make_double2(0.0, 0.0);
        for (int k = 0; k < D; ++k) {
          const double2 q = cmulc(X[k * R + c], A[k * D + r]);  // conj(A[k][r]) * B[k][c]
          y.x += q.x; y.y += q.y;
        }
      }
      Y[i] = y;
    }
    __syncwarp();
    warp_jacobi_any(G, V, rot, D, lane);
    __syncwarp();
    double lmax = 0.0;
    for (int i = 0; i < D; ++i) lmax = fmax(lmax, fabs(G[i * D + i].x));
    const double cut = (herm ? 1.0 : 8.0) * DBL_EPSILON * D * lmax;
    // X = V diag(1 / lambda) V^H Y0 over the eigenvalues above the cut-off
    for (int i = lane; i < D * R; i += 32) {  // T = diag^+ V^H Y0, stored in X
      const int e = i / R, c = i - e * R;
      const double l = G[e * D + e].x;
      double2 t = make_double2(0.0, 0.0);
      if (fabs(l) > cut) {
        for (int k = 0; k < D; ++k) {
          const double2 q = cmulc(Y[k * R + c], V[k * D + e]);  // conj(V[k][e]) * Y0[k][c]
          t.x += q.x; t.y += q.y;
        }
        t.x /= l; t.y /= l;
      }
      X[i] = t;
    }
    __syncwarp();
    for (int i = lane; i < D * R; i += 32) {
      const int r = i / R, c = i - r * R;
      double2 o = make_double2(0.0, 0.0);
      for (int e = 0; e < D; ++e) {
        const double2 q = cmul(V[r * D + e], X[e * R + c]);
        o.x += q.x; o.y += q.y;
      }
      Y[i] = o;
    }
    __syncwarp();
    for (int i = lane; i < D * R; i += 32) X[i] = Y[i];
    __syncwarp();
  }
  for (int i = lane; i < D * R; i += 32) x[(size_t)m * D * R + i] = X[i];
}

// ---- MVDR: w = N^{-1} a / (a^H N^{-1} a) given x = N^{-1} a (beamformer.py:257-258) ----
__global__ void mvdr_scale_kernel(const double2* __restrict__ atf, const double2* __restrict__ x, int n, int D,
                                  double2* __restrict__ w) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n) return;
  double2 den = make_double2(0.0, 0.0);
  for (int d = 0; d < D; ++d) {
    const double2 a = atf[(size_t)m * D + d], v = x[(size_t)m * D + d];
    const double2 p = cmul(make_double2(a.x, -a.y), v);
    den.x += p.x; den.y += p.y;
  }
  for (int d = 0; d < D; ++d) w[(size_t)m * D + d] = cdiv(x[(size_t)m * D + d], den);
}

// ---- Souden MVDR pieces (beamformer.py:601-698) --------------------------------
// mat = phi / max(trace(phi).real, eps); per-bin SNR numerators / denominators for
// every candidate reference channel R: w_R = mat[:, R]
__global__ void souden_kernel(const double2* __restrict__ phi, const double2* __restrict__ target,
                              const double2* __restrict__ noise, int n, int D, double eps, double2* __restrict__ mat,
                              double2* __restrict__ num, double2* __restrict__ den) {
  const int m = blockIdx.x;
  const int R = threadIdx.x;
  if (R >= D) return;
  const double2* __restrict__ ph = phi + (size_t)m * D * D;
  double tr = 0.0;
  for (int d = 0; d < D; ++d) tr += ph[d * D + d].x;
  const double s = 1.0 / fmax(tr, eps);
  for (int d = 0; d < D; ++d) {
    const double2 v = ph[d * D + R];
    mat[(size_t)m * D * D + d * D + R] = make_double2(v.x * s, v.y * s);
  }
  // quadratic forms w^H T w and w^H N w with w = mat[:, R]
  double2 qt = make_double2(0.0, 0.0), qn = make_double2(0.0, 0.0);
  for (int d = 0; d < D; ++d) {
    const double2 wd = make_double2(ph[d * D + R].x * s, -ph[d * D + R].y * s);  // conj(w_d)
    double2 tt = make_double2(0.0, 0.0), nn = make_double2(0.0, 0.0);
    for (int e = 0; e < D; ++e) {
      const double2 we = make_double2(ph[e * D + R].x * s, ph[e * D + R].y * s);
      const double2 a = cmul(target[(size_t)m * D * D + d * D + e], we);
      const double2 b = cmul(noise[(size_t)m * D * D + d * D + e], we);
      tt.x += a.x; tt.y += a.y; nn.x += b.x; nn.y += b.y;
    }
    const double2 a = cmul(wd, tt), b = cmul(wd, nn);
    qt.x += a.x; qt.y += a.y; qn.x += b.x; qn.y += b.y;
  }
  num[(size_t)m * D + R] = qt;
  den[(size_t)m * D + R] = qn;
}

// ---- blind analytic normalisation (beamformer.py:459-488) -----------------------
__global__ void ban_kernel(const double2* __restrict__ vec, const double2* __restrict__ noise, int n, int D,
                           double2* __restrict__ out) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n) return;
  const double2* __restrict__ N = noise + (size_t)m * D * D;
  const double2* __restrict__ w = vec + (size_t)m * D;
  // u = N w ; nominator = sqrt(w^H N N w) = sqrt((N^H w)^H (N w)) ; denominator = |w^H N w|
  double2 nom = make_double2(0.0, 0.0), den = make_double2(0.0, 0.0);
  for (int a = 0; a < D; ++a) {
    double2 left = make_double2(0.0, 0.0);   // sum_x conj(w_x) N[x][a]
    double2 right = make_double2(0.0, 0.0);  // sum_c N[a][c] w_c
    for (int c = 0; c < D; ++c) {
      const double2 p = cmul(make_double2(w[c].x, -w[c].y), N[c * D + a]);
      left.x += p.x; left.y += p.y;
      const double2 q = cmul(N[a * D + c], w[c]);
      right.x += q.x; right.y += q.y;
    }
    const double2 p = cmul(left, right);
    nom.x += p.x; nom.y += p.y;
    const double2 q = cmul(make_double2(w[a].x, -w[a].y), right);
    den.x += q.x; den.y += q.y;
  }
  // complex sqrt of nom, |den|, then |nom_sqrt / den_abs|
  const double nmag = sqrt(sqrt(nom.x * nom.x + nom.y * nom.y));  // |sqrt(z)| = sqrt(|z|)
  const double dmag = sqrt(den.x * den.x + den.y * den.y);        // sqrt(den * conj(den))
  const double scale = dmag != 0.0 ? nmag / dmag : 0.0;
  for (int d = 0; d < D; ++d) out[(size_t)m * D + d] = make_double2(w[d].x * scale, w[d].y * scale);
}

// ---- apply a beamforming vector: out[b][f][t] = sum_d conj(w[b][f][d]) Y[f][d][t] (beamformer.py:572-583);
// blockIdx.z = b runs over beamformers that share one mix (K sources on one STFT), 1 otherwise
template <typename CT>
__global__ void apply_bf_kernel(const double2* __restrict__ w, const CT* __restrict__ Y, int F, int D, int T,
                                double2* __restrict__ out) {
  const int f = blockIdx.y;
  const size_t bf = (size_t)blockIdx.z * F + f;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  double2 s = make_double2(0.0, 0.0);
  for (int d = 0; d < D; ++d) {
    const double2 wd = w[bf * D + d];
    const double2 y = ld_cplx(Y + ((size_t)f * D + d) * T + t);
    s.x += wd.x * y.x + wd.y * y.y;
    s.y += wd.x * y.y - wd.y * y.x;
  }
  out[bf * T + t] = s;
}

// ---- rank-1 PSD approximations (beamformer_wrapper.py:11-69) ---------------------
// out = a a^H * trace(cov) / trace(a a^H)
__global__ void rank_one_kernel(const double2* __restrict__ a, const double2* __restrict__ cov, int n, int D,
                                double2* __restrict__ out) {
  const int m = blockIdx.x;
  if (m >= n) return;
  const double2* __restrict__ am = a + (size_t)m * D;
  double2 tr = make_double2(0.0, 0.0);
  double na = 0.0;
  for (int d = 0; d < D; ++d) {
    const double2 c = cov[(size_t)m * D * D + d * D + d];
    tr.x += c.x; tr.y += c.y;
    na += am[d].x * am[d].x + am[d].y * am[d].y;
  }
  const double2 scale = make_double2(tr.x / na, tr.y / na);
  for (int i = threadIdx.x; i < D * D; i += blockDim.x) {
    const int d = i / D, e = i - d * D;
    out[(size_t)m * D * D + i] = cmul(scale, cmulc(am[d], am[e]));
  }
}

// y = M x per matrix (the "scaled GEV ATF" Phi_nn w, beamformer_wrapper.py:27-46)
__global__ void matvec_kernel(const double2* __restrict__ M, const double2* __restrict__ x, int n, int D,
                              double2* __restrict__ y) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n) return;
  for (int d = 0; d < D; ++d) {
    double2 s = make_double2(0.0, 0.0);
    for (int e = 0; e < D; ++e) {
      const double2 p = cmul(M[(size_t)m * D * D + d * D + e], x[(size_t)m * D + e]);
      s.x += p.x; s.y += p.y;
    }
    y[(size_t)m * D + d] = s;
  }
}

// ---- PSD assembly from the slot sums of the M-step kernels (beamformer.py:59-160) ----
// part (F, NCH, K, NS + 1) -> psd (F, K, D, D); scale: 0 = none, 1 = 1 / max(sum mask, 1e-10)
// (beamformer.py:127-131), 2 = 1 / T (no mask, beamformer.py:114-117)
__global__ void psd_finalize_kernel(const double* __restrict__ part, int nch, int F, int K, int D, int T, int scale,
                                    double2* __restrict__ psd) {
  const int f = blockIdx.x, k = blockIdx.y;
  const int NS = D * D;
  const double* __restrict__ p0 = part + ((size_t)f * nch * K + k) * (NS + 1);
  double sm = 0.0;
  for (int c = 0; c < nch; ++c) sm += p0[(size_t)c * K * (NS + 1) + NS];
  const double sc = scale == 1 ? 1.0 / fmax(sm, 1e-10) : (scale == 2 ? 1.0 / (double)T : 1.0);
  double2* __restrict__ o = psd + ((size_t)f * K + k) * NS;
  double* od = reinterpret_cast<double*>(o);
  for (int s = threadIdx.x; s < NS; s += blockDim.x) {
    double v = 0.0;
    for (int c = 0; c < nch; ++c) v += p0[(size_t)c * K * (NS + 1) + s];
    v *= sc;
    const SlotInfo si = slot_info(D, s);
    if (si.kind == 0) { od[2 * (si.d * D + si.d)] = v; od[2 * (si.d * D + si.d) + 1] = 0.0; }
    else if (si.kind == 1) { od[2 * (si.d * D + si.e)] = v; od[2 * (si.e * D + si.d)] = v; }
    else { od[2 * (si.d * D + si.e) + 1] = -v; od[2 * (si.e * D + si.d) + 1] = v; }
  }
}

}  // namespace pbb
