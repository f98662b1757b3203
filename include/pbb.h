/* pb_bss_b200 C ABI -- the drop-in boundary of the pb_bss EM / beamforming hot path.
 *
 * pb_bss (the reference) is a pure NumPy library; its only native seam is the
 * import-time hook around two Cython LAPACK loops
 * (pb_bss/extraction/beamformer.py:38-56 ->
 *  pb_bss/extraction/cythonized/get_gev_vector.pyx:42,
 *  pb_bss/extraction/cythonized/c_eig.pyx:14).  This library moves the whole
 * per-bin hot path behind one C ABI; each entry point names the reference
 * function(s) it replaces.  A maintainer binds it with ctypes exactly like
 * pb_bss_b200/_lib.py does.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller and the library never frees
 *    caller memory (pbb_cacgmm_fit additionally accepts PINNED host pointers, see there).
 *    The library itself owns three small things per device, created on first use and kept
 *    for the life of the process: a non-blocking side stream with two events (streamed
 *    upload of pbb_cacgmm_fit), a cache of at most 16 task-order tables (4 bytes per task,
 *    streamed upload only) and the occupancy numbers of its persistent kernels;
 *  - `stream` is a cudaStream_t passed as void*; all work is stream-ordered
 *    and asynchronous, no entry synchronises the device or the stream;
 *  - limits: D < 35, K < 20 (the reference asserts the same, cacgmm.py:197,249-250);
 *    pbb_streamed_task_order packs the bin into 16 bits (F <= 65535; pbb_cacgmm_fit only uses
 *    such a table for F <= 4096 and schedules larger problems without it);
 *  - thread safety: entries may be called concurrently from several host threads on
 *    different streams.  pbb_last_error() is thread-local.  Two streamed fits (pinned-host
 *    input) on the SAME device are serialised while they enqueue, because they share the
 *    side stream; everything else only reads library state;
 *  - return value: 0 = ok, -i = argument i (1-based) invalid (LAPACK INFO<0
 *    convention, cf. get_gev_vector.pyx:130-147), > 0 = CUDA runtime error
 *    code; pbb_last_error() gives the message (thread-local);
 *  - numerical failures (non-finite covariance, not-positive-definite noise
 *    PSD, ...) are reported through a caller-provided device status word
 *    `int* status` (0 = ok, else 1 + index of the first failing matrix/bin),
 *    which the caller reads after synchronising -- the reference raises
 *    AssertionError / ValueError at the same places;
 *  - all arithmetic is IEEE fp64; `dtype` only selects the STORAGE type of the
 *    complex observation (PBB_C64 = float2, PBB_C128 = double2).
 */
#ifndef PBB_H_
#define PBB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PBB_C64 0
#define PBB_C128 1

/* covariance_norm of CACGMMTrainer.fit (pb_bss/distribution/cacgmm.py:152) */
#define PBB_NORM_NONE 0       /* covariance_norm=False */
#define PBB_NORM_EIGENVALUE 1 /* 'eigenvalue' (default) */
#define PBB_NORM_TRACE 2      /* 'trace' */

/* weight_constant_axis (pb_bss/distribution/mixture_model_utils.py:133-203) */
#define PBB_WEIGHT_TIME 0     /* (-1,): one weight per (bin, class) */
#define PBB_WEIGHT_CONST 1    /* -2: constant 1/K */
#define PBB_WEIGHT_TIED_TIME 2 /* (-3,): frequency-tied weights, one per (class, frame): array (K, T) */
#define PBB_WEIGHT_TIED 3     /* (-3, -1): frequency-tied, one per class: array (K) */

const char* pbb_last_error(void);
int pbb_version(void);

/* Launch accounting and optional CUDA-event timing of the library's own kernels
 * (used by bench.py for `gpu_launches` and the roofline of the dominant kernel). */
long long pbb_launch_count(void);
void pbb_profile_enable(int on);
void pbb_profile_reset(void);
/* Sums the recorded launches per kernel, returns the kernel with the largest
 * total device time (ms) and its launch count, and clears the records.
 * Synchronises on the recorded events.  Return value: number of distinct kernels. */
int pbb_profile_dominant(char* name, int name_len, double* total_ms, int* launches);
void pbb_profile_dump(void); /* every recorded launch to stderr */

/* ------------------------------------------------------------------------
 * Observation normalisation.
 * swap=1: pb_bss/distribution/complex_angular_central_gaussian.py:34-55
 *         (unit norm over D, zero vectors stay zero, output (F, D, T));
 * swap=0: pb_bss/distribution/complex_watson.py:16-29 (output (F, T, D)).
 * y: (F, T, D) complex of `dtype`; z: same dtype. */
int pbb_normalize_observation(const void* y, void* z, int F, int T, int D,
                              int dtype, int swap, void* stream);

/* ------------------------------------------------------------------------
 * cACGMM (pb_bss/distribution/cacgmm.py).
 *
 * Device model of one fit: for every (bin f, class k)
 *   eigenvectors (F, K, D, D) complex128 row-major, column e = e-th vector
 *   eigenvalues  (F, K, D)    float64, ascending
 *   weight       (F, K)       float64
 * = CACGMM.weight / cacg.covariance_eigenvectors / covariance_eigenvalues
 * (cacgmm.py:58-62, complex_angular_central_gaussian.py:78-79).
 */
typedef struct pbb_cacgmm_options {
  int iterations;          /* > 0 */
  int covariance_norm;     /* PBB_NORM_* */
  int weight_mode;         /* PBB_WEIGHT_* */
  int hermitize;           /* accepted for API parity; the scatter matrix is
                              accumulated in Hermitian form either way */
  double affiliation_eps;  /* clip of the posterior, cacgmm.py:154 */
  double eigenvalue_floor; /* cacgmm.py:155 */
  int frames_per_block;    /* 0 = library default; tuning knob */
  int reserved;            /* bit 0: force the multi-kernel (non-persistent) path;
                              bit 1: no streamed upload (see pbb_cacgmm_fit) */
} pbb_cacgmm_options;

/* Host-only helper (no GPU needed): the task order pbb_cacgmm_fit uses for a streamed upload.
 * order (HOST, F * iterations ints): order[ticket] = bin | iteration << 16.  `arrive` bins join
 * per time slot, at most `cap` tasks run per slot; every (bin, it) comes after (bin, it - 1). */
int pbb_streamed_task_order(int F, int iterations, int arrive, int cap, int* order);

/* Host only: which persistent kernel pbb_cacgmm_fit runs for a device-resident problem on a GPU with `sms` SMs, and
 * how one EM iteration of a bin is split.  lean = no saliency / activity mask / user-supplied model
 * and eigenvalue_floor in the product-softmax range; streamed = pinned host input.
 * *kernel: 0 = em_ws_kernel (task kernel, D = 8), 1 = em_sticky_kernel (one cluster of *split CTAs per bin for the
 * whole fit), 2 = em_persistent_kernel (D = 4 / 6, full variant); *split = parts per bin-iteration (1 = none).  The
 * environment overrides of the library (PBB_TSPLIT, PBB_STICKY, PBB_EM_KERNEL) are not applied here. */
int pbb_em_dispatch(int F, int T, int D, int K, int lean, int streamed, int sms, int* kernel, int* split);

/* Bytes of scratch pbb_cacgmm_fit / _predict need for this problem size. */
size_t pbb_cacgmm_workspace_bytes(int F, int T, int D, int K);

/* CACGMMTrainer.fit (cacgmm.py:142-280): full EM loop on the device.
 *  y            (F, T, D) complex `dtype`, un-normalised STFT
 *  init_aff     (F, K, T) float64 initial affiliations, or NULL for a warm
 *               start from the model already stored in
 *               eigenvectors/eigenvalues/weight (cacgmm.py:229-234)
 *  saliency     (F, T) float64 or NULL
 *  activity     (F, K, T) uint8 source_activity_mask or NULL
 *  outputs      eigenvectors, eigenvalues, weight as described above
 *  status       device int, see header comment
 * Host buffers: y, init_aff and the three outputs may also be PINNED
 * (page-locked, mapped) host memory; the kernels then read / write them in place
 * over PCIe.  With y pinned and init_aff given, a small loader kernel on an
 * internal side stream streams the bins in while the EM kernel already iterates
 * on the bins that have arrived (task order: see em_persistent.cuh; the order
 * table, 4 bytes per task, lives in a small library-owned device cache).  The
 * call stays asynchronous with respect to `stream`.
 */
int pbb_cacgmm_fit(const void* y, int dtype, int F, int T, int D, int K,
                   const double* init_aff, const double* saliency,
                   const uint8_t* activity, const pbb_cacgmm_options* opt,
                   void* eigenvectors, double* eigenvalues, double* weight,
                   void* workspace, size_t workspace_bytes, int* status,
                   void* stream);

/* CACGMM.predict / _predict / log_likelihood (cacgmm.py:64-138): one E-step.
 *  affiliation  (F, K, T) float64 out (may be NULL)
 *  quadratic    (F, K, T) float64 out (may be NULL)
 *  loglik       (F) float64 out, per-bin sum_t logsumexp_k log_pdf (may be NULL)
 *  affiliation_eps: 0 for predict (cacgmm.py:73)
 *  weight / weight_mode: (F, K) for PBB_WEIGHT_TIME, ignored for _CONST, (K, T) for
 *  _TIED_TIME and (K) for _TIED (frequency-tied weights, mixture_model_utils.py:187-190)
 */
int pbb_cacgmm_predict(const void* y, int dtype, int F, int T, int D, int K,
                       const void* eigenvectors, const double* eigenvalues,
                       const double* weight, int weight_mode,
                       const uint8_t* activity, double affiliation_eps,
                       double* affiliation, double* quadratic, double* loglik,
                       void* workspace, size_t workspace_bytes, int* status,
                       void* stream);

/* One M-step from given affiliations and quadratic forms:
 * CACGMMTrainer._m_step (cacgmm.py:315-345) =
 * estimate_mixture_weight + ComplexAngularCentralGaussianTrainer._fit
 * (complex_angular_central_gaussian.py:253-342) + from_covariance (:81-132).
 *  quadratic may be NULL (= ones, the first iteration, cacgmm.py:210). */
int pbb_cacgmm_mstep(const void* y, int dtype, int F, int T, int D, int K,
                     const double* affiliation, const double* quadratic,
                     const double* saliency, const pbb_cacgmm_options* opt,
                     void* eigenvectors, double* eigenvalues, double* weight,
                     void* workspace, size_t workspace_bytes, int* status,
                     void* stream);

/* estimate_mixture_weight with weight_constant_axis=(-3,) / (-3, -1)
 * (mixture_model_utils.py:133-203): weight_kt[k][t] = mean over bins of
 * affiliation[f][k][t]; flags bit 0: additionally weight_k[k] = mean over t; flags bit 1:
 * the saliency form (:192-203, what CWMMTrainer uses, cwmm.py:129-130): the result is
 * L1-normalised over the classes (zero norm -> 1e-10). */
int pbb_mixture_weight_over_bins(const double* affiliation, int F, int K, int T,
                                 int flags, double* weight_kt,
                                 double* weight_k, void* stream);

/* ------------------------------------------------------------------------
 * Integrated spatial + spectral model (pb_bss/distribution/gcacgmm.py): cACG of the observation combined with a
 * Gaussian over per-(bin, frame) embeddings (F, T, E).  The spatial part reuses pbb_cacgmm_predict (quadratic form)
 * and pbb_cacgmm_mstep; these entries are the pieces around them.
 * ------------------------------------------------------------------------ */

/* ComplexAngularCentralGaussian._log_pdf (cacg.py:198-201) from the quadratic form: log_pdf = -D log max(|q|, tiny)
 * - sum_d log eigenvalues.  quadratic, log_pdf (F, K, T); eigenvalues (F, K, D). */
int pbb_cacg_log_pdf(const double* quadratic, const double* eigenvalues, int F, int K,
                     int T, int D, double* log_pdf, void* stream);

/* DiagonalGaussian / SphericalGaussian.log_pdf (gaussian.py:57-135) of every embedding under every class:
 * embedding (F, T, E), mean / precision_cholesky (K, E) (spherical: the scalar repeated E times), log_det (K)
 * -> log_pdf (F, K, T).  E <= 64.  diagonal != 0 evaluates the reference's DiagonalGaussian expression, whose einsum
 * '...dD,...nD->...nd' (gaussian.py:79-87) contracts precision_cholesky[d][:] of EVERY class d with the centred
 * observation and sums the squares over d -- reproduced as is, because the model is a drop-in.  diagonal == 2:
 * VonMisesFisher.log_pdf (von_mises_fisher.py:66-81) for the vMF + cACG model: precision_cholesky[k][0] carries the
 * concentration, log_det[k] the log normaliser; out = concentration <mean, x / max(||x||, tiny)> - log_norm. */
int pbb_gaussian_log_pdf(const double* embedding, const double* mean,
                         const double* precision_cholesky, const double* log_det,
                         int F, int T, int E, int K, int diagonal, double* log_pdf,
                         void* stream);

/* GaussianTrainer._fit (gaussian.py:155-193) over the F*T embeddings with weights weight (F, K, T): mean (K, E),
 * covariance (K, E) ('diagonal') or (K) ('spherical'); two passes (mean, then centred second moments), fixed
 * summation order.  scratch: pbb_gaussian_fit_scratch_doubles doubles. */
size_t pbb_gaussian_fit_scratch_doubles(int F, int E, int K);
int pbb_gaussian_fit(const double* embedding, const double* weight, int F, int T, int E,
                     int K, int spherical, double* mean, double* covariance,
                     double* scratch, void* stream);

/* log_pdf_to_affiliation (mixture_model_utils.py:7-55) of scale_a * log_pdf_a + scale_b * log_pdf_b (log_pdf_b may be
 * null) with the weight layouts of pbb_cacgmm_predict; inline_pa != 0:
 * log_pdf_to_affiliation_for_integration_models_with_inline_pa (:58-130) -- per bin the classes of log_pdf_a are
 * re-paired with those of log_pdf_b by the first permutation (itertools order) that maximises the auxiliary function;
 * permutation (F, K) int32 (may be null) receives the choice.  K <= 6. */
int pbb_log_pdf_to_affiliation(const double* log_pdf_a, const double* log_pdf_b,
                               double scale_a, double scale_b, const double* weight,
                               int weight_mode, const uint8_t* activity,
                               double affiliation_eps, int inline_pa, int F, int K, int T,
                               double* affiliation, int* permutation, void* stream);

/* Per-bin class weights of the integrated models (gcacgmm.py:286-291, weight_constant_axis (-1,)):
 * weight[f][k] = sum_t m[f][k][t] / sum_k sum_t m[f][k][t]. */
int pbb_class_weight(const double* masked_affiliation, int F, int K, int T, double* weight,
                     void* stream);

/* ------------------------------------------------------------------------
 * Complex Watson mixture model (pb_bss/distribution/cwmm.py, complex_watson.py).
 *
 * Device model: mode (F, K, D) complex128, concentration (F, K), weight (F, K)
 * = CWMM.weight / complex_watson.mode / .concentration (cwmm.py:21-24).
 * The inverse hypergeometric ratio is the quadratic B-spline of
 * ComplexWatsonTrainer.spline (complex_watson.py:237-256); the caller builds
 * it once with the reference's recipe and passes its knots spline_t[n + 3]
 * and coefficients spline_c[n] (device pointers); it is model state, like the
 * reference's cached_property. */
size_t pbb_cwmm_workspace_bytes(int F, int T, int D, int K);

/* CWMMTrainer.fit / _fit / _m_step (cwmm.py:76-240), affiliation_eps = 0.
 * init_aff (F, K, T) is required (cwmm.py:121-127 draws it on the host). */
int pbb_cwmm_fit(const void* y, int dtype, int F, int T, int D, int K,
                 const double* init_aff, const double* saliency,
                 int iterations, int weight_mode, const double* spline_t,
                 const double* spline_c, int spline_n,
                 double max_concentration, void* mode, double* concentration,
                 double* weight, void* workspace, size_t workspace_bytes,
                 int* status, void* stream);

/* CWMM.predict (cwmm.py:26-52): affiliation (F, K, T) out.  weight: (F, K) for
 * PBB_WEIGHT_TIME, ignored (1/K) for PBB_WEIGHT_CONST, (K, T) / (K) for the
 * frequency-tied PBB_WEIGHT_TIED_TIME / PBB_WEIGHT_TIED (weight_constant_axis (-3,) / (-3, -1)). */
int pbb_cwmm_predict(const void* y, int dtype, int F, int T, int D, int K,
                     const void* mode, const double* concentration,
                     const double* weight, int weight_mode, double* affiliation,
                     void* workspace, size_t workspace_bytes, int* status,
                     void* stream);

/* ------------------------------------------------------------------------
 * Batched Hermitian eigendecomposition, ascending eigenvalues
 * (np.linalg.eigh as used in complex_angular_central_gaussian.py:95 and
 * pb_bss/utils.py:154).  a: (n, D, D) complex128 (only read), w: (n, D),
 * v: (n, D, D) complex128, columns are eigenvectors. */
int pbb_heig_batched(const void* a, int n, int D, double* w, void* v,
                     int* status, void* stream);

/* ------------------------------------------------------------------------
 * Beamforming side (pb_bss/extraction/beamformer.py).  All small matrices and
 * vectors are complex128; the observation may be complex64 or complex128.
 */

/* get_power_spectral_density_matrix (beamformer.py:59-160) for observation
 * (F, D, T) and mask (F, K, T) float64 (or NULL: plain average over time,
 * K must be 1).  normalize: divide by max(sum_t mask, 1e-10) (:127-131).
 * psd: (F, K, D, D). */
size_t pbb_psd_workspace_bytes(int F, int T, int D, int K);
int pbb_power_spectral_density(const void* observation, int dtype, int F,
                               int D, int T, const double* mask, int K,
                               int normalize, void* psd, void* workspace,
                               size_t workspace_bytes, void* stream);

/* get_gev_vector (beamformer.py:292-411): eigenvector of the largest
 * generalised eigenvalue of (target, noise), normalised like LAPACK zhegvd
 * ITYPE=1 (w^H noise w = 1).  Replaces _c_get_gev_vector
 * (cythonized/get_gev_vector.pyx:42-150); unlike it, matrices are row-major
 * (n, D, D).  status = 1 + index of the first pair whose noise matrix is not
 * positive definite (the Cython code raises ValueError there, :130-147). */
int pbb_gev_batched(const void* target_psd, const void* noise_psd, int n,
                    int D, void* w, int* status, void* stream);

/* np.linalg.solve for a batch: a (n, D, D), b (n, D, R) -> x (n, D, R), partial
 * pivoting.  hermitize != 0 solves with (a + a^H) / 2.  status flags singular
 * matrices (the reference falls back to lstsq, math/solve.py:95-114). */
int pbb_solve_batched(const void* a, const void* b, int n, int D, int R,
                      int hermitize, void* x, int* status, void* stream);

/* get_mvdr_vector (beamformer.py:230-260): w = N^-1 a / (a^H N^-1 a) with the
 * noise PSD hermitised first.  atf (n, D), noise_psd (n, D, D), w (n, D);
 * scratch: n * D complex128. */
int pbb_mvdr(const void* atf, const void* noise_psd, int n, int D, void* w,
             void* scratch, int* status, void* stream);

/* Pieces of get_mvdr_vector_souden (beamformer.py:601-698): phi = solve(noise,
 * target) -> mat = phi / max(trace(phi).real, eps) and, for every candidate
 * reference channel R, the per-bin numerator / denominator of the SNR
 * (num, den: (n, D) complex128) and their sums over the n bins (num_sum, den_sum:
 * (D) complex128) that get_optimal_reference_channel divides (beamformer.py:616-624). */
int pbb_souden(const void* phi, const void* target_psd, const void* noise_psd,
               int n, int D, double eps, void* mat, void* num, void* den,
               void* num_sum, void* den_sum, void* stream);

/* blind_analytic_normalization (beamformer.py:459-488). */
int pbb_blind_analytic_normalization(const void* vector, const void* noise_psd,
                                     int n, int D, void* out, void* stream);

/* Rank-1 PSD approximation a a^H * trace(cov) / trace(a a^H)
 * (get_pca_rank_one_estimate / get_gev_rank_one_estimate, beamformer_wrapper.py:11-69). */
int pbb_rank_one_estimate(const void* vector, const void* covariance, int n,
                          int D, void* out, void* stream);

/* out = matrix @ vector per batch entry (the scaled GEV ATF Phi_nn w, beamformer_wrapper.py:27-46). */
int pbb_matvec_batched(const void* matrix, const void* vector, int n, int D,
                       void* out, void* stream);

/* apply_beamforming_vector (beamformer.py:572-583): out[f][t] = sum_d conj(w[f][d]) mix[f][d][t]. */
int pbb_apply_beamforming_vector(const void* vector, const void* mix, int dtype,
                                 int F, int D, int T, void* out, void* stream);

/* The same for B beamformers per bin that share ONE mix (the K sources of a separation on one STFT; the reference
 * broadcasts the mix in its einsum): vector (B, F, D), mix (F, D, T), out (B, F, T).  B, F <= 65535. */
int pbb_apply_beamforming_vector_shared(const void* vector, const void* mix, int dtype,
                                        int B, int F, int D, int T, void* out, void* stream);

/* ------------------------------------------------------------------------
 * Frequency permutation alignment (pb_bss/permutation_alignment.py).
 */

/* DHTVPermutationAlignment.calculate_mapping (:295-355), similarity 'cos',
 * greedy assignment (:525-553).  mask (K, F, T) float64 is only read;
 * plan: nplan triples (iterations, start, end) as produced by
 * alignment_plan (:204-293) -- a small HOST array (it drives the launch
 * sequence); features (K, F, T) and centroid
 * (pbb_dhtv_scratch_doubles doubles) are device scratch; mapping (K, F) int64 out.
 * The reference's early exit is reproduced with device-side flags. */
size_t pbb_dhtv_scratch_doubles(int K, int T, const int* plan, int nplan);
int pbb_dhtv_mapping(const double* mask, int K, int F, int T, const int* plan,
                     int nplan, double* features, double* centroid,
                     long long* mapping, void* stream);
/* The same with the reference's options (permutation_alignment.py:133-163): metric 0 = 'multiply' (raw masks, inner
 * product), 1 = 'cos' (the default: features and centroid L2-normalised over time), 2 = 'euclidean' (minus the
 * distance); algorithm 0 = 'greedy', 1 = 'optimal' (brute force over the K! permutations).  The whole plan runs in
 * one launch: a single thread-block cluster that keeps a segment's feature rows in distributed shared memory when the
 * widest segment fits (<= 16 bins per CTA of a 16- or 8-CTA cluster and <= 200 KB; the reference's plans do), else one
 * cooperative launch with grid-wide barriers.  The integer mapping is the same on both. */
int pbb_dhtv_mapping_ex(const double* mask, int K, int F, int T, const int* plan,
                        int nplan, double* features, double* centroid,
                        long long* mapping, int metric, int algorithm, void* stream);

/* apply_mapping (:54-104): out[k, f, :] = mask[mapping[k, f], f, :]. */
int pbb_apply_mapping(const double* mask, const long long* mapping, int K,
                      int F, int T, double* out, void* stream);

/* _ScoreMatrix.multiply / cos / euclidean (:380-420): scores (F, K, K) with
 * scores[f][k_reference][k_mask] over the T frames of bin f.  Source k of a
 * side starts at base + k * source_stride (elements), its bins are T apart, so
 * the shifted views mask[:, 1:] / mask[:, :-1] of GreedyPermutationAlignment
 * (:702) are passed without a copy.  metric: */
enum { PBB_SCORE_MULTIPLY = 0, PBB_SCORE_COS = 1, PBB_SCORE_EUCLIDEAN = 2 };
int pbb_score_matrix(const double* mask, const double* reference,
                     long long mask_source_stride,
                     long long reference_source_stride, int K, int F, int T,
                     int metric, double* scores, void* stream);

/* _mapping_from_score_matrix (:458-590): scores (F, K, K) -> mapping (K, F)
 * int64.  algorithm 0 = 'greedy', 1 = 'optimal' (first best of
 * itertools.permutations).  *status = 1 + bin of a non-finite score matrix
 * (the reference raises ValueError('score matrix is infeasible')). */
enum { PBB_ASSIGN_GREEDY = 0, PBB_ASSIGN_OPTIMAL = 1 };
int pbb_mapping_from_score_matrix(const double* scores, int F, int K,
                                  int algorithm, long long* mapping,
                                  int* status, void* stream);

/* GreedyPermutationAlignment.calculate_mapping (:700-712), last step:
 * pair_mapping (K, F-1) = mapping of every bin to its lower neighbour;
 * mapping (K, F): column 0 identity, column f = pair[mapping[:, f-1], f-1]. */
int pbb_chain_mapping(const long long* pair_mapping, int K, int F,
                      long long* mapping, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PBB_H_ */
