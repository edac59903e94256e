/*
 * oak_b200.h -- C ABI of the B200-native OAK Gram / SGPR-statistics / Sobol hot path.
 *
 * Drop-in boundary for the reference's gpflow-Kernel-shaped Python API
 * (amzn/orthogonal-additive-gaussian-processes).  Each entry point names the reference
 * interface it replaces (file:line relative to the reference repository root).
 *
 * Conventions
 *   - plain C, no torch / C++ types; every function returns 0 on success, non-zero on
 *     failure; oak_last_error() returns a thread-local message for the last failure.
 *   - all array arguments named d_* are DEVICE pointers owned by the caller; h_* are HOST
 *     pointers.  `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - matrices are row-major FP64; `ld*` is the row stride in elements.
 *   - nothing is allocated behind the caller's back except inside oak_spec_create()
 *     (a few KB of parameters) and the per-process cuBLAS / cuSOLVER handles.
 *   - there is NO CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef OAK_B200_H
#define OAK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OAK_MAX_DEPTH 16 /* max_interaction_depth supported by the register-tiled kernels */

/* per-dimension sub-kernel kinds (oak/oak_kernel.py:128-189 picks one per input dim) */
enum {
  OAK_DIM_RBF = 0,         /* OrthogonalRBFKernel        oak/ortho_rbf_kernel.py:20-177      */
  OAK_DIM_BINARY = 1,      /* OrthogonalBinary           oak/ortho_binary_kernel.py:13-59    */
  OAK_DIM_CATEGORICAL = 2  /* OrthogonalCategorical      oak/ortho_categorical_kernel.py:14-74 */
};

/* input measures (oak/input_measures.py:16-78); NONE = unconstrained RBF
 * (constrain_orthogonal=False, oak/oak_kernel.py:191-210) */
enum {
  OAK_MEASURE_NONE = 0,
  OAK_MEASURE_GAUSSIAN = 1,
  OAK_MEASURE_UNIFORM = 2,
  OAK_MEASURE_EMPIRICAL = 3,
  OAK_MEASURE_MOG = 4
};

/* how e_n is formed from the per-dimension kernels */
enum {
  OAK_ESP_NEWTON_GIRARD = 0, /* power sums + Newton-Girard, oak/oak_kernel.py:236-249 (the reference) */
  OAK_ESP_DIRECT = 1         /* e_n += k_d * e_{n-1}; same polynomial, better conditioned, fewer
                                FP64 instructions: what the Python host layer passes by default */
};

/* One input dimension.  All pointers are HOST pointers, read during oak_spec_create(). */
typedef struct oak_dim_desc {
  int32_t type;     /* OAK_DIM_*                                                        */
  int32_t column;   /* column of X this sub-kernel acts on (active_dims, oak_kernel.py:75-76) */
  int32_t measure;  /* OAK_MEASURE_* (RBF only)                                         */
  int32_t count;    /* #locations (EMPIRICAL) / #components (MOG) / #categories (CATEGORICAL) */
  int32_t rank;     /* CATEGORICAL: columns of W (ortho_categorical_kernel.py:22)       */
  int32_t reserved;
  double lengthscale; /* RBF base_kernel.lengthscales                                   */
  double variance;    /* RBF base_kernel.variance, or the discrete kernel's .variance   */
  double m0;          /* GAUSSIAN: mu       UNIFORM: a       BINARY: p0                 */
  double m1;          /* GAUSSIAN: var      UNIFORM: b                                  */
  const double* v0;   /* EMPIRICAL: location[count]  MOG: means[count]      CATEGORICAL: W[count*rank] row-major */
  const double* v1;   /* EMPIRICAL: weights[count]   MOG: variances[count]  CATEGORICAL: kappa[count]            */
  const double* v2;   /*                             MOG: weights[count]    CATEGORICAL: p[count]                */
} oak_dim_desc;

/* The composite kernel: replaces the state held by OAKKernel (oak/oak_kernel.py:59-221). */
typedef struct oak_kernel_desc {
  int32_t num_dims;                /* number of sub-kernels                                   */
  int32_t depth;                   /* max_interaction_depth                                   */
  int32_t share_var_across_orders; /* oak_kernel.py:212-221                                   */
  int32_t esp_algorithm;           /* OAK_ESP_*                                               */
  const double* variances;         /* HOST: sigma^2_0..depth (share_var) or sigma^2_0 only    */
  const oak_dim_desc* dims;        /* HOST: num_dims entries                                  */
} oak_kernel_desc;

typedef struct oak_spec oak_spec; /* opaque, device-resident parameter block */

/* ---- housekeeping ------------------------------------------------------------------ */
const char* oak_last_error(void);
int oak_version(void);
/* number of CUDA devices visible (0 => every compute call fails loudly) */
int oak_device_count(void);

/* Packs the hyper-parameters once per evaluation (they are gpflow Parameters read at call
 * time in the reference, oak_kernel.py:251-265).  Computes var_s() for every constrained
 * RBF dim (ortho_rbf_kernel.py:65-78, 94-97, 109-120, 138-152) and the B tables of the
 * discrete dims (ortho_binary_kernel.py:29-38, ortho_categorical_kernel.py:34-53). */
int oak_spec_create(const oak_kernel_desc* desc, void* stream, oak_spec** out);
int oak_spec_destroy(oak_spec* spec);
/* var_s() of constrained RBF sub-kernel `dim` (the closure at ortho_rbf_kernel.py:154-155);
 * synchronises the stream. */
int oak_spec_var_s_f64(const oak_spec* spec, int32_t dim, double* h_var_s, void* stream);

/* ---- per-point prologue ------------------------------------------------------------ */
/* Bytes of the prepared-point block for n points (feature-major, padded). */
size_t oak_points_bytes(const oak_spec* spec, int64_t n);
/* X: (n x ldx) row-major device matrix.  Writes, per dim and point, the scaled coordinate
 * x/(sqrt(2) l) and the normalised correction cov_X_s(x)/sqrt(var_s())
 * (ortho_rbf_kernel.py:47-152, 163-167), or the int32 category index
 * (ortho_binary_kernel.py:47-51). */
int oak_prepare_points_f64(const oak_spec* spec, const double* d_X, int64_t n, int64_t ldx,
                           void* d_points, void* stream);

/* ---- Gram / cross-covariance ------------------------------------------------------- */
/* Replaces OAKKernel.K(X, X2) (oak_kernel.py:251-265) incl. compute_additive_terms
 * (:223-249) and every sub-kernel K (ortho_rbf_kernel.py:157-172,
 * ortho_binary_kernel.py:40-53, ortho_categorical_kernel.py:55-68).
 * Computes rows [row_begin,row_end) of K(X, X2) into d_K[(i-row_begin)*ldk + j].
 * d_points2 == NULL means X2 = X (symmetric; with row range == [0,n) only the lower
 * triangle of tiles is evaluated and mirrored). */
int oak_gram_f64(const oak_spec* spec, const void* d_points, int64_t n, const void* d_points2,
                 int64_t n2, int64_t row_begin, int64_t row_end, double* d_K, int64_t ldk,
                 void* stream);

/* Lower trapezoid of the symmetric Gram K(X, X): rows [row_begin,row_end) x columns
 * [0,row_end), evaluating only tiles that intersect the lower triangle (entries above the
 * diagonal outside the diagonal tiles are left untouched).  This is the per-rank unit of the
 * row-strip sharded symmetric Gram; no mirroring, no collective.
 * d_K[(i-row_begin)*ldk + j]. */
int oak_gram_lower_f64(const oak_spec* spec, const void* d_points, int64_t n, int64_t row_begin,
                       int64_t row_end, double* d_K, int64_t ldk, void* stream);

/* oak_gram_lower_f64 plus the mirror image of the strip: d_Kt is a (row_end x (row_end - row_begin)) block,
 * d_Kt[j*ldkt + (i-row_begin)] = K(i, j) for the entries of the off-diagonal tiles (tiles on the diagonal are
 * written in full to d_K).  The strips of all ranks then hold the whole symmetric matrix between them -- the
 * same product as the single-GPU oak_gram_f64 call -- still without a collective. */
int oak_gram_lower_mirror_f64(const oak_spec* spec, const void* d_points, int64_t n, int64_t row_begin,
                              int64_t row_end, double* d_K, int64_t ldk, double* d_Kt, int64_t ldkt,
                              void* stream);

/* Replaces OAKKernel.K_diag(X) (oak_kernel.py:267-278). d_out[n]. */
int oak_gram_diag_f64(const oak_spec* spec, const void* d_points, int64_t n, double* d_out,
                      void* stream);

/* Replaces KernelComponenent.K (oak_kernel.py:300-322): sigma^2_|S| * prod_{d in S} k_d.
 * h_subset: `order` indices into the spec's dims (order may be 0: constant term). */
int oak_component_gram_f64(const oak_spec* spec, const int32_t* h_subset, int32_t order,
                           const void* d_points, int64_t n, const void* d_points2, int64_t n2,
                           double* d_K, int64_t ldk, void* stream);

/* Replaces KernelComponenent.K_diag (oak_kernel.py:324-335). d_out[n]. */
int oak_component_diag_f64(const oak_spec* spec, const int32_t* h_subset, int32_t order,
                           const void* d_points, int64_t n, double* d_out, void* stream);

/* Replaces OAKKernel.compute_additive_terms on explicit arrays (oak_kernel.py:223-249):
 * d_mats is (num_mats x len), d_out is ((depth+1) x len) = e_0..e_depth, element-wise. */
int oak_additive_terms_f64(const double* d_mats, int32_t num_mats, int32_t depth, int64_t len,
                           double* d_out, void* stream);

/* Fused per-component prediction (oak/utils.py:491-530 get_prediction_component):
 * out[c*n + i] = sigma^2_|S_c| * sum_j prod_{d in S_c} k_d(x_i, z_j) * alpha[j], for a batch
 * of components; d_subsets is (num_components x max_order) int32, padded with -1. */
int oak_component_predict_f64(const oak_spec* spec, const int32_t* d_subsets,
                              int32_t num_components, int32_t max_order, const void* d_points,
                              int64_t n, const void* d_points_cond, int64_t m,
                              const double* d_alpha, double* d_out, void* stream);

/* out[i - row_begin] = sum_j K(x_i, x2_j) alpha_j for rows [row_begin, row_end) of `points`
 * (row_begin a multiple of 64): the mean of gpflow's predict_f, Kus^T alpha
 * (oak/model_utils.py:429-443), fused so that the N* x M matrix is never formed.
 * max_interaction_depth <= 8. */
int oak_gram_matvec_f64(const oak_spec* spec, const void* d_points, int64_t n, int64_t row_begin,
                        int64_t row_end, const void* d_points2, int64_t n2, const double* d_alpha,
                        double* d_out, void* stream);

/* Host-buffer convenience (the reference-facing call: NumPy in, NumPy out).  Copies X (and
 * X2) to the device, runs prepare + gram in row blocks and streams K back, overlapping the
 * D2H copies with compute.  h_X2 == NULL => X2 = X. d_work must hold oak_gram_host_work_bytes. */
size_t oak_gram_host_work_bytes(const oak_spec* spec, int64_t n, int64_t n2, int64_t ldx,
                                int64_t block_rows);
int oak_gram_host_f64(const oak_spec* spec, const double* h_X, int64_t n, const double* h_X2,
                      int64_t n2, int64_t ldx, double* h_K, int64_t ldk, int64_t block_rows,
                      void* d_work, void* stream);

/* Symmetric K(X, X) through host buffers, lower trapezoid only: rows [row_begin, row_end) are produced in
 * blocks of block_rows and only columns [0, end of the block) of each block are copied back -- half the PCIe
 * bytes of oak_gram_host_f64 for the whole matrix.  h_K[(i - row_begin)*ldk + j] for j <= i (and the rest of the
 * 64 x 64 tiles on the diagonal); entries further right are left untouched unless mirror != 0 (full row range
 * only), which fills the strict upper triangle on the host from the lower one.  h_X: all n points (n x ldx). */
size_t oak_gram_host_lower_work_bytes(const oak_spec* spec, int64_t n, int64_t ldx, int64_t row_end,
                                      int64_t block_rows);
int oak_gram_host_lower_f64(const oak_spec* spec, const double* h_X, int64_t n, int64_t ldx, int64_t row_begin,
                            int64_t row_end, double* h_K, int64_t ldk, int64_t block_rows, int mirror,
                            void* d_work, void* stream);

/* ---- input pipeline on the device ------------------------------------------ */
/* np.unique(X[:, col], return_counts=True) of a row-major (n x ldx) device matrix: sorted distinct values and
 * their multiplicities (sort + run-length on order-preserving 64-bit keys: values are the input bit patterns,
 * counts exact).  Replaces the per-column host calls that build the empirical-measure locations / weights
 * (oak/model_utils.py:334-344) and the category frequencies p (oak/model_utils.py:736-739).
 * d_vals[n], d_counts[n]: the first *d_num entries are written. */
size_t oak_column_unique_work_bytes(int64_t n);
int oak_column_unique_f64(const double* d_X, int64_t n, int64_t ldx, int64_t col, double* d_vals,
                          int64_t* d_counts, int32_t* d_num, void* d_work, void* stream);
/* d_out[0] = mean of column col (p0 = 1 - mean of a binary feature, oak/model_utils.py:731); fixed summation
 * order, exact for 0/1 columns.  d_work: ceil(n / 4096) + 1 doubles. */
int oak_column_mean_f64(const double* d_X, int64_t n, int64_t ldx, int64_t col, double* d_out, void* d_work,
                        void* stream);

/* ---- SGPR statistics --------------------------------------------------------------- */
/* Layout of the packed statistics vector (doubles): Phi[M*M] | Kuf_y[M] | sum_kdiag | yty.
 * This is the single buffer all-reduced (sum) across ranks. */
size_t oak_sgpr_stats_count(int64_t m);
size_t oak_sgpr_stats_work_bytes(int64_t m, int64_t chunk);
/* Replaces the Kuf build + A A^T contraction of SGPR.elbo / get_model_sufficient_statistics
 * (oak/utils.py:180-191; gpflow Kuf at utils.py:184).  Streams the local n points in chunks:
 * Kuf chunk (M x chunk, L2 resident) -> Phi += Kuf Kuf^T (cuBLAS DSYRK), Kuf_y += Kuf y,
 * sum_kdiag += sum K_diag(X), yty += y^T y.  Accumulates INTO d_stats (caller zeroes it).
 * The contraction runs on the FP64 tensor cores (csrc/oak_syrk.cu). */
int oak_sgpr_stats_f64(const oak_spec* spec, const void* d_pointsZ, int64_t m,
                       const void* d_pointsX, const double* d_y, int64_t n_local, int64_t chunk,
                       double* d_stats, void* d_work, void* stream);
/* Same, keeping Kuf for a following backward pass: chunk c is written to d_kuf_store + c*m*chunk
 * as an m x chunk row-major block (ld = chunk); d_kuf_store holds ceil(n_local/chunk) blocks. */
int oak_sgpr_stats_keep_f64(const oak_spec* spec, const void* d_pointsZ, int64_t m,
                            const void* d_pointsX, const double* d_y, int64_t n_local, int64_t chunk,
                            double* d_stats, void* d_work, double* d_kuf_store, void* stream);

/* M^3 tail of SGPR.elbo (gpflow 2.2.1; re-derived at oak/utils.py:187-198): Cholesky of
 * Kuu + jitter I, whitening of Phi, Cholesky of B, c, the bound and alpha = L^-T LB^-T c.
 * d_Kuu (M x M) and d_stats are overwritten.  d_out: [elbo, logdetB_half, trace_AAT, c^T c];
 * d_alpha[M] may be NULL. */
size_t oak_sgpr_finish_work_bytes(int64_t m);
int oak_sgpr_finish_f64(double* d_Kuu, double* d_stats, int64_t m, int64_t n_total, double noise,
                        double jitter, double* d_out, double* d_alpha, void* d_work,
                        void* stream);

/* ---- SGPR, factor-first path (default since round 2) ---------------------------------------
 * gpflow's SGPR.elbo whitens Kuf BEFORE the contraction: A = L^-1 Kuf / sigma, then A A^T
 * (oak/utils.py:186-190).  Forming Phi = Kuf Kuf^T first and whitening it once in the tail is cheaper
 * (no M^2 n triangular product) but its rounding is amplified by cond(Kuu): 2e-6 relative ELBO error at
 * cond(Kuu) = 2.6e8 (long lengthscales), 1e-13 at 1e4.  Here L = chol(Kuu + jitter I) is computed first
 * and the route is chosen ON THE DEVICE from a condition estimate, so no host synchronisation is needed:
 *   route 0  Phi statistics + one whitening in the tail          (cond estimate < threshold)
 *   route 1  per chunk A_r = L^-1 Kuf_r, Psi = sum A_r A_r^T      (gpflow's operation order)
 * Both routes all-reduce the same M*M + M + 2 doubles.
 *
 * d_fac (oak_sgpr_factor_count(m) doubles, 16-byte aligned): column-major matrix with leading dimension
 * LD = oak_sgpr_factor_ld(m) = 2 * roundup8(m) and m columns: rows [0, m) hold L (lower triangle), rows
 * [LD/2, LD/2 + m) hold L^-T (upper triangle) -- i.e. the row-major view d_fac[r*LD + LD/2 + c] is L^-1 --
 * followed by a 16-double header {cond estimate, ||Kuu||_1, lambda_max(Kuu^-1) estimate, route,
 * info of chol(Kuu), sum log diag L, threshold, ...} and scratch. */
size_t oak_sgpr_factor_count(int64_t m);
int64_t oak_sgpr_factor_ld(int64_t m);
/* Kuu(iv, kernel) + jitter I (oak/utils.py:185), L = chol(Kuu) (:188), L^-1, the condition estimate and
 * the route flag.  route: -1 = decide from the estimate (cond_threshold <= 0: default 3e5), 0 / 1 = forced. */
int oak_sgpr_factor_f64(const oak_spec* spec, const void* d_pointsZ, int64_t m, double jitter, int route,
                        double cond_threshold, double* d_fac, void* stream);
/* The chunk loop of oak_sgpr_stats_f64 on the route stored in d_fac: d_stats accumulates
 * Phi (route 0) or Psi (route 1) | Kuf y | sum K_diag | y^T y.  d_kuf_store may be NULL (else every chunk's
 * Kuf block is kept as in oak_sgpr_stats_keep_f64). */
size_t oak_sgpr_stats2_work_bytes(int64_t m, int64_t chunk);
int oak_sgpr_stats2_f64(const oak_spec* spec, const void* d_pointsZ, int64_t m, double* d_fac,
                        const void* d_pointsX, const double* d_y, int64_t n_local, int64_t chunk,
                        double* d_stats, void* d_work, double* d_kuf_store, void* stream);
/* oak_sgpr_factor_f64 followed by oak_sgpr_stats2_f64 as one call that takes the factorisation off the
 * critical path (it is replicated on every rank): the first chunk's Kuf tiles (oak/utils.py:184 -- they need
 * neither L nor the route) are launched first on `stream` and leave `overlap_ctas` SMs free; the Kuu tiles,
 * [L ; L^-T], the condition estimate and the route flag run on an internal side stream capped at that many
 * CTAs; `stream` joins the side stream before the first kernel that reads the flag, so on return everything
 * is ordered on `stream`.
 * overlap_ctas: 0 = serial (exactly the two calls above), > 0 = that many CTAs for the factorisation,
 * < 0 = automatic (4 or 8 when the first chunk is long enough to cover it, else serial; OAK_SGPR_OVERLAP overrides).
 * Results are bit-identical for every value of overlap_ctas. */
int oak_sgpr_factor_stats_f64(const oak_spec* spec, const void* d_pointsZ, int64_t m, double jitter, int route,
                              double cond_threshold, double* d_fac, const void* d_pointsX, const double* d_y,
                              int64_t n_local, int64_t chunk, double* d_stats, void* d_work, double* d_kuf_store,
                              int overlap_ctas, void* stream);
/* Tail of SGPR.elbo (oak/utils.py:190-198) from the (all-reduced) statistics: B = I + A A^T, LB = chol(B),
 * c = LB^-1 A err / sigma, the bound, alpha = L^-T LB^-T c.  d_LB: oak_sgpr_lb_ld(m) * m doubles, column-major
 * with that leading dimension: rows [0, m) = LB (lower triangle), row roundup8(m) = c^T.
 * d_out[8] = {elbo, sum log diag LB, tr(A A^T), c^T c, info of chol(Kuu), info of chol(B), route, cond
 * estimate}: a non-zero info (leading minor not positive definite) is reported there, not by a host
 * synchronisation.  d_alpha[m] may be NULL. */
int64_t oak_sgpr_lb_ld(int64_t m);
size_t oak_sgpr_finish2_work_bytes(int64_t m);
int oak_sgpr_finish2_f64(double* d_fac, double* d_stats, int64_t m, int64_t n_total, double noise,
                         double* d_out, double* d_alpha, double* d_LB, void* d_work, void* stream);

/* ---- k-means for the inducing-point initialisation (csrc/oak_kmeans.cu) -----------------------
 * Replaces the scikit-learn KMeans(n_clusters, random_state=0).fit(X).cluster_centers_ calls of
 * oak/model_utils.py:31-41, 376-383 and oak/utils.py:549-552, 570-573 on the continuous columns: the same
 * algorithm (centring, k-means++ with 2 + log k local trials, Lloyd with first-index ties, centres = sums / counts,
 * strict / tolerance stopping rule); the random numbers are drawn on the host from numpy's RandomState exactly as
 * scikit-learn draws them, everything O(N) runs here.  All reductions have a fixed order (deterministic). */
size_t oak_kmeans_work_bytes(int64_t n, int64_t d, int64_t k, int64_t trials);
/* d_Xc (n x d, contiguous) = X - column means; d_stats[0..d) = the means, d_stats[d] = mean of the column variances
 * (sklearn's tolerance scale, _kmeans.py:_tolerance). */
int oak_kmeans_center_f64(const double* d_X, int64_t n, int64_t d, int64_t ldx, double* d_Xc, double* d_stats,
                          void* d_work, void* stream);
/* One k-means++ round (_kmeans.py:_kmeans_plusplus): candidates d_cand[t] = searchsorted(cumsum(d_closest),
 * d_thresholds[t]) clipped to n - 1 (d_thresholds == NULL: d_cand is given), d_dist[t*n + i] = min(d_closest[i],
 * |x_i - x_cand[t]|^2) (d_closest == NULL: no min), d_pots[t] = sum_i d_dist[t*n + i].  d_cum: n doubles. */
int oak_kmeanspp_round_f64(const double* d_Xc, int64_t n, int64_t d, const double* d_closest,
                           const double* d_thresholds, int64_t trials, int64_t* d_cand, double* d_cum, double* d_dist,
                           double* d_pots, void* d_work, void* stream);
/* One Lloyd iteration (_k_means_lloyd.pyx lloyd_iter_chunked_dense, _k_means_common.pyx _average_centers /
 * _center_shift).  d_labels (int32): previous labels in, new labels out.  d_out: 4 x 8 bytes: [0] double, sum of the
 * squared centre shifts; [2] uint64, number of empty clusters (their rows of d_centers_new keep the old centre);
 * [3] uint64, number of labels that changed.  update_centers == 0: labels only.  d <= 64 runs the register-tiled
 * assignment, wider inputs a plain one. */
int oak_kmeans_lloyd_f64(const double* d_Xc, int64_t n, int64_t d, const double* d_centers, int64_t k,
                         int32_t* d_labels, double* d_centers_new, double* d_sums, double* d_counts, double* d_out,
                         int update_centers, void* d_work, void* stream);

/* E step + M-step sums of a one-dimensional spherical Gaussian mixture: the MOG input measure fitted by
 * GaussianMixture(n_components=K, random_state=0, covariance_type="spherical") in oak/model_utils.py:753-770
 * (sklearn/mixture/_gaussian_mixture.py: _estimate_log_gaussian_prob, _estimate_gaussian_parameters).
 * d_par: 4 x K doubles = means | precisions | log precisions_cholesky | log weights; d_labels (int32, nullable): hard
 * responsibilities of the k-means initialisation instead of the E step.  d_out[3K + 1] = sum_i resp_ik |
 * sum_i resp_ik x_i | sum_i resp_ik x_i^2 | sum_i log p(x_i).  K <= 16. */
size_t oak_gmm1d_work_bytes(int64_t n, int64_t K);
int oak_gmm1d_estep_f64(const double* d_x, int64_t n, int64_t K, const double* d_par, const int32_t* d_labels,
                        double* d_out, void* d_work, void* stream);

/* ---- dense building blocks of the tails (csrc/oak_chol.cu, csrc/oak_pgemm.cu) ---------------- */
/* Cholesky factorisation with border rows in one cooperative launch, replacing the tf.linalg.cholesky +
 * tf.linalg.triangular_solve pairs of oak/utils.py:188-195.  d_A: column-major, leading dimension ld
 * (equivalently: the UPPER triangle of a row-major matrix); the n x n symmetric block is read and
 * overwritten in its lower triangle by L; rows [n, rows) -- stored `gap` rows further down -- are
 * overwritten by Border * L^-T.  border_identity != 0 declares that border row n + i holds e_i on entry
 * (rows - n == n), which lets the kernel skip the still-zero part.  *d_info: 0, or k > 0 when the leading
 * minor of order k is not positive definite.  d_logdet (may be NULL): sum_i log L_ii. */
int oak_chol_f64(double* d_A, int64_t n, int64_t rows, int64_t gap, int64_t ld, int border_identity,
                 int32_t* d_info, double* d_logdet, void* stream);
/* C (m x n, ldc) = T (m x kd, ldt) * B (kd x n, ldb) [+ u v^T], row-major, on the FP64 tensor cores:
 * the whitening product A_r = L^-1 Kuf_r of route 1 (lower != 0: T is lower triangular with zeros above
 * the diagonal) and the cotangent W = 2 G_Phi Kuf + g_b y^T of the training step.  Leading dimensions
 * even, bases 16-byte aligned, ldb >= n rounded up to even.  d_work: oak_panel_gemm_work_bytes(). */
size_t oak_panel_gemm_work_bytes(void);
int oak_panel_gemm_f64(const double* d_T, int64_t ldt, const double* d_B, int64_t ldb, double* d_C,
                       int64_t ldc, int64_t m, int64_t kd, int64_t n, int lower, const double* d_u,
                       const double* d_v, void* d_work, void* stream);

/* GPR: log N(y; 0, K + noise I) and alpha = (K + noise I)^-1 y (oak/utils.py:206-211;
 * gpflow GPR.log_marginal_likelihood).  d_K (n x n) is overwritten by its Cholesky factor. */
size_t oak_gpr_finish_work_bytes(int64_t n);
int oak_gpr_finish_f64(double* d_K, const double* d_y, int64_t n, double noise, double* d_lml,
                       double* d_alpha, void* d_work, void* stream);

/* ---- Sobol indices ------------------------------------------------------------------ */
/* Replaces compute_L / compute_L_binary_kernel / compute_L_categorical_kernel /
 * compute_L_empirical_measure (oak/utils.py:221-335): the m x m matrix L_d for sub-kernel
 * `dim` at the conditioning points (unit order-variance; the caller's variance enters in
 * oak_sobol_quadforms_f64).  d_Xcond is the raw (m x ldx) conditioning matrix. */
int oak_sobol_L_f64(const oak_spec* spec, int32_t dim, const double* d_Xcond, int64_t m,
                    int64_t ldx, double delta, double mu, double* d_L, int64_t ldl,
                    void* d_work, void* stream);
size_t oak_sobol_L_work_bytes(const oak_spec* spec, int32_t dim, int64_t m);
/* The four closed-form terms f1..f4 of the Gaussian-measure L (oak/utils.py:116-165, eq. 44-47 of the paper)
 * on their own, elementwise over n paired points: d_out[k * n + i] = f_{k+1}(x_i, y_i, sigma, lengthscale,
 * delta, mu), k = 0..3 (f3(x, y) = f2(y, x)).  The same device functions build L in oak_sobol_L_f64. */
int oak_sobol_gaussian_terms_f64(const double* d_x, const double* d_y, int64_t n, double sigma,
                                 double lengthscale, double delta, double mu, double* d_out, void* stream);
/* Replaces the component loop of compute_sobol_oak (oak/utils.py:369-432):
 * out[c] = scale[c] * alpha^T (prod_{d in S_c} L_d) alpha.  d_Lstack: (num_dims x m x m); d_work:
 * oak_sobol_quadforms_work_bytes bytes (row-split partial sums, folded in a fixed order). */
size_t oak_sobol_quadforms_work_bytes(int32_t num_components, int64_t m);
int oak_sobol_quadforms_f64(const double* d_Lstack, int32_t num_dims, int64_t m,
                            const int32_t* d_subsets, const double* d_scale,
                            int32_t num_components, int32_t max_order, const double* d_alpha,
                            double* d_out, void* d_work, void* stream);

/* ---- backward tiles (training) ------------------------------------------------------ */
/* Replaces TensorFlow's autodiff through OAKKernel.K / K_diag inside the gpflow objectives
 * (oak/oak_kernel.py:223-278 differentiated; optimiser loop at oak/model_utils.py:168-175).
 * For a cotangent W = d objective / d K (rows [row_begin,row_end) of points x all of points2,
 * pitch ldw; d_points2 == NULL => the same point set):
 *   d_grad[i]             += sum W * dK/d lengthscale_i   (i = sub-kernel in the caller's order;
 *                            RBF sub-kernels under any measure -- entries of discrete
 *                            sub-kernels are left untouched)
 *   d_grad[num_dims + n]  += sum W * e_n = dK/d sigma2_n  (n = 0..max_interaction_depth)
 *   d_grad[num_dims + depth + 1 + t] += cotangent of entry t of the discrete kernels' table blob
 *                            (binary / categorical B tables and their diagonals; the caller chains
 *                            it to W, kappa, variance -- oak_spec_table_layout gives the offsets)
 *   d_grad[count - num_dims + i] += sum W * dK/d s2_i   (base variance of RBF sub-kernel i)
 * d_grad has count = oak_backward_grad_count(spec) entries.  max_interaction_depth <= 8.  d_work: oak_gram_backward_work_bytes(spec, n_points).
 * Empirical / uniform / MOG dims need the per-point derivative block d c^/dl of both point sets, written
 * by oak_prepare_backward_f64 from the prepared points (oak_backward_points_bytes bytes each);
 * pass NULL when the kernel has none. */
size_t oak_backward_grad_count(const oak_spec* spec);
int oak_spec_table_layout(const oak_spec* spec, int32_t dim, int32_t* offset, int32_t* count);
size_t oak_backward_points_bytes(const oak_spec* spec, int64_t n);
int oak_prepare_backward_f64(const oak_spec* spec, const void* d_points, int64_t n, void* d_dpoints,
                             void* stream);
size_t oak_gram_backward_work_bytes(const oak_spec* spec, int64_t n);
int oak_gram_backward_f64(const oak_spec* spec, const void* d_points, const void* d_dpoints, int64_t n,
                          int64_t row_begin, int64_t row_end, const void* d_points2,
                          const void* d_dpoints2, int64_t n2, const double* d_W, int64_t ldw,
                          double* d_grad, void* d_work, void* stream);
/* The same contraction over ALL rows of d_points, plus the gradient with respect to the row points
 * themselves -- the inducing points Z of SGPR when the reference leaves them trainable
 * (zfixed=False, oak/model_utils.py:98-101; TensorFlow autodiff through OAKKernel.K(Z, X) and K(Z, Z)):
 *   d_grad_rows[i * ldg + k] += sum_j W_ij * dK(x_i, y_j) / d x_{i,k}    (k = sub-kernel index, caller's order)
 * for RBF sub-kernels under every measure; discrete sub-kernels receive nothing (tf.cast / tf.gather carry
 * no gradient).  Only the FIRST argument of K is differentiated: for a symmetric objective over K(Z, Z)
 * pass W + W^T.  ldg >= num_dims.  d_work: oak_gram_backward_rows_work_bytes(spec, n, n2). */
size_t oak_gram_backward_rows_work_bytes(const oak_spec* spec, int64_t n, int64_t n2);
int oak_gram_backward_rows_f64(const oak_spec* spec, const void* d_points, const void* d_dpoints, int64_t n,
                               const void* d_points2, const void* d_dpoints2, int64_t n2, const double* d_W,
                               int64_t ldw, double* d_grad, double* d_grad_rows, int64_t ldg, void* d_work,
                               void* stream);
/* Same for wscale * sum_i w_i K_diag(x_i) (d_w == NULL => all ones). */
int oak_gram_diag_backward_f64(const oak_spec* spec, const void* d_points, const void* d_dpoints,
                               int64_t n, const double* d_w, double wscale, double* d_grad,
                               void* d_work, void* stream);

/* ---- whitened SVGP (diagonal q) with a Bernoulli likelihood ---------------------------------
 * Replaces, around the Kuf tiles, what gpflow.models.SVGP(kernel=OAK, likelihood=Bernoulli(invlink=inv_logit),
 * whiten=True, q_diag=True) computes in the reference's classification run
 * (examples/uci/uci_classification_train.py:108-135: elbo under BFGS, predict_f, predict_log_density).
 * A = L^-1 Kuf (m x n, row-major, pitch lda) comes from oak_gram_f64 and one triangular solve. */
enum oak_link {
  OAK_LINK_LOGIT = 0,  /* sigmoid(f) (1 - 2 jitter) + jitter: the reference's inv_logit (:43-45)           */
  OAK_LINK_PROBIT = 1  /* 0.5 (1 + erf(f / sqrt 2)) (1 - 2 jitter) + jitter: gpflow's default inv_probit    */
};
/* mean_i = sum_r A_ri q_mu_r;  var_i = kdiag_i - sum_r A_ri^2 (1 - q_sqrt_r^2)   (gpflow base_conditional, white) */
int oak_svgp_moments_f64(const double* d_A, int64_t lda, int32_t m, int64_t n, const double* d_q_mu,
                         const double* d_q_sqrt, const double* d_kdiag, double* d_mean, double* d_var,
                         void* stream);
/* Cotangents of the above: d_Abar (m x n, pitch ldb) = q_mu gmean^T - 2 (1 - q_sqrt^2) A diag(gvar);
 * d_gq_sqrt[r] += 2 q_sqrt_r sum_i gvar_i A_ri^2. */
int oak_svgp_moments_backward_f64(const double* d_A, int64_t lda, int32_t m, int64_t n, const double* d_q_mu,
                                  const double* d_q_sqrt, const double* d_gmean, const double* d_gvar,
                                  double* d_Abar, int64_t ldb, double* d_gq_sqrt, void* stream);
/* Gauss-Hermite expectations of the Bernoulli log density under N(mean_i, var_i) (gpflow
 * Bernoulli.variational_expectations / predict_log_density, NDiagGHQuadrature): nodes d_gh_x (already times
 * sqrt 2) and weights d_gh_w (already over sqrt pi), n_gh <= 64.  Outputs (each may be NULL):
 * d_varexp[i] = sum_k w_k log p(y_i | f_k), its derivatives d_gmean / d_gvar with respect to mean_i / var_i,
 * d_logdensity[i] = log sum_k w_k p(y_i | f_k).  y_i == 1 selects p, anything else 1 - p (tf.where(y == 1, ..)). */
int oak_bernoulli_quadrature_f64(const double* d_mean, const double* d_var, const double* d_y, int64_t n,
                                 int32_t link, double jitter, const double* d_gh_x, const double* d_gh_w,
                                 int32_t n_gh, double* d_varexp, double* d_gmean, double* d_gvar,
                                 double* d_logdensity, void* stream);

/* ---- input pipeline: per-column normalising flow (oak/normalising_flow.py) ---------------------
 * y = sinh((asinh(z) + skewness) * tailweight), z = (u + shift) * scale, u = use_log ? log(x - offset) : x
 * -- the bijector chain of normalising_flow.py:46-56 with TFP 0.11's SinhArcsinh (setup.py:33).  One strided
 * column per call (stride in doubles), so the columns of a row-major X are transformed in place in the
 * layout oak_prepare_points_f64 reads (model_utils.py:179-191 apply_normalise_flow). */
int oak_flow_forward_f64(const double* d_x, int64_t n, int64_t stride_in, double offset, int32_t use_log,
                         double shift, double scale, double skewness, double tailweight, double* d_y,
                         int64_t stride_out, void* stream);
/* KL_objective (normalising_flow.py:76-81) J = mean(y^2 / 2) - mean(log |dy/dx|) of one column and its
 * gradient in the optimiser's variables: d_out5 = [J, dJ/d log(scale), dJ/d shift, dJ/d skewness,
 * dJ/d log(tailweight)] -- what gpflow.optimizers.Scipy (L-BFGS-B) evaluates per iteration
 * (model_utils.py:313-317).  d_work: oak_flow_objective_work_bytes(n). */
size_t oak_flow_objective_work_bytes(int64_t n);
int oak_flow_objective_f64(const double* d_x, int64_t n, int64_t stride, double offset, int32_t use_log,
                           double log_scale, double shift, double skewness, double log_tailweight,
                           double* d_out5, void* d_work, void* stream);

/* ---- measurement helpers ----------------------------------------------------------- */
/* Dependent-chain DFMA microbenchmark: writes achieved FP64 issue slots / second to
 * *h_slots_per_s (1 slot = one DFMA = 2 flop).  This is the measured FP64 roofline peak. */
int oak_measure_fp64_peak(double seconds, double* h_slots_per_s, void* stream);
/* Number of kernels launched by this library since load (bench.py's gpu_launches). */
int64_t oak_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* OAK_B200_H */
