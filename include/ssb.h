/*
 * ssb.h -- C-ABI of the B200-native iterative demixing path (ILRMA / AuxIVA).
 *
 * The reference (tky823/ssspy) is pure Python/NumPy and has no FFI layer; the boundary this
 * library sits behind is its separator-class API.  Every entry point below
 * names the reference function it replaces (path:line under the reference tree).  The Python
 * host mirror (ssspy_b200/) binds these with ctypes; INTEGRATION.md shows the stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - plain C, no exceptions; every function returns 0 on success, non-zero on error;
 *     ssb_last_error() returns the message of the last failure on the calling thread.
 *   - all array arguments are DEVICE pointers owned by the caller unless marked "host";
 *     complex64 = interleaved (re, im) float pairs, complex128 = interleaved doubles.
 *   - layouts are the reference's with a leading batch axis of independent mixtures:
 *       X[B,N,I,J] c64   mixture STFT            (ssspy/bss/ilrma.py:840, input)
 *       W[B,I,N,N] c64   demixing filters, rows are w_n^H   (ilrma.py:186-188)
 *       Y[B,N,I,J] c64   separated STFT          (ilrma.py:292-295)
 *       T[B,N,I,K] f32   NMF basis, V[B,N,K,J] f32 NMF activation  (ilrma.py:256-268)
 *   - every launch is asynchronous on the caller's stream (a cudaStream_t passed as void*);
 *     no hidden host synchronisation.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef SSB_H_
#define SSB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSB_VERSION 100

/* model: which separator class the plan mirrors */
enum {
  SSB_MODEL_ILRMA_GAUSS = 0, /* ssspy/bss/ilrma.py:582 GaussILRMA      */
  SSB_MODEL_IVA_LAPLACE = 1, /* ssspy/bss/iva.py:2976 AuxLaplaceIVA    */
  SSB_MODEL_IVA_GAUSS = 2,   /* ssspy/bss/iva.py:3131 AuxGaussIVA      */
  SSB_MODEL_FASTMNMF_GAUSS = 3 /* ssspy/bss/mnmf.py:1076 FastGaussMNMF (n_sources == n_channels): the W slot of
                                  ssb_plan_bind holds the diagonaliser Q[B,I,N,N] c64 (mnmf.py:557-559), the
                                  `variance` slot the spatial property D[B,I,N,N] f32 (mnmf.py:594-596);
                                  `spatial` selects the diagonaliser algorithm IP1 / IP2 (mnmf.py:1430-1447);
                                  ssb_plan_separate is the multichannel Wiener filter (mnmf.py:1174-1217) */
  ,
  SSB_MODEL_ILRMA_T = 4,  /* ssspy/bss/ilrma.py:1992 TILRMA: Student-t source model, model_param = dof nu > 0
                             (MM and ME source updates, IP1 / IP2 / ISS1) */
  SSB_MODEL_ILRMA_GGD = 5 /* ssspy/bss/ilrma.py:3337 GGDILRMA: generalised Gaussian, model_param = beta in (0, 2)
                             (MM only, IP1 / IP2 / ISS1) */
  ,
  SSB_MODEL_FDICA_LAPLACE = 6 /* ssspy/bss/fdica.py:1527 AuxLaplaceFDICA: per-bin weights 2 / floor(2 |y|), IP1 / IP2
                                 (fdica.py:1065-1245); ssb_plan_permutation_* is its permutation alignment */
};
/* spatial_algorithm (ssspy/bss/ilrma.py:27, ssspy/bss/iva.py:44) */
enum { SSB_SPATIAL_IP1 = 0, SSB_SPATIAL_IP2 = 1, SSB_SPATIAL_ISS1 = 2,
       SSB_SPATIAL_ISS2 = 3 /* pairwise ISS, ssspy/bss/_update_spatial_model.py:197-314; uses `pairs` */,
       SSB_SPATIAL_IPA = 4 /* iterative projection with adjustment, _update_spatial_model.py:398-513; state in Y like
                              ISS; uses `ipa_normalization` / `ipa_newton_iter` (GaussILRMA and AuxIVA only) */ };
/* source_algorithm (ssspy/bss/ilrma.py:28) */
enum { SSB_SOURCE_MM = 0, SSB_SOURCE_ME = 1 };
/* sub-steps of the ILRMA source model for ssb_update_source_part */
enum { SSB_PART_LATENT = 0, SSB_PART_BASIS = 1, SSB_PART_ACTIVATION = 2 };
/* flooring_fn (ssspy/special/flooring.py:6-18): max(x,eps) | x+eps | identity */
enum { SSB_FLOOR_MAX = 0, SSB_FLOOR_ADD = 1, SSB_FLOOR_NONE = 2 };
/* normalization (ssspy/bss/ilrma.py:333-363) */
enum { SSB_NORM_NONE = 0, SSB_NORM_POWER = 1, SSB_NORM_PROJECTION_BACK = 2 };

#define SSB_MAX_SOURCES 8
#define SSB_MAX_BASIS 64
#define SSB_MAX_PAIRS 64

typedef struct ssb_config {
  int32_t model;
  int32_t spatial;
  int32_t source;
  int32_t n_batch;   /* B: independent mixtures (extension; the reference has B = 1) */
  int32_t n_sources; /* N = n_channels, 2..SSB_MAX_SOURCES (determined case, ilrma.py:180-181) */
  int32_t n_bins;    /* I */
  int32_t n_frames;  /* J */
  int32_t n_basis;   /* K (ILRMA only), 1..SSB_MAX_BASIS */
  float domain;      /* p in (0, 2] (ilrma.py:786) */
  int32_t flooring;
  float eps;         /* 1e-10 in the reference (ssspy/special/flooring.py:3) */
  int32_t normalization;
  int32_t reference_id; /* projection-back reference channel */
  int32_t n_pairs;      /* IP2: number of (m, n) pairs per iteration */
  int32_t pairs[2 * SSB_MAX_PAIRS]; /* host copy of pair_selector(N), already wrapped into [0, N)
                                       (ssspy/utils/select_pair.py:35-44; negative indices wrap as
                                       NumPy indexing does, tests/package/bss/test_update_spatial_model.py:19-24) */
  int32_t fast_path; /* 1: use the fused sm_100a kernels where the configuration allows (default),
                        0: always the modular kernels */
  float model_param; /* TILRMA: degree of freedom nu (ilrma.py:2160); GGDILRMA: shape beta (ilrma.py:3496);
                        ignored by the other models */
  int32_t partitioning; /* ILRMA family only, 1: partitioning function (ilrma.py:201-245): T is [B,I,K] and V is
                           [B,K,J], shared by the sources, and the `variance` slot of ssb_plan_bind holds the
                           latent variable Z[B,N,K] f32 (ilrma.py:219-226); power normalisation only */
  int32_t ipa_normalization; /* IPA: lqpqm_normalization (ilrma.py:749, iva.py:1579; default 1) */
  int32_t ipa_newton_iter;   /* IPA: newton_iter, Newton-Raphson updates of the LQPQM root (default 1) */
  int32_t no_whitening;      /* 0 (default): the demixing-filter modes (IP1 / IP2; FastGaussMNMF diagonaliser) iterate in
                                the whitened domain z = L^-1 x, C = mean_j x x^H = L L^H (fp64), W~ = W L: the updates
                                of _update_spatial_model.py:63-76, :317-395 are equivariant under this change of basis
                                and y = W x = W~ z, but complex64 rounding is no longer amplified by cond(C).  W stays
                                the reference's demix_filter at the boundary.  1: iterate on X, W directly (A/B, tests) */
} ssb_config;

typedef struct ssb_plan ssb_plan;

const char* ssb_last_error(void);
int ssb_version(void);
/* number of visible CUDA devices (0 => every compute call fails loudly) */
int ssb_device_count(int* count);

/* Device status flags raised by kernels since the last fetch on the current device; reads and clears them after the
 * work enqueued on `stream` (this call synchronises the stream).  SSB_STATUS_SINGULAR: a pivoting solve / inverse met an
 * exactly zero pivot, the case in which the reference raises numpy.linalg.LinAlgError("Singular matrix")
 * (ssspy/linalg/_solve.py:15, ssspy/algorithm/projection_back.py:89, :110); the host classes re-raise it. */
#define SSB_STATUS_SINGULAR 1
int ssb_status_fetch(int* flags /* host */, void* stream);

/* launch accounting: total kernel launches issued by this library in the process */
int ssb_launch_count(unsigned long long* count);
/* per-kernel device timing: events are recorded after every launch between begin and end;
 * ssb_profile_end synchronises and writes "kernel_name launches total_ms" lines into buf */
int ssb_profile_begin(void* stream);
int ssb_profile_end(char* buf, size_t buf_bytes);

/* ---- plan: one bound problem instance = the state of one separator object ------------------ */
int ssb_plan_create(const ssb_config* cfg, ssb_plan** plan);
int ssb_plan_destroy(ssb_plan* plan);
/* bytes of scratch the plan needs (caller allocates, e.g. a torch uint8 tensor) */
int ssb_plan_workspace_bytes(const ssb_plan* plan, size_t* bytes);
/* Bind caller-owned device buffers.  W may be NULL in ISS modes (state lives in Y,
 * ssspy/bss/ilrma.py:897-898); T, V NULL for IVA; variance[B,N,J] f32 only for IVA_GAUSS
 * (ssspy/bss/iva.py:3317).  workspace must hold ssb_plan_workspace_bytes() bytes. */
int ssb_plan_bind(ssb_plan* plan, const void* X, void* W, void* Y, void* T, void* V, void* variance,
                  void* workspace, size_t workspace_bytes);
/* change the flooring policy of later calls (update_once(flooring_fn=...), ilrma.py:900-922) */
int ssb_plan_set_flooring(ssb_plan* plan, int flooring, float eps);
/* one-time work after bind or after X changed (per-bin unweighted covariance used by the power
 * normalisation) */
int ssb_plan_prepare(ssb_plan* plan, void* stream);

/* GaussILRMA.update_once (ilrma.py:900-922) / AuxIVA.update_once (iva.py:1699-1734,3319-3337) */
int ssb_update_once(ssb_plan* plan, void* stream);
/* n_iter x update_once; if loss != NULL (double[(n_iter)*B], device) the loss after every
 * iteration is written there (ssspy/bss/base.py:68-73) */
int ssb_run(ssb_plan* plan, int n_iter, double* loss, void* stream);
/* update_source_model: ilrma.py:924-978 (MM/ME basis+activation); iva.py:3465-3473 (variance) */
int ssb_update_source_model(ssb_plan* plan, void* stream);
/* one sub-step of it on its own (ILRMA family): update_latent_{mm,me} ilrma.py:1007-1049, :1206-1247;
 * update_basis_{mm,me} :1051-1128, :1249-1325; update_activation_{mm,me} :1130-1204, :1327-1401.
 * part = SSB_PART_*; the MM / ME rule is the plan's `source` */
int ssb_update_source_part(ssb_plan* plan, int part, void* stream);
/* update_spatial_model: ilrma.py:1403-1438 (IP1/IP2/ISS1); AuxIVA update_once_{ip1,ip2,iss1}
 * iva.py:1736-1966 */
int ssb_update_spatial_model(ssb_plan* plan, void* stream);
/* normalize: ilrma.py:333-514 */
int ssb_normalize(ssb_plan* plan, void* stream);
/* compute_loss: ilrma.py:1910-1967, iva.py:200-222,2177-2192; loss is double[B] (device) */
int ssb_compute_loss(ssb_plan* plan, double* loss, void* stream);
/* restore_scale / apply_projection_back: ilrma.py:538-565,1969-1979; iva.py:238-267,2194-2204.
 * W-modes: W <- projection_back(W), Y <- W X.  ISS modes: Y <- projection_back(Y, X). */
int ssb_restore_scale(ssb_plan* plan, void* stream);
/* restore_scale with scale_restoration="minimal_distortion_principle" (ilrma.py:567-579,1981-1989; iva.py:269-281,
 * 2206-2214): Y <- mdp(Y | W X, X); in W modes W is refitted as Y X^H (X X^H)^-1 */
int ssb_restore_scale_mdp(ssb_plan* plan, void* stream);
/* Y <- W X with the plan's current W (ilrma.py:272-295); no-op in ISS modes */
int ssb_plan_separate(ssb_plan* plan, void* stream);

/* ---- standalone batched operators (no plan) -------------------------------------------------- */
/* separate: Y[b,n,i,j] = sum_m W[b,i,n,m] X[b,m,i,j]   (ilrma.py:272-295, iva.py:171-194) */
int ssb_separate(const void* X, const void* W, void* Y, int B, int N, int I, int J, void* stream);
/* weighted covariance U[b,i,s,:,:] = (1/J) sum_j phi[...] x x^H for the n_src sources listed in
 * src (host int32[n_src], NULL = 0..n_src-1).  phi element (b, src[s], i, j) is read at
 * phi[b*phi_sb + src[s]*phi_sn + i*phi_si + j] (phi_si = 0 for IVA weights [N,J])
 * (ilrma.py:1500-1505, iva.py:1785-1791).  U is c64 [B,I,n_src,N,N]. */
int ssb_weighted_covariance(const void* X, const float* phi, long long phi_sb, long long phi_sn,
                            long long phi_si, const int32_t* src, int n_src, void* U, int B, int N,
                            int I, int J, void* stream);
/* update_by_ip1 (ssspy/bss/_update_spatial_model.py:17-78): W[n_mat,N,N] c64 in place,
 * U[n_mat,N,N,N] c64 */
int ssb_update_by_ip1(void* W, const void* U, int n_mat, int N, int flooring, float eps, void* stream);
/* update_by_ip2 (_update_spatial_model.py:81-143,317-395): pairs = host int32[2*n_pairs] in [0,N);
 * U[n_mat,N,N,N] c64 holds all sources */
int ssb_update_by_ip2(void* W, const void* U, int n_mat, int N, const int32_t* pairs, int n_pairs,
                      int flooring, float eps, void* stream);
/* update_by_ip2_one_pair (_update_spatial_model.py:317-395): U_pair[n_mat,2,N,N] */
int ssb_update_by_ip2_one_pair(void* W, const void* U_pair, int n_mat, int N, int m, int n, int flooring,
                               float eps, void* stream);
/* update_by_iss1 (_update_spatial_model.py:146-194): Y[B,N,I,J] c64 in place; phi addressed as in
 * ssb_weighted_covariance */
int ssb_update_by_iss1(void* Y, const float* phi, long long phi_sb, long long phi_sn, long long phi_si,
                       int B, int N, int I, int J, int flooring, float eps, void* stream);
/* update_by_iss2 (_update_spatial_model.py:197-314): pairwise ISS on Y[B,N,I,J] in place; pairs[2*n_pairs] on the host,
 * indices already wrapped into [0, N) (the reference's default is (0,1),(2,3),..., :233-234) */
int ssb_update_by_iss2(void* Y, const float* phi, long long phi_sb, long long phi_sn, long long phi_si, int B, int N,
                       int I, int J, const int32_t* pairs, int n_pairs, int flooring, float eps, void* stream);
/* update_by_ipa (_update_spatial_model.py:398-513, ssspy/linalg/lqpqm.py:13-292): iterative projection with
 * adjustment on Y[B,N,I,J] in place; phi addressed as in ssb_weighted_covariance; normalization / max_iter are the
 * reference's `normalization` and `max_iter` arguments */
int ssb_update_by_ipa(void* Y, const float* phi, long long phi_sb, long long phi_sn, long long phi_si, int B, int N,
                      int I, int J, int normalization, int max_iter, int flooring, float eps, void* stream);
/* correlation-based permutation solver (ssspy/algorithm/permutation_alignment.py:12-121), two phases around the
 * host-side argsort of the per-bin correlations (numpy.argsort on float64, as the reference):
 *   ssb_permutation_correlation: corr[B,I] f64 = sum_j (sum_n P_n)^2, P = |Y| / floor(norm over the sources)
 *   ssb_permutation_align: visits the bins of mixture b in order[b, 0..I-1] (int32, device), permutes Y[B,N,I,J] (and
 *   the rows of W[B,I,N,N] when W != NULL) in place and writes the chosen permutations perms[B,I,N] (int32).
 * The plan forms operate on the bound Y = W X (FDICA: recomputed first) and W. */
int ssb_permutation_correlation(const void* Y, double* corr, int B, int N, int I, int J, int flooring, float eps,
                                void* stream);
int ssb_permutation_align(void* Y, void* W, const int32_t* order, int32_t* perms, int B, int N, int I, int J,
                          int flooring, float eps, void* stream);
int ssb_plan_permutation_correlation(ssb_plan* plan, double* corr, void* stream);
int ssb_plan_permutation_align(ssb_plan* plan, const int32_t* order, int32_t* perms, void* stream);
/* projection_back, filter form (ssspy/algorithm/projection_back.py:87-99):
 * Wout[m,n,:] = W[m,n,:] * (W[m]^-1)[ref, n]; Wout may alias W. */
int ssb_projection_back_w(const void* W, void* Wout, int n_mat, int N, int reference_id, void* stream);
/* projection_back, spectrogram form (projection_back.py:100-121): scale[b,i,c,n] =
 * (X Y^H (Y Y^H)^-1)[c,n]; Yout[b,n,i,:] = Y[b,n,i,:] * scale[b,i,ref,n]; Yout may alias Y.
 * scale_out (c64 [B,I,N,N], REQUIRED: it is also the kernel's scratch) receives the full scale matrix
 * (the reference_id=None case reads it); Yout may be NULL to compute only the scale. */
int ssb_projection_back_y(const void* Y, const void* X, void* Yout, void* scale_out, int B, int N, int I,
                          int J, int reference_id, void* stream);

/* minimal_distortion_principle (ssspy/algorithm/minimal_distortion_principle.py:6-43): Yout[b,n,i,:] =
 * conj(z) Y[b,n,i,:], z = sum_j Y conj(X[b,ref,i,:]) / sum_j |Y|^2; Yout may alias Y */
int ssb_minimal_distortion_principle(const void* Y, const void* X, void* Yout, int B, int N, int I, int J,
                                     int reference_id, void* stream);

/* ---- ssspy.linalg helpers, batched over n_mat small matrices, complex128 in/out -------------- */
/* ILRMABase.compute_logdet (ilrma.py:524-536): out[m] = log|det W_m| (double, device) for n_mat complex64 N x N
 * matrices */
int ssb_logdet(const void* W, double* out, int n_mat, int N, void* stream);
/* ILRMABase.reconstruct_nmf (ilrma.py:297-328): R[b,n,i,j] = sum_k T[b,n,i,k] V[b,n,k,j] (Z == NULL), or with the
 * partitioning function sum_k Z[b,n,k] T[b,i,k] V[b,k,j]; all float32 on the device */
int ssb_reconstruct_nmf(const float* T, const float* V, const float* Z, float* R, int B, int N, int I, int J, int K,
                        void* stream);
/* inv2 (ssspy/linalg/inv.py:4-54) generalised to N x N (np.linalg.inv call sites:
 * projection_back.py:89,110; ilrma.py:495,504,1944) */
int ssb_inv(const void* A, void* Ainv, int n_mat, int N, void* stream);
/* solve (ssspy/linalg/_solve.py:9-21): A[n_mat,N,N], B[n_mat,N,R] -> X[n_mat,N,R] */
int ssb_solve(const void* A, const void* B, void* X, int n_mat, int N, int R, void* stream);
/* eigh / eigh2 (ssspy/linalg/eigh.py:8-207): Hermitian A (and positive definite B unless NULL),
 * type 1: A z = l B z, 2: A B z = l z, 3: B A z = l z; ascending eigenvalues lamb[n_mat,N] f64,
 * eigenvectors in the columns of Z[n_mat,N,N] c128 */
int ssb_eigh(const void* A, const void* B, int type, double* lamb, void* Z, int n_mat, int N,
             void* stream);
/* cbrt (ssspy/linalg/cubic.py:4-22) of n complex128 values: cbrt(|x|) exp(i arg(x) / 3) */
int ssb_cbrt(const void* x, void* y, long long n, void* stream);
/* solve_cubic (ssspy/linalg/polynomial.py:9-104): the three roots of x^3 + A x^2 + B x + C for n coefficient triples
 * (complex128), roots[3][n] in the reference's order (Cardano; P == 0 handled as in :82-92) */
int ssb_solve_cubic(const void* A, const void* B, const void* C, void* roots, long long n, void* stream);
/* lqpqm2 (ssspy/linalg/lqpqm.py:13-119): H[n_bins,M,M] c128 (positive semidefinite), v[n_bins,M] c128, z[n_bins] f64
 * -> y[n_bins,M] c128; flooring / eps as in ssb_config; singular_mode 0: ||v|| < flooring(0) (the reference's
 * default "flooring"), 1: ||v|| == 0 (singular_fn=None); max_iter Newton-Raphson steps (:122-219) */
int ssb_lqpqm2(const void* H, const void* v, const double* z, void* y, int n_bins, int M, int flooring, double eps,
               int singular_mode, int max_iter, void* stream);

/* ---- STFT front / back end (not part of ssspy: its notebooks call scipy.signal.stft / istft with window="hann",
 * nperseg=n_fft, noverlap=n_fft-hop, e.g. notebooks/BSS/ILRMA/GaussILRMA-IP1-MM.ipynb; same conventions here:
 * boundary="zeros", padded=True, one-sided, scaling="spectrum") ---- */
/* number of frames scipy.signal.stft produces for n_samples samples */
int ssb_stft_frames(long long n_samples, int nperseg, int hop, int* n_frames);
/* x[n_rows, n_samples] f64, window[nperseg] f64 (device), window_sum = sum(window) -> Z[n_rows, nperseg/2+1, n_frames]
 * c128; nperseg a power of two in [16, 8192] */
int ssb_stft(const double* x, const double* window, double window_sum, void* Z, int n_rows, long long n_samples,
             int nperseg, int hop, void* stream);
/* Z[n_rows, nperseg/2+1, n_frames] c128 -> y[n_rows, nperseg + (n_frames-1) hop - 2 (nperseg/2)] f64 by weighted
 * overlap-add; seg[n_rows, n_frames, nperseg] f64 is scratch */
int ssb_istft(const void* Z, const double* window, double window_sum, double* y, double* seg, int n_rows, int n_frames,
              int nperseg, int hop, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SSB_H_ */
