// C-ABI of libssb.so: error plumbing, the plan object (= device-side state of one separator) and the
// per-iteration orchestration that mirrors GaussILRMA.update_once / AuxIVA.update_once.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <new>

#include "ssb_fused.h"
#include "ssb_kernels.h"

// ---- errors -------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

void ssb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---- launch accounting + optional per-kernel timing ----------------------------------------------
// Every kernel launch of the library goes through ssb_check_launch.  With profiling enabled an
// event is recorded after each launch; kernels of one stream run back to back, so the gap between
// consecutive events is that launch's device time (the first gap starts at ssb_profile_begin).
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>
static std::atomic<unsigned long long> g_launches{0};
static std::atomic<bool> g_prof_on{false};
static std::mutex g_prof_mu;  // guards g_prof_events / g_prof_start (launches may come from several host threads)
static std::vector<std::pair<const char*, cudaEvent_t>> g_prof_events;
static cudaEvent_t g_prof_start = nullptr;

int ssb_check_launch(const char* what, cudaStream_t st) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    ssb_set_error("CUDA launch of '%s' failed: %s", what, cudaGetErrorString(e));
    return 1;
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (g_prof_on.load(std::memory_order_relaxed)) {
    cudaEvent_t ev;
    if (cudaEventCreate(&ev) == cudaSuccess) {
      cudaEventRecord(ev, st);
      std::lock_guard<std::mutex> lk(g_prof_mu);
      g_prof_events.push_back({what, ev});
    }
  }
  return 0;
}

extern "C" int ssb_launch_count(unsigned long long* count) {
  *count = g_launches.load();
  return 0;
}

extern "C" int ssb_profile_begin(void* stream) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& pe : g_prof_events) cudaEventDestroy(pe.second);
  g_prof_events.clear();
  if (!g_prof_start) SSB_CUDA(cudaEventCreate(&g_prof_start));
  SSB_CUDA(cudaEventRecord(g_prof_start, (cudaStream_t)stream));
  g_prof_on = true;
  return 0;
}

// Stops profiling, synchronises, and writes "name count total_ms\n" lines (sorted by total time).
extern "C" int ssb_profile_end(char* buf, size_t buf_bytes) {
  g_prof_on = false;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  std::map<std::string, std::pair<int, double>> agg;
  cudaEvent_t prev = g_prof_start;
  for (auto& pe : g_prof_events) {
    SSB_CUDA(cudaEventSynchronize(pe.second));
    float ms = 0.f;
    SSB_CUDA(cudaEventElapsedTime(&ms, prev, pe.second));
    auto& a = agg[pe.first];
    a.first += 1;
    a.second += ms;
    prev = pe.second;
  }
  for (auto& pe : g_prof_events) cudaEventDestroy(pe.second);
  g_prof_events.clear();
  std::vector<std::pair<double, std::string>> order;
  for (auto& kv : agg) order.push_back({-kv.second.second, kv.first});
  std::sort(order.begin(), order.end());
  size_t off = 0;
  if (buf && buf_bytes) buf[0] = 0;
  for (auto& o : order) {
    auto& a = agg[o.second];
    int n = snprintf(buf + off, buf_bytes > off ? buf_bytes - off : 0, "%s %d %.6f\n", o.second.c_str(), a.first,
                     a.second);
    if (n < 0 || off + n >= buf_bytes) break;
    off += n;
  }
  return 0;
}

// ---- device status word -----------------------------------------------------------------------------------------
static int* g_status_dev[SSB_MAX_DEVICES] = {};
static std::mutex g_status_mu;

int* ssb_status_word() {
  const int dev = ssb_current_device();
  std::lock_guard<std::mutex> lk(g_status_mu);
  if (g_status_dev[dev] == nullptr) {
    int* p = nullptr;
    if (cudaMalloc(&p, sizeof(int)) == cudaSuccess && cudaMemset(p, 0, sizeof(int)) == cudaSuccess) g_status_dev[dev] = p;
    else cudaGetLastError();
  }
  return g_status_dev[dev];  // NULL only when the device cannot allocate 4 bytes: the next launch fails loudly
}

// Reads and clears the status word of the current device after everything enqueued on `stream` (synchronises it).
extern "C" int ssb_status_fetch(int* flags, void* stream) {
  SSB_REQUIRE(flags != nullptr, "flags output is NULL");
  int* w = ssb_status_word();
  SSB_REQUIRE(w != nullptr, "no device status word (is a CUDA device available?)");
  cudaStream_t st = (cudaStream_t)stream;
  SSB_CUDA(cudaMemcpyAsync(flags, w, sizeof(int), cudaMemcpyDeviceToHost, st));
  SSB_CUDA(cudaMemsetAsync(w, 0, sizeof(int), st));
  SSB_CUDA(cudaStreamSynchronize(st));
  return 0;
}

extern "C" const char* ssb_last_error(void) { return g_err; }
extern "C" int ssb_version(void) { return SSB_VERSION; }
extern "C" int ssb_device_count(int* count) {
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  if (e != cudaSuccess) {
    c = 0;
    cudaGetLastError();
  }
  *count = c;
  return 0;
}

// ---- plan ---------------------------------------------------------------------------------------
struct ssb_plan {
  ssb_config cfg;
  const cf* X = nullptr;
  cf* W = nullptr;
  cf* Y = nullptr;
  float* T = nullptr;
  float* V = nullptr;
  float* variance = nullptr;
  char* ws = nullptr;
  size_t ws_bytes = 0;
  // workspace carve-up
  float* big = nullptr;     // [B,N,I,J] f32: power spectrogram P, later the weights phi (ILRMA)
  float* phi_iva = nullptr; // [B,N,J]
  float* r2 = nullptr;      // [B,N,J]
  cf* U = nullptr;          // [B,I,N,N,N]
  cf* C = nullptr;          // [B,I,N,N]  unweighted covariance (power normalisation)
  cf* S = nullptr;          // [B,I,N,N]  cross-solve result / recovered W
  cf* scale = nullptr;      // [B,I,N]
  double* psi2 = nullptr;   // [B,N]
  double* rowloss = nullptr;  // [B,N,I]
  double* logdet = nullptr;   // [B,I]
  ssb_fused_ws fused;       // extra scratch of the fused kernels
  bool bound = false, prepared = false;
  float* big2 = nullptr;    // [B,N,I,J] f32 second elementwise scratch (FastGaussMNMF: H)
  cd* qinv = nullptr;       // [B,I,N,N] c128 (FastGaussMNMF separate)
  float* big3 = nullptr;    // [B,N,I,J] f32 Lambda = T V (FastGaussMNMF, tensor-core kernel)
  // partitioning function: Teff[B,N,I,K] = z t, raw sums of the basis- / activation-type sweeps
  float* teff = nullptr;
  float *gnum = nullptr, *gden = nullptr;  // [B,N,I,K]
  float *hnum = nullptr, *hden = nullptr;  // [B,N,K,J]
  // whitened-domain iteration (ssb_whiten.cu): the kernels of the demixing-filter modes read Xk = M X and update
  // Wk = W M^-1; X / W stay the caller's (reference-domain) buffers, synchronised at the C-ABI boundary
  const cf* Xk = nullptr;
  cf* Wk = nullptr;
  bool whiten = false;
  cf* Xw = nullptr;       // [B,N,I,J]
  cf* Ww = nullptr;       // [B,I,N,N]
  cf* Wexp = nullptr;     // [B,I,N,N] the W last exported / imported, bit for bit
  cd* C64 = nullptr;      // [B,I,N,N] fp64 covariance of the mixture
  cd* Mw = nullptr;       // [B,I,N,N] whitening matrix M = L^-1 (C = L L^H)
  cd* Mwinv = nullptr;    // [B,I,N,N] M^-1 = L
  double* ldM = nullptr;  // [B,I] log|det M|
  int* wsync = nullptr;   // [B,I] 1: Wk holds the whitened image of Wexp
  // AuxIVA-ISS1: r2[b,n,j] = sum_i |y|^2 comes out of the apply sweep of the previous iteration (ssb_spatial.cu
  // k_iss1_cov); trusted only inside ssb_run, where nothing else touches Y between two iterations
  float* r2part = nullptr;  // [B, groups, N, J]
  bool r2_valid = false;
  // FastGaussMNMF inside ssb_run: km_spatial leaves Z2 = |Q x|^2 of the new filters in `big` for the next iteration's
  // source model; zscale[b, m] = 1 / psi_m^2 when the power normalisation rescaled Q after it was written
  bool in_run = false, z2_valid = false, z2_scaled = false;
  float* zscale = nullptr;
  bool part() const { return cfg.partitioning != 0; }
  bool mnmf() const { return cfg.model == SSB_MODEL_FASTMNMF_GAUSS; }
  // modes whose state lives in Y (no demixing filter): ISS1 / ISS2 / IPA
  bool iss() const {
    return (cfg.spatial == SSB_SPATIAL_ISS1 || cfg.spatial == SSB_SPATIAL_ISS2 || cfg.spatial == SSB_SPATIAL_IPA) &&
           !mnmf();
  }
  bool fdica() const { return cfg.model == SSB_MODEL_FDICA_LAPLACE; }
  bool ilrma() const {
    return cfg.model == SSB_MODEL_ILRMA_GAUSS || cfg.model == SSB_MODEL_ILRMA_T || cfg.model == SSB_MODEL_ILRMA_GGD;
  }
};

namespace {

size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

struct Carver {
  char* base;
  size_t off = 0;
  template <typename T>
  T* take(size_t n) {
    T* p = base ? (T*)(base + off) : nullptr;
    off += align_up(n * sizeof(T));
    return p;
  }
};

size_t carve(ssb_plan* p, char* base) {
  const ssb_config& c = p->cfg;
  const size_t B = c.n_batch, N = c.n_sources, I = c.n_bins, J = c.n_frames;
  Carver cv{base};
  // bins rounded up to 16 (the cooperative kernels keep P in 16 x 16 tiles) + 16 rows of slack: the per-source
  // activation kernel prefetches whole 16-bin tiles of P without clamping
  p->big = (p->ilrma() || p->mnmf() || p->fdica()) ? cv.take<float>(B * N * ((I + 15) / 16 * 16) * J + 16 * J) : nullptr;
  p->big2 = p->mnmf() ? cv.take<float>(B * N * I * J) : nullptr;
  p->qinv = p->mnmf() ? cv.take<cd>(B * I * N * N) : nullptr;
  p->big3 = p->mnmf() ? cv.take<float>(B * N * I * J) : nullptr;
  p->zscale = p->mnmf() ? cv.take<float>(B * N) : nullptr;
  p->phi_iva = cv.take<float>(B * N * J);
  p->r2 = cv.take<float>(B * N * J);
  const bool iva = c.model == SSB_MODEL_IVA_LAPLACE || c.model == SSB_MODEL_IVA_GAUSS;
  p->r2part = (iva && c.spatial == SSB_SPATIAL_ISS1 && ssbk_iss1_emits_r2(c.n_sources, c.n_frames))
                  ? cv.take<float>(B * (size_t)ssbk_iss1_r2_groups(c.n_bins) * N * J)
                  : nullptr;
  p->U = cv.take<cf>(B * I * N * N * N);
  p->C = cv.take<cf>(B * I * N * N);
  p->S = cv.take<cf>(B * I * N * N);
  p->scale = cv.take<cf>(B * I * N);
  p->psi2 = cv.take<double>(B * N);
  p->rowloss = cv.take<double>(B * N * I);
  p->logdet = cv.take<double>(B * I);
  p->whiten = !p->iss() && !c.no_whitening;
  if (p->whiten) {
    p->Xw = cv.take<cf>(B * N * I * J);
    p->Ww = cv.take<cf>(B * I * N * N);
    p->Wexp = cv.take<cf>(B * I * N * N);
    p->C64 = cv.take<cd>(B * I * N * N);
    p->Mw = cv.take<cd>(B * I * N * N);
    p->Mwinv = cv.take<cd>(B * I * N * N);
    p->ldM = cv.take<double>(B * I);
    p->wsync = cv.take<int>(B * I);
  }
  if (p->part()) {
    const size_t K = c.n_basis;
    p->teff = cv.take<float>(B * N * I * K);
    p->gnum = cv.take<float>(B * N * I * K);
    p->gden = cv.take<float>(B * N * I * K);
    p->hnum = cv.take<float>(B * N * K * J);
    p->hden = cv.take<float>(B * N * K * J);
  }
  cv.off += ssb_fused_carve(&p->fused, &c, base ? base + cv.off : nullptr);
  return cv.off;
}

int validate(const ssb_config* c) {
  SSB_REQUIRE(c != nullptr, "config is NULL");
  SSB_REQUIRE(c->model >= 0 && c->model <= 6, "unknown model %d", c->model);
  if (c->model == SSB_MODEL_FDICA_LAPLACE)
    SSB_REQUIRE(c->spatial == SSB_SPATIAL_IP1 || c->spatial == SSB_SPATIAL_IP2, "Not support spatial algorithm id %d.",
                c->spatial);
  SSB_REQUIRE(c->spatial >= 0 && c->spatial <= 4, "Not support spatial algorithm id %d.", c->spatial);
  if (c->spatial == SSB_SPATIAL_IPA) {
    SSB_REQUIRE(c->model != SSB_MODEL_ILRMA_T, "IPA is not supported for t-ILRMA.");
    SSB_REQUIRE(c->model != SSB_MODEL_ILRMA_GGD, "IPA is not supported for GGD-ILRMA.");
    SSB_REQUIRE(c->ipa_newton_iter >= 0, "newton_iter=%d must be non-negative", c->ipa_newton_iter);
  }
  SSB_REQUIRE(c->source == SSB_SOURCE_MM || c->source == SSB_SOURCE_ME, "Not support source algorithm id %d.",
              c->source);
  SSB_REQUIRE(c->n_batch >= 1, "n_batch must be >= 1");
  SSB_REQUIRE(c->n_sources >= 2 && c->n_sources <= SSB_MAX_SOURCES, "n_sources=%d unsupported (2..%d)", c->n_sources,
              SSB_MAX_SOURCES);
  SSB_REQUIRE(c->n_bins >= 1 && c->n_frames >= 1, "empty input (n_bins=%d, n_frames=%d)", c->n_bins, c->n_frames);
  if (c->model == SSB_MODEL_FASTMNMF_GAUSS) {
    SSB_REQUIRE(c->n_basis >= 1 && c->n_basis <= SSB_MAX_BASIS, "n_basis=%d unsupported (1..%d)", c->n_basis,
                SSB_MAX_BASIS);
    SSB_REQUIRE(c->spatial == SSB_SPATIAL_IP1 || c->spatial == SSB_SPATIAL_IP2, "Not support diagonalizer algorithm id %d.",
                c->spatial);
    SSB_REQUIRE(c->normalization == SSB_NORM_NONE || c->normalization == SSB_NORM_POWER,
                "Normalization %d is not implemented.", c->normalization);
  }
  if (c->model == SSB_MODEL_ILRMA_T) SSB_REQUIRE(c->model_param > 0.f, "dof must be positive (got %g)", c->model_param);
  if (c->model == SSB_MODEL_ILRMA_GGD) {
    SSB_REQUIRE(c->model_param > 0.f && c->model_param < 2.f, "Shape parameter 2 shoule be chosen from (0, 2).");
    SSB_REQUIRE(c->source == SSB_SOURCE_MM, "Not support source algorithm id %d for GGDILRMA.", c->source);
  }
  if (c->model == SSB_MODEL_ILRMA_GAUSS || c->model == SSB_MODEL_ILRMA_T || c->model == SSB_MODEL_ILRMA_GGD) {
    SSB_REQUIRE(c->n_basis >= 1 && c->n_basis <= SSB_MAX_BASIS, "n_basis=%d unsupported (1..%d)", c->n_basis,
                SSB_MAX_BASIS);
    SSB_REQUIRE(c->domain > 0.f && c->domain <= 2.f, "domain parameter should be chosen from [0, 2].");
    SSB_REQUIRE(c->source != SSB_SOURCE_ME || c->domain == 2.f,
                "domain parameter should be 2 when you specify ME algorithm.");
    SSB_REQUIRE(c->normalization >= 0 && c->normalization <= 2, "Normalization %d is not implemented.",
                c->normalization);
  }
  if (c->partitioning) {
    SSB_REQUIRE(c->model == SSB_MODEL_ILRMA_GAUSS || c->model == SSB_MODEL_ILRMA_T || c->model == SSB_MODEL_ILRMA_GGD,
                "the partitioning function is defined for the ILRMA family only");
    SSB_REQUIRE(c->normalization != SSB_NORM_PROJECTION_BACK,
                "Projection-back-based normalization is not applicable with partitioning function.");
  }
  SSB_REQUIRE(c->flooring >= 0 && c->flooring <= 2, "unknown flooring mode %d", c->flooring);
  SSB_REQUIRE(c->reference_id >= 0 && c->reference_id < c->n_sources, "reference_id=%d out of range",
              c->reference_id);
  if (c->spatial == SSB_SPATIAL_IP2 || c->spatial == SSB_SPATIAL_ISS2) {
    SSB_REQUIRE(c->n_pairs >= 0 && c->n_pairs <= SSB_MAX_PAIRS, "n_pairs=%d exceeds %d", c->n_pairs, SSB_MAX_PAIRS);
    for (int q = 0; q < c->n_pairs; ++q) {
      int m = c->pairs[2 * q], n = c->pairs[2 * q + 1];
      SSB_REQUIRE(m >= 0 && m < c->n_sources && n >= 0 && n < c->n_sources && m != n,
                  "invalid pair (%d, %d) for n_sources=%d", m, n, c->n_sources);
    }
  }
  return 0;
}

#define TRY(x)            \
  do {                    \
    if (int rc_ = (x)) return rc_; \
  } while (0)

int require_bound(const ssb_plan* p) {
  SSB_REQUIRE(p != nullptr, "plan is NULL");
  SSB_REQUIRE(p->bound, "plan has no buffers bound (call ssb_plan_bind first)");
  return 0;
}

// Whitened-domain boundary (ssb_whiten.cu): before a plan call reads the filters, bins whose caller-visible W changed
// since the last export are re-imported (W~ = W M^-1); after a call that updated them, W = W~ M is written back.
int w_enter(ssb_plan* p, cudaStream_t st) {
  if (!p->whiten) return 0;
  SSB_REQUIRE(p->prepared, "plan not prepared (call ssb_plan_prepare after bind)");
  return ssbk_w_import(p->W, p->Wexp, p->Ww, p->Mwinv, p->wsync, p->cfg.n_batch * p->cfg.n_bins, p->cfg.n_sources, st);
}
int w_exit(ssb_plan* p, cudaStream_t st) {
  if (!p->whiten) return 0;
  return ssbk_w_export(p->Ww, p->Mw, p->W, p->Wexp, p->wsync, p->cfg.n_batch * p->cfg.n_bins, p->cfg.n_sources, st);
}

// P <- |Y|^2 with Y = W X (W modes) or the stored Y (ISS modes)
int power_spectrogram(ssb_plan* p, cudaStream_t st) {
  const ssb_config& c = p->cfg;
  if (p->iss()) return ssbk_abs2(p->Y, p->big, (size_t)c.n_batch * c.n_sources * c.n_bins * c.n_frames, st);
  return ssbk_separate(p->Xk, p->Wk, nullptr, p->big, c.n_batch, c.n_sources, c.n_bins, c.n_frames, st);
}

// log|det W_i| for every (b,i); ISS modes first recover W = Y X^H (X X^H)^-1
int logdets(ssb_plan* p, cudaStream_t st) {
  const ssb_config& c = p->cfg;
  const cf* W = p->Wk;
  if (p->iss()) {
    TRY(ssbk_cross_solve(p->Y, p->Xk, p->S, c.n_batch, c.n_sources, c.n_bins, c.n_frames, st));
    W = p->S;
  }
  TRY(ssbk_logdet(W, p->logdet, c.n_batch * c.n_bins, c.n_sources, st));
  // whitened filters: log|det W| = log|det W~| + log|det M|
  if (p->whiten) TRY(ssbk_add_logdet(p->logdet, p->ldM, c.n_batch * c.n_bins, st));
  return 0;
}

// partitioning function: latent, basis, activation in this order, each from a fresh sweep over P with the
// model R_n = Teff_n V (ilrma.py:972-975)
int ilrma_source_part(ssb_plan* p, cudaStream_t st) {
  const ssb_config& c = p->cfg;
  const int B = c.n_batch, N = c.n_sources, I = c.n_bins, J = c.n_frames, K = c.n_basis;
  float* Z = p->variance;
  TRY(power_spectrogram(p, st));
  TRY(ssbk_part_teff(Z, p->T, p->teff, B, N, I, K, st));
  TRY(ssbk_nmf_basis(p->big, p->teff, p->V, B * N, I, J, K, c.domain, c.source, c.model, c.model_param, c.flooring,
                     c.eps, st, N, p->gnum, p->gden));
  TRY(ssbk_part_latent(p->gnum, p->gden, p->T, Z, B, N, I, K, c.domain, c.source, c.model, c.model_param, st));
  TRY(ssbk_part_teff(Z, p->T, p->teff, B, N, I, K, st));
  TRY(ssbk_nmf_basis(p->big, p->teff, p->V, B * N, I, J, K, c.domain, c.source, c.model, c.model_param, c.flooring,
                     c.eps, st, N, p->gnum, p->gden));
  TRY(ssbk_part_basis(p->gnum, p->gden, Z, p->T, B, N, I, K, c.domain, c.source, c.model, c.model_param, c.flooring,
                      c.eps, st));
  TRY(ssbk_part_teff(Z, p->T, p->teff, B, N, I, K, st));
  TRY(ssbk_nmf_activation(p->big, p->teff, p->V, B * N, I, J, K, c.domain, c.source, c.model, c.model_param,
                          c.flooring, c.eps, st, N, p->hnum, p->hden));
  return ssbk_part_activation(p->hnum, p->hden, p->V, B, N, K, J, c.domain, c.source, c.model, c.model_param,
                              c.flooring, c.eps, st);
}

int ilrma_source(ssb_plan* p, cudaStream_t st) {
  const ssb_config& c = p->cfg;
  const int BN = c.n_batch * c.n_sources;
  if (p->part()) return ilrma_source_part(p, st);
  TRY(power_spectrogram(p, st));
  TRY(ssbk_nmf_basis(p->big, p->T, p->V, BN, c.n_bins, c.n_frames, c.n_basis, c.domain, c.source, c.model,
                     c.model_param, c.flooring, c.eps, st));
  TRY(ssbk_nmf_activation(p->big, p->T, p->V, BN, c.n_bins, c.n_frames, c.n_basis, c.domain, c.source, c.model,
                          c.model_param, c.flooring, c.eps, st));
  return 0;
}

// one sub-step of the source model on its own: the reference's update_latent_* / update_basis_* / update_activation_*
// (ilrma.py:1007-1204, :1206-1401), each starting from a fresh power spectrogram exactly as the reference methods do.
// The kernel sequences are the corresponding slices of ilrma_source / ilrma_source_part.
int ilrma_source_substep(ssb_plan* p, int part, cudaStream_t st) {
  const ssb_config& c = p->cfg;
  const int B = c.n_batch, N = c.n_sources, I = c.n_bins, J = c.n_frames, K = c.n_basis;
  SSB_REQUIRE(part >= SSB_PART_LATENT && part <= SSB_PART_ACTIVATION, "Invalid source-model part %d.", part);
  SSB_REQUIRE(part != SSB_PART_LATENT || p->part(), "The latent variable exists only with partitioning=True.");
  TRY(power_spectrogram(p, st));
  if (p->part()) {
    float* Z = p->variance;
    TRY(ssbk_part_teff(Z, p->T, p->teff, B, N, I, K, st));
    if (part == SSB_PART_ACTIVATION) {
      TRY(ssbk_nmf_activation(p->big, p->teff, p->V, B * N, I, J, K, c.domain, c.source, c.model, c.model_param,
                              c.flooring, c.eps, st, N, p->hnum, p->hden));
      return ssbk_part_activation(p->hnum, p->hden, p->V, B, N, K, J, c.domain, c.source, c.model, c.model_param,
                                  c.flooring, c.eps, st);
    }
    TRY(ssbk_nmf_basis(p->big, p->teff, p->V, B * N, I, J, K, c.domain, c.source, c.model, c.model_param, c.flooring,
                       c.eps, st, N, p->gnum, p->gden));
    if (part == SSB_PART_LATENT)
      return ssbk_part_latent(p->gnum, p->gden, p->T, Z, B, N, I, K, c.domain, c.source, c.model, c.model_param, st);
    return ssbk_part_basis(p->gnum, p->gden, Z, p->T, B, N, I, K, c.domain, c.source, c.model, c.model_param,
                           c.flooring, c.eps, st);
  }
  if (part == SSB_PART_BASIS)
    return ssbk_nmf_basis(p->big, p->T, p->V, B * N, I, J, K, c.domain, c.source, c.model, c.model_param, c.flooring,
                          c.eps, st);
  return ssbk_nmf_activation(p->big, p->T, p->V, B * N, I, J, K, c.domain, c.source, c.model, c.model_param,
                             c.flooring, c.eps, st);
}

int ilrma_spatial(ssb_plan* p, cudaStream_t st) {
  const ssb_config& c = p->cfg;
  const int B = c.n_batch, N = c.n_sources, I = c.n_bins, J = c.n_frames;
  // the Student-t / GGD weights depend on the current power spectrogram (ilrma.py:2920-2934, :3992-4010)
  if (c.model != SSB_MODEL_ILRMA_GAUSS) TRY(power_spectrogram(p, st));
  if (p->part()) TRY(ssbk_part_teff(p->variance, p->T, p->teff, B, N, I, c.n_basis, st));
  TRY(ssbk_nmf_phi(p->part() ? p->teff : p->T, p->V, p->big, p->big, B * N, I, J, c.n_basis, c.domain, c.model,
                   c.model_param, c.flooring, c.eps, st, p->part() ? N : 1));
  const long long sb = (long long)N * I * J, sn = (long long)I * J, si = J;
  if (c.spatial == SSB_SPATIAL_ISS1) return ssbk_iss1(p->Y, p->big, sb, sn, si, B, N, I, J, c.flooring, c.eps, st);
  if (c.spatial == SSB_SPATIAL_ISS2)
    return ssbk_iss2(p->Y, p->big, sb, sn, si, B, N, I, J, c.pairs, c.n_pairs, c.flooring, c.eps, st);
  if (c.spatial == SSB_SPATIAL_IPA)  // ilrma.py:1813-1908
    return ssbk_ipa(p->Y, p->big, sb, sn, si, B, N, I, J, c.ipa_normalization, c.ipa_newton_iter, c.flooring, c.eps, st);
  TRY(ssbk_wcov(p->Xk, p->big, sb, sn, si, nullptr, N, p->U, B, N, I, J, st));
  if (c.spatial == SSB_SPATIAL_IP1) return ssbk_ip1(p->Wk, p->U, B * I, N, c.flooring, c.eps, st);
  return ssbk_ip2(p->Wk, p->U, B * I, N, c.pairs, c.n_pairs, N, nullptr, c.flooring, c.eps, st);
}

int ilrma_normalize(ssb_plan* p, cudaStream_t st) {
  const ssb_config& c = p->cfg;
  const int B = c.n_batch, N = c.n_sources, I = c.n_bins, J = c.n_frames, K = c.n_basis;
  if (c.normalization == SSB_NORM_POWER) {
    // partitioning: psi rescales Z and T through k_part_normalize (ilrma.py:424-430), not T[n] directly
    float* Tn = p->part() ? nullptr : p->T;
    if (p->iss()) {
      TRY(ssbk_psi_from_y(p->Y, p->psi2, B, N, I, J, st));
      TRY(ssbk_apply_psi(p->psi2, Tn, nullptr, p->Y, B, N, I, J, K, c.domain, c.flooring, c.eps, st));
    } else {
      SSB_REQUIRE(p->prepared, "plan not prepared (call ssb_plan_prepare after bind)");
      TRY(ssbk_psi_from_cov(p->Wk, p->C, p->psi2, B, N, I, st));
      TRY(ssbk_apply_psi(p->psi2, Tn, p->Wk, nullptr, B, N, I, J, K, c.domain, c.flooring, c.eps, st));
    }
    if (p->part()) return ssbk_part_normalize(p->psi2, p->variance, p->T, B, N, I, K, c.domain, c.flooring, c.eps, st);
    return 0;
  }
  if (c.normalization == SSB_NORM_PROJECTION_BACK) {
    if (p->iss()) {
      TRY(ssbk_cross_solve(p->Xk, p->Y, p->S, B, N, I, J, st));
      TRY(ssbk_scale_rows(p->Y, p->S, p->Y, B, N, I, J, c.reference_id, st));
      return ssbk_scale_basis(p->T, p->S + (size_t)c.reference_id * N, (long long)N * N, 1, B, N, I, K, c.domain, st);
    }
    if (p->whiten) TRY(ssbk_pb_whitened(p->Wk, p->Mwinv, p->scale, B * I, N, c.reference_id, st));
    else TRY(ssbk_pb_w(p->Wk, p->Wk, p->scale, B * I, N, c.reference_id, st));
    return ssbk_scale_basis(p->T, p->scale, N, 1, B, N, I, K, c.domain, st);
  }
  return 0;
}

int ilrma_loss(ssb_plan* p, double* loss, cudaStream_t st) {
  const ssb_config& c = p->cfg;
  const int B = c.n_batch, N = c.n_sources, I = c.n_bins, J = c.n_frames;
  TRY(power_spectrogram(p, st));
  TRY(logdets(p, st));
  if (p->part()) TRY(ssbk_part_teff(p->variance, p->T, p->teff, B, N, I, c.n_basis, st));
  TRY(ssbk_nmf_rowloss(p->big, p->part() ? p->teff : p->T, p->V, p->rowloss, B * N, I, J, c.n_basis, c.domain, c.model,
                       c.model_param, st, p->part() ? N : 1));
  return ssbk_ilrma_loss_reduce(p->rowloss, p->logdet, loss, B, N, I, st);
}

// r2 over all sources from the current state
int iva_norm_all(ssb_plan* p, cudaStream_t st) {
  const ssb_config& c = p->cfg;
  if (p->r2_valid) return 0;  // r2 of the current Y was emitted by the last ISS1 apply sweep
  return ssbk_iva_norm2(p->Xk, p->iss() ? nullptr : p->Wk, p->Y, nullptr, c.n_sources, p->r2, c.n_batch, c.n_sources,
                        c.n_bins, c.n_frames, st);
}

int iva_source(ssb_plan* p, cudaStream_t st) {
  const ssb_config& c = p->cfg;
  if (c.model != SSB_MODEL_IVA_GAUSS) return 0;
  TRY(iva_norm_all(p, st));
  return ssbk_iva_phi(p->r2, p->variance, 1, nullptr, c.n_sources, nullptr, c.model, c.n_batch, c.n_sources, c.n_bins,
                      c.n_frames, c.flooring, c.eps, st);
}

int iva_spatial(ssb_plan* p, cudaStream_t st) {
  const ssb_config& c = p->cfg;
  const int B = c.n_batch, N = c.n_sources, I = c.n_bins, J = c.n_frames;
  if (c.spatial == SSB_SPATIAL_IP2) {
    for (int q = 0; q < c.n_pairs; ++q) {
      const int pr[2] = {c.pairs[2 * q], c.pairs[2 * q + 1]};
      const int uidx[2] = {0, 1};
      TRY(ssbk_iva_norm2(p->Xk, p->Wk, nullptr, pr, 2, p->r2, B, N, I, J, st));
      TRY(ssbk_iva_phi(p->r2, p->variance, 0, pr, 2, p->phi_iva, c.model, B, N, I, J, c.flooring, c.eps, st));
      if (c.fast_path && (J % 16) == 0) TRY(ssb_fused_cov_w(p->Xk, p->phi_iva, 2LL * J, J, 0, 2, p->U, B, N, I, J, st));
      else TRY(ssbk_wcov(p->Xk, p->phi_iva, 2LL * J, J, 0, nullptr, 2, p->U, B, N, I, J, st));
      TRY(ssbk_ip2(p->Wk, p->U, B * I, N, pr, 1, 2, uidx, c.flooring, c.eps, st));
    }
    return 0;
  }
  TRY(iva_norm_all(p, st));
  TRY(ssbk_iva_phi(p->r2, p->variance, 0, nullptr, N, p->phi_iva, c.model, B, N, I, J, c.flooring, c.eps, st));
  if (c.spatial == SSB_SPATIAL_ISS1) {
    TRY(ssbk_iss1(p->Y, p->phi_iva, (long long)N * J, J, 0, B, N, I, J, c.flooring, c.eps, st, p->r2part, p->r2));
    p->r2_valid = p->r2part != nullptr;
    return 0;
  }
  if (c.spatial == SSB_SPATIAL_ISS2)  // iva.py:1968-2066: weights once, then every pair
    return ssbk_iss2(p->Y, p->phi_iva, (long long)N * J, J, 0, B, N, I, J, c.pairs, c.n_pairs, c.flooring, c.eps, st);
  if (c.spatial == SSB_SPATIAL_IPA)  // iva.py:2068-2176
    return ssbk_ipa(p->Y, p->phi_iva, (long long)N * J, J, 0, B, N, I, J, c.ipa_normalization, c.ipa_newton_iter,
                    c.flooring, c.eps, st);
  if (c.fast_path && (J % 16) == 0) TRY(ssb_fused_cov_w(p->Xk, p->phi_iva, (long long)N * J, J, 0, N, p->U, B, N, I, J, st));
  else TRY(ssbk_wcov(p->Xk, p->phi_iva, (long long)N * J, J, 0, nullptr, N, p->U, B, N, I, J, st));
  return ssbk_ip1(p->Wk, p->U, B * I, N, c.flooring, c.eps, st);
}

int iva_loss(ssb_plan* p, double* loss, cudaStream_t st) {
  const ssb_config& c = p->cfg;
  TRY(iva_norm_all(p, st));
  TRY(logdets(p, st));
  return ssbk_iva_loss(p->r2, p->variance, p->logdet, loss, c.model, c.n_batch, c.n_sources, c.n_bins, c.n_frames, st);
}

// ---- AuxLaplaceFDICA (fdica.py:1065-1245): per-bin weights, the IP kernels of the IVA path ------------------------
int fdica_spatial(ssb_plan* p, cudaStream_t st) {
  const ssb_config& c = p->cfg;
  const int B = c.n_batch, N = c.n_sources, I = c.n_bins, J = c.n_frames;
  const bool fast = c.fast_path && (J % 16) == 0;
  if (c.spatial == SSB_SPATIAL_IP2) {  // weights recomputed for every pair with the current W (fdica.py:1218-1243)
    for (int q = 0; q < c.n_pairs; ++q) {
      const int pr[2] = {c.pairs[2 * q], c.pairs[2 * q + 1]};
      const int uidx[2] = {0, 1};
      TRY(ssbk_fdica_phi(p->Xk, p->Wk, p->big, pr, 2, B, N, I, J, c.flooring, c.eps, st));
      if (fast) TRY(ssb_fused_cov_w(p->Xk, p->big, 2LL * I * J, (long long)I * J, J, 2, p->U, B, N, I, J, st));
      else TRY(ssbk_wcov(p->Xk, p->big, 2LL * I * J, (long long)I * J, J, nullptr, 2, p->U, B, N, I, J, st));
      TRY(ssbk_ip2(p->Wk, p->U, B * I, N, pr, 1, 2, uidx, c.flooring, c.eps, st));
    }
    return 0;
  }
  TRY(ssbk_fdica_phi(p->Xk, p->Wk, p->big, nullptr, N, B, N, I, J, c.flooring, c.eps, st));
  if (fast) TRY(ssb_fused_cov_w(p->Xk, p->big, (long long)N * I * J, (long long)I * J, J, N, p->U, B, N, I, J, st));
  else TRY(ssbk_wcov(p->Xk, p->big, (long long)N * I * J, (long long)I * J, J, nullptr, N, p->U, B, N, I, J, st));
  return ssbk_ip1(p->Wk, p->U, B * I, N, c.flooring, c.eps, st);
}

int fdica_loss(ssb_plan* p, double* loss, cudaStream_t st) {
  const ssb_config& c = p->cfg;
  TRY(ssbk_fdica_rowloss(p->Xk, p->Wk, p->rowloss, c.n_batch, c.n_sources, c.n_bins, c.n_frames, st));
  TRY(ssbk_logdet(p->Wk, p->logdet, c.n_batch * c.n_bins, c.n_sources, st));
  if (p->whiten) TRY(ssbk_add_logdet(p->logdet, p->ldM, c.n_batch * c.n_bins, st));
  return ssbk_ilrma_loss_reduce(p->rowloss, p->logdet, loss, c.n_batch, 1, c.n_bins, st);
}

// ---- FastGaussMNMF: W slot = diagonaliser Q[B,I,N,N] c64, variance slot = spatial D[B,I,N,N] f32 ----------
// Lambda = T V on the tensor pipe when the fused kernel covers the shape, else NULL (the consumers then
// contract over K themselves)
int mnmf_lambda(ssb_plan* p, const float** lam, cudaStream_t st) {
  const ssb_config& c = p->cfg;
  *lam = nullptr;
  if (c.fast_path && c.n_basis <= 32 && (c.n_frames % 16) == 0) {
    TRY(ssb_fused_phi(&c, p->T, p->V, p->big3, 0, st));
    *lam = p->big3;
  }
  return 0;
}

// four sources, K <= 16: G / H and Lambda are formed inside the update kernels from Z2 = |Q x|^2 and D; no Lambda, G, H
// arrays at all.  SSB_MNMF_FUSED=0 (read once) keeps the array path.
bool mnmf_fused_source(const ssb_plan* p) {
  static const int fused_src = getenv("SSB_MNMF_FUSED") != nullptr ? atoi(getenv("SSB_MNMF_FUSED")) : 1;
  const ssb_config& c = p->cfg;
  return fused_src && c.fast_path && c.n_sources == 4 && c.n_basis <= 16 && (c.n_frames % 16) == 0 && p->fused.bytes > 0;
}

int mnmf_source(ssb_plan* p, cudaStream_t st) {
  const ssb_config& c = p->cfg;
  const int B = c.n_batch, N = c.n_sources, I = c.n_bins, J = c.n_frames, K = c.n_basis;
  // tensor-core updates (ssb_coop.cu) when the shape allows, else the CUDA-core contractions
  const bool tc = c.fast_path && c.n_basis <= 32 && (J % 16) == 0 && p->fused.bytes > 0;
  if (tc && !p->fused.zeroed) {
    SSB_CUDA(cudaMemsetAsync(p->fused.base, 0, p->fused.bytes, st));
    p->fused.zeroed = true;
  }
  // four sources, K <= 16: G / H and Lambda are formed inside the update kernels from Z2 = |Q x|^2 (one pass over X)
  // and D; no Lambda, G, H arrays at all.  SSB_MNMF_FUSED=0 (read once) keeps the array path below.
  if (tc && mnmf_fused_source(p)) {
    if (!p->in_run) p->z2_valid = false;  // outside ssb_run anything may have changed Q between two calls
    if (!p->z2_valid) {
      TRY(ssbk_mnmf_z2(p->Xk, p->Wk, p->big, B, N, I, J, st));
      p->z2_scaled = false;
    }
    const float* zs = p->z2_scaled ? p->zscale : nullptr;
    TRY(ssb_coop_mnmf_update(&c, 0, p->big, zs, p->variance, p->T, p->V, p->fused.base, st));
    return ssb_coop_mnmf_update(&c, 1, p->big, zs, p->variance, p->T, p->V, p->fused.base, st);
  }
  const float* lam;
  TRY(mnmf_lambda(p, &lam, st));
  TRY(ssbk_mnmf_gh(p->Xk, p->T, p->V, lam, p->Wk, p->variance, p->big, p->big2, B, N, I, J, K, st));
  if (tc) TRY(ssb_coop_update_ab(&c, 0, p->big, p->big2, p->T, p->V, p->fused.base, st));
  else TRY(ssbk_nmf_basis_ab(p->big, p->big2, p->T, p->V, B * N, I, J, K, c.flooring, c.eps, st));
  TRY(mnmf_lambda(p, &lam, st));
  TRY(ssbk_mnmf_gh(p->Xk, p->T, p->V, lam, p->Wk, p->variance, p->big, p->big2, B, N, I, J, K, st));
  if (tc) return ssb_coop_update_ab(&c, 1, p->big, p->big2, p->T, p->V, p->fused.base, st);
  return ssbk_nmf_activation_ab(p->big, p->big2, p->T, p->V, B * N, I, J, K, c.flooring, c.eps, st);
}

int mnmf_spatial(ssb_plan* p, cudaStream_t st) {
  const ssb_config& c = p->cfg;
  const int B = c.n_batch, N = c.n_sources, I = c.n_bins, J = c.n_frames, K = c.n_basis;
  const float* lam;
  TRY(mnmf_lambda(p, &lam, st));
  if (lam != nullptr && c.fast_path && (J % 16) == 0) {
    // the weights 1 / L_m are formed inside the covariance kernel from Lambda and D (no phi array, no km_phi pass)
    TRY(ssb_fused_cov_lambda(p->Xk, lam, p->variance, p->U, B, N, I, J, st));
  } else {
    TRY(ssbk_mnmf_phi(p->Xk, p->T, p->V, lam, p->Wk, p->variance, p->big, B, N, I, J, K, st));
    if (c.fast_path && (J % 16) == 0)
      TRY(ssb_fused_cov_w(p->Xk, p->big, (long long)N * I * J, (long long)I * J, J, N, p->U, B, N, I, J, st));
    else
      TRY(ssbk_wcov(p->Xk, p->big, (long long)N * I * J, (long long)I * J, J, nullptr, N, p->U, B, N, I, J, st));
  }
  if (c.spatial == SSB_SPATIAL_IP1) TRY(ssbk_ip1(p->Wk, p->U, B * I, N, c.flooring, c.eps, st));
  else TRY(ssbk_ip2(p->Wk, p->U, B * I, N, c.pairs, c.n_pairs, N, nullptr, c.flooring, c.eps, st));
  // inside ssb_run the sweep also leaves Z2 of the new filters for the next iteration's source model (SSB_MNMF_Z2EMIT=0:
  // a separate km_z2 pass per iteration instead)
  static const int emit_on = getenv("SSB_MNMF_Z2EMIT") != nullptr ? atoi(getenv("SSB_MNMF_Z2EMIT")) : 1;
  const bool emit = emit_on && p->in_run && lam != nullptr && mnmf_fused_source(p);
  TRY(ssbk_mnmf_spatial(p->Xk, p->T, p->V, lam, p->Wk, p->variance, p->rowloss, B, N, I, J, K, 1, st,
                        emit ? p->big : nullptr));
  p->z2_valid = emit;
  p->z2_scaled = false;
  return 0;
}

int mnmf_normalize(ssb_plan* p, bool have_zsum, cudaStream_t st) {
  const ssb_config& c = p->cfg;
  const int B = c.n_batch, N = c.n_sources, I = c.n_bins, J = c.n_frames, K = c.n_basis;
  if (!have_zsum) TRY(ssbk_mnmf_spatial(p->Xk, p->T, p->V, nullptr, p->Wk, p->variance, p->rowloss, B, N, I, J, K, 0, st));
  TRY(ssbk_mnmf_normalize(p->rowloss, p->Wk, p->variance, B, N, I, J, c.flooring, c.eps, st,
                          p->z2_valid ? p->zscale : nullptr));
  p->z2_scaled = p->z2_valid;
  return 0;
}

int mnmf_loss(ssb_plan* p, double* loss, cudaStream_t st) {
  const ssb_config& c = p->cfg;
  const int B = c.n_batch, N = c.n_sources, I = c.n_bins, J = c.n_frames, K = c.n_basis;
  const float* lam;
  TRY(mnmf_lambda(p, &lam, st));
  TRY(ssbk_mnmf_rowloss(p->Xk, p->T, p->V, lam, p->Wk, p->variance, p->rowloss, B, N, I, J, K, st));
  TRY(ssbk_logdet(p->Wk, p->logdet, B * I, N, st));
  if (p->whiten) TRY(ssbk_add_logdet(p->logdet, p->ldM, B * I, st));
  return ssbk_ilrma_loss_reduce(p->rowloss, p->logdet, loss, B, 1, I, st);
}

}  // namespace

extern "C" int ssb_plan_create(const ssb_config* cfg, ssb_plan** plan) {
  SSB_REQUIRE(plan != nullptr, "plan out-pointer is NULL");
  TRY(validate(cfg));
  ssb_plan* p = new (std::nothrow) ssb_plan();
  SSB_REQUIRE(p != nullptr, "out of host memory");
  p->cfg = *cfg;
  *plan = p;
  return 0;
}

extern "C" int ssb_plan_destroy(ssb_plan* plan) {
  delete plan;
  return 0;
}

extern "C" int ssb_plan_workspace_bytes(const ssb_plan* plan, size_t* bytes) {
  SSB_REQUIRE(plan != nullptr && bytes != nullptr, "NULL argument");
  ssb_plan tmp = *plan;
  *bytes = carve(&tmp, nullptr);
  return 0;
}

extern "C" int ssb_plan_bind(ssb_plan* p, const void* X, void* W, void* Y, void* T, void* V, void* variance,
                             void* workspace, size_t workspace_bytes) {
  SSB_REQUIRE(p != nullptr, "plan is NULL");
  SSB_REQUIRE(X != nullptr && Y != nullptr, "X and Y must be bound");
  SSB_REQUIRE(p->iss() || W != nullptr, "W must be bound unless spatial_algorithm is ISS");
  SSB_REQUIRE(!(p->ilrma() || p->mnmf()) || (T != nullptr && V != nullptr), "T and V must be bound for ILRMA / MNMF");
  SSB_REQUIRE(p->cfg.model != SSB_MODEL_IVA_GAUSS || variance != nullptr, "variance must be bound for AuxGaussIVA");
  SSB_REQUIRE(!p->mnmf() || variance != nullptr, "spatial (D) must be bound in the variance slot for FastGaussMNMF");
  SSB_REQUIRE(!p->part() || variance != nullptr, "latent (Z) must be bound in the variance slot with partitioning");
  size_t need = 0;
  TRY(ssb_plan_workspace_bytes(p, &need));
  SSB_REQUIRE(workspace != nullptr && workspace_bytes >= need, "workspace too small: %zu < %zu bytes", workspace_bytes,
              need);
  p->X = (const cf*)X;
  p->W = (cf*)W;
  p->Y = (cf*)Y;
  p->T = (float*)T;
  p->V = (float*)V;
  p->variance = (float*)variance;
  p->ws = (char*)workspace;
  p->ws_bytes = workspace_bytes;
  carve(p, p->ws);
  p->Xk = p->whiten ? p->Xw : p->X;
  p->Wk = p->whiten ? p->Ww : p->W;
  p->bound = true;
  p->prepared = false;
  return 0;
}

extern "C" int ssb_plan_set_flooring(ssb_plan* p, int flooring, float eps) {
  SSB_REQUIRE(p != nullptr, "plan is NULL");
  SSB_REQUIRE(flooring >= 0 && flooring <= 2, "unknown flooring mode %d", flooring);
  p->cfg.flooring = flooring;
  p->cfg.eps = eps;
  return 0;
}

extern "C" int ssb_plan_prepare(ssb_plan* p, void* stream) {
  TRY(require_bound(p));
  const ssb_config& c = p->cfg;
  cudaStream_t st = (cudaStream_t)stream;
  if (p->whiten) {
    // C = L L^H in fp64, Z = L^-1 X: the slab every kernel of the iteration reads from now on (ssb_whiten.cu)
    TRY(ssbk_whiten_prepare(p->X, p->C64, p->Mw, p->Mwinv, p->ldM, p->Xw, p->wsync, c.n_batch, c.n_sources, c.n_bins,
                            c.n_frames, st));
  }
  if (p->mnmf()) {
    p->prepared = true;
    return 0;
  }
  if (p->ilrma() && !p->iss()) {
    // C_i = mean_j x x^H, constant over the iterations
    TRY(ssbk_wcov(p->Xk, nullptr, 0, 0, 0, nullptr, 1, p->C, c.n_batch, c.n_sources, c.n_bins, c.n_frames, st));
  }
  TRY(ssb_fused_prepare(&p->fused, &c, p->Xk, st));
  p->prepared = true;
  return 0;
}

extern "C" int ssb_update_source_model(ssb_plan* p, void* stream) {
  TRY(require_bound(p));
  p->r2_valid = false;
  TRY(w_enter(p, (cudaStream_t)stream));
  if (p->mnmf()) return mnmf_source(p, (cudaStream_t)stream);
  if (p->fdica()) return 0;  // no source parameters
  return p->ilrma() ? ilrma_source(p, (cudaStream_t)stream) : iva_source(p, (cudaStream_t)stream);
}

extern "C" int ssb_update_source_part(ssb_plan* p, int part, void* stream) {
  TRY(require_bound(p));
  SSB_REQUIRE(p->ilrma(), "update_source_part is defined for the ILRMA family only");
  p->fused.vs_valid = false;
  TRY(w_enter(p, (cudaStream_t)stream));
  return ilrma_source_substep(p, part, (cudaStream_t)stream);
}

extern "C" int ssb_update_spatial_model(ssb_plan* p, void* stream) {
  TRY(require_bound(p));
  p->r2_valid = false;
  cudaStream_t st = (cudaStream_t)stream;
  TRY(w_enter(p, st));
  if (p->mnmf()) TRY(mnmf_spatial(p, st));
  else if (p->fdica()) TRY(fdica_spatial(p, st));
  else TRY(p->ilrma() ? ilrma_spatial(p, st) : iva_spatial(p, st));
  return w_exit(p, st);
}

extern "C" int ssb_normalize(ssb_plan* p, void* stream) {
  TRY(require_bound(p));
  cudaStream_t st = (cudaStream_t)stream;
  TRY(w_enter(p, st));
  if (p->mnmf()) {
    TRY(mnmf_normalize(p, false, st));
  } else {
    SSB_REQUIRE(p->ilrma(), "normalize is defined for ILRMA only");
    TRY(ilrma_normalize(p, st));
  }
  return w_exit(p, st);
}

namespace {
int loss_impl(ssb_plan* p, double* loss, cudaStream_t st) {
  if (p->mnmf()) return mnmf_loss(p, loss, st);
  if (p->fdica()) return fdica_loss(p, loss, st);
  return p->ilrma() ? ilrma_loss(p, loss, st) : iva_loss(p, loss, st);
}

int update_once_impl(ssb_plan* p, cudaStream_t st) {
  if (p->mnmf()) {  // mnmf.py:1278-1303
    TRY(mnmf_source(p, st));
    TRY(mnmf_spatial(p, st));
    if (p->cfg.normalization != SSB_NORM_NONE) TRY(mnmf_normalize(p, true, st));
    return 0;
  }
  if (p->ilrma()) {
    if (p->cfg.fast_path && ssb_fused_supported(&p->cfg) && p->iss()) {
      const ssb_config& c = p->cfg;
      const int B = c.n_batch, N = c.n_sources, I = c.n_bins, J = c.n_frames;
      TRY(ssb_fused_source_iss(&c, p->Y, p->T, p->V, p->big, st));
      TRY(ssb_fused_phi(&c, p->T, p->V, p->big, 1, st));
      if (c.spatial == SSB_SPATIAL_ISS2)
        TRY(ssbk_iss2(p->Y, p->big, (long long)N * I * J, (long long)I * J, J, B, N, I, J, c.pairs, c.n_pairs,
                      c.flooring, c.eps, st));
      else if (c.spatial == SSB_SPATIAL_IPA)
        TRY(ssbk_ipa(p->Y, p->big, (long long)N * I * J, (long long)I * J, J, B, N, I, J, c.ipa_normalization,
                     c.ipa_newton_iter, c.flooring, c.eps, st));
      else
        TRY(ssbk_iss1(p->Y, p->big, (long long)N * I * J, (long long)I * J, J, B, N, I, J, c.flooring, c.eps, st));
      if (c.normalization != SSB_NORM_NONE) TRY(ilrma_normalize(p, st));
      return 0;
    }
    if (p->cfg.fast_path && ssb_fused_supported(&p->cfg)) {
      const ssb_config& c = p->cfg;
      TRY(ssb_fused_source_and_cov(&c, &p->fused, p->Xk, p->Wk, p->T, p->V, p->big, p->U, st));
      // power normalisation without a pass over X: the IP kernels emit q[b,i,n] = Re(w_n C_i w_n^H) and one small
      // kernel reduces it over the bins and rescales T and W
      const bool pw = c.normalization == SSB_NORM_POWER;
      SSB_REQUIRE(!pw || p->prepared, "plan not prepared (call ssb_plan_prepare after bind)");
      const cf* Cq = pw ? p->C : nullptr;
      if (c.spatial == SSB_SPATIAL_IP1) {
        if (c.n_sources == 2)
          TRY(ssb_fused_ip1_n2(p->Wk, p->U, Cq, p->rowloss, c.n_batch * c.n_bins, c.flooring, c.eps, st));
        else
          TRY(ssbk_ip1(p->Wk, p->U, c.n_batch * c.n_bins, c.n_sources, c.flooring, c.eps, st, Cq, p->rowloss));
      } else {
        TRY(ssbk_ip2(p->Wk, p->U, c.n_batch * c.n_bins, c.n_sources, c.pairs, c.n_pairs, c.n_sources, nullptr,
                     c.flooring, c.eps, st, Cq, p->rowloss));
      }
      if (pw)
        return ssb_fused_normalize(p->rowloss, p->T, p->Wk, c.n_batch, c.n_sources, c.n_bins, c.n_basis, c.domain,
                                   c.flooring, c.eps, st);
      if (c.normalization != SSB_NORM_NONE) TRY(ilrma_normalize(p, st));
      return 0;
    }
    TRY(ilrma_source(p, st));
    TRY(ilrma_spatial(p, st));
    if (p->cfg.normalization != SSB_NORM_NONE) TRY(ilrma_normalize(p, st));
    return 0;
  }
  if (p->fdica()) return fdica_spatial(p, st);
  TRY(iva_source(p, st));
  return iva_spatial(p, st);
}
}  // namespace

extern "C" int ssb_update_once(ssb_plan* p, void* stream) {
  TRY(require_bound(p));
  // the caller may have rewritten V since the last call: the pre-split copy is only trusted inside ssb_run
  p->fused.vs_valid = false;
  p->r2_valid = false;
  TRY(w_enter(p, (cudaStream_t)stream));
  TRY(update_once_impl(p, (cudaStream_t)stream));
  p->r2_valid = false;
  return w_exit(p, (cudaStream_t)stream);
}

extern "C" int ssb_compute_loss(ssb_plan* p, double* loss, void* stream) {
  TRY(require_bound(p));
  SSB_REQUIRE(loss != nullptr, "loss output is NULL");
  p->r2_valid = false;
  TRY(w_enter(p, (cudaStream_t)stream));
  return loss_impl(p, loss, (cudaStream_t)stream);
}

namespace {
// Iterations fused across the update_once boundary (GaussILRMA-IP1, two sources, inside ssb_run): the spatial update of
// iteration t and the basis update of iteration t + 1 as one TMA tile kernel (ssb_tma.cu).  On when SSB_TMA has bit 2
// (ssb_fused_iter_fusable checks it); SSB_FUSE_ITER=0 (read at every call so that tests can toggle it) switches it off.
int fuse_iter_enabled(const ssb_config*) {
  const char* e = getenv("SSB_FUSE_ITER");
  return e ? (atoi(e) & 1) : 1;
}

// n_iter x update_once (ilrma.py:900-922) regrouped as
//   [T, V]_1, { [U, IP1]_t + [T]_(t+1), [V]_(t+1), [normalise]_t }_(t = 1 .. n-1), [U, IP1, normalise]_n
int run_fused_iterations(ssb_plan* p, int n_iter, cudaStream_t st) {
  const ssb_config& c = p->cfg;
  const bool pw = c.normalization == SSB_NORM_POWER;
  SSB_REQUIRE(!pw || p->prepared, "plan not prepared (call ssb_plan_prepare after bind)");
  TRY(ssb_fused_source_and_cov(&c, &p->fused, p->Xk, p->Wk, p->T, p->V, p->big, p->U, st, 1));
  for (int it = 1; it < n_iter; ++it) {
    TRY(ssb_fused_spatial_source(&c, &p->fused, p->Xk, p->Wk, p->T, p->V, p->big, p->rowloss, st));
    if (pw)
      TRY(ssb_fused_normalize(p->rowloss, p->T, p->Wk, c.n_batch, c.n_sources, c.n_bins, c.n_basis, c.domain, c.flooring,
                              c.eps, st));
  }
  TRY(ssb_fused_source_and_cov(&c, &p->fused, p->Xk, p->Wk, p->T, p->V, p->big, p->U, st, 2));
  TRY(ssb_fused_ip1_n2(p->Wk, p->U, pw ? p->C : nullptr, p->rowloss, c.n_batch * c.n_bins, c.flooring, c.eps, st));
  if (pw)
    TRY(ssb_fused_normalize(p->rowloss, p->T, p->Wk, c.n_batch, c.n_sources, c.n_bins, c.n_basis, c.domain, c.flooring,
                            c.eps, st));
  return 0;
}
}  // namespace

extern "C" int ssb_run(ssb_plan* p, int n_iter, double* loss, void* stream) {
  TRY(require_bound(p));
  p->fused.vs_valid = false;
  p->r2_valid = false;
  if (n_iter <= 0) return 0;
  // the whole loop stays in the whitened domain: one import before, one export after
  TRY(w_enter(p, (cudaStream_t)stream));
  if (loss == nullptr && n_iter >= 2 && p->ilrma() && p->cfg.fast_path && !p->iss() &&
      ssb_fused_iter_fusable(&p->cfg, &p->fused) && fuse_iter_enabled(&p->cfg)) {
    const int rc = run_fused_iterations(p, n_iter, (cudaStream_t)stream);
    p->fused.vs_valid = false;
    if (rc) return rc;
    return w_exit(p, (cudaStream_t)stream);
  }
  {
    struct RunScope {  // state handed from one iteration to the next is only trusted between the iterations of this loop
      ssb_plan* q;
      explicit RunScope(ssb_plan* q_) : q(q_) { q->in_run = true; q->z2_valid = false; }
      ~RunScope() { q->in_run = false; q->z2_valid = false; }
    } scope(p);
    for (int it = 0; it < n_iter; ++it) {
      TRY(update_once_impl(p, (cudaStream_t)stream));
      if (loss) TRY(loss_impl(p, loss + (size_t)it * p->cfg.n_batch, (cudaStream_t)stream));
    }
  }
  p->fused.vs_valid = false;
  p->r2_valid = false;
  return w_exit(p, (cudaStream_t)stream);
}

extern "C" int ssb_plan_separate(ssb_plan* p, void* stream) {
  TRY(require_bound(p));
  TRY(w_enter(p, (cudaStream_t)stream));
  if (p->mnmf()) {  // multichannel Wiener filter, mnmf.py:1174-1217
    const ssb_config& c = p->cfg;
    // the filter acts on the mixture itself: Q^-1 = M^-1 Q~^-1 in fp64, X in the caller's domain
    return ssbk_mnmf_separate(p->X, p->T, p->V, p->Wk, p->variance, p->qinv, p->Y, c.n_batch, c.n_sources, c.n_bins,
                              c.n_frames, c.n_basis, c.reference_id, c.flooring, c.eps, (cudaStream_t)stream,
                              p->whiten ? p->Mwinv : nullptr);
  }
  if (p->iss()) return 0;
  const ssb_config& c = p->cfg;
  return ssbk_separate(p->Xk, p->Wk, p->Y, nullptr, c.n_batch, c.n_sources, c.n_bins, c.n_frames, (cudaStream_t)stream);
}

extern "C" int ssb_restore_scale(ssb_plan* p, void* stream) {
  TRY(require_bound(p));
  SSB_REQUIRE(!p->mnmf(), "FastGaussMNMF has no scale restoration (its separate() is the Wiener filter)");
  const ssb_config& c = p->cfg;
  cudaStream_t st = (cudaStream_t)stream;
  const int B = c.n_batch, N = c.n_sources, I = c.n_bins, J = c.n_frames;
  if (p->iss()) {
    TRY(ssbk_cross_solve(p->Xk, p->Y, p->S, B, N, I, J, st));
    return ssbk_scale_rows(p->Y, p->S, p->Y, B, N, I, J, c.reference_id, st);
  }
  TRY(w_enter(p, st));
  if (p->whiten) TRY(ssbk_pb_whitened(p->Wk, p->Mwinv, nullptr, B * I, N, c.reference_id, st));
  else TRY(ssbk_pb_w(p->Wk, p->Wk, nullptr, B * I, N, c.reference_id, st));
  TRY(w_exit(p, st));
  return ssbk_separate(p->Xk, p->Wk, p->Y, nullptr, B, N, I, J, st);
}

// restore_scale with the minimal distortion principle (ilrma.py:567-579, :1981-1989; iva.py:269-281,
// :2206-2214): Y <- mdp(W X | Y, X); W modes also refit W = Y X^H (X X^H)^-1
extern "C" int ssb_restore_scale_mdp(ssb_plan* p, void* stream) {
  TRY(require_bound(p));
  SSB_REQUIRE(!p->mnmf(), "FastGaussMNMF has no scale restoration");
  const ssb_config& c = p->cfg;
  cudaStream_t st = (cudaStream_t)stream;
  const int B = c.n_batch, N = c.n_sources, I = c.n_bins, J = c.n_frames;
  if (!p->iss()) {
    TRY(w_enter(p, st));
    TRY(ssbk_separate(p->Xk, p->Wk, p->Y, nullptr, B, N, I, J, st));
  }
  // the reference channel and the refit are in the caller's domain: X and W, not their whitened images (the next
  // entry point re-imports W because it no longer matches the last export)
  TRY(ssbk_mdp(p->Y, p->X, p->Y, B, N, I, J, c.reference_id, st));
  if (!p->iss()) TRY(ssbk_cross_solve(p->Y, p->X, p->W, B, N, I, J, st));
  return 0;
}

// ---- standalone operators -----------------------------------------------------------------------
extern "C" int ssb_minimal_distortion_principle(const void* Y, const void* X, void* Yout, int B, int N, int I, int J,
                                                int reference_id, void* stream) {
  SSB_REQUIRE(Y && X && Yout, "NULL argument");
  if (B <= 0 || I <= 0 || J <= 0) return 0;
  return ssbk_mdp((const cf*)Y, (const cf*)X, (cf*)Yout, B, N, I, J, reference_id, (cudaStream_t)stream);
}

extern "C" int ssb_separate(const void* X, const void* W, void* Y, int B, int N, int I, int J, void* stream) {
  SSB_REQUIRE(X && W && Y, "NULL argument");
  if (B <= 0 || I <= 0 || J <= 0) return 0;
  return ssbk_separate((const cf*)X, (const cf*)W, (cf*)Y, nullptr, B, N, I, J, (cudaStream_t)stream);
}

extern "C" int ssb_weighted_covariance(const void* X, const float* phi, long long phi_sb, long long phi_sn,
                                       long long phi_si, const int32_t* src, int n_src, void* U, int B, int N, int I,
                                       int J, void* stream) {
  SSB_REQUIRE(X && U, "NULL argument");
  if (B <= 0 || I <= 0) return 0;
  return ssbk_wcov((const cf*)X, phi, phi_sb, phi_sn, phi_si, src, n_src, (cf*)U, B, N, I, J, (cudaStream_t)stream);
}

extern "C" int ssb_update_by_ip1(void* W, const void* U, int n_mat, int N, int flooring, float eps, void* stream) {
  SSB_REQUIRE(W && U, "NULL argument");
  if (n_mat <= 0) return 0;
  return ssbk_ip1((cf*)W, (const cf*)U, n_mat, N, flooring, eps, (cudaStream_t)stream);
}

extern "C" int ssb_update_by_ip2(void* W, const void* U, int n_mat, int N, const int32_t* pairs, int n_pairs,
                                 int flooring, float eps, void* stream) {
  SSB_REQUIRE(W && U && (pairs || n_pairs == 0), "NULL argument");
  if (n_mat <= 0) return 0;
  return ssbk_ip2((cf*)W, (const cf*)U, n_mat, N, pairs, n_pairs, N, nullptr, flooring, eps, (cudaStream_t)stream);
}

extern "C" int ssb_update_by_ip2_one_pair(void* W, const void* U_pair, int n_mat, int N, int m, int n, int flooring,
                                          float eps, void* stream) {
  SSB_REQUIRE(W && U_pair, "NULL argument");
  if (n_mat <= 0) return 0;
  const int pr[2] = {m, n};
  const int uidx[2] = {0, 1};
  return ssbk_ip2((cf*)W, (const cf*)U_pair, n_mat, N, pr, 1, 2, uidx, flooring, eps, (cudaStream_t)stream);
}

extern "C" int ssb_update_by_iss1(void* Y, const float* phi, long long phi_sb, long long phi_sn, long long phi_si,
                                  int B, int N, int I, int J, int flooring, float eps, void* stream) {
  SSB_REQUIRE(Y && phi, "NULL argument");
  if (B <= 0 || I <= 0 || J <= 0) return 0;
  return ssbk_iss1((cf*)Y, phi, phi_sb, phi_sn, phi_si, B, N, I, J, flooring, eps, (cudaStream_t)stream);
}

extern "C" int ssb_update_by_iss2(void* Y, const float* phi, long long phi_sb, long long phi_sn, long long phi_si,
                                  int B, int N, int I, int J, const int32_t* pairs, int n_pairs, int flooring,
                                  float eps, void* stream) {
  SSB_REQUIRE(Y && phi && (pairs || n_pairs == 0), "NULL argument");
  if (B <= 0 || I <= 0 || J <= 0 || n_pairs == 0) return 0;
  return ssbk_iss2((cf*)Y, phi, phi_sb, phi_sn, phi_si, B, N, I, J, pairs, n_pairs, flooring, eps,
                   (cudaStream_t)stream);
}

extern "C" int ssb_update_by_ipa(void* Y, const float* phi, long long phi_sb, long long phi_sn, long long phi_si,
                                 int B, int N, int I, int J, int normalization, int max_iter, int flooring, float eps,
                                 void* stream) {
  SSB_REQUIRE(Y && phi, "NULL argument");
  if (B <= 0 || I <= 0 || J <= 0) return 0;
  return ssbk_ipa((cf*)Y, phi, phi_sb, phi_sn, phi_si, B, N, I, J, normalization, max_iter, flooring, eps,
                  (cudaStream_t)stream);
}

extern "C" int ssb_permutation_correlation(const void* Y, double* corr, int B, int N, int I, int J, int flooring,
                                           float eps, void* stream) {
  SSB_REQUIRE(Y && corr, "NULL argument");
  if (B <= 0 || I <= 0 || J <= 0) return 0;
  return ssbk_perm_corr((const cf*)Y, corr, B, N, I, J, flooring, eps, (cudaStream_t)stream);
}

extern "C" int ssb_permutation_align(void* Y, void* W, const int32_t* order, int32_t* perms, int B, int N, int I, int J,
                                     int flooring, float eps, void* stream) {
  SSB_REQUIRE(Y && order && perms, "NULL argument");
  if (B <= 0 || I <= 0 || J <= 0) return 0;
  return ssbk_perm_align((cf*)Y, (cf*)W, order, perms, B, N, I, J, flooring, eps, (cudaStream_t)stream);
}

// solve_permutation_by_correlation (fdica.py:257-281): Y = W X first, then the two phases on the bound buffers
extern "C" int ssb_plan_permutation_correlation(ssb_plan* p, double* corr, void* stream) {
  TRY(require_bound(p));
  const ssb_config& c = p->cfg;
  SSB_REQUIRE(p->Wk != nullptr && corr != nullptr, "permutation alignment needs demixing filters");
  cudaStream_t st = (cudaStream_t)stream;
  TRY(w_enter(p, st));
  TRY(ssbk_separate(p->Xk, p->Wk, p->Y, nullptr, c.n_batch, c.n_sources, c.n_bins, c.n_frames, st));
  return ssbk_perm_corr(p->Y, corr, c.n_batch, c.n_sources, c.n_bins, c.n_frames, c.flooring, c.eps, st);
}

extern "C" int ssb_plan_permutation_align(ssb_plan* p, const int32_t* order, int32_t* perms, void* stream) {
  TRY(require_bound(p));
  const ssb_config& c = p->cfg;
  SSB_REQUIRE(p->Wk != nullptr && order != nullptr && perms != nullptr, "NULL argument");
  TRY(w_enter(p, (cudaStream_t)stream));
  // a row permutation of W is the same row permutation of W~ = W M^-1
  TRY(ssbk_perm_align(p->Y, p->Wk, order, perms, c.n_batch, c.n_sources, c.n_bins, c.n_frames, c.flooring, c.eps,
                      (cudaStream_t)stream));
  return w_exit(p, (cudaStream_t)stream);
}

extern "C" int ssb_projection_back_w(const void* W, void* Wout, int n_mat, int N, int reference_id, void* stream) {
  SSB_REQUIRE(W && Wout, "NULL argument");
  if (n_mat <= 0) return 0;
  return ssbk_pb_w((const cf*)W, (cf*)Wout, nullptr, n_mat, N, reference_id, (cudaStream_t)stream);
}

extern "C" int ssb_projection_back_y(const void* Y, const void* X, void* Yout, void* scale_out, int B, int N, int I,
                                     int J, int reference_id, void* stream) {
  SSB_REQUIRE(Y && X && scale_out, "NULL argument (scale_out is required scratch of B*I*N*N complex64)");
  if (B <= 0 || I <= 0 || J <= 0) return 0;
  TRY(ssbk_cross_solve((const cf*)X, (const cf*)Y, (cf*)scale_out, B, N, I, J, (cudaStream_t)stream));
  if (Yout == nullptr) return 0;
  return ssbk_scale_rows((const cf*)Y, (const cf*)scale_out, (cf*)Yout, B, N, I, J, reference_id,
                         (cudaStream_t)stream);
}
