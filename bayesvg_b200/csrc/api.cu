// api.cu -- C-ABI entry points (include/bayesvg_b200.h) and the host-side drivers of the fit:
// data staging, log_prob, mean-field ADVI (eta adaptation + stochastic gradient ascent + ELBO
// convergence test), draws + summaries.  Algorithms follow SURVEY.md A.4, which restates what
// R/findSpatiallyVariableFeaturesBayes.R:346-364 asks CmdStan to do.
#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdint>
#include <map>
#include <vector>

#include "bvg_vec.cuh"

static thread_local char g_err[1024] = "";
void bvg_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" const char *bvg_last_error(void) { return g_err; }
extern "C" int bvg_version(void) { return BVG_ABI_VERSION; }
extern "C" int bvg_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}
extern "C" void bvg_config_default(bvg_config *c) {
  memset(c, 0, sizeof(*c));
  c->abi_version = BVG_ABI_VERSION;
  c->likelihood = BVG_GAUSSIAN;
  c->precision = BVG_PREC_FAST;
  c->engine = BVG_ENGINE_AUTO;
  c->device = 0;
  c->n_iter = 3000;
  c->mle_init = 1;
  c->mle_iter = 1000;
  c->n_draws = 1000;
  c->elbo_samples = 150;
  c->eval_elbo = 100;
  c->adapt_iter = 50;
  c->lbfgs_history = 25;
  c->verbose = 0;
  c->tol_rel_obj = 0.01;
  c->eta = 0.0;
  c->seed = 312;
  c->rank = 0;
  c->world_size = 1;
  c->gene_offset = 0;
  c->G_total = 0;
  c->elbo_suffstat = 1;   /* ELBO estimates from sufficient statistics where they exist (gaussian, hierarchical intercept, k <= 30) */
  if (const char *e = getenv("BVG_ELBO_SUFFSTAT")) c->elbo_suffstat = e[0] == '1' ? 1 : 0;   /* the test suite keeps the dense estimates */
  c->depth_adjust = 0;
  c->algorithm = BVG_ALG_MEANFIELD;
  c->pf_num_paths = 2;        /* n.cores = 2 (:94) */
  c->pf_max_lbfgs_iters = 200;
  c->pf_single_draws = 1000;
  c->pf_init_radius = 0.01;
}

int basis_build_device(cudaStream_t st, const double *d_X, int64_t M, const double *d_C, int k, double ls, int kernel,
                       double nu, double period, double *d_Q, double *d_work);
void comm_destroy(bvg_fit *f);

static int g_no_elbo_batch = 0;   // diagnostics: BVG_NO_ELBO_BATCH=1 keeps one k_prep launch per ELBO draw
static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
static int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }
static int comm_failed(const Ctl &h);

// ---- lifecycle ----------------------------------------------------------------------------------
extern "C" int bvg_fit_create(const bvg_config *cfg, bvg_fit **out) {
  if (!cfg || !out) { bvg_set_error("null argument"); return BVG_ERR_ARG; }
  if (cfg->abi_version != BVG_ABI_VERSION) { bvg_set_error("ABI version mismatch: caller %d, library %d", cfg->abi_version, BVG_ABI_VERSION); return BVG_ERR_ARG; }
  if (cfg->likelihood != BVG_GAUSSIAN && cfg->likelihood != BVG_NB) { bvg_set_error("Please provide a valid likelihood."); return BVG_ERR_ARG; }
  if (cfg->precision != BVG_PREC_FP64 && cfg->precision != BVG_PREC_FAST) { bvg_set_error("invalid precision"); return BVG_ERR_ARG; }
  if (cfg->engine == BVG_ENGINE_TCGEN05 && cfg->precision != BVG_PREC_FAST) { bvg_set_error("the tcgen05 engine has no fp64 kind; use precision = FAST"); return BVG_ERR_ARG; }
  if (cfg->n_iter < 1 || cfg->elbo_samples < 1 || cfg->eval_elbo < 1 || cfg->adapt_iter < 1 || cfg->n_draws < 2) { bvg_set_error("iteration / sample counts must be positive"); return BVG_ERR_ARG; }
  if (cfg->world_size < 1 || cfg->rank < 0 || cfg->rank >= cfg->world_size) { bvg_set_error("bad rank/world_size"); return BVG_ERR_ARG; }
  if (cfg->algorithm < BVG_ALG_MEANFIELD || cfg->algorithm > BVG_ALG_PATHFINDER) { bvg_set_error("Please provide a valid variational inference approximation algorithm."); return BVG_ERR_ARG; }
  if (cfg->algorithm == BVG_ALG_FULLRANK && cfg->world_size > 1) { bvg_set_error("algorithm = 'fullrank' couples all parameters and cannot be gene-sharded"); return BVG_ERR_ARG; }
  if (cfg->algorithm == BVG_ALG_PATHFINDER && cfg->world_size > 1) { bvg_set_error("algorithm = 'pathfinder' is not built for gene-sharded fits"); return BVG_ERR_ARG; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    bvg_set_error("no CUDA device: bayesvg_b200 has no CPU fallback");
    return BVG_ERR_CUDA;
  }
  if (cfg->device < 0 || cfg->device >= ndev) { bvg_set_error("device %d out of range (have %d)", cfg->device, ndev); return BVG_ERR_ARG; }
  CUDA_TRY(cudaSetDevice(cfg->device));
  bvg_fit *f = new bvg_fit();
  memset(f, 0, sizeof(*f));
  { const char *e = getenv("BVG_NO_ELBO_BATCH"); g_no_elbo_batch = e && e[0] == '1'; }
  { const char *e = getenv("BVG_NO_PERSIST"); f->no_persist = e && e[0] == '1'; }
  { const char *e = getenv("BVG_NO_MLE_FP64"); f->no_mle_fp64 = e && e[0] == '1'; }
  f->cfg = *cfg;
  CUDA_TRY(cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking));
  CUDA_TRY(cudaEventCreate(&f->ev0));
  CUDA_TRY(cudaEventCreate(&f->ev1));
  *out = f;
  return BVG_OK;
}

// The three large per-fit blocks (the staged Y, the H2D staging buffers, the L-BFGS work space) come from a small process-wide
// cache of exact-size device blocks: cudaMalloc / cudaFree of 30 - 120 MB cost 1 ms typically and 20 - 400 ms when the driver
// maps or trims memory, which on a 0.2 s fit is the difference between runs (bench.py lists every run).  A released block
// is kept (up to 16 blocks / 1.5 GB per process, blocks >= 4 MB) and handed to the next fit that asks for the same size --
// the next sample of an R session, the next fit of a test suite.  The owner synchronises its streams before releasing.
#include <mutex>
namespace {
struct BigBlock { int device; void *p; size_t bytes; };
std::mutex g_big_mu;
std::vector<BigBlock> g_big;
size_t g_big_total = 0;
}  // namespace
int bvg_big_alloc(int device, void **p, size_t bytes) {
  {
    std::lock_guard<std::mutex> lk(g_big_mu);
    for (size_t i = 0; i < g_big.size(); ++i)
      if (g_big[i].device == device && g_big[i].bytes == bytes) {
        *p = g_big[i].p;
        g_big_total -= bytes;
        g_big.erase(g_big.begin() + i);
        return BVG_OK;
      }
  }
  if (cudaMalloc(p, bytes) != cudaSuccess) {
    cudaGetLastError();
    // out of memory: give the cached blocks back to the driver and try once more
    std::lock_guard<std::mutex> lk(g_big_mu);
    for (auto &b : g_big) cudaFree(b.p);
    g_big.clear(); g_big_total = 0;
    CUDA_TRY(cudaMalloc(p, bytes));
  }
  return BVG_OK;
}
void bvg_big_free(int device, void *p, size_t bytes) {
  if (!p) return;
  {
    std::lock_guard<std::mutex> lk(g_big_mu);
    if (bytes >= ((size_t)4 << 20) && g_big.size() < 16 && g_big_total + bytes <= ((size_t)3 << 29)) {
      g_big.push_back({device, p, bytes});
      g_big_total += bytes;
      return;
    }
  }
  cudaFree(p);
}

// ---- pooled device allocations (bvg_internal.cuh) -----------------------------------------------------------------
namespace {
struct PoolDev { int state = 0; cudaStream_t st = nullptr; cudaMemPool_t pool = nullptr; };   // state: 0 unknown, 1 pooled, 2 plain
PoolDev g_pool[64];
std::mutex g_pool_mu;
PoolDev *pool_of_current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { cudaGetLastError(); return nullptr; }
  PoolDev &pd = g_pool[dev];
  if (pd.state == 0) {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    if (pd.state == 0) {
      int ok = 0;
      uint64_t keep = (uint64_t)8 << 30;
      if (getenv("BVG_NO_POOL") == nullptr && cudaDeviceGetAttribute(&ok, cudaDevAttrMemoryPoolsSupported, dev) == cudaSuccess && ok &&
          cudaDeviceGetDefaultMemPool(&pd.pool, dev) == cudaSuccess &&
          cudaMemPoolSetAttribute(pd.pool, cudaMemPoolAttrReleaseThreshold, &keep) == cudaSuccess &&
          cudaStreamCreateWithFlags(&pd.st, cudaStreamNonBlocking) == cudaSuccess) pd.state = 1;
      else { cudaGetLastError(); pd.state = 2; }
    }
  }
  return pd.state == 1 ? &pd : nullptr;
}
}  // namespace
cudaError_t bvg_pool_malloc(void **p, size_t bytes) {
  PoolDev *pd = pool_of_current_device();
  if (!pd) return (cudaMalloc)(p, bytes);
  const cudaError_t e = cudaMallocAsync(p, bytes, pd->st);
  if (e != cudaSuccess) return e;
  return cudaStreamSynchronize(pd->st);   // the block is complete: usable on every stream
}
cudaError_t bvg_pool_free(void *p) {
  if (!p) return cudaSuccess;
  PoolDev *pd = pool_of_current_device();
  if (!pd) return (cudaFree)(p);
  cudaDeviceSynchronize();                // cudaFree's implicit wait: nothing may still use the block
  if (cudaFreeAsync(p, pd->st) == cudaSuccess) return cudaSuccess;
  cudaGetLastError();
  return (cudaFree)(p);                   // a block that did not come from the pool
}

extern "C" void bvg_release_cached_memory(void) {
  {
    std::lock_guard<std::mutex> lk(g_big_mu);
    for (auto &b : g_big) cudaFree(b.p);
    g_big.clear();
    g_big_total = 0;
  }
  for (auto &pd : g_pool)
    if (pd.state == 1) { cudaStreamSynchronize(pd.st); cudaMemPoolTrimTo(pd.pool, 0); }
}

static void free_model(bvg_fit *f) {
  // nothing of this model may be in flight when its blocks go back to the cache (cudaFree used to wait implicitly)
  if (f->stream) cudaStreamSynchronize(f->stream);
  if (f->copy_stream) cudaStreamSynchronize(f->copy_stream);
  shadow_destroy(f);
  lbfgs_cache_destroy(f);
  fr_destroy(f);
  cudaFree(f->eb_d);
  f->eb_d = nullptr; f->eb_cap = 0;
  tc_destroy(f);
  tcbig_destroy(f);
  suff_destroy(f);
  bvg_big_free(f->cfg.device, f->Y, f->Y_bytes); f->Y_bytes = 0;
  cudaFree(f->Phi); cudaFree(f->dvec); cudaFree(f->summary); cudaFree(f->draws);
  cudaFree(f->m.C); cudaFree(f->m.Ppart); cudaFree(f->m.Spart); cudaFree(f->m.blk); cudaFree(f->m.ctl);
  cudaFree((void *)f->m.hist_val);
  cudaFree((void *)f->m.depths);
  f->m.depths = nullptr;
  cudaFree(f->m.hterm);
  f->m.hterm = nullptr;
  cudaFree(f->m.elbo_lps);
  f->m.elbo_lps = nullptr; f->elbo_cap = 0;
  f->Y = f->Phi = nullptr; f->dvec = f->summary = f->draws = nullptr;
  f->m.C = f->m.Ppart = f->m.Spart = nullptr; f->m.blk = nullptr; f->m.ctl = nullptr; f->m.hist_val = nullptr;
}

extern "C" void bvg_fit_destroy(bvg_fit *f) {
  if (!f) return;
  cudaSetDevice(f->cfg.device);
  cudaStreamSynchronize(f->stream);
  comm_destroy(f);
  free_model(f);
  cudaFree(f->phi64);
  cudaFree(f->flush);
  bvg_big_free(f->cfg.device, f->stage_buf, f->stage_bytes); f->stage_buf = nullptr; f->stage_bytes = 0;
  if (f->copy_stream) {
    cudaStreamDestroy(f->copy_stream);
    for (int b = 0; b < 2; ++b) { cudaEventDestroy(f->ev_copied[b]); cudaEventDestroy(f->ev_staged[b]); }
  }
  delete[] f->h_theta_mle;
  delete[] f->pf_idx;
  delete[] f->h_depths;
  cudaEventDestroy(f->ev0); cudaEventDestroy(f->ev1);
  cudaStreamDestroy(f->stream);
  delete f;
}

static int resolve_engine(bvg_fit *f) {
  const bool fp64 = f->cfg.precision == BVG_PREC_FP64;
  const int KP = (int)round_up(f->m.k + 1, 8);
  f->engine = f->cfg.engine;
  // tcgen05: the fused kernel for k + 1 <= 24, the two-GEMM large-basis kernels (k_lik_tcbig.cu) above that
  if (f->engine == BVG_ENGINE_AUTO) f->engine = (!fp64 && tc_available()) ? BVG_ENGINE_TCGEN05 : BVG_ENGINE_SIMT;
  f->big_tc = f->engine == BVG_ENGINE_TCGEN05 && KP > 24;
  f->st.engine_used = f->engine;
  return BVG_OK;
}

// allocate everything that depends on (M, G, k) once both phi and y are known
static int finalize_model(bvg_fit *f) {
  DevModel &m = f->m;
  const bool fp64 = f->cfg.precision == BVG_PREC_FP64;
  const size_t ts = fp64 ? 8 : 4;
  m.lik = f->cfg.likelihood;
  m.KP = (int)round_up(m.k + 1, 8);
  m.Gpad = round_up(m.G, BVG_GENE_TILE);
  m.G_total = f->cfg.G_total > 0 ? f->cfg.G_total : m.G;
  m.gene_offset = f->cfg.gene_offset;
  if (m.gene_offset < 0 || m.gene_offset + m.G > m.G_total) { bvg_set_error("gene shard [%lld, %lld) outside G_total = %lld", (long long)m.gene_offset, (long long)(m.gene_offset + m.G), (long long)m.G_total); return BVG_ERR_ARG; }
  m.depth = f->cfg.depth_adjust ? 1 : 0;
  m.L = make_layout(m.lik, m.G, m.k, m.depth);
  m.LG = make_layout(m.lik, m.G_total, m.k, m.depth);
  m.seed = f->cfg.seed;
  const int64_t D = m.L.D;
  // state vectors: mu | omega | hist(2D) | zeta | grad | init | weights
  CUDA_TRY(cudaMalloc(&f->dvec, sizeof(double) * 8 * D));
  CUDA_TRY(cudaMemsetAsync(f->dvec, 0, sizeof(double) * 8 * D, f->stream));
  m.mu = f->dvec; m.omega = m.mu + D; m.hist = m.omega + D; m.zeta = m.hist + 2 * D; m.grad = m.zeta + D;
  f->scratch = m.grad + D;  // [D] theta_init, then [D] weights
  BVG_TRY(vec_make_weights(f->stream, f->scratch + D, m, f->cfg.rank));
  // split the spots so that (gene tiles x splits) covers the machine a few times over
  int nsm = 148;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, f->cfg.device);
  const int ntiles = (int)(m.Gpad / BVG_GENE_TILE);
  const int64_t chunk = f->big_tc ? 128 : (f->engine == BVG_ENGINE_TCGEN05 ? 64 : 32);  // spots per pipeline chunk of the likelihood kernel
  const int64_t nchunks = (m.M + chunk - 1) / chunk;
  // SIMT: ~4 resident CTAs per SM.  tcgen05: one ~220 KB CTA per SM, so the grid runs in whole waves of nsm CTAs;
  // pick the split count that minimises waves x (chunks per CTA + ~4 chunk-times of prologue / drain per wave).
  // (With "as many CTAs as fit in ONE wave" a 79-tile problem -- config 4 on one GPU -- ran on 79 of 148 SMs.)
  int nsplit;
  if (f->big_tc) {
    nsplit = std::max(1, nsm / ntiles);  // forward kernel: one wave of (gene tiles x splits) CTAs
  } else if (f->engine == BVG_ENGINE_TCGEN05) {
    int64_t best_cost = INT64_MAX;
    nsplit = 1;
    const int64_t smax = std::max<int64_t>(1, std::min<int64_t>(nchunks / 6, 96));
    for (int64_t s = 1; s <= smax; ++s) {
      const int64_t c = (nchunks + s - 1) / s, se = (nchunks + c - 1) / c;
      const int64_t waves = (ntiles * se + nsm - 1) / nsm, cost = waves * (c + 4);
      if (cost < best_cost) { best_cost = cost; nsplit = (int)se; }
    }
  } else {
    nsplit = (nsm * 4 + ntiles - 1) / ntiles;
  }
  nsplit = (int)std::max<int64_t>(1, std::min<int64_t>(nsplit, nchunks));
  const int64_t cps = (nchunks + nsplit - 1) / nsplit;
  m.nsplit = (int)((nchunks + cps - 1) / cps);
  // coefficient rows: [Gpad][KP] in the compute type, or (tcgen05) two [Gpad][32] fp32 arrays (tf32 hi / lo)
  const size_t c_bytes = std::max<size_t>(ts * m.Gpad * m.KP, (size_t)2 * 4 * m.Gpad * 32);
  CUDA_TRY(cudaMalloc(&m.C, c_bytes));
  CUDA_TRY(cudaMemsetAsync(m.C, 0, c_bytes, f->stream));
  CUDA_TRY(cudaMalloc(&m.Ppart, ts * (size_t)m.nsplit * m.Gpad * m.KP));
  CUDA_TRY(cudaMalloc(&m.Spart, sizeof(double) * (size_t)m.nsplit * m.Gpad * 2));  // always fp64
  CUDA_TRY(cudaMemsetAsync(m.Ppart, 0, ts * (size_t)m.nsplit * m.Gpad * m.KP, f->stream));
  CUDA_TRY(cudaMemsetAsync(m.Spart, 0, sizeof(double) * (size_t)m.nsplit * m.Gpad * 2, f->stream));
  // one gene per warp, at most one CTA per SM: 16 warps per CTA, or 21 when that lets every gene be served in a
  // single pass (a warp that owns two genes runs two dependent latency chains back to back)
  f->gene_wpb = (m.G > (int64_t)16 * nsm && m.G <= (int64_t)(BVG_GENE_MAX_THREADS / 32) * nsm) ? BVG_GENE_MAX_THREADS / 32 : 16;
  if (m.k > BVG_SMALL_K) f->gene_wpb = 16;  // k_gene_big
  if (m.k > BVG_SMALL_K) BVG_TRY(gene_big_setup());
  m.nblk_gene = (int)std::min<int64_t>((m.G + f->gene_wpb - 1) / f->gene_wpb, nsm);
  f->h_t_grad = 0;
  const int NS = bvg_ns(m.k);
  CUDA_TRY(cudaMalloc(&m.blk, sizeof(double) * (((size_t)m.nblk_gene + 1) * NS + 4 + m.k + m.G + 48)));
  m.tot = m.blk + (size_t)m.nblk_gene * NS;
  m.shx = m.tot + NS;
  m.Aexp = m.shx + 4 + m.k;
  m.U = (float *)(m.blk + ((((size_t)m.nblk_gene + 1) * NS + 4 + m.k + m.G + 1) & ~(size_t)1));  // 96 floats, 16-byte aligned (float4 loads)
  m.W = (float *)m.C;             // tcgen05 engine: the coefficient-row allocation holds the fp32 draw rows instead
  CUDA_TRY(cudaMalloc(&m.ctl, sizeof(Ctl)));
  CUDA_TRY(cudaMemsetAsync(m.ctl, 0, sizeof(Ctl), f->stream));
  if (m.peer.world > 1) m.peer.fail = &m.ctl->comm_fail;  // a model re-staged after bvg_fit_peer_init
  CUDA_TRY(cudaMalloc(&f->summary, sizeof(double) * 8 * m.G));
  CUDA_TRY(cudaMalloc(&f->draws, sizeof(double) * m.G * f->cfg.n_draws));
  // Phi' in the compute type
  f->Mpad = round_up(m.M, f->big_tc ? 128 : 64);
  if (!f->big_tc) {
    CUDA_TRY(cudaMalloc(&f->Phi, ts * f->Mpad * m.KP));
    if (fp64) BVG_TRY(stage_phi<double>(f->phi64, (double *)f->Phi, m.M, f->Mpad, m.k, m.KP, f->stream));
    else BVG_TRY(stage_phi<float>(f->phi64, (float *)f->Phi, m.M, f->Mpad, m.k, m.KP, f->stream));
  }
  if (f->big_tc) BVG_TRY(tcbig_setup(f));
  else if (f->engine == BVG_ENGINE_TCGEN05) BVG_TRY(tc_setup(f));
  if (m.depth) {
    double *d = nullptr;
    CUDA_TRY(cudaMalloc(&d, sizeof(double) * m.G));
    CUDA_TRY(cudaMemsetAsync(d, 0, sizeof(double) * m.G, f->stream));
    m.depths = d;
    if (f->have_depths) {
      if (f->n_depths != m.G) { bvg_set_error("gene_depths has %lld entries but y has %lld genes", (long long)f->n_depths, (long long)m.G); return BVG_ERR_ARG; }
      CUDA_TRY(cudaMemcpyAsync(d, f->h_depths, sizeof(double) * m.G, cudaMemcpyHostToDevice, f->stream));
    }
  }
  CUDA_TRY(cudaStreamSynchronize(f->stream));
  return BVG_OK;
}

extern "C" int bvg_fit_set_gene_depths(bvg_fit *f, const double *gene_depths, int64_t G) {
  if (!f || !gene_depths || G < 1) { bvg_set_error("null or empty gene_depths"); return BVG_ERR_ARG; }
  if (!f->cfg.depth_adjust) { bvg_set_error("gene_depths belong to the depth-adjusted model: create the fit with depth_adjust = 1"); return BVG_ERR_STATE; }
  for (int64_t g = 0; g < G; ++g)
    if (!std::isfinite(gene_depths[g])) { bvg_set_error("gene_depths[%lld] is not finite (a gene with zero total count?)", (long long)g); return BVG_ERR_ARG; }
  if (f->have_y && G != f->m.G) { bvg_set_error("gene_depths has %lld entries but y has %lld genes", (long long)G, (long long)f->m.G); return BVG_ERR_ARG; }
  CUDA_TRY(cudaSetDevice(f->cfg.device));
  delete[] f->h_depths;
  f->h_depths = new double[G];
  memcpy(f->h_depths, gene_depths, sizeof(double) * G);
  f->n_depths = G;
  f->have_depths = true;
  if (f->have_y) {
    CUDA_TRY(cudaMemcpyAsync((void *)f->m.depths, f->h_depths, sizeof(double) * G, cudaMemcpyHostToDevice, f->stream));
    CUDA_TRY(cudaStreamSynchronize(f->stream));
  }
  return BVG_OK;
}

extern "C" int bvg_fit_set_phi(bvg_fit *f, const double *phi, int64_t M, int k) {
  if (!f || !phi) { bvg_set_error("null argument"); return BVG_ERR_ARG; }
  if (M < 1 || k < 1) { bvg_set_error("phi must be M x k with M, k >= 1"); return BVG_ERR_ARG; }
  if (k > BVG_MAX_K) { bvg_set_error("n.basis.fns = %d exceeds this build's limit of %d", k, BVG_MAX_K); return BVG_ERR_ARG; }
  CUDA_TRY(cudaSetDevice(f->cfg.device));
  const double t0 = now_s();
  if (f->have_y) { free_model(f); f->have_y = false; }  // a new basis invalidates the staged model; set y again
  if (f->have_phi) { cudaFree(f->phi64); f->phi64 = nullptr; }
  CUDA_TRY(cudaMalloc(&f->phi64, sizeof(double) * M * k));
  CUDA_TRY(cudaMemcpyAsync(f->phi64, phi, sizeof(double) * M * k, cudaMemcpyHostToDevice, f->stream));
  f->m.M = M; f->m.k = k;
  f->have_phi = true;
  CUDA_TRY(cudaStreamSynchronize(f->stream));
  f->st.sec_h2d += now_s() - t0;
  return BVG_OK;
}

template <typename SRC>
static int set_y_impl(bvg_fit *f, const SRC *y, int64_t M, int64_t G) {
  if (!f || !y) { bvg_set_error("null argument"); return BVG_ERR_ARG; }
  if (!f->have_phi) { bvg_set_error("call bvg_fit_set_phi before bvg_fit_set_y"); return BVG_ERR_STATE; }
  if (M != f->m.M) { bvg_set_error("y has %lld spots but phi has %lld rows", (long long)M, (long long)f->m.M); return BVG_ERR_ARG; }
  if (G < 1) { bvg_set_error("y must have at least one gene"); return BVG_ERR_ARG; }
  CUDA_TRY(cudaSetDevice(f->cfg.device));
  const double t0 = now_s();
  // same shape as the resident model (a refit on new expression values): keep every allocation, the staged
  // basis and the tensor maps; only Y is restaged
  const bool reuse = f->have_y && f->m.G == G && f->Y != nullptr;
  shadow_destroy(f);   // the fp64 copy of Y belongs to the previous data
  if (f->have_y && !reuse) free_model(f);
  f->have_y = f->have_mle = f->have_vi = f->have_summary = false;
  const bool fp64 = f->cfg.precision == BVG_PREC_FP64;
  f->m.G = G;
  if (!reuse) BVG_TRY(resolve_engine(f));
  const bool tiled = f->engine == BVG_ENGINE_TCGEN05 && !f->big_tc;
  f->tiledY = tiled;
  f->ldY = fp64 ? M : round_up(M, f->big_tc ? 128 : 64);
  const size_t ts = fp64 ? 8 : 4;
  // tcgen05 engine: Y is stored TILE-MAJOR, [gene tile][32-spot chunk][128 genes][32 spots], so that every
  // TMA box (and every 64-spot pipeline stage) is one contiguous 16 KB (32 KB) run of HBM instead of 128
  // separate 128-byte pieces 4*ldY bytes apart
  const size_t y_elems = tiled ? (size_t)round_up(G, BVG_GENE_TILE) * f->ldY : (size_t)G * f->ldY;
  if (!reuse) {
    BVG_TRY(bvg_big_alloc(f->cfg.device, &f->Y, ts * y_elems));
    f->Y_bytes = ts * y_elems;
    if (tiled) CUDA_TRY(cudaMemsetAsync(f->Y, 0, ts * y_elems, f->stream));  // pads stay zero: staging writes real entries only
  }
  // stage through two bounded device buffers (host layout -> compute type + padding): the H2D copy of chunk
  // i+1 runs on a second stream while chunk i is converted
  const int64_t genes_per_chunk = std::max<int64_t>(1, std::min<int64_t>(G, (int64_t)(32u << 20) / (M * (int64_t)sizeof(SRC))));
  const size_t stage_bytes = 2 * sizeof(SRC) * genes_per_chunk * M;
  if (f->stage_bytes < stage_bytes) {
    bvg_big_free(f->cfg.device, f->stage_buf, f->stage_bytes);
    f->stage_buf = nullptr; f->stage_bytes = 0;
    BVG_TRY(bvg_big_alloc(f->cfg.device, &f->stage_buf, stage_bytes));
    f->stage_bytes = stage_bytes;
  }
  if (!f->copy_stream) {
    CUDA_TRY(cudaStreamCreateWithFlags(&f->copy_stream, cudaStreamNonBlocking));
    for (int b = 0; b < 2; ++b) { CUDA_TRY(cudaEventCreateWithFlags(&f->ev_copied[b], cudaEventDisableTiming)); CUDA_TRY(cudaEventCreateWithFlags(&f->ev_staged[b], cudaEventDisableTiming)); }
  }
  CUDA_TRY(cudaStreamSynchronize(f->stream));  // earlier work on the compute stream no longer reads Y
  const bool timing = getenv("BVG_TIMING") != nullptr;
  const double t_alloc = now_s();
  int64_t ci = 0;
  for (int64_t g0 = 0; g0 < G; g0 += genes_per_chunk, ++ci) {
    const int b = (int)(ci & 1);
    const int64_t ng = std::min(genes_per_chunk, G - g0);
    SRC *stage = (SRC *)f->stage_buf + (size_t)b * genes_per_chunk * M;
    if (ci >= 2) CUDA_TRY(cudaStreamWaitEvent(f->copy_stream, f->ev_staged[b], 0));  // buffer b has been consumed
    CUDA_TRY(cudaMemcpyAsync(stage, y + g0 * M, sizeof(SRC) * ng * M, cudaMemcpyHostToDevice, f->copy_stream));
    CUDA_TRY(cudaEventRecord(f->ev_copied[b], f->copy_stream));
    CUDA_TRY(cudaStreamWaitEvent(f->stream, f->ev_copied[b], 0));
    int rc;
    if (tiled) rc = stage_y_tiled<SRC>(stage, (float *)f->Y, M, f->ldY, g0, ng, f->stream);
    else if (fp64) rc = stage_y<SRC, double>(stage, (double *)f->Y + g0 * f->ldY, M, f->ldY, ng, f->stream);
    else rc = stage_y<SRC, float>(stage, (float *)f->Y + g0 * f->ldY, M, f->ldY, ng, f->stream);
    if (rc != BVG_OK) return rc;
    CUDA_TRY(cudaEventRecord(f->ev_staged[b], f->stream));
  }
  // nb: histogram of distinct counts for the lgamma / digamma terms (they depend on y and phi_nb only);
  // the host pass overlaps the copies above
  if (f->cfg.likelihood == BVG_NB) {
    std::map<int64_t, double> h;
    double lg1 = 0;
    const int64_t n = M * G;
    std::vector<double> dense(1 << 16, 0.0);
    for (int64_t i = 0; i < n; ++i) {
      const double v = (double)y[i];
      if (!(v >= 0) || v != std::floor(v)) { cudaStreamSynchronize(f->copy_stream); cudaStreamSynchronize(f->stream); bvg_set_error("nb likelihood needs non-negative integer counts (found %g)", v); return BVG_ERR_ARG; }
      const int64_t c = (int64_t)v;
      if (c < (int64_t)dense.size()) dense[c] += 1.0; else h[c] += 1.0;
    }
    std::vector<double> hv, hc;
    for (size_t c = 0; c < dense.size(); ++c) if (dense[c] > 0) { hv.push_back((double)c); hc.push_back(dense[c]); }
    for (auto &kv : h) { hv.push_back((double)kv.first); hc.push_back(kv.second); }
    for (size_t i = 0; i < hv.size(); ++i) lg1 += hc[i] * lgamma(hv[i] + 1.0);
    CUDA_TRY(cudaStreamSynchronize(f->stream));
    if (f->m.hist_val) { cudaFree((void *)f->m.hist_val); f->m.hist_val = nullptr; }
    double *d = nullptr;
    CUDA_TRY(cudaMalloc(&d, sizeof(double) * 2 * hv.size()));
    CUDA_TRY(cudaMemcpyAsync(d, hv.data(), sizeof(double) * hv.size(), cudaMemcpyHostToDevice, f->stream));
    CUDA_TRY(cudaMemcpyAsync(d + hv.size(), hc.data(), sizeof(double) * hv.size(), cudaMemcpyHostToDevice, f->stream));
    CUDA_TRY(cudaStreamSynchronize(f->stream));
    f->m.hist_val = d; f->m.hist_cnt = d + hv.size(); f->m.nhist = (int)hv.size(); f->m.lgam_y1 = lg1;
    cudaFree(f->m.hterm);
    f->m.hterm = nullptr;
  }
  CUDA_TRY(cudaStreamSynchronize(f->stream));
  const double t_copy = now_s();
  if (!reuse) BVG_TRY(finalize_model(f));
  else suff_destroy(f);  // sufficient statistics belong to the previous Y
  if (timing) fprintf(stderr, "set_y: allocations %.1f ms, copy + staging of %.1f MB %.1f ms, finalize_model %.1f ms\n", 1e3 * (t_alloc - t0),
                      (double)(sizeof(SRC) * M * G) / 1e6, 1e3 * (t_copy - t_alloc), 1e3 * (now_s() - t_copy));
  // nb, tcgen05 engine: the per-count lgamma / digamma terms are evaluated by 31 idle lanes per likelihood CTA --
  // only when there is a lane for every distinct count; otherwise k_gene's finish evaluates them itself
  if (f->cfg.likelihood == BVG_NB && f->engine == BVG_ENGINE_TCGEN05 && !f->big_tc &&
      (int64_t)(f->m.Gpad / BVG_GENE_TILE) * f->m.nsplit * 31 >= f->m.nhist)
    CUDA_TRY(cudaMalloc(&f->m.hterm, sizeof(double) * 2 * f->m.nhist));
  f->have_y = true;
  f->st.sec_h2d += now_s() - t0;
  return BVG_OK;
}
extern "C" int bvg_fit_set_y_f64(bvg_fit *f, const double *y, int64_t M, int64_t G) { return set_y_impl<double>(f, y, M, G); }
extern "C" int bvg_fit_set_y_i32(bvg_fit *f, const int32_t *y, int64_t M, int64_t G) {
  if (f && f->cfg.likelihood != BVG_NB) { bvg_set_error("integer counts are the nb likelihood's input; gaussian takes doubles"); return BVG_ERR_ARG; }
  return set_y_impl<int32_t>(f, y, M, G);
}

static int need_ready(bvg_fit *f, bool need_depths = true) {
  if (!f) { bvg_set_error("null fit"); return BVG_ERR_ARG; }
  if (!f->have_phi || !f->have_y) { bvg_set_error("set phi and y first"); return BVG_ERR_STATE; }
  if (need_depths && f->cfg.depth_adjust && !f->have_depths) { bvg_set_error("gene.depth.adjust: call bvg_fit_set_gene_depths first"); return BVG_ERR_STATE; }
  CUDA_TRY(cudaSetDevice(f->cfg.device));
  return BVG_OK;
}
extern "C" int bvg_fit_dim(bvg_fit *f, int64_t *D) {
  BVG_TRY(need_ready(f, false));
  *D = f->m.L.D;
  return BVG_OK;
}

extern "C" int bvg_fit_shape(bvg_fit *f, int64_t *M, int64_t *G, int *k, int64_t *D) {
  BVG_TRY(need_ready(f, false));
  if (M) *M = f->m.M;
  if (G) *G = f->m.G;
  if (k) *k = f->m.k;
  if (D) *D = f->m.L.D;
  return BVG_OK;
}

// ---- one evaluation -----------------------------------------------------------------------------
int launch_lik(bvg_fit *f, bool backward) {
  if (f->big_tc) return launch_lik_tcbig(f, backward);
  return f->engine == BVG_ENGINE_TCGEN05 ? launch_lik_tc(f, backward) : launch_lik_simt(f, backward);
}
// zeta already holds the point (or is drawn by prep when mode is a draw mode)
int eval_log_prob_device(bvg_fit *f, int mode, int jac, int propto, bool backward) {
  const int prep = mode == MODE_ADVI_UPDATE ? PREP_GRAD : (mode == MODE_ELBO_DRAW ? PREP_ELBO : (mode == MODE_POST_DRAW ? PREP_DRAW : PREP_NONE));
  BVG_TRY(launch_prep(f, prep));
  BVG_TRY(launch_lik(f, backward));
  BVG_TRY(launch_gene_finish(f, mode, jac, propto));
  return BVG_OK;
}

extern "C" int bvg_log_prob(bvg_fit *f, const double *theta, int jacobian, int propto, double *lp, double *grad) {
  BVG_TRY(need_ready(f));
  if (!theta || !lp) { bvg_set_error("null argument"); return BVG_ERR_ARG; }
  const int64_t D = f->m.L.D;
  CUDA_TRY(cudaMemcpyAsync(f->m.zeta, theta, sizeof(double) * D, cudaMemcpyHostToDevice, f->stream));
  BVG_TRY(eval_log_prob_device(f, grad ? MODE_GRAD_OUT : MODE_LP_ONLY, jacobian, propto, grad != nullptr));
  Ctl h;
  CUDA_TRY(cudaMemcpyAsync(&h, f->m.ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, f->stream));
  if (grad) CUDA_TRY(cudaMemcpyAsync(grad, f->m.grad, sizeof(double) * D, cudaMemcpyDeviceToHost, f->stream));
  CUDA_TRY(cudaStreamSynchronize(f->stream));
  BVG_TRY(comm_failed(h));
  *lp = h.lp;
  return BVG_OK;
}

// ---- mean-field ADVI ----------------------------------------------------------------------------
static int advi_reset(bvg_fit *f, const double *d_init /* device, NULL keeps mu */, bool zero_omega) {
  const int64_t D = f->m.L.D;
  if (d_init) CUDA_TRY(cudaMemcpyAsync(f->m.mu, d_init, sizeof(double) * D, cudaMemcpyDeviceToDevice, f->stream));
  if (zero_omega) CUDA_TRY(cudaMemsetAsync(f->m.omega, 0, sizeof(double) * D, f->stream));
  CUDA_TRY(cudaMemsetAsync(f->m.hist, 0, sizeof(double) * 2 * D, f->stream));
  return BVG_OK;
}
static int ctl_set(bvg_fit *f, const Ctl &h0) {
  Ctl h = h0;
  h.es = h.eta / sqrt((double)(h.it < 1 ? 1 : h.it));
  f->h_t_grad = h.t_grad;
  CUDA_TRY(cudaMemcpyAsync(f->m.ctl, &h, sizeof(Ctl), cudaMemcpyHostToDevice, f->stream));
  CUDA_TRY(cudaStreamSynchronize(f->stream));  // h lives on the caller's stack
  return BVG_OK;
}
static int comm_failed(const Ctl &h) {
  if (!h.comm_fail) return BVG_OK;
  bvg_set_error("multi-GPU exchange timed out: a peer rank never delivered its sums (BVG_PEER_TIMEOUT_S raises the limit)");
  return BVG_ERR_COMM;
}
static int ctl_get(bvg_fit *f, Ctl *h) {
  CUDA_TRY(cudaMemcpyAsync(h, f->m.ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, f->stream));
  CUDA_TRY(cudaStreamSynchronize(f->stream));
  return comm_failed(*h);
}
// n consecutive ELBO-gradient evaluations.  tcgen05 engine: only the first one launches k_prep; every
// k_gene draws the next evaluation's zeta itself ("draw-ahead"), and the likelihood kernel builds its
// coefficient rows from zeta, so an evaluation is two launches.
static int advi_steps(bvg_fit *f, int n) {
  const bool ahead = f->engine == BVG_ENGINE_TCGEN05 && !f->big_tc && f->m.k <= BVG_SMALL_K;
  if (n > 0 && persist_ok(f)) {
    // the whole stretch in one cooperative launch (k_advi_persist.cu): no host involvement between evaluations
    BVG_TRY(launch_prep(f, PREP_GRAD));
    BVG_TRY(launch_advi_persist(f, n));
    f->h_t_grad += (uint32_t)n;
    f->st.grad_evals += n;
    return BVG_OK;
  }
  for (int i = 0; i < n; ++i) {
    if (!ahead || i == 0) BVG_TRY(launch_prep(f, PREP_GRAD));
    BVG_TRY(launch_lik(f, true));
    BVG_TRY(launch_gene_finish(f, MODE_ADVI_UPDATE, 1, 1, ahead && i + 1 < n));
  }
  f->st.grad_evals += n;
  return BVG_OK;
}
// ELBO = mean_s log_prob<propto=false, jacobian=true>(zeta_s) + entropy; non-finite draws are dropped
static int advi_elbo(bvg_fit *f, int S, double *out) {
  Ctl h;
  BVG_TRY(ctl_get(f, &h));
  if (f->cfg.elbo_suffstat && f->cfg.likelihood == BVG_GAUSSIAN && !f->cfg.depth_adjust && f->m.k <= BVG_SMALL_K) {
    // all S draws in two launches, from per-gene sufficient statistics (k_suff.cu)
    std::vector<double> lps(S);
    BVG_TRY(suff_elbo(f, S, lps.data()));
    const int64_t D = f->m.L.D;
    BVG_TRY(vec_dot_sum(f->stream, f->m.omega, nullptr, f->scratch + D, D, f->m.tot));
    BVG_TRY(comm_allreduce(f, f->m.tot, 2));
    double hs[2];
    CUDA_TRY(cudaMemcpyAsync(hs, f->m.tot, sizeof(hs), cudaMemcpyDeviceToHost, f->stream));
    h.t_elbo += (uint32_t)S;
    BVG_TRY(ctl_set(f, h));
    f->st.elbo_evals++;
    double acc = 0;
    int ok = 0;
    for (double v : lps) if (std::isfinite(v)) { acc += v; ++ok; }
    if (!ok) { *out = -INFINITY; return BVG_OK; }
    *out = acc / ok + 0.5 * (double)f->m.LG.D * (1.0 + BVG_LOG_TWO_PI) + hs[1];
    if (std::isnan(*out)) *out = -INFINITY;
    return BVG_OK;
  }
  const int64_t D = f->m.L.D;
  if (persist_ok(f) && !g_no_elbo_batch) {
    // fused tcgen05 engine: ONE batched k_prep launch draws all S points (mu / omega do not change between them), ONE
    // cooperative launch (k_elbo_persist) streams the S forward-only passes with no barrier between draws and leaves the S
    // log densities -- local shares on a gene-sharded fit, all-reduced once below
    const int k = f->m.k;
    const size_t nd = (size_t)D + 4 + k + (size_t)f->m.G, nf = (size_t)f->m.Gpad * 32 + 96;
    if (f->eb_cap < S) {
      cudaFree(f->eb_d);
      f->eb_d = nullptr; f->eb_cap = 0;
      CUDA_TRY(cudaMalloc(&f->eb_d, (size_t)S * (nd * sizeof(double) + nf * sizeof(float))));
      CUDA_TRY(cudaMemsetAsync(f->eb_d, 0, (size_t)S * (nd * sizeof(double) + nf * sizeof(float)), f->stream));  // W pads stay zero
      f->eb_cap = S;
    }
    if (f->elbo_cap < S) {
      cudaFree(f->m.elbo_lps);
      f->m.elbo_lps = nullptr; f->elbo_cap = 0;
      CUDA_TRY(cudaMalloc(&f->m.elbo_lps, sizeof(double) * S));
      f->elbo_cap = S;
    }
    double *bz = f->eb_d, *bs = bz + (size_t)f->eb_cap * D, *ba = bs + (size_t)f->eb_cap * (4 + k);
    float *bw = (float *)(ba + (size_t)f->eb_cap * f->m.G), *bu = bw + (size_t)f->eb_cap * f->m.Gpad * 32;
    DevModel &m = f->m;
    double *z0 = m.zeta, *s0 = m.shx, *a0 = m.Aexp;
    float *w0 = m.W, *u0 = m.U;
    m.zeta = bz; m.shx = bs; m.Aexp = ba; m.W = bw; m.U = bu;
    m.elbo_local = f->comm_mode ? 1 : 0;
    m.rank_id = f->cfg.rank;
    int rc = launch_prep_batch(f, PREP_ELBO, S);
    if (rc == BVG_OK) rc = launch_elbo_persist(f, S, m.elbo_lps);
    m.zeta = z0; m.shx = s0; m.Aexp = a0; m.W = w0; m.U = u0;
    m.elbo_local = 0;
    BVG_TRY(rc);
    if (f->comm_mode) BVG_TRY(comm_allreduce(f, f->m.elbo_lps, S));
    std::vector<double> lps(S);
    CUDA_TRY(cudaMemcpyAsync(lps.data(), f->m.elbo_lps, sizeof(double) * S, cudaMemcpyDeviceToHost, f->stream));
    BVG_TRY(vec_dot_sum(f->stream, f->m.omega, nullptr, f->scratch + D, D, f->m.tot));
    BVG_TRY(comm_allreduce(f, f->m.tot, 2));
    double hs[2];
    CUDA_TRY(cudaMemcpyAsync(hs, f->m.tot, sizeof(hs), cudaMemcpyDeviceToHost, f->stream));
    h.t_elbo += (uint32_t)S;
    BVG_TRY(ctl_set(f, h));   // synchronises the stream
    f->st.elbo_evals++;
    double acc = 0;
    int ok = 0;
    for (double v : lps) if (std::isfinite(v)) { acc += v; ++ok; }
    if (!ok) { *out = -INFINITY; return BVG_OK; }
    *out = acc / ok + 0.5 * (double)f->m.LG.D * (1.0 + BVG_LOG_TWO_PI) + hs[1];
    if (std::isnan(*out)) *out = -INFINITY;
    return BVG_OK;
  }
  h.elbo_acc = 0; h.elbo_ok = 0;
  BVG_TRY(ctl_set(f, h));
  if (f->comm_mode) {
    // gene-sharded: every draw leaves its LOCAL share of the log density in elbo_lps[draw] with no exchange; one
    // all-reduce of the S scalars (and of the entropy sum) follows
    if (f->elbo_cap < S) {
      cudaFree(f->m.elbo_lps);
      f->m.elbo_lps = nullptr; f->elbo_cap = 0;
      CUDA_TRY(cudaMalloc(&f->m.elbo_lps, sizeof(double) * S));
      f->elbo_cap = S;
    }
    f->m.elbo_local = 1;
    f->m.rank_id = f->cfg.rank;
    int rc = BVG_OK;
    for (int s = 0; s < S && rc == BVG_OK; ++s) rc = eval_log_prob_device(f, MODE_ELBO_DRAW, 1, 0, false);
    f->m.elbo_local = 0;
    BVG_TRY(rc);
    BVG_TRY(comm_allreduce(f, f->m.elbo_lps, S));
    std::vector<double> lps(S);
    CUDA_TRY(cudaMemcpyAsync(lps.data(), f->m.elbo_lps, sizeof(double) * S, cudaMemcpyDeviceToHost, f->stream));
    BVG_TRY(vec_dot_sum(f->stream, f->m.omega, nullptr, f->scratch + D, D, f->m.tot));
    BVG_TRY(comm_allreduce(f, f->m.tot, 2));
    double hs[2];
    CUDA_TRY(cudaMemcpyAsync(hs, f->m.tot, sizeof(hs), cudaMemcpyDeviceToHost, f->stream));
    CUDA_TRY(cudaStreamSynchronize(f->stream));
    f->st.elbo_evals++;
    double acc = 0;
    int ok = 0;
    for (double v : lps) if (std::isfinite(v)) { acc += v; ++ok; }
    if (!ok) { *out = -INFINITY; return BVG_OK; }
    *out = acc / ok + 0.5 * (double)f->m.LG.D * (1.0 + BVG_LOG_TWO_PI) + hs[1];
    if (std::isnan(*out)) *out = -INFINITY;
    return BVG_OK;
  }
  if (f->engine == BVG_ENGINE_TCGEN05 && !f->big_tc && f->m.k <= BVG_SMALL_K && !g_no_elbo_batch) {
    // fused tcgen05 engine: mu / omega do not change between the S draws, so ONE batched k_prep launch makes all of them
    // (throughput-bound: ~0.2 ms for 150 x 66,045 Philox / Box-Muller / exp chains) instead of S latency-bound launches of
    // ~11 us each; every draw is then two launches (forward-only likelihood pass, per-gene kernel + finish)
    const int k = f->m.k;
    const size_t nd = (size_t)D + 4 + k + (size_t)f->m.G, nf = (size_t)f->m.Gpad * 32 + 96;
    if (f->eb_cap < S) {
      cudaFree(f->eb_d);
      f->eb_d = nullptr; f->eb_cap = 0;
      CUDA_TRY(cudaMalloc(&f->eb_d, (size_t)S * (nd * sizeof(double) + nf * sizeof(float))));
      CUDA_TRY(cudaMemsetAsync(f->eb_d, 0, (size_t)S * (nd * sizeof(double) + nf * sizeof(float)), f->stream));  // W pads stay zero
      f->eb_cap = S;
    }
    double *bz = f->eb_d, *bs = bz + (size_t)f->eb_cap * D, *ba = bs + (size_t)f->eb_cap * (4 + k);
    float *bw = (float *)(ba + (size_t)f->eb_cap * f->m.G), *bu = bw + (size_t)f->eb_cap * f->m.Gpad * 32;
    DevModel &m = f->m;
    double *z0 = m.zeta, *s0 = m.shx, *a0 = m.Aexp;
    float *w0 = m.W, *u0 = m.U;
    m.zeta = bz; m.shx = bs; m.Aexp = ba; m.W = bw; m.U = bu;
    int rc = launch_prep_batch(f, PREP_ELBO, S);
    for (int s = 0; s < S && rc == BVG_OK; ++s) {
      m.zeta = bz + (size_t)s * D; m.shx = bs + (size_t)s * (4 + k); m.Aexp = ba + (size_t)s * m.G;
      m.W = bw + (size_t)s * m.Gpad * 32; m.U = bu + (size_t)s * 96;
      rc = launch_lik(f, false);
      if (rc == BVG_OK) rc = launch_gene_finish(f, MODE_ELBO_DRAW, 1, 0);
    }
    m.zeta = z0; m.shx = s0; m.Aexp = a0; m.W = w0; m.U = u0;
    BVG_TRY(rc);
  } else {
    for (int s = 0; s < S; ++s) BVG_TRY(eval_log_prob_device(f, MODE_ELBO_DRAW, 1, 0, false));
  }
  BVG_TRY(vec_dot_sum(f->stream, f->m.omega, nullptr, f->scratch + D, D, f->m.tot));
  BVG_TRY(comm_allreduce(f, f->m.tot, 2));
  double hs[2];
  CUDA_TRY(cudaMemcpyAsync(hs, f->m.tot, sizeof(hs), cudaMemcpyDeviceToHost, f->stream));
  BVG_TRY(ctl_get(f, &h));
  f->st.elbo_evals++;
  if (h.elbo_ok == 0) { *out = -INFINITY; return BVG_OK; }
  const double Dt = (double)f->m.LG.D;
  *out = h.elbo_acc / h.elbo_ok + 0.5 * Dt * (1.0 + BVG_LOG_TWO_PI) + hs[1];
  if (std::isnan(*out)) *out = -INFINITY;
  return BVG_OK;
}
static int set_phase(bvg_fit *f, double eta) {  // restart the step-size schedule, keep the draw counters
  Ctl h;
  BVG_TRY(ctl_get(f, &h));
  h.it = 1; h.eta = eta; h.bad = 0;
  return ctl_set(f, h);
}

// ---- the variational family behind the driver: meanfield (device-resident control block) or fullrank (k_fullrank.cu) ----
static inline bool is_fr(const bvg_fit *f) { return f->cfg.algorithm == BVG_ALG_FULLRANK; }
static int fam_reset(bvg_fit *f, const double *d_init) { return is_fr(f) ? fr_reset(f, d_init, true) : advi_reset(f, d_init, true); }
static int fam_steps(bvg_fit *f, int n) { return is_fr(f) ? fr_steps(f, n) : advi_steps(f, n); }
static int fam_elbo(bvg_fit *f, int S, double *out) { return is_fr(f) ? fr_elbo(f, S, out) : advi_elbo(f, S, out); }
static int fam_phase(bvg_fit *f, double eta) {
  if (is_fr(f)) { fr_set_phase(f, eta); return BVG_OK; }
  return set_phase(f, eta);
}
static int fam_bad(bvg_fit *f, int *bad) {
  if (is_fr(f)) return fr_bad(f, bad);
  Ctl hh;
  BVG_TRY(ctl_get(f, &hh));
  *bad = hh.bad;
  return BVG_OK;
}

static int run_advi(bvg_fit *f, const double *d_init) {
  const bvg_config &c = f->cfg;
  const int64_t D = f->m.L.D;
  double *init = f->scratch;  // device copy of theta_init
  if (d_init) CUDA_TRY(cudaMemcpyAsync(init, d_init, sizeof(double) * D, cudaMemcpyDeviceToDevice, f->stream));
  else CUDA_TRY(cudaMemsetAsync(init, 0, sizeof(double) * D, f->stream));
  Ctl h;
  memset(&h, 0, sizeof(h));
  h.it = 1;
  BVG_TRY(ctl_set(f, h));
  f->n_trace = 0;
  f->st.grad_evals = 0; f->st.elbo_evals = 0; f->st.converged = 0;
  double eta = c.eta;
  const double t_adapt = now_s();
  if (eta <= 0) {  // adapt_eta
    static const double seq[5] = {100, 10, 1, 0.1, 0.01};
    BVG_TRY(fam_reset(f, init));
    double e0;
    BVG_TRY(fam_elbo(f, c.elbo_samples, &e0));
    if (!std::isfinite(e0)) { bvg_set_error("ADVI: the ELBO is not finite at the initial point"); return BVG_ERR_NUMERIC; }
    double ebest = -INFINITY, eta_best = 0;
    for (int q = 0;; ++q) {
      BVG_TRY(fam_reset(f, init));
      BVG_TRY(fam_phase(f, seq[q]));
      BVG_TRY(fam_steps(f, c.adapt_iter));
      double e;
      BVG_TRY(fam_elbo(f, c.elbo_samples, &e));
      if (c.verbose) fprintf(stderr, "[bayesvg_b200] adapt eta = %-6g ELBO = %.6g (init %.6g)\n", seq[q], e, e0);
      if (e < ebest && ebest > e0) break;
      if (q < 4) { ebest = e; eta_best = seq[q]; }
      else {
        if (e > e0) { ebest = e; eta_best = seq[q]; break; }
        bvg_set_error("ADVI: all proposed step-sizes failed; the model may be ill-conditioned");
        return BVG_ERR_NUMERIC;
      }
    }
    eta = eta_best;
  }
  f->st.eta = eta;
  f->st.sec_adapt = now_s() - t_adapt;
  const double t_sga = now_s();
  BVG_TRY(fam_reset(f, init));
  BVG_TRY(fam_phase(f, eta));
  const int cbn = (int)std::max(0.1 * c.n_iter / c.eval_elbo, 2.0);
  std::vector<double> cb(cbn, 0.0), tmp(cbn);
  int cbc = 0, cbh = 0, conv = 0, it = 0;
  double elbo = 0, elbo_prev;
  while (it < c.n_iter) {
    const int n = std::min(c.eval_elbo - it % c.eval_elbo, c.n_iter - it);
    BVG_TRY(fam_steps(f, n));
    it += n;
    if (it % c.eval_elbo == 0) {
      elbo_prev = elbo;
      BVG_TRY(fam_elbo(f, c.elbo_samples, &elbo));
      int bad = 0;
      BVG_TRY(fam_bad(f, &bad));
      if (bad) { bvg_set_error("ADVI: non-finite gradient during stochastic gradient ascent (%d evaluations)", bad); return BVG_ERR_NUMERIC; }
      const double d = std::fabs((elbo - elbo_prev) / elbo);
      cb[cbh] = d; cbh = (cbh + 1) % cbn; if (cbc < cbn) ++cbc;
      double mean = 0;
      for (int i = 0; i < cbc; ++i) { mean += cb[i]; tmp[i] = cb[i]; }
      mean /= cbc;
      std::sort(tmp.begin(), tmp.begin() + cbc);
      const double med = tmp[cbc / 2];
      if (f->n_trace < 64) { double *r = f->trace + 4 * f->n_trace++; r[0] = it; r[1] = elbo; r[2] = mean; r[3] = med; }
      if (c.verbose) fprintf(stderr, "[bayesvg_b200] %6d  ELBO %-14.6g delta_mean %-10.4g delta_med %-10.4g\n", it, elbo, mean, med);
      if (mean < c.tol_rel_obj) conv |= 1;
      if (med < c.tol_rel_obj) conv |= 2;
      if (conv) break;
    }
  }
  CUDA_TRY(cudaStreamSynchronize(f->stream));
  f->st.advi_iters = it; f->st.converged = conv; f->st.elbo = elbo;
  f->st.sec_sga = now_s() - t_sga;
  f->have_vi = true;
  return BVG_OK;
}

extern "C" int bvg_fit_mle(bvg_fit *f, const double *theta0, double *theta_out) {
  BVG_TRY(need_ready(f));
  const int64_t D = f->m.L.D;
  const double t0 = now_s();
  double *d0 = nullptr;
  if (theta0) {
    CUDA_TRY(cudaMemcpyAsync(f->m.mu, theta0, sizeof(double) * D, cudaMemcpyHostToDevice, f->stream));
    d0 = f->m.mu;
  }
  BVG_TRY(lbfgs_run(f, d0, f->scratch));
  if (!f->h_theta_mle) f->h_theta_mle = new double[D];
  CUDA_TRY(cudaMemcpyAsync(f->h_theta_mle, f->scratch, sizeof(double) * D, cudaMemcpyDeviceToHost, f->stream));
  CUDA_TRY(cudaStreamSynchronize(f->stream));
  if (theta_out) memcpy(theta_out, f->h_theta_mle, sizeof(double) * D);
  f->have_mle = true;
  f->st.sec_mle = now_s() - t0;
  return BVG_OK;
}

extern "C" int bvg_fit_advi(bvg_fit *f, const double *theta_init) {
  BVG_TRY(need_ready(f));
  const int64_t D = f->m.L.D;
  const double *src = theta_init ? theta_init : (f->have_mle ? f->h_theta_mle : nullptr);
  if (src) {
    CUDA_TRY(cudaMemcpyAsync(f->m.grad, src, sizeof(double) * D, cudaMemcpyHostToDevice, f->stream));
    return run_advi(f, f->m.grad);
  }
  return run_advi(f, nullptr);
}

extern "C" int bvg_fit_summarize(bvg_fit *f) {
  BVG_TRY(need_ready(f));
  if (!f->have_vi) { bvg_set_error("no variational fit to summarise"); return BVG_ERR_STATE; }
  const double t0 = now_s();
  if (is_fr(f)) BVG_TRY(fr_amplitude_draws(f));
  BVG_TRY(launch_summary(f, is_fr(f)));
  CUDA_TRY(cudaStreamSynchronize(f->stream));
  f->have_summary = true;
  f->st.sec_draws = now_s() - t0;
  return BVG_OK;
}

extern "C" int bvg_fit_pathfinder_single(bvg_fit *f, int path, double *amp, double *lp_ratio, double *info, double *elbos) {
  BVG_TRY(need_ready(f));
  if (f->cfg.algorithm != BVG_ALG_PATHFINDER || !amp || !lp_ratio || path < 0 || path > 255) { bvg_set_error("bvg_fit_pathfinder_single needs algorithm = pathfinder, 0 <= path <= 255 and output buffers"); return BVG_ERR_ARG; }
  const size_t n = (size_t)f->cfg.pf_single_draws * f->m.G;
  double *d_amp = nullptr;
  CUDA_TRY(cudaMalloc(&d_amp, sizeof(double) * n));
  int rc = pf_single(f, path, d_amp, 0, lp_ratio, info, elbos);
  if (rc == BVG_OK && (cudaMemcpyAsync(amp, d_amp, sizeof(double) * n, cudaMemcpyDeviceToHost, f->stream) != cudaSuccess ||
                       cudaStreamSynchronize(f->stream) != cudaSuccess)) { bvg_set_error("pathfinder: copy failed"); rc = BVG_ERR_CUDA; }
  cudaFree(d_amp);
  return rc;
}
extern "C" int bvg_fit_run(bvg_fit *f) {
  BVG_TRY(need_ready(f));
  const double t0 = now_s();
  if (f->cfg.algorithm == BVG_ALG_PATHFINDER) {   // :366-375: no MLE initialisation (:109), no ADVI
    f->have_mle = false;
    f->st.grad_evals = 0; f->st.elbo_evals = 0;
    const double t1 = now_s();
    BVG_TRY(pf_run(f));
    f->st.sec_sga = now_s() - t1;
    const double t2 = now_s();
    BVG_TRY(launch_summary(f, true));
    CUDA_TRY(cudaStreamSynchronize(f->stream));
    f->have_summary = true;
    f->st.sec_draws = now_s() - t2;
    f->st.sec_total = now_s() - t0 + f->st.sec_h2d;
    return BVG_OK;
  }
  if (f->cfg.mle_init) BVG_TRY(bvg_fit_mle(f, nullptr, nullptr));
  else f->have_mle = false;
  BVG_TRY(bvg_fit_advi(f, nullptr));
  BVG_TRY(bvg_fit_summarize(f));
  f->st.sec_total = now_s() - t0 + f->st.sec_h2d;
  return BVG_OK;
}

// ---- fine-grained hooks -------------------------------------------------------------------------
extern "C" int bvg_fit_set_variational(bvg_fit *f, const double *mu, const double *omega) {
  BVG_TRY(need_ready(f));
  if (!mu) { bvg_set_error("null mu"); return BVG_ERR_ARG; }
  const int64_t D = f->m.L.D;
  CUDA_TRY(cudaMemcpyAsync(f->m.mu, mu, sizeof(double) * D, cudaMemcpyHostToDevice, f->stream));
  if (omega) CUDA_TRY(cudaMemcpyAsync(f->m.omega, omega, sizeof(double) * D, cudaMemcpyHostToDevice, f->stream));
  else CUDA_TRY(cudaMemsetAsync(f->m.omega, 0, sizeof(double) * D, f->stream));
  if (is_fr(f)) BVG_TRY(fr_reset(f, nullptr, true));   // full-rank family: L = I; bvg_fit_set_cholesky installs a factor
  CUDA_TRY(cudaStreamSynchronize(f->stream));
  f->have_vi = true;
  return BVG_OK;
}
extern "C" int bvg_fit_get_cholesky(bvg_fit *f, double *L_packed) {
  BVG_TRY(need_ready(f));
  if (!is_fr(f) || !L_packed) { bvg_set_error("bvg_fit_get_cholesky needs algorithm = fullrank and a buffer of D (D + 1) / 2 doubles"); return BVG_ERR_ARG; }
  return fr_get_cholesky(f, L_packed);
}
extern "C" int bvg_fit_set_cholesky(bvg_fit *f, const double *L_packed) {
  BVG_TRY(need_ready(f));
  if (!is_fr(f) || !L_packed) { bvg_set_error("bvg_fit_set_cholesky needs algorithm = fullrank and D (D + 1) / 2 doubles"); return BVG_ERR_ARG; }
  BVG_TRY(fr_set_cholesky(f, L_packed));
  f->have_vi = true;
  return BVG_OK;
}
extern "C" int bvg_fit_get_variational(bvg_fit *f, double *mu, double *omega) {
  BVG_TRY(need_ready(f));
  const int64_t D = f->m.L.D;
  if (is_fr(f)) BVG_TRY(fr_log_diag(f));   // omega := log |L_ii|
  if (mu) CUDA_TRY(cudaMemcpyAsync(mu, f->m.mu, sizeof(double) * D, cudaMemcpyDeviceToHost, f->stream));
  if (omega) CUDA_TRY(cudaMemcpyAsync(omega, f->m.omega, sizeof(double) * D, cudaMemcpyDeviceToHost, f->stream));
  CUDA_TRY(cudaStreamSynchronize(f->stream));
  return BVG_OK;
}
extern "C" int bvg_fit_advi_steps(bvg_fit *f, int n, double eta, int reset_counters, float *ms_out) {
  BVG_TRY(need_ready(f));
  if (n < 0 || !(eta > 0)) { bvg_set_error("n >= 0 and eta > 0 required"); return BVG_ERR_ARG; }
  Ctl h;
  BVG_TRY(ctl_get(f, &h));
  if (is_fr(f)) h.t_grad = f->h_t_grad;   // full-rank steps count their gradient draws on the host
  if (reset_counters) {
    memset(&h, 0, sizeof(h)); h.it = 1;
    if (is_fr(f)) { BVG_TRY(fr_reset(f, nullptr, false)); fr_set_phase(f, eta); }   // keeps (mu, L), zeroes the step-size history
    else BVG_TRY(advi_reset(f, nullptr, false));
  }
  if (h.it < 1) h.it = 1;
  h.eta = eta;
  BVG_TRY(ctl_set(f, h));
  if (is_fr(f)) BVG_TRY(fr_setup(f));
  CUDA_TRY(cudaEventRecord(f->ev0, f->stream));
  BVG_TRY(fam_steps(f, n));
  CUDA_TRY(cudaEventRecord(f->ev1, f->stream));
  CUDA_TRY(cudaEventSynchronize(f->ev1));
  if (ms_out) CUDA_TRY(cudaEventElapsedTime(ms_out, f->ev0, f->ev1));
  f->have_vi = true;
  return BVG_OK;
}
extern "C" int bvg_fit_elbo(bvg_fit *f, int n_samples, double *elbo_out, float *ms_out) {
  BVG_TRY(need_ready(f));
  if (n_samples < 1 || !elbo_out) { bvg_set_error("bad argument"); return BVG_ERR_ARG; }
  CUDA_TRY(cudaEventRecord(f->ev0, f->stream));
  BVG_TRY(fam_elbo(f, n_samples, elbo_out));
  CUDA_TRY(cudaEventRecord(f->ev1, f->stream));
  CUDA_TRY(cudaEventSynchronize(f->ev1));
  if (ms_out) CUDA_TRY(cudaEventElapsedTime(ms_out, f->ev0, f->ev1));
  return BVG_OK;
}
// L2 eviction: WRITE a buffer larger than L2, then READ it back.  After the write pass the L2 is full of
// dirty lines, whose write-back would be billed to whatever runs next (a 60 MB read would drag 60 MB of
// evictions with it); the read pass leaves only clean lines of the scratch buffer behind.
__global__ void k_flush(float *p, size_t n, int pass) {
  float acc = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    if (pass == 0) p[i] = (float)i;
    else acc += __ldcg(p + i);
  }
  if (pass == 1 && acc == 123.456f) p[0] = acc;  // keep the loads alive
}
static int ensure_flush(bvg_fit *f) {
  if (!f->flush) {
    f->flush_bytes = (size_t)512 << 20;  // > 126 MB L2
    CUDA_TRY(cudaMalloc(&f->flush, f->flush_bytes));
  }
  return BVG_OK;
}
extern "C" int bvg_fit_flush_l2(bvg_fit *f) {
  BVG_TRY(need_ready(f));
  BVG_TRY(ensure_flush(f));
  k_flush<<<148 * 8, 256, 0, f->stream>>>((float *)f->flush, f->flush_bytes / 4, 0);
  k_flush<<<148 * 8, 256, 0, f->stream>>>((float *)f->flush, f->flush_bytes / 4, 1);
  CUDA_TRY(cudaGetLastError());
  return BVG_OK;
}
extern "C" int bvg_fit_time_lik_kernel(bvg_fit *f, int n, int flush_l2, int backward, float *ms_avg_out) {
  BVG_TRY(need_ready(f));
  if (n < 1 || !ms_avg_out) { bvg_set_error("bad argument"); return BVG_ERR_ARG; }
  if (flush_l2 == 2) {   // input larger than L2, no flush inside the timed region
    int ncopy = 4;
    if (const char *e = getenv("BVG_ROT_COPIES")) { const int v = atoi(e); if (v >= 1 && v <= 64) ncopy = v; }
    return tc_time_rotating(f, n, ncopy, backward != 0, ms_avg_out);
  }
  if (flush_l2) BVG_TRY(ensure_flush(f));
  double total = 0;
  if (flush_l2) {
    for (int i = 0; i < n; ++i) {
      k_flush<<<148 * 8, 256, 0, f->stream>>>((float *)f->flush, f->flush_bytes / 4, 0);
      k_flush<<<148 * 8, 256, 0, f->stream>>>((float *)f->flush, f->flush_bytes / 4, 1);
      CUDA_TRY(cudaEventRecord(f->ev0, f->stream));
      BVG_TRY(launch_lik(f, backward != 0));
      CUDA_TRY(cudaEventRecord(f->ev1, f->stream));
      CUDA_TRY(cudaEventSynchronize(f->ev1));
      float ms;
      CUDA_TRY(cudaEventElapsedTime(&ms, f->ev0, f->ev1));
      total += ms;
    }
  } else {
    CUDA_TRY(cudaEventRecord(f->ev0, f->stream));
    for (int i = 0; i < n; ++i) BVG_TRY(launch_lik(f, backward != 0));
    CUDA_TRY(cudaEventRecord(f->ev1, f->stream));
    CUDA_TRY(cudaEventSynchronize(f->ev1));
    float ms;
    CUDA_TRY(cudaEventElapsedTime(&ms, f->ev0, f->ev1));
    total = ms;
  }
  *ms_avg_out = (float)(total / n);
  return BVG_OK;
}

// diagnostics: %globaltimer stamps (ns) inside one k_gene launch: [0] block 0 start, [1] after setup,
// [2] after the gene loop, [3] after the block partials, [4] last block enters finish, [5] after the
// cross-block reduction, [6] after lp / shared update, [7] end
// (stamps are SM cycle counters and are only ordered by data dependencies; for attribution without probe effects
// use bvg_debug_time_gene)
extern "C" int bvg_debug_gene_trace(bvg_fit *f, long long *out16, int ahead) {
  BVG_TRY(need_ready(f));
  long long *d = nullptr;
  CUDA_TRY(cudaMalloc(&d, 16 * sizeof(long long)));
  CUDA_TRY(cudaMemsetAsync(d, 0, 16 * sizeof(long long), f->stream));
  BVG_TRY(launch_prep(f, PREP_GRAD));
  BVG_TRY(launch_lik(f, true));
  f->m.dbg = d;
  // ahead bit 0: draw-ahead; bit 2: gradient-out mode (idempotent); bit 1: run an untraced launch first, so the
  // traced one starts with whatever k_gene left in the instruction / data caches
  const int mode = (ahead & 4) ? MODE_GRAD_OUT : MODE_ADVI_UPDATE;
  if (ahead & 2) { f->m.dbg = nullptr; BVG_TRY(launch_gene_finish(f, mode, 1, 1, 0)); f->m.dbg = d; }
  int rc = launch_gene_finish(f, mode, 1, 1, ahead & 1);
  f->m.dbg = nullptr;
  if (rc == BVG_OK && (cudaMemcpyAsync(out16, d, 16 * sizeof(long long), cudaMemcpyDeviceToHost, f->stream) != cudaSuccess || cudaStreamSynchronize(f->stream) != cudaSuccess)) rc = BVG_ERR_CUDA;
  cudaFree(d);
  return rc;
}

// diagnostics: average device time (ms) of n back-to-back k_gene launches in a given mode, with or without
// the fused finish (fuse = 0: per-gene part only, 3: + cross-gene finish) -- attribution without in-kernel probes
extern int g_gene_fuse_override;
extern "C" int bvg_debug_time_gene(bvg_fit *f, int n, int mode, int fuse, int ahead, float *ms_avg) {
  BVG_TRY(need_ready(f));
  if (n < 1 || !ms_avg || f->comm_mode) { bvg_set_error("bad argument"); return BVG_ERR_ARG; }
  BVG_TRY(launch_prep(f, PREP_GRAD));
  BVG_TRY(launch_lik(f, true));
  g_gene_fuse_override = fuse;
  int rc = launch_gene_finish(f, mode, 1, 1, ahead);
  if (rc == BVG_OK && cudaEventRecord(f->ev0, f->stream) != cudaSuccess) rc = BVG_ERR_CUDA;
  for (int i = 0; i < n && rc == BVG_OK; ++i) rc = launch_gene_finish(f, mode, 1, 1, ahead);
  g_gene_fuse_override = -1;
  BVG_TRY(rc);
  CUDA_TRY(cudaEventRecord(f->ev1, f->stream));
  CUDA_TRY(cudaEventSynchronize(f->ev1));
  float ms;
  CUDA_TRY(cudaEventElapsedTime(&ms, f->ev0, f->ev1));
  *ms_avg = ms / n;
  if (fuse != 3) CUDA_TRY(cudaMemsetAsync(&f->m.ctl->ticket, 0, sizeof(uint32_t), f->stream));
  return BVG_OK;
}

// diagnostics: average device time (ms) of the four stages of one ELBO-gradient evaluation
extern "C" int bvg_debug_stage_times(bvg_fit *f, int n, float *out4) {
  BVG_TRY(need_ready(f));
  cudaEvent_t ev[5];
  for (auto &e : ev) CUDA_TRY(cudaEventCreate(&e));
  double acc[4] = {0, 0, 0, 0};
  Ctl h;
  BVG_TRY(ctl_get(f, &h));
  if (h.it < 1) h.it = 1;
  if (!(h.eta > 0)) h.eta = 0.1;
  BVG_TRY(ctl_set(f, h));
  for (int i = 0; i < n; ++i) {
    CUDA_TRY(cudaEventRecord(ev[0], f->stream));
    BVG_TRY(launch_prep(f, PREP_GRAD));
    CUDA_TRY(cudaEventRecord(ev[1], f->stream));
    BVG_TRY(launch_lik(f, true));
    CUDA_TRY(cudaEventRecord(ev[2], f->stream));
    BVG_TRY(launch_gene_finish(f, MODE_ADVI_UPDATE, 1, 1));
    CUDA_TRY(cudaEventRecord(ev[3], f->stream));
    CUDA_TRY(cudaEventRecord(ev[4], f->stream));
    CUDA_TRY(cudaEventSynchronize(ev[4]));
    for (int q = 0; q < 4; ++q) { float ms; CUDA_TRY(cudaEventElapsedTime(&ms, ev[q], ev[q + 1])); acc[q] += ms; }
  }
  for (int q = 0; q < 4; ++q) out4[q] = (float)(acc[q] / n);
  for (auto &e : ev) cudaEventDestroy(e);
  return BVG_OK;
}

// ---- results ------------------------------------------------------------------------------------
extern "C" int bvg_fit_get_summary(bvg_fit *f, double *out) {
  BVG_TRY(need_ready(f));
  if (!f->have_summary) { bvg_set_error("call bvg_fit_summarize (or bvg_fit_run) first"); return BVG_ERR_STATE; }
  const int64_t G = f->m.G;
  CUDA_TRY(cudaMemcpyAsync(out, f->summary, sizeof(double) * 8 * G, cudaMemcpyDeviceToHost, f->stream));
  CUDA_TRY(cudaStreamSynchronize(f->stream));
  // amplitude_mean_rank: arrange(desc(amplitude_mean)) then row_number()  (:429-430); within a shard
  std::vector<int64_t> idx(G);
  for (int64_t g = 0; g < G; ++g) idx[g] = g;
  // descending, NaN last (dplyr::arrange(desc()) puts NA at the end): a strict weak ordering also when a mean is NaN
  std::stable_sort(idx.begin(), idx.end(), [&](int64_t a, int64_t b) {
    const double x = out[a], y = out[b];
    const bool nx = std::isnan(x), ny = std::isnan(y);
    if (nx || ny) return !nx && ny;
    return x > y;
  });
  for (int64_t r = 0; r < G; ++r) out[7 * G + idx[r]] = (double)(r + 1);
  return BVG_OK;
}
extern "C" int bvg_fit_get_draws(bvg_fit *f, double *out) {
  BVG_TRY(need_ready(f));
  if (!f->have_summary) { bvg_set_error("call bvg_fit_summarize (or bvg_fit_run) first"); return BVG_ERR_STATE; }
  CUDA_TRY(cudaMemcpyAsync(out, f->draws, sizeof(double) * f->m.G * f->cfg.n_draws, cudaMemcpyDeviceToHost, f->stream));
  CUDA_TRY(cudaStreamSynchronize(f->stream));
  return BVG_OK;
}
// ---- full posterior draws (the content of CmdStan's variational output CSV / cmdstanr's $draws()) -----------------
// columns: lp__ (0) | log_p__ | log_g__ | the D parameters CONSTRAINED, Stan declaration order | beta0[G] (hierarchical
// intercept only) | amplitude_sq[G]                                     (approxGP.stan:25-28; SURVEY.md A.4 "output")
static int64_t posterior_ncol(const bvg_fit *f) { return 3 + f->m.L.D + (f->m.depth ? 1 : 2) * f->m.G; }
extern "C" int bvg_fit_posterior_ncol(bvg_fit *f, int64_t *ncol) {
  BVG_TRY(need_ready(f, false));
  if (!ncol) { bvg_set_error("null argument"); return BVG_ERR_ARG; }
  *ncol = posterior_ncol(f);
  return BVG_OK;
}
// one draw: constrain zeta into column-major out[col][d]; per-block partial sums of eps^2 for log_g__
__global__ void __launch_bounds__(256) k_store_draw(DevModel m, int64_t n, int64_t d, double *__restrict__ out, double *__restrict__ part,
                                                    const double *__restrict__ eta /* full-rank family: the draw's standard normals */,
                                                    const double *__restrict__ wgt /* gene-sharded fit: 0 for the replicated gene-shared parameters on ranks > 0 */) {
  __shared__ double red[8];
  const Layout &L = m.L;
  const int64_t D = L.D, G = m.G, ntp = (m.depth ? 1 : 2) * G;
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < D + ntp; i += (int64_t)gridDim.x * blockDim.x) {
    double v;
    if (i < D) {
      const double z = m.zeta[i];
      const bool pos = (!m.depth && i == L.sigma_beta0) || i == L.sigma_amp || (i >= L.amp && i < L.amp + G) ||
                       (i >= L.sigma_alpha && i < L.sigma_alpha + m.k) || i == L.tail;
      v = pos ? exp(z) : z;
      const double e = eta ? eta[i] : (z - m.mu[i]) * exp(-m.omega[i]);
      acc += (wgt ? wgt[i] : 1.0) * e * e;
    } else if (!m.depth && i < D + G) {
      v = m.zeta[L.mu_beta0] + exp(m.zeta[L.sigma_beta0]) * m.zeta[L.z_beta0 + (i - D)];   // beta0[g], approxGP.stan:26
    } else {
      v = exp(2.0 * m.zeta[L.amp + (i - D - (m.depth ? 0 : G))]);                           // amplitude_sq[g], :27
    }
    out[(3 + i) * n + d] = v;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    part[blockIdx.x] = s;
  }
}
__global__ void k_store_draw_finish(const double *__restrict__ part, int nblk, const double *__restrict__ lps, int64_t n, int64_t d,
                                    double *__restrict__ out) {
  double s = 0;
  for (int b = 0; b < nblk; ++b) s += part[b];  // fixed order
  out[d] = 0.0;                 // lp__ is 0 for variational output [U]
  out[n + d] = lps[d];          // log_p__ = log_prob<propto = false, jacobian = true>(zeta_d)
  out[2 * n + d] = -0.5 * s;    // log_g__ = -1/2 sum eps^2: the draw's log density under the approximation up to a constant [U]
}
static const int kStoreBlocks = 148 * 2;
int fr_store_draw(bvg_fit *f, int64_t n, int64_t d, double *d_out, const double *d_eta) {
  double *d_part = d_out + (size_t)posterior_ncol(f) * n;
  k_store_draw<<<kStoreBlocks, 256, 0, f->stream>>>(f->m, n, d, d_out, d_part, d_eta, nullptr);
  k_store_draw_finish<<<1, 1, 0, f->stream>>>(d_part, kStoreBlocks, f->m.elbo_lps, n, d, d_out);
  f->st.kernel_launches += 2;
  CUDA_TRY(cudaGetLastError());
  return BVG_OK;
}
extern "C" int bvg_fit_get_posterior(bvg_fit *f, int n, double *out) {
  BVG_TRY(need_ready(f));
  if (!f->have_vi) { bvg_set_error("no variational fit to draw from"); return BVG_ERR_STATE; }
  if (n < 1 || !out) { bvg_set_error("bad argument"); return BVG_ERR_ARG; }
  // gene-sharded fit: this rank's columns (its slice of the unconstrained vector, beta0 / amplitude_sq of its genes).
  // log_p__ and log_g__ are sums over ALL parameters: every rank forms its local share per draw (the replicated
  // gene-shared parameters count on rank 0 only) and ONE all-reduce of the 2 n scalars follows the n draws.
  const bool sharded = f->comm_mode != 0;
  const int64_t ncol = posterior_ncol(f);
  double *d_out = nullptr, *d_part = nullptr;
  const int nblk = kStoreBlocks;
  CUDA_TRY(cudaMalloc(&d_out, sizeof(double) * (size_t)ncol * n + sizeof(double) * nblk));
  d_part = d_out + (size_t)ncol * n;
  if (f->elbo_cap < n) {
    cudaFree(f->m.elbo_lps);
    f->m.elbo_lps = nullptr; f->elbo_cap = 0;
    if (cudaMalloc(&f->m.elbo_lps, sizeof(double) * n) != cudaSuccess) { cudaFree(d_out); bvg_set_error("out of device memory"); return BVG_ERR_CUDA; }
    f->elbo_cap = n;
  }
  Ctl h;
  int rc = ctl_get(f, &h);
  h.t_draw = 0;
  if (rc == BVG_OK) rc = ctl_set(f, h);
  if (rc == BVG_OK && is_fr(f)) rc = fr_posterior(f, n, d_out);   // zeta = mu + L eta, ten draws per pass over L
  if (sharded) { f->m.elbo_local = 1; f->m.rank_id = f->cfg.rank; }
  for (int d = 0; d < n && rc == BVG_OK && !is_fr(f); ++d) {
    rc = eval_log_prob_device(f, MODE_POST_DRAW, 1, 0, false);
    if (rc != BVG_OK) break;
    k_store_draw<<<nblk, 256, 0, f->stream>>>(f->m, n, d, d_out, d_part, nullptr, sharded ? f->scratch + f->m.L.D : nullptr);
    k_store_draw_finish<<<1, 1, 0, f->stream>>>(d_part, nblk, f->m.elbo_lps, n, d, d_out);
    f->st.kernel_launches += 2;
  }
  f->m.elbo_local = 0;
  if (rc == BVG_OK && sharded) rc = comm_allreduce(f, d_out + n, 2 * n);   // columns log_p__ | log_g__ are contiguous
  if (rc == BVG_OK && (cudaGetLastError() != cudaSuccess ||
                       cudaMemcpyAsync(out, d_out, sizeof(double) * (size_t)ncol * n, cudaMemcpyDeviceToHost, f->stream) != cudaSuccess ||
                       cudaStreamSynchronize(f->stream) != cudaSuccess)) {
    bvg_set_error("posterior draws failed: %s", cudaGetErrorString(cudaGetLastError()));
    rc = BVG_ERR_CUDA;
  }
  cudaFree(d_out);
  return rc;
}

extern "C" int bvg_fit_get_trace(bvg_fit *f, double *out, int cap, int *n) {
  if (!f || !out || !n) { bvg_set_error("null argument"); return BVG_ERR_ARG; }
  const int m = std::min(cap, f->n_trace);
  memcpy(out, f->trace, sizeof(double) * 4 * m);
  *n = m;
  return BVG_OK;
}
extern "C" int bvg_fit_get_stats(bvg_fit *f, bvg_fit_stats *out) {
  if (!f || !out) { bvg_set_error("null argument"); return BVG_ERR_ARG; }
  *out = f->st;
  return BVG_OK;
}

// ---- basis builder ------------------------------------------------------------------------------
extern "C" int bvg_basis_build(int device, const double *coords, int64_t M, const double *centers, int k, double lscale,
                               int kernel, double nu, double period, double *phi_out) {
  if (!coords || !centers || !phi_out) { bvg_set_error("null argument"); return BVG_ERR_ARG; }
  if (M < 1 || k < 1 || k > M) { bvg_set_error("need 1 <= k <= M"); return BVG_ERR_ARG; }
  if (kernel < 0 || kernel > 2) { bvg_set_error("Please provide a valid covariance kernel."); return BVG_ERR_ARG; }
  if (kernel == BVG_KERNEL_MATERN && !(nu == 0.5 || nu == 1.5 || nu == 2.5)) { bvg_set_error("When utilizing the Matern kernel you must provide a valid smoothness parameter value."); return BVG_ERR_ARG; }
  if (!(lscale > 0)) { bvg_set_error("length-scale must be positive"); return BVG_ERR_ARG; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); bvg_set_error("no CUDA device: bayesvg_b200 has no CPU fallback"); return BVG_ERR_CUDA; }
  CUDA_TRY(cudaSetDevice(device));
  double *d = nullptr;
  const size_t n = (size_t)2 * M + 2 * k + (size_t)2 * M * k + 3 * k;
  CUDA_TRY(cudaMalloc(&d, sizeof(double) * n));
  double *dX = d, *dC = dX + 2 * M, *dQ = dC + 2 * k, *dW = dQ + (size_t)M * k;
  cudaStream_t st;
  CUDA_TRY(cudaStreamCreate(&st));
  CUDA_TRY(cudaMemcpyAsync(dX, coords, sizeof(double) * 2 * M, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(dC, centers, sizeof(double) * 2 * k, cudaMemcpyHostToDevice, st));
  int rc = basis_build_device(st, dX, M, dC, k, lscale, kernel, nu, period, dQ, dW);
  if (rc == BVG_OK) {
    if (cudaMemcpyAsync(phi_out, dQ, sizeof(double) * M * k, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) {
      bvg_set_error("basis build failed: %s", cudaGetErrorString(cudaGetLastError()));
      rc = BVG_ERR_CUDA;
    }
  }
  cudaStreamDestroy(st);
  cudaFree(d);
  return rc;
}
