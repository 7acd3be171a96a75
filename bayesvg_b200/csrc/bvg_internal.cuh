// bvg_internal.cuh -- shared declarations of the CUDA engine (not part of the C ABI).
//
// Data layout in HBM:
//   Y     : [G][ldY]  gene-major (= R's M x G column-major), fp64 (PREC_FP64) or fp32 (PREC_FAST),
//           ldY = M rounded up to 64 spots and zero padded in FAST mode.
//   Phi'  : [Mpad][KP] row-major, KP = k+1 rounded up to 8; column k is the all-ones column that
//           carries the intercept beta0[g] forward and R_g = sum_s r backward; padded rows are zero.
//   C     : [Gpad][KP] coefficient rows c'_g = (A_g * alpha_{1..k,g}, beta0_g, 0...) rebuilt each eval.
//   theta-like vectors (mu, omega, hist, zeta, grad): fp64, Stan declaration order.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/bayesvg_b200.h"

#define BVG_SMALL_K 30         // k+2 per-gene parameters map onto one warp: the tuned k_prep / k_gene / sufficient-statistic kernels
#define BVG_MAX_K 510          // general path (k_prep_big / k_gene_big / k_lik_gen): KP = round_up(k+1, 8) <= 512
#define BVG_GENE_TILE 128      // genes per likelihood CTA (= TMEM lanes in the tcgen05 kernel)
#define BVG_GENE_MAX_THREADS 672  // k_gene: up to 21 gene-warps per CTA (one CTA per SM), so 148 SMs hold 3,108 genes in ONE pass
#define BVG_LOG_TWO_PI 1.8378770664093454835606594728112
#define BVG_LOG_PI 1.1447298858494001741434273513531
#define BVG_LOG_TWO 0.69314718055994530941723212145818

enum { STREAM_GRAD = 1, STREAM_ELBO = 2, STREAM_DRAW = 3 };

// slots of the cross-gene reduction vector (the only thing that crosses GPUs each evaluation)
enum { S_R = 0, S_RZ, S_DU, S_DU2, S_SSE /* nb: sum of y*eta - (y+phi)*L */, S_Z2, S_U, S_L, S_LG, S_DG, S_AP };
__host__ __device__ inline int bvg_ns(int k) { return S_AP + 2 * k; }

struct Layout {
  int64_t mu_beta0, sigma_beta0, z_beta0, mu_amp, sigma_amp, amp, mu_alpha, sigma_alpha, z_alpha, tail, D;
};
// inst/approxGP.stan:12-23 and inst/approxGP4.stan:12-23 (declaration order); depth-adjusted programs
// inst/approxGP2.stan:13-23 / approxGP3.stan:13-23: beta0, beta1, sigma_y | phi_nb, amplitude[G], mu_amplitude,
// sigma_amplitude, mu_alpha[k], sigma_alpha[k], z_alpha_t[k,G] -- there mu_beta0 / sigma_beta0 are the slots of
// beta0 / beta1 (beta1 unconstrained) and z_beta0 does not exist.
__host__ __device__ inline Layout make_layout(int lik, int64_t G, int k, int depth) {
  Layout L;
  L.mu_beta0 = 0; L.sigma_beta0 = 1; L.z_beta0 = 2;
  if (depth) {
    L.z_beta0 = -1; L.tail = 2; L.amp = 3; L.mu_amp = 3 + G; L.sigma_amp = 4 + G; L.mu_alpha = 5 + G;
    L.sigma_alpha = L.mu_alpha + k; L.z_alpha = L.sigma_alpha + k;
    L.D = 5 + G + 2 * (int64_t)k + (int64_t)k * G;
    return L;
  }
  if (lik == BVG_GAUSSIAN) {
    L.mu_amp = 2 + G; L.sigma_amp = 3 + G; L.amp = 4 + G; L.mu_alpha = 4 + 2 * G;
    L.sigma_alpha = L.mu_alpha + k; L.z_alpha = L.sigma_alpha + k; L.tail = L.z_alpha + (int64_t)k * G;
  } else {
    L.tail = 2 + G; L.amp = 3 + G; L.mu_amp = 3 + 2 * G; L.sigma_amp = 4 + 2 * G; L.mu_alpha = 5 + 2 * G;
    L.sigma_alpha = L.mu_alpha + k; L.z_alpha = L.sigma_alpha + k;
  }
  L.D = 5 + 2 * G + 2 * (int64_t)k + (int64_t)k * G;
  return L;
}
// compact index of the 2k+5 gene-shared parameters:
//   0 mu_beta0, 1 log sigma_beta0, 2 mu_amplitude, 3 log sigma_amplitude, 4 log tail,
//   5..5+k mu_alpha, 5+k..5+2k log sigma_alpha
__host__ __device__ inline int64_t shared_index(const Layout &L, int k, int i) {
  switch (i) {
    case 0: return L.mu_beta0;
    case 1: return L.sigma_beta0;
    case 2: return L.mu_amp;
    case 3: return L.sigma_amp;
    case 4: return L.tail;
    default: return i < 5 + k ? L.mu_alpha + (i - 5) : L.sigma_alpha + (i - 5 - k);
  }
}

// device-resident control block: everything that changes between iterations lives here so the
// same launch sequence can be replayed without host involvement
struct Ctl {
  uint32_t t_grad;   // running counter of gradient draws   (Philox counter word 2, stream 1)
  uint32_t t_elbo;   // running counter of ELBO draws        (stream 2)
  int32_t it;        // 1-based iteration inside the current phase (step-size schedule eta/sqrt(it))
  int32_t bad;       // number of evaluations whose gradient was non-finite
  double eta;
  double lp;         // last log density
  double elbo_acc;   // sum of finite log densities of the current ELBO estimate
  int32_t elbo_ok;
  uint32_t ticket;   // last-block-done counter of k_gene (fused finish)
  double es;         // eta / sqrt(it): kept next to `it` so that k_gene starts without a sqrt + divide chain
  uint32_t t_draw;   // running counter of posterior draws (stream 3; bvg_fit_get_posterior)
  uint32_t comm_fail; // set by a peer exchange (or the persistent ADVI kernel) whose partner never arrived: BVG_ERR_COMM on the host
};

// ---- peer-memory exchange (multi-GPU) ------------------------------------------
// Every rank owns one exchange buffer, mapped into every peer with CUDA IPC (one process per GPU).
// Layout: cells[2 parities][world ranks][2 * BVG_PEER_CAP] of 8-byte cells {32 data bits, 32-bit
// sequence number}: a cell is valid exactly when its flag equals the current sequence number, so one
// 8-byte store over NVLink carries data + arrival flag and no fence / second round trip is needed.
#define BVG_MAX_PEERS 8
#define BVG_PEER_CAP 1024  // doubles per exchange
struct PeerCtx {
  int world, rank;
  uint32_t *seq;                 // exchanges completed so far (advances in lockstep on every rank)
  uint2 *cells[BVG_MAX_PEERS];   // cells[r] = rank r's exchange buffer as mapped here (cells[rank] is local)
  long long timeout;             // SM cycles a poll may wait for a peer (BVG_PEER_TIMEOUT_S, default 120 s) before it gives up
  uint32_t *fail;                // -> Ctl::comm_fail of this rank
};

struct DevModel {
  int lik, k, KP, nsplit;
  int depth, pad0;           // depth-adjusted program (approxGP2/3): the intercept "z" of gene g is the datum depths[g]
  const double *depths;      // [G] gene_depths
  int64_t M, G, Gpad, gene_offset, G_total;
  Layout L, LG;  // local layout (this rank's genes) and global layout (RNG indices)
  uint64_t seed;
  double *mu, *omega, *hist, *zeta, *grad;
  Ctl *ctl;
  void *C;       // [Gpad][KP]
  void *Ppart;   // [nsplit][Gpad][KP]
  void *Spart;   // [nsplit][Gpad][2], always fp64
  double *blk;   // [nblk_gene][NS]
  double *tot;   // [NS]
  double *shx;   // [4+k] exp of the shared log-scales of the current draw: sigma_beta0, sigma_amplitude, tail, sigma_alpha[k];
                 //       [3+k] = 1 / sigma_y^2 (gaussian) or 1 (nb)
  double *Aexp;  // [G] amplitude_sq of the current draw
  // tcgen05 engine: the current draw as the likelihood kernel wants it, fp32.  W[g][0..k) = z_alpha_t[,g],
  // W[g][k] = z_beta0[g], W[g][k+1] = amplitude_sq[g]; U[0..k) = mu_alpha, U[32..32+k) = sigma_alpha,
  // U[64] = mu_beta0, U[65] = sigma_beta0.  The kernel forms c'_g = (A_g (mu_alpha + sigma_alpha z), mu_beta0 + sigma_beta0 z_b).
  float *W;      // [Gpad][32]
  float *U;      // [96]
  int nblk_gene;
  int nhist;                 // nb: distinct count values
  const double *hist_val;    // nb: value c
  const double *hist_cnt;    // nb: #pairs with y == c
  double lgam_y1;            // nb: sum lgamma(y+1) over local pairs
  double *hterm;             // nb, tcgen05 engine: [nhist][2] per-count terms written by the likelihood kernel (else NULL)
  const double *suffU;       // [G][k+1] Phi'^T y_g   (sufficient-statistic ELBO, gaussian)
  const double *suffS2;      // [G] sum_s y^2
  const double *suffGram;    // [(k+1)^2] Phi'^T Phi'
  long long *dbg;            // diagnostics: %globaltimer stamps (NULL unless bvg_debug_* is running)
  PeerCtx peer;              // world <= 1: no in-kernel exchange
  // gene-sharded ELBO estimates: a draw's log density is linear in the cross-gene sums, so every rank forms its
  // LOCAL share (gene-shared prior terms on rank 0 only) with no exchange per draw, stores it in elbo_lps[draw],
  // and ONE all-reduce of the S scalars follows the S draws.
  int elbo_local, rank_id;
  double *elbo_lps;
};

// Sum-all-reduce of vals[0..n) (global memory, n <= min(BVG_PEER_CAP, blockDim.x)) across the ranks of p,
// executed by ALL threads of ONE CTA.  Push: every 8-byte cell goes straight into each peer's buffer;
// pull: poll the local buffer until every peer's cells carry this exchange's sequence number.  The sum
// runs in rank order on every rank, so all ranks end with bit-identical results (the replicated
// gene-shared parameters stay identical without a broadcast).  Parity double-buffering is enough: a
// rank can only be one exchange ahead of a peer it is waiting for.
// A peer that never arrives (crashed rank, a rank stuck in host code for longer than the time-out) must not hang
// the GPU: the poll gives up, raises Ctl::comm_fail (the host turns it into BVG_ERR_COMM at its next read of the
// control block) and returns NaN so that nothing downstream looks like a result.
__device__ inline double peer_poll(const PeerCtx &p, const uint2 *src, uint32_t seq, long long t0) {
  uint32_t lo, fl, hi, fh;
  for (;;) {
    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(fl) : "l"(src) : "memory");
    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(hi), "=r"(fh) : "l"(src + 1) : "memory");
    if (fl == seq && fh == seq) break;
    if (clock64() - t0 > p.timeout) {
      if (p.fail) *(volatile uint32_t *)p.fail = 1u;
      return __longlong_as_double(0x7ff8000000000000LL);
    }
  }
  return __hiloint2double((int)hi, (int)lo);
}
// scratch: shared memory for (world - 1) * n doubles, or NULL.  With scratch every (peer, value) pair is polled by
// its own thread, so the wait is one round trip instead of world - 1 of them in sequence.
__device__ inline void peer_allreduce_cta(const PeerCtx &p, double *vals, int n, double *scratch) {
  __shared__ uint32_t s_seq;
  __syncthreads();  // the caller's writes to vals[] (same CTA) are visible
  if (threadIdx.x == 0) s_seq = *p.seq + 1;
  __syncthreads();
  const uint32_t seq = s_seq;
  const size_t per_rank = 2 * BVG_PEER_CAP, slot = (size_t)(seq & 1u) * p.world * per_rank;
  for (int i = threadIdx.x; i < 2 * n; i += blockDim.x) {
    const double v = vals[i >> 1];
    const uint32_t half = (i & 1) ? (uint32_t)__double2hiint(v) : (uint32_t)__double2loint(v);
    for (int r = 0; r < p.world; ++r) {
      if (r == p.rank) continue;
      uint2 *dst = p.cells[r] + slot + (size_t)p.rank * per_rank + i;
      asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(dst), "r"(half), "r"(seq) : "memory");
    }
  }
  const long long t0 = clock64();
  const uint2 *mine = p.cells[p.rank] + slot;
  double s = 0.0;
  const int c = threadIdx.x;
  if (scratch) {
    for (int t = threadIdx.x; t < n * (p.world - 1); t += blockDim.x) {
      const int ri = t / n, cc = t - ri * n, r = ri < p.rank ? ri : ri + 1;
      scratch[t] = peer_poll(p, mine + (size_t)r * per_rank + 2 * cc, seq, t0);
    }
    __syncthreads();
    if (c < n)
      for (int r = 0; r < p.world; ++r) s += r == p.rank ? vals[c] : scratch[(r < p.rank ? r : r - 1) * n + c];  // rank order
  } else if (c < n) {
    for (int r = 0; r < p.world; ++r) s += r == p.rank ? vals[c] : peer_poll(p, mine + (size_t)r * per_rank + 2 * c, seq, t0);
  }
  __syncthreads();  // every push has read vals[] before it is overwritten
  if (c < n) vals[c] = s;
  if (threadIdx.x == 0) *p.seq = seq;
  __syncthreads();
}

// ---- Philox4x32-10 + Box-Muller: the RNG specification of DESIGN.md section 5 ----------------
__host__ __device__ inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4]) {
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__host__ __device__ inline double bvg_normal(uint64_t seed, uint32_t stream, uint32_t t, uint64_t i) {
  uint32_t o[4];
  philox4x32_10((uint32_t)i, (uint32_t)(i >> 32), t, stream, (uint32_t)seed, (uint32_t)(seed >> 32), o);
  const double u1 = ((double)(((uint64_t)o[0] << 21) | (o[2] >> 11)) + 0.5) * (1.0 / 9007199254740992.0);
  const double u2 = ((double)o[1] + 0.5) * (1.0 / 4294967296.0);
  return sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925286766559 * u2);
}

// Programmatic dependent launch: every kernel of the evaluation chain is launched with the
// programmatic-stream-serialization attribute, so its CTAs may start (and run their prologue) while the
// previous kernel drains; griddep_wait() blocks until the previous grid has completed and its writes
// are visible.  launch_dependents only says "my successor may start launching".
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

__device__ inline double dev_digamma(double x) {
  double r = 0.0;
  while (x < 10.0) { r -= 1.0 / x; x += 1.0; }
  const double f = 1.0 / (x * x);
  const double t = f * (-1.0 / 12.0 + f * (1.0 / 120.0 + f * (-1.0 / 252.0 + f * (1.0 / 240.0 + f * (-1.0 / 132.0 + f * (691.0 / 32760.0 + f * (-1.0 / 12.0)))))));
  return r + log(x) - 0.5 / x + t;
}

__device__ inline double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Log density from the cross-gene sums tot[] and the gene-shared entries of the point zt (full unconstrained layout) with their
// exponentials shx (DevModel::shx): the last-warp job of finish_body, and the ELBO mode of the persistent kernel.
// inst/approxGP.stan:37-49 / approxGP4.stan:37-49 / approxGP2-3.stan:36-48 of the reference.  Warp-collective (all 32 lanes of
// one warp call it); the value is valid on lane 0.  loc: local share of a gene-sharded draw (gene counts of this shard only,
// terms of the gene-shared parameters weighted by shw = 1 on rank 0, 0 elsewhere).
__device__ inline double bvg_lp_from_sums(const DevModel &m, const double *tot, const double *zt, const double *shx, int jac,
                                          int propto, bool loc, double shw, int lane) {
  const int k = m.k;
  const Layout &L = m.L;
  const double mb = zt[L.mu_beta0], lsb = zt[L.sigma_beta0], sb = shx[0];
  const double ma = zt[L.mu_amp], lsa = zt[L.sigma_amp], sa = shx[1];
  const double lt = zt[L.tail], tl = shx[2];
  const double Gt = loc ? (double)m.G : (double)m.G_total, N = (double)m.M * Gt;
  const double j1 = jac ? 1.0 : 0.0;
  double pa = 0;  // prior terms of mu_alpha, sigma_alpha + jacobian of sigma_alpha
  for (int j = lane; j < k; j += 32) {
    const double a = zt[L.mu_alpha + j], ls = zt[L.sigma_alpha + j], s = shx[3 + j];
    pa += -a * a / 8.0 - 0.5 * s * s + j1 * ls;
  }
  pa = warp_sum(pa);
  double lp, lsh;  // lp: terms carrying cross-gene sums or gene counts; lsh: gene-shared parameters only
  if (m.lik == BVG_GAUSSIAN) { lp = -N * lt - 0.5 * tot[S_SSE] / (tl * tl); lsh = -0.5 * tl * tl; }
  else { lp = tot[S_SSE] + tot[S_LG]; lsh = -log1p(tl * tl / 4.0); }
  if (m.depth && m.lik == BVG_GAUSSIAN) lsh += 0.5 * tl * tl - tl * tl / 8.0;  // sigma_y ~ normal(0, 2) (approxGP2.stan:44)
  lp += -0.5 * tot[S_Z2];
  lsh += -mb * mb / 8.0 - (m.depth ? sb * sb / 8.0 : 0.5 * sb * sb) - ma * ma / 8.0 - 0.5 * sa * sa;
  lp += -Gt * lsa - tot[S_U] - 0.5 * tot[S_DU2] / (sa * sa);
  lsh += pa;
  if (jac) { lp += tot[S_U]; lsh += (m.depth ? 0.0 : lsb) + lsa + lt; }
  if (!propto) {
    if (!m.depth) {
      if (m.lik == BVG_GAUSSIAN) {
        lp += -0.5 * BVG_LOG_TWO_PI * (N + (double)k * Gt + 2.0 * Gt);
        lsh += -0.5 * BVG_LOG_TWO_PI * (2.0 * k + 5.0) - (k + 2.0) * BVG_LOG_TWO;
      } else {
        lp += -0.5 * BVG_LOG_TWO_PI * ((double)k * Gt + 2.0 * Gt);
        lsh += -0.5 * BVG_LOG_TWO_PI * (2.0 * k + 4.0) - (k + 2.0) * BVG_LOG_TWO - BVG_LOG_PI;
      }
    } else {  // approxGP2/3: no z_beta0 terms; normal(0, 2) on beta0, beta1, mu_amplitude, mu_alpha[k] (and sigma_y)
      if (m.lik == BVG_GAUSSIAN) {
        lp += -0.5 * BVG_LOG_TWO_PI * (N + (double)k * Gt + Gt);
        lsh += -0.5 * BVG_LOG_TWO_PI * (2.0 * k + 5.0) - (k + 4.0) * BVG_LOG_TWO;
      } else {
        lp += -0.5 * BVG_LOG_TWO_PI * ((double)k * Gt + Gt);
        lsh += -0.5 * BVG_LOG_TWO_PI * (2.0 * k + 4.0) - (k + 3.0) * BVG_LOG_TWO - BVG_LOG_PI;
      }
    }
  }
  return lp + shw * lsh;
}

// trial point of the device-resident L-BFGS line search (lbfgs_dev.cu): k_prep forms zeta = x0 + *a * p itself when x0 != NULL
struct LsPoint { const double *x0, *p, *a; };

// ---- host-side fit object ---------------------------------------------------------------------
struct NcclApi;
struct bvg_fit {
  bvg_config cfg;
  int engine;  // resolved bvg_engine
  cudaStream_t stream;
  cudaEvent_t ev0, ev1;
  DevModel m;
  int64_t ldY, Mpad;
  void *Y;             // T [G][ldY]
  size_t Y_bytes;
  void *Phi;           // T [Mpad][KP]
  double *phi64;       // k x M? kept for basis-independent checks: col-major M x k copy
  double *dvec;        // backing store of mu|omega|hist(2D)|zeta|grad
  double *scratch;     // D-length scratch vectors for L-BFGS etc.
  double *summary;     // [8][G] device
  double *draws;       // [G][n_draws] device
  void *flush;         // L2 flush buffer
  size_t flush_bytes;
  void *stage_buf;     // two H2D staging buffers of set_y (kept across calls)
  size_t stage_bytes;
  cudaStream_t copy_stream;
  cudaEvent_t ev_copied[2], ev_staged[2];
  bool have_phi, have_y, have_mle, have_vi, have_summary, have_depths;
  double *h_depths;    // host copy of gene_depths (uploaded when the model is finalised)
  int64_t n_depths;
  double *h_theta_mle;
  // trace / stats
  double trace[4 * 64];
  int n_trace;
  bvg_fit_stats st;
  // tcgen05 engine state (k_lik_tc.cu)
  void *tc;
  void *tcbig;         // large-basis tcgen05 engine state (k_lik_tcbig.cu); non-NULL selects it
  bool big_tc;         // engine == TCGEN05 with k + 1 > 24
  int gene_wpb;        // gene-warps per k_gene CTA
  uint32_t h_t_grad;   // host mirror of ctl->t_grad (lets k_gene draw ahead before its PDL wait)
  double *suff;        // sufficient statistics + scratch (k_suff.cu), lazily built
  bool tiledY;         // Y is stored tile-major (tcgen05 engine)
  // comm: 0 = single GPU, 1 = NCCL all-reduce between kernels, 2 = in-kernel exchange over peer memory
  int comm_mode;
  NcclApi *nccl;
  void *comm;
  void *fr;            // full-rank ADVI state (k_fullrank.cu), allocated on first use
  int *pf_idx;         // pathfinder: indices of the PSIS-resampled draws (host)
  double *eb_d;        // batched ELBO draws: [cap][D | 4 + k | G] doubles, then [cap][Gpad * 32 | 96] floats
  int eb_cap;
  int elbo_cap;        // capacity of m.elbo_lps
  void *peer_buf;      // this rank's exchange buffer (cudaMalloc, exported with cudaIpcGetMemHandle)
  void *peer_map[BVG_MAX_PEERS];  // peers' buffers opened with cudaIpcOpenMemHandle
  bool no_persist;     // diagnostics / tests (BVG_NO_PERSIST=1): keep the two-launch evaluation chain instead of k_advi_persist
  void *shadow;        // fp64 evaluator of a fast-precision fit (fp64_shadow.cu), built when the L-BFGS line search hits the fast engines' noise floor
  LsPoint ls;          // non-NULL x0 only while lbfgs_run_device() is enqueueing evaluations
  // work space of lbfgs_run_device(), kept with the fit: cudaFree / cudaFreeHost of a 27 MB block cost anything between 1 and
  // 400 ms on these boxes (measured: the "teardown" of an L-BFGS run), so the block is freed with the model, not per run
  void *lb_dev_buf; size_t lb_dev_bytes;
  void *lb_pinned;     // 512 bytes of pinned host memory (status words)
  cudaEvent_t lb_evt[2];
  bool no_mle_fp64;    // diagnostics (BVG_NO_MLE_FP64=1): let the fast path's L-BFGS end with TERM_LSFAIL as in round 1
};

// Device allocations of a fit come from the device's stream-ordered memory pool (cudaMallocAsync / cudaFreeAsync, release
// threshold 8 GB) instead of cudaMalloc / cudaFree: a released block stays mapped for the next allocation, where cudaFree /
// cudaMalloc cost the driver anything from 0.1 to 700 ms per call on these boxes (bench.py: `sec_close`, `runs_detail`).
// Same semantics as the calls they stand in for (the free waits for the device first, as cudaFree does).  comm.cu opts out:
// the peer-exchange buffer is exported with cudaIpcGetMemHandle, which needs a cudaMalloc block.
cudaError_t bvg_pool_malloc(void **p, size_t bytes);   // api.cu
cudaError_t bvg_pool_free(void *p);
#ifndef BVG_NO_POOL
#define cudaMalloc(pp, n) bvg_pool_malloc((void **)(pp), (n))
#define cudaFree(p) bvg_pool_free((void *)(p))
#endif

void bvg_set_error(const char *fmt, ...);
#define CUDA_TRY(x)                                                                          \
  do {                                                                                       \
    cudaError_t e_ = (x);                                                                    \
    if (e_ != cudaSuccess) {                                                                 \
      bvg_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_));       \
      return BVG_ERR_CUDA;                                                                   \
    }                                                                                        \
  } while (0)
#define BVG_TRY(x)              \
  do {                          \
    int r_ = (x);               \
    if (r_ != BVG_OK) return r_; \
  } while (0)

// launchers (one per translation unit)
int launch_prep(bvg_fit *f, int mode);                       // k_model.cu: draw + transform -> C
int launch_prep_batch(bvg_fit *f, int mode, int B);          // k_model.cu: B draws in one launch (fused tcgen05 engine)
int launch_lik(bvg_fit *f, bool backward);                   // k_lik_simt.cu / k_lik_tc.cu dispatch
int launch_lik_simt(bvg_fit *f, bool backward);
int launch_lik_tc(bvg_fit *f, bool backward);
int launch_lik_tcbig(bvg_fit *f, bool backward);             // k_lik_tcbig.cu
int tcbig_setup(bvg_fit *f);
void tcbig_destroy(bvg_fit *f);
bool tc_available();                                         // false until the tcgen05 kernel is linked in
int tc_setup(bvg_fit *f);                                    // build TMA descriptors etc.
int tc_time_rotating(bvg_fit *f, int n, int ncopy, bool bwd, float *ms_avg);  // measurement: back-to-back launches over rotating copies of Y (> L2)
void tc_destroy(bvg_fit *f);
bool persist_ok(bvg_fit *f);                                 // k_advi_persist.cu: a stretch of ADVI iterations in one cooperative launch
int launch_advi_persist(bvg_fit *f, int n_iter);             //   (the current draw must be in place: launch_prep(PREP_GRAD) first)
int launch_elbo_persist(bvg_fit *f, int S, double *d_lps);   //   all S forward-only draws of an ELBO estimate in one cooperative launch
void persist_destroy(bvg_fit *f);
int shadow_enter(bvg_fit *f);                                // fp64_shadow.cu: evaluate in fp64 (SIMT kernels, Y widened on the device) until shadow_leave
void shadow_leave(bvg_fit *f);
void shadow_destroy(bvg_fit *f);
int bvg_big_alloc(int device, void **p, size_t bytes);       // api.cu: exact-size cache of the large per-fit device blocks
void bvg_big_free(int device, void *p, size_t bytes);
void lbfgs_cache_destroy(bvg_fit *f);                        // lbfgs_dev.cu: the cached work space of lbfgs_run_device
bool shadow_active(const bvg_fit *f);
int gene_big_setup();                                        // k_model.cu: per-device function attributes of k_gene_big
int launch_gene_finish(bvg_fit *f, int mode, int jacobian, int propto, int ahead = 0);  // per-gene backward (+ update, + next draw) + cross-gene finish
int comm_allreduce(bvg_fit *f, double *buf, int n);          // comm.cu
int suff_elbo(bvg_fit *f, int S, double *lps_host);          // k_suff.cu: S log densities from sufficient statistics
void suff_destroy(bvg_fit *f);
int eval_log_prob_device(bvg_fit *f, int mode, int jac, int propto, bool backward);  // api.cu
// full-rank ADVI (k_fullrank.cu): q = N(mu, L L^T), L packed lower-triangular in HBM
int fr_setup(bvg_fit *f);
void fr_destroy(bvg_fit *f);
int fr_reset(bvg_fit *f, const double *d_init, bool reset_L);
void fr_set_phase(bvg_fit *f, double eta);
int fr_bad(bvg_fit *f, int *bad);
int fr_steps(bvg_fit *f, int n);
int fr_elbo(bvg_fit *f, int S, double *out);
int fr_amplitude_draws(bvg_fit *f);
int fr_posterior(bvg_fit *f, int n, double *d_out);
int fr_get_cholesky(bvg_fit *f, double *L_out);
int fr_set_cholesky(bvg_fit *f, const double *L_in);
int fr_log_diag(bvg_fit *f);
// Pathfinder (k_pathfinder.cu)
int pf_single(bvg_fit *f, int path, double *d_amp, int64_t row0, double *h_lr, double *info, double *elbos);
int pf_run(bvg_fit *f);
int fr_store_draw(bvg_fit *f, int64_t n, int64_t d, double *d_out, const double *d_eta);  // api.cu: one row of the posterior matrix
int launch_summary(bvg_fit *f, bool given = false);   // k_summary.cu; given: summarise the draws already in f->draws
// MODE_POST_DRAW: log_p__ of a posterior draw (Philox stream 3, counter ctl->t_draw), stored in elbo_lps[t_draw]
enum { MODE_GRAD_OUT = 0, MODE_ADVI_UPDATE = 1, MODE_ELBO_DRAW = 2, MODE_LP_ONLY = 3, MODE_POST_DRAW = 4 };
enum { PREP_NONE = 0, PREP_GRAD = 1, PREP_ELBO = 2, PREP_DRAW = 3 };
