// k_fullrank.cu -- algorithm = "fullrank" (R/findSpatiallyVariableFeaturesBayes.R:108, :357-364 with algorithm =
// "fullrank"): ADVI with a full-rank Gaussian q = N(mu, L L^T) [U: stan::variational::normal_fullrank].
//
//   zeta = mu + L eta,  grad mu = g,  grad L_ij = g_i eta_j (j <= i) + [i == j] / L_ii,  entropy = D/2 (1 + log 2 pi) + sum log |L_ii|
//   and Stan's step-size sequence s <- 0.9 s + 0.1 grad^2, x += eta / sqrt(it) * grad / (1 + sqrt(s)) elementwise on (mu, L).
//
// Layout: L and its step-size history H are PACKED lower-triangular, row-major (row i at i (i + 1) / 2), fp64, both
// resident in HBM: D (D + 1) / 2 * 16 bytes = 34.9 GB at BASELINE config 2 (D = 66,045) -- a size at which the
// reference's CPU path (two dense D x D Eigen matrices, 70 GB, O(D^2) per iteration on one thread) is not practical
// and one 180 GB B200 is.  The log density / gradient is the same likelihood engine as for meanfield; what is new is
// O(D^2) streaming work, HBM-bound by construction:
//   k_fr_update : ONE pass over (L, H) per iteration does the update of row i AND the next iteration's transform
//                 zeta'_i = mu'_i + sum_j L'_ij eta'_j (32 bytes of traffic per entry, the algorithmic minimum for
//                 read + write of two fp64 arrays); rows are dealt round-robin to CTAs so every CTA gets the same
//                 mix of short and long rows.
//   k_fr_gemv   : zeta_b = mu + L eta_b for a batch of up to FR_NB draws per pass over L (ELBO estimates: 150 draws
//                 in 15 passes instead of 150; posterior draws likewise; summaries only touch the amplitude rows).
// Gene sharding is not supported here: L couples every pair of parameters.
#include <vector>

#include "bvg_internal.cuh"
#include "bvg_vec.cuh"

#define FR_NB 10        // draws per pass over L
#define FR_THREADS 256
#define FR_U 4           // entries per thread and trip in the update pass

struct FrState {
  double *L, *H;        // packed [D (D + 1) / 2]
  double *eta[2];       // current / next gradient draw
  double *etab, *zb;    // [FR_NB][D] batch draws and their transforms
  int *ok;              // [0] 1 = finite gradient, [1] count of failed evaluations
  size_t n;
  int it;               // 1-based iteration inside the current phase
  double step;          // eta of the phase
  int cur;              // which eta buffer holds the draw that produced m.zeta
};

static inline size_t tri(int64_t D) { return (size_t)D * (size_t)(D + 1) / 2; }

__device__ inline double fr_block_sum(double v, double *red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0;
#pragma unroll
  for (int w = 0; w < FR_THREADS / 32; ++w) s += red[w];
  return s;
}

__global__ void k_fr_init(double *__restrict__ L, double *__restrict__ H, int64_t D, int reset_L) {
  const size_t n = (size_t)D * (size_t)(D + 1) / 2;
  for (int64_t i = blockIdx.x; i < D; i += gridDim.x) {
    const size_t r0 = (size_t)i * (size_t)(i + 1) / 2;
    for (int64_t j = threadIdx.x; j <= i; j += blockDim.x) {
      H[r0 + j] = 0.0;
      if (reset_L) L[r0 + j] = j == i ? 1.0 : 0.0;   // L = I [U: normal_fullrank(cont_params)]
    }
  }
  (void)n;
}
// out[b][i] = normal(seed, stream, t0 + b, i)
__global__ void k_fr_normals(uint64_t seed, uint32_t stream, uint32_t t0, int nb, int64_t D, double *__restrict__ out) {
  const int64_t n = (int64_t)nb * D;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = q / D, i = q - b * D;
    out[q] = bvg_normal(seed, stream, t0 + (uint32_t)b, (uint64_t)i);
  }
}
// zb[b][i] = mu[i] + sum_{j <= i} L[i][j] etab[b][j] for rows r0 <= i < r1, b < nb <= FR_NB.
// A CTA takes FR_R consecutive rows at a time: a thread loads the nb normals of its column j ONCE and uses them for
// all FR_R rows (the first version, one row per CTA, re-read them per row: 80 bytes of L2 traffic per 8-byte entry of
// L, 27 ms per pass at D = 66,045 where HBM allows 3); L is streamed past L2 (ld.global.cs).
#define FR_R 4
__global__ void __launch_bounds__(FR_THREADS, 2) k_fr_gemv(const double *__restrict__ L, int64_t D, int64_t r0, int64_t r1,
                                                           const double *__restrict__ mu, const double *__restrict__ etab,
                                                           double *__restrict__ zb, int nb) {
  __shared__ double red[FR_THREADS / 32][FR_R][FR_NB];
  for (int64_t ib = r0 + (int64_t)blockIdx.x * FR_R; ib < r1; ib += (int64_t)gridDim.x * FR_R) {
    const int nrow = (int)(r1 - ib < FR_R ? r1 - ib : FR_R);
    const int64_t imax = ib + nrow - 1;
    const double *rowp[FR_R];
#pragma unroll
    for (int r = 0; r < FR_R; ++r) rowp[r] = L + (size_t)(ib + r) * (size_t)(ib + r + 1) / 2;
    double acc[FR_R][FR_NB];
#pragma unroll
    for (int r = 0; r < FR_R; ++r)
#pragma unroll
      for (int b = 0; b < FR_NB; ++b) acc[r][b] = 0.0;
    for (int64_t j = threadIdx.x; j <= imax; j += FR_THREADS) {
      double l[FR_R];
#pragma unroll
      for (int r = 0; r < FR_R; ++r) l[r] = (r < nrow && j <= ib + r) ? __ldcs(rowp[r] + j) : 0.0;
      double e[FR_NB];
#pragma unroll
      for (int b = 0; b < FR_NB; ++b) e[b] = b < nb ? etab[(size_t)b * D + j] : 0.0;
#pragma unroll
      for (int r = 0; r < FR_R; ++r)
#pragma unroll
        for (int b = 0; b < FR_NB; ++b) acc[r][b] = fma(l[r], e[b], acc[r][b]);
    }
#pragma unroll
    for (int r = 0; r < FR_R; ++r)
#pragma unroll
      for (int b = 0; b < FR_NB; ++b) acc[r][b] = warp_sum(acc[r][b]);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
      for (int r = 0; r < FR_R; ++r)
#pragma unroll
        for (int b = 0; b < FR_NB; ++b) red[threadIdx.x >> 5][r][b] = acc[r][b];
    }
    __syncthreads();
    if (threadIdx.x < FR_R * FR_NB) {
      const int r = threadIdx.x / FR_NB, bb = threadIdx.x - r * FR_NB;
      if (r < nrow && bb < nb) {
        double s = 0;
#pragma unroll
        for (int w = 0; w < FR_THREADS / 32; ++w) s += red[w][r][bb];
        zb[(size_t)bb * D + ib + r] = mu[ib + r] + s;
      }
    }
  }
}
// ok[0] = the gradient evaluation produced a finite log density and gradient [U: calc_grad's check_finite]
__global__ void __launch_bounds__(1024) k_fr_check(const double *__restrict__ g, int64_t D, const Ctl *__restrict__ ctl, int *__restrict__ ok) {
  __shared__ int bad;
  if (threadIdx.x == 0) bad = !isfinite(ctl->lp);
  __syncthreads();
  int b = 0;
  for (int64_t i = threadIdx.x; i < D; i += blockDim.x) b |= !isfinite(g[i]);
  if (b) bad = 1;
  __syncthreads();
  if (threadIdx.x == 0) { ok[0] = !bad; if (bad) ok[1] += 1; }
}
// One pass over (L, H): the step of iteration `it` (gradient g at the draw e0) and the transform of the next draw e1.
__global__ void __launch_bounds__(FR_THREADS) k_fr_update(double *__restrict__ L, double *__restrict__ H, int64_t D,
                                                          const double *__restrict__ g, const double *__restrict__ e0,
                                                          const double *__restrict__ e1, double *__restrict__ mu,
                                                          double *__restrict__ hmu, double *__restrict__ zeta,
                                                          const int *__restrict__ okp, int first, double es) {
  __shared__ double red[FR_THREADS / 32];
  __shared__ double s_dinv;
  const bool ok = okp[0] != 0;  // a failed evaluation contributes a zero gradient (adapt_eta's catch; SGA reports it)
  for (int64_t i = blockIdx.x; i < D; i += gridDim.x) {
    const size_t r0 = (size_t)i * (size_t)(i + 1) / 2;
    double *row = L + r0, *hrow = H + r0;
    if (threadIdx.x == 0) s_dinv = 1.0 / row[i];   // the entropy term uses L_ii BEFORE the update
    __syncthreads();
    const double gi = ok ? g[i] : 0.0, dinv = s_dinv;
    double acc = 0.0;
    // FR_U independent entries per thread and trip: enough loads in flight to cover the HBM latency (one entry per
    // trip ran at 2.9 TB/s); L and H stream past L2 (ld/st.global.cs) so that Y and the draws stay resident there
    for (int64_t jb = threadIdx.x; jb <= i; jb += FR_U * FR_THREADS) {
      double hv[FR_U], lv[FR_U], a0[FR_U], a1[FR_U];
#pragma unroll
      for (int u = 0; u < FR_U; ++u) {
        const int64_t j = jb + u * FR_THREADS;
        const bool in = j <= i;
        hv[u] = in ? __ldcs(hrow + j) : 0.0;
        lv[u] = in ? __ldcs(row + j) : 0.0;
        a0[u] = in ? e0[j] : 0.0;
        a1[u] = in ? e1[j] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < FR_U; ++u) {
        const int64_t j = jb + u * FR_THREADS;
        if (j <= i) {
          const double gl = ok ? gi * a0[u] + (j == i ? dinv : 0.0) : 0.0;
          const double h = first ? hv[u] + gl * gl : 0.9 * hv[u] + 0.1 * gl * gl;
          __stcs(hrow + j, h);
          const double l = lv[u] + es * gl / (1.0 + sqrt(h));
          __stcs(row + j, l);
          acc = fma(l, a1[u], acc);
        }
      }
    }
    acc = fr_block_sum(acc, red);
    if (threadIdx.x == 0) {
      double h = hmu[i];
      h = first ? h + gi * gi : 0.9 * h + 0.1 * gi * gi;
      hmu[i] = h;
      const double m = mu[i] + es * gi / (1.0 + sqrt(h));
      mu[i] = m;
      zeta[i] = m + acc;
    }
    __syncthreads();
  }
}
__global__ void k_fr_diag(const double *__restrict__ L, int64_t D, double *__restrict__ out, int log_abs) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < D; i += (int64_t)gridDim.x * blockDim.x) {
    const double d = L[(size_t)i * (size_t)(i + 1) / 2 + i];
    out[i] = log_abs ? (d != 0.0 ? log(fabs(d)) : 0.0) : d;
  }
}
// draws[g][t0 + b] = exp(zb[b][amp + g])
__global__ void k_fr_amp_draws(const double *__restrict__ zb, int64_t D, int64_t amp, int64_t G, int nd, int t0, int nb,
                               double *__restrict__ draws) {
  const int64_t n = G * nb;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t g = q / nb;
    const int b = (int)(q - g * nb);
    draws[g * nd + t0 + b] = exp(zb[(size_t)b * D + amp + g]);
  }
}

static int fr_grid(int64_t rows) { return (int)std::min<int64_t>(rows, 148 * 4); }
// persistent grid of the update pass: exactly the CTAs that are resident at once (rows are dealt round-robin, so a
// partial second wave -- 592 CTAs on 444 slots in the first version -- leaves two thirds of the machine idle at the end)
static int fr_grid_update(int64_t rows) {
  static int per_sm = 0, nsm = 148;
  if (!per_sm) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_fr_update, FR_THREADS, 0) != cudaSuccess || per_sm < 1) per_sm = 2;
  }
  return (int)std::min<int64_t>(rows, (int64_t)nsm * per_sm);
}
static int fr_grid_gemv(int64_t rows) { return (int)std::min<int64_t>((rows + FR_R - 1) / FR_R, 148 * 2); }

int fr_setup(bvg_fit *f) {
  if (f->fr) return BVG_OK;
  if (f->comm_mode || f->cfg.world_size > 1) { bvg_set_error("algorithm = 'fullrank' couples all parameters and cannot be gene-sharded"); return BVG_ERR_ARG; }
  const int64_t D = f->m.L.D;
  const size_t n = tri(D), bytes = sizeof(double) * (2 * n + 2 * (size_t)D + 2 * (size_t)FR_NB * D) + 64;
  size_t fr_free = 0, total = 0;
  CUDA_TRY(cudaMemGetInfo(&fr_free, &total));
  if (bytes > fr_free) {
    bvg_set_error("algorithm = 'fullrank' needs %.1f GB of device memory for the Cholesky factor and its step-size history (D = %lld); %.1f GB free",
                  bytes / 1e9, (long long)D, fr_free / 1e9);
    return BVG_ERR_ARG;
  }
  FrState *s = new FrState();
  memset(s, 0, sizeof(*s));
  if (cudaMalloc(&s->L, bytes) != cudaSuccess) { delete s; cudaGetLastError(); bvg_set_error("out of device memory (fullrank, %.1f GB)", bytes / 1e9); return BVG_ERR_CUDA; }
  s->H = s->L + n; s->eta[0] = s->H + n; s->eta[1] = s->eta[0] + D; s->etab = s->eta[1] + D; s->zb = s->etab + (size_t)FR_NB * D;
  s->ok = (int *)(s->zb + (size_t)FR_NB * D);
  s->n = n; s->it = 1; s->step = 0.1;
  f->fr = s;
  return BVG_OK;
}
void fr_destroy(bvg_fit *f) {
  FrState *s = (FrState *)f->fr;
  if (!s) return;
  cudaFree(s->L);
  delete s;
  f->fr = nullptr;
}
// mu <- d_init (NULL keeps it); L <- I when reset_L; histories and the failure counter <- 0
int fr_reset(bvg_fit *f, const double *d_init, bool reset_L) {
  BVG_TRY(fr_setup(f));
  FrState *s = (FrState *)f->fr;
  const int64_t D = f->m.L.D;
  if (d_init) CUDA_TRY(cudaMemcpyAsync(f->m.mu, d_init, sizeof(double) * D, cudaMemcpyDeviceToDevice, f->stream));
  CUDA_TRY(cudaMemsetAsync(f->m.hist, 0, sizeof(double) * D, f->stream));
  k_fr_init<<<fr_grid(D), 256, 0, f->stream>>>(s->L, s->H, D, reset_L ? 1 : 0);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemsetAsync(s->ok, 0, 2 * sizeof(int), f->stream));
  s->it = 1;
  f->st.kernel_launches += 1;
  return BVG_OK;
}
void fr_set_phase(bvg_fit *f, double eta) {
  FrState *s = (FrState *)f->fr;
  s->it = 1; s->step = eta;
}
int fr_bad(bvg_fit *f, int *bad) {
  FrState *s = (FrState *)f->fr;
  int h[2];
  CUDA_TRY(cudaMemcpyAsync(h, s->ok, sizeof(h), cudaMemcpyDeviceToHost, f->stream));
  CUDA_TRY(cudaStreamSynchronize(f->stream));
  *bad = h[1];
  return BVG_OK;
}
// n consecutive ELBO-gradient evaluations + steps.  The first one of a call transforms its own draw; afterwards the
// update kernel of iteration t produces zeta of iteration t + 1 in the same pass over L.
int fr_steps(bvg_fit *f, int n) {
  BVG_TRY(fr_setup(f));
  FrState *s = (FrState *)f->fr;
  const int64_t D = f->m.L.D;
  if (n <= 0) return BVG_OK;
  k_fr_normals<<<148, 256, 0, f->stream>>>(f->m.seed, STREAM_GRAD, f->h_t_grad, 1, D, s->eta[0]);
  k_fr_gemv<<<fr_grid_gemv(D), FR_THREADS, 0, f->stream>>>(s->L, D, 0, D, f->m.mu, s->eta[0], f->m.zeta, 1);
  CUDA_TRY(cudaGetLastError());
  f->st.kernel_launches += 2;
  s->cur = 0;
  for (int i = 0; i < n; ++i) {
    BVG_TRY(eval_log_prob_device(f, MODE_GRAD_OUT, 1, 1, true));
    k_fr_check<<<1, 1024, 0, f->stream>>>(f->m.grad, D, f->m.ctl, s->ok);
    f->h_t_grad += 1;
    k_fr_normals<<<148, 256, 0, f->stream>>>(f->m.seed, STREAM_GRAD, f->h_t_grad, 1, D, s->eta[s->cur ^ 1]);
    k_fr_update<<<fr_grid_update(D), FR_THREADS, 0, f->stream>>>(s->L, s->H, D, f->m.grad, s->eta[s->cur], s->eta[s->cur ^ 1], f->m.mu,
                                                         f->m.hist, f->m.zeta, s->ok, s->it == 1 ? 1 : 0, s->step / sqrt((double)s->it));
    CUDA_TRY(cudaGetLastError());
    f->st.kernel_launches += 3;
    s->cur ^= 1;
    s->it += 1;
  }
  f->st.grad_evals += n;
  return BVG_OK;
}
// lps[0..S): log_prob<propto = false, jacobian = true> at zeta_s = mu + L eta_s, eta_s = normal(stream, t0 + s, .)
static int fr_log_probs(bvg_fit *f, uint32_t stream, uint32_t t0, int S, std::vector<double> &lps, double *d_store /* optional [ncol][S] */,
                        int64_t store_n) {
  FrState *s = (FrState *)f->fr;
  const int64_t D = f->m.L.D;
  if (f->elbo_cap < S) {
    cudaFree(f->m.elbo_lps);
    f->m.elbo_lps = nullptr; f->elbo_cap = 0;
    CUDA_TRY(cudaMalloc(&f->m.elbo_lps, sizeof(double) * S));
    f->elbo_cap = S;
  }
  Ctl h;
  CUDA_TRY(cudaMemcpyAsync(&h, f->m.ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, f->stream));
  CUDA_TRY(cudaStreamSynchronize(f->stream));
  h.t_draw = 0;
  CUDA_TRY(cudaMemcpyAsync(f->m.ctl, &h, sizeof(Ctl), cudaMemcpyHostToDevice, f->stream));
  CUDA_TRY(cudaStreamSynchronize(f->stream));
  double *zeta0 = f->m.zeta;
  int rc = BVG_OK;
  for (int b0 = 0; b0 < S && rc == BVG_OK; b0 += FR_NB) {
    const int nb = std::min(FR_NB, S - b0);
    k_fr_normals<<<148 * 2, 256, 0, f->stream>>>(f->m.seed, stream, t0 + (uint32_t)b0, nb, D, s->etab);
    k_fr_gemv<<<fr_grid_gemv(D), FR_THREADS, 0, f->stream>>>(s->L, D, 0, D, f->m.mu, s->etab, s->zb, nb);
    f->st.kernel_launches += 2;
    for (int b = 0; b < nb && rc == BVG_OK; ++b) {
      f->m.zeta = s->zb + (size_t)b * D;   // the launchers copy the model descriptor by value
      rc = launch_prep(f, PREP_NONE);
      if (rc == BVG_OK) rc = launch_lik(f, false);
      if (rc == BVG_OK) rc = launch_gene_finish(f, MODE_POST_DRAW, 1, 0);
      if (rc == BVG_OK && d_store) rc = fr_store_draw(f, store_n, b0 + b, d_store, s->etab + (size_t)b * D);
    }
  }
  f->m.zeta = zeta0;
  BVG_TRY(rc);
  CUDA_TRY(cudaGetLastError());
  lps.resize(S);
  CUDA_TRY(cudaMemcpyAsync(lps.data(), f->m.elbo_lps, sizeof(double) * S, cudaMemcpyDeviceToHost, f->stream));
  CUDA_TRY(cudaStreamSynchronize(f->stream));
  return BVG_OK;
}
int fr_elbo(bvg_fit *f, int S, double *out) {
  BVG_TRY(fr_setup(f));
  FrState *s = (FrState *)f->fr;
  const int64_t D = f->m.L.D;
  Ctl h;
  CUDA_TRY(cudaMemcpyAsync(&h, f->m.ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, f->stream));
  CUDA_TRY(cudaStreamSynchronize(f->stream));
  const uint32_t t0 = h.t_elbo;
  std::vector<double> lps;
  BVG_TRY(fr_log_probs(f, STREAM_ELBO, t0, S, lps, nullptr, 0));
  CUDA_TRY(cudaMemcpyAsync(&h, f->m.ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, f->stream));
  CUDA_TRY(cudaStreamSynchronize(f->stream));
  h.t_elbo = t0 + (uint32_t)S;
  CUDA_TRY(cudaMemcpyAsync(f->m.ctl, &h, sizeof(Ctl), cudaMemcpyHostToDevice, f->stream));
  // entropy: D/2 (1 + log 2 pi) + sum log |L_ii|
  // (log |L_ii| is kept in the omega slot, unused by this family: bvg_fit_get_variational then reports it)
  k_fr_diag<<<148, 256, 0, f->stream>>>(s->L, D, f->m.omega, 1);
  BVG_TRY(vec_dot_sum(f->stream, f->m.omega, nullptr, f->scratch + D, D, f->m.tot));
  double hs[2];
  CUDA_TRY(cudaMemcpyAsync(hs, f->m.tot, sizeof(hs), cudaMemcpyDeviceToHost, f->stream));
  CUDA_TRY(cudaStreamSynchronize(f->stream));
  f->st.elbo_evals++;
  f->st.kernel_launches += 2;
  double acc = 0;
  int okn = 0;
  for (double v : lps) if (std::isfinite(v)) { acc += v; ++okn; }
  if (!okn) { *out = -INFINITY; return BVG_OK; }
  *out = acc / okn + 0.5 * (double)D * (1.0 + BVG_LOG_TWO_PI) + hs[1];
  if (std::isnan(*out)) *out = -INFINITY;
  return BVG_OK;
}
// n.draws draws of amplitude[g] = exp(mu_u + sum_j L[u][j] eta_j) into f->draws ([G][n_draws]); only rows
// amp .. amp + G of L are read
int fr_amplitude_draws(bvg_fit *f) {
  BVG_TRY(fr_setup(f));
  FrState *s = (FrState *)f->fr;
  const int64_t D = f->m.L.D, amp = f->m.L.amp, G = f->m.G;
  const int nd = f->cfg.n_draws;
  for (int t0 = 0; t0 < nd; t0 += FR_NB) {
    const int nb = std::min(FR_NB, nd - t0);
    k_fr_normals<<<148 * 2, 256, 0, f->stream>>>(f->m.seed, STREAM_DRAW, (uint32_t)t0, nb, D, s->etab);
    k_fr_gemv<<<fr_grid_gemv(G), FR_THREADS, 0, f->stream>>>(s->L, D, amp, amp + G, f->m.mu, s->etab, s->zb, nb);
    k_fr_amp_draws<<<148, 256, 0, f->stream>>>(s->zb, D, amp, G, nd, t0, nb, f->draws);
    f->st.kernel_launches += 3;
  }
  CUDA_TRY(cudaGetLastError());
  return BVG_OK;
}
// posterior draws of the whole model (bvg_fit_get_posterior): draw d = mu + L normal(stream 3, d, .)
int fr_posterior(bvg_fit *f, int n, double *d_out) {
  BVG_TRY(fr_setup(f));
  std::vector<double> lps;
  return fr_log_probs(f, STREAM_DRAW, 0, n, lps, d_out, n);
}
int fr_get_cholesky(bvg_fit *f, double *L_out) {
  BVG_TRY(fr_setup(f));
  FrState *s = (FrState *)f->fr;
  CUDA_TRY(cudaMemcpyAsync(L_out, s->L, sizeof(double) * s->n, cudaMemcpyDeviceToHost, f->stream));
  CUDA_TRY(cudaStreamSynchronize(f->stream));
  return BVG_OK;
}
int fr_set_cholesky(bvg_fit *f, const double *L_in) {
  BVG_TRY(fr_setup(f));
  FrState *s = (FrState *)f->fr;
  CUDA_TRY(cudaMemcpyAsync(s->L, L_in, sizeof(double) * s->n, cudaMemcpyHostToDevice, f->stream));
  CUDA_TRY(cudaStreamSynchronize(f->stream));
  return BVG_OK;
}
// omega[i] <- log |L_ii| (what bvg_fit_get_variational reports for a full-rank fit)
int fr_log_diag(bvg_fit *f) {
  BVG_TRY(fr_setup(f));
  FrState *s = (FrState *)f->fr;
  k_fr_diag<<<148, 256, 0, f->stream>>>(s->L, f->m.L.D, f->m.omega, 1);
  CUDA_TRY(cudaGetLastError());
  return BVG_OK;
}
