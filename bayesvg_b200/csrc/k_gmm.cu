// k_gmm.cu -- the step AFTER the SVG fit: clusterSVGsBayes() (R/clusterSVGsBayes.R:63-319), i.e. ADVI on inst/clusterSVGs.stan,
// a diagonal-covariance Gaussian mixture over the PCA embedding of the SVGs (N <= a few thousand genes, D = n.PCs = 30, K =
// n.clust = 5: P = K - 1 + 2 K D = 304 parameters; default algorithm "fullrank", 30,000 iterations, 300 ELBO draws every 100).
//
// The model is tiny; what costs the reference its time is 30,000 + 90,000 sequential log-density evaluations through an
// autodiff tape.  Here ONE CTA owns the whole problem and runs a batch of iterations per launch without returning to the host:
// draw -> transform (mean-field, or zeta = mu + L eta with a warp per row of the packed factor) -> responsibilities (thread per
// gene, X transposed so the reads coalesce) -> gradient sums (thread per (cluster, PC, segment of genes), fixed-order combine)
// -> reverse pass through the stick-breaking simplex transform -> Stan's step-size sequence.  ELBO estimates and the posterior
// pass (draws, soft assignments of :215-239, log-likelihood of :281-284) are loops inside one launch each as well.
// State lives in global memory (L2-resident: 0.4 MB at the defaults); nothing here is bandwidth- or tensor-bound -- the design
// goal is zero launches and zero host round trips inside the 100-iteration stretches between convergence checks.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <vector>

#include "bvg_internal.cuh"

#define GMM_THREADS 1024
#define GMM_MAXK 16

struct GmmDev {
  int64_t N;
  int D, K, P, fullrank;
  int r_smem;           // 1: the responsibilities of an evaluation fit in shared memory (N K doubles <= 96 KB), else they go through m.r
  const double *XT;     // [D][N]
  double *mu, *var, *hmu, *hvar, *zeta, *grad, *eta, *r;   // var: omega [P] or packed L; r: [N][K] responsibilities
  double *scal;         // [0] lp of the last evaluation, [1] ELBO sum, [2] ELBO finite count, [3] bad evaluations, [4] loglik sum
  uint64_t seed;
  double lpconst;       // the parameter-free terms of log_prob<propto = false>: normal / dirichlet normalising constants (host-computed:
                        // an lgamma evaluated by all 1,024 threads of the CTA cost 5 us of its one SM's fp64 pipe per evaluation)
};

__device__ inline double gmm_inv_logit(double a) {
  if (a >= 0) return 1.0 / (1.0 + exp(-a));
  const double e = exp(a);
  return e / (1.0 + e);
}
__device__ inline double gmm_log1p_exp(double a) { return a > 0 ? a + log1p(exp(-a)) : log1p(exp(a)); }

// shared-memory layout of one evaluation: (mu, 1/sigma)[K D] | theta[K] | zk[K] | stick[K] | lconst[K] | R[K] | mu[K D] | sigma[K D] | red[64] | part[...]
struct GmmSh {
  double *theta, *zk, *stick, *lconst, *R, *mu, *sigma, *red, *part;
  double2 *mi;   // (mu, 1 / sigma) per (cluster, PC)
  double *rs;    // responsibilities [N][K] (shared memory when they fit, else global)
};
__device__ inline GmmSh gmm_sh(double *base, int K, int D, int nseg_part = 0, double *r_global = nullptr, int r_smem = 0) {
  GmmSh s;
  s.mi = reinterpret_cast<double2 *>(base);   // first, 16-byte aligned
  base += 2 * K * D;
  s.theta = base; s.zk = s.theta + K; s.stick = s.zk + K; s.lconst = s.stick + K; s.R = s.lconst + K;
  s.mu = s.R + K; s.sigma = s.mu + K * D; s.red = s.sigma + K * D; s.part = s.red + 64;
  s.rs = r_smem ? s.part + nseg_part : r_global;
  return s;
}
__host__ __device__ inline int gmm_part_doubles(int K, int D) {
  const int KD = K * D, a = GMM_THREADS / KD, nseg = a < 1 ? 1 : (a > 32 ? 32 : a);
  return 2 * nseg * KD + nseg * K;
}
__device__ inline double gmm_block_sum(double v, double *red) {   // fixed tree: deterministic
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0;
  for (int w = 0; w < GMM_THREADS / 32; ++w) s += red[w];
  __syncthreads();
  return s;
}
// constrain zeta into shared memory; returns the log-Jacobian (every thread gets it)
__device__ inline double gmm_constrain(const GmmDev &m, const GmmSh &s, const double *zeta) {
  const int K = m.K, D = m.D, tid = threadIdx.x;
  if (tid == 0) {
    double lj = 0, st = 1.0;
    for (int q = 0; q < K - 1; ++q) {
      const double a = zeta[q] - log((double)(K - 1 - q)), z = gmm_inv_logit(a);
      s.zk[q] = z; s.stick[q] = st;
      s.theta[q] = st * z;
      lj += log(st) - gmm_log1p_exp(-a) - gmm_log1p_exp(a);
      st -= s.theta[q];
    }
    s.theta[K - 1] = st;
    s.red[63] = lj;
  }
  const double *um = zeta + K - 1, *us = um + K * D;
  for (int q = tid; q < K * D; q += GMM_THREADS) {
    const int j = q / D, d = q - j * D;
    s.mu[q] = um[j + K * d];          // matrix[K, D] is column-major in the unconstrained vector
    s.sigma[q] = exp(us[q]);
    s.mi[q] = make_double2(um[j + K * d], exp(-us[q]));
  }
  __syncthreads();
  double ls = 0;
  for (int q = tid; q < K * D; q += GMM_THREADS) ls += us[q];
  const double lj = s.red[63];
  ls = gmm_block_sum(ls, s.red);
  if (tid < K) {
    double sl = 0;
    for (int d = 0; d < D; ++d) sl += us[tid * D + d];   // log sigma is the unconstrained coordinate itself
    s.lconst[tid] = log(s.theta[tid]) - 0.5 * (D * BVG_LOG_TWO_PI + 2.0 * sl);
  }
  __syncthreads();
  return lj + ls;
}
// sum_i log_sum_exp_j(...) with the responsibilities left in m.r (and, if soft, accumulated into soft[i][j] with weight wsoft).
// KT = K rounded up to 4 / 8 / 16: the per-cluster accumulators are fully unrolled into registers, and (mu, 1 / sigma) of a
// (cluster, PC) pair come in ONE 16-byte shared-memory read.  ncu of the first version (runtime K: accumulators in local
// memory, two 8-byte reads per term): 116 k warp instructions per evaluation, mio_throttle + short_scoreboard on top.
template <int KT>
__device__ inline double gmm_mixture_t(const GmmDev &m, const GmmSh &s, bool store_r, double *soft, double wsoft) {
  const int K = m.K, D = m.D;
  double acc = 0;
  for (int64_t i = threadIdx.x; i < m.N; i += GMM_THREADS) {
    double lc[KT];
#pragma unroll
    for (int j = 0; j < KT; ++j) lc[j] = 0.0;
    for (int d = 0; d < D; ++d) {
      const double x = m.XT[(size_t)d * m.N + i];   // coalesced over the genes of a warp; read once for all clusters
#pragma unroll
      for (int j = 0; j < KT; ++j)
        if (j < K) { const double2 v = s.mi[j * D + d]; const double u = (x - v.x) * v.y; lc[j] = fma(u, u, lc[j]); }
    }
    double mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < KT; ++j)
      if (j < K) { lc[j] = s.lconst[j] - 0.5 * lc[j]; mx = lc[j] > mx ? lc[j] : mx; }
    double se = 0;
#pragma unroll
    for (int j = 0; j < KT; ++j)
      if (j < K) { lc[j] = exp(lc[j] - mx); se += lc[j]; }
    acc += mx + log(se);
    const double ise = 1.0 / se;
#pragma unroll
    for (int j = 0; j < KT; ++j)
      if (j < K) {
        const double r = lc[j] * ise;
        if (store_r) s.rs[(size_t)i * K + j] = r;
        if (soft) soft[(size_t)i * K + j] += wsoft * r;
      }
  }
  return gmm_block_sum(acc, s.red);
}
__device__ inline double gmm_mixture(const GmmDev &m, const GmmSh &s, bool store_r, double *soft, double wsoft) {
  if (m.K <= 4) return gmm_mixture_t<4>(m, s, store_r, soft, wsoft);
  if (m.K <= 8) return gmm_mixture_t<8>(m, s, store_r, soft, wsoft);
  return gmm_mixture_t<16>(m, s, store_r, soft, wsoft);
}
// one evaluation of log_prob<propto, jacobian> (+ gradient into m.grad) at zeta
__device__ inline double gmm_eval(const GmmDev &m, const GmmSh &s, const double *zeta, int jac, int propto, bool want_grad) {
  const int K = m.K, D = m.D, KD = K * D, tid = threadIdx.x;
  const double lj = gmm_constrain(m, s, zeta);
  double pr = 0;
  for (int q = tid; q < KD; q += GMM_THREADS) pr += -0.5 * s.mu[q] * s.mu[q] / 49.0 - 0.5 * s.sigma[q] * s.sigma[q] / 9.0;
  if (tid < K) pr += log(s.theta[tid]);
  pr = gmm_block_sum(pr, s.red);
  double lp = pr + gmm_mixture(m, s, want_grad, nullptr, 0.0);
  if (!propto) lp += m.lpconst;
  if (jac) lp += lj;
  if (!want_grad) return lp;
  // gradient sums over the genes: segments of the gene range per (j, d) pair, combined in a fixed order
  const int nseg = max(1, min(GMM_THREADS / KD, 32));
  const int64_t per = (m.N + nseg - 1) / nseg;
  for (int w = tid; w < KD * nseg; w += GMM_THREADS) {
    const int q = w % KD, sg = w / KD, j = q / D, d = q - j * D;
    const int64_t i0 = (int64_t)sg * per, i1 = i0 + per < m.N ? i0 + per : m.N;
    const double muq = s.mu[q];
    double a = 0, b = 0, rs = 0;
#pragma unroll 4
    for (int64_t i = i0; i < i1; ++i) {
      const double r = s.rs[(size_t)i * K + j], df = m.XT[(size_t)d * m.N + i] - muq;
      a = fma(r, df, a);
      b = fma(r * df, df, b);
      rs += r;
    }
    s.part[(size_t)(0 * nseg + sg) * KD + q] = a;
    s.part[(size_t)(1 * nseg + sg) * KD + q] = b;
    if (d == 0) s.part[(size_t)2 * nseg * KD + sg * K + j] = rs;
  }
  __syncthreads();
  const double j1 = jac ? 1.0 : 0.0;
  double *gm = m.grad + K - 1, *gs = gm + KD;
  for (int q = tid; q < KD; q += GMM_THREADS) {
    const int j = q / D, d = q - j * D;
    double a = 0, b = 0, Rj = 0;
    for (int sg = 0; sg < nseg; ++sg) {
      a += s.part[(size_t)(0 * nseg + sg) * KD + q];
      b += s.part[(size_t)(1 * nseg + sg) * KD + q];
      Rj += s.part[(size_t)2 * nseg * KD + sg * K + j];
    }
    const double sgm = s.sigma[q];
    gm[j + K * d] = a / (sgm * sgm) - s.mu[q] / 49.0;
    gs[q] = (b / (sgm * sgm * sgm) - Rj / sgm - sgm / 9.0) * sgm + j1;
    if (d == 0) s.R[j] = Rj;
  }
  __syncthreads();
  if (tid == 0) {   // reverse pass through the stick-breaking transform
    double gx[GMM_MAXK];
    for (int j = 0; j < K; ++j) gx[j] = (s.R[j] + 1.0) / s.theta[j];
    double gst = gx[K - 1];
    for (int q = K - 2; q >= 0; --q) {
      const double z = s.zk[q], st = s.stick[q];
      const double gz = gx[q] * st - gst * st + j1 * (1.0 / z - 1.0 / (1.0 - z));
      gst = gx[q] * z + gst * (1.0 - z) + j1 / st;
      m.grad[q] = gz * z * (1.0 - z);
    }
  }
  __syncthreads();
  return lp;
}
// zeta = mu + exp(omega) eta (mean-field) or mu + L eta (full-rank, warp per row)
__device__ inline void gmm_transform(const GmmDev &m, uint32_t stream, uint32_t t) {
  const int P = m.P, tid = threadIdx.x;
  for (int i = tid; i < P; i += GMM_THREADS) m.eta[i] = bvg_normal(m.seed, stream, t, (uint64_t)i);
  __syncthreads();
  if (!m.fullrank) {
    for (int i = tid; i < P; i += GMM_THREADS) m.zeta[i] = m.mu[i] + exp(m.var[i]) * m.eta[i];
  } else {
    const int lane = tid & 31, warp = tid >> 5;
    for (int i = warp; i < P; i += GMM_THREADS / 32) {
      const double *row = m.var + (size_t)i * (i + 1) / 2;
      double a = 0;
      for (int j = lane; j <= i; j += 32) a = fma(row[j], m.eta[j], a);
      a = warp_sum(a);
      if (lane == 0) m.zeta[i] = m.mu[i] + a;
    }
  }
  __syncthreads();
}
extern __shared__ __align__(16) double gmm_smem[];

__global__ void __launch_bounds__(GMM_THREADS) k_gmm_logprob(GmmDev m, int jac, int propto, int want_grad) {
  const GmmSh s = gmm_sh(gmm_smem, m.K, m.D, gmm_part_doubles(m.K, m.D), m.r, m.r_smem);
  const double lp = gmm_eval(m, s, m.zeta, jac, propto, want_grad != 0);
  if (threadIdx.x == 0) m.scal[0] = lp;
}
// n iterations of stochastic gradient ascent: iteration counter it0.., gradient draws t0..
__global__ void __launch_bounds__(GMM_THREADS) k_gmm_steps(GmmDev m, int n, int it0, uint32_t t0, double eta_s) {
  const GmmSh s = gmm_sh(gmm_smem, m.K, m.D, gmm_part_doubles(m.K, m.D), m.r, m.r_smem);
  const int P = m.P, tid = threadIdx.x;
  for (int q = 0; q < n; ++q) {
    const int it = it0 + q;
    gmm_transform(m, STREAM_GRAD, t0 + (uint32_t)q);
    const double lp = gmm_eval(m, s, m.zeta, 1, 1, true);
    int bad = !isfinite(lp);
    for (int i = tid; i < P; i += GMM_THREADS) bad |= !isfinite(m.grad[i]);
    bad = __syncthreads_or(bad);
    if (bad && tid == 0) m.scal[3] += 1.0;
    const double es = eta_s / sqrt((double)it);
    const bool first = it == 1;
    if (!m.fullrank) {
      for (int i = tid; i < P; i += GMM_THREADS) {
        const double g = bad ? 0.0 : m.grad[i];
        const double go = bad ? 0.0 : g * (m.zeta[i] - m.mu[i]) + 1.0;
        double hm = m.hmu[i], ho = m.hvar[i];
        hm = first ? hm + g * g : 0.9 * hm + 0.1 * g * g;
        ho = first ? ho + go * go : 0.9 * ho + 0.1 * go * go;
        m.hmu[i] = hm; m.hvar[i] = ho;
        m.mu[i] += es * g / (1.0 + sqrt(hm));
        m.var[i] += es * go / (1.0 + sqrt(ho));
      }
    } else {
      const int lane = tid & 31, warp = tid >> 5;
      for (int i = warp; i < P; i += GMM_THREADS / 32) {
        double *row = m.var + (size_t)i * (i + 1) / 2, *hrow = m.hvar + (size_t)i * (i + 1) / 2;
        const double g = bad ? 0.0 : m.grad[i], dinv = bad ? 0.0 : 1.0 / row[i];   // L_ii before the update (read by every lane first)
        __syncwarp();
        for (int j = lane; j <= i; j += 32) {
          const double gl = bad ? 0.0 : g * m.eta[j] + (j == i ? dinv : 0.0);
          double h = hrow[j];
          h = first ? h + gl * gl : 0.9 * h + 0.1 * gl * gl;
          hrow[j] = h;
          row[j] += es * gl / (1.0 + sqrt(h));
        }
        if (lane == 0) {
          double h = m.hmu[i];
          h = first ? h + g * g : 0.9 * h + 0.1 * g * g;
          m.hmu[i] = h;
          m.mu[i] += es * g / (1.0 + sqrt(h));
        }
      }
    }
    __syncthreads();
  }
}
// S draws of an ELBO estimate, spread over the CTAs of the grid (the draws are independent: mu and the scale do not change):
// CTA b evaluates draws [b per, (b + 1) per) with its own zeta / eta scratch and leaves (sum of finite log_prob<propto = false,
// jacobian = true>, count) in part[b]; the host adds the partials in CTA order.  Three quarters of all evaluations of a default fit
// are ELBO draws (300 every 100 iterations): one CTA for all of them took 5 of the 7 seconds.
__global__ void __launch_bounds__(GMM_THREADS) k_gmm_elbo(GmmDev m, int S, uint32_t t0, double *__restrict__ zs, double *__restrict__ part) {
  const GmmSh s = gmm_sh(gmm_smem, m.K, m.D);
  GmmDev mm = m;
  mm.zeta = zs + (size_t)blockIdx.x * 2 * m.P;
  mm.eta = mm.zeta + m.P;
  const int per = (S + gridDim.x - 1) / gridDim.x, q0 = blockIdx.x * per, q1 = q0 + per < S ? q0 + per : S;
  double acc = 0, ok = 0;
  for (int q = q0; q < q1; ++q) {
    gmm_transform(mm, STREAM_ELBO, t0 + (uint32_t)q);
    const double lp = gmm_eval(mm, s, mm.zeta, 1, 0, false);
    if (isfinite(lp)) { acc += lp; ok += 1.0; }
  }
  if (threadIdx.x == 0) { part[2 * blockIdx.x] = acc; part[2 * blockIdx.x + 1] = ok; }
  if (blockIdx.x == 0) {   // entropy: P/2 (1 + log 2 pi) + sum omega  |  sum log |L_ii|
    double e = 0;
    for (int i = threadIdx.x; i < m.P; i += GMM_THREADS) {
      if (!m.fullrank) e += m.var[i];
      else { const double d = m.var[(size_t)i * (i + 1) / 2 + i]; e += d != 0.0 ? log(fabs(d)) : 0.0; }
    }
    e = gmm_block_sum(e, s.red);
    if (threadIdx.x == 0) m.scal[5] = e + 0.5 * m.P * (1.0 + BVG_LOG_TWO_PI);
  }
}
// posterior pass (R/clusterSVGsBayes.R:201-291): nd draws (Philox stream 3) -> constrained draws, soft assignments, log-likelihood
__global__ void __launch_bounds__(GMM_THREADS) k_gmm_post(GmmDev m, int nd, double *__restrict__ draws, double *__restrict__ soft) {
  const GmmSh s = gmm_sh(gmm_smem, m.K, m.D);
  const int K = m.K, D = m.D, KD = K * D, tid = threadIdx.x;
  for (int64_t q = tid; q < m.N * K; q += GMM_THREADS) soft[q] = 0.0;
  __syncthreads();
  double ll = 0;
  for (int t = 0; t < nd; ++t) {
    gmm_transform(m, STREAM_DRAW, (uint32_t)t);
    gmm_constrain(m, s, m.zeta);
    if (draws) {
      double *o = draws + (size_t)t * (K + 2 * KD);
      for (int q = tid; q < K; q += GMM_THREADS) o[q] = s.theta[q];
      for (int q = tid; q < KD; q += GMM_THREADS) { const int j = q / D, d = q - j * D; o[K + j + K * d] = s.mu[q]; o[K + KD + q] = s.sigma[q]; }
    }
    ll += gmm_mixture(m, s, false, soft, 1.0 / nd);
  }
  if (tid == 0) m.scal[4] = ll / nd;
}
__global__ void k_gmm_reset(GmmDev m, const double *init, int reset_var) {
  const int P = m.P;
  const size_t nv = m.fullrank ? (size_t)P * (P + 1) / 2 : (size_t)P;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (size_t)gridDim.x * blockDim.x) {
    m.hvar[i] = 0.0;
    if (reset_var) m.var[i] = 0.0;
    if (i < (size_t)P) { m.hmu[i] = 0.0; if (init) m.mu[i] = init[i]; }
  }
}
__global__ void k_gmm_eye(GmmDev m) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m.P; i += gridDim.x * blockDim.x) m.var[(size_t)i * (i + 1) / 2 + i] = 1.0;
}

// B points in one launch, one CTA each (the ELBO draws of a Pathfinder iterate; B = 1 with a gradient for its L-BFGS):
// CTA b evaluates thetas[b]; gradients need the per-gene responsibilities, in shared memory when they fit, else in slice b of r.
__global__ void __launch_bounds__(GMM_THREADS) k_gmm_logprob_batch(GmmDev m, const double *__restrict__ thetas, double *__restrict__ lps,
                                                                   double *__restrict__ grads, double *__restrict__ rbatch, int jac, int propto) {
  GmmDev mb = m;
  const int b = blockIdx.x;
  if (grads) mb.grad = grads + (size_t)b * m.P;
  if (rbatch) mb.r = rbatch + (size_t)b * m.N * m.K;
  const GmmSh s = gmm_sh(gmm_smem, m.K, m.D, gmm_part_doubles(m.K, m.D), mb.r, m.r_smem);
  const double lp = gmm_eval(mb, s, thetas + (size_t)b * m.P, jac, propto, grads != nullptr);
  if (threadIdx.x == 0) lps[b] = lp;
}
// the posterior pass of k_gmm_post for GIVEN unconstrained draws (Pathfinder's PSIS-resampled draws instead of draws of a
// variational family): constrained draws, soft assignments (R/clusterSVGsBayes.R:215-239), log-likelihood (:281-284)
__global__ void __launch_bounds__(GMM_THREADS) k_gmm_post_given(GmmDev m, int nd, const double *__restrict__ thetas, double *__restrict__ draws,
                                                                double *__restrict__ soft) {
  const GmmSh s = gmm_sh(gmm_smem, m.K, m.D);
  const int K = m.K, D = m.D, KD = K * D, tid = threadIdx.x;
  for (int64_t q = tid; q < m.N * K; q += GMM_THREADS) soft[q] = 0.0;
  __syncthreads();
  double ll = 0;
  for (int t = 0; t < nd; ++t) {
    gmm_constrain(m, s, thetas + (size_t)t * m.P);
    if (draws) {
      double *o = draws + (size_t)t * (K + 2 * KD);
      for (int q = tid; q < K; q += GMM_THREADS) o[q] = s.theta[q];
      for (int q = tid; q < KD; q += GMM_THREADS) { const int j = q / D, d = q - j * D; o[K + j + K * d] = s.mu[q]; o[K + KD + q] = s.sigma[q]; }
    }
    ll += gmm_mixture(m, s, false, soft, 1.0 / nd);
    __syncthreads();
  }
  if (tid == 0) m.scal[4] = ll / nd;
}

namespace {
struct GmmHost {
  GmmDev m;
  char *base = nullptr;
  cudaStream_t st = nullptr;
  size_t smem = 0;
  double *d_init = nullptr, *d_zs = nullptr, *d_part = nullptr;
  int elbo_grid = 1;
  // handle form (bvg_gmm_open): batch buffers of bvg_gmm_eval, grown on demand
  double *d_batch = nullptr;   // [cap][P] points | [cap] lp | [gcap][P] gradients | [gcap][N K] responsibilities (when not in shared memory)
  int cap = 0, gcap = 0;
  int device = 0;
};
int gmm_setup(GmmHost &h, int device, const double *X, int64_t N, int D, int K, int fullrank, uint64_t seed) {
  if (!X || N < 1 || D < 1 || D > 256 || K < 2 || K > GMM_MAXK) { bvg_set_error("gmm: need N >= 1, 1 <= D <= 256, 2 <= K <= %d (Please provide a valid number of clusters.)", GMM_MAXK); return BVG_ERR_ARG; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); bvg_set_error("no CUDA device: bayesvg_b200 has no CPU fallback"); return BVG_ERR_CUDA; }
  CUDA_TRY(cudaSetDevice(device));
  const int P = K - 1 + 2 * K * D, KD = K * D;
  const size_t nv = fullrank ? (size_t)P * (P + 1) / 2 : (size_t)P;
  const int EG = 148;   // CTAs of an ELBO estimate
  const size_t nd = (size_t)D * N + 5 * (size_t)P + 2 * nv + (size_t)N * K + 16 + (size_t)P + (size_t)EG * 2 * P + 2 * EG;
  CUDA_TRY(cudaMalloc(&h.base, sizeof(double) * nd));
  CUDA_TRY(cudaMemset(h.base, 0, sizeof(double) * nd));
  CUDA_TRY(cudaStreamCreate(&h.st));
  double *p = (double *)h.base;
  GmmDev &m = h.m;
  m.N = N; m.D = D; m.K = K; m.P = P; m.fullrank = fullrank; m.seed = seed;
  m.lpconst = KD * (-0.5 * BVG_LOG_TWO_PI - std::log(7.0)) + KD * (-0.5 * BVG_LOG_TWO_PI - std::log(3.0)) + std::lgamma(2.0 * K);
  double *XT = p; p += (size_t)D * N;
  m.XT = XT;
  m.mu = p; p += P; m.hmu = p; p += P; m.zeta = p; p += P; m.grad = p; p += P; m.eta = p; p += P;
  m.var = p; p += nv; m.hvar = p; p += nv; m.r = p; p += (size_t)N * K; m.scal = p; p += 16; h.d_init = p; p += P; h.d_zs = p; p += (size_t)EG * 2 * P; h.d_part = p;
  h.elbo_grid = EG;
  // X is N x D column-major (an R matrix) = [D][N]: already the transposed layout the kernels want
  CUDA_TRY(cudaMemcpyAsync(XT, X, sizeof(double) * (size_t)D * N, cudaMemcpyHostToDevice, h.st));
  const int nseg = std::max(1, std::min(GMM_THREADS / KD, 32));
  h.smem = sizeof(double) * (5 * (size_t)K + 4 * (size_t)KD + 64 + (size_t)2 * nseg * KD + (size_t)nseg * K);
  m.r_smem = sizeof(double) * (size_t)N * K <= 96 * 1024 ? 1 : 0;   // then the gradient sums read them at shared-memory latency
  if (m.r_smem) h.smem += sizeof(double) * (size_t)N * K;
  if (h.smem > 200 * 1024) { bvg_set_error("gmm: K * D too large for one CTA's shared memory"); return BVG_ERR_ARG; }
  CUDA_TRY(cudaFuncSetAttribute(k_gmm_logprob, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h.smem));
  CUDA_TRY(cudaFuncSetAttribute(k_gmm_steps, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h.smem));
  CUDA_TRY(cudaFuncSetAttribute(k_gmm_elbo, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h.smem));
  CUDA_TRY(cudaFuncSetAttribute(k_gmm_post, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h.smem));
  CUDA_TRY(cudaFuncSetAttribute(k_gmm_logprob_batch, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h.smem));
  CUDA_TRY(cudaFuncSetAttribute(k_gmm_post_given, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h.smem));
  h.device = device;
  return BVG_OK;
}
void gmm_free(GmmHost &h) {
  if (h.st) cudaStreamDestroy(h.st);
  cudaFree(h.base);
  cudaFree(h.d_batch);
}
int gmm_reset(GmmHost &h) {   // mu <- init, variational scale <- 0 (mean-field) / I (full-rank), histories <- 0
  CUDA_TRY(cudaMemsetAsync(h.m.scal + 3, 0, sizeof(double), h.st));   // failed-evaluation counter of the phase
  k_gmm_reset<<<148, 256, 0, h.st>>>(h.m, h.d_init, 1);
  if (h.m.fullrank) k_gmm_eye<<<8, 256, 0, h.st>>>(h.m);
  CUDA_TRY(cudaGetLastError());
  return BVG_OK;
}
int gmm_elbo(GmmHost &h, int S, uint32_t &te, double *out) {
  const int grid = std::min(S, h.elbo_grid);
  k_gmm_elbo<<<grid, GMM_THREADS, h.smem, h.st>>>(h.m, S, te, h.d_zs, h.d_part);
  te += (uint32_t)S;
  double sc[6], part[2 * 148];
  CUDA_TRY(cudaMemcpyAsync(sc, h.m.scal, sizeof(sc), cudaMemcpyDeviceToHost, h.st));
  CUDA_TRY(cudaMemcpyAsync(part, h.d_part, sizeof(double) * 2 * grid, cudaMemcpyDeviceToHost, h.st));
  CUDA_TRY(cudaStreamSynchronize(h.st));
  double acc = 0, okn = 0;
  for (int b = 0; b < grid; ++b) { acc += part[2 * b]; okn += part[2 * b + 1]; }   // CTA order: deterministic
  *out = okn > 0 ? acc / okn + sc[5] : -INFINITY;
  if (std::isnan(*out)) *out = -INFINITY;
  return BVG_OK;
}
double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
}  // namespace

// log_prob<propto, jacobian> and its gradient at an unconstrained vector (the parity hook for clusterSVGs.stan)
extern "C" int bvg_gmm_log_prob(int device, const double *X, int64_t N, int D, int K, const double *theta, int jacobian, int propto,
                                double *lp, double *grad) {
  if (!theta || !lp) { bvg_set_error("null argument"); return BVG_ERR_ARG; }
  GmmHost h;
  int rc = gmm_setup(h, device, X, N, D, K, 0, 0);
  if (rc == BVG_OK) {
    if (cudaMemcpyAsync(h.m.zeta, theta, sizeof(double) * h.m.P, cudaMemcpyHostToDevice, h.st) != cudaSuccess) rc = BVG_ERR_CUDA;
    k_gmm_logprob<<<1, GMM_THREADS, h.smem, h.st>>>(h.m, jacobian, propto, grad ? 1 : 0);
    if (rc == BVG_OK && (cudaMemcpyAsync(lp, h.m.scal, sizeof(double), cudaMemcpyDeviceToHost, h.st) != cudaSuccess ||
                         (grad && cudaMemcpyAsync(grad, h.m.grad, sizeof(double) * h.m.P, cudaMemcpyDeviceToHost, h.st) != cudaSuccess) ||
                         cudaStreamSynchronize(h.st) != cudaSuccess)) rc = BVG_ERR_CUDA;
    if (rc == BVG_ERR_CUDA) bvg_set_error("gmm: %s", cudaGetErrorString(cudaGetLastError()));
  }
  gmm_free(h);
  return rc;
}

// The whole clusterSVGsBayes() fit (:153-160 with algorithm meanfield / fullrank, init = 0) + the posterior pass (:201-291).
//   opts[0] n_iter (30000), [1] n_draws (1000), [2] elbo_samples (300), [3] eval_elbo (100), [4] adapt_iter (50), [5] eta (<= 0: adapt),
//   [6] tol_rel_obj (0.01)
//   mu_out [P], var_out [P] (omega) or [P (P + 1) / 2] (packed L), draws_out [n_draws][K + 2 K D] (theta | mu column-major | sigma)
//   or NULL, soft_out [N][K] row-major, stats_out[8] = {eta, iterations, converged, last ELBO, log-likelihood, seconds, grad evals, elbo evals}
extern "C" int bvg_gmm_fit(int device, const double *X, int64_t N, int D, int K, int fullrank, uint64_t seed, const double *opts,
                           double *mu_out, double *var_out, double *draws_out, double *soft_out, double *stats_out) {
  if (!opts || !soft_out || !stats_out) { bvg_set_error("null argument"); return BVG_ERR_ARG; }
  const int n_iter = (int)opts[0], n_draws = (int)opts[1], S = (int)opts[2], eval_elbo = (int)opts[3], adapt_iter = (int)opts[4];
  const double tol = opts[6];
  if (n_iter < 1 || n_draws < 1 || S < 1 || eval_elbo < 1 || adapt_iter < 1) { bvg_set_error("gmm: iteration / sample counts must be positive"); return BVG_ERR_ARG; }
  GmmHost h;
  const double t_start = now_s();
  int rc = gmm_setup(h, device, X, N, D, K, fullrank ? 1 : 0, seed);
  double *d_draws = nullptr, *d_soft = nullptr;
  auto fail = [&](int code) { cudaFree(d_draws); cudaFree(d_soft); gmm_free(h); return code; };
  if (rc != BVG_OK) return fail(rc);
  const int P = h.m.P;
  const size_t nv = fullrank ? (size_t)P * (P + 1) / 2 : (size_t)P;
  uint32_t tg = 0, te = 0;
  int gevals = 0, eevals = 0;
  double eta = opts[5];
  if (eta <= 0) {   // adapt_eta [U: stan::variational::advi::adapt_eta]
    static const double seq[5] = {100, 10, 1, 0.1, 0.01};
    if ((rc = gmm_reset(h)) != BVG_OK) return fail(rc);
    double e0;
    if ((rc = gmm_elbo(h, S, te, &e0)) != BVG_OK) return fail(rc);
    ++eevals;
    if (!std::isfinite(e0)) { bvg_set_error("ADVI: the ELBO is not finite at the initial point"); return fail(BVG_ERR_NUMERIC); }
    double ebest = -INFINITY, eta_best = 0;
    for (int q = 0;; ++q) {
      if ((rc = gmm_reset(h)) != BVG_OK) return fail(rc);
      k_gmm_steps<<<1, GMM_THREADS, h.smem, h.st>>>(h.m, adapt_iter, 1, tg, seq[q]);
      tg += (uint32_t)adapt_iter; gevals += adapt_iter;
      double e;
      if ((rc = gmm_elbo(h, S, te, &e)) != BVG_OK) return fail(rc);
      ++eevals;
      if (e < ebest && ebest > e0) break;
      if (q < 4) { ebest = e; eta_best = seq[q]; }
      else {
        if (e > e0) { ebest = e; eta_best = seq[q]; break; }
        bvg_set_error("ADVI: all proposed step-sizes failed; the model may be ill-conditioned");
        return fail(BVG_ERR_NUMERIC);
      }
    }
    eta = eta_best;
  }
  if ((rc = gmm_reset(h)) != BVG_OK) return fail(rc);
  const int cbn = (int)std::max(0.1 * n_iter / eval_elbo, 2.0);
  std::vector<double> cb(cbn, 0.0), tmp(cbn);
  int cbc = 0, cbh = 0, conv = 0, it = 0;
  double elbo = 0, elbo_prev;
  while (it < n_iter) {
    const int n = std::min(eval_elbo - it % eval_elbo, n_iter - it);
    k_gmm_steps<<<1, GMM_THREADS, h.smem, h.st>>>(h.m, n, it + 1, tg, eta);
    tg += (uint32_t)n; gevals += n;
    it += n;
    if (it % eval_elbo == 0) {
      elbo_prev = elbo;
      if ((rc = gmm_elbo(h, S, te, &elbo)) != BVG_OK) return fail(rc);
      ++eevals;
      double sc[4];
      if (cudaMemcpy(sc, h.m.scal, sizeof(sc), cudaMemcpyDeviceToHost) != cudaSuccess) { bvg_set_error("gmm: readback failed"); return fail(BVG_ERR_CUDA); }
      if (sc[3] > 0) { bvg_set_error("ADVI: non-finite gradient during stochastic gradient ascent (%d evaluations)", (int)sc[3]); return fail(BVG_ERR_NUMERIC); }
      const double d = std::fabs((elbo - elbo_prev) / elbo);
      cb[cbh] = d; cbh = (cbh + 1) % cbn; if (cbc < cbn) ++cbc;
      double mean = 0;
      for (int i = 0; i < cbc; ++i) { mean += cb[i]; tmp[i] = cb[i]; }
      mean /= cbc;
      std::sort(tmp.begin(), tmp.begin() + cbc);
      const double med = tmp[cbc / 2];
      if (mean < tol) conv |= 1;
      if (med < tol) conv |= 2;
      if (conv) break;
    }
  }
  // posterior pass
  if (cudaMalloc(&d_soft, sizeof(double) * (size_t)N * K) != cudaSuccess || (draws_out && cudaMalloc(&d_draws, sizeof(double) * (size_t)n_draws * (K + 2 * K * D)) != cudaSuccess)) {
    cudaGetLastError(); bvg_set_error("out of device memory"); return fail(BVG_ERR_CUDA);
  }
  k_gmm_post<<<1, GMM_THREADS, h.smem, h.st>>>(h.m, n_draws, d_draws, d_soft);
  double sc[6];
  bool ok = cudaMemcpyAsync(sc, h.m.scal, sizeof(sc), cudaMemcpyDeviceToHost, h.st) == cudaSuccess &&
            cudaMemcpyAsync(soft_out, d_soft, sizeof(double) * (size_t)N * K, cudaMemcpyDeviceToHost, h.st) == cudaSuccess &&
            (!mu_out || cudaMemcpyAsync(mu_out, h.m.mu, sizeof(double) * P, cudaMemcpyDeviceToHost, h.st) == cudaSuccess) &&
            (!var_out || cudaMemcpyAsync(var_out, h.m.var, sizeof(double) * nv, cudaMemcpyDeviceToHost, h.st) == cudaSuccess) &&
            (!draws_out || cudaMemcpyAsync(draws_out, d_draws, sizeof(double) * (size_t)n_draws * (K + 2 * K * D), cudaMemcpyDeviceToHost, h.st) == cudaSuccess) &&
            cudaStreamSynchronize(h.st) == cudaSuccess && cudaGetLastError() == cudaSuccess;
  if (!ok) { bvg_set_error("gmm: %s", cudaGetErrorString(cudaGetLastError())); return fail(BVG_ERR_CUDA); }
  stats_out[0] = eta; stats_out[1] = it; stats_out[2] = conv; stats_out[3] = elbo; stats_out[4] = sc[4]; stats_out[5] = now_s() - t_start;
  stats_out[6] = gevals; stats_out[7] = eevals;
  return fail(BVG_OK);
}

// ---- handle form: the embedding stays on the device between evaluations (Pathfinder: R/clusterSVGsBayes.R:162-170) ----
extern "C" int bvg_gmm_open(int device, const double *X, int64_t N, int D, int K, void **handle) {
  if (!handle) { bvg_set_error("null argument"); return BVG_ERR_ARG; }
  GmmHost *h = new GmmHost();
  const int rc = gmm_setup(*h, device, X, N, D, K, 0, 0);
  if (rc != BVG_OK || cudaStreamSynchronize(h->st) != cudaSuccess) { gmm_free(*h); delete h; *handle = nullptr; return rc != BVG_OK ? rc : BVG_ERR_CUDA; }
  *handle = h;
  return BVG_OK;
}
extern "C" void bvg_gmm_close(void *handle) {
  if (!handle) return;
  GmmHost *h = (GmmHost *)handle;
  cudaSetDevice(h->device);
  gmm_free(*h);
  delete h;
}
// lp_out[b] = log_prob<propto, jacobian>(thetas[b]) for b < B; grad_out ([B][P]) may be NULL
extern "C" int bvg_gmm_eval(void *handle, const double *thetas, int B, int jacobian, int propto, double *lp_out, double *grad_out) {
  if (!handle || !thetas || !lp_out || B < 1) { bvg_set_error("null or empty argument"); return BVG_ERR_ARG; }
  GmmHost &h = *(GmmHost *)handle;
  CUDA_TRY(cudaSetDevice(h.device));
  const int P = h.m.P;
  const size_t nk = (size_t)h.m.N * h.m.K;
  const int gneed = grad_out ? B : 0;
  if (B > h.cap || gneed > h.gcap) {
    const int cap = std::max(B, h.cap), gcap = std::max(gneed, h.gcap);
    cudaFree(h.d_batch);
    h.d_batch = nullptr; h.cap = h.gcap = 0;
    const size_t nd = (size_t)cap * (P + 1) + (size_t)gcap * P + (h.m.r_smem ? 0 : (size_t)gcap * nk);
    CUDA_TRY(cudaMalloc(&h.d_batch, sizeof(double) * std::max<size_t>(nd, 1)));
    h.cap = cap; h.gcap = gcap;
  }
  double *d_th = h.d_batch, *d_lp = d_th + (size_t)h.cap * P, *d_g = d_lp + h.cap, *d_r = d_g + (size_t)h.gcap * P;
  CUDA_TRY(cudaMemcpyAsync(d_th, thetas, sizeof(double) * (size_t)B * P, cudaMemcpyHostToDevice, h.st));
  k_gmm_logprob_batch<<<B, GMM_THREADS, h.smem, h.st>>>(h.m, d_th, d_lp, grad_out ? d_g : nullptr, (grad_out && !h.m.r_smem) ? d_r : nullptr, jacobian, propto);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(lp_out, d_lp, sizeof(double) * B, cudaMemcpyDeviceToHost, h.st));
  if (grad_out) CUDA_TRY(cudaMemcpyAsync(grad_out, d_g, sizeof(double) * (size_t)B * P, cudaMemcpyDeviceToHost, h.st));
  CUDA_TRY(cudaStreamSynchronize(h.st));
  return BVG_OK;
}
// posterior pass for given unconstrained draws: draws_out [nd][K + 2 K D] (theta | mu column-major | sigma) or NULL, soft_out [N][K]
extern "C" int bvg_gmm_posterior(void *handle, const double *thetas, int nd, double *draws_out, double *soft_out, double *loglik_out) {
  if (!handle || !thetas || !soft_out || nd < 1) { bvg_set_error("null or empty argument"); return BVG_ERR_ARG; }
  GmmHost &h = *(GmmHost *)handle;
  CUDA_TRY(cudaSetDevice(h.device));
  const int P = h.m.P, K = h.m.K, D = h.m.D;
  const size_t nk = (size_t)h.m.N * K, ncol = (size_t)K + 2 * (size_t)K * D;
  double *buf = nullptr;
  CUDA_TRY(cudaMalloc(&buf, sizeof(double) * ((size_t)nd * P + nk + (draws_out ? (size_t)nd * ncol : 0))));
  double *d_th = buf, *d_soft = d_th + (size_t)nd * P, *d_draws = draws_out ? d_soft + nk : nullptr;
  bool ok = cudaMemcpyAsync(d_th, thetas, sizeof(double) * (size_t)nd * P, cudaMemcpyHostToDevice, h.st) == cudaSuccess;
  k_gmm_post_given<<<1, GMM_THREADS, h.smem, h.st>>>(h.m, nd, d_th, d_draws, d_soft);
  double sc[6];
  ok = ok && cudaMemcpyAsync(sc, h.m.scal, sizeof(sc), cudaMemcpyDeviceToHost, h.st) == cudaSuccess &&
       cudaMemcpyAsync(soft_out, d_soft, sizeof(double) * nk, cudaMemcpyDeviceToHost, h.st) == cudaSuccess &&
       (!draws_out || cudaMemcpyAsync(draws_out, d_draws, sizeof(double) * (size_t)nd * ncol, cudaMemcpyDeviceToHost, h.st) == cudaSuccess) &&
       cudaStreamSynchronize(h.st) == cudaSuccess && cudaGetLastError() == cudaSuccess;
  cudaFree(buf);
  if (!ok) { bvg_set_error("gmm: %s", cudaGetErrorString(cudaGetLastError())); return BVG_ERR_CUDA; }
  if (loglik_out) *loglik_out = sc[4];
  return BVG_OK;
}
