// k_pathfinder.cu -- algorithm = "pathfinder" (R/findSpatiallyVariableFeaturesBayes.R:366-375:
//   mod$pathfinder(init = 0.01, num_paths = n.cores, max_lbfgs_iters = 200, num_elbo_draws = elbo.samples, history_size = 25,
//                  draws = n.draws))  [U: stan::services::pathfinder::pathfinder_lbfgs_single / _multi]
//
// Along the L-BFGS path of f = -log_prob<propto, jacobian = true> every iterate (x, g) gets the Gaussian approximation
// N(xc, Sigma) implied by the last <= J curvature pairs and the diagonal alpha (oracle/bvg_oracle.c states the algebra
// and checks it against the BFGS recursion).  Everything of length D stays on the device:
//   W = [Y sqrt(alpha); C / sqrt(alpha)]  (2n x D, C = -R^{-1} S'),  xc,  alpha,  the K normals U and the K draws PHI;
// the host only sees the O(n^2) pieces (S'Y, Y' alpha Y, W W', W U) and does the <= 50 x 50 Cholesky factorisations and
// triangular solves.  The products over D are one kernel, k_pf_prod: out[i][j] = sum_d A[i][d] B[j][d] w[d], chunked
// over D with a fixed-order second stage (deterministic, no atomics).  The log densities of the draws come from the same
// likelihood engine as everything else (forward-only pass, propto = 1, jacobian = 1).
#include <algorithm>
#include <cmath>
#include <vector>

#include "bvg_internal.cuh"
#include "bvg_vec.cuh"

#define PF_MAXJ 25
#define PF_NCH 96          // D chunks of the product kernel
#define PF_TILE 32         // columns per shared-memory tile
enum { STREAM_PF_INIT = 5, STREAM_PF_ELBO = 6, STREAM_PF_DRAW = 7, STREAM_PF_PSIS = 8 };

struct PfApprox {           // host part of an approximation; the D-length parts live in PfState
  int n = 0, ok = 0;
  double Lc[50 * 50], La[50 * 50], logdet = 0;
};
struct PfState {
  double *alpha, *Sh, *Yh, *xprev, *gprev, *s, *y, *W, *xc, *Wb, *xcb, *alphab, *U, *PHI, *part, *out, *coef;
  double *base;
  int nh = 0, head = 0;
  PfApprox cur, best;
  double best_elbo = -INFINITY;
  int best_iter = -1, n_iter = 0;
  std::vector<double> elbos;
};

// out_part[chunk][i][j] = sum over the chunk's d of A[i][d] B[j][d] w[d]   (w may be NULL)
__global__ void __launch_bounds__(256) k_pf_prod(const double *__restrict__ A, int na, const double *__restrict__ Bm, int nbm,
                                                 const double *__restrict__ w, int64_t D, double *__restrict__ part) {
  extern __shared__ double sh[];   // [na][PF_TILE] | [nbm][PF_TILE]
  double *sa = sh, *sb = sh + na * PF_TILE;
  const int64_t per = (D + gridDim.x - 1) / gridDim.x, d0 = (int64_t)blockIdx.x * per, d1 = d0 + per < D ? d0 + per : D;
  const int npair = na * nbm;
  double acc[10];
#pragma unroll
  for (int q = 0; q < 10; ++q) acc[q] = 0.0;
  for (int64_t t0 = d0; t0 < d1; t0 += PF_TILE) {
    for (int e = threadIdx.x; e < (na + nbm) * PF_TILE; e += blockDim.x) {
      const int r = e / PF_TILE, c = e - r * PF_TILE;
      const int64_t d = t0 + c;
      double v = 0.0;
      if (d < d1) v = r < na ? A[(size_t)r * D + d] * (w ? w[d] : 1.0) : Bm[(size_t)(r - na) * D + d];
      sh[e] = v;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 10; ++q) {
      const int pr = threadIdx.x + q * 256;
      if (pr < npair) {
        const int i = pr / nbm, j = pr - i * nbm;
        double s = acc[q];
#pragma unroll 8
        for (int c = 0; c < PF_TILE; ++c) s = fma(sa[i * PF_TILE + c], sb[j * PF_TILE + c], s);
        acc[q] = s;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < 10; ++q) {
    const int pr = threadIdx.x + q * 256;
    if (pr < npair) part[(size_t)blockIdx.x * npair + pr] = acc[q];
  }
}
__global__ void k_pf_prod_sum(const double *__restrict__ part, int nch, int npair, double *__restrict__ out) {
  const int pr = blockIdx.x * blockDim.x + threadIdx.x;
  if (pr >= npair) return;
  double s = 0;
  for (int c = 0; c < nch; ++c) s += part[(size_t)c * npair + pr];   // fixed order
  out[pr] = s;
}
// s = x - xprev, y = g - gprev
__global__ void k_pf_diff(const double *__restrict__ x, const double *__restrict__ g, const double *__restrict__ xp,
                          const double *__restrict__ gp, double *__restrict__ s, double *__restrict__ y, int64_t D) {
  for (int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; d < D; d += (int64_t)gridDim.x * blockDim.x) {
    s[d] = x[d] - xp[d];
    y[d] = g[d] - gp[d];
  }
}
// form_diag [U]: alpha <- y's / (y'ay / alpha + y^2 - (y'ay / s'a^{-1}s) (s / alpha)^2)
__global__ void k_pf_alpha(double *__restrict__ alpha, const double *__restrict__ s, const double *__restrict__ y, double ys,
                           double yay, double sias, int64_t D) {
  for (int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; d < D; d += (int64_t)gridDim.x * blockDim.x) {
    const double a = alpha[d], sa = s[d] / a;
    alpha[d] = ys / (yay / a + y[d] * y[d] - (yay / sias) * sa * sa);
  }
}
// aux rows used by the scalar products: r0 = 1 / alpha (weight), r1 = sqrt(alpha) g, r2 = log alpha, r3 = 1
__global__ void k_pf_aux(const double *__restrict__ alpha, const double *__restrict__ g, double *__restrict__ aux, int64_t D) {
  for (int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; d < D; d += (int64_t)gridDim.x * blockDim.x) {
    const double a = alpha[d];
    aux[d] = 1.0 / a;
    aux[D + d] = sqrt(a) * (g ? g[d] : 0.0);
    aux[2 * D + d] = log(a);
    aux[3 * D + d] = 1.0;
  }
}
// W[i] = Y[h_i] sqrt(alpha);  W[n + i] = (sum_{q >= i} C[i][q] S[h_q]) / sqrt(alpha);  hist[q] = slot of pair q (oldest first)
struct PfIdx { int h[PF_MAXJ]; };
__global__ void k_pf_buildW(const double *__restrict__ Sh, const double *__restrict__ Yh, const double *__restrict__ alpha,
                            const double *__restrict__ Cm, int n, PfIdx idx, int64_t D, double *__restrict__ W) {
  __shared__ double sc[PF_MAXJ * PF_MAXJ];
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) sc[e] = Cm[e];
  __syncthreads();
  for (int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; d < D; d += (int64_t)gridDim.x * blockDim.x) {
    const double sq = sqrt(alpha[d]), isq = 1.0 / sq;
    double sv[PF_MAXJ];
    for (int q = 0; q < n; ++q) sv[q] = Sh[(size_t)idx.h[q] * D + d];
    for (int i = 0; i < n; ++i) {
      W[(size_t)i * D + d] = Yh[(size_t)idx.h[i] * D + d] * sq;
      double cv = 0;
      for (int q = i; q < n; ++q) cv = fma(sc[i * n + q], sv[q], cv);
      W[(size_t)(n + i) * D + d] = cv * isq;
    }
  }
}
// xc = x - alpha g - sqrt(alpha) sum_r W[r] coef[r]
__global__ void k_pf_center(const double *__restrict__ x, const double *__restrict__ g, const double *__restrict__ alpha,
                            const double *__restrict__ W, int n2, const double *__restrict__ coef, int64_t D, double *__restrict__ xc) {
  __shared__ double sc[2 * PF_MAXJ];
  if ((int)threadIdx.x < n2) sc[threadIdx.x] = coef[threadIdx.x];
  __syncthreads();
  for (int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; d < D; d += (int64_t)gridDim.x * blockDim.x) {
    double s = 0;
    for (int r = 0; r < n2; ++r) s = fma(W[(size_t)r * D + d], sc[r], s);
    const double a = alpha[d];
    xc[d] = x[d] - a * g[d] - sqrt(a) * s;
  }
}
__global__ void k_pf_normals(uint64_t seed, uint32_t stream, uint32_t t0, int K, int64_t D, double *__restrict__ U) {
  const int64_t n = (int64_t)K * D;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = q / D, d = q - k * D;
    U[q] = bvg_normal(seed, stream, t0 + (uint32_t)k, (uint64_t)d);
  }
}
// PHI[k] = xc + sqrt(alpha) (U[k] + sum_r W[r] CC[r][k])
__global__ void __launch_bounds__(256) k_pf_phi(const double *__restrict__ xc, const double *__restrict__ alpha, const double *__restrict__ W,
                                                int n2, const double *__restrict__ CC /* [n2][K] */, const double *__restrict__ U, int K,
                                                int64_t D, double *__restrict__ PHI) {
  extern __shared__ double scc[];
  for (int e = threadIdx.x; e < n2 * K; e += blockDim.x) scc[e] = CC[e];
  __syncthreads();
  for (int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; d < D; d += (int64_t)gridDim.x * blockDim.x) {
    double wv[2 * PF_MAXJ];
    for (int r = 0; r < n2; ++r) wv[r] = W[(size_t)r * D + d];
    const double c = xc[d], sq = sqrt(alpha[d]);
    for (int k = 0; k < K; ++k) {
      double s = U[(size_t)k * D + d];
      for (int r = 0; r < n2; ++r) s = fma(wv[r], scc[r * K + k], s);
      PHI[(size_t)k * D + d] = c + sq * s;
    }
  }
}
// amp[(row0 + k)][g] = exp(PHI[k][amp + g])
__global__ void k_pf_amp(const double *__restrict__ PHI, int64_t D, int64_t amp, int64_t G, int K, int64_t row0, double *__restrict__ out) {
  const int64_t n = (int64_t)K * G;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = q / G, g = q - k * G;
    out[(row0 + k) * G + g] = exp(PHI[(size_t)k * D + amp + g]);
  }
}
// draws[g][r] = amp[idx[r]][g]  (the layout k_summary reads)
__global__ void k_pf_gather(const double *__restrict__ amp, const int *__restrict__ idx, int ns, int64_t G, double *__restrict__ draws) {
  const int64_t n = (int64_t)ns * G;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t g = q / ns, r = q - g * ns;
    draws[q] = amp[(size_t)idx[r] * G + g];
  }
}
__global__ void k_pf_init(uint64_t seed, uint32_t path, double radius, int64_t D, double *__restrict__ x) {
  for (int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; d < D; d += (int64_t)gridDim.x * blockDim.x) {
    uint32_t o[4];
    philox4x32_10((uint32_t)d, (uint32_t)((uint64_t)d >> 32), path, STREAM_PF_INIT, (uint32_t)seed, (uint32_t)(seed >> 32), o);
    const double u = ((double)(((uint64_t)o[0] << 21) | (o[2] >> 11)) + 0.5) * (1.0 / 9007199254740992.0);
    x[d] = radius * (2.0 * u - 1.0);
  }
}

namespace {
int chol_lower(double *A, int n) {
  for (int j = 0; j < n; ++j) {
    double d = A[j * n + j];
    for (int q = 0; q < j; ++q) d -= A[j * n + q] * A[j * n + q];
    if (!(d > 0.0) || !std::isfinite(d)) return 1;
    d = std::sqrt(d);
    A[j * n + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double v = A[i * n + j];
      for (int q = 0; q < j; ++q) v -= A[i * n + q] * A[j * n + q];
      A[i * n + j] = v / d;
    }
    for (int i = 0; i < j; ++i) A[i * n + j] = 0.0;
  }
  return 0;
}
// out (host) [na][nbm] = A B' (weighted)
int pf_prod(bvg_fit *f, PfState *st, const double *A, int na, const double *Bm, int nbm, const double *w, double *h_out) {
  const int64_t D = f->m.L.D;
  const int npair = na * nbm;
  if (npair == 0) return BVG_OK;
  if (npair > 2560) { bvg_set_error("pathfinder: product too large"); return BVG_ERR_ARG; }
  const size_t sh = sizeof(double) * (size_t)(na + nbm) * PF_TILE;
  k_pf_prod<<<PF_NCH, 256, sh, f->stream>>>(A, na, Bm, nbm, w, D, st->part);
  k_pf_prod_sum<<<(npair + 255) / 256, 256, 0, f->stream>>>(st->part, PF_NCH, npair, st->out);
  f->st.kernel_launches += 2;
  CUDA_TRY(cudaMemcpyAsync(h_out, st->out, sizeof(double) * npair, cudaMemcpyDeviceToHost, f->stream));
  CUDA_TRY(cudaStreamSynchronize(f->stream));
  return BVG_OK;
}
const int kGrid = 148 * 4;

// the approximation at (x, g) from the current history: device W, xc; host factors in st->cur
int pf_build(bvg_fit *f, PfState *st, const double *d_x, const double *d_g, int J) {
  const int64_t D = f->m.L.D;
  const int n = st->nh, n2 = 2 * n;
  PfApprox &a = st->cur;
  a.n = n; a.ok = 0;
  PfIdx idx;
  for (int q = 0; q < n; ++q) idx.h[q] = (st->head - n + q + 2 * J) % J;
  // compact copies of the live history rows, oldest first (the product kernel wants contiguous rows)
  double *Sc = st->PHI, *Yc = st->PHI + (size_t)PF_MAXJ * D;   // PHI is free while an approximation is being built
  for (int q = 0; q < n; ++q) {
    CUDA_TRY(cudaMemcpyAsync(Sc + (size_t)q * D, st->Sh + (size_t)idx.h[q] * D, sizeof(double) * D, cudaMemcpyDeviceToDevice, f->stream));
    CUDA_TRY(cudaMemcpyAsync(Yc + (size_t)q * D, st->Yh + (size_t)idx.h[q] * D, sizeof(double) * D, cudaMemcpyDeviceToDevice, f->stream));
  }
  double *aux = st->U;   // [4][D]: 1/alpha, sqrt(alpha) g, log alpha, 1   (U is free here too)
  k_pf_aux<<<kGrid, 256, 0, f->stream>>>(st->alpha, d_g, aux, D);
  f->st.kernel_launches += 1;
  std::vector<double> SY((size_t)n * n), YaY((size_t)n * n), Rinv((size_t)n * n, 0.0), Cm((size_t)n * n, 0.0);
  BVG_TRY(pf_prod(f, st, Sc, n, Yc, n, nullptr, SY.data()));
  BVG_TRY(pf_prod(f, st, Yc, n, Yc, n, st->alpha, YaY.data()));
  for (int col = 0; col < n; ++col)
    for (int i = n - 1; i >= 0; --i) {
      double v = i == col ? 1.0 : 0.0;
      for (int q = i + 1; q < n; ++q) v -= SY[i * n + q] * Rinv[q * n + col];
      Rinv[i * n + col] = i > col ? 0.0 : v / SY[i * n + i];
    }
  for (int i = 0; i < n * n; ++i) Cm[i] = -Rinv[i];
  if (n > 0) {
    CUDA_TRY(cudaMemcpyAsync(st->coef, Cm.data(), sizeof(double) * n * n, cudaMemcpyHostToDevice, f->stream));
    CUDA_TRY(cudaStreamSynchronize(f->stream));
    for (int q = 0; q < n; ++q) idx.h[q] = q;   // the compact copies are already in order
    k_pf_buildW<<<kGrid, 128, 0, f->stream>>>(Sc, Yc, st->alpha, st->coef, n, idx, D, st->W);
    f->st.kernel_launches += 1;
  }
  // Gram of W, and the products with sqrt(alpha) g (t2 | t1) and sum log alpha
  double *G = a.Lc;
  std::vector<double> tv((size_t)std::max(n2, 1)), misc(1);
  BVG_TRY(pf_prod(f, st, st->W, n2, st->W, n2, nullptr, G));
  BVG_TRY(pf_prod(f, st, st->W, n2, aux + D, 1, nullptr, tv.data()));
  BVG_TRY(pf_prod(f, st, aux + 2 * D, 1, aux + 3 * D, 1, nullptr, misc.data()));
  if (n2 > 0 && chol_lower(G, n2)) return BVG_OK;   // linearly dependent history: no approximation at this iterate
  double Mk[50 * 50], T[50 * 50], H[50 * 50];
  std::fill(Mk, Mk + n2 * n2, 0.0);
  for (int i = 0; i < n; ++i) {
    Mk[i * n2 + n + i] = Mk[(n + i) * n2 + i] = 1.0;
    for (int j = 0; j < n; ++j) Mk[(n + i) * n2 + n + j] = YaY[i * n + j] + (i == j ? SY[i * n + i] : 0.0);
  }
  for (int i = 0; i < n2; ++i)
    for (int j = 0; j < n2; ++j) {
      double v = 0;
      for (int q = 0; q < n2; ++q) v += Mk[i * n2 + q] * G[q * n2 + j];
      T[i * n2 + j] = v;
    }
  for (int i = 0; i < n2; ++i)
    for (int j = 0; j < n2; ++j) {
      double v = i == j ? 1.0 : 0.0;
      for (int q = 0; q < n2; ++q) v += G[q * n2 + i] * T[q * n2 + j];
      H[i * n2 + j] = v;
    }
  for (int i = 0; i < n2; ++i)
    for (int j = 0; j < i; ++j) { const double v = 0.5 * (H[i * n2 + j] + H[j * n2 + i]); H[i * n2 + j] = H[j * n2 + i] = v; }
  std::copy(H, H + n2 * n2, a.La);
  if (n2 > 0 && chol_lower(a.La, n2)) return BVG_OK;
  a.logdet = 0.5 * misc[0];
  for (int i = 0; i < n2; ++i) a.logdet += std::log(a.La[i * n2 + i]);
  // xc = x - alpha g - sqrt(alpha) W' [t1; v],  t2 = tv[0..n), t1 = tv[n..2n), v = t2 + (YaY + Dk) t1
  std::vector<double> coef((size_t)std::max(n2, 1), 0.0);
  for (int i = 0; i < n; ++i) {
    double s = tv[i];
    for (int j = 0; j < n; ++j) s += (YaY[i * n + j] + (i == j ? SY[i * n + i] : 0.0)) * tv[n + j];
    coef[i] = tv[n + i];
    coef[n + i] = s;
  }
  CUDA_TRY(cudaMemcpyAsync(st->coef, coef.data(), sizeof(double) * std::max(n2, 1), cudaMemcpyHostToDevice, f->stream));
  CUDA_TRY(cudaStreamSynchronize(f->stream));
  k_pf_center<<<kGrid, 256, 0, f->stream>>>(d_x, d_g, st->alpha, st->W, n2, st->coef, D, st->xc);
  f->st.kernel_launches += 1;
  CUDA_TRY(cudaGetLastError());
  a.ok = 1;
  return BVG_OK;
}
// K draws from an approximation (device W / xc / alpha given): PHI, log q (host), log p (host)
int pf_draws(bvg_fit *f, PfState *st, const PfApprox &a, const double *W, const double *xc, const double *alpha, uint32_t stream,
             uint32_t t0, int K, std::vector<double> &lq, std::vector<double> &lp) {
  const int64_t D = f->m.L.D;
  const int n2 = 2 * a.n;
  k_pf_normals<<<kGrid, 256, 0, f->stream>>>(f->m.seed, stream, t0, K, D, st->U);
  f->st.kernel_launches += 1;
  std::vector<double> WU((size_t)std::max(n2, 1) * K), UU((size_t)K * K), CC((size_t)std::max(n2, 1) * K, 0.0);
  BVG_TRY(pf_prod(f, st, W, n2, st->U, K, nullptr, WU.data()));
  lq.resize(K); lp.resize(K);
  // u'u: the diagonal of U U' would waste K^2; K products of one row each are cheap enough through the same kernel
  {
    const int step = 50;
    for (int k0 = 0; k0 < K; k0 += step) {
      const int kk = std::min(step, K - k0);
      BVG_TRY(pf_prod(f, st, st->U + (size_t)k0 * D, kk, st->U + (size_t)k0 * D, kk, nullptr, UU.data()));
      for (int k = 0; k < kk; ++k) lq[k0 + k] = -a.logdet - 0.5 * (UU[(size_t)k * kk + k] + (double)D * BVG_LOG_TWO_PI);
    }
  }
  for (int k = 0; k < K; ++k) {
    double u1[50], z[50], cc[50];
    for (int i = 0; i < n2; ++i) {
      double s = WU[(size_t)i * K + k];
      for (int q = 0; q < i; ++q) s -= a.Lc[i * n2 + q] * u1[q];
      u1[i] = s / a.Lc[i * n2 + i];
    }
    for (int i = 0; i < n2; ++i) {
      double s = -u1[i];
      for (int q = 0; q <= i; ++q) s += a.La[i * n2 + q] * u1[q];
      z[i] = s;
    }
    for (int i = n2 - 1; i >= 0; --i) {
      double s = z[i];
      for (int q = i + 1; q < n2; ++q) s -= a.Lc[q * n2 + i] * cc[q];
      cc[i] = s / a.Lc[i * n2 + i];
    }
    for (int i = 0; i < n2; ++i) CC[(size_t)i * K + k] = cc[i];
  }
  CUDA_TRY(cudaMemcpyAsync(st->coef, CC.data(), sizeof(double) * std::max(n2, 1) * K, cudaMemcpyHostToDevice, f->stream));
  CUDA_TRY(cudaStreamSynchronize(f->stream));
  k_pf_phi<<<kGrid, 256, sizeof(double) * std::max(n2, 1) * K, f->stream>>>(xc, alpha, W, n2, st->coef, st->U, K, D, st->PHI);
  f->st.kernel_launches += 1;
  // log p of every draw: forward-only pass of the likelihood engine, log_prob<propto = true, jacobian = true>
  if (f->elbo_cap < K) {
    cudaFree(f->m.elbo_lps);
    f->m.elbo_lps = nullptr; f->elbo_cap = 0;
    CUDA_TRY(cudaMalloc(&f->m.elbo_lps, sizeof(double) * K));
    f->elbo_cap = K;
  }
  Ctl h;
  CUDA_TRY(cudaMemcpyAsync(&h, f->m.ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, f->stream));
  CUDA_TRY(cudaStreamSynchronize(f->stream));
  h.t_draw = 0;
  CUDA_TRY(cudaMemcpyAsync(f->m.ctl, &h, sizeof(Ctl), cudaMemcpyHostToDevice, f->stream));
  CUDA_TRY(cudaStreamSynchronize(f->stream));
  double *zeta0 = f->m.zeta;
  int rc = BVG_OK;
  for (int k = 0; k < K && rc == BVG_OK; ++k) {
    f->m.zeta = st->PHI + (size_t)k * D;
    rc = launch_prep(f, PREP_NONE);
    if (rc == BVG_OK) rc = launch_lik(f, false);
    if (rc == BVG_OK) rc = launch_gene_finish(f, MODE_POST_DRAW, 1, 1);
  }
  f->m.zeta = zeta0;
  BVG_TRY(rc);
  CUDA_TRY(cudaMemcpyAsync(lp.data(), f->m.elbo_lps, sizeof(double) * K, cudaMemcpyDeviceToHost, f->stream));
  CUDA_TRY(cudaStreamSynchronize(f->stream));
  return BVG_OK;
}
}  // namespace

// Pareto-smoothed importance weights [U: stan psis_weights; Zhang & Stephens 2009 gpdfit]; returns the k estimate
double pf_psis_weights(const std::vector<double> &lr_in, std::vector<double> &w) {
  const int S = (int)lr_in.size();
  double mx = -INFINITY, khat = NAN;
  for (double v : lr_in) mx = std::max(mx, v);
  std::vector<double> lr(S);
  for (int i = 0; i < S; ++i) lr[i] = lr_in[i] - mx;
  const int tail = (int)std::fmin(0.2 * S, 3.0 * std::sqrt((double)S));
  if (tail >= 5) {
    std::vector<std::pair<double, int>> srt(S);
    for (int i = 0; i < S; ++i) srt[i] = {lr[i], i};
    std::stable_sort(srt.begin(), srt.end(), [](const std::pair<double, int> &x, const std::pair<double, int> &y) { return x.first < y.first; });
    const double cutoff = srt[S - tail - 1].first, ec = std::exp(cutoff);
    const int n = tail;
    std::vector<double> x(n);
    for (int i = 0; i < n; ++i) x[i] = std::exp(srt[S - tail + i].first) - ec;
    if (x[n - 1] > 0 && std::isfinite(x[n - 1])) {
      const int M = 30 + (int)std::floor(std::sqrt((double)n));
      const double xstar = x[(int)std::floor(n / 4.0 + 0.5) - 1], prior = 3.0;
      std::vector<double> th(M), lt(M);
      for (int j = 0; j < M; ++j) {
        th[j] = 1.0 / x[n - 1] + (1.0 - std::sqrt((double)M / (j + 0.5))) / (prior * xstar);
        double kk = 0;
        for (int i = 0; i < n; ++i) kk += std::log1p(-th[j] * x[i]);
        kk /= n;
        lt[j] = n * (std::log(-th[j] / kk) - kk - 1.0);
      }
      double theta = 0;
      for (int j = 0; j < M; ++j) {
        double sw = 0;
        for (int q = 0; q < M; ++q) sw += std::exp(lt[q] - lt[j]);
        theta += th[j] / sw;
      }
      double k = 0;
      for (int i = 0; i < n; ++i) k += std::log1p(-theta * x[i]);
      k /= n;
      const double sigma = -k / theta;
      k = k * n / (n + 10.0) + 0.5 * 10.0 / (n + 10.0);
      khat = k;
      if (std::isfinite(k) && std::isfinite(sigma))
        for (int i = 0; i < n; ++i) {
          const double pq = (i + 0.5) / n;
          const double q = (std::fabs(k) < 1e-300 ? -sigma * std::log1p(-pq) : sigma * std::expm1(-k * std::log1p(-pq)) / k) + ec;
          lr[srt[S - tail + i].second] = std::fmin(std::log(q), 0.0);
        }
    }
  }
  double lse = 0;
  for (int i = 0; i < S; ++i) lse += std::exp(lr[i]);
  w.resize(S);
  for (int i = 0; i < S; ++i) w[i] = std::exp(lr[i]) / lse;
  return khat;
}

// One path.  d_amp (device) [num_draws][G] rows row0..; h_lr (host) [num_draws]; info = {lbfgs_iters, lbfgs_term, best_iter,
// n_pairs, best_elbo}; elbos (host, optional) [max_iters + 1]
int pf_single(bvg_fit *f, int path, double *d_amp, int64_t row0, double *h_lr, double *info, double *elbos) {
  const bvg_config &c = f->cfg;
  const int64_t D = f->m.L.D, G = f->m.G;
  const int J = std::min(std::max(c.lbfgs_history, 1), PF_MAXJ), K = c.elbo_samples, nd = c.pf_single_draws, KB = std::max(K, 50);
  if (f->comm_mode || c.world_size > 1) { bvg_set_error("algorithm = 'pathfinder' is not built for gene-sharded fits"); return BVG_ERR_ARG; }
  if (K < 1 || K > 50 || nd < 1 || c.pf_max_lbfgs_iters < 1 || c.pf_max_lbfgs_iters > 4095 || nd > 65536) {
    bvg_set_error("pathfinder: need 1 <= num_elbo_draws <= 50, 1 <= max_lbfgs_iters <= 4095, 1 <= single-path draws <= 65536");
    return BVG_ERR_ARG;
  }
  PfState st;
  const size_t nrow = 3 + 2 + 2 + 2 * (size_t)J /* Sh, Yh */ + 4 * (size_t)J /* W, Wb */ + 2 * (size_t)KB /* U, PHI */;
  const size_t extra = (size_t)PF_NCH * 2560 + 2560 + 2560;
  CUDA_TRY(cudaMalloc(&st.base, sizeof(double) * (nrow * D + extra)));
  double *p = st.base;
  st.alpha = p; p += D; st.xprev = p; p += D; st.gprev = p; p += D; st.s = p; p += D; st.y = p; p += D; st.xc = p; p += D; st.xcb = p; p += D;
  st.Sh = p; p += (size_t)J * D; st.Yh = p; p += (size_t)J * D; st.W = p; p += 2 * (size_t)J * D; st.Wb = p; p += 2 * (size_t)J * D;
  st.U = p; p += (size_t)KB * D; st.PHI = p; p += (size_t)KB * D;
  st.part = p; p += (size_t)PF_NCH * 2560; st.out = p; p += 2560; st.coef = p;
  double *alphab = nullptr;   // the best iterate's alpha
  auto fail = [&](int code) { cudaFree(st.base); cudaFree(alphab); return code; };
  if (cudaMalloc(&alphab, sizeof(double) * D) != cudaSuccess) { cudaGetLastError(); bvg_set_error("out of device memory"); return fail(BVG_ERR_CUDA); }
  st.alphab = alphab;
  if (KB < 2 * PF_MAXJ || KB < 4) { bvg_set_error("pathfinder: internal buffer sizing"); return fail(BVG_ERR_ARG); }
  int rc = vec_fill(f->stream, st.alpha, 1.0, D);
  if (rc != BVG_OK) return fail(rc);
  st.elbos.assign((size_t)c.pf_max_lbfgs_iters + 1, NAN);
  CUDA_TRY(cudaFuncSetAttribute(k_pf_phi, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * 50 * 50)));
  k_pf_init<<<kGrid, 256, 0, f->stream>>>(c.seed, (uint32_t)path, c.pf_init_radius, D, st.xc);   // xc doubles as theta0
  LbfgsOpt opt;
  opt.jacobian = 1; opt.propto = 1; opt.max_iter = c.pf_max_lbfgs_iters; opt.history = J;
  opt.on_iterate = [&](int it, const double *d_x, const double *d_g, double) -> int {
    if (it > 0) {
      k_pf_diff<<<kGrid, 256, 0, f->stream>>>(d_x, d_g, st.xprev, st.gprev, st.s, st.y, D);
      k_pf_aux<<<kGrid, 256, 0, f->stream>>>(st.alpha, nullptr, st.U, D);   // U[0] = 1 / alpha
      f->st.kernel_launches += 2;
      double ys, yy, yay, sias;
      BVG_TRY(pf_prod(f, &st, st.y, 1, st.s, 1, nullptr, &ys));
      BVG_TRY(pf_prod(f, &st, st.y, 1, st.y, 1, nullptr, &yy));
      BVG_TRY(pf_prod(f, &st, st.y, 1, st.y, 1, st.alpha, &yay));
      BVG_TRY(pf_prod(f, &st, st.s, 1, st.s, 1, st.U, &sias));
      if (ys > 0 && std::fabs(yy / ys) <= 1e12) {   // check_curve [U]
        k_pf_alpha<<<kGrid, 256, 0, f->stream>>>(st.alpha, st.s, st.y, ys, yay, sias, D);
        CUDA_TRY(cudaMemcpyAsync(st.Sh + (size_t)st.head * D, st.s, sizeof(double) * D, cudaMemcpyDeviceToDevice, f->stream));
        CUDA_TRY(cudaMemcpyAsync(st.Yh + (size_t)st.head * D, st.y, sizeof(double) * D, cudaMemcpyDeviceToDevice, f->stream));
        st.head = (st.head + 1) % J;
        if (st.nh < J) ++st.nh;
        f->st.kernel_launches += 1;
      }
      BVG_TRY(pf_build(f, &st, d_x, d_g, J));
      double elbo = -INFINITY;
      if (st.cur.ok) {
        std::vector<double> lq, lp;
        const uint32_t t0 = ((uint32_t)path << 24) | ((uint32_t)it << 12);
        BVG_TRY(pf_draws(f, &st, st.cur, st.W, st.xc, st.alpha, STREAM_PF_ELBO, t0, K, lq, lp));
        double acc = 0;
        bool fin = true;
        for (int k = 0; k < K && fin; ++k) {
          if (!std::isfinite(lp[k]) || !std::isfinite(lq[k])) fin = false; else acc += lp[k] - lq[k];
        }
        if (fin) elbo = acc / K;
        f->st.elbo_evals++;
      }
      st.elbos[it] = elbo;
      if (elbo > st.best_elbo) {
        st.best_elbo = elbo; st.best_iter = it; st.best = st.cur;
        CUDA_TRY(cudaMemcpyAsync(st.Wb, st.W, sizeof(double) * 2 * (size_t)st.cur.n * D, cudaMemcpyDeviceToDevice, f->stream));
        CUDA_TRY(cudaMemcpyAsync(st.xcb, st.xc, sizeof(double) * D, cudaMemcpyDeviceToDevice, f->stream));
        CUDA_TRY(cudaMemcpyAsync(st.alphab, st.alpha, sizeof(double) * D, cudaMemcpyDeviceToDevice, f->stream));
      }
      st.n_iter = it;
    }
    CUDA_TRY(cudaMemcpyAsync(st.xprev, d_x, sizeof(double) * D, cudaMemcpyDeviceToDevice, f->stream));
    CUDA_TRY(cudaMemcpyAsync(st.gprev, d_g, sizeof(double) * D, cudaMemcpyDeviceToDevice, f->stream));
    return BVG_OK;
  };
  // theta0 must survive the first callbacks: L-BFGS copies it into its own buffer before evaluating
  rc = lbfgs_run(f, st.xc, f->scratch, &opt);
  if (rc != BVG_OK) return fail(rc);
  if (st.best_iter < 0) { bvg_set_error("Pathfinder: no iterate of path %d produced a finite ELBO estimate", path); return fail(BVG_ERR_NUMERIC); }
  for (int k0 = 0; k0 < nd; k0 += KB) {
    const int kk = std::min(KB, nd - k0);
    std::vector<double> lq, lp;
    const uint32_t t0 = ((uint32_t)path << 16) | (uint32_t)k0;
    rc = pf_draws(f, &st, st.best, st.Wb, st.xcb, st.alphab, STREAM_PF_DRAW, t0, kk, lq, lp);
    if (rc != BVG_OK) return fail(rc);
    for (int k = 0; k < kk; ++k) h_lr[k0 + k] = std::isfinite(lp[k]) ? lp[k] - lq[k] : -INFINITY;
    k_pf_amp<<<kGrid, 256, 0, f->stream>>>(st.PHI, D, f->m.L.amp, G, kk, row0 + k0, d_amp);
    f->st.kernel_launches += 1;
  }
  if (cudaStreamSynchronize(f->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) { bvg_set_error("pathfinder: CUDA failure"); return fail(BVG_ERR_CUDA); }
  if (info) { info[0] = f->st.mle_iters; info[1] = f->st.mle_term; info[2] = st.best_iter; info[3] = st.best.n; info[4] = st.best_elbo; }
  if (elbos) std::copy(st.elbos.begin(), st.elbos.end(), elbos);
  return fail(BVG_OK);
}

// multi-path Pathfinder -> f->draws ([G][n_draws], PSIS-resampled amplitudes) ready for k_summary
int pf_run(bvg_fit *f) {
  const bvg_config &c = f->cfg;
  const int64_t G = f->m.G;
  const int P = c.pf_num_paths, nd = c.pf_single_draws, S = P * nd, ns = c.n_draws;
  if (P < 1 || P > 255) { bvg_set_error("pathfinder: 1 <= num_paths <= 255"); return BVG_ERR_ARG; }
  double *d_amp = nullptr;
  int *d_idx = nullptr;
  CUDA_TRY(cudaMalloc(&d_amp, sizeof(double) * (size_t)S * G));
  if (cudaMalloc(&d_idx, sizeof(int) * ns) != cudaSuccess) { cudaFree(d_amp); cudaGetLastError(); bvg_set_error("out of device memory"); return BVG_ERR_CUDA; }
  std::vector<double> lr(S), w;
  std::vector<int> idx(ns);
  int rc = BVG_OK;
  f->n_trace = 0;
  int iters_total = 0;
  double best = -INFINITY;
  for (int p = 0; p < P && rc == BVG_OK; ++p) {
    double info[5];
    rc = pf_single(f, p, d_amp, (int64_t)p * nd, lr.data() + (size_t)p * nd, info, nullptr);
    if (rc == BVG_OK) {
      if (f->n_trace < 64) { double *r = f->trace + 4 * f->n_trace++; r[0] = p; r[1] = info[2]; r[2] = info[4]; r[3] = info[0]; }
      iters_total += (int)info[0];
      best = std::max(best, info[4]);
      if (c.verbose) fprintf(stderr, "[bayesvg_b200] pathfinder path %d: %d L-BFGS iterations, best ELBO %.6g at iterate %d (%d pairs)\n", p, (int)info[0], info[4], (int)info[2], (int)info[3]);
    }
  }
  double khat = NAN;
  if (rc == BVG_OK) {
    if (P > 1) {
      khat = pf_psis_weights(lr, w);
      for (int r = 0; r < ns; ++r) {   // inverse-CDF resampling with replacement, Philox stream 8
        uint32_t o[4];
        philox4x32_10((uint32_t)r, 0u, 0u, STREAM_PF_PSIS, (uint32_t)c.seed, (uint32_t)(c.seed >> 32), o);
        const double uu = ((double)(((uint64_t)o[0] << 21) | (o[2] >> 11)) + 0.5) * (1.0 / 9007199254740992.0);
        double cs = 0;
        int pick = S - 1;
        for (int i = 0; i < S; ++i) { cs += w[i]; if (uu < cs) { pick = i; break; } }
        idx[r] = pick;
      }
    } else
      for (int r = 0; r < ns; ++r) idx[r] = r % nd;
    if (cudaMemcpyAsync(d_idx, idx.data(), sizeof(int) * ns, cudaMemcpyHostToDevice, f->stream) != cudaSuccess) rc = BVG_ERR_CUDA;
    if (rc == BVG_OK) {
      k_pf_gather<<<kGrid, 256, 0, f->stream>>>(d_amp, d_idx, ns, G, f->draws);
      f->st.kernel_launches += 1;
      if (cudaStreamSynchronize(f->stream) != cudaSuccess) rc = BVG_ERR_CUDA;
    }
    if (rc == BVG_ERR_CUDA) bvg_set_error("pathfinder: %s", cudaGetErrorString(cudaGetLastError()));
  }
  cudaFree(d_amp); cudaFree(d_idx);
  BVG_TRY(rc);
  f->st.advi_iters = iters_total; f->st.elbo = best; f->st.eta = khat; f->st.converged = 1;
  if (!f->pf_idx) f->pf_idx = new int[8192];
  for (int r = 0; r < ns && r < 8192; ++r) f->pf_idx[r] = idx[r];
  return BVG_OK;
}
