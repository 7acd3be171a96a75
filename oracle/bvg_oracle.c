/*
 * bvg_oracle.c -- fp64 CPU restatement of the bayesVG SVG hot path.  TEST INFRASTRUCTURE ONLY
 * (see bvg_oracle.h: "PARITY UNPINNED").  Each function cites the reference lines it follows;
 * paths are relative to /root/reference.  [U] marks upstream CmdStan/Stan behaviour restated
 * from its published algorithm (not vendored in the reference).
 *
 * Build: make -C oracle   (gcc -O3 -march=x86-64-v3 -fopenmp -shared)
 */
#include "bvg_oracle.h"
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define LOG_TWO_PI 1.8378770664093454835606594728112
#define LOG_PI 1.1447298858494001741434273513531
#define LOG_TWO 0.69314718055994530941723212145818

/* ------------------------------------------------------------------------------------------
 * Parameter layout: Stan declaration order = serialisation order of the unconstrained vector
 * [U].  inst/approxGP.stan:12-23 (gaussian) and inst/approxGP4.stan:12-23 (nb).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int64_t mu_beta0, sigma_beta0, z_beta0, mu_amp, sigma_amp, amp, mu_alpha, sigma_alpha, z_alpha,
      tail; /* tail = sigma_y (gaussian) | phi_nb (nb) */
  int64_t D;
} layout_t;

static layout_t make_layout(int model, int64_t G, int k) {
  layout_t L;
  L.mu_beta0 = 0;
  L.sigma_beta0 = 1;
  L.z_beta0 = 2;
  if (model == BVGO_GAUSSIAN_DEPTH || model == BVGO_NB_DEPTH) {
    /* approxGP2.stan:13-23 / approxGP3.stan:13-23: beta0, beta1, sigma_y | phi_nb, amplitude[G], mu_amplitude,
     * sigma_amplitude, mu_alpha[k], sigma_alpha[k], z_alpha_t[k,G].  mu_beta0 / sigma_beta0 name the slots of
     * beta0 / beta1 (beta1 is unconstrained); there is no z_beta0. */
    L.z_beta0 = -1;
    L.tail = 2;
    L.amp = 3;
    L.mu_amp = 3 + G;
    L.sigma_amp = 4 + G;
    L.mu_alpha = 5 + G;
    L.sigma_alpha = L.mu_alpha + k;
    L.z_alpha = L.sigma_alpha + k;
    L.D = 5 + G + 2 * (int64_t)k + (int64_t)k * G;
    return L;
  }
  if (model == BVGO_GAUSSIAN) { /* approxGP.stan:13-22 */
    L.mu_amp = 2 + G;
    L.sigma_amp = 3 + G;
    L.amp = 4 + G;
    L.mu_alpha = 4 + 2 * G;
    L.sigma_alpha = L.mu_alpha + k;
    L.z_alpha = L.sigma_alpha + k;
    L.tail = L.z_alpha + (int64_t)k * G;
  } else { /* approxGP4.stan:13-22: phi_nb after z_beta0, amplitude before its hyper-parameters */
    L.tail = 2 + G;
    L.amp = 3 + G;
    L.mu_amp = 3 + 2 * G;
    L.sigma_amp = 4 + 2 * G;
    L.mu_alpha = 5 + 2 * G;
    L.sigma_alpha = L.mu_alpha + k;
    L.z_alpha = L.sigma_alpha + k;
  }
  L.D = 5 + 2 * G + 2 * (int64_t)k + (int64_t)k * G;
  return L;
}

int64_t bvgo_dim(int model, int64_t G, int k) { return make_layout(model, G, k).D; }

struct bvgo_ctx {
  int model;
  int64_t M, G;
  int k;
  const double *phi, *y;
  const double *depths; /* gene_depths[G] of approxGP2/3.stan:9 (NULL for the hierarchical-intercept models) */
  int nthreads;
  layout_t L;
  /* model == BVGO_GMM (inst/clusterSVGs.stan): X [N][Dx] row-major, K clusters; L.D = K - 1 + 2 K Dx */
  int64_t gN;
  int gD, gK;
  const double *gX;
};
void bvgo_set_depths(bvgo_ctx *c, const double *gene_depths) { c->depths = gene_depths; }

bvgo_ctx *bvgo_create(int model, int64_t M, int64_t G, int k, const double *phi, const double *y,
                      int nthreads) {
  bvgo_ctx *c = (bvgo_ctx *)calloc(1, sizeof(*c));
  c->model = model;
  c->M = M;
  c->G = G;
  c->k = k;
  c->phi = phi;
  c->y = y;
  c->nthreads = nthreads > 0 ? nthreads : 1;
  c->L = make_layout(model, G, k);
  return c;
}
void bvgo_destroy(bvgo_ctx *c) { free(c); }

/* digamma: recurrence to x >= 10 then the asymptotic series (Abramowitz & Stegun 6.3.18) */
double bvgo_digamma(double x) {
  double r = 0.0;
  while (x < 10.0) {
    r -= 1.0 / x;
    x += 1.0;
  }
  double f = 1.0 / (x * x);
  double t = f * (-1.0 / 12.0 +
                  f * (1.0 / 120.0 +
                       f * (-1.0 / 252.0 +
                            f * (1.0 / 240.0 + f * (-1.0 / 132.0 + f * (691.0 / 32760.0 + f * (-1.0 / 12.0)))))));
  return r + log(x) - 0.5 / x + t;
}

static inline double lgam(double x) { int sg; return lgamma_r(x, &sg); } /* thread-safe */
static inline double log1p_exp(double a) { return a > 0 ? a + log1p(exp(-a)) : log1p(exp(a)); }
static inline double inv_logit(double a) {
  if (a >= 0) return 1.0 / (1.0 + exp(-a));
  double e = exp(a);
  return e / (1.0 + e);
}

/* ------------------------------------------------------------------------------------------
 * log density + analytic gradient.
 *   gaussian: inst/approxGP.stan:25-49 ; nb: inst/approxGP4.stan:25-49
 *   alpha_t   = mu_alpha 1' + diag(sigma_alpha) z_alpha_t         (:31-32)
 *   phi_alpha = phi * alpha_t                                     (:33)
 *   w[i] = phi_alpha[spot_id[i], gene_id[i]] -- identity gather since N = M*G, gene-major
 *          (R/findSpatiallyVariableFeaturesBayes.R:176-183, :290-291)
 *   y ~ normal(beta0[g] + amplitude_sq[g] * w, sigma_y)           (approxGP.stan:48)
 *   y ~ neg_binomial_2_log(beta0[g] + amplitude_sq[g] * w, phi_nb) (approxGP4.stan:48)
 * The only O(N) quantities are R_g = sum_s r, P_g = Phi' r, SSE.
 * ---------------------------------------------------------------------------------------- */
#define SB 256 /* spot block kept in L1 */

static double gmm_log_prob(bvgo_ctx *c, const double *th, int jac, int propto, double *grad);
double bvgo_log_prob(bvgo_ctx *c, const double *th, int jac, int propto, double *grad) {
  if (c->model == BVGO_GMM) return gmm_log_prob(c, th, jac, propto, grad);
  const layout_t L = c->L;
  const int64_t M = c->M, G = c->G;
  const int k = c->k;
  const int nb = c->model == BVGO_NB || c->model == BVGO_NB_DEPTH;
  /* depth-adjusted models (approxGP2.stan:46, approxGP3.stan:46): the per-gene intercept is beta0 + beta1 *
   * gene_depths[g]; in the formulas below (mb, sb, zb) = (beta0, beta1, gene_depths[g]) and beta1 has no
   * constraint, no z_beta0 prior and a normal(0, 2) prior (approxGP2.stan:39) */
  const int dep = c->model == BVGO_GAUSSIAN_DEPTH || c->model == BVGO_NB_DEPTH;
  const double mb = th[L.mu_beta0], lsb = th[L.sigma_beta0], sb = dep ? lsb : exp(lsb);
  const double ma = th[L.mu_amp], lsa = th[L.sigma_amp], sa = exp(lsa);
  const double lt = th[L.tail], tl = exp(lt); /* sigma_y or phi_nb */
  const double *mua = th + L.mu_alpha, *lsal = th + L.sigma_alpha;
  double *sal = (double *)malloc(sizeof(double) * k);
  for (int j = 0; j < k; ++j) sal[j] = exp(lsal[j]);
  const double inv_s2 = nb ? 1.0 : 1.0 / (tl * tl);
  const double lg_phi = nb ? lgam(tl) : 0.0, dg_phi = nb ? bvgo_digamma(tl) : 0.0;

  /* reductions over genes */
  double sse = 0, lik = 0, sumR = 0, sumRz = 0, sdu = 0, sdu2 = 0, sz2 = 0, su = 0, dphi = 0;
  double *APs = (double *)calloc(2 * (size_t)k, sizeof(double)), *APz = APs + k;

#pragma omp parallel num_threads(c->nthreads)
  {
    double *cj = (double *)malloc(sizeof(double) * (4 * (size_t)k + 2 * SB));
    double *al = cj + k, *P = al + k, *lAP = P + k, *f = lAP + k, *r = f + SB;
    double t_sse = 0, t_lik = 0, t_R = 0, t_Rz = 0, t_du = 0, t_du2 = 0, t_z2 = 0, t_u = 0, t_dphi = 0;
    double *t_APs = (double *)calloc(2 * (size_t)k, sizeof(double)), *t_APz = t_APs + k;
#pragma omp for schedule(static)
    for (int64_t g = 0; g < G; ++g) {
      const double zb = dep ? c->depths[g] : th[L.z_beta0 + g], u = th[L.amp + g];
      const double *z = th + L.z_alpha + g * k;
      const double beta = mb + sb * zb; /* approxGP.stan:26 */
      const double A = exp(2.0 * u);    /* amplitude_sq, :27 */
      for (int j = 0; j < k; ++j) {
        al[j] = mua[j] + sal[j] * z[j]; /* :32 */
        cj[j] = A * al[j];
        P[j] = 0;
        t_z2 += z[j] * z[j];
      }
      if (!dep) t_z2 += zb * zb;
      const double *yg = c->y + g * M;
      double Rg = 0, sse_g = 0, lik_g = 0, dphi_g = 0;
      for (int64_t s0 = 0; s0 < M; s0 += SB) {
        const int n = (int)((M - s0) < SB ? (M - s0) : SB);
        for (int i = 0; i < n; ++i) f[i] = beta;
        for (int j = 0; j < k; ++j) {
          const double *pj = c->phi + (int64_t)j * M + s0;
          const double cc = cj[j];
          for (int i = 0; i < n; ++i) f[i] += pj[i] * cc; /* :33 + :48 mean */
        }
        if (!nb) {
          for (int i = 0; i < n; ++i) {
            const double e = yg[s0 + i] - f[i];
            r[i] = e;
            sse_g += e * e;
            Rg += e;
          }
        } else {
          for (int i = 0; i < n; ++i) {
            const double yy = yg[s0 + i], eta = f[i], d = eta - lt;
            const double Lq = log1p_exp(d), sg = inv_logit(d);
            /* neg_binomial_2_log_lpmf [U]: lchoose(y+phi-1,y) + y eta - phi L - y (log phi + L) */
            lik_g += lgam(yy + tl) - lgam(yy + 1.0) - lg_phi + yy * eta - tl * Lq - yy * (lt + Lq);
            const double ri = yy - (yy + tl) * sg;
            r[i] = ri;
            Rg += ri;
            if (grad) dphi_g += bvgo_digamma(yy + tl) - dg_phi - Lq - ri / tl;
          }
        }
        if (grad)
          for (int j = 0; j < k; ++j) {
            const double *pj = c->phi + (int64_t)j * M + s0;
            double a = 0;
            for (int i = 0; i < n; ++i) a += pj[i] * r[i];
            P[j] += a;
          }
      }
      t_sse += sse_g;
      t_lik += lik_g;
      t_dphi += dphi_g;
      Rg *= inv_s2;
      const double du = u - ma;
      t_du += du;
      t_du2 += du * du;
      t_u += u;
      if (grad) {
        double aP = 0;
        for (int j = 0; j < k; ++j) {
          const double Pj = P[j] * inv_s2;
          aP += al[j] * Pj;
          const double AP = A * Pj;
          t_APs[j] += AP;
          t_APz[j] += AP * z[j];
          grad[L.z_alpha + g * k + j] = sal[j] * AP - z[j];
        }
        t_R += Rg;
        t_Rz += Rg * zb;
        if (!dep) grad[L.z_beta0 + g] = sb * Rg - zb;
        grad[L.amp + g] = 2.0 * A * aP - du / (sa * sa) - (jac ? 0.0 : 1.0);
      }
    }
#pragma omp critical
    {
      sse += t_sse; lik += t_lik; sumR += t_R; sumRz += t_Rz; sdu += t_du; sdu2 += t_du2;
      sz2 += t_z2; su += t_u; dphi += t_dphi;
      for (int j = 0; j < k; ++j) { APs[j] += t_APs[j]; APz[j] += t_APz[j]; }
    }
    free(t_APs);
    free(cj);
  }

  const double N = (double)M * (double)G;
  double lp = 0;
  if (!nb) lp += -N * lt - 0.5 * sse * inv_s2; /* normal_lpdf, approxGP.stan:48 */
  else lp += lik;
  lp += -0.5 * sz2;                                              /* :38-39 std_normal on z's */
  if (!dep) lp += -mb * mb / 8.0 - 0.5 * sb * sb - ma * ma / 8.0 - 0.5 * sa * sa; /* :40-43 */
  else lp += -mb * mb / 8.0 - sb * sb / 8.0 - ma * ma / 8.0 - 0.5 * sa * sa;  /* approxGP2.stan:38-39, :42-43 */
  lp += -(double)G * lsa - su - 0.5 * sdu2 / (sa * sa);          /* :44 lognormal */
  for (int j = 0; j < k; ++j) lp += -mua[j] * mua[j] / 8.0 - 0.5 * sal[j] * sal[j]; /* :45-46 */
  if (!nb) lp += dep ? -tl * tl / 8.0 : -0.5 * tl * tl; /* :47 sigma_y ~ std_normal; approxGP2.stan:44 normal(0, 2) */
  else lp += -log1p(tl * tl / 4.0);                              /* approxGP4.stan:41 cauchy(0,2) */
  if (jac) {
    lp += (dep ? 0.0 : lsb) + lsa + su + lt;
    for (int j = 0; j < k; ++j) lp += lsal[j];
  }
  if (!propto) {
    if (!dep) {
      if (!nb)
        lp += -0.5 * LOG_TWO_PI * (N + (double)k * G + 2.0 * G + 2.0 * k + 5.0) - (k + 2.0) * LOG_TWO;
      else /* cauchy: -log(pi*2); T[0,]: -log(1/2) [U] */
        lp += -0.5 * LOG_TWO_PI * ((double)k * G + 2.0 * G + 2.0 * k + 4.0) - (k + 2.0) * LOG_TWO - LOG_PI;
    } else {
      /* approxGP2: N likelihood terms, kG std-normal z, G lognormal, normal(0,2) on beta0, beta1, mu_amplitude,
       * mu_alpha[k], sigma_y (k + 4, each also -log 2), std_normal on sigma_amplitude, sigma_alpha[k] (k + 1);
       * approxGP3: no likelihood / sigma_y normal terms, the truncated cauchy contributes -log(pi) */
      if (!nb)
        lp += -0.5 * LOG_TWO_PI * (N + (double)k * G + (double)G + 2.0 * k + 5.0) - (k + 4.0) * LOG_TWO;
      else
        lp += -0.5 * LOG_TWO_PI * ((double)k * G + (double)G + 2.0 * k + 4.0) - (k + 3.0) * LOG_TWO - LOG_PI;
    }
  }
  if (grad) {
    const double j1 = jac ? 1.0 : 0.0;
    grad[L.mu_beta0] = sumR - mb / 4.0;
    grad[L.sigma_beta0] = dep ? sumRz - sb / 4.0 : sb * (sumRz - sb) + j1;
    grad[L.mu_amp] = sdu / (sa * sa) - ma / 4.0;
    grad[L.sigma_amp] = -(double)G + sdu2 / (sa * sa) - sa * sa + j1;
    for (int j = 0; j < k; ++j) {
      grad[L.mu_alpha + j] = APs[j] - mua[j] / 4.0;
      grad[L.sigma_alpha + j] = sal[j] * (APz[j] - sal[j]) + j1;
    }
    if (!nb) grad[L.tail] = -N + sse * inv_s2 - (dep ? tl * tl / 4.0 : tl * tl) + j1;
    else grad[L.tail] = tl * dphi - 0.5 * tl * tl / (1.0 + tl * tl / 4.0) + j1;
  }
  free(APs);
  free(sal);
  return lp;
}

/* ------------------------------------------------------------------------------------------
 * RNG specification shared with the CUDA engine (DESIGN.md "RNG"): Philox4x32-10 (Salmon et
 * al., SC'11), key = seed, counter = (i_lo, i_hi, t, stream); one standard normal per call via
 * Box-Muller on a 53-bit and a 32-bit uniform.  Stan itself uses a boost generator seeded with
 * random.seed (findSpatiallyVariableFeaturesBayes.R:347,358); bit-identical draws to CmdStan are
 * neither possible nor required.
 * ---------------------------------------------------------------------------------------- */
void bvgo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

double bvgo_normal(uint64_t seed, uint32_t stream, uint32_t t, uint64_t i) {
  uint32_t ctr[4] = {(uint32_t)i, (uint32_t)(i >> 32), t, stream};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)}, o[4];
  bvgo_philox4x32_10(ctr, key, o);
  const double u1 = ((double)(((uint64_t)o[0] << 21) | (o[2] >> 11)) + 0.5) * (1.0 / 9007199254740992.0);
  const double u2 = ((double)o[1] + 0.5) * (1.0 / 4294967296.0);
  return sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925286766559 * u2);
}
enum { STREAM_GRAD = 1, STREAM_ELBO = 2, STREAM_DRAW = 3 };

/* ------------------------------------------------------------------------------------------
 * Meanfield ADVI [U: stan::variational::advi<normal_meanfield>], invoked by
 * R/findSpatiallyVariableFeaturesBayes.R:357-364 (iter = n.iter, draws = n.draws,
 * elbo_samples = elbo.samples; grad_samples = 1, adapt_iter = 50, eval_elbo = 100,
 * tol_rel_obj = 0.01 left at CmdStan defaults).  SURVEY A.4.
 * ---------------------------------------------------------------------------------------- */
double bvgo_elbo(bvgo_ctx *c, const double *mu, const double *om, int S, uint64_t seed, uint32_t t0) {
  const int64_t D = c->L.D;
  double *zeta = (double *)malloc(sizeof(double) * D);
  double acc = 0;
  int ok = 0;
  for (int s = 0; s < S; ++s) {
    for (int64_t i = 0; i < D; ++i) zeta[i] = mu[i] + exp(om[i]) * bvgo_normal(seed, STREAM_ELBO, t0 + s, i);
    const double lp = bvgo_log_prob(c, zeta, 1, 0, NULL); /* log_prob<propto=false, jacobian=true> */
    if (isfinite(lp)) { acc += lp; ++ok; } /* non-finite draws are dropped */
  }
  free(zeta);
  if (ok == 0) return -INFINITY;
  double ent = 0.5 * (double)D * (1.0 + LOG_TWO_PI);
  for (int64_t i = 0; i < D; ++i) ent += om[i];
  /* Stan divides by n_monte_carlo_elbo after re-drawing failed samples; with no failures the two
   * agree.  We divide by the number of finite draws. */
  return acc / ok + ent;
}

int bvgo_advi_step(bvgo_ctx *c, double *mu, double *om, double *hist, double eta, int it,
                   uint64_t seed, uint32_t t) {
  const int64_t D = c->L.D;
  double *zeta = (double *)malloc(sizeof(double) * 3 * D), *eps = zeta + D, *g = eps + D;
  for (int64_t i = 0; i < D; ++i) {
    eps[i] = bvgo_normal(seed, STREAM_GRAD, t, i);
    zeta[i] = mu[i] + exp(om[i]) * eps[i];
  }
  const double lp = bvgo_log_prob(c, zeta, 1, 1, g);
  int finite = isfinite(lp);
  for (int64_t i = 0; i < D && finite; ++i) finite = isfinite(g[i]);
  const double es = eta / sqrt((double)it);
  for (int64_t i = 0; i < D; ++i) {
    /* failing gradient -> zero gradient (adapt_eta's catch) */
    const double gm = finite ? g[i] : 0.0;
    const double go = finite ? g[i] * eps[i] * exp(om[i]) + 1.0 : 0.0;
    double *hm = hist + i, *ho = hist + D + i;
    if (it == 1) { *hm += gm * gm; *ho += go * go; }
    else { *hm = 0.9 * *hm + 0.1 * gm * gm; *ho = 0.9 * *ho + 0.1 * go * go; }
    mu[i] += es * gm / (1.0 + sqrt(*hm));
    om[i] += es * go / (1.0 + sqrt(*ho));
  }
  free(zeta);
  return finite ? 0 : 1;
}

static int cmp_d(const void *a, const void *b) {
  const double x = *(const double *)a, y = *(const double *)b;
  return (x > y) - (x < y);
}

/* Full-rank family [U: stan::variational::normal_fullrank]: q = N(mu, L L^T), L lower triangular, stored PACKED row-major
 * (row i at i (i + 1) / 2).  zeta = mu + L eta; grad mu = g; grad L_ij = g_i eta_j (j <= i) + [i == j] / L_ii (entropy);
 * entropy = D/2 (1 + log 2 pi) + sum log |L_ii|; the step-size sequence acts elementwise on (mu, L) as for meanfield. */
static void fr_transform(int64_t D, const double *mu, const double *Lc, const double *eta, double *zeta) {
  for (int64_t i = 0; i < D; ++i) {
    const double *row = Lc + (size_t)i * (i + 1) / 2;
    double z = 0;
    for (int64_t j = 0; j <= i; ++j) z += row[j] * eta[j];
    zeta[i] = mu[i] + z;
  }
}
double bvgo_fr_elbo(bvgo_ctx *c, const double *mu, const double *Lc, int S, uint64_t seed, uint32_t t0) {
  const int64_t D = c->L.D;
  double *zeta = (double *)malloc(sizeof(double) * 2 * D), *eta = zeta + D;
  double acc = 0;
  int ok = 0;
  for (int s = 0; s < S; ++s) {
    for (int64_t i = 0; i < D; ++i) eta[i] = bvgo_normal(seed, STREAM_ELBO, t0 + s, i);
    fr_transform(D, mu, Lc, eta, zeta);
    const double lp = bvgo_log_prob(c, zeta, 1, 0, NULL);
    if (isfinite(lp)) { acc += lp; ++ok; }
  }
  free(zeta);
  if (ok == 0) return -INFINITY;
  double ent = 0.5 * (double)D * (1.0 + LOG_TWO_PI);
  for (int64_t i = 0; i < D; ++i) {
    const double d = Lc[(size_t)i * (i + 1) / 2 + i];
    if (d != 0.0) ent += log(fabs(d));
  }
  return acc / ok + ent;
}
/* hist = [D for mu | D (D + 1) / 2 for L] */
int bvgo_fr_step(bvgo_ctx *c, double *mu, double *Lc, double *hist, double eta_s, int it, uint64_t seed, uint32_t t) {
  const int64_t D = c->L.D;
  double *zeta = (double *)malloc(sizeof(double) * 3 * D), *eps = zeta + D, *g = eps + D;
  for (int64_t i = 0; i < D; ++i) eps[i] = bvgo_normal(seed, STREAM_GRAD, t, i);
  fr_transform(D, mu, Lc, eps, zeta);
  const double lp = bvgo_log_prob(c, zeta, 1, 1, g);
  int finite = isfinite(lp);
  for (int64_t i = 0; i < D && finite; ++i) finite = isfinite(g[i]);
  const double es = eta_s / sqrt((double)it);
  double *hL = hist + D;
  for (int64_t i = 0; i < D; ++i) {
    const double gm = finite ? g[i] : 0.0;
    double *row = Lc + (size_t)i * (i + 1) / 2, *hrow = hL + (size_t)i * (i + 1) / 2;
    const double dinv = 1.0 / row[i]; /* entropy term uses L BEFORE the update */
    for (int64_t j = 0; j <= i; ++j) {
      const double gl = finite ? g[i] * eps[j] + (j == i ? dinv : 0.0) : 0.0;
      if (it == 1) hrow[j] += gl * gl; else hrow[j] = 0.9 * hrow[j] + 0.1 * gl * gl;
      row[j] += es * gl / (1.0 + sqrt(hrow[j]));
    }
    if (it == 1) hist[i] += gm * gm; else hist[i] = 0.9 * hist[i] + 0.1 * gm * gm;
    mu[i] += es * gm / (1.0 + sqrt(hist[i]));
  }
  free(zeta);
  return finite ? 0 : 1;
}
static void fam_init(int fullrank, int64_t D, const double *th0, double *mu, double *var) {
  memcpy(mu, th0, sizeof(double) * D);
  if (!fullrank) { memset(var, 0, sizeof(double) * D); return; }
  memset(var, 0, sizeof(double) * (size_t)D * (D + 1) / 2);
  for (int64_t i = 0; i < D; ++i) var[(size_t)i * (i + 1) / 2 + i] = 1.0; /* L = I */
}
static int advi_run(bvgo_ctx *c, const double *th0, const bvgo_advi_cfg *cfg, int fullrank, double *mu, double *om,
                    bvgo_advi_result *res, double *trace, int trace_cap) {
  const int64_t D = c->L.D;
  const size_t nh = fullrank ? (size_t)D + (size_t)D * (D + 1) / 2 : 2 * (size_t)D;
  double *hist = (double *)calloc(nh, sizeof(double));
#define FAM_STEP(eta_, it_, t_) (fullrank ? bvgo_fr_step(c, mu, om, hist, eta_, it_, cfg->seed, t_) : bvgo_advi_step(c, mu, om, hist, eta_, it_, cfg->seed, t_))
#define FAM_ELBO(t_) (fullrank ? bvgo_fr_elbo(c, mu, om, cfg->elbo_samples, cfg->seed, t_) : bvgo_elbo(c, mu, om, cfg->elbo_samples, cfg->seed, t_))
  uint32_t tg = 0, te = 0; /* running draw counters */
  int gevals = 0, eevals = 0;
  double eta = cfg->eta;
  memset(res, 0, sizeof(*res));
  if (eta <= 0) { /* adapt_eta */
    static const double seq[5] = {100, 10, 1, 0.1, 0.01};
    fam_init(fullrank, D, th0, mu, om);
    const double e0 = FAM_ELBO(te);
    te += cfg->elbo_samples; ++eevals;
    if (!isfinite(e0)) { free(hist); return -2; }
    double ebest = -INFINITY, e = -INFINITY, eta_best = 0;
    for (int q = 0;; ++q) {
      fam_init(fullrank, D, th0, mu, om);
      for (int it = 1; it <= cfg->adapt_iter; ++it) {
        FAM_STEP(seq[q], it, tg++);
        ++gevals;
      }
      e = FAM_ELBO(te);
      te += cfg->elbo_samples; ++eevals;
      if (isnan(e)) e = -INFINITY;
      if (e < ebest && ebest > e0) break; /* previous eta stands */
      if (q < 4) { ebest = e; eta_best = seq[q]; }
      else {
        if (e > e0) { ebest = e; eta_best = seq[q]; break; }
        free(hist);
        return -3; /* "All proposed step-sizes failed" */
      }
      memset(hist, 0, sizeof(double) * nh);
    }
    eta = eta_best;
  }
  /* stochastic_gradient_ascent */
  fam_init(fullrank, D, th0, mu, om);
  memset(hist, 0, sizeof(double) * nh);
  int cbn = (int)fmax(0.1 * cfg->iter / cfg->eval_elbo, 2.0);
  double *cb = (double *)malloc(sizeof(double) * 2 * cbn), *tmp = cb + cbn;
  int cbc = 0, cbh = 0, ntr = 0, conv = 0, it;
  double elbo = 0, elbo_prev;
  for (it = 1; it <= cfg->iter; ++it) {
    if (FAM_STEP(eta, it, tg++)) {
      /* main loop: a failing gradient is an error in Stan; report it */
      free(cb); free(hist);
      res->eta = eta; res->iters = it; res->grad_evals = gevals + it;
      return -4;
    }
    if (it % cfg->eval_elbo == 0) {
      elbo_prev = elbo;
      elbo = FAM_ELBO(te);
      te += cfg->elbo_samples; ++eevals;
      const double d = fabs((elbo - elbo_prev) / elbo);
      cb[cbh] = d; cbh = (cbh + 1) % cbn; if (cbc < cbn) ++cbc;
      double mean = 0;
      for (int i = 0; i < cbc; ++i) { mean += cb[i]; tmp[i] = cb[i]; }
      mean /= cbc;
      qsort(tmp, cbc, sizeof(double), cmp_d);
      /* Stan's circ_buff_median: nth_element at size/2 */
      const double med = tmp[cbc / 2];
      if (trace && ntr < trace_cap) {
        trace[4 * ntr + 0] = it; trace[4 * ntr + 1] = elbo; trace[4 * ntr + 2] = mean; trace[4 * ntr + 3] = med;
        ++ntr;
      }
      if (mean < cfg->tol_rel_obj) conv |= 1;
      if (med < cfg->tol_rel_obj) conv |= 2;
      if (conv) break;
    }
  }
  if (it > cfg->iter) it = cfg->iter;
  res->eta = eta; res->iters = it; res->grad_evals = gevals + it; res->elbo_evals = eevals;
  res->converged = conv; res->elbo = elbo; res->n_trace = ntr;
  free(cb);
  free(hist);
  return 0;
}
#undef FAM_STEP
#undef FAM_ELBO
int bvgo_advi(bvgo_ctx *c, const double *th0, const bvgo_advi_cfg *cfg, double *mu, double *om,
              bvgo_advi_result *res, double *trace, int trace_cap) {
  return advi_run(c, th0, cfg, 0, mu, om, res, trace, trace_cap);
}
/* algorithm = "fullrank" (:357-364 with algorithm = "fullrank"): Lc receives the packed Cholesky factor */
int bvgo_advi_fullrank(bvgo_ctx *c, const double *th0, const bvgo_advi_cfg *cfg, double *mu, double *Lc,
                       bvgo_advi_result *res, double *trace, int trace_cap) {
  return advi_run(c, th0, cfg, 1, mu, Lc, res, trace, trace_cap);
}

/* ------------------------------------------------------------------------------------------
 * L-BFGS [U: stan::optimization::BFGSMinimizer<LBFGSUpdate>], invoked by
 * R/findSpatiallyVariableFeaturesBayes.R:346-353 (init = 0, jacobian = FALSE, iter = mle.iter,
 * history_size = 25).  Minimises f = -log_prob<propto=false, jacobian=false>.
 * ---------------------------------------------------------------------------------------- */
typedef struct { bvgo_ctx *c; int64_t D; int fevals; int jac, propto; } fobj;
static int feval(fobj *o, const double *x, double *f, double *g) {
  const double lp = bvgo_log_prob(o->c, x, o->jac, o->propto, g);
  o->fevals++;
  *f = -lp;
  int ok = isfinite(lp);
  for (int64_t i = 0; i < o->D; ++i) { g[i] = -g[i]; if (!isfinite(g[i])) ok = 0; }
  return ok ? 0 : 1;
}
static double dot(const double *a, const double *b, int64_t n) {
  double s = 0;
  for (int64_t i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}
/* minimiser on [lo,hi] of the cubic through (0,0) slope df0 and (x1,f1) slope df1 */
static double cubic_interp(double df0, double x1, double f1, double df1, double lo, double hi) {
  const double c3 = (-12 * f1 + 6 * x1 * (df0 + df1)) / (x1 * x1 * x1);
  const double c2 = -(4 * df0 + 2 * df1) / x1 + 6 * f1 / (x1 * x1);
  const double c1 = df0;
  const double ts = sqrt(c2 * c2 - 2.0 * c1 * c3);
  const double s1 = -(c2 + ts) / c3, s2 = -(c2 - ts) / c3;
#define CUB(x) ((x) * ((x) * ((x) * c3 / 3.0 + c2) / 2.0 + c1))
  double minx = lo, minf = CUB(lo), t = CUB(hi);
  if (t < minf) { minf = t; minx = hi; }
  if (lo < s1 && s1 < hi) { t = CUB(s1); if (t < minf) { minf = t; minx = s1; } }
  if (lo < s2 && s2 < hi) { t = CUB(s2); if (t < minf) { minf = t; minx = s2; } }
#undef CUB
  return minx;
}
static int ls_zoom(fobj *o, double *alpha, double *x1, double *f1, double *g1, const double *x0,
                   double f0, double c1dfp, double c2dfp, const double *p, double alo, double aloF,
                   double aloD, double ahi, double ahiF, double ahiD) {
  const int64_t D = o->D;
  for (int itn = 1;; ++itn) {
    if (fabs(alo - ahi) < 1e-16) return 1;
    double a;
    if (itn % 5 == 0) a = 0.5 * (alo + ahi);
    else {
      const double d1 = aloD + ahiD - 3 * (aloF - ahiF) / (alo - ahi);
      double d2 = sqrt(d1 * d1 - aloD * ahiD);
      if (ahi < alo) d2 = -d2;
      a = ahi - (ahi - alo) * (ahiD + d2 - d1) / (ahiD - aloD + 2 * d2);
      const double mn = fmin(alo, ahi), mx = fmax(alo, ahi), w = fabs(alo - ahi);
      if (!isfinite(a) || a < mn + 0.01 * w || a > mx - 0.01 * w) a = 0.5 * (alo + ahi);
    }
    for (int64_t i = 0; i < D; ++i) x1[i] = x0[i] + a * p[i];
    while (feval(o, x1, f1, g1)) {
      a = 0.5 * (a + fmin(alo, ahi));
      if (fabs(fmin(alo, ahi) - a) < 1e-16) return 1;
      for (int64_t i = 0; i < D; ++i) x1[i] = x0[i] + a * p[i];
    }
    const double nd = dot(g1, p, D);
    *alpha = a;
    if (*f1 > f0 + a * c1dfp || *f1 >= aloF) { ahi = a; ahiF = *f1; ahiD = nd; }
    else {
      if (fabs(nd) <= -c2dfp) return 0;
      if (nd * (ahi - alo) >= 0) { ahi = alo; ahiF = aloF; ahiD = aloD; }
      alo = a; aloF = *f1; aloD = nd;
    }
  }
}
static int wolfe_ls(fobj *o, double *alpha, double *x1, double *f1, double *g1, const double *p,
                    const double *x0, double f0, const double *g0) {
  const int64_t D = o->D;
  const double c1 = 1e-4, c2 = 0.9, min_alpha = 1e-12;
  const int max_its = 20, max_restarts = 10;
  const double dfp = dot(g0, p, D), c1dfp = c1 * dfp, c2dfp = c2 * dfp;
  double a0 = min_alpha, prevF = f0, prevD = dfp;
  int nits = 0, restarts = 0;
  for (;;) {
    if (nits >= max_its) return 1;
    for (int64_t i = 0; i < D; ++i) x1[i] = x0[i] + *alpha * p[i];
    if (feval(o, x1, f1, g1)) {
      if (restarts >= max_restarts) return 1;
      *alpha = 0.5 * (a0 + *alpha);
      ++restarts;
      continue;
    }
    restarts = 0;
    const double nd = dot(g1, p, D);
    if (*f1 > f0 + *alpha * c1dfp || (*f1 >= prevF && nits > 0))
      return ls_zoom(o, alpha, x1, f1, g1, x0, f0, c1dfp, c2dfp, p, a0, prevF, prevD, *alpha, *f1, nd);
    if (fabs(nd) <= -c2dfp) return 0;
    if (nd >= 0)
      return ls_zoom(o, alpha, x1, f1, g1, x0, f0, c1dfp, c2dfp, p, *alpha, *f1, nd, a0, prevF, prevD);
    a0 = *alpha; prevF = *f1; prevD = nd;
    *alpha *= 10.0;
    ++nits;
  }
}

typedef struct { double v; int64_t i; } rank_t;
static double summarise_gene(double *x, double *d, int nd, int64_t G, int64_t g, double *out);
static void fill_ranks(rank_t *rk, int64_t G, double *out);
/* cb (optional): called with the initial point (it = 0) and after every accepted iterate; x, g = grad of f = -lp */
typedef void (*lbfgs_cb)(void *user, int it, const double *x, const double *g, double f);
static int lbfgs_core(bvgo_ctx *c, const double *th0, const bvgo_lbfgs_cfg *cfg, int jac, int propto, double *out,
                      bvgo_lbfgs_result *res, lbfgs_cb cb, void *user) {
  const int64_t D = c->L.D;
  const int m = cfg->history;
  fobj o = {c, D, 0, jac, propto};
  double *buf = (double *)malloc(sizeof(double) * D * (size_t)(8 + 2 * m));
  double *xk = buf, *gk = xk + D, *pk = gk + D, *xk1 = pk + D, *gk1 = xk1 + D, *pk1 = gk1 + D,
         *sk = pk1 + D, *yk = sk + D, *Sh = yk + D, *Yh = Sh + (size_t)m * D;
  double *rho = (double *)malloc(sizeof(double) * 2 * m), *alph = rho + m;
  int nh = 0, head = 0; /* circular history */
  double fk, fk1 = 0, gamma = 1, alpha = cfg->init_alpha, alphak1 = 0;
  memcpy(xk, th0, sizeof(double) * D);
  int term = 0, it = 0;
  if (feval(&o, xk, &fk, gk)) { free(buf); free(rho); return -1; }
  if (cb) cb(user, 0, xk, gk, fk);
  for (int64_t i = 0; i < D; ++i) pk[i] = -gk[i];
  const double eps = DBL_EPSILON;
  while (!term) {
    ++it;
    int resetB = (it == 1) ? 1 : 0;
    for (;;) {
      if (resetB) for (int64_t i = 0; i < D; ++i) pk[i] = -gk[i];
      if (it > 1 && resetB != 2)
        alpha = fmin(1.0, 1.01 * cubic_interp(dot(gk1, pk1, D), alphak1, fk - fk1, dot(gk, pk, D), 1e-12, 1.0));
      else alpha = cfg->init_alpha;
      const int rc = wolfe_ls(&o, &alpha, xk1, &fk1, gk1, pk, xk, fk, gk);
      if (rc) {
        if (resetB) { term = -1; break; } /* TERM_LSFAIL */
        resetB = 2;
        continue;
      }
      break;
    }
    if (term) break;
    /* swap so that k is the newest iterate */
    { double t = fk; fk = fk1; fk1 = t; }
    { double *t; t = xk; xk = xk1; xk1 = t; t = gk; gk = gk1; gk1 = t; t = pk; pk = pk1; pk1 = t; }
    for (int64_t i = 0; i < D; ++i) { sk[i] = xk[i] - xk1[i]; yk[i] = gk[i] - gk1[i]; }
    if (cb) cb(user, it, xk, gk, fk);
    const double gnorm = sqrt(dot(gk, gk, D)), snorm = sqrt(dot(sk, sk, D));
    const double skyk = dot(sk, yk, D), ykyk = dot(yk, yk, D);
    if (resetB) {
      const double B0 = ykyk / skyk;
      nh = 0; head = 0;
      for (int64_t i = 0; i < D; ++i) pk1[i] /= B0;
      alphak1 = alpha * B0;
    } else alphak1 = alpha;
    gamma = skyk / ykyk;
    memcpy(Sh + (size_t)head * D, sk, sizeof(double) * D);
    memcpy(Yh + (size_t)head * D, yk, sizeof(double) * D);
    rho[head] = 1.0 / skyk;
    head = (head + 1) % m;
    if (nh < m) ++nh;
    /* convergence tests */
    if (fabs(fk1 - fk) < cfg->tol_obj) term = 1;
    else if (fabs(fk1 - fk) < cfg->tol_rel_obj * eps * fmax(fabs(fk1), fmax(fabs(fk), 1.0))) term = 2;
    else if (gnorm < cfg->tol_grad) term = 3;
    else if (snorm < cfg->tol_param) term = 4;
    else if (it >= cfg->max_iter) term = 5;
    /* new direction: two-loop recursion, pk = -H gk */
    for (int64_t i = 0; i < D; ++i) pk[i] = -gk[i];
    for (int q = 0; q < nh; ++q) {
      const int h = (head - 1 - q + 2 * m) % m;
      alph[h] = rho[h] * dot(Sh + (size_t)h * D, pk, D);
      const double a = alph[h];
      const double *Y = Yh + (size_t)h * D;
      for (int64_t i = 0; i < D; ++i) pk[i] -= a * Y[i];
    }
    for (int64_t i = 0; i < D; ++i) pk[i] *= gamma;
    for (int q = nh - 1; q >= 0; --q) {
      const int h = (head - 1 - q + 2 * m) % m;
      const double b = rho[h] * dot(Yh + (size_t)h * D, pk, D);
      const double *S = Sh + (size_t)h * D;
      const double a = alph[h] - b;
      for (int64_t i = 0; i < D; ++i) pk[i] += a * S[i];
    }
    if (!term) { /* relative gradient test: g' H g / max(|f|, 1) */
      const double rg = -dot(gk, pk, D) / fmax(fabs(fk), 1.0);
      if (rg < cfg->tol_rel_grad * eps) term = 6;
    }
  }
  memcpy(out, xk, sizeof(double) * D);
  res->iters = it; res->fevals = o.fevals; res->term = term; res->lp = -fk;
  free(buf);
  free(rho);
  return 0;
}
int bvgo_lbfgs(bvgo_ctx *c, const double *th0, const bvgo_lbfgs_cfg *cfg, double *out, bvgo_lbfgs_result *res) {
  return lbfgs_core(c, th0, cfg, 0, 0, out, res, NULL, NULL);
}

/* ------------------------------------------------------------------------------------------
 * Pathfinder [U: stan::services::pathfinder::pathfinder_lbfgs_single / _multi; Zhang, Carpenter, Gelman, Vehtari 2022],
 * invoked by R/findSpatiallyVariableFeaturesBayes.R:366-375: init = 0.01, num_paths = n.cores, max_lbfgs_iters = 200,
 * num_elbo_draws = elbo.samples (50), history_size = 25, draws = n.draws (PSIS draws); single-path draws 1000 [U].
 * Along the L-BFGS path of f = -log_prob<propto, jacobian = true>, every iterate (x, g) gets a Gaussian approximation
 * N(xc, Sigma) from the last <= J curvature pairs (S, Y) that pass check_curve and the recursively updated diagonal
 * alpha (form_diag):  Sigma = diag(alpha) + [diag(alpha) Y | C'] [[0, I], [I, Y' diag(alpha) Y + Dk]] [...]',
 * C = -R^{-1} S' (R = upper triangle of S'Y, Dk = its diagonal), xc = x - Sigma g.  Sampling goes through the thin
 * factorisation W' = Q Rk of W = [Y diag(sqrt alpha); C diag(1 / sqrt alpha)] and La = chol(Rk Mk Rk' + I):
 * phi = xc + sqrt(alpha) (Q La Q'u + u - Q Q'u), log q = -log|La| - 1/2 sum log alpha - 1/2 (u'u + D log 2 pi).
 * The iterate with the largest ELBO estimate (num_elbo_draws draws) supplies the path's draws; several paths are
 * combined by Pareto-smoothed importance resampling on lp - log q.
 * Rk comes from a Cholesky factorisation of W W' here (Stan: Householder QR of W'): same Q up to column signs.
 * ---------------------------------------------------------------------------------------- */
double bvgo_uniform(uint64_t seed, uint32_t stream, uint32_t t, uint64_t i) {
  uint32_t ctr[4] = {(uint32_t)i, (uint32_t)(i >> 32), t, stream};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)}, o[4];
  bvgo_philox4x32_10(ctr, key, o);
  return ((double)(((uint64_t)o[0] << 21) | (o[2] >> 11)) + 0.5) * (1.0 / 9007199254740992.0);
}
enum { STREAM_PF_INIT = 5, STREAM_PF_ELBO = 6, STREAM_PF_DRAW = 7, STREAM_PF_PSIS = 8 };
static int chol_lower(double *A, int n) { /* in place, row-major; returns 1 if not positive definite */
  for (int j = 0; j < n; ++j) {
    double d = A[j * n + j];
    for (int q = 0; q < j; ++q) d -= A[j * n + q] * A[j * n + q];
    if (!(d > 0.0) || !isfinite(d)) return 1;
    d = sqrt(d);
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
typedef struct {
  int n;                 /* curvature pairs in use */
  double *W;             /* [2n][D] */
  double *xc, *alpha;    /* [D] */
  double Lc[50 * 50];    /* lower Cholesky factor of W W' (Rk = Lc') */
  double La[50 * 50];    /* chol(Rk Mk Rk' + I), lower */
  double logdet;         /* sum log La_ii + 1/2 sum log alpha */
  int ok;
} pf_approx;
typedef struct {
  bvgo_ctx *c; int64_t D; const bvgo_pf_cfg *cfg; int path;
  double *alpha, *Sh, *Yh, *xprev, *gprev, *work; int nh, head;
  pf_approx cur, best; double best_elbo; int best_iter; int n_iter;
  double *elbos; int elbo_cap;
} pf_state;
static void pf_build(pf_state *st, const double *x, const double *g) {
  const int64_t D = st->D;
  const int J = st->cfg->history, n = st->nh, n2 = 2 * n;
  pf_approx *a = &st->cur;
  a->n = n; a->ok = 0;
  memcpy(a->alpha, st->alpha, sizeof(double) * D);
  const double *al = st->alpha;
  /* history rows, oldest first */
  const double *S[25], *Y[25];
  for (int q = 0; q < n; ++q) { const int h = (st->head - n + q + 2 * J) % J; S[q] = st->Sh + (size_t)h * D; Y[q] = st->Yh + (size_t)h * D; }
  double SY[25 * 25], YaY[25 * 25], Rinv[25 * 25];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      double sy = 0, yy = 0;
      for (int64_t d = 0; d < D; ++d) { sy += S[i][d] * Y[j][d]; yy += Y[i][d] * al[d] * Y[j][d]; }
      SY[i * n + j] = sy; YaY[i * n + j] = yy;
    }
  /* Rinv: inverse of the upper triangle of SY (back substitution, column by column) */
  for (int col = 0; col < n; ++col)
    for (int i = n - 1; i >= 0; --i) {
      double v = i == col ? 1.0 : 0.0;
      for (int q = i + 1; q < n; ++q) v -= SY[i * n + q] * Rinv[q * n + col];
      Rinv[i * n + col] = i > col ? 0.0 : v / SY[i * n + i];
    }
  /* W = [Y sqrt(alpha); C / sqrt(alpha)], C = -Rinv S */
  double *W = a->W;
  for (int i = 0; i < n; ++i)
    for (int64_t d = 0; d < D; ++d) {
      W[(size_t)i * D + d] = Y[i][d] * sqrt(al[d]);
      double cv = 0;
      for (int q = i; q < n; ++q) cv -= Rinv[i * n + q] * S[q][d];
      W[(size_t)(n + i) * D + d] = cv / sqrt(al[d]);
    }
  /* Gram, Rk, Mk, La */
  double *G = a->Lc, Mk[50 * 50], T[50 * 50], H[50 * 50];
  for (int i = 0; i < n2; ++i)
    for (int j = 0; j <= i; ++j) {
      double v = 0;
      for (int64_t d = 0; d < D; ++d) v += W[(size_t)i * D + d] * W[(size_t)j * D + d];
      G[i * n2 + j] = G[j * n2 + i] = v;
    }
  if (n2 > 0 && chol_lower(G, n2)) return; /* linearly dependent history: no approximation at this iterate */
  memset(Mk, 0, sizeof(Mk));
  for (int i = 0; i < n; ++i) {
    Mk[i * n2 + n + i] = Mk[(n + i) * n2 + i] = 1.0;
    for (int j = 0; j < n; ++j) Mk[(n + i) * n2 + n + j] = YaY[i * n + j] + (i == j ? SY[i * n + i] : 0.0);
  }
  /* H = Rk Mk Rk' + I with Rk = Lc' :  H = Lc' Mk Lc + I */
  for (int i = 0; i < n2; ++i)
    for (int j = 0; j < n2; ++j) {
      double v = 0;
      for (int q = 0; q < n2; ++q) v += Mk[i * n2 + q] * G[q * n2 + j]; /* (Mk Lc)_ij */
      T[i * n2 + j] = v;
    }
  for (int i = 0; i < n2; ++i)
    for (int j = 0; j < n2; ++j) {
      double v = i == j ? 1.0 : 0.0;
      for (int q = 0; q < n2; ++q) v += G[q * n2 + i] * T[q * n2 + j]; /* (Lc' T)_ij */
      H[i * n2 + j] = v;
    }
  for (int i = 0; i < n2; ++i) for (int j = 0; j < i; ++j) { const double v = 0.5 * (H[i * n2 + j] + H[j * n2 + i]); H[i * n2 + j] = H[j * n2 + i] = v; }
  memcpy(a->La, H, sizeof(double) * n2 * n2);
  if (n2 > 0 && chol_lower(a->La, n2)) return;
  a->logdet = 0;
  for (int i = 0; i < n2; ++i) a->logdet += log(a->La[i * n2 + i]);
  double sl = 0;
  for (int64_t d = 0; d < D; ++d) sl += log(al[d]);
  a->logdet += 0.5 * sl;
  /* xc = x - Sigma g:  t1 = C g, t2 = Y (alpha g), v = t2 + (YaY + Dk) t1; Sigma g = alpha g + alpha (Y' t1) + C' v */
  double t1[25], t2[25], v[25];
  for (int i = 0; i < n; ++i) {
    double a1 = 0, a2 = 0;
    for (int64_t d = 0; d < D; ++d) { a1 += W[(size_t)(n + i) * D + d] * sqrt(al[d]) * g[d]; a2 += Y[i][d] * al[d] * g[d]; }
    t1[i] = a1; t2[i] = a2;
  }
  for (int i = 0; i < n; ++i) {
    double s = t2[i];
    for (int j = 0; j < n; ++j) s += (YaY[i * n + j] + (i == j ? SY[i * n + i] : 0.0)) * t1[j];
    v[i] = s;
  }
  for (int64_t d = 0; d < D; ++d) {
    double yt = 0, cv = 0;
    for (int i = 0; i < n; ++i) { yt += Y[i][d] * t1[i]; cv += W[(size_t)(n + i) * D + d] * sqrt(al[d]) * v[i]; }
    a->xc[d] = x[d] - (al[d] * g[d] + al[d] * yt + cv);
  }
  a->ok = 1;
}
/* one draw from an approximation: u given (length D) -> phi, returns log q */
static double pf_draw(const pf_approx *a, int64_t D, const double *u, double *phi) {
  const int n2 = 2 * a->n;
  double wu[50], u1[50], z[50], cc[50];
  double uu = 0;
  for (int64_t d = 0; d < D; ++d) uu += u[d] * u[d];
  for (int i = 0; i < n2; ++i) {
    double s = 0;
    for (int64_t d = 0; d < D; ++d) s += a->W[(size_t)i * D + d] * u[d];
    wu[i] = s;
  }
  /* u1 = Rk^{-T} (W u) = Lc^{-1} (W u) */
  for (int i = 0; i < n2; ++i) {
    double s = wu[i];
    for (int q = 0; q < i; ++q) s -= a->Lc[i * n2 + q] * u1[q];
    u1[i] = s / a->Lc[i * n2 + i];
  }
  for (int i = 0; i < n2; ++i) { /* z = La u1 - u1 */
    double s = -u1[i];
    for (int q = 0; q <= i; ++q) s += a->La[i * n2 + q] * u1[q];
    z[i] = s;
  }
  /* cc = Rk^{-1} z = Lc^{-T} z */
  for (int i = n2 - 1; i >= 0; --i) {
    double s = z[i];
    for (int q = i + 1; q < n2; ++q) s -= a->Lc[q * n2 + i] * cc[q];
    cc[i] = s / a->Lc[i * n2 + i];
  }
  for (int64_t d = 0; d < D; ++d) {
    double s = u[d];
    for (int i = 0; i < n2; ++i) s += a->W[(size_t)i * D + d] * cc[i];
    phi[d] = a->xc[d] + sqrt(a->alpha[d]) * s;
  }
  return -a->logdet - 0.5 * (uu + (double)D * LOG_TWO_PI);
}
static void pf_copy(pf_approx *dst, const pf_approx *src, int64_t D) {
  dst->n = src->n; dst->logdet = src->logdet; dst->ok = src->ok;
  memcpy(dst->W, src->W, sizeof(double) * 2 * (size_t)src->n * D);
  memcpy(dst->xc, src->xc, sizeof(double) * D);
  memcpy(dst->alpha, src->alpha, sizeof(double) * D);
  memcpy(dst->Lc, src->Lc, sizeof(src->Lc));
  memcpy(dst->La, src->La, sizeof(src->La));
}
static void pf_cb(void *user, int it, const double *x, const double *g, double f) {
  pf_state *st = (pf_state *)user;
  const int64_t D = st->D;
  const int J = st->cfg->history;
  (void)f;
  if (it > 0) {
    double *s = st->work, *y = s + D;
    double ys = 0, yy = 0;
    for (int64_t d = 0; d < D; ++d) { s[d] = x[d] - st->xprev[d]; y[d] = g[d] - st->gprev[d]; ys += y[d] * s[d]; yy += y[d] * y[d]; }
    if (ys > 0 && fabs(yy / ys) <= 1e12) { /* check_curve */
      double yay = 0, sias = 0;
      for (int64_t d = 0; d < D; ++d) { yay += y[d] * st->alpha[d] * y[d]; sias += s[d] * s[d] / st->alpha[d]; }
      for (int64_t d = 0; d < D; ++d) { /* form_diag */
        const double a = st->alpha[d], sa = s[d] / a;
        st->alpha[d] = ys / (yay / a + y[d] * y[d] - (yay / sias) * sa * sa);
      }
      memcpy(st->Sh + (size_t)st->head * D, s, sizeof(double) * D);
      memcpy(st->Yh + (size_t)st->head * D, y, sizeof(double) * D);
      st->head = (st->head + 1) % J;
      if (st->nh < J) ++st->nh;
    }
    pf_build(st, x, g);
    double elbo = -INFINITY;
    if (st->cur.ok) {
      const int K = st->cfg->num_elbo_draws;
      double *u = st->work, *phi = u + D, acc = 0;
      int fin = 1;
      for (int k = 0; k < K && fin; ++k) {
        const uint32_t t = ((uint32_t)st->path << 24) | ((uint32_t)it << 12) | (uint32_t)k;
        for (int64_t d = 0; d < D; ++d) u[d] = bvgo_normal(st->cfg->seed, STREAM_PF_ELBO, t, d);
        const double lq = pf_draw(&st->cur, D, u, phi);
        const double lp = bvgo_log_prob(st->c, phi, 1, 1, NULL);
        if (!isfinite(lp) || !isfinite(lq)) fin = 0; else acc += lp - lq;
      }
      if (fin) elbo = acc / K;
    }
    if (it < st->elbo_cap) st->elbos[it] = elbo;
    if (elbo > st->best_elbo) { st->best_elbo = elbo; st->best_iter = it; pf_copy(&st->best, &st->cur, D); }
    st->n_iter = it;
  }
  memcpy(st->xprev, x, sizeof(double) * D);
  memcpy(st->gprev, g, sizeof(double) * D);
}
static void pf_alloc(pf_approx *a, int64_t D, int J);
/* test hook: the approximation built from given (alpha, S, Y [n][D] oldest first, x, g); nu draws u [nu][D] -> phi, log q */
int bvgo_pf_approx_test(int64_t D, int n, const double *alpha, const double *S, const double *Y, const double *x,
                        const double *g, int nu, const double *u, double *phi, double *logq, double *xc) {
  bvgo_pf_cfg cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.history = n > 0 ? n : 1;
  pf_state st;
  memset(&st, 0, sizeof(st));
  st.D = D; st.cfg = &cfg; st.nh = n; st.head = n % cfg.history;
  st.alpha = (double *)alpha; st.Sh = (double *)S; st.Yh = (double *)Y;
  pf_alloc(&st.cur, D, cfg.history);
  pf_build(&st, x, g);
  const int ok = st.cur.ok;
  if (ok) {
    memcpy(xc, st.cur.xc, sizeof(double) * D);
    for (int k = 0; k < nu; ++k) logq[k] = pf_draw(&st.cur, D, u + (size_t)k * D, phi + (size_t)k * D);
  }
  free(st.cur.W); free(st.cur.xc);
  return ok ? 0 : 1;
}
static void pf_alloc(pf_approx *a, int64_t D, int J) {
  a->W = (double *)malloc(sizeof(double) * 2 * (size_t)J * D);
  a->xc = (double *)malloc(sizeof(double) * 2 * D);
  a->alpha = a->xc + D;
  a->n = 0; a->ok = 0;
}
static double *pf_phi_out = NULL;  /* bvgo_pf_single_phi: the unconstrained draws of the path, [num_draws][D] */
/* one path: amp [num_draws][G] (exp of the amplitude coordinates), lp_ratio [num_draws] = lp - log q;
 * elbos (optional) [max_lbfgs_iters + 1]: the ELBO estimate at every iterate */
int bvgo_pf_single(bvgo_ctx *c, const bvgo_pf_cfg *cfg, int path, double *amp, double *lp_ratio, bvgo_pf_info *info,
                   double *elbos) {
  const int64_t D = c->L.D, G = c->G;
  const int J = cfg->history;
  if (J < 1 || J > 25) return -1;
  pf_state st;
  memset(&st, 0, sizeof(st));
  st.c = c; st.D = D; st.cfg = cfg; st.path = path;
  st.alpha = (double *)malloc(sizeof(double) * ((size_t)D * (5 + 2 * (size_t)J)));
  st.xprev = st.alpha + D; st.gprev = st.xprev + D; st.work = st.gprev + D; st.Sh = st.work + 2 * D; st.Yh = st.Sh + (size_t)J * D;
  for (int64_t d = 0; d < D; ++d) st.alpha[d] = 1.0;
  pf_alloc(&st.cur, D, J); pf_alloc(&st.best, D, J);
  st.best_elbo = -INFINITY; st.best_iter = -1;
  st.elbo_cap = cfg->max_lbfgs_iters + 1;
  st.elbos = (double *)malloc(sizeof(double) * st.elbo_cap);
  for (int i = 0; i < st.elbo_cap; ++i) st.elbos[i] = NAN;
  double *th0 = (double *)malloc(sizeof(double) * 2 * D), *out = th0 + D;
  for (int64_t d = 0; d < D; ++d) th0[d] = cfg->init_radius * (2.0 * bvgo_uniform(cfg->seed, STREAM_PF_INIT, (uint32_t)path, d) - 1.0);
  bvgo_lbfgs_cfg lc = {cfg->max_lbfgs_iters, J, 1e-3, 1e-12, 1e4, 1e-8, 1e7, 1e-8};
  bvgo_lbfgs_result lr;
  memset(&lr, 0, sizeof(lr));
  int rc = lbfgs_core(c, th0, &lc, 1, 1, out, &lr, pf_cb, &st);
  if (rc == 0 && st.best_iter < 0) rc = -5; /* no iterate produced a finite ELBO */
  if (rc == 0) {
    double *u = st.work, *phi = u + D;
    for (int k = 0; k < cfg->num_draws; ++k) {
      const uint32_t t = ((uint32_t)path << 16) | (uint32_t)k;
      for (int64_t d = 0; d < D; ++d) u[d] = bvgo_normal(cfg->seed, STREAM_PF_DRAW, t, d);
      const double lq = pf_draw(&st.best, D, u, phi);
      const double lp = bvgo_log_prob(c, phi, 1, 1, NULL);
      lp_ratio[k] = isfinite(lp) ? lp - lq : -INFINITY;
      if (pf_phi_out) memcpy(pf_phi_out + (size_t)k * D, phi, sizeof(double) * D);
      for (int64_t g = 0; g < G; ++g) amp[(size_t)k * G + g] = exp(phi[c->L.amp + g]);
    }
  }
  if (info) { info->lbfgs_iters = lr.iters; info->lbfgs_term = lr.term; info->best_iter = st.best_iter; info->best_elbo = st.best_elbo; info->n_pairs = st.best.n; }
  if (elbos) memcpy(elbos, st.elbos, sizeof(double) * st.elbo_cap);
  free(th0); free(st.elbos); free(st.cur.W); free(st.cur.xc); free(st.best.W); free(st.best.xc); free(st.alpha);
  return rc;
}
/* the same path with its draws in the unconstrained space (the mixture program of clusterSVGsBayes(algorithm = "pathfinder"),
 * R/clusterSVGsBayes.R:162-170, has no amplitude block to summarise) */
int bvgo_pf_single_phi(bvgo_ctx *c, const bvgo_pf_cfg *cfg, int path, double *phi /*[num_draws][D]*/, double *lp_ratio,
                       bvgo_pf_info *info, double *elbos) {
  double *amp = (double *)malloc(sizeof(double) * (size_t)cfg->num_draws * (size_t)(c->G > 0 ? c->G : 1));
  pf_phi_out = phi;
  const int rc = bvgo_pf_single(c, cfg, path, amp, lp_ratio, info, elbos);
  pf_phi_out = NULL;
  free(amp);
  return rc;
}
/* Pareto-smoothed importance weights [U: stan::services::psis::psis_weights; Vehtari et al. 2015, Zhang & Stephens 2009] */
typedef struct { double v; int i; } lr_t;
static int cmp_lr(const void *a, const void *b) { const double x = ((const lr_t *)a)->v, y = ((const lr_t *)b)->v; return (x > y) - (x < y); }
double bvgo_psis_weights(const double *lr_in, int S, double *w) {
  double mx = -INFINITY, khat = NAN;
  for (int i = 0; i < S; ++i) if (lr_in[i] > mx) mx = lr_in[i];
  double *lr = (double *)malloc(sizeof(double) * S);
  for (int i = 0; i < S; ++i) lr[i] = lr_in[i] - mx;
  const int tail = (int)fmin(0.2 * S, 3.0 * sqrt((double)S));
  if (tail >= 5) {
    lr_t *srt = (lr_t *)malloc(sizeof(lr_t) * S);
    for (int i = 0; i < S; ++i) { srt[i].v = lr[i]; srt[i].i = i; }
    qsort(srt, S, sizeof(lr_t), cmp_lr);
    const double cutoff = srt[S - tail - 1].v, ec = exp(cutoff);
    const int n = tail;
    double *x = (double *)malloc(sizeof(double) * n);
    for (int i = 0; i < n; ++i) x[i] = exp(srt[S - tail + i].v) - ec; /* ascending */
    if (x[n - 1] > 0 && isfinite(x[n - 1])) {
      /* gpdfit */
      const int M = 30 + (int)floor(sqrt((double)n));
      const double xstar = x[(int)floor(n / 4.0 + 0.5) - 1], prior = 3.0;
      double *th = (double *)malloc(sizeof(double) * 2 * M), *lt = th + M;
      for (int j = 0; j < M; ++j) {
        th[j] = 1.0 / x[n - 1] + (1.0 - sqrt((double)M / (j + 0.5))) / (prior * xstar);
        double kk = 0;
        for (int i = 0; i < n; ++i) kk += log1p(-th[j] * x[i]);
        kk /= n;
        lt[j] = n * (log(-th[j] / kk) - kk - 1.0);
      }
      double theta = 0;
      for (int j = 0; j < M; ++j) {
        double sw = 0;
        for (int q = 0; q < M; ++q) sw += exp(lt[q] - lt[j]);
        theta += th[j] / sw;
      }
      double k = 0;
      for (int i = 0; i < n; ++i) k += log1p(-theta * x[i]);
      k /= n;
      const double sigma = -k / theta;
      k = k * n / (n + 10.0) + 0.5 * 10.0 / (n + 10.0); /* weakly informative prior on k */
      khat = k;
      free(th);
      if (isfinite(k) && isfinite(sigma)) {
        for (int i = 0; i < n; ++i) {
          const double pq = (i + 0.5) / n;
          const double q = (fabs(k) < 1e-300 ? -sigma * log1p(-pq) : sigma * expm1(-k * log1p(-pq)) / k) + ec;
          lr[srt[S - tail + i].i] = fmin(log(q), 0.0); /* truncated at the largest raw ratio */
        }
      }
    }
    free(x); free(srt);
  }
  double lse = 0;
  for (int i = 0; i < S; ++i) lse += exp(lr[i]);
  for (int i = 0; i < S; ++i) w[i] = exp(lr[i]) / lse;
  free(lr);
  return khat;
}
/* multi-path Pathfinder: summary G x 8 of the PSIS-resampled amplitude draws; idx (optional) [num_psis_draws] */
int bvgo_pathfinder(bvgo_ctx *c, const bvgo_pf_cfg *cfg, double *summary, bvgo_pf_info *infos, int *idx_out, double *khat) {
  const int64_t G = c->G;
  const int P = cfg->num_paths, nd = cfg->num_draws, S = P * nd, ns = cfg->num_psis_draws;
  double *amp = (double *)malloc(sizeof(double) * ((size_t)S * G + 2 * (size_t)S)), *lr = amp + (size_t)S * G, *w = lr + S;
  int rc = 0;
  for (int p = 0; p < P && rc == 0; ++p) rc = bvgo_pf_single(c, cfg, p, amp + (size_t)p * nd * G, lr + (size_t)p * nd, infos ? infos + p : NULL, NULL);
  if (rc) { free(amp); return rc; }
  int *idx = (int *)malloc(sizeof(int) * ns);
  if (P > 1) {
    const double kh = bvgo_psis_weights(lr, S, w);
    if (khat) *khat = kh;
    for (int r = 0; r < ns; ++r) { /* inverse-CDF resampling with replacement */
      const double uu = bvgo_uniform(cfg->seed, STREAM_PF_PSIS, 0, (uint64_t)r);
      double cs = 0;
      int pick = S - 1;
      for (int i = 0; i < S; ++i) { cs += w[i]; if (uu < cs) { pick = i; break; } }
      idx[r] = pick;
    }
  } else {
    if (khat) *khat = NAN;
    for (int r = 0; r < ns; ++r) idx[r] = r % nd;
  }
  double *x = (double *)malloc(sizeof(double) * 2 * ns), *d = x + ns;
  rank_t *rk = (rank_t *)malloc(sizeof(rank_t) * G);
  for (int64_t g = 0; g < G; ++g) {
    for (int r = 0; r < ns; ++r) x[r] = amp[(size_t)idx[r] * G + g];
    rk[g].v = summarise_gene(x, d, ns, G, g, summary);
    rk[g].i = g;
  }
  fill_ranks(rk, G, summary);
  if (idx_out) memcpy(idx_out, idx, sizeof(int) * ns);
  free(rk); free(x); free(idx); free(amp);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * Posterior draws + amplitude summary.
 * R/findSpatiallyVariableFeaturesBayes.R:419-432: fit_vi$summary("amplitude") over the n.draws
 * draws -> mean, median, sd (n-1), mad (1.4826 * median|x - med|), q5, q95 (quantile type 7) [U:
 * posterior::default_summary_measures], dispersion = sd^2/mean (:428), rank by mean desc (:429-430).
 * ---------------------------------------------------------------------------------------- */
static double quantile7(const double *sorted, int n, double p) {
  const double h = (n - 1) * p;
  const int lo = (int)floor(h);
  const int hi = lo + 1 < n ? lo + 1 : lo;
  return sorted[lo] + (h - lo) * (sorted[hi] - sorted[lo]);
}
static int cmp_rank(const void *a, const void *b) {
  const rank_t *x = (const rank_t *)a, *y = (const rank_t *)b;
  if (x->v > y->v) return -1;
  if (x->v < y->v) return 1;
  return (x->i > y->i) - (x->i < y->i);
}
/* statistics of one gene's nd draws x (sorted in place; d = scratch of nd) -> out columns 0..6 */
static double summarise_gene(double *x, double *d, int nd, int64_t G, int64_t g, double *out) {
  double mean = 0;
  for (int t = 0; t < nd; ++t) mean += x[t];
  mean /= nd;
  double ss = 0;
  for (int t = 0; t < nd; ++t) ss += (x[t] - mean) * (x[t] - mean);
  const double sd = sqrt(ss / (nd - 1));
  qsort(x, nd, sizeof(double), cmp_d);
  const double med = quantile7(x, nd, 0.5);
  for (int t = 0; t < nd; ++t) d[t] = fabs(x[t] - med);
  qsort(d, nd, sizeof(double), cmp_d);
  out[0 * G + g] = mean;
  out[1 * G + g] = med;
  out[2 * G + g] = sd;
  out[3 * G + g] = 1.4826 * quantile7(d, nd, 0.5);
  out[4 * G + g] = quantile7(x, nd, 0.05);
  out[5 * G + g] = quantile7(x, nd, 0.95);
  out[6 * G + g] = sd * sd / mean;
  return mean;
}
static void fill_ranks(rank_t *rk, int64_t G, double *out) {
  qsort(rk, G, sizeof(rank_t), cmp_rank);
  for (int64_t r = 0; r < G; ++r) out[7 * G + rk[r].i] = (double)(r + 1);
}
int bvgo_summary(int model, int64_t G, int k, const double *mu, const double *om, int nd,
                 uint64_t seed, int64_t gene_offset, int64_t G_total, double *out) {
  const layout_t L = make_layout(model, G, k), LG = make_layout(model, G_total, k);
  double *x = (double *)malloc(sizeof(double) * 2 * nd), *d = x + nd;
  rank_t *rk = (rank_t *)malloc(sizeof(rank_t) * G);
  for (int64_t g = 0; g < G; ++g) {
    const double m = mu[L.amp + g], s = exp(om[L.amp + g]);
    const uint64_t gi = (uint64_t)(LG.amp + gene_offset + g); /* global index of amplitude[g] */
    for (int t = 0; t < nd; ++t) x[t] = exp(m + s * bvgo_normal(seed, STREAM_DRAW, (uint32_t)t, gi));
    rk[g].v = summarise_gene(x, d, nd, G, g, out);
    rk[g].i = g;
  }
  fill_ranks(rk, G, out);
  free(rk);
  free(x);
  return 0;
}
/* algorithm = "fullrank": amplitude[g] = exp(mu_u + sum_j L[u, j] eta_j), eta_j = normal(seed, stream 3, draw t, j) */
int bvgo_summary_fullrank(int model, int64_t G, int k, const double *mu, const double *Lc, int nd, uint64_t seed,
                          double *out) {
  const layout_t L = make_layout(model, G, k);
  const int64_t nrow = L.amp + G; /* rows of L needed: every row up to the last amplitude */
  double *x = (double *)malloc(sizeof(double) * 2 * nd), *d = x + nd;
  double *eta = (double *)malloc(sizeof(double) * (size_t)nd * nrow);
  rank_t *rk = (rank_t *)malloc(sizeof(rank_t) * G);
  for (int t = 0; t < nd; ++t)
    for (int64_t j = 0; j < nrow; ++j) eta[(size_t)t * nrow + j] = bvgo_normal(seed, STREAM_DRAW, (uint32_t)t, (uint64_t)j);
  for (int64_t g = 0; g < G; ++g) {
    const int64_t i = L.amp + g;
    const double *row = Lc + (size_t)i * (i + 1) / 2;
    for (int t = 0; t < nd; ++t) {
      double z = 0;
      for (int64_t j = 0; j <= i; ++j) z += row[j] * eta[(size_t)t * nrow + j];
      x[t] = exp(mu[i] + z);
    }
    rk[g].v = summarise_gene(x, d, nd, G, g, out);
    rk[g].i = g;
  }
  fill_ranks(rk, G, out);
  free(rk); free(eta); free(x);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * inst/clusterSVGs.stan (R/clusterSVGsBayes.R:139-160): diagonal-covariance Gaussian mixture over the PCA embedding of the
 * SVGs.  Parameters in declaration order (:8-12): simplex[K] theta (K - 1 unconstrained, stick-breaking [U:
 * stan::math::simplex_constrain as published in the Stan reference manual; the reference pins no CmdStan version and recent
 * releases may map the simplex differently -- that changes the shape of the variational family, not the model]),
 * matrix[K, D] mu (column-major), array[K] vector<lower=0>[D] sigma (log).
 *   to_vector(mu) ~ normal(0, 7); sigma[i] ~ normal(0, 3); theta ~ dirichlet(2);                      (:15-19)
 *   target += log_sum_exp_j(log theta_j - 0.5 (D log 2 pi + 2 sum log sigma_j) - 0.5 |(x_i - mu_j) / sigma_j|^2)   (:20-31)
 * ---------------------------------------------------------------------------------------- */
bvgo_ctx *bvgo_gmm_create(const double *X, int64_t N, int D, int K) {
  bvgo_ctx *c = (bvgo_ctx *)calloc(1, sizeof(*c));
  c->model = BVGO_GMM;
  c->gX = X; c->gN = N; c->gD = D; c->gK = K;
  c->nthreads = 1;
  c->G = K; /* unused by the drivers */
  c->L.D = (int64_t)K - 1 + 2 * (int64_t)K * D;
  return c;
}
/* constrained parameters from the unconstrained vector; returns the log-Jacobian; zk / stick kept for the backward pass */
static double gmm_constrain(const double *th, int K, int D, double *theta, double *mu, double *sigma, double *zk, double *stick) {
  double lj = 0, st = 1.0;
  for (int q = 0; q < K - 1; ++q) {
    const double a = th[q] - log((double)(K - 1 - q)), z = inv_logit(a);
    if (zk) { zk[q] = z; stick[q] = st; }
    theta[q] = st * z;
    lj += log(st) - log1p_exp(-a) - log1p_exp(a);
    st -= theta[q];
  }
  theta[K - 1] = st;
  const double *um = th + K - 1, *us = um + (size_t)K * D;
  for (int d = 0; d < D; ++d)
    for (int j = 0; j < K; ++j) mu[j * D + d] = um[j + K * d]; /* matrix[K, D] column-major -> mu[j][d] */
  for (int j = 0; j < K; ++j)
    for (int d = 0; d < D; ++d) { sigma[j * D + d] = exp(us[j * D + d]); lj += us[j * D + d]; }
  return lj;
}
static double gmm_log_prob(bvgo_ctx *c, const double *th, int jac, int propto, double *grad) {
  const int K = c->gK, D = c->gD;
  const int64_t N = c->gN;
  double theta[64], zk[64], stick[64];
  double *mu = (double *)malloc(sizeof(double) * (4 * (size_t)K * D + 3 * K)), *sigma = mu + K * D, *gmu = sigma + K * D, *gsg = gmu + K * D;
  double *lc = gsg + K * D, *R = lc + K, *lconst = R + K;
  const double lj = gmm_constrain(th, K, D, theta, mu, sigma, zk, stick);
  double lp = 0;
  for (int q = 0; q < K * D; ++q) {
    lp += -0.5 * mu[q] * mu[q] / 49.0 - 0.5 * sigma[q] * sigma[q] / 9.0;
    gmu[q] = -mu[q] / 49.0; gsg[q] = -sigma[q] / 9.0;
  }
  if (!propto) lp += K * D * (-0.5 * LOG_TWO_PI - log(7.0)) + K * D * (-0.5 * LOG_TWO_PI - log(3.0)) + lgam(2.0 * K);
  for (int j = 0; j < K; ++j) {
    lp += log(theta[j]); /* dirichlet(2): (alpha - 1) log theta */
    double sl = 0;
    for (int d = 0; d < D; ++d) sl += log(sigma[j * D + d]);
    lconst[j] = log(theta[j]) - 0.5 * (D * LOG_TWO_PI + 2.0 * sl);
    R[j] = 0;
  }
  for (int64_t i = 0; i < N; ++i) {
    const double *x = c->gX + (size_t)i * D;
    double mx = -INFINITY;
    for (int j = 0; j < K; ++j) {
      double ss = 0;
      for (int d = 0; d < D; ++d) { const double u = (x[d] - mu[j * D + d]) / sigma[j * D + d]; ss += u * u; }
      lc[j] = lconst[j] - 0.5 * ss;
      if (lc[j] > mx) mx = lc[j];
    }
    double se = 0;
    for (int j = 0; j < K; ++j) se += exp(lc[j] - mx);
    lp += mx + log(se);
    if (grad)
      for (int j = 0; j < K; ++j) {
        const double r = exp(lc[j] - mx) / se;
        R[j] += r;
        for (int d = 0; d < D; ++d) {
          const double df = x[d] - mu[j * D + d], sg = sigma[j * D + d];
          gmu[j * D + d] += r * df / (sg * sg);
          gsg[j * D + d] += r * (df * df / (sg * sg * sg) - 1.0 / sg);
        }
      }
  }
  if (jac) lp += lj;
  if (grad) {
    const double j1 = jac ? 1.0 : 0.0;
    double gx[64];
    for (int j = 0; j < K; ++j) gx[j] = (R[j] + 1.0) / theta[j];
    double gst = gx[K - 1]; /* reverse pass through the stick-breaking transform */
    for (int q = K - 2; q >= 0; --q) {
      const double z = zk[q], st = stick[q];
      const double gz = gx[q] * st - gst * st + j1 * (1.0 / z - 1.0 / (1.0 - z));
      gst = gx[q] * z + gst * (1.0 - z) + j1 / st;
      grad[q] = gz * z * (1.0 - z);
    }
    double *gm = grad + K - 1, *gs = gm + (size_t)K * D;
    for (int d = 0; d < D; ++d)
      for (int j = 0; j < K; ++j) gm[j + K * d] = gmu[j * D + d];
    for (int q = 0; q < K * D; ++q) gs[q] = gsg[q] * sigma[q] + j1;
  }
  free(mu);
  return lp;
}
/* posterior pieces of R/clusterSVGsBayes.R:201-291 from n_draws draws of the approximation (Philox stream 3, draw t):
 * draws [n_draws][K + 2 K D] constrained (theta | mu column-major | sigma), soft [N][K] = mean over draws of the per-draw
 * responsibilities (:215-239), *loglik = mean over draws of sum_i log_lik_i (:281-284) */
int bvgo_gmm_post(bvgo_ctx *c, const double *mu_v, const double *var, int fullrank, int nd, uint64_t seed, double *draws,
                  double *soft, double *loglik) {
  const int K = c->gK, D = c->gD;
  const int64_t N = c->gN, P = c->L.D;
  double *eta = (double *)malloc(sizeof(double) * (2 * (size_t)P + 2 * (size_t)K * D + 3 * K)), *zeta = eta + P, *mu = zeta + P, *sigma = mu + K * D;
  double *lc = sigma + K * D, *lconst = lc + K, theta[64];
  memset(soft, 0, sizeof(double) * (size_t)N * K);
  double ll = 0;
  for (int t = 0; t < nd; ++t) {
    for (int64_t i = 0; i < P; ++i) eta[i] = bvgo_normal(seed, STREAM_DRAW, (uint32_t)t, (uint64_t)i);
    if (fullrank) fr_transform(P, mu_v, var, eta, zeta);
    else for (int64_t i = 0; i < P; ++i) zeta[i] = mu_v[i] + exp(var[i]) * eta[i];
    gmm_constrain(zeta, K, D, theta, mu, sigma, NULL, NULL);
    if (draws) {
      double *o = draws + (size_t)t * (K + 2 * K * D);
      for (int j = 0; j < K; ++j) o[j] = theta[j];
      for (int d = 0; d < D; ++d) for (int j = 0; j < K; ++j) o[K + j + K * d] = mu[j * D + d];
      for (int q = 0; q < K * D; ++q) o[K + K * D + q] = sigma[q];
    }
    for (int j = 0; j < K; ++j) {
      double sl = 0;
      for (int d = 0; d < D; ++d) sl += log(sigma[j * D + d]);
      lconst[j] = log(theta[j]) - 0.5 * (D * LOG_TWO_PI + 2.0 * sl);
    }
    for (int64_t i = 0; i < N; ++i) {
      const double *x = c->gX + (size_t)i * D;
      double mx = -INFINITY, se = 0;
      for (int j = 0; j < K; ++j) {
        double ss = 0;
        for (int d = 0; d < D; ++d) { const double u = (x[d] - mu[j * D + d]) / sigma[j * D + d]; ss += u * u; }
        lc[j] = lconst[j] - 0.5 * ss;
        if (lc[j] > mx) mx = lc[j];
      }
      for (int j = 0; j < K; ++j) se += exp(lc[j] - mx);
      ll += mx + log(se);
      for (int j = 0; j < K; ++j) soft[(size_t)i * K + j] += exp(lc[j] - mx) / se / nd;
    }
  }
  *loglik = ll / nd;
  free(eta);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * Basis builder.  R/findSpatiallyVariableFeaturesBayes.R:265-281 (kernel between every spot and
 * every k-means centroid on SQUARED distances), R/maternKernel.R:22-26, R/expQuadKernel.R:17,
 * R/periodicKernel.R:20, then :283-288 qr.Q(qr(phi, LAPACK = TRUE)) = thin Q of column-pivoted
 * Householder QR [U: LAPACK dgeqp3 + dorgqr].
 * ---------------------------------------------------------------------------------------- */
static double kern(double d2, double ls, int kernel, double nu, double period) {
  if (kernel == BVGO_KERNEL_EXP_QUAD) return exp(-d2 / (2.0 * ls * ls));
  if (kernel == BVGO_KERNEL_PERIODIC) {
    const double sn = sin(M_PI * sqrt(d2) / period);
    return exp(-2.0 * sn * sn / (ls * ls));
  }
  /* Matern: (2^(1-nu)/Gamma(nu)) s^nu K_nu(s), s = sqrt(2 nu) sqrt(d)/l; non-finite (d = 0) -> 1.
   * For nu in {1/2, 3/2, 5/2} (the only values accepted, :106) the Bessel form equals: */
  const double s = sqrt(2.0 * nu) * sqrt(d2) / ls;
  if (nu == 0.5) return exp(-s);
  if (nu == 1.5) return (1.0 + s) * exp(-s);
  return (1.0 + s + s * s / 3.0) * exp(-s);
}
static double nrm2(const double *x, int64_t n) {
  /* scaled 2-norm (dnrm2 semantics) */
  double scale = 0, ssq = 1;
  for (int64_t i = 0; i < n; ++i)
    if (x[i] != 0) {
      const double a = fabs(x[i]);
      if (scale < a) { ssq = 1 + ssq * (scale / a) * (scale / a); scale = a; }
      else ssq += (a / scale) * (a / scale);
    }
  return scale * sqrt(ssq);
}
int bvgo_basis(const double *X, int64_t M, const double *C, int k, double ls, int kernel, double nu,
               double period, double *Q, double *raw) {
  double *A = (double *)malloc(sizeof(double) * M * k);
  for (int j = 0; j < k; ++j)
    for (int64_t s = 0; s < M; ++s) {
      const double dx = X[s] - C[j], dy = X[M + s] - C[k + j];
      A[(int64_t)j * M + s] = kern(dx * dx + dy * dy, ls, kernel, nu, period);
    }
  if (raw) memcpy(raw, A, sizeof(double) * M * k);
  double *vn1 = (double *)malloc(sizeof(double) * 3 * k), *vn2 = vn1 + k, *tau = vn2 + k;
  for (int j = 0; j < k; ++j) vn1[j] = vn2[j] = nrm2(A + (int64_t)j * M, M);
  const double tol3z = sqrt(DBL_EPSILON);
  for (int j = 0; j < k && j < M; ++j) {
    int pv = j;
    for (int c2 = j + 1; c2 < k; ++c2) if (vn1[c2] > vn1[pv]) pv = c2; /* idamax: first maximum */
    if (pv != j) {
      for (int64_t s = 0; s < M; ++s) { double t = A[(int64_t)pv * M + s]; A[(int64_t)pv * M + s] = A[(int64_t)j * M + s]; A[(int64_t)j * M + s] = t; }
      vn1[pv] = vn1[j]; vn2[pv] = vn2[j];
    }
    double *a = A + (int64_t)j * M;
    /* dlarfg */
    const double alpha = a[j], xn = nrm2(a + j + 1, M - j - 1);
    if (xn == 0) tau[j] = 0;
    else {
      const double beta = -copysign(hypot(alpha, xn), alpha);
      tau[j] = (beta - alpha) / beta;
      const double sc = 1.0 / (alpha - beta);
      for (int64_t s = j + 1; s < M; ++s) a[s] *= sc;
      a[j] = beta;
    }
    /* apply H_j to trailing columns */
    for (int c2 = j + 1; c2 < k; ++c2) {
      double *b = A + (int64_t)c2 * M;
      double w = b[j];
      for (int64_t s = j + 1; s < M; ++s) w += a[s] * b[s];
      w *= tau[j];
      b[j] -= w;
      for (int64_t s = j + 1; s < M; ++s) b[s] -= w * a[s];
      /* dlaqp2 partial-norm downdate */
      if (vn1[c2] != 0) {
        double t = 1.0 - (fabs(b[j]) / vn1[c2]) * (fabs(b[j]) / vn1[c2]);
        t = t > 0 ? t : 0;
        const double t2 = t * (vn1[c2] / vn2[c2]) * (vn1[c2] / vn2[c2]);
        if (t2 <= tol3z) { vn1[c2] = nrm2(b + j + 1, M - j - 1); vn2[c2] = vn1[c2]; }
        else vn1[c2] *= sqrt(t);
      }
    }
  }
  /* dorg2r: Q = H_0 ... H_{k-1} [I; 0] */
  memset(Q, 0, sizeof(double) * M * k);
  for (int j = 0; j < k; ++j) Q[(int64_t)j * M + j] = 1.0;
  for (int j = k - 1; j >= 0; --j) {
    const double *a = A + (int64_t)j * M;
    for (int c2 = j; c2 < k; ++c2) {
      double *b = Q + (int64_t)c2 * M;
      double w = b[j];
      for (int64_t s = j + 1; s < M; ++s) w += a[s] * b[s];
      w *= tau[j];
      b[j] -= w;
      for (int64_t s = j + 1; s < M; ++s) b[s] -= w * a[s];
    }
  }
  free(vn1);
  free(A);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * k-means centroids, R/findSpatiallyVariableFeaturesBayes.R:200-204.  Lloyd-Forgy; sums in 64-bit fixed point
 * (2^-32 for coordinates, 2^-20 for the within-SS) so that the device version, whose atomics add in any order,
 * gives the same bits.
 * ---------------------------------------------------------------------------------------- */
int bvgo_kmeans(const double *X, int64_t M, int k, int iter_max, int nstart, uint64_t seed, double *centers_out,
                double *tot_withinss, int32_t *info) {
  const double FIX = 4294967296.0, SSFIX = 1048576.0;
  if (M < 1 || k < 1 || k > M || iter_max < 1 || nstart < 1) return -1;
  int64_t *rows = (int64_t *)malloc(sizeof(int64_t) * k);
  int32_t *lab = (int32_t *)malloc(sizeof(int32_t) * M);
  double *C = (double *)malloc(sizeof(double) * 2 * k);
  int64_t *acc = (int64_t *)malloc(sizeof(int64_t) * 3 * k);
  uint64_t best_ss = ~(uint64_t)0;
  for (int r = 0; r < nstart; ++r) {
    for (int j = 0; j < k; ++j) /* k distinct rows: first Philox value (counter (j, attempt, r, 4)) mod M not yet drawn */
      for (uint32_t a = 0;; ++a) {
        uint32_t ctr[4] = {(uint32_t)j, a, (uint32_t)r, 4u}, key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)}, o[4];
        bvgo_philox4x32_10(ctr, key, o);
        const int64_t row = (int64_t)((((uint64_t)o[0] << 32) | o[1]) % (uint64_t)M);
        int dup = 0;
        for (int q = 0; q < j && !dup; ++q) dup = rows[q] == row;
        if (!dup) { rows[j] = row; break; }
      }
    for (int j = 0; j < k; ++j) { C[j] = X[rows[j]]; C[k + j] = X[M + rows[j]]; }
    for (int64_t s = 0; s < M; ++s) lab[s] = -1;
    int it = 0, conv = 0;
    uint64_t ss;
    for (;;) {
      int64_t changed = 0;
      ss = 0;
      memset(acc, 0, sizeof(int64_t) * 3 * k);
      for (int64_t s = 0; s < M; ++s) {
        const double x = X[s], y = X[M + s];
        double best = INFINITY;
        int bj = 0;
        for (int j = 0; j < k; ++j) {
          const double dx = x - C[j], dy = y - C[k + j];
          const double d = fma(dx, dx, dy * dy);
          if (d < best) { best = d; bj = j; }
        }
        changed += lab[s] != bj;
        lab[s] = bj;
        ss += (uint64_t)llrint(best * SSFIX);
        acc[bj] += (int64_t)llrint(x * FIX);
        acc[k + bj] += (int64_t)llrint(y * FIX);
        acc[2 * k + bj] += 1;
      }
      if (changed == 0) { conv = 1; break; }
      if (it == iter_max) break;
      for (int j = 0; j < k; ++j)
        if (acc[2 * k + j]) {
          C[j] = (double)acc[j] / (double)acc[2 * k + j] * (1.0 / FIX);
          C[k + j] = (double)acc[k + j] / (double)acc[2 * k + j] * (1.0 / FIX);
        }
      ++it;
    }
    if (ss < best_ss) {
      best_ss = ss;
      memcpy(centers_out, C, sizeof(double) * 2 * k);
      if (info) { info[0] = r; info[1] = it; info[2] = conv; }
    }
  }
  if (tot_withinss) *tot_withinss = (double)best_ss / SSFIX;
  free(rows); free(lab); free(C); free(acc);
  return 0;
}
