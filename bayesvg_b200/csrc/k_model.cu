// k_model.cu -- the O(kG) part of one log-density / ELBO-gradient evaluation:
//   k_prep   : Monte Carlo draw zeta = mu + exp(omega) * eps (counter-based Philox), constrain,
//              and build the coefficient rows c'_g = (A_g alpha_g, beta0_g) fed to the likelihood GEMM
//   k_gene   : per-gene backward pass from (P_g, R_g, SSE_g) + fused mean-field ADVI update of the
//              per-gene parameters + block partials of the 2k+10 cross-gene sums
//   k_finish : fixed-order reduction of those sums, log density, gradient / update of the 2k+5
//              gene-shared parameters.
// Formulas: SURVEY.md A.3 (restating inst/approxGP.stan:25-49 and inst/approxGP4.stan:25-49 of the
// reference); ADVI update: SURVEY.md A.4 [stan::variational::normal_meanfield::calc_grad].
#include "bvg_internal.cuh"

// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_prep(DevModel m, int mode, int tc) {
  extern __shared__ double sh[];  // [2k+5] shared zeta | [k] sigma_alpha
  const int k = m.k, nsh = 2 * k + 5;
  const uint32_t stream = mode == PREP_GRAD ? STREAM_GRAD : STREAM_ELBO;
  const uint32_t t = mode == PREP_GRAD ? m.ctl->t_grad : m.ctl->t_elbo;
  for (int i = threadIdx.x; i < nsh; i += blockDim.x) {
    const int64_t li = shared_index(m.L, k, i);
    double z;
    if (mode != PREP_NONE) {
      z = m.mu[li] + exp(m.omega[li]) * bvg_normal(m.seed, stream, t, (uint64_t)shared_index(m.LG, k, i));
      if (blockIdx.x == 0) m.zeta[li] = z;
    } else {
      z = m.zeta[li];
    }
    sh[i] = z;
  }
  __syncthreads();
  double *sal = sh + nsh;
  for (int j = threadIdx.x; j < k; j += blockDim.x) sal[j] = exp(sh[5 + k + j]);
  __syncthreads();
  const double mb = sh[0], sb = exp(sh[1]);
  const int lane = threadIdx.x & 31;
  const int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (g >= m.G) return;
  const int64_t gg = m.gene_offset + g;
  // lane 0: z_beta0[g], lane 1: log amplitude[g]
  double v = 0;
  if (lane < 2) {
    const int64_t li = lane == 0 ? m.L.z_beta0 + g : m.L.amp + g;
    if (mode != PREP_NONE) {
      const int64_t gi = lane == 0 ? m.LG.z_beta0 + gg : m.LG.amp + gg;
      v = m.mu[li] + exp(m.omega[li]) * bvg_normal(m.seed, stream, t, (uint64_t)gi);
      m.zeta[li] = v;
    } else {
      v = m.zeta[li];
    }
  }
  const double zb = __shfl_sync(0xffffffffu, v, 0), u = __shfl_sync(0xffffffffu, v, 1);
  const double A = exp(2.0 * u);        // amplitude_sq (approxGP.stan:27)
  const double beta = mb + sb * zb;     // beta0        (approxGP.stan:26)
  // tc != 0: rows of width 32, split into tf32 hi / lo parts for the 3xTF32 tensor-core contraction
  const int cw = tc ? 32 : m.KP;
  T *Crow = (T *)m.C + g * cw, *Clo = (T *)m.C + (m.Gpad + g) * cw;
  auto put = [&](int j, double v) {
    if (!tc) { Crow[j] = (T)v; return; }
    const float x = (float)v;
    uint32_t hb;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(x));
    const float hi = __uint_as_float(hb);
    Crow[j] = (T)hi;
    Clo[j] = (T)(x - hi);
  };
  for (int j = lane; j < k; j += 32) {
    const int64_t li = m.L.z_alpha + g * k + j;
    double z;
    if (mode != PREP_NONE) {
      z = m.mu[li] + exp(m.omega[li]) * bvg_normal(m.seed, stream, t, (uint64_t)(m.LG.z_alpha + gg * k + j));
      m.zeta[li] = z;
    } else {
      z = m.zeta[li];
    }
    put(j, A * (sh[5 + j] + sal[j] * z));  // A_g * alpha_t[j,g] (approxGP.stan:32)
  }
  if (lane == 0) put(k, beta);
}

int launch_prep(bvg_fit *f, int mode) {
  const int wpb = 8;
  const int grid = (int)((f->m.G + wpb - 1) / wpb);
  const size_t sh = sizeof(double) * (3 * f->m.k + 5);
  if (f->cfg.precision == BVG_PREC_FP64) k_prep<double><<<grid, wpb * 32, sh, f->stream>>>(f->m, mode, 0);
  else k_prep<float><<<grid, wpb * 32, sh, f->stream>>>(f->m, mode, f->engine == BVG_ENGINE_TCGEN05);
  f->st.kernel_launches++;
  CUDA_TRY(cudaGetLastError());
  return BVG_OK;
}

// ------------------------------------------------------------------------------------------------
// mean-field ADVI update of one parameter: grad_mu = g, grad_omega = g*eps*exp(omega)+1,
// s <- (it==1 ? g^2 : 0.9 s + 0.1 g^2),  x += eta/sqrt(it) * grad / (1 + sqrt(s))
__device__ inline void advi_update(const DevModel &m, int64_t li, int64_t gi, double gv, bool fin, uint32_t t,
                                   int it, double es) {
  const double om = m.omega[li];
  const double eps = bvg_normal(m.seed, STREAM_GRAD, t, (uint64_t)gi);
  const double gm = fin ? gv : 0.0;
  const double go = fin ? gv * eps * exp(om) + 1.0 : 0.0;
  double hm = m.hist[li], ho = m.hist[m.L.D + li];
  if (it == 1) { hm += gm * gm; ho += go * go; }
  else { hm = 0.9 * hm + 0.1 * gm * gm; ho = 0.9 * ho + 0.1 * go * go; }
  m.hist[li] = hm;
  m.hist[m.L.D + li] = ho;
  m.mu[li] += es * gm / (1.0 + sqrt(hm));
  m.omega[li] = om + es * go / (1.0 + sqrt(ho));
}

template <typename T>
__global__ void __launch_bounds__(256) k_gene(DevModel m, int mode, int jac) {
  extern __shared__ double sm[];  // [nwarps][NS]
  const int k = m.k, KP = m.KP, NS = bvg_ns(k);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const Layout &L = m.L;
  const double *zt = m.zeta;
  const double sb = exp(zt[L.sigma_beta0]), ma = zt[L.mu_amp], sa = exp(zt[L.sigma_amp]);
  const double tail = exp(zt[L.tail]);
  const double inv_s2 = m.lik == BVG_GAUSSIAN ? 1.0 / (tail * tail) : 1.0;
  const double mual = lane < k ? zt[L.mu_alpha + lane] : 0.0;
  const double sal = lane < k ? exp(zt[L.sigma_alpha + lane]) : 0.0;
  const uint32_t t = m.ctl->t_grad;
  const int it = m.ctl->it;
  const double es = m.ctl->eta / sqrt((double)it);
  double aR = 0, aRz = 0, aDu = 0, aDu2 = 0, aS0 = 0, aS1 = 0, aZ2 = 0, aU = 0, aAP = 0, aAPz = 0;
  const T *Pp = (const T *)m.Ppart, *Sp = (const T *)m.Spart;
  for (int64_t g = (int64_t)blockIdx.x * nw + warp; g < m.G; g += (int64_t)gridDim.x * nw) {
    double Pj = 0, s0 = 0;
    for (int sp = 0; sp < m.nsplit; ++sp) {  // fixed order -> deterministic
      const int64_t row = (int64_t)sp * m.Gpad + g;
      if (lane < KP) Pj += (double)Pp[row * KP + lane];
      if (lane < 2) s0 += (double)Sp[row * 2 + lane];
    }
    const double s1 = __shfl_sync(0xffffffffu, s0, 1);
    s0 = __shfl_sync(0xffffffffu, s0, 0);
    Pj *= inv_s2;
    const double R = __shfl_sync(0xffffffffu, Pj, k);  // ones column of Phi' -> R_g = sum_s r_sg
    const double zj = lane < k ? zt[L.z_alpha + g * k + lane] : 0.0;
    const double zb = zt[L.z_beta0 + g], u = zt[L.amp + g];
    const double A = exp(2.0 * u), alpha = mual + sal * zj;
    const double aP = warp_sum(lane < k ? alpha * Pj : 0.0);
    const double z2 = warp_sum(zj * zj) + zb * zb;
    const double AP = lane < k ? A * Pj : 0.0;
    const double du = u - ma;
    aAP += AP; aAPz += AP * zj;
    aR += R; aRz += R * zb; aDu += du; aDu2 += du * du; aS0 += s0; aS1 += s1; aZ2 += z2; aU += u;
    if (mode == MODE_GRAD_OUT || mode == MODE_ADVI_UPDATE) {
      const int64_t gg = m.gene_offset + g;
      double gv = 0;
      int64_t li = 0, gi = 0;
      bool has = true;
      if (lane < k) { gv = sal * AP - zj; li = L.z_alpha + g * k + lane; gi = m.LG.z_alpha + gg * k + lane; }
      else if (lane == k) { gv = sb * R - zb; li = L.z_beta0 + g; gi = m.LG.z_beta0 + gg; }
      else if (lane == k + 1) { gv = 2.0 * A * aP - du / (sa * sa) - (jac ? 0.0 : 1.0); li = L.amp + g; gi = m.LG.amp + gg; }
      else has = false;
      if (mode == MODE_GRAD_OUT) {
        if (has) m.grad[li] = gv;
      } else {
        const bool fin = __all_sync(0xffffffffu, has ? isfinite(gv) : true);
        if (!fin && lane == 0) atomicAdd(&m.ctl->bad, 1);
        if (has) advi_update(m, li, gi, gv, fin, t, it, es);
      }
    }
  }
  double *row = sm + warp * NS;
  if (lane == 0) {
    row[S_R] = aR; row[S_RZ] = aRz; row[S_DU] = aDu; row[S_DU2] = aDu2; row[S_SSE] = aS0; row[S_Z2] = aZ2;
    row[S_U] = aU; row[S_L] = aS1; row[S_LG] = 0; row[S_DG] = 0;
  }
  if (lane < k) { row[S_AP + lane] = aAP; row[S_AP + k + lane] = aAPz; }
  __syncthreads();
  for (int c = threadIdx.x; c < NS; c += blockDim.x) {
    double s = 0;
    for (int w = 0; w < nw; ++w) s += sm[w * NS + c];
    m.blk[(int64_t)blockIdx.x * NS + c] = s;
  }
}

int launch_gene(bvg_fit *f, int mode, int jac) {
  const int wpb = 8;
  const size_t sh = sizeof(double) * wpb * bvg_ns(f->m.k);
  if (f->cfg.precision == BVG_PREC_FP64) k_gene<double><<<f->m.nblk_gene, wpb * 32, sh, f->stream>>>(f->m, mode, jac);
  else k_gene<float><<<f->m.nblk_gene, wpb * 32, sh, f->stream>>>(f->m, mode, jac);
  f->st.kernel_launches++;
  CUDA_TRY(cudaGetLastError());
  return BVG_OK;
}

// ------------------------------------------------------------------------------------------------
__device__ inline double dev_digamma(double x) {
  double r = 0.0;
  while (x < 10.0) { r -= 1.0 / x; x += 1.0; }
  const double f = 1.0 / (x * x);
  const double t = f * (-1.0 / 12.0 + f * (1.0 / 120.0 + f * (-1.0 / 252.0 + f * (1.0 / 240.0 + f * (-1.0 / 132.0 + f * (691.0 / 32760.0 + f * (-1.0 / 12.0)))))));
  return r + log(x) - 0.5 / x + t;
}

// phase bit 0: reduce block partials (+ nb histogram terms) into tot[]; bit 1: finish.
// Between the two phases a multi-GPU fit all-reduces tot[] (comm.cu).
__global__ void __launch_bounds__(256) k_finish(DevModel m, int mode, int jac, int propto, int phase) {
  __shared__ double red[2][8];
  __shared__ double s_lp;
  const int k = m.k, NS = bvg_ns(k), tid = threadIdx.x;
  const Layout &L = m.L;
  const double *zt = m.zeta;
  if (phase & 1) {
    for (int c = tid; c < NS; c += blockDim.x) {
      double s = 0;
      for (int b = 0; b < m.nblk_gene; ++b) s += m.blk[(int64_t)b * NS + c];
      m.tot[c] = s;
    }
    if (m.lik == BVG_NB) {
      // sum_sg lgamma(y+phi) - lgamma(phi) and its digamma derivative depend only on (y, phi):
      // evaluate on the histogram of distinct counts
      const double phi = exp(zt[L.tail]), lg0 = lgamma(phi), dg0 = dev_digamma(phi);
      double lg = 0, dg = 0;
      for (int i = tid; i < m.nhist; i += blockDim.x) {
        const double c = m.hist_val[i], n = m.hist_cnt[i];
        lg += n * (lgamma(c + phi) - lg0);
        dg += n * (dev_digamma(c + phi) - dg0);
      }
      lg = warp_sum(lg); dg = warp_sum(dg);
      if ((tid & 31) == 0) { red[0][tid >> 5] = lg; red[1][tid >> 5] = dg; }
      __syncthreads();
      if (tid == 0) {
        double a = 0, b = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += red[0][w]; b += red[1][w]; }
        m.tot[S_LG] = a - m.lgam_y1;
        m.tot[S_DG] = b;
      }
    }
    __syncthreads();
  }
  if (!(phase & 2)) return;
  const double *tot = m.tot;
  const double mb = zt[L.mu_beta0], lsb = zt[L.sigma_beta0], sb = exp(lsb);
  const double ma = zt[L.mu_amp], lsa = zt[L.sigma_amp], sa = exp(lsa);
  const double lt = zt[L.tail], tl = exp(lt);
  const double Gt = (double)m.G_total, N = (double)m.M * Gt;
  const double j1 = jac ? 1.0 : 0.0;
  if (tid < 32) {  // warp 0: log density
    double pa = 0;  // prior terms of mu_alpha, sigma_alpha + jacobian of sigma_alpha
    for (int j = tid; j < k; j += 32) {
      const double a = zt[L.mu_alpha + j], ls = zt[L.sigma_alpha + j], s = exp(ls);
      pa += -a * a / 8.0 - 0.5 * s * s + j1 * ls;
    }
    pa = warp_sum(pa);
    if (tid == 0) {
      double lp;
      if (m.lik == BVG_GAUSSIAN) lp = -N * lt - 0.5 * tot[S_SSE] / (tl * tl) - 0.5 * tl * tl;
      else lp = tot[S_SSE] + tot[S_LG] - log1p(tl * tl / 4.0);
      lp += -0.5 * tot[S_Z2] - mb * mb / 8.0 - 0.5 * sb * sb - ma * ma / 8.0 - 0.5 * sa * sa;
      lp += -Gt * lsa - tot[S_U] - 0.5 * tot[S_DU2] / (sa * sa);
      lp += pa;
      if (jac) lp += lsb + lsa + tot[S_U] + lt;
      if (!propto) {
        if (m.lik == BVG_GAUSSIAN)
          lp += -0.5 * BVG_LOG_TWO_PI * (N + (double)k * Gt + 2.0 * Gt + 2.0 * k + 5.0) - (k + 2.0) * BVG_LOG_TWO;
        else
          lp += -0.5 * BVG_LOG_TWO_PI * ((double)k * Gt + 2.0 * Gt + 2.0 * k + 4.0) - (k + 2.0) * BVG_LOG_TWO - BVG_LOG_PI;
      }
      s_lp = lp;
    }
  }
  __syncthreads();
  const double lp = s_lp;
  const uint32_t t = m.ctl->t_grad;
  const int it = m.ctl->it;
  if (mode == MODE_GRAD_OUT || mode == MODE_ADVI_UPDATE) {
    const double es = m.ctl->eta / sqrt((double)it);
    for (int i = tid; i < 2 * k + 5; i += blockDim.x) {
      double gv;
      if (i == 0) gv = tot[S_R] - mb / 4.0;
      else if (i == 1) gv = sb * (tot[S_RZ] - sb) + j1;
      else if (i == 2) gv = tot[S_DU] / (sa * sa) - ma / 4.0;
      else if (i == 3) gv = -Gt + tot[S_DU2] / (sa * sa) - sa * sa + j1;
      else if (i == 4) {
        if (m.lik == BVG_GAUSSIAN) gv = -N + tot[S_SSE] / (tl * tl) - tl * tl + j1;
        else gv = tl * (tot[S_DG] - tot[S_L] - tot[S_R] / tl) - 0.5 * tl * tl / (1.0 + tl * tl / 4.0) + j1;
      } else if (i < 5 + k) gv = tot[S_AP + (i - 5)] - zt[L.mu_alpha + (i - 5)] / 4.0;
      else {
        const int j = i - 5 - k;
        const double s = exp(zt[L.sigma_alpha + j]);
        gv = s * (tot[S_AP + k + j] - s) + j1;
      }
      const int64_t li = shared_index(L, k, i);
      if (mode == MODE_GRAD_OUT) m.grad[li] = gv;
      else {
        const bool fin = isfinite(gv);
        if (!fin) atomicAdd(&m.ctl->bad, 1);
        advi_update(m, li, shared_index(m.LG, k, i), gv, fin, t, it, es);
      }
    }
  }
  __syncthreads();
  if (tid == 0) {
    m.ctl->lp = lp;
    if (mode == MODE_ADVI_UPDATE) { m.ctl->t_grad = t + 1; m.ctl->it = it + 1; }
    else if (mode == MODE_ELBO_DRAW) {
      if (isfinite(lp)) { m.ctl->elbo_acc += lp; m.ctl->elbo_ok += 1; }
      m.ctl->t_elbo += 1;
    }
  }
}

int launch_finish(bvg_fit *f, int mode, int jac, int propto) {
  if (f->comm) {
    k_finish<<<1, 256, 0, f->stream>>>(f->m, mode, jac, propto, 1);
    CUDA_TRY(cudaGetLastError());
    BVG_TRY(comm_allreduce(f, f->m.tot, bvg_ns(f->m.k)));
    k_finish<<<1, 256, 0, f->stream>>>(f->m, mode, jac, propto, 2);
    f->st.kernel_launches += 2;
  } else {
    k_finish<<<1, 256, 0, f->stream>>>(f->m, mode, jac, propto, 3);
    f->st.kernel_launches++;
  }
  CUDA_TRY(cudaGetLastError());
  return BVG_OK;
}
