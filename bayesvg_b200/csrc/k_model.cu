// k_model.cu -- the O(kG) part of one log-density / ELBO-gradient evaluation:
//   k_prep   : Monte Carlo draw zeta = mu + exp(omega) * eps (counter-based Philox), constrain,
//              and build the coefficient rows c'_g = (A_g alpha_g, beta0_g) fed to the likelihood GEMM
//   k_gene   : per-gene backward pass from (P_g, R_g, SSE_g) + fused mean-field ADVI update of the
//              per-gene parameters + block partials of the 2k+10 cross-gene sums
//   k_finish : fixed-order reduction of those sums, log density, gradient / update of the 2k+5
//              gene-shared parameters.
// Formulas: SURVEY.md A.3 (restating inst/approxGP.stan:25-49 and inst/approxGP4.stan:25-49 of the
// reference); ADVI update: SURVEY.md A.4 [stan::variational::normal_meanfield::calc_grad].
#include <algorithm>

#include "bvg_internal.cuh"

// ------------------------------------------------------------------------------------------------
// k_prep: ONE thread per parameter.  A block owns `gpb` genes ((k+2) parameters each) and, redundantly,
// the 2k+5 gene-shared parameters, so every normal draw of the block is generated in one parallel
// phase (Philox + Box-Muller in fp64 is a ~1000-cycle dependent chain; the first version chained
// three of them per block).
template <typename T>
__global__ void __launch_bounds__(256) k_prep(DevModel m, int mode, int tc, int gpb, LsPoint ls) {
  extern __shared__ double sh[];  // [2k+5] shared zeta | [k] sigma_alpha | [gpb][k+2] gene zeta | [gpb] A_g
  griddep_wait();    // mu / omega come from the previous evaluation's k_gene
  griddep_launch();
  const int k = m.k, nsh = 2 * k + 5, npg = k + 2, tid = threadIdx.x;
  double *sal = sh + nsh, *zg = sal + k, *Ag = zg + gpb * npg;
  const uint32_t stream = mode == PREP_GRAD ? STREAM_GRAD : (mode == PREP_DRAW ? STREAM_DRAW : STREAM_ELBO);
  // batched form (gridDim.y = B draws, tcgen05 engine): draw b of the batch uses counter t + b and leaves its zeta / shx /
  // Aexp / W / U in slice b of the batch buffers the model descriptor points at (advi_elbo: all draws of an ELBO estimate in
  // one throughput-bound launch instead of one latency-bound launch per draw)
  const uint32_t bb = blockIdx.y;
  const uint32_t t = (mode == PREP_GRAD ? m.ctl->t_grad : (mode == PREP_DRAW ? m.ctl->t_draw : m.ctl->t_elbo)) + bb;
  double *const zeta_o = m.zeta + (size_t)bb * m.L.D, *const shx_o = m.shx + (size_t)bb * (4 + k), *const Aexp_o = m.Aexp + (size_t)bb * m.G;
  float *const W_o = m.W + (size_t)bb * m.Gpad * 32, *const U_o = m.U + (size_t)bb * 96;
  // ---- phase 1: every parameter's zeta ----
  const int t2 = tid - nsh, gl = t2 >= 0 ? t2 / npg : -1, j = t2 >= 0 ? t2 - gl * npg : 0;
  const int64_t g = (int64_t)blockIdx.x * gpb + gl;
  const bool is_sh = tid < nsh, is_gene = gl >= 0 && gl < gpb && g < m.G;
  if (is_gene && m.depth && j == k) {
    zg[gl * npg + k] = m.depths[g];  // depth-adjusted programs: the intercept "z" of the gene is data (approxGP2.stan:46)
  } else if (is_sh || is_gene) {
    int64_t li, gi;
    if (is_sh) { li = shared_index(m.L, k, tid); gi = shared_index(m.LG, k, tid); }
    else {
      const int64_t gg = m.gene_offset + g;
      if (j < k) { li = m.L.z_alpha + g * k + j; gi = m.LG.z_alpha + gg * k + j; }
      else if (j == k) { li = m.L.z_beta0 + g; gi = m.LG.z_beta0 + gg; }
      else { li = m.L.amp + g; gi = m.LG.amp + gg; }
    }
    double z;
    if (mode != PREP_NONE) {
      z = m.mu[li] + exp(m.omega[li]) * bvg_normal(m.seed, stream, t, (uint64_t)gi);
      if (is_gene || blockIdx.x == 0) zeta_o[li] = z;
    } else if (ls.x0) {  // trial point of the device-resident line search (lbfgs_dev.cu): x0 + a p, a read from the device
      z = ls.x0[li] + *ls.a * ls.p[li];
      if (is_gene || blockIdx.x == 0) zeta_o[li] = z;
    } else {
      z = m.zeta[li];
    }
    if (is_sh) sh[tid] = z; else zg[gl * npg + j] = z;
  }
  __syncthreads();
  // ---- phase 2: the exponentials (sigma_alpha, sigma_beta0 in place, A_g = amplitude^2) ----
  // (also published for k_gene / finish_body: shx = [sigma_beta0, sigma_amplitude, tail, sigma_alpha[k]], Aexp[g])
  if (tid < k) { const double e = exp(sh[5 + k + tid]); sal[tid] = e; if (blockIdx.x == 0) shx_o[3 + tid] = e; }
  else if (tid == k) { const double e = m.depth ? sh[1] : exp(sh[1]); sh[1] = e; if (blockIdx.x == 0) shx_o[0] = e; }  // sigma_beta0 | beta1
  else if (tid > k && tid <= k + gpb) {
    const int q = tid - k - 1;
    const double e = exp(2.0 * zg[q * npg + k + 1]);  // approxGP.stan:27
    Ag[q] = e;
    if ((int64_t)blockIdx.x * gpb + q < m.G) Aexp_o[(int64_t)blockIdx.x * gpb + q] = e;
  } else if (blockIdx.x == 0 && tid == k + gpb + 1) shx_o[1] = exp(sh[3]);
  else if (blockIdx.x == 0 && tid == k + gpb + 2) {
    const double e = exp(sh[4]);
    shx_o[2] = e;
    shx_o[3 + k] = m.lik == BVG_GAUSSIAN ? 1.0 / (e * e) : 1.0;
  }
  __syncthreads();
  // ---- phase 3: coefficient rows c'_g = (A_g alpha_t[,g], beta0_g) ----
  if (tc) {  // tcgen05 engine: publish the draw in fp32 (W rows, U vector); the likelihood kernel forms c'_g itself
    if (is_gene) W_o[g * 32 + j] = (float)(j == k + 1 ? Ag[gl] : zg[gl * npg + j]);
    if (blockIdx.x == 0 && tid < k) { U_o[tid] = (float)sh[5 + tid]; U_o[32 + tid] = (float)sal[tid]; }
    if (blockIdx.x == 0 && tid == k) { U_o[64] = (float)sh[0]; U_o[65] = (float)sh[1]; }
    return;
  }
  if (!is_gene || j == k + 1) return;
  const double v = j < k ? Ag[gl] * (sh[5 + j] + sal[j] * zg[gl * npg + j])  // A_g * alpha_t[j,g] (approxGP.stan:32)
                         : sh[0] + sh[1] * zg[gl * npg + k];                  // beta0 (approxGP.stan:26)
  ((T *)m.C)[g * m.KP + j] = (T)v;
}


// ------------------------------------------------------------------------------------------------
// k_prep_big: the same three phases for any k <= BVG_MAX_K (config 5: k ~ 400).  A block owns gpb genes and
// walks its parameters in strides of the block size.
template <typename T>
__global__ void __launch_bounds__(256) k_prep_big(DevModel m, int mode, int gpb, LsPoint ls) {
  extern __shared__ double sh[];  // [2k+5] shared zeta | [k] sigma_alpha | [gpb][k+2] gene zeta | [gpb] A_g
  griddep_wait();
  griddep_launch();
  const int k = m.k, nsh = 2 * k + 5, npg = k + 2, tid = threadIdx.x;
  double *sal = sh + nsh, *zg = sal + k, *Ag = zg + gpb * npg;
  const uint32_t stream = mode == PREP_GRAD ? STREAM_GRAD : (mode == PREP_DRAW ? STREAM_DRAW : STREAM_ELBO);
  const uint32_t t = mode == PREP_GRAD ? m.ctl->t_grad : (mode == PREP_DRAW ? m.ctl->t_draw : m.ctl->t_elbo);
  const int64_t gbase = (int64_t)blockIdx.x * gpb;
  for (int i = tid; i < nsh + gpb * npg; i += blockDim.x) {
    const bool is_sh = i < nsh;
    const int t2 = i - nsh, gl = is_sh ? 0 : t2 / npg, j = is_sh ? 0 : t2 - gl * npg;
    const int64_t g = gbase + gl;
    if (!is_sh && g >= m.G) continue;
    if (!is_sh && m.depth && j == k) { zg[gl * npg + k] = m.depths[g]; continue; }  // the intercept "z" is data (approxGP2.stan:46)
    int64_t li, gi;
    if (is_sh) { li = shared_index(m.L, k, i); gi = shared_index(m.LG, k, i); }
    else {
      const int64_t gg = m.gene_offset + g;
      if (j < k) { li = m.L.z_alpha + g * k + j; gi = m.LG.z_alpha + gg * k + j; }
      else if (j == k) { li = m.L.z_beta0 + g; gi = m.LG.z_beta0 + gg; }
      else { li = m.L.amp + g; gi = m.LG.amp + gg; }
    }
    double z;
    if (mode != PREP_NONE) {
      z = m.mu[li] + exp(m.omega[li]) * bvg_normal(m.seed, stream, t, (uint64_t)gi);
      if (!is_sh || blockIdx.x == 0) m.zeta[li] = z;
    } else if (ls.x0) {  // trial point of the device-resident line search (lbfgs_dev.cu)
      z = ls.x0[li] + *ls.a * ls.p[li];
      if (!is_sh || blockIdx.x == 0) m.zeta[li] = z;
    } else {
      z = m.zeta[li];
    }
    if (is_sh) sh[i] = z; else zg[gl * npg + j] = z;
  }
  __syncthreads();
  for (int i = tid; i < k + 3 + gpb; i += blockDim.x) {
    if (i < k) { const double e = exp(sh[5 + k + i]); sal[i] = e; if (blockIdx.x == 0) m.shx[3 + i] = e; }
    else if (i == k) { if (blockIdx.x == 0) m.shx[1] = exp(sh[3]); }
    else if (i == k + 1) {
      if (blockIdx.x == 0) { const double e = exp(sh[4]); m.shx[2] = e; m.shx[3 + k] = m.lik == BVG_GAUSSIAN ? 1.0 / (e * e) : 1.0; }
    } else if (i == k + 2) { /* sigma_beta0: written after the barrier below (sh[1] is still read as a log here) */ }
    else {
      const int q = i - k - 3;
      const double e = exp(2.0 * zg[q * npg + k + 1]);  // approxGP.stan:27
      Ag[q] = e;
      if (gbase + q < m.G) m.Aexp[gbase + q] = e;
    }
  }
  __syncthreads();
  if (tid == 0) { const double e = m.depth ? sh[1] : exp(sh[1]); sh[1] = e; if (blockIdx.x == 0) m.shx[0] = e; }  // sigma_beta0 | beta1
  __syncthreads();
  for (int i = tid; i < gpb * (k + 1); i += blockDim.x) {
    const int gl = i / (k + 1), j = i - gl * (k + 1);
    const int64_t g = gbase + gl;
    if (g >= m.G) continue;
    const double v = j < k ? Ag[gl] * (sh[5 + j] + sal[j] * zg[gl * npg + j])  // A_g * alpha_t[j,g] (approxGP.stan:32)
                           : sh[0] + sh[1] * zg[gl * npg + k];                  // beta0 (approxGP.stan:26)
    ((T *)m.C)[g * m.KP + j] = (T)v;
  }
}

int launch_prep(bvg_fit *f, int mode) {
  const int k = f->m.k;
  if (k > BVG_SMALL_K) {
    const int gpb = std::max(1, std::min(8, 512 / (k + 2)));
    const int grid = (int)((f->m.G + gpb - 1) / gpb);
    const size_t sh = sizeof(double) * (3 * k + 5 + gpb * (k + 3));
    if (f->cfg.precision == BVG_PREC_FP64) CUDA_TRY(launch_pdl(k_prep_big<double>, dim3(grid), dim3(256), sh, f->stream, f->m, mode, gpb, f->ls));
    else CUDA_TRY(launch_pdl(k_prep_big<float>, dim3(grid), dim3(256), sh, f->stream, f->m, mode, gpb, f->ls));
    f->st.kernel_launches++;
    return BVG_OK;
  }
  const int gpb = (256 - (2 * k + 5)) / (k + 2);
  const int grid = (int)((f->m.G + gpb - 1) / gpb);
  const size_t sh = sizeof(double) * (3 * k + 5 + gpb * (k + 3));
  if (f->cfg.precision == BVG_PREC_FP64) CUDA_TRY(launch_pdl(k_prep<double>, dim3(grid), dim3(256), sh, f->stream, f->m, mode, 0, gpb, f->ls));
  else CUDA_TRY(launch_pdl(k_prep<float>, dim3(grid), dim3(256), sh, f->stream, f->m, mode, (int)(f->engine == BVG_ENGINE_TCGEN05 && !f->big_tc), gpb, f->ls));
  f->st.kernel_launches++;
  return BVG_OK;
}

// B draws of one stream in ONE launch (tcgen05 engine, k <= BVG_SMALL_K): slice b of the buffers f->m.{zeta, shx, Aexp, W, U}
// currently point at receives draw (counter + b).  The caller points the descriptor at batch buffers first.
int launch_prep_batch(bvg_fit *f, int mode, int B) {
  const int k = f->m.k;
  if (k > BVG_SMALL_K || f->cfg.precision == BVG_PREC_FP64 || f->engine != BVG_ENGINE_TCGEN05 || f->big_tc) { bvg_set_error("batched draws need the fused tcgen05 engine"); return BVG_ERR_STATE; }
  const int gpb = (256 - (2 * k + 5)) / (k + 2);
  const int grid = (int)((f->m.G + gpb - 1) / gpb);
  const size_t sh = sizeof(double) * (3 * k + 5 + gpb * (k + 3));
  CUDA_TRY(launch_pdl(k_prep<float>, dim3(grid, B), dim3(256), sh, f->stream, f->m, mode, 1, gpb, LsPoint{nullptr, nullptr, nullptr}));
  f->st.kernel_launches++;
  return BVG_OK;
}

// ------------------------------------------------------------------------------------------------
// mean-field ADVI update of one parameter: grad_mu = g, grad_omega = g*eps*exp(omega)+1
// = g*(zeta - mu) + 1 (zeta = mu + exp(omega)*eps is the stored draw, so neither eps nor exp(omega) is
// recomputed),  s <- (it==1 ? g^2 : 0.9 s + 0.1 g^2),  x += eta/sqrt(it) * grad / (1 + sqrt(s))
__device__ inline void advi_update(const DevModel &m, int64_t li, double gv, bool fin, int it, double es) {
  const double mu = m.mu[li];
  const double gm = fin ? gv : 0.0;
  const double go = fin ? gv * (m.zeta[li] - mu) + 1.0 : 0.0;
  double hm = m.hist[li], ho = m.hist[m.L.D + li];
  if (it == 1) { hm += gm * gm; ho += go * go; }
  else { hm = 0.9 * hm + 0.1 * gm * gm; ho = 0.9 * ho + 0.1 * go * go; }
  m.hist[li] = hm;
  m.hist[m.L.D + li] = ho;
  m.mu[li] = mu + es * gm / (1.0 + sqrt(hm));
  m.omega[li] += es * go / (1.0 + sqrt(ho));
}

// stamps use the SM cycle counter: a %globaltimer read stalls the warp for about a microsecond (four of them
// in a row looked like a 3.6 us "load latency" in an earlier trace)
__device__ inline long long gtime() { return clock64(); }
#define DBG(i) do { if (m.dbg && threadIdx.x == 0) m.dbg[i] = gtime(); } while (0)
#define DBG0(i) do { if (m.dbg && threadIdx.x == 0 && blockIdx.x == 0) m.dbg[i] = gtime(); } while (0)

// ------------------------------------------------------------------------------------------------
// finish_body (one CTA, >= 128 threads).  phase bit 0: fixed-order reduction of the per-block
// partials (+ nb histogram terms) into tot[]; bit 1: log density, gradient / ADVI update of the 2k+5
// gene-shared parameters, counters.  A multi-GPU fit all-reduces tot[] between the two phases.
// ahead != 0 (ADVI update only): also draw the NEXT evaluation's zeta of the gene-shared parameters and
// publish their exponentials, so that no k_prep launch sits between this evaluation and the next.
__device__ void finish_body(const DevModel &m, int mode, int jac, int propto, int phase, int ahead, double *scratch = nullptr) {
  __shared__ double red[2][32];
  __shared__ double s_lp;
  const int k = m.k, NS = bvg_ns(k), tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const Layout &L = m.L;
  const double *zt = m.zeta;
  const int nwarps = blockDim.x >> 5;
  // The threads that own a gene-shared parameter (tid < 2k+5) fetch everything their update needs -- and draw
  // the next evaluation's normal -- up front; when this CTA also reduces (phase 3) their warps are left out of
  // the reduction, so those loads and the Philox / Box-Muller chain overlap it instead of following it.
  const bool big = 2 * k + 5 > (int)blockDim.x - 32;  // large basis: shared parameters are walked in a loop (no prefetch, no draw-ahead)
  const bool gthr = !big && (phase & 2) && (mode == MODE_GRAD_OUT || mode == MODE_ADVI_UPDATE) && tid < 2 * k + 5;
  const int gw = (2 * k + 5 + 31) >> 5;  // warps holding gradient threads
  const int w0 = (!big && (phase & 3) == 3 && (mode == MODE_GRAD_OUT || mode == MODE_ADVI_UPDATE) && nwarps >= gw + 4) ? gw : 0;
  int64_t li_sh = 0;
  double p_mu = 0, p_om = 0, p_hm = 0, p_ho = 0, p_zeta = 0, p_eps = 0;
  if (gthr) {
    li_sh = shared_index(L, k, tid);
    p_zeta = zt[li_sh];
    if (mode == MODE_ADVI_UPDATE) {
      p_mu = m.mu[li_sh]; p_om = m.omega[li_sh]; p_hm = m.hist[li_sh]; p_ho = m.hist[L.D + li_sh];
      if (ahead) p_eps = bvg_normal(m.seed, STREAM_GRAD, m.ctl->t_grad + 1, (uint64_t)shared_index(m.LG, k, tid));
    }
  }
  if (phase & 1) {
    // every reducing warp owns up to 4 columns and requests all their rows at once (nblk_gene <= 148 rows:
    // 5 per lane), then sums in a fixed order and finishes with a fixed shuffle tree -> deterministic
    const int warp_r = warp - w0, nwr = nwarps - w0;
    if (warp >= w0) {
      for (int cbase = 0; cbase < NS; cbase += 4 * nwr) {  // one pass when 2k+10 <= 4 * reducing warps (k <= 27 with 16 warps)
        double acc[4] = {0, 0, 0, 0};
        for (int b0 = lane; b0 < m.nblk_gene; b0 += 160) {
          double v[4][5];
#pragma unroll
          for (int ci = 0; ci < 4; ++ci)
#pragma unroll
            for (int q = 0; q < 5; ++q) {
              const int c = cbase + warp_r + ci * nwr, b = b0 + 32 * q;
              v[ci][q] = (c < NS && b < m.nblk_gene) ? __ldcg(&m.blk[(int64_t)c * m.nblk_gene + b]) : 0.0;
            }
#pragma unroll
          for (int ci = 0; ci < 4; ++ci)
#pragma unroll
            for (int q = 0; q < 5; ++q) acc[ci] += v[ci][q];
        }
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          const int c = cbase + warp_r + ci * nwr;
          const double t = warp_sum(acc[ci]);
          if (lane == 0 && c < NS) m.tot[c] = t;
        }
      }
    }
    __syncthreads();
    DBG(5);
    if (m.lik == BVG_NB) {
      // sum_sg lgamma(y+phi) - lgamma(phi) and its digamma derivative depend only on (y, phi):
      // evaluate on the histogram of distinct counts
      double lg = 0, dg = 0;
      if (m.hterm) {
        // tcgen05 engine: idle lanes of the likelihood kernel already evaluated n_c (lgamma(c + phi) - lgamma(phi)) and
        // the digamma analogue for every distinct count c; only the additions are left for this serial section
        for (int i = tid; i < m.nhist; i += blockDim.x) { lg += m.hterm[2 * i]; dg += m.hterm[2 * i + 1]; }
      } else {
        const double phi = m.shx[2], lg0 = lgamma(phi), dg0 = dev_digamma(phi);
        for (int i = tid; i < m.nhist; i += blockDim.x) {
          const double c = m.hist_val[i], n = m.hist_cnt[i];
          lg += n * (lgamma(c + phi) - lg0);
          dg += n * (dev_digamma(c + phi) - dg0);
        }
      }
      lg = warp_sum(lg); dg = warp_sum(dg);
      if (lane == 0) { red[0][warp] = lg; red[1][warp] = dg; }
      __syncthreads();
      if (tid == 0) {
        double a = 0, b = 0;
        for (int w = 0; w < nwarps; ++w) { a += red[0][w]; b += red[1][w]; }
        m.tot[S_LG] = a - m.lgam_y1;
        m.tot[S_DG] = b;
      }
      __syncthreads();
    }
  }
  if (!(phase & 2)) return;
  // multi-GPU, fused: the cross-gene sums of every gene shard meet here, inside the kernel that produced
  // them -- stores into the peers' exchange buffers over NVLink, no extra launch (bvg_internal.cuh)
  const bool loc = m.elbo_local && (mode == MODE_ELBO_DRAW || mode == MODE_POST_DRAW);  // local share of a sharded ELBO / posterior draw: no exchange
  if ((phase & 1) && m.peer.world > 1 && !loc) peer_allreduce_cta(m.peer, m.tot, NS, scratch);
  const double shw = (loc && m.rank_id != 0) ? 0.0 : 1.0;   // weight of the terms that involve gene-shared parameters only
  const double *tot = m.tot;
  const double mb = zt[L.mu_beta0], lsb = zt[L.sigma_beta0], sb = m.shx[0];
  const double ma = zt[L.mu_amp], lsa = zt[L.sigma_amp], sa = m.shx[1];
  const double lt = zt[L.tail], tl = m.shx[2];
  const double Gt = loc ? (double)m.G : (double)m.G_total, N = (double)m.M * Gt;
  const double j1 = jac ? 1.0 : 0.0;
  const int it = m.ctl->it;
  if (warp == (blockDim.x >> 5) - 1) {  // log density (last warp, concurrently with the gradient threads below)
    const double lp = bvg_lp_from_sums(m, tot, zt, m.shx, jac, propto, loc, shw, lane);
    if (lane == 0) s_lp = lp;
  } else if (big && (mode == MODE_GRAD_OUT || mode == MODE_ADVI_UPDATE)) {
    const double es = m.ctl->es;
    for (int i = tid; i < 2 * k + 5; i += blockDim.x - 32) {  // every warp but the last (which forms the log density)
      double gv;
      if (i == 0) gv = tot[S_R] - mb / 4.0;
      else if (i == 1) gv = m.depth ? tot[S_RZ] - sb / 4.0 : sb * (tot[S_RZ] - sb) + j1;
      else if (i == 2) gv = tot[S_DU] / (sa * sa) - ma / 4.0;
      else if (i == 3) gv = -Gt + tot[S_DU2] / (sa * sa) - sa * sa + j1;
      else if (i == 4) {
        if (m.lik == BVG_GAUSSIAN) gv = -N + tot[S_SSE] / (tl * tl) - (m.depth ? tl * tl / 4.0 : tl * tl) + j1;
        else gv = tl * (tot[S_DG] - tot[S_L] - tot[S_R] / tl) - 0.5 * tl * tl / (1.0 + tl * tl / 4.0) + j1;
      } else if (i < 5 + k) gv = tot[S_AP + (i - 5)] - zt[L.mu_alpha + (i - 5)] / 4.0;
      else {
        const int j = i - 5 - k;
        const double s = m.shx[3 + j];
        gv = s * (tot[S_AP + k + j] - s) + j1;
      }
      const int64_t li = shared_index(L, k, i);
      if (mode == MODE_GRAD_OUT) m.grad[li] = gv;
      else {
        const bool fin = isfinite(gv);
        if (!fin) atomicAdd(&m.ctl->bad, 1);
        advi_update(m, li, gv, fin, it, es);
      }
    }
  } else if ((mode == MODE_GRAD_OUT || mode == MODE_ADVI_UPDATE) && tid < 2 * k + 5) {
    const double es = m.ctl->es;
    const int i = tid;
    double gv;
    if (i == 0) gv = tot[S_R] - mb / 4.0;
    else if (i == 1) gv = m.depth ? tot[S_RZ] - sb / 4.0 : sb * (tot[S_RZ] - sb) + j1;
    else if (i == 2) gv = tot[S_DU] / (sa * sa) - ma / 4.0;
    else if (i == 3) gv = -Gt + tot[S_DU2] / (sa * sa) - sa * sa + j1;
    else if (i == 4) {
      if (m.lik == BVG_GAUSSIAN) gv = -N + tot[S_SSE] / (tl * tl) - (m.depth ? tl * tl / 4.0 : tl * tl) + j1;
      else gv = tl * (tot[S_DG] - tot[S_L] - tot[S_R] / tl) - 0.5 * tl * tl / (1.0 + tl * tl / 4.0) + j1;
    } else if (i < 5 + k) gv = tot[S_AP + (i - 5)] - zt[L.mu_alpha + (i - 5)] / 4.0;
    else {
      const int j = i - 5 - k;
      const double s = m.shx[3 + j];
      gv = s * (tot[S_AP + k + j] - s) + j1;
    }
    if (mode == MODE_GRAD_OUT) m.grad[li_sh] = gv;
    else {
      // mean-field update (advi_update) on the prefetched state, then the next evaluation's draw of this parameter
      const bool fin = isfinite(gv);
      if (!fin) atomicAdd(&m.ctl->bad, 1);
      const double gm = fin ? gv : 0.0, go = fin ? gv * (p_zeta - p_mu) + 1.0 : 0.0;
      if (it == 1) { p_hm += gm * gm; p_ho += go * go; }
      else { p_hm = 0.9 * p_hm + 0.1 * gm * gm; p_ho = 0.9 * p_ho + 0.1 * go * go; }
      m.hist[li_sh] = p_hm;
      m.hist[L.D + li_sh] = p_ho;
      p_mu += es * gm / (1.0 + sqrt(p_hm));
      p_om += es * go / (1.0 + sqrt(p_ho));
      m.mu[li_sh] = p_mu;
      m.omega[li_sh] = p_om;
    }
  }
  __syncthreads();
  DBG(6);
  if (ahead && !big && mode == MODE_ADVI_UPDATE && tid < 2 * k + 5) {
    // every thread of the CTA has finished reading this evaluation's shared zeta (barrier above)
    const double z = p_mu + exp(p_om) * p_eps;
    m.zeta[li_sh] = z;
    if (tid == 0) m.U[64] = (float)z;
    else if (tid == 1) { const double e = m.depth ? z : exp(z); m.shx[0] = e; m.U[65] = (float)e; }  // sigma_beta0 | beta1
    else if (tid == 3) m.shx[1] = exp(z);
    else if (tid == 4) { const double e = exp(z); m.shx[2] = e; m.shx[3 + k] = m.lik == BVG_GAUSSIAN ? 1.0 / (e * e) : 1.0; }
    else if (tid >= 5 + k) { const double e = exp(z); m.shx[3 + tid - 5 - k] = e; m.U[32 + tid - 5 - k] = (float)e; }
    else if (tid >= 5) m.U[tid - 5] = (float)z;
  }
  if (mode == MODE_ADVI_UPDATE && tid == (big ? 0 : 2 * k + 5)) m.ctl->es = m.ctl->eta / sqrt((double)(it + 1));
  if (tid == 0) {
    const double lp = s_lp;
    m.ctl->lp = lp;
    if (mode == MODE_ADVI_UPDATE) { m.ctl->t_grad += 1; m.ctl->it = it + 1; }
    else if (mode == MODE_ELBO_DRAW) {
      if (loc) { m.elbo_lps[m.ctl->elbo_ok] = lp; m.ctl->elbo_ok += 1; }  // local share; the host all-reduces the S slots
      else if (isfinite(lp)) { m.ctl->elbo_acc += lp; m.ctl->elbo_ok += 1; }
      m.ctl->t_elbo += 1;
    } else if (mode == MODE_POST_DRAW) {
      m.elbo_lps[m.ctl->t_draw] = lp;
      m.ctl->t_draw += 1;
    }
  }
}

__global__ void __launch_bounds__(512) k_finish(DevModel m, int mode, int jac, int propto, int phase, int ahead) {
  finish_body(m, mode, jac, propto, phase, ahead);
}

// ------------------------------------------------------------------------------------------------
// k_gene: one warp per gene (lane j < k: z_alpha_t[j,g]; lane k: z_beta0[g]; lane k+1: amplitude[g]).
// Everything a gene needs is requested up front (8 split partials in flight per lane, the lane's own
// zeta / mu / omega / step-size history) so the warp pays ~2 L2 round trips, not one per load.
// The last CTA to retire also runs finish_body (fuse = 3: everything; fuse = 1, multi-GPU: only the
// cross-block reduction, because the NCCL all-reduce of tot[] comes next), saving a launch per evaluation.
// ahead != 0 (ADVI update): every lane also draws its parameter's zeta for the NEXT evaluation right after
// the update (Philox counter t_grad + 1), so consecutive ADVI evaluations are two launches -- likelihood,
// k_gene -- instead of three; k_prep then only runs at the start of a batch of steps.
// tg_next = Philox counter of that next draw (the host mirrors the device counter): the normal of a warp's first
// gene is generated BEFORE the wait on the likelihood kernel, i.e. while its slower CTAs are still running.
// SPEC = 1 is the instantiation the ADVI loop runs (mode = ADVI update, jacobian on, hierarchical intercept): with those
// three fixed at compile time the kernel fits the 80 registers that 21 warps per CTA allow (warps are allocated in
// fours, so 21 count as 24) without spilling.
template <typename T, int SPEC>
__global__ void __launch_bounds__(BVG_GENE_MAX_THREADS) k_gene(DevModel m, int mode_rt, int jac_rt, int propto, int fuse, int ahead, uint32_t tg_next) {
  const int mode = SPEC ? (int)MODE_ADVI_UPDATE : mode_rt, jac = SPEC ? 1 : jac_rt;
  const bool dep = SPEC ? false : m.depth != 0;
  extern __shared__ double sm[];  // [nwarps][NS]
  __shared__ int s_last;
  const int k = m.k, KP = m.KP, NS = bvg_ns(k);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int64_t gfirst = (int64_t)blockIdx.x * nw + warp;
  auto global_index = [&](int64_t g) {
    const int64_t gg = m.gene_offset + g;
    return (uint64_t)(lane < k ? m.LG.z_alpha + gg * k + lane : (lane == k ? m.LG.z_beta0 + gg : m.LG.amp + gg));
  };
  const bool has = lane < k + 2 && !(dep && lane == k);  // depth-adjusted programs have no z_beta0: lane k carries gene_depths[g]
  double eps_first = 0.0;
  if (ahead && has && gfirst < m.G) eps_first = bvg_normal(m.seed, STREAM_GRAD, tg_next, global_index(gfirst));
  griddep_wait();    // Ppart / Spart come from the likelihood kernel
  griddep_launch();
  const Layout &L = m.L;
  const double *zt = m.zeta;
  const double sb = m.shx[0], ma = zt[L.mu_amp], sa = m.shx[1];  // exponentials precomputed by k_prep
  const double inv_s2 = m.shx[3 + k];  // 1 / sigma_y^2 (gaussian) | 1 (nb), published with the draw
  const double mual = lane < k ? zt[L.mu_alpha + lane] : 0.0;
  const double sal = lane < k ? m.shx[3 + lane] : 0.0;
  const int it = m.ctl->it;
  const double es = m.ctl->es;        // eta / sqrt(it)
  const bool upd = mode == MODE_ADVI_UPDATE;
  DBG0(0);
  double aR = 0, aRz = 0, aDu = 0, aDu2 = 0, aS0 = 0, aS1 = 0, aZ2 = 0, aU = 0, aAP = 0, aAPz = 0;
  const T *Pp = (const T *)m.Ppart;
  const double *Sp = (const double *)m.Spart;
  for (int64_t g = (int64_t)blockIdx.x * nw + warp; g < m.G; g += (int64_t)gridDim.x * nw) {
    const int64_t li = lane < k ? L.z_alpha + g * k + lane : (lane == k ? L.z_beta0 + g : L.amp + g);
    const double zi = has ? zt[li] : ((dep && lane == k) ? m.depths[g] : 0.0);  // this lane's own parameter value
    const double Aexp_g = m.Aexp[g];
    double mu_i = 0, om_i = 0, hm = 0, ho = 0;
    if (upd && has) { mu_i = m.mu[li]; om_i = m.omega[li]; hm = m.hist[li]; ho = m.hist[L.D + li]; }
    double Pj = 0, s0 = 0;
    for (int sp0 = 0; sp0 < m.nsplit; sp0 += 8) {  // fixed order -> deterministic
      // unconditional loads from clamped (always valid) addresses: all 16 go out in one round trip; the
      // predicated form was compiled into several dependent load -> convert rounds
      T pv[8];
      double sv[8];
      const int lp = lane < KP ? lane : KP - 1, ls = lane & 1;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int sp = sp0 + q < m.nsplit ? sp0 + q : m.nsplit - 1;
        const int64_t row = (int64_t)sp * m.Gpad + g;
        pv[q] = Pp[row * KP + lp];
        sv[q] = Sp[row * 2 + ls];
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const bool ok = sp0 + q < m.nsplit;
        Pj += (ok && lane < KP) ? (double)pv[q] : 0.0;
        s0 += (ok && lane < 2) ? sv[q] : 0.0;
      }
    }
    const double s1 = __shfl_sync(0xffffffffu, s0, 1);
    s0 = __shfl_sync(0xffffffffu, s0, 0);
    Pj *= inv_s2;
    const double R = __shfl_sync(0xffffffffu, Pj, k);  // ones column of Phi' -> R_g = sum_s r_sg
    const double zj = lane < k ? zi : 0.0;
    const double zb = __shfl_sync(0xffffffffu, zi, k), u = __shfl_sync(0xffffffffu, zi, k + 1);
    const double A = Aexp_g, alpha = mual + sal * zj;
    const double aP = warp_sum(lane < k ? alpha * Pj : 0.0);
    const double z2 = warp_sum(zj * zj) + (dep ? 0.0 : zb * zb);
    const double AP = lane < k ? A * Pj : 0.0;
    const double du = u - ma;
    aAP += AP; aAPz += AP * zj;
    aR += R; aRz += R * zb; aDu += du; aDu2 += du * du; aS0 += s0; aS1 += s1; aZ2 += z2; aU += u;
    if (mode == MODE_GRAD_OUT || upd) {
      double gv = 0;
      if (lane < k) gv = sal * AP - zj;
      else if (lane == k) gv = sb * R - zb;
      else if (lane == k + 1) gv = 2.0 * A * aP - du / (sa * sa) - (jac ? 0.0 : 1.0);
      if (!upd) {
        if (has) m.grad[li] = gv;
      } else {
        // mean-field ADVI update (see advi_update); a gene whose gradient is not finite keeps its parameters
        const bool fin = __all_sync(0xffffffffu, has ? isfinite(gv) : true);
        if (!fin && lane == 0) atomicAdd(&m.ctl->bad, 1);
        if (has) {
          const double gm = fin ? gv : 0.0, go = fin ? gv * (zi - mu_i) + 1.0 : 0.0;
          if (it == 1) { hm += gm * gm; ho += go * go; }
          else { hm = 0.9 * hm + 0.1 * gm * gm; ho = 0.9 * ho + 0.1 * go * go; }
          m.hist[li] = hm;
          m.hist[L.D + li] = ho;
          const double mu_n = mu_i + es * gm / (1.0 + sqrt(hm)), om_n = om_i + es * go / (1.0 + sqrt(ho));
          m.mu[li] = mu_n;
          m.omega[li] = om_n;
          if (ahead) {  // next evaluation's draw of this parameter (what k_prep would compute)
            const double eps = g == gfirst ? eps_first : bvg_normal(m.seed, STREAM_GRAD, tg_next, global_index(g));
            const double zn = mu_n + exp(om_n) * eps;
            m.zeta[li] = zn;
            if (lane == k + 1) {
              const double An = exp(2.0 * zn);  // approxGP.stan:27
              m.Aexp[g] = An;
              m.W[g * 32 + lane] = (float)An;
            } else {
              m.W[g * 32 + lane] = (float)zn;
            }
          }
        }
      }
    }
  }
  DBG0(2);
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
    m.blk[(int64_t)c * gridDim.x + blockIdx.x] = s;  // column-major [NS][nblk]: finish_body reads it coalesced
  }
  DBG0(3);
  if (!fuse) return;
  // last-CTA-done: the barrier orders this CTA's blk[] stores before thread 0's RELEASE on the ticket (release is
  // cumulative), and the last CTA's ACQUIRE on the same ticket + its barrier make every CTA's stores visible to all
  // of its threads -- no full-strength __threadfence (MEMBAR.SC) on either side
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned ticket;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(ticket) : "l"(&m.ctl->ticket) : "memory");
    s_last = ticket == gridDim.x - 1;
    if (s_last) m.ctl->ticket = 0;
  }
  __syncthreads();
  if (!s_last) return;
  DBG(4);
  // fuse = 3: reduce + (exchange) + finish; fuse = 1 (NCCL): reduce only, the all-reduce follows.  The warps' smem
  // rows are free by now and serve as the exchange's polling scratch when they are large enough.
  finish_body(m, mode, jac, propto, fuse, ahead, (m.peer.world - 1) * NS <= nw * NS ? sm : nullptr);
  DBG(7);
}


// ------------------------------------------------------------------------------------------------
// k_gene_big: k_gene for any k <= BVG_MAX_K.  One warp per gene; the k+2 parameters of the gene are walked in
// chunks of 32 lanes.  The amplitude gradient needs sum_j alpha_j P_j over ALL chunks, so lane 0 updates it after
// the walk.  The per-basis cross-gene sums (S_AP slots) accumulate in the warp's shared-memory row.
template <typename T>
__global__ void __launch_bounds__(512) k_gene_big(DevModel m, int mode, int jac, int propto, int fuse) {
  extern __shared__ double sm[];  // [nwarps][NS]
  __shared__ int s_last;
  griddep_wait();
  griddep_launch();
  const int k = m.k, KP = m.KP, NS = bvg_ns(k);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const Layout &L = m.L;
  const double *zt = m.zeta;
  const double sb = m.shx[0], ma = zt[L.mu_amp], sa = m.shx[1], inv_s2 = m.shx[3 + k];
  const int it = m.ctl->it;
  const double es = m.ctl->es;
  const bool upd = mode == MODE_ADVI_UPDATE, wr = mode == MODE_GRAD_OUT || upd;
  double *row = sm + warp * NS;
  for (int c = lane; c < NS; c += 32) row[c] = 0.0;
  __syncwarp();
  double aR = 0, aRz = 0, aDu = 0, aDu2 = 0, aS0 = 0, aS1 = 0, aZ2 = 0, aU = 0;
  const T *Pp = (const T *)m.Ppart;
  const double *Sp = (const double *)m.Spart;
  for (int64_t g = (int64_t)blockIdx.x * nw + warp; g < m.G; g += (int64_t)gridDim.x * nw) {
    const double A = m.Aexp[g];
    double s0 = 0, s1 = 0;
    for (int sp = lane; sp < m.nsplit; sp += 32) {
      const int64_t r = (int64_t)sp * m.Gpad + g;
      s0 += (double)Sp[r * 2];
      s1 += (double)Sp[r * 2 + 1];
    }
    // lanes hold different splits: a fixed shuffle tree keeps the sum deterministic
    s0 = warp_sum(s0); s1 = warp_sum(s1);
    double aP = 0, z2 = 0, R = 0, zb = 0, u = 0;
    for (int j0 = 0; j0 < k + 2; j0 += 32) {
      const int j = j0 + lane;
      if (j >= k + 2) continue;
      const int64_t li = j < k ? L.z_alpha + g * k + j : (j == k ? L.z_beta0 + g : L.amp + g);
      const bool datum = m.depth && j == k;  // depth-adjusted programs: gene_depths[g] in place of z_beta0[g]
      const double z = datum ? m.depths[g] : zt[li];
      if (j == k + 1) { u = z; continue; }
      double Pj = 0;
      for (int sp0 = 0; sp0 < m.nsplit; sp0 += 8) {  // 8 independent loads per round trip, summed in a fixed order
        T pv[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int sp = sp0 + q < m.nsplit ? sp0 + q : m.nsplit - 1;
          pv[q] = Pp[((int64_t)sp * m.Gpad + g) * KP + j];
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) Pj += sp0 + q < m.nsplit ? (double)pv[q] : 0.0;
      }
      Pj *= inv_s2;
      double gv;
      if (j < k) {
        const double sal = m.shx[3 + j], alpha = zt[L.mu_alpha + j] + sal * z, AP = A * Pj;
        aP += alpha * Pj;
        z2 += z * z;
        row[S_AP + j] += AP;
        row[S_AP + k + j] += AP * z;
        gv = sal * AP - z;
      } else {  // j == k: the ones column of Phi' -> R_g = sum_s r_sg
        R = Pj; zb = z;
        gv = sb * R - zb;
        if (datum) continue;
      }
      if (mode == MODE_GRAD_OUT) m.grad[li] = gv;
      else if (upd) {
        const bool fin = isfinite(gv);
        if (!fin) atomicAdd(&m.ctl->bad, 1);
        advi_update(m, li, gv, fin, it, es);
      }
    }
    aP = warp_sum(aP); z2 = warp_sum(z2); R = warp_sum(R); zb = warp_sum(zb); u = warp_sum(u);
    if (!m.depth) z2 += zb * zb;
    const double du = u - ma;
    aR += R; aRz += R * zb; aDu += du; aDu2 += du * du; aS0 += s0; aS1 += s1; aZ2 += z2; aU += u;
    if (wr && lane == 0) {
      const int64_t li = L.amp + g;
      const double gv = 2.0 * A * aP - du / (sa * sa) - (jac ? 0.0 : 1.0);
      if (mode == MODE_GRAD_OUT) m.grad[li] = gv;
      else {
        const bool fin = isfinite(gv);
        if (!fin) atomicAdd(&m.ctl->bad, 1);
        advi_update(m, li, gv, fin, it, es);
      }
    }
  }
  __syncwarp();
  if (lane == 0) {
    row[S_R] = aR; row[S_RZ] = aRz; row[S_DU] = aDu; row[S_DU2] = aDu2; row[S_SSE] = aS0; row[S_Z2] = aZ2;
    row[S_U] = aU; row[S_L] = aS1; row[S_LG] = 0; row[S_DG] = 0;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < NS; c += blockDim.x) {
    double s = 0;
    for (int w = 0; w < nw; ++w) s += sm[w * NS + c];
    m.blk[(int64_t)c * gridDim.x + blockIdx.x] = s;
  }
  if (!fuse) return;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned ticket;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(ticket) : "l"(&m.ctl->ticket) : "memory");
    s_last = ticket == gridDim.x - 1;
    if (s_last) m.ctl->ticket = 0;
  }
  __syncthreads();
  if (!s_last) return;
  finish_body(m, mode, jac, propto, fuse, 0);
}

// function attributes are per device: called from finalize_model (after cudaSetDevice) for every large-basis model
int gene_big_setup() {
  CUDA_TRY(cudaFuncSetAttribute(k_gene_big<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  CUDA_TRY(cudaFuncSetAttribute(k_gene_big<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  return BVG_OK;
}

int g_gene_fuse_override = -1;  // diagnostics (bvg_debug_time_gene): force the fuse argument
int launch_gene_finish(bvg_fit *f, int mode, int jac, int propto, int ahead) {
  const int wpb = f->gene_wpb;
  const size_t sh = sizeof(double) * wpb * bvg_ns(f->m.k);
  // single GPU or peer-memory exchange: everything in one launch; NCCL (or 2k+10 > CTA size): reduce, all-reduce, finish
  const bool split = f->comm_mode == 1 || (f->comm_mode == 2 && bvg_ns(f->m.k) > wpb * 32);
  const int fuse = g_gene_fuse_override >= 0 ? g_gene_fuse_override : (split ? 1 : 3);
  const uint32_t tg_next = f->h_t_grad + 1;
  if (f->m.k > BVG_SMALL_K) {
    // large basis: 16 gene-warps per CTA ([warps][2k+10] shared-memory rows: 104 KB at k = 400); no draw-ahead
    const int wb = 16;
    const size_t shb = sizeof(double) * wb * bvg_ns(f->m.k);
    const bool splitb = f->comm_mode != 0;  // 2k+10 sums exceed one CTA's exchange width: reduce, all-reduce, finish
    const int fuseb = g_gene_fuse_override >= 0 ? g_gene_fuse_override : (splitb ? 1 : 3);
    if (f->cfg.precision == BVG_PREC_FP64) CUDA_TRY(launch_pdl(k_gene_big<double>, dim3(f->m.nblk_gene), dim3(wb * 32), shb, f->stream, f->m, mode, jac, propto, fuseb));
    else CUDA_TRY(launch_pdl(k_gene_big<float>, dim3(f->m.nblk_gene), dim3(wb * 32), shb, f->stream, f->m, mode, jac, propto, fuseb));
    f->st.kernel_launches++;
    if (fuseb == 1) {
      if (!(f->m.elbo_local && (mode == MODE_ELBO_DRAW || mode == MODE_POST_DRAW))) BVG_TRY(comm_allreduce(f, f->m.tot, bvg_ns(f->m.k)));
      k_finish<<<1, 512, 0, f->stream>>>(f->m, mode, jac, propto, 2, 0);
      f->st.kernel_launches += 1;
      CUDA_TRY(cudaGetLastError());
    }
    if (mode == MODE_ADVI_UPDATE) f->h_t_grad++;
    return BVG_OK;
  }
  const bool spec = mode == MODE_ADVI_UPDATE && jac == 1 && !f->m.depth;
  if (f->cfg.precision == BVG_PREC_FP64) {
    if (spec) CUDA_TRY(launch_pdl(k_gene<double, 1>, dim3(f->m.nblk_gene), dim3(wpb * 32), sh, f->stream, f->m, mode, jac, propto, fuse, ahead, tg_next));
    else CUDA_TRY(launch_pdl(k_gene<double, 0>, dim3(f->m.nblk_gene), dim3(wpb * 32), sh, f->stream, f->m, mode, jac, propto, fuse, ahead, tg_next));
  } else {
    if (spec) CUDA_TRY(launch_pdl(k_gene<float, 1>, dim3(f->m.nblk_gene), dim3(wpb * 32), sh, f->stream, f->m, mode, jac, propto, fuse, ahead, tg_next));
    else CUDA_TRY(launch_pdl(k_gene<float, 0>, dim3(f->m.nblk_gene), dim3(wpb * 32), sh, f->stream, f->m, mode, jac, propto, fuse, ahead, tg_next));
  }
  f->st.kernel_launches++;
  if (fuse == 1) {
    if (!(f->m.elbo_local && (mode == MODE_ELBO_DRAW || mode == MODE_POST_DRAW))) BVG_TRY(comm_allreduce(f, f->m.tot, bvg_ns(f->m.k)));
    k_finish<<<1, 512, 0, f->stream>>>(f->m, mode, jac, propto, 2, ahead);
    f->st.kernel_launches += 1;
    CUDA_TRY(cudaGetLastError());
  }
  if (mode == MODE_ADVI_UPDATE) f->h_t_grad++;  // mirrors ctl->t_grad (incremented by the finish of every ADVI evaluation)
  return BVG_OK;
}
