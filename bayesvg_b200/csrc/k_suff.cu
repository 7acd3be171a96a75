// k_suff.cu -- forward-only ELBO estimates from per-gene sufficient statistics (gaussian likelihood).
//
// With Phi' = [Phi, 1] and c'_g = (A_g alpha_g, beta0_g) the gaussian residual sum of squares is the
// quadratic form
//     SSE_g = sum_s y_sg^2 - 2 c'_g . (Phi'^T y_g) + c'_g^T (Phi'^T Phi') c'_g          
// so log_prob<propto=false, jacobian=true>(zeta), which is all stan::variational::advi::calc_ELBO needs
// (150 draws every 100 iterations; R/findSpatiallyVariableFeaturesBayes.R:357-364), costs O(k^2 G) per
// draw instead of a dense O(MkG) pass.  The statistics are computed once, in fp64, from the Y the dense
// kernels read.  All S draws of one estimate run in TWO launches.  The gradient path never uses this.
#include "bvg_internal.cuh"

__device__ inline double blk_sum256(double v, double *red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
  return s;
}

// one CTA per gene: U[g][j] = sum_s Phi'_sj y_sg (j <= k), S2[g] = sum_s y^2
template <typename T>
__global__ void __launch_bounds__(256) k_suff_stats(const T *__restrict__ Y, int64_t ldY, int tiled, const double *__restrict__ phi,
                                                    int64_t M, int k, double *__restrict__ U, double *__restrict__ S2) {
  __shared__ double red[8];
  const int64_t g = blockIdx.x;
  const int64_t tile = g / BVG_GENE_TILE, r = g - tile * BVG_GENE_TILE, nch32 = ldY / 32;
  double acc[BVG_SMALL_K + 2];
  for (int j = 0; j < k + 2; ++j) acc[j] = 0;
  for (int64_t s = threadIdx.x; s < M; s += blockDim.x) {
    const int64_t yi = tiled ? ((tile * nch32 + s / 32) * BVG_GENE_TILE + r) * 32 + (s & 31) : g * ldY + s;
    const double y = (double)Y[yi];
    for (int j = 0; j < k; ++j) acc[j] += phi[(int64_t)j * M + s] * y;
    acc[k] += y;
    acc[k + 1] += y * y;
  }
  for (int j = 0; j < k + 2; ++j) {
    const double t = blk_sum256(acc[j], red);
    if (threadIdx.x == 0) { if (j <= k) U[g * (k + 1) + j] = t; else S2[g] = t; }
  }
}
// Gram[a][b] = sum_s Phi'_sa Phi'_sb, one CTA per entry
__global__ void __launch_bounds__(256) k_suff_gram(const double *__restrict__ phi, int64_t M, int k, double *__restrict__ Gram) {
  __shared__ double red[8];
  const int a = blockIdx.x / (k + 1), b = blockIdx.x % (k + 1);
  double s = 0;
  for (int64_t i = threadIdx.x; i < M; i += blockDim.x) {
    const double pa = a < k ? phi[(int64_t)a * M + i] : 1.0, pb = b < k ? phi[(int64_t)b * M + i] : 1.0;
    s += pa * pb;
  }
  s = blk_sum256(s, red);
  if (threadIdx.x == 0) Gram[blockIdx.x] = s;
}

// grid (gene blocks, S draws).  Same draw / constrain phases as k_prep (stream = ELBO, draw t_elbo + d),
// then the quadratic form per gene; block partials part[(d*nblk + b)*4 + {SSE, Z2, U, DU2}]
__global__ void __launch_bounds__(256) k_suff_elbo(DevModel m, int gpb, double *__restrict__ part) {
  extern __shared__ double sh[];  // [2k+5] shared zeta | [k] sigma_alpha | [gpb][k+2] gene zeta | [gpb] A_g | [gpb][k+1] c | [gpb][k+1] terms
  const int k = m.k, nsh = 2 * k + 5, npg = k + 2, K1 = k + 1, tid = threadIdx.x;
  double *sal = sh + nsh, *zg = sal + k, *Ag = zg + gpb * npg, *cc = Ag + gpb, *tm = cc + gpb * K1;
  const uint32_t t = m.ctl->t_elbo + blockIdx.y;
  const int t2 = tid - nsh, gl = t2 >= 0 ? t2 / npg : -1, j = t2 >= 0 ? t2 - gl * npg : 0;
  const int64_t g = (int64_t)blockIdx.x * gpb + gl;
  const bool is_sh = tid < nsh, is_gene = gl >= 0 && gl < gpb && g < m.G;
  if (is_sh || is_gene) {
    int64_t li, gi;
    if (is_sh) { li = shared_index(m.L, k, tid); gi = shared_index(m.LG, k, tid); }
    else {
      const int64_t gg = m.gene_offset + g;
      if (j < k) { li = m.L.z_alpha + g * k + j; gi = m.LG.z_alpha + gg * k + j; }
      else if (j == k) { li = m.L.z_beta0 + g; gi = m.LG.z_beta0 + gg; }
      else { li = m.L.amp + g; gi = m.LG.amp + gg; }
    }
    const double z = m.mu[li] + exp(m.omega[li]) * bvg_normal(m.seed, STREAM_ELBO, t, (uint64_t)gi);
    if (is_sh) sh[tid] = z; else zg[gl * npg + j] = z;
  }
  __syncthreads();
  if (tid < k) sal[tid] = exp(sh[5 + k + tid]);
  else if (tid == k) sh[1] = exp(sh[1]);
  else if (tid > k && tid <= k + gpb) { const int q = tid - k - 1; Ag[q] = exp(2.0 * zg[q * npg + k + 1]); }
  __syncthreads();
  if (is_gene && j <= k)
    cc[gl * K1 + j] = j < k ? Ag[gl] * (sh[5 + j] + sal[j] * zg[gl * npg + j]) : sh[0] + sh[1] * zg[gl * npg + k];
  __syncthreads();
  // quadratic form: thread (gene, j) computes c_j * (sum_i Gram[j][i] c_i - 2 U[g][j])
  if (is_gene && j <= k) {
    double q = 0;
    for (int i = 0; i < K1; ++i) q += m.suffGram[j * K1 + i] * cc[gl * K1 + i];
    tm[gl * K1 + j] = cc[gl * K1 + j] * (q - 2.0 * m.suffU[g * K1 + j]);
  }
  __syncthreads();
  // one thread per gene adds its terms in a fixed order, then 4 threads add the genes of the block
  double *gs = cc;  // reuse: [gpb][4]
  double sse = 0, z2 = 0, u = 0, du2 = 0;
  const int64_t gq = (int64_t)blockIdx.x * gpb + tid;
  if (tid < gpb && gq < m.G) {
    sse = m.suffS2[gq];
    for (int i = 0; i < K1; ++i) { sse += tm[tid * K1 + i]; const double z = zg[tid * npg + i]; z2 += z * z; }
    u = zg[tid * npg + k + 1];
    const double du = u - sh[2];
    du2 = du * du;
  }
  __syncthreads();
  if (tid < gpb) { gs[tid * 4 + 0] = sse; gs[tid * 4 + 1] = z2; gs[tid * 4 + 2] = u; gs[tid * 4 + 3] = du2; }
  __syncthreads();
  if (tid < 4) {
    double s = 0;
    for (int q = 0; q < gpb; ++q) s += gs[q * 4 + tid];
    part[((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * 4 + tid] = s;
  }
}

// one CTA per draw.  phase 1: tot4[d][4] = fixed-order sum of the block partials; phase 2: lp_d =
// log_prob<propto=false, jacobian=true> from the (all-reduced) sums and the draw's shared parameters
__global__ void __launch_bounds__(128) k_suff_finish(DevModel m, int nblk, const double *__restrict__ part, double *__restrict__ tot4,
                                                     double *__restrict__ lps, int phase) {
  __shared__ double sz[2 * BVG_SMALL_K + 5];
  __shared__ double red[4];
  const int k = m.k, d = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (phase & 1) {  // warp q sums column q
    double s = 0;
    for (int b0 = lane; b0 < nblk; b0 += 256) {
      double v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) { const int b = b0 + 32 * q; v[q] = b < nblk ? part[((int64_t)d * nblk + b) * 4 + warp] : 0.0; }
#pragma unroll
      for (int q = 0; q < 8; ++q) s += v[q];
    }
    s = warp_sum(s);
    if (lane == 0) tot4[d * 4 + warp] = s;
    __syncthreads();
  }
  if (!(phase & 2)) return;
  const uint32_t t = m.ctl->t_elbo + d;
  if (tid < 2 * k + 5) {
    const int64_t li = shared_index(m.L, k, tid);
    sz[tid] = m.mu[li] + exp(m.omega[li]) * bvg_normal(m.seed, STREAM_ELBO, t, (uint64_t)shared_index(m.LG, k, tid));
  }
  __syncthreads();
  double pa = 0;
  if (tid < k) { const double a = sz[5 + tid], ls = sz[5 + k + tid], s = exp(ls); pa = -a * a / 8.0 - 0.5 * s * s + ls; }
  pa = warp_sum(pa);
  if (lane == 0) red[warp] = pa;
  __syncthreads();
  if (tid == 0) {
    pa = red[0] + red[1] + red[2] + red[3];
    const double mb = sz[0], lsb = sz[1], sb = exp(lsb), ma = sz[2], lsa = sz[3], sa = exp(lsa), lt = sz[4], tl = exp(lt);
    const double Gt = (double)m.G_total, N = (double)m.M * Gt;
    const double *T = tot4 + d * 4;  // SSE, Z2, U, DU2
    double lp = -N * lt - 0.5 * T[0] / (tl * tl) - 0.5 * tl * tl;
    lp += -0.5 * T[1] - mb * mb / 8.0 - 0.5 * sb * sb - ma * ma / 8.0 - 0.5 * sa * sa;
    lp += -Gt * lsa - T[2] - 0.5 * T[3] / (sa * sa) + pa;
    lp += lsb + lsa + T[2] + lt;
    lp += -0.5 * BVG_LOG_TWO_PI * (N + (double)k * Gt + 2.0 * Gt + 2.0 * k + 5.0) - (k + 2.0) * BVG_LOG_TWO;
    lps[d] = lp;
  }
}

// ------------------------------------------------------------------------------------------------
struct SuffState { double *U, *S2, *Gram, *part, *tot4, *lps; int gpb, nblk, cap; };

static int suff_build(bvg_fit *f, int S) {
  DevModel &m = f->m;
  const int k = m.k, K1 = k + 1;
  const int gpb = (256 - (2 * k + 5)) / (k + 2), nblk = (int)((m.G + gpb - 1) / gpb);
  if (f->suff) {
    SuffState *st = (SuffState *)f->suff;
    if (st->cap >= S) return BVG_OK;
    cudaFree(st->U);
    delete st;
    f->suff = nullptr;
  }
  SuffState *st = new SuffState();
  st->gpb = gpb; st->nblk = nblk; st->cap = S;
  const size_t n = (size_t)m.G * K1 + m.G + (size_t)K1 * K1 + (size_t)S * nblk * 4 + (size_t)S * 4 + S;
  CUDA_TRY(cudaMalloc(&st->U, sizeof(double) * n));
  st->S2 = st->U + (size_t)m.G * K1; st->Gram = st->S2 + m.G; st->part = st->Gram + (size_t)K1 * K1;
  st->tot4 = st->part + (size_t)S * nblk * 4; st->lps = st->tot4 + (size_t)S * 4;
  if (f->cfg.precision == BVG_PREC_FP64)
    k_suff_stats<double><<<(unsigned)m.G, 256, 0, f->stream>>>((const double *)f->Y, f->ldY, 0, f->phi64, m.M, k, st->U, st->S2);
  else
    k_suff_stats<float><<<(unsigned)m.G, 256, 0, f->stream>>>((const float *)f->Y, f->ldY, f->tiledY ? 1 : 0, f->phi64, m.M, k, st->U, st->S2);
  k_suff_gram<<<K1 * K1, 256, 0, f->stream>>>(f->phi64, m.M, k, st->Gram);
  CUDA_TRY(cudaGetLastError());
  f->suff = (double *)st;
  m.suffU = st->U; m.suffS2 = st->S2; m.suffGram = st->Gram;
  return BVG_OK;
}

void suff_destroy(bvg_fit *f) {
  if (!f->suff) return;
  SuffState *st = (SuffState *)f->suff;
  cudaFree(st->U);
  delete st;
  f->suff = nullptr;
  f->m.suffU = f->m.suffS2 = f->m.suffGram = nullptr;
}

// S log densities log_prob<false, true>(zeta_d), d = 0..S-1, draws t_elbo .. t_elbo+S-1 (the caller advances t_elbo)
int suff_elbo(bvg_fit *f, int S, double *lps_host) {
  if (f->cfg.likelihood != BVG_GAUSSIAN) { bvg_set_error("sufficient-statistic ELBO exists for the gaussian likelihood only"); return BVG_ERR_ARG; }
  if (f->cfg.depth_adjust) { bvg_set_error("sufficient-statistic ELBO is not available for the depth-adjusted model"); return BVG_ERR_ARG; }
  if (f->m.k > BVG_SMALL_K) { bvg_set_error("sufficient-statistic ELBO supports n.basis.fns <= %d", BVG_SMALL_K); return BVG_ERR_ARG; }
  BVG_TRY(suff_build(f, S));
  SuffState *st = (SuffState *)f->suff;
  const DevModel &m = f->m;
  const int k = m.k;
  const size_t sh = sizeof(double) * (3 * k + 5 + st->gpb * (k + 2) + st->gpb + 2 * st->gpb * (k + 1) + 8);
  k_suff_elbo<<<dim3(st->nblk, S), 256, sh, f->stream>>>(m, st->gpb, st->part);
  if (f->comm_mode) {
    k_suff_finish<<<S, 128, 0, f->stream>>>(m, st->nblk, st->part, st->tot4, st->lps, 1);
    BVG_TRY(comm_allreduce(f, st->tot4, 4 * S));
    k_suff_finish<<<S, 128, 0, f->stream>>>(m, st->nblk, st->part, st->tot4, st->lps, 2);
    f->st.kernel_launches += 3;
  } else {
    k_suff_finish<<<S, 128, 0, f->stream>>>(m, st->nblk, st->part, st->tot4, st->lps, 3);
    f->st.kernel_launches += 2;
  }
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(lps_host, st->lps, sizeof(double) * S, cudaMemcpyDeviceToHost, f->stream));
  CUDA_TRY(cudaStreamSynchronize(f->stream));
  return BVG_OK;
}
