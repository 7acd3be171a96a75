// k_summary.cu -- posterior draws of amplitude[g] and their summary statistics, one CTA per gene.
//
// Replaces mod$variational(..., draws = n.draws) + fit_vi$summary(variables = "amplitude") of the
// reference (R/findSpatiallyVariableFeaturesBayes.R:357-364, :419-432): under the mean-field
// approximation amplitude[g] = exp(mu_u + exp(omega_u) * eps); the n.draws draws are summarised with
// posterior::default_summary_measures() semantics (mean, median, sd with n-1, mad = 1.4826 *
// median|x - med|, q5 / q95 with quantile type 7), then amplitude_dispersion = sd^2 / mean (:428).
// The rank column (:429-430) is filled on the host in bvg_fit_get_summary.
#include "bvg_internal.cuh"

__device__ inline void bitonic_sort(double *a, int n) {  // n power of two, whole CTA
  for (int k = 2; k <= n; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const double x = a[i], y = a[ixj];
          const bool up = (i & k) == 0;
          if ((x > y) == up) { a[i] = y; a[ixj] = x; }
        }
      }
      __syncthreads();
    }
}
__device__ inline double quantile7(const double *s, int n, double p) {
  const double h = (n - 1) * p;
  const int lo = (int)floor(h), hi = lo + 1 < n ? lo + 1 : lo;
  return s[lo] + (h - lo) * (s[hi] - s[lo]);
}
__device__ inline double block_sum(double v, double *red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
  return s;
}

__global__ void __launch_bounds__(256) k_summary(DevModel m, int nd, int npad, double *__restrict__ out, double *__restrict__ draws, int given) {
  extern __shared__ double sm[];
  __shared__ double red[8];
  double *x = sm, *d = sm + npad;
  const int64_t g = blockIdx.x, li = m.L.amp + g;
  const uint64_t gi = (uint64_t)(m.LG.amp + m.gene_offset + g);
  const double mu = m.mu[li], sd_u = exp(m.omega[li]);
  double acc = 0, nbad = 0;
  for (int t = threadIdx.x; t < npad; t += blockDim.x) {
    double v = INFINITY;
    if (t < nd) {
      v = given ? draws[g * nd + t]   // full-rank family: the draws were made by k_fullrank.cu
                : exp(mu + sd_u * bvg_normal(m.seed, STREAM_DRAW, (uint32_t)t, gi));
      acc += v;
      if (!isfinite(v)) nbad += 1.0;
      if (draws && !given) draws[g * nd + t] = v;
    }
    x[t] = v;
  }
  // a non-finite draw (overflowing exp, NaN mu / omega) would collide with the +inf padding and break the
  // comparisons of the sorting network: such a gene gets NaN statistics instead of silently wrong quantiles
  if (block_sum(nbad, red) > 0.0) {
    if (threadIdx.x < 7) out[threadIdx.x * m.G + g] = NAN;
    if (threadIdx.x == 7) out[7 * m.G + g] = 0.0;
    return;
  }
  const double mean = block_sum(acc, red) / nd;
  acc = 0;
  for (int t = threadIdx.x; t < nd; t += blockDim.x) acc += (x[t] - mean) * (x[t] - mean);
  const double sd = sqrt(block_sum(acc, red) / (nd - 1));
  __syncthreads();
  bitonic_sort(x, npad);
  const double med = quantile7(x, nd, 0.5);
  for (int t = threadIdx.x; t < npad; t += blockDim.x) d[t] = t < nd ? fabs(x[t] - med) : INFINITY;
  __syncthreads();
  bitonic_sort(d, npad);
  if (threadIdx.x == 0) {
    const int64_t G = m.G;
    out[0 * G + g] = mean;
    out[1 * G + g] = med;
    out[2 * G + g] = sd;
    out[3 * G + g] = 1.4826 * quantile7(d, nd, 0.5);
    out[4 * G + g] = quantile7(x, nd, 0.05);
    out[5 * G + g] = quantile7(x, nd, 0.95);
    out[6 * G + g] = sd * sd / mean;
    out[7 * G + g] = 0.0;
  }
}

int launch_summary(bvg_fit *f, bool given) {
  const int nd = f->cfg.n_draws;
  int npad = 2;
  while (npad < nd) npad <<= 1;
  const size_t sh = sizeof(double) * 2 * npad;
  if (nd < 2 || sh > 200 * 1024) { bvg_set_error("n_draws = %d unsupported (2..8192)", nd); return BVG_ERR_ARG; }
  CUDA_TRY(cudaFuncSetAttribute(k_summary, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
  k_summary<<<(unsigned)f->m.G, 256, sh, f->stream>>>(f->m, nd, npad, f->summary, f->draws, given ? 1 : 0);
  f->st.kernel_launches++;
  CUDA_TRY(cudaGetLastError());
  return BVG_OK;
}
