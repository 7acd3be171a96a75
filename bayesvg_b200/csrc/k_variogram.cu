// k_variogram.cu -- lscale.estimator = "variogram" (R/findSpatiallyVariableFeaturesBayes.R:208-260): for every gene an empirical
// semivariogram of its expression over the scaled spot coordinates (gstat::variogram(z ~ 1): cutoff = bounding-box diagonal / 3,
// 15 distance classes) and a weighted least-squares fit of nugget + Matern(partial sill, range; kappa = kernel.smoothness)
// (gstat::fit.variogram, weights N_h / h^2 [U]); the length-scale is the median fitted range.
//
// The reference runs this gene by gene in a doSNOW pool -- M^2 / 2 point pairs per gene through gstat's C code.  Here the pair
// geometry is shared by all genes: a warp = 32 consecutive genes of one spot i walks j > i, computes the pair's distance class once
// and loads the 32 genes' values of spot j in one coalesced 128-byte request (expr is genes x spots column-major, i.e. spot-major,
// exactly as R holds expr_mtx).  Per-class sums live in private shared-memory slots and are reduced in a fixed order: deterministic.
// The 3-parameter fits (15 points each) run one gene per thread (Levenberg-Marquardt with the analytic Jacobian).
#include <algorithm>
#include <cmath>
#include <vector>

#include "bvg_internal.cuh"

#define VG_NB 15        // distance classes (gstat default: cutoff / 15)
#define VG_IY 8         // spots i per CTA pass
#define VG_FIX 1099511627776.0   // 2^40: fixed-point distance sums (deterministic atomics)

// class counts and distance sums (gene independent): np[b], sum_h[b] as 2^-40 fixed point
__global__ void __launch_bounds__(256) k_vg_classes(const double *__restrict__ X, int64_t M, double cutoff, double width,
                                                    unsigned long long *__restrict__ out /* [2][VG_NB] */) {
  __shared__ unsigned long long s[2 * VG_NB];
  if (threadIdx.x < 2 * VG_NB) s[threadIdx.x] = 0ull;
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (int64_t)gridDim.x * blockDim.x) {
    const double xi = X[i], yi = X[M + i];
    for (int64_t j = i + 1; j < M; ++j) {
      const double dx = xi - X[j], dy = yi - X[M + j];
      const double h = sqrt(fma(dx, dx, dy * dy));
      if (h > 0.0 && h <= cutoff) {
        int b = (int)floor(h / width);
        b = b >= VG_NB ? VG_NB - 1 : b;
        atomicAdd(&s[b], 1ull);
        atomicAdd(&s[VG_NB + b], (unsigned long long)llrint(h * VG_FIX));
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < 2 * VG_NB && s[threadIdx.x]) atomicAdd(&out[threadIdx.x], s[threadIdx.x]);
}

// part[chunk][g][b] = sum over the chunk's spots i and all j > i in class b of (z_gi - z_gj)^2;  Z is [M][G] (spot-major)
__global__ void __launch_bounds__(32 * VG_IY) k_vg_gamma(const double *__restrict__ X, const double *__restrict__ Z, int64_t M, int64_t G,
                                                         double cutoff, double width, int64_t i_per_chunk, double *__restrict__ part) {
  __shared__ double acc[VG_IY][VG_NB][32];
  const int lane = threadIdx.x & 31, iy = threadIdx.x >> 5;
  const int64_t g = (int64_t)blockIdx.x * 32 + lane;
  const bool live = g < G;
#pragma unroll
  for (int b = 0; b < VG_NB; ++b) acc[iy][b][lane] = 0.0;
  const int64_t i0 = (int64_t)blockIdx.y * i_per_chunk, i1 = i0 + i_per_chunk < M ? i0 + i_per_chunk : M;
  for (int64_t i = i0 + iy; i < i1; i += VG_IY) {
    const double xi = X[i], yi = X[M + i];
    const double zi = live ? Z[(size_t)i * G + g] : 0.0;
    for (int64_t j = i + 1; j < M; ++j) {
      const double dx = xi - X[j], dy = yi - X[M + j];
      const double h = sqrt(fma(dx, dx, dy * dy));   // the same expression as k_vg_classes: identical class membership
      if (h > 0.0 && h <= cutoff) {   // uniform over the warp: every lane shares (i, j)
        int b = (int)floor(h / width);
        b = b >= VG_NB ? VG_NB - 1 : b;
        const double d = zi - (live ? Z[(size_t)j * G + g] : 0.0);
        acc[iy][b][lane] = fma(d, d, acc[iy][b][lane]);
      }
    }
  }
  __syncthreads();
  if (iy == 0 && live) {
    double *o = part + ((size_t)blockIdx.y * G + g) * VG_NB;
#pragma unroll
    for (int b = 0; b < VG_NB; ++b) {
      double s = 0;
#pragma unroll
      for (int q = 0; q < VG_IY; ++q) s += acc[q][b][lane];   // fixed order
      o[b] = s;
    }
  }
}

__device__ inline void vg_rho(double x, int kind, double *rho, double *drho) {  // Matern correlation in h / range and its derivative
  const double e = exp(-x);
  if (kind == 0) { *rho = e; *drho = -e; }                                        // kappa = 0.5
  else if (kind == 1) { *rho = (1.0 + x) * e; *drho = -x * e; }                   // 1.5
  else { *rho = (1.0 + x + x * x / 3.0) * e; *drho = -(x + x * x) / 3.0 * e; }    // 2.5
}
__device__ inline double vg_wss(const double *h, const double *gam, const double *w, int n, int kind, const double *p) {
  double s = 0;
  for (int q = 0; q < n; ++q) {
    double r, dr;
    vg_rho(h[q] / p[2], kind, &r, &dr);
    const double e = gam[q] - (p[0] + p[1] * (1.0 - r));
    s += w[q] * e * e;
  }
  return s;
}
// one gene per thread: gamma_b = sum / (2 np_b), weights np_b / dist_b^2, start (0.2 var, 0.8 var, median dist) (:241-245)
__global__ void k_vg_fit(const double *__restrict__ part, int nchunk, int64_t G, const double *__restrict__ cls /* [2][VG_NB]: np, mean dist */,
                         const double *__restrict__ Z, int64_t M, int kind, double *__restrict__ ranges) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  double h[VG_NB], gam[VG_NB], w[VG_NB], hs[VG_NB];
  int n = 0;
  for (int b = 0; b < VG_NB; ++b) {
    if (cls[b] <= 0) continue;   // empty classes are dropped
    double s = 0;
    for (int c = 0; c < nchunk; ++c) s += part[((size_t)c * G + g) * VG_NB + b];
    h[n] = cls[VG_NB + b];
    gam[n] = s / (2.0 * cls[b]);
    w[n] = cls[b] / (h[n] * h[n]);
    hs[n] = h[n];
    ++n;
  }
  // stats::var(gene_expr) and median(emp_vario$dist)
  double mean = 0;
  for (int64_t s = 0; s < M; ++s) mean += Z[(size_t)s * G + g];
  mean /= (double)M;
  double var = 0;
  for (int64_t s = 0; s < M; ++s) { const double d = Z[(size_t)s * G + g] - mean; var += d * d; }
  var /= (double)(M - 1);
  for (int a = 1; a < n; ++a) { const double v = hs[a]; int q = a - 1; while (q >= 0 && hs[q] > v) { hs[q + 1] = hs[q]; --q; } hs[q + 1] = v; }
  double p[3] = {0.2 * var, 0.8 * var, n ? ((n & 1) ? hs[n / 2] : 0.5 * (hs[n / 2 - 1] + hs[n / 2])) : NAN};
  if (n < 3 || !(var > 0.0)) { ranges[g] = NAN; return; }
  double lam = 1e-3, f = vg_wss(h, gam, w, n, kind, p);
  bool ok = isfinite(f);
  for (int it = 0; it < 200 && ok; ++it) {
    double A[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, gr[3] = {0, 0, 0};
    for (int q = 0; q < n; ++q) {
      double r, dr;
      const double x = h[q] / p[2];
      vg_rho(x, kind, &r, &dr);
      const double J[3] = {1.0, 1.0 - r, p[1] * dr * x / p[2]};   // d model / d (nugget, psill, range)
      const double e = gam[q] - (p[0] + p[1] * (1.0 - r));
      for (int a = 0; a < 3; ++a) {
        gr[a] += w[q] * J[a] * e;
        for (int c = 0; c < 3; ++c) A[a][c] += w[q] * J[a] * J[c];
      }
    }
    bool stepped = false;
    for (int tries = 0; tries < 30; ++tries) {
      double B[3][4];
      for (int a = 0; a < 3; ++a) { for (int c = 0; c < 3; ++c) B[a][c] = A[a][c] + (a == c ? lam * A[a][a] : 0.0); B[a][3] = gr[a]; }
      bool sing = false;
      for (int a = 0; a < 3 && !sing; ++a) {   // Gaussian elimination with partial pivoting
        int pv = a;
        for (int r2 = a + 1; r2 < 3; ++r2) if (fabs(B[r2][a]) > fabs(B[pv][a])) pv = r2;
        if (!(fabs(B[pv][a]) > 0.0)) { sing = true; break; }
        if (pv != a) for (int c = 0; c < 4; ++c) { const double t = B[a][c]; B[a][c] = B[pv][c]; B[pv][c] = t; }
        for (int r2 = a + 1; r2 < 3; ++r2) { const double m2 = B[r2][a] / B[a][a]; for (int c = a; c < 4; ++c) B[r2][c] -= m2 * B[a][c]; }
      }
      double dlt[3] = {0, 0, 0};
      if (!sing) {
        for (int a = 2; a >= 0; --a) { double s = B[a][3]; for (int c = a + 1; c < 3; ++c) s -= B[a][c] * dlt[c]; dlt[a] = s / B[a][a]; }
      }
      const double q3[3] = {p[0] + dlt[0], p[1] + dlt[1], p[2] + dlt[2]};
      const double fn = (!sing && q3[2] > 0.0) ? vg_wss(h, gam, w, n, kind, q3) : INFINITY;
      if (isfinite(fn) && fn <= f) {
        const double rel = (f - fn) / (f > 0 ? f : 1.0);
        p[0] = q3[0]; p[1] = q3[1]; p[2] = q3[2];
        f = fn;
        lam = lam * 0.3 > 1e-12 ? lam * 0.3 : 1e-12;
        stepped = true;
        if (rel < 1e-12) it = 1000;   // converged
        break;
      }
      lam *= 10.0;
      if (lam > 1e12) break;
    }
    if (!stepped) break;
  }
  ranges[g] = (ok && isfinite(p[2]) && p[2] > 0.0) ? p[2] : NAN;
}

extern "C" int bvg_variogram_lscale(int device, const double *coords, int64_t M, const double *expr, int64_t G, double kappa,
                                    double *ranges_out, double *lscale_out) {
  if (!coords || !expr || !lscale_out) { bvg_set_error("null argument"); return BVG_ERR_ARG; }
  if (M < 3 || G < 1) { bvg_set_error("variogram: need at least 3 spots and 1 gene"); return BVG_ERR_ARG; }
  const int kind = kappa == 0.5 ? 0 : (kappa == 1.5 ? 1 : (kappa == 2.5 ? 2 : -1));
  if (kind < 0) { bvg_set_error("variogram: kernel.smoothness must be 0.5, 1.5 or 2.5 (closed-form Matern variogram models)"); return BVG_ERR_ARG; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); bvg_set_error("no CUDA device: bayesvg_b200 has no CPU fallback"); return BVG_ERR_CUDA; }
  CUDA_TRY(cudaSetDevice(device));
  // gstat::variogram defaults: cutoff = diagonal of the bounding box / 3, width = cutoff / 15
  double lo[2] = {INFINITY, INFINITY}, hi[2] = {-INFINITY, -INFINITY};
  for (int64_t s = 0; s < M; ++s)
    for (int c = 0; c < 2; ++c) { const double v = coords[(size_t)c * M + s]; if (!std::isfinite(v)) { bvg_set_error("variogram: non-finite coordinate"); return BVG_ERR_ARG; } lo[c] = std::min(lo[c], v); hi[c] = std::max(hi[c], v); }
  const double cutoff = std::sqrt((hi[0] - lo[0]) * (hi[0] - lo[0]) + (hi[1] - lo[1]) * (hi[1] - lo[1])) / 3.0, width = cutoff / VG_NB;
  if (!(cutoff > 0)) { bvg_set_error("variogram: all spots coincide"); return BVG_ERR_ARG; }
  const int64_t i_per_chunk = std::max<int64_t>(VG_IY, (M + 63) / 64);
  const int nchunk = (int)((M + i_per_chunk - 1) / i_per_chunk);
  char *d = nullptr;
  const size_t bytes = sizeof(double) * (2 * (size_t)M + (size_t)M * G + (size_t)nchunk * G * VG_NB + 2 * VG_NB + (size_t)G) + sizeof(unsigned long long) * 2 * VG_NB;
  CUDA_TRY(cudaMalloc(&d, bytes));
  double *dX = (double *)d, *dZ = dX + 2 * M, *dPart = dZ + (size_t)M * G, *dCls = dPart + (size_t)nchunk * G * VG_NB, *dR = dCls + 2 * VG_NB;
  unsigned long long *dCnt = (unsigned long long *)(dR + G);
  cudaStream_t st = nullptr;
  int rc = BVG_OK;
  std::vector<double> ranges((size_t)G), cls(2 * VG_NB);
  unsigned long long hc[2 * VG_NB];
#define VG_CHECK(x) if ((x) != cudaSuccess) { bvg_set_error("bvg_variogram_lscale: %s", cudaGetErrorString(cudaGetLastError())); rc = BVG_ERR_CUDA; goto done; }
  VG_CHECK(cudaStreamCreate(&st));
  VG_CHECK(cudaMemcpyAsync(dX, coords, sizeof(double) * 2 * M, cudaMemcpyHostToDevice, st));
  VG_CHECK(cudaMemcpyAsync(dZ, expr, sizeof(double) * (size_t)M * G, cudaMemcpyHostToDevice, st));
  VG_CHECK(cudaMemsetAsync(dCnt, 0, sizeof(unsigned long long) * 2 * VG_NB, st));
  k_vg_classes<<<(int)std::min<int64_t>((M + 255) / 256, 148 * 8), 256, 0, st>>>(dX, M, cutoff, width, dCnt);
  k_vg_gamma<<<dim3((unsigned)((G + 31) / 32), (unsigned)nchunk), 32 * VG_IY, 0, st>>>(dX, dZ, M, G, cutoff, width, i_per_chunk, dPart);
  VG_CHECK(cudaMemcpyAsync(hc, dCnt, sizeof(hc), cudaMemcpyDeviceToHost, st));
  VG_CHECK(cudaStreamSynchronize(st));
  for (int b = 0; b < VG_NB; ++b) { cls[b] = (double)hc[b]; cls[VG_NB + b] = hc[b] ? (double)hc[VG_NB + b] / VG_FIX / (double)hc[b] : 0.0; }
  VG_CHECK(cudaMemcpyAsync(dCls, cls.data(), sizeof(double) * 2 * VG_NB, cudaMemcpyHostToDevice, st));
  k_vg_fit<<<(int)((G + 127) / 128), 128, 0, st>>>(dPart, nchunk, G, dCls, dZ, M, kind, dR);
  VG_CHECK(cudaMemcpyAsync(ranges.data(), dR, sizeof(double) * G, cudaMemcpyDeviceToHost, st));
  VG_CHECK(cudaStreamSynchronize(st));
  VG_CHECK(cudaGetLastError());
  {
    if (ranges_out) std::copy(ranges.begin(), ranges.end(), ranges_out);
    std::vector<double> v;
    for (double r : ranges) if (std::isfinite(r)) v.push_back(r);   // median(vario_ranges, na.rm = TRUE) (:260)
    if (v.empty()) { bvg_set_error("variogram: no gene produced a fitted range"); rc = BVG_ERR_NUMERIC; goto done; }
    std::sort(v.begin(), v.end());
    *lscale_out = (v.size() & 1) ? v[v.size() / 2] : 0.5 * (v[v.size() / 2 - 1] + v[v.size() / 2]);
  }
done:
  if (st) cudaStreamDestroy(st);
  cudaFree(d);
  return rc;
}
