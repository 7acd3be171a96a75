// k_vec.cu -- small vector kernels: data staging (fp64/int32 host layout -> device layout),
// masked reductions, and the device side of the vector-free L-BFGS (lbfgs.cu).
#include "bvg_vec.cuh"

// ---- staging ------------------------------------------------------------------------------------
template <typename SRC, typename DST>
__global__ void k_stage_y(const SRC *__restrict__ src, DST *__restrict__ dst, int64_t M, int64_t ldY, int64_t ngenes) {
  // src: [ngenes][M] contiguous (R column-major M x G slice); dst: [ngenes][ldY], zero padded
  const int64_t n = ngenes * ldY;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t g = i / ldY, s = i - g * ldY;
    dst[i] = s < M ? (DST)src[g * M + s] : (DST)0;
  }
}
template <typename SRC, typename DST>
int stage_y(const SRC *src, DST *dst, int64_t M, int64_t ldY, int64_t ngenes, cudaStream_t st) {
  const int64_t n = ngenes * ldY;
  const int grid = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
  k_stage_y<SRC, DST><<<grid > 0 ? grid : 1, 256, 0, st>>>(src, dst, M, ldY, ngenes);
  CUDA_TRY(cudaGetLastError());
  return BVG_OK;
}
template int stage_y<double, double>(const double *, double *, int64_t, int64_t, int64_t, cudaStream_t);
template int stage_y<double, float>(const double *, float *, int64_t, int64_t, int64_t, cudaStream_t);
template int stage_y<int32_t, double>(const int32_t *, double *, int64_t, int64_t, int64_t, cudaStream_t);
template int stage_y<int32_t, float>(const int32_t *, float *, int64_t, int64_t, int64_t, cudaStream_t);

// tile-major layout for the tcgen05 engine: dst[((tile*nch32 + c32)*128 + r)*32 + i], gene = 128*tile + r,
// spot = 32*c32 + i (dst is zero-filled beforehand: padded genes / spots stay zero)
template <typename SRC>
__global__ void k_stage_y_tiled(const SRC *__restrict__ src, float *__restrict__ dst, int64_t M, int64_t ldY, int64_t g0, int64_t ngenes) {
  const int64_t n = ngenes * M, nch32 = ldY / 32;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t gl = i / M, s = i - gl * M, g = g0 + gl;
    const int64_t tile = g / BVG_GENE_TILE, r = g - tile * BVG_GENE_TILE, c32 = s / 32, w = s - c32 * 32;
    dst[((tile * nch32 + c32) * BVG_GENE_TILE + r) * 32 + w] = (float)src[i];
  }
}
template <typename SRC>
int stage_y_tiled(const SRC *src, float *dst, int64_t M, int64_t ldY, int64_t g0, int64_t ngenes, cudaStream_t st) {
  k_stage_y_tiled<SRC><<<148 * 16, 256, 0, st>>>(src, dst, M, ldY, g0, ngenes);
  CUDA_TRY(cudaGetLastError());
  return BVG_OK;
}
template int stage_y_tiled<double>(const double *, float *, int64_t, int64_t, int64_t, int64_t, cudaStream_t);
template int stage_y_tiled<int32_t>(const int32_t *, float *, int64_t, int64_t, int64_t, int64_t, cudaStream_t);

// Phi (M x k col-major fp64) -> Phi' [Mpad][KP] row-major with the ones column at index k
template <typename T>
__global__ void k_stage_phi(const double *__restrict__ phi, T *__restrict__ out, int64_t M, int64_t Mpad, int k, int KP) {
  const int64_t n = Mpad * KP;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = i / KP;
    const int j = (int)(i - s * KP);
    T v = 0;
    if (s < M) v = j < k ? (T)phi[(int64_t)j * M + s] : (j == k ? (T)1 : (T)0);
    out[i] = v;
  }
}
template <typename T>
int stage_phi(const double *phi, T *out, int64_t M, int64_t Mpad, int k, int KP, cudaStream_t st) {
  k_stage_phi<T><<<148 * 4, 256, 0, st>>>(phi, out, M, Mpad, k, KP);
  CUDA_TRY(cudaGetLastError());
  return BVG_OK;
}
template int stage_phi<double>(const double *, double *, int64_t, int64_t, int, int, cudaStream_t);
template int stage_phi<float>(const double *, float *, int64_t, int64_t, int, int, cudaStream_t);

// ---- reductions ---------------------------------------------------------------------------------
// out[r*3 + q] = sum_i B[r][i] * v_q[i] * w[i]   for the three vectors v = (B[ra], B[rb], B[rc]);
// one block per (row r, chunk), second pass adds chunks in fixed order.
__global__ void __launch_bounds__(256) k_multidot3(const double *__restrict__ B, int64_t D, int nrows, int ra, int rb, int rc,
                                                   const double *__restrict__ w, double *__restrict__ part, int nchunk) {
  __shared__ double red[3][8];
  const int r = blockIdx.x, ch = blockIdx.y;
  const int64_t per = (D + nchunk - 1) / nchunk, i0 = ch * per, i1 = (i0 + per < D) ? i0 + per : D;
  const double *b = B + (int64_t)r * D, *va = B + (int64_t)ra * D, *vb = B + (int64_t)rb * D, *vc = B + (int64_t)rc * D;
  double a0 = 0, a1 = 0, a2 = 0;
  for (int64_t i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    const double x = b[i] * w[i];
    a0 += x * va[i]; a1 += x * vb[i]; a2 += x * vc[i];
  }
  a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a0; red[1][threadIdx.x >> 5] = a1; red[2][threadIdx.x >> 5] = a2; }
  __syncthreads();
  if (threadIdx.x < 3) {
    double s = 0;
    for (int q = 0; q < 8; ++q) s += red[threadIdx.x][q];
    part[((int64_t)r * nchunk + ch) * 3 + threadIdx.x] = s;
  }
}
__global__ void k_chunk_sum(const double *__restrict__ part, double *__restrict__ out, int n, int nchunk, int width) {
  // out[r*width + q] = sum_ch part[(r*nchunk + ch)*width + q]
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * width) return;
  const int r = i / width, q = i - r * width;
  double s = 0;
  for (int ch = 0; ch < nchunk; ++ch) s += part[((int64_t)r * nchunk + ch) * width + q];
  out[i] = s;
}
int vec_multidot3(cudaStream_t st, const double *B, int64_t D, int nrows, int ra, int rb, int rc, const double *w,
                  double *part, double *out) {
  const int nchunk = VEC_NCHUNK;
  k_multidot3<<<dim3(nrows, nchunk), 256, 0, st>>>(B, D, nrows, ra, rb, rc, w, part, nchunk);
  k_chunk_sum<<<(nrows * 3 + 127) / 128, 128, 0, st>>>(part, out, nrows, nchunk, 3);
  CUDA_TRY(cudaGetLastError());
  return BVG_OK;
}

// out[0] = sum a*b*w, out[1] = sum a*w   (single block: D is O(1e5))
__global__ void __launch_bounds__(1024) k_dot_sum(const double *__restrict__ a, const double *__restrict__ b,
                                                  const double *__restrict__ w, int64_t D, double *__restrict__ out) {
  __shared__ double red[2][32];
  double s0 = 0, s1 = 0;
  for (int64_t i = threadIdx.x; i < D; i += blockDim.x) {
    const double x = a[i] * w[i];
    s0 += x * (b ? b[i] : 0.0);
    s1 += x;
  }
  s0 = warp_sum(s0); s1 = warp_sum(s1);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s0; red[1][threadIdx.x >> 5] = s1; }
  __syncthreads();
  if (threadIdx.x < 2) {
    double s = 0;
    for (int q = 0; q < 32; ++q) s += red[threadIdx.x][q];
    out[threadIdx.x] = s;
  }
}
int vec_dot_sum(cudaStream_t st, const double *a, const double *b, const double *w, int64_t D, double *out) {
  k_dot_sum<<<1, 1024, 0, st>>>(a, b, w, D, out);
  CUDA_TRY(cudaGetLastError());
  return BVG_OK;
}

// ---- L-BFGS vector updates ----------------------------------------------------------------------
__global__ void k_axpy_to(double *__restrict__ out, const double *__restrict__ x, double a, const double *__restrict__ p, int64_t D) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < D; i += (int64_t)gridDim.x * blockDim.x) out[i] = x[i] + a * p[i];
}
int vec_axpy_to(cudaStream_t st, double *out, const double *x, double a, const double *p, int64_t D) {
  k_axpy_to<<<(int)((D + 255) / 256), 256, 0, st>>>(out, x, a, p, D);
  CUDA_TRY(cudaGetLastError());
  return BVG_OK;
}
// accept a step: s = x1 - x0, y = (-grad1) - g0, then x0 <- x1, g0 <- -grad1 (g is the gradient of f = -lp)
__global__ void k_accept(double *__restrict__ srow, double *__restrict__ yrow, double *__restrict__ x0, double *__restrict__ g0,
                         const double *__restrict__ x1, const double *__restrict__ grad1, int64_t D) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < D; i += (int64_t)gridDim.x * blockDim.x) {
    const double xn = x1[i], gn = -grad1[i];
    if (srow) { srow[i] = xn - x0[i]; yrow[i] = gn - g0[i]; }
    x0[i] = xn;
    g0[i] = gn;
  }
}
int vec_accept(cudaStream_t st, double *srow, double *yrow, double *x0, double *g0, const double *x1, const double *grad1, int64_t D) {
  k_accept<<<(int)((D + 255) / 256), 256, 0, st>>>(srow, yrow, x0, g0, x1, grad1, D);
  CUDA_TRY(cudaGetLastError());
  return BVG_OK;
}
// p = sum_r coef[r] * B[r]
__global__ void k_lincomb(double *__restrict__ p, const double *__restrict__ B, int64_t D, int nrows, LinCoef coef) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < D; i += (int64_t)gridDim.x * blockDim.x) {
    double s = 0;
    for (int r = 0; r < nrows; ++r)
      if (coef.d[r] != 0.0) s += coef.d[r] * B[(int64_t)r * D + i];
    p[i] = s;
  }
}
int vec_lincomb(cudaStream_t st, double *p, const double *B, int64_t D, int nrows, const LinCoef &coef) {
  k_lincomb<<<(int)((D + 255) / 256), 256, 0, st>>>(p, B, D, nrows, coef);
  CUDA_TRY(cudaGetLastError());
  return BVG_OK;
}
__global__ void k_fill(double *x, double v, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] = v;
}
int vec_fill(cudaStream_t st, double *x, double v, int64_t n) {
  k_fill<<<(int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184), 256, 0, st>>>(x, v, n);
  CUDA_TRY(cudaGetLastError());
  return BVG_OK;
}
// weight vector: 1 for per-gene parameters; for the 2k+5 gene-shared parameters 1 on rank 0 and 0
// elsewhere, so that all-reduced dot products count them once
__global__ void k_make_weights(double *w, DevModel m, int rank) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m.L.D; i += (int64_t)gridDim.x * blockDim.x) w[i] = 1.0;
}
__global__ void k_mask_shared(double *w, DevModel m) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 2 * m.k + 5) w[shared_index(m.L, m.k, i)] = 0.0;
}
int vec_make_weights(cudaStream_t st, double *w, const DevModel &m, int rank) {
  k_make_weights<<<(int)((m.L.D + 255) / 256), 256, 0, st>>>(w, m, rank);
  if (rank != 0) k_mask_shared<<<1, 128, 0, st>>>(w, m);
  CUDA_TRY(cudaGetLastError());
  return BVG_OK;
}
