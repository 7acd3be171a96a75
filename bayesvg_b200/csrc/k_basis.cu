// k_basis.cu -- basis builder on the device, fp64.
//
// Replaces R/findSpatiallyVariableFeaturesBayes.R:265-288 of the reference:
//   phi[, i] <- kernel(rowSums((X - c_i)^2), lscale, ...)   (R/expQuadKernel.R:17, R/maternKernel.R:22-26,
//                                                            R/periodicKernel.R:20; d is the SQUARED distance)
//   phi_ortho <- qr.Q(qr(phi, LAPACK = TRUE))                 (column-pivoted Householder, dgeqp3 + dorgqr)
// One CTA per column; each Householder step is two launches (pivot + reflector, then trailing update
// with dlaqp2's partial-norm downdate), with no host synchronisation inside the factorisation.
#include <float.h>

#include "bvg_internal.cuh"

__device__ inline double kern_eval(double d2, double ls, int kernel, double nu, double period) {
  if (kernel == BVG_KERNEL_EXP_QUAD) return exp(-d2 / (2.0 * ls * ls));
  if (kernel == BVG_KERNEL_PERIODIC) {
    const double sn = sin(M_PI * sqrt(d2) / period);
    return exp(-2.0 * sn * sn / (ls * ls));
  }
  // Matern nu in {1/2, 3/2, 5/2}: closed forms of (2^(1-nu)/Gamma(nu)) s^nu K_nu(s); s = 0 -> 1
  const double s = sqrt(2.0 * nu) * sqrt(d2) / ls;
  if (nu == 0.5) return exp(-s);
  if (nu == 1.5) return (1.0 + s) * exp(-s);
  return (1.0 + s + s * s / 3.0) * exp(-s);
}

__global__ void k_kernel_matrix(const double *__restrict__ X, int64_t M, const double *__restrict__ C, int k, double ls,
                                int kernel, double nu, double period, double *__restrict__ A) {
  const int64_t n = M * k;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = i / M, s = i - j * M;
    const double dx = X[s] - C[j], dy = X[M + s] - C[k + j];
    A[i] = kern_eval(dx * dx + dy * dy, ls, kernel, nu, period);
  }
}

__device__ inline double blk_sum(double v, double *red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
  return s;
}

__global__ void __launch_bounds__(512) k_col_norms(const double *__restrict__ A, int64_t M, double *vn1, double *vn2) {
  __shared__ double red[16];
  const double *a = A + (int64_t)blockIdx.x * M;
  double s = 0;
  for (int64_t i = threadIdx.x; i < M; i += blockDim.x) s += a[i] * a[i];
  s = sqrt(blk_sum(s, red));
  if (threadIdx.x == 0) vn1[blockIdx.x] = vn2[blockIdx.x] = s;
}

// step j: choose the pivot (first column of largest partial norm), swap it into place, build the
// reflector (dlarfg): beta = -sign(alpha) ||x||, tau = (beta - alpha)/beta, v = x / (alpha - beta)
__global__ void __launch_bounds__(1024) k_qr_pivot_house(double *__restrict__ A, int64_t M, int k, int j, double *vn1,
                                                          double *vn2, double *tau) {
  __shared__ double red[32];
  __shared__ int s_pv;
  if (threadIdx.x == 0) {
    int pv = j;
    for (int c = j + 1; c < k; ++c) if (vn1[c] > vn1[pv]) pv = c;
    s_pv = pv;
  }
  __syncthreads();
  const int pv = s_pv;
  double *a = A + (int64_t)j * M;
  if (pv != j) {
    double *b = A + (int64_t)pv * M;
    for (int64_t i = threadIdx.x; i < M; i += blockDim.x) { const double t = a[i]; a[i] = b[i]; b[i] = t; }
    if (threadIdx.x == 0) { vn1[pv] = vn1[j]; vn2[pv] = vn2[j]; }
  }
  __syncthreads();
  double s = 0;
  for (int64_t i = j + 1 + threadIdx.x; i < M; i += blockDim.x) s += a[i] * a[i];
  const double xn = sqrt(blk_sum(s, red));
  const double alpha = a[j];
  __syncthreads();
  if (xn == 0.0) { if (threadIdx.x == 0) tau[j] = 0.0; return; }
  const double beta = -copysign(hypot(alpha, xn), alpha);
  const double sc = 1.0 / (alpha - beta);
  for (int64_t i = j + 1 + threadIdx.x; i < M; i += blockDim.x) a[i] *= sc;
  if (threadIdx.x == 0) { tau[j] = (beta - alpha) / beta; a[j] = beta; }
}

// apply H_j = I - tau v v' (v_j = 1) to column c = j+1+blockIdx.x of `T` using reflector column j of A.
// downdate != 0: T is A itself and the partial column norms are downdated (dlaqp2).
__global__ void __launch_bounds__(512) k_qr_apply(const double *__restrict__ A, double *__restrict__ T, int64_t M, int j, int c0,
                                                  const double *__restrict__ tau, double *vn1, double *vn2, int downdate) {
  __shared__ double red[16];
  __shared__ int s_recompute;
  const int c = c0 + blockIdx.x;
  const double *a = A + (int64_t)j * M;
  double *b = T + (int64_t)c * M;
  double w = 0;
  for (int64_t i = j + 1 + threadIdx.x; i < M; i += blockDim.x) w += a[i] * b[i];
  w = (blk_sum(w, red) + b[j]) * tau[j];
  __syncthreads();
  for (int64_t i = j + 1 + threadIdx.x; i < M; i += blockDim.x) b[i] -= w * a[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    b[j] -= w;
    s_recompute = 0;
    if (downdate && vn1[c] != 0.0) {
      double t = fabs(b[j]) / vn1[c];
      t = 1.0 - t * t;
      t = t > 0 ? t : 0;
      const double r = vn1[c] / vn2[c], t2 = t * r * r;
      if (t2 <= sqrt(DBL_EPSILON)) s_recompute = 1;
      else vn1[c] *= sqrt(t);
    }
  }
  __syncthreads();
  if (downdate && s_recompute) {
    double s = 0;
    for (int64_t i = j + 1 + threadIdx.x; i < M; i += blockDim.x) s += b[i] * b[i];
    s = sqrt(blk_sum(s, red));
    if (threadIdx.x == 0) vn1[c] = vn2[c] = s;
  }
}

__global__ void k_eye(double *Q, int64_t M, int k) {
  const int64_t n = M * k;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = i / M, s = i - j * M;
    Q[i] = s == j ? 1.0 : 0.0;
  }
}

// d_X: M x 2, d_C: k x 2 (device).  d_Q: M x k out.  d_work: M*k + 3k doubles.
int basis_build_device(cudaStream_t st, const double *d_X, int64_t M, const double *d_C, int k, double ls, int kernel,
                       double nu, double period, double *d_Q, double *d_work) {
  double *A = d_work, *vn1 = A + M * k, *vn2 = vn1 + k, *tau = vn2 + k;
  k_kernel_matrix<<<148 * 4, 256, 0, st>>>(d_X, M, d_C, k, ls, kernel, nu, period, A);
  k_col_norms<<<k, 512, 0, st>>>(A, M, vn1, vn2);
  const int steps = (int)(k < M ? k : M);
  for (int j = 0; j < steps; ++j) {
    k_qr_pivot_house<<<1, 1024, 0, st>>>(A, M, k, j, vn1, vn2, tau);
    if (j + 1 < k) k_qr_apply<<<k - j - 1, 512, 0, st>>>(A, A, M, j, j + 1, tau, vn1, vn2, 1);
  }
  k_eye<<<148 * 4, 256, 0, st>>>(d_Q, M, k);
  for (int j = steps - 1; j >= 0; --j) k_qr_apply<<<k - j, 512, 0, st>>>(A, d_Q, M, j, j, tau, vn1, vn2, 0);
  CUDA_TRY(cudaGetLastError());
  return BVG_OK;
}
