// bvg_vec.cuh -- declarations of the small vector kernels in k_vec.cu
#pragma once
#include "bvg_internal.cuh"

#define VEC_NCHUNK 8
#define LBFGS_MAX_HISTORY 32
struct LinCoef { double d[2 * LBFGS_MAX_HISTORY + 1]; };

#include <functional>
struct LbfgsOpt {
  int jacobian = 0, propto = 0, max_iter = 0, history = 0;
  std::function<int(int it, const double *d_x, const double *d_g, double fval)> on_iterate;
};
int lbfgs_run(bvg_fit *f, const double *d_theta0, double *d_out, const LbfgsOpt *opt = nullptr);  // lbfgs.cu (dispatch + host-driven form)
bool lbfgs_device_ok(const bvg_fit *f, const LbfgsOpt *opt);                                       // lbfgs_dev.cu: control flow on the device
int lbfgs_run_device(bvg_fit *f, const double *d_theta0, double *d_out, const LbfgsOpt *opt);

template <typename SRC, typename DST>
int stage_y(const SRC *src, DST *dst, int64_t M, int64_t ldY, int64_t ngenes, cudaStream_t st);
template <typename SRC>
int stage_y_tiled(const SRC *src, float *dst, int64_t M, int64_t ldY, int64_t g0, int64_t ngenes, cudaStream_t st);
template <typename T>
int stage_phi(const double *phi, T *out, int64_t M, int64_t Mpad, int k, int KP, cudaStream_t st);
int vec_multidot3(cudaStream_t st, const double *B, int64_t D, int nrows, int ra, int rb, int rc, const double *w,
                  double *part, double *out);
int vec_dot_sum(cudaStream_t st, const double *a, const double *b, const double *w, int64_t D, double *out);
int vec_axpy_to(cudaStream_t st, double *out, const double *x, double a, const double *p, int64_t D);
int vec_accept(cudaStream_t st, double *srow, double *yrow, double *x0, double *g0, const double *x1, const double *grad1, int64_t D);
int vec_lincomb(cudaStream_t st, double *p, const double *B, int64_t D, int nrows, const LinCoef &coef);
int vec_fill(cudaStream_t st, double *x, double v, int64_t n);
int vec_make_weights(cudaStream_t st, double *w, const DevModel &m, int rank);
