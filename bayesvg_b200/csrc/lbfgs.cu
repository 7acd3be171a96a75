// lbfgs.cu -- L-BFGS initialisation at the penalised MLE, device-resident.
//
// Replaces mod$optimize(data, init = 0, jacobian = FALSE, iter = mle.iter, algorithm = "lbfgs",
// history_size = 25) of the reference (R/findSpatiallyVariableFeaturesBayes.R:346-353), i.e.
// CmdStan's BFGSMinimizer<LBFGSUpdate> with its default tolerances and Wolfe line search
// (SURVEY.md A.4).  All D-vectors stay on the GPU; the host sees only scalars.  The two-loop
// recursion is evaluated in its "vector-free" form (Chen et al., NIPS 2014): the search direction is
// a linear combination of the 2m+1 basis vectors {s_i, y_i, g}, whose coefficients depend only on
// the (2m+1)^2 Gram matrix; each iteration adds three rows to that matrix with ONE fused kernel and
// forms the direction with one more, instead of 4m dependent dot/axpy launches.
#include <float.h>
#include <stdlib.h>

#include <algorithm>
#include <cmath>
#include <functional>
#include <vector>

#include "bvg_vec.cuh"

namespace {

using EvalFn = std::function<int(double alpha, double *f, double *dg)>;  // 0 = finite

double cubic_interp(double df0, double x1, double f1, double df1, double lo, double hi) {
  const double c3 = (-12 * f1 + 6 * x1 * (df0 + df1)) / (x1 * x1 * x1);
  const double c2 = -(4 * df0 + 2 * df1) / x1 + 6 * f1 / (x1 * x1);
  const double c1 = df0;
  const double ts = std::sqrt(c2 * c2 - 2.0 * c1 * c3);
  const double r1 = -(c2 + ts) / c3, r2 = -(c2 - ts) / c3;
  auto val = [&](double x) { return x * (x * (x * c3 / 3.0 + c2) / 2.0 + c1); };
  double bx = lo, bf = val(lo);
  for (double cand : {hi, r1, r2}) {
    if (cand == hi || (lo < cand && cand < hi)) {
      const double v = val(cand);
      if (v < bf) { bf = v; bx = cand; }
    }
  }
  return bx;
}

int zoom(const EvalFn &ev, double *alpha, double *f1, double *dg1, double f0, double c1dfp, double c2dfp, double alo,
         double aloF, double aloD, double ahi, double ahiF, double ahiD) {
  for (int itn = 1;; ++itn) {
    if (std::fabs(alo - ahi) < 1e-16) return 1;
    double a;
    if (itn % 5 == 0) a = 0.5 * (alo + ahi);
    else {
      const double d1 = aloD + ahiD - 3 * (aloF - ahiF) / (alo - ahi);
      double d2 = std::sqrt(d1 * d1 - aloD * ahiD);
      if (ahi < alo) d2 = -d2;
      a = ahi - (ahi - alo) * (ahiD + d2 - d1) / (ahiD - aloD + 2 * d2);
      const double mn = std::min(alo, ahi), mx = std::max(alo, ahi), w = std::fabs(alo - ahi);
      if (!std::isfinite(a) || a < mn + 0.01 * w || a > mx - 0.01 * w) a = 0.5 * (alo + ahi);
    }
    while (ev(a, f1, dg1)) {
      a = 0.5 * (a + std::min(alo, ahi));
      if (std::fabs(std::min(alo, ahi) - a) < 1e-16) return 1;
    }
    *alpha = a;
    if (*f1 > f0 + a * c1dfp || *f1 >= aloF) { ahi = a; ahiF = *f1; ahiD = *dg1; }
    else {
      if (std::fabs(*dg1) <= -c2dfp) return 0;
      if (*dg1 * (ahi - alo) >= 0) { ahi = alo; ahiF = aloF; ahiD = aloD; }
      alo = a; aloF = *f1; aloD = *dg1;
    }
  }
}

int wolfe(const EvalFn &ev, double *alpha, double *f1, double *dg1, double f0, double dfp) {
  const double c1dfp = 1e-4 * dfp, c2dfp = 0.9 * dfp;
  double a0 = 1e-12, prevF = f0, prevD = dfp;
  int nits = 0, restarts = 0;
  for (;;) {
    if (nits >= 20) return 1;
    if (ev(*alpha, f1, dg1)) {
      if (restarts >= 10) return 1;
      *alpha = 0.5 * (a0 + *alpha);
      ++restarts;
      continue;
    }
    restarts = 0;
    if (*f1 > f0 + *alpha * c1dfp || (*f1 >= prevF && nits > 0))
      return zoom(ev, alpha, f1, dg1, f0, c1dfp, c2dfp, a0, prevF, prevD, *alpha, *f1, *dg1);
    if (std::fabs(*dg1) <= -c2dfp) return 0;
    if (*dg1 >= 0) return zoom(ev, alpha, f1, dg1, f0, c1dfp, c2dfp, *alpha, *f1, *dg1, a0, prevF, prevD);
    a0 = *alpha; prevF = *f1; prevD = *dg1;
    *alpha *= 10.0;
    ++nits;
  }
}

}  // namespace


// d_theta0: device vector or NULL (= 0, the reference's init = 0).  Result left in d_out (device).
// opt (optional): jacobian / propto of the objective, iteration cap, and a callback that sees the initial point (it = 0)
// and every accepted iterate (device x, device gradient of f = -lp, f) -- Pathfinder builds its approximations there.
int lbfgs_run(bvg_fit *f, const double *d_theta0, double *d_out, const LbfgsOpt *opt) {
  if (lbfgs_device_ok(f, opt)) return lbfgs_run_device(f, d_theta0, d_out, opt);  // no host round trip per evaluation (lbfgs_dev.cu)
  const DevModel &m = f->m;
  const int64_t D = m.L.D;
  const int jac = opt ? opt->jacobian : 0, propto = opt ? opt->propto : 0;
  const int mh = std::min<int>(std::max(opt && opt->history > 0 ? opt->history : f->cfg.lbfgs_history, 1), LBFGS_MAX_HISTORY), nb = 2 * mh + 1;
  cudaStream_t st = f->stream;
  double *B = nullptr, *x0, *p, *w, *part, *dots, *scal;
  const size_t nd = (size_t)nb * D + 3 * (size_t)D + (size_t)nb * VEC_NCHUNK * 3 + (size_t)nb * 3 + 8;
  CUDA_TRY(cudaMalloc(&B, nd * sizeof(double)));
  CUDA_TRY(cudaMemsetAsync(B, 0, nd * sizeof(double), st));
  x0 = B + (size_t)nb * D; p = x0 + D; w = p + D; part = w + D; dots = part + (size_t)nb * VEC_NCHUNK * 3; scal = dots + nb * 3;
  double *g0 = B + (size_t)(2 * mh) * D;  // the gradient row of the basis
  int rc = BVG_OK;
  double *h_scal = nullptr;  // pinned: [0] grad.p, [1] sum grad, [2] lp  -- one small D2H per evaluation
  auto fail = [&](int code) { shadow_leave(f); cudaFree(B); cudaFreeHost(h_scal); return code; };
  if (cudaMallocHost(&h_scal, 4 * sizeof(double)) != cudaSuccess) { bvg_set_error("cudaMallocHost failed"); return fail(BVG_ERR_CUDA); }
  if ((rc = vec_make_weights(st, w, m, f->cfg.rank)) != BVG_OK) return fail(rc);
  if (d_theta0) CUDA_TRY(cudaMemcpyAsync(x0, d_theta0, D * sizeof(double), cudaMemcpyDeviceToDevice, st));

  int fevals = 0;
  // f(x0 + alpha p) and its directional derivative; leaves x in m.zeta and grad(lp) in m.grad
  EvalFn ev = [&](double alpha, double *fv, double *dg) -> int {
    if (vec_axpy_to(st, m.zeta, x0, alpha, p, D) != BVG_OK) return 2;
    if (eval_log_prob_device(f, MODE_GRAD_OUT, jac, propto, true) != BVG_OK) return 2;
    if (vec_dot_sum(st, m.grad, p, w, D, scal) != BVG_OK) return 2;
    if (f->comm_mode && comm_allreduce(f, scal, 2) != BVG_OK) return 2;
    if (cudaMemcpyAsync(scal + 2, &m.ctl->lp, sizeof(double), cudaMemcpyDeviceToDevice, st) != cudaSuccess) return 2;
    if (cudaMemcpyAsync(h_scal, scal, 3 * sizeof(double), cudaMemcpyDeviceToHost, st) != cudaSuccess) return 2;
    if (cudaStreamSynchronize(st) != cudaSuccess) return 2;
    ++fevals;
    *fv = -h_scal[2];
    *dg = -h_scal[0];
    return (std::isfinite(h_scal[2]) && std::isfinite(h_scal[0]) && std::isfinite(h_scal[1])) ? 0 : 1;
  };

  std::vector<double> Gm((size_t)nb * nb, 0.0), hd((size_t)nb * 3), delta(nb), al(mh);
  double fk, dg;
  if (ev(0.0, &fk, &dg)) { bvg_set_error("L-BFGS: log density is not finite at the initial point"); return fail(BVG_ERR_NUMERIC); }
  // g0 = -grad, p = -g0 = grad
  if ((rc = vec_accept(st, nullptr, nullptr, x0, g0, m.zeta, m.grad, D)) != BVG_OK) return fail(rc);
  if (opt && opt->on_iterate && (rc = opt->on_iterate(0, x0, g0, fk)) != BVG_OK) return fail(rc);
  CUDA_TRY(cudaMemcpyAsync(p, m.grad, D * sizeof(double), cudaMemcpyDeviceToDevice, st));
  if ((rc = vec_dot_sum(st, g0, g0, w, D, scal)) != BVG_OK) return fail(rc);
  if (f->comm_mode && (rc = comm_allreduce(f, scal, 2)) != BVG_OK) return fail(rc);
  CUDA_TRY(cudaMemcpyAsync(h_scal, scal, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  double gg = h_scal[0];   // g.g
  double dfp = -gg;        // g.p for p = -g
  int nh = 0, head = 0, it = 0, term = 0, switch_iter = 0;
  int force_at = 0;
  if (const char *e = getenv("BVG_MLE_FP64_AT")) force_at = atoi(e);
  double fk1 = 0, alpha = 1e-3, alphak1 = 0, dfp_prev = 0, gamma = 1;
  const double eps = DBL_EPSILON, tol_obj = 1e-12, tol_rel_obj = 1e4, tol_grad = 1e-8, tol_rel_grad = 1e7, tol_param = 1e-8;
  const int max_iter = opt && opt->max_iter > 0 ? opt->max_iter : f->cfg.mle_iter;
  while (!term && max_iter > 0) {
    ++it;
    int resetB = (it == 1) ? 1 : 0;
    double f1 = 0, dg1 = 0;
    for (;;) {
      if (resetB) {  // p = -g
        LinCoef c{};
        c.d[2 * mh] = -1.0;
        if ((rc = vec_lincomb(st, p, B, D, nb, c)) != BVG_OK) return fail(rc);
        dfp = -gg;
      }
      if (it > 1 && resetB != 2) alpha = std::min(1.0, 1.01 * cubic_interp(dfp_prev, alphak1, fk - fk1, dfp, 1e-12, 1.0));
      else alpha = 1e-3;
      int ls = wolfe(ev, &alpha, &f1, &dg1, fk, dfp);
      if (force_at > 0 && it == force_at && !shadow_active(f)) { ls = 1; resetB = resetB ? resetB : 2; }  // tests: hand over at a chosen iteration
      if (ls) {
        if (resetB) {
          // even steepest descent finds no acceptable step.  Fast precision: the search is comparing the evaluator's noise
          // (fp64_shadow.cu) -- continue from this very point with fp64 evaluations, fresh curvature history
          if (f->cfg.precision == BVG_PREC_FAST && !opt && !f->no_mle_fp64) {
            if ((rc = shadow_enter(f)) != BVG_OK) return fail(rc);
            switch_iter = it;
            if (ev(0.0, &fk, &dg1)) { bvg_set_error("L-BFGS: log density is not finite where the fp64 evaluator takes over"); return fail(BVG_ERR_NUMERIC); }
            if ((rc = vec_accept(st, nullptr, nullptr, x0, g0, m.zeta, m.grad, D)) != BVG_OK) return fail(rc);
            if ((rc = vec_dot_sum(st, g0, g0, w, D, scal)) != BVG_OK) return fail(rc);
            if (f->comm_mode && (rc = comm_allreduce(f, scal, 2)) != BVG_OK) return fail(rc);
            CUDA_TRY(cudaMemcpyAsync(h_scal, scal, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            gg = h_scal[0];
            resetB = 2;   // steepest descent from the re-evaluated point, initial step as after any failed search
            continue;
          }
          term = -1;
          break;
        }
        resetB = 2;
        continue;
      }
      break;
    }
    if (term) break;
    // accept: s, y rows; x0 <- x1; g0 <- g1
    if ((rc = vec_accept(st, B + (size_t)head * D, B + (size_t)(mh + head) * D, x0, g0, m.zeta, m.grad, D)) != BVG_OK) return fail(rc);
    fk1 = fk; fk = f1;
    dfp_prev = dfp;
    if (opt && opt->on_iterate && (rc = opt->on_iterate(it, x0, g0, fk)) != BVG_OK) return fail(rc);
    if (resetB) {  // history cleared before the push (LBFGSUpdate::update(reset = true))
      for (auto &v : Gm) v = 0.0;
      if (head != 0) {  // move the fresh pair to slot 0
        CUDA_TRY(cudaMemcpyAsync(B, B + (size_t)head * D, D * sizeof(double), cudaMemcpyDeviceToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(B + (size_t)mh * D, B + (size_t)(mh + head) * D, D * sizeof(double), cudaMemcpyDeviceToDevice, st));
        head = 0;
      }
      nh = 0;
    }
    if ((rc = vec_multidot3(st, B, D, nb, head, mh + head, 2 * mh, w, part, dots)) != BVG_OK) return fail(rc);
    if (f->comm_mode && (rc = comm_allreduce(f, dots, nb * 3)) != BVG_OK) return fail(rc);
    CUDA_TRY(cudaMemcpyAsync(hd.data(), dots, nb * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    const int rs = head, ry = mh + head, rg = 2 * mh;
    for (int r = 0; r < nb; ++r) {
      Gm[(size_t)rs * nb + r] = Gm[(size_t)r * nb + rs] = hd[r * 3 + 0];
      Gm[(size_t)ry * nb + r] = Gm[(size_t)r * nb + ry] = hd[r * 3 + 1];
      Gm[(size_t)rg * nb + r] = Gm[(size_t)r * nb + rg] = hd[r * 3 + 2];
    }
    const double skyk = Gm[(size_t)rs * nb + ry], ykyk = Gm[(size_t)ry * nb + ry];
    const double snorm = std::sqrt(Gm[(size_t)rs * nb + rs]);
    gg = Gm[(size_t)rg * nb + rg];
    const double gnorm = std::sqrt(gg);
    if (resetB) {
      const double B0 = ykyk / skyk;
      dfp_prev /= B0;          // _pk_1 /= B0fact
      alphak1 = alpha * B0;
    } else alphak1 = alpha;
    gamma = skyk / ykyk;
    head = (head + 1) % mh;
    if (nh < mh) ++nh;
    if (std::fabs(fk1 - fk) < tol_obj) term = 1;
    else if (std::fabs(fk1 - fk) < tol_rel_obj * eps * std::max(std::fabs(fk1), std::max(std::fabs(fk), 1.0))) term = 2;
    else if (gnorm < tol_grad) term = 3;
    else if (snorm < tol_param) term = 4;
    else if (it >= max_iter) term = 5;
    // vector-free two-loop recursion on the Gram matrix
    std::fill(delta.begin(), delta.end(), 0.0);
    delta[rg] = -1.0;
    auto col_dot = [&](int col) { double s = 0; for (int j = 0; j < nb; ++j) if (delta[j] != 0.0) s += delta[j] * Gm[(size_t)j * nb + col]; return s; };
    for (int q = 0; q < nh; ++q) {
      const int h = (head - 1 - q + 2 * mh) % mh;
      al[h] = col_dot(h) / Gm[(size_t)h * nb + mh + h];
      delta[mh + h] -= al[h];
    }
    for (auto &d : delta) d *= gamma;
    for (int q = nh - 1; q >= 0; --q) {
      const int h = (head - 1 - q + 2 * mh) % mh;
      const double b = col_dot(mh + h) / Gm[(size_t)h * nb + mh + h];
      delta[h] += al[h] - b;
    }
    LinCoef c{};
    for (int j = 0; j < nb; ++j) c.d[j] = delta[j];
    if ((rc = vec_lincomb(st, p, B, D, nb, c)) != BVG_OK) return fail(rc);
    dfp = col_dot(rg);  // g . p
    if (!term) {
      const double rgd = -dfp / std::max(std::fabs(fk), 1.0);
      if (rgd < tol_rel_grad * eps) term = 6;
    }
  }
  CUDA_TRY(cudaMemcpyAsync(d_out, x0, D * sizeof(double), cudaMemcpyDeviceToDevice, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  f->st.mle_iters = it;
  f->st.mle_fevals = fevals;
  f->st.mle_term = term;
  f->st.mle_fp64_from = switch_iter;
  shadow_leave(f);
  f->st.mle_lp = -fk;
  cudaFree(B);
  cudaFreeHost(h_scal);
  return BVG_OK;
}
