// lbfgs_dev.cu -- the L-BFGS initialisation with its control flow ON THE DEVICE.
//
// Same algorithm as lbfgs.cu (CmdStan's BFGSMinimizer<LBFGSUpdate> + Wolfe line search behind mod$optimize,
// R/findSpatiallyVariableFeaturesBayes.R:346-353, in the vector-free Gram form), but the host no longer sees a scalar per
// evaluation.  The line search (bracketing, zoom, cubic interpolation), the termination tests and the two-loop recursion on
// the Gram matrix are state machines that live in one device struct (LbDev); the host only enqueues "ticks" -- one function
// evaluation each -- and reads a 64-byte status every LB_BATCH ticks, one batch behind the one it is enqueueing, so the
// stream never drains:
//
//   k_prep, likelihood pass, k_gene + finish   the evaluation chain (log density + gradient); k_prep forms the trial point
//                zeta = x0 + a p itself, with a read from the device (LsPoint)
//   k_lb_gdot    g.p over all SMs; its last CTA adds the partials (+ the peer exchange of a gene-sharded fit) and runs ONE
//                step of the Wolfe search: next trial step, or accept
//   k_lb_dots    (only after an accepted step) s = x1 - x0, y = g1 - g0 formed on the fly and stored as history rows, three
//                rows of the Gram matrix; its last CTA adds the chunks, exchanges, tests for termination, runs the two-loop
//                recursion and opens the next iteration's search
//   k_lb_dir     (only after an accepted step / when the direction is new) x0 <- x1, g0 <- g1;
//                p = sum of the active basis rows times their coefficients
//
// A tick whose kernels have nothing to do (search still running: no accept; run finished) costs their launch only.
// Host-driven lbfgs.cu remains for the callers that need the host at every iterate (Pathfinder's on_iterate) and for the
// NCCL transport (its all-reduce is a host call between two kernels).
#include <float.h>
#include <stdlib.h>

#include <algorithm>
#include <chrono>
#include <cmath>

#include "bvg_vec.cuh"

#define LB_NB (2 * LBFGS_MAX_HISTORY + 1)
#define LB_BATCH 16
#define LB_NCHUNK 40   // element chunks of the Gram-row kernel
#define LB_RG 8        // basis rows per CTA of the Gram-row kernel
#define LB_GDOT_MAX 592 // CTAs of the g.p kernel (4 per SM)
enum { PH_LS = 0, PH_DONE = 1, PH_NEED_SHADOW = 2 };
enum { LS_WOLFE = 0, LS_ZOOM = 1 };

struct LbDev {
  // ---- status block: the first 64 bytes are what the host reads ----
  int32_t phase, term, it, fevals, switch_iter, accept, fresh_dir, pad0;
  double fk;
  double pad1[3];
  // ---- configuration ----
  int32_t mh, nb, max_iter, force_at, allow_shadow, shadow_on;
  // ---- outer iteration ----
  int32_t resetB, nh, head, wslot;   // wslot: history slot the accepted pair is written to
  double fk1, gg, dfp, dfp_prev, alpha, alphak1, gamma;
  // ---- line search ----
  int32_t ls_state, nits, restarts, itn;
  double a_try, f0, c1dfp, c2dfp, a0, prevF, prevD;
  double alo, aloF, aloD, ahi, ahiF, ahiD;
  double f1, dg1;
  uint32_t ticket, ticket2, ticket3, pad2;
  int32_t n_accept, n_applied;   // accepted steps / accepted steps whose x0 <- x1, g0 <- g1 has been applied (k_lb_dir)
  int32_t nact, pad3;          // the direction as a compact list: p = sum_q act_coef[q] * B[act_idx[q]]
  int32_t act_idx[LB_NB + 1];
  double act_coef[LB_NB];
  double Gm[LB_NB * LB_NB];
  long long dbg[8];            // clock64 stamps of the last CTA of k_lb_dots (BVG_LBFGS_TRACE=1 prints them)
};

namespace {

// ---------------------------------------------------------------------------------------------------------------------
// scalar control (one thread).  Statement for statement the host code of lbfgs.cu, unrolled into "what to do after the
// evaluation that just came back".
__device__ double lb_cubic_interp(double df0, double x1, double f1, double df1, double lo, double hi) {
  const double c3 = (-12 * f1 + 6 * x1 * (df0 + df1)) / (x1 * x1 * x1);
  const double c2 = -(4 * df0 + 2 * df1) / x1 + 6 * f1 / (x1 * x1);
  const double c1 = df0;
  const double ts = sqrt(c2 * c2 - 2.0 * c1 * c3);
  const double r1 = -(c2 + ts) / c3, r2 = -(c2 - ts) / c3;
  double bx = lo, bf = lo * (lo * (lo * c3 / 3.0 + c2) / 2.0 + c1);
  const double cand[3] = {hi, r1, r2};
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const double c = cand[q];
    if (q == 0 || c == hi || (lo < c && c < hi)) {
      const double v = c * (c * (c * c3 / 3.0 + c2) / 2.0 + c1);
      if (v < bf) { bf = v; bx = c; }
    }
  }
  return bx;
}

enum { R_EVAL = 0, R_SUCCESS = 1, R_FAIL = 2 };

// next trial step of the zoom phase (the head of zoom()'s loop)
__device__ int lb_zoom_next(LbDev &s) {
  if (fabs(s.alo - s.ahi) < 1e-16) return R_FAIL;
  double a;
  if (s.itn % 5 == 0) a = 0.5 * (s.alo + s.ahi);
  else {
    const double d1 = s.aloD + s.ahiD - 3 * (s.aloF - s.ahiF) / (s.alo - s.ahi);
    double d2 = sqrt(d1 * d1 - s.aloD * s.ahiD);
    if (s.ahi < s.alo) d2 = -d2;
    a = s.ahi - (s.ahi - s.alo) * (s.ahiD + d2 - d1) / (s.ahiD - s.aloD + 2 * d2);
    const double mn = fmin(s.alo, s.ahi), mx = fmax(s.alo, s.ahi), w = fabs(s.alo - s.ahi);
    if (!isfinite(a) || a < mn + 0.01 * w || a > mx - 0.01 * w) a = 0.5 * (s.alo + s.ahi);
  }
  s.a_try = a;
  return R_EVAL;
}
__device__ int lb_zoom_init(LbDev &s, double alo, double aloF, double aloD, double ahi, double ahiF, double ahiD) {
  s.ls_state = LS_ZOOM; s.itn = 1;
  s.alo = alo; s.aloF = aloF; s.aloD = aloD; s.ahi = ahi; s.ahiF = ahiF; s.ahiD = ahiD;
  return lb_zoom_next(s);
}
// the evaluation at a_try came back as (bad, f1, dg1)
__device__ int lb_ls_step(LbDev &s, bool bad, double f1, double dg1) {
  if (s.ls_state == LS_WOLFE) {
    if (bad) {
      if (s.restarts >= 10) return R_FAIL;
      s.alpha = 0.5 * (s.a0 + s.alpha);
      s.a_try = s.alpha;
      ++s.restarts;
      return R_EVAL;
    }
    s.restarts = 0;
    if (f1 > s.f0 + s.alpha * s.c1dfp || (f1 >= s.prevF && s.nits > 0)) return lb_zoom_init(s, s.a0, s.prevF, s.prevD, s.alpha, f1, dg1);
    if (fabs(dg1) <= -s.c2dfp) return R_SUCCESS;
    if (dg1 >= 0) return lb_zoom_init(s, s.alpha, f1, dg1, s.a0, s.prevF, s.prevD);
    s.a0 = s.alpha; s.prevF = f1; s.prevD = dg1;
    s.alpha *= 10.0;
    s.a_try = s.alpha;
    ++s.nits;
    if (s.nits >= 20) return R_FAIL;
    return R_EVAL;
  }
  // zoom
  double a = s.a_try;
  if (bad) {
    a = 0.5 * (a + fmin(s.alo, s.ahi));
    if (fabs(fmin(s.alo, s.ahi) - a) < 1e-16) return R_FAIL;
    s.a_try = a;
    return R_EVAL;
  }
  s.alpha = a;
  if (f1 > s.f0 + a * s.c1dfp || f1 >= s.aloF) { s.ahi = a; s.ahiF = f1; s.ahiD = dg1; }
  else {
    if (fabs(dg1) <= -s.c2dfp) return R_SUCCESS;
    if (dg1 * (s.ahi - s.alo) >= 0) { s.ahi = s.alo; s.ahiF = s.aloF; s.ahiD = s.aloD; }
    s.alo = a; s.aloF = f1; s.aloD = dg1;
  }
  ++s.itn;
  return lb_zoom_next(s);
}
// open the search of the current iteration (or reopen it along steepest descent after a failed one: resetB = 2)
__device__ void lb_ls_begin(LbDev &s) {
  if (s.resetB) {  // p = -g
    s.nact = 1; s.act_idx[0] = 2 * s.mh; s.act_coef[0] = -1.0;
    s.fresh_dir = 1;
    s.dfp = -s.gg;
  }
  if (s.it > 1 && s.resetB != 2) s.alpha = fmin(1.0, 1.01 * lb_cubic_interp(s.dfp_prev, s.alphak1, s.fk - s.fk1, s.dfp, 1e-12, 1.0));
  else s.alpha = 1e-3;
  s.a_try = s.alpha;
  s.c1dfp = 1e-4 * s.dfp; s.c2dfp = 0.9 * s.dfp; s.f0 = s.fk;
  s.a0 = 1e-12; s.prevF = s.fk; s.prevD = s.dfp;
  s.nits = 0; s.restarts = 0; s.ls_state = LS_WOLFE;
  s.accept = 0;
}
// the search of this iteration found no acceptable step
__device__ void lb_ls_failed(LbDev &s) {
  if (s.resetB) {
    // even steepest descent finds nothing: a fast-precision fit hands over to the fp64 evaluator (host: fp64_shadow.cu)
    if (s.allow_shadow && !s.shadow_on) { s.phase = PH_NEED_SHADOW; return; }
    s.term = -1;
    s.phase = PH_DONE;
    return;
  }
  s.resetB = 2;
  lb_ls_begin(s);
}

// ---------------------------------------------------------------------------------------------------------------------
__global__ void k_lb_begin(LbDev *s) { lb_ls_begin(*s); s->phase = PH_LS; }

// After an accepted step: x0 <- x1, g0 <- g1 (the gradient row of the basis; k_lb_dots has read the old ones).  With a new
// direction: p = sum of the active basis rows {s_i, y_i, g} times their coefficients (32 rows in flight per thread).
__global__ void __launch_bounds__(256) k_lb_dir(LbDev *s, double *__restrict__ p, double *B, double *__restrict__ x0,
                                                const double *__restrict__ x1, const double *__restrict__ grad1, int64_t D) {
  __shared__ int sidx[LB_NB + 31];
  __shared__ double scoef[LB_NB + 31];
  __shared__ int s_last;
  griddep_wait();     // programmatic dependent launch: this grid's launch latency hides under the previous kernel
  griddep_launch();
  const bool upd = s->n_accept != s->n_applied, fresh = s->phase == PH_LS && s->fresh_dir;
  if (!upd && !fresh) return;
  const int nact = fresh ? s->nact : 0, rg = 2 * s->mh;
  if (threadIdx.x < LB_NB + 31) {
    const bool ok = (int)threadIdx.x < nact;
    sidx[threadIdx.x] = ok ? s->act_idx[threadIdx.x] : rg;
    scoef[threadIdx.x] = ok ? s->act_coef[threadIdx.x] : 0.0;
  }
  __syncthreads();
  double *g0 = B + (int64_t)rg * D;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < D; i += (int64_t)gridDim.x * blockDim.x) {
    double gn = 0;
    if (upd) { gn = -grad1[i]; x0[i] = x1[i]; g0[i] = gn; }
    if (!fresh) continue;
    double pi = 0;
    for (int r0 = 0; r0 < nact; r0 += 32) {
      double bv[32];
#pragma unroll
      for (int q = 0; q < 32; ++q) bv[q] = __ldcg(B + (int64_t)sidx[r0 + q] * D + i);
#pragma unroll
      for (int q = 0; q < 32; ++q) pi += scoef[r0 + q] * ((upd && sidx[r0 + q] == rg) ? gn : bv[q]);   // the g row was just replaced by this thread
    }
    p[i] = pi;
  }
  if (!upd) return;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned ticket;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(ticket) : "l"(&s->ticket3) : "memory");
    s_last = ticket == gridDim.x - 1;
    if (s_last) { s->ticket3 = 0; s->n_applied = s->n_accept; }
  }
}

// after the evaluation: g.p and sum g (finiteness) over all SMs; the last CTA adds the partials in a fixed order, runs the
// exchange of a gene-sharded fit and takes one step of the line search
__global__ void __launch_bounds__(256) k_lb_gdot(LbDev *s, DevModel m, const double *__restrict__ p, const double *__restrict__ w,
                                                 double *__restrict__ part, double *scal) {
  extern __shared__ double xscratch[];   // (world - 1) * 2 doubles of polling scratch
  __shared__ double red[2][8];
  __shared__ int s_last;
  griddep_wait();
  griddep_launch();
  if (s->phase != PH_LS) return;
  const int64_t D = m.L.D;
  const double *__restrict__ grad = m.grad;
  double s0 = 0, s1 = 0;
#pragma unroll 4
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < D; i += (int64_t)gridDim.x * blockDim.x) {
    const double x = grad[i] * w[i];
    s0 += x * p[i];
    s1 += x;
  }
  s0 = warp_sum(s0); s1 = warp_sum(s1);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s0; red[1][threadIdx.x >> 5] = s1; }
  __syncthreads();
  if (threadIdx.x < 2) {
    double t = 0;
    for (int q = 0; q < 8; ++q) t += red[threadIdx.x][q];
    part[2 * blockIdx.x + threadIdx.x] = t;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned ticket;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(ticket) : "l"(&s->ticket2) : "memory");
    s_last = ticket == gridDim.x - 1;
    if (s_last) s->ticket2 = 0;
  }
  __syncthreads();
  if (!s_last) return;
  s0 = 0; s1 = 0;
  for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) { s0 += __ldcg(&part[2 * b]); s1 += __ldcg(&part[2 * b + 1]); }
  s0 = warp_sum(s0); s1 = warp_sum(s1);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s0; red[1][threadIdx.x >> 5] = s1; }
  __syncthreads();
  if (threadIdx.x < 2) {
    double t = 0;
    for (int q = 0; q < 8; ++q) t += red[threadIdx.x][q];
    scal[threadIdx.x] = t;
  }
  if (m.peer.world > 1) peer_allreduce_cta(m.peer, scal, 2, xscratch);
  else __syncthreads();
  if (threadIdx.x != 0) return;
  const double gp = scal[0], gs = scal[1], lp = m.ctl->lp;
  LbDev &S = *s;
  S.fevals += 1;
  S.fresh_dir = 0;
  const bool bad = !(isfinite(lp) && isfinite(gp) && isfinite(gs));
  const double f1 = -lp, dg1 = -gp;
  S.f1 = f1; S.dg1 = dg1;
  int r = lb_ls_step(S, bad, f1, dg1);
  if (r == R_SUCCESS && S.force_at > 0 && S.it == S.force_at && !S.shadow_on) {  // tests: hand over at a chosen iteration
    r = R_FAIL;
    if (!S.resetB) S.resetB = 2;
  }
  if (r == R_SUCCESS) {
    S.accept = 1;
    S.n_accept += 1;
    S.wslot = S.resetB ? 0 : S.head;   // a reset clears the history before the push: the fresh pair goes to slot 0
  } else if (r == R_FAIL) lb_ls_failed(S);
}

// Gram rows of the new (s, y, g); the last CTA finishes the iteration
// The accepted step is formed here on the fly -- s = x1 - x0, y = g1 - g0 from (zeta, grad) and the still unchanged (x0, g0 row) --
// and the row-group-0 CTAs store the s / y rows; x0 <- x1 and g0 <- g1 wait for k_lb_dir (other CTAs of this grid still read them).
__global__ void __launch_bounds__(256, 2) k_lb_dots(LbDev *s, DevModel m, double *B, const double *__restrict__ x0, const double *__restrict__ w,
                                                 double *__restrict__ part, double *dots) {
  extern __shared__ double sh[];   // last CTA: Gram matrix [nb * nb] | polling scratch
  __shared__ double red[LB_RG * 3][8];
  __shared__ int s_last;
  __shared__ double s_rho[LBFGS_MAX_HISTORY], s_dots[LB_NB * 3];
  griddep_wait();
  griddep_launch();
  if (s->phase != PH_LS || !s->accept) return;
  const long long t_start = clock64();
  const int64_t D = m.L.D;
  const int nb = s->nb, mh = s->mh, nchunk = LB_NCHUNK;   // gridDim.y == LB_NCHUNK
  const int rs = s->wslot, ry = mh + rs, rg = 2 * mh;
  {
    // a CTA owns LB_RG basis rows and one chunk of elements: the new (s, y, g) and the weights are read once per LB_RG rows
    const int r0 = blockIdx.x * LB_RG, ch = blockIdx.y;
    const int64_t per = (D + nchunk - 1) / nchunk, i0 = ch * per, i1 = (i0 + per < D) ? i0 + per : D;
    const double *__restrict__ x1 = m.zeta, *__restrict__ grad1 = m.grad;
    const double *g0row = B + (int64_t)rg * D;
    double *srow = B + (int64_t)rs * D, *yrow = B + (int64_t)ry * D;
    const double *brow[LB_RG];
    int kind[LB_RG];   // 0: stored row, 1 / 2 / 3: the new s / y / g
#pragma unroll
    for (int q = 0; q < LB_RG; ++q) {
      const int r = r0 + q < nb ? r0 + q : nb - 1;
      kind[q] = r == rs ? 1 : (r == ry ? 2 : (r == rg ? 3 : 0));
      brow[q] = kind[q] ? g0row : B + (int64_t)r * D;   // the rows being replaced are never read
    }
    double acc[LB_RG][3];
#pragma unroll
    for (int q = 0; q < LB_RG; ++q) acc[q][0] = acc[q][1] = acc[q][2] = 0.0;
#pragma unroll 2
    for (int64_t i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
      const double xn = x1[i], gn = -grad1[i], si = xn - x0[i], yi = gn - __ldcg(g0row + i);
      const double wi = w[i], xs = si * wi, xy = yi * wi, xg = gn * wi;
      double bv[LB_RG];
#pragma unroll
      for (int q = 0; q < LB_RG; ++q) bv[q] = __ldcg(brow[q] + i);
#pragma unroll
      for (int q = 0; q < LB_RG; ++q) {
        const double b = kind[q] == 0 ? bv[q] : (kind[q] == 1 ? si : (kind[q] == 2 ? yi : gn));
        acc[q][0] += b * xs; acc[q][1] += b * xy; acc[q][2] += b * xg;
      }
      if (blockIdx.x == 0) { srow[i] = si; yrow[i] = yi; }
    }
#pragma unroll
    for (int q = 0; q < LB_RG; ++q)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const double t = warp_sum(acc[q][c]);
        if ((threadIdx.x & 31) == 0) red[q * 3 + c][threadIdx.x >> 5] = t;
      }
    __syncthreads();
    if (threadIdx.x < LB_RG * 3) {
      const int q = threadIdx.x / 3, c = threadIdx.x - q * 3;
      double t = 0;
      for (int wq = 0; wq < 8; ++wq) t += red[threadIdx.x][wq];
      if (r0 + q < nb) part[((int64_t)(r0 + q) * nchunk + ch) * 3 + c] = t;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned ticket;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(ticket) : "l"(&s->ticket) : "memory");
    s_last = ticket == gridDim.x * gridDim.y - 1;
    if (s_last) s->ticket = 0;
  }
  __syncthreads();
  if (!s_last) return;
  if (threadIdx.x == 0) { s->dbg[0] = t_start; s->dbg[1] = clock64(); }
  // ---- last CTA: chunk sums in a fixed order, exchange, Gram update, termination tests, two-loop recursion ----
  LbDev &S = *s;
  double *G = sh;
  const bool reset = S.resetB != 0;
  const int n2 = nb * nb;
  double gv[17];   // the old Gram matrix: all of a thread's loads in flight, under the chunk sums
#pragma unroll
  for (int q = 0; q < 17; ++q) { const int i = threadIdx.x + q * 256; gv[q] = (!reset && i < n2) ? __ldcg(&S.Gm[i]) : 0.0; }
  if (threadIdx.x < nb * 3) {
    const int r = threadIdx.x / 3, q = threadIdx.x - r * 3;
    const double *pp = part + ((int64_t)r * LB_NCHUNK) * 3 + q;
    double pv[LB_NCHUNK];
#pragma unroll
    for (int ch = 0; ch < LB_NCHUNK; ++ch) pv[ch] = __ldcg(pp + ch * 3);
    double t = 0;
#pragma unroll
    for (int ch = 0; ch < LB_NCHUNK; ++ch) t += pv[ch];
    s_dots[threadIdx.x] = t;
    if (m.peer.world > 1) dots[threadIdx.x] = t;
  }
#pragma unroll
  for (int q = 0; q < 17; ++q) { const int i = threadIdx.x + q * 256; if (i < n2) G[i] = gv[q]; }
  if (m.peer.world > 1) {
    peer_allreduce_cta(m.peer, dots, nb * 3, sh + nb * nb);
    if (threadIdx.x < nb * 3) s_dots[threadIdx.x] = dots[threadIdx.x];
  }
  __syncthreads();
  if (threadIdx.x == 0) S.dbg[2] = clock64();
  // rows / columns of the new s, then y, then g (later ones overwrite the shared corners, as the host loop does)
  for (int c = 0; c < 3; ++c) {
    const int rr = c == 0 ? rs : (c == 1 ? ry : rg);
    if (threadIdx.x < nb) { const double d = s_dots[threadIdx.x * 3 + c]; G[rr * nb + threadIdx.x] = d; G[threadIdx.x * nb + rr] = d; }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < n2; i += blockDim.x) S.Gm[i] = G[i];
  if (threadIdx.x >= 32) return;
  if (threadIdx.x == 0) S.dbg[3] = clock64();
  // ---- one warp: scalars (all lanes compute the same values), then the recursion with v = Gm * delta kept up to date ----
  const int lane = threadIdx.x;
  const double eps = DBL_EPSILON, tol_obj = 1e-12, tol_rel_obj = 1e4, tol_grad = 1e-8, tol_rel_grad = 1e7, tol_param = 1e-8;
  const double fk1 = S.fk, fk = S.f1;   // fk1 <- fk; fk <- f1
  const double dfp_old = S.dfp;
  const double skyk = G[rs * nb + ry], ykyk = G[ry * nb + ry];
  const double snorm = sqrt(G[rs * nb + rs]);
  const double gg = G[rg * nb + rg];
  const double gnorm = sqrt(gg);
  double dfp_prev = dfp_old, alphak1;
  if (reset) {
    const double B0 = ykyk / skyk;
    dfp_prev /= B0;          // _pk_1 /= B0fact
    alphak1 = S.alpha * B0;
  } else alphak1 = S.alpha;
  const double gamma = skyk / ykyk;
  int head = reset ? 0 : S.head, nh = reset ? 0 : S.nh;
  head = (head + 1) % mh;
  if (nh < mh) ++nh;
  const int it = S.it;
  int term = 0;
  if (fabs(fk1 - fk) < tol_obj) term = 1;
  else if (fabs(fk1 - fk) < tol_rel_obj * eps * fmax(fabs(fk1), fmax(fabs(fk), 1.0))) term = 2;
  else if (gnorm < tol_grad) term = 3;
  else if (snorm < tol_param) term = 4;
  else if (it >= S.max_iter) term = 5;
  // lane owns basis indices lane, lane + 32, lane + 64
  double dl[3] = {0, 0, 0}, v[3];
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const int j = lane + 32 * q;
    v[q] = j < nb ? -G[j * nb + rg] : 0.0;   // delta = -e_g
    if (j == rg) dl[q] = -1.0;
  }
  auto vget = [&](int idx) {
    const int q = idx >> 5;
    const double mine = q == 0 ? v[0] : (q == 1 ? v[1] : v[2]);
    return __shfl_sync(0xffffffffu, mine, idx & 31);
  };
  auto axpy_col = [&](double c, int col, int dpos) {  // delta[dpos] += c; v += c * Gm[:, col]
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const int j = lane + 32 * q;
      if (j < nb) v[q] += c * G[col * nb + j];
      if (j == dpos) dl[q] += c;
    }
  };
  if (lane == 0) S.dbg[4] = clock64();
  if (lane < mh) s_rho[lane] = 1.0 / G[lane * nb + mh + lane];   // 1 / (s_h . y_h): the divisions leave the serial chain
  __syncwarp();
  double my_al = 0;   // lane q keeps the alpha of step q (every lane computes the same value)
  {
    int h = head == 0 ? mh - 1 : head - 1;
    for (int q = 0; q < nh; ++q) {
      const double al = vget(h) * s_rho[h];
      if (lane == q) my_al = al;
      axpy_col(-al, mh + h, mh + h);
      h = h == 0 ? mh - 1 : h - 1;
    }
  }
#pragma unroll
  for (int q = 0; q < 3; ++q) { v[q] *= gamma; dl[q] *= gamma; }
  {
    int h = head - nh; if (h < 0) h += mh;   // the oldest pair
    for (int q = nh - 1; q >= 0; --q) {
      const double al = __shfl_sync(0xffffffffu, my_al, q);
      const double b = vget(mh + h) * s_rho[h];
      axpy_col(al - b, h, h);
      h = h + 1 == mh ? 0 : h + 1;
    }
  }
  const double dfp = vget(rg);  // g . p
  int nact = 0;   // the direction as a compact list of (basis row, coefficient)
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const int j = lane + 32 * q;
    const bool on = j < nb && dl[q] != 0.0;
    const unsigned msk = __ballot_sync(0xffffffffu, on);
    if (on) {
      const int pos = nact + __popc(msk & ((1u << lane) - 1u));
      S.act_idx[pos] = j;
      S.act_coef[pos] = dl[q];
    }
    nact += __popc(msk);
  }
  if (!term) {
    const double rgd = -dfp / fmax(fabs(fk), 1.0);
    if (rgd < tol_rel_grad * eps) term = 6;
  }
  __syncwarp();
  if (lane != 0) return;
  S.dbg[5] = clock64();
  S.fk1 = fk1; S.fk = fk; S.gg = gg; S.dfp_prev = dfp_prev; S.alphak1 = alphak1; S.gamma = gamma;
  S.head = head; S.nh = nh; S.dfp = dfp; S.nact = nact;
  S.fresh_dir = 1;
  S.accept = 0;
  if (term) { S.term = term; S.phase = PH_DONE; return; }
  S.it = it + 1;
  S.resetB = 0;
  lb_ls_begin(S);
  S.dbg[6] = clock64();
}

struct LbStat { int32_t phase, term, it, fevals, switch_iter, accept, fresh_dir, pad0; double fk; double pad1[3]; };

}  // namespace

// Pinned host memory for the status words: ONE 32 KB block per process, cut into 512-byte slots and never freed --
// cudaMallocHost / cudaFreeHost take the driver's big locks and were measured at up to 0.4 s apiece on these boxes, which
// is more than a whole fit.  A slot belongs to a fit from its first L-BFGS run until the model is freed.
#include <mutex>
static std::mutex g_pin_mu;
static char *g_pin_block = nullptr;
static uint64_t g_pin_used = 0;   // bitmap of the 64 slots
static void *pinned_slot_acquire() {
  std::lock_guard<std::mutex> lk(g_pin_mu);
  if (!g_pin_block && cudaMallocHost(&g_pin_block, 64 * 512) != cudaSuccess) { cudaGetLastError(); g_pin_block = nullptr; return nullptr; }
  for (int i = 0; i < 64; ++i)
    if (!(g_pin_used >> i & 1ull)) { g_pin_used |= 1ull << i; return g_pin_block + 512 * i; }
  return nullptr;   // more than 64 fits with a live L-BFGS work space in one process
}
static void pinned_slot_release(void *p) {
  std::lock_guard<std::mutex> lk(g_pin_mu);
  if (p && g_pin_block) g_pin_used &= ~(1ull << (((char *)p - g_pin_block) / 512));
}

bool lbfgs_device_ok(const bvg_fit *f, const LbfgsOpt *opt) {
  if (opt && opt->on_iterate) return false;   // Pathfinder needs the host at every iterate
  if (f->comm_mode == 1) return false;        // NCCL: the all-reduce is a host call between kernels
  if (const char *e = getenv("BVG_LBFGS_HOST")) if (e[0] == '1') return false;
  return true;
}

static double lb_now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int lbfgs_run_device(bvg_fit *f, const double *d_theta0, double *d_out, const LbfgsOpt *opt) {
  const double t_enter = lb_now();
  double t_setup = 0, t_loop = 0, t_waited = 0;
  int n_batches = 0;
  const DevModel &m = f->m;
  const int64_t D = m.L.D;
  const int jac = opt ? opt->jacobian : 0, propto = opt ? opt->propto : 0;
  const int mh = std::min<int>(std::max(opt && opt->history > 0 ? opt->history : f->cfg.lbfgs_history, 1), LBFGS_MAX_HISTORY), nb = 2 * mh + 1;
  const int max_iter = opt && opt->max_iter > 0 ? opt->max_iter : f->cfg.mle_iter;
  cudaStream_t st = f->stream;
  double *B = nullptr, *x0, *p, *w, *part, *dots, *scal;
  LbDev *dev = nullptr;
  LbStat *h_stat = nullptr;
  double *h_scal = nullptr;
  cudaEvent_t evt[2] = {nullptr, nullptr};
  const size_t nd = (size_t)nb * D + 3 * (size_t)D + (size_t)nb * LB_NCHUNK * 3 + (size_t)nb * 3 + 8 + 2 * LB_GDOT_MAX;
  auto cleanup = [&]() {
    f->ls = LsPoint{nullptr, nullptr, nullptr};
    shadow_leave(f);
    // B, dev, the pinned words and the events belong to the fit's cache (lbfgs_cache_destroy)
  };
  auto fail = [&](int code) { cudaStreamSynchronize(st); cleanup(); return code; };
#define LB_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { bvg_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); return fail(BVG_ERR_CUDA); } } while (0)
#define LB_TRY(x) do { int r_ = (x); if (r_ != BVG_OK) return fail(r_); } while (0)
  {
    const size_t need = nd * sizeof(double) + 256 + sizeof(LbDev);
    if (f->lb_dev_bytes < need) {
      bvg_big_free(f->cfg.device, f->lb_dev_buf, f->lb_dev_bytes);
      f->lb_dev_buf = nullptr; f->lb_dev_bytes = 0;
      LB_TRY(bvg_big_alloc(f->cfg.device, &f->lb_dev_buf, need));
      f->lb_dev_bytes = need;
    }
    if (!f->lb_pinned) {
      f->lb_pinned = pinned_slot_acquire();
      if (!f->lb_pinned) { bvg_set_error("L-BFGS: no pinned host slot (cudaMallocHost failed, or more than 64 live fits)"); return fail(BVG_ERR_CUDA); }
      for (auto &e : f->lb_evt) LB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    B = (double *)f->lb_dev_buf;
    dev = (LbDev *)((char *)f->lb_dev_buf + ((nd * sizeof(double) + 255) & ~(size_t)255));
    h_stat = (LbStat *)f->lb_pinned;
    h_scal = (double *)((char *)f->lb_pinned + 256);
    evt[0] = f->lb_evt[0]; evt[1] = f->lb_evt[1];
  }
  LB_CUDA(cudaMemsetAsync(B, 0, nd * sizeof(double), st));
  x0 = B + (size_t)nb * D; p = x0 + D; w = p + D; part = w + D; dots = part + (size_t)nb * LB_NCHUNK * 3; scal = dots + nb * 3;
  double *gpart = scal + 8;
  double *g0 = B + (size_t)(2 * mh) * D;
  LB_TRY(vec_make_weights(st, w, m, f->cfg.rank));
  if (d_theta0) LB_CUDA(cudaMemcpyAsync(x0, d_theta0, D * sizeof(double), cudaMemcpyDeviceToDevice, st));

  // value, gradient and g.g at the point in x0 (host-synchronous: start of the run and the fp64 hand-over only)
  double fk = 0, gg = 0;
  auto eval_at_x0 = [&]() -> int {
    BVG_TRY(vec_axpy_to(st, m.zeta, x0, 0.0, p, D));
    BVG_TRY(eval_log_prob_device(f, MODE_GRAD_OUT, jac, propto, true));
    BVG_TRY(vec_accept(st, nullptr, nullptr, x0, g0, m.zeta, m.grad, D));
    BVG_TRY(vec_dot_sum(st, g0, g0, w, D, scal));
    if (f->comm_mode) BVG_TRY(comm_allreduce(f, scal, 2));
    CUDA_TRY(cudaMemcpyAsync(scal + 2, &m.ctl->lp, sizeof(double), cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(h_scal, scal, 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    fk = -h_scal[2]; gg = h_scal[0];
    return (std::isfinite(h_scal[2]) && std::isfinite(h_scal[0]) && std::isfinite(h_scal[1])) ? BVG_OK : BVG_ERR_NUMERIC;
  };
  int rc = eval_at_x0();
  if (rc == BVG_ERR_NUMERIC) bvg_set_error("L-BFGS: log density is not finite at the initial point");
  if (rc != BVG_OK) return fail(rc);
  int fevals_host = 1;

  LbDev *h = new LbDev();
  memset(h, 0, sizeof(LbDev));
  h->phase = PH_LS; h->mh = mh; h->nb = nb; h->max_iter = max_iter;
  if (const char *e = getenv("BVG_MLE_FP64_AT")) h->force_at = atoi(e);
  h->allow_shadow = (f->cfg.precision == BVG_PREC_FAST && !opt && !f->no_mle_fp64) ? 1 : 0;
  h->it = 1; h->resetB = 1; h->fk = fk; h->gg = gg; h->gamma = 1;
  int term = 0, it = 0, switch_iter = 0;
  if (max_iter > 0) {
    cudaError_t ce = cudaMemcpyAsync(dev, h, sizeof(LbDev), cudaMemcpyHostToDevice, st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    if (ce != cudaSuccess) { delete h; bvg_set_error("L-BFGS: upload of the control block failed: %s", cudaGetErrorString(ce)); return fail(BVG_ERR_CUDA); }
    const int vgrid = (int)std::min<int64_t>((D + 255) / 256, 148 * 8);
    const int ggrid = (int)std::max<int64_t>(1, std::min<int64_t>((D + 511) / 512, LB_GDOT_MAX));
    k_lb_begin<<<1, 1, 0, st>>>(dev);
    k_lb_dir<<<vgrid, 256, 0, st>>>(dev, p, B, x0, m.zeta, m.grad, D);
    const LsPoint lsp{x0, p, &dev->a_try};
    const size_t sh_ls = sizeof(double) * 2 * std::max(1, m.peer.world - 1);
    const size_t sh_dots = sizeof(double) * ((size_t)nb * nb + (size_t)nb * 3 * std::max(1, m.peer.world - 1));
    if (sh_dots > 48 * 1024) cudaFuncSetAttribute(k_lb_dots, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh_dots);
    auto tick = [&]() -> int {
      f->ls = lsp;   // k_prep forms zeta = x0 + a p
      const int erc = eval_log_prob_device(f, MODE_GRAD_OUT, jac, propto, true);
      f->ls = LsPoint{nullptr, nullptr, nullptr};
      BVG_TRY(erc);
      CUDA_TRY(launch_pdl(k_lb_gdot, dim3(ggrid), dim3(256), sh_ls, st, dev, f->m, (const double *)p, (const double *)w, gpart, scal));
      CUDA_TRY(launch_pdl(k_lb_dots, dim3((nb + LB_RG - 1) / LB_RG, LB_NCHUNK), dim3(256), sh_dots, st, dev, f->m, B, (const double *)x0, (const double *)w, part, dots));
      CUDA_TRY(launch_pdl(k_lb_dir, dim3(vgrid), dim3(256), (size_t)0, st, dev, p, B, x0, (const double *)m.zeta, (const double *)m.grad, D));
      f->st.kernel_launches += 3;
      CUDA_TRY(cudaGetLastError());
      return BVG_OK;
    };
    t_setup = lb_now();
    const long long max_ticks = (long long)max_iter * 64 + 4096;
    long long ticks = 0;
    bool pending[2] = {false, false};
    int slot = 0;
    bool done = false;
    while (!done) {
      for (int t = 0; t < LB_BATCH; ++t) if ((rc = tick()) != BVG_OK) { delete h; return fail(rc); }
      ticks += LB_BATCH;
      cudaError_t e1 = cudaMemcpyAsync(h_stat + slot, dev, sizeof(LbStat), cudaMemcpyDeviceToHost, st);
      if (e1 == cudaSuccess) e1 = cudaEventRecord(evt[slot], st);
      if (e1 != cudaSuccess) { delete h; bvg_set_error("L-BFGS: %s", cudaGetErrorString(e1)); return fail(BVG_ERR_CUDA); }
      pending[slot] = true;
      const int prev = slot ^ 1;
      slot ^= 1;
      if (!pending[prev]) continue;
      ++n_batches;
      const double tw0 = lb_now();
      const cudaError_t es_ = cudaEventSynchronize(evt[prev]);
      t_waited += lb_now() - tw0;
      if (es_ != cudaSuccess) { delete h; bvg_set_error("L-BFGS: device error in the search"); return fail(BVG_ERR_CUDA); }
      pending[prev] = false;
      const LbStat &s = h_stat[prev];
      if (s.phase == PH_DONE) { done = true; break; }
      if (s.phase == PH_NEED_SHADOW) {
        // the fast evaluator's noise floor: continue from this very point with fp64 evaluations, fresh curvature history
        cudaError_t e2 = cudaStreamSynchronize(st);
        pending[0] = pending[1] = false;
        if (e2 == cudaSuccess) e2 = cudaMemcpy(h, dev, sizeof(LbDev), cudaMemcpyDeviceToHost);
        if (e2 != cudaSuccess) { delete h; bvg_set_error("L-BFGS: %s", cudaGetErrorString(e2)); return fail(BVG_ERR_CUDA); }
        if ((rc = shadow_enter(f)) != BVG_OK) { delete h; return fail(rc); }
        switch_iter = h->it;
        rc = eval_at_x0();
        ++fevals_host;
        if (rc == BVG_ERR_NUMERIC) bvg_set_error("L-BFGS: log density is not finite where the fp64 evaluator takes over");
        if (rc != BVG_OK) { delete h; return fail(rc); }
        h->fk = fk; h->gg = gg; h->resetB = 2; h->shadow_on = 1; h->switch_iter = switch_iter; h->phase = PH_LS;
        e2 = cudaMemcpy(dev, h, sizeof(LbDev), cudaMemcpyHostToDevice);
        if (e2 != cudaSuccess) { delete h; bvg_set_error("L-BFGS: %s", cudaGetErrorString(e2)); return fail(BVG_ERR_CUDA); }
        k_lb_begin<<<1, 1, 0, st>>>(dev);
        k_lb_dir<<<vgrid, 256, 0, st>>>(dev, p, B, x0, m.zeta, m.grad, D);
        continue;
      }
      if (ticks > max_ticks) { delete h; bvg_set_error("L-BFGS: the device-resident search did not terminate"); return fail(BVG_ERR_STATE); }
    }
    cudaError_t e3 = cudaStreamSynchronize(st);
    t_loop = lb_now();
    if (e3 == cudaSuccess) e3 = cudaMemcpy(h, dev, sizeof(LbDev), cudaMemcpyDeviceToHost);
    if (e3 != cudaSuccess) { delete h; bvg_set_error("L-BFGS: %s", cudaGetErrorString(e3)); return fail(BVG_ERR_CUDA); }
    term = h->term; it = h->it; fk = h->fk; switch_iter = h->switch_iter;
    if (const char *e = getenv("BVG_LBFGS_TRACE")) if (e[0] == '1')
      fprintf(stderr, "k_lb_dots last CTA (cycles): partials+ticket %lld, chunk sums+exchange %lld, Gram %lld, scalars %lld, recursion %lld, next search %lld\n",
              h->dbg[1] - h->dbg[0], h->dbg[2] - h->dbg[1], h->dbg[3] - h->dbg[2], h->dbg[4] - h->dbg[3], h->dbg[5] - h->dbg[4], h->dbg[6] - h->dbg[5]);
    fevals_host += h->fevals;
  }
  delete h;
  LB_CUDA(cudaMemcpyAsync(d_out, x0, D * sizeof(double), cudaMemcpyDeviceToDevice, st));
  LB_CUDA(cudaStreamSynchronize(st));
  f->st.mle_iters = it;
  f->st.mle_fevals = fevals_host;
  f->st.mle_term = term;
  f->st.mle_fp64_from = switch_iter;
  f->st.mle_lp = -fk;
  cleanup();
  if (const char *e = getenv("BVG_LBFGS_TRACE")) if (e[0] == '1')
    fprintf(stderr, "lbfgs_run_device: setup %.2f ms, loop %.2f ms (%d batches, %.2f ms of it waiting for the device), teardown %.2f ms\n",
            1e3 * (t_setup - t_enter), 1e3 * (t_loop - t_setup), n_batches, 1e3 * t_waited, 1e3 * (lb_now() - t_loop));
  return BVG_OK;
#undef LB_CUDA
#undef LB_TRY
}

void lbfgs_cache_destroy(bvg_fit *f) {
  bvg_big_free(f->cfg.device, f->lb_dev_buf, f->lb_dev_bytes);
  f->lb_dev_buf = nullptr; f->lb_dev_bytes = 0;
  if (f->lb_pinned) {
    pinned_slot_release(f->lb_pinned);
    f->lb_pinned = nullptr;
    for (auto &e : f->lb_evt) { if (e) cudaEventDestroy(e); e = nullptr; }
  }
}
