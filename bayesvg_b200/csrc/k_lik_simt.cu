// k_lik_simt.cu -- CUDA-core likelihood kernel (fp64 parity path, and fp32 cross-check of the
// tcgen05 kernel).  One pass over Y per evaluation:
//     f_sg = sum_j Phi'_sj c'_jg                       (inst/approxGP.stan:33 + :48 mean; the ones
//                                                       column of Phi' carries beta0[g])
//     gaussian: r = y - f,  SSE_g += r^2               (approxGP.stan:48)
//     nb:       r = y - (y+phi) sigmoid(f - log phi)   (approxGP4.stan:48, d/d eta of
//               ll += y f - (y+phi) log1p_exp(f - log phi) - y log phi     neg_binomial_2_log)
//     P'_g += Phi'_s r   (k+1 values: Phi^T r and R_g = sum_s r)
// Decomposition: CTA = 128 genes x one contiguous range of spots; thread = gene, so every
// reduction over spots is a serial in-register sum (no shuffles, no atomics); the spot ranges
// ("splits") of one gene tile write separate partial rows that k_gene adds in fixed order.
#include "bvg_internal.cuh"

#define ST 32  // spots per staged tile

template <typename T> __device__ inline T t_exp(T x);
template <> __device__ inline float t_exp<float>(float x) { return __expf(x); }
template <> __device__ inline double t_exp<double>(double x) { return exp(x); }
template <typename T> __device__ inline T t_log1p(T x);
template <> __device__ inline float t_log1p<float>(float x) { return log1pf(x); }
template <> __device__ inline double t_log1p<double>(double x) { return log1p(x); }

template <typename T, int KP, int LIK, bool BWD>
__global__ void __launch_bounds__(BVG_GENE_TILE) k_lik_simt(const T *__restrict__ Y, int64_t ldY,
                                                            const T *__restrict__ Phi, const T *__restrict__ C,
                                                            T *__restrict__ Ppart, double *__restrict__ Spart, int64_t M,
                                                            int64_t G, int64_t Gpad, int chunks_per_split,
                                                            const double *__restrict__ zeta, int64_t tail_idx) {
  griddep_wait();    // the coefficient rows C come from k_prep
  griddep_launch();
  __shared__ T Ys[BVG_GENE_TILE][ST + 1];
  __shared__ __align__(16) T Ps[ST * KP];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t g0 = (int64_t)blockIdx.x * BVG_GENE_TILE, g = g0 + tid;
  const int split = blockIdx.y;
  T c[KP], P[KP];
#pragma unroll
  for (int j = 0; j < KP; ++j) { c[j] = C[g * KP + j]; P[j] = 0; }  // C has Gpad rows, pads are zero
  T acc0 = 0, acc1 = 0;
  T lt = 0, phi = 0;
  if (LIK == BVG_NB) { lt = (T)zeta[tail_idx]; phi = t_exp(lt); }
  const int64_t s_begin = (int64_t)split * chunks_per_split * ST;
  for (int ch = 0; ch < chunks_per_split; ++ch) {
    const int64_t s0 = s_begin + (int64_t)ch * ST;
    if (s0 >= M) break;
    // stage the Y tile (coalesced: one 32-spot run of one gene per warp instruction)
    for (int r = warp; r < BVG_GENE_TILE; r += BVG_GENE_TILE / 32) {
      const int64_t gr = g0 + r, s = s0 + lane;
      Ys[r][lane] = (gr < G && s < M) ? Y[gr * ldY + s] : (T)0;
    }
    for (int i = tid; i < ST * KP; i += BVG_GENE_TILE) Ps[i] = Phi[s0 * KP + i];  // Phi' has Mpad zero-padded rows
    __syncthreads();
    const int ns = (int)((M - s0) < ST ? (M - s0) : ST);
#pragma unroll 4
    for (int s = 0; s < ST; ++s) {
      const T y = Ys[tid][s];
      const T *p = Ps + s * KP;
      T f = 0;
#pragma unroll
      for (int j = 0; j < KP; ++j) f += p[j] * c[j];
      T r;
      if (LIK == BVG_GAUSSIAN) {
        r = y - f;
        acc0 += r * r;
      } else {
        const T d = f - lt;
        // L = log1p_exp(d), sg = sigmoid(d), numerically stable on both sides
        const T e = t_exp(-fabs(d));
        const T Lq = (d > 0 ? d : (T)0) + t_log1p(e);
        const T sg = (d >= 0 ? (T)1 : e) / ((T)1 + e);
        r = y - (y + phi) * sg;
        if (s < ns) {  // padded spots have a zero Phi' row (no P contribution) but a non-zero nb term
          acc0 += y * f - (y + phi) * Lq - y * lt;
          acc1 += Lq;
        }
      }
      if (BWD) {
#pragma unroll
        for (int j = 0; j < KP; ++j) P[j] += p[j] * r;
      }
    }
    __syncthreads();
  }
  const int64_t row = (int64_t)split * Gpad + g;
  if (BWD) {
#pragma unroll
    for (int j = 0; j < KP; ++j) Ppart[row * KP + j] = P[j];
  }
  Spart[row * 2 + 0] = (double)acc0;
  Spart[row * 2 + 1] = (double)acc1;
}

template <typename T, int KP>
static int launch_t(bvg_fit *f, bool bwd) {
  const DevModel &m = f->m;
  const int ntiles = (int)(m.Gpad / BVG_GENE_TILE);
  const int64_t nchunks = (m.M + ST - 1) / ST;
  const int cps = (int)((nchunks + m.nsplit - 1) / m.nsplit);
  dim3 grid(ntiles, m.nsplit);
#define LAUNCH(LIK, BWD)                                                                                      \
  CUDA_TRY(launch_pdl(k_lik_simt<T, KP, LIK, BWD>, grid, dim3(BVG_GENE_TILE), 0, f->stream, (const T *)f->Y, f->ldY, \
                      (const T *)f->Phi, (const T *)m.C, (T *)m.Ppart, (double *)m.Spart, m.M, m.G, m.Gpad, cps,     \
                      (const double *)m.zeta, m.L.tail))
  if (m.lik == BVG_GAUSSIAN) { if (bwd) LAUNCH(BVG_GAUSSIAN, true); else LAUNCH(BVG_GAUSSIAN, false); }
  else { if (bwd) LAUNCH(BVG_NB, true); else LAUNCH(BVG_NB, false); }
#undef LAUNCH
  f->st.kernel_launches++;
  return BVG_OK;
}


// ------------------------------------------------------------------------------------------------
// k_lik_gen: the same fused pass for a LARGE basis (KP > 32, up to 512: config 5 has k ~ 400), where a thread
// can no longer hold a whole coefficient row.  CTA = 32 genes x one contiguous range of spots, 256 threads,
// chunks of 32 spots, the basis dimension walked in blocks of 32:
//   forward : thread (gene gt, spot group sg) accumulates f for spots sg, sg+8, sg+16, sg+24 over the k-blocks
//             (Phi' block [32 spots][32] and C block [32][32 genes] staged in shared memory);
//   residual: as in k_lik_simt; r goes to shared memory;
//   backward: thread (gene gt, basis group jg) owns P'[j][gt] for j = jg, jg+8, ... (NKB*4 register accumulators
//             for the CTA's whole spot range), Phi' blocks staged again.
// It is the fp64 parity path and the fp32 fallback for k > 23; the tcgen05 kernel covers k <= 23.
template <typename T, int NKB, int LIK, bool BWD>
__global__ void __launch_bounds__(256) k_lik_gen(const T *__restrict__ Y, int64_t ldY, const T *__restrict__ Phi,
                                                 const T *__restrict__ C, T *__restrict__ Ppart, double *__restrict__ Spart,
                                                 int64_t M, int64_t G, int64_t Gpad, int KP, int chunks_per_split,
                                                 const double *__restrict__ zeta, int64_t tail_idx) {
  griddep_wait();
  griddep_launch();
  __shared__ T Ys[32][33];   // [gene][spot]
  __shared__ T Pk[32][33];   // [spot][basis within block]
  __shared__ T Ck[32][33];   // [basis within block][gene]
  __shared__ T Rs[32][33];   // [spot][gene]
  __shared__ T red[2][8][32];
  const int tid = threadIdx.x, gt = tid & 31, grp = tid >> 5;  // grp = spot group (forward) = basis group (backward)
  const int64_t g0 = (int64_t)blockIdx.x * 32;
  const int split = blockIdx.y;
  T Pacc[NKB * 4];
#pragma unroll
  for (int i = 0; i < NKB * 4; ++i) Pacc[i] = 0;
  T acc0 = 0, acc1 = 0, lt = 0, phi = 0;
  if (LIK == BVG_NB) { lt = (T)zeta[tail_idx]; phi = t_exp(lt); }
  const int nkb = (KP + 31) / 32;
  const int64_t s_begin = (int64_t)split * chunks_per_split * 32;
  for (int ch = 0; ch < chunks_per_split; ++ch) {
    const int64_t s0 = s_begin + (int64_t)ch * 32;
    if (s0 >= M) break;
    for (int r = grp; r < 32; r += 8) {  // Y tile: one 32-spot run of one gene per warp instruction
      const int64_t gr = g0 + r, s = s0 + gt;
      Ys[r][gt] = (gr < G && s < M) ? Y[gr * ldY + s] : (T)0;
    }
    T f[4] = {0, 0, 0, 0};
    for (int kb = 0; kb < nkb; ++kb) {
      __syncthreads();  // previous users of Pk / Ck are done (and Ys is complete on the first pass)
      for (int r = grp; r < 32; r += 8) {
        const int j = kb * 32 + gt;
        Pk[r][gt] = j < KP ? Phi[(s0 + r) * KP + j] : (T)0;   // Phi' has zero-padded rows up to Mpad
        Ck[gt][r] = j < KP ? C[(g0 + r) * KP + j] : (T)0;     // C has Gpad rows, pads are zero
      }
      __syncthreads();
#pragma unroll 8
      for (int j = 0; j < 32; ++j) {
        const T c = Ck[j][gt];
#pragma unroll
        for (int i = 0; i < 4; ++i) f[i] += Pk[grp + 8 * i][j] * c;
      }
    }
    const int ns = (int)((M - s0) < 32 ? (M - s0) : 32);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int s = grp + 8 * i;
      const T y = Ys[gt][s];
      T r;
      if (LIK == BVG_GAUSSIAN) {
        r = y - f[i];
        acc0 += r * r;
      } else {
        const T d = f[i] - lt;
        const T e = t_exp(-fabs(d));
        const T Lq = (d > 0 ? d : (T)0) + t_log1p(e);
        const T sg = (d >= 0 ? (T)1 : e) / ((T)1 + e);
        r = y - (y + phi) * sg;
        if (s < ns) {
          acc0 += y * f[i] - (y + phi) * Lq - y * lt;
          acc1 += Lq;
        }
      }
      Rs[s][gt] = r;
    }
    if (BWD) {
#pragma unroll
      for (int kb = 0; kb < NKB; ++kb) {
        if (kb < nkb) {
          __syncthreads();  // Rs complete (first pass) / previous Pk consumed
          for (int r = grp; r < 32; r += 8) {
            const int j = kb * 32 + gt;
            Pk[r][gt] = j < KP ? Phi[(s0 + r) * KP + j] : (T)0;
          }
          __syncthreads();
#pragma unroll 8
          for (int s = 0; s < 32; ++s) {
            const T r = Rs[s][gt];
#pragma unroll
            for (int i = 0; i < 4; ++i) Pacc[kb * 4 + i] += Pk[s][grp + 8 * i] * r;
          }
        }
      }
    }
    __syncthreads();
  }
  const int64_t g = g0 + gt;
  const int64_t row = (int64_t)split * Gpad + g;
  if (BWD) {
#pragma unroll
    for (int kb = 0; kb < NKB; ++kb)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int j = kb * 32 + grp + 8 * i;
        if (j < KP) Ppart[row * KP + j] = Pacc[kb * 4 + i];
      }
  }
  red[0][grp][gt] = acc0;
  red[1][grp][gt] = acc1;
  __syncthreads();
  if (grp == 0) {
    T a = 0, b = 0;
    for (int q = 0; q < 8; ++q) { a += red[0][q][gt]; b += red[1][q][gt]; }  // fixed order
    Spart[row * 2 + 0] = (double)a;
    Spart[row * 2 + 1] = (double)b;
  }
}

template <typename T, int NKB>
static int launch_gen(bvg_fit *f, bool bwd) {
  const DevModel &m = f->m;
  const int64_t nchunks = (m.M + 31) / 32;
  const int cps = (int)((nchunks + m.nsplit - 1) / m.nsplit);
  dim3 grid((unsigned)(m.Gpad / 32), (unsigned)m.nsplit);
#define LAUNCHG(LIK, BWD)                                                                                          \
  CUDA_TRY(launch_pdl(k_lik_gen<T, NKB, LIK, BWD>, grid, dim3(256), 0, f->stream, (const T *)f->Y, f->ldY, (const T *)f->Phi, \
                      (const T *)m.C, (T *)m.Ppart, (double *)m.Spart, m.M, m.G, m.Gpad, m.KP, cps, (const double *)m.zeta, m.L.tail))
  if (m.lik == BVG_GAUSSIAN) { if (bwd) LAUNCHG(BVG_GAUSSIAN, true); else LAUNCHG(BVG_GAUSSIAN, false); }
  else { if (bwd) LAUNCHG(BVG_NB, true); else LAUNCHG(BVG_NB, false); }
#undef LAUNCHG
  f->st.kernel_launches++;
  return BVG_OK;
}

template <typename T>
static int launch_kp(bvg_fit *f, bool bwd) {
  if (f->m.KP > 32) {
    const int nkb = (f->m.KP + 31) / 32;
    if (nkb <= 2) return launch_gen<T, 2>(f, bwd);
    if (nkb <= 4) return launch_gen<T, 4>(f, bwd);
    if (nkb <= 8) return launch_gen<T, 8>(f, bwd);
    if (nkb <= 16) return launch_gen<T, 16>(f, bwd);
    bvg_set_error("unsupported padded basis width %d", f->m.KP);
    return BVG_ERR_ARG;
  }
  switch (f->m.KP) {
    case 8: return launch_t<T, 8>(f, bwd);
    case 16: return launch_t<T, 16>(f, bwd);
    case 24: return launch_t<T, 24>(f, bwd);
    case 32: return launch_t<T, 32>(f, bwd);
  }
  bvg_set_error("unsupported padded basis width %d", f->m.KP);
  return BVG_ERR_ARG;
}

int launch_lik_simt(bvg_fit *f, bool bwd) {
  return f->cfg.precision == BVG_PREC_FP64 ? launch_kp<double>(f, bwd) : launch_kp<float>(f, bwd);
}
