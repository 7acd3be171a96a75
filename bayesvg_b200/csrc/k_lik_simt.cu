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
                                                            T *__restrict__ Ppart, T *__restrict__ Spart, int64_t M,
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
  Spart[row * 2 + 0] = acc0;
  Spart[row * 2 + 1] = acc1;
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
                      (const T *)f->Phi, (const T *)m.C, (T *)m.Ppart, (T *)m.Spart, m.M, m.G, m.Gpad, cps,     \
                      (const double *)m.zeta, m.L.tail))
  if (m.lik == BVG_GAUSSIAN) { if (bwd) LAUNCH(BVG_GAUSSIAN, true); else LAUNCH(BVG_GAUSSIAN, false); }
  else { if (bwd) LAUNCH(BVG_NB, true); else LAUNCH(BVG_NB, false); }
#undef LAUNCH
  f->st.kernel_launches++;
  return BVG_OK;
}

template <typename T>
static int launch_kp(bvg_fit *f, bool bwd) {
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
