// k_lik_tc.cu -- the fused likelihood kernel for sm_100a: TMA-fed tcgen05 (BF16x3 split) contractions
// with TMEM accumulators AND TMEM-resident A operands, one pass over Y per evaluation.
//
// Per CTA: one tile of 128 genes (= the 128 TMEM lanes) x one contiguous range of spots, walked in
// chunks of 64 spots.  For every chunk
//   MMA1  D1[128 genes x 64 spots] = C'[128 x 32] . Phi'_chunk[64 spots x 32]^T       (forward:  f = beta0 + A * Phi alpha;
//                                                                                      inst/approxGP.stan:33 and :48 of the reference)
//   epilogue (8 warps, thread = gene row x 32 spots): f <- TMEM, y <- smem (the TMA-written, 128B-
//         swizzled fp32 Y tile), r = y - f (gaussian) | y - (y+phi) sigmoid(f - log phi) (nb), SSE /
//         log-lik partials in registers (serial per thread: no shuffles, no atomics), r -> TMEM as two
//         packed bf16 parts with tcgen05.st
//   MMA2  D2[128 genes x 32]      += R[128 x 64 spots] . Phi'^T_chunk[32 x 64 spots]^T (backward: Phi^T r and R_g = 1^T r)
// Both 128-row operands (coefficient rows C' and residuals R) live in TENSOR MEMORY (A-from-TMEM form
// of tcgen05.mma); shared memory only carries what TMA brings in.  D2 stays in TMEM for the CTA's whole
// spot range and is read once.
// "BF16x3": x = x1 + x2 (x1 = bf16(x), x2 = bf16(x - x1)); every product is x1*y1 + x2*y1 + x1*y2, i.e.
// ~2^-16 relative error per product with fp32 accumulation -- and kind::f16 moves K = 16 per instruction
// at twice the TF32 MAC rate.  History (profiles/README.md): 3xTF32 with 32-spot chunks needed 24 small
// MMAs per 32 spots and ran at the tensor pipe's limit for that shape (~30 cycles per 128x32x8 MMA,
// 19 us per launch at cfg2); this version needs 9 per 32 spots.
//
// Warp roles (384 threads, 1 CTA/SM): warp 0 = TMA producer of the Y ring, warp 11 = TMA producer of the
// basis ring, warp 1 = TMEM owner + forward-MMA issuer, warp 2 = backward-MMA issuer, warps 3..10 =
// epilogue (two warps per TMEM lane quarter, one 32-spot half each).  MMA warps stay convergent and
// issue from one elected lane (elect.sync).
// Pipelines: Y ring full/empty (5 x 32 KB, released by the epilogue as soon as y is in registers), basis
// ring full/empty (2 x 24 KB, released by the MMA that read it), D1 full/empty (2 TMEM buffers), R full/free (2).
#include <cuda.h>
#include <string.h>
#include <cuda_bf16.h>

#include <type_traits>

#include "bvg_internal.cuh"

#define TC_THREADS 384
#define TC_NY 3       // Y ring: a stage lives from its TMA issue until the epilogue has y in registers (3 x 32 KB in flight per SM = 14 MB chip-wide)
#define TC_NB 5       // basis ring (24 KB per stage): a stage lives from TMA issue to the END of MMA2 of its chunk (~3,300 cycles); with 3 stages that lifetime, not the tensor pipe, set the chunk period (1,060 cycles measured; tcgen05 floor 480)
#define TC_CHUNK 64   // spots per chunk
#define TC_KW 32      // coefficient row width (k+1 <= 24 padded to 32)
#define YT_BYTES (BVG_GENE_TILE * 128)   // 16 KB: 128 genes x 32 spots fp32
#define FT_BYTES (64 * 128)              // 8 KB: 64 spots x 64 bf16 (first 32 columns = Phi' row)
#define TT_BYTES (32 * 128)              // 4 KB: 32 basis rows x 64 spots bf16
#define YSTAGE_BYTES (2 * YT_BYTES)
#define BSTAGE_BYTES (2 * FT_BYTES + 2 * TT_BYTES)
#define TC_SMEM_BYTES (TC_NY * YSTAGE_BYTES + TC_NB * BSTAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/)
// TMEM column map (512 allocated)
#define TM_C1 0      // 16 columns: bf16 part 1 of the coefficient rows (32 values, packed in pairs)
#define TM_C2 16
#define TM_D1 64     // + 64 * buf : forward accumulator, 64 spots
#define TM_R 192     // + 64 * buf : residual part 1 at +0 (32 columns = 64 spots packed), part 2 at +32
#define TM_D2 320    // 32 columns: backward accumulator
#define TM_COLS 512

struct TcArgs {
  float *Ppart;
  double *Spart;   // fp64: a float store of the fp64 running sums cost ~1e-8 of the log density (L-BFGS noise floor)
  // coefficient rows c'_g = (A_g alpha_t[,g], beta0_g) are built here from the current draw (approxGP.stan:26-27, :32),
  // published in fp32 by k_prep / k_gene: W[g] = (z_alpha_t[,g], z_beta0[g], A_g), U = (mu_alpha | sigma_alpha | mu_beta0, sigma_beta0)
  const float *W, *U;
  int k;
  int64_t M, G, Gpad;
  int KP, cps, nchunks, nk1, nch32;  // nk1 = ceil((k+1)/16): K steps of the forward MMA; nch32 = ldY / 32
  const double *zeta;
  int64_t tail_idx;
  // nb: the lgamma / digamma terms of the log density depend only on (distinct count c, phi_nb); the idle lanes of
  // the basis-producer warp evaluate them here, off the serial finish of k_gene (SURVEY.md A.3 histogram trick)
  const double *hist_val, *hist_cnt;
  double *hterm;
  int nhist;
  long long *trace;  // optional [chunk][8] clock64 stamps of CTA (0,0) (diagnostics)
};

struct TcState {
  CUtensorMap tmY, tmF1, tmF2, tmT1, tmT2;
  __nv_bfloat16 *F1, *F2, *T1, *T2;
};

#include "tc_ptx.cuh"

// ------------------------------------------------------------------------------------------------
template <int LIK, bool BWD>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_lik_tc(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmF1,
         const __grid_constant__ CUtensorMap tmF2, const __grid_constant__ CUtensorMap tmT1,
         const __grid_constant__ CUtensorMap tmT2, TcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const bool tr0 = a.trace && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0;
  if (tr0) { a.trace[8 * 30 + 0] = clock64(); asm volatile("mov.u64 %0, %globaltimer;" : "=l"(a.trace[8 * 30 + 4])); }
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024 B alignment
  const uint32_t sB0 = base + TC_NY * YSTAGE_BYTES, sBar = sB0 + TC_NB * BSTAGE_BYTES;
  // barrier map (8 B each)
  const uint32_t bFullY = sBar, bEmptyY = bFullY + 8 * TC_NY, bFullB = bEmptyY + 8 * TC_NY, bEmptyB = bFullB + 8 * TC_NB;
  const uint32_t bD1Full = bEmptyB + 8 * TC_NB, bD1Empty = bD1Full + 16, bRFull = bD1Empty + 16, bRFree = bRFull + 16;
  const uint32_t bCFull = bRFree + 16, bD2Full = bCFull + 8, sTmemPtr = bD2Full + 8;
  __shared__ double s_half[2][BVG_GENE_TILE];
  auto stY = [&](int s, int h) { return base + s * YSTAGE_BYTES + h * YT_BYTES; };
  auto stF = [&](int s, int w) { return sB0 + s * BSTAGE_BYTES + w * FT_BYTES; };
  auto stT = [&](int s, int w) { return sB0 + s * BSTAGE_BYTES + 2 * FT_BYTES + w * TT_BYTES; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, split = blockIdx.y;
  const int c_begin = split * a.cps;
  int n = a.nchunks - c_begin;
  n = n < a.cps ? n : a.cps;
  if (n < 0) n = 0;
  constexpr uint32_t NEPI = 256;

  griddep_launch();  // PDL: the successor (k_gene) may begin launching; it waits for this grid to finish
  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_NY; ++s) { mbar_init(bFullY + 8 * s, 1); mbar_init(bEmptyY + 8 * s, NEPI); }
    for (int s = 0; s < TC_NB; ++s) { mbar_init(bFullB + 8 * s, 1); mbar_init(bEmptyB + 8 * s, 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bD1Full + 8 * b, 1);
      mbar_init(bD1Empty + 8 * b, NEPI);
      mbar_init(bRFull + 8 * b, NEPI);
      mbar_init(bRFree + 8 * b, 1);
    }
    mbar_init(bCFull, NEPI);
    mbar_init(bD2Full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sTmemPtr), "r"((uint32_t)TM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(sTmemPtr));
  if (tr0) a.trace[8 * 30 + 1] = clock64();

  if (warp == 0) {
    // ===================== TMA producer, Y ring (Y is written by no kernel: no PDL wait) =====
    if (lane == 0) {
      for (int c = 0; c < n; ++c) {
        const int s = c % TC_NY;
        if (c >= TC_NY) mbar_wait(bEmptyY + 8 * s, ((c / TC_NY) - 1) & 1);
        mbar_expect_tx(bFullY + 8 * s, YSTAGE_BYTES);
        // one 32 KB contiguous run of the tile-major Y: rows [0,128) = spots spot0..+31, rows [128,256) = +32..+63
        tma_load_2d(stY(s, 0), &tmY, bFullY + 8 * s, 0, (tile * a.nch32 + 2 * (c_begin + c)) * BVG_GENE_TILE);
      }
    }
  } else if (warp == 11) {
    // ===================== TMA producer, basis ring (lane 0); nb histogram terms (lanes 1..31) =====
    auto produce = [&](int c0, int c1) {
      for (int c = c0; c < c1; ++c) {
        const int s = c % TC_NB;
        if (c >= TC_NB) mbar_wait(bEmptyB + 8 * s, ((c / TC_NB) - 1) & 1);
        const int spot0 = (c_begin + c) * TC_CHUNK;
        mbar_expect_tx(bFullB + 8 * s, 2 * FT_BYTES + (BWD ? 2 * TT_BYTES : 0));
        tma_load_2d(stF(s, 0), &tmF1, bFullB + 8 * s, 0, spot0);
        tma_load_2d(stF(s, 1), &tmF2, bFullB + 8 * s, 0, spot0);
        if (BWD) {
          tma_load_2d(stT(s, 0), &tmT1, bFullB + 8 * s, spot0, 0);
          tma_load_2d(stT(s, 1), &tmT2, bFullB + 8 * s, spot0, 0);
        }
      }
    };
    const int n0 = n < TC_NB ? n : TC_NB;
    if (lane == 0) produce(0, n0);  // the whole ring is in flight before anything else happens in this warp
    if (LIK == BVG_NB && a.hterm) {
      // the other 31 lanes are idle: each evaluates the lgamma / digamma term of one distinct count (the ring above
      // gives lane 0 several chunk-times of slack while it waits for them)
      __syncwarp();
      const int e = (int)((blockIdx.y * gridDim.x + blockIdx.x) * 31 + lane - 1);
      if (lane > 0 && e < a.nhist) {
        griddep_wait();  // phi_nb of this draw comes from k_prep / the previous k_gene
        const double phi = exp(a.zeta[a.tail_idx]);
        const double cv = a.hist_val[e], nv = a.hist_cnt[e];
        a.hterm[2 * e] = nv * (lgamma(cv + phi) - lgamma(phi));
        a.hterm[2 * e + 1] = nv * (dev_digamma(cv + phi) - dev_digamma(phi));
      }
      __syncwarp();
    }
    if (lane == 0) produce(n0, n);
  } else if (warp == 1) {
    // ===================== forward MMA issuer (whole warp convergent, one elected lane issues) =====
    constexpr uint32_t idesc = make_idesc_bf16(128, 64);
    const bool tr = a.trace && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0;
    if (n > 0) mbar_wait(bCFull, 0);
    for (int c = 0; c < n; ++c) {
      const int s = c % TC_NB, b = c & 1;
      if (tr) a.trace[c * 8 + 0] = clock64();
      mbar_wait(bFullB + 8 * s, (c / TC_NB) & 1);
      if (c >= 2) mbar_wait(bD1Empty + 8 * b, ((c >> 1) - 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t d1 = tmem + TM_D1 + 64 * b;
        const uint64_t f1 = make_desc(stF(s, 0)), f2 = make_desc(stF(s, 1));
        for (int kk = 0; kk < a.nk1; ++kk) tc_mma_bf16_ts(d1, tmem + TM_C1 + 8 * kk, f1 + 2 * kk, idesc, kk ? 1u : 0u);
        for (int kk = 0; kk < a.nk1; ++kk) tc_mma_bf16_ts(d1, tmem + TM_C2 + 8 * kk, f1 + 2 * kk, idesc, 1u);
        for (int kk = 0; kk < a.nk1; ++kk) tc_mma_bf16_ts(d1, tmem + TM_C1 + 8 * kk, f2 + 2 * kk, idesc, 1u);
        tc_commit(bD1Full + 8 * b);
        if (!BWD) tc_commit(bEmptyB + 8 * s);  // forward-only: the basis stage is free once MMA1 has read it
      }
      __syncwarp();
      if (tr) a.trace[c * 8 + 1] = clock64();
    }
  } else if (warp == 2) {
    // ===================== backward MMA issuer =====================
    if (BWD) {
      constexpr uint32_t idesc = make_idesc_bf16(128, 32);
      const uint32_t d2 = tmem + TM_D2;
      const bool tr = a.trace && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0;
      for (int c = 0; c < n; ++c) {
        const int s = c % TC_NB, b = c & 1;
        mbar_wait(bFullB + 8 * s, (c / TC_NB) & 1);      // basis tiles of this stage (acquire)
        mbar_wait(bRFull + 8 * b, (c >> 1) & 1);         // residuals published by the epilogue
        tc_fence_after();
        if (elect_one()) {
          const uint32_t r1 = tmem + TM_R + 64 * b, r2 = r1 + 32;
          const uint64_t t1 = make_desc(stT(s, 0)), t2 = make_desc(stT(s, 1));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) tc_mma_bf16_ts(d2, r1 + 8 * kk, t1 + 2 * kk, idesc, (c | kk) ? 1u : 0u);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) tc_mma_bf16_ts(d2, r2 + 8 * kk, t1 + 2 * kk, idesc, 1u);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) tc_mma_bf16_ts(d2, r1 + 8 * kk, t2 + 2 * kk, idesc, 1u);
          tc_commit(bEmptyB + 8 * s);  // basis stage free for its producer
          tc_commit(bRFree + 8 * b);   // residual buffer free for the epilogue of chunk c + 2
          if (c == n - 1) tc_commit(bD2Full);
        }
        __syncwarp();
        if (tr) a.trace[c * 8 + 2] = clock64();
      }
    }
  } else {
    // ===================== epilogue: 8 warps =====================
    const int e = warp - 3, q = warp & 3, h = e >> 2;  // TMEM lane quarter of this warp, 32-spot half of the chunk
    const int r = 32 * q + lane;                       // gene row inside the tile = TMEM lane
    const uint32_t trow = tmem + ((uint32_t)(32 * q) << 16);
    const int64_t g = (int64_t)tile * BVG_GENE_TILE + r;
    {  // coefficient rows -> TMEM (A operand of MMA1), once per CTA, as two packed bf16 parts
      // PDL: everything above (barriers, TMEM, and the producer's first TMA loads) overlaps the tail of
      // k_prep; its output C is first touched here
      griddep_wait();
      float cv[16];
      {
        const float4 *pw = reinterpret_cast<const float4 *>(a.W + g * TC_KW + 16 * h);
        const float4 *pm = reinterpret_cast<const float4 *>(a.U + 16 * h), *ps = reinterpret_cast<const float4 *>(a.U + 32 + 16 * h);
        const float Ag = a.W[g * TC_KW + a.k + 1], mb = a.U[64], sb = a.U[65];
        float w[16], um[16], us[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 x = pw[i], y = __ldg(pm + i), z = __ldg(ps + i);
          w[4 * i] = x.x; w[4 * i + 1] = x.y; w[4 * i + 2] = x.z; w[4 * i + 3] = x.w;
          um[4 * i] = y.x; um[4 * i + 1] = y.y; um[4 * i + 2] = y.z; um[4 * i + 3] = y.w;
          us[4 * i] = z.x; us[4 * i + 1] = z.y; us[4 * i + 2] = z.z; us[4 * i + 3] = z.w;
        }
        const bool live = g < a.G;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int j = 16 * h + i;
          float v = 0.f;
          if (j < a.k) v = Ag * fmaf(us[i], w[i], um[i]);
          else if (j == a.k) v = fmaf(sb, w[i], mb);
          cv[i] = live ? v : 0.f;
        }
      }
      uint32_t c1[8], c2[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) split_bf16x2(cv[2 * i], cv[2 * i + 1], c1[i], c2[i]);
      tmem_st8u(trow + TM_C1 + 8 * h, c1);
      tmem_st8u(trow + TM_C2 + 8 * h, c2);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bCFull);
    }
    // per-chunk sums (32 terms) in fp32, running sums over the CTA's spot range in fp64: a serial fp32 sum over
    // thousands of spots loses ~1e-5 of the log density, enough to stall the L-BFGS line search at config-4 scale
    double sum0 = 0.0, sum1 = 0.0;
    float lt = 0.f, phi = 0.f;
    if (LIK == BVG_NB) { lt = (float)a.zeta[a.tail_idx]; phi = __expf(lt); }
    const bool tr = a.trace && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 96;
    for (int c = 0; c < n; ++c) {
      const int s = c % TC_NY, b = c & 1;
      mbar_wait(bFullY + 8 * s, (c / TC_NY) & 1);      // Y tile landed (acquire for this thread)
      if (tr) a.trace[c * 8 + 3] = clock64();
      mbar_wait(bD1Full + 8 * b, (c >> 1) & 1);        // MMA1 of this chunk retired
      tc_fence_after();
      if (tr) a.trace[c * 8 + 4] = clock64();
      float f[32];
      tmem_ld32(trow + TM_D1 + 64 * b + 32 * h, f);
      if (tr) a.trace[c * 8 + 5] = clock64();
      tc_fence_before();
      mbar_arrive(bD1Empty + 8 * b);
      const uint32_t yrow = stY(s, h) + r * 128;
      const int64_t spot0 = (int64_t)(c_begin + c) * TC_CHUNK + 32 * h;
      uint32_t p1[16], p2[16];
      float acc0 = 0.f, acc1 = 0.f;
      // masked = std::true_type only for the one chunk that straddles M (nb: padded spots have a non-zero term)
      auto body = [&](auto masked, int nvalid) {
        constexpr bool MASKED = decltype(masked)::value;
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          const uint32_t off = (uint32_t)((c4 ^ (r & 7)) << 4);  // 128B swizzle: 16B chunk index ^ (row & 7)
          const float4 y4 = lds128(yrow + off);
          const float yv[4] = {y4.x, y4.y, y4.z, y4.w};
          float res[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float ff = f[4 * c4 + i];
            if (LIK == BVG_GAUSSIAN) {
              res[i] = yv[i] - ff;
              acc0 = fmaf(res[i], res[i], acc0);
            } else {
              // log1p_exp(d) and sigmoid(d) from ONE exp2, one reciprocal and one log2 of t = 1 + e^{-|d|} in (1, 2]:
              // three MUFU operations and ~12 FMA-pipe instructions per element (log1pf + IEEE division cost ~4x
              // that); absolute error ~1e-7 per element, far inside the 1e-3 gate of the fast path.
              //   ll_sg = y eta - (y + phi) L - y log phi = y d - (y + phi) L,  d = eta - log phi   (approxGP4.stan:48)
              const float d = ff - lt;
              const float ex = ex2_approx(-fabsf(d) * 1.4426950408889634f);
              const float t = 1.f + ex;
              const float rt = rcp_approx(t);
              const float Lq = fmaf(lg2_approx(t), 0.6931471805599453f, fmaxf(d, 0.f));
              const float sg = (d >= 0.f ? 1.f : ex) * rt;
              const float yp = yv[i] + phi;
              res[i] = fmaf(-yp, sg, yv[i]);
              if (!MASKED || 4 * c4 + i < nvalid) {
                acc0 = fmaf(yv[i], d, acc0);
                acc0 = fmaf(-yp, Lq, acc0);
                acc1 += Lq;
              }
            }
          }
          split_bf16x2(res[0], res[1], p1[2 * c4], p2[2 * c4]);
          split_bf16x2(res[2], res[3], p1[2 * c4 + 1], p2[2 * c4 + 1]);
        }
      };
      const int64_t left = a.M - spot0;
      if (LIK == BVG_GAUSSIAN || left >= 32) body(std::false_type{}, 32);
      else body(std::true_type{}, left > 0 ? (int)left : 0);
      mbar_arrive(bEmptyY + 8 * s);  // y is in registers: the Y stage can be refilled
      sum0 += (double)acc0;
      if (LIK == BVG_NB) sum1 += (double)acc1;
      if (BWD) {
        if (tr) a.trace[c * 8 + 6] = clock64();
        if (c >= 2) mbar_wait(bRFree + 8 * b, ((c >> 1) - 1) & 1);  // MMA2 of chunk c-2 has consumed this buffer
        tc_fence_after();
        tmem_st16u(trow + TM_R + 64 * b + 16 * h, p1);
        tmem_st16u(trow + TM_R + 64 * b + 32 + 16 * h, p2);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(bRFull + 8 * b);
        if (tr) a.trace[c * 8 + 7] = clock64();
      }
    }
    // ---- drain: P' = D2 (k+1 values per gene) and the scalar partials ----
    const int64_t row = (int64_t)split * a.Gpad + g;
    if (BWD) {
      float p[16];
      if (n > 0) {
        mbar_wait(bD2Full, 0);
        tc_fence_after();
        tmem_ld16(trow + TM_D2 + 16 * h, p);
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) p[i] = 0.f;
      }
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (16 * h + i < a.KP) a.Ppart[row * a.KP + 16 * h + i] = p[i];
    }
    if (h == 1) { s_half[0][r] = sum0; s_half[1][r] = sum1; }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (h == 0) {
      a.Spart[row * 2 + 0] = sum0 + s_half[0][r];
      a.Spart[row * 2 + 1] = sum1 + s_half[1][r];
    }
  }
  if (tr0) a.trace[8 * 30 + 2] = clock64();
  tc_fence_before();
  __syncthreads();
  if (tr0) { a.trace[8 * 30 + 3] = clock64(); asm volatile("mov.u64 %0, %globaltimer;" : "=l"(a.trace[8 * 30 + 5])); }
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)TM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// host side: operand staging, tensor maps, launch
// ------------------------------------------------------------------------------------------------
// Phi (M x k col-major fp64) -> bf16 part 1 / part 2 of Phi' in both orientations:
//   F*: [Mpad][64]  (row = spot, 128 B; columns 0..k = Phi' row, rest zero)   B operand of the forward MMA
//   T*: [32][Mpad]  (row = basis index, spots contiguous)                      B operand of the backward MMA
__global__ void k_tc_stage_phi(const double *__restrict__ phi, int64_t M, int64_t Mpad, int k, __nv_bfloat16 *__restrict__ F1,
                               __nv_bfloat16 *__restrict__ F2, __nv_bfloat16 *__restrict__ T1, __nv_bfloat16 *__restrict__ T2) {
  const int64_t n = Mpad * 64;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = i / 64;
    const int j = (int)(i - s * 64);
    float v = 0.f;
    if (s < M && j <= k) v = j < k ? (float)phi[(int64_t)j * M + s] : 1.f;
    const __nv_bfloat16 b1 = __float2bfloat16_rn(v), b2 = __float2bfloat16_rn(v - __bfloat162float(b1));
    F1[i] = b1; F2[i] = b2;
    if (j < 32) { T1[(int64_t)j * Mpad + s] = b1; T2[(int64_t)j * Mpad + s] = b2; }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}
// 2-D tensor [rows][cols] (cols contiguous, row pitch ld elements), box = box_rows x 128 bytes, SWIZZLE_128B
int tc_make_map(CUtensorMap *m, void *ptr, bool bf16, uint64_t cols, uint64_t rows, uint64_t ld, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { bvg_set_error("cuTensorMapEncodeTiled is not available from this driver"); return BVG_ERR_CUDA; }
  const size_t es_bytes = bf16 ? 2 : 4;
  cuuint64_t dims[2] = {cols, rows}, strides[1] = {ld * es_bytes};
  cuuint32_t box[2] = {(cuuint32_t)(128 / es_bytes), box_rows}, es[2] = {1, 1};
  const CUresult r = enc(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { bvg_set_error("cuTensorMapEncodeTiled failed with %d", (int)r); return BVG_ERR_CUDA; }
  return BVG_OK;
}

bool tc_available() { return true; }

int tc_setup(bvg_fit *f) {
  const DevModel &m = f->m;
  if (m.k + 1 > 24) { bvg_set_error("tcgen05 engine: k + 1 = %d exceeds 24", m.k + 1); return BVG_ERR_ARG; }
  TcState *t = new TcState();
  f->tc = t;
  const size_t nF = (size_t)f->Mpad * 64, nT = (size_t)32 * f->Mpad;
  CUDA_TRY(cudaMalloc(&t->F1, sizeof(__nv_bfloat16) * (2 * nF + 2 * nT)));
  t->F2 = t->F1 + nF; t->T1 = t->F2 + nF; t->T2 = t->T1 + nT;
  k_tc_stage_phi<<<148 * 2, 256, 0, f->stream>>>(f->phi64, m.M, f->Mpad, m.k, t->F1, t->F2, t->T1, t->T2);
  CUDA_TRY(cudaGetLastError());
  // Y is tile-major: rows of 32 spots, 128 gene rows per (tile, 32-spot chunk); one box = one 64-spot stage
  BVG_TRY(tc_make_map(&t->tmY, f->Y, false, 32, (uint64_t)(m.Gpad / BVG_GENE_TILE) * (uint64_t)(f->ldY / 32) * BVG_GENE_TILE, 32, 2 * BVG_GENE_TILE));
  BVG_TRY(tc_make_map(&t->tmF1, t->F1, true, 64, (uint64_t)f->Mpad, 64, 64));
  BVG_TRY(tc_make_map(&t->tmF2, t->F2, true, 64, (uint64_t)f->Mpad, 64, 64));
  BVG_TRY(tc_make_map(&t->tmT1, t->T1, true, (uint64_t)f->Mpad, 32, (uint64_t)f->Mpad, 32));
  BVG_TRY(tc_make_map(&t->tmT2, t->T2, true, (uint64_t)f->Mpad, 32, (uint64_t)f->Mpad, 32));
  CUDA_TRY(cudaFuncSetAttribute(k_lik_tc<BVG_GAUSSIAN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
  CUDA_TRY(cudaFuncSetAttribute(k_lik_tc<BVG_GAUSSIAN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
  CUDA_TRY(cudaFuncSetAttribute(k_lik_tc<BVG_NB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
  CUDA_TRY(cudaFuncSetAttribute(k_lik_tc<BVG_NB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
  return BVG_OK;
}

void tc_destroy(bvg_fit *f) {
  TcState *t = (TcState *)f->tc;
  if (!t) return;
  cudaFree(t->F1);
  delete t;
  f->tc = nullptr;
}

static long long *g_tc_trace = nullptr;
int launch_lik_tc(bvg_fit *f, bool bwd);
// diagnostics: one backward launch with per-chunk clock64 stamps of CTA (0,0): [chunk][8] =
// mma: begin, after mma1 issue, after mma2 issue | epilogue warp 2 lane 0: Y ready, D1 ready, after tcgen05.ld,
// after math, after R published
extern "C" int bvg_debug_tc_trace(bvg_fit *f, long long *out, int nchunk_cap) {
  if (!f || !f->tc || !out) { bvg_set_error("tcgen05 engine not active"); return BVG_ERR_STATE; }
  CUDA_TRY(cudaSetDevice(f->cfg.device));
  long long *d = nullptr;
  CUDA_TRY(cudaMalloc(&d, sizeof(long long) * 8 * nchunk_cap));
  CUDA_TRY(cudaMemsetAsync(d, 0, sizeof(long long) * 8 * nchunk_cap, f->stream));
  g_tc_trace = d;
  int rc = launch_lik_tc(f, true);
  g_tc_trace = nullptr;
  if (rc == BVG_OK && (cudaMemcpyAsync(out, d, sizeof(long long) * 8 * nchunk_cap, cudaMemcpyDeviceToHost, f->stream) != cudaSuccess || cudaStreamSynchronize(f->stream) != cudaSuccess)) rc = BVG_ERR_CUDA;
  cudaFree(d);
  return rc;
}

int launch_lik_tc(bvg_fit *f, bool bwd) {
  const DevModel &m = f->m;
  TcState *t = (TcState *)f->tc;
  TcArgs a;
  memset(&a, 0, sizeof(a));  // padding bytes take part in the change check below
  a.Ppart = (float *)m.Ppart; a.Spart = (double *)m.Spart;
  a.W = m.W; a.U = m.U; a.k = m.k;
  a.hist_val = m.hist_val; a.hist_cnt = m.hist_cnt; a.nhist = m.nhist;
  a.hterm = m.lik == BVG_NB ? m.hterm : nullptr;  // allocated only when every distinct count gets a lane (api.cu)
  a.nk1 = (m.k + 1 + 15) / 16;
  a.nch32 = (int)(f->ldY / 32);
  a.M = m.M; a.G = m.G; a.Gpad = m.Gpad; a.KP = m.KP;
  a.nchunks = (int)((m.M + TC_CHUNK - 1) / TC_CHUNK);
  a.cps = (a.nchunks + m.nsplit - 1) / m.nsplit;
  a.zeta = m.zeta; a.tail_idx = m.L.tail;
  a.trace = g_tc_trace;
  dim3 grid((unsigned)(m.Gpad / BVG_GENE_TILE), (unsigned)m.nsplit);
#define TCL(LIK, BWD) CUDA_TRY(launch_pdl(k_lik_tc<LIK, BWD>, grid, dim3(TC_THREADS), (size_t)TC_SMEM_BYTES, f->stream, t->tmY, t->tmF1, t->tmF2, t->tmT1, t->tmT2, a))
  if (m.lik == BVG_GAUSSIAN) { if (bwd) TCL(BVG_GAUSSIAN, true); else TCL(BVG_GAUSSIAN, false); }
  else { if (bwd) TCL(BVG_NB, true); else TCL(BVG_NB, false); }
#undef TCL
  f->st.kernel_launches++;
  return BVG_OK;
}
