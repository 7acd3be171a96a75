// tc_roles.cuh -- the warp roles of the fused tcgen05 likelihood pass (one 128-gene tile x one spot range), shared by
//   k_lik_tc        (k_lik_tc.cu): one pass per launch, and
//   k_advi_persist  (k_advi_persist.cu): the whole ADVI stretch in one cooperative launch, one pass per iteration.
// Every role takes a RUNNING chunk counter q0 (chunks this CTA has pushed through its pipelines since the barriers
// were initialised): stage indices and mbarrier phase parities derive from q = q0 + c, so the rings, the two D1 / R
// TMEM buffers and their barriers simply keep turning from one pass to the next.  q0 = 0 gives the one-shot kernel.
//
// Per chunk of 64 spots (inst/approxGP.stan:33 and :48 of the reference are the work being done):
//   MMA1  D1[128 genes x 64 spots] = C'[128 x 32] . Phi'_chunk[64 spots x 32]^T
//   epilogue (8 warps, thread = gene row x 32 spots): f <- TMEM, y <- smem (TMA-written, 128B-swizzled fp32 Y tile),
//         residual + per-thread partial sums, r -> TMEM as two packed bf16 parts
//   MMA2  D2[128 genes x 32] += R[128 x 64 spots] . Phi'^T_chunk[32 x 64 spots]^T
// 384 threads: warp 0 = Y producer, warp 1 = TMEM owner + forward-MMA issuer, warp 2 = backward-MMA issuer, warps 3..10 =
// epilogue (two warps per TMEM lane quarter, one 32-spot half of the chunk each), warp 11 = basis producer.
// Measured alternatives (scripts/mma_probe.cu, profiles/r2_mma_probe.json, profiles/README.md): sixteen epilogue warps of 16
// spots at 96 registers (640 threads) did NOT shorten the chunk period (1,350-1,450 cycles against 1,165 with eight warps):
// all epilogue warps move through wait -> tcgen05.ld -> math -> tcgen05.st -> publish in lockstep (they share one set of
// barriers), the math phase is issue-bound either way (~130 instructions per 16 spots and warp), and the latencies of the
// other phases are exposed, not overlapped.  Neither the tensor pipe (18 MMAs of a chunk retire in ~400 cycles back to back,
// from one issuer or two, one accumulator or several) nor tcgen05.ld (latency ~155 cycles, >= 390 B/cycle/SM with 16 warps)
// nor L2 bandwidth (19 % of peak) bounds the chunk.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include <type_traits>

#include "bvg_internal.cuh"
#include "tc_ptx.cuh"

#define TC_THREADS 384
#define TC_WARP_Y 0
#define TC_WARP_MMA1 1
#define TC_WARP_MMA2 2
#define TC_WARP_EPI0 3    // first epilogue warp (3..10)
#define TC_WARP_BASIS 11
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
// TMEM column map (512 allocated).  One accumulator per product: dependent tcgen05.mma's into one accumulator pipeline at
// full rate (scripts/mma_probe.cu: 17 / 33 cycles per N = 32 / 64 MMA with one accumulator or four), so splitting the three
// BF16x3 terms over separate accumulators only added tcgen05.ld traffic.
#define TM_C1 0      // 16 columns: bf16 part 1 of the coefficient rows (32 values, packed in pairs)
#define TM_C2 16
#define TM_D1 64     // + 64 * buf : forward accumulator, 64 spots
#define TM_R 192     // + 64 * buf : residual part 1 at +0 (32 columns = 64 spots packed), part 2 at +32
#define TM_D2 320    // 32 columns: backward accumulator
#define TM_COLS 512
#define TC_NEPI 256u  // epilogue threads (8 warps)

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
  void *persist;   // PersistBuf of the persistent ADVI kernel (k_advi_persist.cu), allocated on first use
};

// shared-memory / TMEM addresses of one CTA's pipelines
struct TcCtx {
  uint32_t base, sB0;
  uint32_t bFullY, bEmptyY, bFullB, bEmptyB, bD1Full, bD1Empty, bRFull, bRFree, bCFull, bD2Full, sTmemPtr;
  uint32_t tmem;
  long long timeout;   // SM cycles any wait may last before the kernel traps (never hang the GPU on a protocol bug)
  __device__ __forceinline__ uint32_t stY(int s, int h) const { return base + s * YSTAGE_BYTES + h * YT_BYTES; }
  __device__ __forceinline__ uint32_t stF(int s, int w) const { return sB0 + s * BSTAGE_BYTES + w * FT_BYTES; }
  __device__ __forceinline__ uint32_t stT(int s, int w) const { return sB0 + s * BSTAGE_BYTES + 2 * FT_BYTES + w * TT_BYTES; }
};

__device__ __forceinline__ void mbar_wait_to(uint32_t bar, uint32_t parity, long long timeout) {
  uint32_t done = 0, spins = 0;
  long long t0 = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    // the watchdog reads the clock only every 64th failed probe: a waiting warp spins on three instructions, not ten,
    // and leaves the issue slots of its scheduler to the epilogue warps
    if ((++spins & 63u) == 0u) {
      const long long t = clock64();
      if (t0 == 0) t0 = t;
      else if (t - t0 > timeout) __trap();
    }
  }
}

__device__ __forceinline__ TcCtx tc_ctx_make(uint8_t *smem_raw, long long timeout) {
  TcCtx x;
  x.base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024 B alignment
  x.sB0 = x.base + TC_NY * YSTAGE_BYTES;
  const uint32_t sBar = x.sB0 + TC_NB * BSTAGE_BYTES;
  // barrier map (8 B each)
  x.bFullY = sBar; x.bEmptyY = x.bFullY + 8 * TC_NY; x.bFullB = x.bEmptyY + 8 * TC_NY; x.bEmptyB = x.bFullB + 8 * TC_NB;
  x.bD1Full = x.bEmptyB + 8 * TC_NB; x.bD1Empty = x.bD1Full + 16; x.bRFull = x.bD1Empty + 16; x.bRFree = x.bRFull + 16;
  x.bCFull = x.bRFree + 16; x.bD2Full = x.bCFull + 8; x.sTmemPtr = x.bD2Full + 8;
  x.tmem = 0;
  x.timeout = timeout;
  return x;
}

// barrier init (thread 0), TMEM allocation (warp 1), CTA-wide sync; leaves the TMEM base address in x.tmem
__device__ __forceinline__ void tc_setup_cta(TcCtx &x) {
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_NY; ++s) { mbar_init(x.bFullY + 8 * s, 1); mbar_init(x.bEmptyY + 8 * s, TC_NEPI); }
    for (int s = 0; s < TC_NB; ++s) { mbar_init(x.bFullB + 8 * s, 1); mbar_init(x.bEmptyB + 8 * s, 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(x.bD1Full + 8 * b, 1);
      mbar_init(x.bD1Empty + 8 * b, TC_NEPI);
      mbar_init(x.bRFull + 8 * b, TC_NEPI);
      mbar_init(x.bRFree + 8 * b, 1);
    }
    mbar_init(x.bCFull, TC_NEPI);
    mbar_init(x.bD2Full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(x.sTmemPtr), "r"((uint32_t)TM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x.tmem) : "r"(x.sTmemPtr));
}
__device__ __forceinline__ void tc_teardown_cta(const TcCtx &x) {
  tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(x.tmem), "r"((uint32_t)TM_COLS) : "memory");
  }
}

// ===================== TMA producer, Y ring (ONE lane).  Y is written by no kernel: no dependency on the draw =====
// chunks [c0, c1) of this pass (the persistent kernel issues the first chunks of the NEXT pass ahead of time)
__device__ __forceinline__ void tc_role_produce_y(const TcCtx &x, const CUtensorMap *tmY, int tile, int nch32, int c_begin, int c0, int c1, uint32_t q0) {
  for (int c = c0; c < c1; ++c) {
    const uint32_t q = q0 + c;
    const int s = (int)(q % TC_NY);
    if (q >= TC_NY) mbar_wait_to(x.bEmptyY + 8 * s, ((q / TC_NY) - 1) & 1, x.timeout);
    mbar_expect_tx(x.bFullY + 8 * s, YSTAGE_BYTES);
    // one 32 KB contiguous run of the tile-major Y: rows [0,128) = spots spot0..+31, rows [128,256) = +32..+63
    tma_load_2d(x.stY(s, 0), tmY, x.bFullY + 8 * s, 0, (tile * nch32 + 2 * (c_begin + c)) * BVG_GENE_TILE);
  }
}
// ===================== TMA producer, basis ring (ONE lane): chunks [c0, c1) of this pass =====
template <bool BWD>
__device__ __forceinline__ void tc_role_produce_basis(const TcCtx &x, const CUtensorMap *tmF1, const CUtensorMap *tmF2, const CUtensorMap *tmT1,
                                                      const CUtensorMap *tmT2, int c_begin, int c0, int c1, uint32_t q0) {
  for (int c = c0; c < c1; ++c) {
    const uint32_t q = q0 + c;
    const int s = (int)(q % TC_NB);
    if (q >= TC_NB) mbar_wait_to(x.bEmptyB + 8 * s, ((q / TC_NB) - 1) & 1, x.timeout);
    const int spot0 = (c_begin + c) * TC_CHUNK;
    mbar_expect_tx(x.bFullB + 8 * s, 2 * FT_BYTES + (BWD ? 2 * TT_BYTES : 0));
    tma_load_2d(x.stF(s, 0), tmF1, x.bFullB + 8 * s, 0, spot0);
    tma_load_2d(x.stF(s, 1), tmF2, x.bFullB + 8 * s, 0, spot0);
    if (BWD) {
      tma_load_2d(x.stT(s, 0), tmT1, x.bFullB + 8 * s, spot0, 0);
      tma_load_2d(x.stT(s, 1), tmT2, x.bFullB + 8 * s, spot0, 0);
    }
  }
}
// the forward MMAs of one accumulator (BF16x3: C1 F1 + C2 F1 + C1 F2), NK1 = ceil((k + 1) / 16) K steps each, fully unrolled:
// with a run-time trip count ptxas built a 16-way unrolled loop with remainder branches around six instructions, and ISSUING
// them took ~650 cycles per chunk (in-kernel trace) -- longer than the 200 cycles the tensor pipe needs to execute them
template <int NK1>
__device__ __forceinline__ void tc_issue_mma1(uint32_t d1, uint32_t tmem, uint64_t f1, uint64_t f2, uint32_t idesc) {
#pragma unroll
  for (int kk = 0; kk < NK1; ++kk) tc_mma_bf16_ts(d1, tmem + TM_C1 + 8 * kk, f1 + 2 * kk, idesc, kk ? 1u : 0u);
#pragma unroll
  for (int kk = 0; kk < NK1; ++kk) tc_mma_bf16_ts(d1, tmem + TM_C2 + 8 * kk, f1 + 2 * kk, idesc, 1u);
#pragma unroll
  for (int kk = 0; kk < NK1; ++kk) tc_mma_bf16_ts(d1, tmem + TM_C1 + 8 * kk, f2 + 2 * kk, idesc, 1u);
}
// ===================== forward MMA issuer (whole warp convergent, one elected lane issues) =====
// pass = index of this pass since barrier init (parity of the "coefficient rows are in TMEM" barrier)
template <bool BWD>
__device__ __forceinline__ void tc_role_mma1(const TcCtx &x, int n, int nk1, uint32_t q0, uint32_t pass, long long *trace) {
  constexpr uint32_t idesc = make_idesc_bf16(128, 64);
  const int lane = threadIdx.x & 31;
  const bool tr = trace && lane == 0;
  if (n > 0) mbar_wait_to(x.bCFull, pass & 1, x.timeout);
  for (int c = 0; c < n; ++c) {
    const uint32_t q = q0 + c;
    const int s = (int)(q % TC_NB), b = (int)(q & 1);
    if (tr) trace[c * 8 + 0] = clock64();
    mbar_wait_to(x.bFullB + 8 * s, (q / TC_NB) & 1, x.timeout);
    if (q >= 2) mbar_wait_to(x.bD1Empty + 8 * b, ((q >> 1) - 1) & 1, x.timeout);
    tc_fence_after();
    if (elect_one()) {
      const uint32_t d1 = x.tmem + TM_D1 + 64 * b;
      const uint64_t f1 = make_desc(x.stF(s, 0)), f2 = make_desc(x.stF(s, 1));
      if (nk1 == 2) tc_issue_mma1<2>(d1, x.tmem, f1, f2, idesc);
      else tc_issue_mma1<1>(d1, x.tmem, f1, f2, idesc);
      tc_commit(x.bD1Full + 8 * b);
      if (!BWD) tc_commit(x.bEmptyB + 8 * s);  // forward-only: the basis stage is free once MMA1 has read it
    }
    __syncwarp();
    if (tr) trace[c * 8 + 1] = clock64();
  }
}
// ===================== backward MMA issuer =====================
__device__ __forceinline__ void tc_role_mma2(const TcCtx &x, int n, uint32_t q0, long long *trace) {
  constexpr uint32_t idesc = make_idesc_bf16(128, 32);
  const uint32_t d2 = x.tmem + TM_D2;
  const int lane = threadIdx.x & 31;
  const bool tr = trace && lane == 0;
  for (int c = 0; c < n; ++c) {
    const uint32_t q = q0 + c;
    const int s = (int)(q % TC_NB), b = (int)(q & 1);
    mbar_wait_to(x.bFullB + 8 * s, (q / TC_NB) & 1, x.timeout);      // basis tiles of this stage (acquire)
    mbar_wait_to(x.bRFull + 8 * b, (q >> 1) & 1, x.timeout);         // residuals published by the epilogue
    tc_fence_after();
    if (elect_one()) {
      const uint32_t r1 = x.tmem + TM_R + 64 * b, r2 = r1 + 32;
      const uint64_t t1 = make_desc(x.stT(s, 0)), t2 = make_desc(x.stT(s, 1));
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) tc_mma_bf16_ts(d2, r1 + 8 * kk, t1 + 2 * kk, idesc, (c | kk) ? 1u : 0u);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) tc_mma_bf16_ts(d2, r2 + 8 * kk, t1 + 2 * kk, idesc, 1u);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) tc_mma_bf16_ts(d2, r1 + 8 * kk, t2 + 2 * kk, idesc, 1u);
      tc_commit(x.bEmptyB + 8 * s);  // basis stage free for its producer
      tc_commit(x.bRFree + 8 * b);   // residual buffer free for the epilogue of chunk c + 2
      if (c == n - 1) tc_commit(x.bD2Full);
    }
    __syncwarp();
    if (tr) trace[c * 8 + 2] = clock64();
  }
}

// ===================== epilogue threads: coefficient rows -> TMEM (A operand of MMA1), once per pass =====
// cv[i] = column 16 h + i of this thread's gene row
__device__ __forceinline__ void tc_epi_store_coeff(const TcCtx &x, uint32_t trow, int h, const float *cv) {
  uint32_t c1[8], c2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) split_bf16x2(cv[2 * i], cv[2 * i + 1], c1[i], c2[i]);
  tmem_st8u(trow + TM_C1 + 8 * h, c1);
  tmem_st8u(trow + TM_C2 + 8 * h, c2);
  tmem_st_wait();
  tc_fence_before();
  mbar_arrive(x.bCFull);
}
// w[i] = column 16 h + i of the gene's draw row (z_alpha_t[,g], z_beta0[g] | gene_depths[g], ...), Ag = amplitude_sq[g];
// um / us = the matching 16 entries of mu_alpha / sigma_alpha; mb, sb = mu_beta0, sigma_beta0 (beta0, beta1)
__device__ __forceinline__ void tc_epi_form_coeff(int k, int h, bool live, float Ag, float mb, float sb, const float *w, const float *um,
                                                  const float *us, float *cv) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int j = 16 * h + i;
    float v = 0.f;
    if (j < k) v = Ag * fmaf(us[i], w[i], um[i]);      // A_g * alpha_t[j,g]  (approxGP.stan:27, :32)
    else if (j == k) v = fmaf(sb, w[i], mb);           // beta0[g]            (approxGP.stan:26 | approxGP2.stan:46)
    cv[i] = live ? v : 0.f;
  }
}

// ===================== epilogue main loop (thread = gene row r x 32-spot half h of the chunk) =====
// per-chunk sums (32 terms) in fp32, running sums over the CTA's spot range in fp64: a serial fp32 sum over
// thousands of spots loses ~1e-5 of the log density, enough to stall the L-BFGS line search at config-4 scale
template <int LIK, bool BWD>
__device__ __forceinline__ void tc_role_epilogue(const TcCtx &x, uint32_t trow, int r, int h, int c_begin, int n, uint32_t q0, int64_t M, float lt,
                                                 float phi, double &sum0, double &sum1, long long *trace) {
  const bool tr = trace != nullptr;
  for (int c = 0; c < n; ++c) {
    const uint32_t q = q0 + c;
    const int s = (int)(q % TC_NY), b = (int)(q & 1);
    mbar_wait_to(x.bFullY + 8 * s, (q / TC_NY) & 1, x.timeout);      // Y tile landed (acquire for this thread)
    if (tr) trace[c * 8 + 3] = clock64();
    mbar_wait_to(x.bD1Full + 8 * b, (q >> 1) & 1, x.timeout);        // MMA1 of this chunk retired
    tc_fence_after();
    if (tr) trace[c * 8 + 4] = clock64();
    float f[32];
    tmem_ld32(trow + TM_D1 + 64 * b + 32 * h, f);
    if (tr) trace[c * 8 + 5] = clock64();
    tc_fence_before();
    mbar_arrive(x.bD1Empty + 8 * b);
    const uint32_t yrow = x.stY(s, h) + r * 128;
    const int64_t spot0 = (int64_t)(c_begin + c) * TC_CHUNK + 32 * h;
    uint32_t p1[16], p2[16];
    float acc0 = 0.f, acc1 = 0.f;
    uint64_t accp = 0ull, accq = 0ull;   // packed pairs of running sums (gaussian: squares; nb: log-lik terms, log1p_exp terms)
    // masked = std::true_type only for the one chunk that straddles M (nb: padded spots have a non-zero term)
    auto body = [&](auto masked, int nvalid) {
      constexpr bool MASKED = decltype(masked)::value;
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4) {
        const uint32_t off = (uint32_t)((c4 ^ (r & 7)) << 4);  // 128B swizzle: 16B chunk index ^ (row & 7)
        const float4 y4 = lds128(yrow + off);
        const float yv[4] = {y4.x, y4.y, y4.z, y4.w};
        float res[4];
        if (LIK == BVG_GAUSSIAN) {
          // packed fp32 pairs: residual, squared-error accumulation and the bf16 remainder are one instruction per TWO
          // elements each (the epilogue is bound by issue slots: ~5.25 instructions per element before, ~3.75 now)
          const uint64_t ra = f2_sub(f2_pack(yv[0], yv[1]), f2_pack(f[4 * c4], f[4 * c4 + 1]));
          const uint64_t rb = f2_sub(f2_pack(yv[2], yv[3]), f2_pack(f[4 * c4 + 2], f[4 * c4 + 3]));
          accp = f2_fma(ra, ra, accp);
          accp = f2_fma(rb, rb, accp);
          split_bf16x2_pk(ra, p1[2 * c4], p2[2 * c4]);
          split_bf16x2_pk(rb, p1[2 * c4 + 1], p2[2 * c4 + 1]);
          continue;
        }
        if (LIK == BVG_NB && !MASKED) {
          // negative binomial, packed: per element TWO MUFU operations (ex2, rcp) instead of three -- log1p(e^{-|d|}) comes
          // from a degree-8 polynomial in u = e^{-|d|} in (0, 1] on the FMA pipe (fp32 Horner error 1.2e-7 absolute, the same
          // as lg2.approx x ln 2 had) -- and everything that is an add / multiply / fma is one instruction per two elements.
          // The round-1 form (3 MUFU + ~17 scalar instructions per element) was bound by the MUFU pipe: 1,536 cycles per
          // chunk and scheduler against 950 for the gaussian epilogue.
          //   ll_sg = y d - (y + phi) L,  d = eta - log phi,  L = log1p_exp(d)   (approxGP4.stan:48)
          const uint64_t one2 = f2_pack(1.f, 1.f), mone2 = f2_pack(-1.f, -1.f), lt2 = f2_pack(lt, lt), nphi2 = f2_pack(-phi, -phi);
#pragma unroll
          for (int hp = 0; hp < 2; ++hp) {
            const uint64_t y2 = f2_pack(yv[2 * hp], yv[2 * hp + 1]);
            const uint64_t d2 = f2_sub(f2_pack(f[4 * c4 + 2 * hp], f[4 * c4 + 2 * hp + 1]), lt2);
            float d0, d1;
            f2_unpack(d2, d0, d1);
            const float e0 = ex2_approx(-fabsf(d0) * 1.4426950408889634f), e1 = ex2_approx(-fabsf(d1) * 1.4426950408889634f);
            const uint64_t u2 = f2_pack(e0, e1);
            float t0, t1;
            f2_unpack(f2_fma(u2, one2, one2), t0, t1);
            const uint64_t rt2 = f2_pack(rcp_approx(t0), rcp_approx(t1));
            uint64_t L2 = f2_pack(-0.006006604991853237f, -0.006006604991853237f);
            L2 = f2_fma(L2, u2, f2_pack(0.03426460176706314f, 0.03426460176706314f));
            L2 = f2_fma(L2, u2, f2_pack(-0.09229041635990143f, -0.09229041635990143f));
            L2 = f2_fma(L2, u2, f2_pack(0.1649981290102005f, 0.1649981290102005f));
            L2 = f2_fma(L2, u2, f2_pack(-0.2394333779811859f, -0.2394333779811859f));
            L2 = f2_fma(L2, u2, f2_pack(0.33144664764404297f, 0.33144664764404297f));
            L2 = f2_fma(L2, u2, f2_pack(-0.49982550740242004f, -0.49982550740242004f));
            L2 = f2_fma(L2, u2, f2_pack(0.999993622303009f, 0.999993622303009f));
            L2 = f2_fma(L2, u2, f2_pack(3.910905377324525e-08f, 3.910905377324525e-08f));   // log1p(u)
            const uint64_t Lq2 = f2_fma(L2, one2, f2_pack(fmaxf(d0, 0.f), fmaxf(d1, 0.f)));      // + max(d, 0) = log1p_exp(d)
            const uint64_t sg2 = f2_fma(f2_pack(d0 >= 0.f ? 1.f : e0, d1 >= 0.f ? 1.f : e1), rt2, 0ull);   // sigmoid(d)
            const uint64_t nyp2 = f2_fma(y2, mone2, nphi2);                                     // -(y + phi)
            const uint64_t r2 = f2_fma(nyp2, sg2, y2);                                          // y - (y + phi) sigmoid(d)
            accp = f2_fma(y2, d2, accp);
            accp = f2_fma(nyp2, Lq2, accp);
            accq = f2_fma(Lq2, one2, accq);
            split_bf16x2_pk(r2, p1[2 * c4 + hp], p2[2 * c4 + hp]);
          }
          continue;
        }
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
    const int64_t left = M - spot0;
    if (LIK == BVG_GAUSSIAN || left >= 32) body(std::false_type{}, 32);
    else body(std::true_type{}, left > 0 ? (int)left : 0);
    mbar_arrive(x.bEmptyY + 8 * s);  // y is in registers: the Y stage can be refilled
    if (LIK == BVG_GAUSSIAN) { float a0, a1; f2_unpack(accp, a0, a1); acc0 = a0 + a1; }
    else { float a0, a1, b0, b1; f2_unpack(accp, a0, a1); f2_unpack(accq, b0, b1); acc0 += a0 + a1; acc1 += b0 + b1; }
    sum0 += (double)acc0;
    if (LIK == BVG_NB) sum1 += (double)acc1;
    if (BWD) {
      if (tr) trace[c * 8 + 6] = clock64();
      if (q >= 2) mbar_wait_to(x.bRFree + 8 * b, ((q >> 1) - 1) & 1, x.timeout);  // MMA2 of chunk q-2 has consumed this buffer
      tc_fence_after();
      tmem_st16u(trow + TM_R + 64 * b + 16 * h, p1);
      tmem_st16u(trow + TM_R + 64 * b + 32 + 16 * h, p2);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(x.bRFull + 8 * b);
      if (tr) trace[c * 8 + 7] = clock64();
    }
  }
}
// P' = D2 (this thread's 16 columns 16 h .. 16 h + 15 of its gene row), once per pass
__device__ __forceinline__ void tc_epi_read_d2(const TcCtx &x, uint32_t trow, int h, int n, uint32_t pass, float *p) {
  if (n > 0) {
    mbar_wait_to(x.bD2Full, pass & 1, x.timeout);
    tc_fence_after();
    tmem_ld16(trow + TM_D2 + 16 * h, p);
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) p[i] = 0.f;
  }
}
// the thread's 16 values of P' as float4 stores (16 h + i < KP, KP a multiple of 8): 32 lanes x 16 B spread over 3 KB
// cost the LSU ~24 wavefronts per instruction; the scalar form (16 stores of 4 B at a 96 B lane stride) cost 32 each --
// 4,096 wavefront-cycles per CTA and pass, 3,500 cycles of the drain in the in-kernel trace
__device__ __forceinline__ void tc_epi_store_p(float *dst_row, int h, int KP, const float *p) {
#pragma unroll
  for (int q = 0; q < 4; ++q)
    if (16 * h + 4 * q < KP) reinterpret_cast<float4 *>(dst_row + 16 * h)[q] = make_float4(p[4 * q], p[4 * q + 1], p[4 * q + 2], p[4 * q + 3]);
}
