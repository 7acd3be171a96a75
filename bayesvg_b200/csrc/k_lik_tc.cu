// k_lik_tc.cu -- the fused likelihood kernel for sm_100a: TMA-fed tcgen05 (TF32, 3-term split)
// contractions with TMEM accumulators AND TMEM-resident A operands, one pass over Y per evaluation.
//
// Per CTA: one tile of 128 genes (= the 128 TMEM lanes) x one contiguous range of spots, walked
// in chunks of 32 spots (one 128-byte swizzle row of fp32).  For every chunk
//   MMA1  D1[128 genes x 32 spots]  = C'[128 x 32] . Phi'_chunk[32 spots x 32]^T      (forward:  f = beta0 + A * Phi alpha;
//                                                                                       inst/approxGP.stan:33 and :48 of the reference)
//   epilogue (8 warps, thread = gene row x 16 spots): f <- TMEM, y <- smem (the TMA-written Y tile),
//         r = y - f (gaussian) | y - (y+phi) sigmoid(f - log phi) (nb), SSE / log-lik partials in
//         registers, r -> TMEM (tf32 hi and lo parts) with tcgen05.st
//   MMA2  D2[128 genes x 32]       += R[128 x 32 spots] . Phi'^T_chunk[32 x 32 spots]^T (backward: Phi^T r and R_g = 1^T r)
// Both 128-row operands (the coefficient rows C' and the residuals R) live in TENSOR MEMORY (A-from-
// TMEM form of tcgen05.mma): shared memory only carries what TMA brings in (Y tile + four 4 KB basis
// tiles per chunk) and is read once by the epilogue (Y) and once per term by the tensor core (basis).
// With both A operands in shared memory the kernel was shared-memory-bandwidth bound (~200 KB of
// smem traffic per 16 KB of Y; profiles/README.md, "v1").  D2 stays in TMEM for the whole spot range.
// "3xTF32": every product is evaluated as hi*hi + lo*hi + hi*lo with hi = rna_tf32(x), lo = x - hi, so
// the tensor-core result carries ~2^-21 relative error instead of TF32's 2^-11.
//
// Warp roles (352 threads, 1 CTA/SM): warp 0 = TMA producer, warp 1 = TMEM owner + forward-MMA issuer,
// warp 2 = backward-MMA issuer (two issue streams: the MMAs are small, so instruction issue, not the
// tensor pipe, is the scarce resource), warps 3..10 = epilogue (two warps per TMEM lane quarter, 16
// columns each).  MMA warps stay convergent and issue from one elected lane (elect.sync).
// Pipelines: smem stage full/empty (6 stages), D1 full/empty (2 TMEM buffers), R ready/free (2).
#include <cuda.h>

#include "bvg_internal.cuh"

#define TC_THREADS 352
#define TC_STAGES 6
#define TC_CHUNK 32   // spots per chunk
#define TC_KW 32      // padded operand width (k+1 -> 32 floats = one 128B swizzle row)
#define TILE_A_BYTES (BVG_GENE_TILE * 128)  // 16 KB: 128 rows x 128 B
#define TILE_B_BYTES (32 * 128)             // 4 KB: 32 rows x 128 B
#define STAGE_BYTES (TILE_A_BYTES + 4 * TILE_B_BYTES)
#define TC_SMEM_BYTES (TC_STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/)
// TMEM column map (512 allocated)
#define TM_CHI 0
#define TM_CLO 32
#define TM_D1 64     // + 32 * buf
#define TM_R 128     // + 64 * buf : hi at +0, lo at +32
#define TM_D2 256
#define TM_COLS 512

struct TcArgs {
  float *Ppart, *Spart;
  const float *CH, *CL;  // [Gpad][32] tf32 hi / lo coefficient rows written by k_prep
  int64_t M, G, Gpad;
  int KP, cps, nchunks;
  const double *zeta;
  int64_t tail_idx;
  long long *trace;  // optional [chunk][8] clock64 stamps of CTA (0,0) (diagnostics)
};

struct TcState {
  CUtensorMap tmY, tmPhiH, tmPhiL, tmPhiTH, tmPhiTL;
  float *PhiH, *PhiL, *PhiTH, *PhiTL;
};

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  const long long t0 = clock64();
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if (clock64() - t0 > 4000000000LL) __trap();  // watchdog: never hang the GPU on a protocol bug
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]^T, kind::tf32, M = 128, N = 32, K = 8 (A: 128 lanes x 8 columns)
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start >> 4,
// LBO = 1 (unused for swizzled K-major), SBO = 1024 B between 8-row groups, version 1, layout 2.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, K-major both, N>>3, M>>4
__device__ __forceinline__ constexpr uint32_t make_idesc(int Mdim, int Ndim) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(Ndim >> 3) << 17) | ((uint32_t)(Mdim >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float *v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
        "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
        "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])),
        "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
        "r"(__float_as_uint(v[15]))
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(p));
  return p != 0;
}
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}

// ------------------------------------------------------------------------------------------------
template <int LIK, bool BWD>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_lik_tc(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmPhiH,
         const __grid_constant__ CUtensorMap tmPhiL, const __grid_constant__ CUtensorMap tmPhiTH,
         const __grid_constant__ CUtensorMap tmPhiTL, TcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const bool tr0 = a.trace && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0;
  if (tr0) { a.trace[8 * 30 + 0] = clock64(); asm volatile("mov.u64 %0, %globaltimer;" : "=l"(a.trace[8 * 30 + 4])); }
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024 B alignment
  const uint32_t sBar = base + TC_STAGES * STAGE_BYTES;
  // barrier map (8 B each)
  const uint32_t bFull = sBar, bEmpty = bFull + 8 * TC_STAGES;
  const uint32_t bD1Full = bEmpty + 8 * TC_STAGES, bD1Empty = bD1Full + 16, bRFull = bD1Empty + 16, bRFree = bRFull + 16;
  const uint32_t bCFull = bRFree + 16, bD2Full = bCFull + 8, sTmemPtr = bD2Full + 8;
  __shared__ float s_half[2][BVG_GENE_TILE];
  auto stY = [&](int s) { return base + s * STAGE_BYTES; };
  auto stPh = [&](int s, int w) { return base + s * STAGE_BYTES + TILE_A_BYTES + w * TILE_B_BYTES; };  // w: 0 PhiH 1 PhiL 2 PhiTH 3 PhiTL

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, split = blockIdx.y;
  const int c_begin = split * a.cps;
  int n = a.nchunks - c_begin;
  n = n < a.cps ? n : a.cps;
  if (n < 0) n = 0;
  constexpr uint32_t NEPI = 256;

  griddep_launch();  // PDL: the successor (k_gene) may begin launching; it waits for this grid to finish
  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(bFull + 8 * s, 1);
      mbar_init(bEmpty + 8 * s, BWD ? 1 : NEPI);
    }
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
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int c = 0; c < n; ++c) {
        const int s = c % TC_STAGES;
        if (c >= TC_STAGES) mbar_wait(bEmpty + 8 * s, ((c / TC_STAGES) - 1) & 1);
        const int spot0 = (c_begin + c) * TC_CHUNK;
        mbar_expect_tx(bFull + 8 * s, TILE_A_BYTES + (BWD ? 4 : 2) * TILE_B_BYTES);
        tma_load_2d(stY(s), &tmY, bFull + 8 * s, spot0, tile * BVG_GENE_TILE);
        tma_load_2d(stPh(s, 0), &tmPhiH, bFull + 8 * s, 0, spot0);
        tma_load_2d(stPh(s, 1), &tmPhiL, bFull + 8 * s, 0, spot0);
        if (BWD) {
          tma_load_2d(stPh(s, 2), &tmPhiTH, bFull + 8 * s, spot0, 0);
          tma_load_2d(stPh(s, 3), &tmPhiTL, bFull + 8 * s, spot0, 0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== forward MMA issuer (whole warp convergent, one elected lane issues) =====
    constexpr uint32_t idesc = make_idesc(128, 32);
    const bool tr = a.trace && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0;
    if (n > 0) mbar_wait(bCFull, 0);
    for (int c = 0; c < n; ++c) {
      const int s = c % TC_STAGES, b = c & 1;
      if (tr) a.trace[c * 8 + 0] = clock64();
      mbar_wait(bFull + 8 * s, (c / TC_STAGES) & 1);
      if (c >= 2) mbar_wait(bD1Empty + 8 * b, ((c >> 1) - 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t d1 = tmem + TM_D1 + 32 * b;
        const uint64_t bh = make_desc(stPh(s, 0)), bl = make_desc(stPh(s, 1));
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) tc_mma_tf32_ts(d1, tmem + TM_CHI + 8 * kk, bh + 2 * kk, idesc, kk ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) tc_mma_tf32_ts(d1, tmem + TM_CLO + 8 * kk, bh + 2 * kk, idesc, 1u);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) tc_mma_tf32_ts(d1, tmem + TM_CHI + 8 * kk, bl + 2 * kk, idesc, 1u);
        tc_commit(bD1Full + 8 * b);
      }
      __syncwarp();
      if (tr) a.trace[c * 8 + 1] = clock64();
    }
  } else if (warp == 2) {
    // ===================== backward MMA issuer =====================
    if (BWD) {
      constexpr uint32_t idesc = make_idesc(128, 32);
      const uint32_t d2 = tmem + TM_D2;
      const bool tr = a.trace && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0;
      for (int c = 0; c < n; ++c) {
        const int s = c % TC_STAGES, b = c & 1;
        mbar_wait(bFull + 8 * s, (c / TC_STAGES) & 1);   // basis tiles of this stage (acquire)
        mbar_wait(bRFull + 8 * b, (c >> 1) & 1);         // residuals published by the epilogue
        tc_fence_after();
        if (elect_one()) {
          const uint32_t rh = tmem + TM_R + 64 * b, rl = rh + 32;
          const uint64_t th = make_desc(stPh(s, 2)), tl = make_desc(stPh(s, 3));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) tc_mma_tf32_ts(d2, rh + 8 * kk, th + 2 * kk, idesc, (c | kk) ? 1u : 0u);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) tc_mma_tf32_ts(d2, rl + 8 * kk, th + 2 * kk, idesc, 1u);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) tc_mma_tf32_ts(d2, rh + 8 * kk, tl + 2 * kk, idesc, 1u);
          tc_commit(bEmpty + 8 * s);   // smem stage free for the producer
          tc_commit(bRFree + 8 * b);   // residual buffer free for the epilogue of chunk c + 2
          if (c == n - 1) tc_commit(bD2Full);
        }
        __syncwarp();
        if (tr) a.trace[c * 8 + 2] = clock64();
      }
    }
  } else {
    // ===================== epilogue: 8 warps =====================
    const int e = warp - 3, q = warp & 3, h = e >> 2;  // TMEM lane quarter of this warp, column half
    const int r = 32 * q + lane;                       // gene row inside the tile = TMEM lane
    const uint32_t trow = tmem + ((uint32_t)(32 * q) << 16);
    const int64_t g = (int64_t)tile * BVG_GENE_TILE + r;
    {  // coefficient rows -> TMEM (A operand of MMA1), once per CTA
      // PDL: everything above (barriers, TMEM, and the producer's first TMA loads of Y / basis tiles, which
      // no kernel writes) overlaps the tail of k_prep; its output C is first touched here
      griddep_wait();
      float ch[16], cl[16];
      const float4 *ph = reinterpret_cast<const float4 *>(a.CH + g * TC_KW + 16 * h);
      const float4 *pl = reinterpret_cast<const float4 *>(a.CL + g * TC_KW + 16 * h);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 x = ph[i], y = pl[i];
        ch[4 * i] = x.x; ch[4 * i + 1] = x.y; ch[4 * i + 2] = x.z; ch[4 * i + 3] = x.w;
        cl[4 * i] = y.x; cl[4 * i + 1] = y.y; cl[4 * i + 2] = y.z; cl[4 * i + 3] = y.w;
      }
      tmem_st16(trow + TM_CHI + 16 * h, ch);
      tmem_st16(trow + TM_CLO + 16 * h, cl);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bCFull);
    }
    float acc0 = 0.f, acc1 = 0.f, lt = 0.f, phi = 0.f;
    if (LIK == BVG_NB) { lt = (float)a.zeta[a.tail_idx]; phi = __expf(lt); }
    const bool tr = a.trace && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 96;
    for (int c = 0; c < n; ++c) {
      const int s = c % TC_STAGES, b = c & 1;
      mbar_wait(bFull + 8 * s, (c / TC_STAGES) & 1);   // Y tile landed (acquire for this thread)
      if (tr) a.trace[c * 8 + 3] = clock64();
      mbar_wait(bD1Full + 8 * b, (c >> 1) & 1);        // MMA1 of this chunk retired
      tc_fence_after();
      if (tr) a.trace[c * 8 + 4] = clock64();
      float f[16];
      tmem_ld16(trow + TM_D1 + 32 * b + 16 * h, f);
      if (tr) a.trace[c * 8 + 5] = clock64();
      tc_fence_before();
      mbar_arrive(bD1Empty + 8 * b);
      const uint32_t yrow = stY(s) + r * 128;
      const int64_t spot0 = (int64_t)(c_begin + c) * TC_CHUNK + 16 * h;
      float hi[16], lo[16];
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        const uint32_t off = (uint32_t)(((4 * h + c4) ^ (r & 7)) << 4);  // 128B swizzle: 16B chunk index ^ (row & 7)
        const float4 y4 = lds128(yrow + off);
        const float yv[4] = {y4.x, y4.y, y4.z, y4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float ff = f[4 * c4 + i];
          float res;
          if (LIK == BVG_GAUSSIAN) {
            res = yv[i] - ff;
            acc0 = fmaf(res, res, acc0);
          } else {
            const float d = ff - lt;
            const float ex = __expf(-fabsf(d));
            const float Lq = fmaxf(d, 0.f) + log1pf(ex);
            const float sg = (d >= 0.f ? 1.f : ex) / (1.f + ex);
            res = yv[i] - (yv[i] + phi) * sg;
            if (spot0 + 4 * c4 + i < a.M) {
              acc0 += yv[i] * ff - (yv[i] + phi) * Lq - yv[i] * lt;
              acc1 += Lq;
            }
          }
          hi[4 * c4 + i] = tf32_rna(res);
          lo[4 * c4 + i] = res - hi[4 * c4 + i];
        }
      }
      if (BWD) {
        if (tr) a.trace[c * 8 + 6] = clock64();
        if (c >= 2) mbar_wait(bRFree + 8 * b, ((c >> 1) - 1) & 1);  // MMA2 of chunk c-2 has consumed this buffer
        tc_fence_after();
        tmem_st16(trow + TM_R + 64 * b + 16 * h, hi);
        tmem_st16(trow + TM_R + 64 * b + 32 + 16 * h, lo);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(bRFull + 8 * b);
        if (tr) a.trace[c * 8 + 7] = clock64();
      } else {
        mbar_arrive(bEmpty + 8 * s);
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
    if (h == 1) { s_half[0][r] = acc0; s_half[1][r] = acc1; }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (h == 0) {
      a.Spart[row * 2 + 0] = acc0 + s_half[0][r];
      a.Spart[row * 2 + 1] = acc1 + s_half[1][r];
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
// Phi (M x k col-major fp64) -> tf32 hi/lo splits of Phi' in both orientations
__global__ void k_tc_stage_phi(const double *__restrict__ phi, int64_t M, int64_t Mpad, int k, float *__restrict__ PhiH,
                               float *__restrict__ PhiL, float *__restrict__ PhiTH, float *__restrict__ PhiTL) {
  const int64_t n = Mpad * TC_KW;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = i / TC_KW;
    const int j = (int)(i - s * TC_KW);
    float v = 0.f;
    if (s < M) v = j < k ? (float)phi[(int64_t)j * M + s] : (j == k ? 1.f : 0.f);
    uint32_t hb;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v));
    const float hi = __uint_as_float(hb), lo = v - hi;
    PhiH[i] = hi; PhiL[i] = lo;
    PhiTH[(int64_t)j * Mpad + s] = hi; PhiTL[(int64_t)j * Mpad + s] = lo;
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
// 2-D fp32 tensor [rows][cols] (cols contiguous, row pitch ld floats), box = box_rows x 32 floats, SWIZZLE_128B
static int make_map(CUtensorMap *m, void *ptr, uint64_t cols, uint64_t rows, uint64_t ld, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { bvg_set_error("cuTensorMapEncodeTiled is not available from this driver"); return BVG_ERR_CUDA; }
  cuuint64_t dims[2] = {cols, rows}, strides[1] = {ld * sizeof(float)};
  cuuint32_t box[2] = {32, box_rows}, es[2] = {1, 1};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
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
  const size_t nb = sizeof(float) * (size_t)f->Mpad * TC_KW;
  CUDA_TRY(cudaMalloc(&t->PhiH, 4 * nb));
  t->PhiL = t->PhiH + (size_t)f->Mpad * TC_KW; t->PhiTH = t->PhiL + (size_t)f->Mpad * TC_KW; t->PhiTL = t->PhiTH + (size_t)f->Mpad * TC_KW;
  k_tc_stage_phi<<<148 * 2, 256, 0, f->stream>>>(f->phi64, m.M, f->Mpad, m.k, t->PhiH, t->PhiL, t->PhiTH, t->PhiTL);
  CUDA_TRY(cudaGetLastError());
  BVG_TRY(make_map(&t->tmY, f->Y, (uint64_t)f->ldY, (uint64_t)m.G, (uint64_t)f->ldY, BVG_GENE_TILE));
  BVG_TRY(make_map(&t->tmPhiH, t->PhiH, TC_KW, (uint64_t)f->Mpad, TC_KW, 32));
  BVG_TRY(make_map(&t->tmPhiL, t->PhiL, TC_KW, (uint64_t)f->Mpad, TC_KW, 32));
  BVG_TRY(make_map(&t->tmPhiTH, t->PhiTH, (uint64_t)f->Mpad, TC_KW, (uint64_t)f->Mpad, 32));
  BVG_TRY(make_map(&t->tmPhiTL, t->PhiTL, (uint64_t)f->Mpad, TC_KW, (uint64_t)f->Mpad, 32));
  CUDA_TRY(cudaFuncSetAttribute(k_lik_tc<BVG_GAUSSIAN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
  CUDA_TRY(cudaFuncSetAttribute(k_lik_tc<BVG_GAUSSIAN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
  CUDA_TRY(cudaFuncSetAttribute(k_lik_tc<BVG_NB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
  CUDA_TRY(cudaFuncSetAttribute(k_lik_tc<BVG_NB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
  return BVG_OK;
}

void tc_destroy(bvg_fit *f) {
  TcState *t = (TcState *)f->tc;
  if (!t) return;
  cudaFree(t->PhiH);
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
  a.Ppart = (float *)m.Ppart; a.Spart = (float *)m.Spart;
  a.CH = (const float *)m.C; a.CL = a.CH + (size_t)m.Gpad * TC_KW;  // k_prep writes the hi/lo coefficient rows here
  a.M = m.M; a.G = m.G; a.Gpad = m.Gpad; a.KP = m.KP;
  a.nchunks = (int)((m.M + TC_CHUNK - 1) / TC_CHUNK);
  a.cps = (a.nchunks + m.nsplit - 1) / m.nsplit;
  a.zeta = m.zeta; a.tail_idx = m.L.tail;
  a.trace = g_tc_trace;
  dim3 grid((unsigned)(m.Gpad / BVG_GENE_TILE), (unsigned)m.nsplit);
#define TCL(LIK, BWD) CUDA_TRY(launch_pdl(k_lik_tc<LIK, BWD>, grid, dim3(TC_THREADS), (size_t)TC_SMEM_BYTES, f->stream, t->tmY, t->tmPhiH, t->tmPhiL, t->tmPhiTH, t->tmPhiTL, a))
  if (m.lik == BVG_GAUSSIAN) { if (bwd) TCL(BVG_GAUSSIAN, true); else TCL(BVG_GAUSSIAN, false); }
  else { if (bwd) TCL(BVG_NB, true); else TCL(BVG_NB, false); }
#undef TCL
  f->st.kernel_launches++;
  return BVG_OK;
}
