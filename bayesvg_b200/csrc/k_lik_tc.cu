// k_lik_tc.cu -- the fused likelihood kernel for sm_100a: TMA-fed tcgen05 (TF32, 3-term split)
// contractions with TMEM accumulators, one pass over Y per evaluation.
//
// Per CTA: one tile of 128 genes (= the 128 TMEM lanes) x one contiguous range of spots, walked
// in chunks of 32 spots (one 128-byte swizzle row of fp32).  For every chunk
//   MMA1  D1[128 genes x 32 spots]  = C'[128 x 32] . Phi'_chunk[32 spots x 32]^T      (forward:  f = beta0 + A * Phi alpha;
//                                                                                       inst/approxGP.stan:33 and :48 of the reference)
//   epilogue (8 warps, thread = gene row x 16 spots): f <- TMEM, y <- smem (the TMA-written Y tile),
//         r = y - f (gaussian) | y - (y+phi) sigmoid(f - log phi) (nb), SSE / log-lik partials in
//         registers, r written BACK IN PLACE over the Y tile (tf32 hi part) + a second tile (lo part)
//   MMA2  D2[128 genes x 32]       += R[128 x 32 spots] . Phi'^T_chunk[32 x 32 spots]^T (backward: Phi^T r and R_g = 1^T r)
// D2 stays in TMEM for the whole spot range and is read once at the end.  All operands are K-major,
// 128B-swizzled, written by TMA (or, for R, by the epilogue in the same swizzled layout).
// "3xTF32": every product is evaluated as hi*hi + lo*hi + hi*lo with hi = rna_tf32(x), lo = x - hi, so
// the tensor-core result carries ~2^-21 relative error instead of TF32's 2^-11.
//
// Warp roles (320 threads, 1 CTA/SM): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2..9 = epilogue (two warps per TMEM lane quarter, 16 columns each).
// Pipelines: smem stage full/empty (3 stages), D1 full/empty (2 TMEM buffers), residual-ready.
#include <cuda.h>

#include "bvg_internal.cuh"

#define TC_THREADS 320
#define TC_STAGES 3
#define TC_CHUNK 32   // spots per chunk
#define TC_KW 32      // padded operand width (k+1 -> 32 floats = one 128B swizzle row)
#define TILE_A_BYTES (BVG_GENE_TILE * 128)  // 16 KB: 128 rows x 128 B
#define TILE_B_BYTES (32 * 128)             // 4 KB: 32 rows x 128 B
#define STAGE_BYTES (2 * TILE_A_BYTES + 4 * TILE_B_BYTES)
#define TC_SMEM_BYTES (2 * TILE_A_BYTES + TC_STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/)

struct TcArgs {
  float *Ppart, *Spart;
  int64_t M, G, Gpad;
  int KP, cps, nchunks;
  const double *zeta;
  int64_t tail_idx;
};

struct TcState {
  CUtensorMap tmY, tmPhiH, tmPhiL, tmPhiTH, tmPhiTL, tmCH, tmCL;
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
// D[tmem] (+)= A[smem] . B[smem]^T, kind::tf32, M = 128, N = 32, K = 8
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
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
__device__ __forceinline__ void sts128(uint32_t a, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ------------------------------------------------------------------------------------------------
template <int LIK, bool BWD>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_lik_tc(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmPhiH,
         const __grid_constant__ CUtensorMap tmPhiL, const __grid_constant__ CUtensorMap tmPhiTH,
         const __grid_constant__ CUtensorMap tmPhiTL, const __grid_constant__ CUtensorMap tmCH,
         const __grid_constant__ CUtensorMap tmCL, TcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024 B alignment
  const uint32_t sCH = base, sCL = base + TILE_A_BYTES, sStage0 = base + 2 * TILE_A_BYTES;
  const uint32_t sBar = sStage0 + TC_STAGES * STAGE_BYTES;
  // barrier map (8 B each)
  const uint32_t bFull = sBar, bEmpty = sBar + 8 * TC_STAGES, bRes = sBar + 16 * TC_STAGES;
  const uint32_t bD1Full = sBar + 24 * TC_STAGES, bD1Empty = bD1Full + 16, bCFull = bD1Empty + 16, bD2Full = bCFull + 8;
  const uint32_t sTmemPtr = bD2Full + 8;
  __shared__ float s_half[2][BVG_GENE_TILE];
  auto stY = [&](int s) { return sStage0 + s * STAGE_BYTES; };
  auto stL = [&](int s) { return sStage0 + s * STAGE_BYTES + TILE_A_BYTES; };
  auto stPh = [&](int s, int w) { return sStage0 + s * STAGE_BYTES + 2 * TILE_A_BYTES + w * TILE_B_BYTES; };  // w: 0 PhiH 1 PhiL 2 PhiTH 3 PhiTL

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, split = blockIdx.y;
  const int c_begin = split * a.cps;
  int n = a.nchunks - c_begin;
  n = n < a.cps ? n : a.cps;
  if (n < 0) n = 0;
  constexpr uint32_t NEPI = 256;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(bFull + 8 * s, 1);
      mbar_init(bEmpty + 8 * s, BWD ? 1 : NEPI);
      mbar_init(bRes + 8 * s, NEPI);
    }
    for (int b = 0; b < 2; ++b) { mbar_init(bD1Full + 8 * b, 1); mbar_init(bD1Empty + 8 * b, NEPI); }
    mbar_init(bCFull, 1);
    mbar_init(bD2Full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM: 128 columns (2 x 32 for D1, 32 for D2; power of two)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sTmemPtr), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(sTmemPtr));

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0 && n > 0) {
      mbar_expect_tx(bCFull, 2 * TILE_A_BYTES);
      tma_load_2d(sCH, &tmCH, bCFull, 0, tile * BVG_GENE_TILE);
      tma_load_2d(sCL, &tmCL, bCFull, 0, tile * BVG_GENE_TILE);
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
    // ===================== MMA issuer =====================
    if (lane == 0 && n > 0) {
      constexpr uint32_t idesc = make_idesc(128, 32);
      const uint32_t d2 = tmem + 64;
      auto mma1 = [&](int c) {
        const int s = c % TC_STAGES, b = c & 1;
        mbar_wait(bFull + 8 * s, (c / TC_STAGES) & 1);
        if (c >= 2) mbar_wait(bD1Empty + 8 * b, ((c >> 1) - 1) & 1);
        tc_fence_after();
        const uint32_t d1 = tmem + 32 * b;
        const uint32_t A[3] = {sCH, sCL, sCH}, B[3] = {stPh(s, 0), stPh(s, 0), stPh(s, 1)};
#pragma unroll
        for (int t = 0; t < 3; ++t)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            tc_mma_tf32(d1, make_desc(A[t] + 32 * kk), make_desc(B[t] + 32 * kk), idesc, (t | kk) ? 1u : 0u);
        tc_commit(bD1Full + 8 * b);
      };
      auto mma2 = [&](int c) {
        const int s = c % TC_STAGES;
        mbar_wait(bRes + 8 * s, (c / TC_STAGES) & 1);
        tc_fence_after();
        const uint32_t A[3] = {stY(s), stL(s), stY(s)}, B[3] = {stPh(s, 2), stPh(s, 2), stPh(s, 3)};
#pragma unroll
        for (int t = 0; t < 3; ++t)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            tc_mma_tf32(d2, make_desc(A[t] + 32 * kk), make_desc(B[t] + 32 * kk), idesc, (c | t | kk) ? 1u : 0u);
        tc_commit(bEmpty + 8 * s);
      };
      mbar_wait(bCFull, 0);
      for (int c = 0; c < n; ++c) {
        mma1(c);
        if (BWD && c >= 1) mma2(c - 1);
      }
      if (BWD) {
        mma2(n - 1);
        tc_commit(bD2Full);
      }
    }
  } else {
    // ===================== epilogue: 8 warps =====================
    const int e = warp - 2, q = warp & 3, h = e >> 2;  // TMEM lane quarter of this warp, column half
    const int r = 32 * q + lane;                       // gene row inside the tile = TMEM lane
    const uint32_t trow = tmem + ((uint32_t)(32 * q) << 16);
    float acc0 = 0.f, acc1 = 0.f, lt = 0.f, phi = 0.f;
    if (LIK == BVG_NB) { lt = (float)a.zeta[a.tail_idx]; phi = __expf(lt); }
    for (int c = 0; c < n; ++c) {
      const int s = c % TC_STAGES, b = c & 1;
      mbar_wait(bFull + 8 * s, (c / TC_STAGES) & 1);   // Y tile landed (acquire for this thread)
      mbar_wait(bD1Full + 8 * b, (c >> 1) & 1);        // MMA1 of this chunk retired
      tc_fence_after();
      float f[16];
      tmem_ld16(trow + 32 * b + 16 * h, f);
      tc_fence_before();
      mbar_arrive(bD1Empty + 8 * b);
      const uint32_t yrow = stY(s) + r * 128, lrow = stL(s) + r * 128;
      const int64_t spot0 = (int64_t)(c_begin + c) * TC_CHUNK + 16 * h;
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        const uint32_t off = (uint32_t)(((4 * h + c4) ^ (r & 7)) << 4);  // 128B swizzle: 16B chunk index ^ (row & 7)
        const float4 y4 = lds128(yrow + off);
        const float yv[4] = {y4.x, y4.y, y4.z, y4.w};
        float hi[4], lo[4];
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
          hi[i] = tf32_rna(res);
          lo[i] = res - hi[i];
        }
        if (BWD) {
          sts128(yrow + off, make_float4(hi[0], hi[1], hi[2], hi[3]));
          sts128(lrow + off, make_float4(lo[0], lo[1], lo[2], lo[3]));
        }
      }
      if (BWD) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the MMA (async proxy)
        mbar_arrive(bRes + 8 * s);
      } else {
        mbar_arrive(bEmpty + 8 * s);
      }
    }
    // ---- drain: P' = D2 (k+1 values per gene) and the scalar partials ----
    const int64_t g = (int64_t)tile * BVG_GENE_TILE + r;
    const int64_t row = (int64_t)split * a.Gpad + g;
    if (BWD) {
      float p[16];
      if (n > 0) {
        mbar_wait(bD2Full, 0);
        tc_fence_after();
        tmem_ld16(trow + 64 + 16 * h, p);
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
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
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
  float *CH = (float *)m.C, *CL = CH + (size_t)m.Gpad * TC_KW;  // k_prep writes the hi/lo coefficient rows here
  BVG_TRY(make_map(&t->tmY, f->Y, (uint64_t)f->ldY, (uint64_t)m.G, (uint64_t)f->ldY, BVG_GENE_TILE));
  BVG_TRY(make_map(&t->tmPhiH, t->PhiH, TC_KW, (uint64_t)f->Mpad, TC_KW, 32));
  BVG_TRY(make_map(&t->tmPhiL, t->PhiL, TC_KW, (uint64_t)f->Mpad, TC_KW, 32));
  BVG_TRY(make_map(&t->tmPhiTH, t->PhiTH, (uint64_t)f->Mpad, TC_KW, (uint64_t)f->Mpad, 32));
  BVG_TRY(make_map(&t->tmPhiTL, t->PhiTL, (uint64_t)f->Mpad, TC_KW, (uint64_t)f->Mpad, 32));
  BVG_TRY(make_map(&t->tmCH, CH, TC_KW, (uint64_t)m.Gpad, TC_KW, BVG_GENE_TILE));
  BVG_TRY(make_map(&t->tmCL, CL, TC_KW, (uint64_t)m.Gpad, TC_KW, BVG_GENE_TILE));
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

int launch_lik_tc(bvg_fit *f, bool bwd) {
  const DevModel &m = f->m;
  TcState *t = (TcState *)f->tc;
  TcArgs a;
  a.Ppart = (float *)m.Ppart; a.Spart = (float *)m.Spart;
  a.M = m.M; a.G = m.G; a.Gpad = m.Gpad; a.KP = m.KP;
  a.nchunks = (int)((m.M + TC_CHUNK - 1) / TC_CHUNK);
  a.cps = (a.nchunks + m.nsplit - 1) / m.nsplit;
  a.zeta = m.zeta; a.tail_idx = m.L.tail;
  dim3 grid((unsigned)(m.Gpad / BVG_GENE_TILE), (unsigned)m.nsplit);
#define TCL(LIK, BWD) k_lik_tc<LIK, BWD><<<grid, TC_THREADS, TC_SMEM_BYTES, f->stream>>>(t->tmY, t->tmPhiH, t->tmPhiL, t->tmPhiTH, t->tmPhiTL, t->tmCH, t->tmCL, a)
  if (m.lik == BVG_GAUSSIAN) { if (bwd) TCL(BVG_GAUSSIAN, true); else TCL(BVG_GAUSSIAN, false); }
  else { if (bwd) TCL(BVG_NB, true); else TCL(BVG_NB, false); }
#undef TCL
  f->st.kernel_launches++;
  CUDA_TRY(cudaGetLastError());
  return BVG_OK;
}
