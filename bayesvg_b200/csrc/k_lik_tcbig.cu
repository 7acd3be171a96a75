// k_lik_tcbig.cu -- the likelihood pass on tensor cores for a LARGE basis (24 <= k+1 <= 512; BASELINE config 5 has
// k ~ 400), where the whole coefficient row no longer fits the fused kernel's TMEM / shared-memory budget.
// Two tcgen05 GEMM kernels share one warp-specialised main loop (TMA -> 3-stage smem ring -> tcgen05.mma with both
// operands K-major in shared memory, 128 x 128 fp32 accumulator in TMEM, BF16x3 split: A1 B1 + A2 B1 + A1 B2):
//   forward  (k_tcbig<FWD>): D[128 genes x 128 spots] = C'[genes x K] . Phi'[spots x K]^T           (approxGP.stan:33, :48 mean)
//             epilogue (thread = gene row): f <- TMEM, y <- global, residual r, SSE / nb terms in registers (fp64 running
//             sums), r -> global as two bf16 parts R1, R2 [gene][spot]
//   backward (k_tcbig<BWD>): D[128 basis x 128 genes] = Phi'^T[basis x spots] . R[genes x spots]^T   (P' = Phi'^T r)
//             epilogue (thread = basis row): D -> Ppart[split][gene][basis]
// The residual makes one round trip through HBM (4 B per pair written + read) -- at k ~ 400 the contraction needs
// ~4.8k flop per pair (x3 for the split), so the pass stays tensor-bound: 2.4 TFLOP per evaluation of a config-5 shard
// against 6 GB of traffic.  The per-gene kernels (k_prep_big, k_gene_big) are the same as for the CUDA-core path.
#include <string.h>

#include "bvg_internal.cuh"
#include "tc_ptx.cuh"

#define TB_THREADS 320   // warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2-9 epilogue (two per TMEM lane quarter)
#define TB_STAGES 3
#define TB_TILE_BYTES (128 * 128)                 // 128 rows x 64 bf16 (one 128-byte swizzled row per operand row)
#define TB_STAGE_BYTES (4 * TB_TILE_BYTES)        // A1 | A2 | B1 | B2
#define TB_STAGING_BYTES (8 * 4096)             // forward epilogue: one 4 KB transpose buffer per epilogue warp
#define TB_SMEM_BYTES (TB_STAGES * TB_STAGE_BYTES + 1024 + 256 + TB_STAGING_BYTES)

int tc_make_map(CUtensorMap *m, void *ptr, bool bf16, uint64_t cols, uint64_t rows, uint64_t ld, uint32_t box_rows);

struct TbArgs {
  // forward
  const float *Y;           // [G][ldY] fp32
  int64_t ldY, M, G, Gpad, Mpad;
  __nv_bfloat16 *R1, *R2;   // [Gpad][Mpad]
  double *Spart;            // [nsplit][Gpad][2] fp64: the per-(split, gene) sums are accumulated in fp64 and stay fp64
  // backward
  float *Ppart;             // [nsplit][Gpad][KP]
  int KP, k1;               // padded row width of Ppart, k + 1
  int nkb;                  // forward: K blocks of 64 basis columns
  int tiles_per_split;      // spot tiles (128 spots) per split
  int ntiles_m;             // spot tiles in total
  const double *zeta;
  int64_t tail_idx;
  int write_r;
};

struct TbState {
  CUtensorMap tmC1, tmC2, tmF1, tmF2, tmT1, tmT2, tmR1, tmR2;
  __nv_bfloat16 *F1, *F2, *T1, *T2, *R1, *R2, *C1, *C2;
  int Kp, njb;
};

// MODE 0: forward (A = C' parts, B = F parts, K = basis), MODE 1: backward (A = T parts, B = R parts, K = spots)
template <int MODE, int LIK>
__global__ void __launch_bounds__(TB_THREADS, 1)
k_tcbig(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
        const __grid_constant__ CUtensorMap tmB1, const __grid_constant__ CUtensorMap tmB2, TbArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sBar = base + TB_STAGES * TB_STAGE_BYTES;
  const uint32_t bFull = sBar, bEmpty = bFull + 8 * TB_STAGES, bDFull = bEmpty + 8 * TB_STAGES, bDEmpty = bDFull + 16, sTmem = bDEmpty + 16;
  const uint32_t sStage = sBar + 256;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // work decomposition
  //   forward : blockIdx.x = gene tile, blockIdx.y = split; tiles = spot tiles of the split; K loop = nkb basis blocks
  //   backward: blockIdx.x = basis block, blockIdx.y = gene tile, blockIdx.z = split; one tile; K loop = spot blocks of the
  //             split.  Basis blocks vary fastest so that the CTAs sharing an R tile (same gene tile and split) and the
  //             CTAs sharing a Phi'^T slice (same split) run in the same wave and hit L2: with splits fastest every
  //             basis block re-read R from HBM and the kernel sat on the HBM roofline (4 GB per launch, 628 us).
  const int gt = MODE == 1 ? blockIdx.y : blockIdx.x, split = MODE == 1 ? blockIdx.z : blockIdx.y, jb = MODE == 1 ? blockIdx.x : 0;
  const int t_begin = split * a.tiles_per_split;
  int nt = a.ntiles_m - t_begin;
  nt = nt < a.tiles_per_split ? nt : a.tiles_per_split;
  if (nt < 0) nt = 0;
  const int ntile = MODE == 0 ? nt : (nt > 0 ? 1 : 0);   // accumulator tiles this CTA produces
  const int nk = MODE == 0 ? a.nkb : 2 * nt;              // K blocks (64 wide) per accumulator tile

  griddep_launch();
  if (threadIdx.x == 0) {
    for (int s = 0; s < TB_STAGES; ++s) { mbar_init(bFull + 8 * s, 1); mbar_init(bEmpty + 8 * s, 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(bDFull + 8 * b, 1); mbar_init(bDEmpty + 8 * b, 256); }  // two TMEM accumulators
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sTmem), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(sTmem));
  griddep_wait();  // operands written by the previous kernel (C' parts by k_prep_big, R parts by the forward kernel)

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int it = 0;
      for (int t = 0; t < ntile; ++t) {
        for (int kb = 0; kb < nk; ++kb, ++it) {
          const int s = it % TB_STAGES;
          if (it >= TB_STAGES) mbar_wait(bEmpty + 8 * s, ((it / TB_STAGES) - 1) & 1);
          const uint32_t st = base + s * TB_STAGE_BYTES;
          mbar_expect_tx(bFull + 8 * s, TB_STAGE_BYTES);
          int ac, ar, bc, br;
          if (MODE == 0) { ac = kb * 64; ar = gt * 128; bc = kb * 64; br = (t_begin + t) * 128; }
          else { ac = t_begin * 128 + kb * 64; ar = jb * 128; bc = ac; br = gt * 128; }
          tma_load_2d(st, &tmA1, bFull + 8 * s, ac, ar);
          tma_load_2d(st + TB_TILE_BYTES, &tmA2, bFull + 8 * s, ac, ar);
          tma_load_2d(st + 2 * TB_TILE_BYTES, &tmB1, bFull + 8 * s, bc, br);
          tma_load_2d(st + 3 * TB_TILE_BYTES, &tmB2, bFull + 8 * s, bc, br);
          // The 3-stage ring covers ~2 K blocks (~0.8 us); the basis tiles stream from HBM: ask L2 for the boxes of the
          // NEXT spot tile.  (The same prefetch 6 K blocks ahead in the backward kernel made it 8 % slower -- it is bound
          // by L2 -> SM bandwidth, 13 TB/s aggregate, not by HBM latency -- and was removed.)
          if (MODE == 0 && t + 1 < ntile) { tma_prefetch_2d(&tmB1, bc, br + 128); tma_prefetch_2d(&tmB2, bc, br + 128); }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (convergent warp, one elected lane) =====================
    constexpr uint32_t idesc = make_idesc_bf16(128, 128);
    int it = 0;
    for (int t = 0; t < ntile; ++t) {
      const int b = t & 1;  // accumulator buffer: the epilogue of tile t - 1 overlaps this tile's main loop
      if (t >= 2) mbar_wait(bDEmpty + 8 * b, ((t >> 1) - 1) & 1);  // the epilogue has drained this buffer (tile t - 2)
      tc_fence_after();
      const uint32_t dacc = tmem + 128 * b;
      for (int kb = 0; kb < nk; ++kb, ++it) {
        const int s = it % TB_STAGES;
        mbar_wait(bFull + 8 * s, (it / TB_STAGES) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t st = base + s * TB_STAGE_BYTES;
          const uint64_t a1 = make_desc(st), a2 = make_desc(st + TB_TILE_BYTES), b1 = make_desc(st + 2 * TB_TILE_BYTES),
                         b2 = make_desc(st + 3 * TB_TILE_BYTES);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {  // 4 x K = 16 inside the 64-wide swizzled block: +32 bytes = +2 in descriptor units
            tc_mma_bf16_ss(dacc, a1 + 2 * ks, b1 + 2 * ks, idesc, (kb | ks) ? 1u : 0u);
            tc_mma_bf16_ss(dacc, a2 + 2 * ks, b1 + 2 * ks, idesc, 1u);
            tc_mma_bf16_ss(dacc, a1 + 2 * ks, b2 + 2 * ks, idesc, 1u);
          }
          tc_commit(bEmpty + 8 * s);
          if (kb == nk - 1) tc_commit(bDFull + 8 * b);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue: 8 warps; thread = accumulator row (TMEM lane) x one 64-column half ================
    const int q = warp & 3, h = (warp - 2) >> 2;  // TMEM lane quarter this warp may access, column half
    const int r = 32 * q + lane;                  // accumulator row
    const uint32_t trow = tmem + ((uint32_t)(32 * q) << 16) + 64 * h;
    if (MODE == 0) {
      // Global traffic of the epilogue is COALESCED through a 4 KB per-warp staging buffer (32 rows x 128 bytes, 16-byte
      // chunks XOR-swizzled by row so that both access patterns are bank-conflict free).  The TMEM layout fixes
      // "thread = gene row", and Y / R are gene-major: the first version let every lane read / write its own row
      // directly, i.e. 32 different 128-byte lines per instruction -- ncu (profiles/r1b_ncu_full_k_tcbig.csv): tensor
      // pipe 33 % active, the kernel sat on the L1 tag stage.  Now a Y chunk (32 genes x 32 spots) is fetched with 8
      // lanes per row (4 full lines per instruction) into registers one chunk AHEAD, transposed through the buffer, and
      // the two bf16 residual parts go back the same way (8 rows x 64 bytes per store instruction).
      const int64_t g0 = (int64_t)gt * 128 + 32 * q;   // first gene row of this warp
      const int64_t g = g0 + lane;
      const uint32_t sbuf = sStage + (uint32_t)(warp - 2) * 4096u;
      double sum0 = 0.0, sum1 = 0.0;
      float lt = 0.f, phi = 0.f;
      if (LIK == BVG_NB) { lt = (float)a.zeta[a.tail_idx]; phi = __expf(lt); }
      const int yr = lane >> 3, yc = lane & 7;         // Y pattern: instruction i covers rows 4 i + yr, 16-byte chunk yc
      const int rr = lane & 7, rc = lane >> 3;         // R pattern: instruction i covers rows 8 i + rr, chunk rc of the part
      const int nchunk = 2 * ntile;                    // chunk n = (tile n >> 1, 32-spot half n & 1)
      float4 yq[8];
      auto load_y = [&](int n) {
        const float *src = a.Y + ((int64_t)(t_begin + (n >> 1)) * 128 + 64 * h + 32 * (n & 1) + 4 * yc);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int64_t gr = g0 + 4 * i + yr;
          yq[i] = gr < a.G ? __ldg(reinterpret_cast<const float4 *>(src + gr * a.ldY)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      if (nchunk > 0) load_y(0);
      for (int n = 0; n < nchunk; ++n) {
        const int t = n >> 1, c32 = n & 1, b = t & 1;
        const int64_t s0 = (int64_t)(t_begin + t) * 128 + 64 * h + 32 * c32;
        // ---- y: coalesced registers -> staging -> this thread's row ----
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = 4 * i + yr;
          sts128(sbuf + row * 128 + ((yc ^ (row & 7)) << 4), yq[i]);
        }
        __syncwarp();
        float4 y4[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) y4[c] = lds128m(sbuf + lane * 128 + ((c ^ (lane & 7)) << 4));
        __syncwarp();
        if (n + 1 < nchunk) load_y(n + 1);   // in flight during this chunk's wait + math
        if (n + 3 < nchunk && yc == 0) {     // and the chunk after the next one on its way from HBM to L2
          const float *nx = a.Y + ((int64_t)(t_begin + ((n + 3) >> 1)) * 128 + 64 * h + 32 * ((n + 3) & 1));
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int64_t gr = g0 + 4 * i + yr;
            if (gr < a.G) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + gr * a.ldY));
          }
        }
        if (c32 == 0) {
          mbar_wait(bDFull + 8 * b, (t >> 1) & 1);
          tc_fence_after();
        }
        float f[32];
        tmem_ld32(trow + 128 * b + 32 * c32, f);
        uint32_t p1[16], p2[16];
        float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          const float4 yv4 = y4[c4];
          const float yv[4] = {yv4.x, yv4.y, yv4.z, yv4.w};
          float res[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float ff = f[4 * c4 + i];
            if (LIK == BVG_GAUSSIAN) {
              res[i] = yv[i] - ff;
              acc0 = fmaf(res[i], res[i], acc0);
            } else {  // see k_lik_tc.cu: one ex2 / rcp / lg2 per element
              const float d = ff - lt;
              const float ex = ex2_approx(-fabsf(d) * 1.4426950408889634f);
              const float tt = 1.f + ex;
              const float rt = rcp_approx(tt);
              const float Lq = fmaf(lg2_approx(tt), 0.6931471805599453f, fmaxf(d, 0.f));
              const float sg = (d >= 0.f ? 1.f : ex) * rt;
              const float yp = yv[i] + phi;
              res[i] = fmaf(-yp, sg, yv[i]);
              if (s0 + 4 * c4 + i < a.M) {
                acc0 = fmaf(yv[i], d, acc0);
                acc0 = fmaf(-yp, Lq, acc0);
                acc1 += Lq;
              }
            }
          }
          split_bf16x2(res[0], res[1], p1[2 * c4], p2[2 * c4]);
          split_bf16x2(res[2], res[3], p1[2 * c4 + 1], p2[2 * c4 + 1]);
        }
        if (g < a.G) {
          sum0 += (double)acc0;
          if (LIK == BVG_NB) sum1 += (double)acc1;
        }
        if (c32 == 1) {   // both halves of the accumulator are in registers: hand the buffer back before the stores
          tc_fence_before();
          mbar_arrive(bDEmpty + 8 * b);
        }
        if (a.write_r) {
          // ---- r: this thread's row (part 1 = chunks 0-3, part 2 = chunks 4-7) -> staging -> coalesced stores ----
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            sts128u(sbuf + lane * 128 + ((w ^ (lane & 7)) << 4), p1[4 * w], p1[4 * w + 1], p1[4 * w + 2], p1[4 * w + 3]);
            sts128u(sbuf + lane * 128 + (((4 + w) ^ (lane & 7)) << 4), p2[4 * w], p2[4 * w + 1], p2[4 * w + 2], p2[4 * w + 3]);
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int row = 8 * i + rr;
            const int64_t off = (g0 + row) * a.Mpad + s0 + 8 * rc;
            *reinterpret_cast<float4 *>(a.R1 + off) = lds128m(sbuf + row * 128 + ((rc ^ (row & 7)) << 4));
            *reinterpret_cast<float4 *>(a.R2 + off) = lds128m(sbuf + row * 128 + (((4 + rc) ^ (row & 7)) << 4));
          }
          __syncwarp();
        }
      }
      // the two column halves of a row meet through the (now idle) staging buffer of the h = 1 warp (warp + 4 of the h = 0 one)
      const uint32_t sred = sStage + (uint32_t)(warp - 2 + (h == 0 ? 4 : 0)) * 4096u + lane * 16;
      if (h == 1) st_shared_f64x2(sred, sum0, sum1);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (h == 0) {
        double o0, o1;
        ld_shared_f64x2(sred, o0, o1);
        const int64_t row = (int64_t)split * a.Gpad + g;
        a.Spart[row * 2 + 0] = sum0 + o0;
        a.Spart[row * 2 + 1] = sum1 + o1;
      }
    } else {
      const int j = jb * 128 + r;  // basis index of this row
      if (ntile > 0) {
        mbar_wait(bDFull, 0);
        tc_fence_after();
      }
#pragma unroll 1
      for (int c32 = 0; c32 < 2; ++c32) {
        float p[32];
        if (ntile > 0) tmem_ld32(trow + 32 * c32, p);
        else {
#pragma unroll
          for (int i = 0; i < 32; ++i) p[i] = 0.f;
        }
        if (j < a.KP) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int64_t g = (int64_t)gt * 128 + 64 * h + 32 * c32 + i;
            a.Ppart[((int64_t)split * a.Gpad + g) * a.KP + j] = j < a.k1 ? p[i] : 0.f;   // lanes = consecutive j: coalesced per gene
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// operand staging
// ------------------------------------------------------------------------------------------------
// Phi (M x k col-major fp64) -> bf16 parts of Phi' = [Phi, 1]:  F*: [Mpad][Kp] (row = spot), T*: [Kp][Mpad] (row = basis)
__global__ void k_tb_stage_phi(const double *__restrict__ phi, int64_t M, int64_t Mpad, int k, int Kp, __nv_bfloat16 *__restrict__ F1,
                               __nv_bfloat16 *__restrict__ F2, __nv_bfloat16 *__restrict__ T1, __nv_bfloat16 *__restrict__ T2) {
  const int64_t n = Mpad * Kp;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = i / Kp;
    const int j = (int)(i - s * Kp);
    float v = 0.f;
    if (s < M && j <= k) v = j < k ? (float)phi[(int64_t)j * M + s] : 1.f;
    const __nv_bfloat16 b1 = __float2bfloat16_rn(v), b2 = __float2bfloat16_rn(v - __bfloat162float(b1));
    F1[i] = b1; F2[i] = b2;
    T1[(int64_t)j * Mpad + s] = b1; T2[(int64_t)j * Mpad + s] = b2;
  }
}
// coefficient rows (fp32 [Gpad][KP], written by k_prep_big) -> bf16 parts [Gpad][Kp]
__global__ void k_tb_split_c(const float *__restrict__ C, int64_t G, int KP, int Kp, __nv_bfloat16 *__restrict__ C1, __nv_bfloat16 *__restrict__ C2) {
  griddep_wait();
  griddep_launch();
  const int64_t n = G * KP;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t g = i / KP;
    const int j = (int)(i - g * KP);
    const float v = C[i];
    const __nv_bfloat16 b1 = __float2bfloat16_rn(v), b2 = __float2bfloat16_rn(v - __bfloat162float(b1));
    C1[g * Kp + j] = b1; C2[g * Kp + j] = b2;
  }
}

int tcbig_setup(bvg_fit *f) {
  const DevModel &m = f->m;
  TbState *t = new TbState();
  memset(t, 0, sizeof(*t));
  f->tcbig = t;
  t->Kp = (int)((m.k + 1 + 63) / 64 * 64);
  t->njb = (m.k + 1 + 127) / 128;
  const size_t nF = (size_t)f->Mpad * t->Kp, nR = (size_t)m.Gpad * f->Mpad, nC = (size_t)m.Gpad * t->Kp;
  CUDA_TRY(cudaMalloc(&t->F1, sizeof(__nv_bfloat16) * (4 * nF + 2 * nR + 2 * nC)));
  t->F2 = t->F1 + nF; t->T1 = t->F2 + nF; t->T2 = t->T1 + nF; t->R1 = t->T2 + nF; t->R2 = t->R1 + nR; t->C1 = t->R2 + nR; t->C2 = t->C1 + nC;
  CUDA_TRY(cudaMemsetAsync(t->F1, 0, sizeof(__nv_bfloat16) * (4 * nF + 2 * nR + 2 * nC), f->stream));
  k_tb_stage_phi<<<148 * 4, 256, 0, f->stream>>>(f->phi64, m.M, f->Mpad, m.k, t->Kp, t->F1, t->F2, t->T1, t->T2);
  CUDA_TRY(cudaGetLastError());
  // every operand is K-major with 128-byte (64 bf16) swizzled rows; boxes are 128 rows x 64 columns
  BVG_TRY(tc_make_map(&t->tmC1, t->C1, true, (uint64_t)t->Kp, (uint64_t)m.Gpad, (uint64_t)t->Kp, 128));
  BVG_TRY(tc_make_map(&t->tmC2, t->C2, true, (uint64_t)t->Kp, (uint64_t)m.Gpad, (uint64_t)t->Kp, 128));
  BVG_TRY(tc_make_map(&t->tmF1, t->F1, true, (uint64_t)t->Kp, (uint64_t)f->Mpad, (uint64_t)t->Kp, 128));
  BVG_TRY(tc_make_map(&t->tmF2, t->F2, true, (uint64_t)t->Kp, (uint64_t)f->Mpad, (uint64_t)t->Kp, 128));
  BVG_TRY(tc_make_map(&t->tmT1, t->T1, true, (uint64_t)f->Mpad, (uint64_t)t->Kp, (uint64_t)f->Mpad, 128));   // rows beyond Kp: zero fill
  BVG_TRY(tc_make_map(&t->tmT2, t->T2, true, (uint64_t)f->Mpad, (uint64_t)t->Kp, (uint64_t)f->Mpad, 128));
  BVG_TRY(tc_make_map(&t->tmR1, t->R1, true, (uint64_t)f->Mpad, (uint64_t)m.Gpad, (uint64_t)f->Mpad, 128));
  BVG_TRY(tc_make_map(&t->tmR2, t->R2, true, (uint64_t)f->Mpad, (uint64_t)m.Gpad, (uint64_t)f->Mpad, 128));
  CUDA_TRY(cudaFuncSetAttribute(k_tcbig<0, BVG_GAUSSIAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, TB_SMEM_BYTES));
  CUDA_TRY(cudaFuncSetAttribute(k_tcbig<0, BVG_NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, TB_SMEM_BYTES));
  CUDA_TRY(cudaFuncSetAttribute(k_tcbig<1, BVG_GAUSSIAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, TB_SMEM_BYTES));
  return BVG_OK;
}

void tcbig_destroy(bvg_fit *f) {
  TbState *t = (TbState *)f->tcbig;
  if (!t) return;
  cudaFree(t->F1);
  delete t;
  f->tcbig = nullptr;
}

int launch_lik_tcbig(bvg_fit *f, bool bwd) {
  const DevModel &m = f->m;
  TbState *t = (TbState *)f->tcbig;
  TbArgs a;
  memset(&a, 0, sizeof(a));
  a.Y = (const float *)f->Y; a.ldY = f->ldY; a.M = m.M; a.G = m.G; a.Gpad = m.Gpad; a.Mpad = f->Mpad;
  a.R1 = t->R1; a.R2 = t->R2; a.Spart = (double *)m.Spart; a.Ppart = (float *)m.Ppart; a.KP = m.KP; a.k1 = m.k + 1;
  a.nkb = t->Kp / 64;
  a.ntiles_m = (int)(f->Mpad / 128);
  a.tiles_per_split = (a.ntiles_m + m.nsplit - 1) / m.nsplit;
  a.zeta = m.zeta; a.tail_idx = m.L.tail;
  a.write_r = bwd ? 1 : 0;
  // coefficient rows -> bf16 parts
  k_tb_split_c<<<148 * 2, 256, 0, f->stream>>>((const float *)m.C, m.G, m.KP, t->Kp, t->C1, t->C2);
  CUDA_TRY(cudaGetLastError());
  const dim3 gridF((unsigned)(m.Gpad / 128), (unsigned)m.nsplit, 1), gridB((unsigned)t->njb, (unsigned)(m.Gpad / 128), (unsigned)m.nsplit);
  if (m.lik == BVG_GAUSSIAN) CUDA_TRY(launch_pdl(k_tcbig<0, BVG_GAUSSIAN>, gridF, dim3(TB_THREADS), (size_t)TB_SMEM_BYTES, f->stream, t->tmC1, t->tmC2, t->tmF1, t->tmF2, a));
  else CUDA_TRY(launch_pdl(k_tcbig<0, BVG_NB>, gridF, dim3(TB_THREADS), (size_t)TB_SMEM_BYTES, f->stream, t->tmC1, t->tmC2, t->tmF1, t->tmF2, a));
  f->st.kernel_launches += 2;
  if (bwd) {
    CUDA_TRY(launch_pdl(k_tcbig<1, BVG_GAUSSIAN>, gridB, dim3(TB_THREADS), (size_t)TB_SMEM_BYTES, f->stream, t->tmT1, t->tmT2, t->tmR1, t->tmR2, a));
    f->st.kernel_launches++;
  }
  return BVG_OK;
}
