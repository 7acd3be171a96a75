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

#include <vector>
#include <cuda_bf16.h>

#include "tc_roles.cuh"

// ------------------------------------------------------------------------------------------------
// one pass per launch; the warp roles live in tc_roles.cuh (shared with the persistent ADVI kernel)
template <int LIK, bool BWD>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_lik_tc(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmF1,
         const __grid_constant__ CUtensorMap tmF2, const __grid_constant__ CUtensorMap tmT1,
         const __grid_constant__ CUtensorMap tmT2, TcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const bool tr0 = a.trace && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0;
  if (tr0) { a.trace[8 * 30 + 0] = clock64(); asm volatile("mov.u64 %0, %globaltimer;" : "=l"(a.trace[8 * 30 + 4])); }
  TcCtx x = tc_ctx_make(smem_raw, 4000000000LL);
  __shared__ double s_half[2][BVG_GENE_TILE];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, split = blockIdx.y;
  const int c_begin = split * a.cps;
  int n = a.nchunks - c_begin;
  n = n < a.cps ? n : a.cps;
  if (n < 0) n = 0;
  long long *const trace00 = (a.trace && blockIdx.x == 0 && blockIdx.y == 0) ? a.trace : nullptr;

  griddep_launch();  // PDL: the successor (k_gene) may begin launching; it waits for this grid to finish
  tc_setup_cta(x);
  if (tr0) a.trace[8 * 30 + 1] = clock64();

  if (warp == TC_WARP_Y) {
    if (lane == 0) tc_role_produce_y(x, &tmY, tile, a.nch32, c_begin, 0, n, 0u);
  } else if (warp == TC_WARP_BASIS) {
    // ===================== TMA producer, basis ring (lane 0); nb histogram terms (lanes 1..31) =====
    const int n0 = n < TC_NB ? n : TC_NB;
    if (lane == 0) tc_role_produce_basis<BWD>(x, &tmF1, &tmF2, &tmT1, &tmT2, c_begin, 0, n0, 0u);  // the whole ring is in flight before anything else happens in this warp
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
    if (lane == 0) tc_role_produce_basis<BWD>(x, &tmF1, &tmF2, &tmT1, &tmT2, c_begin, n0, n, 0u);
  } else if (warp == TC_WARP_MMA1) {
    tc_role_mma1<BWD>(x, n, a.nk1, 0u, 0u, trace00);
  } else if (warp == TC_WARP_MMA2) {
    if (BWD) tc_role_mma2(x, n, 0u, trace00);
  } else {
    // ===================== epilogue: 8 warps =====================
    const int e = warp - TC_WARP_EPI0, qd = warp & 3, h = e >> 2;  // TMEM lane quarter of this warp, 32-spot half of the chunk
    const int r = 32 * qd + lane;                                  // gene row inside the tile = TMEM lane
    const uint32_t trow = x.tmem + ((uint32_t)(32 * qd) << 16);
    const int64_t g = (int64_t)tile * BVG_GENE_TILE + r;
    {  // coefficient rows -> TMEM (A operand of MMA1), once per CTA, as two packed bf16 parts
      // PDL: everything above (barriers, TMEM, and the producer's first TMA loads) overlaps the tail of
      // k_prep; its output is first touched here
      griddep_wait();
      float cv[16];
      {
        const float4 *pw = reinterpret_cast<const float4 *>(a.W + g * TC_KW + 16 * h);
        const float4 *pm = reinterpret_cast<const float4 *>(a.U + 16 * h), *ps = reinterpret_cast<const float4 *>(a.U + 32 + 16 * h);
        const float Ag = a.W[g * TC_KW + a.k + 1], mb = a.U[64], sb = a.U[65];
        float w[16], um[16], us[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 xx = pw[i], y = __ldg(pm + i), z = __ldg(ps + i);
          w[4 * i] = xx.x; w[4 * i + 1] = xx.y; w[4 * i + 2] = xx.z; w[4 * i + 3] = xx.w;
          um[4 * i] = y.x; um[4 * i + 1] = y.y; um[4 * i + 2] = y.z; um[4 * i + 3] = y.w;
          us[4 * i] = z.x; us[4 * i + 1] = z.y; us[4 * i + 2] = z.z; us[4 * i + 3] = z.w;
        }
        tc_epi_form_coeff(a.k, h, g < a.G, Ag, mb, sb, w, um, us, cv);
      }
      tc_epi_store_coeff(x, trow, h, cv);
    }
    double sum0 = 0.0, sum1 = 0.0;
    float lt = 0.f, phi = 0.f;
    if (LIK == BVG_NB) { lt = (float)a.zeta[a.tail_idx]; phi = __expf(lt); }
    tc_role_epilogue<LIK, BWD>(x, trow, r, h, c_begin, n, 0u, a.M, lt, phi, sum0, sum1, (trace00 && threadIdx.x == 32 * TC_WARP_EPI0) ? trace00 : nullptr);
    // ---- drain: P' = D2 (k+1 values per gene) and the scalar partials ----
    const int64_t row = (int64_t)split * a.Gpad + g;
    if (BWD) {
      float p[16];
      tc_epi_read_d2(x, trow, h, n, 0u, p);
      tc_epi_store_p(a.Ppart + row * a.KP, h, a.KP, p);
    }
    if (h == 1) { s_half[0][r] = sum0; s_half[1][r] = sum1; }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (h == 0) {
      a.Spart[row * 2 + 0] = sum0 + s_half[0][r];
      a.Spart[row * 2 + 1] = sum1 + s_half[1][r];
    }
  }
  if (tr0) a.trace[8 * 30 + 2] = clock64();
  tc_teardown_cta(x);
  if (tr0) { a.trace[8 * 30 + 3] = clock64(); asm volatile("mov.u64 %0, %globaltimer;" : "=l"(a.trace[8 * 30 + 5])); }
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
  persist_destroy(f);
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

// Measurement support (bvg_fit_time_lik_kernel, mode 2): the likelihood kernel timed back to back on an input LARGER THAN L2
// without flush kernels inside the timed region -- `ncopy` copies of the tile-major Y (4 x 60 MB at config 2 > 126 MB of L2),
// launch i reads copy i mod ncopy, so every launch streams its Y from HBM while the launches queue behind each other as they
// do in a real evaluation chain.  One event pair around all n launches.
int tc_time_rotating(bvg_fit *f, int n, int ncopy, bool bwd, float *ms_avg) {
  TcState *t = (TcState *)f->tc;
  if (!t || !f->tiledY) { bvg_set_error("rotating-copy timing needs the fused tcgen05 engine"); return BVG_ERR_STATE; }
  const DevModel &m = f->m;
  const size_t y_bytes = sizeof(float) * (size_t)m.Gpad * f->ldY;
  std::vector<void *> copies(ncopy, nullptr);
  std::vector<CUtensorMap> maps(ncopy);
  const CUtensorMap own = t->tmY;
  int rc = BVG_OK;
  for (int c = 0; c < ncopy && rc == BVG_OK; ++c) {
    if (cudaMalloc(&copies[c], y_bytes) != cudaSuccess) { cudaGetLastError(); bvg_set_error("out of device memory for the rotating copies of Y"); rc = BVG_ERR_CUDA; break; }
    if (cudaMemcpyAsync(copies[c], f->Y, y_bytes, cudaMemcpyDeviceToDevice, f->stream) != cudaSuccess) { rc = BVG_ERR_CUDA; break; }
    rc = tc_make_map(&maps[c], copies[c], false, 32, (uint64_t)(m.Gpad / BVG_GENE_TILE) * (uint64_t)(f->ldY / 32) * BVG_GENE_TILE, 32, 2 * BVG_GENE_TILE);
  }
  if (rc == BVG_OK) {
    for (int i = 0; i < ncopy && rc == BVG_OK; ++i) { t->tmY = maps[i]; rc = launch_lik_tc(f, bwd); }   // warm-up: one pass over every copy
    if (rc == BVG_OK && cudaEventRecord(f->ev0, f->stream) != cudaSuccess) rc = BVG_ERR_CUDA;
    for (int i = 0; i < n && rc == BVG_OK; ++i) { t->tmY = maps[i % ncopy]; rc = launch_lik_tc(f, bwd); }
    if (rc == BVG_OK && (cudaEventRecord(f->ev1, f->stream) != cudaSuccess || cudaEventSynchronize(f->ev1) != cudaSuccess)) rc = BVG_ERR_CUDA;
    float ms = 0.f;
    if (rc == BVG_OK && cudaEventElapsedTime(&ms, f->ev0, f->ev1) != cudaSuccess) rc = BVG_ERR_CUDA;
    *ms_avg = ms / (float)n;
  }
  t->tmY = own;
  cudaStreamSynchronize(f->stream);
  for (void *c : copies) cudaFree(c);
  if (rc == BVG_ERR_CUDA) bvg_set_error("rotating-copy timing: %s", cudaGetErrorString(cudaGetLastError()));
  return rc;
}
