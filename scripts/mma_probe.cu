// mma_probe.cu -- cycles per tcgen05.mma (kind::f16, M = 128, K = 16) issued back to back by one thread of one CTA per SM:
// A operand from TMEM ("ts") or shared memory ("ss"), N = 32 .. 256, one accumulator (dependent chain) or four in rotation.
// Answers: what does one small MMA of the fused likelihood kernel cost, and does it depend on N / the A source / the chain?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe scripts/mma_probe.cu && ./mma_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../bayesvg_b200/csrc/tc_ptx.cuh"

template <int N, int SS, int NACC>
__global__ void __launch_bounds__(128, 1) k_probe(int nmma, long long *out) {
  extern __shared__ uint8_t raw[];
  __shared__ uint32_t s_tmem;
  __shared__ uint64_t s_bar;
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t bar = smem_u32(&s_bar);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) ((uint32_t *)(raw + (base - smem_u32(raw))))[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  if (warp == 0) {
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t bdesc = make_desc(base), adesc = make_desc(base + 32 * 1024);
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 2; ++rep) {   // rep 0 warms up
      t0 = clock64();
      if (elect_one()) {
        constexpr uint32_t DW = N < 64 ? 32 : N;   // accumulator pitch in TMEM columns
#pragma unroll 8
        for (int i = 0; i < nmma; ++i) {   // nmma is a multiple of 8: addresses are compile-time constants of the unrolled body
          const uint32_t d = tmem + 256 + (uint32_t)(i & (NACC - 1)) * DW;
          if (SS) tc_mma_bf16_ss(d, adesc + 2 * (i & 3), bdesc + 2 * (i & 3), idesc, 1u);
          else tc_mma_bf16_ts(d, tmem + 8 * (i & 7), bdesc + 2 * (i & 3), idesc, 1u);
        }
        tc_commit(bar);
      }
      __syncwarp();
      mbar_wait(bar, rep & 1);
      t1 = clock64();
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// tcgen05.ld / tcgen05.st throughput: `nw` warps (lane quarter = warp & 3) each move `nrep` x (32 lanes x 32 columns x 4 B = 4 KB);
// mode 0: loads only, 1: stores only, 2: loads while warp 0's elected lane streams N = 32 TS MMAs (how much do they slow each other?)
__global__ void __launch_bounds__(640, 1) k_tmem_probe(int nw, int mode, int nrep, long long *out) {
  extern __shared__ uint8_t raw[];
  __shared__ uint32_t s_tmem;
  __shared__ uint64_t s_bar;
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u, bar = smem_u32(&s_bar);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) ((uint32_t *)(raw + (base - smem_u32(raw))))[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  long long t0 = clock64(), t1 = t0;
  if (warp == 0 && mode == 2) {
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t bdesc = make_desc(base);
    if (elect_one()) {
#pragma unroll 8
      for (int i = 0; i < 8 * nrep; ++i) tc_mma_bf16_ts(tmem + 448, tmem + 8 * (i & 7), bdesc + 2 * (i & 3), idesc, 1u);
      tc_commit(bar);
    }
    __syncwarp();
    mbar_wait(bar, 0);
    t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = t1 - t0;
  } else if (warp >= 4 && warp < 4 + nw) {
    const uint32_t trow = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + 64 + 32 * ((warp - 4) >> 2);
    float f[32];
    uint32_t u[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) u[i] = i;
    float acc = 0.f;
    for (int r = 0; r < nrep; ++r) {
      if (mode == 1) { tmem_st16u(trow, u); tmem_st16u(trow + 16, u); tmem_st_wait(); }
      else if (mode == 3) {   // same 4 KB as two x16 loads in flight, one wait
        tmem_ld16_sum2(trow, trow + 16, f); acc += f[3];
      } else if (mode == 4) { // 16x256b shape: 16 lanes x 256 bit x 8 = 4 KB per warp instruction
        uint32_t q[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.16x256b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
            "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]),
              "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15]), "=r"(q[16]),
              "=r"(q[17]), "=r"(q[18]), "=r"(q[19]), "=r"(q[20]), "=r"(q[21]), "=r"(q[22]), "=r"(q[23]), "=r"(q[24]),
              "=r"(q[25]), "=r"(q[26]), "=r"(q[27]), "=r"(q[28]), "=r"(q[29]), "=r"(q[30]), "=r"(q[31])
            : "r"(trow));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        acc += __uint_as_float(q[5]);
      } else if (mode == 5) {   // 4 x (32x32b.x8), one wait
        float g0[8], g1[8], g2[8];
        tmem_ld8x3(trow, trow + 8, trow + 16, g0, g1, g2); tmem_ld8(trow + 24, f); acc += g0[1] + g1[2] + g2[3] + f[4];
      } else { tmem_ld32(trow, f); acc += f[7]; }
    }
    t1 = clock64();
    if (acc == 123.f) out[3] = 1;
    if (threadIdx.x == 128 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// two issuing warps (as the fused likelihood kernel has: forward N = 64 from one warp, backward N = 32 from another), each
// with its own commit barrier: do the streams interleave at full rate?  mode 0: both warps; 1: warp 0 only issues both kinds
__global__ void __launch_bounds__(128, 1) k_two_issuers(int mode, int ngroups, long long *out) {
  extern __shared__ uint8_t raw[];
  __shared__ uint32_t s_tmem;
  __shared__ uint64_t s_bar[2];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) ((uint32_t *)(raw + (base - smem_u32(raw))))[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&s_bar[0]), 1); mbar_init(smem_u32(&s_bar[1]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  constexpr uint32_t id64 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  constexpr uint32_t id32 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint64_t bdesc = make_desc(base);
  const long long t0 = clock64();
  if (warp < 2 && (mode == 0 || warp == 0)) {
    if (elect_one()) {
      for (int g = 0; g < ngroups; ++g) {   // one "chunk": 6 forward MMAs (N = 64) and 12 backward MMAs (N = 32)
        if (warp == 0) {
#pragma unroll
          for (int i = 0; i < 6; ++i) tc_mma_bf16_ts(tmem + 256 + 64 * (g & 1), tmem + 8 * (i & 3), bdesc + 2 * (i & 1), id64, 1u);
        }
        if (warp == 1 || mode == 1) {
#pragma unroll
          for (int i = 0; i < 12; ++i) tc_mma_bf16_ts(tmem + 448, tmem + 64 + 8 * (i & 7), bdesc + 2 * (i & 3), id32, 1u);
        }
      }
      tc_commit(smem_u32(&s_bar[warp]));
    }
    __syncwarp();
    mbar_wait(smem_u32(&s_bar[warp]), 0);
    if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) out[warp] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

int main() {
  long long *d, h;
  cudaMalloc(&d, 32);
  const int nmma = 256;
  printf("{\"what\": \"cycles per tcgen05.mma kind::f16 M=128 K=16, %d back-to-back from one thread, 148 CTAs\", \"rows\": [", nmma);
  bool first = true;
#define RUN(N_, SS_, NACC_)                                                                                            \
  do {                                                                                                                 \
    cudaFuncSetAttribute(k_probe<N_, SS_, NACC_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);              \
    k_probe<N_, SS_, NACC_><<<148, 128, 64 * 1024>>>(nmma, d);                                                         \
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("\n error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; } \
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);                                                                       \
    printf("%s{\"a_from\": \"%s\", \"N\": %d, \"accumulators\": %d, \"cycles_per_mma\": %.1f}", first ? "" : ", ", SS_ ? "smem" : "tmem", N_, NACC_, (double)h / nmma); \
    first = false;                                                                                                     \
  } while (0)
  RUN(32, 0, 1); RUN(32, 0, 4); RUN(64, 0, 1); RUN(64, 0, 4); RUN(128, 0, 1); RUN(128, 0, 2); RUN(256, 0, 1);
  RUN(32, 1, 1); RUN(32, 1, 4); RUN(64, 1, 1); RUN(64, 1, 4); RUN(128, 1, 1); RUN(256, 1, 1);
  printf("], \"tmem\": [");
  cudaFuncSetAttribute(k_tmem_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  first = true;
  for (int mode = 0; mode < 6; ++mode)
    for (int nw : {4, 8, 16}) {
      long long hh[4] = {0, 0, 0, 0};
      cudaMemset(d, 0, 32);
      k_tmem_probe<<<148, 640, 64 * 1024>>>(nw, mode, 256, d);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("\n error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
      cudaMemcpy(hh, d, 32, cudaMemcpyDeviceToHost);
      printf("%s{\"mode\": \"%s\", \"warps\": %d, \"cycles_per_4KB_warp_op\": %.1f, \"bytes_per_cycle_sm\": %.1f, \"mma_cycles_each\": %.1f}", first ? "" : ", ",
             mode == 0 ? "ld 32x32b.x32" : (mode == 1 ? "st" : (mode == 2 ? "ld+mma" : (mode == 3 ? "ld 2 x (32x32b.x16), one wait" : (mode == 4 ? "ld 16x256b.x8" : "ld 3 x x8 one wait + x8")))), nw, (double)hh[0] / 256, 4096.0 * nw * 256 / (double)hh[0], mode == 2 ? (double)hh[1] / (8 * 256) : 0.0);
      first = false;
    }
  printf("], \"two_issuers\": [");
  cudaFuncSetAttribute(k_two_issuers, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int mode = 0; mode < 2; ++mode) {
    long long hh[4] = {0, 0, 0, 0};
    cudaMemset(d, 0, 32);
    k_two_issuers<<<148, 128, 64 * 1024>>>(mode, 64, d);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("\n error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    cudaMemcpy(hh, d, 32, cudaMemcpyDeviceToHost);
    printf("%s{\"issuers\": %d, \"cycles_per_chunk_of_6xN64_plus_12xN32\": %.1f, \"floor\": 402}", mode ? ", " : "", mode ? 1 : 2, (double)(hh[0] > hh[1] ? hh[0] : hh[1]) / 64);
  }
  printf("]}\n");
  return 0;
}
