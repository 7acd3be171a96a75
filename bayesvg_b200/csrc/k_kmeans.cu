// k_kmeans.cu -- centroids of the scaled spot coordinates on the device.
//
// Replaces R/findSpatiallyVariableFeaturesBayes.R:200-204 of the reference,
//   kmeans_centers <- stats::kmeans(spatial_mtx, centers = k, iter.max = 100, nstart = 10)$centers
// for the sizes where the host version stops being "cheap" (BASELINE config 5: 200,000 spots x 400 centres x 100
// iterations x 10 starts = 8e10 distance evaluations).  R runs Hartigan-Wong from starts drawn with its ambient RNG,
// so its centroids are not reproducible run to run and parity on them is statistical by
// construction; this is Lloyd-Forgy from `nstart` Philox-drawn sets of k distinct rows, best total within-cluster sum
// of squares kept -- the same objective, the same iter.max / nstart meaning.
//
// Determinism: coordinate sums and the within-SS are accumulated as 64-bit FIXED-POINT integers (2^-32 and 2^-20
// resolution), so the shared-memory / global atomics commute exactly: the result is bit-reproducible and equal to the
// serial CPU restatement (oracle/bvg_oracle.c::bvgo_kmeans), which the parity test checks with array_equal.
// One thread per spot; centres and the per-CTA accumulators live in shared memory (k <= 510: 20 KB).
#include <algorithm>
#include <vector>

#include "bvg_internal.cuh"

#define KM_FIX 4294967296.0   // 2^32: |x| < 256 and M <= 2^22 keep every sum below 2^62
#define KM_SSFIX 1048576.0    // 2^20
#define KM_STREAM 4u          // Philox stream of the start draws (1-3: DESIGN.md section 4.3)

// state of one start: iteration count, done / converged flags, within-SS of its final state (device-resident control: the
// host looks at the `done` flags of all starts once per KM_CHUNK iterations instead of synchronising after every one)
struct KmStart { int32_t it, done, conv, pad; unsigned long long ss; };
#define KM_CHUNK 8

// blockIdx.y = start: the nstart runs advance in lockstep, one launch per iteration for all of them; a start that has
// stopped is skipped, so its labels, centres and within-SS stay those of its own last iteration
__global__ void __launch_bounds__(256) k_km_assign(const double *__restrict__ X, int64_t M, const double *__restrict__ Call, int k,
                                                   int32_t *__restrict__ label_all, unsigned long long *__restrict__ acc_all /* per start: [3][k] + changed + ss */,
                                                   const KmStart *__restrict__ stt) {
  extern __shared__ unsigned long long s_acc[];  // [3k] sums | then 2k doubles of centres
  const int r = blockIdx.y;
  if (stt[r].done) return;
  const double *C = Call + (size_t)r * 2 * k;
  int32_t *label = label_all + (size_t)r * M;
  unsigned long long *acc = acc_all + (size_t)r * (3 * k + 2);
  double *sc = (double *)(s_acc + 3 * k);
  for (int i = threadIdx.x; i < 3 * k; i += blockDim.x) s_acc[i] = 0ull;
  for (int i = threadIdx.x; i < 2 * k; i += blockDim.x) sc[i] = C[i];
  __syncthreads();
  unsigned long long changed = 0, ss = 0;
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < M; s += (int64_t)gridDim.x * blockDim.x) {
    const double x = X[s], y = X[M + s];
    double best = INFINITY;
    int bj = 0;
    for (int j = 0; j < k; ++j) {
      const double dx = x - sc[j], dy = y - sc[k + j];
      const double d = fma(dx, dx, dy * dy);  // the oracle evaluates exactly this expression
      if (d < best) { best = d; bj = j; }     // ties -> lowest index
    }
    changed += label[s] != bj;
    label[s] = bj;
    ss += (unsigned long long)llrint(best * KM_SSFIX);
    atomicAdd(&s_acc[bj], (unsigned long long)llrint(x * KM_FIX));
    atomicAdd(&s_acc[k + bj], (unsigned long long)llrint(y * KM_FIX));
    atomicAdd(&s_acc[2 * k + bj], 1ull);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < k; i += blockDim.x)
    if (s_acc[2 * k + i]) {
      atomicAdd(&acc[i], s_acc[i]);
      atomicAdd(&acc[k + i], s_acc[k + i]);
      atomicAdd(&acc[2 * k + i], s_acc[2 * k + i]);
    }
  for (int o = 16; o; o >>= 1) {
    changed += __shfl_xor_sync(0xffffffffu, changed, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (changed) atomicAdd(&acc[3 * k], changed);
    atomicAdd(&acc[3 * k + 1], ss);
  }
}

// one thread block per start: stop test of the iteration just assigned (no label changed -> converged; iter.max reached ->
// stop, the within-SS was just measured against the final centres), else centre = mean of its spots (an empty cluster keeps
// its centre); clears the accumulators for the next iteration
__global__ void k_km_update(unsigned long long *__restrict__ acc_all, double *__restrict__ Call, int k, int iter_max, KmStart *__restrict__ stt) {
  const int r = blockIdx.x;
  KmStart &st = stt[r];
  if (st.done) return;
  unsigned long long *acc = acc_all + (size_t)r * (3 * k + 2);
  double *C = Call + (size_t)r * 2 * k;
  __shared__ int s_stop;
  if (threadIdx.x == 0) {
    const unsigned long long changed = acc[3 * k];
    s_stop = changed == 0ull || st.it == iter_max;
    if (s_stop) { st.ss = acc[3 * k + 1]; st.conv = changed == 0ull; }
  }
  __syncthreads();
  const bool stop = s_stop != 0;
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    const unsigned long long n = acc[2 * k + j];
    if (!stop && n) {
      C[j] = (double)(long long)acc[j] / (double)n * (1.0 / KM_FIX);
      C[k + j] = (double)(long long)acc[k + j] / (double)n * (1.0 / KM_FIX);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * k + 2; i += blockDim.x) acc[i] = 0ull;
  if (threadIdx.x == 0) { if (stop) st.done = 1; else st.it += 1; }
}

// start r: k distinct rows; row j = first Philox value (counter (j, attempt, r, stream 4)) mod M not drawn before
static void km_start_rows(uint64_t seed, int r, int64_t M, int k, int64_t *rows) {
  for (int j = 0; j < k; ++j)
    for (uint32_t a = 0;; ++a) {
      uint32_t o[4];
      philox4x32_10((uint32_t)j, a, (uint32_t)r, KM_STREAM, (uint32_t)seed, (uint32_t)(seed >> 32), o);
      const int64_t row = (int64_t)((((uint64_t)o[0] << 32) | o[1]) % (uint64_t)M);
      bool dup = false;
      for (int q = 0; q < j && !dup; ++q) dup = rows[q] == row;
      if (!dup) { rows[j] = row; break; }
    }
}

extern "C" int bvg_kmeans(int device, const double *coords, int64_t M, int k, int iter_max, int nstart, uint64_t seed,
                          double *centers_out, double *tot_withinss, int32_t *info) {
  if (!coords || !centers_out) { bvg_set_error("null argument"); return BVG_ERR_ARG; }
  if (M < 1 || k < 1 || k > BVG_MAX_K || iter_max < 1 || nstart < 1) { bvg_set_error("need M >= 1, 1 <= k <= %d, iter.max >= 1, nstart >= 1", BVG_MAX_K); return BVG_ERR_ARG; }
  if (k > M) { bvg_set_error("more cluster centers than distinct data points."); return BVG_ERR_ARG; }  // stats::kmeans' message
  if (M > ((int64_t)1 << 22)) { bvg_set_error("bvg_kmeans: M <= 4,194,304"); return BVG_ERR_ARG; }
  for (int64_t i = 0; i < 2 * M; ++i)
    if (!(fabs(coords[i]) < 256.0)) { bvg_set_error("bvg_kmeans: coordinates must be finite and scaled (|x| < 256; :160 z-scores them)"); return BVG_ERR_ARG; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); bvg_set_error("no CUDA device: bayesvg_b200 has no CPU fallback"); return BVG_ERR_CUDA; }
  CUDA_TRY(cudaSetDevice(device));
  const size_t nacc = (size_t)3 * k + 2;
  char *d = nullptr;
  const size_t bytes = sizeof(double) * (2 * (size_t)M + 2 * (size_t)k * nstart) + sizeof(unsigned long long) * nacc * nstart +
                       sizeof(KmStart) * (size_t)nstart + sizeof(int32_t) * (size_t)M * nstart;
  CUDA_TRY(cudaMalloc(&d, bytes));
  double *dX = (double *)d, *dC = dX + 2 * M;
  unsigned long long *dAcc = (unsigned long long *)(dC + 2 * (size_t)k * nstart);
  KmStart *dSt = (KmStart *)(dAcc + nacc * nstart);
  int32_t *dLab = (int32_t *)(dSt + nstart);
  cudaStream_t st = nullptr;
  int rc = BVG_OK;
  std::vector<int64_t> rows(k);
  std::vector<double> c0(2 * (size_t)k * nstart);
  std::vector<KmStart> hs(nstart);
  unsigned long long best_ss = ~0ull;
  int best_start = -1, best_iters = 0, best_conv = 0;
  const int grid = (int)std::min<int64_t>((M + 255) / 256, 148 * 8);
  const size_t smem = sizeof(unsigned long long) * 3 * k + sizeof(double) * 2 * k;
#define KM_CHECK(x) if ((x) != cudaSuccess) { bvg_set_error("bvg_kmeans: %s", cudaGetErrorString(cudaGetLastError())); rc = BVG_ERR_CUDA; goto done; }
  KM_CHECK(cudaStreamCreate(&st));
  KM_CHECK(cudaMemcpyAsync(dX, coords, sizeof(double) * 2 * M, cudaMemcpyHostToDevice, st));
  for (int r = 0; r < nstart; ++r) {
    km_start_rows(seed, r, M, k, rows.data());
    for (int j = 0; j < k; ++j) { c0[(size_t)r * 2 * k + j] = coords[rows[j]]; c0[(size_t)r * 2 * k + k + j] = coords[M + rows[j]]; }
  }
  KM_CHECK(cudaMemcpyAsync(dC, c0.data(), sizeof(double) * 2 * k * nstart, cudaMemcpyHostToDevice, st));
  KM_CHECK(cudaMemsetAsync(dLab, 0xff, sizeof(int32_t) * M * nstart, st));
  KM_CHECK(cudaMemsetAsync(dAcc, 0, sizeof(unsigned long long) * nacc * nstart + sizeof(KmStart) * nstart, st));
  // iteration = assign every spot to its nearest centre, then move the centres; a start stops when no label changed (the
  // centres then ARE the means of the final labels and ss is the within-SS of the final state) or at iter.max
  for (int done_all = 0, itc = 0; !done_all && itc <= iter_max + KM_CHUNK; itc += KM_CHUNK) {
    for (int q = 0; q < KM_CHUNK; ++q) {
      k_km_assign<<<dim3(grid, nstart), 256, smem, st>>>(dX, M, dC, k, dLab, dAcc, dSt);
      k_km_update<<<nstart, 128, 0, st>>>(dAcc, dC, k, iter_max, dSt);
    }
    KM_CHECK(cudaMemcpyAsync(hs.data(), dSt, sizeof(KmStart) * nstart, cudaMemcpyDeviceToHost, st));
    KM_CHECK(cudaStreamSynchronize(st));
    done_all = 1;
    for (int r = 0; r < nstart; ++r) done_all &= hs[r].done;
  }
  for (int r = 0; r < nstart; ++r)
    if (hs[r].ss < best_ss) { best_ss = hs[r].ss; best_start = r; best_iters = hs[r].it; best_conv = hs[r].conv; }   // ties: the first start
  KM_CHECK(cudaMemcpyAsync(centers_out, dC + (size_t)best_start * 2 * k, sizeof(double) * 2 * k, cudaMemcpyDeviceToHost, st));
  KM_CHECK(cudaStreamSynchronize(st));
  if (tot_withinss) *tot_withinss = (double)best_ss / KM_SSFIX;
  if (info) { info[0] = best_start; info[1] = best_iters; info[2] = best_conv; }
done:
  if (st) cudaStreamDestroy(st);
  cudaFree(d);
  return rc;
}
