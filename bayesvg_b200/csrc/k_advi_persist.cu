// k_advi_persist.cu -- a whole stretch of mean-field ADVI iterations in ONE cooperative launch (sm_100a).
//
// The two-launch evaluation (k_lik_tc -> k_gene + finish) spends about half of its 27 us at config 2 outside the
// likelihood pass: a second launch, a round trip of the per-(split, gene) partials through a separate kernel, a serial
// "last CTA" that reduces the cross-gene sums and updates the 2k+5 gene-shared parameters, and ~3.4 k cycles of barrier /
// TMEM / descriptor prologue per launch.  Here the CTAs of the likelihood grid stay resident for the whole stretch
// (inst/approxGP.stan:25-49 of the reference evaluated n times, stan::variational::advi::stochastic_gradient_ascent [U]):
//
//   per iteration, per CTA
//     pass      : the fused tcgen05 likelihood pass of tc_roles.cuh over the CTA's (gene tile, spot range) item(s); the TMA
//                 rings, mbarriers and TMEM buffers keep turning from one pass to the next, and the producers issue the first
//                 stages of the NEXT pass before the CTA leaves the current one (Y and the basis do not depend on the draw)
//     tile sync : the nsplit CTAs of a gene tile publish P' / SSE partials and meet on a per-tile arrival counter
//     gene phase: every CTA of the tile owns ~128 / nsplit of its genes: fixed-order sum of the split partials, per-gene
//                 gradient, mean-field update, next draw (its normal is generated while the CTA waits on the tile counter)
//     grid sync : one arrival counter for the block rows of the 2k+10 cross-gene sums
//     finish    : EVERY CTA reduces the rows in the same fixed order and updates its own replica of the gene-shared
//                 parameters (identical arithmetic on identical inputs: no broadcast, no serial last CTA); multi-GPU: CTA 0
//                 pushes the local sums into the peers' exchange buffers, every CTA polls the local buffer and sums in rank
//                 order (bvg_internal.cuh protocol), so the ranks' replicas stay bit-identical as well
//
// Determinism: every reduction has a fixed order; no floating-point atomics.  The per-gene arithmetic is k_gene's, statement
// for statement (the split sums are bit-identical); the cross-gene sums are added in a different (fixed) order than
// finish_body's, so trajectories agree with the two-launch path to rounding (tests/test_gpu_parity.py).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "tc_roles.cuh"

#define PS_NW (TC_THREADS / 32)  // 12 warps
#define PS_ROW 56                // 2k+10 <= 56 (k <= 23)
#define PS_GROW 64               // row pitch of the block rows in global memory
#define PS_NROWS 150             // rows per parity: gridDim.x <= 148 block rows, the rest stay zero (6 groups x 25 rows)
#define PS_FX (3 * PS_GROW + 8)  // per buffer: three limbs of 64 words, poison word
#define PS_NSH 52                // 2k+5 <= 51
#define PS_NPRE 24               // gene slots of a CTA's first item whose next normals are drawn ahead of the gene phase
#define PS_KW 25                 // k + 2 <= 25 parameters per gene
#define PS_NRM_LANES ((PS_NW - 2) * 32)   // warps 2 and 11 (backward-MMA issuer, basis producer) own the 2k+5 gene-shared parameters between passes; the other ten draw normals

struct PersistArgs {
  TcArgs a;
  DevModel m;
  int n_iter, nitems, nsplit, gs;   // gs = genes of a tile owned by each of its splits in the gene phase
  int pre_y, pre_b;                 // stages the producers run ahead into the next pass (<= TC_NY, <= TC_NB - 1)
  int xchg_all;                     // multi-GPU: 1 = every CTA polls the exchange buffer itself, 0 = CTA 0 polls and publishes the totals
  uint32_t *tile_cnt;               // [ntiles] arrival counters (zeroed before the launch; monotonic inside it)
  uint32_t *grid_cnt;               // [1]
  double *rows;                     // [2 parities][gridDim.x][PS_GROW] block rows of the cross-gene sums
  double *xtot;                     // [2 parities][PS_GROW] multi-GPU: the exchanged totals, published by CTA 0
  unsigned long long *fx;           // [3][PS_FX] fixed-point accumulators of the cross-gene sums (see the grid meeting)
  uint32_t *grid_cnt2;              // [1] second meeting of an iteration whose sums left the accumulators' range
  int force_rows;                   // tests (BVG_PERSIST_FORCE_ROWS=1): always take that second form
  uint32_t *xflag;                  // [1] multi-GPU: iterations whose totals CTA 0 has published
  long long timeout;                // SM cycles any intra-GPU wait may last before the kernel traps
  long long *trace;                 // optional [32 iterations][8] clock64 stamps of CTA 0, then [32 chunks][8] of the pass of iteration 8 (diagnostics)
};

// state that lives across passes (static shared memory; the TMA rings own the dynamic part)
struct PersistSmem {
  double nrm[PS_NPRE * PS_KW];  // normals of the next draw of the per-gene parameters this CTA owns (first item, slots 0..23)
  double nrm_sh[2][PS_NSH];     // normals of the next draw of the gene-shared parameters (double-buffered: read while the next are drawn)
  double st[4][PS_NSH];         // replica of the gene-shared parameters: mu, omega, step-size history of mu / omega
  double zsh[PS_NSH];           // their current draw (compact index of bvg_internal.cuh: shared_index)
  double shx[32];               // exponentials of the current draw (DevModel::shx)
  double tot[PS_ROW];           // cross-gene sums of this evaluation
  float U[96];                  // DevModel::U of the current draw
};
// scratch between passes = the ONE basis-ring stage the producer leaves unfilled when it runs ahead into the next pass
struct PersistScratch {
  double wrow[PS_NW][PS_ROW];         // per-warp partial rows | reduce scratch | exchange scratch
  double half[2][BVG_GENE_TILE];      // sums of the second 32-spot half of every gene row (drain of a pass)
};
static_assert(sizeof(PersistScratch) <= BSTAGE_BYTES, "scratch must fit one basis stage");

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(uint32_t *p, uint32_t v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// one thread: spin until *p >= target (monotonic counter)
__device__ __forceinline__ void wait_counter(const uint32_t *p, uint32_t target, long long timeout) {
  const long long t0 = clock64();
  uint32_t spins = 0;
  while ((int32_t)(ld_acquire_u32(p) - target) < 0) {
    if ((++spins & 255u) == 0u && clock64() - t0 > timeout) __trap();  // protocol bug: fail fast instead of hanging the GPU
  }
}

template <int LIK>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_advi_persist(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmF1,
               const __grid_constant__ CUtensorMap tmF2, const __grid_constant__ CUtensorMap tmT1,
               const __grid_constant__ CUtensorMap tmT2, const __grid_constant__ PersistArgs p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ PersistSmem S;
  const TcArgs &a = p.a;
  const DevModel &m = p.m;
  TcCtx x = tc_ctx_make(smem_raw, p.timeout);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cta = blockIdx.x, ncta = gridDim.x;
  const int k = m.k, KP = m.KP, NS = bvg_ns(k), NSH = 2 * k + 5, npg = k + 2;
  const bool dep = m.depth != 0;
  const Layout &L = m.L;
  const bool tr = p.trace && cta == 0 && tid == 0;
  const uint32_t smem_base_u32 = smem_u32(smem_raw);

  tc_setup_cta(x);
  // replica of the gene-shared state and of the current draw (made by k_prep just before this launch)
  if (tid < NSH) {
    const int64_t li = shared_index(L, k, tid);
    S.zsh[tid] = m.zeta[li];
    S.st[0][tid] = m.mu[li]; S.st[1][tid] = m.omega[li]; S.st[2][tid] = m.hist[li]; S.st[3][tid] = m.hist[L.D + li];
  }
  if (tid < 4 + k) S.shx[tid] = m.shx[tid];
  if (tid < 96) S.U[tid] = m.U[tid];
  const int it0 = m.ctl->it;
  const uint32_t tg0 = m.ctl->t_grad;
  const double eta = m.ctl->eta;
  uint32_t xseq = m.peer.world > 1 ? *m.peer.seq : 0u;
  bool peer_dead = false;
  int nfallback = 0;   // iterations of this launch whose sums met as rows (uniform over the grid)

  const int nmy = cta < p.nitems ? (p.nitems - cta + ncta - 1) / ncta : 0;
  auto item_geom = [&](int j, int &tile, int &split, int &c_begin, int &n) {
    const int id = cta + j * ncta;
    tile = id / p.nsplit;
    split = id - tile * p.nsplit;
    c_begin = split * a.cps;
    n = a.nchunks - c_begin;
    n = n < a.cps ? n : a.cps;
    if (n < 0) n = 0;
  };
  // global (RNG) index of parameter prm of gene g
  auto global_index = [&](int64_t g, int prm) {
    const int64_t gg = m.gene_offset + g;
    return (uint64_t)(prm < k ? m.LG.z_alpha + gg * k + prm : (prm == k ? m.LG.z_beta0 + gg : m.LG.amp + gg));
  };
  const bool has = lane < npg && !(dep && lane == k);  // depth-adjusted programs have no z_beta0: lane k carries gene_depths[g]
  // the genes whose next normals are drawn ahead: the first PS_NPRE slots of this CTA's first item
  int tile0 = 0, split0 = 0, cb0 = 0, n0_ = 0;
  if (nmy > 0) item_geom(0, tile0, split0, cb0, n0_);
  const int r_lo0 = split0 * p.gs;
  const int r_hi0 = (r_lo0 + p.gs < BVG_GENE_TILE) ? r_lo0 + p.gs : BVG_GENE_TILE;
  // (with nsplit * gs > 128 the last splits of a tile own no genes: r_lo0 >= r_hi0)
  const int nslot0 = (nmy > 0 && r_hi0 > r_lo0) ? ((r_hi0 - r_lo0) < PS_NPRE ? (r_hi0 - r_lo0) : PS_NPRE) : 0;
  // normals of draw `t` for the ahead slots and (buffer sb) for the gene-shared parameters; nl = this thread's index among
  // the nthr threads doing it
  auto draw_normals = [&](uint32_t t, int sb, int nl, int nthr) {
    const int NN = nslot0 * npg + NSH;
    for (int nid = nl; nid < NN; nid += nthr) {
      if (nid < nslot0 * npg) {
        const int slot = nid / npg, prm = nid - slot * npg;
        const int64_t g = (int64_t)tile0 * BVG_GENE_TILE + r_lo0 + slot;
        if (g < m.G && !(dep && prm == k)) S.nrm[slot * PS_KW + prm] = bvg_normal(m.seed, STREAM_GRAD, t, global_index(g, prm));
      } else {
        const int si = nid - nslot0 * npg;
        S.nrm_sh[sb][si] = bvg_normal(m.seed, STREAM_GRAD, t, (uint64_t)shared_index(m.LG, k, si));
      }
    }
  };
  if (p.n_iter > 1) draw_normals(tg0 + 1u, 0, tid, TC_THREADS);
  __syncthreads();

  uint32_t q0 = 0, pass = 0;
  int preY = 0, preB = 0;  // chunks of the upcoming pass the producers have already issued
  const float *Pp = a.Ppart;
  const double *Sp = a.Spart;
  // epilogue threads of quarters 0 / 1: the fp32 draw row of their gene, fetched right after the grid barrier of the
  // previous iteration (first item only)
  float4 wpre[4] = {};
  float wAg = 0.f;

  for (int i = 0; i < p.n_iter; ++i) {
    const int it = it0 + i;
    const double es = eta / sqrt((double)it);  // Ctl::es
    const bool ahead = i + 1 < p.n_iter;       // the last evaluation of a stretch draws nothing (as advi_steps)
    const uint32_t tg_next = tg0 + (uint32_t)i + 1u;
    if (tr && i < 32) p.trace[i * 16 + 0] = clock64();
    double hlg = 0.0, hdg = 0.0;  // nb: this CTA's share of the lgamma / digamma histogram sums (basis-producer warp)
    PersistScratch *SC = nullptr;

    // =================================== likelihood passes ===================================
    for (int j = 0; j < nmy; ++j) {
      int tile, split, c_begin, n;
      item_geom(j, tile, split, c_begin, n);
      const bool has_next = j + 1 < nmy || ahead;
      int ntile = 0, nsplit_ = 0, nc_begin = 0, nn = 0;
      if (has_next) item_geom(j + 1 < nmy ? j + 1 : 0, ntile, nsplit_, nc_begin, nn);
      // the basis producer runs at most TC_NB - 1 stages ahead into the next pass: the stage of this pass's LAST chunk stays
      // free until the next pass begins and is the CTA's scratch for the drain and for everything between the passes
      const int preBn = has_next ? (nn < p.pre_b ? nn : p.pre_b) : 0;
      SC = reinterpret_cast<PersistScratch *>(smem_raw + (x.stF((int)((q0 + (uint32_t)n + (uint32_t)preBn) % TC_NB), 0) - smem_base_u32));
      long long *const ctrace = (p.trace && cta == 0 && i == 8 && j == 0 && n <= 31) ? p.trace + 32 * 16 : nullptr;
      if (warp == TC_WARP_Y) {
        const int pre = has_next ? (nn < p.pre_y ? nn : p.pre_y) : 0;
        if (lane == 0) {
          tc_role_produce_y(x, &tmY, tile, a.nch32, c_begin, preY, n, q0);
          tc_role_produce_y(x, &tmY, ntile, a.nch32, nc_begin, 0, pre, q0 + (uint32_t)n);
        }
        preY = pre;
        __syncwarp();
      } else if (warp == TC_WARP_BASIS) {
        if (lane == 0) {
          tc_role_produce_basis<true>(x, &tmF1, &tmF2, &tmT1, &tmT2, c_begin, preB, n, q0);
          tc_role_produce_basis<true>(x, &tmF1, &tmF2, &tmT1, &tmT2, nc_begin, 0, preBn, q0 + (uint32_t)n);
        }
        preB = preBn;
        double lg = 0.0, dg = 0.0;
        if (LIK == BVG_NB && j == 0) {
          // lgamma / digamma terms of the log density depend only on (distinct count c, phi_nb): lanes 1..31 take them in turn
          if (lane > 0) {
            const double phi = S.shx[2], lg0 = lgamma(phi), dg0 = dev_digamma(phi);
            for (int e = cta * 31 + lane - 1; e < a.nhist; e += ncta * 31) {
              const double cv = a.hist_val[e], nv = a.hist_cnt[e];
              lg += nv * (lgamma(cv + phi) - lg0);
              dg += nv * (dev_digamma(cv + phi) - dg0);
            }
          }
        }
        __syncwarp();
        if (LIK == BVG_NB && j == 0) { hlg = warp_sum(lg); hdg = warp_sum(dg); }
      } else if (warp == TC_WARP_MMA1) {
        tc_role_mma1<true>(x, n, a.nk1, q0, pass, ctrace);
      } else if (warp == TC_WARP_MMA2) {
        tc_role_mma2(x, n, q0, ctrace);
      } else {
        // ----- epilogue warps: coefficient rows -> TMEM, the chunk loop, drain -----
        const int e = warp - TC_WARP_EPI0, qd = warp & 3, h = e >> 2;
        const int r = 32 * qd + lane;
        const uint32_t trow = x.tmem + ((uint32_t)(32 * qd) << 16);
        const int64_t g = (int64_t)tile * BVG_GENE_TILE + r;
        const bool etr = ctrace && tid == 32 * TC_WARP_EPI0;
        {
          float cv[16];
          if (!(j == 0 && i > 0)) {  // not prefetched
            const float4 *pw = reinterpret_cast<const float4 *>(a.W + g * TC_KW + 16 * h);
#pragma unroll
            for (int q = 0; q < 4; ++q) wpre[q] = __ldcg(pw + q);
            wAg = __ldcg(a.W + g * TC_KW + k + 1);
          }
          {  // c'_g = (A_g (mu_alpha + sigma_alpha z), mu_beta0 + sigma_beta0 z_b): tc_epi_form_coeff with U read from the replica
            const float mb = S.U[64], sbf = S.U[65];
            const bool live = g < a.G;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float wv[4] = {wpre[q].x, wpre[q].y, wpre[q].z, wpre[q].w};
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const int jj = 16 * h + 4 * q + t;
                float v = 0.f;
                if (jj < k) v = wAg * fmaf(S.U[32 + jj], wv[t], S.U[jj]);   // A_g * alpha_t[j,g]  (approxGP.stan:27, :32)
                else if (jj == k) v = fmaf(sbf, wv[t], mb);                 // beta0[g]            (approxGP.stan:26 | approxGP2.stan:46)
                cv[4 * q + t] = live ? v : 0.f;
              }
            }
          }
          tc_epi_store_coeff(x, trow, h, cv);
        }
        if (etr) ctrace[31 * 8 + 0] = clock64();
        double sum0 = 0.0, sum1 = 0.0;
        float lt = 0.f, phi = 0.f;
        if (LIK == BVG_NB) { lt = (float)S.zsh[4]; phi = __expf(lt); }
        tc_role_epilogue<LIK, true>(x, trow, r, h, c_begin, n, q0, a.M, lt, phi, sum0, sum1, etr ? ctrace : nullptr);
        if (etr) ctrace[31 * 8 + 1] = clock64();
        const int64_t row = (int64_t)split * a.Gpad + g;
        float pv[16];
        tc_epi_read_d2(x, trow, h, n, pass, pv);
        if (etr) ctrace[31 * 8 + 2] = clock64();
        tc_epi_store_p(a.Ppart + row * KP, h, KP, pv);
        // the scratch stage is free once the last backward MMA of this pass has retired (every epilogue thread waited for it above)
        if (h == 1) { SC->half[0][r] = sum0; SC->half[1][r] = sum1; }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (h == 0) {
          a.Spart[row * 2 + 0] = sum0 + SC->half[0][r];
          a.Spart[row * 2 + 1] = sum1 + SC->half[1][r];
        }
        if (etr) ctrace[31 * 8 + 3] = clock64();
      }
      q0 += (uint32_t)n;
      ++pass;
      __syncthreads();  // this CTA's partials are written
      if (tid == 0) red_release_add(p.tile_cnt + tile, 1u);
    }
    if (tr && i < 32) p.trace[i * 16 + 1] = clock64();

    // =================================== gene phases ===================================
    double aR = 0, aRz = 0, aDu = 0, aDu2 = 0, aS0 = 0, aS1 = 0, aZ2 = 0, aU = 0, aAP = 0, aAPz = 0;
    {
      const double sb = S.shx[0], ma = S.zsh[2], sa = S.shx[1], inv_s2 = S.shx[3 + k];
      const double mual = lane < k ? S.zsh[5 + lane] : 0.0, sal = lane < k ? S.shx[3 + lane] : 0.0;
      for (int j = 0; j < nmy; ++j) {
        int tile, split, c_begin, n;
        item_geom(j, tile, split, c_begin, n);
        const int r_lo = split * p.gs, r_hi = (r_lo + p.gs < BVG_GENE_TILE) ? r_lo + p.gs : BVG_GENE_TILE;
        const uint32_t target = (uint32_t)p.nsplit * (uint32_t)(i + 1);
        struct GS { int64_t g, li; double zi, mu, om, hm, ho, A, eps; };
        // everything of a gene that does not depend on the other CTAs: this lane's parameter state and its next normal
        auto load = [&](int slot, GS &s) {
          const int rr = r_lo + slot;
          s.g = rr < r_hi ? (int64_t)tile * BVG_GENE_TILE + rr : m.G;
          s.li = 0; s.zi = 0; s.mu = 0; s.om = 0; s.hm = 0; s.ho = 0; s.A = 0; s.eps = 0;
          if (s.g < m.G) {
            const int64_t g = s.g;
            s.li = lane < k ? L.z_alpha + g * k + lane : (lane == k ? L.z_beta0 + g : L.amp + g);
            s.zi = has ? m.zeta[s.li] : ((dep && lane == k) ? m.depths[g] : 0.0);
            s.A = m.Aexp[g];
            if (has) { s.mu = m.mu[s.li]; s.om = m.omega[s.li]; s.hm = m.hist[s.li]; s.ho = m.hist[L.D + s.li]; }
            if (ahead && has) s.eps = (j == 0 && slot < PS_NPRE) ? S.nrm[slot * PS_KW + lane] : bvg_normal(m.seed, STREAM_GRAD, tg_next, global_index(g, lane));
          }
        };
        // the splits' partials of a gene (fixed order -> deterministic, k_gene's order): one pointer per array, bumped by a
        // constant stride (the indexed form spent ~40 % of the gene phase's instructions on 64-bit address arithmetic)
        auto fetch = [&](const GS &s, double &Pj, double &s0) {
          Pj = 0; s0 = 0;
          if (s.g >= m.G) return;  // warp-uniform
          const int64_t g = s.g;
          const float *pp = Pp + g * KP + (lane < KP ? lane : KP - 1);
          const double *sp_ = Sp + g * 2 + (lane & 1);
          const int64_t pstride = m.Gpad * KP, sstride = m.Gpad * 2;
          for (int sp0 = 0; sp0 < p.nsplit; sp0 += 8) {
            float pv[8];
            double sv[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const bool ok = sp0 + q < p.nsplit;
              pv[q] = ok ? __ldcg(pp) : 0.f;
              sv[q] = (ok && lane < 2) ? __ldcg(sp_) : 0.0;
              if (ok) { pp += pstride; sp_ += sstride; }
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              Pj += lane < KP ? (double)pv[q] : 0.0;
              s0 += sv[q];
            }
          }
        };
        auto process = [&](const GS &s, double Pj, double s0) {
          if (s.g >= m.G) return;  // warp-uniform
          const int64_t g = s.g;
          const double s1 = __shfl_sync(0xffffffffu, s0, 1);
          s0 = __shfl_sync(0xffffffffu, s0, 0);
          Pj *= inv_s2;
          const double R = __shfl_sync(0xffffffffu, Pj, k);  // ones column of Phi' -> R_g = sum_s r_sg
          const double zi = s.zi, zj = lane < k ? zi : 0.0;
          const double zb = __shfl_sync(0xffffffffu, zi, k), u = __shfl_sync(0xffffffffu, zi, k + 1);
          const double A = s.A, alpha = mual + sal * zj;
          const double aP = warp_sum(lane < k ? alpha * Pj : 0.0);
          const double z2 = warp_sum(zj * zj) + (dep ? 0.0 : zb * zb);
          const double AP = lane < k ? A * Pj : 0.0;
          const double du = u - ma;
          aAP += AP; aAPz += AP * zj;
          aR += R; aRz += R * zb; aDu += du; aDu2 += du * du; aS0 += s0; aS1 += s1; aZ2 += z2; aU += u;
          double gv = 0;
          if (lane < k) gv = sal * AP - zj;
          else if (lane == k) gv = sb * R - zb;
          else if (lane == k + 1) gv = 2.0 * A * aP - du / (sa * sa);  // jacobian on
          // mean-field ADVI update (advi_update of k_model.cu); a gene whose gradient is not finite keeps its parameters
          const bool fin = __all_sync(0xffffffffu, has ? isfinite(gv) : true);
          if (!fin && lane == 0) atomicAdd(&m.ctl->bad, 1);
          if (has) {
            double hm = s.hm, ho = s.ho;
            const double gm = fin ? gv : 0.0, go = fin ? gv * (zi - s.mu) + 1.0 : 0.0;
            if (it == 1) { hm += gm * gm; ho += go * go; }
            else { hm = 0.9 * hm + 0.1 * gm * gm; ho = 0.9 * ho + 0.1 * go * go; }
            m.hist[s.li] = hm;
            m.hist[L.D + s.li] = ho;
            const double mu_n = s.mu + es * gm / (1.0 + sqrt(hm)), om_n = s.om + es * go / (1.0 + sqrt(ho));
            m.mu[s.li] = mu_n;
            m.omega[s.li] = om_n;
            if (ahead) {  // next evaluation's draw of this parameter
              const double zn = mu_n + exp(om_n) * s.eps;
              m.zeta[s.li] = zn;
              if (lane == k + 1) {
                const double An = exp(2.0 * zn);  // approxGP.stan:27
                m.Aexp[g] = An;
                m.W[g * 32 + lane] = (float)An;
              } else {
                m.W[g * 32 + lane] = (float)zn;
              }
            }
          }
        };
        GS s0, s1;
        if (tr && i < 32 && j == 0) p.trace[i * 16 + 8] = clock64();
        load(warp, s0);
        if (tr && i < 32 && j == 0) p.trace[i * 16 + 9] = clock64();
        load(warp + PS_NW, s1);
        if (tr && i < 32 && j == 0) p.trace[i * 16 + 6] = clock64();
        if (lane == 0) wait_counter(p.tile_cnt + tile, target, p.timeout);
        __syncwarp();
        if (tr && i < 32 && j == 0) p.trace[i * 16 + 7] = clock64();
        double P0, P1, q0s, q1s;
        fetch(s0, P0, q0s);   // both genes' partials are in flight before either gene's arithmetic starts
        fetch(s1, P1, q1s);
        process(s0, P0, q0s);
        if (tr && i < 32 && j == 0) p.trace[i * 16 + 10] = clock64();
        process(s1, P1, q1s);
        if (tr && i < 32 && j == 0) p.trace[i * 16 + 11] = clock64();
        for (int slot = warp + 2 * PS_NW; r_lo + slot < r_hi; slot += PS_NW) {
          GS s2;
          double P2, q2s;
          load(slot, s2);
          fetch(s2, P2, q2s);
          process(s2, P2, q2s);
        }
      }
    }
    if (tr && i < 32) p.trace[i * 16 + 2] = clock64();

    // =================================== block row of the cross-gene sums ===================================
    {
      double *row = SC->wrow[warp];
      if (lane == 0) {
        row[S_R] = aR; row[S_RZ] = aRz; row[S_DU] = aDu; row[S_DU2] = aDu2; row[S_SSE] = aS0; row[S_Z2] = aZ2;
        row[S_U] = aU; row[S_L] = aS1;
        row[S_LG] = (LIK == BVG_NB && warp == TC_WARP_BASIS) ? hlg : 0.0;
        row[S_DG] = (LIK == BVG_NB && warp == TC_WARP_BASIS) ? hdg : 0.0;
      }
      if (lane < k) { row[S_AP + lane] = aAP; row[S_AP + k + lane] = aAPz; }
    }
    if (tr && i < 32) p.trace[i * 16 + 12] = clock64();
    __syncthreads();
    if (tr && i < 32) p.trace[i * 16 + 13] = clock64();
    // The block rows meet in FIXED-POINT accumulators: every CTA adds its row with three 64-bit integer reductions per sum
    // (value * 2^72 = l2 * 2^104 + l1 * 2^52 + l0 with signed 52-bit limbs: |value| < 2^84, resolution 2^-72, and thousands of
    // limbs cannot overflow a 64-bit word).  Integer addition commutes, so the totals are bit-reproducible whatever the order of
    // arrival -- like the fixed-order sum of 148 rows they replace -- and after the meeting a CTA reads 3 x (2k+10) words
    // instead of 148 rows (the reduction went from 4.4 k to ~1.7 k cycles at cfg2).  A sum that is not finite or out of that
    // range (a diverging step-size trial of adapt_eta) sets the poison word: then every CTA publishes its row as before and
    // the rows are added after a second meeting (the same answer as ever, just slower).  Three buffers: CTA 0 clears the next
    // one before it arrives.
    unsigned long long *const fxb = p.fx + (size_t)(i % 3) * PS_FX;
    if (tid < NS) {
      double s = 0;
#pragma unroll
      for (int w = 0; w < PS_NW; ++w) s += SC->wrow[w][tid];
      if (fabs(s) < 19342813113834066795298816.0 && !p.force_rows) {   // 2^84 (false for NaN)
        // signed limbs, truncated towards zero: every remainder is the exact low part of the value
        const double a2 = trunc(s * (1.0 / 4294967296.0));            // units of 2^32
        const double r1 = (s - a2 * 4294967296.0) * 1048576.0;        // |r1| < 2^52: units of 2^-20, exact
        const double a1 = trunc(r1);
        const unsigned long long l2 = (unsigned long long)(long long)a2, l1 = (unsigned long long)(long long)a1,
                                 l0 = (unsigned long long)(long long)((r1 - a1) * 4503599627370496.0);   // units of 2^-72, truncated
        asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(fxb + tid), "l"(l2) : "memory");
        asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(fxb + PS_GROW + tid), "l"(l1) : "memory");
        asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(fxb + 2 * PS_GROW + tid), "l"(l0) : "memory");
      } else {
        asm volatile("red.relaxed.gpu.global.or.b64 [%0], %1;" ::"l"(fxb + 3 * PS_GROW), "l"(1ull << tid) : "memory");
      }
    } else if (cta == 0 && tid >= 64 && tid < 64 + 3 * PS_GROW + 1) {
      p.fx[(size_t)((i + 1) % 3) * PS_FX + (tid - 64)] = 0ull;   // last read two iterations ago; next written after this meeting
    }
    // grid-wide arrival counter (all CTAs are co-resident: cooperative launch)
    __syncthreads();
    if (tr && i < 32) p.trace[i * 16 + 14] = clock64();
    if (tid == 0) {
      red_release_add(p.grid_cnt, 1u);
      if (tr && i < 32) p.trace[i * 16 + 15] = clock64();
    }
    // while the other CTAs arrive: the normals the NEXT gene phase and the next shared update will use (the gene phase of
    // this iteration has finished with S.nrm; buffer (i + 1) & 1 of the shared normals was last read one iteration ago)
    if (i + 2 < p.n_iter) draw_normals(tg0 + (uint32_t)i + 2u, (i + 1) & 1, tid, TC_THREADS);
    if (tid == 0) wait_counter(p.grid_cnt, (uint32_t)ncta * (uint32_t)(i + 1), p.timeout);
    __syncthreads();
    if (tr && i < 32) p.trace[i * 16 + 3] = clock64();

    // =================================== finish, replicated in every CTA ===================================
    if (ahead && nmy > 0 && warp >= TC_WARP_EPI0 && warp < TC_WARP_EPI0 + 8) {
      // the next pass's draw rows are final (every CTA has passed its gene phases): fetch them under the reduction
      const int h = (warp - TC_WARP_EPI0) >> 2;
      const int64_t g = (int64_t)tile0 * BVG_GENE_TILE + 32 * (warp & 3) + lane;
      const float4 *pw = reinterpret_cast<const float4 *>(a.W + g * TC_KW + 16 * h);
#pragma unroll
      for (int q = 0; q < 4; ++q) wpre[q] = __ldcg(pw + q);
      wAg = __ldcg(a.W + g * TC_KW + k + 1);
    }
    const bool reducer = m.peer.world <= 1 || cta == 0 || p.xchg_all;  // multi-GPU: only CTA 0 needs the local sums (it exchanges and publishes the totals)
    unsigned long long L2 = 0, L1 = 0, L0 = 0, pw = 0;
    if (tid < NS) {   // limbs and poison word in one round trip
      L2 = __ldcg(fxb + tid); L1 = __ldcg(fxb + PS_GROW + tid); L0 = __ldcg(fxb + 2 * PS_GROW + tid);
      pw = __ldcg(fxb + 3 * PS_GROW);
    }
    const int poison = __syncthreads_or(pw != 0ull);   // the same word in every CTA: a uniform decision
    if (!poison) {
      if (reducer && tid < NS) {
        double t = (double)(long long)L0 * (1.0 / 4722366482869645213696.0) + (double)(long long)L1 * (1.0 / 1048576.0);   // 2^-72, 2^-20
        t += (double)(long long)L2 * 4294967296.0;
        if (LIK == BVG_NB && tid == S_LG) t -= m.lgam_y1;
        S.tot[tid] = t;
      }
    } else {
      // out of the accumulators' range, or not finite: the rows themselves meet (round 2's first form of this reduction)
      double *const grow = p.rows;
      if (tid < NS) {
        double s = 0;
#pragma unroll
        for (int w = 0; w < PS_NW; ++w) s += SC->wrow[w][tid];
        grow[(size_t)cta * PS_GROW + tid] = s;
      }
      __syncthreads();
      if (tid == 0) {
        red_release_add(p.grid_cnt2, 1u);
        wait_counter(p.grid_cnt2, (uint32_t)ncta * (uint32_t)(nfallback + 1), p.timeout);
      }
      ++nfallback;
      __syncthreads();
      if (reducer) {
        const int c = tid & 63, grp = tid >> 6;
        if (c < NS) {
          double acc = 0;
          for (int b = grp; b < ncta; b += 6) acc += __ldcg(grow + (size_t)b * PS_GROW + c);
          SC->wrow[grp][c] = acc;   // every thread of this CTA has read its own rows (barrier above)
        }
      }
      __syncthreads();
      if (reducer && tid < NS) {
        double t = 0;
#pragma unroll
        for (int q = 0; q < 6; ++q) t += SC->wrow[q][tid];
        if (LIK == BVG_NB && tid == S_LG) t -= m.lgam_y1;
        S.tot[tid] = t;
      }
    }
    __syncthreads();
    if (m.peer.world > 1) {
      // multi-GPU: the local sums of every gene shard meet here (bvg_internal.cuh: 8-byte cells {32 data bits, sequence
      // number}).  ONE CTA talks to the peers: it pushes this rank's sums, polls the local exchange buffer, sums in rank
      // order (bit-identical totals on every rank) and publishes them to the other CTAs behind a flag.  (All 144 CTAs polling
      // the 700 cells themselves cost ~10 us per evaluation at 8 GPUs: the polling storm sat on the very L2 lines the
      // NVLink writes had to land in.)
      double *const xt = p.xtot + (size_t)(i & 1) * PS_GROW;
      if (cta == 0 || p.xchg_all) {
        const PeerCtx &pc = m.peer;
        ++xseq;
        const size_t per_rank = 2 * BVG_PEER_CAP, slot = (size_t)(xseq & 1u) * pc.world * per_rank;
        if (cta == 0) {
          // one thread per (peer, cell): all remote stores of the exchange leave in one instruction per thread
          for (int q = tid; q < 2 * NS * (pc.world - 1); q += TC_THREADS) {
            const int ri = q / (2 * NS), cq = q - ri * 2 * NS, r = ri < pc.rank ? ri : ri + 1;
            const double v = S.tot[cq >> 1];
            const uint32_t hv = (cq & 1) ? (uint32_t)__double2hiint(v) : (uint32_t)__double2loint(v);
            uint2 *dst = pc.cells[r] + slot + (size_t)pc.rank * per_rank + cq;
            asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(dst), "r"(hv), "r"(xseq) : "memory");
          }
        }
        double *scratch = &SC->wrow[0][0];  // (world - 1) * NS <= 7 * 56 doubles
        const uint2 *mine = pc.cells[pc.rank] + slot;
        const long long t0 = clock64();
        for (int t = tid; t < NS * (pc.world - 1); t += TC_THREADS) {
          const int ri = t / NS, cc = t - ri * NS, r = ri < pc.rank ? ri : ri + 1;
          scratch[t] = peer_dead ? __longlong_as_double(0x7ff8000000000000LL) : peer_poll(pc, mine + (size_t)r * per_rank + 2 * cc, xseq, t0);
        }
        __syncthreads();
        double sx = 0;
        if (tid < NS)
          for (int r = 0; r < pc.world; ++r) sx += r == pc.rank ? S.tot[tid] : scratch[(r < pc.rank ? r : r - 1) * NS + tid];  // rank order
        if (!peer_dead && *(volatile uint32_t *)pc.fail) peer_dead = true;  // a peer never arrived: stop polling, finish with NaN
        __syncthreads();
        if (tid < NS) { S.tot[tid] = sx; if (!p.xchg_all) xt[tid] = sx; }
        __syncthreads();
        if (tid == 0 && !p.xchg_all) red_release_add(p.xflag, 1u);
      } else {
        if (tid == 0) wait_counter(p.xflag, (uint32_t)(i + 1), p.timeout);
        __syncthreads();
        if (tid < NS) S.tot[tid] = __ldcg(xt + tid);
        __syncthreads();
      }
    }
    if (tr && i < 32) p.trace[i * 16 + 4] = clock64();
    // warps 2, 11: gradient and mean-field update of the 2k+5 gene-shared parameters (finish_body of k_model.cu, jacobian on);
    // every other warp meanwhile draws the normals the NEXT gene phase and the next shared update will use
    double zn = 0;
    const bool upd_warp = warp == TC_WARP_MMA2 || warp == TC_WARP_BASIS;
    const int ut = warp == TC_WARP_MMA2 ? lane : 32 + lane;
    if (upd_warp) {
      if (ut < NSH) {
        const double *tot = S.tot;
        const double mb = S.zsh[0], sbb = S.shx[0], maa = S.zsh[2], saa = S.shx[1], tl = S.shx[2];
        const double Gt = (double)m.G_total, N = (double)m.M * Gt;
        double gv;
        if (ut == 0) gv = tot[S_R] - mb / 4.0;
        else if (ut == 1) gv = dep ? tot[S_RZ] - sbb / 4.0 : sbb * (tot[S_RZ] - sbb) + 1.0;
        else if (ut == 2) gv = tot[S_DU] / (saa * saa) - maa / 4.0;
        else if (ut == 3) gv = -Gt + tot[S_DU2] / (saa * saa) - saa * saa + 1.0;
        else if (ut == 4) {
          if (LIK == BVG_GAUSSIAN) gv = -N + tot[S_SSE] / (tl * tl) - (dep ? tl * tl / 4.0 : tl * tl) + 1.0;
          else gv = tl * (tot[S_DG] - tot[S_L] - tot[S_R] / tl) - 0.5 * tl * tl / (1.0 + tl * tl / 4.0) + 1.0;
        } else if (ut < 5 + k) gv = tot[S_AP + (ut - 5)] - S.zsh[ut] / 4.0;
        else {
          const double s = S.shx[3 + ut - 5 - k];
          gv = s * (tot[S_AP + k + (ut - 5 - k)] - s) + 1.0;
        }
        const bool fin = isfinite(gv);
        if (!fin && cta == 0) atomicAdd(&m.ctl->bad, 1);
        double p_mu = S.st[0][ut], p_om = S.st[1][ut], p_hm = S.st[2][ut], p_ho = S.st[3][ut];
        const double gm = fin ? gv : 0.0, go = fin ? gv * (S.zsh[ut] - p_mu) + 1.0 : 0.0;
        if (it == 1) { p_hm += gm * gm; p_ho += go * go; }
        else { p_hm = 0.9 * p_hm + 0.1 * gm * gm; p_ho = 0.9 * p_ho + 0.1 * go * go; }
        p_mu += es * gm / (1.0 + sqrt(p_hm));
        p_om += es * go / (1.0 + sqrt(p_ho));
        S.st[0][ut] = p_mu; S.st[1][ut] = p_om; S.st[2][ut] = p_hm; S.st[3][ut] = p_ho;
        if (ahead) zn = p_mu + exp(p_om) * S.nrm_sh[i & 1][ut];
      }
    }
    __syncthreads();  // every thread has read this evaluation's shared draw
    if (ahead && upd_warp && ut < NSH) {
      S.zsh[ut] = zn;
      if (ut == 0) S.U[64] = (float)zn;
      else if (ut == 1) { const double e = dep ? zn : exp(zn); S.shx[0] = e; S.U[65] = (float)e; }  // sigma_beta0 | beta1
      else if (ut == 3) S.shx[1] = exp(zn);
      else if (ut == 4) { const double e = exp(zn); S.shx[2] = e; S.shx[3 + k] = LIK == BVG_GAUSSIAN ? 1.0 / (e * e) : 1.0; }
      else if (ut >= 5 + k) { const double e = exp(zn); S.shx[3 + ut - 5 - k] = e; S.U[32 + ut - 5 - k] = (float)e; }
      else if (ut >= 5) S.U[ut - 5] = (float)zn;
    }
    // the scratch stage goes back to the TMA ring: order this CTA's generic-proxy accesses before the async-proxy refill
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tr && i < 32) p.trace[i * 16 + 5] = clock64();
  }

  // the replica of CTA 0 goes back to the model (all replicas are identical)
  if (cta == 0) {
    if (tid < NSH) {
      const int64_t li = shared_index(L, k, tid);
      m.zeta[li] = S.zsh[tid];
      m.mu[li] = S.st[0][tid]; m.omega[li] = S.st[1][tid]; m.hist[li] = S.st[2][tid]; m.hist[L.D + li] = S.st[3][tid];
    }
    if (tid < 4 + k) m.shx[tid] = S.shx[tid];
    if (tid < 96) m.U[tid] = S.U[tid];
    if (tid == 0) {
      m.ctl->it = it0 + p.n_iter;
      m.ctl->t_grad = tg0 + (uint32_t)p.n_iter;
      m.ctl->es = eta / sqrt((double)(it0 + p.n_iter));
      if (m.peer.world > 1) *m.peer.seq = xseq;
    }
  }
  tc_teardown_cta(x);
}

// ------------------------------------------------------------------------------------------------
// k_elbo_persist: all S forward-only draws of ONE ELBO estimate (stan::variational::advi::calc_ELBO [U]: the mean of
// log_prob<propto = false, jacobian = true> over S draws of q) in one cooperative launch -- SURVEY.md 2.1's "K5".
// The draws do not depend on each other, so there is NO barrier between them: a CTA streams its (gene tile, spot range)
// item(s) through the forward-only pass once per draw -- rings, barriers and TMEM keep turning -- and leaves one row of
// partial sums per draw; the only grid-wide synchronisation is the one at the end, after which one warp per draw adds the
// rows in a fixed order and forms the draw's log density.  The draws themselves (zeta, exponentials, fp32 rows W / U of all
// S draws) come from ONE batched k_prep launch.  While the epilogue warps work, the otherwise idle backward-MMA warp adds
// the prior terms of the CTA's own genes, and (nb) the basis producer's idle lanes the lgamma terms of their distinct counts.
struct ElboArgs {
  TcArgs a;
  DevModel m;                 // zeta / shx / Aexp / W / U point at slice 0 of the batch buffers
  int S, nitems, nsplit, gs;
  int loc, rank_id;           // gene-sharded: local share of every draw (the host all-reduces the S values)
  int pre_y, pre_b;
  uint32_t *grid_cnt;         // [1], zeroed before the launch
  double *erows;              // [S][gridDim.x][8]: SSE | sum L | sum z^2 | sum u | sum (u - mu_amplitude)^2 | lgamma terms | - | -
  double *lps;                // [S] out
  long long timeout;
};

template <int LIK>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_elbo_persist(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmF1,
               const __grid_constant__ CUtensorMap tmF2, const __grid_constant__ CUtensorMap tmT1,
               const __grid_constant__ CUtensorMap tmT2, const __grid_constant__ ElboArgs p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ double s_wsum[2][8][2];   // [draw parity][epilogue warp][SSE | L]
  __shared__ double s_tot[PS_NW][PS_ROW];
  const TcArgs &a = p.a;
  const DevModel &m = p.m;
  TcCtx x = tc_ctx_make(smem_raw, p.timeout);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cta = blockIdx.x, ncta = gridDim.x;
  const int k = m.k;
  const bool dep = m.depth != 0;
  const Layout &L = m.L;
  const int64_t D = L.D;
  tc_setup_cta(x);
  const int nmy = cta < p.nitems ? (p.nitems - cta + ncta - 1) / ncta : 0;
  auto item_geom = [&](int j, int &tile, int &split, int &c_begin, int &n) {
    const int id = cta + j * ncta;
    tile = id / p.nsplit;
    split = id - tile * p.nsplit;
    c_begin = split * a.cps;
    n = a.nchunks - c_begin;
    n = n < a.cps ? n : a.cps;
    if (n < 0) n = 0;
  };
  uint32_t q0 = 0, pass = 0;
  int preY = 0, preB = 0;
  for (int s = 0; s < p.S; ++s) {
    const double *zt = m.zeta + (size_t)s * D, *shx = m.shx + (size_t)s * (4 + k);
    const float *Ws = m.W + (size_t)s * m.Gpad * 32, *Us = m.U + (size_t)s * 96;
    double *erow = p.erows + ((size_t)s * ncta + cta) * 8;
    double t0 = 0.0, t1 = 0.0;                 // epilogue threads: this draw's sums over the CTA's items
    double aZ2 = 0, aU = 0, aDU2 = 0;          // warp 2: prior terms of the CTA's own genes
    double hlg = 0.0;
    for (int j = 0; j < nmy; ++j) {
      int tile, split, c_begin, n;
      item_geom(j, tile, split, c_begin, n);
      const bool has_next = j + 1 < nmy || s + 1 < p.S;
      int ntile = 0, nsplit_ = 0, nc_begin = 0, nn = 0;
      if (has_next) item_geom(j + 1 < nmy ? j + 1 : 0, ntile, nsplit_, nc_begin, nn);
      if (warp == TC_WARP_Y) {
        const int pre = has_next ? (nn < p.pre_y ? nn : p.pre_y) : 0;
        if (lane == 0) {
          tc_role_produce_y(x, &tmY, tile, a.nch32, c_begin, preY, n, q0);
          tc_role_produce_y(x, &tmY, ntile, a.nch32, nc_begin, 0, pre, q0 + (uint32_t)n);
        }
        preY = pre;
        __syncwarp();
      } else if (warp == TC_WARP_BASIS) {
        const int pre = has_next ? (nn < p.pre_b ? nn : p.pre_b) : 0;
        if (lane == 0) {
          tc_role_produce_basis<false>(x, &tmF1, &tmF2, &tmT1, &tmT2, c_begin, preB, n, q0);
          tc_role_produce_basis<false>(x, &tmF1, &tmF2, &tmT1, &tmT2, nc_begin, 0, pre, q0 + (uint32_t)n);
        }
        preB = pre;
        double lg = 0.0;
        if (LIK == BVG_NB && j == 0 && lane > 0) {
          const double phi = shx[2], lg0 = lgamma(phi);
          for (int e = cta * 31 + lane - 1; e < a.nhist; e += ncta * 31) lg += a.hist_cnt[e] * (lgamma(a.hist_val[e] + phi) - lg0);
        }
        __syncwarp();
        if (LIK == BVG_NB && j == 0) hlg = warp_sum(lg);
      } else if (warp == TC_WARP_MMA1) {
        tc_role_mma1<false>(x, n, a.nk1, q0, pass, nullptr);
      } else if (warp == TC_WARP_MMA2) {
        // no backward MMA in a forward-only pass: this warp adds the prior terms of the genes this CTA owns.  Their
        // parameters are three contiguous runs of the draw (z_alpha_t[, g], z_beta0[g], amplitude[g] for g in [g_lo, g_hi)):
        // every lane keeps 16 loads in flight (a gene-by-gene loop paid one HBM round trip per gene -- 22 per draw, 17 us --
        // and made this warp, not the pass, the critical path of the kernel: 2.4 ms for 150 draws)
        const int r_lo = split * p.gs, r_hi = (r_lo + p.gs < BVG_GENE_TILE) ? r_lo + p.gs : BVG_GENE_TILE;
        const int64_t g_lo = (int64_t)tile * BVG_GENE_TILE + r_lo;
        int64_t g_hi = (int64_t)tile * BVG_GENE_TILE + r_hi;
        if (g_hi > m.G) g_hi = m.G;
        const int ng = g_hi > g_lo ? (int)(g_hi - g_lo) : 0;
        const double ma = zt[L.mu_amp];
        double z2 = 0, su = 0, sdu2 = 0;
        auto run = [&](const double *base, int cnt, int what) {   // what: 0 = sum v^2, 1 = sums of v and (v - ma)^2
          for (int i0 = 0; i0 < cnt; i0 += 32 * 16) {
            double v[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) { const int i = i0 + 32 * q + lane; v[q] = __ldcg(base + (i < cnt ? i : 0)); }
#pragma unroll
            for (int q = 0; q < 16; ++q) {
              const bool ok = i0 + 32 * q + lane < cnt;
              if (what == 0) z2 += ok ? v[q] * v[q] : 0.0;
              else { su += ok ? v[q] : 0.0; const double du = v[q] - ma; sdu2 += ok ? du * du : 0.0; }
            }
          }
        };
        if (ng > 0) {
          run(zt + L.z_alpha + g_lo * k, ng * k, 0);
          if (!dep) run(zt + L.z_beta0 + g_lo, ng, 0);
          run(zt + L.amp + g_lo, ng, 1);
        }
        aZ2 += warp_sum(z2); aU += warp_sum(su); aDU2 += warp_sum(sdu2);
      } else {
        const int e = warp - TC_WARP_EPI0, qd = warp & 3, h = e >> 2;
        const int r = 32 * qd + lane;
        const uint32_t trow = x.tmem + ((uint32_t)(32 * qd) << 16);
        const int64_t g = (int64_t)tile * BVG_GENE_TILE + r;
        {
          float cv[16];
          const float4 *pw = reinterpret_cast<const float4 *>(Ws + g * TC_KW + 16 * h);
          const float4 *pm = reinterpret_cast<const float4 *>(Us + 16 * h), *ps = reinterpret_cast<const float4 *>(Us + 32 + 16 * h);
          const float Ag = __ldg(Ws + g * TC_KW + k + 1), mb = __ldg(Us + 64), sb = __ldg(Us + 65);
          float w[16], um[16], us[16];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 xx = __ldg(pw + i), y = __ldg(pm + i), z = __ldg(ps + i);
            w[4 * i] = xx.x; w[4 * i + 1] = xx.y; w[4 * i + 2] = xx.z; w[4 * i + 3] = xx.w;
            um[4 * i] = y.x; um[4 * i + 1] = y.y; um[4 * i + 2] = y.z; um[4 * i + 3] = y.w;
            us[4 * i] = z.x; us[4 * i + 1] = z.y; us[4 * i + 2] = z.z; us[4 * i + 3] = z.w;
          }
          tc_epi_form_coeff(k, h, g < a.G, Ag, mb, sb, w, um, us, cv);
          tc_epi_store_coeff(x, trow, h, cv);
        }
        double sum0 = 0.0, sum1 = 0.0;
        float lt = 0.f, phi = 0.f;
        if (LIK == BVG_NB) { lt = (float)zt[L.tail]; phi = __expf(lt); }
        tc_role_epilogue<LIK, false>(x, trow, r, h, c_begin, n, q0, a.M, lt, phi, sum0, sum1, nullptr);
        if (g < a.G) { t0 += sum0; t1 += sum1; }
      }
      q0 += (uint32_t)n;
      ++pass;
    }
    // ---- this CTA's row of the draw (no CTA-wide barrier: only the eight epilogue warps meet) ----
    if (warp >= TC_WARP_EPI0 && warp < TC_WARP_EPI0 + 8) {
      const int e = warp - TC_WARP_EPI0;
      t0 = warp_sum(t0);
      if (LIK == BVG_NB) t1 = warp_sum(t1);
      if (lane == 0) { s_wsum[s & 1][e][0] = t0; s_wsum[s & 1][e][1] = t1; }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (e == 0 && lane == 0) {
        double b0 = 0, b1 = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) { b0 += s_wsum[s & 1][w][0]; b1 += s_wsum[s & 1][w][1]; }
        erow[0] = b0; erow[1] = b1;
      }
    } else if (warp == TC_WARP_MMA2) {
      if (lane == 0) { erow[2] = aZ2; erow[3] = aU; erow[4] = aDU2; erow[6] = 0.0; erow[7] = 0.0; }
    } else if (warp == TC_WARP_BASIS) {
      if (lane == 0) erow[5] = hlg;
    }
  }
  // ---- the one grid-wide meeting, then one warp per draw: rows -> sums -> log density ----
  __syncthreads();
  if (tid == 0) {
    red_release_add(p.grid_cnt, 1u);
    wait_counter(p.grid_cnt, (uint32_t)ncta, p.timeout);
  }
  __syncthreads();
  for (int s = cta * PS_NW + warp; s < p.S; s += ncta * PS_NW) {
    const int c = lane & 7, rg = lane >> 3;   // 8 columns x 4 row groups
    double acc = 0;
    for (int b = rg; b < ncta; b += 4) acc += __ldcg(p.erows + ((size_t)s * ncta + b) * 8 + c);
    const double a1 = __shfl_sync(0xffffffffu, acc, c + 8), a2 = __shfl_sync(0xffffffffu, acc, c + 16), a3 = __shfl_sync(0xffffffffu, acc, c + 24);
    acc = ((acc + a1) + a2) + a3;   // fixed order
    double *tot = s_tot[warp];
    if (lane < 8) {
      const int slot = lane == 0 ? S_SSE : lane == 1 ? S_L : lane == 2 ? S_Z2 : lane == 3 ? S_U : lane == 4 ? S_DU2 : lane == 5 ? S_LG : -1;
      if (slot >= 0) tot[slot] = (LIK == BVG_NB && slot == S_LG) ? acc - m.lgam_y1 : acc;
    }
    __syncwarp();
    const double lp = bvg_lp_from_sums(m, tot, m.zeta + (size_t)s * D, m.shx + (size_t)s * (4 + k), 1, 0, p.loc != 0,
                                       (p.loc && p.rank_id != 0) ? 0.0 : 1.0, lane);
    if (lane == 0) p.lps[s] = lp;
    __syncwarp();
  }
  tc_teardown_cta(x);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct PersistBuf {
  uint32_t *cnt;    // [ntiles] tile counters, grid counter, exchange flag
  double *rows;     // [2][ncta][PS_GROW] block rows, then [2][PS_GROW] exchanged totals
  unsigned long long *fx;  // [3][PS_FX] fixed-point accumulators of k_advi_persist's grid meeting
  long long *trace; // [32][8]
  double *erows;    // ELBO mode: [S][ncta][8]
  int erows_cap;
  int ntiles, ncta;
  bool attr_set;
};

static int g_persist_trace = 0;

void persist_destroy(bvg_fit *f) {
  TcState *t = (TcState *)f->tc;
  if (!t || !t->persist) return;
  PersistBuf *b = (PersistBuf *)t->persist;
  cudaFree(b->cnt);
  cudaFree(b->rows);
  cudaFree(b->fx);
  cudaFree(b->trace);
  cudaFree(b->erows);
  delete b;
  t->persist = nullptr;
}

// the persistent kernel serves the fused tcgen05 engine (k + 1 <= 24), single GPU or the in-kernel peer exchange
bool persist_ok(bvg_fit *f) {
  const DevModel &m = f->m;
  if (f->engine != BVG_ENGINE_TCGEN05 || f->big_tc || !f->tc || f->no_persist) return false;
  if (m.k + 2 > PS_KW || bvg_ns(m.k) > PS_ROW || 2 * m.k + 5 >= PS_NSH) return false;
  if (f->comm_mode == 1) return false;  // NCCL transport: the all-reduce is a separate launch
  // many (tile, split) items per CTA: the gene phases of a CTA's items run one after the other and only the first item's
  // normals are drawn ahead -- measured at config 4 on ONE GPU (1,027 items on 148 SMs): 436 us per evaluation against 405 us
  // for the two-launch chain; at <= 2 items per CTA (config 4 on 2+ GPUs) the persistent kernel wins (206 against 217 us)
  int nsm = 148;
  if (cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, f->cfg.device) != cudaSuccess) { cudaGetLastError(); nsm = 148; }
  if ((int64_t)(m.Gpad / BVG_GENE_TILE) * m.nsplit > (int64_t)4 * nsm) return false;
  return true;
}

int launch_advi_persist(bvg_fit *f, int n_iter) {
  const DevModel &m = f->m;
  TcState *t = (TcState *)f->tc;
  int nsm = 148, coop = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, f->cfg.device));
  CUDA_TRY(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, f->cfg.device));
  if (!coop) { bvg_set_error("device does not support cooperative launches"); return BVG_ERR_STATE; }
  const int ntiles = (int)(m.Gpad / BVG_GENE_TILE), nitems = ntiles * m.nsplit;
  const int ncta = std::min(nitems, std::min(nsm, 148));   // PS_NROWS block rows per parity
  PersistBuf *b = (PersistBuf *)t->persist;
  if (b && (b->ntiles != ntiles || b->ncta != ncta)) { persist_destroy(f); b = nullptr; }
  if (!b) {
    b = new PersistBuf();
    memset(b, 0, sizeof(*b));
    t->persist = b;
    b->ntiles = ntiles; b->ncta = ncta;
    CUDA_TRY(cudaMalloc(&b->cnt, sizeof(uint32_t) * (ntiles + 4)));
    CUDA_TRY(cudaMalloc(&b->rows, sizeof(double) * 2 * (PS_NROWS + 1) * PS_GROW));
    CUDA_TRY(cudaMemsetAsync(b->rows, 0, sizeof(double) * 2 * (PS_NROWS + 1) * PS_GROW, f->stream));
    CUDA_TRY(cudaMalloc(&b->trace, sizeof(long long) * (32 * 16 + 32 * 8 + 32 * 4)));
    CUDA_TRY(cudaFuncSetAttribute(k_advi_persist<BVG_GAUSSIAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    CUDA_TRY(cudaFuncSetAttribute(k_advi_persist<BVG_NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    CUDA_TRY(cudaFuncSetAttribute(k_elbo_persist<BVG_GAUSSIAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    CUDA_TRY(cudaFuncSetAttribute(k_elbo_persist<BVG_NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    b->attr_set = true;
    int occ = 0;
    if (m.lik == BVG_GAUSSIAN) CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_advi_persist<BVG_GAUSSIAN>, TC_THREADS, TC_SMEM_BYTES));
    else CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_advi_persist<BVG_NB>, TC_THREADS, TC_SMEM_BYTES));
    if (occ < 1) { bvg_set_error("persistent ADVI kernel does not fit an SM"); return BVG_ERR_STATE; }
  }
  CUDA_TRY(cudaMemsetAsync(b->cnt, 0, sizeof(uint32_t) * (ntiles + 4), f->stream));
  if (!b->fx) CUDA_TRY(cudaMalloc(&b->fx, sizeof(unsigned long long) * 3 * PS_FX));
  CUDA_TRY(cudaMemsetAsync(b->fx, 0, sizeof(unsigned long long) * 3 * PS_FX, f->stream));
  PersistArgs p;
  memset(&p, 0, sizeof(p));
  TcArgs &a = p.a;
  a.Ppart = (float *)m.Ppart; a.Spart = (double *)m.Spart;
  a.W = m.W; a.U = m.U; a.k = m.k;
  a.hist_val = m.hist_val; a.hist_cnt = m.hist_cnt; a.nhist = m.nhist;
  a.hterm = nullptr;
  a.nk1 = (m.k + 1 + 15) / 16;
  a.nch32 = (int)(f->ldY / 32);
  a.M = m.M; a.G = m.G; a.Gpad = m.Gpad; a.KP = m.KP;
  a.nchunks = (int)((m.M + TC_CHUNK - 1) / TC_CHUNK);
  a.cps = (a.nchunks + m.nsplit - 1) / m.nsplit;
  a.zeta = m.zeta; a.tail_idx = m.L.tail;
  p.m = m;
  p.n_iter = n_iter; p.nitems = nitems; p.nsplit = m.nsplit;
  p.gs = (BVG_GENE_TILE + m.nsplit - 1) / m.nsplit;
  p.pre_y = TC_NY; p.pre_b = TC_NB - 1;
  if (const char *e = getenv("BVG_PERSIST_PRE")) { int y = 0, bb = 0; if (sscanf(e, "%d,%d", &y, &bb) == 2) { p.pre_y = std::max(0, std::min(y, TC_NY)); p.pre_b = std::max(0, std::min(bb, TC_NB - 1)); } }
  { const char *e = getenv("BVG_XCHG_ALL"); p.xchg_all = !(e && e[0] == '0'); }   // measured at 8 GPUs (profiles/README.md): every CTA polling is the faster form
  p.tile_cnt = b->cnt; p.grid_cnt = b->cnt + ntiles;
  p.rows = b->rows;
  p.fx = b->fx;
  p.grid_cnt2 = b->cnt + ntiles + 2;
  { const char *e = getenv("BVG_PERSIST_FORCE_ROWS"); p.force_rows = e && e[0] == '1'; }
  p.xtot = b->rows + (size_t)2 * PS_NROWS * PS_GROW;
  p.xflag = b->cnt + ntiles + 1;
  p.timeout = 8000000000LL;  // ~4 s: longer than any legitimate wait inside one GPU
  if (m.peer.world > 1 && m.peer.timeout > p.timeout) p.timeout = 2 * m.peer.timeout;  // the rings idle while a peer is awaited
  p.trace = g_persist_trace ? b->trace : nullptr;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)ncta); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = TC_SMEM_BYTES; cfg.stream = f->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  if (m.lik == BVG_GAUSSIAN) CUDA_TRY(cudaLaunchKernelEx(&cfg, k_advi_persist<BVG_GAUSSIAN>, t->tmY, t->tmF1, t->tmF2, t->tmT1, t->tmT2, p));
  else CUDA_TRY(cudaLaunchKernelEx(&cfg, k_advi_persist<BVG_NB>, t->tmY, t->tmF1, t->tmF2, t->tmT1, t->tmT2, p));
  f->st.kernel_launches++;
  return BVG_OK;
}

// All S draws of an ELBO estimate in one cooperative launch.  The model descriptor's zeta / shx / Aexp / W / U must point at
// slice 0 of the batch buffers filled by launch_prep_batch(PREP_ELBO, S); d_lps receives S log densities (local shares when
// the fit is gene-sharded and m.elbo_local is set).
int launch_elbo_persist(bvg_fit *f, int S, double *d_lps) {
  const DevModel &m = f->m;
  TcState *t = (TcState *)f->tc;
  int nsm = 148, coop = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, f->cfg.device));
  CUDA_TRY(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, f->cfg.device));
  if (!coop) { bvg_set_error("device does not support cooperative launches"); return BVG_ERR_STATE; }
  const int ntiles = (int)(m.Gpad / BVG_GENE_TILE), nitems = ntiles * m.nsplit;
  const int ncta = std::min(nitems, std::min(nsm, 148));   // PS_NROWS block rows per parity
  PersistBuf *b = (PersistBuf *)t->persist;
  if (b && (b->ntiles != ntiles || b->ncta != ncta)) { persist_destroy(f); b = nullptr; }
  if (!b) {
    b = new PersistBuf();
    memset(b, 0, sizeof(*b));
    t->persist = b;
    b->ntiles = ntiles; b->ncta = ncta;
    CUDA_TRY(cudaMalloc(&b->cnt, sizeof(uint32_t) * (ntiles + 4)));
    CUDA_TRY(cudaMalloc(&b->rows, sizeof(double) * 2 * (PS_NROWS + 1) * PS_GROW));
    CUDA_TRY(cudaMemsetAsync(b->rows, 0, sizeof(double) * 2 * (PS_NROWS + 1) * PS_GROW, f->stream));
    CUDA_TRY(cudaMalloc(&b->trace, sizeof(long long) * (32 * 16 + 32 * 8 + 32 * 4)));
  }
  if (!b->attr_set) {
    CUDA_TRY(cudaFuncSetAttribute(k_advi_persist<BVG_GAUSSIAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    CUDA_TRY(cudaFuncSetAttribute(k_advi_persist<BVG_NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    CUDA_TRY(cudaFuncSetAttribute(k_elbo_persist<BVG_GAUSSIAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    CUDA_TRY(cudaFuncSetAttribute(k_elbo_persist<BVG_NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    b->attr_set = true;
  }
  if (b->erows_cap < S) {
    cudaFree(b->erows);
    b->erows = nullptr; b->erows_cap = 0;
    CUDA_TRY(cudaMalloc(&b->erows, sizeof(double) * (size_t)S * ncta * 8));
    b->erows_cap = S;
  }
  CUDA_TRY(cudaMemsetAsync(b->cnt, 0, sizeof(uint32_t) * (ntiles + 4), f->stream));
  ElboArgs p;
  memset(&p, 0, sizeof(p));
  TcArgs &a = p.a;
  a.k = m.k;
  a.hist_val = m.hist_val; a.hist_cnt = m.hist_cnt; a.nhist = m.nhist;
  a.nk1 = (m.k + 1 + 15) / 16;
  a.nch32 = (int)(f->ldY / 32);
  a.M = m.M; a.G = m.G; a.Gpad = m.Gpad; a.KP = m.KP;
  a.nchunks = (int)((m.M + TC_CHUNK - 1) / TC_CHUNK);
  a.cps = (a.nchunks + m.nsplit - 1) / m.nsplit;
  p.m = m;
  p.S = S; p.nitems = nitems; p.nsplit = m.nsplit;
  p.gs = (BVG_GENE_TILE + m.nsplit - 1) / m.nsplit;
  p.loc = m.elbo_local; p.rank_id = m.rank_id;
  p.pre_y = TC_NY; p.pre_b = TC_NB;
  p.grid_cnt = b->cnt + ntiles;
  p.erows = b->erows;
  p.lps = d_lps;
  p.timeout = 8000000000LL;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)ncta); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = TC_SMEM_BYTES; cfg.stream = f->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  if (m.lik == BVG_GAUSSIAN) CUDA_TRY(cudaLaunchKernelEx(&cfg, k_elbo_persist<BVG_GAUSSIAN>, t->tmY, t->tmF1, t->tmF2, t->tmT1, t->tmT2, p));
  else CUDA_TRY(cudaLaunchKernelEx(&cfg, k_elbo_persist<BVG_NB>, t->tmY, t->tmF1, t->tmF2, t->tmT1, t->tmT2, p));
  f->st.kernel_launches++;
  return BVG_OK;
}

// diagnostics: clock64 stamps of CTA 0 for the first iterations of one stretch: [iteration][8] =
// start, passes done, gene phases done, grid barrier passed, sums reduced (+ exchanged), shared update done, normals drawn
// (before the tile wait), tile counter reached; rows 32.. = per-chunk stamps of the pass of iteration 8 (tc_roles.cuh
// layout; row 63: coefficient rows stored, chunk loop done, D2 read, partials written)
extern "C" int bvg_debug_persist_trace(bvg_fit *f, int n_iter, long long *out, int cap_iters) {
  if (!f || !out || !persist_ok(f)) { bvg_set_error("persistent ADVI kernel not active"); return BVG_ERR_STATE; }
  CUDA_TRY(cudaSetDevice(f->cfg.device));
  BVG_TRY(launch_prep(f, PREP_GRAD));
  g_persist_trace = 1;
  PersistBuf *b0 = (PersistBuf *)((TcState *)f->tc)->persist;
  if (b0) CUDA_TRY(cudaMemsetAsync(b0->trace, 0, sizeof(long long) * (32 * 16 + 32 * 8 + 32 * 4), f->stream));
  int rc = launch_advi_persist(f, n_iter);
  g_persist_trace = 0;
  BVG_TRY(rc);
  f->h_t_grad += (uint32_t)n_iter;
  PersistBuf *b = (PersistBuf *)((TcState *)f->tc)->persist;
  (void)n_iter;
  CUDA_TRY(cudaMemcpyAsync(out, b->trace, sizeof(long long) * (32 * 16 + 32 * 8 + 32 * 4), cudaMemcpyDeviceToHost, f->stream));
  CUDA_TRY(cudaStreamSynchronize(f->stream));
  return BVG_OK;
}
