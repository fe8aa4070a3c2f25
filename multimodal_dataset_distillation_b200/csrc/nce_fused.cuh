// InfoNCE block of one student step for small batches without tensor-core launches.
//
//   reference sites: distill.py:548-551 (logits, bidirectional cross-entropy) and their part of 562-567 / 606
//
// The logits are a B x B matrix (100 x 100 at the Flickr shape: 23 MFLOP, 40 KB) sitting between two full-width row
// kernels.  As tensor-core GEMMs they were the worst-shaped launches of the step -- S = s X Yn^T needs a 36-way split-K
// (36 partial slabs summed by the consumer) and dYn = s G^T X has K = 100 and fills 18 SMs -- and the softmax
// statistics took three more dependent launches.  Two CUDA-core kernels replace the five:
//   small_scores_kernel   S = scale * X Y^T, 8 x 12 output tiles; 12 warps per CTA = six 4 x 4 micro-tiles x two halves of
//                         K, lanes striding their half in 128-bit loads (exact fp32 FMA, fixed order: bit-reproducible),
//                         no split-K slabs in HBM;
//   nce_gx_kernel         every CTA loads S (40 KB) into shared memory, derives the softmax statistics itself (redundantly:
//                         2 exp per logit, cheaper than a launch boundary), forms G = (Pr + Pc - 2I) / 2B and produces its
//                         16-column block of G^T X from shared memory; four CTAs also write G, Pr, Pc, the log-sum-exps and
//                         the loss for the reverse sweep;
//   nce_t_gx_kernel       the tangent of the same block from the saved Pr, Pc (no exp at all): rho, kappa, Gd, L_dot,
//                         dlr / dscale and Gd^T X.
// exp is the hardware ex2 path (__expf, relative error ~2^-21: far inside the 1e-4 budget); log-sum-exps use logf.
#pragma once
#include "common.cuh"
#include "tc_gemm.cuh"      // mbarrier / shared-address helpers

namespace vldd {

#ifdef VLDD_NCE_TIMELINE
__device__ long long g_nce_tl[3][16];      // [kernel][slot]: globaltimer of CTA 5, thread 0 at phase boundaries (dev harness)
__device__ __forceinline__ long long nce_gtime() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define NTL(kern, slot) do { if (blockIdx.x == 5 % gridDim.x && blockIdx.y == 0 && threadIdx.x == 0) g_nce_tl[kern][slot] = nce_gtime(); } while (0)
#else
#define NTL(kern, slot) do {} while (0)
#endif

// One bulk asynchronous copy global -> shared (TMA engine, no tensor map): everything a CTA needs is requested in one go
// and lands behind a single mbarrier -- these kernels are latency-bound, a loop of dependent loads costs ~1 us per trip.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tc::smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(tc::smem_u32(bar))
               : "memory");
}

constexpr int kNceGxThreads = 512;
constexpr int kNceGxCols = 16;        // output columns per CTA
constexpr int kNceGxParts = kNceGxThreads / 128;

// shared-memory images keep the global layout [B][ld] (ld = B rounded up to 32: conflict-free rows and columns)
inline size_t nce_gx_smem_bytes(int B, int ld) {
  return (3 * (size_t)B * ld + (size_t)B * kNceGxCols + 3 * 128 * 16 + kNceGxParts * 128 + 8 * 128 + 64) * sizeof(float) + 64;
}
constexpr int kScoresThreads = 384;
constexpr int kScoresTI = 8, kScoresTJ = 12;
inline size_t scores_smem_bytes(int d) { return (size_t)(kScoresTI + kScoresTJ) * d * sizeof(float) + 64; }
inline bool nce_small_ok(int B, int d, int ld, const void* a, const void* b, const void* c) {
  return B >= 1 && B <= 128 && d % 8 == 0 && ld % 4 == 0 && aligned16(a) && aligned16(b) && aligned16(c) &&
         nce_gx_smem_bytes(B, ld) <= 220 * 1024 && scores_smem_bytes(d) <= 220 * 1024;
}

// ---- S = scale * X Y^T -------------------------------------------------------------------------------------
// CTA = 8 x 12 output tile = six 4 x 4 micro-tiles x two halves of K (12 warps); 9 x 13 = 117 CTAs at B = 100: one wave.
// The 20 operand rows (184 KB at d = 2304) are requested with 40 bulk copies up front, one mbarrier per half of K.
__global__ void __launch_bounds__(kScoresThreads) small_scores_kernel(const float* __restrict__ X, const float* __restrict__ Y, int B,
                                                                      int d, const float* __restrict__ scale, float* __restrict__ S,
                                                                      int ld) {
  extern __shared__ __align__(16) uint8_t smem_sc_raw[];
  __shared__ float part[2][6][16];                      // [K half][micro-tile][4 x 4]
  __shared__ __align__(8) uint64_t bars[2];
  float* rows = reinterpret_cast<float*>(smem_sc_raw);    // [20][d]: 8 rows of X, then 12 rows of Y
  NTL(0, 0);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int n4 = d >> 2, half4 = n4 >> 1;                 // d % 8 == 0: both halves are whole float4s
  if (threadIdx.x == 0) {
    tc::mbar_init(&bars[0], 1);
    tc::mbar_init(&bars[1], 1);
    tc::fence_barrier_init();
  }
  __syncthreads();
  pdl_enter();
  NTL(0, 1);
  if (w == 0) {
    const uint32_t half_bytes = (uint32_t)half4 * 16u;
    if (lane == 0) {
      tc::mbar_expect_tx(&bars[0], half_bytes * (kScoresTI + kScoresTJ));
      tc::mbar_expect_tx(&bars[1], half_bytes * (kScoresTI + kScoresTJ));
    }
    __syncwarp();
    if (lane < kScoresTI + kScoresTJ) {
      const bool isx = lane < kScoresTI;
      const int r = isx ? blockIdx.y * kScoresTI + lane : blockIdx.x * kScoresTJ + (lane - kScoresTI);
      const float* src = (isx ? X : Y) + (size_t)min(r, B - 1) * d;          // rows past B: clamped, never stored
      bulk_g2s(rows + (size_t)lane * d, src, half_bytes, &bars[0]);
      bulk_g2s(rows + (size_t)lane * d + 4 * half4, src + 4 * half4, half_bytes, &bars[1]);
    }
  }
  const int micro = w % 6, kh = w / 6;
  const float4* xr[4];
  const float4* yr[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    xr[a] = reinterpret_cast<const float4*>(rows + (size_t)(4 * (micro / 3) + a) * d);
    yr[a] = reinterpret_cast<const float4*>(rows + (size_t)(kScoresTI + 4 * (micro % 3) + a) * d);
  }
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  tc::mbar_wait(&bars[kh], 0);
  NTL(0, 2);
  const int kend = (kh + 1) * half4;
#pragma unroll 2
  for (int k = kh * half4 + lane; k < kend; k += 32) {
    float4 xv[4], yv[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) { xv[a] = xr[a][k]; yv[a] = yr[a][k]; }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        acc[a][b] = fmaf(xv[a].x, yv[b].x, acc[a][b]);
        acc[a][b] = fmaf(xv[a].y, yv[b].y, acc[a][b]);
        acc[a][b] = fmaf(xv[a].z, yv[b].z, acc[a][b]);
        acc[a][b] = fmaf(xv[a].w, yv[b].w, acc[a][b]);
      }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const float v = warp_sum(acc[a][b]);
      if (lane == a * 4 + b) part[kh][micro][lane] = v;
    }
  NTL(0, 3);
  __syncthreads();
  if (threadIdx.x < 96) {
    const int mt = threadIdx.x >> 4, e = threadIdx.x & 15;
    const int i = blockIdx.y * kScoresTI + 4 * (mt / 3) + (e >> 2), j = blockIdx.x * kScoresTJ + 4 * (mt % 3) + (e & 3);
    const float v = part[0][mt][e] + part[1][mt][e];
    if (i < B && j < B) S[(size_t)i * ld + j] = (*scale) * v;
  }
  NTL(0, 4);
}

// ---- shared pieces of the two G^T X kernels ---------------------------------------------------------------------
// out[j, c0 + c] = sum_i Gs[i][j] * Xs[i][c]  for the CTA's 16 columns.  Thread u < 4 * ceil(B/4) of each group owns a 4 (j) x 4 (c)
// block; up to four groups split the summation index and are combined through shared memory in a fixed order.
__device__ __forceinline__ void gtx_block(const float* __restrict__ Gs, int lds, const float* __restrict__ Xs, int B, int d, int c0,
                                          float* __restrict__ comb, float* __restrict__ out) {
  const int per = ((B + 3) / 4) * 4;                  // threads per group (<= 128)
  const int ngroups = min(4, (int)blockDim.x / per);
  const int t = threadIdx.x, g = t / per, u = t - g * per;
  const int jg = u >> 2, cg = u & 3;
  const int chunk = (B + ngroups - 1) / ngroups;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  if (g < ngroups) {
    const int ib = g * chunk, ie = min(B, ib + chunk);
    for (int i = ib; i < ie; ++i) {
      const float4 gq = *reinterpret_cast<const float4*>(Gs + (size_t)i * lds + 4 * jg);
      const float4 x = *reinterpret_cast<const float4*>(Xs + i * kNceGxCols + 4 * cg);
      const float gv[4] = {gq.x, gq.y, gq.z, gq.w}, xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(gv[a], xv[b], acc[a][b]);
    }
    if (g > 0) {
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) comb[((g - 1) * 128 + u) * 16 + a * 4 + b] = acc[a][b];
    }
  }
  __syncthreads();
  if (g == 0) {
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int j = 4 * jg + a, c = c0 + 4 * cg;
      if (j < B && c < d) {
        float o[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          float v = acc[a][b];
          for (int gg = 1; gg < ngroups; ++gg) v += comb[((gg - 1) * 128 + u) * 16 + a * 4 + b];
          o[b] = v;
        }
        *reinterpret_cast<float4*>(out + (size_t)j * d + c) = make_float4(o[0], o[1], o[2], o[3]);   // d % 4 == 0, c % 4 == 0
      }
    }
  }
}
// [B, ld] global matrices -> shared memory (same layout) and the CTA's 16-column block of X, all behind one mbarrier.
// Called by every thread; thread 0 arms the barrier, threads < B each fetch one 64-byte row segment of X.
__device__ __forceinline__ void bulk_load_inputs(uint64_t* bar, int n_mats, const float* const* mats, float* const* mats_s, int B, int ld,
                                                 const float* __restrict__ X, int d, int c0, float* __restrict__ Xs) {
  const int cols = min(kNceGxCols, d - c0);              // multiple of 4
  if (threadIdx.x == 0) {
    tc::mbar_init(bar, 1);
    tc::fence_barrier_init();
  }
  if (cols < kNceGxCols)
    for (int e = threadIdx.x; e < B * kNceGxCols; e += blockDim.x) Xs[e] = 0.f;
  __syncthreads();
  pdl_enter();
  if (threadIdx.x == 0) {
    tc::mbar_expect_tx(bar, (uint32_t)(n_mats * B * ld * 4 + B * cols * 4));
    for (int m = 0; m < n_mats; ++m) bulk_g2s(mats_s[m], mats[m], (uint32_t)(B * ld * 4), bar);
  }
  __syncthreads();                                       // the expectation is armed before any row copy can complete
  if ((int)threadIdx.x < B) bulk_g2s(Xs + threadIdx.x * kNceGxCols, X + (size_t)threadIdx.x * d + c0, (uint32_t)(cols * 4), bar);
  tc::mbar_wait(bar, 0);
}
__device__ __forceinline__ void store_square(const float* __restrict__ Ms, int lds, int B, int ld, float* __restrict__ M) {
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int i = w; i < B; i += nw)
    for (int j = lane; j < B; j += 32) M[(size_t)i * ld + j] = Ms[(size_t)i * lds + j];
}
// column reductions with all threads: thread (j = t % 128, part = t / 128) covers rows part, part + parts, ...
template <class F>
__device__ __forceinline__ float col_reduce_max(int B, float* __restrict__ tmp, F f) {
  const int j = threadIdx.x & 127, part = threadIdx.x >> 7, parts = blockDim.x >> 7;
  float m = -INFINITY;
  if (j < B)
    for (int i = part; i < B; i += parts) m = fmaxf(m, f(i, j));
  tmp[part * 128 + j] = m;
  __syncthreads();
  float r = -INFINITY;
  for (int p = 0; p < parts; ++p) r = fmaxf(r, tmp[p * 128 + j]);
  __syncthreads();
  return r;
}
template <class F>
__device__ __forceinline__ float col_reduce_sum(int B, float* __restrict__ tmp, F f) {
  const int j = threadIdx.x & 127, part = threadIdx.x >> 7, parts = blockDim.x >> 7;
  float s = 0.f;
  if (j < B)
    for (int i = part; i < B; i += parts) s += f(i, j);
  tmp[part * 128 + j] = s;
  __syncthreads();
  float r = 0.f;
  for (int p = 0; p < parts; ++p) r += tmp[p * 128 + j];
  __syncthreads();
  return r;
}

// ---- primal: statistics, G, loss, dYn_raw = G^T X ---------------------------------------------------------------
__global__ void __launch_bounds__(kNceGxThreads) nce_gx_kernel(const float* __restrict__ S, const float* __restrict__ X, int B, int ld,
                                                                int d, float* __restrict__ lse_r_out, float* __restrict__ lse_c_out,
                                                                float* __restrict__ G_out, float* __restrict__ Pr_out,
                                                                float* __restrict__ Pc_out, float* __restrict__ loss_out,
                                                                float* __restrict__ out) {
  extern __shared__ __align__(16) float smem_nce[];
  __shared__ __align__(8) uint64_t bar;
  const int lds = ld;
  const size_t mat = (size_t)B * lds;
  float* Ss = smem_nce;                              // S, then G
  float* Es = Ss + mat;                              // exp(S - row max), then Pr
  float* Fs = Es + mat;                              // exp(S - column max), then Pc
  float* Xs = Fs + mat;                              // [B][16]
  float* comb = Xs + B * kNceGxCols;                 // [3][128][16]
  float* tmp = comb + 3 * 128 * 16;                  // [parts][128]
  float* lse_r = tmp + kNceGxParts * 128;            // [128] each
  float* lse_c = lse_r + 128;
  float* rinv = lse_c + 128;
  float* cinv = rinv + 128;
  float* diag = cinv + 128;
  float* scratch = diag + 128;
  const int c0 = blockIdx.x * kNceGxCols;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  NTL(1, 0);
  {
    const float* mats[1] = {S};
    float* mats_s[1] = {Ss};
    bulk_load_inputs(&bar, 1, mats, mats_s, B, ld, X, d, c0, Xs);
  }
  NTL(1, 1);
  // rows: one warp per row
  for (int i = w; i < B; i += nw) {
    const float* row = Ss + (size_t)i * lds;
    float m = -INFINITY;
    for (int j = lane; j < B; j += 32) m = fmaxf(m, row[j]);
    m = warp_max(m);
    float s = 0.f;
    for (int j = lane; j < B; j += 32) {
      const float e = __expf(row[j] - m);
      Es[(size_t)i * lds + j] = e;
      s += e;
    }
    s = warp_sum(s);
    if (lane == 0) { rinv[i] = 1.0f / s; lse_r[i] = m + logf(s); diag[i] = row[i]; }
  }
  NTL(1, 2);
  // columns
  const float cm = col_reduce_max(B, tmp, [&](int i, int j) { return Ss[(size_t)i * lds + j]; });
  const float cs = col_reduce_sum(B, tmp, [&](int i, int j) {
    const float f = __expf(Ss[(size_t)i * lds + j] - cm);
    Fs[(size_t)i * lds + j] = f;
    return f;
  });
  if (threadIdx.x < B) { cinv[threadIdx.x] = 1.0f / cs; lse_c[threadIdx.x] = cm + logf(cs); }
  __syncthreads();
  NTL(1, 3);
  // G = (Pr + Pc - 2I) / 2B
  const float inv2B = 0.5f / B;
  for (int i = w; i < B; i += nw) {
    const float ri = rinv[i];
    for (int j = lane; j < B; j += 32) {
      const size_t ij = (size_t)i * lds + j;
      const float pr = Es[ij] * ri, pc = Fs[ij] * cinv[j];
      Es[ij] = pr;
      Fs[ij] = pc;
      Ss[ij] = (pr + pc - (i == j ? 2.0f : 0.f)) * inv2B;
    }
  }
  __syncthreads();
  NTL(1, 4);
  // what the reverse sweep reads from HBM, written by four different CTAs (or by fewer, if the grid is that small)
  const int nb = gridDim.x;
  if ((int)blockIdx.x == 0 % nb) store_square(Ss, lds, B, ld, G_out);
  if ((int)blockIdx.x == 1 % nb) store_square(Es, lds, B, ld, Pr_out);
  if ((int)blockIdx.x == 2 % nb) store_square(Fs, lds, B, ld, Pc_out);
  if ((int)blockIdx.x == 3 % nb) {
    if (threadIdx.x < B) { lse_r_out[threadIdx.x] = lse_r[threadIdx.x]; lse_c_out[threadIdx.x] = lse_c[threadIdx.x]; }
    if (loss_out != nullptr) {
      float acc = 0.f;
      for (int r = threadIdx.x; r < B; r += blockDim.x) acc += (lse_r[r] - diag[r]) + (lse_c[r] - diag[r]);
      acc = block_sum<float>(acc, scratch);
      if (threadIdx.x == 0) *loss_out = acc * inv2B;
    }
  }
  NTL(1, 5);
  gtx_block(Ss, lds, Xs, B, d, c0, comb, out);
  NTL(1, 6);
}

// ---- tangent: rho, kappa, Gd, L_dot, dlr / dscale, dYnd_raw = Gd^T X ---------------------------------------------
//   rho_i = sum_j Pr Sd, kap_j = sum_i Pc Sd;  Gd = (Pr (Sd - rho_i) + Pc (Sd - kap_j)) / 2B
//   Ld = sum G o Sd = (sum rho + sum kap - 2 tr Sd) / 2B;  GdS = sum Gd o S;  dlr -= Ld;  dscale -= lr (GdS + Ld) / scale
//   (one CTA applies the two scalar updates: the steps of the reverse sweep are serialised on the stream)
__global__ void __launch_bounds__(kNceGxThreads) nce_t_gx_kernel(const float* __restrict__ S, const float* __restrict__ Sd,
                                                                  const float* __restrict__ Pr, const float* __restrict__ Pc,
                                                                  const float* __restrict__ X, int B, int ld, int d,
                                                                  const float* __restrict__ lr, const float* __restrict__ scale,
                                                                  float* __restrict__ Gd_out, float* __restrict__ dlr,
                                                                  float* __restrict__ dscale, float* __restrict__ out) {
  extern __shared__ __align__(16) float smem_nce[];
  __shared__ __align__(8) uint64_t bar;
  const int lds = ld;
  const size_t mat = (size_t)B * lds;
  float* Ds = smem_nce;                              // Sd, then Gd
  float* Es = Ds + mat;                              // Pr
  float* Fs = Es + mat;                              // Pc
  float* Xs = Fs + mat;
  float* comb = Xs + B * kNceGxCols;
  float* tmp = comb + 3 * 128 * 16;
  float* rho = tmp + kNceGxParts * 128;
  float* kaps = rho + 128;
  float* scratch = kaps + 128 * 4;
  const int c0 = blockIdx.x * kNceGxCols;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  NTL(2, 0);
  {
    const float* mats[3] = {Sd, Pr, Pc};
    float* mats_s[3] = {Ds, Es, Fs};
    bulk_load_inputs(&bar, 3, mats, mats_s, B, ld, X, d, c0, Xs);
  }
  NTL(2, 1);
  for (int i = w; i < B; i += nw) {
    float a = 0.f;
    for (int j = lane; j < B; j += 32) a = fmaf(Es[(size_t)i * lds + j], Ds[(size_t)i * lds + j], a);
    a = warp_sum(a);
    if (lane == 0) rho[i] = a;
  }
  NTL(2, 2);
  const float kap = col_reduce_sum(B, tmp, [&](int i, int j) { return Fs[(size_t)i * lds + j] * Ds[(size_t)i * lds + j]; });
  if (threadIdx.x < 128) kaps[threadIdx.x] = kap;
  __syncthreads();
  NTL(2, 3);
  const float inv2B = 0.5f / B;
  const bool scalar_cta = (int)blockIdx.x == 1 % (int)gridDim.x;
  float gds = 0.f, tr = 0.f;
  for (int i = w; i < B; i += nw) {
    const float rh = rho[i];
    for (int j = lane; j < B; j += 32) {
      const size_t ij = (size_t)i * lds + j;
      const float sd = Ds[ij];
      const float g = (Es[ij] * (sd - rh) + Fs[ij] * (sd - kaps[j])) * inv2B;
      Ds[ij] = g;
      if (scalar_cta) {
        gds = fmaf(g, S[(size_t)i * ld + j], gds);
        if (i == j) tr += sd;
      }
    }
  }
  __syncthreads();
  NTL(2, 4);
  if (blockIdx.x == 0) store_square(Ds, lds, B, ld, Gd_out);
  if (scalar_cta) {
    float srk = 0.f;
    for (int r = threadIdx.x; r < B; r += blockDim.x) srk += rho[r] + kaps[r];
    srk = block_sum<float>(srk, scratch);
    tr = block_sum<float>(tr, scratch);
    gds = block_sum<float>(gds, scratch);
    if (threadIdx.x == 0) {
      const float Ld = (srk - 2.0f * tr) * inv2B;
      *dlr -= Ld;
      *dscale -= (*lr) * (gds + Ld) / (*scale);
    }
  }
  NTL(2, 5);
  gtx_block(Ds, lds, Xs, B, d, c0, comb, out);
  NTL(2, 6);
}

}  // namespace vldd
