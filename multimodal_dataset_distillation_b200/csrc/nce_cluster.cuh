// Bidirectional InfoNCE (primal and tangent) as ONE launch each: a thread-block cluster of 8 CTAs owns the B x B logit
// matrix, rows are split across the CTAs and the column statistics are exchanged through distributed shared memory.
//
//   reference: distill.py:548-551 (logits, 2 x cross_entropy) and its double backward (distill.py:562-567, 606)
//   formulas:  oracle/distill_ref.py::step_first_order / step_tangent ; DESIGN.md section 4
//
// Alternative (VLDD_NCE=cluster) to nce_rows + nce_cols + nce_grad (3 launches) and nce_t_rows + nce_t_cols + nce_t_grad +
// nce_t_finish (4 launches) of head_kernels.cuh for batches whose row blocks fit shared memory (B <= ~440).  Not the default:
// it measured 1.5% slower per iteration at B = 100 (engine.cu::nce_fused has the numbers); scores are bit-identical.
// All sums run in a fixed order (slabs, lanes, CTA ranks): results are bit-reproducible run to run.
#pragma once
#include <cooperative_groups.h>

#include "head_kernels.cuh"

namespace vldd {

namespace cg = cooperative_groups;

constexpr int kNceCluster = 8;      // portable maximum
constexpr int kNceThreads = 512;

inline size_t nce_cluster_smem_bytes(int B, int ld) {
  const int rows_per = ceil_div(B, kNceCluster);
  return (2 * (size_t)rows_per * ld + 5 * (size_t)ld + 64) * sizeof(float);   // tangent kernel: Sd and S row blocks
}
inline bool nce_cluster_ok(int B, int ld) { return nce_cluster_smem_bytes(B, ld) <= 200 * 1024; }

template <typename... KArgs, typename... Args>
inline void launch_cluster_k(void (*kern)(KArgs...), int cluster, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(cluster);
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// Pass 0 of both kernels: this CTA's rows of (scale * sum of the split-K slabs) -> shared memory, two elements per thread
// per trip with all 2 x 12 slab loads in flight (the 36-slab reduction is what the kernel's time is made of).
// Same association as sum_slabs_ilp, so the scores are bit-identical to the multi-launch path.
template <bool kStageS>
__device__ __forceinline__ void nce_reduce_rows(const float* __restrict__ part, int splits, size_t stride, float sc, int B,
                                                int ld, int r0, int nrows, float* __restrict__ out_l,
                                                float* __restrict__ out_g, const float* __restrict__ S_g,
                                                float* __restrict__ S_l) {
  const int nelem = nrows * B;
  for (int e = threadIdx.x; e < nelem; e += blockDim.x) {
    const int r = e / B, j = e - r * B;
    const size_t ij = (size_t)(r0 + r) * ld + j;
    float sv = 0.f;
    if (kStageS) sv = S_g[ij];                                              // in flight together with the slab loads
    const float v = sc * sum_slabs_ilp(part, splits, stride, (size_t)(r0 + r) * B + j);
    out_l[(size_t)r * ld + j] = v;
    if (out_g != nullptr) out_g[ij] = v;
    if (kStageS) S_l[(size_t)r * ld + j] = sv;
  }
}

// S = scale * sum_slabs ; lse_r, lse_c ; G = (softmax_rows + softmax_cols - 2I) / 2B ; loss
__global__ void __launch_bounds__(kNceThreads) nce_cluster_kernel(const float* __restrict__ part, int splits, size_t stride,
                                                                  const float* __restrict__ scale, int B, int ld,
                                                                  float* __restrict__ S, float* __restrict__ lse_r,
                                                                  float* __restrict__ lse_c, float* __restrict__ G,
                                                                  float* __restrict__ loss_out) {
  pdl_enter();
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ float nce_sm[];
  const int rank = (int)cluster.block_rank();
  const int rows_per = ceil_div(B, kNceCluster);
  const int r0 = rank * rows_per;
  const int nrows = max(min(B, r0 + rows_per) - r0, 0);
  float* Sl = nce_sm;                         // [rows_per][ld]   this CTA's rows of S
  float* lr_s = Sl + (size_t)rows_per * ld;   // [<= ld]          lse of those rows
  float* cm = lr_s + ld;                      // [ld]             column max over the local rows
  float* cs = cm + ld;                        // [ld]             column sum exp(. - cm)
  float* lc_s = cs + ld;                      // [ld]             full column lse
  float* red = lc_s + ld;                     // [32] warp partials, [32 + q] per-rank loss partials (rank 0 only)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const float sc = *scale, inv2B = 0.5f / B;

  nce_reduce_rows<false>(part, splits, stride, sc, B, ld, r0, nrows, Sl, S, nullptr, nullptr);
  __syncthreads();
  for (int r = warp; r < nrows; r += nw) {
    float mx = -INFINITY;
    for (int j = lane; j < B; j += 32) mx = fmaxf(mx, Sl[(size_t)r * ld + j]);
    mx = warp_max(mx);
    float se = 0.f;
    for (int j = lane; j < B; j += 32) se += expf(Sl[(size_t)r * ld + j] - mx);
    se = warp_sum(se);
    if (lane == 0) {
      const float l = mx + logf(se);
      lr_s[r] = l;
      lse_r[r0 + r] = l;
    }
  }
  for (int j = threadIdx.x; j < B; j += blockDim.x) {
    float m = -INFINITY;
    for (int r = 0; r < nrows; ++r) m = fmaxf(m, Sl[(size_t)r * ld + j]);
    float s = 0.f;
    for (int r = 0; r < nrows; ++r) s += expf(Sl[(size_t)r * ld + j] - m);
    cm[j] = m;
    cs[j] = s;
  }
  cluster.sync();
  for (int j = threadIdx.x; j < B; j += blockDim.x) {
    float ms[kNceCluster], ss[kNceCluster], M = -INFINITY;
#pragma unroll
    for (int q = 0; q < kNceCluster; ++q) {
      ms[q] = cluster.map_shared_rank(cm, q)[j];
      ss[q] = cluster.map_shared_rank(cs, q)[j];
      M = fmaxf(M, ms[q]);
    }
    float tot = 0.f;
#pragma unroll
    for (int q = 0; q < kNceCluster; ++q)
      if (ss[q] > 0.f) tot += ss[q] * expf(ms[q] - M);
    const float l = M + logf(tot);
    lc_s[j] = l;
    if (rank == 0) lse_c[j] = l;
  }
  __syncthreads();
  float lpart = 0.f;
  for (int r = warp; r < nrows; r += nw) {
    const int i = r0 + r;
    const float li = lr_s[r];
    for (int j = lane; j < B; j += 32) {
      const float s = Sl[(size_t)r * ld + j];
      float g = expf(s - li) + expf(s - lc_s[j]);
      if (j == i) g -= 2.0f;
      G[(size_t)i * ld + j] = g * inv2B;
    }
    if (lane == 0) {
      const float sii = Sl[(size_t)r * ld + i];
      lpart += (li - sii) + (lc_s[i] - sii);
    }
  }
  if (lane == 0) red[warp] = lpart;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < nw; ++w) t += red[w];
    cluster.map_shared_rank(red, 0)[32 + rank] = t;      // push into rank 0: nobody reads this CTA after the barrier
  }
  cluster.sync();
  if (rank == 0 && threadIdx.x == 0 && loss_out != nullptr) {
    float t = 0.f;
    for (int q = 0; q < kNceCluster; ++q) t += red[32 + q];
    *loss_out = t * inv2B;
  }
}

// Tangent: Sd = scale * sum_slabs ; rho_i = sum_j Pr Sd ; kap_j = sum_i Pc Sd ;
//          Gd = (Pr (Sd - rho_i) + Pc (Sd - kap_j)) / 2B ; Ld = sum G Sd ; dlr -= Ld ; dscale -= lr (sum Gd S + Ld) / scale
__global__ void __launch_bounds__(kNceThreads) nce_t_cluster_kernel(const float* __restrict__ part, int splits, size_t stride,
                                                                    const float* __restrict__ scale,
                                                                    const float* __restrict__ S, const float* __restrict__ lse_r,
                                                                    const float* __restrict__ lse_c, const float* __restrict__ G,
                                                                    int B, int ld, float* __restrict__ Gd,
                                                                    const float* __restrict__ lr, float* __restrict__ dlr,
                                                                    float* __restrict__ dscale) {
  pdl_enter();
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ float nce_sm[];
  const int rank = (int)cluster.block_rank();
  const int rows_per = ceil_div(B, kNceCluster);
  const int r0 = rank * rows_per;
  const int nrows = max(min(B, r0 + rows_per) - r0, 0);
  float* Sdl = nce_sm;                         // [rows_per][ld]   this CTA's rows of Sd
  float* Sl = Sdl + (size_t)rows_per * ld;     // [rows_per][ld]   ... and of S
  float* rho_s = Sl + (size_t)rows_per * ld;   // [<= ld]
  float* kp = rho_s + ld;                      // [ld]   column partial over the local rows
  float* kap_s = kp + ld;                      // [ld]
  float* lc_s = kap_s + ld;                    // [ld]
  float* red = lc_s + ld;                      // [0,32) A / [32,64) B warp partials, [64,80) per-rank partials (rank 0)
  float* redB = red + 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const float sc = *scale, inv2B = 0.5f / B;

  for (int j = threadIdx.x; j < B; j += blockDim.x) lc_s[j] = lse_c[j];
  nce_reduce_rows<true>(part, splits, stride, sc, B, ld, r0, nrows, Sdl, nullptr, S, Sl);
  __syncthreads();
  float accA = 0.f, accB = 0.f;
  for (int r = warp; r < nrows; r += nw) {
    const int i = r0 + r;
    const float l = lse_r[i];
    float a = 0.f, b = 0.f;
    for (int j0 = 0; j0 < B; j0 += 128) {                       // G: four independent loads per lane per trip
      float g[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + 32 * u + lane;
        g[u] = j < B ? G[(size_t)i * ld + j] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + 32 * u + lane;
        if (j < B) {
          const float v = Sdl[(size_t)r * ld + j];
          a = fmaf(expf(Sl[(size_t)r * ld + j] - l), v, a);
          b = fmaf(g[u], v, b);
        }
      }
    }
    a = warp_sum(a);
    b = warp_sum(b);
    if (lane == 0) { rho_s[r] = a; accA += b; }
  }
  for (int j = threadIdx.x; j < B; j += blockDim.x) {
    const float lc = lc_s[j];
    float k = 0.f;
    for (int r = 0; r < nrows; ++r) k = fmaf(expf(Sl[(size_t)r * ld + j] - lc), Sdl[(size_t)r * ld + j], k);
    kp[j] = k;
  }
  cluster.sync();
  for (int j = threadIdx.x; j < B; j += blockDim.x) {
    float k = 0.f;
#pragma unroll
    for (int q = 0; q < kNceCluster; ++q) k += cluster.map_shared_rank(kp, q)[j];
    kap_s[j] = k;
  }
  __syncthreads();
  for (int r = warp; r < nrows; r += nw) {
    const int i = r0 + r;
    const float l = lse_r[i], rh = rho_s[r];
    float c = 0.f;
    for (int j = lane; j < B; j += 32) {
      const float s = Sl[(size_t)r * ld + j], sd = Sdl[(size_t)r * ld + j];
      const float g = (expf(s - l) * (sd - rh) + expf(s - lc_s[j]) * (sd - kap_s[j])) * inv2B;
      Gd[(size_t)i * ld + j] = g;
      c = fmaf(g, s, c);
    }
    c = warp_sum(c);
    if (lane == 0) accB += c;
  }
  if (lane == 0) { red[warp] = accA; redB[warp] = accB; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float ta = 0.f, tb = 0.f;
    for (int w = 0; w < nw; ++w) { ta += red[w]; tb += redB[w]; }
    float* r0red = cluster.map_shared_rank(red, 0);
    r0red[64 + rank] = ta;
    r0red[64 + kNceCluster + rank] = tb;
  }
  cluster.sync();
  if (rank == 0 && threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int q = 0; q < kNceCluster; ++q) { a += red[64 + q]; b += red[64 + kNceCluster + q]; }
    *dlr -= a;
    *dscale -= (*lr) * (b + a) / (*scale);
  }
}

}  // namespace vldd
