// Retrieval scoring (Path 2): rank of the ground-truth item per query row, recall@{1,5,10} counts,
// the similarity GEMM and the reference's top-k/-100 fill.
//
//   reference sites: epoch.py:219-244 / epoch_original.py:115-161 (itm_eval), epoch_original.py:94-105 (sims, topk fill)
//
// Rank definition (deterministic form of np.argsort(row)[::-1], ties broken by index -- see oracle/retrieval_ref.py):
//   rank(c) = #{j : s_j > s_c} + #{j < c : s_j == s_c};  i2t takes the minimum over the image's GT captions,
//   which is the rank of the GT caption with the highest score (lowest index among equals).
#include "common.cuh"
#include "gemm_dispatch.cuh"
#include "kernels.h"

namespace vldd {

// One CTA per query row.  The row is read from HBM exactly once (4 B / pair); the <=C ground-truth entries are
// gathered first (a few extra sectors).  Integer result => bit-exact by construction.
template <int THREADS>
__global__ void __launch_bounds__(THREADS) ranks_rows_kernel(const float* __restrict__ S, int64_t ld, int ncols,
                                                             const int32_t* __restrict__ gt_ptr,
                                                             const int32_t* __restrict__ gt_idx,
                                                             int32_t* __restrict__ ranks, int vec) {
  pdl_enter();
  __shared__ int scratch[34];
  __shared__ float s_best;
  __shared__ int c_best;
  const int r = blockIdx.x;
  const float* __restrict__ row = S + (size_t)r * ld;
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    const int beg = gt_ptr ? gt_ptr[r] : r, end = gt_ptr ? gt_ptr[r + 1] : r + 1;
    float bs = -INFINITY;
    int bc = INT_MAX;
    bool any = false;
    for (int e = beg + lane; e < end; e += 32) {
      const int c = gt_idx[e];
      if (c < 0 || c >= ncols) continue;
      const float s = row[c];
      if (!any || s > bs || (s == bs && c < bc)) { bs = s; bc = c; any = true; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, bs, o);
      const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
      const bool oa = __shfl_xor_sync(0xffffffffu, (int)any, o);
      if (oa && (!any || os > bs || (os == bs && oc < bc))) { bs = os; bc = oc; any = true; }
    }
    if (lane == 0) { s_best = bs; c_best = any ? bc : -1; }
  }
  __syncthreads();
  const float sg = s_best;
  const int cg = c_best;
  if (cg < 0) {  // no valid ground truth for this row: never retrieved
    if (threadIdx.x == 0) ranks[r] = ncols;
    return;
  }
  int cnt = 0;
  if (vec) {
    const int n4 = ncols >> 2;
    for (int i = threadIdx.x; i < n4; i += THREADS) {
      const float4 v = ldg_stream4(row + 4 * i);
      const int j = 4 * i;
      cnt += (v.x > sg) + (v.x == sg && j + 0 < cg);
      cnt += (v.y > sg) + (v.y == sg && j + 1 < cg);
      cnt += (v.z > sg) + (v.z == sg && j + 2 < cg);
      cnt += (v.w > sg) + (v.w == sg && j + 3 < cg);
    }
    for (int j = (n4 << 2) + threadIdx.x; j < ncols; j += THREADS) {
      const float v = row[j];
      cnt += (v > sg) + (v == sg && j < cg);
    }
  } else {
    for (int j = threadIdx.x; j < ncols; j += THREADS) {
      const float v = row[j];
      cnt += (v > sg) + (v == sg && j < cg);
    }
  }
  const int total = block_sum<int>(cnt, scratch);
  if (threadIdx.x == 0) ranks[r] = total;
}

int ranks_rows(const float* S, int64_t ld, int nrows, int ncols, const int32_t* gt_ptr, const int32_t* gt_idx,
               int32_t* ranks, cudaStream_t st) {
  if (nrows <= 0) return VLDD_OK;
  const int vec = aligned16(S) && (ld % 4 == 0);
  if (ncols > 4096)
    launch_k(ranks_rows_kernel<256>, nrows, 256, 0, st, S, ld, ncols, gt_ptr, gt_idx, ranks, vec);
  else
    launch_k(ranks_rows_kernel<128>, nrows, 128, 0, st, S, ld, ncols, gt_ptr, gt_idx, ranks, vec);
  return check_launch("ranks_rows");
}

// Column-wise ranking on the SAME matrix: rank of row gt_row[c] inside column c (text->image retrieval on S_i2t without
// materialising the transpose).  Block = 32 columns x 8 row lanes; each warp reads 128 contiguous bytes of a row.
// counts must be zero on entry; integer atomics => order-independent, bit-exact.
constexpr int kColRowsPerBlock = 256;
__global__ void __launch_bounds__(256) ranks_cols_kernel(const float* __restrict__ S, int64_t ld, int nrows, int ncols,
                                                         const int32_t* __restrict__ gt_row, int32_t* __restrict__ counts) {
  pdl_enter();
  __shared__ int red[8][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const int r0 = blockIdx.y * kColRowsPerBlock, r1 = min(nrows, r0 + kColRowsPerBlock);
  int cnt = 0;
  if (c < ncols) {
    const int g = gt_row[c];
    if (g >= 0 && g < nrows) {
      const float thr = S[(size_t)g * ld + c];
      int r = r0 + w;
      for (; r + 24 < r1; r += 32) {            // 4 independent row loads in flight per thread
        const float v0 = S[(size_t)r * ld + c], v1 = S[(size_t)(r + 8) * ld + c], v2 = S[(size_t)(r + 16) * ld + c],
                    v3 = S[(size_t)(r + 24) * ld + c];
        cnt += (v0 > thr) + (v0 == thr && r < g);
        cnt += (v1 > thr) + (v1 == thr && r + 8 < g);
        cnt += (v2 > thr) + (v2 == thr && r + 16 < g);
        cnt += (v3 > thr) + (v3 == thr && r + 24 < g);
      }
      for (; r < r1; r += 8) {
        const float v = S[(size_t)r * ld + c];
        cnt += (v > thr) + (v == thr && r < g);
      }
    } else if (blockIdx.y == 0 && w == 0) {
      cnt = nrows;                                // no valid ground truth: never retrieved
    }
  }
  red[w][lane] = cnt;
  __syncthreads();
  if (w == 0 && c < ncols) {
    int t = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][lane];
    if (t) atomicAdd(counts + c, t);
  }
}

// 128-bit form: a warp reads 512 contiguous bytes of a row (each lane four neighbouring columns), eight warps take rows
// r0 + w, r0 + w + 8, ...; four independent row loads per thread in flight.  Needs 16-byte aligned rows (ld % 4 == 0).
__global__ void __launch_bounds__(256) ranks_cols_vec4_kernel(const float* __restrict__ S, int64_t ld, int nrows, int ncols,
                                                              const int32_t* __restrict__ gt_row, int32_t* __restrict__ counts) {
  pdl_enter();
  __shared__ int red[8][128];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 128 + lane * 4;
  const int r0 = blockIdx.y * kColRowsPerBlock, r1 = min(nrows, r0 + kColRowsPerBlock);
  int cnt[4] = {0, 0, 0, 0};
  if (c0 < ncols) {                              // ncols % 4 == 0 on this path, so c0 + 3 < ncols as well
    int g[4];
    float thr[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      g[j] = gt_row[c0 + j];
      const bool ok = g[j] >= 0 && g[j] < nrows;
      thr[j] = ok ? S[(size_t)g[j] * ld + c0 + j] : INFINITY;      // no ground truth: nothing counts, the finalise step fixes it
      if (!ok) g[j] = -1;
    }
    auto tally = [&](const float4 v, int r) {
      cnt[0] += (v.x > thr[0]) + (v.x == thr[0] && r < g[0]);
      cnt[1] += (v.y > thr[1]) + (v.y == thr[1] && r < g[1]);
      cnt[2] += (v.z > thr[2]) + (v.z == thr[2] && r < g[2]);
      cnt[3] += (v.w > thr[3]) + (v.w == thr[3] && r < g[3]);
    };
    int r = r0 + w;
    for (; r + 24 < r1; r += 32) {
      const float4 v0 = ldg_stream4(S + (size_t)r * ld + c0), v1 = ldg_stream4(S + (size_t)(r + 8) * ld + c0),
                   v2 = ldg_stream4(S + (size_t)(r + 16) * ld + c0), v3 = ldg_stream4(S + (size_t)(r + 24) * ld + c0);
      tally(v0, r); tally(v1, r + 8); tally(v2, r + 16); tally(v3, r + 24);
    }
    for (; r < r1; r += 8) tally(ldg_stream4(S + (size_t)r * ld + c0), r);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) red[w][lane * 4 + j] = cnt[j];
  __syncthreads();
  if (threadIdx.x < 128) {
    const int c = blockIdx.x * 128 + threadIdx.x;
    int t = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
    if (c < ncols && t) atomicAdd(counts + c, t);
  }
}

int ranks_cols(const float* S, int64_t ld, int nrows, int ncols, const int32_t* gt_row, int32_t* ranks, cudaStream_t st) {
  if (ncols <= 0) return VLDD_OK;
  cudaError_t e = cudaMemsetAsync(ranks, 0, (size_t)ncols * sizeof(int32_t), st);
  if (e != cudaSuccess) { set_error("memset failed: %s", cudaGetErrorString(e)); return VLDD_ERR_CUDA; }
  if (nrows <= 0) return VLDD_OK;
  if (aligned16(S) && ld % 4 == 0 && ncols % 4 == 0) {
    dim3 grid(ceil_div(ncols, 128), ceil_div(nrows, kColRowsPerBlock));
    launch_k(ranks_cols_vec4_kernel, grid, 256, 0, st, S, ld, nrows, ncols, gt_row, ranks);
    return check_launch("ranks_cols");
  }
  dim3 grid(ceil_div(ncols, 32), ceil_div(nrows, kColRowsPerBlock));
  launch_k(ranks_cols_kernel, grid, 256, 0, st, S, ld, nrows, ncols, gt_row, ranks);
  return check_launch("ranks_cols");
}

// ---- caption-sharded form (multi-GPU): the score matrix holds columns [col_offset, col_offset + ncols) of the full
// one.  Step 1: best ground-truth candidate per row among the LOCAL columns; step 2 (after the candidates of all
// shards are merged on the host side of the collective): count local columns that beat the global threshold.
__global__ void __launch_bounds__(32) best_gt_rows_kernel(const float* __restrict__ S, int64_t ld, int ncols, int col_offset,
                                                          const int32_t* __restrict__ gt_ptr,
                                                          const int32_t* __restrict__ gt_idx, float* __restrict__ best_score,
                                                          int32_t* __restrict__ best_idx) {
  pdl_enter();
  const int r = blockIdx.x, lane = threadIdx.x;
  const float* __restrict__ row = S + (size_t)r * ld;
  float bs = -INFINITY;
  int bc = INT_MAX;
  bool any = false;
  for (int e = gt_ptr[r] + lane; e < gt_ptr[r + 1]; e += 32) {
    const int c = gt_idx[e] - col_offset;            // global -> local column
    if (c < 0 || c >= ncols) continue;
    const float s = row[c];
    const int cg = c + col_offset;
    if (!any || s > bs || (s == bs && cg < bc)) { bs = s; bc = cg; any = true; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float os = __shfl_xor_sync(0xffffffffu, bs, o);
    const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
    const bool oa = __shfl_xor_sync(0xffffffffu, (int)any, o);
    if (oa && (!any || os > bs || (os == bs && oc < bc))) { bs = os; bc = oc; any = true; }
  }
  if (lane == 0) { best_score[r] = any ? bs : -INFINITY; best_idx[r] = any ? bc : -1; }
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS) count_rows_kernel(const float* __restrict__ S, int64_t ld, int ncols,
                                                             int col_offset, const float* __restrict__ thr_score,
                                                             const int32_t* __restrict__ thr_idx,
                                                             int32_t* __restrict__ counts) {
  pdl_enter();
  __shared__ int scratch[34];
  const int r = blockIdx.x;
  const float* __restrict__ row = S + (size_t)r * ld;
  const float sg = thr_score[r];
  const int cg = thr_idx[r] - col_offset;              // may lie outside [0, ncols): then only the sign of (j < cg) matters
  int cnt = 0;
  if (thr_idx[r] >= 0) {
    for (int j = threadIdx.x; j < ncols; j += THREADS) {
      const float v = row[j];
      cnt += (v > sg) + (v == sg && j < cg);
    }
  }
  const int total = block_sum<int>(cnt, scratch);
  if (threadIdx.x == 0) counts[r] = total;
}

int best_gt_rows(const float* S, int64_t ld, int nrows, int ncols, int col_offset, const int32_t* gt_ptr,
                 const int32_t* gt_idx, float* best_score, int32_t* best_idx, cudaStream_t st) {
  if (nrows <= 0) return VLDD_OK;
  launch_k(best_gt_rows_kernel, nrows, 32, 0, st, S, ld, ncols, col_offset, gt_ptr, gt_idx, best_score, best_idx);
  return check_launch("best_gt_rows");
}

int count_rows(const float* S, int64_t ld, int nrows, int ncols, int col_offset, const float* thr_score,
               const int32_t* thr_idx, int32_t* counts, cudaStream_t st) {
  if (nrows <= 0) return VLDD_OK;
  launch_k(count_rows_kernel<256>, nrows, 256, 0, st, S, ld, ncols, col_offset, thr_score, thr_idx, counts);
  return check_launch("count_rows");
}

// counts3 += (#ranks<1, #ranks<5, #ranks<10)     (epoch.py:227-229,236-238)
__global__ void __launch_bounds__(256) recall_counts_kernel(const int32_t* __restrict__ ranks, int n,
                                                            int32_t* __restrict__ counts3) {
  pdl_enter();
  __shared__ int scratch[34];
  int c1 = 0, c5 = 0, c10 = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int r = ranks[i];
    c1 += r < 1; c5 += r < 5; c10 += r < 10;
  }
  c1 = block_sum<int>(c1, scratch);
  c5 = block_sum<int>(c5, scratch);
  c10 = block_sum<int>(c10, scratch);
  if (threadIdx.x == 0) {
    atomicAdd(counts3 + 0, c1);
    atomicAdd(counts3 + 1, c5);
    atomicAdd(counts3 + 2, c10);
  }
}

int recall_counts(const int32_t* ranks, int n, int32_t* counts3, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(counts3, 0, 3 * sizeof(int32_t), st);
  if (e != cudaSuccess) { set_error("memset failed: %s", cudaGetErrorString(e)); return VLDD_ERR_CUDA; }
  if (n <= 0) return VLDD_OK;
  int grid = ceil_div(n, 256);
  if (grid > num_sms() * 4) grid = num_sms() * 4;
  launch_k(recall_counts_kernel, grid, 256, 0, st, ranks, n, counts3);
  return check_launch("recall_counts");
}

// S_i2t[i,t] = scale * <img_i, txt_t>;  S_t2i = transpose (second GEMM with swapped operands so both stores coalesce)
int sim_scores(const float* img, const float* txt, int I, int T, int D, float scale, float* S_i2t, float* S_t2i,
               cudaStream_t st) {
  if (I <= 0 || T <= 0) return VLDD_OK;
  if (S_i2t) {
    GemmOperands g = gemm_ops(img, D, txt, D, I, T, D);
    int rc = gemm_store<true, true>(g, S_i2t, T, scale, st);
    if (rc) return rc;
    rc = check_launch("sim_scores i2t");
    if (rc) return rc;
  }
  if (S_t2i) {
    GemmOperands g = gemm_ops(txt, D, img, D, T, I, D);
    int rc = gemm_store<true, true>(g, S_t2i, I, scale, st);
    if (rc) return rc;
    rc = check_launch("sim_scores t2i");
    if (rc) return rc;
  }
  return VLDD_OK;
}

// ---- nearest neighbour by cosine similarity (distill.py:89-95: sklearn cosine_similarity + np.argmax per query) ----------
__global__ void __launch_bounds__(256) unit_rows_kernel(const float* __restrict__ x, int d, float* __restrict__ out) {
  pdl_enter();
  __shared__ float scratch[34];
  const size_t base = (size_t)blockIdx.x * d;
  float s = 0.f;
  for (int j = threadIdx.x; j < d; j += blockDim.x) s = fmaf(x[base + j], x[base + j], s);
  s = block_sum<float>(s, scratch);
  // sklearn.preprocessing.normalize leaves all-zero rows unchanged (norm 0 -> divide by 1)
  const float inv = s > 0.f ? 1.0f / sqrtf(s) : 1.0f;
  for (int j = threadIdx.x; j < d; j += blockDim.x) out[base + j] = x[base + j] * inv;
}
// first index of the row maximum (np.argmax semantics), optionally the maximum itself
__global__ void __launch_bounds__(256) argmax_rows_kernel(const float* __restrict__ S, int64_t ld, int ncols,
                                                          int32_t* __restrict__ idx_out, float* __restrict__ val_out) {
  pdl_enter();
  __shared__ float sv[8];
  __shared__ int si[8];
  const float* row = S + (size_t)blockIdx.x * ld;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int j = threadIdx.x; j < ncols; j += blockDim.x) {
    const float v = row[j];
    if (v > best || (v == best && j < bi) || bi == 0x7fffffff) { best = v; bi = j; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (oi != 0x7fffffff && (bi == 0x7fffffff || ov > best || (ov == best && oi < bi))) { best = ov; bi = oi; }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { sv[wid] = best; si[wid] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
      if (si[w] != 0x7fffffff && (bi == 0x7fffffff || sv[w] > best || (sv[w] == best && si[w] < bi))) { best = sv[w]; bi = si[w]; }
    idx_out[blockIdx.x] = bi == 0x7fffffff ? 0 : bi;
    if (val_out) val_out[blockIdx.x] = best;
  }
}

size_t nearest_rows_workspace_bytes(int Q, int T, int D) {
  if (Q <= 0 || T <= 0 || D <= 0) return 0;
  return ((size_t)Q * D + (size_t)T * D + (size_t)Q * T) * sizeof(float) + 3 * 256;
}

int nearest_rows(const float* query, const float* bank, int Q, int T, int D, int32_t* idx_out, float* cos_out,
                 void* workspace, cudaStream_t st) {
  auto align = [](size_t b) { return (b + 255) / 256 * 256; };
  char* base = reinterpret_cast<char*>(workspace);
  float* qn = reinterpret_cast<float*>(base);
  float* bn = reinterpret_cast<float*>(base + align((size_t)Q * D * sizeof(float)));
  float* S = reinterpret_cast<float*>(base + align((size_t)Q * D * sizeof(float)) + align((size_t)T * D * sizeof(float)));
  launch_k(unit_rows_kernel, Q, 256, 0, st, query, D, qn);
  launch_k(unit_rows_kernel, T, 256, 0, st, bank, D, bn);
  int rc = sim_scores(qn, bn, Q, T, D, 1.0f, S, nullptr, st);
  if (rc) return rc;
  launch_k(argmax_rows_kernel, Q, 256, 0, st, (const float*)S, (int64_t)T, T, idx_out, cos_out);
  return check_launch("nearest_rows");
}

// ---- fused similarity + ranking: the score matrix is never written to HBM --------------------------------------------
// Two passes of the SAME tcgen05 GEMM kernel (bit-identical tile values): pass 1 visits only the tiles that hold a
// ground-truth pair and extracts those scores; pass 2 visits every tile and counts, per row and per column, the entries
// ranked ahead of the ground truth (integer atomics).  Result == ranking the materialised matrix, bit for bit.
__global__ void __launch_bounds__(256) mark_gt_tiles_kernel(const int32_t* __restrict__ gt_ptr, const int32_t* __restrict__ gt_idx,
                                                            const int32_t* __restrict__ txt2img, int I, int T, int tiles_n,
                                                            int col_offset, int* __restrict__ flags) {
  pdl_enter();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < I) {
    for (int e = gt_ptr[t]; e < gt_ptr[t + 1]; ++e) {
      const int c = gt_idx[e] - col_offset;
      if (c >= 0 && c < T) flags[(t / tc::BM) * tiles_n + c / 128] = 1;
    }
  }
  if (t < T) {
    const int g = txt2img[t];
    if (g >= 0 && g < I) flags[(g / tc::BM) * tiles_n + t / 128] = 1;
  }
}
__global__ void __launch_bounds__(256) compact_tiles_kernel(const int* __restrict__ flags, int n, int* __restrict__ list,
                                                            int* __restrict__ count) {
  pdl_enter();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n && flags[t]) list[atomicAdd(count, 1)] = t;
}
// Best ground-truth candidate per image among the captions of THIS product (columns [col_offset, col_offset + T) of the full
// problem): score (+inf when the image has none here -- as a threshold "+inf" means nothing ranks ahead) and the GLOBAL caption
// index (-1: none).  Also the validity of every local caption's own ground truth.
__global__ void __launch_bounds__(256) rank_thresholds_kernel(const int32_t* __restrict__ gt_ptr, const int32_t* __restrict__ gt_idx,
                                                              const float* __restrict__ gt_val, const int32_t* __restrict__ txt2img,
                                                              int I, int T, int col_offset, float* __restrict__ cand_score,
                                                              int32_t* __restrict__ cand_idx, int32_t* __restrict__ col_thr_idx) {
  pdl_enter();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < I) {
    float bs = INFINITY;
    int bc = -1;
    for (int e = gt_ptr[t]; e < gt_ptr[t + 1]; ++e) {
      const int cg = gt_idx[e], c = cg - col_offset;
      if (c < 0 || c >= T) continue;
      const float sc = gt_val[e];
      if (bc < 0 || sc > bs || (sc == bs && cg < bc)) { bs = sc; bc = cg; }
    }
    cand_score[t] = bs;
    cand_idx[t] = bc;
  }
  if (t < T) {
    const int g = txt2img[t];
    col_thr_idx[t] = (g >= 0 && g < I) ? g : -1;
  }
}
__global__ void __launch_bounds__(256) rank_finalize_kernel(const float* __restrict__ row_thr, const int32_t* __restrict__ col_thr_idx,
                                                            int I, int T, int invalid_row_rank, int32_t* __restrict__ ranks_i2t,
                                                            int32_t* __restrict__ ranks_t2i) {
  pdl_enter();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < I && row_thr[t] == INFINITY) ranks_i2t[t] = invalid_row_rank;      // no valid ground truth: never retrieved
  if (t < T && col_thr_idx[t] < 0) ranks_t2i[t] = I;
}

// ---- screened count pass: bf16x3 screen, exact 3xTF32 decision of the borderline pairs ----------------------------------
// x = hi + lo + r with hi = bf16(x), lo = bf16(x - hi): |r| <= 2^-18 |x|.  The screen computes sum (hi hi + hi lo + lo hi): it
// drops lo lo and the r terms (<= 3 * 2^-18 sum |x_k y_k|) and accumulates in the tensor core's truncating fp32 (<= 2^-23 of
// the running magnitude per MMA: K/16 * 3 MMAs here, K/8 * 3 in the 3xTF32 product it is compared with).  With
// sum |x_k y_k| <= |x| |y| the two scores of a pair differ by at most  alpha * eps_rel(K) * |x| |y|;  the factor 2 is margin.
inline float screen_eps_rel(int D) { return 2.0f * (3.0f / 262144.0f + (float)D * (3.0f / 16.0f + 3.0f / 8.0f) / 8388608.0f); }

// (max_norm: one int-punned float per side, atomicMax -- norms are non-negative, so integer order == float order)
__global__ void __launch_bounds__(256) split_rows_bf16_kernel(const float* __restrict__ x, int D, uint16_t* __restrict__ hi,
                                                              uint16_t* __restrict__ lo, int* __restrict__ max_norm) {
  pdl_enter();
  __shared__ float scratch[34];
  const size_t base = (size_t)blockIdx.x * D;
  float ss = 0.f;
  for (int j = threadIdx.x; j < D; j += blockDim.x) {
    const float v = x[base + j];
    const uint32_t u = __float_as_uint(v);
    const uint32_t hr = (u + 0x7FFFu + ((u >> 16) & 1u)) & 0xFFFF0000u;             // round to nearest even (finite inputs)
    const float h = __uint_as_float(hr);
    const float l = v - h;                                                             // exact
    const uint32_t ul = __float_as_uint(l);
    const uint32_t lr = (ul + 0x7FFFu + ((ul >> 16) & 1u)) & 0xFFFF0000u;
    hi[base + j] = (uint16_t)(hr >> 16);
    lo[base + j] = (uint16_t)(lr >> 16);
    ss = fmaf(v, v, ss);
  }
  ss = block_sum<float>(ss, scratch);
  if (threadIdx.x == 0) atomicMax(max_norm, __float_as_int(sqrtf(ss) * 1.000001f));
}
// the same split, one warp per row, 8 elements per lane and trip: two 128-bit loads, one 128-bit store per output (D % 8 == 0,
// 16-byte aligned rows)
__device__ __forceinline__ void split_bf16(float v, uint32_t& h16, uint32_t& l16) {
  const uint32_t u = __float_as_uint(v);
  const uint32_t hr = (u + 0x7FFFu + ((u >> 16) & 1u)) & 0xFFFF0000u;
  const float l = v - __uint_as_float(hr);
  const uint32_t ul = __float_as_uint(l);
  h16 = hr >> 16;
  l16 = ((ul + 0x7FFFu + ((ul >> 16) & 1u)) >> 16) & 0xFFFFu;
}
__global__ void __launch_bounds__(256) split_rows_bf16_v8_kernel(const float* __restrict__ x, int rows, int D,
                                                                 uint16_t* __restrict__ hi, uint16_t* __restrict__ lo,
                                                                 int* __restrict__ max_norm) {
  pdl_enter();
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  float worst = 0.f;
  for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += gridDim.x * wpb) {
    const size_t base = (size_t)row * D;
    float ss = 0.f;
    for (int j = lane * 8; j < D; j += 256) {
      const float4 a = ldg_stream4(x + base + j), b = ldg_stream4(x + base + j + 4);
      const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
      uint32_t h[8], l[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        split_bf16(v[i], h[i], l[i]);
        ss = fmaf(v[i], v[i], ss);
      }
      *reinterpret_cast<uint4*>(hi + base + j) = make_uint4(h[0] | (h[1] << 16), h[2] | (h[3] << 16), h[4] | (h[5] << 16), h[6] | (h[7] << 16));
      *reinterpret_cast<uint4*>(lo + base + j) = make_uint4(l[0] | (l[1] << 16), l[2] | (l[3] << 16), l[4] | (l[5] << 16), l[6] | (l[7] << 16));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    worst = fmaxf(worst, ss);
  }
  // (the per-row sum runs in a different order than the block kernel's: the bound it feeds carries a 1e-6 relative margin)
  if (lane == 0 && worst > 0.f) atomicMax(max_norm, __float_as_int(sqrtf(worst) * 1.000001f));
}
inline void split_rows_bf16(const float* x, int rows, int D, uint16_t* hi, uint16_t* lo, int* max_norm, cudaStream_t st) {
  const bool v8 = D % 8 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(hi) | reinterpret_cast<uintptr_t>(lo)) & 15) == 0;
  if (v8) launch_k(split_rows_bf16_v8_kernel, std::min(ceil_div(rows, 8), num_sms() * 16), 256, 0, st, x, rows, D, hi, lo, max_norm);
  else launch_k(split_rows_bf16_kernel, rows, 256, 0, st, x, D, hi, lo, max_norm);
}
__global__ void screen_band_kernel(const int* __restrict__ max_norms, float eps_alpha, float* __restrict__ band) {
  pdl_enter();
  if (threadIdx.x == 0) *band = eps_alpha * __int_as_float(max_norms[0]) * __int_as_float(max_norms[1]) * 1.000002f;
}
// n_pairs = borderline pairs to decide (0 when the list overflowed: the exact pass then recounts everything), the diagonal
// tiles of the gathered product, and the work count of the exact fall-back pass
__global__ void __launch_bounds__(256) decide_prepare_kernel(const int32_t* __restrict__ amb_count, int cap, int tiles_n_gather,
                                                             int exact_tiles, int32_t* __restrict__ n_pairs,
                                                             int* __restrict__ tile_list, int* __restrict__ tile_count,
                                                             int* __restrict__ ident_list, int* __restrict__ fb_count) {
  pdl_enter();
  const int cnt = *amb_count;
  const bool overflow = cnt > cap;
  const int n = overflow ? 0 : cnt;
  const int nt = (n + 127) / 128;
  const int t = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
  for (int i = t; i < nt; i += stride) tile_list[i] = i * (tiles_n_gather + 1);
  for (int i = t; i < exact_tiles; i += stride) ident_list[i] = i;
  if (t == 0) { *n_pairs = n; *tile_count = nt; *fb_count = overflow ? exact_tiles : 0; }
}
__global__ void __launch_bounds__(256) gather_pairs_kernel(const tc::AmbiguousPair* __restrict__ list, const int32_t* __restrict__ n_pairs,
                                                           const float* __restrict__ img, const float* __restrict__ txt, int D,
                                                           float* __restrict__ Ag, float* __restrict__ Bg) {
  pdl_enter();
  const int n = *n_pairs;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int p = warp; p < n; p += nwarps) {
    const tc::AmbiguousPair a = list[p];
    const float4* si = reinterpret_cast<const float4*>(img + (size_t)a.m * D);
    const float4* sj = reinterpret_cast<const float4*>(txt + (size_t)a.n * D);
    float4* di = reinterpret_cast<float4*>(Ag + (size_t)p * D);
    float4* dj = reinterpret_cast<float4*>(Bg + (size_t)p * D);
    for (int k = lane; k < (D >> 2); k += 32) { di[k] = si[k]; dj[k] = sj[k]; }
  }
}
// the borderline list overflowed: forget the screen's counts, the exact pass recounts every tile
__global__ void __launch_bounds__(256) fallback_reset_kernel(const int* __restrict__ fb_count, int I, int T, int32_t* __restrict__ ranks_i2t,
                                                             int32_t* __restrict__ ranks_t2i) {
  pdl_enter();
  if (*fb_count == 0) return;
  const int t = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
  for (int i = t; i < I; i += stride) ranks_i2t[i] = 0;
  for (int i = t; i < T; i += stride) ranks_t2i[i] = 0;
}

// borderline-list capacity: pairs / 2048, at least 4096, at most what 2 GB of gathered rows hold (VLDD_SCREEN_CAP: test hook)
int screen_cap(int I, int T, int D) {
  static int forced = -1;
  if (forced < 0) { const char* e = getenv("VLDD_SCREEN_CAP"); forced = e ? atoi(e) : 0; }
  long long cap = forced > 0 ? forced : (long long)I * T / 2048;
  const long long hi = (1ll << 31) / (8ll * D), top = 1ll << 20;
  if (forced <= 0 && cap < 4096) cap = 4096;
  if (cap > hi) cap = hi;
  if (cap > top) cap = top;
  if (cap < 128) cap = 128;
  return (int)((cap + 127) / 128 * 128);
}
// the screen pays for its extra launches (two splits, prepare, gather, decide, fall-back) from ~4 M pairs on
bool screen_wanted(const float* img, const float* txt, int I, int T, int D) {
  static long long min_pairs = -1;
  if (min_pairs < 0) { const char* e = getenv("VLDD_SCREEN_MIN_PAIRS"); min_pairs = e ? atoll(e) : 4000000ll; }
  return D % 8 == 0 && aligned16(img) && aligned16(txt) && (long long)I * T >= min_pairs;
}

namespace {
struct FusedWs {
  float *gt_val, *col_val, *row_thr; int32_t *row_thr_idx, *col_thr_idx; int *flags, *list, *count;
  // screened pass
  uint16_t *img_hi, *img_lo, *txt_hi, *txt_lo; int* max_norms; float* band; tc::AmbiguousPair* amb; int32_t *amb_count, *n_pairs;
  int *tile_list, *tile_count, *ident_list, *fb_count; float *Ag, *Bg;
  size_t bytes;
};
void carve_fused(FusedWs& w, void* base, int I, int T, int D, int nnz, bool screen) {
  auto al = [](size_t b) { return (b + 255) / 256 * 256; };
  char* p = reinterpret_cast<char*>(base);
  const size_t tiles = (size_t)ceil_div(I, tc::BM) * ceil_div(T, 128);
  auto take = [&](size_t b) { char* r = p; p += al(b); return r; };
  w.gt_val = (float*)take((size_t)nnz * 4 + 4);
  w.col_val = (float*)take((size_t)T * 4);
  w.row_thr = (float*)take((size_t)I * 4);
  w.row_thr_idx = (int32_t*)take((size_t)I * 4);
  w.col_thr_idx = (int32_t*)take((size_t)T * 4);
  w.flags = (int*)take(tiles * 4);
  w.list = (int*)take(tiles * 4);
  w.count = (int*)take(256);
  if (screen) {
    const int cap = screen_cap(I, T, D);
    w.img_hi = (uint16_t*)take((size_t)I * D * 2); w.img_lo = (uint16_t*)take((size_t)I * D * 2);
    w.txt_hi = (uint16_t*)take((size_t)T * D * 2); w.txt_lo = (uint16_t*)take((size_t)T * D * 2);
    w.amb = (tc::AmbiguousPair*)take((size_t)cap * sizeof(tc::AmbiguousPair));
    w.amb_count = (int32_t*)take(256); w.n_pairs = w.amb_count + 1; w.tile_count = (int*)(w.amb_count + 2); w.fb_count = (int*)(w.amb_count + 3);
    w.max_norms = (int*)(w.amb_count + 4); w.band = (float*)(w.amb_count + 6);
    w.tile_list = (int*)take((size_t)(cap / 128) * 4);
    w.ident_list = (int*)take(tiles * 4);
    w.Ag = (float*)take((size_t)cap * D * 4); w.Bg = (float*)take((size_t)cap * D * 4);
  }
  w.bytes = (size_t)(p - reinterpret_cast<char*>(base));
}
}  // namespace

size_t sim_rank_fused_workspace_bytes(int I, int T, int D, int nnz) {
  FusedWs w;
  carve_fused(w, nullptr, I, T, D, nnz, D % 8 == 0);       // sized for the screened pass whenever the shape allows it
  return w.bytes + 256;
}

bool sim_rank_fused_ok(const float* img, const float* txt, int I, int T, int D) {
  return tc_enabled() && tc::gemm_ok<true, true>(gemm_ops(img, D, txt, D, I, T, D));
}

// Phase A: the 3xTF32 scores at the ground-truth positions (only the tiles that hold one are computed) -> per image the best
// local candidate (score, global caption index); the per-caption thresholds stay in the workspace for phase B.
int sim_rank_fused_candidates(const float* img, const float* txt, int I, int T, int D, float scale, const int32_t* txt2img,
                              const int32_t* gt_ptr, const int32_t* gt_idx, int nnz, int col_offset, float* cand_score,
                              int32_t* cand_idx, void* workspace, cudaStream_t st) {
  const int tiles_m = ceil_div(I, tc::BM), tiles_n = ceil_div(T, 128), tiles = tiles_m * tiles_n;
  FusedWs w;
  carve_fused(w, workspace, I, T, D, nnz, D % 8 == 0);
  VLDD_CUDA(cudaMemsetAsync(w.flags, 0, (size_t)tiles * 4, st));
  VLDD_CUDA(cudaMemsetAsync(w.count, 0, 4, st));
  const int nmax = I > T ? I : T;
  launch_k(mark_gt_tiles_kernel, ceil_div(nmax, 256), 256, 0, st, gt_ptr, gt_idx, txt2img, I, T, tiles_n, col_offset, w.flags);
  launch_k(compact_tiles_kernel, ceil_div(tiles, 256), 256, 0, st, (const int*)w.flags, tiles, w.list, w.count);
  const GemmOperands g = gemm_ops(img, D, txt, D, I, T, D);
  int rc = tc::launch<true, true, 3>(g, 1, tc::EpiRankExtract{scale, gt_ptr, gt_idx, w.gt_val, txt2img, w.col_val, col_offset}, st, w.list,
                                     w.count);
  if (rc) return rc;
  launch_k(rank_thresholds_kernel, ceil_div(nmax, 256), 256, 0, st, gt_ptr, gt_idx, (const float*)w.gt_val, txt2img, I, T, col_offset,
           cand_score, cand_idx, w.col_thr_idx);
  return check_launch("sim_rank_fused_candidates");
}

// Phase B: per image the number of LOCAL captions ranked ahead of its threshold (score, caption index in local numbering -- any
// integer: a threshold that lives in another shard simply never wins or loses the index tie-break), per local caption its final
// text -> image rank.  Same workspace as phase A.  Rows whose threshold is +inf get `invalid_row_rank`.
int sim_rank_fused_count(const float* img, const float* txt, int I, int T, int D, float scale, const float* thr_score,
                         const int32_t* thr_idx_local, int nnz, int invalid_row_rank, int32_t* ranks_i2t, int32_t* ranks_t2i,
                         void* workspace, cudaStream_t st) {
  const int tiles_m = ceil_div(I, tc::BM), tiles_n = ceil_div(T, 128), tiles = tiles_m * tiles_n;
  const bool screen = screen_wanted(img, txt, I, T, D);
  FusedWs w;
  carve_fused(w, workspace, I, T, D, nnz, D % 8 == 0);
  VLDD_CUDA(cudaMemsetAsync(ranks_i2t, 0, (size_t)I * 4, st));
  VLDD_CUDA(cudaMemsetAsync(ranks_t2i, 0, (size_t)T * 4, st));
  const int nmax = I > T ? I : T;
  const GemmOperands g = gemm_ops(img, D, txt, D, I, T, D);
  int rc;
  if (!screen) {
    // exact: every tile, count the entries ranked ahead of the ground truth per row and per column
    rc = tc::launch<true, true, 3>(g, 1, tc::EpiRankCount{scale, thr_score, thr_idx_local, ranks_i2t, w.col_val, w.col_thr_idx, ranks_t2i}, st);
    if (rc) return rc;
  } else {
    // screen: the same count from a bf16x3 product at half the tensor time; pairs within the error band of their threshold go
    // to a list ...
    const int cap = screen_cap(I, T, D);
    VLDD_CUDA(cudaMemsetAsync(w.amb_count, 0, 32, st));          // list count, pair count, tile count, fall-back count, max norms, band
    split_rows_bf16(img, I, D, w.img_hi, w.img_lo, w.max_norms, st);
    split_rows_bf16(txt, T, D, w.txt_hi, w.txt_lo, w.max_norms + 1, st);
    launch_k(screen_band_kernel, 1, 32, 0, st, (const int*)w.max_norms, screen_eps_rel(D) * fabsf(scale), w.band);
    rc = tc::launch_bf16x3<tc::EpiRankScreen, 256>(
        w.img_hi, w.img_lo, w.txt_hi, w.txt_lo, I, T, D,
        tc::EpiRankScreen{scale, thr_score, thr_idx_local, ranks_i2t, w.col_val, w.col_thr_idx, ranks_t2i, w.band, w.amb, w.amb_count, cap},
        st);
    if (rc) return rc;
    // ... decided exactly: the listed pairs are gathered into a [P, D] x [P, D]^T problem whose diagonal tiles the 3xTF32 kernel
    // computes -- the same arithmetic, k order and epilogue scaling as phase A, hence the same bits as the materialised matrix
    launch_k(decide_prepare_kernel, 64, 256, 0, st, (const int32_t*)w.amb_count, cap, cap / 128, tiles, w.n_pairs, w.tile_list,
             w.tile_count, w.ident_list, w.fb_count);
    launch_k(gather_pairs_kernel, num_sms() * 4, 256, 0, st, (const tc::AmbiguousPair*)w.amb, (const int32_t*)w.n_pairs, img, txt, D, w.Ag,
             w.Bg);
    const GemmOperands gg = gemm_ops(w.Ag, D, w.Bg, D, cap, cap, D);
    rc = tc::launch<true, true, 3>(gg, 1,
                                   tc::EpiPairDecide{scale, w.amb, w.n_pairs, thr_score, thr_idx_local, ranks_i2t, w.col_val, w.col_thr_idx,
                                                     ranks_t2i},
                                   st, w.tile_list, w.tile_count);
    if (rc) return rc;
    // list overflow (far more near-ties than the capacity provides for): recount everything exactly
    launch_k(fallback_reset_kernel, num_sms(), 256, 0, st, (const int*)w.fb_count, I, T, ranks_i2t, ranks_t2i);
    rc = tc::launch<true, true, 3>(g, 1, tc::EpiRankCount{scale, thr_score, thr_idx_local, ranks_i2t, w.col_val, w.col_thr_idx, ranks_t2i}, st,
                                   w.ident_list, w.fb_count);
    if (rc) return rc;
  }
  launch_k(rank_finalize_kernel, ceil_div(nmax, 256), 256, 0, st, thr_score, (const int32_t*)w.col_thr_idx, I, T, invalid_row_rank, ranks_i2t,
           ranks_t2i);
  return check_launch("sim_rank_fused_count");
}

int sim_rank_fused(const float* img, const float* txt, int I, int T, int D, float scale, const int32_t* txt2img,
                   const int32_t* gt_ptr, const int32_t* gt_idx, int nnz, int32_t* ranks_i2t, int32_t* ranks_t2i,
                   void* workspace, cudaStream_t st) {
  FusedWs w;
  carve_fused(w, workspace, I, T, D, nnz, D % 8 == 0);
  int rc = sim_rank_fused_candidates(img, txt, I, T, D, scale, txt2img, gt_ptr, gt_idx, nnz, 0, w.row_thr, w.row_thr_idx, workspace, st);
  if (rc) return rc;
  return sim_rank_fused_count(img, txt, I, T, D, scale, w.row_thr, w.row_thr_idx, nnz, T, ranks_i2t, ranks_t2i, workspace, st);
}

// Keep each row's k largest entries, everything else := fill (epoch_original.py:95-105).
// Radix select on the order-preserving integer image of the float, 4 passes of 8 bits, one CTA per row;
// among entries equal to the k-th value the lower column indices are kept (stable, as the oracle does).
__device__ __forceinline__ unsigned int float_key(float f) {
  const unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);   // larger float -> larger key
}

__global__ void __launch_bounds__(256) topk_fill_rows_kernel(const float* __restrict__ S, float* __restrict__ out,
                                                             int ncols, int k, float fill) {
  pdl_enter();
  __shared__ int hist[256];
  __shared__ unsigned int sel_prefix;
  __shared__ int sel_remaining;
  __shared__ int tie_base[257];
  const float* __restrict__ row = S + (size_t)blockIdx.x * ncols;
  float* __restrict__ orow = out + (size_t)blockIdx.x * ncols;
  if (k >= ncols) {
    for (int j = threadIdx.x; j < ncols; j += blockDim.x) orow[j] = row[j];
    return;
  }
  unsigned int prefix = 0, mask = 0;
  int remaining = k;  // how many still to take among keys matching the prefix
  for (int pass = 3; pass >= 0; --pass) {
    const int shift = pass * 8;
    hist[threadIdx.x] = 0;
    __syncthreads();
    for (int j = threadIdx.x; j < ncols; j += blockDim.x) {
      const unsigned int key = float_key(row[j]);
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int rem = remaining, b = 255;
      for (; b > 0; --b) {
        if (hist[b] >= rem) break;
        rem -= hist[b];
      }
      sel_prefix = prefix | ((unsigned int)b << shift);
      sel_remaining = rem;
    }
    __syncthreads();
    prefix = sel_prefix;
    remaining = sel_remaining;
    mask |= 255u << shift;
    __syncthreads();
  }
  // prefix == key of the k-th largest; take every key > prefix and the first `remaining` (by column) keys == prefix
  const int chunk = (ncols + blockDim.x - 1) / blockDim.x;
  const int j0 = threadIdx.x * chunk, j1 = min(ncols, j0 + chunk);
  int ties = 0;
  for (int j = j0; j < j1; ++j) ties += float_key(row[j]) == prefix;
  tie_base[threadIdx.x + 1] = ties;
  if (threadIdx.x == 0) tie_base[0] = 0;
  __syncthreads();
  if (threadIdx.x == 0)
    for (int t = 1; t <= (int)blockDim.x; ++t) tie_base[t] += tie_base[t - 1];
  __syncthreads();
  int seen = tie_base[threadIdx.x];
  for (int j = j0; j < j1; ++j) {
    const float v = row[j];
    const unsigned int key = float_key(v);
    bool keep = key > prefix;
    if (key == prefix) { keep = seen < remaining; ++seen; }
    orow[j] = keep ? v : fill;
  }
}

int topk_fill_rows(const float* S, float* out, int nrows, int ncols, int k, float fill, cudaStream_t st) {
  if (nrows <= 0 || ncols <= 0) return VLDD_OK;
  launch_k(topk_fill_rows_kernel, nrows, 256, 0, st, S, out, ncols, k, fill);
  return check_launch("topk_fill_rows");
}

}  // namespace vldd
