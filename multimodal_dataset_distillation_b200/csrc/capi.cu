// extern "C" surface of libvldd_b200.so (declared in include/vldd_b200.h).  Thin: argument checks, then the
// C++ launchers.  No torch types cross this boundary.
#include <mutex>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/vldd_b200.h"
#include "common.cuh"
#include "engine.h"
#include "kernels.h"

namespace vldd {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("VLDD_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: kernel launch failed: %s", what, cudaGetErrorString(e));
    return VLDD_ERR_CUDA;
  }
  return VLDD_OK;
}

// grow-only device scratch for the *_host entry points
static std::mutex g_scratch_mu;
static void* g_scratch = nullptr;
static size_t g_scratch_bytes = 0;

static int host_scratch(size_t bytes, void** out) {
  if (bytes > g_scratch_bytes) {
    if (g_scratch) cudaFree(g_scratch);
    g_scratch = nullptr;
    g_scratch_bytes = 0;
    size_t want = bytes + bytes / 8;
    cudaError_t e = cudaMalloc(&g_scratch, want);
    if (e != cudaSuccess) {
      set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
      return VLDD_ERR_CUDA;
    }
    g_scratch_bytes = want;
  }
  *out = g_scratch;
  return VLDD_OK;
}

}  // namespace vldd

using namespace vldd;

static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

extern "C" {

int vldd_version(void) { return 100; }
const char* vldd_last_error(void) { return g_err; }
unsigned long long vldd_kernel_launch_count(void) { return launch_counter().load(); }

int vldd_flat_sgd_step(const float* theta, const float* grad, const float* lr, float* out, int64_t n, void* stream) {
  VLDD_REQUIRE(n >= 0 && (n == 0 || (theta && grad && lr && out)), "flat_sgd_step: null pointer or negative n");
  return flat_sgd_step(theta, grad, lr, out, n, S(stream));
}

size_t vldd_match_loss_scratch_bytes(void) { return (size_t)match_loss_scratch_bytes(); }

int vldd_match_loss_fwd(const float* theta_K, const float* theta_tgt, const float* theta_0, int64_t n, float* out3,
                        void* scratch, void* stream) {
  VLDD_REQUIRE(n >= 0 && theta_K && theta_tgt && theta_0 && out3 && scratch, "match_loss_fwd: null pointer");
  return match_loss_fwd(theta_K, theta_tgt, theta_0, n, out3, scratch, S(stream));
}

int vldd_match_loss_bwd(const float* theta_K, const float* theta_tgt, const float* num_den, const float* gout, float* a,
                        int64_t n, void* stream) {
  VLDD_REQUIRE(n >= 0 && theta_K && theta_tgt && num_den && a, "match_loss_bwd: null pointer");
  return match_loss_bwd(theta_K, theta_tgt, num_den, gout, a, n, S(stream));
}

int vldd_momentum_sgd(float* p, const float* g, float* buf, float lr, float momentum, int first, int64_t n,
                      void* stream) {
  VLDD_REQUIRE(n >= 0 && (n == 0 || (p && g && buf)), "momentum_sgd: null pointer");
  return momentum_sgd(p, g, buf, lr, momentum, first, n, S(stream));
}

int vldd_ranks_from_scores(const float* scores_i2t, const float* scores_t2i, int n_img, int n_txt,
                           const int32_t* txt2img, const int32_t* img2txt_ptr, const int32_t* img2txt_idx,
                           int32_t* ranks_i2t, int32_t* ranks_t2i, void* stream) {
  VLDD_REQUIRE(n_img >= 0 && n_txt >= 0, "ranks_from_scores: negative sizes");
  if (scores_i2t) {
    VLDD_REQUIRE(img2txt_ptr && img2txt_idx && ranks_i2t, "ranks_from_scores: i2t needs img2txt CSR and ranks_i2t");
    int rc = ranks_rows(scores_i2t, n_txt, n_img, n_txt, img2txt_ptr, img2txt_idx, ranks_i2t, S(stream));
    if (rc) return rc;
  }
  if (scores_t2i) {
    VLDD_REQUIRE(txt2img && ranks_t2i, "ranks_from_scores: t2i needs txt2img and ranks_t2i");
    int rc = ranks_rows(scores_t2i, n_img, n_txt, n_img, nullptr, txt2img, ranks_t2i, S(stream));
    if (rc) return rc;
  }
  return VLDD_OK;
}

int vldd_ranks_cols(const float* scores_i2t, int n_img, int n_txt, const int32_t* txt2img, int32_t* ranks_t2i, void* stream) {
  VLDD_REQUIRE(n_img >= 0 && n_txt >= 0 && (n_txt == 0 || (scores_i2t && txt2img && ranks_t2i)), "ranks_cols: bad arguments");
  return ranks_cols(scores_i2t, n_txt, n_img, n_txt, txt2img, ranks_t2i, S(stream));
}

int vldd_rank_best_gt(const float* scores, int rows, int cols, int col_offset, const int32_t* gt_ptr,
                      const int32_t* gt_idx, float* best_score, int32_t* best_idx, void* stream) {
  VLDD_REQUIRE(rows >= 0 && cols >= 0 && (rows == 0 || (scores && gt_ptr && gt_idx && best_score && best_idx)),
               "rank_best_gt: bad arguments");
  return best_gt_rows(scores, cols, rows, cols, col_offset, gt_ptr, gt_idx, best_score, best_idx, S(stream));
}

int vldd_rank_count(const float* scores, int rows, int cols, int col_offset, const float* thr_score,
                    const int32_t* thr_idx, int32_t* counts, void* stream) {
  VLDD_REQUIRE(rows >= 0 && cols >= 0 && (rows == 0 || (scores && thr_score && thr_idx && counts)),
               "rank_count: bad arguments");
  return count_rows(scores, cols, rows, cols, col_offset, thr_score, thr_idx, counts, S(stream));
}

int vldd_recall_counts(const int32_t* ranks, int n, int32_t* counts3, void* stream) {
  VLDD_REQUIRE(n >= 0 && counts3 && (n == 0 || ranks), "recall_counts: null pointer");
  return recall_counts(ranks, n, counts3, S(stream));
}

int vldd_sim_scores(const float* img, const float* txt, int n_img, int n_txt, int dim, float scale, float* scores_i2t,
                    float* scores_t2i, void* stream) {
  VLDD_REQUIRE(n_img >= 0 && n_txt >= 0 && dim > 0 && img && txt, "sim_scores: bad arguments");
  return sim_scores(img, txt, n_img, n_txt, dim, scale, scores_i2t, scores_t2i, S(stream));
}

int vldd_topk_fill(const float* scores, float* out, int rows, int cols, int k, float fill, void* stream) {
  VLDD_REQUIRE(rows >= 0 && cols >= 0 && k >= 0 && scores && out && scores != out, "topk_fill: bad arguments");
  return topk_fill_rows(scores, out, rows, cols, k, fill, S(stream));
}

size_t vldd_sim_rank_workspace_bytes(int n_img, int n_txt, int dim) {
  (void)dim;
  if (n_img <= 0 || n_txt <= 0) return 256;
  return (size_t)n_img * (size_t)n_txt * sizeof(float) + 256;     // one [I,T] score matrix; the transpose is never built
}

int vldd_sim_rank(const float* img, const float* txt, int n_img, int n_txt, int dim, float scale,
                  const int32_t* txt2img, const int32_t* img2txt_ptr, const int32_t* img2txt_idx, int32_t* ranks_i2t,
                  int32_t* ranks_t2i, void* workspace, size_t workspace_bytes, void* stream) {
  VLDD_REQUIRE(n_img >= 0 && n_txt >= 0 && dim > 0 && img && txt, "sim_rank: bad arguments");
  VLDD_REQUIRE(txt2img && img2txt_ptr && img2txt_idx && ranks_i2t && ranks_t2i, "sim_rank: null ground truth / outputs");
  if (workspace == nullptr || workspace_bytes < vldd_sim_rank_workspace_bytes(n_img, n_txt, dim)) {
    set_error("sim_rank: workspace too small (need %zu bytes)", vldd_sim_rank_workspace_bytes(n_img, n_txt, dim));
    return VLDD_ERR_WORKSPACE;
  }
  if (n_img == 0 || n_txt == 0) return VLDD_OK;
  float* s_i2t = reinterpret_cast<float*>(workspace);
  int rc = sim_scores(img, txt, n_img, n_txt, dim, scale, s_i2t, nullptr, S(stream));
  if (rc) return rc;
  rc = ranks_rows(s_i2t, n_txt, n_img, n_txt, img2txt_ptr, img2txt_idx, ranks_i2t, S(stream));   // image -> text: rows
  if (rc) return rc;
  return ranks_cols(s_i2t, n_txt, n_img, n_txt, txt2img, ranks_t2i, S(stream));                   // text -> image: columns
}

size_t vldd_sim_rank_fused_workspace_bytes(int n_img, int n_txt, int dim, int nnz) {
  if (n_img <= 0 || n_txt <= 0 || nnz < 0) return 256;
  return sim_rank_fused_workspace_bytes(n_img, n_txt, dim, nnz);
}

int vldd_sim_rank_fused(const float* img, const float* txt, int n_img, int n_txt, int dim, float scale,
                        const int32_t* txt2img, const int32_t* img2txt_ptr, const int32_t* img2txt_idx, int nnz,
                        int32_t* ranks_i2t, int32_t* ranks_t2i, void* workspace, size_t workspace_bytes, void* stream) {
  VLDD_REQUIRE(n_img >= 0 && n_txt >= 0 && dim > 0 && nnz >= 0 && img && txt, "sim_rank_fused: bad arguments");
  VLDD_REQUIRE(txt2img && img2txt_ptr && img2txt_idx && ranks_i2t && ranks_t2i, "sim_rank_fused: null ground truth / outputs");
  if (n_img == 0 || n_txt == 0) return VLDD_OK;
  VLDD_REQUIRE(sim_rank_fused_ok(img, txt, n_img, n_txt, dim),
               "sim_rank_fused: operands do not satisfy the tensor-map constraints (16-byte aligned, dim %% 4 == 0); use "
               "vldd_sim_rank");
  if (workspace == nullptr || workspace_bytes < vldd_sim_rank_fused_workspace_bytes(n_img, n_txt, dim, nnz)) {
    set_error("sim_rank_fused: workspace too small (need %zu bytes)", vldd_sim_rank_fused_workspace_bytes(n_img, n_txt, dim, nnz));
    return VLDD_ERR_WORKSPACE;
  }
  return sim_rank_fused(img, txt, n_img, n_txt, dim, scale, txt2img, img2txt_ptr, img2txt_idx, nnz, ranks_i2t, ranks_t2i,
                        workspace, S(stream));
}

// caption-sharded form of the fused ranking: phase A (candidates) and phase B (counts) with the exchange in between left to
// the caller (dist.sharded_ranks_fused: all-gather of the candidates, merge, all-reduce of the row counts)
int vldd_sim_rank_fused_candidates(const float* img, const float* txt, int n_img, int n_txt, int dim, float scale,
                                   const int32_t* txt2img, const int32_t* img2txt_ptr, const int32_t* img2txt_idx, int nnz,
                                   int col_offset, float* cand_score, int32_t* cand_idx, void* workspace, size_t workspace_bytes,
                                   void* stream) {
  VLDD_REQUIRE(n_img > 0 && n_txt > 0 && dim > 0 && nnz >= 0 && img && txt, "sim_rank_fused_candidates: bad arguments");
  VLDD_REQUIRE(txt2img && img2txt_ptr && img2txt_idx && cand_score && cand_idx, "sim_rank_fused_candidates: null ground truth / outputs");
  VLDD_REQUIRE(sim_rank_fused_ok(img, txt, n_img, n_txt, dim),
               "sim_rank_fused_candidates: operands do not satisfy the tensor-map constraints (16-byte aligned, dim %% 4 == 0)");
  if (workspace == nullptr || workspace_bytes < vldd_sim_rank_fused_workspace_bytes(n_img, n_txt, dim, nnz)) {
    set_error("sim_rank_fused_candidates: workspace too small (need %zu bytes)", vldd_sim_rank_fused_workspace_bytes(n_img, n_txt, dim, nnz));
    return VLDD_ERR_WORKSPACE;
  }
  return sim_rank_fused_candidates(img, txt, n_img, n_txt, dim, scale, txt2img, img2txt_ptr, img2txt_idx, nnz, col_offset, cand_score,
                                   cand_idx, workspace, S(stream));
}

int vldd_sim_rank_fused_count(const float* img, const float* txt, int n_img, int n_txt, int dim, float scale, const float* thr_score,
                              const int32_t* thr_idx_local, int nnz, int invalid_row_rank, int32_t* row_counts, int32_t* ranks_t2i,
                              void* workspace, size_t workspace_bytes, void* stream) {
  VLDD_REQUIRE(n_img > 0 && n_txt > 0 && dim > 0 && nnz >= 0 && img && txt && thr_score && thr_idx_local && row_counts && ranks_t2i,
               "sim_rank_fused_count: bad arguments");
  if (workspace == nullptr || workspace_bytes < vldd_sim_rank_fused_workspace_bytes(n_img, n_txt, dim, nnz)) {
    set_error("sim_rank_fused_count: workspace too small (need %zu bytes)", vldd_sim_rank_fused_workspace_bytes(n_img, n_txt, dim, nnz));
    return VLDD_ERR_WORKSPACE;
  }
  return sim_rank_fused_count(img, txt, n_img, n_txt, dim, scale, thr_score, thr_idx_local, nnz, invalid_row_rank, row_counts, ranks_t2i,
                              workspace, S(stream));
}

int vldd_itm_eval_host(const float* scores_i2t_host, const float* scores_t2i_host, int n_img, int n_txt,
                       const int32_t* txt2img_host, const int32_t* img2txt_ptr_host, const int32_t* img2txt_idx_host,
                       int32_t* ranks_i2t_host, int32_t* ranks_t2i_host, double* result9, void* stream) {
  VLDD_REQUIRE(n_img > 0 && n_txt > 0, "itm_eval: empty score matrix (%d x %d)", n_img, n_txt);
  VLDD_REQUIRE(scores_i2t_host && scores_t2i_host && txt2img_host && img2txt_ptr_host && img2txt_idx_host && result9,
               "itm_eval: null pointer");
  std::lock_guard<std::mutex> lock(g_scratch_mu);
  cudaStream_t st = S(stream);
  const size_t IT = (size_t)n_img * n_txt;
  const int nnz = img2txt_ptr_host[n_img];
  VLDD_REQUIRE(nnz >= 0, "itm_eval: bad img2txt CSR");
  auto al = [](size_t b) { return (b + 255) / 256 * 256; };
  const size_t o_s1 = 0, o_s2 = o_s1 + al(IT * 4), o_t2i = o_s2 + al(IT * 4), o_ptr = o_t2i + al((size_t)n_txt * 4),
               o_idx = o_ptr + al((size_t)(n_img + 1) * 4), o_ri = o_idx + al((size_t)nnz * 4 + 4),
               o_rt = o_ri + al((size_t)n_img * 4), o_cnt = o_rt + al((size_t)n_txt * 4), total = o_cnt + 256;
  void* base = nullptr;
  int rc = host_scratch(total, &base);
  if (rc) return rc;
  char* b = reinterpret_cast<char*>(base);
  float* d_s1 = reinterpret_cast<float*>(b + o_s1);
  float* d_s2 = reinterpret_cast<float*>(b + o_s2);
  int32_t* d_t2i = reinterpret_cast<int32_t*>(b + o_t2i);
  int32_t* d_ptr = reinterpret_cast<int32_t*>(b + o_ptr);
  int32_t* d_idx = reinterpret_cast<int32_t*>(b + o_idx);
  int32_t* d_ri = reinterpret_cast<int32_t*>(b + o_ri);
  int32_t* d_rt = reinterpret_cast<int32_t*>(b + o_rt);
  int32_t* d_cnt = reinterpret_cast<int32_t*>(b + o_cnt);
  VLDD_CUDA(cudaMemcpyAsync(d_t2i, txt2img_host, (size_t)n_txt * 4, cudaMemcpyHostToDevice, st));
  VLDD_CUDA(cudaMemcpyAsync(d_ptr, img2txt_ptr_host, (size_t)(n_img + 1) * 4, cudaMemcpyHostToDevice, st));
  if (nnz) VLDD_CUDA(cudaMemcpyAsync(d_idx, img2txt_idx_host, (size_t)nnz * 4, cudaMemcpyHostToDevice, st));
  VLDD_CUDA(cudaMemcpyAsync(d_s1, scores_i2t_host, IT * 4, cudaMemcpyHostToDevice, st));
  rc = ranks_rows(d_s1, n_txt, n_img, n_txt, d_ptr, d_idx, d_ri, st);
  if (rc) return rc;
  VLDD_CUDA(cudaMemcpyAsync(d_s2, scores_t2i_host, IT * 4, cudaMemcpyHostToDevice, st));
  rc = ranks_rows(d_s2, n_img, n_txt, n_img, nullptr, d_t2i, d_rt, st);
  if (rc) return rc;
  rc = recall_counts(d_ri, n_img, d_cnt, st);
  if (rc) return rc;
  rc = recall_counts(d_rt, n_txt, d_cnt + 4, st);
  if (rc) return rc;
  int32_t cnt[8];
  VLDD_CUDA(cudaMemcpyAsync(cnt, d_cnt, sizeof(cnt), cudaMemcpyDeviceToHost, st));
  if (ranks_i2t_host) VLDD_CUDA(cudaMemcpyAsync(ranks_i2t_host, d_ri, (size_t)n_img * 4, cudaMemcpyDeviceToHost, st));
  if (ranks_t2i_host) VLDD_CUDA(cudaMemcpyAsync(ranks_t2i_host, d_rt, (size_t)n_txt * 4, cudaMemcpyDeviceToHost, st));
  VLDD_CUDA(cudaStreamSynchronize(st));
  const double tr1 = 100.0 * cnt[0] / n_img, tr5 = 100.0 * cnt[1] / n_img, tr10 = 100.0 * cnt[2] / n_img;
  const double ir1 = 100.0 * cnt[4] / n_txt, ir5 = 100.0 * cnt[5] / n_txt, ir10 = 100.0 * cnt[6] / n_txt;
  const double trm = (tr1 + tr5 + tr10) / 3, irm = (ir1 + ir5 + ir10) / 3;
  result9[0] = tr1; result9[1] = tr5; result9[2] = tr10; result9[3] = trm;
  result9[4] = ir1; result9[5] = ir5; result9[6] = ir10; result9[7] = irm;
  result9[8] = (trm + irm) / 2;
  return VLDD_OK;
}

size_t vldd_proj_head_workspace_bytes(int rows, int dt, int d) { return proj_head_workspace_bytes(rows, dt, d); }

int vldd_proj_head_forward(const float* theta, const float* Y, const float* mask, int rows, int dt, int d, float* z,
                           float* zn, void* workspace, size_t workspace_bytes, void* stream) {
  VLDD_REQUIRE(theta && Y && (z || zn), "proj_head_forward: null pointer");
  return proj_head_forward(theta, Y, mask, rows, dt, d, z, zn, workspace, workspace_bytes, S(stream));
}

size_t vldd_contrastive_step_workspace_bytes(int B, int dt, int d) { return contrastive_step_workspace_bytes(B, dt, d); }

int vldd_contrastive_step(const float* theta, const float* Y, const float* U, const float* scale, const float* mask,
                          int B, int dt, int d, float* loss, float* g_theta, float* dY, float* dU, float* dscale,
                          void* workspace, size_t workspace_bytes, void* stream) {
  VLDD_REQUIRE(theta && Y && U && scale && loss && g_theta, "contrastive_step: null pointer");
  return contrastive_step(theta, Y, U, scale, mask, B, dt, d, loss, g_theta, dY, dU, dscale, workspace, workspace_bytes,
                          S(stream));
}

int vldd_clip_loss(const float* theta, const float* Y, const float* U, const float* scale, const float* mask, int B, int dt,
                   int d, float* loss, int32_t* top1, float* g_theta, float* dY, float* dU, float* dscale, void* workspace,
                   size_t workspace_bytes, void* stream) {
  VLDD_REQUIRE(theta && Y && U && scale && loss && g_theta, "clip_loss: null pointer");
  return clip_loss(theta, Y, U, scale, mask, B, dt, d, loss, top1, g_theta, dY, dU, dscale, workspace, workspace_bytes,
                   S(stream));
}

size_t vldd_infonce_workspace_bytes(int B, int d) { return infonce_workspace_bytes(B, d); }

int vldd_infonce_grad(const float* xn, const float* yn, const float* scale, int B, int d, float* loss, float* dxn, float* dyn,
                      float* dscale, void* workspace, size_t workspace_bytes, void* stream) {
  VLDD_REQUIRE(xn && yn && scale && loss, "infonce_grad: null pointer");
  return infonce_grad(xn, yn, scale, B, d, loss, dxn, dyn, dscale, workspace, workspace_bytes, S(stream));
}

int vldd_infonce_hvp(const float* xn, const float* yn, const float* scale, const float* cx, const float* cy, const float* cs,
                     int B, int d, float* Ldot, float* hx, float* hy, float* hs, void* workspace, size_t workspace_bytes,
                     void* stream) {
  VLDD_REQUIRE(xn && yn && scale && cx && cy && cs && Ldot && hx && hy && hs, "infonce_hvp: null pointer");
  return infonce_hvp(xn, yn, scale, cx, cy, cs, B, d, Ldot, hx, hy, hs, workspace, workspace_bytes, S(stream));
}

size_t vldd_nearest_rows_workspace_bytes(int n_query, int n_bank, int dim) {
  return nearest_rows_workspace_bytes(n_query, n_bank, dim);
}

int vldd_nearest_rows(const float* query, const float* bank, int n_query, int n_bank, int dim, int32_t* idx_out,
                      float* cos_out, void* workspace, size_t workspace_bytes, void* stream) {
  VLDD_REQUIRE(n_query >= 0 && n_bank > 0 && dim > 0, "nearest_rows: bad sizes (n_query=%d n_bank=%d dim=%d)", n_query, n_bank, dim);
  if (n_query == 0) return VLDD_OK;
  VLDD_REQUIRE(query && bank && idx_out && workspace, "nearest_rows: null pointer");
  if (workspace_bytes < nearest_rows_workspace_bytes(n_query, n_bank, dim)) {
    set_error("nearest_rows: workspace too small: need %zu bytes, got %zu", nearest_rows_workspace_bytes(n_query, n_bank, dim),
              workspace_bytes);
    return VLDD_ERR_WORKSPACE;
  }
  return nearest_rows(query, bank, n_query, n_bank, dim, idx_out, cos_out, workspace, S(stream));
}

size_t vldd_bench_skinny_gemm_workspace_bytes(int M, int N, int K) { return skinny_gemm_workspace_bytes(M, N, K); }

int vldd_bench_skinny_gemm(const float* A, const float* W, int M, int N, int K, float* partial, size_t partial_bytes,
                           int* splits_out, void* stream) {
  return skinny_gemm_partial(A, W, M, N, K, partial, partial_bytes, splits_out, S(stream));
}

size_t vldd_unrolled_match_workspace_bytes(int N, int B, int K, int dt, int d) {
  return unrolled_match_workspace_bytes(N, B, K, dt, d);
}

int vldd_unrolled_match(const float* theta0, const float* theta_tgt, const float* Y, const float* U, const float* lr,
                        const float* scale, const int64_t* perms, float* masks, float dropout_p,
                        unsigned long long* rng_state, int N, int B, int K, int dt, int d, float* out5, float* ce, float* dY,
                        float* dU, float* theta_K, void* workspace, size_t workspace_bytes, void* stream) {
  VLDD_REQUIRE(theta0 && theta_tgt && Y && U && lr && scale && (K == 0 || perms) && out5 && dY && dU,
               "unrolled_match: null pointer");
  return unrolled_match(theta0, theta_tgt, Y, U, lr, scale, perms, masks, dropout_p, rng_state, N, B, K, dt, d, out5, ce, dY,
                        dU, theta_K, workspace, workspace_bytes, S(stream));
}

int vldd_match_final(const float* theta_K, const float* theta_tgt, const float* den, int64_t n, float* out3, float* adjoint,
                     void* scratch, void* stream) {
  VLDD_REQUIRE(n > 0 && theta_K && theta_tgt && den && adjoint && scratch, "match_final: null pointer or n <= 0");
  const int rc = match_final_pass(theta_K, theta_tgt, den, n, adjoint, scratch, S(stream));
  if (rc || out3 == nullptr) return rc;         // out3 == NULL: the streaming pass alone (block partials stay in `scratch`)
  return match_final_finish(den, n, out3, scratch, S(stream));
}

int vldd_outer_update(float* U, const float* gU, float* bufU, int64_t nU, float lr_img, float* Y, const float* gY, float* bufY,
                      int64_t nY, float lr_txt, float* syn_lr_img, float* syn_lr_txt, const float* g_lr_img, const float* g_lr_txt,
                      float* buf_lr, float lr_lr, float momentum, int first, float grad_scale, const float* loss, int* skipped,
                      void* stream) {
  VLDD_REQUIRE(nU >= 0 && nY >= 0 && (nU == 0 || (U && gU && bufU)) && (nY == 0 || (Y && gY && bufY)),
               "outer_update: null pointer or negative size");
  VLDD_REQUIRE((syn_lr_img == nullptr && syn_lr_txt == nullptr) || buf_lr != nullptr,
               "outer_update: the student learning rates need their momentum buffer buf_lr[2]");
  return outer_update(U, gU, bufU, nU, lr_img, Y, gY, bufY, nY, lr_txt, syn_lr_img, syn_lr_txt, g_lr_img, g_lr_txt, buf_lr, lr_lr,
                      momentum, first, grad_scale, loss, skipped, S(stream));
}

int vldd_dropout_masks(float* masks, int64_t n, float p, unsigned long long* rng_state, int advance, void* stream) {
  VLDD_REQUIRE(n >= 0 && (n == 0 || masks) && rng_state && p >= 0.f && p < 1.f, "dropout_masks: bad argument");
  return dropout_masks(masks, n, p, rng_state, advance, S(stream));
}

}  // extern "C"
