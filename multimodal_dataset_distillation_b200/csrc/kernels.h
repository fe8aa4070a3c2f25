// Internal C++ launcher declarations (the public C ABI is include/vldd_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace vldd {

// streaming.cu
int flat_sgd_step(const float* theta, const float* grad, const float* lr, float* out, int64_t n, cudaStream_t st);
int64_t match_loss_scratch_bytes();
int match_loss_fwd(const float* thK, const float* tgt, const float* th0, int64_t n, float* out3, void* scratch,
                   cudaStream_t st);
int match_loss_bwd(const float* thK, const float* tgt, const float* num_den, const float* gout, float* a, int64_t n,
                   cudaStream_t st);
int momentum_sgd(float* p, const float* g, float* buf, float lr, float momentum, int first, int64_t n,
                 cudaStream_t st);
int stage_segment(const float* th0_src, const float* tgt_src, float* th0_dst, float* tgt_dst, int64_t n, float* den_out,
                  void* scratch, cudaStream_t st);
// per-call addresses (theta_0, theta*, minibatch indices) reach the replayed launch graph through a 3-pointer device table
int set_stage_sources(const float* th0_src, const float* tgt_src, const void* perms, void* table, void* scratch, cudaStream_t st);
int stage_segment_indirect(const void* table, float* th0_dst, float* tgt_dst, int64_t n, float* den_out, void* scratch,
                           cudaStream_t st);
int match_final_pass(const float* thK, const float* tgt, const float* den, int64_t n, float* a, void* scratch, cudaStream_t st);
int match_final_finish(const float* den, int64_t n, float* out3, void* scratch, cudaStream_t finish_st);
const double* match_final_parts(const void* scratch);   // block partials left by match_final_pass ...
int match_final_n_parts(int64_t n);                     // ... and how many there are
int outer_update(float* U, const float* gU, float* bufU, int64_t nU, float lrU, float* Y, const float* gY, float* bufY,
                 int64_t nY, float lrY, float* syn_lr_img, float* syn_lr_txt, const float* g_lr_img, const float* g_lr_txt,
                 float* buf_lr, float lr_lr, float momentum, int first, float gscale, const float* loss, int* skipped,
                 cudaStream_t st);
int dropout_masks(float* masks, int64_t n, float p, unsigned long long* state, int advance, cudaStream_t st);

// retrieval.cu
int ranks_rows(const float* S, int64_t ld, int nrows, int ncols, const int32_t* gt_ptr, const int32_t* gt_idx,
               int32_t* ranks, cudaStream_t st);
int ranks_cols(const float* S, int64_t ld, int nrows, int ncols, const int32_t* gt_row, int32_t* ranks, cudaStream_t st);
int best_gt_rows(const float* S, int64_t ld, int nrows, int ncols, int col_offset, const int32_t* gt_ptr,
                 const int32_t* gt_idx, float* best_score, int32_t* best_idx, cudaStream_t st);
int count_rows(const float* S, int64_t ld, int nrows, int ncols, int col_offset, const float* thr_score,
               const int32_t* thr_idx, int32_t* counts, cudaStream_t st);
int recall_counts(const int32_t* ranks, int n, int32_t* counts3, cudaStream_t st);
int sim_scores(const float* img, const float* txt, int I, int T, int D, float scale, float* S_i2t, float* S_t2i,
               cudaStream_t st);
size_t sim_rank_fused_workspace_bytes(int I, int T, int D, int nnz);
bool sim_rank_fused_ok(const float* img, const float* txt, int I, int T, int D);
int sim_rank_fused_candidates(const float* img, const float* txt, int I, int T, int D, float scale, const int32_t* txt2img,
                              const int32_t* gt_ptr, const int32_t* gt_idx, int nnz, int col_offset, float* cand_score,
                              int32_t* cand_idx, void* workspace, cudaStream_t st);
int sim_rank_fused_count(const float* img, const float* txt, int I, int T, int D, float scale, const float* thr_score,
                         const int32_t* thr_idx_local, int nnz, int invalid_row_rank, int32_t* ranks_i2t, int32_t* ranks_t2i,
                         void* workspace, cudaStream_t st);
int sim_rank_fused(const float* img, const float* txt, int I, int T, int D, float scale, const int32_t* txt2img,
                   const int32_t* gt_ptr, const int32_t* gt_idx, int nnz, int32_t* ranks_i2t, int32_t* ranks_t2i,
                   void* workspace, cudaStream_t st);
size_t nearest_rows_workspace_bytes(int Q, int T, int D);
int nearest_rows(const float* query, const float* bank, int Q, int T, int D, int32_t* idx_out, float* cos_out,
                 void* workspace, cudaStream_t st);
int topk_fill_rows(const float* S, float* out, int nrows, int ncols, int k, float fill, cudaStream_t st);

}  // namespace vldd
