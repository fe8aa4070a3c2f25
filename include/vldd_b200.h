/* vldd_b200 -- C ABI of the B200-native hot path of vision-language trajectory-matching distillation.
 *
 * The reference (kushal-bhargav/multimodal_dataset_distillation) has no FFI / operator interface: its hot
 * path is inline Python.  Each entry point below therefore cites the reference
 * LINES it replaces; the Python names that sit on top of it (ReparamModule, itm_eval, epoch_test,
 * evaluate_synset, the distill.py CLI) are mirrored in multimodal_dataset_distillation_b200/*.py and bind
 * these symbols through ctypes (INTEGRATION.md shows the stub a maintainer adds to the reference).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to contiguous fp32 / int32 / int64 data unless the name ends in
 *     `_host`; scalars that the reference keeps as tensors (syn_lr, logit scale) are device scalars so no
 *     call synchronises the host;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); calls only enqueue work,
 *     except the `*_host` entry points, which copy in, run, copy out and synchronise `stream`;
 *   - return value: 0 on success, negative on error (VLDD_ERR_*); vldd_last_error() gives the message of
 *     the calling thread's last failure.  There is no CPU fallback: without a CUDA device calls fail.
 *   - the library owns no memory across calls except a grow-only device scratch used by `*_host` calls;
 *     workspaces are caller-provided (sizes from the matching *_workspace_bytes function).
 */
#ifndef VLDD_B200_H_
#define VLDD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VLDD_ERR_ARG (-1)
#define VLDD_ERR_CUDA (-2)
#define VLDD_ERR_WORKSPACE (-3)

int vldd_version(void);
const char* vldd_last_error(void);

/* ---- flat-parameter streaming ops --------------------------------------------------------------- */

/* out = theta - (*lr) * grad.   distill.py:582-583 (`student_params[-1] - syn_lr * grad`).  12 B/param. */
int vldd_flat_sgd_step(const float* theta, const float* grad, const float* lr, float* out, int64_t n, void* stream);

/* out3 = {sum (theta_K-theta_tgt)^2, sum (theta_0-theta_tgt)^2, their ratio}.   distill.py:588-598
 * (4x mse_loss(reduction="sum") + division).  `scratch`: vldd_match_loss_scratch_bytes() bytes whose first 16 bytes
 * are zero before the first use.  12 B/param, deterministic. */
size_t vldd_match_loss_scratch_bytes(void);
int vldd_match_loss_fwd(const float* theta_K, const float* theta_tgt, const float* theta_0, int64_t n, float* out3,
                        void* scratch, void* stream);
/* a = g * 2 (theta_K - theta_tgt) / den, g = gout ? *gout : 1; num_den = out3 of the forward.   distill.py:606. */
int vldd_match_loss_bwd(const float* theta_K, const float* theta_tgt, const float* num_den, const float* gout, float* a,
                        int64_t n, void* stream);

/* buf = first ? g : momentum*buf + g;  p -= lr*buf   (in place).   torch.optim.SGD(momentum=0.5) built at
 * distill.py:233-241 and stepped at distill.py:611-613. */
int vldd_momentum_sgd(float* p, const float* g, float* buf, float lr, float momentum, int first, int64_t n, void* stream);

/* ---- retrieval scoring --------------------------------------------------------------------------- */

/* Ranks of the ground truth for both directions given the two score matrices.   epoch.py:219-244 /
 * epoch_original.py:115-161 (`itm_eval`), rank = #{s_j > s_gt} + #{j < gt : s_j == s_gt}, i2t = min over the
 * image's captions.  img2txt is CSR (ptr[I+1], idx[]); txt2img[T].  Either direction may be skipped by passing
 * NULL scores.  Integer results: bit-exact. */
int vldd_ranks_from_scores(const float* scores_i2t, const float* scores_t2i, int n_img, int n_txt,
                           const int32_t* txt2img, const int32_t* img2txt_ptr, const int32_t* img2txt_idx,
                           int32_t* ranks_i2t, int32_t* ranks_t2i, void* stream);
/* Text->image ranks taken column-wise from the [n_img, n_txt] image->text matrix (no transpose needed): rank of row
 * txt2img[t] inside column t.   epoch.py:231-235 applied to sims_matrix.t() (epoch_original.py:101). */
int vldd_ranks_cols(const float* scores_i2t, int n_img, int n_txt, const int32_t* txt2img, int32_t* ranks_t2i, void* stream);
/* Caption-sharded ranking (multi-GPU; `scores` holds columns [col_offset, col_offset+cols) of the full matrix, ground
 * truth indices are GLOBAL).  best_gt: per row the best local ground-truth (score, global index; -inf / -1 if none).
 * count: per row #{local j : s_j > thr_score or (s_j == thr_score and global j < thr_idx)}.  Summed over shards this is
 * the rank of epoch.py:219-244 with the tie rule above. */
int vldd_rank_best_gt(const float* scores, int rows, int cols, int col_offset, const int32_t* gt_ptr,
                      const int32_t* gt_idx, float* best_score, int32_t* best_idx, void* stream);
int vldd_rank_count(const float* scores, int rows, int cols, int col_offset, const float* thr_score,
                    const int32_t* thr_idx, int32_t* counts, void* stream);
/* counts3 = {#ranks<1, #ranks<5, #ranks<10}.   epoch.py:227-229,236-238. */
int vldd_recall_counts(const int32_t* ranks, int n, int32_t* counts3, void* stream);
/* S_i2t[I,T] = scale * img @ txt^T and/or S_t2i[T,I] (either may be NULL).   epoch_original.py:94,101. */
int vldd_sim_scores(const float* img, const float* txt, int n_img, int n_txt, int dim, float scale, float* scores_i2t,
                    float* scores_t2i, void* stream);
/* out = S with every row's k largest kept and the rest := fill.   epoch_original.py:95-99, 102-105 (k=128, -100). */
int vldd_topk_fill(const float* scores, float* out, int rows, int cols, int k, float fill, void* stream);
/* Embeddings -> ranks of both directions (similarity + ranking without returning the score matrices).
 * epoch_original.py:94 + itm_eval.  workspace: vldd_sim_rank_workspace_bytes(). */
size_t vldd_sim_rank_workspace_bytes(int n_img, int n_txt, int dim);
int vldd_sim_rank(const float* img, const float* txt, int n_img, int n_txt, int dim, float scale,
                  const int32_t* txt2img, const int32_t* img2txt_ptr, const int32_t* img2txt_idx, int32_t* ranks_i2t,
                  int32_t* ranks_t2i, void* workspace, size_t workspace_bytes, void* stream);

/* Same result as vldd_sim_rank without ever writing the score matrix: passes of the tcgen05 GEMM whose epilogues
 * (1) extract the ground-truth scores from the tiles that contain them (3xTF32) and (2) count, per image row and per
 * caption column, the entries ranked ahead of the ground truth.  From ~4 M pairs on (and dim % 8 == 0) pass 2 is a
 * SCREEN: a bf16x3 product (half the tensor time) counts every pair whose distance to its threshold exceeds the proven
 * error band alpha * eps(dim) * |img_m| * |txt_n|; the pairs inside the band (ties included) are listed, gathered and
 * decided by the 3xTF32 kernel itself, so the ranks are bit-identical to the materialised path; if the list overflows
 * (capacity ~ pairs / 2048) the exact count pass runs over everything.  nnz = img2txt_ptr[n_img] (number of CSR
 * entries).  Requires 16-byte aligned embeddings and dim % 4 == 0 (tensor-map constraints).  Workspace:
 * O(n_img + n_txt + tiles), plus -- for the screen -- the bf16 hi / lo copies of both embedding sets and the gathered
 * rows of the listed pairs (<= 2 GB). */
size_t vldd_sim_rank_fused_workspace_bytes(int n_img, int n_txt, int dim, int nnz);
int vldd_sim_rank_fused(const float* img, const float* txt, int n_img, int n_txt, int dim, float scale,
                        const int32_t* txt2img, const int32_t* img2txt_ptr, const int32_t* img2txt_idx, int nnz,
                        int32_t* ranks_i2t, int32_t* ranks_t2i, void* workspace, size_t workspace_bytes, void* stream);

/* The same fused ranking for ONE CAPTION SHARD (multi-GPU: images replicated, captions [col_offset, col_offset + n_txt) of
 * the full set on this rank; epoch_original.py:94-105 + epoch.py:219-244 over the whole set).  img2txt_idx holds GLOBAL
 * caption ids.  Phase A: per image the best ground-truth candidate among the local captions, cand_score[n_img] (+inf: none
 * here) and cand_idx[n_img] (global caption id, -1: none).  The caller merges the candidates of all shards (higher score,
 * then lower index) and calls phase B with thr_score (+inf: the image has no ground truth anywhere) and
 * thr_idx_local = global id - col_offset (any integer): row_counts[n_img] = local captions ranked ahead of the threshold
 * (to be summed over the shards; `invalid_row_rank` for +inf rows), ranks_t2i[n_txt] = final text -> image ranks of the
 * local captions.  Both phases take the SAME workspace (vldd_sim_rank_fused_workspace_bytes), phase B reads what phase A
 * left in it. */
int vldd_sim_rank_fused_candidates(const float* img, const float* txt, int n_img, int n_txt, int dim, float scale,
                                   const int32_t* txt2img, const int32_t* img2txt_ptr, const int32_t* img2txt_idx, int nnz,
                                   int col_offset, float* cand_score, int32_t* cand_idx, void* workspace, size_t workspace_bytes,
                                   void* stream);
int vldd_sim_rank_fused_count(const float* img, const float* txt, int n_img, int n_txt, int dim, float scale, const float* thr_score,
                              const int32_t* thr_idx_local, int nnz, int invalid_row_rank, int32_t* row_counts, int32_t* ranks_t2i,
                              void* workspace, size_t workspace_bytes, void* stream);

/* Host-buffer drop-in for `itm_eval(scores_i2t, scores_t2i, txt2img, img2txt)` (numpy arrays in the reference):
 * copies the matrices to the device, ranks, copies ranks back and fills result9 in the reference's key order
 * {txt_r1, txt_r5, txt_r10, txt_r_mean, img_r1, img_r5, img_r10, img_r_mean, r_mean}.  ranks_*_host may be NULL. */
int vldd_itm_eval_host(const float* scores_i2t_host, const float* scores_t2i_host, int n_img, int n_txt,
                       const int32_t* txt2img_host, const int32_t* img2txt_ptr_host, const int32_t* img2txt_idx_host,
                       int32_t* ranks_i2t_host, int32_t* ranks_t2i_host, double* result9, void* stream);

/* ---- text_projection head + InfoNCE + unroll ------------------------------------------------------ */
/* Flat layout of theta (reparam_module.py:28-51 applied to networks.py:625-646):
 *   [projection.weight d x dt | projection.bias d | fc.weight d x d | fc.bias d | layer_norm.weight d | layer_norm.bias d] */

/* z = ProjectionHead(Y) (networks.py:639-646; mask = pre-scaled dropout mask or NULL for eval), zn = z/|z|
 * (epoch_original.py:78).  z or zn may be NULL. */
size_t vldd_proj_head_workspace_bytes(int rows, int dt, int d);
int vldd_proj_head_forward(const float* theta, const float* Y, const float* mask, int rows, int dt, int d, float* z,
                           float* zn, void* workspace, size_t workspace_bytes, void* stream);

/* One contrastive step on a whole batch: loss and first-order gradients.   distill.py:524-551 + 562-567
 * (forward, normalise, logits = scale * Xn Yn^T, (CE + CE^T)/2, grad wrt the flat text parameters).
 * U = image-encoder outputs [B,d].  Outputs (any of dY,dU,dscale may be NULL): loss[1], g_theta[P], dY[B,dt],
 * dU[B,d], dscale[1]. */
size_t vldd_contrastive_step_workspace_bytes(int B, int dt, int d);
int vldd_contrastive_step(const float* theta, const float* Y, const float* U, const float* scale, const float* mask,
                          int B, int dt, int d, float* loss, float* g_theta, float* dY, float* dU, float* dscale,
                          void* workspace, size_t workspace_bytes, void* stream);

/* CLIPModel_full.forward from the encoder outputs on, with the gradients loss.backward() produces.   networks.py:866-889:
 *   txt = text_projection(Y; theta) (868-870), row-normalise both (873-874), logits = scale * Xn Yn^T (877-878; the
 *   reference passes scale = exp(log(1/0.07))), loss = (CE(logits) + CE(logits^T)) / 2 (881-882),
 *   top1 = {#rows whose argmax is the diagonal, #columns whose argmax is the diagonal} (884-885; acc = (top1[0]+top1[1])/2).
 * Same workspace and gradient outputs as vldd_contrastive_step; top1 may be NULL.   Used by epoch.py:59-98 (`epoch`). */
int vldd_clip_loss(const float* theta, const float* Y, const float* U, const float* scale, const float* mask, int B, int dt,
                   int d, float* loss, int32_t* top1, float* g_theta, float* dY, float* dU, float* dscale, void* workspace,
                   size_t workspace_bytes, void* stream);

/* Bidirectional InfoNCE on already row-normalised features as a twice-differentiable node ("Mode B": any image tower
 * under PyTorch autograd, e.g. pixels -> NFNet through ReparamModule as in distill.py:524-567, with the loss of
 * distill.py:548-551 and autograd.grad(create_graph=True) on top of it).
 *   vldd_infonce_grad: loss[1] = (CE(S) + CE(S^T)) / 2 with S = scale * xn yn^T; dxn[B,d], dyn[B,d], dscale[1] (nullable)
 *                      = its gradient.
 *   vldd_infonce_hvp:  for a direction (cx[B,d], cy[B,d], cs[1]):  Ldot[1] = <grad L, direction> and
 *                      hx[B,d], hy[B,d], hs[1] = the gradient of Ldot w.r.t. (xn, yn, scale), i.e. the Hessian of the loss
 *                      applied to the direction -- what the backward of the first-order gradients needs.
 * scale, cs: device scalars.  Stateless: both take the same workspace (vldd_infonce_workspace_bytes). */
size_t vldd_infonce_workspace_bytes(int B, int d);
int vldd_infonce_grad(const float* xn, const float* yn, const float* scale, int B, int d, float* loss, float* dxn, float* dyn,
                      float* dscale, void* workspace, size_t workspace_bytes, void* stream);
int vldd_infonce_hvp(const float* xn, const float* yn, const float* scale, const float* cx, const float* cy, const float* cs,
                     int B, int d, float* Ldot, float* hx, float* hy, float* hs, void* workspace, size_t workspace_bytes,
                     void* stream);

/* Nearest bank row per query by cosine similarity, first index on ties.   distill.py:89-95 (`nearest_neighbor`:
 * sklearn cosine_similarity(query, database) + np.argmax per query; rows are L2-normalised, all-zero rows left as is).
 * idx_out[n_query] int32; cos_out[n_query] (nullable) = the winning cosine.  Workspace holds the normalised copies and
 * the [n_query, n_bank] cosine matrix. */
size_t vldd_nearest_rows_workspace_bytes(int n_query, int n_bank, int dim);
int vldd_nearest_rows(const float* query, const float* bank, int n_query, int n_bank, int dim, int32_t* idx_out,
                      float* cos_out, void* workspace, size_t workspace_bytes, void* stream);

/* The whole inner loop of one expert segment and its backward.   distill.py:509-606:
 *   for k < K: idx = perms[k] (510-511); g = grad(InfoNCE(head(Y[idx]; theta_k), U[idx]), theta_k) (524-567);
 *              theta_{k+1} = theta_k - lr g (583);
 *   loss = |theta_K - theta_tgt|^2 / |theta_0 - theta_tgt|^2 (588-598);  backward to Y, U, lr, scale (606).
 * perms: int64 [K,B] (unique indices per row);  masks: [K,B,d] pre-scaled dropout masks or NULL.
 * dropout_p > 0 (needs masks != NULL and rng_state): the engine DRAWS the K masks itself into `masks` at the head of its
 * launch graph (networks.py:636,643 nn.Dropout(p), students in train mode distill.py:446-447) -- Philox4x32-10 keyed by
 * rng_state[0] (seed) with rng_state[1] (draws so far, advanced by one per call) in the counter -- and the reverse sweep
 * reads the same buffer, i.e. replays the same masks.  dropout_p == 0: `masks` is used as given (parity mode).
 * Outputs: out5 = {num, den, loss, dloss/dlr, dloss/dscale}; ce[K] per-step contrastive losses (nullable);
 * dY[N,dt]; dU[N,d]; theta_K[P] (nullable). */
size_t vldd_unrolled_match_workspace_bytes(int N, int B, int K, int dt, int d);
int vldd_unrolled_match(const float* theta0, const float* theta_tgt, const float* Y, const float* U, const float* lr,
                        const float* scale, const int64_t* perms, float* masks, float dropout_p,
                        unsigned long long* rng_state, int N, int B, int K, int dt, int d, float* out5, float* ce, float* dY,
                        float* dU, float* theta_K, void* workspace, size_t workspace_bytes, void* stream);

/* Numerator and adjoint of the matching loss in one pass (distill.py:588-598 + the first step of 606):
 *   out3 = {num = |theta_K - theta_tgt|^2, *den, num / *den};  adjoint = 2 (theta_K - theta_tgt) / *den.
 * `den` = |theta_0 - theta_tgt|^2 is a device scalar computed beforehand (the engine accumulates it while it stages the
 * segment).  12 B / parameter; scratch: vldd_match_loss_scratch_bytes().  Two launches: the streaming pass (adjoint + fp64
 * block partials of the numerator) and a one-block finish that adds the partials in index order; out3 == NULL skips the
 * finish (the unroll engine folds it into the last kernel of its launch graph instead). */
int vldd_match_final(const float* theta_K, const float* theta_tgt, const float* den, int64_t n, float* out3, float* adjoint,
                     void* scratch, void* stream);

/* The three torch.optim.SGD(momentum) steps of one outer iteration in ONE launch (distill.py:233-241, 603-613):
 *   buf = first ? g : momentum * buf + g;  p -= lr * buf   for U ("image_syn", Mode A), Y (text_syn) and the two
 * learnable student learning rates *syn_lr_img, *syn_lr_txt (device scalars, each nullable; momentum in buf_lr[0], buf_lr[1])
 * with gradients *g_lr_img (nullable = 0; the fork's logit-scale path, distill.py:548) and *g_lr_txt.  Gradients are multiplied by grad_scale first (1 = sum over
 * segments / ranks, 1/segments = mean).  If `loss` is given and *loss is not finite nothing is updated and *skipped = 1
 * (the reference breaks before stepping, distill.py:599-600). */
int vldd_outer_update(float* U, const float* gU, float* bufU, int64_t nU, float lr_img, float* Y, const float* gY, float* bufY,
                      int64_t nY, float lr_txt, float* syn_lr_img, float* syn_lr_txt, const float* g_lr_img, const float* g_lr_txt,
                      float* buf_lr, float lr_lr, float momentum, int first, float grad_scale, const float* loss, int* skipped,
                      void* stream);

/* Pre-scaled dropout masks (0 or 1/(1-p)) from the engine's generator: Philox4x32-10, key = rng_state[0], counter =
 * (element index / 4, rng_state[1]); advance != 0 bumps rng_state[1] afterwards.  networks.py:629,636. */
int vldd_dropout_masks(float* masks, int64_t n, float p, unsigned long long* rng_state, int advance, void* stream);

/* ---- measurement hook ------------------------------------------------------------------------------ */
/* One launch of the weight-streaming GEMM the engine issues for networks.py:642 (`fc`: f = h W2^T) inside the unroll:
 * partial[z][M*N] (z < *splits_out) are split-K slabs of A[M,K] @ W[N,K]^T computed by the tcgen05 3xTF32 kernel.
 * Used by bench.py to time the dominant kernel in isolation for the roofline figure. */
size_t vldd_bench_skinny_gemm_workspace_bytes(int M, int N, int K);
int vldd_bench_skinny_gemm(const float* A, const float* W, int M, int N, int K, float* partial, size_t partial_bytes,
                           int* splits_out, void* stream);

/* Kernels of this library launched by the process so far (a replayed CUDA graph counts its kernel nodes on every
 * replay).  Measurement aid: bench.py reports the difference across its timed region as `gpu_launches`. */
unsigned long long vldd_kernel_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* VLDD_B200_H_ */
