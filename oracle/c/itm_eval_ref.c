/* Oracle (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py): plain-C restatement of the integer part of the
 * retrieval path, independent of the numpy restatement in oracle/retrieval_ref.py.
 *
 *   reference: epoch.py:219-244 / epoch_original.py:115-161 (itm_eval)
 *     image->text: rank_i = min over the image's ground-truth captions c of the position of c in the row sorted by
 *                  descending score (epoch.py:222-226); text->image: position of txt2img[t] in row t (231-235);
 *                  recall@k = 100 * #(rank < k) / n for k in {1, 5, 10}; the three means (239-243).
 *   position in a descending sort, ties broken by index (the north star's rule; numpy's argsort order among equal
 *   scores is unspecified):  pos(c) = #{j : s_j > s_c} + #{j < c : s_j == s_c}.
 *
 * Built by oracle/c/Makefile into oracle/_ref/libitm_ref.so and loaded by oracle/retrieval_ref.py::c_oracle().
 */
#include <stdint.h>

static int32_t position(const float* row, int n, int c) {
  const float sc = row[c];
  int32_t ahead = 0;
  for (int j = 0; j < n; ++j) ahead += (row[j] > sc) || (row[j] == sc && j < c);
  return ahead;
}

/* scores [rows, cols] row-major; ground truth of row r = gt_idx[gt_ptr[r] .. gt_ptr[r+1]) ; ranks[rows] */
void itm_ranks_rows(const float* scores, int rows, int cols, const int32_t* gt_ptr, const int32_t* gt_idx,
                    int32_t* ranks) {
  for (int r = 0; r < rows; ++r) {
    int32_t best = 1000000000;
    for (int e = gt_ptr[r]; e < gt_ptr[r + 1]; ++e) {
      const int32_t p = position(scores + (int64_t)r * cols, cols, gt_idx[e]);
      if (p < best) best = p;
    }
    ranks[r] = best;
  }
}

/* out9 = txt_r1, txt_r5, txt_r10, txt_r_mean, img_r1, img_r5, img_r10, img_r_mean, r_mean   (epoch.py:227-243) */
void itm_result(const int32_t* ranks_i2t, int n_img, const int32_t* ranks_t2i, int n_txt, double* out9) {
  const int ks[3] = {1, 5, 10};
  for (int d = 0; d < 2; ++d) {
    const int32_t* r = d == 0 ? ranks_i2t : ranks_t2i;
    const int n = d == 0 ? n_img : n_txt;
    double mean = 0.0;
    for (int k = 0; k < 3; ++k) {
      int64_t hits = 0;
      for (int i = 0; i < n; ++i) hits += r[i] < ks[k];
      out9[4 * d + k] = n > 0 ? 100.0 * (double)hits / (double)n : 0.0;
      mean += out9[4 * d + k];
    }
    out9[4 * d + 3] = mean / 3.0;
  }
  out9[8] = (out9[3] + out9[7]) / 2.0;
}
