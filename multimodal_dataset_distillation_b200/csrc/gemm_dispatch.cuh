// GEMM front end of the engine: tcgen05 (3xTF32, TMA-fed) when the operands satisfy the tensor-map constraints,
// fp32 CUDA-core kernel otherwise (tiny test shapes, unaligned leading dimensions, MN-major extents not % 32).
// The product has ONE arithmetic path (no backend dispatch): 3xTF32 on the tensor cores; the CUDA-core kernel only takes
// shapes a tensor map cannot describe.  A developer build with -DVLDD_DEV_GEMM_SWITCH (VLDD_NVCC_DEFS, build.py) restores
// the A/B switches used while the kernels were brought up: VLDD_GEMM=simt (CUDA-core kernel everywhere) and
// VLDD_GEMM=tf32 (one tf32 product per fp32 product, ~6e-4 relative error per GEMM).  Neither exists in the shipped library.
#pragma once
#include <cstdlib>
#include <cstring>

#include "gemm_simt.cuh"
#include "tc_gemm_host.cuh"

// pipeline depth per GEMM class (64 KB of shared memory per stage; 0 = the kernel's default of 3).  Measured on B200 at the
// Flickr shape (ms / iteration): partial/axpy = 3/3: 2.53, 3/1: 2.54, 2/1: 2.74, 2/2: 2.81 -- depth matters for the skinny
// split-K GEMMs, and shrinking the weight-gradient CTAs to co-reside with them buys nothing.
#ifndef VLDD_STAGES_PARTIAL
#define VLDD_STAGES_PARTIAL 0
#endif
#ifndef VLDD_STAGES_AXPY
#define VLDD_STAGES_AXPY 0
#endif

namespace vldd {

#ifdef VLDD_DEV_GEMM_SWITCH
inline bool tc_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("VLDD_GEMM");
    v = (e && strcmp(e, "simt") == 0) ? 0 : 1;
  }
  return v == 1;
}
inline bool tf32_single_pass() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("VLDD_GEMM");
    v = (e && strcmp(e, "tf32") == 0) ? 1 : 0;
  }
  return v == 1;
}
#else
constexpr bool tc_enabled() { return true; }
constexpr bool tf32_single_pass() { return false; }
#endif

inline int simt_pick_splits(int M, int N, int Ktot) {
  const int tiles = ceil_div(M, GBM) * ceil_div(N, GBN);
  const int nkb = ceil_div(Ktot, GBK);
  int s = (num_sms() + tiles - 1) / tiles;
  const int max_s = nkb / 2 > 0 ? nkb / 2 : 1;
  if (s > max_s) s = max_s;
  return s < 1 ? 1 : s;
}
// upper bound of the split count either backend may choose (workspace sizing)
inline int max_splits(int M, int N, int Ktot) {
  const int a = simt_pick_splits(M, N, Ktot), b = tc::pick_splits(M, N, Ktot), c = tc::pick_splits(M, N, Ktot, 64);
  return a > b ? (a > c ? a : c) : (b > c ? b : c);
}

// Tile width of the split-K GEMMs (profiles/gemm_sweep_r02*.txt, B200, GEMM + slab consumer timed as a dependent pair).
// The main loop is bound by the shared-memory port: per k-block the A tile crosses it twice (TMA write, splitter read) and the
// B tile six times (TMA write, splitter read, B_lo write, three MMA reads) = 32 KB + 6 * BN * 128 B at 128 B/clk, i.e.
// 0.50 us at BN = 128 and 0.32 us at BN = 64, and a 64-wide tile needs half the K splits to fill one wave, so its consumer
// sums half as many slabs: p (K = 768) 10.1 -> 8.8 us, S 10.9 -> 9.8, dY (N = 768) 12.2 -> 11.0.  The wide-N products with
// K >= 2304 keep 128 columns: f / dh (K = 2304) are a tie as a pair (13.2 vs 13.1 us) but the GEMM alone is 9.1 vs 10.1 us, and the
// K = 4608 tangent products lose (18.1 vs 19.1 us): with 18-36 k-blocks per CTA the re-read of the activation tile by twice as many
// n-tiles outweighs the slabs.
// (one M tile only: with many M tiles there is no split-K to save and the wider tile does more MMA work per byte of shared memory)
inline bool narrow_tile(const GemmOperands& g) { return g.M <= tc::BM && !(g.K0 + g.K1 >= 2048 && g.N >= 1024); }
inline int tc_partial_splits(const GemmOperands& g) { return tc::pick_splits(g.M, g.N, g.K0 + g.K1, narrow_tile(g) ? 64 : 128); }

// partial slabs: part[z][M*N], returns the split count through *splits
template <bool AK, bool BKm>
inline int gemm_partial(const GemmOperands& g, float* part, int* splits, cudaStream_t st, int old_mask = 0) {
  const int Kt = g.K0 + g.K1;
  if (tc_enabled() && tc::gemm_ok<AK, BKm>(g)) {
#ifdef VLDD_DEV_GEMM_SWITCH
    if (tf32_single_pass()) {
      *splits = tc::pick_splits(g.M, g.N, Kt);
      return tc::launch<AK, BKm, 1, tc::EpiPartial>(g, *splits, tc::EpiPartial{part, (long long)g.M * g.N}, st, nullptr, nullptr, old_mask);
    }
#endif
    *splits = tc_partial_splits(g);
#ifndef VLDD_NO_TMA_STORE
    if (tc::slabs_tma_ok(part, g.M, g.N)) {          // slabs leave through TMA stores (one instruction per 32 x 32 block)
      if (narrow_tile(g))
        return tc::launch<AK, BKm, 3, tc::EpiPartialTma, VLDD_STAGES_PARTIAL, 64>(g, *splits, tc::EpiPartialTma{part, (long long)g.M * g.N}, st, nullptr, nullptr, old_mask);
      return tc::launch<AK, BKm, 3, tc::EpiPartialTma, VLDD_STAGES_PARTIAL>(g, *splits, tc::EpiPartialTma{part, (long long)g.M * g.N}, st, nullptr, nullptr, old_mask);
    }
#endif
    if (narrow_tile(g))
      return tc::launch<AK, BKm, 3, tc::EpiPartial, VLDD_STAGES_PARTIAL, 64>(g, *splits, tc::EpiPartial{part, (long long)g.M * g.N}, st, nullptr, nullptr, old_mask);
    return tc::launch<AK, BKm, 3, tc::EpiPartial, VLDD_STAGES_PARTIAL>(g, *splits, tc::EpiPartial{part, (long long)g.M * g.N}, st, nullptr, nullptr, old_mask);
  }
  *splits = simt_pick_splits(g.M, g.N, Kt);
  launch_gemm<AK, BKm>(g, *splits, part, EpiStore{}, st);
  return VLDD_OK;
}
// C = alpha * A B
template <bool AK, bool BKm>
inline int gemm_store(const GemmOperands& g, float* C, int ldc, float alpha, cudaStream_t st, int old_mask = 0) {
  if (tc_enabled() && tc::gemm_ok<AK, BKm>(g)) {
#ifdef VLDD_DEV_GEMM_SWITCH
    if (tf32_single_pass()) return tc::launch<AK, BKm, 1>(g, 1, tc::EpiScale{C, ldc, alpha}, st, nullptr, nullptr, old_mask);
#endif
    // Tile width by a two-term model of the persistent loop: rounds of tiles per SM x shared-memory traffic per k-block
    // (32 KB for the A tile + 0.75 KB per column of B: TMA write, split read, lo write, three MMA reads).  Narrow tiles when
    // even they give every CTA at most one tile (the engine's M = 100 products: 36 CTAs x 64 columns instead of 18 x 128),
    // 96 columns when that saves a round (1000 x 5000: 424 tiles = 3 rounds of 104 KB instead of 320 tiles = 3 rounds of
    // 128 KB: 45.2 vs 51.3 us, csrc/dev/gemm_sweep_test `flickr`).
    const int tiles_m = ceil_div(g.M, tc::BM), sms = num_sms();
    auto cost = [&](int bn) { return ceil_div(tiles_m * ceil_div(g.N, bn), sms) * (32.0 + 0.75 * bn); };
    const tc::EpiScale epi{C, ldc, alpha};
    if (tiles_m * ceil_div(g.N, 64) <= sms) return tc::launch<AK, BKm, 3, tc::EpiScale, 0, 64>(g, 1, epi, st, nullptr, nullptr, old_mask);
    if (cost(96) < cost(128)) return tc::launch<AK, BKm, 3, tc::EpiScale, 0, 96>(g, 1, epi, st, nullptr, nullptr, old_mask);
    return tc::launch<AK, BKm, 3>(g, 1, epi, st, nullptr, nullptr, old_mask);
  }
  launch_gemm<AK, BKm>(g, 1, nullptr, EpiStore{C, ldc, alpha}, st);
  return VLDD_OK;
}
// dst = src - (*lr) * A B      (src nullable)
template <bool AK, bool BKm>
inline int gemm_axpy(const GemmOperands& g, const float* src, float* dst, int ld, const float* lr, cudaStream_t st,
                     int old_mask = 0) {
#ifdef VLDD_DEV_GEMM_SWITCH
  if (tc_enabled() && tc::gemm_ok<AK, BKm>(g) && tf32_single_pass())
    return tc::launch<AK, BKm, 1, tc::EpiAxpyTC>(g, 1, tc::EpiAxpyTC{src, dst, ld, lr}, st, nullptr, nullptr, old_mask);
#endif
  if (tc_enabled() && tc::gemm_ok<AK, BKm>(g)) {
    // 128 x 96 tiles when they divide N: 2304 x 2304 -> 432 work items = 2.92 per SM (three even rounds) instead of
    // 324 = 2.19 (28 CTAs run a third round while 120 idle)
    // (8 epilogue warps -- two per TMEM lane quadrant -- keep twice the `src` loads in flight: 15.1 -> 13.8 us for K = 100;
    //  with two K segments the 4-stage pipeline they leave room for costs what they gain: 18.6 vs 18.8 us)
    if (g.N % 96 == 0 && ceil_div(g.M, tc::BM) * (g.N / 128) > num_sms()) {
      if (g.K1 == 0)
        return tc::launch<AK, BKm, 3, tc::EpiAxpyTC, VLDD_STAGES_AXPY, 96, 8>(g, 1, tc::EpiAxpyTC{src, dst, ld, lr}, st, nullptr, nullptr, old_mask);
      return tc::launch<AK, BKm, 3, tc::EpiAxpyTC, VLDD_STAGES_AXPY, 96>(g, 1, tc::EpiAxpyTC{src, dst, ld, lr}, st, nullptr, nullptr, old_mask);
    }
    if (g.N % 96 == 0 && g.K1 == 0)
      return tc::launch<AK, BKm, 3, tc::EpiAxpyTC, VLDD_STAGES_AXPY, 96, 8>(g, 1, tc::EpiAxpyTC{src, dst, ld, lr}, st, nullptr, nullptr, old_mask);
    return tc::launch<AK, BKm, 3, tc::EpiAxpyTC, VLDD_STAGES_AXPY>(g, 1, tc::EpiAxpyTC{src, dst, ld, lr}, st, nullptr, nullptr, old_mask);
  }
  launch_gemm<AK, BKm>(g, 1, nullptr, EpiAxpy{src, dst, ld, lr}, st);
  return VLDD_OK;
}

}  // namespace vldd
