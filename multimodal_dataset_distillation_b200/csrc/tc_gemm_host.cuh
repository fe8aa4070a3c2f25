// Host side of the tcgen05 GEMM: tensor-map construction (cached: operands live at fixed workspace addresses, so
// each map is encoded once), eligibility checks and the launcher.
#pragma once
#include <mutex>
#include <unordered_map>

#include "gemm_simt.cuh"
#include "tc_gemm.cuh"

namespace vldd {
namespace tc {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct MapKey {
  const void* ptr; int rows, K, ld, kmajor, box_rows;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && rows == o.rows && K == o.K && ld == o.ld && kmajor == o.kmajor && box_rows == o.box_rows;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    h ^= (size_t)k.rows * 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h ^= (size_t)k.K * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2);
    h ^= (size_t)(k.ld * 2 + k.kmajor + 1024 * k.box_rows) * 0x165667B19E3779F9ull + (h << 6) + (h >> 2);
    return h;
  }
};

// operand with `rows` = its M or N extent.  kmajor: element (row,k) at ptr[row*ld + k]; else at ptr[k*ld + row].
inline bool operand_ok(const float* ptr, int rows, int K, int ld, bool kmajor) {
  if (ptr == nullptr || !aligned16(ptr) || ld % 4 != 0 || K <= 0 || rows <= 0) return false;
  if (!kmajor && rows % 32 != 0) return false;
  return true;
}

inline int get_map(const float* ptr, int rows, int K, int ld, bool kmajor, CUtensorMap* out, int box_rows = BM) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  const MapKey key{ptr, rows, K, ld, kmajor ? 1 : 0, box_rows};
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return VLDD_OK; }
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return VLDD_ERR_CUDA; }
  CUtensorMap m;
  CUresult r;
  if (kmajor) {
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    const cuuint32_t es[2] = {1, 1};
    r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, es,
           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    const cuuint64_t dims[3] = {32, (cuuint64_t)K, (cuuint64_t)(rows / 32)};
    const cuuint64_t strides[2] = {(cuuint64_t)ld * 4, 128};
    const cuuint32_t box[3] = {32, (cuuint32_t)BK, (cuuint32_t)(box_rows / 32)};
    const cuuint32_t es[3] = {1, 1, 1};
    r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), dims, strides, box, es,
           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) for rows=%d K=%d ld=%d kmajor=%d", (int)r, rows, K, ld, (int)kmajor);
    return VLDD_ERR_CUDA;
  }
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, m);
  *out = m;
  return VLDD_OK;
}

// K-major bf16 operand [rows, K] (row stride ld elements): 2-D map, box 64 x box_rows, SWIZZLE_128B
inline int get_map_bf16(const void* ptr, int rows, int K, int ld, CUtensorMap* out, int box_rows) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  const MapKey key{ptr, rows, K, ld, 2, box_rows};
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return VLDD_OK; }
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return VLDD_ERR_CUDA; }
  CUtensorMap m;
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  const cuuint32_t es[2] = {1, 1};
  const CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(bf16) failed (%d) for rows=%d K=%d ld=%d", (int)r, rows, K, ld);
    return VLDD_ERR_CUDA;
  }
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, m);
  *out = m;
  return VLDD_OK;
}

// output map of the TMA-store epilogue: slabs part[z][M][N] fp32 as a 3-D tensor {N, M, splits}, box 32 x 32 x 1, SWIZZLE_128B
inline int get_map_slabs(const float* part, int M, int N, int splits, CUtensorMap* out) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  const MapKey key{part, M, N, splits, 3, 32};
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return VLDD_OK; }
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return VLDD_ERR_CUDA; }
  CUtensorMap m;
  const cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)M, (cuuint64_t)splits};
  const cuuint64_t strides[2] = {(cuuint64_t)N * 4, (cuuint64_t)M * N * 4};
  const cuuint32_t box[3] = {32, 32, 1};
  const cuuint32_t es[3] = {1, 1, 1};
  const CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(part), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(slabs) failed (%d) for M=%d N=%d splits=%d", (int)r, M, N, splits);
    return VLDD_ERR_CUDA;
  }
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, m);
  *out = m;
  return VLDD_OK;
}
inline bool slabs_tma_ok(const float* part, int M, int N) { return aligned16(part) && N % 4 == 0 && ((long long)M * N) % 4 == 0; }

// C = A B^T from pre-split bf16 operands (A = A_hi + A_lo [M, K], B = B_hi + B_lo [N, K], K-major, K % 8 == 0, 16-byte
// aligned): three bf16 tensor-core products per k-step (tc_gemm.cuh, kSplit == 6), 128 x BN tiles.
template <class Epi, int BN = 256>
inline int launch_bf16x3(const void* A_hi, const void* A_lo, const void* B_hi, const void* B_lo, int M, int N, int K, Epi epi,
                         cudaStream_t st, const int* work_list = nullptr, const int* work_count = nullptr) {
  using C = Cfg<6, false, 0, BN, 4>;
  auto kern = tc_gemm_kernel<true, true, 6, Epi, 0, BN, 4>;
  static bool configured[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { set_error("tc::launch_bf16x3: bad device ordinal"); return VLDD_ERR_CUDA; }
  if (!configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(smem=%d) failed: %s", C::kSmemBytes, cudaGetErrorString(e)); return VLDD_ERR_CUDA; }
    configured[dev] = true;
  }
  if (K % 8 != 0 || !aligned16(A_hi) || !aligned16(A_lo) || !aligned16(B_hi) || !aligned16(B_lo)) {
    set_error("launch_bf16x3: K %% 8 != 0 or unaligned operand");
    return VLDD_ERR_ARG;
  }
  Maps maps;
  int rc = get_map_bf16(A_hi, M, K, K, &maps.a0, BM);
  if (!rc) rc = get_map_bf16(A_lo, M, K, K, &maps.a1, BM);
  if (!rc) rc = get_map_bf16(B_hi, N, K, K, &maps.b0, BN);
  if (!rc) rc = get_map_bf16(B_lo, N, K, K, &maps.b1, BN);
  if (rc) return rc;
  const int work = ceil_div(N, BN) * ceil_div(M, BM);
  dim3 grid(work < num_sms() ? work : num_sms());
  launch_k(kern, grid, C::kThreads, C::kSmemBytes, st, maps, M, N, K, 0, 1, epi, work_list, work_count, 0);
  return VLDD_OK;
}

template <bool A_KMAJOR, bool B_KMAJOR>
inline bool gemm_ok(const GemmOperands& g) {
  if (!operand_ok(g.A0, g.M, g.K0, g.lda0, A_KMAJOR) || !operand_ok(g.B0, g.N, g.K0, g.ldb0, B_KMAJOR)) return false;
  if (g.K1 > 0 && (!operand_ok(g.A1, g.M, g.K1, g.lda1, A_KMAJOR) || !operand_ok(g.B1, g.N, g.K1, g.ldb1, B_KMAJOR)))
    return false;
  return true;
}

// Background GEMMs (launched on a side lane next to critical-path kernels) leave some SMs to the main lane: the caller caps
// the persistent grid for the duration of a scope (n > 0: at most n CTAs; -f: about 1/f of the SMs, rounded so that every CTA
// runs the same number of tiles; 0: no cap).
inline int& grid_cap() { static thread_local int cap = 0; return cap; }
struct GridCapScope {
  int saved;
  explicit GridCapScope(int cap) : saved(grid_cap()) { grid_cap() = cap; }
  ~GridCapScope() { grid_cap() = saved; }
};

template <bool A_KMAJOR, bool B_KMAJOR, int kSplit, class Epi, int kStagesT = 0, int BN = 128, int kEpiWarps = 4>
inline int launch(const GemmOperands& g, int splits, Epi epi, cudaStream_t st, const int* work_list = nullptr,
                  const int* work_count = nullptr, int old_mask = 0) {
  using C = Cfg<kSplit, kSplit == 3, kStagesT, BN, kEpiWarps, epi_uses_tma<Epi>::value>;
  auto kern = tc_gemm_kernel<A_KMAJOR, B_KMAJOR, kSplit, Epi, kStagesT, BN, kEpiWarps>;
  // the attribute is per device (a process may drive several GPUs): one flag per device ordinal and template instance
  static bool configured[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { set_error("tc::launch: bad device ordinal"); return VLDD_ERR_CUDA; }
  if (!configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(smem=%d) failed: %s", C::kSmemBytes, cudaGetErrorString(e)); return VLDD_ERR_CUDA; }
    configured[dev] = true;
  }
  Maps maps;
  int rc = get_map(g.A0, g.M, g.K0, g.lda0, A_KMAJOR, &maps.a0);
  if (rc) return rc;
  rc = get_map(g.B0, g.N, g.K0, g.ldb0, B_KMAJOR, &maps.b0, BN);
  if (rc) return rc;
  if (g.K1 > 0) {
    rc = get_map(g.A1, g.M, g.K1, g.lda1, A_KMAJOR, &maps.a1);
    if (rc) return rc;
    rc = get_map(g.B1, g.N, g.K1, g.ldb1, B_KMAJOR, &maps.b1, BN);
    if (rc) return rc;
  } else {
    maps.a1 = maps.a0;
    maps.b1 = maps.b0;
  }
  maps.c = maps.a0;
  if constexpr (epi_uses_tma<Epi>::value) {
    rc = get_map_slabs(epi.part, g.M, g.N, splits, &maps.c);
    if (rc) return rc;
  }
  const int work = ceil_div(g.N, BN) * ceil_div(g.M, BM) * splits;
  int ctas = work < num_sms() ? work : num_sms();                         // persistent: one CTA per SM at most
  if (grid_cap() > 0 && ctas > grid_cap()) ctas = grid_cap();
  if (grid_cap() < 0) {                                                   // -f: about 1/f of the SMs, in whole rounds of tiles
    const int target = num_sms() / (-grid_cap());
    if (work > target) ctas = ceil_div(work, ceil_div(work, target));
  }
  dim3 grid(ctas);
  launch_k(kern, grid, C::kThreads, C::kSmemBytes, st, maps, g.M, g.N, g.K0, g.K1, splits, epi, work_list, work_count, old_mask);
  return VLDD_OK;
}

// number of K splits so that tiles x splits fills ONE wave of the 148 SMs without spilling into a second one (the
// kernel runs one CTA per SM: 18 tiles x 9 splits = 162 CTAs would cost two waves), each split keeping >= 2 k-blocks
inline int pick_splits(int M, int N, int Ktot, int bn = 128) {
  const int tiles = ceil_div(M, BM) * ceil_div(N, bn);
  const int nkb = ceil_div(Ktot, BK);
  int s = num_sms() / tiles;
  const int max_s = nkb / 2 > 0 ? nkb / 2 : 1;
  if (s > max_s) s = max_s;
  return s < 1 ? 1 : s;
}

}  // namespace tc
}  // namespace vldd
