// tcgen05 / TMEM / TMA GEMM for sm_100a with fp32-grade accuracy (3xTF32 operand split).
//
//   C[m,n] = sum_k opA(A)[m,k] * opB(B)[k,n]      (+ optional second K segment A1/B1, same shapes)
//
// Persistent: CTA b handles work items b, b + grid, ... (one item = one 128x128 output tile x one K split); the
// accumulator is double-buffered in TMEM so the epilogue of item i overlaps the main loop of item i+1.
// Warp roles (320 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor loads of the fp32 A/B k-blocks (128 x 32 floats each, SWIZZLE_128B)
//   warp 1      TMEM allocator + MMA issuer: tcgen05.mma.cta_group::1.kind::tf32, M=128 N=128 K=8, accumulator in TMEM
//   warps 6..9  epilogue: tcgen05.ld (32 lanes x 32 columns per instruction), 32x32 transpose through shared memory,
//               fused epilogue (partial-slab store / alpha store / theta - lr*acc) with coalesced 128-byte rows
//   warps 2..5  operand splitter.  tf32 keeps 10 mantissa bits and the tensor core TRUNCATES fp32 operands,
//               so the landed fp32 tile is the hi operand as is and only lo = x - trunc_tf32(x) is produced (second
//               buffer at the same swizzled offsets); the MMA warp issues lo*hi + hi*lo + hi*hi -> error ~2^-21 per
//               product instead of 2^-11.  For K-major A the hi / lo rows go straight into TENSOR MEMORY
//               (tcgen05.st, one TMEM lane per row) and the MMAs take A from TMEM: no A traffic on the shared-memory port.
// Operand layouts: K-major  = row-major [rows, K]  (2-D tensor map, box 32 x 128);
//                  MN-major = row-major [K, rows]  (3-D tensor map {32, K, rows/32}, box 32 x 32 x 4), used for
//                  dH = dF W2 and the weight-gradient GEMMs dW = dF^T H, whose operands are contiguous along M/N.
#pragma once
#include <cuda.h>

#include <type_traits>

#include "common.cuh"

namespace vldd {
namespace tc {

constexpr int BM = 128, BK = 32;                    // BK floats = 128 bytes = one swizzle row; BN is a template parameter
constexpr int UMMA_K = 8;                           // tf32
constexpr int TILE_BYTES = BM * BK * 4;             // 16 KB per operand per stage
constexpr int kBaseThreads = 192;                   // producer, MMA, 4 splitter warps; + 32 per epilogue warp (4 or 8 of them)

// kSplit == 6 ("bf16x3"): operands arrive PRE-SPLIT as bf16 hi / lo pairs (x = hi + lo + O(2^-18 |x|)), K-major only, four
//   tensor maps (a0 = A_hi, a1 = A_lo, b0 = B_hi, b1 = B_lo; no second K segment); per 16-element k-step the MMA warp issues
//   lo*hi + hi*lo + hi*hi as kind::f16 -- half the tensor time of 3xTF32 and no splitter work.  Error ~3 * 2^-18 per product:
//   used as the SCREEN of the fused retrieval ranking, whose borderline pairs are re-decided in 3xTF32.
// Stage layout (kSplit == 3):  K-major A  -> [A | B | B_lo]            48 KB x 4 stages; A_hi / A_lo live in TMEM
//                              MN-major A -> [A | B | A_lo | B_lo]     64 KB x 3 stages
//               (kSplit == 1):              [A | B]                    32 KB x 6 stages
template <int kSplit, bool A_TMEM, int kStagesT = 0, int BN = 128, int kEpiWarps = 4, bool kTmaEpi = false>
struct Cfg {
  static_assert(kEpiWarps == 4 || kEpiWarps == 8, "one or two epilogue warps per TMEM lane quadrant");
  static constexpr int kThreads = kBaseThreads + 32 * kEpiWarps;
  static constexpr int kTileB = BN * BK * 4;                              // B tile bytes (12 KB at BN = 96)
  static constexpr int kStageBytes = kSplit == 6 ? 2 * TILE_BYTES + 2 * kTileB
                                     : kSplit == 3 ? (A_TMEM ? TILE_BYTES + 2 * kTileB : 2 * TILE_BYTES + 2 * kTileB)
                                                   : TILE_BYTES + kTileB;
  // Pipeline depth.  One stage's round trip (TMA issue -> data landed ~0.8-1.0 us, operand split ~0.25 us, its 12 MMAs
  // ~0.4 us, commit -> producer) is ~1.8 us, so with d stages a k-block costs max(MMA time, 1.8 us / d): narrower N tiles
  // buy depth (shared memory: 48 / 40 / 32 KB per stage at BN = 128 / 96 / 64; tensor memory: 64 columns of A per stage).
  static constexpr int kStagesDefault =
      kSplit == 6 ? (BN > 128 ? 2 : 3)
                  : (kSplit == 3 ? (A_TMEM ? (BN <= 64 ? 6 : (BN <= 96 ? 5 : 4)) : 3) : 6) - (kEpiWarps == 8 ? 1 : 0);
  static constexpr int kStages = kStagesT > 0 ? kStagesT : kStagesDefault;
  // private patches of the epilogue warps: 32 x 36 floats (padded, conflict-free transposed reads) for the register -> global
  // epilogues; 32 x 32 floats in the 128-byte-swizzled box layout (1 KB aligned) when the tile leaves through a TMA store
  static constexpr int kEpiPatchBytes = kTmaEpi ? 4096 : 32 * 36 * 4;
  static constexpr int kEpiBytes = kEpiWarps * kEpiPatchBytes;
  static constexpr int kSmemBytes = kStages * kStageBytes + kEpiBytes + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int kAccStride = BN;                                   // accumulator a lives in TMEM columns [a*BN, a*BN + BN)
  static constexpr uint32_t kTmemA = 2 * BN;                              // first column of the staged A operand (hi | lo per stage)
  static constexpr int kTmemNeed = (kSplit == 3 && A_TMEM) ? 2 * BN + kStages * 64 : 2 * BN;
  static constexpr int kTmemCols = kTmemNeed <= 32 ? 32 : kTmemNeed <= 64 ? 64 : kTmemNeed <= 128 ? 128 : kTmemNeed <= 256 ? 256 : 512;
  static_assert(kTmemNeed <= 512, "tensor memory: 2 accumulators + 64 columns of A per stage must fit 512 columns");
  static_assert(kSmemBytes <= 227 * 1024, "shared memory per CTA");
  static_assert(3 * kStages * 8 + 4 * 8 + 8 <= 256, "barrier block");
};

// ---- PTX wrappers -----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// One lane of the (fully converged) warp: tcgen05.mma / TMA / commit are issued from uniform registers, and ptxas only
// keeps their operands there when the surrounding control flow is warp-uniform.  Under `if (lane == 0)` it emitted a
// divergence "waterfall" around every UTCHMMA (ELECT, 4 x R2UR.BROADCAST, UTCHMMA, BRA.U.ANY: ~90 cycles per MMA, more
// than the 64 cycles the tensor core needs for 128x128x8), which made every k-block cost 0.58 us whatever the tile width.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y)
               : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int z) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
      : "memory");
}
// TMA store of a [1 x 32 x 32] fp32 box (slab z, rows y.., columns x..) from shared memory; rows / columns past the tensor's extent
// are clipped by the hardware
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int x, int y, int z) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(smem_u32(src)), "r"(x),
               "r"(y), "r"(z)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand from tensor memory (lane = row, 8 consecutive 32-bit columns = the K=8 tf32 values of that row)
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// fp32 -> the tf32 value the tensor core actually uses: kind::tf32 TRUNCATES the 13 low mantissa bits (measured with
// csrc/dev/tc_gemm_test: lo = x - trunc(x) gives 1e-6 relative error, lo = x - rna(x) gives 7e-4)
__device__ __forceinline__ float tf32_lo(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- descriptors --------------------------------------------------------------------------------------
// shared-memory matrix descriptor, SWIZZLE_128B, sm_100 version bit set (cute::UMMA::SmemDescriptor layout)
//   layout_type 2 = SWIZZLE_128B (K-major operands), 1 = SWIZZLE_128B_BASE32B (the only legal layout for MN-major
//   tf32 operands: 128-byte rows whose four 32-byte segments are XOR-ed with (row & 3), TMA SWIZZLE_128B_ATOM_32B)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // version = 1 (Blackwell)
  d |= (uint64_t)layout_type << 61;
  return d;
}
// instruction descriptor for kind::tf32, fp32 accumulate (cute::UMMA::InstrDescriptor layout)
__host__ __device__ constexpr uint32_t instr_desc_tf32(bool a_mn_major, bool b_mn_major, int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// instruction descriptor for kind::f16 with bf16 operands, fp32 accumulate, both operands K-major
__host__ __device__ constexpr uint32_t instr_desc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- epilogues -------------------------------------------------------------------------------------------
enum EpiKind { kEpiStore = 0, kEpiRankExtract = 1, kEpiRankCount = 2, kEpiRankScreen = 3, kEpiPairDecide = 4, kEpiStoreTma = 5 };

struct EpiPartial {   // raw fp32 tile -> slab z of [splits][M*N]
  static constexpr int kKind = kEpiStore;
  float* part; long long stride;
  __device__ __forceinline__ float* row_ptr(int m, int N, int z) const { return part + (size_t)z * stride + (size_t)m * N; }
  __device__ __forceinline__ float coef() const { return 1.0f; }
  __device__ __forceinline__ float apply(float acc, float, float) const { return acc; }
  __device__ __forceinline__ const float* src_row(int) const { return nullptr; }
};
struct EpiPartialTma {   // the same slabs, written by TMA stores (Maps::c: 3-D map {N, M, splits}, box 32 x 32 x 1, SWIZZLE_128B): each
  static constexpr int kKind = kEpiStoreTma;   // warp parks a 32 x 32 block in shared memory and one instruction ships it; rows >= M
  static constexpr bool kTmaStore = true;      // are clipped by the hardware.  Needs N % 4 == 0 and 16-byte aligned slabs.
  float* part; long long stride;
  __device__ __forceinline__ float* row_ptr(int m, int N, int z) const { return part + (size_t)z * stride + (size_t)m * N; }
  __device__ __forceinline__ float coef() const { return 1.0f; }
  __device__ __forceinline__ float apply(float acc, float, float) const { return acc; }
  __device__ __forceinline__ const float* src_row(int) const { return nullptr; }
};
template <class E, class = void> struct epi_uses_tma : std::false_type {};
template <class E> struct epi_uses_tma<E, std::void_t<decltype(E::kTmaStore)>> : std::bool_constant<E::kTmaStore> {};
struct EpiScale {     // C = alpha * acc
  static constexpr int kKind = kEpiStore;
  float* C; int ldc; float alpha;
  __device__ __forceinline__ float* row_ptr(int m, int, int) const { return C + (size_t)m * ldc; }
  __device__ __forceinline__ float coef() const { return alpha; }
  __device__ __forceinline__ float apply(float acc, float, float c) const { return c * acc; }
  __device__ __forceinline__ const float* src_row(int) const { return nullptr; }
};
struct EpiAxpyTC {    // dst = src - (*lr) * acc   (src nullable = 0)
  static constexpr int kKind = kEpiStore;
  const float* src; float* dst; int ld; const float* lr;
  __device__ __forceinline__ float* row_ptr(int m, int, int) const { return dst + (size_t)m * ld; }
  __device__ __forceinline__ const float* src_row(int m) const { return src ? src + (size_t)m * ld : nullptr; }
  __device__ __forceinline__ float coef() const { return *lr; }
  __device__ __forceinline__ float apply(float acc, float srcv, float c) const { return srcv - c * acc; }
};

// Retrieval epilogues: the score tile S = alpha * acc never leaves the SM.
//  pass 1 (extract): only the tiles that contain a ground-truth pair are computed; the scores at the ground-truth
//                    positions are written out (gt_val[e] for CSR entry e of image m; col_val[t] = S[txt2img[t], t]).
//  pass 2 (count):   every tile; per row   rank_row[m] += #{n : S > thr or (S == thr and n < thr_idx)}   (image -> text)
//                                per column rank_col[n] += #{m : S > thr or (S == thr and m < thr_idx)}  (text -> image)
//                    integer atomics: order-independent, and both passes see bit-identical tile values (same kernel,
//                    same K order), so the result equals ranking the materialised matrix.
struct EpiRankExtract {
  static constexpr int kKind = kEpiRankExtract;
  float alpha;
  __device__ __forceinline__ float coef() const { return alpha; }
  __device__ __forceinline__ const float* src_row(int) const { return nullptr; }
  __device__ __forceinline__ float* row_ptr(int, int, int) const { return nullptr; }
  __device__ __forceinline__ float apply(float a, float, float) const { return a; }
  const int32_t* gt_ptr; const int32_t* gt_idx; float* gt_val;      // img2txt CSR and its score slots
  const int32_t* col_gt; float* col_val;                            // txt2img and its score slots
  int col_offset;                                                   // gt_idx holds GLOBAL caption ids; this product covers
};                                                                  // captions [col_offset, col_offset + N) (caption shards)
struct EpiRankCount {
  static constexpr int kKind = kEpiRankCount;
  float alpha;
  __device__ __forceinline__ float coef() const { return alpha; }
  __device__ __forceinline__ const float* src_row(int) const { return nullptr; }
  __device__ __forceinline__ float* row_ptr(int, int, int) const { return nullptr; }
  __device__ __forceinline__ float apply(float a, float, float) const { return a; }
  const float* row_thr; const int32_t* row_thr_idx; int32_t* row_cnt;   // [M]
  const float* col_thr; const int32_t* col_thr_idx; int32_t* col_cnt;   // [N]
};

//  screened count: the tile comes from a CHEAPER product (bf16x3) whose distance to the 3xTF32 score of the same pair is
//                    bounded by *band (one scalar: alpha * eps(K) * max |img row| * max |txt row|).  Entries clearly ahead of the
//                    threshold are counted, entries clearly behind are not, entries inside the band are appended to a list
//                    (pair + direction) and decided exactly afterwards.  Ties (and the ground-truth entries themselves) always
//                    fall inside the band.  The hot loop is branch-free: two compares and two adds per entry and direction;
//                    only a chunk that holds a borderline entry (#{v >= thr - band} != #{v > thr + band}) is rescanned.
struct AmbiguousPair { int32_t m, n, dir; };      // dir 1: image -> text (row count of image m), 2: text -> image (column count of n)
struct EpiRankScreen {
  static constexpr int kKind = kEpiRankScreen;
  float alpha;
  __device__ __forceinline__ float coef() const { return alpha; }
  __device__ __forceinline__ const float* src_row(int) const { return nullptr; }
  __device__ __forceinline__ float* row_ptr(int, int, int) const { return nullptr; }
  __device__ __forceinline__ float apply(float a, float, float) const { return a; }
  const float* row_thr; const int32_t* row_thr_idx; int32_t* row_cnt;    // [M]
  const float* col_thr; const int32_t* col_thr_idx; int32_t* col_cnt;    // [N]
  const float* band;                                                     // device scalar
  AmbiguousPair* list; int32_t* list_count; int32_t list_cap;
  __device__ __forceinline__ void push(int m, int n, int dir) const {
    const int at = atomicAdd(list_count, 1);
    if (at < list_cap) list[at] = AmbiguousPair{m, n, dir};
  }
};
//  pair decide:      tile t of a [P, D] x [P, D]^T product of GATHERED rows (pair p = image m_p against caption n_p): only the
//                    diagonal is read; S = alpha * acc[p, p] is the 3xTF32 score of the pair, compared like the exact count.
struct EpiPairDecide {
  static constexpr int kKind = kEpiPairDecide;
  float alpha;
  __device__ __forceinline__ float coef() const { return alpha; }
  __device__ __forceinline__ const float* src_row(int) const { return nullptr; }
  __device__ __forceinline__ float* row_ptr(int, int, int) const { return nullptr; }
  __device__ __forceinline__ float apply(float a, float, float) const { return a; }
  const AmbiguousPair* list; const int32_t* n_pairs;
  const float* row_thr; const int32_t* row_thr_idx; int32_t* row_cnt;
  const float* col_thr; const int32_t* col_thr_idx; int32_t* col_cnt;
};

#ifdef VLDD_TC_TIMELINE
__device__ long long g_timeline[148 * 10 * 16];   // [cta][slot]: globaltimer at phase boundaries (developer harness only)
__device__ __forceinline__ long long gtime() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define TL(slot) do { if (lane == 0 && warp < 10) g_timeline[(blockIdx.x * 10 + warp) * 16 + (slot)] = gtime(); } while (0)
#define TL_ITEM(base, item) do { if ((item) < 4) TL((base) + 2 * (item)); } while (0)     // per work item stamps, first 4 items
#else
#define TL(slot) do {} while (0)
#define TL_ITEM(base, item) do {} while (0)
#endif

#ifdef VLDD_TC_DEBUG
__constant__ uint32_t g_dbg[4];   // {lbo, sbo, step, unused} overrides for MN-major operands (developer harness only)
#endif

struct Maps {
  CUtensorMap a0, b0, a1, b1;
  CUtensorMap c;      // output map of the TMA-store epilogue (EpiPartialTma), unused otherwise
};

// ---- the kernel ----------------------------------------------------------------------------------------
// Work item w (one per (m-tile, n-tile, k-split)) -> CTA blockIdx.x, blockIdx.x + gridDim.x, ... (persistent loop).
struct WorkItem { int m0, n0, z, kb_begin, n_kb; };

template <bool A_KMAJOR, bool B_KMAJOR, int kSplit, class Epi, int kStagesT = 0, int BN = 128, int kEpiWarps = 4>
__global__ void __launch_bounds__(kBaseThreads + 32 * kEpiWarps, 1)
tc_gemm_kernel(const __grid_constant__ Maps maps, int M, int N, int K0, int K1, int splits, Epi epi,
               const int* __restrict__ work_list, const int* __restrict__ work_count, int old_mask) {
  static_assert(BN % 32 == 0 && BN >= 64 && (BN <= 128 || (kSplit == 6 && BN <= 256)),
                "N tile: 64, 96 or 128 (32-column epilogue chunks, MN-major boxes); up to 256 for the bf16 mode");
  static_assert(kSplit != 6 || (A_KMAJOR && B_KMAJOR), "bf16x3 operands are K-major");
  constexpr bool A_TMEM = kSplit == 3;                 // hi/lo of the A tile are staged in tensor memory (either major)
  using C = Cfg<kSplit, A_TMEM, kStagesT, BN, kEpiWarps, epi_uses_tma<Epi>::value>;
  constexpr int TILE_B = C::kTileB;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* epi_stage = reinterpret_cast<float*>(smem + C::kStages * C::kStageBytes);        // one patch per epilogue warp
  constexpr int kPatchFloats = C::kEpiPatchBytes / 4;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes + C::kEpiBytes);
  uint64_t* full = bars;                       // TMA landed
  uint64_t* ready = bars + C::kStages;         // split done (kSplit == 3 only)
  uint64_t* empty = bars + 2 * C::kStages;     // MMAs reading the stage have completed
  uint64_t* tmem_full = bars + 3 * C::kStages;       // [2] accumulator complete
  uint64_t* tmem_empty = bars + 3 * C::kStages + 2;  // [2] accumulator drained by the epilogue warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * C::kStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  TL(0);
  const int tiles_m = (M + BM - 1) / BM, tiles_n = (N + BN - 1) / BN;
  constexpr int BKE = kSplit == 6 ? 64 : BK;               // elements of K per k-block (128 bytes either way)
  const int nkb0 = (K0 + BKE - 1) / BKE, nkb1 = (K1 + BKE - 1) / BKE, nkb = nkb0 + nkb1;
  const int per = (nkb + splits - 1) / splits;
  // optional device-side work list (tile indices chosen by an earlier kernel, e.g. "tiles holding a ground-truth pair");
  // it is written by a predecessor kernel, so it may only be read after pdl_wait() -- see below
  int total_work = tiles_m * tiles_n * splits;
  constexpr bool kRaster = Epi::kKind != kEpiStore && Epi::kKind != kEpiStoreTma;      // the ranking kernels (many m-tiles)
  auto decode = [&](int w) {
    WorkItem it;
    if (work_list != nullptr) w = work_list[w];
    const int tile = w % (tiles_m * tiles_n);
    it.z = w / (tiles_m * tiles_n);
    if (kRaster && work_list == nullptr && tiles_m > 1) {
      // bands of kRasterM m-tiles, m fastest inside a band: the CTAs running at one time cover up to kRasterM m-tiles x a few
      // n-tiles, so the A band stays in L2 for a whole sweep over n and every B tile is fetched from DRAM once per BAND instead
      // of once per m-tile (25k x 125k x 768 screen: 60 GB of DRAM reads with n-fastest order, B = 384 MB > L2)
#ifndef VLDD_RASTER_M
#define VLDD_RASTER_M 64      // 25k x 125k screen: 8 -> 9.94 ms, 16 -> 9.60-9.77, 32 -> 9.47, 64 -> 9.34-9.43, 128 -> 9.46, all m-tiles -> 10.82
#endif
      constexpr int kRasterM = VLDD_RASTER_M;
      const int band = tile / (kRasterM * tiles_n), r = tile - band * kRasterM * tiles_n;
      const int h = min(kRasterM, tiles_m - band * kRasterM);
      it.m0 = (band * kRasterM + r % h) * BM;
      it.n0 = (r / h) * BN;
    } else {
      it.n0 = (tile % tiles_n) * BN;          // one m-tile (or listed tiles): consecutive CTAs take consecutive n-tiles
      it.m0 = (tile / tiles_n) * BM;
    }
    it.kb_begin = it.z * per;
    it.n_kb = max(min(nkb, it.kb_begin + per) - it.kb_begin, 0);
    return it;
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.a0);
    tma_prefetch_desc(&maps.b0);
    if (K1 > 0) { tma_prefetch_desc(&maps.a1); tma_prefetch_desc(&maps.b1); }
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&ready[s], 128);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 32 * kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above is independent of the preceding kernel: under programmatic dependent launch it overlaps that
  // kernel's tail; from here on global memory written by predecessors is read.
  // `old_mask` (bit 0: A operands, bit 1: B operands) marks operands that were complete before the programmatic
  // predecessor even started (weights written a step ago, gathered minibatches): the producer issues their TMA loads for
  // the first pipeline stages BEFORE griddepcontrol.wait, so the weight fetch overlaps the predecessor's execution.
  // (Safe: every kernel of the library triggers its dependents only after its own wait returned, so when this grid
  // runs, everything older than its immediate predecessor has completed and is visible.)
  TL(1);
  if (warp != 0) {
    pdl_enter();
    if (work_count != nullptr) total_work = *work_count;
  }
  TL(2);

  // byte offsets of the tiles inside a stage
  constexpr int kHalfN = ((BN / 32 + 1) / 2) * 32;                   // column split of the last item's drain (64 of 96 / 128)
  //   A in TMEM: [A | B | B_lo]      A in smem: [A | A_lo | B | B_lo]   (A / A_lo and B / B_lo adjacent: one split loop each)
  constexpr int OFF_A = 0;
  constexpr int OFF_ALO = TILE_BYTES;                               // only when A is fed from shared memory
  constexpr int OFF_B = (kSplit == 6) ? 2 * TILE_BYTES : ((A_TMEM || kSplit != 3) ? TILE_BYTES : 2 * TILE_BYTES);
  constexpr int OFF_BLO = OFF_B + TILE_B;

  // Drain columns [c_begin, c_end) of accumulator `acc` for work item `it`: TMEM lane quadrant is fixed by warp index % 4.
  // Each thread drains 32 columns of its own row (tcgen05.ld 32x32b.x32), the warp transposes the 32x32 block through a
  // private shared-memory patch and then touches global memory as 4 rows x 128 contiguous bytes per instruction.
  [[maybe_unused]] int tl_item = 0;
  auto drain = [&](const WorkItem& it, int acc, uint32_t parity, int c_begin, int c_end, int c_step, float* stage, bool release) {
    const int quad = warp & 3;
    constexpr int LDS = 36;                                    // padded row stride (floats): conflict-free float4 access
    const int rsub = lane >> 3, q4 = (lane & 7) * 4;
    const float coef = epi.coef();
    const int m0 = it.m0, n0 = it.n0, z = it.z;
    // `src` operand of the axpy epilogue: 8 independent 128-bit loads per thread per 32-column chunk, issued one chunk
    // ahead (the first chunk before the accumulator is even ready): latency hides behind the MMAs / TMEM drain
    float4 sv[8];
    auto load_src = [&](int c) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int m = m0 + quad * 32 + i * 4 + rsub, nb = n0 + c + q4;
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#ifdef VLDD_EXP_NOSRC
        const float* src = nullptr;
#else
        const float* src = (m < M && nb < N) ? epi.src_row(m) : nullptr;
#endif
        if (src != nullptr) {
          if (nb + 3 < N && ((reinterpret_cast<uintptr_t>(src + nb) & 15) == 0)) {
            t = *reinterpret_cast<const float4*>(src + nb);
          } else {
            t.x = src[nb];
            if (nb + 1 < N) t.y = src[nb + 1];
            if (nb + 2 < N) t.z = src[nb + 2];
            if (nb + 3 < N) t.w = src[nb + 3];
          }
        }
        sv[i] = t;
      }
    };
    if constexpr (Epi::kKind == kEpiStore) load_src(c_begin);
    mbar_wait(&tmem_full[acc], parity);
    tc_fence_after();
    TL_ITEM(8, tl_item);
    [[maybe_unused]] int row_count = 0;
    [[maybe_unused]] float row_thr = 0.f;
    [[maybe_unused]] int row_thr_idx = -1;
    const int my_m = m0 + quad * 32 + lane;                     // the accumulator row this thread drains
    if constexpr (Epi::kKind == kEpiRankCount) {
      if (my_m < M) { row_thr = epi.row_thr[my_m]; row_thr_idx = epi.row_thr_idx[my_m]; }
    }
    // screened count: thresholds with the band folded in; +inf for a row / column without ground truth (never counted, never
    // borderline).  The column thresholds of the NEXT chunk are fetched while the current one is processed.
    [[maybe_unused]] float rhi = INFINITY, rlo = INFINITY, band = 0.f, cthr_next = INFINITY;
    auto col_thr_of = [&](int c) {
      const int n = n0 + c + lane;
      float t = INFINITY;
      if constexpr (Epi::kKind == kEpiRankScreen || Epi::kKind == kEpiRankCount) {
        if (c < c_end && n < N && epi.col_thr_idx[n] >= 0) t = epi.col_thr[n];
      }
      return t;
    };
    [[maybe_unused]] auto col_idx_of = [&](int c) {
      const int n = n0 + c + lane;
      int t = -1;
      if constexpr (Epi::kKind == kEpiRankCount) {
        if (c < c_end && n < N) t = epi.col_thr_idx[n];
      }
      return t;
    };
    [[maybe_unused]] int cidx_next = -1;
    if constexpr (Epi::kKind == kEpiRankCount) { cthr_next = col_thr_of(c_begin); cidx_next = col_idx_of(c_begin); }
    if constexpr (Epi::kKind == kEpiRankScreen) {
      band = *epi.band;
      if (my_m < M) { const float t = epi.row_thr[my_m]; rhi = t + band; rlo = t - band; row_thr_idx = epi.row_thr_idx[my_m]; }   // +inf: no ground truth
      cthr_next = col_thr_of(c_begin);
    }
#pragma unroll 1
    for (int c = c_begin; c < c_end; c += c_step) {
      float v[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * C::kAccStride + c);
      if (it.n_kb > 0) {
        tmem_ld_32x32(taddr, v);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
      if (release && c + c_step >= c_end) {      // this warp's last chunk is read: hand the TMEM buffer back to the MMA warp
        tc_fence_before();
        mbar_arrive(&tmem_empty[acc]);
      }
      if constexpr (Epi::kKind != kEpiStore && Epi::kKind != kEpiStoreTma) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= coef;               // the score, exactly as EpiScale would have stored it
      }
      [[maybe_unused]] float cthr = INFINITY;
      [[maybe_unused]] int cidx = -1;
      if constexpr (Epi::kKind == kEpiRankCount) {
        // branch-free hot loop: #{v > thr} and #{v == thr}; only a chunk with an exact tie runs the index tie-break
        cthr = cthr_next; cidx = cidx_next;
        cthr_next = col_thr_of(c + c_step); cidx_next = col_idx_of(c + c_step);
        if (my_m < M) {                        // (a row without ground truth carries thr = +inf: nothing is ahead of it)
          const int jmax = N - (n0 + c);
          int gt = 0, eq = 0;
          if (jmax >= 32) {
#pragma unroll
            for (int j = 0; j < 32; ++j) { gt += v[j] > row_thr; eq += v[j] == row_thr; }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) { gt += (j < jmax) && v[j] > row_thr; eq += (j < jmax) && v[j] == row_thr; }
          }
          row_count += gt;
          if (eq) {
#pragma unroll
            for (int j = 0; j < 32; ++j) row_count += (j < jmax) && v[j] == row_thr && (n0 + c + j) < row_thr_idx;
          }
        }
      }
      if constexpr (Epi::kKind == kEpiRankScreen) {
        cthr = cthr_next;
        cthr_next = col_thr_of(c + c_step);
        // image -> text: this thread's row against the 32 captions of the chunk
        const int jmax = N - (n0 + c);                   // >= 32 except in the last tile of a row of tiles
        int gt = 0, ge = 0;
        if (jmax >= 32) {
#pragma unroll
          for (int j = 0; j < 32; ++j) { gt += v[j] > rhi; ge += v[j] >= rlo; }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) { gt += (j < jmax) && v[j] > rhi; ge += (j < jmax) && v[j] >= rlo; }
        }
        row_count += gt;
        if (ge != gt) {                                  // rare: a borderline entry in this row of the chunk
#pragma unroll
          for (int j = 0; j < 32; ++j)      // (the threshold entry itself is borderline by construction and never ahead of itself: not listed)
            if (j < jmax && v[j] >= rlo && !(v[j] > rhi) && n0 + c + j != row_thr_idx) epi.push(my_m, n0 + c + j, 1);
        }
      }
      if constexpr (Epi::kKind == kEpiStoreTma) {
        // the 32 x 32 block goes to shared memory in the box layout of the output map (row = 128 bytes, 16-byte chunk q of row r
        // at q ^ (r & 7)) and leaves with ONE TMA store; the previous chunk's store must have finished READING the patch first
        if (elect_one()) bulk_wait_read0();
        __syncwarp();
        uint8_t* patch = reinterpret_cast<uint8_t*>(stage) + lane * 128;
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<float4*>(patch + ((q ^ (lane & 7)) << 4)) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        fence_proxy_async();
        __syncwarp();
        if (elect_one()) {
          tma_store_3d(&maps.c, stage, n0 + c, m0 + quad * 32, z);
          bulk_commit();
        }
        __syncwarp();
      } else {
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(stage + lane * LDS + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        __syncwarp();
      }
      if constexpr (Epi::kKind == kEpiStoreTma) {
        // (shipped above)
      } else if constexpr (Epi::kKind == kEpiStore) {
        float4 cur[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) cur[i] = sv[i];
        if (c + c_step < c_end) load_src(c + c_step);
        const int nb = n0 + c + q4;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int m = m0 + quad * 32 + i * 4 + rsub;
          if (m >= M || nb >= N) continue;
          const float4 a = *reinterpret_cast<const float4*>(stage + (i * 4 + rsub) * LDS + q4);
          float* out = epi.row_ptr(m, N, z) + nb;
          const float4 o = make_float4(epi.apply(a.x, cur[i].x, coef), epi.apply(a.y, cur[i].y, coef),
                                       epi.apply(a.z, cur[i].z, coef), epi.apply(a.w, cur[i].w, coef));
#ifdef VLDD_EXP_NOSTORE
          if (o.x == 123.456f) *reinterpret_cast<float4*>(out) = o;
          else if (false) {
#else
          if (nb + 3 < N && ((reinterpret_cast<uintptr_t>(out) & 15) == 0)) {
            *reinterpret_cast<float4*>(out) = o;
          } else {
#endif
            out[0] = o.x;
            if (nb + 1 < N) out[1] = o.y;
            if (nb + 2 < N) out[2] = o.z;
            if (nb + 3 < N) out[3] = o.w;
          }
        }
      } else if constexpr (Epi::kKind == kEpiRankExtract) {
        // rows: the ground-truth captions of image my_m that fall into this 32-column chunk
        if (my_m < M) {
          for (int e = epi.gt_ptr[my_m]; e < epi.gt_ptr[my_m + 1]; ++e) {
            const int cidx = epi.gt_idx[e] - epi.col_offset - (n0 + c);
            if (cidx >= 0 && cidx < 32) epi.gt_val[e] = stage[lane * LDS + cidx];
          }
        }
        // columns: lane j owns caption n; its image row may sit in this warp's 32-row quadrant
        const int n = n0 + c + lane;
        if (n < N) {
          const int g = epi.col_gt[n] - (m0 + quad * 32);
          if (g >= 0 && g < 32) epi.col_val[n] = stage[g * LDS + lane];
        }
      } else if constexpr (Epi::kKind == kEpiRankScreen) {   // text -> image: lane j owns caption n, 32 image rows of the quadrant
        const int n = n0 + c + lane;
        const float chi = cthr + band, clo = cthr - band;      // +inf: no ground truth / past the last caption
        const int mbase = m0 + quad * 32;
        const int nvalid = min(32, M - mbase);
        int gt = 0, ge = 0;
        if (nvalid == 32) {
#pragma unroll 16
          for (int rr = 0; rr < 32; ++rr) { const float sv_ = stage[rr * LDS + lane]; gt += sv_ > chi; ge += sv_ >= clo; }
        } else {
          for (int rr = 0; rr < nvalid; ++rr) { const float sv_ = stage[rr * LDS + lane]; gt += sv_ > chi; ge += sv_ >= clo; }
        }
        if (gt) atomicAdd(epi.col_cnt + n, gt);
        if (ge != gt) {
          const int self = epi.col_thr_idx[n];                 // (n < N here: past the last caption the band is infinite and ge == gt)
          for (int rr = 0; rr < nvalid; ++rr) {
            const float sv_ = stage[rr * LDS + lane];
            if (sv_ >= clo && !(sv_ > chi) && mbase + rr != self) epi.push(mbase + rr, n, 2);
          }
        }
      } else if constexpr (Epi::kKind == kEpiPairDecide) {   // diagonal tiles of the gathered product: pair p = row p = column p
        if (c == quad * 32 && m0 == n0) {
          const int p = m0 + quad * 32 + lane;
          if (p < *epi.n_pairs) {
            const float sc = stage[lane * LDS + lane];
            const AmbiguousPair a = epi.list[p];
            if (a.dir == 1) {
              const float thr = epi.row_thr[a.m];
              if (sc > thr || (sc == thr && a.n < epi.row_thr_idx[a.m])) atomicAdd(epi.row_cnt + a.m, 1);
            } else {
              const float thr = epi.col_thr[a.n];
              if (sc > thr || (sc == thr && a.m < epi.col_thr_idx[a.n])) atomicAdd(epi.col_cnt + a.n, 1);
            }
          }
        }
      } else {   // kEpiRankCount, column direction: lane j counts the 32 rows of this quadrant for caption n
        const int n = n0 + c + lane;
        const int mbase = m0 + quad * 32;
        const int nvalid = min(32, M - mbase);
        int gt = 0, eq = 0;                    // cthr = +inf: no ground truth / past the last caption -> both stay 0
        if (nvalid == 32) {
#pragma unroll 16
          for (int rr = 0; rr < 32; ++rr) { const float sv_ = stage[rr * LDS + lane]; gt += sv_ > cthr; eq += sv_ == cthr; }
        } else {
          for (int rr = 0; rr < nvalid; ++rr) { const float sv_ = stage[rr * LDS + lane]; gt += sv_ > cthr; eq += sv_ == cthr; }
        }
        if (eq) {
          for (int rr = 0; rr < nvalid; ++rr) gt += stage[rr * LDS + lane] == cthr && mbase + rr < cidx;
        }
        if (gt) atomicAdd(epi.col_cnt + n, gt);
      }
    }
    if constexpr (Epi::kKind == kEpiRankCount || Epi::kKind == kEpiRankScreen) {
      if (row_count) atomicAdd(epi.row_cnt + my_m, row_count);
    }
    // (TMA stores: only the READS of the patch have to be over before the next chunk reuses it -- checked there; the writes
    //  themselves are awaited once, before the CTA exits)
    TL_ITEM(9, tl_item);
    ++tl_item;
  };

  if (warp == 0) {
    // ===== TMA producer: the whole warp walks the loop (uniform control flow), one elected lane issues =====
    auto issue = [&](const WorkItem& it, int i, int s, bool load_a, bool load_b) {
      const int kb = it.kb_begin + i;
      if constexpr (kSplit == 6) {        // four bf16 tiles: A_hi | A_lo | B_hi | B_lo, 64 elements of K each
        uint8_t* st0 = smem + s * C::kStageBytes;
        const int k = kb * 64;
        tma_load_2d(st0 + OFF_A, &maps.a0, &full[s], k, it.m0);
        tma_load_2d(st0 + OFF_ALO, &maps.a1, &full[s], k, it.m0);
        tma_load_2d(st0 + OFF_B, &maps.b0, &full[s], k, it.n0);
        tma_load_2d(st0 + OFF_BLO, &maps.b1, &full[s], k, it.n0);
        return;
      }
      const bool seg1 = kb >= nkb0;
      const int k = (seg1 ? kb - nkb0 : kb) * BK;
      const CUtensorMap* ma = seg1 ? &maps.a1 : &maps.a0;
      const CUtensorMap* mb = seg1 ? &maps.b1 : &maps.b0;
      uint8_t* sa = smem + s * C::kStageBytes + OFF_A;
      uint8_t* sb = smem + s * C::kStageBytes + OFF_B;
      if (load_a) { if (A_KMAJOR) tma_load_2d(sa, ma, &full[s], k, it.m0); else tma_load_3d(sa, ma, &full[s], 0, k, it.m0 / 32); }
      if (load_b) { if (B_KMAJOR) tma_load_2d(sb, mb, &full[s], k, it.n0); else tma_load_3d(sb, mb, &full[s], 0, k, it.n0 / 32); }
    };
    // early loads of the operands that do not depend on the programmatic predecessor (first item, first kStages k-blocks)
    int pre = 0;
    const bool a_old = (old_mask & 1) != 0, b_old = (old_mask & 2) != 0;
    if (old_mask != 0 && work_list == nullptr && work_count == nullptr && (int)blockIdx.x < total_work) {
      const WorkItem it0 = decode(blockIdx.x);
      pre = min(C::kStages, it0.n_kb);
      for (int i = 0; i < pre; ++i) {
        if (elect_one()) {
          mbar_expect_tx(&full[i], TILE_BYTES + TILE_B);        // both tiles of the stage; the rest follows after the wait
          issue(it0, i, i, a_old, b_old);
        }
        __syncwarp();
      }
    }
    pdl_enter();
    if (work_count != nullptr) total_work = *work_count;
    int g = 0;                                                    // k-blocks issued so far (across work items)
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      const WorkItem it = decode(w);
      for (int i = 0; i < it.n_kb; ++i, ++g) {
        const int s = g % C::kStages, round = g / C::kStages;
        if (g < pre) {
          if (elect_one()) issue(it, i, s, !a_old, !b_old);
        } else {
          if (round > 0) mbar_wait(&empty[s], (round - 1) & 1);
          if (elect_one()) {
            mbar_expect_tx(&full[s], kSplit == 6 ? 2 * (TILE_BYTES + TILE_B) : TILE_BYTES + TILE_B);
            issue(it, i, s, true, true);
          }
        }
        __syncwarp();
        if (g == 0) TL(3);
      }
    }
    TL(4);
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp walks the loop (uniform control flow), one elected lane issues =====
    {
      constexpr uint32_t idesc = instr_desc_tf32(A_TMEM ? false : !A_KMAJOR, !B_KMAJOR, BM, BN);   // TMEM A is row-per-lane
      // K-major  (SWIZZLE_128B):         8-row groups 1024 B apart (SBO), k-step of 8 floats = +32 B inside the row
      // MN-major (SWIZZLE_128B_BASE32B): 32-wide MN chunks 4096 B apart (LBO), 4-k-row groups 512 B apart (SBO),
      //                                  k-step of 8 k-rows = +1024 B
#ifdef VLDD_TC_DEBUG
      const uint32_t a_lbo = A_KMAJOR ? 16 : g_dbg[0], a_sbo = A_KMAJOR ? 1024 : g_dbg[1], a_step = A_KMAJOR ? 32 : g_dbg[2];
      const uint32_t b_lbo = B_KMAJOR ? 16 : g_dbg[0], b_sbo = B_KMAJOR ? 1024 : g_dbg[1], b_step = B_KMAJOR ? 32 : g_dbg[2];
#else
      constexpr uint32_t a_lbo = A_KMAJOR ? 16 : 4096, a_sbo = A_KMAJOR ? 1024 : 512, a_step = A_KMAJOR ? 32 : 1024;
      constexpr uint32_t b_lbo = B_KMAJOR ? 16 : 4096, b_sbo = B_KMAJOR ? 1024 : 512, b_step = B_KMAJOR ? 32 : 1024;
#endif
      constexpr uint32_t a_lt = A_KMAJOR ? 2 : 1, b_lt = B_KMAJOR ? 2 : 1;
      int g = 0, local = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++local) {
        const WorkItem it = decode(w);
        const int acc = local & 1;                                  // accumulator buffer: TMEM columns [BN*acc, +BN)
        if (local >= 2) mbar_wait(&tmem_empty[acc], ((local >> 1) - 1) & 1);
        tc_fence_after();
        const uint32_t tacc = tmem_base + acc * C::kAccStride;
        uint32_t accumulate = 0;
        for (int i = 0; i < it.n_kb; ++i, ++g) {
          const int s = g % C::kStages, round = g / C::kStages;
          mbar_wait(kSplit == 3 ? &ready[s] : &full[s], round & 1);
          if (g == 0) TL(3);
          tc_fence_after();
          const uint32_t st = smem_u32(smem + s * C::kStageBytes);
          const uint32_t sa = st + OFF_A, sb = st + OFF_B, sal = st + OFF_ALO, sbl = st + OFF_BLO;
          const uint32_t ta_hi = tmem_base + C::kTmemA + s * 64, ta_lo = ta_hi + 32;   // TMEM columns of this stage's A
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t db = smem_desc(sb + k * b_step, b_lbo, b_sbo, b_lt);
              if constexpr (kSplit == 6) {
                constexpr uint32_t idesc16 = instr_desc_bf16(BM, BN);
                const uint64_t dbl = smem_desc(sbl + k * b_step, b_lbo, b_sbo, b_lt);
                const uint64_t da = smem_desc(sa + k * a_step, a_lbo, a_sbo, a_lt);
                const uint64_t dal = smem_desc(sal + k * a_step, a_lbo, a_sbo, a_lt);
                umma_bf16(tacc, dal, db, idesc16, accumulate);
                umma_bf16(tacc, da, dbl, idesc16, 1);
                umma_bf16(tacc, da, db, idesc16, 1);
              } else if (kSplit == 3) {
                const uint64_t dbl = smem_desc(sbl + k * b_step, b_lbo, b_sbo, b_lt);
                if (A_TMEM) {
                  umma_tf32_ts(tacc, ta_lo + k * UMMA_K, db, idesc, accumulate);
                  umma_tf32_ts(tacc, ta_hi + k * UMMA_K, dbl, idesc, 1);
                  umma_tf32_ts(tacc, ta_hi + k * UMMA_K, db, idesc, 1);
                } else {
                  const uint64_t da = smem_desc(sa + k * a_step, a_lbo, a_sbo, a_lt);
                  const uint64_t dal = smem_desc(sal + k * a_step, a_lbo, a_sbo, a_lt);
                  umma_tf32(tacc, dal, db, idesc, accumulate);
                  umma_tf32(tacc, da, dbl, idesc, 1);
                  umma_tf32(tacc, da, db, idesc, 1);
                }
              } else {
                const uint64_t da = smem_desc(sa + k * a_step, a_lbo, a_sbo, a_lt);
                umma_tf32(tacc, da, db, idesc, accumulate);
              }
              accumulate = 1;
            }
            umma_commit(&empty[s]);
          }
          __syncwarp();
          accumulate = 1;
        }
        if (elect_one()) umma_commit(&tmem_full[acc]);
        __syncwarp();
        TL_ITEM(8, local);
      }
      TL(4);
    }
  } else if (warp < 6) {
    // ===== operand splitter (kSplit == 3) =====
    // The tensor core truncates fp32 operands to tf32, so the landed fp32 tile IS the hi operand; only lo = x - trunc(x)
    // has to be produced (same swizzled offsets as the source tile, so no layout knowledge is needed for B / MN-major A).
    if (kSplit == 3) {
      const int t = threadIdx.x - 64;                               // 0..127
      const int row = (warp & 3) * 32 + lane;                       // TMEM lane owned by this thread
      int g = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        const WorkItem it = decode(w);
        for (int i = 0; i < it.n_kb; ++i, ++g) {
          const int s = g % C::kStages, round = g / C::kStages;
          mbar_wait(&full[s], round & 1);
          if (g == 0) TL(3);
          uint8_t* stg = smem + s * C::kStageBytes;
          if (A_TMEM) {
            uint32_t hi[32], lo[32];
            if (A_KMAJOR) {
              // row `row` of the K-major A tile: 8 x 16-byte chunks, chunk c stored at c ^ (row & 7) (SWIZZLE_128B)
              const uint8_t* arow = stg + OFF_A + row * 128;
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                const float4 v = *reinterpret_cast<const float4*>(arow + ((c ^ (row & 7)) << 4));
                hi[4 * c + 0] = __float_as_uint(v.x); lo[4 * c + 0] = __float_as_uint(tf32_lo(v.x));
                hi[4 * c + 1] = __float_as_uint(v.y); lo[4 * c + 1] = __float_as_uint(tf32_lo(v.y));
                hi[4 * c + 2] = __float_as_uint(v.z); lo[4 * c + 2] = __float_as_uint(tf32_lo(v.z));
                hi[4 * c + 3] = __float_as_uint(v.w); lo[4 * c + 3] = __float_as_uint(tf32_lo(v.w));
              }
            } else {
              // MN-major A tile [chunk = row/32][k][32 floats], 32-byte segments XOR-ed with (k & 3) (SWIZZLE_128B_ATOM_32B):
              // this thread gathers column `row` over the 32 k-rows (a transpose; each k is one conflict-free warp read)
              const uint8_t* acol = stg + OFF_A + (warp & 3) * 4096 + (lane & 7) * 4;
#pragma unroll
              for (int k = 0; k < 32; ++k) {
                const float v = *reinterpret_cast<const float*>(acol + k * 128 + (((lane >> 3) ^ (k & 3)) << 5));
                hi[k] = __float_as_uint(v);
                lo[k] = __float_as_uint(tf32_lo(v));
              }
            }
            const uint32_t ta = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + C::kTmemA + s * 64;
            tmem_st_32x32(ta, hi);
            tmem_st_32x32(ta + 32, lo);
            const float4* bsrc = reinterpret_cast<const float4*>(stg + OFF_B);
            float4* blo = reinterpret_cast<float4*>(stg + OFF_BLO);
#pragma unroll
            for (int j = t; j < TILE_B / 16; j += 128) {
              const float4 v = bsrc[j];
              blo[j] = make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
            }
            tmem_st_wait();
            tc_fence_before();
          } else {
            const float4* asrc = reinterpret_cast<const float4*>(stg + OFF_A);
            float4* alo = reinterpret_cast<float4*>(stg + OFF_ALO);
#pragma unroll 4
            for (int j = t; j < TILE_BYTES / 16; j += 128) {
              const float4 v = asrc[j];
              alo[j] = make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
            }
            const float4* bsrc = reinterpret_cast<const float4*>(stg + OFF_B);
            float4* blo = reinterpret_cast<float4*>(stg + OFF_BLO);
#pragma unroll 4
            for (int j = t; j < TILE_B / 16; j += 128) {
              const float4 v = bsrc[j];
              blo[j] = make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
            }
          }
          fence_proxy_async();
          mbar_arrive(&ready[s]);
          if (g == 0) TL(4);
        }
      }
      TL(5);
    }
  } else {
    // ===== epilogue warps (6..): drain accumulator `acc` of work item i while the other roles run item i+1 =====
    // TMEM lane quadrant = warp % 4.  With 8 epilogue warps two warps share a quadrant and take alternate 32-column chunks:
    // twice the loads of the axpy epilogue's `src` operand in flight (that stream, not the MMAs, bounds the weight-gradient
    // GEMMs: 16 KB in flight per SM against ~0.6 us of L2 latency is 4 TB/s over the chip).
    constexpr int kGroups = kEpiWarps / 4;
    const int grp = (warp - 6) >> 2;
    int local = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++local) {
      const WorkItem it = decode(w);
      const bool last = w + (int)gridDim.x >= total_work;
      if (kGroups == 1) {
        // the CTA's last item has nothing to overlap with: the (by then idle) splitter warps take the upper two chunks
        drain(it, local & 1, (local >> 1) & 1, 0, (last && kSplit == 3) ? kHalfN : BN, 32, epi_stage + (warp - 6) * kPatchFloats, !last);
      } else {
        drain(it, local & 1, (local >> 1) & 1, 32 * grp, BN, 32 * kGroups, epi_stage + (warp - 6) * kPatchFloats, true);
      }
      if (local == 0) TL(6);
    }
    TL(7);
  }
  if (kEpiWarps == 4 && kSplit == 3 && warp >= 2 && warp < 6 && total_work > (int)blockIdx.x) {
    const int n_items = (total_work - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int local = n_items - 1;
    const WorkItem it = decode((int)blockIdx.x + local * (int)gridDim.x);
    // pipeline stage 0 is free once tmem_full of the last item has fired (drain waits for it before touching it)
    drain(it, local & 1, (local >> 1) & 1, kHalfN, BN, 32, reinterpret_cast<float*>(smem) + (warp - 2) * kPatchFloats, false);
  }
  if constexpr (Epi::kKind == kEpiStoreTma) {
    // the patches must outlive the stores' READS of them; the global writes themselves are ordinary in-flight stores at exit and
    // are performed before the grid counts as complete (what a dependent's griddepcontrol.wait / the stream order waits for)
    // (waiting for full completion instead measured 1.6038 vs 1.5973 ms per iteration)
    if (warp >= 2 && elect_one()) bulk_wait_read0();
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, C::kTmemCols);
  TL(8);
}

}  // namespace tc
}  // namespace vldd
