// fp32 CUDA-core GEMM used as the exact-fp32 contraction primitive of the engine.
//
//   C[m,n] = sum_k opA(A)[m,k] * opB(B)[k,n]          (+ optional second K segment: A1/B1)
//
// * 128x128x16 CTA tile, 256 threads, 8x8 register tile per thread (two 4-wide halves per axis so
//   shared-memory reads are float4 and conflict-free on the B side, broadcast on the A side).
// * Operand contiguity is a template parameter: K-contiguous ("row-major [rows,K]") tiles are loaded
//   with float4 along k and transposed into smem; MN-contiguous ("[K,rows]") tiles are copied directly.
// * Two K segments let the tangent GEMMs  dH W2^T + H V2^T,  dF^T H + ...  run as ONE launch.
// * split-K across blockIdx.z writes partial slabs (deterministic: the consumer kernel sums the slabs
//   in fixed order); splits == 1 runs the fused epilogue functor instead.
#pragma once
#include "common.cuh"

namespace vldd {

struct GemmOperands {
  const float* A0; const float* B0; int lda0; int ldb0; int K0;
  const float* A1; const float* B1; int lda1; int ldb1; int K1;   // K1 == 0 -> unused
  int M; int N;
};

constexpr int GBM = 128, GBN = 128, GBK = 16, GPAD = 4;

// Fetch this thread's share (2 x float4) of a 128(rows) x 16(k) tile into registers, zero-filled out of range.
// KMAJOR: source element (row, k) at src[row*ld + k];  else at src[k*ld + row].
template <bool KMAJOR>
__device__ __forceinline__ void fetch_tile(float4 (&reg)[2], const float* __restrict__ src, int ld, int row0, int nrows,
                                           int k0, int kend, bool vec_ok) {
  const int t = threadIdx.x;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    if (KMAJOR) {
      const int gr = row0 + (t >> 2) + 64 * r, gk = k0 + (t & 3) * 4;
      if (gr < nrows) {
        const float* p = src + (size_t)gr * ld + gk;
        if (vec_ok && gk + 3 < kend) {
          q = *reinterpret_cast<const float4*>(p);
        } else {
          if (gk + 0 < kend) q.x = p[0];
          if (gk + 1 < kend) q.y = p[1];
          if (gk + 2 < kend) q.z = p[2];
          if (gk + 3 < kend) q.w = p[3];
        }
      }
    } else {
      const int gk = k0 + (t >> 5) + 8 * r, gr = row0 + (t & 31) * 4;
      if (gk < kend) {
        const float* p = src + (size_t)gk * ld + gr;
        if (vec_ok && gr + 3 < nrows) {
          q = *reinterpret_cast<const float4*>(p);
        } else {
          if (gr + 0 < nrows) q.x = p[0];
          if (gr + 1 < nrows) q.y = p[1];
          if (gr + 2 < nrows) q.z = p[2];
          if (gr + 3 < nrows) q.w = p[3];
        }
      }
    }
    reg[r] = q;
  }
}
// Commit the fetched registers to smem laid out [k][row].
template <bool KMAJOR>
__device__ __forceinline__ void commit_tile(float (*dst)[GBM + GPAD], const float4 (&reg)[2]) {
  const int t = threadIdx.x;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    if (KMAJOR) {
      const int row = (t >> 2) + 64 * r, kq = (t & 3) * 4;
      dst[kq + 0][row] = reg[r].x;
      dst[kq + 1][row] = reg[r].y;
      dst[kq + 2][row] = reg[r].z;
      dst[kq + 3][row] = reg[r].w;
    } else {
      const int k = (t >> 5) + 8 * r, rq = (t & 31) * 4;
      *reinterpret_cast<float4*>(&dst[k][rq]) = reg[r];
    }
  }
}

// Epilogue functors -----------------------------------------------------------------------------
struct EpiStore {          // C = alpha * acc
  float* C; int ldc; float alpha;
  __device__ __forceinline__ void operator()(int m, int n, float v) const { C[(size_t)m * ldc + n] = alpha * v; }
};
struct EpiAxpy {           // dst = src - (*lr) * acc     (theta_{k+1} = theta_k - lr dW ; a_k = a_{k+1} - lr H a); src nullable = 0
  const float* src; float* dst; int ld; const float* lr;
  __device__ __forceinline__ void operator()(int m, int n, float v) const {
    const size_t i = (size_t)m * ld + n;
    dst[i] = (src ? src[i] : 0.f) - (*lr) * v;
  }
};

template <bool A_KMAJOR, bool B_KMAJOR, class Epi>
__global__ void __launch_bounds__(256) gemm_simt_kernel(GemmOperands g, float* __restrict__ partials,
                                                        long long part_stride, Epi epi, int vecA0, int vecB0,
                                                        int vecA1, int vecB1) {
  pdl_enter();
  __shared__ __align__(16) float As[2][GBK][GBM + GPAD];
  __shared__ __align__(16) float Bs[2][GBK][GBN + GPAD];
  const int m0 = blockIdx.y * GBM, n0 = blockIdx.x * GBN;
  const int nkb0 = (g.K0 + GBK - 1) / GBK, nkb1 = (g.K1 + GBK - 1) / GBK, nkb = nkb0 + nkb1;
  const int splits = gridDim.z;
  const int per = (nkb + splits - 1) / splits;
  const int kb_begin = blockIdx.z * per, kb_end = min(nkb, kb_begin + per);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 ra[2], rb[2];
  auto fetch = [&](int kb) {
    if (kb < nkb0) {
      fetch_tile<A_KMAJOR>(ra, g.A0, g.lda0, m0, g.M, kb * GBK, g.K0, vecA0);
      fetch_tile<B_KMAJOR>(rb, g.B0, g.ldb0, n0, g.N, kb * GBK, g.K0, vecB0);
    } else {
      fetch_tile<A_KMAJOR>(ra, g.A1, g.lda1, m0, g.M, (kb - nkb0) * GBK, g.K1, vecA1);
      fetch_tile<B_KMAJOR>(rb, g.B1, g.ldb1, n0, g.N, (kb - nkb0) * GBK, g.K1, vecB1);
    }
  };
  if (kb_begin < kb_end) {
    fetch(kb_begin);
    commit_tile<A_KMAJOR>(As[0], ra);
    commit_tile<B_KMAJOR>(Bs[0], rb);
  }
  __syncthreads();
  for (int kb = kb_begin; kb < kb_end; ++kb) {
    const int buf = (kb - kb_begin) & 1;
    const bool more = kb + 1 < kb_end;
    if (more) fetch(kb + 1);          // global loads in flight while this k-block is multiplied
#pragma unroll
    for (int k = 0; k < GBK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) {
      commit_tile<A_KMAJOR>(As[buf ^ 1], ra);
      commit_tile<B_KMAJOR>(Bs[buf ^ 1], rb);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= g.N) continue;
      if (partials != nullptr)
        partials[(size_t)blockIdx.z * part_stride + (size_t)m * g.N + n] = acc[i][j];
      else
        epi(m, n, acc[i][j]);
    }
  }
}

inline int vec_ok(const float* p, int ld) { return p != nullptr && aligned16(p) && (ld % 4 == 0); }

// partials != nullptr -> partial slabs [splits][M*N] written there (epilogue ignored)
// partials == nullptr -> splits must be 1; epi(m, n, acc) is applied
template <bool A_KMAJOR, bool B_KMAJOR, class Epi>
inline void launch_gemm(const GemmOperands& g, int splits, float* partials, Epi epi, cudaStream_t st) {
  dim3 grid(ceil_div(g.N, GBN), ceil_div(g.M, GBM), splits);
  launch_k(gemm_simt_kernel<A_KMAJOR, B_KMAJOR, Epi>, grid, 256, 0, st, 
      g, partials, (long long)g.M * g.N, epi, vec_ok(g.A0, g.lda0), vec_ok(g.B0, g.ldb0),
      g.K1 ? vec_ok(g.A1, g.lda1) : 0, g.K1 ? vec_ok(g.B1, g.ldb1) : 0);
}

inline GemmOperands gemm_ops(const float* A, int lda, const float* B, int ldb, int M, int N, int K) {
  GemmOperands g{};
  g.A0 = A; g.B0 = B; g.lda0 = lda; g.ldb0 = ldb; g.K0 = K;
  g.A1 = nullptr; g.B1 = nullptr; g.lda1 = 0; g.ldb1 = 0; g.K1 = 0;
  g.M = M; g.N = N;
  return g;
}
inline GemmOperands gemm_ops2(const float* A0, int lda0, const float* B0, int ldb0, int K0, const float* A1, int lda1,
                              const float* B1, int ldb1, int K1, int M, int N) {
  GemmOperands g = gemm_ops(A0, lda0, B0, ldb0, M, N, K0);
  g.A1 = A1; g.B1 = B1; g.lda1 = lda1; g.ldb1 = ldb1; g.K1 = K1;
  return g;
}

}  // namespace vldd
