// Register-resident float4 versions of the four LayerNorm/normalise row kernels (one CTA of 256 threads per row,
// d % 4 == 0, d <= 1024 * NV).  Same math as the generic kernels in head_kernels.cuh (which remain the fallback for
// other d); each row is read from global memory once and kept in registers across the dependent reductions.
#pragma once
#include "head_kernels.cuh"

namespace vldd {

__device__ __forceinline__ float hsum(float4 a) { return (a.x + a.y) + (a.z + a.w); }
__device__ __forceinline__ float hdot(float4 a, float4 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w))); }

// two simultaneous block sums (one barrier round trip instead of two)
__device__ __forceinline__ float2 block_sum2(float a, float b, float2* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  a = warp_sum(a);
  b = warp_sum(b);
  __syncthreads();
  if (lane == 0) scratch[wid] = make_float2(a, b);
  __syncthreads();
  float2 t = lane < nw ? scratch[lane] : make_float2(0.f, 0.f);
  t.x = warp_sum(t.x);
  t.y = warp_sum(t.y);
  return t;
}
__device__ __forceinline__ float block_sum1(float a, float2* scratch) { return block_sum2(a, 0.f, scratch).x; }

// These kernels are latency-bound (100 rows x 9 KB, L2-resident): what they cost is the number of dependent L2 round
// trips.  Every operand a later phase needs is therefore loaded up front, next to the slab loads (predicated 128-bit
// loads, all in flight together; ~150 registers per thread at NV = 3, one CTA per SM anyway), so that after each block
// reduction only arithmetic and stores remain.
#define VLDD_LDP(ok, ptr) ((ok) ? ld4(ptr) : f4(0.f))

// acc[i] = sum over slabs of element (base + 4 * (threadIdx.x + 256 i)), all NV x splits loads issued before the first add
template <int NV>
__device__ __forceinline__ void load_slab_rows(float4 (&acc)[NV], const float* __restrict__ part, int splits, size_t stride,
                                               size_t base, int d4) {
  if (splits <= kMaxSlabsMlp) {
    float4 t[NV][kMaxSlabsMlp];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int j = threadIdx.x + 256 * i;
#pragma unroll
      for (int z = 0; z < kMaxSlabsMlp; ++z) t[i][z] = VLDD_LDP(j < d4 && z < splits, part + (size_t)z * stride + base + 4 * j);
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      acc[i] = t[i][0];
#pragma unroll
      for (int z = 1; z < kMaxSlabsMlp; ++z) acc[i] = acc[i] + t[i][z];
    }
  } else {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int j = threadIdx.x + 256 * i;
      acc[i] = j < d4 ? sum_slabs4(part, splits, stride, base + 4 * j) : f4(0.f);
    }
  }
}

template <int NV>
__global__ void __launch_bounds__(256) ln_fwd_v4_kernel(const float* __restrict__ part, int splits, size_t stride,
                                                        const float* __restrict__ b2, const float* __restrict__ mask,
                                                        const float* __restrict__ p, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, int d, float* __restrict__ rhat,
                                                        float* __restrict__ z_out, float* __restrict__ yn,
                                                        float* __restrict__ rstd_out, float* __restrict__ nz_out) {
  pdl_enter();
  __shared__ float2 scratch[32];
  const int row = blockIdx.x, d4 = d >> 2;
  const size_t base = (size_t)row * d;
  float4 r[NV], bb[NV], mk[NV], pp[NV], gm[NV], bt[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = threadIdx.x + 256 * i;
    const bool ok = j < d4;
    bb[i] = VLDD_LDP(ok, b2 + 4 * j);
    mk[i] = VLDD_LDP(ok && mask != nullptr, mask + base + 4 * j);
    pp[i] = VLDD_LDP(ok, p + base + 4 * j);
    gm[i] = VLDD_LDP(ok, gamma + 4 * j);
    bt[i] = VLDD_LDP(ok, beta + 4 * j);
  }
  load_slab_rows<NV>(r, part, splits, stride, base, d4);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = threadIdx.x + 256 * i;
    if (j < d4) {
      float4 f = r[i] + bb[i];
      if (mask) f = f * mk[i];
      r[i] = f + pp[i];
      s += hsum(r[i]);
    } else {
      r[i] = f4(0.f);
    }
  }
  const float mu = block_sum1(s, scratch) / d;
  float v = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = threadIdx.x + 256 * i;
    if (j < d4) { r[i] = r[i] - f4(mu); v += hdot(r[i], r[i]); }
  }
  const float rstd = rsqrtf(block_sum1(v, scratch) / d + kLnEps);
  float zz = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = threadIdx.x + 256 * i;
    if (j < d4) {
      const float4 rh = rstd * r[i];
      if (rhat) st4(rhat + base + 4 * j, rh);
      r[i] = gm[i] * rh + bt[i];
      zz += hdot(r[i], r[i]);
    }
  }
  const float nz = sqrtf(block_sum1(zz, scratch));
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = threadIdx.x + 256 * i;
    if (j < d4) {
      if (z_out) st4(z_out + base + 4 * j, r[i]);
      if (yn) st4(yn + base + 4 * j, make_float4(r[i].x / nz, r[i].y / nz, r[i].z / nz, r[i].w / nz));
    }
  }
  if (threadIdx.x == 0) {
    if (rstd_out) rstd_out[row] = rstd;
    if (nz_out) nz_out[row] = nz;
  }
}

template <int NV>
__global__ void __launch_bounds__(256) norm_ln_bwd_v4_kernel(const float* __restrict__ raw, const float* __restrict__ scale,
                                                             const float* __restrict__ yn, const float* __restrict__ nz_p,
                                                             const float* __restrict__ rhat, const float* __restrict__ rstd_p,
                                                             const float* __restrict__ gamma, const float* __restrict__ mask,
                                                             int d, float* __restrict__ dyn, float* __restrict__ q_out,
                                                             float* __restrict__ dz, float* __restrict__ dr,
                                                             float* __restrict__ df) {
  pdl_enter();
  __shared__ float2 scratch[32];
  const int row = blockIdx.x, d4 = d >> 2;
  const size_t base = (size_t)row * d;
  float4 a[NV], y[NV], rh[NV], gm[NV], mk[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = threadIdx.x + 256 * i;
    const bool ok = j < d4;
    a[i] = VLDD_LDP(ok, raw + base + 4 * j);
    y[i] = VLDD_LDP(ok, yn + base + 4 * j);
    rh[i] = VLDD_LDP(ok, rhat + base + 4 * j);
    gm[i] = VLDD_LDP(ok, gamma + 4 * j);
    mk[i] = VLDD_LDP(ok && mask != nullptr, mask + base + 4 * j);
  }
  const float sc = *scale, inz = 1.0f / nz_p[row], rstd = rstd_p[row];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = threadIdx.x + 256 * i;
    if (j < d4) {
      a[i] = sc * a[i];
      st4(dyn + base + 4 * j, a[i]);
      s += hdot(y[i], a[i]);
    }
  }
  const float q = block_sum1(s, scratch);
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = threadIdx.x + 256 * i;
    if (j < d4) {
      const float4 dzv = inz * (a[i] - q * y[i]);
      st4(dz + base + 4 * j, dzv);
      a[i] = gm[i] * dzv;     // drhat
      s1 += hsum(a[i]);
      s2 += hdot(a[i], rh[i]);
    }
  }
  const float2 mm = block_sum2(s1, s2, scratch);
  const float m1 = mm.x / d, m2 = mm.y / d;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = threadIdx.x + 256 * i;
    if (j < d4) {
      const float4 v = rstd * (a[i] - f4(m1) - m2 * rh[i]);
      st4(dr + base + 4 * j, v);
      st4(df + base + 4 * j, mask ? v * mk[i] : v);
    }
  }
  if (threadIdx.x == 0) q_out[row] = q;
}

template <int NV>
__global__ void __launch_bounds__(256) ln_tangent_v4_kernel(const float* __restrict__ part, int splits, size_t stride,
                                                            const float* __restrict__ c2, const float* __restrict__ mask,
                                                            const float* __restrict__ pd, const float* __restrict__ rhat,
                                                            const float* __restrict__ rstd_p, const float* __restrict__ yn,
                                                            const float* __restrict__ nz_p, const float* __restrict__ gamma,
                                                            const float* __restrict__ gammad, const float* __restrict__ betad,
                                                            int d, float* __restrict__ rhatd, float* __restrict__ ynd,
                                                            float* __restrict__ t_out, float* __restrict__ nzd_out) {
  pdl_enter();
  __shared__ float2 scratch[32];
  const int row = blockIdx.x, d4 = d >> 2;
  const size_t base = (size_t)row * d;
  float4 rd[NV], rh[NV], y[NV], cc[NV], mk[NV], pp[NV], gm[NV], gd[NV], bd[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = threadIdx.x + 256 * i;
    const bool ok = j < d4;
    cc[i] = VLDD_LDP(ok, c2 + 4 * j);
    mk[i] = VLDD_LDP(ok && mask != nullptr, mask + base + 4 * j);
    pp[i] = VLDD_LDP(ok, pd + base + 4 * j);
    rh[i] = VLDD_LDP(ok, rhat + base + 4 * j);
    y[i] = VLDD_LDP(ok, yn + base + 4 * j);
    gm[i] = VLDD_LDP(ok, gamma + 4 * j);
    gd[i] = VLDD_LDP(ok, gammad + 4 * j);
    bd[i] = VLDD_LDP(ok, betad + 4 * j);
  }
  load_slab_rows<NV>(rd, part, splits, stride, base, d4);
  const float rstd = rstd_p[row], nz = nz_p[row];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = threadIdx.x + 256 * i;
    if (j < d4) {
      float4 f = rd[i] + cc[i];
      if (mask) f = f * mk[i];
      rd[i] = f + pp[i];
      s1 += hsum(rd[i]);
      s2 += hdot(rh[i], rd[i]);
    }
  }
  const float2 mm = block_sum2(s1, s2, scratch);
  const float mrd = mm.x / d, t = mm.y / d;
  float s3 = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = threadIdx.x + 256 * i;
    if (j < d4) {
      const float4 rhd = rstd * (rd[i] - f4(mrd) - t * rh[i]);
      st4(rhatd + base + 4 * j, rhd);
      rd[i] = gd[i] * rh[i] + gm[i] * rhd + bd[i];   // zd
      s3 += hdot(y[i], rd[i]);
    }
  }
  const float nzd = block_sum1(s3, scratch);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = threadIdx.x + 256 * i;
    if (j < d4) {
      const float4 o = rd[i] - nzd * y[i];
      st4(ynd + base + 4 * j, make_float4(o.x / nz, o.y / nz, o.z / nz, o.w / nz));
    }
  }
  if (threadIdx.x == 0) { t_out[row] = t; nzd_out[row] = nzd; }
}

template <int NV>
__global__ void __launch_bounds__(256) norm_ln_bwd_tangent_v4_kernel(
    const float* __restrict__ raw, const float* __restrict__ scale, const float* __restrict__ yn,
    const float* __restrict__ ynd, const float* __restrict__ dyn, const float* __restrict__ q_p,
    const float* __restrict__ nz_p, const float* __restrict__ nzd_p, const float* __restrict__ dz,
    const float* __restrict__ rhat, const float* __restrict__ rhatd, const float* __restrict__ rstd_p,
    const float* __restrict__ t_p, const float* __restrict__ dr, const float* __restrict__ gamma,
    const float* __restrict__ gammad, const float* __restrict__ mask, int d, float* __restrict__ dzd,
    float* __restrict__ drd, float* __restrict__ dfd) {
  pdl_enter();
  __shared__ float2 scratch[32];
  const int row = blockIdx.x, d4 = d >> 2;
  const size_t base = (size_t)row * d;
  float4 a[NV], y[NV], yd[NV], dy[NV], dzv[NV], gm[NV], gd[NV], rh[NV], rhd[NV], drv[NV], mk[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = threadIdx.x + 256 * i;
    const bool ok = j < d4;
    a[i] = VLDD_LDP(ok, raw + base + 4 * j);
    y[i] = VLDD_LDP(ok, yn + base + 4 * j);
    yd[i] = VLDD_LDP(ok, ynd + base + 4 * j);
    dy[i] = VLDD_LDP(ok, dyn + base + 4 * j);
    dzv[i] = VLDD_LDP(ok, dz + base + 4 * j);
    gm[i] = VLDD_LDP(ok, gamma + 4 * j);
    gd[i] = VLDD_LDP(ok, gammad + 4 * j);
    rh[i] = VLDD_LDP(ok, rhat + base + 4 * j);
    rhd[i] = VLDD_LDP(ok, rhatd + base + 4 * j);
    drv[i] = VLDD_LDP(ok, dr + base + 4 * j);
    mk[i] = VLDD_LDP(ok && mask != nullptr, mask + base + 4 * j);
  }
  const float sc = *scale, nz = nz_p[row], nzd = nzd_p[row], q = q_p[row], rstd = rstd_p[row], t = t_p[row];
  const float inz = 1.0f / nz;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = threadIdx.x + 256 * i;
    if (j < d4) {
      a[i] = sc * a[i];           // dynd
      s += hdot(yd[i], dy[i]) + hdot(y[i], a[i]);
    }
  }
  const float qd = block_sum1(s, scratch);
  float s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = threadIdx.x + 256 * i;
    if (j < d4) {
      const float4 v = inz * (a[i] - q * yd[i] - qd * y[i]) - (nzd * inz) * dzv[i];
      st4(dzd + base + 4 * j, v);
      const float4 drh = gm[i] * dzv[i];
      a[i] = gd[i] * dzv[i] + gm[i] * v;      // drhatd
      s1 += hsum(a[i]);
      s2 += hdot(a[i], rh[i]) + hdot(drh, rhd[i]);
      s3 += hdot(drh, rh[i]);
    }
  }
  const float2 mm = block_sum2(s1, s2, scratch);
  const float m1d = mm.x / d, m2d = mm.y / d;
  const float m2 = block_sum1(s3, scratch) / d;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = threadIdx.x + 256 * i;
    if (j < d4) {
      const float4 v = (-(rstd * t)) * drv[i] + rstd * (a[i] - f4(m1d) - m2 * rhd[i] - m2d * rh[i]);
      st4(drd + base + 4 * j, v);
      st4(dfd + base + 4 * j, mask ? v * mk[i] : v);
    }
  }
}

inline bool row_v4_ok(int d) { return d % 4 == 0 && d <= 3072; }
#define VLDD_UNPACK(...) __VA_ARGS__
#define VLDD_ROW_V4_DISPATCH(d, KERNEL, CFG, ARGS)                                         \
  do {                                                                                     \
    if ((d) <= 1024) launch_k(KERNEL<1>, VLDD_UNPACK CFG, VLDD_UNPACK ARGS);               \
    else if ((d) <= 2048) launch_k(KERNEL<2>, VLDD_UNPACK CFG, VLDD_UNPACK ARGS);          \
    else launch_k(KERNEL<3>, VLDD_UNPACK CFG, VLDD_UNPACK ARGS);                           \
  } while (0)

}  // namespace vldd
