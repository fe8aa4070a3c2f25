// HBM-streaming kernels of the distill path: flat-parameter SGD step, matching-loss reduction and its
// backward, outer momentum SGD.  All are 128-bit vectorised, grid-stride, sized to a multiple of the SM count.
//
//   reference sites:  distill.py:582-583 (theta - lr*g), 588-598 (mse_loss sums, ratio), 233-241 + 611-613
//   (torch.optim.SGD momentum=0.5 on image_syn / text_syn / syn_lr).
#include "common.cuh"
#include "kernels.h"

namespace vldd {

constexpr int kStreamThreads = 256;
constexpr int kStreamBlocksPerSM = 8;

static inline int stream_grid(int64_t n_vec) {
  int64_t want = ceil_div64(n_vec, kStreamThreads);
  int64_t cap = (int64_t)num_sms() * kStreamBlocksPerSM;
  return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

// theta_next = theta - lr * grad                                      (12 B / param)
__global__ void __launch_bounds__(kStreamThreads) flat_sgd_step_kernel(const float* __restrict__ theta,
                                                                       const float* __restrict__ grad,
                                                                       const float* __restrict__ lr_p,
                                                                       float* __restrict__ out, int64_t n, int vec) {
  pdl_enter();
  const float lr = *lr_p;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  if (vec) {
    const int64_t n4 = n >> 2;
    for (int64_t i = tid; i < n4; i += stride) {
      const float4 t = ldg_stream4(theta + 4 * i), g = ldg_stream4(grad + 4 * i);
      stg_stream4(out + 4 * i, make_float4(t.x - lr * g.x, t.y - lr * g.y, t.z - lr * g.z, t.w - lr * g.w));
    }
    for (int64_t i = (n4 << 2) + tid; i < n; i += stride) out[i] = theta[i] - lr * grad[i];
  } else {
    for (int64_t i = tid; i < n; i += stride) out[i] = theta[i] - lr * grad[i];
  }
}

// num = sum (theta_K - theta*)^2, den = sum (theta_0 - theta*)^2       (12 B / param)
// Deterministic: per-thread fp32 partial over a fixed index set, block sums in fp64 written to
// block_partials[2*gridDim.x]; the last block to finish (ticket) adds them in index order.
__global__ void __launch_bounds__(kStreamThreads) match_loss_fwd_kernel(const float* __restrict__ thK,
                                                                        const float* __restrict__ tgt,
                                                                        const float* __restrict__ th0, int64_t n,
                                                                        int vec, double* __restrict__ block_partials,
                                                                        unsigned int* __restrict__ ticket,
                                                                        float* __restrict__ out3) {
  pdl_enter();
  __shared__ double scratch[34];
  __shared__ bool is_last;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  float num = 0.f, den = 0.f;
  if (vec) {
    const int64_t n4 = n >> 2;
    for (int64_t i = tid; i < n4; i += stride) {
      const float4 a = ldg_stream4(thK + 4 * i), t = ldg_stream4(tgt + 4 * i), b = ldg_stream4(th0 + 4 * i);
      float d;
      d = a.x - t.x; num = fmaf(d, d, num); d = b.x - t.x; den = fmaf(d, d, den);
      d = a.y - t.y; num = fmaf(d, d, num); d = b.y - t.y; den = fmaf(d, d, den);
      d = a.z - t.z; num = fmaf(d, d, num); d = b.z - t.z; den = fmaf(d, d, den);
      d = a.w - t.w; num = fmaf(d, d, num); d = b.w - t.w; den = fmaf(d, d, den);
    }
    for (int64_t i = (n4 << 2) + tid; i < n; i += stride) {
      float d = thK[i] - tgt[i]; num = fmaf(d, d, num);
      d = th0[i] - tgt[i]; den = fmaf(d, d, den);
    }
  } else {
    for (int64_t i = tid; i < n; i += stride) {
      float d = thK[i] - tgt[i]; num = fmaf(d, d, num);
      d = th0[i] - tgt[i]; den = fmaf(d, d, den);
    }
  }
  const double bn = block_sum<double>((double)num, scratch);
  const double bd = block_sum<double>((double)den, scratch);
  if (threadIdx.x == 0) {
    block_partials[2 * blockIdx.x] = bn;
    block_partials[2 * blockIdx.x + 1] = bd;
    __threadfence();
    const unsigned int t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    double sn = 0.0, sd = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
      sn += block_partials[2 * b];
      sd += block_partials[2 * b + 1];
    }
    sn = block_sum<double>(sn, scratch);
    sd = block_sum<double>(sd, scratch);
    if (threadIdx.x == 0) {
      out3[0] = (float)sn;
      out3[1] = (float)sd;
      out3[2] = (float)(sn / sd);
      *ticket = 0u;  // re-arm for the next launch (stream-ordered)
    }
  }
}

// a = gout * 2 (theta_K - theta*) / den                                 (12 B / param: 2 reads + 1 write)
__global__ void __launch_bounds__(kStreamThreads) match_loss_bwd_kernel(const float* __restrict__ thK,
                                                                        const float* __restrict__ tgt,
                                                                        const float* __restrict__ num_den,
                                                                        const float* __restrict__ gout,
                                                                        float* __restrict__ a, int64_t n, int vec) {
  pdl_enter();
  const float c = (gout ? *gout : 1.0f) * 2.0f / num_den[1];
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  if (vec) {
    const int64_t n4 = n >> 2;
    for (int64_t i = tid; i < n4; i += stride) {
      const float4 x = ldg_stream4(thK + 4 * i), t = ldg_stream4(tgt + 4 * i);
      stg_stream4(a + 4 * i, make_float4(c * (x.x - t.x), c * (x.y - t.y), c * (x.z - t.z), c * (x.w - t.w)));
    }
    for (int64_t i = (n4 << 2) + tid; i < n; i += stride) a[i] = c * (thK[i] - tgt[i]);
  } else {
    for (int64_t i = tid; i < n; i += stride) a[i] = c * (thK[i] - tgt[i]);
  }
}

// buf = first ? g : momentum*buf + g ;  p -= lr * buf                   (16 B / element: 3 reads + 2 writes = 20; in place)
__global__ void __launch_bounds__(kStreamThreads) momentum_sgd_kernel(float* __restrict__ p,
                                                                      const float* __restrict__ g,
                                                                      float* __restrict__ buf, float lr, float momentum,
                                                                      int first, int64_t n, int vec) {
  pdl_enter();
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  if (vec) {
    const int64_t n4 = n >> 2;
    for (int64_t i = tid; i < n4; i += stride) {
      const float4 gg = ldg_stream4(g + 4 * i);
      float4 b = gg;
      if (!first) {
        const float4 ob = *reinterpret_cast<const float4*>(buf + 4 * i);
        b = make_float4(fmaf(momentum, ob.x, gg.x), fmaf(momentum, ob.y, gg.y), fmaf(momentum, ob.z, gg.z),
                        fmaf(momentum, ob.w, gg.w));
      }
      float4 pp = *reinterpret_cast<const float4*>(p + 4 * i);
      pp.x -= lr * b.x; pp.y -= lr * b.y; pp.z -= lr * b.z; pp.w -= lr * b.w;
      *reinterpret_cast<float4*>(buf + 4 * i) = b;
      *reinterpret_cast<float4*>(p + 4 * i) = pp;
    }
    for (int64_t i = (n4 << 2) + tid; i < n; i += stride) {
      const float b = first ? g[i] : fmaf(momentum, buf[i], g[i]);
      buf[i] = b;
      p[i] -= lr * b;
    }
  } else {
    for (int64_t i = tid; i < n; i += stride) {
      const float b = first ? g[i] : fmaf(momentum, buf[i], g[i]);
      buf[i] = b;
      p[i] -= lr * b;
    }
  }
}


// ---- fused matching-loss passes of the unroll engine ---------------------------------------------------
// The reference computes  loss = |theta_K - theta*|^2 / |theta_0 - theta*|^2  and backpropagates it (distill.py:588-598, 606).
// The denominator does not depend on the student, so it is accumulated while the segment is staged into the engine's
// workspace (the copy has to happen anyway: the CUDA graph needs fixed addresses), and ONE pass at the end of the unroll
// produces the numerator and the adjoint a_K = 2 (theta_K - theta*) / den together: 3 vector passes on the critical path
// (read theta_K, theta*, write a_K: 85 MB at the Flickr shape) instead of the 6 of match_loss_fwd + match_loss_bwd (170 MB).
struct BlockReduceOut { double* parts; unsigned int* ticket; };

// ticketed deterministic finish: block sums in fp64, the last block adds them in index order and calls `fin(sum)`
template <class Fin>
__device__ __forceinline__ void finish_sum(float v, double* __restrict__ parts, unsigned int* __restrict__ ticket, Fin fin) {
  __shared__ double scratch[34];
  __shared__ bool is_last;
  const double b = block_sum<double>((double)v, scratch);
  if (threadIdx.x == 0) {
    parts[blockIdx.x] = b;
    __threadfence();
    is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    double s = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) s += parts[i];
    s = block_sum<double>(s, scratch);
    if (threadIdx.x == 0) {
      fin(s);
      *ticket = 0u;   // re-arm (stream-ordered)
    }
  }
}

// th0_dst = th0_src ; tgt_dst = tgt_src ; den = sum (th0 - tgt)^2            (16 B / param: 2 reads + 2 writes)
__device__ __forceinline__ void stage_segment_body(const float* __restrict__ th0_src, const float* __restrict__ tgt_src,
                                                   float* __restrict__ th0_dst, float* __restrict__ tgt_dst, int64_t n, int vec,
                                                   double* __restrict__ parts, unsigned int* __restrict__ ticket,
                                                   float* __restrict__ den_out) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  float den = 0.f;
  if (vec) {
    const int64_t n4 = n >> 2;
    int64_t i = tid;
    for (; i + stride < n4; i += 2 * stride) {       // two independent 128-bit loads per operand in flight
      const float4 a0 = ldg_stream4(th0_src + 4 * i), t0 = ldg_stream4(tgt_src + 4 * i);
      const float4 a1 = ldg_stream4(th0_src + 4 * (i + stride)), t1 = ldg_stream4(tgt_src + 4 * (i + stride));
      *reinterpret_cast<float4*>(th0_dst + 4 * i) = a0; *reinterpret_cast<float4*>(tgt_dst + 4 * i) = t0;
      *reinterpret_cast<float4*>(th0_dst + 4 * (i + stride)) = a1; *reinterpret_cast<float4*>(tgt_dst + 4 * (i + stride)) = t1;
      float d;
      d = a0.x - t0.x; den = fmaf(d, d, den); d = a0.y - t0.y; den = fmaf(d, d, den);
      d = a0.z - t0.z; den = fmaf(d, d, den); d = a0.w - t0.w; den = fmaf(d, d, den);
      d = a1.x - t1.x; den = fmaf(d, d, den); d = a1.y - t1.y; den = fmaf(d, d, den);
      d = a1.z - t1.z; den = fmaf(d, d, den); d = a1.w - t1.w; den = fmaf(d, d, den);
    }
    for (; i < n4; i += stride) {
      const float4 a0 = ldg_stream4(th0_src + 4 * i), t0 = ldg_stream4(tgt_src + 4 * i);
      *reinterpret_cast<float4*>(th0_dst + 4 * i) = a0; *reinterpret_cast<float4*>(tgt_dst + 4 * i) = t0;
      float d;
      d = a0.x - t0.x; den = fmaf(d, d, den); d = a0.y - t0.y; den = fmaf(d, d, den);
      d = a0.z - t0.z; den = fmaf(d, d, den); d = a0.w - t0.w; den = fmaf(d, d, den);
    }
    for (int64_t j = (n4 << 2) + tid; j < n; j += stride) {
      const float a = th0_src[j], t = tgt_src[j];
      th0_dst[j] = a; tgt_dst[j] = t;
      den = fmaf(a - t, a - t, den);
    }
  } else {
    for (int64_t j = tid; j < n; j += stride) {
      const float a = th0_src[j], t = tgt_src[j];
      th0_dst[j] = a; tgt_dst[j] = t;
      den = fmaf(a - t, a - t, den);
    }
  }
  finish_sum(den, parts, ticket, [&](double s) { *den_out = (float)s; });
}
__global__ void __launch_bounds__(kStreamThreads) stage_segment_kernel(const float* __restrict__ th0_src,
                                                                       const float* __restrict__ tgt_src,
                                                                       float* __restrict__ th0_dst, float* __restrict__ tgt_dst,
                                                                       int64_t n, int vec, double* __restrict__ parts,
                                                                       unsigned int* __restrict__ ticket, float* __restrict__ den_out) {
  pdl_enter();
  stage_segment_body(th0_src, tgt_src, th0_dst, tgt_dst, n, vec, parts, ticket, den_out);
}
// The same pass with the two SOURCE addresses read from a device table: the kernel's arguments are then the same for every
// call, so it can be the first node of the engine's replayed launch graph (the sources change every iteration; the table is
// written by set_stage_sources_kernel just before the graph is launched).
__global__ void __launch_bounds__(kStreamThreads) stage_segment_indirect_kernel(const float* const* __restrict__ table,
                                                                                float* __restrict__ th0_dst,
                                                                                float* __restrict__ tgt_dst, int64_t n,
                                                                                double* __restrict__ parts,
                                                                                unsigned int* __restrict__ ticket,
                                                                                float* __restrict__ den_out) {
  pdl_enter();
  const float* th0_src = table[0];
  const float* tgt_src = table[1];
  const int vec = ((reinterpret_cast<uintptr_t>(th0_src) | reinterpret_cast<uintptr_t>(tgt_src) | reinterpret_cast<uintptr_t>(th0_dst) |
                    reinterpret_cast<uintptr_t>(tgt_dst)) & 15) == 0;
  stage_segment_body(th0_src, tgt_src, th0_dst, tgt_dst, n, vec, parts, ticket, den_out);
}
__global__ void set_stage_sources_kernel(const void** table, const float* th0_src, const float* tgt_src, const void* perms,
                                         unsigned int* ticket) {
  pdl_enter();
  if (threadIdx.x == 0) { table[0] = th0_src; table[1] = tgt_src; table[2] = perms; *ticket = 0u; }
}

// num = sum (thK - tgt)^2 ; a = 2 (thK - tgt) / den                      (12 B / param: 2 reads + 1 write)
// The pass leaves fp64 block partials; match_finish_kernel adds them in index order (deterministic) and writes
// out3 = {num, den, num / den}.  Nothing on the reverse sweep's critical path needs `num`, so the engine runs the finish on a
// side branch; keeping it out of the streaming kernel also removes that kernel's serial tail (ticket -> partial loads ->
// block sum: ~3 us after the last byte has moved, the difference between 0.68 and 0.85 of HBM peak).
__global__ void __launch_bounds__(kStreamThreads) match_final_kernel(const float* __restrict__ thK, const float* __restrict__ tgt,
                                                                     const float* __restrict__ den_p, int64_t n, int vec,
                                                                     double* __restrict__ parts, float* __restrict__ a) {
  pdl_enter();
  __shared__ double scratch[34];
  const float c = 2.0f / *den_p;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  float num = 0.f;
  if (vec) {
    const int64_t n4 = n >> 2;
    int64_t i = tid;
    for (; i + stride < n4; i += 2 * stride) {
      const float4 x0 = ldg_stream4(thK + 4 * i), t0 = ldg_stream4(tgt + 4 * i);
      const float4 x1 = ldg_stream4(thK + 4 * (i + stride)), t1 = ldg_stream4(tgt + 4 * (i + stride));
      const float4 d0 = make_float4(x0.x - t0.x, x0.y - t0.y, x0.z - t0.z, x0.w - t0.w);
      const float4 d1 = make_float4(x1.x - t1.x, x1.y - t1.y, x1.z - t1.z, x1.w - t1.w);
      num = fmaf(d0.x, d0.x, num); num = fmaf(d0.y, d0.y, num); num = fmaf(d0.z, d0.z, num); num = fmaf(d0.w, d0.w, num);
      num = fmaf(d1.x, d1.x, num); num = fmaf(d1.y, d1.y, num); num = fmaf(d1.z, d1.z, num); num = fmaf(d1.w, d1.w, num);
      // (L1::no_allocate store: +4 % in dev/stream_bench_test; the line still lands in L2, where the first reverse step finds it)
      stg_stream4(a + 4 * i, make_float4(c * d0.x, c * d0.y, c * d0.z, c * d0.w));
      stg_stream4(a + 4 * (i + stride), make_float4(c * d1.x, c * d1.y, c * d1.z, c * d1.w));
    }
    for (; i < n4; i += stride) {
      const float4 x0 = ldg_stream4(thK + 4 * i), t0 = ldg_stream4(tgt + 4 * i);
      const float4 d0 = make_float4(x0.x - t0.x, x0.y - t0.y, x0.z - t0.z, x0.w - t0.w);
      num = fmaf(d0.x, d0.x, num); num = fmaf(d0.y, d0.y, num); num = fmaf(d0.z, d0.z, num); num = fmaf(d0.w, d0.w, num);
      stg_stream4(a + 4 * i, make_float4(c * d0.x, c * d0.y, c * d0.z, c * d0.w));
    }
    for (int64_t j = (n4 << 2) + tid; j < n; j += stride) {
      const float d = thK[j] - tgt[j];
      num = fmaf(d, d, num);
      a[j] = c * d;
    }
  } else {
    for (int64_t j = tid; j < n; j += stride) {
      const float d = thK[j] - tgt[j];
      num = fmaf(d, d, num);
      a[j] = c * d;
    }
  }
  const double b = block_sum<double>((double)num, scratch);
  if (threadIdx.x == 0) parts[blockIdx.x] = b;
}
__global__ void __launch_bounds__(256) match_finish_kernel(const double* __restrict__ parts, int n_parts, const float* __restrict__ den_p,
                                                           float* __restrict__ out3) {
  pdl_enter();
  __shared__ double scratch[34];
  double s = 0.0;
  for (int i = threadIdx.x; i < n_parts; i += blockDim.x) s += parts[i];
  s = block_sum<double>(s, scratch);
  if (threadIdx.x == 0) {
    const float den = *den_p;
    out3[0] = (float)s;
    out3[1] = den;
    out3[2] = (float)(s / (double)den);
  }
}

// ---- outer update: the three torch.optim.SGD(momentum = 0.5) steps of distill.py:233-241, 603-613 in ONE launch ---------
// buf = first ? g : momentum * buf + g ; p -= lr * buf   for the synthetic image embeddings, the synthetic text and the two
// learnable student learning rates.  `gscale` multiplies every gradient first (1 for the sum over segments / ranks,
// 1 / segments for their mean).  A non-finite loss (distill.py:599-600: the reference stops BEFORE stepping) leaves every
// parameter untouched and raises *skipped.
struct OuterUpdateArgs {
  float* p[2]; const float* g[2]; float* buf[2]; int64_t n[2]; float lr[2];
  float* syn_lr[2];          // {syn_lr_img, syn_lr_txt}: one-element device scalars, each nullable
  float* buf_lr; float lr_lr;
  const float* g_lr_img;     // nullable: gradient of syn_lr_img (the fork's logit-scale path), else 0
  const float* g_lr_txt;
  const float* loss;         // nullable
  int* skipped;              // nullable
  float momentum, gscale; int first;
};
__global__ void __launch_bounds__(kStreamThreads) outer_update_kernel(const OuterUpdateArgs A) {
  pdl_enter();
  if (A.loss != nullptr && !isfinite(*A.loss)) {
    if (A.skipped != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *A.skipped = 1;
    return;
  }
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    float* __restrict__ p = A.p[t]; const float* __restrict__ g = A.g[t]; float* __restrict__ buf = A.buf[t];
    const float lr = A.lr[t];
    for (int64_t i = tid; i < A.n[t]; i += stride) {
      const float gg = A.gscale * g[i];
      const float b = A.first ? gg : fmaf(A.momentum, buf[i], gg);
      buf[i] = b;
      p[i] -= lr * b;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < 2 && A.syn_lr[threadIdx.x] != nullptr) {
    const float* gp = threadIdx.x == 0 ? A.g_lr_img : A.g_lr_txt;
    const float gg = gp ? A.gscale * (*gp) : 0.f;
    const float b = A.first ? gg : fmaf(A.momentum, A.buf_lr[threadIdx.x], gg);
    A.buf_lr[threadIdx.x] = b;
    *A.syn_lr[threadIdx.x] -= A.lr_lr * b;
  }
}

// ---- dropout masks of the student steps (networks.py:636,643: nn.Dropout(0.1), students in train mode distill.py:446-447) ----
// Philox4x32-10 keyed by state[0] (seed), counter = (element group, state[1] = draws so far); element e of the launch gets
// lane e % 4 of group e / 4.  mask = keep ? 1 / (1 - p) : 0, keep iff u >= p with u uniform in [0, 1) (24 bits).
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__global__ void __launch_bounds__(kStreamThreads) dropout_masks_kernel(float* __restrict__ masks, int64_t n, float p,
                                                                       const unsigned long long* __restrict__ state) {
  pdl_enter();
  const unsigned long long seed = state[0], draw = state[1];
  const float keep_scale = 1.0f / (1.0f - p);
  const int64_t groups = (n + 3) >> 2;
  for (int64_t gidx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; gidx < groups; gidx += (int64_t)gridDim.x * blockDim.x) {
    uint32_t r[4];
    philox4x32_10((uint32_t)gidx, (uint32_t)(gidx >> 32), (uint32_t)draw, (uint32_t)(draw >> 32), (uint32_t)seed,
                  (uint32_t)(seed >> 32), r);
    float m[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) m[j] = ((r[j] >> 8) * (1.0f / 16777216.0f) >= p) ? keep_scale : 0.f;
    const int64_t e = gidx << 2;
    if (e + 3 < n && ((reinterpret_cast<uintptr_t>(masks + e) & 15) == 0)) {
      *reinterpret_cast<float4*>(masks + e) = make_float4(m[0], m[1], m[2], m[3]);
    } else {
      for (int j = 0; j < 4 && e + j < n; ++j) masks[e + j] = m[j];
    }
  }
}
__global__ void rng_advance_kernel(unsigned long long* state) {
  pdl_enter();
  if (threadIdx.x == 0) state[1] += 1ull;
}

// ---- host launchers --------------------------------------------------------------------------------
int flat_sgd_step(const float* theta, const float* grad, const float* lr, float* out, int64_t n, cudaStream_t st) {
  if (n <= 0) return VLDD_OK;
  const int vec = aligned16(theta) && aligned16(grad) && aligned16(out);
  launch_k(flat_sgd_step_kernel, stream_grid(n / 4 + 1), kStreamThreads, 0, st, theta, grad, lr, out, n, vec);
  return check_launch("flat_sgd_step");
}

int64_t match_loss_scratch_bytes() { return (int64_t)(2 * kMaxSMs * kStreamBlocksPerSM) * sizeof(double) + 16; }

int match_loss_fwd(const float* thK, const float* tgt, const float* th0, int64_t n, float* out3, void* scratch,
                   cudaStream_t st) {
  // scratch layout: [ticket u32 | pad][block partials f64 x 2*grid]; the ticket must be zero on first use
  unsigned int* ticket = reinterpret_cast<unsigned int*>(scratch);
  double* parts = reinterpret_cast<double*>(reinterpret_cast<char*>(scratch) + 16);
  const int vec = aligned16(thK) && aligned16(tgt) && aligned16(th0);
  const int grid = stream_grid(n / 4 + 1);
  launch_k(match_loss_fwd_kernel, grid, kStreamThreads, 0, st, thK, tgt, th0, n, vec, parts, ticket, out3);
  return check_launch("match_loss_fwd");
}

int match_loss_bwd(const float* thK, const float* tgt, const float* num_den, const float* gout, float* a, int64_t n,
                   cudaStream_t st) {
  if (n <= 0) return VLDD_OK;
  const int vec = aligned16(thK) && aligned16(tgt) && aligned16(a);
  launch_k(match_loss_bwd_kernel, stream_grid(n / 4 + 1), kStreamThreads, 0, st, thK, tgt, num_den, gout, a, n, vec);
  return check_launch("match_loss_bwd");
}

int stage_segment(const float* th0_src, const float* tgt_src, float* th0_dst, float* tgt_dst, int64_t n, float* den_out,
                  void* scratch, cudaStream_t st) {
  unsigned int* ticket = reinterpret_cast<unsigned int*>(scratch);
  double* parts = reinterpret_cast<double*>(reinterpret_cast<char*>(scratch) + 16);
  const int vec = aligned16(th0_src) && aligned16(tgt_src) && aligned16(th0_dst) && aligned16(tgt_dst);
  launch_k(stage_segment_kernel, stream_grid(n / 8 + 1), kStreamThreads, 0, st, th0_src, tgt_src, th0_dst, tgt_dst, n, vec, parts,
           ticket, den_out);
  return check_launch("stage_segment");
}

int set_stage_sources(const float* th0_src, const float* tgt_src, const void* perms, void* table, void* scratch, cudaStream_t st) {
  launch_k(set_stage_sources_kernel, 1, 32, 0, st, reinterpret_cast<const void**>(table), th0_src, tgt_src, perms,
           reinterpret_cast<unsigned int*>(scratch));
  return check_launch("set_stage_sources");
}
int stage_segment_indirect(const void* table, float* th0_dst, float* tgt_dst, int64_t n, float* den_out, void* scratch,
                           cudaStream_t st) {
  unsigned int* ticket = reinterpret_cast<unsigned int*>(scratch);
  double* parts = reinterpret_cast<double*>(reinterpret_cast<char*>(scratch) + 16);
  launch_k(stage_segment_indirect_kernel, stream_grid(n / 8 + 1), kStreamThreads, 0, st,
           reinterpret_cast<const float* const*>(table), th0_dst, tgt_dst, n, parts, ticket, den_out);
  return check_launch("stage_segment_indirect");
}

// streaming pass on `st`; the finish on `finish_st` -- the CALLER orders finish_st after the pass (same stream, or an event edge)
int match_final_pass(const float* thK, const float* tgt, const float* den, int64_t n, float* a, void* scratch, cudaStream_t st) {
  double* parts = reinterpret_cast<double*>(reinterpret_cast<char*>(scratch) + 16);
  const int vec = aligned16(thK) && aligned16(tgt) && aligned16(a);
  launch_k(match_final_kernel, stream_grid(n / 8 + 1), kStreamThreads, 0, st, thK, tgt, den, n, vec, parts, a);
  return check_launch("match_final_pass");
}
const double* match_final_parts(const void* scratch) { return reinterpret_cast<const double*>(reinterpret_cast<const char*>(scratch) + 16); }
int match_final_n_parts(int64_t n) { return stream_grid(n / 8 + 1); }
int match_final_finish(const float* den, int64_t n, float* out3, void* scratch, cudaStream_t finish_st) {
  const double* parts = reinterpret_cast<const double*>(reinterpret_cast<char*>(scratch) + 16);
  launch_k(match_finish_kernel, 1, 256, 0, finish_st, parts, stream_grid(n / 8 + 1), den, out3);
  return check_launch("match_final_finish");
}

int outer_update(float* U, const float* gU, float* bufU, int64_t nU, float lrU, float* Y, const float* gY, float* bufY,
                 int64_t nY, float lrY, float* syn_lr_img, float* syn_lr_txt, const float* g_lr_img, const float* g_lr_txt,
                 float* buf_lr, float lr_lr, float momentum, int first, float gscale, const float* loss, int* skipped,
                 cudaStream_t st) {
  OuterUpdateArgs A;
  A.p[0] = U; A.g[0] = gU; A.buf[0] = bufU; A.n[0] = nU; A.lr[0] = lrU;
  A.p[1] = Y; A.g[1] = gY; A.buf[1] = bufY; A.n[1] = nY; A.lr[1] = lrY;
  A.syn_lr[0] = syn_lr_img; A.syn_lr[1] = syn_lr_txt; A.buf_lr = buf_lr; A.lr_lr = lr_lr;
  A.g_lr_img = g_lr_img; A.g_lr_txt = g_lr_txt; A.loss = loss; A.skipped = skipped;
  A.momentum = momentum; A.gscale = gscale; A.first = first;
  const int64_t nmax = nU > nY ? nU : nY;
  launch_k(outer_update_kernel, stream_grid(nmax + 1), kStreamThreads, 0, st, A);
  return check_launch("outer_update");
}

int dropout_masks(float* masks, int64_t n, float p, unsigned long long* state, int advance, cudaStream_t st) {
  if (n <= 0) return VLDD_OK;
  launch_k(dropout_masks_kernel, stream_grid(n / 4 + 1), kStreamThreads, 0, st, masks, n, p, (const unsigned long long*)state);
  if (advance) launch_k(rng_advance_kernel, 1, 32, 0, st, state);
  return check_launch("dropout_masks");
}

int momentum_sgd(float* p, const float* g, float* buf, float lr, float momentum, int first, int64_t n,
                 cudaStream_t st) {
  if (n <= 0) return VLDD_OK;
  const int vec = aligned16(p) && aligned16(g) && aligned16(buf);
  launch_k(momentum_sgd_kernel, stream_grid(n / 4 + 1), kStreamThreads, 0, st, p, g, buf, lr, momentum, first, n, vec);
  return check_launch("momentum_sgd");
}

}  // namespace vldd
