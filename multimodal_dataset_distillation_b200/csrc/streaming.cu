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

int momentum_sgd(float* p, const float* g, float* buf, float lr, float momentum, int first, int64_t n,
                 cudaStream_t st) {
  if (n <= 0) return VLDD_OK;
  const int vec = aligned16(p) && aligned16(g) && aligned16(buf);
  launch_k(momentum_sgd_kernel, stream_grid(n / 4 + 1), kStreamThreads, 0, st, p, g, buf, lr, momentum, first, n, vec);
  return check_launch("momentum_sgd");
}

}  // namespace vldd
