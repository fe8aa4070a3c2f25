// Unroll engine (Path 1): K flat-parameter SGD steps of the text_projection head on the bidirectional InfoNCE
// loss, the normalised parameter-matching loss, and the hand-written reverse sweep that replaces
// `grand_loss.backward()` through `autograd.grad(create_graph=True)`.
//
//   reference sites: distill.py:509-583 (unroll), 584-598 (matching loss), 606 (backward through the unroll)
//   algorithm:       DESIGN.md section 4 == oracle/distill_ref.py::unrolled_match_manual
//
// Forward sweep, step k:   g_k = dL/dtheta(theta_k; Y_b, Xn_b, s),  theta_{k+1} = theta_k - lr g_k   (update fused
//                          into the dW GEMM epilogues; every activation of the step is kept -- ~12 MB/step)
// Reverse sweep, step k:   with v = a_{k+1}:  tangent pass of the whole first-order step along theta_dot = v gives
//                          L_dot = <g_k, v>, H_k v, d/dY, d/dXn, d/ds;   a_k = a_{k+1} - lr H_k v (fused into the
//                          tangent dW GEMM epilogues), dlr -= L_dot, dY[perm] -= lr dY_dot, ...
#include "common.cuh"
#include "gemm_dispatch.cuh"
#include "head_kernels.cuh"
#include "nce_cluster.cuh"
#include "nce_fused.cuh"
#include "row_kernels_v4.cuh"
#include "kernels.h"
#include "engine.h"

#include <mutex>
#include <vector>

namespace vldd {

namespace {

constexpr int kMaxDevicesNce = 64;

struct Dims {
  int N, B, K, dt, d, Bp;   // Bp: leading dimension of the B x B matrices (multiple of 32 for the MN-major tensor maps)
  int64_t P, oW1, ob1, oW2, ob2, og, obt;
};

Dims make_dims(int N, int B, int K, int dt, int d) {
  Dims m;
  m.N = N; m.B = B; m.K = K; m.dt = dt; m.d = d; m.Bp = (B + 31) / 32 * 32;
  m.oW1 = 0;
  m.ob1 = (int64_t)d * dt;
  m.oW2 = m.ob1 + d;
  m.ob2 = m.oW2 + (int64_t)d * d;
  m.og = m.ob2 + d;
  m.obt = m.og + d;
  m.P = m.obt + d;
  return m;
}

struct Saved {  // activations of one forward step, all fp32
  float *Yb, *Xb, *p, *h, *rhat, *yn, *dyn, *dz, *dr, *df, *dh, *dp;  // [B,dt] [B,d] then [B,d] x10
  float *rstd, *nz, *q, *lse_r, *lse_c;                               // [B]
  float *S, *G, *Pr, *Pc;                                             // [B,B] (Pr, Pc: row / column softmax, small-batch path)
};

struct Bump {
  char* base; size_t off, cap;
  float* f(size_t n) {
    size_t bytes = ((n * sizeof(float) + 255) / 256) * 256;
    float* r = base ? reinterpret_cast<float*>(base + off) : nullptr;
    off += bytes;
    return r;
  }
};

struct Work {
  Saved sv[64];
  float *traj, *tgt, *adj0, *adj1, *Xn, *un, *dXn;
  float *pa, *pb, *pc, *pe;            // GEMM partial slabs / raw outputs (pc, pe: side-stream branches)
  float *pd, *hd, *rhatd, *ynd, *dzd, *drd, *dfd, *dpd, *t, *nzd, *Sd, *Gd, *rho, *kap, *rowA, *rowB;
  float *neg_one;                      // device constant -1 (first-order API)
  int* bad_index;                      // set by the gather kernel when a minibatch index is out of range
  void* ml_scratch;
  void* stage_scratch;                 // ticket + block partials of the staging kernel
  void* stage_table;                   // {theta_0, theta*, perms}: the per-call addresses (written outside the graph)
  int64_t* perms_copy;                 // [K, B] minibatch indices, copied by the gather kernel for the reverse sweep
  float* den;                          // |theta_0 - theta*|^2, produced while the segment is staged
  size_t bytes;
};

size_t partial_floats(const Dims& m) {
  const int B = m.B, d = m.d, dt = m.dt;
  size_t mx = 0;
  auto upd = [&](int M, int N, int Kt) {
    size_t v = (size_t)max_splits(M, N, Kt) * M * N;
    if (v > mx) mx = v;
  };
  upd(B, d, dt); upd(B, d, d); upd(B, d, 2 * d); upd(B, B, d); upd(B, d, B); upd(B, d, 2 * B); upd(B, dt, d);
  upd(B, dt, 2 * d);
  return mx;
}

void carve(Work& w, const Dims& m, void* base) {
  Bump b{reinterpret_cast<char*>(base), 0, 0};
  const size_t Bd = (size_t)m.B * m.d, BB = (size_t)m.B * m.Bp, Bpd = (size_t)m.Bp * m.d;
  w.traj = b.f((size_t)(m.K + 1) * m.P);
  w.tgt = b.f(m.P);
  w.adj0 = b.f(m.P);
  w.adj1 = b.f(m.P);
  w.Xn = b.f((size_t)m.N * m.d);
  w.un = b.f(m.N);
  w.dXn = b.f((size_t)m.N * m.d);
  for (int k = 0; k < m.K; ++k) {
    Saved& s = w.sv[k];
    s.Yb = b.f((size_t)m.B * m.dt);
    s.Xb = b.f(Bd); s.p = b.f(Bd); s.h = b.f(Bd); s.rhat = b.f(Bd); s.yn = b.f(Bd); s.dyn = b.f(Bd);
    s.dz = b.f(Bd); s.dr = b.f(Bd); s.df = b.f(Bd); s.dh = b.f(Bd); s.dp = b.f(Bd);
    s.rstd = b.f(m.B); s.nz = b.f(m.B); s.q = b.f(m.B); s.lse_r = b.f(m.B); s.lse_c = b.f(m.B);
    s.S = b.f(BB); s.G = b.f(BB); s.Pr = b.f(BB); s.Pc = b.f(BB);
  }
  const size_t pf = partial_floats(m);
  w.pa = b.f(pf > Bpd ? pf : Bpd);
  w.pb = b.f(pf > Bpd ? pf : Bpd);
  w.pc = b.f(Bd);
  w.pe = b.f((size_t)max_splits(m.B, m.dt, 2 * m.d) * m.B * m.dt);
  w.pd = b.f(Bd); w.hd = b.f(Bd); w.rhatd = b.f(Bd); w.ynd = b.f(Bd); w.dzd = b.f(Bd); w.drd = b.f(Bd);
  w.dfd = b.f(Bd); w.dpd = b.f(Bd);
  w.t = b.f(m.B); w.nzd = b.f(m.B); w.rho = b.f(m.B); w.kap = b.f(m.B); w.rowA = b.f(m.B); w.rowB = b.f(m.B);
  w.Sd = b.f(BB); w.Gd = b.f(BB);
  w.neg_one = b.f(4);
  w.bad_index = reinterpret_cast<int*>(b.f(4));
  w.ml_scratch = b.f((size_t)match_loss_scratch_bytes() / sizeof(float) + 8);
  w.stage_scratch = b.f((size_t)match_loss_scratch_bytes() / sizeof(float) + 8);
  w.den = b.f(4);
  w.stage_table = b.f(8);
  w.perms_copy = reinterpret_cast<int64_t*>(b.f(2 * (size_t)(m.K > 0 ? m.K : 1) * m.B));
  w.bytes = b.off;
}

__global__ void fill_kernel(float* p, float v, int n) {
  pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void reduce_slabs_kernel(const float* __restrict__ part, int splits, size_t stride, size_t n,
                                    float* __restrict__ out) {
  pdl_enter();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = sum_slabs(part, splits, stride, i);
}
__global__ void __launch_bounds__(256) dot_over_scale_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                             size_t n, const float* __restrict__ scale,
                                                             float* __restrict__ out) {
  pdl_enter();
  __shared__ float scratch[34];
  float acc = 0.f;
  for (size_t i = threadIdx.x; i < n; i += blockDim.x) acc = fmaf(a[i], b[i], acc);
  acc = block_sum<float>(acc, scratch);
  if (threadIdx.x == 0) *out = acc / (*scale);
}

#define CHECK_RC(x) do { int rc__ = (x); if (rc__) return rc__; } while (0)
// Everything a call has to clear, in ONE launch (as memset nodes these were 2 K + 7 serial ~1.1 us graph nodes ahead of the
// first kernel): the accumulators (dY, dXn, dlr / dscale, the matching-loss scratch, the bad-index flag) and S, G, Sd, Gd, whose
// padding columns [B, Bp) must read as zeros (they feed GEMMs as extra, all-zero rows).
constexpr int kMaxZeroSegs = 24;
struct ZeroSegs {
  void* p[kMaxZeroSegs];
  unsigned long long bytes[kMaxZeroSegs];      // multiples of 4
};
__global__ void __launch_bounds__(256) zero_segments_kernel(ZeroSegs z) {
  pdl_enter();
  char* p = static_cast<char*>(z.p[blockIdx.y]);
  const size_t bytes = z.bytes[blockIdx.y];
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (size_t)gridDim.x * blockDim.x;
  if (((reinterpret_cast<uintptr_t>(p) | bytes) & 15) == 0) {
    for (size_t i = t; i < bytes / 16; i += nt) reinterpret_cast<float4*>(p)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
    for (size_t i = t; i < bytes / 4; i += nt) reinterpret_cast<float*>(p)[i] = 0.f;
  }
}
struct ZeroList {
  ZeroSegs z;
  int n = 0;
  cudaStream_t st;
  int flush() {
    if (n == 0) return VLDD_OK;
    launch_k(zero_segments_kernel, dim3(16, n), 256, 0, st, z);
    n = 0;
    return check_launch("zero_segments");
  }
  int add(void* p, size_t bytes) {
    if (bytes == 0) return VLDD_OK;
    z.p[n] = p; z.bytes[n] = bytes;
    if (++n == kMaxZeroSegs) return flush();
    return VLDD_OK;
  }
};
int zero_square_matrices(const Dims& m, Work& w, ZeroList& zl) {
  if (m.Bp == m.B) return VLDD_OK;
  // S, G, Pr, Pc of a step are carved next to each other (carve()): S and G are cleared as the run [S, G + B*Bp)
  const size_t bytes = (size_t)m.B * m.Bp * sizeof(float);
  for (int k = 0; k < m.K; ++k) {
    if (w.sv[k].G == w.sv[k].S + (size_t)m.B * m.Bp) CHECK_RC(zl.add(w.sv[k].S, 2 * bytes));
    else { CHECK_RC(zl.add(w.sv[k].S, bytes)); CHECK_RC(zl.add(w.sv[k].G, bytes)); }
  }
  if (w.Gd == w.Sd + (size_t)m.B * m.Bp) CHECK_RC(zl.add(w.Sd, 2 * bytes));
  else { CHECK_RC(zl.add(w.Sd, bytes)); CHECK_RC(zl.add(w.Gd, bytes)); }
  return VLDD_OK;
}

// grid of the element-wise kernels that take the 128-bit path when the row length allows (4 elements per thread)
inline int ew_grid(size_t n);
inline int ew_grid4(size_t n, int d) { return (d & 3) == 0 ? ew_grid(n / 4) : ew_grid(n); }
inline int ew_grid(size_t n) {
  size_t g = (n + 255) / 256;
  const size_t cap = (size_t)num_sms() * 8;
  return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}


// VLDD_EARLY_LOADS=0 disables the pre-wait operand loads of the tensor-core GEMMs (A/B runs)
bool early_loads() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("VLDD_EARLY_LOADS");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

// InfoNCE as one cluster launch (nce_cluster.cuh) is opt-in, VLDD_NCE=cluster.  Measured on B200 at the Flickr shape
// (bench.py, ms / iteration): three / four small launches 1.906, one 8-CTA cluster launch 1.936 -- the 36-slab logits
// reduction (1.4 MB from L2) is spread over 100 SMs by the row kernels but over only 8 by the cluster, and under
// programmatic dependent launch consecutive small kernels cost little more than their execution time.
bool nce_fused(int B, int ld) {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("VLDD_NCE");
    mode = (e && strcmp(e, "cluster") == 0) ? 1 : 0;
    if (mode == 1) {
      const int cap = 200 * 1024;
      if (cudaFuncSetAttribute(nce_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, cap) != cudaSuccess ||
          cudaFuncSetAttribute(nce_t_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, cap) != cudaSuccess) {
        cudaGetLastError();
        mode = 0;
      }
    }
  }
  return mode == 1 && nce_cluster_ok(B, ld);
}

// The two-kernel CUDA-core InfoNCE block (nce_fused.cuh: exact-fp32 logits without split-K slabs, then one kernel that
// derives the softmax statistics in every CTA and produces G^T X from shared memory) is OPT-IN, VLDD_NCE=fused, for batches
// of up to ~120 pairs.  Measured on B200 at the Flickr shape (bench.py, ms / iteration): tensor-core GEMMs + row / column
// kernels 1.694, fused block 1.892 (first version with per-thread load loops: 2.19).  The five launches it replaces cost
// 3-5 us each; the replacement's phases (bulk load, row pass, column pass, G, G^T X) are dependent stages of ~2-6 us
// INSIDE one CTA at 4 warps per scheduler (csrc/dev/nce_fused_test.cu prints the phase times): 22 us against ~20 us.
// On this chain a launch boundary (~1 us with programmatic dependent launch) is cheaper than a block-wide dependent phase.
bool nce_small(int B, int d, int ld, const void* a, const void* b, const void* c) {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("VLDD_NCE");
    mode = (e && strcmp(e, "fused") == 0) ? 1 : 0;
  }
  if (mode != 1 || !nce_small_ok(B, d, ld, a, b, c)) return false;
  static bool configured[kMaxDevicesNce] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevicesNce) return false;
  if (!configured[dev]) {
    const int cap = 224 * 1024;            // three B x B matrices + operand blocks: 141 KB at B = 100, 216 KB at B = 124
    if (cudaFuncSetAttribute(small_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, cap) != cudaSuccess ||
        cudaFuncSetAttribute(nce_gx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, cap) != cudaSuccess ||
        cudaFuncSetAttribute(nce_t_gx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, cap) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    configured[dev] = true;
  }
  return true;
}

// VLDD_PROFILE=1: serialise everything on one stream and time every launch with events (developer aid; prints a
// per-call-site table to stderr at the end of vldd_unrolled_match).
struct ProfMark { const char* label; cudaEvent_t ev; };
std::vector<ProfMark> g_prof;
bool prof_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("VLDD_PROFILE"); v = (e && strcmp(e, "1") == 0) ? 1 : 0; }
  return v == 1;
}
void prof_mark(const char* label, cudaStream_t st) {
  if (!prof_enabled()) return;
  cudaEvent_t e;
  cudaEventCreate(&e);
  cudaEventRecord(e, st);
  g_prof.push_back(ProfMark{label, e});
}
void prof_report() {
  if (!prof_enabled() || g_prof.size() < 2) return;
  cudaDeviceSynchronize();
  struct Acc { const char* label; double ms; int n; };
  std::vector<Acc> acc;
  double total = 0;
  for (size_t i = 1; i < g_prof.size(); ++i) {
    float ms = 0;
    cudaEventElapsedTime(&ms, g_prof[i - 1].ev, g_prof[i].ev);
    total += ms;
    bool found = false;
    for (auto& a : acc) if (strcmp(a.label, g_prof[i].label) == 0) { a.ms += ms; a.n++; found = true; break; }
    if (!found) acc.push_back(Acc{g_prof[i].label, ms, 1});
  }
  fprintf(stderr, "[vldd profile] one call, serialised on one stream: %.1f us\n", total * 1e3);
  for (auto& a : acc) fprintf(stderr, "[vldd profile] %9.1f us %5.1f%% x%3d  %s\n", a.ms * 1e3, 100 * a.ms / total, a.n, a.label);
  for (auto& m : g_prof) cudaEventDestroy(m.ev);
  g_prof.clear();
}
#define MARK(label) prof_mark(label, st)

// Independent branches of a step (weight-gradient GEMMs, the dXn / dY side products) run on two side streams that
// fork from / join into the main stream with events; under graph capture this becomes parallel branches of the graph.
struct Lanes {
  cudaStream_t main, s1, s2;
  bool s1_busy, s2_busy;
  // reverse sweep: side-lane work of the PREVIOUS step that the next step has not waited for yet (tangent_step)
  cudaEvent_t ev_small = nullptr, ev_s1_tail = nullptr, ev_s2 = nullptr;
  // the dY product of the previous reverse step, not launched yet (run_pending_dy)
  struct PendingDy { bool live = false; const float *dpd, *W1, *dp, *V1; const int64_t* perm; } dy;
};
std::mutex g_lane_mu;
// side streams, event pool and capture stream are per device (a process may drive several GPUs)
constexpr int kMaxDevices = 64;
struct DeviceState {
  cudaStream_t side1 = nullptr, side2 = nullptr, capture = nullptr;
  std::vector<cudaEvent_t> events;
  size_t event_next = 0;
};
DeviceState g_dev[kMaxDevices];
DeviceState* g_cur = nullptr;        // state of the device the current call runs on (set under g_graph_mu / g_lane_mu)

int current_device_state(DeviceState** out) {
  int dev = 0;
  VLDD_CUDA(cudaGetDevice(&dev));
  VLDD_REQUIRE(dev >= 0 && dev < kMaxDevices, "device ordinal %d out of range", dev);
  *out = &g_dev[dev];
  return VLDD_OK;
}

bool relaxed_joins() {          // VLDD_JOIN=step restores the whole-step join of the reverse sweep (developer comparison)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("VLDD_JOIN");
    v = (e && strcmp(e, "step") == 0) ? 0 : 1;
  }
  return v == 1 && !prof_enabled();
}
bool env_on(const char* name, bool dflt) {
  const char* e = getenv(name);
  if (!e) return dflt;
  return strcmp(e, "0") != 0;
}
// (stream priorities -- side lanes lowest, capture stream highest -- were measured and change nothing: the persistent GEMM CTAs
//  of a side product hold their SMs until they are done, 1.5991 vs 1.5981 ms per iteration)
// the dY product of reverse step k runs in step k-1's InfoNCE window (few busy SMs) instead of next to step k-1's first GEMM
bool defer_dy() { static const bool v = env_on("VLDD_DEFER_DY", true); return v; }
int dw2_grid_cap() {              // persistent CTAs of the d x d weight GEMMs (side lanes); 0 = all SMs, -2 = half of them
  static int v = -1000;
  if (v == -1000) { const char* e = getenv("VLDD_DW2_CTAS"); v = e ? atoi(e) : -2; }
  return v;
}
int dw2t_grid_cap() {
  static int v = -1000;
  if (v == -1000) { const char* e = getenv("VLDD_DW2T_CTAS"); v = e ? atoi(e) : -1001; }
  return v == -1001 ? dw2_grid_cap() : v;
}
int dy_grid_cap() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("VLDD_DY_CTAS"); v = e ? atoi(e) : 96; }
  return v;
}
int lanes_init(Lanes& L, cudaStream_t main) {
  std::lock_guard<std::mutex> lock(g_lane_mu);
  CHECK_RC(current_device_state(&g_cur));
  if (g_cur->side1 == nullptr) {
    VLDD_CUDA(cudaStreamCreateWithFlags(&g_cur->side1, cudaStreamNonBlocking));
    VLDD_CUDA(cudaStreamCreateWithFlags(&g_cur->side2, cudaStreamNonBlocking));
  }
  L.main = main; L.s1 = g_cur->side1; L.s2 = g_cur->side2; L.s1_busy = false; L.s2_busy = false;
  if (prof_enabled()) { L.s1 = main; L.s2 = main; }
  g_cur->event_next = 0;
  return VLDD_OK;
}
int lane_edge(cudaStream_t from, cudaStream_t to) {   // everything enqueued on `from` so far happens-before later work on `to`
  if (from == to) return VLDD_OK;
  if (g_cur->event_next == g_cur->events.size()) {
    cudaEvent_t e;
    VLDD_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    g_cur->events.push_back(e);
  }
  cudaEvent_t e = g_cur->events[g_cur->event_next++];
  VLDD_CUDA(cudaEventRecord(e, from));
  VLDD_CUDA(cudaStreamWaitEvent(to, e, 0));
  return VLDD_OK;
}
int lane_record(cudaStream_t on, cudaEvent_t* out) {     // an event after everything enqueued on `on` so far
  if (g_cur->event_next == g_cur->events.size()) {
    cudaEvent_t e;
    VLDD_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    g_cur->events.push_back(e);
  }
  *out = g_cur->events[g_cur->event_next++];
  VLDD_CUDA(cudaEventRecord(*out, on));
  return VLDD_OK;
}
int lane_wait(cudaStream_t st, cudaEvent_t* ev) {        // consume a pending event (no-op when there is none)
  if (*ev == nullptr) return VLDD_OK;
  VLDD_CUDA(cudaStreamWaitEvent(st, *ev, 0));
  *ev = nullptr;
  return VLDD_OK;
}
int lanes_join(Lanes& L) {
  L.ev_small = L.ev_s1_tail = L.ev_s2 = nullptr;          // subsumed by the full joins below
  if (L.s1_busy) { CHECK_RC(lane_edge(L.s1, L.main)); L.s1_busy = false; }
  if (L.s2_busy) { CHECK_RC(lane_edge(L.s2, L.main)); L.s2_busy = false; }
  return VLDD_OK;
}

// ---------------------------------------------------------------------------------------------------
// One forward (first-order) step.  th_src/th_dst: theta_k -> theta_{k+1} (th_src == nullptr: dst = -lr*g).
// ---------------------------------------------------------------------------------------------------
// `chain`: the call sits inside the engine's own launch chain (unrolled_match), where every kernel triggers its
// dependents only after its own griddepcontrol.wait has returned; operands produced two or more kernels upstream
// (weights of this step, the gathered minibatch) may then be fetched before the wait (tc_gemm.cuh, old_mask:
// bit 0 = A operands, bit 1 = B operands).  step_index < 0 or !chain: no early loads.
int forward_step(const Dims& m, Work& w, Saved& s, const float* th, const float* upd_src, float* upd_dst,
                 const float* lr, const float* scale, const float* mask, float* ce_out, Lanes& L, bool chain = false,
                 int step_index = 0) {
  const int kOldA = chain ? 1 : 0, kOldB = chain ? 2 : 0;
  // theta_k[W1] comes from the side-stream GEMM joined right here (one hop); in step 0 from the staging pass, which is this
  // GEMM's programmatic predecessor on the main lane: never an early load.  The gathered minibatch (side lane 1 in step 0,
  // joined by an event = complete before this kernel may start; steps old afterwards) always is.
  const int p_mask = chain ? 1 : 0;
  (void)step_index;
  cudaStream_t st = L.main;
  CHECK_RC(lanes_join(L));          // theta_k must be complete (previous step's weight-gradient branches)
  const int B = m.B, d = m.d, dt = m.dt, Bp = m.Bp;
  const size_t Bd = (size_t)B * d;
  const float *W1 = th + m.oW1, *b1 = th + m.ob1, *W2 = th + m.oW2, *b2 = th + m.ob2, *gam = th + m.og, *bet = th + m.obt;
  // p = Yb W1^T + b1 ; h = gelu(p)
  int sp = 1;
  CHECK_RC((gemm_partial<true, true>(gemm_ops(s.Yb, dt, W1, dt, B, d, dt), w.pa, &sp, st, p_mask)));
  prof_mark("gemm_partial<true,true> A=s.Yb", st);
  launch_k(epi_p_kernel, ew_grid4(Bd, d), 256, 0, st, w.pa, sp, Bd, b1, B, d, s.p, s.h);
  prof_mark("epi_p_kernel", st);
  // f = h W2^T + b2 ; r = mask f + p ; LN ; normalise
  CHECK_RC((gemm_partial<true, true>(gemm_ops(s.h, d, W2, d, B, d, d), w.pa, &sp, st, kOldB)));
  prof_mark("gemm_partial<true,true> A=s.h", st);
  if (row_v4_ok(d))
    VLDD_ROW_V4_DISPATCH(d, ln_fwd_v4_kernel, (B, 256, 0, st), (w.pa, sp, Bd, b2, mask, s.p, gam, bet, d, s.rhat, nullptr,
                                                                   s.yn, s.rstd, s.nz));
  else
    launch_k(ln_fwd_kernel, B, 256, d * sizeof(float), st, w.pa, sp, Bd, b2, mask, s.p, gam, bet, d, s.rhat, nullptr, s.yn,
                                                     s.rstd, s.nz);
  prof_mark("ln_fwd_kernel", st);
  // S = scale * Xb Yn^T ; lse ; G ; loss ; dyn_raw[j,:] = sum_i G[i,j] Xb[i,:]
  if (nce_small(B, d, Bp, s.Xb, s.yn, w.pb)) {
    launch_k(small_scores_kernel, dim3(ceil_div(B, kScoresTJ), ceil_div(B, kScoresTI)), kScoresThreads, scores_smem_bytes(d), st, (const float*)s.Xb, (const float*)s.yn, B, d, scale,
             s.S, Bp);
    prof_mark("small_scores_kernel", st);
    launch_k(nce_gx_kernel, ceil_div(d, kNceGxCols), kNceGxThreads, nce_gx_smem_bytes(B, Bp), st, (const float*)s.S,
             (const float*)s.Xb, B, Bp, d, s.lse_r, s.lse_c, s.G, s.Pr, s.Pc, ce_out, w.pb);
    prof_mark("nce_gx_kernel", st);
  } else {
  CHECK_RC((gemm_partial<true, true>(gemm_ops(s.Xb, d, s.yn, d, B, B, d), w.pa, &sp, st, kOldA)));
  prof_mark("gemm_partial<true,true> A=s.Xb", st);
  if (nce_fused(B, Bp)) {
    launch_cluster_k(nce_cluster_kernel, kNceCluster, kNceThreads, nce_cluster_smem_bytes(B, Bp), st, w.pa, sp, (size_t)B * B,
                     scale, B, Bp, s.S, s.lse_r, s.lse_c, s.G, ce_out);
    prof_mark("nce_cluster_kernel", st);
  } else {
    launch_k(nce_rows_kernel, B, 128, 0, st, w.pa, sp, (size_t)B * B, scale, B, Bp, s.S, s.lse_r);
    prof_mark("nce_rows_kernel", st);
    launch_k(nce_cols_kernel, B, 128, 0, st, s.S, B, Bp, s.lse_c);
    prof_mark("nce_cols_kernel", st);
    launch_k(nce_grad_kernel, B, 128, 0, st, s.S, s.lse_r, s.lse_c, B, Bp, s.G, ce_out);
    prof_mark("nce_grad_kernel", st);
  }
  // dyn_raw[j,:] = sum_i G[i,j] Xb[i,:]
  // (G is stored with leading dimension Bp and zero padding columns: rows B..Bp-1 of the product are zeros)
  CHECK_RC((gemm_store<false, false>(gemm_ops(s.G, Bp, s.Xb, d, Bp, d, B), w.pb, d, 1.0f, st, kOldB)));
  prof_mark("gemm_store<false,false> A=s.G", st);
  }
  if (row_v4_ok(d))
    VLDD_ROW_V4_DISPATCH(d, norm_ln_bwd_v4_kernel, (B, 256, 0, st), (w.pb, scale, s.yn, s.nz, s.rhat, s.rstd, gam, mask, d,
                                                                        s.dyn, s.q, s.dz, s.dr, s.df));
  else
    launch_k(norm_ln_bwd_kernel, B, 256, 0, st, w.pb, scale, s.yn, s.nz, s.rhat, s.rstd, gam, mask, d, s.dyn, s.q, s.dz, s.dr,
                                          s.df);
  prof_mark("norm_ln_bwd_kernel", st);
  // branch 1: theta_{k+1}[W2] = theta_k[W2] - lr df^T h   (needs only df, h)
  CHECK_RC(lane_edge(st, L.s1));
  L.s1_busy = true;
  {
    tc::GridCapScope cap(dw2_grid_cap());
    CHECK_RC((gemm_axpy<false, false>(gemm_ops(s.df, d, s.h, d, d, d, B), upd_src ? upd_src + m.oW2 : nullptr,
                                      upd_dst + m.oW2, d, lr, L.s1)));
  }
  prof_mark("gemm_axpy<false,false> A=s.df", L.s1);
  // dh = df W2 ; dp = dh gelu'(p) + dr
  CHECK_RC((gemm_partial<true, false>(gemm_ops(s.df, d, W2, d, B, d, d), w.pa, &sp, st, kOldB)));
  prof_mark("gemm_partial<true,false> A=s.df", st);
  launch_k(epi_dp_kernel, ew_grid4(Bd, d), 256, 0, st, w.pa, sp, Bd, s.p, s.dr, Bd, s.dh, s.dp);
  prof_mark("epi_dp_kernel", st);
  // branch 2: theta_{k+1}[W1] = theta_k[W1] - lr dp^T Yb ;  main: small params
  CHECK_RC(lane_edge(st, L.s2));
  L.s2_busy = true;
  CHECK_RC((gemm_axpy<false, false>(gemm_ops(s.dp, d, s.Yb, dt, d, dt, B), upd_src ? upd_src + m.oW1 : nullptr,
                                    upd_dst + m.oW1, dt, lr, L.s2)));
  prof_mark("gemm_axpy<false,false> A=s.dp", L.s2);
  launch_k(colsum_update_kernel, ceil_div(d, 16), 256, 0, st, 
      s.dp, s.df, s.dz, s.rhat, B, d, lr, upd_src ? upd_src + m.ob1 : nullptr, upd_dst + m.ob1,
      upd_src ? upd_src + m.ob2 : nullptr, upd_dst + m.ob2, upd_src ? upd_src + m.og : nullptr, upd_dst + m.og,
      upd_src ? upd_src + m.obt : nullptr, upd_dst + m.obt);
  prof_mark("colsum_update_kernel", st);
  return check_launch("forward_step");
}

// ---------------------------------------------------------------------------------------------------
// One reverse step: tangent of the first-order step along theta_dot = v (= a_{k+1}); writes a_k = v - lr H v
// and accumulates dlr, dscale, dY, dXn.
// ---------------------------------------------------------------------------------------------------
int run_pending_dy(const Dims& m, Work& w, const float* lr, float* dY, Lanes& L, bool full_grid = false);
int tangent_step(const Dims& m, Work& w, const Saved& s, const float* th, const float* v, float* a_out,
                 const float* lr, const float* scale, const float* mask, const int64_t* perm, float* dY, float* dlr,
                 float* dscale, Lanes& L, bool last = false) {
  // early operand loads (see forward_step): theta_k and the saved activations are steps old.  v = a_{k+1} is NOT old for the
  // first GEMM: its W1 block is written by the previous reverse step's last main-stream kernel (or, in the first reverse
  // step, by the matching-loss backward kernel), i.e. by the immediate predecessor; from the second GEMM on it is.
  const int kOldA = early_loads() ? 1 : 0, kOldB = early_loads() ? 2 : 0;
  cudaStream_t st = L.main;
  const int B = m.B, d = m.d, dt = m.dt, Bp = m.Bp;
  const size_t Bd = (size_t)B * d;
  const float *W1 = th + m.oW1, *W2 = th + m.oW2, *gam = th + m.og;
  const float *V1 = v + m.oW1, *c1 = v + m.ob1, *V2 = v + m.oW2, *c2 = v + m.ob2, *gamd = v + m.og, *betd = v + m.obt;
  // pd = Yb V1^T + c1 ; hd = gelu'(p) pd
  int sp = 1;
  CHECK_RC((gemm_partial<true, true>(gemm_ops(s.Yb, dt, V1, dt, B, d, dt), w.pa, &sp, st, kOldA)));
  prof_mark("gemm_partial<true,true> A=s.Yb", st);
  // The previous reverse step is not joined as a whole (its dY product would sit between its last weight GEMM and this step's
  // first one): the GEMM above needs only v[W1], written on the main stream.  From here on: c1 / c2 / gamd / betd come from
  // the previous step's column-sum kernel (side lane 1), V2 from its W2 GEMM (side lane 2), which also still reads w.hd.
  CHECK_RC(lane_wait(st, &L.ev_small));
  CHECK_RC(lane_wait(st, &L.ev_s2));
  launch_k(epi_pd_kernel, ew_grid4(Bd, d), 256, 0, st, w.pa, sp, Bd, c1, s.p, B, d, w.pd, w.hd);
  prof_mark("epi_pd_kernel", st);
  // fd = hd W2^T + h V2^T + c2 ; LN / normalise tangents
  CHECK_RC((gemm_partial<true, true>(gemm_ops2(w.hd, d, W2, d, d, s.h, d, V2, d, d, B, d), w.pa, &sp, st, kOldB)));
  prof_mark("gemm_partial<true,true> A=w.hd", st);
  if (row_v4_ok(d))
    VLDD_ROW_V4_DISPATCH(d, ln_tangent_v4_kernel, (B, 256, 0, st), (w.pa, sp, Bd, c2, mask, w.pd, s.rhat, s.rstd, s.yn, s.nz,
                                                                       gam, gamd, betd, d, w.rhatd, w.ynd, w.t, w.nzd));
  else
    launch_k(ln_tangent_kernel, B, 256, d * sizeof(float), st, w.pa, sp, Bd, c2, mask, w.pd, s.rhat, s.rstd, s.yn, s.nz, gam,
                                                         gamd, betd, d, w.rhatd, w.ynd, w.t, w.nzd);
  prof_mark("ln_tangent_kernel", st);
  if (L.dy.live) {
    // the previous step's dY product: from here to the end of the InfoNCE block the main lane keeps few SMs busy
    CHECK_RC(lane_edge(st, L.s1));
    CHECK_RC(run_pending_dy(m, w, lr, dY, L));
    CHECK_RC(lane_record(L.s1, &L.ev_s1_tail));
  }
  // Sd = scale Xb Ynd^T ; rho, kappa, Gd ; L_dot ; dlr, dscale ; dynd_raw[j,:] = sum_i Gd[i,j] Xb[i,:]
  const bool small = nce_small(B, d, Bp, s.Xb, w.ynd, w.pb);
  const bool fused_nce = !small && nce_fused(B, Bp);
  if (small) {
    launch_k(small_scores_kernel, dim3(ceil_div(B, kScoresTJ), ceil_div(B, kScoresTI)), kScoresThreads, scores_smem_bytes(d), st, (const float*)s.Xb, (const float*)w.ynd, B, d, scale,
             w.Sd, Bp);
    prof_mark("small_scores_kernel", st);
    launch_k(nce_t_gx_kernel, ceil_div(d, kNceGxCols), kNceGxThreads, nce_gx_smem_bytes(B, Bp), st, (const float*)s.S,
             (const float*)w.Sd, (const float*)s.Pr, (const float*)s.Pc, (const float*)s.Xb, B, Bp, d, lr, scale, w.Gd, dlr,
             dscale, w.pb);
    prof_mark("nce_t_gx_kernel", st);
  } else {
  CHECK_RC((gemm_partial<true, true>(gemm_ops(s.Xb, d, w.ynd, d, B, B, d), w.pa, &sp, st, kOldA)));
  prof_mark("gemm_partial<true,true> A=s.Xb", st);
  if (fused_nce) {
    launch_cluster_k(nce_t_cluster_kernel, kNceCluster, kNceThreads, nce_cluster_smem_bytes(B, Bp), st, w.pa, sp,
                     (size_t)B * B, scale, s.S, s.lse_r, s.lse_c, s.G, B, Bp, w.Gd, lr, dlr, dscale);
    prof_mark("nce_t_cluster_kernel", st);
  } else {
    launch_k(nce_t_rows_kernel, B, 128, 0, st, w.pa, sp, (size_t)B * B, scale, s.S, s.lse_r, s.G, B, Bp, w.Sd, w.rho, w.rowA,
             (const float*)nullptr);
    prof_mark("nce_t_rows_kernel", st);
    launch_k(nce_t_cols_kernel, B, 128, 0, st, s.S, s.lse_c, w.Sd, B, Bp, w.kap);
    prof_mark("nce_t_cols_kernel", st);
    launch_k(nce_t_grad_kernel, B, 128, 0, st, s.S, s.lse_r, s.lse_c, w.Sd, w.rho, w.kap, B, Bp, w.Gd, w.rowB);
    prof_mark("nce_t_grad_kernel", st);
  }
  }
  // branch 1: dXn_dot = scale (Gd Yn + G Ynd)  ->  dXn[perm] -= lr * scale * raw
  CHECK_RC(lane_edge(st, L.s1));
  L.s1_busy = true;
  CHECK_RC((gemm_store<true, false>(gemm_ops2(w.Gd, Bp, s.yn, d, B, s.G, Bp, w.ynd, d, B, B, d), w.pc, d, 1.0f, L.s1)));
  prof_mark("gemm_store<true,false> A=w.Gd", L.s1);
  launch_k(scatter_add_rows_kernel, B, 256, 0, L.s1, w.pc, 1, Bd, perm, d, lr, scale, w.dXn, m.N);
  prof_mark("scatter_add_rows_kernel", L.s1);
  if (!small && !fused_nce) {        // the dlr / dscale accumulation is off the critical path: side stream, ordered step to step
    launch_k(nce_t_finish_kernel, 1, 128, 0, L.s1, w.rowA, w.rowB, B, lr, scale, dlr, dscale);
    prof_mark("nce_t_finish_kernel", L.s1);
  }
  if (!small) {
    // dynd_raw[j,:] = sum_i Gd[i,j] Xb[i,:]
    CHECK_RC((gemm_store<false, false>(gemm_ops(w.Gd, Bp, s.Xb, d, Bp, d, B), w.pb, d, 1.0f, st, kOldB)));
    prof_mark("gemm_store<false,false> A=w.Gd", st);
  }
  if (row_v4_ok(d))
    VLDD_ROW_V4_DISPATCH(d, norm_ln_bwd_tangent_v4_kernel, (B, 256, 0, st), (w.pb, scale, s.yn, w.ynd, s.dyn, s.q, s.nz, w.nzd, s.dz, s.rhat, w.rhatd, s.rstd,
                                             w.t, s.dr, gam, gamd, mask, d, w.dzd, w.drd, w.dfd));
  else
    launch_k(norm_ln_bwd_tangent_kernel, B, 256, d * sizeof(float), st, w.pb, scale, s.yn, w.ynd, s.dyn, s.q, s.nz, w.nzd,
                                                                  s.dz, s.rhat, w.rhatd, s.rstd, w.t, s.dr, gam, gamd,
                                                                  mask, d, w.dzd, w.drd, w.dfd);
  prof_mark("norm_ln_bwd_tangent_kernel", st);
  // `last` (k = 0): a_0, the adjoint of the expert's start parameters, is nobody's input (the outputs are dY, dXn, dlr,
  // dscale) -- its two weight GEMMs and the column sums are not launched at all
  // branch 2: a_k[W2] = a_{k+1}[W2] - lr (dfd^T h + df^T hd)   (fused into the GEMM epilogue)
  if (!last) {
    CHECK_RC(lane_edge(st, L.s2));
    L.s2_busy = true;
    {
      tc::GridCapScope cap(dw2t_grid_cap());
      CHECK_RC((gemm_axpy<false, false>(gemm_ops2(w.dfd, d, s.h, d, B, s.df, d, w.hd, d, B, d, d), v + m.oW2, a_out + m.oW2, d,
                                        lr, L.s2)));
    }
    prof_mark("gemm_axpy<false,false> A=w.dfd", L.s2);
  }
  // dhd = dfd W2 + df V2 ; dpd
  CHECK_RC((gemm_partial<true, false>(gemm_ops2(w.dfd, d, W2, d, d, s.df, d, V2, d, d, B, d), w.pa, &sp, st, kOldB)));
  prof_mark("gemm_partial<true,false> A=w.dfd", st);
  // the previous step's dY product (side lane 1) reads w.dpd and v's buffer-mate a_out[W1]: both are rewritten from here on
  CHECK_RC(lane_wait(st, &L.ev_s1_tail));
  launch_k(epi_dpd_kernel, ew_grid4(Bd, d), 256, 0, st, w.pa, sp, Bd, s.p, w.pd, s.dh, w.drd, Bd, w.dpd);
  prof_mark("epi_dpd_kernel", st);
  // branch 1 (after dXn): small parameters of a_k by column sums (the next step's second kernel needs them), then
  // dY_dot = dpd W1 + dp V1  ->  dY[perm] -= lr * (.), which nothing needs before the end of the sweep
  CHECK_RC(lane_edge(st, L.s1));
  if (!last) {
    launch_k(colsum_tangent_update_kernel, ceil_div(d, 16), 256, 0, L.s1, 
        w.dpd, w.dfd, w.dzd, s.dz, s.rhat, w.rhatd, B, d, lr, v + m.ob1, a_out + m.ob1, v + m.ob2, a_out + m.ob2,
        v + m.og, a_out + m.og, v + m.obt, a_out + m.obt);
    prof_mark("colsum_tangent_update_kernel", L.s1);
    if (relaxed_joins()) CHECK_RC(lane_record(L.s1, &L.ev_small));
  }
  L.dy.live = true; L.dy.dpd = w.dpd; L.dy.W1 = W1; L.dy.dp = s.dp; L.dy.V1 = V1; L.dy.perm = perm;
  if (last || !(relaxed_joins() && defer_dy())) CHECK_RC(run_pending_dy(m, w, lr, dY, L, /*full_grid=*/last));
  if (last) return check_launch("tangent_step");            // the caller joins the lanes
  // main: a_k[W1] = a_{k+1}[W1] - lr dpd^T Yb
  CHECK_RC((gemm_axpy<false, false>(gemm_ops(w.dpd, d, s.Yb, dt, d, dt, B), v + m.oW1, a_out + m.oW1, dt, lr, st, kOldB)));
  prof_mark("gemm_axpy<false,false> A=w.dpd", st);
  if (relaxed_joins()) {
    // the next reverse step waits for exactly what it touches, where it touches it (see its head); the caller joins after
    // the last step
    if (!L.dy.live) CHECK_RC(lane_record(L.s1, &L.ev_s1_tail));
    CHECK_RC(lane_record(L.s2, &L.ev_s2));
  } else {
    CHECK_RC(lanes_join(L));          // a_k, dXn, dY complete before the next reverse step reuses the scratch buffers
  }
  return check_launch("tangent_step");
}

// dY_dot = dpd W1 + dp V1  ->  dY[perm] -= lr * (.) of the step recorded in L.dy, on side lane 1 (capped grid: it runs next
// to critical-path kernels)
int run_pending_dy(const Dims& m, Work& w, const float* lr, float* dY, Lanes& L, bool full_grid) {
  if (!L.dy.live) return VLDD_OK;
  L.dy.live = false;
  const int B = m.B, d = m.d, dt = m.dt;
  int sp_y = 1;
  {
    tc::GridCapScope cap(full_grid ? 0 : dy_grid_cap());
    CHECK_RC((gemm_partial<true, false>(gemm_ops2(L.dy.dpd, d, L.dy.W1, dt, d, L.dy.dp, d, L.dy.V1, dt, d, B, dt), w.pe, &sp_y, L.s1)));
  }
  prof_mark("gemm_partial<true,false> A=w.dpd", L.s1);
  launch_k(scatter_add_rows_kernel, B, 256, 0, L.s1, w.pe, sp_y, (size_t)B * dt, L.dy.perm, dt, lr, nullptr, dY, m.N);
  prof_mark("scatter_add_rows_kernel", L.s1);
  return VLDD_OK;
}

int validate(int N, int B, int K, int dt, int d) {
  VLDD_REQUIRE(N > 0 && B > 0 && B <= N, "need 0 < B <= N (got N=%d B=%d)", N, B);
  VLDD_REQUIRE(K >= 0 && K <= 64, "syn_steps K=%d out of range [0,64]", K);
  VLDD_REQUIRE(dt > 0 && d > 0, "bad dims dt=%d d=%d", dt, d);
  VLDD_REQUIRE((size_t)d * sizeof(float) <= 48 * 1024, "projection dim d=%d too large for the row kernels", d);
  return VLDD_OK;
}

}  // namespace

size_t unrolled_match_workspace_bytes(int N, int B, int K, int dt, int d) {
  if (validate(N, B, K, dt, d)) return 0;
  Work w;
  carve(w, make_dims(N, B, K, dt, d), nullptr);
  return w.bytes;
}

// ---------------------------------------------------------------------------------------------------
// Body of one call, everything after theta_0 / theta_tgt have been staged into the workspace.  All addresses it
// touches are workspace addresses or the caller's (Y, U, lr, scale, masks, outputs), so the launch sequence (~290 kernels) is
// captured ONCE per such address set into a CUDA graph and replayed afterwards; theta_0, theta* and the minibatch indices, which
// change every call, are reached through the workspace's pointer table (set_stage_sources).
// ---------------------------------------------------------------------------------------------------
static int unrolled_match_body(const Dims& m, Work& w, const float* Y, const float* U, const float* lr,
                               const float* scale, const int64_t* perms, const float* masks, float dropout_p,
                               unsigned long long* rng_state, float* out5, float* ce, float* dY, float* dU, float* theta_K,
                               cudaStream_t st) {
  const int N = m.N, B = m.B, K = m.K, dt = m.dt, d = m.d;
  const size_t Bd = (size_t)B * d;
  Lanes L;
  CHECK_RC(lanes_init(L, st));
  // Three independent branches open the call: the staging pass on the main lane (21 us of pure copy), the accumulator clears,
  // the row normalisation of U and the minibatch gather on side lane 1, the dropout masks on side lane 2; step 0 joins them.
  CHECK_RC(lane_edge(st, L.s1));
  CHECK_RC(lane_edge(st, L.s2));
  L.s1_busy = true;
  CHECK_RC(stage_segment_indirect(w.stage_table, w.traj, w.tgt, m.P, w.den, w.stage_scratch, st));
  {
    ZeroList zl;
    zl.st = L.s1;
    CHECK_RC(zl.add(w.ml_scratch, 16));
    CHECK_RC(zl.add(w.bad_index, sizeof(int)));
    CHECK_RC(zl.add(out5 + 3, 2 * sizeof(float)));
    CHECK_RC(zl.add(dY, (size_t)N * dt * sizeof(float)));
    CHECK_RC(zl.add(w.dXn, (size_t)N * d * sizeof(float)));
    CHECK_RC(zero_square_matrices(m, w, zl));
    CHECK_RC(zl.flush());
  }
  MARK("start");
  // fresh dropout masks for the K student steps (networks.py:636,643), drawn by the engine itself on a side branch; the
  // reverse sweep reads the same buffer, i.e. replays the same masks.  First use: the LayerNorm kernel of step 0.
  if (masks != nullptr && dropout_p > 0.f && rng_state != nullptr && K > 0) {
    L.s2_busy = true;
    CHECK_RC(dropout_masks(const_cast<float*>(masks), (int64_t)K * (int64_t)Bd, dropout_p, rng_state, 1, L.s2));
    prof_mark("dropout_masks", L.s2);
  }
  launch_k(row_normalise_kernel, N, 256, 0, L.s1, U, d, w.Xn, w.un);
  prof_mark("row_normalise", L.s1);
  // minibatches of all K steps in one launch (distill.py:510-513)
  if (K > 0) {
    const size_t step_stride = K > 1 ? (size_t)(w.sv[1].Yb - w.sv[0].Yb) : 0;
    launch_k(gather_all_kernel, dim3(B, K, 2), 256, 0, L.s1, Y, (const float*)w.Xn,
             reinterpret_cast<const int64_t* const*>(static_cast<char*>(w.stage_table) + 2 * sizeof(void*)), w.perms_copy, B, dt, d, w.sv[0].Yb, w.sv[0].Xb,
             step_stride, N, w.bad_index);
    prof_mark("gather_all", L.s1);
  }
  // forward unroll
  for (int k = 0; k < K; ++k) {
    Saved& s = w.sv[k];
    const float* th = w.traj + (size_t)k * m.P;
    CHECK_RC(forward_step(m, w, s, th, th, w.traj + (size_t)(k + 1) * m.P, lr, scale, masks ? masks + k * Bd : nullptr,
                          ce ? ce + k : nullptr, L, early_loads(), k));
  }
  CHECK_RC(lanes_join(L));
  const float* thK = w.traj + (size_t)K * m.P;
  // numerator + adjoint a_K = 2 (theta_K - theta*) / den in one pass (den was accumulated while the segment was staged)
  // (the pass leaves fp64 block partials of the numerator; nothing in the reverse sweep reads it, so the call's last kernel
  //  -- finalize_kernel below -- adds them up together with the index check: no serial tail in the streaming pass)
  CHECK_RC(match_final_pass(thK, w.tgt, w.den, m.P, w.adj0, w.ml_scratch, st));
  MARK("match_final");
  if (theta_K) VLDD_CUDA(cudaMemcpyAsync(theta_K, thK, m.P * sizeof(float), cudaMemcpyDeviceToDevice, st));
  // reverse sweep
  float* a_cur = w.adj0;
  float* a_nxt = w.adj1;
  for (int k = K - 1; k >= 0; --k) {
    const float* th = w.traj + (size_t)k * m.P;
    CHECK_RC(tangent_step(m, w, w.sv[k], th, a_cur, a_nxt, lr, scale, masks ? masks + k * Bd : nullptr,
                          w.perms_copy + (size_t)k * B, dY, out5 + 3, out5 + 4, L, /*last=*/k == 0));
    float* t = a_cur; a_cur = a_nxt; a_nxt = t;
  }
  if (L.dy.live) {                      // the last step's dY product
    CHECK_RC(lane_edge(st, L.s1));
    CHECK_RC(run_pending_dy(m, w, lr, dY, L));
  }
  CHECK_RC(lanes_join(L));
  launch_k(row_normalise_bwd_kernel, N, 256, 0, st, w.Xn, w.un, w.dXn, nullptr, d, dU);
  MARK("row_normalise_bwd");
  launch_k(finalize_kernel, 1, 256, 0, st, match_final_parts(w.ml_scratch), match_final_n_parts(m.P), (const float*)w.den,
           (const int*)w.bad_index, out5);
  prof_report();
  return check_launch("unrolled_match");
}

namespace {
struct GraphKey {
  const void* p[13];
  int dims[5];
  float dropout_p;
  bool operator==(const GraphKey& o) const { return memcmp(this, &o, sizeof(GraphKey)) == 0; }
};
struct GraphEntry { GraphKey key; cudaGraphExec_t exec; uint64_t last_use; unsigned long long kernels; };
std::mutex g_graph_mu;
std::vector<GraphEntry> g_graphs;
uint64_t g_graph_clock = 0;
constexpr size_t kMaxGraphs = 32;

bool graphs_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("VLDD_GRAPH");
    v = (e && strcmp(e, "0") == 0) ? 0 : 1;
  }
  return v == 1;
}
}  // namespace

int unrolled_match(const float* theta0, const float* theta_tgt, const float* Y, const float* U, const float* lr,
                   const float* scale, const int64_t* perms, const float* masks, float dropout_p,
                   unsigned long long* rng_state, int N, int B, int K, int dt, int d, float* out5, float* ce, float* dY,
                   float* dU, float* theta_K, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  CHECK_RC(validate(N, B, K, dt, d));
  const Dims m = make_dims(N, B, K, dt, d);
  Work w;
  carve(w, m, workspace);
  if (workspace == nullptr || workspace_bytes < w.bytes) {
    set_error("workspace too small: need %zu bytes, got %zu", w.bytes, workspace_bytes);
    return VLDD_ERR_WORKSPACE;
  }
  VLDD_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "dropout_p=%f out of range [0,1)", dropout_p);
  VLDD_REQUIRE(!(dropout_p > 0.f) || (masks != nullptr && rng_state != nullptr),
               "dropout_p > 0 needs a [K,B,d] mask buffer and an rng state {seed, draws}");
  // The segment is staged into the workspace by the FIRST kernel of the body (same pass: denominator |theta_0 - theta*|^2 of
  // the matching loss).  Its two source addresses change every iteration, so it reads them from a device table written here,
  // outside the replayed graph.  (As a separate launch ahead of the graph the 22 us pass was followed by the graph's 2 us root
  // kernel and ~10 us of launch latency before any second-level node ran -- profiles/timeline_r02z_steady.txt; as the root it
  // covers that latency.)
  CHECK_RC(set_stage_sources(theta0, theta_tgt, perms, w.stage_table, w.stage_scratch, st));
  if (!graphs_enabled() || prof_enabled()) {
    std::lock_guard<std::mutex> lock(g_graph_mu);   // the side streams / event pool are process-wide
    return unrolled_match_body(m, w, Y, U, lr, scale, perms, masks, dropout_p, rng_state, out5, ce, dY, dU, theta_K, st);
  }

  GraphKey key;
  memset(&key, 0, sizeof(key));
  // (theta_0, theta* and perms are not part of the key: the graph reaches them through the device table)
  const void* ptrs[13] = {workspace, Y, U, lr, scale, nullptr, masks, out5, ce, dY, dU, theta_K, rng_state};
  memcpy(key.p, ptrs, sizeof(ptrs));
  const int dims[5] = {N, B, K, dt, d};
  memcpy(key.dims, dims, sizeof(dims));
  key.dropout_p = dropout_p;
  std::lock_guard<std::mutex> lock(g_graph_mu);
  cudaGraphExec_t exec = nullptr;
  unsigned long long graph_kernels = 0;
  for (auto& e : g_graphs)
    if (e.key == key) { exec = e.exec; e.last_use = ++g_graph_clock; graph_kernels = e.kernels; break; }
  if (exec == nullptr) {
    DeviceState* ds = nullptr;
    CHECK_RC(current_device_state(&ds));
    if (ds->capture == nullptr) VLDD_CUDA(cudaStreamCreateWithFlags(&ds->capture, cudaStreamNonBlocking));
    cudaStream_t g_capture_stream = ds->capture;
    VLDD_CUDA(cudaStreamBeginCapture(g_capture_stream, cudaStreamCaptureModeThreadLocal));
    const unsigned long long before = launch_counter().load();
    const int rc = unrolled_match_body(m, w, Y, U, lr, scale, perms, masks, dropout_p, rng_state, out5, ce, dY, dU, theta_K,
                                       g_capture_stream);
    cudaGraph_t graph = nullptr;
    const cudaError_t ce_end = cudaStreamEndCapture(g_capture_stream, &graph);
    graph_kernels = launch_counter().exchange(before) - before;   // recorded, not run: counted per replay below
    if (rc != VLDD_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (ce_end != cudaSuccess || graph == nullptr) {
      set_error("CUDA graph capture failed: %s", cudaGetErrorString(ce_end));
      return VLDD_ERR_CUDA;
    }
    const cudaError_t ce_inst = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ce_inst != cudaSuccess) { set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(ce_inst)); return VLDD_ERR_CUDA; }
    if (g_graphs.size() >= kMaxGraphs) {
      size_t victim = 0;
      for (size_t i = 1; i < g_graphs.size(); ++i)
        if (g_graphs[i].last_use < g_graphs[victim].last_use) victim = i;
      cudaGraphExecDestroy(g_graphs[victim].exec);
      g_graphs.erase(g_graphs.begin() + victim);
    }
    g_graphs.push_back(GraphEntry{key, exec, ++g_graph_clock, graph_kernels});
  }
  VLDD_CUDA(cudaGraphLaunch(exec, st));
  launch_counter().fetch_add(graph_kernels, std::memory_order_relaxed);
  return VLDD_OK;
}

// First-order contrastive step on the whole batch (config 2): loss, g_theta, dY, dU, dscale.
size_t contrastive_step_workspace_bytes(int B, int dt, int d) { return unrolled_match_workspace_bytes(B, B, 1, dt, d); }

int contrastive_step(const float* theta, const float* Y, const float* U, const float* scale, const float* mask, int B,
                     int dt, int d, float* loss, float* g_theta, float* dY, float* dU, float* dscale, void* workspace,
                     size_t workspace_bytes, cudaStream_t st) {
  CHECK_RC(validate(B, B, 1, dt, d));
  const Dims m = make_dims(B, B, 1, dt, d);
  Work w;
  carve(w, m, workspace);
  if (workspace == nullptr || workspace_bytes < w.bytes) {
    set_error("workspace too small: need %zu bytes, got %zu", w.bytes, workspace_bytes);
    return VLDD_ERR_WORKSPACE;
  }
  Saved& s = w.sv[0];
  const size_t Bd = (size_t)B * d;
  launch_k(fill_kernel, 1, 32, 0, st, w.neg_one, -1.0f, 4);
  {
    ZeroList zl;
    zl.st = st;
    CHECK_RC(zero_square_matrices(m, w, zl));
    CHECK_RC(zl.flush());
  }
  launch_k(row_normalise_kernel, B, 256, 0, st, U, d, w.Xn, w.un);
  VLDD_CUDA(cudaMemcpyAsync(s.Yb, Y, (size_t)B * dt * sizeof(float), cudaMemcpyDeviceToDevice, st));
  VLDD_CUDA(cudaMemcpyAsync(s.Xb, w.Xn, Bd * sizeof(float), cudaMemcpyDeviceToDevice, st));
  std::lock_guard<std::mutex> lock(g_graph_mu);
  Lanes L;
  CHECK_RC(lanes_init(L, st));
  CHECK_RC(forward_step(m, w, s, theta, nullptr, g_theta, w.neg_one, scale, mask, loss, L));
  CHECK_RC(lanes_join(L));
  // dY = dp W1
  if (dY) {
    int sp = 1;
    CHECK_RC((gemm_partial<true, false>(gemm_ops(s.dp, d, theta + m.oW1, dt, B, dt, d), w.pa, &sp, st)));
    launch_k(reduce_slabs_kernel, ew_grid((size_t)B * dt), 256, 0, st, w.pa, sp, (size_t)B * dt, (size_t)B * dt, dY);
  }
  // dU = normalise_bwd(scale * G Yn)
  if (dU) {
    CHECK_RC((gemm_store<true, false>(gemm_ops(s.G, m.Bp, s.yn, d, B, d, B), w.pb, d, 1.0f, st)));
    launch_k(row_normalise_bwd_kernel, B, 256, 0, st, w.Xn, w.un, w.pb, scale, d, dU);
  }
  // dscale = sum(G * S) / scale
  if (dscale) launch_k(dot_over_scale_kernel, 1, 256, 0, st, s.G, s.S, (size_t)B * m.Bp, scale, dscale);
  return check_launch("contrastive_step");
}

// Top-1 retrieval hits inside the batch, both directions (networks.py:884-886: argmax over rows / columns of the logits
// equals the diagonal index; torch.argmax returns the FIRST maximum, so a tie counts only for the lowest index).
//   top1[0] = #{i : argmax_j S_ij == i}     top1[1] = #{j : argmax_i S_ij == j}
__global__ void __launch_bounds__(128) nce_top1_kernel(const float* __restrict__ S, int B, int ld, int32_t* __restrict__ top1) {
  pdl_enter();
  __shared__ int scratch[34];
  const int i = blockIdx.x;
  const float sii = S[(size_t)i * ld + i];
  int ahead_r = 0, ahead_c = 0;
  for (int j = threadIdx.x; j < B; j += blockDim.x) {
    const float r = S[(size_t)i * ld + j], c = S[(size_t)j * ld + i];
    ahead_r += (r > sii) || (r == sii && j < i);
    ahead_c += (c > sii) || (c == sii && j < i);
  }
  ahead_r = block_sum<int>(ahead_r, scratch);
  ahead_c = block_sum<int>(ahead_c, scratch);
  if (threadIdx.x == 0) {
    if (ahead_r == 0) atomicAdd(top1 + 0, 1);
    if (ahead_c == 0) atomicAdd(top1 + 1, 1);
  }
}

// CLIPModel_full.forward from the encoder outputs on (networks.py:866-889): text head, normalise, logits, symmetric
// cross-entropy, top-1 counters -- plus the first-order gradients (what loss.backward() hands the two optimisers).
int clip_loss(const float* theta, const float* Y, const float* U, const float* scale, const float* mask, int B, int dt,
              int d, float* loss, int32_t* top1, float* g_theta, float* dY, float* dU, float* dscale, void* workspace,
              size_t workspace_bytes, cudaStream_t st) {
  CHECK_RC(contrastive_step(theta, Y, U, scale, mask, B, dt, d, loss, g_theta, dY, dU, dscale, workspace, workspace_bytes, st));
  if (top1 != nullptr) {
    const Dims m = make_dims(B, B, 1, dt, d);
    Work w;
    carve(w, m, workspace);
    VLDD_CUDA(cudaMemsetAsync(top1, 0, 2 * sizeof(int32_t), st));
    launch_k(nce_top1_kernel, B, 128, 0, st, (const float*)w.sv[0].S, B, m.Bp, top1);
  }
  return check_launch("clip_loss");
}

// ---------------------------------------------------------------------------------------------------
// "Mode B" building block: the bidirectional InfoNCE loss on already-normalised features as a node that PyTorch can
// differentiate TWICE (distill.py:548-551 under autograd.grad(create_graph=True) with an arbitrary image tower, e.g.
// pixels -> NFNet under ReparamModule, distill.py:524-567).  Two stateless entry points:
//   infonce_grad: L, dL/dxn = s G yn, dL/dyn = s G^T xn, dL/ds = sum(G o S) / s
//   infonce_hvp : for a direction (cx, cy, cs) -- the cotangents autograd hands to the first-order gradients --
//                 Ldot = <grad L, direction> and the gradient of Ldot w.r.t. (xn, yn, s):
//                   Sd = cs S/s + s (cx yn^T + xn cy^T) ; Gd = Hess_S(L) Sd (nce_t_* kernels)
//                   hx = s (Gd yn + G cy) + cs G yn ; hy = s (Gd^T xn + G^T cx) + cs G^T xn
//                   hs = (sum Gd o S + sum G o Sd) / s - cs sum(G o S) / s^2
// ---------------------------------------------------------------------------------------------------
struct NceWork {
  float *S, *G, *Sd, *Gd;                        // [B, Bp]
  float *lse_r, *lse_c, *rho, *kap, *rowA, *rowB, *rowC;   // [B]
  float *part;                                   // split-K slabs of the B x B GEMMs
  float *t1, *t2;                                // [Bp, d] GEMM outputs
  float *scal;                                   // 8 scalars
  size_t bytes;
};
static void carve_nce(NceWork& w, int B, int d, void* base) {
  Bump b{reinterpret_cast<char*>(base), 0, 0};
  const int Bp = (B + 31) / 32 * 32;
  const size_t BB = (size_t)B * Bp;
  w.S = b.f(BB); w.G = b.f(BB); w.Sd = b.f(BB); w.Gd = b.f(BB);
  w.lse_r = b.f(B); w.lse_c = b.f(B); w.rho = b.f(B); w.kap = b.f(B); w.rowA = b.f(B); w.rowB = b.f(B); w.rowC = b.f(B);
  w.part = b.f((size_t)max_splits(B, B, 2 * d) * B * B);
  w.t1 = b.f((size_t)Bp * d); w.t2 = b.f((size_t)Bp * d);
  w.scal = b.f(8);
  w.bytes = b.off;
}
size_t infonce_workspace_bytes(int B, int d) {
  if (B <= 0 || d <= 0) return 0;
  NceWork w;
  carve_nce(w, B, d, nullptr);
  return w.bytes;
}
// out = (*a) * x + (b ? (*b) * y : 0)
__global__ void __launch_bounds__(256) lincomb_kernel(const float* __restrict__ x, const float* __restrict__ a,
                                                      const float* __restrict__ y, const float* __restrict__ b, size_t n,
                                                      float* __restrict__ out) {
  pdl_enter();
  const float ca = *a, cb = b ? *b : 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = y ? fmaf(cb, y[i], ca * x[i]) : ca * x[i];
}
// rowC[i] = sum_j G_ij S_ij
__global__ void __launch_bounds__(128) rowdot_kernel(const float* __restrict__ G, const float* __restrict__ S, int B, int ld,
                                                     float* __restrict__ out) {
  pdl_enter();
  __shared__ float scratch[34];
  float a = 0.f;
  for (int j = threadIdx.x; j < B; j += blockDim.x) a = fmaf(G[(size_t)blockIdx.x * ld + j], S[(size_t)blockIdx.x * ld + j], a);
  a = block_sum<float>(a, scratch);
  if (threadIdx.x == 0) out[blockIdx.x] = a;
}
// Ldot = sum rowA ; hs = (sum rowB + sum rowA) / s - cs * sum rowC / s^2
__global__ void __launch_bounds__(128) nce_hvp_finish_kernel(const float* __restrict__ rowA, const float* __restrict__ rowB,
                                                             const float* __restrict__ rowC, int B,
                                                             const float* __restrict__ scale, const float* __restrict__ cs,
                                                             float* __restrict__ Ldot, float* __restrict__ hs) {
  pdl_enter();
  __shared__ float scratch[34];
  float a = 0.f, b = 0.f, c = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) { a += rowA[i]; b += rowB[i]; c += rowC[i]; }
  a = block_sum<float>(a, scratch);
  b = block_sum<float>(b, scratch);
  c = block_sum<float>(c, scratch);
  if (threadIdx.x == 0) {
    const float s = *scale;
    *Ldot = a;
    *hs = (b + a) / s - (*cs) * c / (s * s);
  }
}

// S, lse, G (and the loss) from normalised features: shared first half of both entry points
static int nce_primal(NceWork& w, const float* xn, const float* yn, const float* scale, int B, int d, float* loss,
                      cudaStream_t st) {
  const int Bp = (B + 31) / 32 * 32;
  if (Bp != B) {
    const size_t bytes = (size_t)B * Bp * sizeof(float);
    VLDD_CUDA(cudaMemsetAsync(w.S, 0, bytes, st));
    VLDD_CUDA(cudaMemsetAsync(w.G, 0, bytes, st));
    VLDD_CUDA(cudaMemsetAsync(w.Sd, 0, bytes, st));
    VLDD_CUDA(cudaMemsetAsync(w.Gd, 0, bytes, st));
  }
  int sp = 1;
  CHECK_RC((gemm_partial<true, true>(gemm_ops(xn, d, yn, d, B, B, d), w.part, &sp, st)));
  launch_k(nce_rows_kernel, B, 128, 0, st, w.part, sp, (size_t)B * B, scale, B, Bp, w.S, w.lse_r);
  launch_k(nce_cols_kernel, B, 128, 0, st, w.S, B, Bp, w.lse_c);
  launch_k(nce_grad_kernel, B, 128, 0, st, w.S, w.lse_r, w.lse_c, B, Bp, w.G, loss);
  return VLDD_OK;
}

int infonce_grad(const float* xn, const float* yn, const float* scale, int B, int d, float* loss, float* dxn, float* dyn,
                 float* dscale, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  VLDD_REQUIRE(B > 0 && d > 0, "infonce: bad dims B=%d d=%d", B, d);
  NceWork w;
  carve_nce(w, B, d, workspace);
  if (workspace == nullptr || workspace_bytes < w.bytes) {
    set_error("workspace too small: need %zu bytes, got %zu", w.bytes, workspace_bytes);
    return VLDD_ERR_WORKSPACE;
  }
  const int Bp = (B + 31) / 32 * 32;
  const size_t Bd = (size_t)B * d;
  CHECK_RC(nce_primal(w, xn, yn, scale, B, d, loss, st));
  if (dxn) {
    CHECK_RC((gemm_store<true, false>(gemm_ops(w.G, Bp, yn, d, B, d, B), w.t1, d, 1.0f, st)));
    launch_k(lincomb_kernel, ew_grid(Bd), 256, 0, st, (const float*)w.t1, scale, (const float*)nullptr, (const float*)nullptr, Bd, dxn);
  }
  if (dyn) {
    CHECK_RC((gemm_store<false, false>(gemm_ops(w.G, Bp, xn, d, Bp, d, B), w.t2, d, 1.0f, st)));
    launch_k(lincomb_kernel, ew_grid(Bd), 256, 0, st, (const float*)w.t2, scale, (const float*)nullptr, (const float*)nullptr, Bd, dyn);
  }
  if (dscale) launch_k(dot_over_scale_kernel, 1, 256, 0, st, (const float*)w.G, (const float*)w.S, (size_t)B * Bp, scale, dscale);
  return check_launch("infonce_grad");
}

int infonce_hvp(const float* xn, const float* yn, const float* scale, const float* cx, const float* cy, const float* cs,
                int B, int d, float* Ldot, float* hx, float* hy, float* hs, void* workspace, size_t workspace_bytes,
                cudaStream_t st) {
  VLDD_REQUIRE(B > 0 && d > 0, "infonce_hvp: bad dims B=%d d=%d", B, d);
  NceWork w;
  carve_nce(w, B, d, workspace);
  if (workspace == nullptr || workspace_bytes < w.bytes) {
    set_error("workspace too small: need %zu bytes, got %zu", w.bytes, workspace_bytes);
    return VLDD_ERR_WORKSPACE;
  }
  const int Bp = (B + 31) / 32 * 32;
  const size_t Bd = (size_t)B * d;
  CHECK_RC(nce_primal(w, xn, yn, scale, B, d, nullptr, st));
  // Sd = cs S/s + s (cx yn^T + xn cy^T)
  int sp = 1;
  CHECK_RC((gemm_partial<true, true>(gemm_ops2(cx, d, yn, d, d, xn, d, cy, d, d, B, B), w.part, &sp, st)));
  launch_k(nce_t_rows_kernel, B, 128, 0, st, w.part, sp, (size_t)B * B, scale, w.S, w.lse_r, w.G, B, Bp, w.Sd, w.rho, w.rowA, cs);
  launch_k(nce_t_cols_kernel, B, 128, 0, st, w.S, w.lse_c, w.Sd, B, Bp, w.kap);
  launch_k(nce_t_grad_kernel, B, 128, 0, st, w.S, w.lse_r, w.lse_c, w.Sd, w.rho, w.kap, B, Bp, w.Gd, w.rowB);
  launch_k(rowdot_kernel, B, 128, 0, st, (const float*)w.G, (const float*)w.S, B, Bp, w.rowC);
  launch_k(nce_hvp_finish_kernel, 1, 128, 0, st, (const float*)w.rowA, (const float*)w.rowB, (const float*)w.rowC, B, scale, cs,
           Ldot, hs);
  // hx = s (Gd yn + G cy) + cs (G yn)
  CHECK_RC((gemm_store<true, false>(gemm_ops2(w.Gd, Bp, yn, d, B, w.G, Bp, cy, d, B, B, d), w.t1, d, 1.0f, st)));
  CHECK_RC((gemm_store<true, false>(gemm_ops(w.G, Bp, yn, d, B, d, B), w.t2, d, 1.0f, st)));
  launch_k(lincomb_kernel, ew_grid(Bd), 256, 0, st, (const float*)w.t1, scale, (const float*)w.t2, cs, Bd, hx);
  // hy = s (Gd^T xn + G^T cx) + cs (G^T xn)
  CHECK_RC((gemm_store<false, false>(gemm_ops2(w.Gd, Bp, xn, d, B, w.G, Bp, cx, d, B, Bp, d), w.t1, d, 1.0f, st)));
  CHECK_RC((gemm_store<false, false>(gemm_ops(w.G, Bp, xn, d, Bp, d, B), w.t2, d, 1.0f, st)));
  launch_k(lincomb_kernel, ew_grid(Bd), 256, 0, st, (const float*)w.t1, scale, (const float*)w.t2, cs, Bd, hy);
  return check_launch("infonce_hvp");
}

// text_projection forward over `rows` embeddings (eval-mode when mask == nullptr): z = LN(...), zn = z/|z|.
//   reference: epoch_original.py:77-78 (text head over the cached BERT test embeddings, then normalise)
size_t proj_head_workspace_bytes(int rows, int dt, int d) {
  if (rows <= 0 || dt <= 0 || d <= 0) return 0;
  const size_t Rd = (size_t)rows * d;
  const size_t s1 = (size_t)max_splits(rows, d, dt) * Rd, s2 = (size_t)max_splits(rows, d, d) * Rd;
  const size_t part = s1 > s2 ? s1 : s2;
  return (part + 2 * Rd) * sizeof(float) + 1024;
}

int proj_head_forward(const float* theta, const float* Y, const float* mask, int rows, int dt, int d, float* z,
                      float* zn, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  VLDD_REQUIRE(rows > 0 && dt > 0 && d > 0, "bad dims rows=%d dt=%d d=%d", rows, dt, d);
  VLDD_REQUIRE((size_t)d * sizeof(float) <= 48 * 1024, "projection dim d=%d too large for the row kernels", d);
  const size_t need = proj_head_workspace_bytes(rows, dt, d);
  if (workspace == nullptr || workspace_bytes < need) {
    set_error("workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
    return VLDD_ERR_WORKSPACE;
  }
  const Dims m = make_dims(rows, rows, 0, dt, d);
  const size_t Rd = (size_t)rows * d;
  float* p = reinterpret_cast<float*>(workspace);
  float* h = p + Rd;
  float* part = h + Rd;
  int sp = 1;
  CHECK_RC((gemm_partial<true, true>(gemm_ops(Y, dt, theta + m.oW1, dt, rows, d, dt), part, &sp, st)));
  launch_k(epi_p_kernel, ew_grid4(Rd, d), 256, 0, st, part, sp, Rd, theta + m.ob1, rows, d, p, h);
  CHECK_RC((gemm_partial<true, true>(gemm_ops(h, d, theta + m.oW2, d, rows, d, d), part, &sp, st)));
  if (row_v4_ok(d))
    VLDD_ROW_V4_DISPATCH(d, ln_fwd_v4_kernel, (rows, 256, 0, st), (part, sp, Rd, theta + m.ob2, mask, p, theta + m.og,
                                                                      theta + m.obt, d, nullptr, z, zn, nullptr, nullptr));
  else
    launch_k(ln_fwd_kernel, rows, 256, d * sizeof(float), st, part, sp, Rd, theta + m.ob2, mask, p, theta + m.og,
                                                        theta + m.obt, d, nullptr, z, zn, nullptr, nullptr);
  return check_launch("proj_head_forward");
}

// Measurement hook (bench.py roofline): exactly the weight-streaming GEMM launch the engine issues for f = h W2^T
// (split-K partial slabs, 3xTF32 tcgen05 kernel), on caller-provided operands.
size_t skinny_gemm_workspace_bytes(int M, int N, int K) { return (size_t)max_splits(M, N, K) * M * N * sizeof(float); }

int skinny_gemm_partial(const float* A, const float* W, int M, int N, int K, float* partial, size_t partial_bytes,
                        int* splits, cudaStream_t st) {
  VLDD_REQUIRE(A && W && partial && splits && M > 0 && N > 0 && K > 0, "skinny_gemm_partial: bad arguments");
  if (partial_bytes < skinny_gemm_workspace_bytes(M, N, K)) {
    set_error("skinny_gemm_partial: workspace too small");
    return VLDD_ERR_WORKSPACE;
  }
  CHECK_RC((gemm_partial<true, true>(gemm_ops(A, K, W, K, M, N, K), partial, splits, st)));
  return check_launch("skinny_gemm_partial");
}

}  // namespace vldd
