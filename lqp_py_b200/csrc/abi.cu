// C ABI (include/lqpb.h): argument validation, workspace carving, kernel orchestration, profiling.
// Host code only; every kernel lives in scale.cu / factor.cu / iterate.cu / backward.cu / lu.cu.
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <string>
#include <type_traits>
#include "layout.cuh"

namespace lqpb {
template <typename T>
cudaError_t launch_lu_factor(int B, int N, const T* A, T* LU, int32_t* piv, cudaStream_t st);
template <typename T>
cudaError_t launch_lu_solve(int B, int N, int nrhs, const T* LU, const int32_t* piv, const T* rhs, T* x, int negate,
                            cudaStream_t st);
template <typename T>
cudaError_t launch_outer(int B, int N, int M, const T* a, const T* b, T* C, cudaStream_t st);
cudaError_t launch_mapped_copy(void* dst, const void* src, size_t bytes, cudaStream_t st);   // hostio.cu
template <typename T>
int iterate_split_size(const FwdWs<T>& w);                                                    // iterate_split.cu
}  // namespace lqpb

using namespace lqpb;

namespace {

thread_local std::string g_err;
bool g_prof_on = false;   // process-wide switch: autograd runs the backward on its own thread
constexpr int kMaxDev = 64;

constexpr int kMaxSeg = 16;
struct ProfState {
  bool have_events = false;
  cudaEvent_t ev[8 + 2 * kMaxSeg + 2 * kMaxSeg];
  // forward: 0 start, 1 after scale, [factor segs], [iterate segs], fin0, fin1 ; backward: 4..7
  int n_fac = 0, n_it = 0;
  bool fwd_valid = false, bwd_valid = false;
  int launches = 0, it_launches = 0, fac_launches = 0;
  int prep_launches = 0;     // kernels a forward call enqueued for the backward (stage 1)
  cudaEvent_t fac0[kMaxSeg], fac1[kMaxSeg], it0[kMaxSeg], it1[kMaxSeg];
};
// Profiling state (only touched while lqpb_profile_enable(1) is in force): one record per DEVICE, because CUDA events
// belong to the device they were created on; calls made with profiling off write to a per-thread scratch record, so
// the solver itself keeps no process-global state.
ProfState g_prof_tab[kMaxDev];
thread_local ProfState g_prof_scratch;

int cur_dev() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= kMaxDev) d = 0;
  return d;
}
ProfState& prof_state() { return g_prof_on ? g_prof_tab[cur_dev()] : g_prof_scratch; }

void prof_init(ProfState& g_prof) {
  if (g_prof.have_events) return;
  for (auto& e : g_prof.ev) cudaEventCreate(&e);
  for (int i = 0; i < kMaxSeg; ++i) {
    cudaEventCreate(&g_prof.fac0[i]);
    cudaEventCreate(&g_prof.fac1[i]);
    cudaEventCreate(&g_prof.it0[i]);
    cudaEventCreate(&g_prof.it1[i]);
  }
  g_prof.have_events = true;
}

int fail(int code, const char* what, cudaError_t ce = cudaSuccess) {
  g_err = what;
  if (ce != cudaSuccess) {
    g_err += ": ";
    g_err += cudaGetErrorString(ce);
  }
  return code;
}

#define CK(call, what)                                          \
  do {                                                          \
    cudaError_t ce__ = (call);                                  \
    if (ce__ != cudaSuccess) return fail(LQPB_E_CUDA, what, ce__); \
  } while (0)

int check_device() {
  static thread_local int ok_dev = -1;
  int dev = 0;
  cudaError_t ce = cudaGetDevice(&dev);
  if (ce != cudaSuccess) return fail(LQPB_E_CUDA, "cudaGetDevice", ce);
  if (dev == ok_dev) return LQPB_OK;
  int major = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10) return fail(LQPB_E_NOT_BLACKWELL, "lqpb kernels are built for sm_100a (B200) only");
  ok_dev = dev;
  return LQPB_OK;
}

struct HostCtrl {
  Ctrl* pinned = nullptr;                     // page-locked host memory: usable from every device (UVA)
  cudaEvent_t seg_done[kMaxDev] = {};         // events belong to a device: one per device, created on first use
  ~HostCtrl() {}
};
thread_local HostCtrl g_hctrl;

// ---- host-buffer calls (lqpb_forward_host_* / lqpb_backward_host_*): the batch is cut into chunks of whole
// problems; a copy stream moves chunk c + 1 over PCIe while the compute stream runs the per-problem setup
// (forward) or the whole adjoint chain (backward) of chunk c.
constexpr int kMaxChunks = 16;
struct HostPipe {
  int dev = -1;
  cudaStream_t cs;
  cudaEvent_t fork, vec, done, ev[kMaxChunks];
};
thread_local HostPipe g_pipe_tab[kMaxDev];      // streams and events belong to a device: one set per device ordinal
#define g_pipe (g_pipe_tab[cur_dev()])

cudaError_t pipe_init() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess || g_pipe.dev == dev) return e;
  if ((e = cudaStreamCreateWithFlags(&g_pipe.cs, cudaStreamNonBlocking)) != cudaSuccess) return e;
  cudaEvent_t* evs[] = {&g_pipe.fork, &g_pipe.vec, &g_pipe.done};
  for (auto p : evs)
    if ((e = cudaEventCreateWithFlags(p, cudaEventDisableTiming)) != cudaSuccess) return e;
  for (auto& p : g_pipe.ev)
    if ((e = cudaEventCreateWithFlags(&p, cudaEventDisableTiming)) != cudaSuccess) return e;
  g_pipe.dev = dev;
  return cudaSuccess;
}

int pick_chunks(int requested, int B) {
  static const int env_chunks = [] { const char* e = getenv("LQPB_HOST_CHUNKS"); return e ? atoi(e) : 0; }();
  int c = requested > 0 ? requested : (env_chunks > 0 ? env_chunks : 4);
  if (c > kMaxChunks) c = kMaxChunks;
  while (c > 1 && B / c < 16) --c;       // chunks of fewer than 16 problems are pure launch latency
  return c < 1 ? 1 : c;
}

// chunk c of C covers problems [chunk_lo(c), chunk_lo(c + 1)): equal sizes (the per-chunk kernel chains are latency-
// bound, ~0.45 ms for 32 problems at dz=500; smaller first / last chunks were measured and did not pay)
int chunk_lo(int B, int C, int c) { return (int)((long long)B * c / C); }

template <typename T>
struct HostFwd {      // host side of a forward call
  const T *Q, *p, *A, *b, *lb, *ub;
  T* x;
  int chunks;
};
template <typename T>
struct HostBwd {      // host side of a backward call
  const T* dl_dz;
  T *dQ, *dp, *dA, *db, *dlb, *dub;
  int chunks;
};

// problems [b0, b0 + bc) of a carved workspace (every array is per-problem contiguous; ctrl / flags are shared)
template <typename T>
FwdWs<T> slice_fwd(const FwdWs<T>& w, int b0, int bc) {
  FwdWs<T> s = w;
  s.B = bc;
  const size_t o = (size_t)b0, te = (size_t)kTcBlock * kTcBlock, mm = w.m > 0 ? w.m : 1;
  s.Qp += o * Pack<T>::elems(w.n);
  s.Kp += o * Pack<T>::elems(w.n);
  if (w.tc) {
    s.W += o * ((size_t)w.nb * (w.nb + 1) / 2) * te;
    s.Vg += o * w.nb * te;
    s.Wg += o * w.nb * te;
    s.Pb += o * w.nb * te;
  } else {
    s.W += o * w.np * w.np;
    s.Vg += o * w.np * kTile;
    s.Wg += o * w.np * kTile;
  }
  T** vecs[] = {&s.D, &s.pt, &s.lbt, &s.ubt, &s.c, &s.z, &s.u, &s.xs};
  for (auto v : vecs) *v += o * w.ld;
  s.At += o * mm * w.ld;
  s.Gt += o * mm * w.ld;
  s.Sinv += o * mm * mm;
  s.bt += o * mm;
  s.E += o * mm;
  s.rho += o; s.rho_cand += o; s.pnorm += o; s.ratio += o;
  s.fro_part += o * w.n_fro;
  s.chk += 4 * o;
  s.wants += o;
  if (w.z0) s.z0 += o * w.n;
  if (w.u0) s.u0 += o * w.n;
  if (w.rho_in) s.rho_in += o;
  return s;
}
template <typename T>
BwdWs<T> slice_bwd(const BwdWs<T>& w, int b0, int bc) {
  BwdWs<T> s = w;
  s.B = bc;
  const size_t o = (size_t)b0, te = (size_t)kTcBlock * kTcBlock;
  if (w.tc) {
    s.W += o * ((size_t)w.nb * (w.nb + 1) / 2) * te;
    s.Vg += o * w.nb * te;
    s.Wg += o * w.nb * te;
    s.Pb += o * w.nb * te;
  } else {
    s.W += o * w.np * w.np;
    s.Vg += o * w.np * kTile;
    s.Wg += o * w.np * kTile;
  }
  s.mask += o * w.ld;
  s.dv += o * w.ld;
  s.dvec += o * w.ld;
  s.dnu += o * (w.m > 0 ? w.m : 1);
  return s;
}
template <typename P>
P* off(P* p, size_t elems) { return p ? p + elems : nullptr; }

// ---- unrolled mode: what a reverse sweep needs of one operator segment (rho and the KKT inverse it belongs to)
template <typename T>
struct Snap {
  T *Kp, *Gt, *Sinv, *c, *rho;
  size_t bytes;
};
template <typename T>
Snap<T> carve_snap(void* base, int B, int n, int m) {
  Snap<T> s;
  char* p = static_cast<char*>(base);
  size_t o = 0;
  auto take = [&](size_t count) {
    void* r = p ? p + o : nullptr;
    o = round_up_sz(o + count * sizeof(T), 256);
    return (T*)r;
  };
  const size_t Bn = (size_t)B, mm = m > 0 ? m : 1, ld = round_up(n, Vec<T>::N);
  s.Kp = take(Bn * Pack<T>::elems(n));
  s.Gt = take(Bn * mm * ld);
  s.Sinv = take(Bn * mm * mm);
  s.c = take(Bn * ld);
  s.rho = take(Bn);
  s.bytes = o;
  return s;
}
template <typename T>
struct UnrollRec {        // recording run of a forward solve (lqpb_unroll_forward_*)
  Tape<T> tape;
  void* snap;             // max_seg snapshots of carve_snap().bytes each
  int max_seg;
  int32_t* seg_start;     // host, max_seg + 1 entries: first iteration of every operator segment, then n_iter
  int32_t* wants;         // device, (max_seg - 1, B): do_rho_update flags applied by each adaptive-rho update
  int true_max_iters;     // first-pass mode (snap == nullptr): cfg->max_iters before it was cut to the tape capacity
};

// forward call that also enqueues the dl_dz-independent part of the backward (stage 1 of backward_impl)
template <typename T>
struct BwdPrep {
  void* ws;
  size_t ws_bytes;
  int kkt;
};
template <typename T>
int backward_impl(int B, int n, int m, const T* dl_dz, const T* x, const T* u, const T* lams, const T* nus,
                  const T* Q, const T* A, const T* lb, const T* ub, const T* rho_dev, double rho_scalar, T* dQ,
                  T* dp, T* dA, T* db, T* dlb, T* dub, void* ws, size_t ws_bytes, void* stream, bool kkt = false,
                  int32_t* any_bounds = nullptr, const HostBwd<T>* host = nullptr, int stage = 0);

template <typename T>
int factor_forward(const FwdWs<T>& w, bool first, cudaStream_t st) {
  GjArgs<T> a{};
  a.n = w.n; a.m = w.m; a.np = w.np;
  a.src = w.Qp; a.lds = 0;          // packed symmetric Q~ (scale.cu)
  a.diag_shift = w.rho; a.diag_const = T(0);
  a.mask = nullptr; a.ldm = 0;
  a.Arows = w.At; a.lda = w.ld; a.a_diag = T(0);
  a.W = w.W; a.Vg = w.Vg; a.Wg = w.Wg;
  a.dst = w.Kp; a.ldd = w.ld; a.G21 = w.Gt; a.K22 = w.Sinv;
  a.bt = w.bt; a.c_out = w.m > 0 ? w.c : nullptr;
  ProfState& g_prof = prof_state();
  if (w.tc) {
    int l = 0;
    // the first factorisation of a call finds the H block already in place (scale_pack_kernel)
    CK(launch_tc_inverse(w.B, a, w.Pb, w.nb, first, st, &l), "tensor-core inverse (forward)");
    g_prof.launches += l;
    g_prof.fac_launches += l;
    return LQPB_OK;
  }
  CK(launch_gj_inverse<T>(w.B, a, st), "gj_inverse (forward)");
  g_prof.launches += 1;
  g_prof.fac_launches += 1;
  return LQPB_OK;
}

template <typename T>
int forward_impl(const lqpb_config* cfg, int B, int n, int m, const T* Q, const T* p, const T* A, const T* b,
                 const T* lb, const T* ub, T* x, T* z, T* u, T* lams, T* nus, T* rho_out, lqpb_info* info, void* ws,
                 size_t ws_bytes, void* stream, const HostFwd<T>* host = nullptr, const UnrollRec<T>* rec = nullptr,
                 const BwdPrep<T>* prep = nullptr, const T* z0 = nullptr, const T* u0 = nullptr,
                 void* async_ctrl = nullptr, int32_t* deferred = nullptr, const T* rho_in = nullptr) {
  if (host && (!host->Q || !host->p || !host->lb || !host->ub || (m > 0 && (!host->A || !host->b))))
    return fail(LQPB_E_ARG, "null host pointer argument");
  if (!cfg || !Q || !p || !lb || !ub || !x || !z || !u || !lams || !rho_out || !info || !ws)
    return fail(LQPB_E_ARG, "null pointer argument");
  if (B <= 0 || n <= 0 || m < 0) return fail(LQPB_E_ARG, "bad dimensions");
  if (m > 0 && (!A || !b || !nus)) return fail(LQPB_E_ARG, "A, b and nus are required when m > 0");
  if (m > kMaxM) return fail(LQPB_E_ARG, "more than 256 equality rows are not supported");
  if ((z0 == nullptr) != (u0 == nullptr)) return fail(LQPB_E_ARG, "a warm start needs both z0 and u0");
  if (cfg->max_iters < 1 || cfg->check_solved < 1 || cfg->adaptive_rho_iter < 1)
    return fail(LQPB_E_ARG, "max_iters, check_solved and adaptive_rho_iter must be >= 1");
  int rc = check_device();
  if (rc) return rc;
  FwdWs<T> w = carve_fwd<T>(ws, B, n, m);
  if (w.bytes > ws_bytes) return fail(LQPB_E_WORKSPACE, "workspace too small");
  w.z0 = z0;
  w.u0 = u0;
  w.rho_in = rho_in;
  cudaStream_t st = (cudaStream_t)stream;
  if (!g_hctrl.pinned) CK(cudaMallocHost(&g_hctrl.pinned, sizeof(Ctrl)), "cudaMallocHost");
  Ctrl* hc = g_hctrl.pinned;

  const bool prof = g_prof_on;
  ProfState& g_prof = prof_state();
  if (prof) prof_init(g_prof);
  g_prof.launches = g_prof.it_launches = g_prof.fac_launches = 0;
  g_prof.n_fac = g_prof.n_it = 0;
  g_prof.fwd_valid = false;

  CK(cudaMemsetAsync(w.ctrl, 0, sizeof(Ctrl), st), "memset ctrl");
  if (prof) cudaEventRecord(g_prof.ev[0], st);
  // ---- small problems: scaling, factorisation, the ADMM loop (adaptive-rho refactorisations included, on the device)
  //      and the finalisation are ONE launch (iterate_row.cu, FUSED) -- unless the caller needs the operators to stay
  //      in the workspace afterwards (recording pass of the unrolled mode)
  if (!rec && !prep && cfg->keep_operators == 0 && forward_fused_applies<T>(*cfg, w)) {
    if (host) {
      const size_t sv = (size_t)B * n * sizeof(T);
      CK(cudaMemcpyAsync((void*)p, host->p, sv, cudaMemcpyHostToDevice, st), "H2D p");
      CK(cudaMemcpyAsync((void*)lb, host->lb, sv, cudaMemcpyHostToDevice, st), "H2D lb");
      CK(cudaMemcpyAsync((void*)ub, host->ub, sv, cudaMemcpyHostToDevice, st), "H2D ub");
      if (m > 0) {
        CK(cudaMemcpyAsync((void*)A, host->A, (size_t)B * m * n * sizeof(T), cudaMemcpyHostToDevice, st), "H2D A");
        CK(cudaMemcpyAsync((void*)b, host->b, (size_t)B * m * sizeof(T), cudaMemcpyHostToDevice, st), "H2D b");
      }
      CK(cudaMemcpyAsync((void*)Q, host->Q, (size_t)B * n * n * sizeof(T), cudaMemcpyHostToDevice, st), "H2D Q");
    }
    CK(launch_bound_flags<T>(w, lb, ub, st), "bound_flags");
    // asynchronous form (lqpb_forward_async_*): the host waits for the bound flags alone -- the one thing the caller's
    // Python needs before it can return (reference :33-38) -- and collects iter / status whenever it wants them.
    // Stream order: flags kernel, copy of the control block (flags valid), event, solve, copy of the control block.
    const bool go_async = async_ctrl != nullptr && !host && !cfg->verbose;
    cudaEvent_t flags_done = nullptr;
    if (go_async) {
      cudaEvent_t& ev = g_hctrl.seg_done[cur_dev()];
      if (!ev) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming), "event");
      flags_done = ev;
      CK(cudaMemcpyAsync(async_ctrl, w.ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, st), "copy ctrl (flags)");
      CK(cudaEventRecord(flags_done, st), "record (flags)");
    }
    if (prof) cudaEventRecord(g_prof.ev[1], st);
    if (prof) cudaEventRecord(g_prof.it0[0], st);
    bool taken = false;
    CK(launch_forward_fused<T>(*cfg, w, Q, p, A, b, lb, ub, x, z, u, lams, nus, rho_out, &g_prof.it_launches, st, &taken),
       "fused forward");
    if (!taken) return fail(LQPB_E_CUDA, "fused forward kernel did not launch");
    if (prof) {
      cudaEventRecord(g_prof.it1[0], st);
      cudaEventRecord(g_prof.ev[2], st);
      cudaEventRecord(g_prof.ev[3], st);
    }
    g_prof.n_it = 1;
    g_prof.launches += 2;
    if (go_async) {
      CK(cudaEventSynchronize(flags_done), "synchronize (bound flags)");
      const Ctrl* ac = static_cast<const Ctrl*>(async_ctrl);
      info->iter = -1;                 // pending: lqpb_forward_collect fills these once the stream has drained
      info->status = 0;
      info->n_factor = 0;
      info->any_lb = ac->any_lb;
      info->any_ub = ac->any_ub;
      info->n_log = 0;
      CK(cudaMemcpyAsync(async_ctrl, w.ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, st), "copy ctrl (result)");
      if (deferred) *deferred = 1;
      g_prof.fwd_valid = prof;
      return LQPB_OK;
    }
    CK(cudaMemcpyAsync(hc, w.ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, st), "copy ctrl");
    if (host && host->x) CK(cudaMemcpyAsync(host->x, x, (size_t)B * n * sizeof(T), cudaMemcpyDeviceToHost, st), "D2H x");
    CK(cudaStreamSynchronize(st), "synchronize (fused forward)");
    if (hc->status != LQPB_STATUS_CONVERGED && hc->status != LQPB_STATUS_MAX_ITERS && hc->status != LQPB_STATUS_BREAKDOWN)
      return fail(LQPB_E_CUDA, "fused forward kernel ended without a status");
    info->iter = hc->iter;
    info->status = hc->status;
    info->n_factor = 1 + hc->pad0;
    info->any_lb = hc->any_lb;
    info->any_ub = hc->any_ub;
    info->n_log = cfg->verbose ? hc->n_log : 0;
    if (cfg->verbose) {
      memcpy(info->log_iter, hc->log_iter, sizeof(info->log_iter));
      memcpy(info->log_primal, hc->log_primal, sizeof(info->log_primal));
      memcpy(info->log_dual, hc->log_dual, sizeof(info->log_dual));
    }
    g_prof.fwd_valid = prof;
    return LQPB_OK;
  }
  bool factored = false;
  if (host) {
    // ---- inputs on the host: the vectors first (they decide any_lb / any_ub for the whole batch), then Q chunk
    //      by chunk on the copy stream; scaling + rho + factorisation of a chunk start as soon as it has landed
    // (the one-CTA-per-problem Gauss-Jordan kernel of the fp64 / small-problem path takes as long for 32 problems as
    // for 128, so its chains do not pipeline: one chunk there)
    const int C = w.tc ? pick_chunks(host->chunks, B) : 1;
    CK(pipe_init(), "copy stream");
    cudaStream_t cs = g_pipe.cs;
    CK(cudaEventRecord(g_pipe.fork, st), "fork");
    CK(cudaStreamWaitEvent(cs, g_pipe.fork, 0), "fork wait");
    const size_t sv = (size_t)B * n * sizeof(T);
    CK(cudaMemcpyAsync((void*)p, host->p, sv, cudaMemcpyHostToDevice, cs), "H2D p");
    CK(cudaMemcpyAsync((void*)lb, host->lb, sv, cudaMemcpyHostToDevice, cs), "H2D lb");
    CK(cudaMemcpyAsync((void*)ub, host->ub, sv, cudaMemcpyHostToDevice, cs), "H2D ub");
    if (m > 0) {
      CK(cudaMemcpyAsync((void*)A, host->A, (size_t)B * m * n * sizeof(T), cudaMemcpyHostToDevice, cs), "H2D A");
      CK(cudaMemcpyAsync((void*)b, host->b, (size_t)B * m * sizeof(T), cudaMemcpyHostToDevice, cs), "H2D b");
    }
    CK(cudaEventRecord(g_pipe.vec, cs), "vec event");
    // (Sending only the lower triangle of the symmetric Q was measured twice and lost both times: strided 3-D DMA
    // copies of the block-lower part -- 56 % of the bytes -- take as long as the full contiguous copy, and a kernel
    // pulling the triangle from page-locked host memory with 16-byte loads is slower still)
    for (int c = 0; c < C; ++c) {
      const int b0 = chunk_lo(B, C, c), bc = chunk_lo(B, C, c + 1) - b0;
      CK(cudaMemcpyAsync((void*)(Q + (size_t)b0 * n * n), host->Q + (size_t)b0 * n * n, (size_t)bc * n * n * sizeof(T),
                         cudaMemcpyHostToDevice, cs), "H2D Q");
      CK(cudaEventRecord(g_pipe.ev[c], cs), "chunk event");
    }
    CK(cudaStreamWaitEvent(st, g_pipe.vec, 0), "vec wait");
    CK(launch_bound_flags<T>(w, lb, ub, st), "bound_flags");
    g_prof.launches += 1;
    if (prof) cudaEventRecord(g_prof.ev[1], st);
    if (prof) cudaEventRecord(g_prof.fac0[0], st);
    for (int c = 0; c < C; ++c) {
      const int b0 = chunk_lo(B, C, c), bc = chunk_lo(B, C, c + 1) - b0;
      const FwdWs<T> wc = slice_fwd(w, b0, bc);
      CK(cudaStreamWaitEvent(st, g_pipe.ev[c], 0), "chunk wait");
      CK(launch_scale<T>(*cfg, wc, Q + (size_t)b0 * n * n, p + (size_t)b0 * n, off(A, (size_t)b0 * m * n),
                         off(b, (size_t)b0 * m), lb + (size_t)b0 * n, ub + (size_t)b0 * n, st), "scale");
      CK(launch_select_rho<T>(*cfg, wc, st), "select_rho");
      g_prof.launches += 3 + (cfg->scale ? 1 : 0);      // (colmax), scale_vec, scale_pack, select_rho: kernels only
      rc = factor_forward<T>(wc, true, st);
      if (rc) return rc;
    }
    if (prof) cudaEventRecord(g_prof.fac1[0], st);
    g_prof.n_fac = 1;
    factored = true;
  } else {
    CK(launch_scale<T>(*cfg, w, Q, p, A, b, lb, ub, st), "scale");
    CK(launch_select_rho<T>(*cfg, w, st), "select_rho");
    g_prof.launches += 3 + (cfg->scale ? 1 : 0);        // (colmax), scale_vec, scale_pack, select_rho: kernels only
    if (prof) cudaEventRecord(g_prof.ev[1], st);
  }

  int n_factor = 0, i0 = 0, skip = 0;
  while (true) {
    if (!(factored && n_factor == 0)) {
      if (prof && g_prof.n_fac < kMaxSeg) cudaEventRecord(g_prof.fac0[g_prof.n_fac], st);
      rc = factor_forward<T>(w, n_factor == 0, st);
      if (rc) return rc;
      if (prof && g_prof.n_fac < kMaxSeg) cudaEventRecord(g_prof.fac1[g_prof.n_fac++], st);
    }
    if (rec) {      // keep the operators of this segment for its reverse sweep
      if (n_factor >= rec->max_seg)
        return fail(rec->snap ? LQPB_E_ARG : LQPB_E_TAPE, "recording run needs more operator segments than announced");
      if (rec->snap) {
        const Snap<T> sn = carve_snap<T>((char*)rec->snap + (size_t)n_factor * carve_snap<T>(nullptr, B, n, m).bytes, B, n, m);
        const size_t mm = m > 0 ? m : 1, sz = sizeof(T);
        CK(cudaMemcpyAsync(sn.Kp, w.Kp, (size_t)B * Pack<T>::elems(n) * sz, cudaMemcpyDeviceToDevice, st), "snapshot K11");
        CK(cudaMemcpyAsync(sn.Gt, w.Gt, (size_t)B * mm * w.ld * sz, cudaMemcpyDeviceToDevice, st), "snapshot K21");
        CK(cudaMemcpyAsync(sn.Sinv, w.Sinv, (size_t)B * mm * mm * sz, cudaMemcpyDeviceToDevice, st), "snapshot K22");
        CK(cudaMemcpyAsync(sn.c, w.c, (size_t)B * w.ld * sz, cudaMemcpyDeviceToDevice, st), "snapshot c");
        CK(cudaMemcpyAsync(sn.rho, w.rho, (size_t)B * sz, cudaMemcpyDeviceToDevice, st), "snapshot rho");
      }
      rec->seg_start[n_factor] = i0;
    }
    ++n_factor;
    if (prof && g_prof.n_it < kMaxSeg) cudaEventRecord(g_prof.it0[g_prof.n_it], st);
    CK(launch_iterate<T>(*cfg, w, i0, skip, nus, &g_prof.it_launches, st, rec ? &rec->tape : nullptr), "iterate");
    if (prof && g_prof.n_it < kMaxSeg) cudaEventRecord(g_prof.it1[g_prof.n_it++], st);
    if (prof) cudaEventRecord(g_prof.ev[2], st);
    CK(launch_finalize<T>(w, x, z, u, lams, rho_out, st), "finalize");
    if (prof) cudaEventRecord(g_prof.ev[3], st);
    g_prof.launches += 2;
    // (the control block travels as SM stores into the page-locked buffer, not through the D2H copy engine: that engine
    // may be in the middle of another call's 128 MB of gradients -- solve-ahead, DESIGN.md 4a -- and would hold the
    // end of this solve up until it is through)
    CK(launch_mapped_copy(hc, w.ctrl, sizeof(Ctrl), st), "copy ctrl");
    g_prof.launches += 1;
    if (host && host->x)     // wasted (and overwritten later) only in the rare segment that ends in a refactorisation
      CK(cudaMemcpyAsync(host->x, x, (size_t)B * n * sizeof(T), cudaMemcpyDeviceToHost, st), "D2H x");
    if (prep) {
      // the mask, the assembly and the block LDL^T of the adjoint system depend on the solution only, not on dl_dz:
      // they are queued behind the solve NOW, and the host waits for the solve alone -- the GPU then works through
      // them while the caller is back in Python on its way to .backward() (wasted only if this segment ends in a
      // refactorisation, or if no backward follows)
      cudaEvent_t& seg_done = g_hctrl.seg_done[cur_dev()];
      if (!seg_done) CK(cudaEventCreateWithFlags(&seg_done, cudaEventDisableTiming), "event");
      CK(cudaEventRecord(seg_done, st), "record (segment end)");
      const int fwd_launches = g_prof.launches;
      rc = backward_impl<T>(B, n, m, nullptr, x, u, lams, nus, Q, A, lb, ub, nullptr, 0.0, nullptr, nullptr, nullptr,
                            nullptr, nullptr, nullptr, prep->ws, prep->ws_bytes, stream, prep->kkt != 0, nullptr, nullptr, 1);
      g_prof.launches = fwd_launches;
      if (rc) return rc;
      CK(cudaEventSynchronize(seg_done), "synchronize (segment end)");
    } else {
      CK(cudaStreamSynchronize(st), "synchronize (segment end)");
    }
    if (hc->status == 3) {
      i0 = hc->next_i;
      skip = 1;
      if (rec && n_factor < rec->max_seg)     // the flags the kernel just applied to rho (w.wants is rewritten at the next check)
        CK(cudaMemcpyAsync(rec->wants + (size_t)(n_factor - 1) * B, w.wants, (size_t)B * sizeof(int), cudaMemcpyDeviceToDevice,
                           st), "snapshot wants");
      CK(cudaMemsetAsync(&w.ctrl->status, 0, sizeof(int), st), "reset status");
      continue;
    }
    break;
  }
  if (hc->status != LQPB_STATUS_CONVERGED && hc->status != LQPB_STATUS_MAX_ITERS && hc->status != LQPB_STATUS_BREAKDOWN)
    return fail(LQPB_E_CUDA, "iteration kernel ended without a status");
  if (rec && !rec->snap) {          // first-pass mode: the tape only has to be long enough
    if (hc->status == LQPB_STATUS_MAX_ITERS && rec->true_max_iters > rec->tape.n_iter)
      return fail(LQPB_E_TAPE, "the solve did not converge within the tape capacity");
    rec->seg_start[n_factor] = hc->iter + 1;
  } else if (rec) {
    if (n_factor != rec->max_seg || hc->iter != rec->tape.n_iter - 1)
      return fail(LQPB_E_CUDA, "recording run did not reproduce the forward solve");
    rec->seg_start[n_factor] = rec->tape.n_iter;
  }
  info->iter = hc->iter;
  info->status = hc->status;
  info->n_factor = n_factor;
  info->any_lb = hc->any_lb;
  info->any_ub = hc->any_ub;
  info->n_log = cfg->verbose ? hc->n_log : 0;
  if (cfg->verbose) {
    memcpy(info->log_iter, hc->log_iter, sizeof(info->log_iter));
    memcpy(info->log_primal, hc->log_primal, sizeof(info->log_primal));
    memcpy(info->log_dual, hc->log_dual, sizeof(info->log_dual));
  }
  g_prof.fwd_valid = prof;
  return LQPB_OK;
}

template <typename T>
int backward_impl(int B, int n, int m, const T* dl_dz, const T* x, const T* u, const T* lams, const T* nus,
                  const T* Q, const T* A, const T* lb, const T* ub, const T* rho_dev, double rho_scalar, T* dQ,
                  T* dp, T* dA, T* db, T* dlb, T* dub, void* ws, size_t ws_bytes, void* stream, bool kkt,
                  int32_t* any_bounds, const HostBwd<T>* host, int stage) {
  // stage 0: the whole backward.  stage 1 ("prepare", enqueued by the forward call): mask / KKT diagonal, assembly and
  // block LDL^T of the adjoint system -- everything that does not depend on dl_dz.  stage 2 ("finish"): substitution
  // with dl_dz and the gradients, on a workspace stage 1 left prepared.  Stages exist for the tensor-core path only.
  if (host && !host->dl_dz) return fail(LQPB_E_ARG, "null host pointer argument");
  if ((!dl_dz && stage != 1) || !x || (!u && !kkt) || !lams || !Q || !lb || !ub || !ws)
    return fail(LQPB_E_ARG, "null pointer argument");
  if (B <= 0 || n <= 0 || m < 0) return fail(LQPB_E_ARG, "bad dimensions");
  if (m > 0 && (!A || !nus)) return fail(LQPB_E_ARG, "A and nus are required when m > 0");
  if (m > kMaxM) return fail(LQPB_E_ARG, "more than 256 equality rows are not supported");
  int rc = check_device();
  if (rc) return rc;
  BwdWs<T> w = carve_bwd<T>(ws, B, n, m);
  if (w.bytes > ws_bytes) return fail(LQPB_E_WORKSPACE, "workspace too small");
  if (stage != 0 && (!w.tc || (host && stage == 1)))
    return fail(LQPB_E_ARG, "staged backward needs the tensor-core path (and device buffers for the prepare stage)");
  cudaStream_t st = (cudaStream_t)stream;
  const bool prof = g_prof_on;
  ProfState& g_prof = prof_state();
  if (prof) prof_init(g_prof);
  int bwd_fac_launches = 0;
  if (stage != 2) {
    g_prof.bwd_valid = false;
    if (stage == 0) g_prof.launches = 0;
    if (prof) cudaEventRecord(g_prof.ev[4], st);
    if (kkt) CK(cudaMemsetAsync(w.flags, 0, 4 * sizeof(int), st), "memset flags");
  }
  // device-buffer call: one chunk = the whole batch.  Host-buffer call: the adjoint chain of chunk c runs on the
  // compute stream while the copy stream returns the dQ rows of chunk c - 1 to the host.
  const int C = (host && w.tc) ? pick_chunks(host->chunks, B) : 1;
  cudaStream_t cs = nullptr;
  if (host) {
    CK(pipe_init(), "copy stream");
    cs = g_pipe.cs;
    // cudaMemcpyDefault: the "host" dl_dz may already be a device copy (a caller that uploaded it itself so that its own
    // H2D traffic -- e.g. the prefetch of the next batch -- is ordered behind it); the copy engine serves H2D in order
    if (host->dl_dz != dl_dz)
      CK(cudaMemcpyAsync((void*)dl_dz, host->dl_dz, (size_t)B * n * sizeof(T), cudaMemcpyDefault, st), "H2D dl_dz");
  }
  for (int c = 0; c < C; ++c) {
    const int b0 = chunk_lo(B, C, c), bc = chunk_lo(B, C, c + 1) - b0;
    const size_t o = (size_t)b0;
    const BwdWs<T> wc = C > 1 ? slice_bwd(w, b0, bc) : w;
    const T *xc = x + o * n, *uc = off(u, o * n), *lamc = lams + o * 2 * n, *nuc = off(nus, o * m);
    const T *Qc = Q + o * n * n, *Ac = off(A, o * m * n), *lbc = lb + o * n, *ubc = ub + o * n, *gc = dl_dz + o * n;
    if (stage != 2) {
      if (kkt) CK(launch_bwd_kkt_prep<T>(wc, xc, lamc, lbc, ubc, st), "bwd_kkt_prep");
      else CK(launch_bwd_mask<T>(wc, xc, uc, lbc, ubc, st), "bwd_mask");
    }
    {
      GjArgs<T> a{};
      a.n = n; a.m = m; a.np = wc.np;
      a.src = Qc; a.lds = n;
      a.diag_shift = nullptr; a.diag_const = kkt ? T(0) : T(1e-8);   // :392 small regulariser on the whole diagonal
      a.diag_vec = kkt ? wc.dvec : nullptr;                          // KKT mode: Q + G^T diag(lam / s) G, no regulariser
      a.mask = wc.mask; a.ldm = wc.ld;
      a.Arows = Ac; a.lda = n; a.a_diag = kkt ? T(0) : T(1e-8);
      a.W = wc.W; a.Vg = wc.Vg; a.Wg = wc.Wg;
      a.dst = nullptr; a.ldd = wc.ld; a.G21 = nullptr; a.K22 = nullptr;
      a.bt = nullptr; a.c_out = nullptr;
      a.rhs_g = gc; a.sol_x = wc.dv; a.sol_nu = wc.dnu;
      bool done = false;
      if (wc.tc) {
        int l = 0;
        CK(launch_tc_ldl_solve(bc, a, wc.Pb, wc.nb, st, &l, stage), "tensor-core LDL solve (backward)");
        bwd_fac_launches += l;
        done = true;
      }
      if (!done) {
        CK(launch_ldl_solve<T>(bc, a, st), "ldl_solve (backward)");
        bwd_fac_launches += 1;
      }
    }
    if (stage == 1) {
      if (prof) cudaEventRecord(g_prof.ev[5], st);
      g_prof.prep_launches = 1 + bwd_fac_launches;
      return LQPB_OK;
    }
    if (prof && c == C - 1 && stage == 0) cudaEventRecord(g_prof.ev[5], st);
    if (prof && c == C - 1) cudaEventRecord(g_prof.ev[6], st);
    CK(launch_bwd_grads<T>(wc, gc, xc, uc, lamc, nuc, Qc, Ac, off(rho_dev, o), rho_scalar, off(dQ, o * n * n),
                           off(dp, o * n), off(dA, o * m * n), off(db, o * m), off(dlb, o * n), off(dub, o * n), st,
                           kkt ? lbc : nullptr, kkt ? ubc : nullptr),
       "bwd_grads");
    if (host && host->dQ && dQ) {
      CK(cudaEventRecord(g_pipe.ev[c], st), "chunk event");
      CK(cudaStreamWaitEvent(cs, g_pipe.ev[c], 0), "chunk wait");
      CK(cudaMemcpyAsync(host->dQ + o * n * n, dQ + o * n * n, (size_t)bc * n * n * sizeof(T), cudaMemcpyDeviceToHost,
                         cs), "D2H dQ");
    }
  }
  if (prof) cudaEventRecord(g_prof.ev[7], st);
  g_prof.launches = (stage == 2 ? 1 + g_prof.prep_launches : 2 * C) + bwd_fac_launches;
  g_prof.bwd_valid = prof;
  if (host) {
    // the small gradients follow the last chunk on the compute stream; then wait for both streams
    const size_t sv = (size_t)B * n * sizeof(T);
    if (host->dp && dp) CK(cudaMemcpyAsync(host->dp, dp, sv, cudaMemcpyDeviceToHost, st), "D2H dp");
    if (host->dlb && dlb) CK(cudaMemcpyAsync(host->dlb, dlb, sv, cudaMemcpyDeviceToHost, st), "D2H dlb");
    if (host->dub && dub) CK(cudaMemcpyAsync(host->dub, dub, sv, cudaMemcpyDeviceToHost, st), "D2H dub");
    if (m > 0 && host->dA && dA)
      CK(cudaMemcpyAsync(host->dA, dA, (size_t)B * m * n * sizeof(T), cudaMemcpyDeviceToHost, st), "D2H dA");
    if (m > 0 && host->db && db)
      CK(cudaMemcpyAsync(host->db, db, (size_t)B * m * sizeof(T), cudaMemcpyDeviceToHost, st), "D2H db");
    if (kkt && any_bounds)
      CK(cudaMemcpyAsync(any_bounds, w.flags, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, st), "copy bound flags");
    CK(cudaEventRecord(g_pipe.done, cs), "copy done");
    CK(cudaStreamWaitEvent(st, g_pipe.done, 0), "join");
    CK(cudaStreamSynchronize(st), "synchronize (host gradients)");
    return LQPB_OK;
  }
  if (kkt && any_bounds) {
    CK(cudaMemcpyAsync(any_bounds, w.flags, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, st), "copy bound flags");
    CK(cudaStreamSynchronize(st), "synchronize (bound flags)");
  }
  return LQPB_OK;
}

// ---- unrolled mode: recording pass and reverse sweep on the operators a forward call left in its workspace
template <typename T>
int unroll_record_impl(const lqpb_config* cfg, int B, int n, int m, int n_iter, void* ws, size_t ws_bytes, T* tx,
                       T* tz, T* tu, T* tnu, void* stream) {
  if (!cfg || !ws || !tx || !tz || !tu || (m > 0 && !tnu)) return fail(LQPB_E_ARG, "null pointer argument");
  if (B <= 0 || n <= 0 || m < 0 || m > kMaxM || n_iter < 1) return fail(LQPB_E_ARG, "bad dimensions");
  int rc = check_device();
  if (rc) return rc;
  FwdWs<T> w = carve_fwd<T>(ws, B, n, m);
  if (w.bytes > ws_bytes) return fail(LQPB_E_WORKSPACE, "workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  if (!g_hctrl.pinned) CK(cudaMallocHost(&g_hctrl.pinned, sizeof(Ctrl)), "cudaMallocHost");
  Ctrl* hc = g_hctrl.pinned;
  // same start as the forward solve: z = u = 0, no check results yet; any_lb / any_ub are kept
  CK(cudaMemsetAsync(w.ctrl, 0, offsetof(Ctrl, any_lb), st), "reset ctrl");
  CK(cudaMemsetAsync((char*)w.ctrl + offsetof(Ctrl, n_log), 0, sizeof(Ctrl) - offsetof(Ctrl, n_log), st), "reset ctrl");
  CK(cudaMemsetAsync(w.z, 0, (size_t)B * w.ld * sizeof(T), st), "reset z");
  CK(cudaMemsetAsync(w.u, 0, (size_t)B * w.ld * sizeof(T), st), "reset u");
  CK(cudaMemsetAsync(w.wants, 0, (size_t)B * sizeof(int), st), "reset wants");
  lqpb_config c = *cfg;
  c.max_iters = n_iter;
  c.verbose = 0;
  Tape<T> tape{n_iter, tx, tz, tu, tnu};
  int l = 0;
  CK(launch_iterate<T>(c, w, 0, 0, (T*)nullptr, &l, st, &tape), "iterate (recording pass)");
  CK(cudaMemcpyAsync(hc, w.ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, st), "copy ctrl");
  CK(cudaStreamSynchronize(st), "synchronize (recording pass)");
  if (hc->status == 3) return fail(LQPB_E_ARG, "recording pass hit an adaptive-rho update (the forward solve must have n_factor == 1)");
  if ((hc->status != LQPB_STATUS_CONVERGED && hc->status != LQPB_STATUS_MAX_ITERS) || hc->iter != n_iter - 1)
    return fail(LQPB_E_CUDA, "recording pass did not reproduce the forward solve");
  return LQPB_OK;
}

template <typename T>
int unroll_forward_impl(const lqpb_config* cfg, int B, int n, int m, int n_iter, int n_seg, const T* Q, const T* p,
                        const T* A, const T* b, const T* lb, const T* ub, T* x, T* z, T* u, T* lams, T* nus, T* rho_out,
                        T* tx, T* tz, T* tu, T* tnu, void* snap, size_t snap_bytes, int32_t* seg_start, int32_t* wants,
                        lqpb_info* info, void* ws, size_t ws_bytes, void* stream) {
  if (!cfg || !tx || !tz || !tu || (m > 0 && !tnu) || !seg_start || (n_seg > 1 && (!wants || !snap)))
    return fail(LQPB_E_ARG, "null pointer argument");
  if (n_iter < 1 || n_seg < 1) return fail(LQPB_E_ARG, "bad dimensions");
  if (snap && B > 0 && n > 0 && m >= 0 && carve_snap<T>(nullptr, B, n, m).bytes * (size_t)n_seg > snap_bytes)
    return fail(LQPB_E_WORKSPACE, "snapshot buffer too small");
  lqpb_config c = *cfg;
  if (c.max_iters > n_iter) c.max_iters = n_iter;      // the tape has n_iter rows: the run can never write past them
  if (snap) c.verbose = 0;                             // a second pass stays silent; a first pass is the solve itself
  UnrollRec<T> rec{Tape<T>{n_iter, tx, tz, tu, tnu}, snap, n_seg, seg_start, wants, cfg->max_iters};
  return forward_impl<T>(&c, B, n, m, Q, p, A, b, lb, ub, x, z, u, lams, nus, rho_out, info, ws, ws_bytes, stream, nullptr,
                         &rec);
}

template <typename T>
int unroll_backward_impl(int B, int n, int m, int n_iter, int k_lo, int k_hi, void* ws, size_t ws_bytes, const void* snap,
                         const T* gx, const T* gz_last, const T* gu_last, const T* gzprev_last, const T* tx, const T* tz,
                         const T* tu, const T* tnu, T* tw, T* twnu, T* gQ, T* gp, T* gA, T* gb, T* glb, T* gub, T* grho,
                         T* gz_in, T* gu_in, void* stream) {
  if (!ws || !tx || !tz || !tu || !tw || !gp || !glb || !gub || !grho) return fail(LQPB_E_ARG, "null pointer argument");
  if (m > 0 && (!tnu || !twnu || !gb)) return fail(LQPB_E_ARG, "tape_nu, tape_wnu and gb are required when m > 0");
  if (B <= 0 || n <= 0 || m < 0 || m > kMaxM || n_iter < 1 || k_lo < 0 || k_hi < k_lo || k_hi >= n_iter)
    return fail(LQPB_E_ARG, "bad dimensions");
  int rc = check_device();
  if (rc) return rc;
  FwdWs<T> w = carve_fwd<T>(ws, B, n, m);
  if (w.bytes > ws_bytes) return fail(LQPB_E_WORKSPACE, "workspace too small");
  if (snap) {                       // operators of an earlier segment; bounds and flags stay those of the workspace
    const Snap<T> sn = carve_snap<T>(const_cast<void*>(snap), B, n, m);
    w.Kp = sn.Kp; w.Gt = sn.Gt; w.Sinv = sn.Sinv; w.c = sn.c; w.rho = sn.rho;
  }
  Tape<T> tape{n_iter, (T*)tx, (T*)tz, (T*)tu, (T*)tnu};
  UnrollGrads<T> g{k_lo, k_hi, gx, gz_last, gu_last, gzprev_last, gz_in, gu_in, tw, twnu, gQ, gp, gA, gb, glb, gub, grho};
  int l = 0;
  CK(launch_unroll_reverse<T>(w, tape, g, &l, (cudaStream_t)stream), "unroll reverse sweep");
  prof_state().launches = l;
  return LQPB_OK;
}

}  // namespace

extern "C" {

int lqpb_abi_version(void) { return LQPB_ABI_VERSION; }
const char* lqpb_last_error(void) { return g_err.c_str(); }
void lqpb_profile_enable(int on) { g_prof_on = on != 0; }

void lqpb_profile_get(lqpb_profile* out) {
  if (!out) return;
  memset(out, 0, sizeof(*out));
  ProfState& g_prof = prof_state();
  auto el = [](cudaEvent_t a, cudaEvent_t b) {
    float ms = 0.f;
    cudaEventSynchronize(b);
    cudaEventElapsedTime(&ms, a, b);
    return ms;
  };
  if (g_prof.fwd_valid) {
    out->scale_ms = el(g_prof.ev[0], g_prof.ev[1]);
    for (int i = 0; i < g_prof.n_fac; ++i) out->factor_ms += el(g_prof.fac0[i], g_prof.fac1[i]);
    for (int i = 0; i < g_prof.n_it; ++i) out->iterate_ms += el(g_prof.it0[i], g_prof.it1[i]);
    out->finalize_ms = el(g_prof.ev[2], g_prof.ev[3]);
  }
  if (g_prof.bwd_valid) {
    out->bwd_factor_ms = el(g_prof.ev[4], g_prof.ev[5]);
    out->bwd_solve_ms = el(g_prof.ev[5], g_prof.ev[6]);
    out->bwd_grad_ms = el(g_prof.ev[6], g_prof.ev[7]);
  }
  out->iterate_launches = g_prof.it_launches;
  out->factor_launches = g_prof.fac_launches;
  out->kernel_launches = g_prof.launches;
}

size_t lqpb_forward_workspace_bytes_f32(int B, int n, int m) { return carve_fwd<float>(nullptr, B, n, m).bytes; }
size_t lqpb_forward_workspace_bytes_f64(int B, int n, int m) { return carve_fwd<double>(nullptr, B, n, m).bytes; }
size_t lqpb_backward_workspace_bytes_f32(int B, int n, int m) { return carve_bwd<float>(nullptr, B, n, m).bytes; }
size_t lqpb_backward_workspace_bytes_f64(int B, int n, int m) { return carve_bwd<double>(nullptr, B, n, m).bytes; }

int lqpb_forward_f32(const lqpb_config* cfg, int B, int n, int m, const float* Q, const float* p, const float* A,
                     const float* b, const float* lb, const float* ub, float* x, float* z, float* u, float* lams,
                     float* nus, float* rho_out, lqpb_info* info, void* workspace, size_t workspace_bytes,
                     void* stream) {
  return forward_impl<float>(cfg, B, n, m, Q, p, A, b, lb, ub, x, z, u, lams, nus, rho_out, info, workspace,
                             workspace_bytes, stream);
}
int lqpb_forward_f64(const lqpb_config* cfg, int B, int n, int m, const double* Q, const double* p, const double* A,
                     const double* b, const double* lb, const double* ub, double* x, double* z, double* u,
                     double* lams, double* nus, double* rho_out, lqpb_info* info, void* workspace,
                     size_t workspace_bytes, void* stream) {
  return forward_impl<double>(cfg, B, n, m, Q, p, A, b, lb, ub, x, z, u, lams, nus, rho_out, info, workspace,
                              workspace_bytes, stream);
}

int lqpb_backward_f32(int B, int n, int m, const float* dl_dz, const float* x, const float* u, const float* lams,
                      const float* nus, const float* Q, const float* A, const float* lb, const float* ub,
                      const float* rho_dev, double rho_scalar, float* dQ, float* dp, float* dA, float* db,
                      float* dlb, float* dub, void* workspace, size_t workspace_bytes, void* stream) {
  return backward_impl<float>(B, n, m, dl_dz, x, u, lams, nus, Q, A, lb, ub, rho_dev, rho_scalar, dQ, dp, dA, db, dlb,
                              dub, workspace, workspace_bytes, stream);
}
int lqpb_backward_f64(int B, int n, int m, const double* dl_dz, const double* x, const double* u, const double* lams,
                      const double* nus, const double* Q, const double* A, const double* lb, const double* ub,
                      const double* rho_dev, double rho_scalar, double* dQ, double* dp, double* dA, double* db,
                      double* dlb, double* dub, void* workspace, size_t workspace_bytes, void* stream) {
  return backward_impl<double>(B, n, m, dl_dz, x, u, lams, nus, Q, A, lb, ub, rho_dev, rho_scalar, dQ, dp, dA, db,
                               dlb, dub, workspace, workspace_bytes, stream);
}

#define WARM_ENTRY(SFX, T)                                                                                          \
  int lqpb_forward_warm_##SFX(const lqpb_config* cfg, int B, int n, int m, const T* Q, const T* p, const T* A,     \
                              const T* b, const T* lb, const T* ub, const T* z0, const T* u0, const T* rho0,      \
                              T* x, T* z, T* u, T* lams, T* nus, T* rho_out, lqpb_info* info, void* workspace,     \
                              size_t workspace_bytes, void* stream) {                                              \
    return forward_impl<T>(cfg, B, n, m, Q, p, A, b, lb, ub, x, z, u, lams, nus, rho_out, info, workspace,         \
                           workspace_bytes, stream, nullptr, nullptr, nullptr, z0, u0, nullptr, nullptr, rho0);    \
  }                                                                                                                \
  int lqpb_solution_status_##SFX(const lqpb_config* cfg, int B, int n, int m, void* workspace,                     \
                                 size_t workspace_bytes, T* residuals, int32_t* converged, void* stream) {         \
    if (!cfg || !workspace || !residuals || !converged || B <= 0 || n <= 0 || m < 0)                               \
      return fail(LQPB_E_ARG, "bad argument");                                                                     \
    int rc = check_device();                                                                                       \
    if (rc) return rc;                                                                                             \
    FwdWs<T> w = carve_fwd<T>(workspace, B, n, m);                                                                 \
    if (w.bytes > workspace_bytes) return fail(LQPB_E_WORKSPACE, "workspace too small");                           \
    CK(launch_status<T>(*cfg, w, residuals, converged, (cudaStream_t)stream), "status");                           \
    return LQPB_OK;                                                                                                \
  }
WARM_ENTRY(f32, float)
WARM_ENTRY(f64, double)

#define ASYNC_ENTRY(SFX, T)                                                                                         \
  int lqpb_forward_async_##SFX(const lqpb_config* cfg, int B, int n, int m, const T* Q, const T* p, const T* A,    \
                               const T* b, const T* lb, const T* ub, const T* z0, const T* u0, T* x, T* z, T* u,   \
                               T* lams, T* nus, T* rho_out, void* pinned_ctrl, lqpb_info* info, void* workspace,   \
                               size_t workspace_bytes, void* stream, int32_t* deferred) {                          \
    if (!pinned_ctrl || !deferred) return fail(LQPB_E_ARG, "null pointer argument");                               \
    *deferred = 0;                                                                                                 \
    return forward_impl<T>(cfg, B, n, m, Q, p, A, b, lb, ub, x, z, u, lams, nus, rho_out, info, workspace,         \
                           workspace_bytes, stream, nullptr, nullptr, nullptr, z0, u0, pinned_ctrl, deferred);     \
  }
ASYNC_ENTRY(f32, float)
ASYNC_ENTRY(f64, double)

size_t lqpb_ctrl_bytes(void) { return sizeof(Ctrl); }

#define REGIME_ENTRY(SFX, T)                                                                               \
  int lqpb_iterate_regime_##SFX(const lqpb_config* cfg, int B, int n, int m) {                             \
    if (!cfg || B <= 0 || n <= 0 || m < 0 || check_device()) return -1;                                    \
    const FwdWs<T> w = carve_fwd<T>(nullptr, B, n, m);                                                     \
    if (cfg->keep_operators == 0 && forward_fused_applies<T>(*cfg, w)) return LQPB_REGIME_FUSED_ROWS;      \
    if (iterate_rows_applies<T>(*cfg, w)) return LQPB_REGIME_ROWS;                                         \
    if (iterate_resident_applies<T>(*cfg, w)) return LQPB_REGIME_PACKED_RESIDENT;                          \
    if (iterate_split_size<T>(w) > 0) return LQPB_REGIME_STREAM_SPLIT;                                     \
    return LQPB_REGIME_STREAM;                                                                             \
  }
REGIME_ENTRY(f32, float)
REGIME_ENTRY(f64, double)

int lqpb_forward_collect(const void* pinned_ctrl, const lqpb_config* cfg, lqpb_info* info) {
  if (!pinned_ctrl || !cfg || !info) return fail(LQPB_E_ARG, "null pointer argument");
  const Ctrl* hc = static_cast<const Ctrl*>(pinned_ctrl);
  if (hc->status != LQPB_STATUS_CONVERGED && hc->status != LQPB_STATUS_MAX_ITERS && hc->status != LQPB_STATUS_BREAKDOWN)
    return fail(LQPB_E_CUDA, "the solve has not finished (synchronise the stream first) or ended without a status");
  info->iter = hc->iter;
  info->status = hc->status;
  info->n_factor = 1 + hc->pad0;
  info->any_lb = hc->any_lb;
  info->any_ub = hc->any_ub;
  info->n_log = 0;
  return LQPB_OK;
}

#define PREP_ENTRY(SFX, T)                                                                                         \
  int lqpb_forward_prep_##SFX(const lqpb_config* cfg, int B, int n, int m, const T* Q, const T* p, const T* A,     \
                              const T* b, const T* lb, const T* ub, T* x, T* z, T* u, T* lams, T* nus, T* rho_out, \
                              lqpb_info* info, void* workspace, size_t workspace_bytes, void* bwd_workspace,       \
                              size_t bwd_workspace_bytes, int kkt, int32_t* prepared, void* stream) {              \
    const bool can = bwd_workspace != nullptr && B > 0 && n > 0 && m >= 0 && tc_factor_enabled<T>(n, m);           \
    if (prepared) *prepared = can ? 1 : 0;                                                                         \
    BwdPrep<T> pr{bwd_workspace, bwd_workspace_bytes, kkt};                                                        \
    return forward_impl<T>(cfg, B, n, m, Q, p, A, b, lb, ub, x, z, u, lams, nus, rho_out, info, workspace,         \
                           workspace_bytes, stream, nullptr, nullptr, can ? &pr : nullptr);                        \
  }                                                                                                                \
  int lqpb_backward_finish_##SFX(int B, int n, int m, int kkt, const T* dl_dz, const T* x, const T* u,            \
                                 const T* lams, const T* nus, const T* Q, const T* A, const T* lb, const T* ub,   \
                                 const T* rho_dev, double rho_scalar, T* dQ, T* dp, T* dA, T* db, T* dlb, T* dub, \
                                 void* workspace, size_t workspace_bytes, void* stream) {                          \
    return backward_impl<T>(B, n, m, dl_dz, x, u, lams, nus, Q, A, lb, ub, rho_dev, rho_scalar, dQ, dp, dA, db,   \
                            dlb, dub, workspace, workspace_bytes, stream, kkt != 0, nullptr, nullptr, 2);          \
  }
PREP_ENTRY(f32, float)
PREP_ENTRY(f64, double)

#define HOST_ENTRY(SFX, T)                                                                                         \
  int lqpb_forward_host_##SFX(const lqpb_config* cfg, int B, int n, int m, const T* hQ, const T* hp, const T* hA,  \
                              const T* hb, const T* hlb, const T* hub, T* Q, T* p, T* A, T* b, T* lb, T* ub, T* x, \
                              T* z, T* u, T* lams, T* nus, T* rho_out, T* hx, lqpb_info* info, void* workspace,    \
                              size_t workspace_bytes, void* stream, int chunks, void* bwd_workspace,               \
                              size_t bwd_workspace_bytes, int bwd_kkt, int32_t* prepared) {                        \
    HostFwd<T> h{hQ, hp, hA, hb, hlb, hub, hx, chunks};                                                            \
    const bool can = bwd_workspace != nullptr && B > 0 && n > 0 && m >= 0 && tc_factor_enabled<T>(n, m);           \
    if (prepared) *prepared = can ? 1 : 0;                                                                         \
    BwdPrep<T> pr{bwd_workspace, bwd_workspace_bytes, bwd_kkt};                                                    \
    return forward_impl<T>(cfg, B, n, m, Q, p, A, b, lb, ub, x, z, u, lams, nus, rho_out, info, workspace,         \
                           workspace_bytes, stream, &h, nullptr, can ? &pr : nullptr);                             \
  }                                                                                                                \
  int lqpb_backward_host_##SFX(int B, int n, int m, int kkt, const T* h_dl_dz, T* dl_dz, const T* x, const T* u,  \
                               const T* lams, const T* nus, const T* Q, const T* A, const T* lb, const T* ub,     \
                               const T* rho_dev, double rho_scalar, T* dQ, T* dp, T* dA, T* db, T* dlb, T* dub,   \
                               T* hdQ, T* hdp, T* hdA, T* hdb, T* hdlb, T* hdub, int32_t* any_bounds,             \
                               void* workspace, size_t workspace_bytes, void* stream, int chunks, int prepared) { \
    HostBwd<T> h{h_dl_dz, hdQ, hdp, hdA, hdb, hdlb, hdub, chunks};                                                 \
    return backward_impl<T>(B, n, m, dl_dz, x, u, lams, nus, Q, A, lb, ub, rho_dev, rho_scalar, dQ, dp, dA, db,   \
                            dlb, dub, workspace, workspace_bytes, stream, kkt != 0, any_bounds, &h,                \
                            prepared ? 2 : 0);                                                                     \
  }
HOST_ENTRY(f32, float)
HOST_ENTRY(f64, double)

int lqpb_backward_kkt_f32(int B, int n, int m, const float* dl_dz, const float* x, const float* lams,
                          const float* nus, const float* Q, const float* A, const float* lb, const float* ub, float* dQ,
                          float* dp, float* dA, float* db, float* dlb, float* dub, int32_t* any_bounds, void* workspace,
                          size_t workspace_bytes, void* stream) {
  return backward_impl<float>(B, n, m, dl_dz, x, nullptr, lams, nus, Q, A, lb, ub, nullptr, 0.0, dQ, dp, dA, db, dlb,
                              dub, workspace, workspace_bytes, stream, true, any_bounds);
}
int lqpb_backward_kkt_f64(int B, int n, int m, const double* dl_dz, const double* x, const double* lams,
                          const double* nus, const double* Q, const double* A, const double* lb, const double* ub,
                          double* dQ, double* dp, double* dA, double* db, double* dlb, double* dub, int32_t* any_bounds,
                          void* workspace, size_t workspace_bytes, void* stream) {
  return backward_impl<double>(B, n, m, dl_dz, x, nullptr, lams, nus, Q, A, lb, ub, nullptr, 0.0, dQ, dp, dA, db, dlb,
                               dub, workspace, workspace_bytes, stream, true, any_bounds);
}

#define UNROLL_ENTRY(SFX, T)                                                                                       \
  size_t lqpb_unroll_snapshot_bytes_##SFX(int B, int n, int m) { return carve_snap<T>(nullptr, B, n, m).bytes; }    \
  int lqpb_unroll_record_##SFX(const lqpb_config* cfg, int B, int n, int m, int n_iter, void* workspace,            \
                               size_t workspace_bytes, T* tape_x, T* tape_z, T* tape_u, T* tape_nu, void* stream) { \
    return unroll_record_impl<T>(cfg, B, n, m, n_iter, workspace, workspace_bytes, tape_x, tape_z, tape_u,          \
                                 tape_nu, stream);                                                                  \
  }                                                                                                                \
  int lqpb_unroll_forward_##SFX(const lqpb_config* cfg, int B, int n, int m, int n_iter, int n_seg, const T* Q,     \
                                const T* p, const T* A, const T* b, const T* lb, const T* ub, T* x, T* z, T* u,     \
                                T* lams, T* nus, T* rho_out, T* tape_x, T* tape_z, T* tape_u, T* tape_nu,           \
                                void* snapshots, size_t snapshot_bytes, int32_t* seg_start, int32_t* wants,         \
                                lqpb_info* info, void* workspace, size_t workspace_bytes, void* stream) {           \
    return unroll_forward_impl<T>(cfg, B, n, m, n_iter, n_seg, Q, p, A, b, lb, ub, x, z, u, lams, nus, rho_out,     \
                                  tape_x, tape_z, tape_u, tape_nu, snapshots, snapshot_bytes, seg_start, wants,     \
                                  info, workspace, workspace_bytes, stream);                                        \
  }                                                                                                                \
  int lqpb_unroll_backward_##SFX(int B, int n, int m, int n_iter, int k_lo, int k_hi, void* workspace,             \
                                 size_t workspace_bytes, const void* snapshot, const T* g_x, const T* g_z,         \
                                 const T* g_u, const T* g_zprev, const T* tape_x, const T* tape_z,                 \
                                 const T* tape_u, const T* tape_nu, T* tape_w, T* tape_wnu, T* gQ, T* gp, T* gA,   \
                                 T* gb, T* glb, T* gub, T* grho, T* gz_in, T* gu_in, void* stream) {               \
    return unroll_backward_impl<T>(B, n, m, n_iter, k_lo, k_hi, workspace, workspace_bytes, snapshot, g_x, g_z,    \
                                   g_u, g_zprev, tape_x, tape_z, tape_u, tape_nu, tape_w, tape_wnu, gQ, gp, gA,    \
                                   gb, glb, gub, grho, gz_in, gu_in, stream);                                      \
  }
UNROLL_ENTRY(f32, float)
UNROLL_ENTRY(f64, double)

#define SCALE_GRAD_ENTRY(SFX, T)                                                                                   \
  size_t lqpb_unroll_scale_grad_scratch_elems_##SFX(int B, int n) { return (size_t)B * ((n + 31) / 32 + 1) * n; }  \
  int lqpb_unroll_scale_grad_##SFX(int B, int n, T* G, const T* Q, const T* D, const T* coef, T* gD, T* scratch,   \
                                   void* stream) {                                                                 \
    if (!G || !Q || !scratch || B <= 0 || n <= 0 || (D != nullptr) != (gD != nullptr))                             \
      return fail(LQPB_E_ARG, "bad argument");                                                                     \
    int rc = check_device();                                                                                       \
    if (rc) return rc;                                                                                             \
    CK(launch_scale_grad<T>(B, n, G, Q, D, coef, gD, scratch, (cudaStream_t)stream), "scale_grad");                \
    return LQPB_OK;                                                                                                \
  }
SCALE_GRAD_ENTRY(f32, float)
SCALE_GRAD_ENTRY(f64, double)

#define SCALE_VEC_ENTRY(SFX, T)                                                                                     \
  int lqpb_unroll_scaled_vectors_##SFX(int B, int n, int m, void* workspace, size_t workspace_bytes, T* D, T* pt,   \
                                       T* At, T* bt, T* lbt, T* ubt, T* E, void* stream) {                          \
    if (!workspace || !D || !pt || !lbt || !ubt || B <= 0 || n <= 0 || m < 0 || (m > 0 && (!At || !bt || !E)))      \
      return fail(LQPB_E_ARG, "bad argument");                                                                     \
    int rc = check_device();                                                                                       \
    if (rc) return rc;                                                                                             \
    FwdWs<T> w = carve_fwd<T>(workspace, B, n, m);                                                                 \
    if (w.bytes > workspace_bytes) return fail(LQPB_E_WORKSPACE, "workspace too small");                           \
    CK(launch_scaled_vectors<T>(w, D, pt, At, bt, lbt, ubt, E, (cudaStream_t)stream), "scaled_vectors");           \
    return LQPB_OK;                                                                                                \
  }                                                                                                                \
  int lqpb_unroll_scale_vec_grad_##SFX(int B, int n, int m, int beta_auto, double beta, int use_lb, int use_ub,    \
                                       const T* colmax, const T* p, const T* A, const T* b, const T* lb,           \
                                       const T* ub, const T* D, const T* E, const T* gD, const T* gD2,             \
                                       const T* gpt, const T* gAt, const T* gbt, const T* glbt, const T* gubt,     \
                                       T* gcolmax, T* gp, T* gA, T* gb, T* glb, T* gub, void* stream) {            \
    if (!colmax || !p || !lb || !ub || !D || !gcolmax || !gp || !glb || !gub || B <= 0 || n <= 0 || m < 0 ||       \
        (m > 0 && (!A || !b || !E || !gA || !gb)))                                                                 \
      return fail(LQPB_E_ARG, "bad argument");                                                                     \
    int rc = check_device();                                                                                       \
    if (rc) return rc;                                                                                             \
    ScaleVecGrad<T> a{n, m, beta_auto, use_lb, use_ub, (T)beta, colmax, p, A, b, lb, ub, D, E, gD, gD2, gpt, gAt,  \
                      gbt, glbt, gubt, gcolmax, gp, gA, gb, glb, gub};                                              \
    CK(launch_scale_vec_grad<T>(B, a, (cudaStream_t)stream), "scale_vec_grad");                                    \
    return LQPB_OK;                                                                                                \
  }                                                                                                                \
  int lqpb_unroll_colmax_##SFX(int B, int n, const T* Q, T* colmax, void* stream) {                                \
    if (!Q || !colmax || B <= 0 || n <= 0) return fail(LQPB_E_ARG, "bad argument");                                \
    int rc = check_device();                                                                                       \
    if (rc) return rc;                                                                                             \
    CK(launch_colmax_plain<T>(B, n, Q, colmax, (cudaStream_t)stream), "colmax");                                   \
    return LQPB_OK;                                                                                                \
  }                                                                                                                \
  int lqpb_unroll_colmax_grad_##SFX(int B, int n, const T* Q, const T* colmax, const T* gcolmax, T* G, void* stream) { \
    if (!Q || !colmax || !gcolmax || !G || B <= 0 || n <= 0) return fail(LQPB_E_ARG, "bad argument");              \
    int rc = check_device();                                                                                       \
    if (rc) return rc;                                                                                             \
    CK(launch_colmax_scatter<T>(B, n, Q, colmax, gcolmax, G, (cudaStream_t)stream), "colmax_grad");                \
    return LQPB_OK;                                                                                                \
  }
SCALE_VEC_ENTRY(f32, float)
SCALE_VEC_ENTRY(f64, double)

#define LU_ENTRY(SFX, T)                                                                                          \
  int lqpb_lu_factor_##SFX(int B, int N, const T* A, T* LU, int32_t* piv, void* stream) {                          \
    if (!A || !LU || !piv || B <= 0 || N <= 0) return fail(LQPB_E_ARG, "bad argument");                           \
    int rc = check_device();                                                                                       \
    if (rc) return rc;                                                                                             \
    CK(launch_lu_factor<T>(B, N, A, LU, piv, (cudaStream_t)stream), "lu_factor");                                  \
    return LQPB_OK;                                                                                                \
  }                                                                                                                \
  int lqpb_lu_solve_##SFX(int B, int N, int nrhs, const T* LU, const int32_t* piv, const T* rhs, T* x,            \
                          int negate_rhs, void* stream) {                                                          \
    if (!LU || !piv || !rhs || !x || B <= 0 || N <= 0 || nrhs <= 0) return fail(LQPB_E_ARG, "bad argument");       \
    int rc = check_device();                                                                                       \
    if (rc) return rc;                                                                                             \
    CK(launch_lu_solve<T>(B, N, nrhs, LU, piv, rhs, x, negate_rhs, (cudaStream_t)stream), "lu_solve");             \
    return LQPB_OK;                                                                                                \
  }                                                                                                                \
  int lqpb_outer_##SFX(int B, int N, int M, const T* a, const T* b, T* C, void* stream) {                          \
    if (!a || !b || !C || B <= 0 || N <= 0 || M <= 0) return fail(LQPB_E_ARG, "bad argument");                    \
    int rc = check_device();                                                                                       \
    if (rc) return rc;                                                                                             \
    CK(launch_outer<T>(B, N, M, a, b, C, (cudaStream_t)stream), "outer");                                          \
    return LQPB_OK;                                                                                                \
  }
LU_ENTRY(f32, float)
LU_ENTRY(f64, double)

// developer / diagnostic entry (tools/tc_check.py): inverse of B symmetric N x N matrices (N multiple of 128)
// through the tensor-core sweep of tcfactor.cu.  work: lqpb_dev_tc_inverse_work_bytes(B, N) bytes of device memory.
size_t lqpb_dev_tc_inverse_work_bytes(int B, int N) {
  const size_t nb = (size_t)N / 128;
  return (size_t)B * (nb * (nb + 1) / 2 + 3 * nb) * 128 * 128 * sizeof(float);
}
int lqpb_dev_tc_inverse_f32(int B, int N, const float* A, float* Ainv, void* work, void* stream) {
  if (!A || !Ainv || !work || B <= 0 || N <= 0 || N % 128) return fail(LQPB_E_ARG, "bad argument");
  int rc = check_device();
  if (rc) return rc;
  CK(launch_tc_dev_inverse(B, N, A, Ainv, (float*)work, (cudaStream_t)stream), "tc_dev_inverse");
  return LQPB_OK;
}

}  // extern "C"
