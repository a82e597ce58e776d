// K3 -- the persistent ADMM iteration kernel, and K4 -- finalisation.
//
// Restates the loop of the reference forward solver, lqp_py/solve_box_qp_admm_torch.py:235-313:
//   :237-256  adaptive-rho decision at the top of iteration i from the residuals of the previous check
//             (the refactorisation itself is done by the host between two launches of this kernel)
//   :259-268  rhs = -p~ + rho (z - u);  x = M^-1 [rhs; b~]   ->  here  x = K11 rhs + c  (factor.cu)
//   :271-282  z = clamp(x + u), r = x - z, s = rho (z - z_prev), u += r
//   :285-313  every check_solved iterations: six inf-norms (one needs Q~ x), tolerances, and the
//             GLOBAL stop test "all problems optimal"
//   :327      nus = (last KKT solve)[n:] * E   ->  here  nu = S^-1 (G^T rhs - b~) * E
//
// Design: one persistent CTA per problem (problems are strided over the grid when B exceeds the
// number of resident CTAs).  A dedicated producer warp streams the symmetric operator K11 -- the
// only O(n^2) data of an iteration -- from HBM into a shared-memory ring with 1-D bulk TMA copies
// (cp.async.bulk + mbarrier complete_tx); it runs ahead across iteration boundaries because K11
// does not change, so HBM never idles while the consumers do the O(n) vector update.  Because K11
// is symmetric, a row panel is also a column panel: consumer thread (g, t) owns 16 bytes of
// columns and accumulates  x[cols] += K[r][cols] * rhs[r]  over the panel rows r = g mod NG --
// conflict-free 16-byte shared loads, no shuffles; the NG partial sums are combined once per
// iteration.  All problems advance in lock step and stop together: every check_solved
// iterations each CTA publishes its flags, a grid barrier (atomic counter) makes the decision
// global, exactly like torch.all(is_optimal) in the reference.  No host round trip per iteration.
#include "layout.cuh"

namespace lqpb {

constexpr int kIterConsumers = 512;
constexpr int kIterThreads = kIterConsumers + 32;   // + one producer warp
constexpr int kConsBar = 1;                         // named barrier of the consumer threads
constexpr int kMaxChunks = 4;                       // max 16-byte column chunks per consumer thread

struct IterGeom {
  int tpr;          // threads per panel row (each owns `cpt` 16-byte chunks of columns)
  int ng;           // row groups; group g owns the `rpg` contiguous panel rows [g*rpg, (g+1)*rpg)
  int cpt;          // chunks per thread (template parameter of the kernel)
  int rpg;          // rows per group and panel (1..8, dispatched to unrolled code)
  int rows;         // rows per panel (stage) = ng * rpg; every panel is full: the last one is shifted
                    // up to end at row n and the rows it shares with its predecessor are masked out
  int panels;       // panels per matrix
  int stages;       // ring depth
  int stage_elems;  // elements per stage
};

// One panel of the column sweep: acc[c][:] += sum_rr K[g*rpg + rr][cols_c] * v[g*rpg + rr].
template <typename T, int CPT, int RPG>
__device__ __forceinline__ void sweep_panel(const T* __restrict__ P, const T* __restrict__ vp, int skip, int row_base,
                                            int ld, int col0, int col_stride, T (&acc)[CPT][Vec<T>::N]) {
  constexpr int VN = Vec<T>::N;
  using V4 = typename Vec<T>::type;
  T vr[RPG];
#pragma unroll
  for (int rr = 0; rr < RPG; ++rr) vr[rr] = (row_base + rr >= skip) ? vp[row_base + rr] : T(0);
#pragma unroll
  for (int rr = 0; rr < RPG; ++rr) {
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const int col = col0 + c * col_stride;
      if (CPT == 1 || col < ld) {
        const V4 kv = *reinterpret_cast<const V4*>(P + (size_t)(row_base + rr) * ld + col);
        const T* kp = reinterpret_cast<const T*>(&kv);
#pragma unroll
        for (int e = 0; e < VN; ++e) acc[c][e] += kp[e] * vr[rr];
      }
    }
  }
}

template <typename T, int CPT>
__global__ void __launch_bounds__(kIterThreads, 1)
iterate_kernel(lqpb_config cfg, FwdWs<T> w, int i0, int skip_rho_check, T* nus_out, IterGeom geo) {
  constexpr int VN = Vec<T>::N;
  using V4 = typename Vec<T>::type;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int n = w.n, m = w.m, ld = w.ld;
  T* ring = reinterpret_cast<T*>(smem_raw);
  T* part = ring + (size_t)geo.stages * geo.stage_elems;   // [ng][ld]
  T* v = part + (size_t)geo.ng * ld;                        // [ld] rhs of the x-update
  T* xs = v + ld;                                           // [ld] x~ of this iteration
  T* Ds = xs + ld;                                          // [ld] D
  T* tdot = Ds + ld;                                        // [max(m,1)] G^T rhs
  T* red = tdot + (m > 0 ? round_up(m, 4) : 4);             // [6][16] reduction scratch
  uint64_t* full = reinterpret_cast<uint64_t*>(red + 6 * 16 + 4);
  uint64_t* empty = full + geo.stages;
  __shared__ int s_dec[4];

  const int tid = threadIdx.x;
  const int wid = tid >> 5, lane = tid & 31;
  const bool is_producer = wid == (kIterConsumers >> 5);
  const int nprob = (w.B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  Ctrl* ctrl = w.ctrl;

  if (tid == 0) {
    for (int s = 0; s < geo.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kIterConsumers >> 5);
    }
    fence_mbar_init();
  }
  __syncthreads();

  const bool any_lb = ctrl->any_lb != 0, any_ub = ctrl->any_ub != 0;
  int last_wants = ctrl->last_wants, last_rout = ctrl->last_ratio_out;
  const int check = cfg.check_solved;
  const T eps_abs = (T)cfg.eps_abs, eps_rel = (T)cfg.eps_rel, zc = (T)cfg.zero_clamp;
  const T thr = (T)cfg.adaptive_rho_threshold, ar_tol = (T)cfg.adaptive_rho_tol, ar_tol_inv = (T)(1.0 / cfg.adaptive_rho_tol);

  // consumer thread geometry
  const int g = tid / geo.tpr, t = tid % geo.tpr;
  const bool gemv_active = !is_producer && g < geo.ng && t * VN < ld;
  const int nwc = kIterConsumers >> 5;
  const int R = geo.rows;
  const int last_start = n - R;      // first row of the (shifted) last panel

  int stage = 0;            // ring position and phase: same schedule in producer and consumers
  uint32_t phase = 0;
  bool have_v = false;      // v already holds the rhs of this iteration (single-problem CTAs)
  int i = i0;
  int status = 0;
  unsigned barrier_epoch = 0;

  while (true) {
    // ---------------- adaptive rho (:237-256): decided from the previous check, applied before iteration i
    if (cfg.adaptive_rho && i > 0 && i < cfg.adaptive_rho_max_iter && (i % cfg.adaptive_rho_iter) == 0 &&
        !(i == i0 && skip_rho_check)) {
      if (last_wants && last_rout) {
        if (!is_producer) {
          for (int k = tid; k < nprob; k += kIterConsumers) {
            const int b = blockIdx.x + k * gridDim.x;
            if (w.wants[b]) {
              T r = w.rho[b] * w.ratio[b];
              r = t_min(t_max(r, (T)cfg.rho_min), (T)cfg.rho_max);
              w.rho[b] = r;
            }
          }
        }
        status = 3;
        break;
      }
    }
    const bool is_check = (i % check) == 0;
    const bool is_last = i == cfg.max_iters - 1;
    const bool maybe_final = is_check || is_last;

    if (is_producer) {
      // ======================= producer warp: stream K11 (and Q~ at checks) through the ring
      if (lane == 0) {
        for (int k = 0; k < nprob; ++k) {
          const int b = blockIdx.x + k * gridDim.x;
          for (int pass = 0; pass < (is_check ? 2 : 1); ++pass) {
            const T* src = (pass == 0 ? w.K : w.Qs) + (size_t)b * n * ld;
            const uint32_t bytes = (uint32_t)(R * ld * sizeof(T));
            for (int pn = 0; pn < geo.panels; ++pn) {
              mbar_wait(&empty[stage], phase ^ 1u);
              const int start = min(pn * R, last_start);
              mbar_arrive_expect_tx(&full[stage], bytes);
              tma_load_1d(ring + (size_t)stage * geo.stage_elems, src + (size_t)start * ld, bytes, &full[stage]);
              if (++stage == geo.stages) { stage = 0; phase ^= 1u; }
            }
          }
        }
      }
      __syncwarp();
    } else {
      // ======================= consumers
      int cta_notopt = 0, cta_wants = 0, cta_rout = 0;
      for (int k = 0; k < nprob; ++k) {
        const int b = blockIdx.x + k * gridDim.x;
        const size_t vo = (size_t)b * ld;
        const T rho = w.rho[b];
        if (!have_v) {
          for (int e = tid; e < ld; e += kIterConsumers)
            v[e] = e < n ? -w.pt[vo + e] + rho * (w.z[vo + e] - w.u[vo + e]) : T(0);
          bar_sync(kConsBar, kIterConsumers);
        }
        // ---- x~ = K11 v : column sweep over the streamed row panels
        T acc[CPT][VN];
#pragma unroll
        for (int c = 0; c < CPT; ++c)
#pragma unroll
          for (int e = 0; e < VN; ++e) acc[c][e] = T(0);
        {
          const int row_base = g * geo.rpg, col0 = t * VN, cstride = geo.tpr * VN;
          for (int pn = 0; pn < geo.panels; ++pn) {
            mbar_wait(&full[stage], phase);
            if (gemv_active) {
              const int start = min(pn * R, last_start);
              const int skip = pn * R - start;          // rows already covered by the previous panel
              const T* P = ring + (size_t)stage * geo.stage_elems;
              const T* vp = v + start;
              switch (geo.rpg) {
                case 1: sweep_panel<T, CPT, 1>(P, vp, skip, row_base, ld, col0, cstride, acc); break;
                case 2: sweep_panel<T, CPT, 2>(P, vp, skip, row_base, ld, col0, cstride, acc); break;
                case 3: sweep_panel<T, CPT, 3>(P, vp, skip, row_base, ld, col0, cstride, acc); break;
                case 4: sweep_panel<T, CPT, 4>(P, vp, skip, row_base, ld, col0, cstride, acc); break;
                case 5: sweep_panel<T, CPT, 5>(P, vp, skip, row_base, ld, col0, cstride, acc); break;
                case 6: sweep_panel<T, CPT, 6>(P, vp, skip, row_base, ld, col0, cstride, acc); break;
                case 7: sweep_panel<T, CPT, 7>(P, vp, skip, row_base, ld, col0, cstride, acc); break;
                default: sweep_panel<T, CPT, 8>(P, vp, skip, row_base, ld, col0, cstride, acc); break;
              }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);
            if (++stage == geo.stages) { stage = 0; phase ^= 1u; }
          }
        }
        if (gemv_active) {
#pragma unroll
          for (int c = 0; c < CPT; ++c) {
            const int col = (t + c * geo.tpr) * VN;
            if (CPT == 1 || col < ld) {
#pragma unroll
              for (int e = 0; e < VN; ++e) part[(size_t)g * ld + col + e] = acc[c][e];
            }
          }
        }
        bar_sync(kConsBar, kIterConsumers);
        // ---- K21 rhs for nu (:327), from the rhs of THIS solve (before v is overwritten)
        if (maybe_final && m > 0) {
          const T* Gt = w.Gt + (size_t)b * m * ld;
          for (int l = wid; l < m; l += nwc) {
            T d = T(0);
            for (int e = lane; e < n; e += 32) d += Gt[(size_t)l * ld + e] * v[e];
            d = warp_sum(d);
            if (lane == 0) tdot[l] = d;
          }
          bar_sync(kConsBar, kIterConsumers);
        }
        // ---- element-wise ADMM update (:271-282) and the rhs of the next iteration (:259-262)
        T mx_p = T(0), mx_d = T(0), mx_x = T(0), mx_z = T(0), mx_y = T(0);
        for (int e = tid; e < n; e += kIterConsumers) {
          T x = T(0);
          for (int gg = 0; gg < geo.ng; ++gg) x += part[(size_t)gg * ld + e];
          x += w.c[vo + e];
          const T z_prev = w.z[vo + e], u_prev = w.u[vo + e];
          T zn = x + u_prev;
          if (any_lb) zn = t_max(zn, w.lbt[vo + e]);
          if (any_ub) zn = t_min(zn, w.ubt[vo + e]);
          const T r = x - zn;
          const T sres = rho * (zn - z_prev);
          const T un = u_prev + r;
          w.z[vo + e] = zn;
          w.u[vo + e] = un;
          v[e] = -w.pt[vo + e] + rho * (zn - un);
          if (maybe_final) {
            xs[e] = x;
            w.xs[vo + e] = x;
          }
          if (is_check) {
            const T d = w.D[vo + e];
            Ds[e] = d;
            mx_p = t_max(mx_p, t_abs(d * r));
            mx_d = t_max(mx_d, t_abs(d * sres));
            mx_x = t_max(mx_x, t_abs(d * x));
            mx_z = t_max(mx_z, t_abs(d * zn));
            mx_y = t_max(mx_y, t_abs(rho * d * un));
          }
        }
        if (maybe_final)
          for (int e = n + tid; e < ld; e += kIterConsumers) xs[e] = T(0);
        have_v = (nprob == 1);
        bar_sync(kConsBar, kIterConsumers);
        if (maybe_final && m > 0 && tid < m) {     // nu = K21 rhs + K22 b~, unscaled by E (:327)
          const T* K22 = w.Sinv + (size_t)b * m * m;
          T a = tdot[tid];
          for (int l = 0; l < m; ++l) a += K22[tid * m + l] * w.bt[(size_t)b * m + l];
          nus_out[(size_t)b * m + tid] = a * w.E[(size_t)b * m + tid];
        }
        if (is_check) {
          // ---- ||Q~ x~ / D||_inf (:299): row dots over the streamed Q~ panels, one warp per row
          T mx_q = T(0);
          for (int pn = 0; pn < geo.panels; ++pn) {
            mbar_wait(&full[stage], phase);
            const int start = min(pn * R, last_start);
            const int skip = pn * R - start;
            const T* P = ring + (size_t)stage * geo.stage_elems;
            for (int r = skip + wid; r < R; r += nwc) {
              T d = T(0);
              for (int col = lane * VN; col < ld; col += 32 * VN) {
                const V4 kv = *reinterpret_cast<const V4*>(P + (size_t)r * ld + col);
                const V4 xv = *reinterpret_cast<const V4*>(xs + col);
                const T* kp = reinterpret_cast<const T*>(&kv);
                const T* xp = reinterpret_cast<const T*>(&xv);
#pragma unroll
                for (int e = 0; e < VN; ++e) d += kp[e] * xp[e];
              }
              d = warp_sum(d);
              mx_q = t_max(mx_q, t_abs(d / Ds[start + r]));
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);
            if (++stage == geo.stages) { stage = 0; phase ^= 1u; }
          }
          // ---- block reduction of the six maxima
          mx_p = warp_max(mx_p); mx_d = warp_max(mx_d); mx_x = warp_max(mx_x);
          mx_z = warp_max(mx_z); mx_y = warp_max(mx_y); mx_q = warp_max(mx_q);
          if (lane == 0) {
            red[0 * 16 + wid] = mx_p; red[1 * 16 + wid] = mx_d; red[2 * 16 + wid] = mx_x;
            red[3 * 16 + wid] = mx_z; red[4 * 16 + wid] = mx_y; red[5 * 16 + wid] = mx_q;
          }
          bar_sync(kConsBar, kIterConsumers);
          if (tid == 0) {
            T mm[6];
            for (int a = 0; a < 6; ++a) {
              T r = red[a * 16];
              for (int ww = 1; ww < nwc; ++ww) r = t_max(r, red[a * 16 + ww]);
              mm[a] = r;
            }
            const T primal = mm[0], dual = mm[1];
            const T tol_p_rel = t_max(t_max(mm[2], mm[3]), zc);                      // :301
            const T tol_p = eps_abs + eps_rel * tol_p_rel;                           // :302
            const T tol_d_rel = t_max(t_max(t_max(mm[4], mm[5]), w.pnorm[b]), zc);   // :303
            const T tol_d = eps_abs + eps_rel * tol_d_rel;                           // :304
            const bool optimal = (primal < tol_p) && (dual < tol_d);                // :307-309
            const bool wants = (primal > t_max(tol_p, thr)) || (dual > t_max(tol_d, thr));   // :310-311
            const T num = t_max(primal / tol_p_rel, zc), den = t_max(dual / tol_d_rel, zc);  // :239-242
            const T ratio = t_sqrt(num / den);                                       // :243
            w.chk[4 * b + 0] = primal; w.chk[4 * b + 1] = dual;
            w.chk[4 * b + 2] = tol_p_rel; w.chk[4 * b + 3] = tol_d_rel;
            w.wants[b] = wants ? 1 : 0;
            w.ratio[b] = ratio;
            cta_notopt += optimal ? 0 : 1;
            cta_wants |= wants ? 1 : 0;
            cta_rout |= (ratio > ar_tol || ratio < ar_tol_inv) ? 1 : 0;              // :244-245
            if (cfg.verbose) {
              const int ci = i / check;
              if (ci < LQPB_LOG_CAP) {
                atomic_max_nonneg(&ctrl->log_primal[ci], (double)primal);
                atomic_max_nonneg(&ctrl->log_dual[ci], (double)dual);
                ctrl->log_iter[ci] = i;
              }
            }
          }
          bar_sync(kConsBar, kIterConsumers);   // red[] reusable
        }
      }
      // ---- publish this CTA's flags and make the decision global (:312 torch.all)
      if (is_check && tid == 0) {
        int* slot = ctrl->slot[(i / check) & 3];
        if (cta_notopt) atomicAdd(&slot[0], cta_notopt);
        if (cta_wants) atomicOr(&slot[1], 1);
        if (cta_rout) atomicOr(&slot[2], 1);
        __threadfence();
        atomicAdd(&ctrl->barrier, 1u);
        const unsigned target = (barrier_epoch + 1) * gridDim.x;
        while (ld_acquire_u32(&ctrl->barrier) < target) {
        }
        __threadfence();
        s_dec[0] = *(volatile int*)&slot[0];
        s_dec[1] = *(volatile int*)&slot[1];
        s_dec[2] = *(volatile int*)&slot[2];
        if (blockIdx.x == 0) {
          int* nxt = ctrl->slot[((i / check) + 2) & 3];
          nxt[0] = 0; nxt[1] = 0; nxt[2] = 0;
          ctrl->last_wants = s_dec[1];
          ctrl->last_ratio_out = s_dec[2];
          if (cfg.verbose) ctrl->n_log = min(i / check + 1, LQPB_LOG_CAP);
          __threadfence();
        }
      }
    }
    if (is_check) {
      ++barrier_epoch;
      __syncthreads();
      const int notopt = s_dec[0];
      last_wants = s_dec[1];
      last_rout = s_dec[2];
      __syncthreads();
      if (notopt == 0) { status = 1; break; }
    }
    if (is_last) { status = 2; break; }
    ++i;
  }
  if (blockIdx.x == 0 && tid == 0) {
    ctrl->status = status;
    if (status == 3) ctrl->next_i = i;
    else ctrl->iter = i;
  }
}

// ---------------------------------------------------------------------------------------------
// K4: undo the scaling and split the duals (:315-331).  No-op while a refactorisation is pending.
template <typename T>
__global__ void finalize_kernel(FwdWs<T> w, T* x, T* z, T* u, T* lams, T* rho_out) {
  if (w.ctrl->status == 3) return;
  const int b = blockIdx.y;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= w.n) return;
  const size_t vo = (size_t)b * w.ld + e, o = (size_t)b * w.n + e;
  const T d = w.D[vo], rho = w.rho[b];
  x[o] = d * w.xs[vo];
  z[o] = d * w.z[vo];
  const T uu = w.u[vo] / d;
  u[o] = uu;
  const T y = uu * rho;
  lams[(size_t)b * 2 * w.n + e] = (-y > T(0)) ? -y : T(0);
  lams[(size_t)b * 2 * w.n + w.n + e] = (y > T(0)) ? y : T(0);
  if (e == 0) rho_out[b] = rho;
}

template <typename T>
static IterGeom make_geom(const FwdWs<T>& w, size_t* smem_bytes, int max_smem) {
  IterGeom g{};
  const int vn = Vec<T>::N;
  const int chunks = w.ld / vn;
  int tpr;
  if (chunks <= 32) { tpr = 1; while (tpr < chunks) tpr <<= 1; }
  else tpr = round_up(chunks, 32);
  if (tpr > kIterConsumers) tpr = kIterConsumers;
  g.tpr = tpr;
  g.cpt = (chunks + tpr - 1) / tpr;
  const size_t row_bytes = (size_t)w.ld * sizeof(T);
  int target = (int)(40960 / row_bytes);          // ~40 KB stages
  if (target < 1) target = 1;
  if (target > w.n) target = w.n;
  int ng = kIterConsumers / tpr;
  if (ng > 16) ng = 16;
  while (ng > target) ng >>= 1;                    // power of two <= target rows
  if (ng < 1) ng = 1;
  g.ng = ng;
  int rpg = target / ng;
  if (rpg > 8) rpg = 8;
  if (rpg < 1) rpg = 1;
  // prefer a panel height that divides n (no shifted last panel), searching a little below the target
  for (int cand = rpg; cand >= 1 && cand >= rpg - 2; --cand)
    if (w.n % (cand * ng) == 0) { rpg = cand; break; }
  g.rpg = rpg;
  g.rows = rpg * ng;
  g.panels = (w.n + g.rows - 1) / g.rows;
  g.stage_elems = (int)(round_up_sz((size_t)g.rows * row_bytes, 128) / sizeof(T));
  const size_t fixed = ((size_t)g.ng * w.ld + 3 * (size_t)w.ld + (w.m > 0 ? round_up(w.m, 4) : 4) + 6 * 16 + 4) * sizeof(T) +
                       2 * 16 * sizeof(uint64_t) + 256;
  const size_t stage_bytes = (size_t)g.stage_elems * sizeof(T);
  int stages = (int)(((size_t)max_smem - fixed) / stage_bytes);
  if (stages > 8) stages = 8;
  g.stages = stages;
  *smem_bytes = fixed + stages * stage_bytes;
  return g;
}

template <typename T, int CPT>
static cudaError_t launch_iterate_cpt(const lqpb_config& cfg, const FwdWs<T>& w, int i0, int skip_rho_check,
                                      T* nus_out, IterGeom geo, size_t smem, int grid, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(iterate_kernel<T, CPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  lqpb_config c = cfg;
  FwdWs<T> ww = w;
  void* args[] = {&c, &ww, &i0, &skip_rho_check, &nus_out, &geo};
  return cudaLaunchCooperativeKernel((void*)iterate_kernel<T, CPT>, dim3(grid), dim3(kIterThreads), args, smem, st);
}

template <typename T>
cudaError_t launch_iterate(const lqpb_config& cfg, const FwdWs<T>& w, int i0, int skip_rho_check, T* nus_out,
                           int* launches, cudaStream_t st) {
  int dev = 0, max_smem = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  size_t smem = 0;
  IterGeom geo = make_geom(w, &smem, max_smem - 1024);
  if (geo.stages < 2 || geo.cpt > kMaxChunks) return cudaErrorInvalidConfiguration;
  const int grid = w.B < sms ? w.B : sms;   // one CTA per SM: all CTAs co-resident (needed by the grid barrier)
  e = cudaMemsetAsync(&w.ctrl->barrier, 0, sizeof(unsigned), st);
  if (e != cudaSuccess) return e;
  if (geo.cpt == 1) e = launch_iterate_cpt<T, 1>(cfg, w, i0, skip_rho_check, nus_out, geo, smem, grid, st);
  else if (geo.cpt == 2) e = launch_iterate_cpt<T, 2>(cfg, w, i0, skip_rho_check, nus_out, geo, smem, grid, st);
  else e = launch_iterate_cpt<T, 4>(cfg, w, i0, skip_rho_check, nus_out, geo, smem, grid, st);
  if (launches) ++*launches;
  return e;
}

template <typename T>
cudaError_t launch_finalize(const FwdWs<T>& w, T* x, T* z, T* u, T* lams, T* rho_out, cudaStream_t st) {
  dim3 grid((w.n + 127) / 128, w.B);
  finalize_kernel<T><<<grid, 128, 0, st>>>(w, x, z, u, lams, rho_out);
  return cudaGetLastError();
}

#define INST(T)                                                                                            \
  template cudaError_t launch_iterate<T>(const lqpb_config&, const FwdWs<T>&, int, int, T*, int*, cudaStream_t); \
  template cudaError_t launch_finalize<T>(const FwdWs<T>&, T*, T*, T*, T*, T*, cudaStream_t);
INST(float)
INST(double)
#undef INST

}  // namespace lqpb
