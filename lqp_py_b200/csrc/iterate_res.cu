// K3, resident regime -- the ADMM iteration kernel for problems whose x-update operator fits in shared memory.
//
// Same loop as iterate.cu (reference lqp_py/solve_box_qp_admm_torch.py:235-313: x-update, clamp, dual update,
// residual norms, GLOBAL stop test, adaptive-rho trigger; :327 nus), same packed operators, same arithmetic per
// problem -- what changes is where the operator lives.  SURVEY App. C names three on-chip regimes for the iteration
// operator; iterate.cu is the streaming one (dz >= ~350 in fp32: the operator set of the batch is re-read from L2 /
// HBM every iteration through per-warp bulk-TMA rings).  Here the packed K11 of every problem a CTA owns is copied into
// shared memory ONCE (bulk TMA, one mbarrier) and stays there for the whole solve, together with the problem's
// vectors (rhs, z, u, p~, lb~, ub~, c, D, x~); the scaled Q~ needed by the stop checks stays resident too when it fits
// and is otherwise read straight from L2 at the checks.  An iteration then costs shared-memory bandwidth only.
//
// Work split: a CTA has 16 warps in `ngroups` groups of `gw` warps; a group owns whole problems (slot q of the CTA is
// problem blockIdx.x + q * gridDim.x, group g takes slots g, g + ngroups, ...) and synchronises with its own named
// barrier (or __syncwarp when gw == 1), so small problems advance independently inside a CTA -- dz = 10 runs one
// problem per warp -- and the CTA only meets at the stop checks.  All problems still advance in lock step and stop
// together (reference :312): at a check every CTA publishes its flags and waits at a grid-wide barrier.  When the
// whole batch fits the shared memory of <= 8 CTAs the grid is launched as ONE thread-block cluster and that barrier is
// the hardware cluster barrier (barrier.cluster, ~0.2 us) instead of a round trip through L2 atomics -- which matters
// at dz < 25, where the reference checks after every single iteration.
#include <cstring>
#include "itergeom.cuh"

namespace lqpb {

#ifdef LQPB_PHASE_TIMERS
__device__ long long g_res_cycles[16];
#define RT0() long long rt__ = clock64()
#define RADD(k) do { long long n__ = clock64(); if (blockIdx.x == 0 && threadIdx.x == 0) g_res_cycles[k] += n__ - rt__; rt__ = n__; } while (0)
#else
#define RT0()
#define RADD(k)
#endif

constexpr int kResWarps = 16;
constexpr int kResThreads = kResWarps * 32;
constexpr int kResVecs = 9;          // v, z, u, p~, lb~, ub~, c, D, x~ per resident problem

struct ResGeom {
  int G;          // CTAs
  int ppc;        // resident problem slots per CTA
  int gw;         // warps per group
  int ngroups;    // groups per CTA
  int qres;       // 1: Q~ resident as well
  int cluster;    // 1: the grid is ONE cluster (barrier.cluster); 0: cooperative launch, atomic grid barrier
  int nt, nbc, ntiles, np;
  int mpad;       // K21 rhs scratch entries per group
  size_t prob_elems;   // elements per problem block
  size_t group_elems;  // elements per group scratch block
};

__device__ __forceinline__ void cluster_arrive_release() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait_acquire() {
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ int ld_acquire_s32(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <typename T>
__global__ void __launch_bounds__(kResThreads, 1)
iterate_res_kernel(lqpb_config cfg, FwdWs<T> w, int i0, int skip_rho_check, T* nus_out, ResGeom geo) {
  using P = Pack<T>;
  constexpr int TC = P::TC, TILE = P::TILE;
  using V4 = typename Vec<T>::type;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t load_bar;
  __shared__ int s_dec[4];
  __shared__ int s_flags[4];
  const int n = w.n, m = w.m, ld = w.ld, np = geo.np;
  const int ntv = geo.nt, ntiles = geo.ntiles, gw = geo.gw, ngroups = geo.ngroups;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int grp = wid / gw, wg = wid % gw;          // group, warp inside the group
  const int gtid = wg * 32 + lane, gthreads = gw * 32;
  const int nprob = (w.B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // slots in use (<= ppc)
  Ctrl* ctrl = w.ctrl;

  T* base = reinterpret_cast<T*>(smem_raw);
  auto prob_K = [&](int q) { return base + (size_t)q * geo.prob_elems; };
  auto prob_Q = [&](int q) { return prob_K(q) + (size_t)ntiles * TILE; };
  auto prob_vec = [&](int q, int k) { return prob_K(q) + (size_t)(1 + geo.qres) * ntiles * TILE + (size_t)k * np; };
  T* gscr = base + (size_t)geo.ppc * geo.prob_elems + (size_t)grp * geo.group_elems;
  T* xpart = gscr;                                  // [gw][np]
  T* tdot = xpart + (size_t)gw * np;                // [mpad]
  T* red = tdot + geo.mpad;                         // [6][16]
  T* pscal = base + (size_t)geo.ppc * geo.prob_elems + (size_t)ngroups * geo.group_elems;   // [ppc][2]: rho, pnorm

  // ---- one-time load: operators by bulk TMA, vectors by plain loads
  if (tid == 0) {
    mbar_init(&load_bar, 1);
    fence_mbar_init();
    s_flags[0] = s_flags[1] = s_flags[2] = s_flags[3] = 0;
  }
  __syncthreads();
  if (tid == 0) {
    const uint32_t total = (uint32_t)((size_t)nprob * (1 + geo.qres) * ntiles * TILE * sizeof(T));
    mbar_arrive_expect_tx(&load_bar, total);
  }
  __syncthreads();
  for (int t = tid; t < nprob * ntiles; t += kResThreads) {
    const int q = t / ntiles, tl = t % ntiles;
    const int b = blockIdx.x + q * gridDim.x;
    tma_load_1d(prob_K(q) + (size_t)tl * TILE, w.Kp + ((size_t)b * ntiles + tl) * TILE, (uint32_t)(TILE * sizeof(T)),
                &load_bar);
    if (geo.qres)
      tma_load_1d(prob_Q(q) + (size_t)tl * TILE, w.Qp + ((size_t)b * ntiles + tl) * TILE, (uint32_t)(TILE * sizeof(T)),
                  &load_bar);
  }
  for (int t = tid; t < nprob * np; t += kResThreads) {
    const int q = t / np, e = t % np;
    const int b = blockIdx.x + q * gridDim.x;
    const size_t vo = (size_t)b * ld + e;
    const bool in = e < n;
    const T rho = w.rho[b];
    const T z = in ? w.z[vo] : T(0), u = in ? w.u[vo] : T(0), pt = in ? w.pt[vo] : T(0);
    prob_vec(q, 0)[e] = in ? -pt + rho * (z - u) : T(0);     // rhs of the first iteration (:259-262)
    prob_vec(q, 1)[e] = z;
    prob_vec(q, 2)[e] = u;
    prob_vec(q, 3)[e] = pt;
    prob_vec(q, 4)[e] = in ? w.lbt[vo] : T(0);
    prob_vec(q, 5)[e] = in ? w.ubt[vo] : T(0);
    prob_vec(q, 6)[e] = in ? w.c[vo] : T(0);
    prob_vec(q, 7)[e] = in ? w.D[vo] : T(1);
    prob_vec(q, 8)[e] = T(0);
  }
  for (int t = tid; t < ngroups * gw * np; t += kResThreads)
    (base + (size_t)geo.ppc * geo.prob_elems + (size_t)(t / (gw * np)) * geo.group_elems)[t % (gw * np)] = T(0);
  for (int q = tid; q < nprob; q += kResThreads) {
    const int b = blockIdx.x + q * gridDim.x;
    pscal[2 * q] = w.rho[b];
    pscal[2 * q + 1] = w.pnorm[b];
  }
  mbar_wait(&load_bar, 0);
  __syncthreads();

  const bool any_lb = ctrl->any_lb != 0, any_ub = ctrl->any_ub != 0;
  int last_wants = ctrl->last_wants, last_rout = ctrl->last_ratio_out;
  const int check = cfg.check_solved;
  const T eps_abs = (T)cfg.eps_abs, eps_rel = (T)cfg.eps_rel, zc = (T)cfg.zero_clamp;
  const T thr = (T)cfg.adaptive_rho_threshold, ar_tol = (T)cfg.adaptive_rho_tol, ar_tol_inv = (T)(1.0 / cfg.adaptive_rho_tol);

  // ---- this warp's run of tiles inside a problem's tile sequence (the same for every matrix)
  const int run_lo = (int)((long long)wg * ntiles / gw);
  const int run_len = (int)((long long)(wg + 1) * ntiles / gw) - run_lo;
  int Jc_first = 0, I_first = 0;
  {
    int rem = run_lo;
    while (Jc_first < geo.nbc && rem >= ntv - Jc_first / P::R) { rem -= ntv - Jc_first / P::R; ++Jc_first; }
    I_first = Jc_first / P::R + rem;
  }
  T* const xp = xpart + (size_t)wg * np;
  auto group_sync = [&]() {
    if (gw == 1) __syncwarp();
    else bar_sync(1 + grp, gthreads);
  };

  // one symmetric pass over the warp's run of tiles starting at `tiles` (shared memory, or global memory for a Q~
  // that is not resident):  xp += (this warp's share of)  S vec
  auto sym_pass = [&](const T* tiles, const T* vec, bool global_src) {
    if (run_len == 0) return;
    int Jc = Jc_first, I = I_first;
    SymAcc<T> sa;
    auto flush_cols = [&]() {
      const T tot = sa.reduce(lane);
      __syncwarp();
      if (lane < TC) xp[Jc * TC + lane] += tot;
      __syncwarp();
    };
    sa.load_vJ(vec + Jc * TC);
    bool dirty = false;
    for (int r = 0; r < run_len; ++r) {
      const T* tp = tiles + (size_t)(run_lo + r) * TILE;
      V4 kv[8];
      if (global_src) {
#pragma unroll
        for (int k = 0; k < 8; ++k) kv[k] = __ldg(reinterpret_cast<const V4*>(tp + P::in_tile(lane, k)));
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) kv[k] = *reinterpret_cast<const V4*>(tp + P::in_tile(lane, k));
      }
      const T vI = vec[I * kPackRows + lane];
      xp[I * kPackRows + lane] += sa.apply(kv, vI);
      __syncwarp();
      dirty = true;
      if (++I == ntv) {
        flush_cols();
        dirty = false;
        ++Jc;
        I = Jc / P::R;
        if (r + 1 < run_len) sa.load_vJ(vec + Jc * TC);
      }
    }
    if (dirty) flush_cols();
  };

  int i = i0;
  int status = 0;
  unsigned barrier_epoch = 0;

  while (true) {
    // ---------------- adaptive rho (:237-256): decided from the previous check, applied before iteration i
    if (cfg.adaptive_rho && i > 0 && i < cfg.adaptive_rho_max_iter && (i % cfg.adaptive_rho_iter) == 0 &&
        !(i == i0 && skip_rho_check)) {
      if (last_wants && last_rout) {
        for (int k = tid; k < nprob; k += kResThreads) {
          const int b = blockIdx.x + k * gridDim.x;
          if (w.wants[b]) {
            T r = w.rho[b] * w.ratio[b];
            r = t_min(t_max(r, (T)cfg.rho_min), (T)cfg.rho_max);
            w.rho[b] = r;
          }
        }
        status = 3;
        break;
      }
    }
    const bool is_check = (i % check) == 0;
    const bool is_last = i == cfg.max_iters - 1;
    const bool maybe_final = is_check || is_last;

    RT0();
    for (int q = grp; q < nprob; q += ngroups) {
      const int b = blockIdx.x + q * gridDim.x;
      T* v = prob_vec(q, 0);
      T* zs = prob_vec(q, 1);
      T* us = prob_vec(q, 2);
      const T* pts = prob_vec(q, 3);
      const T* lbs = prob_vec(q, 4);
      const T* ubs = prob_vec(q, 5);
      const T* cs = prob_vec(q, 6);
      const T* Ds = prob_vec(q, 7);
      T* xs = prob_vec(q, 8);
      const T rho = pscal[2 * q];
      // ---- x~ = K11 v (+ c below)
      sym_pass(prob_K(q), v, false);
      RADD(0);
      group_sync();
      RADD(1);
      // ---- K21 rhs for nu (:327), from the rhs of THIS solve (before v is overwritten)
      if (maybe_final && m > 0) {
        const T* Gt = w.Gt + (size_t)b * m * ld;
        for (int l = wg; l < m; l += gw) {
          T d = T(0);
          for (int e = lane; e < n; e += 32) d += Gt[(size_t)l * ld + e] * v[e];
          d = warp_sum(d);
          if (lane == 0) tdot[l] = d;
        }
        group_sync();
      }
      RADD(2);
      // ---- element-wise ADMM update (:271-282) and the rhs of the next iteration (:259-262)
      T mx_p = T(0), mx_d = T(0), mx_x = T(0), mx_z = T(0), mx_y = T(0);
      for (int e = gtid; e < n; e += gthreads) {
        T x = T(0);
        for (int ww = 0; ww < gw; ++ww) {
          x += xpart[(size_t)ww * np + e];
          xpart[(size_t)ww * np + e] = T(0);
        }
        x += cs[e];
        const T z_prev = zs[e], u_prev = us[e];
        T zn = x + u_prev;
        if (any_lb) zn = t_max(zn, lbs[e]);
        if (any_ub) zn = t_min(zn, ubs[e]);
        const T r = x - zn;
        const T sres = rho * (zn - z_prev);
        const T un = u_prev + r;
        zs[e] = zn;
        us[e] = un;
        v[e] = -pts[e] + rho * (zn - un);
        xs[e] = x;
        if (is_check) {
          if (!(t_abs(x) < t_inf<T>())) s_flags[3] = 1;    // NaN / inf iterate: numerical breakdown (benign race: all write 1)
          const T d = Ds[e];
          mx_p = t_max(mx_p, t_abs(d * r));
          mx_d = t_max(mx_d, t_abs(d * sres));
          mx_x = t_max(mx_x, t_abs(d * x));
          mx_z = t_max(mx_z, t_abs(d * zn));
          mx_y = t_max(mx_y, t_abs(rho * d * un));
        }
      }
      RADD(3);
      group_sync();
      RADD(4);
      if (maybe_final && m > 0) {                 // nu = K21 rhs + K22 b~, unscaled by E (:327)
        const T* K22 = w.Sinv + (size_t)b * m * m;
        for (int r = gtid; r < m; r += gthreads) {
          T a = tdot[r];
          for (int l = 0; l < m; ++l) a += K22[r * m + l] * w.bt[(size_t)b * m + l];
          nus_out[(size_t)b * m + r] = a * w.E[(size_t)b * m + r];
        }
      }
      RADD(5);
      if (is_check) {
        // ---- ||Q~ x~ / D||_inf (:299): the same symmetric sweep over the packed Q~ tiles
        if (geo.qres) sym_pass(prob_Q(q), xs, false);
        else sym_pass(w.Qp + (size_t)b * ntiles * TILE, xs, true);
        RADD(6);
        group_sync();
        T mx_q = T(0);
        for (int e = gtid; e < n; e += gthreads) {
          T y = T(0);
          for (int ww = 0; ww < gw; ++ww) {
            y += xpart[(size_t)ww * np + e];
            xpart[(size_t)ww * np + e] = T(0);
          }
          mx_q = t_max(mx_q, t_abs(y / Ds[e]));
        }
        mx_p = warp_max(mx_p); mx_d = warp_max(mx_d); mx_x = warp_max(mx_x);
        mx_z = warp_max(mx_z); mx_y = warp_max(mx_y); mx_q = warp_max(mx_q);
        if (lane == 0) {
          red[0 * 16 + wg] = mx_p; red[1 * 16 + wg] = mx_d; red[2 * 16 + wg] = mx_x;
          red[3 * 16 + wg] = mx_z; red[4 * 16 + wg] = mx_y; red[5 * 16 + wg] = mx_q;
        }
        group_sync();
        if (gtid == 0) {
          T mm[6];
          for (int a = 0; a < 6; ++a) {
            T r = red[a * 16];
            for (int ww = 1; ww < gw; ++ww) r = t_max(r, red[a * 16 + ww]);
            mm[a] = r;
          }
          const T primal = mm[0], dual = mm[1];
          const T tol_p_rel = t_max(t_max(mm[2], mm[3]), zc);                      // :301
          const T tol_p = eps_abs + eps_rel * tol_p_rel;                           // :302
          const T tol_d_rel = t_max(t_max(t_max(mm[4], mm[5]), pscal[2 * q + 1]), zc);   // :303
          const T tol_d = eps_abs + eps_rel * tol_d_rel;                           // :304
          const bool optimal = (primal < tol_p) && (dual < tol_d);                // :307-309
          const bool wants = (primal > t_max(tol_p, thr)) || (dual > t_max(tol_d, thr));   // :310-311
          const T num = t_max(primal / tol_p_rel, zc), den = t_max(dual / tol_d_rel, zc);  // :239-242
          const T ratio = t_sqrt(num / den);                                       // :243
          w.chk[4 * b + 0] = primal; w.chk[4 * b + 1] = dual;
          w.chk[4 * b + 2] = tol_p_rel; w.chk[4 * b + 3] = tol_d_rel;
          w.wants[b] = wants ? 1 : 0;
          w.ratio[b] = ratio;
          if (!optimal) atomicAdd(&s_flags[0], 1);
          if (wants) atomicOr(&s_flags[1], 1);
          if (ratio > ar_tol || ratio < ar_tol_inv) atomicOr(&s_flags[2], 1);     // :244-245
          if (cfg.verbose) {
            const int ci = i / check;
            if (ci < LQPB_LOG_CAP) {
              atomic_max_nonneg(&ctrl->log_primal[ci], (double)primal);
              atomic_max_nonneg(&ctrl->log_dual[ci], (double)dual);
              ctrl->log_iter[ci] = i;
            }
          }
        }
        group_sync();   // red[] reusable
        RADD(7);
      }
    }
    // ---- publish this CTA's flags and make the decision global (:312 torch.all)
    if (is_check) {
      __syncthreads();
      int* slot = ctrl->slot[(i / check) & 3];
      if (tid == 0) {
        if (s_flags[0]) atomicAdd(&slot[0], s_flags[0]);
        if (s_flags[1]) atomicOr(&slot[1], 1);
        if (s_flags[2]) atomicOr(&slot[2], 1);
        if (s_flags[3]) atomicOr(&slot[3], 1);
        s_flags[0] = s_flags[1] = s_flags[2] = s_flags[3] = 0;
        __threadfence();
      }
      if (geo.cluster) {
        if (gridDim.x > 1) {
          cluster_arrive_release();
          cluster_wait_acquire();
        }
      } else if (tid == 0) {
        atomicAdd(&ctrl->barrier, 1u);
        const unsigned target = (barrier_epoch + 1) * gridDim.x;
        while (ld_acquire_u32(&ctrl->barrier) < target) {
        }
        __threadfence();
      }
      if (tid == 0) {
        s_dec[0] = ld_acquire_s32(&slot[0]);
        s_dec[1] = ld_acquire_s32(&slot[1]);
        s_dec[2] = ld_acquire_s32(&slot[2]);
        s_dec[3] = ld_acquire_s32(&slot[3]);
        if (blockIdx.x == 0) {
          int* nxt = ctrl->slot[((i / check) + 2) & 3];
          nxt[0] = 0; nxt[1] = 0; nxt[2] = 0; nxt[3] = 0;
          ctrl->last_wants = s_dec[1];
          ctrl->last_ratio_out = s_dec[2];
          if (cfg.verbose) ctrl->n_log = min(i / check + 1, LQPB_LOG_CAP);
          __threadfence();
        }
      }
      ++barrier_epoch;
      __syncthreads();
      RADD(8);
      const int notopt = s_dec[0];
      last_wants = s_dec[1];
      last_rout = s_dec[2];
      const int broken = s_dec[3];
      __syncthreads();
      if (broken) { status = 4; break; }           // LQPB_STATUS_BREAKDOWN
      if (notopt == 0) { status = 1; break; }
    }
    if (is_last) { status = 2; break; }
    ++i;
  }
  // ---- the state goes back to the workspace: finalize_kernel (and a relaunch after a refactorisation) read it there
  __syncthreads();
  for (int t = tid; t < nprob * n; t += kResThreads) {
    const int q = t / n, e = t % n;
    const int b = blockIdx.x + q * gridDim.x;
    const size_t vo = (size_t)b * ld + e;
    w.z[vo] = prob_vec(q, 1)[e];
    w.u[vo] = prob_vec(q, 2)[e];
    w.xs[vo] = prob_vec(q, 8)[e];
  }
  if (blockIdx.x == 0 && tid == 0) {
    ctrl->status = status;
    if (status == 3) ctrl->next_i = i;
    else ctrl->iter = i;
  }
}

// ---------------------------------------------------------------------------------------------
// Plan: which problems live where.  Returns false when the resident regime does not apply (operator too large for the
// shared memory of the SMs the batch can use) -- the caller then takes the streaming kernel of iterate.cu.
template <typename T>
bool plan_resident(const FwdWs<T>& w, const lqpb_config& cfg, int max_smem, int n_sm, ResGeom* out, size_t* smem_bytes) {
  using P = Pack<T>;
  ResGeom g{};
  g.nt = P::nt(w.n);
  g.nbc = P::nbc(w.n);
  g.ntiles = P::ntiles(w.n);
  g.np = kPackRows * g.nt;
  g.mpad = w.m > 0 ? round_up(w.m, 4) : 4;
  const size_t mat = (size_t)g.ntiles * P::TILE;          // elements
  const size_t budget = (size_t)max_smem / sizeof(T);     // elements
  // warps per group: about three tiles per warp and iteration, a power of two
  int gw = 1;
  while (gw < kResWarps && g.ntiles > 3 * gw) gw *= 2;
  auto fits = [&](int ppc, int gwx, int qres, size_t* total) {
    const int ng = kResWarps / gwx;
    const size_t prob = (size_t)(1 + qres) * mat + (size_t)kResVecs * g.np;
    const size_t grp = (size_t)gwx * g.np + g.mpad + 6 * 16;
    const size_t tot = (size_t)ppc * prob + (size_t)ng * grp + 2 * (size_t)ppc + 64;
    *total = tot;
    return tot <= budget;
  };
  // candidate grids: few CTAs as one cluster when the stop test runs (almost) every iteration, else one CTA per SM
  const bool chatty = cfg.check_solved <= 2;
  int grids[3];
  int ng = 0;
  if (chatty) {
    for (int G = 1; G <= 8; G *= 2)
      if (ng < 1) {
        size_t tot;
        const int ppc = (w.B + G - 1) / G;
        if (fits(ppc, gw, 1, &tot) || fits(ppc, gw, 0, &tot)) grids[ng++] = G;
      }
  }
  grids[ng++] = w.B < n_sm ? w.B : n_sm;
  for (int c = 0; c < ng; ++c) {
    const int G = grids[c];
    const int ppc = (w.B + G - 1) / G;
    int gwx = gw;
    while (gwx < kResWarps && kResWarps / gwx > ppc) gwx *= 2;      // no idle groups: fewer, wider groups
    for (int qres = 1; qres >= 0; --qres) {
      size_t tot;
      if (!fits(ppc, gwx, qres, &tot)) continue;
      g.G = G; g.ppc = ppc; g.gw = gwx; g.ngroups = kResWarps / gwx; g.qres = qres;
      g.cluster = (G <= 8) ? 1 : 0;
      g.prob_elems = (size_t)(1 + qres) * mat + (size_t)kResVecs * g.np;
      g.group_elems = (size_t)gwx * g.np + g.mpad + 6 * 16;
      *smem_bytes = tot * sizeof(T);
      *out = g;
      return true;
    }
  }
  return false;
}

template <typename T>
cudaError_t launch_iterate_split(const lqpb_config& cfg, const FwdWs<T>& w, int i0, int skip_rho_check, T* nus_out,
                                 int* launches, cudaStream_t st, bool* taken);      // iterate_split.cu

template <typename T>
cudaError_t launch_iterate_resident(const lqpb_config& cfg, const FwdWs<T>& w, int i0, int skip_rho_check, T* nus_out,
                                    int* launches, cudaStream_t st, bool* taken) {
  *taken = false;
  {
    const char* e = getenv("LQPB_ITER");          // developer switch (A/B measurements): stream = always iterate.cu
    if (e && (!strcmp(e, "stream") || !strcmp(e, "rows"))) return cudaSuccess;
  }
  int dev = 0, max_smem = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  ResGeom geo{};
  size_t smem = 0;
  if (!plan_resident(w, cfg, max_smem - 2048, sms, &geo, &smem))      // streamed regime: small batches split every
    return launch_iterate_split<T>(cfg, w, i0, skip_rho_check, nus_out, launches, st, taken);   // problem over a cluster
  e = cudaMemsetAsync(&w.ctrl->barrier, 0, sizeof(unsigned), st);
  if (e != cudaSuccess) return e;
  void* kern = (void*)iterate_res_kernel<T>;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  lqpb_config c = cfg;
  FwdWs<T> ww = w;
  if (geo.cluster) {
    cudaLaunchConfig_t lc{};
    lc.gridDim = dim3(geo.G);
    lc.blockDim = dim3(kResThreads);
    lc.dynamicSmemBytes = smem;
    lc.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = geo.G;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    lc.attrs = at;
    lc.numAttrs = 1;
    e = cudaLaunchKernelEx(&lc, iterate_res_kernel<T>, c, ww, i0, skip_rho_check, nus_out, geo);
  } else {
    void* args[] = {&c, &ww, &i0, &skip_rho_check, &nus_out, &geo};
    e = cudaLaunchCooperativeKernel(kern, dim3(geo.G), dim3(kResThreads), args, smem, st);
  }
  if (e != cudaSuccess) return e;
  if (launches) ++*launches;
  *taken = true;
  return cudaSuccess;
}

template <typename T>
bool iterate_resident_applies(const lqpb_config& cfg, const FwdWs<T>& w) {
  const char* e = getenv("LQPB_ITER");
  if (e && (!strcmp(e, "stream") || !strcmp(e, "rows"))) return false;
  int dev = 0, max_smem = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  ResGeom geo{};
  size_t smem = 0;
  return plan_resident(w, cfg, max_smem - 2048, sms, &geo, &smem);
}
template bool iterate_resident_applies<float>(const lqpb_config&, const FwdWs<float>&);
template bool iterate_resident_applies<double>(const lqpb_config&, const FwdWs<double>&);

#define INST(T)                                                                                                  \
  template cudaError_t launch_iterate_resident<T>(const lqpb_config&, const FwdWs<T>&, int, int, T*, int*, cudaStream_t, \
                                                  bool*);
INST(float)
INST(double)
#undef INST

}  // namespace lqpb

#ifdef LQPB_PHASE_TIMERS
// developer aid: clock64 totals of thread 0 of CTA 0 per phase of the resident iteration kernel
extern "C" void lqpb_debug_res_cycles(long long* out, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, lqpb::g_res_cycles, sizeof(long long) * 16);
  if (reset) {
    long long z[16] = {0};
    cudaMemcpyToSymbol(lqpb::g_res_cycles, z, sizeof(z));
  }
}
#endif
