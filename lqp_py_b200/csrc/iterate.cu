// K3 -- the persistent ADMM iteration kernel, and K4 -- finalisation.
//
// Restates the loop of the reference forward solver, lqp_py/solve_box_qp_admm_torch.py:235-313:
//   :237-256  adaptive-rho decision at the top of iteration i from the residuals of the previous check
//             (the refactorisation itself is done by the host between two launches of this kernel)
//   :259-268  rhs = -p~ + rho (z - u);  x = M^-1 [rhs; b~]   ->  here  x = K11 rhs + c  (factor.cu)
//   :271-282  z = clamp(x + u), r = x - z, s = rho (z - z_prev), u += r
//   :285-313  every check_solved iterations: six inf-norms (one needs Q~ x), tolerances, and the
//             GLOBAL stop test "all problems optimal"
//   :327      nus = (last KKT solve)[n:] * E   ->  here  nu = K21 rhs + K22 b~, times E
//
// Design: one persistent CTA per problem (problems are strided over the grid when B exceeds the number
// of resident CTAs).  The only O(n^2) data of an iteration is the symmetric operator K11 (and Q~ at the
// checks); both are stored as packed lower triangles (layout.cuh Pack<T>: 4 KB tiles, block-column major,
// chunk-rotated rows, halved diagonal), so an iteration streams n(n+32)/2 instead of n^2 elements -- at
// dz=500, B=128 the fp32 operator set (70 MB) even stays resident in the 126 MB L2 across iterations.
// The tile sequence of a matrix is cut into one contiguous run per warp.  Every warp owns a private ring
// of `depth` 4 KB slots and feeds it itself with 1-D bulk TMA copies (cp.async.bulk + mbarrier
// complete_tx): after consuming a slot, lane 0 immediately re-arms it with the tile `depth` positions
// ahead in the warp's own stream -- which runs on across pass and iteration boundaries because the
// operators do not change, so the memory system never idles during the O(n) vector phase and there is no
// producer/consumer handshake between warps at all.  A tile T (block row I, block column J) is applied
// symmetrically from shared memory: lane l reads row l (conflict-free thanks to the rotation) and
// accumulates  x_I[l] += T[l,:] v_J  (one register) and  x_J[:] += T[l,:] v_I[l]  (TC registers, combined
// across lanes with a butterfly of shuffles once per block column).  Partial sums go to a per-warp
// slice of shared memory with a fixed summation order: results are deterministic and independent of the
// position of a problem in the batch.  All problems advance in lock step and stop together: every
// check_solved iterations each CTA publishes its flags and a grid barrier (atomic counter) makes the
// decision global, exactly like torch.all(is_optimal) in the reference.  No host round trip per iteration.
#include "itergeom.cuh"

namespace lqpb {

// TAPE = true is the recording pass of the unrolled mode (lqpb_unroll_record_*): identical arithmetic and
// decisions, plus x~_i, z_i, u_i and the equality part nu_i of EVERY KKT solve written to the tape.
template <typename T, bool TAPE>
__global__ void __launch_bounds__(kIterMaxThreads, 1)
iterate_kernel(lqpb_config cfg, FwdWs<T> w, int i0, int skip_rho_check, T* nus_out, IterGeom geo, Tape<T> tape) {
  using P = Pack<T>;
  constexpr int TC = P::TC, TILE = P::TILE;
  using V4 = typename Vec<T>::type;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int n = w.n, m = w.m, ld = w.ld, np = geo.np;
  const int nthreads = blockDim.x, nwarps = geo.nwarps, depth = geo.depth;
  const int ntv = geo.nt, ntiles = geo.ntiles;
  T* ring = reinterpret_cast<T*>(smem_raw);                 // [nwarps][depth][TILE]
  T* xpart = ring + (size_t)nwarps * depth * TILE;          // [nwarps][np] per-warp partial sums of K v
  T* v = xpart + (size_t)nwarps * np;                       // [np] rhs of the x-update (zero padded)
  T* xs = v + np;                                           // [np] x~ of this iteration (zero padded)
  T* Ds = xs + np;                                          // [np] D
  T* tdot = Ds + np;                                        // [max(m,1)] K21 rhs
  T* red = tdot + (m > 0 ? round_up(m, 4) : 4);             // [6][16] reduction scratch
  uint64_t* full = reinterpret_cast<uint64_t*>(red + 6 * 16 + 4);   // [nwarps][depth]
  __shared__ int s_dec[4];

  const int tid = threadIdx.x;
  const int wid = tid >> 5, lane = tid & 31;
  const int nprob = (w.B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  Ctrl* ctrl = w.ctrl;

  if (tid == 0) {
    for (int s = 0; s < nwarps * depth; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  for (int e = tid; e < nwarps * np; e += nthreads) xpart[e] = T(0);
  for (int e = tid; e < np; e += nthreads) { v[e] = T(0); xs[e] = T(0); Ds[e] = T(1); }
  __syncthreads();

  const bool any_lb = ctrl->any_lb != 0, any_ub = ctrl->any_ub != 0;
  int last_wants = ctrl->last_wants, last_rout = ctrl->last_ratio_out;
  const int check = cfg.check_solved;
  const T eps_abs = (T)cfg.eps_abs, eps_rel = (T)cfg.eps_rel, zc = (T)cfg.zero_clamp;
  const T thr = (T)cfg.adaptive_rho_threshold, ar_tol = (T)cfg.adaptive_rho_tol, ar_tol_inv = (T)(1.0 / cfg.adaptive_rho_tol);

  // ---- this warp's run of tiles (the same for every matrix) and the tile it starts with
  const int run_lo = (int)((long long)wid * ntiles / nwarps);
  const int run_len = (int)((long long)(wid + 1) * ntiles / nwarps) - run_lo;
  int Jc_first = 0, I_first = 0;
  {
    int rem = run_lo;
    while (Jc_first < geo.nbc && rem >= ntv - Jc_first / P::R) { rem -= ntv - Jc_first / P::R; ++Jc_first; }
    I_first = Jc_first / P::R + rem;
  }
  T* const ring_w = ring + (size_t)wid * depth * TILE;
  uint64_t* const full_w = full + wid * depth;
  T* const xp = xpart + (size_t)wid * np;

  // ---- the warp's tile stream: position of the next tile to fetch (p_*) and ring bookkeeping.  The stream
  //      is  for i: for problem k: K11 run, then (check iterations) Q~ run;  it is fetched `depth` tiles ahead.
  int p_i = i0, p_k = 0, p_pass = 0, p_r = 0, p_slot = 0;
  int c_slot = 0;
  uint32_t c_phase = 0;
  int in_flight = 0;
  uint64_t pol_keep = 0, pol_stream = 0;
  if (lane == 0) { pol_keep = l2_policy_evict_last(); pol_stream = l2_policy_evict_first(); }
  auto issue_next = [&]() {
    if (run_len == 0 || p_i >= cfg.max_iters) return;
    if (lane == 0) {
      const int b = blockIdx.x + p_k * gridDim.x;
      const T* src = (p_pass == 0 ? w.Kp : w.Qp) + ((size_t)b * ntiles + run_lo + p_r) * TILE;
      mbar_arrive_expect_tx(&full_w[p_slot], (uint32_t)(TILE * sizeof(T)));
      tma_load_1d_hint(ring_w + (size_t)p_slot * TILE, src, (uint32_t)(TILE * sizeof(T)), &full_w[p_slot],
                       p_pass == 0 ? pol_keep : pol_stream);
    }
    ++in_flight;
    if (++p_slot == depth) p_slot = 0;
    if (++p_r == run_len) {
      p_r = 0;
      if (++p_pass == ((p_i % check) == 0 ? 2 : 1)) {
        p_pass = 0;
        if (++p_k == nprob) { p_k = 0; ++p_i; }
      }
    }
  };
  for (int d = 0; d < depth; ++d) issue_next();

  // ---- one symmetric pass over the warp's run:  xp += (this warp's share of)  S vec,  S = K11 or Q~
  auto sym_pass = [&](const T* vec) {
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
    auto load_vJ = [&]() { sa.load_vJ(vec + Jc * TC); };
    bool dirty = false;
    for (int r = 0; r < run_len; ++r) {
      mbar_wait(&full_w[c_slot], c_phase);
      const T* tp = ring_w + (size_t)c_slot * TILE;
      V4 kv[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) kv[k] = *reinterpret_cast<const V4*>(tp + P::in_tile(lane, k));
      const T vI = vec[I * kPackRows + lane];
      xp[I * kPackRows + lane] += sa.apply(kv, vI);
      __syncwarp();                       // every lane has consumed the slot: it can be re-armed
      if (++c_slot == depth) { c_slot = 0; c_phase ^= 1u; }
      --in_flight;
      issue_next();
      dirty = true;
      if (++I == ntv) {
        flush_cols();
        dirty = false;
        ++Jc;
        I = Jc / P::R;
        if (r + 1 < run_len) load_vJ();
      }
    }
    if (dirty) flush_cols();
  };

  bool have_v = false;      // v already holds the rhs of this iteration (single-problem CTAs)
  int i = i0;
  int status = 0;
  unsigned barrier_epoch = 0;

  while (true) {
    // ---------------- adaptive rho (:237-256): decided from the previous check, applied before iteration i
    if (cfg.adaptive_rho && i > 0 && i < cfg.adaptive_rho_max_iter && (i % cfg.adaptive_rho_iter) == 0 &&
        !(i == i0 && skip_rho_check)) {
      if (last_wants && last_rout) {
        for (int k = tid; k < nprob; k += nthreads) {
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
    const bool maybe_final = TAPE || is_check || is_last;

    int cta_notopt = 0, cta_wants = 0, cta_rout = 0, cta_bad = 0;
    for (int k = 0; k < nprob; ++k) {
      const int b = blockIdx.x + k * gridDim.x;
      const size_t vo = (size_t)b * ld;
      const T rho = w.rho[b];
      if (!have_v) {
        for (int e = tid; e < n; e += nthreads) v[e] = -w.pt[vo + e] + rho * (w.z[vo + e] - w.u[vo + e]);
        __syncthreads();
      }
      // ---- x~ = K11 v (+ c below): symmetric sweep over the packed tiles
      sym_pass(v);
      __syncthreads();
      // ---- K21 rhs for nu (:327), from the rhs of THIS solve (before v is overwritten)
      if (maybe_final && m > 0) {
        const T* Gt = w.Gt + (size_t)b * m * ld;
        for (int l = wid; l < m; l += nwarps) {
          T d = T(0);
          for (int e = lane; e < n; e += 32) d += Gt[(size_t)l * ld + e] * v[e];
          d = warp_sum(d);
          if (lane == 0) tdot[l] = d;
        }
        __syncthreads();
      }
      // ---- element-wise ADMM update (:271-282) and the rhs of the next iteration (:259-262)
      T mx_p = T(0), mx_d = T(0), mx_x = T(0), mx_z = T(0), mx_y = T(0);
      for (int e = tid; e < n; e += nthreads) {
        T x = T(0);
        for (int ww = 0; ww < nwarps; ++ww) {
          x += xpart[(size_t)ww * np + e];
          xpart[(size_t)ww * np + e] = T(0);
        }
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
        if (TAPE) {
          const size_t to = ((size_t)b * tape.n_iter + i) * n + e;
          tape.x[to] = x;
          tape.z[to] = zn;
          tape.u[to] = un;
        }
        if (is_check) {
          if (!(t_abs(x) < t_inf<T>())) cta_bad = 1;       // NaN / inf iterate: numerical breakdown
          const T d = w.D[vo + e];
          Ds[e] = d;
          mx_p = t_max(mx_p, t_abs(d * r));
          mx_d = t_max(mx_d, t_abs(d * sres));
          mx_x = t_max(mx_x, t_abs(d * x));
          mx_z = t_max(mx_z, t_abs(d * zn));
          mx_y = t_max(mx_y, t_abs(rho * d * un));
        }
      }
      have_v = (nprob == 1);
      __syncthreads();
      if (maybe_final && m > 0 && tid < m) {     // nu = K21 rhs + K22 b~, unscaled by E (:327)
        const T* K22 = w.Sinv + (size_t)b * m * m;
        T a = tdot[tid];
        for (int l = 0; l < m; ++l) a += K22[tid * m + l] * w.bt[(size_t)b * m + l];
        if (!TAPE || nus_out) nus_out[(size_t)b * m + tid] = a * w.E[(size_t)b * m + tid];
        if (TAPE) tape.nu[((size_t)b * tape.n_iter + i) * m + tid] = a;
      }
      if (is_check) {
        // ---- ||Q~ x~ / D||_inf (:299): the same symmetric sweep over the packed Q~ tiles
        sym_pass(xs);
        __syncthreads();
        T mx_q = T(0);
        for (int e = tid; e < n; e += nthreads) {
          T y = T(0);
          for (int ww = 0; ww < nwarps; ++ww) {
            y += xpart[(size_t)ww * np + e];
            xpart[(size_t)ww * np + e] = T(0);
          }
          mx_q = t_max(mx_q, t_abs(y / Ds[e]));
        }
        // ---- block reduction of the six maxima
        mx_p = warp_max(mx_p); mx_d = warp_max(mx_d); mx_x = warp_max(mx_x);
        mx_z = warp_max(mx_z); mx_y = warp_max(mx_y); mx_q = warp_max(mx_q);
        if (lane == 0) {
          red[0 * 16 + wid] = mx_p; red[1 * 16 + wid] = mx_d; red[2 * 16 + wid] = mx_x;
          red[3 * 16 + wid] = mx_z; red[4 * 16 + wid] = mx_y; red[5 * 16 + wid] = mx_q;
        }
        __syncthreads();
        if (tid == 0) {
          T mm[6];
          for (int a = 0; a < 6; ++a) {
            T r = red[a * 16];
            for (int ww = 1; ww < nwarps; ++ww) r = t_max(r, red[a * 16 + ww]);
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
        __syncthreads();   // red[] reusable
      }
    }
    // ---- publish this CTA's flags and make the decision global (:312 torch.all)
    if (is_check) {
      cta_bad = __syncthreads_or(cta_bad);
      if (tid == 0) {
        int* slot = ctrl->slot[(i / check) & 3];
        if (cta_notopt) atomicAdd(&slot[0], cta_notopt);
        if (cta_wants) atomicOr(&slot[1], 1);
        if (cta_rout) atomicOr(&slot[2], 1);
        if (cta_bad) atomicOr(&slot[3], 1);
        __threadfence();
        atomicAdd(&ctrl->barrier, 1u);
        const unsigned target = (barrier_epoch + 1) * gridDim.x;
        while (ld_acquire_u32(&ctrl->barrier) < target) {
        }
        __threadfence();
        s_dec[0] = *(volatile int*)&slot[0];
        s_dec[1] = *(volatile int*)&slot[1];
        s_dec[2] = *(volatile int*)&slot[2];
        s_dec[3] = *(volatile int*)&slot[3];
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
      const int notopt = s_dec[0];
      last_wants = s_dec[1];
      last_rout = s_dec[2];
      const int broken = s_dec[3];
      __syncthreads();
      if (broken) { status = 4; break; }           // LQPB_STATUS_BREAKDOWN: some iterate is NaN / inf
      if (notopt == 0) { status = 1; break; }
    }
    if (is_last) { status = 2; break; }
    ++i;
  }
  // ---- drain the tiles that were fetched ahead (a CTA must not exit with bulk copies in flight)
  while (in_flight > 0) {
    mbar_wait(&full_w[c_slot], c_phase);
    if (++c_slot == depth) { c_slot = 0; c_phase ^= 1u; }
    --in_flight;
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
cudaError_t launch_iterate(const lqpb_config& cfg, const FwdWs<T>& w, int i0, int skip_rho_check, T* nus_out,
                           int* launches, cudaStream_t st, const Tape<T>* tape) {
  // regime dispatch: operators that fit in shared memory stay there for the whole solve
  // (iterate_res.cu); everything else -- and the recording pass of the unrolled mode -- streams them (this file)
  if (!tape) {
    bool taken = false;
    cudaError_t er = launch_iterate_rows<T>(cfg, w, i0, skip_rho_check, nus_out, launches, st, &taken);
    if (er != cudaSuccess || taken) return er;
    er = launch_iterate_resident<T>(cfg, w, i0, skip_rho_check, nus_out, launches, st, &taken);
    if (er != cudaSuccess || taken) return er;
  }
  int dev = 0, max_smem = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  size_t smem = 0;
  IterGeom geo{};
  if (!make_geom(w, max_smem - 1024, &geo, &smem)) return cudaErrorInvalidConfiguration;
  const int grid = w.B < sms ? w.B : sms;   // one CTA per SM: all CTAs co-resident (needed by the grid barrier)
  e = cudaMemsetAsync(&w.ctrl->barrier, 0, sizeof(unsigned), st);
  if (e != cudaSuccess) return e;
  void* kern = tape ? (void*)iterate_kernel<T, true> : (void*)iterate_kernel<T, false>;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  lqpb_config c = cfg;
  FwdWs<T> ww = w;
  Tape<T> tp = tape ? *tape : Tape<T>{};
  void* args[] = {&c, &ww, &i0, &skip_rho_check, &nus_out, &geo, &tp};
  e = cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(geo.nwarps * 32), args, smem, st);
  if (launches) ++*launches;
  return e;
}

// Per-problem record of the LAST stop check (reference :286-311 evaluates these and throws them away; a solve that
// runs out of iterations is silent there): residual norms, the relative scales of their tolerances, and whether the
// problem satisfied its own stop test.  out: (B, 4) = [primal, dual, tol_primal, tol_dual], converged: (B) int.
template <typename T>
__global__ void status_kernel(lqpb_config cfg, FwdWs<T> w, T* out, int* converged) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= w.B) return;
  const T primal = w.chk[4 * b + 0], dual = w.chk[4 * b + 1];
  const T tol_p = (T)cfg.eps_abs + (T)cfg.eps_rel * w.chk[4 * b + 2];
  const T tol_d = (T)cfg.eps_abs + (T)cfg.eps_rel * w.chk[4 * b + 3];
  out[4 * b + 0] = primal;
  out[4 * b + 1] = dual;
  out[4 * b + 2] = tol_p;
  out[4 * b + 3] = tol_d;
  converged[b] = (primal < tol_p && dual < tol_d) ? 1 : 0;
}

template <typename T>
cudaError_t launch_status(const lqpb_config& cfg, const FwdWs<T>& w, T* out, int* converged, cudaStream_t st) {
  status_kernel<T><<<(w.B + 127) / 128, 128, 0, st>>>(cfg, w, out, converged);
  return cudaGetLastError();
}

template <typename T>
cudaError_t launch_finalize(const FwdWs<T>& w, T* x, T* z, T* u, T* lams, T* rho_out, cudaStream_t st) {
  dim3 grid((w.n + 127) / 128, w.B);
  finalize_kernel<T><<<grid, 128, 0, st>>>(w, x, z, u, lams, rho_out);
  return cudaGetLastError();
}

#define INST(T)                                                                                            \
  template cudaError_t launch_iterate<T>(const lqpb_config&, const FwdWs<T>&, int, int, T*, int*, cudaStream_t, \
                                         const Tape<T>*);                                                  \
  template cudaError_t launch_finalize<T>(const FwdWs<T>&, T*, T*, T*, T*, T*, cudaStream_t);                \
  template cudaError_t launch_status<T>(const lqpb_config&, const FwdWs<T>&, T*, int*, cudaStream_t);
INST(float)
INST(double)
#undef INST

}  // namespace lqpb
