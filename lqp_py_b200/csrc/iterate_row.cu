// K3, small-problem regime -- the ADMM iteration kernel for problems whose operators fit in shared memory as FULL
// matrices (dz <= ~150 in fp32, ~105 in fp64: the dz = 10, 50, 100 configurations of the reference's Experiment 1).
//
// Same loop as iterate.cu (reference lqp_py/solve_box_qp_admm_torch.py:235-313, :327) with the same global semantics
// (all problems advance in lock step, stop together, adaptive-rho trigger from the previous check).  SURVEY App. C:
// at these sizes an iteration is bound by instruction latency and synchronisation, not by bytes, so the kernel is
// organised to have as few dependent steps per iteration as possible:
//   * the packed K11 (and Q~) of every problem a CTA owns are expanded ONCE into dense row-major matrices in shared
//     memory; an x-update is then a plain row-times-vector product -- `lpr` lanes share a row, each reads 16-byte
//     chunks of the row and of the broadcast rhs, a log2(lpr)-step shuffle completes the dot product -- with no
//     partial sums across warps and no column accumulators;
//   * the lane that owns row r finishes the iteration for that coordinate in registers (clamp, dual update, next
//     rhs into the OTHER rhs buffer, residual maxima): ONE group barrier per iteration;
//   * a group of `gw` warps owns whole problems, groups run independently (dz = 10: one problem per warp, the
//     barrier is a __syncwarp); the CTA meets only at the stop checks;
//   * a stop check publishes the CTA's three flags AND its barrier arrival with ONE 64-bit reduction
//     (arrivals | wants << 16 | ratio_out << 28 | not_optimal << 40 | breakdown << 52); the word that completes the count carries the
//     global decision, so the grid barrier costs one L2 round trip.  When the batch fits <= 8 CTAs the grid is one
//     thread-block cluster and the wait is the hardware cluster barrier (0.26 us measured against 1.3 us);
//   * nus (:327) is formed once, after the loop, from the rhs of the last solve (still intact in its buffer).
//
// FUSED = true is the whole forward solve of such a problem in ONE launch (lqpb_forward_* takes it when the recording
// pass of the unrolled mode is not involved): the group first builds everything else in shared memory too -- the
// scaling of :161-197 (column norms, zero guard, the q10 / q90 quantiles of D by a bitonic sort, blend, Q~ = D Q D,
// p~, A~, E, b~, lb~, ub~), rho of :200-203, the KKT matrix of :206-212 and its inverse by an in-place symmetric
// Gauss-Jordan sweep (H pivots first, equality rows last; replaces lu_factor, :215) -- then runs the loop above and
// finishes with :315-327 (un-scaling, split duals, nus) straight into the caller's tensors.  The adaptive-rho update of
// :237-256 happens ON THE DEVICE here: the flagged problems rebuild and re-sweep their KKT matrix in place and the
// loop carries on -- no status 3, no host round trip -- so the host never has to look at the solve before it ends.
#include <cstring>
#include "itergeom.cuh"

namespace lqpb {

constexpr int kRowWarps = 16;
constexpr int kRowThreads = kRowWarps * 32;
constexpr int kRowVecs = 10;         // v0, v1, z, u, p~, lb~, ub~, c, D, x~ per resident problem
constexpr int kPS = 8;               // per-problem scalars in shared memory: rho, ||p||_inf, and the record of the last
                                     // stop check [primal, dual, tol_primal_rel, tol_dual_rel, ratio, wants]

struct RowGeom {
  int G, ppc, gw, ngroups, cluster;
  int lpr;        // lanes per row (power of two)
  int ldk;        // row stride of the dense matrices (elements); chosen so that the lanes of a quarter warp hit
                  // distinct banks
  int nch;        // 16-byte chunks per row that hold data
  int ldv;        // padded vector length (covers nch chunks)
  int fused;      // 1: FUSED kernel (K block is the (n + m) x (n + m) KKT matrix, A~ rows resident, sort / sweep scratch)
  int P2;         // FUSED: power of two >= n (bitonic sort of D)
  size_t prob_elems, group_elems;
};

// raw problem data and outputs of the FUSED kernel (the caller's tensors, reference layouts)
template <typename T>
struct FusedIO {
  const T *Q, *p, *A, *b, *lb, *ub;
  T *x, *z, *u, *lams, *rho_out;
};

__device__ __forceinline__ void row_cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void row_cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void red_release_add_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

template <typename T, bool FUSED>
__global__ void __launch_bounds__(kRowThreads, 1)
iterate_row_kernel(lqpb_config cfg, FwdWs<T> w, int i0, int skip_rho_check, T* nus_out, RowGeom geo, FusedIO<T> io) {
  using P = Pack<T>;
  constexpr int VN = P::VN;
  using V4 = typename Vec<T>::type;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ unsigned long long s_word;
  __shared__ unsigned long long s_slots[2][8];      // cluster mode: the packed check words of the CTAs of the cluster
  __shared__ int s_flags[2][4];                     // per check parity: not_optimal, wants, ratio_out, breakdown
  const int n = w.n, m = w.m, ld = w.ld;
  const int gw = geo.gw, ngroups = geo.ngroups, lpr = geo.lpr, ldk = geo.ldk, nch = geo.nch, ldv = geo.ldv;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int grp = wid / gw, wg = wid % gw;
  const int gtid = wg * 32 + lane, gthreads = gw * 32;
  const int sub = gtid & (lpr - 1), rslot = gtid / lpr, rpp = gthreads / lpr;   // my chunk phase, my row slot, rows per pass
  const int nprob = (w.B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  Ctrl* ctrl = w.ctrl;

  T* base = reinterpret_cast<T*>(smem_raw);
  const int N = n + m;                                      // FUSED: order of the KKT matrix held in the K block
  const int krows = FUSED ? N : n;
  auto prob_K = [&](int q) { return base + (size_t)q * geo.prob_elems; };
  auto prob_Q = [&](int q) { return prob_K(q) + (size_t)krows * ldk; };
  auto prob_At = [&](int q) { return prob_Q(q) + (size_t)n * ldk; };            // FUSED: A~ rows [m][ldk]
  auto prob_vec = [&](int q, int k) { return prob_Q(q) + (size_t)(n + (FUSED ? m : 0)) * ldk + (size_t)k * ldv; };
  T* gscr = base + (size_t)geo.ppc * geo.prob_elems + (size_t)grp * geo.group_elems;
  T* red = gscr;                                            // [6][16]
  double* dred = reinterpret_cast<double*>(gscr + 96);      // [16]   (FUSED)
  T* sortbuf = gscr + 96 + 16 * (int)(sizeof(double) / sizeof(T));   // [P2]  (FUSED)
  T* csv = sortbuf + geo.P2;                                // [ldk]  (FUSED) pivot column of a sweep step
  T* pscal = base + (size_t)geo.ppc * geo.prob_elems + (size_t)ngroups * geo.group_elems;   // [ppc][kPS]

  // ---- one-time load: zero the problem blocks, then either expand the packed lower triangles a factorisation left in
  //      the workspace into dense symmetric matrices, or (FUSED) leave the set-up to the groups below
  for (size_t t = tid; t < (size_t)nprob * geo.prob_elems; t += kRowThreads) base[t] = T(0);
  if (tid < 8) s_flags[tid >> 2][tid & 3] = 0;
  __syncthreads();
  // cluster mode: a CTA's shared memory may only be written by its peers (DSMEM stores at the stop checks) once it has
  // started executing -- one cluster barrier up front establishes that (compute-sanitizer racecheck: "block that might not
  // have entered yet" without it)
  if (geo.cluster && gridDim.x > 1) {
    row_cluster_arrive();
    row_cluster_wait();
  }
  if constexpr (!FUSED) {
    const int ntv = P::nt(n), ntiles = P::ntiles(n);
    const size_t per = (size_t)ntiles * P::TILE;
    for (size_t t = tid; t < (size_t)nprob * per; t += kRowThreads) {
      const int q = (int)(t / per);
      const size_t o = t % per;
      const int tile = (int)(o / P::TILE), in = (int)(o % P::TILE);
      // inverse of Pack::in_tile (chunk-major): in = (k * 32 + l) * VN + e
      const int e = in % VN, l = (in / VN) % kPackRows, k = in / (VN * kPackRows);
      int Jc = 0, rem = tile;
      while (rem >= ntv - Jc / P::R) { rem -= ntv - Jc / P::R; ++Jc; }
      const int I = Jc / P::R + rem;
      const int i = I * kPackRows + l, j = Jc * P::TC + k * VN + e;
      if (i < n && j <= i) {
        const int b = blockIdx.x + q * gridDim.x;
        const size_t go = (size_t)b * per + o;
        T kv = w.Kp[go], qv = w.Qp[go];
        if (i == j) { kv += kv; qv += qv; }          // the packed layout stores the diagonal halved
        prob_K(q)[(size_t)i * ldk + j] = kv;
        prob_K(q)[(size_t)j * ldk + i] = kv;
        prob_Q(q)[(size_t)i * ldk + j] = qv;
        prob_Q(q)[(size_t)j * ldk + i] = qv;
      }
    }
    for (int t = tid; t < nprob * n; t += kRowThreads) {
      const int q = t / n, e = t % n;
      const int b = blockIdx.x + q * gridDim.x;
      const size_t vo = (size_t)b * ld + e;
      const T rho = w.rho[b], z = w.z[vo], u = w.u[vo], pt = w.pt[vo];
      prob_vec(q, 0)[e] = -pt + rho * (z - u);       // rhs of the first iteration (:259-262)
      prob_vec(q, 2)[e] = z;
      prob_vec(q, 3)[e] = u;
      prob_vec(q, 4)[e] = pt;
      prob_vec(q, 5)[e] = w.lbt[vo];
      prob_vec(q, 6)[e] = w.ubt[vo];
      prob_vec(q, 7)[e] = w.c[vo];
      prob_vec(q, 8)[e] = w.D[vo];
    }
    for (int q = tid; q < nprob; q += kRowThreads) {
      const int b = blockIdx.x + q * gridDim.x;
      pscal[kPS * q] = w.rho[b];
      pscal[kPS * q + 1] = w.pnorm[b];
      for (int k = 0; k < 4; ++k) pscal[kPS * q + 2 + k] = w.chk[4 * b + k];
      pscal[kPS * q + 6] = w.ratio[b];
      pscal[kPS * q + 7] = (T)w.wants[b];
    }
    __syncthreads();
  }

  const bool any_lb = ctrl->any_lb != 0, any_ub = ctrl->any_ub != 0;
  int last_wants = ctrl->last_wants, last_rout = ctrl->last_ratio_out;
  const int check = cfg.check_solved;
  const T eps_abs = (T)cfg.eps_abs, eps_rel = (T)cfg.eps_rel, zc = (T)cfg.zero_clamp;
  const T thr = (T)cfg.adaptive_rho_threshold, ar_tol = (T)cfg.adaptive_rho_tol, ar_tol_inv = (T)(1.0 / cfg.adaptive_rho_tol);
  auto group_sync = [&]() {
    if (gw == 1) __syncwarp();
    else bar_sync(1 + grp, gthreads);
  };
  // dot product of row r of a dense matrix with a vector: the lpr lanes of the row take the chunks sub, sub + lpr, ...
  auto row_dot = [&](const T* M, const T* vec, int r) -> T {
    T a0 = T(0), a1 = T(0);
    if (r < n) {
      const T* row = M + (size_t)r * ldk;
      for (int c = sub; c < nch; c += lpr) {
        const V4 mv = *reinterpret_cast<const V4*>(row + c * VN);
        const V4 vv = *reinterpret_cast<const V4*>(vec + c * VN);
        const T* mp = reinterpret_cast<const T*>(&mv);
        const T* vp = reinterpret_cast<const T*>(&vv);
#pragma unroll
        for (int e = 0; e < VN; e += 2) {
          a0 += mp[e] * vp[e];
          a1 += mp[e + 1] * vp[e + 1];
        }
      }
    }
    T a = a0 + a1;
    for (int o = lpr >> 1; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    return a;
  };

  // ---- FUSED: group-level reductions, the KKT build + sweep, and the set-up of every problem of this group
  auto gsum_d = [&](double v) -> double {
    v = warp_sum(v);
    if (gw == 1) return v;
    if (lane == 0) dred[wg] = v;
    group_sync();
    double r = 0.0;
    for (int ww = 0; ww < gw; ++ww) r += dred[ww];
    group_sync();
    return r;
  };
  auto gmax = [&](T v) -> T {
    v = warp_max(v);
    if (gw == 1) return v;
    if (lane == 0) red[wg] = v;
    group_sync();
    T r = red[0];
    for (int ww = 1; ww < gw; ++ww) r = t_max(r, red[ww]);
    group_sync();
    return r;
  };
  // K block <- inverse of the KKT matrix [[Q~ + rho I, A~^T], [A~, 0]] (:206-215): build, in-place symmetric
  // Gauss-Jordan sweep without pivoting (H pivots are positive, the equality rows are swept last, when their pivot
  // block has become -(A~ H^-1 A~^T)), negate; then c = K21^T b~
  auto build_and_sweep = [&](int q, int b, T rho) {
    T* M = prob_K(q);
    const T* Qm = prob_Q(q);
    const T* At = prob_At(q);
    for (int t = gtid; t < N * ldk; t += gthreads) {
      const int i = t / ldk, j = t % ldk;
      T v = T(0);
      if (j < N) {
        if (i < n && j < n) v = Qm[(size_t)i * ldk + j] + (i == j ? rho : T(0));
        else if (i >= n && j < n) v = At[(size_t)(i - n) * ldk + j];
        else if (i < n && j >= n) v = At[(size_t)(j - n) * ldk + i];
      }
      M[t] = v;
    }
    group_sync();
    const int nchN = (N + VN - 1) / VN;
    for (int s = 0; s < N; ++s) {
      for (int j = gtid; j < nchN * VN; j += gthreads) csv[j] = j < N ? M[(size_t)s * ldk + j] : T(0);   // row s == column s
      group_sync();
      const T piv = T(1) / csv[s];
      const int cs = s / VN, es = s % VN;
      // 16-byte chunks, the lane mapping of row_dot: conflict-free LDS.128 / STS.128 under the padded row stride
      for (int r0 = 0; r0 < N; r0 += rpp) {
        const int i = r0 + rslot;
        if (i < N) {
          T* row = M + (size_t)i * ldk;
          const T ci = csv[i] * piv;
          for (int c = sub; c < nchN; c += lpr) {
            const V4 cv = *reinterpret_cast<const V4*>(csv + c * VN);
            V4 rv = *reinterpret_cast<const V4*>(row + c * VN);
            const T* cp = reinterpret_cast<const T*>(&cv);
            T* rp = reinterpret_cast<T*>(&rv);
            if (i == s) {
#pragma unroll
              for (int e = 0; e < VN; ++e) rp[e] = (c == cs && e == es) ? -piv : cp[e] * piv;
            } else {
#pragma unroll
              for (int e = 0; e < VN; ++e) rp[e] = (c == cs && e == es) ? ci : rp[e] - ci * cp[e];
            }
            *reinterpret_cast<V4*>(row + c * VN) = rv;
          }
        }
      }
      group_sync();
    }
    for (int t = gtid; t < N * ldk; t += gthreads) M[t] = -M[t];
    group_sync();
    T* cvec = prob_vec(q, 7);
    for (int i = gtid; i < n; i += gthreads) {
      T a = T(0);
      for (int l = 0; l < m; ++l) a += M[(size_t)(n + l) * ldk + i] * w.bt[(size_t)b * m + l];
      cvec[i] = a;
    }
    group_sync();
  };
  if constexpr (FUSED) {
    const bool boxed = ctrl->any_lb != 0 || ctrl->any_ub != 0;       // bound_flags_kernel ran before this launch
    for (int q = grp; q < nprob; q += ngroups) {
      const int b = blockIdx.x + q * gridDim.x;
      T* Qm = prob_Q(q);
      T* At = prob_At(q);
      T* Dv = prob_vec(q, 8);
      const T* Qb = io.Q + (size_t)b * n * n;
      for (int t = gtid; t < n * n; t += gthreads) Qm[(size_t)(t / n) * ldk + (t % n)] = Qb[t];
      group_sync();
      if (cfg.scale) {
        // column inf-norms (:163), zero guard (:164-168), D = sqrt(1 / norm) (:170)
        for (int j = gtid; j < n; j += gthreads) {
          T mx = T(0);
          for (int i = 0; i < n; ++i) mx = t_max(mx, t_abs(Qm[(size_t)i * ldk + j]));
          Dv[j] = mx;
        }
        group_sync();
        double part = 0.0;
        for (int j = gtid; j < n; j += gthreads) part += (double)Dv[j];
        const double tot = gsum_d(part);
        const T floor_v = t_max((T)(tot / n), T(1e-6));
        for (int j = gtid; j < n; j += gthreads) {
          T qn = Dv[j];
          if (qn <= T(0)) qn = t_max(qn, floor_v);
          Dv[j] = t_sqrt(T(1) / qn);
        }
        group_sync();
        // beta = 1 - q10(D) / q90(D) (:171-174): bitonic sort + torch.quantile's linear interpolation
        T beta = (T)cfg.beta;
        if (cfg.beta_auto) {
          const int P2 = geo.P2;
          for (int j = gtid; j < P2; j += gthreads) sortbuf[j] = j < n ? Dv[j] : t_inf<T>();
          group_sync();
          for (int k = 2; k <= P2; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
              for (int ii = gtid; ii < P2; ii += gthreads) {
                const int ixj = ii ^ j;
                if (ixj > ii) {
                  const T a = sortbuf[ii], c = sortbuf[ixj];
                  const bool asc = (ii & k) == 0;
                  if ((a > c) == asc) { sortbuf[ii] = c; sortbuf[ixj] = a; }
                }
              }
              group_sync();
            }
          }
          T qv[2];
          const T qs[2] = {T(0.10), T(0.90)};
          for (int k = 0; k < 2; ++k) {
            const T rank = qs[k] * T(n - 1);
            const T lo = floor(rank), hi = ceil(rank);
            const T a = sortbuf[(int)lo], c = sortbuf[(int)hi], wgt = rank - lo;
            qv[k] = wgt < T(0.5) ? a + wgt * (c - a) : c - (c - a) * (T(1) - wgt);      // torch's lerp
          }
          beta = T(1) - qv[0] / qv[1];
          group_sync();
        }
        // D <- (1 - beta) D + beta mean(D) (:175)
        part = 0.0;
        for (int j = gtid; j < n; j += gthreads) part += (double)Dv[j];
        const T mean = (T)(gsum_d(part) / n);
        for (int j = gtid; j < n; j += gthreads) Dv[j] = (T(1) - beta) * Dv[j] + beta * mean;
      } else {
        for (int j = gtid; j < n; j += gthreads) Dv[j] = T(1);
      }
      group_sync();
      // vectors (:127, :177, :192-194, :221-223 or the warm start)
      T pmax = T(0);
      for (int j = gtid; j < n; j += gthreads) {
        const T d = Dv[j];
        T pv = io.p[(size_t)b * n + j], l = io.lb[(size_t)b * n + j], uu = io.ub[(size_t)b * n + j];
        pmax = t_max(pmax, t_abs(pv));
        if (cfg.scale) { pv = d * pv; l = l / d; uu = uu / d; }
        prob_vec(q, 4)[j] = pv;
        prob_vec(q, 5)[j] = l;
        prob_vec(q, 6)[j] = uu;
        prob_vec(q, 2)[j] = w.z0 ? w.z0[(size_t)b * n + j] / d : T(0);
        prob_vec(q, 3)[j] = w.u0 ? w.u0[(size_t)b * n + j] * d : T(0);
      }
      pmax = gmax(pmax);
      // Q~ = (D_i Q_ij) D_j (:176) in place, ||Q~||_F^2 in double (:201), rho (:156-158, :200-203)
      double fro = 0.0;
      for (int t = gtid; t < n * n; t += gthreads) {
        const int i = t / n, j = t % n;
        T v = Qm[(size_t)i * ldk + j];
        if (cfg.scale) v = (Dv[i] * v) * Dv[j];
        Qm[(size_t)i * ldk + j] = v;
        fro += (double)v * (double)v;
      }
      fro = gsum_d(fro);
      T rc = (T)sqrt(fro) / (T)sqrt((double)n);
      rc = t_min(t_max(rc, (T)cfg.rho_min), (T)cfg.rho_max);
      T rho = cfg.rho_auto ? rc : (w.rho_in ? w.rho_in[b] : (T)cfg.rho);
      if (!boxed) rho = T(0);
      // equality rows: A~ = E (A D), b~ = E b (:179-190)
      for (int l = 0; l < m; ++l) {
        const T* Al = io.A + ((size_t)b * m + l) * n;
        T mx = T(0);
        for (int j = gtid; j < n; j += gthreads) {
          const T v = Al[j] * Dv[j];
          At[(size_t)l * ldk + j] = cfg.scale ? v : Al[j];
          mx = t_max(mx, t_abs(v));
        }
        mx = gmax(mx);
        if (gtid == 0) sortbuf[l % geo.P2] = mx;      // m <= P2 is guaranteed by the plan
        group_sync();
      }
      if (m > 0) {
        if (gtid == 0) {
          double sden = 0.0;
          for (int l = 0; l < m; ++l) sden += (double)sortbuf[l];
          const T fl = t_max((T)(sden / m), T(1e-6));
          for (int l = 0; l < m; ++l) {
            T r = sortbuf[l];
            if (r <= T(0)) r = t_max(r, fl);
            const T e = cfg.scale ? T(1) / r : T(1);
            sortbuf[l] = e;
            w.E[(size_t)b * m + l] = e;
            w.bt[(size_t)b * m + l] = cfg.scale ? e * io.b[(size_t)b * m + l] : io.b[(size_t)b * m + l];
          }
        }
        group_sync();
        if (cfg.scale)
          for (int t = gtid; t < m * n; t += gthreads) At[(size_t)(t / n) * ldk + (t % n)] *= sortbuf[t / n];
        group_sync();
      }
      if (gtid == 0) {
        pscal[kPS * q] = rho;
        pscal[kPS * q + 1] = pmax;
        pscal[kPS * q + 2] = pscal[kPS * q + 3] = pscal[kPS * q + 4] = pscal[kPS * q + 5] = T(0);
        pscal[kPS * q + 6] = T(1);
        pscal[kPS * q + 7] = T(0);
      }
      build_and_sweep(q, b, rho);
      for (int j = gtid; j < n; j += gthreads)
        prob_vec(q, 0)[j] = -prob_vec(q, 4)[j] + rho * (prob_vec(q, 2)[j] - prob_vec(q, 3)[j]);   // :259-262
      group_sync();
    }
    __syncthreads();
  }

  int i = i0, cur = 0;       // cur: which rhs buffer holds the rhs of iteration i
  int status = 0;
  int last_check = -1;       // iteration of the most recent stop check of THIS launch

  while (true) {
    // ---------------- adaptive rho (:237-256): decided from the previous check, applied before iteration i
    if (cfg.adaptive_rho && i > 0 && i < cfg.adaptive_rho_max_iter && (i % cfg.adaptive_rho_iter) == 0 &&
        !(i == i0 && skip_rho_check)) {
      if (last_wants && last_rout) {
        if constexpr (FUSED) {
          // on the device: every flagged problem of this group updates rho, rebuilds and re-sweeps its KKT matrix in
          // place and re-forms the rhs of THIS iteration with the new rho (u is not rescaled, like the reference)
          for (int q = grp; q < nprob; q += ngroups) {
            const int b = blockIdx.x + q * gridDim.x;
            if (pscal[kPS * q + 7] != T(0)) {
              T r = pscal[kPS * q] * pscal[kPS * q + 6];
              r = t_min(t_max(r, (T)cfg.rho_min), (T)cfg.rho_max);
              group_sync();
              if (gtid == 0) pscal[kPS * q] = r;
              build_and_sweep(q, b, r);
              T* v = prob_vec(q, cur);
              for (int j = gtid; j < n; j += gthreads) v[j] = -prob_vec(q, 4)[j] + r * (prob_vec(q, 2)[j] - prob_vec(q, 3)[j]);
              group_sync();
            }
          }
          if (blockIdx.x == 0 && tid == 0) ctrl->pad0 += 1;        // refactorisations done on the device
        } else {
          for (int k = tid; k < nprob; k += kRowThreads) {
            const int b = blockIdx.x + k * gridDim.x;
            if (pscal[kPS * k + 7] != T(0)) {
              T r = pscal[kPS * k] * pscal[kPS * k + 6];
              r = t_min(t_max(r, (T)cfg.rho_min), (T)cfg.rho_max);
              w.rho[b] = r;
            }
          }
          status = 3;
          break;
        }
      }
    }
    const bool is_check = (i % check) == 0;
    const bool is_last = i == cfg.max_iters - 1;

    for (int q = grp; q < nprob; q += ngroups) {
      const int b = blockIdx.x + q * gridDim.x;
      (void)b;
      const T* v = prob_vec(q, cur);
      T* vn = prob_vec(q, cur ^ 1);
      T* zs = prob_vec(q, 2);
      T* us = prob_vec(q, 3);
      const T* pts = prob_vec(q, 4);
      const T* lbs = prob_vec(q, 5);
      const T* ubs = prob_vec(q, 6);
      const T* cs = prob_vec(q, 7);
      const T* Ds = prob_vec(q, 8);
      T* xs = prob_vec(q, 9);
      const T rho = pscal[kPS * q];
      // ---- x~ = K11 v + c, then the element-wise ADMM update (:271-282) by the lane that owns the row
      T mx_p = T(0), mx_d = T(0), mx_x = T(0), mx_z = T(0), mx_y = T(0);
      for (int r0 = 0; r0 < n; r0 += rpp) {
        const int r = r0 + rslot;
        const T dot = row_dot(prob_K(q), v, r);
        if (sub == 0 && r < n) {
          const T x = dot + cs[r];
          const T z_prev = zs[r], u_prev = us[r];
          T zn = x + u_prev;
          if (any_lb) zn = t_max(zn, lbs[r]);
          if (any_ub) zn = t_min(zn, ubs[r]);
          const T res = x - zn;
          const T sres = rho * (zn - z_prev);
          const T un = u_prev + res;
          zs[r] = zn;
          us[r] = un;
          vn[r] = -pts[r] + rho * (zn - un);        // rhs of the next iteration (:259-262)
          xs[r] = x;
          if (is_check) {
            if (!(t_abs(x) < t_inf<T>())) s_flags[(i / check) & 1][3] = 1;  // NaN / inf iterate: breakdown (benign race)
            const T d = Ds[r];
            mx_p = t_max(mx_p, t_abs(d * res));
            mx_d = t_max(mx_d, t_abs(d * sres));
            mx_x = t_max(mx_x, t_abs(d * x));
            mx_z = t_max(mx_z, t_abs(d * zn));
            mx_y = t_max(mx_y, t_abs(rho * d * un));
          }
        }
      }
      group_sync();
      if (is_check) {
        // ---- ||Q~ x~ / D||_inf (:299)
        T mx_q = T(0);
        for (int r0 = 0; r0 < n; r0 += rpp) {
          const int r = r0 + rslot;
          const T dot = row_dot(prob_Q(q), xs, r);
          if (sub == 0 && r < n) mx_q = t_max(mx_q, t_abs(dot / Ds[r]));
        }
        mx_p = warp_max(mx_p); mx_d = warp_max(mx_d); mx_x = warp_max(mx_x);
        mx_z = warp_max(mx_z); mx_y = warp_max(mx_y); mx_q = warp_max(mx_q);
        if (gw > 1) {
          if (lane == 0) {
            red[0 * 16 + wg] = mx_p; red[1 * 16 + wg] = mx_d; red[2 * 16 + wg] = mx_x;
            red[3 * 16 + wg] = mx_z; red[4 * 16 + wg] = mx_y; red[5 * 16 + wg] = mx_q;
          }
          group_sync();
        }
        if (gtid == 0) {
          T mm[6] = {mx_p, mx_d, mx_x, mx_z, mx_y, mx_q};
          if (gw > 1) {
            for (int a = 0; a < 6; ++a) {
              T r = red[a * 16];
              for (int ww = 1; ww < gw; ++ww) r = t_max(r, red[a * 16 + ww]);
              mm[a] = r;
            }
          }
          const T primal = mm[0], dual = mm[1];
          const T tol_p_rel = t_max(t_max(mm[2], mm[3]), zc);                      // :301
          const T tol_p = eps_abs + eps_rel * tol_p_rel;                           // :302
          const T tol_d_rel = t_max(t_max(t_max(mm[4], mm[5]), pscal[kPS * q + 1]), zc);   // :303
          const T tol_d = eps_abs + eps_rel * tol_d_rel;                           // :304
          const bool optimal = (primal < tol_p) && (dual < tol_d);                // :307-309
          const bool wants = (primal > t_max(tol_p, thr)) || (dual > t_max(tol_d, thr));   // :310-311
          const T num = t_max(primal / tol_p_rel, zc), den = t_max(dual / tol_d_rel, zc);  // :239-242
          const T ratio = t_sqrt(num / den);                                       // :243
          // the record of this check stays in shared memory (the adaptive-rho update and the exit read it there)
          pscal[kPS * q + 2] = primal; pscal[kPS * q + 3] = dual;
          pscal[kPS * q + 4] = tol_p_rel; pscal[kPS * q + 5] = tol_d_rel;
          pscal[kPS * q + 6] = ratio;
          pscal[kPS * q + 7] = wants ? T(1) : T(0);
          int* fl = s_flags[(i / check) & 1];
          if (!optimal) atomicOr(&fl[0], 1);
          if (wants) atomicOr(&fl[1], 1);
          if (ratio > ar_tol || ratio < ar_tol_inv) atomicOr(&fl[2], 1);          // :244-245
          if (cfg.verbose) {
            const int ci = i / check;
            if (ci < LQPB_LOG_CAP) {
              atomic_max_nonneg(&ctrl->log_primal[ci], (double)primal);
              atomic_max_nonneg(&ctrl->log_dual[ci], (double)dual);
              ctrl->log_iter[ci] = i;
            }
          }
        }
        if (gw > 1) group_sync();   // red[] reusable
      }
    }
    // ---- the global decision (:312 torch.all): one packed word per CTA and check carries its arrival and its flags.
    //      The CTA flags are double-buffered by check parity (tid 0 clears a buffer after the barrier of its check; the
    //      next writers of that buffer are two checks away, behind the next barrier), so a check costs ONE __syncthreads
    //      plus the barrier itself.
    if (is_check) {
      const int par = (i / check) & 1;
      __syncthreads();
      const int* fl = s_flags[par];
      const unsigned long long mine = 1ull | (fl[1] ? (1ull << 16) : 0ull) | (fl[2] ? (1ull << 28) : 0ull) |
                                      (fl[0] ? (1ull << 40) : 0ull) | (fl[3] ? (1ull << 52) : 0ull);
      unsigned long long v = mine;
      if (geo.cluster) {
        // the grid is ONE cluster: every CTA drops its word into slot [parity][its rank] of every CTA's shared memory
        // (DSMEM stores), the hardware cluster barrier orders them, and every thread adds up its CTA's copy -- no
        // global memory traffic at all at a check (a single-CTA grid needs nothing)
        if (gridDim.x > 1) {
          if (tid < (int)gridDim.x) {
            const uint32_t laddr = smem_u32(&s_slots[par][blockIdx.x]);
            uint32_t raddr;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(laddr), "r"(tid));
            asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(raddr), "l"(mine) : "memory");
          }
          row_cluster_arrive();
          row_cluster_wait();
          v = 0ull;
          for (int r = 0; r < (int)gridDim.x; ++r) v += s_slots[par][r];
        } else {
          __syncthreads();        // every thread has read the flags before tid 0 clears them
        }
        if (tid == 0) s_flags[par][0] = s_flags[par][1] = s_flags[par][2] = s_flags[par][3] = 0;
      } else {
        unsigned long long* word = reinterpret_cast<unsigned long long*>(&ctrl->slot[(i / check) & 3][0]);
        if (tid == 0) {
          red_release_add_u64(word, mine);
          do {
            v = ld_acquire_u64(word);
          } while ((unsigned)(v & 0xffffull) < gridDim.x);
          s_word = v;
          if (blockIdx.x == 0) {
            // slot of the check after next: every CTA has read it (it passed the previous barrier); the store is
            // ordered before this CTA's next arrival, which every other CTA acquires before it can reach that check
            unsigned long long* nxt = reinterpret_cast<unsigned long long*>(&ctrl->slot[((i / check) + 2) & 3][0]);
            *nxt = 0ull;
          }
        }
        __syncthreads();
        v = s_word;
        if (tid == 0) s_flags[par][0] = s_flags[par][1] = s_flags[par][2] = s_flags[par][3] = 0;
      }
      last_wants = ((v >> 16) & 0xfffull) != 0;
      last_rout = ((v >> 28) & 0xfffull) != 0;
      const bool all_optimal = ((v >> 40) & 0xfffull) == 0;
      const bool broken = (v >> 52) != 0;
      last_check = i;
      if (broken) { status = 4; break; }           // LQPB_STATUS_BREAKDOWN: some iterate is NaN / inf
      if (all_optimal) { status = 1; break; }
    }
    if (is_last) { status = 2; break; }
    ++i;
    cur ^= 1;
  }
  // ---- epilogue: nus of the LAST solve (:327) from its rhs, which is still intact in buffer `cur` (status 1 / 2 / 4:
  //      the loop left before flipping; status 3 left before iteration i ran: nothing to report yet), and the state
  __syncthreads();
  if (status != 3 && m > 0) {      // (after a breakdown the values are NaN like everything else)
    for (int q = grp; q < nprob; q += ngroups) {
      const int b = blockIdx.x + q * gridDim.x;
      const T* v = prob_vec(q, cur);
      for (int l = wg; l < m; l += gw) {
        T d = T(0);
        if constexpr (FUSED) {
          const T* Kl = prob_K(q) + (size_t)(n + l) * ldk;        // K21 row l | K22 row l
          for (int e = lane; e < n; e += 32) d += Kl[e] * v[e];
          d = warp_sum(d);
          if (lane == 0) {
            T a = d;
            for (int l2 = 0; l2 < m; ++l2) a += Kl[n + l2] * w.bt[(size_t)b * m + l2];
            nus_out[(size_t)b * m + l] = a * w.E[(size_t)b * m + l];
          }
        } else {
          const T* Gt = w.Gt + (size_t)b * m * ld;
          const T* K22 = w.Sinv + (size_t)b * m * m;
          for (int e = lane; e < n; e += 32) d += Gt[(size_t)l * ld + e] * v[e];
          d = warp_sum(d);
          if (lane == 0) {
            T a = d;
            for (int l2 = 0; l2 < m; ++l2) a += K22[l * m + l2] * w.bt[(size_t)b * m + l2];
            nus_out[(size_t)b * m + l] = a * w.E[(size_t)b * m + l];
          }
        }
      }
    }
  }
  for (int t = tid; t < nprob * n; t += kRowThreads) {
    const int q = t / n, e = t % n;
    const int b = blockIdx.x + q * gridDim.x;
    if constexpr (FUSED) {
      // :315-323 undo the scaling, split the duals -- straight into the caller's tensors
      const size_t o = (size_t)b * n + e;
      const T d = prob_vec(q, 8)[e], rho = pscal[kPS * q];
      io.x[o] = d * prob_vec(q, 9)[e];
      io.z[o] = d * prob_vec(q, 2)[e];
      const T uu = prob_vec(q, 3)[e] / d;
      io.u[o] = uu;
      const T y = uu * rho;
      io.lams[(size_t)b * 2 * n + e] = (-y > T(0)) ? -y : T(0);
      io.lams[(size_t)b * 2 * n + n + e] = (y > T(0)) ? y : T(0);
      if (e == 0) io.rho_out[b] = rho;
    } else {
      const size_t vo = (size_t)b * ld + e;
      w.z[vo] = prob_vec(q, 2)[e];
      w.u[vo] = prob_vec(q, 3)[e];
      w.xs[vo] = prob_vec(q, 9)[e];
    }
  }
  for (int q = tid; q < nprob; q += kRowThreads) {      // record of the last check, rho (status kernel, relaunch, warm restarts)
    const int b = blockIdx.x + q * gridDim.x;
    for (int k = 0; k < 4; ++k) w.chk[4 * b + k] = pscal[kPS * q + 2 + k];
    w.ratio[b] = pscal[kPS * q + 6];
    w.wants[b] = pscal[kPS * q + 7] != T(0) ? 1 : 0;
    if (FUSED) w.rho[b] = pscal[kPS * q];
  }
  if (blockIdx.x == 0 && tid == 0) {
    ctrl->last_wants = last_wants;                 // flags of the most recent check (a relaunch after status 3 reads them)
    ctrl->last_ratio_out = last_rout;
    if (cfg.verbose && last_check >= 0) ctrl->n_log = min(last_check / check + 1, LQPB_LOG_CAP);
    ctrl->status = status;
    if (status == 3) ctrl->next_i = i;
    else ctrl->iter = i;
  }
}

// ---------------------------------------------------------------------------------------------
template <typename T>
static bool plan_rows(const FwdWs<T>& w, const lqpb_config& cfg, int max_smem, int n_sm, bool fused, RowGeom* out,
                      size_t* smem_bytes) {
  constexpr int VN = Vec<T>::N;
  const int n = w.n, m = w.m, N = n + m;
  RowGeom g{};
  g.fused = fused ? 1 : 0;
  g.nch = (n + VN - 1) / VN;
  g.ldv = g.nch * VN;
  g.P2 = 32;
  while (g.P2 < n || g.P2 < m) g.P2 <<= 1;
  const size_t budget = (size_t)max_smem / sizeof(T);
  auto geom = [&](int ppc, int gw, RowGeom* gg, size_t* total) {
    const int gthreads = gw * 32;
    // lanes per row: the cheapest of the powers of two (passes over the rows x chunk steps per lane)
    int best_lpr = 1;
    long best_cost = -1;
    for (int lpr = 1; lpr <= 32; lpr *= 2) {
      const int rpp = gthreads / lpr;
      const long passes = (n + rpp - 1) / rpp;
      const long steps = (g.nch + lpr - 1) / lpr;
      int lg = 0;
      while ((1 << lg) < lpr) ++lg;
      const long cost = passes * (steps * (2 + VN) + 2 * lg + 24);
      if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_lpr = lpr; }
    }
    gg->lpr = best_lpr;
    // row stride: a multiple of the 16-byte chunk with (stride in 4-byte banks) = 4 lpr mod 32 for lpr <= 8, so that
    // the 8 / lpr rows a quarter warp touches per LDS.128 phase fall into disjoint bank groups
    const int chunk_banks = 4;                         // 16 bytes
    int ldk_banks = ((fused ? N : n) + VN - 1) / VN * chunk_banks;
    if (best_lpr < 8) {
      const int want = (chunk_banks * best_lpr) % 32;
      while (ldk_banks % 32 != want) ldk_banks += chunk_banks;
    }
    gg->ldk = ldk_banks * 4 / (int)sizeof(T);
    gg->prob_elems = (size_t)(fused ? N + n + m : 2 * n) * gg->ldk + (size_t)kRowVecs * g.ldv;
    gg->prob_elems = (gg->prob_elems + VN - 1) / VN * VN;
    gg->group_elems = 6 * 16;
    if (fused) gg->group_elems += 16 * (sizeof(double) / sizeof(T)) + g.P2 + gg->ldk;
    gg->group_elems = (gg->group_elems + VN - 1) / VN * VN;
    const int ng = kRowWarps / gw;
    *total = (size_t)ppc * gg->prob_elems + (size_t)ng * gg->group_elems + kPS * (size_t)ppc + 64;
    return *total <= budget;
  };
  // warps per problem: enough lanes for ~one pass with 4 lanes per row, at most the whole CTA
  int gw = 1;
  while (gw < kRowWarps && gw * 32 < 2 * n) gw *= 2;
  const bool chatty = cfg.check_solved <= 2;
  int grids[2], ngr = 0;
  if (chatty) {
    // one cluster of <= 8 CTAs: the smallest grid in which every group owns at most one problem, else the largest
    // that fits
    int pick = 0;
    for (int G = 1; G <= 8; G *= 2) {
      RowGeom t = g;
      size_t tot;
      const int ppc = (w.B + G - 1) / G;
      if (!geom(ppc, gw, &t, &tot)) continue;
      pick = G;
      if (ppc <= kRowWarps / gw) break;
    }
    if (pick) grids[ngr++] = pick;
  }
  grids[ngr++] = w.B < n_sm ? w.B : n_sm;
  for (int c = 0; c < ngr; ++c) {
    const int G = grids[c];
    const int ppc = (w.B + G - 1) / G;
    int gwx = gw;
    while (gwx < kRowWarps && kRowWarps / gwx > ppc) gwx *= 2;      // no idle groups: fewer, wider groups
    RowGeom t = g;
    size_t tot;
    if (!geom(ppc, gwx, &t, &tot)) continue;
    t.G = G; t.ppc = ppc; t.gw = gwx; t.ngroups = kRowWarps / gwx;
    t.cluster = (G <= 8) ? 1 : 0;
    *smem_bytes = tot * sizeof(T);
    *out = t;
    return true;
  }
  return false;
}

template <typename T, bool FUSED>
static cudaError_t launch_rows(const lqpb_config& cfg, const FwdWs<T>& w, int i0, int skip_rho_check, T* nus_out,
                               const FusedIO<T>& io, int* launches, cudaStream_t st, bool* taken) {
  *taken = false;
  int dev = 0, max_smem = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  RowGeom geo{};
  size_t smem = 0;
  if (!plan_rows(w, cfg, max_smem - 2048, sms, FUSED, &geo, &smem)) return cudaSuccess;
  e = cudaMemsetAsync(&w.ctrl->barrier, 0, sizeof(unsigned), st);
  if (e != cudaSuccess) return e;
  void* kern = (void*)iterate_row_kernel<T, FUSED>;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  lqpb_config c = cfg;
  FwdWs<T> ww = w;
  FusedIO<T> ioc = io;
  if (geo.cluster) {
    cudaLaunchConfig_t lc{};
    lc.gridDim = dim3(geo.G);
    lc.blockDim = dim3(kRowThreads);
    lc.dynamicSmemBytes = smem;
    lc.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = geo.G;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    lc.attrs = at;
    lc.numAttrs = 1;
    e = cudaLaunchKernelEx(&lc, iterate_row_kernel<T, FUSED>, c, ww, i0, skip_rho_check, nus_out, geo, ioc);
  } else {
    void* args[] = {&c, &ww, &i0, &skip_rho_check, &nus_out, &geo, &ioc};
    e = cudaLaunchCooperativeKernel(kern, dim3(geo.G), dim3(kRowThreads), args, smem, st);
  }
  if (e != cudaSuccess) return e;
  if (launches) ++*launches;
  *taken = true;
  return cudaSuccess;
}

template <typename T>
cudaError_t launch_iterate_rows(const lqpb_config& cfg, const FwdWs<T>& w, int i0, int skip_rho_check, T* nus_out,
                                int* launches, cudaStream_t st, bool* taken) {
  *taken = false;
  const char* e = getenv("LQPB_ITER");          // developer switch (A/B measurements): auto | rows | packed | stream
  if (e && (!strcmp(e, "stream") || !strcmp(e, "packed"))) return cudaSuccess;
  return launch_rows<T, false>(cfg, w, i0, skip_rho_check, nus_out, FusedIO<T>{}, launches, st, taken);
}

// The whole forward solve of a small problem in one launch (FUSED kernel).  The caller has zeroed the control block and
// run bound_flags_kernel (any_lb / any_ub of the whole batch) on the same stream.  *taken = false: does not apply.
template <typename T>
cudaError_t launch_forward_fused(const lqpb_config& cfg, const FwdWs<T>& w, const T* Q, const T* p, const T* A, const T* b,
                                 const T* lb, const T* ub, T* x, T* z, T* u, T* lams, T* nus, T* rho_out, int* launches,
                                 cudaStream_t st, bool* taken) {
  *taken = false;
  const char* e = getenv("LQPB_FUSED");         // developer switch (A/B measurements): 0 = separate kernels
  if (e && e[0] == '0') return cudaSuccess;
  const char* it = getenv("LQPB_ITER");
  if (it && strcmp(it, "auto") && strcmp(it, "rows")) return cudaSuccess;
  const FusedIO<T> io{Q, p, A, b, lb, ub, x, z, u, lams, rho_out};
  return launch_rows<T, true>(cfg, w, 0, 0, nus, io, launches, st, taken);
}

template <typename T>
bool forward_fused_applies(const lqpb_config& cfg, const FwdWs<T>& w) {
  const char* e = getenv("LQPB_FUSED");
  if (e && e[0] == '0') return false;
  const char* it = getenv("LQPB_ITER");
  if (it && strcmp(it, "auto") && strcmp(it, "rows")) return false;
  int dev = 0, max_smem = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  RowGeom geo{};
  size_t smem = 0;
  return plan_rows(w, cfg, max_smem - 2048, sms, true, &geo, &smem);
}
template <typename T>
bool iterate_rows_applies(const lqpb_config& cfg, const FwdWs<T>& w) {
  const char* it = getenv("LQPB_ITER");
  if (it && (!strcmp(it, "stream") || !strcmp(it, "packed"))) return false;
  int dev = 0, max_smem = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  RowGeom geo{};
  size_t smem = 0;
  return plan_rows(w, cfg, max_smem - 2048, sms, false, &geo, &smem);
}
template bool iterate_rows_applies<float>(const lqpb_config&, const FwdWs<float>&);
template bool iterate_rows_applies<double>(const lqpb_config&, const FwdWs<double>&);
template bool forward_fused_applies<float>(const lqpb_config&, const FwdWs<float>&);
template bool forward_fused_applies<double>(const lqpb_config&, const FwdWs<double>&);

#define INST(T)                                                                                                      \
  template cudaError_t launch_iterate_rows<T>(const lqpb_config&, const FwdWs<T>&, int, int, T*, int*, cudaStream_t, \
                                              bool*);                                                                \
  template cudaError_t launch_forward_fused<T>(const lqpb_config&, const FwdWs<T>&, const T*, const T*, const T*,    \
                                               const T*, const T*, const T*, T*, T*, T*, T*, T*, T*, int*,           \
                                               cudaStream_t, bool*);
INST(float)
INST(double)
#undef INST

}  // namespace lqpb
