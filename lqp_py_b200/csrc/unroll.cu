// Unrolled mode (control['unroll'] = True): the reverse sweep through the recorded ADMM iterations.
//
// The reference differentiates the loop of lqp_py/solve_box_qp_admm_torch.py:259-282 with autograd; every linear
// solve goes through TorchLULayer (lqp_py/lu_layer.py:18-58), whose backward is
//     dx = M^-1 (-g),   dl_dA = dx xv^T,   dl_db = -dx                                   (lu_layer.py:52-54)
// With (z_k, u_k) the state after iteration k, xv_k = [x_k; nu_k] = M^-1 [-p~ + rho (z_{k-1} - u_{k-1}); b~],
// z_k = clamp(x_k + u_{k-1}), u_k = u_{k-1} + x_k - z_k, the adjoint recursion from the last iteration down is
//     h    = gz - gu                       (adjoint of z_k once u_k = u_{k-1} + x_k - z_k is undone)
//     t    = h where z_k is strictly inside the box, else 0;  the clamped entries send h to lb~ / ub~ (ties: halves)
//     gx   = gu + t  (+ the incoming adjoint of the returned x~ at the last iteration)
//     [w; wnu] = M^-1 [gx; 0] = [K11 gx; K21 gx]          <- the same symmetric GEMV the forward iteration streams
//     gp~ -= w,  gb~ += wnu,  grho += w . (z_{k-1} - u_{k-1}) - w . x_k   (rhs term and the trace of dM = -w xv^T)
//     gz   = rho w,   gu = gu + t - rho w
// and the matrix adjoints are rank-n_iter products over the tape, formed once at the end instead of one dense
// N x N outer product per iteration as in the reference:
//     dQ~ = - sum_k w_k x_k^T   (NOT symmetric, like the reference's),   dA~ = - sum_k (wnu_k x_k^T + nu_k w_k^T).
// A sweep covers the iteration range [k_lo, k_hi] of one operator segment (rho and K11 constant): it takes the
// adjoints of (x, z, u)_{k_hi} and z_{k_hi - 1} and returns those of (z, u)_{k_lo - 1}, so that the host can chain
// segments around an adaptive-rho update (reference :237-256), whose ratio reads the state of the last check.
// Problems are independent here (no stop test): one persistent CTA per problem, K11 streamed from its packed lower
// triangle through the per-warp bulk-TMA rings of the forward kernel (itergeom.cuh), prefetching across iterations.
#include "itergeom.cuh"

namespace lqpb {

template <typename T>
__global__ void __launch_bounds__(kIterMaxThreads, 1)
unroll_reverse_kernel(FwdWs<T> w, Tape<T> tape, UnrollGrads<T> g, IterGeom geo) {
  using P = Pack<T>;
  constexpr int TC = P::TC, TILE = P::TILE;
  using V4 = typename Vec<T>::type;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int n = w.n, m = w.m, ld = w.ld, np = geo.np, K = tape.n_iter;     // K: rows of the tape per problem
  const int k_lo = g.k_lo, k_hi = g.k_hi, nk = g.k_hi - g.k_lo + 1;            // iterations swept here, k_hi down to k_lo
  const int nthreads = blockDim.x, nwarps = geo.nwarps, depth = geo.depth;
  const int ntv = geo.nt, ntiles = geo.ntiles;
  T* ring = reinterpret_cast<T*>(smem_raw);                 // [nwarps][depth][TILE]
  T* xpart = ring + (size_t)nwarps * depth * TILE;          // [nwarps][np] per-warp partial sums of K11 gx
  T* v = xpart + (size_t)nwarps * np;                       // [np] gx of this iteration (zero padded)
  T* gz = v + np;                                           // [np] adjoint of z_k
  T* gu = gz + np;                                          // [np] adjoint of u_k
  T* tdot = gu + np;                                        // [max(m,1)] K21 gx
  T* red = tdot + (m > 0 ? round_up(m, 4) : 4);             // [6][16] reduction scratch
  uint64_t* full = reinterpret_cast<uint64_t*>(red + 6 * 16 + 4);   // [nwarps][depth]

  const int tid = threadIdx.x;
  const int wid = tid >> 5, lane = tid & 31;
  const int nprob = (w.B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (tid == 0) {
    for (int s = 0; s < nwarps * depth; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  for (int e = tid; e < nwarps * np; e += nthreads) xpart[e] = T(0);
  for (int e = tid; e < np; e += nthreads) { v[e] = T(0); gz[e] = T(0); gu[e] = T(0); }
  __syncthreads();

  const bool any_lb = w.ctrl->any_lb != 0, any_ub = w.ctrl->any_ub != 0;

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

  // ---- tile stream of this warp: for problem: for iteration: its K11 run; fetched `depth` tiles ahead
  int p_k = 0, p_it = 0, p_r = 0, p_slot = 0;
  int c_slot = 0;
  uint32_t c_phase = 0;
  int in_flight = 0;
  uint64_t pol_keep = 0;
  if (lane == 0) pol_keep = l2_policy_evict_last();
  auto issue_next = [&]() {
    if (run_len == 0 || p_k >= nprob) return;
    if (lane == 0) {
      const int b = blockIdx.x + p_k * gridDim.x;
      const T* src = w.Kp + ((size_t)b * ntiles + run_lo + p_r) * TILE;
      mbar_arrive_expect_tx(&full_w[p_slot], (uint32_t)(TILE * sizeof(T)));
      tma_load_1d_hint(ring_w + (size_t)p_slot * TILE, src, (uint32_t)(TILE * sizeof(T)), &full_w[p_slot], pol_keep);
    }
    ++in_flight;
    if (++p_slot == depth) p_slot = 0;
    if (++p_r == run_len) {
      p_r = 0;
      if (++p_it == nk) { p_it = 0; ++p_k; }
    }
  };
  for (int d = 0; d < depth; ++d) issue_next();

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
      __syncwarp();
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

  for (int kp = 0; kp < nprob; ++kp) {
    const int b = blockIdx.x + kp * gridDim.x;
    const size_t vo = (size_t)b * ld, go = (size_t)b * n;
    const size_t tb = (size_t)b * K * n;
    const T rho = w.rho[b];
    const T* lbt = w.lbt + vo;
    const T* ubt = w.ubt + vo;
    T grho_acc = T(0), gb_acc = T(0);

    // adjoint of (x_k, z_k, u_k) folded into gx: the elementwise head of iteration k (needs gz, gu of k + 1)
    auto head = [&](int k, int e, T gz_e, T gu_e) {
      // z_k = min(max(v, lb), ub), v = x_k + u_{k-1} (:272-276), with torch's rule for ties: maximum / minimum send
      // half of the gradient to each argument when they are equal (a pinned variable has lb == ub)
      const size_t to = tb + (size_t)k * n + e;
      const T vk = tape.x[to] + (k > 0 ? tape.u[to - n] : T(0));
      const T h = gz_e - gu_e;
      T z1 = vk, c_v = T(1), c_lb = T(0), c_ub = T(0);
      if (any_lb) {
        const T l = lbt[e];
        c_v = vk > l ? T(1) : (vk == l ? T(0.5) : T(0));
        c_lb = T(1) - c_v;
        z1 = t_max(vk, l);
      }
      if (any_ub) {
        const T ub_e = ubt[e];
        const T c_z1 = z1 < ub_e ? T(1) : (z1 == ub_e ? T(0.5) : T(0));
        c_ub = T(1) - c_z1;
        c_v *= c_z1;
        c_lb *= c_z1;
      }
      const T t = c_v * h;
      if (c_ub != T(0)) g.gub[go + e] += c_ub * h;
      if (c_lb != T(0)) g.glb[go + e] += c_lb * h;
      v[e] = gu_e + t + ((k == k_hi && g.gx) ? g.gx[go + e] : T(0));
      gu[e] = gu_e + t;
    };
    for (int e = tid; e < n; e += nthreads) {
      g.gp[go + e] = T(0);
      g.glb[go + e] = T(0);
      g.gub[go + e] = T(0);
      head(k_hi, e, g.gz_last ? g.gz_last[go + e] : T(0), g.gu_last ? g.gu_last[go + e] : T(0));
    }
    __syncthreads();

    for (int k = k_hi; k >= k_lo; --k) {
      sym_pass(v);                                            // xpart += K11 gx
      if (m > 0) {                                            // wnu = K21 gx
        const T* Gt = w.Gt + (size_t)b * m * ld;
        for (int l = wid; l < m; l += nwarps) {
          T d = T(0);
          for (int e = lane; e < n; e += 32) d += Gt[(size_t)l * ld + e] * v[e];
          d = warp_sum(d);
          if (lane == 0) tdot[l] = d;
        }
      }
      __syncthreads();
      for (int e = tid; e < n; e += nthreads) {
        T we = T(0);
        for (int ww = 0; ww < nwarps; ++ww) {
          we += xpart[(size_t)ww * np + e];
          xpart[(size_t)ww * np + e] = T(0);
        }
        g.tw[tb + (size_t)k * n + e] = we;
        g.gp[go + e] -= we;
        T zu = T(0);
        if (k > 0) zu = tape.z[tb + (size_t)(k - 1) * n + e] - tape.u[tb + (size_t)(k - 1) * n + e];
        grho_acc += we * (zu - tape.x[tb + (size_t)k * n + e]);
        T gz_e = rho * we;
        if (k == k_hi && g.gzprev_last) gz_e += g.gzprev_last[go + e];   // z_{k_hi - 1} is also an output of the segment
        const T gu_e = gu[e] - rho * we;
        if (k > k_lo) {
          head(k - 1, e, gz_e, gu_e);
        } else {                                                        // adjoints of the state the segment started from
          if (g.gz_in) g.gz_in[go + e] = gz_e;
          if (g.gu_in) g.gu_in[go + e] = gu_e;
        }
      }
      if (m > 0 && tid < m) {
        g.twnu[((size_t)b * K + k) * m + tid] = tdot[tid];
        gb_acc += tdot[tid];
      }
      __syncthreads();
    }
    // ---- per-problem scalars
    grho_acc = warp_sum(grho_acc);
    if (lane == 0) red[wid] = grho_acc;
    __syncthreads();
    if (tid == 0) {
      T s = T(0);
      for (int ww = 0; ww < nwarps; ++ww) s += red[ww];
      g.grho[b] = s;
    }
    if (m > 0 && tid < m) g.gb[(size_t)b * m + tid] = gb_acc;
    for (int e = tid; e < np; e += nthreads) { gz[e] = T(0); gu[e] = T(0); v[e] = T(0); }
    __syncthreads();
  }
  while (in_flight > 0) {
    mbar_wait(&full_w[c_slot], c_phase);
    if (++c_slot == depth) { c_slot = 0; c_phase ^= 1u; }
    --in_flight;
  }
}

// ---------------------------------------------------------------------------------------------
// out[b][i][j] = - sum_k ( U[b][k][i] V[b][k][j] + U2[b][k][i] V2[b][k][j] ),  i < rows, j < cols, k < K.
// 64 x 64 output tile per CTA, 4 x 4 per thread, the K dimension staged through shared memory 16 at a time.
template <typename T>
__global__ void __launch_bounds__(256)
tape_outer_kernel(const T* __restrict__ U, int su, const T* __restrict__ V, int sv, const T* __restrict__ U2, int su2,
                  const T* __restrict__ V2, int sv2, int rows, int cols, int kb, int k_lo, int K, T* __restrict__ out) {
  // kb = tape rows per problem (batch stride), the sum runs over the K rows k_lo .. k_lo + K - 1
  constexpr int TB = 64, KC = 16;
  __shared__ T us[KC][TB + 4], vs[KC][TB + 4];
  const int b = blockIdx.z;
  const int i0 = blockIdx.y * TB, j0 = blockIdx.x * TB;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  T acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[a][c] = T(0);
  const int npairs = U2 ? 2 : 1;
  for (int pr = 0; pr < npairs; ++pr) {
    const T* Ub = pr == 0 ? U + ((size_t)b * kb + k_lo) * su : U2 + ((size_t)b * kb + k_lo) * su2;
    const T* Vb = pr == 0 ? V + ((size_t)b * kb + k_lo) * sv : V2 + ((size_t)b * kb + k_lo) * sv2;
    const int s_u = pr == 0 ? su : su2, s_v = pr == 0 ? sv : sv2;
    for (int k0 = 0; k0 < K; k0 += KC) {
      for (int t = threadIdx.x; t < KC * TB; t += 256) {
        const int kk = t / TB, c = t % TB;
        const bool kin = k0 + kk < K;
        us[kk][c] = (kin && i0 + c < rows) ? Ub[(size_t)(k0 + kk) * s_u + i0 + c] : T(0);
        vs[kk][c] = (kin && j0 + c < cols) ? Vb[(size_t)(k0 + kk) * s_v + j0 + c] : T(0);
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < KC; ++kk) {
        T ua[4], va[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) { ua[a] = us[kk][ty * 4 + a]; va[a] = vs[kk][tx * 4 + a]; }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[a][c] += ua[a] * va[c];
      }
      __syncthreads();
    }
  }
  T* ob = out + (size_t)b * rows * cols;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int i = i0 + ty * 4 + a;
    if (i >= rows) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int j = j0 + tx * 4 + c;
      if (j < cols) ob[(size_t)i * cols + j] = -acc[a][c];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Adjoint of the scaling of Q and of the rho selection (reference :176, :200-203), fused:  with G the adjoint of
// Q~ = D Q D delivered by the sweep and  c = grho / (n rho)  (zero where rho was clamped),
//     Gt = G + c Q~,     gQ_ij = D_i Gt_ij D_j   (written over G),     gD_j = sum_i Gt_ij Q_ij D_i + sum_i Gt_ji Q_ji D_i.
// grid = (row blocks of 32, B); a warp owns 4 rows: the row sums (second term, index j = row) are warp reductions,
// the column sums (first term) go through per-warp slices of shared memory and leave as one partial row per CTA;
// scale_grad_reduce_kernel adds the partials in a fixed order (deterministic).  D == nullptr: no scaling (D = 1).
constexpr int kSgThreads = 256, kSgRows = 32;

template <typename T>
__global__ void __launch_bounds__(kSgThreads)
scale_grad_kernel(T* __restrict__ G, const T* __restrict__ Q, const T* __restrict__ D, const T* __restrict__ coef,
                  T* __restrict__ part, int n, int nblk) {
  extern __shared__ __align__(16) unsigned char sg_smem[];
  T* colacc = reinterpret_cast<T*>(sg_smem);            // [8 warps][n]
  const int b = blockIdx.y, blk = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const T* Db = D ? D + (size_t)b * n : nullptr;
  const T c = coef ? coef[b] : T(0);
  T* Gb = G + (size_t)b * n * n;
  const T* Qb = Q + (size_t)b * n * n;
  T* ca = colacc + (size_t)warp * n;
  for (int j = lane; j < n; j += 32) ca[j] = T(0);
  T* rowpart = part + ((size_t)b * (nblk + 1) + nblk) * n;       // last partial row: the row sums
  for (int r = 0; r < kSgRows / 8; ++r) {
    const int i = blk * kSgRows + warp * (kSgRows / 8) + r;
    if (i >= n) break;
    const T di = Db ? Db[i] : T(1);
    T racc = T(0);
    for (int j = lane; j < n; j += 32) {
      const size_t o = (size_t)i * n + j;
      const T q = Qb[o], dj = Db ? Db[j] : T(1);
      const T gt = Gb[o] + c * (di * q * dj);
      Gb[o] = di * gt * dj;
      const T t = gt * q;
      ca[j] += t * di;
      racc += t * dj;
    }
    racc = warp_sum(racc);
    if (lane == 0) rowpart[i] = racc;
  }
  __syncthreads();
  T* cp = part + ((size_t)b * (nblk + 1) + blk) * n;
  for (int j = tid; j < n; j += kSgThreads) {
    T s = T(0);
#pragma unroll
    for (int w8 = 0; w8 < kSgThreads / 32; ++w8) s += colacc[(size_t)w8 * n + j];
    cp[j] = s;
  }
}

template <typename T>
__global__ void scale_grad_reduce_kernel(const T* __restrict__ part, T* __restrict__ gD, int n, int nblk) {
  const int b = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const T* pb = part + (size_t)b * (nblk + 1) * n;
  T s = pb[(size_t)nblk * n + j];
  for (int k = 0; k < nblk; ++k) s += pb[(size_t)k * n + j];
  gD[(size_t)b * n + j] = s;
}

template <typename T>
cudaError_t launch_scale_grad(int B, int n, T* G, const T* Q, const T* D, const T* coef, T* gD, T* part, cudaStream_t st) {
  const int nblk = (n + kSgRows - 1) / kSgRows;
  const size_t smem = (size_t)(kSgThreads / 32) * n * sizeof(T);
  cudaError_t e = cudaFuncSetAttribute(scale_grad_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  dim3 grid(nblk, B);
  scale_grad_kernel<T><<<grid, kSgThreads, smem, st>>>(G, Q, D, coef, part, n, nblk);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  if (gD) {
    dim3 g2((n + 127) / 128, B);
    scale_grad_reduce_kernel<T><<<g2, 128, 0, st>>>(part, gD, n, nblk);
  }
  return cudaGetLastError();
}
template cudaError_t launch_scale_grad<float>(int, int, float*, const float*, const float*, const float*, float*, float*, cudaStream_t);
template cudaError_t launch_scale_grad<double>(int, int, double*, const double*, const double*, const double*, double*, double*, cudaStream_t);

// ---------------------------------------------------------------------------------------------
// The O(B n) part of the scaling map (:161-197) for the unrolled mode: what autograd walks through between the caller's
// (p, A, b, lb, ub, column norms of Q) and the scaled problem (D, p~, A~, b~, lb~, ub~).
//   scaled_vectors_kernel  -- forward VALUES: copied out of the workspace of the recording solve (the kernels computed
//                             them already; same numbers the loop used)
//   scale_vec_grad_kernel  -- the adjoint of the whole map in one kernel per problem, with torch's (sub)gradient rules
//                             for the operators the reference uses: inf-norms split their gradient evenly among exact
//                             ties, torch.quantile sends (1 - w, w) to the two order statistics it interpolates,
//                             where(bad, maximum(norm, floor), norm) routes the guarded entries to the mean
template <typename T>
__global__ void scaled_vectors_kernel(FwdWs<T> w, T* D, T* pt, T* At, T* bt, T* lbt, T* ubt, T* E) {
  const int b = blockIdx.y, n = w.n, m = w.m, ld = w.ld;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const size_t vo = (size_t)b * ld + j, o = (size_t)b * n + j;
    D[o] = w.D[vo];
    pt[o] = w.pt[vo];
    lbt[o] = w.lbt[vo];
    ubt[o] = w.ubt[vo];
    for (int l = 0; l < m; ++l) At[((size_t)b * m + l) * n + j] = w.At[((size_t)b * m + l) * ld + j];
  }
  if (blockIdx.x == 0)
    for (int l = threadIdx.x; l < m; l += blockDim.x) {
      bt[(size_t)b * m + l] = w.bt[(size_t)b * m + l];
      E[(size_t)b * m + l] = w.E[(size_t)b * m + l];
    }
}

constexpr int kSvThreads = 256;

template <typename T>
__global__ void __launch_bounds__(kSvThreads) scale_vec_grad_kernel(ScaleVecGrad<T> a, int P2) {
  extern __shared__ __align__(16) unsigned char sv_raw[];
  double* gDt = reinterpret_cast<double*>(sv_raw);             // [n] total adjoint of D, then of D0
  double* dscr = gDt + a.n;                                    // [32]
  T* D0 = reinterpret_cast<T*>(dscr + 32);                     // [n]
  T* cg = D0 + a.n;                                            // [n] guarded column norms
  T* skey = cg + a.n;                                          // [P2]
  int* sidx = reinterpret_cast<int*>(skey + P2);               // [P2]
  __shared__ double s_tot;
  const int b = blockIdx.x, tid = threadIdx.x, n = a.n, m = a.m;
  const size_t vo = (size_t)b * n;
  auto bsum = [&](double v) -> double {
    v = warp_sum(v);
    __syncthreads();
    if ((tid & 31) == 0) dscr[tid >> 5] = v;
    __syncthreads();
    if (tid == 0) {
      double r = 0.0;
      for (int k = 0; k < kSvThreads / 32; ++k) r += dscr[k];
      s_tot = r;
    }
    __syncthreads();
    return s_tot;
  };
  // ---- forward quantities again: guarded norms, D0, beta, mean
  double part = 0.0;
  for (int j = tid; j < n; j += kSvThreads) part += (double)a.colmax[vo + j];
  const double mean_c = bsum(part) / n;
  const T floor_c = t_max((T)mean_c, T(1e-6));
  for (int j = tid; j < n; j += kSvThreads) {
    T c = a.colmax[vo + j];
    if (c <= T(0)) c = t_max(c, floor_c);
    cg[j] = c;
    D0[j] = t_sqrt(T(1) / c);
  }
  __syncthreads();
  T beta = a.beta, q10 = T(0), q90 = T(1);
  int ilo[2] = {0, 0}, ihi[2] = {0, 0};
  T wq[2] = {T(0), T(0)};
  if (a.beta_auto) {
    for (int j = tid; j < P2; j += kSvThreads) { skey[j] = j < n ? D0[j] : t_inf<T>(); sidx[j] = j; }
    __syncthreads();
    for (int k = 2; k <= P2; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = tid; i < P2; i += kSvThreads) {
          const int ixj = i ^ j;
          if (ixj > i) {
            const T x = skey[i], y = skey[ixj];
            const bool asc = (i & k) == 0;
            if ((x > y) == asc) {
              skey[i] = y; skey[ixj] = x;
              const int t = sidx[i]; sidx[i] = sidx[ixj]; sidx[ixj] = t;
            }
          }
        }
        __syncthreads();
      }
    const T qs[2] = {T(0.10), T(0.90)};
    T qv[2];
    for (int k = 0; k < 2; ++k) {
      const T rank = qs[k] * T(n - 1);
      const T lo = floor(rank), hi = ceil(rank);
      const T x = skey[(int)lo], y = skey[(int)hi], wgt = rank - lo;
      qv[k] = wgt < T(0.5) ? x + wgt * (y - x) : y - (y - x) * (T(1) - wgt);
      ilo[k] = sidx[(int)lo]; ihi[k] = sidx[(int)hi]; wq[k] = wgt;
    }
    q10 = qv[0]; q90 = qv[1];
    beta = T(1) - q10 / q90;
  }
  part = 0.0;
  for (int j = tid; j < n; j += kSvThreads) part += (double)D0[j];
  const double meanD0 = bsum(part) / n;

  // ---- adjoint of the vectors: everything that reaches D, and the gradients of p, lb, ub
  for (int j = tid; j < n; j += kSvThreads) {
    const T d = a.D[vo + j];
    double g = (a.gD ? (double)a.gD[vo + j] : 0.0) + (a.gD2 ? (double)a.gD2[vo + j] : 0.0);
    if (a.gpt) {
      g += (double)(a.gpt[vo + j] * a.p[vo + j]);
      a.gp[vo + j] = a.gpt[vo + j] * d;
    } else {
      a.gp[vo + j] = T(0);
    }
    if (a.use_lb && a.glbt) {
      const T gl = a.glbt[vo + j];
      g += (double)(gl * (-a.lb[vo + j] / (d * d)));      // 0 * inf = NaN on an infinite entry, like torch's division
      a.glb[vo + j] = gl / d;
    } else {
      a.glb[vo + j] = T(0);
    }
    if (a.use_ub && a.gubt) {
      const T gu = a.gubt[vo + j];
      g += (double)(gu * (-a.ub[vo + j] / (d * d)));
      a.gub[vo + j] = gu / d;
    } else {
      a.gub[vo + j] = T(0);
    }
    gDt[j] = g;
  }
  __syncthreads();
  // ---- equality rows: A~ = E (A D), b~ = E b, E = 1 / guard(||A D||_inf per row)
  if (m > 0) {
    // raw row norms, their mean (for the guard) -- m is small, one row at a time
    double rsum = 0.0;
    for (int l = 0; l < m; ++l) {
      const T* Al = a.A + ((size_t)b * m + l) * n;
      T mx = T(0);
      for (int j = tid; j < n; j += kSvThreads) mx = t_max(mx, t_abs(Al[j] * a.D[vo + j]));
      mx = warp_max(mx);
      __syncthreads();
      if ((tid & 31) == 0) dscr[tid >> 5] = (double)mx;
      __syncthreads();
      double r = dscr[0];
      for (int k = 1; k < kSvThreads / 32; ++k) r = r > dscr[k] ? r : dscr[k];
      skey[l % P2] = (T)r;         // (reuse: m <= P2 is guaranteed by the launcher)
      rsum += r;
      __syncthreads();
    }
    const bool mean_active = (T)(rsum / m) > T(1e-6);
    double gfloor_r = 0.0;
    for (int l = 0; l < m; ++l) {
      const T* Al = a.A + ((size_t)b * m + l) * n;
      const T e = a.E[(size_t)b * m + l], rraw = skey[l % P2];
      // gE_l = sum_j gAt_lj A_lj D_j + gbt_l b_l
      part = 0.0;
      if (a.gAt)
        for (int j = tid; j < n; j += kSvThreads) part += (double)(a.gAt[((size_t)b * m + l) * n + j] * (Al[j] * a.D[vo + j]));
      double gE = bsum(part);
      if (a.gbt) gE += (double)(a.gbt[(size_t)b * m + l] * a.b[(size_t)b * m + l]);
      const double gr = -gE * (double)e * (double)e;            // E = 1 / r
      const bool bad = rraw <= T(0);
      if (bad) gfloor_r += gr;                                  // maximum(r, floor) picks the floor
      // ties of the inf-norm split the gradient evenly
      part = 0.0;
      if (!bad)
        for (int j = tid; j < n; j += kSvThreads) part += (t_abs(Al[j] * a.D[vo + j]) == rraw) ? 1.0 : 0.0;
      const double cnt = bad ? 1.0 : bsum(part);
      for (int j = tid; j < n; j += kSvThreads) {
        const T ad = Al[j] * a.D[vo + j];
        double gad = 0.0;                                       // adjoint of (A D)_lj
        if (a.gAt) gad += (double)(a.gAt[((size_t)b * m + l) * n + j] * e);
        if (!bad && t_abs(ad) == rraw) gad += gr * (ad > T(0) ? 1.0 : (ad < T(0) ? -1.0 : 0.0)) / cnt;
        a.gA[((size_t)b * m + l) * n + j] = (T)(gad * (double)a.D[vo + j]);
        gDt[j] += gad * (double)Al[j];
      }
      if (tid == 0) a.gb[(size_t)b * m + l] = a.gbt ? a.gbt[(size_t)b * m + l] * e : T(0);
      __syncthreads();
    }
    // guarded rows send their gradient to the mean of the raw norms, i.e. evenly to every row's maximisers
    if (gfloor_r != 0.0 && mean_active) {
      for (int l = 0; l < m; ++l) {
        const T* Al = a.A + ((size_t)b * m + l) * n;
        const T rraw = skey[l % P2];
        part = 0.0;
        for (int j = tid; j < n; j += kSvThreads) part += (t_abs(Al[j] * a.D[vo + j]) == rraw) ? 1.0 : 0.0;
        const double cnt = bsum(part);
        for (int j = tid; j < n; j += kSvThreads) {
          const T ad = Al[j] * a.D[vo + j];
          if (t_abs(ad) == rraw) {
            const double gad = (gfloor_r / m) * (ad > T(0) ? 1.0 : (ad < T(0) ? -1.0 : 0.0)) / cnt;
            a.gA[((size_t)b * m + l) * n + j] += (T)(gad * (double)a.D[vo + j]);
            gDt[j] += gad * (double)Al[j];
          }
        }
        __syncthreads();
      }
    }
  }
  __syncthreads();
  // ---- D = (1 - beta) D0 + beta mean(D0), beta = 1 - q10 / q90
  part = 0.0;
  double part2 = 0.0;
  for (int j = tid; j < n; j += kSvThreads) {
    part += gDt[j];
    part2 += gDt[j] * (meanD0 - (double)D0[j]);
  }
  const double S = bsum(part);
  const double gbeta = bsum(part2);
  for (int j = tid; j < n; j += kSvThreads) gDt[j] = (1.0 - (double)beta) * gDt[j] + (double)beta * S / n;
  __syncthreads();
  if (a.beta_auto && tid == 0) {
    const double gq10 = -gbeta / (double)q90, gq90 = gbeta * (double)q10 / ((double)q90 * (double)q90);
    const double gq[2] = {gq10, gq90};
    for (int k = 0; k < 2; ++k) {
      gDt[ilo[k]] += gq[k] * (1.0 - (double)wq[k]);
      gDt[ihi[k]] += gq[k] * (double)wq[k];
    }
  }
  __syncthreads();
  // ---- D0 = sqrt(1 / c'), guard
  part = 0.0;
  for (int j = tid; j < n; j += kSvThreads) {
    const double gc = gDt[j] * (-0.5) * (double)D0[j] / (double)cg[j];
    const bool bad = a.colmax[vo + j] <= T(0);
    if (bad) part += gc;
    gDt[j] = bad ? 0.0 : gc;
  }
  const double gfloor = bsum(part);
  const double spread = ((T)mean_c > T(1e-6)) ? gfloor / n : 0.0;
  for (int j = tid; j < n; j += kSvThreads) a.gcolmax[vo + j] = (T)(gDt[j] + spread);
}

// Column inf-norms of Q (:163) and their adjoint.  Both kernels: CTA = 32 columns of one problem, 8 warps; warp w walks rows
// w, w + 8, ... with lane = column (one 128-byte segment of a row per warp access), the per-warp results meet in shared
// memory.  colmax_scatter: G[i][j] += sign(Q_ij) g_j / (number of maximisers of column j) on every maximiser -- torch's
// inf-norm backward, exact ties included -- added IN PLACE to the dense adjoint of Q that scale_grad_kernel left in G:
// pass 1 counts the maximisers, pass 2 (same traversal, the rows are L2 hits) writes the few entries that are maximisers.
constexpr int kCmWarps = 8;
template <typename T>
__global__ void __launch_bounds__(32 * kCmWarps) colmax_plain_kernel(const T* __restrict__ Q, T* __restrict__ out, int n) {
  __shared__ T part[kCmWarps][32];
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, j = blockIdx.x * 32 + lane;
  const T* Qb = Q + (size_t)b * n * n;
  T mx = T(0);
  if (j < n) {
#pragma unroll 4
    for (int i = warp; i < n; i += kCmWarps) mx = t_max(mx, t_abs(Qb[(size_t)i * n + j]));
  }
  part[warp][lane] = mx;
  __syncthreads();
  if (warp == 0 && j < n) {
#pragma unroll
    for (int w = 1; w < kCmWarps; ++w) mx = t_max(mx, part[w][lane]);
    out[(size_t)b * n + j] = mx;
  }
}
template <typename T>
__global__ void __launch_bounds__(32 * kCmWarps) colmax_scatter_kernel(const T* __restrict__ Q, const T* __restrict__ colmax,
                                                                      const T* __restrict__ g, T* __restrict__ G, int n) {
  __shared__ int part[kCmWarps][32];
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, j = blockIdx.x * 32 + lane;
  const bool col = j < n;
  const T gj = col ? g[(size_t)b * n + j] : T(0), mx = col ? colmax[(size_t)b * n + j] : T(0);
  const bool live = col && gj != T(0) && mx > T(0);          // sign(0) = 0: zero columns carry no gradient
  const T* Qb = Q + (size_t)b * n * n;
  T* Gb = G + (size_t)b * n * n;
  int cnt = 0;
  if (live) {
#pragma unroll 4
    for (int i = warp; i < n; i += kCmWarps) cnt += (t_abs(Qb[(size_t)i * n + j]) == mx) ? 1 : 0;
  }
  part[warp][lane] = cnt;
  __syncthreads();
  cnt = 0;
#pragma unroll
  for (int w = 0; w < kCmWarps; ++w) cnt += part[w][lane];
  if (!live || cnt == 0) return;
  const T share = gj / (T)cnt;
#pragma unroll 4
  for (int i = warp; i < n; i += kCmWarps) {
    const T q = Qb[(size_t)i * n + j];
    if (t_abs(q) == mx) Gb[(size_t)i * n + j] += q > T(0) ? share : -share;
  }
}
template <typename T>
cudaError_t launch_colmax_plain(int B, int n, const T* Q, T* out, cudaStream_t st) {
  dim3 grid((n + 31) / 32, B);
  colmax_plain_kernel<T><<<grid, 32 * kCmWarps, 0, st>>>(Q, out, n);
  return cudaGetLastError();
}
template <typename T>
cudaError_t launch_colmax_scatter(int B, int n, const T* Q, const T* colmax, const T* g, T* G, cudaStream_t st) {
  dim3 grid((n + 31) / 32, B);
  colmax_scatter_kernel<T><<<grid, 32 * kCmWarps, 0, st>>>(Q, colmax, g, G, n);
  return cudaGetLastError();
}
template cudaError_t launch_colmax_plain<float>(int, int, const float*, float*, cudaStream_t);
template cudaError_t launch_colmax_plain<double>(int, int, const double*, double*, cudaStream_t);
template cudaError_t launch_colmax_scatter<float>(int, int, const float*, const float*, const float*, float*, cudaStream_t);
template cudaError_t launch_colmax_scatter<double>(int, int, const double*, const double*, const double*, double*, cudaStream_t);

template <typename T>
cudaError_t launch_scaled_vectors(const FwdWs<T>& w, T* D, T* pt, T* At, T* bt, T* lbt, T* ubt, T* E, cudaStream_t st) {
  dim3 grid((w.n + 255) / 256, w.B);
  scaled_vectors_kernel<T><<<grid, 256, 0, st>>>(w, D, pt, At, bt, lbt, ubt, E);
  return cudaGetLastError();
}
template <typename T>
cudaError_t launch_scale_vec_grad(int B, const ScaleVecGrad<T>& a, cudaStream_t st) {
  int P2 = 64;
  while (P2 < a.n || P2 < a.m) P2 <<= 1;
  const size_t smem = (size_t)(a.n + 32) * sizeof(double) + (size_t)(2 * a.n + P2) * sizeof(T) + (size_t)P2 * sizeof(int) + 64;
  cudaError_t e = cudaFuncSetAttribute(scale_vec_grad_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  scale_vec_grad_kernel<T><<<B, kSvThreads, smem, st>>>(a, P2);
  return cudaGetLastError();
}
template cudaError_t launch_scaled_vectors<float>(const FwdWs<float>&, float*, float*, float*, float*, float*, float*, float*, cudaStream_t);
template cudaError_t launch_scaled_vectors<double>(const FwdWs<double>&, double*, double*, double*, double*, double*, double*, double*, cudaStream_t);
template cudaError_t launch_scale_vec_grad<float>(int, const ScaleVecGrad<float>&, cudaStream_t);
template cudaError_t launch_scale_vec_grad<double>(int, const ScaleVecGrad<double>&, cudaStream_t);

template <typename T>
cudaError_t launch_unroll_reverse(const FwdWs<T>& w, const Tape<T>& tape, const UnrollGrads<T>& g, int* launches,
                                  cudaStream_t st) {
  int dev = 0, max_smem = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  size_t smem = 0;
  IterGeom geo{};
  if (!make_geom(w, max_smem - 1024, &geo, &smem)) return cudaErrorInvalidConfiguration;
  const int grid = w.B < sms ? w.B : sms;
  e = cudaFuncSetAttribute(unroll_reverse_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  unroll_reverse_kernel<T><<<grid, geo.nwarps * 32, smem, st>>>(w, tape, g, geo);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  if (launches) ++*launches;
  const int n = w.n, m = w.m, kb = tape.n_iter, K = g.k_hi - g.k_lo + 1;
  if (g.gQ) {       // dQ~ = - sum_k w_k x_k^T
    dim3 grid2((n + 63) / 64, (n + 63) / 64, w.B);
    tape_outer_kernel<T><<<grid2, 256, 0, st>>>(g.tw, n, tape.x, n, (const T*)nullptr, 0, (const T*)nullptr, 0, n, n, kb,
                                                g.k_lo, K, g.gQ);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (launches) ++*launches;
  }
  if (g.gA && m > 0) {   // dA~ = - sum_k (wnu_k x_k^T + nu_k w_k^T)
    dim3 grid2((n + 63) / 64, (m + 63) / 64, w.B);
    tape_outer_kernel<T><<<grid2, 256, 0, st>>>(g.twnu, m, tape.x, n, tape.nu, m, g.tw, n, m, n, kb, g.k_lo, K, g.gA);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (launches) ++*launches;
  }
  return cudaSuccess;
}

#define INST(T)                                                                                                   \
  template cudaError_t launch_unroll_reverse<T>(const FwdWs<T>&, const Tape<T>&, const UnrollGrads<T>&, int*, \
                                                cudaStream_t);
INST(float)
INST(double)
#undef INST

}  // namespace lqpb
