// K1 -- problem scaling, rho candidate and state reset: three kernels, so that the two passes over Q are spread
// over thousands of CTAs instead of one CTA per problem (which left the pass latency-bound at 1.4 TB/s):
//   colmax_kernel      grid (row chunks, B)  column inf-norms of Q (atomic max into the D buffer)
//   scale_vec_kernel   grid B                D, beta, the scaled vectors, equality rows, flags, state reset
//   scale_pack_kernel  grid (tile chunks, B) Q~ = D Q D in the packed layout (+ the block-lower tiles of the
//                                            tensor-core factorisation), per-CTA partial sums of ||Q~||_F^2
// select_rho_kernel (factor.cu) adds the partial sums in a fixed order: rho is deterministic.
//
// Restates the setup of the reference forward solver, lqp_py/solve_box_qp_admm_torch.py:
//   :127      p_norm = ||p||_inf on the unscaled p
//   :129-130  any_lb / any_ub (OR-reduced over the batch into the control block)
//   :161-194  D from the column inf-norms of Q, beta from the 10% / 90% quantiles of D,
//             Q~ = D Q D, p~ = D p, A~ = E (A D), b~ = E b, lb~ = lb / D, ub~ = ub / D
//   :200-203  rho candidate = clamp(||Q~||_F / sqrt(n), rho_min, rho_max)
//   :221-223  x = z = u = 0
// HBM traffic: Q is read 1.5 times (column norms over the full matrix, then the lower triangle for the
// scaling), the packed lower triangle of Q~ written once (and once more as block-lower tiles when the
// tensor-core factorisation is used).
#include "layout.cuh"

namespace lqpb {

constexpr int kScaleThreads = 512;

template <typename T>
__device__ __forceinline__ void smem_atomic_max_nonneg(T* addr, T v);
template <>
__device__ __forceinline__ void smem_atomic_max_nonneg<float>(float* addr, float v) {
  atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
}
template <>
__device__ __forceinline__ void smem_atomic_max_nonneg<double>(double* addr, double v) {
  atomicMax(reinterpret_cast<long long*>(addr), __double_as_longlong(v));
}

// torch's lerp (aten/src/ATen/native/Lerp.h), used by torch.quantile's linear interpolation
template <typename T>
__device__ __forceinline__ T torch_lerp(T a, T b, T w) {
  return w < T(0.5) ? a + w * (b - a) : b - (b - a) * (T(1) - w);
}

// ---------------------------------------------------------------------------------------------
// Column inf-norms (:163).  CTA (rc, b) scans rows [rc * rows, (rc + 1) * rows) of problem b, thread = 16-byte
// column chunk; the per-thread maxima are merged with an integer atomic max (values are non-negative, so the
// result does not depend on the order).  The D buffer (zeroed by the host) receives the norms.
constexpr int kColmaxThreads = 128;
constexpr int kColmaxRows = 32;

template <typename T, int VW>
__device__ __forceinline__ void colmax_rows(const T* __restrict__ Qb, int n, int r0, int r1, T* __restrict__ out, int tid) {
  const int chunks = n / VW;
  for (int c = tid; c < chunks; c += kColmaxThreads) {
    T mx[VW];
#pragma unroll
    for (int e = 0; e < VW; ++e) mx[e] = T(0);
#pragma unroll 8
    for (int i = r0; i < r1; ++i) {
      alignas(16) T v[VW];
      if (VW == 1) v[0] = Qb[(size_t)i * n + c];
      else *reinterpret_cast<typename Vec<T>::type*>(v) =
               *reinterpret_cast<const typename Vec<T>::type*>(Qb + (size_t)i * n + c * VW);
#pragma unroll
      for (int e = 0; e < VW; ++e) mx[e] = t_max(mx[e], t_abs(v[e]));
    }
#pragma unroll
    for (int e = 0; e < VW; ++e) atomic_max_nonneg(&out[c * VW + e], mx[e]);
  }
}

template <typename T>
__global__ void __launch_bounds__(kColmaxThreads) colmax_kernel(FwdWs<T> w, const T* __restrict__ Q) {
  const int b = blockIdx.y, n = w.n;
  const int r0 = blockIdx.x * kColmaxRows, r1 = min(r0 + kColmaxRows, n);
  const T* Qb = Q + (size_t)b * n * n;
  T* out = w.D + (size_t)b * w.ld;
  const bool vec_ok = (n % Vec<T>::N) == 0 && ((uintptr_t)Q % 16) == 0;
  if (vec_ok) colmax_rows<T, Vec<T>::N>(Qb, n, r0, r1, out, threadIdx.x);
  else colmax_rows<T, 1>(Qb, n, r0, r1, out, threadIdx.x);
}

// ---------------------------------------------------------------------------------------------
// Per-problem vector work: D from the column norms, beta, blend; p~, lb~, ub~, flags, p_norm, state reset;
// equality rows A~ = E (A D), b~ = E b.
template <typename T>
__global__ void __launch_bounds__(kScaleThreads)
scale_vec_kernel(lqpb_config cfg, FwdWs<T> w, const T* __restrict__ p, const T* __restrict__ A,
                 const T* __restrict__ bvec, const T* __restrict__ lb, const T* __restrict__ ub, int P2, int SB) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* Ds = reinterpret_cast<T*>(smem_raw);           // [ld]   column norms, then D
  T* sortbuf = Ds + w.ld;                           // [SB = max(P2, m)]: sort buffer of P2 entries, later the m row norms
  T* scratch = sortbuf + SB;                        // [32]
  double* dscratch = reinterpret_cast<double*>(scratch + 32);  // [32]
  __shared__ T s_beta, s_mean;

  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const int n = w.n, m = w.m, ld = w.ld;
  const size_t vo = (size_t)b * ld;

  if (cfg.scale) {
    for (int j = tid; j < ld; j += kScaleThreads) Ds[j] = w.D[vo + j];     // column inf-norms (colmax_kernel)
    __syncthreads();
    // mean of the norms (:166), zero guard (:164-168), D = sqrt(1/norm) (:170)
    double part = 0.0;
    for (int j = tid; j < n; j += kScaleThreads) part += (double)Ds[j];
    const double tot = group_sum(part, dscratch, tid, kScaleThreads, 0);
    const T floor_v = t_max((T)(tot / n), T(1e-6));
    for (int j = tid; j < n; j += kScaleThreads) {
      T q = Ds[j];
      if (q <= T(0)) q = t_max(q, floor_v);
      Ds[j] = t_sqrt(T(1) / q);
    }
    __syncthreads();
    // beta = 1 - q10(D)/q90(D) (:171-174): bitonic sort + torch.quantile's linear interpolation
    if (cfg.beta_auto) {
      for (int j = tid; j < P2; j += kScaleThreads) sortbuf[j] = j < n ? Ds[j] : t_inf<T>();
      __syncthreads();
      for (int k = 2; k <= P2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
          for (int i = tid; i < P2; i += kScaleThreads) {
            const int ixj = i ^ j;
            if (ixj > i) {
              const T a = sortbuf[i], c = sortbuf[ixj];
              const bool asc = (i & k) == 0;
              if ((a > c) == asc) { sortbuf[i] = c; sortbuf[ixj] = a; }
            }
          }
          __syncthreads();
        }
      }
      if (tid == 0) {
        T v[2];
        const T qs[2] = {T(0.10), T(0.90)};
        for (int k = 0; k < 2; ++k) {
          const T rank = qs[k] * T(n - 1);
          const T lo = floor(rank), hi = ceil(rank);
          v[k] = torch_lerp(sortbuf[(int)lo], sortbuf[(int)hi], rank - lo);
        }
        s_beta = T(1) - v[0] / v[1];
      }
    } else if (tid == 0) {
      s_beta = (T)cfg.beta;
    }
    // D <- (1-beta) D + beta mean(D) (:175)
    part = 0.0;
    for (int j = tid; j < n; j += kScaleThreads) part += (double)Ds[j];
    const double dsum = group_sum(part, dscratch, tid, kScaleThreads, 0);
    if (tid == 0) s_mean = (T)(dsum / n);
    __syncthreads();
    const T beta = s_beta, mean = s_mean;
    for (int j = tid; j < ld; j += kScaleThreads) Ds[j] = j < n ? (T(1) - beta) * Ds[j] + beta * mean : T(1);
    __syncthreads();
  } else {
    for (int j = tid; j < ld; j += kScaleThreads) Ds[j] = T(1);
    __syncthreads();
  }

  // ---- vectors: p~, lb~, ub~, state reset, flags, p_norm
  T pmax = T(0);
  int f_lb = 0, f_ub = 0;
  for (int j = tid; j < ld; j += kScaleThreads) {
    T d = Ds[j], pv = T(0), l = T(0), uu = T(0);
    if (j < n) {
      pv = p[(size_t)b * n + j];
      l = lb[(size_t)b * n + j];
      uu = ub[(size_t)b * n + j];
      pmax = t_max(pmax, t_abs(pv));
      f_lb |= (l > -t_inf<T>());
      f_ub |= (uu < t_inf<T>());
      if (cfg.scale) { pv = d * pv; l = l / d; uu = uu / d; }
    }
    w.D[vo + j] = d;
    w.pt[vo + j] = pv;
    w.lbt[vo + j] = l;
    w.ubt[vo + j] = uu;
    // ADMM state: zero (reference :221-223), or a warm start from the caller's unscaled z, u of an earlier solve
    // (z = D z~, u = u~ / D, :316-318)
    w.z[vo + j] = (w.z0 && j < n) ? w.z0[(size_t)b * n + j] / d : T(0);
    w.u[vo + j] = (w.u0 && j < n) ? w.u0[(size_t)b * n + j] * d : T(0);
    w.c[vo + j] = T(0);
    w.xs[vo + j] = T(0);
  }
  pmax = group_max(pmax, scratch, tid, kScaleThreads, 0);
  f_lb = __syncthreads_or(f_lb);
  f_ub = __syncthreads_or(f_ub);
  if (tid == 0) {
    w.pnorm[b] = pmax;
    if (f_lb) atomicOr(&w.ctrl->any_lb, 1);
    if (f_ub) atomicOr(&w.ctrl->any_ub, 1);
    w.ratio[b] = T(1);
    w.wants[b] = 0;
    w.chk[4 * b + 0] = w.chk[4 * b + 1] = w.chk[4 * b + 2] = w.chk[4 * b + 3] = T(0);
  }

  // ---- equality rows: A~ = E (A D), b~ = E b (:179-190)
  if (m > 0) {
    T* rown = sortbuf;  // reuse: [m] row norms (m <= SB)
    for (int l = 0; l < m; ++l) {
      const T* Al = A + ((size_t)b * m + l) * n;
      T mx = T(0);
      for (int j = tid; j < n; j += kScaleThreads) mx = t_max(mx, t_abs(Al[j] * Ds[j]));
      mx = group_max(mx, scratch, tid, kScaleThreads, 0);
      if (tid == 0) rown[l] = mx;
    }
    __syncthreads();
    if (cfg.scale) {
      if (tid == 0) {
        double s = 0.0;
        for (int l = 0; l < m; ++l) s += (double)rown[l];
        const T fl = t_max((T)(s / m), T(1e-6));
        for (int l = 0; l < m; ++l) {
          T r = rown[l];
          if (r <= T(0)) r = t_max(r, fl);
          rown[l] = T(1) / r;  // E
        }
      }
    } else {
      if (tid < m) rown[tid] = T(1);
      for (int l = tid + kScaleThreads; l < m; l += kScaleThreads) rown[l] = T(1);
    }
    __syncthreads();
    for (int l = 0; l < m; ++l) {
      const T* Al = A + ((size_t)b * m + l) * n;
      T* Atl = w.At + ((size_t)b * m + l) * ld;
      const T e = rown[l];
      for (int j = tid; j < ld; j += kScaleThreads) {
        T v = T(0);
        if (j < n) v = cfg.scale ? e * (Al[j] * Ds[j]) : Al[j];
        Atl[j] = v;
      }
    }
    for (int l = tid; l < m; l += kScaleThreads) {
      w.E[(size_t)b * m + l] = rown[l];
      w.bt[(size_t)b * m + l] = cfg.scale ? rown[l] * bvec[(size_t)b * m + l] : bvec[(size_t)b * m + l];
    }
  }
}


// ---------------------------------------------------------------------------------------------
// Q~ = (D_i Q_ij) D_j for the lower triangle (:176), written in the packed symmetric layout (Pack<T>: diagonal
// halved, zero fill above the diagonal and in the padding) and, for the tensor-core factorisation, as
// block-lower 128 x 128 tiles (tcfactor.cu; entries with i, j < n only -- the equality rows, the padding and the
// diagonal shift are added by tc_fixup_kernel).  One warp per packed tile: lanes run along the tile columns, so
// the reads of Q are 128-byte row segments and every tile row is written as one full line.  Each CTA leaves its
// share of ||Q~||_F^2 (:201; off-diagonal entries counted twice) in fro_part[b][chunk].
constexpr int kPackThreads = 256;
constexpr int kPackWarps = kPackThreads / 32;

template <typename T>
__global__ void __launch_bounds__(kPackThreads)
scale_pack_kernel(lqpb_config cfg, FwdWs<T> w, const T* __restrict__ Q) {
  using P = Pack<T>;
  __shared__ double wsum[kPackWarps];
  const int b = blockIdx.y, n = w.n, tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const T* Qb = Q + (size_t)b * n * n;
  const T* Ds = w.D + (size_t)b * w.ld;
  T* Qpb = w.Qp + (size_t)b * P::elems(n);
  const bool do_scale = cfg.scale != 0;
  const int ntv = P::nt(n), ntl = P::ntiles(n);
  const int c = lane % P::TC, k = c / P::VN, e = c % P::VN;
  T* blb = nullptr;
  if (w.tc) blb = w.W + (size_t)b * ((size_t)w.nb * (w.nb + 1) / 2) * kTcBlock * kTcBlock;
  double fro = 0.0;
  const int t = blockIdx.x * kPackWarps + warp;
  if (t < ntl) {
    int Jc = 0, rem = t;
    while (rem >= ntv - Jc / P::R) { rem -= ntv - Jc / P::R; ++Jc; }
    const int I = Jc / P::R + rem;
    T* tp = Qpb + (size_t)t * P::TILE;
    const int j = Jc * P::TC + c;
    const T dj = j < n ? Ds[j] : T(0);
#pragma unroll 16
    for (int l0 = 0; l0 < kPackRows; l0 += P::R) {
      const int l = l0 + lane / P::TC, i = I * kPackRows + l;
      T v = T(0);
      if (i < n && j <= i) {
        v = Qb[(size_t)i * n + j];
        if (do_scale) v = (Ds[i] * v) * dj;
        fro += (i == j ? 1.0 : 2.0) * (double)v * (double)v;
        if (blb) {
          const int Ib = i >> 7, Jb = j >> 7;
          blb[((size_t)(Ib * (Ib + 1) / 2 + Jb) * kTcBlock + (i & 127)) * kTcBlock + (j & 127)] = v;
        }
        if (i == j) v *= T(0.5);
      }
      tp[P::in_tile(l, k, e)] = v;
    }
  }
  fro = warp_sum(fro);
  if (lane == 0) wsum[warp] = fro;
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int q = 0; q < kPackWarps; ++q) s += wsum[q];
    w.fro_part[(size_t)b * w.n_fro + blockIdx.x] = s;
  }
}

// any_lb / any_ub (:129-130) over the WHOLE batch, for the chunked host-buffer forward: the per-chunk
// scale_vec_kernel only sees its own problems, but rho = 0 (:157-158) is a decision about all of them.
template <typename T>
__global__ void bound_flags_kernel(FwdWs<T> w, const T* __restrict__ lb, const T* __restrict__ ub) {
  const size_t total = (size_t)w.B * w.n;
  int f_lb = 0, f_ub = 0;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    f_lb |= lb[e] > -t_inf<T>();
    f_ub |= ub[e] < t_inf<T>();
  }
  f_lb = __syncthreads_or(f_lb);
  f_ub = __syncthreads_or(f_ub);
  if (threadIdx.x == 0) {
    if (f_lb) atomicOr(&w.ctrl->any_lb, 1);
    if (f_ub) atomicOr(&w.ctrl->any_ub, 1);
  }
}

template <typename T>
cudaError_t launch_bound_flags(const FwdWs<T>& w, const T* lb, const T* ub, cudaStream_t st) {
  const size_t total = (size_t)w.B * w.n;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 296) blocks = 296;
  bound_flags_kernel<T><<<blocks, 256, 0, st>>>(w, lb, ub);
  return cudaGetLastError();
}

template <typename T>
cudaError_t launch_scale(const lqpb_config& cfg, const FwdWs<T>& w, const T* Q, const T* p, const T* A, const T* b,
                         const T* lb, const T* ub, cudaStream_t st) {
  cudaError_t e;
  if (cfg.scale) {
    e = cudaMemsetAsync(w.D, 0, (size_t)w.B * w.ld * sizeof(T), st);
    if (e != cudaSuccess) return e;
    dim3 g((w.n + kColmaxRows - 1) / kColmaxRows, w.B);
    colmax_kernel<T><<<g, kColmaxThreads, 0, st>>>(w, Q);
  }
  int P2 = 64;
  while (P2 < w.n) P2 <<= 1;
  const int SB = P2 > w.m ? P2 : round_up(w.m, 4);
  const size_t smem = (size_t)(w.ld + SB + 32) * sizeof(T) + 32 * sizeof(double) + 16;
  e = cudaFuncSetAttribute(scale_vec_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  scale_vec_kernel<T><<<w.B, kScaleThreads, smem, st>>>(cfg, w, p, A, b, lb, ub, P2, SB);
  dim3 gp(w.n_fro, w.B);
  scale_pack_kernel<T><<<gp, kPackThreads, 0, st>>>(cfg, w, Q);
  return cudaGetLastError();
}

template cudaError_t launch_bound_flags<float>(const FwdWs<float>&, const float*, const float*, cudaStream_t);
template cudaError_t launch_bound_flags<double>(const FwdWs<double>&, const double*, const double*, cudaStream_t);
template cudaError_t launch_scale<float>(const lqpb_config&, const FwdWs<float>&, const float*, const float*,
                                         const float*, const float*, const float*, const float*, cudaStream_t);
template cudaError_t launch_scale<double>(const lqpb_config&, const FwdWs<double>&, const double*, const double*,
                                          const double*, const double*, const double*, const double*, cudaStream_t);

}  // namespace lqpb
