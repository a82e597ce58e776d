// K1 -- problem scaling, rho candidate and state reset (one CTA per problem).
//
// Restates the setup of the reference forward solver, lqp_py/solve_box_qp_admm_torch.py:
//   :127      p_norm = ||p||_inf on the unscaled p
//   :129-130  any_lb / any_ub (OR-reduced over the batch into the control block)
//   :161-194  D from the column inf-norms of Q, beta from the 10% / 90% quantiles of D,
//             Q~ = D Q D, p~ = D p, A~ = E (A D), b~ = E b, lb~ = lb / D, ub~ = ub / D
//   :200-203  rho candidate = clamp(||Q~||_F / sqrt(n), rho_min, rho_max)
//   :221-223  x = z = u = 0
// HBM traffic: Q is read 1.5 times (column norms over the full matrix, then the lower triangle for the
// scaling), the packed lower triangle of Q~ written once.
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

// Column max-abs of an n x n row-major matrix (row stride n) into smem cm[] (pre-zeroed), VW-wide loads.
template <typename T, int VW>
__device__ __forceinline__ void col_absmax(const T* __restrict__ Qb, int n, T* cm, int tid) {
  const int chunks = n / VW;
  const int tpc = chunks < kScaleThreads ? round_up(chunks, 32) : kScaleThreads;
  const int ng = kScaleThreads / tpc;
  const int g = tid / tpc, t = tid % tpc;
  if (g >= ng) return;
  for (int c = t; c < chunks; c += tpc) {
    T mx[VW];
#pragma unroll
    for (int e = 0; e < VW; ++e) mx[e] = T(0);
#pragma unroll 8
    for (int i = g; i < n; i += ng) {
      alignas(16) T v[VW];
      if (VW == 1) v[0] = Qb[(size_t)i * n + c];
      else *reinterpret_cast<typename Vec<T>::type*>(v) =
               *reinterpret_cast<const typename Vec<T>::type*>(Qb + (size_t)i * n + c * VW);
#pragma unroll
      for (int e = 0; e < VW; ++e) mx[e] = t_max(mx[e], t_abs(v[e]));
    }
#pragma unroll
    for (int e = 0; e < VW; ++e) smem_atomic_max_nonneg(&cm[c * VW + e], mx[e]);
  }
}

// Q~ = (D_i Q_ij) D_j for the lower triangle, written in the packed symmetric layout (Pack<T>: diagonal
// halved, zero fill above the diagonal and in the padding); returns this thread's share of ||Q~||_F^2
// (off-diagonal entries counted twice -- Q~ is symmetric).  One warp per tile: lanes run along the tile
// columns, so the reads of Q are 128-byte row segments and every tile row is written as one full line.
template <typename T>
__device__ __forceinline__ double scale_pack(const T* __restrict__ Qb, T* __restrict__ Qpb, int n, const T* Ds,
                                             bool do_scale, int tid) {
  using P = Pack<T>;
  const int ntv = P::nt(n), ntl = P::ntiles(n);
  const int warp = tid >> 5, lane = tid & 31, nw = kScaleThreads / 32;
  const int c = lane % P::TC, k = c / P::VN, e = c % P::VN;
  double fro = 0.0;
  for (int t = warp; t < ntl; t += nw) {
    int Jc = 0, rem = t;
    while (rem >= ntv - Jc / P::R) { rem -= ntv - Jc / P::R; ++Jc; }
    const int I = Jc / P::R + rem;
    T* tp = Qpb + (size_t)t * P::TILE;
    const int j = Jc * P::TC + c;
    const T dj = j < n ? Ds[j] : T(0);
#pragma unroll 4
    for (int l0 = 0; l0 < kPackRows; l0 += P::R) {
      const int l = l0 + lane / P::TC, i = I * kPackRows + l;
      T v = T(0);
      if (i < n && j <= i) {
        v = Qb[(size_t)i * n + j];
        if (do_scale) v = (Ds[i] * v) * dj;
        fro += (i == j ? 1.0 : 2.0) * (double)v * (double)v;
        if (i == j) v *= T(0.5);
      }
      tp[l * P::TC + ((k + l) & 7) * P::VN + e] = v;
    }
  }
  return fro;
}

template <typename T>
__global__ void __launch_bounds__(kScaleThreads)
scale_kernel(lqpb_config cfg, FwdWs<T> w, const T* __restrict__ Q, const T* __restrict__ p,
             const T* __restrict__ A, const T* __restrict__ bvec, const T* __restrict__ lb,
             const T* __restrict__ ub, int P2) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* Ds = reinterpret_cast<T*>(smem_raw);           // [ld]   column norms, then D
  T* sortbuf = Ds + w.ld;                           // [P2]
  T* scratch = sortbuf + P2;                        // [32]
  double* dscratch = reinterpret_cast<double*>(scratch + 32);  // [32]
  __shared__ T s_beta, s_mean;

  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const int n = w.n, m = w.m, ld = w.ld;
  const size_t vo = (size_t)b * ld;
  const T* Qb = Q + (size_t)b * n * n;
  T* Qpb = w.Qp + (size_t)b * Pack<T>::elems(n);

  // ---- column inf-norms of Q (:163); thread layout: groups of rows x columns
  if (cfg.scale) {
    for (int j = tid; j < ld; j += kScaleThreads) Ds[j] = T(0);
    __syncthreads();
    const bool vec_ok = (n % Vec<T>::N) == 0 && ((uintptr_t)Q % 16) == 0;
    if (vec_ok) col_absmax<T, Vec<T>::N>(Qb, n, Ds, tid);
    else col_absmax<T, 1>(Qb, n, Ds, tid);
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

  // ---- Q~ = (D_i Q_ij) D_j (:176), Frobenius norm (:201), written packed (lower triangle only)
  const double fro = scale_pack<T>(Qb, Qpb, n, Ds, cfg.scale != 0, tid);
  const double fro_tot = group_sum(fro, dscratch, tid, kScaleThreads, 0);

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
    w.z[vo + j] = T(0);
    w.u[vo + j] = T(0);
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
    T fr = (T)sqrt(fro_tot);
    T r = fr / (T)sqrt((double)n);
    r = t_min(t_max(r, (T)cfg.rho_min), (T)cfg.rho_max);
    w.rho_cand[b] = r;
    w.ratio[b] = T(1);
    w.wants[b] = 0;
    w.chk[4 * b + 0] = w.chk[4 * b + 1] = w.chk[4 * b + 2] = w.chk[4 * b + 3] = T(0);
  }

  // ---- equality rows: A~ = E (A D), b~ = E b (:179-190)
  if (m > 0) {
    T* rown = sortbuf;  // reuse: [m] row norms (m <= kMaxM <= P2)
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

template <typename T>
cudaError_t launch_scale(const lqpb_config& cfg, const FwdWs<T>& w, const T* Q, const T* p, const T* A, const T* b,
                         const T* lb, const T* ub, cudaStream_t st) {
  int P2 = 64;
  while (P2 < w.n) P2 <<= 1;
  const size_t smem = (size_t)(w.ld + P2 + 32) * sizeof(T) + 32 * sizeof(double) + 16;
  cudaError_t e = cudaFuncSetAttribute(scale_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  scale_kernel<T><<<w.B, kScaleThreads, smem, st>>>(cfg, w, Q, p, A, b, lb, ub, P2);
  return cudaGetLastError();
}

template cudaError_t launch_scale<float>(const lqpb_config&, const FwdWs<float>&, const float*, const float*,
                                         const float*, const float*, const float*, const float*, cudaStream_t);
template cudaError_t launch_scale<double>(const lqpb_config&, const FwdWs<double>&, const double*, const double*,
                                          const double*, const double*, const double*, const double*, cudaStream_t);

}  // namespace lqpb
