// K5 / K6 -- fixed-point (implicit) backward of the ADMM box-QP layer.
//
// Restates lqp_py/solve_box_qp_admm_torch.py:349-432 (torch_solve_box_qp_grad):
//   :360-365  dpi = 1 except where x + u > ub or x + u < lb            -> bwd_mask_kernel
//   :378-393  lhs = [[dpi*Q with diag += rho (1-dpi), dpi*A^T], [A, 0]] + 1e-8 I ; solve lhs d = [-dpi*dl_dz; 0]
//             The masked rows decouple (their solution is exactly 0), so the system is solved on the
//             free set F as the symmetric  [[Q_FF + 1e-8 I, A_F^T], [A_F, 1e-8 I]]  (SURVEY App. A.5):
//             factor.cu solves that masked KKT system (equality rows included) by a tiled block LDL^T
//             elimination with the right-hand side carried along (launch_ldl_solve): dv, dnu.
//   :396-427  dp = dv, dQ = 1/2 (dv x^T + x dv^T), dA = dnu x^T + nus dv^T, db = -dnu,
//             dlam = (-dl_dz - Q dv - A^T dnu) / (rho u | 1), dlb = dlam lams[:n], dub = -dlam lams[n:]
//
// KKT backward (backward='kkt', lqp_py/solve_box_qp_admm_torch.py:435-584, torch_solve_box_qp_grad_kkt):
//   :446-451  G = [-I; I], h = [-lb; ub], slacks = clamp(h - G x, 1e-8), lams = clamp(lams, 1e-8)
//   :465-493  lhs = [[Q, G^T diag(lams), A^T], [G, -diag(slacks), 0], [A, 0, 0]],  :496-507 solve lhs d = [-dl_dz; 0; 0]
//             The 2n inequality rows are eliminated in closed form -- dlam = (G dx) / slacks -- which leaves the
//             symmetric  [[Q + diag(lam_lo / s_lo + lam_hi / s_hi), A^T], [A, 0]] [dx; dnu] = [-dl_dz; 0]:
//             bwd_kkt_prep_kernel builds that diagonal (and mask = 1), the block LDL^T solver does the rest.
//   :521-584  dp = dx, dQ = 1/2 (dx x^T + x dx^T), dA = dnu x^T + nus dx^T, db = -dnu, dh = -lams dlam,
//             dlb = -dh[:n] = -lam_lo dx / s_lo,  dub = dh[n:] = -lam_hi dx / s_hi
//             (an infinite bound has slack inf and contributes exactly 0 here; the reference's dense system
//             contains -inf in that case and returns NaN for every gradient)
#include "layout.cuh"

namespace lqpb {

template <typename T>
__global__ void bwd_mask_kernel(BwdWs<T> w, const T* __restrict__ x, const T* __restrict__ u,
                                const T* __restrict__ lb, const T* __restrict__ ub) {
  const int b = blockIdx.y;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= w.ld) return;
  T f = T(0);
  if (e < w.n) {
    const size_t o = (size_t)b * w.n + e;
    const T s = x[o] + u[o];
    f = (s > ub[o] || s < lb[o]) ? T(0) : T(1);
  }
  w.mask[(size_t)b * w.ld + e] = f;
}

template <typename T>
cudaError_t launch_bwd_mask(const BwdWs<T>& w, const T* x, const T* u, const T* lb, const T* ub, cudaStream_t st) {
  dim3 grid((w.ld + 127) / 128, w.B);
  bwd_mask_kernel<T><<<grid, 128, 0, st>>>(w, x, u, lb, ub);
  return cudaGetLastError();
}

template <typename T>
__global__ void bwd_kkt_prep_kernel(BwdWs<T> w, const T* __restrict__ x, const T* __restrict__ lams,
                                    const T* __restrict__ lb, const T* __restrict__ ub) {
  const int b = blockIdx.y;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  int f_lb = 0, f_ub = 0;
  if (e < w.ld) {
    T f = T(0), d = T(0);
    if (e < w.n) {
      const size_t o = (size_t)b * w.n + e;
      const T l = lb[o], u = ub[o], xv = x[o];
      f_lb = l > -t_inf<T>();
      f_ub = u < t_inf<T>();
      const T s_lo = t_max(xv - l, T(1e-8)), s_hi = t_max(u - xv, T(1e-8));          // :450
      const T l_lo = t_max(lams[(size_t)b * 2 * w.n + e], T(1e-8));                  // :451
      const T l_hi = t_max(lams[(size_t)b * 2 * w.n + w.n + e], T(1e-8));
      d = l_lo / s_lo + l_hi / s_hi;
      f = T(1);
    }
    w.mask[(size_t)b * w.ld + e] = f;
    w.dvec[(size_t)b * w.ld + e] = d;
  }
  f_lb = __syncthreads_or(f_lb);
  f_ub = __syncthreads_or(f_ub);
  if (threadIdx.x == 0) {
    if (f_lb) atomicOr(&w.flags[0], 1);
    if (f_ub) atomicOr(&w.flags[1], 1);
  }
}

template <typename T>
cudaError_t launch_bwd_kkt_prep(const BwdWs<T>& w, const T* x, const T* lams, const T* lb, const T* ub,
                                cudaStream_t st) {
  // w.flags is zeroed once per call by the C ABI (abi.cu): the batch may arrive here in several chunks
  dim3 grid((w.ld + 127) / 128, w.B);
  bwd_kkt_prep_kernel<T><<<grid, 128, 0, st>>>(w, x, lams, lb, ub);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Gradient assembly.  grid = (row chunks, B); each CTA owns kGradRows rows of one problem:
// Q dv row dots (one warp per row), dlam/dlb/dub for those rows, and the dQ rows (streaming write).
constexpr int kGradThreads = 256;
constexpr int kGradRows = 32;

template <typename T>
__global__ void __launch_bounds__(kGradThreads)
bwd_grads_kernel(BwdWs<T> w, const T* __restrict__ dl_dz, const T* __restrict__ x, const T* __restrict__ u,
                 const T* __restrict__ lams, const T* __restrict__ nus, const T* __restrict__ Q,
                 const T* __restrict__ A, const T* __restrict__ rho_dev, T rho_scalar, T* __restrict__ dQ,
                 T* __restrict__ dp, T* __restrict__ dA, T* __restrict__ db, T* __restrict__ dlb,
                 T* __restrict__ dub, const T* __restrict__ lb_kkt, const T* __restrict__ ub_kkt) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = w.n, m = w.m, ld = w.ld;
  T* dvs = reinterpret_cast<T*>(smem_raw);   // [n]
  T* xsh = dvs + n;                          // [n]
  const int b = blockIdx.y, r0 = blockIdx.x * kGradRows;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const T* dv = w.dv + (size_t)b * ld;
  const T* dnu = w.dnu + (size_t)b * (m > 0 ? m : 1);
  for (int j = tid; j < n; j += kGradThreads) {
    dvs[j] = dv[j];
    xsh[j] = x[(size_t)b * n + j];
  }
  __syncthreads();
  const int r1 = min(r0 + kGradRows, n);
  const T rho = rho_dev ? rho_dev[b] : rho_scalar;

  if (lb_kkt) {
    // KKT mode (:521-584): dlb = -lam_lo dx / s_lo, dub = -lam_hi dx / s_hi with the clamps of :450-451
    for (int i = r0 + tid; i < r1; i += kGradThreads) {
      const size_t o = (size_t)b * n + i;
      const T xv = xsh[i], dx = dvs[i];
      if (dlb) {
        const T s_lo = t_max(xv - lb_kkt[o], T(1e-8));
        dlb[o] = -(t_max(lams[(size_t)b * 2 * n + i], T(1e-8)) * (dx / s_lo));
      }
      if (dub) {
        const T s_hi = t_max(ub_kkt[o] - xv, T(1e-8));
        dub[o] = -(t_max(lams[(size_t)b * 2 * n + n + i], T(1e-8)) * (dx / s_hi));
      }
    }
  } else if (dlb || dub) {
    using V4 = typename Vec<T>::type;
    constexpr int VN = Vec<T>::N;
    const bool vec_ok = (n % VN) == 0 && (reinterpret_cast<uintptr_t>(Q) & 15) == 0;   // rows start on 16-byte boundaries
    for (int i = r0 + wid; i < r1; i += kGradThreads / 32) {
      const T* Qi = Q + ((size_t)b * n + i) * n;
      T acc = T(0);
      if (vec_ok) {
        T a4[VN];
#pragma unroll
        for (int e = 0; e < VN; ++e) a4[e] = T(0);
#pragma unroll 4
        for (int j = lane * VN; j < n; j += 32 * VN) {
          const V4 q4 = __ldcs(reinterpret_cast<const V4*>(Qi + j));      // streamed once: do not keep it in L1 / L2
          const V4 d4 = *reinterpret_cast<const V4*>(dvs + j);
          const T* qp = reinterpret_cast<const T*>(&q4);
          const T* dp4 = reinterpret_cast<const T*>(&d4);
#pragma unroll
          for (int e = 0; e < VN; ++e) a4[e] += qp[e] * dp4[e];
        }
#pragma unroll
        for (int e = 0; e < VN; ++e) acc += a4[e];
      } else {
        for (int j = lane; j < n; j += 32) acc += Qi[j] * dvs[j];
      }
      acc = warp_sum(acc);
      if (lane == 0) {
        const size_t o = (size_t)b * n + i;
        T kkt = -dl_dz[o] - acc;                                   // :417
        if (m > 0) {
          T at = T(0);
          for (int l = 0; l < m; ++l) at += A[((size_t)b * m + l) * n + i] * dnu[l];
          kkt -= at;                                               // :419
        }
        T div = rho * u[o];                                        // :420-421
        if (div == T(0)) div = T(1);
        const T dlam = kkt / div;
        if (dlb) dlb[o] = dlam * lams[(size_t)b * 2 * n + i];      // :426
        if (dub) dub[o] = -dlam * lams[(size_t)b * 2 * n + n + i]; // :427
      }
    }
  }
  if (dp)
    for (int i = r0 + tid; i < r1; i += kGradThreads) dp[(size_t)b * n + i] = dvs[i];   // :400
  if (dQ) {                                                                              // :403-404
    using V4 = typename Vec<T>::type;
    constexpr int VN = Vec<T>::N;
    if ((n % VN) == 0 && (reinterpret_cast<uintptr_t>(dQ) & 15) == 0) {
      // every thread keeps its columns' x_j and dv_j / 2 in registers and writes them for all rows of the chunk:
      // 16-byte streaming stores, no shared-memory traffic in the row loop
      for (int j = tid * VN; j < n; j += kGradThreads * VN) {
        const V4 x4 = *reinterpret_cast<const V4*>(xsh + j);
        const V4 d4 = *reinterpret_cast<const V4*>(dvs + j);
        const T* xp = reinterpret_cast<const T*>(&x4);
        const T* dp4 = reinterpret_cast<const T*>(&d4);
        T xj[VN], hj[VN];
#pragma unroll
        for (int e = 0; e < VN; ++e) { xj[e] = xp[e]; hj[e] = T(0.5) * dp4[e]; }
#pragma unroll 4
        for (int i = r0; i < r1; ++i) {
          const T hdi = T(0.5) * dvs[i], xi = xsh[i];
          V4 o4;
          T* op = reinterpret_cast<T*>(&o4);
#pragma unroll
          for (int e = 0; e < VN; ++e) op[e] = hdi * xj[e] + hj[e] * xi;
          __stcs(reinterpret_cast<V4*>(dQ + ((size_t)b * n + i) * n + j), o4);
        }
      }
    } else {
      for (int i = r0; i < r1; ++i) {
        const T hdi = T(0.5) * dvs[i], xi = xsh[i];
        T* row = dQ + ((size_t)b * n + i) * n;
        for (int j = tid; j < n; j += kGradThreads) row[j] = hdi * xsh[j] + (T(0.5) * dvs[j]) * xi;
      }
    }
  }
  if (m > 0 && blockIdx.x == 0) {
    if (db)
      for (int l = tid; l < m; l += kGradThreads) db[(size_t)b * m + l] = -dnu[l];     // :411
  }
  if (m > 0 && dA) {
    // rows of dA are spread over the row-chunk CTAs: CTA c handles l = c, c + gridDim.x, ...
    for (int l = blockIdx.x; l < m; l += gridDim.x) {                                   // :412
      const T dn = dnu[l], nu = nus[(size_t)b * m + l];
      T* row = dA + ((size_t)b * m + l) * n;
      for (int j = tid; j < n; j += kGradThreads) row[j] = dn * xsh[j] + nu * dvs[j];
    }
  }
}

template <typename T>
cudaError_t launch_bwd_grads(const BwdWs<T>& w, const T* dl_dz, const T* x, const T* u, const T* lams, const T* nus,
                             const T* Q, const T* A, const T* rho_dev, double rho_scalar, T* dQ, T* dp, T* dA, T* db,
                             T* dlb, T* dub, cudaStream_t st, const T* lb_kkt, const T* ub_kkt) {
  const size_t smem = (size_t)2 * w.n * sizeof(T);
  cudaError_t e = cudaFuncSetAttribute(bwd_grads_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  dim3 grid((w.n + kGradRows - 1) / kGradRows, w.B);
  bwd_grads_kernel<T><<<grid, kGradThreads, smem, st>>>(w, dl_dz, x, u, lams, nus, Q, A, rho_dev, (T)rho_scalar, dQ,
                                                        dp, dA, db, dlb, dub, lb_kkt, ub_kkt);
  return cudaGetLastError();
}

#define INST(T)                                                                                                       \
  template cudaError_t launch_bwd_mask<T>(const BwdWs<T>&, const T*, const T*, const T*, const T*, cudaStream_t);    \
  template cudaError_t launch_bwd_kkt_prep<T>(const BwdWs<T>&, const T*, const T*, const T*, const T*, cudaStream_t); \
  template cudaError_t launch_bwd_grads<T>(const BwdWs<T>&, const T*, const T*, const T*, const T*, const T*,        \
                                           const T*, const T*, const T*, double, T*, T*, T*, T*, T*, T*, cudaStream_t, \
                                           const T*, const T*);
INST(float)
INST(double)
#undef INST

}  // namespace lqpb
