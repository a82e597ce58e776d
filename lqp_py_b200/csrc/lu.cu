// lu_layer kernels -- general (indefinite) batched LU with partial pivoting and cached-factor solves.
//
// Replaces the torch calls inside lqp_py/lu_layer.py:
//   :10,:30  torch.linalg.lu_factor(A)            -> lu_factor_kernel  (LAPACK getrf layout: packed L\U,
//                                                    1-based int32 pivots, first-max tie break)
//   :33,:52  torch.linalg.lu_solve(LU, P, b)      -> lu_solve_kernel   (reuses the cached factors; :52 solves
//                                                    with -dl_dx, hence `negate_rhs`)
//   :53      torch.matmul(dx, x^T)                -> outer_kernel
// Only reached in the reference's unroll mode (solve_box_qp_admm_torch.py:216-217,264-265), where A is the
// symmetric indefinite KKT matrix, so a pivoted LU (not Cholesky) is required.  One CTA per matrix.
#include "layout.cuh"

namespace lqpb {

constexpr int kLuThreads = 512;

template <typename T>
__global__ void __launch_bounds__(kLuThreads)
lu_factor_kernel(int N, const T* __restrict__ A, T* LUall, int32_t* __restrict__ pivall) {
  __shared__ T s_val[kLuThreads / 32];
  __shared__ int s_idx[kLuThreads / 32];
  __shared__ int s_piv;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  T* LU = LUall + (size_t)b * N * N;
  const T* Ab = A + (size_t)b * N * N;
  int32_t* piv = pivall + (size_t)b * N;
  if (LU != Ab)
    for (int e = tid; e < N * N; e += kLuThreads) LU[e] = Ab[e];
  __syncthreads();
  for (int k = 0; k < N; ++k) {
    // pivot search: first index of max |LU[i][k]|, i >= k
    T best = T(-1);
    int bi = N;
    for (int i = k + tid; i < N; i += kLuThreads) {
      const T a = t_abs(LU[(size_t)i * N + k]);
      if (a > best) { best = a; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const T ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) { s_val[wid] = best; s_idx[wid] = bi; }
    __syncthreads();
    if (tid == 0) {
      T bb = s_val[0];
      int ii = s_idx[0];
      for (int q = 1; q < kLuThreads / 32; ++q)
        if (s_val[q] > bb || (s_val[q] == bb && s_idx[q] < ii)) { bb = s_val[q]; ii = s_idx[q]; }
      s_piv = ii;
      piv[k] = ii + 1;
    }
    __syncthreads();
    const int p = s_piv;
    if (p != k) {
      for (int j = tid; j < N; j += kLuThreads) {
        const T a = LU[(size_t)k * N + j];
        LU[(size_t)k * N + j] = LU[(size_t)p * N + j];
        LU[(size_t)p * N + j] = a;
      }
    }
    __syncthreads();
    const T inv = T(1) / LU[(size_t)k * N + k];
    __syncthreads();
    for (int i = k + 1 + tid; i < N; i += kLuThreads) LU[(size_t)i * N + k] *= inv;
    __syncthreads();
    // trailing rank-1 update, threads along the contiguous j
    const int rem = N - k - 1;
    for (int e = tid; e < rem * rem; e += kLuThreads) {
      const int i = k + 1 + e / rem, j = k + 1 + e % rem;
      LU[(size_t)i * N + j] -= LU[(size_t)i * N + k] * LU[(size_t)k * N + j];
    }
    __syncthreads();
  }
}

constexpr int kSolveThreads = 256;

template <typename T>
__global__ void __launch_bounds__(kSolveThreads, 2)
lu_solve_kernel(int N, int nrhs, const T* __restrict__ LUall, const int32_t* __restrict__ pivall,
                const T* __restrict__ rhs, T* __restrict__ xout, int negate) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* y = reinterpret_cast<T*>(smem_raw);   // [N] current right-hand side / solution
  __shared__ T blk[32];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  constexpr int NW = kSolveThreads / 32;
  const T* LU = LUall + (size_t)b * N * N;
  const int32_t* piv = pivall + (size_t)b * N;
  for (int c = 0; c < nrhs; ++c) {
    __syncthreads();
    for (int i = tid; i < N; i += kSolveThreads) {
      const T v = rhs[((size_t)b * N + i) * nrhs + c];
      y[i] = negate ? -v : v;
    }
    __syncthreads();
    if (tid == 0)
      for (int k = 0; k < N; ++k) {
        const int p = piv[k] - 1;
        if (p != k) { const T a = y[k]; y[k] = y[p]; y[p] = a; }
      }
    __syncthreads();
    // forward substitution, unit lower L, 32 rows at a time
    for (int k0 = 0; k0 < N; k0 += 32) {
      const int kb = min(32, N - k0);
      for (int r = wid; r < kb; r += NW) {
        const T* row = LU + (size_t)(k0 + r) * N;
        T acc = T(0);
        for (int k = lane; k < k0; k += 32) acc += row[k] * y[k];
        acc = warp_sum(acc);
        if (lane == 0) blk[r] = y[k0 + r] - acc;
      }
      __syncthreads();
      if (wid == 0) {
        T mine = lane < kb ? blk[lane] : T(0);
        for (int cc = 0; cc < kb; ++cc) {
          const T yc = __shfl_sync(0xffffffffu, mine, cc);
          if (lane > cc && lane < kb) mine -= LU[(size_t)(k0 + lane) * N + k0 + cc] * yc;
        }
        if (lane < kb) y[k0 + lane] = mine;
      }
      __syncthreads();
    }
    // back substitution with U
    for (int k1 = N; k1 > 0; k1 -= 32) {
      const int k0 = max(0, k1 - 32), kb = k1 - k0;
      for (int r = wid; r < kb; r += NW) {
        const T* row = LU + (size_t)(k0 + r) * N;
        T acc = T(0);
        for (int k = k1 + lane; k < N; k += 32) acc += row[k] * y[k];
        acc = warp_sum(acc);
        if (lane == 0) blk[r] = y[k0 + r] - acc;
      }
      __syncthreads();
      if (wid == 0) {
        T mine = lane < kb ? blk[lane] : T(0);
        for (int cc = kb - 1; cc >= 0; --cc) {
          if (lane == cc) mine = mine / LU[(size_t)(k0 + cc) * N + k0 + cc];
          const T xc = __shfl_sync(0xffffffffu, mine, cc);
          if (lane < cc) mine -= LU[(size_t)(k0 + lane) * N + k0 + cc] * xc;
        }
        if (lane < kb) y[k0 + lane] = mine;
      }
      __syncthreads();
    }
    for (int i = tid; i < N; i += kSolveThreads) xout[((size_t)b * N + i) * nrhs + c] = y[i];
  }
}

template <typename T>
__global__ void outer_kernel(int N, int M, const T* __restrict__ a, const T* __restrict__ bv, T* __restrict__ C) {
  const int b = blockIdx.y;
  const size_t total = (size_t)N * M;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(e / M), j = (int)(e % M);
    C[(size_t)b * total + e] = a[(size_t)b * N + i] * bv[(size_t)b * M + j];
  }
}

template <typename T>
cudaError_t launch_lu_factor(int B, int N, const T* A, T* LU, int32_t* piv, cudaStream_t st) {
  lu_factor_kernel<T><<<B, kLuThreads, 0, st>>>(N, A, LU, piv);
  return cudaGetLastError();
}
template <typename T>
cudaError_t launch_lu_solve(int B, int N, int nrhs, const T* LU, const int32_t* piv, const T* rhs, T* x, int negate,
                            cudaStream_t st) {
  const size_t smem = (size_t)N * sizeof(T);
  cudaError_t e = cudaFuncSetAttribute(lu_solve_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  lu_solve_kernel<T><<<B, kSolveThreads, smem, st>>>(N, nrhs, LU, piv, rhs, x, negate);
  return cudaGetLastError();
}
template <typename T>
cudaError_t launch_outer(int B, int N, int M, const T* a, const T* b, T* C, cudaStream_t st) {
  const size_t total = (size_t)N * M;
  int gx = (int)((total + 255) / 256);
  if (gx > 1024) gx = 1024;
  outer_kernel<T><<<dim3(gx, B), 256, 0, st>>>(N, M, a, b, C);
  return cudaGetLastError();
}

#define INST(T)                                                                                         \
  template cudaError_t launch_lu_factor<T>(int, int, const T*, T*, int32_t*, cudaStream_t);             \
  template cudaError_t launch_lu_solve<T>(int, int, int, const T*, const int32_t*, const T*, T*, int,   \
                                          cudaStream_t);                                                \
  template cudaError_t launch_outer<T>(int, int, int, const T*, const T*, T*, cudaStream_t);
INST(float)
INST(double)
#undef INST

}  // namespace lqpb
