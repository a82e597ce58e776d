// K2 -- batched dense inversion of the SPD matrix H = Q~ + rho I (forward) or of the masked
// adjoint matrix (backward) by a tiled, symmetric Gauss-Jordan ("sweep") elimination, plus the
// Schur-complement step that folds the equality rows into the x-update operator.
//
// This replaces the reference's batched LU of the KKT matrix
//   torch.linalg.lu_factor(M), M = [[Q~ + rho I, A~^T], [A~, 0]]   solve_box_qp_admm_torch.py:206-215, :252-254
// and the fresh LU inside torch.linalg.solve of the backward (:393).  Instead of LU factors the
// iteration kernel streams the explicit symmetric operator
//   K11 = H^-1 - G S^-1 G^T,  G = H^-1 A~^T,  S = A~ G,   c = G S^-1 b~,   nu = S^-1 (G^T r - b~)
// which is the top-left block of M^-1 (identical iterates to ~1e-15).
//
// One CTA per problem.  The lower triangle lives in HBM/L2 (np x np, np = n padded to 64 with an
// identity block); step k inverts the 32 x 32 pivot tile in shared memory, forms the column panels
// V = A[:,k] and W = V * inv(A_kk) (stored k-major so the trailing update reads them as 16-byte
// vectors), and applies the rank-32 update C -= W V^T to every 64 x 64 lower macro tile with
// 4 x 4 register tiles.  After np/32 steps the buffer holds -(H^-1); the epilogue negates,
// mirrors and compacts it into the row stride the iteration kernel streams.
#include "layout.cuh"

namespace lqpb {

constexpr int kGroup = 256;        // threads per macro-tile group (16 x 16 threads, 4 x 4 each)

template <typename T> struct GjCfg;
template <> struct GjCfg<float>  { static constexpr int NT = 1024; };
template <> struct GjCfg<double> { static constexpr int NT = 512; };

template <typename T, int NT>
__global__ void __launch_bounds__(NT)
gj_inverse_kernel(int n, int np, const T* __restrict__ src, int lds, const T* __restrict__ diag_shift, T diag_const,
                  const T* __restrict__ mask, int ldm, T* Wall, T* Vall, T* Wgall, T* dst, int ldd) {
  constexpr int NG = NT / kGroup;
  constexpr int VN = Vec<T>::N;
  using V4 = typename Vec<T>::type;
  __shared__ T Ps[kTile][kTile + 1];
  extern __shared__ __align__(16) unsigned char gj_smem[];
  T(*Wt)[kTile][kMacro] = reinterpret_cast<T(*)[kTile][kMacro]>(gj_smem);   // [NG][32][64] k-major W panel tile
  T(*Vt)[kTile][kMacro] = Wt + NG;                                          // [NG][32][64] k-major V panel tile

  const int b = blockIdx.x, tid = threadIdx.x;
  T* Wb = Wall + (size_t)b * np * np;
  T* Vb = Vall + (size_t)b * np * kTile;    // k-major: Vb[c * np + i]
  T* Wg = Wgall + (size_t)b * np * kTile;
  const T* srcb = src + (size_t)b * n * lds;
  const T* maskb = mask ? mask + (size_t)b * ldm : nullptr;
  const T shift = (diag_shift ? diag_shift[b] : T(0)) + diag_const;

  // ---- prologue: lower triangle of the (masked, shifted, identity-padded) matrix
  for (int idx = tid; idx < np * np; idx += NT) {
    const int i = idx / np, j = idx - i * np;
    if (j > i) continue;
    T v = T(0);
    if (i < n) {   // j <= i < n
      const T fi = maskb ? maskb[i] : T(1), fj = maskb ? maskb[j] : T(1);
      const bool keep = (fi != T(0)) && (fj != T(0));
      if (keep) v = srcb[(size_t)i * lds + j];
      if (i == j) v = keep ? v + shift : T(1);
    } else if (i == j) {
      v = T(1);
    }
    Wb[idx] = v;
  }
  __syncthreads();

  const int nt = np / kTile;        // sweep steps
  const int nm = np / kMacro;       // macro tiles per side
  const int nmac = nm * (nm + 1) / 2;
  const int g = tid / kGroup, gt = tid % kGroup;
  const int ty = gt / 16, tx = gt % 16;

  for (int k = 0; k < nt; ++k) {
    const int k0 = k * kTile;
    // ---- A: pivot tile -> shared (mirrored), swept in place: Ps <- -(A_kk)^-1
    for (int e = tid; e < kTile * kTile; e += NT) {
      const int r = e / kTile, c = e % kTile;
      Ps[r][c] = r >= c ? Wb[(size_t)(k0 + r) * np + k0 + c] : Wb[(size_t)(k0 + c) * np + k0 + r];
    }
    __syncthreads();
    for (int s = 0; s < kTile; ++s) {
      T nv[(kTile * kTile + NT - 1) / NT];
      const T piv = T(1) / Ps[s][s];
      int q = 0;
      for (int e = tid; e < kTile * kTile; e += NT, ++q) {
        const int r = e / kTile, c = e % kTile;
        const T ars = Ps[r][s], asc = Ps[s][c], arc = Ps[r][c];
        T v;
        if (r == s && c == s) v = -piv;
        else if (r == s) v = asc * piv;
        else if (c == s) v = ars * piv;
        else v = arc - ars * asc * piv;
        nv[q] = v;
      }
      __syncthreads();
      q = 0;
      for (int e = tid; e < kTile * kTile; e += NT, ++q) Ps[e / kTile][e % kTile] = nv[q];
      __syncthreads();
    }
    // ---- B: column panels V (before) and W = V * inv(A_kk) = -(V * Ps), k-major in HBM/L2.
    // Pass 1 copies the panel (rows below the pivot block are contiguous, rows above are stored
    // transposed); pass 2 re-reads it (L1/L2 hits, coalesced) in 4 chunks of 8 outputs to keep registers low.
    for (int idx = tid; idx < np * (kTile / VN); idx += NT) {
      const int i = idx / (kTile / VN), cv = (idx % (kTile / VN)) * VN;   // consecutive threads -> one 128B row segment
      if (i >= k0 + kTile) {
        const V4 t4 = *reinterpret_cast<const V4*>(Wb + (size_t)i * np + k0 + cv);
        const T* tp = reinterpret_cast<const T*>(&t4);
#pragma unroll
        for (int q = 0; q < VN; ++q) Vb[(size_t)(cv + q) * np + i] = tp[q];
      } else if (i >= k0) {
#pragma unroll
        for (int q = 0; q < VN; ++q) Vb[(size_t)(cv + q) * np + i] = T(0);
      }
    }
    for (int idx = tid; idx < k0 * kTile; idx += NT) {
      const int c = idx / k0, i = idx - c * k0;            // consecutive threads -> consecutive i
      Vb[(size_t)c * np + i] = Wb[(size_t)(k0 + c) * np + i];
    }
    __syncthreads();
    for (int idx = tid; idx < np * 4; idx += NT) {
      const int c0 = (idx / np) * 8, i = idx % np;         // consecutive threads -> consecutive i
      T acc[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q] = T(0);
#pragma unroll 8
      for (int c = 0; c < kTile; ++c) {
        const T vc = Vb[(size_t)c * np + i];
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] -= vc * Ps[c][c0 + q];
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) Wg[(size_t)(c0 + q) * np + i] = acc[q];
    }
    __syncthreads();
    // ---- C: rank-32 update of every lower 64 x 64 macro tile, one group of 256 threads per tile
    for (int t = g; t < nmac; t += NG) {
      int I = (int)((sqrtf(8.f * t + 1.f) - 1.f) * 0.5f);
      while ((I + 1) * (I + 2) / 2 <= t) ++I;
      while (I * (I + 1) / 2 > t) --I;
      const int J = t - I * (I + 1) / 2;
      const int i_base = I * kMacro, j_base = J * kMacro;
      bar_sync(1 + g, kGroup);   // previous tile's reads of Wt/Vt are done
      for (int e = gt; e < kTile * kMacro / VN; e += kGroup) {
        const int kk = e / (kMacro / VN), cc = (e % (kMacro / VN)) * VN;
        *reinterpret_cast<V4*>(&Wt[g][kk][cc]) = *reinterpret_cast<const V4*>(Wg + (size_t)kk * np + i_base + cc);
        *reinterpret_cast<V4*>(&Vt[g][kk][cc]) = *reinterpret_cast<const V4*>(Vb + (size_t)kk * np + j_base + cc);
      }
      bar_sync(1 + g, kGroup);
      T acc[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][c] = T(0);
#pragma unroll 8
      for (int kk = 0; kk < kTile; ++kk) {
        T wa[4], vb[4];
#pragma unroll
        for (int q = 0; q < 4; q += VN) {
          *reinterpret_cast<V4*>(&wa[q]) = *reinterpret_cast<const V4*>(&Wt[g][kk][ty * 4 + q]);
          *reinterpret_cast<V4*>(&vb[q]) = *reinterpret_cast<const V4*>(&Vt[g][kk][tx * 4 + q]);
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[a][c] += wa[a] * vb[c];
      }
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        T* cp = Wb + (size_t)(i_base + ty * 4 + a) * np + j_base + tx * 4;
#pragma unroll
        for (int q = 0; q < 4; q += VN) {
          V4 cv = *reinterpret_cast<V4*>(cp + q);
          T* cvp = reinterpret_cast<T*>(&cv);
#pragma unroll
          for (int r = 0; r < VN; ++r) cvp[r] -= acc[a][q + r];
          *reinterpret_cast<V4*>(cp + q) = cv;
        }
      }
    }
    __syncthreads();
    // ---- D: write the swept column panel and pivot tile back
    for (int idx = tid; idx < k0 * kTile; idx += NT) {          // rows above the pivot block: stored transposed
      const int c = idx / k0, i = idx - c * k0;                   // consecutive threads -> consecutive i
      Wb[(size_t)(k0 + c) * np + i] = Wg[(size_t)c * np + i];
    }
    for (int idx = tid; idx < (np - k0 - kTile) * kTile; idx += NT) {   // rows below: consecutive threads -> consecutive c
      const int i = k0 + kTile + idx / kTile, c = idx % kTile;
      Wb[(size_t)i * np + k0 + c] = Wg[(size_t)c * np + i];
    }
    for (int e = tid; e < kTile * kTile; e += NT) {
      const int r = e / kTile, c = e % kTile;
      if (r >= c) Wb[(size_t)(k0 + r) * np + k0 + c] = Ps[r][c];
    }
    __syncthreads();
  }

  // ---- epilogue: dst = -(Wb) mirrored to a full symmetric n x n matrix with row stride ldd
  T* dstb = dst + (size_t)b * n * ldd;
  T(*ts)[kTile + 1] = Ps;
  const int ntl = nt * (nt + 1) / 2;
  for (int t = 0; t < ntl; ++t) {
    int I = (int)((sqrtf(8.f * t + 1.f) - 1.f) * 0.5f);
    while ((I + 1) * (I + 2) / 2 <= t) ++I;
    while (I * (I + 1) / 2 > t) --I;
    const int J = t - I * (I + 1) / 2;
    __syncthreads();
    for (int e = tid; e < kTile * kTile; e += NT) {
      const int r = e / kTile, c = e % kTile;
      const int gi = I * kTile + r, gj = J * kTile + c;
      T v = T(0);
      if (I != J || r >= c) v = Wb[(size_t)gi * np + gj];
      else v = Wb[(size_t)gj * np + gi];   // diagonal tile: mirror inside the tile
      ts[r][c] = -v;
    }
    __syncthreads();
    for (int e = tid; e < kTile * kTile; e += NT) {
      const int r = e / kTile, c = e % kTile;
      int gi = I * kTile + r, gj = J * kTile + c;
      if (gi < n && gj < n) dstb[(size_t)gi * ldd + gj] = ts[r][c];
      if (I != J) {
        gi = J * kTile + r; gj = I * kTile + c;   // transposed tile, coalesced along c
        if (gi < n && gj < n) dstb[(size_t)gi * ldd + gj] = ts[c][r];
      }
    }
  }
  // zero the row padding so that padded columns never contribute
  for (int idx = tid; idx < n * (ldd - n); idx += NT) {
    const int i = idx / (ldd - n), j = n + idx % (ldd - n);
    dstb[(size_t)i * ldd + j] = T(0);
  }
}

template <typename T>
cudaError_t launch_gj_inverse(int B, int n, int np, const T* src, int lds, const T* diag_shift, T diag_const,
                              const T* mask, int ldm, T* W, T* Vg, T* Wg, T* dst, int ldd, cudaStream_t st) {
  constexpr int NT = GjCfg<T>::NT;
  const size_t smem = (size_t)2 * (NT / kGroup) * kTile * kMacro * sizeof(T);
  cudaError_t e = cudaFuncSetAttribute(gj_inverse_kernel<T, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  gj_inverse_kernel<T, NT><<<B, NT, smem, st>>>(n, np, src, lds, diag_shift, diag_const, mask, ldm, W, Vg, Wg, dst, ldd);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// rho selection (:156-158, :200-203): rho = 0 if the whole batch is unbounded, the Frobenius
// candidate if control['rho'] is None, else the user's scalar.
template <typename T>
__global__ void select_rho_kernel(lqpb_config cfg, FwdWs<T> w) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= w.B) return;
  const bool boxed = w.ctrl->any_lb || w.ctrl->any_ub;
  T r = cfg.rho_auto ? w.rho_cand[b] : (T)cfg.rho;
  if (!boxed) r = T(0);
  w.rho[b] = r;
}

template <typename T>
cudaError_t launch_select_rho(const lqpb_config& cfg, const FwdWs<T>& w, cudaStream_t st) {
  select_rho_kernel<T><<<(w.B + 127) / 128, 128, 0, st>>>(cfg, w);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Schur complement of the equality rows (m >= 1): on entry K = H^-1 (symmetric, stride ld).
//   G^T = A~ H^-1  (m x n),  S = A~ G,  K <- K - G S^-1 G^T,  c = G S^-1 b~
constexpr int kSchurThreads = 512;
constexpr int kSchurChunk = 8;

template <typename T>
__global__ void __launch_bounds__(kSchurThreads) schur_kernel(FwdWs<T> w, T* Ht_all) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* Ac = reinterpret_cast<T*>(smem_raw);        // [kSchurChunk][ld] chunk of A~ rows
  T* S = Ac + kSchurChunk * w.ld;                // [m][m+1]
  T* y = S + w.m * (w.m + 1);                    // [m]
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int n = w.n, m = w.m, ld = w.ld;
  T* Kb = w.K + (size_t)b * n * ld;
  const T* At = w.At + (size_t)b * m * ld;
  T* Gt = w.Gt + (size_t)b * m * ld;
  T* Ht = Ht_all + (size_t)b * m * ld;
  T* Sinv = w.Sinv + (size_t)b * m * m;

  // G^T[l][i] = sum_j K[j][i] A~[l][j]  (K symmetric: column sweep over contiguous rows)
  for (int l0 = 0; l0 < m; l0 += kSchurChunk) {
    const int lc = min(kSchurChunk, m - l0);
    __syncthreads();
    for (int e = tid; e < lc * ld; e += kSchurThreads) Ac[e] = At[(size_t)l0 * ld + e];
    __syncthreads();
    for (int i = tid; i < ld; i += kSchurThreads) {
      T acc[kSchurChunk];
#pragma unroll
      for (int q = 0; q < kSchurChunk; ++q) acc[q] = T(0);
      if (i < n) {
#pragma unroll 4
        for (int j = 0; j < n; ++j) {
          const T kv = Kb[(size_t)j * ld + i];
#pragma unroll
          for (int q = 0; q < kSchurChunk; ++q)
            if (q < lc) acc[q] += kv * Ac[q * ld + j];
        }
      }
#pragma unroll
      for (int q = 0; q < kSchurChunk; ++q)
        if (q < lc) Gt[(size_t)(l0 + q) * ld + i] = acc[q];
    }
  }
  __syncthreads();
  // S = A~ G  (warp per entry)
  for (int e = wid; e < m * m; e += kSchurThreads / 32) {
    const int l = e / m, l2 = e % m;
    T acc = T(0);
    for (int i = lane; i < n; i += 32) acc += At[(size_t)l * ld + i] * Gt[(size_t)l2 * ld + i];
    acc = warp_sum(acc);
    if (lane == 0) S[l * (m + 1) + l2] = acc;
  }
  __syncthreads();
  // S <- -(S^-1) by the symmetric sweep (S is SPD), then Sinv = -S
  for (int s = 0; s < m; ++s) {
    const T piv = T(1) / S[s * (m + 1) + s];
    __syncthreads();
    T nv[(kMaxM * kMaxM + kSchurThreads - 1) / kSchurThreads];
    int q = 0;
    for (int e = tid; e < m * m; e += kSchurThreads, ++q) {
      const int r = e / m, c = e % m;
      const T ars = S[r * (m + 1) + s], asc = S[s * (m + 1) + c], arc = S[r * (m + 1) + c];
      T v;
      if (r == s && c == s) v = -piv;
      else if (r == s) v = asc * piv;
      else if (c == s) v = ars * piv;
      else v = arc - ars * asc * piv;
      nv[q] = v;
    }
    __syncthreads();
    q = 0;
    for (int e = tid; e < m * m; e += kSchurThreads, ++q) S[(e / m) * (m + 1) + e % m] = nv[q];
    __syncthreads();
  }
  for (int e = tid; e < m * m; e += kSchurThreads) {
    const T v = -S[(e / m) * (m + 1) + e % m];
    S[(e / m) * (m + 1) + e % m] = v;
    Sinv[e] = v;
  }
  __syncthreads();
  // y = Sinv b~ ;  c = G y ;  H^T = Sinv G^T
  if (tid < m) {
    T acc = T(0);
    for (int l = 0; l < m; ++l) acc += S[tid * (m + 1) + l] * w.bt[(size_t)b * m + l];
    y[tid] = acc;
  }
  __syncthreads();
  for (int i = tid; i < ld; i += kSchurThreads) {
    T acc = T(0);
    if (i < n)
      for (int l = 0; l < m; ++l) acc += Gt[(size_t)l * ld + i] * y[l];
    w.c[(size_t)b * ld + i] = acc;
    for (int l = 0; l < m; ++l) {
      T h = T(0);
      if (i < n)
        for (int l2 = 0; l2 < m; ++l2) h += S[l * (m + 1) + l2] * Gt[(size_t)l2 * ld + i];
      Ht[(size_t)l * ld + i] = h;
    }
  }
  __syncthreads();
  // K <- K - G H  (K_ij -= sum_l G^T[l][i] H^T[l][j])
  for (int idx = tid; idx < n * ld; idx += kSchurThreads) {
    const int i = idx / ld, j = idx - i * ld;
    if (j >= n) continue;
    T acc = T(0);
    for (int l = 0; l < m; ++l) acc += Gt[(size_t)l * ld + i] * Ht[(size_t)l * ld + j];
    Kb[idx] -= acc;
  }
}

template <typename T>
cudaError_t launch_schur(const FwdWs<T>& w, cudaStream_t st) {
  if (w.m <= 0) return cudaSuccess;
  const size_t smem = (size_t)(kSchurChunk * w.ld + w.m * (w.m + 1) + w.m + 8) * sizeof(T);
  cudaError_t e = cudaFuncSetAttribute(schur_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  // H^T scratch: the Gauss-Jordan work matrix is free at this point (np*np >= m*ld for m <= kMaxM <= np)
  schur_kernel<T><<<w.B, kSchurThreads, smem, st>>>(w, w.W);
  return cudaGetLastError();
}

#define INST(T)                                                                                                    \
  template cudaError_t launch_gj_inverse<T>(int, int, int, const T*, int, const T*, T, const T*, int, T*, T*, T*, T*, \
                                            int, cudaStream_t);                                                    \
  template cudaError_t launch_select_rho<T>(const lqpb_config&, const FwdWs<T>&, cudaStream_t);                    \
  template cudaError_t launch_schur<T>(const FwdWs<T>&, cudaStream_t);
INST(float)
INST(double)
#undef INST

}  // namespace lqpb
