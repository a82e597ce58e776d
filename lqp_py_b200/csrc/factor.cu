// K2 -- batched dense inversion of the symmetric quasi-definite KKT matrix
//     M = [[H, A^T], [A, d I]],   H = Q~ + rho I (forward, d = 0)  or  masked Q + 1e-8 I (backward, d = 1e-8)
// by a tiled, symmetric Gauss-Jordan ("sweep") elimination without pivoting: every H pivot is positive
// and the equality rows are swept last, when their pivot block has become the negative-definite
// -(A H^-1 A^T) + d I.
//
// This replaces the reference's batched LU of the same matrix
//   torch.linalg.lu_factor(M)        solve_box_qp_admm_torch.py:206-215, :252-254
// and the fresh LU inside torch.linalg.solve of the backward (:393).  Instead of LU factors the
// iteration kernel streams the explicit symmetric top-left block K11 of M^-1 = [[K11, K21^T], [K21, K22]]:
//   x = K11 rhs + K21^T b~,   nu = K21 rhs + K22 b~        (identical iterates to ~1e-15)
// The m equality rows live in the identity padding that the 64 x 64 tiling needs anyway, so they cost nothing.
//
// One CTA per problem.  The lower triangle lives in HBM/L2 (np x np, np = n + m padded to 64 with an
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
__global__ void __launch_bounds__(NT) gj_inverse_kernel(GjArgs<T> a) {
  constexpr int NG = NT / kGroup;
  constexpr int VN = Vec<T>::N;
  using V4 = typename Vec<T>::type;
  __shared__ T Ps[kTile][kTile + 1];
  extern __shared__ __align__(16) unsigned char gj_smem[];
  T(*Wt)[kTile][kMacro] = reinterpret_cast<T(*)[kTile][kMacro]>(gj_smem);   // [NG][32][64] k-major W panel tile
  T(*Vt)[kTile][kMacro] = Wt + NG;                                          // [NG][32][64] k-major V panel tile

  const int n = a.n, m = a.m, np = a.np;
  const int b = blockIdx.x, tid = threadIdx.x;
  T* Wb = a.W + (size_t)b * np * np;
  T* Vb = a.Vg + (size_t)b * np * kTile;    // k-major: Vb[c * np + i]
  T* Wg = a.Wg + (size_t)b * np * kTile;
  const T* srcb = a.src + (size_t)b * n * a.lds;
  const T* maskb = a.mask ? a.mask + (size_t)b * a.ldm : nullptr;
  const T* Ab = (m > 0) ? a.Arows + (size_t)b * m * a.lda : nullptr;
  const T shift = (a.diag_shift ? a.diag_shift[b] : T(0)) + a.diag_const;

  // ---- prologue: lower triangle of the KKT matrix [[H, A^T], [A, a_diag I]] embedded in the np x np
  //      work matrix: H masked / shifted, the m equality rows right below it (inside the padding that
  //      the tiling needs anyway), identity on the rest of the padding.
  for (int idx = tid; idx < np * np; idx += NT) {
    const int i = idx / np, j = idx - i * np;
    if (j > i) continue;
    T v = T(0);
    if (i < n) {   // j <= i < n
      const T fi = maskb ? maskb[i] : T(1), fj = maskb ? maskb[j] : T(1);
      const bool keep = (fi != T(0)) && (fj != T(0));
      if (keep) v = srcb[(size_t)i * a.lds + j];
      if (i == j) v = keep ? v + shift : T(1);
    } else if (i < n + m) {
      if (j < n) {
        v = Ab[(size_t)(i - n) * a.lda + j];
        if (maskb) v *= maskb[j];
      } else if (i == j) {
        v = a.a_diag;
      }
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

  // ---- epilogue: -(Wb) = KKT^-1 = [[K11, K21^T], [K21, K22]]; K11 is mirrored to a full symmetric n x n
  //      matrix with row stride ldd, K21 (m x n) and K22 (m x m) go to their own buffers.
  T* dstb = a.dst + (size_t)b * n * a.ldd;
  T* g21 = (m > 0) ? a.G21 + (size_t)b * m * a.ldd : nullptr;
  T* k22 = (m > 0) ? a.K22 + (size_t)b * m * m : nullptr;
  const int ldd = a.ldd;
  auto emit = [&](int gi, int gj, T val) {
    if (gi < n) {
      if (gj < n) dstb[(size_t)gi * ldd + gj] = val;
    } else if (gi < n + m) {
      if (gj < n) g21[(size_t)(gi - n) * ldd + gj] = val;
      else if (gj < n + m) k22[(size_t)(gi - n) * m + (gj - n)] = val;
    }
  };
  T(*ts)[kTile + 1] = Ps;
  const int ntl = nt * (nt + 1) / 2;
  for (int t = 0; t < ntl; ++t) {
    int I = (int)((sqrtf(8.f * t + 1.f) - 1.f) * 0.5f);
    while ((I + 1) * (I + 2) / 2 <= t) ++I;
    while (I * (I + 1) / 2 > t) --I;
    const int J = t - I * (I + 1) / 2;
    if (J * kTile >= n + m) continue;      // identity padding only
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
      emit(I * kTile + r, J * kTile + c, ts[r][c]);
      if (I != J) emit(J * kTile + r, I * kTile + c, ts[c][r]);   // transposed tile, coalesced along c
    }
  }
  // zero the row padding so that padded columns never contribute
  if (ldd > n) {
    for (int idx = tid; idx < (n + m) * (ldd - n); idx += NT) {
      const int i = idx / (ldd - n), j = n + idx % (ldd - n);
      if (i < n) dstb[(size_t)i * ldd + j] = T(0);
      else g21[(size_t)(i - n) * ldd + j] = T(0);
    }
  }
  // c = K12 b~ = K21^T b~  (constant part of the x-update)
  if (a.c_out) {
    __syncthreads();
    T* cb = a.c_out + (size_t)b * ldd;
    for (int i = tid; i < ldd; i += NT) {
      T acc = T(0);
      if (i < n)
        for (int l = 0; l < m; ++l) acc += g21[(size_t)l * ldd + i] * a.bt[(size_t)b * m + l];
      cb[i] = acc;
    }
  }
}

template <typename T>
cudaError_t launch_gj_inverse(int B, const GjArgs<T>& a, cudaStream_t st) {
  constexpr int NT = GjCfg<T>::NT;
  const size_t smem = (size_t)2 * (NT / kGroup) * kTile * kMacro * sizeof(T);
  cudaError_t e = cudaFuncSetAttribute(gj_inverse_kernel<T, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  gj_inverse_kernel<T, NT><<<B, NT, smem, st>>>(a);
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

#define INST(T)                                                                                \
  template cudaError_t launch_gj_inverse<T>(int, const GjArgs<T>&, cudaStream_t);              \
  template cudaError_t launch_select_rho<T>(const lqpb_config&, const FwdWs<T>&, cudaStream_t);
INST(float)
INST(double)
#undef INST

}  // namespace lqpb
