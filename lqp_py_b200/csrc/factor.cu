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
// vectors), and applies the rank-32 update C -= W V^T to every 64 x 64 lower macro tile on the tensor
// cores (3xTF32 mma.sync for fp32, DMMA for fp64).  After np/32 steps the buffer holds -(M^-1); the epilogue negates
// it and writes K11 in the packed symmetric layout (lower triangle, layout.cuh Pack<T>) the iteration kernel streams.
#include "layout.cuh"

namespace lqpb {

constexpr int kGroup = 256;        // threads per macro-tile group (16 x 16 threads, 4 x 4 each)

#ifdef LQPB_PHASE_TIMERS
__device__ long long g_phase_cycles[16];
#define PHASE_T0() long long t__ = clock64()
#define PHASE_ADD(k) do { __syncthreads(); if (blockIdx.x == 0 && threadIdx.x == 0) { long long n__ = clock64(); \
    g_phase_cycles[k] += n__ - t__; t__ = n__; } } while (0)
#else
#define PHASE_T0()
#define PHASE_ADD(k)
#endif

template <typename T> struct GjCfg;
template <> struct GjCfg<float>  { static constexpr int NT = 1024; };
template <> struct GjCfg<double> { static constexpr int NT = 1024; };

// ---------------------------------------------------------------------------------------------
// Rank-32 update of one 64 x 64 macro tile, C -= W V^T, by a group of 256 threads with the two 32 x 64
// panel tiles staged k-major in shared memory ("staged" path: used when the panels of a whole step do
// not fit in shared memory, i.e. fp64 and np > 512).
//   fp64: DMMA (mma.sync m8n8k4, the only tensor-core path that takes fp64 operands), 8 warps as 2 x 4,
//         each a 32 x 16 sub-tile; row stride 68 makes every fragment load bank-conflict free.
//   fp32: 4 x 4 register tiles of FFMA.  (A 3xTF32 mma.sync variant was measured in round 1: not faster --
//         the update is latency-, not FLOP-bound -- and 10x less accurate, so fp32 stays on the FP32 pipe.)
template <typename T> struct TileCfg;
template <> struct TileCfg<float>  { static constexpr int LDS = 64; static constexpr int ARRAYS = 2; };
template <> struct TileCfg<double> { static constexpr int LDS = 68; static constexpr int ARRAYS = 2; };

// per-group shared scratch: the staged panel tiles, or a 32 x 33 transpose tile in the epilogue
template <typename T> struct GroupSmem {
  static constexpr int staged = TileCfg<T>::ARRAYS * kTile * TileCfg<T>::LDS;
  static constexpr int value = staged > kTile * (kTile + 1) ? staged : kTile * (kTile + 1);
};

__device__ __forceinline__ void mma_f64(double (&d)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d[0]), "+d"(d[1])
               : "d"(a), "d"(b));
}

// float: smem = [W | V], each [32][64]
__device__ __forceinline__ void tile_update(float* sm, const float* Wg, const float* Vb, float* Wb, int np, int i_base,
                                            int j_base, int gt, int bar) {
  constexpr int L = TileCfg<float>::LDS;
  float* Wt = sm;
  float* Vt = sm + kTile * L;
  bar_sync(bar, kGroup);   // previous tile's reads are done
  for (int e = gt; e < kTile * kMacro / 4; e += kGroup) {
    const int kk = e / (kMacro / 4), cc = (e % (kMacro / 4)) * 4;
    *reinterpret_cast<float4*>(Wt + kk * L + cc) = *reinterpret_cast<const float4*>(Wg + (size_t)kk * np + i_base + cc);
    *reinterpret_cast<float4*>(Vt + kk * L + cc) = *reinterpret_cast<const float4*>(Vb + (size_t)kk * np + j_base + cc);
  }
  bar_sync(bar, kGroup);
  const int ty = gt / 16, tx = gt % 16;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
#pragma unroll 8
  for (int kk = 0; kk < kTile; ++kk) {
    const float4 wa = *reinterpret_cast<const float4*>(Wt + kk * L + ty * 4);
    const float4 vb = *reinterpret_cast<const float4*>(Vt + kk * L + tx * 4);
    const float w[4] = {wa.x, wa.y, wa.z, wa.w}, v[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[a][c] += w[a] * v[c];
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    float4* cp = reinterpret_cast<float4*>(Wb + (size_t)(i_base + ty * 4 + a) * np + j_base + tx * 4);
    float4 cv = *cp;
    cv.x -= acc[a][0]; cv.y -= acc[a][1]; cv.z -= acc[a][2]; cv.w -= acc[a][3];
    *cp = cv;
  }
}

// double: smem = [W | V], each [32][68] (W negated); DMMA m8n8k4
__device__ __forceinline__ void tile_update(double* sm, const double* Wg, const double* Vb, double* Wb, int np,
                                            int i_base, int j_base, int gt, int bar) {
  constexpr int L = TileCfg<double>::LDS;
  double* Ws = sm;
  double* Vs = sm + kTile * L;
  bar_sync(bar, kGroup);
  for (int e = gt; e < kTile * kMacro / 2; e += kGroup) {
    const int kk = e / (kMacro / 2), cc = (e % (kMacro / 2)) * 2;
    double2 w2 = *reinterpret_cast<const double2*>(Wg + (size_t)kk * np + i_base + cc);
    const double2 v2 = *reinterpret_cast<const double2*>(Vb + (size_t)kk * np + j_base + cc);
    w2.x = -w2.x; w2.y = -w2.y;
    *reinterpret_cast<double2*>(Ws + kk * L + cc) = w2;
    *reinterpret_cast<double2*>(Vs + kk * L + cc) = v2;
  }
  bar_sync(bar, kGroup);
  const int warp = gt >> 5, lane = gt & 31, gid = lane >> 2, tig = lane & 3;
  const int wi = (warp >> 2) * 32, wj = (warp & 3) * 16;
  double acc[4][2][2];
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int nj = 0; nj < 2; ++nj) {
      const double2 c = *reinterpret_cast<const double2*>(Wb + (size_t)(i_base + wi + mi * 8 + gid) * np + j_base + wj +
                                                          nj * 8 + 2 * tig);
      acc[mi][nj][0] = c.x; acc[mi][nj][1] = c.y;
    }
#pragma unroll
  for (int k4 = 0; k4 < kTile; k4 += 4) {
    double af[4], bf[2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) af[mi] = Ws[(k4 + tig) * L + wi + mi * 8 + gid];
#pragma unroll
    for (int nj = 0; nj < 2; ++nj) bf[nj] = Vs[(k4 + tig) * L + wj + nj * 8 + gid];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
      for (int nj = 0; nj < 2; ++nj) mma_f64(acc[mi][nj], af[mi], bf[nj]);
  }
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int nj = 0; nj < 2; ++nj)
      *reinterpret_cast<double2*>(Wb + (size_t)(i_base + wi + mi * 8 + gid) * np + j_base + wj + nj * 8 + 2 * tig) =
          make_double2(acc[mi][nj][0], acc[mi][nj][1]);
}

// ---------------------------------------------------------------------------------------------
// Resident-panel path: one warp updates a 16 x 32 sub-tile  C -= W V^T  from the k-major panels in shared
// memory (row stride lp): lanes form a 4 x 8 grid with a 4 x 4 register tile each; the C loads are issued
// first and stay in flight during the k loop.
// (Measured in round 1 and rejected for fp32: mma.sync TF32 with the 3xTF32 split.  With zero-initialised
//  accumulators it is as accurate as FFMA, but the legacy HMMA path on sm_100 runs at about FFMA rate, so
//  three MMAs per product made the update 13% slower.  Only tcgen05 would pay off here -- see DESIGN.md.)
template <typename T>
__device__ __forceinline__ void subtile_update(const T* Wg, const T* Vb, int lp, T* Wb, int np, int i0, int j0,
                                               int lane) {
  constexpr int VN = Vec<T>::N;
  using V4 = typename Vec<T>::type;
  i0 += (lane >> 3) * 4;
  j0 += (lane & 7) * 4;
  T cr[4][4], acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int q = 0; q < 4; q += VN)
      *reinterpret_cast<V4*>(&cr[r][q]) = *reinterpret_cast<const V4*>(Wb + (size_t)(i0 + r) * np + j0 + q);
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = T(0);
#pragma unroll 8
  for (int kk = 0; kk < kTile; ++kk) {
    T wa[4], vb[4];
#pragma unroll
    for (int q = 0; q < 4; q += VN) {
      *reinterpret_cast<V4*>(&wa[q]) = *reinterpret_cast<const V4*>(Wg + (size_t)kk * lp + i0 + q);
      *reinterpret_cast<V4*>(&vb[q]) = *reinterpret_cast<const V4*>(Vb + (size_t)kk * lp + j0 + q);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] += wa[r] * vb[c];
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
#pragma unroll
    for (int c = 0; c < 4; ++c) cr[r][c] -= acc[r][c];
#pragma unroll
    for (int q = 0; q < 4; q += VN)
      *reinterpret_cast<V4*>(Wb + (size_t)(i0 + r) * np + j0 + q) = *reinterpret_cast<V4*>(&cr[r][q]);
  }
}

// Symmetric sweep of the 32 x 32 pivot tile by 4 warps (threads t = 0..127; thread = column t % 32, rows
// t / 32 + 4 q), ping-pong between two shared tiles so that each of the 32 steps needs one 128-thread
// named barrier.  The tile starts in P0 and, after an even number of steps, ends in P0 = -(A_kk)^-1.
constexpr int kSweepBar = 8;        // named barrier id of the 4 sweep warps
template <typename T>
__device__ __forceinline__ void sweep_pivot_tile(T (*P0)[kTile + 1], T (*P1)[kTile + 1], int t) {
  const int c = t & 31, w = t >> 5;
  T(*src)[kTile + 1] = P0;
  T(*dst)[kTile + 1] = P1;
  for (int s = 0; s < kTile; ++s) {
    const T piv = T(1) / src[s][s];
    const T asc = src[s][c];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int r = w + 4 * q;
      const T ars = src[r][s], arc = src[r][c];
      T v;
      if (r == s) v = (c == s) ? -piv : asc * piv;
      else v = (c == s) ? ars * piv : arc - ars * asc * piv;
      dst[r][c] = v;
    }
    bar_sync(kSweepBar, 128);
    T(*tmp)[kTile + 1] = src; src = dst; dst = tmp;
  }
}

// LDL = false: full symmetric sweep, the buffer ends as -(M^-1) and the epilogue extracts K11 / K21 / K22.
// LDL = true : elimination restricted to the trailing sub-matrix (block LDL^T, one third of the flops),
//              forward substitution of the embedded right-hand side on the fly, back substitution at the
//              end: solves M d = [-dpi*dl_dz; 0] for the backward pass without forming the inverse.
// RES = true : the two k-major column panels of a step (2 x 32 x np elements) stay resident in shared
//              memory; the trailing update then needs no staging and no barriers: every warp walks its own
//              16 x 32 sub-tiles (4 x 4 per lane), with the C loads in flight during the k loop.
// RES = false: panels in HBM/L2 scratch (Vg/Wg), macro tiles staged per 256-thread group (tile_update).
template <typename T, int NT, bool LDL, bool RES>
__global__ void __launch_bounds__(NT) gj_inverse_kernel(GjArgs<T> a) {
  constexpr int NG = NT / kGroup;
  constexpr int NW = NT / 32;
  constexpr int VN = Vec<T>::N;
  using V4 = typename Vec<T>::type;
  __shared__ T Pbuf[2][kTile][kTile + 1];     // pivot tiles: current step / look-ahead (and sweep ping-pong)
  __shared__ int pool_next;                    // RES: dynamic warp-tile counter of the trailing update
  extern __shared__ __align__(16) unsigned char gj_smem[];
  constexpr int kGroupSmem = GroupSmem<T>::value;                           // elements of per-group scratch
  const int n = a.n, m = a.m, np = a.np;
  T* tiles = reinterpret_cast<T*>(gj_smem);                                 // staged: [NG][kGroupSmem]; RES: V | W panels
  const int lp = RES ? np + 8 : np;                                         // row stride of the k-major panels
  const int scratch_elems = (RES && 2 * kTile * lp > NG * kGroupSmem) ? 2 * kTile * lp : NG * kGroupSmem;
  T* ys = tiles + scratch_elems;                                            // [np] LDL: running rhs (forward subst.)
  T* wv = ys + np;                                                          // [np] LDL: D^-1 L^-1 rhs, then the solution

  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  T* Wb = a.W + (size_t)b * np * np;
  T* Vb = RES ? tiles : a.Vg + (size_t)b * np * kTile;                      // k-major: Vb[c * lp + i]
  T* Wg = RES ? tiles + kTile * lp : a.Wg + (size_t)b * np * kTile;
  const bool packed_src = a.lds == 0;
  const int ntv_src = Pack<T>::nt(n);
  const T* srcb = packed_src ? a.src + (size_t)b * Pack<T>::elems(n) : a.src + (size_t)b * n * a.lds;
  const T* maskb = a.mask ? a.mask + (size_t)b * a.ldm : nullptr;
  const T* Ab = (m > 0) ? a.Arows + (size_t)b * m * a.lda : nullptr;
  const T shift = (a.diag_shift ? a.diag_shift[b] : T(0)) + a.diag_const;
  const T* dvecb = a.diag_vec ? a.diag_vec + (size_t)b * a.ldm : nullptr;

  PHASE_T0();
  // ---- prologue: lower triangle of the KKT matrix [[H, A^T], [A, a_diag I]] embedded in the np x np
  //      work matrix: H masked / shifted, the m equality rows right below it (inside the padding that
  //      the tiling needs anyway), identity on the rest of the padding.  One warp per row, coalesced.
  for (int i = warp; i < np; i += NW) {
    T* wrow = Wb + (size_t)i * np;
    if (i < n) {
      const T fi = maskb ? maskb[i] : T(1);
      const T* srow = srcb + (size_t)i * a.lds;
      for (int j = lane; j <= i; j += 32) {
        const T fj = maskb ? maskb[j] : T(1);
        const bool keep = (fi != T(0)) && (fj != T(0));
        T v = T(0);
        if (keep) {
          if (packed_src) {
            v = srcb[Pack<T>::offset(i, j, ntv_src)];
            if (i == j) v += v;                 // the packed layout stores the diagonal halved
          } else {
            v = srow[j];
          }
        }
        if (i == j) v = keep ? v + shift + (dvecb ? dvecb[i] : T(0)) : T(1);
        wrow[j] = v;
      }
    } else if (i < n + m) {
      const T* arow = Ab + (size_t)(i - n) * a.lda;
      for (int j = lane; j <= i; j += 32) {
        T v = T(0);
        if (j < n) v = maskb ? arow[j] * maskb[j] : arow[j];
        else if (j == i) v = a.a_diag;
        wrow[j] = v;
      }
    } else {
      for (int j = lane; j <= i; j += 32) wrow[j] = (j == i) ? T(1) : T(0);
    }
  }
  if (LDL) {
    for (int i = tid; i < np; i += NT) {
      ys[i] = (i < n) ? -(maskb[i] * a.rhs_g[(size_t)b * n + i]) : T(0);   // [-dpi*dl_dz; 0]  (:368-375)
      wv[i] = T(0);
    }
  }
  __syncthreads();
  PHASE_ADD(0);

  const int nt = np / kTile;        // sweep steps
  const int nm = np / kMacro;       // macro tiles per side
  const int nmac = nm * (nm + 1) / 2;
  const int g = tid / kGroup, gt = tid % kGroup;

  for (int k = 0; k < nt; ++k) {
    const int k0 = k * kTile;
    // ---- A: the swept pivot tile Ps = -(A_kk)^-1.  RES: it was produced during the previous step's
    //         trailing update (look-ahead), except for k = 0.  While warps 0-3 sweep, the other warps copy
    //         the column panel V = A[:, k] k-major: rows below the pivot block are contiguous 128-byte
    //         segments, rows above are stored transposed in block row k (LDL: rows above are finished,
    //         their panel entries are zero).
    T(*Ps)[kTile + 1] = Pbuf[k & 1];
    T(*Pscratch)[kTile + 1] = Pbuf[(k & 1) ^ 1];
    const bool need_sweep = !RES || k == 0;
    if (need_sweep) {
      for (int e = tid; e < kTile * kTile; e += NT) {
        const int r = e / kTile, c = e % kTile;
        Ps[r][c] = r >= c ? Wb[(size_t)(k0 + r) * np + k0 + c] : Wb[(size_t)(k0 + c) * np + k0 + r];
      }
      __syncthreads();
    }
    if (need_sweep && tid < 128) {
      sweep_pivot_tile<T>(Ps, Pscratch, tid);
    } else {
      const int first = need_sweep ? 128 : 0;
      for (int i = tid - first; i < np; i += NT - first) {
        if (i >= k0 + kTile) {
          const V4* rowp = reinterpret_cast<const V4*>(Wb + (size_t)i * np + k0);
#pragma unroll
          for (int cv = 0; cv < kTile / VN; ++cv) {
            const V4 t4 = rowp[cv];
            const T* tp = reinterpret_cast<const T*>(&t4);
#pragma unroll
            for (int q = 0; q < VN; ++q) Vb[(size_t)(cv * VN + q) * lp + i] = tp[q];
          }
        } else if (i >= k0) {
#pragma unroll
          for (int c = 0; c < kTile; ++c) Vb[(size_t)c * lp + i] = T(0);
        } else {
#pragma unroll 8
          for (int c = 0; c < kTile; ++c) Vb[(size_t)c * lp + i] = LDL ? T(0) : Wb[(size_t)(k0 + c) * np + i];
        }
      }
    }
    __syncthreads();
    PHASE_ADD(1);
    // ---- B: W = V * inv(A_kk) = -(V * Ps), k-major panel; the swept panel is also written straight back
    //         into the matrix (the trailing update adds zero to block row / column k, so the order is safe).
    for (int idx = tid; idx < np * 4; idx += NT) {
      const int c0 = (idx / np) * 8, i = idx % np;         // consecutive threads -> consecutive i
      T acc[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q] = T(0);
#pragma unroll 8
      for (int c = 0; c < kTile; ++c) {
        const T vc = Vb[(size_t)c * lp + i];
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] -= vc * Ps[c][c0 + q];
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) Wg[(size_t)(c0 + q) * lp + i] = acc[q];
      if (i >= k0 + kTile) {                               // rows below: 32 contiguous bytes of row i
#pragma unroll
        for (int q = 0; q < 8; q += VN)
          *reinterpret_cast<V4*>(Wb + (size_t)i * np + k0 + c0 + q) = *reinterpret_cast<V4*>(&acc[q]);
      } else if (!LDL && i < k0) {                         // rows above: transposed, coalesced along i
#pragma unroll
        for (int q = 0; q < 8; ++q) Wb[(size_t)(k0 + c0 + q) * np + i] = acc[q];
      }
    }
    for (int e = tid; e < kTile * kTile; e += NT) {
      const int r = e / kTile, c = e % kTile;
      if (r >= c) Wb[(size_t)(k0 + r) * np + k0 + c] = Ps[r][c];
    }
    if (tid == 0) pool_next = 0;
    __syncthreads();
    PHASE_ADD(2);
    if (LDL) {
      // forward substitution with this block column: z_k = ys[k0..], w_k = inv(A_kk) z_k, ys[i] -= L_ik z_k
      if (tid < kTile) {
        T acc = T(0);
        for (int c = 0; c < kTile; ++c) acc -= Ps[tid][c] * ys[k0 + c];
        wv[k0 + tid] = acc;
      }
      for (int i = k0 + kTile + tid; i < np; i += NT) {
        T acc = T(0);
#pragma unroll 8
        for (int c = 0; c < kTile; ++c) acc += Wg[(size_t)c * lp + i] * ys[k0 + c];
        ys[i] -= acc;
      }
    }
    PHASE_ADD(3);
    // ---- C: rank-32 update of the lower 64 x 64 macro tiles (LDL: trailing ones only)
    const int I0 = LDL ? (k0 + kTile) / kMacro : 0;
    const int nma = LDL ? (nm - I0) * (nm - I0 + 1) / 2 : nmac;
    if (RES) {
      // Look-ahead: warps 0-3 first update the macro tile that holds the NEXT pivot tile, read it back and
      // sweep it into the other pivot buffer while the remaining warps work through the other sub-tiles;
      // sub-tiles are handed out through a shared counter, so the four warps simply join in late.
      const int k1 = k0 + kTile;
      const int t_la = (k + 1 < nt) ? ((k1 / kMacro - I0) * (k1 / kMacro - I0 + 1) / 2 + (k1 / kMacro - I0)) : -1;
      auto do_subtile = [&](int wt) {
        const int t = wt >> 3, sub = wt & 7;
        int I = (int)((sqrtf(8.f * t + 1.f) - 1.f) * 0.5f);
        while ((I + 1) * (I + 2) / 2 <= t) ++I;
        while (I * (I + 1) / 2 > t) --I;
        const int J = t - I * (I + 1) / 2 + I0;
        I += I0;
        subtile_update(Wg, Vb, lp, Wb, np, I * kMacro + (sub >> 1) * 16, J * kMacro + (sub & 1) * 32, lane);
      };
      if (t_la >= 0 && warp < 4) {
        do_subtile(t_la * 8 + warp);
        do_subtile(t_la * 8 + 4 + warp);
        bar_sync(kSweepBar, 128);
        T(*Pn)[kTile + 1] = Pbuf[(k + 1) & 1];
        for (int e = tid; e < kTile * kTile; e += 128) {
          const int r = e / kTile, c = e % kTile;
          Pn[r][c] = r >= c ? Wb[(size_t)(k1 + r) * np + k1 + c] : Wb[(size_t)(k1 + c) * np + k1 + r];
        }
        bar_sync(kSweepBar, 128);
        sweep_pivot_tile<T>(Pn, Pbuf[k & 1], tid);
      }
      while (true) {
        int wt = 0;
        if (lane == 0) wt = atomicAdd(&pool_next, 1);
        wt = __shfl_sync(0xffffffffu, wt, 0);
        if (wt >= nma * 8) break;
        if ((wt >> 3) == t_la) continue;
        do_subtile(wt);
      }
    } else {
      for (int t = g; t < nma; t += NG) {
        int I = (int)((sqrtf(8.f * t + 1.f) - 1.f) * 0.5f);
        while ((I + 1) * (I + 2) / 2 <= t) ++I;
        while (I * (I + 1) / 2 > t) --I;
        const int J = t - I * (I + 1) / 2 + I0;
        I += I0;
        tile_update(tiles + g * kGroupSmem, Wg, Vb, Wb, np, I * kMacro, J * kMacro, gt, 1 + g);
      }
    }
    __syncthreads();
    PHASE_ADD(4);
  }

  if (LDL) {
    // ---- back substitution d_k = w_k - sum_{I>k} L_Ik^T d_I, bottom up; wv becomes the solution in place
    T* red = reinterpret_cast<T*>(gj_smem);      // [NT/32][32] (the panel tiles are no longer needed)
    for (int k = nt - 1; k >= 0; --k) {
      const int k0 = k * kTile;
      T acc = T(0);
      for (int i = k0 + kTile + warp; i < np; i += NW) acc += Wb[(size_t)i * np + k0 + lane] * wv[i];
      red[warp * 32 + lane] = acc;
      __syncthreads();
      if (tid < kTile) {
        T tot = T(0);
        for (int q = 0; q < NW; ++q) tot += red[q * 32 + tid];
        wv[k0 + tid] -= tot;
      }
      __syncthreads();
    }
    for (int i = tid; i < a.ldd; i += NT) a.sol_x[(size_t)b * a.ldd + i] = i < n ? wv[i] : T(0);
    for (int l = tid; l < m; l += NT) a.sol_nu[(size_t)b * m + l] = wv[n + l];
    return;
  }
  // ---- epilogue: -(Wb) = KKT^-1 = [[K11, K21^T], [K21, K22]].  K11 goes out in the packed symmetric layout the
  //      iteration kernel streams (lower triangle only, diagonal halved: Pack<T>), one warp per 4 KB tile with the
  //      lanes along the tile columns (coalesced row segments in, full lines out); K21 (m x n, row stride ldd,
  //      zero padded) and K22 (m x m) go to their own small buffers.
  {
    using P = Pack<T>;
    T* dstb = a.dst + (size_t)b * P::elems(n);
    const int ntv = P::nt(n), ntl = P::ntiles(n);
    const int c = lane % P::TC, kc = c / P::VN, ec = c % P::VN;
    for (int t = warp; t < ntl; t += NW) {
      int Jc = 0, rem = t;
      while (rem >= ntv - Jc / P::R) { rem -= ntv - Jc / P::R; ++Jc; }
      const int I = Jc / P::R + rem;
      T* tp = dstb + (size_t)t * P::TILE;
      const int j = Jc * P::TC + c;
#pragma unroll 4
      for (int l0 = 0; l0 < kPackRows; l0 += P::R) {
        const int l = l0 + lane / P::TC, i = I * kPackRows + l;
        T v = T(0);
        if (i < n && j <= i) {
          v = -Wb[(size_t)i * np + j];
          if (i == j) v *= T(0.5);
        }
        tp[P::in_tile(l, kc, ec)] = v;
      }
    }
  }
  T* g21 = (m > 0) ? a.G21 + (size_t)b * m * a.ldd : nullptr;
  T* k22 = (m > 0) ? a.K22 + (size_t)b * m * m : nullptr;
  const int ldd = a.ldd;
  for (int r = warp; r < m; r += NW) {
    const T* wrow = Wb + (size_t)(n + r) * np;
    for (int j = lane; j < ldd; j += 32) g21[(size_t)r * ldd + j] = j < n ? -wrow[j] : T(0);
    for (int q = lane; q < m; q += 32)
      k22[(size_t)r * m + q] = q <= r ? -wrow[n + q] : -Wb[(size_t)(n + q) * np + n + r];
  }
  PHASE_ADD(6);
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

template <typename T, bool LDL, bool RES>
static cudaError_t launch_gj_res(int B, const GjArgs<T>& a, size_t smem, cudaStream_t st) {
  constexpr int NT = GjCfg<T>::NT;
  cudaError_t e = cudaFuncSetAttribute(gj_inverse_kernel<T, NT, LDL, RES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) return e;
  gj_inverse_kernel<T, NT, LDL, RES><<<B, NT, smem, st>>>(a);
  return cudaGetLastError();
}
template <bool LDL>
static cudaError_t launch_gj(int B, const GjArgs<float>& a, cudaStream_t st) {
  size_t res_elems = (size_t)2 * kTile * (a.np + 8);
  const size_t grp_elems = (size_t)(GjCfg<float>::NT / kGroup) * GroupSmem<float>::value;
  if (res_elems < grp_elems) res_elems = grp_elems;
  const size_t res_smem = (res_elems + 2 * (size_t)a.np) * sizeof(float);
  if (res_smem <= 200 * 1024) return launch_gj_res<float, LDL, true>(B, a, res_smem, st);
  const size_t smem = ((size_t)(GjCfg<float>::NT / kGroup) * GroupSmem<float>::value + 2 * (size_t)a.np) * sizeof(float);
  return launch_gj_res<float, LDL, false>(B, a, smem, st);
}
template <bool LDL>
static cudaError_t launch_gj(int B, const GjArgs<double>& a, cudaStream_t st) {
  const size_t smem = ((size_t)(GjCfg<double>::NT / kGroup) * GroupSmem<double>::value + 2 * (size_t)a.np) * sizeof(double);
  return launch_gj_res<double, LDL, false>(B, a, smem, st);
}
template <typename T>
cudaError_t launch_gj_inverse(int B, const GjArgs<T>& a, cudaStream_t st) { return launch_gj<false>(B, a, st); }
template <typename T>
cudaError_t launch_ldl_solve(int B, const GjArgs<T>& a, cudaStream_t st) { return launch_gj<true>(B, a, st); }

// ---------------------------------------------------------------------------------------------
// rho selection (:156-158, :200-203): rho = 0 if the whole batch is unbounded, the Frobenius
// candidate if control['rho'] is None, else the user's scalar.
template <typename T>
__global__ void select_rho_kernel(lqpb_config cfg, FwdWs<T> w) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= w.B) return;
  const bool boxed = w.ctrl->any_lb || w.ctrl->any_ub;
  // :200-203 rho candidate = clamp(||Q~||_F / sqrt(n)); the partial sums are added in a fixed order
  double tot = 0.0;
  for (int q = 0; q < w.n_fro; ++q) tot += w.fro_part[(size_t)b * w.n_fro + q];
  T rc = (T)sqrt(tot) / (T)sqrt((double)w.n);
  rc = t_min(t_max(rc, (T)cfg.rho_min), (T)cfg.rho_max);
  w.rho_cand[b] = rc;
  T r = cfg.rho_auto ? rc : (w.rho_in ? w.rho_in[b] : (T)cfg.rho);
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
  template cudaError_t launch_ldl_solve<T>(int, const GjArgs<T>&, cudaStream_t);               \
  template cudaError_t launch_select_rho<T>(const lqpb_config&, const FwdWs<T>&, cudaStream_t);
INST(float)
INST(double)
#undef INST

}  // namespace lqpb

#ifdef LQPB_PHASE_TIMERS
// developer aid (tools/gj_phases.py): per-phase clock64 totals of block 0
extern "C" void lqpb_debug_phase_cycles(long long* out, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, lqpb::g_phase_cycles, sizeof(long long) * 16);
  if (reset) {
    long long z[16] = {0};
    cudaMemcpyToSymbol(lqpb::g_phase_cycles, z, sizeof(z));
  }
}
#endif
