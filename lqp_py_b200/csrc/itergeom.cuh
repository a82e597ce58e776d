// Pieces shared by the two persistent K11-streaming kernels: the ADMM iteration kernel (iterate.cu) and the reverse
// sweep of the unrolled mode (unroll.cu) -- CTA geometry / shared-memory plan and the butterfly column reduction.
#pragma once
#include <cstdlib>
#include "layout.cuh"

namespace lqpb {

constexpr int kIterMaxWarps = 16;
constexpr int kIterMaxThreads = kIterMaxWarps * 32;
constexpr int kIterMaxDepth = 8;

struct IterGeom {
  int nwarps;       // warps per CTA (each owns a run of tiles and a private ring)
  int depth;        // ring slots (4 KB tiles) per warp
  int nt;           // block rows of the packed layout
  int nbc;          // block columns
  int ntiles;       // tiles per matrix
  int np;           // padded vector length = 32 * nt
};

// Butterfly transpose-reduce: on entry lane l holds its own partial sums acc[0..TC) for the TC columns of a
// block column; on exit acc[0] of lane l is the total (over the 32 lanes) of column (l mod TC).
template <typename T, int TC>
__device__ __forceinline__ void reduce_cols(T (&acc)[TC], int lane) {
  if (TC == 16) {
#pragma unroll
    for (int k = 0; k < TC; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 16);
  }
#pragma unroll
  for (int s = (TC == 32 ? 16 : 8); s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int k = 0; k < s; ++k) {
      const T send = up ? acc[k] : acc[k + s];
      const T keep = up ? acc[k + s] : acc[k];
      acc[k] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
}

// Register state of a warp's symmetric pass over packed tiles: the TC entries v_J of the current block column and the
// TC running column sums.  apply() consumes one tile row (8 rotated 16-byte chunks, lane = row): returns the row
// sum  T[l,:] v_J  and adds  T[l,:] v_I[l]  to the column sums.  (Packed FFMA2 was measured for fp32 and lost:
// 9.5 vs 9.3 us per ADMM iteration at dz=500 -- the pass waits on the tile stream, not on instruction issue.)
template <typename T>
struct SymAcc {
  static constexpr int TC = Pack<T>::TC, VN = Pack<T>::VN;
  using V4 = typename Vec<T>::type;
  T vJ[TC], colacc[TC];
  __device__ __forceinline__ void load_vJ(const T* src) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const V4 t4 = *reinterpret_cast<const V4*>(src + k * VN);
      const T* tp = reinterpret_cast<const T*>(&t4);
#pragma unroll
      for (int e = 0; e < VN; ++e) vJ[k * VN + e] = tp[e];
    }
#pragma unroll
    for (int c = 0; c < TC; ++c) colacc[c] = T(0);
  }
  __device__ __forceinline__ T apply(const V4 (&kv)[8], T vI) {
    T rs[4] = {T(0), T(0), T(0), T(0)};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const T* kp = reinterpret_cast<const T*>(&kv[k]);
#pragma unroll
      for (int e = 0; e < VN; ++e) {
        rs[k & 3] += kp[e] * vJ[k * VN + e];
        colacc[k * VN + e] += kp[e] * vI;
      }
    }
    return (rs[0] + rs[1]) + (rs[2] + rs[3]);
  }
  // column totals: on exit the return value of lane l is the sum over the 32 lanes of column (l mod TC)
  __device__ __forceinline__ T reduce(int lane) {
    reduce_cols<T, TC>(colacc, lane);
    return colacc[0];
  }
};

inline int env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  return (s && *s) ? atoi(s) : dflt;
}

// Shared-memory plan: `nwarps` private rings of `depth` 4 KB slots + the per-warp partial sums + vectors.
// Default: as many warps as fit with at least two slots each (measured on B200: warps matter more than ring
// depth -- 16 x 2 beats 12 x 4 and 8 x 6 at dz=500), then as many slots as fit; small problems get fewer,
// busier warps.  LQPB_ITER_WARPS / LQPB_ITER_DEPTH override the plan (tuning aid, tools/iter_tune.py).
template <typename T>
inline bool make_geom(const FwdWs<T>& w, int max_smem, IterGeom* out, size_t* smem_bytes) {
  using P = Pack<T>;
  IterGeom g{};
  g.nt = P::nt(w.n);
  g.nbc = P::nbc(w.n);
  g.ntiles = P::ntiles(w.n);
  g.np = kPackRows * g.nt;
  const size_t tile_bytes = (size_t)P::TILE * sizeof(T);
  auto fixed = [&](int nw) {
    return ((size_t)nw * g.np + 3 * (size_t)g.np + (w.m > 0 ? round_up(w.m, 4) : 4) + 6 * 16 + 4) * sizeof(T) +
           (size_t)nw * kIterMaxDepth * sizeof(uint64_t) + 128;
  };
  const int want_w = env_int("LQPB_ITER_WARPS", 0), want_d = env_int("LQPB_ITER_DEPTH", 0);
  int nw_max = kIterMaxWarps;
  if (!want_w && nw_max > round_up(g.ntiles, 4)) nw_max = round_up(g.ntiles, 4) < 4 ? 4 : round_up(g.ntiles, 4);
  for (int nw = want_w ? want_w : nw_max; nw >= 4; nw -= 2) {
    if (nw > kIterMaxWarps) continue;
    if (fixed(nw) + 2 * nw * tile_bytes > (size_t)max_smem) {
      if (want_w) return false;
      continue;
    }
    int d = (int)(((size_t)max_smem - fixed(nw)) / (nw * tile_bytes));
    if (d > kIterMaxDepth) d = kIterMaxDepth;
    if (want_d && want_d >= 2 && want_d <= d) d = want_d;
    g.nwarps = nw;
    g.depth = d;
    *smem_bytes = fixed(nw) + (size_t)nw * d * tile_bytes;
    *out = g;
    return true;
  }
  return false;
}


}  // namespace lqpb
