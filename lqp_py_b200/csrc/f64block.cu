// K2 in fp64 (n + m > 128): the blocked sweep of tcfactor.cu with fp64 tensor-core tile products (DMMA).
//
// Same algorithm, tile storage ("block lower", 128 x 128 tiles of doubles) and outputs as the fp32 per-phase kernels of
// tcfactor.cu (blocked symmetric Gauss-Jordan sweep of M = [[H, A^T], [A, d I]] for the forward inverse, block LDL^T + solve
// for the backward; replaces torch.linalg.lu_factor of solve_box_qp_admm_torch.py:206-215, :252-254 and the LU inside
// torch.linalg.solve, :393):
//   for k:  P = inv(M_kk)                       d_pivot_kernel   (register-tiled rank-1 sweep, FP64 pipe, one CTA / problem)
//           W_i = M_ik P           (i != k)     d_tile_kernel<PANEL>   also keeps V_i = old M_ik and stores M_ik <- W_i
//           M_ij -= W_i V_j^T      (i,j != k)   d_tile_kernel<TRAIL>
// Why: round 1's fp64 path (factor.cu, one CTA per problem, rank-32 updates) re-streams the 1 MB lower triangle of every
// problem 16 times -- 6.2 GB of DRAM traffic for 0.4 GB of data, 2.2 ms -- because a rank-32 step touches the whole matrix.
// With 128-wide blocks every tile is read and written once per block step (4 steps instead of 16 passes), and the O(n^3)
// work becomes 128 x 128 x 128 products on the fp64 tensor-core MMA (mma.sync.m16n8k16.f64: 2048 FMAs per instruction)
// fed from shared memory: row-major operand slabs with a row stride of 36 doubles, for which the A and B fragment loads of
// a warp (8 rows x 4 k, 8 bytes per lane) are the conflict-free minimum of two wavefronts.
#include <atomic>
#include "tcmma.cuh"

namespace lqpb {

struct DArgs {
  double* M;       // B * nb(nb+1)/2 tiles
  double* Wbuf;    // B * nb tiles : W_i of the current step (rows above the pivot row; below it W_i lives in M_ik)
  double* Vbuf;    // B * nb tiles : V_i = M_ik before the step
  double* Pbuf;    // B * nb tiles : P_k = inv(M_kk) (all kept: the LDL solve needs them)
  int nb, k, ldl;
};

// D (16 x 8) += A (16 x KK) B (KK x 8), fp64 tensor-core MMA.  Fragments (PTX ISA, .f64 mma.m16n8k*; gid = lane / 4,
// tig = lane % 4): a[2 q + h] = A[gid + 8 h][tig + 4 q], b[q] = B[tig + 4 q][gid], d[2 h + e] = D[gid + 8 h][2 tig + e].
__device__ __forceinline__ void dmma16x8x16(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, "
      "{%0,%1,%2,%3};"
      : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
      : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]),
        "d"(b[2]), "d"(b[3]));
}

// ------------------------------------------------------------------ tile products
constexpr int kDThreads = 512;                 // 16 warps as 4 x 4, each a 32 x 32 sub-tile of the 128 x 128 output
constexpr int kDLd = 36;                       // row stride (doubles) of a row-major operand slab of 32 k-columns
constexpr int kDLdT = 132;                     // row stride of a k-major slab (transposed source tile): 32 k-rows x 128
constexpr int kDSlab = kTB * kDLd;             // doubles per operand slab (>= 32 * kDLdT)
constexpr int kDSmem = 2 * 2 * kDSlab * 8;     // two stages of X | Y

struct DJob {
  const double* xsrc;
  const double* ysrc;
  double* vdst;
  double* wdst;
  double* mdst;
  int xtrans;
};

template <int MODE>
__device__ __forceinline__ DJob d_decode_job(const DArgs& a, int job_global) {
  const int k = a.k, nb = a.nb;
  const int span = a.ldl ? nb - 1 - k : nb - 1;
  const int jobs = MODE == 0 ? span : span * (span + 1) / 2;
  const int b = job_global / jobs, job = job_global % jobs;
  int i, j = 0;
  if (MODE == 0) {
    i = a.ldl ? k + 1 + job : (job < k ? job : job + 1);
  } else {
    int ii = 0;
    while ((ii + 1) * (ii + 2) / 2 <= job) ++ii;
    const int jj = job - ii * (ii + 1) / 2;
    if (a.ldl) { i = k + 1 + ii; j = k + 1 + jj; }
    else { i = ii < k ? ii : ii + 1; j = jj < k ? jj : jj + 1; }
  }
  const size_t ntile = (size_t)nb * (nb + 1) / 2;
  double* Mb = a.M + (size_t)b * ntile * kTBE;
  double* Wb = a.Wbuf + (size_t)b * nb * kTBE;
  double* Vb = a.Vbuf + (size_t)b * nb * kTBE;
  DJob t;
  if (MODE == 0) {
    t.xtrans = i < k;
    t.mdst = Mb + (t.xtrans ? bl_tile(k, i) : bl_tile(i, k));
    t.xsrc = t.mdst;
    t.ysrc = a.Pbuf + ((size_t)b * nb + k) * kTBE;
    t.vdst = Vb + (size_t)i * kTBE;
    t.wdst = Wb + (size_t)i * kTBE;
  } else {
    t.xtrans = 0;
    t.xsrc = i > k ? Mb + bl_tile(i, k) : Wb + (size_t)i * kTBE;     // W_i lives in M_ik below the pivot row
    t.ysrc = Vb + (size_t)j * kTBE;
    t.mdst = Mb + bl_tile(i, j);
    t.vdst = nullptr;
    t.wdst = nullptr;
  }
  return t;
}

// D = X Y^T (128 x 128 x 128) per job; MODE 0 (PANEL): X = M_ik (stored tile, transposed when i < k), Y = P_k, outputs
// V_i = X (raw copy), W_i -> Wbuf_i (i < k) and the M tile (transposed when i < k); MODE 1 (TRAIL): C -= D.
// Persistent CTAs, jobs round-robin.  K runs in four slabs of 32 columns: while the warps multiply slab s out of one
// shared-memory stage, the registers hold slab s + 1 (16-byte loads issued before the multiply), one barrier per slab.
template <int MODE>
__global__ void __launch_bounds__(kDThreads, 1) d_tile_kernel(DArgs a, int total_jobs) {
  extern __shared__ __align__(16) unsigned char d_smem[];
  double* stage0 = reinterpret_cast<double*>(d_smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gid = lane >> 2, tig = lane & 3;
  const int wi = (warp >> 2) * 32, wj = (warp & 3) * 32;      // the warp's 32 x 32 sub-tile
  // staging map: 16 lanes cover the 32 k-columns of a row (256 contiguous bytes), 32 rows per pass, 4 passes
  const int srow = tid >> 4, sk = (tid & 15) * 2;

  double2 xr[4], yr[4];
  auto gload = [&](const DJob& jb, int s) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int r = srow + 32 * t;
      if (!jb.xtrans) {
        xr[t] = __ldcg(reinterpret_cast<const double2*>(jb.xsrc + (size_t)r * kTB + 32 * s + sk));
      } else {
        // stored tile T[k][r] (X = T^T): k-row 32 s + (tid >> 6) + 8 t, 64 lanes cover its 128 columns
        xr[t] = __ldcg(reinterpret_cast<const double2*>(jb.xsrc + (size_t)(32 * s + (tid >> 6) + 8 * t) * kTB + (tid & 63) * 2));
      }
      yr[t] = __ldcg(reinterpret_cast<const double2*>(jb.ysrc + (size_t)r * kTB + 32 * s + sk));
    }
  };
  auto sstore = [&](const DJob& jb, int s, double* sx, double* sy) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int r = srow + 32 * t;
      if (!jb.xtrans) {
        *reinterpret_cast<double2*>(sx + r * kDLd + sk) = xr[t];
        if (MODE == 0) *reinterpret_cast<double2*>(jb.vdst + (size_t)r * kTB + 32 * s + sk) = xr[t];
      } else {
        const int kr = (tid >> 6) + 8 * t, c2 = (tid & 63) * 2;
        *reinterpret_cast<double2*>(sx + kr * kDLdT + c2) = xr[t];
        if (MODE == 0) {                  // V_i = X in its logical orientation: V[r][k] = T[k][r]
          jb.vdst[(size_t)c2 * kTB + 32 * s + kr] = xr[t].x;
          jb.vdst[(size_t)(c2 + 1) * kTB + 32 * s + kr] = xr[t].y;
        }
      }
      *reinterpret_cast<double2*>(sy + r * kDLd + sk) = yr[t];
    }
  };

  int job = blockIdx.x;
  DJob cur{};
  if (job < total_jobs) {
    cur = d_decode_job<MODE>(a, job);
    gload(cur, 0);
  }
#pragma unroll 1
  for (; job < total_jobs; job += gridDim.x) {
    // [16-row block][8-column block][fragment].  TRAIL starts from -C (the tile's loads travel while slab 0 is staged, and no
    // registers beyond the accumulators are needed) and stores -(acc) = C - X Y^T
    double acc[2][4][4];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
          if (MODE == 1) {
            const int r = wi + 16 * mi + 8 * h + gid, c = wj + 8 * ni + 2 * tig;
            const double2 cv = __ldcg(reinterpret_cast<const double2*>(cur.mdst + (size_t)r * kTB + c));
            acc[mi][ni][2 * h] = -cv.x;
            acc[mi][ni][2 * h + 1] = -cv.y;
          } else {
            acc[mi][ni][2 * h] = acc[mi][ni][2 * h + 1] = 0.0;
          }
        }
    DJob nxt = cur;
    const bool more = job + (int)gridDim.x < total_jobs;
    if (more) nxt = d_decode_job<MODE>(a, job + gridDim.x);
#pragma unroll 1
    for (int s = 0; s < 4; ++s) {
      double* sx = stage0 + (s & 1) * (2 * kDSlab);
      double* sy = sx + kDSlab;
      sstore(cur, s, sx, sy);
      if (s < 3) gload(cur, s + 1);
      else if (more) gload(nxt, 0);
      __syncthreads();                   // slab s is staged; everybody has finished multiplying slab s - 1 (other stage)
#pragma unroll
      for (int k16 = 0; k16 < 2; ++k16) {
        double af[2][8], bf[4][4];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
          for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int row = wi + 16 * mi + 8 * h + gid, kk = 16 * k16 + 4 * q + tig;
              af[mi][2 * q + h] = cur.xtrans ? sx[kk * kDLdT + row] : sx[row * kDLd + kk];
            }
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
#pragma unroll
          for (int q = 0; q < 4; ++q) bf[ni][q] = sy[(wj + 8 * ni + gid) * kDLd + 16 * k16 + 4 * q + tig];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
          for (int ni = 0; ni < 4; ++ni) dmma16x8x16(acc[mi][ni], af[mi], bf[ni]);
      }
    }
    // ---- epilogue: lane (gid, tig) holds D[wi + 16 mi + 8 h + gid][wj + 8 ni + 2 tig .. + 1] in acc[mi][ni][2 h .. + 1]
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
          const int r = wi + 16 * mi + 8 * h + gid, c = wj + 8 * ni + 2 * tig;
          const double2 d = make_double2(acc[mi][ni][2 * h], acc[mi][ni][2 * h + 1]);
          if (MODE == 1) {
            *reinterpret_cast<double2*>(cur.mdst + (size_t)r * kTB + c) = make_double2(-d.x, -d.y);
          } else if (!cur.xtrans) {
            *reinterpret_cast<double2*>(cur.mdst + (size_t)r * kTB + c) = d;
          } else {
            *reinterpret_cast<double2*>(cur.wdst + (size_t)r * kTB + c) = d;
            cur.mdst[(size_t)c * kTB + r] = d.x;             // M_ki = W_i^T
            cur.mdst[(size_t)(c + 1) * kTB + r] = d.y;
          }
        }
    cur = nxt;
    __syncthreads();                     // the last slab's reads are done before the next job restages stage 1 ... 0
  }
}

// ------------------------------------------------------------------ pivot block inverse
// One CTA (512 threads) per problem: symmetric sweep of the 128 x 128 pivot tile M_kk held in registers.  Thread (ta, tb) of a
// 32 x 16 grid owns rows 4 ta .. 4 ta + 3 and the eight columns tb, tb + 16, ..., tb + 112 -- interleaved, so that the 16
// threads of a row group read 16 consecutive doubles of a published pivot row (with 8 consecutive columns per thread 63 % of
// the kernel's shared-memory wavefronts were bank conflicts).  Per sweep step the pivot row and column are broadcast
// through double-buffered shared vectors, one __syncthreads per step.  The tile ends as -(M_kk)^-1: P = inv(M_kk) goes to
// Pbuf[k], -P back into the matrix.  (The rank-8 look-ahead form of the fp32 kernel needs twice the registers in fp64.)
__global__ void __launch_bounds__(512, 1) d_pivot_kernel(DArgs a) {
  __shared__ __align__(16) double rowbuf[2][kTB];
  __shared__ __align__(16) double colbuf[2][kTB];
  const int b = blockIdx.x, tid = threadIdx.x, ta = tid >> 4, tb = tid & 15;
  const int k = a.k, nb = a.nb;
  double* tile = a.M + ((size_t)b * ((size_t)nb * (nb + 1) / 2)) * kTBE + bl_tile(k, k);
  double acc[4][8];
  // the lower triangle is the reference copy: entries above the diagonal are read transposed
#pragma unroll
  for (int rr = 0; rr < 4; ++rr)
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int r = 4 * ta + rr, c = tb + 16 * q;
      acc[rr][q] = r >= c ? tile[(size_t)r * kTB + c] : tile[(size_t)c * kTB + r];
    }
  // pivot s = 16 q + 4 t4 + sr: column register q and row register sr are compile-time, the owners (tb == 4 t4 + sr,
  // ta == 4 q + t4) are found at run time
#pragma unroll
  for (int q = 0; q < 8; ++q) {
#pragma unroll 1
    for (int t4 = 0; t4 < 4; ++t4) {
#pragma unroll
      for (int sr = 0; sr < 4; ++sr) {
        const int s = 16 * q + 4 * t4 + sr, par = sr & 1;
        const int ra = 4 * q + t4, cb = 4 * t4 + sr;
        if (ta == ra) {          // publish pivot row s
#pragma unroll
          for (int qq = 0; qq < 8; ++qq) rowbuf[par][tb + 16 * qq] = acc[sr][qq];
        }
        if (tb == cb) {          // publish pivot column s
#pragma unroll
          for (int rr = 0; rr < 4; ++rr) colbuf[par][4 * ta + rr] = acc[rr][q];
        }
        __syncthreads();
        const double piv = 1.0 / rowbuf[par][s];
        double asc[8], tr[4];
#pragma unroll
        for (int qq = 0; qq < 8; ++qq) asc[qq] = rowbuf[par][tb + 16 * qq];
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) tr[rr] = colbuf[par][4 * ta + rr] * piv;
#pragma unroll
        for (int rr = 0; rr < 4; ++rr)
#pragma unroll
          for (int qq = 0; qq < 8; ++qq) acc[rr][qq] = fma(-tr[rr], asc[qq], acc[rr][qq]);
        if (tb == cb) {          // own column s: a[r][s] <- a[r][s] / piv
#pragma unroll
          for (int rr = 0; rr < 4; ++rr) acc[rr][q] = tr[rr];
        }
        if (ta == ra) {          // own row s: a[s][c] <- a[s][c] / piv, a[s][s] <- -1 / piv
#pragma unroll
          for (int qq = 0; qq < 8; ++qq) acc[sr][qq] = (tb == cb && qq == q) ? -piv : asc[qq] * piv;
        }
      }
    }
  }
  double* P = a.Pbuf + ((size_t)b * nb + k) * kTBE;
#pragma unroll
  for (int rr = 0; rr < 4; ++rr)
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const size_t o = (size_t)(4 * ta + rr) * kTB + tb + 16 * q;
      tile[o] = acc[rr][q];
      P[o] = -acc[rr][q];
    }
}

// ------------------------------------------------------------------ pivot block inverse, blocked by 8 (DMMA)
// The same sweep 8 pivots at a time (the scheme of pivot8_body, tcmma.cuh): with S the 8 pivot indices of a step and R the rest,
//     P = inv(A_SS),  U = P A_S:,  A_RR -= U_R^T A_SR,  A_SR <- U,  A_RS <- U^T,  A_SS <- -P.
// 16 work warps hold the tile as accumulator fragments of the fp64 tensor-core MMA -- warp w owns rows 8 w .. 8 w + 7 as
// sixteen 8 x 8 blocks, lane (gid, tig) the entries [8 w + gid][8 t + 2 tig .. + 1] -- so the rank-8 update is 32
// DMMA.8x8x4 per warp and step (A fragment = -U^T of the warp's rows, B fragments = the published pivot rows) instead of
// 256 DFMAs per thread, the pivot rows of a step live in ONE warp and the pivot columns in ONE block index.  A 17th warp
// inverts the next 8 x 8 diagonal block (look-ahead: D' = A_S'S' - U_S'^T A_SS', two entries per lane, shuffles) while the
// update runs.  Two barriers per step; rows of step sb + 1 and the raw diagonal block of step sb + 2 are published at the
// end of step sb into the other halves of double-buffered arrays.
constexpr int kDP8Threads = 544;
constexpr int kDPLd = 132;                 // row stride of the published rows / of U: fragment loads at the 2-wavefront minimum

__device__ __forceinline__ void d_inv8x8_warp(double& e0, double& e1, int lane) {
  const int i = lane >> 2, jq = lane & 3;
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    const double d_s0 = __shfl_sync(0xffffffffu, e0, s * 4 + jq);       // row s, my two columns
    const double d_s1 = __shfl_sync(0xffffffffu, e1, s * 4 + jq);
    const double d_isa = __shfl_sync(0xffffffffu, e0, i * 4 + (s >> 1)); // my row, column s
    const double d_isb = __shfl_sync(0xffffffffu, e1, i * 4 + (s >> 1));
    const double d_ssa = __shfl_sync(0xffffffffu, e0, s * 4 + (s >> 1));
    const double d_ssb = __shfl_sync(0xffffffffu, e1, s * 4 + (s >> 1));
    const double d_is = (s & 1) ? d_isb : d_isa;
    const double piv = 1.0 / ((s & 1) ? d_ssb : d_ssa);
    const double ci = d_is * piv;
    const int j0 = 2 * jq;
    if (i != s) {
      e0 = (j0 == s) ? ci : fma(-ci, d_s0, e0);
      e1 = (j0 + 1 == s) ? ci : fma(-ci, d_s1, e1);
    } else {
      e0 = (j0 == s) ? -piv : d_s0 * piv;
      e1 = (j0 + 1 == s) ? -piv : d_s1 * piv;
    }
  }
  e0 = -e0;
  e1 = -e1;
}

__device__ __forceinline__ void dmma8x8x4(double (&d)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d[0]), "+d"(d[1])
               : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(kDP8Threads, 1) d_pivot8_kernel(DArgs a) {
  __shared__ __align__(16) double rowbuf[2][8][kDPLd];   // rows S of step parity
  __shared__ __align__(16) double ubuf[8][kDPLd];        // U = P A_S:
  __shared__ __align__(16) double pbuf[2][8][8];         // P of step parity
  __shared__ __align__(16) double dbuf[2][8][8];         // raw diagonal block
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gid = lane >> 2, tig = lane & 3;
  const bool la = warp == 16;
  const int k = a.k, nb = a.nb;
  double* tile = a.M + ((size_t)b * ((size_t)nb * (nb + 1) / 2)) * kTBE + bl_tile(k, k);
  double* P = a.Pbuf + ((size_t)b * nb + k) * kTBE;
  double acc[16][2];
#pragma unroll
  for (int t = 0; t < 16; ++t) acc[t][0] = acc[t][1] = 0.0;       // defined on every path (see pivot8_body)
  const int r = 8 * warp + gid;                                     // the lane's row (work warps)
  if (!la) {
    // the lower triangle is the reference copy: entries above the diagonal are read transposed
#pragma unroll
    for (int t = 0; t < 16; ++t) {
      const int c = 8 * t + 2 * tig;
      if (t < warp) {
        const double2 v = *reinterpret_cast<const double2*>(tile + (size_t)r * kTB + c);
        acc[t][0] = v.x;
        acc[t][1] = v.y;
      } else {
        acc[t][0] = r >= c ? tile[(size_t)r * kTB + c] : tile[(size_t)c * kTB + r];
        acc[t][1] = r >= c + 1 ? tile[(size_t)r * kTB + c + 1] : tile[(size_t)(c + 1) * kTB + r];
      }
    }
    if (warp == 0) {
#pragma unroll
      for (int t = 0; t < 16; ++t) *reinterpret_cast<double2*>(&rowbuf[0][gid][8 * t + 2 * tig]) = make_double2(acc[t][0], acc[t][1]);
    }
#pragma unroll
    for (int t = 0; t < 2; ++t)
      if (warp == t) *reinterpret_cast<double2*>(&dbuf[t][gid][2 * tig]) = make_double2(acc[t][0], acc[t][1]);
  }
  __syncthreads();
  if (la) {
    const int i = lane >> 2, j0 = 2 * (lane & 3);
    double e0 = dbuf[0][i][j0], e1 = dbuf[0][i][j0 + 1];
    d_inv8x8_warp(e0, e1, lane);
    *reinterpret_cast<double2*>(&pbuf[0][i][j0]) = make_double2(e0, e1);
  }
  __syncthreads();

#pragma unroll 1
  for (int sb = 0; sb < kTB / 8; ++sb) {
    const int cur = sb & 1;
    const double(*Pc)[8] = pbuf[cur];
    const double(*rows)[kDPLd] = rowbuf[cur];
    // ---- B. U = P A_S: (8 x 128): thread -> column c = tid % 128, rows s = tid / 128 and tid / 128 + 4
    if (!la) {
      const int c = tid & 127, s0 = tid >> 7;
      double rv[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) rv[t] = rows[t][c];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int s = s0 + 4 * hh;
        double u = Pc[s][0] * rv[0];
#pragma unroll
        for (int t = 1; t < 8; ++t) u = fma(Pc[s][t], rv[t], u);
        ubuf[s][c] = u;
      }
    }
    __syncthreads();
    if (la) {
      // ---- C'. look-ahead: D' = A_S'S' - U_S'^T A_SS', P' = inv(D')
      if (sb + 1 < kTB / 8) {
        const int i = lane >> 2, j0 = 2 * (lane & 3), c0 = 8 * (sb + 1);
        double e0 = dbuf[cur ^ 1][i][j0], e1 = dbuf[cur ^ 1][i][j0 + 1];
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          const double ui = ubuf[s][c0 + i];
          e0 = fma(-ui, rows[s][c0 + j0], e0);
          e1 = fma(-ui, rows[s][c0 + j0 + 1], e1);
        }
        d_inv8x8_warp(e0, e1, lane);
        *reinterpret_cast<double2*>(&pbuf[cur ^ 1][i][j0]) = make_double2(e0, e1);
      }
    } else {
      // ---- C. rank-8 update:  A_rc -= sum_s U[s][r] * A_S[s][c]  as two DMMA k-steps per 8 x 8 block
#pragma unroll
      for (int k4 = 0; k4 < 2; ++k4) {
        const double af = -ubuf[4 * k4 + tig][r];                       // A[row gid][k tig] = -U[k][row]
#pragma unroll
        for (int t = 0; t < 16; ++t) dmma8x8x4(acc[t], af, rows[4 * k4 + tig][8 * t + gid]);   // B[k tig][col gid]
      }
#pragma unroll
      for (int t = 0; t < 16; ++t) {
        if (t == sb) {               // columns S of the lane's row:  A_rS <- U^T
          acc[t][0] = ubuf[2 * tig][r];
          acc[t][1] = ubuf[2 * tig + 1][r];
        }
      }
      if (warp == sb) {              // rows S:  A_Sc <- U, and A_SS <- -P
#pragma unroll
        for (int t = 0; t < 16; ++t) {
          const double2 u = *reinterpret_cast<const double2*>(&ubuf[gid][8 * t + 2 * tig]);
          const double2 p = *reinterpret_cast<const double2*>(&Pc[gid][2 * tig]);
          acc[t][0] = t == sb ? -p.x : u.x;
          acc[t][1] = t == sb ? -p.y : u.y;
        }
      }
      // ---- publish for the next steps (the other halves of the double buffers; last read one step ago)
      if (warp == sb + 1) {
#pragma unroll
        for (int t = 0; t < 16; ++t)
          *reinterpret_cast<double2*>(&rowbuf[cur ^ 1][gid][8 * t + 2 * tig]) = make_double2(acc[t][0], acc[t][1]);
      }
      if (warp == sb + 2) {
#pragma unroll
        for (int t = 0; t < 16; ++t)
          if (t == sb + 2) *reinterpret_cast<double2*>(&dbuf[cur][gid][2 * tig]) = make_double2(acc[t][0], acc[t][1]);
      }
    }
    __syncthreads();
  }
  if (la) return;
#pragma unroll
  for (int t = 0; t < 16; ++t) {
    const size_t o = (size_t)r * kTB + 8 * t + 2 * tig;
    *reinterpret_cast<double2*>(tile + o) = make_double2(acc[t][0], acc[t][1]);
    *reinterpret_cast<double2*>(P + o) = make_double2(-acc[t][0], -acc[t][1]);
  }
}

// ------------------------------------------------------------------ assemble / fix up / extract (as tcfactor.cu, in fp64)
__global__ void __launch_bounds__(256) d_assemble_kernel(GjArgs<double> a, double* __restrict__ Mout, int nb) {
  const int b = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
  int I = 0;
  while ((I + 1) * (I + 2) / 2 <= tile) ++I;
  const int J = tile - I * (I + 1) / 2;
  const int n = a.n, m = a.m;
  const bool packed_src = a.lds == 0;
  const int ntv = Pack<double>::nt(n);
  const double* srcb = packed_src ? a.src + (size_t)b * Pack<double>::elems(n) : a.src + (size_t)b * n * a.lds;
  const double* maskb = a.mask ? a.mask + (size_t)b * a.ldm : nullptr;
  const double* Ab = (m > 0) ? a.Arows + (size_t)b * m * a.lda : nullptr;
  const double shift = (a.diag_shift ? a.diag_shift[b] : 0.0) + a.diag_const;
  const double* dvecb = a.diag_vec ? a.diag_vec + (size_t)b * a.ldm : nullptr;
  double* dst = Mout + ((size_t)b * ((size_t)nb * (nb + 1) / 2)) * kTBE + bl_tile(I, J);
  for (int e = tid; e < kTBE; e += 256) {
    const int r = e >> 7, c = e & 127;
    const int i = I * kTB + r, j = J * kTB + c;
    if (j > i) continue;     // the strict upper triangle of a diagonal tile is never read
    double v = 0.0;
    if (i < n) {
      const double fi = maskb ? maskb[i] : 1.0, fj = maskb ? maskb[j] : 1.0;
      const bool keep = fi != 0.0 && fj != 0.0;
      if (keep) {
        if (packed_src) {
          v = srcb[Pack<double>::offset(i, j, ntv)];
          if (i == j) v += v;                          // the packed layout stores the diagonal halved
        } else {
          v = srcb[(size_t)i * a.lds + j];
        }
      }
      if (i == j) v = keep ? v + shift + (dvecb ? dvecb[i] : 0.0) : 1.0;
    } else if (i < n + m) {
      if (j < n) {
        const double av = Ab[(size_t)(i - n) * a.lda + j];
        v = maskb ? av * maskb[j] : av;
      } else if (j == i) {
        v = a.a_diag;
      }
    } else {
      v = (i == j) ? 1.0 : 0.0;
    }
    dst[e] = v;
  }
}

__global__ void __launch_bounds__(256) d_fixup_kernel(GjArgs<double> a, double* __restrict__ Mout, int nb) {
  const int b = blockIdx.x, tid = threadIdx.x;
  const int n = a.n, m = a.m, np = nb * kTB;
  double* Mb = Mout + ((size_t)b * ((size_t)nb * (nb + 1) / 2)) * kTBE;
  const double shift = (a.diag_shift ? a.diag_shift[b] : 0.0) + a.diag_const;
  for (int i = tid; i < n; i += 256) Mb[bl_off(i, i)] += shift;
  const double* Ab = (m > 0) ? a.Arows + (size_t)b * m * a.lda : nullptr;
  for (int e = tid; e < (np - n) * np; e += 256) {
    const int i = n + e / np, j = e % np;
    if (j > i) continue;
    double v = 0.0;
    if (i < n + m) {
      if (j < n) v = Ab[(size_t)(i - n) * a.lda + j];
      else if (j == i) v = a.a_diag;
    } else if (j == i) {
      v = 1.0;
    }
    Mb[bl_off(i, j)] = v;
  }
}

__global__ void __launch_bounds__(512) d_extract_kernel(GjArgs<double> a, const double* __restrict__ Min, int nb) {
  using P = Pack<double>;
  constexpr int NW = 512 / 32;
  const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = a.n, m = a.m;
  const double* Mb = Min + ((size_t)b * ((size_t)nb * (nb + 1) / 2)) * kTBE;
  double* dstb = a.dst + (size_t)b * P::elems(n);
  const int ntv = P::nt(n), ntl = P::ntiles(n);
  const int c = lane % P::TC, kc = c / P::VN, ec = c % P::VN;
  for (int t = blockIdx.x * NW + warp; t < ntl; t += gridDim.x * NW) {
    int Jc = 0, rem = t;
    while (rem >= ntv - Jc / P::R) { rem -= ntv - Jc / P::R; ++Jc; }
    const int I = Jc / P::R + rem;
    double* tp = dstb + (size_t)t * P::TILE;
    const int j = Jc * P::TC + c;
#pragma unroll 8
    for (int l0 = 0; l0 < kPackRows; l0 += P::R) {
      const int l = l0 + lane / P::TC, i = I * kPackRows + l;
      double v = 0.0;
      if (i < n && j <= i) {
        v = -Mb[bl_off(i, j)];
        if (i == j) v *= 0.5;
      }
      tp[P::in_tile(l, kc, ec)] = v;
    }
  }
  if (blockIdx.x != 0) return;
  double* g21 = (m > 0) ? a.G21 + (size_t)b * m * a.ldd : nullptr;
  double* k22 = (m > 0) ? a.K22 + (size_t)b * m * m : nullptr;
  const int ldd = a.ldd;
  for (int r = warp; r < m; r += NW) {
    for (int j = lane; j < ldd; j += 32) g21[(size_t)r * ldd + j] = j < n ? -Mb[bl_off(n + r, j)] : 0.0;
    for (int q = lane; q < m; q += 32)
      k22[(size_t)r * m + q] = q <= r ? -Mb[bl_off(n + r, n + q)] : -Mb[bl_off(n + q, n + r)];
  }
  if (a.c_out) {
    __syncthreads();
    double* cb = a.c_out + (size_t)b * ldd;
    for (int i = tid; i < ldd; i += 512) {
      double acc = 0.0;
      if (i < n)
        for (int l = 0; l < m; ++l) acc += g21[(size_t)l * ldd + i] * a.bt[(size_t)b * m + l];
      cb[i] = acc;
    }
  }
}

// ------------------------------------------------------------------ block LDL^T solve (backward)
// M = L D L^T with L_ik = W_i of step k (stored in tile (i,k)), D_k^-1 = P_k.  Solves M d = [-mask * dl_dz; 0]:
// forward  y_i -= L_ik y_k,  z_k = P_k y_k,  backward d_k = z_k - sum_{i>k} L_ik^T d_i.  One CTA per problem.
__global__ void __launch_bounds__(512) d_ldl_solve_kernel(GjArgs<double> a, const double* __restrict__ Min,
                                                          const double* __restrict__ Pin, int nb) {
  extern __shared__ __align__(16) unsigned char ldl_smem_d[];
  constexpr int NT = 512, NW = NT / 32;
  const int np = nb * kTB, n = a.n, m = a.m;
  double* y = reinterpret_cast<double*>(ldl_smem_d);   // [np]
  double* z = y + np;                                  // [np]
  double* red = z + np;                                // [4][128]
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double* Mb = Min + ((size_t)b * ((size_t)nb * (nb + 1) / 2)) * kTBE;
  const double* Pb = Pin + (size_t)b * nb * kTBE;
  const double* maskb = a.mask + (size_t)b * a.ldm;
  for (int i = tid; i < np; i += NT) y[i] = (i < n) ? -(maskb[i] * a.rhs_g[(size_t)b * n + i]) : 0.0;   // :368-375
  __syncthreads();
  // a warp takes 4 rows per turn: lane covers columns 4 lane .. + 3 of the 128-wide block row
  auto row_dots = [&](const double* rows, const double* vec, double (&out)[4]) {
    const double2 v0 = *reinterpret_cast<const double2*>(vec + 4 * lane), v1 = *reinterpret_cast<const double2*>(vec + 4 * lane + 2);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const double2 l0 = *reinterpret_cast<const double2*>(rows + (size_t)q * kTB + 4 * lane);
      const double2 l1 = *reinterpret_cast<const double2*>(rows + (size_t)q * kTB + 4 * lane + 2);
      out[q] = l0.x * v0.x + l0.y * v0.y + l1.x * v1.x + l1.y * v1.y;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) out[q] = warp_sum(out[q]);
  };
  for (int k = 0; k + 1 < nb; ++k) {
    for (int r4 = 4 * warp; r4 < (nb - 1 - k) * kTB; r4 += 4 * NW) {
      const int i = k + 1 + r4 / kTB, r = r4 % kTB;
      double acc[4];
      row_dots(Mb + bl_tile(i, k) + (size_t)r * kTB, y + k * kTB, acc);
      if (lane == 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) y[i * kTB + r + q] -= acc[q];
      }
    }
    __syncthreads();
  }
  for (int r4 = 4 * warp; r4 < np; r4 += 4 * NW) {
    const int k = r4 / kTB, r = r4 % kTB;
    double acc[4];
    row_dots(Pb + (size_t)k * kTBE + (size_t)r * kTB, y + k * kTB, acc);
    if (lane == 0) {
#pragma unroll
      for (int q = 0; q < 4; ++q) z[r4 + q] = acc[q];
    }
  }
  __syncthreads();
  const int c = tid & 127, chunk = tid >> 7;        // 4 row chunks of 32
  for (int k = nb - 2; k >= 0; --k) {
    double acc = 0.0;
    for (int i = k + 1; i < nb; ++i) {
      const double* L = Mb + bl_tile(i, k);
#pragma unroll 8
      for (int r = chunk * 32; r < chunk * 32 + 32; ++r) acc += L[(size_t)r * kTB + c] * z[i * kTB + r];
    }
    red[chunk * kTB + c] = acc;
    __syncthreads();
    if (tid < kTB) z[k * kTB + tid] -= red[tid] + red[kTB + tid] + red[2 * kTB + tid] + red[3 * kTB + tid];
    __syncthreads();
  }
  for (int i = tid; i < a.ldd; i += NT) a.sol_x[(size_t)b * a.ldd + i] = i < n ? z[i] : 0.0;
  for (int l = tid; l < m; l += NT) a.sol_nu[(size_t)b * m + l] = z[n + l];
}

// ------------------------------------------------------------------ host orchestration
static cudaError_t d_sweep(int B, const DArgs& base, bool ldl, cudaStream_t st, int* launches) {
  static std::atomic<int> sm_of_dev[64];          // function attributes and the SM count are per device
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const bool slot_ok = dev >= 0 && dev < 64;
  int n_sm = slot_ok ? sm_of_dev[dev].load(std::memory_order_acquire) : 0;
  if (n_sm == 0) {
    e = cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(d_tile_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDSmem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(d_tile_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDSmem);
    if (e != cudaSuccess) return e;
    if (slot_ok) sm_of_dev[dev].store(n_sm, std::memory_order_release);
  }
  DArgs a = base;
  a.ldl = ldl ? 1 : 0;
  const int nb = a.nb;
  for (int k = 0; k < nb; ++k) {
    a.k = k;
    static const bool piv_v1 = [] { const char* e = getenv("LQPB_D_PIVOT"); return e && e[0] == '1'; }();
    if (piv_v1) d_pivot_kernel<<<B, 512, 0, st>>>(a);            // developer switch: rank-1 register sweep (A/B)
    else d_pivot8_kernel<<<B, kDP8Threads, 0, st>>>(a);
    ++*launches;
    const int span = ldl ? nb - 1 - k : nb - 1;
    if (span > 0) {
      const int jp = B * span, jt = B * (span * (span + 1) / 2);
      d_tile_kernel<0><<<jp < n_sm ? jp : n_sm, kDThreads, kDSmem, st>>>(a, jp);
      d_tile_kernel<1><<<jt < n_sm ? jt : n_sm, kDThreads, kDSmem, st>>>(a, jt);
      *launches += 2;
    }
  }
  return cudaGetLastError();
}

cudaError_t launch_tc_inverse(int B, const GjArgs<double>& a, double* Pbuf, int nb, bool prebuilt, cudaStream_t st,
                              int* launches) {
  if (prebuilt) {
    d_fixup_kernel<<<B, 256, 0, st>>>(a, a.W, nb);
  } else {
    dim3 ga(nb * (nb + 1) / 2, B);
    d_assemble_kernel<<<ga, 256, 0, st>>>(a, a.W, nb);
  }
  ++*launches;
  DArgs t{a.W, a.Wg, a.Vg, Pbuf, nb, 0, 0};
  cudaError_t e = d_sweep(B, t, false, st, launches);
  if (e != cudaSuccess) return e;
  const int ntl = Pack<double>::ntiles(a.n);
  dim3 ge((ntl + 15) / 16, B);
  d_extract_kernel<<<ge, 512, 0, st>>>(a, a.W, nb);
  ++*launches;
  return cudaGetLastError();
}

cudaError_t launch_tc_ldl_solve(int B, const GjArgs<double>& a, double* Pbuf, int nb, cudaStream_t st, int* launches,
                                int stage) {
  if (stage != 2) {
    dim3 ga(nb * (nb + 1) / 2, B);
    d_assemble_kernel<<<ga, 256, 0, st>>>(a, a.W, nb);
    ++*launches;
    DArgs t{a.W, a.Wg, a.Vg, Pbuf, nb, 0, 1};
    cudaError_t e = d_sweep(B, t, true, st, launches);
    if (e != cudaSuccess) return e;
  }
  if (stage != 1) {
    const size_t smem = ((size_t)2 * nb * kTB + 4 * kTB) * sizeof(double);
    if (smem > 48 * 1024) {                         // more than 23 block rows: per-launch opt-in (cheap next to the solve)
      cudaError_t e = cudaFuncSetAttribute(d_ldl_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
    }
    d_ldl_solve_kernel<<<B, 512, smem, st>>>(a, a.W, Pbuf, nb);
    ++*launches;
  }
  return cudaGetLastError();
}

}  // namespace lqpb
