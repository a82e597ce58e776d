// K2 (fp32, n + m > 128) -- the batched KKT factorisation on the 5th-generation tensor cores.
//
// Same mathematics as factor.cu (symmetric Gauss-Jordan "sweep" of the quasi-definite KKT matrix
//     M = [[H, A^T], [A, d I]],  H = Q~ + rho I (forward, d = 0)  |  masked Q + 1e-8 I (backward, d = 1e-8),
// replacing torch.linalg.lu_factor(M) of solve_box_qp_admm_torch.py:206-215, :252-254 and the fresh LU inside
// torch.linalg.solve of the backward, :393), but blocked with 128 x 128 blocks so that every O(n^3) term is a
// 128 x 128 x 128 product issued as tcgen05.mma (kind::tf32, accumulators in TMEM):
//
//   for k = 0 .. nb-1                                   (sweep of block k; nb = np / 128)
//     P    = inv(M_kk)                 tc_pivot_kernel  (128 x 128 register-tiled sweep, FP32 pipe, one CTA / problem)
//     W_i  = M_ik P       (i != k)     tc_tile_kernel<PANEL>   also keeps V_i = old M_ik and stores M_ik <- W_i
//     M_ij = M_ij - W_i V_j^T (i,j!=k) tc_tile_kernel<TRAIL>
//     M_kk = -P
//   after nb sweeps the buffer holds -(M^-1); tc_extract_kernel writes K11 in the packed layout of the
//   iteration kernel plus K21, K22 and c = K21^T b~.
//   LDL mode (backward): only i, j > k are touched (block LDL^T, one third of the products), every P_k is kept,
//   and tc_ldl_solve_kernel does the block forward / backward substitution for the single right-hand side.
//
// fp32 accuracy on TF32 tensor cores: every operand x is split as x = hi + lo with hi = rna_tf32(x) and
// lo = rna_tf32(x - hi); a product is accumulated as hi*hi + lo*hi + hi*lo (fp32 accumulators in TMEM, zero
// initialised; the C tile is added in registers by the epilogue).  The neglected lo*lo term is <= 2^-22 relative.
//
// Storage ("block lower"): only tiles (I, J) with I >= J are kept, tile index I (I + 1) / 2 + J, each 128 x 128
// row-major and contiguous; diagonal tiles hold both triangles.  Operands are staged in shared memory by the
// CTA's threads in the canonical K-major SWIZZLE_128B layout (8-row x 128-byte atoms, 16-byte chunk index
// XOR row % 8), 32 k-columns (one 16 KB slab per operand part) at a time, while the previous slab's MMAs run.
#include <atomic>
#include "tcmma.cuh"

namespace lqpb {


// ------------------------------------------------------------------ the tile product kernel
// MODE 0 (PANEL): W_i = X P with X = M_ik (tile (i,k), or tile (k,i) read transposed when i < k), Y = P_k.
//                 Side effects: V_i = X (raw copy), Wbuf_i = W_i, M_ik <- W_i (transposed store when i < k).
// MODE 1 (TRAIL): M_ij <- M_ij - W_i V_j^T for the lower tiles i >= j (i, j != k; LDL: i >= j > k).
//
// Persistent CTAs (one per SM, 512 threads), jobs = output tiles handed out round-robin.  The two 128 x 128
// operand tiles of a job are fetched into registers (16 x 16 bytes per thread) one job ahead, so the global /
// L2 latency is covered by the previous tile's MMAs and epilogue.  Per job the four 32-column K slabs are split
// into hi / lo parts and stored in the canonical K-major SWIZZLE_128B layout into one of two shared-memory
// stages (Xhi | Xlo | Yhi | Ylo, 64 KB each): the split of slab s + 1 overlaps the 12 MMAs of slab s, whose
// completion (tcgen05.commit -> mbarrier) frees the stage again.  Accumulators: 2 x 128 TMEM columns.
#ifdef LQPB_PHASE_TIMERS
__device__ long long g_tc_cycles[16];
#define TC_T0() long long tc_t__ = clock64()
#define TC_ADD(k) do { if (blockIdx.x == 0 && threadIdx.x == 0) { long long n__ = clock64(); \
    atomicAdd((unsigned long long*)&g_tc_cycles[(MODE) * 8 + (k)], (unsigned long long)(n__ - tc_t__)); tc_t__ = n__; } } while (0)
#else
#define TC_T0()
#define TC_ADD(k)
#endif

struct TcJob {
  const float* xsrc;
  const float* ysrc;
  float* vdst;      // PANEL: raw copy of X
  float* wdst;      // PANEL: W_i
  float* mdst;      // PANEL: tile of M that receives W_i (transposed if xtrans) ; TRAIL: the C tile
  int xtrans;
};

template <int MODE>
__device__ __forceinline__ TcJob tc_decode_job(const TcArgs& a, int job_global) {
  const int k = a.k, nb = a.nb;
  const int span = a.ldl ? nb - 1 - k : nb - 1;                  // block indices taking part (besides k)
  const int jobs = MODE == 0 ? span : span * (span + 1) / 2;
  const int b = job_global / jobs, job = job_global % jobs;
  int i, j = 0;
  if (MODE == 0) {
    i = a.ldl ? k + 1 + job : (job < k ? job : job + 1);
  } else {
    int ii = (int)((sqrtf(8.f * job + 1.f) - 1.f) * 0.5f);
    while ((ii + 1) * (ii + 2) / 2 <= job) ++ii;
    while (ii * (ii + 1) / 2 > job) --ii;
    const int jj = job - ii * (ii + 1) / 2;
    if (a.ldl) { i = k + 1 + ii; j = k + 1 + jj; }
    else { i = ii < k ? ii : ii + 1; j = jj < k ? jj : jj + 1; }
  }
  const size_t ntile = (size_t)nb * (nb + 1) / 2;
  float* Mb = a.M + (size_t)b * ntile * kTBE;
  float* Wb = a.Wbuf + (size_t)b * nb * kTBE;
  float* Vb = a.Vbuf + (size_t)b * nb * kTBE;
  TcJob t;
  if (MODE == 0) {
    t.xtrans = i < k;
    t.mdst = Mb + (t.xtrans ? bl_tile(k, i) : bl_tile(i, k));
    t.xsrc = t.mdst;
    t.ysrc = a.Pbuf + ((size_t)b * nb + k) * kTBE;
    t.vdst = Vb + (size_t)i * kTBE;
    t.wdst = Wb + (size_t)i * kTBE;
  } else {
    t.xtrans = 0;
    t.xsrc = Wb + (size_t)i * kTBE;
    t.ysrc = Vb + (size_t)j * kTBE;
    t.mdst = Mb + bl_tile(i, j);
    t.vdst = nullptr;
    t.wdst = nullptr;
  }
  return t;
}

template <int MODE>
__global__ void __launch_bounds__(kTcThreads, 1) tc_tile_kernel(TcArgs a, int total_jobs) {
  extern __shared__ unsigned char tc_smem_raw[];
  __shared__ __align__(8) uint64_t mma_bar[2];
  __shared__ __align__(8) uint64_t c_bar;      // TRAIL: the C tile has landed in shared memory
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- shared memory carve (1024-byte aligned for the 128-byte swizzle), barriers, TMEM
  const uint32_t sbase = (smem_u32(tc_smem_raw) + 1023u) & ~1023u;
  unsigned char* sptr = tc_smem_raw + (sbase - smem_u32(tc_smem_raw));
  if (warp == 0) tmem_alloc(&tmem_slot, kTcCols);
  if (tid == 32) {
    mbar_init(&mma_bar[0], 1);
    mbar_init(&mma_bar[1], 1);
    mbar_init(&c_bar, 1);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  uint32_t ph[2] = {0u, 0u};          // parity of the next completion to wait for, per stage barrier
  uint32_t c_ph = 0u;
  // TRAIL: the 64 KB C tile of a job (contiguous in the block-lower storage) is fetched by ONE bulk TMA copy at
  // the start of the job and waits in shared memory until the epilogue: the read-modify-write of C then has no
  // global-load latency left (it used to queue behind the 128 KB operand prefetch of the next job).
  float* cbuf = reinterpret_cast<float*>(sptr + 8 * kSlabBytes);

  float4 xr[8], yr[8];                 // the whole X / Y tile pair of one job: [slab * 2 + t]
  auto gload = [&](const TcJob& jb) {
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int q = tid + kTcThreads * t, r = q >> 3, ch = q & 7;
        if (!jb.xtrans) {
          xr[2 * s + t] = *reinterpret_cast<const float4*>(jb.xsrc + (size_t)r * kTB + 32 * s + 4 * ch);
        } else {
          const int r4 = 4 * (warp + 16 * t);      // logical rows r4 .. r4+3 at logical column 32 s + lane
          xr[2 * s + t] = *reinterpret_cast<const float4*>(jb.xsrc + (size_t)(32 * s + lane) * kTB + r4);
        }
        yr[2 * s + t] = *reinterpret_cast<const float4*>(jb.ysrc + (size_t)r * kTB + 32 * s + 4 * ch);
      }
  };

  int job = blockIdx.x;
  TcJob cur{};
  if (job < total_jobs) {
    cur = tc_decode_job<MODE>(a, job);
    gload(cur);
  }
  TC_T0();
#pragma unroll 1
  for (; job < total_jobs; job += gridDim.x) {
    if (MODE == 1 && tid == 0) {      // the previous job's reads of cbuf ended before its closing __syncthreads
      mbar_arrive_expect_tx(&c_bar, (uint32_t)(kTBE * sizeof(float)));
      tma_load_1d(cbuf, cur.mdst, (uint32_t)(kTBE * sizeof(float)), &c_bar);
    }
    TC_ADD(0);
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int buf = s & 1;
      unsigned char* sXh = sptr + buf * (4 * kSlabBytes);
      unsigned char* sXl = sXh + kSlabBytes;
      unsigned char* sYh = sXh + 2 * kSlabBytes;
      unsigned char* sYl = sXh + 3 * kSlabBytes;
      if (s >= 2) {                   // the MMAs of slab s - 2 have finished reading this stage
        mbar_wait(&mma_bar[buf], ph[buf]);
        ph[buf] ^= 1u;
      }
      TC_ADD(1);
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int q = tid + kTcThreads * t, r = q >> 3, ch = q & 7;
        const uint32_t off = (uint32_t)r * 128u + (uint32_t)((ch ^ (r & 7)) << 4);
        float4 hi, lo;
        const float4 xv = xr[2 * s + t];
        if (!cur.xtrans) {
          split4(xv, hi, lo);
          *reinterpret_cast<float4*>(sXh + off) = hi;
          *reinterpret_cast<float4*>(sXl + off) = lo;
          if (MODE == 0) *reinterpret_cast<float4*>(cur.vdst + (size_t)r * kTB + 32 * s + 4 * ch) = xv;
        } else {
          const int r4 = 4 * (warp + 16 * t);
          const float xe[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int rr = r4 + e;
            const uint32_t o = (uint32_t)rr * 128u + (uint32_t)((((lane >> 2) ^ (rr & 7)) << 4) + ((lane & 3) << 2));
            float h1, l1;
            split_tf32(xe[e], h1, l1);
            *reinterpret_cast<float*>(sXh + o) = h1;
            *reinterpret_cast<float*>(sXl + o) = l1;
            if (MODE == 0) cur.vdst[(size_t)rr * kTB + 32 * s + lane] = xe[e];
          }
        }
        split4(yr[2 * s + t], hi, lo);
        *reinterpret_cast<float4*>(sYh + off) = hi;
        *reinterpret_cast<float4*>(sYl + off) = lo;
      }
      TC_ADD(2);
      fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
      __syncthreads();
      TC_ADD(3);
      if (tid == 0) {
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint32_t ko = (uint32_t)kk * 32u;   // 8 tf32 = 32 bytes along K inside the swizzle atom
          const uint64_t dXh = umma_desc(smem_u32(sXh) + ko), dXl = umma_desc(smem_u32(sXl) + ko);
          const uint64_t dYh = umma_desc(smem_u32(sYh) + ko), dYl = umma_desc(smem_u32(sYl) + ko);
          // The tensor core truncates the fp32 accumulator after every MMA, an error proportional to the
          // accumulator's magnitude: the two small cross terms therefore get their own accumulator (columns
          // 128..255), so that the large hi*hi sum sees 16 instead of 48 roundings; the epilogue adds the two.
          const uint32_t tcross = tmem + (a.acc2 ? 128u : 0u);
          umma_tf32(tcross, dXl, dYh, kIdescTf32, (s | kk) ? 1u : 0u);
          umma_tf32(tcross, dXh, dYl, kIdescTf32, 1u);
          umma_tf32(tmem, dXh, dYh, kIdescTf32, (a.acc2 && !(s | kk)) ? 0u : 1u);
        }
        umma_commit(&mma_bar[buf]);
      }
      TC_ADD(4);
    }
    // ---- the operands of the next job start their trip now (the registers are free again)
    const TcJob done = cur;
    if (job + (int)gridDim.x < total_jobs) {
      cur = tc_decode_job<MODE>(a, job + gridDim.x);
      gload(cur);
    }
    // ---- all MMAs of this tile (slabs 2 and 3 are the outstanding commits)
    mbar_wait(&mma_bar[0], ph[0]);
    ph[0] ^= 1u;
    mbar_wait(&mma_bar[1], ph[1]);
    ph[1] ^= 1u;
    tc_fence_after();
    TC_ADD(5);

    // ---- epilogue.  TMEM -> registers gives every thread 16 consecutive columns of ITS row (32 (warp % 4) +
    // lane; warp / 4 selects a 32-column group): written straight to global memory that is 32 different rows
    // per instruction.  The first stage is free now, so the tile is staged there (16-byte chunk index XOR row:
    // conflict-free both ways) and then moved with full-row 512-byte warp accesses.
    float* stage = reinterpret_cast<float*>(sptr);
    {
      const int r = 32 * (warp & 3) + lane;
      const uint32_t trow = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
#pragma unroll 1
      for (int g = 0; g < 2; ++g) {
        const int c0 = (warp >> 2) * 32 + 16 * g;
        uint32_t v[16];
        float f[16];
        tmem_ld16(trow + (uint32_t)c0, v);
        tmem_ld_wait(v);
#pragma unroll
        for (int e = 0; e < 16; ++e) f[e] = __uint_as_float(v[e]);
        if (a.acc2) {
          tmem_ld16(trow + 128u + (uint32_t)c0, v);
          tmem_ld_wait(v);
#pragma unroll
          for (int e = 0; e < 16; ++e) f[e] += __uint_as_float(v[e]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int ch = (c0 >> 2) + q;
          *reinterpret_cast<float4*>(stage + r * kTB + ((ch ^ (r & 31)) << 2)) =
              make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
        }
      }
    }
    tc_fence_before();              // TMEM reads are complete before the next job's first MMA overwrites it
    __syncthreads();
    TC_ADD(6);
    constexpr int NW = kTcThreads / 32;
    if (MODE == 1) {
      mbar_wait(&c_bar, c_ph);
      c_ph ^= 1u;
#pragma unroll 4
      for (int rr = warp; rr < kTB; rr += NW) {
        const float4 d = *reinterpret_cast<const float4*>(stage + rr * kTB + ((lane ^ (rr & 31)) << 2));
        float4 c = *reinterpret_cast<const float4*>(cbuf + rr * kTB + 4 * lane);
        c.x -= d.x; c.y -= d.y; c.z -= d.z; c.w -= d.w;
        *reinterpret_cast<float4*>(done.mdst + (size_t)rr * kTB + 4 * lane) = c;
      }
    } else {
#pragma unroll 4
      for (int rr = warp; rr < kTB; rr += NW) {
        const float4 d = *reinterpret_cast<const float4*>(stage + rr * kTB + ((lane ^ (rr & 31)) << 2));
        *reinterpret_cast<float4*>(done.wdst + (size_t)rr * kTB + 4 * lane) = d;
        if (!done.xtrans) *reinterpret_cast<float4*>(done.mdst + (size_t)rr * kTB + 4 * lane) = d;
      }
      if (done.xtrans) {
        // M_ki = W_i^T: output row c (= W column), 4 consecutive W rows per lane (transposed read of the stage)
#pragma unroll 2
        for (int c = warp; c < kTB; c += NW) {
          float o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int wr = 4 * lane + e;
            o[e] = stage[wr * kTB + ((((c >> 2) ^ (wr & 31)) << 2) | (c & 3))];
          }
          *reinterpret_cast<float4*>(done.mdst + (size_t)c * kTB + 4 * lane) = make_float4(o[0], o[1], o[2], o[3]);
        }
      }
    }
    __syncthreads();                // the stage is rewritten by the next job's first slab
    TC_ADD(7);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, kTcCols);
}

// ------------------------------------------------------------------ pivot block inverse
// One CTA (512 threads) per problem: symmetric sweep of the 128 x 128 pivot tile M_kk held in registers
// (thread (ta, tb) of a 32 x 16 grid owns the 4 x 8 sub-block rows 4 ta.., columns 8 tb..); per sweep step the
// pivot row and column are broadcast through double-buffered shared vectors, one __syncthreads per step, and
// every thread applies the rank-1 update to its 32 entries.  The tile ends as -(M_kk)^-1: P = inv(M_kk) goes to
// Pbuf[k], -P back into the matrix.
__global__ void __launch_bounds__(kPivThreads, 1) tc_pivot_kernel(TcArgs a) {
  __shared__ __align__(16) float rowbuf[2][kTB];
  __shared__ __align__(16) float colbuf[2][kTB];
  const int b = blockIdx.x, tid = threadIdx.x, ta = tid >> 4, tb = tid & 15;   // rows 4 ta.., columns 8 tb..
  const int k = a.k, nb = a.nb;
  float* tile = a.M + ((size_t)b * ((size_t)nb * (nb + 1) / 2)) * kTBE + bl_tile(k, k);
  float acc[4][8];
  // the lower triangle is the reference copy: entries above the diagonal are read transposed
#pragma unroll
  for (int rr = 0; rr < 4; ++rr)
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int r = 4 * ta + rr, c = 8 * tb + q;
      acc[rr][q] = r >= c ? tile[(size_t)r * kTB + c] : tile[(size_t)c * kTB + r];
    }
  // the step loop is unrolled by 8 so that the row / column index inside the thread's tile is a
  // compile-time constant: no dynamically indexed registers
#pragma unroll 1
  for (int sb = 0; sb < kTB / 8; ++sb) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int s = 8 * sb + j, par = j & 1;
      const int sr = j & 3, ra = 2 * sb + (j >> 2);   // pivot row s = 4 ra + sr ; pivot column s = 8 sb + j
      if (ta == ra) {          // publish pivot row s
        *reinterpret_cast<float4*>(&rowbuf[par][8 * tb]) = make_float4(acc[sr][0], acc[sr][1], acc[sr][2], acc[sr][3]);
        *reinterpret_cast<float4*>(&rowbuf[par][8 * tb + 4]) = make_float4(acc[sr][4], acc[sr][5], acc[sr][6], acc[sr][7]);
      }
      if (tb == sb)            // publish pivot column s
        *reinterpret_cast<float4*>(&colbuf[par][4 * ta]) = make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
      __syncthreads();
      const float piv = __frcp_rn(rowbuf[par][s]);
      const float4 a0 = *reinterpret_cast<const float4*>(&rowbuf[par][8 * tb]);
      const float4 a1 = *reinterpret_cast<const float4*>(&rowbuf[par][8 * tb + 4]);
      const float4 c0 = *reinterpret_cast<const float4*>(&colbuf[par][4 * ta]);
      const float asc[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float tr[4] = {c0.x * piv, c0.y * piv, c0.z * piv, c0.w * piv};
#pragma unroll
      for (int rr = 0; rr < 4; ++rr)
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[rr][q] = fmaf(-tr[rr], asc[q], acc[rr][q]);
      if (tb == sb) {          // own column s: a[r][s] <- a[r][s] / piv
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) acc[rr][j] = tr[rr];
      }
      if (ta == ra) {          // own row s: a[s][c] <- a[s][c] / piv, a[s][s] <- -1 / piv
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[sr][q] = (tb == sb && q == j) ? -piv : asc[q] * piv;
      }
    }
  }
  float* P = a.Pbuf + ((size_t)b * nb + k) * kTBE;
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) {
    const size_t o = (size_t)(4 * ta + rr) * kTB + 8 * tb;
    *reinterpret_cast<float4*>(tile + o) = make_float4(acc[rr][0], acc[rr][1], acc[rr][2], acc[rr][3]);
    *reinterpret_cast<float4*>(tile + o + 4) = make_float4(acc[rr][4], acc[rr][5], acc[rr][6], acc[rr][7]);
    *reinterpret_cast<float4*>(P + o) = make_float4(-acc[rr][0], -acc[rr][1], -acc[rr][2], -acc[rr][3]);
    *reinterpret_cast<float4*>(P + o + 4) = make_float4(-acc[rr][4], -acc[rr][5], -acc[rr][6], -acc[rr][7]);
  }
}

__global__ void __launch_bounds__(kPiv8Threads, 1) tc_pivot8_kernel(TcArgs a) {
  const int b = blockIdx.x, k = a.k, nb = a.nb;
  float* tile = a.M + ((size_t)b * ((size_t)nb * (nb + 1) / 2)) * kTBE + bl_tile(k, k);
  __shared__ __align__(16) float scratch[kPivScratchFloats];
  pivot8_body(tile, a.Pbuf + ((size_t)b * nb + k) * kTBE, scratch);
}

// (Round 2 measured the same sweep with 16 pivots per step -- 8 rank-16 updates, rows S in two warps, the look-ahead warp
// forming and inverting a 16 x 16 block by a ping-pong sweep in shared memory; git history, "tc_pivot16_kernel".  Correct
// (all parity tests), but the look-ahead chain -- 16 serial shared-memory sweep steps behind a 16-step D' update on ONE
// warp -- became the critical path: 84 us per launch against 39 us here.  Rejected; the 8 x 8 inverse in registers with
// shuffles is what keeps the look-ahead off the critical path.)
// (A tensor-core variant of the pivot-block inverse was built and measured in round 1 -- the 16 rank-8 updates as
// tcgen05.mma M = N = 128, K = 8 instructions accumulating in TMEM, rows read back with tcgen05.ld, exact 8 x 8
// diagonal blocks in a side buffer, look-ahead inverse on a 17th warp; git history, "LQPB_TC_PIVOT=m".  Its inverse
// was as accurate as this one (8.9e-7 against fp64), but a sequential chain of 16 steps pays the fixed latencies of
// the asynchronous path 16 times -- fence.proxy.async + tcgen05.commit / mbarrier + tcgen05.ld round trips, 2.6 us
// per step measured -- and the launch took 52.7 us against 41 us for the FP32-pipe kernel above.  Rejected.)

// ------------------------------------------------------------------ assemble the KKT matrix (block-lower tiles)
// Same embedding as the prologue of gj_inverse_kernel: H masked / shifted, the m equality rows right below it,
// identity on the padding.  grid = (tiles, B).
__global__ void __launch_bounds__(256) tc_assemble_kernel(GjArgs<float> a, float* __restrict__ Mout, int nb) {
  const int b = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
  int I = (int)((sqrtf(8.f * tile + 1.f) - 1.f) * 0.5f);
  while ((I + 1) * (I + 2) / 2 <= tile) ++I;
  while (I * (I + 1) / 2 > tile) --I;
  const int J = tile - I * (I + 1) / 2;
  const int n = a.n, m = a.m;
  const bool packed_src = a.lds == 0;
  const int ntv = Pack<float>::nt(n);
  const float* srcb = packed_src ? a.src + (size_t)b * Pack<float>::elems(n) : a.src + (size_t)b * n * a.lds;
  const float* maskb = a.mask ? a.mask + (size_t)b * a.ldm : nullptr;
  const float* Ab = (m > 0) ? a.Arows + (size_t)b * m * a.lda : nullptr;
  const float shift = (a.diag_shift ? a.diag_shift[b] : 0.f) + a.diag_const;
  const float* dvecb = a.diag_vec ? a.diag_vec + (size_t)b * a.ldm : nullptr;
  float* dst = Mout + ((size_t)b * ((size_t)nb * (nb + 1) / 2)) * kTBE + bl_tile(I, J);
  // fast path for tiles that lie entirely inside the H block of a dense row-major source (the backward's Q): 16-byte
  // loads, four of them in flight per thread; rows / columns that cross n, the equality rows and the padding take
  // the element-wise path below.  Entries above the diagonal of a diagonal tile are written too (never read).
  if (!packed_src && (I + 1) * kTB <= n && (a.lds & 3) == 0 && (reinterpret_cast<uintptr_t>(srcb) & 15) == 0) {
#pragma unroll 4
    for (int e4 = tid; e4 < kTBE / 4; e4 += 256) {
      const int r = e4 >> 5, c = (e4 & 31) << 2;
      const int i = I * kTB + r, j = J * kTB + c;
      float4 v = *reinterpret_cast<const float4*>(srcb + (size_t)i * a.lds + j);
      if (maskb) {
        const float fi = maskb[i];
        const float4 fj = *reinterpret_cast<const float4*>(maskb + j);
        v.x = (fi != 0.f && fj.x != 0.f) ? v.x : 0.f;
        v.y = (fi != 0.f && fj.y != 0.f) ? v.y : 0.f;
        v.z = (fi != 0.f && fj.z != 0.f) ? v.z : 0.f;
        v.w = (fi != 0.f && fj.w != 0.f) ? v.w : 0.f;
      }
      if (I == J && i >= j && i < j + 4) {       // the diagonal entry of row i is one of these four
        const bool keep = !maskb || maskb[i] != 0.f;
        const float add = shift + (dvecb ? dvecb[i] : 0.f);
        const int d = i - j;
        if (d == 0) v.x = keep ? v.x + add : 1.f;
        else if (d == 1) v.y = keep ? v.y + add : 1.f;
        else if (d == 2) v.z = keep ? v.z + add : 1.f;
        else v.w = keep ? v.w + add : 1.f;
      }
      *reinterpret_cast<float4*>(dst + (size_t)r * kTB + c) = v;
    }
    return;
  }
  for (int e = tid; e < kTBE; e += 256) {
    const int r = e >> 7, c = e & 127;
    const int i = I * kTB + r, j = J * kTB + c;
    if (j > i) continue;     // the strict upper triangle of a diagonal tile is never read (see tc_pivot_kernel)
    float v = 0.f;
    if (i < n) {
      const float fi = maskb ? maskb[i] : 1.f, fj = maskb ? maskb[j] : 1.f;
      const bool keep = fi != 0.f && fj != 0.f;
      if (keep) {
        if (packed_src) {
          v = srcb[Pack<float>::offset(i, j, ntv)];
          if (i == j) v += v;                          // the packed layout stores the diagonal halved
        } else {
          v = srcb[(size_t)i * a.lds + j];
        }
      }
      if (i == j) v = keep ? v + shift + (dvecb ? dvecb[i] : 0.f) : 1.f;
    } else if (i < n + m) {
      if (j < n) {
        const float av = Ab[(size_t)(i - n) * a.lda + j];
        v = maskb ? av * maskb[j] : av;
      } else if (j == i) {
        v = a.a_diag;
      }
    } else {
      v = (i == j) ? 1.f : 0.f;
    }
    dst[e] = v;
  }
}

// When scale_pack_kernel has already written the H block (entries i, j < n, no diagonal shift) into the
// block-lower tiles, only the shift, the m equality rows and the identity padding are missing.  grid = B.
__global__ void __launch_bounds__(256) tc_fixup_kernel(GjArgs<float> a, float* __restrict__ Mout, int nb) {
  const int b = blockIdx.x, tid = threadIdx.x;
  const int n = a.n, m = a.m, np = nb * kTB;
  float* Mb = Mout + ((size_t)b * ((size_t)nb * (nb + 1) / 2)) * kTBE;
  const float shift = (a.diag_shift ? a.diag_shift[b] : 0.f) + a.diag_const;
  for (int i = tid; i < n; i += 256) Mb[bl_off(i, i)] += shift;
  const float* Ab = (m > 0) ? a.Arows + (size_t)b * m * a.lda : nullptr;
  for (int e = tid; e < (np - n) * np; e += 256) {
    const int i = n + e / np, j = e % np;
    if (j > i) continue;
    float v = 0.f;
    if (i < n + m) {
      if (j < n) v = Ab[(size_t)(i - n) * a.lda + j];
      else if (j == i) v = a.a_diag;
    } else if (j == i) {
      v = 1.f;
    }
    Mb[bl_off(i, j)] = v;
  }
}

// ------------------------------------------------------------------ extract K11 (packed), K21, K22, c from -(M^-1)
__global__ void __launch_bounds__(512) tc_extract_kernel(GjArgs<float> a, const float* __restrict__ Min, int nb) {
  using P = Pack<float>;
  constexpr int NW = 512 / 32;
  // grid = (tile chunks, B): one packed tile per warp, 16 per CTA; the CTA of chunk 0 also writes K21, K22 and c
  const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = a.n, m = a.m;
  const float* Mb = Min + ((size_t)b * ((size_t)nb * (nb + 1) / 2)) * kTBE;
  float* dstb = a.dst + (size_t)b * P::elems(n);
  const int ntv = P::nt(n), ntl = P::ntiles(n);
  const int c = lane % P::TC, kc = c / P::VN, ec = c % P::VN;
  for (int t = blockIdx.x * NW + warp; t < ntl; t += gridDim.x * NW) {
    int Jc = 0, rem = t;
    while (rem >= ntv - Jc / P::R) { rem -= ntv - Jc / P::R; ++Jc; }
    const int I = Jc / P::R + rem;
    float* tp = dstb + (size_t)t * P::TILE;
    const int j = Jc * P::TC + c;
#pragma unroll 16
    for (int l0 = 0; l0 < kPackRows; l0 += P::R) {
      const int l = l0 + lane / P::TC, i = I * kPackRows + l;
      float v = 0.f;
      if (i < n && j <= i) {
        v = -Mb[bl_off(i, j)];
        if (i == j) v *= 0.5f;
      }
      tp[P::in_tile(l, kc, ec)] = v;
    }
  }
  if (blockIdx.x != 0) return;
  float* g21 = (m > 0) ? a.G21 + (size_t)b * m * a.ldd : nullptr;
  float* k22 = (m > 0) ? a.K22 + (size_t)b * m * m : nullptr;
  const int ldd = a.ldd;
  for (int r = warp; r < m; r += NW) {
    for (int j = lane; j < ldd; j += 32) g21[(size_t)r * ldd + j] = j < n ? -Mb[bl_off(n + r, j)] : 0.f;
    for (int q = lane; q < m; q += 32)
      k22[(size_t)r * m + q] = q <= r ? -Mb[bl_off(n + r, n + q)] : -Mb[bl_off(n + q, n + r)];
  }
  if (a.c_out) {
    __syncthreads();
    float* cb = a.c_out + (size_t)b * ldd;
    for (int i = tid; i < ldd; i += 512) {
      float acc = 0.f;
      if (i < n)
        for (int l = 0; l < m; ++l) acc += g21[(size_t)l * ldd + i] * a.bt[(size_t)b * m + l];
      cb[i] = acc;
    }
  }
}

// ------------------------------------------------------------------ block LDL^T solve (backward)
// M = L D L^T with L_ik = W_i of step k (stored in tile (i,k)), D_k^-1 = P_k.  Solves M d = [-mask * dl_dz; 0]:
// forward  y_i -= L_ik y_k,  z_k = P_k y_k,  backward d_k = z_k - sum_{i>k} L_ik^T d_i.  One CTA per problem.
__global__ void __launch_bounds__(512) tc_ldl_solve_kernel(GjArgs<float> a, const float* __restrict__ Min,
                                                           const float* __restrict__ Pin, int nb) {
  extern __shared__ __align__(16) unsigned char ldl_smem[];
  constexpr int NT = 512, NW = NT / 32;
  const int np = nb * kTB, n = a.n, m = a.m;
  float* y = reinterpret_cast<float*>(ldl_smem);   // [np]
  float* z = y + np;                               // [np]
  float* red = z + np;                             // [4][128]
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* Mb = Min + ((size_t)b * ((size_t)nb * (nb + 1) / 2)) * kTBE;
  const float* Pb = Pin + (size_t)b * nb * kTBE;
  const float* maskb = a.mask + (size_t)b * a.ldm;
  for (int i = tid; i < np; i += NT) y[i] = (i < n) ? -(maskb[i] * a.rhs_g[(size_t)b * n + i]) : 0.f;   // :368-375
  __syncthreads();
  for (int k = 0; k + 1 < nb; ++k) {
    const float4 yk = *reinterpret_cast<const float4*>(y + k * kTB + 4 * lane);
    // a warp takes 8 consecutive rows per turn (kTB is a multiple of 8, so they share the tile): 8 row loads in flight
    for (int r8 = 8 * warp; r8 < (nb - 1 - k) * kTB; r8 += 8 * NW) {
      const int i = k + 1 + r8 / kTB, r = r8 % kTB;
      const float* Lr = Mb + bl_tile(i, k) + (size_t)r * kTB + 4 * lane;
      float4 l4[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) l4[q] = *reinterpret_cast<const float4*>(Lr + (size_t)q * kTB);
      float acc[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q] = l4[q].x * yk.x + l4[q].y * yk.y + l4[q].z * yk.z + l4[q].w * yk.w;
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q] = warp_sum(acc[q]);
      if (lane == 0) {
#pragma unroll
        for (int q = 0; q < 8; ++q) y[i * kTB + r + q] -= acc[q];
      }
    }
    __syncthreads();
  }
  for (int r8 = 8 * warp; r8 < np; r8 += 8 * NW) {
    const int k = r8 / kTB, r = r8 % kTB;
    const float4 yk = *reinterpret_cast<const float4*>(y + k * kTB + 4 * lane);
    const float* Pr = Pb + (size_t)k * kTBE + (size_t)r * kTB + 4 * lane;
    float4 p4[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) p4[q] = *reinterpret_cast<const float4*>(Pr + (size_t)q * kTB);
    float acc[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = p4[q].x * yk.x + p4[q].y * yk.y + p4[q].z * yk.z + p4[q].w * yk.w;
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = warp_sum(acc[q]);
    if (lane == 0) {
#pragma unroll
      for (int q = 0; q < 8; ++q) z[r8 + q] = acc[q];
    }
  }
  __syncthreads();
  const int c = tid & 127, chunk = tid >> 7;        // 4 row chunks of 32
  for (int k = nb - 2; k >= 0; --k) {
    float acc = 0.f;
    for (int i = k + 1; i < nb; ++i) {
      const float* L = Mb + bl_tile(i, k);
#pragma unroll 16
      for (int r = chunk * 32; r < chunk * 32 + 32; ++r) acc += L[(size_t)r * kTB + c] * z[i * kTB + r];
    }
    red[chunk * kTB + c] = acc;
    __syncthreads();
    if (tid < kTB) z[k * kTB + tid] -= red[tid] + red[kTB + tid] + red[2 * kTB + tid] + red[3 * kTB + tid];
    __syncthreads();
  }
  for (int i = tid; i < a.ldd; i += NT) a.sol_x[(size_t)b * a.ldd + i] = i < n ? z[i] : 0.f;
  for (int l = tid; l < m; l += NT) a.sol_nu[(size_t)b * m + l] = z[n + l];
}

// ------------------------------------------------------------------ host orchestration
// The pivot kernel is latency-bound (one CTA per problem, a serial sweep of 128 steps) and the tile kernels are
// bound by memory latency, so the batch is cut into `groups` slices that run the same pivot -> panel -> trail
// chain on their own streams: the pivot sweep of one slice overlaps the tile products of the others.  Slice 0
// stays on the caller's stream; the auxiliary streams fork from / join into it with events, so the call keeps
// its stream-ordered semantics.
constexpr int kTcMaxGroups = 8;
struct TcStreams {
  int dev = -1;
  cudaStream_t aux[kTcMaxGroups - 1];
  cudaEvent_t fork, join[kTcMaxGroups - 1];
};
static thread_local TcStreams g_tcs_tab[64];     // streams and events belong to a device: one set per device ordinal
#define g_tcs (g_tcs_tab[tc_dev_slot()])
static int tc_dev_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  return dev;
}

static cudaError_t tc_streams_init() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (g_tcs.dev == dev) return cudaSuccess;
  for (int i = 0; i < kTcMaxGroups - 1; ++i) {
    e = cudaStreamCreateWithFlags(&g_tcs.aux[i], cudaStreamNonBlocking);
    if (e != cudaSuccess) return e;
    e = cudaEventCreateWithFlags(&g_tcs.join[i], cudaEventDisableTiming);
    if (e != cudaSuccess) return e;
  }
  e = cudaEventCreateWithFlags(&g_tcs.fork, cudaEventDisableTiming);
  if (e != cudaSuccess) return e;
  g_tcs.dev = dev;
  return cudaSuccess;
}

static int tc_groups(int B) {
  static const int cfg = [] {
    const char* e = getenv("LQPB_TC_GROUPS");     // developer switch; default chosen from measurements
    int c = e ? atoi(e) : 2;   // measured at dz=500, B=128: forward sweep 0.644 (1) / 0.593 (2) / 0.602 (3) / 0.608 ms (4)
    if (c < 1) c = 1;
    if (c > kTcMaxGroups) c = kTcMaxGroups;
    return c;
  }();
  int g = cfg;
  while (g > 1 && B / g < 16) --g;                // slices of fewer than 16 problems no longer fill the tile grids
  return g;
}

static cudaError_t tc_sweep(int B, const TcArgs& base, bool ldl, cudaStream_t st, int* launches) {
  // function attributes and the SM count are PER DEVICE (a process may drive several GPUs, and autograd calls in from
  // its own thread): one slot per device ordinal, published with release / acquire
  static std::atomic<int> sm_of_dev[64];
  int dev = 0;
  cudaError_t e0 = cudaGetDevice(&dev);
  if (e0 != cudaSuccess) return e0;
  const int slot = (dev >= 0 && dev < 64) ? dev : 0;
  int n_sm = (dev >= 0 && dev < 64) ? sm_of_dev[slot].load(std::memory_order_acquire) : 0;
  if (n_sm == 0) {
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaFuncSetAttribute(tc_tile_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(tc_tile_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemTrail);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) sm_of_dev[slot].store(n_sm, std::memory_order_release);
  }
  TcArgs a0 = base;
  a0.ldl = ldl ? 1 : 0;
  {
    const char* e = getenv("LQPB_TC_ACC2");     // developer switch (accuracy A/B); default on
    a0.acc2 = (e && e[0] == '0') ? 0 : 1;
  }
  // The fused persistent kernel (tcfused.cu: one launch for the whole sweep, CTA b owns problem b) where it is the faster
  // form -- batches that fill the device once and sweeps of up to 4 block rows (measured, DESIGN 5a: dz = 250 / 500 at
  // B = 128 a little ahead; B = 32, B = 256 and dz = 1000 behind the per-phase kernels, which spread a problem's tiles over
  // all SMs) -- else the per-phase kernels below.  LQPB_TC_FUSED=0 / 1 forces either (A/B measurements; bit-identical).
  static const int fused_cfg = [] { const char* e = getenv("LQPB_TC_FUSED"); return e ? (e[0] == '0' ? 0 : 1) : -1; }();
  const bool fused = fused_cfg >= 0 ? fused_cfg == 1 : (a0.nb <= 4 && B >= 64 && B <= n_sm);
  if (fused) {
    ++*launches;
    static const int dbg = [] { const char* e = getenv("LQPB_FU_DBG"); return e ? atoi(e) : 0; }();   // timing experiments only
    a0.k = dbg;
    return launch_tc_fused(B, a0, st);
  }
  const int nb = a0.nb;
  const int G = tc_groups(B);
  if (G > 1) {
    cudaError_t e = tc_streams_init();
    if (e != cudaSuccess) return e;
    e = cudaEventRecord(g_tcs.fork, st);
    if (e != cudaSuccess) return e;
    for (int g = 1; g < G; ++g) {
      e = cudaStreamWaitEvent(g_tcs.aux[g - 1], g_tcs.fork, 0);
      if (e != cudaSuccess) return e;
    }
  }
  const size_t ntile = (size_t)nb * (nb + 1) / 2;
  // k outermost: the slices are enqueued round-robin, which is also the order the hardware should start them in
  for (int k = 0; k < nb; ++k) {
    for (int g = 0; g < G; ++g) {
      const int b0 = (int)((long long)B * g / G), bc = (int)((long long)B * (g + 1) / G) - b0;
      if (bc <= 0) continue;
      cudaStream_t s = g == 0 ? st : g_tcs.aux[g - 1];
      TcArgs a = a0;
      a.M = a0.M + (size_t)b0 * ntile * kTBE;
      a.Wbuf = a0.Wbuf + (size_t)b0 * nb * kTBE;
      a.Vbuf = a0.Vbuf + (size_t)b0 * nb * kTBE;
      a.Pbuf = a0.Pbuf + (size_t)b0 * nb * kTBE;
      a.k = k;
      static const bool piv_v1 = [] { const char* e = getenv("LQPB_TC_PIVOT"); return e && e[0] == '1'; }();
      if (piv_v1) tc_pivot_kernel<<<bc, kPivThreads, 0, s>>>(a);      // developer switch: rank-1 sweep (A/B)
      else tc_pivot8_kernel<<<bc, kPiv8Threads, 0, s>>>(a);
      ++*launches;
      const int span = ldl ? nb - 1 - k : nb - 1;
      if (span > 0) {
        const int jp = bc * span, jt = bc * (span * (span + 1) / 2);
        tc_tile_kernel<0><<<jp < n_sm ? jp : n_sm, kTcThreads, kTcSmem, s>>>(a, jp);
        tc_tile_kernel<1><<<jt < n_sm ? jt : n_sm, kTcThreads, kTcSmemTrail, s>>>(a, jt);
        *launches += 2;
      }
    }
  }
  for (int g = 1; g < G; ++g) {
    cudaError_t e = cudaEventRecord(g_tcs.join[g - 1], g_tcs.aux[g - 1]);
    if (e != cudaSuccess) return e;
    e = cudaStreamWaitEvent(st, g_tcs.join[g - 1], 0);
    if (e != cudaSuccess) return e;
  }
  return cudaGetLastError();
}

// forward: inverse of the KKT matrix, outputs as launch_gj_inverse (a.W = block-lower work matrix,
// a.Vg / a.Wg = panel buffers (nb tiles per problem each), Pbuf = nb tiles per problem)
cudaError_t launch_tc_inverse(int B, const GjArgs<float>& a, float* Pbuf, int nb, bool prebuilt, cudaStream_t st,
                              int* launches) {
  if (prebuilt) {
    tc_fixup_kernel<<<B, 256, 0, st>>>(a, a.W, nb);
  } else {
    dim3 ga(nb * (nb + 1) / 2, B);
    tc_assemble_kernel<<<ga, 256, 0, st>>>(a, a.W, nb);
  }
  ++*launches;
  TcArgs t{a.W, a.Wg, a.Vg, Pbuf, nb, 0, 0, 1};
  cudaError_t e = tc_sweep(B, t, false, st, launches);
  if (e != cudaSuccess) return e;
  {
    const int ntl = Pack<float>::ntiles(a.n);
    dim3 ge((ntl + 15) / 16, B);
    tc_extract_kernel<<<ge, 512, 0, st>>>(a, a.W, nb);
  }
  ++*launches;
  return cudaGetLastError();
}

// backward: block LDL^T + solve, outputs as launch_ldl_solve
// stage 0: everything; 1: assemble + block LDL^T only (independent of the right-hand side); 2: substitution only
cudaError_t launch_tc_ldl_solve(int B, const GjArgs<float>& a, float* Pbuf, int nb, cudaStream_t st, int* launches,
                                int stage) {
  if (stage != 2) {
    dim3 ga(nb * (nb + 1) / 2, B);
    tc_assemble_kernel<<<ga, 256, 0, st>>>(a, a.W, nb);
    ++*launches;
    TcArgs t{a.W, a.Wg, a.Vg, Pbuf, nb, 0, 1, 1};
    cudaError_t e = tc_sweep(B, t, true, st, launches);
    if (e != cudaSuccess) return e;
  }
  if (stage != 1) {
    const size_t smem = ((size_t)2 * nb * kTB + 4 * kTB) * sizeof(float);
    tc_ldl_solve_kernel<<<B, 512, smem, st>>>(a, a.W, Pbuf, nb);
    ++*launches;
  }
  return cudaGetLastError();
}

// ------------------------------------------------------------------ developer entry: dense inverse of B SPD /
// quasi-definite N x N matrices (N multiple of 128) through the tensor-core sweep; used by tools/tc_check.py
__global__ void tc_dev_pack_kernel(const float* __restrict__ A, float* __restrict__ Mout, int N, int nb) {
  const int b = blockIdx.y, tile = blockIdx.x;
  int I = (int)((sqrtf(8.f * tile + 1.f) - 1.f) * 0.5f);
  while ((I + 1) * (I + 2) / 2 <= tile) ++I;
  while (I * (I + 1) / 2 > tile) --I;
  const int J = tile - I * (I + 1) / 2;
  float* dst = Mout + ((size_t)b * ((size_t)nb * (nb + 1) / 2)) * kTBE + bl_tile(I, J);
  for (int e = threadIdx.x; e < kTBE; e += blockDim.x) {
    const int i = I * kTB + (e >> 7), j = J * kTB + (e & 127);
    if (j <= i) dst[e] = A[((size_t)b * N + i) * N + j];
  }
}
__global__ void tc_dev_unpack_kernel(const float* __restrict__ Min, float* __restrict__ Ainv, int N, int nb) {
  const int b = blockIdx.y;
  const float* Mb = Min + ((size_t)b * ((size_t)nb * (nb + 1) / 2)) * kTBE;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < (size_t)N * N; e += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(e / N), j = (int)(e % N);
    Ainv[(size_t)b * N * N + e] = -(i >= j ? Mb[bl_off(i, j)] : Mb[bl_off(j, i)]);
  }
}
cudaError_t launch_tc_dev_inverse(int B, int N, const float* A, float* Ainv, float* work, cudaStream_t st) {
  const int nb = N / kTB;
  const size_t ntile = (size_t)nb * (nb + 1) / 2;
  float* M = work;
  float* Wb = M + (size_t)B * ntile * kTBE;
  float* Vb = Wb + (size_t)B * nb * kTBE;
  float* Pb = Vb + (size_t)B * nb * kTBE;
  dim3 g((unsigned)ntile, B);
  tc_dev_pack_kernel<<<g, 256, 0, st>>>(A, M, N, nb);
  TcArgs t{M, Wb, Vb, Pb, nb, 0, 0, 1};
  int launches = 0;
  cudaError_t e = tc_sweep(B, t, false, st, &launches);
  if (e != cudaSuccess) return e;
  dim3 g2(64, B);
  tc_dev_unpack_kernel<<<g2, 256, 0, st>>>(M, Ainv, N, nb);
  return cudaGetLastError();
}

}  // namespace lqpb

#ifdef LQPB_PHASE_TIMERS
// developer aid (tools/tc_phases.py): clock64 totals of thread 0 of CTA 0, [mode][phase]
extern "C" void lqpb_debug_tc_cycles(long long* out, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, lqpb::g_tc_cycles, sizeof(long long) * 16);
  if (reset) {
    long long z[16] = {0};
    cudaMemcpyToSymbol(lqpb::g_tc_cycles, z, sizeof(z));
  }
}
#endif
