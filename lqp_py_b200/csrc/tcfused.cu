// K2 (fp32, n + m > 128), fused form: the WHOLE block sweep of a problem inside one persistent CTA.
//
// Same mathematics, same tile storage and bit-identical results as the per-phase kernels of tcfactor.cu (blocked
// symmetric Gauss-Jordan sweep / block LDL^T of the KKT matrix with 128 x 128 blocks, every tile product as three TF32
// tcgen05.mma chains on hi / lo split operands, replacing torch.linalg.lu_factor of solve_box_qp_admm_torch.py:206-215,
// :252-254 and the LU inside torch.linalg.solve, :393), but scheduled differently: round 1 ran pivot -> PANEL -> TRAIL as
// 3 launches per block step and batch slice (48 launches per training step), every launch paying its own pipeline fill,
// the pivot launches leaving 84 of 148 SMs idle and the slices competing for SMs.  Here CTA b owns problem b (b + grid,
// ...) for all nb block steps: no grid-wide dependency is left, so there is ONE launch per factorisation and a tile
// product reads operands its own SM wrote moments before (L2 hits).  Measurements, and the shapes for which the
// per-phase kernels stay the better form, are in DESIGN.md 5a; the dispatch is tc_sweep (tcfactor.cu).
//
// Inside the CTA (17 warps) a block step is
//   pivot   all 17 warps: pivot8_body (tcmma.cuh) inverts M_kk on the FP32 pipe
//   PANEL   W_i = M_ik P_k     } warp-specialised pipeline over the step's tile jobs:
//   TRAIL   M_ij -= W_i V_j^T  }
//     warps  8-15  staging : operand slabs (32 K-columns of X and Y) global -> registers (one slab in flight per thread,
//                            every register refilled with the next slab's data as soon as it has been stored)
//                            -> hi / lo split -> K-major SWIZZLE_128B shared-memory stage (2 stages of 64 KB); PANEL
//                            also writes the raw copy V_i
//     warp   16    MMA     : one thread issues the 12 tcgen05.mma of a slab when its stage is full, tcgen05.commit frees
//                            the stage; two accumulator sets in TMEM (2 x 256 columns: hi*hi | cross terms)
//     warps  0-7   epilogue: TMEM -> registers -> swizzled shared tile -> coalesced global stores; TRAIL fetches the first
//                            half of its C tile into registers while the job's MMAs run, the second while it stores the first
//   so the split of job j + 1, the MMAs of job j and the write-back of job j - 1 overlap.  Stage hand-over is by
//   mbarriers (full / empty per stage, accfull / accempty per accumulator set); phases are separated by __syncthreads,
//   which is also what orders a phase's global writes before the next phase's reads (same CTA; operands are read with
//   ld.global.cg).
#include <atomic>
#include "tcmma.cuh"

namespace lqpb {

constexpr int kFuThreads = kPiv8Threads;             // 17 warps (-> 96 registers per thread: 5 warps share one scheduler's file)
constexpr int kFuEpi = 256;                          // epilogue group: warps 0-7 (TMEM lane quarter warp % 4, column half warp / 4)
constexpr int kFuRole = 256;                         // staging group: warps 8-15; warp 16: MMA issuer (look-ahead warp of the pivot inverse)
constexpr int kFuPrefetchAllNb = 4;                  // up to this many block rows a step's whole tile set is prefetched at once
constexpr int kFuCols = 512;                         // all of TMEM: two accumulator sets
constexpr int kFuSmem = 8 * kSlabBytes + kTBE * 4 + 1024;   // 2 operand stages + the epilogue tile, 1024-byte aligned

#ifdef LQPB_PHASE_TIMERS
// developer aid (tools/tc_fused_phases.py): CTA 0's timeline of its first problem in ns ([0] start, then per block step: pivot
// end, PANEL end, TRAIL end) and, from [32], the ns the roles of CTA 0 spent waiting: [32] staging on a free stage, [33] MMA
// issuer on a full stage, [34] MMA issuer on a drained accumulator, [35] epilogue on a finished accumulator
__device__ unsigned long long g_fu_ns[512];
__device__ __forceinline__ unsigned long long fu_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define FU_MARK(slot) do { if (blockIdx.x == 0 && threadIdx.x == 0 && b == 0) g_fu_ns[slot] = fu_now(); } while (0)
#define FU_TS(slot) do { if (blockIdx.x == 0 && st == 0 && c.k == 0 && (slot) < 512) g_fu_ns[slot] = fu_now(); } while (0)
#define FU_WAIT(slot, stmt) do { const unsigned long long t0__ = fu_now(); stmt; \
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) atomicAdd(&g_fu_ns[slot], fu_now() - t0__); } while (0)
#else
#define FU_MARK(slot)
#define FU_TS(slot)
#define FU_WAIT(slot, stmt) stmt
#endif

struct FuJob {
  const float* xsrc;
  const float* ysrc;
  float* vdst;      // PANEL: raw copy of X
  float* wdst;      // PANEL, i < k: W_i (row-major copy for TRAIL)
  float* mdst;      // PANEL: tile of M that receives W_i (transposed if xtrans) ; TRAIL: the C tile
  int xtrans;
};

struct FuCtx {
  unsigned char* sptr;     // operand stages (1024-byte aligned)
  float* obuf;             // epilogue tile
  uint64_t *full, *empty, *accfull, *accempty;
  uint32_t tmem;
  float *Mb, *Wb, *Vb;
  const float* Pk;
  int nb, k, ldl;
  int dbg;                 // developer switch (LQPB_FU_DBG): 1 staging does no work, 2 no MMAs, 4 epilogue does no work (timing only)
};

template <int MODE>
__device__ __forceinline__ FuJob fu_job(const FuCtx& c, int job) {
  const int k = c.k;
  int i, j = 0;
  if (MODE == 0) {
    i = c.ldl ? k + 1 + job : (job < k ? job : job + 1);
  } else {
    int ii = 0;
    while ((ii + 1) * (ii + 2) / 2 <= job) ++ii;
    const int jj = job - ii * (ii + 1) / 2;
    if (c.ldl) { i = k + 1 + ii; j = k + 1 + jj; }
    else { i = ii < k ? ii : ii + 1; j = jj < k ? jj : jj + 1; }
  }
  FuJob t;
  if (MODE == 0) {
    t.xtrans = i < k;
    t.mdst = c.Mb + (t.xtrans ? bl_tile(k, i) : bl_tile(i, k));
    t.xsrc = t.mdst;
    t.ysrc = c.Pk;
    t.vdst = c.Vb + (size_t)i * kTBE;
    t.wdst = c.Wb + (size_t)i * kTBE;
  } else {
    t.xtrans = 0;
    t.xsrc = i > k ? c.Mb + bl_tile(i, k) : c.Wb + (size_t)i * kTBE;   // W_i lives in M_ik below the pivot row
    t.ysrc = c.Vb + (size_t)j * kTBE;
    t.mdst = c.Mb + bl_tile(i, j);
    t.vdst = nullptr;
    t.wdst = nullptr;
  }
  return t;
}

// One phase (PANEL or TRAIL) of one block step of one problem: njobs tile products through the pipeline.  sc0 / jc0 =
// slabs / jobs that went through the pipeline before (they fix stage / accumulator indices and mbarrier parities).
template <int MODE>
__device__ __forceinline__ void fu_epilogue(const FuCtx& c, int njobs, uint32_t jc0) {
  int tid = threadIdx.x;
  asm volatile("" : "+r"(tid));          // opaque: see fu_opaque_tid
  const int ew = tid >> 5, lane = tid & 31;      // warps 0-7: TMEM lane quarter ew % 4, column half ew / 4
  const int r = 32 * (ew & 3) + lane;                               // tile row of pass 1
  const uint32_t trow = ((uint32_t)(32 * (ew & 3)) << 16);
  FuJob cur = fu_job<MODE>(c, 0);
#pragma unroll 1
  for (int job = 0; job < njobs; ++job) {
    const uint32_t jc = jc0 + job;
    const uint32_t buf = jc & 1u, v = jc >> 1;
    if (MODE == 1 && c.nb > kFuPrefetchAllNb && tid == 0 && job + 2 < njobs)     // C tile of the job after the next one -> L2
      l2_prefetch_bulk(fu_job<MODE>(c, job + 2).mdst, kTBE * 4);
    float4 creg[8];                    // TRAIL: batch 0 of the C tile travels while the job's MMAs run
    if (MODE == 1) {
#pragma unroll
      for (int t = 0; t < 8; ++t) creg[t] = ldcg_pinned(cur.mdst + (size_t)(ew + 8 * t) * kTB + 4 * lane);
    }
    FU_WAIT(35, mbar_wait_backoff(&c.accfull[buf], v & 1u, 64));
    tc_fence_after();
    const uint32_t tacc = c.tmem + buf * 256u + trow;
    // pass 1: this thread's row, its 64 columns in 4 groups of 16 (every tcgen05.ld / wait round trip costs ~0.3 us: loads
    // as fat as the register budget allows) -> swizzled tile
#pragma unroll 1
    for (int g = 0; g < ((c.dbg & 4) ? 0 : 4); ++g) {
      const int c0 = 64 * (ew >> 2) + 16 * g;
      uint32_t vh[16], vc[16];
      tmem_ld16(tacc + (uint32_t)c0, vh);
      tmem_ld16(tacc + 128u + (uint32_t)c0, vc);
      tmem_ld_wait(vh);
      tmem_ld_wait(vc);
#pragma unroll
      for (int e = 0; e < 16; ++e) vh[e] = __float_as_uint(__uint_as_float(vh[e]) + __uint_as_float(vc[e]));
      if (MODE == 0 && cur.xtrans) {
        // M_ki = W_i^T straight from the registers: for every column the 32 lanes (= 32 consecutive rows of W) write 128
        // contiguous bytes of row c0 + e of the transposed tile
#pragma unroll
        for (int e = 0; e < 16; ++e) cur.mdst[(size_t)(c0 + e) * kTB + r] = __uint_as_float(vh[e]);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int ch = (c0 >> 2) + q;
        *reinterpret_cast<float4*>(c.obuf + r * kTB + ((ch ^ (r & 31)) << 2)) =
            make_float4(__uint_as_float(vh[4 * q]), __uint_as_float(vh[4 * q + 1]), __uint_as_float(vh[4 * q + 2]),
                        __uint_as_float(vh[4 * q + 3]));
      }
    }
    tc_fence_before();                 // the TMEM reads are complete: the MMA warp may overwrite this accumulator set
    __syncwarp();
    if (lane == 0) mbar_arrive(&c.accempty[buf]);
    // TRAIL: the C tile, rows ew + 8 t, chunk `lane`, in 2 batches of 8 rows: in pass 2 every register is refilled with the
    // row of the next batch as soon as its row has been written
    if (c.dbg & 4) {
      bar_sync(1, kFuEpi);
      bar_sync(1, kFuEpi);
      if (job + 1 < njobs) cur = fu_job<MODE>(c, job + 1);
      continue;
    }
    bar_sync(1, kFuEpi);
    // pass 2: full rows, 512 bytes per warp access
    if (MODE == 1) {
#pragma unroll 1
      for (int bt = 0; bt < 2; ++bt) {
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const int rr = ew + 8 * (8 * bt + t);
          const float4 d = *reinterpret_cast<const float4*>(c.obuf + rr * kTB + ((lane ^ (rr & 31)) << 2));
          float4 o = creg[t];
          if (bt < 1) creg[t] = ldcg_pinned(cur.mdst + (size_t)(rr + 64) * kTB + 4 * lane);
          o.x -= d.x; o.y -= d.y; o.z -= d.z; o.w -= d.w;
          *reinterpret_cast<float4*>(cur.mdst + (size_t)rr * kTB + 4 * lane) = o;
        }
      }
    } else {
      float* dst = cur.xtrans ? cur.wdst : cur.mdst;
#pragma unroll 8
      for (int t = 0; t < 16; ++t) {
        const int rr = ew + 8 * t;
        const float4 d = *reinterpret_cast<const float4*>(c.obuf + rr * kTB + ((lane ^ (rr & 31)) << 2));
        *reinterpret_cast<float4*>(dst + (size_t)rr * kTB + 4 * lane) = d;
      }
    }
    bar_sync(1, kFuEpi);               // the tile is rewritten by pass 1 of the next job
    if (job + 1 < njobs) cur = fu_job<MODE>(c, job + 1);
  }
}

// Staging of one slab: split the registers of slab (jb, s) into the stage, refilling every register with the next slab's
// data (jn, sn) as soon as it has been consumed.  XT (PANEL jobs above the pivot row): X is the transpose of the stored tile.
template <int MODE, bool XT>
__device__ __forceinline__ void fu_stage_slab(unsigned char* sXh, const FuJob& jb, int s, const FuJob& jn, int sn, bool more,
                                              int st, int sw, int lane, float4 (&xr)[4], float4 (&yr)[4], int dbg = 0) {
  unsigned char* sXl = sXh + kSlabBytes;
  unsigned char* sYh = sXh + 2 * kSlabBytes;
  unsigned char* sYl = sXh + 3 * kSlabBytes;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int q = st + kFuRole * t, rr = q >> 3, ch = q & 7;
    const uint32_t off = (uint32_t)rr * 128u + (uint32_t)((ch ^ (rr & 7)) << 4);
    float4 hi, lo;
    const float4 xv = xr[t];
    if (more) {
      if (MODE == 0 && jn.xtrans)        // logical rows 4 (sw + 8 t) .. + 3 at logical column 32 sn + lane
        xr[t] = ldcg_pinned(jn.xsrc + (size_t)(32 * sn + lane) * kTB + 4 * (sw + 8 * t));
      else
        xr[t] = ldcg_pinned(jn.xsrc + (size_t)rr * kTB + 32 * sn + 4 * ch);
    }
    if (!XT) {
      split4(xv, hi, lo);
      *reinterpret_cast<float4*>(sXh + off) = hi;
      *reinterpret_cast<float4*>(sXl + off) = lo;
      if (MODE == 0 && !(dbg & 8)) *reinterpret_cast<float4*>(jb.vdst + (size_t)rr * kTB + 32 * s + 4 * ch) = xv;
    } else {
      const int r4 = 4 * (sw + 8 * t);
      const float xe[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r2 = r4 + e;
        const uint32_t o = (uint32_t)r2 * 128u + (uint32_t)((((lane >> 2) ^ (r2 & 7)) << 4) + ((lane & 3) << 2));
        float h1, l1;
        split_tf32(xe[e], h1, l1);
        *reinterpret_cast<float*>(sXh + o) = h1;
        *reinterpret_cast<float*>(sXl + o) = l1;
        if (MODE == 0) jb.vdst[(size_t)r2 * kTB + 32 * s + lane] = xe[e];
      }
    }
    const float4 yv = yr[t];
    if (more) yr[t] = ldcg_pinned(jn.ysrc + (size_t)rr * kTB + 32 * sn + 4 * ch);
    split4(yv, hi, lo);
    *reinterpret_cast<float4*>(sYh + off) = hi;
    *reinterpret_cast<float4*>(sYl + off) = lo;
  }
}

template <int MODE>
__device__ __forceinline__ void fu_stage(const FuCtx& c, int njobs, uint32_t sc0) {
  int tid = threadIdx.x;
  asm volatile("" : "+r"(tid));
  const int warp = tid >> 5, lane = tid & 31;
  const int st = tid - kFuRole, sw = warp - 8;
  // one slab of loads in flight per thread: every register is refilled with the next slab's data as soon as it has been
  // split and stored (the operands are L2 hits: written by this SM moments ago, or prefetched during the pivot inverse)
  float4 xr[4], yr[4];
  const int nsl = 4 * njobs;
  FuJob jcur = fu_job<MODE>(c, 0);
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int q = st + kFuRole * t, rr = q >> 3, ch = q & 7;
    if (MODE == 0 && jcur.xtrans) xr[t] = ldcg_pinned(jcur.xsrc + (size_t)lane * kTB + 4 * (sw + 8 * t));
    else xr[t] = ldcg_pinned(jcur.xsrc + (size_t)rr * kTB + 4 * ch);
    yr[t] = ldcg_pinned(jcur.ysrc + (size_t)rr * kTB + 4 * ch);
  }
#pragma unroll 1
  for (int L = 0; L < nsl; ++L) {
    const int s = L & 3, sn = (s + 1) & 3;
    const bool more = L + 1 < nsl;
    FuJob jnext = jcur;
    if (s == 3 && more) jnext = fu_job<MODE>(c, (L + 1) >> 2);
    const uint32_t sc = sc0 + (uint32_t)L;
    const uint32_t stage = sc & 1u, u = sc >> 1;
    FU_WAIT(32, mbar_wait_backoff(&c.empty[stage], (u & 1u) ^ 1u, 32));       // the MMAs that read this stage two slabs ago are done
    unsigned char* sXh = c.sptr + stage * (4 * kSlabBytes);
    if (c.dbg & 1) {}
    else if (MODE == 0 && jcur.xtrans) fu_stage_slab<MODE, true>(sXh, jcur, s, jnext, sn, more, st, sw, lane, xr, yr);
    else fu_stage_slab<MODE, false>(sXh, jcur, s, jnext, sn, more, st, sw, lane, xr, yr, c.dbg);
    fence_proxy_async();             // generic-proxy smem writes -> visible to the tensor core (async proxy)
    __syncwarp();
    if (lane == 0) mbar_arrive(&c.full[stage]);
    jcur = jnext;
  }
}

// The MMA issuer (one thread of warp 16).  tcgen05.mma issue is NOT free-running: with both operands in shared memory a
// 128 x 128 x 8 TF32 instruction occupies the tensor pipe for ~128 cycles (the operand fetch, 8 KB, runs at 64 B / cycle)
// and the issuing thread stalls for about that long per instruction (measured: 0.8 us per 12 instructions), so the issuer
// needs a warp of its own -- inside the staging group it serialised staging and MMAs.
template <int MODE>
__device__ __forceinline__ void fu_mma(const FuCtx& c, int njobs, uint32_t sc0, uint32_t jc0) {
  if ((threadIdx.x & 31) == 0) {
#pragma unroll 1
    for (int job = 0; job < njobs; ++job) {
      const uint32_t jc = jc0 + job;
      const uint32_t buf = jc & 1u, v = jc >> 1;
      FU_WAIT(34, mbar_wait(&c.accempty[buf], (v & 1u) ^ 1u));    // the epilogue has drained this accumulator set
      tc_fence_after();
      // The tensor core truncates the fp32 accumulator after every MMA, an error proportional to the accumulator's
      // magnitude: the two small cross terms get their own accumulator (columns +128), so that the large hi*hi sum sees
      // 16 instead of 48 roundings; the epilogue adds the two.
      const uint32_t tacc = c.tmem + buf * 256u, tcross = tacc + 128u;
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const uint32_t sc = sc0 + 4u * (uint32_t)job + (uint32_t)s;
        const uint32_t stage = sc & 1u, u = sc >> 1;
        FU_WAIT(33, mbar_wait(&c.full[stage], u & 1u));
        tc_fence_after();
        const uint32_t aXh = smem_u32(c.sptr) + stage * (4 * kSlabBytes);
        const uint32_t aXl = aXh + kSlabBytes, aYh = aXh + 2 * kSlabBytes, aYl = aXh + 3 * kSlabBytes;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          if (c.dbg & 2) break;
          const uint32_t ko = (uint32_t)kk * 32u;   // 8 tf32 = 32 bytes along K inside the swizzle atom
          const uint64_t dXh = umma_desc(aXh + ko), dXl = umma_desc(aXl + ko);
          const uint64_t dYh = umma_desc(aYh + ko), dYl = umma_desc(aYl + ko);
          umma_tf32(tcross, dXl, dYh, kIdescTf32, (s | kk) ? 1u : 0u);
          umma_tf32(tcross, dXh, dYl, kIdescTf32, 1u);
          umma_tf32(tacc, dXh, dYh, kIdescTf32, (s | kk) ? 1u : 0u);
        }
        umma_commit(&c.empty[stage]);
      }
      umma_commit(&c.accfull[buf]);
    }
  }
  __syncwarp();
}

template <int MODE>
__device__ __forceinline__ void fu_phase(const FuCtx& c, int njobs, uint32_t sc0, uint32_t jc0) {
  const int warp = threadIdx.x >> 5;
  if (warp < 8) fu_epilogue<MODE>(c, njobs, jc0);
  else if (warp < 16) fu_stage<MODE>(c, njobs, sc0);
  else fu_mma<MODE>(c, njobs, sc0, jc0);
}

__device__ __forceinline__ void fu_pivot(float* tile, float* P, float* scratch) { pivot8_body(tile, P, scratch); }

__global__ void __launch_bounds__(kFuThreads, 1) tc_fused_kernel(TcArgs a, int B) {
  extern __shared__ unsigned char fu_smem_raw[];
  __shared__ __align__(8) uint64_t bars[8];      // full[2] | empty[2] | accfull[2] | accempty[2]
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t sbase = (smem_u32(fu_smem_raw) + 1023u) & ~1023u;
  unsigned char* sptr = fu_smem_raw + (sbase - smem_u32(fu_smem_raw));
  if (warp == 0) tmem_alloc(&tmem_slot, kFuCols);
  if (tid == 32) {
    mbar_init(&bars[0], 8);      // one arrival per staging warp
    mbar_init(&bars[1], 8);
    mbar_init(&bars[2], 1);      // tcgen05.commit
    mbar_init(&bars[3], 1);
    mbar_init(&bars[4], 1);
    mbar_init(&bars[5], 1);
    mbar_init(&bars[6], 8);      // one arrival per epilogue warp
    mbar_init(&bars[7], 8);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // the per-step context lives in shared memory: nothing but (b, k, sc0, jc0) stays in registers across the pivot-block
  // inverse, whose inner loop needs the whole register budget (96 per thread with 17 warps)
  __shared__ FuCtx c;
  if (tid == 0) {
    c.sptr = sptr;
    c.obuf = reinterpret_cast<float*>(sptr + 8 * kSlabBytes);
    c.full = &bars[0];
    c.empty = &bars[2];
    c.accfull = &bars[4];
    c.accempty = &bars[6];
    c.tmem = tmem_slot;
    c.nb = a.nb;
    c.ldl = a.ldl;
    c.dbg = a.k;
  }
  const int nb = a.nb;
  const size_t ntile = (size_t)nb * (nb + 1) / 2;
  uint32_t sc0 = 0u, jc0 = 0u;
#pragma unroll 1
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    FU_MARK(0);
#pragma unroll 1
    for (int k = 0; k < nb; ++k) {
      if (tid == 0) {                    // read by the phases after the barrier below; the previous step's phases are over
        c.Mb = a.M + (size_t)b * ntile * kTBE;
        c.Wb = a.Wbuf + (size_t)b * nb * kTBE;
        c.Vb = a.Vbuf + (size_t)b * nb * kTBE;
        c.Pk = a.Pbuf + ((size_t)b * nb + k) * kTBE;
        c.k = k;
      }
      // While the pivot block is inverted (FP32 pipe, no memory traffic) the tiles the step is going to read travel to L2:
      // the batch's working set exceeds L2, and the roles keep only one slab (staging) / eight rows (epilogue) of loads in
      // flight, so a DRAM-latency load stalls them (measured: 9 us per tile job, all of it load latency).  Forward: the
      // PANEL operands M_ik and the TRAIL C tiles; larger sweeps prefetch the C tiles job by job (fu_epilogue).
      if (tid == 7 * 32) {
        const float* Mp = a.M + (size_t)b * ntile * kTBE;
        for (int I = 0; I < nb; ++I)
          for (int J = 0; J <= I; ++J) {
            const bool panel = a.ldl ? (J == k && I > k) : ((J == k) != (I == k));
            const bool trail = a.ldl ? (J > k) : (I != k && J != k);
            if (panel || (trail && nb <= kFuPrefetchAllNb)) l2_prefetch_bulk(Mp + bl_tile(I, J), kTBE * 4);
          }
      }
      fu_pivot(a.M + (size_t)b * ntile * kTBE + bl_tile(k, k), a.Pbuf + ((size_t)b * nb + k) * kTBE,
                                   reinterpret_cast<float*>(sptr));   // scratch = operand stage 0 (the pipeline is drained)
      __syncthreads();
      FU_MARK(1 + 3 * k);
      const int span = a.ldl ? nb - 1 - k : nb - 1;
      if (span <= 0) continue;
      fu_phase<0>(c, span, sc0, jc0);
      sc0 += 4u * (uint32_t)span;
      jc0 += (uint32_t)span;
      __syncthreads();
      FU_MARK(2 + 3 * k);
      const int nt = span * (span + 1) / 2;
      fu_phase<1>(c, nt, sc0, jc0);
      sc0 += 4u * (uint32_t)nt;
      jc0 += (uint32_t)nt;
      __syncthreads();
      FU_MARK(3 + 3 * k);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_slot, kFuCols);
}

cudaError_t launch_tc_fused(int B, const TcArgs& a, cudaStream_t st) {
  // function attributes and the SM count are PER DEVICE (a process may drive several GPUs, and autograd calls in from its
  // own thread): one slot per device ordinal, published with release / acquire
  static std::atomic<int> sm_of_dev[64];
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const bool slot_ok = dev >= 0 && dev < 64;
  int n_sm = slot_ok ? sm_of_dev[dev].load(std::memory_order_acquire) : 0;
  if (n_sm == 0) {
    e = cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(tc_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFuSmem);
    if (e != cudaSuccess) return e;
    if (slot_ok) sm_of_dev[dev].store(n_sm, std::memory_order_release);
  }
  tc_fused_kernel<<<B < n_sm ? B : n_sm, kFuThreads, kFuSmem, st>>>(a, B);
  return cudaGetLastError();
}

}  // namespace lqpb

#ifdef LQPB_PHASE_TIMERS
extern "C" void lqpb_debug_fu_ns(unsigned long long* out, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, lqpb::g_fu_ns, sizeof(unsigned long long) * 512);
  if (reset) {
    unsigned long long z[512] = {0};
    cudaMemcpyToSymbol(lqpb::g_fu_ns, z, sizeof(z));
  }
}
#endif
