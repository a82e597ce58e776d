// Shared pieces of the tensor-core factorisation kernels (tcfactor.cu: one launch per sweep phase; tcfused.cu: the whole
// sweep of a problem in one persistent CTA): tile geometry, tcgen05 / TMEM PTX wrappers, the TF32 hi / lo split and the
// 128 x 128 pivot-block inverse (FP32 pipe, 16 rank-8 steps with a look-ahead warp).
#pragma once
#include "layout.cuh"

namespace lqpb {

constexpr int kTB = 128;                 // block edge
constexpr int kTBE = kTB * kTB;          // elements per tile
constexpr int kSlabBytes = kTB * 128;    // 128 rows x 32 fp32
constexpr int kTcThreads = 512;
constexpr int kTcCols = 256;             // TMEM columns per CTA: hi*hi accumulator | cross-term accumulator
constexpr int kTcSmem = 8 * kSlabBytes + 1024;   // two stages of Xhi | Xlo | Yhi | Ylo, manually aligned to 1024 B
constexpr int kTcSmemTrail = kTcSmem + kTBE * 4; // + the C tile of the job, fetched by one bulk TMA copy

__host__ __device__ inline size_t bl_tile(int I, int J) { return (size_t)(I * (I + 1) / 2 + J) * kTBE; }
__host__ __device__ inline size_t bl_off(int i, int j) {   // element (i, j), tile row >= tile column
  return bl_tile(i >> 7, j >> 7) + (size_t)(i & 127) * kTB + (j & 127);
}

struct TcArgs {
  float* M;       // B * nb(nb+1)/2 tiles
  float* Wbuf;    // B * nb tiles : W_i of the current step
  float* Vbuf;    // B * nb tiles : V_i = M_ik before the step
  float* Pbuf;    // B * nb tiles : P_k = inv(M_kk) (all kept: the LDL solve needs them)
  int nb, k, ldl;
  int acc2;       // 1: cross terms (lo*hi + hi*lo) accumulate in their own TMEM tile (see tc_tile_kernel)
};

// ------------------------------------------------------------------ tcgen05 PTX wrappers
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {   // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {     // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, 128 x 128 x 8, TF32 inputs, FP32 accumulate; issued by ONE thread
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrive when every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 16 consecutive columns: thread `lane` of warp w reads TMEM lane 32 (w % 4) + lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
// wait for the tcgen05.ld; the registers are in/out operands so that no use of them is scheduled above the wait
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&v)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
               :
               : "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4 in
// bits [0,14), leading byte offset (unused for swizzled K-major, canonical value 1) in [16,30), stride byte
// offset = 1024 B between 8-row groups in [32,46), descriptor version 1 in [46,48), layout type 2 in [61,64).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2),
// both K-major (bits 15, 16 = 0), N >> 3 in bits [17,23), M >> 4 in bits [24,29)
constexpr uint32_t kIdescTf32 = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

// the whole block sweep (forward inverse or LDL^T, a.ldl) of B problems in one launch, CTA b owns problem b (tcfused.cu)
cudaError_t launch_tc_fused(int B, const TcArgs& a, cudaStream_t st);

// asynchronous prefetch of `bytes` (multiple of 16) contiguous bytes into L2: one instruction, no registers, no completion
__device__ __forceinline__ void l2_prefetch_bulk(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// 16-byte global load through L2 that stays where it is written (volatile asm, memory clobber): the pipelined roles of the
// fused kernel place their prefetches by hand, and ptxas otherwise hoists whole batches of them and spills
__device__ __forceinline__ float4 ldcg_pinned(const float* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

// hi = x rounded to TF32 (10-bit mantissa, round to nearest / ties away: add half an ulp to the magnitude and
// clear the 13 low bits -- 2 integer ops instead of cvt.rna.tf32.f32), lo = x - hi (exact, |lo| <= 2^-11 |x|)
// rounded to TF32 the same way: the neglected part is <= 2^-22 |x|.
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
  lo = __uint_as_float((__float_as_uint(x - hi) + 0x1000u) & 0xffffe000u);
}
__device__ __forceinline__ void split4(const float4& x, float4& hi, float4& lo) {
  split_tf32(x.x, hi.x, lo.x);
  split_tf32(x.y, hi.y, lo.y);
  split_tf32(x.z, hi.z, lo.z);
  split_tf32(x.w, hi.w, lo.w);
}

// ------------------------------------------------------------------ pivot block inverse, blocked by 8
// Same register tiling as tc_pivot_kernel (16 work warps, thread (ta, tb) owns rows 4 ta.., columns 8 tb..), but
// the 128 pivots are swept 8 at a time: 16 rank-8 updates instead of 128 barrier-separated rank-1 updates.  With S
// the 8 pivot indices of a step and R the rest, the sweep of S is
//     P = inv(A_SS),  U = P A_S:,  A_RR -= U_R^T A_SR  (A_RS = A_SR^T),  A_SR <- U,  A_RS <- U^T,  A_SS <- -P.
// The 8 x 8 inverse is a serial chain of 8 reciprocals; a 17th "pivot warp" takes it off the critical path by
// look-ahead: as soon as U of step sb is known it forms the NEXT diagonal block
//     D' = A_S'S' - U_S'^T A_SS'           (A_S'S' published raw by its owner warp at the top of the step)
// and inverts it (two entries per lane, pivot row / column exchanged with shuffles) while the 16 work warps
// apply the rank-8 update of step sb.
//   A. warp sb publishes its rows (= rows S, final since the previous update); warp sb + 1 publishes raw A_S'S'
//   B. work warps: U = P A_S: (two entries per thread)
//   C. work warps: rank-8 update of the 4 x 8 tiles, owners of rows / columns S overwrite them with U / U^T / -P;
//      pivot warp: P' = inv(D')
constexpr int kPivThreads = 512;
constexpr int kPiv8Threads = kPivThreads + 32;

// lane L holds entries (i, j0) and (i, j0 + 1) of an 8 x 8 symmetric block, i = L / 4, j0 = 2 (L % 4); on exit the
// block is inv(block).  Same sweep formulas as the rank-1 kernel; the sign flip at the end gives +inverse.
__device__ __forceinline__ void inv8x8_warp(float& e0, float& e1, int lane) {
  const int i = lane >> 2, jq = lane & 3;
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    const float d_s0 = __shfl_sync(0xffffffffu, e0, s * 4 + jq);       // row s, my two columns
    const float d_s1 = __shfl_sync(0xffffffffu, e1, s * 4 + jq);
    const float d_isa = __shfl_sync(0xffffffffu, e0, i * 4 + (s >> 1)); // my row, column s
    const float d_isb = __shfl_sync(0xffffffffu, e1, i * 4 + (s >> 1));
    const float d_ssa = __shfl_sync(0xffffffffu, e0, s * 4 + (s >> 1));
    const float d_ssb = __shfl_sync(0xffffffffu, e1, s * 4 + (s >> 1));
    const float d_is = (s & 1) ? d_isb : d_isa;
    const float piv = __frcp_rn((s & 1) ? d_ssb : d_ssa);
    const float ci = d_is * piv;
    const int j0 = 2 * jq;
    if (i != s) {
      e0 = (j0 == s) ? ci : fmaf(-ci, d_s0, e0);
      e1 = (j0 + 1 == s) ? ci : fmaf(-ci, d_s1, e1);
    } else {
      e0 = (j0 == s) ? -piv : d_s0 * piv;
      e1 = (j0 + 1 == s) ? -piv : d_s1 * piv;
    }
  }
  e0 = -e0;
  e1 = -e1;
}

// All kPiv8Threads threads of the CTA call this (it synchronises with __syncthreads): `tile` = M_kk (lower triangle is the
// reference copy) ends as -(M_kk)^-1, P receives +(M_kk)^-1.  Two barriers per step: the owners publish the rows of step
// sb + 1 (final once their update of step sb is done) and the raw diagonal block of step sb + 2 at the END of step sb, into
// the other half of double-buffered arrays.  `scratch`: kPivScratchFloats floats of shared memory.
constexpr int kPivScratchFloats = 4 * 8 * kTB + 4 * 64;
__device__ __forceinline__ void pivot8_body(float* tile, float* P, float* scratch) {
  float(*rowbuf)[8][kTB] = reinterpret_cast<float(*)[8][kTB]>(scratch);             // [2][8][128] rows S of step parity
  float(*ubuf)[kTB] = reinterpret_cast<float(*)[kTB]>(scratch + 2 * 8 * kTB);         // [8][128]    U = P A_S:
  float(*nubuf)[kTB] = reinterpret_cast<float(*)[kTB]>(scratch + 3 * 8 * kTB);        // [8][128]    -U (no negations in the hot loop)
  float(*pbuf)[8][8] = reinterpret_cast<float(*)[8][8]>(scratch + 4 * 8 * kTB);       // [2][8][8]   P of step parity
  float(*dbuf)[8][8] = reinterpret_cast<float(*)[8][8]>(scratch + 4 * 8 * kTB + 128); // [2][8][8]   raw diagonal block
  // work thread (ta, tb): rows 4 ta .. 4 ta + 3, columns 4 tb .. 4 tb + 3 (acc[.][0..3]) and 64 + 4 tb .. + 3
  // (acc[.][4..7]): the 16 threads of a row group read 256 contiguous bytes of a published row per access
  const int tid = threadIdx.x, ta = (tid >> 4) & 31, tb = tid & 15;
  const int warp = tid >> 5, lane = tid & 31;
  const bool pivot_warp = warp == kPivThreads / 32;
  // defined on EVERY path (the tile is only written and read by the work warps; a compiler that cannot correlate the
  // predicates would otherwise keep it alive across the caller's loops)
  float acc[4][8];
#pragma unroll
  for (int rr = 0; rr < 4; ++rr)
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[rr][q] = 0.f;
  // publish the diagonal block `blk` (8 x 8, held by the four threads ta = 2 blk, 2 blk + 1, tb = tB, tB + 1 of warp blk)
  auto publish_diag = [&](int blk, float(*dst)[8]) {
    const int tB = 2 * (blk & 7), gB = blk >> 3, h = ta & 1;
    if (warp == blk && (tb == tB || tb == tB + 1)) {
#pragma unroll
      for (int rr = 0; rr < 4; ++rr)
        *reinterpret_cast<float4*>(&dst[4 * h + rr][4 * (tb - tB)]) =
            gB ? make_float4(acc[rr][4], acc[rr][5], acc[rr][6], acc[rr][7])
               : make_float4(acc[rr][0], acc[rr][1], acc[rr][2], acc[rr][3]);
    }
  };
  auto publish_rows = [&](int blk, float(*dst)[kTB]) {          // rows 8 blk .. 8 blk + 7 live in warp blk
    if (warp == blk) {
      const int h = ta & 1;
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
        *reinterpret_cast<float4*>(&dst[4 * h + rr][4 * tb]) = make_float4(acc[rr][0], acc[rr][1], acc[rr][2], acc[rr][3]);
        *reinterpret_cast<float4*>(&dst[4 * h + rr][64 + 4 * tb]) = make_float4(acc[rr][4], acc[rr][5], acc[rr][6], acc[rr][7]);
      }
    }
  };
  if (!pivot_warp) {
    // the lower triangle is the reference copy: entries above the diagonal are read transposed
#pragma unroll
    for (int rr = 0; rr < 4; ++rr)
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int r = 4 * ta + rr, c = (q < 4 ? 4 * tb + q : 60 + 4 * tb + q);
        acc[rr][q] = __ldcg(r >= c ? tile + (size_t)r * kTB + c : tile + (size_t)c * kTB + r);
      }
    publish_rows(0, rowbuf[0]);
    publish_diag(0, dbuf[0]);           // first inverse
    publish_diag(1, dbuf[1]);           // look-ahead of step 0
  }
  __syncthreads();
  if (pivot_warp) {
    const int i = lane >> 2, j0 = 2 * (lane & 3);
    float e0 = dbuf[0][i][j0], e1 = dbuf[0][i][j0 + 1];
    inv8x8_warp(e0, e1, lane);
    *reinterpret_cast<float2*>(&pbuf[0][i][j0]) = make_float2(e0, e1);
  }
  __syncthreads();

#pragma unroll 1
  for (int sb = 0; sb < kTB / 8; ++sb) {
    const int cur = sb & 1;
    const float(*Pc)[8] = pbuf[cur];
    const float(*rows)[kTB] = rowbuf[cur];
    // columns S = 8 sb .. 8 sb + 7 belong to threads tb = tS, tS + 1 (4 columns each), register half gS
    const int tS = 2 * (sb & 7), gS = sb >> 3;
    // ---- B. U = P A_S: (8 x 128): thread -> column c = tid % 128, rows s = tid / 128 and tid / 128 + 4
    if (!pivot_warp) {
      const int c = tid & 127, s0 = tid >> 7;
      float rv[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) rv[t] = rows[t][c];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int s = s0 + 4 * hh;
        const float4 p0 = *reinterpret_cast<const float4*>(&Pc[s][0]);
        const float4 p1 = *reinterpret_cast<const float4*>(&Pc[s][4]);
        float u = p0.x * rv[0];
        u = fmaf(p0.y, rv[1], u); u = fmaf(p0.z, rv[2], u); u = fmaf(p0.w, rv[3], u);
        u = fmaf(p1.x, rv[4], u); u = fmaf(p1.y, rv[5], u); u = fmaf(p1.z, rv[6], u); u = fmaf(p1.w, rv[7], u);
        ubuf[s][c] = u;
        nubuf[s][c] = -u;
      }
    }
    __syncthreads();
    if (pivot_warp) {
      // ---- C'. look-ahead: D' = A_S'S' - U_S'^T A_SS', P' = inv(D')
      if (sb + 1 < kTB / 8) {
        const int i = lane >> 2, j0 = 2 * (lane & 3), c0 = 8 * (sb + 1);
        float e0 = dbuf[cur ^ 1][i][j0], e1 = dbuf[cur ^ 1][i][j0 + 1];
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          const float ui = ubuf[s][c0 + i];
          e0 = fmaf(-ui, rows[s][c0 + j0], e0);
          e1 = fmaf(-ui, rows[s][c0 + j0 + 1], e1);
        }
        inv8x8_warp(e0, e1, lane);
        *reinterpret_cast<float2*>(&pbuf[cur ^ 1][i][j0]) = make_float2(e0, e1);
      }
    } else {
      // ---- C. rank-8 update of the thread's tile:  A_rc -= sum_s U[s][r] * A_S[s][c]
#pragma unroll 4
      for (int s = 0; s < 8; ++s) {
        const float4 nu4 = *reinterpret_cast<const float4*>(&nubuf[s][4 * ta]);     // -U[s][rows of the thread]
        const float4 r0 = *reinterpret_cast<const float4*>(&rows[s][4 * tb]);
        const float4 r1 = *reinterpret_cast<const float4*>(&rows[s][64 + 4 * tb]);
        const float nuu[4] = {nu4.x, nu4.y, nu4.z, nu4.w};
        const float2 rc[4] = {make_float2(r0.x, r0.y), make_float2(r0.z, r0.w), make_float2(r1.x, r1.y),
                              make_float2(r1.z, r1.w)};
        // packed FFMA2: two of the thread's 32 entries per issued instruction (the update is issue-bound)
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          const float2 nu = make_float2(nuu[rr], nuu[rr]);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 t = ffma2(nu, rc[j], make_float2(acc[rr][2 * j], acc[rr][2 * j + 1]));
            acc[rr][2 * j] = t.x;
            acc[rr][2 * j + 1] = t.y;
          }
        }
      }
      if (tb == tS || tb == tS + 1) {     // columns S of the thread's rows:  A_rS <- U^T
        const int s0 = tb != tS ? 4 : 0;  // local columns 4..7 of S belong to thread tS + 1
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 v = *reinterpret_cast<const float4*>(&ubuf[s0 + j][4 * ta]);
          if (gS) { acc[0][4 + j] = v.x; acc[1][4 + j] = v.y; acc[2][4 + j] = v.z; acc[3][4 + j] = v.w; }
          else    { acc[0][j] = v.x; acc[1][j] = v.y; acc[2][j] = v.z; acc[3][j] = v.w; }
        }
      }
      if (warp == sb) {          // rows S:  A_Sc <- U, and A_SS <- -P
        const int h = ta & 1;
        const bool own = tb == tS || tb == tS + 1;
        const int lo = 4 * (tb - tS);
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          float4 u0 = *reinterpret_cast<const float4*>(&ubuf[4 * h + rr][4 * tb]);
          float4 u1 = *reinterpret_cast<const float4*>(&ubuf[4 * h + rr][64 + 4 * tb]);
          if (own) {
            const float4 p = *reinterpret_cast<const float4*>(&Pc[4 * h + rr][lo]);
            const float4 np4 = make_float4(-p.x, -p.y, -p.z, -p.w);
            if (gS) u1 = np4; else u0 = np4;
          }
          acc[rr][0] = u0.x; acc[rr][1] = u0.y; acc[rr][2] = u0.z; acc[rr][3] = u0.w;
          acc[rr][4] = u1.x; acc[rr][5] = u1.y; acc[rr][6] = u1.z; acc[rr][7] = u1.w;
        }
      }
      // ---- publish for the next steps (the other halves of the double buffers; last read one step ago)
      publish_rows(sb + 1, rowbuf[cur ^ 1]);
      publish_diag(sb + 2, dbuf[cur]);
    }
    __syncthreads();
  }
  if (pivot_warp) return;
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) {
    const size_t o = (size_t)(4 * ta + rr) * kTB + 4 * tb;
    *reinterpret_cast<float4*>(tile + o) = make_float4(acc[rr][0], acc[rr][1], acc[rr][2], acc[rr][3]);
    *reinterpret_cast<float4*>(tile + o + 64) = make_float4(acc[rr][4], acc[rr][5], acc[rr][6], acc[rr][7]);
    *reinterpret_cast<float4*>(P + o) = make_float4(-acc[rr][0], -acc[rr][1], -acc[rr][2], -acc[rr][3]);
    *reinterpret_cast<float4*>(P + o + 64) = make_float4(-acc[rr][4], -acc[rr][5], -acc[rr][6], -acc[rr][7]);
  }
}

}  // namespace lqpb
