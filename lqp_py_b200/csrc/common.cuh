// Shared device helpers for the sm_100a kernels: mbarrier / bulk-TMA PTX wrappers,
// warp and block reductions, 16-byte vector types, and the workspace layout.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace lqpb {

constexpr int kWarp = 32;

template <typename T> struct Vec;   // 16-byte vector of T
template <> struct Vec<float>  { using type = float4;  static constexpr int N = 4; };
template <> struct Vec<double> { using type = double2; static constexpr int N = 2; };

__host__ __device__ inline int round_up(int a, int b) { return (a + b - 1) / b * b; }
__host__ __device__ inline size_t round_up_sz(size_t a, size_t b) { return (a + b - 1) / b * b; }

template <typename T> __device__ __forceinline__ T t_abs(T v) { return v < T(0) ? -v : v; }
template <typename T> __device__ __forceinline__ T t_max(T a, T b) { return a > b ? a : b; }
template <typename T> __device__ __forceinline__ T t_min(T a, T b) { return a < b ? a : b; }
__device__ __forceinline__ float t_sqrt(float v) { return sqrtf(v); }
__device__ __forceinline__ double t_sqrt(double v) { return sqrt(v); }
template <typename T> __device__ __forceinline__ T t_inf();
template <> __device__ __forceinline__ float t_inf<float>() { return __int_as_float(0x7f800000); }
template <> __device__ __forceinline__ double t_inf<double>() { return __longlong_as_double(0x7ff0000000000000LL); }

// packed FP32 FMA (Blackwell FFMA2: two IEEE fused multiply-adds per issued instruction), d = a * b + c per half
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                     rc = *reinterpret_cast<unsigned long long*>(&c), rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}

// ---------------------------------------------------------------- warp / block reductions
template <typename T> __device__ __forceinline__ T warp_max(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = t_max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// named barrier over `count` threads (count multiple of 32); id 0 is __syncthreads
__device__ __forceinline__ void bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// Block-wide max over `nthreads` threads (the calling group, synchronised through named
// barrier `bar`); `scratch` holds >= 32 T.  All threads get the result.
template <typename T>
__device__ __forceinline__ T group_max(T v, T* scratch, int tid, int nthreads, int bar) {
  v = warp_max(v);
  const int w = tid >> 5, nw = nthreads >> 5;
  bar_sync(bar, nthreads);                 // protect scratch from the previous use
  if ((tid & 31) == 0) scratch[w] = v;
  bar_sync(bar, nthreads);
  T r = scratch[0];
  for (int k = 1; k < nw; ++k) r = t_max(r, scratch[k]);
  return r;
}
template <typename T>
__device__ __forceinline__ T group_sum(T v, T* scratch, int tid, int nthreads, int bar) {
  v = warp_sum(v);
  const int w = tid >> 5, nw = nthreads >> 5;
  bar_sync(bar, nthreads);
  if ((tid & 31) == 0) scratch[w] = v;
  bar_sync(bar, nthreads);
  T r = scratch[0];
  for (int k = 1; k < nw; ++k) r += scratch[k];
  return r;
}

// ---------------------------------------------------------------- mbarrier + bulk TMA (PTX)
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(
                   smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// the same for a whole group of waiting warps: back off between polls, so that the pollers do not take issue slots and
// shared-memory bandwidth from the warps that are working on the same SM
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity, unsigned ns) {
  while (!mbar_try_wait(bar, parity)) __nanosleep(ns);
}
// 1-D bulk async copy global -> shared (TMA engine, SASS UBLKCP), completion on an mbarrier.
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// same with an L2 cache-policy hint
__device__ __forceinline__ void tma_load_1d_hint(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar,
                                                 uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
          "r"(smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}

// ---------------------------------------------------------------- global-memory helpers
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// atomic max for non-negative floating point values stored as T
__device__ __forceinline__ void atomic_max_nonneg(float* addr, float v) {
  atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
}
__device__ __forceinline__ void atomic_max_nonneg(double* addr, double v) {
  atomicMax(reinterpret_cast<long long*>(addr), __double_as_longlong(v));
}

// ---------------------------------------------------------------- control block (device ints)
// One per forward call, zeroed by the host before the first launch.
struct Ctrl {
  unsigned barrier;        // monotonically increasing grid-barrier counter
  int status;              // 0 running, 1 converged, 2 max_iters, 3 refactor requested
  int iter;                // loop index at exit
  int next_i;              // first iteration of the next segment (status 3)
  int any_lb, any_ub;      // OR over the batch (set by the scale kernel)
  int n_log;
  int pad0;
  int slot[4][4];          // per-check reduction slots: [not_optimal, any_wants_update, any_ratio_out, unused]
  int last_wants, last_ratio_out;  // flags of the most recent check (survive a relaunch)
  int pad1[2];
  int log_iter[64];
  double log_primal[64];
  double log_dual[64];
};

}  // namespace lqpb
