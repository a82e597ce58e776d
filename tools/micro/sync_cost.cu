// Developer microbenchmark (GPU box): cost of the synchronisation primitives the iteration kernels use at a stop check.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/sync_cost tools/micro/sync_cost.cu && gpurun_out/sync_cost
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned ld_acq(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_rlx(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// mode 0: __threadfence ; 1: fence.acq_rel.gpu ; 2: red.add + __threadfence ; 3: ld.acquire.gpu ; 4: ld.relaxed.gpu ;
// 5: atom.add (returning) ; 6: __syncthreads ; 7: st + fence.acq_rel.gpu ; 8: red.release.gpu
__global__ void single(int mode, int reps, unsigned* g, double* out) {
  __shared__ unsigned long long t0;
  unsigned acc = 0;
  if (threadIdx.x == 0) t0 = gtime();
  __syncthreads();
  for (int r = 0; r < reps; ++r) {
    if (mode == 6) { __syncthreads(); continue; }
    if (threadIdx.x == 0) {
      switch (mode) {
        case 0: __threadfence(); break;
        case 1: asm volatile("fence.acq_rel.gpu;" ::: "memory"); break;
        case 2: atomicAdd(&g[0], 1u); __threadfence(); break;
        case 3: acc += ld_acq(&g[r & 7]); break;
        case 4: acc += ld_rlx(&g[r & 7]); break;
        case 5: acc += atomicAdd(&g[1], 1u); break;
        case 7: g[2] = r; asm volatile("fence.acq_rel.gpu;" ::: "memory"); break;
        case 8: asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(&g[3]) : "memory"); break;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) { out[mode] = double(gtime() - t0) / reps; if (acc == 12345) g[9] = acc; }
}

// grid barrier variants over gridDim.x CTAs, reps times. mode 0: atomicAdd + spin ld.acquire (as iterate.cu) with
// __threadfence both sides; 1: red.release + spin ld.acquire, no extra fences; 2: cooperative groups grid.sync()
__global__ void gridbar(int mode, int reps, unsigned* bar, double* out) {
  cg::grid_group grid = cg::this_grid();
  __shared__ unsigned long long t0;
  if (threadIdx.x == 0) t0 = gtime();
  __syncthreads();
  for (int r = 0; r < reps; ++r) {
    if (mode == 2) { grid.sync(); continue; }
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned target = (unsigned)(r + 1) * gridDim.x;
      if (mode == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        while (ld_acq(bar) < target) {}
        __threadfence();
      } else {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
        while (ld_acq(bar) < target) {}
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = double(gtime() - t0) / reps;
}

__global__ void clusterbar(int reps, double* out) {
  __shared__ unsigned long long t0;
  if (threadIdx.x == 0) t0 = gtime();
  __syncthreads();
  for (int r = 0; r < reps; ++r) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = double(gtime() - t0) / reps;
}

int main() {
  unsigned* g; double* out;
  cudaMalloc(&g, 4096); cudaMemset(g, 0, 4096);
  cudaMallocManaged(&out, 64 * sizeof(double));
  const char* names[] = {"__threadfence", "fence.acq_rel.gpu", "red.add + __threadfence", "ld.acquire.gpu", "ld.relaxed.gpu",
                         "atom.add returning", "__syncthreads (512 thr)", "st + fence.acq_rel.gpu", "red.release.gpu"};
  for (int m = 0; m < 9; ++m) {
    single<<<1, 512>>>(m, 2000, g, out);
    cudaDeviceSynchronize();
    printf("%-28s %8.1f ns\n", names[m], out[m]);
  }
  for (int G : {1, 8, 32, 64, 128, 148})
    for (int mode = 0; mode < 3; ++mode) {
      cudaMemset(g, 0, 4096);
      int reps = 1000; double* o = out + 32; unsigned* bar = g;
      void* args[] = {&mode, &reps, &bar, &o};
      cudaError_t e = cudaLaunchCooperativeKernel((void*)gridbar, dim3(G), dim3(512), args, 0, 0);
      cudaDeviceSynchronize();
      printf("grid barrier G=%3d mode %d (%s): %8.1f ns %s\n", G, mode,
             mode == 0 ? "fence+atomicAdd+spin+fence" : mode == 1 ? "red.release+spin acquire" : "cg grid.sync", out[32],
             e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  for (int G : {1, 2, 4, 8}) {
    cudaLaunchConfig_t lc{};
    lc.gridDim = dim3(G); lc.blockDim = dim3(512);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = G; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    lc.attrs = at; lc.numAttrs = 1;
    int reps = 1000; double* o = out + 40;
    cudaError_t e = cudaLaunchKernelEx(&lc, clusterbar, reps, o);
    cudaDeviceSynchronize();
    printf("cluster barrier size %d: %8.1f ns %s\n", G, out[40], e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
