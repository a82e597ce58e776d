// Small transfers between device memory and page-locked (mapped) host memory done by a KERNEL instead of the copy
// engines.  A copy engine finishes the transfer it has started before it looks at the next one, so a 256 KB result or a
// control block queued while a 128 MB batch is on its way waits for the whole batch; loads / stores issued by the SMs go
// over PCIe beside the DMA traffic.  Used where the host-buffer pipeline keeps both engines busy with bulk transfers
// (solve-ahead: DESIGN.md 4a): the control block at the end of a forward segment, x of a solved-ahead batch, dl_dz.
#include "common.cuh"
#include "../../include/lqpb.h"

namespace lqpb {

__global__ void __launch_bounds__(256) mapped_copy_kernel(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src,
                                                          size_t n4, int vec16) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (vec16) {
    const size_t n16 = n4 / 4;
    uint4* d = reinterpret_cast<uint4*>(dst);
    const uint4* s = reinterpret_cast<const uint4*>(src);
    for (size_t k = i; k < n16; k += stride) d[k] = s[k];
    for (size_t k = n16 * 4 + i; k < n4; k += stride) dst[k] = src[k];
  } else {
    for (; i < n4; i += stride) dst[i] = src[i];
  }
  __threadfence_system();
}

cudaError_t launch_mapped_copy(void* dst, const void* src, size_t bytes, cudaStream_t st) {
  if (bytes == 0) return cudaSuccess;
  if (bytes % 4 != 0 || ((uintptr_t)dst | (uintptr_t)src) % 4 != 0) return cudaErrorInvalidValue;
  const size_t n4 = bytes / 4;
  const int vec16 = (((uintptr_t)dst | (uintptr_t)src) % 16 == 0) ? 1 : 0;
  const size_t items = vec16 ? (n4 + 3) / 4 : n4;
  int blocks = (int)((items + 255) / 256);
  if (blocks > 64) blocks = 64;       // few CTAs: this runs beside the solve's kernels and must not take SMs from them
  if (blocks < 1) blocks = 1;
  mapped_copy_kernel<<<blocks, 256, 0, st>>>((uint32_t*)dst, (const uint32_t*)src, n4, vec16);
  return cudaGetLastError();
}

}  // namespace lqpb

extern "C" int lqpb_copy_mapped(void* dst, const void* src, size_t bytes, void* stream) {
  if (!dst || !src) return LQPB_E_ARG;
  return lqpb::launch_mapped_copy(dst, src, bytes, (cudaStream_t)stream) == cudaSuccess ? LQPB_OK : LQPB_E_CUDA;
}
