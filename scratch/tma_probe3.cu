#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// V: bit0 = my TMA asm, bit1 = my init+fence, bit2 = my expect_tx (before TMA) + my wait loop
template <int V>
__global__ void kernel(const __grid_constant__ CUtensorMap tensor_map, int x, int y, int* out) {
  __shared__ alignas(128) int smem_buffer[32][32];
  __shared__ alignas(8) uint64_t rawbar;
#pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ barrier bar;
  if (V & 4) {
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&rawbar)), "r"(1));
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&rawbar)), "r"(4096) : "memory");
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(smem_u32(smem_buffer)), "l"(reinterpret_cast<uint64_t>(&tensor_map)), "r"(smem_u32(&rawbar)), "r"(x), "r"(y) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&rawbar)), "r"(0) : "memory");
  } else {
    if (threadIdx.x == 0) {
      init(&bar, blockDim.x);
      cde::fence_proxy_async_shared_cta();
    }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
      if (V & 1) {
        uint64_t* nb = cuda::device::barrier_native_handle(bar);
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(smem_u32(smem_buffer)), "l"(reinterpret_cast<uint64_t>(&tensor_map)), "r"(smem_u32(nb)), "r"(x), "r"(y) : "memory");
      } else {
        cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
      }
      token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
    } else {
      token = bar.arrive();
    }
    bar.wait(std::move(token));
  }
  out[threadIdx.x] = smem_buffer[0][threadIdx.x % 32];
}

typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  int V = argc > 1 ? atoi(argv[1]) : 0; int X0 = argc > 2 ? atoi(argv[2]) : 64;
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  PFN enc = (PFN)p;
  const int W = 256, Hh = 256;
  std::vector<int> h(W * Hh);
  for (int i = 0; i < W * Hh; ++i) h[i] = i;
  int *d, *out; cudaMalloc(&d, W * Hh * 4); cudaMalloc(&out, 128 * 4);
  cudaMemcpy(d, h.data(), W * Hh * 4, cudaMemcpyHostToDevice);
  CUtensorMap m;
  cuuint64_t size[2] = {W, Hh};
  cuuint64_t stride[1] = {W * sizeof(int)};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t es[2] = {1, 1};
  CUresult rc = enc(&m, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, d, size, stride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (V == 0) kernel<0><<<1, 128>>>(m, 64, 3, out);
  if (V == 1) kernel<1><<<1, 128>>>(m, 64, 3, out);
  if (V == 4) kernel<4><<<1, 128>>>(m, X0, 3, out);
  cudaError_t e = cudaDeviceSynchronize();
  int r[4] = {0,0,0,0}; cudaMemcpy(r, out, 16, cudaMemcpyDeviceToHost);
  printf("V=%d encode rc=%d sync: %s out %d %d expect %d\n", V, (int)rc, cudaGetErrorString(e), r[0], r[1], 3 * W + X0);
  return 0;
}
