// canonical CUDA-programming-guide TMA example (libcu++ wrappers)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <stdio.h>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;

__global__ void kernel(const __grid_constant__ CUtensorMap tensor_map, int x, int y, int* out) {
  __shared__ alignas(128) int smem_buffer[32][32];
#pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ barrier bar;
  if (threadIdx.x == 0) {
    init(&bar, blockDim.x);
    cde::fence_proxy_async_shared_cta();
  }
  __syncthreads();
  barrier::arrival_token token;
  if (threadIdx.x == 0) {
    cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
    token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
  } else {
    token = bar.arrive();
  }
  bar.wait(std::move(token));
  out[threadIdx.x] = smem_buffer[0][threadIdx.x % 32];
}

typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
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
  printf("encode rc=%d\n", (int)rc);
  kernel<<<1, 128>>>(m, 64, 3, out);
  cudaError_t e = cudaDeviceSynchronize();
  printf("sync: %s\n", cudaGetErrorString(e));
  int r[4]; cudaMemcpy(r, out, 16, cudaMemcpyDeviceToHost);
  printf("out %d %d %d expect %d\n", r[0], r[1], r[2], 3 * W + 64);
  return 0;
}
