// standalone probe: which f64 TMA box shapes work on this GPU
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int RANK>
__global__ void probe(const __grid_constant__ CUtensorMap tmap, double* out, int nbytes, int c0, int c1, int c2, int c3, int fmode) {
  __shared__ alignas(128) unsigned char buf[4096];
  __shared__ alignas(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    if (fmode == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (fmode == 1) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (fmode == 2) { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(nbytes) : "memory");
    if (RANK == 4)
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                   ::"r"(smem_u32(buf)), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(smem_u32(&bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(smem_u32(buf)), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(smem_u32(&bar)), "r"(c0), "r"(c1) : "memory");
  }
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
  const double* b = reinterpret_cast<const double*>(buf);
  for (int i = threadIdx.x; i < nbytes / 8; i += blockDim.x) out[i] = b[i];
}

typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  int only = argc > 1 ? atoi(argv[1]) : -1; int ci = -1; int fmode = argc > 2 ? atoi(argv[2]) : 0; int l2 = argc > 3 ? atoi(argv[3]) : 1;
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  PFN enc = (PFN)p;
  printf("entry %p q=%d\n", p, (int)q);
  const int X = 22, Y = 26, Z = 50, V = 5;
  const size_t n = (size_t)V * X * Y * Z;
  std::vector<double> h(n);
  for (size_t i = 0; i < n; ++i) h[i] = (double)i;
  double *d, *out; cudaMalloc(&d, n * 8); cudaMalloc(&out, 4096);
  cudaMemcpy(d, h.data(), n * 8, cudaMemcpyHostToDevice);
  struct Case { int rank; int b0, b1, b2, b3; CUtensorMapDataType dt; int esz; const char* name; };
  Case cases[] = {
    {4, 40, 1, 1, 5, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, "f64 rank4 box40x1x1x5"},
    {4, 32, 1, 1, 5, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, "f64 rank4 box32x1x1x5"},
    {4, 40, 1, 1, 1, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, "f64 rank4 box40x1x1x1"},
    {2, 40, 2, 0, 0, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, "f64 rank2 box40x2"},
    {4, 80, 1, 1, 5, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, "u32 rank4 box80x1x1x5 (same bytes)"},
    {4, 40, 1, 1, 5, CU_TENSOR_MAP_DATA_TYPE_UINT64, 8, "u64 rank4 box40x1x1x5"},
    {2, 32, 32, 0, 0, CU_TENSOR_MAP_DATA_TYPE_INT32, 4, "CANON i32 rank2 256x256 box32x32"},
    {2, 32, 2, 0, 0, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, "f64 rank2 dims(64,26) pitch512 box32x2"},
    {2, 40, 2, 0, 0, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, "f64 rank2 dims(64,26) pitch512 box40x2"},
    {2, 32, 2, 0, 0, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, "f64 rank2 dims(50,26) pitch400 box32x2"},
  };
  for (auto& c : cases) {
    ++ci; if (only >= 0 && ci != only) continue;
    CUtensorMap m;
    const int s = 8 / c.esz;
    cuuint64_t dims4[4] = {(cuuint64_t)Z * s, Y, X, V};
    cuuint64_t str4[3] = {(cuuint64_t)Z * 8, (cuuint64_t)Y * Z * 8, (cuuint64_t)X * Y * Z * 8};
    if (ci == 6) { dims4[0] = 256; dims4[1] = 256; str4[0] = 1024; }
    if (ci == 7 || ci == 8) { dims4[0] = 64; dims4[1] = 26; str4[0] = 512; }
    cuuint32_t box[4] = {(cuuint32_t)c.b0, (cuuint32_t)c.b1, (cuuint32_t)c.b2, (cuuint32_t)c.b3};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult rc = enc(&m, c.dt, c.rank, d, dims4, str4, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                      l2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    int nb = c.b0 * c.esz * c.b1 * (c.rank == 4 ? c.b2 * c.b3 : 1);
    printf("%-40s encode rc=%d bytes=%d : ", c.name, (int)rc, nb);
    if (rc != CUDA_SUCCESS) { printf("\n"); continue; }
    cudaMemset(out, 0, 4096);
    if (c.rank == 4) probe<4><<<1, 128>>>(m, out, nb, 3 * s, 7, 9, 0, fmode); else probe<2><<<1, 128>>>(m, out, nb, 3 * s, 7, 0, 0, fmode);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("LAUNCH ERROR %s\n", cudaGetErrorString(e)); return 1; }
    double r[4]; cudaMemcpy(r, out, 32, cudaMemcpyDeviceToHost);
    double expect = (double)((9 * Y + 7) * Z + 3);
    printf("out[0..2]=%.0f %.0f %.0f expect %.0f\n", r[0], r[1], r[2], expect);
  }
  return 0;
}
