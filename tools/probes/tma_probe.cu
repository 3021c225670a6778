#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <vector>
#define BW 32
#define BH 16
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
// variant: 0 = 3d grid_constant, 1 = 3d descriptor in global memory, 2 = 2d grid_constant, 3 = 1-D bulk copy only
template <int V>
__global__ void probe(const __grid_constant__ CUtensorMap tm, const CUtensorMap* gtm, const float* src, float* out, int x0, int y0, int z0, int sd, int* status) {
  extern __shared__ unsigned char dyn[];
  __shared__ uint64_t bar;
  float* stg = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(dyn) + 1023) & ~(uintptr_t)1023);
  const unsigned bytes = BW * BH * sd * 4;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
    if (V == 0)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n"
                   ::"r"(smem_u32(stg)), "l"(&tm), "r"(smem_u32(&bar)), "r"(x0), "r"(y0), "r"(z0) : "memory");
    if (V == 1)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n"
                   ::"r"(smem_u32(stg)), "l"(gtm), "r"(smem_u32(&bar)), "r"(x0), "r"(y0), "r"(z0) : "memory");
    if (V == 2)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n"
                   ::"r"(smem_u32(stg)), "l"(&tm), "r"(smem_u32(&bar)), "r"(x0), "r"(y0) : "memory");
    if (V == 3)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                   ::"r"(smem_u32(stg)), "l"(src), "r"(bytes), "r"(smem_u32(&bar)) : "memory");
  }
  unsigned done = 0; unsigned spin = 0;
  while (!done && spin < (1u << 22)) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    spin++;
  }
  if (threadIdx.x == 0) { status[0] = done; status[1] = spin; }
  __syncthreads();
  for (int e = threadIdx.x; e < BW * BH * sd; e += blockDim.x) out[e] = stg[e];
}
typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                              const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
  const int V = argc > 1 ? atoi(argv[1]) : 0, sw = argc > 2 ? atoi(argv[2]) : 0, sd = argc > 3 ? atoi(argv[3]) : 4;
  const int nx = 640, ny = 320, ns = 8;
  std::vector<float> h((size_t)nx * ny * ns);
  for (int s = 0; s < ns; s++) for (int j = 0; j < ny; j++) for (int i = 0; i < nx; i++) h[((size_t)s * ny + j) * nx + i] = s * 1000000.f + j * 1000.f + i;
  float* d; cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  float* out; cudaMalloc(&out, BW * BH * 8 * 4);
  int* st; cudaMalloc(&st, 8);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  CUtensorMap tm;
  CUresult r;
  if (V == 2) {
    cuuint64_t dims[2] = {nx, ny}, strides[1] = {(cuuint64_t)nx * 4};
    cuuint32_t box[2] = {BW, BH}, es[2] = {1, 1};
    r = ((encode_fn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        sw ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    cuuint64_t dims[3] = {nx, ny, ns}, strides[2] = {(cuuint64_t)nx * 4, (cuuint64_t)nx * ny * 4};
    cuuint32_t box[3] = {BW, BH, (cuuint32_t)sd}, es[3] = {1, 1, 1};
    r = ((encode_fn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        sw ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  printf("V=%d sw=%d sd=%d encode -> %d\n", V, sw, sd, (int)r);
  CUtensorMap* gtm; cudaMalloc(&gtm, sizeof(CUtensorMap)); cudaMemcpy(gtm, &tm, sizeof(CUtensorMap), cudaMemcpyHostToDevice);
  const int sde = V == 2 ? 1 : sd;
  int x0 = 5, y0 = 7;
  if (V == 0) { cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 50176); probe<0><<<1, 256, 50176>>>(tm, gtm, d, out, x0, y0, 0, sde, st); }
  if (V == 1) { cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 50176); probe<1><<<1, 256, 50176>>>(tm, gtm, d, out, x0, y0, 0, sde, st); }
  if (V == 2) { cudaFuncSetAttribute(probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 50176); probe<2><<<1, 256, 50176>>>(tm, gtm, d, out, x0, y0, 0, sde, st); }
  if (V == 3) { cudaFuncSetAttribute(probe<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 50176); probe<3><<<1, 256, 50176>>>(tm, gtm, d, out, x0, y0, 0, sde, st); }
  cudaError_t e = cudaDeviceSynchronize();
  int hs[2] = {0, 0}; cudaMemcpy(hs, st, 8, cudaMemcpyDeviceToHost);
  printf("  err=%s done=%d spin=%d\n", cudaGetErrorString(e), hs[0], hs[1]);
  if (e == cudaSuccess && V != 3) {
    std::vector<float> ho(BW * BH * sde); cudaMemcpy(ho.data(), out, ho.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int s = 0; s < sde; s++) for (int r_ = 0; r_ < BH; r_++) for (int c = 0; c < BW; c++) {
      int gi = x0 + c, gj = y0 + r_;
      float want = s * 1000000.f + gj * 1000.f + gi;
      int idx = s * BW * BH + r_ * BW + (sw ? ((((c >> 2) ^ (r_ & 7)) << 2) | (c & 3)) : c);
      if (ho[idx] != want) { if (bad < 3) printf("  mismatch s=%d r=%d c=%d got %f want %f\n", s, r_, c, ho[idx], want); bad++; }
    }
    printf("  layout mismatches: %d\n", bad);
  }
  return 0;
}
