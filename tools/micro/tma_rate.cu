// Micro-benchmark (measurement aid): TMA tensor loads that produce the conv engine's operand layout [k-chunk][pixel][16 B] directly:
// 3-D view of x[pixels][C] bf16 as (8 elements, pixels (stride 2C bytes), C/8 chunks (stride 16 bytes)), box {8, R, 4}.
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ void mbar_init(uint32_t b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint32_t b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t b, uint32_t ph) {
  asm volatile("{\n.reg .pred p;\nW: mbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n@!p bra W;\n}" ::"r"(b), "r"(ph) : "memory");
}

template <int NBUF>
__global__ void __launch_bounds__(128) k(const __grid_constant__ CUtensorMap map, int C, int R, long pixels_per_cta, int passes, long long* out, int inner_ch) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bars[NBUF];
  const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(sm);
  const uint32_t bytes = (uint32_t)R * (inner_ch ? inner_ch * 2u : 64u);
  if (threadIdx.x == 0) {
    for (int i = 0; i < NBUF; ++i) mbar_init((uint32_t)__cvta_generic_to_shared(&bars[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    const long pbase = (long)blockIdx.x * pixels_per_cta;
    const int rounds = passes / NBUF;
    for (int r = 0; r < rounds; ++r) {
      const int pix = (int)(pbase + ((r * 7) & 63) * R);
#pragma unroll
      for (int i = 0; i < NBUF; ++i) {
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&bars[i]);
        mbar_expect(bar, bytes);
        if (inner_ch == 0)
          asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                       ::"r"(s0 + i * bytes), "l"(&map), "r"(0), "r"(pix + i * R), "r"((i & 3) * 4), "r"(bar) : "memory");
        else
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                       ::"r"(s0 + i * bytes), "l"(&map), "r"((i % (C / inner_ch)) * inner_ch), "r"(pix + (i / (C / inner_ch)) * R), "r"(bar) : "memory");
      }
#pragma unroll
      for (int i = 0; i < NBUF; ++i) mbar_wait((uint32_t)__cvta_generic_to_shared(&bars[i]), r & 1);
    }
    out[0] = clock64() - t0;
  }
}

int main() {
  void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
  EncodeFn enc = (EncodeFn)fn;
  uint8_t* x; long long* d_out;
  cudaMalloc(&x, 80ull << 20); cudaMemset(x, 1, 80ull << 20); cudaMalloc(&d_out, 64);
  for (int C : {128}) for (int R : {256}) {
    const long npix = (64l << 20) / 2 / C;
    CUtensorMap m;
    cuuint64_t dims[3] = {8, (cuuint64_t)npix, (cuuint64_t)C / 8};
    cuuint64_t strides[2] = {(cuuint64_t)C * 2, 16};
    cuuint32_t box[3] = {8, (cuuint32_t)R, 4}, es[3] = {1, 1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed C=%d R=%d rc=%d\n", C, R, (int)r); continue; }
    const int passes = 400 * 256 / R;
    const long ppc = npix / 148;
    cudaFuncSetAttribute(k<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * R * 64 + 1024);
    for (int rep = 0; rep < 2; ++rep) k<4><<<148, 128, 4 * R * 64 + 1024>>>(m, C, R, ppc, passes, d_out, 0);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
    printf("tma 3d box{8,%3d,4} C=%3d : %7.0f clk per box of %5d B -> %5.1f B/clk/SM  err=%d\n", R, C, (double)h / passes, R * 64, (double)R * 64 * passes / (double)h, (int)e);
  }
  for (int C : {128}) for (int inner : {32, 64}) for (int R : {64, 128, 256}) {
    const long npix = (64l << 20) / 2 / C;
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)npix};
    cuuint64_t strides[1] = {(cuuint64_t)C * 2};
    cuuint32_t box[2] = {(cuuint32_t)inner, (cuuint32_t)R}, es[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     inner == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed C=%d R=%d rc=%d\n", C, R, (int)r); continue; }
    const int passes = 400 * 256 / R;
    const long ppc = npix / 148;
    for (int nb : {1, 2, 4, 8}) {
      const int smem = nb * R * inner * 2 + 1024;
      if (smem > 200 * 1024) continue;
      auto kern = nb == 1 ? k<1> : (nb == 2 ? k<2> : (nb == 4 ? k<4> : k<8>));
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      for (int rep = 0; rep < 2; ++rep) kern<<<148, 128, smem>>>(m, C, R, ppc, passes, d_out, inner);
      cudaError_t e = cudaDeviceSynchronize();
      long long h = 0; cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
      printf("tma 2d box{%d ch,%3d px} SW%d C=%3d nbuf=%d : %7.0f clk per round, box %5d B -> %5.1f B/clk/SM  err=%d\n", inner, R, inner * 2, C, nb, (double)h / (passes / nb), R * inner * 2,
             (double)R * inner * 2 * passes / (double)h, (int)e);
    }
  }
  return 0;
}
