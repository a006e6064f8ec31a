// Microbenchmark 3: L2 -> shared-memory streaming rate of cp.async.bulk when every CTA streams the SAME weight image
// (the conv engine's access pattern) vs. CTA-private images, for several stage sizes / ring depths.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/l2_stream tools/l2_stream.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// image_bytes: size of the streamed image; priv: 1 = each CTA streams its own copy (offset by blockIdx * image_bytes)
// skew: CTA b starts at stage offset (b * skew) % nchunks  (de-synchronises the CTAs' positions in the shared image)
__global__ void __launch_bounds__(128, 1) k(const uint8_t* g, int image_bytes, int stage_bytes, int nst, int priv, int skew, int reps, long long* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar[16];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint8_t* src = g + (priv ? (size_t)blockIdx.x * image_bytes : 0);
    const int nchunks = image_bytes / stage_bytes;
    const long total = (long)nchunks * reps;
    uint32_t ph[16];
    for (int i = 0; i < 16; ++i) ph[i] = 0;
    long issued = 0;
    int chunk = (int)(((long)blockIdx.x * skew) % nchunks);
    long long t0 = clock64();
    for (int s = 0; s < nst && issued < total; ++s, ++issued) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(stage_bytes) : "memory");
      bulk_g2s(smem_u32(smem) + s * stage_bytes, src + (size_t)chunk * stage_bytes, stage_bytes, smem_u32(&bar[s]));
      if (++chunk == nchunks) chunk = 0;
    }
    long done = 0;
    int s = 0;
    while (done < total) {
      while (!try_wait(smem_u32(&bar[s]), ph[s])) {}
      ph[s] ^= 1;
      ++done;
      if (issued < total) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(stage_bytes) : "memory");
        bulk_g2s(smem_u32(smem) + s * stage_bytes, src + (size_t)chunk * stage_bytes, stage_bytes, smem_u32(&bar[s]));
        if (++chunk == nchunks) chunk = 0;
        ++issued;
      }
      if (++s == nst) s = 0;
    }
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
}

int main() {
  const int grid = 148;
  long long* d;
  uint8_t* g;
  const size_t gbytes = (size_t)grid * 1024 * 1024;
  cudaMalloc(&d, grid * 8);
  cudaMalloc(&g, gbytes);
  cudaMemset(g, 1, gbytes);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 192 * 1024);
  struct C { int image, stage, nst, priv, skew; };
  const C cs[] = {
      {294912, 8192, 12, 0, 0},  {294912, 8192, 12, 1, 0},  {294912, 8192, 12, 0, 7},  {294912, 8192, 6, 0, 0},   {294912, 8192, 6, 0, 7},
      {294912, 4096, 16, 0, 0},  {294912, 16384, 8, 0, 0},  {294912, 16384, 8, 0, 5},  {294912, 16384, 8, 1, 0},  {589824, 8192, 12, 0, 0},
      {589824, 8192, 12, 0, 11}, {294912, 24576, 6, 0, 0},  {294912, 24576, 6, 0, 5},  {73728, 8192, 12, 0, 0},   {73728, 8192, 12, 0, 3},
  };
  for (const C& c : cs) {
    const int reps = 8;
    for (int it = 0; it < 2; ++it) k<<<grid, 128, c.stage * c.nst>>>(g, c.image, c.stage, c.nst, c.priv, c.skew, reps, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
    long long mx = 0, mn = 1LL << 60;
    double avg = 0;
    for (int i = 0; i < grid; ++i) { mx = h[i] > mx ? h[i] : mx; mn = h[i] < mn ? h[i] : mn; avg += (double)h[i] / grid; }
    const double bytes = (double)c.image * reps;
    printf("image=%6d stage=%5d nst=%2d priv=%d skew=%2d : %.1f B/clk/SM avg (min-CTA %.1f, max-CTA %.1f) => chip %.0f B/clk %s\n", c.image, c.stage, c.nst,
           c.priv, c.skew, bytes / avg, bytes / mx, bytes / mn, bytes / avg * grid, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
