// Micro-benchmark (measurement aid, not product): issue cost of 16-byte cp.async (LDGSTS) with the conv producers' access pattern.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldgsts_rate ldgsts_rate.cu ; run: ./ldgsts_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__device__ __forceinline__ void cp16(uint32_t dst, const void* src, bool ok) {
  if (MODE == 0) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
  if (MODE == 1) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
  if (MODE == 2) { uint32_t sz = ok ? 16u : 0u; asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory"); }
}

// each CTA walks its own contiguous slice of x: [pixels][C] bf16; a "pass" = 32 channels of a 320-pixel window
template <int MODE, int ITEMS>
__global__ void __launch_bounds__(256) k(const uint8_t* x, int C, long pixels_per_cta, int passes, int depth, long long* out, int planes_pitch) {
  extern __shared__ __align__(128) uint8_t sm[];
  const int tid = threadIdx.x, kc = tid & 3, px0 = tid >> 2;
  const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(sm);
  const uint8_t* base = x + (long)blockIdx.x * pixels_per_cta * C * 2;
  long long t_issue = 0, t_all0 = clock64();
  int inflight = 0;
  for (int c = 0; c < passes; ++c) {
    const int buf = c % 4;
    const long pix0 = (long)(c / (C / 32)) * 256 % (pixels_per_cta - 512);   // tile start
    const int cb = (c % (C / 32)) * 32 + kc * 8;
    const uint8_t* src = base + (pix0 * C + cb) * 2;
    const uint32_t dst = s0 + buf * (4 * planes_pitch) + kc * planes_pitch + px0 * 16;
    const long long t0 = clock64();
    if (MODE == 3) {
      uint4 v[ITEMS];
#pragma unroll
      for (int j = 0; j < ITEMS; ++j) v[j] = __ldcg(reinterpret_cast<const uint4*>(src + (long)(px0 + 64 * j) * C * 2));
#pragma unroll
      for (int j = 0; j < ITEMS; ++j) asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst + j * 64 * 16), "r"(v[j].x), "r"(v[j].y), "r"(v[j].z), "r"(v[j].w) : "memory");
    } else {
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) cp16<MODE>(dst + j * 64 * 16, src + (long)(px0 + 64 * j) * C * 2, true);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    t_issue += clock64() - t0;
    if (++inflight > depth) {
      if (depth == 1) asm volatile("cp.async.wait_group 1;" ::: "memory");
      else if (depth == 2) asm volatile("cp.async.wait_group 2;" ::: "memory");
      else asm volatile("cp.async.wait_group 0;" ::: "memory");
      --inflight;
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  if (tid == 0 && blockIdx.x == 3) { out[0] = t_issue; out[1] = clock64() - t_all0; }
}

static long g_mb = 64; static int g_ctas = 148;
template <int MODE, int ITEMS>
void run(const char* name, const uint8_t* x, int C, int threads_note, int depth, long long* d_out) {
  const int passes = 400, pitch = 460 * 16;
  const long ppc = 1024 * 1024 * g_mb / 2 / C / 148;   // g_mb MB tensor over 148 CTAs
  cudaFuncSetAttribute(k<MODE, ITEMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 4 * pitch);
  for (int r = 0; r < 2; ++r) k<MODE, ITEMS><<<g_ctas, 256, 4 * 4 * pitch>>>(x, C, ppc, passes, depth, d_out, pitch);
  cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, d_out, sizeof h, cudaMemcpyDeviceToHost);
  printf("[%ld MB, %d CTAs] %-28s C=%3d items=%d depth=%d : issue %6.0f clk/pass (%5.1f clk per LDGSTS per warp), total %6.0f clk/pass -> %5.1f B/clk/SM  err=%d\n", g_mb, g_ctas, name, C, ITEMS,
         depth, (double)h[0] / passes, (double)h[0] / passes / ITEMS, (double)h[1] / passes, 256.0 * ITEMS * 16 / ((double)h[1] / passes), (int)cudaGetLastError());
}

int main() {
  uint8_t* x; long long* d_out;
  cudaMalloc(&x, 80ull << 20); cudaMemset(x, 1, 80ull << 20); cudaMalloc(&d_out, 64);
  for (int cfg = 0; cfg < 4; ++cfg) {
    g_mb = cfg == 0 ? 64 : 16; g_ctas = cfg == 2 ? 37 : (cfg == 3 ? 74 : 148);
    int C = 128;
    for (int depth : {1, 2, 3}) run<0, 5>("cg", x, C, 0, depth, d_out);
    run<1, 5>("ca", x, C, 0, 2, d_out);
    run<2, 5>("cg zfill-operand", x, C, 0, 2, d_out);
    run<0, 7>("cg", x, C, 0, 2, d_out);
    run<3, 5>("ldg.cg + sts", x, C, 0, 2, d_out);
  }
  return 0;
}
