// Plumbing test (measurement / bring-up aid, not product): TMA im2col load of the conv engine's padded flat window + tcgen05.mma reading
// it through SWIZZLE_64B K-major descriptors at arbitrary row offsets (the "shifted view" of a 3x3 tap).  Checks window contents and the
// conv result against a CPU loop.   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I../../diffusion_model_nemo_b200/csrc ...
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "tc_ptx.cuh"
using namespace dmn::tc;

typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const int*, const int*, cuuint32_t,
                                   cuuint32_t, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

constexpr int B = 3, H = 8, W = 8, C = 32, N = 32, Wv = W + 1, S = (H + 1) * Wv, HALO = Wv + 1, P = 128 + 2 * HALO;

__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t saddr, uint32_t base_off) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(512u >> 4) << 32) | (1ull << 46) | ((uint64_t)(base_off & 7) << 49) | (4ull << 61);
}

__global__ void __launch_bounds__(128) k(const __grid_constant__ CUtensorMap map, const __nv_bfloat16* wpk, int m0, int bo_mode, uint8_t* a_dump, float* d_out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  uint8_t* sA = sm;                         // P x 64 B (SW64), 1024-aligned
  uint8_t* sW = sm + 10240;                 // 9 taps x [4 kc][32 n][16 B] = 9 x 2048
  __shared__ uint64_t bars[2];
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { mbar_init(smem_u32(&bars[0]), 1); mbar_init(smem_u32(&bars[1]), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&tslot), 32);
  for (int i = tid; i < 9 * 2048 / 16; i += 128) reinterpret_cast<uint4*>(sW)[i] = reinterpret_cast<const uint4*>(wpk)[i];
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tslot;
  if (tid == 0) {
    int f = m0 - HALO;
    int n = f >= 0 ? f / S : -((-f + S - 1) / S);
    int rem = f - n * S;
    int h = rem / Wv, w = rem - h * Wv;
    mbar_arrive_expect_tx(smem_u32(&bars[0]), P * 64);
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6], {%7, %8};"
                 ::"r"(smem_u32(sA)), "l"(&map), "r"(0), "r"(w), "r"(h), "r"(n), "r"(smem_u32(&bars[0])), "h"((uint16_t)0), "h"((uint16_t)0) : "memory");
  }
  mbar_wait(smem_u32(&bars[0]), 0);
  for (int i = tid; i < P * 64 / 16; i += 128) reinterpret_cast<uint4*>(a_dump)[i] = reinterpret_cast<const uint4*>(sA)[i];
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc(128, N);
    for (int t = 0; t < 9; ++t) {
      const int dy = t / 3 - 1, dx = t % 3 - 1;
      const uint32_t a0 = smem_u32(sA) + (uint32_t)(HALO + dy * Wv + dx) * 64u;
      for (int ks = 0; ks < 2; ++ks) {
        const uint32_t a = a0 + ks * 32;
        uint32_t bo = 0;
        if (bo_mode == 1) bo = (a0 >> 7) & 3;
        if (bo_mode == 2) bo = (a0 >> 7) & 7;
        if (bo_mode == 3) bo = (a0 >> 6) & 7;
        umma_bf16(tmem, make_desc_sw64(a, bo), make_desc(smem_u32(sW) + t * 2048 + ks * 2 * 512, 512, 128), idesc, (t | ks) ? 1u : 0u);
      }
    }
    umma_commit(smem_u32(&bars[1]));
  }
  mbar_wait(smem_u32(&bars[1]), 0);
  tc_fence_after();
  uint32_t r[32];
  tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), r);
  for (int j = 0; j < 32; ++j) d_out[(warp * 32 + lane) * N + j] = __uint_as_float(r[j]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 32);
}

__global__ void kdown(const __grid_constant__ CUtensorMap map, int f, int sx, int sy, int Pd, uint8_t* a_dump) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bar;
  const int Wd = W / 2 + 1, Sd = (H / 2 + 1) * Wd;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    int n = f >= 0 ? f / Sd : -((-f + Sd - 1) / Sd);
    int rem = f - n * Sd;
    int r = rem / Wd, c = rem - r * Wd;
    mbar_arrive_expect_tx(smem_u32(&bar), Pd * 64);
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6], {%7, %8};"
                 ::"r"(smem_u32(sm)), "l"(&map), "r"(0), "r"(2 * c - 1), "r"(2 * r - 1), "r"(n), "r"(smem_u32(&bar)), "h"((uint16_t)sx), "h"((uint16_t)sy) : "memory");
  }
  mbar_wait(smem_u32(&bar), 0);
  for (int i = threadIdx.x; i < Pd * 64 / 16; i += blockDim.x) reinterpret_cast<uint4*>(a_dump)[i] = reinterpret_cast<const uint4*>(sm)[i];
}

int main() {
  void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &qr);
  if (!fn) { printf("no cuTensorMapEncodeIm2col\n"); return 1; }
  EncodeIm2colFn enc = (EncodeIm2colFn)fn;
  std::vector<__nv_bfloat16> x(B * H * W * C), wq(N * C * 9), wpk(9 * 4 * N * 8);
  std::vector<float> xf(x.size()), wf(wq.size());
  srand(1);
  for (size_t i = 0; i < x.size(); ++i) { x[i] = __float2bfloat16((rand() % 17 - 8) / 8.f); xf[i] = __bfloat162float(x[i]); }
  for (size_t i = 0; i < wq.size(); ++i) { wq[i] = __float2bfloat16((rand() % 9 - 4) / 4.f); wf[i] = __bfloat162float(wq[i]); }
  for (int t = 0; t < 9; ++t) for (int kc = 0; kc < 4; ++kc) for (int n = 0; n < N; ++n) for (int e = 0; e < 8; ++e)
    wpk[((t * 4 + kc) * N + n) * 8 + e] = wq[(n * C + kc * 8 + e) * 9 + t];      // w[n][c][tap]
  __nv_bfloat16 *dx, *dw; uint8_t* da; float* dd;
  cudaMalloc(&dx, x.size() * 2); cudaMalloc(&dw, wpk.size() * 2); cudaMalloc(&da, P * 64); cudaMalloc(&dd, 128 * N * 4);
  cudaMemcpy(dx, x.data(), x.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dw, wpk.data(), wpk.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap m;
  cuuint64_t dims[4] = {C, W, H, B};
  cuuint64_t strides[3] = {C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  int lo[2] = {0, 0}, hi[2] = {1, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult rc = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dx, dims, strides, lo, hi, C, P, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d\n", (int)rc);
  if (rc) return 1;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 10240 + 9 * 2048);
  for (int m0 : {0, 128}) for (int bo = 0; bo < 4; ++bo) {
    cudaMemset(da, 0xff, P * 64); cudaMemset(dd, 0, 128 * N * 4);
    k<<<1, 128, 10240 + 9 * 2048>>>(m, dw, m0, bo, da, dd);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<__nv_bfloat16> a(P * 32); std::vector<float> d(128 * N);
    cudaMemcpy(a.data(), da, P * 64, cudaMemcpyDeviceToHost); cudaMemcpy(d.data(), dd, 128 * N * 4, cudaMemcpyDeviceToHost);
    // window check (un-swizzle SW64: 16-byte chunk ^= (row >> 1) & 3)
    int bad_w = 0;
    auto xat = [&](int f, int c) -> float {       // padded flat index -> value
      if (f < 0) return 0.f;
      int n = f / S, rem = f % S, h = rem / Wv, w = rem % Wv;
      if (n >= B || h >= H || w >= W) return 0.f;
      return xf[((n * H + h) * W + w) * C + c];
    };
    for (int r = 0; r < P; ++r) for (int c = 0; c < C; ++c) {
      const int chunk = (c / 8) ^ ((r >> 1) & 3);
      const float got = __bfloat162float(a[r * 32 + chunk * 8 + (c & 7)]);
      if (got != xat(m0 - HALO + r, c)) { if (bad_w < 5) printf("   window mismatch row %d c %d got %f want %f\n", r, c, got, xat(m0 - HALO + r, c)); ++bad_w; }
    }
    double maxerr = 0;
    for (int mm = 0; mm < 128; ++mm) for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int t = 0; t < 9; ++t) { const int dy = t / 3 - 1, dxx = t % 3 - 1; for (int c = 0; c < C; ++c) ref += xat(m0 + mm + dy * Wv + dxx, c) * wf[(n * C + c) * 9 + t]; }
      maxerr = fmax(maxerr, fabs(ref - d[mm * N + n]));
    }
    printf("m0=%3d bo_mode=%d err=%d : window mismatches %d, conv max abs err %.4g\n", m0, bo, (int)e, bad_w, maxerr);
  }
  // ---- stride-2 window (GEO_DOWN): positions (r, c) of a (H/2+1) x (W/2+1) grid read pixel (2r-1+sy, 2c-1+sx) ----
  for (int up = 0; up <= 1; ++up) {
    CUtensorMap md;
    int lo2[2] = {-1, -1}, hi2[2] = {up, up};
    cuuint32_t es2[4] = {1, 2, 2, 1};
    const int Pd = 64, Wd = W / 2 + 1, Sd = (H / 2 + 1) * Wd;
    CUresult rc2 = enc(&md, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dx, dims, strides, lo2, hi2, C, Pd, es2, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("down encode (upper=%d) rc=%d\n", up, (int)rc2);
    if (rc2) continue;
    uint8_t* da2; cudaMalloc(&da2, Pd * 64);
    for (int f : {0, 7, 30}) for (int sub = 0; sub < 4; ++sub) {
      const int sy = sub >> 1, sx = sub & 1;
      cudaMemset(da2, 0xff, Pd * 64);
      kdown<<<1, 128, Pd * 64 + 1024>>>(md, f, sx, sy, Pd, da2);
      cudaError_t e = cudaDeviceSynchronize();
      std::vector<__nv_bfloat16> a(Pd * 32);
      cudaMemcpy(a.data(), da2, Pd * 64, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int r = 0; r < Pd; ++r) for (int c = 0; c < C; ++c) {
        const int ff = f + r, n = ff / Sd, rem = ff % Sd, pr = rem / Wd, pc = rem % Wd;
        const int iy = 2 * pr - 1 + sy, ix = 2 * pc - 1 + sx;
        const float want = (n < B && iy >= 0 && iy < H && ix >= 0 && ix < W) ? xf[((n * H + iy) * W + ix) * C + c] : 0.f;
        const int chunk = (c / 8) ^ ((r >> 1) & 3);
        const float got = __bfloat162float(a[r * 32 + chunk * 8 + (c & 7)]);
        if (got != want) { if (bad < 3) printf("   down mismatch f=%d row %d c %d got %f want %f\n", f, r, c, got, want); ++bad; }
      }
      printf("down upper=%d f=%2d sub=(%d,%d) err=%d mismatches %d\n", up, f, sy, sx, (int)e, bad);
    }
  }
  return 0;
}
