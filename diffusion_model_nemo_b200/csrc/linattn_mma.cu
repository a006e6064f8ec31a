// linattn_mma.cu -- LinearAttention core (reference parts/mha.py:44-58) for bf16 activations on the tensor cores.
//
//   q = softmax_d(q) * scale        (per token, per head, over the 32 head channels;  scale AFTER the softmax, mha.py:54)
//   k = softmax_n(k)                (per channel, over the N tokens of the image)
//   ctx[h][d][e] = sum_n k[d,n] v[e,n]          out[n][e] = sum_d ctx[h][d][e] q[n][d]
//
// qkv: bf16 [B][N][384] (channel = which*128 + head*32 + d), out: bf16 [B][N][128]; heads = 4, dim_head = 32.
// One CTA (8 warps) per image, two passes over the image's tokens in 64-token tiles that stream through a cp.async double
// buffer:
//   pass A  k|v tiles: ONLINE column softmax of k (running max per channel, accumulators rescaled when it moves), P = exp(k - max)
//           written back in place as bf16, ctx += P^T V with mma.sync m16n8k16 (operands via ldmatrix.trans straight from the
//           token-major tile), column sums in fp32 -> ctx / sum as bf16 in shared memory
//   pass B  q tiles: per-token softmax in place, out = q ctx with mma.sync (ctx fragments live in registers), result written
//           in place and stored with full 256-byte rows.
// HBM-bound by design: qkv is read once (k|v in pass A, q in pass B) and out written once = 4 x 128 x 2 B per token.
// The 0.5 % of the U-Net's FLOPs that live here do not justify a tcgen05/TMEM pipeline; legacy warp-level MMA keeps the
// kernel at the memory roofline instead of the FP32-FMA roofline.  fp32 activations use the CUDA-core kernel in kernels_simt.cu.
#include "common.cuh"
#include "ops.h"
#include "tc_ptx.cuh"

namespace dmn {
namespace la {

using tc::cp_async16;
using tc::cp_async_commit;
using tc::cp_async_wait;
using tc::lds128;
using tc::smem_u32;
using tc::sts128;

constexpr int TN = 64;        // tokens per tile
constexpr int KV_LD = 528;    // bytes per token row of a k|v tile: 256 channels x 2 B + 16 B pad (ldmatrix rows hit distinct banks)
constexpr int Q_LD = 272;     // bytes per token row of a q tile: 128 channels x 2 B + 16 B pad
constexpr int CTX_LD = 80;    // bytes per d row of the bf16 context: 32 e x 2 B + 16 B pad
constexpr int kTileBytes = 2 * TN * KV_LD;                         // double buffer (the q tiles reuse it)
constexpr int kCtxBytes = 4 * 32 * CTX_LD;
constexpr int kSmem = kTileBytes + kCtxBytes + 128 * 4 * 10;       // + fmax, fac, red[4], ksum[4]

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float* d, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ld_bf16(uint32_t addr) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
  return __uint_as_float((uint32_t)v << 16);
}
__device__ __forceinline__ void st_bf16(uint32_t addr, bf16 v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(*reinterpret_cast<unsigned short*>(&v)) : "memory");
}
__device__ __forceinline__ void st_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

__global__ void __launch_bounds__(256, 2) linattn_mma_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, int N) {
  extern __shared__ __align__(128) uint8_t sm[];
  pdl_trigger();
  pdl_wait();
  const uint32_t tiles = smem_u32(sm);
  const uint32_t ctxs = tiles + kTileBytes;
  float* fmax_s = reinterpret_cast<float*>(sm + kTileBytes + kCtxBytes);   // running column max of k
  float* fac_s = fmax_s + 128;                                              // exp(old max - new max) of this tile
  float* red_s = fac_s + 128;                                               // [4][128] per-quarter tile max
  float* ksum_s = red_s + 512;                                              // [4][128] per-quarter column sums
  const int b = blockIdx.x, t = threadIdx.x, w = t >> 5, lane = t & 31;
  const bf16* base = qkv + (long)b * N * 384;
  const int ntiles = (N + TN - 1) / TN;
  const int h = w >> 1;                      // head of this warp in both MMA phases
  const int lmat = lane >> 3, li = lane & 7;   // ldmatrix: which 8x8 matrix this lane addresses, row within it
  const int g = lane >> 2, tq = lane & 3;      // mma fragment coordinates

  // ---------------------------------------------------------------- pass A: context ----------------------------------------
  auto load_kv = [&](int tile, int buf) {
    const uint32_t dst0 = tiles + (uint32_t)(buf * TN * KV_LD);
    for (int i = t; i < TN * 32; i += 256) {
      const int r = i >> 5, c16 = i & 31, n = tile * TN + r;
      const uint32_t dst = dst0 + (uint32_t)(r * KV_LD + c16 * 16);
      if (n < N) cp_async16(dst, base + (long)n * 384 + 128 + c16 * 8);
      else sts128(dst, make_uint4(0, 0, 0, 0));
    }
    cp_async_commit();
  };
  if (t < 128) fmax_s[t] = -INFINITY;
  // column-softmax role: a PAIR of adjacent channels (one 32-bit word per row) and a quarter of the tile's tokens
  const int c2 = (t & 63) * 2, qtr = t >> 6;
  float ksum0 = 0.f, ksum1 = 0.f;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int dh = (w & 1) * 16;               // d half of this warp's 16 x 32 context block
  load_kv(0, 0);
  for (int tile = 0; tile < ntiles; ++tile) {
    const int buf = tile & 1;
    const uint32_t tb = tiles + (uint32_t)(buf * TN * KV_LD);
    if (tile + 1 < ntiles) { load_kv(tile + 1, buf ^ 1); cp_async_wait<1>(); }
    else cp_async_wait<0>();
    __syncthreads();
    const int nvalid = min(TN, N - tile * TN);
    // tile column max (this thread: channels c2, c2 + 1, rows [16*qtr, 16*qtr + 16)); the values stay in registers for the exp pass
    uint32_t kw[16];
    {
      float m0 = -INFINITY, m1 = -INFINITY;
      const uint32_t a0 = tb + (uint32_t)(qtr * 16 * KV_LD + c2 * 2);
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(kw[r]) : "r"(a0 + (uint32_t)(r * KV_LD)));
        if (qtr * 16 + r < nvalid) {
          m0 = fmaxf(m0, __uint_as_float(kw[r] << 16));
          m1 = fmaxf(m1, __uint_as_float(kw[r] & 0xffff0000u));
        }
      }
      red_s[qtr * 128 + c2] = m0;
      red_s[qtr * 128 + c2 + 1] = m1;
    }
    __syncthreads();
    if (t < 128) {
      const float mo = fmax_s[t], mn = fmaxf(fmaxf(mo, fmaxf(red_s[t], red_s[128 + t])), fmaxf(red_s[256 + t], red_s[384 + t]));
      fac_s[t] = __expf(mo - mn);            // first tile: exp(-inf) = 0
      fmax_s[t] = mn;
    }
    __syncthreads();
    // P = exp(k - max) in place (bf16); column sums of the ROUNDED values so numerator and denominator agree
    {
      const float mn0 = fmax_s[c2], mn1 = fmax_s[c2 + 1];
      ksum0 *= fac_s[c2];
      ksum1 *= fac_s[c2 + 1];
      const uint32_t a0 = tb + (uint32_t)(qtr * 16 * KV_LD + c2 * 2);
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        uint32_t pw = 0u;
        if (qtr * 16 + r < nvalid) {
          const bf16 p0 = __float2bfloat16_rn(__expf(__uint_as_float(kw[r] << 16) - mn0));
          const bf16 p1 = __float2bfloat16_rn(__expf(__uint_as_float(kw[r] & 0xffff0000u) - mn1));
          ksum0 += __bfloat162float(p0);
          ksum1 += __bfloat162float(p1);
          pw = (uint32_t)(*reinterpret_cast<const unsigned short*>(&p0)) | ((uint32_t)(*reinterpret_cast<const unsigned short*>(&p1)) << 16);
        }
        st_u32(a0 + (uint32_t)(r * KV_LD), pw);
      }
    }
    __syncthreads();
    // ctx[h][dh + 0..15][0..31] = fac * ctx + P^T V over the tile's 64 tokens
    {
      const float f0 = fac_s[h * 32 + dh + g], f1 = fac_s[h * 32 + dh + g + 8];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) { acc[nt][0] *= f0; acc[nt][1] *= f0; acc[nt][2] *= f1; acc[nt][3] *= f1; }
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t a[4];
        // A[m = d][k = token] = P[token][d]: transposed 8x8 loads; matrices (d lo, tok lo), (d hi, tok lo), (d lo, tok hi), (d hi, tok hi)
        ldsm_x4_t(tb + (uint32_t)((ks * 16 + (lmat >> 1) * 8 + li) * KV_LD + (h * 32 + dh + (lmat & 1) * 8) * 2), a[0], a[1], a[2], a[3]);
#pragma unroll
        for (int pr = 0; pr < 2; ++pr) {
          uint32_t b0, b1, b2, b3;
          // B[k = token][n = e] = V[token][e]: matrices (tok lo, e lo), (tok hi, e lo), (tok lo, e hi), (tok hi, e hi)
          ldsm_x4_t(tb + (uint32_t)((ks * 16 + (lmat & 1) * 8 + li) * KV_LD + 256 + (h * 32 + (2 * pr + (lmat >> 1)) * 8) * 2), b0, b1, b2, b3);
          mma16816(acc[2 * pr], a, b0, b1);
          mma16816(acc[2 * pr + 1], a, b2, b3);
        }
      }
    }
    __syncthreads();      // the next iteration's prefetch overwrites the buffer read here one tile later
  }
  ksum_s[qtr * 128 + c2] = ksum0;
  ksum_s[qtr * 128 + c2 + 1] = ksum1;
  __syncthreads();
  {
    const int ch0 = h * 32 + dh + g, ch1 = ch0 + 8;
    const float i0 = 1.f / ((ksum_s[ch0] + ksum_s[128 + ch0]) + (ksum_s[256 + ch0] + ksum_s[384 + ch0]));
    const float i1 = 1.f / ((ksum_s[ch1] + ksum_s[128 + ch1]) + (ksum_s[256 + ch1] + ksum_s[384 + ch1]));
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const uint32_t a = ctxs + (uint32_t)((h * 32 + dh + g) * CTX_LD + (nt * 8 + 2 * tq) * 2);
      st_u32(a, pack_bf16x2(acc[nt][0] * i0, acc[nt][1] * i0));
      st_u32(a + 8 * CTX_LD, pack_bf16x2(acc[nt][2] * i1, acc[nt][3] * i1));
    }
  }
  __syncthreads();

  // ---------------------------------------------------------------- pass B: output -----------------------------------------
  // B[k = d][n = e] = ctx[h][d][e] (constant over the pass): fragments in registers, [k-step][n-tile][2]
  uint32_t bq[2][4][2];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int pr = 0; pr < 2; ++pr)
      ldsm_x4_t(ctxs + (uint32_t)((h * 32 + ks * 16 + (lmat & 1) * 8 + li) * CTX_LD + ((2 * pr + (lmat >> 1)) * 8) * 2), bq[ks][2 * pr][0],
                bq[ks][2 * pr][1], bq[ks][2 * pr + 1][0], bq[ks][2 * pr + 1][1]);
  auto load_q = [&](int tile, int buf) {
    const uint32_t dst0 = tiles + (uint32_t)(buf * TN * Q_LD);
    for (int i = t; i < TN * 16; i += 256) {
      const int r = i >> 4, c16 = i & 15, n = tile * TN + r;
      const uint32_t dst = dst0 + (uint32_t)(r * Q_LD + c16 * 16);
      if (n < N) cp_async16(dst, base + (long)n * 384 + c16 * 8);
      else sts128(dst, make_uint4(0, 0, 0, 0));
    }
    cp_async_commit();
  };
  const float scale = rsqrtf(32.f);
  const int th = (w & 1) * 32;               // token half of this warp's 32 x 32 output block
  load_q(0, 0);
  for (int tile = 0; tile < ntiles; ++tile) {
    const int buf = tile & 1;
    const uint32_t tb = tiles + (uint32_t)(buf * TN * Q_LD);
    if (tile + 1 < ntiles) { load_q(tile + 1, buf ^ 1); cp_async_wait<1>(); }
    else cp_async_wait<0>();
    __syncthreads();
    // softmax over the 32 channels of (token r, head hh), times scale, in place
    {
      const int r = t & 63, hh = t >> 6;
      const uint32_t a = tb + (uint32_t)(r * Q_LD + hh * 64);
      float q[32];
      float qm = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 u = lds128(a + j * 16);
        unpack8(u, q + 8 * j);
      }
#pragma unroll
      for (int d = 0; d < 32; ++d) qm = fmaxf(qm, q[d]);
      float qs = 0.f;
#pragma unroll
      for (int d = 0; d < 32; ++d) { q[d] = __expf(q[d] - qm); qs += q[d]; }
      const float qn = scale / qs;
#pragma unroll
      for (int d = 0; d < 32; ++d) q[d] *= qn;
#pragma unroll
      for (int j = 0; j < 4; ++j) sts128(a + j * 16, pack8(q + 8 * j));
    }
    __syncthreads();
    // out[th + 0..31][h*32 + 0..31] = q ctx, written back in place (this warp is the only reader / writer of the block)
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      float o[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        uint32_t a[4];
        // A[m = token][k = d] = q[token][d]: matrices (tok lo, d lo), (tok hi, d lo), (tok lo, d hi), (tok hi, d hi)
        ldsm_x4(tb + (uint32_t)((th + mt * 16 + (lmat & 1) * 8 + li) * Q_LD + (h * 32 + ks * 16 + (lmat >> 1) * 8) * 2), a[0], a[1], a[2], a[3]);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma16816(o[nt], a, bq[ks][nt][0], bq[ks][nt][1]);
      }
      __syncwarp();
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const uint32_t a = tb + (uint32_t)((th + mt * 16 + g) * Q_LD + (h * 32 + nt * 8 + 2 * tq) * 2);
        st_u32(a, pack_bf16x2(o[nt][0], o[nt][1]));
        st_u32(a + 8 * Q_LD, pack_bf16x2(o[nt][2], o[nt][3]));
      }
    }
    __syncthreads();
    // coalesced store: 16 lanes x 16 B per 256-byte output row
    for (int i = t; i < TN * 16; i += 256) {
      const int r = i >> 4, c16 = i & 15, n = tile * TN + r;
      if (n < N) *reinterpret_cast<uint4*>(out + ((long)b * N + n) * 128 + c16 * 8) = lds128(tb + (uint32_t)(r * Q_LD + c16 * 16));
    }
    __syncthreads();
  }
}

}  // namespace la

int linattn_core_bf16_mma(const void* qkv, void* out, int B, int N, cudaStream_t st) {
  static DeviceOnce attr;
  if (attr.first()) {
    DMN_CUDA_CHECK(cudaFuncSetAttribute(la::linattn_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, la::kSmem));
  }
  DMN_CUDA_CHECK(launch_pdl(la::linattn_mma_kernel, dim3(B), dim3(256), la::kSmem, st, (const bf16*)qkv, (bf16*)out, N));
  count_launch();
  DMN_LAUNCH_CHECK("linattn_mma");
  return 0;
}

}  // namespace dmn
