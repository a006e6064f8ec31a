// conv_tcgen05.cu -- implicit-GEMM convolution on the 5th-generation tensor cores (sm_100a only).
//
// D[m][n] = sum_{tap, ci} A[m + delta(tap)][ci] * W[n][tap][ci]       (bf16 operands, fp32 accumulate in TMEM)
//
// "Shifted-view" formulation.  The NHWC activation tensor is addressed through a FLAT padded pixel index
//   f = img * S + (y + pad) * Wp + (x + pad),   Wp = W + pad, S = (H + pad) * Wp,  pad = 1 for 3x3, 0 for 1x1
// in which one zero column / zero row is shared by neighbouring rows / images, so that every filter tap is a
// constant offset delta = (ky-1)*Wp + (kx-1) in f.  A CTA owns 256 consecutive flat positions (two 128-row
// accumulators).  Its input window [m0 - halo, m0 + 256 + halo) is loaded ONCE per 32-channel pass into
// shared memory by the producer warps -- which also apply the fused prologue (GroupNorm-apply of the producing
// conv, SiLU, time-embedding add; zero padding stays zero) -- in the UMMA K-major, no-swizzle canonical layout
//   A_smem[kchunk (8 channels = 16 B)][pixel]      (LBO = PA*16 B between k-chunks, SBO = 128 B between 8-row groups)
// so the operand of tap t is the same buffer with the start address advanced by delta*16 B: nine MMAs per
// k-step read one resident tile; nothing is re-fetched from L2.  Weights are pre-blocked on the host into
// [n_tile][pass][tap][kchunk][n] 16-byte items and streamed with 1-D bulk async copies (cp.async.bulk ->
// UBLKCP, the TMA engine's non-tensor mode) through a 6-stage mbarrier ring.
//
// Warp roles (192 threads): warps 0-3 operand producers, then epilogue (TMEM -> registers -> +bias, +residual,
// GroupNorm statistics, bf16 -> global); warp 4 weight loader; warp 5 TMEM allocator + single-thread MMA issuer.
// Resources per CTA: <= 100 KB shared memory and 256 TMEM columns, so two CTAs share an SM and one CTA's
// epilogue / operand ramp overlaps the other's MMA main loop.
//
// Garbage rows: flat positions that fall on a pad column/row are computed and discarded (1 - HW/S of the MMA
// work: 6 % at 32x32, 11 % at 16x16, 21 % at 8x8, 36 % at 4x4).
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "ops.h"

namespace dmn {
namespace tc {

constexpr int kThreads = 192;
constexpr int kProducerThreads = 128;
constexpr int kMT = 2;                 // 128-row accumulators per CTA
constexpr int kMcta = 128 * kMT;
constexpr int kCk = 32;                // channels per pass (4 k-chunks of 8)
constexpr int kStagesB = 6;
constexpr int kMaxItems = 13;          // 16-byte operand items per producer thread per pass
constexpr int kNimgMax = 20;           // images a 256-position window may touch
constexpr int kGroupsMax = 32;         // GroupNorm groups of the prologue
constexpr int kOgMax = 16;             // output-statistics groups per N tile

struct Params {
  ConvP c;
  int S, Wp, pad, halo, P, PA, HW;
  int ksize, ntap, NT, n_pass;
  long total_flat;
  uint32_t lbo_a, sbo_a, lbo_b, sbo_b;   // bytes
  uint32_t tmem_cols;
  int cpg_in, cpg_out;
  float inv_cnt_in;
};

// ---------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug traps (CUDA error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, "
      "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor, K-major, SWIZZLE_NONE (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout_type [61,64) = 0)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) |
         (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 [4,6)=1, a/b_format BF16 [7,10)/[10,13)=1,
// a/b K-major (bits 15/16 = 0), N>>3 at [17,23), M>>4 at [24,29)
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct FlatPos {
  int img, pix;     // pix < 0: padding / out of range
};
__device__ __forceinline__ FlatPos decode(long f, const Params& p) {
  FlatPos r;
  r.img = -1;
  r.pix = -1;
  if (f < 0 || f >= p.total_flat) return r;
  const int img = (int)(f / p.S);
  const int rem = (int)(f - (long)img * p.S);
  const int row = rem / p.Wp, col = rem - row * p.Wp;
  r.img = img;
  if (row >= p.pad && col >= p.pad) r.pix = (row - p.pad) * (p.Wp - p.pad) + (col - p.pad);
  return r;
}

// ---------------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 2) conv_tcgen05_kernel(const Params p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const long m0 = (long)blockIdx.x * kMcta;
  const int n_tile = blockIdx.y;
  const int n0 = n_tile * p.NT;

  // ---- shared memory carve-up ----
  const uint32_t a_bytes = 4u * p.PA * 16u;           // one A buffer (4 k-chunks)
  const uint32_t b_bytes = 4u * p.NT * 16u;           // one B stage
  uint8_t* sA = smem;
  uint8_t* sB = sA + 2 * a_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + kStagesB * b_bytes);
  uint64_t* full_b = bars;
  uint64_t* empty_b = bars + kStagesB;
  uint64_t* full_a = bars + 2 * kStagesB;
  uint64_t* empty_a = full_a + 2;
  uint64_t* acc_full = empty_a + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  float* s_bias = reinterpret_cast<float*>(tmem_slot + 2);
  float2* s_gn = reinterpret_cast<float2*>(s_bias + 128);               // [kNimgMax][kGroupsMax] (mean, rstd)
  float* s_ost = reinterpret_cast<float*>(s_gn + kNimgMax * kGroupsMax);  // [kNimgMax][kOgMax][2]

  // images touched by this CTA's window
  long f_lo = m0 - p.halo;
  if (f_lo < 0) f_lo = 0;
  const int img_lo = (int)(f_lo / p.S);

  // ---- one-time setup ----
  if (tid == 0) {
    for (int i = 0; i < kStagesB; ++i) { mbar_init(smem_u32(&full_b[i]), 1); mbar_init(smem_u32(&empty_b[i]), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&full_a[i]), kProducerThreads); mbar_init(smem_u32(&empty_a[i]), 1); }
    mbar_init(smem_u32(acc_full), 1);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc(smem_u32(tmem_slot), p.tmem_cols);
  // zero both operand buffers once: padding positions are never written again
  for (uint32_t i = tid; i < 2 * a_bytes / 16; i += kThreads) reinterpret_cast<uint4*>(sA)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < p.NT; i += kThreads) s_bias[i] = p.c.bias ? p.c.bias[n0 + i] : 0.f;
  for (int i = tid; i < kNimgMax * kOgMax * 2; i += kThreads) s_ost[i] = 0.f;
  if (p.c.pro & PRO_GN) {
    for (int i = tid; i < kNimgMax * p.c.pgroups; i += kThreads) {
      const int il = i / p.c.pgroups, g = i - il * p.c.pgroups;
      const int img = img_lo + il;
      float mean = 0.f, rstd = 0.f;
      if (img < p.c.B) gn_mean_rstd(p.c.pstats + ((long)img * p.c.pgroups + g) * 2, p.inv_cnt_in, kGnEps, mean, rstd);
      s_gn[il * kGroupsMax + g] = make_float2(mean, rstd);
    }
  }
  fence_proxy_async();      // the zero fill must be visible to the tensor-core (async) proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // =============================== operand producers ===============================
    const int kc = tid & 3;
    int goff[kMaxItems];      // img*HW + pix, or -1
    int imgl[kMaxItems];
#pragma unroll
    for (int j = 0; j < kMaxItems; ++j) {
      const int pixel = (tid >> 2) + 32 * j;
      goff[j] = -1;
      imgl[j] = 0;
      if (pixel < p.P) {
        const FlatPos fp = decode(m0 - p.halo + pixel, p);
        if (fp.pix >= 0) {
          goff[j] = fp.img * p.HW + fp.pix;
          imgl[j] = fp.img - img_lo;
        }
      }
    }
    const bf16* src1 = (const bf16*)p.c.src1;
    const bf16* src2 = (const bf16*)p.c.src2;
    const float* temb_base = nullptr;
    if (p.c.pro & PRO_TEMB) temb_base = p.c.temb + (p.c.d_row ? (long)(*p.c.d_row) * p.c.temb_rstride : 0);

    for (int c = 0; c < p.n_pass; ++c) {
      const int buf = c & 1;
      mbar_wait(smem_u32(&empty_a[buf]), ((c >> 1) & 1) ^ 1);
      const int cb = c * kCk + kc * 8;        // first channel of this thread's k-chunk
      const bf16* src;
      int Cs, cofs;
      if (cb < p.c.C1) { src = src1; Cs = p.c.C1; cofs = cb; }
      else { src = src2; Cs = p.c.C2; cofs = cb - p.c.C1; }
      uint4 raw[kMaxItems];
#pragma unroll
      for (int j = 0; j < kMaxItems; ++j)
        if (goff[j] >= 0) raw[j] = __ldg(reinterpret_cast<const uint4*>(src + (long)goff[j] * Cs + cofs));
      uint8_t* dstbase = sA + buf * a_bytes + (uint32_t)kc * p.lbo_a;
      if (p.c.pro == PRO_NONE) {
#pragma unroll
        for (int j = 0; j < kMaxItems; ++j)
          if (goff[j] >= 0) *reinterpret_cast<uint4*>(dstbase + ((tid >> 2) + 32 * j) * 16) = raw[j];
      } else {
        float ga[8], be[8], te[8];
        {
          const float4 g0 = *reinterpret_cast<const float4*>(p.c.pgamma + cb), g1 = *reinterpret_cast<const float4*>(p.c.pgamma + cb + 4);
          const float4 b0 = *reinterpret_cast<const float4*>(p.c.pbeta + cb), b1 = *reinterpret_cast<const float4*>(p.c.pbeta + cb + 4);
          ga[0] = g0.x; ga[1] = g0.y; ga[2] = g0.z; ga[3] = g0.w; ga[4] = g1.x; ga[5] = g1.y; ga[6] = g1.z; ga[7] = g1.w;
          be[0] = b0.x; be[1] = b0.y; be[2] = b0.z; be[3] = b0.w; be[4] = b1.x; be[5] = b1.y; be[6] = b1.z; be[7] = b1.w;
        }
        const bool temb_shared = (p.c.pro & PRO_TEMB) && p.c.temb_bstride == 0;
#pragma unroll
        for (int e = 0; e < 8; ++e) te[e] = 0.f;
        if (temb_shared) {
          const float4 t0 = *reinterpret_cast<const float4*>(temb_base + cb), t1 = *reinterpret_cast<const float4*>(temb_base + cb + 4);
          te[0] = t0.x; te[1] = t0.y; te[2] = t0.z; te[3] = t0.w; te[4] = t1.x; te[5] = t1.y; te[6] = t1.z; te[7] = t1.w;
        }
        const int g = cb / p.cpg_in;
#pragma unroll
        for (int j = 0; j < kMaxItems; ++j) {
          if (goff[j] < 0) continue;
          const float2 mr = s_gn[imgl[j] * kGroupsMax + g];
          if ((p.c.pro & PRO_TEMB) && !temb_shared) {
            const float* tp = temb_base + (long)(img_lo + imgl[j]) * p.c.temb_bstride + cb;
            const float4 t0 = *reinterpret_cast<const float4*>(tp), t1 = *reinterpret_cast<const float4*>(tp + 4);
            te[0] = t0.x; te[1] = t0.y; te[2] = t0.z; te[3] = t0.w; te[4] = t1.x; te[5] = t1.y; te[6] = t1.z; te[7] = t1.w;
          }
          float v[8];
          unpack8(raw[j], v);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float t = (v[e] - mr.x) * mr.y;
            t = fmaf(t, ga[e], be[e]);
            if (p.c.pro & PRO_SILU) t = silu_fast(t);
            v[e] = t + te[e];
          }
          *reinterpret_cast<uint4*>(dstbase + ((tid >> 2) + 32 * j) * 16) = pack8(v);
        }
      }
      fence_proxy_async();
      mbar_arrive(smem_u32(&full_a[buf]));
    }

    // =============================== epilogue ===============================
    mbar_wait(smem_u32(acc_full), 0);
    tc_fence_after();
    bf16* out = (bf16*)p.c.out;
    const bf16* res = (const bf16*)p.c.res;
#pragma unroll 1
    for (int mt = 0; mt < kMT; ++mt) {
      const int row = warp * 32 + lane;
      const FlatPos fp = decode(m0 + mt * 128 + row, p);
      const bool valid = fp.pix >= 0;
      const long orow = valid ? ((long)fp.img * p.HW + fp.pix) * p.c.Cout + n0 : 0;
      // statistics bookkeeping: is the warp inside one image?
      const int my_img = valid ? fp.img - img_lo : -1;
      const int ref_img = __reduce_max_sync(0xffffffffu, my_img);
      const bool uniform = __all_sync(0xffffffffu, my_img == ref_img || my_img < 0);
      int cur_g = -1;
      float s = 0.f, ss = 0.f;
      auto flush = [&]() {
        if (cur_g < 0) return;
        const int gl = cur_g - n0 / p.cpg_out;   // group index local to this N tile
        if (uniform) {
          const float a = warp_sum(s), b = warp_sum(ss);
          if (lane == 0 && ref_img >= 0) {
            atomicAdd(&s_ost[(ref_img * kOgMax + gl) * 2], a);
            atomicAdd(&s_ost[(ref_img * kOgMax + gl) * 2 + 1], b);
          }
        } else if (valid) {
          atomicAdd(&s_ost[(my_img * kOgMax + gl) * 2], s);
          atomicAdd(&s_ost[(my_img * kOgMax + gl) * 2 + 1], ss);
        }
        s = ss = 0.f;
      };
      for (int ch = 0; ch < p.NT / 32; ++ch) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(mt * p.NT + ch * 32), r);
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + s_bias[ch * 32 + j];
        if (res && valid) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 rr = *reinterpret_cast<const uint4*>(res + orow + ch * 32 + q * 8);
            float rf[8];
            unpack8(rr, rf);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[q * 8 + e] += rf[e];
          }
        }
        if (valid) {
#pragma unroll
          for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(out + orow + ch * 32 + q * 8) = pack8(v + q * 8);
        }
        if (p.c.ostats) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int g = (n0 + ch * 32 + q * 8) / p.cpg_out;
            if (g != cur_g) { flush(); cur_g = g; }
            if (valid) {
#pragma unroll
              for (int e = 0; e < 8; ++e) { s += v[q * 8 + e]; ss += v[q * 8 + e] * v[q * 8 + e]; }
            }
          }
        }
      }
      if (p.c.ostats) flush();
    }
    tc_fence_before();
  } else if (warp == 4) {
    // =============================== weight loader ===============================
    if (lane == 0) {
      const uint8_t* wsrc = (const uint8_t*)p.c.w + (size_t)n_tile * p.n_pass * p.ntap * b_bytes;
      const int total = p.n_pass * p.ntap;
      for (int s = 0; s < total; ++s) {
        const int st = s % kStagesB;
        mbar_wait(smem_u32(&empty_b[st]), ((s / kStagesB) & 1) ^ 1);
        mbar_arrive_expect_tx(smem_u32(&full_b[st]), b_bytes);
        bulk_g2s(smem_u32(sB + st * b_bytes), wsrc + (size_t)s * b_bytes, b_bytes, smem_u32(&full_b[st]));
      }
    }
  } else {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(128, p.NT);
      const int kh = p.ksize >> 1;
      int s = 0;
      for (int c = 0; c < p.n_pass; ++c) {
        const int buf = c & 1;
        mbar_wait(smem_u32(&full_a[buf]), (c >> 1) & 1);
        tc_fence_after();
        const uint32_t abase = smem_u32(sA + buf * a_bytes);
        for (int t = 0; t < p.ntap; ++t, ++s) {
          const int st = s % kStagesB;
          mbar_wait(smem_u32(&full_b[st]), (s / kStagesB) & 1);
          tc_fence_after();
          const int ky = t / p.ksize, kx = t - ky * p.ksize;
          const int delta = (ky - kh) * p.Wp + (kx - kh);
          const uint32_t bbase = smem_u32(sB + st * b_bytes);
#pragma unroll
          for (int k16 = 0; k16 < 2; ++k16) {
            const uint64_t bdesc = make_desc(bbase + 2 * k16 * p.lbo_b, p.lbo_b, p.sbo_b);
#pragma unroll
            for (int mt = 0; mt < kMT; ++mt) {
              const uint64_t adesc =
                  make_desc(abase + 2 * k16 * p.lbo_a + (uint32_t)(p.halo + mt * 128 + delta) * 16u, p.lbo_a, p.sbo_a);
              umma_bf16(tmem_base + (uint32_t)(mt * p.NT), adesc, bdesc, idesc, (c | t | k16) ? 1u : 0u);
            }
          }
          umma_commit(smem_u32(&empty_b[st]));     // frees the weight stage once these MMAs retire
        }
        umma_commit(smem_u32(&empty_a[buf]));      // frees the operand buffer
      }
      umma_commit(smem_u32(acc_full));
    }
  }

  __syncthreads();
  // flush the CTA's GroupNorm statistics
  if (p.c.ostats) {
    const int og_tile = (p.NT + p.cpg_out - 1) / p.cpg_out;
    for (int i = tid; i < kNimgMax * og_tile; i += kThreads) {
      const int il = i / og_tile, gl = i - il * og_tile;
      const int img = img_lo + il;
      const float a = s_ost[(il * kOgMax + gl) * 2], b = s_ost[(il * kOgMax + gl) * 2 + 1];
      if (img < p.c.B && (a != 0.f || b != 0.f)) {
        float* dst = p.c.ostats + ((long)img * p.c.ogroups + n0 / p.cpg_out + gl) * 2;
        atomicAdd(dst, a);
        atomicAdd(dst + 1, b);
      }
    }
  }
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

static int pick_nt(int cout) { return cout % 128 == 0 ? 128 : (cout % 64 == 0 ? 64 : (cout % 32 == 0 ? 32 : 0)); }

static bool fill_params(const ConvP& c, Params& p) {
  if (c.mode != CONV_SAME || (c.ksize != 1 && c.ksize != 3)) return false;
  if (c.C1 <= 0 || c.C1 % kCk || c.C2 % kCk) return false;
  p.NT = pick_nt(c.Cout);
  if (!p.NT) return false;
  p.c = c;
  p.ksize = c.ksize;
  p.ntap = c.ksize * c.ksize;
  p.pad = c.ksize == 3 ? 1 : 0;
  p.Wp = c.Win + p.pad;
  p.S = (c.Hin + p.pad) * p.Wp;
  p.HW = c.Hin * c.Win;
  p.halo = c.ksize == 3 ? p.Wp + 1 : 0;
  p.P = kMcta + 2 * p.halo;
  p.PA = p.P;
  while (p.PA % 8 != 2) ++p.PA;
  if ((4 * p.P + kProducerThreads - 1) / kProducerThreads > kMaxItems) return false;
  if (kMcta / p.S + 3 > kNimgMax) return false;
  p.n_pass = (c.C1 + c.C2) / kCk;
  p.total_flat = (long)c.B * p.S;
  p.lbo_a = (uint32_t)p.PA * 16u;
  p.sbo_a = 128u;
  p.lbo_b = (uint32_t)p.NT * 16u;
  p.sbo_b = 128u;
  p.tmem_cols = (uint32_t)(kMT * p.NT);
  if (p.tmem_cols < 32) p.tmem_cols = 32;
  p.cpg_in = 1;
  p.inv_cnt_in = 0.f;
  if (c.pro & PRO_GN) {
    if (c.C2 != 0 || c.pgroups <= 0 || c.pgroups > kGroupsMax || c.C1 % c.pgroups) return false;
    p.cpg_in = c.C1 / c.pgroups;
    if (p.cpg_in % 8) return false;
    p.inv_cnt_in = 1.f / (float)(p.HW * p.cpg_in);
  } else if (c.pro != PRO_NONE) {
    return false;   // SiLU / temb only come together with the GroupNorm apply
  }
  p.cpg_out = 1;
  if (c.ogroups > 0) {
    if (c.Cout % c.ogroups) return false;
    p.cpg_out = c.Cout / c.ogroups;
    if (p.cpg_out % 8) return false;
    if ((p.NT + p.cpg_out - 1) / p.cpg_out > kOgMax) return false;
    if (p.cpg_out < p.NT && p.NT % p.cpg_out) return false;
    if (p.cpg_out > p.NT && p.cpg_out % p.NT) return false;
  }
  return true;
}

static size_t smem_bytes(const Params& p) {
  return (size_t)2 * 4 * p.PA * 16 + (size_t)kStagesB * 4 * p.NT * 16 + (2 * kStagesB + 5) * 8 + 16 + 128 * 4 +
         (size_t)kNimgMax * kGroupsMax * 8 + (size_t)kNimgMax * kOgMax * 2 * 4 + 128;
}

}  // namespace tc

bool conv_tcgen05_supported(const ConvP& c) {
  tc::Params p;
  if (!tc::fill_params(c, p)) return false;
  return tc::smem_bytes(p) <= 113 * 1024;
}

size_t conv_tcgen05_weight_bytes(int mode, int ksize, int cin, int cout) {
  const int taps = (mode == CONV_SAME) ? ksize * ksize : 16;
  return (size_t)cout * cin * taps * 2;
}

// [n_tile][pass][tap][kchunk 0..3][n 0..NT-1][8 channels] bf16
void conv_tcgen05_pack_weights(int mode, int ksize, int cin, int cout, const float* w, void* dst_host) {
  (void)mode;
  const int NT = tc::pick_nt(cout);
  const int taps = ksize * ksize;
  const int n_pass = cin / tc::kCk;
  bf16* dst = (bf16*)dst_host;
  size_t o = 0;
  for (int nt = 0; nt < cout / NT; ++nt)
    for (int c = 0; c < n_pass; ++c)
      for (int t = 0; t < taps; ++t)
        for (int kc = 0; kc < 4; ++kc)
          for (int n = 0; n < NT; ++n)
            for (int e = 0; e < 8; ++e) {
              const int co = nt * NT + n, ci = c * tc::kCk + kc * 8 + e;
              dst[o++] = __float2bfloat16_rn(w[((long)co * cin + ci) * taps + t]);
            }
}

int conv_tcgen05(const ConvP& c, cudaStream_t st) {
  tc::Params p;
  if (!tc::fill_params(c, p)) return fail(-2, "conv_tcgen05: unsupported convolution shape");
  static const bool swap = [] {
    const char* e = getenv("DMN_UMMA_SWAP_LBO_SBO");
    return e && e[0] == '1';
  }();
  if (swap) {
    std::swap(p.lbo_a, p.sbo_a);
    std::swap(p.lbo_b, p.sbo_b);
  }
  const size_t smem = tc::smem_bytes(p);
  static bool attr_set = false;
  if (!attr_set) {
    DMN_CUDA_CHECK(cudaFuncSetAttribute(tc::conv_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
    attr_set = true;
  }
  dim3 grid((unsigned)((p.total_flat + tc::kMcta - 1) / tc::kMcta), (unsigned)(c.Cout / p.NT));
  tc::conv_tcgen05_kernel<<<grid, tc::kThreads, smem, st>>>(p);
  count_launch();
  DMN_LAUNCH_CHECK("conv_tcgen05");
  return 0;
}

}  // namespace dmn
