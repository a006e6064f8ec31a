// conv_tcgen05.cu -- implicit-GEMM convolution on the 5th-generation tensor cores (sm_100a only).
//
// D[m][n] = sum_{tap, ci} A[m + delta(tap)][ci] * W[n][tap][ci]       (bf16 operands, fp32 accumulate in TMEM)
//
// "Shifted-view" formulation.  Every convolution of the U-Net is expressed over a VIRTUAL image addressed by a flat
// pixel index f = img * S + row * Wv + col in which each filter tap is a constant offset delta in f:
//
//   GEO_SAME  3x3 pad 1 : Wv = W+1, S = (H+1)*Wv; row 0 / col 0 of every image block are zero pads SHARED with the
//             (1x1 pad 0)  neighbouring row / image, delta = (ky-1)*Wv + (kx-1).  1x1: Wv = W, S = H*W, delta = 0.
//   GEO_DOWN  Conv k4 s2 p1 == 2x2 stride-1 conv over the shifted space-to-depth image: virtual pixel (u,v) holds the
//             2x2 input patch rows (2u-1, 2u) x cols (2v-1, 2v) as 4*C channels; Wv = W/2+1, delta = du*Wv + dv.
//   GEO_UP    ConvTranspose k4 s2 p1 == four sub-pixel phases (py,px); each phase is a 2x2 conv over the padded input
//             (layout of GEO_SAME 3x3) with its own 4 taps/weights; the phase is folded into the N-tile index.
//   GEO_INIT  7x7 pad 3 stem on the fp32 NCHW sampler state: the kx direction is packed into channels
//             (virtual channel = kx*Cin + ch, 7*Cin <= 32), 3 zero rows shared between images, delta = (ky-3)*W.
//
// A CTA owns 256 consecutive flat positions (two 128-row accumulators).  Its input window is loaded ONCE per
// 32-channel pass into shared memory by the producer warps -- which also apply the fused prologue (GroupNorm-apply of
// the producing conv, SiLU, time-embedding add; padding stays zero) -- in the UMMA K-major, no-swizzle canonical layout
//   A_smem[kchunk (8 channels = 16 B)][pixel]      (LBO = PA*16 B between k-chunks, SBO = 128 B between 8-row groups)
// so the operand of tap t is the same buffer with the start address advanced by delta*16 B: all taps of a k-step read
// one resident tile and nothing is re-fetched from L2.  Weights are pre-blocked on the host into
// [n_tile][pass][tap][kchunk][n] 16-byte items and streamed with 1-D bulk async copies (cp.async.bulk -> UBLKCP, the TMA
// engine's non-tensor mode) through a 6-stage mbarrier ring.
//
// Warp roles (192 threads): warps 0-3 operand producers, then epilogue (TMEM -> registers -> +bias, +residual,
// GroupNorm statistics, bf16 -> global); warp 4 weight loader; warp 5 TMEM allocator + single-thread MMA issuer.
// Resources per CTA: <= 113 KB shared memory and 256 TMEM columns, so two CTAs share an SM and one CTA's
// epilogue / operand ramp overlaps the other's MMA main loop.
//
// Garbage rows: flat positions that fall on a pad column/row are computed and discarded (1 - HW/S of the MMA
// work for 3x3: 6 % at 32x32, 11 % at 16x16, 21 % at 8x8, 36 % at 4x4).
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "ops.h"

namespace dmn {
namespace tc {

enum { GEO_SAME = 0, GEO_DOWN = 1, GEO_UP = 2, GEO_INIT = 3 };

constexpr int kThreads = 192;
constexpr int kProducerThreads = 128;
constexpr int kMTmax = 2;              // 128-row accumulators per CTA: 2, or 1 when two would leave most SMs idle
constexpr int kMcta = 128 * kMTmax;    // table sizing
constexpr int kCk = 32;                // channels per pass (4 k-chunks of 8)
constexpr int kStagesMax = 8;         // weight-ring depth is chosen per launch (4..8) to fit shared memory
constexpr int kMaxItems = 13;          // 16-byte operand items per producer thread per pass (P <= 416)
constexpr int kNimgMax = 20;           // images a 256-position window may touch
constexpr int kGroupsMax = 32;         // GroupNorm groups of the prologue
constexpr int kOgMax = 16;             // output-statistics groups per N tile
constexpr int kSegMax = 4;             // images a warp's 32 consecutive rows may touch (S >= 16)

struct Params {
  ConvP c;
  int geo;
  int S, Wv, pad, halo_lo, P, PA, n_abuf, mt, mcta;
  int H, W, HW;            // input image
  int ksize, ntap, NT, n_pass, tiles_per_phase;
  long total_flat;
  uint32_t lbo_a, sbo_a, lbo_b, sbo_b;   // bytes
  uint32_t tmem_cols;
  int cpg_in, cpg_out, cpg_in_shift, cpg_out_shift, nstage, nt_shift;
  float inv_cnt_in;
  long long* trace;        // debug: per-role clock64 timeline of CTA (trace_cta, 0); null in production
  int trace_cta;
  // GEO_INIT extras
  const float* cls_w;
  const int64_t* classes;
  int pad_class;
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

// debug timeline (DMN_TC_TRACE=1): slot layout documented in tools/trace_conv.py
__device__ long long g_trace[1024];
#define TRACE(slot)                                                                     \
  do {                                                                                  \
    if (p.trace && blockIdx.x == (unsigned)p.trace_cta && blockIdx.y == 0) p.trace[(slot)] = clock64(); \
  } while (0)

// flat virtual position -> (image, virtual row, virtual col); img < 0 when out of range.  32-bit arithmetic only
// (the host guarantees total_flat < 2^30): 64-bit divisions here used to cost ~7 us per CTA.
struct VPos {
  int img, row, col;
};
__device__ __forceinline__ VPos vdecode(int f, const Params& p) {
  VPos r;
  r.img = -1;
  r.row = r.col = 0;
  if (f < 0 || f >= (int)p.total_flat) return r;
  const unsigned uf = (unsigned)f, uS = (unsigned)p.S, uW = (unsigned)p.Wv;
  const unsigned img = uf / uS;
  const unsigned rem = uf - img * uS;
  const unsigned row = rem / uW;
  r.img = (int)img;
  r.row = (int)row;
  r.col = (int)(rem - row * uW);
  return r;
}
// tap offset in flat positions
template <int GEO>
__device__ __forceinline__ int tap_delta(const Params& p, int t, int phase) {
  if (GEO == GEO_SAME) {
    const int kh = p.ksize >> 1, ky = t / p.ksize, kx = t - ky * p.ksize;
    return (ky - kh) * p.Wv + (kx - kh);
  } else if (GEO == GEO_DOWN) {
    return (t >> 1) * p.Wv + (t & 1);
  } else if (GEO == GEO_UP) {
    const int py = phase >> 1, px = phase & 1, a = t >> 1, b = t & 1;
    const int dy = py ? (a ? 0 : 1) : (a ? -1 : 0);
    const int dx = px ? (b ? 0 : 1) : (b ? -1 : 0);
    return dy * p.Wv + dx;
  } else {
    return (t - 3) * p.Wv;
  }
}
// waits of the non-critical roles back off so that their polling does not steal issue slots from the MMA thread
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  int n = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(64);
    if ((++n & 1023) == 0 && clock64() - t0 > 4000000000LL) __trap();
  }
}

// ---------------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------------
template <int GEO>
__global__ void __launch_bounds__(kThreads, 2) conv_tcgen05_kernel(const Params p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int m0 = (int)blockIdx.x * p.mcta;
  const int n_tile = blockIdx.y;
  const int phase = (GEO == GEO_UP) ? n_tile / p.tiles_per_phase : 0;
  const int n0 = (GEO == GEO_UP ? n_tile % p.tiles_per_phase : n_tile) * p.NT;
  const int nst = p.nstage;

  // ---- shared memory carve-up ----
  const uint32_t a_bytes = 4u * p.PA * 16u;           // one A buffer (4 k-chunks)
  const uint32_t b_bytes = 4u * p.NT * 16u;           // one B stage
  uint8_t* sA = smem;
  uint8_t* sB = sA + p.n_abuf * a_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + nst * b_bytes);
  uint64_t* full_b = bars;
  uint64_t* empty_b = bars + kStagesMax;
  uint64_t* full_a = bars + 2 * kStagesMax;
  uint64_t* empty_a = full_a + 2;
  uint64_t* acc_full = empty_a + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  int* s_delta = reinterpret_cast<int*>(tmem_slot + 2);                                     // [16] tap offsets
  float* s_bias = reinterpret_cast<float*>(s_delta + 16);
  float2* s_gn = reinterpret_cast<float2*>(s_bias + 128);                                   // [kNimgMax][kGroupsMax] (mean, rstd)
  float* s_part = reinterpret_cast<float*>(s_gn + kNimgMax * kGroupsMax);                   // [8 (mt,warp)][kSegMax][kOgMax][2]
  int* s_partkey = reinterpret_cast<int*>(s_part + 8 * kSegMax * kOgMax * 2);               // [8][kSegMax] image key or -1
  int* s_opix = s_partkey + 8 * kSegMax;                                                    // [256] output pixel or -1
  int* s_oimg = s_opix + kMcta;                                                             // [256] image - img_lo (clamped)
  int* s_pix = s_oimg + kMcta;                                                              // [P] operand source or -1
  int* s_pimg = s_pix + p.P;                                                                // [P] image - img_lo

  int f_lo = m0 - p.halo_lo;
  if (f_lo < 0) f_lo = 0;
  const int img_lo = f_lo / p.S;

  if (tid == 0) TRACE(0);
  // ---- one-time setup ----
  if (warp == 4) {          // one lane per barrier
    if (lane < nst) { mbar_init(smem_u32(&full_b[lane]), 1); mbar_init(smem_u32(&empty_b[lane]), 1); }
    if (lane >= 8 && lane < 10) { mbar_init(smem_u32(&full_a[lane - 8]), kProducerThreads); mbar_init(smem_u32(&empty_a[lane - 8]), 1); }
    if (lane == 10) mbar_init(smem_u32(acc_full), 1);
    fence_barrier_init();
    if (lane >= 16 && lane < 16 + p.ntap) s_delta[lane - 16] = tap_delta<GEO>(p, lane - 16, phase);
  }
  if (warp == 5) tmem_alloc(smem_u32(tmem_slot), p.tmem_cols);
  // zero the operand buffers once: padding positions are never written again (GEO_DOWN writes its own zeros)
  for (uint32_t i = tid; i < p.n_abuf * a_bytes / 16; i += kThreads) reinterpret_cast<uint4*>(sA)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < p.NT; i += kThreads) s_bias[i] = p.c.bias ? p.c.bias[n0 + i] : 0.f;
  for (int i = tid; i < 8 * kSegMax; i += kThreads) s_partkey[i] = -1;
  // operand source table (one decode per window pixel instead of one per thread-item)
  if (GEO != GEO_INIT) {
    for (int pixel = tid; pixel < p.P; pixel += kThreads) {
      const VPos v = vdecode(m0 - p.halo_lo + pixel, p);
      int g = -1, il = 0;
      if (v.img >= 0) {
        if (GEO == GEO_DOWN) {
          g = (v.img << 14) | (v.row << 7) | v.col;
        } else if (v.row >= p.pad && v.col >= p.pad) {
          g = v.img * p.HW + (v.row - p.pad) * p.W + (v.col - p.pad);
          il = v.img - img_lo;
        }
      }
      s_pix[pixel] = g;
      s_pimg[pixel] = il;
    }
  }
  // output row table
  for (int r = tid; r < p.mcta; r += kThreads) {
    const VPos v = vdecode(m0 + r, p);
    bool valid = v.img >= 0;
    int opix = -1;
    if (valid) {
      if (GEO == GEO_SAME) {
        valid = v.row >= p.pad && v.col >= p.pad;
        opix = v.img * p.HW + (v.row - p.pad) * p.W + (v.col - p.pad);
      } else if (GEO == GEO_DOWN) {
        valid = v.row < (p.H >> 1) && v.col < (p.W >> 1);
        opix = v.img * (p.HW >> 2) + v.row * (p.W >> 1) + v.col;
      } else if (GEO == GEO_UP) {
        valid = v.row >= 1 && v.col >= 1;
        opix = v.img * (p.HW << 2) + (2 * (v.row - 1) + (phase >> 1)) * (2 * p.W) + 2 * (v.col - 1) + (phase & 1);
      } else {
        valid = v.row >= 3;
        opix = v.img * p.HW + (v.row - 3) * p.W + v.col;
      }
    }
    int il = (v.img >= 0 ? v.img : p.c.B - 1) - img_lo;
    il = il < 0 ? 0 : (il >= kNimgMax ? kNimgMax - 1 : il);
    s_opix[r] = valid ? opix : -1;
    s_oimg[r] = il;
  }
  if (GEO == GEO_SAME && (p.c.pro & PRO_GN)) {
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
  if (tid == 0) TRACE(1);

  if (warp < 4) {
    // =============================== operand producers ===============================
    const int kc = tid & 3;
    if (GEO == GEO_INIT) {
      // one pass: virtual channel vc = kx*Cin + ch  (7*Cin <= 32); source is the fp32 NCHW sampler state
      const float* x = (const float*)p.c.src1;
      const int Cin = p.c.C1;
      for (int pixel = tid >> 2; pixel < p.P; pixel += 32) {
        const VPos v = vdecode(m0 - p.halo_lo + pixel, p);
        if (v.img < 0 || v.row < 3) continue;          // shared zero rows
        const int iy = v.row - 3;
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int vc = kc * 8 + e;
          const int kx = vc / Cin, ch = vc - kx * Cin;
          const int ix = v.col + kx - 3;
          f[e] = (kx < 7 && ix >= 0 && ix < p.W) ? __ldg(x + (((long)v.img * Cin + ch) * p.H + iy) * p.W + ix) : 0.f;
        }
        *reinterpret_cast<uint4*>(sA + (uint32_t)kc * p.lbo_a + pixel * 16) = pack8(f);
      }
      fence_proxy_async();
      mbar_arrive(smem_u32(&full_a[0]));
    } else {
      // per-thread item table: smem pixel (tid>>2) + 32 j, k-chunk kc
      int goff[kMaxItems];      // SAME/UP: img*HW + pix (or -1); DOWN: packed (img, u, v) (or -1)
      int imgl[kMaxItems];
#pragma unroll
      for (int j = 0; j < kMaxItems; ++j) {
        const int pixel = (tid >> 2) + 32 * j;
        goff[j] = -1;
        imgl[j] = 0;
        if (pixel < p.P) {
          goff[j] = s_pix[pixel];
          imgl[j] = s_pimg[pixel];
        }
      }
      const bf16* src1 = (const bf16*)p.c.src1;
      const bf16* src2 = (const bf16*)p.c.src2;
      const float* temb_base = nullptr;
      if (GEO == GEO_SAME && (p.c.pro & PRO_TEMB)) temb_base = p.c.temb + (p.c.d_row ? (long)(*p.c.d_row) * p.c.temb_rstride : 0);

      for (int c = 0; c < p.n_pass; ++c) {
        const int buf = c & 1;
        mbar_wait_relaxed(smem_u32(&empty_a[buf]), ((c >> 1) & 1) ^ 1);
        if (tid == 0 && c < 64) TRACE(16 + 4 * c);
        int cb = c * kCk + kc * 8;        // first (virtual) channel of this thread's k-chunk
        int sy = 0, sx = 0;
        if (GEO == GEO_DOWN) {            // virtual channel = sub * C + ci, sub = sy*2 + sx
          const int sub = cb / p.c.C1;
          cb -= sub * p.c.C1;
          sy = sub >> 1;
          sx = sub & 1;
        }
        const bf16* src;
        int Cs, cofs;
        if (cb < p.c.C1) { src = src1; Cs = p.c.C1; cofs = cb; }
        else { src = src2; Cs = p.c.C2; cofs = cb - p.c.C1; }
        uint4 raw[kMaxItems];
#pragma unroll
        for (int j = 0; j < kMaxItems; ++j) {
          raw[j] = make_uint4(0, 0, 0, 0);
          if (goff[j] >= 0) {
            if (GEO == GEO_DOWN) {
              const int img = goff[j] >> 14, iy = 2 * ((goff[j] >> 7) & 127) - 1 + sy, ix = 2 * (goff[j] & 127) - 1 + sx;
              if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W)
                raw[j] = __ldg(reinterpret_cast<const uint4*>(src + ((long)img * p.HW + iy * p.W + ix) * Cs + cofs));
            } else {
              raw[j] = __ldg(reinterpret_cast<const uint4*>(src + (long)goff[j] * Cs + cofs));
            }
          }
        }
        uint8_t* dstbase = sA + buf * a_bytes + (uint32_t)kc * p.lbo_a;
        if (GEO != GEO_SAME || p.c.pro == PRO_NONE) {
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
          const int g = cb >> p.cpg_in_shift;
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
            const float sc = mr.y, sh = -mr.x * mr.y;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              float t = fmaf(v[e], sc, sh);
              t = fmaf(t, ga[e], be[e]);
              if (p.c.pro & PRO_SILU) t = silu_fast(t);
              v[e] = t + te[e];
            }
            *reinterpret_cast<uint4*>(dstbase + ((tid >> 2) + 32 * j) * 16) = pack8(v);
          }
        }
        fence_proxy_async();
        mbar_arrive(smem_u32(&full_a[buf]));
        if (tid == 0 && c < 64) TRACE(17 + 4 * c);
      }
    }

    // =============================== epilogue ===============================
    if (tid == 0) TRACE(2);
    mbar_wait_relaxed(smem_u32(acc_full), 0);
    tc_fence_after();
    if (tid == 0) TRACE(3);
    bf16* out = (bf16*)p.c.out;
    const bf16* res = (const bf16*)p.c.res;
    // The operand / weight buffers are dead once the accumulators are complete: reuse them as a staging tile so that the
    // global stores are full 256-byte rows (row stride padded by 16 B => conflict-free 16-byte shared stores).
    const uint32_t rs = (uint32_t)p.NT * 2u + 16u;
#pragma unroll 1
    for (int mt = 0; mt < p.mt; ++mt) {
      uint8_t* stage = smem + (uint32_t)mt * 128u * rs;
      const int row = mt * 128 + warp * 32 + lane;
      const int opix = s_opix[row];
      const bool valid = opix >= 0;
      const long orow = valid ? (long)opix * p.c.Cout + n0 : 0;
      const float* cls_row = nullptr;
      if (GEO == GEO_INIT && p.cls_w && valid)
        cls_row = p.cls_w + (long)(p.classes ? (int)p.classes[img_lo + s_oimg[row]] : p.pad_class) * p.c.Cout + n0;
      // statistics: segmented warp reduction keyed by image (rows of a warp are consecutive flat positions, so the key is
      // non-decreasing); the last lane of every segment stores the partial into a slot owned by (mt, warp, segment):
      // no atomics, fixed summation order => deterministic
      const int key = s_oimg[row];
      unsigned segmask = 0;
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        const int k2 = __shfl_up_sync(0xffffffffu, key, 1 << i);
        if (lane >= (1 << i) && k2 == key) segmask |= 1u << i;
      }
      const int knext = __shfl_down_sync(0xffffffffu, key, 1);
      const bool tail = (lane == 31) || (knext != key);
      const unsigned tails = __ballot_sync(0xffffffffu, tail);
      int seg = __popc(tails & ((1u << lane) - 1u));
      seg = seg < kSegMax ? seg : kSegMax - 1;
      // per-thread partial sums per 16-channel pair (statistics groups have >= 16 channels): 16 independent chains
      float sa[8], qa[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) sa[i] = qa[i] = 0.f;
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        if (ch < (p.NT >> 5)) {
          uint32_t r[32];
          if (tid == 0) TRACE(300 + (mt * 4 + ch) * 4);
          tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(mt * p.NT + ch * 32), r);
          if (tid == 0) TRACE(301 + (mt * 4 + ch) * 4);
          float vv[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) vv[j] = __uint_as_float(r[j]) + s_bias[ch * 32 + j];
          if (GEO == GEO_INIT && cls_row) {
#pragma unroll
            for (int j = 0; j < 32; ++j) vv[j] += __ldg(cls_row + ch * 32 + j);
          }
          if (res && valid) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint4 rr = *reinterpret_cast<const uint4*>(res + orow + ch * 32 + q * 8);
              float rf[8];
              unpack8(rr, rf);
#pragma unroll
              for (int e = 0; e < 8; ++e) vv[q * 8 + e] += rf[e];
            }
          }
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<uint4*>(stage + (uint32_t)(warp * 32 + lane) * rs + (uint32_t)(ch * 64 + q * 16)) = pack8(vv + q * 8);
          if (tid == 0) TRACE(302 + (mt * 4 + ch) * 4);
          if (p.c.ostats && valid) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              sa[ch * 2 + (j >> 4)] += vv[j];
              qa[ch * 2 + (j >> 4)] = fmaf(vv[j], vv[j], qa[ch * 2 + (j >> 4)]);
            }
          }
          if (tid == 0) TRACE(303 + (mt * 4 + ch) * 4);
        }
      }
      if (tid == 0) TRACE(400 + mt * 4);
      if (p.c.ostats) {
        // one segmented scan for all 16 partials (independent shuffle chains), then the segment tails combine the pairs into
        // groups and store them into the slot owned by (mt, warp, segment): no atomics, fixed order => deterministic
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          const bool take = segmask & (1u << i);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float ta = __shfl_up_sync(0xffffffffu, sa[k], 1 << i), tb = __shfl_up_sync(0xffffffffu, qa[k], 1 << i);
            if (take) { sa[k] += ta; qa[k] += tb; }
          }
        }
        if (tail) {
          const int sidx = (mt * 4 + warp) * kSegMax + seg;
          float* slot = s_part + sidx * kOgMax * 2;
          s_partkey[sidx] = key;
          const int sh = p.cpg_out_shift;
          if (sh <= 4) {
#pragma unroll
            for (int g = 0; g < 8; ++g) { slot[g * 2] = sa[g]; slot[g * 2 + 1] = qa[g]; }
          } else if (sh == 5) {
#pragma unroll
            for (int g = 0; g < 4; ++g) { slot[g * 2] = sa[2 * g] + sa[2 * g + 1]; slot[g * 2 + 1] = qa[2 * g] + qa[2 * g + 1]; }
          } else if (sh == 6) {
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              slot[g * 2] = (sa[4 * g] + sa[4 * g + 1]) + (sa[4 * g + 2] + sa[4 * g + 3]);
              slot[g * 2 + 1] = (qa[4 * g] + qa[4 * g + 1]) + (qa[4 * g + 2] + qa[4 * g + 3]);
            }
          } else {
            slot[0] = ((sa[0] + sa[1]) + (sa[2] + sa[3])) + ((sa[4] + sa[5]) + (sa[6] + sa[7]));
            slot[1] = ((qa[0] + qa[1]) + (qa[2] + qa[3])) + ((qa[4] + qa[5]) + (qa[6] + qa[7]));
          }
        }
      }
      // copy-out: every thread hands ITS OWN staged row (NT*2 contiguous bytes in shared and in global memory) to the bulk
      // copy engine: no barrier, no copy loop, full-line global writes
      if (tid == 0) TRACE(401 + mt * 4);
      fence_proxy_async();
      if (tid == 0) TRACE(402 + mt * 4);
      if (valid) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + orow),
                     "r"(smem_u32(stage + (uint32_t)(warp * 32 + lane) * rs)), "r"((uint32_t)p.NT * 2u)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      if (tid == 0) TRACE(403 + mt * 4);
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (tid == 0) TRACE(408);
    if (tid == 0) TRACE(4);
    tc_fence_before();
  } else if (warp == 4) {
    // =============================== weight loader ===============================
    if (lane == 0) {
      const uint8_t* wsrc = (const uint8_t*)p.c.w + (size_t)n_tile * p.n_pass * p.ntap * b_bytes;
      const int total = p.n_pass * p.ntap;
      int st = 0, ph = 1;
      for (int s = 0; s < total; ++s) {
        mbar_wait_relaxed(smem_u32(&empty_b[st]), ph);
        mbar_arrive_expect_tx(smem_u32(&full_b[st]), b_bytes);
        bulk_g2s(smem_u32(sB + st * b_bytes), wsrc + (size_t)s * b_bytes, b_bytes, smem_u32(&full_b[st]));
        if (s == 0) TRACE(5);
        if (s == total - 1) TRACE(6);
        if (++st == nst) { st = 0; ph ^= 1; }
      }
    }
  } else {
    // =============================== MMA issuer ===============================
    // One thread feeds the tensor core; its instruction stream is the critical path, so everything that does not depend on
    // the stage is hoisted: descriptor high words, LBO fields, the four (k16, mt) operand offsets.
    if (lane == 0) {
      const uint32_t idesc = make_idesc(128, p.NT);
      const uint32_t hi_a = (p.sbo_a >> 4) | (1u << 14), hi_b = (p.sbo_b >> 4) | (1u << 14);
      const uint32_t lbo_a_f = ((p.lbo_a >> 4) & 0x3FFFu) << 16, lbo_b_f = ((p.lbo_b >> 4) & 0x3FFFu) << 16;
      const uint32_t a_units0 = (smem_u32(sA) >> 4) + (uint32_t)p.halo_lo;          // 16-byte units
      const uint32_t a_buf_units = a_bytes >> 4;
      const uint32_t a_k16 = 2u * (p.lbo_a >> 4);
      const uint32_t b_units0 = smem_u32(sB) >> 4, b_stage_units = b_bytes >> 4, b_k16 = 2u * (p.lbo_b >> 4);
      const uint32_t d1 = tmem_base + (uint32_t)p.NT;
      const bool two = p.mt == 2;
      int st = 0, ph = 0;
      for (int c = 0; c < p.n_pass; ++c) {
        const int buf = c & 1;
        mbar_wait(smem_u32(&full_a[buf]), (c >> 1) & 1);
        tc_fence_after();
        if (c < 64) TRACE(18 + 4 * c);
        const uint32_t au = a_units0 + (uint32_t)buf * a_buf_units;
        for (int t = 0; t < p.ntap; ++t) {
          mbar_wait(smem_u32(&full_b[st]), ph);
          tc_fence_after();
          const uint32_t a0 = au + (uint32_t)s_delta[t];
          const uint32_t b0 = b_units0 + (uint32_t)st * b_stage_units;
          const uint32_t acc = (c | t) ? 1u : 0u;
          {
            const uint64_t bd = ((uint64_t)hi_b << 32) | ((b0 & 0x3FFFu) | lbo_b_f);
            const uint64_t ad0 = ((uint64_t)hi_a << 32) | ((a0 & 0x3FFFu) | lbo_a_f);
            const uint64_t ad1 = ((uint64_t)hi_a << 32) | (((a0 + 128u) & 0x3FFFu) | lbo_a_f);
            umma_bf16(tmem_base, ad0, bd, idesc, acc);
            if (two) umma_bf16(d1, ad1, bd, idesc, acc);
          }
          {
            const uint64_t bd = ((uint64_t)hi_b << 32) | (((b0 + b_k16) & 0x3FFFu) | lbo_b_f);
            const uint64_t ad0 = ((uint64_t)hi_a << 32) | (((a0 + a_k16) & 0x3FFFu) | lbo_a_f);
            const uint64_t ad1 = ((uint64_t)hi_a << 32) | (((a0 + a_k16 + 128u) & 0x3FFFu) | lbo_a_f);
            umma_bf16(tmem_base, ad0, bd, idesc, 1u);
            if (two) umma_bf16(d1, ad1, bd, idesc, 1u);
          }
          umma_commit(smem_u32(&empty_b[st]));     // frees the weight stage once these MMAs retire
          if (++st == nst) { st = 0; ph ^= 1; }
        }
        umma_commit(smem_u32(&empty_a[buf]));      // frees the operand buffer
        if (c < 64) TRACE(19 + 4 * c);
      }
      umma_commit(smem_u32(acc_full));
    }
  }

  __syncthreads();
  // flush the CTA's GroupNorm statistics: fixed-order sum over the (mt, warp, segment) slots, then one fixed-point
  // integer atomic per (image, group) => deterministic
  if (p.c.ostats) {
    const int og_tile = (p.NT + (1 << p.cpg_out_shift) - 1) >> p.cpg_out_shift;
    for (int i = tid; i < kNimgMax * og_tile; i += kThreads) {
      const int il = i / og_tile, gl = i - il * og_tile;
      const int img = img_lo + il;
      float a = 0.f, b = 0.f;
      bool any = false;
#pragma unroll 4
      for (int k = 0; k < 8 * kSegMax; ++k) {
        if (s_partkey[k] == il) {
          a += s_part[(k * kOgMax + gl) * 2];
          b += s_part[(k * kOgMax + gl) * 2 + 1];
          any = true;
        }
      }
      if (any && img < p.c.B) stat_add(p.c.ostats + ((long)img * p.c.ogroups + (n0 >> p.cpg_out_shift) + gl) * 2, a, b);
    }
  }
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
  if (tid == 0) TRACE(7);
}

static size_t smem_fixed_bytes(const Params& p) {
  return (size_t)p.n_abuf * 4 * p.PA * 16 + (2 * kStagesMax + 5) * 8 + 16 + 16 * 4 + 128 * 4 + (size_t)kNimgMax * kGroupsMax * 8 +
         (size_t)8 * kSegMax * kOgMax * 2 * 4 + 8 * kSegMax * 4 + 2 * kMcta * 4 + 2 * (size_t)p.P * 4 + 128;
}
static size_t smem_bytes(const Params& p) { return smem_fixed_bytes(p) + (size_t)p.nstage * 4 * p.NT * 16; }
constexpr size_t kSmemLimit = 113 * 1024;
// deepest weight ring (<= kStagesMax, >= 4) that still lets two CTAs share an SM
static bool pick_stages(Params& p) {
  const size_t fixed = smem_fixed_bytes(p), stage = (size_t)4 * p.NT * 16;
  if (fixed + 4 * stage > kSmemLimit) return false;
  int n = (int)((kSmemLimit - fixed) / stage);
  p.nstage = n > kStagesMax ? kStagesMax : n;
  // the epilogue stages both 128-row output tiles in the (then dead) operand + weight buffers
  return (size_t)p.mt * 128 * (p.NT * 2 + 16) <= (size_t)p.n_abuf * 4 * p.PA * 16 + (size_t)p.nstage * stage;
}

static int pick_nt(int cout) { return cout % 128 == 0 ? 128 : (cout % 64 == 0 ? 64 : (cout % 32 == 0 ? 32 : 0)); }

static bool fill_params(const ConvP& c, int geo, Params& p) {
  p = Params();
  p.c = c;
  p.geo = geo;
  p.NT = pick_nt(c.Cout);
  if (!p.NT) return false;
  p.H = c.Hin; p.W = c.Win; p.HW = c.Hin * c.Win;
  p.ksize = c.ksize;
  p.tiles_per_phase = c.Cout / p.NT;
  p.nt_shift = p.NT == 128 ? 7 : (p.NT == 64 ? 6 : 5);
  int halo_hi = 0;
  if (geo == GEO_SAME) {
    if (c.ksize != 1 && c.ksize != 3) return false;
    if (c.C1 <= 0 || c.C1 % kCk || c.C2 % kCk) return false;
    p.ntap = c.ksize * c.ksize;
    p.pad = c.ksize == 3 ? 1 : 0;
    p.Wv = c.Win + p.pad;
    p.S = (c.Hin + p.pad) * p.Wv;
    p.halo_lo = halo_hi = c.ksize == 3 ? p.Wv + 1 : 0;
    p.n_pass = (c.C1 + c.C2) / kCk;
  } else if (geo == GEO_DOWN) {
    if (c.C2 != 0 || c.C1 % kCk || (c.Hin & 1) || (c.Win & 1) || c.pro != PRO_NONE || c.ostats || c.res) return false;
    if (c.Hin / 2 + 1 > 127 || c.B >= (1 << 17)) return false;
    p.ntap = 4;
    p.pad = 0;
    p.Wv = c.Win / 2 + 1;
    p.S = (c.Hin / 2 + 1) * p.Wv;
    p.halo_lo = 0;
    halo_hi = p.Wv + 1;
    p.n_pass = 4 * c.C1 / kCk;
  } else if (geo == GEO_UP) {
    if (c.C2 != 0 || c.C1 % kCk || c.pro != PRO_NONE || c.ostats || c.res) return false;
    p.ntap = 4;
    p.pad = 1;
    p.Wv = c.Win + 1;
    p.S = (c.Hin + 1) * p.Wv;
    p.halo_lo = halo_hi = p.Wv + 1;
    p.n_pass = c.C1 / kCk;
  } else {
    if (7 * c.C1 > kCk || c.C2 != 0) return false;
    p.ntap = 7;
    p.pad = 0;
    p.Wv = c.Win;
    p.S = (c.Hin + 3) * p.Wv;
    p.halo_lo = halo_hi = 3 * p.Wv;
    p.n_pass = 1;
  }
  p.n_abuf = (geo == GEO_INIT) ? 1 : 2;
  {
    // two accumulators (256 rows) per CTA halve the weight traffic per FLOP; with few tiles one accumulator keeps more SMs busy
    const long tiles2 = ((long)c.B * p.S + 255) / 256 * (geo == GEO_UP ? 4 : 1) * (c.Cout / p.NT);
    p.mt = tiles2 >= 240 ? 2 : 1;
    p.mcta = 128 * p.mt;
  }
  p.P = p.mcta + p.halo_lo + halo_hi;
  p.PA = p.P;
  while (p.PA % 8 != 2) ++p.PA;
  if (geo != GEO_INIT && (4 * p.P + kProducerThreads - 1) / kProducerThreads > kMaxItems) return false;
  if (geo == GEO_INIT) p.P = p.P;   // (the stem's producer loops over the window; no per-thread item table)
  if (kMcta / p.S + 3 > kNimgMax || p.S < 16) return false;
  p.total_flat = (long)c.B * p.S;
  p.lbo_a = (uint32_t)p.PA * 16u;
  p.sbo_a = 128u;
  p.lbo_b = (uint32_t)p.NT * 16u;
  p.sbo_b = 128u;
  p.tmem_cols = (uint32_t)(p.mt * p.NT);
  if (p.tmem_cols < 32) p.tmem_cols = 32;
  if (p.total_flat + 4096 >= (1L << 30) || (long)c.B * 4 * p.HW >= (1L << 30)) return false;   // 32-bit index arithmetic
  p.cpg_in = 1;
  p.cpg_in_shift = p.cpg_out_shift = 0;
  p.inv_cnt_in = 0.f;
  if (c.pro & PRO_GN) {
    if (geo != GEO_SAME || c.C2 != 0 || c.pgroups <= 0 || c.pgroups > kGroupsMax || c.C1 % c.pgroups) return false;
    p.cpg_in = c.C1 / c.pgroups;
    if (p.cpg_in % 8 || (p.cpg_in & (p.cpg_in - 1))) return false;
    while ((1 << p.cpg_in_shift) < p.cpg_in) ++p.cpg_in_shift;
    p.inv_cnt_in = 1.f / (float)(p.HW * p.cpg_in);
  } else if (c.pro != PRO_NONE) {
    return false;   // SiLU / temb only come together with the GroupNorm apply
  }
  p.cpg_out = 1;
  if (c.ogroups > 0) {
    if (c.Cout % c.ogroups) return false;
    p.cpg_out = c.Cout / c.ogroups;
    if (p.cpg_out % 16 || (p.cpg_out & (p.cpg_out - 1))) return false;
    while ((1 << p.cpg_out_shift) < p.cpg_out) ++p.cpg_out_shift;
    if ((p.NT + p.cpg_out - 1) / p.cpg_out > kOgMax) return false;
    if (p.cpg_out < p.NT && p.NT % p.cpg_out) return false;
    if (p.cpg_out > p.NT && p.cpg_out % p.NT) return false;
  }
  return pick_stages(p);
}

static int geo_of(const ConvP& c) { return c.mode == CONV_SAME ? GEO_SAME : (c.mode == CONV_DOWN ? GEO_DOWN : GEO_UP); }

template <int GEO>
static int launch(Params p, cudaStream_t st) {
  static const bool trace_on = [] {
    const char* e = getenv("DMN_TC_TRACE");
    return e && e[0] == '1';
  }();
  if (trace_on) {
    void* sym = nullptr;
    DMN_CUDA_CHECK(cudaGetSymbolAddress(&sym, g_trace));
    p.trace = (long long*)sym;
    const long nx = (p.total_flat + p.mcta - 1) / p.mcta;
    p.trace_cta = (int)(nx / 2);
  }
  static bool attr_set = false;
  if (!attr_set) {
    DMN_CUDA_CHECK(cudaFuncSetAttribute(conv_tcgen05_kernel<GEO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
    attr_set = true;
  }
  const unsigned gy = (unsigned)((GEO == GEO_UP ? 4 : 1) * (p.c.Cout / p.NT));
  dim3 grid((unsigned)((p.total_flat + p.mcta - 1) / p.mcta), gy);
  conv_tcgen05_kernel<GEO><<<grid, kThreads, smem_bytes(p), st>>>(p);
  count_launch();
  DMN_LAUNCH_CHECK("conv_tcgen05");
  return 0;
}

}  // namespace tc

bool conv_tcgen05_supported(const ConvP& c) {
  tc::Params p;
  return tc::fill_params(c, tc::geo_of(c), p);
}
bool init_conv_tcgen05_supported(int Cin, int S, int Cout, int B) {
  ConvP c;
  c.C1 = Cin; c.Hin = c.Win = S; c.Cout = Cout; c.B = B;
  tc::Params p;
  return tc::fill_params(c, tc::GEO_INIT, p);
}

size_t conv_tcgen05_weight_bytes(int mode, int ksize, int cin, int cout) {
  const int taps = (mode == CONV_SAME) ? ksize * ksize : 16;
  return (size_t)cout * cin * taps * 2;
}
size_t init_conv_tcgen05_weight_bytes(int cout) { return (size_t)cout * 7 * tc::kCk * 2; }

// blocked bf16 weight image [n_tile][pass][tap][kchunk 0..3][n 0..NT-1][8 channels]
void conv_tcgen05_pack_weights(int mode, int ksize, int cin, int cout, const float* w, void* dst_host) {
  const int NT = tc::pick_nt(cout);
  bf16* dst = (bf16*)dst_host;
  size_t o = 0;
  if (mode == CONV_SAME) {
    const int taps = ksize * ksize;
    for (int nt = 0; nt < cout / NT; ++nt)
      for (int c = 0; c < cin / tc::kCk; ++c)
        for (int t = 0; t < taps; ++t)
          for (int kc = 0; kc < 4; ++kc)
            for (int n = 0; n < NT; ++n)
              for (int e = 0; e < 8; ++e) {
                const int co = nt * NT + n, ci = c * tc::kCk + kc * 8 + e;
                dst[o++] = __float2bfloat16_rn(w[((long)co * cin + ci) * taps + t]);
              }
  } else if (mode == CONV_DOWN) {
    // Conv2d weight [co][ci][ky][kx], k4 s2 p1.  virtual channel = sub*C + ci, sub = sy*2+sx; tap t = du*2+dv;
    // ky = 2*du + sy, kx = 2*dv + sx
    for (int nt = 0; nt < cout / NT; ++nt)
      for (int c = 0; c < 4 * cin / tc::kCk; ++c)
        for (int t = 0; t < 4; ++t)
          for (int kc = 0; kc < 4; ++kc)
            for (int n = 0; n < NT; ++n)
              for (int e = 0; e < 8; ++e) {
                const int vc = c * tc::kCk + kc * 8 + e, sub = vc / cin, ci = vc % cin;
                const int ky = 2 * (t >> 1) + (sub >> 1), kx = 2 * (t & 1) + (sub & 1);
                const int co = nt * NT + n;
                dst[o++] = __float2bfloat16_rn(w[(((long)co * cin + ci) * 4 + ky) * 4 + kx]);
              }
  } else {
    // ConvTranspose2d weight [ci][co][ky][kx], k4 s2 p1.  n_tile = phase*(cout/NT) + ct, phase = py*2+px;
    // tap t = a*2+b: dy = py ? (a ? 0 : +1) : (a ? -1 : 0) (dx likewise); ky = py + 1 - 2*dy
    for (int ph = 0; ph < 4; ++ph)
      for (int ct = 0; ct < cout / NT; ++ct)
        for (int c = 0; c < cin / tc::kCk; ++c)
          for (int t = 0; t < 4; ++t) {
            const int py = ph >> 1, px = ph & 1, a = t >> 1, b = t & 1;
            const int dy = py ? (a ? 0 : 1) : (a ? -1 : 0), dx = px ? (b ? 0 : 1) : (b ? -1 : 0);
            const int ky = py + 1 - 2 * dy, kx = px + 1 - 2 * dx;
            for (int kc = 0; kc < 4; ++kc)
              for (int n = 0; n < NT; ++n)
                for (int e = 0; e < 8; ++e) {
                  const int ci = c * tc::kCk + kc * 8 + e, co = ct * NT + n;
                  dst[o++] = __float2bfloat16_rn(w[(((long)ci * cout + co) * 4 + ky) * 4 + kx]);
                }
          }
  }
}

// init conv weight [co][ch][7][7] -> [n_tile][tap ky][kchunk][n][8], virtual channel vc = kx*Cin + ch
void init_conv_tcgen05_pack_weights(int cin, int cout, const float* w, void* dst_host) {
  const int NT = tc::pick_nt(cout);
  bf16* dst = (bf16*)dst_host;
  size_t o = 0;
  for (int nt = 0; nt < cout / NT; ++nt)
    for (int ky = 0; ky < 7; ++ky)
      for (int kc = 0; kc < 4; ++kc)
        for (int n = 0; n < NT; ++n)
          for (int e = 0; e < 8; ++e) {
            const int vc = kc * 8 + e, kx = vc / cin, ch = vc % cin, co = nt * NT + n;
            dst[o++] = __float2bfloat16_rn(kx < 7 ? w[(((long)co * cin + ch) * 7 + ky) * 7 + kx] : 0.f);
          }
}

// debug: copy the trace of the last traced launch to the host (DMN_TC_TRACE=1)
int conv_tcgen05_read_trace(long long* out, int n) {
  if (n > 1024) n = 1024;
  DMN_CUDA_CHECK(cudaDeviceSynchronize());
  DMN_CUDA_CHECK(cudaMemcpyFromSymbol(out, tc::g_trace, (size_t)n * sizeof(long long)));
  return 0;
}

int conv_tcgen05(const ConvP& c, cudaStream_t st) {
  tc::Params p;
  const int geo = tc::geo_of(c);
  if (!tc::fill_params(c, geo, p)) return fail(-2, "conv_tcgen05: unsupported convolution shape");
  static const bool swap = [] {
    const char* e = getenv("DMN_UMMA_SWAP_LBO_SBO");
    return e && e[0] == '1';
  }();
  if (swap) {   // negative control for the descriptor encoding (tests only)
    std::swap(p.lbo_a, p.sbo_a);
    std::swap(p.lbo_b, p.sbo_b);
  }
  if (geo == tc::GEO_SAME) return tc::launch<tc::GEO_SAME>(p, st);
  if (geo == tc::GEO_DOWN) return tc::launch<tc::GEO_DOWN>(p, st);
  return tc::launch<tc::GEO_UP>(p, st);
}

int init_conv_tcgen05(const InitConvP& q, cudaStream_t st) {
  ConvP c;
  c.src1 = q.x; c.C1 = q.Cin; c.B = q.B; c.Hin = c.Win = q.S; c.Hout = c.Wout = q.S; c.Cout = q.Cout;
  c.w = q.w; c.bias = q.bias; c.out = q.out;
  tc::Params p;
  if (!tc::fill_params(c, tc::GEO_INIT, p)) return fail(-2, "init_conv_tcgen05: unsupported shape");
  p.cls_w = q.cls_w;
  p.classes = q.classes;
  p.pad_class = q.pad_class;
  return tc::launch<tc::GEO_INIT>(p, st);
}

}  // namespace dmn
