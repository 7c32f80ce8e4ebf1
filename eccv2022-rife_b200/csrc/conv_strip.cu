// Fused pair of 3x3 convolutions for the thin layers (C = 16: Head.cnn1 + cnn2; C = 32: two ResConvs of IFBlock 4;
// train_log/IFNet_HDv3.py:28-29, 45-54, 63-72), kx-stacked tcgen05 MMAs on "strip" tiles.
//
// Why a third conv kernel (measured on B200, tools/tmem_bench.cu + tools/umma_bench2.cu, profiles/r02_*):
//   * a tcgen05.mma with M = 128, K = 16 from smem costs ~49 cycles for every N <= 64 (the 4 KB A tile is re-read from
//     smem by each MMA) and 56 cycles at N = 96.  conv_halo.cu issues one MMA per (tap, K step): 18 x 49 = 880 cycles
//     per 128 pixels at C = 32, 9 x 49 at C = 16 -- the thin layers sit on that issue floor, and on HBM at the same time
//     (each layer reads and writes its activation tensor once).
//   * here the three horizontal taps of a kernel row are stacked into N: per (ky, K step) ONE MMA multiplies the pixel
//     strip shifted by ky - 1 rows with the weight tile [W(ky,0) | W(ky,1) | W(ky,2)] (N = 3C): 6 x 56 = 337 cycles per
//     128 pixels at C = 32, 3 x 49 at C = 16.  Accumulator column group kx of pixel m holds the contribution of INPUT
//     pixel m to output pixel m - kx + 1, so   out[x] = D0[x - 1] + D1[x] + D2[x + 1]   -- a +-1 lane shift in the
//     epilogue (two warp shuffles per channel);
//   * the strip: activations are planar-8 [B][C/8][H][W][8] (conv_halo.cu); ONE TMA box {32 px, 20 rows, C/8 planes}
//     lands as planes of [20][32][8] halfs = 16 bytes per pixel, 32 pixels per row.  An M-tile is 4 whole rows = 128
//     consecutive 16-byte pixels (tcgen05 no-swizzle K-major canonical layout with SBO = 128 B, LBO = plane pitch), so
//     every epilogue warp (= one TMEM lane quarter) owns one tile row, lane = column, and the lane shift never crosses a
//     warp: lanes 0 and 31 are the halo columns;
//   * two layers per launch: layer A's epilogue writes its 18 x 30 output (zero outside the image = the second conv's
//     zero padding) back into shared memory in the same layout, layer B's MMAs read it from there, so the pair reads
//     and writes HBM once (the intermediate activation never leaves the SM).  Cost: 9 M-tiles per 16 x 28 output
//     pixels of two layers where 7 would do without halo recompute.
// Pipeline per CTA (persistent, static round robin over tiles): TMA producer warp, MMA issuer warp, 8 epilogue warps
// (two per TMEM lane quarter, alternating M-tiles), 2 input stages, 2 intermediate buffers, 4 TMEM accumulator slots;
// the MMAs of layer A of tile i+1 are issued before layer B of tile i so that the tensor pipe works while the
// intermediate of tile i is being written.
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>

#include <mutex>
#include <utility>

#include "conv_strip.cuh"

namespace rife {
namespace {

constexpr int SP = 32;                  // strip pitch = pixels per smem row = lanes of an epilogue warp
constexpr int SR = 20;                  // rows of the input window
constexpr int S_TR = 16, S_TC = 28;     // output tile (rows 2..17, columns 2..29 of the window)
constexpr int NMT_A = 5, NMT_B = 4;     // M-tiles (4 rows each): layer A rows 1..20 (19, 20 unused), layer B rows 2..17
constexpr int PLANE = SR * SP * 16;     // bytes per 8-channel plane of a window
constexpr int NSLOT = 4;                // TMEM accumulator slots
// epilogue warps: C = 32 runs one CTA per SM with 16 (four per SM sub-partition: the epilogue is a long dependent chain
// -- TMEM load, shuffles, residual, pack -- and needs that many to keep the issue slots busy); C = 16 runs two CTAs per
// SM with 8 each.  Warps w, w + 4, ... share a TMEM lane quarter; group g = (w - 2) / 4 takes the M-tiles with seq % NG == g.
template <int C> struct EpiW { static constexpr int value = C == 32 ? 16 : 8; };

struct StripParams {
  int B, H, W, tiles_x, tiles_y, ntiles;
  int bf;                      // 0 = fp16, 1 = bf16 operands and activations (RIFE_PRECISION_BF16)
  int resA, resB;              // layer tail: 0 = lrelu(acc + bias), 1 = lrelu((acc + bias) * beta + x)   (ResConv)
  const __half* w;             // [layer][ky][C/8][3C][8]
  const float* bias[2];
  const float* beta[2];
  __half* out;
  long long* dbg;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ float lrelu_max(float x) { return fmaxf(x, 0.2f * x); }   // == x >= 0 ? x : 0.2 x
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  while (true) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
    if (clock64() - t0 > 4000000000LL) __trap();  // a mis-programmed pipeline traps instead of hanging the GPU
  }
}
__device__ __forceinline__ void mbar_wait_t(uint32_t bar, uint32_t parity, long long& acc) {
  const long long t0 = clock64();
  mbar_wait(bar, parity);
  acc += clock64() - t0;
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ uint64_t make_desc_noswz(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version 1 (sm_100); layout type 0 = SWIZZLE_NONE
  return d;
}
template <bool ACC>
__device__ __forceinline__ void umma_f16_c(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc) {
  if constexpr (ACC)
    asm volatile("tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, 1;" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc) : "memory");
  else
    asm volatile("tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, 0;" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// three 16-column accumulator groups (the kx = 0, 1, 2 partial sums of 16 output channels) with one wait
__device__ __forceinline__ void tmem_ld3x16(uint32_t a0, uint32_t a1, uint32_t a2, float (&d0)[16], float (&d1)[16],
                                            float (&d2)[16]) {
  uint32_t r[48];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%48];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%49];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47}, [%50];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]),
        "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]),
        "=r"(r[46]), "=r"(r[47])
      : "r"(a0), "r"(a1), "r"(a2)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    d0[i] = __uint_as_float(r[i]);
    d1[i] = __uint_as_float(r[16 + i]);
    d2[i] = __uint_as_float(r[32 + i]);
  }
}

template <int C>
struct StripCfg {
  static constexpr int G = C / 8, KSTEPS = C / 16, N = 3 * C;
  static constexpr int SLOT = C == 32 ? 128 : 64;       // TMEM columns per accumulator slot (N = 96 / 48 used)
  static constexpr int TMEM_COLS = NSLOT * SLOT;        // 512 / 256
  static constexpr int BUF = G * PLANE;                 // bytes of one window (all planes)
  static constexpr int WBYTES = 9 * C * C * 2;          // one layer's weights
  static constexpr int SMEM = 2 * WBYTES + 4 * BUF + 2048 + 1024;   // + over-read slack + alignment
};

// One M-tile row of one layer: lane = window column, `row` = window row.  TO_SMEM: layer A (write the intermediate,
// zero outside the image); else layer B (write the output tensor, valid columns / rows only).
template <int C, bool TO_SMEM>
__device__ __forceinline__ void epi_row(const StripParams& p, uint32_t tacc, const float* s_bias, const float* s_beta,
                                        bool res, uint32_t res_smem, uint32_t dst_smem, int b, int gy, int gx, int lane) {
  using K = StripCfg<C>;
  const bool bf = p.bf != 0;
  bool keep;
  if constexpr (TO_SMEM) keep = gy >= 0 && gy < p.H && gx >= 0 && gx < p.W && lane >= 1 && lane <= SP - 2;
  else keep = gy < p.H && gx < p.W && lane >= 2 && lane <= SP - 3;
#pragma unroll
  for (int h = 0; h < C / 16; ++h) {
    float d0[16], d1[16], d2[16], v[16];
    tmem_ld3x16(tacc + 16 * h, tacc + C + 16 * h, tacc + 2 * C + 16 * h, d0, d1, d2);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float up = __shfl_up_sync(0xffffffffu, d0[i], 1);      // kx = 0 partial sum of input pixel x - 1
      const float dn = __shfl_down_sync(0xffffffffu, d2[i], 1);    // kx = 2 partial sum of input pixel x + 1
      v[i] = (up + d1[i]) + dn;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 bb = *reinterpret_cast<const float4*>(s_bias + 16 * h + 4 * i);
      v[4 * i] += bb.x; v[4 * i + 1] += bb.y; v[4 * i + 2] += bb.z; v[4 * i + 3] += bb.w;
    }
#pragma unroll
    for (int g2 = 0; g2 < 2; ++g2) {
      const int g = 2 * h + g2;
      float a[8];
      if (res) {
        uint4 u;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w)
                     : "r"(res_smem + (uint32_t)(g * PLANE)));
        float rr[8];
        unpack8x(u, bf, rr);
        const float4 b0 = *reinterpret_cast<const float4*>(s_beta + 8 * g);
        const float4 b1 = *reinterpret_cast<const float4*>(s_beta + 8 * g + 4);
        const float be[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = lrelu_max(fmaf(v[8 * g2 + i], be[i], rr[i]));
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = lrelu_max(v[8 * g2 + i]);
      }
      if constexpr (TO_SMEM) {
        uint4 u = pack8x(a, bf);
        if (!keep) u = make_uint4(0u, 0u, 0u, 0u);
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(dst_smem + (uint32_t)(g * PLANE)), "r"(u.x), "r"(u.y),
                     "r"(u.z), "r"(u.w)
                     : "memory");
      } else {
        if (keep) st8x(p.out + ((((size_t)b * K::G + g) * p.H + gy) * p.W + gx) * 8, a, bf);
      }
    }
  }
}

template <int C>
__global__ void __launch_bounds__(64 + 32 * EpiW<C>::value, C == 16 ? 2 : 1)
conv_strip2_kernel(const __grid_constant__ CUtensorMap tm_in, const StripParams p) {
  using K = StripCfg<C>;
  constexpr int NUM_EPI_WARPS = EpiW<C>::value;
  constexpr int NG = NUM_EPI_WARPS / 4;         // epilogue groups
  constexpr int NTHREADS = 64 + 32 * NUM_EPI_WARPS;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t w_full;
  __shared__ __align__(8) uint64_t in_full[2], in_empty[2], mid_full[2], mid_empty[2];
  __shared__ __align__(8) uint64_t tm_full[NSLOT], tm_empty[NSLOT];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float s_bias[2][32];
  __shared__ __align__(16) float s_beta[2][32];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t w_smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t in_smem = w_smem + 2 * K::WBYTES;          // 2 input stages (WBYTES is a multiple of 1024 for C = 16, 32)
  const uint32_t mid_smem = in_smem + 2 * K::BUF;           // 2 intermediate buffers (+ 2 KB over-read slack behind them)

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_in) : "memory");
    mbar_init(smem_u32(&w_full), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&in_full[i]), 1);
      mbar_init(smem_u32(&in_empty[i]), 1 + NUM_EPI_WARPS);   // layer A's MMAs (commit) + its epilogues (residual reads)
      mbar_init(smem_u32(&mid_full[i]), NUM_EPI_WARPS);       // every epilogue warp has written its rows of layer A
      mbar_init(smem_u32(&mid_empty[i]), 1 + NUM_EPI_WARPS);  // layer B's MMAs (commit) + its epilogues (residual reads)
    }
    for (int i = 0; i < NSLOT; ++i) {
      mbar_init(smem_u32(&tm_full[i]), 1);
      mbar_init(smem_u32(&tm_empty[i]), 4);                   // the four warps of the group that drains this slot
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "r"((uint32_t)K::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp >= 2) {
    for (int i = threadIdx.x - 64; i < 2 * 32; i += NTHREADS - 64) {
      const int l = i >> 5, c = i & 31;
      s_bias[l][c] = c < C ? p.bias[l][c] : 0.f;
      s_beta[l][c] = (c < C && p.beta[l]) ? p.beta[l][c] : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const int ntl = (int)blockIdx.x < p.ntiles ? (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int per_img = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- producer: both layers' weights once, then one window per tile
      const uint32_t wb = smem_u32(&w_full);
      mbar_expect_tx(wb, 2u * K::WBYTES);
      for (int off = 0; off < 2 * K::WBYTES; off += 18432)
        bulk_load(w_smem + off, reinterpret_cast<const uint8_t*>(p.w) + off, (uint32_t)min(18432, 2 * K::WBYTES - off), wb);
      pdl_wait();  // the activations are the previous kernel's output; everything else in this CTA is downstream of these loads
      for (int k = 0; k < ntl; ++k) {
        const int tile = (int)blockIdx.x + k * (int)gridDim.x;
        const int b = tile / per_img, r = tile - b * per_img;
        const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
        const int s = k & 1;
        mbar_wait(smem_u32(&in_empty[s]), (((uint32_t)k >> 1) & 1u) ^ 1u);
        const uint32_t fb = smem_u32(&in_full[s]);
        mbar_expect_tx(fb, (uint32_t)K::BUF);
        tma_load_4d(in_smem + s * K::BUF, &tm_in, fb, (tx * S_TC - 2) * 8, ty * S_TR - 2, 0, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- MMA issuer: A(t0); [A(t_k), B(t_k-1)] ...; B(t_last)
      const uint32_t idesc = (1u << 4) | (p.bf ? ((1u << 7) | (1u << 10)) : 0u) | ((uint32_t)(K::N >> 3) << 17) |
                             ((uint32_t)(128 >> 4) << 24);
      const uint64_t a_desc0 = make_desc_noswz(0, (uint32_t)PLANE, 128u);
      const uint64_t b_desc0 = make_desc_noswz(0, (uint32_t)K::N * 16u, 128u);
      mbar_wait(smem_u32(&w_full), 0);
      uint32_t seq = 0;
      long long t_te = 0, t_in = 0, t_mid = 0;
      const long long t_begin = clock64();
      auto issue_layer = [&](uint32_t a_buf, uint32_t w_layer, int row_first, int nmt) {
        for (int j = 0; j < nmt; ++j, ++seq) {
          const uint32_t slot = seq & (NSLOT - 1);
          mbar_wait_t(smem_u32(&tm_empty[slot]), ((seq >> 2) & 1u) ^ 1u, t_te);
          tc_fence_after();
          const uint32_t acc = tmem_base + slot * K::SLOT;
          // window row (row_first + 4j + ky - 1), column 0: 128 consecutive pixels from there on
          const uint32_t a16 = (a_buf + (uint32_t)((row_first + 4 * j - 1) * SP * 16)) >> 4;
          const uint32_t b16 = w_layer >> 4;
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kk = 0; kk < K::KSTEPS; ++kk) {
              const uint64_t da = a_desc0 + (uint64_t)(a16 + (uint32_t)(ky * SP + kk * 2 * (PLANE >> 4)));
              const uint64_t db = b_desc0 + (uint64_t)(b16 + (uint32_t)((ky * K::KSTEPS + kk) * 2 * K::N));
              if (ky == 0 && kk == 0) umma_f16_c<false>(acc, da, db, idesc);
              else umma_f16_c<true>(acc, da, db, idesc);
            }
          umma_commit(smem_u32(&tm_full[slot]));
        }
      };
      for (int k = 0; k <= ntl; ++k) {
        if (k < ntl) {
          const int s = k & 1;
          mbar_wait_t(smem_u32(&in_full[s]), ((uint32_t)k >> 1) & 1u, t_in);
          tc_fence_after();
          issue_layer(in_smem + s * K::BUF, w_smem, 1, NMT_A);
          umma_commit(smem_u32(&in_empty[s]));
        }
        if (k >= 1) {
          const int m = (k - 1) & 1;
          mbar_wait_t(smem_u32(&mid_full[m]), ((uint32_t)(k - 1) >> 1) & 1u, t_mid);
          tc_fence_after();
          issue_layer(mid_smem + m * K::BUF, w_smem + K::WBYTES, 2, NMT_B);
          umma_commit(smem_u32(&mid_empty[m]));
        }
      }
      if (p.dbg) {
        p.dbg[blockIdx.x * 8 + 0] = t_te; p.dbg[blockIdx.x * 8 + 1] = t_in; p.dbg[blockIdx.x * 8 + 2] = t_mid;
        p.dbg[blockIdx.x * 8 + 3] = clock64() - t_begin;
      }
    }
  } else {
    // ---------------- epilogue warps: TMEM lane quarter q = warp % 4 = row of the M-tile; group = M-tile parity
    const int q = warp & 3;
    const int grp = (warp - 2) >> 2;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t seq = 0;
    long long t_tf = 0, t_me = 0;
    const long long t_begin = clock64();
    for (int k = 0; k <= ntl; ++k) {
      if (k < ntl) {
        // layer A of tile k: window rows 1..18 -> intermediate buffer k & 1
        const int tile = (int)blockIdx.x + k * (int)gridDim.x;
        const int b = tile / per_img, r = tile - b * per_img;
        const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
        const int s = k & 1;
        const uint32_t ibuf = in_smem + s * K::BUF, mbuf = mid_smem + s * K::BUF;
        bool first = true;
        for (int j = 0; j < NMT_A; ++j, ++seq) {
          if ((int)(seq & (uint32_t)(NG - 1)) != grp) continue;
          if (first) {
            // the intermediate buffer is free (layer B of tile k - 2 has read it); the window is visible to this warp
            mbar_wait_t(smem_u32(&mid_empty[s]), (((uint32_t)k >> 1) & 1u) ^ 1u, t_me);
            mbar_wait(smem_u32(&in_full[s]), ((uint32_t)k >> 1) & 1u);
            first = false;
          }
          const uint32_t slot = seq & (NSLOT - 1);
          mbar_wait_t(smem_u32(&tm_full[slot]), (seq >> 2) & 1u, t_tf);
          tc_fence_after();
          const int row = 1 + 4 * j + q;
          if (row <= SR - 2) {
            const uint32_t px = (uint32_t)((row * SP + lane) * 16);
            epi_row<C, true>(p, trow + slot * K::SLOT, s_bias[0], s_beta[0], p.resA != 0, ibuf + px, mbuf + px, b,
                             ty * S_TR - 2 + row, tx * S_TC - 2 + lane, lane);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&tm_empty[slot]));
        }
        // this warp's rows of the intermediate are written: publish them to the async proxy (the MMAs of layer B)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(smem_u32(&mid_full[s]));
          mbar_arrive(smem_u32(&in_empty[s]));
        }
      }
      if (k >= 1) {
        // layer B of tile k - 1: window rows 2..17 -> the output tensor
        const int tile = (int)blockIdx.x + (k - 1) * (int)gridDim.x;
        const int b = tile / per_img, r = tile - b * per_img;
        const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
        const int m = (k - 1) & 1;
        const uint32_t mbuf = mid_smem + m * K::BUF;
        for (int j = 0; j < NMT_B; ++j, ++seq) {
          if ((int)(seq & (uint32_t)(NG - 1)) != grp) continue;
          const uint32_t slot = seq & (NSLOT - 1);
          mbar_wait_t(smem_u32(&tm_full[slot]), (seq >> 2) & 1u, t_tf);
          tc_fence_after();
          const int row = 2 + 4 * j + q;
          const uint32_t px = (uint32_t)((row * SP + lane) * 16);
          epi_row<C, false>(p, trow + slot * K::SLOT, s_bias[1], s_beta[1], p.resB != 0, mbuf + px, 0u, b,
                            ty * S_TR - 2 + row, tx * S_TC - 2 + lane, lane);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&tm_empty[slot]));
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&mid_empty[m]));
      }
    }
    if (p.dbg && warp == 2 && lane == 0) {
      p.dbg[blockIdx.x * 8 + 4] = t_tf; p.dbg[blockIdx.x * 8 + 5] = t_me; p.dbg[blockIdx.x * 8 + 6] = clock64() - t_begin;
      p.dbg[blockIdx.x * 8 + 7] = ntl;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)K::TMEM_COLS)
                 : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return (EncodeTiledFn)f;
  }();
  return fn;
}

int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

int set_smem_attr(const void* fn, int bytes) {
  static std::vector<std::pair<const void*, int>> seen;
  static std::mutex mu;   // one host thread per GPU may launch concurrently
  std::lock_guard<std::mutex> lock(mu);
  int dev = 0;
  cudaGetDevice(&dev);
  for (auto& s : seen)
    if (s.first == fn && s.second == dev) return RIFE_OK;
  RIFE_CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  seen.emplace_back(fn, dev);
  return RIFE_OK;
}

}  // namespace

bool strip_fits(const HostLayer& h) {
  return h.kind == CONV_3x3 && h.taps == 9 && h.convt_co == 0 && (h.cin == 16 || h.cin == 32) && h.npad == h.cin && h.n == h.cin;
}

int strip_pack_weights(const HostLayer& h, std::vector<__half>* packed, bool bf16) {
  if (!strip_fits(h)) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "strip_pack_weights: the strip kernel takes 3x3 c -> c layers with c = 16 or 32");
  const int C = h.cin, G = C / 8, N = 3 * C;
  packed->assign((size_t)3 * G * N * 8, host_cvt16(0.f, bf16));
  for (int ky = 0; ky < 3; ++ky)
    for (int g = 0; g < G; ++g)
      for (int kx = 0; kx < 3; ++kx)
        for (int co = 0; co < C; ++co)
          for (int c = 0; c < 8; ++c)
            (*packed)[((((size_t)ky * G + g) * N) + kx * C + co) * 8 + c] =
                host_cvt16(h.w[((size_t)(ky * 3 + kx) * h.cin + g * 8 + c) * h.npad + co], bf16);
  return RIFE_OK;
}

int launch_conv_strip2(const StripPair& L, const ConvIO& io, cudaStream_t st) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) RIFE_FAIL(RIFE_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  if (L.c != 16 && L.c != 32) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "conv_strip2: c must be 16 or 32");
  if (!L.w || !L.bias[0] || !L.bias[1] || !io.in || !io.out) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "conv_strip2: null argument");
  if ((L.res[0] && !L.beta[0]) || (L.res[1] && !L.beta[1])) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "conv_strip2: a ResConv layer needs beta");
  StripParams p{};
  p.B = io.B; p.H = io.H; p.W = io.W;
  p.tiles_x = (io.W + S_TC - 1) / S_TC;
  p.tiles_y = (io.H + S_TR - 1) / S_TR;
  p.ntiles = p.tiles_x * p.tiles_y * io.B;
  p.bf = io.in_bf16 ? 1 : 0;
  p.resA = L.res[0]; p.resB = L.res[1];
  p.w = L.w;
  p.bias[0] = L.bias[0]; p.bias[1] = L.bias[1];
  p.beta[0] = L.beta[0]; p.beta[1] = L.beta[1];
  p.out = (__half*)io.out;
  const int G = L.c / 8;
  CUtensorMap tm;
  {
    cuuint64_t dims[4] = {(cuuint64_t)io.W * 8, (cuuint64_t)io.H, (cuuint64_t)G, (cuuint64_t)io.B};
    cuuint64_t strides[3] = {(cuuint64_t)io.W * 16, (cuuint64_t)io.H * io.W * 16, (cuuint64_t)G * io.H * io.W * 16};
    cuuint32_t box[4] = {(cuuint32_t)(SP * 8), (cuuint32_t)SR, (cuuint32_t)G, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(io.in), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) RIFE_FAIL(RIFE_ERR_CUDA, "cuTensorMapEncodeTiled(strip) failed: " + std::to_string((int)r));
  }
  const int per_sm = L.c == 16 ? 2 : 1;
  int grid = num_sms() * per_sm;
  if (grid > p.ntiles) grid = p.ntiles;
  static const bool debug = getenv("RIFE_HK_DEBUG") != nullptr;
  if (debug) {
    cudaMalloc((void**)&p.dbg, (size_t)grid * 64);
    cudaMemsetAsync(p.dbg, 0, (size_t)grid * 64, st);
  }
  if (L.c == 16) {
    RIFE_TRY(set_smem_attr((const void*)conv_strip2_kernel<16>, StripCfg<16>::SMEM));
    RIFE_CUDA_TRY(launch_pdl(conv_strip2_kernel<16>, dim3(grid), dim3(64 + 32 * EpiW<16>::value), StripCfg<16>::SMEM, st, tm, p));
  } else {
    RIFE_TRY(set_smem_attr((const void*)conv_strip2_kernel<32>, StripCfg<32>::SMEM));
    RIFE_CUDA_TRY(launch_pdl(conv_strip2_kernel<32>, dim3(grid), dim3(64 + 32 * EpiW<32>::value), StripCfg<32>::SMEM, st, tm, p));
  }
  RIFE_LAUNCH_CHECK();
  if (debug) {
    std::vector<long long> h((size_t)grid * 8);
    cudaStreamSynchronize(st);
    cudaMemcpy(h.data(), p.dbg, (size_t)grid * 64, cudaMemcpyDeviceToHost);
    cudaFree(p.dbg);
    double a[8] = {0};
    for (int i = 0; i < grid; ++i)
      for (int k = 0; k < 8; ++k) a[k] += (double)h[(size_t)i * 8 + k] / grid;
    fprintf(stderr, "[strip] c=%d grid=%d tiles/cta=%.1f | mma: wait tm_empty %.0f, in_full %.0f, mid_full %.0f, total %.0f | epi(w2): wait tm_full %.0f, mid_empty %.0f, total %.0f\n",
            L.c, grid, a[7], a[0], a[1], a[2], a[3], a[4], a[5], a[6]);
  }
  return RIFE_OK;
}

}  // namespace rife
