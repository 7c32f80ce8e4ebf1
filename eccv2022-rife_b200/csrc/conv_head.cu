// The whole Head -- cnn0 (3 -> 16, stride 2) -> cnn1 -> cnn2 (16 -> 16) -> cnn3 (ConvTranspose 16 -> 4) + the assembly of the
// 8-channel SRC pixel -- as ONE persistent tcgen05 kernel (train_log/IFNet_HDv3.py:24-43, called on both frames at :128-129).
//
// Layer by layer the four convs move 80 B per frame pixel through HBM where 24 B are compulsory (the s2d frame in, SRC out):
// each of them sits at 0.8-0.9 of the HBM roofline and at its MMA issue floor at once.  Fusing them
// needs two things:
//   * the activations of a tile stay in shared memory: a tile of 58 x 26 half-resolution output pixels is computed from a
//     65 x 33 window of the s2d frame, every layer shrinking the valid region by one pixel per side (halo recompute:
//     cnn0 64 x 32, cnn1 62 x 30, cnn2 60 x 28, cnn3 58 x 26); two planar-8 buffers of 70 x 36 pixels ping-pong;
//   * cheaper MMAs, or the fused kernel would sit on the issue floor of the four layers together (0.84 ms): an M = 128,
//     K = 16 MMA costs ~49 cycles for every N <= 64 (its A tile, 4 KB, is what it reads), so ROW-STACKING makes N = 64 out of
//     N = 16: the GEMM row is a "super-pixel" of R = 4 vertically adjacent pixels, its N = 4 x 16 columns are the 16 output
//     channels of each of the 4 rows.  Input row j in {-1 .. R} (relative to the super-pixel's first row) feeds output row
//     ro through the tap dy = j - ro, so the B operand of (j, dx) is the block row [W(j - 0, dx) | W(j - 1, dx) | ..] with
//     zero blocks where |dy| > 1: a window of one "master" strip [0 0 0 W(+1) W(0) W(-1) 0 0 0] per dx.  A 3x3 layer takes
//     3 (R + 2) = 18 MMAs per 512 pixels instead of 36, and -- unlike stacking the horizontal taps (conv_strip.cu) -- the
//     epilogue has nothing to add up: column block ro IS output row ro.  The A operand needs no new layout: the super-pixel's
//     rows are R rows of the same planar-8 smem tile, reached through the descriptor's 8-row-group stride (SBO = R x pitch).
// M-tile = 64 rows x 8 columns; a layer's region is 4 such strips; 16 M-tiles, 256 MMAs per output tile.
// Roles: warp 0 TMA producer (weights once, one input window per tile), warp 1 MMA issuer, 8 epilogue warps (two per TMEM
// lane quarter, splitting the 4 stacked rows).  The layers are pipelined strip by strip: strip k of layer l + 1 reads
// strips k and k + 1 of layer l, so it starts as soon as those two epilogues have landed in shared memory (one mbarrier per
// (layer, strip), one phase per tile); 4 TMEM accumulator buffers of 64 columns.
// Zero padding (IFNet_HDv3.py:27-29, padding = 1 at every layer): the TMA box zero-fills outside the frame, and the
// epilogues of cnn0..cnn2 write zeros for positions outside the frame.
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>

#include <mutex>
#include <utility>
#include <vector>

#include "conv_head.cuh"
#include "tc_dev.cuh"
#include "tc_host.cuh"

namespace rife {
namespace {

constexpr int HF_R = 4;                       // stacked rows per GEMM row
constexpr int HF_TH = 58, HF_TW = 26;         // output tile (half-resolution pixels)
constexpr int HF_BP = 36;                     // buffer row pitch (pixels)
constexpr int HF_BROWS = 70;
constexpr int HF_PLANE = HF_BROWS * HF_BP * 16;   // 40320 bytes: one 8-channel plane
constexpr int HF_BUF = 2 * HF_PLANE;              // 16 channels
constexpr int HF_IN_ROWS = HF_TH + 7;             // 65 rows of the s2d frame feed a tile
constexpr int HF_IN_BYTES = 2 * HF_IN_ROWS * HF_BP * 16;
constexpr int HF_NBLK0 = 2 * (HF_R - 1) + 2, HF_NBLK = 2 * (HF_R - 1) + 3;   // blocks per master strip: cnn0 (2 dy) / 3x3
constexpr int HF_W0 = 2 * HF_NBLK0 * 512;         // bytes: 2 dx x 8 blocks x [2 k-chunks][16 rows][16 B]
constexpr int HF_W1 = 3 * HF_NBLK * 512;
constexpr int HF_W_BYTES = HF_W0 + 3 * HF_W1;     // 49664
constexpr int HF_SMEM = HF_W_BYTES + 2 * HF_BUF + 128;
constexpr int HF_THREADS = 64 + 32 * 8 + 32;   // producer, issuer 0, 8 epilogue warps, issuer 1
constexpr int HF_NACC = 4;                        // TMEM accumulator buffers of 64 columns

struct HFParams {
  const __half* w;
  const float* bias;
  const __half* head_in;
  __half* src;
  int F, H2, W2;
  int tiles_x, tiles_y, ntiles;
  int bf;
  long long* dbg;   // RIFE_HF_DEBUG=1: per-CTA role timers (8 x int64), else null
  int mode;         // timing experiments (RIFE_HF_MODE, wrong results): 1 = the epilogue stores nothing, 2 = nor reads TMEM
};

// the MMAs of one 64 x 8 strip of a layer: NDX horizontal taps x NJ input rows, every operand offset an immediate.
// a_base: descriptor of buffer position (r0 - 1, c0 - 1); b_base: descriptor of the layer's first master strip.
template <int NDX, int NJ, int DMAX, int NBLK, int XI = 0, int JI = 0>
__device__ __forceinline__ void hf_issue(uint32_t acc, uint64_t a_base, uint64_t b_base, uint32_t idesc) {
  if constexpr (XI < NDX) {
    constexpr int j = JI - 1;                                          // input row relative to the super-pixel's first row
    constexpr int a_off = JI * HF_BP + XI;                             // 16-byte units (dx = XI - 1)
    constexpr int b_off = XI * (NBLK * 32) + (HF_R - 1 + DMAX - j) * 16;   // block window of the master strip
    tc::umma_f16<!(XI == 0 && JI == 0)>(acc, a_base + (uint64_t)a_off, b_base + (uint64_t)b_off, idesc);
    if constexpr (JI + 1 < NJ) hf_issue<NDX, NJ, DMAX, NBLK, XI, JI + 1>(acc, a_base, b_base, idesc);
    else hf_issue<NDX, NJ, DMAX, NBLK, XI + 1, 0>(acc, a_base, b_base, idesc);
  }
}

__device__ __forceinline__ void sts128(uint32_t a, const uint4& u) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
}

__global__ void __launch_bounds__(HF_THREADS, 1)
head_fused_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ HFParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t w_full, in_full, in_free;
  __shared__ __align__(8) uint64_t strip_ready[3][4];
  __shared__ __align__(8) uint64_t t_full[HF_NACC];
  __shared__ __align__(8) uint64_t t_empty[HF_NACC];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float s_bias[4][16];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t w_smem = (tc::smem_u32(smem_raw) + 127u) & ~127u;
  const uint32_t buf0 = w_smem + HF_W_BYTES, buf1 = buf0 + HF_BUF;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_in) : "memory");
    tc::mbar_init(tc::smem_u32(&w_full), 1);
    tc::mbar_init(tc::smem_u32(&in_full), 1);
    tc::mbar_init(tc::smem_u32(&in_free), 2);   // one commit per MMA issuer
    for (int l = 0; l < 3; ++l)
      for (int k = 0; k < 4; ++k) tc::mbar_init(tc::smem_u32(&strip_ready[l][k]), 8);   // one arrive per epilogue warp
    for (int i = 0; i < HF_NACC; ++i) {
      tc::mbar_init(tc::smem_u32(&t_full[i]), 1);
      tc::mbar_init(tc::smem_u32(&t_empty[i]), 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(&tmem_slot)),
                 "r"((uint32_t)(HF_NACC * 64))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x >= 64 && threadIdx.x < 128) (&s_bias[0][0])[threadIdx.x - 64] = p.bias[threadIdx.x - 64];
  // both buffers start as zeros: an MMA multiplies rows that only feed OTHER output rows by the zero blocks of the weight
  // strips, and 0 x (NaN / Inf bit patterns of never-written shared memory) would poison valid outputs.  From here on the
  // buffers only ever hold zero fill, frame data and earlier tiles' activations.
  for (uint32_t o = threadIdx.x * 16u; o < (uint32_t)(2 * HF_BUF); o += HF_THREADS * 16u)
    sts128(buf0 + o, make_uint4(0u, 0u, 0u, 0u));
  tc::fence_async_smem();
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  const uint32_t tmem_base = tmem_slot;
  const int tiles_per_frame = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- producer: the weight strips once, then one input window per tile into buffer 1
      const uint32_t wb = tc::smem_u32(&w_full);
      tc::mbar_expect_tx(wb, (uint32_t)HF_W_BYTES);
      for (int off = 0; off < HF_W_BYTES; off += 16384)
        tc::bulk_load(w_smem + off, reinterpret_cast<const uint8_t*>(p.w) + off, (uint32_t)min(16384, HF_W_BYTES - off), wb);
      pdl_wait();   // the s2d frames are the previous kernel's output (common.cuh: PDL contract)
      int it = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
        const int f = tile / tiles_per_frame, r = tile - f * tiles_per_frame;
        const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
        // buffer 1 is free again once the MMAs of the previous tile's cnn2 (its last readers) have completed
        if (it > 0) tc::mbar_wait_backoff(tc::smem_u32(&in_free), (uint32_t)(it - 1) & 1u);
        const uint32_t fb = tc::smem_u32(&in_full);
        tc::mbar_expect_tx(fb, (uint32_t)HF_IN_BYTES);
        // s2d frames [F][2][H2][W2][8] fp16 as 32-bit words {4 W2, H2, 2, F}: 36 pixels x 65 rows from (ox - 4, oy - 4) per
        // plane; outside the frame = zero fill = cnn0's padding
        tc::tma_load_4d(buf1, &tm_in, fb, 4 * (tx * HF_TW - 4), ty * HF_TH - 4, 0, f);
        tc::tma_load_4d(buf1 + HF_PLANE, &tm_in, fb, 4 * (tx * HF_TW - 4), ty * HF_TH - 4, 1, f);
      }
    }
  } else if (warp == 1 || warp == 10) {
    if (lane == 0) {
      // ---------------- MMA issuers: two threads on two sub-partitions, even / odd strips.  One thread tops out at ~50-75
      // cycles per tcgen05.mma in instruction issue alone (descriptor arithmetic + the uniform-register hand-off around every
      // UTCHMMA), which is the whole cost of an N = 64, K = 16 MMA; two of them keep the tensor pipe fed.  The epilogue warps
      // walk the groups in order, so when they write for layer l every MMA of layer l - 1 (either issuer) has completed.
      const int me = warp == 1 ? 0 : 1;
      tc::mbar_wait(tc::smem_u32(&w_full), 0);
      // instruction descriptor: D = f32, A / B = f16 or bf16 (bits 7-9 / 10-12), N = 64, M = 128
      const uint32_t idesc16 = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t idescbf = idesc16 | (1u << 7) | (1u << 10);
      const uint64_t a_desc0 = tc::make_desc_noswz(0, (uint32_t)HF_PLANE, (uint32_t)(HF_R * HF_BP * 16));
      const uint64_t b_l0 = tc::make_desc_noswz(w_smem, (uint32_t)(HF_NBLK0 * 256), 128u);
      const uint64_t b_l1 = tc::make_desc_noswz(w_smem + HF_W0, (uint32_t)(HF_NBLK * 256), 128u);
      uint32_t n = 0;
      int it = 0;
      long long t_in = 0, t_sr = 0, t_te = 0;
      const long long t_begin = clock64();
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
        const uint32_t tph = (uint32_t)it & 1u;
#pragma unroll 1
        for (int l = 0; l < 4; ++l) {
          const uint32_t in_buf = (l & 1) ? buf0 : buf1;
          const uint32_t idesc = (l >= 1 && p.bf) ? idescbf : idesc16;
#pragma unroll 1
          for (int k = 0; k < 4; ++k, ++n) {
            if ((k & 1) != me) continue;
            if (l == 0) {
              if (k < 2) tc::mbar_wait_t(tc::smem_u32(&in_full), tph, t_in);
            } else {
              tc::mbar_wait_t(tc::smem_u32(&strip_ready[l - 1][k < 3 ? k + 1 : 3]), tph, t_sr);
            }
            const uint32_t tb = n & (HF_NACC - 1);
            tc::mbar_wait_t(tc::smem_u32(&t_empty[tb]), ((n / HF_NACC) & 1u) ^ 1u, t_te);
            tc::fence_after();
            // strip k of layer l: rows 1 + l .., columns 1 + l + 8 k ..; a_base points one row up and one pixel left
            const uint64_t a_base = a_desc0 + (uint64_t)((in_buf + (uint32_t)((l * HF_BP + l + 8 * k) * 16)) >> 4);
            const uint32_t acc = tmem_base + tb * 64u;
            if (l == 0) hf_issue<2, HF_R + 1, 0, HF_NBLK0>(acc, a_base, b_l0, idesc);
            else hf_issue<3, HF_R + 2, 1, HF_NBLK>(acc, a_base, b_l1 + (uint64_t)(((l - 1) * HF_W1) >> 4), idesc);
            tc::umma_commit(tc::smem_u32(&t_full[tb]));
            if (l == 2 && k >= 2) tc::umma_commit(tc::smem_u32(&in_free));   // this issuer's last reader of buffer 1
          }
        }
      }
      if (p.dbg && me == 0) {
        p.dbg[blockIdx.x * 8 + 0] = t_in; p.dbg[blockIdx.x * 8 + 1] = t_sr; p.dbg[blockIdx.x * 8 + 2] = t_te;
        p.dbg[blockIdx.x * 8 + 3] = clock64() - t_begin;
      }
    }
  } else if (warp >= 2 && warp < 10) {
    // ---------------- epilogue warps: TMEM lane quarter warp % 4 (two warps each); thread = super-pixel (P, c) of the strip;
    // the two warps of a quarter take stacked rows {0, 1} and {2, 3}
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int P = 4 * q + (lane >> 3), c = lane & 7;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    const size_t hw2 = (size_t)p.H2 * p.W2;
    uint32_t n = 0;
    int it = 0;
    long long t_tf = 0;
    const long long t_begin = clock64();
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int f = tile / tiles_per_frame, r = tile - f * tiles_per_frame;
      const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
      const int gy0 = ty * HF_TH - 4, gx0 = tx * HF_TW - 4;   // frame position of buffer (0, 0)
#pragma unroll 1
      for (int l = 0; l < 4; ++l) {
        const uint32_t out_buf = (l & 1) ? buf1 : buf0;
        const bool out_bf = p.bf != 0;
        float bias[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) bias[i] = s_bias[l][i];
#pragma unroll 1
        for (int k = 0; k < 4; ++k, ++n) {
          const uint32_t tb = n & (HF_NACC - 1);
          tc::mbar_wait_t(tc::smem_u32(&t_full[tb]), (n / HF_NACC) & 1u, t_tf);
          tc::fence_after();
          const int bc = 1 + l + 8 * k + c;
          const int x = gx0 + bc;
          const int br0 = 1 + l + 4 * P + 2 * half;          // this thread's two stacked rows: br0, br0 + 1
          const int y0 = gy0 + br0;
          const bool xin = x >= 0 && x < p.W2;
          const bool in0 = xin && y0 >= 0 && y0 < p.H2, in1 = xin && y0 + 1 >= 0 && y0 + 1 < p.H2;
          if (p.mode & 1) {
            // (timing experiment: nothing stored)
          } else if (l < 3) {
            // cnn0 .. cnn2: LeakyReLU(0.2); zero outside the frame (the next layer's padding).  Both rows' accumulators are
            // fetched before the one wait.
            float v0[16], v1[16];
            tc::tmem_ld16_nowait(trow + tb * 64u + (uint32_t)((2 * half) * 16), v0);
            tc::tmem_ld16_nowait(trow + tb * 64u + (uint32_t)((2 * half + 1) * 16), v1);
            tc::tmem_ld_wait(v0, v1);
            const uint32_t dst = out_buf + (uint32_t)((br0 * HF_BP + bc) * 16);
            float a[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = in0 ? lrelu02(v0[i] + bias[i]) : 0.f;
            sts128(dst, pack8x(a, out_bf));
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = in0 ? lrelu02(v0[8 + i] + bias[8 + i]) : 0.f;
            sts128(dst + HF_PLANE, pack8x(a, out_bf));
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = in1 ? lrelu02(v1[i] + bias[i]) : 0.f;
            sts128(dst + HF_BP * 16, pack8x(a, out_bf));
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = in1 ? lrelu02(v1[8 + i] + bias[8 + i]) : 0.f;
            sts128(dst + HF_BP * 16 + HF_PLANE, pack8x(a, out_bf));
          } else {
            // cnn3 (ConvTranspose 16 -> 4, no activation): the complete 8-channel SRC pixel [f0..f3 | r g b a] of the four
            // full-resolution pixels (2y + py, 2x + px); the image channels come from the s2d frame (channel par * 3 + c =
            // RGB of pixel par, channel 12 + par = alpha) -- conv_halo.cu: EPI_HEADT.  The image loads of both rows go out
            // before the accumulators are waited for.
            const bool colv = bc >= 4 && bc < 4 + HF_TW;
            const bool st0 = in0 && colv && br0 >= 4 && br0 < 4 + HF_TH, st1 = in1 && colv && br0 + 1 >= 4 && br0 + 1 < 4 + HF_TH;
            const size_t hp = ((size_t)f * 2 * p.H2 + (st0 ? y0 : (st1 ? y0 + 1 : 0))) * p.W2 + (st0 || st1 ? x : 0);
            const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
            uint4 i00 = z4, i01 = z4, i10 = z4, i11 = z4;
            if (st0) {
              i00 = *reinterpret_cast<const uint4*>(p.head_in + hp * 8);
              i01 = *reinterpret_cast<const uint4*>(p.head_in + (hp + hw2) * 8);
            }
            if (st1) {
              const size_t hq = st0 ? hp + p.W2 : hp;
              i10 = *reinterpret_cast<const uint4*>(p.head_in + hq * 8);
              i11 = *reinterpret_cast<const uint4*>(p.head_in + (hq + hw2) * 8);
            }
            float v0[16], v1[16];
            tc::tmem_ld16_nowait(trow + tb * 64u + (uint32_t)((2 * half) * 16), v0);
            tc::tmem_ld16_nowait(trow + tb * 64u + (uint32_t)((2 * half + 1) * 16), v1);
            tc::tmem_ld_wait(v0, v1);
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
              if (!(rr ? st1 : st0)) continue;
              const float* v = rr ? v1 : v0;
              float im[16];
              unpack8x(rr ? i10 : i00, false, reinterpret_cast<float(&)[8]>(im[0]));
              unpack8x(rr ? i11 : i01, false, reinterpret_cast<float(&)[8]>(im[8]));
              const int y = y0 + rr;
#pragma unroll
              for (int par = 0; par < 4; ++par) {
                const int oy = 2 * y + (par >> 1), ox = 2 * x + (par & 1);
                const float o8[8] = {v[4 * par] + bias[4 * par], v[4 * par + 1] + bias[4 * par + 1],
                                     v[4 * par + 2] + bias[4 * par + 2], v[4 * par + 3] + bias[4 * par + 3],
                                     im[par * 3], im[par * 3 + 1], im[par * 3 + 2], im[12 + par]};
                st8(p.src + (((size_t)f * 2 * p.H2 + oy) * 2 * p.W2 + ox) * SRC_C, o8);
              }
            }
          }
          tc::fence_before();
          if (l < 3) tc::fence_async_smem();   // generic-proxy stores -> the next layer's MMAs (async proxy)
          __syncwarp();
          if (lane == 0) {
            tc::mbar_arrive(tc::smem_u32(&t_empty[tb]));
            if (l < 3) tc::mbar_arrive(tc::smem_u32(&strip_ready[l][k]));
          }
        }
      }
    }
    if (p.dbg && warp == 2 && lane == 0) {
      p.dbg[blockIdx.x * 8 + 4] = t_tf; p.dbg[blockIdx.x * 8 + 5] = clock64() - t_begin; p.dbg[blockIdx.x * 8 + 6] = it;
    }
  }
  tc::fence_before();
  __syncthreads();
  if (warp == 1) {
    tc::fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(HF_NACC * 64))
                 : "memory");
  }
}

using tch::EncodeTiledFn;
using tch::encode_fn;
using tch::opt_in_smem;
using tch::sm_count;

}  // namespace

bool head_fused_fits(const HostLayer* h) {
  if (h[0].kind != CONV_S2D || h[0].taps != 4 || h[0].cin != 16 || h[0].npad != 16) return false;
  for (int l = 1; l < 4; ++l)
    if (h[l].kind != CONV_3x3 || h[l].taps != 9 || h[l].cin != 16 || h[l].npad != 16) return false;
  return h[3].convt_co == 4 && h[1].convt_co == 0 && h[2].convt_co == 0;
}

int head_fused_pack(const HostLayer* h, bool bf16, std::vector<__half>* packed, std::vector<float>* bias) {
  if (!head_fused_fits(h)) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "head_fused_pack: not the 16-channel Head");
  packed->assign(HF_W_BYTES / 2, host_cvt16(0.f, false));
  bias->assign(4 * 16, 0.f);
  size_t off = 0;   // in halves
  for (int l = 0; l < 4; ++l) {
    const int ndx = l == 0 ? 2 : 3, ndy = l == 0 ? 2 : 3, dmax = l == 0 ? 0 : 1;
    const int nblk = 2 * (HF_R - 1) + ndy, nrows = nblk * 16;
    const bool bf = bf16 && l >= 1;   // cnn0 reads the fp16 s2d frame
    for (int xi = 0; xi < ndx; ++xi) {
      const int dx = xi - 1;
      // master strip [k-chunk][row = block * 16 + n][8]: block u holds W(dy = dmax - (u - (R - 1)), dx), zero elsewhere
      for (int u = HF_R - 1; u < HF_R - 1 + ndy; ++u) {
        const int dy = dmax - (u - (HF_R - 1));
        const int t = l == 0 ? (dy + 1) * 2 + (dx + 1) : (dy + 1) * 3 + (dx + 1);
        for (int kc = 0; kc < 2; ++kc)
          for (int n = 0; n < 16; ++n)
            for (int e = 0; e < 8; ++e)
              (*packed)[off + ((size_t)kc * nrows + u * 16 + n) * 8 + e] =
                  host_cvt16(h[l].w[((size_t)t * 16 + kc * 8 + e) * 16 + n], bf);
      }
      off += (size_t)2 * nrows * 8;
    }
    for (int n = 0; n < 16; ++n) (*bias)[l * 16 + n] = h[l].bias[n];
  }
  if (off * 2 != (size_t)HF_W_BYTES) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "head_fused_pack: internal size mismatch");
  return RIFE_OK;
}

int launch_head_fused(const HeadFused& L, const void* head_in, void* src, int F, int H2, int W2, cudaStream_t st) {
  if (!L.w || !L.bias || L.w_bytes != HF_W_BYTES) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "head_fused: weights not packed");
  EncodeTiledFn enc = encode_fn();
  if (!enc) RIFE_FAIL(RIFE_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  HFParams p{};
  p.w = L.w; p.bias = L.bias;
  p.head_in = (const __half*)head_in; p.src = (__half*)src;
  p.F = F; p.H2 = H2; p.W2 = W2;
  p.tiles_x = (W2 + HF_TW - 1) / HF_TW;
  p.tiles_y = (H2 + HF_TH - 1) / HF_TH;
  p.ntiles = p.tiles_x * p.tiles_y * F;
  p.bf = L.bf16 ? 1 : 0;
  CUtensorMap tm_in;
  {
    cuuint64_t dims[4] = {(cuuint64_t)W2 * 4, (cuuint64_t)H2, 2, (cuuint64_t)F};
    cuuint64_t strides[3] = {(cuuint64_t)W2 * 16, (cuuint64_t)H2 * W2 * 16, (cuuint64_t)2 * H2 * W2 * 16};
    cuuint32_t box[4] = {(cuuint32_t)(HF_BP * 4), (cuuint32_t)HF_IN_ROWS, 1, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&tm_in, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, const_cast<void*>(head_in), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) RIFE_FAIL(RIFE_ERR_CUDA, "cuTensorMapEncodeTiled(s2d frames) failed: " + std::to_string((int)r));
  }
  RIFE_TRY(opt_in_smem((const void*)head_fused_kernel, HF_SMEM));
  int grid = sm_count();
  if (grid > p.ntiles) grid = p.ntiles;
  static const bool debug = getenv("RIFE_HF_DEBUG") != nullptr;
  static const int mode = getenv("RIFE_HF_MODE") ? atoi(getenv("RIFE_HF_MODE")) : 0;
  p.mode = mode;
  p.dbg = nullptr;
  if (debug) {
    cudaMalloc((void**)&p.dbg, (size_t)grid * 64);
    cudaMemsetAsync(p.dbg, 0, (size_t)grid * 64, st);
  }
  RIFE_CUDA_TRY(launch_pdl(head_fused_kernel, dim3(grid), dim3(HF_THREADS), HF_SMEM, st, tm_in, p));
  RIFE_LAUNCH_CHECK();
  if (debug) {
    std::vector<long long> h((size_t)grid * 8);
    cudaStreamSynchronize(st);
    cudaMemcpy(h.data(), p.dbg, (size_t)grid * 64, cudaMemcpyDeviceToHost);
    cudaFree(p.dbg);
    double a[8] = {0};
    for (int i = 0; i < grid; ++i)
      for (int k = 0; k < 8; ++k) a[k] += (double)h[(size_t)i * 8 + k] / grid;
    fprintf(stderr, "[hf] grid=%d tiles/cta=%.1f | mma: wait in_full %.0f, strip_ready %.0f, t_empty %.0f, total %.0f | epi(w2): wait t_full %.0f, total %.0f\n",
            grid, a[6], a[0], a[1], a[2], a[3], a[4], a[5]);
  }
  return RIFE_OK;
}

}  // namespace rife
