// Block-input gather fused into conv0.0 of the scale-1 IFBlock (block 4 of the 1080p pyramid; 28 -> 16 channels, stride 2:
// train_log/IFNet_HDv3.py:57-62,79-83,141-152; model/warplayer.py:8-22).
//
// Un-fused, the gather writes the s2d block input `xin` (64 B per full-resolution pixel) to HBM and conv0.0 reads it straight
// back: 2 x 2.1 GB per 16-pair step at 1080p, 19 % of the DRAM traffic of a pass.  Here one persistent CTA per SM
//   * builds the block input of a 16 x 16 tile of the s2d grid (32 x 32 full-resolution pixels + the top row and left
//     column conv0.0's stride-2 3x3 window reaches into: 33 x 33 = 1089 pixels, 6.3 % recomputed) with NG gather warps:
//     lazy flow on top of the fp32 flow buffer, the two backward warps, the concat -- the arithmetic of
//     block_input_h_kernel<1> (pixel_kernels.cu), operation for operation -- and stores every pixel's 4 x 16 B channel groups
//     into shared memory in the planar-8 / tcgen05 no-swizzle K-major canonical layout conv_halo.cu's TMA box would have
//     produced ([plane = parity * 4 + group][17 rows][17 cols][8 ch]);
//   * the tile's window of the flow buffer (33 x 33 x 16 B) and of the newest tmp level (18 x 18 x 32 B: block 3's output at
//     half resolution) arrive by TMA two tiles ahead, so a pixel's dependent chain holds ONE global round trip (the eight
//     warp taps of SRC) instead of three, and the 4x-redundant tap loads of the 2x up-sampling hit shared memory;
//   * one elected thread issues the 2 x 18 tcgen05 MMAs of the tile (2 M-tiles of 16 x 8 s2d pixels; N = 16, K = 16; the 14
//     structurally-zero (tap, parity) products of the s2d formulation are not issued, conv_halo.cu) into a double-buffered
//     TMEM accumulator;
//   * four epilogue warps add bias, apply LeakyReLU(0.2) and store the activation space-to-depth, planar-8, exactly as
//     conv_halo.cu's EPI_LRELU_S2D does (conv0.1 reads it unchanged).
// The MMA and epilogue roles wait with a back-off (they are idle most of a tile's time and would otherwise spend issue slots
// of the gather warps on polling).  Two smem stages: the gather of tile i+1 overlaps the MMAs and the epilogue of tile i.
// A gather warp takes 32-pixel chunks (one interior row of the tile, the halo row, the halo column, the corner) of a
// continuous (tile, chunk) stream -- chunk c of the CTA's t-th tile belongs to warp (c + 35 t) mod NG -- so the 35 chunks of
// a tile spread evenly over any number of warps in steady state; every warp arrives once per tile.
// `xin` is never written; the flow summed over [0, level) goes to a second flow buffer (halo pixels of neighbouring tiles
// read the first one, so it cannot be updated in place).
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>

#include <mutex>
#include <utility>
#include <vector>

#include "conv_gather.cuh"
#include "gather.cuh"
#include "tc_dev.cuh"
#include "tc_host.cuh"

namespace rife {
namespace {

constexpr int GT = 16;                          // tile edge on the s2d grid
constexpr int GP = GT + 1;                      // rows / row pitch (pixels) of a plane: one halo row / column
constexpr int G_PLANE16 = GP * GP;              // 16-byte units per plane
constexpr int G_PLANE_BYTES = G_PLANE16 * 16;   // 4624
constexpr int G_NPLANES = 16;                   // 4 parities x 32 channels / 8
constexpr int G_STAGE_BYTES = G_NPLANES * G_PLANE_BYTES;   // 73984 (a multiple of 128)
constexpr int G_KSTEPS = 8, G_NE = 16;
constexpr int G_W_BYTES = 4 * G_NPLANES * G_NE * 16;       // 4 taps x [cin/8][ne][8] fp16 = 16384
constexpr int G_EDGE = 2 * GT + 1;              // 33 full-resolution rows / columns per tile
constexpr int G_PX = G_EDGE * G_EDGE;           // 1089
constexpr int G_CHUNKS = (G_PX + 31) / 32;      // 35
constexpr int G_CTRL_THREADS = 32 + 128;        // MMA warp + 4 epilogue warps
// staged inputs of one tile: flow-buffer window [33][33] float4 and tmp-level window [18][18][16] fp16
constexpr int G_TE = GT + 2;                    // 18 rows / columns of the half-resolution level under a 33 x 33 tile
// (the two 16-byte halves of a tmp pixel land in separate planes: lanes that read neighbouring pixels then touch consecutive
// 16-byte slots instead of every other one -- no shared-memory bank conflicts on the eight tap loads of a pixel)
constexpr int G_FB_BYTES = G_EDGE * G_EDGE * 16;                // 17424
constexpr int G_FB_AL = (G_FB_BYTES + 127) / 128 * 128;         // 17536
constexpr int G_TM_BYTES = 2 * G_TE * G_TE * 16;                // 10368 bytes arrive per window
constexpr int G_TM_HALF = (G_TE * G_TE * 16 + 127) / 128 * 128; // one plane (channels 0..7 or 8..15), 128-byte aligned for TMA
constexpr int G_TM_AL = 2 * G_TM_HALF;                          // 10496
constexpr int G_IN_BYTES = G_FB_AL + G_TM_AL;                   // per input stage
constexpr int G_SMEM = G_W_BYTES + 2 * G_STAGE_BYTES + 2 * G_IN_BYTES + 128;

struct GParams {
  const __half* src0;
  const __half* src1;
  const float* tstep;
  float* fbuf_out;
  int B, H, W;           // full resolution
  int th, tw;            // the newest tmp level (half resolution)
  int tiles_x, tiles_y, ntiles;
  WarpGeom geom;
  const __half* w;       // conv_halo.cu packing: [tap][cin/8][ne][8]
  const float* bias;
  __half* out;           // [B][8][H/4][W/4][8]
  int out_bf;
  int src_margin;        // >= 0: L2 prefetch of the SRC windows with this many pixels of margin; < 0: off
};

// structural zeros of the stride-2 conv on the s2d input (conv_halo.cu: mma_nonzero<1, 8, E, KK>)
template <int E, int KK>
__device__ __forceinline__ constexpr bool g_nonzero() {
  constexpr int par = KK / (G_KSTEPS / 4), py = par >> 1, px = par & 1;
  constexpr int dy = E / 2 - 1, dx = E % 2 - 1;
  return (dy == 0 || py == 1) && (dx == 0 || px == 1);
}
template <int E, int KK = 0>
__device__ __forceinline__ constexpr int g_first_kk() {
  if constexpr (g_nonzero<E, KK>()) return KK;
  else return g_first_kk<E, KK + 1>();
}
// the 18 MMAs of one M-tile, every operand offset an immediate
template <int E = 0, int KK = 0>
__device__ __forceinline__ void g_issue_tile(uint32_t acc, uint64_t a_base, uint64_t b_base, uint32_t idesc) {
  if constexpr (E < 4) {
    if constexpr (g_nonzero<E, KK>()) {
      constexpr int dy = E / 2 - 1, dx = E % 2 - 1;
      constexpr int a_off = 2 * KK * G_PLANE16 + (dy + 1) * GP + (dx + 1);
      constexpr int b_off = E * (2 * G_KSTEPS) * G_NE + 2 * KK * G_NE;
      constexpr bool first = E == 0 && KK == g_first_kk<0>();
      tc::umma_f16<!first>(acc, a_base + (uint64_t)a_off, b_base + (uint64_t)b_off, idesc);
    }
    if constexpr (KK + 1 < G_KSTEPS) g_issue_tile<E, KK + 1>(acc, a_base, b_base, idesc);
    else g_issue_tile<E + 1, 0>(acc, a_base, b_base, idesc);
  }
}

__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 u;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(a));
  return u;
}
__device__ __forceinline__ void sts128(uint32_t a, const uint4& u) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
}

// level_taps_h1() of gather.cuh (newest level: flow + mask / feat) with the four taps read from the staged window
__device__ __forceinline__ void level_taps_h1_s(uint32_t r0, uint32_t r1, uint32_t c0, uint32_t c1, const float (&wq)[4], float sc,
                                                float (&flow)[4], __half2 (&mfh)[6]) {
  float acc[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t a = ((k >> 1) ? r1 : r0) + ((k & 1) ? c1 : c0);
    const uint4 lo = lds128(a), hi = lds128(a + G_TM_HALF);
    const __half2* hl = reinterpret_cast<const __half2*>(&lo);
    const __half2* hh = reinterpret_cast<const __half2*>(&hi);
    const float2 f01 = __half22float2(hl[0]), f23 = __half22float2(hl[1]);
    const float e[4] = {f01.x, f01.y, f23.x, f23.y};
    const __half2 wk = __float2half2_rn(wq[k]);
    const __half2 v[6] = {hl[2], hl[3], hh[0], hh[1], hh[2], hh[3]};   // channels 4..15
#pragma unroll
    for (int i = 0; i < 6; ++i) mfh[i] = k == 0 ? __hmul2(wk, v[i]) : __hfma2(wk, v[i], mfh[i]);
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[c] = k == 0 ? wq[0] * e[c] : fmaf(wq[k], e[c], acc[c]);
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) flow[c] = fmaf(acc[c], sc, flow[c]);
}

template <int NG>
__global__ void __launch_bounds__(G_CTRL_THREADS + 32 * NG, 1)
gather_conv00_kernel(const __grid_constant__ CUtensorMap tm_fb, const __grid_constant__ CUtensorMap tm_tmp,
                     const __grid_constant__ CUtensorMap tm_s0, const __grid_constant__ CUtensorMap tm_s1,
                     const __grid_constant__ GParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[2];
  __shared__ __align__(8) uint64_t a_empty[2];
  __shared__ __align__(8) uint64_t t_full[2];
  __shared__ __align__(8) uint64_t t_empty[2];
  __shared__ __align__(8) uint64_t in_full[2];
  __shared__ __align__(8) uint64_t w_full;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float s_bias[G_NE];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t w_smem = (tc::smem_u32(smem_raw) + 127u) & ~127u;
  const uint32_t a_smem = w_smem + G_W_BYTES;
  const uint32_t in_smem = a_smem + 2 * G_STAGE_BYTES;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_fb) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_tmp) : "memory");
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(tc::smem_u32(&a_full[i]), NG);   // one arrive per gather warp
      tc::mbar_init(tc::smem_u32(&a_empty[i]), 1);   // tcgen05.commit
      tc::mbar_init(tc::smem_u32(&t_full[i]), 1);    // tcgen05.commit
      tc::mbar_init(tc::smem_u32(&t_empty[i]), 4);   // one arrive per epilogue warp
      tc::mbar_init(tc::smem_u32(&in_full[i]), 1);   // TMA transaction bytes
    }
    tc::mbar_init(tc::smem_u32(&w_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(&tmem_slot)), "r"(64u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x >= 32 && threadIdx.x < 32 + G_NE) s_bias[threadIdx.x - 32] = p.bias[threadIdx.x - 32];
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  const uint32_t tmem_base = tmem_slot;
  const int tiles_per_img = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- weights (constants: loaded before the dependency wait), input windows two tiles ahead, the MMAs
      const uint32_t wb = tc::smem_u32(&w_full);
      tc::mbar_expect_tx(wb, (uint32_t)G_W_BYTES);
      tc::bulk_load(w_smem, p.w, (uint32_t)G_W_BYTES, wb);
      pdl_wait();   // the flow buffer and the tmp level are earlier kernels' outputs (common.cuh: PDL contract)
      auto prefetch = [&](int tile, int s) {
        const int b = tile / tiles_per_img, r = tile - b * tiles_per_img;
        const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
        const uint32_t bar = tc::smem_u32(&in_full[s]);
        const uint32_t dst = in_smem + (uint32_t)s * G_IN_BYTES;
        tc::mbar_expect_tx(bar, (uint32_t)(G_FB_BYTES + G_TM_BYTES));
        // flow buffer [B][H][W][4] fp32 as {4W, H, B}: 33 pixels x 33 rows from (32 tx - 1, 32 ty - 1); outside = zero fill
        tc::tma_load_3d(dst, &tm_fb, bar, 4 * (tx * 2 * GT - 1), ty * 2 * GT - 1, b);
        // tmp level [B][th][tw][16] fp16 as {4 words, 2 halves, tw, th, B}: 18 x 18 pixels from (16 tx - 1, 16 ty - 1), one
        // plane per 16-byte half
        tc::tma_load_5d(dst + G_FB_AL, &tm_tmp, bar, 0, 0, tx * GT - 1, ty * GT - 1, b);
        tc::tma_load_5d(dst + G_FB_AL + G_TM_HALF, &tm_tmp, bar, 0, 1, tx * GT - 1, ty * GT - 1, b);
        if (p.src_margin >= 0) {
          // the tile's window of both SRC frames (+ a margin for the flow) on its way into L2: the warp taps then pay an
          // L2 hit instead of a DRAM round trip.  A hint only -- taps beyond the margin go to DRAM as before.
          const int fr = p.geom.src_div == 1 ? b : b / p.geom.src_div;
          tc::tma_prefetch_3d(&tm_s0, 4 * (tx * 2 * GT - 1 - p.src_margin), ty * 2 * GT - 1 - p.src_margin, fr);
          tc::tma_prefetch_3d(&tm_s1, 4 * (tx * 2 * GT - 1 - p.src_margin), ty * 2 * GT - 1 - p.src_margin, fr);
        }
      };
      if ((int)blockIdx.x < p.ntiles) prefetch(blockIdx.x, 0);
      if ((int)(blockIdx.x + gridDim.x) < p.ntiles) prefetch(blockIdx.x + gridDim.x, 1);
      tc::mbar_wait(wb, 0);
      // instruction descriptor: D = f32, A / B = f16, N >> 3, M >> 4
      const uint32_t idesc = (1u << 4) | ((uint32_t)(G_NE >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint64_t a_desc0 = tc::make_desc_noswz(0, (uint32_t)G_PLANE_BYTES, (uint32_t)GP * 16u);
      const uint64_t b_base = tc::make_desc_noswz(w_smem, (uint32_t)G_NE * 16u, 128u);
      int it = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
        const int s = it & 1;
        const uint32_t ph = ((uint32_t)it >> 1) & 1u;
        tc::mbar_wait_backoff(tc::smem_u32(&a_full[s]), ph);
        // every gather warp is done with tile `it`, hence with input stage s: refill it for tile it + 2
        if (tile + 2 * (int)gridDim.x < p.ntiles) prefetch(tile + 2 * (int)gridDim.x, s);
        tc::mbar_wait_backoff(tc::smem_u32(&t_empty[s]), ph ^ 1u);
        tc::fence_after();
        const uint32_t sa16 = (a_smem + (uint32_t)s * G_STAGE_BYTES) >> 4;
#pragma unroll
        for (int cx = 0; cx < 2; ++cx)
          g_issue_tile<>(tmem_base + (uint32_t)(s * 2 * G_NE + cx * G_NE), a_desc0 + (uint64_t)(sa16 + (uint32_t)cx * 8u), b_base,
                         idesc);
        tc::umma_commit(tc::smem_u32(&a_empty[s]));
        tc::umma_commit(tc::smem_u32(&t_full[s]));
      }
    }
  } else if (warp < 5) {
    // ---------------- epilogue: TMEM lane quarter warp % 4; thread = GEMM row m = ty_ * 8 + tx_ of both M-tiles
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int ty_ = m >> 3, tx_ = m & 7;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    const int H2 = p.H >> 1, W2 = p.W >> 1, H4 = p.H >> 2, W4 = p.W >> 2;
    const size_t pl = (size_t)H4 * W4 * 8;   // elements per planar-8 group of the destination
    float bias[G_NE];
#pragma unroll
    for (int i = 0; i < G_NE; ++i) bias[i] = s_bias[i];
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int s = it & 1;
      const uint32_t ph = ((uint32_t)it >> 1) & 1u;
      const int b = tile / tiles_per_img, r = tile - b * tiles_per_img;
      const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
      tc::mbar_wait_backoff(tc::smem_u32(&t_full[s]), ph);
      tc::fence_after();
#pragma unroll
      for (int cx = 0; cx < 2; ++cx) {
        float v[16];
        tc::tmem_ld16(trow + (uint32_t)(s * 2 * G_NE + cx * G_NE), v);
        const int y = ty * GT + ty_, x = tx * GT + cx * 8 + tx_;   // s2d grid
        if (y < H2 && x < W2) {
          // EPI_LRELU_S2D of conv_halo.cu: planar-8 on the quarter-resolution grid, channel = parity * 16 + n
          const int par = (y & 1) * 2 + (x & 1);
          __half* dst = p.out + ((size_t)b * 8 + par * 2) * pl + ((size_t)(y >> 1) * W4 + (x >> 1)) * 8;
          float a[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) a[i] = lrelu02(v[i] + bias[i]);
          st8x(dst, a, p.out_bf);
#pragma unroll
          for (int i = 0; i < 8; ++i) a[i] = lrelu02(v[8 + i] + bias[8 + i]);
          st8x(dst + pl, a, p.out_bf);
        }
      }
      tc::fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(tc::smem_u32(&t_empty[s]));
    }
  } else {
    // ---------------- gather warps
    pdl_wait();   // SRC is an earlier kernel's output
    const int gw = warp - 5;
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
    const int HW = p.H * p.W;
    // (b, ty, tx) of the CTA's tiles advance by a fixed step: no integer division per tile in the 27 gather warps
    int b = blockIdx.x / tiles_per_img, ty, tx;
    {
      const int r = blockIdx.x - b * tiles_per_img;
      ty = r / p.tiles_x; tx = r - ty * p.tiles_x;
    }
    const int db = gridDim.x / tiles_per_img, dr = gridDim.x - db * tiles_per_img;
    const int dty = dr / p.tiles_x, dtx = dr - dty * p.tiles_x;
    int first = gw;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      if (it > 0) {
        tx += dtx; ty += dty; b += db;
        if (tx >= p.tiles_x) { tx -= p.tiles_x; ++ty; }
        if (ty >= p.tiles_y) { ty -= p.tiles_y; ++b; }
        first -= G_CHUNKS % NG;   // chunk c of the CTA's t-th tile belongs to warp (c + 35 t) mod NG
        if (first < 0) first += NG;
      }
      const int s = it & 1;
      const uint32_t ph = ((uint32_t)it >> 1) & 1u;
      const int y0 = ty * 2 * GT - 1, x0 = tx * 2 * GT - 1;   // full-resolution position of the halo corner
      const int ty0 = ty * GT - 1, tx0 = tx * GT - 1;         // origin of the staged tmp window
      const int fr = p.geom.src_div == 1 ? b : b / p.geom.src_div;
      const __half* s0 = p.src0 + (size_t)fr * HW * SRC_C;
      const __half* s1 = p.src1 + (size_t)fr * HW * SRC_C;
      float* fbo = p.fbuf_out + (size_t)b * HW * 4;
      const __half tsh = __float2half_rn(p.tstep[b]);
      const uint32_t stage = a_smem + (uint32_t)s * G_STAGE_BYTES;
      const uint32_t fbs = in_smem + (uint32_t)s * G_IN_BYTES, tms = fbs + G_FB_AL;
      tc::mbar_wait(tc::smem_u32(&a_empty[s]), ph ^ 1u);   // the MMAs of the tile two steps back have read this stage
      tc::mbar_wait(tc::smem_u32(&in_full[s]), ph);        // this tile's flow-buffer and tmp windows have landed
#pragma unroll 1
      for (int c = first; c < G_CHUNKS; c += NG) {
        // chunks 0..31: interior rows (ry = c + 1, rx = 1..32: 16-byte-slot aligned shared-memory stores, 512-byte aligned
        // flow-buffer stores); 32: the top halo row (rx = 1..32); 33: the left halo column (ry = 1..32); 34: the corner
        int ry, rx;
        if (c < 2 * GT) { ry = c + 1; rx = lane + 1; }
        else if (c == 2 * GT) { ry = 0; rx = lane + 1; }
        else if (c == 2 * GT + 1) { ry = lane + 1; rx = 0; }
        else { ry = 0; rx = lane == 0 ? 0 : -1; }
        if (rx >= 0) {
          const int idx = ry * G_EDGE + rx;
          const int y = y0 + ry, x = x0 + rx;
          uint4 u[4];
          if (y >= 0 && x >= 0 && y < p.H && x < p.W) {
            // ---- gather_s1_px() of gather.cuh for one coarser level on top of the flow buffer (level = fbuf_levels + 1)
            float flow[4];
            {
              const uint4 f = lds128(fbs + (uint32_t)idx * 16u);
              flow[0] = __uint_as_float(f.x); flow[1] = __uint_as_float(f.y);
              flow[2] = __uint_as_float(f.z); flow[3] = __uint_as_float(f.w);
            }
            __half2 mfh[6];
            {
              const Lerp2 ly = make_lerp2(y, p.th, 0.5f), lx = make_lerp2(x, p.tw, 0.5f);
              const float ay = ly.a0, ax = lx.a0;
              const float wq[4] = {(1.f - ay) * (1.f - ax), (1.f - ay) * ax, ay * (1.f - ax), ay * ax};
              level_taps_h1_s(tms + (uint32_t)((ly.i0 - ty0) * (G_TE * 16)), tms + (uint32_t)((ly.i1 - ty0) * (G_TE * 16)),
                              (uint32_t)((lx.i0 - tx0) * 16), (uint32_t)((lx.i1 - tx0) * 16), wq, 2.f, flow, mfh);
            }
            if (ry > 0 && rx > 0) st4(fbo + (y * p.W + x) * 4, flow);
            __half2 A[4], Cc[4];
            warp8_h2(s0, p.H, p.W, p.geom, x, y, flow[0], flow[1], A);   // IFNet_HDv3.py:145,155
            warp8_h2(s1, p.H, p.W, p.geom, x, y, flow[2], flow[3], Cc);  // IFNet_HDv3.py:146,156
            __half2* h0 = reinterpret_cast<__half2*>(&u[0]);
            __half2* h1 = reinterpret_cast<__half2*>(&u[1]);
            __half2* h2 = reinterpret_cast<__half2*>(&u[2]);
            __half2* h3 = reinterpret_cast<__half2*>(&u[3]);
            h0[0] = A[2];                                                      // warped img0 r, g
            h0[1] = __halves2half2(__low2half(A[3]), __low2half(Cc[2]));       // img0 b, img1 r
            h0[2] = __halves2half2(__high2half(Cc[2]), __low2half(Cc[3]));     // img1 g, b
            h0[3] = A[0]; h1[0] = A[1];                                        // warped f0
            h1[1] = Cc[0]; h1[2] = Cc[1];                                      // warped f1
            h1[3] = __halves2half2(tsh, __low2half(mfh[0]));                   // timestep, mask
            h2[0] = __halves2half2(__high2half(mfh[0]), __low2half(mfh[1]));   // feat 0..7
            h2[1] = __halves2half2(__high2half(mfh[1]), __low2half(mfh[2]));
            h2[2] = __halves2half2(__high2half(mfh[2]), __low2half(mfh[3]));
            h2[3] = __halves2half2(__high2half(mfh[3]), __low2half(mfh[4]));
            h3[0] = __floats2half2_rn(flow[0], flow[1]);   // interpolate(flow, 1/s) * 1/s with s = 1 (IFNet_HDv3.py:81)
            h3[1] = __floats2half2_rn(flow[2], flow[3]);
            h3[2] = __float2half2_rn(0.f);
            h3[3] = h3[2];
          } else {
            // conv0.0's zero padding (top / left), or beyond the image (bottom / right: only feeds outputs that are not stored)
            u[0] = zero4; u[1] = zero4; u[2] = zero4; u[3] = zero4;
          }
          // s2d pixel ((ry + 1) >> 1, (rx + 1) >> 1) of the stage (row / column 0 = halo), parity ((ry + 1) & 1, (rx + 1) & 1)
          const int par = ((ry + 1) & 1) * 2 + ((rx + 1) & 1);
          const uint32_t a = stage + (uint32_t)(par * 4 * G_PLANE_BYTES + (((ry + 1) >> 1) * GP + ((rx + 1) >> 1)) * 16);
#pragma unroll
          for (int k = 0; k < 4; ++k) sts128(a + (uint32_t)(k * G_PLANE_BYTES), u[k]);
        }
      }
      tc::fence_async_smem();   // generic-proxy stores -> the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(tc::smem_u32(&a_full[s]));
    }
  }
  tc::fence_before();
  __syncthreads();
  if (warp == 0) {
    tc::fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64u) : "memory");
  }
}

using tch::EncodeTiledFn;
using tch::encode_fn;
using tch::opt_in_smem;
using tch::sm_count;

template <int NG>
int launch_ng(const CUtensorMap& tm_fb, const CUtensorMap& tm_tmp, const CUtensorMap& tm_s0, const CUtensorMap& tm_s1,
              const GParams& p, cudaStream_t st) {
  RIFE_TRY(opt_in_smem((const void*)gather_conv00_kernel<NG>, G_SMEM));
  int grid = sm_count();
  if (grid > p.ntiles) grid = p.ntiles;
  RIFE_CUDA_TRY(launch_pdl(gather_conv00_kernel<NG>, dim3(grid), dim3(G_CTRL_THREADS + 32 * NG), G_SMEM, st, tm_fb, tm_tmp, tm_s0, tm_s1,
                           p));
  RIFE_LAUNCH_CHECK();
  return RIFE_OK;
}

}  // namespace

bool gather_conv_fits(const HaloLayer& L) {
  return L.kind == CONV_S2D && L.cin == 4 * XIN_C && L.nentries == 4 && L.ne == G_NE && L.ncols == G_NE && !L.stream &&
         L.w_bytes == G_W_BYTES && L.ps_cols == 0;
}

bool gather_conv_levels_fit(int level, const TmpLevel* levels, int fbuf_levels, int H, int W) {
  return level >= 2 && fbuf_levels == level - 1 && levels[level - 1].s == 2 && levels[level - 1].pitch == TMP_C &&
         levels[level - 1].h * 2 == H && levels[level - 1].w * 2 == W;
}

int launch_gather_conv00(const PixelCtx& c, int level, const TmpLevel* levels, const float* fbuf_in, float* fbuf_out,
                         int fbuf_levels, const HaloLayer& conv00, void* out, int out_bf16, cudaStream_t st) {
  if (!gather_conv_fits(conv00)) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "gather_conv00: conv0.0 is not the 4 x 32 -> 16 channel s2d layer");
  if (!c.half_act || c.any_up) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "gather_conv00: fp16 / bf16 pipeline, scale-1 block only");
  if (c.H % 4 || c.W % 4) RIFE_FAIL(RIFE_ERR_SHAPE, "gather_conv00: H and W must be multiples of 4");
  if (!gather_conv_levels_fit(level, levels, fbuf_levels, c.H, c.W))
    RIFE_FAIL(RIFE_ERR_INVALID_ARG, "gather_conv00: needs the flow buffer over [0, level - 1) and one half-resolution tmp level on top");
  if (!fbuf_in || !fbuf_out || fbuf_out == fbuf_in) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "gather_conv00: two distinct flow buffers");
  EncodeTiledFn enc = encode_fn();
  if (!enc) RIFE_FAIL(RIFE_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  const TmpLevel& tl = levels[level - 1];
  GParams p{};
  p.src0 = (const __half*)c.src0; p.src1 = (const __half*)c.src1;
  p.tstep = c.tstep;
  p.fbuf_out = fbuf_out;
  p.B = c.B; p.H = c.H; p.W = c.W;
  p.th = tl.h; p.tw = tl.w;
  p.tiles_x = (c.W / 2 + GT - 1) / GT;
  p.tiles_y = (c.H / 2 + GT - 1) / GT;
  p.ntiles = p.tiles_x * p.tiles_y * c.B;
  p.geom = make_geom(c.H, c.W);
  p.geom.src_div = c.src_div;
  p.w = conv00.w; p.bias = conv00.bias;
  p.out = (__half*)out; p.out_bf = out_bf16 ? 1 : 0;
  CUtensorMap tm_fb, tm_tmp;
  {
    cuuint64_t dims[3] = {(cuuint64_t)c.W * 4, (cuuint64_t)c.H, (cuuint64_t)c.B};
    cuuint64_t strides[2] = {(cuuint64_t)c.W * 16, (cuuint64_t)c.H * c.W * 16};
    cuuint32_t box[3] = {(cuuint32_t)(G_EDGE * 4), (cuuint32_t)G_EDGE, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&tm_fb, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(fbuf_in), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) RIFE_FAIL(RIFE_ERR_CUDA, "cuTensorMapEncodeTiled(flow buffer) failed: " + std::to_string((int)r));
  }
  {
    // [B][th][tw][16] fp16 seen as 32-bit words: {4 words = 16 B, 2 halves, tw, th, B}; one box = one half of 18 x 18 pixels
    cuuint64_t dims[5] = {4, 2, (cuuint64_t)tl.w, (cuuint64_t)tl.h, (cuuint64_t)c.B};
    cuuint64_t strides[4] = {16, 32, (cuuint64_t)tl.w * 32, (cuuint64_t)tl.h * tl.w * 32};
    cuuint32_t box[5] = {4, 1, (cuuint32_t)G_TE, (cuuint32_t)G_TE, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&tm_tmp, CU_TENSOR_MAP_DATA_TYPE_UINT32, 5, const_cast<void*>(tl.ptr), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) RIFE_FAIL(RIFE_ERR_CUDA, "cuTensorMapEncodeTiled(tmp level) failed: " + std::to_string((int)r));
  }
  // SRC frames [F][H][W][8] fp16 as {4 W words, H, F}: prefetch box = (33 + 2 margin)^2 pixels
  static const int margin = getenv("RIFE_GC_SRC_MARGIN") ? atoi(getenv("RIFE_GC_SRC_MARGIN")) : 4;
  p.src_margin = margin > 15 ? 15 : margin;
  CUtensorMap tm_s0 = tm_fb, tm_s1 = tm_fb;
  if (p.src_margin >= 0) {
    const int nfr = (c.B + c.src_div - 1) / c.src_div;   // frames of src0 / src1 this (sub-)batch warps
    const int e = G_EDGE + 2 * p.src_margin;
    cuuint64_t strides[2] = {(cuuint64_t)c.W * 16, (cuuint64_t)c.H * c.W * 16};
    cuuint32_t box[3] = {(cuuint32_t)(e * 4), (cuuint32_t)e, 1};
    cuuint32_t es[3] = {1, 1, 1};
    for (int k = 0; k < 2; ++k) {
      cuuint64_t dims[3] = {(cuuint64_t)c.W * 4, (cuuint64_t)c.H, (cuuint64_t)nfr};
      CUresult r = enc(k ? &tm_s1 : &tm_s0, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<void*>(k ? c.src1 : c.src0), dims,
                       strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) RIFE_FAIL(RIFE_ERR_CUDA, "cuTensorMapEncodeTiled(SRC) failed: " + std::to_string((int)r));
    }
  }
  // gather warps per CTA (one CTA per SM): 27 = a full 1024-thread CTA at 64 registers per thread
  static const int ng = getenv("RIFE_GC_WARPS") ? atoi(getenv("RIFE_GC_WARPS")) : 27;
  if (ng <= 16) return launch_ng<16>(tm_fb, tm_tmp, tm_s0, tm_s1, p, st);
  if (ng <= 20) return launch_ng<20>(tm_fb, tm_tmp, tm_s0, tm_s1, p, st);
  if (ng <= 24) return launch_ng<24>(tm_fb, tm_tmp, tm_s0, tm_s1, p, st);
  return launch_ng<27>(tm_fb, tm_tmp, tm_s0, tm_s1, p, st);
}

}  // namespace rife
