// Persistent halo-tile tcgen05 convolution for the thin layers (C <= 64: Head, IFBlock 3 and 4 -- 75% of the
// network's FLOPs and nearly all of its conv bytes; train_log/IFNet_HDv3.py:24-43, 57-76).
//
// Why a second conv kernel: with C <= 64 the per-tap operand tiles are tiny, so a kernel that re-loads the
// activation tile once per tap (conv_umma.cu) moves 9x the activation bytes L2->SM and exposes a TMA->MMA->epilogue
// latency chain per 128-pixel tile.  Here
//   * activations live in HBM as [B][C/8][H][W][8] fp16 ("planar-8"); ONE TMA box {(TW+R)*8, TH+R, C/8, 1} brings
//     the whole (TH+R) x (TW+R) halo of a 16 x 8 pixel tile into smem as C/8 planes of [rows][cols][8ch];
//   * that layout is exactly the tcgen05 no-swizzle K-major canonical layout when the tile is 8 pixels wide
//     (8 consecutive pixels = one 8x16B core matrix, SBO = halo row pitch, LBO = plane pitch), so every tap of the
//     3x3 / 2x2 stencil is the SAME smem tile read through a descriptor whose start address is shifted by
//     (dy * pitch + dx) * 16 bytes -- no im2col, no re-load;
//   * the layer's weights stay resident in smem for the CTA's lifetime; CTAs are persistent (one or two per SM)
//     and walk tiles with a TMA ring several tiles deep and two TMEM accumulator buffers, so the epilogue of tile
//     i overlaps the MMAs of tile i+1 and the loads of tiles i+2..;
//   * ConvTranspose2d(4,2,1)+PixelShuffle (lastconv) runs as 4 output parities x 4 taps (16 MMAs groups of N=64)
//     instead of 9 dense taps of N=208/256.
#include <cuda.h>
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <mutex>
#include <utility>

#include "conv_halo.cuh"

namespace rife {
namespace {

constexpr int TW = 8, TH = 16;  // M-tile: 16 rows x 8 pixels = 128 GEMM rows, row m = ty*8 + tx
constexpr int MAX_MMA = 128;    // (schedule entries) x (K steps) per M-tile
constexpr int MAX_THREADS = 64 + 32 * 8;  // 2 control warps + 4 or 8 epilogue warps
constexpr int MAX_STAGES = 8;
constexpr int MAX_ENTRIES = 16;
constexpr int SMEM_BUDGET = 220 * 1024;

struct HaloParams {
  int B, H, W, tiles_x, tiles_y, ntiles;
  int kgroups, ksteps, R, mt;  // mt = vertically stacked M-tiles per pipeline step (they share one halo load)
  int plane_bytes, a_stage_bytes, a_tx_bytes, stages;
  int nentries, ne, ncols, tmem_cols, n_valid, w_bytes, wtile_bytes;
  int nb, wtile_al;  // weight streaming: ring of nb tiles of wtile_al bytes (0 = weights resident)
  int epi, out_c, cpp;
  int in_bf, out_bf;  // operand / activation-output storage type: 0 = fp16, 1 = bf16 (RIFE_PRECISION_BF16)
  const __half* w;
  const float* bias;
  const float* beta;
  const __half* res;
  __half* out;
  // ResConv chain (nlayers > 1): the same c -> c layer shape applied nlayers times inside ONE cooperative launch, ping-pong
  // between two activation tensors, a grid-wide barrier (atomic counter + spin) instead of a kernel boundary between layers
  int nlayers;
  const __half* w_l[8];
  const float* bias_l[8];
  const float* beta_l[8];
  __half* out_l[2];        // destination of even / odd layers
  unsigned int* sync;      // grid barrier: sync[0] arrival counter, sync[1] finished-CTA counter; both zero when a launch
                           // starts and again when it ends (the last CTA to finish resets them), so a launch carries no
                           // host-side state and can be replayed from a CUDA graph
  unsigned int sync_base;  // always 0 (kept for the wait arithmetic)
  unsigned int* tile_ctr;  // dynamic tile scheduling: zero-initialised counter of this launch (null: static round robin)
  int pdl_trig;    // launch_dependents: 0 implicit at CTA exit (default; measured best), 1 after the prologue, 2 after the last MMA
  long long* dbg;  // RIFE_HK_DEBUG=1: per-CTA role timers (8 x int64), else null
  signed char e_dy[MAX_ENTRIES], e_dx[MAX_ENTRIES];
  unsigned char e_col16[MAX_ENTRIES], e_first[MAX_ENTRIES];
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
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
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

// no-swizzle K-major descriptor: 8 rows x 16 B core matrices; LBO = byte stride between the two 8-element K chunks
// of one MMA, SBO = byte stride between consecutive 8-row groups
__device__ __forceinline__ uint64_t make_desc_noswz(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version 1 (sm_100); layout type 0 = SWIZZLE_NONE
  return d;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// the same load without its wait, and the wait for all loads issued so far.  tmem_regs_ready() ties the loaded registers to
// the wait (volatile asm statements keep their order; the empty one makes every later use depend on it), so that several
// accumulator chunks can be in flight before one wait.
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
        "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait_all() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_regs_ready(float (&v)[16]) {
  asm volatile(""
               : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]), "+f"(v[8]),
                 "+f"(v[9]), "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15])
               :
               : "memory");
}
__device__ __forceinline__ void store16h(__half* dst, const float (&v)[16]) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = v[i];
  st8(dst, a);
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = v[8 + i];
  st8(dst + 8, a);
}

// grid-wide barrier of the ResConv chain: every CTA adds 1 after its layer-l stores are visible; the TMA producer of
// layer l + 1 spins until all gridDim.x CTAs have arrived (all CTAs are co-resident: cooperative launch)
__device__ __forceinline__ void grid_arrive(unsigned int* ctr) {
  __threadfence();
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
}
__device__ __forceinline__ void grid_wait(const unsigned int* ctr, unsigned int target) {
  const long long t0 = clock64();
  while (true) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    if ((int)(v - target) >= 0) break;   // wrap-safe
    if (clock64() - t0 > 4000000000LL) __trap();
    __nanosleep(64);
  }
  asm volatile("fence.proxy.async;" ::: "memory");  // other CTAs' generic-proxy stores -> this CTA's TMA (async proxy) loads
}

__device__ __forceinline__ int ring_next_tile(const int* ring, const uint64_t* bars, int seq) {
  mbar_wait(smem_u32(&bars[seq & 15]), ((uint32_t)seq >> 4) & 1u);
  return *reinterpret_cast<const volatile int*>(&ring[seq & 15]);
}

template <int I> struct IntC { static constexpr int value = I; };
template <int N, int I = 0, class F>
__device__ __forceinline__ void static_for(F& f) {
  if constexpr (I < N) {
    f(IntC<I>{});
    static_for<N, I + 1>(f);
  }
}

// Compile-time MMA schedules (must match halo_pack_weights): KIND 0 = 3x3 taps, 1 = 2x2 taps on the s2d grid,
// 2 = ConvTranspose(4,2,1)+PixelShuffle as (output parity, tap) pairs.
template <int KIND> struct Sched;
template <> struct Sched<0> {
  static constexpr int NENT = 9, R = 2;
  static constexpr int dy(int e) { return e / 3 - 1; }
  static constexpr int dx(int e) { return e % 3 - 1; }
  static constexpr int col16(int) { return 0; }
  static constexpr bool first(int e) { return e == 0; }
};
template <> struct Sched<1> {
  static constexpr int NENT = 4, R = 1;
  static constexpr int dy(int e) { return e / 2 - 1; }
  static constexpr int dx(int e) { return e % 2 - 1; }
  static constexpr int col16(int) { return 0; }
  static constexpr bool first(int e) { return e == 0; }
};
template <> struct Sched<2> {
  static constexpr int NENT = 16, R = 2;
  static constexpr int dy(int e) { return ((e >> 2) >> 1) ? ((e >> 1) & 1) : ((e >> 1) & 1) - 1; }
  static constexpr int dx(int e) { return ((e >> 2) & 1) ? (e & 1) : (e & 1) - 1; }
  static constexpr int col16(int e) { return (e >> 2) * 4; }
  static constexpr bool first(int e) { return (e & 3) == 0; }
};

template <bool ACC>
__device__ __forceinline__ void umma_f16_c(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc) {
  if constexpr (ACC)
    asm volatile("tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, 1;" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc) : "memory");
  else
    asm volatile("tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, 0;" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc) : "memory");
}

// Stride-2 3x3 convs run as 2x2 taps over the space-to-depth input, channel = parity * cs + c (api.cu: pack_conv3x3_s2d).
// Only 9 of the 16 (tap, parity) pairs carry weights -- tap (dy, dx) reads parity (py, px) iff (dy == 0 || py == 1) and
// (dx == 0 || px == 1); the other 7 are structural zeros.  With cs a multiple of 16 every K-step lies inside one parity, so
// 14 of the 32 MMAs of a tile (KSTEPS = 8) multiply by zero and are not issued at all.
template <int KIND, int KSTEPS, int E, int KK>
__device__ __forceinline__ constexpr bool mma_nonzero() {
  if constexpr (KIND == 1 && KSTEPS % 4 == 0) {
    constexpr int par = KK / (KSTEPS / 4), py = par >> 1, px = par & 1;
    return (Sched<1>::dy(E) == 0 || py == 1) && (Sched<1>::dx(E) == 0 || px == 1);
  } else {
    return true;
  }
}
template <int KIND, int KSTEPS, int E, int KK = 0>
__device__ __forceinline__ constexpr int first_nonzero_kk() {
  if constexpr (mma_nonzero<KIND, KSTEPS, E, KK>()) return KK;
  else return first_nonzero_kk<KIND, KSTEPS, E, KK + 1>();
}

// fully unrolled issue of one M-tile's MMAs: every operand offset is an immediate
template <int KIND, int KSTEPS, int NE, int MT, int E = 0, int KK = 0>
__device__ __forceinline__ void issue_tile(uint32_t acc, uint64_t a_base, uint64_t b_base, uint32_t idesc) {
  using S = Sched<KIND>;
  if constexpr (E < S::NENT) {
    if constexpr (mma_nonzero<KIND, KSTEPS, E, KK>()) {
      constexpr int pitch = TW + S::R;
      constexpr int plane16 = (TH * MT + S::R) * pitch;
      constexpr int a_off = 2 * KK * plane16 + (S::dy(E) + 1) * pitch + (S::dx(E) + 1);
      constexpr int b_off = E * (2 * KSTEPS) * NE + 2 * KK * NE;
      constexpr int d_col = KIND == 2 ? (E >> 2) * NE : S::col16(E) * 16;  // parity schedule: one NE-column group per output parity
      constexpr bool first = S::first(E) && KK == first_nonzero_kk<KIND, KSTEPS, E>();   // overwrites the accumulator
      umma_f16_c<!first>(acc + d_col, a_base + (uint64_t)a_off, b_base + (uint64_t)b_off, idesc);
    }
    if constexpr (KK + 1 < KSTEPS) issue_tile<KIND, KSTEPS, NE, MT, E, KK + 1>(acc, a_base, b_base, idesc);
    else issue_tile<KIND, KSTEPS, NE, MT, E + 1, 0>(acc, a_base, b_base, idesc);
  }
}


// weight-streaming variant: the MMAs of ONE schedule entry (one weight tile) for all MT stacked M-tiles
template <int KIND, int KSTEPS, int NE, int MT, int E, int M = 0, int KK = 0>
__device__ __forceinline__ void issue_entry(uint32_t acc0, uint32_t ncols, uint64_t a_base0, uint32_t mt_step,
                                            uint64_t b_base, uint32_t idesc, int mt_eff) {
  using S = Sched<KIND>;
  if constexpr (M < MT) {
    if (M == 0 || M < mt_eff) {  // stacked M-tiles below the image's last row are not computed
      constexpr int pitch = TW + S::R;
      constexpr int plane16 = (TH * MT + S::R) * pitch;
      constexpr int a_off = 2 * KK * plane16 + (S::dy(E) + 1) * pitch + (S::dx(E) + 1);
      constexpr int d_col = KIND == 2 ? (E >> 2) * NE : S::col16(E) * 16;
      umma_f16_c<!(S::first(E) && KK == 0)>(acc0 + (uint32_t)M * ncols + d_col,
                                            a_base0 + (uint64_t)((uint32_t)M * mt_step + a_off), b_base + (uint64_t)(2 * KK * NE),
                                            idesc);
    }
    if constexpr (KK + 1 < KSTEPS) issue_entry<KIND, KSTEPS, NE, MT, E, M, KK + 1>(acc0, ncols, a_base0, mt_step, b_base, idesc, mt_eff);
    else issue_entry<KIND, KSTEPS, NE, MT, E, M + 1, 0>(acc0, ncols, a_base0, mt_step, b_base, idesc, mt_eff);
  }
}

// Per-tile epilogue context of one thread (= one pixel column position tx_ of the tile; rows come from (mi, ty_)).
struct EpiCtx {
  int b, x, y0;           // batch, pixel x, first row of this thread in the step (mi = 0)
  uint32_t res_smem;      // smem address of the residual for (mi = 0, plane 0)
  size_t plane;           // elements per planar-8 channel group of the destination
  __half* out;            // destination tensor of this layer
};

// Fused layer tail for 16 accumulator columns [16j, 16j+16) of M-tile `mi` (bias already in smem).
template <int EPI>
__device__ __forceinline__ void epi_chunk(const HaloParams& p, const EpiCtx& c, const float* s_bias, const float* s_beta,
                                          int pitch, int mi, int j, float (&v)[16]) {
  const int n0 = j * 16;
  const int y = c.y0 + mi * TH;
  if (y >= p.H || c.x >= p.W) return;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 bb = *reinterpret_cast<const float4*>(s_bias + n0 + 4 * i);
    v[4 * i] += bb.x; v[4 * i + 1] += bb.y; v[4 * i + 2] += bb.z; v[4 * i + 3] += bb.w;
  }
  if constexpr (EPI == EPI_LRELU || EPI == EPI_LINEAR || EPI == EPI_LRELU_S2D) {
    if constexpr (EPI != EPI_LINEAR) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = lrelu02(v[i]);
    }
    if (n0 >= p.n_valid) return;
    __half* dst;
    size_t pl;
    if constexpr (EPI == EPI_LRELU_S2D) {
      // planar-8 on the half-resolution s2d grid, channel = par*out_c + n
      const int H2 = p.H >> 1, W2 = p.W >> 1;
      const int G = (4 * p.out_c) >> 3, g = (((y & 1) * 2 + (c.x & 1)) * p.out_c + n0) >> 3;
      pl = (size_t)H2 * W2 * 8;
      dst = c.out + ((size_t)c.b * G + g) * pl + ((size_t)(y >> 1) * W2 + (c.x >> 1)) * 8;
    } else {
      pl = c.plane;
      dst = c.out + ((size_t)c.b * (p.out_c >> 3) + (n0 >> 3)) * pl + ((size_t)y * p.W + c.x) * 8;
    }
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = v[i];
    st8x(dst, a, p.out_bf);
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = v[8 + i];
    st8x(dst + pl, a, p.out_bf);
  } else if constexpr (EPI == EPI_RES) {
    const size_t o = ((size_t)c.b * (p.out_c >> 3) + (n0 >> 3)) * c.plane + ((size_t)y * p.W + c.x) * 8;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      // residual x[pixel][n0+8h .. +7] = plane (n0/8 + h) of the halo tile at (row+1, col+1)
      uint4 u;
      const uint32_t ra = c.res_smem + (uint32_t)(((n0 >> 3) + h) * p.plane_bytes + mi * TH * pitch * 16);
      asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(ra));
      float rr[8], a[8];
      unpack8x(u, p.in_bf, rr);
      const float4 b0 = *reinterpret_cast<const float4*>(s_beta + n0 + 8 * h);
      const float4 b1 = *reinterpret_cast<const float4*>(s_beta + n0 + 8 * h + 4);
      const float be[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = lrelu02(v[8 * h + i] * be[i] + rr[i]);
      st8x(c.out + o + h * c.plane, a, p.out_bf);
    }
  } else if constexpr (EPI == EPI_PS) {
    const int par = j >> 2, sub = j & 3;
    const int oy = 4 * y + 2 * (par >> 1) + (sub >> 1), ox = 4 * c.x + 2 * (par & 1) + (sub & 1);
    store16h(c.out + (((size_t)c.b * 4 * p.H + oy) * 4 * p.W + ox) * TMP_C, v);
  } else if constexpr (EPI == EPI_HEADT) {
    // Head.cnn3 (ConvTranspose 16 -> 4): writes the complete 8-channel SRC pixel [f0..f3 | r g b a] of the four
    // full-resolution pixels (2y+py, 2x+px); the image channels come from the s2d Head input (p.res, planar-8
    // [2B][2][H][W][8]: ch par*3+c = RGB of pixel par, ch 12+par = alpha), so the input-packing kernel never
    // touches SRC and every store is a full 16-byte pixel.
    const size_t hp = ((size_t)c.b * 2 * p.H + y) * p.W + c.x;
    float im[16];
    {
      float a[8];
      ld8(p.res + hp * 8, a);
#pragma unroll
      for (int i = 0; i < 8; ++i) im[i] = a[i];
      ld8(p.res + (hp + (size_t)p.H * p.W) * 8, a);
#pragma unroll
      for (int i = 0; i < 8; ++i) im[8 + i] = a[i];
    }
#pragma unroll
    for (int par = 0; par < 4; ++par) {
      const int oy = 2 * y + (par >> 1), ox = 2 * c.x + (par & 1);
      float f[8] = {v[4 * par], v[4 * par + 1], v[4 * par + 2], v[4 * par + 3],
                    im[par * 3], im[par * 3 + 1], im[par * 3 + 2], im[12 + par]};
      st8(c.out + (((size_t)c.b * 2 * p.H + oy) * 2 * p.W + ox) * SRC_C, f);
    }
  } else if constexpr (EPI == EPI_CONVT) {
    for (int i = 0; i < 16; ++i) {
      const int n = n0 + i;
      if (n >= p.n_valid) break;
      const int par = n / p.cpp, co = n - par * p.cpp;
      const int oy = 2 * y + (par >> 1), ox = 2 * c.x + (par & 1);
      c.out[(((size_t)c.b * 2 * p.H + oy) * 2 * p.W + ox) * p.out_c + co] = __float2half_rn(v[i]);
    }
  }
}

__device__ __forceinline__ void epi_chunk_rt(const HaloParams& p, const EpiCtx& c, const float* s_bias,
                                             const float* s_beta, int pitch, int mi, int j, float (&v)[16]) {
  switch (p.epi) {
    case EPI_LRELU: epi_chunk<EPI_LRELU>(p, c, s_bias, s_beta, pitch, mi, j, v); break;
    case EPI_LINEAR: epi_chunk<EPI_LINEAR>(p, c, s_bias, s_beta, pitch, mi, j, v); break;
    case EPI_LRELU_S2D: epi_chunk<EPI_LRELU_S2D>(p, c, s_bias, s_beta, pitch, mi, j, v); break;
    case EPI_RES: epi_chunk<EPI_RES>(p, c, s_bias, s_beta, pitch, mi, j, v); break;
    case EPI_PS: epi_chunk<EPI_PS>(p, c, s_bias, s_beta, pitch, mi, j, v); break;
    case EPI_HEADT: epi_chunk<EPI_HEADT>(p, c, s_bias, s_beta, pitch, mi, j, v); break;
    case EPI_CONVT: epi_chunk<EPI_CONVT>(p, c, s_bias, s_beta, pitch, mi, j, v); break;
    default: break;
  }
}

// KIND < 0: generic table-driven schedule (any layer that fits); otherwise the compile-time schedule <KIND,KSTEPS,NE,MT>
// EPI >= 0 additionally fixes the epilogue at compile time (fully unrolled over M-tiles and 16-column chunks).
// Epilogue warps: 4 (one per SM sub-partition); the column-heavy PixelShuffle epilogue (256 columns per M-tile) uses 8,
// warps w and w+4 sharing a TMEM lane quarter and splitting the output rows.
// Measured per layer shape (B200, 16 x 1080p pairs): the second warp per quarter pays for the ConvTranspose tails (EPI_PS /
// EPI_PS8 / Head.cnn3's EPI_HEADT: 0.407 -> 0.364 ms) and for the s2d conv0.1 layers (N >= 32: 0.078 -> 0.074, 0.047 -> 0.043
// ms); it costs the ResConv layers (0.56 -> 0.59 ms: their epilogue shares the smem tile with the MMAs) and the N = 16 Head
// layers (0.212 -> 0.229 ms).
template <int KIND, int NE, int EPI> struct EpiWarps {
  static constexpr int value =
      (EPI == EPI_PS || EPI == EPI_PS8 || EPI == EPI_HEADT || (EPI == EPI_LRELU && KIND == 1 && NE >= 32)) ? 8 : 4;
};

// STREAM: the layer's weights do not fit in smem next to the activation ring (c >= 96): an extra producer warp streams
// one weight tile per schedule entry through a ring of p.nb slots, re-using every tile for the MT stacked M-tiles.
template <int KIND, int KSTEPS, int NE, int MT, int EPI, bool STREAM = false>
__global__ void __launch_bounds__(64 + 32 * EpiWarps<KIND, NE, EPI>::value + (STREAM ? 32 : 0))
conv_halo_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, const HaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[MAX_STAGES];
  __shared__ __align__(8) uint64_t a_empty[MAX_STAGES];
  __shared__ __align__(8) uint64_t w_full;
  __shared__ __align__(8) uint64_t b_full[MAX_STAGES];
  __shared__ __align__(8) uint64_t b_empty[MAX_STAGES];
  __shared__ __align__(8) uint64_t t_full[2];
  __shared__ __align__(8) uint64_t t_empty[2];
  // tile ids travel from the TMA producer to the other roles through a 16-entry ring (entry seq & 15, barrier phase
  // (seq >> 4) & 1; -1 = end of a layer): the producer takes its tiles either round robin or -- tile_ctr != null -- from a
  // device-wide counter, so that a CTA that finishes early (cheap edge tiles, a faster SM) simply takes more tiles.  The
  // producer is never more than stages + 2 < 16 tiles ahead of the slowest reader (a_empty / t_empty back-pressure).
  __shared__ __align__(8) uint64_t tile_bar[16];
  __shared__ int tile_ring[16];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float s_bias[256];
  __shared__ __align__(16) float s_beta[128];
  // per-MMA operand tables, built once: the issuing thread then spends ~5 instructions per tcgen05.mma
  // (generic table-driven schedule only: the compile-time schedules do not spend static smem on them)
  constexpr int NTAB = KIND < 0 ? MAX_MMA : 1;
  __shared__ __align__(8) uint64_t tab_b[NTAB];  // complete B (weights) descriptor
  __shared__ uint32_t tab_a[NTAB];               // A start-address field relative to the stage base (16-byte units)
  __shared__ uint32_t tab_d[NTAB];               // accumulator column offset | (accumulate << 31)

  constexpr int NUM_EPI_WARPS = EpiWarps<KIND, NE, EPI>::value;
  constexpr int NUM_THREADS = 64 + 32 * NUM_EPI_WARPS;  // without the optional weight-streaming warp
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (p.dbg && threadIdx.x == 0) {
    long long g;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
    p.dbg[blockIdx.x * 16 + 8] = g;
  }
  const uint32_t w_smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_smem = w_smem + (STREAM ? (uint32_t)(p.nb * p.wtile_al) : (((uint32_t)p.w_bytes + 1023u) & ~1023u));

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_a) : "memory");
    if (p.nlayers > 1) asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_b) : "memory");
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&a_full[s]), 1);
      // ResConv epilogues read the residual (= centre tap of the input tile) from the smem stage itself, so the
      // stage is released by the MMA commit AND the four epilogue warps
      mbar_init(smem_u32(&a_empty[s]), p.epi == EPI_RES ? 1 + NUM_EPI_WARPS : 1);
    }
    mbar_init(smem_u32(&w_full), 1);
    for (int i = 0; i < 16; ++i) mbar_init(smem_u32(&tile_bar[i]), 1);
    if (STREAM) {
      for (int i = 0; i < p.nb; ++i) {
        mbar_init(smem_u32(&b_full[i]), 1);
        mbar_init(smem_u32(&b_empty[i]), 1);
      }
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&t_full[i]), 1);
      mbar_init(smem_u32(&t_empty[i]), NUM_EPI_WARPS);  // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp >= 2 && warp < 2 + NUM_EPI_WARPS) {
    for (int i = threadIdx.x - 64; i < p.ncols; i += NUM_THREADS - 64) {
      s_bias[i] = p.bias[i];
      if (p.beta && i < 128) s_beta[i] = p.beta[i];
    }
  }
  if (KIND < 0 && warp == 1) {
    const uint64_t b_desc0 = make_desc_noswz(0, (uint32_t)p.ne * 16u, 128u);
    for (int i = lane; i < p.nentries * p.ksteps; i += 32) {
      const int e = i / p.ksteps, kk = i - e * p.ksteps;
      tab_a[i] = (uint32_t)((2 * kk) * (p.plane_bytes >> 4) + (p.e_dy[e] + 1) * (TW + p.R) + (p.e_dx[e] + 1));
      tab_b[i] = b_desc0 + (uint64_t)((w_smem + (uint32_t)(e * p.wtile_bytes) + (uint32_t)(2 * kk * p.ne * 16)) >> 4);
      tab_d[i] = (uint32_t)p.e_col16[e] * 16u | ((p.e_first[e] && kk == 0) ? 0u : 0x80000000u);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (p.dbg && threadIdx.x == 0) {
    long long g;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
    p.dbg[blockIdx.x * 16 + 9] = g;
  }
  if (p.pdl_trig == 1) pdl_trigger();  // the next kernel's CTAs may start their prologue as SMs free up (they block in pdl_wait())
  const uint32_t tmem_base = tmem_slot;
  const int pitch = TW + p.R;  // halo row pitch in pixels
  const int n_mma = p.nentries * p.ksteps;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- producer: resident weights (per layer), then one halo box per tile
      const uint32_t wb = smem_u32(&w_full);
      if (!STREAM) {
        mbar_expect_tx(wb, (uint32_t)p.w_bytes);
        for (int off = 0; off < p.w_bytes; off += 32768) {
          const int n = min(32768, p.w_bytes - off);
          bulk_load(w_smem + off, reinterpret_cast<const uint8_t*>(p.w) + off, (uint32_t)n, wb);
        }
      }
      // weights are constants; the activations are the previous kernel's output.  Everything else in this CTA
      // (MMAs, epilogue loads from the smem tile, output stores) is downstream of these loads.
      pdl_wait();
      int stage = 0, seq = 0;
      uint32_t phase = 0;
      long long t_wait = 0;
      for (int l = 0; l < p.nlayers; ++l) {
        if (l > 0) {
          // every CTA has stored layer l-1 (that includes this CTA's epilogues, hence all of its MMAs: the weight
          // buffer is free again)
          grid_wait(p.sync, p.sync_base + (unsigned int)l * gridDim.x);
          if (!STREAM) {
            mbar_expect_tx(wb, (uint32_t)p.w_bytes);
            for (int off = 0; off < p.w_bytes; off += 32768) {
              const int n = min(32768, p.w_bytes - off);
              bulk_load(w_smem + off, reinterpret_cast<const uint8_t*>(p.w_l[l]) + off, (uint32_t)n, wb);
            }
          }
        }
        const CUtensorMap* tm = (l & 1) ? &tm_b : &tm_a;
        int tile = blockIdx.x;
        while (tile < p.ntiles) {
          const int tx = tile % p.tiles_x;
          const int r = tile / p.tiles_x;
          const int ty = r % p.tiles_y, b = r / p.tiles_y;
          // stacked M-tiles that start below the last image row (H not a multiple of 16 * MT) are skipped by the MMA
          // issuer and the epilogue: their count rides in the ring entry's top bits
          const int mt_eff = min(p.mt, (p.H - ty * p.mt * TH + TH - 1) / TH);
          mbar_wait_t(smem_u32(&a_empty[stage]), phase ^ 1, t_wait);
          tile_ring[seq & 15] = tile | (mt_eff << 28);
          mbar_arrive(smem_u32(&tile_bar[seq & 15]));
          ++seq;
          const uint32_t fb = smem_u32(&a_full[stage]);
          mbar_expect_tx(fb, (uint32_t)p.a_tx_bytes);
          tma_load_4d(a_smem + stage * p.a_stage_bytes, tm, fb, (tx * TW - 1) * 8, ty * TH * p.mt - 1, 0, b);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
          tile = p.tile_ctr ? (int)(gridDim.x + atomicAdd(p.tile_ctr, 1u)) : tile + (int)gridDim.x;
        }
        // end of this layer: an entry of -1 that takes a stage like a tile does (the MMA issuer learns about it through
        // a_full, the other roles through tile_bar), completed by a plain arrive instead of a TMA transaction
        mbar_wait_t(smem_u32(&a_empty[stage]), phase ^ 1, t_wait);
        tile_ring[seq & 15] = -1;
        mbar_arrive(smem_u32(&tile_bar[seq & 15]));
        ++seq;
        mbar_arrive(smem_u32(&a_full[stage]));
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      if (p.dbg) p.dbg[blockIdx.x * 16 + 0] = t_wait;
    }
  } else if (STREAM && warp == 2 + NUM_EPI_WARPS) {
    if (lane == 0) {
      // ---------------- weight-tile producer: entry e of every tile -> ring slot (runs ahead across layers)
      int bs = 0, seq = 0;
      uint32_t bph = 0;
      for (int l = 0; l < p.nlayers; ++l) {
        const uint8_t* wl = reinterpret_cast<const uint8_t*>(p.nlayers > 1 ? p.w_l[l] : p.w);
        while (ring_next_tile(tile_ring, tile_bar, seq++) >= 0) {
          for (int e = 0; e < p.nentries; ++e) {
            mbar_wait(smem_u32(&b_empty[bs]), bph ^ 1);
            const uint32_t fb = smem_u32(&b_full[bs]);
            mbar_expect_tx(fb, (uint32_t)p.wtile_bytes);
            const uint8_t* src = wl + (size_t)e * p.wtile_bytes;
            for (int off = 0; off < p.wtile_bytes; off += 32768)
              bulk_load(w_smem + bs * p.wtile_al + off, src + off, (uint32_t)min(32768, p.wtile_bytes - off), fb);
            if (++bs == p.nb) { bs = 0; bph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- MMA issuer
      // instruction descriptor: D = f32, A / B format 0 = f16 or 1 = bf16 (bits 7-9 / 10-12), N >> 3, M >> 4
      const uint32_t idesc = (1u << 4) | (p.in_bf ? ((1u << 7) | (1u << 10)) : 0u) | ((uint32_t)(p.ne >> 3) << 17) |
                             ((uint32_t)(128 >> 4) << 24);
      const uint64_t a_desc0 = make_desc_noswz(0, (uint32_t)p.plane_bytes, (uint32_t)pitch * 16u);
      const uint32_t mt_step = (uint32_t)(TH * pitch);  // 16 halo rows, in 16-byte units (one pixel = 16 B per plane)
      long long t_w = 0, t_te = 0, t_af = 0;
      const long long t_begin = clock64();
      int stage = 0, it = 0, bs = 0, seq = 0;
      uint32_t phase = 0, bph = 0;
      for (int l = 0; l < p.nlayers; ++l) {
      if (!STREAM) mbar_wait_t(smem_u32(&w_full), (uint32_t)(l & 1), t_w);
      for (;; ++it) {
        // the stage's a_full (acquire) also publishes the ring entry the producer wrote before arming it: no second wait
        // on this thread, whose per-tile instruction count bounds the thin layers (18 MMAs per tile in block 4's conv0.0)
        mbar_wait_t(smem_u32(&a_full[stage]), phase, t_af);
        const int entry = *reinterpret_cast<const volatile int*>(&tile_ring[seq & 15]);
        ++seq;
        if (entry < 0) {  // end of the layer: hand the stage back like a tile would
          umma_commit(smem_u32(&a_empty[stage]));
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
          break;
        }
        const int mt_eff = (KIND < 0 || MT > 1) ? (entry >> 28) : 1;
        const int buf = it & 1;
        mbar_wait_t(smem_u32(&t_empty[buf]), (((uint32_t)it >> 1) & 1u) ^ 1u, t_te);
        tc_fence_after();
        const uint32_t sa16 = (a_smem + stage * p.a_stage_bytes) >> 4;
        if constexpr (KIND >= 0 && STREAM) {
          const uint64_t b_desc0 = make_desc_noswz(0, (uint32_t)NE * 16u, 128u);
          const uint32_t acc0 = tmem_base + (uint32_t)(buf * MT * p.ncols);
          auto one = [&](auto ec) {
            constexpr int E = decltype(ec)::value;
            mbar_wait_t(smem_u32(&b_full[bs]), bph, t_w);
            tc_fence_after();
            issue_entry<KIND, KSTEPS, NE, MT, E>(acc0, (uint32_t)p.ncols, a_desc0 + (uint64_t)sa16, mt_step,
                                                 b_desc0 + (uint64_t)((w_smem + (uint32_t)(bs * p.wtile_al)) >> 4), idesc, mt_eff);
            umma_commit(smem_u32(&b_empty[bs]));
            if (++bs == p.nb) { bs = 0; bph ^= 1; }
          };
          static_for<Sched<KIND>::NENT>(one);
        } else if constexpr (KIND >= 0) {
          const uint64_t b_base = make_desc_noswz(w_smem, (uint32_t)NE * 16u, 128u);
#pragma unroll
          for (int m = 0; m < MT; ++m)
            if (m < mt_eff)
              issue_tile<KIND, KSTEPS, NE, MT>(tmem_base + (uint32_t)((buf * MT + m) * p.ncols),
                                               a_desc0 + (uint64_t)(sa16 + (uint32_t)m * mt_step), b_base, idesc);
        } else {
          for (int m = 0; m < mt_eff; ++m) {
            const uint64_t a_base = a_desc0 + (uint64_t)(sa16 + (uint32_t)m * mt_step);
            const uint32_t acc = tmem_base + (uint32_t)((buf * p.mt + m) * p.ncols);
#pragma unroll 4
            for (int i = 0; i < n_mma; ++i) {
              const uint32_t d = tab_d[i];
              umma_f16(acc + (d & 0xFFFFu), a_base + tab_a[i], tab_b[i], idesc, d >> 31);
            }
          }
        }
        umma_commit(smem_u32(&a_empty[stage]));
        umma_commit(smem_u32(&t_full[buf]));
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      }
      if (p.pdl_trig == 2) pdl_trigger();  // only the last epilogue is left: let the next kernel's CTAs set up meanwhile
      if (p.dbg) {
        long long g;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
        p.dbg[blockIdx.x * 16 + 10] = g;
        p.dbg[blockIdx.x * 16 + 1] = t_w;
        p.dbg[blockIdx.x * 16 + 2] = t_te;
        p.dbg[blockIdx.x * 16 + 3] = t_af;
        p.dbg[blockIdx.x * 16 + 4] = clock64() - t_begin;
      }
    }
  } else {
    // ---------------- epilogue warps: TMEM lanes 32*(warp%4)..+31; warps w and w+4 share a lane quarter and split
    // the 16-column chunks (even / odd) between them
    const int q = warp & 3;
    const int ehalf = (warp - 2) >> 2;
    constexpr int ESPLIT = NUM_EPI_WARPS / 4;  // 1 or 2 warps per lane quarter
    const int m = q * 32 + lane;
    const int ty_ = m >> 3, tx_ = m & 7;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    const int nchunk16 = p.ncols >> 4;
    const size_t plane = (size_t)p.H * p.W * 8;  // elements per planar-8 channel group
    int it = 0, stage = 0, seq = 0;
    uint32_t phase = 0;
    long long t_tf = 0;
    const long long t_begin = clock64();
    for (int l = 0; l < p.nlayers; ++l) {
    __half* const out_l = p.nlayers > 1 ? p.out_l[l & 1] : p.out;
    for (;; ++it) {
      const int entry = ring_next_tile(tile_ring, tile_bar, seq++);
      if (entry < 0) {  // end of the layer: the entry took a stage
        if ((EPI >= 0 ? EPI : p.epi) == EPI_RES) {
          mbar_wait(smem_u32(&a_full[stage]), phase);
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&a_empty[stage]));
        }
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
        break;
      }
      const int tile = entry & 0x0FFFFFFF;
      const int tx = tile % p.tiles_x;
      const int r = tile / p.tiles_x;
      const int ty = r % p.tiles_y, b = r / p.tiles_y;
      const int buf = it & 1;
      EpiCtx c;
      c.b = b; c.x = tx * TW + tx_; c.plane = plane; c.out = out_l;
      c.res_smem = a_smem + stage * p.a_stage_bytes + (uint32_t)(((ty_ + 1) * pitch + tx_ + 1) * 16);
      if ((EPI >= 0 ? EPI : p.epi) == EPI_RES) mbar_wait(smem_u32(&a_full[stage]), phase);  // acquire the TMA-written tile
      mbar_wait_t(smem_u32(&t_full[buf]), ((uint32_t)it >> 1) & 1u, t_tf);
      tc_fence_after();
      if constexpr (KIND == 2 && EPI == EPI_PS) {
        // ConvTranspose + PixelShuffle: every input pixel yields a 4x4 patch of 16-channel tmp pixels (32 B each).
        // Stage one output row (py, dy') of the warp's 4 x 8 input pixels in smem, then write 512 contiguous bytes
        // per store instruction (a thread-per-pixel store would touch 32 different 128-byte lines per instruction).
        __shared__ uint4 ps_stage[NUM_EPI_WARPS][4][64];
        uint4(*stg)[64] = ps_stage[warp - 2];
        const int tyl = lane >> 3;
        const int swz_w = (tx_ + 2 * tyl) & 7;
#pragma unroll 1
        for (int mi = 0; mi < MT; ++mi) {
          const int yw = (ty * MT + mi) * TH + q * 4;  // first input row of this warp
#pragma unroll 1
          for (int r = (4 / ESPLIT) * ehalf; r < (4 / ESPLIT) * (ehalf + 1); ++r) {
            const int py = r >> 1, dyp = r & 1;
#pragma unroll
            for (int o = 0; o < 4; ++o) {
              const int j = (py * 2 + (o >> 1)) * 4 + dyp * 2 + (o & 1);
              float v[16];
              tmem_ld16(trow + (uint32_t)((buf * MT + mi) * 256 + j * 16), v);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 bb = *reinterpret_cast<const float4*>(s_bias + j * 16 + 4 * i);
                v[4 * i] += bb.x; v[4 * i + 1] += bb.y; v[4 * i + 2] += bb.z; v[4 * i + 3] += bb.w;
              }
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                uint4 u;
                __half2* hh = reinterpret_cast<__half2*>(&u);
#pragma unroll
                for (int i = 0; i < 4; ++i) hh[i] = __floats2half2_rn(v[8 * h + 2 * i], v[8 * h + 2 * i + 1]);
                stg[tyl][((tx_ * 4 + o) * 2 + h) ^ swz_w] = u;
              }
            }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int idx = k * 32 + lane;
              const int row = idx >> 6, col = idx & 63, txc = col >> 3;
              const uint4 u = stg[row][col ^ ((txc + 2 * row) & 7)];
              const int y = yw + row, x = tx * TW + txc;
              if (y < p.H && x < p.W) {
                const int oy = 4 * y + 2 * py + dyp;
                __half* dst = out_l + (((size_t)b * 4 * p.H + oy) * 4 * p.W + (size_t)4 * tx * TW) * TMP_C + col * 8;
                *reinterpret_cast<uint4*>(dst) = u;
              }
            }
            __syncwarp();
          }
        }
      } else if constexpr (KIND == 2 && EPI == EPI_PS8) {
        // Last block of the network: only flow (4) + mask (1) are consumed downstream (IFNet_HDv3.py:152,158; the feat
        // channels of block 4 are dropped by the reference too), so the tmp pixel is 8 channels = 16 B and the
        // accumulator has 4 parities x (4 sub-pixels x 8) = 128 columns.  One staged output row (py, dy') of the warp's
        // 4 x 8 input pixels = 4 x 512 contiguous bytes.
        __shared__ uint4 ps8_stage[NUM_EPI_WARPS][4][32];
        uint4(*stg)[32] = ps8_stage[warp - 2];
        const int tyl = lane >> 3;
        const int swz_w = tx_ >> 1;
#pragma unroll 1
        for (int mi = 0; mi < MT; ++mi) {
          const int yw = (ty * MT + mi) * TH + q * 4;  // first input row of this warp
#pragma unroll 1
          for (int r = (4 / ESPLIT) * ehalf; r < (4 / ESPLIT) * (ehalf + 1); ++r) {
            const int py = r >> 1, dyp = r & 1;
#pragma unroll
            for (int o = 0; o < 2; ++o) {  // o = px: columns (py*2+px)*32 + dy'*16 + dx'*8 + k
              const int j = (py * 2 + o) * 2 + dyp;
              float v[16];
              tmem_ld16(trow + (uint32_t)((buf * MT + mi) * 128 + j * 16), v);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 bb = *reinterpret_cast<const float4*>(s_bias + j * 16 + 4 * i);
                v[4 * i] += bb.x; v[4 * i + 1] += bb.y; v[4 * i + 2] += bb.z; v[4 * i + 3] += bb.w;
              }
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                uint4 u;
                __half2* hh = reinterpret_cast<__half2*>(&u);
#pragma unroll
                for (int i = 0; i < 4; ++i) hh[i] = __floats2half2_rn(v[8 * h + 2 * i], v[8 * h + 2 * i + 1]);
                stg[tyl][(tx_ * 4 + o * 2 + h) ^ swz_w] = u;
              }
            }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int idx = k * 32 + lane;
              const int row = idx >> 5, col = idx & 31, txc = col >> 2;
              const uint4 u = stg[row][col ^ (txc >> 1)];
              const int y = yw + row, x = tx * TW + txc;
              if (y < p.H && x < p.W) {
                const int oy = 4 * y + 2 * py + dyp;
                __half* dst = out_l + (((size_t)b * 4 * p.H + oy) * 4 * p.W + (size_t)4 * tx * TW) * 8 + col * 8;
                *reinterpret_cast<uint4*>(dst) = u;
              }
            }
            __syncwarp();
          }
        }
      } else if constexpr (KIND >= 0 && EPI >= 0) {
        constexpr int NCOLS = KIND == 2 ? 256 : NE;
        c.y0 = ty * MT * TH + ty_;
#pragma unroll
        for (int mi = 0; mi < MT; ++mi) {
          if (mi > 0 && (ty * MT + mi) * TH >= p.H) continue;   // M-tile below the image: not computed (see the MMA issuer)
#ifndef RIFE_EPI_NO_BATCH
          if constexpr (ESPLIT == 1 && NCOLS <= 64) {
            // all of the M-tile's accumulator chunks are requested before the one wait (TMEM latency paid once per M-tile)
            float v[NCOLS / 16][16];
#pragma unroll
            for (int j = 0; j < NCOLS / 16; ++j) tmem_ld16_nowait(trow + (uint32_t)((buf * MT + mi) * NCOLS + j * 16), v[j]);
            tmem_ld_wait_all();
#pragma unroll
            for (int j = 0; j < NCOLS / 16; ++j) tmem_regs_ready(v[j]);
#pragma unroll
            for (int j = 0; j < NCOLS / 16; ++j) {
              epi_chunk<EPI>(p, c, s_bias, s_beta, pitch, mi, j, v[j]);
              __syncwarp();
            }
            continue;
          }
#endif
#pragma unroll
          for (int j = 0; j < NCOLS / 16; ++j) {
            if (ESPLIT == 2 && ((mi * (NCOLS / 16) + j) & 1) != ehalf) continue;
            float v[16];
            tmem_ld16(trow + (uint32_t)((buf * MT + mi) * NCOLS + j * 16), v);
            epi_chunk<EPI>(p, c, s_bias, s_beta, pitch, mi, j, v);
            __syncwarp();
          }
        }
      } else {
        c.y0 = ty * p.mt * TH + ty_;
        for (int jm = ehalf; jm < nchunk16 * p.mt; jm += ESPLIT) {
          const int mi = jm / nchunk16, j = jm - mi * nchunk16;
          if (mi > 0 && (ty * p.mt + mi) * TH >= p.H) continue;
          float v[16];
          __syncwarp();
          tmem_ld16(trow + (uint32_t)((buf * p.mt + mi) * p.ncols + j * 16), v);
          epi_chunk_rt(p, c, s_bias, s_beta, pitch, mi, j, v);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(smem_u32(&t_empty[buf]));
        if ((EPI >= 0 ? EPI : p.epi) == EPI_RES) mbar_arrive(smem_u32(&a_empty[stage]));
      }
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
    if (l + 1 < p.nlayers) {
      // layer boundary: this layer's stores published grid-wide, next layer's bias / beta loaded once every epilogue
      // warp is done with the current ones
      __threadfence();
      asm volatile("fence.proxy.async;" ::: "memory");
      asm volatile("bar.sync 1, %0;" ::"r"(32 * NUM_EPI_WARPS) : "memory");
      if (threadIdx.x == 64) grid_arrive(p.sync);
      for (int i = threadIdx.x - 64; i < p.ncols; i += NUM_THREADS - 64) {
        s_bias[i] = p.bias_l[l + 1][i];
        if (i < 128) s_beta[i] = p.beta_l[l + 1][i];
      }
      asm volatile("bar.sync 1, %0;" ::"r"(32 * NUM_EPI_WARPS) : "memory");
    }
    }
    if (p.dbg && warp == 2 && lane == 0) {  // timers of the first epilogue warp
      p.dbg[blockIdx.x * 16 + 5] = t_tf;
      p.dbg[blockIdx.x * 16 + 6] = clock64() - t_begin;
      p.dbg[blockIdx.x * 16 + 7] = it;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                 : "memory");
  }
  if (p.dbg && threadIdx.x == 0) {
    long long g;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
    p.dbg[blockIdx.x * 16 + 11] = g;
  }
  if (p.nlayers > 1 && threadIdx.x == 0) {
    // chain: every CTA that gets here has passed all of its barrier waits; the last one to finish re-arms the counters
    __threadfence();
    const unsigned int d = atomicAdd(p.sync + 1, 1u);
    if (d == gridDim.x - 1) {
      p.sync[0] = 0u;
      p.sync[1] = 0u;
      __threadfence();
    }
  }
}

// ---------------------------------------------------------------------------------------- layout converters (tests)
__device__ __forceinline__ __half cvt16_dev(float x, bool bf) {
  if (!bf) return __float2half_rn(x);
  const __nv_bfloat16 b = __float2bfloat16_rn(x);
  return *reinterpret_cast<const __half*>(&b);
}
__device__ __forceinline__ float cvt16_to_f(__half h, bool bf) {
  if (!bf) return __half2float(h);
  return __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(&h));
}
__global__ void nhwc_f32_to_planar8_kernel(const float* __restrict__ x, __half* __restrict__ y, int B, int H, int W, int C,
                                           bool bf) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * H * W * C) return;
  const int c = (int)(i % C);
  const size_t pix = i / C;
  const size_t hw = (size_t)H * W;
  const size_t b = pix / hw, p = pix % hw;
  y[((b * (C / 8) + c / 8) * hw + p) * 8 + (c & 7)] = cvt16_dev(x[i], bf);
}
__global__ void planar8_to_nhwc_f32_kernel(const __half* __restrict__ x, float* __restrict__ y, int B, int H, int W, int C,
                                           bool bf) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * H * W * C) return;
  const int c = (int)(i % C);
  const size_t pix = i / C;
  const size_t hw = (size_t)H * W;
  const size_t b = pix / hw, p = pix % hw;
  y[i] = cvt16_to_f(x[((b * (C / 8) + c / 8) * hw + p) * 8 + (c & 7)], bf);
}
__global__ void f16_to_f32_k(const __half* __restrict__ x, float* __restrict__ y, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = __half2float(x[i]);
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

}  // namespace

// 0 = does not fit, 1 = weights resident in smem, 2 = weights streamed per schedule entry (templated shapes only:
// the c = 96 / 128 ResConv and lastconv layers of IFBlock 2 / 1)
int halo_mode(const HostLayer& h, int ps_layout) {
  if (h.cin % 16) return 0;
  const int R = h.kind == CONV_3x3 ? 2 : 1;
  const int a_stage = ((h.cin / 8) * (TH + R) * (TW + R) * 16 + 1023) / 1024 * 1024;
  const int ne = ps_layout == 2 ? 32 : (ps_layout ? 64 : h.npad);
  const int w_bytes = (ps_layout ? 16 : h.taps) * (h.cin / 8) * ne * 16;
  const int budget = SMEM_BUDGET - (ps_layout ? 32 * 1024 : 0);
  if (h.cin <= 128 && (ps_layout || h.npad <= 128) && w_bytes + 2 * a_stage + 2048 <= budget) return 1;
  const bool templated = h.kind == CONV_3x3 && (h.cin == 96 || h.cin == 128) &&
                         (ps_layout ? h.convt_co == 52 : (h.npad == h.cin && h.convt_co == 0));
  const int wtile = ((h.cin / 8) * ne * 16 + 1023) / 1024 * 1024;
  if (templated && 2 * a_stage + 2 * wtile + 2048 <= budget) return 2;
  return 0;
}
bool halo_fits(const HostLayer& h, int ps_layout) { return halo_mode(h, ps_layout) != 0; }

int halo_pack_weights(const HostLayer& h, int ps_layout, std::vector<__half>* packed, std::vector<float>* bias,
                      std::vector<float>* beta, HaloLayer* L, bool bf16) {
  if (!halo_fits(h, ps_layout)) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "halo_pack_weights: layer does not fit the halo kernel");
  const int kg = h.cin / 8;
  *L = HaloLayer();
  L->kind = h.kind; L->cin = h.cin;
  beta->clear();
  if (ps_layout) {
    if (h.convt_co != 52 || h.kind != CONV_3x3) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "halo_pack_weights: PS layout is the 52-channel lastconv");
    // ps_layout 2: the narrow variant for the network's last block -- flow (4) + mask (1) only, 8 channels per tmp pixel
    const int pc = ps_layout == 2 ? 8 : 16, nk = ps_layout == 2 ? 5 : 13, ne = 4 * pc;
    L->nentries = 16; L->ne = ne; L->ncols = 4 * ne; L->n_valid = 4 * ne; L->ps_cols = ps_layout == 2 ? 2 : 1;
    packed->assign((size_t)16 * kg * ne * 8, __float2half_rn(0.f));
    bias->assign(4 * ne, 0.f);
    int e = 0;
    for (int par = 0; par < 4; ++par) {
      const int py = par >> 1, px = par & 1;
      for (int sub = 0; sub < 4; ++sub)
        for (int k = 0; k < nk; ++k) (*bias)[par * ne + sub * pc + k] = h.bias[par * 52 + 4 * k + sub];
      for (int dyi = 0; dyi < 2; ++dyi)
        for (int dxi = 0; dxi < 2; ++dxi, ++e) {
          const int dy = py ? dyi : dyi - 1, dx = px ? dxi : dxi - 1;  // taps that feed this output parity
          const int t = (dy + 1) * 3 + (dx + 1);
          L->e_dy[e] = (signed char)dy; L->e_dx[e] = (signed char)dx;
          L->e_col16[e] = (unsigned char)(par * ne / 16);
          L->e_first[e] = (dyi == 0 && dxi == 0) ? 1 : 0;
          for (int g = 0; g < kg; ++g)
            for (int sub = 0; sub < 4; ++sub)
              for (int k = 0; k < nk; ++k) {
                const int n = sub * pc + k, on = par * 52 + 4 * k + sub;
                for (int c = 0; c < 8; ++c)
                  (*packed)[(((size_t)e * kg + g) * ne + n) * 8 + c] =
                      host_cvt16(h.w[((size_t)t * h.cin + g * 8 + c) * h.npad + on], bf16);
              }
        }
    }
  } else {
    L->nentries = h.taps; L->ne = h.npad; L->ncols = h.npad; L->n_valid = h.n;
    packed->assign((size_t)h.taps * kg * h.npad * 8, __float2half_rn(0.f));
    bias->assign(h.bias.begin(), h.bias.end());
    if (!h.beta.empty()) beta->assign(h.beta.begin(), h.beta.end());
    const int side = h.kind == CONV_3x3 ? 3 : 2;
    for (int t = 0; t < h.taps; ++t) {
      L->e_dy[t] = (signed char)(t / side - 1); L->e_dx[t] = (signed char)(t % side - 1);
      L->e_col16[t] = 0; L->e_first[t] = t == 0;
      for (int g = 0; g < kg; ++g)
        for (int n = 0; n < h.npad; ++n)
          for (int c = 0; c < 8; ++c)
            (*packed)[(((size_t)t * kg + g) * h.npad + n) * 8 + c] =
                host_cvt16(h.w[((size_t)t * h.cin + g * 8 + c) * h.npad + n], bf16);
    }
  }
  L->w_bytes = (int)(packed->size() * sizeof(__half));
  L->stream = halo_mode(h, ps_layout) == 2 ? 1 : 0;
  return RIFE_OK;
}

static int set_smem_attr(const void* fn) {
  // once per (function, device): opt in to > 48 KB dynamic shared memory
  static std::vector<std::pair<const void*, int>> seen;
  static std::mutex mu;   // one host thread per GPU may launch concurrently
  std::lock_guard<std::mutex> lock(mu);
  int dev = 0;
  cudaGetDevice(&dev);
  for (auto& s : seen)
    if (s.first == fn && s.second == dev) return RIFE_OK;
  cudaFuncAttributes fa;
  RIFE_CUDA_TRY(cudaFuncGetAttributes(&fa, fn));
  const int max_dyn = 227 * 1024 - (int)fa.sharedSizeBytes;  // 227 KB opt-in limit per CTA, static part included
  RIFE_CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn));
  seen.emplace_back(fn, dev);
  return RIFE_OK;
}

namespace {
struct HaloChain {  // nlayers ResConv layers of one shape: layer l reads (l even ? io.in : io.out), writes the other one
  const HaloLayer* const* layers;
  int n;
  unsigned int* sync;
  unsigned int* sync_host;  // host mirror of the counter's value after everything enqueued so far
};

// cooperative launch (all CTAs co-resident: the chain's grid barrier spins) -- or the usual PDL launch
template <typename... KArgs>
int launch_hk(void (*kernel)(KArgs...), int grid, int threads, int smem, cudaStream_t st, const CUtensorMap& tm_a,
              const CUtensorMap& tm_b, const HaloParams& p, bool coop) {
  if (!coop) {
    RIFE_CUDA_TRY(launch_pdl(kernel, dim3(grid), dim3(threads), smem, st, tm_a, tm_b, p));
    return RIFE_OK;
  }
  int nb = 0;
  RIFE_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, threads, smem));
  if (grid > nb * num_sms()) return 1;  // not co-resident: the caller launches layer by layer
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  static int with_pdl = pdl_enabled() ? 1 : 0;   // cooperative + programmatic serialization; dropped if the driver refuses
  if (with_pdl) {
    cfg.numAttrs = 2;
    if (cudaLaunchKernelEx(&cfg, kernel, tm_a, tm_b, p) == cudaSuccess) return RIFE_OK;
    cudaGetLastError();
    with_pdl = 0;
  }
  cfg.numAttrs = 1;
  RIFE_CUDA_TRY(cudaLaunchKernelEx(&cfg, kernel, tm_a, tm_b, p));
  return RIFE_OK;
}

int launch_conv_halo_impl(const HaloLayer& L, const ConvIO& io, cudaStream_t st, const HaloChain* chain);
}  // namespace

int launch_conv_halo(const HaloLayer& L, const ConvIO& io, cudaStream_t st) { return launch_conv_halo_impl(L, io, st, nullptr); }

int launch_conv_halo_chain(const HaloLayer* const* layers, int n, const ConvIO& io, unsigned int* sync,
                           unsigned int* sync_host, cudaStream_t st) {
  if (n < 2 || n > 8 || !sync || !sync_host) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "conv_halo chain: 2..8 layers and a sync counter");
  if (io.epi != EPI_RES) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "conv_halo chain: ResConv layers only");
  for (int i = 1; i < n; ++i)
    if (layers[i]->cin != layers[0]->cin || layers[i]->ne != layers[0]->ne || layers[i]->nentries != layers[0]->nentries ||
        layers[i]->w_bytes != layers[0]->w_bytes || layers[i]->stream != layers[0]->stream || !layers[i]->beta)
      RIFE_FAIL(RIFE_ERR_INVALID_ARG, "conv_halo chain: layers must share one shape");
  HaloChain ch{layers, n, sync, sync_host};
  return launch_conv_halo_impl(*layers[0], io, st, &ch);
}

namespace {
int launch_conv_halo_impl(const HaloLayer& L, const ConvIO& io, cudaStream_t st, const HaloChain* chain) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) RIFE_FAIL(RIFE_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  if (io.epi == EPI_PS && L.ps_cols != 1) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "conv_halo: EPI_PS needs the PixelShuffle layout");
  if (io.epi == EPI_PS8 && L.ps_cols != 2) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "conv_halo: EPI_PS8 needs the narrow PixelShuffle layout");
  if (io.epi == EPI_RES && (!L.beta || !io.res)) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "conv_halo: EPI_RES needs beta and res");
  if (io.epi == EPI_RES && L.cin != io.out_c) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "conv_halo: EPI_RES is the c -> c ResConv (residual = input)");
  if (io.epi == EPI_HEADT && !io.res) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "conv_halo: EPI_HEADT needs the s2d Head input as io.res");
  if (io.epi == EPI_HEADT && L.ncols != 16) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "conv_halo: EPI_HEADT expects N = 16");
  HaloParams p{};
  p.B = io.B; p.H = io.H; p.W = io.W;
  p.kgroups = L.cin / 8; p.ksteps = L.cin / 16;
  p.R = L.kind == CONV_3x3 ? 2 : 1;
  if (L.nentries * p.ksteps > MAX_MMA) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "conv_halo: MMA schedule too long");
  // stack up to 4 M-tiles per pipeline step for thin layers: amortises the per-step barrier chain and the halo
  const int wtile_al = (p.kgroups * L.ne * 16 + 1023) / 1024 * 1024;
  {
    // resident: all weights + >= 3 activation stages; streamed: 2 activation stages + >= 3 weight tiles
    const int w_al0 = L.stream ? 3 * wtile_al : (L.w_bytes + 1023) / 1024 * 1024;
    const int nst = L.stream ? 2 : 3;
    static const int mt_max = []() { const char* e = getenv("RIFE_HK_MT"); return e ? atoi(e) : 2; }();
    int mt = mt_max < 1 ? 1 : (mt_max > 4 ? 4 : mt_max);
    if (L.ps_cols) mt = 1;
    while (mt > 1 && (2 * mt * L.ncols > 512 || io.H < TH * mt ||
                      w_al0 + nst * (p.kgroups * (TH * mt + p.R) * (TW + p.R) * 16 + 1024) + 2048 > SMEM_BUDGET))
      mt >>= 1;
    p.mt = mt;
  }
  p.tiles_x = (io.W + TW - 1) / TW;
  p.tiles_y = (io.H + TH * p.mt - 1) / (TH * p.mt);
  p.ntiles = p.tiles_x * p.tiles_y * io.B;
  p.plane_bytes = (TH * p.mt + p.R) * (TW + p.R) * 16;
  p.a_tx_bytes = p.kgroups * p.plane_bytes;
  p.a_stage_bytes = (p.a_tx_bytes + 1023) / 1024 * 1024;
  p.nentries = L.nentries; p.ne = L.ne; p.ncols = L.ncols; p.n_valid = L.n_valid;
  p.w_bytes = L.w_bytes;
  p.wtile_bytes = p.kgroups * L.ne * 16;
  p.tmem_cols = 32;
  while (p.tmem_cols < 2 * p.mt * L.ncols) p.tmem_cols <<= 1;
  // the PixelShuffle epilogues stage rows in static smem (8 warps x 4 KB / 2 KB)
  const int static_extra = io.epi == EPI_PS ? 32 * 1024 : (io.epi == EPI_PS8 ? 16 * 1024 : 0);
  const int budget = SMEM_BUDGET - static_extra;
  int w_al = (L.w_bytes + 1023) / 1024 * 1024;
  int stages;
  p.nb = 0; p.wtile_al = wtile_al;
  if (L.stream) {
    stages = 2;
    int nb = (budget - 2048 - stages * p.a_stage_bytes) / wtile_al;
    if (nb > MAX_STAGES) nb = MAX_STAGES;
    if (nb > L.nentries) nb = L.nentries;
    if (nb < 2) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "conv_halo: streamed layer does not fit in shared memory");
    // spare smem goes to a third activation stage
    if (budget - 2048 - nb * wtile_al >= 3 * p.a_stage_bytes) stages = 3;
    p.nb = nb;
    w_al = nb * wtile_al;
  } else {
    stages = (budget - w_al - 2048) / p.a_stage_bytes;
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (stages < 2) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "conv_halo: layer does not fit in shared memory");
  }
  int smem = w_al + stages * p.a_stage_bytes + 1024;
  // co-resident CTAs per SM hide the per-step latency chain of thin layers when smem and TMEM allow
  int per_sm = 1;
  static const int persm_max = []() { const char* e = getenv("RIFE_HK_PERSM"); return e ? atoi(e) : 4; }();
  for (int k = L.stream ? 1 : persm_max; k >= 2; --k) {
    if (k * p.tmem_cols > 512) continue;
    int s2 = stages;
    const int lim = static_extra ? 212 * 1024 : 224 * 1024;  // static smem is not part of `smem`: leave room for it
    while (s2 > 3 && k * (w_al + s2 * p.a_stage_bytes + 2048 + static_extra) > lim) --s2;
    if (k * (w_al + s2 * p.a_stage_bytes + 2048 + static_extra) <= lim) {
      per_sm = k;
      stages = s2;
      smem = w_al + stages * p.a_stage_bytes + 1024;
      break;
    }
  }
  p.stages = stages;
  p.epi = io.epi; p.out_c = io.out_c; p.cpp = io.co_per_par;
  p.in_bf = io.in_bf16 ? 1 : 0; p.out_bf = io.out_bf16 ? 1 : 0;
  p.w = L.w; p.bias = L.bias; p.beta = L.beta;
  p.res = (const __half*)io.res; p.out = (__half*)io.out;
  p.tile_ctr = chain ? nullptr : io.tile_ctr;   // (a chain would need one counter per layer)
  for (int e = 0; e < L.nentries; ++e) {
    p.e_dy[e] = L.e_dy[e]; p.e_dx[e] = L.e_dx[e]; p.e_col16[e] = L.e_col16[e]; p.e_first[e] = L.e_first[e];
  }
  CUtensorMap tm_a;
  {
    cuuint64_t dims[4] = {(cuuint64_t)io.W * 8, (cuuint64_t)io.H, (cuuint64_t)p.kgroups, (cuuint64_t)io.B};
    cuuint64_t strides[3] = {(cuuint64_t)io.W * 16, (cuuint64_t)io.H * io.W * 16, (cuuint64_t)p.kgroups * io.H * io.W * 16};
    cuuint32_t box[4] = {(cuuint32_t)((TW + p.R) * 8), (cuuint32_t)(TH * p.mt + p.R), (cuuint32_t)p.kgroups, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&tm_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(io.in), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) RIFE_FAIL(RIFE_ERR_CUDA, "cuTensorMapEncodeTiled(halo) failed: " + std::to_string((int)r));
  }
  CUtensorMap tm_b = tm_a;
  p.nlayers = 1;
  if (chain) {
    cuuint64_t dims[4] = {(cuuint64_t)io.W * 8, (cuuint64_t)io.H, (cuuint64_t)p.kgroups, (cuuint64_t)io.B};
    cuuint64_t strides[3] = {(cuuint64_t)io.W * 16, (cuuint64_t)io.H * io.W * 16, (cuuint64_t)p.kgroups * io.H * io.W * 16};
    cuuint32_t box[4] = {(cuuint32_t)((TW + p.R) * 8), (cuuint32_t)(TH * p.mt + p.R), (cuuint32_t)p.kgroups, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&tm_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, io.out, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) RIFE_FAIL(RIFE_ERR_CUDA, "cuTensorMapEncodeTiled(halo chain) failed: " + std::to_string((int)r));
    p.nlayers = chain->n;
    for (int l = 0; l < chain->n; ++l) {
      p.w_l[l] = chain->layers[l]->w; p.bias_l[l] = chain->layers[l]->bias; p.beta_l[l] = chain->layers[l]->beta;
    }
    p.out_l[0] = (__half*)io.out;
    p.out_l[1] = (__half*)const_cast<void*>(io.in);
    p.sync = chain->sync;
    p.sync_base = 0;
  }
  const bool coop = chain != nullptr;
  int grid = num_sms() * per_sm;
  if (grid > p.ntiles) grid = p.ntiles;
  // with only a few tiles per CTA the counter's round trip costs more than the balance gains (block 2: 3.6 tiles per CTA)
  static const int dyn_min = getenv("RIFE_HK_DYN_MIN") ? atoi(getenv("RIFE_HK_DYN_MIN")) : 4;
  if (p.ntiles < dyn_min * grid) p.tile_ctr = nullptr;
  static const bool debug = getenv("RIFE_HK_DEBUG") != nullptr;
  p.dbg = nullptr;
  static const int pdl_trig = getenv("RIFE_PDL_HALO_TRIGGER") ? atoi(getenv("RIFE_PDL_HALO_TRIGGER")) : 0;
  p.pdl_trig = pdl_trig;
  // (an early trigger does not pay even for the kernels that own a whole SM: measured 5.52 -> 5.99 ms per 16-pair step --
  // the dependent of the dependent is launched early too and can take a freed SM before the kernel in between has placed its
  // CTA there)
  if (debug) {
    cudaMalloc((void**)&p.dbg, (size_t)grid * 128);
    cudaMemsetAsync(p.dbg, 0, (size_t)grid * 128, st);
  }
  const int kind = L.ps_cols ? 2 : (L.kind == CONV_3x3 ? 0 : 1);
  bool done = false;
#define RIFE_HK_CASE(K, KS, N_, MT_, E_)                                                                     \
  if (!done && p.mt == MT_ && kind == K && p.ksteps == KS && p.ne == N_ && p.epi == E_) {                   \
    RIFE_TRY(set_smem_attr((const void*)conv_halo_kernel<K, KS, N_, MT_, E_>));                              \
    { const int rc_ = launch_hk(conv_halo_kernel<K, KS, N_, MT_, E_>, grid, 64 + 32 * EpiWarps<K, N_, E_>::value, smem, st, tm_a, tm_b, p, coop); if (rc_) return rc_; } \
    done = true;                                                                                             \
  }
  RIFE_HK_CASE(1, 1, 16, 2, EPI_LRELU)      // Head.cnn0            (s2d, cin 16 -> 16)
  RIFE_HK_CASE(0, 1, 16, 2, EPI_LRELU)      // Head.cnn1/2          (3x3, 16 -> 16)
  RIFE_HK_CASE(0, 1, 16, 2, EPI_HEADT)      // Head.cnn3            (ConvT as 3x3, 16 -> 4x4)
  RIFE_HK_CASE(1, 8, 16, 1, EPI_LRELU_S2D)  // block4.conv0.0       (s2d, 128 -> 16)
  RIFE_HK_CASE(1, 4, 32, 2, EPI_LRELU)      // block4.conv0.1       (s2d, 64 -> 32)
  RIFE_HK_CASE(0, 2, 32, 2, EPI_RES)        // block4 ResConv       (3x3, 32 -> 32)
  RIFE_HK_CASE(2, 2, 64, 1, EPI_PS)         // block4.lastconv      (ConvT+PS, 32 -> 4x64)
  RIFE_HK_CASE(2, 2, 32, 1, EPI_PS8)        // block4.lastconv, flow + mask only (ConvT+PS, 32 -> 4x32)
  RIFE_HK_CASE(1, 8, 32, 1, EPI_LRELU_S2D)  // block3.conv0.0       (s2d, 128 -> 32)
  RIFE_HK_CASE(1, 8, 64, 1, EPI_LRELU)      // block3.conv0.1       (s2d, 128 -> 64)
  RIFE_HK_CASE(0, 4, 64, 2, EPI_RES)        // block3 ResConv       (3x3, 64 -> 64)
  RIFE_HK_CASE(2, 4, 64, 1, EPI_PS)         // block3.lastconv      (ConvT+PS, 64 -> 4x64)
#undef RIFE_HK_CASE
#define RIFE_HK_STREAM(K, KS, N_, MT_, E_)                                                                   \
  if (!done && L.stream && p.mt == MT_ && kind == K && p.ksteps == KS && p.ne == N_ && p.epi == E_) {       \
    RIFE_TRY(set_smem_attr((const void*)conv_halo_kernel<K, KS, N_, MT_, E_, true>));                        \
    { const int rc_ = launch_hk(conv_halo_kernel<K, KS, N_, MT_, E_, true>, grid, 96 + 32 * EpiWarps<K, N_, E_>::value, smem, st, tm_a, tm_b, p, coop); if (rc_) return rc_; } \
    done = true;                                                                                             \
  }
  RIFE_HK_STREAM(0, 6, 96, 2, EPI_RES)      // block2 ResConv       (3x3, 96 -> 96), two stacked M-tiles
  RIFE_HK_STREAM(0, 6, 96, 1, EPI_RES)
  RIFE_HK_STREAM(0, 8, 128, 1, EPI_RES)     // block1 ResConv       (3x3, 128 -> 128)
  RIFE_HK_STREAM(2, 6, 64, 1, EPI_PS)       // block2.lastconv      (ConvT+PS, 96 -> 4x64)
  RIFE_HK_STREAM(2, 8, 64, 1, EPI_PS)       // block1.lastconv      (ConvT+PS, 128 -> 4x64)
#undef RIFE_HK_STREAM
  if (!done && L.stream && coop) return 1;
  if (!done && L.stream) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "conv_halo: no streamed instantiation for this layer / epilogue");
  if (!done && p.epi == EPI_PS8) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "conv_halo: no EPI_PS8 instantiation for this layer");
  if (!done) {
    RIFE_TRY(set_smem_attr((const void*)conv_halo_kernel<-1, 0, 0, 0, -1>));
    if (coop) return 1;  // no compile-time instantiation for this shape: the caller launches layer by layer
    RIFE_TRY(launch_hk(conv_halo_kernel<-1, 0, 0, 0, -1>, grid, 64 + 32 * EpiWarps<-1, 0, -1>::value, smem, st, tm_a, tm_b, p, false));
  }
  RIFE_LAUNCH_CHECK();
  if (debug) {
    std::vector<long long> h((size_t)grid * 16);
    cudaStreamSynchronize(st);
    cudaMemcpy(h.data(), p.dbg, (size_t)grid * 128, cudaMemcpyDeviceToHost);
    cudaFree(p.dbg);
    double a[8] = {0};
    for (int i = 0; i < grid; ++i)
      for (int k = 0; k < 8; ++k) a[k] += (double)h[(size_t)i * 16 + k] / grid;
    // slots 8..11: %globaltimer (ns) at CTA entry, after the prologue's __syncthreads, at the end of the MMA role, at CTA exit
    long long g0 = LLONG_MAX, g1 = 0;
    double pro = 0, life = 0, tail = 0;
    for (int i = 0; i < grid; ++i) {
      const long long* r = &h[(size_t)i * 16];
      g0 = std::min(g0, r[8]); g1 = std::max(g1, r[11]);
      pro += (double)(r[9] - r[8]) / grid; life += (double)(r[11] - r[8]) / grid; tail += (double)(r[11] - r[10]) / grid;
    }
    long long first_exit = LLONG_MAX, last_entry = 0;
    for (int i = 0; i < grid; ++i) { first_exit = std::min(first_exit, h[(size_t)i * 16 + 11]); last_entry = std::max(last_entry, h[(size_t)i * 16 + 8]); }
    fprintf(stderr, "[hk-t] span first entry -> last exit %.1f us | last entry at +%.1f us, first exit at +%.1f us | per CTA: prologue %.1f us, lifetime %.1f us, after the last MMA %.1f us\n",
            (g1 - g0) / 1e3, (last_entry - g0) / 1e3, (first_exit - g0) / 1e3, pro / 1e3, life / 1e3, tail / 1e3);
    fprintf(stderr, "[hk] kind=%d cin=%d ne=%d mt=%d grid=%d per_sm=%d stages=%d steps/cta=%.1f | prod wait a_empty %.0f | mma: wait w %.0f, t_empty %.0f, a_full %.0f, total %.0f | epi: wait t_full %.0f, total %.0f\n",
            kind, L.cin, L.ne, p.mt, grid, per_sm, p.stages, a[7], a[0], a[1], a[2], a[3], a[4], a[5], a[6]);
  }
  return RIFE_OK;
}

}  // namespace

int launch_conv_halo_f32io(const HaloLayer& L, const ConvIO& io, cudaStream_t st) {
  // NHWC fp32 in (cin channels) -> planar-8 fp16; planar / pixel-major fp16 out -> NHWC fp32
  const size_t nin = (size_t)io.B * io.H * io.W * L.cin;
  size_t nout = 0;
  bool planar_out = false;
  int oh = io.H, ow = io.W, oc = io.out_c;
  switch (io.epi) {
    case EPI_PS: nout = (size_t)io.B * 16 * io.H * io.W * TMP_C; break;
    case EPI_PS8: nout = (size_t)io.B * 16 * io.H * io.W * 8; break;
    case EPI_HEADT:
    case EPI_CONVT: nout = (size_t)io.B * 4 * io.H * io.W * io.out_c; break;
    case EPI_LRELU_S2D: RIFE_FAIL(RIFE_ERR_INVALID_ARG, "conv_halo_f32io: s2d store is pipeline-only");
    default: nout = (size_t)io.B * io.H * io.W * io.out_c; planar_out = true; break;
  }
  const size_t nres = io.res ? (size_t)io.B * io.H * io.W * io.out_c : 0;
  __half *hin = nullptr, *hout = nullptr, *hres = nullptr;
  auto cleanup = [&]() {
    cudaStreamSynchronize(st);
    cudaFree(hin); cudaFree(hout); cudaFree(hres);
  };
  cudaError_t e = cudaMalloc((void**)&hin, nin * 2);
  if (e == cudaSuccess) e = cudaMalloc((void**)&hout, nout * 2);
  if (e == cudaSuccess && nres) e = cudaMalloc((void**)&hres, nres * 2);
  if (e == cudaSuccess) e = cudaMemsetAsync(hout, 0, nout * 2, st);
  if (e != cudaSuccess) {
    cleanup();
    RIFE_FAIL(RIFE_ERR_CUDA, std::string("conv_halo_f32io: ") + cudaGetErrorString(e));
  }
  nhwc_f32_to_planar8_kernel<<<(unsigned)((nin + 255) / 256), 256, 0, st>>>((const float*)io.in, hin, io.B, io.H, io.W, L.cin,
                                                                           io.in_bf16 != 0);
  if (nres)
    nhwc_f32_to_planar8_kernel<<<(unsigned)((nres + 255) / 256), 256, 0, st>>>((const float*)io.res, hres, io.B, io.H, io.W,
                                                                             io.out_c, io.in_bf16 != 0);
  ConvIO h = io;
  h.in = hin; h.res = hres; h.out = hout;
  int rc = launch_conv_halo(L, h, st);
  if (rc == RIFE_OK) {
    if (planar_out)
      planar8_to_nhwc_f32_kernel<<<(unsigned)((nout + 255) / 256), 256, 0, st>>>(hout, (float*)io.out, io.B, oh, ow, oc,
                                                                                 io.out_bf16 != 0);
    else
      f16_to_f32_k<<<(unsigned)((nout + 255) / 256), 256, 0, st>>>(hout, (float*)io.out, nout);
    if (cudaGetLastError() != cudaSuccess) rc = RIFE_ERR_CUDA;
  }
  cleanup();
  return rc;
}

}  // namespace rife
