// tcgen05 (5th-gen tensor core) implicit-GEMM convolution for sm_100a, fp16 operands / fp32 TMEM accumulators.
//
//   out[p][n] = sum_{tap} sum_{ci} act[p + off(tap)][ci] * w[tap][ci][n]        (p = pixel of an NHWC tensor)
//
// One CTA per 128-pixel tile (TH x TW output pixels = the 128 rows of one UMMA M=128 accumulator).
//   warp 0   : TMA producer.  Per (tap, K chunk): one 4-D tiled TMA load of the activation box
//              {KC ch, TW, TH, 1} at (x0+dx, y0+dy) -- out-of-image coordinates (the conv's zero padding and the
//              ragged last tiles) are zero-filled by the TMA unit -- plus one 2-D load of the weight tile
//              {KC, Npad}.  Both land K-major with the hardware swizzle that matches KC*2 bytes (128B/64B/32B).
//   warp 1   : allocates TMEM, then one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=Npad,
//              K=16) straight from the swizzled smem tiles; tcgen05.commit releases each smem stage back to the
//              producer and finally signals the epilogue.
//   warps 2-5: epilogue.  tcgen05.ld (32 lanes x 16 columns) -> registers -> fused tail of the layer
//              (bias, LeakyReLU, beta*x + residual, space-to-depth store, ConvTranspose parity scatter +
//              PixelShuffle) -> 16/32-byte vector stores of fp16 NHWC.
// Reference semantics: train_log/IFNet_HDv3.py:9-14 (conv), :24-43 (Head), :45-54 (ResConv), :73-76 (lastconv).
#include <cuda.h>

#include "conv_umma.cuh"

namespace rife {
namespace {

constexpr int TILE_M = 128;
constexpr int NUM_THREADS = 192;
constexpr int MAX_STAGES = 8;
constexpr int MAX_TAPS = 9;

struct UmmaParams {
  int B, H, W;
  int tw, th, tiles_x, tiles_y;
  int ntaps, nchunks, kc, stages;
  int npad, n_valid, tmem_cols, nw;  // nw = accumulator columns per CTA (N split over blockIdx.y)
  int a_bytes, b_bytes;  // per stage (b rounded up to 1024)
  int epi, out_c, cpp, planar_out, res_pitch, out_f32;
  int in_bf, out_bf;  // operand / activation-output storage type: 0 = fp16, 1 = bf16 (RIFE_PRECISION_BF16)
  unsigned long long kmask;  // bit ks = tap * nchunks + chunk set: that K block has non-zero weights (all ones if > 64 blocks)
  const float* bias;
  const float* beta;
  const __half* res;
  __half* out;
  signed char dy[MAX_TAPS], dx[MAX_TAPS];
};

// ------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded wait: a mis-programmed pipeline traps instead of hanging the GPU
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
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major operand tile, rows of `kc` fp16 (= swizzle span), 8-row swizzle atoms stacked every 8*kc*2 bytes
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);            // start address       [0,14)
  d |= (uint64_t)1 << 16;                            // LBO (unused for swizzled K-major) [16,30)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;  // stride byte offset  [32,46)
  d |= (uint64_t)1 << 46;                            // descriptor version 1 (sm_100)     [46,48)
  d |= (uint64_t)layout_type << 61;                  // swizzle mode        [61,64)
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

__device__ __forceinline__ void store16h(__half* dst, const float (&v)[16], bool bf = false) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = v[i];
  st8x(dst, a, bf);
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = v[8 + i];
  st8x(dst + 8, a, bf);
}

// ------------------------------------------------------------------------------------------ kernel
// grid = (pixel tiles, N splits): each CTA computes `nw` accumulator columns [blockIdx.y * nw, +nw) of one 128-pixel
// tile.  Splitting N gives the small-spatial / wide-channel layers (IFBlock 0-2) enough CTAs to cover the machine
// and shrinks the per-stage weight tile, so the TMA ring is deeper.
template <int EPI>
__global__ void __launch_bounds__(NUM_THREADS) conv_umma_kernel(const __grid_constant__ CUtensorMap tm_a,
                                                                const __grid_constant__ CUtensorMap tm_b,
                                                                const UmmaParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float s_bias[256];
  __shared__ __align__(16) float s_beta[256];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tiles_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B atoms need 1024-byte alignment
  const uint32_t stage_bytes = (uint32_t)(p.a_bytes + p.b_bytes);

  int tile = blockIdx.x;
  const int tx = tile % p.tiles_x;
  tile /= p.tiles_x;
  const int ty = tile % p.tiles_y;
  const int b = tile / p.tiles_y;
  const int x0 = tx * p.tw, y0 = ty * p.th;
  const int n_off = blockIdx.y * p.nw;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_b) : "memory");
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(&tmem_full_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp >= 2) {
    for (int i = threadIdx.x - 64; i < p.nw; i += NUM_THREADS - 64) {
      s_bias[i] = p.bias[n_off + i];
      if (EPI == EPI_RES || EPI >= EPI_PRELU) s_beta[i] = p.beta[n_off + i];
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // no explicit launch_dependents: measured on B200, early-resident dependents of the conv kernels unbalance the next
  // kernel's CTA placement (-10 % at 8 pairs); the implicit trigger at CTA exit keeps the launch-latency overlap.
  pdl_wait();  // activations and residual come from the previous kernel (bias / slope / weights are constants)
  const uint32_t tmem_base = tmem_slot;
  const int ksteps = p.ntaps * p.nchunks;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = (uint32_t)(p.a_bytes + p.nw * p.kc * 2);
      for (int ks = 0; ks < ksteps; ++ks) {
        if (ks < 64 && !((p.kmask >> ks) & 1ull)) continue;  // structurally zero K block (stride-2 convs on s2d input)
        const int tap = ks / p.nchunks, chunk = ks - tap * p.nchunks;
        mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
        const uint32_t fb = smem_u32(&full_bar[stage]);
        mbar_expect_tx(fb, tx_bytes);
        const uint32_t sa = tiles_base + stage * stage_bytes;
        tma_load_4d(sa, &tm_a, fb, chunk * p.kc, x0 + p.dx[tap], y0 + p.dy[tap], b);
        tma_load_2d(sa + p.a_bytes, &tm_b, fb, 0, ks * p.npad + n_off);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- MMA issuer
      // instruction descriptor: D=f32 [4,6)=1, A=B=f16 (0), both K-major, N>>3 at [17,23), M>>4 at [24,29)
      // (A / B format bits 7-9 / 10-12: 1 = bf16)
      const uint32_t idesc = (1u << 4) | (p.in_bf ? ((1u << 7) | (1u << 10)) : 0u) | ((uint32_t)(p.nw >> 3) << 17) |
                             ((uint32_t)(TILE_M >> 4) << 24);
      const uint32_t sbo = 8u * (uint32_t)p.kc * 2u;
      const uint32_t layout = p.kc == 64 ? 2u : (p.kc == 32 ? 4u : 6u);  // SWIZZLE_128B / 64B / 32B
      const uint64_t desc0 = make_smem_desc(0, sbo, layout);
      const uint32_t a16 = (uint32_t)p.a_bytes >> 4;
      const int ksub = p.kc >> 4;
      int stage = 0;
      uint32_t phase = 0, accum = 0;
      for (int ks = 0; ks < ksteps; ++ks) {
        if (ks < 64 && !((p.kmask >> ks) & 1ull)) continue;
        mbar_wait(smem_u32(&full_bar[stage]), phase);
        tc_fence_after();
        const uint64_t da = desc0 + (uint64_t)((tiles_base + stage * stage_bytes) >> 4);
        const uint64_t db = da + a16;
        // advance 16 fp16 = 32 bytes along K inside the swizzle span
        if (ksub == 4) {
          umma_f16(tmem_base, da, db, idesc, accum);
          umma_f16(tmem_base, da + 2, db + 2, idesc, 1u);
          umma_f16(tmem_base, da + 4, db + 4, idesc, 1u);
          umma_f16(tmem_base, da + 6, db + 6, idesc, 1u);
        } else if (ksub == 2) {
          umma_f16(tmem_base, da, db, idesc, accum);
          umma_f16(tmem_base, da + 2, db + 2, idesc, 1u);
        } else {
          umma_f16(tmem_base, da, db, idesc, accum);
        }
        accum = 1u;
        umma_commit(smem_u32(&empty_bar[stage]));
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      umma_commit(smem_u32(&tmem_full_bar));
    }
  } else {
    // ---------------- epilogue: warp w owns TMEM lanes 32*(w%4) .. +31  (accumulator row = pixel of the tile)
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int py_ = m / p.tw, px_ = m - py_ * p.tw;
    const int y = y0 + py_, x = x0 + px_;
    const bool valid = (y < p.H) && (x < p.W);
    const int nchunk16 = p.nw >> 4;
    const size_t pix = ((size_t)b * p.H + y) * p.W + x;
    // ResConv: fetch the residual while the MMAs run (it does not depend on them)
    constexpr int MAXRES = 6;
    uint4 resv[2 * MAXRES];
    if constexpr (EPI == EPI_RES || EPI == EPI_PRELU_ADD) {
      if (valid) {
        const uint4* rp = reinterpret_cast<const uint4*>(p.res + pix * p.res_pitch + n_off);
#pragma unroll
        for (int j = 0; j < MAXRES; ++j)
          if (j < nchunk16) { resv[2 * j] = __ldg(rp + 2 * j); resv[2 * j + 1] = __ldg(rp + 2 * j + 1); }
      }
    }
    mbar_wait(smem_u32(&tmem_full_bar), 0);
    tc_fence_after();
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int jb = 0; jb < nchunk16; jb += MAXRES) {
#pragma unroll
      for (int jj = 0; jj < MAXRES; ++jj) {
        const int j = jb + jj;
        if (j >= nchunk16) break;
        float v[16];
        __syncwarp();
        tmem_ld16(trow + (uint32_t)(j * 16), v);
        if (!valid) continue;
        const int nl = j * 16;        // column inside this CTA's slice
        const int n0 = n_off + nl;    // global GEMM column
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 bb = *reinterpret_cast<const float4*>(s_bias + nl + 4 * i);
          v[4 * i] += bb.x; v[4 * i + 1] += bb.y; v[4 * i + 2] += bb.z; v[4 * i + 3] += bb.w;
        }
        if constexpr (EPI == EPI_LRELU || EPI == EPI_LINEAR) {
          if constexpr (EPI == EPI_LRELU) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = lrelu02(v[i]);
          }
          if (n0 < p.n_valid) {
            if (EPI == EPI_LINEAR && p.out_f32) {
              // flow / mask / refine heads keep fp32 (no rounding of a value that is later scaled and accumulated)
              float4* d = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + pix * p.out_c + n0);
#pragma unroll
              for (int i = 0; i < 4; ++i) d[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            } else if (p.planar_out) {
              // planar-8 destination [B][out_c/8][H][W][8]: the consumer is the halo kernel (conv_halo.cu)
              const size_t hw = (size_t)p.H * p.W;
              __half* d = p.out + (((size_t)b * (p.out_c >> 3) + (n0 >> 3)) * hw + (size_t)y * p.W + x) * 8;
              float a8[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) a8[i] = v[i];
              st8x(d, a8, p.out_bf);
#pragma unroll
              for (int i = 0; i < 8; ++i) a8[i] = v[8 + i];
              st8x(d + hw * 8, a8, p.out_bf);
            } else {
              store16h(p.out + pix * p.out_c + n0, v, p.out_bf);
            }
          }
        } else if constexpr (EPI == EPI_LRELU_S2D) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = lrelu02(v[i]);
          const size_t o = ((((size_t)b * (p.H >> 1) + (y >> 1)) * (p.W >> 1) + (x >> 1)) * 4 + ((y & 1) * 2 + (x & 1)));
          if (n0 < p.n_valid) store16h(p.out + o * p.out_c + n0, v, p.out_bf);
        } else if constexpr (EPI == EPI_RES) {
          uint4 r2[2];
          if (jb == 0) { r2[0] = resv[2 * jj]; r2[1] = resv[2 * jj + 1]; }
          else {  // slices wider than 16*MAXRES columns: fetch late (not used by the network's shapes)
            const uint4* rp = reinterpret_cast<const uint4*>(p.res + pix * p.res_pitch + n0);
            r2[0] = __ldg(rp); r2[1] = __ldg(rp + 1);
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float rr[8];
            unpack8x(r2[h], p.in_bf, rr);
            const float4 b0 = *reinterpret_cast<const float4*>(s_beta + nl + 8 * h);
            const float4 b1 = *reinterpret_cast<const float4*>(s_beta + nl + 8 * h + 4);
            const float be[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) v[8 * h + i] = lrelu02(v[8 * h + i] * be[i] + rr[i]);
          }
          store16h(p.out + pix * p.out_c + n0, v, p.out_bf);
        } else if constexpr (EPI == EPI_PRELU || EPI == EPI_PRELU_ADD) {
          // ECCV-paper net: per-channel PReLU (slopes in s_beta), optionally "+ res" afterwards (model/IFNet.py:46)
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = v[i] >= 0.f ? v[i] : v[i] * s_beta[nl + i];
          if constexpr (EPI == EPI_PRELU_ADD) {
            uint4 r2[2];
            if (jb == 0) { r2[0] = resv[2 * jj]; r2[1] = resv[2 * jj + 1]; }
            else {
              const uint4* rp = reinterpret_cast<const uint4*>(p.res + pix * p.res_pitch + n0);
              r2[0] = __ldg(rp); r2[1] = __ldg(rp + 1);
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const __half2* hh = reinterpret_cast<const __half2*>(&r2[h]);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 f = __half22float2(hh[i]);
                v[8 * h + 2 * i] += f.x;
                v[8 * h + 2 * i + 1] += f.y;
              }
            }
          }
          if (n0 < p.n_valid) store16h(p.out + pix * p.out_c + n0, v);
        } else if constexpr (EPI == EPI_CONVT_PRELU) {
          if ((p.cpp & 15) == 0) {
            // a 16-column chunk lies inside one output parity: one 32-byte store per output pixel
            if (n0 < p.n_valid) {
              const int par = n0 / p.cpp, co = n0 - par * p.cpp;
              const int oy = 2 * y + (par >> 1), ox = 2 * x + (par & 1);
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = v[i] >= 0.f ? v[i] : v[i] * s_beta[nl + i];
              store16h(p.out + (((size_t)b * 2 * p.H + oy) * 2 * p.W + ox) * p.out_c + co, v);
            }
          } else
          for (int i = 0; i < 16; ++i) {
            const int n = n0 + i;
            if (n >= p.n_valid) break;
            const int par = n / p.cpp, co = n - par * p.cpp;
            const int oy = 2 * y + (par >> 1), ox = 2 * x + (par & 1);
            const float r = v[i] >= 0.f ? v[i] : v[i] * s_beta[nl + i];
            p.out[(((size_t)b * 2 * p.H + oy) * 2 * p.W + ox) * p.out_c + co] = __float2half_rn(r);
          }
        } else if constexpr (EPI == EPI_PS) {
          // permuted columns: global chunk jg = par*4 + (dy*2+dx) holds the 13 (+3 zero) channels of one output pixel
          const int jg = n0 >> 4;
          const int par = jg >> 2, sub = jg & 3;
          const int oy = 4 * y + 2 * (par >> 1) + (sub >> 1), ox = 4 * x + 2 * (par & 1) + (sub & 1);
          store16h(p.out + (((size_t)b * 4 * p.H + oy) * 4 * p.W + ox) * TMP_C, v);
        } else if constexpr (EPI == EPI_HEADT) {
#pragma unroll
          for (int par = 0; par < 4; ++par) {
            const int oy = 2 * y + (par >> 1), ox = 2 * x + (par & 1);
            float f[4] = {v[4 * par], v[4 * par + 1], v[4 * par + 2], v[4 * par + 3]};
            st4(p.out + (((size_t)b * 2 * p.H + oy) * 2 * p.W + ox) * p.out_c, f);
          }
        } else if constexpr (EPI == EPI_CONVT) {
          for (int i = 0; i < 16; ++i) {
            const int n = n0 + i;
            if (n >= p.n_valid) break;
            const int par = n / p.cpp, co = n - par * p.cpp;
            const int oy = 2 * y + (par >> 1), ox = 2 * x + (par & 1);
            const size_t o = (((size_t)b * 2 * p.H + oy) * 2 * p.W + ox) * p.out_c + co;
            if (p.out_f32) reinterpret_cast<float*>(p.out)[o] = v[i];
            else p.out[o] = __float2half_rn(v[i]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                 : "memory");
  }
}

// ------------------------------------------------------------------------------------------ casts for the f32io wrapper
__global__ void f32_to_f16_kernel(const float* __restrict__ x, __half* __restrict__ y, size_t n, bool bf) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (bf) {
    const __nv_bfloat16 b = __float2bfloat16_rn(x[i]);
    y[i] = *reinterpret_cast<const __half*>(&b);
  } else {
    y[i] = __float2half_rn(x[i]);
  }
}
__global__ void f16_to_f32_kernel(const __half* __restrict__ x, float* __restrict__ y, size_t n, bool bf) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = bf ? __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(&x[i])) : __half2float(x[i]);
}

// ------------------------------------------------------------------------------------------ tensor maps
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

CUtensorMapSwizzle swizzle_for(int kc) {
  return kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

}  // namespace

int umma_pack_weights(const HostLayer& h, bool ps_layout, std::vector<__half>* packed, std::vector<float>* bias,
                      std::vector<float>* beta, UmmaLayer* L, bool bf16) {
  if (h.cin % 16) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "umma_pack_weights: cin must be a multiple of 16");
  const int kc = (h.cin % 64 == 0) ? 64 : ((h.cin % 32 == 0) ? 32 : 16);
  const int nchunks = h.cin / kc;
  std::vector<int> colmap;  // new column -> old column (or -1 = zero column)
  if (ps_layout) {
    if (h.convt_co != 52) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "umma_pack_weights: PixelShuffle layout is the 52-channel lastconv");
    colmap.assign(256, -1);
    for (int par = 0; par < 4; ++par)
      for (int sub = 0; sub < 4; ++sub)
        for (int k = 0; k < 13; ++k) colmap[(par * 4 + sub) * 16 + k] = par * 52 + 4 * k + sub;  // co = 4k + 2dy + dx
  } else {
    colmap.resize(h.npad);
    for (int i = 0; i < h.npad; ++i) colmap[i] = i;
  }
  const int np = (int)colmap.size();
  if (np % 16 || (np > 256 && np % 128)) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "umma_pack_weights: N must be a multiple of 16 (of 128 above 256)");
  packed->assign((size_t)h.taps * nchunks * np * kc, __float2half_rn(0.f));
  for (int t = 0; t < h.taps; ++t)
    for (int ch = 0; ch < nchunks; ++ch)
      for (int n = 0; n < np; ++n) {
        const int on = colmap[n];
        if (on < 0) continue;
        for (int k = 0; k < kc; ++k)
          (*packed)[(((size_t)t * nchunks + ch) * np + n) * kc + k] =
              host_cvt16(h.w[((size_t)t * h.cin + ch * kc + k) * h.npad + on], bf16);
      }
  bias->assign(np, 0.f);
  beta->clear();
  if (!h.beta.empty()) beta->assign(np, 0.f);
  for (int n = 0; n < np; ++n) {
    if (colmap[n] < 0) continue;
    (*bias)[n] = h.bias[colmap[n]];
    if (!h.beta.empty()) (*beta)[n] = h.beta[colmap[n]];
  }
  L->kind = h.kind; L->taps = h.taps; L->cin = h.cin; L->kc = kc;
  {
    // K blocks (tap, kc-channel chunk) whose weights are all zero -- the structural zeros of a stride-2 conv written as 2x2
    // taps over the space-to-depth input -- are neither loaded nor multiplied
    const int nch = h.cin / kc;
    unsigned long long m = ~0ull;
    if (h.taps * nch <= 64) {
      m = 0;
      for (int t = 0; t < h.taps; ++t)
        for (int k = 0; k < nch; ++k) {
          bool nz = false;
          for (int c = k * kc; c < (k + 1) * kc && !nz; ++c)
            for (int n = 0; n < h.npad && !nz; ++n) nz = h.w[((size_t)t * h.cin + c) * h.npad + n] != 0.f;
          if (nz) m |= 1ull << (t * nch + k);
        }
      if (!m) m = 1;  // an all-zero layer still has to write its accumulator
    }
    L->kmask = m;
  }
  L->n = ps_layout ? 256 : h.n;
  L->npad = np;
  L->ps_cols = ps_layout ? 1 : 0;
  return RIFE_OK;
}

int launch_conv_umma(const UmmaLayer& L, const ConvIO& io, cudaStream_t st) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) RIFE_FAIL(RIFE_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  if (io.epi == EPI_PS && !L.ps_cols) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "conv_umma: EPI_PS needs the PixelShuffle column layout");
  if (io.epi == EPI_RES && (!L.beta || !io.res)) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "conv_umma: EPI_RES needs beta and res");
  if (io.epi >= EPI_PRELU && !L.beta) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "conv_umma: PReLU epilogues need the slope vector");
  if (io.epi == EPI_PRELU_ADD && !io.res) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "conv_umma: EPI_PRELU_ADD needs res");
  if (io.out_f32 && io.epi != EPI_LINEAR && io.epi != EPI_CONVT) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "conv_umma: fp32 output is for the activation-free epilogues only");
  if (io.epi == EPI_HEADT && L.npad != 16) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "conv_umma: EPI_HEADT expects N = 16");

  UmmaParams p{};
  p.B = io.B; p.H = io.H; p.W = io.W;
  p.tw = 16; p.th = 8;
  p.tiles_x = (io.W + p.tw - 1) / p.tw;
  p.tiles_y = (io.H + p.th - 1) / p.th;
  p.ntaps = L.taps; p.kc = L.kc; p.nchunks = L.cin / L.kc;
  p.npad = L.npad; p.n_valid = L.n;
  // N split: 64-column slices when the layer is wide and has too few pixel tiles to fill the machine
  p.nw = L.npad;
  if (L.npad >= 128 && L.npad % 64 == 0 && p.tiles_x * p.tiles_y * io.B <= 148 / 2 + 16 && io.epi != EPI_HEADT) p.nw = 64;
  if (p.nw > 256) p.nw = (L.npad % 256 == 0) ? 256 : 128;  // one MMA / TMEM accumulator holds at most 256 columns
  p.tmem_cols = 32;
  while (p.tmem_cols < p.nw) p.tmem_cols <<= 1;
  p.a_bytes = TILE_M * L.kc * 2;
  p.b_bytes = ((p.nw * L.kc * 2) + 1023) / 1024 * 1024;
  const int stage_bytes = p.a_bytes + p.b_bytes;
  int stages = (104 * 1024) / stage_bytes;  // two CTAs per SM stay co-resident (no second wave, overlapped epilogues)
  if (stages < 2) stages = 2;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  // a grid of at most one CTA per SM (the coarse blocks' layers) has the SM's shared memory to itself: a deeper ring keeps
  // more of the latency-bound (tap, chunk) loads in flight.  RIFE_UMMA_DEEP=0 restores the two-CTA budget.
  {
    static const bool deep = !(getenv("RIFE_UMMA_DEEP") && atoi(getenv("RIFE_UMMA_DEEP")) == 0);
    static const int nsm = []() {
      int n = 0, dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
      return n > 0 ? n : 148;
    }();
    const long long ctas = (long long)p.tiles_x * p.tiles_y * io.B * (L.npad / p.nw);
    if (deep && ctas <= nsm) {
      stages = (198 * 1024) / stage_bytes;
      if (stages > MAX_STAGES) stages = MAX_STAGES;
    }
  }
  const int ksteps = p.ntaps * p.nchunks;
  if (stages > ksteps) stages = ksteps;
  p.stages = stages;
  p.epi = io.epi; p.out_c = io.out_c; p.cpp = io.co_per_par; p.planar_out = io.planar_out;
  p.out_f32 = io.out_f32;
  p.res_pitch = io.res_pitch ? io.res_pitch : io.out_c;
  p.bias = L.bias; p.beta = L.beta;
  p.res = (const __half*)io.res; p.out = (__half*)io.out;
  p.in_bf = io.in_bf16 ? 1 : 0; p.out_bf = io.out_bf16 ? 1 : 0;
  p.kmask = L.kmask;
  if (L.kind == CONV_3x3) {
    for (int t = 0; t < 9; ++t) { p.dy[t] = (signed char)(t / 3 - 1); p.dx[t] = (signed char)(t % 3 - 1); }
  } else {
    for (int t = 0; t < 4; ++t) { p.dy[t] = (signed char)(t / 2 - 1); p.dx[t] = (signed char)(t % 2 - 1); }
  }

  CUtensorMap tm_a, tm_b;
  {
    cuuint64_t dims[4] = {(cuuint64_t)L.cin, (cuuint64_t)io.W, (cuuint64_t)io.H, (cuuint64_t)io.B};
    const cuuint64_t pitch = io.in_pitch ? io.in_pitch : L.cin;  // the conv may read a channel slice of a wider tensor
    cuuint64_t strides[3] = {pitch * 2, (cuuint64_t)io.W * pitch * 2, (cuuint64_t)io.H * io.W * pitch * 2};
    cuuint32_t box[4] = {(cuuint32_t)L.kc, (cuuint32_t)p.tw, (cuuint32_t)p.th, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&tm_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(io.in), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(L.kc), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) RIFE_FAIL(RIFE_ERR_CUDA, "cuTensorMapEncodeTiled(activation) failed: " + std::to_string((int)r));
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)L.kc, (cuuint64_t)L.taps * p.nchunks * L.npad};
    cuuint64_t strides[1] = {(cuuint64_t)L.kc * 2};
    cuuint32_t box[2] = {(cuuint32_t)L.kc, (cuuint32_t)p.nw};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tm_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(L.w), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(L.kc), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) RIFE_FAIL(RIFE_ERR_CUDA, "cuTensorMapEncodeTiled(weights) failed: " + std::to_string((int)r));
  }
  const int smem = stages * stage_bytes + 1024;
  const dim3 grid((unsigned)(p.tiles_x * p.tiles_y * io.B), (unsigned)(L.npad / p.nw));
#define RIFE_UMMA_LAUNCH(E_)                                                                                   \
  case E_: {                                                                                                   \
    static bool attr_set[64] = {};                                                                             \
    int dev = 0;                                                                                               \
    cudaGetDevice(&dev);                                                                                       \
    if (dev < 64 && !attr_set[dev]) {                                                                          \
      RIFE_CUDA_TRY(cudaFuncSetAttribute(conv_umma_kernel<E_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); \
      attr_set[dev] = true;                                                                                    \
    }                                                                                                          \
    RIFE_CUDA_TRY(launch_pdl(conv_umma_kernel<E_>, grid, dim3(NUM_THREADS), smem, st, tm_a, tm_b, p));          \
  } break;
  switch (io.epi) {
    RIFE_UMMA_LAUNCH(EPI_LRELU)
    RIFE_UMMA_LAUNCH(EPI_LRELU_S2D)
    RIFE_UMMA_LAUNCH(EPI_RES)
    RIFE_UMMA_LAUNCH(EPI_PS)
    RIFE_UMMA_LAUNCH(EPI_HEADT)
    RIFE_UMMA_LAUNCH(EPI_LINEAR)
    RIFE_UMMA_LAUNCH(EPI_CONVT)
    RIFE_UMMA_LAUNCH(EPI_PRELU)
    RIFE_UMMA_LAUNCH(EPI_PRELU_ADD)
    RIFE_UMMA_LAUNCH(EPI_CONVT_PRELU)
    default: RIFE_FAIL(RIFE_ERR_INVALID_ARG, "conv_umma: bad epilogue");
  }
#undef RIFE_UMMA_LAUNCH
  RIFE_LAUNCH_CHECK();
  return RIFE_OK;
}

int launch_conv_umma_f32io(const UmmaLayer& L, const ConvIO& io, cudaStream_t st) {
  size_t nin = (size_t)io.B * io.H * io.W * L.cin, nout = 0, nres = 0;
  switch (io.epi) {
    case EPI_PS: nout = (size_t)io.B * 16 * io.H * io.W * TMP_C; break;
    case EPI_HEADT:
    case EPI_CONVT: nout = (size_t)io.B * 4 * io.H * io.W * io.out_c; break;
    default: nout = (size_t)io.B * io.H * io.W * io.out_c; break;
  }
  if (io.res) nres = (size_t)io.B * io.H * io.W * io.out_c;
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
    RIFE_FAIL(RIFE_ERR_CUDA, std::string("conv_umma_f32io: ") + cudaGetErrorString(e));
  }
  f32_to_f16_kernel<<<(unsigned)((nin + 255) / 256), 256, 0, st>>>((const float*)io.in, hin, nin, io.in_bf16 != 0);
  if (nres) f32_to_f16_kernel<<<(unsigned)((nres + 255) / 256), 256, 0, st>>>((const float*)io.res, hres, nres, io.in_bf16 != 0);
  ConvIO h = io;
  h.in = hin; h.res = hres; h.out = hout;
  int rc = launch_conv_umma(L, h, st);
  if (rc == RIFE_OK) {
    // PixelShuffle / ConvTranspose scatter epilogues always store fp16
    const bool obf = io.out_bf16 && (io.epi == EPI_LRELU || io.epi == EPI_LINEAR || io.epi == EPI_RES || io.epi == EPI_LRELU_S2D);
    f16_to_f32_kernel<<<(unsigned)((nout + 255) / 256), 256, 0, st>>>(hout, (float*)io.out, nout, obf);
    if (cudaGetLastError() != cudaSuccess) rc = RIFE_ERR_CUDA;
  }
  cleanup();
  return rc;
}

}  // namespace rife
