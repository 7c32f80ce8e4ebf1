// fp32 FMA implicit-GEMM convolution (RIFE_PRECISION_FP32 backbone): NHWC activations, every conv of the
// network expressed as a stride-1 multi-tap GEMM  out[p][n] = sum_t sum_ci in[p + off_t][ci] * w[t][ci][n]
//   * 3x3 / stride 1        : 9 taps                                   (ResConv, Head.cnn1/2; IFNet_HDv3.py:28-29,48)
//   * 3x3 / stride 2        : 4 taps on the space-to-depth input        (conv0.*, Head.cnn0; IFNet_HDv3.py:27,60-61)
//   * ConvTranspose2d(4,2,1): 9 taps, N = 4 parities x Cout (zero weights where a tap does not feed a parity)
//                                                                       (lastconv, Head.cnn3; IFNet_HDv3.py:30,74)
// with the layer's elementwise tail fused into the epilogue (bias, beta*x + residual, LeakyReLU, PixelShuffle
// scatter).  CTA tile: 8 x 16 output pixels x NT output channels, 256 threads, K chunked by 8 input channels.
#include "common.cuh"

namespace rife {
namespace {

constexpr int TH = 8, TW = 16, KC = 8, NTHREADS = 256;

template <int KIND, int NT>
__global__ void __launch_bounds__(NTHREADS) conv_f32_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                            const float* __restrict__ bias,
                                                            const float* __restrict__ beta,
                                                            const float* __restrict__ res, float* __restrict__ out,
                                                            int B, int H, int W, int cin, int n_valid, int npad,
                                                            int epi, int out_c, int cpp, int in_pitch, int res_pitch) {
  constexpr int RH = (KIND == CONV_3x3) ? 2 : 1;  // halo extent: taps cover rows/cols [-1, RH-1]
  constexpr int NTAP = (RH + 1) * (RH + 1);
  constexpr int IH = TH + RH, IW = TW + RH;
  constexpr int TN = NT / 4;                 // threads along N (4 couts each)
  constexpr int GROUPS = NTHREADS / TN;      // pixel groups
  constexpr int PPT = (TH * TW) / GROUPS;    // pixels per thread (consecutive along x)
  constexpr int GPR = TW / PPT;              // groups per tile row
  static_assert(PPT >= 1 && GPR >= 1 && TW % PPT == 0, "tile config");

  __shared__ float in_s[IH][KC][IW + 1];
  __shared__ __align__(16) float w_s[NTAP][KC][NT];

  const int tid = threadIdx.x;
  const int tn = tid % TN, g = tid / TN;
  const int row = g / GPR, xs = (g % GPR) * PPT;
  const int ntiles_n = (npad + NT - 1) / NT;
  const int b = blockIdx.z / ntiles_n, nt = blockIdx.z % ntiles_n;
  const int n0 = nt * NT;
  const int y0 = blockIdx.y * TH, x0 = blockIdx.x * TW;

  float acc[PPT][4];
#pragma unroll
  for (int j = 0; j < PPT; ++j)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[j][q] = 0.f;

  const float* in_b = in + (size_t)b * H * W * in_pitch;
  for (int ci0 = 0; ci0 < cin; ci0 += KC) {
    // ---- stage the input halo tile (transposed to [row][kc][col]) and this chunk's weights
    for (int i = tid; i < IH * IW * 2; i += NTHREADS) {
      const int half = i & 1, pix = i >> 1;
      const int r = pix / IW, c = pix % IW;
      const int gy = y0 + r - 1, gx = x0 + c - 1;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gy >= 0 && gy < H && gx >= 0 && gx < W)
        v = *reinterpret_cast<const float4*>(in_b + ((size_t)gy * W + gx) * in_pitch + ci0 + half * 4);
      in_s[r][half * 4 + 0][c] = v.x;
      in_s[r][half * 4 + 1][c] = v.y;
      in_s[r][half * 4 + 2][c] = v.z;
      in_s[r][half * 4 + 3][c] = v.w;
    }
    for (int i = tid; i < NTAP * KC * (NT / 4); i += NTHREADS) {
      const int c4 = i % (NT / 4), kc = (i / (NT / 4)) % KC, t = i / ((NT / 4) * KC);
      const int n = n0 + c4 * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n < npad) v = *reinterpret_cast<const float4*>(w + ((size_t)t * cin + ci0 + kc) * npad + n);
      *reinterpret_cast<float4*>(&w_s[t][kc][c4 * 4]) = v;
    }
    __syncthreads();
    // ---- FMA
#pragma unroll
    for (int dyi = 0; dyi <= RH; ++dyi) {
#pragma unroll
      for (int kc = 0; kc < KC; ++kc) {
        float a[PPT + RH];
#pragma unroll
        for (int j = 0; j < PPT + RH; ++j) a[j] = in_s[row + dyi][kc][xs + j];
#pragma unroll
        for (int dxi = 0; dxi <= RH; ++dxi) {
          const float4 bw = *reinterpret_cast<const float4*>(&w_s[dyi * (RH + 1) + dxi][kc][tn * 4]);
#pragma unroll
          for (int j = 0; j < PPT; ++j) {
            acc[j][0] = fmaf(a[j + dxi], bw.x, acc[j][0]);
            acc[j][1] = fmaf(a[j + dxi], bw.y, acc[j][1]);
            acc[j][2] = fmaf(a[j + dxi], bw.z, acc[j][2]);
            acc[j][3] = fmaf(a[j + dxi], bw.w, acc[j][3]);
          }
        }
      }
    }
    __syncthreads();
  }

  // ---- epilogue
  const int n = n0 + tn * 4;
  if (n >= n_valid) return;
  const float4 bv = *reinterpret_cast<const float4*>(bias + n);
  float4 be = make_float4(1.f, 1.f, 1.f, 1.f);
  if (epi == EPI_RES || epi >= EPI_PRELU) be = *reinterpret_cast<const float4*>(beta + n);
  const int y = y0 + row;
  if (y >= H) return;
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    const int x = x0 + xs + j;
    if (x >= W) continue;
    float v[4] = {acc[j][0] + bv.x, acc[j][1] + bv.y, acc[j][2] + bv.z, acc[j][3] + bv.w};
    switch (epi) {
      case EPI_LRELU:
      case EPI_LINEAR: {
        if (epi == EPI_LRELU) {
#pragma unroll
          for (int q = 0; q < 4; ++q) v[q] = lrelu02(v[q]);
        }
        st4(out + (((size_t)b * H + y) * W + x) * out_c + n, v);
      } break;
      case EPI_LRELU_S2D: {
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = lrelu02(v[q]);
        const size_t o = ((((size_t)b * (H >> 1) + (y >> 1)) * (W >> 1) + (x >> 1)) * 4 + ((y & 1) * 2 + (x & 1)));
        st4(out + o * out_c + n, v);
      } break;
      case EPI_RES: {
        const size_t o = (((size_t)b * H + y) * W + x) * out_c + n;
        float r[4];
        ld4(res + (((size_t)b * H + y) * W + x) * res_pitch + n, r);
        v[0] = lrelu02(v[0] * be.x + r[0]);
        v[1] = lrelu02(v[1] * be.y + r[1]);
        v[2] = lrelu02(v[2] * be.z + r[2]);
        v[3] = lrelu02(v[3] * be.w + r[3]);
        st4(out + o, v);
      } break;
      case EPI_PRELU:
      case EPI_PRELU_ADD: {
        v[0] = v[0] >= 0.f ? v[0] : v[0] * be.x;
        v[1] = v[1] >= 0.f ? v[1] : v[1] * be.y;
        v[2] = v[2] >= 0.f ? v[2] : v[2] * be.z;
        v[3] = v[3] >= 0.f ? v[3] : v[3] * be.w;
        const size_t pix = ((size_t)b * H + y) * W + x;
        if (epi == EPI_PRELU_ADD) {
          float r[4];
          ld4(res + pix * res_pitch + n, r);
#pragma unroll
          for (int q = 0; q < 4; ++q) v[q] += r[q];
        }
        st4(out + pix * out_c + n, v);
      } break;
      case EPI_PS: {
        // ConvT output (2y+py, 2x+px), channel co = 4k + 2dy + dx  ->  PixelShuffle -> (4y+2py+dy, 4x+2px+dx), channel k
        const int par = n / cpp, co = n % cpp, k = co >> 2;
        const int py = par >> 1, px = par & 1;
        const int Ho = 4 * H, Wo = 4 * W;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int oy = 4 * y + 2 * py + (q >> 1), ox = 4 * x + 2 * px + (q & 1);
          out[(((size_t)b * Ho + oy) * Wo + ox) * TMP_C + k] = v[q];
        }
      } break;
      case EPI_HEADT:
      case EPI_CONVT:
      case EPI_CONVT_PRELU: {
        if (epi == EPI_CONVT_PRELU) {
          v[0] = v[0] >= 0.f ? v[0] : v[0] * be.x;
          v[1] = v[1] >= 0.f ? v[1] : v[1] * be.y;
          v[2] = v[2] >= 0.f ? v[2] : v[2] * be.z;
          v[3] = v[3] >= 0.f ? v[3] : v[3] * be.w;
        }
        const int par = n / cpp, co = n % cpp;
        const int oy = 2 * y + (par >> 1), ox = 2 * x + (par & 1);
        st4(out + (((size_t)b * 2 * H + oy) * 2 * W + ox) * out_c + co, v);
      } break;
      default: break;
    }
  }
}

template <int KIND, int NT>
int launch_t(const ConvLayerF32& L, const ConvIO& io, cudaStream_t st) {
  const int ntn = (L.npad + NT - 1) / NT;
  dim3 grid((io.W + TW - 1) / TW, (io.H + TH - 1) / TH, io.B * ntn);
  conv_f32_kernel<KIND, NT><<<grid, NTHREADS, 0, st>>>((const float*)io.in, L.w, L.bias, L.beta,
                                                       (const float*)io.res, (float*)io.out, io.B, io.H, io.W, L.cin,
                                                       L.n, L.npad, io.epi, io.out_c, io.co_per_par,
                                                       io.in_pitch ? io.in_pitch : L.cin,
                                                       io.res_pitch ? io.res_pitch : io.out_c);
  RIFE_LAUNCH_CHECK();
  return RIFE_OK;
}

}  // namespace

int launch_conv_f32(const ConvLayerF32& L, const ConvIO& io, cudaStream_t st) {
  if (L.cin % KC != 0 || L.npad % 4 != 0 || L.n % 4 != 0) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "conv_f32: channel padding");
  if (io.epi == EPI_RES && (!L.beta || !io.res)) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "conv_f32: EPI_RES needs beta and res");
  if (io.epi >= EPI_PRELU && !L.beta) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "conv_f32: PReLU epilogues need the slope vector");
  if (io.epi == EPI_PRELU_ADD && !io.res) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "conv_f32: EPI_PRELU_ADD needs res");
  const bool wide = L.npad > 32, mid = L.npad > 16;
  if (L.kind == CONV_3x3) {
    if (wide) return launch_t<CONV_3x3, 64>(L, io, st);
    if (mid) return launch_t<CONV_3x3, 32>(L, io, st);
    return launch_t<CONV_3x3, 16>(L, io, st);
  } else if (L.kind == CONV_S2D) {
    if (wide) return launch_t<CONV_S2D, 64>(L, io, st);
    if (mid) return launch_t<CONV_S2D, 32>(L, io, st);
    return launch_t<CONV_S2D, 16>(L, io, st);
  }
  RIFE_FAIL(RIFE_ERR_INVALID_ARG, "conv_f32: bad kind");
}

}  // namespace rife
