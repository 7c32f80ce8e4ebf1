// C ABI of librife_b200.so (include/rife_b200.h): weight repacking, plan/workspace, the IFNet pipeline.
// Host-side orchestration of train_log/IFNet_HDv3.py:119-169 and train_log/RIFE_HDv3.py:56-60.
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "common.cuh"
#include "conv_gather.cuh"
#include "conv_halo.cuh"
#include "conv_head.cuh"
#include "conv_strip.cuh"
#include "conv_umma.cuh"

namespace rife {

std::atomic<uint64_t> g_launch_count{0};
bool pdl_enabled() {
  static const bool on = !(getenv("RIFE_NO_PDL") && atoi(getenv("RIFE_NO_PDL")));
  return on;
}
static thread_local std::string t_error;
void set_error(const std::string& msg) { t_error = msg; }

static const int kBlockC[5] = {192, 128, 96, 64, 32};  // train_log/IFNet_HDv3.py:95-99
static const int kLastCo = 52;                          // ConvTranspose2d(c, 4*13, 4, 2, 1), IFNet_HDv3.py:74

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static bool is_device_ptr(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// --------------------------------------------------------------------------------------------------
// host-side repacking into the implicit-GEMM layouts [tap][cin_pad][npad]
// Conv2d OIHW (co, ci, 3, 3), stride 1
static HostLayer pack_conv3x3(const float* W, const float* b, int co, int ci, int cin_pad) {
  HostLayer L;
  L.kind = CONV_3x3; L.taps = 9; L.cin = cin_pad; L.n = round_up(co, 4); L.npad = round_up(co, 16);
  L.w.assign((size_t)9 * cin_pad * L.npad, 0.f);
  L.bias.assign(L.npad, 0.f);
  for (int o = 0; o < co; ++o) {
    if (b) L.bias[o] = b[o];
    for (int i = 0; i < ci; ++i)
      for (int ky = 0; ky < 3; ++ky)
        for (int kx = 0; kx < 3; ++kx)
          L.w[((size_t)(ky * 3 + kx) * cin_pad + i) * L.npad + o] = W[(((size_t)o * ci + i) * 3 + ky) * 3 + kx];
  }
  return L;
}

// Conv2d OIHW (co, ci, 3, 3), stride 2, expressed on the space-to-depth input whose channel index is
// par * cs + i  (par = py*2+px): taps (dy,dx) in {-1,0}^2 with ky = 2*dy + py + 1, kx = 2*dx + px + 1.
static HostLayer pack_conv3x3_s2d(const float* W, const float* b, int co, int ci, int cs, int cin_pad) {
  HostLayer L;
  L.kind = CONV_S2D; L.taps = 4; L.cin = cin_pad; L.n = round_up(co, 4); L.npad = round_up(co, 16);
  L.w.assign((size_t)4 * cin_pad * L.npad, 0.f);
  L.bias.assign(L.npad, 0.f);
  for (int o = 0; o < co; ++o) {
    if (b) L.bias[o] = b[o];
    for (int i = 0; i < ci; ++i)
      for (int dy = -1; dy <= 0; ++dy)
        for (int dx = -1; dx <= 0; ++dx)
          for (int py = 0; py < 2; ++py)
            for (int px = 0; px < 2; ++px) {
              const int ky = 2 * dy + py + 1, kx = 2 * dx + px + 1;
              if (ky < 0 || ky > 2 || kx < 0 || kx > 2) continue;
              const int t = (dy + 1) * 2 + (dx + 1);
              const int c = (py * 2 + px) * cs + i;
              L.w[((size_t)t * cin_pad + c) * L.npad + o] = W[(((size_t)o * ci + i) * 3 + ky) * 3 + kx];
            }
  }
  return L;
}

// ConvTranspose2d IOHW (ci, co, 4, 4), stride 2, pad 1 as a 3x3-tap GEMM with N = 4 parities x co:
// out[(2y+py, 2x+px), o] = sum_{dy,dx} in[(y+dy, x+dx)] * W[., o, py+1-2dy, px+1-2dx]   
static HostLayer pack_convT(const float* W, const float* b, int ci, int co, int cin_pad, int co_pad = 0) {
  HostLayer L;
  const int cp = co_pad ? co_pad : co;  // columns per output parity (>= co, zero-padded)
  L.kind = CONV_3x3; L.taps = 9; L.cin = cin_pad; L.n = 4 * cp; L.npad = round_up(4 * cp, 16);
  L.convt_co = cp;
  L.w.assign((size_t)9 * cin_pad * L.npad, 0.f);
  L.bias.assign(L.npad, 0.f);
  for (int par = 0; par < 4; ++par) {
    const int py = par >> 1, px = par & 1;
    for (int o = 0; o < co; ++o) {
      if (b) L.bias[par * cp + o] = b[o];
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          const int ky = py + 1 - 2 * dy, kx = px + 1 - 2 * dx;
          if (ky < 0 || ky > 3 || kx < 0 || kx > 3) continue;
          const int t = (dy + 1) * 3 + (dx + 1);
          for (int i = 0; i < ci; ++i)
            L.w[((size_t)t * cin_pad + i) * L.npad + par * cp + o] = W[(((size_t)i * co + o) * 4 + ky) * 4 + kx];
        }
    }
  }
  return L;
}

struct DevLayer {
  ConvLayerF32 f32;   // always present (fp32 FMA path)
  UmmaLayer h16;      // present when precision == FP16: NHWC tcgen05 kernel (conv_umma.cu)
  HaloLayer hk;       // present when precision == FP16 and the layer fits the halo kernel (conv_halo.cu)
  bool use_hk = false;
  // RIFE_PRECISION_BF16: operand type of this layer's input + weights, and storage type of its activation output
  bool in_bf16 = false, out_bf16 = false;
};

static inline bool is_half_mode(int precision) { return precision == RIFE_PRECISION_FP16 || precision == RIFE_PRECISION_BF16; }

}  // namespace rife

using namespace rife;

struct rife_weights {
  int device = 0;
  int precision = RIFE_PRECISION_FP32;
  bool force_nhwc = false;  // test hook: never pick the halo kernel
  bool head_hk = false;     // Head runs on the halo kernel with planar-8 activations
  // per block: 0 = NHWC tcgen05 kernel everywhere; 1 = halo kernel everywhere (planar-8 activations);
  // 2 = conv0.* on the NHWC kernel (conv0.1 stores planar-8), ResConv x8 + lastconv on the halo kernel (streamed weights)
  int blk_hk[5] = {0, 0, 0, 0, 0};
  std::vector<void*> allocs;
  // conv_strip.cu: Head.cnn1 + cnn2 as one launch, and the ResConv layers (2j, 2j + 1) of IFBlock 4 as one launch each
  StripPair strip_head;
  bool has_strip_head = false;
  StripPair strip_res4[4];
  bool has_strip_res4 = false;
  DevLayer head[4];
  HeadFused head_fused;     // the four Head layers as one kernel (conv_head.cu); 16-bit modes
  bool has_head_fused = false;
  struct Block {
    DevLayer conv00, conv01, res[8], last;
    HaloLayer last8;       // block 4 only, fp16 halo path: lastconv restricted to flow + mask (8-channel tmp pixel)
    bool has_last8 = false;
    int c;
  } blk[5];
};

static int upload(rife_weights* w, const void* host, size_t bytes, void** dev) {
  RIFE_CUDA_TRY(cudaMalloc(dev, bytes));
  w->allocs.push_back(*dev);
  RIFE_CUDA_TRY(cudaMemcpy(*dev, host, bytes, cudaMemcpyHostToDevice));
  return RIFE_OK;
}

// in_bf / out_bf (RIFE_PRECISION_BF16 only): see DevLayer
static int upload_layer(rife_weights* w, const HostLayer& h, DevLayer* d, bool ps_layout = false, bool in_bf = false,
                        bool out_bf = false) {
  void *pw, *pb, *pbeta = nullptr;
  RIFE_TRY(upload(w, h.w.data(), h.w.size() * 4, &pw));
  RIFE_TRY(upload(w, h.bias.data(), h.bias.size() * 4, &pb));
  if (!h.beta.empty()) RIFE_TRY(upload(w, h.beta.data(), h.beta.size() * 4, &pbeta));
  d->f32.w = (const float*)pw;
  d->f32.bias = (const float*)pb;
  d->f32.beta = (const float*)pbeta;
  d->f32.kind = h.kind; d->f32.cin = h.cin; d->f32.n = h.n; d->f32.npad = h.npad;
  d->in_bf16 = in_bf; d->out_bf16 = out_bf;
  if (is_half_mode(w->precision)) {
    void *ph = nullptr, *phb = nullptr, *phbeta = nullptr;
    std::vector<__half> packed;
    std::vector<float> hb, hbeta;
    RIFE_TRY(umma_pack_weights(h, ps_layout, &packed, &hb, &hbeta, &d->h16, in_bf));
    RIFE_TRY(upload(w, packed.data(), packed.size() * sizeof(__half), &ph));
    RIFE_TRY(upload(w, hb.data(), hb.size() * 4, &phb));
    if (!hbeta.empty()) RIFE_TRY(upload(w, hbeta.data(), hbeta.size() * 4, &phbeta));
    d->h16.w = (const __half*)ph;
    d->h16.bias = (const float*)phb;
    d->h16.beta = (const float*)phbeta;
    if (!w->force_nhwc && halo_fits(h, ps_layout)) {
      void *pk = nullptr, *pkb = nullptr, *pkbeta = nullptr;
      RIFE_TRY(halo_pack_weights(h, ps_layout, &packed, &hb, &hbeta, &d->hk, in_bf));
      RIFE_TRY(upload(w, packed.data(), packed.size() * sizeof(__half), &pk));
      RIFE_TRY(upload(w, hb.data(), hb.size() * 4, &pkb));
      if (!hbeta.empty()) RIFE_TRY(upload(w, hbeta.data(), hbeta.size() * 4, &pkbeta));
      d->hk.w = (const __half*)pk;
      d->hk.bias = (const float*)pkb;
      d->hk.beta = (const float*)pkbeta;
      d->use_hk = true;
    }
  }
  return RIFE_OK;
}

static bool strip_enabled() {
  // RIFE_STRIP=1: Head.cnn1+cnn2 and the ResConv pairs of IFBlock 4 run on the fused kx-stacked strip kernel
  // (conv_strip.cu).  Off by default: measured on B200 (16 x 1080p pairs) it is slower than the per-layer halo kernel --
  // b4.res 0.78 vs 0.58 ms, head.cnn1+2 0.58 vs 0.42 ms -- because a tcgen05.mma's operand reads already saturate the
  // SM's shared-memory bandwidth (N = 96: 7 KB per 56 cycles) and the epilogue's residual loads, intermediate stores and
  // shuffles compete for it; see DESIGN.md section 5.
  static const bool on = getenv("RIFE_STRIP") && atoi(getenv("RIFE_STRIP")) != 0;
  return on;
}

// two consecutive 3x3 c -> c layers (c = 16 / 32) packed for the fused strip kernel; biases / betas are those already
// uploaded for the per-layer kernels
static int upload_strip_pair(rife_weights* w, const HostLayer& a, const HostLayer& b, const DevLayer& da, const DevLayer& db,
                             bool res, bool bf, StripPair* out) {
  std::vector<__half> pa, pb;
  RIFE_TRY(strip_pack_weights(a, &pa, bf));
  RIFE_TRY(strip_pack_weights(b, &pb, bf));
  pa.insert(pa.end(), pb.begin(), pb.end());
  void* pw = nullptr;
  RIFE_TRY(upload(w, pa.data(), pa.size() * sizeof(__half), &pw));
  out->w = (const __half*)pw;
  out->bias[0] = da.f32.bias; out->bias[1] = db.f32.bias;
  out->beta[0] = res ? da.f32.beta : nullptr; out->beta[1] = res ? db.f32.beta : nullptr;
  out->res[0] = out->res[1] = res ? 1 : 0;
  out->c = a.cin;
  return RIFE_OK;
}

extern "C" {

int rife_abi_version(void) { return RIFE_B200_ABI_VERSION; }
const char* rife_last_error(void) { return t_error.c_str(); }
uint64_t rife_launch_count(void) { return g_launch_count.load(); }

int rife_weights_create(rife_weights** out, int n, const char* const* names, const float* const* data,
                        const int64_t* numel, int precision) {
  if (!out || !names || !data || !numel) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "rife_weights_create: null argument");
  if (precision != RIFE_PRECISION_FP32 && !is_half_mode(precision))
    RIFE_FAIL(RIFE_ERR_DTYPE, "rife_weights_create: precision must be RIFE_PRECISION_FP32, RIFE_PRECISION_FP16 or RIFE_PRECISION_BF16");
  // bf16 mode: every conv whose input is itself a conv activation runs bf16 x bf16 and every conv activation is stored
  // as bf16; the layers that read a pixel-kernel product (Head.cnn0: the s2d frame, conv0.0: the block input) keep fp16
  // operands, and the tensors the pixel kernels read (SRC, tmp) stay fp16 -- see DESIGN.md section 2
  const bool bf = precision == RIFE_PRECISION_BF16;
  std::map<std::string, std::pair<const float*, int64_t>> sd;
  for (int i = 0; i < n; ++i) sd[names[i]] = {data[i], numel[i]};
  std::string missing;
  auto get = [&](const std::string& k, int64_t want) -> const float* {
    auto it = sd.find(k);
    if (it == sd.end() || it->second.second != want) {
      if (missing.empty())
        missing = k + (it == sd.end() ? " (absent)" : " (numel " + std::to_string(it->second.second) + ", want " +
                                                           std::to_string(want) + ")");
      return nullptr;
    }
    return it->second.first;
  };
  auto* w = new rife_weights();
  w->precision = precision;
  cudaGetDevice(&w->device);
  int st = RIFE_OK;
  auto fail = [&](int code) {
    rife_weights_destroy(w);
    return code;
  };
  // ---- Head (train_log/IFNet_HDv3.py:27-30)
  {
    const float *w0 = get("encode.cnn0.weight", 16 * 3 * 9), *b0 = get("encode.cnn0.bias", 16);
    const float *w1 = get("encode.cnn1.weight", 16 * 16 * 9), *b1 = get("encode.cnn1.bias", 16);
    const float *w2 = get("encode.cnn2.weight", 16 * 16 * 9), *b2 = get("encode.cnn2.bias", 16);
    const float *w3 = get("encode.cnn3.weight", 16 * 4 * 16), *b3 = get("encode.cnn3.bias", 4);
    if (!missing.empty()) { set_error("rife_weights_create: bad tensor " + missing); return fail(RIFE_ERR_WEIGHTS); }
    if ((st = upload_layer(w, pack_conv3x3_s2d(w0, b0, 16, 3, 3, 16), &w->head[0], false, false, bf))) return fail(st);
    if ((st = upload_layer(w, pack_conv3x3(w1, b1, 16, 16, 16), &w->head[1], false, bf, bf))) return fail(st);
    if ((st = upload_layer(w, pack_conv3x3(w2, b2, 16, 16, 16), &w->head[2], false, bf, bf))) return fail(st);
    if ((st = upload_layer(w, pack_convT(w3, b3, 16, 4, 16), &w->head[3], false, bf, false))) return fail(st);
    if (is_half_mode(precision) && !w->force_nhwc) {
      const HostLayer hh[4] = {pack_conv3x3_s2d(w0, b0, 16, 3, 3, 16), pack_conv3x3(w1, b1, 16, 16, 16),
                               pack_conv3x3(w2, b2, 16, 16, 16), pack_convT(w3, b3, 16, 4, 16)};
      if (head_fused_fits(hh)) {
        std::vector<__half> packed;
        std::vector<float> hb;
        void *pw = nullptr, *pb = nullptr;
        if ((st = head_fused_pack(hh, bf, &packed, &hb))) return fail(st);
        if ((st = upload(w, packed.data(), packed.size() * sizeof(__half), &pw))) return fail(st);
        if ((st = upload(w, hb.data(), hb.size() * 4, &pb))) return fail(st);
        w->head_fused.w = (const __half*)pw;
        w->head_fused.bias = (const float*)pb;
        w->head_fused.bf16 = bf ? 1 : 0;
        w->head_fused.w_bytes = (int)(packed.size() * sizeof(__half));
        w->has_head_fused = true;
      }
    }
    if (is_half_mode(precision) && strip_enabled() && w->head[1].use_hk && w->head[2].use_hk) {
      const HostLayer h1 = pack_conv3x3(w1, b1, 16, 16, 16), h2 = pack_conv3x3(w2, b2, 16, 16, 16);
      if (strip_fits(h1) && strip_fits(h2)) {
        if ((st = upload_strip_pair(w, h1, h2, w->head[1], w->head[2], false, bf, &w->strip_head))) return fail(st);
        w->has_strip_head = true;
      }
    }
  }
  // ---- IFBlocks (train_log/IFNet_HDv3.py:57-76, 95-99)
  for (int i = 0; i < 5; ++i) {
    const int c = kBlockC[i], p = (i == 0) ? 15 : 28;
    const std::string pre = "block" + std::to_string(i) + ".";
    auto& B = w->blk[i];
    B.c = c;
    const float *w00 = get(pre + "conv0.0.0.weight", (int64_t)(c / 2) * p * 9), *b00 = get(pre + "conv0.0.0.bias", c / 2);
    const float *w01 = get(pre + "conv0.1.0.weight", (int64_t)c * (c / 2) * 9), *b01 = get(pre + "conv0.1.0.bias", c);
    const float *wl = get(pre + "lastconv.0.weight", (int64_t)c * kLastCo * 16), *bl = get(pre + "lastconv.0.bias", kLastCo);
    if (!missing.empty()) { set_error("rife_weights_create: bad tensor " + missing); return fail(RIFE_ERR_WEIGHTS); }
    if ((st = upload_layer(w, pack_conv3x3_s2d(w00, b00, c / 2, p, XIN_C, 4 * XIN_C), &B.conv00, false, false, bf))) return fail(st);
    if ((st = upload_layer(w, pack_conv3x3_s2d(w01, b01, c, c / 2, c / 2, 2 * c), &B.conv01, false, bf, bf))) return fail(st);
    for (int j = 0; j < 8; ++j) {
      const std::string rp = pre + "convblock." + std::to_string(j) + ".";
      const float *wr = get(rp + "conv.weight", (int64_t)c * c * 9), *br = get(rp + "conv.bias", c);
      const float* be = get(rp + "beta", c);
      if (!missing.empty()) { set_error("rife_weights_create: bad tensor " + missing); return fail(RIFE_ERR_WEIGHTS); }
      HostLayer L = pack_conv3x3(wr, br, c, c, c);
      L.beta.assign(L.npad, 0.f);
      for (int k = 0; k < c; ++k) L.beta[k] = be[k];
      if ((st = upload_layer(w, L, &B.res[j], false, bf, bf))) return fail(st);
    }
    if (i == 4 && is_half_mode(precision) && strip_enabled() && B.res[0].use_hk) {
      bool ok = true;
      for (int j = 0; j < 8 && ok; j += 2) {
        HostLayer hp[2];
        for (int k = 0; k < 2; ++k) {
          const std::string rp = pre + "convblock." + std::to_string(j + k) + ".";
          hp[k] = pack_conv3x3(get(rp + "conv.weight", (int64_t)c * c * 9), get(rp + "conv.bias", c), c, c, c);
        }
        ok = strip_fits(hp[0]) && strip_fits(hp[1]);
        if (ok && (st = upload_strip_pair(w, hp[0], hp[1], B.res[j], B.res[j + 1], true, bf, &w->strip_res4[j / 2]))) return fail(st);
      }
      w->has_strip_res4 = ok;
    }
    const HostLayer hl = pack_convT(wl, bl, c, kLastCo, c);
    if ((st = upload_layer(w, hl, &B.last, /*ps_layout=*/true, bf, false))) return fail(st);
    if (i == 4 && B.last.use_hk && halo_fits(hl, 2)) {
      // nothing downstream reads the last block's feat channels (IFNet_HDv3.py:148-159): Model.inference runs a lastconv
      // with 5 of the 13 output channels
      std::vector<__half> packed;
      std::vector<float> hb, hbeta;
      void *pk = nullptr, *pkb = nullptr;
      if ((st = halo_pack_weights(hl, 2, &packed, &hb, &hbeta, &B.last8, bf))) return fail(st);
      if ((st = upload(w, packed.data(), packed.size() * sizeof(__half), &pk))) return fail(st);
      if ((st = upload(w, hb.data(), hb.size() * 4, &pkb))) return fail(st);
      B.last8.w = (const __half*)pk;
      B.last8.bias = (const float*)pkb;
      B.has_last8 = true;
    }
    bool tail = B.last.use_hk;
    for (int j = 0; j < 8; ++j) tail = tail && B.res[j].use_hk;
    w->blk_hk[i] = (tail && B.conv00.use_hk && B.conv01.use_hk) ? 1 : (tail ? 2 : 0);
  }
  w->head_hk = w->head[0].use_hk && w->head[1].use_hk && w->head[2].use_hk && w->head[3].use_hk;
  *out = w;
  return RIFE_OK;
}

int rife_weights_destroy(rife_weights* w) {
  if (!w) return RIFE_OK;
  for (void* p : w->allocs) cudaFree(p);
  delete w;
  return RIFE_OK;
}

}  // extern "C"

// --------------------------------------------------------------------------------------------------
// plan
constexpr int kTileCtrs = 1024;

struct PlanLayout {
  size_t src, head_in, head_a, head_b, xin, xfull, c0, act_a, act_b, tmp[5], fbuf, fbuf2, gridx, gridy, tstep, sync, ctrs, io_tab, stage_in0, stage_in1,
      stage_out, total;
  int s[5];   // block i works at H / s[i] * u[i]: scale_list[i] = s[i] / u[i] (one of the two is 1)
  int u[5];
  bool generic = false;   // scale_list is not the halving pyramid of RIFE_HDv3.py:58: generic pixel kernels only
};

// P frame pairs, T timesteps per pair (outputs B = P*T, pair-major), `seq`: the pairs are the P consecutive pairs of
// ONE run of P+1 frames (Head features computed once per frame instead of twice)
// `sl`: the five entries of IFNet.forward's scale_list (IFNet_HDv3.py:119,139-148), each a power of two in [1/4, 64]
static int plan_layout(int P, int T, int seq, int C, int H, int W, const float* sl5, int precision, PlanLayout* L) {
  if (P < 1 || T < 1 || H < 1 || W < 1) RIFE_FAIL(RIFE_ERR_SHAPE, "plan: pairs, timesteps, H, W must be positive");
  const int B = P * T, F = seq ? P + 1 : 2 * P;
  if (C != 3 && C != 4) RIFE_FAIL(RIFE_ERR_SHAPE, "plan: C must be 3 or 4 (IFNet_HDv3.py:121-123 splits an RGB/RGBA pair)");
  if (precision != RIFE_PRECISION_FP32 && !is_half_mode(precision)) RIFE_FAIL(RIFE_ERR_DTYPE, "plan: bad precision");
  bool any_up = false;
  L->generic = false;
  for (int i = 0; i < 5; ++i) {
    const float sl = sl5[i];
    int e = 0;
    const float m = frexpf(sl, &e);   // sl = m * 2^e, m = 0.5 for a power of two
    if (!(sl >= 0.25f && sl <= 64.f) || m != 0.5f)
      RIFE_FAIL(RIFE_ERR_SHAPE, "plan: every scale_list entry must be a power of two in [1/4, 64] (inference_video.py:92 with "
                                "scale_list = [16, 8, 4, 2, 1] / scale gives exactly those); got " + std::to_string(sl));
    L->s[i] = sl >= 1.f ? (int)lroundf(sl) : 1;
    L->u[i] = sl >= 1.f ? 1 : (int)lroundf(1.f / sl);   // entry < 1: the block works above full resolution
    any_up = any_up || L->u[i] > 1;
    if (i > 0 && sl5[i - 1] != 2.f * sl) L->generic = true;
    // block i works on (H / s * u, W / s * u) and halves it twice more (conv0): that size must be a multiple of 4
    if (H % L->s[i] || W % L->s[i] || (H / L->s[i] * L->u[i]) % 4 || (W / L->s[i] * L->u[i]) % 4)
      RIFE_FAIL(RIFE_ERR_SHAPE, "plan: H and W must be multiples of 64/scale (in general of 4 * scale_list[i] for every "
                                "block, IFNet_HDv3.py:79-81 + two stride-2 convs); got " + std::to_string(H) + "x" + std::to_string(W) +
                                " with scale_list[" + std::to_string(i) + "] = " + std::to_string(sl));
  }
  const size_t E = is_half_mode(precision) ? 2 : 4;
  const size_t N = (size_t)B * H * W, NF = (size_t)F * H * W;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 1024);
    return o;
  };
  L->src = take(NF * SRC_C * E);
  L->head_in = take((NF / 4) * 16 * E);
  L->head_a = take((NF / 4) * 16 * E);
  L->head_b = take((NF / 4) * 16 * E);
  size_t xin = 0, c0 = 0, act = 0;
  for (int i = 0; i < 5; ++i) {
    const size_t n = N / ((size_t)L->s[i] * L->s[i]) * L->u[i] * L->u[i];
    const size_t c = kBlockC[i];
    xin = xin > n * XIN_C ? xin : n * XIN_C;
    c0 = c0 > (n / 4) * (c / 2) ? c0 : (n / 4) * (c / 2);
    act = act > (n / 16) * c ? act : (n / 16) * c;
  }
  L->xin = take(xin * E);
  L->xfull = take(any_up ? N * XIN_C * E : 0);   // full-resolution block input before its up-sampling (scale > 1 only)
  L->c0 = take(c0 * E);
  L->act_a = take(act * E);
  L->act_b = take(act * E);
  for (int i = 0; i < 5; ++i) L->tmp[i] = take(N / ((size_t)L->s[i] * L->s[i]) * L->u[i] * L->u[i] * TMP_C * E);
  L->fbuf = take(N * 4 * 4);  // full-resolution fp32 flow carried between the last stages
  // second flow buffer: the gather fused into the scale-1 block's conv0.0 (conv_gather.cu) cannot update fbuf in place
  L->fbuf2 = take(is_half_mode(precision) && !any_up && !L->generic && L->s[4] == 1 ? N * 4 * 4 : 0);
  L->gridx = take((size_t)W * 4);
  L->gridy = take((size_t)H * 4);
  L->tstep = take((size_t)B * 4);
  L->sync = take(256);   // grid-barrier counter of the ResConv chain kernel
  L->ctrs = take(kTileCtrs * 4);   // one dynamic-tile-scheduling counter per halo-kernel launch of a call
  L->io_tab = take(256);           // {img0, img1, merged} of the current call, read by a replayed CUDA graph
  L->stage_in0 = take((size_t)(seq ? P + 1 : P) * H * W * C * 4);
  L->stage_in1 = take((size_t)(seq ? 0 : P) * H * W * C * 4);
  L->stage_out = take(N * C * 4);
  L->total = off;
  return RIFE_OK;
}

struct rife_plan {
  const rife_weights* w = nullptr;
  int B = 0, C = 0, H = 0, W = 0, precision = 0, device = 0;   // B = P * T outputs per call
  int P = 0, T = 1, seq = 0;                                    // frame pairs, timesteps per pair, sequence mode
  int nf0() const { return seq ? P + 1 : P; }                   // frames the caller passes in img0 / img1
  int nf1() const { return seq ? 0 : P; }
  int fh = 0, fw = 0;  // frame size seen by the caller (rife_plan_set_frame_size); 0 = H, W
  int chunk = 0;       // pairs per pass of the full-resolution stages (rife_plan_set_chunk); 0 = default
  unsigned int sync_host = 0;  // value of the ResConv chain's grid-barrier counter after everything enqueued so far
  int next_ctr = 0;            // tile-scheduling counters handed out to the halo launches of the current call
  std::vector<float> ts_host;  // timesteps last written to the device buffer, and the stream they were written on
  cudaStream_t ts_stream = nullptr;
  bool ts_valid = false;
  PlanLayout L{};
  char* ws = nullptr;
  void* at(size_t off) const { return ws + off; }
  // optional per-stage CUDA-event timing (rife_plan_set_profiling)
  bool profile = false;
  struct Stage {
    std::string name;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
  };
  std::vector<Stage> stages;
  size_t n_stage = 0;  // stages recorded by the last run
  // rife_inference_host_pipelined: "staging-in filled", "compute done", "staging-out drained"
  cudaEvent_t ev_in = nullptr, ev_compute = nullptr, ev_out = nullptr;
  // CUDA-graph replay of rife_inference for small batches: the ~66 dependent launches of one
  // call are captured once per (io dtype, frame size) and replayed; the frame pointers travel through L.io_tab
  cudaGraphExec_t gexec = nullptr;
  cudaStream_t cap_stream = nullptr;   // capture happens on a stream of the plan's own: the caller's may be the legacy default stream
  int g_io = -1, g_fh = 0, g_fw = 0, g_chunk = 0, g_launches = 0;
  int graph_state = 0;   // 0 = not captured yet, 1 = captured, -1 = capture failed once: plain launches from then on
  int plain_calls = 0;   // the first call always runs un-captured (one-time function attributes, lazy driver state)
  ~rife_plan() {
    if (gexec) cudaGraphExecDestroy(gexec);
    if (cap_stream) cudaStreamDestroy(cap_stream);
    for (cudaEvent_t e : {ev_in, ev_compute, ev_out})
      if (e) cudaEventDestroy(e);
    for (auto& s : stages) {
      if (s.e0) cudaEventDestroy(s.e0);
      if (s.e1) cudaEventDestroy(s.e1);
    }
  }
};

// brackets one launch with events when profiling is on
static bool nvtx_on() {
  // RIFE_NVTX=1: one NVTX range per pipeline stage (same names as rife_plan_get_profile) around its launches, for
  // ncu --nvtx / Nsight timelines; header-only NVTX v3, nothing happens without an attached tool
  static const bool on = getenv("RIFE_NVTX") && atoi(getenv("RIFE_NVTX")) != 0;
  return on;
}

struct StageTimer {
  rife_plan* p;
  cudaStream_t st;
  rife_plan::Stage* s = nullptr;
  bool range = false;
  StageTimer(rife_plan* p_, cudaStream_t st_, const std::string& name) : p(p_), st(st_) {
    if (nvtx_on()) {
      nvtxRangePushA(name.c_str());
      range = true;
    }
    if (!p->profile) return;
    if (p->n_stage == p->stages.size()) {
      p->stages.emplace_back();
      cudaEventCreate(&p->stages.back().e0);
      cudaEventCreate(&p->stages.back().e1);
    }
    s = &p->stages[p->n_stage++];
    s->name = name;
    cudaEventRecord(s->e0, st);
  }
  ~StageTimer() {
    if (s) cudaEventRecord(s->e1, st);
    if (range) nvtxRangePop();
  }
};
#define RIFE_STAGE(name, expr)          \
  do {                                  \
    StageTimer _t(p, st, name);         \
    RIFE_TRY(expr);                     \
  } while (0)

extern "C" {

size_t rife_plan_workspace_bytes(int B, int C, int H, int W, float scale, int precision) {
  return rife_plan_workspace_bytes_ex(B, 1, 0, C, H, W, scale, precision);
}

static int scale_to_list(float scale, float* sl5) {
  if (!(scale == 1.f || scale == 0.5f || scale == 0.25f || scale == 2.f || scale == 4.f))
    RIFE_FAIL(RIFE_ERR_SHAPE, "plan: scale must be one of 0.25, 0.5, 1, 2, 4 (inference_video.py:92)");
  const float base[5] = {16.f, 8.f, 4.f, 2.f, 1.f};
  for (int i = 0; i < 5; ++i) sl5[i] = base[i] / scale;   // scale_list = [16, 8, 4, 2, 1] / scale, RIFE_HDv3.py:58
  return RIFE_OK;
}

size_t rife_plan_workspace_bytes_ex(int pairs, int timesteps, int sequence, int C, int H, int W, float scale, int precision) {
  float sl[5];
  if (scale_to_list(scale, sl) != RIFE_OK) return 0;
  return rife_plan_workspace_bytes_list(pairs, timesteps, sequence, C, H, W, sl, precision);
}

size_t rife_plan_workspace_bytes_list(int pairs, int timesteps, int sequence, int C, int H, int W, const float* scale_list5,
                                      int precision) {
  PlanLayout L;
  if (!scale_list5) return 0;
  if (plan_layout(pairs, timesteps, sequence, C, H, W, scale_list5, precision, &L) != RIFE_OK) return 0;
  return L.total;
}

int rife_plan_create(rife_plan** out, const rife_weights* w, int B, int C, int H, int W, float scale, void* workspace,
                     size_t workspace_bytes, const float* gridx_host, const float* gridy_host) {
  return rife_plan_create_ex(out, w, B, 1, 0, C, H, W, scale, workspace, workspace_bytes, gridx_host, gridy_host);
}

int rife_plan_create_ex(rife_plan** out, const rife_weights* w, int pairs, int timesteps, int sequence, int C, int H, int W,
                        float scale, void* workspace, size_t workspace_bytes, const float* gridx_host,
                        const float* gridy_host) {
  float sl[5];
  RIFE_TRY(scale_to_list(scale, sl));
  return rife_plan_create_list(out, w, pairs, timesteps, sequence, C, H, W, sl, workspace, workspace_bytes, gridx_host, gridy_host);
}

int rife_plan_create_list(rife_plan** out, const rife_weights* w, int pairs, int timesteps, int sequence, int C, int H, int W,
                          const float* scale_list5, void* workspace, size_t workspace_bytes, const float* gridx_host,
                          const float* gridy_host) {
  if (!out || !w || !workspace || !scale_list5) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "rife_plan_create: null argument");
  PlanLayout L;
  RIFE_TRY(plan_layout(pairs, timesteps, sequence ? 1 : 0, C, H, W, scale_list5, w->precision, &L));
  const int B = pairs * timesteps;
  if (workspace_bytes < L.total)
    RIFE_FAIL(RIFE_ERR_WORKSPACE, "rife_plan_create: workspace has " + std::to_string(workspace_bytes) + " bytes, need " +
                                      std::to_string(L.total));
  if (((uintptr_t)workspace) % 256) RIFE_FAIL(RIFE_ERR_WORKSPACE, "rife_plan_create: workspace must be 256-byte aligned");
  if (!is_device_ptr(workspace)) RIFE_FAIL(RIFE_ERR_NOT_DEVICE_PTR, "rife_plan_create: workspace is not device memory");
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev != w->device) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "rife_plan_create: weights live on another device");
  auto* p = new rife_plan();
  p->w = w; p->B = B; p->C = C; p->H = H; p->W = W; p->precision = w->precision; p->device = dev;
  p->P = pairs; p->T = timesteps; p->seq = sequence ? 1 : 0;
  p->L = L;
  // 1024-byte alignment for TMA/UMMA-facing buffers: offsets are 1024-aligned, so align the base too
  p->ws = (char*)workspace;
  cudaError_t e = cudaMemset(p->ws, 0, L.total);
  if (e == cudaSuccess) {
    if (gridx_host) e = cudaMemcpy(p->at(L.gridx), gridx_host, (size_t)W * 4, cudaMemcpyHostToDevice);
    else if (launch_fill_grid((float*)p->at(L.gridx), W, 0) != RIFE_OK) e = cudaErrorUnknown;
  }
  if (e == cudaSuccess) {
    if (gridy_host) e = cudaMemcpy(p->at(L.gridy), gridy_host, (size_t)H * 4, cudaMemcpyHostToDevice);
    else if (launch_fill_grid((float*)p->at(L.gridy), H, 0) != RIFE_OK) e = cudaErrorUnknown;
  }
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    set_error(std::string("rife_plan_create: ") + cudaGetErrorString(e));
    delete p;
    return RIFE_ERR_CUDA;
  }
  *out = p;
  return RIFE_OK;
}

int rife_plan_destroy(rife_plan* p) {
  delete p;
  return RIFE_OK;
}

int rife_plan_set_frame_size(rife_plan* p, int h, int w) {
  if (!p) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "rife_plan_set_frame_size: null plan");
  if (h < 1 || w < 1 || h > p->H || w > p->W)
    RIFE_FAIL(RIFE_ERR_SHAPE, "rife_plan_set_frame_size: (h, w) must lie within the plan's padded size");
  p->fh = h; p->fw = w;
  return RIFE_OK;
}

int rife_plan_set_chunk(rife_plan* p, int pairs) {
  if (!p) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "rife_plan_set_chunk: null plan");
  if (pairs < 0) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "rife_plan_set_chunk: pairs must be >= 0 (0 = default)");
  p->chunk = pairs;
  return RIFE_OK;
}

int rife_plan_set_profiling(rife_plan* p, int enable) {
  if (!p) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "rife_plan_set_profiling: null plan");
  p->profile = enable != 0;
  p->n_stage = 0;
  return RIFE_OK;
}

int rife_plan_get_profile(rife_plan* p, const char** names, float* ms, int cap) {
  if (!p || !names || !ms) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "rife_plan_get_profile: null argument");
  int n = 0;
  for (size_t i = 0; i < p->n_stage && n < cap; ++i, ++n) {
    names[n] = p->stages[i].name.c_str();
    if (cudaEventElapsedTime(&ms[n], p->stages[i].e0, p->stages[i].e1) != cudaSuccess) {
      cudaGetLastError();
      ms[n] = -1.f;
    }
  }
  return n;
}

}  // extern "C"

// --------------------------------------------------------------------------------------------------
// pipeline
static bool narrow_off() {
  // test hooks: 16-channel tmp for the last block too / generic eval kernel instead of the fast merge
  static const bool off = getenv("RIFE_NO_NARROW_LAST") != nullptr || getenv("RIFE_NO_FAST_MERGE") != nullptr;
  return off;
}

static int chain_min_c(const rife_plan* p) {
  // RIFE_HK_CHAIN=c: the 8 ResConv layers of every block with >= c channels run as ONE cooperative launch with grid
  // barriers (conv_halo.cu).  Off by default, at every batch size -- measured on B200: 16 x 1080p pairs 2542 vs 2611
  // frames/s; one 1080p pair 0.895 vs 0.877 ms, one 448x256 pair 0.54 vs 0.50 ms (programmatic dependent launch already
  // hides most of a kernel boundary, a cooperative grid has to wait for the previous kernel to drain completely, and a
  // grid-wide barrier costs about what a boundary does).  The barrier re-arms itself, so the chain is CUDA-graph safe.
  (void)p;
  static const int env = getenv("RIFE_HK_CHAIN") ? atoi(getenv("RIFE_HK_CHAIN")) : 0;
  return env > 0 ? env : 1 << 30;
}

static int default_chunk_pairs() {
  // pairs per pass of the 1/2- and full-resolution stages; measured on B200 at 1080p
  static const int v = getenv("RIFE_CHUNK_PAIRS") ? atoi(getenv("RIFE_CHUNK_PAIRS")) : 0;
  return v;
}

static bool head_fused_on() {
  // RIFE_HEAD_FUSED=1: the four Head layers as ONE kernel with their activations in shared memory (conv_head.cu).  Off by
  // default -- measured on B200 (16 x 1080p pairs): 1.10 ms against 0.97 ms for the four per-layer launches.  It moves 3.7x
  // less HBM traffic and needs half the MMAs (row-stacked N = 64), but one CTA per SM then carries MMA operand reads (48
  // L1-data-pipe wavefronts per MMA: 43 % of the pipe), the epilogues' shared-memory stores and the TMEM reads of four layers
  // on the same SM: tensor pipe 29 % active, epilogue warps 80 % busy (profiles/r02_ncu_head_fused.md).
  static const bool on = getenv("RIFE_HEAD_FUSED") && atoi(getenv("RIFE_HEAD_FUSED")) != 0;
  return on;
}

static bool gather_conv_on() {
  // RIFE_GATHER_CONV=0: the scale-1 block's input gather and its conv0.0 as two kernels with `xin` in HBM between them
  static const bool on = !(getenv("RIFE_GATHER_CONV") && atoi(getenv("RIFE_GATHER_CONV")) == 0);
  return on;
}

static bool dyn_tiles_on() {
  static const bool on = !(getenv("RIFE_HK_DYN") && atoi(getenv("RIFE_HK_DYN")) == 0);  // A/B: static round robin
  return on;
}

// halo-kernel launches take the call's next tile-scheduling counter (zeroed by the call's first kernel)
static int run_conv(rife_plan* p, const DevLayer& L, const ConvIO& io_, bool hk, cudaStream_t st) {
  if (is_half_mode(p->precision)) {
    ConvIO io = io_;
    io.in_bf16 = L.in_bf16 ? 1 : 0; io.out_bf16 = L.out_bf16 ? 1 : 0;
    if (!hk) return launch_conv_umma(L.h16, io, st);
    ConvIO h = io;
    if (dyn_tiles_on() && p->next_ctr < kTileCtrs) h.tile_ctr = (unsigned int*)p->at(p->L.ctrs) + p->next_ctr++;
    return launch_conv_halo(L.hk, h, st);
  }
  return launch_conv_f32(L.f32, io_, st);
}

// timesteps: rewritten only when they (or the stream) changed since the last call on this plan, and then through kernel
// parameters -- a cudaMemcpyAsync here would sit on the H2D copy engine behind the next calls' frame uploads
// (rife_inference_host_pipelined) and stall the compute stream for up to a whole frame-batch transfer
static int update_tstep(rife_plan* p, const float* timestep_host, cudaStream_t st) {
  const int B = p->B;
  if (!(p->ts_valid && p->ts_stream == st && (int)p->ts_host.size() == B &&
        memcmp(p->ts_host.data(), timestep_host, (size_t)B * 4) == 0)) {
    RIFE_TRY(launch_set_tstep((float*)p->at(p->L.tstep), timestep_host, B, st));
    p->ts_host.assign(timestep_host, timestep_host + B);
    p->ts_stream = st;
    p->ts_valid = true;
  }
  return RIFE_OK;
}

// graph_capture: the launches are being captured into the plan's CUDA graph -- the first / last kernel take the frame
// pointers from the plan's table and the timesteps are already on the device
static int run_pipeline(rife_plan* p, const void* img0, const void* img1, const float* timestep_host,
                        const rife_ifnet_outputs* outs, int io_dtype, cudaStream_t st, bool graph_capture = false) {
  if (!p || !img0 || (!img1 && !p->seq) || !timestep_host || !outs) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "rife: null argument");
  if (io_dtype != RIFE_F32 && io_dtype != RIFE_F16 && io_dtype != RIFE_U8)
    RIFE_FAIL(RIFE_ERR_DTYPE, "rife: io_dtype must be RIFE_F32, RIFE_F16 or RIFE_U8");
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev != p->device) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "rife: plan belongs to another device");
  const rife_weights* w = p->w;
  const PlanLayout& L = p->L;
  const int B = p->B, H = p->H, W = p->W;
  const bool half_act = is_half_mode(p->precision);
  const size_t E = half_act ? 2 : 4;
  const size_t N = (size_t)B * H * W;

  if (!graph_capture) RIFE_TRY(update_tstep(p, timestep_host, st));

  PixelCtx c;
  c.B = B; c.C = p->C; c.H = H; c.W = W;
  c.h = p->fh; c.w = p->fw;
  c.gridx = (const float*)p->at(L.gridx);
  c.gridy = (const float*)p->at(L.gridy);
  c.tstep = (const float*)p->at(L.tstep);
  // SRC frames: [img0's frames | img1's frames]; output b = pair * T + t warps frames (pair, pair + P), or in sequence
  // mode (one run of P + 1 frames) the neighbours (pair, pair + 1)
  c.F = p->nf0() + p->nf1(); c.nf0 = p->nf0(); c.src_div = p->T;
  c.src0 = p->at(L.src);
  c.src1 = (char*)p->at(L.src) + (size_t)(p->seq ? 1 : p->P) * H * W * SRC_C * E;
  c.half_act = half_act;
  c.planar_head = half_act && w->head_hk;
  c.tile_ctrs = (unsigned int*)p->at(L.ctrs); c.n_tile_ctrs = kTileCtrs;
  if (graph_capture) c.io_table = (const void* const*)p->at(L.io_tab);
  p->next_ctr = 0;
  for (int i = 0; i < 5; ++i) c.any_up = c.any_up || L.u[i] > 1;
  c.any_up = c.any_up || L.generic;   // an arbitrary scale_list takes the generic (level-by-level) pixel kernels too

  // ---- inputs + Head (IFNet_HDv3.py:128-129, both frames as one batch of 2B)
  p->n_stage = 0;
  RIFE_STAGE("prep", launch_prep(c, img0, img1, io_dtype, p->at(L.src), p->at(L.head_in), st));
  if (c.planar_head && w->has_head_fused && head_fused_on()) {
    // cnn0 -> cnn1 -> cnn2 -> cnn3 + SRC assembly in one kernel: the 16-channel activations stay in shared memory
    RIFE_STAGE("head.fused", launch_head_fused(w->head_fused, p->at(L.head_in), p->at(L.src), c.F, H / 2, W / 2, st));
  } else {
    ConvIO io;
    io.B = c.F; io.H = H / 2; io.W = W / 2;
    io.in = p->at(L.head_in); io.out = p->at(L.head_a); io.epi = EPI_LRELU; io.out_c = 16;
    RIFE_STAGE("head.cnn0", run_conv(p, w->head[0], io, c.planar_head, st));
    const void* head_feat = p->at(L.head_a);
    if (c.planar_head && w->has_strip_head) {
      // cnn1 + cnn2 in one launch: the activation between them never leaves shared memory (conv_strip.cu)
      io.in = p->at(L.head_a); io.out = p->at(L.head_b);
      io.in_bf16 = io.out_bf16 = w->head[1].in_bf16 ? 1 : 0;
      RIFE_STAGE("head.cnn1+2", launch_conv_strip2(w->strip_head, io, st));
      io.in_bf16 = io.out_bf16 = 0;
      head_feat = p->at(L.head_b);
    } else {
      io.in = p->at(L.head_a); io.out = p->at(L.head_b);
      RIFE_STAGE("head.cnn1", run_conv(p, w->head[1], io, c.planar_head, st));
      io.in = p->at(L.head_b); io.out = p->at(L.head_a);
      RIFE_STAGE("head.cnn2", run_conv(p, w->head[2], io, c.planar_head, st));
    }
    io.in = head_feat; io.out = p->at(L.src); io.epi = EPI_HEADT; io.out_c = SRC_C; io.co_per_par = 4;
    io.res = c.planar_head ? p->at(L.head_in) : nullptr;  // halo kernel: assembles the full SRC pixel (see conv_halo.cu)
    RIFE_STAGE("head.cnn3", run_conv(p, w->head[3], io, c.planar_head, st));
  }

  // ---- 5 IFBlocks (IFNet_HDv3.py:139-157)
  // The blocks that work at <= 1/4 resolution touch < 7 % of the pixels and are bound by per-layer latency: they run once
  // over the whole batch.  The stages at 1/2 and full resolution are bound by HBM / L2 and run in passes of `chunk` pairs
  // through the same (chunk-sized, L2-friendlier) work buffers.  One pass = the whole batch unless the caller asked for
  // per-level outputs.
  TmpLevel lev[5];
  float* fbuf = (float*)p->at(L.fbuf);
  const bool per_level_outputs = outs->mask || outs->flow[0] || outs->flow[1] || outs->flow[2] || outs->flow[3] ||
                                 outs->flow[4] || outs->warped0[0] || outs->warped0[1] || outs->warped0[2] ||
                                 outs->warped0[3] || outs->warped0[4] || outs->warped1[0] || outs->warped1[1] ||
                                 outs->warped1[2] || outs->warped1[3] || outs->warped1[4];
  int chunk_pairs = p->chunk > 0 ? p->chunk : default_chunk_pairs();
  if (per_level_outputs || c.any_up || chunk_pairs <= 0 || chunk_pairs > p->P) chunk_pairs = p->P;
  int first_chunked = 5;  // blocks [first_chunked, 5) run per pass
  if (chunk_pairs < p->P)
    for (int i = 4; i >= 1 && L.s[i] <= 2 && L.u[i] == 1; --i) first_chunked = i;
  const size_t io_esz = io_dtype == RIFE_F32 ? 4 : (io_dtype == RIFE_F16 ? 2 : 1);
  const size_t out_frame = (size_t)p->C * (p->fh ? p->fh : H) * (p->fw ? p->fw : W) * io_esz;

  int fbuf_levels = 0;  // number of levels summed in fbuf (0 = nothing stored yet)
  auto run_block = [&](int i, const PixelCtx& cc, int b0, bool last_pass) -> int {
    // block i for the outputs [b0, b0 + cc.B): cc is the (sub-)batch view, lev[j < i] already point at its tmp levels
    const int s = L.s[i], up = L.u[i], c_ = kBlockC[i];
    const int hs = H / s * up, wsz = W / s * up;
    const int Bc = cc.B;
    const std::string bn = "b" + std::to_string(i);
    const int hkm = half_act ? w->blk_hk[i] : 0;
    const bool hk = hkm == 1;            // conv0.* on the halo kernel (planar-8 block input)
    const bool hk_tail = hkm != 0;       // ResConv x8 + lastconv on the halo kernel
    // the last block at block scale 1 (1080p pyramid: block 4): input gather fused into conv0.0
    const bool fused = hk && i == 4 && s == 1 && up == 1 && !c.any_up && gather_conv_on() && w->blk[i].conv00.use_hk &&
                       gather_conv_fits(w->blk[i].conv00.hk) && gather_conv_levels_fit(i, lev, fbuf_levels, H, W);
    float* fb_eval = fbuf;               // flow buffer the evaluation after this block reads
    if (up > 1) {
      // scale_list[i] < 1: the block input is the full-resolution one, bilinearly up-sampled (IFNet_HDv3.py:79-81)
      RIFE_STAGE(bn + ".input", launch_block_input(cc, i, lev, 1, p->at(L.xfull), false, fbuf, fbuf_levels, i > 0, st));
      RIFE_STAGE(bn + ".upsample", launch_upsample_xin(cc, p->at(L.xfull), p->at(L.xin), up, hk, st));
    } else if (fused) {
      // gather + conv0.0 in one kernel: the block input lives in shared memory only (conv_gather.cu)
      RIFE_STAGE(bn + ".input+conv0.0",
                 launch_gather_conv00(cc, i, lev, fbuf, (float*)p->at(L.fbuf2), fbuf_levels, w->blk[i].conv00.hk, p->at(L.c0),
                                      w->blk[i].conv00.out_bf16 ? 1 : 0, st));
      fb_eval = (float*)p->at(L.fbuf2);
    } else {
      RIFE_STAGE(bn + ".input", launch_block_input(cc, i, lev, s, p->at(L.xin), hk, fbuf, fbuf_levels, i > 0 && s <= 2, st));
    }
    if (i > 0 && s <= 2) fbuf_levels = i;  // every full-resolution pixel was sampled: fbuf = flow after blocks [0, i)
    ConvIO io;
    io.B = Bc;
    // conv0.0: stride 2 on (hs, wsz) == 2x2 taps on the s2d grid (hs/2, wsz/2); output stored s2d again
    io.H = hs / 2; io.W = wsz / 2; io.in = p->at(L.xin); io.out = p->at(L.c0); io.epi = EPI_LRELU_S2D; io.out_c = c_ / 2;
    if (!fused) RIFE_STAGE(bn + ".conv0.0", run_conv(p, w->blk[i].conv00, io, hk, st));
    // conv0.1: stride 2 on (hs/2, wsz/2) == 2x2 taps on (hs/4, wsz/4)
    io.H = hs / 4; io.W = wsz / 4; io.in = p->at(L.c0); io.out = p->at(L.act_a); io.epi = EPI_LRELU; io.out_c = c_;
    io.planar_out = (hk_tail && !hk) ? 1 : 0;
    RIFE_STAGE(bn + ".conv0.1", run_conv(p, w->blk[i].conv01, io, hk, st));
    io.planar_out = 0;
    void* cur = p->at(L.act_a);
    void* nxt = p->at(L.act_b);
    bool chained = false;
    if (hk_tail && c_ >= chain_min_c(p)) {
      // ResConv x8 (IFNet_HDv3.py:63-72) as ONE cooperative launch with grid barriers between the layers; the result
      // of the 8th layer lands in act_a again
      const HaloLayer* ls[8];
      for (int j = 0; j < 8; ++j) ls[j] = &w->blk[i].res[j].hk;
      io.in = cur; io.res = cur; io.out = nxt; io.epi = EPI_RES; io.out_c = c_;
      io.in_bf16 = io.out_bf16 = w->blk[i].res[0].in_bf16 ? 1 : 0;
      int rc;
      {
        StageTimer _t(p, st, bn + ".res");
        rc = launch_conv_halo_chain(ls, 8, io, (unsigned int*)p->at(L.sync), &p->sync_host, st);
      }
      io.in_bf16 = io.out_bf16 = 0;
      if (rc < 0) return rc;
      chained = rc == 0;   // 1: this shape cannot be chained (no instantiation / not co-resident)
    }
    if (!chained && i == 4 && hk_tail && w->has_strip_res4) {
      // ResConv layers (2j, 2j + 1) as one launch each (conv_strip.cu)
      for (int j = 0; j < 4; ++j) {
        io.in = cur; io.res = nullptr; io.out = nxt; io.epi = EPI_RES; io.out_c = c_;
        io.in_bf16 = io.out_bf16 = w->blk[i].res[0].in_bf16 ? 1 : 0;
        RIFE_STAGE(bn + ".res", launch_conv_strip2(w->strip_res4[j], io, st));
        io.in_bf16 = io.out_bf16 = 0;
        void* t = cur; cur = nxt; nxt = t;
      }
      chained = true;
    }
    if (!chained) {
      for (int j = 0; j < 8; ++j) {  // ResConv x8 (IFNet_HDv3.py:63-72)
        io.in = cur; io.res = cur; io.out = nxt; io.epi = EPI_RES; io.out_c = c_;
        RIFE_STAGE(bn + ".res", run_conv(p, w->blk[i].res[j], io, hk_tail, st));
        void* t = cur; cur = nxt; nxt = t;
      }
    }
    // lastconv + PixelShuffle -> tmp_i at (hs, wsz): the batch's own slice for a whole-batch block, the front of the
    // buffer for a per-pass block
    const bool chunked = i >= first_chunked;
    const size_t tmp_item = (size_t)hs * wsz * TMP_C * E;
    char* tmp_ptr = (char*)p->at(L.tmp[i]) + (chunked ? 0 : (size_t)b0 * tmp_item);
    io.in = cur; io.res = nullptr; io.out = tmp_ptr; io.epi = EPI_PS; io.out_c = TMP_C; io.co_per_par = kLastCo;
    lev[i].ptr = tmp_ptr; lev[i].h = hs; lev[i].w = wsz; lev[i].s = up > 1 ? -up : s; lev[i].pitch = TMP_C;
    // Model.inference (only `merged` wanted) at scale <= 1: the merge kernel needs flow + mask of the last block only
    const bool narrow = i == 4 && hk_tail && w->blk[i].has_last8 && !narrow_off() && s == 1 && up == 1 && fbuf_levels == 4 &&
                        !c.any_up && outs->merged && !per_level_outputs;
    if (narrow) {
      io.epi = EPI_PS8; io.out_c = 8;
      io.in_bf16 = w->blk[i].last.in_bf16 ? 1 : 0; io.out_bf16 = 0;
      lev[i].pitch = 8;
      if (dyn_tiles_on() && p->next_ctr < kTileCtrs) io.tile_ctr = (unsigned int*)p->at(L.ctrs) + p->next_ctr++;
      RIFE_STAGE(bn + ".lastconv", launch_conv_halo(w->blk[i].last8, io, st));
      io.tile_ctr = nullptr;
    } else {
      RIFE_STAGE(bn + ".lastconv", run_conv(p, w->blk[i].last, io, hk_tail, st));
    }

    EvalOut eo;
    eo.flow = outs->flow[i];
    eo.warped0 = outs->warped0[i];
    eo.warped1 = outs->warped1[i];
    if (i == 4) {
      eo.mask = outs->mask;
      eo.merged = outs->merged ? (char*)outs->merged + (size_t)b0 * out_frame : nullptr;
    }
    if (eo.flow || eo.warped0 || eo.warped1 || eo.mask || eo.merged)
      RIFE_STAGE(i == 4 ? std::string("merge") : bn + ".eval", launch_eval(cc, i + 1, lev, eo, io_dtype, fb_eval, fbuf_levels, st));
    (void)last_pass;
    return RIFE_OK;
  };

  for (int i = 0; i < first_chunked; ++i) RIFE_TRY(run_block(i, c, 0, true));
  if (first_chunked < 5) {
    const int fb0 = fbuf_levels;
    const void* lev_base[5];
    for (int j = 0; j < first_chunked; ++j) lev_base[j] = lev[j].ptr;
    for (int p0 = 0; p0 < p->P; p0 += chunk_pairs) {
      const int np = std::min(chunk_pairs, p->P - p0);
      PixelCtx cc = c;
      cc.B = np * p->T;
      const int b0 = p0 * p->T;
      cc.tstep = c.tstep + b0;
      cc.src0 = (const char*)c.src0 + (size_t)p0 * H * W * SRC_C * E;
      cc.src1 = (const char*)c.src1 + (size_t)p0 * H * W * SRC_C * E;
      for (int j = 0; j < first_chunked; ++j)
        lev[j].ptr = (const char*)lev_base[j] + (size_t)b0 * lev[j].h * lev[j].w * TMP_C * E;
      fbuf_levels = fb0;
      for (int i = first_chunked; i < 5; ++i) RIFE_TRY(run_block(i, cc, b0, p0 + np >= p->P));
    }
  }
  return RIFE_OK;
}

extern "C" {

int rife_ifnet_forward(rife_plan* p, const void* img0, const void* img1, const float* timestep_host,
                       const rife_ifnet_outputs* outs, int io_dtype, void* stream) {
  return run_pipeline(p, img0, img1, timestep_host, outs, io_dtype, (cudaStream_t)stream);
}

static int graph_max_outputs() {
  // RIFE_GRAPH_MAX=n: calls with at most n outputs replay a CUDA graph (default 4: above that the launches are hidden
  // behind the kernels anyway); 0 turns graphs off
  static const int v = getenv("RIFE_GRAPH_MAX") ? atoi(getenv("RIFE_GRAPH_MAX")) : 4;
  return v;
}

int rife_inference(rife_plan* p, const void* img0, const void* img1, const float* timestep_host, void* out, int io_dtype,
                   void* stream) {
  if (!out) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "rife_inference: out is null");
  rife_ifnet_outputs o;
  memset(&o, 0, sizeof(o));
  o.merged = out;
  cudaStream_t st = (cudaStream_t)stream;
  static const bool hk_debug = getenv("RIFE_HK_DEBUG") != nullptr;
  const bool want_graph = p && img0 && timestep_host && p->B <= graph_max_outputs() && !p->profile && p->graph_state >= 0 &&
                          !hk_debug && (p->chunk <= 0 || p->chunk >= p->P) && default_chunk_pairs() <= 0 &&
                          (io_dtype == RIFE_F32 || io_dtype == RIFE_F16 || io_dtype == RIFE_U8);
  if (!want_graph || p->plain_calls < 1) {
    if (p) p->plain_calls++;
    return run_pipeline(p, img0, img1, timestep_host, &o, io_dtype, st);
  }
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev != p->device) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "rife: plan belongs to another device");
  if (!p->gexec || p->g_io != io_dtype || p->g_fh != p->fh || p->g_fw != p->fw || p->g_chunk != p->chunk) {
    if (p->gexec) { cudaGraphExecDestroy(p->gexec); p->gexec = nullptr; }
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
      cudaGetLastError();   // the caller is capturing this stream itself: just enqueue
      return run_pipeline(p, img0, img1, timestep_host, &o, io_dtype, st);
    }
    const uint64_t n0 = g_launch_count.load();
    cudaGraph_t g = nullptr;
    int rc = RIFE_ERR_CUDA;
    std::string why = "cudaStreamCreate";
    if (p->cap_stream || cudaStreamCreateWithFlags(&p->cap_stream, cudaStreamNonBlocking) == cudaSuccess) {
      why = "cudaStreamBeginCapture";
      if (cudaStreamBeginCapture(p->cap_stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        rc = run_pipeline(p, img0, img1, timestep_host, &o, io_dtype, p->cap_stream, /*graph_capture=*/true);
        if (rc != RIFE_OK) why = std::string("launch during capture: ") + rife_last_error();
        const cudaError_t e = cudaStreamEndCapture(p->cap_stream, &g);
        if (e != cudaSuccess || !g) {
          if (rc == RIFE_OK) why = std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e);
          rc = RIFE_ERR_CUDA;
        }
      }
    }
    const int captured = (int)(g_launch_count.load() - n0);
    g_launch_count.fetch_sub((uint64_t)captured);   // captured, not launched
    if (rc == RIFE_OK) {
      const cudaError_t e = cudaGraphInstantiate(&p->gexec, g, 0);
      if (e != cudaSuccess) { rc = RIFE_ERR_CUDA; why = std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e); }
    }
    if (g) cudaGraphDestroy(g);
    if (rc != RIFE_OK) {
      cudaGetLastError();
      p->gexec = nullptr;
      p->graph_state = -1;   // e.g. a driver that cannot capture programmatic dependent launches: plain launches from now on
      static std::atomic<bool> told{false};
      if (!told.exchange(true)) fprintf(stderr, "librife_b200: CUDA-graph capture unavailable (%s); using plain launches\n", why.c_str());
      return run_pipeline(p, img0, img1, timestep_host, &o, io_dtype, st);
    }
    p->graph_state = 1;
    p->g_io = io_dtype; p->g_fh = p->fh; p->g_fw = p->fw; p->g_chunk = p->chunk; p->g_launches = captured;
  }
  RIFE_TRY(update_tstep(p, timestep_host, st));
  RIFE_TRY(launch_set_io_table((const void**)p->at(p->L.io_tab), img0, img1, out, st));
  RIFE_CUDA_TRY(cudaGraphLaunch(p->gexec, st));
  g_launch_count.fetch_add((uint64_t)p->g_launches);
  return RIFE_OK;
}

int rife_inference_host(rife_plan* p, const void* img0_host, const void* img1_host, const float* timestep_host,
                        void* out_host, int io_dtype, void* stream) {
  if (!p || !img0_host || (!img1_host && !p->seq) || !out_host) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "rife_inference_host: null argument");
  if (io_dtype != RIFE_F32 && io_dtype != RIFE_F16 && io_dtype != RIFE_U8) RIFE_FAIL(RIFE_ERR_DTYPE, "rife_inference_host: bad io_dtype");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t esz = io_dtype == RIFE_F32 ? 4 : (io_dtype == RIFE_F16 ? 2 : 1);
  const size_t frame = (size_t)p->C * (p->fh ? p->fh : p->H) * (p->fw ? p->fw : p->W) * esz;
  RIFE_CUDA_TRY(cudaMemcpyAsync(p->at(p->L.stage_in0), img0_host, frame * p->nf0(), cudaMemcpyHostToDevice, st));
  if (p->nf1()) RIFE_CUDA_TRY(cudaMemcpyAsync(p->at(p->L.stage_in1), img1_host, frame * p->nf1(), cudaMemcpyHostToDevice, st));
  RIFE_TRY(rife_inference(p, p->at(p->L.stage_in0), p->at(p->L.stage_in1), timestep_host, p->at(p->L.stage_out), io_dtype,
                          stream));
  RIFE_CUDA_TRY(cudaMemcpyAsync(out_host, p->at(p->L.stage_out), frame * p->B, cudaMemcpyDeviceToHost, st));
  return RIFE_OK;
}

int rife_inference_host_pipelined(rife_plan* p, const void* img0_host, const void* img1_host, const float* timestep_host,
                                  void* out_host, int io_dtype, void* stream_h2d, void* stream_compute, void* stream_d2h) {
  if (!p || !img0_host || (!img1_host && !p->seq) || !out_host) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "rife_inference_host_pipelined: null argument");
  if (io_dtype != RIFE_F32 && io_dtype != RIFE_F16 && io_dtype != RIFE_U8) RIFE_FAIL(RIFE_ERR_DTYPE, "rife_inference_host_pipelined: bad io_dtype");
  cudaStream_t s_in = (cudaStream_t)stream_h2d, s_c = (cudaStream_t)stream_compute, s_out = (cudaStream_t)stream_d2h;
  if (!p->ev_in) {
    RIFE_CUDA_TRY(cudaEventCreateWithFlags(&p->ev_in, cudaEventDisableTiming));
    RIFE_CUDA_TRY(cudaEventCreateWithFlags(&p->ev_compute, cudaEventDisableTiming));
    RIFE_CUDA_TRY(cudaEventCreateWithFlags(&p->ev_out, cudaEventDisableTiming));
    // first use: nothing to wait for
    RIFE_CUDA_TRY(cudaEventRecord(p->ev_compute, s_c));
    RIFE_CUDA_TRY(cudaEventRecord(p->ev_out, s_out));
  }
  const size_t esz = io_dtype == RIFE_F32 ? 4 : (io_dtype == RIFE_F16 ? 2 : 1);
  const size_t frame = (size_t)p->C * (p->fh ? p->fh : p->H) * (p->fw ? p->fw : p->W) * esz;
  // H2D: the previous inference on this plan must have consumed its staging-in buffers
  RIFE_CUDA_TRY(cudaStreamWaitEvent(s_in, p->ev_compute, 0));
  RIFE_CUDA_TRY(cudaMemcpyAsync(p->at(p->L.stage_in0), img0_host, frame * p->nf0(), cudaMemcpyHostToDevice, s_in));
  if (p->nf1()) RIFE_CUDA_TRY(cudaMemcpyAsync(p->at(p->L.stage_in1), img1_host, frame * p->nf1(), cudaMemcpyHostToDevice, s_in));
  RIFE_CUDA_TRY(cudaEventRecord(p->ev_in, s_in));
  // compute: frames resident, previous result of this plan already copied out
  RIFE_CUDA_TRY(cudaStreamWaitEvent(s_c, p->ev_in, 0));
  RIFE_CUDA_TRY(cudaStreamWaitEvent(s_c, p->ev_out, 0));
  RIFE_TRY(rife_inference(p, p->at(p->L.stage_in0), p->at(p->L.stage_in1), timestep_host, p->at(p->L.stage_out), io_dtype,
                          stream_compute));
  RIFE_CUDA_TRY(cudaEventRecord(p->ev_compute, s_c));
  // D2H
  RIFE_CUDA_TRY(cudaStreamWaitEvent(s_out, p->ev_compute, 0));
  RIFE_CUDA_TRY(cudaMemcpyAsync(out_host, p->at(p->L.stage_out), frame * p->B, cudaMemcpyDeviceToHost, s_out));
  RIFE_CUDA_TRY(cudaEventRecord(p->ev_out, s_out));
  return RIFE_OK;
}

// ---- single operators ----------------------------------------------------------------------------
int rife_warp(const void* input, const void* flow, void* out, int B, int C, int H, int W, int dtype,
              const float* gridx_dev, const float* gridy_dev, void* stream) {
  if (!input || !flow || !out) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "rife_warp: null argument");
  if (B < 1 || C < 1 || H < 1 || W < 1) RIFE_FAIL(RIFE_ERR_SHAPE, "rife_warp: empty shape");
  if (!is_device_ptr(input) || !is_device_ptr(flow) || !is_device_ptr(out))
    RIFE_FAIL(RIFE_ERR_NOT_DEVICE_PTR, "rife_warp: tensors must be CUDA device memory (no CPU fallback)");
  cudaStream_t st = (cudaStream_t)stream;
  float* tmpg = nullptr;
  if (!gridx_dev || !gridy_dev) {
    RIFE_CUDA_TRY(cudaMallocAsync((void**)&tmpg, (size_t)(W + H) * 4, st));
    RIFE_TRY(launch_fill_grid(tmpg, W, st));
    RIFE_TRY(launch_fill_grid(tmpg + W, H, st));
    gridx_dev = tmpg;
    gridy_dev = tmpg + W;
  }
  int rc = launch_warp_op(input, flow, out, B, C, H, W, dtype, gridx_dev, gridy_dev, st);
  if (tmpg) cudaFreeAsync(tmpg, st);
  return rc;
}

int rife_resize_bilinear(const void* x, void* out, int B, int C, int H, int W, int num, int den, float post_mul,
                         void* stream) {
  if (!x || !out) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "rife_resize_bilinear: null argument");
  auto pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
  if (!pow2(num) || !pow2(den) || (num != 1 && den != 1))
    RIFE_FAIL(RIFE_ERR_SHAPE, "rife_resize_bilinear: scale_factor must be 2^k or 2^-k");
  if ((H * (long long)num) % den || (W * (long long)num) % den) RIFE_FAIL(RIFE_ERR_SHAPE, "rife_resize_bilinear: size not divisible");
  if (!is_device_ptr(x) || !is_device_ptr(out)) RIFE_FAIL(RIFE_ERR_NOT_DEVICE_PTR, "rife_resize_bilinear: device memory required");
  return launch_resize_op((const float*)x, (float*)out, B, C, H, W, num, den, post_mul, (cudaStream_t)stream);
}

// Test entry points: repack on the fly (synchronous), run the same kernels/epilogues as the pipeline.
static int run_single_conv(HostLayer& h, int precision, const float* x_nchw, int B, int Cin, int H, int W, bool s2d,
                           const float* residual_nchw, float* y_nchw, int epi, int out_c_store, int cpp, int out_C,
                           int out_H, int out_W, int out_Cpad, cudaStream_t st) {
  rife_weights tmpw;
  // test hooks: 3 = RIFE_PRECISION_FP16 through the NHWC tcgen05 kernel only, 4 = the same for RIFE_PRECISION_BF16
  tmpw.force_nhwc = precision == 3 || precision == 4;
  if (precision == 3) precision = RIFE_PRECISION_FP16;
  if (precision == 4) precision = RIFE_PRECISION_BF16;
  tmpw.precision = precision;
  cudaGetDevice(&tmpw.device);
  DevLayer d;
  const bool bf = precision == RIFE_PRECISION_BF16;
  // bf16 mode: bf16 operands; the activation epilogues store bf16, the scatter epilogues (PixelShuffle / ConvT) fp16
  const bool obf = bf && (epi == EPI_LRELU || epi == EPI_LINEAR || epi == EPI_RES);
  int rc = upload_layer(&tmpw, h, &d, epi == EPI_PS, bf, obf);
  float *xin = nullptr, *res = nullptr, *yout = nullptr;
  const int gh = s2d ? H / 2 : H, gw = s2d ? W / 2 : W;  // GEMM row grid
  const int cin_act = s2d ? h.cin / 4 : h.cin;
  auto cleanup = [&]() {
    cudaStreamSynchronize(st);
    if (xin) cudaFree(xin);
    if (res) cudaFree(res);
    if (yout) cudaFree(yout);
    for (void* p : tmpw.allocs) cudaFree(p);
    tmpw.allocs.clear();
  };
  if (rc != RIFE_OK) { cleanup(); return rc; }
#define RIFE_SC_TRY(e) do { cudaError_t _e = (e); if (_e != cudaSuccess) { set_error(std::string(#e) + ": " + cudaGetErrorString(_e)); cleanup(); return RIFE_ERR_CUDA; } } while (0)
#define RIFE_SC(e) do { int _r = (e); if (_r != RIFE_OK) { cleanup(); return _r; } } while (0)
  const size_t nin = (size_t)B * H * W * cin_act;
  const size_t nout = (size_t)B * out_H * out_W * out_Cpad;
  RIFE_SC_TRY(cudaMalloc((void**)&xin, nin * 4));
  RIFE_SC_TRY(cudaMalloc((void**)&yout, nout * 4));
  RIFE_SC_TRY(cudaMemsetAsync(yout, 0, nout * 4, st));
  RIFE_SC(launch_nchw_to_nhwc_f32(x_nchw, xin, B, Cin, H, W, cin_act, s2d, st));
  if (residual_nchw) {
    RIFE_SC_TRY(cudaMalloc((void**)&res, (size_t)B * gh * gw * out_c_store * 4));
    RIFE_SC(launch_nchw_to_nhwc_f32(residual_nchw, res, B, out_C, gh, gw, out_c_store, false, st));
  }
  ConvIO io;
  io.B = B; io.H = gh; io.W = gw; io.in = xin; io.res = res; io.out = yout; io.epi = epi; io.out_c = out_c_store;
  io.co_per_par = cpp;
  io.in_bf16 = bf ? 1 : 0; io.out_bf16 = obf ? 1 : 0;
  if (is_half_mode(precision)) {
    // streamed-weight halo instantiations exist for the ResConv / lastconv epilogues only
    if (d.use_hk && !(d.hk.stream && epi != EPI_RES && epi != EPI_PS)) RIFE_SC(launch_conv_halo_f32io(d.hk, io, st));
    else RIFE_SC(launch_conv_umma_f32io(d.h16, io, st));
  } else {
    RIFE_SC(launch_conv_f32(d.f32, io, st));
  }
  RIFE_SC(launch_nhwc_to_nchw_f32(yout, y_nchw, B, out_C, out_H, out_W, out_Cpad, st));
  cleanup();
  return RIFE_OK;
#undef RIFE_SC_TRY
#undef RIFE_SC
}

static int fetch_host(const float* dev, size_t n, std::vector<float>* h) {
  h->resize(n);
  RIFE_CUDA_TRY(cudaMemcpy(h->data(), dev, n * 4, cudaMemcpyDeviceToHost));
  return RIFE_OK;
}

int rife_conv3x3(const float* x, const float* w, const float* bias, const float* beta, const float* residual, float* y,
                 int B, int Cin, int H, int W, int Cout, int stride, int act, int precision, void* stream) {
  if (!x || !w || !y) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "rife_conv3x3: null argument");
  if (stride != 1 && stride != 2) RIFE_FAIL(RIFE_ERR_SHAPE, "rife_conv3x3: stride must be 1 or 2");
  if (stride == 2 && (H % 2 || W % 2)) RIFE_FAIL(RIFE_ERR_SHAPE, "rife_conv3x3: stride 2 needs even H, W");
  if ((beta == nullptr) != (residual == nullptr)) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "rife_conv3x3: beta and residual go together");
  if (!is_device_ptr(x) || !is_device_ptr(w) || !is_device_ptr(y)) RIFE_FAIL(RIFE_ERR_NOT_DEVICE_PTR, "rife_conv3x3: device memory required");
  RIFE_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  std::vector<float> hw, hb, hbeta;
  RIFE_TRY(fetch_host(w, (size_t)Cout * Cin * 9, &hw));
  if (bias) RIFE_TRY(fetch_host(bias, Cout, &hb));
  const int cs = round_up(Cin, 16);
  HostLayer L = stride == 2 ? pack_conv3x3_s2d(hw.data(), bias ? hb.data() : nullptr, Cout, Cin, cs, 4 * cs)
                            : pack_conv3x3(hw.data(), bias ? hb.data() : nullptr, Cout, Cin, cs);
  int epi = act ? EPI_LRELU : EPI_LINEAR;
  if (beta) {
    if (stride != 1 || !act) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "rife_conv3x3: the residual form is stride 1 + LeakyReLU");
    RIFE_TRY(fetch_host(beta, Cout, &hbeta));
    L.beta.assign(L.npad, 0.f);
    for (int k = 0; k < Cout; ++k) L.beta[k] = hbeta[k];
    epi = EPI_RES;
  }
  const int oh = H / stride, ow = W / stride;
  return run_single_conv(L, precision, x, B, Cin, H, W, stride == 2, residual, y, epi, L.npad, 0, Cout, oh, ow, L.npad,
                         (cudaStream_t)stream);
}

int rife_conv_transpose4x4s2(const float* x, const float* w, const float* bias, float* y, int B, int Cin, int H, int W,
                             int Cout, int pixel_shuffle, int precision, void* stream) {
  if (!x || !w || !y) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "rife_conv_transpose4x4s2: null argument");
  if (pixel_shuffle && Cout != kLastCo) RIFE_FAIL(RIFE_ERR_SHAPE, "rife_conv_transpose4x4s2: the fused PixelShuffle form is the 52-channel lastconv");
  if (Cout % 4) RIFE_FAIL(RIFE_ERR_SHAPE, "rife_conv_transpose4x4s2: Cout must be a multiple of 4");
  if (!is_device_ptr(x) || !is_device_ptr(w) || !is_device_ptr(y)) RIFE_FAIL(RIFE_ERR_NOT_DEVICE_PTR, "rife_conv_transpose4x4s2: device memory required");
  RIFE_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  std::vector<float> hw, hb;
  RIFE_TRY(fetch_host(w, (size_t)Cin * Cout * 16, &hw));
  if (bias) RIFE_TRY(fetch_host(bias, Cout, &hb));
  HostLayer L = pack_convT(hw.data(), bias ? hb.data() : nullptr, Cin, Cout, round_up(Cin, 16));
  if (pixel_shuffle)
    return run_single_conv(L, precision, x, B, Cin, H, W, false, nullptr, y, EPI_PS, TMP_C, Cout, Cout / 4, 4 * H, 4 * W,
                           TMP_C, (cudaStream_t)stream);
  const int cpad = round_up(Cout, 4);
  return run_single_conv(L, precision, x, B, Cin, H, W, false, nullptr, y, EPI_CONVT, cpad, Cout, Cout, 2 * H, 2 * W, cpad,
                         (cudaStream_t)stream);
}

}  // extern "C"

// --------------------------------------------------------------------------------------------------
// Generic NHWC layer / operator API (ECCV-paper network path; fp32).  See include/rife_b200.h.
struct rife_layer {
  rife_weights store;  // owns the device allocations
  DevLayer d;
  int kind = 0, cin_store = 0, cout = 0, co_pad = 0;
};

extern "C" {

int rife_layer_create(rife_layer** out, int kind, const float* w_host, const float* bias_host, const float* slope_host,
                      int cin, int cout, int cin_store, int precision) {
  if (!out || !w_host) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "rife_layer_create: null argument");
  if (kind < 0 || kind > 2) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "rife_layer_create: kind must be 0 (3x3 s1), 1 (3x3 s2) or 2 (ConvTranspose 4x4 s2)");
  if (precision != RIFE_PRECISION_FP32 && precision != RIFE_PRECISION_FP16) RIFE_FAIL(RIFE_ERR_DTYPE, "rife_layer_create: bad precision");
  const bool h16 = precision == RIFE_PRECISION_FP16;
  const int mult = (h16 && kind != 1) ? 16 : 8;  // the tcgen05 kernel consumes K in chunks of 16 channels (kind 1: 4 * cin_store)
  if (cin < 1 || cout < 1 || cin_store < cin || cin_store % mult)
    RIFE_FAIL(RIFE_ERR_SHAPE, "rife_layer_create: cin_store must be >= cin and a multiple of " + std::to_string(mult));
  auto* l = new rife_layer();
  l->store.precision = precision;
  l->store.force_nhwc = true;  // generic layers run on the NHWC kernels (no planar-8 halo variant)
  cudaGetDevice(&l->store.device);
  l->kind = kind; l->cin_store = cin_store; l->cout = cout;
  HostLayer h;
  if (kind == 0) h = pack_conv3x3(w_host, bias_host, cout, cin, cin_store);
  else if (kind == 1) h = pack_conv3x3_s2d(w_host, bias_host, cout, cin, cin_store, 4 * cin_store);
  else { l->co_pad = round_up(cout, h16 ? 8 : 4); h = pack_convT(w_host, bias_host, cin, cout, cin_store, l->co_pad); }
  if (slope_host) {
    h.beta.assign(h.npad, 0.f);
    if (kind == 2) { for (int par = 0; par < 4; ++par) for (int o = 0; o < cout; ++o) h.beta[par * l->co_pad + o] = slope_host[o]; }
    else { for (int o = 0; o < cout; ++o) h.beta[o] = slope_host[o]; }
  }
  int rc = upload_layer(&l->store, h, &l->d);
  if (rc != RIFE_OK) { rife_layer_destroy(l); return rc; }
  *out = l;
  return RIFE_OK;
}

int rife_layer_destroy(rife_layer* l) {
  if (!l) return RIFE_OK;
  for (void* p : l->store.allocs) cudaFree(p);
  delete l;
  return RIFE_OK;
}

int rife_layer_run(const rife_layer* l, const void* in, int in_pitch, void* out, int out_dtype, int out_pitch, int B, int H,
                   int W, const void* res, int res_pitch, int act, void* stream) {
  if (!l || !in || !out) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "rife_layer_run: null argument");
  const bool h16 = l->store.precision == RIFE_PRECISION_FP16;
  if (out_dtype != (h16 ? RIFE_F16 : RIFE_F32) && !(h16 && out_dtype == RIFE_F32 && act == 0))
    RIFE_FAIL(RIFE_ERR_DTYPE, "rife_layer_run: out_dtype must match the layer precision (fp16 layers may write fp32 when act == 0)");
  ConvIO io;
  io.B = B; io.H = H; io.W = W; io.in = in; io.out = out; io.res = res;
  io.in_pitch = in_pitch; io.res_pitch = res_pitch; io.out_c = out_pitch;
  if (l->kind == 2) {
    if (act != 0 && act != 2) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "rife_layer_run: ConvTranspose supports act 0 (none) or 2 (PReLU)");
    if (out_pitch < l->co_pad) RIFE_FAIL(RIFE_ERR_SHAPE, "rife_layer_run: out_pitch too small for the ConvTranspose output");
    io.epi = act == 2 ? EPI_CONVT_PRELU : EPI_CONVT;
    io.co_per_par = l->co_pad;
  } else {
    switch (act) {
      case 0: io.epi = EPI_LINEAR; break;
      case 1: io.epi = EPI_LRELU; break;
      case 2: io.epi = EPI_PRELU; break;
      case 3: io.epi = EPI_PRELU_ADD; break;
      default: RIFE_FAIL(RIFE_ERR_INVALID_ARG, "rife_layer_run: bad act");
    }
    const int nout = h16 ? l->d.h16.npad : l->d.f32.n;
    if (out_pitch < nout) RIFE_FAIL(RIFE_ERR_SHAPE, "rife_layer_run: out_pitch smaller than the padded output channels (" + std::to_string(nout) + ")");
  }
  if (act == 3 && !res) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "rife_layer_run: act 3 needs res");
  if (l->kind == 1 && in_pitch != 0 && in_pitch != 4 * l->cin_store) RIFE_FAIL(RIFE_ERR_SHAPE, "rife_layer_run: stride-2 layers read a dense s2d tensor");
  if (h16) {
    // TMA / 16-byte vector access: pitches in multiples of 8 halfs, 16-byte aligned slice bases
    io.out_f32 = out_dtype == RIFE_F32;
    if ((in_pitch % 8) || (out_pitch % 8) || (res && res_pitch % 8) || ((uintptr_t)in & 15) || ((uintptr_t)out & 15) || ((uintptr_t)res & 15))
      RIFE_FAIL(RIFE_ERR_SHAPE, "rife_layer_run: fp16 layers need pitches in multiples of 8 and 16-byte aligned slices");
    return launch_conv_umma(l->d.h16, io, (cudaStream_t)stream);
  }
  return launch_conv_f32(l->d.f32, io, (cudaStream_t)stream);
}

static bool bad_dt(int dt) { return dt != RIFE_F32 && dt != RIFE_F16; }

int rife_nhwc_from_nchw(const void* src, int io_dtype, void* dst, int dst_dtype, int B, int C, int H, int W, int pitch,
                        int off, void* stream) {
  if (!src || !dst) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "rife_nhwc_from_nchw: null argument");
  return launch_nchw_to_nhwc_slice(src, io_dtype, dst, dst_dtype, B, C, H, W, pitch, off, (cudaStream_t)stream);
}
int rife_nhwc_to_nchw(const void* src, int src_dtype, int pitch, int off, void* dst, int io_dtype, int B, int C, int H,
                      int W, int sigmoid, void* stream) {
  if (!src || !dst) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "rife_nhwc_to_nchw: null argument");
  return launch_nhwc_slice_to_nchw(src, src_dtype, pitch, off, dst, io_dtype, B, C, H, W, sigmoid, (cudaStream_t)stream);
}
int rife_nhwc_resize(const void* src, int src_dtype, int spitch, int soff, void* dst, int dst_dtype, int dpitch, int doff,
                     int B, int C, int Hs, int Ws, int Hd, int Wd, float mul, int accumulate, void* stream) {
  if (!src || !dst) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "rife_nhwc_resize: null argument");
  auto pow2 = [](int a, int b) { int r = a > b ? a / b : b / a; return (a > b ? a % b : b % a) == 0 && (r & (r - 1)) == 0; };
  if (!pow2(Hs, Hd) || !pow2(Ws, Wd) || (long long)Hs * Wd != (long long)Hd * Ws) RIFE_FAIL(RIFE_ERR_SHAPE, "rife_nhwc_resize: power-of-two isotropic factors only");
  return launch_resize_nhwc(src, src_dtype, spitch, soff, dst, dst_dtype, dpitch, doff, B, C, Hs, Ws, Hd, Wd, mul, accumulate,
                            (cudaStream_t)stream);
}
int rife_nhwc_warp(const void* src, int src_dtype, int spitch, int soff, const void* flow, int flow_dtype, int fpitch,
                   int foff, void* dst, int dst_dtype, int dpitch, int doff, int B, int C, int H, int W,
                   const float* gridx_dev, const float* gridy_dev, void* stream) {
  if (!src || !flow || !dst || !gridx_dev || !gridy_dev) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "rife_nhwc_warp: null argument");
  return launch_warp_nhwc(src, src_dtype, spitch, soff, flow, flow_dtype, fpitch, foff, dst, dst_dtype, dpitch, doff, B, C, H,
                          W, gridx_dev, gridy_dev, (cudaStream_t)stream);
}
int rife_nhwc_s2d(const void* src, int src_dtype, int spitch, int soff, int C, void* dst, int dst_dtype, int cpad, int B,
                  int H, int W, void* stream) {
  if (!src || !dst) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "rife_nhwc_s2d: null argument");
  if (H % 2 || W % 2 || cpad < C) RIFE_FAIL(RIFE_ERR_SHAPE, "rife_nhwc_s2d: even H, W and cpad >= C required");
  return launch_s2d_nhwc(src, src_dtype, spitch, soff, C, dst, dst_dtype, cpad, B, H, W, (cudaStream_t)stream);
}
int rife_scene_ssim(const void* img0, const void* img1, int io_dtype, int B, int C, int h, int w, int pad_multiple,
                    float* ssim_out, void* stream) {
  if (!img0 || !img1 || !ssim_out) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "rife_scene_ssim: null argument");
  if (B < 1 || C < 3 || h < 1 || w < 1 || pad_multiple < 1) RIFE_FAIL(RIFE_ERR_SHAPE, "rife_scene_ssim: need B >= 1, C >= 3, h, w, pad_multiple >= 1");
  if (!is_device_ptr(img0) || !is_device_ptr(img1) || !is_device_ptr(ssim_out))
    RIFE_FAIL(RIFE_ERR_NOT_DEVICE_PTR, "rife_scene_ssim: device memory required (no CPU fallback)");
  const int Hp = round_up(h, pad_multiple), Wp = round_up(w, pad_multiple);
  return launch_scene_ssim(img0, img1, io_dtype, B, C, h, w, Hp, Wp, ssim_out, (cudaStream_t)stream);
}
int rife_nhwc_fill(void* dst, int dst_dtype, int pitch, int off, int B, int C, int H, int W, float value, void* stream) {
  if (!dst) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "rife_nhwc_fill: null argument");
  if (B < 1 || C < 1 || H < 1 || W < 1 || off < 0 || off + C > pitch) RIFE_FAIL(RIFE_ERR_SHAPE, "rife_nhwc_fill: bad slice");
  return launch_fill_nhwc(dst, dst_dtype, pitch, off, B, C, H, W, value, (cudaStream_t)stream);
}
int rife_eccv_merge(const void* w0, int p0, int o0, const void* w1, int p1, int o1, const void* mask, int pm, int om,
                    int w_dtype, const void* refine, int refine_dtype, int pr, int orf, void* out, int io_dtype, int B,
                    int H, int W, void* stream) {
  if (!w0 || !w1 || !mask || !out) RIFE_FAIL(RIFE_ERR_INVALID_ARG, "rife_eccv_merge: null argument");
  if (bad_dt(w_dtype) || (refine && bad_dt(refine_dtype))) RIFE_FAIL(RIFE_ERR_DTYPE, "rife_eccv_merge: bad tensor dtype");
  return launch_eccv_merge(w0, p0, o0, w1, p1, o1, mask, pm, om, w_dtype, refine, refine_dtype, pr, orf, out, io_dtype, B, H,
                           W, (cudaStream_t)stream);
}

}  // extern "C"
