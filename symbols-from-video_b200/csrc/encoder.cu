// encoder.cu -- KL-f8 encoder handle: weight ingestion/repacking and the forward
// schedule.  Mirrors Encoder.forward (ldm/modules/diffusionmodules/model.py:434-459),
// ResnetBlock.forward (:121-141), Downsample.forward (:72-79), AttnBlock.forward
// (:178-202) and AutoencoderKL.encode (ldm/models/autoencoder.py:324-328).
//
// Data layout in HBM (per chunk of Bc frames):
//   x stream      fp32 NHWC   (residual stream; never rounded to 16 bit)
//   operands      16-bit NHWC (bf16 or fp16): GroupNorm+SiLU outputs, conv1 outputs,
//                 and a 16-bit copy of x where a conv consumes x directly
//                 (downsample, nin_shortcut)
//   GN statistics fp64 [Bc][32][2]
// In fp32 check mode the operand buffers hold fp32 and every GEMM runs on the
// CUDA-core kernel.
#include "common.cuh"
#include "encoder.h"
#include <math.h>
#include <string.h>

namespace sfv {

// ------------------------------------------------------------------ weights
static constexpr float kConvInScale = 4096.f;
static const SfvTensor* find_tensor(const SfvTensor* t, int n, const std::string& name) {
  for (int i = 0; i < n; ++i)
    if (t[i].name && name == t[i].name) return &t[i];
  return nullptr;
}

static uint16_t host_to_16(float v, int fmt) {
  if (fmt == FMT_BF16) {
    uint32_t u; memcpy(&u, &v, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);   // NaN
    const uint32_t r = 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)((u + r) >> 16);
  }
  __half h = __float2half_rn(v);
  uint16_t o; memcpy(&o, &h, 2);
  return o;
}

static float host_from_16(uint16_t b, int fmt) {
  if (fmt == FMT_BF16) { const uint32_t u = (uint32_t)b << 16; float v; memcpy(&v, &u, 4); return v; }
  __half h; memcpy(&h, &b, 2);
  return __half2float(h);
}

int DeviceBlob::upload(const void* host, size_t bytes, void** out) {
  void* d = nullptr;
  SFV_CUDA(cudaMalloc(&d, bytes ? bytes : 16));
  allocs.push_back(d);
  if (bytes) SFV_CUDA(cudaMemcpy(d, host, bytes, cudaMemcpyHostToDevice));
  *out = d;
  return 0;
}
void DeviceBlob::release() {
  for (void* p : allocs) cudaFree(p);
  allocs.clear();
}

// OIHW fp32 host weights -> (a) fp32 [K][Cout] for the CUDA-core kernel,
// (b) 16-bit [Cout_pad][K] K-major rows for UMMA; k = (r*ks + s)*Cin + ci.
float weight_scale_for(int fmt, double maxabs) {
  if (fmt != FMT_F16 || !(maxabs > 0.0) || !std::isfinite(maxabs)) return 1.f;
  int e = (int)floor(log2(16384.0 / maxabs));          // maxabs * 2^e in [2^13, 2^14)
  if (e > 24) e = 24;
  if (e < -24) e = -24;
  return (float)ldexp(1.0, e);
}

int make_conv_from_host(DeviceBlob& blob, const float* w, const float* b, int Cout, int Cin, int ks,
                        int fmt, bool want16, ConvW* out) {
  out->Cin = Cin; out->Cout = Cout; out->ks = ks;
  const int K = ks * ks * Cin;
  out->cout_pad = (Cout + 15) / 16 * 16;
  std::vector<float> w32((size_t)K * Cout);
  for (int o = 0; o < Cout; ++o)
    for (int i = 0; i < Cin; ++i)
      for (int t = 0; t < ks * ks; ++t)
        w32[((size_t)t * Cin + i) * Cout + o] = w[((size_t)o * Cin + i) * ks * ks + t];
  SFV_TRY(blob.upload(w32.data(), w32.size() * 4, (void**)&out->w32));
  std::vector<float> bias(out->cout_pad, 0.f);
  if (b) for (int o = 0; o < Cout; ++o) bias[o] = b[o];
  SFV_TRY(blob.upload(bias.data(), bias.size() * 4, (void**)&out->bias));
  out->w16 = nullptr;
  out->w_scale = 1.f;
  if (want16 && Cin % 64 == 0) {
    double mx = 0;
    for (size_t i = 0; i < (size_t)Cout * Cin * ks * ks; ++i) mx = fmax(mx, fabs((double)w[i]));
    const float sc = out->w_scale = weight_scale_for(fmt, mx);
    std::vector<uint16_t> w16((size_t)out->cout_pad * K, 0);
    for (int o = 0; o < Cout; ++o)
      for (int i = 0; i < Cin; ++i)
        for (int t = 0; t < ks * ks; ++t)
          w16[(size_t)o * K + (size_t)t * Cin + i] = host_to_16(w[((size_t)o * Cin + i) * ks * ks + t] * sc, fmt);
    SFV_TRY(blob.upload(w16.data(), w16.size() * 2, &out->w16));
  }
  return 0;
}

// uint8-fed tensor-core conv_in: x = (2u-255)/255 exactly, so with A = 2u-255 (an odd integer <= 255 in
// magnitude: exact in bf16 and fp16, and the zero padding of x stays 0) the layer is A . (w/255).  The weight is
// split hi + lo so the product keeps ~16 (bf16) / 22 (fp16) mantissa bits; kConvInScale keeps lo out of the
// fp16 subnormals and is undone by the epilogue's alpha.  w: OIHW [128][3][3][3].
int make_conv_in_u8(DeviceBlob& blob, const float* w, int fmt, ConvW* out) {
  std::vector<uint16_t> wt((size_t)128 * 64, 0);
  for (int o = 0; o < 128; ++o)
    for (int k = 0; k < 27; ++k) {
      const int tap = k / 3, c = k % 3;                       // k = (dy*3+dx)*3 + c
      const double v = (double)w[((size_t)o * 3 + c) * 9 + tap] * (double)kConvInScale / 255.0;
      const uint16_t hi = host_to_16((float)v, fmt);
      const uint16_t lo = host_to_16((float)(v - (double)host_from_16(hi, fmt)), fmt);
      wt[(size_t)o * 64 + k] = hi;
      wt[(size_t)o * 64 + 27 + k] = lo;
    }
  return blob.upload(wt.data(), wt.size() * 2, &out->w16_u8);
}

static int get_conv(DeviceBlob& blob, const SfvTensor* t, int n, const std::string& name, int Cout,
                    int Cin, int ks, int fmt, bool want16, ConvW* out) {
  const SfvTensor* w = find_tensor(t, n, name + ".weight");
  const SfvTensor* b = find_tensor(t, n, name + ".bias");
  if (!w || !b) return fail(SFV_ERR_MISSING_KEY, "missing state-dict key %s.{weight,bias}", name.c_str());
  if (w->ndim != 4 || w->shape[0] != Cout || w->shape[1] != Cin || w->shape[2] != ks || w->shape[3] != ks ||
      b->shape[0] != Cout)
    return fail(SFV_ERR_MISSING_KEY, "%s: expected [%d,%d,%d,%d]", name.c_str(), Cout, Cin, ks, ks);
  return make_conv_from_host(blob, w->host_data, b->host_data, Cout, Cin, ks, fmt, want16, out);
}

static int get_norm(DeviceBlob& blob, const SfvTensor* t, int n, const std::string& name, int C, NormW* out) {
  const SfvTensor* w = find_tensor(t, n, name + ".weight");
  const SfvTensor* b = find_tensor(t, n, name + ".bias");
  if (!w || !b || w->shape[0] != C || b->shape[0] != C)
    return fail(SFV_ERR_MISSING_KEY, "missing/mis-shaped state-dict key %s.{weight,bias} [%d]", name.c_str(), C);
  out->C = C;
  SFV_TRY(blob.upload(w->host_data, (size_t)C * 4, (void**)&out->gamma));
  SFV_TRY(blob.upload(b->host_data, (size_t)C * 4, (void**)&out->beta));
  return 0;
}

// xc_scale: the fused nin_shortcut reads the 16-bit x copy, which holds xc_scale * x -> its weights carry 1 / xc_scale
static int get_res(DeviceBlob& blob, const SfvTensor* t, int n, const std::string& name, int Cin, int Cout,
                   int fmt, bool want16, ResW* r, float xc_scale) {
  SFV_TRY(get_norm(blob, t, n, name + ".norm1", Cin, &r->n1));
  SFV_TRY(get_conv(blob, t, n, name + ".conv1", Cout, Cin, 3, fmt, want16, &r->c1));
  SFV_TRY(get_norm(blob, t, n, name + ".norm2", Cout, &r->n2));
  SFV_TRY(get_conv(blob, t, n, name + ".conv2", Cout, Cout, 3, fmt, want16, &r->c2));
  r->has_nin = Cin != Cout;
  if (r->has_nin) {
    SFV_TRY(get_conv(blob, t, n, name + ".nin_shortcut", Cout, Cin, 1, fmt, want16, &r->nin));
    if (want16) {
      // conv2 and the 1x1 nin_shortcut write the same output (x + h, model.py:136-141): fold the shortcut into
      // conv2's GEMM as Cin extra K columns  [W2 (tap-major) | Wnin],  bias b2 + bnin
      const SfvTensor* w2 = find_tensor(t, n, name + ".conv2.weight");
      const SfvTensor* b2 = find_tensor(t, n, name + ".conv2.bias");
      const SfvTensor* wn = find_tensor(t, n, name + ".nin_shortcut.weight");
      const SfvTensor* bn = find_tensor(t, n, name + ".nin_shortcut.bias");
      const int K = 9 * Cout + Cin;
      std::vector<uint16_t> w16((size_t)Cout * K, 0);
      std::vector<float> bias(Cout);
      const float nin_gain = 1.f / xc_scale;
      double mx = 0;
      for (size_t i = 0; i < (size_t)Cout * Cout * 9; ++i) mx = fmax(mx, fabs((double)w2->host_data[i]));
      for (size_t i = 0; i < (size_t)Cout * Cin; ++i) mx = fmax(mx, fabs((double)wn->host_data[i] * nin_gain));
      const float sc = weight_scale_for(fmt, mx);
      for (int o = 0; o < Cout; ++o) {
        for (int i = 0; i < Cout; ++i)
          for (int tp = 0; tp < 9; ++tp)
            w16[(size_t)o * K + (size_t)tp * Cout + i] = host_to_16(w2->host_data[((size_t)o * Cout + i) * 9 + tp] * sc, fmt);
        for (int i = 0; i < Cin; ++i)
          w16[(size_t)o * K + 9 * Cout + i] = host_to_16(wn->host_data[(size_t)o * Cin + i] * nin_gain * sc, fmt);
        bias[o] = b2->host_data[o] + bn->host_data[o];
      }
      ConvW& f = r->c2n;
      f.w_scale = sc;
      f.Cin = Cout; f.Cout = Cout; f.ks = 3; f.cout_pad = Cout; f.extra_k = Cin;
      SFV_TRY(blob.upload(w16.data(), w16.size() * 2, &f.w16));
      SFV_TRY(blob.upload(bias.data(), bias.size() * 4, (void**)&f.bias));
    }
  }
  return 0;
}

int encoder_build(SfvEncoder* e, const SfvTensor* t, int n) {
  const bool tc = e->prec != SFV_PREC_F32;
  const int fmt = e->fmt_w;          // every weight tensor is stored in the weight format
  // accept both "encoder.x" and "first_stage_model.encoder.x" (get_percep_embeddings.py:34-39)
  std::string pre = "";
  if (!find_tensor(t, n, "encoder.conv_in.weight") && find_tensor(t, n, "first_stage_model.encoder.conv_in.weight"))
    pre = "first_stage_model.";
  const std::string p = pre + "encoder.";
  static const int mult[4] = {1, 2, 4, 4};
  SFV_TRY(get_conv(e->blob, t, n, p + "conv_in", 128, 3, 3, fmt, false, &e->conv_in));
  if (e->prec != SFV_PREC_F32) {
    const SfvTensor* w = find_tensor(t, n, p + "conv_in.weight");
    SFV_TRY(make_conv_in_u8(e->blob, w->host_data, fmt, &e->conv_in));
  }
  int cin = 128;
  for (int l = 0; l < 4; ++l) {
    const int cout = 128 * mult[l];
    for (int b = 0; b < 2; ++b) {
      SFV_TRY(get_res(e->blob, t, n, p + "down." + std::to_string(l) + ".block." + std::to_string(b), cin, cout,
                      fmt, tc, &e->down[l][b], e->xc_scale));
      cin = cout;
    }
    if (l != 3)
      SFV_TRY(get_conv(e->blob, t, n, p + "down." + std::to_string(l) + ".downsample.conv", cin, cin, 3, fmt, tc,
                       &e->ds[l]));
  }
  SFV_TRY(get_res(e->blob, t, n, p + "mid.block_1", 512, 512, fmt, tc, &e->mid1, e->xc_scale));
  SFV_TRY(get_res(e->blob, t, n, p + "mid.block_2", 512, 512, fmt, tc, &e->mid2, e->xc_scale));
  SFV_TRY(get_norm(e->blob, t, n, p + "mid.attn_1.norm", 512, &e->attn_norm));
  SFV_TRY(get_conv(e->blob, t, n, p + "mid.attn_1.q", 512, 512, 1, fmt, tc, &e->q));
  SFV_TRY(get_conv(e->blob, t, n, p + "mid.attn_1.k", 512, 512, 1, fmt, tc, &e->k));
  SFV_TRY(get_conv(e->blob, t, n, p + "mid.attn_1.v", 512, 512, 1, fmt, tc, &e->v));
  // proj_out's A operand is the attention output (fmt_attn); tcgen05 kind::f16 takes A and B in ONE 16-bit format (a
  // descriptor with a_format != b_format faults as an illegal instruction on B200), so its weights follow fmt_attn
  SFV_TRY(get_conv(e->blob, t, n, p + "mid.attn_1.proj_out", 512, 512, 1, e->fmt_attn, tc, &e->proj));
  {  // fused q|k projection: one GEMM with N = 1024
    const SfvTensor* qw = find_tensor(t, n, p + "mid.attn_1.q.weight");
    const SfvTensor* kw = find_tensor(t, n, p + "mid.attn_1.k.weight");
    const SfvTensor* qb = find_tensor(t, n, p + "mid.attn_1.q.bias");
    const SfvTensor* kb = find_tensor(t, n, p + "mid.attn_1.k.bias");
    std::vector<float> w(1024 * 512), b(1024);
    memcpy(w.data(), qw->host_data, 512 * 512 * 4);
    memcpy(w.data() + 512 * 512, kw->host_data, 512 * 512 * 4);
    memcpy(b.data(), qb->host_data, 512 * 4);
    memcpy(b.data() + 512, kb->host_data, 512 * 4);
    SFV_TRY(make_conv_from_host(e->blob, w.data(), b.data(), 1024, 512, 1, fmt, tc, &e->qk));
  }
  SFV_TRY(get_norm(e->blob, t, n, p + "norm_out", 512, &e->norm_out));
  {  // conv_out (3x3, 512->8) with quant_conv (1x1, 8->8) folded in:
     // W'[o] = sum_m Wq[o][m] Wc[m],  b' = Wq bc + bq   (autoencoder.py:325-326)
    const SfvTensor* cw = find_tensor(t, n, p + "conv_out.weight");
    const SfvTensor* cb = find_tensor(t, n, p + "conv_out.bias");
    const SfvTensor* qw = find_tensor(t, n, pre + "quant_conv.weight");
    const SfvTensor* qb = find_tensor(t, n, pre + "quant_conv.bias");
    if (!cw || !cb || !qw || !qb || cw->shape[0] != 8 || cw->shape[1] != 512 || qw->shape[0] != 8 || qw->shape[1] != 8)
      return fail(SFV_ERR_MISSING_KEY, "missing/mis-shaped conv_out / quant_conv");
    const int per = 512 * 9;
    std::vector<float> w((size_t)8 * per), b(8);
    for (int o = 0; o < 8; ++o) {
      for (int j = 0; j < per; ++j) {
        double a = 0;
        for (int m = 0; m < 8; ++m) a += (double)qw->host_data[o * 8 + m] * cw->host_data[(size_t)m * per + j];
        w[(size_t)o * per + j] = (float)a;
      }
      double a = qb->host_data[o];
      for (int m = 0; m < 8; ++m) a += (double)qw->host_data[o * 8 + m] * cb->host_data[m];
      b[o] = (float)a;
    }
    SFV_TRY(make_conv_from_host(e->blob, w.data(), b.data(), 8, 512, 3, fmt, tc, &e->conv_out));
  }
  return 0;
}

// ------------------------------------------------------------------ op dispatch
static void choose_tile(int Wo, int Ho, int* BW, int* BH) {
  double best = -1; int bw = 128;
  for (int w = 128; w >= 8; w >>= 1) {
    const int h = 128 / w;
    const double util = ((double)Wo * Ho) / ((double)ceil_div(Wo, w) * w * ceil_div(Ho, h) * h);
    if (util > best + 1e-9) { best = util; bw = w; }
  }
  *BW = bw; *BH = 128 / bw;
}

static int pick_block_n(int cout_pad) {
  if (cout_pad % 256 == 0) return 256;
  if (cout_pad % 128 == 0) return 128;
  if (cout_pad % 64 == 0) return 64;
  if (cout_pad % 32 == 0) return 32;
  return 16;
}

// conv on the tensor-core path.  in16: NHWC 16-bit [N,H,W,Cin].
int conv_tc(const ConvW& w, TcFmt fmt, const void* in16, int N, int H, int W, int stride, int pad_lo,
            int pad_hi, const void* residual, float* out_f32, void* out_16, int relu, cudaStream_t s,
            double* gn_stats, const void* a2_16, float in_scale, float out16_scale, int res16, float res_mul, int sat_check,
            const XfIn* xf) {
  SFV_CHECK(w.w16 != nullptr, "conv_tc: layer has no 16-bit weights (Cin=%d)", w.Cin);
  const int ks = w.ks, Cin = w.Cin;
  int Ho, Wo;
  TcGemmArgs a;
  memset(&a, 0, sizeof(a));
  a.a = in16; a.fmt = fmt.a; a.fmt_split = 1; a.fmt_b = fmt.b; a.fmt_out = fmt.out;
  a.out16_scale = out16_scale;
  if (stride == 1) {
    Ho = H + pad_lo + pad_hi - ks + 1; Wo = W + pad_lo + pad_hi - ks + 1;
    choose_tile(Wo, Ho, &a.BW, &a.BH);
    a.a_rank = 4;
    a.a_dims[0] = Cin; a.a_dims[1] = W; a.a_dims[2] = H; a.a_dims[3] = N;
    a.a_strides[1] = (unsigned long long)Cin * 2; a.a_strides[2] = a.a_strides[1] * W; a.a_strides[3] = a.a_strides[2] * H;
    a.a_box[0] = 64; a.a_box[1] = a.BW; a.a_box[2] = a.BH; a.a_box[3] = 1;
    a.dim_x = 1; a.dim_y = 2; a.dim_n = 3;
    a.halo_ok = (ks == 3 && pad_lo == 1 && pad_hi == 1);
    a.ntaps = ks * ks;
    for (int r = 0; r < ks; ++r)
      for (int c = 0; c < ks; ++c) {
        TcTap& t = a.taps[r * ks + c];
        t.o[1] = c - pad_lo; t.o[2] = r - pad_lo;
        a.tap_k[r * ks + c] = (r * ks + c) * Cin;
      }
  } else {
    SFV_CHECK(stride == 2 && H % 2 == 0 && W % 2 == 0, "conv_tc: stride-2 needs even H, W");
    Ho = (H + pad_lo + pad_hi - ks) / 2 + 1; Wo = (W + pad_lo + pad_hi - ks) / 2 + 1;
    choose_tile(Wo, Ho, &a.BW, &a.BH);
    a.a_rank = 5;
    a.a_dims[0] = 2ull * Cin; a.a_dims[1] = W / 2; a.a_dims[2] = 2; a.a_dims[3] = H / 2; a.a_dims[4] = N;
    a.a_strides[1] = 2ull * Cin * 2; a.a_strides[2] = (unsigned long long)W * Cin * 2;
    a.a_strides[3] = 2ull * W * Cin * 2; a.a_strides[4] = (unsigned long long)H * W * Cin * 2;
    a.a_box[0] = 64; a.a_box[1] = a.BW; a.a_box[2] = 1; a.a_box[3] = a.BH; a.a_box[4] = 1;
    a.dim_x = 1; a.dim_y = 3; a.dim_n = 4;
    a.ntaps = ks * ks;
    for (int r = 0; r < ks; ++r)
      for (int c = 0; c < ks; ++c) {
        TcTap& t = a.taps[r * ks + c];
        const int qx = c - pad_lo, qy = r - pad_lo;
        const int px = qx & 1, py = qy & 1;
        t.o[0] = px * Cin; t.o[1] = (qx - px) / 2; t.o[2] = py; t.o[3] = (qy - py) / 2;
        a.tap_k[r * ks + c] = (r * ks + c) * Cin;
      }
  }
  a.kchunks = Cin / 64;
  if (w.extra_k) {
    SFV_CHECK(a2_16 != nullptr && stride == 1 && w.extra_k % 64 == 0, "conv_tc: fused 1x1 branch needs its input tensor");
    a.a2 = a2_16; a.a2_cin = w.extra_k; a.a2_k0 = ks * ks * Cin;
  }
  a.b = w.w16; a.b_rows = w.cout_pad; a.b_k = (unsigned long long)ks * ks * Cin + w.extra_k;
  a.b_row_stride = a.b_k * 2; a.b_batched = 0;
  a.Wo = Wo; a.Ho = Ho; a.Nimg = N; a.Cout = w.Cout;
  a.block_n = pick_block_n(w.cout_pad);
  a.alpha = 1.f / (w.w_scale * in_scale); a.bias = w.bias; a.residual = residual;   // exact: both are powers of two
  a.res16 = res16; a.res_mul = res_mul; a.sat_check = sat_check;
  if (xf) {
    a.xf_stats = xf->stats; a.xf_gamma = xf->gamma; a.xf_beta = xf->beta; a.xf_in_mul = xf->in_mul; a.xf_silu = xf->silu;
    a.xf_check = xf->check;
  }
  a.out_f32 = out_f32; a.out_16 = out_16; a.ldo = w.Cout; a.relu = relu;
  if (gn_stats) {   // fused GroupNorm(32) statistics of the output; the CALLER zeroed the accumulators
    a.gn_stats = gn_stats; a.gn_cpg = w.Cout / 32;
  }
  return launch_tc_gemm(a, s);
}

// conv_in on the tensor pipe, fed from uint8 HWC frames (see the weight preparation in encoder_build)
int conv_in_tc(const ConvW& w, int fmt, const unsigned char* u8, int N, int H, int W, float* out_f32, double* gn_stats,
               cudaStream_t s, void* out_16, int fmt_out, float out16_scale) {
  SFV_CHECK(w.w16_u8 != nullptr, "conv_in_tc: no uint8 weights");
  TcGemmArgs a;
  memset(&a, 0, sizeof(a));
  a.u8_src = u8; a.fmt = fmt;
  a.BW = 128; a.BH = 1;
  a.dim_x = a.dim_y = a.dim_n = -1;
  a.ntaps = 1; a.kchunks = 1; a.tap_k[0] = 0;
  a.b = w.w16_u8; a.b_rows = 128; a.b_k = 64; a.b_row_stride = 128; a.b_batched = 0;
  a.Wo = W; a.Ho = H; a.Nimg = N; a.Cout = 128; a.block_n = 128;
  a.alpha = 1.f / kConvInScale; a.bias = w.bias;
  a.out_f32 = out_f32; a.ldo = 128;
  if (out_16) { a.out_16 = out_16; a.fmt_split = 1; a.fmt_b = fmt; a.fmt_out = fmt_out; a.out16_scale = out16_scale; }
  if (gn_stats) { a.gn_stats = gn_stats; a.gn_cpg = 4; }      // caller zeroed the accumulators
  return launch_tc_gemm(a, s);
}

int conv_f32(const ConvW& w, const void* in, int src_kind, int N, int H, int W, int stride, int pad_lo,
             int pad_hi, const float* residual, float* out, int relu, float in_scale, cudaStream_t s,
             void* out16, int fmt16, int range_check16) {
  IgemmArgs a;
  memset(&a, 0, sizeof(a));
  if (range_check16) {
    DevState* ds = nullptr;
    SFV_TRY(dev_state(&ds));
    a.err16 = ds->err_flag;
  }
  a.x = in; a.src_kind = src_kind; a.w = w.w32; a.w_sk = w.Cout; a.w_sn = 1; a.w_batch = 0;
  a.bias = w.bias; a.residual = residual; a.y = out; a.y16 = out16; a.fmt16 = fmt16;
  a.N = N; a.H = H; a.W = W; a.Cin = w.Cin; a.Cout = w.Cout;
  a.ksize = w.ks; a.stride = stride; a.pad = pad_lo;
  a.Ho = (H + pad_lo + pad_hi - w.ks) / stride + 1;
  a.Wo = (W + pad_lo + pad_hi - w.ks) / stride + 1;
  a.relu = relu; a.alpha = 1.f; a.in_scale = in_scale; a.ldy = w.Cout;
  return launch_igemm_f32(a, s);
}

// ------------------------------------------------------------------ forward
// SFV_ATTN_FUSED=0 falls back to materialised fp32 scores + a separate softmax kernel (A/B measurements)
static const bool g_attn_fused = []() { const char* e = getenv("SFV_ATTN_FUSED"); return !(e && atoi(e) == 0); }();

namespace {

constexpr int kStatSlots = 40;     // conv_in + 2 per ResnetBlock (10) + nin / Downsample / proj_out launches (<= 28), with slack

struct Plan {
  void *xa, *xb;            // residual stream ping-pong: fp32, or (stream16) 16-bit scaled
  void *oa, *ob, *x16, *x16b;
  double* stats; double* stat_slots; int stat_stride;   // check-mode statistics; per-launch slots of fused statistics
  float* S; void* P; float* moments;
  int attn_chunk;
};

void make_plan(bool tc, bool s16, int Bc, int H, int W, Arena& ar, Plan* p) {
  const size_t E0 = (size_t)Bc * H * W * 128;
  const size_t osz = tc ? 2 : 4;
  const size_t L = (size_t)(H / 8) * (W / 8);
  const size_t Lp0 = (L + 7) / 8 * 8;
  p->xa = ar.take(E0 * (s16 ? 2 : 4));
  p->xb = ar.take(E0 * (s16 ? 2 : 4));
  p->oa = ar.take(E0 * osz);
  p->ob = ar.take(E0 * osz);
  // fp32 stream: 16-bit copy of x feeding a downsample (also V^T in attention) and of a downsample output feeding
  // nin_shortcut.  16-bit stream: x is its own operand copy; only V^T needs room.
  p->x16 = tc ? ar.take(s16 ? (size_t)Bc * 512 * Lp0 * 2 : E0 * osz) : nullptr;
  p->x16b = (tc && !s16) ? ar.take(E0 / 4 * osz) : nullptr;
  p->stats = (double*)ar.take(sizeof(double) * 2 * 32 * Bc);
  // fused GroupNorm statistics: every producing launch of a forward gets its own slot, all zeroed by one memset
  p->stat_stride = 2 * 32 * Bc;
  p->stat_slots = (double*)ar.take(sizeof(double) * p->stat_stride * kStatSlots);
  const size_t Lp = (L + 7) / 8 * 8;                  // row pitch of S / P / V^T (16-byte rows for TMA)
  // the fp32 score matrix S exists only in the check mode and in the unfused A/B path; the tensor-core path keeps the
  // scores on chip and stores P (16 bit) only
  const bool need_S = !tc || !g_attn_fused;
  const size_t per_img = L * Lp * ((need_S ? 4 : 0) + (tc ? 2 : 0));
  size_t na = (size_t)(1024ull << 20) / (per_img ? per_img : 1);
  if (na < 1) na = 1;
  if (na > (size_t)Bc) na = Bc;
  p->attn_chunk = (int)na;
  p->S = need_S ? (float*)ar.take(na * L * Lp * 4) : nullptr;
  p->P = tc ? ar.take(na * L * Lp * 2) : nullptr;
  p->moments = (float*)ar.take((size_t)Bc * L * 8 * 4);
}

struct Fwd {
  SfvEncoder* e; cudaStream_t s; bool tc; int fmt; Plan pl; int N;
  bool s16 = false;          // residual stream stored as 16-bit * xc_scale

  // GroupNorm(32, eps 1e-6)(+SiLU).  `ready`: statistics already accumulated by the producing
  // kernel's epilogue (tensor-core mode); otherwise a statistics pass runs first (check mode).
  int gn(const NormW& nw, const void* in, bool in_is16, long long HW, int silu, void* out, const double* ready,
         float in_mul = 1.f) {
    if (!ready) {
      SFV_TRY(launch_gn_stats(in, in_is16, fmt, N, HW, nw.C, 32, pl.stats, s));
      ready = pl.stats;
    }
    return launch_gn_apply(in, in_is16, ready, nw.gamma, nw.beta, out, tc, fmt, N, HW, nw.C, 32, 1e-6f, silu, s,
                           tc && e->range_check, in_mul);
  }
  // GroupNorm of the residual stream x (fp32, or 16-bit scaled by xc_scale)
  int gn_x(const NormW& nw, const void* x, long long HW, int silu, void* out) {
    return gn(nw, x, s16, HW, silu, out, sx(), s16 ? 1.f / e->xc_scale : 1.f);
  }
  // A convolution that produces the next residual-stream tensor `xo` (+ optional residual `res` from the stream,
  // + optional 16-bit operand copy for the fp32-stream mode): handles both stream representations.
  int conv_x(const ConvW& w, const void* in_op, float in_scale, int H, int W, int stride, const void* res, void* xo,
             void* copy, const void* a2 = nullptr, const XfIn* xf = nullptr) {
    const int pad_lo = (w.ks == 3 && stride == 1) ? 1 : 0;
    const int pad_hi = (w.ks == 3) ? 1 : 0;
    const float xs = e->xc_scale;
    if (!tc) return conv_f32(w, in_op, SRC_NHWC_F32, N, H, W, stride, pad_lo, pad_hi, (const float*)res, (float*)xo, 0, 1.f, s);
    if (s16)
      return conv_tc(w, cf(), in_op, N, H, W, stride, pad_lo, pad_hi, res, nullptr, xo, 0, s, nsx(), a2, in_scale, xs,
                     res != nullptr, 1.f / xs, 0, xf);
    return conv_tc(w, cf(), in_op, N, H, W, stride, pad_lo, pad_hi, res, (float*)xo, copy, 0, s, nsx(), a2, in_scale, xs);
  }
  // formats of a conv whose A operand is a GroupNorm / conv1 output (fmt) and whose 16-bit output is one too
  TcFmt cf() const { return TcFmt{fmt, e->fmt_w, fmt}; }
  // operand-typed input -> fp32 stream and/or operand-typed output (+ fused GN statistics of the output)
  // in_scale: in_op holds in_scale * x (16-bit copies of the residual stream); out_scale: same for out_op
  int conv(const ConvW& w, const void* in_op, int H, int W, int stride, const float* residual,
           float* out_stream, void* out_op, double* stats_out, float in_scale = 1.f, float out_scale = 1.f) {
    // 3x3 s1: pad 1/1;  1x1: none;  3x3 s2 (Downsample): zero pad right/bottom only (model.py:75-77)
    const int pad_lo = (w.ks == 3 && stride == 1) ? 1 : 0;
    const int pad_hi = (w.ks == 3) ? 1 : 0;
    if (tc) return conv_tc(w, cf(), in_op, N, H, W, stride, pad_lo, pad_hi, residual, out_stream, out_op, 0, s, stats_out,
                           nullptr, in_scale, out_scale);
    float* y = out_stream ? out_stream : (float*)out_op;
    return conv_f32(w, in_op, SRC_NHWC_F32, N, H, W, stride, pad_lo, pad_hi, residual, y, 0, 1.f, s);
  }
  // statistics buffers: sx holds the stats of the current stream x (written by whichever kernel
  // produced x), sh those of conv1's output; both null in check mode.
  // A producer takes a fresh, already zeroed slot (nsx / nsh); consumers read the current one (sx / sh).
  double* sx_p = nullptr; double* sh_p = nullptr; int stat_slot = 0;
  double* take_slot() {
    if (!(tc && e->fused_stats)) return nullptr;
    if (stat_slot >= kStatSlots) return nullptr;          // (cannot happen with this graph) the consumer then runs its own statistics pass
    return pl.stat_slots + (size_t)(stat_slot++) * pl.stat_stride;
  }
  double* nsx() { return sx_p = take_slot(); }
  double* nsh() { return sh_p = take_slot(); }
  double* sx() const { return sx_p; }
  double* sh() const { return sh_p; }

  // x: residual stream (C=Cin), x_op: 16-bit operand copy of x (needed only when r.has_nin; == x with a 16-bit stream).
  // Writes the block output to `xo` (and, fp32 stream only, its operand copy to pl.x16 if want_copy).
  int resblock(const ResW& r, const void* x, const void* x_op, int H, int W, void* xo, bool want_copy,
               const void** xo_op) {
    const long long HW = (long long)H * W;
    // conv1 = conv(silu(GN1(x))), conv2 input = silu(GN2(h)).  With a 16-bit stream both GroupNorms can run inside the
    // consuming conv (transform warps normalise the raw A tile in shared memory) wherever that kernel variant exists;
    // elsewhere the stand-alone apply pass writes the operand first.
    const bool fuse = s16 && e->gn_fuse && e->fused_stats;
    XfIn xf1{sx(), r.n1.gamma, r.n1.beta, 1.f / e->xc_scale, 1, e->range_check ? 1 : 0};
    // (with fusion on, conv1's epilogue range-checks h as it stores it: the stand-alone apply pass that used to read h
    // and check it may not run)
    const int chk_h = (fuse && e->range_check) ? 1 : 0;
    double* const h_stats = nsh();             // conv1's epilogue accumulates the statistics of h here
    int st = fuse ? conv_tc(r.c1, cf(), x, N, H, W, 1, 1, 1, nullptr, nullptr, pl.ob, 0, s, h_stats, nullptr, 1.f, 1.f, 0, 1.f, chk_h, &xf1)
                  : TC_NOT_FUSABLE;
    if (st == TC_NOT_FUSABLE) {
      SFV_TRY(gn_x(r.n1, x, HW, 1, pl.oa));
      if (tc) SFV_TRY(conv_tc(r.c1, cf(), pl.oa, N, H, W, 1, 1, 1, nullptr, nullptr, pl.ob, 0, s, h_stats, nullptr, 1.f, 1.f, 0, 1.f, chk_h));
      else SFV_TRY(conv(r.c1, pl.oa, H, W, 1, nullptr, nullptr, pl.ob, h_stats));
    } else if (st != 0) {
      return st;
    }
    XfIn xf2{sh(), r.n2.gamma, r.n2.beta, 1.f, 1, e->range_check ? 1 : 0};
    const void* res = x;
    void* copy = (tc && want_copy && !s16) ? pl.x16 : nullptr;
    const float xs = e->xc_scale;              // the 16-bit copies of x (or x itself) hold xs * x
    auto conv2 = [&](const void* in, const XfIn* xf) -> int {
      if (r.has_nin && tc && e->fuse_nin) {
        // x' = nin(x) + conv2(a2): one GEMM, the 1x1 shortcut rides along as extra K chunks read from x's 16-bit copy
        // (its weights carry 1 / xs, see get_res)
        return conv_x(r.c2n, in, 1.f, H, W, 1, nullptr, xo, copy, x_op, xf);
      }
      if (r.has_nin) {
        SFV_TRY(conv_x(r.nin, x_op, tc ? xs : 1.f, H, W, 1, nullptr, xo, nullptr));
        res = xo;
      }
      return conv_x(r.c2, in, 1.f, H, W, 1, res, xo, copy, nullptr, xf);
    };
    st = fuse ? conv2(pl.ob, &xf2) : TC_NOT_FUSABLE;
    if (st == TC_NOT_FUSABLE) {
      SFV_TRY(gn(r.n2, pl.ob, tc, HW, 1, pl.oa, sh()));
      st = conv2(pl.oa, nullptr);
    }
    if (st != 0) return st;
    *xo_op = tc ? (s16 ? (const void*)xo : (const void*)copy) : (const void*)xo;
    return 0;
  }

  int attention(const void* x, int h, int w, void* xo) {
    const int L = h * w;
    const int C = 512;
    const float scale = 1.0f / sqrtf((float)C);
    SFV_TRY(gn_x(e->attn_norm, x, L, 0, pl.oa));                   // hn (no SiLU)
    if (tc) {
      const int Lp = (L + 7) / 8 * 8;                       // token counts need not be a multiple of 8: padded row pitch
      uint16_t* qk = (uint16_t*)pl.ob;                      // [N][L][1024]: q | k
      uint16_t* vT = (uint16_t*)pl.x16;                     // [N][512][Lp]
      const int fa = e->fmt_attn;                            // q, k, V^T, P, O: range is data dependent -> bf16 in MIXED mode
      SFV_TRY(conv_tc(e->qk, TcFmt{fmt, e->fmt_w, fa}, pl.oa, N, 1, L, 1, 0, 0, nullptr, nullptr, qk, 0, s, nullptr));
      // bias b_v is added after P V (rows of P sum to 1)
      SFV_TRY(vT_tc(e->v, TcFmt{e->fmt_w, fmt, fa}, pl.oa, vT, N, L, s));
      uint16_t* O = (uint16_t*)pl.oa;                       // [N][L][512]  (hn is dead after the two GEMMs above)
      for (int n0 = 0; n0 < N; n0 += pl.attn_chunk) {
        const int nn = (N - n0) < pl.attn_chunk ? (N - n0) : pl.attn_chunk;
        SFV_TRY(attention_tc(fa, qk + (size_t)n0 * L * 1024, 1024, qk + (size_t)n0 * L * 1024 + 512, 1024,
                             vT + (size_t)n0 * C * Lp, e->v.bias, pl.S, pl.P, O + (size_t)n0 * L * C, nn, L, C,
                             scale, s));
      }
      if (s16)
        SFV_TRY(conv_tc(e->proj, TcFmt{fa, fa, fmt}, O, N, h, w, 1, 0, 0, x, nullptr, xo, 0, s, nsx(), nullptr, 1.f,
                        e->xc_scale, 1, 1.f / e->xc_scale));
      else
        SFV_TRY(conv_tc(e->proj, TcFmt{fa, fa, fmt}, O, N, h, w, 1, 0, 0, x, (float*)xo, nullptr, 0, s, nsx()));
    } else {
      float* q = (float*)pl.ob;
      float* k = q + (size_t)N * L * C;
      float* v = k + (size_t)N * L * C;
      SFV_TRY(conv_f32(e->q, pl.oa, SRC_NHWC_F32, N, h, w, 1, 0, 0, nullptr, q, 0, 1.f, s));
      SFV_TRY(conv_f32(e->k, pl.oa, SRC_NHWC_F32, N, h, w, 1, 0, 0, nullptr, k, 0, 1.f, s));
      SFV_TRY(conv_f32(e->v, pl.oa, SRC_NHWC_F32, N, h, w, 1, 0, 0, nullptr, v, 0, 1.f, s));
      float* O = (float*)pl.oa;
      for (int n0 = 0; n0 < N; n0 += pl.attn_chunk) {
        const int nn = (N - n0) < pl.attn_chunk ? (N - n0) : pl.attn_chunk;
        SFV_TRY(attention_f32(q + (size_t)n0 * L * C, k + (size_t)n0 * L * C, v + (size_t)n0 * L * C,
                              O + (size_t)n0 * L * C, pl.S, nn, L, C, scale, s));
      }
      SFV_TRY(conv_f32(e->proj, O, SRC_NHWC_F32, N, h, w, 1, 0, 0, (const float*)x, (float*)xo, 0, 1.f, s));
    }
    return 0;
  }
};

}  // namespace


// Tensor-core attention core for a chunk of images whose score matrices fit S/P:
//   S = scale * Q K^T (fp32) ; P = softmax(S) (16-bit) ; O = P V + b_v (16-bit)
// q16/k16: [N][L][*] rows with pitch q_ld/k_ld elements; vT16: [N][C][Lp], S/P: [N][L][Lp], Lp = L rounded up to 8.
int attention_tc(int fmt, const void* q16, long long q_ld, const void* k16, long long k_ld,
                 const void* vT16, const float* v_bias, float* S, void* P, void* O16, int N, int L, int C,
                 float scale, cudaStream_t s) {
  const int Lp = (L + 7) / 8 * 8;       // row pitch of S, P and V^T
  {
    TcGemmArgs a; memset(&a, 0, sizeof(a));
    a.a = q16; a.fmt = fmt; a.a_rank = 3;
    a.a_dims[0] = C; a.a_dims[1] = L; a.a_dims[2] = N;
    a.a_strides[1] = (unsigned long long)q_ld * 2; a.a_strides[2] = (unsigned long long)L * q_ld * 2;
    a.a_box[0] = 64; a.a_box[1] = 128; a.a_box[2] = 1;
    a.dim_x = 1; a.dim_y = -1; a.dim_n = 2;
    a.ntaps = 1; a.kchunks = ceil_div(C, 64);
    a.b = k16; a.b_rows = L; a.b_k = C; a.b_row_stride = (unsigned long long)k_ld * 2;
    a.b_batch_stride = (unsigned long long)L * k_ld * 2; a.b_batched = 1;
    a.BW = 128; a.BH = 1; a.Wo = L; a.Ho = 1; a.Nimg = N; a.Cout = L; a.block_n = 256;
    a.alpha = scale; a.ldo = Lp;
    if (g_attn_fused) {            // P = softmax(scale * Q K^T) straight from the GEMM: the fp32 scores stay on chip
      a.out_16 = P; a.softmax_mode = 1;
      SFV_TRY(launch_tc_gemm(a, s));
    } else {
      a.out_f32 = S;
      SFV_TRY(launch_tc_gemm(a, s));
      SFV_TRY(launch_softmax_rows(S, P, 1, fmt, (long long)N * L, L, s, Lp));
    }
  }
  {
    TcGemmArgs a; memset(&a, 0, sizeof(a));
    a.a = P; a.fmt = fmt; a.a_rank = 3;
    a.a_dims[0] = L; a.a_dims[1] = L; a.a_dims[2] = N;
    a.a_strides[1] = (unsigned long long)Lp * 2; a.a_strides[2] = (unsigned long long)L * Lp * 2;
    a.a_box[0] = 64; a.a_box[1] = 128; a.a_box[2] = 1;
    a.dim_x = 1; a.dim_y = -1; a.dim_n = 2;
    a.ntaps = 1; a.kchunks = ceil_div(L, 64);
    a.b = vT16; a.b_rows = C; a.b_k = L; a.b_row_stride = (unsigned long long)Lp * 2;
    a.b_batch_stride = (unsigned long long)C * Lp * 2; a.b_batched = 1;
    a.BW = 128; a.BH = 1; a.Wo = L; a.Ho = 1; a.Nimg = N; a.Cout = C; a.block_n = pick_block_n(C);
    a.alpha = 1.f; a.bias = v_bias; a.out_16 = O16; a.ldo = C;
    SFV_TRY(launch_tc_gemm(a, s));
  }
  return 0;
}

// V^T[n][c][token] = sum_k Wv[c][k] x[n][token][k]: the value projection computed
// directly in the K-major layout the P V GEMM needs as its B operand.
int vT_tc(const ConvW& v, TcFmt fmt, const void* x16, void* vT16, int N, int L, cudaStream_t s) {
  const int C = v.Cin;
  TcGemmArgs a; memset(&a, 0, sizeof(a));
  a.a = v.w16; a.fmt = fmt.a; a.fmt_split = 1; a.fmt_b = fmt.b; a.fmt_out = fmt.out; a.a_rank = 2;
  a.a_dims[0] = C; a.a_dims[1] = v.cout_pad; a.a_strides[1] = (unsigned long long)C * 2;
  a.a_box[0] = 64; a.a_box[1] = 128;
  a.dim_x = 1; a.dim_y = -1; a.dim_n = -1;
  a.ntaps = 1; a.kchunks = C / 64;
  a.b = x16; a.b_rows = L; a.b_k = C; a.b_row_stride = (unsigned long long)C * 2;
  a.b_batch_stride = (unsigned long long)L * C * 2; a.b_batched = 1;
  a.BW = 128; a.BH = 1; a.Wo = v.Cout; a.Ho = 1; a.Nimg = N; a.Cout = L; a.block_n = 256;
  a.alpha = 1.f / v.w_scale; a.out_16 = vT16; a.ldo = (L + 7) / 8 * 8;
  return launch_tc_gemm(a, s);
}

// fp32 attention core on the CUDA-core GEMM: S = scale q k^T, softmax (in place), O = P v
int attention_f32(const float* q, const float* k, const float* v, float* O, float* S, int N, int L, int C,
                  float scale, cudaStream_t s) {
  IgemmArgs a; memset(&a, 0, sizeof(a));
  a.x = q; a.src_kind = SRC_NHWC_F32; a.w = k; a.w_sk = 1; a.w_sn = C; a.w_batch = (long long)L * C;
  a.y = S; a.N = N; a.H = 1; a.W = L; a.Cin = C; a.Ho = 1; a.Wo = L; a.Cout = L;
  a.ksize = 1; a.stride = 1; a.pad = 0; a.alpha = scale; a.in_scale = 1.f; a.ldy = L;
  SFV_TRY(launch_igemm_f32(a, s));
  SFV_TRY(launch_softmax_rows(S, S, 0, 0, (long long)N * L, L, s));
  memset(&a, 0, sizeof(a));
  a.x = S; a.src_kind = SRC_NHWC_F32; a.w = v; a.w_sk = C; a.w_sn = 1; a.w_batch = (long long)L * C;
  a.y = O; a.N = N; a.H = 1; a.W = L; a.Cin = L; a.Ho = 1; a.Wo = L; a.Cout = C;
  a.ksize = 1; a.stride = 1; a.pad = 0; a.alpha = 1.f; a.in_scale = 1.f; a.ldy = C;
  return launch_igemm_f32(a, s);
}

size_t encoder_workspace(const SfvEncoder* e, int B, int H, int W) {
  const int Bc = B < e->chunk ? B : e->chunk;
  Arena ar(nullptr, 0);
  Plan p;
  make_plan(e->prec != SFV_PREC_F32, e->prec != SFV_PREC_F32 && e->stream16, Bc, H, W, ar, &p);
  return ar.off + 1024;
}

int encoder_forward(SfvEncoder* e, const void* x, int src_kind, int B, int H, int W, float* params,
                    float* logvar, float* stdv, float* var, void* ws, size_t ws_bytes, float* const* taps,
                    cudaStream_t s) {
  SFV_CHECK(B >= 1 && H >= 8 && W >= 8 && H % 8 == 0 && W % 8 == 0, "encoder: H, W must be multiples of 8 (got %dx%d)", H, W);
  SFV_CHECK(ws != nullptr && ws_bytes >= encoder_workspace(e, B, H, W), "encoder: workspace too small (%zu < %zu)",
            ws_bytes, encoder_workspace(e, B, H, W));
  {
    int dev = -1;
    SFV_CUDA(cudaGetDevice(&dev));
    SFV_CHECK(dev == e->device, "encoder: handle was created on device %d but device %d is current (one handle per device)",
              e->device, dev);
  }
  const bool tc = e->prec != SFV_PREC_F32;
  const int chunk = B < e->chunk ? B : e->chunk;
  const int h8 = H / 8, w8 = W / 8;
  const long long L = (long long)h8 * w8;
  static const long long tapC[SFV_NUM_TAPS] = {128, 128, 128, 256, 256, 512, 512, 512, 512, 128, 256, 512, 512, 512, 512, 8};
  for (int b0 = 0; b0 < B; b0 += chunk) {
    Fwd f; f.e = e; f.s = s; f.tc = tc; f.fmt = e->fmt; f.N = (B - b0) < chunk ? (B - b0) : chunk;
    Arena ar(ws, ws_bytes);
    make_plan(tc, tc && e->stream16, chunk, H, W, ar, &f.pl);
    f.s16 = tc && e->stream16;
    const int N = f.N;
    const bool s16 = f.s16;
    // taps are fp32 NHWC copies of block outputs (layer-wise parity); a 16-bit stream is widened (and un-scaled)
    auto tap = [&](int idx, const void* src, int hh, int ww, bool stream = true) -> int {
      if (!taps || !taps[idx]) return 0;
      const size_t per = (size_t)hh * ww * tapC[idx];
      if (stream && s16)
        return launch_16_to_f32_scaled(src, taps[idx] + (size_t)b0 * per, (long long)per * N, e->fmt, 1.f / e->xc_scale, s);
      SFV_CUDA(cudaMemcpyAsync(taps[idx] + (size_t)b0 * per, src, per * N * 4, cudaMemcpyDeviceToDevice, s));
      return 0;
    };
    // conv_in straight from the boundary layout (fp32 NCHW or uint8 HWC)
    const char* xin = (const char*)x + (size_t)b0 * 3 * H * W * (src_kind == SRC_NHWC_U8 ? 1 : 4);
    if (tc && e->fused_stats)       // every statistics slot of this forward, once
      SFV_CUDA(cudaMemsetAsync(f.pl.stat_slots, 0, sizeof(double) * f.pl.stat_stride * kStatSlots, s));
    double* const st_in = f.nsx();
    if (tc && src_kind == SRC_NHWC_U8 && e->conv_in.w16_u8 && e->conv_in_tc && W % 8 == 0 && ((uintptr_t)xin & 3) == 0)
      SFV_TRY(conv_in_tc(e->conv_in, e->fmt_w, (const unsigned char*)xin, N, H, W, s16 ? nullptr : (float*)f.pl.xa, st_in, s,
                         s16 ? f.pl.xa : nullptr, e->fmt, e->xc_scale));
    else
      SFV_TRY(launch_conv_in(xin, src_kind, e->conv_in.w32, e->conv_in.bias, (float*)f.pl.xa, st_in, N, H, W, s,
                             s16 ? f.pl.xa : nullptr, e->fmt, e->xc_scale));
    SFV_TRY(tap(0, f.pl.xa, H, W));
    void* cur = f.pl.xa; void* oth = f.pl.xb;
    // operand view of the stream for convs that read x directly: the stream itself (fp32 check mode, 16-bit stream) or
    // a 16-bit copy written next to the fp32 stream where one is needed
    const void* cur_op = (!tc || s16) ? (const void*)cur : nullptr;
    int ch = H, cw = W;
    for (int l = 0; l < 4; ++l) {
      for (int b = 0; b < 2; ++b) {
        // the block's output is consumed directly by a conv (downsample) only after block 1 of levels 0..2
        const bool want_copy = (b == 1 && l != 3);
        SFV_TRY(f.resblock(e->down[l][b], cur, cur_op, ch, cw, oth, want_copy, &cur_op));
        std::swap(cur, oth);
        SFV_TRY(tap(1 + l * 2 + b, cur, ch, cw));
      }
      if (l != 3) {
        // downsample output feeds down.(l+1).block.0, whose nin_shortcut (levels 1, 2) reads x directly
        const bool want_copy = tc && !s16 && e->down[l + 1][0].has_nin;
        SFV_TRY(f.conv_x(e->ds[l], cur_op, tc ? e->xc_scale : 1.f, ch, cw, 2, nullptr, oth, want_copy ? f.pl.x16b : nullptr));
        ch /= 2; cw /= 2;
        std::swap(cur, oth);
        cur_op = (!tc || s16) ? (const void*)cur : (want_copy ? (const void*)f.pl.x16b : nullptr);
        SFV_TRY(tap(9 + l, cur, ch, cw));
      }
    }
    SFV_TRY(f.resblock(e->mid1, cur, cur_op, ch, cw, oth, false, &cur_op));
    std::swap(cur, oth);
    SFV_TRY(tap(12, cur, ch, cw));
    SFV_TRY(f.attention(cur, ch, cw, oth));
    std::swap(cur, oth);
    SFV_TRY(tap(13, cur, ch, cw));
    SFV_TRY(f.resblock(e->mid2, cur, nullptr, ch, cw, oth, false, &cur_op));
    std::swap(cur, oth);
    SFV_TRY(tap(14, cur, ch, cw));
    SFV_TRY(f.gn_x(e->norm_out, cur, L, 1, f.pl.oa));
    SFV_TRY(f.conv(e->conv_out, f.pl.oa, ch, cw, 1, nullptr, f.pl.moments, nullptr, nullptr));
    SFV_TRY(tap(15, f.pl.moments, ch, cw, false));
    SFV_TRY(launch_head(f.pl.moments, params + (size_t)b0 * 8 * L, logvar + (size_t)b0 * 4 * L,
                        stdv ? stdv + (size_t)b0 * 4 * L : nullptr, var ? var + (size_t)b0 * 4 * L : nullptr, N,
                        (int)L, s));
  }
  return 0;
}

}  // namespace sfv
