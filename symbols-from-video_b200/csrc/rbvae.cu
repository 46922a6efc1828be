// rbvae.cu -- RBVAE encoder-half handle (percep: 256 channels / 4 LSTM layers,
// contrastive: 64 / 2).  Reference: models/percep_RBVAE/percep_RBVAE_model.py:46-68,
// 94-107,172-191; models/contrastive_RBVAE/contrastive_RBVAE_model.py:45-67,93-106,171-190.
//
// precision fp32: every layer on CUDA cores in fp32.  bf16 / fp16: the two C -> C stride-2 convs (97 % of the
// RBVAE FLOPs) run on the tcgen05 kernel with 16-bit operands and fp32 accumulation; conv.0, fc, the LSTM and the
// threshold stay fp32 -- the binary code is the sign of a small LSTM state.
#include "common.cuh"
#include "encoder.h"
#include <string.h>

namespace sfv {

static const SfvTensor* find_t(const SfvTensor* t, int n, const std::string& name) {
  for (int i = 0; i < n; ++i)
    if (t[i].name && name == t[i].name) return &t[i];
  return nullptr;
}

int rbvae_build(SfvRbvae* r, const SfvTensor* t, int n) {
  const SfvTensor* w0 = find_t(t, n, "encoder_cnn.conv.0.weight");
  const SfvTensor* fcw = find_t(t, n, "encoder_cnn.fc.weight");
  const SfvTensor* fcb = find_t(t, n, "encoder_cnn.fc.bias");
  if (!w0 || !fcw || !fcb) return fail(SFV_ERR_MISSING_KEY, "rbvae: missing encoder_cnn.conv.0 / fc tensors");
  r->channels = (int)w0->shape[0];
  if (w0->shape[1] != r->in_channels)
    return fail(SFV_ERR_MISSING_KEY, "rbvae: conv.0 expects %lld input channels, got %d",
                (long long)w0->shape[1], r->in_channels);
  r->L = (int)fcw->shape[0];
  int h = r->in_h, w = r->in_w;
  for (int i = 0; i < 3; ++i) { h = (h + 2 - 3) / 2 + 1; w = (w + 2 - 3) / 2 + 1; }
  r->fh = h; r->fw = w;
  const long long K = (long long)r->channels * h * w;
  if (fcw->shape[1] != K)
    return fail(SFV_ERR_MISSING_KEY, "rbvae: fc.in_features %lld does not match %d*%d*%d for a %dx%d input",
                (long long)fcw->shape[1], r->channels, h, w, r->in_h, r->in_w);
  const char* names[3] = {"encoder_cnn.conv.0", "encoder_cnn.conv.3", "encoder_cnn.conv.6"};
  ConvW* cw[3] = {&r->c0, &r->c1, &r->c2};
  int cin = r->in_channels;
  for (int i = 0; i < 3; ++i) {
    const SfvTensor* ww = find_t(t, n, std::string(names[i]) + ".weight");
    const SfvTensor* bb = find_t(t, n, std::string(names[i]) + ".bias");
    if (!ww || !bb || ww->shape[0] != r->channels || ww->shape[1] != cin || ww->shape[2] != 3)
      return fail(SFV_ERR_MISSING_KEY, "rbvae: missing/mis-shaped %s", names[i]);
    SFV_TRY(make_conv_from_host(r->blob, ww->host_data, bb->host_data, r->channels, cin, 3, r->fmt,
                                r->prec != SFV_PREC_F32 && i > 0, cw[i]));
    if (i == 0 && r->prec != SFV_PREC_F32 && cin == 3 && r->channels == 64) {
      // contrastive conv.0 (3 -> 64, stride 2) as a 1x1 GEMM over im2col rows [hi(x) 27 | lo(x) 27 | 0 10]: the weight
      // row is [w | w | 0], k = (r*3+s)*3 + c  (launch_rb_im2col builds the matching A rows)
      std::vector<float> w1((size_t)64 * 64, 0.f);
      for (int o = 0; o < 64; ++o)
        for (int c = 0; c < 3; ++c)
          for (int t = 0; t < 9; ++t) {
            const float v = ww->host_data[((size_t)o * 3 + c) * 9 + t];
            w1[(size_t)o * 64 + t * 3 + c] = v;
            w1[(size_t)o * 64 + 27 + t * 3 + c] = v;
          }
      SFV_TRY(make_conv_from_host(r->blob, w1.data(), bb->host_data, 64, 64, 1, r->fmt, true, &r->c0_tc));
      // "pixel pairing": two horizontally adjacent output pixels form one 256-byte row, the weight is block diagonal
      // diag(W, W).  Same bytes, but the GEMM runs as a 128-wide kernel with eight epilogue warps -- this layer is one
      // k-chunk deep, i.e. purely epilogue bound, and the zero blocks cost nothing on the tensor pipe.
      std::vector<float> w2((size_t)128 * 128, 0.f), b2(128);
      for (int o = 0; o < 64; ++o) {
        for (int k = 0; k < 64; ++k) {
          w2[(size_t)o * 128 + k] = w1[(size_t)o * 64 + k];
          w2[(size_t)(64 + o) * 128 + 64 + k] = w1[(size_t)o * 64 + k];
        }
        b2[o] = b2[64 + o] = bb->host_data[o];
      }
      SFV_TRY(make_conv_from_host(r->blob, w2.data(), b2.data(), 128, 128, 1, r->fmt, true, &r->c0_tc2));
    }
    cin = r->channels;
  }
  {  // fc weight: reference flatten order is (C,H,W) (nn.Flatten on NCHW); ours is NHWC
    std::vector<float> p((size_t)r->L * K);
    const int C = r->channels, HW = h * w;
    for (int l = 0; l < r->L; ++l)
      for (int c = 0; c < C; ++c)
        for (int q = 0; q < HW; ++q)
          p[(size_t)l * K + (size_t)q * C + c] = fcw->host_data[(size_t)l * K + (size_t)c * HW + q];
    SFV_TRY(r->blob.upload(p.data(), p.size() * 4, (void**)&r->fc_w));
    SFV_TRY(r->blob.upload(fcb->host_data, (size_t)r->L * 4, (void**)&r->fc_b));
  }
  int layers = 0;
  while (find_t(t, n, "encoder_rnn.lstm.weight_ih_l" + std::to_string(layers))) ++layers;
  if (layers < 1) return fail(SFV_ERR_MISSING_KEY, "rbvae: no encoder_rnn.lstm.weight_ih_l0");
  r->layers = layers;
  const int L = r->L;
  std::vector<float> wi((size_t)layers * 4 * L * L), wh((size_t)layers * 4 * L * L), bs((size_t)layers * 4 * L);
  for (int l = 0; l < layers; ++l) {
    const std::string sfx = "_l" + std::to_string(l);
    const SfvTensor* a = find_t(t, n, "encoder_rnn.lstm.weight_ih" + sfx);
    const SfvTensor* b = find_t(t, n, "encoder_rnn.lstm.weight_hh" + sfx);
    const SfvTensor* c = find_t(t, n, "encoder_rnn.lstm.bias_ih" + sfx);
    const SfvTensor* d = find_t(t, n, "encoder_rnn.lstm.bias_hh" + sfx);
    if (!a || !b || !c || !d || a->shape[0] != 4 * L || a->shape[1] != L || b->shape[0] != 4 * L || b->shape[1] != L)
      return fail(SFV_ERR_MISSING_KEY, "rbvae: missing/mis-shaped LSTM layer %d (hidden must equal latent_dim %d)", l, L);
    memcpy(&wi[(size_t)l * 4 * L * L], a->host_data, (size_t)4 * L * L * 4);
    memcpy(&wh[(size_t)l * 4 * L * L], b->host_data, (size_t)4 * L * L * 4);
    for (int j = 0; j < 4 * L; ++j) bs[(size_t)l * 4 * L + j] = c->host_data[j] + d->host_data[j];
  }
  SFV_TRY(r->blob.upload(wi.data(), wi.size() * 4, (void**)&r->w_ih));
  SFV_TRY(r->blob.upload(wh.data(), wh.size() * 4, (void**)&r->w_hh));
  SFV_TRY(r->blob.upload(bs.data(), bs.size() * 4, (void**)&r->lstm_b));
  return 0;
}

static void rb_dims(const SfvRbvae* r, int h[4], int w[4]) {
  h[0] = r->in_h; w[0] = r->in_w;
  for (int i = 1; i < 4; ++i) { h[i] = (h[i - 1] - 1) / 2 + 1; w[i] = (w[i - 1] - 1) / 2 + 1; }
}

size_t rbvae_workspace(const SfvRbvae* r, int N) {
  int h[4], w[4];
  rb_dims(r, h, w);
  Arena ar(nullptr, 0);
  ar.take((size_t)N * h[1] * w[1] * r->channels * 4);
  ar.take((size_t)N * h[2] * w[2] * r->channels * 4);
  int KS, splits;
  fc_plan((long long)h[3] * w[3] * r->channels, r->L, &KS, &splits);
  ar.take((size_t)splits * N * r->L * 4);       // fc split-K partial sums
  ar.take((size_t)N * r->L * 4);
  return ar.off + 1024;
}

int rbvae_encode(SfvRbvae* r, const float* x, int B, int T, float in_scale, const float* u, float noise_ratio,
                 float temperature, int hard, float* h_out, float* z_out, uint32_t* codes, void* ws,
                 size_t ws_bytes, cudaStream_t s) {
  const int N = B * T;
  SFV_CHECK(B >= 1 && T >= 1, "rbvae: empty batch");
  SFV_CHECK(noise_ratio == 0.f || u != nullptr, "rbvae: noise_ratio != 0 needs the uniform draws U");
  SFV_CHECK(temperature > 0.f, "rbvae: temperature must be > 0");
  SFV_CHECK(ws && ws_bytes >= rbvae_workspace(r, N), "rbvae: workspace too small");
  int h[4], w[4];
  rb_dims(r, h, w);
  Arena ar(ws, ws_bytes);
  float* a1 = (float*)ar.take((size_t)N * h[1] * w[1] * r->channels * 4);
  float* a2 = (float*)ar.take((size_t)N * h[2] * w[2] * r->channels * 4);
  int KS, splits;
  const long long Kfc = (long long)h[3] * w[3] * r->channels;
  fc_plan(Kfc, r->L, &KS, &splits);
  float* partial = (float*)ar.take((size_t)splits * N * r->L * 4);
  float* hbuf = (float*)ar.take((size_t)N * r->L * 4);
  // frames may exceed the grid.z limit of the GEMM kernel -> slices of 32768
  for (int n0 = 0; n0 < N; n0 += 32768) {
    const int nn = (N - n0) < 32768 ? (N - n0) : 32768;
    const float* xi = x + (size_t)n0 * r->in_channels * h[0] * w[0];
    float* a1i = a1 + (size_t)n0 * h[1] * w[1] * r->channels;
    float* a2i = a2 + (size_t)n0 * h[2] * w[2] * r->channels;
    // conv(s2,p1)+ReLU, conv(s2,p1)+ReLU, conv(s2,p1)   (dropout is identity in eval)
    const bool tc = r->prec != SFV_PREC_F32 && r->c1.w16 && r->c2.w16 && h[1] % 2 == 0 && w[1] % 2 == 0 &&
                    h[2] % 2 == 0 && w[2] % 2 == 0;
    // contrastive first layer (3 -> 64 on full-resolution frames) has its own write-bound kernel
    const bool c0_direct = r->in_channels == 3 && r->channels == 64;
    if (tc) {
      // 16-bit operands for the two C->C convs (97 % of the RBVAE FLOPs) on the tcgen05 kernel;
      // conv.0 (Cin = 3 or 4) stays on CUDA cores and emits the 16-bit operand directly
      const TcFmt cf{r->fmt_act, r->fmt, r->fmt_act};
      const int chk = r->range_check ? 1 : 0;
      static const bool c0_tc_on = []() { const char* e = getenv("SFV_RB_CONV0_TC"); return !(e && atoi(e) == 0); }();
      if (c0_direct && r->c0_tc.w16 && c0_tc_on) {
        // conv.0 on the tensor pipe: im2col rows (hi | lo split of the fp32 input), then a 1x1 GEMM + bias + ReLU IN PLACE
        // over those rows (a tile's 64-channel output row is its own 128-byte input row)
        SFV_TRY(launch_rb_im2col(xi, a1i, r->fmt_act, nn, h[0], w[0], in_scale, s));
        if (w[1] % 2 == 0 && r->c0_tc2.w16)
          SFV_TRY(conv_tc(r->c0_tc2, cf, a1i, nn, h[1], w[1] / 2, 1, 0, 0, nullptr, nullptr, a1i, 1, s, nullptr, nullptr, 1.f,
                          1.f, 0, 1.f, chk));
        else
          SFV_TRY(conv_tc(r->c0_tc, cf, a1i, nn, h[1], w[1], 1, 0, 0, nullptr, nullptr, a1i, 1, s, nullptr, nullptr, 1.f, 1.f,
                          0, 1.f, chk));
      } else if (c0_direct) {
        SFV_TRY(launch_rb_conv0(xi, r->c0.w32, r->c0.bias, a1i, 1, r->fmt_act, nn, h[0], w[0], in_scale, s));
      } else {
        SFV_TRY(conv_f32(r->c0, xi, SRC_NCHW_F32, nn, h[0], w[0], 2, 1, 1, nullptr, nullptr, 1, in_scale, s, a1i, r->fmt_act,
                         chk));
      }
      SFV_TRY(conv_tc(r->c1, cf, a1i, nn, h[1], w[1], 2, 1, 1, nullptr, nullptr, a2i, 1, s, nullptr, nullptr, 1.f, 1.f, 0, 1.f, chk));
      SFV_TRY(conv_tc(r->c2, cf, a2i, nn, h[2], w[2], 2, 1, 1, nullptr, a1i, nullptr, 0, s));
    } else {
      if (c0_direct) SFV_TRY(launch_rb_conv0(xi, r->c0.w32, r->c0.bias, a1i, 0, r->fmt, nn, h[0], w[0], in_scale, s));
      else SFV_TRY(conv_f32(r->c0, xi, SRC_NCHW_F32, nn, h[0], w[0], 2, 1, 1, nullptr, a1i, 1, in_scale, s));
      SFV_TRY(conv_f32(r->c1, a1i, SRC_NHWC_F32, nn, h[1], w[1], 2, 1, 1, nullptr, a2i, 1, 1.f, s));
      // third conv output reuses a1 (dead after conv 2)
      SFV_TRY(conv_f32(r->c2, a2i, SRC_NHWC_F32, nn, h[2], w[2], 2, 1, 1, nullptr, a1i, 0, 1.f, s));
    }
    SFV_TRY(launch_fc(a1i, r->fc_w, partial + (size_t)n0 * r->L, nn, N, Kfc, r->L, s));
  }
  float* hb = h_out ? h_out : hbuf;
  // fc partial sums + bias are reduced inside the LSTM kernel's load: fc -> LSTM -> threshold -> pack, no logits tensor
  return launch_lstm_code(partial, splits, r->fc_b, B, T, r->L, r->layers, r->w_ih, r->w_hh, r->lstm_b, u, noise_ratio,
                          temperature, hard, hb, z_out, codes, s);
}

}  // namespace sfv
