// encoder.h -- handle structs shared by encoder.cu / rbvae.cu / api.cu
#pragma once
#include "common.cuh"

namespace sfv {

struct DeviceBlob {
  std::vector<void*> allocs;
  int upload(const void* host, size_t bytes, void** out);
  void release();
};

struct ConvW {
  int Cin = 0, Cout = 0, ks = 0, cout_pad = 0;
  int extra_k = 0;         // channels of a fused 1x1 branch appended to K (nin_shortcut folded into conv2)
  float* w32 = nullptr;    // [ks*ks*Cin][Cout] fp32 (CUDA-core kernel)
  void* w16 = nullptr;     // [cout_pad][ks*ks*Cin] 16-bit, K-major (UMMA B operand)
  float* bias = nullptr;   // [cout_pad]
  float w_scale = 1.f;     // w16 holds w * w_scale (a power of two chosen per layer for fp16 weights); the epilogue's alpha undoes it
  void* w16_u8 = nullptr;  // conv_in only: [128][64] 16-bit, k<27: hi(w*4096/255), 27..53: lo, for the uint8-fed tcgen05 path
};
struct NormW { float* gamma = nullptr; float* beta = nullptr; int C = 0; };
struct ResW { NormW n1, n2; ConvW c1, c2, nin, c2n; bool has_nin = false; };   // c2n: conv2 with nin_shortcut fused along K

// power-of-two scale that puts max|w| in [2^13, 2^14) for fp16 weights (keeps small weights out of the subnormals), 1 for bf16
float weight_scale_for(int fmt, double maxabs);
int make_conv_from_host(DeviceBlob& blob, const float* w, const float* b, int Cout, int Cin, int ks,
                        int fmt, bool want16, ConvW* out);
// 16-bit formats of one tensor-core GEMM: A operand, B operand, 16-bit output
struct TcFmt { int a, b, out; };
// GroupNorm(32)(+SiLU) to be applied to the conv's raw 16-bit input inside the kernel (TcGemmArgs::xf_*)
struct XfIn { const double* stats; const float* gamma; const float* beta; float in_mul; int silu; int check; };
inline TcFmt tcfmt(int f) { return TcFmt{f, f, f}; }
int make_conv_in_u8(DeviceBlob& blob, const float* w_oihw, int fmt, ConvW* out);
int conv_in_tc(const ConvW& w, int fmt, const unsigned char* u8, int N, int H, int W, float* out_f32, double* gn_stats,
               cudaStream_t s, void* out_16 = nullptr, int fmt_out = 0, float out16_scale = 1.f);
// in_scale: the A tensor holds in_scale * (true input) (MIXED mode x copies); out16_scale: out_16 = to16(out16_scale * y)
// residual: fp32, or (res16) a 16-bit tensor in fmt.out added as res_mul * value (the scaled 16-bit residual stream)
int conv_tc(const ConvW& w, TcFmt fmt, const void* in16, int N, int H, int W, int stride, int pad_lo,
            int pad_hi, const void* residual, float* out_f32, void* out_16, int relu, cudaStream_t s,
            double* gn_stats = nullptr, const void* a2_16 = nullptr, float in_scale = 1.f, float out16_scale = 1.f,
            int res16 = 0, float res_mul = 1.f, int sat_check = 0, const XfIn* xf = nullptr);
inline int conv_tc(const ConvW& w, int fmt, const void* in16, int N, int H, int W, int stride, int pad_lo,
                   int pad_hi, const float* residual, float* out_f32, void* out_16, int relu, cudaStream_t s,
                   double* gn_stats = nullptr, const void* a2_16 = nullptr) {
  return conv_tc(w, tcfmt(fmt), in16, N, H, W, stride, pad_lo, pad_hi, residual, out_f32, out_16, relu, s, gn_stats, a2_16);
}
int conv_f32(const ConvW& w, const void* in, int src_kind, int N, int H, int W, int stride, int pad_lo,
             int pad_hi, const float* residual, float* out, int relu, float in_scale, cudaStream_t s,
             void* out16 = nullptr, int fmt16 = 0, int range_check16 = 0);
int attention_f32(const float* q, const float* k, const float* v, float* O, float* S, int N, int L, int C,
                  float scale, cudaStream_t s);
// fmt.a: format of the weights (they are the A operand here), fmt.b: of x16, fmt.out: of V^T
int vT_tc(const ConvW& v, TcFmt fmt, const void* x16, void* vT16, int N, int L, cudaStream_t s);
inline int vT_tc(const ConvW& v, int fmt, const void* x16, void* vT16, int N, int L, cudaStream_t s) {
  return vT_tc(v, tcfmt(fmt), x16, vT16, N, L, s);
}
int attention_tc(int fmt, const void* q16, long long q_ld, const void* k16, long long k_ld,
                 const void* vT16, const float* v_bias, float* S, void* P, void* O16, int N, int L, int C,
                 float scale, cudaStream_t s);

}  // namespace sfv

struct SfvEncoder {
  int prec = 0, fmt = 0, chunk = 16;
  int device = -1;             // CUDA device the weights live on; forward calls must run with it current
  // 16-bit operand formats (see SfvPrecision in sfv.h): fmt = activations written by GroupNorm / conv1 / x copies,
  // fmt_w = weights, fmt_attn = q, k, V^T, P and the attention output; xc_scale = scale of the 16-bit x copies
  int fmt_w = 0, fmt_attn = 0;
  float xc_scale = 1.f;
  bool range_check = false;    // MIXED: fp16 store sites are range-checked on the device
  bool stream16 = false;       // MIXED: the residual stream x itself is stored as fp16 * xc_scale (no fp32 copy of x in HBM)
  bool gn_fuse = false;        // 16-bit stream: GroupNorm+SiLU applied inside the consuming conv where a transform variant exists
  bool fuse_nin = true;        // nin_shortcut folded into conv2's GEMM (tensor-core modes)
  bool fused_stats = true;     // GroupNorm statistics from the producing kernel's epilogue (tensor-core modes)
  bool conv_in_tc = true;      // uint8-fed conv_in on the tensor pipe (tensor-core modes; SFV_CONV_IN_TC=0: CUDA cores)
  sfv::DeviceBlob blob;
  sfv::ConvW conv_in, ds[3], q, k, v, qk, proj, conv_out;
  sfv::ResW down[4][2], mid1, mid2;
  sfv::NormW attn_norm, norm_out;
};

struct SfvRbvae {
  int device = -1;
  int in_channels = 0, in_h = 0, in_w = 0, channels = 0, layers = 0, L = 0;
  int prec = 0, fmt = 0;              // SFV_PREC_F32: all fp32; BF16/FP16/MIXED: the two C->C stride-2 convs on tcgen05
  int fmt_act = 0;                    // fmt = weight format, fmt_act = activation format (MIXED: fp16 weights, bf16 activations)
  int fh = 0, fw = 0;                 // feature map after the three stride-2 convs
  sfv::DeviceBlob blob;
  sfv::ConvW c0, c1, c2;
  sfv::ConvW c0_tc;                   // contrastive conv.0 as a 1x1 GEMM over im2col rows: [64][w | w | 0] (K = 64)
  sfv::ConvW c0_tc2;                  // the same over pixel PAIRS: [128][128] = diag(c0_tc, c0_tc)
  bool range_check = false;           // MIXED: fp16 activation stores are range-checked on the device
  float* fc_w = nullptr;              // [L][fh*fw*channels], permuted to NHWC flatten order
  float* fc_b = nullptr;
  float *w_ih = nullptr, *w_hh = nullptr, *lstm_b = nullptr;   // [layers][4L][L], [layers][4L]
};

// decoder half (training-side forward): decoder_rnn + ConvDecoder, fp32 on CUDA cores
struct SfvRbvaeDecoder {
  int device = -1;
  int out_channels = 0, out_h = 0, out_w = 0, channels = 0, layers = 0, L = 0;
  int fh = 0, fw = 0;                 // feature map the fc layer produces (out / 8)
  sfv::DeviceBlob blob;
  float* ph_w[3][4] = {};             // the three ConvTranspose2d as four sub-pixel phase kernels each: [(1+a)(1+b)*Cin][Cout]
  float* dc_bias[3] = {};
  int dc_cout[3] = {};
  float* fc_w = nullptr;              // [channels*fh*fw][L] (reference layout)
  float* fc_b = nullptr;
  float *w_ih = nullptr, *w_hh = nullptr, *lstm_b = nullptr;   // decoder_rnn: [layers][4L][L], [layers][4L]
};

namespace sfv {
int rbvae_decoder_build(SfvRbvaeDecoder* r, const SfvTensor* t, int n);
size_t rbvae_decoder_workspace(const SfvRbvaeDecoder* r, int N);
int rbvae_decode(SfvRbvaeDecoder* r, const float* z_seq, int B, int T, float* d_seq, float* x_recon, void* ws, size_t ws_bytes,
                 cudaStream_t s);
int launch_mse(const float* a, const float* b, long long n, float* out, cudaStream_t s);
int launch_l1(const float* q, long long n, float lamb, float* out, cudaStream_t s);
int launch_kl_binary_concrete(const float* q_logits, long long rows, int L, float p, float eps, float* out, cudaStream_t s);
int launch_contrast(const float* x1, const float* x2, const float* label, int rows, int D, float margin, int cosine, float* out,
                    cudaStream_t s);
int launch_triplet(const float* a, const float* p, const float* n, int rows, int D, float margin, float eps, int swap, float* out,
                   cudaStream_t s);
int encoder_build(SfvEncoder* e, const SfvTensor* t, int n);
size_t encoder_workspace(const SfvEncoder* e, int B, int H, int W);
int encoder_forward(SfvEncoder* e, const void* x, int src_kind, int B, int H, int W, float* params,
                    float* logvar, float* stdv, float* var, void* ws, size_t ws_bytes, float* const* taps,
                    cudaStream_t s);
}
