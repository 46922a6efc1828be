// decoder.cu -- RBVAE decoder half and the training losses: forward values only, fp32 on CUDA cores.
// Not on the precompute hot path: the reference runs these in its training / validation loops
// (models/percep_RBVAE/percep_RBVAE_train.py:526-527, 606).
//
//   decoder_rnn   nn.LSTM(L, L, layers) over z_seq              percep_RBVAE_model.py:110-122,160
//   ConvDecoder   fc(L -> C*fh*fw), reshape (C, fh, fw), 3 x ConvTranspose2d(3, stride 2, pad 1, output_padding 1)
//                 with ReLU between, Sigmoid at the end         percep_RBVAE_model.py:71-91,163-166
//                 (contrastive: 64 channels, 2-layer LSTM        contrastive_RBVAE_model.py:70-90,109-121)
//   losses        l1_loss, recon_loss, triplet_loss, kl_binary_concrete, contrast_loss
//                                                               percep_RBVAE_train.py:27-107
//
// ConvTranspose2d(k=3, s=2, p=1, output_padding=1) by sub-pixel phases: output pixel (2y+a, 2x+b) only sees the taps
// whose parity matches -- a = 0: ky = 1 on input row y; a = 1: ky = 2 on row y and ky = 0 on row y+1 (likewise in x).
// Each of the four phases is a small dense convolution (1x1, 1x2, 2x1, 2x2 taps, 9 in total instead of the 36 a
// zero-stuffed 3x3 convolution would execute) run by the fp32 implicit-GEMM kernel with a strided output.
#include "common.cuh"
#include "encoder.h"
#include <string.h>

namespace sfv {
namespace {

const SfvTensor* find_d(const SfvTensor* t, int n, const std::string& name) {
  for (int i = 0; i < n; ++i)
    if (t[i].name && name == t[i].name) return &t[i];
  return nullptr;
}

// h[n][(y*fw + x)*C + c] = bias[j] + sum_k d[n][k] * w[j][k],  j = c*fh*fw + y*fw + x  (the reference reshapes the fc
// output to (C, fh, fw); ours is NHWC)
__global__ void fc_decode_kernel(const float* __restrict__ d, const float* __restrict__ w, const float* __restrict__ b,
                                 float* __restrict__ out, int N, int L, int C, int HW) {
  const long long F = (long long)C * HW;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * F) return;
  const int n = (int)(i / F);
  const long long r = i - (long long)n * F;       // NHWC offset inside the frame
  const int q = (int)(r / C), c = (int)(r - (long long)q * C);
  const long long j = (long long)c * HW + q;
  const float* wr = w + j * L;
  const float* dr = d + (long long)n * L;
  float a = 0.f;
  for (int k = 0; k < L; ++k) a = fmaf(wr[k], dr[k], a);
  out[i] = a + b[j];
}

// NHWC conv output -> sigmoid -> NCHW  (nn.Sigmoid at the end of ConvDecoder.deconv)
__global__ void sigmoid_nchw_kernel(const float* __restrict__ in, float* __restrict__ out, int C, long long HW, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // NCHW index
  if (i >= total) return;
  const long long q = i % HW;
  const long long nc = i / HW;
  const int c = (int)(nc % C);
  const long long n = nc / C;
  const float v = in[(n * HW + q) * C + c];
  out[i] = 1.f / (1.f + expf(-v));
}

// ---- losses: double accumulators, fixed summation order (deterministic) --------------------------------------------
__device__ double block_sum(double v) {
  __shared__ double sh[32];
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x < 32) {
    t = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  return t;       // valid in thread 0
}

// The element-wise losses run as ONE cluster of eight CTAs: every CTA reduces its stride of the data, the partial sums
// meet in CTA 0 through distributed shared memory in rank order (deterministic, no scratch buffer, no atomics).
constexpr int kLossCtas = 8;
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// block_sum over the block, then over the cluster; the total is valid in thread 0 of CTA 0
__device__ double cluster_sum(double v) {
  __shared__ double part;
  v = block_sum(v);
  if (threadIdx.x == 0) part = v;
  cluster_barrier();
  double t = 0.0;
  if (cluster_rank() == 0 && threadIdx.x == 0) {
    const uint32_t local = (uint32_t)__cvta_generic_to_shared(&part);
    for (uint32_t r = 0; r < (uint32_t)kLossCtas; ++r) {
      uint32_t remote;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(r));
      double pv;
      asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(pv) : "r"(remote) : "memory");
      t += pv;
    }
  }
  cluster_barrier();            // nobody leaves while CTA 0 may still read its shared memory
  return t;
}

// F.mse_loss(a, b): mean of squared differences
__global__ void __cluster_dims__(kLossCtas, 1, 1) mse_kernel(const float* a, const float* b, long long n, float* out) {
  double s = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)kLossCtas * blockDim.x) {
    const float d = a[i] - b[i];
    s += (double)(d * d);
  }
  s = cluster_sum(s);
  if (blockIdx.x == 0 && threadIdx.x == 0) *out = (float)(s / (double)n);
}
// lamb * torch.norm(q, p=1)
__global__ void __cluster_dims__(kLossCtas, 1, 1) l1_kernel(const float* q, long long n, float lamb, float* out) {
  double s = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)kLossCtas * blockDim.x)
    s += (double)fabsf(q[i]);
  s = cluster_sum(s);
  if (blockIdx.x == 0 && threadIdx.x == 0) *out = lamb * (float)s;
}
// kl_binary_concrete (percep_RBVAE_train.py:52-77): q = clamp(sigmoid(logit), eps, 1 - eps);
// kl = q (log(q + eps) - log p) + (1 - q)(log(1 - q + eps) - log(1 - p)); sum over the latent dim, mean over the rest
__global__ void __cluster_dims__(kLossCtas, 1, 1) kl_kernel(const float* q_logits, long long rows, int L, float log_p, float log_1mp,
                                                             float eps, float* out) {
  double s = 0.0;
  const long long n = rows * L;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)kLossCtas * blockDim.x) {
    float q = 1.f / (1.f + expf(-q_logits[i]));
    q = fminf(fmaxf(q, eps), 1.0f - eps);
    const float kl = q * (logf(q + eps) - log_p) + (1.0f - q) * (logf((1.0f - q) + eps) - log_1mp);
    s += (double)kl;
  }
  s = cluster_sum(s);
  if (blockIdx.x == 0 && threadIdx.x == 0) *out = (float)(s / (double)rows);
}
// F.pairwise_distance(x, y, p=2, eps): || x - y + eps ||_2 per row
__device__ float row_dist(const float* x, const float* y, int D, float eps) {
  float s = 0.f;
  for (int k = 0; k < D; ++k) { const float d = x[k] - y[k] + eps; s = fmaf(d, d, s); }
  return sqrtf(s);
}
// contrast_loss (percep_RBVAE_train.py:80-107): dist = pairwise_distance or 1 - cosine_similarity;
// mean((1 - label) dist^2 + label clamp(margin - dist, 0)^2)
__global__ void contrast_kernel(const float* x1, const float* x2, const float* label, int rows, int D, float margin, int cosine,
                                float* out) {
  double s = 0.0;
  for (int r = threadIdx.x; r < rows; r += blockDim.x) {
    const float* a = x1 + (long long)r * D;
    const float* b = x2 + (long long)r * D;
    float dist;
    if (cosine) {
      // F.cosine_similarity(dim=1, eps=1e-8): (a . b) / (max(||a||, eps) * max(||b||, eps))
      float ab = 0.f, aa = 0.f, bb = 0.f;
      for (int k = 0; k < D; ++k) { ab = fmaf(a[k], b[k], ab); aa = fmaf(a[k], a[k], aa); bb = fmaf(b[k], b[k], bb); }
      dist = 1.f - ab / (fmaxf(sqrtf(aa), 1e-8f) * fmaxf(sqrtf(bb), 1e-8f));
    } else {
      dist = row_dist(a, b, D, 1e-6f);
    }
    const float lb = label[r];
    const float m = fmaxf(margin - dist, 0.f);
    s += (double)((1.f - lb) * dist * dist + lb * m * m);
  }
  s = block_sum(s);
  if (threadIdx.x == 0) *out = (float)(s / (double)rows);
}
// F.triplet_margin_loss(anchor, pos, neg, margin, p=2, eps, swap, reduction='mean')
__global__ void triplet_kernel(const float* a, const float* p, const float* n, int rows, int D, float margin, float eps, int swap,
                               float* out) {
  double s = 0.0;
  for (int r = threadIdx.x; r < rows; r += blockDim.x) {
    const float* ar = a + (long long)r * D;
    const float* pr = p + (long long)r * D;
    const float* nr = n + (long long)r * D;
    const float dap = row_dist(ar, pr, D, eps);
    float dan = row_dist(ar, nr, D, eps);
    if (swap) dan = fminf(dan, row_dist(pr, nr, D, eps));
    s += (double)fmaxf(margin + dap - dan, 0.f);
  }
  s = block_sum(s);
  if (threadIdx.x == 0) *out = (float)(s / (double)rows);
}

}  // namespace

int rbvae_decoder_build(SfvRbvaeDecoder* r, const SfvTensor* t, int n) {
  const SfvTensor* fcw = find_d(t, n, "decoder_cnn.fc.weight");
  const SfvTensor* fcb = find_d(t, n, "decoder_cnn.fc.bias");
  const SfvTensor* w0 = find_d(t, n, "decoder_cnn.deconv.0.weight");
  if (!fcw || !fcb || !w0) return fail(SFV_ERR_MISSING_KEY, "rbvae decoder: missing decoder_cnn.fc / deconv.0 tensors");
  r->L = (int)fcw->shape[1];
  r->channels = (int)w0->shape[0];                       // ConvTranspose2d weight is [Cin][Cout][3][3]
  SFV_CHECK(r->out_h % 8 == 0 && r->out_w % 8 == 0, "rbvae decoder: output %dx%d is not a multiple of 8", r->out_h, r->out_w);
  r->fh = r->out_h / 8; r->fw = r->out_w / 8;
  const long long F = (long long)r->channels * r->fh * r->fw;
  if (fcw->shape[0] != F || fcb->shape[0] != F)
    return fail(SFV_ERR_MISSING_KEY, "rbvae decoder: fc.out_features %lld does not match %d*%d*%d for a %dx%d output",
                (long long)fcw->shape[0], r->channels, r->fh, r->fw, r->out_h, r->out_w);
  SFV_CHECK(r->L >= 1 && r->L <= 256, "rbvae decoder: latent_dim %d out of range [1,256]", r->L);
  SFV_TRY(r->blob.upload(fcw->host_data, (size_t)F * r->L * 4, (void**)&r->fc_w));
  SFV_TRY(r->blob.upload(fcb->host_data, (size_t)F * 4, (void**)&r->fc_b));
  const char* names[3] = {"decoder_cnn.deconv.0", "decoder_cnn.deconv.3", "decoder_cnn.deconv.6"};
  for (int i = 0; i < 3; ++i) {
    const SfvTensor* ww = find_d(t, n, std::string(names[i]) + ".weight");
    const SfvTensor* bb = find_d(t, n, std::string(names[i]) + ".bias");
    const int cin = r->channels, cout = i == 2 ? r->out_channels : r->channels;
    if (!ww || !bb || ww->shape[0] != cin || ww->shape[1] != cout || ww->shape[2] != 3 || ww->shape[3] != 3)
      return fail(SFV_ERR_MISSING_KEY, "rbvae decoder: missing/mis-shaped %s", names[i]);
    // phase (a, b), tap (r, s) of its (1+a) x (1+b) kernel reads input (y + r, x + s) with the transposed-convolution
    // weight [ci][co][ky][kx], ky = a ? (r ? 0 : 2) : 1, kx likewise; igemm layout [(r*kw + s)*Cin + ci][Cout]
    r->dc_cout[i] = cout;
    SFV_TRY(r->blob.upload(bb->host_data, (size_t)cout * 4, (void**)&r->dc_bias[i]));
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b) {
        const int kh = 1 + a, kw = 1 + b;
        std::vector<float> w((size_t)kh * kw * cin * cout);
        for (int rr = 0; rr < kh; ++rr)
          for (int ss = 0; ss < kw; ++ss) {
            const int ky = a ? (rr ? 0 : 2) : 1, kx = b ? (ss ? 0 : 2) : 1;
            for (int ci = 0; ci < cin; ++ci)
              for (int co = 0; co < cout; ++co)
                w[((size_t)(rr * kw + ss) * cin + ci) * cout + co] = ww->host_data[(((size_t)ci * cout + co) * 3 + ky) * 3 + kx];
          }
        SFV_TRY(r->blob.upload(w.data(), w.size() * 4, (void**)&r->ph_w[i][a * 2 + b]));
      }
  }
  int layers = 0;
  while (find_d(t, n, "decoder_rnn.lstm.weight_ih_l" + std::to_string(layers))) ++layers;
  if (layers < 1) return fail(SFV_ERR_MISSING_KEY, "rbvae decoder: no decoder_rnn.lstm.weight_ih_l0");
  r->layers = layers;
  const int L = r->L;
  std::vector<float> wi((size_t)layers * 4 * L * L), wh((size_t)layers * 4 * L * L), bs((size_t)layers * 4 * L);
  for (int l = 0; l < layers; ++l) {
    const std::string sfx = "_l" + std::to_string(l);
    const SfvTensor* a = find_d(t, n, "decoder_rnn.lstm.weight_ih" + sfx);
    const SfvTensor* b = find_d(t, n, "decoder_rnn.lstm.weight_hh" + sfx);
    const SfvTensor* c = find_d(t, n, "decoder_rnn.lstm.bias_ih" + sfx);
    const SfvTensor* d = find_d(t, n, "decoder_rnn.lstm.bias_hh" + sfx);
    if (!a || !b || !c || !d || a->shape[0] != 4 * L || a->shape[1] != L || b->shape[0] != 4 * L || b->shape[1] != L)
      return fail(SFV_ERR_MISSING_KEY, "rbvae decoder: missing/mis-shaped LSTM layer %d (hidden must equal latent_dim %d)", l, L);
    memcpy(&wi[(size_t)l * 4 * L * L], a->host_data, (size_t)4 * L * L * 4);
    memcpy(&wh[(size_t)l * 4 * L * L], b->host_data, (size_t)4 * L * L * 4);
    for (int j = 0; j < 4 * L; ++j) bs[(size_t)l * 4 * L + j] = c->host_data[j] + d->host_data[j];
  }
  SFV_TRY(r->blob.upload(wi.data(), wi.size() * 4, (void**)&r->w_ih));
  SFV_TRY(r->blob.upload(wh.data(), wh.size() * 4, (void**)&r->w_hh));
  SFV_TRY(r->blob.upload(bs.data(), bs.size() * 4, (void**)&r->lstm_b));
  return 0;
}

// frames that go through the transposed convolutions together (bounds the workspace: ~4.5 MB per 88x160 frame)
static size_t dec_bytes_a(const SfvRbvaeDecoder* r, int n) {      // fc output, second layer's output
  return (size_t)n * (r->out_h / 2) * (r->out_w / 2) * r->channels * 4;
}
static size_t dec_bytes_b(const SfvRbvaeDecoder* r, int n) {      // first layer's output, last layer's output
  const size_t l0 = (size_t)n * (r->out_h / 4) * (r->out_w / 4) * r->channels * 4;
  const size_t l2 = (size_t)n * r->out_h * r->out_w * r->out_channels * 4;
  return l0 > l2 ? l0 : l2;
}
static int dec_slice(const SfvRbvaeDecoder* r, int N) {
  const long long per = (long long)(dec_bytes_a(r, 1) + dec_bytes_b(r, 1));
  long long k = (1ll << 30) / (per > 0 ? per : 1);
  if (k < 1) k = 1;
  if (k > 4096) k = 4096;
  return (int)(k < N ? k : N);
}

size_t rbvae_decoder_workspace(const SfvRbvaeDecoder* r, int N) {
  const int ns = dec_slice(r, N);
  Arena ar(nullptr, 0);
  ar.take((size_t)N * r->L * 4);                                              // d_seq when the caller does not want it
  ar.take(dec_bytes_a(r, ns));
  ar.take(dec_bytes_b(r, ns));
  return ar.off + 1024;
}

int rbvae_decode(SfvRbvaeDecoder* r, const float* z_seq, int B, int T, float* d_seq, float* x_recon, void* ws, size_t ws_bytes,
                 cudaStream_t s) {
  const int N = B * T;
  SFV_CHECK(B >= 1 && T >= 1, "rbvae decoder: empty batch");
  SFV_CHECK(z_seq && x_recon, "rbvae decoder: null argument");
  SFV_CHECK(ws && ws_bytes >= rbvae_decoder_workspace(r, N), "rbvae decoder: workspace too small");
  const int ns = dec_slice(r, N);
  Arena ar(ws, ws_bytes);
  float* dbuf = (float*)ar.take((size_t)N * r->L * 4);
  float* bufA = (float*)ar.take(dec_bytes_a(r, ns));
  float* bufB = (float*)ar.take(dec_bytes_b(r, ns));
  float* d = d_seq ? d_seq : dbuf;
  // decoder_rnn: the stacked LSTM over z_seq (no fc partials, no noise, no code)
  SFV_TRY(launch_lstm_code(z_seq, 0, nullptr, B, T, r->L, r->layers, r->w_ih, r->w_hh, r->lstm_b, nullptr, 0.f, 1.f, 0, d, nullptr,
                           nullptr, s));
  const int C = r->channels, Co = r->out_channels;
  for (int n0 = 0; n0 < N; n0 += ns) {
    const int nn = (N - n0) < ns ? (N - n0) : ns;
    int h = r->fh, w = r->fw;
    {  // fc -> [nn][fh][fw][C]
      const long long tot = (long long)nn * C * h * w;
      fc_decode_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(d + (size_t)n0 * r->L, r->fc_w, r->fc_b, bufA, nn, r->L, C, h * w);
      SFV_LAUNCH_OK();
    }
    float* cur = bufA;
    float* nxt = bufB;
    for (int i = 0; i < 3; ++i) {
      // cur [nn][h][w][C] -> nxt [nn][2h][2w][Cout]: four sub-pixel phases, (+ReLU after the first two layers)
      for (int ph = 0; ph < 4; ++ph) {
        IgemmArgs g;
        memset(&g, 0, sizeof(g));
        g.x = cur; g.src_kind = SRC_NHWC_F32; g.w = r->ph_w[i][ph]; g.w_sk = r->dc_cout[i]; g.w_sn = 1;
        g.bias = r->dc_bias[i]; g.y = nxt;
        g.N = nn; g.H = h; g.W = w; g.Cin = C; g.Ho = h; g.Wo = w; g.Cout = r->dc_cout[i];
        g.ksize = 1 + (ph >> 1); g.ksize_x = 1 + (ph & 1); g.stride = 1; g.pad = 0;
        g.relu = i < 2 ? 1 : 0; g.alpha = 1.f; g.in_scale = 1.f; g.ldy = r->dc_cout[i];
        g.osy = 2; g.osx = 2; g.ooy = ph >> 1; g.oox = ph & 1; g.oHf = 2 * h; g.oWf = 2 * w;
        SFV_TRY(launch_igemm_f32(g, s));
      }
      h *= 2; w *= 2;
      float* t = cur; cur = nxt; nxt = t;
    }
    // cur [nn][H][W][Co] -> sigmoid -> x_recon [nn][Co][H][W]
    const long long HW = (long long)h * w;
    const long long tot = (long long)nn * Co * HW;
    sigmoid_nchw_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(cur, x_recon + (size_t)n0 * Co * HW, Co, HW, tot);
    SFV_LAUNCH_OK();
  }
  return 0;
}

// ---- losses --------------------------------------------------------------------------------------------------------
int launch_mse(const float* a, const float* b, long long n, float* out, cudaStream_t s) {
  SFV_CHECK(a && b && out && n >= 1, "mse: bad arguments");
  mse_kernel<<<kLossCtas, 1024, 0, s>>>(a, b, n, out);
  SFV_LAUNCH_OK();
  return 0;
}
int launch_l1(const float* q, long long n, float lamb, float* out, cudaStream_t s) {
  SFV_CHECK(q && out && n >= 1, "l1: bad arguments");
  l1_kernel<<<kLossCtas, 1024, 0, s>>>(q, n, lamb, out);
  SFV_LAUNCH_OK();
  return 0;
}
int launch_kl_binary_concrete(const float* q_logits, long long rows, int L, float p, float eps, float* out, cudaStream_t s) {
  SFV_CHECK(q_logits && out && rows >= 1 && L >= 1 && p > 0.f && p < 1.f, "kl_binary_concrete: bad arguments");
  // the reference takes np.log of the python float p (float64) and lets torch round it to the tensor's float32
  kl_kernel<<<kLossCtas, 1024, 0, s>>>(q_logits, rows, L, (float)log((double)p), (float)log(1.0 - (double)p), eps, out);
  SFV_LAUNCH_OK();
  return 0;
}
int launch_contrast(const float* x1, const float* x2, const float* label, int rows, int D, float margin, int cosine, float* out,
                    cudaStream_t s) {
  SFV_CHECK(x1 && x2 && label && out && rows >= 1 && D >= 1, "contrast_loss: bad arguments");
  contrast_kernel<<<1, 256, 0, s>>>(x1, x2, label, rows, D, margin, cosine, out);
  SFV_LAUNCH_OK();
  return 0;
}
int launch_triplet(const float* a, const float* p, const float* n, int rows, int D, float margin, float eps, int swap, float* out,
                   cudaStream_t s) {
  SFV_CHECK(a && p && n && out && rows >= 1 && D >= 1, "triplet_loss: bad arguments");
  triplet_kernel<<<1, 256, 0, s>>>(a, p, n, rows, D, margin, eps, swap, out);
  SFV_LAUNCH_OK();
  return 0;
}

}  // namespace sfv
