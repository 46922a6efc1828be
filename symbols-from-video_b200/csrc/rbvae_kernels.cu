// rbvae_kernels.cu -- tail of the RBVAE encoder:
// fc -> stacked LSTM -> binary-concrete threshold -> bit-packed code.
// Reference: models/percep_RBVAE/percep_RBVAE_model.py:17-44 (binary_concrete_logits),
// :61-67 (fc), :94-107 (EncoderRNN), :172-191 (encode).
#include "common.cuh"

namespace sfv {
namespace {

// ---- fc (percep_RBVAE_model.py:61,67): logits[n][l] = bias[l] + sum_k x[n][k] w[l][k] ------------------------------
// HBM-bound skinny GEMM (L <= 256 outputs, K up to 262144 inputs per frame): every byte of x and of w is read from
// HBM exactly once.  Split-K: block s owns the K-slice [s*KS, (s+1)*KS), keeps its slice of the weight in shared
// memory (transposed to [k][l], so lane = l reads conflict-free and x is a broadcast) and streams the same slice of
// EVERY frame past it, eight frames at a time (register blocking: two 16-byte broadcast loads of x feed eight FMAs).
// It writes partial[s][n][l]; the sum over s (fixed order: deterministic, no atomics) and the bias are folded into
// the LSTM kernel's load, so the logits never exist as a tensor of their own (fc -> LSTM -> threshold -> pack).
constexpr int kFcThreads = 256;
constexpr int kFcFrames = 8;
constexpr int kFcWFloats = 8192;           // weight slice per block: 32 KB (four blocks per SM: the k loop is latency bound)
constexpr int kFcFramesPerBlock = 32;      // frames one block streams past its weight slice (grid.y covers the rest)

__global__ void __launch_bounds__(kFcThreads) fc_splitk_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                               float* __restrict__ partial, int N, long long n_stride,
                                                               long long K, int L, int Lp, int KS) {
  extern __shared__ float fsm[];
  const int Lw = Lp + 1;                             // padded row: the transposing stores below hit 32 different banks
  float* wt = fsm;                                   // [KS][Lp + 1]
  float* xs = fsm + (size_t)((KS * Lw + 3) & ~3);    // [KS][kFcFrames], 16-byte aligned
  float* red = xs + (size_t)KS * kFcFrames;          // [8 warps][Lp][kFcFrames]
  const int s = blockIdx.x;
  const long long k0 = (long long)s * KS;
  const int kn = (int)min((long long)KS, K - k0);    // valid k in this slice
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // stage the weight slice, transposed: coalesced reads along k, writes [k][l]
  for (int i = threadIdx.x; i < Lp * KS; i += kFcThreads) {
    const int l = i / KS, k = i - l * KS;
    wt[k * Lw + l] = (l < L && k < kn) ? w[(long long)l * K + k0 + k] : 0.f;
  }
  const int lchunks = Lp >> 5;
  const int kw0 = warp * (KS >> 3), kw1 = kw0 + (KS >> 3);     // this warp's k sub-slice
  // the slice of the NEXT frame group is fetched into registers while the current one is being reduced, so the DRAM
  // latency of the streaming operand is not paid once per group
  const int per_thread = (kFcFrames * KS + kFcThreads - 1) / kFcThreads;      // <= 16 for KS <= 512
  // blockIdx.y: this block's share of the frames (kFcFramesPerBlock), so the per-block chain of dependent groups stays short
  const int n_begin = blockIdx.y * kFcFramesPerBlock;
  const int n_end = min(N, n_begin + kFcFramesPerBlock);
  float nxt[16];
  auto fetch = [&](int n0) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int i = threadIdx.x + j * kFcThreads;
      float v = 0.f;
      if (j < per_thread && i < kFcFrames * KS) {
        const int f = i / KS, k = i - f * KS;
        const int n = n0 + f;
        if (n < n_end && k < kn) v = __ldg(x + (long long)n * K + k0 + k);
      }
      nxt[j] = v;
    }
  };
  fetch(n_begin);
  for (int n0 = n_begin; n0 < n_end; n0 += kFcFrames) {
    __syncthreads();                                 // previous group's xs / red are consumed (also covers the wt stage)
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int i = threadIdx.x + j * kFcThreads;
      if (j < per_thread && i < kFcFrames * KS) {
        const int f = i / KS, k = i - f * KS;
        xs[k * kFcFrames + f] = nxt[j];
      }
    }
    if (n0 + kFcFrames < n_end) fetch(n0 + kFcFrames);
    __syncthreads();
    for (int lc = 0; lc < lchunks; ++lc) {
      const int l = lc * 32 + lane;
      float a[kFcFrames];
#pragma unroll
      for (int f = 0; f < kFcFrames; ++f) a[f] = 0.f;
#pragma unroll 4
      for (int k = kw0; k < kw1; ++k) {
        const float4 x0 = *reinterpret_cast<const float4*>(xs + k * kFcFrames);
        const float4 x1 = *reinterpret_cast<const float4*>(xs + k * kFcFrames + 4);
        const float wv = wt[k * Lw + l];
        a[0] = fmaf(x0.x, wv, a[0]); a[1] = fmaf(x0.y, wv, a[1]); a[2] = fmaf(x0.z, wv, a[2]); a[3] = fmaf(x0.w, wv, a[3]);
        a[4] = fmaf(x1.x, wv, a[4]); a[5] = fmaf(x1.y, wv, a[5]); a[6] = fmaf(x1.z, wv, a[6]); a[7] = fmaf(x1.w, wv, a[7]);
      }
      float* rd = red + ((size_t)warp * Lp + l) * kFcFrames;
      *reinterpret_cast<float4*>(rd) = make_float4(a[0], a[1], a[2], a[3]);
      *reinterpret_cast<float4*>(rd + 4) = make_float4(a[4], a[5], a[6], a[7]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < Lp * kFcFrames; i += kFcThreads) {
      const int l = i / kFcFrames, f = i - l * kFcFrames;
      const int n = n0 + f;
      if (l < L && n < n_end) {
        float t = 0.f;
#pragma unroll
        for (int wv = 0; wv < 8; ++wv) t += red[((size_t)wv * Lp + l) * kFcFrames + f];
        partial[((long long)s * n_stride + n) * L + l] = t;
      }
    }
  }
}

__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + expf(-v)); }

// One block per sequence b; blockDim = 4*Lpad (Lpad = L rounded up to 32).
// Stacked LSTM from zero state, PyTorch gate order (i, f, g, o); layer-major so
// that layer l consumes the whole output sequence of layer l-1 (in place in hbuf).
// Then binary_concrete_logits on the top layer's hidden state.
// `logits`: [n][L] ready values (splits == 0), or the fc layer's split-K partial sums [splits][n_total][L] with `fc_bias`.
__global__ void lstm_code_kernel(const float* logits, int splits, long long n_total, const float* fc_bias, int T, int L, int layers,
                                 const float* w_ih, const float* w_hh, const float* bias,
                                 const float* u, float noise_ratio, float temperature, int hard,
                                 float* hbuf, float* z_out, uint32_t* codes) {
  extern __shared__ float sh[];
  float* h_prev = sh;            // [L]
  float* c_st = sh + L;          // [L]
  float* xin = sh + 2 * L;       // [L]
  float* gates = sh + 3 * L;     // [4L]
  const int b = blockIdx.x;
  const int j = threadIdx.x;
  float* seq = hbuf + (long long)b * T * L;
  for (int t = 0; t < T; ++t)
    for (int i = j; i < L; i += blockDim.x) {
      const long long bt = (long long)b * T + t;
      float v;
      if (splits == 0) v = logits[bt * L + i];
      else {
        v = fc_bias[i];
        for (int sp = 0; sp < splits; ++sp) v += logits[((long long)sp * n_total + bt) * L + i];   // fixed order
      }
      seq[t * L + i] = v;
    }
  __syncthreads();
  for (int l = 0; l < layers; ++l) {
    const float* Wi = w_ih + (long long)l * 4 * L * L;
    const float* Wh = w_hh + (long long)l * 4 * L * L;
    const float* bs = bias + (long long)l * 4 * L;     // b_ih + b_hh (pre-summed on the host)
    if (j < L) { h_prev[j] = 0.f; c_st[j] = 0.f; }
    __syncthreads();
    for (int t = 0; t < T; ++t) {
      if (j < L) xin[j] = seq[t * L + j];
      __syncthreads();
      if (j < 4 * L) {
        float a = 0.f, r = 0.f;
        const float* wi = Wi + (long long)j * L;
        const float* wh = Wh + (long long)j * L;
        for (int i = 0; i < L; ++i) { a = fmaf(wi[i], xin[i], a); r = fmaf(wh[i], h_prev[i], r); }
        gates[j] = (a + r) + bs[j];
      }
      __syncthreads();
      if (j < L) {
        const float ig = sigmoidf_(gates[j]), fg = sigmoidf_(gates[L + j]);
        const float gg = tanhf(gates[2 * L + j]), og = sigmoidf_(gates[3 * L + j]);
        const float c = fg * c_st[j] + ig * gg;
        const float h = og * tanhf(c);
        c_st[j] = c; h_prev[j] = h;
        seq[t * L + j] = h;
      }
      __syncthreads();
    }
  }
  // binary concrete + bit pack
  const int words = (L + 31) / 32;
  for (int t = 0; t < T; ++t) {
    const long long bt = (long long)b * T + t;
    bool bit = false;
    if (j < L) {
      const float h = seq[t * L + j];
      float noise = 0.f;
      if (u != nullptr && noise_ratio != 0.f) {
        const float uu = u[bt * L + j];
        noise = noise_ratio * (logf(uu + 1e-8f) - logf(1.0f - uu + 1e-8f));
      }
      float y = sigmoidf_((h + noise) / temperature);
      bit = y > 0.5f;
      if (hard) { const float yh = bit ? 1.f : 0.f; y = (yh - y) + y; }
      if (z_out) z_out[bt * L + j] = y;
    }
    const unsigned m = __ballot_sync(0xffffffffu, bit);
    if (codes && (j & 31) == 0 && (j >> 5) < words) codes[bt * words + (j >> 5)] = m;
  }
}

__global__ void hamming_kernel(const uint32_t* a, int Na, const uint32_t* b, int Nb, int words, int* out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (j >= Nb) return;
  int d = 0;
  for (int w = 0; w < words; ++w) d += __popc(a[(long long)i * words + w] ^ b[(long long)j * words + w]);
  out[(long long)i * Nb + j] = d;
}

}  // namespace

// split-K geometry of the fc layer for (K, L): slice length KS and number of slices
void fc_plan(long long K, int L, int* KS, int* splits) {
  const int Lp = (L + 31) / 32 * 32;
  int ks = kFcWFloats / Lp;
  if (ks > 512) ks = 512;                       // 8 frames x 512 k = 16 prefetch registers per thread
  static const int ks_env = []() { const char* e = getenv("SFV_FC_KS"); return e ? atoi(e) : 0; }();   // experiments
  if (ks_env > 0 && ks_env < ks) ks = ks_env;
  ks = ks / 8 * 8;                              // eight warps split the slice evenly
  if (ks < 8) ks = 8;
  if ((long long)ks > K) ks = (int)((K + 7) / 8 * 8);
  *KS = ks;
  *splits = (int)((K + ks - 1) / ks);
}

// partial: [splits][n_stride][L] (fc_plan), rows 0..N-1 of every slice written; the consumer (launch_lstm_code) adds the
// bias and the slices
int launch_fc(const float* x, const float* w, float* partial, int N, long long n_stride, long long K, int L, cudaStream_t s) {
  SFV_CHECK(L >= 1 && L <= 256, "fc: latent_dim %d out of range [1,256]", L);
  int KS, splits;
  fc_plan(K, L, &KS, &splits);
  const int Lp = (L + 31) / 32 * 32;
  const size_t smem = ((size_t)((KS * (Lp + 1) + 3) & ~3) + (size_t)KS * kFcFrames + (size_t)8 * Lp * kFcFrames) * sizeof(float);
  static unsigned long long attr_devs = 0;
  int dev = 0;
  SFV_CUDA(cudaGetDevice(&dev));
  if (first_use_on_device(attr_devs, dev))
    SFV_CUDA(cudaFuncSetAttribute(fc_splitk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  SFV_CHECK(smem <= 160 * 1024, "fc: shared-memory plan too large (%zu)", smem);
  ProfScope prof(PROF_OTHER, ((double)N * K + (double)L * K) * 4.0, s);
  fc_splitk_kernel<<<dim3(splits, (N + kFcFramesPerBlock - 1) / kFcFramesPerBlock), kFcThreads, smem, s>>>(x, w, partial, N, n_stride, K,
                                                                                                        L, Lp, KS);
  SFV_LAUNCH_OK();
  return 0;
}

int launch_lstm_code(const float* logits, int splits, const float* fc_bias, int B, int T, int L, int layers, const float* w_ih,
                     const float* w_hh, const float* bias, const float* u, float noise_ratio,
                     float temperature, int hard, float* h_out, float* z_out, uint32_t* codes,
                     cudaStream_t s) {
  SFV_CHECK(splits == 0 || fc_bias != nullptr, "lstm: split-K partials need the fc bias");
  SFV_CHECK(L >= 1 && L <= 256, "lstm: latent_dim %d out of range [1,256]", L);
  SFV_CHECK(h_out != nullptr, "lstm: h buffer required");
  const int Lpad = (L + 31) / 32 * 32;
  const int threads = 4 * Lpad;
  lstm_code_kernel<<<B, threads, 7 * L * sizeof(float), s>>>(logits, splits, (long long)B * T, fc_bias, T, L, layers, w_ih, w_hh, bias, u,
                                                             noise_ratio, temperature, hard, h_out, z_out,
                                                             codes);
  SFV_LAUNCH_OK();
  return 0;
}

int launch_hamming(const uint32_t* a, int Na, const uint32_t* b, int Nb, int words, int* out,
                   cudaStream_t s) {
  SFV_CHECK(Na <= 65535, "hamming: Na too large");
  hamming_kernel<<<dim3(ceil_div(Nb, 128), Na), 128, 0, s>>>(a, Na, b, Nb, words, out);
  SFV_LAUNCH_OK();
  return 0;
}

}  // namespace sfv
