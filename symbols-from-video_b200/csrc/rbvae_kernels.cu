// rbvae_kernels.cu -- tail of the RBVAE encoder:
// fc -> stacked LSTM -> binary-concrete threshold -> bit-packed code.
// Reference: models/percep_RBVAE/percep_RBVAE_model.py:17-44 (binary_concrete_logits),
// :61-67 (fc), :94-107 (EncoderRNN), :172-191 (encode).
#include "common.cuh"

namespace sfv {
namespace {

// y[n][l] = bias[l] + sum_k x[n][k] * w[l][k].  grid (L, N), block 256.
__global__ void __launch_bounds__(256) fc_kernel(const float* x, const float* w, const float* bias,
                                                 float* y, long long K, int L) {
  __shared__ float red[8];
  const int l = blockIdx.x, n = blockIdx.y;
  const float* xr = x + (long long)n * K;
  const float* wr = w + (long long)l * K;
  float acc = 0.f;
  const long long K4 = K & ~3ll;
  for (long long k = (long long)threadIdx.x * 4; k < K4; k += 1024) {
    const float4 a = *reinterpret_cast<const float4*>(xr + k);
    const float4 b = *reinterpret_cast<const float4*>(wr + k);
    acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc);
    acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
  }
  for (long long k = K4 + threadIdx.x; k < K; k += 256) acc = fmaf(xr[k], wr[k], acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i];
    y[(long long)n * L + l] = s + bias[l];
  }
}

__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + expf(-v)); }

// One block per sequence b; blockDim = 4*Lpad (Lpad = L rounded up to 32).
// Stacked LSTM from zero state, PyTorch gate order (i, f, g, o); layer-major so
// that layer l consumes the whole output sequence of layer l-1 (in place in hbuf).
// Then binary_concrete_logits on the top layer's hidden state.
__global__ void lstm_code_kernel(const float* logits, int T, int L, int layers,
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
    for (int i = j; i < L; i += blockDim.x) seq[t * L + i] = logits[((long long)b * T + t) * L + i];
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

int launch_fc(const float* x, const float* w, const float* bias, float* y, int N, long long K, int L,
              float*, int, cudaStream_t s) {
  SFV_CHECK(N <= 65535, "fc: N too large");
  fc_kernel<<<dim3(L, N), 256, 0, s>>>(x, w, bias, y, K, L);
  SFV_LAUNCH_OK();
  return 0;
}

int launch_lstm_code(const float* logits, int B, int T, int L, int layers, const float* w_ih,
                     const float* w_hh, const float* bias, const float* u, float noise_ratio,
                     float temperature, int hard, float* h_out, float* z_out, uint32_t* codes,
                     cudaStream_t s) {
  SFV_CHECK(L >= 1 && L <= 256, "lstm: latent_dim %d out of range [1,256]", L);
  SFV_CHECK(h_out != nullptr, "lstm: h buffer required");
  const int Lpad = (L + 31) / 32 * 32;
  const int threads = 4 * Lpad;
  lstm_code_kernel<<<B, threads, 7 * L * sizeof(float), s>>>(logits, T, L, layers, w_ih, w_hh, bias, u,
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
