// eval_kernels.cu -- the evaluation helpers either side of the hot path:
//   * state consistency on packed codes
//       scripts/evaluation/state_consistency_eval/embedding_matching.py:275-297
//       (np.unique(axis=0) + "matches the most common vector" per state)
//   * robustness perturbations on uint8 frames
//       embedding_matching.py:141-193 (add_gaussian_noise / add_occlusion, followed by
//       T.ToPILImage() at :243 -- i.e. ToTensor's x/255, the perturbation, mul(255).byte()).
// Integer / byte work: results are bit-exact with the reference's arithmetic.
#include "common.cuh"

namespace sfv {
namespace {

constexpr int kMaxWords = 8;       // latent_dim <= 256 (same limit as the LSTM kernel)
constexpr int kTile = 256;

// One thread per frame i: how many frames of the same state carry the same code.  The state's
// most common code is the one with the largest such count; ties give the same percentage.
__global__ void __launch_bounds__(kTile) code_multiplicity_kernel(const uint32_t* __restrict__ codes,
                                                                  const int* __restrict__ labels, long long n,
                                                                  int words, int n_states, int* best, int* count) {
  __shared__ uint32_t s_code[kTile * kMaxWords];
  __shared__ int s_lab[kTile];
  const long long i = (long long)blockIdx.x * kTile + threadIdx.x;
  uint32_t mine[kMaxWords];
  int lab = -1;
  if (i < n) {
    lab = labels[i];
    for (int w = 0; w < words; ++w) mine[w] = codes[i * words + w];
  }
  int same = 0;
  for (long long j0 = 0; j0 < n; j0 += kTile) {
    const long long j = j0 + threadIdx.x;
    __syncthreads();
    s_lab[threadIdx.x] = j < n ? labels[j] : -2;
    for (int w = 0; w < words; ++w) s_code[w * kTile + threadIdx.x] = j < n ? codes[j * words + w] : 0u;
    __syncthreads();
    for (int t = 0; t < kTile; ++t) {
      bool eq = s_lab[t] == lab;
      for (int w = 0; w < words; ++w) eq = eq && (s_code[w * kTile + t] == mine[w]);
      same += eq ? 1 : 0;
    }
  }
  if (i < n && lab >= 0 && lab < n_states) {
    atomicMax(&best[lab], same);
    atomicAdd(&count[lab], 1);
  }
}

// HWC uint8 frame, CHW fp32 noise (the layout the reference's randn_like(ToTensor(img)) draws in).
__global__ void perturb_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int B, int H, int W,
                               const float* __restrict__ noise, float mean, float stdv,
                               const int* __restrict__ occ_xy, int osz) {
  const long long total = (long long)B * H * W * 3;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % 3);
    const long long p = e / 3;
    const int x = (int)(p % W);
    const int y = (int)((p / W) % H);
    const long long b = p / ((long long)W * H);
    uint8_t v = in[e];
    if (noise) {
      // ToTensor: float(u8)/255 ; + (randn*std + mean) ; clamp[0,1] ; ToPILImage: mul(255).byte()
      const float t = __fdiv_rn((float)v, 255.f);
      const float nz = __fadd_rn(__fmul_rn(noise[((b * 3 + c) * H + y) * (long long)W + x], stdv), mean);
      float s = __fadd_rn(t, nz);
      s = fminf(fmaxf(s, 0.f), 1.f);
      v = (uint8_t)(int)__fmul_rn(s, 255.f);
    }
    if (occ_xy) {
      const int ox = occ_xy[2 * b], oy = occ_xy[2 * b + 1];
      if (x >= ox && x < ox + osz && y >= oy && y < oy + osz) v = 127;   // 0.5 * 255 -> byte() truncates
    }
    out[e] = v;
  }
}

}  // namespace

int launch_state_consistency(const uint32_t* codes, const int* labels, long long n, int words, int n_states,
                             int* best, int* count, cudaStream_t s) {
  SFV_CHECK(words >= 1 && words <= kMaxWords, "state_consistency: words %d out of range [1,%d]", words, kMaxWords);
  SFV_CHECK(n >= 0 && n_states >= 1, "state_consistency: bad sizes");
  SFV_CUDA(cudaMemsetAsync(best, 0, sizeof(int) * n_states, s));
  SFV_CUDA(cudaMemsetAsync(count, 0, sizeof(int) * n_states, s));
  if (n == 0) return 0;
  code_multiplicity_kernel<<<(unsigned)ceil_div(n, (long long)kTile), kTile, 0, s>>>(codes, labels, n, words, n_states,
                                                                                   best, count);
  SFV_LAUNCH_OK();
  return 0;
}

int launch_perturb(const uint8_t* in, uint8_t* out, int B, int H, int W, const float* noise, float mean, float stdv,
                   const int* occ_xy, int osz, cudaStream_t s) {
  SFV_CHECK(B >= 0 && H >= 1 && W >= 1, "perturb: bad sizes");
  SFV_CHECK(osz >= 0 && osz <= H && osz <= W, "perturb: occlusion square %d larger than the %dx%d frame", osz, W, H);
  if (osz == 0) occ_xy = nullptr;
  if (B == 0) return 0;
  const long long total = (long long)B * H * W * 3;
  const int blocks = (int)(ceil_div(total, 256LL) < 148 * 16 ? ceil_div(total, 256LL) : 148 * 16);
  perturb_kernel<<<blocks, 256, 0, s>>>(in, out, B, H, W, noise, mean, stdv, occ_xy, osz);
  SFV_LAUNCH_OK();
  return 0;
}

}  // namespace sfv
