// resize.cu -- uint8 frame resize + normalise.
//
// Replaces load_img (src/stable-diffusion/get_percep_embeddings.py:48-71): PIL
// Image.resize(..., LANCZOS) followed by /255, HWC->CHW, 2x-1.  The arithmetic
// lives in Pillow (pillow==10.2.0 pinned at reference requirements.txt:113;
// libImaging/Resample.c): a separable two-pass convolution with a Lanczos a=3
// kernel stretched by the down-scale factor, coefficients normalised and
// quantised to 22-bit fixed point, uint8 rounding/clipping after EACH pass
// (horizontal first), passes skipped when the size along that axis is unchanged.
// The coefficient tables are computed on the host in double precision exactly
// as precompute_coeffs()/normalize_coeffs_8bpc() do; the two integer passes run
// on the GPU, so the uint8 result is bit-identical to Image.resize.
#include "common.cuh"
#include <math.h>
#include <mutex>

namespace sfv {
namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;

struct CoeffTable { int in_size, out_size, ksize; int* d_bounds; int* d_kk; };
std::mutex g_mu;
std::vector<CoeffTable> g_tables;

double sinc_filter(double x) {
  if (x == 0.0) return 1.0;
  x = x * M_PI;
  return sin(x) / x;
}
double lanczos_filter(double x) {
  if (-3.0 <= x && x < 3.0) return sinc_filter(x) * sinc_filter(x / 3);
  return 0.0;
}

int get_table(int in_size, int out_size, CoeffTable* out) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (const CoeffTable& t : g_tables)
    if (t.in_size == in_size && t.out_size == out_size) { *out = t; return 0; }
  const double support0 = 3.0;
  double scale = (double)in_size / out_size, filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = support0 * filterscale;
  const int ksize = (int)ceil(support) * 2 + 1;
  std::vector<int> bounds((size_t)out_size * 2);
  std::vector<int> kk((size_t)out_size * ksize, 0);
  std::vector<double> k(ksize);
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    double ww = 0.0;
    const double ss = 1.0 / filterscale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    for (int x = 0; x < xmax; ++x) {
      const double w = lanczos_filter((x + xmin - center + 0.5) * ss);
      k[x] = w;
      ww += w;
    }
    for (int x = 0; x < xmax; ++x) {
      double v = k[x];
      if (ww != 0.0) v /= ww;
      kk[(size_t)xx * ksize + x] = v < 0 ? (int)(-0.5 + v * (1 << kPrecisionBits)) : (int)(0.5 + v * (1 << kPrecisionBits));
    }
    bounds[xx * 2] = xmin;
    bounds[xx * 2 + 1] = xmax;
  }
  CoeffTable t;
  t.in_size = in_size; t.out_size = out_size; t.ksize = ksize;
  SFV_CUDA(cudaMalloc(&t.d_bounds, bounds.size() * 4));
  SFV_CUDA(cudaMalloc(&t.d_kk, kk.size() * 4));
  SFV_CUDA(cudaMemcpy(t.d_bounds, bounds.data(), bounds.size() * 4, cudaMemcpyHostToDevice));
  SFV_CUDA(cudaMemcpy(t.d_kk, kk.data(), kk.size() * 4, cudaMemcpyHostToDevice));
  g_tables.push_back(t);
  *out = t;
  return 0;
}

__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= kPrecisionBits;
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// in [B][H][Win][3] -> out [B][H][Wout][3]
__global__ void resize_h_kernel(const uint8_t* in, uint8_t* out, const int* bounds, const int* kk, int ksize,
                                int H, int Win, int Wout, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // (b*H + y)*Wout + xo
  if (i >= total) return;
  const int xo = (int)(i % Wout);
  const long long row = i / Wout;
  const int xmin = bounds[xo * 2], xmax = bounds[xo * 2 + 1];
  const int* k = kk + (long long)xo * ksize;
  const uint8_t* p = in + (row * Win + xmin) * 3;
  int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
  for (int x = 0; x < xmax; ++x) {
    const int c = k[x];
    s0 += p[x * 3] * c; s1 += p[x * 3 + 1] * c; s2 += p[x * 3 + 2] * c;
  }
  uint8_t* o = out + i * 3;
  o[0] = clip8(s0); o[1] = clip8(s1); o[2] = clip8(s2);
}

// in [B][Hin][W][3] -> out [B][Hout][W][3]
__global__ void resize_v_kernel(const uint8_t* in, uint8_t* out, const int* bounds, const int* kk, int ksize,
                                int Hin, int Hout, int W, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // ((b*Hout + yo)*W + x)*3 + c
  if (i >= total) return;
  const int rowlen = W * 3;
  const int xc = (int)(i % rowlen);
  const long long t = i / rowlen;
  const int yo = (int)(t % Hout);
  const long long b = t / Hout;
  const int ymin = bounds[yo * 2], ymax = bounds[yo * 2 + 1];
  const int* k = kk + (long long)yo * ksize;
  const uint8_t* p = in + (b * Hin + ymin) * rowlen + xc;
  int s = 1 << (kPrecisionBits - 1);
  for (int y = 0; y < ymax; ++y) s += p[(long long)y * rowlen] * k[y];
  out[i] = clip8(s);
}

// u8 HWC [B][H][W][3] -> fp32 NCHW, 2*(v/255)-1
__global__ void normalise_kernel(const uint8_t* in, float* out, int HW, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // (b*HW + p)
  if (i >= total) return;
  const long long b = i / HW; const int p = (int)(i - b * HW);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float f = (float)in[i * 3 + c] / 255.0f;
    out[(b * 3 + c) * HW + p] = 2.f * f - 1.f;
  }
}

}  // namespace

int resize_workspace(int B, int Hs, int Ws, int H, int W, size_t* bytes) {
  if (!bytes || B < 1 || Hs < 1 || Ws < 1 || H < 1 || W < 1) return fail(SFV_ERR_INVALID, "resize: bad shape");
  *bytes = align_up((size_t)B * Hs * W * 3, 1024) + align_up((size_t)B * H * W * 3, 1024) + 1024;
  return 0;
}

int resize_normalise(const uint8_t* frames, int B, int Hs, int Ws, int H, int W, float* out_nchw, uint8_t* out_u8,
                     void* ws, size_t ws_bytes, cudaStream_t s) {
  size_t need = 0;
  SFV_TRY(resize_workspace(B, Hs, Ws, H, W, &need));
  SFV_CHECK(frames != nullptr, "resize: null frames");
  SFV_CHECK(ws != nullptr && ws_bytes >= need, "resize: workspace too small");
  Arena ar(ws, ws_bytes);
  uint8_t* mid = (uint8_t*)ar.take((size_t)B * Hs * W * 3);
  uint8_t* fin = (uint8_t*)ar.take((size_t)B * H * W * 3);
  const uint8_t* cur = frames;
  if (Ws != W) {   // horizontal pass (skipped by Pillow when the width is unchanged)
    CoeffTable t;
    SFV_TRY(get_table(Ws, W, &t));
    const long long total = (long long)B * Hs * W;
    resize_h_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(cur, mid, t.d_bounds, t.d_kk, t.ksize, Hs, Ws, W, total);
    SFV_LAUNCH_OK();
    cur = mid;
  }
  if (Hs != H) {
    CoeffTable t;
    SFV_TRY(get_table(Hs, H, &t));
    uint8_t* dst = out_u8 ? out_u8 : fin;
    const long long total = (long long)B * H * W * 3;
    resize_v_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(cur, dst, t.d_bounds, t.d_kk, t.ksize, Hs, H, W, total);
    SFV_LAUNCH_OK();
    cur = dst;
  }
  if (out_u8 && cur != out_u8) {
    SFV_CUDA(cudaMemcpyAsync(out_u8, cur, (size_t)B * H * W * 3, cudaMemcpyDeviceToDevice, s));
    cur = out_u8;
  }
  if (out_nchw) {
    const long long total = (long long)B * H * W;
    normalise_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(cur, out_nchw, H * W, total);
    SFV_LAUNCH_OK();
  }
  return 0;
}

}  // namespace sfv
