// conv_in.cu -- Encoder.conv_in (3x3, 3 -> 128, pad 1; model.py:383-387,439) fed
// straight from the boundary layout: uint8 HWC frames (load_img's /255 and 2x-1
// fused, get_percep_embeddings.py:67-71) or the reference's fp32 NCHW tensor.
//
// K = 27 does not fit UMMA and the layer is write-bound (512 B of fp32 NHWC per
// pixel against 3456 FMAs), so it runs on CUDA cores: a warp owns one pixel row
// segment, lane l owns output channels 4l..4l+3 (= GroupNorm group l of
// down.0.block.0.norm1), its 27x4 weights live in registers, the input patch is
// staged in shared memory and read as warp broadcasts, and each pixel is written
// as one coalesced 512-byte row.  The GroupNorm statistics of the output are
// accumulated on the fly (fp32 per lane, fp64 atomics per block).
#include "common.cuh"

namespace sfv {
namespace {

constexpr int TH = 8, TW = 32;

template <int SRC>
__global__ void __launch_bounds__(256, 1)
conv_in_kernel(const void* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
               float* __restrict__ y, double* __restrict__ stats, int H, int W, int tiles_x, int tiles,
               uint16_t* __restrict__ y16, int fmt16, float y16_scale) {
  __shared__ float patch[TH + 2][TW + 2][3];
  __shared__ float red[8][32][2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = blockIdx.y;
  float wr[27][4];
#pragma unroll
  for (int k = 0; k < 27; ++k) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(w + k * 128 + lane * 4));
    wr[k][0] = t.x; wr[k][1] = t.y; wr[k][2] = t.z; wr[k][3] = t.w;
  }
  const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + lane * 4));
  float s_sum = 0.f, s_sq = 0.f;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    const int y0 = ty * TH, x0 = tx * TW;
    __syncthreads();
    for (int i = threadIdx.x; i < (TH + 2) * (TW + 2) * 3; i += 256) {
      const int c = i % 3;
      const int cc = (i / 3) % (TW + 2);
      const int r = i / (3 * (TW + 2));
      const int iy = y0 + r - 1, ix = x0 + cc - 1;
      float v = 0.f;
      if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
        if (SRC == SRC_NHWC_U8) {
          const uint8_t u = reinterpret_cast<const uint8_t*>(x)[(((long long)n * H + iy) * W + ix) * 3 + c];
          const float f = (float)u / 255.0f;
          v = 2.f * f - 1.f;
        } else {
          v = reinterpret_cast<const float*>(x)[(((long long)n * 3 + c) * H + iy) * W + ix];
        }
      }
      patch[r][cc][c] = v;
    }
    __syncthreads();
    const int oy = y0 + warp;
    if (oy < H) {
      const long long row0 = (((long long)n * H + oy) * W + x0) * 128 + lane * 4;
      float* yrow = y + row0;
#pragma unroll 2
      for (int px = 0; px < TW; ++px) {
        if (x0 + px >= W) break;
        float a0 = b4.x, a1 = b4.y, a2 = b4.z, a3 = b4.w;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int sx = 0; sx < 3; ++sx)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const float v = patch[warp + r][px + sx][c];
              const int k = (r * 3 + sx) * 3 + c;
              a0 = fmaf(v, wr[k][0], a0); a1 = fmaf(v, wr[k][1], a1);
              a2 = fmaf(v, wr[k][2], a2); a3 = fmaf(v, wr[k][3], a3);
            }
        if (y16) {        // 16-bit (scaled) residual stream: one coalesced 256-byte row per pixel
          uint2 u; u.x = pack2_16(a0 * y16_scale, a1 * y16_scale, fmt16); u.y = pack2_16(a2 * y16_scale, a3 * y16_scale, fmt16);
          *reinterpret_cast<uint2*>(y16 + row0 + (long long)px * 128) = u;
        } else {
          *reinterpret_cast<float4*>(yrow + (long long)px * 128) = make_float4(a0, a1, a2, a3);
        }
        s_sum += (a0 + a1) + (a2 + a3);
        s_sq += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
      }
    }
  }
  if (stats) {
    red[warp][lane][0] = s_sum; red[warp][lane][1] = s_sq;
    __syncthreads();
    if (threadIdx.x < 64) {
      const int g = threadIdx.x >> 1, m = threadIdx.x & 1;
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) t += red[k][g][m];
      atomicAdd(&stats[((long long)n * 32 + g) * 2 + m], (double)t);
    }
  }
}

// Contrastive RBVAE first layer (contrastive_RBVAE_model.py:50-52): conv3x3 3 -> 64, stride 2, pad 1, + ReLU, fed from
// the reference's fp32 NCHW frames in [0,1].  Same shape of problem as conv_in (K = 27, write-bound: 128 B of 16-bit
// NHWC per output pixel): a warp owns one output row segment, lane l owns channels 2l, 2l+1 with their 27x2 weights
// in registers, the (2*TH+1) x (2*TW+1) x 3 input patch is staged in shared memory with coalesced NCHW reads and read
// back as warp broadcasts, each pixel leaves as one coalesced 128-byte (16-bit) or 256-byte (fp32) row.
constexpr int RTH = 8, RTW = 32;
template <bool OUT16>
__global__ void __launch_bounds__(256)
rb_conv0_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                void* __restrict__ y, int fmt, int H, int W, int Ho, int Wo, float in_scale, int tiles_x, int tiles) {
  constexpr int PH = 2 * RTH + 1, PW = 2 * RTW + 1;
  __shared__ float patch[PH][PW][3];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = blockIdx.y;
  float wr[27][2];
#pragma unroll
  for (int k = 0; k < 27; ++k) {
    const float2 t = __ldg(reinterpret_cast<const float2*>(w + k * 64 + lane * 2));
    wr[k][0] = t.x; wr[k][1] = t.y;
  }
  const float2 b2 = __ldg(reinterpret_cast<const float2*>(bias + lane * 2));
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    const int y0 = ty * RTH, x0 = tx * RTW;
    __syncthreads();
    for (int i = threadIdx.x; i < PH * PW * 3; i += 256) {
      const int cc = i % PW;
      const int r = (i / PW) % PH;
      const int c = i / (PW * PH);
      const int iy = 2 * y0 - 1 + r, ix = 2 * x0 - 1 + cc;
      float v = 0.f;
      if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = x[(((long long)n * 3 + c) * H + iy) * W + ix] * in_scale;
      patch[r][cc][c] = v;
    }
    __syncthreads();
    const int oy = y0 + warp;
    if (oy < Ho) {
      const long long row = (((long long)n * Ho + oy) * Wo + x0) * 64 + lane * 2;
#pragma unroll 2
      for (int px = 0; px < RTW; ++px) {
        if (x0 + px >= Wo) break;
        float a0 = b2.x, a1 = b2.y;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int sx = 0; sx < 3; ++sx)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const float v = patch[2 * warp + r][2 * px + sx][c];
              const int k = (r * 3 + sx) * 3 + c;
              a0 = fmaf(v, wr[k][0], a0); a1 = fmaf(v, wr[k][1], a1);
            }
        a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f);
        if (OUT16) reinterpret_cast<uint32_t*>(y)[(row + (long long)px * 64) >> 1] = pack2_16(a0, a1, fmt);
        else *reinterpret_cast<float2*>(reinterpret_cast<float*>(y) + row + (long long)px * 64) = make_float2(a0, a1);
      }
    }
  }
}

// Contrastive conv.0 on the tensor pipe (contrastive_RBVAE_model.py:50-52), step 1 of 2: one thread per OUTPUT pixel gathers
// its 3x3x3 stride-2 patch from the fp32 NCHW frame and writes one 128-byte row of the GEMM's A operand:
// k = (r*3+s)*3 + c < 27 holds hi = to16(x), k + 27 holds lo = to16(x - hi), k >= 54 is zero.  The weights are [w | w | 0]
// (rbvae.cu), so A . B^T = sum_k (hi + lo) w: the input enters with ~fp32 precision, only the weights are rounded.
// Step 2 is the ordinary tcgen05 1x1 GEMM over these rows, IN PLACE: a tile's output rows (64 channels x 2 B) are its own
// input rows, so no second buffer exists and the rows are still in L2 when they are rewritten.
__global__ void __launch_bounds__(256) rb_im2col_kernel(const float* __restrict__ x, uint4* __restrict__ a16, int fmt,
                                                         int H, int W, int Ho, int Wo, float in_scale, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int ox = (int)(i % Wo);
  const int oy = (int)((i / Wo) % Ho);
  const long long n = i / ((long long)Wo * Ho);
  float v[27];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int iy = 2 * oy - 1 + r;
#pragma unroll
    for (int sx = 0; sx < 3; ++sx) {
      const int ix = 2 * ox - 1 + sx;
      const bool ok = iy >= 0 && iy < H && ix >= 0 && ix < W;
#pragma unroll
      for (int c = 0; c < 3; ++c)
        v[(r * 3 + sx) * 3 + c] = ok ? __ldg(x + ((n * 3 + c) * H + iy) * W + ix) * in_scale : 0.f;
    }
  }
  uint32_t wd[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) wd[k] = 0u;
  uint16_t hv[54];
#pragma unroll
  for (int k = 0; k < 27; ++k) {
    const uint16_t hi = f32_to_16(v[k], fmt);
    hv[k] = hi;
    hv[27 + k] = f32_to_16(v[k] - f16_to_32(hi, fmt), fmt);
  }
#pragma unroll
  for (int k = 0; k < 54; ++k) wd[k >> 1] |= (uint32_t)hv[k] << (16 * (k & 1));
  uint4* dst = a16 + i * 8;
#pragma unroll
  for (int q = 0; q < 8; ++q) dst[q] = make_uint4(wd[4 * q], wd[4 * q + 1], wd[4 * q + 2], wd[4 * q + 3]);
}

}  // namespace

int launch_rb_im2col(const float* x, void* a16, int fmt, int N, int H, int W, float in_scale, cudaStream_t s) {
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const long long total = (long long)N * Ho * Wo;
  ProfScope prof(PROF_OTHER, (double)N * (3.0 * H * W * 4 + (double)Ho * Wo * 128), s, "rb_im2col");
  rb_im2col_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(x, (uint4*)a16, fmt, H, W, Ho, Wo, in_scale, total);
  SFV_LAUNCH_OK();
  return 0;
}

// x fp32 NCHW [N,3,H,W]; w fp32 [27][64] (k = (r*3+s)*3 + c); y NHWC [N,Ho,Wo,64] fp32 (y16 == 0) or 16-bit.
int launch_rb_conv0(const float* x, const float* w, const float* bias, void* y, int y16, int fmt, int N, int H, int W,
                    float in_scale, cudaStream_t s) {
  SFV_CHECK(N <= 65535, "rb_conv0: batch too large");
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const int tiles_x = ceil_div(Wo, RTW), tiles = tiles_x * ceil_div(Ho, RTH);
  const int per = ceil_div(148 * 4, N);                  // about four blocks per SM over the whole batch
  const int gx = tiles < per ? tiles : (per < 1 ? 1 : per);
  ProfScope prof(PROF_IGEMM, 2.0 * N * (double)Ho * Wo * 64 * 27, s);
  if (y16) rb_conv0_kernel<true><<<dim3(gx, N), 256, 0, s>>>(x, w, bias, y, fmt, H, W, Ho, Wo, in_scale, tiles_x, tiles);
  else rb_conv0_kernel<false><<<dim3(gx, N), 256, 0, s>>>(x, w, bias, y, fmt, H, W, Ho, Wo, in_scale, tiles_x, tiles);
  SFV_LAUNCH_OK();
  return 0;
}

// x: uint8 HWC [N,H,W,3] or fp32 NCHW [N,3,H,W]; w: fp32 [27][128]; y: fp32 NHWC [N,H,W,128];
// stats_or_null: fp64 [N][32][2] (pre-zeroed) receives sum / sum of squares per GroupNorm group.
int launch_conv_in(const void* x, int src_kind, const float* w, const float* bias, float* y, double* stats,
                   int N, int H, int W, cudaStream_t s, void* y16, int fmt16, float y16_scale) {
  SFV_CHECK(src_kind == SRC_NHWC_U8 || src_kind == SRC_NCHW_F32, "conv_in: unsupported source layout");
  SFV_CHECK(N <= 65535, "conv_in: batch too large");
  const int tiles_x = ceil_div(W, TW), tiles = tiles_x * ceil_div(H, TH);
  const int gx = tiles < 160 ? tiles : 160;
  ProfScope prof(PROF_IGEMM, 2.0 * N * (double)H * W * 128 * 27, s);
  if (src_kind == SRC_NHWC_U8)
    conv_in_kernel<SRC_NHWC_U8><<<dim3(gx, N), 256, 0, s>>>(x, w, bias, y, stats, H, W, tiles_x, tiles, (uint16_t*)y16, fmt16, y16_scale);
  else
    conv_in_kernel<SRC_NCHW_F32><<<dim3(gx, N), 256, 0, s>>>(x, w, bias, y, stats, H, W, tiles_x, tiles, (uint16_t*)y16, fmt16, y16_scale);
  SFV_LAUNCH_OK();
  return 0;
}

}  // namespace sfv
