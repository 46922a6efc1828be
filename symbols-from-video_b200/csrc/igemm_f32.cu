// igemm_f32.cu -- CUDA-core fp32 implicit GEMM.
//
// Serves (a) the "fp32 check mode" of every convolution / 1x1 / attention GEMM
// (north-star tolerance 1e-4 needs a non-tensor-core path), (b) the layers whose
// shapes do not fit UMMA and are bandwidth-bound anyway: conv_in (K = 27, fed
// straight from fp32 NCHW or uint8 HWC frames with load_img's /255, 2x-1 fused
// in the gather; get_percep_embeddings.py:67-71) and the RBVAE convolutions
// (percep_RBVAE_model.py:50-58).
//
//   y[img, m, n] = alpha * sum_k A(img, m, k) * B(img, k, n) + bias[n] (+ residual) (ReLU)
//   m = oy*Wo + ox,  k = (r*ks + s)*Cin + ci,  A = x[img, oy*stride + r - pad, ox*stride + s - pad, ci]
//
// 64x64 output tile per 256-thread block, BK = 16, 4x4 register micro-tile.
#include "common.cuh"

namespace sfv {
namespace {

constexpr int BM = 64, BN = 64, BK = 16;

__device__ __forceinline__ float load_src(const IgemmArgs& a, int img, int iy, int ix, int ci) {
  if (iy < 0 || iy >= a.H || ix < 0 || ix >= a.W) return 0.f;
  if (a.src_kind == SRC_NHWC_F32) {
    return a.in_scale * reinterpret_cast<const float*>(a.x)[(((long long)img * a.H + iy) * a.W + ix) * a.Cin + ci];
  } else if (a.src_kind == SRC_NCHW_F32) {
    return a.in_scale * reinterpret_cast<const float*>(a.x)[(((long long)img * a.Cin + ci) * a.H + iy) * a.W + ix];
  } else {
    const uint8_t v = reinterpret_cast<const uint8_t*>(a.x)[(((long long)img * a.H + iy) * a.W + ix) * a.Cin + ci];
    const float f = (float)v / 255.0f;       // load_img: astype(float32)/255.0, then 2*x-1
    return 2.f * f - 1.f;
  }
}

__global__ void __launch_bounds__(256) igemm_f32_kernel(const IgemmArgs a) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int img = blockIdx.z;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int M = a.Ho * a.Wo;
  const int ksx = a.ksize_x ? a.ksize_x : a.ksize;      // kernel width (height = ksize)
  const int K = a.ksize * ksx * a.Cin;
  const float* Bp = a.w + (long long)img * a.w_batch;

  // A-gather assignment: thread -> (m = tid % 64, k = tid / 64 + 4 i)
  const int am = tid & 63;
  const int ak = tid >> 6;
  const int m = m0 + am;
  const bool m_ok = m < M;
  const int oy = m_ok ? m / a.Wo : 0;
  const int ox = m_ok ? m - oy * a.Wo : 0;
  // vectorised path: NHWC fp32 with Cin % 16 == 0 -> a BK chunk is 16 contiguous channels of one tap
  const bool vec = (a.src_kind == SRC_NHWC_F32) && (a.Cin % BK == 0);
  const int vm = tid >> 2;          // 64 rows x 4 float4
  const int vq = tid & 3;
  const int vmm = m0 + vm;
  const bool vm_ok = vmm < M;
  const int voy = vm_ok ? vmm / a.Wo : 0;
  const int vox = vm_ok ? vmm - voy * a.Wo : 0;

  const int tx = tid & 15;   // n micro-tile
  const int ty = tid >> 4;   // m micro-tile
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    if (vec) {
      const int tap = k0 / a.Cin;
      const int ci0 = k0 - tap * a.Cin;
      const int r = tap / ksx, s = tap - r * ksx;
      const int iy = voy * a.stride + r - a.pad, ix = vox * a.stride + s - a.pad;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (vm_ok && iy >= 0 && iy < a.H && ix >= 0 && ix < a.W) {
        v = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.x) +
                                             (((long long)img * a.H + iy) * a.W + ix) * a.Cin + ci0 + vq * 4);
        v.x *= a.in_scale; v.y *= a.in_scale; v.z *= a.in_scale; v.w *= a.in_scale;
      }
      As[vq * 4 + 0][vm] = v.x; As[vq * 4 + 1][vm] = v.y; As[vq * 4 + 2][vm] = v.z; As[vq * 4 + 3][vm] = v.w;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int kk = ak + 4 * i;
        const int k = k0 + kk;
        float v = 0.f;
        if (m_ok && k < K) {
          const int tap = k / a.Cin;
          const int ci = k - tap * a.Cin;
          const int r = tap / ksx, s = tap - r * ksx;
          v = load_src(a, img, oy * a.stride + r - a.pad, ox * a.stride + s - a.pad, ci);
        }
        As[kk][am] = v;
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int kk = ak + 4 * i;
      const int k = k0 + kk;
      const int n = n0 + am;
      float v = 0.f;
      if (k < K && n < a.Cout) v = Bp[(long long)k * a.w_sk + (long long)n * a.w_sn];
      Bs[kk][am] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float aa[4] = {av.x, av.y, av.z, av.w};
      const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int mm = m0 + ty * 4 + i;
    if (mm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= a.Cout) continue;
      float v = acc[i][j] * a.alpha;
      if (a.bias) v += a.bias[n];
      long long o;
      if (a.osy) {
        const int py = mm / a.Wo, px = mm - py * a.Wo;
        o = (((long long)img * a.oHf + (py * a.osy + a.ooy)) * a.oWf + (px * a.osx + a.oox)) * a.ldy + n;
      } else {
        o = a.nchw_out ? (((long long)img * a.Cout + n) * M + mm) : (((long long)img * M + mm) * a.ldy + n);
      }
      if (a.residual) v += a.residual[o];
      if (a.relu) v = fmaxf(v, 0.f);
      if (a.y) a.y[o] = v;
      if (a.y16) {
        if (a.err16 && a.fmt16 == FMT_F16 && !(fabsf(v) <= 65504.f)) atomicCAS(a.err16, 0, kErrRangeBase + SITE_XCOPY);
        reinterpret_cast<uint16_t*>(a.y16)[o] = f32_to_16(v, a.fmt16);
      }
    }
  }
}

}  // namespace

int launch_igemm_f32(const IgemmArgs& a, cudaStream_t s) {
  SFV_CHECK(a.N >= 1 && a.N <= 65535, "igemm: N=%d out of range", a.N);
  const int M = a.Ho * a.Wo;
  dim3 grid(ceil_div(M, BM), ceil_div(a.Cout, BN), a.N);
  SFV_CHECK(grid.y <= 65535, "igemm: Cout too large");
  ProfScope prof(PROF_IGEMM, 2.0 * M * (double)a.N * a.Cout * a.ksize * (a.ksize_x ? a.ksize_x : a.ksize) * a.Cin, s);
  igemm_f32_kernel<<<grid, 256, 0, s>>>(a);
  SFV_LAUNCH_OK();
  return 0;
}

}  // namespace sfv
