// norm.cu -- GroupNorm(32, eps=1e-6)+SiLU passes and the other HBM-bound
// elementwise kernels of the encoder (reference
// ldm/modules/diffusionmodules/model.py:33-39, distributions.py:24-37).
//
// All activations are NHWC so a thread owns 4 consecutive channels of one pixel
// (16 B fp32 / 8 B 16-bit vector access, coalesced across the warp) and those 4
// channels always fall in one group (C/32 is 4, 8 or 16).
// Statistics: fp32 per-thread partial sums -> fp64 shared/global atomics, so the
// E[x^2]-E[x]^2 subtraction happens in double.
#include "common.cuh"

namespace sfv {
namespace {

constexpr int kPixPerBlock = 512;

template <bool IS16>
__device__ __forceinline__ void load4(const void* base, long long idx, int fmt, float v[4]) {
  if (IS16) {
    const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(base) + idx);
    v[0] = f16_to_32((uint16_t)(u.x & 0xFFFF), fmt); v[1] = f16_to_32((uint16_t)(u.x >> 16), fmt);
    v[2] = f16_to_32((uint16_t)(u.y & 0xFFFF), fmt); v[3] = f16_to_32((uint16_t)(u.y >> 16), fmt);
  } else {
    const float4 f = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx);
    v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
  }
}

// grid (ceil(HW / kPixPerBlock), N); block 256.  VEC = 4 channels per thread.
template <bool IS16>
__global__ void __launch_bounds__(256) gn_stats_kernel(const void* x, int fmt, long long HW, int C,
                                                       int G, double* stats) {
  __shared__ double sh[64][2];
  const int n = blockIdx.y;
  const int tpp = C >> 2;                 // threads per pixel
  const int rows = 256 / tpp;             // pixels per pass (tpp in {8..64} for C<=256; C=512 -> 2)
  const int tc = threadIdx.x % tpp;
  const int tr = threadIdx.x / tpp;
  const int cpg = C / G;
  const int g = (tc * 4) / cpg;
  if (threadIdx.x < 2 * G) (&sh[0][0])[threadIdx.x] = 0.0;
  __syncthreads();
  const long long p0 = (long long)blockIdx.x * kPixPerBlock;
  const long long p1 = min(p0 + (long long)kPixPerBlock, HW);
  float s = 0.f, ss = 0.f;
  if (tr < rows) {
    for (long long p = p0 + tr; p < p1; p += rows) {
      float v[4];
      load4<IS16>(x, ((long long)n * HW + p) * C + tc * 4, fmt, v);
      s += (v[0] + v[1]) + (v[2] + v[3]);
      ss += (v[0] * v[0] + v[1] * v[1]) + (v[2] * v[2] + v[3] * v[3]);
    }
    atomicAdd(&sh[g][0], (double)s);
    atomicAdd(&sh[g][1], (double)ss);
  }
  __syncthreads();
  if (threadIdx.x < 2 * G)
    atomicAdd(&stats[(long long)n * G * 2 + threadIdx.x], (&sh[0][0])[threadIdx.x]);
}

// scalar fallback for channel counts whose groups are not multiples of 4 (op tests only)
__global__ void gn_stats_scalar_kernel(const float* x, long long HW, int C, int G, double* stats) {
  const int n = blockIdx.y, g = blockIdx.x;
  const int cpg = C / G;
  double s = 0, ss = 0;
  for (long long i = threadIdx.x; i < HW * cpg; i += blockDim.x) {
    const long long p = i / cpg; const int c = g * cpg + (int)(i % cpg);
    const float v = x[((long long)n * HW + p) * C + c];
    s += v; ss += (double)v * v;
  }
  atomicAdd(&stats[((long long)n * G + g) * 2], s);
  atomicAdd(&stats[((long long)n * G + g) * 2 + 1], ss);
}

__device__ __forceinline__ float silu(float v) { return v / (1.f + __expf(-v)); }

// SiLU for 16-bit outputs: t * sigmoid(t) = h + h * tanh(h), h = t / 2 -- ONE MUFU op (tanh.approx.f32, relative
// error 2^-11, i.e. the output format's own rounding step) instead of ex2 + rcp.  The apply pass moves 4 bytes per
// element with 16-bit input and output, which puts two MUFU ops per element (16 lanes/clk/SM) level with HBM time.
__device__ __forceinline__ float silu_tanh(float t) {
  const float h = 0.5f * t;
  float th;
  asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(h));
  return fmaf(h, th, h);
}

// fp16 range check on packed pairs: bit 15 of a half of the result is set iff that half's magnitude bits are
// >= 0x7BFF (65504 = the value a saturating conversion produces, inf, NaN).  OR-accumulated per thread.
__device__ __forceinline__ uint32_t f16_sat_bits(uint32_t w) { return (w & 0x7FFF7FFFu) + 0x04010401u; }

// GroupNorm apply (+SiLU): y = silu((x - mean_g) * rstd_g * gamma + beta).
// A thread owns 8 consecutive channels (32 B fp32 / 16 B 16-bit in, 16 B 16-bit out) and
// walks pixels with a 4-deep unrolled load batch so enough bytes are in flight to cover
// HBM latency.  grid (pixel slabs, N); block 256.
template <bool IN16>
__device__ __forceinline__ void load8(const void* base, long long idx, int fmt, float v[8]) {
  if (IN16) {
    const uint4 u = __ldcs(reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(base) + idx));
    unpack2_16(u.x, fmt, v[0], v[1]); unpack2_16(u.y, fmt, v[2], v[3]);
    unpack2_16(u.z, fmt, v[4], v[5]); unpack2_16(u.w, fmt, v[6], v[7]);
  } else {
    const float4 a = __ldcs(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx));
    const float4 b = __ldcs(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx + 4));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
}

template <bool IN16, bool OUT16>
__global__ void __launch_bounds__(256, 3) gn_apply_kernel(const void* __restrict__ x, const double* __restrict__ stats,
                                                          const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, void* __restrict__ y,
                                                          int fmt, long long HW, int C, int G, float eps,
                                                          int do_silu, int pix_per_block, int* err, float in_mul) {
  __shared__ float sh_mean[64], sh_rstd[64];
  const int n = blockIdx.y;
  const int cpg = C / G;
  const bool chk = err != nullptr && fmt == FMT_F16;
  uint32_t sat_in = 0, sat_out = 0;
  if (threadIdx.x < G) {
    const double cnt = (double)HW * cpg;
    const double m = stats[((long long)n * G + threadIdx.x) * 2] / cnt;
    double var = stats[((long long)n * G + threadIdx.x) * 2 + 1] / cnt - m * m;
    if (var < 0) var = 0;
    sh_mean[threadIdx.x] = (float)m;
    sh_rstd[threadIdx.x] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  const int tpp = C >> 3;                 // threads per pixel
  const int rows = 256 / tpp;
  const int tc = threadIdx.x % tpp;
  const int tr = threadIdx.x / tpp;
  float sc[8], sf[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = tc * 8 + j;
    const int g = c / cpg;
    const float gr = gamma[c] * sh_rstd[g];
    sc[j] = gr * in_mul;                       // the stored input is (true value) / in_mul (scaled 16-bit stream)
    sf[j] = beta[c] - sh_mean[g] * gr;
  }
  const long long p0 = (long long)blockIdx.x * pix_per_block;
  const long long p1 = min(p0 + (long long)pix_per_block, HW);
  const long long base = (long long)n * HW * C + tc * 8;
  // 128 bytes per thread in flight (x 1024 resident threads per SM): the raw words stay packed in
  // registers until they are consumed, so the kernel fits 64 registers / 4 blocks per SM.
  constexpr int U = IN16 ? 8 : 4;
  constexpr int W = IN16 ? 1 : 2;         // uint4 words per pixel per thread
  for (long long p = p0 + tr; p < p1; p += (long long)rows * U) {
    uint4 raw[U][W];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long pp = p + (long long)u * rows;
      if (pp < p1) {
        if (IN16) {
          raw[u][0] = __ldcs(reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(x) + base + pp * C));
        } else {
          const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(x) + base + pp * C);
          raw[u][0] = __ldcs(src);
          raw[u][W - 1] = __ldcs(src + 1);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long pp = p + (long long)u * rows;
      if (pp >= p1) break;
      float v[8];
      if (IN16) {
        if (chk) sat_in |= f16_sat_bits(raw[u][0].x) | f16_sat_bits(raw[u][0].y) | f16_sat_bits(raw[u][0].z) | f16_sat_bits(raw[u][0].w);
        unpack2_16(raw[u][0].x, fmt, v[0], v[1]); unpack2_16(raw[u][0].y, fmt, v[2], v[3]);
        unpack2_16(raw[u][0].z, fmt, v[4], v[5]); unpack2_16(raw[u][0].w, fmt, v[6], v[7]);
      } else {
        v[0] = __uint_as_float(raw[u][0].x); v[1] = __uint_as_float(raw[u][0].y);
        v[2] = __uint_as_float(raw[u][0].z); v[3] = __uint_as_float(raw[u][0].w);
        v[4] = __uint_as_float(raw[u][W - 1].x); v[5] = __uint_as_float(raw[u][W - 1].y);
        v[6] = __uint_as_float(raw[u][W - 1].z); v[7] = __uint_as_float(raw[u][W - 1].w);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float t = fmaf(v[j], sc[j], sf[j]);
        v[j] = do_silu ? (OUT16 ? silu_tanh(t) : __fdividef(t, 1.f + __expf(-t))) : t;
      }
      const long long idx = base + pp * C;
      if (OUT16) {
        uint4 o;
        o.x = pack2_16(v[0], v[1], fmt); o.y = pack2_16(v[2], v[3], fmt);
        o.z = pack2_16(v[4], v[5], fmt); o.w = pack2_16(v[6], v[7], fmt);
        if (chk) sat_out |= f16_sat_bits(o.x) | f16_sat_bits(o.y) | f16_sat_bits(o.z) | f16_sat_bits(o.w);
        *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(y) + idx) = o;
      } else {
        float* o = reinterpret_cast<float*>(y) + idx;
        *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
      }
    }
  }
  if (chk) {
    if (sat_in & 0x80008000u) atomicCAS(err, 0, kErrRangeBase + SITE_GN_IN);
    if (sat_out & 0x80008000u) atomicCAS(err, 0, kErrRangeBase + SITE_GN_OUT);
  }
}

// ---- bulk-async fed variant -----------------------------------------------------
// A block's slab of pixels is one contiguous byte range in NHWC, so the input is streamed
// through a 4-slot shared-memory ring with cp.async.bulk (the TMA engine keeps 64 KB per block in
// flight with no registers or issue slots spent on address generation); the 256 threads normalise
// from shared memory and write 16-byte coalesced outputs.  grid (slabs, N), block 256, 64 KB smem.
constexpr int kBulkSlots = 4;
constexpr int kBulkPieceBytes = 16384;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One ring piece of the 16-bit -> 16-bit apply pass with the format, the range check and the activation fixed at
// compile time: under the power cap the SMs run at ~1.4 GHz and the general loop's ~100 instructions per 8 elements
// (format / check / activation branches, 64-bit address arithmetic, two add-and-mask range checks) made this
// HBM-bound pass partly issue-bound.  Here: ~50.  With SiLU the 1/2 of h = t/2 is folded into scale and shift
// (exact: a power of two), so h is the FMA's result; range check = running |max| of the packed halves (NaN-propagating),
// compared once at the end.  Bit-identical to the general loop.
template <int FMT, bool CHK, bool SILU>
__device__ __forceinline__ void gn_piece16(const uint8_t* __restrict__ buf, int C, int tc, int tr, int rows, int cntp,
                                           const float (&sc)[8], const float (&sf)[8], uint4* __restrict__ yrow, long long ystep,
                                           __half2& mxi, __half2& mxo) {
  const uint8_t* src = buf + ((size_t)tr * C + tc * 8) * 2;
  const size_t sstep = (size_t)rows * C * 2;
#pragma unroll 2
  for (int r = tr; r < cntp; r += rows, src += sstep, yrow += ystep) {
    const uint4 u = *reinterpret_cast<const uint4*>(src);
    if (CHK) {
      mxi = __hmax2_nan(__hmax2_nan(mxi, __habs2(*reinterpret_cast<const __half2*>(&u.x))), __habs2(*reinterpret_cast<const __half2*>(&u.y)));
      mxi = __hmax2_nan(__hmax2_nan(mxi, __habs2(*reinterpret_cast<const __half2*>(&u.z))), __habs2(*reinterpret_cast<const __half2*>(&u.w)));
    }
    float v[8];
    unpack2_16(u.x, FMT, v[0], v[1]); unpack2_16(u.y, FMT, v[2], v[3]);
    unpack2_16(u.z, FMT, v[4], v[5]); unpack2_16(u.w, FMT, v[6], v[7]);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float h = fmaf(v[j], sc[j], sf[j]);            // SILU: sc, sf already carry the 1/2
      if (SILU) {
        float th;
        asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(h));
        v[j] = fmaf(h, th, h);
      } else {
        v[j] = h;
      }
    }
    uint4 o;
    o.x = pack2_16(v[0], v[1], FMT); o.y = pack2_16(v[2], v[3], FMT);
    o.z = pack2_16(v[4], v[5], FMT); o.w = pack2_16(v[6], v[7], FMT);
    if (CHK) {
      mxo = __hmax2_nan(__hmax2_nan(mxo, __habs2(*reinterpret_cast<const __half2*>(&o.x))), __habs2(*reinterpret_cast<const __half2*>(&o.y)));
      mxo = __hmax2_nan(__hmax2_nan(mxo, __habs2(*reinterpret_cast<const __half2*>(&o.z))), __habs2(*reinterpret_cast<const __half2*>(&o.w)));
    }
    *yrow = o;
  }
}

template <bool IN16, bool OUT16>
__global__ void __launch_bounds__(256, 3) gn_apply_bulk_kernel(const void* __restrict__ x, const double* __restrict__ stats,
                                                               const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, void* __restrict__ y,
                                                               int fmt, long long HW, int C, int G, float eps,
                                                               int do_silu, int pix_per_block, int* err, int* wd_err, float in_mul) {
  extern __shared__ __align__(128) uint8_t ring[];
  __shared__ float sh_mean[64], sh_rstd[64];
  __shared__ __align__(8) uint64_t full[kBulkSlots];
  const int n = blockIdx.y;
  const int cpg = C / G;
  const bool chk = err != nullptr && fmt == FMT_F16;
  uint32_t sat_in = 0, sat_out = 0;
  constexpr int ES = IN16 ? 2 : 4;
  const int pix_bytes = C * ES;
  const int ppp = kBulkPieceBytes / pix_bytes;            // pixels per piece
  const long long p0 = (long long)blockIdx.x * pix_per_block;
  const long long p1 = min(p0 + (long long)pix_per_block, HW);
  const int npieces = (int)((p1 - p0 + ppp - 1) / ppp);
  const uint8_t* src = reinterpret_cast<const uint8_t*>(x) + ((long long)n * HW + p0) * pix_bytes;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kBulkSlots; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&full[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < G) {
    const double cnt = (double)HW * cpg;
    const double m = stats[((long long)n * G + threadIdx.x) * 2] / cnt;
    double var = stats[((long long)n * G + threadIdx.x) * 2 + 1] / cnt - m * m;
    if (var < 0) var = 0;
    sh_mean[threadIdx.x] = (float)m;
    sh_rstd[threadIdx.x] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  auto issue = [&](int piece) {
    const long long pp0 = (long long)piece * ppp;
    const long long cntp = min((long long)ppp, (p1 - p0) - pp0);
    const uint32_t bytes = (uint32_t)(cntp * pix_bytes);
    const uint32_t bar = smem_addr(&full[piece % kBulkSlots]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(ring + (piece % kBulkSlots) * kBulkPieceBytes)), "l"(src + pp0 * pix_bytes), "r"(bytes), "r"(bar)
                 : "memory");
  };
  if (threadIdx.x == 0)
    for (int i = 0; i < kBulkSlots && i < npieces; ++i) issue(i);
  const int tpp = C >> 3;
  const int rows = 256 / tpp;
  const int tc = threadIdx.x % tpp;
  const int tr = threadIdx.x / tpp;
  float sc[8], sf[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = tc * 8 + j;
    const int g = c / cpg;
    const float gr = gamma[c] * sh_rstd[g];
    sc[j] = gr * in_mul;                       // the stored input is (true value) / in_mul (scaled 16-bit stream)
    sf[j] = beta[c] - sh_mean[g] * gr;
  }
  // lean 16-bit -> 16-bit path (see gn_piece16): tables with the SiLU's 1/2 folded in, packed-half running maxima
  const bool lean = IN16 && OUT16;
  float sch[8], sfh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { sch[j] = do_silu ? 0.5f * sc[j] : sc[j]; sfh[j] = do_silu ? 0.5f * sf[j] : sf[j]; }
  __half2 mxi = __float2half2_rn(0.f), mxo = __float2half2_rn(0.f);
  for (int piece = 0; piece < npieces; ++piece) {
    const int slot = piece % kBulkSlots;
    const uint32_t bar = smem_addr(&full[slot]);
    const uint32_t parity = (uint32_t)((piece / kBulkSlots) & 1);
    // bounded wait (~2 s): a bulk copy that never completes must surface as an error, not hang the GPU
    uint32_t done = 0;
    unsigned long long t0 = 0;
    for (uint32_t spin = 0; !done; ++spin) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(bar), "r"(parity) : "memory");
      if (!done && (spin & 1023u) == 1023u) {
        if (t0 == 0) t0 = clock64();
        else if (clock64() - t0 > 4000000000ull) { atomicCAS(wd_err, 0, 6); return; }
      }
    }
    const long long pp0 = (long long)piece * ppp;
    const int cntp = (int)min((long long)ppp, (p1 - p0) - pp0);
    const uint8_t* buf = ring + slot * kBulkPieceBytes;
    if (lean) {
      uint4* yrow = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(y) + ((long long)n * HW + p0 + pp0 + tr) * C + tc * 8);
      const long long ystep = (long long)rows * C / 8;
      if (fmt == FMT_F16) {
        if (chk) {
          if (do_silu) gn_piece16<FMT_F16, true, true>(buf, C, tc, tr, rows, cntp, sch, sfh, yrow, ystep, mxi, mxo);
          else gn_piece16<FMT_F16, true, false>(buf, C, tc, tr, rows, cntp, sch, sfh, yrow, ystep, mxi, mxo);
        } else {
          if (do_silu) gn_piece16<FMT_F16, false, true>(buf, C, tc, tr, rows, cntp, sch, sfh, yrow, ystep, mxi, mxo);
          else gn_piece16<FMT_F16, false, false>(buf, C, tc, tr, rows, cntp, sch, sfh, yrow, ystep, mxi, mxo);
        }
      } else {
        if (do_silu) gn_piece16<FMT_BF16, false, true>(buf, C, tc, tr, rows, cntp, sch, sfh, yrow, ystep, mxi, mxo);
        else gn_piece16<FMT_BF16, false, false>(buf, C, tc, tr, rows, cntp, sch, sfh, yrow, ystep, mxi, mxo);
      }
    } else
    for (int r = tr; r < cntp; r += rows) {
      float v[8];
      if (IN16) {
        const uint4 u = *reinterpret_cast<const uint4*>(buf + (r * C + tc * 8) * 2);
        if (chk) sat_in |= f16_sat_bits(u.x) | f16_sat_bits(u.y) | f16_sat_bits(u.z) | f16_sat_bits(u.w);
        unpack2_16(u.x, fmt, v[0], v[1]); unpack2_16(u.y, fmt, v[2], v[3]);
        unpack2_16(u.z, fmt, v[4], v[5]); unpack2_16(u.w, fmt, v[6], v[7]);
      } else {
        const float4 a = *reinterpret_cast<const float4*>(buf + (r * C + tc * 8) * 4);
        const float4 b = *reinterpret_cast<const float4*>(buf + (r * C + tc * 8) * 4 + 16);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float t = fmaf(v[j], sc[j], sf[j]);
        v[j] = do_silu ? (OUT16 ? silu_tanh(t) : __fdividef(t, 1.f + __expf(-t))) : t;
      }
      const long long idx = ((long long)n * HW + p0 + pp0 + r) * C + tc * 8;
      if (OUT16) {
        uint4 o;
        o.x = pack2_16(v[0], v[1], fmt); o.y = pack2_16(v[2], v[3], fmt);
        o.z = pack2_16(v[4], v[5], fmt); o.w = pack2_16(v[6], v[7], fmt);
        if (chk) sat_out |= f16_sat_bits(o.x) | f16_sat_bits(o.y) | f16_sat_bits(o.z) | f16_sat_bits(o.w);
        *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(y) + idx) = o;
      } else {
        float* o = reinterpret_cast<float*>(y) + idx;
        *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
      }
    }
    __syncthreads();                       // everyone is done reading this slot
    if (threadIdx.x == 0 && piece + kBulkSlots < npieces) issue(piece + kBulkSlots);
  }
  if (chk) {
    if (sat_in & 0x80008000u) atomicCAS(err, 0, kErrRangeBase + SITE_GN_IN);
    if (sat_out & 0x80008000u) atomicCAS(err, 0, kErrRangeBase + SITE_GN_OUT);
    if (lean) {      // |x| >= 65504 (what a saturating conversion leaves), inf or NaN in either half
      if (!(__low2float(mxi) < 65504.f) || !(__high2float(mxi) < 65504.f)) atomicCAS(err, 0, kErrRangeBase + SITE_GN_IN);
      if (!(__low2float(mxo) < 65504.f) || !(__high2float(mxo) < 65504.f)) atomicCAS(err, 0, kErrRangeBase + SITE_GN_OUT);
    }
  }
}

__global__ void gn_apply_scalar_kernel(const float* x, const double* stats, const float* gamma,
                                       const float* beta, float* y, long long HW, int C, int G,
                                       float eps, int do_silu, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const long long n = i / ((long long)HW * C);
  const int cpg = C / G, g = c / cpg;
  const double cnt = (double)HW * cpg;
  const double m = stats[(n * G + g) * 2] / cnt;
  double var = stats[(n * G + g) * 2 + 1] / cnt - m * m;
  if (var < 0) var = 0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  float t = (x[i] - (float)m) * rstd * gamma[c] + beta[c];
  y[i] = do_silu ? silu(t) : t;
}

__global__ void f32_to_16_kernel(const float* x, uint16_t* y, long long n, int fmt) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 f = *reinterpret_cast<const float4*>(x + i);
    uint2 u; u.x = pack2_16(f.x, f.y, fmt); u.y = pack2_16(f.z, f.w, fmt);
    *reinterpret_cast<uint2*>(y + i) = u;
  } else {
    for (long long j = i; j < n; ++j) y[j] = f32_to_16(x[j], fmt);
  }
}

__global__ void f16_to_32_kernel(const uint16_t* x, float* y, long long n, int fmt, float mul) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = f16_to_32(x[i], fmt) * mul;
}

// one block per row; cols up to 16384 (mid-block attention at 1024^2)
template <bool OUT16>
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* x, void* y, int fmt, int cols) {
  __shared__ float red[8];
  __shared__ float bcast;
  const float* r = x + (long long)blockIdx.x * cols;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float mx = -INFINITY;
  for (int c = threadIdx.x * 4; c < cols; c += 1024) {
    const float4 f = *reinterpret_cast<const float4*>(r + c);
    mx = fmaxf(mx, fmaxf(fmaxf(f.x, f.y), fmaxf(f.z, f.w)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = red[0];
    for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
    bcast = m;
  }
  __syncthreads();
  mx = bcast;
  float sum = 0.f;
  for (int c = threadIdx.x * 4; c < cols; c += 1024) {
    const float4 f = *reinterpret_cast<const float4*>(r + c);
    sum += (expf(f.x - mx) + expf(f.y - mx)) + (expf(f.z - mx) + expf(f.w - mx));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  __syncthreads();
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i];
    bcast = 1.f / s;
  }
  __syncthreads();
  const float inv = bcast;
  for (int c = threadIdx.x * 4; c < cols; c += 1024) {
    const float4 f = *reinterpret_cast<const float4*>(r + c);
    const float a = expf(f.x - mx) * inv, b = expf(f.y - mx) * inv, cc = expf(f.z - mx) * inv,
                d = expf(f.w - mx) * inv;
    const long long o = (long long)blockIdx.x * cols + c;
    if (OUT16) {
      uint2 u; u.x = pack2_16(a, b, fmt); u.y = pack2_16(cc, d, fmt);
      *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(y) + o) = u;
    } else {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + o) = make_float4(a, b, cc, d);
    }
  }
}

// any row length / row pitch (token counts that are not a multiple of 8, tiny latents in check mode):
// one warp per row, scalar accesses; rows of x have pitch ldx, rows of y pitch ldy (elements)
__global__ void softmax_rows_scalar_kernel(const float* x, long long ldx, void* y, long long ldy, int y_is16, int fmt,
                                           long long rows, int cols) {
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + r * ldx;
  float mx = -INFINITY;
  for (int c = lane; c < cols; c += 32) mx = fmaxf(mx, xr[c]);
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int c = lane; c < cols; c += 32) sum += expf(xr[c] - mx);
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.f / sum;
  for (int c = lane; c < cols; c += 32) {
    const float v = expf(xr[c] - mx) * inv;
    if (y_is16) reinterpret_cast<uint16_t*>(y)[r * ldy + c] = (uint16_t)(pack2_16(v, 0.f, fmt) & 0xffffu);
    else reinterpret_cast<float*>(y)[r * ldy + c] = v;
  }
}

// DiagonalGaussian head (distributions.py:24-31): moments NHWC [N,HW,8] ->
//   parameters NCHW [N,8,HW] (raw; mean = channels 0..3), logvar = clamp(ch 4..7, -30, 20),
//   std = exp(0.5 logvar), var = exp(logvar)   (each NCHW [N,4,HW]; std/var optional)
__global__ void head_kernel(const float* mom, float* params, float* logvar, float* stdv, float* var, int HW,
                            long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // (n, p)
  if (i >= total) return;
  const long long n = i / HW; const int p = (int)(i - n * HW);
  const float4 a = *reinterpret_cast<const float4*>(mom + i * 8);
  const float4 b = *reinterpret_cast<const float4*>(mom + i * 8 + 4);
  const float m[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int c = 0; c < 8; ++c) params[(n * 8 + c) * HW + p] = m[c];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float lv = fminf(fmaxf(m[4 + c], -30.f), 20.f);
    const long long o = (n * 4 + c) * HW + p;
    logvar[o] = lv;
    if (stdv) stdv[o] = expf(0.5f * lv);
    if (var) var[o] = expf(lv);
  }
}

__global__ void sample_kernel(const float* mean, const float* logvar, const float* noise, float scale,
                              float* out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = mean[i];
  if (noise) v += expf(0.5f * logvar[i]) * noise[i];
  out[i] = scale * v;
}

}  // namespace

int launch_zero(void* p, size_t bytes, cudaStream_t s) {
  SFV_CUDA(cudaMemsetAsync(p, 0, bytes, s));
  return 0;
}

int launch_gn_stats(const void* x, int x_is16, int fmt, int N, long long HW, int C, int G, double* stats,
                    cudaStream_t s) {
  SFV_CHECK(G >= 1 && C % G == 0, "group_norm: bad C=%d G=%d", C, G);
  SFV_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * N * G, s));
  const int cpg = C / G;
  ProfScope prof(PROF_GN_STATS, (double)N * HW * C * (x_is16 ? 2 : 4), s);
  if (cpg % 4 == 0 && C % 4 == 0 && 256 % (C / 4) == 0 && C <= 1024 && G <= 32) {
    dim3 grid((unsigned)((HW + kPixPerBlock - 1) / kPixPerBlock), N);
    if (x_is16) gn_stats_kernel<true><<<grid, 256, 0, s>>>(x, fmt, HW, C, G, stats);
    else gn_stats_kernel<false><<<grid, 256, 0, s>>>(x, fmt, HW, C, G, stats);
  } else {
    SFV_CHECK(!x_is16, "group_norm: scalar path is fp32 only");
    gn_stats_scalar_kernel<<<dim3(G, N), 256, 0, s>>>((const float*)x, HW, C, G, stats);
  }
  SFV_LAUNCH_OK();
  return 0;
}

int launch_gn_apply(const void* x, int x_is16, const double* stats, const float* gamma, const float* beta,
                    void* y, int y_is16, int fmt, int N, long long HW, int C, int G, float eps, int silu,
                    cudaStream_t s, int range_check, float in_mul) {
  const int cpg = C / G;
  DevState* ds = nullptr;
  SFV_TRY(dev_state(&ds));
  int* err = range_check ? ds->err_flag : nullptr;
  int* wd = ds->err_flag;
  char tag[64];
  snprintf(tag, sizeof(tag), "gn N=%d HW=%lld C=%d in%d out%d", N, HW, C, x_is16 ? 16 : 32, y_is16 ? 16 : 32);
  ProfScope prof(PROF_GN_APPLY, (double)N * HW * C * ((x_is16 ? 2 : 4) + (y_is16 ? 2 : 4)), s, tag);
  if (C % 8 == 0 && 256 % (C / 8) == 0 && C <= 2048 && G <= 64 && C % G == 0) {
    // slabs sized so that the grid is a few waves of 148 SMs x 8 resident blocks
    long long want = (HW * N + 148 * 16 - 1) / (148 * 16);
    int ppb = (int)(want < 64 ? 64 : (want > 1024 ? 1024 : want));
    const int rows = 256 / (C / 8);
    ppb = (ppb + rows * 8 - 1) / (rows * 8) * (rows * 8);
    dim3 grid((unsigned)((HW + ppb - 1) / ppb), N);
    static int bulk = -1;
    if (bulk < 0) { const char* e = getenv("SFV_GN_BULK"); bulk = e ? atoi(e) : 1; }
    const int pix_bytes = C * (x_is16 ? 2 : 4);
    if (bulk && kBulkPieceBytes % pix_bytes == 0 && ((uintptr_t)x & 15) == 0) {
      static unsigned long long attr_devs = 0;      // cudaFuncSetAttribute is per device
      if (first_use_on_device(attr_devs, ds->dev)) {
        SFV_CUDA(cudaFuncSetAttribute(gn_apply_bulk_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBulkSlots * kBulkPieceBytes));
        SFV_CUDA(cudaFuncSetAttribute(gn_apply_bulk_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBulkSlots * kBulkPieceBytes));
        SFV_CUDA(cudaFuncSetAttribute(gn_apply_bulk_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBulkSlots * kBulkPieceBytes));
        SFV_CUDA(cudaFuncSetAttribute(gn_apply_bulk_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBulkSlots * kBulkPieceBytes));
      }
      const size_t sm = kBulkSlots * kBulkPieceBytes;
      if (x_is16 && y_is16) gn_apply_bulk_kernel<true, true><<<grid, 256, sm, s>>>(x, stats, gamma, beta, y, fmt, HW, C, G, eps, silu, ppb, err, wd, in_mul);
      else if (!x_is16 && y_is16) gn_apply_bulk_kernel<false, true><<<grid, 256, sm, s>>>(x, stats, gamma, beta, y, fmt, HW, C, G, eps, silu, ppb, err, wd, in_mul);
      else if (x_is16 && !y_is16) gn_apply_bulk_kernel<true, false><<<grid, 256, sm, s>>>(x, stats, gamma, beta, y, fmt, HW, C, G, eps, silu, ppb, err, wd, in_mul);
      else gn_apply_bulk_kernel<false, false><<<grid, 256, sm, s>>>(x, stats, gamma, beta, y, fmt, HW, C, G, eps, silu, ppb, err, wd, in_mul);
      SFV_LAUNCH_OK();
      return 0;
    }
    if (x_is16 && y_is16) gn_apply_kernel<true, true><<<grid, 256, 0, s>>>(x, stats, gamma, beta, y, fmt, HW, C, G, eps, silu, ppb, err, in_mul);
    else if (!x_is16 && y_is16) gn_apply_kernel<false, true><<<grid, 256, 0, s>>>(x, stats, gamma, beta, y, fmt, HW, C, G, eps, silu, ppb, err, in_mul);
    else if (x_is16 && !y_is16) gn_apply_kernel<true, false><<<grid, 256, 0, s>>>(x, stats, gamma, beta, y, fmt, HW, C, G, eps, silu, ppb, err, in_mul);
    else gn_apply_kernel<false, false><<<grid, 256, 0, s>>>(x, stats, gamma, beta, y, fmt, HW, C, G, eps, silu, ppb, err, in_mul);
  } else {
    SFV_CHECK(!x_is16 && !y_is16, "group_norm: scalar path is fp32 only");
    const long long total = (long long)N * HW * C;
    gn_apply_scalar_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(
        (const float*)x, stats, gamma, beta, (float*)y, HW, C, G, eps, silu, total);
  }
  SFV_LAUNCH_OK();
  (void)cpg;
  return 0;
}

int launch_f32_to_16(const float* x, void* y, long long n, int fmt, cudaStream_t s) {
  const long long nt = (n + 3) / 4;
  f32_to_16_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, s>>>(x, (uint16_t*)y, n, fmt);
  SFV_LAUNCH_OK();
  return 0;
}

int launch_16_to_f32_scaled(const void* x, float* y, long long n, int fmt, float mul, cudaStream_t s) {
  f16_to_32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>((const uint16_t*)x, y, n, fmt, mul);
  SFV_LAUNCH_OK();
  return 0;
}
int launch_16_to_f32(const void* x, float* y, long long n, int fmt, cudaStream_t s) {
  return launch_16_to_f32_scaled(x, y, n, fmt, 1.f, s);
}

int launch_softmax_rows(const float* x, void* y, int y_is16, int fmt, long long rows, int cols,
                        cudaStream_t s, long long ld) {
  SFV_CHECK(rows < (1ll << 31), "softmax: too many rows");
  if (ld <= 0) ld = cols;
  if (cols % 4 != 0 || ld != cols) {
    ProfScope prof(PROF_SOFTMAX, (double)rows * cols * (4 + (y_is16 ? 2 : 4)), s);
    softmax_rows_scalar_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(x, ld, y, ld, y_is16, fmt, rows, cols);
    SFV_LAUNCH_OK();
    return 0;
  }
  ProfScope prof(PROF_SOFTMAX, (double)rows * cols * (4 + (y_is16 ? 2 : 4)), s);
  if (y_is16) softmax_rows_kernel<true><<<(unsigned)rows, 256, 0, s>>>(x, y, fmt, cols);
  else softmax_rows_kernel<false><<<(unsigned)rows, 256, 0, s>>>(x, y, fmt, cols);
  SFV_LAUNCH_OK();
  return 0;
}

int launch_head(const float* mom, float* params, float* logvar, float* stdv, float* var, int N, int HW,
                cudaStream_t s) {
  const long long total = (long long)N * HW;
  head_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(mom, params, logvar, stdv, var, HW, total);
  SFV_LAUNCH_OK();
  return 0;
}

int launch_sample(const float* mean, const float* logvar, const float* noise, float scale, float* out,
                  long long n, cudaStream_t s) {
  sample_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(mean, logvar, noise, scale, out, n);
  SFV_LAUNCH_OK();
  return 0;
}

}  // namespace sfv
