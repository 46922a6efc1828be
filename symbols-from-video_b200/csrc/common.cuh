// common.cuh -- shared host/device helpers for libsfv (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include <map>
#include <atomic>

#include "../../include/sfv.h"

namespace sfv {

extern thread_local std::string g_last_error;
extern std::atomic<long long> g_launches;

int fail(int code, const char* fmt, ...);

#define SFV_CUDA(expr)                                                              \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess)                                                          \
      return sfv::fail(SFV_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,   \
                       cudaGetErrorString(_e));                                     \
  } while (0)

#define SFV_TRY(expr)            \
  do {                           \
    int _s = (expr);             \
    if (_s != 0) return _s;      \
  } while (0)

#define SFV_CHECK(cond, ...)                                    \
  do {                                                          \
    if (!(cond)) return sfv::fail(SFV_ERR_INVALID, __VA_ARGS__); \
  } while (0)

// every kernel launch goes through this so the library can report a launch count
#define SFV_LAUNCH_OK()                                                                    \
  do {                                                                                     \
    sfv::g_launches.fetch_add(1, std::memory_order_relaxed);                               \
    cudaError_t _e = cudaGetLastError();                                                   \
    if (_e != cudaSuccess)                                                                 \
      return sfv::fail(SFV_ERR_CUDA, "%s:%d launch -> %s", __FILE__, __LINE__,             \
                       cudaGetErrorString(_e));                                            \
  } while (0)

// ---- optional per-kernel-class timing (bench.py roofline): CUDA events on the launching stream ----
enum : int { PROF_TC_GEMM = 0, PROF_IGEMM = 1, PROF_GN_STATS = 2, PROF_GN_APPLY = 3, PROF_SOFTMAX = 4,
             PROF_OTHER = 5, PROF_NUM = 6 };
extern bool g_prof_on;
void prof_begin(int cat, double work, cudaStream_t s, const char* tag);
void prof_end(cudaStream_t s);
struct ProfScope {
  cudaStream_t s; bool on;
  ProfScope(int cat, double work, cudaStream_t st, const char* tag = nullptr) : s(st), on(g_prof_on) {
    if (on) prof_begin(cat, work, s, tag);
  }
  ~ProfScope() { if (on) prof_end(s); }
};

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline long long ceil_div(long long a, long long b) { return (a + b - 1) / b; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// 16-bit operand formats of the tensor-core path.
enum : int { FMT_F16 = 0, FMT_BF16 = 1 };   // == UMMA F16F32Format encoding
// single-format view of a precision (op entry points, RBVAE convs): MIXED behaves as fp16 there
inline int fmt_of_precision(int prec) { return (prec == SFV_PREC_FP16 || prec == SFV_PREC_MIXED) ? FMT_F16 : FMT_BF16; }

// ---- per-device library state ------------------------------------------------
// One record per CUDA device ordinal: the device-side error flag (pipeline watchdog / fp16 range checks) and the SM
// count.  Kernel attributes (cudaFuncSetAttribute is per device) are tracked by per-kernel device bit masks.
constexpr int kMaxDevices = 64;
struct DevState {
  int dev = -1;
  int* err_flag = nullptr;     // device int: 0 ok, 1..5 watchdog role code, >= kErrRangeBase fp16 range violation site
  int num_sms = 0;
};
int dev_state(DevState** out);                // state of the CURRENT device (created on first use)
constexpr int kErrRangeBase = 64;             // err_flag = kErrRangeBase + site
enum : int { SITE_GN_OUT = 1, SITE_GN_IN = 2, SITE_XCOPY = 3 };
// true the first time it is called for (mask, current device): the caller then sets its per-device kernel attributes
inline bool first_use_on_device(unsigned long long& mask, int dev) {
  if (dev < 0 || dev >= kMaxDevices) return true;
  if ((mask >> dev) & 1ull) return false;
  mask |= 1ull << dev;
  return true;
}

// ---- device helpers -------------------------------------------------------
__device__ __forceinline__ uint16_t f32_to_16(float v, int fmt) {
  if (fmt == FMT_BF16) {
    __nv_bfloat16 b = __float2bfloat16_rn(v);
    return *reinterpret_cast<uint16_t*>(&b);
  } else {
    // saturate instead of overflowing to inf: operands are bounded activations
    v = fminf(fmaxf(v, -65504.f), 65504.f);
    __half h = __float2half_rn(v);
    return *reinterpret_cast<uint16_t*>(&h);
  }
}
__device__ __forceinline__ float f16_to_32(uint16_t v, int fmt) {
  if (fmt == FMT_BF16) return __uint_as_float(((uint32_t)v) << 16);
  __half h = *reinterpret_cast<__half*>(&v);
  return __half2float(h);
}
// two fp32 -> packed 16-bit pair (a in the low half), one F2FP instruction; fp16 saturates at +-65504
__device__ __forceinline__ uint32_t pack2_16(float a, float b, int fmt) {
  uint32_t r;
  if (fmt == FMT_BF16) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  else asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
// packed 16-bit pair -> two fp32
__device__ __forceinline__ void unpack2_16(uint32_t w, int fmt, float& a, float& b) {
  if (fmt == FMT_BF16) {
    a = __uint_as_float(w << 16);
    b = __uint_as_float(w & 0xFFFF0000u);
  } else {
    const __half2 h = *reinterpret_cast<const __half2*>(&w);
    const float2 f = __half22float2(h);
    a = f.x; b = f.y;
  }
}

// ---- bump allocator over the caller's workspace ----------------------------
struct Arena {
  char* base; size_t cap; size_t off;
  Arena(void* p, size_t bytes) : base((char*)p), cap(bytes), off(0) {}
  void* take(size_t bytes) {
    size_t o = align_up(off, 1024);
    off = o + bytes;
    return base ? (void*)(base + o) : nullptr;   // base == nullptr: sizing pass
  }
  bool ok() const { return base == nullptr || off <= cap; }
};

// ---- igemm (CUDA-core fp32 implicit GEMM) -----------------------------------
enum : int { SRC_NHWC_F32 = 0, SRC_NCHW_F32 = 1, SRC_NHWC_U8 = 2 };
struct IgemmArgs {
  const void* x; int src_kind;
  const float* w;            // [K = kh*kw*Cin][Cout] fp32, or strided B
  long long w_sk, w_sn;      // strides of B(k,n); conv: (Cout, 1)
  long long w_batch;         // per-image stride of B (attention), 0 for conv
  const float* bias;         // [Cout] or null
  const float* residual;     // fp32, same layout as y, or null
  float* y;                  // fp32 [N, Ho*Wo, ldy] (+ y_off)
  void* y16;                 // optional 16-bit copy
  int fmt16;
  int N, H, W, Cin, Ho, Wo, Cout;
  int ksize, stride, pad;    // pad = top/left padding (bottom/right implied by Ho/Wo)
  int relu;
  float alpha;               // scale applied to the accumulator before bias
  float in_scale;            // scale applied to x on load
  long long ldy;             // output row pitch (elements)
  int nchw_out;              // write y as NCHW [N,Cout,Ho,Wo] instead
  int* err16;                // with fmt16 == FMT_F16: flag a y16 store beyond the fp16 range here (range-checked modes)
  int ksize_x;               // kernel width when it differs from ksize (= height); 0: square
  // strided output (sub-pixel phases of a transposed convolution): pixel (oy, ox) of this launch is written at
  // (oy * osy + ooy, ox * osx + oox) of an [oHf, oWf] image; osy == 0: dense [Ho, Wo] output
  int osy, osx, ooy, oox, oHf, oWf;
};
int launch_igemm_f32(const IgemmArgs& a, cudaStream_t s);
// y16 != nullptr: write to16(y16_scale * value) there (format fmt16) instead of the fp32 y
int launch_conv_in(const void* x, int src_kind, const float* w, const float* bias, float* y, double* stats,
                   int N, int H, int W, cudaStream_t s, void* y16 = nullptr, int fmt16 = 0, float y16_scale = 1.f);
int launch_16_to_f32_scaled(const void* x, float* y, long long n, int fmt, float mul, cudaStream_t s);
int launch_rb_conv0(const float* x, const float* w, const float* bias, void* y, int y16, int fmt, int N, int H, int W,
                    float in_scale, cudaStream_t s);
// contrastive conv.0 on the tensor pipe, step 1: fp32 NCHW frames -> A[N*Ho*Wo][64] 16-bit rows of the stride-2 3x3 patch,
// k < 27: hi(x), 27..53: lo(x) = x - hi (so the 16-bit operand carries ~fp32 input precision), 54..63: 0
int launch_rb_im2col(const float* x, void* a16, int fmt, int N, int H, int W, float in_scale, cudaStream_t s);

// ---- norm / elementwise ------------------------------------------------------
// stats: double [N][G][2] accumulators (sum, sumsq) -- zeroed by the caller/kernel.
int launch_gn_stats(const void* x, int x_is16, int fmt, int N, long long HW, int C, int G,
                    double* stats, cudaStream_t s);
// fmt: 16-bit format of x (if x_is16) and of y (if y_is16).  range_check: with fmt == FMT_F16, flag 16-bit inputs /
// outputs at the fp16 limit in the device error word (MIXED mode: saturation must not be silent).
int launch_gn_apply(const void* x, int x_is16, const double* stats, const float* gamma,
                    const float* beta, void* y, int y_is16, int fmt, int N, long long HW, int C,
                    int G, float eps, int silu, cudaStream_t s, int range_check = 0, float in_mul = 1.f);
// in_mul: x holds (true value) / in_mul (the 16-bit residual stream of MIXED mode is stored times 2^-6 -> in_mul = 64)
int launch_zero(void* p, size_t bytes, cudaStream_t s);
int launch_f32_to_16(const float* x, void* y, long long n, int fmt, cudaStream_t s);
int launch_16_to_f32(const void* x, float* y, long long n, int fmt, cudaStream_t s);
int launch_softmax_rows(const float* x, void* y, int y_is16, int fmt, long long rows, int cols,
                        cudaStream_t s, long long ld = 0)   /* ld: row pitch of x and y in elements (0 = cols) */;
int launch_head(const float* moments_nhwc8, float* params, float* logvar, float* stdv, float* var,
                int N, int HW, cudaStream_t s);
int launch_sample(const float* mean, const float* logvar, const float* noise, float scale,
                  float* out, long long n, cudaStream_t s);
// logits: [B*T][L] (splits == 0) or the fc split-K partials [splits][B*T][L] + fc_bias (summed in the kernel's load)
int launch_lstm_code(const float* logits, int splits, const float* fc_bias, int B, int T, int L, int layers,
                     const float* w_ih, const float* w_hh, const float* bias,
                     const float* u, float noise_ratio, float temperature, int hard,
                     float* h_out, float* z_out, uint32_t* codes, cudaStream_t s);
void fc_plan(long long K, int L, int* KS, int* splits);
int launch_fc(const float* x, const float* w, float* partial, int N, long long n_stride, long long K, int L, cudaStream_t s);
int launch_hamming(const uint32_t* a, int Na, const uint32_t* b, int Nb, int words, int* out,
                   cudaStream_t s);
int launch_state_consistency(const uint32_t* codes, const int* labels, long long n, int words, int n_states,
                             int* best, int* count, cudaStream_t s);
int launch_perturb(const uint8_t* in, uint8_t* out, int B, int H, int W, const float* noise, float mean,
                   float stdv, const int* occ_xy, int osz, cudaStream_t s);

// ---- tcgen05 conv / GEMM -----------------------------------------------------
struct TcTap { int o[5]; };
struct TcGemmArgs {
  // A operand: 16-bit tensor described for TMA as up to 5 dims (dim0 = K/channels).
  const void* a; int a_rank;
  unsigned long long a_dims[5];
  unsigned long long a_strides[5];   // bytes; strides[0] ignored (contiguous)
  unsigned a_box[5];                 // box[0] = 64
  int dim_x, dim_y, dim_n;           // which A dims the tile x / y / image coordinates feed (-1: none)
  // B operand: [batch][Nrows][K] 16-bit, K contiguous
  const void* b; unsigned long long b_rows, b_k; unsigned long long b_row_stride, b_batch_stride;
  int b_batched;
  int ntaps; TcTap taps[9]; int tap_k[9];  // tap_k: k offset in B for this tap
  const void* a2; int a2_cin;        // optional second A tensor [N,Ho,Wo,a2_cin] read as one extra 1x1 tap (fused shortcut)
  int a2_k0;                         // its k offset in B
  int halo_ok;                       // 3x3 stride-1 conv whose taps are row-major with x offsets -1,0,+1
  int kchunks;                       // 64-wide K chunks per tap
  // tiling of the output
  int BW, BH;                        // tile = BH rows x BW cols (BH*BW == 128)
  int Wo, Ho, Nimg, Cout;
  int block_n;
  // epilogue
  float alpha; const float* bias;
  const void* residual;              // fp32 [.., ldo], or with res16 a 16-bit tensor (format fmt_out) added as res_mul * value
  int res16; float res_mul;
  float* out_f32; void* out_16; int fmt; long long ldo; int relu;
  // fmt is the 16-bit format of A.  With fmt_split set, B and the 16-bit output carry their own formats (MIXED
  // mode: UMMA's instruction descriptor takes a_format and b_format independently); otherwise all three are fmt.
  int fmt_split, fmt_b, fmt_out;
  // out_16 = to16(out16_scale * value) (0 means 1): the 16-bit copies of the residual stream are stored times 2^-6
  // in MIXED mode; a scaled fp16 store is range-checked in the epilogue (error site SITE_XCOPY).
  float out16_scale;
  int sat_check;                     // range-check an (unscaled) fp16 out_16 store in the epilogue (error site SITE_XCOPY)
  // Fused GroupNorm(32)(+SiLU) of the A operand (XF kernels): `a` is the RAW 16-bit input, normalised in shared memory
  // by transform warps with these statistics [Nimg][32][2] / affine.  xf_in_mul: multiplier of the stored input (64 for
  // the scaled residual stream).  Only the HALO BLOCK_N = 256 pair kernel has the variant: launch_tc_gemm returns
  // TC_NOT_FUSABLE (and launches nothing) for any other launch, the caller then runs the stand-alone apply pass.
  const double* xf_stats; const float* xf_gamma; const float* xf_beta; float xf_in_mul; int xf_silu; int xf_check;
  double* gn_stats; int gn_cpg;      // optional fused GroupNorm partial sums [Nimg][Cout/gn_cpg][2] (pre-zeroed)
  // conv_in mode: instead of a 16-bit A tensor, uint8 HWC frames [Nimg][Ho][Wo][3]; the producer warp builds the
  // 3x3x3 patch rows (2u-255, zero padded, duplicated for the hi/lo weight split) straight into the swizzled A tile
  const unsigned char* u8_src;
  // attention scores: out_16 = softmax over the Cout axis of alpha * (A . B^T), fp32 scores never stored (two passes
  // over the n-tiles inside one CTA pair; needs out_16 only: no bias / residual / fp32 output / statistics)
  int softmax_mode;
};
constexpr int TC_NOT_FUSABLE = 1;
int launch_tc_gemm(const TcGemmArgs& a, cudaStream_t s);
int tc_check_device_error(cudaStream_t s);   // sync + read the watchdog flag

}  // namespace sfv
