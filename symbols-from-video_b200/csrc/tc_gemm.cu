// tc_gemm.cu -- tcgen05 / TMEM / TMA implicit-GEMM for sm_100a.
//
// One warp-specialised persistent kernel serves every dense contraction of the
// hot path:
//   D[128 pixels x BLOCK_N] += sum over taps, 64-wide K chunks  A_tap[128 x 64] * B[BLOCK_N x 64]^T
// * A tiles are fetched by TMA from the NHWC 16-bit activation tensor as a
//   (64 channels x BW x BH) box whose origin is shifted by the filter tap, so
//   im2col never exists in memory; out-of-bounds box elements (the zero padding
//   of the convolution, partial tiles, K tails) are zero-filled by the TMA unit.
//   Stride-2 convolutions use a 5-D view (c + parity_x*C, x/2, parity_y, y/2, n)
//   of the same tensor, so they need no element strides and no padded copy
//   (reference: Downsample.forward pads (0,1,0,1), model.py:72-79).
// * B tiles (weights, or K / V^T for attention) are K-major rows, 128B-swizzled.
// * tcgen05.mma (kind::f16, M=128, N=BLOCK_N, K=16) accumulates into TMEM;
//   two accumulator stages let the epilogue of tile i overlap the MMAs of tile i+1.
// * Epilogue warps read TMEM with tcgen05.ld and apply alpha, bias, residual,
//   ReLU, then write fp32 and/or 16-bit NHWC rows (and optional GroupNorm
//   partial sums for the consumer's normalisation).
//
// * Variants selected per launch:
//     HALO      3x3 stride-1 row tiles: one 130-pixel A box serves the three horizontal taps of a filter row
//     NCTA = 2  CTA pairs (tcgen05 cta_group::2) share a 256-pixel tile and split the B tile
//     a2        a second A tensor map appends a fused 1x1 branch (nin_shortcut) as extra K chunks
//     u8_src    conv_in: the producer warp builds the A tile from uint8 frames (exact 2u-255 operand)
//     softmax   attention scores: two passes over the key tiles, P = softmax(alpha S) stored in 16 bit
//   Epilogue outputs leave through swizzled staging tiles and TMA bulk stores; residual tiles arrive by TMA.
//
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer + TMEM
// allocator, warps 2.. = epilogue (4 or 8 warps; TMEM lane quarter = warp_idx % 4).
//
// Environment knobs (experiments / A-B measurements, defaults are the tuned choices): SFV_NCTA, SFV_EPI, SFV_HALO,
// SFV_EPI_SLOTS, SFV_RES_PREFETCH, SFV_TC_DEBUG (role-cycle accounting, tools/tc_debug.py).
#include "common.cuh"
#include <cuda.h>

namespace sfv {

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;            // 64 x 16-bit = 128 B = one swizzle row
constexpr unsigned long long kWatchdogCycles = 4000000000ull;  // ~2 s

// Division by a launch-constant divisor without the ~150-cycle IDIV sequence (round-up method of
// Granlund & Montgomery): q = (umulhi(n, mul) + n) >> shr, exact for 0 <= n < 2^31.
struct FastDiv {
  uint32_t d, mul, shr;
  __device__ __forceinline__ int div(int n) const { return (int)((__umulhi((uint32_t)n, mul) + (uint32_t)n) >> shr); }
  __device__ __forceinline__ void divmod(int n, int& q, int& r) const { q = div(n); r = n - q * (int)d; }
};
static FastDiv make_fastdiv(int d) {
  FastDiv f; f.d = (uint32_t)d; f.shr = 0;
  while ((1u << f.shr) < (uint32_t)d) ++f.shr;
  f.mul = (uint32_t)((((uint64_t)1 << 32) * (((uint64_t)1 << f.shr) - (uint64_t)d)) / (uint64_t)d + 1);
  return f;
}

struct TcParams {
  int dim_x, dim_y, dim_n;
  int ntaps, kchunks;
  int tap_o[9][5];
  int tap_k[9];
  int b_batched;
  int BW_log2, BW, BH;
  int tiles_x, tiles_y, n_tiles_m, n_tiles_n, n_tiles, n_units;
  FastDiv fd_ntn, fd_tx, fd_ty;        // n_tiles_n, tiles_x, tiles_y
  int Wo, Ho, Cout, n_img;
  float alpha;
  const float* bias;
  const float* residual;               // fp32 residual, or (res16) a 16-bit tensor holding res_scale^-1 ... see res16
  int res16; float res_mul;            // res16: the residual is 16-bit (format fmt_out) and is added as res_mul * value
  int res_slot;                        // bytes of one residual / fp32 staging slot: 4096 (fp32 tile) or 2048 (16-bit residual)
  float* out_f32;
  void* out_16;
  int fmt_a, fmt_b, fmt_out;           // 16-bit formats of the A operand, the B operand and out_16
  float out16_scale;                   // out_16 = to16(out16_scale * value); != 1 also range-checks fp16 stores
  // scaled-domain epilogue (16-bit residual stream): alpha, the bias copy and the residual multiplier already carry the
  // stream scale s, so the finished value IS the stored value (no per-element multiply); GroupNorm partial sums are
  // un-scaled when flushed (stat_mul = 1/s for sums, its square for sums of squares; exact, s is a power of two)
  float bias_mul, stat_mul; int sat_check;
  // XF kernels (GroupNorm + SiLU applied to the A operand inside the kernel, see the transform warps): statistics of
  // the raw input [img][32][2] (sum, sum of squares), affine, element count per (image, channel), multiplier of the
  // stored input (64 for the scaled residual stream), SiLU on/off, range check of the fp16 inputs / outputs
  const double* xf_stats; const float* xf_gamma; const float* xf_beta;
  long long xf_hw; int xf_cin; float xf_in_mul; int xf_silu; int xf_check;
  long long ldo;
  int relu;
  double* gn_stats; int gn_cpg; int gn_groups;   // fused GroupNorm partial sums: channels per group, groups
  int epi_mode;
  int epi_fast;                        // lean epilogue of the level-0 launches is applicable (host-side conditions)
  int res_inplace;                     // lean epilogue: the output tile is built in the residual's ring slot and stored from there
  int halo_base_offset;
  int num_stages, res_bufs, h16_slots;   // shared-memory plan of this launch
  int res_prefetch;                    // L2-prefetch the residual one tile ahead of its TMA load
  const unsigned char* u8_src;         // conv_in mode (see TcGemmArgs::u8_src)
  // softmax mode (attention scores): a CTA (pair) owns whole rows -- it walks all n-tiles of its m-tile twice, pass 0
  // keeps the running row max / sum in registers, pass 1 recomputes the scores and stores P = softmax(alpha * S) in
  // 16 bit.  The fp32 scores never leave the SM.
  int softmax_mode, sm_per, n_groups;  // sm_per = 2 * n_tiles_n units per m-tile group
  FastDiv fd_per;
  int a2_kchunks, a2_k0;               // fused 1x1 branch: extra k-chunks read through the second A map
  unsigned long long* dbg;             // optional per-CTA role cycle counters [grid][8]
  int* err;                            // device watchdog flag
};

// unit -> (n_tile, m_tile, tx, ty, img) with multiply-shift divisions
#ifndef SFV_TC_FINE_DEBUG
#define SFV_TC_FINE_DEBUG 0
#endif
constexpr bool kFineDbg = SFV_TC_FINE_DEBUG != 0;   // per-phase cycle counters inside the epilogue chunk loop (58 CS2R per tile)
struct TileCoord { int n_tile, m_tile, tx, ty, img, pass; };
template <int NCTA>
__device__ __forceinline__ TileCoord tile_coord(const TcParams& p, int unit, int rank) {
  TileCoord t;
  int um;
  if (p.softmax_mode) {
    int k;
    p.fd_per.divmod(unit, um, k);          // unit = group * sm_per + pass * n_tiles_n + n_tile
    t.pass = k >= p.n_tiles_n ? 1 : 0;
    t.n_tile = k - t.pass * p.n_tiles_n;
  } else {
    p.fd_ntn.divmod(unit, um, t.n_tile);
    t.pass = 0;
  }
  t.m_tile = um * NCTA + rank;
  int q;
  p.fd_tx.divmod(t.m_tile, q, t.tx);
  p.fd_ty.divmod(q, t.img, t.ty);
  return t;
}

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// Bounded wait: a pipeline bug must surface as an error code, not as a hung GPU.
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, volatile int* abort_flag,
                                          int* err, int code) {
  if (mbar_try_wait(bar, parity)) return true;
  unsigned long long t0 = clock64();
  for (uint32_t spin = 0;; ++spin) {
    if (mbar_try_wait(bar, parity)) return true;
    if ((spin & 255u) == 255u) {
      if (*abort_flag) return false;
      if (clock64() - t0 > kWatchdogCycles) {
        *abort_flag = 1;
        atomicCAS(err, 0, code);
        return false;
      }
    }
  }
}
// One lane of a converged warp (the compiler keeps warp-uniform operands in uniform registers
// when the tcgen05 / TMA instructions are issued this way instead of from a divergent lane-0 branch).
__device__ __forceinline__ uint32_t elect_one_sync() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, %1;\n\t"
      "@px mov.s32 %0, 1;\n\t}"
      : "+r"(pred)
      : "r"(0xFFFFFFFFu));
  return pred;
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// L2 prefetch of a tile (no shared memory, no completion tracking)
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* map, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
               ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// shared -> global tile store (clips out-of-bounds elements); completion tracked by bulk groups
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
      ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
// order generic-proxy shared-memory writes before async-proxy (TMA) accesses
// explicit shared-space 128-bit accesses on 32-bit addresses (generic LD/ST cost 64-bit address arithmetic per access)
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void sts128u(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {       // 2^x, 2 ulp; x <= 0 here
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T ; kind::f16 covers fp16 and bf16 operands.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrive once all previously issued tcgen05.mma of this thread retire.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- cta_group::2 (CTA pair) variants -------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Both CTAs of the pair issue their own loads into their own shared memory; the transaction
// bytes are reported to the LEADER CTA's barrier (peer bit of the cluster address cleared).
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ void tma_load_5d_2cta(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                                 int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_2cta(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                                 int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
// M = 256 over the CTA pair: each CTA supplies its 128 rows of A and half of B's rows, and
// receives its 128 accumulator rows in its own TMEM.  Issued by the leader CTA only.
__device__ __forceinline__ void umma_f16_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                              uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this shared-memory offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_2cta(uint32_t bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"(mask)
      : "memory");
}
// Remote arrive on the barrier at the same offset in CTA `cta`.  Default (.release.cta) semantics as in CUTLASS'
// ClusterBarrier::arrive: a .cluster-scope release costs a full membar (+L1 invalidate, ~1600 cycles per tile measured);
// the TMEM hand-over is ordered by tcgen05.wait::ld + tcgen05.fence::before_thread_sync.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_local, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(bar_local), "r"(cta)
      : "memory");
}

// UMMA shared-memory descriptor, K-major, 128-byte swizzle, 8-row atoms 1024 B apart
// (bit layout: cute::UMMA::SmemDescriptor; version=1 for sm_100).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t base_offset = 0) {
  uint64_t d = 0;
  d |= (uint64_t)(base_offset & 7u) << 49;       // swizzle phase of a start address that is not 1024 B aligned
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);        // start address
  d |= (uint64_t)1 << 16;                        // leading byte offset (unused for SW128 K-major)
  d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset: 8 rows * 128 B
  d |= (uint64_t)1 << 46;                        // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
  return d;
}
// UMMA instruction descriptor (cute::UMMA::InstrDescriptor): fp32 accumulate,
// A/B format fmt (0 = f16, 1 = bf16), both K-major, M = 128, N = n.
__host__ __device__ constexpr uint32_t make_idesc(int fmt_a, int fmt_b, int n, int m) {
  return (1u << 4) | ((uint32_t)fmt_a << 7) | ((uint32_t)fmt_b << 10) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}


// ---- fused GroupNorm statistics (epilogue) -----------------------------------
// Each lane holds NV partial sums for its pixel row (sum and sum of squares of NG
// channel groups).  Recursive halving: at every step a lane keeps one half of its
// values and adds the partner's copy of that half, so NV values cost NV-1 (+log)
// shuffles instead of 5*NV; value `idx` ends, fully reduced over the 32 rows, in
// the lanes whose upper bits spell idx.
template <int NV>
__device__ __forceinline__ float halving_reduce(float (&v)[NV], int lane, int& idx) {
  int base = 0;
  int off = 16;
#pragma unroll
  for (int cnt = NV; cnt > 1; cnt >>= 1) {
    const bool up = (lane & off) != 0;
    const int half = cnt >> 1;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = up ? v[i] : v[i + half];
      const float keep = up ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
    base += up ? half : 0;
    off >>= 1;
  }
  float r = v[0];
  for (; off > 0; off >>= 1) r += __shfl_xor_sync(0xffffffffu, r, off);
  idx = base;
  return r;
}

// f: this lane's 32 finished output values (columns c0..c0+31 of its pixel row).
template <int CPG>
__device__ __forceinline__ void epi_stats_chunk(const float* f, bool valid, int lane, float* acc_w, int group0) {
  constexpr int NG = 32 / CPG;
  constexpr int NV = 2 * NG;
  float v[NV];
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int j = 0; j < CPG; ++j) {
      const float t = f[g * CPG + j];
      s += t; q = fmaf(t, t, q);
    }
    v[2 * g] = valid ? s : 0.f; v[2 * g + 1] = valid ? q : 0.f;      // clipped rows of a partial tile contribute nothing
  }
  int idx;
  const float r = halving_reduce<NV>(v, lane, idx);
  if ((lane & (32 / NV - 1)) == 0) acc_w[group0 * 2 + idx] += r;   // one owner lane per (group, moment): no race
}

// Epilogue staging (per epilogue warp): a ring of kResBufs fp32 tiles (32 rows x 128 B, 128B-swizzled) that
// first receive the residual by TMA and then hold the fp32 output for the TMA store, and two 16-bit
// output tiles (32 rows x 64 B, 64B-swizzled).
// The epilogue of a 128 x 128 tile costs about as many cycles as its 72 MMAs when four single warps do it
// (latency-bound instruction stream), so BLOCK_N = 128 kernels run TWO warps per TMEM lane quarter, each
// taking half of the columns (kSets = 2, 8 epilogue warps, 2-slot rings); wider tiles keep 4 warps.
constexpr int kAuxBytes = 512 /*barriers*/ + 2048 /*GN partials*/ + 4096 /*bias copies*/;
// residual ring depth per epilogue warp: the load of chunk g + depth - 1 is issued while chunk g is processed.  A
// 16-bit residual tile is 2 KB, so four slots (three chunks of lookahead, ~7000 cycles at level 0) cost what two
// fp32 slots did; with one chunk of lookahead the epilogue sat on the TMA latency and the MMA warp on tmem_empty.
constexpr int kMaxResBufs = 4;          // 20 pipeline barriers + 8 warps x 4 + TMEM pointer fit the 512-byte barrier block

// HALO: for 3x3 stride-1 convolutions tiled as 128-pixel row segments, one TMA box of 130 pixels feeds the
// three horizontal taps of a filter row; the MMAs read it at start addresses shifted by one pixel
// (128 B) per tap.  Cuts the A operand's TMA->smem traffic 3x on the layers where the shared-memory
// port, not the tensor pipe, bounds the main loop (Cout = 128).
// Per-lane partial sums only (no cross-lane traffic): acc[2g], acc[2g+1] += sum / sum of squares of group g of this
// lane's row.  Used by the 8-warp epilogue, which reduces across lanes once per image instead of once per chunk.
template <int CPG>
__device__ __forceinline__ void epi_stats_lane(const float* f, bool valid, float* acc) {
  if (!valid) return;                     // only lanes of a partial tile's clipped rows diverge here
#pragma unroll
  for (int g = 0; g < 32 / CPG; ++g) {
    float s = acc[2 * g], q = acc[2 * g + 1];
#pragma unroll
    for (int j = 0; j < CPG; ++j) {
      const float t = f[g * CPG + j];
      s += t; q = fmaf(t, t, q);
    }
    acc[2 * g] = s; acc[2 * g + 1] = q;
  }
}

// conv_in mode, a lane's four adjacent pixels x0 .. x0+3 of row y.
// conv_in_values: the 3 x 18 bytes around them (6 aligned words per filter row; byte 0 of the first word is one byte
// before pixel x0-1) -> 2u-255 as floats, 0 outside the frame.  INTERIOR skips the zero-padding selects.
template <bool INTERIOR>
__device__ __forceinline__ void conv_in_values(const uint32_t (&w)[3][6], int x0, int y, int W, int H, float (&v)[3][18]) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int yy = y + r - 1;
    const bool row_ok = INTERIOR || (yy >= 0 && yy < H);
#pragma unroll
    for (int i = 0; i < 18; ++i) {
      const int bi = i + 1;
      const float t = fmaf((float)((w[r][bi >> 2] >> (8 * (bi & 3))) & 0xffu), 2.f, -255.f);
      if constexpr (INTERIOR) v[r][i] = t;
      else { const int xx = x0 - 1 + i / 3; v[r][i] = (row_ok && xx >= 0 && xx < W) ? t : 0.f; }
    }
  }
}
// conv_in_store_rows: rows m0 .. m0+3 of the 128B-swizzled A tile; k = (dy*3+dx)*3+c for k < 27, repeated at 27..53.
template <int FMT>
__device__ __forceinline__ void conv_in_store_rows(const float (&v)[3][18], uint32_t tile_s, int m0) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int m = m0 + j;
    const uint32_t row_s = tile_s + m * 128;
    uint32_t wd[27];
#pragma unroll
    for (int i = 0; i < 27; ++i) {                            // 54 values = 27 packed pairs
      const int e0 = 2 * i, e1 = 2 * i + 1;
      const int k0 = e0 < 27 ? e0 : e0 - 27, k1 = e1 < 27 ? e1 : e1 - 27;
      wd[i] = pack2_16(v[k0 / 9][3 * j + k0 % 9], v[k1 / 9][3 * j + k1 % 9], FMT);
    }
#pragma unroll
    for (int c = 0; c < 8; ++c)
      sts128u(row_s + ((c ^ (m & 7)) << 4), 4 * c < 27 ? wd[4 * c] : 0u, 4 * c + 1 < 27 ? wd[4 * c + 1] : 0u,
              4 * c + 2 < 27 ? wd[4 * c + 2] : 0u, 4 * c + 3 < 27 ? wd[4 * c + 3] : 0u);
  }
}

constexpr int kXfWarps = 8;             // transform warps of an XF kernel (two per scheduler: the transform is MUFU / latency bound)
constexpr int kXfRowStep = 4 * kXfWarps; // rows of the A box covered by one pass of all transform threads (8 units per row)
template <int BLOCK_N, int NCTA, bool HALO = false, bool XF = false>
struct Cfg {
  static constexpr int kGroup = HALO ? 3 : 1;                           // filter taps per pipeline stage
  static constexpr int kATxBytes = HALO ? 130 * 128 : kBlockM * kBlockK * 2;
  static constexpr int kABytes = HALO ? 17 * 1024 : kBlockM * kBlockK * 2;   // keeps the B tiles 1024 B aligned
  static constexpr int kBBytes = (BLOCK_N / NCTA) * kBlockK * 2;      // a CTA pair splits B's rows
  static constexpr int kStageBytes = kABytes + kGroup * kBBytes;
  static constexpr int kTxBytes = kATxBytes + kGroup * kBBytes;
  // epilogue warps per TMEM lane quarter: two (each half of the columns) for BLOCK_N = 128 and for the non-HALO
  // BLOCK_N = 256 kernels (attention GEMMs, 64x64 convs: short K, epilogue-bound with four warps)
  static constexpr int kSets = (BLOCK_N == 128 || (BLOCK_N == 256 && !HALO)) ? 2 : 1;
  static constexpr int kEpiWarps = 4 * kSets;
  static constexpr int kThreads = 64 + 32 * kEpiWarps + (XF ? 32 * kXfWarps : 0);
  static constexpr int kXfBytes = XF ? 4096 : 0;       // per-image (scale, shift) table of the fused GroupNorm, [Cin <= 512] float2
  // The staging rings are sized per launch (a conv that writes only 16-bit outputs needs no fp32 ring), and
  // whatever shared memory is left becomes pipeline stages: res_bufs fp32 slots (0, 2 or 3) and h16_slots
  // (0 or 2) per epilogue warp, see smem_plan().
  static constexpr int kMaxStages = 8;
  static constexpr int kSmemLimit = 232448;
  static __host__ __device__ constexpr int epi_bytes(int res_bufs, int h16_slots, int res_slot = 4096) {
    return BLOCK_N >= 32 ? kEpiWarps * (res_bufs * res_slot + h16_slots * 2048) : 0;
  }
  static __host__ constexpr int stages_for(int res_bufs, int h16_slots, int res_slot = 4096) {
    const int n = (kSmemLimit - 1024 - kAuxBytes - kXfBytes - epi_bytes(res_bufs, h16_slots, res_slot)) / kStageBytes;
    return n > kMaxStages ? kMaxStages : n;
  }
  static constexpr int kTmemCols = (2 * BLOCK_N) < 32 ? 32 : (2 * BLOCK_N);
  static constexpr int kChunk = BLOCK_N < 32 ? 16 : 32;
  static __host__ constexpr int smem_bytes(int stages, int res_bufs, int h16_slots, int res_slot = 4096) {
    return stages * kStageBytes + epi_bytes(res_bufs, h16_slots, res_slot) + kAuxBytes + kXfBytes + 1024 /*align slack*/;
  }
};

__device__ __forceinline__ float xf_silu_tanh(float t) {      // t * sigmoid(t) = h + h tanh(h), h = t/2 (as norm.cu's silu_tanh)
  const float h = 0.5f * t;
  float th;
  asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(h));
  return fmaf(h, th, h);
}

// One stage of a transform thread: its 16-byte units (8 channels c_log*8.. of rows r0, r0 + 32, r0 + 64, r0 + 96, and for
// the 16 threads with r0 < 2 also r0 + 128) are all loaded first, then normalised with every value independent, then
// stored.  Straight-line code of ~200 instructions: a per-unit load / compute / store loop serialised on the LDS and
// MUFU latencies (ncu: the first use of the loaded word was the top stall), the nine-unit unrolled form of the
// four-warp version was instruction-fetch bound.  Columns outside the image (the convolution's zero padding, filled
// by TMA) are computed but never written back.
template <int FMT>
__device__ __forceinline__ void xf_stage(uint32_t sa, int r0, int c_log, int x_first, int Wo, const float (&sc)[8],
                                         const float (&sf)[8], bool silu, bool chk, __half2& mx_out) {
  constexpr int kMain = 128 / kXfRowStep;              // units every thread owns (rows < 128)
  uint32_t w[kMain + 1][4];
  const bool tail = r0 + kMain * kXfRowStep < 130;     // rows 128, 129
#pragma unroll
  for (int u = 0; u <= kMain; ++u) {
    const int r = r0 + kXfRowStep * u;
    if (u < kMain || tail) {
      const float4 q = lds128(sa + r * 128 + ((c_log ^ (r & 7)) << 4));
      w[u][0] = __float_as_uint(q.x); w[u][1] = __float_as_uint(q.y); w[u][2] = __float_as_uint(q.z); w[u][3] = __float_as_uint(q.w);
    }
  }
#pragma unroll
  for (int u = 0; u <= kMain; ++u) {
    if (u < kMain || tail) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float a, b;
        unpack2_16(w[u][i], FMT, a, b);
        a = fmaf(a, sc[2 * i], sf[2 * i]); b = fmaf(b, sc[2 * i + 1], sf[2 * i + 1]);
        if (silu) { a = xf_silu_tanh(a); b = xf_silu_tanh(b); }
        w[u][i] = pack2_16(a, b, FMT);
      }
    }
  }
#pragma unroll
  for (int u = 0; u <= kMain; ++u) {
    const int r = r0 + kXfRowStep * u;
    const int px = x_first + r;
    if ((u < kMain || tail) && px >= 0 && px < Wo) {
      if (FMT == FMT_F16 && chk) {
        mx_out = __hmax2(__hmax2(mx_out, __habs2(*reinterpret_cast<const __half2*>(&w[u][0]))),
                         __habs2(*reinterpret_cast<const __half2*>(&w[u][1])));
        mx_out = __hmax2(__hmax2(mx_out, __habs2(*reinterpret_cast<const __half2*>(&w[u][2]))),
                         __habs2(*reinterpret_cast<const __half2*>(&w[u][3])));
      }
      sts128u(sa + r * 128 + ((c_log ^ (r & 7)) << 4), w[u][0], w[u][1], w[u][2], w[u][3]);
    }
  }
}

template <int BLOCK_N, int NCTA, bool HALO, bool XF = false>
__global__ void __launch_bounds__(Cfg<BLOCK_N, NCTA, HALO, XF>::kThreads, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmO32,
               const __grid_constant__ CUtensorMap tmO16, const TcParams p) {
  using C = Cfg<BLOCK_N, NCTA, HALO, XF>;
  // NCTA == 2: the kernel runs as clusters of two CTAs (one SM pair) that share one
  // 256-pixel x BLOCK_N tile; rank 0 issues the MMAs for both (tcgen05 cta_group::2).
  const uint32_t rank = NCTA == 2 ? cluster_ctarank() : 0u;
  const int unit0 = blockIdx.x / NCTA, unit_step = gridDim.x / NCTA;
  // it-th work unit of this CTA (pair), or -1: round-robin tiles, or -- softmax mode -- whole m-tile groups
  auto unit_of = [&](int it) -> int {
    if (!p.softmax_mode) { const int u = unit0 + it * unit_step; return u < p.n_units ? u : -1; }
    int g, k;
    p.fd_per.divmod(it, g, k);
    const int grp = unit0 + g * unit_step;
    return grp < p.n_groups ? grp * p.sm_per + k : -1;
  };
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B atoms must start on 1024-byte boundaries
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int num_stages = p.num_stages;
  const int kResBufs = p.res_bufs;                                        // fp32 staging slots per epilogue warp (runtime)
  uint8_t* epi_f32 = smem + num_stages * C::kStageBytes;                  // 1024-aligned (stage sizes are)
  const int kResSlot = p.res_slot;                                        // 4096, or 2048 when the residual is 16-bit
  uint8_t* epi_h16 = epi_f32 + C::kEpiWarps * kResBufs * kResSlot;
  uint64_t* bars = (uint64_t*)(smem + num_stages * C::kStageBytes + C::epi_bytes(kResBufs, p.h16_slots, kResSlot));
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + C::kMaxStages;
  uint64_t* tfull_bar = bars + 2 * C::kMaxStages;
  uint64_t* tempty_bar = bars + 2 * C::kMaxStages + 2;
  uint64_t* res_bar = bars + 2 * C::kMaxStages + 4;                        // [epilogue warps][kMaxResBufs]
  // XF: afull_bar = this CTA's A box has landed (local), xf_bar = the A tiles of the stage are transformed (leader's)
  uint64_t* afull_bar = res_bar + C::kEpiWarps * kMaxResBufs;
  uint64_t* xf_bar = afull_bar + C::kMaxStages;
  uint32_t* tmem_ptr = (uint32_t*)(XF ? xf_bar + C::kMaxStages : afull_bar);
  volatile int* abort_flag = (volatile int*)(tmem_ptr + 1);
  float* gn_acc = (float*)((uint8_t*)bars + 512);        // [epilogue warps][512 / warps floats]: (sum, sumsq) per group
  float* bias_all = gn_acc + 512;                        // [epilogue warps][1024 / warps floats]
  float2* xf_tab = (float2*)(bias_all + 1024);           // XF: [Cin] (scale, shift) of the current image

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < num_stages; ++i) {
      mbar_init(smem_u32(&full_bar[i]), 1);
      mbar_init(smem_u32(&empty_bar[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&tfull_bar[i]), 1);
      mbar_init(smem_u32(&tempty_bar[i]), C::kEpiWarps * NCTA);
    }
    for (int i = 0; i < C::kEpiWarps * kMaxResBufs; ++i) mbar_init(smem_u32(&res_bar[i]), 1);
    if constexpr (XF) {
      for (int i = 0; i < num_stages; ++i) {
        mbar_init(smem_u32(&afull_bar[i]), 1);
        mbar_init(smem_u32(&xf_bar[i]), kXfWarps * NCTA);
      }
    }
    *abort_flag = 0;
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 512; i += C::kThreads) gn_acc[i] = 0.f;
  if (warp == 1) {
    if constexpr (NCTA == 2) tmem_alloc_2cta(smem_u32(tmem_ptr), C::kTmemCols);
    else tmem_alloc(smem_u32(tmem_ptr), C::kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (NCTA == 2) cluster_sync_all();     // peer barriers are initialised before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int k_iters = (p.ntaps / C::kGroup) * p.kchunks + p.a2_kchunks;

  if (warp == 0) {
    // ===================== TMA producer (converged warp, one elected lane issues) =====================
    {
      int stage = 0; uint32_t phase = 0;
      bool ok = true;
      unsigned long long t_wait = 0, t_start = clock64();
      if constexpr (NCTA == 1 && !HALO && BLOCK_N == 128) {
        if (p.u8_src) {
          // ---- conv_in: A rows are built here from the uint8 frame, one k-chunk per tile ----
          // Row m of the tile = pixel (y, x0+m); k = (dy*3+dx)*3 + c for k < 27 holds 2u-255 of pixel
          // (y+dy-1, x+dx-1) (0 outside the frame = the reference's zero padding of the normalised image),
          // k = 27..53 repeats them (the weights are split hi + lo), k = 54..63 are zero.  A lane owns 4 adjacent
          // pixels: per filter row it needs 18 consecutive bytes = 6 aligned words, loaded in bounds (indices are
          // clamped; whatever lies outside the frame is masked) one tile ahead: the words are turned into values
          // first, then the same registers receive the next tile's loads while the rows are packed and stored.
          const int W3w = p.Wo * 3 / 4;                       // words per frame row
          const uint32_t* src32 = reinterpret_cast<const uint32_t*>(p.u8_src);
          uint32_t w[3][6];
          auto fetch = [&](const TileCoord& tc) {
            const int wi0 = 3 * (tc.tx * 32 + lane) - 1;      // word holding the byte before pixel x0-1 (x0 = tx*128 + 4*lane)
#pragma unroll
            for (int r = 0; r < 3; ++r) {
              const int yy = min(max(tc.ty + r - 1, 0), p.Ho - 1);
              const uint32_t* rowp = src32 + ((long long)tc.img * p.Ho + yy) * W3w;
#pragma unroll
              for (int i = 0; i < 6; ++i) w[r][i] = __ldg(rowp + min(max(wi0 + i, 0), W3w - 1));
            }
          };
          TileCoord tc = tile_coord<NCTA>(p, unit0 < p.n_units ? unit0 : 0, 0);
          fetch(tc);
#pragma unroll 1
          for (int unit = unit0; unit < p.n_units && ok; unit += unit_step) {
            const int x0 = tc.tx * 128 + 4 * lane, y = tc.ty, n_tile = tc.n_tile;
            float v[3][18];
            const bool interior = __all_sync(0xffffffffu, y >= 1 && y + 1 < p.Ho && x0 >= 1 && x0 + 4 < p.Wo);
            if (interior) conv_in_values<true>(w, x0, y, p.Wo, p.Ho, v);
            else conv_in_values<false>(w, x0, y, p.Wo, p.Ho, v);
            if (unit + unit_step < p.n_units) { tc = tile_coord<NCTA>(p, unit + unit_step, 0); fetch(tc); }
            const unsigned long long tw = p.dbg ? clock64() : 0;
            ok = mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1, abort_flag, p.err, 1);
            if (p.dbg) t_wait += clock64() - tw;
            if (!ok) break;
            const uint32_t tile_s = smem_u32(smem + stage * C::kStageBytes);
            if (p.fmt_a == FMT_BF16) conv_in_store_rows<FMT_BF16>(v, tile_s, 4 * lane);
            else conv_in_store_rows<FMT_F16>(v, tile_s, 4 * lane);
            fence_proxy_async();
            __syncwarp();
            if (elect_one_sync()) {
              const uint32_t fb = smem_u32(&full_bar[stage]);
              mbar_arrive_expect_tx(fb, C::kBBytes);
              tma_load_3d(tile_s + C::kABytes, &tmB, fb, 0, n_tile * BLOCK_N, 0);
            }
            __syncwarp();
            if (++stage == num_stages) { stage = 0; phase ^= 1; }
          }
          if (p.dbg && lane == 0) { p.dbg[blockIdx.x * 16 + 0] = clock64() - t_start; p.dbg[blockIdx.x * 16 + 1] = t_wait; }
          ok = false;                                          // skip the TMA producer loop below
        }
      }
      for (int it = 0; ok; ++it) {
        const int unit = unit_of(it);
        if (unit < 0) break;
        const TileCoord tc = tile_coord<NCTA>(p, unit, (int)rank);   // m_tile may be a phantom tile (>= n_tiles_m): all OOB -> zeros
        const int n_tile = tc.n_tile, m_tile = tc.m_tile, tx = tc.tx, ty = tc.ty, img = tc.img;
        int base[5] = {0, 0, 0, 0, 0};
        if (p.dim_x >= 0) base[p.dim_x] += tx * p.BW;
        if (p.dim_y >= 0) base[p.dim_y] += ty * p.BH;
        if (p.dim_n >= 0) base[p.dim_n] += img;
        if (p.dim_n < 0 && p.dim_y < 0 && m_tile >= p.n_tiles_m) base[p.dim_x] = 0x3fffffff;   // phantom row block
        for (int tap = 0; tap < p.ntaps && ok; tap += C::kGroup) {
          const int c1 = base[1] + p.tap_o[tap][1], c2 = base[2] + p.tap_o[tap][2];
          const int c3 = base[3] + p.tap_o[tap][3], c4 = base[4] + p.tap_o[tap][4];
          for (int kc = 0; kc < p.kchunks; ++kc) {
            const unsigned long long tw = p.dbg ? clock64() : 0;
            ok = mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1, abort_flag, p.err, 1);
            if (p.dbg) t_wait += clock64() - tw;
            if (!ok) break;
            const uint32_t sa = smem_u32(smem + stage * C::kStageBytes);
            const uint32_t fb = smem_u32(&full_bar[stage]);
            if (elect_one_sync()) {
              // HALO: tap is the left tap of a filter row; its box is 130 pixels wide and serves taps tap..tap+2
              if constexpr (NCTA == 2 && XF) {
                // XF: this CTA's A box completes on its OWN barrier (its transform warps wait there); the weight tiles
                // of both CTAs complete on the leader's full barrier as usual
                const uint32_t ab = smem_u32(&afull_bar[stage]);
                mbar_arrive_expect_tx(ab, C::kATxBytes);
                tma_load_5d(sa, &tmA, ab, p.tap_o[tap][0] + kc * kBlockK, c1, c2, c3, c4);
                if (rank == 0) mbar_arrive_expect_tx(fb, 2 * C::kGroup * C::kBBytes);
#pragma unroll
                for (int g = 0; g < C::kGroup; ++g)
                  tma_load_3d_2cta(sa + C::kABytes + g * C::kBBytes, &tmB, fb, p.tap_k[tap + g] + kc * kBlockK,
                                   n_tile * BLOCK_N + (int)rank * (BLOCK_N / 2), p.b_batched ? img : 0);
              } else if constexpr (NCTA == 2) {
                // the leader's barrier collects the bytes of both CTAs' loads
                if (rank == 0) mbar_arrive_expect_tx(fb, 2 * C::kTxBytes);
                tma_load_5d_2cta(sa, &tmA, fb, p.tap_o[tap][0] + kc * kBlockK, c1, c2, c3, c4);
#pragma unroll
                for (int g = 0; g < C::kGroup; ++g)
                  tma_load_3d_2cta(sa + C::kABytes + g * C::kBBytes, &tmB, fb, p.tap_k[tap + g] + kc * kBlockK,
                                   n_tile * BLOCK_N + (int)rank * (BLOCK_N / 2), p.b_batched ? img : 0);
              } else {
                mbar_arrive_expect_tx(fb, C::kTxBytes);
                tma_load_5d(sa, &tmA, fb, p.tap_o[tap][0] + kc * kBlockK, c1, c2, c3, c4);
#pragma unroll
                for (int g = 0; g < C::kGroup; ++g)
                  tma_load_3d(sa + C::kABytes + g * C::kBBytes, &tmB, fb, p.tap_k[tap + g] + kc * kBlockK,
                              n_tile * BLOCK_N, p.b_batched ? img : 0);
              }
            }
            __syncwarp();
            if (++stage == num_stages) { stage = 0; phase ^= 1; }
          }
        }
        {
          // fused 1x1 branch (nin_shortcut): same pixels, no tap shift, channels of the second tensor
          // (one plain 128-pixel A box + one B tile per stage, also inside a HALO kernel)
          constexpr int kTx1 = kBlockM * kBlockK * 2 + C::kBBytes;
          for (int kc = 0; kc < p.a2_kchunks && ok; ++kc) {
            ok = mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1, abort_flag, p.err, 1);
            if (!ok) break;
            const uint32_t sa = smem_u32(smem + stage * C::kStageBytes);
            const uint32_t fb = smem_u32(&full_bar[stage]);
            if (elect_one_sync()) {
              if constexpr (NCTA == 2 && XF) {
                const uint32_t ab = smem_u32(&afull_bar[stage]);
                mbar_arrive_expect_tx(ab, kBlockM * kBlockK * 2);
                tma_load_5d(sa, &tmA2, ab, kc * kBlockK, base[1], base[2], base[3], base[4]);
                if (rank == 0) mbar_arrive_expect_tx(fb, 2 * C::kBBytes);
                tma_load_3d_2cta(sa + C::kABytes, &tmB, fb, p.a2_k0 + kc * kBlockK,
                                 n_tile * BLOCK_N + (int)rank * (BLOCK_N / 2), 0);
              } else if constexpr (NCTA == 2) {
                if (rank == 0) mbar_arrive_expect_tx(fb, 2 * kTx1);
                tma_load_5d_2cta(sa, &tmA2, fb, kc * kBlockK, base[1], base[2], base[3], base[4]);
                tma_load_3d_2cta(sa + C::kABytes, &tmB, fb, p.a2_k0 + kc * kBlockK,
                                 n_tile * BLOCK_N + (int)rank * (BLOCK_N / 2), 0);
              } else {
                mbar_arrive_expect_tx(fb, kTx1);
                tma_load_5d(sa, &tmA2, fb, kc * kBlockK, base[1], base[2], base[3], base[4]);
                tma_load_3d(sa + C::kABytes, &tmB, fb, p.a2_k0 + kc * kBlockK, n_tile * BLOCK_N, 0);
              }
            }
            __syncwarp();
            if (++stage == num_stages) { stage = 0; phase ^= 1; }
          }
        }
      }
      if (p.dbg && lane == 0) { p.dbg[blockIdx.x * 16 + 0] = clock64() - t_start; p.dbg[blockIdx.x * 16 + 1] = t_wait; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (converged warp, one elected lane issues; leader CTA only in pair mode) =====
    if (rank == 0) {
      const uint32_t idesc = make_idesc(p.fmt_a, p.fmt_b, BLOCK_N, kBlockM * NCTA);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      bool ok = true;
      unsigned long long t_full = 0, t_tempty = 0, t_start = clock64();
      for (int it = 0; ok; ++it) {
        if (unit_of(it) < 0) break;
        const unsigned long long tw0 = p.dbg ? clock64() : 0;
        ok = mbar_wait(smem_u32(&tempty_bar[acc]), acc_phase ^ 1, abort_flag, p.err, 2);
        if (p.dbg) t_tempty += clock64() - tw0;
        if (!ok) break;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
        for (int it = 0; it < k_iters; ++it) {
          const unsigned long long tw = p.dbg ? clock64() : 0;
          ok = mbar_wait(smem_u32(&full_bar[stage]), phase, abort_flag, p.err, 3);
          if (p.dbg) t_full += clock64() - tw;
          if (!ok) break;
          if constexpr (XF) {      // the A tiles of both CTAs have been normalised in place
            ok = mbar_wait(smem_u32(&xf_bar[stage]), phase, abort_flag, p.err, 7);
            if (!ok) break;
          }
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * C::kStageBytes);
          const int main_iters = k_iters - p.a2_kchunks;
          if (elect_one_sync()) {
#pragma unroll
            for (int g = 0; g < C::kGroup; ++g) {
              if (g > 0 && it >= main_iters) break;      // fused 1x1 branch stages carry a single tap
              // HALO: tap g of the filter row reads the same box one pixel (128 B) further right.  The 128B swizzle
              // is applied to absolute smem address bits by both TMA and UMMA, so a start address that is
              // 128-byte (not 1024-byte) aligned needs no descriptor base offset (verified on B200).
              const uint64_t da = make_smem_desc(sa + g * 128, p.halo_base_offset ? (uint32_t)g : 0u);
              const uint64_t db = make_smem_desc(sa + C::kABytes + g * C::kBBytes);
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k) {
                // advance 16 elements (32 B) along K inside the swizzle atom: +2 in the >>4 address field
                const uint32_t accum = (it > 0 || g > 0 || k > 0) ? 1u : 0u;
                if constexpr (NCTA == 2)
                  umma_f16_2cta(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, accum);
                else
                  umma_f16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, accum);
              }
            }
            if constexpr (NCTA == 2) {
              umma_commit_2cta(smem_u32(&empty_bar[stage]), 3);             // frees the stage in both CTAs
              if (it == k_iters - 1) umma_commit_2cta(smem_u32(&tfull_bar[acc]), 3);
            } else {
              umma_commit(smem_u32(&empty_bar[stage]));           // frees the smem stage when MMAs retire
              if (it == k_iters - 1) umma_commit(smem_u32(&tfull_bar[acc]));  // accumulator complete
            }
          }
          __syncwarp();
          if (++stage == num_stages) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if (p.dbg && lane == 0) { p.dbg[blockIdx.x * 16 + 2] = clock64() - t_start; p.dbg[blockIdx.x * 16 + 3] = t_full; p.dbg[blockIdx.x * 16 + 4] = t_tempty; }
    }
  } else if (XF && warp >= 2 + C::kEpiWarps) {
    // ===================== transform warps (XF kernels): GroupNorm (+SiLU) of the A operand, in place ==============
    // The A tensor of this launch is the RAW 16-bit input (conv1: the residual stream, stored x 2^-6; conv2: conv1's
    // output).  Each landed A box is normalised where it sits -- y = silu(v * scale[c] + shift[c]) with the same fp32
    // arithmetic, in the same order, as the stand-alone apply pass (norm.cu), so the operand the tensor pipe sees is
    // bit-identical -- and the separate HBM pass (read + write of the whole activation) disappears.  Thread t owns the
    // 16-byte unit (8 channels) c_log = t & 7 of rows (t >> 3) + 16 j: its 8 (scale, shift) pairs are loaded once per
    // stage, and the 32 lanes of a warp touch 4 whole 128-byte rows (conflict free under the 128B swizzle).  Rows of
    // the box that lie outside the image are the convolution's zero padding (TMA zero fill) and stay untouched.
    if constexpr (XF) {
      const int t = (warp - 2 - C::kEpiWarps) * 32 + lane;
      const int c_log = t & 7, r0 = t >> 3;
      const int cpg = p.xf_cin >> 5;
      int stage = 0; uint32_t phase = 0;
      int cur_img = -1;
      bool ok = true;
      const bool chk = p.xf_check && p.fmt_a == FMT_F16;
      __half2 mx_out = __float2half2_rn(0.f);      // range check of the normalised operand (its raw input was checked when written)
      unsigned long long xt_wait = 0, xt_work = 0, xt_start = clock64();
      auto hand_over = [&]() {            // the stage's A tile is final: tell the leader's MMA warp
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (NCTA == 2 && rank != 0) mbar_arrive_cluster(smem_u32(&xf_bar[stage]), 0);
          else mbar_arrive(smem_u32(&xf_bar[stage]));
        }
        if (++stage == num_stages) { stage = 0; phase ^= 1; }
      };
      for (int it = 0; ok; ++it) {
        const int unit = unit_of(it);
        if (unit < 0) break;
        const TileCoord tc = tile_coord<NCTA>(p, unit, (int)rank);
        const bool tile_ok = tc.m_tile < p.n_tiles_m;
        if (tile_ok && tc.img != cur_img) {
          // (scale, shift) of every input channel for this image, from the producer's fp64 sums
          asm volatile("bar.sync 8, %0;" ::"n"(32 * kXfWarps) : "memory");      // nobody still reads the previous image's table
          const double cnt = (double)p.xf_hw * cpg;
          const double* st = p.xf_stats + (long long)tc.img * 64;
          for (int c = t; c < p.xf_cin; c += 32 * kXfWarps) {
            const int g = c / cpg;
            const double m = st[g * 2] / cnt;
            double var = st[g * 2 + 1] / cnt - m * m;
            if (var < 0) var = 0;
            const float rstd = (float)(1.0 / sqrt(var + 1e-6));
            const float gr = __ldg(p.xf_gamma + c) * rstd;
            xf_tab[c] = make_float2(gr * p.xf_in_mul, __ldg(p.xf_beta + c) - (float)m * gr);
          }
          asm volatile("bar.sync 8, %0;" ::"n"(32 * kXfWarps) : "memory");
          cur_img = tc.img;
        }
        const int x_first = tc.tx * 128 - 1;                   // image column of box row 0 (left tap of a filter row)
        for (int tap = 0; tap < p.ntaps && ok; tap += C::kGroup) {
          const int yy = tc.ty + p.tap_o[tap][2];
          const bool row_in = tile_ok && yy >= 0 && yy < p.Ho;
          for (int kc = 0; kc < p.kchunks; ++kc) {
            const unsigned long long tq0 = p.dbg ? clock64() : 0;
            ok = mbar_wait(smem_u32(&afull_bar[stage]), phase, abort_flag, p.err, 8);
            if (!ok) break;
            const unsigned long long tq1 = p.dbg ? clock64() : 0;
            xt_wait += tq1 - tq0;
            if (row_in) {
              const uint32_t sa = smem_u32(smem + stage * C::kStageBytes);
              float sc[8], sf[8];
              const float4* tb = reinterpret_cast<const float4*>(xf_tab + kc * kBlockK + c_log * 8);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float4 v = tb[q];
                sc[2 * q] = v.x; sf[2 * q] = v.y; sc[2 * q + 1] = v.z; sf[2 * q + 1] = v.w;
              }
              if (p.fmt_a == FMT_F16) xf_stage<FMT_F16>(sa, r0, c_log, x_first, p.Wo, sc, sf, p.xf_silu != 0, chk, mx_out);
              else xf_stage<FMT_BF16>(sa, r0, c_log, x_first, p.Wo, sc, sf, p.xf_silu != 0, chk, mx_out);
            }
            hand_over();
            if (p.dbg) xt_work += clock64() - tq1;
          }
        }
        for (int kc = 0; kc < p.a2_kchunks && ok; ++kc) {     // fused 1x1 branch: raw operand, nothing to normalise
          ok = mbar_wait(smem_u32(&afull_bar[stage]), phase, abort_flag, p.err, 8);
          if (!ok) break;
          hand_over();
        }
      }
      if (p.dbg && t == 0) {
        p.dbg[blockIdx.x * 16 + 13] = clock64() - xt_start; p.dbg[blockIdx.x * 16 + 14] = xt_wait; p.dbg[blockIdx.x * 16 + 15] = xt_work;
      }
      if (chk) {
        if (__low2float(mx_out) >= 65504.f || __high2float(mx_out) >= 65504.f) atomicCAS(p.err, 0, kErrRangeBase + SITE_GN_OUT);
      }
    }
  } else {
    // ===================== epilogue: 4 warps, TMEM lane quarter = warp % 4 =====================
    const int quarter = warp & 3;
    const int eset = (warp - 2) >> 2;                     // which half of the columns (kSets == 2), else 0
    const int ew = eset * 4 + quarter;                    // epilogue warp slot for the per-warp smem regions
    constexpr int kColsPerSet = BLOCK_N / C::kSets;
    const int cbase = eset * kColsPerSet;                 // first column of this warp inside the tile
    const int row = quarter * 32 + lane;
    const int yy = row >> p.BW_log2;
    const int xx = row & (p.BW - 1);
    // origin of this warp's 32 rows inside the tile (always a bx x by pixel rectangle, bx = min(BW, 32))
    const int wx = (quarter * 32) & (p.BW - 1), wy = (quarter * 32) >> p.BW_log2;
    float* acc_w = gn_acc + ew * (512 / C::kEpiWarps);
    float* bias_w = bias_all + ew * (1024 / C::kEpiWarps);
    int gn_img = -1, gn_nt = 0;
    // GroupNorm partials are kept per warp in shared memory across the tiles of one image and
    // flushed with fp64 global atomics when the (image, n_tile) key changes.
    auto gn_flush = [&]() {
      __syncwarp();
      const int ng2 = 2 * (kColsPerSet / p.gn_cpg);
      const int g0 = (gn_nt * BLOCK_N + cbase) / p.gn_cpg;
      for (int t = lane; t < ng2; t += 32) {
        const float val = acc_w[t];
        acc_w[t] = 0.f;

        const int grp = g0 + (t >> 1);
        if (grp < p.gn_groups)
          atomicAdd(&p.gn_stats[((long long)gn_img * p.gn_groups + grp) * 2 + (t & 1)],
                    (double)val * (double)((t & 1) ? p.stat_mul * p.stat_mul : p.stat_mul));
      }
      __syncwarp();
    };
    // kSets == 2 (Cout = 128, 4 channels per group): per-lane register partials for this warp's 16 groups,
    // reduced across the 32 rows once per image (one 32-value recursive-halving pass, 31 shuffles)
    constexpr bool kRegStats = C::kSets == 2 && BLOCK_N == 128;
    float gacc[kRegStats ? 32 : 1];
#pragma unroll
    for (int i = 0; i < (kRegStats ? 32 : 1); ++i) gacc[i] = 0.f;
    auto gn_flush_regs = [&]() {
      if constexpr (kRegStats) {
        int idx;
        const float r = halving_reduce<32>(gacc, lane, idx);
        const int grp = (gn_nt * BLOCK_N + cbase) / 4 + (idx >> 1);
        if (grp < p.gn_groups)
          atomicAdd(&p.gn_stats[((long long)gn_img * p.gn_groups + grp) * 2 + (idx & 1)],
                    (double)r * (double)((idx & 1) ? p.stat_mul * p.stat_mul : p.stat_mul));
#pragma unroll
        for (int i = 0; i < 32; ++i) gacc[i] = 0.f;
      }
    };
    int bias_nt = -1;
    auto load_bias = [&](int n_tile) {      // per-warp shared copy: L1 is ~empty with this much smem carved out
      if (n_tile == bias_nt) return;
      __syncwarp();
      for (int j = lane; j < kColsPerSet; j += 32) {
        const int c = n_tile * BLOCK_N + cbase + j;
        bias_w[j] = (p.bias && c < p.Cout) ? __ldg(p.bias + c) * p.bias_mul : 0.f;
      }
      bias_nt = n_tile;
      __syncwarp();
    };
    int acc = 0; uint32_t acc_phase = 0;
    bool ok = true;
    unsigned long long t_tfull = 0, t_start = clock64();
    unsigned long long t_e[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};   // debug: residual wait | tmem wait | math+stats+staging | store issue+wait | tile prologue | tile epilogue
    constexpr int kNChunk = kColsPerSet / C::kChunk;      // chunks of 32 columns this warp handles per tile
    const bool use_tma = C::kChunk == 32 && p.epi_mode == 1;
    const bool has_res = p.residual != nullptr;
    const bool epi_fast = p.epi_fast && use_tma && (has_res ? kResBufs >= 2 && (p.res_inplace || p.h16_slots == 2) : p.h16_slots == 2);
    // ---- residual prefetch stream (TMA): global chunk index g = tile_seq * kNChunk + c -> ring slot g % kResBufs
    uint8_t* f32_w = epi_f32 + ew * (kResBufs * kResSlot);
    const uint32_t f32_s = smem_u32(f32_w), bias_s = smem_u32(bias_w);
    uint8_t* h16_w = epi_h16 + ew * (p.h16_slots * 2048);
    const uint32_t h16_s = smem_u32(h16_w);
    // a staging slot is rewritten (by the next chunk's math or the next residual load) one chunk after the bulk
    // store that reads it was committed, unless its ring has a single slot: then wait for that store right away
    const bool deep_rings = (kResBufs == 0 || kResBufs >= 2) && (p.h16_slots == 0 || p.h16_slots >= 2);
    uint64_t* res_bar_w = res_bar + ew * kMaxResBufs;
    auto issue_residual = [&](int g, int slot) {      // whole warp calls; one elected lane issues; slot == g % kResBufs
      const int seq = g / kNChunk, c = g - seq * kNChunk;
      const int unit = unit0 + seq * unit_step;
      if (unit >= p.n_units) return;
      const TileCoord tc = tile_coord<NCTA>(p, unit, (int)rank);
      const int col0 = tc.n_tile * BLOCK_N + cbase + c * 32;
      if (tc.m_tile >= p.n_tiles_m || col0 >= p.Cout) return;     // consumer skips the same chunks
      const int tx = tc.tx, ty = tc.ty, img = tc.img;
      // the ring is only 1-2 chunks deep (shared memory goes to the operand pipeline), which would expose the
      // full DRAM latency of a tensor written two launches ago: pull the chunk one tile further ahead into L2
      if (elect_one_sync()) {
        const uint32_t bar = smem_u32(&res_bar_w[slot]);
        mbar_arrive_expect_tx(bar, (uint32_t)kResSlot);
        tma_load_4d(smem_u32(f32_w + slot * kResSlot), &tmR, bar, col0, tx * p.BW + wx, ty * p.BH + wy, img);
      }
      if (p.res_prefetch) {       // experiment (off by default, measured slower): L2 prefetch one tile further ahead
        const int unit_pf = unit + unit_step;
        const TileCoord tp = tile_coord<NCTA>(p, unit_pf < p.n_units ? unit_pf : unit, (int)rank);
        if (unit_pf < p.n_units && tp.m_tile < p.n_tiles_m && elect_one_sync())
          tma_prefetch_4d(&tmR, tp.n_tile * BLOCK_N + cbase + c * 32, tp.tx * p.BW + wx, tp.ty * p.BH + wy, tp.img);
      }
      __syncwarp();
    };
    int g_cur = 0;                          // global chunk counter of this warp
    int rslot = 0; uint32_t rphase = 0;     // g_cur % kResBufs and (g_cur / kResBufs) & 1, kept incrementally
    if (use_tma && has_res) {
      for (int g = 0; g < (kResBufs > 1 ? kResBufs - 1 : 1); ++g) issue_residual(g, g);
    }
    float sm_m = -INFINITY, sm_l = 0.f, sm_inv = 0.f;      // softmax mode: this lane's row max (log2 domain), sum, max + log2(sum)
    for (int it = 0; ok; ++it) {
      const int unit = unit_of(it);
      if (unit < 0) break;
      const unsigned long long t_top = p.dbg ? clock64() : 0;
      const TileCoord tc = tile_coord<NCTA>(p, unit, (int)rank);
      const int n_tile = tc.n_tile, m_tile = tc.m_tile, tx = tc.tx, ty = tc.ty, img = tc.img;
      const bool tile_ok = m_tile < p.n_tiles_m;
      const int x = tx * p.BW + xx, y = ty * p.BH + yy;
      const bool row_ok = tile_ok && (x < p.Wo) && (y < p.Ho);
      const long long row_off = (((long long)img * p.Ho + y) * p.Wo + x) * p.ldo;
      if (p.gn_stats && tile_ok && (img != gn_img || n_tile != gn_nt)) {
        if (gn_img >= 0) { if (kRegStats && p.gn_cpg == 4) gn_flush_regs(); else gn_flush(); }
        gn_img = img; gn_nt = n_tile;
      }
      load_bias(n_tile);
      const unsigned long long tw = p.dbg ? clock64() : 0;
      if (p.dbg) t_e[4] += tw - t_top;
      ok = mbar_wait(smem_u32(&tfull_bar[acc]), acc_phase, abort_flag, p.err, 4);
      if (p.dbg) t_tfull += clock64() - tw;
      if (!ok) break;
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BLOCK_N);
      bool handed_back = false;                // the lean path returns the accumulator stage before its math
      if constexpr (C::kChunk == 32) {
        if (use_tma) {
          // ---- asynchronous, coalesced epilogue: residual in by TMA (prefetched kResBufs-1 chunks ahead),
          //      outputs out by TMA store; each lane only ever touches its own 128-byte (64-byte) row of the
          //      swizzled staging tiles, so shared-memory accesses are conflict-free.
          const int sw = lane & 7;                 // 128B swizzle: 16-byte chunk j of row r lives at j ^ (r & 7)
          const int sw16 = (lane >> 1) & 3;        // 64B swizzle:  chunk j of row r lives at j ^ ((r >> 1) & 3)
          if (p.softmax_mode) {
            // ---- attention scores: t = alpha * log2(e) * s.  Pass 0: online row max / sum (one row per lane, no
            //      cross-lane traffic); pass 1: P = 2^(t - max - log2(sum)), stored in 16 bit by TMA.  Columns past the
            //      last key are masked.  The TMEM load of chunk c+1 is in flight while chunk c is processed, and the
            //      max / sum reductions run as four independent chains (a single warp per scheduler hides no latency).
            const float a2 = p.alpha * 1.4426950408889634f;
            if (tc.pass == 0 && n_tile == 0) { sm_m = -INFINITY; sm_l = 0.f; }
            if (tc.pass == 1 && n_tile == 0) {
              if constexpr (C::kSets == 2) {
                // two warps share a row (column halves): merge their (max, sum) through shared memory; the pair of
                // warps of one TMEM lane quarter meets on its own named barrier (both walk the same unit sequence)
                float* ex = gn_acc;                               // [epilogue warp][lane][2], GroupNorm is off in this mode
                ex[(ew * 32 + lane) * 2] = sm_m; ex[(ew * 32 + lane) * 2 + 1] = sm_l;
                asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
                const int other = (eset ^ 1) * 4 + quarter;
                const float m1 = ex[(other * 32 + lane) * 2], l1 = ex[(other * 32 + lane) * 2 + 1];
                const float m = fmaxf(sm_m, m1);
                sm_l = sm_l * ex2_approx(sm_m - m) + l1 * ex2_approx(m1 - m);
                sm_m = m;
                asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");   // scratch may be rewritten after this
              }
              sm_inv = sm_m + log2f(sm_l);
            }
            auto chunk = [&](const uint32_t (&v)[32], int c) {
              const int col0 = n_tile * BLOCK_N + cbase + c * 32;
              if (!(tile_ok && col0 < p.Cout)) return;
              float f[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
              if (p.Cout - col0 < 32) {
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = (j < p.Cout - col0) ? f[j] : -INFINITY;
              }
              if (tc.pass == 0) {
                float m4[4] = {f[0], f[1], f[2], f[3]};
#pragma unroll
                for (int j = 4; j < 32; ++j) m4[j & 3] = fmaxf(m4[j & 3], f[j]);
                const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
                const float m_new = fmaxf(sm_m, mx * a2);       // alpha > 0
                float l4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int j = 0; j < 32; ++j) l4[j & 3] += ex2_approx(fmaf(f[j], a2, -m_new));
                sm_l = sm_l * ex2_approx(sm_m - m_new) + ((l4[0] + l4[1]) + (l4[2] + l4[3]));
                sm_m = m_new;
              } else {
                const int hslot = p.h16_slots == 2 ? (g_cur & 1) : 0;
                const uint32_t hb = h16_s + hslot * 2048 + lane * 64;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  uint32_t w4[4];
#pragma unroll
                  for (int q = 0; q < 4; ++q)
                    w4[q] = pack2_16(ex2_approx(fmaf(f[8 * j + 2 * q], a2, -sm_inv)),
                                     ex2_approx(fmaf(f[8 * j + 2 * q + 1], a2, -sm_inv)), p.fmt_out);
                  sts128u(hb + ((j ^ sw16) << 4), w4[0], w4[1], w4[2], w4[3]);
                }
                fence_proxy_async();
                __syncwarp();
                if (elect_one_sync()) {
                  tma_store_4d(&tmO16, h16_s + hslot * 2048, col0, tx * p.BW + wx, ty * p.BH + wy, img);
                  tma_store_commit();
                  if (deep_rings) tma_store_wait_read<1>(); else tma_store_wait_read<0>();
                }
                __syncwarp();
              }
              ++g_cur;
            };
            uint32_t va[32], vb[32];
            tmem_ld32(t_row + cbase, va);
#pragma unroll 1
            for (int c = 0; c < kNChunk; c += 2) {
              tmem_ld_wait();
              if (c + 1 < kNChunk) tmem_ld32(t_row + cbase + (c + 1) * 32, vb);
              chunk(va, c);
              if (c + 1 < kNChunk) {
                tmem_ld_wait();
                if (c + 2 < kNChunk) tmem_ld32(t_row + cbase + (c + 2) * 32, va);
                chunk(vb, c + 1);
              }
            }
          } else if (kRegStats && epi_fast) {
            if constexpr (kRegStats) {
              // ---- lean path of the level-0 launches (128-wide tiles, two 32-column chunks per warp, fp16 stream or
              //      conv1 output, 4 channels per group, optional 16-bit residual; whole tiles only): the same arithmetic
              //      as the general loop below with formats and options fixed at compile time.  Both TMEM loads are in
              //      flight together and the accumulator stage goes back to the MMA warp before the math starts.
              uint32_t va[32], vb[32];
              tmem_ld32(t_row + cbase, va);
              tmem_ld32(t_row + cbase + 32, vb);
              unsigned long long tq = kFineDbg && p.dbg ? clock64() : 0;
              tmem_ld_wait();
              if (kFineDbg && p.dbg) { const unsigned long long t1 = clock64(); t_e[1] += t1 - tq; tq = t1; }
              tc_fence_before();
              __syncwarp();
              if (lane == 0) {
                if constexpr (NCTA == 2) mbar_arrive_cluster(smem_u32(&tempty_bar[acc]), 0);
                else mbar_arrive(smem_u32(&tempty_bar[acc]));
              }
              handed_back = true;
              const int ox = tx * p.BW + wx, oy = ty * p.BH + wy;
              const int colw = n_tile * BLOCK_N + cbase;
              const float alpha = p.alpha;
              // three-slot ring: the residual chunks to prefetch while this tile is processed are the same two column
              // chunks of this CTA's NEXT tile -> its coordinates are computed once per tile
              const bool ring3 = has_res && kResBufs == 3;
              bool nvalid = false; int ncol = 0, nox = 0, noy = 0, nimg = 0;
              if (ring3 && unit + unit_step < p.n_units) {
                const TileCoord tn = tile_coord<NCTA>(p, unit + unit_step, (int)rank);
                nvalid = true; ncol = tn.n_tile * BLOCK_N + cbase; nox = tn.tx * p.BW + wx; noy = tn.ty * p.BH + wy; nimg = tn.img;
              }
              auto fast_chunk = [&](const uint32_t (&v)[32], const int c) -> bool {
                const int slot = rslot;
                const uint32_t bw = bias_s + c * 128;
                float f[32];
                if (kFineDbg && p.dbg) tq = clock64();
                if (has_res) {
                  if (!mbar_wait(smem_u32(&res_bar_w[slot]), rphase, abort_flag, p.err, 5)) return false;
                  if (kFineDbg && p.dbg) { const unsigned long long t1 = clock64(); t_e[0] += t1 - tq; tq = t1; }
                  const uint32_t rb = f32_s + slot * kResSlot + lane * 64;
                  const float rm = p.res_mul;
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const float4 q = lds128(rb + ((j ^ sw16) << 4));
                    const float4 b0 = lds128(bw + j * 32), b1 = lds128(bw + j * 32 + 16);
                    float r8[8];
                    unpack2_16(__float_as_uint(q.x), FMT_F16, r8[0], r8[1]); unpack2_16(__float_as_uint(q.y), FMT_F16, r8[2], r8[3]);
                    unpack2_16(__float_as_uint(q.z), FMT_F16, r8[4], r8[5]); unpack2_16(__float_as_uint(q.w), FMT_F16, r8[6], r8[7]);
                    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i) f[8 * j + i] = fmaf(r8[i], rm, fmaf(__uint_as_float(v[8 * j + i]), alpha, bb[i]));
                  }
                } else {
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    const float4 b4 = lds128(bw + j * 16);
                    f[4 * j] = fmaf(__uint_as_float(v[4 * j]), alpha, b4.x);
                    f[4 * j + 1] = fmaf(__uint_as_float(v[4 * j + 1]), alpha, b4.y);
                    f[4 * j + 2] = fmaf(__uint_as_float(v[4 * j + 2]), alpha, b4.z);
                    f[4 * j + 3] = fmaf(__uint_as_float(v[4 * j + 3]), alpha, b4.w);
                  }
                }
                if (kFineDbg && p.dbg) { const unsigned long long t1 = clock64(); t_e[6] += t1 - tq; tq = t1; }
                epi_stats_lane<4>(f, row_ok, gacc + c * 16);
                if (kFineDbg && p.dbg) { const unsigned long long t1 = clock64(); t_e[7] += t1 - tq; tq = t1; }
                // staging tile of the store: with a residual, the ring slot the residual arrived in (same 32 x 64 B
                // swizzled layout; every lane overwrites exactly the 64 bytes it has just read), which leaves the
                // shared memory of the separate output tiles to a fourth pipeline stage
                const uint32_t ht = (has_res && p.res_inplace) ? f32_s + slot * kResSlot : h16_s + (g_cur & 1) * 2048;
                const uint32_t hb = ht + lane * 64;
                __half2 mx = __float2half2_rn(0.f);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const uint32_t w0 = pack2_16(f[8 * j], f[8 * j + 1], FMT_F16), w1 = pack2_16(f[8 * j + 2], f[8 * j + 3], FMT_F16);
                  const uint32_t w2 = pack2_16(f[8 * j + 4], f[8 * j + 5], FMT_F16), w3 = pack2_16(f[8 * j + 6], f[8 * j + 7], FMT_F16);
                  mx = __hmax2(__hmax2(mx, __habs2(*reinterpret_cast<const __half2*>(&w0))), __habs2(*reinterpret_cast<const __half2*>(&w1)));
                  mx = __hmax2(__hmax2(mx, __habs2(*reinterpret_cast<const __half2*>(&w2))), __habs2(*reinterpret_cast<const __half2*>(&w3)));
                  sts128u(hb + ((j ^ sw16) << 4), w0, w1, w2, w3);
                }
                if (p.sat_check && row_ok && (__low2float(mx) >= 65504.f || __high2float(mx) >= 65504.f))
                  atomicCAS(p.err, 0, kErrRangeBase + SITE_XCOPY);
                if (kFineDbg && p.dbg) { const unsigned long long t1 = clock64(); t_e[8] += t1 - tq; tq = t1; }
                fence_proxy_async();
                __syncwarp();
                if (kFineDbg && p.dbg) { const unsigned long long t1 = clock64(); t_e[2] += t1 - tq; tq = t1; }
                if (elect_one_sync()) {
                  tma_store_4d(&tmO16, ht, colw + c * 32, ox, oy, img);
                  tma_store_commit();
                  tma_store_wait_read<1>();                      // the previous chunk's staging tile is free again
                }
                __syncwarp();
                if (kFineDbg && p.dbg) { const unsigned long long t1 = clock64(); t_e[3] += t1 - tq; tq = t1; }
                if (ring3) {                                     // refill the slot freed by the wait above
                  if (nvalid && elect_one_sync()) {
                    const int fs = slot == 0 ? 2 : slot - 1;
                    const uint32_t bar = smem_u32(&res_bar_w[fs]);
                    mbar_arrive_expect_tx(bar, (uint32_t)kResSlot);
                    tma_load_4d(f32_s + fs * kResSlot, &tmR, bar, ncol + c * 32, nox, noy, nimg);
                  }
                  __syncwarp();
                } else if (has_res) {
                  issue_residual(g_cur + kResBufs - 1, slot == 0 ? kResBufs - 1 : slot - 1);
                }
                ++g_cur;
                rslot = (rslot + 1 == kResBufs) ? 0 : rslot + 1;
                rphase ^= (rslot == 0);
                return true;
              };
              ok = fast_chunk(va, 0) && fast_chunk(vb, 1);
              if (!ok) break;
            }
          } else
#pragma unroll (kRegStats ? 2 : 1)
          for (int c = 0; c < kNChunk; ++c, ++g_cur, rslot = (rslot + 1 == kResBufs) ? 0 : rslot + 1, rphase ^= (rslot == 0)) {
            const int c0 = c * 32;                              // column inside this warp's share
            const int col0 = n_tile * BLOCK_N + cbase + c0;
            const bool chunk_ok = tile_ok && col0 < p.Cout;     // warp-uniform
            uint32_t v[32];
            tmem_ld32(t_row + cbase + c0, v);
            const int slot = kResBufs ? rslot : 0;
            const uint32_t fb = f32_s + slot * kResSlot + lane * 128;     // fp32 tile row (only meaningful with 4096-byte slots)
            const uint32_t bw = bias_s + c0 * 4;
            float f[32];
            unsigned long long tq = kFineDbg && p.dbg ? clock64() : 0;
            if (has_res && chunk_ok) {
              ok = mbar_wait(smem_u32(&res_bar_w[slot]), rphase, abort_flag, p.err, 5);
              if (!ok) break;
              if (kFineDbg && p.dbg) { const unsigned long long t1 = clock64(); t_e[0] += t1 - tq; tq = t1; }
              tmem_ld_wait();
              if (p.res16) {
                // 16-bit residual stream (MIXED: fp16 x 2^-6): 32 rows x 64 B, 64B-swizzled; value = res_mul * stored
                const uint32_t rb = f32_s + slot * kResSlot + lane * 64;
                const float rm = p.res_mul;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float4 q = lds128(rb + ((j ^ sw16) << 4));
                  float r8[8];
                  unpack2_16(__float_as_uint(q.x), p.fmt_out, r8[0], r8[1]); unpack2_16(__float_as_uint(q.y), p.fmt_out, r8[2], r8[3]);
                  unpack2_16(__float_as_uint(q.z), p.fmt_out, r8[4], r8[5]); unpack2_16(__float_as_uint(q.w), p.fmt_out, r8[6], r8[7]);
                  const float4 b0 = lds128(bw + j * 32), b1 = lds128(bw + j * 32 + 16);
                  f[8 * j] = fmaf(r8[0], rm, fmaf(__uint_as_float(v[8 * j]), p.alpha, b0.x));
                  f[8 * j + 1] = fmaf(r8[1], rm, fmaf(__uint_as_float(v[8 * j + 1]), p.alpha, b0.y));
                  f[8 * j + 2] = fmaf(r8[2], rm, fmaf(__uint_as_float(v[8 * j + 2]), p.alpha, b0.z));
                  f[8 * j + 3] = fmaf(r8[3], rm, fmaf(__uint_as_float(v[8 * j + 3]), p.alpha, b0.w));
                  f[8 * j + 4] = fmaf(r8[4], rm, fmaf(__uint_as_float(v[8 * j + 4]), p.alpha, b1.x));
                  f[8 * j + 5] = fmaf(r8[5], rm, fmaf(__uint_as_float(v[8 * j + 5]), p.alpha, b1.y));
                  f[8 * j + 6] = fmaf(r8[6], rm, fmaf(__uint_as_float(v[8 * j + 6]), p.alpha, b1.z));
                  f[8 * j + 7] = fmaf(r8[7], rm, fmaf(__uint_as_float(v[8 * j + 7]), p.alpha, b1.w));
                }
              } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 r4 = lds128(fb + ((j ^ sw) << 4));
                const float4 b4 = lds128(bw + j * 16);
                f[4 * j] = fmaf(__uint_as_float(v[4 * j]), p.alpha, b4.x) + r4.x;
                f[4 * j + 1] = fmaf(__uint_as_float(v[4 * j + 1]), p.alpha, b4.y) + r4.y;
                f[4 * j + 2] = fmaf(__uint_as_float(v[4 * j + 2]), p.alpha, b4.z) + r4.z;
                f[4 * j + 3] = fmaf(__uint_as_float(v[4 * j + 3]), p.alpha, b4.w) + r4.w;
              }
              }
            } else {
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 b4 = lds128(bw + j * 16);
                f[4 * j] = fmaf(__uint_as_float(v[4 * j]), p.alpha, b4.x);
                f[4 * j + 1] = fmaf(__uint_as_float(v[4 * j + 1]), p.alpha, b4.y);
                f[4 * j + 2] = fmaf(__uint_as_float(v[4 * j + 2]), p.alpha, b4.z);
                f[4 * j + 3] = fmaf(__uint_as_float(v[4 * j + 3]), p.alpha, b4.w);
              }
            }
            if (chunk_ok) {
              if (p.relu) {
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
              }
              if (kFineDbg && p.dbg) { const unsigned long long t1 = clock64(); t_e[6] += t1 - tq; tq = t1; }
              if (p.gn_stats) {   // statistics of the finished fp32 values (bias and residual included)
                if (kRegStats && p.gn_cpg == 4) epi_stats_lane<4>(f, row_ok, gacc + (kRegStats ? c * 16 : 0));
                else if (p.gn_cpg == 4) epi_stats_chunk<4>(f, row_ok, lane, acc_w, c0 / 4);
                else if (p.gn_cpg == 8) epi_stats_chunk<8>(f, row_ok, lane, acc_w, c0 / 8);
                else epi_stats_chunk<16>(f, row_ok, lane, acc_w, c0 / 16);
              }
              if (kFineDbg && p.dbg) { const unsigned long long t1 = clock64(); t_e[7] += t1 - tq; tq = t1; }
              if (p.out_f32) {
#pragma unroll
                for (int j = 0; j < 8; ++j) sts128(fb + ((j ^ sw) << 4), f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
              }
              const int hslot = p.h16_slots == 2 ? (g_cur & 1) : 0;
              const uint32_t hb = h16_s + hslot * 2048 + lane * 64;
              if (p.out_16) {
                if (p.out16_scale == 1.f) {
                  __half2 mx = __float2half2_rn(0.f);
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const uint32_t w0 = pack2_16(f[8 * j], f[8 * j + 1], p.fmt_out), w1 = pack2_16(f[8 * j + 2], f[8 * j + 3], p.fmt_out);
                    const uint32_t w2 = pack2_16(f[8 * j + 4], f[8 * j + 5], p.fmt_out), w3 = pack2_16(f[8 * j + 6], f[8 * j + 7], p.fmt_out);
                    // fp16 range check of the stored residual stream: running |max| of the packed pairs (one HMNMX2 per
                    // word); the saturating conversion leaves 65504 in a half that overflowed
                    if (p.sat_check) {
                      mx = __hmax2(mx, __habs2(*reinterpret_cast<const __half2*>(&w0)));
                      mx = __hmax2(mx, __habs2(*reinterpret_cast<const __half2*>(&w1)));
                      mx = __hmax2(mx, __habs2(*reinterpret_cast<const __half2*>(&w2)));
                      mx = __hmax2(mx, __habs2(*reinterpret_cast<const __half2*>(&w3)));
                    }
                    sts128u(hb + ((j ^ sw16) << 4), w0, w1, w2, w3);
                  }
                  if (p.sat_check && row_ok && (__low2float(mx) >= 65504.f || __high2float(mx) >= 65504.f))
                    atomicCAS(p.err, 0, kErrRangeBase + SITE_XCOPY);
                } else {
                  // scaled 16-bit copy of the residual stream (MIXED mode: fp16 x 2^-6); a value beyond the fp16
                  // range must not saturate silently -> error word (only the few launches that write such copies)
                  const float s16 = p.out16_scale;
                  float mx = 0.f;
#pragma unroll
                  for (int j = 0; j < 32; ++j) mx = fmaxf(mx, fabsf(f[j]));
                  if (p.fmt_out == FMT_F16 && row_ok && !(mx * s16 <= 65504.f)) atomicCAS(p.err, 0, kErrRangeBase + SITE_XCOPY);
#pragma unroll
                  for (int j = 0; j < 4; ++j)
                    sts128u(hb + ((j ^ sw16) << 4), pack2_16(f[8 * j] * s16, f[8 * j + 1] * s16, p.fmt_out),
                            pack2_16(f[8 * j + 2] * s16, f[8 * j + 3] * s16, p.fmt_out),
                            pack2_16(f[8 * j + 4] * s16, f[8 * j + 5] * s16, p.fmt_out),
                            pack2_16(f[8 * j + 6] * s16, f[8 * j + 7] * s16, p.fmt_out));
                }
              }
              if (kFineDbg && p.dbg) { const unsigned long long t1 = clock64(); t_e[8] += t1 - tq; tq = t1; }
              fence_proxy_async();
              __syncwarp();
              if (kFineDbg && p.dbg) { const unsigned long long t1 = clock64(); t_e[2] += t1 - tq; tq = t1; }
              if (elect_one_sync()) {
                const int ox = tx * p.BW + wx, oy = ty * p.BH + wy;
                if (p.out_f32) tma_store_4d(&tmO32, f32_s + slot * kResSlot, col0, ox, oy, img);
                if (p.out_16) tma_store_4d(&tmO16, h16_s + hslot * 2048, col0, ox, oy, img);
                tma_store_commit();
                if (deep_rings) tma_store_wait_read<1>();       // the previous chunk's staging tiles are free again
                else tma_store_wait_read<0>();
              }
              __syncwarp();
              if (kFineDbg && p.dbg) { const unsigned long long t1 = clock64(); t_e[3] += t1 - tq; tq = t1; }
            }
            if (has_res) {
              if (!chunk_ok) {                 // skipped chunk: still make sure the slot about to be refilled is free
                if (elect_one_sync()) { if (deep_rings) tma_store_wait_read<1>(); else tma_store_wait_read<0>(); }
                __syncwarp();
              }
              issue_residual(g_cur + (kResBufs > 1 ? kResBufs - 1 : 1),
                             kResBufs > 1 ? (slot == 0 ? kResBufs - 1 : slot - 1) : 0);   // refills the slot freed by the wait above
            }
          }
          if (!ok) break;
        } else {
          // reference epilogue: every lane writes its own pixel row straight from registers
#pragma unroll 1
          for (int c0 = 0; c0 < kColsPerSet; c0 += 32) {
            const int col0 = n_tile * BLOCK_N + cbase + c0;
            if (col0 >= p.Cout) break;
            uint32_t v[32];
            tmem_ld32(t_row + cbase + c0, v);
            tmem_ld_wait();
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) * p.alpha + bias_w[c0 + j];
            if (row_ok) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                if (col0 + j < p.Cout) {
                  if (p.residual && p.res16) {
                    const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(p.residual) + row_off + col0 + j);
                    float r0, r1, r2, r3;
                    unpack2_16(u.x, p.fmt_out, r0, r1); unpack2_16(u.y, p.fmt_out, r2, r3);
                    f[j] = fmaf(r0, p.res_mul, f[j]); f[j + 1] = fmaf(r1, p.res_mul, f[j + 1]);
                    f[j + 2] = fmaf(r2, p.res_mul, f[j + 2]); f[j + 3] = fmaf(r3, p.res_mul, f[j + 3]);
                  } else if (p.residual) {
                    const float4 b = *reinterpret_cast<const float4*>(p.residual + row_off + col0 + j);
                    f[j] += b.x; f[j + 1] += b.y; f[j + 2] += b.z; f[j + 3] += b.w;
                  }
                  if (p.relu) { f[j] = fmaxf(f[j], 0.f); f[j + 1] = fmaxf(f[j + 1], 0.f); f[j + 2] = fmaxf(f[j + 2], 0.f); f[j + 3] = fmaxf(f[j + 3], 0.f); }
                  if (p.out_f32)
                    *reinterpret_cast<float4*>(p.out_f32 + row_off + col0 + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
                  if (p.out_16) {
                    const float s16 = p.out16_scale;
                    if (p.fmt_out == FMT_F16 && s16 != 1.f &&
                        !(fmaxf(fmaxf(fabsf(f[j]), fabsf(f[j + 1])), fmaxf(fabsf(f[j + 2]), fabsf(f[j + 3]))) * s16 <= 65504.f))
                      atomicCAS(p.err, 0, kErrRangeBase + SITE_XCOPY);
                    uint2 u; u.x = pack2_16(f[j] * s16, f[j + 1] * s16, p.fmt_out); u.y = pack2_16(f[j + 2] * s16, f[j + 3] * s16, p.fmt_out);
                    *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p.out_16) + row_off + col0 + j) = u;
                  }
                }
              }
            }
            if (p.gn_stats) {
              if (kRegStats && p.gn_cpg == 4) {
                if (c0 == 0) epi_stats_lane<4>(f, row_ok, gacc); else epi_stats_lane<4>(f, row_ok, gacc + (kRegStats ? 16 : 0));
              }
              else if (p.gn_cpg == 4) epi_stats_chunk<4>(f, row_ok, lane, acc_w, c0 / 4);
              else if (p.gn_cpg == 8) epi_stats_chunk<8>(f, row_ok, lane, acc_w, c0 / 8);
              else epi_stats_chunk<16>(f, row_ok, lane, acc_w, c0 / 16);
            }
          }
        }
      } else {
        // BLOCK_N == 16 (the 8-channel moments head): tiny, written straight from registers
        uint32_t v[16];
        tmem_ld16(t_row, v);
        tmem_ld_wait();
        const int col0 = n_tile * BLOCK_N;
        if (row_ok && col0 < p.Cout) {
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]) * p.alpha + bias_w[j];
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            if (col0 + j < p.Cout) {
              if (p.residual) {
                const float4 b = *reinterpret_cast<const float4*>(p.residual + row_off + col0 + j);
                f[j] += b.x; f[j + 1] += b.y; f[j + 2] += b.z; f[j + 3] += b.w;
              }
              if (p.relu) { f[j] = fmaxf(f[j], 0.f); f[j + 1] = fmaxf(f[j + 1], 0.f); f[j + 2] = fmaxf(f[j + 2], 0.f); f[j + 3] = fmaxf(f[j + 3], 0.f); }
              if (p.out_f32)
                *reinterpret_cast<float4*>(p.out_f32 + row_off + col0 + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
              if (p.out_16) {
                uint2 u; u.x = pack2_16(f[j], f[j + 1], p.fmt_out); u.y = pack2_16(f[j + 2], f[j + 3], p.fmt_out);
                *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p.out_16) + row_off + col0 + j) = u;
              }
            }
          }
        }
      }
      // all TMEM reads of this accumulator stage are done -> hand it back to the MMA warp
      const unsigned long long t_end0 = p.dbg ? clock64() : 0;
      if (!handed_back) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (NCTA == 2) mbar_arrive_cluster(smem_u32(&tempty_bar[acc]), 0);   // the leader's MMA thread waits
          else mbar_arrive(smem_u32(&tempty_bar[acc]));
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      if (p.dbg) t_e[5] += clock64() - t_end0;
    }
    if (use_tma) {      // staging tiles must outlive the bulk stores that read them
      if (elect_one_sync()) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      __syncwarp();
    }
    if (p.gn_stats && gn_img >= 0) { if (kRegStats && p.gn_cpg == 4) gn_flush_regs(); else gn_flush(); }
    if (p.dbg && warp == 2 && lane == 0) {
      p.dbg[blockIdx.x * 16 + 5] = clock64() - t_start; p.dbg[blockIdx.x * 16 + 6] = t_tfull;
      p.dbg[blockIdx.x * 16 + 7] = t_e[0]; p.dbg[blockIdx.x * 16 + 8] = t_e[1]; p.dbg[blockIdx.x * 16 + 9] = t_e[2]; p.dbg[blockIdx.x * 16 + 10] = t_e[3];
      p.dbg[blockIdx.x * 16 + 11] = t_e[4]; p.dbg[blockIdx.x * 16 + 12] = t_e[5];
      if (!XF) { p.dbg[blockIdx.x * 16 + 13] = t_e[6]; p.dbg[blockIdx.x * 16 + 14] = t_e[7]; p.dbg[blockIdx.x * 16 + 15] = t_e[8]; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (NCTA == 2) cluster_sync_all();     // the peer may still be reading our smem / arriving on our barriers
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    if constexpr (NCTA == 2) tmem_dealloc_2cta(tmem_base, C::kTmemCols);
    else tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
DevState g_dev[kMaxDevices];
int g_ncta_max = 2;     // SFV_NCTA=1 disables CTA pairs (A/B experiments)
int g_halo = 2;         // SFV_HALO=0 disables the shared A halo box; 2 also uses it for BLOCK_N = 256 (2-stage pipeline)
int g_epi_slots_auto = 1, g_epi_slots_r = -1, g_epi_slots_h = -1;   // SFV_EPI_SLOTS=auto|fixed|r,h
int g_res_prefetch = 0;  // SFV_RES_PREFETCH=1: L2-prefetch residual chunks one tile ahead (measured slower: 571->561 us off)
int g_halo_boff = 0;    // descriptor base-offset for the shifted taps: measured WRONG on B200 (the swizzle is a function of
                        // the absolute smem address bits), so it stays 0; SFV_HALO_BOFF=1 reproduces the failing variant
int g_epi_mode = 1;
int g_epi_fast = 1;     // SFV_EPI_FAST=0: level-0 launches take the general epilogue loop; 2: lean loop with separate output tiles (A/B)
unsigned long long* g_dbg = nullptr;   // SFV_TC_DEBUG=1: per-CTA role cycle counters, printed after each launch (synchronous)     // SFV_EPI=0: write rows straight from registers; 1: coalesced via smem transpose

int tc_init() {
  if (g_encode) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  SFV_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (!fn || qres != cudaDriverEntryPointSuccess)
    return fail(SFV_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  if (const char* e = getenv("SFV_NCTA")) g_ncta_max = atoi(e);
  if (const char* e = getenv("SFV_EPI")) g_epi_mode = atoi(e);
  if (const char* e = getenv("SFV_EPI_FAST")) g_epi_fast = atoi(e);
  if (const char* e = getenv("SFV_HALO")) g_halo = atoi(e);
  if (const char* e = getenv("SFV_HALO_BOFF")) g_halo_boff = atoi(e);
  if (const char* e = getenv("SFV_RES_PREFETCH")) g_res_prefetch = atoi(e);
  if (const char* e = getenv("SFV_EPI_SLOTS")) {
    if (!strcmp(e, "fixed")) g_epi_slots_auto = 0;
    else if (strcmp(e, "auto")) sscanf(e, "%d,%d", &g_epi_slots_r, &g_epi_slots_h);
  }
  if (const char* e = getenv("SFV_TC_DEBUG")) { if (atoi(e)) SFV_CUDA(cudaMalloc(&g_dbg, 8 * 16 * 256)); }
  g_encode = (EncodeTiledFn)fn;
  return 0;
}

int encode_map(CUtensorMap* m, int fmt, int rank, const void* ptr, const cuuint64_t* dims,
               const cuuint64_t* strides_bytes, const cuuint32_t* box, int swizzle_bytes = 128, bool f32 = false) {
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  if (((uintptr_t)ptr & 15) != 0) return fail(SFV_ERR_INVALID, "TMA base not 16B aligned");
  for (int i = 1; i < rank; ++i)
    if (strides_bytes[i] % 16 != 0)
      return fail(SFV_ERR_INVALID, "TMA stride %d = %llu not a multiple of 16 B", i,
                  (unsigned long long)strides_bytes[i]);
  const CUtensorMapDataType dt = f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                     : (fmt == FMT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
  CUresult r = g_encode(m, dt,
                        (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides_bytes + 1, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SFV_ERR_CUDA, "cuTensorMapEncodeTiled failed: %d", (int)r);
  return 0;
}

template <int BLOCK_N, int NCTA, bool HALO = false, bool XF = false>
int launch_cfg(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& ma2, const CUtensorMap& mr, const CUtensorMap& mo32,
               const CUtensorMap& mo16, const TcParams& p_in, const DevState& ds, cudaStream_t s, const char* tag) {
  using C = Cfg<BLOCK_N, NCTA, HALO, XF>;
  static unsigned long long attr_devs = 0;        // cudaFuncSetAttribute is per device
  if (first_use_on_device(attr_devs, ds.dev))
    SFV_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<BLOCK_N, NCTA, HALO, XF>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  C::kSmemLimit));
  // shared-memory plan: staging rings only where this launch needs them, the rest goes to pipeline stages
  TcParams p = p_in;
  const bool need_f32 = p.epi_mode == 1 && (p.residual || p.out_f32);
  const bool need_h16 = p.epi_mode == 1 && p.out_16;
  // Staging depth vs pipeline depth.  Preferred: deep epilogue rings (residual prefetched 1-2 chunks ahead, stores
  // overlapped).  A HALO stage holds three k-chunks (up to 65 KB), so deep rings can leave only two stages and the
  // MMA warp starves; when the epilogue has slack (>= ~2000 tensor-pipe cycles per 32-column chunk, 4000 when a
  // residual load sits in the chain) single-slot rings buy the third stage instead (measured: 448 -> 379 us on the
  // 512->512 conv2 at 128x128, but 513 -> 551 us on the 256->256 one, hence the threshold).
  p.h16_slots = need_h16 ? 2 : 0;
  // eight epilogue warps hold half as many chunks each: two-slot rings suffice when staging is scarce
  p.res_bufs = !need_f32 ? 0 : (p.res16 ? kMaxResBufs : ((C::kSets == 2 && (need_h16 || BLOCK_N == 256)) ? 2 : 3));
  {
    const int want = HALO ? 3 : 6;      // non-HALO stages hold one k-chunk (24-32 KB): six of them cover the load latency
    const long long k_chunks_total = (long long)p.ntaps * p.kchunks + p.a2_kchunks;
    constexpr int chunks_per_warp = BLOCK_N >= 32 ? BLOCK_N / C::kSets / 32 : 1;
    const long long chunk_budget = k_chunks_total * (BLOCK_N * kBlockM * kBlockK / 4096) / chunks_per_warp;
    const int floor_stages = HALO ? 2 : 3;
    const bool have_slack = chunk_budget >= (p.residual ? 4000 : 2000);
    if (g_epi_slots_auto && (C::stages_for(p.res_bufs, p.h16_slots, p.res_slot) < floor_stages ||
                             (C::stages_for(p.res_bufs, p.h16_slots, p.res_slot) < want && have_slack))) {
      const int need = C::stages_for(p.res_bufs, p.h16_slots, p.res_slot) < floor_stages && !have_slack ? floor_stages : want;
      const int cand[6][2] = {{3, 2}, {3, 1}, {2, 2}, {2, 1}, {1, 2}, {1, 1}};
      int best_r = p.res_bufs, best_h = p.h16_slots;
      for (int i = 0; i < 6; ++i) {
        const int r = need_f32 ? cand[i][0] : 0, h = need_h16 ? cand[i][1] : 0;
        if (r > p.res_bufs || h > p.h16_slots) continue;
        best_r = r; best_h = h;                                  // candidates get shallower: the last one has the most stages
        if (C::stages_for(r, h, p.res_slot) >= need) break;
      }
      p.res_bufs = best_r; p.h16_slots = best_h;
    }
    if (g_epi_slots_r >= 0 && need_f32) p.res_bufs = g_epi_slots_r;       // SFV_EPI_SLOTS=r,h experiment override
    if (g_epi_slots_h >= 0 && need_h16) p.h16_slots = g_epi_slots_h;
  }
  p.res_inplace = 0;
  if (p.epi_fast && p.res16 && C::kSets == 2 && BLOCK_N == 128 && g_epi_fast != 2 && g_epi_slots_r < 0 && g_epi_slots_h < 0) {
    // lean level-0 epilogue with a 16-bit residual: three ring slots per warp serve the residual load, the math and
    // the store of consecutive chunks; no separate output tiles
    p.res_inplace = 1; p.res_bufs = 3; p.h16_slots = 0;
  }
  p.num_stages = C::stages_for(p.res_bufs, p.h16_slots, p.res_slot);
  SFV_CHECK(p.num_stages >= (HALO ? 2 : 3), "tc_gemm: pipeline too shallow (%d stages)", p.num_stages);
  const int smem_bytes = C::smem_bytes(p.num_stages, p.res_bufs, p.h16_slots, p.res_slot);
  const int max_units = ds.num_sms / NCTA;
  const int n_sched = p.softmax_mode ? p.n_groups : p.n_units;      // schedulable items: tiles, or whole m-tile groups
  const int grid = (n_sched < max_units ? n_sched : max_units) * NCTA;
  // algorithmic FLOPs: 2 * (valid output pixels) * Cout * K, K = taps * 64-wide chunks (no tile padding counted)
  const double flops = 2.0 * (double)p.Wo * p.Ho * p.n_img * p.Cout * (double)p.ntaps * p.kchunks * kBlockK;
  // conv_in (K = 27, write-bound) is accounted with the other first-layer kernels, not with the tensor-bound GEMMs
  ProfScope prof(p.u8_src ? PROF_IGEMM : PROF_TC_GEMM, flops, s, tag);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(C::kThreads); cfg.dynamicSmemBytes = smem_bytes; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = NCTA; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  if (g_dbg) SFV_CUDA(cudaMemsetAsync(g_dbg, 0, 8 * 16 * 256, s));
  SFV_CUDA(cudaLaunchKernelEx(&cfg, tc_gemm_kernel<BLOCK_N, NCTA, HALO, XF>, ma, mb, ma2, mr, mo32, mo16, p));
  SFV_LAUNCH_OK();
  if (g_dbg) {
    std::vector<unsigned long long> h(16 * 256);
    SFV_CUDA(cudaStreamSynchronize(s));
    SFV_CUDA(cudaMemcpy(h.data(), g_dbg, 8 * 16 * 256, cudaMemcpyDeviceToHost));
    double a[16] = {0}; int n = 0;
    for (int b = 0; b < grid; b += NCTA) { for (int k = 0; k < 16; ++k) a[k] += (double)h[b * 16 + k]; ++n; }
    for (int k = 0; k < 16; ++k) a[k] /= n;
    fprintf(stderr, "TCDBG %s st=%d r=%d h=%d | tiles/cta %.1f | producer total %.0f wait_empty %.0f | mma total %.0f wait_full %.0f wait_tempty %.0f | epi total %.0f wait_tfull %.0f [res_wait %.0f tmem %.0f fence %.0f store %.0f | tile_pre %.0f tile_post %.0f | fma %.0f stats %.0f pack_sts %.0f]\n",
            tag, p.num_stages, p.res_bufs, p.h16_slots, (double)p.n_units / (grid / NCTA), a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11], a[12], a[13], a[14], a[15]);
  }
  return 0;
}

}  // namespace

// state of the current device, created on first use (error word, SM count)
int dev_state(DevState** out) {
  int dev = 0;
  SFV_CUDA(cudaGetDevice(&dev));
  SFV_CHECK(dev >= 0 && dev < kMaxDevices, "device ordinal %d out of range", dev);
  DevState& d = g_dev[dev];
  if (!d.err_flag) {
    SFV_CUDA(cudaDeviceGetAttribute(&d.num_sms, cudaDevAttrMultiProcessorCount, dev));
    SFV_CUDA(cudaMalloc(&d.err_flag, sizeof(int)));
    SFV_CUDA(cudaMemset(d.err_flag, 0, sizeof(int)));
    d.dev = dev;
  }
  *out = &d;
  return 0;
}

int launch_tc_gemm(const TcGemmArgs& a, cudaStream_t s) {
  SFV_TRY(tc_init());
  DevState* dsp = nullptr;
  SFV_TRY(dev_state(&dsp));
  const DevState& ds = *dsp;
  const int fmt_b = a.fmt_split ? a.fmt_b : a.fmt, fmt_out = a.fmt_split ? a.fmt_out : a.fmt;
  // measured on B200: an instruction descriptor with a_format != b_format under kind::f16 raises an illegal-instruction
  // fault, so the two operands always share a format; only the 16-bit OUTPUT format is free
  SFV_CHECK(fmt_b == a.fmt, "tc_gemm: A and B operands must share one 16-bit format (got %d / %d)", a.fmt, fmt_b);
  SFV_CHECK(a.BW * a.BH == kBlockM && (a.BW & (a.BW - 1)) == 0, "tc_gemm: bad tile %dx%d", a.BW, a.BH);
  SFV_CHECK(a.ntaps >= 1 && a.ntaps <= 9 && a.kchunks >= 1, "tc_gemm: bad taps/kchunks");

  SFV_CHECK(a.block_n <= 256, "tc_gemm: block_n > 256");
  SFV_CHECK(a.ldo % 4 == 0, "tc_gemm: ldo %% 4 != 0");
  CUtensorMap ma, mb, ma2;
  if (a.u8_src)
    SFV_CHECK(a.block_n == 128 && a.BW == 128 && a.BH == 1 && a.ntaps == 1 && a.kchunks == 1 && !a.a2 && a.Wo % 8 == 0 &&
                  ((uintptr_t)a.u8_src & 3) == 0 && a.Cout == 128,
              "tc_gemm: conv_in mode needs 128-pixel row tiles, one k-chunk, Cout 128 and a 4-byte aligned frame");
  // HALO variant: 3x3 stride-1 conv, 128-pixel row-segment tiles, taps ordered row-major with dx = -1,0,+1
  // BLOCK_N = 256: a HALO stage is 65 KB and its kernel has four epilogue warps, so (measured, 8 frames of 512^2)
  //   Cin = 128 (two k-chunks per tap, epilogue-bound)               239 us HALO vs 215 us plain with eight epilogue warps
  //   Cin = 256 with residual + fp32 + 16-bit outputs (two stages)   489 us HALO vs 475 us plain
  // go to the plain kernel; every other 256-wide 3x3 layer is faster with HALO (417 vs 434, 375 vs 396, 385 vs 417 us).
  const bool halo256 = a.block_n == 256 && g_halo >= 2 && a.kchunks >= 4 &&
                       !(a.kchunks < 8 && a.residual && a.out_f32 && a.out_16);
  const bool halo = !a.u8_src && g_halo && g_ncta_max >= 2 && !a.b_batched && a.halo_ok && a.ntaps == 9 && a.BW == 128 && a.BH == 1 &&
                    (a.block_n == 128 || halo256) && a.dim_x == 1 && (long long)ceil_div(a.Wo, 128) * a.Ho * a.Nimg >= 2;
  if (!a.u8_src) {
    cuuint64_t dims[5], strides[5]; cuuint32_t box[5];
    for (int i = 0; i < 5; ++i) {
      dims[i] = i < a.a_rank ? a.a_dims[i] : 1;
      strides[i] = i < a.a_rank ? a.a_strides[i] : (i > 0 ? strides[i - 1] * dims[i - 1] : 2);
      box[i] = i < a.a_rank ? a.a_box[i] : 1;
    }
    if (a.a2) {   // same geometry as A, different channel count
      SFV_CHECK(a.a_rank == 4 && a.a2_cin % 64 == 0, "tc_gemm: fused 1x1 branch needs a plain NHWC stride-1 A tensor");
      cuuint64_t d2[5], s2[5];
      for (int i = 0; i < 5; ++i) { d2[i] = dims[i]; s2[i] = strides[i]; }
      d2[0] = a.a2_cin;
      s2[1] = (cuuint64_t)a.a2_cin * 2; s2[2] = s2[1] * d2[1]; s2[3] = s2[2] * d2[2]; s2[4] = s2[3] * d2[3];
      SFV_TRY(encode_map(&ma2, a.fmt, 5, a.a2, d2, s2, box));   // 16-bit data: the map's element type only matters for OOB fill (zeros)
    }
    if (halo) box[a.dim_x] = 130;
    strides[0] = 2;
    SFV_TRY(encode_map(&ma, a.fmt, 5, a.a, dims, strides, box));
    if (!a.a2) ma2 = ma;
  }
  // CTA pairs (cta_group::2): two adjacent 128-pixel tiles share one B tile, halving the shared-memory
  // operand traffic per MMA.  A batched B (attention) must be the same for both tiles of a pair.
  const int tiles_m_per_img = ceil_div(a.Wo, a.BW) * ceil_div(a.Ho, a.BH);
  int ncta = (g_ncta_max >= 2 && a.block_n >= 32 && (!a.b_batched || tiles_m_per_img % 2 == 0) &&
              tiles_m_per_img * a.Nimg >= 2 && !a.u8_src) ? 2 : 1;
  const bool xf = a.xf_stats != nullptr;
  if (xf && !(halo && a.block_n == 256 && ncta == 2 && a.a_dims[0] <= 512 && a.a_dims[0] % 32 == 0 && a.dim_y == 2))
    return TC_NOT_FUSABLE;       // no transform variant for this launch: the caller runs the stand-alone GroupNorm pass
  {
    cuuint64_t dims[3] = {a.b_k, a.b_rows, a.b_batched ? (cuuint64_t)a.Nimg : 1};
    cuuint64_t strides[3] = {2, a.b_row_stride, a.b_batched ? a.b_batch_stride : a.b_row_stride * a.b_rows};
    cuuint32_t box[3] = {(cuuint32_t)kBlockK, (cuuint32_t)(a.block_n / ncta), 1};
    SFV_TRY(encode_map(&mb, fmt_b, 3, a.b, dims, strides, box));
  }
  if (a.u8_src) ma = ma2 = mb;          // unused by the kernel in this mode
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.u8_src = a.u8_src;
  p.dim_x = a.dim_x; p.dim_y = a.dim_y; p.dim_n = a.dim_n;
  p.ntaps = a.ntaps; p.kchunks = a.kchunks;
  for (int t = 0; t < a.ntaps; ++t) {
    for (int i = 0; i < 5; ++i) p.tap_o[t][i] = a.taps[t].o[i];
    p.tap_k[t] = a.tap_k[t];
  }
  p.b_batched = a.b_batched;
  p.BW = a.BW; p.BH = a.BH; p.BW_log2 = 0;
  while ((1 << p.BW_log2) < a.BW) ++p.BW_log2;
  p.tiles_x = ceil_div(a.Wo, a.BW); p.tiles_y = ceil_div(a.Ho, a.BH);
  p.n_tiles_m = p.tiles_x * p.tiles_y * a.Nimg;
  p.n_tiles_n = ceil_div(a.Cout, a.block_n);
  p.n_tiles = p.n_tiles_m * p.n_tiles_n;
  p.n_units = ceil_div(p.n_tiles_m, ncta) * p.n_tiles_n;
  p.fd_ntn = make_fastdiv(p.n_tiles_n); p.fd_tx = make_fastdiv(p.tiles_x); p.fd_ty = make_fastdiv(p.tiles_y);
  p.softmax_mode = a.softmax_mode;
  p.sm_per = 2 * p.n_tiles_n; p.n_groups = ceil_div(p.n_tiles_m, ncta); p.fd_per = make_fastdiv(p.sm_per);
  if (a.softmax_mode) {
    SFV_CHECK(a.out_16 && !a.out_f32 && !a.residual && !a.bias && !a.gn_stats && !a.relu && !a.a2 && !a.u8_src && a.block_n >= 32,
              "tc_gemm: softmax mode takes a 16-bit output only");
    p.n_units = p.n_groups * p.sm_per;
  }
  p.epi_mode = g_epi_mode;
  p.res_prefetch = g_res_prefetch;
  p.dbg = g_dbg;
  p.a2_kchunks = a.a2 ? a.a2_cin / 64 : 0; p.a2_k0 = a.a2_k0;
  p.halo_base_offset = g_halo_boff;
  p.Wo = a.Wo; p.Ho = a.Ho; p.Cout = a.Cout; p.n_img = a.Nimg;
  p.alpha = a.alpha; p.bias = a.bias; p.residual = (const float*)a.residual;
  p.res16 = a.residual ? a.res16 : 0; p.res_mul = a.res_mul == 0.f ? 1.f : a.res_mul;
  p.res_slot = p.res16 ? 2048 : 4096;
  SFV_CHECK(!p.res16 || !a.out_f32, "tc_gemm: a 16-bit residual goes with a 16-bit output only");
  SFV_CHECK(!p.res16 || a.block_n >= 32, "tc_gemm: 16-bit residual needs block_n >= 32");
  p.out_f32 = a.out_f32; p.out_16 = a.out_16; p.ldo = a.ldo; p.relu = a.relu;
  p.fmt_a = a.fmt; p.fmt_b = fmt_b; p.fmt_out = fmt_out;
  p.out16_scale = a.out16_scale == 0.f ? 1.f : a.out16_scale;
  p.bias_mul = 1.f; p.stat_mul = 1.f; p.sat_check = 0;
  SFV_CHECK(p.out16_scale == 1.f || !a.softmax_mode, "tc_gemm: softmax mode has no scaled 16-bit output");
  if (p.out16_scale != 1.f && a.out_16 && !a.out_f32 && a.block_n >= 32) {
    // the only output is the scaled 16-bit tensor: run the whole epilogue in the scaled domain
    const float sc = p.out16_scale;
    p.alpha *= sc; p.bias_mul = sc; p.res_mul *= sc; p.stat_mul = 1.f / sc;
    p.sat_check = fmt_out == FMT_F16;
    p.out16_scale = 1.f;
  }
  if (a.sat_check && a.out_16 && fmt_out == FMT_F16 && a.block_n >= 32) p.sat_check = 1;

  p.gn_stats = a.gn_stats; p.gn_cpg = a.gn_cpg; p.gn_groups = a.gn_cpg ? a.Cout / a.gn_cpg : 0;
  if (a.gn_stats)
    SFV_CHECK((a.gn_cpg == 4 || a.gn_cpg == 8 || a.gn_cpg == 16) && a.Cout % 32 == 0 && a.block_n >= 32 &&
                  a.block_n / a.gn_cpg <= 64,
              "tc_gemm: fused GroupNorm statistics need 4/8/16 channels per group (got %d)", a.gn_cpg);
  p.err = ds.err_flag;
  p.xf_stats = a.xf_stats; p.xf_gamma = a.xf_gamma; p.xf_beta = a.xf_beta; p.xf_hw = (long long)a.Ho * a.Wo;
  p.xf_cin = (int)a.a_dims[0]; p.xf_in_mul = a.xf_in_mul == 0.f ? 1.f : a.xf_in_mul; p.xf_silu = a.xf_silu; p.xf_check = a.xf_check;
  // epilogue tensor maps: per-warp boxes of 32 channels x (bx x by) pixels over the output / residual tensors
  CUtensorMap mr = ma, mo32 = ma, mo16 = ma;
  const bool tma_epi = g_epi_mode == 1 && a.block_n >= 32 && a.ldo % 8 == 0 &&
                       (!a.out_f32 || ((uintptr_t)a.out_f32 & 15) == 0) && (!a.out_16 || ((uintptr_t)a.out_16 & 15) == 0) &&
                       (!a.residual || ((uintptr_t)a.residual & 15) == 0);
  p.epi_mode = tma_epi ? 1 : 0;
  SFV_CHECK(tma_epi || a.Cout % 4 == 0, "tc_gemm: Cout %% 4 != 0 needs the TMA epilogue (ldo %% 8 == 0, aligned outputs)");
  SFV_CHECK(tma_epi || !a.softmax_mode, "tc_gemm: softmax mode needs the TMA epilogue");
  if (tma_epi) {
    const cuuint32_t bx = a.BW < 32 ? a.BW : 32, by = 32 / bx;
    cuuint64_t dims[4] = {(cuuint64_t)a.Cout, (cuuint64_t)a.Wo, (cuuint64_t)a.Ho, (cuuint64_t)a.Nimg};
    cuuint32_t box[4] = {32, bx, by, 1};
    cuuint64_t st32[4] = {4, (cuuint64_t)a.ldo * 4, (cuuint64_t)a.ldo * 4 * a.Wo, (cuuint64_t)a.ldo * 4 * a.Wo * a.Ho};
    cuuint64_t st16[4] = {2, (cuuint64_t)a.ldo * 2, (cuuint64_t)a.ldo * 2 * a.Wo, (cuuint64_t)a.ldo * 2 * a.Wo * a.Ho};
    if (a.residual && a.res16) SFV_TRY(encode_map(&mr, fmt_out, 4, a.residual, dims, st16, box, 64, false));
    else if (a.residual) SFV_TRY(encode_map(&mr, a.fmt, 4, a.residual, dims, st32, box, 128, true));
    if (a.out_f32) SFV_TRY(encode_map(&mo32, a.fmt, 4, a.out_f32, dims, st32, box, 128, true));
    if (a.out_16) SFV_TRY(encode_map(&mo16, fmt_out, 4, a.out_16, dims, st16, box, 64, false));
  }
  p.epi_fast = g_epi_fast && tma_epi && !a.softmax_mode && a.block_n == 128 && a.out_16 && !a.out_f32 && p.out16_scale == 1.f &&
               fmt_out == FMT_F16 && !a.relu && a.gn_stats && a.gn_cpg == 4 && (!a.residual || p.res16) &&
               a.Cout % a.block_n == 0 && p.n_tiles_m % ncta == 0;
  char tag[64];
  snprintf(tag, sizeof(tag), "M=%dx%dx%d N=%d K=%dx%d bn=%d cta=%d%s res=%d f32=%d o16=%d gn=%d", a.Nimg, a.Ho, a.Wo, a.Cout,
           a.ntaps, a.kchunks * 64, a.block_n, ncta, halo ? "h" : "", a.residual != nullptr, a.out_f32 != nullptr, a.out_16 != nullptr,
           a.gn_stats != nullptr);
  if (xf) {
    snprintf(tag + strlen(tag), sizeof(tag) - strlen(tag), " xf");
    return launch_cfg<256, 2, true, true>(ma, mb, ma2, mr, mo32, mo16, p, ds, s, tag);
  }
  if (halo && a.block_n == 256) return launch_cfg<256, 2, true>(ma, mb, ma2, mr, mo32, mo16, p, ds, s, tag);
  if (halo) return launch_cfg<128, 2, true>(ma, mb, ma2, mr, mo32, mo16, p, ds, s, tag);
  if (ncta == 2) {
    switch (a.block_n) {
      case 256: return launch_cfg<256, 2>(ma, mb, ma2, mr, mo32, mo16, p, ds, s, tag);
      case 128: return launch_cfg<128, 2>(ma, mb, ma2, mr, mo32, mo16, p, ds, s, tag);
      case 64: return launch_cfg<64, 2>(ma, mb, ma2, mr, mo32, mo16, p, ds, s, tag);
      case 32: return launch_cfg<32, 2>(ma, mb, ma2, mr, mo32, mo16, p, ds, s, tag);
      default: break;
    }
  }
  switch (a.block_n) {
    case 256: return launch_cfg<256, 1>(ma, mb, ma2, mr, mo32, mo16, p, ds, s, tag);
    case 128: return launch_cfg<128, 1>(ma, mb, ma2, mr, mo32, mo16, p, ds, s, tag);
    case 64: return launch_cfg<64, 1>(ma, mb, ma2, mr, mo32, mo16, p, ds, s, tag);
    case 32: return launch_cfg<32, 1>(ma, mb, ma2, mr, mo32, mo16, p, ds, s, tag);
    case 16: return launch_cfg<16, 1>(ma, mb, ma2, mr, mo32, mo16, p, ds, s, tag);
    default: return fail(SFV_ERR_INVALID, "tc_gemm: unsupported block_n %d", a.block_n);
  }
}

// Synchronise `s` and read (and clear) the current device's error word.
int tc_check_device_error(cudaStream_t s) {
  int dev = 0;
  SFV_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDevices || !g_dev[dev].err_flag) { SFV_CUDA(cudaStreamSynchronize(s)); return 0; }
  int* flag = g_dev[dev].err_flag;
  int h = 0;
  SFV_CUDA(cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, s));
  SFV_CUDA(cudaStreamSynchronize(s));
  if (h != 0) {
    cudaMemsetAsync(flag, 0, sizeof(int), s);
    if (h >= kErrRangeBase) {
      static const char* site[] = {"?", "a GroupNorm(+SiLU) output", "a GroupNorm input (conv1 output / 16-bit residual stream)",
                                   "a 16-bit store of the residual stream"};
      const int k = h - kErrRangeBase;
      return fail(SFV_ERR_RANGE, "fp16 operand range exceeded at %s: the results of this call are invalid; "
                                 "re-create the encoder with precision bf16", site[(k >= 1 && k <= 3) ? k : 0]);
    }
    return fail(SFV_ERR_DEVICE, "tcgen05 pipeline watchdog tripped (role code %d)", h);
  }
  return 0;
}

}  // namespace sfv
