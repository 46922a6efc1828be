/* sfv.h -- C ABI of the B200-native frame -> KL-f8 latent -> binary-code path.
 *
 * This is the drop-in boundary.  The reference
 * (matt-suncy/symbols-from-video) is pure Python on PyTorch and has no FFI of
 * its own; each entry point below names the reference Python interface it
 * replaces (path:line under /root/reference).  INTEGRATION.md shows the ctypes
 * stub a reference maintainer would add.
 *
 * Conventions
 *  - plain C types only; every pointer is a DEVICE pointer unless its name
 *    starts with host_;  the caller owns all buffers and the CUDA stream.
 *  - every function returns 0 on success, a negative SfvStatus otherwise;
 *    sfv_last_error() returns a thread-local message for the last failure.
 *  - calls are asynchronous on the given stream (cudaStream_t passed as void*);
 *    no hidden allocation or synchronisation after *_create.
 *  - there is no CPU fallback: without a CUDA device every compute call fails
 *    with SFV_ERR_CUDA.
 *  - activations inside the library are NHWC; tensors at the boundary keep the
 *    reference's NCHW fp32 layout.
 */
#ifndef SFV_H_
#define SFV_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum SfvStatus {
  SFV_OK = 0,
  SFV_ERR_INVALID = -1,    /* bad argument / unsupported shape            */
  SFV_ERR_CUDA = -2,       /* CUDA runtime / driver error                  */
  SFV_ERR_MISSING_KEY = -3,/* a state-dict tensor is absent or mis-shaped  */
  SFV_ERR_WORKSPACE = -4,  /* workspace too small                          */
  SFV_ERR_DEVICE = -5,     /* device-side watchdog tripped (pipeline hang) */
  SFV_ERR_RANGE = -6       /* an fp16 operand left the fp16 range (MIXED / FP16 modes): results are invalid,
                              re-create the handle with SFV_PREC_BF16 */
} SfvStatus;

/* Arithmetic mode of the encoder / GEMM operands.
 *  F32   : "fp32 check mode" -- CUDA-core fp32 implicit GEMM everywhere.
 *  BF16  : tcgen05 kind::f16, bf16 operands, fp32 TMEM accumulators,
 *          fp32 residual stream and GroupNorm statistics.
 *  FP16  : same kernels with IEEE half operands (same tensor-pipe rate,
 *          3 more mantissa bits).  Diagnostic mode: conv outputs consumed directly
 *          as operands (q, k, V, raw copies of x) saturate at +-65504.
 *  MIXED : the product default.  fp16 where the operand is bounded by construction --
 *          weights (per-layer power-of-two scale, undone in the epilogue), GroupNorm(+SiLU)
 *          outputs, conv1 outputs (re-normalised by norm2) and the 16-bit copies of the
 *          residual stream that feed Downsample / nin_shortcut (stored times 2^-6: range
 *          +-4.19e6) -- and bf16 where range is data dependent (q, k, V^T, P, attention
 *          output, and proj_out's weights: a tcgen05 GEMM takes both operands in one format).  Every fp16 store site is range-checked on the device: a value that
 *          leaves the fp16 range raises SFV_ERR_RANGE at the next sfv_check_async_error /
 *          synchronising call instead of saturating silently. */
typedef enum SfvPrecision { SFV_PREC_F32 = 0, SFV_PREC_BF16 = 1, SFV_PREC_FP16 = 2, SFV_PREC_MIXED = 3 } SfvPrecision;

/* One named host tensor of a PyTorch state-dict (fp32, contiguous, reference
 * layout: conv weights OIHW, linear [out,in], LSTM [4L,L]). */
typedef struct SfvTensor {
  const char* name;        /* reference key, e.g. "encoder.down.0.block.0.conv1.weight" */
  const float* host_data;
  int32_t ndim;
  int64_t shape[4];
} SfvTensor;

typedef struct SfvEncoder SfvEncoder;   /* AutoencoderKL.encoder + quant_conv */
typedef struct SfvRbvae SfvRbvae;       /* Seq2SeqBinaryVAE encoder half       */
typedef struct SfvRbvaeDecoder SfvRbvaeDecoder;   /* decoder_rnn + ConvDecoder (training-side forward) */

const char* sfv_version(void);
const char* sfv_last_error(void);
/* 1 if a CUDA device with compute capability 10.x is usable, else 0. */
int sfv_device_ok(void);

/* ---- KL-f8 encoder --------------------------------------------------------
 * Replaces AutoencoderKL.__init__/init_from_ckpt weight ingestion
 * (src/stable-diffusion/ldm/models/autoencoder.py:285-322): `tensors` holds the
 * "encoder.*" and "quant_conv.*" entries of the state-dict (kl-f8 ddconfig,
 * configs/stable-diffusion/v1-inference.yaml:46-67).  Weights are repacked on
 * the device at creation (K-major 16-bit tiles for UMMA, tap-major fp32 for the
 * check mode). */
int sfv_encoder_create(const SfvTensor* tensors, int32_t n_tensors, int32_t precision,
                       SfvEncoder** out);
void sfv_encoder_destroy(SfvEncoder* enc);
int sfv_encoder_precision(const SfvEncoder* enc);
/* Frames pushed through the network together (default 16): bounds the workspace
 * and keeps the deeper levels L2-resident; B larger than this is processed in
 * slices inside one sfv_encoder_forward_* call. */
int sfv_encoder_set_chunk(SfvEncoder* enc, int32_t frames);
/* Synchronises `stream` and returns SFV_ERR_DEVICE if a tcgen05 pipeline
 * watchdog tripped in any kernel launched so far on the current device (bounded
 * mbarrier waits turn a would-be hang into an error), or SFV_ERR_RANGE if an fp16
 * operand store left the fp16 range (MIXED / FP16 modes).  The flag is cleared. */
int sfv_check_async_error(void* stream);

/* Bytes of scratch sfv_encoder_forward_* needs for a batch of B frames HxW. */
int sfv_encoder_workspace_bytes(const SfvEncoder* enc, int32_t B, int32_t H, int32_t W,
                                size_t* bytes);

/* Replaces AutoencoderKL.encode(x) (autoencoder.py:324-328) +
 * DiagonalGaussianDistribution.__init__ (ldm/modules/distributions/distributions.py:24-33):
 * x fp32 NCHW [B,3,H,W] in [-1,1]  ->
 *   parameters fp32 NCHW [B,8,H/8,W/8]  raw moments; mean = channels 0..3 (a view,
 *                                        exactly as torch.chunk gives the reference)
 *   logvar     fp32 NCHW [B,4,H/8,W/8]  clamp(parameters[:,4:], -30, 20)
 *   std, var   fp32 NCHW [B,4,H/8,W/8]  exp(0.5 logvar), exp(logvar)   (optional, may be NULL)
 * H, W multiples of 8 (tensor-core modes additionally need (H/8)*(W/8) % 8 == 0).
 * host_taps_or_null: optional HOST array of SFV_NUM_TAPS device pointers; a non-null
 * entry receives that block's output as fp32 NHWC (layer-wise parity bisection). */
int sfv_encoder_forward_nchw(SfvEncoder* enc, const float* x, int32_t B, int32_t H, int32_t W,
                             float* parameters, float* logvar, float* std_or_null, float* var_or_null,
                             void* workspace, size_t workspace_bytes,
                             float* const* host_taps_or_null, void* stream);

/* Same, fed from uint8 HWC RGB frames [B,H,W,3] (device memory, or pinned host
 * memory mapped into the device address space): fuses load_img's /255 and 2x-1
 * (src/stable-diffusion/get_percep_embeddings.py:67-71) into conv_in's gather. */
int sfv_encoder_forward_u8(SfvEncoder* enc, const uint8_t* frames, int32_t B, int32_t H, int32_t W,
                           float* parameters, float* logvar, float* std_or_null, float* var_or_null,
                           void* workspace, size_t workspace_bytes, void* stream);

#define SFV_NUM_TAPS 16
/* tap order: conv_in, down.{0..3}.block.{0,1} (8), down.{0,1,2}.downsample (3),
 * mid.block_1, mid.attn_1, mid.block_2, moments(pre-clamp, 8ch) */

/* LatentDiffusion.get_first_stage_encoding (ldm/models/diffusion/ddpm.py:542-549)
 * + DiagonalGaussianDistribution.sample (distributions.py:35-37):
 * out = scale * (mean + exp(0.5*logvar) * noise);  noise_or_null == NULL gives
 * scale * mean (mode()).  n = element count. */
int sfv_posterior_sample(const float* mean, const float* logvar, const float* noise_or_null,
                         float scale, float* out, int64_t n, void* stream);

/* uint8 HWC frame resize + normalise (replaces load_img,
 * get_percep_embeddings.py:48-71 / load_img_for_sd, embedding_matching.py:318-338):
 * frames [B,Hs,Ws,3] u8 -> out fp32 NCHW [B,3,H,W] = 2*(resized/255)-1.
 * Resampling is PIL's LANCZOS (a=3) with its 8-bit fixed-point two-pass
 * arithmetic, so the u8 result matches Image.resize bit for bit;
 * out_u8_or_null (HWC [B,H,W,3]) receives the resized uint8 frame. */
int sfv_resize_normalise(const uint8_t* frames, int32_t B, int32_t Hs, int32_t Ws,
                         int32_t H, int32_t W, float* out_nchw_or_null,
                         uint8_t* out_u8_or_null, void* workspace, size_t workspace_bytes,
                         void* stream);
int sfv_resize_workspace_bytes(int32_t B, int32_t Hs, int32_t Ws, int32_t H, int32_t W,
                               size_t* bytes);

/* ---- RBVAE encoder half ---------------------------------------------------
 * Replaces Seq2SeqBinaryVAE.__init__/load_state_dict for the encoder half
 * (models/percep_RBVAE/percep_RBVAE_model.py:135-141; contrastive:
 * models/contrastive_RBVAE/contrastive_RBVAE_model.py:134-140).  `tensors`:
 * encoder_cnn.conv.{0,3,6}.*, encoder_cnn.fc.*, encoder_rnn.lstm.*_l{k}.
 * Channel count (256 / 64) and LSTM depth (4 / 2) are read off the tensors;
 * fc.in_features fixes the accepted input H x W (in_h, in_w must be given). */
int sfv_rbvae_create(const SfvTensor* tensors, int32_t n_tensors, int32_t in_channels,
                     int32_t in_h, int32_t in_w, SfvRbvae** out);
/* Same with an arithmetic mode: SFV_PREC_F32 (default of sfv_rbvae_create: every layer fp32, codes
 * bit-exact outside the |h|<1e-3 band) or BF16/FP16 (the two C->C stride-2 convolutions, 97 % of the
 * RBVAE FLOPs, run on the tcgen05 kernel with 16-bit operands; conv.0, fc and the LSTM stay fp32). */
int sfv_rbvae_create_ex(const SfvTensor* tensors, int32_t n_tensors, int32_t in_channels,
                        int32_t in_h, int32_t in_w, int32_t precision, SfvRbvae** out);
void sfv_rbvae_destroy(SfvRbvae* rb);
int sfv_rbvae_latent_dim(const SfvRbvae* rb);
int sfv_rbvae_workspace_bytes(const SfvRbvae* rb, int32_t N, size_t* bytes);

/* Replaces Seq2SeqBinaryVAE.encode / the encoder half of .forward
 * (percep_RBVAE_model.py:143-160,172-191) and binary_concrete_logits (:17-44).
 *  x        fp32 [B,T,C,h,w] (NCHW per frame), pre-multiplied by in_scale
 *           (lets the caller pass the raw posterior mean with in_scale=0.18215)
 *  u_or_null  uniform(0,1) draws [B*T,L] standing in for the reference's
 *           torch.rand (required when noise_ratio != 0)
 *  h_out    fp32 [B,T,L] LSTM hidden state (the thresholded "logit"), or NULL
 *  z_out    fp32 [B,T,L] binary-concrete output (soft, or {0,1} if hard), or NULL
 *  codes_out uint32 [B*T, ceil(L/32)] bit-packed hard code (bit j of word w =
 *           (h+noise > 0) for latent 32w+j), or NULL */
int sfv_rbvae_encode(SfvRbvae* rb, const float* x, int32_t B, int32_t T, float in_scale,
                     const float* u_or_null, float noise_ratio, float temperature, int32_t hard,
                     float* h_out, float* z_out, uint32_t* codes_out,
                     void* workspace, size_t workspace_bytes, void* stream);

/* ---- decoder half and losses: the training-side forward, fp32, forward values only -------------
 * Replaces DecoderRNN.forward + ConvDecoder.forward, i.e. the second half of Seq2SeqBinaryVAE.forward
 * (models/percep_RBVAE/percep_RBVAE_model.py:71-91,110-122,159-168;
 *  models/contrastive_RBVAE/contrastive_RBVAE_model.py:70-90,109-121).  Tensors: the reference keys
 * decoder_cnn.fc.{weight,bias}, decoder_cnn.deconv.{0,3,6}.{weight,bias} (ConvTranspose2d layout [Cin,Cout,3,3]),
 * decoder_rnn.lstm.{weight_ih,weight_hh,bias_ih,bias_hh}_l{k}.  out_h x out_w is the reconstruction's size
 * (a multiple of 8; fc.out_features must equal channels * out_h/8 * out_w/8 -- the reference hard-wires 11x20). */
int sfv_rbvae_decoder_create(const SfvTensor* tensors, int32_t n_tensors, int32_t out_channels, int32_t out_h,
                             int32_t out_w, SfvRbvaeDecoder** out);
void sfv_rbvae_decoder_destroy(SfvRbvaeDecoder* d);
int sfv_rbvae_decoder_workspace_bytes(const SfvRbvaeDecoder* d, int32_t N, size_t* bytes);
/*  z_seq    fp32 [B,T,L]  binary-concrete output of the encoder half
 *  d_seq_out fp32 [B,T,L] decoder LSTM output, or NULL
 *  x_recon  fp32 [B,T,C,out_h,out_w] (NCHW per frame), values in (0,1) */
int sfv_rbvae_decode(SfvRbvaeDecoder* d, const float* z_seq, int32_t B, int32_t T, float* d_seq_out, float* x_recon,
                     void* workspace, size_t workspace_bytes, void* stream);

/* Losses of models/percep_RBVAE/percep_RBVAE_train.py:27-107 (the contrastive trainer uses the same functions).
 * Every result is ONE fp32 value written to device memory `out`.
 *  sfv_loss_mse       recon_loss = F.mse_loss(a, b)                                    (:32-33)
 *  sfv_loss_l1        l1_loss = lamb * torch.norm(q, p=1)                              (:27-29)
 *  sfv_loss_kl_binary_concrete  KL(Bernoulli(sigmoid(q)) || Bernoulli(p)), summed over the L latents, mean over rows (:52-77)
 *  sfv_loss_contrast  contrast_loss(x1, x2, label, margin, dist): cosine != 0 -> 1 - cosine_similarity, else
 *                     F.pairwise_distance                                              (:80-107)
 *  sfv_loss_triplet   F.triplet_margin_loss(a, p, n, margin, p=2, eps, swap, 'mean')   (:35-49) */
int sfv_loss_mse(const float* a, const float* b, int64_t n, float* out, void* stream);
int sfv_loss_l1(const float* q, int64_t n, float lamb, float* out, void* stream);
int sfv_loss_kl_binary_concrete(const float* q_logits, int64_t rows, int32_t L, float p, float eps, float* out, void* stream);
int sfv_loss_contrast(const float* x1, const float* x2, const float* label, int32_t rows, int32_t D, float margin,
                      int32_t cosine, float* out, void* stream);
int sfv_loss_triplet(const float* anchor, const float* pos, const float* neg, int32_t rows, int32_t D, float margin,
                     float eps, int32_t swap, float* out, void* stream);

/* Hamming distance matrix between packed codes (evaluation helper for
 * scripts/evaluation/clustering_eval/embedding_hamming_distance.py:53-87):
 * a [Na,words], b [Nb,words] -> out int32 [Na,Nb]. */
int sfv_hamming(const uint32_t* a, int32_t Na, const uint32_t* b, int32_t Nb, int32_t words,
                int32_t* out, void* stream);

/* State consistency on packed codes
 * (scripts/evaluation/state_consistency_eval/embedding_matching.py:275-297; the same
 * block in models/percep_RBVAE/percep_RBVAE_train.py:473-497): for every state s,
 * best_count[s] = number of frames labelled s whose code equals the state's most common
 * code, state_count[s] = frames labelled s.  percentage_s = best/state, weighted average =
 * sum(best)/sum(state).  codes uint32 [n,words], labels int32 [n] in [0,n_states)
 * (labels outside the range are ignored), words <= 8. */
int sfv_state_consistency(const uint32_t* codes, const int32_t* labels, int64_t n, int32_t words,
                          int32_t n_states, int32_t* best_count, int32_t* state_count, void* stream);

/* Robustness perturbations on uint8 HWC frames (embedding_matching.py:141-193 followed by
 * T.ToPILImage() at :243): out = byte(255 * clamp(u8/255 + (noise*std + mean), 0, 1)) when
 * noise != NULL (noise fp32 [B,3,H,W], the CHW layout randn_like(ToTensor(img)) draws in);
 * then, when occ_size > 0 and occ_xy != NULL (device int32 [B,2] = (x,y) per frame, the
 * reference draws a fresh position per frame), the square [y,y+occ_size) x [x,x+occ_size) is
 * set to 127 (= byte(0.5*255)).  in == out is allowed. */
int sfv_perturb_frames(const uint8_t* frames, uint8_t* out, int32_t B, int32_t H, int32_t W,
                       const float* noise_or_null, float mean, float std,
                       const int32_t* occ_xy_or_null, int32_t occ_size, void* stream);

/* ---- single-operator entry points (parity bisection of the kernels) --------
 * All tensors NHWC; `precision` picks the CUDA-core fp32 kernel or the tcgen05
 * kernel with bf16/fp16 operands (x and w are given in fp32 and converted by a
 * library kernel first, so the operator is tested exactly as the encoder uses it). */
int sfv_op_conv2d(const float* x_nhwc, const float* host_w_oihw, const float* host_bias,
                  const float* residual_or_null, float* y_nhwc,
                  int32_t N, int32_t H, int32_t W, int32_t Cin, int32_t Cout,
                  int32_t ksize, int32_t stride, int32_t pad_lo, int32_t pad_hi,
                  int32_t relu, int32_t precision, void* stream);
/* conv_in (model.py:383-387,439) fed from uint8 HWC frames with load_img's /255, 2x-1 fused
 * (get_percep_embeddings.py:67-71): frames uint8 [N,H,W,3] -> y fp32 NHWC [N,H,W,128].
 * precision 0: CUDA-core kernel; 1/2: the tcgen05 kernel with the exact integer operand 2u-255. */
int sfv_op_conv_in_u8(const uint8_t* frames, const float* host_w_oihw, const float* host_bias, float* y_nhwc,
                      int32_t N, int32_t H, int32_t W, int32_t precision, void* stream);
int sfv_op_group_norm(const float* x_nhwc, const float* gamma, const float* beta, float* y_nhwc,
                      int32_t N, int32_t HW, int32_t C, int32_t groups, float eps, int32_t silu,
                      void* stream);
/* q,k,v fp32 [N,L,C]; out fp32 [N,L,C] = softmax(q k^T * scale) v */
int sfv_op_attention(const float* q, const float* k, const float* v, float* out,
                     int32_t N, int32_t L, int32_t C, float scale, int32_t precision, void* stream);

/* Per-kernel-class device timing for the roofline report (bench.py): while enabled
 * every launch of a class is bracketed by CUDA events on its own stream.
 * category: 0 tcgen05 GEMM/conv (work = algorithmic FLOPs), 1 CUDA-core igemm (FLOPs),
 * 2 GroupNorm statistics, 3 GroupNorm apply, 4 softmax (work = algorithmic bytes), 5 other.
 * sfv_profile_enable(on) resets the counters; sfv_profile_read synchronises the
 * recorded events and returns the accumulated milliseconds, work and launch count. */
int sfv_profile_enable(int32_t on);
int sfv_profile_read(int32_t category, double* ms, double* work, int64_t* launches);
/* Per-launch records since sfv_profile_enable(1): lines "category,ms,work,shape tag". */
const char* sfv_profile_log(void);

/* Number of kernel launches issued by this library since load (bench evidence). */
int64_t sfv_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SFV_H_ */
