"""symbols-from-video_b200 -- B200-native (sm_100a) frame -> KL-f8 latent -> binary code.

The directory name carries a hyphen (it mirrors the reference repo's name), so
import it through the ``sfv_b200`` alias module at the repo root::

    import sfv_b200
    vae = sfv_b200.AutoencoderKL(precision="bf16").cuda()
    post = vae.encode(x)                       # reference call surface
    rb = sfv_b200.Seq2SeqBinaryVAE(4, 4, latent_dim=25, input_hw=(64, 64))
    z = rb.encode(lat[:, None], hard=True, noise_ratio=0.0)

Everything numerical runs in ``libsfv.so`` (hand-written CUDA behind the C ABI in
``include/sfv.h``); importing this package never imports ``oracle/`` and there
is no CPU / eager-PyTorch fallback.
"""
from . import _lib
from ._lib import SfvError, lib
from .autoencoder import (SCALE_FACTOR, KL_F8_DDCONFIG, AutoencoderKL, DiagonalGaussianDistribution,
                          FirstStage, encoder_param_shapes)
from .rbvae import Seq2SeqBinaryVAE, hamming_matrix, unpack_codes
from .pipeline import (EncodeResult, FramePipeline, all_gather_ragged, all_gather_slices, encode_sharded,
                       shard_range)
from .embedding_store import (FlatEmbeddingStore, ShuffledStatePairDataset, frame_key, load_embeddings_npy,
                              lookup_embedding, save_embeddings_npy)
from . import evaluation, feeder, losses, ops, precompute
from .feeder import ArraySource, Feeder, FrameDirSource, VideoSource
from .precompute import PrecomputeResult, precompute_embeddings
from .evaluation import (add_gaussian_noise, add_occlusion, assign_label, calculate_state_consistency,
                         labels_from_flags, perturb_frames, state_consistency)
from .weights import (init_encoder_state_dict, init_rbvae_decoder_state_dict, init_rbvae_state_dict,
                      make_rbvae_responsive, synthetic_frames)

__all__ = [n for n in dir() if not n.startswith("_")]
