"""pytest -m gpu: the parity tests proper (CUDA path through the C ABI vs the CPU oracle)."""
import pytest

pytestmark = pytest.mark.gpu

# Pure bf16 operands are a KNOWN MISS of the north-star latent gate on random-init weights (operand-rounding floor
# 0.9-1.6e-2 > 1e-2, oracle/numerics_model.py).  The product default is "mixed", which passes.
# The bf16 cases assert the same 1e-2 as every other 16-bit mode and are recorded as expected failures -- the gate
# is not widened to fit them.
BF16_MISS = pytest.param("bf16", marks=pytest.mark.xfail(
    reason="bf16 operand-rounding floor 0.9-1.6e-2 exceeds the 1e-2 latent gate; use precision='mixed'", strict=False))


def _c():
    import gpu_checks
    return gpu_checks


@pytest.mark.parametrize("shape", range(13))
@pytest.mark.parametrize("prec", ["fp32", "bf16", "fp16", "mixed"])
def test_conv_tile_geometries(prec, shape):
    c = _c()
    N, H, W, Cin, Cout, ks, st, pad, res = c.CONV_SHAPES[shape]
    c.check_conv(prec, N, H, W, Cin, Cout, ks, st, pad, residual=res, seed=shape)


@pytest.mark.parametrize("shape", range(12))
@pytest.mark.parametrize("prec", ["bf16", "fp32"])
def test_conv_layer_shapes(prec, shape):
    c = _c()
    N, H, W, Cin, Cout, ks, st, pad, res = c.LAYER_SHAPES_64[shape]
    c.check_conv(prec, N, H, W, Cin, Cout, ks, st, pad, residual=res, seed=100 + shape)


def test_conv_small_channels_fp32():
    c = _c()
    c.check_conv("fp32", 2, 17, 23, 3, 128, 3, 1, (1, 1))            # conv_in shape class (K = 27)
    c.check_conv("fp32", 2, 9, 11, 4, 256, 3, 2, (1, 1), relu=True)  # RBVAE conv.0 (odd sizes)
    c.check_conv("fp32", 1, 11, 20, 64, 64, 3, 2, (1, 1))


@pytest.mark.parametrize("C,HW", [(128, 4096), (256, 1000), (512, 77), (64, 333)])
def test_group_norm(C, HW):
    _c().check_group_norm(C, HW, silu=True)
    _c().check_group_norm(C, HW, silu=False, seed=1)


@pytest.mark.parametrize("prec,L", [("fp32", 256), ("bf16", 256), ("fp16", 1024), ("bf16", 144), ("bf16", 125), ("fp16", 77)])
def test_attention(prec, L):
    _c().check_attention(prec, N=2, L=L)


@pytest.mark.parametrize("prec", ["mixed", "fp16", "fp32"])
def test_shape_sweep(prec):
    print(_c().check_shape_sweep(prec))


@pytest.mark.parametrize("prec", ["mixed", "fp16"])
def test_token_count_not_multiple_of_8(prec):
    print(_c().check_odd_token_count(prec))


def test_resize_bit_exact_vs_pil_golden():
    _c().check_resize()


@pytest.mark.parametrize("name", ["kl_f8_seed0_2x64x96_white", "kl_f8_seed1_1x128x128_smooth",
                                  "kl_f8_seed0_2x256x256_white"])
@pytest.mark.parametrize("prec", ["fp32", "mixed", "fp16", BF16_MISS])
def test_encoder_vs_reference_golden(prec, name):
    print(_c().check_encoder_golden(prec, name))


@pytest.mark.parametrize("prec", ["fp32", "mixed", BF16_MISS])
def test_encoder_layerwise(prec):
    print(_c().check_encoder_taps(prec))


def test_chunking_and_batch_independence():
    _c().check_chunking_and_batch_independence("fp32")
    _c().check_chunking_and_batch_independence("mixed")
    _c().check_chunking_and_batch_independence("bf16")


@pytest.mark.parametrize("name", ["rbvae_percep_L25_32x32_T1", "rbvae_percep_L25_64x64_T1",
                                  "rbvae_percep_L100_88x160_T1", "rbvae_percep_L50_32x32_T4",
                                  "rbvae_contrastive_L25_256x256_T1"])
def test_rbvae_vs_reference_golden(name):
    print(_c().check_rbvae_golden(name))


@pytest.mark.parametrize("prec", ["bf16", "fp16", "mixed"])
@pytest.mark.parametrize("name", ["rbvae_percep_L25_64x64_T1", "rbvae_percep_L100_88x160_T1",
                                  "rbvae_contrastive_L25_256x256_T1"])
def test_rbvae_tensor_core_mode(name, prec):
    print(_c().check_rbvae_tensor_core(name, prec))


def test_hamming():
    _c().check_hamming()


@pytest.mark.parametrize("prec", ["fp32", "mixed", "fp16", BF16_MISS])
def test_pipeline_frames_to_codes(prec):
    print(_c().check_pipeline(prec))


def test_full_size_properties_512():
    print(_c().check_full_size_properties("mixed", B=4, R=512))


@pytest.mark.parametrize("prec", ["mixed", "fp16", BF16_MISS])
def test_native_1280x704_frame(prec):
    print(_c().check_native_frame_size(prec))


def test_large_frame_properties_1024():
    print(_c().check_large_frame_properties("mixed", 1024, 2))


@pytest.mark.parametrize("prec", ["fp32", "bf16", "mixed"])
def test_contrastive_rbvae_512(prec):
    print(_c().check_contrastive_512(prec))


@pytest.mark.parametrize("prec", ["fp32", "mixed", "fp16", BF16_MISS])
def test_chinchess_480_frame_code_match(prec):
    print(_c().check_chinchess_video(prec))


@pytest.mark.parametrize("prec", ["fp16", "bf16", "mixed"])
def test_conv_in_tensor_core(prec):
    print(_c().check_conv_in_tensor_core(prec))


def test_evaluation_kernels_bit_exact():
    print(_c().check_evaluation_kernels())


@pytest.mark.parametrize("prec", ["fp32", "mixed", "bf16"])
def test_state_consistency_pipeline(prec):
    print(_c().check_state_consistency_pipeline(prec))


def test_edge_cases_and_errors():
    print(_c().check_edge_cases())


def test_mixed_mode_never_saturates_silently():
    print(_c().check_range_safety())


@pytest.mark.parametrize("prec", ["mixed", "fp32"])
def test_precompute_driver_video_to_reference_npy(prec):
    from oracle import ref_shim
    if ref_shim.video_path() is None:
        pytest.skip("the reference's sample video is not present (oracle/_ref/videos, placed by __graft_entry__.build())")
    print(_c().check_precompute_driver(prec))


def test_resident_dataset_on_device():
    print(_c().check_resident_dataset())


@pytest.mark.parametrize("prec,R,B", [("mixed", 512, 4), ("fp32", 512, 1), ("mixed", 1024, 2), ("fp32", 1024, 1),
                                      pytest.param("bf16", 512, 4, marks=BF16_MISS.marks)])
def test_full_size_against_cuda_fp32_oracle(prec, R, B):
    print(_c().check_full_size_oracle(prec, R, B))


@pytest.mark.parametrize("prec", ["fp32", "mixed"])
def test_encode_host_posterior_sampling(prec):
    print(_c().check_encode_host_sampling(prec))


def test_fused_groupnorm_transform_matches_default_path():
    print(_c().check_gn_fused_transform())


# ---- decoder half + losses: the training-side forward ----
@pytest.mark.parametrize("name", ["rbvae_forward_percep_L25_88x160", "rbvae_forward_contrastive_L25_256x256"])
def test_rbvae_forward_with_decoder_golden(name):
    print(_c().check_decoder_golden(name))


def test_rbvae_decoder_other_shapes_and_errors():
    print(_c().check_decoder_shapes())


def test_training_losses():
    print(_c().check_losses())
