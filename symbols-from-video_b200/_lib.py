"""ctypes binding of libsfv.so (C ABI: include/sfv.h).

There is deliberately no fallback: if the shared library is missing, or no
sm_100 device is present, every compute entry point raises.  Build with
``python -c "import __graft_entry__ as g; g.build()"`` (or ``make -C
symbols-from-video_b200/csrc``).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SFV_LIB_PATH") or os.path.join(_HERE, "libsfv.so")   # override: instrumented debug builds

PREC_F32, PREC_BF16, PREC_FP16, PREC_MIXED = 0, 1, 2, 3
PRECISIONS = {"fp32": PREC_F32, "f32": PREC_F32, "bf16": PREC_BF16, "fp16": PREC_FP16, "f16": PREC_FP16,
              "mixed": PREC_MIXED}
DEFAULT_PRECISION = "mixed"      # fp16 for bounded operands, bf16 for attention q/k/V/P/O (include/sfv.h SfvPrecision)
NUM_TAPS = 16
TAP_NAMES = (["conv_in"] + [f"down.{l}.block.{b}" for l in range(4) for b in range(2)]
             + [f"down.{l}.downsample" for l in range(3)] + ["mid.block_1", "mid.attn_1", "mid.block_2", "moments"])
TAP_CHANNELS = [128, 128, 128, 256, 256, 512, 512, 512, 512, 128, 256, 512, 512, 512, 512, 8]
TAP_DOWN = [0, 0, 0, 1, 1, 2, 2, 3, 3, 1, 2, 3, 3, 3, 3, 3]   # log2 of the spatial reduction


class SfvTensor(C.Structure):
    _fields_ = [("name", C.c_char_p), ("host_data", C.POINTER(C.c_float)),
                ("ndim", C.c_int32), ("shape", C.c_int64 * 4)]


class SfvError(RuntimeError):
    pass


_lib = None

# name -> (restype, argtypes); every symbol include/sfv.h declares
_P = C.c_void_p
SIGNATURES = {
    "sfv_version": (C.c_char_p, []),
    "sfv_last_error": (C.c_char_p, []),
    "sfv_device_ok": (C.c_int, []),
    "sfv_launch_count": (C.c_int64, []),
    "sfv_profile_enable": (C.c_int, [C.c_int32]),
    "sfv_profile_read": (C.c_int, [C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "sfv_profile_log": (C.c_char_p, []),
    "sfv_encoder_create": (C.c_int, [C.POINTER(SfvTensor), C.c_int32, C.c_int32, C.POINTER(_P)]),
    "sfv_encoder_destroy": (None, [_P]),
    "sfv_encoder_precision": (C.c_int, [_P]),
    "sfv_encoder_set_chunk": (C.c_int, [_P, C.c_int32]),
    "sfv_check_async_error": (C.c_int, [_P]),
    "sfv_encoder_workspace_bytes": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_size_t)]),
    "sfv_encoder_forward_nchw": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P, _P,
                                           C.c_size_t, C.POINTER(_P), _P]),
    "sfv_encoder_forward_u8": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P, _P,
                                         C.c_size_t, _P]),
    "sfv_posterior_sample": (C.c_int, [_P, _P, _P, C.c_float, _P, C.c_int64, _P]),
    "sfv_resize_normalise": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P,
                                       C.c_size_t, _P]),
    "sfv_resize_workspace_bytes": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                             C.POINTER(C.c_size_t)]),
    "sfv_rbvae_create": (C.c_int, [C.POINTER(SfvTensor), C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                   C.POINTER(_P)]),
    "sfv_rbvae_create_ex": (C.c_int, [C.POINTER(SfvTensor), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                      C.POINTER(_P)]),
    "sfv_rbvae_destroy": (None, [_P]),
    "sfv_rbvae_latent_dim": (C.c_int, [_P]),
    "sfv_rbvae_workspace_bytes": (C.c_int, [_P, C.c_int32, C.POINTER(C.c_size_t)]),
    "sfv_rbvae_encode": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_float, _P, C.c_float, C.c_float, C.c_int32,
                                   _P, _P, _P, _P, C.c_size_t, _P]),
    "sfv_hamming": (C.c_int, [_P, C.c_int32, _P, C.c_int32, C.c_int32, _P, _P]),
    "sfv_rbvae_decoder_create": (C.c_int, [C.POINTER(SfvTensor), C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                           C.POINTER(_P)]),
    "sfv_rbvae_decoder_destroy": (None, [_P]),
    "sfv_rbvae_decoder_workspace_bytes": (C.c_int, [_P, C.c_int32, C.POINTER(C.c_size_t)]),
    "sfv_rbvae_decode": (C.c_int, [_P, _P, C.c_int32, C.c_int32, _P, _P, _P, C.c_size_t, _P]),
    "sfv_loss_mse": (C.c_int, [_P, _P, C.c_int64, _P, _P]),
    "sfv_loss_l1": (C.c_int, [_P, C.c_int64, C.c_float, _P, _P]),
    "sfv_loss_kl_binary_concrete": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_float, C.c_float, _P, _P]),
    "sfv_loss_contrast": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, C.c_float, C.c_int32, _P, _P]),
    "sfv_loss_triplet": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_int32, _P, _P]),
    "sfv_state_consistency": (C.c_int, [_P, _P, C.c_int64, C.c_int32, C.c_int32, _P, _P, _P]),
    "sfv_perturb_frames": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, _P, C.c_float, C.c_float,
                                     _P, C.c_int32, _P]),
    "sfv_op_conv2d": (C.c_int, [_P, _P, _P, _P, _P] + [C.c_int32] * 11 + [_P]),
    "sfv_op_conv_in_u8": (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P]),
    "sfv_op_group_norm": (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float,
                                    C.c_int32, _P]),
    "sfv_op_attention": (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_int32, _P]),
}


def lib():
    """Load libsfv.so once; raise (never fall back) if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SfvError(f"{LIB_PATH} not found: build it with __graft_entry__.build() "
                           "(there is no CPU / PyTorch fallback for this path)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)          # AttributeError if the ABI lost a symbol
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(status: int):
    if status != 0:
        raise SfvError(f"libsfv error {status}: {lib().sfv_last_error().decode()}")


def require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise SfvError(f"{what} must be a CUDA tensor: this path has no CPU fallback")


def stream_ptr(device=None) -> int:
    """The caller's current stream ON `device` (a tensor's device), not on whatever device happens to be current."""
    return torch.cuda.current_stream(device).cuda_stream


def on_device(t: torch.Tensor):
    """Context manager making `t`'s device current: libsfv handles, kernel attributes and the error word are
    per device, and the C side resolves "the device" through the CUDA current device."""
    return torch.cuda.device(t.device)


def run(fn, dev_tensor: torch.Tensor, *args):
    """Call a libsfv entry point whose last parameter is the stream: `dev_tensor`'s device is made current
    (handles, kernel attributes and the error word are per device) and ITS current stream is passed."""
    dev = dev_tensor.device
    with torch.cuda.device(dev):
        check(fn(*args, stream_ptr(dev)))


def check_async_error(device=None):
    """Synchronise the current stream of `device` and raise SfvError if a device-side check fired since the last
    call: a tcgen05 pipeline watchdog (outputs are garbage) or an fp16 range violation (mixed / fp16 modes).
    Product paths call this wherever they already synchronise."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    with torch.cuda.device(dev):
        check(lib().sfv_check_async_error(stream_ptr(dev)))


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def make_tensor_table(sd: dict, prefixes=None):
    """state-dict -> (SfvTensor array, keep-alive list).  Values are copied to
    contiguous fp32 host arrays in the reference layout."""
    items = [(k, v) for k, v in sd.items()
             if prefixes is None or any(k.startswith(p) for p in prefixes)]
    arr = (SfvTensor * max(len(items), 1))()
    keep = []
    for i, (k, v) in enumerate(items):
        a = np.ascontiguousarray(v.detach().to("cpu", torch.float32).numpy())
        if a.ndim > 4:
            raise SfvError(f"{k}: rank {a.ndim} > 4")
        name = k.encode()
        keep.append((a, name))
        arr[i].name = name
        arr[i].host_data = a.ctypes.data_as(C.POINTER(C.c_float))
        arr[i].ndim = a.ndim
        for d in range(4):
            arr[i].shape[d] = a.shape[d] if d < a.ndim else 1
    return arr, len(items), keep


class Workspace:
    """Grow-only device scratch owned by the Python side (the C ABI never allocates per call)."""

    def __init__(self):
        self.buf = None

    def get(self, nbytes: int, device) -> torch.Tensor:
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != device:
            self.buf = None
            self.buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        return self.buf
