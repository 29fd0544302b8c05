"""ctypes binding of libcgat_b200.so (the C ABI in include/cgat_b200.h).

There is no fallback: if the library is missing, or a kernel rejects its arguments, a RuntimeError is
raised.  Tensors are passed as raw ``data_ptr()`` values; the stream is torch's current stream.
"""
from __future__ import annotations

import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CGAT_B200_LIB", os.path.join(os.path.dirname(_HERE), "libcgat_b200.so"))

F32, BF16 = 0, 1
LAYOUT_SPATIAL, LAYOUT_TEMPORAL = 0, 1
PROJ_LINEAR, PROJ_PRE = 0, 1
MERGE_CONCAT, MERGE_MEAN = 0, 1
ACT_NONE, ACT_RELU, ACT_LRELU, ACT_SIGMOID = 0, 1, 2, 3


class AttnDesc(ctypes.Structure):
    _fields_ = [
        ("n_pix", ctypes.c_int64),
        ("pix_per_sample", ctypes.c_int64),
        ("nodes", ctypes.c_int32),
        ("ci", ctypes.c_int32),
        ("co", ctypes.c_int32),
        ("heads", ctypes.c_int32),
        ("layout", ctypes.c_int32),
        ("proj", ctypes.c_int32),
        ("merge", ctypes.c_int32),
        ("dtype", ctypes.c_int32),
        ("apply_elu", ctypes.c_int32),
        ("alpha", ctypes.c_float),
    ]


class ConvDesc(ctypes.Structure):
    _fields_ = [
        ("n", ctypes.c_int32), ("h", ctypes.c_int32), ("w", ctypes.c_int32), ("cin", ctypes.c_int32),
        ("cout", ctypes.c_int32), ("kh", ctypes.c_int32), ("kw", ctypes.c_int32),
        ("stride", ctypes.c_int32),
        ("pad_top", ctypes.c_int32), ("pad_left", ctypes.c_int32),
        ("ho", ctypes.c_int32), ("wo", ctypes.c_int32),
        ("dtype", ctypes.c_int32),
        ("act", ctypes.c_int32),
        ("groups", ctypes.c_int32),
    ]


class StreamDesc(ctypes.Structure):
    _fields_ = [
        ("nodes", ctypes.c_int32), ("ci", ctypes.c_int32), ("co", ctypes.c_int32), ("heads", ctypes.c_int32),
        ("layout", ctypes.c_int32), ("mapping", ctypes.c_int32), ("transpose_adj", ctypes.c_int32),
        ("wgrad_cols", ctypes.c_int32),
    ]


class LayerDesc(ctypes.Structure):
    _fields_ = [
        ("n", ctypes.c_int32), ("h", ctypes.c_int32), ("w", ctypes.c_int32),
        ("nodes", ctypes.c_int32), ("ci", ctypes.c_int32), ("co", ctypes.c_int32), ("heads", ctypes.c_int32),
        ("layout", ctypes.c_int32), ("merge", ctypes.c_int32), ("apply_elu", ctypes.c_int32),
        ("alpha", ctypes.c_float),
        ("x_layout", ctypes.c_int32),  # X_RECORDS (default) or X_PLANAR
    ]


X_RECORDS, X_PLANAR = 0, 1


_P = ctypes.c_void_p
_I = ctypes.c_int
_I64 = ctypes.c_int64
_F = ctypes.c_float

# name -> argtypes; every symbol declared in include/cgat_b200.h (tests check the export list)
SIGNATURES = {
    "cgat_attn_fwd": [ctypes.POINTER(AttnDesc), _P, _P, _P, _P, _P, _P, _P, _P],
    "cgat_attn_pixstats": [ctypes.POINTER(AttnDesc), _P, _P, _P, _P, _P, _P],
    "cgat_attn_bwd": [ctypes.POINTER(AttnDesc), _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "cgat_attn_pixstats_bwd": [ctypes.POINTER(AttnDesc), _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "cgat_adj_norm_fwd": [_P, _P, _I, _I, _I, _P],
    "cgat_adj_norm_bwd": [_P, _P, _P, _I, _I, _I, _P],
    "cgat_conv2d_fprop": [ctypes.POINTER(ConvDesc), _P, _P, _P, _P, _I, _P, _P],
    "cgat_conv2d_dgrad": [ctypes.POINTER(ConvDesc), _P, _P, _P, _I, _P, _P],
    "cgat_conv2d_wgrad": [ctypes.POINTER(ConvDesc), _P, _P, _P, _P, _I, _P, _P],
    "cgat_conv_tc_supported": [ctypes.POINTER(ConvDesc), _I],
    "cgat_conv_workspace_bytes": [ctypes.POINTER(ConvDesc), _I],
    "cgat_conv_dbias_workspace_bytes": [ctypes.POINTER(ConvDesc)],
    "cgat_conv_stream_supported": [ctypes.POINTER(ConvDesc), _I],
    "cgat_conv_stream_workspace_bytes": [ctypes.POINTER(ConvDesc)],
    "cgat_conv2d_fprop_packed": [ctypes.POINTER(ConvDesc), _P, _P, _P, _P, _P],
    "cgat_conv2d_dgrad_packed": [ctypes.POINTER(ConvDesc), _P, _P, _P, _P],
    "cgat_conv2d_wgrad_partial": [ctypes.POINTER(ConvDesc), _P, _P, _P, ctypes.POINTER(ctypes.c_int32),
                                  ctypes.POINTER(ctypes.c_int32), _P],
    "cgat_stream_wpack_bytes": [ctypes.POINTER(StreamDesc), _I],
    "cgat_stream_prepare": [ctypes.POINTER(StreamDesc), _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "cgat_stream_prepare_clear": [ctypes.POINTER(StreamDesc), _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _P],
    "cgat_stream_param_grads": [ctypes.POINTER(StreamDesc), _P, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P],
    "cgat_layer_supported": [ctypes.POINTER(LayerDesc)],
    "cgat_layer_workspace_bytes": [ctypes.POINTER(LayerDesc)],
    "cgat_layer_fwd": [ctypes.POINTER(LayerDesc), _P, _P, _P, _P, _P, _P, _P, _P],
    "cgat_layer_bwd": [ctypes.POINTER(LayerDesc), _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                       ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32), _P],
    "cgat_layer_train": [ctypes.POINTER(LayerDesc), _P, _P, _P, _P, _P, _P, _P, _F, _P, _P, _P, _P, _P, _P, _P,
                         ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32), _P],
    "cgat_layer_train_fp32": [ctypes.POINTER(LayerDesc), _P, _P, _P, _P, _P, _P, _P, _F, _P, _P, _P, _P, _P, _P, _P,
                              ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32), _P],
    "cgat_stream_finish": [ctypes.POINTER(StreamDesc), _P, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P, _I64,
                           _P, _P, _P, _P, _P, _P, _I64, _P, _P, _P],
    "cgat_stream_finish_mirror": [ctypes.POINTER(StreamDesc), _P, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P,
                                  _I64, _P, _P, _P, _P, _P, _P, _I64, _P, _P, _P, _P, _I, _P, _P],
    "cgat_gat1d_fwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _F, _P],
    "cgat_gat1d_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _F, _P],
    "cgat_loss_fwd_bwd": [_P, _P, _P, _P, _P, _I64, _F, _F, _I, _P],
    "cgat_adam_step": [_P, _P, _P, _P, _P, _I64, _F, _F, _F, _F, _F, _F, _P],
    "cgat_adam_step_at": [_P, _P, _P, _P, _I64, _I64, _F, _F, _F, _F, _F, _F, _P],
    "cgat_cast": [_P, _I, _P, _I, _I64, _P],
    "cgat_p2p_mailbox_bytes": [_I64, _I],
    "cgat_p2p_allreduce_adam": [_P, _I, _I, _P, _P, _P, _P, _P, _I64, _I64, _F, _F, _F, _F, _F, _P],
    "cgat_p2p_allreduce_adam_graph": [_P, _I, _I, _P, _P, _P, _P, _P, _P, _I64, _P],
    "cgat_val_metrics": [_P, _P, _I64, _F, _F, _F, _I, _P, _P],
    "cgat_loader_gather": [_P, _I64, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _F, _I, _P],
    "cgat_loader_gather_planar": [_P, _I64, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _F, _P],
    "cgat_s2d_pad": [_P, _P, _I, _I64, _I, _I, _I, _I, _P],
    "cgat_loader_gather_f32": [_P, _I64, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "cgat_records_to_planar": [_P, _P, _I64, _I, _I, _I, _P],
    "cgat_comm_available": [],
    "cgat_comm_unique_id": [_P],
    "cgat_comm_init": [_I, _I, _P, _P],
    "cgat_flat_allreduce": [_P, _P, _I64, _P],
    "cgat_comm_destroy": [_P],
    "cgat_bn_workspace_bytes": [_I],
    "cgat_bn_stats": [_P, _I, _I64, _I64, _I, _P, _P, _P, _P, _P, _P, _F, _F, _P],
    "cgat_bn_act_fwd": [_P, _P, _I, _I64, _I64, _I, _P, _P, _P, _P, _P, _I, _F, _P],
    "cgat_bn_act_bwd": [_P, _P, _P, _I, _I64, _I64, _I, _P, _P, _P, _P, _P, _I, _F, _I, _P, _P, _P, _I, _P],
    "cgat_bn_workspace_bytes_sets": [_I, _I],
    "cgat_bn_stats_sets": [_P, _I, _I64, _I64, _I, _I, _P, _P, _P, _P, _P, _P, _F, _F, _P],
    "cgat_bn_act_fwd_sets": [_P, _P, _I, _I64, _I64, _I, _I, _P, _P, _P, _P, _P, _I, _F, _P],
    "cgat_bn_act_bwd_sets": [_P, _P, _P, _I, _I64, _I64, _I, _I, _P, _P, _P, _P, _P, _I, _F, _I, _P, _P, _P, _P, _I, _P],
    "cgat_dropout2d_mask": [_P, _I64, _F, ctypes.c_uint64, _P, _P],
    "cgat_maxpool2_fwd": [_P, _P, _P, _I, _I64, _I, _I, _I, _P],
    "cgat_maxpool2_bwd": [_P, _P, _P, _I, _I64, _I, _I, _I, _P],
    "cgat_upcat_fwd": [_P, _P, _P, _I, _I64, _I, _I, _I, _I, _I, _I, _P],
    "cgat_upcat_bwd": [_P, _P, _P, _I, _I64, _I, _I, _I, _I, _I, _I, _P],
    "cgat_pool_hw_workspace_bytes": [_I64, _I64, _I],
    "cgat_pool_hw": [_P, _I, _I64, _I64, _I, _P, _P, _P, _P, _P],
    "cgat_dot_hw": [_P, _P, _I, _I64, _I64, _I, _P, _P, _P],
    "cgat_cbam_mlp_fwd": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P],
    "cgat_cbam_mlp_bwd": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "cgat_gate_channels_fwd": [_P, _P, _P, _I, _I64, _I64, _I, _P],
    "cgat_gate_channels_bwd": [_P, _P, _P, _P, _P, _P, _I, _I64, _I64, _I, _P],
    "cgat_gate_pixels": [_P, _P, _P, _I, _I64, _I64, _I, _P],
    "cgat_chan_pool_fwd": [_P, _P, _P, _I, _I64, _I, _P],
    "cgat_chan_pool_bwd": [_P, _P, _P, _I, _I64, _I, _P],
    "cgat_chan_dot": [_P, _P, _P, _I, _I64, _I, _P],
}

_lib = None


def lib() -> ctypes.CDLL:
    """Load the CUDA library (once).  Raises if it has not been built -- there is no CPU path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C extended-gan_b200/csrc`).  The conv-GAT layer has no CPU fallback."
            )
        L = ctypes.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = ctypes.c_int
        L.cgat_conv_workspace_bytes.restype = ctypes.c_int64
        L.cgat_conv_dbias_workspace_bytes.restype = ctypes.c_int64
        L.cgat_conv_stream_workspace_bytes.restype = ctypes.c_int64
        L.cgat_stream_wpack_bytes.restype = ctypes.c_int64
        L.cgat_layer_workspace_bytes.restype = ctypes.c_int64
        L.cgat_p2p_mailbox_bytes.restype = ctypes.c_int64
        L.cgat_bn_workspace_bytes.restype = ctypes.c_int64
        L.cgat_bn_workspace_bytes_sets.restype = ctypes.c_int64
        L.cgat_pool_hw_workspace_bytes.restype = ctypes.c_int64
        L.cgat_version.restype = ctypes.c_char_p
        L.cgat_last_error.restype = ctypes.c_char_p
        _lib = L
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().cgat_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (rc={rc}): {msg}")


def dtype_tag(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise RuntimeError(f"unsupported dtype {t.dtype}: the CUDA kernels take float32 or bfloat16")


def ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "conv-GAT kernels need CUDA tensors on an sm_100 device (no CPU fallback); got a tensor on "
                f"{t.device}"
            )


# ---- launch accounting / optional per-kernel CUDA-event timing (used by bench.py) -------------------
LAUNCHES = 0  # kernels launched through the C ABI since import (bench.py reports the per-step count)
_PROFILE = None  # None, or {name: [(start_event, end_event), ...]}


def profile_start():
    global _PROFILE
    _PROFILE = {}


def profile_stop():
    """Returns {kernel name: (count, mean milliseconds)}; must be called after a synchronize."""
    global _PROFILE
    prof, _PROFILE = _PROFILE, None
    out = {}
    for name, evs in (prof or {}).items():
        ms = [a.elapsed_time(b) for a, b in evs]
        out[name] = (len(ms), sum(ms) / max(1, len(ms)))
    return out


def call(name: str, *args, launches: int = 1):
    """Invoke one C-ABI entry point on the current stream, count its launches, raise on error."""
    global LAUNCHES
    fn = getattr(lib(), name)
    if _PROFILE is not None:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        rc = fn(*args)
        b.record()
        _PROFILE.setdefault(name, []).append((a, b))
    else:
        rc = fn(*args)
    LAUNCHES += launches
    check(rc, name)


def ptr_array(tensors):
    """Host array of device pointers (``const float* const*`` of the C ABI); None entries become NULL."""
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr
