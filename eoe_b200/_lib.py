"""ctypes binding of libeoe_b200.so (the C ABI declared in include/eoe_b200.h).

No fallback of any kind: if the shared library is missing or a call fails, this raises.
PyTorch is only used for device memory, streams and dtype bookkeeping.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# EOE_B200_LIB selects another build of the same library (A/B timing of kernel versions, tools/ab_encoder.py); there is
# still no fallback: the named file must exist and export the full ABI.
LIB_PATH = os.environ.get("EOE_B200_LIB") or os.path.join(_HERE, "libeoe_b200.so")

EOE_F32, EOE_F16, EOE_BF16, EOE_F16X2 = 0, 1, 2, 3
F16X2 = "f16x2"      # operand_dtype of the precise mode (split fp16 pairs, include/eoe_b200.h EOE_F16X2)
EOE_HEAD_WS_BYTES = 32768
EOE_AUC_IGNORE_NEGATIVE_LABELS = 1
EOE_AUC_WITH_PRC = 2
EOE_AUC_FORCE_TILED = 4
EOE_AUC_FORCE_SINGLE_CTA = 8
EOE_AUC_FORCE_CLUSTER = 16
EOE_AUC_SINGLE_LAUNCH_MAX = 49152
EOE_AUC_STATUS_NONFINITE = 1
EOE_AUC_STATUS_SINGLE_CLASS = 2
EOE_EPI_BIAS, EOE_EPI_BIAS_QUICKGELU, EOE_EPI_BIAS_RESIDUAL_F32, EOE_EPI_PATCH_EMBED = 0, 1, 2, 3
EOE_EPI_LNFOLD_BIAS, EOE_EPI_LNFOLD_QUICKGELU, EOE_EPI_RESIDUAL_STATS = 4, 5, 6
EOE_EPI_LNFOLD_QUICKGELU_X1702 = 8
GELU_SLOPE = 1.702
EOE_ABI_VERSION = 6
EOE_LAYOUT_NCHW, EOE_LAYOUT_NHWC = 0, 1
LAYOUT_RESIZE = 2            # host-side tag only: raw [B,H,W,3] pixels of another size -> eoe_vit_encode_u8_resize

DTYPE_CODE = {torch.float32: EOE_F32, torch.float16: EOE_F16, torch.bfloat16: EOE_BF16}


class EoeError(RuntimeError):
    pass


class VitLayer(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "ln_1_w", "ln_1_b", "in_proj_w", "in_proj_b", "out_proj_w", "out_proj_b",
        "ln_2_w", "ln_2_b", "c_fc_w", "c_fc_b", "c_proj_w", "c_proj_b",
        "in_proj_wf", "in_proj_c1", "in_proj_c2", "c_fc_wf", "c_fc_c1", "c_fc_c2", "c_proj_w_div1702")]


class VitWeights(C.Structure):
    _fields_ = [
        ("patch", C.c_int32), ("resolution", C.c_int32), ("width", C.c_int32), ("heads", C.c_int32),
        ("n_layers", C.c_int32), ("embed_dim", C.c_int32), ("operand_dtype", C.c_int32), ("reserved", C.c_int32),
        ("conv1_w", C.c_void_p), ("class_embedding", C.c_void_p), ("positional_embedding", C.c_void_p),
        ("ln_pre_w", C.c_void_p), ("ln_pre_b", C.c_void_p), ("ln_post_w", C.c_void_p), ("ln_post_b", C.c_void_p),
        ("proj", C.c_void_p), ("layers_host", C.POINTER(VitLayer)),
    ]


_P, _I64, _I, _F, _SZ = C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_size_t

# name -> (restype, argtypes); mirrors include/eoe_b200.h one to one (tests/test_capi_symbols.py checks it)
SIGNATURES = {
    "eoe_abi_version": (_I, []),
    "eoe_strerror": (C.c_char_p, [_I]),
    "eoe_last_cuda_error": (C.c_char_p, []),
    "eoe_launch_count": (C.c_longlong, []),
    "eoe_build_id": (C.c_char_p, []),
    "eoe_hsc_fwd_bwd": (_I, [_P, _I, _P, _I64, _I64, _I64, _P, _P, _P, _P, _P]),
    "eoe_hsc_score": (_I, [_P, _I, _I64, _I64, _P, _P]),
    "eoe_bce_fwd_bwd": (_I, [_P, _I, _P, _I64, _I64, _P, _P, _P, _P, _P]),
    "eoe_bce_score": (_I, [_P, _I, _I64, _I64, _P, _P]),
    "eoe_dsad_fwd_bwd": (_I, [_P, _I, _P, _I64, _I64, _I64, _P, _P, _P, _P, _P]),
    "eoe_dsvdd_fwd_bwd": (_I, [_P, _I, _P, _I64, _I64, _P, _P, _P, _P, _P]),
    "eoe_focal_fwd_bwd": (_I, [_P, _I, _P, _I64, _I64, _F, _F, _P, _P, _P, _P, _P]),
    "eoe_clip_score": (_I, [_P, _I, _P, _I64, _I64, _I64, _F, _P, _P]),
    "eoe_clip_oe_loss_fwd_bwd": (_I, [_P, _I, _P, _P, _I64, _I64, _I64, _F, _I64, _I, _P, _P, _P, _P]),
    "eoe_auc_workspace_bytes": (_SZ, [_I64]),
    "eoe_auc": (_I, [_P, _I, _P, _I64, _I, _P, _SZ, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "eoe_vit_workspace_bytes": (_SZ, [C.POINTER(VitWeights), _I64]),
    "eoe_vit_plan_create": (_I, [C.POINTER(VitWeights), _I64, _P, _SZ, C.POINTER(_P)]),
    "eoe_vit_plan_destroy": (None, [_P]),
    "eoe_vit_encode": (_I, [_P, _P, _I64, _P, _P, _I64, _F, _P, _P]),
    "eoe_vit_encode_u8": (_I, [_P, _P, _I, C.POINTER(C.c_float), C.POINTER(C.c_float), _I64, _P, _P, _I64, _F, _P, _P]),
    "eoe_vit_encode_u8_resize": (_I, [_P, _P, _I64, _I64, C.POINTER(C.c_float), C.POINTER(C.c_float), _I64, _P, _P, _I64, _F, _P, _P]),
    "eoe_resize_geometry": (_I, [_I64, _I64, _I, C.POINTER(C.c_int)]),
    "eoe_vit_profile_enable": (_I, [_P, _I]),
    "eoe_vit_profile_read": (_I, [_P, C.POINTER(C.c_double), C.POINTER(_I64), C.POINTER(C.c_double)]),
    "eoe_debug_set": (None, [_I]),
    "eoe_gemm": (_I, [_P, _P, _P, _P, _I64, _I64, _I64, _I, _I, _P, _I64, _P]),
    "eoe_vit_fold_layernorm": (_I, [_P, _P, _P, _P, _I64, _I64, _I, _P, _P, _P, _P]),
    "eoe_gemm_lnfold": (_I, [_P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I64, _I, _I, _P]),
    "eoe_gemm_residual_stats": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I64, _I, _P]),
    "eoe_layernorm": (_I, [_P, _P, _P, _P, _I, _I64, _I64, _P]),
    "eoe_attention": (_I, [_P, _P, _I64, _I64, _I64, _I, _P]),
    "eoe_attention_causal": (_I, [_P, _P, _I64, _I64, _I64, _I, _P]),
    "eoe_text_embed": (_I, [_P, _P, _P, _P, _I64, _I64, _I64, _I64, _P]),
    "eoe_text_tail": (_I, [_P, _P, _P, _P, _P, _P, _I64, _I64, _I64, _I64, _P]),
}

_lib = None
MISSING = []


def lib():
    """Load libeoe_b200.so once; raise (never fall back) if it is absent or incomplete."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EoeError(
            f"{LIB_PATH} is missing: build it with `python -m eoe_b200.build` "
            "(there is no CPU / PyTorch fallback for the eoe_b200 hot path)")
    l = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(l, name)
        except AttributeError:
            # a stale build: calling the symbol later raises AttributeError (loud, no fallback);
            # tests/test_capi_symbols.py requires every declared symbol to be exported
            MISSING.append(name)
            continue
        fn.restype = res
        fn.argtypes = args
    _lib = l
    return l


def check(rc: int, what: str):
    if rc != 0:
        l = lib()
        msg = l.eoe_strerror(rc).decode()
        detail = l.eoe_last_cuda_error().decode()
        raise EoeError(f"{what} failed: {msg}" + (f" [{detail}]" if rc == -6 and detail else ""))


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise EoeError("eoe_b200 kernels need CUDA tensors (no CPU fallback); got a tensor on " + str(t.device))


def dtype_code(t):
    try:
        return DTYPE_CODE[t.dtype]
    except KeyError:
        raise EoeError(f"unsupported dtype {t.dtype}; use float32, float16 or bfloat16")


_head_ws = {}


def head_workspace(device):
    """One zero-initialised 32 KiB reduction workspace per (device, stream)."""
    key = (device.index if device.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(device).cuda_stream)
    ws = _head_ws.get(key)
    if ws is None:
        ws = torch.zeros(EOE_HEAD_WS_BYTES, dtype=torch.uint8, device=device)
        _head_ws[key] = ws
    return ws
