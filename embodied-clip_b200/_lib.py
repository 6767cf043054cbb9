"""ctypes binding of libembclip_b200.so (C ABI: include/embclip_b200.h).

Fails loudly: a missing library is an ImportError-grade RuntimeError, never a fallback."""
from __future__ import annotations

import ctypes as C
import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libembclip_b200.so")

DTYPE_F16, DTYPE_F32 = 0, 1


class RN50Cfg(C.Structure):
    _fields_ = [("layers", C.c_int32 * 4), ("width", C.c_int32), ("heads", C.c_int32),
                ("output_dim", C.c_int32), ("input_resolution", C.c_int32), ("arch", C.c_int32)]


class ParamInfo(C.Structure):
    _fields_ = [("name", C.c_char * 64), ("dtype", C.c_int32), ("ndim", C.c_int32),
                ("shape", C.c_int64 * 4), ("offset", C.c_uint64), ("nbytes", C.c_uint64)]


class ActInfo(C.Structure):
    _fields_ = [("name", C.c_char * 64), ("dtype", C.c_int32), ("n", C.c_int32), ("h", C.c_int32),
                ("w", C.c_int32), ("c", C.c_int32), ("offset", C.c_uint64)]


class ACCfg(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("feat_channels", "feat_pixels", "compress_hidden", "compress_out", "goal_dims",
                                         "combine_hidden", "combine_out", "hidden", "num_actions", "num_goals",
                                         "trainable_masked_hidden_state")]


class TFCfg(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("kind", "width", "layers", "heads", "output_dim", "patch_size", "input_resolution",
                                         "context_length", "vocab_size")]


TF_VISION, TF_TEXT = 0, 1

# symbol -> (restype, argtypes); this table is also what tests/test_abi.py checks against the header
_VP, _I, _U64, _FP, _LL, _F = C.c_void_p, C.c_int, C.c_uint64, C.c_void_p, C.c_longlong, C.c_float
SIGNATURES = {
    "embclip_last_error": (C.c_char_p, []),
    "embclip_abi_version": (_I, []),
    "embclip_rn50_create": (_I, [C.POINTER(RN50Cfg), C.POINTER(_VP)]),
    "embclip_rn50_destroy": (_I, [_VP]),
    "embclip_rn50_num_params": (_I, [_VP]),
    "embclip_rn50_param_info": (_I, [_VP, _I, C.POINTER(ParamInfo)]),
    "embclip_rn50_blob_bytes": (_U64, [_VP]),
    "embclip_rn50_bind_weights": (_I, [_VP, _VP, _U64]),
    "embclip_rn50_workspace_bytes": (_U64, [_VP, _I]),
    "embclip_rn50_forward": (_I, [_VP, _FP, _I, _FP, _FP, _FP, _VP, _U64, _VP]),
    "embclip_rn50_forward_u8": (_I, [_VP, _VP, C.POINTER(C.c_float), C.POINTER(C.c_float), _I, _FP, _FP, _FP, _VP, _U64, _VP]),
    "embclip_rn50_num_acts": (_I, [_VP]),
    "embclip_rn50_act_info": (_I, [_VP, _I, _I, C.POINTER(ActInfo)]),
    "embclip_rn50_profile": (_I, [_VP, _FP, _I, _FP, _FP, _FP, _VP, _U64, _VP, _VP, _VP, _I]),
    "embclip_rn50_profile_u8": (_I, [_VP, _VP, C.POINTER(C.c_float), C.POINTER(C.c_float), _I, _FP, _FP, _FP, _VP, _U64, _VP, _VP, _VP, _I]),
    "embclip_rn50_launches_per_forward": (_I, [_VP, _I, _I, _I]),
    "embclip_rn50_encode_rows_f16": (_I, [_VP, _VP, _I, C.POINTER(C.c_float), C.POINTER(C.c_float), _I, _VP, _VP, _U64, _VP]),
    "embclip_rn50_export_rows_f16": (_I, [_VP, _I, _VP, _U64, _VP, _VP]),
    "embclip_gemm_f16": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _I, _I, _I, _I, _I, _I, _VP]),
    "embclip_gemm_grouped_f16": (_I, [_VP, _I, _VP, _I, _I, _VP, _VP, _I, _I, _I, _I, _I, _I, _I, _I, _I, _VP]),
    "embclip_conv3x3_f16": (_I, [_VP, _VP, _VP, _VP, _I, _I, _I, _I, _I, _I, _I, _VP]),
    "embclip_avgpool2_f16": (_I, [_VP, _VP, _I, _I, _I, _I, _VP]),
    "embclip_pool2_f16": (_I, [_VP, _VP, _I, _I, _I, _I, _I, _VP]),
    "embclip_bneck_tail_f16": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, C.c_int64, _I, _VP]),
    "embclip_bneck_tail_pool_f16": (_I, [_VP, _VP, _FP, _VP, _VP, _I, _I, _VP, _FP, _VP, C.c_int64, _I, _VP]),
    "embclip_bneck_tail_stream_f16": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, C.c_int64, _I, _I, _I, _VP]),
    "embclip_stem_conv1": (_I, [_VP, _VP, _VP, _VP, _I, _I, _I, _VP]),
    # CLIP transformer towers
    "embclip_tf_create": (_I, [C.POINTER(TFCfg), C.POINTER(_VP)]),
    "embclip_tf_destroy": (_I, [_VP]),
    "embclip_tf_num_params": (_I, [_VP]),
    "embclip_tf_param_info": (_I, [_VP, _I, C.POINTER(ParamInfo)]),
    "embclip_tf_blob_bytes": (_U64, [_VP]),
    "embclip_tf_bind_weights": (_I, [_VP, _VP, _U64]),
    "embclip_tf_workspace_bytes": (_U64, [_VP, _I]),
    "embclip_tf_launches_per_forward": (_I, [_VP]),
    "embclip_vit_forward": (_I, [_VP, _FP, _I, _FP, _VP, _U64, _VP]),
    "embclip_text_forward": (_I, [_VP, _VP, _I, _FP, _VP, _U64, _VP]),
    "embclip_clip_logits": (_I, [_FP, _FP, _I, _I, _I, _F, _FP, _VP]),
    # actor-critic / PPO update
    "embclip_ac_create": (_I, [C.POINTER(ACCfg), C.POINTER(_VP)]),
    "embclip_ac_destroy": (_I, [_VP]),
    "embclip_ac_num_params": (_I, [_VP]),
    "embclip_ac_param_info": (_I, [_VP, _I, C.POINTER(ParamInfo)]),
    "embclip_ac_param_floats": (_U64, [_VP]),
    "embclip_ac_workspace_bytes": (_U64, [_VP, _I, _I]),
    "embclip_ac_num_acts": (_I, [_VP]),
    "embclip_ac_act_info": (_I, [_VP, _I, _I, _I, C.POINTER(ActInfo)]),
    "embclip_ac_pack_features": (_I, [_VP, _FP, _LL, _VP, _VP]),
    "embclip_ac_forward": (_I, [_VP, _FP, _VP, _VP, _FP, _FP, _I, _I, _FP, _FP, _FP, _VP, _U64, _I, _VP]),
    "embclip_ac_act": (_I, [_VP, _FP, _U64, _VP, _VP, _FP, _FP, _I, _FP, _VP, _FP, _FP, _FP, _FP, _VP, _U64, _VP]),
    "embclip_ac_ppo_loss": (_I, [_VP, _FP, _I, _I, _VP, _FP, _FP, _FP, _FP, _F, _F, _F, _F, _FP, _FP, _FP, _VP, _U64, _VP]),
    "embclip_ac_backward": (_I, [_VP, _FP, _VP, _VP, _FP, _FP, _I, _I, _FP, _FP, _FP, _FP, _VP, _U64, _VP]),
    "embclip_gae": (_I, [_FP, _FP, _FP, _I, _I, _F, _F, _FP, _FP, _FP, _F, _VP]),
    "embclip_sumsq_f32": (_I, [_FP, _LL, _FP, _VP]),
    "embclip_adam_clip_step": (_I, [_FP, _FP, _FP, _FP, _LL, _FP, _F, _F, _F, _F, _F, _I, _VP]),
    "embclip_wgrad_f16": (_I, [_VP, _I, _I, _VP, _I, _I, _LL, _FP, _LL, _LL, _FP, _VP]),
    "embclip_gru_geometry": (_I, [_I, _I, C.POINTER(C.c_int)]),
    "embclip_gru_forward": (_I, [_FP, _FP, _FP, _FP, _FP, _FP, _I, _I, _I, _FP, _FP, _FP, _FP, _FP, _VP, _VP]),
    "embclip_gru_backward": (_I, [_FP, _FP, _FP, _FP, _FP, _FP, _FP, _FP, _FP, _FP, _FP, _I, _I, _I, _FP, _FP, _VP, _FP, _FP, _VP, _VP]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library (once).  Raises RuntimeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"embclip_b200: {LIB_PATH} is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU / PyTorch fallback for this path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError here = header / library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class EmbclipError(RuntimeError):
    pass


def check(rc: int) -> int:
    if rc < 0:
        raise EmbclipError(f"embclip_b200 error {rc}: {load().embclip_last_error().decode()}")
    return rc
