"""ctypes binding of the C-ABI shared library (include/medfusion_b200.h).

The library is the product: if it is missing or fails to load, importing any compute path raises —
there is deliberately no PyTorch/CPU fallback.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libmedfusion_b200.so")
MF_MAX_LEVELS = 8


class MedfusionLibError(RuntimeError):
    pass


class UNetConfig(ctypes.Structure):
    _fields_ = [
        ("in_ch", c_int), ("out_ch", c_int), ("depth", c_int),
        ("hid_chs", c_int * MF_MAX_LEVELS), ("kernel_sizes", c_int * MF_MAX_LEVELS),
        ("strides", c_int * MF_MAX_LEVELS), ("num_res_blocks", c_int), ("emb_dim", c_int),
        ("pos_emb_dim", c_int), ("num_classes", c_int), ("norm_groups", c_int),
        ("attention", c_int * MF_MAX_LEVELS), ("deep_supervision", c_int), ("ds_out_ch", c_int),
    ]


class VAEConfig(ctypes.Structure):
    _fields_ = [
        ("emb_channels", c_int), ("out_channels", c_int), ("depth", c_int),
        ("hid_chs", c_int * MF_MAX_LEVELS), ("strides", c_int * MF_MAX_LEVELS), ("norm_groups", c_int),
        ("in_channels", c_int), ("num_embeddings", c_int),
    ]


class SchedTables(ctypes.Structure):
    _fields_ = [(n, c_void_p) for n in (
        "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_mean_coef1",
        "posterior_mean_coef2", "posterior_variance", "betas", "alphas_cumprod")]


class SchedOpts(ctypes.Structure):
    _fields_ = [("d_pred_var", c_void_p), ("d_pred_var_uncond", c_void_p), ("pred_batch_stride", c_int64),
                ("cold_diffusion", c_int), ("sqrt_alphas_cumprod", c_void_p),
                ("sqrt_one_minus_alphas_cumprod", c_void_p), ("T", c_int)]


class StepArgs(ctypes.Structure):
    _fields_ = [
        ("tables", POINTER(SchedTables)), ("d_pred_uncond", c_void_p), ("guidance_scale", c_float),
        ("d_noise", c_void_p), ("d_t_next", c_void_p), ("d_noise_ddim", c_void_p),
        ("objective_is_x0", c_int), ("clip_x0", c_int),
        ("d_x_prior", c_void_p), ("d_x_0", c_void_p), ("d_x_T", c_void_p), ("d_x_next", c_void_p),
        ("uniform_t", c_int),
    ]


# name -> (restype, argtypes); every symbol declared in include/medfusion_b200.h
_P = c_void_p
SIGNATURES = {
    "mf_last_error": (c_char_p, []),
    "mf_abi_version": (c_int, []),
    "mf_saturation_count": (c_int, [POINTER(ctypes.c_ulonglong), c_int, _P]),
    "mf_set_drain_interval": (c_int, [c_int]),
    "mf_set_cta_group": (c_int, [c_int]),
    "mf_set_block_n": (c_int, [c_int]),
    "mf_set_stream_k": (c_int, [c_int]),
    "mf_set_split_fill": (c_int, [c_int]),
    "mf_set_row_patch": (c_int, [c_int]),
    "mf_op_conv_tc_plan": (c_int, [c_int] * 10 + [ctypes.POINTER(c_int)]),
    "mf_set_pdl": (c_int, [c_int]),
    "mf_set_fold_upsample": (c_int, [c_int]),
    "mf_set_stem_on_tc": (c_int, [c_int]),
    "mf_set_gn_variant": (c_int, [c_int]),
    "mf_set_fold_head": (c_int, [c_int]),
    "mf_set_fuse_gn": (c_int, [c_int]),
    "mf_set_attn_tc": (c_int, [c_int]),
    "mf_set_debias_eps": (c_int, [c_float]),
    "mf_unet_create": (c_int, [POINTER(UNetConfig), POINTER(_P)]),
    "mf_unet_destroy": (None, [_P]),
    "mf_unet_param_count": (c_int, [_P]),
    "mf_unet_param_name": (c_char_p, [_P, c_int]),
    "mf_unet_param_shape": (c_int, [_P, c_int, POINTER(c_int64), POINTER(c_int)]),
    "mf_unet_set_param": (c_int, [_P, c_char_p, _P, POINTER(c_int64), c_int, _P]),
    "mf_unet_set_time_freqs": (c_int, [_P, _P, c_int, _P]),
    "mf_unet_workspace_bytes": (c_size_t, [_P, c_int, c_int, c_int]),
    "mf_unet_forward": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, _P, c_size_t, _P]),
    "mf_unet_forward_ex": (c_int, [_P, _P, _P, _P, _P, _P, POINTER(_P), c_int, c_int, c_int, c_int, _P, c_size_t, _P]),
    "mf_unet_forward_step": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, _P, c_size_t, POINTER(StepArgs), _P]),
    "mf_unet_forward_step_cfg": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, _P, c_size_t, POINTER(StepArgs), _P]),
    "mf_unet_profile": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, _P, c_size_t, _P, POINTER(c_float),
                                POINTER(c_int), POINTER(ctypes.c_double), c_int, POINTER(c_int)]),
    "mf_unet_plan_info": (c_int, [_P, POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "mf_vae_create": (c_int, [POINTER(VAEConfig), POINTER(_P)]),
    "mf_vae_destroy": (None, [_P]),
    "mf_vae_param_count": (c_int, [_P]),
    "mf_vae_param_name": (c_char_p, [_P, c_int]),
    "mf_vae_param_shape": (c_int, [_P, c_int, POINTER(c_int64), POINTER(c_int)]),
    "mf_vae_set_param": (c_int, [_P, c_char_p, _P, POINTER(c_int64), c_int, _P]),
    "mf_vae_workspace_bytes": (c_size_t, [_P, c_int, c_int, c_int]),
    "mf_vae_decode": (c_int, [_P, _P, _P, c_int, c_int, c_int, _P, c_size_t, _P]),
    "mf_vae_decode_u8": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, _P, c_size_t, _P]),
    "mf_vae_encode_workspace_bytes": (c_size_t, [_P, c_int, c_int, c_int]),
    "mf_vae_encode": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, _P, c_size_t, _P]),
    "mf_vae_profile": (c_int, [_P, _P, _P, c_int, c_int, c_int, _P, c_size_t, _P, POINTER(c_float), POINTER(c_int),
                               POINTER(ctypes.c_double), c_int, POINTER(c_int)]),
    "mf_vae_plan_info": (c_int, [_P, POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "mf_sched_step": (c_int, [POINTER(SchedTables), _P, _P, _P, c_float, _P, _P, _P, _P, c_int, c_int,
                              _P, _P, _P, _P, c_int, c_int, _P]),
    "mf_sched_step_opts": (c_int, [POINTER(SchedTables), _P, _P, _P, c_float, _P, _P, _P, _P, c_int, c_int,
                                   _P, _P, _P, _P, c_int, c_int, POINTER(SchedOpts), _P]),
    "mf_op_pack_split": (c_int, [_P, _P, c_int64, c_int, c_int, c_int, c_int, _P]),
    "mf_op_unpack_nchw": (c_int, [_P, c_int64, c_int, _P, c_int, c_int, c_int, c_int, _P]),
    "mf_op_prep_weight_tc": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "mf_op_prep_weight_simt": (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P]),
    "mf_op_conv_tc_supported": (c_int, [c_int] * 8),
    "mf_op_conv_tc_stats_chunks": (c_int, [c_int, c_int]),
    "mf_op_conv_tc": (c_int, [_P, c_int64, c_int, _P, c_int64, c_int, c_int, c_int, c_int, _P, _P, c_int, c_int, _P,
                              _P, c_int64, c_int, _P, c_int, c_int, _P]),
    "mf_op_prep_weight_up_tc": (c_int, [_P, _P, _P, c_int, c_int, _P]),
    "mf_op_upconv_tc": (c_int, [_P, c_int64, c_int, c_int, c_int, c_int, _P, _P, c_int, _P, _P, c_int64, _P]),
    "mf_op_conv_simt": (c_int, [_P, c_int64, c_int, c_int, c_int, c_int, c_int, _P, _P, c_int, c_int, c_int, _P,
                                c_int64, c_int, _P]),
    "mf_op_gn_partial": (c_int, [_P, _P, c_int, c_int, c_int, _P]),
    "mf_op_gn_finalize": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, c_float, _P]),
    "mf_op_gn_apply": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int, _P, c_int, _P, c_int64, c_int, c_int, c_int,
                               c_int, _P]),
    "mf_op_attention": (c_int, [_P, _P, _P, c_int, _P, c_int64, c_int, c_int, c_int, c_int, _P]),
    "mf_op_layernorm": (c_int, [_P, c_int64, _P, _P, _P, c_int64, c_int64, c_int, c_float, _P]),
    "mf_op_geglu": (c_int, [_P, _P, c_int64, c_int64, c_int, _P]),
    "mf_op_upsample2x": (c_int, [_P, c_int64, _P, c_int64, c_int, c_int, c_int, c_int, _P]),
    "mf_op_vq_quantize": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, _P]),
}

_lib = None


def load():
    """Load libmedfusion_b200.so (built by `__graft_entry__.build()` / csrc/Makefile). Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MedfusionLibError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no fallback path.")
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover
        raise MedfusionLibError(f"failed to load {LIB_PATH}: {e}") from e
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    # tuning knobs (defaults are the parity-safe, fastest-known settings)
    if os.environ.get("MF_CTA_GROUP"):
        lib.mf_set_cta_group(int(os.environ["MF_CTA_GROUP"]))
    if os.environ.get("MF_FOLD_UPSAMPLE"):
        lib.mf_set_fold_upsample(int(os.environ["MF_FOLD_UPSAMPLE"]))
    if os.environ.get("MF_DEBIAS_EPS"):
        lib.mf_set_debias_eps(float(os.environ["MF_DEBIAS_EPS"]))
    if os.environ.get("MF_PDL"):
        lib.mf_set_pdl(int(os.environ["MF_PDL"]))
    if os.environ.get("MF_STREAM_K"):
        lib.mf_set_stream_k(int(os.environ["MF_STREAM_K"]))
    if os.environ.get("MF_SPLIT_FILL"):
        lib.mf_set_split_fill(int(os.environ["MF_SPLIT_FILL"]))
    if os.environ.get("MF_ROW_PATCH"):
        lib.mf_set_row_patch(int(os.environ["MF_ROW_PATCH"]))
    if os.environ.get("MF_ATTN_TC"):
        lib.mf_set_attn_tc(int(os.environ["MF_ATTN_TC"]))
    if os.environ.get("MF_FUSE_GN"):
        lib.mf_set_fuse_gn(int(os.environ["MF_FUSE_GN"]))
    if os.environ.get("MF_FOLD_HEAD"):
        lib.mf_set_fold_head(int(os.environ["MF_FOLD_HEAD"]))
    if os.environ.get("MF_GN_VARIANT"):
        lib.mf_set_gn_variant(int(os.environ["MF_GN_VARIANT"]))
    if os.environ.get("MF_BLOCK_N"):
        lib.mf_set_block_n(int(os.environ["MF_BLOCK_N"]))
    if os.environ.get("MF_DRAIN_INTERVAL"):
        lib.mf_set_drain_interval(int(os.environ["MF_DRAIN_INTERVAL"]))
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().mf_last_error()
        msg = msg.decode() if msg else "unknown error"
        exc = ValueError if rc == 2 else RuntimeError
        raise exc(f"medfusion_b200 {what} failed (status {rc}): {msg}")
