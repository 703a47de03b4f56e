"""ctypes binding of the C ABI declared in ``include/wd_b200.h``.

There is no CPU fallback: if the shared library is missing or a call fails, this raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_lib", "libwd_b200.so")

WD_OK, WD_IGNORED = 0, 1
VARIANT_UNET, VARIANT_PHOSC = 0, 1
STEP_EPS_ONLY, STEP_DDPM, STEP_DDIM = 0, 1, 2


class WdConfig(C.Structure):
    _fields_ = [("variant", C.c_int), ("in_channels", C.c_int), ("model_channels", C.c_int),
                ("out_channels", C.c_int), ("num_res_blocks", C.c_int), ("n_channel_mult", C.c_int),
                ("channel_mult", C.c_int * 8), ("n_attention_resolutions", C.c_int),
                ("attention_resolutions", C.c_int * 8), ("num_heads", C.c_int), ("num_head_channels", C.c_int),
                ("transformer_depth", C.c_int), ("context_dim", C.c_int), ("vocab_size", C.c_int),
                ("num_classes", C.c_int), ("max_seq_len", C.c_int), ("latent_h", C.c_int), ("latent_w", C.c_int),
                ("add_label_emb", C.c_int), ("phosc_len", C.c_int)]


# name -> (restype, argtypes); must list every symbol of include/wd_b200.h (tests/test_abi.py checks it)
_P, _I, _F, _I64, _U64 = C.c_void_p, C.c_int, C.c_float, C.c_int64, C.c_uint64
SIGNATURES = {
    "wd_last_error": (C.c_char_p, []),
    "wd_version": (_I, []),
    "wd_engine_create": (_I, [C.POINTER(WdConfig), C.POINTER(_P)]),
    "wd_engine_destroy": (None, [_P]),
    "wd_engine_load_param": (_I, [_P, C.c_char_p, _P, C.POINTER(_I64), _I, _P]),
    "wd_engine_set_pos_encoding": (_I, [_P, _P, _P]),
    "wd_engine_finalize_params": (_I, [_P, _P]),
    "wd_engine_reserve": (_I, [_P, _I]),
    "wd_engine_workspace_bytes": (C.c_size_t, [_P]),
    "wd_engine_weight_bytes": (C.c_size_t, [_P]),
    "wd_engine_last_launch_count": (_I, [_P]),
    "wd_engine_set_profiling": (_I, [_P, _I]),
    "wd_engine_profile_read": (_I, [_P, _I, C.POINTER(_I), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                    C.POINTER(_F), C.POINTER(_I)]),
    "wd_encode_context": (_I, [_P, _I, _P, _I, _P, _P]),
    "wd_unet_eval": (_I, [_P, _I, _P, _P, _I64, _P, _P, _P]),
    "wd_sampler_step": (_I, [_P, _I, _P, _I64, _P, _I, C.POINTER(_F), _P, _I, _U64, _U64, _I, _P, _P]),
    "wd_phosc_tokenize": (_I, [_P, _I, _I, _P, _P, _P]),
    "wd_sampler_update": (_I, [_P, _P, _I, _I, _I, C.POINTER(_F), _P, _I, _U64, _U64, _I, _P]),
    "wd_op_groupnorm": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _F, _I, _P]),
    "wd_op_layernorm": (_I, [_P, _P, _P, _P, _I, _I, _F, _P]),
    "wd_op_gemm": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "wd_op_gemm_f16": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "wd_op_conv3x3": (_I, [_P, _P, _P, _P, _I, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "wd_op_conv3x3_gn_silu": (_I, [_P, _P, _P, _P, _I, _P, _P, _F, _P, _P, _I, _I, _I, _I, _I, _P]),
    "wd_op_pack_conv3x3": (_I, [_P, _P, _I, _I, _P]),
    "wd_op_pack_linear": (_I, [_P, _P, _I, _I, _I, _P]),
    "wd_op_pack_vec_geglu": (_I, [_P, _P, _I, _P]),
    "wd_op_attention_small": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _P]),
    "wd_op_q_ctx_attention": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _P]),
    "wd_op_attention": (_I, [_P, _I, _P, _P, _I, _P, _I, _I, _I, _I, _I, _F, _P]),
    "wd_op_gemm_block_n": (_I, []),
    "wd_op_tblock_unet": (_I, [C.POINTER(_P), _I, _I, _I, _I, _I, _P, _P, _P]),
    # ---- training step ----
    "wd_trainer_create": (_I, [C.POINTER(WdConfig), C.POINTER(_P)]),
    "wd_trainer_destroy": (None, [_P]),
    "wd_trainer_bind_param": (_I, [_P, C.c_char_p, _P, _P, C.POINTER(_I64), _I]),
    "wd_trainer_set_pos_encoding": (_I, [_P, _P, _P]),
    "wd_trainer_sync_weights": (_I, [_P, _P]),
    "wd_trainer_forward": (_I, [_P, _I, _P, _P, _P, _P, _I, _P, _P]),
    "wd_trainer_backward": (_I, [_P, _P, _P, _P, _P]),
    "wd_trainer_launch_counts": (_I, [_P, C.POINTER(_I), C.POINTER(_I)]),
    "wd_trainer_num_grad_stages": (_I, [_P, C.POINTER(_I)]),
    "wd_trainer_grad_stage": (_I, [_P, C.c_char_p, C.POINTER(_I)]),
    "wd_trainer_backward_stages": (_I, [_P, _P, _I, _I, _P]),
    "wd_trainer_read_tensor": (_I, [_P, C.c_char_p, _P, C.c_size_t, _P]),
    "wd_trainer_workspace_bytes": (C.c_size_t, [_P]),
    "wd_trainer_weight_bytes": (C.c_size_t, [_P]),
    "wd_adamw_ema_step": (_I, [_P, _P, _P, _P, _P, C.c_size_t, _F, _F, _F, _F, _F, _I, _F, _I, _F, _P]),
    "wd_op_wgrad_linear": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "wd_op_wgrad_conv3x3": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "wd_op_pack_conv3x3_t": (_I, [_P, _P, _I, _I, _P]),
    "wd_op_pack_linear_t": (_I, [_P, _P, _I, _I, _P]),
    "wd_op_groupnorm_bwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _I, _P]),
    "wd_op_layernorm_bwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _F, _P]),
    "wd_op_geglu_fwd": (_I, [_P, _P, _I, _I, _P]),
    "wd_op_geglu_bwd": (_I, [_P, _P, _P, _I, _I, _P]),
    "wd_op_attention_small_bwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _P]),
    # ---- fp32 mode ----
    "wd_f32_create": (_I, [C.POINTER(WdConfig), C.POINTER(_P)]),
    "wd_f32_destroy": (None, [_P]),
    "wd_f32_load_param": (_I, [_P, C.c_char_p, _P, C.POINTER(_I64), _I, _P]),
    "wd_f32_set_pos_encoding": (_I, [_P, _P, _P]),
    "wd_f32_encode_context": (_I, [_P, _I, _P, _I, _P, _P]),
    "wd_f32_unet_eval": (_I, [_P, _I, _P, _P, _I64, _P, _P, _P]),
    "wd_f32_unet_eval_maps": (_I, [_P, _I, _P, _P, _I64, _P, _P, _P]),
    "wd_f32_read_attention_map": (_I, [_P, _I, _I, _P, C.POINTER(_I), C.POINTER(_I), C.POINTER(_I), _P]),
    "wd_f32_read_context": (_I, [_P, _P, C.c_size_t, _P]),
    "wd_set_context": (_I, [_P, _I, _P, _I, _P]),
    "wd_lerp": (_I, [_P, _P, _F, _P, C.c_size_t, _P]),
    "wd_noise_images": (_I, [_P, _P, _P, _I, _P, C.c_uint64, C.c_uint64, C.c_uint32, _P, _P, _I, _I, _P]),
    "wd_mse_workspace_bytes": (C.c_size_t, [C.c_size_t]),
    "wd_mse_loss_grad": (_I, [_P, _P, _P, _P, _P, C.c_size_t, _P]),
    "wd_engine_set_label_mix": (_I, [_P, _I, _I, _I, _F, _P]),
    "wd_f32_set_context": (_I, [_P, _I, _P, _I, _P]),
    "wd_f32_set_label_mix": (_I, [_P, _I, _I, _I, _F, _P]),
    "wd_f32_ctc_head": (_I, [_P, _I, _P, _I, _I, _I, _P, _P]),
    "wd_f32_op_linear": (_I, [_P, _P, _P, _P, _I, _I, _I, _P]),
    "wd_vae_create": (_I, [C.POINTER(_P)]),
    "wd_vae_decode": (_I, [_P, _I, _P, _I, _I, _F, _I, _P, _I, _P]),
    "wd_f32_last_launch_count": (_I, [_P]),
    "wd_f32_workspace_bytes": (C.c_size_t, [_P]),
    "wd_f32_op_conv3x3": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "wd_f32_op_gemm_tc": (_I, [_P, _P, _P, _P, _I, _I, _I, _P]),
    "wd_f32_op_conv3x3_tc": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "wd_f32_op_attention": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _P]),
}

_lib = None


class WdError(RuntimeError):
    pass


def lib():
    """Load libwd_b200.so (built by worddiffusion_b200.build).  Raises if it is missing -- no fallback."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise WdError(f"{LIB_PATH} not found: build it with `python -m worddiffusion_b200.build` "
                          "(the B200 path has no CPU / eager fallback)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc, what=""):
    if rc < 0:
        msg = lib().wd_last_error()
        raise WdError(f"{what}: wd_b200 error {rc}: {msg.decode() if msg else ''}")
    return rc
