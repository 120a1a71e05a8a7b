"""ctypes binding of libvividb200.so — the C ABI declared in include/vivid_b200.h.

The library is GPU-only.  Importing this module never touches CUDA; `lib()` loads the
shared object (building it first if the sources are newer) and raises if it is missing.
There is deliberately no CPU fallback: a product call without the CUDA extension fails loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# (VB_LIB_PATH: A/B measurements against another build of the same ABI — tools only; the product loads the in-tree library)
LIB_PATH = os.environ.get("VB_LIB_PATH") or os.path.join(_HERE, "libvividb200.so")

VB_F32, VB_F16, VB_BF16, VB_F64, VB_U8 = 0, 1, 2, 3, 4
VB_EPI_PLAIN, VB_EPI_QKVNORM = 0, 1
VB_F_MODSILU, VB_F_CLIP, VB_F_RESB_FOLDED = 1, 4, 8
VB_RES_NONE, VB_RES_PLAIN, VB_RES_PIXNORM, VB_RES_SCALED = 0, 1, 2, 3
VB_OUT_NONE, VB_OUT_RAW, VB_OUT_SILU, VB_OUT_NORM, VB_OUT_NORM_SILU = 0, 1, 2, 3, 4
VB_EW_PIXNORM, VB_EW_DOWN_PIXNORM, VB_EW_UP, VB_EW_CAT, VB_EW_SILU = 0, 1, 2, 3, 4

i32, i64, f32, vp = C.c_int32, C.c_int64, C.c_float, C.c_void_p


class WeightPrepDesc(C.Structure):
    _fields_ = [("src", vp), ("dst", vp), ("src_dtype", i32), ("dst_dtype", i32), ("cout", i32), ("cin", i32),
                ("taps", i32), ("cout_pad", i32), ("split", i32), ("seg_a_pad", i32), ("seg_b_pad", i32),
                ("perm_parts", i32), ("perm_dim", i32), ("gain", f32), ("scale_a", f32), ("scale_b", f32)]


class ConvDesc(C.Structure):
    _fields_ = [("x", vp), ("x2", vp), ("w", vp), ("mod", vp), ("res", vp), ("out", vp * 3), ("out_f32", vp),
                ("out_rnorm", vp), ("res_rnorm", vp), ("part_out", vp * 3), ("B", i32), ("H", i32), ("W", i32), ("cin_pad", i32), ("cin2_pad", i32),
                ("cout_pad", i32), ("taps", i32), ("block_n", i32), ("epi_mode", i32), ("flags", i32),
                ("mod_stride", i32), ("ld_f32", i32), ("res_mode", i32), ("out_kind", i32 * 3), ("head_dim", i32),
                ("parts", i32), ("seg_div", i32), ("part_seq", i32 * 3), ("part_off", i32 * 3),
                ("out_scale", f32 * 3), ("res_t", f32), ("clip", f32), ("tune", i32), ("part_ld", i32), ("ks_ws", vp)]


class AttnDesc(C.Structure):
    _fields_ = [("q", vp), ("k", vp), ("v", vp), ("y", vp), ("B", i32), ("heads", i32), ("sq", i32), ("sk", i32),
                ("head_dim", i32), ("zero_keys", i32), ("ld", i32), ("q_prescaled", i32)]


class EwDesc(C.Structure):
    _fields_ = [("a", vp), ("b", vp), ("out", vp), ("out_silu", vp), ("kind", i32),
                ("B", i32), ("H", i32), ("W", i32), ("ca", i32), ("cb", i32), ("wa", f32), ("wb", f32)]


class EmbDesc(C.Structure):
    _fields_ = [("sigma", vp), ("geom", vp), ("freqs", vp), ("phases", vp), ("w_noise", vp), ("w_label", vp),
                ("w_mod", vp), ("emb", vp), ("mod", vp), ("B", i32), ("sigma_n", i32), ("sigma_stride", i32),
                ("cnoise", i32), ("cemb", i32), ("label_dim", i32), ("mod_total", i32), ("geom_rows", i32),
                ("label_balance", f32), ("noise_scale", f32), ("geom_scale", f32)]


class PrecondInDesc(C.Structure):
    _fields_ = [("x", vp), ("cond", vp), ("noise", vp), ("sigma", vp), ("out", vp), ("B", i32), ("R", i32),
                ("cpad", i32), ("sigma_n", i32), ("sigma_stride", i32), ("im2col", i32), ("img_stride", i64),
                ("sigma_data", f32), ("noisy_sr", f32)]


class PrecondOutDesc(C.Structure):
    _fields_ = [("x", vp), ("f", vp), ("sigma", vp), ("d_out", vp), ("B", i32), ("R", i32), ("ldf", i32),
                ("sigma_n", i32), ("sigma_stride", i32), ("img_stride", i64), ("sigma_data", f32)]


class HeunDesc(C.Structure):
    _fields_ = [("d_net", vp), ("d_gnet", vp), ("x_hat", vp), ("d_cur", vp), ("x_next", vp), ("x_out", vp * 2),
                ("sigma_out", vp * 2), ("n", i64), ("phase", i32), ("sigma_n", i32), ("guidance", f32), ("t_hat", f32),
                ("t_next", f32), ("sigma_next", f32)]


class F32ConvDesc(C.Structure):
    _fields_ = [("x", vp), ("w", vp), ("out", vp), ("B", i32), ("H", i32), ("W", i32), ("cin", i32), ("cout", i32),
                ("taps", i32), ("ldo", i32)]


class F32OpDesc(C.Structure):
    _fields_ = [("a", vp), ("b", vp), ("b2", vp), ("mod", vp), ("out", vp), ("out2", vp), ("out3", vp), ("img_stride", i64),
                ("kind", i32), ("flags", i32), ("B", i32), ("H", i32), ("W", i32), ("ca", i32), ("cb", i32),
                ("mod_stride", i32), ("heads", i32), ("parts", i32), ("head_dim", i32), ("seg_div", i32),
                ("part_seq", i32 * 3), ("part_off", i32 * 3), ("wa", f32), ("wb", f32), ("clip", f32)]


VB_F32_ACT, VB_F32_SUM, VB_F32_CAT, VB_F32_DOWN, VB_F32_UP, VB_F32_QKV, VB_F32_PRECOND_IN = range(7)
VB_F32_NORM, VB_F32_MOD, VB_F32_SILU = 1, 2, 4


class StatsDesc(C.Structure):
    _fields_ = [("feat", vp), ("feat2", vp), ("cum_mu", vp), ("cum_sigma", vp), ("ld1", i64), ("ld2", i64),
                ("dtype", i32), ("n", i32), ("f1", i32), ("f2", i32)]


# name -> (restype, argtypes); the smoke/CPU tests check that every one of these is exported.
class IoDesc(C.Structure):
    _fields_ = [("in_x", vp), ("in_src", vp), ("in_sigma", vp), ("in_geom", vp), ("in_cond", vp), ("in_noise", vp), ("out_d", vp),
                ("n_x", i64), ("n_out", i64), ("img_elems", i64), ("geom_dim", i64), ("workspace_bytes", i64)]


NOISE_FN = C.CFUNCTYPE(C.c_int, vp, vp, i64, vp)      # vb_noise_fn(user, dst, n, stream)


class SampleDesc(C.Structure):
    _fields_ = [("net", vp), ("gnet", vp), ("noise", vp), ("t_steps", C.POINTER(C.c_float)), ("workspace", vp), ("x_out", vp),
                ("sr_noise", NOISE_FN), ("sr_noise_user", vp), ("side_stream", vp), ("num_steps", i32), ("net_first_op", i32),
                ("guidance", f32)]


class UNetDesc(C.Structure):
    _fields_ = [("img_resolution", i32), ("in_channels", i32), ("out_channels", i32), ("model_channels", i32), ("num_levels", i32),
                ("channel_mult", i32 * 8), ("num_blocks", i32), ("num_attn_res", i32), ("attn_resolutions", i32 * 8),
                ("extra_attn", i32), ("channels_per_head", i32), ("xattn", i32), ("label_dim", i32), ("cnoise", i32), ("cemb", i32),
                ("label_balance", C.c_double), ("concat_balance", C.c_double), ("res_balance", C.c_double),
                ("attn_balance", C.c_double), ("clip_act", C.c_double)]


class NetDesc(C.Structure):
    _fields_ = [("unet", UNetDesc), ("encoder", UNetDesc), ("has_encoder", i32), ("img_resolution", i32), ("uncond", i32),
                ("super_res", i32), ("dual_source", i32), ("no_time_enc", i32), ("sigma_data", C.c_double), ("noisy_sr", C.c_double)]


class Param(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", vp), ("dtype", i32), ("ndim", i32), ("shape", i64 * 4)]


SIGNATURES = {
    "vb_last_error": (C.c_char_p, []),
    "vb_abi_version": (C.c_int, []),
    "vb_device_check": (C.c_int, []),
    "vb_struct_size": (C.c_int, [C.c_int]),
    "vb_operand_dtype": (C.c_int, []),
    "vb_weight_prep": (C.c_int, [C.POINTER(WeightPrepDesc), vp]),
    "vb_conv": (C.c_int, [C.POINTER(ConvDesc), vp]),
    "vb_conv_ksplit_ws_bytes": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "vb_attn": (C.c_int, [C.POINTER(AttnDesc), vp]),
    "vb_eltwise": (C.c_int, [C.POINTER(EwDesc), vp]),
    "vb_embed": (C.c_int, [C.POINTER(EmbDesc), vp]),
    "vb_precond_in": (C.c_int, [C.POINTER(PrecondInDesc), vp]),
    "vb_precond_out": (C.c_int, [C.POINTER(PrecondOutDesc), vp]),
    "vb_heun": (C.c_int, [C.POINTER(HeunDesc), vp]),
    "vb_logvar": (C.c_int, [vp, C.c_int32, C.c_int32, vp, vp, vp, C.c_int32, vp, vp]),
    "vb_f32_conv": (C.c_int, [C.POINTER(F32ConvDesc), vp]),
    "vb_f32_op": (C.c_int, [C.POINTER(F32OpDesc), vp]),
    "vb_f32_attn": (C.c_int, [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp]),
    "vb_stats_update": (C.c_int, [C.POINTER(StatsDesc), vp]),
    "vb_psnr_u8": (C.c_int, [vp, vp, C.c_int32, C.c_int32, C.c_int64, C.c_int64, vp, vp, vp]),
    "vb_resize": (C.c_int, [vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp]),
    "vb_encode_u8": (C.c_int, [vp, vp, i64, vp]),
    "vb_decode_u8": (C.c_int, [vp, vp, i64, vp]),
    "vb_plan_create": (C.c_int, [C.POINTER(vp)]),
    "vb_plan_destroy": (None, [vp]),
    "vb_plan_add_conv": (C.c_int, [vp, C.POINTER(ConvDesc)]),
    "vb_plan_add_attn": (C.c_int, [vp, C.POINTER(AttnDesc)]),
    "vb_plan_add_eltwise": (C.c_int, [vp, C.POINTER(EwDesc)]),
    "vb_plan_add_embed": (C.c_int, [vp, C.POINTER(EmbDesc)]),
    "vb_plan_add_precond_in": (C.c_int, [vp, C.POINTER(PrecondInDesc)]),
    "vb_plan_add_precond_out": (C.c_int, [vp, C.POINTER(PrecondOutDesc)]),
    "vb_plan_add_heun": (C.c_int, [vp, C.POINTER(HeunDesc)]),
    "vb_spin": (C.c_int, [C.c_int, vp]),
    "vb_set_pdl": (C.c_int, [C.c_int]),
    "vb_debug_conv_cycles": (C.c_int, [C.POINTER(C.c_ulonglong)]),
    "vb_debug_conv_stamps": (C.c_int, [C.POINTER(C.c_longlong)]),
    "vb_plan_num_ops": (C.c_int, [vp]),
    "vb_plan_run": (C.c_int, [vp, C.c_int, C.c_int, vp]),
    "vb_plan_launch_graph": (C.c_int, [vp, vp]),
    "vb_plan_launch_graph_range": (C.c_int, [vp, C.c_int, C.c_int, vp]),
    "vb_plan_query": (C.c_double, [vp, C.c_int]),
    "vb_plan_bind_io": (C.c_int, [vp, C.POINTER(IoDesc)]),
    "vb_denoise": (C.c_int, [vp, vp, vp, vp, i32, vp, i32, vp, vp, vp, vp]),
    "vb_workspace_bytes": (C.c_int64, [vp]),
    "vb_plan_set_inputs": (C.c_int, [vp, vp, vp, i32, vp, vp]),
    "vb_sample_workspace_bytes": (C.c_int64, [vp]),
    "vb_sample": (C.c_int, [C.POINTER(SampleDesc), vp]),
    "vb_net_plan_create": (C.c_int, [C.POINTER(NetDesc), C.POINTER(Param), i32, i32, vp, C.POINTER(vp)]),
    "vb_net_plan_set_weights": (C.c_int, [vp, C.POINTER(Param), i32, vp]),
    "vb_plan_get_io": (C.c_int, [vp, C.POINTER(IoDesc), C.POINTER(i32)]),
    "vb_plan_num_features": (C.c_int, [vp]),
    "vb_plan_get_feature": (C.c_int, [vp, i32, C.POINTER(vp), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]),
    "vb_net_plan_trace": (C.c_int64, [C.POINTER(NetDesc), C.POINTER(Param), i32, i32, C.c_char_p, i64]),
    "vb_trace_desc": (C.c_int64, [i32, vp, C.c_char_p, i64]),
}

STRUCTS = [WeightPrepDesc, ConvDesc, AttnDesc, EwDesc, EmbDesc, PrecondInDesc, PrecondOutDesc, HeunDesc, StatsDesc, F32ConvDesc, F32OpDesc, IoDesc, SampleDesc, UNetDesc, NetDesc, Param]

_lib = None


class VividB200Error(RuntimeError):
    pass


def lib():
    """Load (once) and return the ctypes handle of libvividb200.so. Raises if it cannot be loaded."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VividB200Error(
            f"{LIB_PATH} is missing: build it with `python -m vivid_b200.build` "
            "(vivid_b200 has no CPU fallback; the CUDA extension is required)")
    h = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(h, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    for i, st in enumerate(STRUCTS):
        if h.vb_struct_size(i) != C.sizeof(st):
            raise VividB200Error(f"ABI mismatch: {st.__name__} is {C.sizeof(st)} bytes here, "
                                 f"{h.vb_struct_size(i)} in {LIB_PATH}; rebuild with `python -m vivid_b200.build`")
    _lib = h
    return h


def operand_torch_dtype():
    """torch dtype of GEMM operands / the residual stream in the loaded build (fp16 unless built with -DVB_OP_BF16)."""
    import torch
    return {VB_F16: torch.float16, VB_BF16: torch.bfloat16}[lib().vb_operand_dtype()]


def check(rc, what="vivid_b200 call"):
    if rc != 0:
        msg = lib().vb_last_error()
        raise VividB200Error(f"{what} failed ({rc}): {msg.decode() if msg else '?'}")


def ptr(t):
    """Device/host pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()
