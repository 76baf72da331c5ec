"""ctypes binding of libs2r_b200.so (the C ABI declared in include/s2r_b200.h).

The product path has no CPU or eager fallback: if the shared library is missing, or a call is
made without a CUDA device, the error is raised here and propagates.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libs2r_b200.so")

S2R_MAX_TAPS = 16
ACT_NONE, ACT_RELU, ACT_RELU6, ACT_LEAKY = 0, 1, 2, 3
AUX_NONE, AUX_ADD, AUX_LEAKY_MASK = 0, 1, 2

vp = C.c_void_p
i32 = C.c_int
i64 = C.c_int64
u64 = C.c_uint64
f32 = C.c_float
f64 = C.c_double


class Tap(C.Structure):
    _fields_ = [("base", vp), ("sn", i64), ("sh", i64), ("sw", i64), ("H", i32), ("W", i32),
                ("dh", i32), ("dw", i32), ("wslice", i32), ("_pad", i32), ("wofs", i64)]


class BnTail(C.Structure):
    _fields_ = [("count", f64), ("sums", vp), ("gamma", vp), ("beta", vp), ("running_mean", vp), ("running_var", vp),
                ("mean_invstd", vp), ("scale_shift", vp), ("eps", f32), ("momentum", f32),
                ("clamp_mode", i32), ("channel", i32)]


class BnEvalJob(C.Structure):
    _fields_ = [("gamma", vp), ("beta", vp), ("running_mean", vp), ("running_var", vp), ("mean_invstd", vp),
                ("scale_shift", vp), ("C", i32), ("eps", f32)]


class ConvArgs(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("ntaps", i32), ("taps", Tap * S2R_MAX_TAPS),
                ("N", i32), ("OH", i32), ("OW", i32), ("Cin", i32), ("Cout", i32),
                ("w", vp), ("Cout_pad", i32), ("Kpad", i32),
                ("out", vp), ("on", i64), ("oh", i64), ("ow", i64),
                ("bias", vp), ("act", i32), ("slope", f32), ("aux_mode", i32), ("_pad", i32),
                ("aux", vp), ("an", i64), ("ah", i64), ("aw", i64), ("stats", vp), ("oscale", vp)]


class WgradArgs(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("ntaps", i32), ("taps", Tap * S2R_MAX_TAPS),
                ("N", i32), ("OH", i32), ("OW", i32), ("Cin", i32), ("Cout", i32),
                ("dy", vp), ("dn", i64), ("dh", i64), ("dw", i64),
                ("dweight", vp), ("s_co", i64), ("s_ci", i64)]


class PackJob(C.Structure):
    _fields_ = [("w", vp), ("packed", vp), ("Cout", i32), ("Cin", i32), ("RS", i32), ("transpose", i32),
                ("A_pad", i32), ("B_pad", i32), ("begin", i64), ("end", i64)]


class ResizeJob(C.Structure):
    _fields_ = [("inp", vp), ("out", vp), ("bounds", vp), ("kk", vp), ("W", i32), ("C", i32), ("ksize", i32), ("flip", i32),
                ("axis", i32), ("o0", i32), ("on", i32), ("lines", i32), ("in_pitch", i32), ("base", i32)]


class NearestJob(C.Structure):
    _fields_ = [("inp", vp), ("out", vp), ("xtab", vp), ("ytab", vp), ("W", i32), ("OH", i32), ("OW", i32), ("x0", i32),
                ("y0", i32), ("flip", i32)]


class StageJob(C.Structure):
    _fields_ = [("img", vp), ("label", vp), ("out_img", vp), ("out_label", vp), ("Hs", i32), ("Ws", i32), ("flip", i32),
                ("x1", i32), ("y1", i32), ("_pad", i32)]


class BlurJob(C.Structure):
    _fields_ = [("img", vp), ("tmp", vp), ("out", vp), ("Hs", i32), ("Ws", i32), ("flip", i32), ("x1", i32), ("y1", i32),
                ("ww", C.c_uint32), ("fw", C.c_uint32), ("_pad", C.c_uint32)]


class ParamSlot(C.Structure):
    _fields_ = [("p", vp), ("g", vp), ("s0", vp), ("s1", vp), ("n", i64), ("lr_mult", f32),
                ("_pad", f32)]


# name -> argtypes (the trailing stream argument included); every function returns int
PROTOTYPES = {
    "s2r_conv_fwd": [C.POINTER(ConvArgs), vp],
    "s2r_conv_fwd_mma": [C.POINTER(ConvArgs), vp],
    "s2r_conv_wgrad": [C.POINTER(WgradArgs), vp],
    "s2r_pack_weight": [vp, i32, i32, i32, i32, i32, vp, i32, i32, vp],
    "s2r_pack_weights_multi": [vp, i32, vp],
    "s2r_rowtap_wgrad_scatter": [vp, vp, i32, i32, vp],
    "s2r_resize_bilinear_u8": [vp, i32, i32, i32, i32, i32, i32, vp, vp, i32, i32, vp, vp],
    "s2r_export_prediction_nchw": [vp, i32, i32, i32, i32, vp, vp, i32, i32, vp, vp, i32, vp, vp, vp],
    "s2r_resize_bilinear_u8_multi": [vp, i32, i64, vp],
    "s2r_resize_nearest_u8_multi": [vp, i32, i64, vp],
    "s2r_gaussian_blur3_u8_multi": [vp, i32, i32, i32, vp],
    "s2r_input_stage_u8_multi": [vp, i32, vp, vp, vp, i32, i32, i32, vp],
    "s2r_resize_nearest_u8": [vp, i32, i32, i32, vp, vp, i32, i32, i32, vp, vp],
    "s2r_input_stage_u8": [vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, i32, vp, vp, i32, i32, vp],
    "s2r_wgrad_scatter_taps": [vp, vp, i32, i32, i32, i32, vp],
    "s2r_softmax0_nchw_to_nhwc_pad": [vp, i32, i32, i32, i32, i32, vp, i32, vp],
    "s2r_softmax0_nhwc_pad_bwd": [vp, vp, i32, i32, i32, i32, i32, i32, vp, vp],
    "s2r_softmax0_batch_stats": [vp, i32, i64, vp, vp, vp],
    "s2r_softmax0_nchw_to_nhwc_pad_global": [vp, i32, i32, i32, i32, vp, vp, vp, i32, vp],
    "s2r_softmax0_nhwc_pad_bwd_global": [vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp],
    "s2r_dwconv3x3_fwd": [vp, vp, i32, i32, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp],
    "s2r_dwconv3x3_fwd_bn": [vp, C.POINTER(BnTail), vp, i32, i32, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp],
    "s2r_dwconv3x3_dgrad": [vp, vp, vp, vp, vp, i32, i32, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp],
    "s2r_dwconv3x3_wgrad": [vp, vp, i32, i32, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp],
    "s2r_dwconv3x3_bwd": [vp, vp, vp, vp, vp, i32, i32, i32, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp],
    "s2r_im2col_nchw_f32": [vp, i32, i32, i32, i32, i32, i32, i32, i32, vp, i32, vp],
    "s2r_channel_sums_bf16": [vp, i64, i32, i32, i32, vp, vp],
    "s2r_bn_finalize": [vp, f64, vp, vp, f32, i32, f32, vp, vp, vp, vp, i32, vp],
    "s2r_bn_tail_run": [C.POINTER(BnTail), i32, vp],
    "s2r_bn_eval_multi": [vp, i32, vp],
    "s2r_bn_eval_scale_shift": [vp, vp, vp, vp, f32, vp, vp, i32, vp],
    "s2r_bn_apply_act": [vp, i64, i32, i32, i32, vp, i32, vp, f32, u64, vp, vp, i32, i32, vp],
    "s2r_bn_apply_act_bn": [vp, i64, i32, i32, i32, C.POINTER(BnTail), vp, i32, vp, f32, u64, vp, vp, i32, i32, vp],
    "s2r_bn_bwd_reduce": [vp, i32, i32, vp, i32, i32, vp, vp, i32, f32, u64, vp, i64, i32, vp, vp],
    "s2r_bn_bwd_apply": [vp, i32, i32, vp, i32, i32, vp, vp, i32, f32, u64, vp, vp, f64, i64, i32, vp,
                         i32, i32, vp, vp, i32, i32, i32, vp],
    "s2r_upsample_bilinear_nhwc": [vp, i32, i32, i32, i32, vp, i32, i32, i32, i32, vp],
    "s2r_upsample_bilinear_nhwc_bwd": [vp, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp],
    "s2r_upsample_bilinear_nhwc_to_nchw": [vp, i32, i32, i32, i32, i32, vp, i32, i32, vp],
    "s2r_upsample_bilinear_nchw_bwd_to_nhwc": [vp, i32, i32, i32, i32, vp, i32, i32, i32, vp],
    "s2r_upsample_bilinear_nchw_bwd_to_nhwc_scaled": [vp, i32, i32, i32, i32, vp, i32, i32, i32, vp, vp],
    "s2r_avgpool_nhwc": [vp, i32, i32, i32, i32, i32, f32, vp, vp, vp],
    "s2r_broadcast_nhwc": [vp, i32, i32, i32, f32, i32, vp, i32, i32, vp],
    "s2r_nchw_f32_to_nhwc_bf16": [vp, i32, i32, i64, vp, i32, vp],
    "s2r_nhwc_bf16_to_nchw_f32": [vp, i32, i32, i32, i64, vp, vp],
    "s2r_leaky_relu_bwd_bf16": [vp, vp, vp, i64, f32, vp],
    "s2r_add_bf16": [vp, vp, i64, vp],
    "s2r_add_f64_to_f32": [vp, vp, i32, vp],
    "s2r_softmax_dim0_fwd": [vp, vp, i32, i64, vp],
    "s2r_softmax_dim0_bwd": [vp, vp, vp, i32, i64, vp],
    "s2r_cross_entropy_nchw": [vp, vp, i32, vp, i32, i32, i64, i32, vp, vp, vp],
    "s2r_ratio": [vp, f64, vp, vp],
    "s2r_scale_by_ratio": [vp, i64, vp, vp, f64, vp],
    "s2r_bce_logits_fwd": [vp, vp, f32, i64, vp, vp],
    "s2r_bce_logits_bwd": [vp, vp, f32, i64, vp, vp, vp],
    "s2r_confusion_matrix": [vp, i32, vp, i64, i32, vp, vp, vp],
    "s2r_argmax_confusion_nchw": [vp, vp, i32, i32, i64, i32, vp, vp, vp],
    "s2r_upsample_argmax_confusion_nhwc": [vp, i32, i32, i32, i32, i32, vp, i32, i32, i32, vp, vp],
    "s2r_comm_create": [i32, i32, i32, vp],
    "s2r_comm_open": [vp],
    "s2r_allreduce_small_f64": [vp, i32, vp],
    "s2r_allreduce_small_f64_ch": [vp, i32, i32, vp],
    "s2r_sgd_step": [vp, i32, vp, f32, f32, f32, i32, f32, vp],
    "s2r_adam_step": [vp, i32, vp, f32, f32, f32, f32, f32, vp],
    "s2r_store_f32": [vp, i32, C.POINTER(f32), vp],
}
PLAIN = {"s2r_version": (i32, []), "s2r_last_error": (C.c_char_p, []), "s2r_device_ok": (i32, []),
         "s2r_comm_ready": (i32, []), "s2r_comm_error": (i32, []), "s2r_comm_destroy": (i32, [])}

_lib = None
launches = 0  # number of C-ABI compute calls issued by this process (bench.py reports it)


class S2RError(RuntimeError):
    pass


def lib():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise S2RError(
                "libs2r_b200.so is missing (%s): run __graft_entry__.build(); there is no fallback path"
                % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, argtypes in PROTOTYPES.items():
            fn = getattr(L, name)
            fn.argtypes = argtypes
            fn.restype = i32
        for name, (res, argtypes) in PLAIN.items():
            fn = getattr(L, name)
            fn.argtypes = argtypes
            fn.restype = res
        _lib = L
    return _lib


def last_error():
    return lib().s2r_last_error().decode("utf-8", "replace")


PROFILE = None  # when a list: (name, signature, start_event, end_event) per call (tests/tools/step_profile.py)


def _signature(name, args):
    a = getattr(args[0], "_obj", None) if args else None
    if a is not None and hasattr(a, "ntaps"):
        return "taps=%d N=%d OH=%d OW=%d Cin=%d Cout=%d" % (a.ntaps, a.N, a.OH, a.OW, a.Cin, a.Cout)
    return " ".join(str(v) for v in args if isinstance(v, int) and not isinstance(v, bool))[:60]


def call(name, *args):
    """Invoke a C-ABI entry point and convert a non-zero status into an exception."""
    global launches
    if PROFILE is not None:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib(), name)(*args)
        e1.record()
        PROFILE.append((name, _signature(name, args), e0, e1))
    else:
        rc = getattr(lib(), name)(*args)
    launches += 1
    if rc != 0:
        msg = last_error()
        if rc == -1:
            raise ValueError("%s: %s" % (name, msg))
        if rc == -2:
            raise NotImplementedError("%s: %s" % (name, msg))
        raise S2RError("%s failed (%d): %s" % (name, rc, msg))
    return rc


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise S2RError("s2r_b200 needs a CUDA device (sm_100a); there is no CPU path")
