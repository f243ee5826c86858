"""ctypes binding of libpcgan_kernels.so (the C ABI declared in include/pcgan_kernels.h).

The library is the only compute backend of this package: there is no CPU or
PyTorch fallback.  Importing this module never needs a GPU (the driver uses it
to check that the library loads and exports every declared symbol), but any
launch without one fails loudly with the CUDA error text.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PCGAN_KERNELS_LIB: another build of the same library (A/B measurements of a kernel change: tools/norm_bench.py)
LIB_PATH = os.environ.get("PCGAN_KERNELS_LIB") or os.path.join(_HERE, "libpcgan_kernels.so")

ABI_VERSION = 21
MAX_TAPS = 64

OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA = 0, -1, -2, -3
ACT_NONE, ACT_RELU, ACT_LRELU, ACT_TANH, ACT_SIGMOID = 0, 1, 2, 3, 4
DT_BF16, DT_F32 = 0, 1
HALO_ZERO, HALO_REFLECT = 0, 1
IGEMM_KMAJOR, IGEMM_WGRAD = 0, 1
STATS_NONE, STATS_ON = 0, 1
LOSS_BCE, LOSS_MSE, LOSS_L1, LOSS_ELO_NLL, LOSS_ELO_NLL_SCORE = 0, 1, 2, 3, 4

i32, i64, u32, u64, f32, vp = C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_float, C.c_void_p


class TMap(C.Structure):
    _fields_ = [("dims", u64 * 5), ("strides", u64 * 5), ("box", u32 * 5)]


class Comp(C.Structure):
    _fields_ = [("lo", i32), ("hi", i32), ("stride", i64)]


class IgemmDesc(C.Structure):
    _fields_ = [
        ("kind", i32), ("block_n", i32), ("a", TMap), ("b", TMap),
        ("t_count", i32 * 4),
        ("a_base", i32 * 4), ("a_step", (i32 * 4) * 4),
        ("b_base", i32 * 4), ("b_step", (i32 * 4) * 4),
        ("n_tiles", i32), ("m_tiles", i32), ("ksplit", i32),
        ("num_taps", i32), ("cchunks", i32),
        ("tap_off", (i32 * 4) * MAX_TAPS), ("tap_c0", i32 * MAX_TAPS), ("tap_bk", i32 * MAX_TAPS),
        ("e_base", i32 * 4), ("e_step", (i32 * 4) * 4), ("e_p1", i32 * 4), ("e_p2", i32 * 4),
        ("e_comp", (Comp * 3) * 4),
        ("out_dtype", i32), ("act", i32), ("act_slope", f32), ("n_valid", i32), ("out_cstride", i64),
        ("stats_mode", i32), ("stats_dim", i32), ("stats_comp", i32),
        ("m_valid", i32), ("wg_ncols", i32), ("ldo", i64), ("pair", i32), ("shift_taps", i32), ("shift_cpad", i32),
        ("a_window", i32), ("wg_box_dim", i32), ("tf32", i32), ("stats_div", i32),
    ]


class BatchItem(C.Structure):
    _fields_ = [("src", vp), ("idx", vp), ("dst", vp), ("n", i64)]


class RunningItem(C.Structure):
    _fields_ = [("stats", vp), ("running_mean", vp), ("running_var", vp), ("num_batches_tracked", vp),
                ("groups", i32), ("c", i32), ("count", f32), ("momentum", f32), ("sequential", i32), ("reserved_", i32)]


class PackArgs(C.Structure):
    _fields_ = [("src", vp), ("z", vp), ("mul_out", vp), ("mul_kind", i32), ("dst", vp),
                ("n", i32), ("cs", i32), ("h", i32), ("w", i32), ("ho", i32), ("wo", i32),
                ("cd", i32), ("pad", i32), ("halo", i32), ("dst_n_stride", i64)]


class UnpackArgs(C.Structure):
    _fields_ = [("g", vp), ("dst", vp), ("n", i32), ("c", i32), ("hs", i32), ("ws", i32), ("pad", i32),
                ("cd", i32), ("h", i32), ("w", i32), ("accumulate", i32), ("scale", f32)]


class NormFinalizeArgs(C.Structure):
    _fields_ = [("stats", vp), ("groups", i32), ("c", i32), ("count", f32), ("eps", f32), ("momentum", f32),
                ("gamma", vp), ("beta", vp), ("mean", vp), ("rstd", vp), ("scale", vp), ("shift", vp),
                ("running_mean", vp), ("running_var", vp), ("drop_mask", vp), ("in_groups", i32)]


class NormApplyArgs(C.Structure):
    _fields_ = [("x", vp), ("x_pad", i32), ("res", vp), ("res_pad", i32), ("y", vp), ("y_pad", i32), ("y_halo", i32),
                ("n", i32), ("h", i32), ("w", i32), ("c", i32),
                ("scale", vp), ("shift", vp), ("groups", i32),
                ("res_scale", vp), ("res_shift", vp), ("res_groups", i32),
                ("drop_mask", vp), ("act", i32), ("act_slope", f32), ("post_mask", vp),
                ("stats", vp), ("count", f32), ("eps", f32), ("gamma", vp), ("beta", vp),
                ("mean_out", vp), ("rstd_out", vp), ("scale_out", vp), ("shift_out", vp), ("y_c", i32), ("y_c0", i32)]


class FoldArgs(C.Structure):
    _fields_ = [("gpad", vp), ("g_pad", i32), ("halo", i32), ("add", vp), ("add_pad", i32),
                ("out", vp), ("out_pad", i32), ("n", i32), ("h", i32), ("w", i32), ("c", i32)]


class NormBwdArgs(C.Structure):
    _fields_ = [("dy", vp), ("dy_pad", i32), ("x", vp), ("x_pad", i32), ("res", vp), ("res_pad", i32),
                ("mean", vp), ("rstd", vp), ("scale", vp), ("shift", vp), ("groups", i32),
                ("res_scale", vp), ("res_shift", vp), ("res_groups", i32), ("drop_mask", vp),
                ("act", i32), ("act_slope", f32), ("n", i32), ("h", i32), ("w", i32), ("c", i32),
                ("count", f32), ("sums", vp), ("dx", vp), ("dx_pad", i32), ("dres", vp), ("dres_pad", i32), ("dy_fold", i32), ("post_mask", vp), ("affine", i32)]


class ImageItem(C.Structure):
    _fields_ = [("src", vp), ("h", i32), ("w", i32), ("crop_y", i32), ("crop_x", i32), ("flip", i32), ("reserved_", i32)]


class AugmentArgs(C.Structure):
    _fields_ = [("items", vp), ("dst", vp), ("n", i32), ("load", i32), ("fine", i32)]


class MaxpoolArgs(C.Structure):
    _fields_ = [("x", vp), ("x_pad", i32), ("y", vp), ("y_pad", i32), ("idx", vp),
                ("n", i32), ("h", i32), ("w", i32), ("c", i32), ("pool_pad", i32)]


class LossArgs(C.Structure):
    _fields_ = [("kind", i32), ("p", vp), ("target", vp), ("n", i64), ("per_sample", i64),
                ("weight", f32), ("weight_dev", vp), ("loss", vp), ("grad", vp)]


class AdamItem(C.Structure):
    _fields_ = [("p", vp), ("g", vp), ("m", vp), ("v", vp), ("n", i64)]


_STRUCTS = {
    "pcgan_tmap": TMap, "pcgan_comp": Comp, "pcgan_igemm_desc": IgemmDesc, "pcgan_pack_args": PackArgs,
    "pcgan_unpack_args": UnpackArgs, "pcgan_norm_finalize_args": NormFinalizeArgs,
    "pcgan_norm_apply_args": NormApplyArgs, "pcgan_fold_args": FoldArgs, "pcgan_norm_bwd_args": NormBwdArgs,
    "pcgan_maxpool_args": MaxpoolArgs, "pcgan_image_item": ImageItem, "pcgan_augment_args": AugmentArgs, "pcgan_loss_args": LossArgs, "pcgan_batch_item": BatchItem, "pcgan_running_item": RunningItem,
    "pcgan_adam_item": AdamItem,
}

# name -> (restype, argtypes); every symbol include/pcgan_kernels.h declares
SYMBOLS = {
    "pcgan_abi_version": (C.c_int, []),
    "pcgan_last_error": (C.c_char_p, []),
    "pcgan_sizeof": (i64, [C.c_char_p]),
    "pcgan_igemm_plan_create": (C.c_int, [C.POINTER(IgemmDesc), C.POINTER(vp)]),
    "pcgan_igemm_plan_destroy": (None, [vp]),
    "pcgan_selftest_fastdiv": (C.c_int64, [C.c_uint32, C.c_int32]),
    "pcgan_igemm_run": (C.c_int, [vp, vp, vp, vp, vp, vp, vp]),
    "pcgan_gather_cast_bf16": (C.c_int, [vp, vp, vp, i64, vp]),
    "pcgan_gather_tf32": (C.c_int, [vp, vp, vp, i64, vp]),
    "pcgan_scatter_f32": (C.c_int, [vp, vp, vp, i64, i32, vp]),
    "pcgan_gather_cast_bf16_batched": (C.c_int, [vp, i32, i64, vp]),
    "pcgan_scatter_f32_batched": (C.c_int, [vp, i32, i64, i32, vp]),
    "pcgan_pack_nchw": (C.c_int, [C.POINTER(PackArgs), vp]),
    "pcgan_resize_nchw_fwd": (C.c_int, [vp, vp, i64, i32, i32, i32, i32, vp]),
    "pcgan_resize_nchw_bwd": (C.c_int, [vp, vp, i64, i32, i32, i32, i32, vp]),
    "pcgan_unpack_resize_bwd": (C.c_int, [C.POINTER(UnpackArgs), vp]),
    "pcgan_norm_finalize": (C.c_int, [C.POINTER(NormFinalizeArgs), vp]),
    "pcgan_norm_apply": (C.c_int, [C.POINTER(NormApplyArgs), vp]),
    "pcgan_norm_running_batched": (C.c_int, [vp, i32, i32, vp]),
    "pcgan_halo_fold": (C.c_int, [C.POINTER(FoldArgs), vp]),
    "pcgan_halo_accumulate": (C.c_int, [C.POINTER(FoldArgs), vp]),
    "pcgan_norm_bwd_reduce": (C.c_int, [C.POINTER(NormBwdArgs), vp]),
    "pcgan_norm_bwd_apply": (C.c_int, [C.POINTER(NormBwdArgs), vp]),
    "pcgan_norm_bwd_fused_supported": (C.c_int, [C.POINTER(NormBwdArgs)]),
    "pcgan_norm_bwd_fused": (C.c_int, [C.POINTER(NormBwdArgs), vp]),
    "pcgan_norm_bwd_fused_active_clusters": (C.c_int, []),
    "pcgan_maxpool3x3s2_fwd": (C.c_int, [C.POINTER(MaxpoolArgs), vp]),
    "pcgan_maxpool3x3s2_bwd": (C.c_int, [vp, i32, vp, vp, i32, i32, i32, i32, i32, i32, vp]),
    "pcgan_act_bwd": (C.c_int, [vp, i32, vp, i32, vp, i32, i32, i32, i32, i32, f32, vp]),
    "pcgan_nhwc_cast": (C.c_int, [vp, vp, i32, i32, i32, i32, i32, i32, vp]),
    "pcgan_augment": (C.c_int, [C.POINTER(AugmentArgs), vp]),
    "pcgan_loss": (C.c_int, [C.POINTER(LossArgs), vp]),
    "pcgan_adam": (C.c_int, [vp, vp, vp, vp, i64, vp, f32, f32, f32, vp, vp]),
    "pcgan_adam_batched": (C.c_int, [vp, i32, i64, vp, C.c_double, C.c_double, C.c_double, vp, vp]),
}

_lib = None


class PcganError(RuntimeError):
    pass


def load():
    """Load the shared library (once) and check ABI version and struct layouts."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PcganError(
            "libpcgan_kernels.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C pcgan_b200/csrc`. There is no fallback path." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if lib.pcgan_abi_version() != ABI_VERSION:
        raise PcganError("ABI version mismatch: library %d, binding %d" % (lib.pcgan_abi_version(), ABI_VERSION))
    for name, cls in _STRUCTS.items():
        n = lib.pcgan_sizeof(name.encode())
        if n != C.sizeof(cls):
            raise PcganError("struct %s: library sizeof %d != binding %d" % (name, n, C.sizeof(cls)))
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().pcgan_last_error()
        raise PcganError("%s failed (%d): %s" % (what or "pcgan call", rc, msg.decode() if msg else "?"))
