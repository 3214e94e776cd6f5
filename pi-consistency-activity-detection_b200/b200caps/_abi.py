"""ctypes binding of libb200caps.so (include/b200caps.h).  No math lives here.

The library is REQUIRED: importing any compute entry point without the built shared object
raises -- there is no CPU / PyTorch fallback for the hot path.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200caps.so")

i32, i64, f32, vp = C.c_int32, C.c_int64, C.c_float, C.c_void_p


class ConvClass(C.Structure):
    _fields_ = [("taps", vp), ("w", vp), ("ntaps", i32), ("Qt", i32), ("Qh", i32), ("Qw", i32),
                ("po_t", i32), ("po_h", i32), ("po_w", i32), ("lo_t", i32), ("lo_h", i32), ("lo_w", i32),
                ("h_block", i32), ("pad_", i32)]


class ConvDesc(C.Structure):
    _fields_ = [("inp", vp), ("out", vp), ("bias", vp), ("scale_nc", vp), ("stat_sums", vp),
                ("in_row_stride", i64), ("out_row_stride", i64),
                ("in_c_off", i32), ("out_c_off", i32), ("Cin", i32), ("Cout", i32),
                ("N", i32), ("Ti", i32), ("Hi", i32), ("Wi", i32), ("To", i32), ("Ho", i32), ("Wo", i32),
                ("si_t", i32), ("si_h", i32), ("si_w", i32), ("so_t", i32), ("so_h", i32), ("so_w", i32),
                ("out_fp32", i32), ("relu", i32), ("sigmoid_from", i32), ("accumulate", i32), ("bn_tile", i32),
                ("tap_pitch", i32), ("dtype", i32), ("stat_groups", i32), ("round_out", i32),
                ("nclass", i32), ("out_fold", i32), ("w_sample_stride", i64), ("cls", ConvClass * 8)]


class WgradDesc(C.Structure):
    _fields_ = [("g", vp), ("p", vp), ("dw", vp), ("taps", vp), ("wtap", vp), ("taps_host", vp),
                ("g_row_stride", i64), ("p_row_stride", i64), ("s_p", i64), ("s_g", i64),
                ("g_c_off", i32), ("p_c_off", i32), ("Cg", i32), ("Cp", i32), ("Cg_real", i32), ("Cp_real", i32),
                ("N", i32), ("Tg", i32), ("Hg", i32), ("Wg", i32), ("Tp", i32), ("Hp", i32), ("Wp", i32),
                ("Qt", i32), ("Qh", i32), ("Qw", i32),
                ("sg_t", i32), ("sg_h", i32), ("sg_w", i32), ("sp_t", i32), ("sp_h", i32), ("sp_w", i32),
                ("pp_t", i32), ("pp_h", i32), ("pp_w", i32),
                ("ntaps", i32), ("bn_tile", i32), ("nsplit", i32), ("atomic", i32), ("dtype", i32),
                ("p_fold", i32), ("pad1_", i32), ("dw_sample_stride", i64),
                ("seg_dw", vp * 4), ("seg_begin", i32 * 4), ("nseg", i32), ("pad2_", i32)]


class PackJob(C.Structure):
    _fields_ = [("w", vp), ("packed", vp), ("wtap", vp), ("s_r", i64), ("s_c", i64), ("tap_pitch", i64), ("col_off", i64),
                ("R", i32), ("ntaps", i32), ("C", i32), ("C_real", i32), ("r_off", i32), ("bn_tile", i32), ("nkb", i32),
                ("dtype", i32)]


# name -> argtypes  (restype is always int unless noted)
_SIGS = {
    "b2c_conv_fprop": [C.POINTER(ConvDesc), vp],
    "b2c_conv_wgrad": [C.POINTER(WgradDesc), vp],
    "b2c_pack_weights": [vp, vp, vp, i32, i32, i32, i32, i64, i64, i64, i64, i32, i32, i32, vp],
    "b2c_pack_weights_batched": [vp, vp, i32, i32, vp],
    "b2c_ncdhw_to_ndhwc": [vp, vp, i32, i32, i64, i32, vp],
    "b2c_ndhwc_to_ncdhw_f32": [vp, i64, i32, vp, i32, i32, i64, vp],
    "b2c_u8_clip_to_cl": [vp, vp, i32, i32, i64, i32, vp],
    "b2c_u8_to_f32": [vp, vp, i64, f32, vp],
    "b2c_im2col_small": [vp, vp] + [i32] * 19 + [vp],
    "b2c_bn_sums": [vp, i64, i32, i64, i32, i32, vp, vp],
    "b2c_bn_sums_finalize": [vp, i64, i32, i64, i32, i32, vp, vp, vp, vp, vp, f32, f32, vp],
    "b2c_bn_finalize": [vp, i32, i32, i32, i32, i64, vp, vp, vp, vp, f32, f32, vp],
    "b2c_bn_relu_apply": [vp, i64, i32, i64, i32, i32, vp, vp, vp, vp, vp, i64, i32, i32, vp],
    "b2c_bn_relu_bwd_reduce": [vp, i64, i32, vp, i64, i32, vp, i64, i32, i64, i32, i32, vp, vp, vp, vp, vp, i32, vp],
    "b2c_bn_relu_bwd_apply": [vp, i64, i32, vp, i64, i32, vp, i64, i32, i64, i32, i32, vp, vp, vp, vp, vp, vp, i64, i32,
                              vp, vp, i32, vp],
    "b2c_maxpool_fwd": [vp, i64, i32, vp, i64, i32, vp] + [i32] * 17 + [vp],
    "b2c_maxpool_bwd": [vp, i64, i32, vp, vp, i64, i32] + [i32] * 18 + [vp],
    "b2c_channel_scale": [vp, i64, i32, vp, vp, i64, i32, i32, i64, i32, vp],
    "b2c_act_bwd": [vp, i64, i32, vp, i64, i32, vp, vp, i64, i32, vp, i32, i64, i32, i32, vp],
    "b2c_add": [vp, i64, i32, vp, i64, i32, vp, i64, i32, i64, i32, vp],
    "b2c_stencil27_fwd": [vp, vp, vp, i32, i32, i32, i32, vp],
    "b2c_stencil27_bwd": [vp, vp, vp, i32, i32, i32, i32, i32, vp],
    "b2c_em_routing_fwd": [vp, vp, vp, vp, vp, i64, i32, vp],
    "b2c_em_routing_bwd": [vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, i32, vp],
    "b2c_em_routing_fwd_train": [vp, vp, vp, vp, vp, vp, i64, i32, vp],
    "b2c_em_routing_bwd_state": [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, i32, vp],
    "b2c_primarycaps_finish": [vp, i32, vp, vp, i32, i32, vp],
    "b2c_primarycaps_bwd_prep2": [vp, vp, vp, vp, vp, i32, i32, i32, i32, vp],
    "b2c_rows_to_clips": [vp, vp, i64, i32, i32, i32, i32, i32, vp],
    "b2c_clips_to_rows": [vp, i64, i32, vp, i32, i32, i32, i32, vp],
    "b2c_primarycaps_bwd_prep": [vp, vp, vp, vp, i64, i32, vp],
    "b2c_class_mean_fwd": [vp, vp, i32, i32, i32, vp],
    "b2c_pose_mask_fwd": [vp, vp, vp, i32, i32, i32, vp],
    "b2c_caps_head_bwd": [vp, vp, vp, vp, vp, i32, i32, i32, vp],
    "b2c_seg_loss_fwd": [vp, vp, vp, i32, i64, vp, vp, vp],
    "b2c_seg_loss_bwd": [vp, vp, vp, i32, i64, vp, f32, f32, vp, vp],
    "b2c_spread_loss": [vp, vp, vp, i32, i32, f32, vp, f32, vp, vp],
    "b2c_bv_mask": [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp],
    "b2c_gv_mask": [vp, vp, vp, i32, i32, i32, f32, f32, i32, i32, vp],
    "b2c_cons_reduce": [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp],
    "b2c_cons_finish": [vp, vp, i32, i32, i32, i32, f32, f32, f32, vp, vp],
    "b2c_cons_grad": [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, f32, f32, f32, vp, vp],
    "b2c_adam_step": [vp, vp, vp, vp, i64, f32, f32, f32, f32, vp, f32, vp, vp],
    "b2c_fill_f32": [vp, i64, f32, vp],
    "b2c_set_deterministic": [i32],
    "b2c_set_pool_generic": [i32],
    "b2c_set_precision": [i32],
    "b2c_stem_fold_input": [vp, vp, i32, i32, i32, i32, i32, i32, i32, vp],
    "b2c_clips_to_folded": [vp, i32, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp],
    "b2c_stem_fold_weights": [vp, vp, i32, i32, i32, i32, i32, i32, i32, vp],
    "b2c_stem_unfold_wgrad": [vp, vp, i32, i32, i32, i32, i32, i32, i32, vp],
    "b2c_frame_iou_counts": [vp, vp, vp, i64, i32, vp],
    "b2c_split_bf16": [vp, i64, i32, vp, vp, i64, i32, vp],
    "b2c_tail_weff": [vp, vp, vp, vp, vp, i64, vp, i64, i32, vp, i32, vp],
    "b2c_tail_gather_fwd": [vp, vp, vp, vp, i32, i32, i32, i32, vp],
    "b2c_tail_gather_bwd": [vp, vp, vp, i32, i32, i32, i32, vp],
    "b2c_tail_chain_bwd": [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp],
}

EXPORTS = sorted(list(_SIGS) + ["b2c_last_error", "b2c_version", "b2c_launch_count", "b2c_get_precision",
                                "b2c_em_routing_state_floats"])

_lib = None


def lib():
    """Load (once) and return the shared library; raise loudly when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                f"b200caps: {LIB_PATH} is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU or PyTorch fallback for this path.")
        L = C.CDLL(LIB_PATH)
        for name, args in _SIGS.items():
            fn = getattr(L, name, None)
            if fn is None:          # reported by tests/test_abi.py; calling it raises below
                continue
            fn.argtypes = args
            fn.restype = C.c_int
        L.b2c_last_error.restype = C.c_char_p
        L.b2c_last_error.argtypes = []
        L.b2c_version.restype = C.c_int
        L.b2c_launch_count.restype = C.c_longlong
        L.b2c_get_precision.restype = C.c_int
        L.b2c_get_precision.argtypes = []
        L.b2c_em_routing_state_floats.restype = C.c_int64
        L.b2c_em_routing_state_floats.argtypes = []
        _lib = L
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().b2c_last_error().decode(errors="replace")
        raise RuntimeError(f"b200caps {what} failed (rc={rc}): {msg}")


def call(name: str, *args):
    fn = getattr(lib(), name, None)
    if fn is None:
        raise RuntimeError(f"b200caps: libb200caps.so does not export {name}; rebuild the extension")
    check(fn(*args), name)


def launch_count() -> int:
    return int(lib().b2c_launch_count())
