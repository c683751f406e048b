"""ctypes binding of ``libinstaorder_b200.so`` (the C ABI declared in ``include/instaorder_b200.h``).

The product path has no CPU fallback: if the shared library is missing or a call fails, an exception is raised.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libinstaorder_b200.so")

IO_HEAD_OCC, IO_HEAD_DEPTH, IO_HEAD_ORDERNET = 1, 2, 3
IO_ERR_DEGENERATE = -3


class IoError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libinstaorder_b200 error %d: %s" % (code, msg))
        self.code = code


class PairDesc(C.Structure):
    """``io_pair_desc`` (48 bytes)."""
    _fields_ = [("image_off", C.c_int64), ("mask_a_off", C.c_int64), ("mask_b_off", C.c_int64),
                ("h", C.c_int32), ("w", C.c_int32), ("x", C.c_int32), ("y", C.c_int32), ("s", C.c_int32),
                ("rgb_slot", C.c_int32)]


# numpy structured dtype with the same layout, so descriptors can be built vectorised
PAIR_DESC_DTYPE = [("image_off", "<i8"), ("mask_a_off", "<i8"), ("mask_b_off", "<i8"), ("h", "<i4"), ("w", "<i4"),
                   ("x", "<i4"), ("y", "<i4"), ("s", "<i4"), ("rgb_slot", "<i4")]

_vp, _i, _i64 = C.c_void_p, C.c_int, C.c_int64
_SIGS = {
    "io_abi_version": (_i, []),
    "io_last_error": (C.c_char_p, []),
    "io_pair_enumerate": (_i, [_i, _vp]),
    "io_expand_bbox": (_i, [_vp, _i, C.c_double, _vp]),
    "io_pair_crop_boxes": (_i, [_vp, _vp, _i, _vp]),
    "io_pair_bordering": (_i, [_vp, _i, _i, _i, _vp, _i, _vp, _vp]),
    "io_infer_gt_order": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _vp, _vp]),
    "io_mask_stats": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "io_pair_tensor_row_pitch": (_i64, [_i]),
    "io_pair_tensor_bytes": (_i64, [_i, _i]),
    "io_pair_gather_patch": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "io_image_resize_rgb": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "io_image_square_linear_rgb": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "io_image_resize_linear_rgb": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "io_pair_gather_resize": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp]),
    "io_normalize_lut": (_i, [_vp, _vp, _vp]),
    "io_net_create": (_i, [_vp, _i, _i, _i, C.POINTER(_vp)]),
    "io_net_create_arch": (_i, [_vp, _vp, _vp, _i, _i, _vp, _i, _i, _i, C.POINTER(_vp)]),
    "io_net_feature": (_i, [_vp, _i, C.POINTER(_vp), C.POINTER(_i64)]),
    "io_net_set_inject": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "io_net_destroy": (_i, [_vp]),
    "io_net_load_state": (_i, [_vp, _vp, _vp, _i]),
    "io_net_forward_pairs": (_i, [_vp, _vp, _i, _vp, _vp]),
    "io_net_last_launches": (_i, [_vp]),
    "io_pair_tensor_bytes_hw": (_i64, [_i, _i, _i]),
    "io_image_resize_rgb_hw": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "io_pair_gather_resize_hw": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "io_net_forward_pairs_hw": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    "io_net_profile": (_i, [_vp, _i]),
    "io_net_profile_read": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i]),
    "io_order_decide": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "io_rle_from_string": (_i, [C.c_char_p, _i64, _vp, _i, _vp]),
    "io_rle_from_polygon": (_i, [_vp, _i, _i, _i, _vp, _i, _vp]),
    "io_masks_from_rle": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "io_add_relu": (_i, [_vp, _vp, _vp, _i64, _i, _vp]),
    "io_upsample2x_bilinear": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "io_conv_bn_act": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "io_stem_pool": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp]),
    "io_conv_dual": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _i, _vp, _vp, _i, _i, _vp, _vp]),
    "io_conv_fused_dual": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp]),
    "io_conv_fused_pair": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp]),
    "io_pair_pack_nchw": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp]),
    "io_loss_forward": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, C.c_float, C.c_float, _i, _vp, _vp]),
    "io_metrics_prf": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _vp]),
    "io_metrics_whdr": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp]),
}
_f = C.c_float
_SIGS.update({
    "io_train_create": (_i, [_vp, _i, _i, _i, C.POINTER(_vp)]),
    "io_train_destroy": (_i, [_vp]),
    "io_train_param_count": (_i64, [_vp]),
    "io_train_stat_count": (_i64, [_vp]),
    "io_train_num_segments": (_i, [_vp]),
    "io_train_segment": (_i, [_vp, _i, C.c_char_p, _i, _vp, _vp, _vp]),
    "io_train_bind": (_i, [_vp, _vp, _vp, _vp]),
    "io_train_sync_weights": (_i, [_vp, _vp]),
    "io_train_forward_backward": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _f, _f, _i, _vp, _i, _vp]),
    "io_train_num_buckets": (_i, [_vp]),
    "io_train_bucket": (_i, [_vp, _i, _vp, _vp]),
    "io_train_wait_bucket": (_i, [_vp, _i, _vp]),
    "io_train_sgd_step": (_i, [_vp, _vp, _f, _f, _f, _i, _vp]),
    "io_train_adam_step": (_i, [_vp, _vp, _vp, _f, _f, _f, _f, _i, _vp]),
    "io_train_read_logits": (_i, [_vp, _vp, _vp]),
    "io_train_read_activation": (_i, [_vp, C.c_char_p, _i, _vp, _vp, _vp]),
    "io_train_last_launches": (_i, [_vp]),
    "io_train_profile": (_i, [_vp, _i]),
    "io_train_profile_read": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i]),
    "io_optim_sgd": (_i, [_vp, _vp, _vp, _i64, _f, _f, _f, _i, _vp, _i64, _vp]),
    "io_optim_adam": (_i, [_vp, _vp, _vp, _vp, _i64, _f, _f, _f, _f, _i, _vp, _i64, _vp]),
    "io_conv_wgrad": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _i, _i, _vp, _vp]),
    "io_conv_dgrad": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "io_stem_wgrad": (_i, [_vp, _i, _i, _vp, _vp, _vp]),
    "io_bn_train_forward": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp, _i, _vp, _vp]),
    "io_bn_train_backward": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "io_maxpool_train": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
})
# io_net_load_state takes (net, names, ptrs, numels, n)
_SIGS["io_net_load_state"] = (_i, [_vp, _vp, _vp, _vp, _i])

EXPORTS = sorted(_SIGS)

_lib = None


def lib():
    """Loads the shared library once.  Raises if it has not been built (``python -c 'import __graft_entry__ as g;
    g.build()'`` or ``make -C instaorder_b200/csrc``)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise ImportError("%s is missing: build it with `make -C instaorder_b200/csrc` "
                              "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error():
    return lib().io_last_error().decode("utf-8", "replace")


def check(rc):
    if rc < 0:
        raise IoError(rc, last_error())
    return rc


def ptr(t):
    """Raw pointer of a torch tensor / numpy array (must be contiguous) or None."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        assert t.is_contiguous()
        return t.data_ptr()
    assert t.flags["C_CONTIGUOUS"]
    return t.ctypes.data


def stream_ptr(stream=None):
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return s.cuda_stream
