"""ctypes binding of liblssvc_b200.so (the C-ABI declared in include/lssvc_b200.h).

There is no fallback: if the shared library is missing, or a call reports an error, this raises.
"""
import ctypes
import os
from ctypes import POINTER, Structure, byref, c_char_p, c_double, c_float, c_int32, c_int64, c_uint8, c_uint32, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "liblssvc_b200.so")

MAX_SRC = 3
ACT_NONE, ACT_LRELU = 0, 1
IN_NONE, IN_SQUARE, IN_LRELU = 0, 1, 2
EPI_PLAIN, EPI_GDN, EPI_IGDN, EPI_LAPLACE, EPI_BITPARM, EPI_FOURPART = 0, 1, 2, 3, 4, 5
PREC_TF32, PREC_3XTF32, PREC_H2 = 0, 1, 2


class LssvcError(RuntimeError):
    pass


class CView(Structure):
    _fields_ = [("ptr", c_void_p), ("H", c_int32), ("W", c_int32), ("C", c_int32), ("pitch", c_int32)]


class CConv(Structure):
    _fields_ = [
        ("n_src", c_int32),
        ("src", CView * MAX_SRC),
        ("weight", c_void_p),
        ("bias", c_void_p),
        ("kh", c_int32), ("kw", c_int32), ("stride", c_int32), ("pad", c_int32),
        ("cout", c_int32), ("n_pad", c_int32), ("cin_total", c_int32),
        ("in_transform", c_int32),
        ("in_slope", c_float),
        ("epi", c_int32),
        ("act", c_int32),
        ("slope", c_float),
        ("out_scale", c_float),
        ("pixel_shuffle", c_int32),
        ("out", CView),
        ("res1", CView), ("res2", CView),
        ("out2", CView),
        ("slope2", c_float),
        ("gdn_x", CView),
        ("precision", c_int32),
        ("weight_split", c_void_p),
        ("weight_h2", c_void_p),
        ("cin_pad16", c_int32),
        ("acc_scale", c_float),
        ("ent_y", CView), ("ent_y_hat", CView),
        ("ent_coef", c_void_p),
        ("ent_bits", c_void_p),
        ("ent_sym", c_void_p), ("ent_index", c_void_p),
        ("ent_thr", c_void_p),
        ("ent_n_thr", c_int32),
        ("ent_tile", c_int32),
        ("ent_step", c_int32),
    ]


class CFfn(Structure):
    _fields_ = [("inp", CView), ("out", CView), ("res2", CView), ("hidden", c_int32), ("w1", c_void_p), ("w2", c_void_p),
                ("b1", c_void_p), ("b2", c_void_p), ("scale1", c_float), ("scale2", c_float), ("slope1", c_float),
                ("slope2", c_float)]


class CPw(Structure):
    _fields_ = [("inp", CView), ("out", CView), ("res1", CView), ("res2", CView), ("w", c_void_p), ("bias", c_void_p),
                ("dw_weight", c_void_p), ("dw_bias", c_void_p), ("act", c_int32), ("slope", c_float), ("out_scale", c_float),
                ("acc_scale", c_float)]


_PV = POINTER(CView)
_SIGNATURES = {
    # name: (restype, argtypes)
    "lssvc_abi_version": (c_int32, []),
    "lssvc_device_check": (c_int32, [c_int32]),
    "lssvc_last_error": (c_char_p, []),
    "lssvc_range_flag_fetch": (c_int32, [c_void_p, c_void_p]),
    "lssvc_launch_count": (c_int64, []),
    "lssvc_launch_count_add": (None, [c_int64]),
    "lssvc_conv_hs": (c_int32, [POINTER(CConv), c_void_p]),
    "lssvc_conv_ffn": (c_int32, [POINTER(CFfn), c_void_p]),
    "lssvc_conv_pw": (c_int32, [POINTER(CPw), c_void_p]),
    "lssvc_conv_simt": (c_int32, [POINTER(CConv), c_void_p]),
    "lssvc_conv_head": (c_int32, [POINTER(CConv), c_void_p]),
    "lssvc_conv_head_supported": (c_int32, [POINTER(CConv)]),
    "lssvc_dwconv3x3": (c_int32, [_PV, c_void_p, c_void_p, _PV, c_void_p]),
    "lssvc_deconv3x3_s2": (c_int32, [_PV, c_void_p, c_void_p, c_int32, c_float, _PV, c_void_p]),
    "lssvc_nchw_to_nhwc": (c_int32, [c_void_p, c_int32, _PV, c_void_p]),
    "lssvc_nhwc_to_nchw": (c_int32, [_PV, c_void_p, c_void_p]),
    "lssvc_lrelu_copy": (c_int32, [_PV, c_float, _PV, c_void_p]),
    "lssvc_softmax2_blend": (c_int32, [_PV, _PV, _PV, _PV, c_void_p]),
    "lssvc_flow_warp": (c_int32, [_PV, _PV, c_float, _PV, c_void_p]),
    "lssvc_bilinear_resize": (c_int32, [_PV, c_float, _PV, c_void_p]),
    "lssvc_avgpool2": (c_int32, [_PV, _PV, c_void_p]),
    "lssvc_maxpool2": (c_int32, [_PV, _PV, c_void_p]),
    "lssvc_spynet_prep": (c_int32, [_PV, _PV, _PV, _PV, _PV, c_void_p]),
    "lssvc_offset_diversity": (c_int32, [_PV, _PV, _PV, c_void_p, c_void_p, c_int32, c_int32, c_float, _PV, c_void_p, c_void_p]),
    "lssvc_laplace_quant": (c_int32, [_PV, _PV, _PV, _PV, _PV, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p]),
    "lssvc_four_part_step": (c_int32, [_PV, _PV, c_int32, _PV, _PV, _PV, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p]),
    "lssvc_four_part_index": (c_int32, [_PV, c_int32, c_void_p, c_void_p, c_int32, c_void_p]),
    "lssvc_four_part_dec_step": (c_int32, [c_void_p, _PV, c_int32, _PV, c_void_p]),
    "lssvc_scale_index": (c_int32, [_PV, c_void_p, c_void_p, c_int32, c_void_p]),
    "lssvc_symbols_to_view": (c_int32, [c_void_p, _PV, _PV, c_void_p]),
    "lssvc_gaussian_quant": (c_int32, [_PV, _PV, _PV, _PV, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p]),
    "lssvc_bitparm_quant": (c_int32, [_PV, c_void_p, _PV, c_void_p, c_void_p, c_void_p]),
    "lssvc_eb_quant": (c_int32, [_PV, c_void_p, _PV, c_void_p, c_void_p, c_void_p]),
    "lssvc_sse": (c_int32, [_PV, _PV, c_void_p, c_void_p]),
    "lssvc_yuv420_to_rgb": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_int32, c_int32, c_void_p]),
    "lssvc_resample_1d": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_int32, c_int32, c_void_p,
                                    c_int32, c_void_p]),
    "lssvc_sse_flat": (c_int32, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "lssvc_pmf_to_quantized_cdf": (c_int32, [POINTER(c_float), c_int32, c_int32, POINTER(c_uint32)]),
    "lssvc_rans_encoder_new": (c_void_p, []),
    "lssvc_rans_encoder_free": (None, [c_void_p]),
    "lssvc_rans_encoder_reset": (None, [c_void_p]),
    "lssvc_rans_encode_with_indexes": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int32, c_int32, c_void_p, c_void_p]),
    "lssvc_rans_encoder_flush": (c_int64, [c_void_p, POINTER(POINTER(c_uint8))]),
    "lssvc_rans_decoder_new": (c_void_p, []),
    "lssvc_rans_decoder_free": (None, [c_void_p]),
    "lssvc_rans_decoder_set_stream": (c_int32, [c_void_p, c_char_p, c_int64]),
    "lssvc_rans_decode_stream": (c_int32, [c_void_p, c_void_p, c_int64, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
}

EXPORTS = tuple(_SIGNATURES)
_lib = None


def load():
    """Load the shared library once; raise if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LssvcError(
            f"{LIB_PATH} is missing: run `python -m lssvc_b200.build` (or __graft_entry__.build()); "
            "there is no CPU or PyTorch fallback for the coding path")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


# Host-logic validation without a GPU (tests/test_dry_run.py): every call still goes through the library's
# argument checks on CPU-resident buffers; only the "no CUDA device / launch failed" outcomes are tolerated.
DRY_RUN = False


def check(rc, what=""):
    if DRY_RUN and rc in (-2, -3):
        return
    if rc != 0:
        msg = load().lssvc_last_error().decode("utf-8", "replace")
        raise LssvcError(f"{what or 'lssvc call'} failed ({rc}): {msg}")


def launch_count():
    return int(load().lssvc_launch_count())
