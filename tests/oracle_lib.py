"""ctypes doors onto the test-infrastructure libraries (oracle restatement and, when present,
the compiled unmodified reference).  Imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs ONLY -- never by the product package."""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
YUV_AUTO, YUV_420, YUV_SHARP, YUV_444, YUV_400 = 0, 1, 2, 3, 4

_u8p = C.POINTER(C.c_uint8)


class SjoParams(C.Structure):
    _fields_ = [("yuv_mode", C.c_int), ("method", C.c_int), ("pix_fmt", C.c_int),
                ("quant", (C.c_uint8 * 64) * 2), ("min_quant", (C.c_uint8 * 64) * 2),
                ("q_bias", C.c_int), ("qdelta_max_luma", C.c_int), ("qdelta_max_chroma", C.c_int)]


def build_oracle():
    """(Re)build oracle/liboracle.so and, where /root/reference exists, oracle/_ref."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True, capture_output=True)


_oracle = None
_ref = None


def oracle():
    global _oracle
    if _oracle is None:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            build_oracle()
        L = C.CDLL(path)
        L.sjo_encode.restype = C.c_size_t
        L.sjo_encode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(SjoParams),
                                 C.POINTER(_u8p)]
        L.sjo_sjpeg_encode.restype = C.c_size_t
        L.sjo_sjpeg_encode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int,
                                       C.c_int, C.POINTER(_u8p)]
        L.sjo_free.argtypes = [_u8p]
        L.sjo_default_params.argtypes = [C.POINTER(SjoParams), C.c_float, C.c_int, C.c_int]
        L.sjo_make_rgb.argtypes = [C.c_char, C.c_int, C.c_int, C.c_uint32, C.c_void_p]
        L.sjo_image_to_coeffs.argtypes = [C.c_void_p] + [C.c_int] * 5 + [C.c_void_p]
        L.sjo_image_to_samples.argtypes = [C.c_void_p] + [C.c_int] * 5 + [C.c_void_p]
        L.sjo_quantize_image.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                         C.c_int, C.c_void_p]
        L.sjo_trellis_quantize_image.argtypes = L.sjo_quantize_image.argtypes
        L.sjo_collect_histograms.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.sjo_analyse_histo.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.sjo_symbol_stats.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.sjo_build_optimal_table.restype = C.c_int
        L.sjo_build_optimal_table.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.sjo_quality_to_matrices.argtypes = [C.c_float, C.c_void_p]
        L.sjo_geometry.argtypes = [C.c_int] * 3 + [C.POINTER(C.c_int)] * 5
        L.sjo_encode_planar.restype = C.c_size_t
        L.sjo_encode_planar.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                        C.c_int, C.c_int, C.POINTER(SjoParams), C.POINTER(_u8p)]
        L.sjo_sharp_yuv.restype = C.c_int
        L.sjo_sharp_yuv.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.sjo_riskiness.restype = C.c_int
        L.sjo_riskiness.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_float)]
        _oracle = L
    return _oracle


def ref():
    """The compiled unmodified reference, or None when oracle/_ref was not shipped/built."""
    global _ref
    if _ref is None:
        path = os.path.join(ORACLE_DIR, "_ref", "libsjpeg_ref.so")
        if not os.path.exists(path):
            if os.path.isdir("/root/reference/src"):
                build_oracle()
            if not os.path.exists(path):
                return None
        L = C.CDLL(path)
        L.SjpegEncode.restype = C.c_size_t
        L.SjpegEncode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(_u8p), C.c_float,
                                  C.c_int, C.c_int]
        L.SjpegFreeBuffer.argtypes = [_u8p]
        L.ref_encode_planar.restype = C.c_size_t
        L.ref_encode_planar.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                        C.c_int, C.c_int, C.c_float] + [C.c_int] * 3 + [C.POINTER(_u8p)]
        L.ref_encode_param.restype = C.c_size_t
        L.ref_encode_param.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                       C.c_float] + [C.c_int] * 6 + [C.POINTER(_u8p)]
        L.ref_sharp_yuv.restype = None
        L.ref_sharp_yuv.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_score_table.restype = C.c_void_p
        L.ref_score_table.argtypes = [C.POINTER(C.c_size_t)]
        L.SjpegRiskiness.restype = C.c_int
        L.SjpegRiskiness.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]
        _ref = L
    return _ref


def make_rgb(gen, w, h, seed=7654321):
    out = np.empty((h, w, 3), dtype=np.uint8)
    oracle().sjo_make_rgb(gen.encode(), w, h, seed, out.ctypes.data)
    return out


def _base_ptr(img, stride):
    """address of row 0 for a (possibly negative) stride over a contiguous buffer"""
    return img.ctypes.data


def oracle_encode(rgb, w, h, stride, quality, method, yuv_mode, base=None):
    out = _u8p()
    ptr = base if base is not None else rgb.ctypes.data
    n = oracle().sjo_sjpeg_encode(ptr, w, h, stride, quality, method, yuv_mode, C.byref(out))
    if n == 0:
        return None
    data = C.string_at(out, n)
    oracle().sjo_free(out)
    return data


def oracle_encode_params(rgb, w, h, stride, params, base=None):
    out = _u8p()
    ptr = base if base is not None else rgb.ctypes.data
    n = oracle().sjo_encode(ptr, w, h, stride, C.byref(params), C.byref(out))
    if n == 0:
        return None
    data = C.string_at(out, n)
    oracle().sjo_free(out)
    return data


def ref_encode(rgb, w, h, stride, quality, method, yuv_mode, base=None):
    out = _u8p()
    ptr = base if base is not None else rgb.ctypes.data
    n = ref().SjpegEncode(ptr, w, h, stride, C.byref(out), quality, method, yuv_mode)
    if n == 0:
        return None
    data = C.string_at(out, n)
    ref().SjpegFreeBuffer(out)
    return data


def md5(data):
    return hashlib.md5(data).hexdigest().upper()


# planar kinds of oracle/ref_shim.cc::ref_encode_planar
KIND_YUV420, KIND_YUV444, KIND_NV12, KIND_NV21, KIND_GRAY = 0, 1, 2, 3, 4
KIND_MODE = {KIND_YUV420: YUV_420, KIND_YUV444: YUV_444, KIND_NV12: YUV_420, KIND_NV21: YUV_420, KIND_GRAY: YUV_400}


def make_planes(kind, w, h, seed=5, pad=(5, 3, 7)):
    """Random planes with padded strides for a planar kind.  Returns dict(y, u, v: arrays or None)."""
    rng = np.random.RandomState(seed)
    cw, ch = (w + 1) // 2, (h + 1) // 2
    Y = rng.randint(0, 256, (h, w + pad[0])).astype(np.uint8)
    if kind == KIND_GRAY:
        return {"y": Y, "u": None, "v": None}
    if kind == KIND_YUV420:
        return {"y": Y, "u": rng.randint(0, 256, (ch, cw + pad[1])).astype(np.uint8),
                "v": rng.randint(0, 256, (ch, cw + pad[2])).astype(np.uint8)}
    if kind == KIND_YUV444:
        return {"y": Y, "u": rng.randint(0, 256, (h, w + pad[1])).astype(np.uint8),
                "v": rng.randint(0, 256, (h, w + pad[2])).astype(np.uint8)}
    return {"y": Y, "u": rng.randint(0, 256, (ch, 2 * cw + pad[1])).astype(np.uint8), "v": None}   # interleaved


def planar_args(kind, planes):
    """(y, ys, u, us, v, vs, uv_step) as addresses / strides for sjo_encode_planar and sjb_encode_planar."""
    y, u, v = planes["y"], planes["u"], planes["v"]
    if kind == KIND_GRAY:
        return y.ctypes.data, y.strides[0], None, 0, None, 0, 1
    if kind in (KIND_NV12, KIND_NV21):
        base = u.ctypes.data
        up, vp = (base, base + 1) if kind == KIND_NV12 else (base + 1, base)
        return y.ctypes.data, y.strides[0], up, u.strides[0], vp, u.strides[0], 2
    return y.ctypes.data, y.strides[0], u.ctypes.data, u.strides[0], v.ctypes.data, v.strides[0], 1


def oracle_encode_planar(kind, planes, w, h, quality, method):
    p = SjoParams()
    oracle().sjo_default_params(C.byref(p), float(quality), method, KIND_MODE[kind])
    out = _u8p()
    n = oracle().sjo_encode_planar(*planar_args(kind, planes), w, h, C.byref(p), C.byref(out))
    if n == 0:
        return None
    data = C.string_at(out, n)
    oracle().sjo_free(out)
    return data


def ref_encode_planar(kind, planes, w, h, quality, huffman, adaptive, trellis):
    y, u, v = planes["y"], planes["u"], planes["v"]
    out = _u8p()
    n = ref().ref_encode_planar(kind, y.ctypes.data, y.strides[0], u.ctypes.data if u is not None else None,
                                u.strides[0] if u is not None else 0, v.ctypes.data if v is not None else None,
                                v.strides[0] if v is not None else 0, w, h, float(quality), huffman, adaptive, trellis,
                                C.byref(out))
    if n == 0:
        return None
    data = C.string_at(out, n)
    ref().SjpegFreeBuffer(out)
    return data


# ---- sharp YUV pre-pass and riskiness ----------------------------------------------------------
def _sharp(fn, rgb, w, h, stride, base=None):
    cw, ch = (w + 1) // 2, (h + 1) // 2
    y = np.zeros((h, w), np.uint8)
    u = np.zeros((ch, cw), np.uint8)
    v = np.zeros((ch, cw), np.uint8)
    fn(base if base is not None else rgb.ctypes.data, w, h, stride, y.ctypes.data, u.ctypes.data, v.ctypes.data)
    return y, u, v


def oracle_sharp_yuv(rgb, w, h, stride, base=None):
    return _sharp(oracle().sjo_sharp_yuv, rgb, w, h, stride, base)


def ref_sharp_yuv(rgb, w, h, stride, base=None):
    return _sharp(ref().ref_sharp_yuv, rgb, w, h, stride, base)


def score_table():
    """The reference's generated 343 x 343 riskiness table, read out of the compiled reference
    (None when oracle/_ref did not travel)."""
    if ref() is None:
        return None
    n = C.c_size_t()
    p = ref().ref_score_table(C.byref(n))
    return np.frombuffer(C.string_at(p, n.value), np.uint8).copy()


def oracle_riskiness(rgb, w, h, stride, table):
    risk = C.c_float()
    mode = oracle().sjo_riskiness(rgb.ctypes.data, w, h, stride, table.ctypes.data, C.byref(risk))
    return mode, risk.value


def ref_riskiness(rgb, w, h, stride):
    risk = C.c_float()
    mode = ref().SjpegRiskiness(rgb.ctypes.data, w, h, stride, C.byref(risk))
    return mode, risk.value
