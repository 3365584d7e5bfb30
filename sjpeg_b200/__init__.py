"""sjpeg_b200 -- Python door onto the B200 baseline-JPEG encode path.

The product is the native library ``libsjpeg_b200.so`` (hand-written sm_100a kernels + C++ host
code, built in-tree by ``sjpeg_b200/csrc/Makefile``) and its C ABI ``include/sjpeg_b200.h``.  This
module is only plumbing: a ctypes binding that mirrors the reference's call surface
(``SjpegEncode`` / ``sjpeg::Encode`` with an ``EncoderParam``; /root/reference/src/sjpeg.h:104,
187-292) for tests and benchmarks.  There is no CPU fallback: if the library or a CUDA device is
missing, calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# SJPEG_B200_LIB selects another build of the same library (kernel A/B experiments)
LIB_PATH = os.environ.get("SJPEG_B200_LIB") or os.path.join(_HERE, "libsjpeg_b200.so")

YUV_AUTO, YUV_420, YUV_SHARP, YUV_444, YUV_400 = 0, 1, 2, 3, 4
PIX_RGB, PIX_BGRA, PIX_RGBA = 0, 1, 2
OK, ERR_ARG, ERR_CUDA, ERR_NOMEM, ERR_CAPACITY = 0, -1, -2, -3, -4

_u8p = C.POINTER(C.c_uint8)


class Params(C.Structure):
    """sjb_params (include/sjpeg_b200.h)"""
    _fields_ = [("yuv_mode", C.c_int), ("method", C.c_int), ("pix_fmt", C.c_int),
                ("quant", (C.c_uint8 * 64) * 2), ("min_quant", (C.c_uint8 * 64) * 2),
                ("q_bias", C.c_int), ("qdelta_max_luma", C.c_int), ("qdelta_max_chroma", C.c_int)]


class SjpegB200Error(RuntimeError):
    pass


# every symbol include/sjpeg_b200.h declares, with its signature
_SIGNATURES = {
    "sjb_version": (C.c_uint32, []),
    "sjb_device_count": (C.c_int, []),
    "sjb_context_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "sjb_context_destroy": (None, [C.c_void_p]),
    "sjb_last_error": (C.c_char_p, [C.c_void_p]),
    "sjb_params_default": (None, [C.POINTER(Params), C.c_float, C.c_int, C.c_int]),
    "sjb_quality_to_matrices": (None, [C.c_float, C.c_void_p]),
    "sjb_max_output_size": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "sjb_encode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_longlong,
                             C.POINTER(Params), C.c_void_p, C.c_int, C.c_size_t, C.POINTER(C.c_size_t)]),
    "sjb_fetch_output": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t]),
    "sjb_context_set_search": (C.c_int, [C.c_void_p, C.c_void_p]),
    "sjb_sharp_yuv": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_int]),
    "sjb_set_score_table": (C.c_int, [C.c_void_p, C.c_size_t]),
    "sjb_has_score_table": (C.c_int, []),
    "sjb_riskiness": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_longlong,
                                C.POINTER(C.c_int), C.POINTER(C.c_float)]),
    "sjb_encode_planar": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_void_p,
                                    C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(Params), C.c_void_p,
                                    C.c_int, C.c_size_t, C.POINTER(C.c_size_t)]),
    "sjb_encode_batch": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int,
                                   C.c_longlong, C.POINTER(Params), C.POINTER(C.c_void_p), C.c_int,
                                   C.c_size_t, C.POINTER(C.c_size_t)]),
    "sjb_encode_planar_batch": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.c_longlong, C.POINTER(C.c_void_p),
                                          C.c_longlong, C.POINTER(C.c_void_p), C.c_longlong, C.c_int, C.c_int, C.c_int,
                                          C.c_int, C.POINTER(Params), C.POINTER(C.c_void_p), C.c_int, C.c_size_t,
                                          C.POINTER(C.c_size_t)]),
    "sjb_gather_frames": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_void_p, C.c_size_t,
                                    C.POINTER(C.c_size_t), C.c_int, C.POINTER(C.c_int)]),
    "sjb_host_alloc": (C.c_void_p, [C.c_size_t]),
    "sjb_host_alloc_wc": (C.c_void_p, [C.c_size_t]),
    "sjb_host_free": (None, [C.c_void_p]),
    "sjb_stage_coefficients": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_longlong,
                                         C.POINTER(Params), C.c_int, C.c_void_p, C.c_void_p]),
    "sjb_stage_histogram": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_longlong,
                                      C.POINTER(Params), C.c_void_p]),
    "sjb_stage_adapted_matrices": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_longlong,
                                             C.POINTER(Params), C.c_void_p]),
    "sjb_stage_symbol_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_longlong,
                                         C.POINTER(Params), C.c_void_p, C.c_void_p]),
    "sjb_last_stage_timings": (C.c_int, [C.c_void_p, C.POINTER(C.c_float * 6), C.POINTER(C.c_int)]),
    "sjb_last_timings": (C.c_int, [C.c_void_p, C.POINTER(C.c_float * 3)]),
    "sjb_bench_device": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.c_int, C.c_int,
                                   C.c_longlong, C.POINTER(Params), C.c_int, C.POINTER(C.c_float),
                                   C.POINTER(C.c_float), C.POINTER(C.c_size_t), C.POINTER(C.c_ulonglong)]),
    "sjb_bench_output": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "sjb_bench_f1": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_longlong,
                               C.POINTER(Params), C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    "sjb_stripes_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(Params), C.POINTER(C.c_void_p)]),
    "sjb_stripes_destroy": (None, [C.c_void_p]),
    "sjb_stripes_transform": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_longlong, C.c_void_p]),
    "sjb_stripes_code": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "sjb_stripes_finish": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_size_t,
                                     C.POINTER(C.c_size_t), C.c_void_p, C.c_void_p, C.c_void_p]),
    "sjb_comm_unique_id": (C.c_int, [C.c_void_p]),
    "sjb_comm_create": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "sjb_comm_destroy": (None, [C.c_void_p]),
    "sjb_stripe_rows": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "sjb_stripes_encode": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_longlong,
                                     C.POINTER(Params), C.POINTER(C.c_void_p), C.c_size_t, C.POINTER(C.c_size_t)]),
    "sjb_stripes_assemble": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t),
                                       C.POINTER(C.c_uint), C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "sjb_picture_header": (C.c_int, [C.POINTER(Params), C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                     C.POINTER(C.c_size_t)]),
    # drop-in C entry points (include/sjpeg.h)
    "SjpegVersion": (C.c_uint32, []),
    "SjpegEncode": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(_u8p), C.c_float, C.c_int, C.c_int]),
    "SjpegCompress": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int, C.c_float, C.POINTER(_u8p)]),
    "SjpegFreeBuffer": (None, [_u8p]),
    "SjpegQuantMatrix": (None, [C.c_float, C.c_bool, C.c_void_p]),
    "SjpegEstimateQuality": (C.c_float, [C.c_void_p, C.c_bool]),
    "SjpegDimensions": (C.c_bool, [C.c_void_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "SjpegFindQuantizer": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p]),
    "SjpegRiskiness": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]),
}

_lib = None


def lib():
    """Loads libsjpeg_b200.so (raises if it has not been built: run __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SjpegB200Error("%s is missing: build it with `make -C sjpeg_b200/csrc` "
                                 "(or __graft_entry__.build()); there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            f = getattr(L, name)     # AttributeError if the library lacks a declared symbol
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def exported_symbols():
    return sorted(_SIGNATURES)


def default_params(quality=75.0, method=0, yuv_mode=YUV_420):
    """What SjpegEncode(rgb, w, h, stride, &out, quality, method, yuv_mode) uses (api.cc:32-49)."""
    p = Params()
    lib().sjb_params_default(C.byref(p), float(quality), int(method), int(yuv_mode))
    return p


def _check(ctx, rc, what):
    if rc != OK:
        msg = lib().sjb_last_error(ctx) if ctx else b""
        raise SjpegB200Error("%s failed: rc=%d %s" % (what, rc, (msg or b"").decode()))


class Context:
    """One GPU context (stream + device scratch); not thread-safe, create one per thread."""

    def __init__(self, device=0):
        self._ctx = C.c_void_p()
        if lib().sjb_device_count() <= 0:
            raise SjpegB200Error("no CUDA device: the encode path has no CPU fallback")
        rc = lib().sjb_context_create(int(device), C.byref(self._ctx))
        if rc != OK:
            raise SjpegB200Error("sjb_context_create(device=%d) failed: rc=%d" % (device, rc))

    def close(self):
        if self._ctx:
            lib().sjb_context_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- whole encode -------------------------------------------------------------------------
    def encode(self, pix, width, height, stride, params, base=None):
        """pix: contiguous uint8 numpy array; base: address of row 0 if different (negative
        stride).  Returns the JPEG bytes, or None when the arguments are refused."""
        ptr = base if base is not None else pix.ctypes.data
        size = C.c_size_t(0)
        rc = lib().sjb_encode(self._ctx, ptr, 0, width, height, stride, C.byref(params), None, 0, 0,
                              C.byref(size))
        if rc == ERR_ARG:
            return None
        if rc != ERR_CAPACITY:
            _check(self._ctx, rc if rc != OK else ERR_CUDA, "sjb_encode")
        out = np.empty(size.value, dtype=np.uint8)
        _check(self._ctx, lib().sjb_fetch_output(self._ctx, out.ctypes.data, 0, out.nbytes), "sjb_fetch_output")
        return out.tobytes()

    def encode_planar(self, y, y_stride, u, u_stride, v, v_stride, uv_step, width, height, params):
        """y/u/v: addresses (ints) of the first sample of each plane, or None."""
        size = C.c_size_t(0)
        rc = lib().sjb_encode_planar(self._ctx, y, y_stride, u, u_stride, v, v_stride, uv_step, 0, width, height,
                                     C.byref(params), None, 0, 0, C.byref(size))
        if rc == ERR_ARG:
            return None
        if rc != ERR_CAPACITY:
            _check(self._ctx, rc if rc != OK else ERR_CUDA, "sjb_encode_planar")
        out = np.empty(size.value, dtype=np.uint8)
        _check(self._ctx, lib().sjb_fetch_output(self._ctx, out.ctypes.data, 0, out.nbytes), "sjb_fetch_output")
        return out.tobytes()

    def encode_into(self, ptr, on_device, width, height, stride, params, out_ptr, out_on_device, cap):
        size = C.c_size_t(0)
        rc = lib().sjb_encode(self._ctx, ptr, int(on_device), width, height, stride, C.byref(params),
                              out_ptr, int(out_on_device), cap, C.byref(size))
        _check(self._ctx, rc, "sjb_encode")
        return size.value

    def encode_batch(self, ptrs, on_device, width, height, stride, params, out_ptrs, out_on_device, cap):
        n = len(ptrs)
        a = (C.c_void_p * n)(*ptrs)
        o = (C.c_void_p * n)(*out_ptrs)
        sizes = (C.c_size_t * n)()
        rc = lib().sjb_encode_batch(self._ctx, n, a, int(on_device), width, height, stride, C.byref(params),
                                    o, int(out_on_device), cap, sizes)
        _check(self._ctx, rc, "sjb_encode_batch")
        return list(sizes)

    def encode_planar_batch(self, ys, y_stride, us, u_stride, vs, v_stride, uv_step, on_device, width, height, params,
                            out_ptrs, out_on_device, cap):
        """ys / us / vs: lists of plane addresses (us, vs may be None for 4:0:0)"""
        n = len(ys)
        ya = (C.c_void_p * n)(*ys)
        ua = (C.c_void_p * n)(*us) if us is not None else None
        va = (C.c_void_p * n)(*vs) if vs is not None else None
        o = (C.c_void_p * n)(*out_ptrs)
        sizes = (C.c_size_t * n)()
        rc = lib().sjb_encode_planar_batch(self._ctx, n, ya, y_stride, ua, u_stride, va, v_stride, uv_step, int(on_device),
                                           width, height, C.byref(params), o, int(out_on_device), cap, sizes)
        _check(self._ctx, rc, "sjb_encode_planar_batch")
        return list(sizes)

    # -- stage-level --------------------------------------------------------------------------
    def _nblocks(self, width, height, mode):
        mcu, mb = (16, 6) if mode == YUV_420 else ((8, 3) if mode == YUV_444 else (8, 1))
        return ((width + mcu - 1) // mcu) * ((height + mcu - 1) // mcu) * mb

    def coefficients(self, pix, width, height, stride, params, quantise, base=None):
        nb = self._nblocks(width, height, params.yuv_mode)
        coef = np.empty((nb, 64), dtype=np.int16)
        mask = np.zeros(nb, dtype=np.uint8)
        ptr = base if base is not None else pix.ctypes.data
        rc = lib().sjb_stage_coefficients(self._ctx, ptr, width, height, stride, C.byref(params),
                                          int(quantise), coef.ctypes.data, mask.ctypes.data)
        _check(self._ctx, rc, "sjb_stage_coefficients")
        return coef, mask

    def histogram(self, pix, width, height, stride, params):
        counts = np.zeros((2, 64, 129), dtype=np.int32)
        rc = lib().sjb_stage_histogram(self._ctx, pix.ctypes.data, width, height, stride, C.byref(params),
                                       counts.ctypes.data)
        _check(self._ctx, rc, "sjb_stage_histogram")
        return counts

    def adapted_matrices(self, pix, width, height, stride, params):
        """the matrices kernels A1 derive from the picture's histogram (natural order, clamped to min_quant)"""
        quant = np.zeros((2, 64), dtype=np.uint8)
        rc = lib().sjb_stage_adapted_matrices(self._ctx, pix.ctypes.data, width, height, stride, C.byref(params),
                                              quant.ctypes.data)
        _check(self._ctx, rc, "sjb_stage_adapted_matrices")
        return quant

    def symbol_stats(self, pix, width, height, stride, params):
        ac = np.zeros((2, 256), dtype=np.uint32)
        dc = np.zeros((2, 12), dtype=np.uint32)
        rc = lib().sjb_stage_symbol_stats(self._ctx, pix.ctypes.data, width, height, stride, C.byref(params),
                                          ac.ctypes.data, dc.ctypes.data)
        _check(self._ctx, rc, "sjb_stage_symbol_stats")
        return ac, dc

    def sharp_yuv(self, rgb, width, height, stride, base=None):
        """sjpeg::ApplySharpYUVConversion on the device: returns the (y, u, v) planes."""
        cw, ch = (width + 1) // 2, (height + 1) // 2
        y = np.zeros((height, width), np.uint8)
        u = np.zeros((ch, cw), np.uint8)
        v = np.zeros((ch, cw), np.uint8)
        ptr = base if base is not None else rgb.ctypes.data
        rc = lib().sjb_sharp_yuv(self._ctx, ptr, 0, width, height, stride, y.ctypes.data, u.ctypes.data,
                                 v.ctypes.data, 0)
        _check(self._ctx, rc, "sjb_sharp_yuv")
        return y, u, v

    def riskiness(self, rgb, width, height, stride, base=None):
        """SjpegRiskiness on the device: (recommended yuv mode, risk).  Needs set_score_table()."""
        mode, risk = C.c_int(0), C.c_float(0)
        ptr = base if base is not None else rgb.ctypes.data
        rc = lib().sjb_riskiness(self._ctx, ptr, 0, width, height, stride, C.byref(mode), C.byref(risk))
        _check(self._ctx, rc, "sjb_riskiness")
        return mode.value, risk.value

    def last_timings(self):
        ms = (C.c_float * 3)()
        lib().sjb_last_timings(self._ctx, C.byref(ms))
        return list(ms)

    def last_stage_timings(self):
        """({stage: ms}, pictures in the timed group); stages the method does not run are left out"""
        ms = (C.c_float * 6)()
        frames = C.c_int(0)
        lib().sjb_last_stage_timings(self._ctx, C.byref(ms), C.byref(frames))
        names = ["F1", "H1", "Q1/T1", "S1", "E", "S"]
        return {k: v for k, v in zip(names, ms) if v >= 0}, frames.value

    def bench_device(self, dev_ptrs, width, height, stride, params, iters):
        n = len(dev_ptrs)
        a = (C.c_void_p * n)(*dev_ptrs)
        total, f1 = C.c_float(0), C.c_float(0)
        nbytes, launches = C.c_size_t(0), C.c_ulonglong(0)
        rc = lib().sjb_bench_device(self._ctx, n, a, width, height, stride, C.byref(params), iters,
                                    C.byref(total), C.byref(f1), C.byref(nbytes), C.byref(launches))
        _check(self._ctx, rc, "sjb_bench_device")
        return total.value, f1.value, nbytes.value, launches.value

    def bench_output(self, index, cap=64 << 20):
        """bytes of picture `index` as the last round of bench_device left them in HBM"""
        size = C.c_size_t(0)
        out = np.empty(cap, np.uint8)
        rc = lib().sjb_bench_output(self._ctx, index, out.ctypes.data, cap, C.byref(size))
        _check(self._ctx, rc, "sjb_bench_output")
        return out[:size.value].tobytes()

    def bench_f1(self, dev_ptrs, width, height, stride, params, iters):
        n = len(dev_ptrs)
        a = (C.c_void_p * n)(*dev_ptrs)
        ms, fpl = C.c_float(0), C.c_int(0)
        rc = lib().sjb_bench_f1(self._ctx, n, a, width, height, stride, C.byref(params), iters, C.byref(ms),
                                C.byref(fpl))
        _check(self._ctx, rc, "sjb_bench_f1")
        return ms.value, fpl.value


def set_score_table(table):
    """Hands the reference's generated 343 x 343 riskiness table (sjpeg::kSharpnessScore) to the
    library, process-wide; None clears it."""
    if table is None:
        return lib().sjb_set_score_table(None, 0)
    t = np.ascontiguousarray(table, dtype=np.uint8)
    rc = lib().sjb_set_score_table(t.ctypes.data, t.size)
    if rc != OK:
        raise SjpegB200Error("sjb_set_score_table: rc=%d" % rc)
    return rc


def default_score_table():
    """The table the library finds on its own next to the .so (written at build time where the
    reference sources are present), or None."""
    path = os.path.join(os.path.dirname(LIB_PATH), "sjpeg_score_table.bin")
    if not os.path.exists(path):
        return None
    t = np.fromfile(path, np.uint8)
    return t if t.size == 343 * 343 else None


def sjpeg_encode(rgb, width, height, stride, quality, method, yuv_mode, base=None):
    """The drop-in C entry point SjpegEncode() (api.cc:32-49) through the library's own symbol.
    Returns bytes, or None when it returns 0."""
    out = _u8p()
    ptr = base if base is not None else rgb.ctypes.data
    n = lib().SjpegEncode(ptr, width, height, stride, C.byref(out), float(quality), int(method), int(yuv_mode))
    if n == 0:
        return None
    data = C.string_at(out, n)
    lib().SjpegFreeBuffer(out)
    return data
