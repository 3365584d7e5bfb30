"""Device-side and host-to-host times of BASELINE.json configs 2-4 (one picture at a time through
sjb_encode with a pinned host buffer), next to the compiled reference on one host thread."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as O
import sjpeg_b200 as S
ctx = S.Context(0)
CASES = [("C2", "B", 3840, 2160, 75, 0, S.YUV_420), ("C2", "A", 3840, 2160, 75, 0, S.YUV_420),
         ("C3", "A", 3840, 2160, 90, 1, S.YUV_444), ("C3", "B", 3840, 2160, 90, 1, S.YUV_444),
         ("C4", "B", 7680, 4320, 75, 6, S.YUV_420), ("C4", "B", 7680, 4320, 75, 7, S.YUV_420),
         ("C4", "A", 7680, 4320, 75, 6, S.YUV_420), ("C4", "A", 7680, 4320, 75, 7, S.YUV_420)]
for (cfg, gen, w, h, q, m, mode) in CASES:
    rgb = O.make_rgb(gen, w, h)
    pin = S.lib().sjb_host_alloc(rgb.nbytes)
    C.memmove(pin, rgb.ctypes.data, rgb.nbytes)
    cap = 64 << 20
    out = S.lib().sjb_host_alloc(cap)
    p = S.default_params(q, m, mode)
    best = 1e9
    for _ in range(4):
        t0 = time.perf_counter()
        n = ctx.encode_into(pin, False, w, h, 3 * w, p, out, False, cap)
        best = min(best, time.perf_counter() - t0)
    tim = ctx.last_timings()
    data = bytes((C.c_uint8 * n).from_address(out))
    ref_t = -1.0
    ok = None
    if O.ref() is not None:
        t0 = time.perf_counter(); r = O.ref_encode(rgb, w, h, 3 * w, float(q), m, mode); ref_t = time.perf_counter() - t0
        ok = (r == data)
    print("%s gen%s %dx%d q%d m%d yuv%d: %d B md5 %s equal_ref=%s | host-to-host %.2f ms (%.0f Mpix/s) | device F1-stage %.3f ms, rest %.3f ms, total %.3f ms | reference 1 thread %.1f ms" % (
        cfg, gen, w, h, q, m, mode, n, O.md5(data), ok, best * 1e3, w * h / best / 1e6, tim[0], tim[1], tim[2], ref_t * 1e3))
    S.lib().sjb_host_free(pin); S.lib().sjb_host_free(out)
