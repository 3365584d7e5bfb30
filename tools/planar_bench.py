"""Device time of the F1 stage for planar / semi-planar 4K input next to packed RGB (sjb_encode_planar with
pinned host planes; stage timers of the library)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as O
import sjpeg_b200 as S
ctx = S.Context(0)
w, h = 3840, 2160
rgb = O.make_rgb("B", w, h)
for method in (0, 4):
    p = S.default_params(75, method, S.YUV_420)
    for _ in range(3):
        ctx.encode(rgb, w, h, 3 * w, p)
    st, _ = ctx.last_stage_timings()
    print("packed RGB 4:2:0 m%d: stages ms %s" % (method, {k: round(v, 4) for k, v in st.items()}))
    for kind, name in ((O.KIND_YUV420, "YUV420 planar"), (O.KIND_NV12, "NV12"), (O.KIND_YUV444, "YUV444 planar"), (O.KIND_GRAY, "gray")):
        planes = O.make_planes(kind, w, h, seed=3, pad=(0, 0, 0))
        pp = S.default_params(75, method, O.KIND_MODE[kind])
        for _ in range(3):
            got = ctx.encode_planar(*O.planar_args(kind, planes), w, h, pp)
        st, _ = ctx.last_stage_timings()
        ok = got == O.oracle_encode_planar(kind, planes, w, h, 75, method)
        print("%-14s m%d: stages ms %s  bit-exact=%s" % (name, method, {k: round(v, 4) for k, v in st.items()}, ok))
rgbA = O.make_rgb("A", w, h)
import time
for _ in range(2):
    t0 = time.perf_counter(); got = ctx.encode(rgbA, w, h, 3 * w, S.default_params(75, 4, S.YUV_SHARP)); dt = time.perf_counter() - t0
print("SJPEG_YUV_SHARP 4K m4 host-to-host %.2f ms, stages %s" % (dt * 1e3, {k: round(v, 4) for k, v in ctx.last_stage_timings()[0].items()}))
