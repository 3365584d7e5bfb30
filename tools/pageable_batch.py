import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as O
import sjpeg_b200 as S
ctx = S.Context(0)
for (w, h, n) in ((1920, 1080, 64), (3840, 2160, 16)):
    frames = [O.make_rgb("B", w, h, 7654321 + f) for f in range(n)]
    cap = 4 << 20
    outs = [np.empty(cap, np.uint8) for _ in frames]
    for method in (0, 4):
        p = S.default_params(75, method, S.YUV_420)
        best = 1e9
        for _ in range(6):
            t0 = time.perf_counter()
            sizes = ctx.encode_batch([f.ctypes.data for f in frames], False, w, h, 3 * w, p, [o.ctypes.data for o in outs], False, cap)
            best = min(best, time.perf_counter() - t0)
        ok = outs[0][:sizes[0]].tobytes() == O.oracle_encode(frames[0], w, h, 3 * w, 75.0, method, O.YUV_420)
        print("pageable numpy buffers in and out, %d x %dx%d m%d: %.2f ms  %.2f Gpix/s ok=%s" % (n, w, h, method, best * 1e3, n * w * h / best / 1e9, ok), flush=True)
