"""Small mixed workload for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as O
import sjpeg_b200 as S
ctx = S.Context(0)
ok = True
for (w, h, m, mode) in ((256, 128, 0, S.YUV_420), (203, 117, 4, S.YUV_444), (203, 117, 7, S.YUV_420), (64, 48, 1, S.YUV_400)):
    rgb = O.make_rgb("A", w, h)
    got = ctx.encode(rgb, w, h, 3 * w, S.default_params(75, m, mode))
    ok &= got == O.oracle_encode(rgb, w, h, 3 * w, 75.0, m, mode)
frames = [O.make_rgb("B", 320, 240, 1 + f) for f in range(5)]
outs = [np.empty(1 << 18, np.uint8) for _ in frames]
sizes = ctx.encode_batch([f.ctypes.data for f in frames], False, 320, 240, 960, S.default_params(75, 0, S.YUV_420),
                         [o.ctypes.data for o in outs], False, 1 << 18)
for f, o, s in zip(frames, outs, sizes):
    ok &= o[:s].tobytes() == O.oracle_encode(f, 320, 240, 960, 75.0, 0, S.YUV_420)
print("sanitize workload parity:", ok)
ctx.close()
