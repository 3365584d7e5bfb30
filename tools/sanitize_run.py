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
# batches of several pictures per launch at the adaptive / trellis methods (A1 analysis kernels with
# gridDim.z > 1, per-picture tables), device-resident groups, a planar batch on the bulk-copy F1
import torch
framesA = [O.make_rgb("A", 320, 240, 40 + f) for f in range(5)]
for method, mode in ((4, S.YUV_420), (7, S.YUV_420), (3, S.YUV_444), (6, S.YUV_400)):
    dev = [torch.from_numpy(f.reshape(-1)).cuda() for f in framesA]
    douts = [torch.zeros(1 << 18, dtype=torch.uint8, device="cuda") for _ in framesA]
    sizes = ctx.encode_batch([t.data_ptr() for t in dev], True, 320, 240, 960, S.default_params(75, method, mode),
                             [t.data_ptr() for t in douts], True, 1 << 18)
    for f, o, s in zip(framesA, douts, sizes):
        ok &= bytes(o[:s].cpu().numpy()) == O.oracle_encode(f, 320, 240, 960, 75.0, method, mode)
# whole-picture passes: sharp conversion (cluster of 2 CTAs at 1100 px), AUTO with the score table
for (w, h) in ((66, 34), (1100, 20)):
    rgb = O.make_rgb("A", w, h)
    ok &= ctx.encode(rgb, w, h, 3 * w, S.default_params(75, 0, S.YUV_SHARP)) == \
        O.oracle_encode(rgb, w, h, 3 * w, 75.0, 0, O.YUV_SHARP)
table = O.score_table()
if table is not None:
    S.set_score_table(table)
    rgb = O.make_rgb("A", 203, 117)
    ok &= ctx.riskiness(rgb, 203, 117, 609) == O.oracle_riskiness(rgb, 203, 117, 609, table)
# a picture with several entropy tiles per persistent CTA and blocks longer than the 512-bit slots
rng = np.random.RandomState(2)
noise = rng.randint(0, 256, (160, 1024, 3)).astype(np.uint8)
ok &= ctx.encode(noise, 1024, 160, 3072, S.default_params(98, 0, S.YUV_444)) == \
    O.oracle_encode(noise, 1024, 160, 3072, 98.0, 0, O.YUV_444)
print("sanitize workload parity:", ok)
ctx.close()
