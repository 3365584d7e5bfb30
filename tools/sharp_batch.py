"""Batched SJPEG_YUV_SHARP / SJPEG_YUV_AUTO (the batched SjpegCompress): 16 x 4K device-resident and
from pinned host memory, checked against the oracle.  python tools/sharp_batch.py [n]"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import oracle_lib as O
import sjpeg_b200 as S
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
w, h = 3840, 2160
ctx = S.Context(0)
frames = [O.make_rgb("AB"[f & 1], w, h, 7654321 + f) for f in range(n)]
dev = [torch.from_numpy(f.reshape(-1)).cuda() for f in frames]
cap = 8 << 20
douts = [torch.zeros(cap, dtype=torch.uint8, device="cuda") for _ in range(n)]
for (name, mode, method) in (("YUV_SHARP m0", S.YUV_SHARP, 0), ("YUV_AUTO m4 (SjpegCompress)", S.YUV_AUTO, 4), ("YUV_420 m4", S.YUV_420, 4)):
    p = S.default_params(75, method, mode)
    best = 1e9
    for _ in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        sizes = ctx.encode_batch([t.data_ptr() for t in dev], True, w, h, 3 * w, p, [t.data_ptr() for t in douts], True, cap)
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    got = [bytes(douts[i][:sizes[i]].cpu().numpy()) for i in (0, 1)]
    if mode == S.YUV_AUTO:
        tab = O.score_table()
        want = []
        for i in (0, 1):
            m, _ = O.oracle_riskiness(frames[i], w, h, 3 * w, tab)
            want.append(O.oracle_encode(frames[i], w, h, 3 * w, 75.0, method, m))
    else:
        want = [O.oracle_encode(frames[i], w, h, 3 * w, 75.0, method, mode) for i in (0, 1)]
    print("%-30s %d x 4K device-resident: %.2f ms  %.2f Gpix/s  exact=%s" % (name, n, best * 1e3, n * w * h / best / 1e9, got == want), flush=True)
