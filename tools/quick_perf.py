"""Quick per-stage timing of the device pipeline on one GPU (for A/B runs while tuning kernels):
16 x 4K pictures, gen B and gen A, methods 0 and 4 (+ 4:4:4 m1), outputs checked against the oracle.
  python tools/quick_perf.py [cases: e.g. B0,A0,B4,A4,B1x]   (x = 4:4:4 q90)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import oracle_lib as O  # noqa: E402
import sjpeg_b200 as S  # noqa: E402

cases = (sys.argv[1] if len(sys.argv) > 1 else "B0,A0,B4,A4,B1x,A1x").split(",")
w, h = 3840, 2160
ctx = S.Context(0)
frames_cache = {}
for case in cases:
    gen, method, x = case[0], int(case[1]), case.endswith("x")
    mode, q, n = (S.YUV_444, 90, 8) if x else (S.YUV_420, 75, 16)
    if (gen, n) not in frames_cache:
        frames_cache[(gen, n)] = [O.make_rgb(gen, w, h, 7654321 + f) for f in range(n)]
    frames = frames_cache[(gen, n)]
    dev = [torch.from_numpy(f.reshape(-1)).cuda() for f in frames]
    ptrs = [t.data_ptr() for t in dev]
    p = S.default_params(q, method, mode)
    ctx.bench_device(ptrs, w, h, 3 * w, p, 8)
    iters = 6
    total_ms = min(ctx.bench_device(ptrs, w, h, 3 * w, p, iters)[0] for _ in range(3))
    ok = all(ctx.bench_output(i) == O.oracle_encode(frames[i], w, h, 3 * w, float(q), method, mode) for i in (0, n - 1))
    _, fpl = ctx.last_stage_timings()
    best = None
    for _ in range(4):
        ctx.bench_device(ptrs[:fpl], w, h, 3 * w, p, 1)
        st, _ = ctx.last_stage_timings()
        if best is None or sum(st.values()) < sum(best.values()):
            best = st
    print("%-4s %s m%d q%d: %7.1f Gpix/s  %6.2f us/picture  exact=%s  per launch of %d (us): %s" % (
        case, "444" if x else "420", method, q, n * iters * w * h / total_ms / 1e6, total_ms * 1e3 / (n * iters), ok, fpl,
        "  ".join("%s %.1f" % (k, v * 1e3) for k, v in best.items())), flush=True)
    del dev
    torch.cuda.empty_cache()
