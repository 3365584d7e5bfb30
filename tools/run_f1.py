"""Small driver for profiling: runs the fused F1 kernel (and optionally the whole device pipeline)
on a few 4K frames.  Usage: python tools/run_f1.py [frames] [iters] [mode: f1|full] [w h yuv method]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import oracle_lib as O  # noqa: E402
import sjpeg_b200 as S  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
what = sys.argv[3] if len(sys.argv) > 3 else "f1"
w = int(sys.argv[4]) if len(sys.argv) > 4 else 3840
h = int(sys.argv[5]) if len(sys.argv) > 5 else 2160
yuv = int(sys.argv[6]) if len(sys.argv) > 6 else S.YUV_420
method = int(sys.argv[7]) if len(sys.argv) > 7 else 0
gen = sys.argv[8] if len(sys.argv) > 8 else "B"
ctx = S.Context(0)
p = S.default_params(75, method, yuv)
frames = [O.make_rgb(gen, w, h, 7654321 + f) for f in range(n)]
dev = [torch.from_numpy(f.reshape(-1)).cuda() for f in frames]
ptrs = [t.data_ptr() for t in dev]
import subprocess
def clocks():
    return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw", "--format=csv,noheader"],
                          capture_output=True, text=True).stdout.strip()
if what == "f1":
    for _ in range(3):
        ctx.bench_f1(ptrs, w, h, 3 * w, p, 40)      # warm-up: clocks ramp, I-cache, TLB
    ms, fpl = min(ctx.bench_f1(ptrs, w, h, 3 * w, p, iters) for _ in range(5))
    print("clocks:", clocks(), "frames/launch", fpl)
    ms /= fpl
    print("f1 ms/frame %.5f  -> %.1f GB/s algorithmic" % (ms, (3 * w * h * 2) / ms / 1e6))
else:
    for _ in range(3):
        ctx.bench_device(ptrs, w, h, 3 * w, p, 10)
    total, f1, nbytes, launches = min(ctx.bench_device(ptrs, w, h, 3 * w, p, iters) for _ in range(5))
    print("clocks:", clocks())
    print("device pipeline: %.4f ms/frame, %.1f Mpix/s, jpeg %d B, launches %d" % (
        total / (n * iters), n * iters * w * h / total / 1e3, nbytes, launches))
    t = ctx.last_timings()
    print("last group on lane 0 (ms): F1 stage %.4f, entropy+stuffing %.4f, total %.4f" % (t[0], t[1], t[2]))
