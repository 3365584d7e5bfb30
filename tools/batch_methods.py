"""Throughput of a batch of 4K pictures by method: device-resident (sjb_bench_device) and end to end
(sjb_encode_batch, pinned host input, host output).  Usage: python tools/batch_methods.py [gen] [n]"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import oracle_lib as O
import sjpeg_b200 as S
gen = sys.argv[1] if len(sys.argv) > 1 else "B"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16
W, H = 3840, 2160
ctx = S.Context(0)
frames = [O.make_rgb(gen, W, H, 7654321 + f) for f in range(n)]
dev = [torch.from_numpy(f.reshape(-1)).cuda() for f in frames]
ptrs = [t.data_ptr() for t in dev]
pinned = []
for f in frames:
    p = S.lib().sjb_host_alloc(f.nbytes); C.memmove(p, f.ctypes.data, f.nbytes); pinned.append(p)
cap = 8 << 20
outs = [S.lib().sjb_host_alloc(cap) for _ in range(n)]
for method in (0, 1, 4, 7):
    p = S.default_params(75, method, S.YUV_420)
    for _ in range(2):
        ctx.bench_device(ptrs, W, H, 3 * W, p, 3)
    total = min(ctx.bench_device(ptrs, W, H, 3 * W, p, 10)[0] for _ in range(3))
    stages, fr = ctx.last_stage_timings()
    ok = all(ctx.bench_output(i) == O.oracle_encode(frames[i], W, H, 3 * W, 75.0, method, O.YUV_420) for i in (0, n - 1))
    ctx.encode_batch(pinned, False, W, H, 3 * W, p, outs, False, cap)
    t0 = time.perf_counter()
    for _ in range(5):
        sizes = ctx.encode_batch(pinned, False, W, H, 3 * W, p, outs, False, cap)
    dt = (time.perf_counter() - t0) / 5
    print("gen%s m%d: device %.1f Gpix/s (%.1f us/picture)  e2e %.2f Gpix/s  ok=%s  stages(ms per %d pictures)=%s" % (
        gen, method, n * 10 * W * H / total / 1e6, total / (10 * n) * 1e3, n * W * H / dt / 1e9, ok, fr,
        {k: round(v, 4) for k, v in stages.items()}), flush=True)
