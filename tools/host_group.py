"""e2e throughput of sjb_encode_batch from pinned host memory for one SJB_HOST_BATCH_GROUP setting
(read once per process): python tools/host_group.py W H N method"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import oracle_lib as O
import sjpeg_b200 as S
w, h, n, method = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
ctx = S.Context(0)
frames = [O.make_rgb("B", w, h, 7654321 + f) for f in range(n)]
p = S.default_params(75, method, S.YUV_420)
pinned = []
for f in frames:
    ptr = S.lib().sjb_host_alloc(f.nbytes); C.memmove(ptr, f.ctypes.data, f.nbytes); pinned.append(ptr)
cap = 4 << 20
outs = [S.lib().sjb_host_alloc(cap) for _ in frames]
for _ in range(3):
    sizes = ctx.encode_batch(pinned, False, w, h, 3 * w, p, outs, False, cap)
best = 1e9
for _ in range(6):
    t0 = time.perf_counter()
    sizes = ctx.encode_batch(pinned, False, w, h, 3 * w, p, outs, False, cap)
    best = min(best, time.perf_counter() - t0)
ok = bytes((C.c_uint8 * sizes[0]).from_address(outs[0])) == O.oracle_encode(frames[0], w, h, 3 * w, 75.0, method, O.YUV_420)
print("group %s  %dx%d x %d m%d: %.3f ms  %.2f Gpix/s  ok=%s" % (os.environ.get("SJB_HOST_BATCH_GROUP", "default"), w, h, n, method,
                                                          best * 1e3, n * w * h / best / 1e9, ok), flush=True)
