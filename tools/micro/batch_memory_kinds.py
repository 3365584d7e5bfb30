import ctypes as C, os, sys, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import oracle_lib as O
import sjpeg_b200 as S
ctx = S.Context(0)
w, h, n = 3840, 2160, 16
frames = [O.make_rgb("B", w, h, 7654321 + f) for f in range(n)]
cap = 4 << 20
p = S.default_params(75, 0, S.YUV_420)
def run(inp, outs, label):
    best = 1e9
    for _ in range(6):
        t0 = time.perf_counter()
        sizes = ctx.encode_batch(inp, False, w, h, 3 * w, p, outs, False, cap)
        best = min(best, time.perf_counter() - t0)
    print("%-40s %.2f ms  %.2f Gpix/s" % (label, best * 1e3, n * w * h / best / 1e9), flush=True)
pag_in = [f.ctypes.data for f in frames]
pin_in = []
for f in frames:
    ptr = S.lib().sjb_host_alloc(f.nbytes); C.memmove(ptr, f.ctypes.data, f.nbytes); pin_in.append(ptr)
pag_out_arr = [np.empty(cap, np.uint8) for _ in frames]
pag_out = [o.ctypes.data for o in pag_out_arr]
pin_out = [S.lib().sjb_host_alloc(cap) for _ in frames]
run(pag_in, pag_out, "pageable in, pageable out")
run(pag_in, pin_out, "pageable in, pinned out")
run(pin_in, pag_out, "pinned in, pageable out")
run(pin_in, pin_out, "pinned in, pinned out")
