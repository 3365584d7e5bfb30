"""Encodes one synthetic picture a few times through sjb_encode (pinned host input) -- a small,
fixed command line to put under ncu.  Usage: python tools/one_encode.py gen w h quality method yuv [reps]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O
import sjpeg_b200 as S
gen, w, h, q, m, mode = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 3
ctx = S.Context(0)
rgb = O.make_rgb(gen, w, h)
pin = S.lib().sjb_host_alloc(rgb.nbytes)
C.memmove(pin, rgb.ctypes.data, rgb.nbytes)
cap = 64 << 20
out = S.lib().sjb_host_alloc(cap)
p = S.default_params(q, m, mode)
for _ in range(reps):
    n = ctx.encode_into(pin, False, w, h, 3 * w, p, out, False, cap)
    print(n, ctx.last_timings())
