"""Single-call latency of the drop-in entry point SjpegEncode() (pageable host buffer in, new[]
buffer out) for a few sizes, next to the compiled reference on one host thread."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O
import sjpeg_b200 as S
for (w, h) in ((512, 512), (1920, 1080), (3840, 2160), (7680, 4320)):
    # a rotation of distinct pictures, together larger than the host's last-level cache: every call reads its
    # pixels from DRAM, as a stream of frames does (re-encoding ONE buffer keeps it cache-warm: 4K 0.73 instead of 0.85 ms)
    nsrc = max(3, min(16, (300 << 20) // (3 * w * h)))
    frames = [O.make_rgb("B", w, h, 7654321 + i) for i in range(nsrc)]
    rgb = frames[0]
    for method in (0, 4):
        S.sjpeg_encode(rgb, w, h, 3 * w, 75, method, S.YUV_420)
        t = []
        for i in range(2 * nsrc):
            src = frames[i % nsrc]
            t0 = time.perf_counter(); a = S.sjpeg_encode(src, w, h, 3 * w, 75, method, S.YUV_420); t.append(time.perf_counter() - t0)
        a = S.sjpeg_encode(rgb, w, h, 3 * w, 75, method, S.YUV_420)
        r = []
        if O.ref() is not None:
            for _ in range(3):
                t0 = time.perf_counter(); b = O.ref_encode(rgb, w, h, 3 * w, 75.0, method, O.YUV_420); r.append(time.perf_counter() - t0)
            assert a == b
        print("%dx%d m%d: B200 %.3f ms (%.0f Mpix/s)   reference 1 thread %.3f ms (%.0f Mpix/s)" % (
            w, h, method, 1e3 * min(t), w * h / min(t) / 1e6, 1e3 * min(r) if r else -1, w * h / min(r) / 1e6 if r else -1))
