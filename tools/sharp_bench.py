"""Times the whole-picture passes (sharp RGB->YUV 4:2:0, riskiness) and the SJPEG_YUV_SHARP /
SJPEG_YUV_AUTO encodes on a B200 against the compiled reference on one host thread.
  python tools/sharp_bench.py [W H]"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np   # noqa: E402
import torch         # noqa: E402
import oracle_lib as O   # noqa: E402
import sjpeg_b200 as S   # noqa: E402


def timed(fn, n=5):
    fn()
    best = 1e9
    for _ in range(n):
        t = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t)
    return best * 1e3


def main():
    w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (3840, 2160)
    ctx = S.Context(0)
    table = O.score_table()
    if table is not None:
        S.set_score_table(table)
    for gen in "AB":
        rgb = O.make_rgb(gen, w, h)
        d_rgb = torch.from_numpy(rgb.reshape(-1)).cuda()
        cw, ch = (w + 1) // 2, (h + 1) // 2
        dy = torch.empty(w * h, dtype=torch.uint8, device="cuda")
        du = torch.empty(cw * ch, dtype=torch.uint8, device="cuda")
        dv = torch.empty(cw * ch, dtype=torch.uint8, device="cuda")

        def sharp_dev():
            rc = S.lib().sjb_sharp_yuv(ctx._ctx, d_rgb.data_ptr(), 1, w, h, 3 * w, dy.data_ptr(), du.data_ptr(),
                                       dv.data_ptr(), 1)
            assert rc == 0
        ms_dev = timed(sharp_dev)
        want = O.ref_sharp_yuv(rgb, w, h, 3 * w) if O.ref() else O.oracle_sharp_yuv(rgb, w, h, 3 * w)
        ok = np.array_equal(dy.cpu().numpy().reshape(h, w), want[0]) and \
            np.array_equal(du.cpu().numpy().reshape(ch, cw), want[1]) and np.array_equal(dv.cpu().numpy().reshape(ch, cw), want[2])
        ms_ref = timed(lambda: O.ref_sharp_yuv(rgb, w, h, 3 * w), 2) if O.ref() else float("nan")
        print("gen %s %dx%d sharp-yuv: device-resident %.3f ms, reference CPU (1 thread) %.1f ms, planes equal: %s"
              % (gen, w, h, ms_dev, ms_ref, ok))
        p = S.default_params(75, 0, S.YUV_SHARP)
        ms_enc = timed(lambda: ctx.encode(rgb, w, h, 3 * w, p))
        ms_enc_ref = timed(lambda: O.ref_encode(rgb, w, h, 3 * w, 75.0, 0, O.YUV_SHARP), 2) if O.ref() else float("nan")
        print("          SjpegEncode(YUV_SHARP, m0) host-to-host: %.3f ms vs reference %.1f ms" % (ms_enc, ms_enc_ref))
        if table is not None:
            mode, risk = C.c_int(), C.c_float()

            def risk_dev():
                assert S.lib().sjb_riskiness(ctx._ctx, d_rgb.data_ptr(), 1, w, h, 3 * w, C.byref(mode), C.byref(risk)) == 0
            ms_r = timed(risk_dev)
            ms_r_ref = timed(lambda: O.ref_riskiness(rgb, w, h, 3 * w), 2)
            print("          riskiness: device-resident %.3f ms -> (mode %d, risk %.3f), reference CPU %.1f ms -> %s"
                  % (ms_r, mode.value, risk.value, ms_r_ref, O.ref_riskiness(rgb, w, h, 3 * w)))
            pa = S.default_params(75, 4, S.YUV_AUTO)
            ms_auto = timed(lambda: ctx.encode(rgb, w, h, 3 * w, pa))
            ms_auto_ref = timed(lambda: O.ref_encode(rgb, w, h, 3 * w, 75.0, 4, O.YUV_AUTO), 2)
            print("          SjpegCompress-equivalent (AUTO, m4) host-to-host: %.3f ms vs reference %.1f ms"
                  % (ms_auto, ms_auto_ref))
    ctx.close()


if __name__ == "__main__":
    main()
