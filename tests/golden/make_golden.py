"""Generates tests/golden/ref_md5.json by running the UNMODIFIED reference compiled from
/root/reference/src (oracle/_ref/libsjpeg_ref.so, recipe oracle/Makefile) on the deterministic
inputs of SURVEY.md 8(d).  Run in the build container only:  python tests/golden/make_golden.py
The values for the BASELINE.md rows are cross-checked against that table."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import oracle_lib as O   # noqa: E402

CASES = [
    # gen, w, h, quality, method, yuv_mode
    ("A", 512, 512, 75, 0, 1), ("B", 512, 512, 75, 0, 1),
    ("A", 3840, 2160, 75, 0, 1), ("B", 3840, 2160, 75, 0, 1),
    ("B", 3840, 2160, 75, 4, 1), ("B", 3840, 2160, 75, 7, 1),
    ("A", 3840, 2160, 90, 1, 3), ("B", 3840, 2160, 90, 1, 3),
    ("B", 7680, 4320, 75, 0, 1), ("B", 7680, 4320, 75, 6, 1), ("B", 7680, 4320, 75, 7, 1),
    ("A", 7680, 4320, 75, 6, 1), ("A", 7680, 4320, 75, 8, 1),
    ("B", 1920, 1080, 75, 0, 1),
    ("A", 203, 117, 75, 0, 1), ("A", 203, 117, 90, 4, 3), ("A", 203, 117, 50, 7, 4),
    ("B", 1000, 700, 85, 2, 1), ("A", 1000, 700, 30, 5, 3), ("B", 1000, 700, 95, 8, 4),
    # SJPEG_YUV_SHARP (yuv_mode 2): iterative sharp RGB->YUV420 pre-pass, then the planar encoder
    ("A", 203, 117, 75, 0, 2), ("A", 1000, 700, 85, 1, 2), ("B", 1001, 701, 60, 4, 2),
    ("A", 3840, 2160, 75, 0, 2), ("B", 3840, 2160, 75, 4, 2),
]
# SjpegRiskiness (mode, risk) of the reference on the same generators
RISK_CASES = [("A", 512, 512), ("B", 512, 512), ("A", 3840, 2160), ("B", 3840, 2160), ("A", 203, 117), ("B", 1001, 701)]


def main():
    assert O.ref() is not None, "needs oracle/_ref (build container only)"
    out = []
    for gen, w, h, q, m, mode in CASES:
        rgb = O.make_rgb(gen, w, h)
        data = O.ref_encode(rgb, w, h, 3 * w, float(q), m, mode)
        out.append({"gen": gen, "w": w, "h": h, "quality": q, "method": m, "yuv_mode": mode,
                    "seed": 7654321, "size": len(data), "md5": O.md5(data),
                    "input_md5": O.md5(rgb.tobytes())})
        print(out[-1])
    # config 5: 64 frames, seeds 7654321+f; digest of digests
    digests = []
    total = 0
    for f in range(64):
        rgb = O.make_rgb("B", 1920, 1080, 7654321 + f)
        data = O.ref_encode(rgb, 1920, 1080, 3 * 1920, 75.0, 0, 1)
        digests.append(O.md5(data))
        total += len(data)
    c5 = {"frames": 64, "w": 1920, "h": 1080, "total_size": total,
          "md5_of_md5s": O.md5("".join(digests).encode()), "frame_md5": digests}
    risk = []
    for gen, w, h in RISK_CASES:
        rgb = O.make_rgb(gen, w, h)
        mode, value = O.ref_riskiness(rgb, w, h, 3 * w)
        risk.append({"gen": gen, "w": w, "h": h, "seed": 7654321, "mode": mode, "risk": value})
        print(risk[-1])
    with open(os.path.join(os.path.dirname(__file__), "ref_md5.json"), "w") as fp:
        json.dump({"generator": "tests/golden/make_golden.py", "reference_commit": "6b8cd89",
                   "cases": out, "config5": c5, "riskiness": risk}, fp, indent=1)


if __name__ == "__main__":
    main()
