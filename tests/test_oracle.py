"""CPU tier: the oracle restatement against (a) the committed golden md5s that the compiled,
unmodified reference produced (tests/golden/ref_md5.json) and (b) the compiled reference itself
where oracle/_ref is available (byte equality over a parameter matrix, incl. the edge cases the
reference's own tests exercise: 1x1, clipped MCUs, negative strides, invalid arguments)."""
import json
import os

import numpy as np
import pytest

import oracle_lib as O

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_md5.json")))
SMALL = [c for c in GOLD["cases"] if c["w"] * c["h"] <= 3840 * 2160]


@pytest.mark.parametrize("case", SMALL, ids=lambda c: "%s_%dx%d_q%d_m%d_y%d" % (
    c["gen"], c["w"], c["h"], c["quality"], c["method"], c["yuv_mode"]))
def test_oracle_matches_golden(case):
    rgb = O.make_rgb(case["gen"], case["w"], case["h"], case["seed"])
    assert O.md5(rgb.tobytes()) == case["input_md5"]
    data = O.oracle_encode(rgb, case["w"], case["h"], 3 * case["w"], float(case["quality"]),
                           case["method"], case["yuv_mode"])
    assert len(data) == case["size"]
    assert O.md5(data) == case["md5"]


def test_oracle_config5_frames():
    c5 = GOLD["config5"]
    for f in (0, 1, 63):
        rgb = O.make_rgb("B", c5["w"], c5["h"], 7654321 + f)
        data = O.oracle_encode(rgb, c5["w"], c5["h"], 3 * c5["w"], 75.0, 0, O.YUV_420)
        assert O.md5(data) == c5["frame_md5"][f]


def _patterns(w, h, rng):
    yield O.make_rgb("A", w, h)
    yield O.make_rgb("B", w, h)
    yield rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
    yield (rng.randint(0, 2, (h, w, 3)) * 255).astype(np.uint8)
    cb = (np.indices((h, w)).sum(0) % 2 * 255).astype(np.uint8)
    yield np.ascontiguousarray(np.stack([cb, 255 - cb, cb], -1))


@pytest.mark.skipif(O.ref() is None, reason="oracle/_ref not built (reference sources absent)")
def test_oracle_equals_compiled_reference_matrix():
    rng = np.random.RandomState(1)
    n = 0
    for (w, h) in ((203, 117), (16, 16), (1, 1), (8, 8), (17, 9), (64, 48), (7, 33)):
        for rgb in _patterns(w, h, rng):
            for q in (0, 1, 25, 50, 75, 90, 93, 97, 100):
                for mode in (O.YUV_420, O.YUV_444, O.YUV_400):
                    for m in range(9):
                        a = O.oracle_encode(rgb, w, h, 3 * w, float(q), m, mode)
                        b = O.ref_encode(rgb, w, h, 3 * w, float(q), m, mode)
                        assert a == b, (w, h, q, mode, m)
                        n += 1
    assert n == 7 * 5 * 9 * 3 * 9


@pytest.mark.skipif(O.ref() is None, reason="oracle/_ref not built")
def test_oracle_negative_and_padded_strides():
    w, h = 203, 117
    rgb = O.make_rgb("A", w, h)
    for m, mode in ((0, O.YUV_420), (4, O.YUV_444), (7, O.YUV_400)):
        base = rgb.ctypes.data + (h - 1) * 3 * w
        a = O.oracle_encode(rgb, w, h, -3 * w, 75.0, m, mode, base=base)
        b = O.ref_encode(rgb, w, h, -3 * w, 75.0, m, mode, base=base)
        c = O.ref_encode(np.ascontiguousarray(rgb[::-1]), w, h, 3 * w, 75.0, m, mode)
        assert a == b == c
        padded = np.full((h, 3 * w + 37), 0xAB, np.uint8)
        padded[:, :3 * w] = rgb.reshape(h, 3 * w)
        d = O.oracle_encode(padded, w, h, 3 * w + 37, 75.0, m, mode)
        assert d == O.ref_encode(rgb, w, h, 3 * w, 75.0, m, mode)


def test_oracle_refuses_invalid_arguments():
    rgb = O.make_rgb("A", 16, 16)
    assert O.oracle_encode(rgb, 0, 16, 48, 75.0, 0, O.YUV_420) is None
    assert O.oracle_encode(rgb, 16, -1, 48, 75.0, 0, O.YUV_420) is None
    assert O.oracle_encode(rgb, 16, 16, 47, 75.0, 0, O.YUV_420) is None
    assert O.oracle_encode(rgb, 16, 16, 48, 75.0, 0, 7) is None          # unknown mode
    assert O.oracle_encode(rgb, 16, 16, 48, 75.0, -3, O.YUV_420) == O.oracle_encode(rgb, 16, 16, 48, 75.0, 0, O.YUV_420)
    assert O.oracle_encode(rgb, 16, 16, 48, 75.0, 11, O.YUV_420) == O.oracle_encode(rgb, 16, 16, 48, 75.0, 8, O.YUV_420)


@pytest.mark.skipif(O.ref() is None, reason="oracle/_ref not built")
def test_oracle_planar_inputs_equal_compiled_reference():
    """EncodeYUV420 / YUV444 / NV12 / NV21 / Gray (encoders.cc:256-507), padded strides, clipped MCUs"""
    flags = {0: (0, 0, 0), 1: (1, 0, 0), 3: (0, 1, 0), 4: (1, 1, 0), 7: (1, 1, 1)}
    for (w, h) in ((17, 13), (64, 48), (203, 117), (1, 1), (16, 16), (33, 40)):
        for q in (30, 80, 97):
            for method, fl in flags.items():
                for kind in range(5):
                    planes = O.make_planes(kind, w, h, seed=w + q + kind)
                    a = O.oracle_encode_planar(kind, planes, w, h, q, method)
                    b = O.ref_encode_planar(kind, planes, w, h, q, *fl)
                    assert a is not None and a == b, (w, h, q, method, kind)


@pytest.mark.skipif(O.ref() is None, reason="oracle/_ref not built")
def test_sharp_yuv_and_riskiness_match_reference():
    """Pins oracle/sjpeg_oracle_sharp.c: planes of sjpeg::ApplySharpYUVConversion, (mode, risk) of
    SjpegRiskiness, and whole SJPEG_YUV_SHARP files of the compiled unmodified reference."""
    rng = np.random.RandomState(1)
    table = O.score_table()
    for (w, h) in ((1, 1), (3, 7), (4, 4), (5, 5), (5, 4), (4, 9), (6, 5), (7, 7), (16, 16), (17, 33), (64, 48),
                   (203, 117), (256, 255)):
        sat = np.zeros((h, w, 3), np.uint8)
        sat[:, ::2, 0] = 255
        sat[::2, :, 2] = 255
        gray = np.repeat(rng.randint(0, 256, (h, w, 1)), 3, axis=2).astype(np.uint8)
        for img in (O.make_rgb("A", w, h), O.make_rgb("B", w, h), rng.randint(0, 256, (h, w, 3)).astype(np.uint8),
                    (rng.randint(0, 2, (h, w, 3)) * 255).astype(np.uint8), sat, gray):
            for a, b in zip(O.oracle_sharp_yuv(img, w, h, 3 * w), O.ref_sharp_yuv(img, w, h, 3 * w)):
                assert np.array_equal(a, b), (w, h)
            assert O.oracle_riskiness(img, w, h, 3 * w, table) == O.ref_riskiness(img, w, h, 3 * w), (w, h)
            for m in (0, 4):
                assert O.oracle_encode(img, w, h, 3 * w, 75.0, m, O.YUV_SHARP) == \
                    O.ref_encode(img, w, h, 3 * w, 75.0, m, O.YUV_SHARP), (w, h, m)
