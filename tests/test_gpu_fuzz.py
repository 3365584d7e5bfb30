"""GPU tier, seeded random sweep: picture size, content, yuv mode (incl. SHARP and, with the
reference's score table, AUTO), method, quality, pixel format, stride padding / sign / base
alignment -- every combination must give the oracle's bytes.  Sizes are drawn so that block counts
land on, just below and just above multiples of the 256-block tiles of the entropy kernel."""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu

import os

# one-off longer runs: SJB_FUZZ_CASES=3000 SJB_FUZZ_SEED=7 python -m pytest tests/test_gpu_fuzz.py -m gpu
N_CASES = int(os.environ.get("SJB_FUZZ_CASES", "240"))
SEED = int(os.environ.get("SJB_FUZZ_SEED", "20261017"))


def _draw_size(rng):
    kind = rng.randint(0, 4)
    if kind == 0:      # tiny
        return int(rng.randint(1, 20)), int(rng.randint(1, 20))
    if kind == 1:      # around entropy-tile boundaries: 256 blocks = e.g. 42.67 MCUs (4:2:0) / 85.3 (4:4:4)
        mcus = int(rng.choice([42, 43, 85, 86, 128, 171, 256, 257]))
        return 16 * mcus + int(rng.randint(-15, 16)), int(rng.randint(1, 40))
    if kind == 2:      # tall and narrow
        return int(rng.randint(1, 40)), int(rng.randint(200, 1500))
    return int(rng.randint(20, 700)), int(rng.randint(20, 500))


def _draw_image(rng, w, h):
    kind = rng.randint(0, 5)
    if kind == 0:
        return O.make_rgb("A", w, h, int(rng.randint(1, 1 << 30)))
    if kind == 1:
        return O.make_rgb("B", w, h, int(rng.randint(1, 1 << 30)))
    if kind == 2:
        return rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
    if kind == 3:
        return (rng.randint(0, 2, (h, w, 3)) * 255).astype(np.uint8)
    return np.full((h, w, 3), int(rng.randint(0, 256)), np.uint8)       # flat: all-zero AC everywhere


def test_random_configurations_bit_exact(gpu_ctx):
    import sjpeg_b200 as S
    rng = np.random.RandomState(SEED)
    table = O.score_table()
    if table is not None:
        S.set_score_table(table)
    try:
        for case in range(N_CASES):
            w, h = _draw_size(rng)
            w, h = max(1, w), max(1, h)
            rgb = _draw_image(rng, w, h)
            modes = [S.YUV_420, S.YUV_444, S.YUV_400, S.YUV_SHARP] + ([S.YUV_AUTO] if table is not None else [])
            mode = int(modes[rng.randint(0, len(modes))])
            method = int(rng.randint(0, 9))
            quality = float(rng.choice([0, 5, 30, 50, 75, 90, 95, 100]))
            fmt = int(rng.randint(0, 3)) if mode in (S.YUV_420, S.YUV_444, S.YUV_400) else S.PIX_RGB
            bpp = 3 if fmt == S.PIX_RGB else 4
            if fmt == S.PIX_RGB:
                pix = rgb
            else:
                order = [0, 1, 2] if fmt == S.PIX_RGBA else [2, 1, 0]
                pix = np.dstack([rgb[:, :, order], rng.randint(0, 256, (h, w, 1)).astype(np.uint8)])
            pad = int(rng.choice([0, 0, 1, 3, 16, 61]))
            lead = int(rng.choice([0, 0, 1, 5]))
            flip = bool(rng.randint(0, 2))
            buf = np.full(lead + h * (bpp * w + pad), 0xA5, np.uint8)
            rows = buf[lead:].reshape(h, bpp * w + pad)
            rows[:, :bpp * w] = (pix[::-1] if flip else pix).reshape(h, bpp * w)
            stride = bpp * w + pad
            base = buf.ctypes.data + lead
            if flip:               # bottom-up storage, negative stride: row 0 is the last stored row
                base += (h - 1) * stride
                stride = -stride
            want_mode = mode
            if mode == S.YUV_AUTO:
                want_mode = O.oracle_riskiness(rgb, w, h, 3 * w, table)[0]
            want = O.oracle_encode(rgb, w, h, 3 * w, quality, method, want_mode)
            p = S.default_params(quality, method, mode)
            p.pix_fmt = fmt
            got = gpu_ctx.encode(buf, w, h, stride, p, base=base)
            assert got == want, dict(case=case, w=w, h=h, mode=mode, method=method, quality=quality, fmt=fmt,
                                     pad=pad, lead=lead, flip=flip)
    finally:
        S.set_score_table(S.default_score_table())


def test_pageable_upload_through_the_stager(gpu_ctx):
    """Pictures of 4 MB and more in ordinary (pageable) numpy memory go through the threaded pinned
    ring of csrc/host_stager.cc; sizes that are not multiples of the 128 KB pieces / 2 MB chunks,
    back-to-back calls that reuse ring slots while their DMA is still in flight, and a bottom-up
    picture.  Bytes must equal the oracle's every time."""
    import sjpeg_b200 as S
    rng = np.random.RandomState(77)
    p = S.default_params(75, 0, S.YUV_420)
    cases = [(1366, 1025), (1920, 1080), (2731, 1537), (3840, 2160), (4099, 2309)]
    for rep in range(3):
        for (w, h) in cases:
            rgb = rng.randint(0, 256, (h, w, 3)).astype(np.uint8) if rep == 0 else O.make_rgb("B", w, h, 100 + rep)
            assert rgb.nbytes >= 4 << 20
            want = O.oracle_encode(rgb, w, h, 3 * w, 75.0, 0, O.YUV_420)
            assert gpu_ctx.encode(rgb, w, h, 3 * w, p) == want, (rep, w, h)
            if rep == 2:
                flipped = np.ascontiguousarray(rgb[::-1])
                base = flipped.ctypes.data + (h - 1) * 3 * w
                assert gpu_ctx.encode(flipped, w, h, -3 * w, p, base=base) == want, ("flip", w, h)
    # batch API: 24 pageable 1080p frames back to back over the lanes
    frames = [O.make_rgb("B", 1920, 1080, 500 + f) for f in range(24)]
    outs = [np.empty(1 << 20, np.uint8) for _ in frames]
    sizes = gpu_ctx.encode_batch([f.ctypes.data for f in frames], False, 1920, 1080, 5760, p,
                                 [o.ctypes.data for o in outs], False, 1 << 20)
    for f, o, n in zip(frames, outs, sizes):
        assert o[:n].tobytes() == O.oracle_encode(f, 1920, 1080, 5760, 75.0, 0, O.YUV_420)
