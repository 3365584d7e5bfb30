"""GPU tier, groups of pictures per launch: the paths bench.py times (device-resident input and
output, 8 pictures per launch, several lanes) and the batch API with per-picture quantiser tables,
Huffman tables and headers inside one launch (methods >= 1), on sparse, busy and mixed content.
Every output is compared byte for byte with the oracle."""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu

_want_cache = {}


def _want(key, rgb, w, h, q, method, mode):
    k = (key, w, h, q, method, mode)
    if k not in _want_cache:
        _want_cache[k] = O.oracle_encode(rgb, w, h, 3 * w, float(q), method, mode)
    return _want_cache[k]


def _frame(kind, w, h, seed):
    if kind in ("A", "B"):
        return O.make_rgb(kind, w, h, seed)
    rng = np.random.RandomState(seed)
    if kind == "noise":
        return rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
    if kind == "flat":
        return np.full((h, w, 3), seed & 255, np.uint8)
    raise ValueError(kind)


def _content(name, n, w, h):
    """list of (key, frame): n pictures of one kind, or kinds mixed inside every group"""
    kinds = {"A": ["A"], "B": ["B"], "noise": ["noise"], "mixed": ["A", "noise", "B", "flat", "B", "A"]}[name]
    out = []
    for i in range(n):
        kind = kinds[i % len(kinds)]
        seed = 1000 + i * 17 + len(name)
        out.append(((kind, seed), _frame(kind, w, h, seed)))
    return out


def _encode_batch_device(ctx, frames, w, h, params, cap):
    """device-resident pixels in, device-resident JPEGs out (pix_on_device = out_on_device = 1)"""
    import torch
    dev = [torch.from_numpy(f.reshape(-1)).cuda() for f in frames]
    outs = [torch.zeros(cap, dtype=torch.uint8, device="cuda") for _ in frames]
    torch.cuda.synchronize()
    sizes = ctx.encode_batch([t.data_ptr() for t in dev], True, w, h, 3 * w, params,
                             [t.data_ptr() for t in outs], True, cap)
    torch.cuda.synchronize()
    return [outs[i][:sizes[i]].cpu().numpy().tobytes() for i in range(len(frames))]


def _encode_batch_host(ctx, frames, w, h, params, cap):
    outs = [np.empty(cap, np.uint8) for _ in frames]
    sizes = ctx.encode_batch([f.ctypes.data for f in frames], False, w, h, 3 * w, params,
                             [o.ctypes.data for o in outs], False, cap)
    return [outs[i][:sizes[i]].tobytes() for i in range(len(frames))]


@pytest.mark.parametrize("shape", [(3840, 2160, 16), (1920, 1080, 17), (1920, 1080, 40)], ids=["16x4K", "17x1080p", "40x1080p"])
def test_device_resident_batches(gpu_ctx, shape):
    """the configuration of the headline number: groups of 16 device-resident pictures per launch,
    JPEGs left in device memory; 17 pictures leave a ragged last group of one"""
    import sjpeg_b200 as S
    w, h, n = shape
    frames = [O.make_rgb("B", w, h, 7654321 + f) for f in range(n)]
    p = S.default_params(75, 0, S.YUV_420)
    got = _encode_batch_device(gpu_ctx, frames, w, h, p, 4 << 20)
    for i, f in enumerate(frames):
        assert got[i] == O.oracle_encode(f, w, h, 3 * w, 75.0, 0, O.YUV_420), (w, h, i)


def test_bench_device_outputs(gpu_ctx):
    """sjb_bench_device is what bench.py times: its outputs (sjb_bench_output) after several rounds
    over rotating lanes equal the oracle's, for a sparse and a busy picture set, methods 0 and 4"""
    import torch
    import sjpeg_b200 as S
    w, h, n = 1920, 1080, 16
    for kind, method in (("B", 0), ("A", 0), ("A", 4), ("noise", 1)):
        frames = [_frame(kind, w, h, 4242 + f) for f in range(n)]
        dev = [torch.from_numpy(f.reshape(-1)).cuda() for f in frames]
        p = S.default_params(75, method, S.YUV_420)
        for iters in (1, 3):
            gpu_ctx.bench_device([t.data_ptr() for t in dev], w, h, 3 * w, p, iters)
            for i, f in enumerate(frames):
                assert gpu_ctx.bench_output(i) == _want((kind, 4242 + i), f, w, h, 75, method, O.YUV_420), \
                    (kind, method, iters, i)


@pytest.mark.parametrize("content", ["A", "B", "noise", "mixed"])
@pytest.mark.parametrize("mode", [O.YUV_420, O.YUV_444, O.YUV_400])
@pytest.mark.parametrize("method", [1, 4, 6, 7])
def test_batch_matrix_methods_modes_content(gpu_ctx, method, mode, content):
    """per-picture matrices / Huffman tables / headers inside one launch (gridDim.y > 1), host and
    device-resident groups, group sizes with a ragged tail, several lanes in flight"""
    import sjpeg_b200 as S
    w, h = (331, 203) if content != "noise" else (200, 120)
    p = S.default_params(75, method, mode)
    for n in (3, 9, 17, 40):
        items = _content(content, n, w, h)
        frames = [f for _, f in items]
        want = [_want(k, f, w, h, 75, method, mode) for k, f in items]
        got = _encode_batch_device(gpu_ctx, frames, w, h, p, 1 << 20)
        assert got == want, (method, mode, content, n, "device", [i for i in range(n) if got[i] != want[i]])
        if n in (3, 9):
            got = _encode_batch_host(gpu_ctx, frames, w, h, p, 1 << 20)
            assert got == want, (method, mode, content, n, "host", [i for i in range(n) if got[i] != want[i]])


@pytest.mark.parametrize("method", [0, 4, 7])
def test_busy_batches_multi_iteration_stuffing(gpu_ctx, method):
    """noise and gen-A 1080p pictures: streams of several hundred KB to MBs per picture, so the
    persistent stuffing CTAs run several tiles each while other lanes hold SM slots; blocks longer
    than the 512-bit slots of the entropy kernel; 9 pictures = groups of 8 + 1 (device) or 2 (host)"""
    import sjpeg_b200 as S
    w, h, n = 1920, 1080, 9
    kinds = ["noise", "A", "noise", "B", "A", "noise", "noise", "A", "noise"]
    frames = [_frame(k, w, h, 900 + i) for i, k in enumerate(kinds)]
    p = S.default_params(90 if method == 0 else 75, method, S.YUV_420)
    q = 90 if method == 0 else 75
    want = [_want((kinds[i], 900 + i), f, w, h, q, method, O.YUV_420) for i, f in enumerate(frames)]
    got = _encode_batch_device(gpu_ctx, frames, w, h, p, 16 << 20)
    assert got == want, [i for i in range(n) if got[i] != want[i]]
    got = _encode_batch_host(gpu_ctx, frames, w, h, p, 16 << 20)
    assert got == want, [i for i in range(n) if got[i] != want[i]]


@pytest.mark.parametrize("method", [0, 7])
def test_8k_busy_batch(gpu_ctx, method):
    """three 8K gen-A pictures (groups of 2 at this size): ~1 700 stuffing tiles per picture"""
    import sjpeg_b200 as S
    w, h, n = 7680, 4320, 3
    frames = [O.make_rgb("A", w, h, 31 + f) for f in range(n)]
    p = S.default_params(75, method, S.YUV_420)
    got = _encode_batch_device(gpu_ctx, frames, w, h, p, 48 << 20)
    for i, f in enumerate(frames):
        assert got[i] == O.oracle_encode(f, w, h, 3 * w, 75.0, method, O.YUV_420), (method, i)


def test_batch_error_leaves_no_copies_in_flight(gpu_ctx):
    """a too-small output capacity is reported (sizes still filled) and the context stays usable"""
    import sjpeg_b200 as S
    w, h, n = 640, 360, 5
    frames = [O.make_rgb("A", w, h, 5 + f) for f in range(n)]
    p = S.default_params(75, 0, S.YUV_420)
    outs = [np.empty(64, np.uint8) for _ in frames]
    with pytest.raises(S.SjpegB200Error):
        gpu_ctx.encode_batch([f.ctypes.data for f in frames], False, w, h, 3 * w, p,
                             [o.ctypes.data for o in outs], False, 64)
    got = _encode_batch_host(gpu_ctx, frames, w, h, p, 1 << 20)
    for i, f in enumerate(frames):
        assert got[i] == O.oracle_encode(f, w, h, 3 * w, 75.0, 0, O.YUV_420)


@pytest.mark.parametrize("method", [0, 1, 3, 4, 7])
def test_native_stripe_session_single_rank(gpu_ctx, method):
    """sjb_stripes_encode with a one-rank NCCL communicator: the whole stream-ordered sequence
    (all-reduce of histograms / symbol counts, all-gathers, offsets, stuffing, compaction, assembly)
    runs with degenerate collectives; 11 pictures = one chunk, 37 = three.  Multi-rank runs: bench.py
    --gpus N (config 5) and tools/stripes_nccl.py."""
    import sjpeg_b200 as S
    from sjpeg_b200 import distributed as D
    enc = D.NcclStripeEncoder(gpu_ctx, single=True)
    try:
        for (w, h, mode, q) in ((640, 360, S.YUV_420, 75), (203, 117, S.YUV_444, 90), (64, 200, S.YUV_400, 50)):
            kinds = ["A", "B", "noise", "flat"]
            frames = [_frame(kinds[i % 4], w, h, 50 + i) for i in range(11)]
            p = S.default_params(q, method, mode)
            assert enc.rows(h, mode) == (0, h)
            got = enc.encode([f.ctypes.data for f in frames], False, w, h, 3 * w, p, 1 << 20)
            for i, f in enumerate(frames):
                assert got[i] == O.oracle_encode(f, w, h, 3 * w, float(q), method, mode), (w, h, mode, method, i)
        # several chunks (one group of 16 pictures each, pipelined: the next chunk's uploads are queued before
        # this chunk's exchange): 37 pictures = 16 + 16 + 5, outputs left in the encoder's host buffers
        w, h = 203, 117
        frames = [_frame(["A", "B", "noise", "flat"][i % 4], w, h, 700 + i) for i in range(37)]
        p = S.default_params(75, method, S.YUV_420)
        outs, sizes = enc.encode([f.ctypes.data for f in frames], False, w, h, 3 * w, p, 1 << 18, raw=True)
        for i, f in enumerate(frames):
            assert outs[i][:sizes[i]].tobytes() == O.oracle_encode(f, w, h, 3 * w, 75.0, method, S.YUV_420), (method, i)
        # the context still encodes whole pictures afterwards
        rgb = O.make_rgb("A", 320, 200)
        assert gpu_ctx.encode(rgb, 320, 200, 960, S.default_params(75, 4, S.YUV_420)) == \
            O.oracle_encode(rgb, 320, 200, 960, 75.0, 4, O.YUV_420)
    finally:
        enc.close()


@pytest.mark.parametrize("kind", [O.KIND_YUV420, O.KIND_YUV444, O.KIND_NV12, O.KIND_NV21, O.KIND_GRAY])
def test_planar_fast_path_large_pictures(gpu_ctx, kind):
    """planar / semi-planar sources through the bulk-copy F1 kernel (16-byte aligned planes): sizes
    with several tiles per MCU row, a partial last tile, an odd MCU column and a clipped bottom
    row left to the generic kernel; methods with and without the raw-coefficient pass"""
    import sjpeg_b200 as S
    for (w, h) in ((1920, 1080), (1040, 200), (1296, 72), (272, 528)):
        for q, method in ((75, 0), (90, 4), (60, 7)):
            planes = O.make_planes(kind, w, h, seed=w + q + kind, pad=(16 - w % 16 if w % 16 else 0, 0, 0))
            # make every plane's stride a multiple of 16 so that the fast kernel is eligible
            for name in ("y", "u", "v"):
                a = planes[name]
                if a is not None and a.strides[0] % 16:
                    padded = np.zeros((a.shape[0], (a.shape[1] + 15) // 16 * 16), np.uint8)
                    padded[:, :a.shape[1]] = a
                    planes[name] = padded
            p = S.default_params(q, method, O.KIND_MODE[kind])
            got = gpu_ctx.encode_planar(*O.planar_args(kind, planes), w, h, p)
            assert got == O.oracle_encode_planar(kind, planes, w, h, q, method), (kind, w, h, q, method)


def test_sharp_4k_uses_planar_fast_path(gpu_ctx):
    """SJPEG_YUV_SHARP at 4K: conversion on the device, then the planar 4:2:0 encoder on aligned planes"""
    import sjpeg_b200 as S
    w, h = 3840, 2160
    rgb = O.make_rgb("A", w, h, 99)
    got = gpu_ctx.encode(rgb, w, h, 3 * w, S.default_params(75, 4, S.YUV_SHARP))
    assert got == O.oracle_encode(rgb, w, h, 3 * w, 75.0, 4, O.YUV_SHARP)


def test_batch_auto_and_sharp_modes(gpu_ctx):
    """sjb_encode_batch with SJB_YUV_AUTO (the batched SjpegCompress: riskiness per picture, then one
    sub-batch per mode) and SJB_YUV_SHARP (several conversions in flight, then the planar pipeline):
    a batch whose pictures land in all four modes, host and device-resident input."""
    import torch
    import sjpeg_b200 as S
    table = S.default_score_table() if O.score_table() is None else O.score_table()
    if table is None:
        pytest.skip("no riskiness score table available")
    S.set_score_table(table)
    try:
        w, h = 352, 208
        rng = np.random.RandomState(4)
        sat = np.zeros((h, w, 3), np.uint8)
        sat[:, ::2, 0] = 255
        sat[::2, :, 2] = 255
        gray = np.repeat(rng.randint(0, 256, (h, w, 1)), 3, axis=2).astype(np.uint8)
        base = [O.make_rgb("A", w, h, 3), O.make_rgb("B", w, h, 4), sat, gray, rng.randint(0, 256, (h, w, 3)).astype(np.uint8),
                (rng.randint(0, 2, (h, w, 3)) * 255).astype(np.uint8)]
        frames = [base[i % len(base)] if i < 2 * len(base) else O.make_rgb("A", w, h, 100 + i) for i in range(21)]
        modes = [O.oracle_riskiness(f, w, h, 3 * w, table)[0] for f in frames]
        assert len(set(modes)) >= 3, modes            # the batch really is split
        for method in (4, 0):
            p = S.default_params(80, method, S.YUV_AUTO)
            want = [O.oracle_encode(f, w, h, 3 * w, 80.0, method, m) for f, m in zip(frames, modes)]
            assert _encode_batch_host(gpu_ctx, frames, w, h, p, 1 << 20) == want, ("auto host", method)
            assert _encode_batch_device(gpu_ctx, frames, w, h, p, 1 << 20) == want, ("auto device", method)
            p = S.default_params(80, method, S.YUV_SHARP)
            want = [O.oracle_encode(f, w, h, 3 * w, 80.0, method, O.YUV_SHARP) for f in frames[:9]]
            assert _encode_batch_host(gpu_ctx, frames[:9], w, h, p, 1 << 20) == want, ("sharp host", method)
        # sharp at a size the aligned planar fast path takes (width % 32 == 0), more pictures than streams
        w, h = 640, 368
        frames = [O.make_rgb("A", w, h, 7 + i) for i in range(6)]
        p = S.default_params(75, 4, S.YUV_SHARP)
        want = [O.oracle_encode(f, w, h, 3 * w, 75.0, 4, O.YUV_SHARP) for f in frames]
        assert _encode_batch_device(gpu_ctx, frames, w, h, p, 1 << 20) == want
    finally:
        S.set_score_table(S.default_score_table())


@pytest.mark.parametrize("kind", [O.KIND_YUV420, O.KIND_NV12, O.KIND_NV21, O.KIND_YUV444, O.KIND_GRAY])
def test_planar_batch(gpu_ctx, kind):
    """sjb_encode_planar_batch: planar / semi-planar pictures in groups (host planes with arbitrary strides,
    and device-resident aligned planes as a video decoder leaves them), methods 0 and 4"""
    import torch
    import sjpeg_b200 as S
    for (w, h, n) in ((640, 368, 19), (203, 117, 5)):
        sets = [O.make_planes(kind, w, h, seed=50 + i, pad=(5, 3, 7) if w == 203 else (0, 0, 0)) for i in range(n)]
        args = [O.planar_args(kind, pl) for pl in sets]
        for method in (0, 4):
            p = S.default_params(75, method, O.KIND_MODE[kind])
            want = [O.oracle_encode_planar(kind, pl, w, h, 75, method) for pl in sets]
            outs = [np.empty(1 << 20, np.uint8) for _ in range(n)]
            ys = [a[0] for a in args]
            us = [a[2] for a in args] if kind != O.KIND_GRAY else None
            vs = [a[4] for a in args] if kind != O.KIND_GRAY else None
            sizes = gpu_ctx.encode_planar_batch(ys, args[0][1], us, args[0][3], vs, args[0][5], args[0][6], False, w, h, p,
                                                [o.ctypes.data for o in outs], False, 1 << 20)
            got = [outs[i][:sizes[i]].tobytes() for i in range(n)]
            assert got == want, (kind, w, h, method, "host")
            if w % 32 == 0:      # device-resident planes
                dev = []
                for pl in sets:
                    dev.append({k: (torch.from_numpy(v).cuda() if v is not None else None) for k, v in pl.items()})
                torch.cuda.synchronize()
                ys = [d["y"].data_ptr() for d in dev]
                if kind == O.KIND_GRAY:
                    us = vs = None
                elif kind in (O.KIND_NV12, O.KIND_NV21):
                    us = [d["u"].data_ptr() + (0 if kind == O.KIND_NV12 else 1) for d in dev]
                    vs = [d["u"].data_ptr() + (1 if kind == O.KIND_NV12 else 0) for d in dev]
                else:
                    us = [d["u"].data_ptr() for d in dev]
                    vs = [d["v"].data_ptr() for d in dev]
                douts = [torch.zeros(1 << 20, dtype=torch.uint8, device="cuda") for _ in range(n)]
                sizes = gpu_ctx.encode_planar_batch(ys, args[0][1], us, args[0][3], vs, args[0][5], args[0][6], True, w, h, p,
                                                    [t.data_ptr() for t in douts], True, 1 << 20)
                torch.cuda.synchronize()
                got = [douts[i][:sizes[i]].cpu().numpy().tobytes() for i in range(n)]
                assert got == want, (kind, w, h, method, "device")


def test_random_batches(gpu_ctx):
    """seeded random batches: number of pictures (ragged groups, more groups than lanes), geometry, content
    mixed inside the batch, method, mode, quality, device-resident or host buffers.  One-off longer runs:
    SJB_BATCH_FUZZ_CASES=200 SJB_BATCH_FUZZ_SEED=5 python -m pytest tests/test_gpu_batch.py -m gpu -k random_batches"""
    import os
    import sjpeg_b200 as S
    cases = int(os.environ.get("SJB_BATCH_FUZZ_CASES", "10"))
    rng = np.random.RandomState(int(os.environ.get("SJB_BATCH_FUZZ_SEED", "424242")))
    kinds = ["A", "B", "noise", "flat"]
    for case in range(cases):
        n = int(rng.choice([1, 2, 3, 5, 16, 17, 31, 50, 97, 130]))
        w, h = int(rng.randint(1, 400)), int(rng.randint(1, 260))
        mode = int(rng.choice([S.YUV_420, S.YUV_444, S.YUV_400]))
        method = int(rng.randint(0, 9))
        q = int(rng.choice([5, 50, 75, 90, 98]))
        frames = [_frame(kinds[int(rng.randint(0, 4))], w, h, int(rng.randint(1, 1 << 30))) for _ in range(n)]
        p = S.default_params(q, method, mode)
        want = [O.oracle_encode(f, w, h, 3 * w, float(q), method, mode) for f in frames]
        cap = max(len(x) for x in want) + 4096
        if rng.randint(0, 2):
            got = _encode_batch_device(gpu_ctx, frames, w, h, p, cap)
        else:
            got = _encode_batch_host(gpu_ctx, frames, w, h, p, cap)
        assert got == want, (case, n, w, h, mode, method, q, [i for i in range(n) if got[i] != want[i]][:5])
