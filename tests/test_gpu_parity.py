"""GPU tier: the CUDA path, called through the C ABI (include/sjpeg_b200.h) and through the
drop-in SjpegEncode() symbol, against the oracle on the same inputs -- bit-exact (integer
pipeline) -- and against the committed golden md5s of the compiled reference at full size."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_md5.json")))
MODES = {O.YUV_420: (16, 6), O.YUV_444: (8, 3), O.YUV_400: (8, 1)}


def _nb(w, h, mode):
    mcu, mb = MODES[mode]
    return ((w + mcu - 1) // mcu) * ((h + mcu - 1) // mcu), mb


def _oracle_coeffs(rgb, w, h, stride, mode, fmt=0):
    nm, mb = _nb(w, h, mode)
    out = np.zeros((nm * mb, 64), np.int16)
    O.oracle().sjo_image_to_coeffs(rgb.ctypes.data, w, h, stride, mode, fmt, out.ctypes.data)
    return out


def _oracle_quantised(coeffs, w, h, mode, params):
    nm, mb = _nb(w, h, mode)
    out = np.zeros_like(coeffs)
    q = np.frombuffer(bytes(params.quant), np.uint8).copy()
    mq = np.frombuffer(bytes(params.min_quant), np.uint8).copy()
    O.oracle().sjo_quantize_image(coeffs.ctypes.data, nm, mode, q.ctypes.data, mq.ctypes.data,
                                  params.q_bias, out.ctypes.data)
    return out


def _images(w, h, seed=11):
    rng = np.random.RandomState(seed)
    yield "genA", O.make_rgb("A", w, h)
    yield "genB", O.make_rgb("B", w, h)
    yield "binary", (rng.randint(0, 2, (h, w, 3)) * 255).astype(np.uint8)
    yield "noise", rng.randint(0, 256, (h, w, 3)).astype(np.uint8)


SIZES = [(512, 512), (203, 117), (1, 1), (8, 8), (17, 9), (640, 360), (1920, 1080), (48, 1000), (1000, 16)]


@pytest.mark.parametrize("mode", [O.YUV_420, O.YUV_444, O.YUV_400])
@pytest.mark.parametrize("size", SIZES, ids=lambda s: "%dx%d" % s)
def test_f1_coefficients_bit_exact(gpu_ctx, size, mode):
    """fused convert + fDCT (+ quantise) kernel vs oracle stage dumps"""
    import sjpeg_b200 as S
    w, h = size
    for name, rgb in _images(w, h):
        p = S.default_params(75, 0, mode)
        want = _oracle_coeffs(rgb, w, h, 3 * w, mode)
        got, _ = gpu_ctx.coefficients(rgb, w, h, 3 * w, p, quantise=False)
        assert np.array_equal(got, want), (name, "raw")
        for q in (75, 97):
            p = S.default_params(q, 0, mode)
            wantq = _oracle_quantised(want, w, h, mode, p)
            gotq, mask = gpu_ctx.coefficients(rgb, w, h, 3 * w, p, quantise=True)
            assert np.array_equal(gotq, wantq), (name, "quantised", q)
            chunk_nz = (wantq.reshape(-1, 8, 8) != 0).any(axis=2)      # chunk c = zig-zag 8c..8c+7
            bits = np.zeros(len(chunk_nz), np.uint8)
            for i in range(8):
                bits |= chunk_nz[:, i].astype(np.uint8) << np.uint8(i)
            assert np.array_equal(mask, bits), (name, "chunk bitmap")


@pytest.mark.parametrize("mode", [O.YUV_420, O.YUV_444, O.YUV_400])
@pytest.mark.parametrize("method", list(range(9)))
def test_whole_file_bit_exact_small(gpu_ctx, method, mode):
    import sjpeg_b200 as S
    for (w, h) in ((203, 117), (1, 1), (16, 16), (17, 9), (512, 512), (7, 33)):
        for name, rgb in _images(w, h):
            for q in (0, 50, 75, 93, 100):
                p = S.default_params(q, method, mode)
                got = gpu_ctx.encode(rgb, w, h, 3 * w, p)
                want = O.oracle_encode(rgb, w, h, 3 * w, float(q), method, mode)
                assert got == want, (w, h, name, q, method, mode, len(got), len(want))


def test_strides_and_alignment(gpu_ctx):
    """padded, unaligned and negative strides (reference tests: unit_test.cc:246-342)"""
    import sjpeg_b200 as S
    w, h = 320, 240
    rgb = O.make_rgb("A", w, h)
    for mode in (O.YUV_420, O.YUV_444, O.YUV_400):
        p = S.default_params(75, 0, mode)
        want = O.oracle_encode(rgb, w, h, 3 * w, 75.0, 0, mode)
        for pad in (0, 1, 16, 37, 4000):
            buf = np.full((h, 3 * w + pad), 0x5A, np.uint8)
            buf[:, :3 * w] = rgb.reshape(h, 3 * w)
            assert gpu_ctx.encode(buf, w, h, 3 * w + pad, p) == want, (mode, pad)
        # one byte of misalignment of the base pointer
        flat = np.zeros(rgb.size + 1, np.uint8)
        flat[1:] = rgb.ravel()
        assert gpu_ctx.encode(flat, w, h, 3 * w, p, base=flat.ctypes.data + 1) == want
        # bottom-up picture
        base = rgb.ctypes.data + (h - 1) * 3 * w
        flipped = np.ascontiguousarray(rgb[::-1])
        assert gpu_ctx.encode(rgb, w, h, -3 * w, p, base=base) == \
            O.oracle_encode(flipped, w, h, 3 * w, 75.0, 0, mode)


@pytest.mark.parametrize("size", [(203, 117), (640, 360), (1024, 96)], ids=lambda s: "%dx%d" % s)
def test_rgba_bgra_inputs(gpu_ctx, size):
    """4-byte pixels: generic path (odd strides) and the bulk-copy fast path (16-byte aligned rows)"""
    import sjpeg_b200 as S
    w, h = size
    rgb = O.make_rgb("A", w, h)
    want = {m: O.oracle_encode(rgb, w, h, 3 * w, 75.0, 4, m) for m in (O.YUV_420, O.YUV_444, O.YUV_400)}
    rgba = np.dstack([rgb, np.full((h, w, 1), 77, np.uint8)]).copy()
    bgra = np.ascontiguousarray(rgba[:, :, [2, 1, 0, 3]])
    for mode in want:
        p = S.default_params(75, 4, mode)
        p.pix_fmt = S.PIX_RGBA
        assert gpu_ctx.encode(rgba, w, h, 4 * w, p) == want[mode]
        p.pix_fmt = S.PIX_BGRA
        assert gpu_ctx.encode(bgra, w, h, 4 * w, p) == want[mode]


def test_invalid_arguments_are_refused(gpu_ctx):
    """api.cc:35-36, enc.cc:406-408, unit_test.cc:165-185,393-410"""
    import sjpeg_b200 as S
    rgb = O.make_rgb("A", 16, 16)
    p = S.default_params(75, 0, S.YUV_420)
    assert gpu_ctx.encode(rgb, 0, 16, 48, p) is None
    assert gpu_ctx.encode(rgb, 16, 0, 48, p) is None
    assert gpu_ctx.encode(rgb, 16, 16, 47, p) is None
    assert gpu_ctx.encode(rgb, 65536, 1, 3 * 65536, p) is None
    p.yuv_mode = 9
    assert gpu_ctx.encode(rgb, 16, 16, 48, p) is None
    assert S.sjpeg_encode(rgb, 16, 16, 47, 75, 0, S.YUV_420) is None
    assert S.sjpeg_encode(rgb, 16, 16, 48, 75, 0, 9) is None
    # method is clamped, not refused (unit_test.cc:605-623)
    a = S.sjpeg_encode(rgb, 16, 16, 48, 75, -1, S.YUV_420)
    b = S.sjpeg_encode(rgb, 16, 16, 48, 75, 0, S.YUV_420)
    c = S.sjpeg_encode(rgb, 16, 16, 48, 75, 9, S.YUV_420)
    d = S.sjpeg_encode(rgb, 16, 16, 48, 75, 8, S.YUV_420)
    assert a == b and c == d


def test_delta_limits_beyond_the_candidate_range_are_refused(gpu_ctx):
    """the adaptive analysis tries steps q0 - 12 .. q0 + 12; the reference asserts the limits stay
    inside (histogram.cc:179-183) and reads past its tables otherwise: refused here"""
    import sjpeg_b200 as S
    rgb = O.make_rgb("A", 64, 48)
    for (dl, dc) in ((13, 1), (12, 13), (100, 100)):
        p = S.default_params(75, 4, S.YUV_420)
        p.qdelta_max_luma, p.qdelta_max_chroma = dl, dc
        assert gpu_ctx.encode(rgb, 64, 48, 192, p) is None          # SJB_ERR_ARG
    p = S.default_params(75, 4, S.YUV_420)
    p.qdelta_max_luma, p.qdelta_max_chroma = -20, -13        # nothing to try: the starting matrices are kept
    got = gpu_ctx.encode(rgb, 64, 48, 192, p)
    want = O.oracle_encode_params(rgb, 64, 48, 192, O.SjoParams.from_buffer_copy(bytes(p)))    # same layout
    assert got == want
    # a step of 0 does not exist: zero lower bounds mean "none", zero matrix entries are raised to 1
    for method in (0, 4):
        pz = S.default_params(75, method, S.YUV_420)
        p1 = S.default_params(75, method, S.YUV_420)
        for m in range(2):
            for i in range(64):
                pz.quant[m][i] = 0
                pz.min_quant[m][i] = 0
                p1.quant[m][i] = 1
                p1.min_quant[m][i] = 1
        got = gpu_ctx.encode(rgb, 64, 48, 192, pz)
        assert got is not None and got == gpu_ctx.encode(rgb, 64, 48, 192, p1)
        assert got == O.oracle_encode_params(rgb, 64, 48, 192, O.SjoParams.from_buffer_copy(bytes(p1)))


def test_large_dimension_limits(gpu_ctx):
    """65535 is legal, 65536 is not (unit_test.cc:393-410)"""
    import sjpeg_b200 as S
    w, h = 65535, 3
    rgb = np.zeros((h, w, 3), np.uint8)
    rgb[:, ::7] = 200
    p = S.default_params(75, 0, S.YUV_420)
    assert gpu_ctx.encode(rgb, w, h, 3 * w, p) == O.oracle_encode(rgb, w, h, 3 * w, 75.0, 0, O.YUV_420)
    rgb2 = np.ascontiguousarray(rgb.transpose(1, 0, 2))
    assert gpu_ctx.encode(rgb2, h, w, 3 * h, p) == O.oracle_encode(rgb2, h, w, 3 * h, 75.0, 0, O.YUV_420)


def test_stage_histogram_and_symbol_stats(gpu_ctx):
    import sjpeg_b200 as S
    for (w, h) in ((203, 117), (512, 512)):
        rgb = O.make_rgb("A", w, h)
        for mode in (O.YUV_420, O.YUV_444, O.YUV_400):
            nm, mb = _nb(w, h, mode)
            p = S.default_params(75, 4, mode)
            coeffs = _oracle_coeffs(rgb, w, h, 3 * w, mode)
            want = np.zeros((2, 64, 129), np.int32)
            O.oracle().sjo_collect_histograms(coeffs.ctypes.data, nm, mode, want.ctypes.data)
            got = gpu_ctx.histogram(rgb, w, h, 3 * w, p)
            assert np.array_equal(got[:, :, :128], want[:, :, :128])
            zz = _oracle_quantised(coeffs, w, h, mode, p)
            want_ac = np.zeros((2, 256), np.uint32)
            want_dc = np.zeros((2, 12), np.uint32)
            O.oracle().sjo_symbol_stats(zz.ctypes.data, nm, mode, want_ac.ctypes.data, want_dc.ctypes.data)
            ac, dc = gpu_ctx.symbol_stats(rgb, w, h, 3 * w, p)
            assert np.array_equal(ac, want_ac) and np.array_equal(dc, want_dc)


def test_stage_adapted_matrices(gpu_ctx):
    """kernels A1 (histogram analysis on the device) against the oracle's restatement of
    histogram.cc:126-315 on the oracle's own histogram: the matrices that go into the DQT, for
    several pictures, sizes, qualities, modes, delta limits and minimum matrices"""
    import sjpeg_b200 as S
    rng = np.random.RandomState(11)
    cases = 0
    for (kind, w, h) in (("A", 203, 117), ("B", 512, 512), ("noise", 331, 203), ("A", 1920, 1080), ("B", 64, 48)):
        rgb = rng.randint(0, 256, (h, w, 3)).astype(np.uint8) if kind == "noise" else O.make_rgb(kind, w, h)
        for mode in (O.YUV_420, O.YUV_444, O.YUV_400):
            nm, mb = _nb(w, h, mode)
            coeffs = _oracle_coeffs(rgb, w, h, 3 * w, mode)
            counts = np.zeros((2, 64, 129), np.int32)
            O.oracle().sjo_collect_histograms(coeffs.ctypes.data, nm, mode, counts.ctypes.data)
            for (q, qdl, qdc, tol) in ((75, 12, 1, 0), (30, 12, 12, 0), (93, 4, 0, 0), (50, 12, 1, 40), (98, 12, 6, 0)):
                p = S.default_params(q, 4, mode)
                p.qdelta_max_luma, p.qdelta_max_chroma = qdl, qdc
                quant = np.array([list(p.quant[0]), list(p.quant[1])], np.uint8)
                if tol:     # a restrictive minimum matrix (EncoderParam::SetMinQuantization style)
                    minq = np.maximum(1, (quant.astype(np.int32) * (256 - tol)) >> 8).astype(np.uint8)
                    for m in range(2):
                        for i in range(64):
                            p.min_quant[m][i] = int(minq[m, i])
                else:
                    minq = np.array([list(p.min_quant[0]), list(p.min_quant[1])], np.uint8)
                want = np.ascontiguousarray(quant.copy())
                O.oracle().sjo_analyse_histo(C.c_void_p(counts.ctypes.data), 1 if mode == O.YUV_400 else 3,
                                             C.c_void_p(want.ctypes.data), C.c_void_p(minq.ctypes.data), qdl, qdc)
                want = np.maximum(want, minq)           # FinalizeQuantizer's clamp (quantize.cc:116-148)
                got = gpu_ctx.adapted_matrices(rgb, w, h, 3 * w, p)
                rows = 1 if mode == O.YUV_400 else 2
                assert np.array_equal(got[:rows], want[:rows]), (kind, w, h, mode, q, qdl, qdc, tol)
                cases += 1
    assert cases == 5 * 3 * 5


GOLD_GPU = [c for c in GOLD["cases"]]


@pytest.mark.parametrize("case", GOLD_GPU, ids=lambda c: "%s_%dx%d_q%d_m%d_y%d" % (
    c["gen"], c["w"], c["h"], c["quality"], c["method"], c["yuv_mode"]))
def test_golden_md5_full_size(gpu_ctx, case):
    """BASELINE.json configs 1-4 at full size: md5 of the compiled reference's own output"""
    import sjpeg_b200 as S
    w, h = case["w"], case["h"]
    rgb = O.make_rgb(case["gen"], w, h, case["seed"])
    data = S.sjpeg_encode(rgb, w, h, 3 * w, case["quality"], case["method"], case["yuv_mode"])
    assert data is not None
    assert len(data) == case["size"]
    assert O.md5(data) == case["md5"]


def test_config5_batch_of_frames(gpu_ctx):
    """64 x 1080p (clipped bottom MCU row), batch API; digest of digests of the reference"""
    import sjpeg_b200 as S
    c5 = GOLD["config5"]
    w, h, n = c5["w"], c5["h"], c5["frames"]
    frames = [O.make_rgb("B", w, h, 7654321 + f) for f in range(n)]
    p = S.default_params(75, 0, S.YUV_420)
    cap = 1 << 20
    outs = [np.empty(cap, np.uint8) for _ in range(n)]
    sizes = gpu_ctx.encode_batch([f.ctypes.data for f in frames], False, w, h, 3 * w, p,
                                 [o.ctypes.data for o in outs], False, cap)
    digests = [O.md5(outs[i][:sizes[i]].tobytes()) for i in range(n)]
    assert digests == c5["frame_md5"]
    assert sum(sizes) == c5["total_size"]
    assert O.md5("".join(digests).encode()) == c5["md5_of_md5s"]


def test_repeated_encodes_are_stable(gpu_ctx):
    """self-cleaning bit buffer + cached tables: alternate sizes / methods on one context"""
    import sjpeg_b200 as S
    cases = [(512, 512, 0, O.YUV_420), (203, 117, 4, O.YUV_444), (512, 512, 1, O.YUV_420), (64, 64, 7, O.YUV_400)]
    for _ in range(3):
        for (w, h, m, mode) in cases:
            rgb = O.make_rgb("A", w, h)
            assert gpu_ctx.encode(rgb, w, h, 3 * w, S.default_params(75, m, mode)) == \
                O.oracle_encode(rgb, w, h, 3 * w, 75.0, m, mode)


@pytest.mark.parametrize("parts", [2, 3, 8])
def test_row_stripes_on_one_gpu(gpu_ctx, parts):
    """BASELINE.json config 5, intra-picture striping: the three-phase stripe API (sjb_stripes_*)
    driven for every 'rank' in turn on one GPU, exchange done by hand, assembled with the product's
    own merge rule -- must equal the whole-picture encode.  (The collectives themselves are covered
    by tests/test_distributed_cpu.py over gloo and tools/bench_config5.py over NCCL.)"""
    import sjpeg_b200 as S
    from sjpeg_b200 import distributed as D
    for (w, h, mode, q) in ((1920, 1080, S.YUV_420, 75), (203, 117, S.YUV_444, 90), (320, 200, S.YUV_400, 50),
                            (640, 360, S.YUV_420, 100)):
        n = 3
        frames = [O.make_rgb("A" if i % 2 == 0 else "B", w, h, 100 + i) for i in range(n)]
        params = S.default_params(q, 0, mode)
        plan = D.stripe_plan(h, mode, parts)
        holders = [r for r in range(parts) if plan[r][1] > plan[r][0]]
        backends = {r: D.GpuStripeBackend(gpu_ctx, params) for r in holders}
        last = {}
        for r in holders:
            y0, y1 = plan[r]
            last[r] = backends[r].transform([np.ascontiguousarray(f[y0:y1]) for f in frames], w, y1 - y0, 3 * w)
        bits, prev = {}, None
        for r in holders:
            bits[r] = backends[r].code(np.zeros((n, 3), np.int32) if prev is None else last[prev])
            prev = r
        offs = np.zeros(n, np.uint64)
        meta, blobs = {}, {}
        for r in holders:
            partsb, head, tail, tb = backends[r].finish(offs, r == holders[0], r == holders[-1], 4 << 20)
            meta[r] = [(len(partsb[i]), head[i], tail[i], tb[i]) for i in range(n)]
            blobs[r] = b"".join(partsb)
            offs = offs + bits[r]
        got = D.assemble_striped(backends[holders[0]].header(w, h), holders, meta, blobs)
        for b in backends.values():
            b.close()
        for i in range(n):
            assert got[i] == O.oracle_encode(frames[i], w, h, 3 * w, float(q), 0, mode), (w, h, mode, parts, i)
        # the context must still encode whole pictures correctly afterwards
        assert gpu_ctx.encode(frames[0], w, h, 3 * w, params) == O.oracle_encode(frames[0], w, h, 3 * w, float(q), 0, mode)


# the reference's own API tests: all 17.  Riskiness / YUV_AUTO need the
# reference's generated score table, handed over through SJPEG_B200_SCORE_TABLE.
REFERENCE_TESTS_IN_SCOPE = ["InvalidArguments", "SinkFailure", "CompressionMethod", "Compress", "Dimensions",
                            "QuantMatrix", "LargeDimensions", "EncodeYUV420Strides", "EncodeYUV444Strides",
                            "EncodeNV", "NegativeStrides", "TargetSize", "AllocationFailure", "Threads",
                            "EncodeParams", "MemoryManager", "Riskiness"]


def test_reference_unit_tests_against_the_product_library(gpu_ctx):
    """tests/unit_test.cc of the reference, compiled UNMODIFIED against include/sjpeg.h and linked
    with libsjpeg_b200.so (oracle/Makefile target `conformance`; the binary travels in oracle/_ref)."""
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "unit_test_b200")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/unit_test_b200 not built (reference sources absent at build time)")
    table = O.score_table()
    if table is None:
        pytest.skip("oracle/_ref not shipped: no score table for the Riskiness test")
    table_path = os.path.join(os.path.dirname(exe), "score_table.bin")
    table.tofile(table_path)
    env = dict(os.environ, SJPEG_B200_SCORE_TABLE=table_path)
    res = subprocess.run([exe] + REFERENCE_TESTS_IN_SCOPE, capture_output=True, text=True, timeout=900, env=env)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]
    assert "%d test(s)" % len(REFERENCE_TESTS_IN_SCOPE) in res.stdout and " 0 failure(s)" in res.stdout, res.stdout


@pytest.mark.parametrize("kind", [O.KIND_YUV420, O.KIND_YUV444, O.KIND_NV12, O.KIND_NV21, O.KIND_GRAY])
def test_planar_and_semiplanar_inputs(gpu_ctx, kind):
    """sjb_encode_planar (EncodeYUV420 / YUV444 / NV12 / NV21 / Gray, encoders.cc:256-507) vs oracle"""
    import sjpeg_b200 as S
    for (w, h) in ((17, 13), (64, 48), (203, 117), (1, 1), (640, 360)):
        for q, method in ((80, 4), (30, 0), (97, 7), (75, 1)):
            planes = O.make_planes(kind, w, h, seed=w + q + kind)
            p = S.default_params(q, method, O.KIND_MODE[kind])
            got = gpu_ctx.encode_planar(*O.planar_args(kind, planes), w, h, p)
            assert got == O.oracle_encode_planar(kind, planes, w, h, q, method), (kind, w, h, q, method)


SHARP_SIZES = [(1, 1), (3, 7), (4, 4), (5, 5), (5, 4), (4, 9), (6, 5), (7, 7), (16, 16), (17, 33), (64, 48), (203, 117),
               (256, 255), (640, 481), (1030, 64), (2050, 37), (4100, 21), (8300, 9)]


def _sharp_images(w, h, seed=5):
    rng = np.random.RandomState(seed + w + h)
    yield from _images(w, h, seed)
    sat = np.zeros((h, w, 3), np.uint8)      # saturated stripes: the case the iteration exists for
    sat[:, ::2, 0] = 255
    sat[::2, :, 2] = 255
    yield "saturated", sat
    yield "gray", np.repeat(rng.randint(0, 256, (h, w, 1)), 3, axis=2).astype(np.uint8)


@pytest.mark.parametrize("size", SHARP_SIZES, ids=lambda s: "%dx%d" % s)
def test_sharp_yuv_planes_bit_exact(gpu_ctx, size):
    """sjb_sharp_yuv (import -> pipelined refinement clusters -> finish; ApplySharpYUVConversion,
    yuv_convert.cc:671-695) vs the oracle, plane by plane; widths above 512 / 1024 / 2048 chroma
    columns exercise clusters of 2, 4 and 8 CTAs, 8300 px more than one cell per thread."""
    w, h = size
    for name, rgb in _sharp_images(w, h):
        got = gpu_ctx.sharp_yuv(rgb, w, h, 3 * w)
        for plane, g, want in zip("yuv", got, O.oracle_sharp_yuv(rgb, w, h, 3 * w)):
            assert np.array_equal(g, want), (name, w, h, plane)
    # padded and negative strides
    pad = np.zeros((h, 3 * w + 7), np.uint8)
    rgb = O.make_rgb("A", w, h)
    pad[:, :3 * w] = rgb.reshape(h, 3 * w)
    want = O.oracle_sharp_yuv(rgb, w, h, 3 * w)
    for g, wv in zip(gpu_ctx.sharp_yuv(pad, w, h, pad.strides[0]), want):
        assert np.array_equal(g, wv)
    flipped = np.ascontiguousarray(pad[::-1])
    base = flipped.ctypes.data + (h - 1) * flipped.strides[0]
    for g, wv in zip(gpu_ctx.sharp_yuv(flipped, w, h, -flipped.strides[0], base=base), want):
        assert np.array_equal(g, wv)


@pytest.mark.parametrize("method", [0, 1, 4, 7])
def test_sharp_mode_whole_file(gpu_ctx, method):
    """SJPEG_YUV_SHARP through sjb_encode and the drop-in SjpegEncode() (EncoderSharp420,
    encoders.cc:512-541) vs the oracle; RGBA / BGRA go through the facade's RGB copy (api.cc:208-251)."""
    import sjpeg_b200 as S
    for (w, h) in ((3, 3), (17, 9), (203, 117), (640, 360)):
        for name, rgb in _sharp_images(w, h):
            want = O.oracle_encode(rgb, w, h, 3 * w, 75.0, method, O.YUV_SHARP)
            assert gpu_ctx.encode(rgb, w, h, 3 * w, S.default_params(75, method, S.YUV_SHARP)) == want, (name, w, h)
            assert S.sjpeg_encode(rgb, w, h, 3 * w, 75, method, S.YUV_SHARP) == want, (name, w, h)
    p = S.default_params(75, method, S.YUV_SHARP)
    p.pix_fmt = S.PIX_RGBA
    rgba = np.zeros((9, 17, 4), np.uint8)
    assert gpu_ctx.encode(rgba, 17, 9, 68, p) is None      # C ABI: sharp takes packed RGB only


def test_riskiness_and_auto_mode(gpu_ctx):
    """sjb_riskiness / SjpegRiskiness / SJPEG_YUV_AUTO (jpeg_tools.cc:177-236, encoders.cc:549-551)
    with the reference's own score table vs the oracle and the committed reference values."""
    import sjpeg_b200 as S
    table = O.score_table()
    if table is None:
        pytest.skip("oracle/_ref not shipped: no score table")
    S.set_score_table(None)
    rgb = O.make_rgb("A", 64, 48)
    assert S.lib().sjb_has_score_table() == 0
    with pytest.raises(S.SjpegB200Error):
        gpu_ctx.riskiness(rgb, 64, 48, 192)                  # loud without the table
    # the drop-in facade without a table: AUTO is refused (0 / false), never silently another mode
    assert S.sjpeg_encode(rgb, 64, 48, 192, 75, 0, S.YUV_AUTO) is None
    out = C.POINTER(C.c_uint8)()
    assert S.lib().SjpegCompress(rgb.ctypes.data, 64, 48, 75.0, C.byref(out)) == 0
    assert S.lib().SjpegRiskiness(rgb.ctypes.data, 64, 48, 192, None) == S.YUV_AUTO
    assert S.sjpeg_encode(rgb, 64, 48, 192, 75, 0, S.YUV_420) == O.oracle_encode(rgb, 64, 48, 192, 75.0, 0, O.YUV_420)
    S.set_score_table(table)
    try:
        seen = set()
        for (w, h) in ((1, 1), (2, 2), (9, 2), (2, 9), (64, 48), (203, 117), (641, 359), (1920, 1080)):
            for name, img in _sharp_images(w, h):
                want = O.oracle_riskiness(img, w, h, 3 * w, table)
                assert gpu_ctx.riskiness(img, w, h, 3 * w) == want, (name, w, h)
                risk = C.c_float()
                mode = S.lib().SjpegRiskiness(img.ctypes.data, w, h, 3 * w, C.byref(risk))
                assert (mode, risk.value) == want, (name, w, h)
                seen.add(want[0])
                if w <= 641:
                    for method in (0, 4):
                        got = S.sjpeg_encode(img, w, h, 3 * w, 75, method, S.YUV_AUTO)
                        assert got == O.oracle_encode(img, w, h, 3 * w, 75.0, method, want[0]), (name, w, h, method)
        assert seen == {O.YUV_420, O.YUV_SHARP, O.YUV_444, O.YUV_400}, seen
        for case in GOLD.get("riskiness", []):
            img = O.make_rgb(case["gen"], case["w"], case["h"], case["seed"])
            mode, risk = gpu_ctx.riskiness(img, case["w"], case["h"], 3 * case["w"])
            assert mode == case["mode"] and risk == pytest.approx(case["risk"], abs=0, rel=0), case
    finally:
        S.set_score_table(S.default_score_table())


def test_auto_mode_out_of_the_box(gpu_ctx):
    """SjpegCompress() / default EncoderParam (AUTO + method 4) with the table the library finds by
    itself next to the .so (csrc/Makefile writes it from the reference's score_7.cc at build time)"""
    import sjpeg_b200 as S
    if S.default_score_table() is None:
        pytest.skip("library built without the reference sources: no sjpeg_score_table.bin")
    S.set_score_table(S.default_score_table())
    table = S.default_score_table()
    for name, img in _sharp_images(203, 117):
        want_mode = O.oracle_riskiness(img, 203, 117, 609, table)[0]
        out = C.POINTER(C.c_uint8)()
        n = S.lib().SjpegCompress(img.ctypes.data, 203, 117, 80.0, C.byref(out))
        assert n > 0, name
        data = C.string_at(out, n)
        S.lib().SjpegFreeBuffer(out)
        assert data == O.oracle_encode(img, 203, 117, 609, 80.0, 4, want_mode), (name, want_mode)


def _api_shims():
    """oracle/ref_shim.cc only uses the PUBLIC sjpeg.h API, so the same source compiled against
    include/sjpeg.h + libsjpeg_b200.so gives a C door onto the product's C++ facade."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = os.path.join(root, "tests", "emul", "libapi_shim.so")
    subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-I", os.path.join(root, "include"), "-o", so,
                    os.path.join(root, "oracle", "ref_shim.cc"), "-L", os.path.join(root, "sjpeg_b200"),
                    "-lsjpeg_b200", "-Wl,-rpath," + os.path.join(root, "sjpeg_b200")], check=True)
    prod = C.CDLL(so)
    ref = O.ref()
    for L in (prod, ref):
        L.ref_encode_meta.restype = C.c_size_t
        L.ref_encode_meta.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float] + \
            [C.c_char_p, C.c_size_t] * 4 + [C.c_int, C.POINTER(C.POINTER(C.c_uint8))]
        L.ref_encode_param.restype = C.c_size_t
        L.ref_encode_param.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_float] + \
            [C.c_int] * 6 + [C.POINTER(C.POINTER(C.c_uint8))]
    return prod, ref


def _call_meta(L, free, rgb, w, h, mode, q, exif=b"", iccp=b"", xmp=b"", app=b"", split=0):
    out = C.POINTER(C.c_uint8)()
    n = L.ref_encode_meta(rgb.ctypes.data, w, h, 3 * w, mode, q, exif, len(exif), iccp, len(iccp), xmp, len(xmp),
                          app, len(app), split, C.byref(out))
    if n == 0:
        return None
    data = C.string_at(out, n)
    free(out)
    return data


@pytest.mark.skipif(O.ref() is None, reason="oracle/_ref not shipped")
def test_cpp_facade_params_and_metadata_equal_reference(gpu_ctx):
    """sjpeg::Encode with an EncoderParam through the product's C++ facade vs the compiled
    reference: custom matrices / bias / deltas / trellis flag (api.cc:145-181) and the metadata
    segments EXIF, ICC (multi-chunk), XMP, extended XMP with its MD5 GUID, raw app markers
    (headers.cc:63-180)."""
    import sjpeg_b200 as S
    prod, ref = _api_shims()
    w, h = 203, 117
    rgb = O.make_rgb("A", w, h)
    rng = np.random.RandomState(9)
    # --- EncoderParam fields ---
    quant = rng.randint(1, 120, (2, 64)).astype(np.uint8)
    for mode in (O.YUV_420, O.YUV_444, O.YUV_400):
        for (hf, ad, tr, bias, dl, dc) in ((1, 1, 0, -1, -1, -1), (0, 0, 0, 0x60, -1, -1), (1, 1, 1, 0x78, 6, 3),
                                           (0, 1, 0, 0x90, 12, 12), (1, 0, 1, -1, -1, -1)):
            outs = []
            for L, free in ((prod, S.lib().SjpegFreeBuffer), (ref, ref.SjpegFreeBuffer)):
                out = C.POINTER(C.c_uint8)()
                n = L.ref_encode_param(rgb.ctypes.data, w, h, 3 * w, mode, quant.ctypes.data, 85.0, hf, ad, tr, bias,
                                       dl, dc, C.byref(out))
                outs.append(C.string_at(out, n) if n else None)
                if n:
                    free(out)
            assert outs[0] is not None and outs[0] == outs[1], (mode, hf, ad, tr, bias, dl, dc)
    # --- metadata ---
    note = b'<x:xmpmeta xmpNote:HasExtendedXMP="' + b"0" * 32 + b'" >'
    big_xmp = note + bytes(rng.randint(32, 127, 150000).astype(np.uint8))
    cases = [dict(exif=b"II*\x00" + bytes(rng.randint(0, 256, 300).astype(np.uint8))),
             dict(iccp=bytes(rng.randint(0, 256, 150000).astype(np.uint8))),
             dict(xmp=b"<x:xmpmeta>small</x:xmpmeta>"),
             dict(xmp=big_xmp), dict(xmp=big_xmp, split=40000),
             dict(app=b"\xff\xe5\x00\x06ABCD"),
             dict(exif=b"E" * 10, iccp=b"I" * 70000, xmp=b"X" * 100, app=b"\xff\xe7\x00\x04zz"),
             dict(exif=b"E" * 70000),                     # too large: both must refuse
             dict(xmp=b"Y" * 70000)]                      # extended without the note: both must refuse
    for kw in cases:
        a = _call_meta(prod, S.lib().SjpegFreeBuffer, rgb, w, h, O.YUV_420, 75.0, **kw)
        b = _call_meta(ref, ref.SjpegFreeBuffer, rgb, w, h, O.YUV_420, 75.0, **kw)
        assert a == b, {k: (len(v) if isinstance(v, bytes) else v) for k, v in kw.items()}


@pytest.mark.skipif(O.ref() is None, reason="oracle/_ref not shipped")
def test_target_size_and_psnr_search_equal_reference(gpu_ctx):
    """EncoderParam::passes > 1 (Encoder::LoopScan, dichotomy.cc:113-205): same bytes, same final q
    and measured value as the compiled reference, for size and PSNR targets."""
    import sjpeg_b200 as S
    prod, ref = _api_shims()
    for L in (prod, ref):
        L.ref_encode_search.restype = C.c_size_t
        L.ref_encode_search.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_float,
                                        C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float),
                                        C.POINTER(C.c_float), C.POINTER(C.POINTER(C.c_uint8))]
    w, h = 203, 117
    for gen in ("A", "B"):
        rgb = O.make_rgb(gen, w, h)
        for mode in (O.YUV_420, O.YUV_444, O.YUV_400):
            for (tmode, tval, passes, tol) in ((1, 6000.0, 8, 1.0), (1, 2500.0, 12, 0.5), (2, 36.0, 6, 1.0),
                                               (2, 42.0, 10, 0.2), (0, 0.0, 3, 1.0)):
                for (hf, ad, tr) in ((1, 1, 0), (0, 0, 0), (1, 0, 0), (0, 1, 0), (1, 1, 1)):
                    res = []
                    for L, free in ((prod, S.lib().SjpegFreeBuffer), (ref, ref.SjpegFreeBuffer)):
                        out = C.POINTER(C.c_uint8)()
                        q, v = C.c_float(-1), C.c_float(-1)
                        n = L.ref_encode_search(rgb.ctypes.data, w, h, 3 * w, mode, 70.0, tmode, tval, passes, tol, hf,
                                                ad, tr, C.byref(q), C.byref(v), C.byref(out))
                        res.append((C.string_at(out, n) if n else None, q.value, v.value))
                        if n:
                            free(out)
                    assert res[0][0] is not None and res[0][0] == res[1][0], (gen, mode, tmode, tval, passes, hf, ad, tr,
                                                                              len(res[0][0] or b""), len(res[1][0] or b""))
                    assert res[0][1:] == res[1][1:], (gen, mode, tmode, tval, res[0][1:], res[1][1:])


def test_concurrent_host_threads(gpu_ctx):
    """The library is re-entrant like the reference (unit_test.cc:114-131): 8 host threads call the
    drop-in SjpegEncode() at the same time (one lazily created GPU context per thread)."""
    import threading
    import sjpeg_b200 as S
    cases = [(512, 512, 0, O.YUV_420), (203, 117, 4, O.YUV_444), (640, 360, 7, O.YUV_420), (320, 200, 1, O.YUV_400)]
    want = {}
    imgs = {}
    for (w, h, m, mode) in cases:
        imgs[(w, h)] = O.make_rgb("A", w, h)
        want[(w, h, m, mode)] = O.oracle_encode(imgs[(w, h)], w, h, 3 * w, 75.0, m, mode)
    errors = []

    def work(t):
        try:
            for rep in range(6):
                w, h, m, mode = cases[(t + rep) % len(cases)]
                got = S.sjpeg_encode(imgs[(w, h)], w, h, 3 * w, 75, m, mode)
                if got != want[(w, h, m, mode)]:
                    errors.append((t, rep, w, h, m, mode))
        except Exception as e:
            errors.append((t, repr(e)))

    ths = [threading.Thread(target=work, args=(t,)) for t in range(8)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert not errors, errors
