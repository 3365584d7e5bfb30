"""CPU tier: the product's own host code and per-block __host__ __device__ functions
(sjpeg_b200/csrc/block_ops.cuh, host_codec.cc) compiled by g++ into a CPU emulation of the kernel
pipeline (tests/emul) and compared with the oracle; plus the C-ABI library's load/exports."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMUL_DIR = os.path.join(ROOT, "tests", "emul")
_u8p = C.POINTER(C.c_uint8)


@pytest.fixture(scope="module")
def emul():
    so = os.path.join(EMUL_DIR, "libemul.so")
    if not (os.environ.get("SJB_EMUL_REUSE") == "1" and os.path.exists(so)):   # a child run reuses the parent's build
        tmp = so + ".%d.tmp" % os.getpid()
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", tmp,
                        os.path.join(EMUL_DIR, "emul_main.cc"),
                        os.path.join(ROOT, "sjpeg_b200", "csrc", "host_codec.cc")], check=True)
        os.replace(tmp, so)      # atomic: a process that has the old file mapped keeps it
    E = C.CDLL(so)
    E.emul_encode.restype = C.c_size_t
    E.emul_encode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_longlong, C.c_int, C.c_int, C.c_int,
                              C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(_u8p)]
    E.emul_free.argtypes = [_u8p]
    E.emul_coeffs.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_longlong, C.c_int, C.c_int, C.c_void_p]
    E.emul_sharp_yuv.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p]
    E.emul_analyse_histo.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    E.emul_analyse_histo_device_form.argtypes = E.emul_analyse_histo.argtypes
    E.emul_riskiness.restype = C.c_int
    E.emul_riskiness.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_longlong, C.c_void_p, C.POINTER(C.c_float)]
    return E


def _emul_encode(E, rgb, w, h, q, m, mode):
    quant = np.zeros((2, 64), np.uint8)
    O.oracle().sjo_quality_to_matrices(q, quant.ctypes.data)
    out = _u8p()
    n = E.emul_encode(rgb.ctypes.data, w, h, 3 * w, mode, 0, m, quant.ctypes.data, 0x78, 12, 1, C.byref(out))
    data = C.string_at(out, n)
    E.emul_free(out)
    return data


def test_block_functions_and_host_codec_match_oracle(emul):
    rng = np.random.RandomState(3)
    for (w, h) in ((203, 117), (1, 1), (17, 9), (64, 48), (7, 33), (256, 128)):
        imgs = [O.make_rgb("A", w, h), O.make_rgb("B", w, h),
                (rng.randint(0, 2, (h, w, 3)) * 255).astype(np.uint8),
                rng.randint(0, 256, (h, w, 3)).astype(np.uint8)]
        for rgb in imgs:
            for q in (0, 25, 75, 93, 100):
                for mode in (O.YUV_420, O.YUV_444, O.YUV_400):
                    for m in (0, 1, 3, 4, 7, 8):
                        assert _emul_encode(emul, rgb, w, h, float(q), m, mode) == \
                            O.oracle_encode(rgb, w, h, 3 * w, float(q), m, mode), (w, h, q, mode, m)


def test_raw_coefficients_match_oracle(emul):
    w, h = 203, 117
    rgb = (np.random.RandomState(5).randint(0, 2, (h, w, 3)) * 255).astype(np.uint8)   # extreme swings
    for mode, mcu, mb in ((O.YUV_420, 16, 6), (O.YUV_444, 8, 3), (O.YUV_400, 8, 1)):
        nb = ((w + mcu - 1) // mcu) * ((h + mcu - 1) // mcu) * mb
        a = np.zeros((nb, 64), np.int16)
        b = np.zeros((nb, 64), np.int16)
        emul.emul_coeffs(rgb.ctypes.data, w, h, 3 * w, mode, 0, a.ctypes.data)
        O.oracle().sjo_image_to_coeffs(rgb.ctypes.data, w, h, 3 * w, mode, 0, b.ctypes.data)
        assert np.array_equal(a, b)


def test_library_loads_and_exports_every_declared_symbol():
    import sjpeg_b200
    L = sjpeg_b200.lib()
    header = open(os.path.join(ROOT, "include", "sjpeg_b200.h")).read()
    declared = set(re.findall(r"\b(sjb_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations found"
    for name in declared:
        assert hasattr(L, name), name
    assert declared <= set(sjpeg_b200.exported_symbols())
    assert L.sjb_version() == 0x000101 == L.SjpegVersion()
    # C entry points of include/sjpeg.h
    for name in ("SjpegEncode", "SjpegCompress", "SjpegFreeBuffer", "SjpegDimensions", "SjpegFindQuantizer",
                 "SjpegEstimateQuality", "SjpegQuantMatrix", "SjpegRiskiness"):
        assert hasattr(L, name), name


def test_sharp_yuv_and_riskiness_cell_functions_match_oracle(emul):
    """sharp_ops.cuh (the per-cell code of sharp.cu) run on the CPU in the kernels' own data layout
    -- one state copy per iteration, exit rule applied afterwards -- equals the oracle's in-place loop."""
    rng = np.random.RandomState(3)
    table = O.score_table()
    if table is None:   # the compiled reference did not travel: any table exercises the same code
        table = rng.randint(0, 40, 343 * 343).astype(np.uint8)
    for (w, h) in ((1, 1), (4, 9), (5, 5), (6, 5), (7, 8), (16, 16), (33, 21), (130, 67)):
        sat = np.zeros((h, w, 3), np.uint8)
        sat[:, ::2, 0] = 255
        sat[::2, :, 2] = 255
        for img in (O.make_rgb("A", w, h), O.make_rgb("B", w, h), rng.randint(0, 256, (h, w, 3)).astype(np.uint8),
                    (rng.randint(0, 2, (h, w, 3)) * 255).astype(np.uint8), sat):
            cw, ch = (w + 1) // 2, (h + 1) // 2
            y, u, v = np.zeros((h, w), np.uint8), np.zeros((ch, cw), np.uint8), np.zeros((ch, cw), np.uint8)
            emul.emul_sharp_yuv(img.ctypes.data, w, h, 3 * w, y.ctypes.data, u.ctypes.data, v.ctypes.data)
            for got, want in zip((y, u, v), O.oracle_sharp_yuv(img, w, h, 3 * w)):
                assert np.array_equal(got, want), (w, h)
            risk = C.c_float()
            mode = emul.emul_riskiness(img.ctypes.data, w, h, 3 * w, table.ctypes.data, C.byref(risk))
            assert (mode, risk.value) == O.oracle_riskiness(img, w, h, 3 * w, table), (w, h)


def test_entropy_walk_order_is_a_luma_first_permutation(emul):
    """block_ops.cuh::walk_order (which block of a tile each worker of the entropy kernel walks when
    the tile is busy): a bijection on the tile that lists the luma blocks first, for every phase of
    the tile start inside an MCU and for short last tiles."""
    emul.emul_walk_order.restype = C.c_uint
    emul.emul_walk_order.argtypes = [C.c_uint, C.c_uint, C.c_uint, C.c_int]
    for mb, lb in ((6, 4), (3, 1), (1, 1)):
        for first in list(range(0, 256 * 7, 256)) + [256 * 12345, 4294966784]:
            for count in (256, 255, 100, 7, 2, 1):
                js = [emul.emul_walk_order(first, count, i, mb) for i in range(count)]
                assert sorted(js) == list(range(count)), (mb, first, count)
                chroma = [((first + j) % mb) >= lb for j in js]
                assert chroma == sorted(chroma), (mb, first, count)
                if mb == 1:
                    assert js == list(range(count))


def test_coefficient_layout_is_a_sector_interleaved_bijection(emul):
    """block_ops.cuh coef_block_base / coef_pos_offset / coef_chunk_index: every (block, position)
    owns one int16 of the padded array, chunks are 16-byte aligned runs of 8 positions, and the
    first sectors of four consecutive blocks share one 128-byte line."""
    for f in (emul.emul_coef_offset, emul.emul_coef_chunk_offset, emul.emul_coef_padded_blocks):
        f.restype = C.c_ulonglong
    emul.emul_coef_offset.argtypes = [C.c_ulonglong, C.c_int]
    emul.emul_coef_chunk_offset.argtypes = [C.c_ulonglong, C.c_int]
    emul.emul_coef_padded_blocks.argtypes = [C.c_ulonglong]
    for nb in (1, 2, 3, 4, 5, 6, 7, 8, 9, 48, 49, 50, 51):
        padded = emul.emul_coef_padded_blocks(nb)
        assert padded % 4 == 0 and nb <= padded < nb + 4
        offs = [emul.emul_coef_offset(g, p) for g in range(nb) for p in range(64)]
        assert len(set(offs)) == len(offs) and max(offs) < padded * 64
    for g in (0, 1, 2, 3, 4, 7, 194399, 201326591):
        for c in range(8):
            base = emul.emul_coef_chunk_offset(g, c)
            assert base % 8 == 0
            assert [emul.emul_coef_offset(g, 8 * c + k) for k in range(8)] == list(range(base, base + 8))
        # sector s of blocks 4j..4j+3 = one 128-byte line (64 int16)
        for s in range(4):
            line = {emul.emul_coef_offset(4 * (g // 4) + b, 16 * s + k) // 64 for b in range(4) for k in range(16)}
            assert len(line) == 1
    # block index arithmetic is 32-bit: the largest picture (65535 x 65535, 4:4:4) still fits
    assert emul.emul_coef_offset(3 * 8192 * 8192 - 1, 63) == (3 * 8192 * 8192) * 64 - 1


def test_histogram_analysis_baseline_isa_path(emul):
    """The same comparison in a fresh process with SJPEG_B200_NO_AVX2=1, so that the non-AVX2 build of the
    inner sums (what a CPU without AVX2 would run) is exercised as well."""
    env = dict(os.environ, SJPEG_B200_NO_AVX2="1", SJB_EMUL_REUSE="1", PYTHONPATH=os.path.join(ROOT, "tests"))
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", os.path.join(ROOT, "tests", "test_host_logic.py"),
                        "-k", "equals_oracle_on_hard_histograms"], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "1 passed" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_histogram_analysis_equals_oracle_on_hard_histograms(emul):
    """host_codec.cc::AnalyseHistograms (int64 sums over squeezed bins, one branch-free term) against the
    oracle's line-by-line restatement of histogram.cc:126-315, on histograms chosen to hit what the
    whole-file tests rarely do: counts large enough for the 32-bit products to wrap, single-bin and
    full-width rows, tiny and saturated matrices, restricted minima, every delta limit."""
    rng = np.random.default_rng(20261017)
    lib = O.oracle()
    cases = 0
    for trial in range(160):
        counts = np.zeros((2, 64, 129), np.int64)
        kind = trial % 8
        for m in range(2):
            for pos in range(64):
                width = int(rng.integers(1, 129))
                if kind == 0:      # geometric, photographic
                    row = (rng.integers(1, 200000) * np.exp(-rng.uniform(0.02, 0.6) * np.arange(width))).astype(np.int64)
                elif kind == 1:    # flat and huge: products wrap
                    row = rng.integers(0, 40_000_000, width)
                elif kind == 2:    # single bin
                    row = np.zeros(width, np.int64); row[-1] = rng.integers(1, 1 << 30)
                elif kind == 3:    # sparse spikes
                    row = np.where(rng.random(width) < 0.1, rng.integers(1, 5_000_000, width), 0)
                elif kind == 4:    # nearly empty (density rule)
                    row = np.where(rng.random(width) < 0.02, 1, 0)
                elif kind == 5:    # bimodal
                    row = rng.integers(0, 3000, width) + np.where(np.arange(width) > width // 2, 50000, 0)
                elif kind == 6:    # small counts
                    row = rng.integers(0, 4, width)
                else:              # everything at once
                    row = rng.integers(0, 1 << int(rng.integers(1, 31)), width)
                counts[m, pos, :width] = row
        c32 = np.ascontiguousarray(counts.astype(np.int32))
        for quant_kind in range(3):
            if quant_kind == 0:
                quant = rng.integers(1, 256, (2, 64)).astype(np.uint8)
            elif quant_kind == 1:
                quant = rng.integers(1, 14, (2, 64)).astype(np.uint8)
            else:
                quant = rng.integers(240, 256, (2, 64)).astype(np.uint8)
            minq = np.minimum(quant, rng.integers(1, 256, (2, 64))).astype(np.uint8) if trial % 3 else np.ones((2, 64), np.uint8)
            for (qdl, qdc) in ((12, 1), (0, 0), (5, 12), (12, 12), (1, 0)):
                for comps in (3, 1):
                    a = np.ascontiguousarray(quant.copy()); b = np.ascontiguousarray(quant.copy())
                    emul.emul_analyse_histo(c32.ctypes.data, comps, a.ctypes.data, minq.ctypes.data, qdl, qdc)
                    lib.sjo_analyse_histo(C.c_void_p(c32.ctypes.data), comps, C.c_void_p(b.ctypes.data),
                                          C.c_void_p(minq.ctypes.data), qdl, qdc)
                    assert a.tobytes() == b.tobytes(), (trial, quant_kind, qdl, qdc, comps)
                    # the form the DEVICE runs (block_ops.cuh aq_*, decomposed like the kernels)
                    d = np.ascontiguousarray(quant.copy())
                    emul.emul_analyse_histo_device_form(c32.ctypes.data, comps, d.ctypes.data, minq.ctypes.data, qdl, qdc)
                    assert d.tobytes() == b.tobytes(), ("device form", trial, quant_kind, qdl, qdc, comps)
                    cases += 1
    assert cases == 160 * 3 * 5 * 2


def test_hot_kernels_keep_their_resource_budget():
    """ptxas report of the build (csrc/kernels.ptxas.log, written by the Makefile): the residency the
    design counts on -- 16 CTAs per SM of the 4:2:0 F1 kernel (128 registers, no spill), 5 CTAs of
    the entropy kernel (40 registers, no spill, below 45 KB of shared memory) -- is a build-time
    property; a source change that breaks it shows up here, not as a slower bench line."""
    log = os.path.join(ROOT, "sjpeg_b200", "csrc", "kernels.ptxas.log")
    if not os.path.exists(log):
        pytest.skip("no ptxas log (library not built by the Makefile here)")
    text = open(log).read()
    entries = {}
    for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores,"
                         r" (\d+) bytes spill loads\n.*Used (\d+) registers(?:.*?(\d+) bytes smem)?", text):
        entries[m.group(1)] = dict(stack=int(m.group(2)), spill=int(m.group(3)) + int(m.group(4)), regs=int(m.group(5)),
                                   smem=int(m.group(6) or 0))
    f1 = [v for k, v in entries.items() if "f1_fast_kernelILi1ELb0ELi0E" in k]
    ent = [v for k, v in entries.items() if "entropy_pack_kernel" in k]
    assert len(f1) == 1 and len(ent) == 1, sorted(entries)
    assert f1[0]["regs"] <= 128 and f1[0]["spill"] == 0 and 16 * (f1[0]["smem"] + 1024) <= 228 * 1024, f1
    assert ent[0]["regs"] <= 40 and ent[0]["spill"] == 0 and 5 * (ent[0]["smem"] + 1024) <= 228 * 1024, ent
    # the trellis keeps its node arrays in shared memory: no stack frame (local memory), no spill
    tr = [v for k, v in entries.items() if "trellis_kernel" in k]
    assert len(tr) == 1 and tr[0]["stack"] == 0 and tr[0]["spill"] == 0 and tr[0]["regs"] <= 128, tr


def test_host_helpers_without_gpu():
    import sjpeg_b200
    L = sjpeg_b200.lib()
    for q in (0, 10, 50, 75, 93, 100):
        p = sjpeg_b200.default_params(q, 4, sjpeg_b200.YUV_420)
        ref = np.zeros((2, 64), np.uint8)
        O.oracle().sjo_quality_to_matrices(float(q), ref.ctypes.data)
        assert bytes(p.quant[0]) == ref[0].tobytes() and bytes(p.quant[1]) == ref[1].tobytes()
        m = np.zeros(64, np.uint8)
        L.SjpegQuantMatrix(float(q), False, m.ctypes.data)
        assert m.tobytes() == ref[0].tobytes()
        assert L.SjpegEstimateQuality(m.ctypes.data, False) == pytest.approx(q if q > 0 else 0, abs=1)
    # parsers on an oracle-made file
    data = O.oracle_encode(O.make_rgb("A", 33, 21), 33, 21, 99, 75.0, 0, O.YUV_420)
    w, h, is420 = C.c_int(), C.c_int(), C.c_int()
    assert L.SjpegDimensions(data, len(data), C.byref(w), C.byref(h), C.byref(is420))
    assert (w.value, h.value, is420.value) == (33, 21, 1)
    qm = np.zeros((2, 64), np.uint8)
    assert L.SjpegFindQuantizer(data, len(data), qm.ctypes.data) == 2
    ref = np.zeros((2, 64), np.uint8)
    O.oracle().sjo_quality_to_matrices(75.0, ref.ctypes.data)
    assert np.array_equal(qm, ref)


def test_no_silent_cpu_fallback():
    """Without a device the compute entry points must fail loudly, never produce bytes."""
    import sjpeg_b200
    if sjpeg_b200.lib().sjb_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(sjpeg_b200.SjpegB200Error):
        sjpeg_b200.Context(0)
    rgb = O.make_rgb("A", 16, 16)
    assert sjpeg_b200.sjpeg_encode(rgb, 16, 16, 48, 75, 0, sjpeg_b200.YUV_420) is None


def test_host_stager_threading_is_race_free_and_exact():
    """csrc/host_stager.cc (threaded pinned-ring upload of pageable pictures) against a stand-in CUDA
    runtime (tests/emul/fake_cuda), under ThreadSanitizer: 4 owners x 40 uploads of awkward sizes,
    every byte checked, no data race reported."""
    exe = os.path.join(EMUL_DIR, "stager_tsan")
    subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=thread", "-I", os.path.join(EMUL_DIR, "fake_cuda"),
                    "-o", exe, os.path.join(EMUL_DIR, "stager_main.cc"),
                    os.path.join(ROOT, "sjpeg_b200", "csrc", "host_stager.cc"), "-lpthread"], check=True)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "wrong bytes: 0" in res.stdout
    assert "ThreadSanitizer" not in res.stderr, res.stderr[-4000:]


def test_score_table_is_found_next_to_the_library():
    """SJPEG_YUV_AUTO out of the box: csrc/Makefile leaves the reference's generated riskiness table
    next to the .so where the reference sources are present; a fresh process finds it without any
    call or environment variable, an explicit clear removes it (then AUTO fails loudly)."""
    import subprocess
    import sys
    side = os.path.join(ROOT, "sjpeg_b200", "sjpeg_score_table.bin")
    code = ("import sjpeg_b200 as S; L = S.lib(); a = L.sjb_has_score_table(); S.set_score_table(None); "
            "print(a, L.sjb_has_score_table())")
    env = {k: v for k, v in os.environ.items() if k != "SJPEG_B200_SCORE_TABLE"}
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert out.stdout.split() == (["1", "0"] if os.path.exists(side) else ["0", "0"]), out.stdout
    if os.path.exists(side) and O.ref() is not None:
        assert np.array_equal(np.fromfile(side, np.uint8), O.score_table())
