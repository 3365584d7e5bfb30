"""CPU tier, world_size 2 and 3 over gloo: the multi-rank host logic of sjpeg_b200/distributed.py
(sharding plans, DC / bit-offset exchange, byte alignment, boundary-byte merge, gathers) with a
CPU stand-in for the per-rank compute (the oracle's stripe coder) -- the assembled files must equal
the oracle's whole-picture encode.  The same code path runs on NCCL with GpuStripeBackend
(tests/test_gpu_parity.py::test_striped_single_rank and tools/bench_config5.py)."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as O

_u8p = C.POINTER(C.c_uint8)


class OracleStripeBackend:
    """Test double with the contract of GpuStripeBackend: raw bits from the oracle, then the
    alignment / stuffing / shared-byte rules restated in numpy."""

    def __init__(self, quality=75.0, yuv_mode=O.YUV_420):
        self.p = O.SjoParams()
        O.oracle().sjo_default_params(C.byref(self.p), quality, 0, yuv_mode)
        self.quality, self.mode = quality, yuv_mode

    def header(self, width, height):
        whole = O.oracle_encode(np.zeros((height, width, 3), np.uint8), width, height, 3 * width, self.quality, 0,
                                self.mode)
        return whole[:whole.index(b"\xff\xda") + (14 if self.mode != O.YUV_400 else 10)]

    def _run(self, stripe, width, hs, stride, pred):
        L = O.oracle()
        L.sjo_encode_stripe.restype = C.c_size_t
        L.sjo_encode_stripe.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(O.SjoParams), C.c_void_p,
                                        C.c_void_p, C.POINTER(_u8p), C.POINTER(C.c_uint64)]
        pred = np.ascontiguousarray(pred, np.int32)
        last = np.zeros(3, np.int32)
        out, nbits = _u8p(), C.c_uint64(0)
        n = L.sjo_encode_stripe(stripe.ctypes.data, width, hs, stride, C.byref(self.p), pred.ctypes.data,
                                last.ctypes.data, C.byref(out), C.byref(nbits))
        data = C.string_at(out, n)
        L.sjo_free(out)
        return last, data, nbits.value

    def transform(self, stripes, width, hs, stride):
        self.stripes, self.geom = stripes, (width, hs, stride)
        return np.stack([self._run(s, width, hs, stride, np.zeros(3))[0] for s in stripes])

    def code(self, dc_pred):
        w, hs, stride = self.geom
        self.raw = [self._run(s, w, hs, stride, dc_pred[i])[1:] for i, s in enumerate(self.stripes)]
        return np.array([r[1] for r in self.raw], np.uint64)

    def finish(self, bit_offsets, is_first, is_last, capacity):
        parts, head, tail, tbits = [], [], [], []
        for (data, nbits), off in zip(self.raw, bit_offsets):
            s = int(off) & 7
            bits = np.unpackbits(np.frombuffer(data, np.uint8))[:nbits]
            r = np.concatenate([np.zeros(s, np.uint8), bits])
            end = len(r)
            if is_last and end % 8:
                r = np.concatenate([r, np.ones(8 - end % 8, np.uint8)])     # pad with 1-bits
            full = len(r) // 8
            by = np.packbits(r[:full * 8])
            t_bits = 0 if is_last else end % 8
            t_byte = int(np.packbits(np.concatenate([r[full * 8:], np.zeros(8 - t_bits, np.uint8)]))[0]) if t_bits else 0
            b0 = 1 if (s and not is_first) else 0
            h_byte = int(by[0]) if b0 else 0
            body = bytearray()
            for b in by[b0:]:
                body.append(int(b))
                if b == 0xFF:
                    body.append(0)
            if is_last:
                body += b"\xff\xd9"
            parts.append(bytes(body)); head.append(h_byte); tail.append(t_byte); tbits.append(t_bits)
        return parts, np.array(head, np.uint8), np.array(tail, np.uint8), np.array(tbits, np.uint8)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, case, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from sjpeg_b200 import distributed as D
        w, h, mode, quality, n = case
        frames = [O.make_rgb("A" if i % 2 == 0 else "B", w, h, 100 + i) for i in range(n)]
        # --- stripes -------------------------------------------------------------------------
        plan = D.stripe_plan(h, mode, world)
        y0, y1 = plan[rank]
        stripes = [np.ascontiguousarray(f[y0:y1]) for f in frames]
        res = D.encode_striped(OracleStripeBackend(quality, mode), stripes, w, h, (y0, y1), 3 * w, device="cpu")
        # --- frames --------------------------------------------------------------------------
        a, b = D.shard_frames(n, world)[rank]
        mine = [O.oracle_encode(frames[i], w, h, 3 * w, quality, 0, mode) for i in range(a, b)]
        gathered = D.gather_frames(mine, device="cpu")
        if rank == 0:
            want = [O.oracle_encode(f, w, h, 3 * w, quality, 0, mode) for f in frames]
            q.put(("ok", res == want, gathered == want, [len(x) for x in res], [len(x) for x in want]))
    except Exception as e:   # surface the failure in the parent
        if rank == 0:
            q.put(("err", repr(e)))
        raise
    finally:
        dist.destroy_process_group()


CASES = [(203, 117, O.YUV_420, 75.0, 3), (64, 48, O.YUV_444, 90.0, 2), (320, 200, O.YUV_400, 50.0, 2),
         (48, 40, O.YUV_420, 100.0, 2)]


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("case", CASES, ids=lambda c: "%dx%d_y%d_q%d" % (c[0], c[1], c[2], c[3]))
def test_stripe_and_frame_sharding_over_gloo(world, case):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    msg = q.get(timeout=5)
    assert msg[0] == "ok", msg
    assert msg[1], ("striped output differs from the whole-picture encode", msg[3], msg[4])
    assert msg[2], "frame-sharded gather differs"


def test_plans():
    from sjpeg_b200 import distributed as D
    assert D.shard_frames(64, 8) == [(8 * r, 8 * r + 8) for r in range(8)]
    assert D.shard_frames(5, 3) == [(0, 2), (2, 4), (4, 5)]
    plan = D.stripe_plan(1080, O.YUV_420, 8)           # 68 MCU rows: 9,9,9,9,8,8,8,8
    assert plan[0] == (0, 144) and plan[-1] == (960, 1080)
    assert all(a % 16 == 0 for a, _ in plan) and [b for _, b in plan[:-1]] == [a for a, _ in plan[1:]]
    assert D.stripe_plan(40, O.YUV_420, 8)[3:] == [(40, 40)] * 5


# ---- the library's own host-side pieces of the native (NCCL) stripe path ----------------------
def test_stripe_rows_match_the_python_plan():
    """sjb_stripe_rows (C++, used by sjb_stripes_encode on every rank) == distributed.stripe_plan"""
    from sjpeg_b200 import distributed as D
    for height in (1, 7, 8, 16, 17, 135, 1080, 2160, 4320, 65535):
        for mode in (O.YUV_420, O.YUV_444, O.YUV_400):
            for world in (1, 2, 3, 4, 8):
                plan = D.stripe_plan(height, mode, world)
                for r in range(world):
                    assert D.stripe_rows(height, mode, world, r) == plan[r], (height, mode, world, r)


def _split_like_the_stuffing_kernel(scan_bits, cuts):
    """What each rank's stuffing kernel reports for its stripe (kernels.cu stuff_kernel, stripe mode):
    scan_bits = the unstuffed, padded scan as a 0/1 array (multiple of 8 long); cuts = bit positions
    where one stripe ends and the next begins.  Returns [(bytes, flags)] in stripe order."""
    R = np.packbits(scan_bits)
    bounds = [0] + list(cuts) + [len(scan_bits)]
    out = []
    for r in range(len(bounds) - 1):
        a, b = bounds[r], bounds[r + 1]
        first, last = r == 0, r == len(bounds) - 2
        shift, A = a % 8, a // 8
        end_bits = shift + (b - a)
        b0 = 1 if (shift != 0 and not first) else 0
        b1 = (end_bits + 7) // 8 if last else end_bits // 8
        own = np.zeros(len(scan_bits), np.uint8)
        own[a:b] = scan_bits[a:b]
        mine = np.packbits(own)                       # the stripe's bits alone, zeros elsewhere
        body = bytearray()
        for x in mine[A + b0:A + max(b1, b0)]:
            body.append(int(x))
            if x == 0xFF:
                body.append(0)
        if last:
            body += b"\xff\xd9"
        head = int(mine[A]) if b0 == 1 else 0
        tail_bits = 0 if last else end_bits % 8
        tail = int(mine[A + b1]) if tail_bits else 0
        head_open = 1 if (b0 == 1 and not last and end_bits < 8) else 0
        out.append((bytes(body), head | (tail << 8) | (tail_bits << 16) | (head_open << 24)))
    return out


def test_library_assembles_stripes_incl_several_in_one_byte():
    """sjb_stripes_assemble (rank 0 of the native path): an oracle file cut at arbitrary bit positions
    -- including stripes of one or two bits, three and more meeting inside one byte, cuts on byte
    boundaries and next to 0xFF bytes -- must come back byte for byte."""
    import sjpeg_b200 as S
    L = S.lib()
    rng = np.random.RandomState(12)
    for (w, h, gen, q) in ((64, 48, "A", 75.0), (33, 21, "B", 50.0), (8, 8, "A", 100.0), (120, 16, "N", 95.0)):
        rgb = rng.randint(0, 256, (h, w, 3)).astype(np.uint8) if gen == "N" else O.make_rgb(gen, w, h)
        whole = O.oracle_encode(rgb, w, h, 3 * w, q, 0, O.YUV_420)
        sos = whole.index(b"\xff\xda")
        header, scan = whole[:sos + 14], whole[sos + 14:-2]
        raw = scan.replace(b"\xff\x00", b"\xff")
        bits = np.unpackbits(np.frombuffer(raw, np.uint8))
        nbits = len(bits)
        for trial in range(60):
            k = int(rng.randint(1, 9))
            if trial % 3 == 0:       # clusters of tiny stripes
                base = int(rng.randint(1, nbits - 40))
                cuts = sorted(set(int(base + d) for d in np.cumsum(rng.randint(1, 4, k))))
            elif trial % 3 == 1:     # byte-aligned and near-aligned cuts
                cuts = sorted(set(int(8 * rng.randint(1, nbits // 8) + rng.randint(-1, 2)) for _ in range(k)))
            else:
                cuts = sorted(set(int(c) for c in rng.randint(1, nbits, k)))
            cuts = [c for c in cuts if 0 < c < nbits]
            parts = _split_like_the_stuffing_kernel(bits, cuts)
            n = len(parts)
            bufs = [np.frombuffer(p, np.uint8).copy() if p else np.zeros(1, np.uint8) for p, _ in parts]
            ptrs = (C.c_void_p * n)(*[b.ctypes.data for b in bufs])
            sizes = (C.c_size_t * n)(*[len(p) for p, _ in parts])
            flags = (C.c_uint * n)(*[f for _, f in parts])
            out = np.zeros(len(whole) + 64, np.uint8)
            size = C.c_size_t(0)
            rc = L.sjb_stripes_assemble(header, len(header), n, ptrs, sizes, flags, out.ctypes.data, out.nbytes, C.byref(size))
            assert rc == 0
            assert out[:size.value].tobytes() == whole, (w, h, gen, trial, cuts)
