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
