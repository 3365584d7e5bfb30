"""Multi-GPU use of the encode path: one process per GPU, torch.distributed for the plumbing.

Two ways to shard:

* frames  -- pictures are independent: every rank encodes its own pictures, nothing is exchanged
             on the data path (`shard_frames`, `gather_frames`).  This is what bench.py scales.
* stripes -- one picture is split into horizontal stripes of whole MCU rows, one per rank
             (BASELINE.json config 5 when the input is explicitly partitioned).  The reference
             writes no restart markers (/root/reference/src/headers.cc:242-258) and predicts DC
             across the whole picture (enc.cc:277, entropy.cc:133-136), so the stripes of a picture
             are ONE bit string: ranks exchange (1) the last quantised DCs and (2) their bit counts,
             byte-align their bits to the global offset, stuff the bytes they own, and rank 0
             concatenates, OR-merging the single byte two neighbours share.  All collectives carry
             a few bytes per picture; the payload is the compressed stripes only.

The per-rank compute is behind a small backend protocol so that the exchange/assembly logic can
be tested on CPU (gloo, world_size 2) with a stand-in backend; `GpuStripeBackend` is the product.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import Params, SjpegB200Error, lib, OK


# ------------------------------------------------------------------------------------------------
# sharding plans
# ------------------------------------------------------------------------------------------------
def shard_frames(n_frames, world_size):
    """Contiguous, balanced ranges: rank r encodes frames [out[r][0], out[r][1])."""
    base, extra = divmod(n_frames, world_size)
    out, start = [], 0
    for r in range(world_size):
        cnt = base + (1 if r < extra else 0)
        out.append((start, start + cnt))
        start += cnt
    return out


def stripe_plan(height, yuv_mode, world_size):
    """Pixel-row ranges [y0, y1) per rank: whole MCU rows (16 px for 4:2:0, else 8), balanced,
    the last stripe owning the clipped MCU row.  Ranks beyond the number of MCU rows get (H, H)."""
    mcu = 16 if yuv_mode == 1 else 8
    rows = (height + mcu - 1) // mcu
    out = []
    for (a, b) in shard_frames(rows, world_size):
        out.append((min(a * mcu, height), min(b * mcu, height)))
    return out


# ------------------------------------------------------------------------------------------------
# backends
# ------------------------------------------------------------------------------------------------
class GpuStripeBackend:
    """This rank's stripes on its GPU through the C ABI (sjb_stripes_*)."""

    def __init__(self, ctx, params):
        self.ctx, self.params = ctx, params
        self._s = C.c_void_p()
        self._n = 0
        self._shape = None      # (n, width, stripe_height) of the open session

    def header(self, width, height):
        buf = (C.c_uint8 * 2048)()
        size = C.c_size_t(0)
        rc = lib().sjb_picture_header(C.byref(self.params), width, height, buf, 2048, C.byref(size))
        if rc != OK:
            raise SjpegB200Error("sjb_picture_header rc=%d" % rc)
        return bytes(buf[:size.value])

    def transform(self, stripes, width, stripe_height, stride):
        """stripes: list of contiguous uint8 arrays (this rank's rows of each picture)."""
        n = len(stripes)
        if self._shape != (n, width, stripe_height):
            # a session owns device buffers: keep it across calls of the same geometry (creating
            # one costs several cudaMalloc / cudaFree, each a device-wide synchronisation)
            self.close()
            rc = lib().sjb_stripes_create(self.ctx._ctx, n, width, stripe_height, C.byref(self.params),
                                          C.byref(self._s))
            if rc != OK:
                raise SjpegB200Error("sjb_stripes_create rc=%d" % rc)
            self._shape = (n, width, stripe_height)
        self._n = n
        ptrs = (C.c_void_p * n)(*[s.ctypes.data for s in stripes])
        last = np.zeros((n, 3), np.int32)
        rc = lib().sjb_stripes_transform(self._s, ptrs, 0, stride, last.ctypes.data)
        if rc != OK:
            raise SjpegB200Error("sjb_stripes_transform rc=%d %s" % (rc, lib().sjb_last_error(self.ctx._ctx)))
        return last

    def code(self, dc_pred):
        bits = np.zeros(self._n, np.uint64)
        pred = np.ascontiguousarray(dc_pred, np.int32)
        rc = lib().sjb_stripes_code(self._s, pred.ctypes.data, bits.ctypes.data)
        if rc != OK:
            raise SjpegB200Error("sjb_stripes_code rc=%d" % rc)
        return bits

    def finish(self, bit_offsets, is_first, is_last, capacity):
        n = self._n
        outs = [np.empty(capacity, np.uint8) for _ in range(n)]
        optr = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
        sizes = (C.c_size_t * n)()
        head = np.zeros(n, np.uint8)
        tail = np.zeros(n, np.uint8)
        tbits = np.zeros(n, np.uint8)
        offs = np.ascontiguousarray(bit_offsets, np.uint64)
        rc = lib().sjb_stripes_finish(self._s, offs.ctypes.data, int(is_first), int(is_last), optr, capacity, sizes,
                                      head.ctypes.data, tail.ctypes.data, tbits.ctypes.data)
        if rc != OK:
            raise SjpegB200Error("sjb_stripes_finish rc=%d" % rc)
        return [outs[i][:sizes[i]].tobytes() for i in range(n)], head, tail, tbits

    def close(self):
        if self._s:
            lib().sjb_stripes_destroy(self._s)
            self._s = C.c_void_p()
            self._shape = None


# ------------------------------------------------------------------------------------------------
# collectives (work on gloo/CPU and nccl/CUDA alike)
# ------------------------------------------------------------------------------------------------
def _all_gather_array(arr, device, group=None):
    """all-gather a small fixed-shape numpy array; returns [world, ...]."""
    world = dist.get_world_size(group)
    t = torch.from_numpy(np.ascontiguousarray(arr)).to(device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t, group=group)
    return np.stack([o.cpu().numpy() for o in outs])


def _gather_bytes(payload, device, group=None):
    """Variable-length byte gather to rank 0: sizes first, then padded payloads.
    Returns list of bytes (one per rank) on rank 0, None elsewhere."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    sizes = _all_gather_array(np.array([len(payload)], np.int64), device, group)[:, 0]
    cap = int(sizes.max()) if sizes.size else 0
    buf = np.zeros(max(cap, 1), np.uint8)
    buf[:len(payload)] = np.frombuffer(payload, np.uint8)
    t = torch.from_numpy(buf).to(device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t, group=group)     # tiny payloads: the compressed stripes only
    if rank != 0:
        return None
    return [outs[r].cpu().numpy()[:int(sizes[r])].tobytes() for r in range(world)]


def gather_frames(jpegs, device="cpu", group=None):
    """Frame sharding: collect every rank's list of JPEG byte strings on rank 0 (rank order)."""
    sizes = np.array([len(j) for j in jpegs], np.int64)
    blob = b"".join(jpegs)
    counts = _all_gather_array(np.array([len(jpegs)], np.int64), device, group)[:, 0]
    nmax = int(counts.max())
    padded = np.zeros(max(nmax, 1), np.int64)
    padded[:len(sizes)] = sizes
    all_sizes = _all_gather_array(padded, device, group)
    blobs = _gather_bytes(blob, device, group)
    if blobs is None:
        return None
    out = []
    for r, b in enumerate(blobs):
        pos = 0
        for i in range(int(counts[r])):
            out.append(b[pos:pos + int(all_sizes[r, i])])
            pos += int(all_sizes[r, i])
    return out


# ------------------------------------------------------------------------------------------------
# stripe encoder
# ------------------------------------------------------------------------------------------------
def encode_striped(backend, stripes, width, height, stripe_rows, stride, device="cpu", group=None,
                   capacity=None):
    """Encodes n pictures whose rows are partitioned across the ranks of `group`.

    stripes      this rank's rows [y0, y1) of each picture (list of uint8 arrays, same shape)
    stripe_rows  (y0, y1) of this rank (stripe_plan(...)[rank])
    Returns the n complete JPEG byte strings on rank 0, None on the other ranks.
    """
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n = len(stripes)
    y0, y1 = stripe_rows
    hs = y1 - y0
    active = hs > 0
    # which ranks hold rows at all (tiny pictures may leave the last ranks empty)
    heights = _all_gather_array(np.array([hs], np.int64), device, group)[:, 0]
    holders = [r for r in range(world) if heights[r] > 0]
    is_first = active and rank == holders[0]
    is_last = active and rank == holders[-1]

    # phase 1: transform, hand the last DCs to the next stripe
    last_dc = backend.transform(stripes, width, hs, stride) if active else np.zeros((n, 3), np.int32)
    all_dc = _all_gather_array(last_dc.astype(np.int32), device, group)          # [world, n, 3]
    dc_pred = np.zeros((n, 3), np.int32)
    if active and not is_first:
        dc_pred = all_dc[holders[holders.index(rank) - 1]]

    # phase 2: entropy code, exchange bit counts -> global bit offsets
    bits = backend.code(dc_pred) if active else np.zeros(n, np.uint64)
    all_bits = _all_gather_array(bits.astype(np.int64), device, group)            # [world, n]
    offsets = np.zeros(n, np.int64)
    for r in holders:
        if r == rank:
            break
        offsets += all_bits[r]

    # phase 3: align + stuff own bytes; report the shared boundary bytes
    if capacity is None:
        capacity = max(1 << 16, 4 * width * max(hs, 1))
    if active:
        parts, head, tail, tbits = backend.finish(offsets.astype(np.uint64), is_first, is_last, capacity)
    else:
        parts, head, tail, tbits = [b""] * n, np.zeros(n, np.uint8), np.zeros(n, np.uint8), np.zeros(n, np.uint8)
    meta = np.zeros((n, 4), np.int64)
    meta[:, 0] = [len(p) for p in parts]
    meta[:, 1], meta[:, 2], meta[:, 3] = head, tail, tbits
    all_meta = _all_gather_array(meta, device, group)                              # [world, n, 4]
    blobs = _gather_bytes(b"".join(parts), device, group)
    if rank != 0:
        return None

    header = backend.header(width, height)
    return assemble_striped(header, holders, all_meta, blobs)


def assemble_striped(header, holders, all_meta, blobs):
    """Rank 0: header + stripes in rank order, OR-merging the byte two neighbours share (and
    stuffing it: bit_writer.h:172-196).  all_meta[r][i] = (size, head_byte, tail_byte, tail_bits)
    of picture i on rank r; blobs[r] = that rank's emitted bytes, pictures concatenated."""
    n = len(all_meta[holders[0]]) if holders else 0
    out = []
    pos = {r: 0 for r in holders}
    for i in range(n):
        buf = bytearray(header)
        carry, carry_bits = 0, 0
        for r in holders:
            size, h, t, tb = (int(x) for x in all_meta[r][i])
            if carry_bits:
                merged = carry | h
                buf.append(merged)
                if merged == 0xFF:
                    buf.append(0x00)
            buf += blobs[r][pos[r]:pos[r] + size]
            pos[r] += size
            carry, carry_bits = t, tb
        out.append(bytes(buf))
    return out


# ------------------------------------------------------------------------------------------------
# native path: the exchange runs inside the library over NCCL (csrc/engine_stripes.inl)
# ------------------------------------------------------------------------------------------------
def stripe_rows(height, yuv_mode, world_size, rank):
    """(y0, y1) of a rank, from the library itself (same split as stripe_plan)."""
    y0, y1 = C.c_int(0), C.c_int(0)
    rc = lib().sjb_stripe_rows(height, yuv_mode, world_size, rank, C.byref(y0), C.byref(y1))
    if rc != OK:
        raise SjpegB200Error("sjb_stripe_rows rc=%d" % rc)
    return y0.value, y1.value


class NcclStripeEncoder:
    """One per rank.  torch.distributed is used ONCE, to hand rank 0's NCCL unique id to the other
    ranks; every encode after that is a single collective call into the library, which owns its
    communicator, stream and buffers."""

    def __init__(self, ctx, group=None, single=False):
        self.ctx = ctx
        self._comm = C.c_void_p()
        if single:          # a communicator of one rank: no torch.distributed needed
            self.rank, self.world = 0, 1
            uid = np.zeros(128, np.uint8)
            if lib().sjb_comm_unique_id(uid.ctypes.data) != OK:
                raise SjpegB200Error("sjb_comm_unique_id failed (libnccl.so.2 not loadable?)")
            rc = lib().sjb_comm_create(ctx._ctx, uid.ctypes.data, 0, 1, C.byref(self._comm))
            if rc != OK:
                raise SjpegB200Error("sjb_comm_create rc=%d %s" % (rc, lib().sjb_last_error(ctx._ctx)))
            return
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        uid = np.zeros(128, np.uint8)
        if self.rank == 0:
            rc = lib().sjb_comm_unique_id(uid.ctypes.data)
            if rc != OK:
                raise SjpegB200Error("sjb_comm_unique_id rc=%d (libnccl.so.2 not loadable?)" % rc)
        holder = [uid.tobytes()]
        dist.broadcast_object_list(holder, src=0, group=group)
        uid = np.frombuffer(holder[0], np.uint8).copy()
        rc = lib().sjb_comm_create(ctx._ctx, uid.ctypes.data, self.rank, self.world, C.byref(self._comm))
        if rc != OK:
            raise SjpegB200Error("sjb_comm_create rc=%d %s" % (rc, lib().sjb_last_error(ctx._ctx)))

    def rows(self, height, yuv_mode):
        return stripe_rows(height, yuv_mode, self.world, self.rank)

    def encode(self, stripe_ptrs, on_device, width, height, stride, params, capacity, raw=False):
        """stripe_ptrs: address of row y0 of this rank's stripe of each picture.  Returns the list
        of JPEG byte strings on rank 0, None elsewhere; with raw=True the (reused) output arrays and
        the sizes instead, without copying them into Python byte strings."""
        n = len(stripe_ptrs)
        a = (C.c_void_p * n)(*stripe_ptrs)
        if self.rank == 0:
            # output buffers are kept between calls (fresh megabyte-sized numpy arrays are mmap()ed
            # and page-faulted every time: 0.3 ms per 64 pictures, more than the exchange itself)
            if getattr(self, "_outs_key", None) != (n, capacity):
                self._outs = [np.empty(capacity, np.uint8) for _ in range(n)]
                self._outs_ptrs = (C.c_void_p * n)(*[x.ctypes.data for x in self._outs])
                self._outs_key = (n, capacity)
            outs, o = self._outs, self._outs_ptrs
            sizes = (C.c_size_t * n)()
        else:
            outs, o, sizes = None, None, None
        rc = lib().sjb_stripes_encode(self._comm, n, a, int(on_device), width, height, stride, C.byref(params), o,
                                      capacity, sizes)
        if rc != OK:
            raise SjpegB200Error("sjb_stripes_encode rc=%d %s" % (rc, lib().sjb_last_error(self.ctx._ctx)))
        if self.rank != 0:
            return None
        if raw:
            return outs, [int(sizes[i]) for i in range(n)]
        return [outs[i][:sizes[i]].tobytes() for i in range(n)]

    def gather_frames(self, dev_ptrs, sizes, blob_capacity):
        """Frame sharding: every rank passes the device addresses and sizes of the JPEGs it encoded
        (sjb_encode_batch with out_on_device); returns the list of all JPEG byte strings in rank
        order on rank 0, None elsewhere."""
        n = len(dev_ptrs)
        a = (C.c_void_p * max(n, 1))(*dev_ptrs)
        sz = (C.c_size_t * max(n, 1))(*sizes)
        if self.rank == 0:
            blob = np.empty(blob_capacity, np.uint8)
            cap = 1 << 16
            out_sizes = (C.c_size_t * cap)()
            total = C.c_int(0)
            rc = lib().sjb_gather_frames(self._comm, n, a, sz, blob.ctypes.data, blob.nbytes, out_sizes, cap, C.byref(total))
        else:
            rc = lib().sjb_gather_frames(self._comm, n, a, sz, None, 0, None, 0, None)
        if rc != OK:
            raise SjpegB200Error("sjb_gather_frames rc=%d %s" % (rc, lib().sjb_last_error(self.ctx._ctx)))
        if self.rank != 0:
            return None
        out, pos = [], 0
        for i in range(total.value):
            out.append(blob[pos:pos + out_sizes[i]].tobytes())
            pos += out_sizes[i]
        return out

    def close(self):
        if self._comm:
            lib().sjb_comm_destroy(self._comm)
            self._comm = C.c_void_p()
