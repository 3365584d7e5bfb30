"""Row stripes over NCCL inside the library (sjb_stripes_encode), N ranks of one box:
correctness of every method class / yuv mode against the oracle, then BASELINE.json config 5
(64 x 1920x1080 gen B q75 4:2:0 m0) as frames and as stripes.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29513 tools/stripes_nccl.py [--reps 5] [--method 0]"""
import argparse, ctypes as C, hashlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist
import oracle_lib as O
import sjpeg_b200 as S
from sjpeg_b200 import distributed as D

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--method", type=int, default=0)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29513")
torch.cuda.set_device(local)
dist.init_process_group("gloo", rank=rank, world_size=world)      # bootstrap channel only
ctx = S.Context(local)
enc = D.NcclStripeEncoder(ctx)

# ---- correctness: all method classes, modes, pictures with fewer MCU rows than ranks ----
bad = 0
rng = np.random.RandomState(5)
for (w, h, mode, q) in ((640, 360, S.YUV_420, 75), (203, 117, S.YUV_444, 90), (64, 40, S.YUV_400, 50), (48, 16, S.YUV_420, 30),
                        (8, 64, S.YUV_400, 75)):
    kinds = ["A", "B", "noise", "flat"]
    frames = []
    for i in range(11):
        k = kinds[i % 4]
        frames.append(O.make_rgb(k, w, h, 90 + i) if k in "AB" else (rng.randint(0, 256, (h, w, 3)).astype(np.uint8) if k == "noise"
                      else np.full((h, w, 3), 37 * i & 255, np.uint8)))
    y0, y1 = enc.rows(h, mode)
    mine = [np.ascontiguousarray(f[y0:y1]) if y1 > y0 else np.zeros((1, w, 3), np.uint8) for f in frames]
    for method in (0, 1, 3, 4, 7):
        p = S.default_params(q, method, mode)
        got = enc.encode([m.ctypes.data for m in mine], False, w, h, 3 * w, p, 1 << 20)
        if rank == 0:
            for i, f in enumerate(frames):
                if got[i] != O.oracle_encode(f, w, h, 3 * w, float(q), method, mode):
                    bad += 1
                    print("MISMATCH", w, h, mode, method, i, len(got[i]), flush=True)
if rank == 0:
    print(json.dumps({"check": "stripes over NCCL vs oracle, methods 0/1/3/4/7 x 5 geometries x 11 pictures", "n_gpus": world,
                      "mismatches": bad}), flush=True)

# ---- frames sharded across the ranks, gathered on rank 0 inside the library ----
frames = [O.make_rgb("A" if i % 2 else "B", 320, 200, 300 + i) for i in range(13)]
a, b = D.shard_frames(len(frames), world)[rank]
p = S.default_params(75, 4, S.YUV_420)
dev = [torch.from_numpy(f.reshape(-1)).cuda() for f in frames[a:b]]
douts = [torch.zeros(1 << 18, dtype=torch.uint8, device="cuda") for _ in dev]
torch.cuda.synchronize()
sizes = ctx.encode_batch([t.data_ptr() for t in dev], True, 320, 200, 960, p, [t.data_ptr() for t in douts], True, 1 << 18) if dev else []
got = enc.gather_frames([t.data_ptr() for t in douts], sizes, 8 << 20)
if rank == 0:
    ok = got == [O.oracle_encode(f, 320, 200, 960, 75.0, 4, O.YUV_420) for f in frames]
    print(json.dumps({"check": "frames sharded over ranks + sjb_gather_frames (NCCL) vs oracle", "n_gpus": world, "ok": ok}), flush=True)

# ---- config 5 ----
W, H, N = 1920, 1080, 64
gold = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_md5.json")))["config5"]
frames = [O.make_rgb("B", W, H, 7654321 + f) for f in range(N)]
p = S.default_params(75, args.method, S.YUV_420)
def digest(jpegs):
    return hashlib.md5("".join(hashlib.md5(j).hexdigest().upper() for j in jpegs).encode()).hexdigest().upper()
def pin(arr):
    ptr = S.lib().sjb_host_alloc(arr.nbytes); C.memmove(ptr, arr.ctypes.data, arr.nbytes); return ptr
res = {}
a, b = D.shard_frames(N, world)[rank]
mine = [pin(f) for f in frames[a:b]]
cap = 1 << 20
outs = [S.lib().sjb_host_alloc(cap) for _ in mine]
times = []
for rep in range(args.reps + 1):
    dist.barrier(); t0 = time.perf_counter()
    sizes = ctx.encode_batch(mine, False, W, H, 3 * W, p, outs, False, cap)
    dist.barrier(); times.append(time.perf_counter() - t0)
res["frames (each rank its own pictures, JPEGs stay on their rank)"] = min(times[1:])
y0, y1 = enc.rows(H, S.YUV_420)
stripes = [pin(np.ascontiguousarray(f[y0:y1])) for f in frames]
times = []
for rep in range(args.reps + 1):
    dist.barrier(); t0 = time.perf_counter()
    jpegs = enc.encode(stripes, False, W, H, 3 * W, p, cap)
    dist.barrier(); times.append(time.perf_counter() - t0)
res["stripes (NCCL inside the library, complete JPEGs on rank 0)"] = min(times[1:])
if rank == 0:
    ok = digest(jpegs) == gold["md5_of_md5s"] if args.method == 0 else all(
        jpegs[i] == O.oracle_encode(frames[i], W, H, 3 * W, 75.0, args.method, O.YUV_420) for i in (0, 17, 63))
    for k, t in res.items():
        print(json.dumps({"config5": k, "n_gpus": world, "method": args.method, "seconds": round(t, 5),
                          "Mpix_per_s": round(N * W * H / t / 1e6, 1), "stripes_bit_exact": ok}), flush=True)
enc.close()
dist.destroy_process_group()
