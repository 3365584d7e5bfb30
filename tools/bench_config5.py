"""BASELINE.json config 5: 64 frames of 1920x1080 (gen B, seeds 7654321+f), q75 yuv420 method 0,
on N GPUs of one box, both shardings of SURVEY.md 8(e):
  frames  : each rank encodes 64/N whole pictures, JPEGs gathered on rank 0 (no data-path exchange)
  stripes : every picture split into N row stripes, ranks exchange DC predictors and bit offsets,
            compressed stripes gathered and merged on rank 0 (the literal config-5 wording)
Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
             --master-port 29511 tools/bench_config5.py [--reps 5]
Prints one JSON line per mode on rank 0 and checks the digest-of-digests of BASELINE.md."""
import argparse
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import oracle_lib as O  # noqa: E402
import sjpeg_b200 as S  # noqa: E402
from sjpeg_b200 import distributed as D  # noqa: E402

W, H, N = 1920, 1080, 64


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29512")
    torch.cuda.set_device(local)
    dist.init_process_group("nccl" if world > 1 else "gloo", rank=rank, world_size=world,
                            **({"device_id": torch.device("cuda", local)} if world > 1 else {}))
    device = "cuda" if world > 1 else "cpu"
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_md5.json")))["config5"]
    ctx = S.Context(local)
    params = S.default_params(75, 0, S.YUV_420)
    frames = [O.make_rgb("B", W, H, 7654321 + f) for f in range(N)]     # every rank generates all (cheap)

    def digest(jpegs):
        return hashlib.md5("".join(hashlib.md5(j).hexdigest().upper() for j in jpegs).encode()).hexdigest().upper()

    results = {}
    # ---- frames ------------------------------------------------------------------------------
    a, b = D.shard_frames(N, world)[rank]
    mine = frames[a:b]
    cap = 1 << 20
    outs = [np.empty(cap, np.uint8) for _ in mine]
    times = []
    for rep in range(args.reps + 1):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        sizes = ctx.encode_batch([f.ctypes.data for f in mine], False, W, H, 3 * W, params,
                                 [o.ctypes.data for o in outs], False, cap)
        jpegs = D.gather_frames([outs[i][:sizes[i]].tobytes() for i in range(len(mine))], device)
        dist.barrier()
        times.append(time.perf_counter() - t0)
    if rank == 0:
        results["frames"] = (min(times[1:]), digest(jpegs))
    # ---- stripes -----------------------------------------------------------------------------
    plan = D.stripe_plan(H, S.YUV_420, world)
    y0, y1 = plan[rank]
    stripes = [np.ascontiguousarray(f[y0:y1]) for f in frames]
    backend = D.GpuStripeBackend(ctx, params)
    times = []
    for rep in range(args.reps + 1):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        jpegs = D.encode_striped(backend, stripes, W, H, (y0, y1), 3 * W, device=device)
        dist.barrier()
        times.append(time.perf_counter() - t0)
    if rank == 0:
        results["stripes"] = (min(times[1:]), digest(jpegs))
        for mode, (t, dg) in results.items():
            print(json.dumps({"config": "64x1920x1080 q75 yuv420 m0 (BASELINE.json configs[4])", "sharding": mode,
                              "n_gpus": world, "seconds": round(t, 5), "Mpix_per_s": round(N * W * H / t / 1e6, 1),
                              "digest_of_digests": dg, "matches_reference": dg == gold["md5_of_md5s"],
                              "timing": "host wall clock incl. H2D from pageable numpy, D2H, collectives, assembly"}),
                  flush=True)
    backend.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
