"""Throughput of the drop-in SjpegEncode() called concurrently from T host threads on pageable 4K
frames (one context per thread): python tools/dropin_threads.py [T ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
import oracle_lib as O
frames = [O.make_rgb("B", bench.W, bench.H, 7654321 + f) for f in range(4)]
want = O.oracle_encode(frames[0], bench.W, bench.H, 3 * bench.W, bench.QUALITY, bench.METHOD, O.YUV_420)
for t in [int(x) for x in sys.argv[1:]] or [1, 2, 4, 8, 16]:
    r = bench.run_dropin_threads(frames, want, nthreads=t, seconds=1.0)
    print("%2d threads: %6.2f Gpix/s  %.3f ms per call per thread  exact=%s" % (t, r["e2e_mpix_s"] / 1e3, r["ms_per_call_per_thread"], r["bit_exact_vs_reference"]), flush=True)
