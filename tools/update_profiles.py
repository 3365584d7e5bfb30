"""Turns what tools/collect_profiles.sh left in gpurun_out/r02/ into the files kept under profiles/
(run here after the gpurun call; no GPU needed):  python tools/update_profiles.py"""
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out", "r02")
DST = os.path.join(ROOT, "profiles")


def summary(name, command):
    rep = os.path.join(SRC, name + "_full.ncu-rep")
    if not os.path.exists(rep):
        rep = os.path.join(SRC, name + "_full_raw.csv")      # report dropped on the box to fit gpurun's size limit
    if not os.path.exists(rep):
        return None
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep, command],
                         capture_output=True, text=True).stdout
    open(os.path.join(DST, "r02_%s_ncu_summary.json" % name), "w").write(out)
    return json.loads(out)


def lines(name, kernel, top=45):
    rep = os.path.join(SRC, name + "_full.ncu-rep")
    if not os.path.exists(rep):
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_lines.py"), rep, kernel,
                          os.path.join(ROOT, "sjpeg_b200", "libsjpeg_b200.so"), str(top)], capture_output=True, text=True).stdout
    open(os.path.join(DST, "r02_%s_%s_lines.txt" % (name, kernel.split("_")[0])), "w").write(
        "per source line of %s in the capture %s_full (tools/sass_lines.py): warp instructions executed, threads per\n"
        "instruction, stall samples and the three most frequent stall reasons\n%s" % (kernel, name, out))


cmds = {"f1_420": "tools/run_f1.py 16 3 f1 (16 x 4K gen B, 4:2:0)", "f1_444": "tools/run_f1.py 16 3 f1 3840 2160 3 0 B (4:4:4)",
        "f1_planar": "tools/planar_bench.py (one 4K YUV420 planar picture)", "es_genB": "tools/run_f1.py 16 3 full (gen B)",
        "es_genA": "tools/run_f1.py 16 3 full 3840 2160 1 0 A (gen A)", "m4": "tools/run_f1.py 16 2 full 3840 2160 1 4 B (method 4)",
        "trellis": "tools/one_encode.py A 7680 4320 75 7 1 2 (8K gen A, method 7, two pictures)"}
sums = {k: summary(k, "see tools/collect_profiles.sh: " + v) for k, v in cmds.items()}
lines("es_genB", "entropy_pack")
lines("es_genA", "entropy_pack")
lines("trellis", "trellis_kernel")
lines("m4", "histogram_kernel", 25)
f1 = sums.get("f1_420")
if f1:
    k = f1["kernels"][0]
    rd = float(k["dram__bytes_read.sum"].split()[0]) * 1e6
    wr = float(k["dram__bytes_write.sum"].split()[0]) * 1e6
    json.dump({"kernel": "f1_fast_kernel<420>", "pictures_per_launch": 16, "picture": "3840x2160 gen B",
               "dram_bytes_read": int(rd), "dram_bytes_write": int(wr), "dram_bytes_per_launch": int(rd + wr),
               "algorithmic_bytes_per_launch": 16 * 49766400, "gpu_time_us": k["gpu__time_duration.sum"],
               "source": "profiles/r02_f1_420_ncu_summary.json (ncu --set full, cold L2, one launch of the shape bench.py times)"},
              open(os.path.join(DST, "f1_traffic.json"), "w"), indent=1)
for f, t in (("bench_n1.json", "r02_bench_n1.json"), ("bench_reference_n1.json", "r02_bench_reference_n1.json"),
             ("bench_n2.json", "r02_bench_n2.json"), ("bench_n8.json", "r02_bench_n8.json"), ("latency.txt", "r02_latency.txt"),
             ("quick_perf.txt", "r02_quick_perf.txt"), ("planar.txt", "r02_planar.txt"), ("sharp.txt", "r02_sharp.txt"),
             ("sharp_batch.txt", "r02_sharp_batch.txt"), ("batch_methods.txt", "r02_batch_methods.txt"),
             ("launches_bench.csv", "r02_launches_bench.csv"), ("sanitizer.txt", "r02_sanitizer.txt"),
             ("stripes_n4_m0.txt", "r02_stripes_n4_m0.txt"), ("stripes_n4_m4.txt", "r02_stripes_n4_m4.txt")):
    if os.path.exists(os.path.join(SRC, f)):
        shutil.copy(os.path.join(SRC, f), os.path.join(DST, t))
subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_opcodes.py")], capture_output=True)
print(sorted(x for x in os.listdir(DST) if x.startswith("r02") or x in ("f1_traffic.json", "sass_opcodes.txt")))
