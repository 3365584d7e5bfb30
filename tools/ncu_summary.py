"""Turns an .ncu-rep (ncu --set full) into the JSON summary kept under profiles/.
  python tools/ncu_summary.py gpurun_out/x.ncu-rep "command that produced it" > profiles/rNN_x.json
Runs here (ncu is installed in the build container; no GPU needed to read a report)."""
import csv
import io
import json
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]


def main():
    rep, command = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    if rep.endswith(".csv"):          # the raw page exported on the GPU box (tools/collect_profiles.sh)
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    kernels = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        k = {"kernel": r[ix["Kernel Name"]]}
        for m in KEEP:
            if m in ix:
                k[m] = "%s %s" % (r[ix[m]], units[ix[m]])
        for h, i in ix.items():
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v >= 0.05:
                    k["stall_" + h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = round(v, 3)
        kernels.append(k)
    json.dump({"command": command, "kernels": kernels}, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
