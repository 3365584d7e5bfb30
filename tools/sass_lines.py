"""Per-source-line view of an ncu capture without the GUI.
  python tools/sass_lines.py <report.ncu-rep> <kernel regex> [library.so] [top N]
ncu's CSV export of the source page is SASS only; this joins it -- by instruction offset inside
the kernel -- with the line table of the same kernel from `nvdisasm -g` of the cubin in the
library, and prints, per source line: warp instructions executed (and their share), average
active threads, stall samples and the dominant stall reasons.  Runs here (no GPU needed)."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_rows(rep, kernel):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    out = []
    hdr = None
    for r in rows:
        if r and r[0] == "Address":
            if hdr is not None:
                break                      # first matching launch only
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        out.append(dict(zip(hdr, r)))
    return out


def line_table(lib, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    table = {}
    for f in os.listdir(tmp):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        cur = None
        inside = False
        for ln in txt.splitlines():
            m = re.match(r"^(_Z\w+):$", ln)
            if m:
                inside = re.search(kernel, m.group(1)) is not None and not table
                cur = None
                continue
            if not inside:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m:
                table[int(m.group(1), 16)] = (cur, m.group(2).strip())
        if table:
            break
    return table


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    lib = sys.argv[3] if len(sys.argv) > 3 else "sjpeg_b200/libsjpeg_b200.so"
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
    rows = sass_rows(rep, kernel)
    table = line_table(lib, kernel)
    base = int(rows[0]["Address"], 16)
    per = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
    total_inst = total_samp = 0
    stall_cols = [c for c in rows[0] if c.startswith("stall_") and "Not Issued" not in c]
    mismatch = 0
    for r in rows:
        off = int(r["Address"], 16) - base
        line, text = table.get(off, (None, None))
        if text is None or text.split()[0].lstrip("@!P0123456789T ") .split(".")[0] not in r["Source"]:
            mismatch += 1
        inst = int(r["Instructions Executed"] or 0)
        thr = int(r["Thread Instructions Executed"] or 0)
        samp = int(r["# Samples"] or 0)
        p = per[line]
        p[0] += inst
        p[1] += thr
        p[2] += samp
        for c in stall_cols:
            v = int(r[c] or 0)
            if v:
                p[3][c[6:]] += v
        total_inst += inst
        total_samp += samp
    print("kernel %s: %d SASS instructions, %d warp instructions executed, %d stall samples (%d offsets unmatched)" % (
        kernel, len(rows), total_inst, total_samp, mismatch))
    src = {}
    for (line, p) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        f, n = line if line else ("?", 0)
        if f not in src:
            try:
                src[f] = open(os.path.join(os.path.dirname(os.path.abspath(lib)), "csrc", f)).read().splitlines()
            except OSError:
                src[f] = []
        text = src[f][n - 1].strip()[:70] if 0 < n <= len(src[f]) else ""
        stalls = ", ".join("%s %d" % kv for kv in p[3].most_common(3))
        print("%-14s %5d  inst %9d (%4.1f%%) thr/inst %4.1f  samples %6d (%4.1f%%)  [%s]  %s" % (
            f, n, p[0], 100.0 * p[0] / max(total_inst, 1), p[1] / max(p[0], 1), p[2], 100.0 * p[2] / max(total_samp, 1),
            stalls, text))


if __name__ == "__main__":
    main()
