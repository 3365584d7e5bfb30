"""Writes profiles/sass_opcodes.txt: per kernel of the built library, the number of SASS
instructions and of the opcodes that show which hardware paths it uses.  Runs here (no GPU)."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "sjpeg_b200", "libsjpeg_b200.so")],
                     capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
names = [f.split("\n", 1)[0].strip() for f in funcs[1:]]
dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
keys = ["UBLKCP", "SYNCS", "UCGABAR", "CGAERRBAR", "ATOMS", "ATOMG", "RED", "BAR", "SHFL", "VOTE", "PRMT", "IMAD",
        "LDL", "STL", "LDS", "STS", "LDG", "STG", "HMMA", "UTCMMA", "UTMALDG"]
out = ["SASS opcode evidence: cuobjdump -sass sjpeg_b200/libsjpeg_b200.so (sm_100a), per kernel the number of SASS instructions and",
       "of the opcodes that show which hardware paths it uses (regenerate: python tools/sass_opcodes.py).",
       "  UBLKCP = cp.async.bulk, the 1-D bulk copy of the TMA engine (row strips need no tensor map: no UTMALDG); SYNCS = mbarrier;",
       "  UCGABAR* / CGAERRBAR = cluster barriers (sharp conversion, DSMEM halos); ATOMS / ATOMG / RED = shared / global atomics;",
       "  BAR = CTA and named barriers; SHFL / VOTE = warp collectives; LDL / STL = local memory (spills).",
       "  No tensor-core opcode (HMMA, UTCMMA, ...) is expected or present: the path is integer / byte work (DESIGN.md section 4).",
       "", "%-40s %6s  %s" % ("kernel", "instr", " ".join(keys))]
for f, d in zip(funcs[1:], dem):
    ops = re.findall(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", f)
    c = collections.Counter()
    for o in ops:
        for k in keys:
            if o == k or o.startswith(k + "_"):
                c[k] += 1
    d = re.sub(r"\(.*$", "", d.replace("sjb::(anonymous namespace)::", "").replace("void ", ""))
    out.append("%-40s %6d  %s" % (d[:40], len(ops), " ".join(("%d" % c.get(k, 0)).rjust(len(k)) for k in keys)))
open(os.path.join(ROOT, "profiles", "sass_opcodes.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
