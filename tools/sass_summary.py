#!/usr/bin/env python3
"""cuobjdump -sass of libvnect_b200.so -> profiles/<tag>_sass_summary.txt: per kernel, the counts of the Blackwell-native
mnemonics (UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UBLKCP = TMA, SYNCS = mbarrier) and of the
legacy tensor path (HMMA: must be zero), plus the instructions around the first MMA of every kernel.  Runs without a GPU."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "rXX"
lib = os.path.join(ROOT, "vnect_b200", "lib", "libvnect_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "SYNCS", "HMMA", "LDGSTS", "DFMA", "DADD", "DMUL"]
out = ["# cuobjdump -sass vnect_b200/lib/libvnect_b200.so (sm_100a): mnemonic counts per kernel", ""]
blocks = sass.split("Function : ")[1:]
for name, blk in zip(names, blocks):
    body = blk.split("\n")
    cnt = collections.Counter()
    tma_forms = collections.Counter()
    first_mma = None
    n_inst = 0
    for i, line in enumerate(body):
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        n_inst += 1
        op = m.group(1)
        for k in KEYS:
            if op == k or op.startswith(k + ".") or (k in ("UTCHMMA",) and op.startswith("UTCHMMA")):
                cnt[k] += 1
        if ".2CTA" in op and op.startswith("UTCHMMA"):
            cnt["UTCHMMA.2CTA"] += 1
        if op.startswith("UTMALDG") or op.startswith("UTMASTG"):
            tma_forms[op] += 1
        if first_mma is None and op.startswith("UTCHMMA"):
            first_mma = i
    short = name.replace("vnect::", "").replace("CUtensorMap_st, ", "")
    out.append(f"## {short[:150]}")
    out.append(f"   {n_inst} instructions; " + ", ".join(f"{k} {cnt[k]}" for k in KEYS if cnt[k]))
    if tma_forms:
        out.append("   TMA forms: " + ", ".join(f"{k} x{v}" for k, v in sorted(tma_forms.items())))
    if first_mma is not None:
        out.append("   around the first tcgen05.mma:")
        for line in body[max(0, first_mma - 6):first_mma + 10]:
            t = re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", line.rstrip())
            if re.search(r"/\*[0-9a-f]{4}\*/", t):
                out.append("      " + t.strip())
    out.append("")
open(os.path.join(ROOT, "profiles", f"{tag}_sass_summary.txt"), "w").write("\n".join(out))
print("kernels:", len(blocks))
