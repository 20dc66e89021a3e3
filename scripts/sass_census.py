"""SASS opcode census of the shipped library: which hardware paths each kernel really uses (cuobjdump -sass, no GPU needed).

    python scripts/sass_census.py [lib.so] > profiles/<tag>_sass_census.json

Per kernel: instruction count and the counts of the mnemonics that prove a path — DMMA (FP64 tensor core), UTCHMMA / UTCQMMA
(tcgen05.mma), LDTM / STTM (tensor memory), UTCBAR (tcgen05.commit), UBLKCP (bulk copy engine), UTMALDG (tensor-map TMA),
SYNCS (mbarrier), DFMA / DADD / DMUL (FP64 pipe), FFMA, LDS / STS, LDG / STG, BAR."""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "lentil_b200", "liblentil_b200.so")
WATCH = ["DMMA", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "SYNCS", "DFMA", "DADD", "DMUL", "FFMA", "FADD",
         "FMUL", "LDS", "STS", "LDG", "STG", "BAR", "LDL", "STL", "CCTL", "RED", "ATOMG"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kernels, cur = {}, None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = kernels.setdefault(name.split("(")[0].replace("void ", ""), collections.Counter())
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur is not None:
        cur["instructions"] += 1
        op = m.group(1)
        for w in WATCH:
            if op == w or op.startswith(w + "."):
                cur[w] += 1
res = {"library": os.path.relpath(lib, ROOT), "kernels": {k: dict(v) for k, v in sorted(kernels.items())}}
tot = collections.Counter()
for v in kernels.values():
    tot.update(v)
res["total"] = dict(tot)
json.dump(res, sys.stdout, indent=1)
