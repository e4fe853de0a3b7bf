#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of the in-tree library (the Blackwell-only mnemonics that prove
tcgen05 / TMA / TMEM code): cuobjdump -sass libsqe_b200.so -> profiles/<name>.json

    python scripts/sass_histogram.py profiles/r2b_sass_opcode_histogram.json
"""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "semantic-query-engine_b200", "lib", "libsqe_b200.so")
WATCH = ["UTCHMMA", "UTCIMMA", "UTMALDG", "UBLKCP", "LDTM", "UTCBAR", "IDP", "ACQBULK", "MUFU", "SYNCS", "STS", "LDS"]


def main(out):
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    res, cur = [], None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = {"kernel": m.group(1), "instructions": 0}
            res.append(cur)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            cur["instructions"] += 1
            op = m.group(1)
            for w in WATCH:
                if op.startswith(w):
                    cur[w] = cur.get(w, 0) + 1
    names = [r["kernel"] for r in res]
    dm = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    for r, d in zip(res, dm):
        r["kernel"] = d[:160]
    with open(out, "w") as f:
        json.dump(res, f, indent=0)
    for r in res:
        if any(k in r for k in ("UTCHMMA", "UTCIMMA")):
            print({k: v for k, v in r.items()})


if __name__ == "__main__":
    main(sys.argv[1])
