#!/usr/bin/env python
"""Instruction histogram per kernel of libb200sgd.so (cuobjdump -sass): the mnemonics that show what the
hot path is made of -- TMA (UTMALDG / UBLKCP), async stores into distributed shared memory (STAS),
mbarrier traffic (SYNCS), cluster barriers (UCGABAR), packed fp32 FMAs (FFMA2), cp.async (LDGSTS) -- and
that no tensor-core instruction is present (deliberately: the path is a memory-bound GEMV pair).
    python tools/sass_histogram.py > profiles/r2_sass_histogram.md"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "cuda_sgd_sese_project_b200", "libb200sgd.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
KEYS = ["UTMALDG", "UBLKCP", "LDGSTS", "STAS", "SYNCS", "UCGABAR", "FFMA2", "FFMA", "LDS", "SHFL", "BAR", "LDG", "STG", "ATOM", "RED",
        "HMMA", "UTCMMA", "UTCHMMA", "QGMMA"]
kern = None
hist = collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern).replace("b200sgd::", "").replace("void ", "")
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1)
        hist[kern]["total"] += 1
        base = op.split(".")[0]
        for k in KEYS:
            if base == k or (k == "UCGABAR" and base.startswith("UCGABAR")):
                hist[kern][k] += 1
        if op.startswith("UTMALDG.2D.GATHER4"):
            hist[kern]["GATHER4"] += 1
print("# SASS instruction histogram of `libb200sgd.so` (sm_100a), round 2\n")
print("`cuobjdump -sass` of the built library, static counts per kernel (`tools/sass_histogram.py`).  `UTMALDG` = TMA tensor load "
      "(`GATHER4`: the `tile::gather4` form, four arbitrary rows per instruction), `UBLKCP` = TMA bulk copy (one row), `LDGSTS` = "
      "`cp.async` (permutation indices straight into shared memory), `STAS` = `st.async` into another CTA's shared memory "
      "(reduce-scatter / all-gather of the cluster step tail, partial sums of the cluster kernel), `SYNCS` = mbarrier operations, "
      "`UCGABAR` = cluster barrier, `FFMA2` = packed fp32x2 FMA.  No tensor-core instruction (`HMMA` / `UTC*MMA` / `QGMMA`) anywhere: "
      "the path is a memory-bound GEMV pair (~1 flop per byte).\n")
cols = ["total", "UTMALDG", "GATHER4", "UBLKCP", "LDGSTS", "STAS", "SYNCS", "UCGABAR", "FFMA2", "FFMA", "LDS", "SHFL", "BAR", "ATOM", "RED",
        "HMMA", "UTCMMA", "UTCHMMA", "QGMMA"]
print("| kernel | " + " | ".join(cols) + " |")
print("|---|" + "---|" * len(cols))
for k, h in hist.items():
    print(f"| `{k}` | " + " | ".join(str(h.get(c, 0)) for c in cols) + " |")
tot = collections.Counter()
for h in hist.values():
    tot.update(h)
print("| **all kernels** | " + " | ".join(str(tot.get(c, 0)) for c in cols) + " |")
