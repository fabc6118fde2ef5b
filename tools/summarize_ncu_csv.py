#!/usr/bin/env python
"""Round 2: the ncu reports stay on the GPU box (19 MB each with sources); tools/exp_ncu.sh brings their
`--page raw --csv` and `--page source --csv` exports back.  This turns them into profiles/*.md.
    python tools/summarize_ncu_csv.py full gpurun_out/r2_X_raw.csv gpurun_out/r2_X_source.csv.gz profiles/r2_X.md "<title>" [alg_bytes]
    python tools/summarize_ncu_csv.py launches gpurun_out/r2_launches_Y.csv profiles/r2_launches_Y.md "<title>"
"""
import csv
import gzip
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_dim_x", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg",
]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def full(raw, src, dst, title, alg_bytes=None):
    rows = list(csv.reader(open(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
    out = [f"# {title}\n", f"`ncu --set full --clock-control none --import-source on -s 1 -c 1` on one B200 (the report stays on the "
           f"box; this is its `--page raw` / `--page source` export, `{raw.split('/')[-1]}`).  One launch = one epoch"
           " (the kernel is launched without the cooperative attribute under ncu, `B200SGD_NO_COOP=1`: ncu cannot replay a launch "
           "that is cooperative AND clustered).\n", "| metric | value |", "|---|---|", f"| kernel | `{d['Kernel Name'][0]}` |"]
    for k in KEYS:
        if k in d:
            out.append(f"| `{k}` | {d[k][0]} {d[k][1]} |")
    tot = None
    try:
        rd = float(d["dram__bytes_read.sum"][0]) * UNIT[d["dram__bytes_read.sum"][1]]
        wr = float(d["dram__bytes_write.sum"][0]) * UNIT[d["dram__bytes_write.sum"][1]]
        tot = rd + wr
        out.append(f"| DRAM traffic (read + write) per launch | {tot / 1e6:.1f} MB |")
        if alg_bytes:
            out.append(f"| algorithmic bytes per launch (DESIGN.md 3.1) | {alg_bytes / 1e6:.1f} MB |")
            out.append(f"| traffic / algorithmic | {tot / alg_bytes:.3f} |")
            t = float(d["gpu__time_duration.sum"][0]) * {"ms": 1e-3, "us": 1e-6, "s": 1.0, "ns": 1e-9}[d["gpu__time_duration.sum"][1]]
            out.append(f"| algorithmic GB/s under ncu (cold, serialised) | {alg_bytes / t * 1e-9:.0f} |")
    except Exception:
        pass
    # stall reasons and executed TMA / DSMEM instructions from the source page
    s = list(csv.reader(gzip.open(src, "rt")))
    sh = s[1]
    ix = {h: i for i, h in enumerate(sh)}
    data = [r for r in s[2:] if len(r) == len(sh)]
    stalls = [h for h in sh if h.startswith("stall_") and "Not Issued" not in h and "(Not" not in h]
    nsamp = sum(int(r[ix["# Samples"]] or 0) for r in data)
    st = {h: sum(int(r[ix[h]] or 0) for r in data) for h in stalls}
    ops = OrderedDict((k, 0) for k in ("UBLKCP", "UTMALDG", "LDGSTS", "STAS", "SYNCS", "LDS", "SHFL", "FFMA", "BAR"))
    for r in data:
        srcl = r[ix["Source"]].strip()
        op = srcl.split()[0] if srcl else ""
        if op.startswith("@") and len(srcl.split()) > 1:
            op = srcl.split()[1]
        base = op.split(".")[0]
        if base in ops:
            ops[base] += int(r[ix["Instructions Executed"]] or 0)
    out.append("")
    out.append("Executed warp instructions (source page): " + ", ".join(f"`{k}` {v}" for k, v in ops.items()) + ".")
    out.append(f"\nWarp stall samples ({nsamp} total), share by reason:\n")
    out.append("| reason | share |")
    out.append("|---|---|")
    tot_st = sum(st.values()) or 1
    for h, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]:
        out.append(f"| {h} | {100 * v / tot_st:.1f} % |")
    open(dst, "w").write("\n".join(out) + "\n")
    return tot


def launches(src, dst, title):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
    hdr = None
    for r in csv.reader(open(src)):
        if r and r[0] == "ID":
            hdr = r
    mi, vi, ni = hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Kernel Name")
    gi, bi = hdr.index("Grid Size"), hdr.index("Block Size")
    agg = OrderedDict()
    for r in rows:
        name = r[ni].split("(")[0]
        a = agg.setdefault(name, {"n": 0, "ns": 0.0, "rd": 0.0, "wr": 0.0, "grid": r[gi], "block": r[bi]})
        v = float(r[vi].replace(",", ""))
        if r[mi] == "gpu__time_duration.sum":
            a["n"] += 1
            a["ns"] += v
        elif r[mi] == "dram__bytes_read.sum":
            a["rd"] += v
        elif r[mi] == "dram__bytes_write.sum":
            a["wr"] += v
    total = sum(a["ns"] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# {title}\n\n`ncu --metrics gpu__time_duration.sum[,dram__bytes_*] --clock-control none --csv`, one B200. Per-launch "
                "times are cold-cache and serialised (no overlap between streams): read the SHARES.\n\n")
        f.write("| kernel | launches | total ms | share | ms / launch | DRAM read+write MB / launch | grid | block |\n|---|---|---|---|---|---|---|---|\n")
        for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
            mb = (a["rd"] + a["wr"]) / max(1, a["n"]) / 1e6
            f.write(f"| `{name}` | {a['n']} | {a['ns'] / 1e6:.3f} | {100 * a['ns'] / total:.1f} % | {a['ns'] / 1e6 / max(1, a['n']):.3f} | "
                    f"{mb:.1f} | {a['grid']} | {a['block']} |\n")
        f.write(f"\nTotal {total / 1e6:.2f} ms over {sum(a['n'] for a in agg.values())} launches.\n")


if __name__ == "__main__":
    if sys.argv[1] == "full":
        t = full(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5], float(sys.argv[6]) if len(sys.argv) > 6 else None)
        print(t)
    else:
        launches(sys.argv[2], sys.argv[3], sys.argv[4])
