#!/usr/bin/env python
"""Turns the ncu captures brought back in gpurun_out/ into the tracked summaries under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/launches_bench_c2.csv profiles/r1_launches_bench_c2.md
    python tools/summarize_ncu.py full gpurun_out/prof_r1_c2_epoch.ncu-rep profiles/r1_epoch_c2.md "<title>" [alg_bytes]
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
    agg = OrderedDict()
    for r in rows:
        name = r[4].split("(")[0]
        ns = float(r[-1])
        a = agg.setdefault(name, [0, 0.0, r[8], r[7]])
        a[0] += 1
        a[1] += ns
    total = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list: `python bench.py --steps 2 --warmup 3 --no-cpu` (round 1)\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv`, one B200. Per-launch times "
                "are cold-cache and serialised; read the SHARES. The run has 1 warm-up call of 3 epochs and 2 timed "
                "epochs plus the e2e leg (2 x 2 epochs from host buffers) and the MAE passes.\n\n")
        f.write("| kernel | launches | total ms | share | grid | block |\n|---|---|---|---|---|---|\n")
        for name, (n, ns, grid, block) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{name}` | {n} | {ns / 1e6:.3f} | {100 * ns / total:.1f} % | {grid} | {block} |\n")
        f.write(f"\nTotal {total / 1e6:.2f} ms over {len(rows)} launches. `sgd_epoch_kernel` is the dominant kernel "
                "(the roofline object of bench.py); the seven `shuf_*` / `scan_*` kernels are the bit-exact device "
                "shuffle of one epoch.\n")


KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__cycles_active.avg",
]


def full(src, dst, title, alg_bytes=None, kregex="sgd_epoch_kernel"):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
    srcp = subprocess.run(["ncu", "-i", src, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(srcp)))
    sh = srows[1]
    ix = {h: i for i, h in enumerate(sh)}
    data = srows[2:]
    tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
    stalls = [h for h in sh if h.startswith("stall_") and "Not Issued" not in h]
    st_tot = {h: sum(int(r[ix[h]] or 0) for r in data) for h in stalls}
    ublk = sum(int(r[ix["Instructions Executed"]] or 0) for r in data if "UBLKCP" in r[ix["Source"]])
    utma = sum(int(r[ix["Instructions Executed"]] or 0) for r in data if "UTMALDG" in r[ix["Source"]])
    with open(dst, "w") as f:
        f.write(f"# {title}\n\n`ncu --set full --clock-control none --import-source on -k regex:{kregex} -s 1 -c 1` "
                f"on one B200 (`{src.split('/')[-1]}`, not committed: binary). One launch = one epoch.\n\n")
        f.write("| metric | value |\n|---|---|\n")
        f.write(f"| kernel | `{d['Kernel Name'][0] if 'Kernel Name' in d else 'sgd_epoch_kernel'}` |\n")
        for k in KEYS:
            if k in d:
                f.write(f"| `{k}` | {d[k][0]} {d[k][1]} |\n")
        rd = float(d["dram__bytes_read.sum"][0]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[d["dram__bytes_read.sum"][1]]
        wr = float(d["dram__bytes_write.sum"][0]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[d["dram__bytes_write.sum"][1]]
        f.write(f"| DRAM traffic (read + write) per launch | {(rd + wr) / 1e6:.1f} MB |\n")
        if alg_bytes:
            f.write(f"| algorithmic bytes per launch (DESIGN.md 3.1) | {float(alg_bytes) / 1e6:.1f} MB |\n")
            f.write(f"| traffic / algorithmic | {(rd + wr) / float(alg_bytes):.3f} |\n")
        f.write(f"| `UBLKCP` (TMA bulk copy, one row each) instructions executed | {ublk} |\n")
        f.write(f"| `UTMALDG.2D.GATHER4` (TMA tile::gather4, four rows each) instructions executed | {utma} |\n")
        f.write(f"\nWarp stall samples ({tot} total), share by reason:\n\n| reason | share |\n|---|---|\n")
        ssum = sum(st_tot.values()) or 1
        for h, v in sorted(st_tot.items(), key=lambda kv: -kv[1])[:8]:
            f.write(f"| {h} | {100 * v / ssum:.1f} % |\n")
        f.write("\nHottest SASS instructions (samples, executions):\n\n```\n")
        for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:14]:
            f.write(f"{int(r[ix['# Samples']]):8d} {int(r[ix['Instructions Executed']] or 0):10d}  {r[ix['Source']][:90]}\n")
        f.write("```\n")
    return rd + wr


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        t = full(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5] if len(sys.argv) > 5 else None,
                 sys.argv[6] if len(sys.argv) > 6 else "sgd_epoch_kernel")
        print(int(t))
