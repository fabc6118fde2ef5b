#!/usr/bin/env python
"""Data ingestion at BASELINE config 2's real size (4M x 100 fp32 -> a ~4 GB CSV), SURVEY 8(f)-2:
the reference's read_csv (getline / stringstream / stof, sgd_io.cu:38-77 -- timed through the unmodified
binary oracle/_ref/main, whose "csv read time" line reports exactly that call) against sgd_read_csv
(threaded from_chars), sgd_load_csv (parse overlapped with the upload) and the raw cache.
    python tools/ingest_bench.py [rows] [cols]      -> one JSON object on stdout"""
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import cuda_sgd_sese_project_b200 as pkg

R = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
C = int(sys.argv[2]) if len(sys.argv) > 2 else 100
out = {"rows": R, "cols": C, "values": R * (C + 1), "host_cores": os.cpu_count()}
with tempfile.TemporaryDirectory(dir=os.environ.get("TMPDIR", "/tmp")) as tmp:
    rng = np.random.default_rng(42)
    X = rng.standard_normal((R, C), dtype=np.float32)
    y = rng.standard_normal(R, dtype=np.float32)
    csv = os.path.join(tmp, "c2.csv")
    t0 = time.perf_counter(); pkg.write_csv(csv, X, y); out["write_csv_s"] = time.perf_counter() - t0
    out["csv_bytes"] = os.path.getsize(csv)
    t0 = time.perf_counter(); ok, Xr, yr = pkg.read_csv(csv, R, C); dt = time.perf_counter() - t0
    assert ok and np.array_equal(Xr, X) and np.array_equal(yr, y)
    out["sgd_read_csv"] = {"seconds": dt, "values_per_s": out["values"] / dt}
    del Xr, yr
    with pkg.SGDEngine(R, C, 1000, 0.1) as eng:
        t0 = time.perf_counter(); st = eng.load_csv(csv); dt = time.perf_counter() - t0
        out["sgd_load_csv_overlapped"] = {"seconds": dt, "values_per_s": out["values"] / dt, "parse_seconds": st["parse_seconds"]}
        t0 = time.perf_counter(); st = eng.load_csv(csv, cache=True); dt = time.perf_counter() - t0
        out["sgd_load_csv_writing_cache"] = {"seconds": dt, "values_per_s": out["values"] / dt, "cache_written": st["cache_written"]}
        t0 = time.perf_counter(); st = eng.load_csv(csv, cache=True); dt = time.perf_counter() - t0
        out["sgd_load_csv_from_cache"] = {"seconds": dt, "values_per_s": out["values"] / dt, "from_cache": st["from_cache"]}
        Xd, yd = eng.get_rows(0, 1000)
        assert np.array_equal(Xd, X[:1000]) and np.array_equal(yd, y[:1000])
    # for comparison: create from host arrays (the pre-existing path: H2D of the parsed arrays)
    t0 = time.perf_counter()
    with pkg.SGDEngine(R, C, 1000, 0.1, data=X, labels=y) as eng:
        out["sgd_create_from_host_arrays_s"] = time.perf_counter() - t0
    ref = os.path.join(ROOT, "oracle", "_ref", "main")
    if os.path.exists(ref):
        r = subprocess.run([ref, "0.1", "1", csv, str(R), str(C), "1000", "1"], capture_output=True, text=True, cwd=tmp, timeout=1500)
        m = re.search(r"csv read time = (\S+) ms", r.stdout)
        if m:
            s = float(m.group(1)) * 1e-3
            out["reference_read_csv"] = {"seconds": s, "values_per_s": out["values"] / s}
        else:
            out["reference_read_csv"] = {"error": (r.stdout + r.stderr)[-300:]}
print(json.dumps(out))
