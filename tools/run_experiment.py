#!/usr/bin/env python
"""Parameter sweeps over the drop-in CLI with the reference's experiment contract.

Mirrors the runner half of /root/reference/experiments/plot_parametrized.py (:60-125; the plotting
half needs matplotlib, which this image does not have): same options, same argv for the binary
(`cmd lr epochs csv R C B cudaRun`, plot_parametrized.py:100-105), same JSON file naming and
renaming for batch sweeps (:108-121), same three fields read back (:123-125).  Instead of PDFs it
prints one JSON line per parameter value and a summary table.

Two additions: `--generate` writes the data files first, in the distribution of
data_set_creation/generate_scalable_data.py:30-32 (make_regression-like, fp32, seed 42), because the
reference's `/data` directory is git-ignored; `--cmd` defaults to this repository's `main`.

    python tools/run_experiment.py --generate -i /tmp/data/ -p test_number_of_features- \
        --start 200 --end 1000 --stride 200 -r 10000 -b 1000 -l 0.1 -e 20 --paramName features
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def generate_csv(path: str, R: int, C: int, seed: int = 42) -> None:
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((R, C)).astype(np.float32)
    w = (100.0 * rng.random(C)).astype(np.float32)
    y = (X @ w).astype(np.float32)
    with open(path, "w") as f:
        for row, lab in zip(X, y):
            f.write(",".join("%.9g" % v for v in row) + ",%.9g\n" % lab)


def main() -> int:
    ap = argparse.ArgumentParser(description="Runs the SGD binary over a range of one parameter "
                                             "(interface of experiments/plot_parametrized.py)")
    ap.add_argument("-i", "--input", required=True, help="directory of the data files, with trailing '/'")
    ap.add_argument("-p", "--prefix", required=True, help="file-name prefix (features/samples) or file name (batch)")
    ap.add_argument("--start", type=int, required=True)
    ap.add_argument("--end", type=int, required=True)
    ap.add_argument("--stride", type=int, required=True)
    ap.add_argument("--cuda", type=int, default=0, help="1 = 'plain CUDA' selector, 0 = 'cuBLAS' selector")
    ap.add_argument("-c", "--columns", type=int, default=-1)
    ap.add_argument("-r", "--rows", type=int, default=-1)
    ap.add_argument("-b", "--batchsize", type=int, required=True)
    ap.add_argument("-l", "--learningrate", type=float, default=0.01)
    ap.add_argument("-e", "--epochs", type=int, default=10)
    ap.add_argument("--cmd", default=os.path.join(ROOT, "cuda_sgd_sese_project_b200", "main"))
    ap.add_argument("--paramName", required=True, choices=["features", "samples", "batch"])
    ap.add_argument("--generate", action="store_true", help="write missing data files first")
    a = ap.parse_args()

    if a.columns == -1 and a.paramName == "samples":
        print("Error: Need to provide --columns argument when --paramName is set to 'samples'")
        return 1
    if a.rows == -1 and a.paramName == "features":
        print("Error: Need to provide --rows argument when --paramName is set to 'features'")
        return 1
    if (a.rows == -1 or a.columns == -1) and a.paramName == "batch":
        print("Error: Need to provide both --rows and --columns arguments when --paramName is set to 'batch'")
        return 1

    suffix = "cublas" if a.cuda == 0 else "cuda"
    full_path = a.input + a.prefix
    rows_out = []
    for parameter in range(a.start, a.end + a.stride, a.stride):
        R = parameter if a.paramName == "samples" else a.rows
        C = parameter if a.paramName == "features" else a.columns
        B = parameter if a.paramName == "batch" else a.batchsize
        dataset = full_path if a.paramName == "batch" else full_path + str(parameter) + ".csv"
        if a.generate and not os.path.exists(dataset):
            os.makedirs(os.path.dirname(dataset) or ".", exist_ok=True)
            generate_csv(dataset, R, C)
        run_line = [a.cmd] + [str(v) for v in (a.learningrate, a.epochs, dataset, R, C, B, a.cuda)]
        p = subprocess.run(run_line, capture_output=True, text=True)
        if p.returncode != 0:
            print(p.stdout[-2000:], p.stderr[-2000:], file=sys.stderr)
            return p.returncode
        if a.paramName != "batch":
            json_filename = a.prefix + str(parameter) + "-{}.json".format(suffix)
        else:  # the binary would overwrite it: rename, as plot_parametrized.py:112-118 does
            json_filename = a.prefix.split(".")[0] + "-{}.json".format(suffix)
            parts = json_filename.split("-")
            parts[1] = str(parameter)
            new_filename = "-".join(parts)
            os.rename(json_filename, new_filename)
            json_filename = new_filename
        with open(json_filename) as f:
            vals = json.load(f)
        rec = {"parameter": a.paramName, "value": parameter, "R": R, "C": C, "B": B, "json": json_filename,
               "transfer_time": vals["transfer_time"], "gpu_time": vals["gpu_time"], "error": vals["error"]}
        rows_out.append(rec)
        print(json.dumps(rec), flush=True)
    print("\n%-10s %14s %14s %14s" % (a.paramName, "transfer ms", "gpu_time ms", "error"))
    for r in rows_out:
        print("%-10d %14.3f %14.3f %14.6g" % (r["value"], r["transfer_time"], r["gpu_time"], r["error"]))
    return 0


if __name__ == "__main__":
    sys.exit(main())
