#!/usr/bin/env python
"""Regenerates the committed golden fixtures.  Run from the repo root IN THE BUILD CONTAINER
(needs /root/reference for the reference's own test data):

    python tests/golden/make_golden.py            # everything
    python tests/golden/make_golden.py --no-big   # skip the 2^24 / 2^26 shuffles (~10 s, 400 MB)

Outputs (all committed):
  tests/golden/thirdparty_kat.json   known answers of the REAL glibc rand(), libstdc++
                                     std::random_shuffle and Thrust minstd/uniform_real
                                     implementations of this image (gen_thirdparty_kat.cpp), i.e.
                                     of the third-party code that defines the reference's
                                     permutation (main.cu:90,207) and initial weights
                                     (main.cu:389-395).
  tests/golden/ref_fixtures.npz      the data of the reference's own test fixtures
                                     (c_code/test-data/*) parsed into arrays; the reference
                                     directory does not exist on the GPU box, tests read this.
The expected ANSWERS for those fixtures are the ones the reference's test code states
(testing.cu:54 "8.0 expected", the order vector at testing.cu:80-84, ...) and live in
tests/test_oracle.py beside the citation.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("REFERENCE_DIR", "/root/reference")


def build_gen(tmp: str) -> str:
    exe = os.path.join(tmp, "gen_kat")
    subprocess.run(
        ["g++", "-O2", "-std=c++17", "-I/usr/local/cuda/include",
         "-DTHRUST_DEVICE_SYSTEM=THRUST_DEVICE_SYSTEM_CPP", "-Wno-deprecated-declarations",
         os.path.join(HERE, "gen_thirdparty_kat.cpp"), "-o", exe],
        check=True)
    return exe


def run(exe: str, *args: str):
    out = subprocess.run([exe, *args], check=True, capture_output=True, text=True).stdout
    return json.loads(out)


def thirdparty(big: bool) -> dict:
    with tempfile.TemporaryDirectory() as tmp:
        exe = build_gen(tmp)
        kat = {
            "_provenance": {
                "generator": "tests/golden/gen_thirdparty_kat.cpp via tests/golden/make_golden.py",
                "glibc": subprocess.run(["ldd", "--version"], capture_output=True,
                                        text=True).stdout.splitlines()[0],
                "gxx": subprocess.run(["g++", "--version"], capture_output=True,
                                      text=True).stdout.splitlines()[0],
                "thrust": "CCCL headers of /usr/local/cuda (CUDA 12.9), host-only build",
                "note": "one fresh process per record, so rand() always starts at srand(1)",
            },
            "rand_first": run(exe, "rand", "64"),
            "rand_at": [run(exe, "rand_at", str(k)) for k in (343, 999999, 9999999)],
            "shuffle_full": {str(n): run(exe, "shuffle", str(n), "3", "full") for n in (1, 2, 5, 10, 33)},
            "shuffle_kat": [],
            "weights_hex": run(exe, "weights", "2048"),
        }
        sizes = [1000, 10000, 4000000] + ([16777216, 67108864] if big else [])
        for n in sizes:
            kat["shuffle_kat"].extend(run(exe, "shuffle", str(n), "2"))
    return kat


def _load_csv(path: str, C: int):
    rows = []
    with open(path) as f:
        for line in f:
            line = line.strip()
            if line:
                rows.append([float(t) for t in line.split(",")])
    a = np.asarray(rows, dtype=np.float64)
    assert a.shape[1] == C + 1, (path, a.shape)
    return a[:, :C].astype(np.float32), a[:, C].astype(np.float32)


def ref_fixtures() -> dict:
    td = os.path.join(REF, "c_code", "test-data")
    spec = {  # name -> (file, C)
        "gemv_test": ("gemv_test", 4),
        "permutation_test": ("permutation_test.csv", 4),
        "col_sum_test": ("col_sum_test.csv", 4),
        "matrix_scale_test": ("matrix_scale_test", 4),
        "example": ("example.csv", 3),
        "xy5": ("5xy.csv", 1),
        "correct_prediction": ("correct_prediction.csv", 3),
        "enb2012": ("ENB2012_data_Y1.csv", 8),
    }
    out = {}
    for name, (fn, C) in spec.items():
        X, y = _load_csv(os.path.join(td, fn), C)
        out[name + "_X"] = X
        out[name + "_y"] = y
    out["correct_prediction_true_w"] = np.loadtxt(
        os.path.join(td, "correct_prediction_weights.csv"), delimiter=",", dtype=np.float64)
    return out


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--no-big", action="store_true")
    ap.add_argument("--only", choices=["kat", "fixtures"], default=None)
    a = ap.parse_args()
    if a.only in (None, "kat"):
        kat = thirdparty(not a.no_big)
        with open(os.path.join(HERE, "thirdparty_kat.json"), "w") as f:
            json.dump(kat, f, indent=1)
            f.write("\n")
        print("wrote thirdparty_kat.json")
    if a.only in (None, "fixtures"):
        if not os.path.isdir(REF):
            print("no reference directory; ref_fixtures.npz left as is", file=sys.stderr)
        else:
            np.savez_compressed(os.path.join(HERE, "ref_fixtures.npz"), **ref_fixtures())
            print("wrote ref_fixtures.npz")
    return 0


if __name__ == "__main__":
    sys.exit(main())
