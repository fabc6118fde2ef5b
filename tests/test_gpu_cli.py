"""The drop-in CLI (cuda_sgd_sese_project_b200/main) end to end on a B200: the reference's seven
positionals in, the reference's stdout line and five-key JSON out (main.cu:303-571), results equal
to the oracle's; and the experiment runner with the reference's sweep interface."""
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from oracle import loader as orc
from tests.helpers import make_regression

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "cuda_sgd_sese_project_b200", "main")


def test_cli_binary_is_not_stale():
    # the binary is a build artefact (git-ignored): build() must have produced it after the sources
    src = os.path.join(ROOT, "cuda_sgd_sese_project_b200", "csrc", "main.cpp")
    lib = os.path.join(ROOT, "cuda_sgd_sese_project_b200", "libb200sgd.so")
    assert os.path.exists(EXE), "run __graft_entry__.build() first"
    assert os.path.getmtime(EXE) >= os.path.getmtime(src)
    assert os.path.getmtime(EXE) >= os.path.getmtime(lib) - 1.0   # linked against the library it ships with


def write_csv(path, X, y):
    with open(path, "w") as f:
        for row, lab in zip(X, y):
            f.write(",".join("%.9g" % v for v in row) + ",%.9g\n" % lab)


@pytest.mark.parametrize("cuda_run,suffix,label", [(0, "cublas", "CUBLAS execution"), (1, "cuda", "Plain CUDA execution")])
def test_cli_matches_oracle_and_writes_reference_json(tmp_path, cuda_run, suffix, label):
    R, C, B, lr, epochs = 5000, 20, 250, 0.05, 4
    X, y, _ = make_regression(R, C, seed=2, fp16_round=True)
    csv = tmp_path / "my_data.v1.csv"
    write_csv(csv, X, y)
    wfile = tmp_path / "w.txt"
    r = subprocess.run([EXE, str(lr), str(epochs), str(csv), str(R), str(C), str(B), str(cuda_run),
                        f"--dump-weights={wfile}"], capture_output=True, text=True, cwd=tmp_path, timeout=300)
    assert r.returncode == 0, r.stderr
    assert label in r.stdout and "Data transfer time = " in r.stdout and "GPU time = " in r.stdout
    mae_cli = float(re.search(r"Mean absolute error : (\S+)", r.stdout).group(1))
    js = tmp_path / f"my_data.v1-{suffix}.json"       # get_filename_from_path strips the last suffix only
    vals = json.loads(js.read_text())
    assert sorted(vals) == ["error", "gpu_time", "shuffle_time", "total_gradient_time", "transfer_time"]
    mae_o, w_o, _ = orc.run(X, y, B, lr, epochs)
    assert abs(vals["error"] - mae_o) <= 1e-4 * mae_o + 1e-6 * float(np.mean(np.abs(y)))
    assert abs(mae_cli - vals["error"]) <= 1e-5 * abs(vals["error"]) + 1e-9   # stdout prints 6 digits
    w = np.loadtxt(wfile)
    assert np.linalg.norm(w - w_o) / np.linalg.norm(w_o) <= 1e-4
    assert vals["gpu_time"] > 0 and vals["total_gradient_time"] <= vals["gpu_time"] * 1.01   # true, not re-added


def test_cli_full_batch_docstring_example(tmp_path, fixtures):
    # main.cu:311: ./main 0.00001 10 data/5xy.csv 40 1 0 0
    X, y = fixtures["xy5_X"], fixtures["xy5_y"]
    csv = tmp_path / "5xy.csv"
    write_csv(csv, X, y)
    r = subprocess.run([EXE, "0.00001", "10", str(csv), "40", "1", "0", "0"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0
    mae = json.loads((tmp_path / "5xy-cublas.json").read_text())["error"]
    assert mae == pytest.approx(198.1447, rel=1e-4)


def test_experiment_runner_feature_sweep(tmp_path):
    data = str(tmp_path / "data") + "/"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_experiment.py"), "--generate", "-i", data,
                        "-p", "test_number_of_features-", "--start", "8", "--end", "24", "--stride", "8",
                        "-r", "3000", "-b", "500", "-l", "0.05", "-e", "3", "--paramName", "features"],
                       capture_output=True, text=True, cwd=tmp_path, timeout=600)
    assert r.returncode == 0, r.stderr
    recs = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    assert [x["value"] for x in recs] == [8, 16, 24]
    for x in recs:
        assert os.path.exists(tmp_path / x["json"]) and x["gpu_time"] > 0 and np.isfinite(x["error"])
