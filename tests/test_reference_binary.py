"""Three-way parity on the GPU box: the UNMODIFIED reference (oracle/_ref/main_weights, built by
oracle/build_ref.sh from /root/reference/c_code) vs the CPU oracle vs libb200sgd -- same CSV, same
hyper-parameters, both of the reference's code paths (cudaRun 0 = cuBLAS, 1 = plain CUDA)."""
import os
import re
import subprocess

import numpy as np
import pytest

import cuda_sgd_sese_project_b200 as pkg
from oracle import loader as orc
from tests.helpers import make_regression, rel_l2

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "main_weights")


def write_csv(path, X, y):
    with open(path, "w") as f:
        for row, lab in zip(X, y):
            f.write(",".join("%.9g" % v for v in row) + "," + "%.9g" % lab + "\n")


def run_reference(csv, R, C, B, lr, epochs, cuda_run, cwd):
    r = subprocess.run([REF, repr(lr), str(epochs), csv, str(R), str(C), str(B), str(cuda_run)],
                       capture_output=True, text=True, cwd=cwd, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    mae = float(re.search(r"Mean absolute error : (\S+)", r.stdout).group(1))
    w = np.array([float(t) for t in re.search(r"weights = \[ (.*?)\]", r.stdout).group(1).split()], dtype=np.float64)
    return mae, w, r.stdout


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/main_weights not built")
@pytest.mark.parametrize("R,C,B,lr,epochs", [(2000, 20, 100, 0.05, 5), (10000, 100, 1000, 0.1, 20), (999, 7, 0, 0.01, 3)])
@pytest.mark.parametrize("cuda_run", [0, 1])
def test_three_way_parity(tmp_path, R, C, B, lr, epochs, cuda_run):
    X, y, _ = make_regression(R, C, seed=R + C, fp16_round=True)
    csv = str(tmp_path / "data.csv")
    write_csv(csv, X, y)
    mae_ref, w_ref, out = run_reference(csv, R, C, B, lr, epochs, cuda_run, str(tmp_path))
    assert w_ref.shape == (C,)
    # the reference's JSON exists under the reference's name (main.cu:545 / :557)
    assert os.path.exists(tmp_path / ("data-cuda.json" if cuda_run else "data-cublas.json"))

    mae_o, w_o, _ = orc.run(X, y, B, lr, epochs)
    with pkg.SGDEngine(R, C, B, lr, data=X, labels=y, cudaRun=cuda_run) as eng:
        eng.run(1, epochs)
        w = eng.weights
        mae = eng.calculate_mean_abs_error()
    # reference binary <-> oracle (pins the oracle), and ours <-> reference binary (the real gate)
    assert rel_l2(w_o, w_ref) <= 1e-4, ("oracle vs reference", rel_l2(w_o, w_ref))
    assert rel_l2(w, w_ref) <= 1e-4, ("b200sgd vs reference", rel_l2(w, w_ref))
    assert abs(mae_o - mae_ref) <= 1e-4 * mae_ref + 1e-6
    assert abs(mae - mae_ref) <= 1e-4 * mae_ref + 1e-6
