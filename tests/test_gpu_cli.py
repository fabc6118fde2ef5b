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


def test_load_csv_overlapped_and_raw_cache(tmp_path):
    """sgd_load_csv (parse of piece k+1 overlapped with the upload of piece k) gives the rows sgd_read_csv
    gives; the raw cache <csv>.f32 is written once, used while the CSV is unchanged, rebuilt afterwards."""
    import cuda_sgd_sese_project_b200 as pkg
    R, C = 300_000, 100       # ~330 MB of text: several 32 MB pieces, two pipeline stages
    rng = np.random.default_rng(9)
    X = rng.standard_normal((R, C)).astype(np.float32)
    y = rng.standard_normal(R).astype(np.float32)
    csv = tmp_path / "big.csv"
    pkg.write_csv(str(csv), X, y)
    ok, Xr, yr = pkg.read_csv(str(csv), R, C)
    assert ok and np.array_equal(Xr, X) and np.array_equal(yr, y)
    with pkg.SGDEngine(R, C, 1000, 0.1) as eng:
        st = eng.load_csv(str(csv))
        assert st["rows"] == R and st["from_cache"] == 0 and st["cache_written"] == 0 and st["parse_seconds"] > 0
        Xd, yd = eng.get_rows(0, R)
        assert np.array_equal(Xd, X) and np.array_equal(yd, y)
    with pkg.SGDEngine(R, C, 1000, 0.1) as eng:
        st1 = eng.load_csv(str(csv), cache=True)
        assert st1["from_cache"] == 0 and st1["cache_written"] == 1 and os.path.exists(str(csv) + ".f32")
        st2 = eng.load_csv(str(csv), cache=True)
        assert st2["from_cache"] == 1 and st2["parse_seconds"] == 0
        Xd, yd = eng.get_rows(0, R)
        assert np.array_equal(Xd, X) and np.array_equal(yd, y)
    # a shard of a two-rank layout loads its rows only
    with pkg.SGDEngine(R, C, 1000, 0.1, rank=1, nranks=2) as sh:
        st3 = sh.load_csv(str(csv), cache=True)
        assert st3["from_cache"] == 1 and st3["bytes_uploaded"] == (R // 2) * (C + 1) * 4
        Xs, ys = sh.get_rows(R // 2, R // 2)
        assert np.array_equal(Xs, X[R // 2:]) and np.array_equal(ys, y[R // 2:])
    # the CSV changes -> the cache is stale and is rebuilt
    X[0, 0] = 42.0
    pkg.write_csv(str(csv), X, y)
    with pkg.SGDEngine(R, C, 1000, 0.1) as eng:
        st4 = eng.load_csv(str(csv), cache=True)
        assert st4["from_cache"] == 0 and st4["cache_written"] == 1
        assert eng.get_rows(0, 1)[0][0, 0] == 42.0
    # errors: as sgd_read_csv
    with pkg.SGDEngine(10, 3, 5, 0.1) as eng:
        with pytest.raises(pkg.SgdError) as ei:
            eng.load_csv(str(tmp_path / "missing.csv"))
        assert ei.value.code == pkg._lib.SGD_ERR_IO
        bad = tmp_path / "bad.csv"
        bad.write_text("1,2,3,4,5\n")
        with pytest.raises(pkg.SgdError) as ei:
            eng.load_csv(str(bad))
        assert ei.value.code == pkg._lib.SGD_ERR_PARSE
        short = tmp_path / "short.csv"
        short.write_text("1,2,3,4\n5,6,7,8\n")            # fewer lines than announced: the rest is zero
        eng.load_csv(str(short))
        Xs, ys = eng.get_rows(0, 10)
        assert Xs[1].tolist() == [5.0, 6.0, 7.0] and ys[1] == 8.0 and not Xs[2:].any() and not ys[2:].any()


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
