"""Row-sharded data parallel on real GPUs (needs >= 2 B200s; skipped on a 1-GPU box): every rank is
one process, the gradient exchange runs inside the step kernel over NVLink peer memory."""
import os
import socket
import sys

import numpy as np
import pytest

import cuda_sgd_sese_project_b200 as pkg
from oracle import loader as orc
from tests.helpers import make_regression, rel_l2

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, R, C, B, lr, epochs, out):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import cuda_sgd_sese_project_b200 as pkg
    from cuda_sgd_sese_project_b200.dist import connect_peers
    from tests.helpers import make_regression
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    X, y, _ = make_regression(R, C, seed=11)
    eng = pkg.SGDEngine(R, C, B, lr, device=rank, rank=rank, nranks=world)
    b, e = pkg.shard_range(R, rank, world)
    assert (eng.local_row_begin, eng.local_rows) == (b, e - b)
    eng.load_rows(b, X[b:e], y[b:e])
    connect_peers(eng)
    for epoch in range(1, epochs + 1):
        eng.iteration(epoch)
        perm = eng.permutation()
        local, off = eng.local_schedule()
        hl, hoff = pkg.host_bucket(perm, B, rank, world)
        assert np.array_equal(off, hoff) and np.array_equal(local, hl), "device bucketing differs from the host statement"
    w = eng.weights
    s = torch.tensor([eng.abs_error_sum()], dtype=torch.float64, device="cuda")
    dist.all_reduce(s)
    np.save(out, np.concatenate([w, [np.float32(float(s[0]) / R)]]))
    dist.barrier()
    eng.close()
    dist.destroy_process_group()


def _run(world, R, C, B, lr, epochs, tmp_path):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    port = _free_port()
    outs = [str(tmp_path / f"w{r}.npy") for r in range(world)]
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, world, port, R, C, B, lr, epochs, outs[r])) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    res = [np.load(o) for o in outs]
    for r in res[1:]:
        assert np.array_equal(r, res[0])        # bit-identical replicas
    X, y, _ = make_regression(R, C, seed=11)
    mae_o, w_o, _ = orc.run(X, y, B, lr, epochs)
    assert rel_l2(res[0][:-1], w_o) <= 1e-4
    assert abs(float(res[0][-1]) - mae_o) <= 1e-4 * mae_o + 1e-6 * float(np.mean(np.abs(y)))


@pytest.mark.parametrize("R,C,B", [(40000, 100, 1000), (50000, 512, 4096), (20000, 2048, 512), (33333, 7, 100)])
def test_two_ranks_match_single_rank_oracle(tmp_path, R, C, B):
    _run(2, R, C, B, 0.05, 2, tmp_path)


def test_two_ranks_generic_kernel_direct_push(tmp_path, monkeypatch):
    monkeypatch.setenv("B200SGD_CLUSTER", "0")   # narrow rows on sgd_epoch_kernel: partials pushed to the peers' owners
    _run(2, 40000, 100, 1000, 0.05, 2, tmp_path)


def test_four_ranks(tmp_path):
    _run(4, 60000, 100, 2000, 0.05, 2, tmp_path)


def test_eight_ranks(tmp_path):
    _run(8, 80000, 64, 4000, 0.05, 2, tmp_path)


def _worker_setperm(rank, world, port, R, C, B, out):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import cuda_sgd_sese_project_b200 as pkg
    from cuda_sgd_sese_project_b200.dist import connect_peers
    from oracle import loader as orc
    from tests.helpers import make_regression
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    X, y, _ = make_regression(R, C, seed=11)
    eng = pkg.SGDEngine(R, C, B, 0.01, device=rank, rank=rank, nranks=world)
    b, e = pkg.shard_range(R, rank, world)
    eng.load_rows(b, X[b:e], y[b:e])
    connect_peers(eng)
    g = orc.Rand()
    ind = np.arange(R, dtype=np.int32)
    eng.run(1, 2)                           # shuffle 1 (own, from sgd_create's look-ahead), 2 (shared); 3 is ahead
    g.shuffle(ind); g.shuffle(ind)
    assert np.array_equal(eng.permutation(), ind)
    start = ind[::-1].copy()
    eng.set_permutation(start)              # the map of shuffle 3 is applied to the new permutation instead
    eng.iteration(3)
    g.shuffle(start)
    assert np.array_equal(eng.permutation(), start)
    eng.run(4, 2)
    g.shuffle(start); g.shuffle(start)
    assert np.array_equal(eng.permutation(), start)
    np.save(out, eng.weights)
    dist.barrier()
    eng.close()
    dist.destroy_process_group()


def test_set_permutation_after_a_shared_lookahead(tmp_path):
    import torch
    import torch.multiprocessing as mp
    world = 2
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    port = _free_port()
    outs = [str(tmp_path / f"w{r}.npy") for r in range(world)]
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker_setperm, args=(r, world, port, 60000, 64, 2000, outs[r])) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert np.array_equal(np.load(outs[0]), np.load(outs[1]))
