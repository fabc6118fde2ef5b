"""The N > 1 path on the CPU: world_size-2 gloo processes run the host statement of the row-sharded
data-parallel step (identical global permutation on every rank, per-step bucketing by owner,
rank-ordered gradient sum) and must reproduce the single-rank oracle.  Also exercises the peer-handle
rendezvous of cuda_sgd_sese_project_b200.dist with a stand-in engine.  No GPU."""
import os
import socket
import sys

import numpy as np
import pytest

import cuda_sgd_sese_project_b200 as pkg
from oracle import loader as orc
from tests.helpers import make_regression, rel_l2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_partition_the_rows():
    for R, N in [(10, 3), (4000000, 8), (7, 8), (67108864, 8), (1000, 1)]:
        edges = [pkg.shard_range(R, k, N) for k in range(N)]
        assert edges[0][0] == 0 and edges[-1][1] == R
        for a, b in zip(edges, edges[1:]):
            assert a[1] == b[0]
        per = -(-R // N)
        assert all(e - b <= per for b, e in edges)


def test_host_bucket_keeps_batch_order_and_covers_every_row_once():
    R, B, N = 10007, 250, 3
    perm = np.arange(R, dtype=np.int32)
    pkg.host_shuffle(perm)
    _, nb = pkg.batch_geometry(R, B)
    seen = []
    for k in range(N):
        b, e = pkg.shard_range(R, k, N)
        local, off = pkg.host_bucket(perm, B, k, N)
        assert off[0] == 0 and off[-1] == local.shape[0] and np.all(np.diff(off) >= 0)
        for s in range(nb):
            batch = perm[s * B:(s + 1) * B]
            mine = batch[(batch >= b) & (batch < e)] - b          # batch order preserved
            assert np.array_equal(local[off[s]:off[s + 1]], mine)
        seen.append(local + b)
    allrows = np.sort(np.concatenate(seen))
    assert np.array_equal(allrows, np.sort(perm[: nb * B]))    # tail rows dropped (main.cu:344)


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
    dist.init_process_group("gloo", rank=rank, world_size=world)
    X, y, _ = make_regression(R, C, seed=7)
    b, e = pkg.shard_range(R, rank, world)
    Xl, yl = X[b:e].astype(np.float32), y[b:e].astype(np.float32)     # this rank's shard only
    w = pkg.init_weights(C)
    perm = np.arange(R, dtype=np.int32)
    _, nb = pkg.batch_geometry(R, B)

    class FakeEngine:   # the rendezvous only needs these four members
        nranks = world
        imported = {}

        def peer_export(self):
            return bytes([rank]) * 128

        def peer_import(self, r, h):
            self.imported[r] = h

        def peer_ready(self):
            assert sorted(self.imported) == list(range(world))

    fe = FakeEngine()
    connect_peers(fe)
    assert all(fe.imported[r] == bytes([r]) * 128 for r in range(world))

    for epoch in range(1, epochs + 1):
        pkg.host_shuffle(perm, skip_epochs=epoch - 1)         # every rank: the same global permutation
        local, off = pkg.host_bucket(perm, B, rank, world)
        a = np.float32(-(float(np.float32(lr)) / (epoch ** 0.25)))
        for s in range(nb):
            rows = local[off[s]:off[s + 1]]
            r = Xl[rows] @ w - yl[rows]
            part = torch.from_numpy((Xl[rows].T @ r).astype(np.float32))
            parts = [torch.zeros_like(part) for _ in range(world)]
            dist.all_gather(parts, part)
            g = parts[0].numpy().copy()
            for k in range(1, world):                          # rank order, as the kernel does
                g += parts[k].numpy()
            w = (a * (g / np.float32(B)) + w).astype(np.float32)
    np.save(out, w)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_world_size_2_gloo_matches_single_rank_oracle(tmp_path, world):
    import torch.multiprocessing as mp
    R, C, B, lr, epochs = 3000, 24, 200, 0.05, 2
    port = _free_port()
    outs = [str(tmp_path / f"w{r}.npy") for r in range(world)]
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, world, port, R, C, B, lr, epochs, outs[r])) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    ws = [np.load(o) for o in outs]
    assert np.array_equal(ws[0], ws[1])                      # replicas hold identical bits
    X, y, _ = make_regression(R, C, seed=7)
    _, w_o, _ = orc.run(X, y, B, lr, epochs)
    assert rel_l2(ws[0], w_o) < 1e-5
    # and the oracle's own sharded emulation agrees
    g = orc.Rand()
    ind = np.arange(R, dtype=np.int32)
    w_s = orc.init_weights(C)
    for e in range(1, epochs + 1):
        g.shuffle(ind)
        orc.epoch_sharded(X, y, B, ind, w_s, lr, e, world)
    assert rel_l2(ws[0], w_s) < 1e-5
