"""world_size-2 gloo test of the N>1 path's host logic: instance sharding with no data-path collective.
Each rank solves its contiguous chunk (with the CPU oracle standing in for the device kernels) and the
gathered result must equal the single-process solve bit for bit."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "i-admm-lstm_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from oracle import iadmm_oracle as orc
    from iadmm_b200.dist import shard_instances, gather_batch, shard_range
    B, n, mi, me, h, K = 5, 12, 4, 4, 8, 6
    qp = orc.qp_instances(B, n, mi, me, seed=21)
    prm = orc.lstm_parameters(h, K, seed=21)
    loc = shard_instances({k: qp[k] for k in ("Q", "p", "A0", "zl", "zu")}, rank, world)
    lo, hi = shard_range(B, rank, world)
    assert loc["Q"].shape[0] == hi - lo
    r = orc.solve(prm, K, mi, me, loc["Q"], loc["p"], loc["A0"], loc["zl"], loc["zu"], 6e-6, h)
    x = gather_batch(r.x, B)
    pri = gather_batch(r.pri, B, dim=1)
    if rank == 0:
        full = orc.solve(prm, K, mi, me, qp["Q"], qp["p"], qp["A0"], qp["zl"], qp["zu"], 6e-6, h)
        out["x_equal"] = bool(torch.equal(x, full.x))
        out["pri_equal"] = bool(torch.equal(pri, full.pri))
    dist.destroy_process_group()


def test_shard_ranges():
    sys.path.insert(0, os.path.join(ROOT, "i-admm-lstm_b200"))
    from iadmm_b200.dist import shard_range
    for B in (1, 5, 8, 256, 257):
        for world in (1, 2, 4, 8):
            spans = [shard_range(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_sharded_solve_matches_single():
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 400)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert out["x_equal"] and out["pri_equal"]
