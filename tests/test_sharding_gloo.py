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


def test_balanced_shares():
    sys.path.insert(0, os.path.join(ROOT, "i-admm-lstm_b200"))
    from iadmm_b200.dist import balanced_shares
    assert balanced_shares(512, [1.0, 1.0]) == [256, 256]
    sh = balanced_shares(2048, [440., 465., 430., 470., 450., 455., 445., 460.])
    assert sum(sh) == 2048 and max(sh) - min(sh) <= 25
    assert sh[3] == max(sh) and sh[2] == min(sh)                       # faster GPUs take more instances
    t = [s_ / r for s_, r in zip(sh, [440., 465., 430., 470., 450., 455., 445., 460.])]
    assert max(t) / min(t) < 1.01                                      # finish together within 1 %
    assert balanced_shares(10, [0.0, 1.0]) == [5, 5]                   # no usable measurement: equal shards
    assert balanced_shares(7, [1.0, 100.0]) == [1, 6]                  # nobody is left without work
    assert balanced_shares(600, [1.0, 3.0], cap=320) == [280, 320]     # the cap moves work to the slower rank, nothing is dropped
    assert sorted(balanced_shares(2, [1.0, 1.0, 1.0, 1.0])) == [0, 0, 1, 1]   # fewer instances than ranks: zero-size shares
    import pytest
    with pytest.raises(ValueError):
        balanced_shares(700, [1.0, 3.0], cap=320)                      # impossible under the caps: loud, never silent


def _balance_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "i-admm-lstm_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from iadmm_b200.dist import balance_by_rate
    mine, shares = balance_by_rate(512, 400.0 if rank == 0 else 440.0)
    q.put((rank, mine, shares))
    dist.destroy_process_group()


def test_two_rank_balance_by_rate():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 200
    ps = [ctx.Process(target=_balance_worker, args=(r, 2, port, q)) for r in range(2)]
    for p_ in ps: p_.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p_ in ps: p_.join(timeout=60)
    assert res[0][2] == res[1][2] and sum(res[0][2]) == 512
    assert res[0][1] == res[0][2][0] and res[1][1] == res[1][2][1] and res[1][1] > res[0][1]


def test_two_rank_sharded_solve_matches_single():
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 400)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert out["x_equal"] and out["pri_equal"]


def _dp_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "i-admm-lstm_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from oracle import iadmm_oracle as orc
    from iadmm_b200.dist import shard_instances, allreduce_gradients

    class Params(torch.nn.Module):           # stands in for the LSTM parameter container
        def __init__(self, prm):
            super().__init__()
            for k, v in prm.items():
                setattr(self, k, torch.nn.Parameter(v.clone()))

    B, n, mi, me, h, TL = 4, 10, 3, 4, 8, 3
    qp = orc.qp_instances(B, n, mi, me, seed=23, dtype=torch.float64)
    prm = orc.lstm_parameters(h, TL, seed=23, dtype=torch.float64)

    def window(mod, data):
        Bq = data["Q"].shape[0]
        m = mi + me
        x = torch.zeros((Bq, n, 1), dtype=torch.float64); y = torch.zeros((Bq, m, 1), dtype=torch.float64)
        z = torch.zeros((Bq, m, 1), dtype=torch.float64); xv = torch.zeros((Bq, n + m, 1), dtype=torch.float64)
        H = torch.zeros((Bq, n + m, h), dtype=torch.float64); C = torch.zeros((Bq, n + m, h), dtype=torch.float64)
        p_ = dict(mod.named_parameters())
        loss = 0.0
        for t in range(TL):
            x, y, z, xv, H, C, _ = orc.lstm_step(p_, t, mi, me, x, y, z, xv, 6e-6, H, C, data["Q"], data["p"], data["A0"],
                                                 data["zl"], data["zu"], form="block")
            loss = loss + orc.primal_dual_residuals(x, y, z, data["Q"], data["p"], data["A0"])[2].mean() / TL
        loss.backward()

    mod = Params(prm)
    window(mod, shard_instances({k: qp[k] for k in ("Q", "p", "A0", "zl", "zu")}, rank, world))
    allreduce_gradients(mod)
    # unequal shards (3 + 1 instances): each rank's loss is the mean over ITS instances, so the ranks are weighted B_r / B
    cut = 3
    mine = {k: (qp[k][:cut] if rank == 0 else qp[k][cut:]) for k in ("Q", "p", "A0", "zl", "zu")}
    mod_u = Params(prm)
    window(mod_u, mine)
    allreduce_gradients(mod_u, local_batch=mine["Q"].shape[0])
    if rank == 0:
        full = Params(prm)
        window(full, qp)
        out["max_rel"] = max(float((a.grad - b.grad).norm() / (b.grad.norm() + 1e-300))
                             for a, b in zip(mod.parameters(), full.parameters()))
        out["max_rel_unequal"] = max(float((a.grad - b.grad).norm() / (b.grad.norm() + 1e-300))
                                     for a, b in zip(mod_u.parameters(), full.parameters()))
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_matches_single():
    """Training DP: per-rank window gradients averaged by one all-reduce equal the gradient of
    the same window on the concatenated batch (loss is a batch mean, main.py:347)."""
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29900 + (os.getpid() % 90)
    mp.spawn(_dp_worker, args=(2, port, out), nprocs=2, join=True)
    assert out["max_rel"] < 1e-10
    assert out["max_rel_unequal"] < 1e-10


def test_balanced_shares_properties():
    """Property test (hypothesis): the shares always sum to the total, nobody is left without work, and a faster GPU never
    gets fewer instances than a slower one (up to the one instance of rounding)."""
    sys.path.insert(0, os.path.join(ROOT, "i-admm-lstm_b200"))
    from hypothesis import given, settings, strategies as st
    from iadmm_b200.dist import balanced_shares

    @settings(max_examples=300, deadline=None)
    @given(st.lists(st.floats(min_value=50.0, max_value=1000.0), min_size=1, max_size=8), st.integers(min_value=8, max_value=4096))
    def check(rates, per_rank):
        total = per_rank * len(rates)
        sh = balanced_shares(total, rates)
        assert sum(sh) == total and min(sh) >= 1
        for i in range(len(rates)):
            for j in range(len(rates)):
                if rates[i] > rates[j]:
                    assert sh[i] >= sh[j] - 1
        # finishing times within one instance's worth of the ideal
        t = [s_ / r for s_, r in zip(sh, rates)]
        ideal = total / sum(rates)
        assert max(t) <= ideal + 1.5 / min(rates)

    check()
