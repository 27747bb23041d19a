"""Multi-GPU tests (need >= 2 visible GPUs; skipped otherwise -- run with `gpurun --gpus 2`):
  * one process driving two devices in turn (the per-device kernel opt-ins and SM-count caches);
  * the one collective of the path: the NCCL all-reduce of the flat LSTM gradient buffer per TBPTT window (SURVEY.md
    section 8e / a16): all-reduced gradient == single-GPU gradient on the concatenated batch (main.py:347 batch mean)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _need_two():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")


def test_one_process_two_devices():
    """cuda:0 then cuda:1 in the same process: every kernel family that opts into > 48 KB of dynamic shared memory
    (tensor-core gate kernels, resident kernel, backward NT-GEMM, LU solve) must launch on the second device too, and give
    bit-identical results (same arithmetic, same work decomposition)."""
    _need_two()
    import iadmm_b200 as ia
    from oracle import iadmm_oracle as orc

    def run(dev):
        out = []
        for (B, n, mi, me, h, K, streaming) in ((2, 64, 24, 24, 320, 3, True), (3, 40, 12, 12, 64, 4, False), (2, 48, 16, 16, 208, 3, True)):
            qp = {k: v.to(dev) for k, v in orc.qp_instances(B, n, mi, me, seed=5).items()}
            prm = orc.lstm_parameters(h, K, seed=5)
            model = ia.LSTM(None, 2, h, K, dev)
            with torch.no_grad():
                for k, v in prm.items():
                    getattr(model, k).copy_(v.to(dev))
            sc = ia.Scaling(n, mi + me, 10, dev)
            data = sc.scale_data(qp["Q"], qp["p"], qp["A0"], qp["zl"], qp["zu"])
            with torch.no_grad():
                r = model.solve(K, mi, me, *data, 6e-6, scaling=sc, streaming=streaming)
            out += [r.x.cpu(), r.y.cpu(), r.H.cpu(), r.pri.cpu()]
            # training window (tensor-core forward with saved gates + tcgen05 backward GEMMs)
            m = mi + me
            st = [torch.zeros(s, device=dev) for s in ((B, n, 1), (B, m, 1), (B, m, 1), (B, n + m, 1), (B, n + m, h), (B, n + m, h))]
            model.zero_grad()
            loss, _ = model.train_window(2, mi, me, *data, 6e-6, st)
            out += [loss.detach().cpu().reshape(1), model.U_i.grad.cpu(), model.rho.grad.cpu()]
        # Stage II LU (lu_solve_kernel opts into 200 KB)
        lu = ia.LU(dev)
        kw = dict(Q=qp["Q"], p=qp["p"], A0=qp["A0"], lb=None, ub=None, zl=qp["zl"], zu=qp["zu"])
        model.materialize_kkt = True
        with torch.no_grad():
            st = [torch.zeros(s, device=dev) for s in ((B, n, 1), (B, m, 1), (B, m, 1), (B, n + m, 1), (B, n + m, h), (B, n + m, h))]
            x, y, z, xv, H, C, A_tild, b_tild, rho_vec = model(0, mi, me, *st[:4], 6e-6, st[4], st[5], **kw)
            x, y, z, xv, A_tild, b_tild, l_, p_ = lu(rho_vec, x, y, z, xv, 6e-6, A_tild, None, None, **kw)
        out += [x.cpu(), z.cpu()]
        torch.cuda.synchronize(dev)
        return out

    a = run("cuda:0")
    b = run("cuda:1")
    a2 = run("cuda:0")
    for i, (u, v, w) in enumerate(zip(a, b, a2)):
        assert torch.isfinite(v).all(), i
        assert torch.equal(u, v), i
        assert torch.equal(u, w), i


def _nccl_worker(rank, world, port, out):
    for p_ in (ROOT, os.path.join(ROOT, "i-admm-lstm_b200"), os.path.join(ROOT, "tests")):
        if p_ not in sys.path:
            sys.path.insert(0, p_)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import iadmm_b200 as ia
    from iadmm_b200.dist import allreduce_gradients
    from oracle import iadmm_oracle as orc
    B, n, mi, me, h, TL = 6, 64, 20, 24, 64, 4
    qp = orc.qp_instances(B, n, mi, me, seed=29)
    prm = orc.lstm_parameters(h, TL, seed=29, scale=2.0)
    m = mi + me

    def window(rows):
        model = ia.LSTM(None, 2, h, TL, dev, gate_mode="tc_f16f8")
        with torch.no_grad():
            for k, v in prm.items():
                getattr(model, k).copy_(v.to(dev))
        data = [qp[k][rows].contiguous().to(dev) for k in ("Q", "p", "A0", "zl", "zu")]
        b = data[0].shape[0]
        st = [torch.zeros(s, device=dev) for s in ((b, n, 1), (b, m, 1), (b, m, 1), (b, n + m, 1), (b, n + m, h), (b, n + m, h))]
        loss, _ = model.train_window(TL, mi, me, *data, 6e-6, st, loss_scale=1.0 / TL)
        return model, float(loss), b

    res = {}
    # equal shards (3 + 3) with the plain average, unequal shards (4 + 2) weighted by the local batch
    for tag, cut in (("equal", 3), ("unequal", 4)):
        rows = slice(0, cut) if rank == 0 else slice(cut, B)
        model, loss, b = window(rows)
        allreduce_gradients(model, local_batch=b if tag == "unequal" else None)
        torch.cuda.synchronize()
        if rank == 0:
            full, loss_full, _ = window(slice(0, B))
            errs = {}
            for k in prm:
                g_, f_ = getattr(model, k).grad.double(), getattr(full, k).grad.double()
                if float(f_.abs().max()) > 0:
                    errs[k] = float((g_ - f_).norm() / f_.norm())
            res[tag] = max(errs.values())
        # every rank holds the same gradient after the collective
        chk = torch.cat([p_.grad.reshape(-1) for p_ in model.parameters()]).double().sum()
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        if rank == 0:
            res[tag + "_identical"] = bool(lo == hi)
    # the library's own collective (iadmm_allreduce_grads of the C ABI: ncclAllReduce on a raw communicator) against
    # torch.distributed's: the same two-rank sum, bit for bit
    from iadmm_b200.dist import nccl_comm
    rows = slice(0, 3) if rank == 0 else slice(3, B)
    m_a, _, b_a = window(rows)
    m_b, _, _ = window(rows)
    allreduce_gradients(m_a, local_batch=b_a)
    allreduce_gradients(m_b, local_batch=b_a, comm=nccl_comm())
    torch.cuda.synchronize()
    same = all(torch.equal(pa.grad, pb.grad) for pa, pb in zip(m_a.parameters(), m_b.parameters()))
    flag = torch.tensor([1.0 if same else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        res["abi_collective_identical"] = bool(flag.item() == 1.0)
        out.update(res)
    dist.barrier()
    dist.destroy_process_group()


def test_nccl_allreduced_gradient_equals_single_gpu_gradient():
    _need_two()
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29300 + (os.getpid() % 200)
    mp.spawn(_nccl_worker, args=(2, port, out), nprocs=2, join=True)
    print("NCCL DP gradient vs single GPU:", dict(out))
    assert out["equal_identical"] and out["unequal_identical"] and out["abi_collective_identical"]
    # fp32 kernels: the per-instance adjoints are bit-identical, the only difference is the order of the sum over instances
    assert out["equal"] < 2e-5 and out["unequal"] < 2e-5
