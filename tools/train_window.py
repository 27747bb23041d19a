"""BASELINE config 3: truncated-BPTT training windows on n=1000 QPs (forward + backward kernels + NCCL gradient
all-reduce + the reference's Adam), run like main.py:336-358 through the drop-in modules.

    python tools/train_window.py [--batch 2] [--tl 100] [--windows 3]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/train_window.py ...
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "i-admm-lstm_b200"))
import torch
import torch.distributed as dist
import iadmm_b200 as ia
from iadmm_b200.dist import allreduce_gradients
from bench import device_qp_batch

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=2); ap.add_argument("--tl", type=int, default=100)
ap.add_argument("--windows", type=int, default=3); ap.add_argument("--n", type=int, default=1000)
ap.add_argument("--hidden", type=int, default=800); ap.add_argument("--lr", type=float, default=5e-5)
ap.add_argument("--gate-mode", default="tc_f16f8"); ap.add_argument("--fused", type=int, default=1)
a = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr_ = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr_); dev = torch.device("cuda", lr_)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n, mi, me, h, TL, B = a.n, a.n // 2, a.n // 2, a.hidden, a.tl, a.batch
torch.manual_seed(17)
model = ia.LSTM(None, 2, h, TL, dev, gate_mode=a.gate_mode)
opt = torch.optim.Adam(model.parameters(), lr=a.lr, weight_decay=0.0)          # main.py:191
Q, p, A0, zl, zu = device_qp_batch(B, n, mi, me, 17 + rank, dev)
sc = ia.Scaling(n, mi + me, 10, dev)
Q, p, A0, zl, zu = sc.scale_data(Q, p, A0, zl, zu)
m = mi + me
x = torch.zeros((B, n, 1), device=dev); y = torch.zeros((B, m, 1), device=dev); z = torch.zeros((B, m, 1), device=dev)
xv = torch.zeros((B, n + m, 1), device=dev); H = torch.zeros((B, n + m, h), device=dev); C = torch.zeros((B, n + m, h), device=dev)
times, losses = [], []
for wdw in range(a.windows):
    if os.environ.get("PROFILE_WINDOW") == str(wdw): torch.cuda.profiler.start()
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    t0 = time.perf_counter()
    opt.zero_grad()
    if a.fused:                                      # one library call per window (iadmm_train_window)
        loss, (x, y, z, xv, H, C) = model.train_window(TL, mi, me, Q, p, A0, zl, zu, 6e-6, (x, y, z, xv, H, C), loss_scale=1.0 / TL)
    else:                                            # the reference's loop, one autograd node per iteration
        loss = 0.0
        for t in range(TL):
            x, y, z, xv, H, C, _, _, _ = model(t, mi, me, x, y, z, xv, 6e-6, H, C, Q=Q, p=p, A0=A0, lb=None, ub=None, zl=zl, zu=zu)
            pr, du, tot = ia.primal_dual_loss(x, y, z, Q, p, A0)
            loss = loss + tot.mean() / TL
        loss.backward()
    if world > 1:
        allreduce_gradients(model)                   # ONE all-reduce of the flat gradient buffer per window
    opt.step()
    x, y, z, xv, H, C = (v.detach() for v in (x, y, z, xv, H, C))
    torch.cuda.synchronize()
    times.append(time.perf_counter() - t0); losses.append(float(loss))
    if os.environ.get("PROFILE_WINDOW") == str(wdw): torch.cuda.profiler.stop()
chk = torch.stack([prm.detach().double().sum() for prm in model.parameters()]).sum()
if world > 1:
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    same = bool(lo == hi)
else:
    same = True
if rank == 0:
    print(json.dumps({"workload": f"TBPTT window: n={n}, {mi}+{me}, h={h}, TL={TL}, batch {B}/GPU, fwd+bwd+allreduce+Adam, " +
                                  ("one fused call per window" if a.fused else "per-iteration autograd"),
                      "n_gpus": world, "s_per_window": times, "instances_per_s": world * B / min(times[1:] or times),
                      "loss": losses, "weights_identical_across_ranks": same,
                      "mem_GB": torch.cuda.max_memory_allocated() / 1e9}))
if world > 1:
    dist.destroy_process_group()
