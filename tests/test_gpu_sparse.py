"""Sparse problem families: the bitmap-slab form of Q / A0 (iadmm_sparse_pack,
iadmm_solve_sparse) against the dense path on the densified problem -- which is what the reference computes, since
main.py:243-296 calls .toarray() on every family -- and the on-disk instance format through pinned staging onto the GPU."""
import os
import sys

import numpy as np
import pytest
import torch

from helpers import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _block_words(M):
    """Host-side block occupancy (documents the layout: bit s of word g <-> rows 8g..8g+7, columns 128s..128s+127)."""
    B, rows, n = M.shape
    G, S = (rows + 7) // 8, (n + 127) // 128
    out = np.zeros((B, G), dtype=np.uint64)
    nz = M.cpu().numpy() != 0
    for b in range(B):
        for g in range(G):
            for s_ in range(S):
                if nz[b, 8 * g:8 * g + 8, 128 * s_:128 * s_ + 128].any():
                    out[b, g] |= np.uint64(1) << np.uint64(s_)
    return out


def _unpack(sb):
    """Host-side decode of a SparseBatch (documents the layout of include/iadmm.h: mask | off | vals)."""
    B, rows, n = sb.shape
    S = (n + 127) // 128
    raw = sb.buf.cpu().numpy()
    al = lambda v: (v + 255) // 256 * 256
    mb, ob = al(B * rows * S * 16), al(B * rows * S * 4)
    mask = raw[:B * rows * S * 16].view(np.uint32).reshape(B, rows, S, 4)
    off = raw[mb:mb + B * rows * S * 4].view(np.uint32).reshape(B, rows, S)
    vals = raw[mb + ob:mb + ob + B * sb.cap * 4].view(np.float32).reshape(B, sb.cap)
    out = np.zeros((B, rows, n), dtype=np.float32)
    for b in range(B):
        for r in range(rows):
            for s in range(S):
                k = int(off[b, r, s])
                for w in range(4):
                    word = int(mask[b, r, s, w])
                    for bit in range(32):
                        if word >> bit & 1:
                            out[b, r, s * 128 + w * 32 + bit] = vals[b, k]
                            k += 1
    return out


@pytest.mark.parametrize("shape", [(2, 37, 130), (3, 20, 256), (1, 9, 1000), (2, 0, 16)])
def test_pack_roundtrip(shape):
    import iadmm_b200 as ia
    B, rows, n = shape
    g = torch.Generator().manual_seed(rows + n)
    M = torch.randn((B, rows, n), generator=g) * (torch.rand((B, rows, n), generator=g) < 0.37)
    if rows:
        M[0, 0, :] = 0.0                 # an empty row
        M[-1, -1, :] = 1.5               # a full row
        M[0, rows // 2, 0] = -0.0        # a negative zero is a zero
    sb = ia.SparseBatch.pack(M.to(DEV))
    if rows == 0:
        assert sb is None
        return
    assert torch.equal(sb.nnz.cpu().long(), torch.count_nonzero(M.reshape(B, -1), dim=1))
    assert np.array_equal(_unpack(sb), M.numpy() + 0.0)
    assert sb.bytes_per_instance < 4 * rows * n * 1.05


@pytest.mark.parametrize("shape", [(2, 37, 130), (1, 64, 1000), (3, 100, 260)])
def test_block_mask(shape):
    import iadmm_b200 as ia
    B, rows, n = shape
    g = torch.Generator().manual_seed(rows)
    M = torch.zeros((B, rows, n))
    for b in range(B):                       # a diagonal, a dense band of rows and a few scattered entries
        for i in range(min(rows, n)):
            M[b, i, i] = 1.0 + i
        M[b, rows // 2:rows // 2 + 3, :] = torch.randn((min(3, rows - rows // 2), n), generator=g)
        M[b, 0, n - 1] = -0.0                # a negative zero does not make a block non-empty
    sb = ia.SparseBatch.blocks(M.to(DEV))
    G = (rows + 7) // 8
    words = sb.buf.cpu().numpy()[:B * G * 8].view(np.uint64).reshape(B, G)
    ref = _block_words(M)
    assert np.array_equal(words, ref)
    assert np.array_equal(sb.nonempty.cpu().numpy(), np.array([sum(bin(int(w)).count("1") for w in ref[b]) for b in range(B)]))
    assert 0 < sb.occupancy < 1 and sb.bytes_per_instance <= 4 * rows * n + 8 * G


def _families():
    from iadmm_b200 import data
    yield "Random_QP", data.generate_family_batch("Random_QP", 3, 100, num_ineq=50, seed=1, device=DEV)
    yield "Equality_QP", data.generate_family_batch("Equality_QP", 2, 132, num_eq=60, seed=2, device=DEV)
    yield "SVM", data.generate_family_batch("SVM", 2, 60, num_ineq=40, seed=3, device=DEV)
    yield "QPLIB-like 1 %", data.generate_family_batch("Random_QP", 2, 1100, num_ineq=300, seed=4, device=DEV, density=0.01)
    yield "ragged n", data.generate_family_batch("Random_QP", 2, 203, num_ineq=77, seed=5, device=DEV)
    from bench import device_qp_batch
    Q, p, A0, zl, zu = device_qp_batch(2, 300, 100, 60, 6, DEV)                     # the QP family: diagonal Q, dense A0
    yield "QP (diagonal Q)", dict(Q=Q, p=p, A0=A0, zl=zl, zu=zu, num_var=300, num_ineq=100, num_eq=60)


@pytest.mark.parametrize("mode", ["simt_fp32", "tc_f16f8"])
def test_sparse_solve_is_bit_identical_to_the_densified_problem(mode):
    """Same lane-to-column assignment, same accumulation order: skipping the zeros must not change a single bit of the
    iterates, the residual traces or the metrics -- with Ruiz scaling (which keeps the sparsity pattern) in front."""
    import iadmm_b200 as ia
    from oracle import iadmm_oracle as orc
    for name, d in _families():
        n, mi, me = d["num_var"], d["num_ineq"], d["num_eq"]
        h, K = 32, 6
        prm = orc.lstm_parameters(h, K, seed=7, scale=3.0)
        model = ia.LSTM(None, 2, h, K, DEV, gate_mode=mode)
        with torch.no_grad():
            for k, v in prm.items():
                getattr(model, k).copy_(v.to(DEV))
        sc = ia.Scaling(n, mi + me, 10, DEV)
        data = sc.scale_data(d["Q"], d["p"], d["A0"], d["zl"], d["zu"])
        with torch.no_grad():
            dense = model.solve(K, mi, me, *data, 6e-6, scaling=sc, streaming=True)
            sparse = model.solve(K, mi, me, *data, 6e-6, scaling=sc, sparse=True)
            blocks = model.solve(K, mi, me, *data, 6e-6, scaling=sc, sparse="blocks")
            mixed = model.solve(K, mi, me, *data, 6e-6, scaling=sc, sparse=(ia.SparseBatch.blocks(data[0]), ia.SparseBatch.pack(data[2])))
            auto = model.solve(K, mi, me, *data, 6e-6, scaling=sc, sparse="auto")
        q_sp, a_sp = model.last_sparse
        print(name, mode, "density Q %.3f A0 %.3f" % (ia.SparseBatch.pack(data[0]).density, ia.SparseBatch.pack(data[2]).density),
              "block occupancy Q %.2f A0 %.2f" % (ia.SparseBatch.blocks(data[0]).occupancy, ia.SparseBatch.blocks(data[2]).occupancy),
              "auto:", q_sp.kind if q_sp else "dense", a_sp.kind if a_sp else "dense")
        for r in (sparse, blocks, mixed, auto):
            for k in ("x", "y", "z", "xv", "H", "C", "pri", "dual", "pri_unscaled", "dual_unscaled", "metrics"):
                assert torch.equal(getattr(dense, k), getattr(r, k)), (name, mode, k)
        assert torch.isfinite(dense.x).all()
        # and the dense path is the reference's arithmetic: oracle on the densified problem
        if name == "Random_QP":
            ref = orc.solve(prm, K, mi, me, *(t_.cpu() for t_ in data), 6e-6, h, form="block")
            for k in ("x", "z", "pri", "dual"):
                assert rel_err(getattr(sparse, k), getattr(ref, k)) < 5e-5, (name, k)


def test_sparse_only_pointers_and_errors():
    """A matrix given in sparse form needs no dense copy (NULL dense pointer); giving neither form is an error."""
    import iadmm_b200 as ia
    from iadmm_b200 import data
    from oracle import iadmm_oracle as orc
    d = data.generate_family_batch("SVM", 2, 40, num_ineq=24, seed=9, device=DEV)
    n, mi = d["num_var"], d["num_ineq"]
    h, K = 16, 4
    model = ia.LSTM(None, 2, h, K, DEV)
    with torch.no_grad():
        ref = model.solve(K, mi, 0, d["Q"], d["p"], d["A0"], d["zl"], d["zu"], 6e-6, streaming=True)
        q_sp, a_sp = ia.SparseBatch.pack(d["Q"]), ia.SparseBatch.pack(d["A0"])
        B, m = 2, mi
        st = [torch.zeros(s, device=DEV) for s in ((B, n, 1), (B, m, 1), (B, m, 1), (B, n + m, 1), (B, n + m, h), (B, n + m, h))]
        ws = model._workspace(B, n, m, model._mode(), torch.device(DEV))
        torch.ops.iadmm.solve_sparse(model.packed_weights(), None, q_sp.buf, q_sp.cap, None, d["p"], None, a_sp.buf, a_sp.cap, None,
                                     d["zl"], d["zu"], None, None, None, *st, None, None, None, None, None, ws, mi, 0, h, K, 0, K, 6e-6,
                                     model._mode(), 1)
        for a, b in zip(st, (ref.x, ref.y, ref.z, ref.xv, ref.H, ref.C)):
            assert torch.equal(a, b)
        with pytest.raises(ia.IadmmError, match="neither matrix"):
            torch.ops.iadmm.solve_sparse(model.packed_weights(), d["Q"], None, 0, None, d["p"], d["A0"], None, 0, None, d["zl"], d["zu"],
                                         None, None, None, *st, None, None, None, None, None, ws, mi, 0, h, K, 0, K, 6e-6, model._mode(), 0)


def test_dataset_files_to_gpu_solve(tmp_path):
    """f3 on the GPU: instances written the way generate_data.py writes them (scipy csc matrices for Random_QP, Q0 = Q/2),
    read back through the pinned staging buffers of data.load_batch onto the device, scaled and solved; identical to
    solving the tensors they were written from."""
    import scipy.sparse as sp
    import iadmm_b200 as ia
    from iadmm_b200 import data
    d = data.generate_family_batch("Random_QP", 3, 48, num_ineq=20, seed=11, device=DEV)
    root = data.dataset_dir(str(tmp_path), "Random_QP", 48, 20)
    for i in range(3):
        inst = dict(Q=sp.csc_matrix((d["Q"][i] / 2).cpu().numpy()), p=sp.csc_matrix(d["p"][i].cpu().numpy()),
                    G=sp.csc_matrix(torch.cat((d["A0"][i], -d["A0"][i])).cpu().numpy()),
                    c=sp.csc_matrix(torch.cat((d["zu"][i], -d["zl"][i])).cpu().numpy()), A0=sp.csc_matrix(d["A0"][i].cpu().numpy()),
                    zl=sp.csc_matrix(d["zl"][i].cpu().numpy()), zu=sp.csc_matrix(d["zu"][i].cpu().numpy()))
        data.write_instance(data.instance_path(root, "Random_QP", i), inst)
    batch, sizes = data.load_batch(root, "Random_QP", [0, 1, 2], DEV)
    assert batch["Q"].is_cuda and sizes["num_var"] == 48
    for k in ("Q", "p", "A0", "zl", "zu"):
        assert torch.equal(batch[k], d[k]), k
    model = ia.LSTM(None, 2, 16, 5, DEV)
    sc = ia.Scaling(48, 20, 10, DEV)
    with torch.no_grad():
        a = model.solve(5, 20, 0, *sc.scale_data(*(batch[k] for k in ("Q", "p", "A0", "zl", "zu"))), 6e-6, scaling=sc, sparse="auto")
        b = model.solve(5, 20, 0, *sc.scale_data(*(d[k] for k in ("Q", "p", "A0", "zl", "zu"))), 6e-6, scaling=sc)
    assert torch.equal(a.x, b.x) and torch.equal(a.pri_unscaled, b.pri_unscaled)
