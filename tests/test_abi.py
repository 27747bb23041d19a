"""CPU checks of the C-ABI boundary: the library builds, loads without a GPU, exports every symbol
include/iadmm.h declares, sizes its workspaces, and fails loudly (no fallback) when no B200 is present."""
import ctypes
import os
import re
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def libpath():
    sys.path.insert(0, os.path.join(ROOT, "i-admm-lstm_b200"))
    import build as iadmm_build
    return iadmm_build.build()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "iadmm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(iadmm_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported(libpath):
    L = ctypes.CDLL(libpath)
    names = declared_symbols()
    assert len(names) >= 12
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/iadmm.h but not exported"


def test_binding_covers_header(libpath):
    from iadmm_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    assert _lib.lib().iadmm_abi_version() == 3


def test_workspace_sizes(libpath):
    from iadmm_b200 import _lib
    L = _lib.lib()
    n = ctypes.c_size_t()
    assert L.iadmm_weights_bytes(800, 100, ctypes.byref(n)) == 0
    assert n.value >= 4 * 800 * 800 * 4 + 2 * 4 * 800 * 800 * 2
    assert L.iadmm_solve_workspace_bytes(256, 1000, 1000, 800, 0, ctypes.byref(n)) == 0
    simt = n.value
    assert L.iadmm_solve_workspace_bytes(256, 1000, 1000, 800, 1, ctypes.byref(n)) == 0
    assert n.value > simt > 256 * 2000 * 800 * 4          # holds the second H buffer
    assert L.iadmm_solve_workspace_bytes(4, 100, 100, 60, 1, ctypes.byref(n)) == -6   # h % 8 != 0 -> EMODE
    assert b"hidden_dim" in L.iadmm_last_error()
    assert L.iadmm_solve_workspace_bytes(0, 100, 100, 64, 0, ctypes.byref(n)) == -1   # ESHAPE
    assert L.iadmm_ruiz_workspace_bytes(8, 100, 100, ctypes.byref(n)) == 0 and n.value > 0
    assert L.iadmm_residuals_workspace_bytes(8, 100, 100, ctypes.byref(n)) == 0 and n.value > 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(libpath):
    """Without a B200 the product path raises; it never routes through the oracle or torch."""
    import iadmm_b200 as ia
    assert ia.lib().iadmm_device_check() != 0
    model = ia.LSTM(None, 2, 8, 4, "cpu")
    qp = {k: torch.zeros(s) for k, s in dict(Q=(1, 4, 4), p=(1, 4, 1), A0=(1, 2, 4), zl=(1, 2, 1), zu=(1, 2, 1)).items()}
    with pytest.raises(ia.IadmmError):
        with torch.no_grad():
            model.solve(2, 1, 1, qp["Q"], qp["p"], qp["A0"], qp["zl"], qp["zu"], 1e-6)
    with pytest.raises(ia.IadmmError):
        ia.Scaling(4, 2, 10, "cpu").scale_data(qp["Q"], qp["p"], qp["A0"], qp["zl"], qp["zu"])
    with pytest.raises(ia.IadmmError):
        ia.primal_dual_loss(torch.zeros(1, 4, 1), torch.zeros(1, 2, 1), torch.zeros(1, 2, 1), qp["Q"], qp["p"], qp["A0"])


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "i-admm-lstm_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.replace("no CPU or PyTorch fallback", ""), f"{f} mentions the oracle"


def test_state_dict_contract():
    """Parameter names, order and shapes match the reference (checkpoints load unchanged, lstm.py:21-41)."""
    import numpy as np
    import iadmm_b200 as ia
    g = np.load(os.path.join(ROOT, "tests", "golden", "init_contract.npz"))
    model = ia.LSTM(None, 2, 8, 5, "cpu")
    sd = model.state_dict()
    assert list(sd.keys()) == list(g["keys"])
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(g["shape_" + k])
    assert model.name() == "lstm"


def test_header_is_plain_c_and_links(libpath, tmp_path):
    """include/iadmm.h is the drop-in boundary: it must compile as plain C99 (no C++-isms, no torch/CUDA types) and a C
    program must link against the shared library and call a non-compute entry point."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    src = tmp_path / "c_abi.c"
    src.write_text('#include "iadmm.h"\n#include <stdio.h>\n'
                   'int main(void) { size_t b = 0; int rc = iadmm_weights_bytes(64, 10, &b);\n'
                   '  printf("%d %d %lu\\n", iadmm_abi_version(), rc, (unsigned long)b);\n'
                   '  return (iadmm_abi_version() == IADMM_ABI_VERSION && rc == IADMM_OK && b > 0) ? 0 : 1; }\n')
    exe = tmp_path / "c_abi"
    inc = os.path.join(ROOT, "include")
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", inc, str(src), "-o", str(exe),
                           libpath, "-Wl,-rpath," + os.path.dirname(libpath)])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr


def test_release_library_reads_no_environment(libpath):
    """VERDICT r1 / ADVICE: no environment variable may change what the shipped library computes.  Our sources call
    getenv in exactly one place (dev_env, under IADMM_DEV_SWITCHES, development build only) and none of the release
    object files references it (the one undefined `getenv` of the .so comes from the statically linked cudart)."""
    import subprocess
    csrc = os.path.join(ROOT, "i-admm-lstm_b200", "csrc")
    hits = []
    for f in sorted(os.listdir(csrc)):
        txt = re.sub(r"//[^\n]*", "", open(os.path.join(csrc, f)).read())
        hits += [f for _ in re.findall(r"\bgetenv\s*\(", txt)]
    assert hits == ["api.cu"], hits
    api = open(os.path.join(csrc, "api.cu")).read()
    assert re.search(r"#ifdef IADMM_DEV_SWITCHES\s+return getenv\(name\);", api)
    objdir = os.path.join(ROOT, "i-admm-lstm_b200", "build")
    for f in sorted(os.listdir(objdir)):
        if f.endswith(".o"):
            syms = subprocess.run(["nm", os.path.join(objdir, f)], capture_output=True, text=True).stdout
            assert " U getenv" not in syms, f


def test_custom_ops_registered_cuda_only(libpath):
    """north_star: the forward path calls one thin C-ABI torch custom op.  The ops exist in the dispatcher and have NO CPU
    kernel (a CPU call fails in the dispatcher instead of falling back)."""
    import iadmm_b200  # noqa: F401
    for name in ("solve", "ruiz", "residuals", "build_kkt"):
        assert hasattr(torch.ops.iadmm, name)
    z = torch.zeros(4)
    with pytest.raises(NotImplementedError):
        torch.ops.iadmm.residuals(z, z, z, z.reshape(1, 2, 2), z, z.reshape(1, 2, 2), z, z, z)


def test_gate_mode_table_matches_header():
    from iadmm_b200 import _lib
    src = open(os.path.join(ROOT, "include", "iadmm.h")).read()
    enum = dict((k.lower(), int(v)) for k, v in re.findall(r"IADMM_GATES_([A-Z0-9_]+)\s*=\s*(\d+)", src))
    assert enum == _lib.GATE_MODES


def test_model_pickles_and_deepcopies_without_its_device_scratch():
    """`torch.save(model)` / `copy.deepcopy(model)` (EarlyStopping-style checkpointing of whole modules) carry the 16 parameters;
    workspaces, packed weights and the per-call records of `forward` (weak references) are dropped and rebuilt lazily."""
    import copy
    import io
    import weakref
    import iadmm_b200 as ia
    m = ia.LSTM(None, 2, 8, 3, "cpu")
    t = torch.zeros(3)
    m._chain = {"dims": (1,), "odd": True, "H": (weakref.ref(t), 0, 0), "C": (weakref.ref(t), 0, 0)}
    m._kkt_shared = ((1,), weakref.ref(t), weakref.ref(t), t)
    m._ws = torch.zeros(4, dtype=torch.uint8)
    buf = io.BytesIO()
    torch.save(m, buf)
    buf.seek(0)
    m2 = torch.load(buf, weights_only=False)
    m3 = copy.deepcopy(m)
    for other in (m2, m3):
        assert other._chain is None and other._kkt_shared is None and other._ws is None and other._packed is None
        assert list(other.state_dict().keys()) == list(m.state_dict().keys())
        for k, v in m.state_dict().items():
            assert torch.equal(other.state_dict()[k], v)
    assert m._chain is not None and m._ws is not None            # the original keeps its own
