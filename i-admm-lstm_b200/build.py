"""Build libiadmm_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python i-admm-lstm_b200/build.py [--force] [--verbose]
    IADMM_DEV_BUILD=1 python i-admm-lstm_b200/build.py      # -> libiadmm_b200_dev.so with the development switches

The release library never reads the environment; the IADMM_TC_* / IADMM_RESIDENT / ... development switches used by
tools/ exist only in the development build (-DIADMM_DEV_SWITCHES), which is loaded explicitly with
IADMM_B200_LIB=<path to libiadmm_b200_dev.so>.

The library has no torch / libcuda link-time dependency (cudart is linked statically; the one driver
entry point needed for TMA descriptors is resolved at run time), so it loads on a CPU-only box for the
symbol checks of the CPU test-suite.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "iadmm_b200", "libiadmm_b200.so")
SOURCES = ["api.cu", "pack.cu", "kkt.cu", "sparse.cu", "ruiz.cu", "gates_simt.cu", "gates_tc.cu", "gemm_tc.cu", "train.cu", "lu.cu", "resident.cu", "comm.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unknown-pragmas", "--expt-relaxed-constexpr",
              "-cudart", "static"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, dev=None):
    nvcc = _nvcc()
    dev = (os.environ.get("IADMM_DEV_BUILD") == "1") if dev is None else dev
    target = OUT.replace(".so", "_dev.so") if dev else OUT
    flags = NVCC_FLAGS + (["-DIADMM_DEV_SWITCHES"] if dev else [])
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(HERE, "..", "include", "iadmm.h"))
    objdir = os.path.join(HERE, "build", "dev") if dev else os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = [nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- nvcc {src} ---\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    if force or procs or _stale(target, objs):
        cmd = [nvcc, "-shared", "-o", target] + objs + ["-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-ldl"]
        subprocess.check_call(cmd)
    return target


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
