"""In-tree build of liblmpc_b200.so (nvcc, sm_100a only).  No JIT cache: the .so sits next to this
file so that it travels with the repository snapshot to the GPU box."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "lmpc_b200.cu")
DEPS = [SRC] + [os.path.join(HERE, "csrc", f) for f in ("ftocp_pdip.cuh", "safeset.cuh", "lapbooks.cuh", "probe.cuh")] + \
       [os.path.join(HERE, "..", "include", "lmpc_b200.h")]
OUT = os.path.join(HERE, "liblmpc_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-shared", "-ccbin", "/usr/bin/g++"]


def stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in DEPS)


def build_native(force=False, verbose=False):
    if not force and not stale():
        return OUT
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT, SRC]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout)
    if verbose:
        print(r.stdout)
    return OUT


if __name__ == "__main__":
    build_native(force=True, verbose=True)
