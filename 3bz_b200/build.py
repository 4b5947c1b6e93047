"""Builds libthreebz_cuda.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libthreebz_cuda.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread", "--cudart", "static",
         "-I", os.path.join(ROOT, "include"), "-I", CSRC]


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]


def _stale():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "threebz_cuda.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, extra=(), out=None):
    """out: alternative output path (tuning variants built with -D overrides; see tools/)."""
    if out is None and not force and not _stale():
        return SO
    cmd = [NVCC] + FLAGS + list(extra) + ["-o", out or SO] + sources()
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True,
          extra=["-Xptxas", "-v"] if "--ptxas" in sys.argv else ())
    print(SO)
