"""Build libkrisp_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libkrisp_b200.so")
SOURCES = ["kb_api.cu"]
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith(".cuh")) + [os.path.join("..", "..", "include", "krisp_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "--use_fast_math", "-Xcompiler", "-fPIC,-O2", "-shared", "-Xptxas", "-v"]


def _nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    """Compile the CUDA library if sources are newer than the .so; returns its path."""
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + ["-o", LIB] + [os.path.join(CSRC, f) for f in SOURCES]
    env = dict(os.environ)
    env.pop("CC", None)   # the image exports a gcc wrapper that nvcc must not pick up
    env.pop("CXX", None)
    p = subprocess.run(cmd, capture_output=True, text=True, env=env)
    log = os.path.join(HERE, "build.log")
    with open(log, "w") as fh:
        fh.write(" ".join(cmd) + "\n" + p.stdout + p.stderr)
    if verbose or p.returncode != 0:
        sys.stderr.write(p.stdout + p.stderr)
    if p.returncode != 0:
        raise RuntimeError(f"nvcc failed (see {log})")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
