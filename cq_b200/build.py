"""Build every native piece in-tree.

  cq_b200/libcqgpu.so            CUDA kernels + C-ABI, nvcc for sm_100a only (cross-compiles without a GPU)
  build/cq_gpu, cq_gpu_dump      cq's host C + this repo's dispatcher + libcqgpu (needs the reference sources)
  oracle/liboracle.so            CPU restatement — test infrastructure
  oracle/_ref/*                  the unmodified reference, compiled — test infrastructure

`python -m cq_b200.build [--force]`
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "cq_b200", "csrc")
LIB = os.path.join(ROOT, "cq_b200", "libcqgpu.so")
CQ_REF = os.environ.get("CQ_REF", "/root/reference")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared",
]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_cuda(force=False, verbose=False):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [os.path.join(ROOT, "include", "cq_gpu.h")]
    if not force and not _newer(LIB, srcs):
        return LIB
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [
        "-I" + os.path.join(ROOT, "include"), "-I" + CSRC, "-o", LIB, os.path.join(CSRC, "cqg_api.cu")]
    subprocess.run(cmd, check=True)
    return LIB


def build_oracle():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    if os.path.isdir(os.path.join(CQ_REF, "src")):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref", f"CQ_REF={CQ_REF}"], check=True)


def build_host():
    """The drop-in binaries; only where the reference sources are present."""
    if not os.path.isdir(os.path.join(CQ_REF, "src")):
        return False
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "cq_b200", "host"), f"CQ_REF={CQ_REF}"], check=True)
    return True


def build_all(force=False, verbose=False):
    build_cuda(force, verbose)
    build_oracle()
    build_host()


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built:", LIB)
