"""CPU-side check of the per-query kernels (cqg_jit, cq_b200/csrc/cqg_api.cu): the lean kernel headers must
compile under NVRTC for sm_100a with a plan's shape given as macros — no GPU is needed to compile."""
import ctypes as C
import os

import pytest

from conftest import ROOT

CSRC = os.path.join(ROOT, "cq_b200", "csrc")
INC = os.path.join(ROOT, "include")

SHAPE = """#define CQG_JIT 1
#define CQG_JIT_NWANT 3
#define CQG_JIT_GAP0 0
#define CQG_JIT_GAP1 2
#define CQG_JIT_GAP2 2
#define CQG_JIT_GAP3 0
#define CQG_JIT_NPROG 3
#define CQG_JIT_NLEAF 2
#define CQG_JIT_NGC 1
#define CQG_JIT_CRLF 0
#define CQG_JIT_NAGG 2
#define CQG_JIT_PROG(i) ((i)==0?0:(i)==1?1:(i)==2?-1:0)
#define CQG_JIT_LEAFSLOT(i) ((i)==0?1:(i)==1?0:0)
#define CQG_JIT_LEAFKIND(i) ((i)==0?0:(i)==1?1:0)
#define CQG_JIT_GSLOT(i) ((i)==0?0:0)
#define CQG_JIT_ASLOT(i) ((i)==0?2:(i)==1?1:0)
#define CQG_JIT_AFUNC(i) ((i)==0?3:(i)==1?2:0)
#define CQG_JIT_PKIDW 3
#define CQG_JIT_PKBYTES 64
#define CQG_JIT_PKCOUNT 24
#define CQG_JIT_PKKEYWORD(i) ((i)==0?1:0)
#define CQG_JIT_PKKEYWIDE(i) ((i)==0?1:0)
#define CQG_JIT_PKAGGOFF(i) ((i)==0?32:(i)==1?-1:0)
#define CQG_JIT_PKAGGKEY(i) ((i)==0?-1:(i)==1?0:0)
"""

KERNELS = [
    ("cqg_lean2.cuh", "cqg::lean2_kernel<cqg::Geo<128, 16384, 1, 224>, 8, false, -1>"),
    ("cqg_lean2g.cuh", "cqg::lean2g_kernel<cqg::Geo<128, 16384, 1, 224>, 6>"),
    ("cqg_lean2k.cuh", "cqg::lean2k_kernel<cqg::Geo<128, 16384, 1, 224>, 8>"),
    ("cqg_leanhc.cuh", "cqg::leanhc_kernel<cqg::Geo<128, 16384, 1, 224>, 6>"),
    ("cqg_lean.cuh", "cqg::lean_kernel<cqg::Geo<128, 16384, 1, 992>, 5, true, false, false>"),
    ("cqg_lean.cuh", "cqg::lean_kernel<cqg::Geo<128, 16384, 1, 992>, 6, false, false, true>"),
]


def _nvrtc_libs():
    """Every run-time compiler a process might end up with: the toolkit's, and the one PyTorch bundles (an older CUDA:
    a process that imported torch first resolves libnvrtc.so.12 to that one)."""
    import glob
    import sys
    names = ["libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so"]
    for sp in sys.path:
        names += glob.glob(os.path.join(sp, "nvidia", "cuda_nvrtc", "lib", "libnvrtc.so.*"))
    libs, seen = [], set()
    for name in names:
        try:
            lib = C.CDLL(name)
        except OSError:
            continue
        major, minor = C.c_int(), C.c_int()
        lib.nvrtcVersion(C.byref(major), C.byref(minor))
        if (major.value, minor.value) not in seen:
            seen.add((major.value, minor.value))
            libs.append(lib)
    return libs


@pytest.mark.parametrize("header,name", KERNELS)
def test_lean_kernels_compile_at_run_time(header, name):
    libs = _nvrtc_libs()
    if not libs:
        pytest.skip("libnvrtc not found")
    for nv in libs:
        _compile(nv, header, name)


def _compile(nv, header, name):
    src = (SHAPE + f'#include "{header}"\n').encode()
    prog = C.c_void_p()
    assert nv.nvrtcCreateProgram(C.byref(prog), src, b"cqg_jit.cu", 0, None, None) == 0
    try:
        assert nv.nvrtcAddNameExpression(prog, name.encode()) == 0
        opts = [b"--gpu-architecture=sm_100a", b"--std=c++17", b"-default-device", f"-I{CSRC}".encode(), f"-I{INC}".encode(),
                b"-I/usr/local/cuda/include"]
        rc = nv.nvrtcCompileProgram(prog, len(opts), (C.c_char_p * len(opts))(*opts))
        n = C.c_size_t()
        nv.nvrtcGetProgramLogSize(prog, C.byref(n))
        log = C.create_string_buffer(max(n.value, 1))
        nv.nvrtcGetProgramLog(prog, log)
        assert rc == 0, log.value.decode(errors="replace")[-3000:]
        low = C.c_char_p()
        assert nv.nvrtcGetLoweredName(prog, name.encode(), C.byref(low)) == 0 and low.value
        nv.nvrtcGetCUBINSize(prog, C.byref(n))
        assert n.value > 10_000  # a real cubin, and nothing but this kernel in it (seconds, not minutes, per shape)
        assert n.value < 1_000_000
    finally:
        nv.nvrtcDestroyProgram(C.byref(prog))
