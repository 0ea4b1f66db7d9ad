#!/usr/bin/env python
"""Compile one per-query kernel with NVRTC on the CPU box (as cqg_jit does at run time) and report what ptxas made of it.

  python tools/jit_compile.py cqg_lean2k.cuh 'cqg::lean2k_kernel<cqg::Geo<128, 16384, 1, 224>, 4>' [out.cubin] [--shape group_name]

Shapes are the macro blocks cqg_jit::lean_shape_defs would emit for plans of tests/parity_cases.py."""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "cq_b200", "csrc")
INC = os.path.join(ROOT, "include")

PK = """#define CQG_JIT_PKIDW 3
#define CQG_JIT_PKBYTES 64
#define CQG_JIT_PKCOUNT 24
#define CQG_JIT_PKKEYWORD(i) ((i)==0?1:0)
#define CQG_JIT_PKKEYWIDE(i) ((i)==0?1:0)
#define CQG_JIT_PKAGGOFF(i) ((i)==0?32:(i)==1?-1:0)
#define CQG_JIT_PKAGGKEY(i) ((i)==0?-1:(i)==1?0:0)
"""
SHAPES = {
    # SELECT name, COUNT(*), AVG(height), SUM(age) FROM f WHERE age > 25 GROUP BY name  (name,surname,age,gender,height)
    "group_name": """#define CQG_JIT 1
#define CQG_JIT_NWANT 3
#define CQG_JIT_GAP0 0
#define CQG_JIT_GAP1 2
#define CQG_JIT_GAP2 2
#define CQG_JIT_GAP3 0
#define CQG_JIT_NPROG 1
#define CQG_JIT_NLEAF 1
#define CQG_JIT_NGC 1
#define CQG_JIT_CRLF 0
#define CQG_JIT_NAGG 2
#define CQG_JIT_PROG(i) ((i)==0?0:0)
#define CQG_JIT_LEAFSLOT(i) ((i)==0?1:0)
#define CQG_JIT_LEAFKIND(i) ((i)==0?0:0)
#define CQG_JIT_GSLOT(i) ((i)==0?0:0)
#define CQG_JIT_ASLOT(i) ((i)==0?2:(i)==1?1:0)
#define CQG_JIT_AFUNC(i) ((i)==0?3:(i)==1?2:0)
""" + PK,
}


def nvrtc():
    for name in ("libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so"):
        try:
            return C.CDLL(name)
        except OSError:
            continue
    sys.exit("libnvrtc not found")


def compile_kernel(header, name, shape):
    nv = nvrtc()
    src = (SHAPES[shape] + f'#include "{header}"\n').encode()
    prog = C.c_void_p()
    assert nv.nvrtcCreateProgram(C.byref(prog), src, b"cqg_jit.cu", 0, None, None) == 0
    assert nv.nvrtcAddNameExpression(prog, name.encode()) == 0
    opts = [b"--gpu-architecture=sm_100a", b"--std=c++17", b"-default-device", b"-lineinfo", f"-I{CSRC}".encode(), f"-I{INC}".encode(),
            b"-I/usr/local/cuda/include"]
    rc = nv.nvrtcCompileProgram(prog, len(opts), (C.c_char_p * len(opts))(*opts))
    n = C.c_size_t()
    nv.nvrtcGetProgramLogSize(prog, C.byref(n))
    log = C.create_string_buffer(max(n.value, 1))
    nv.nvrtcGetProgramLog(prog, log)
    if rc != 0:
        sys.exit(log.value.decode(errors="replace")[-6000:])
    nv.nvrtcGetCUBINSize(prog, C.byref(n))
    cubin = C.create_string_buffer(n.value)
    nv.nvrtcGetCUBIN(prog, cubin)
    return cubin.raw


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    shape = "group_name"
    if "--shape" in sys.argv:
        shape = sys.argv[sys.argv.index("--shape") + 1]
        args = [a for a in args if a != shape]
    header, name = args[0], args[1]
    out = args[2] if len(args) > 2 else "/tmp/jit.cubin"
    open(out, "wb").write(compile_kernel(header, name, shape))
    res = subprocess.run(["cuobjdump", "-res-usage", out], capture_output=True, text=True).stdout
    print(res.strip())
    sass = subprocess.run(["cuobjdump", "-sass", out], capture_output=True, text=True).stdout
    lines = [l for l in sass.splitlines() if "/*0" in l and ";" in l]
    print("SASS instructions (static):", len(lines), "->", out)
