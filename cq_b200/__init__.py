"""cq_b200 — B200-native scan / filter / GROUP BY / equi-JOIN hot path of krow89/cq.

csrc/   hand-written sm_100a CUDA kernels + the C-ABI (include/cq_gpu.h) -> libcqgpu.so
host/   the C dispatcher that drops in behind cq's evaluate_query (reference language: C)
engine  ctypes plumbing used by tests/ and bench.py
"""
from . import _abi  # noqa: F401
from .engine import CqError, Plan, Table, csv_config, gpu  # noqa: F401
