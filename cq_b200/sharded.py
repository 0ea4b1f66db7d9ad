"""Byte-range sharded aggregates over the ranks of one node (BASELINE configs 2 and 3 at 2/4/8 GPUs).

Plumbing only: every rank scans the rows whose first byte lies in its slice (cqg_execute_partial), the partial
group records cross NVLink with ONE collective, and are folded together on the device (cqg_partial_merge):

  GatherExchange   few groups: fixed-capacity `all_gather_into_tensor` of the records (no count exchange, no host
                   round trip inside the step), merged and finished on every rank.
  OwnerExchange    many groups (~10^6): records split by owner = hash % world (cqg_partial_export with an owner),
                   `all_to_all_single`, every rank merges and finishes the groups it owns.

A merged group's bare columns are read from its first row by a rank that holds that row; plans run through these
classes therefore carry no bare columns unless every rank views the whole file.
"""
import ctypes as C

import torch

from . import _abi as A
from .engine import _check


class GatherExchange:
    def __init__(self, lib, dist, capacity=64):
        self.lib, self.dist, self.cap = lib, dist, capacity
        self.world = dist.get_world_size()
        self.rec = None
        self.send = self.recv = None

    def _buffers(self, rec):
        if self.rec != rec:
            self.rec = rec
            self.send = torch.zeros(self.cap * rec, dtype=torch.uint8, device="cuda")
            self.recv = torch.empty(self.world * self.cap * rec, dtype=torch.uint8, device="cuda")

    def step(self, table, plan):
        lib = self.lib
        p = C.c_void_p()
        _check(lib, lib.execute_partial(table.handle, C.byref(plan.q), C.byref(p)))
        try:
            self._buffers(lib.partial_record_size(p))
            self.send.zero_()
            got = C.c_int64()
            rc = lib.partial_export(p, 0, 1, self.send.data_ptr(), self.cap, C.byref(got))
            if rc != A.OK and got.value > self.cap:  # more groups than the exchange buffer holds: grow it, all ranks alike
                raise RuntimeError(f"GatherExchange capacity {self.cap} < {got.value} groups")
            _check(lib, rc)
            self.dist.all_gather_into_tensor(self.recv, self.send)
            m = C.c_void_p()
            _check(lib, lib.partial_new_like(p, C.byref(m)))
            try:
                _check(lib, lib.partial_merge(m, self.recv.data_ptr(), self.world * self.cap))
                res = C.POINTER(A.Result)()
                _check(lib, lib.partial_finish(m, table.handle, C.byref(res)))
                c = res.contents
                G = c.n_groups
                out = {"n_groups": G, "count": [c.count[g] for g in range(min(G, self.cap))],
                       "first_offset": [c.first_offset[g] for g in range(min(G, self.cap))],
                       "sum": [[c.sum[a * G + g] for g in range(min(G, self.cap))] for a in range(c.n_aggs)],
                       "kernel_ms": lib.partial_kernel_ms(p), "rows_scanned": lib.partial_rows_scanned(p)}
                lib.result_free(res)
            finally:
                lib.partial_free(m)
        finally:
            lib.partial_free(p)
        return out


class OwnerExchange:
    def __init__(self, lib, dist):
        self.lib, self.dist = lib, dist
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]

    def step(self, table, plan):
        lib, world = self.lib, self.world
        p = C.c_void_p()
        _check(lib, lib.execute_partial(table.handle, C.byref(plan.q), C.byref(p)))
        try:
            rec = lib.partial_record_size(p)
            counts = (C.c_int64 * world)()
            _check(lib, lib.partial_owner_counts(p, world, counts))
            counts = [int(c) for c in counts]
            send = torch.empty(max(sum(counts), 1) * rec, dtype=torch.uint8, device="cuda")
            off = 0
            for o in range(world):
                got = C.c_int64()
                _check(lib, lib.partial_export(p, o, world, send.data_ptr() + off * rec, counts[o], C.byref(got)))
                assert got.value == counts[o]
                off += counts[o]
            self.ev[0].record()
            send_n = torch.tensor(counts, dtype=torch.int64, device="cuda")
            recv_n = torch.empty(world, dtype=torch.int64, device="cuda")
            self.dist.all_to_all_single(recv_n, send_n)
            recv_counts = [int(x) for x in recv_n.tolist()]
            recv = torch.empty(max(sum(recv_counts), 1) * rec, dtype=torch.uint8, device="cuda")
            self.dist.all_to_all_single(recv[:sum(recv_counts) * rec], send[:sum(counts) * rec],
                                        output_split_sizes=[c * rec for c in recv_counts],
                                        input_split_sizes=[c * rec for c in counts])
            self.ev[1].record()
            m = C.c_void_p()
            _check(lib, lib.partial_new_like(p, C.byref(m)))
            try:
                _check(lib, lib.partial_merge(m, recv.data_ptr(), sum(recv_counts)))
                res = C.POINTER(A.Result)()
                _check(lib, lib.partial_finish(m, table.handle, C.byref(res)))
                c = res.contents
                G = c.n_groups
                # cheap whole-result facts for the cross-rank checks: groups owned, rows in them, sum of SUM(agg 1)
                import numpy as np
                cnt = np.ctypeslib.as_array(c.count, shape=(max(G, 1),))[:G]
                out = {"n_groups": G, "rows_in_groups": int(cnt.sum()), "records_sent": sum(counts), "record_bytes": rec,
                       "kernel_ms": lib.partial_kernel_ms(p), "rows_scanned": lib.partial_rows_scanned(p)}
                if c.n_aggs > 1:
                    out["sum_agg1"] = float(np.ctypeslib.as_array(c.sum, shape=(c.n_aggs * max(G, 1),))[G:2 * G].sum())
                lib.result_free(res)
            finally:
                lib.partial_free(m)
            torch.cuda.synchronize()
            out["exchange_ms"] = self.ev[0].elapsed_time(self.ev[1])
        finally:
            lib.partial_free(p)
        return out
