"""Hash-partitioned equi-join over the ranks of one node (BASELINE config 5).

Plumbing only: the C-ABI does the work (cqg_partition_rows, cqg_execute_partial_rows, cqg_partial_*), torch
holds the device buffers and `torch.distributed` (NCCL over NVLink) moves them:

  every rank    scans its byte-range shard of both files, splits the row offsets by key owner
  all-to-all    8-byte global row offsets, once per side (sizes first, then the offsets)
  every rank    builds the join table over the right rows it owns, probes with the left rows it owns, aggregates
  all-gather    the partial group records (they are few next to the rows), merged and finished on every rank

Both tables must be open on every rank over the WHOLE file (the join reads rows in place by offset).
`exchange=None` runs the same steps for `world` simulated ranks in one process (tests on one GPU).
"""
import ctypes as C

import torch

from .engine import _check, decode_result


def partition(lib, table, key_col, world):
    """-> (int64 tensor of offsets on the device, list of per-owner counts, key-class mask) for the table's current shard."""
    h = C.c_void_p()
    _check(lib, lib.partition_rows(table.handle, key_col, world, C.byref(h)))
    try:
        classes = int(lib.rowlist_key_classes(h))
        counts = (C.c_int64 * world)()
        _check(lib, lib.rowlist_counts(h, world, counts))
        counts = [int(c) for c in counts]
        rows = torch.empty(max(sum(counts), 1), dtype=torch.int64, device="cuda")
        _check(lib, lib.rowlist_copy(h, rows.data_ptr(), rows.numel()))
    finally:
        lib.rowlist_free(h)
    return rows[:sum(counts)], counts, classes


class MixedKeyClasses(RuntimeError):
    """Join keys of more than one comparison class: the reference's cross-type "equal" is not reproduced."""


def _check_classes(mask):
    nonnull = mask & ~1
    if nonnull & (nonnull - 1):
        from . import _abi as A
        from .engine import CqError
        raise CqError(A.ERR_UNSUPPORTED, "join key columns mix comparison classes (cross-type value_compare)")


def _agree(dist, err, what):
    """Collective error check: every rank learns whether ANY rank failed, BEFORE the next collective - a rank
    that raised on its own would leave the others blocked in NCCL until the watchdog fires. Re-raises the local
    error, or a generic one on the ranks that were fine."""
    flag = torch.tensor([1 if err is not None else 0], dtype=torch.int64, device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    if int(flag.item()) != 0:
        if err is not None:
            raise err
        raise RuntimeError(f"another rank failed in {what}")


def _all_to_all_rows(rows, counts, dist):
    """segment o of `rows` goes to rank o; returns what this rank received (int64 tensor)."""
    world = dist.get_world_size()
    send_n = torch.tensor(counts, dtype=torch.int64, device="cuda")
    recv_n = torch.empty(world, dtype=torch.int64, device="cuda")
    dist.all_to_all_single(recv_n, send_n)
    recv_counts = [int(x) for x in recv_n.tolist()]
    out = torch.empty(max(sum(recv_counts), 1), dtype=torch.int64, device="cuda")[:sum(recv_counts)]
    dist.all_to_all_single(out, rows, output_split_sizes=recv_counts, input_split_sizes=counts)
    return out


def _merge_and_finish(lib, left, plan, parts_records, like):
    """parts_records: list of (device tensor of records, n). Returns the decoded result."""
    merged = C.c_void_p()
    _check(lib, lib.partial_new_like(like, C.byref(merged)))
    try:
        for buf, n in parts_records:
            if n:
                _check(lib, lib.partial_merge(merged, buf.data_ptr(), n))
        from . import _abi as A
        res = C.POINTER(A.Result)()
        _check(lib, lib.partial_finish(merged, left.handle, C.byref(res)))
        try:
            return decode_result(res.contents, plan)
        finally:
            lib.result_free(res)
    finally:
        lib.partial_free(merged)


def _export(lib, p):
    rec = lib.partial_record_size(p)
    n = lib.partial_count(p)
    buf = torch.empty(max(n, 1) * rec, dtype=torch.uint8, device="cuda")
    got = C.c_int64()
    _check(lib, lib.partial_export(p, 0, 1, buf.data_ptr(), max(n, 1), C.byref(got)))
    return buf, int(got.value), rec


def join_aggregate(lib, left, right, plan, world=None, dist=None):
    """Aggregate over `left JOIN right` (plan.q.join set to `right`), hash-partitioned over the ranks.

    dist given (initialised process group, one rank per GPU): this rank's part of the collective run; every rank
    returns the full result. dist None: `world` ranks simulated one after the other in this process."""
    q = plan.q
    lcol, rcol = q.join.left_col, q.join.right_col
    stats = {"kernel_ms": 0.0, "rows_exchanged": 0}
    if dist is not None:
        world, rank = dist.get_world_size(), dist.get_rank()
        left.set_shard(rank, world)
        right.set_shard(rank, world)
        err = None
        lrows = rrows = None
        lcounts = rcounts = None
        try:
            lrows, lcounts, lcls = partition(lib, left, lcol, world)
            rrows, rcounts, rcls = partition(lib, right, rcol, world)
            classes = lcls | rcls
        except Exception as e:  # a key this rank's shard holds and the kernels decline, out of memory, ...
            err = e
            classes = 0
        finally:
            left.set_shard(0, 1)
            right.set_shard(0, 1)
        _agree(dist, err, "the partition scan")
        # key classes over ALL ranks and both sides (per rank the mix can be invisible): every rank declines together
        cm = torch.tensor([(classes >> b) & 1 for b in range(4)], dtype=torch.int64, device="cuda")
        dist.all_reduce(cm, op=dist.ReduceOp.MAX)
        _check_classes(sum(int(v) << b for b, v in enumerate(cm.tolist())))
        mine_l = _all_to_all_rows(lrows, lcounts, dist)
        mine_r = _all_to_all_rows(rrows, rcounts, dist)
        stats["rows_exchanged"] = int(lrows.numel() + rrows.numel())
        p = C.c_void_p()
        rc = lib.execute_partial_rows(left.handle, C.byref(q), mine_l.data_ptr(), mine_l.numel(), mine_r.data_ptr(),
                                      mine_r.numel(), C.byref(p))
        # a rank that must decline (mixed key classes, ...) makes every rank decline
        flag = torch.tensor([rc], dtype=torch.int64, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        if int(flag.item()) != 0:
            if rc == 0:
                lib.partial_free(p)
                raise RuntimeError("another rank declined the partitioned join")
            _check(lib, rc)
        try:
            stats["kernel_ms"] = lib.partial_kernel_ms(p)
            err, buf, n, rec = None, None, 0, 0
            try:
                buf, n, rec = _export(lib, p)
            except Exception as e:
                err = e
            _agree(dist, err, "the export of the partial aggregates")
            ns = torch.empty(world, dtype=torch.int64, device="cuda")
            dist.all_gather_into_tensor(ns, torch.tensor([n], dtype=torch.int64, device="cuda"))
            ns = [int(x) for x in ns.tolist()]
            nmax = max(max(ns), 1)
            send = torch.zeros(nmax * rec, dtype=torch.uint8, device="cuda")
            send[:n * rec] = buf[:n * rec]
            recv = torch.empty(world * nmax * rec, dtype=torch.uint8, device="cuda")
            dist.all_gather_into_tensor(recv, send)
            parts = [(recv[r * nmax * rec:(r + 1) * nmax * rec], ns[r]) for r in range(world)]
            err, out = None, None
            try:
                out = _merge_and_finish(lib, left, plan, parts, p)
            except Exception as e:
                err = e
            _agree(dist, err, "the merge of the partial aggregates")  # (so that no rank runs ahead into a later collective)
        finally:
            lib.partial_free(p)
        out["stats"] = stats
        return out
    # one process, `world` simulated ranks
    world = world or 2
    lparts, rparts = [], []
    classes = 0
    for r in range(world):
        left.set_shard(r, world)
        right.set_shard(r, world)
        lr, lc, lk = partition(lib, left, lcol, world)
        rr, rc_, rk = partition(lib, right, rcol, world)
        lparts.append((lr, lc))
        rparts.append((rr, rc_))
        classes |= lk | rk
    left.set_shard(0, 1)
    right.set_shard(0, 1)
    _check_classes(classes)

    def received(parts, owner):
        segs = []
        for rows, counts in parts:
            start = sum(counts[:owner])
            segs.append(rows[start:start + counts[owner]])
        return torch.cat(segs) if segs else torch.empty(0, dtype=torch.int64, device="cuda")

    partials, records = [], []
    try:
        for o in range(world):
            ml, mr = received(lparts, o), received(rparts, o)
            p = C.c_void_p()
            _check(lib, lib.execute_partial_rows(left.handle, C.byref(q), ml.data_ptr(), ml.numel(), mr.data_ptr(), mr.numel(),
                                                 C.byref(p)))
            partials.append(p)
            stats["kernel_ms"] += lib.partial_kernel_ms(p)
            buf, n, _ = _export(lib, p)
            records.append((buf, n))
        out = _merge_and_finish(lib, left, plan, records, partials[0])
        out["rows_scanned"] = sum(lib.partial_rows_scanned(p) for p in partials)
    finally:
        for p in partials:
            lib.partial_free(p)
    out["stats"] = stats
    return out
