#!/usr/bin/env python
"""bench.py — CSV GB/s of cq's scan + filter + GROUP BY hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--bytes B] [--no-extra]

One step = one pass of the hot path over one batch of synthetic CSV (the seeded restatement of the reference's
utils/generate_big_dataset.py, generated on the device). The file is 10^10 bytes (far larger than the 126 MB L2,
so nothing is flushed between steps).

Headline (`value`, `roofline`, `e2e`, the reference arm): the metric's own shape, scan + filter + GROUP BY —
BASELINE configs[0]'s query on the 10 GB file:
    SELECT name, COUNT(*), AVG(height), SUM(age) FROM f WHERE age > 25 GROUP BY name
Timed beside it in the same run, W warm-up + K timed steps each, with a roofline object each (`configs`):
    configs[1]  SELECT COUNT(*) FROM f WHERE age > 40                                  (pure parse + filter)
    configs[2]  GROUP BY name, surname, age, height (1 835 776 keys) + SUM/MIN/MAX/AVG (high cardinality)

N = 1: `value` = bytes / CUDA-event time of K steps through cqg_execute on the resident table; `e2e` = the same query
through cqg_table_open(path) on a page-cached file (mmap -> parallel pinned staging -> scan -> result on the host).
N > 1 (torchrun, one rank per GPU): STRONG scaling — the same 10 GB file, every rank holds and scans 1/N of it
(byte-range slices cut at row edges), the partial group records cross NVLink with one NCCL collective and are merged
on the device (cq_b200/sharded.py); value = file bytes / max-over-ranks time. `weak` repeats the headline with 10 GB
per rank. Results are asserted against the ranks' own single-scan answers.

Prints ONE JSON line (rank 0). Keys beyond the base contract: roofline, cpu_baseline, configs, weak, extra.
"""
import argparse
import ctypes as C
import json
import os
import shutil
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "csv_scan_filter_groupby_throughput"
UNIT = "GB/s"
QUERY = "SELECT name, COUNT(*), AVG(height), SUM(age) FROM f WHERE age > 25 GROUP BY name"
WORKLOAD = ("BASELINE configs[0] query shape (scan + filter + GROUP BY) on the 10 GB synthetic CSV of configs[1] "
            f"(seeded restatement of utils/generate_big_dataset.py): {QUERY}")
LEGS = {  # name -> (plan in tests/parity_cases.py, SQL text, kernel)
    "groupby": ("group_name", QUERY, "cqg::lean2k_kernel (compiled for the query; two-word dictionary buckets, per-warp shared-memory accumulators)"),
    "count_where": ("count_age_gt_40", "SELECT COUNT(*) FROM f WHERE age > 40", "cqg::lean2_kernel<ONELEAF>"),
    "high_card": ("group_high_card",
                  "SELECT name, surname, age, height, COUNT(*), SUM(age), MIN(height), MAX(height), AVG(height) FROM f "
                  "GROUP BY name, surname, age, height",
                  "cqg::leanhc_kernel (packed global table) + expand_packed_kernel"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--bytes", type=float, default=float(os.environ.get("CQ_BENCH_BYTES", 1e10)))
    ap.add_argument("--no-extra", action="store_true")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
# clocks sampled during the timed region (profiling recipe: clocks line)
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.samples, self.reasons, self.stop_flag, self.max_mhz = [], set(), False, None
        self.thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv:
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join()
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ---------------------------------------------------------------------------------------------
# the reference arm / CPU baseline: the reference's own CPU implementation on host cores
# ---------------------------------------------------------------------------------------------
def cpu_sample_file(nbytes, seed=1):
    """A bounded sample of the same workload on disk: the first rows of the seeded file."""
    from oracle_lib import generate_bigdata
    rows = max(1000, int(nbytes / 29.89))
    data = generate_bigdata(rows, seed=seed)
    path = f"/tmp/cq_bench_sample_{os.getpid()}.csv"
    with open(path, "wb") as f:
        f.write(data)
    return path, len(data), rows


def cpu_run_once(path):
    """Time the reference CLI (oracle/_ref/cq, compiled from the unmodified sources) on `path`;
    falls back to the CPU restatement when the compiled reference is not there."""
    ref_cq = os.path.join(ROOT, "oracle", "_ref", "cq")
    sql = QUERY.replace(" f ", f" '{path}' ")
    if os.path.exists(ref_cq):
        t0 = time.perf_counter()
        p = subprocess.run([ref_cq, "-q", sql, "-p"], capture_output=True, timeout=3600)
        dt = time.perf_counter() - t0
        if p.returncode == 0:
            return dt, "reference", p.stdout.decode("latin1")
    from oracle_lib import oracle
    import parity_cases as pc
    from cq_b200.engine import Table
    t0 = time.perf_counter()
    with Table.open(path, lib=oracle()) as t:
        r = t.execute(pc.build(pc.plans()["group_name"]))
    dt = time.perf_counter() - t0
    return dt, "port", str(len(r["groups"]))


def cpu_rate_sample(seconds):
    """(path, bytes, rows) of a sample the reference needs about `seconds` for (calibrated on 20 MB, <= 1 GB)."""
    path, n, rows = cpu_sample_file(20e6)
    dt, kind, _ = cpu_run_once(path)
    os.unlink(path)
    nbytes = min(max(n / dt * seconds, 20e6), 1e9)
    return cpu_sample_file(nbytes)


def cpu_baseline(target_seconds=12.0):
    path, n, rows = cpu_rate_sample(target_seconds)
    dt, kind, _ = cpu_run_once(path)
    os.unlink(path)
    return {"value": n / dt / 1e9, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": f"first {rows} rows ({n / 1e6:.1f} MB) of the seeded file, {QUERY}, wall clock {dt:.2f} s, "
                      f"single thread (the reference has no threads)"}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    total = args.steps + args.warmup
    per_step = max(2.0, min(25.0, 170.0 / max(total, 1)))
    path, n, rows = cpu_rate_sample(per_step)
    times = []
    for i in range(total):
        dt, kind, _ = cpu_run_once(path)
        if i >= args.warmup:
            times.append(dt)
    os.unlink(path)
    ms = 1e3 * sum(times) / len(times)
    value = n / (ms / 1e3) / 1e9
    sample = f"first {rows} rows ({n / 1e6:.1f} MB) of the seeded file per step, {QUERY}, single thread"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "sample_bytes": n, "note": "the reference's CPU path on host cores; bounded sample per step"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# the B200 arm
# ---------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    if args.impl == "reference":
        reference_arm(args)
        return

    # stdout carries exactly one JSON line: anything libraries print there (NCCL's version banner,
    # torchrun notices) goes to stderr instead
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist

    import parity_cases as pc
    from cq_b200 import _abi as A
    from cq_b200.engine import Plan, Table, _check, gpu
    from cq_b200.sharded import GatherExchange, OwnerExchange

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        sys.exit("bench.py needs a CUDA device (there is no CPU fallback for the measured path)")
    torch.cuda.set_device(local)
    lib = gpu()
    _check(lib, lib.set_device(local))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = json.load(open(peaks_path))["hbm_gbs"]
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if world == 1:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return int(t.item())

    def make_slice(total_rows, nslices, index, seed=1):
        """rows [index*R/n, (index+1)*R/n) of the seeded file of `total_rows` rows, resident in HBM: a table whose
        byte 0 sits at its global offset in that file (header only in slice 0; the others view data rows only)."""
        r0 = total_rows * index // nslices
        r1 = total_rows * (index + 1) // nslices
        cap = lib.generate_bigdata_bound(r1 - r0, 0) + lib.device_padding()
        buf = torch.empty(cap, dtype=torch.uint8, device="cuda")
        size = C.c_size_t()
        _check(lib, lib.generate_bigdata_range(buf.data_ptr(), cap - lib.device_padding(), r0, r1 - r0, seed, 0,
                                               1 if index == 0 else 0, C.byref(size)))
        from cq_b200.engine import csv_config
        t = Table.from_device(buf.data_ptr(), size.value, cfg=csv_config(has_header=(index == 0)), lib=lib, keep=buf)
        return t, buf, size.value

    # DRAM bytes per CSV byte of each kernel, from THIS round's `ncu --set full` captures (profiles/r02_traffic.json, made by
    # tools/summarise_profiles.py out of gpurun_out/r02_*.ncu-rep: one launch on a 2 GB table each). bench.py cannot run
    # under ncu, so `traffic` is that measured ratio times this launch's algorithmic bytes, and says so.
    traffic_ratio = {}
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        traffic_ratio = {k: v["dram_bytes_per_csv_byte"] for k, v in tj.items()}
    except Exception:
        pass

    def roofline_of(nbytes, kernel_ms, kernel):
        ach = nbytes / (kernel_ms / 1e3) / 1e9
        key = "lean2k" if "lean2k" in kernel else "leanhc" if "leanhc" in kernel else "lean2" if "lean2_kernel" in kernel else None
        ratio = traffic_ratio.get(key)
        return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": (ratio * nbytes) if ratio else None,
                "kernel": kernel, "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": nbytes, "peak_source": peak_src,
                "frac_of_nominal_8TBs": ach / 8000.0,
                "traffic_note": (f"dram__bytes_read + dram__bytes_write = {ratio:.4f} x the table's bytes in this round's ncu --set full capture "
                                 f"of this kernel (profiles/r02_{key}_ncu.txt, one launch on a 2 GB table); scaled to this launch's bytes, "
                                 "not re-measured here") if ratio else "no capture of this kernel under profiles/"}

    total_rows = int(args.bytes / 29.89)

    # =========================================================================================
    # N = 1
    # =========================================================================================
    if world == 1:
        table, buf, nbytes = make_slice(total_rows, 1, 0)
        legs, clocks, launches = {}, None, 0
        for leg, (plan_name, sql, kernel) in LEGS.items():
            plan = pc.build(pc.plans()[plan_name])
            for _ in range(args.warmup):
                last = table.execute_raw(plan)
            sampler = ClockSampler(local) if leg == "groupby" else None
            torch.cuda.synchronize()
            if sampler:
                sampler.start()
            l0 = lib.total_kernel_launches()
            k0 = lib.kernel_launches_named(b"lean2k")
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            kms = []
            e0.record()
            for _ in range(args.steps):
                last = table.execute_raw(plan)
                kms.append(last["kernel_ms"])
            e1.record()
            torch.cuda.synchronize()
            if leg == "groupby" and lib.kernel_launches_named(b"lean2k") == k0:
                kernel = "cqg::lean2g_kernel (the run-time compiler was not available: lean2k_kernel exists only compiled per query)"
            if sampler:
                clocks = sampler.stop()
                launches = lib.total_kernel_launches() - l0
            ms = e0.elapsed_time(e1) / args.steps
            kernel_ms = sum(kms) / len(kms)
            legs[leg] = {"query": sql, "value": nbytes / (ms / 1e3) / 1e9, "unit": UNIT, "ms_per_step": ms,
                         "rows_per_s": last["rows_scanned"] / (ms / 1e3), "groups": last["n_groups"],
                         "first_group_count": int(last["count0"]), "steps": args.steps, "warmup": args.warmup,
                         "kernel_ms_per_step": [round(k, 3) for k in kms],
                         "roofline": roofline_of(nbytes, kernel_ms, kernel)}
        head = legs["groupby"]

        # ---- end to end through the reference-facing call: cqg_table_open(path) on a page-cached file ----
        e2e, e2e_pinned = None, None
        plan = pc.build(pc.plans()["group_name"])
        n_e2e = max(2, min(args.steps, 3))
        pinned = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        pinned.copy_(buf[:nbytes])
        torch.cuda.synchronize()

        def timed(fn):
            fn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                r = fn()
            torch.cuda.synchronize()
            return (time.perf_counter() - t0) / n_e2e, r

        def step_pinned():
            with Table.from_bytes((pinned.data_ptr(), nbytes), lib=lib, pinned=True) as t:
                return t.execute_raw(plan)

        dt, r = timed(step_pinned)
        assert r["count0"] == head["first_group_count"] and r["n_groups"] == head["groups"]
        e2e_pinned = {"value": nbytes / dt / 1e9, "unit": UNIT, "ms_per_step": dt * 1e3,
                      "note": "cqg_table_open_buffer on an already page-locked host buffer (DMA in place) + cqg_execute + close"}
        path = None
        for d in ("/dev/shm", "/tmp"):
            try:
                if shutil.disk_usage(d).free > nbytes * 1.2:
                    path = os.path.join(d, f"cq_bench_{os.getpid()}.csv")
                    break
            except OSError:
                pass
        if path:
            try:
                import numpy as np
                with open(path, "wb") as f:
                    view = np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(pinned.data_ptr()))
                    step = 256 << 20
                    for o in range(0, nbytes, step):
                        f.write(memoryview(view[o:o + step]))

                def step_file():
                    with Table.open(path, lib=lib) as t:
                        return t.execute_raw(plan)

                dt, r = timed(step_file)
                assert r["count0"] == head["first_group_count"] and r["n_groups"] == head["groups"]
                e2e = {"value": nbytes / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": nbytes,
                       "d2h_bytes_per_step": 24 * 16 * 4 + 64, "ms_per_step": dt * 1e3, "steps": n_e2e,
                       "note": f"cqg_table_open('{os.path.dirname(path)}/...', page-cached file: mmap + parallel staging through "
                               "page-locked bounce buffers) + cqg_execute + result on host + cqg_table_close per step"}
            finally:
                try:
                    os.unlink(path)
                except OSError:
                    pass
        if e2e is None:
            e2e = dict(e2e_pinned, h2d_bytes_per_step=nbytes, d2h_bytes_per_step=24 * 16 * 4 + 64, steps=n_e2e)
        e2e["pinned_buffer_route"] = e2e_pinned
        del pinned

        extra = {}
        if not args.no_extra:
            # ---- real-world bytes: the same two queries on 2 GB tables that hold a blank inside one field
            # of every row / end every row in CR LF, beside the clean table of the same size; with the share of 16 KB tiles
            # the lean kernels handed to the general kernel ----
            def stats():
                a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
                lib.last_scan_stats(C.byref(a), C.byref(b), C.byref(c))
                return {"tiles": a.value, "handed_tiles": b.value, "handed_rows": c.value,
                        "handed_tile_share": (b.value / a.value) if a.value else None}

            def variant_legs(kind):
                vt, vbuf, vn = make_slice(int(2e9 / 29.89), 1, 0)
                view = vbuf[:vn]
                if kind != "clean":
                    nl = torch.nonzero(view == 10).flatten()
                    if kind == "dirty_blank":      # `KKK KKKKKK,...`: a blank inside the first field of every data row
                        view[nl[:-1] + 4] = 32
                    elif kind == "crlf":           # every line ends in CR LF (the last digit of the row gives way to the CR)
                        view[nl - 1] = 13
                    del nl
                out = {}
                for leg in ("count_where", "groupby"):
                    pl = pc.build(pc.plans()[LEGS[leg][0]])
                    for _ in range(2):
                        r = vt.execute_raw(pl)
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    kms = []
                    for _ in range(3):
                        r = vt.execute_raw(pl)
                        kms.append(r["kernel_ms"])
                    torch.cuda.synchronize()
                    dt = (time.perf_counter() - t0) / 3
                    out[leg] = {"scan_kernel_gbs": vn / (sum(kms) / 3 / 1e3) / 1e9, "query_gbs": vn / dt / 1e9, "groups": r["n_groups"],
                                "count0": int(r["count0"]), **stats()}
                del vt, vbuf
                torch.cuda.empty_cache()
                return out

            try:
                clean = variant_legs("clean")
                for kind in ("dirty_blank", "crlf"):
                    v = variant_legs(kind)
                    for leg in v:
                        v[leg]["vs_clean"] = v[leg]["scan_kernel_gbs"] / clean[leg]["scan_kernel_gbs"]
                    extra[kind] = v
                extra["clean_2gb"] = clean
            except Exception as ex:
                extra["variants_error"] = repr(ex)
            # ---- BASELINE configs[4] at its stated size: 100 M x 10 M rows, equi-join on an integer key (the seeded
            # generator's `uid` column on both sides: left `...,uid` ~ U{0..R-1}, right one row per draw of the same range),
            # bytes(left) + bytes(right) read once ----
            try:
                scale = args.bytes / 1e10
                Lr, Rr = int(100e6 * scale), int(10e6 * scale)

                def gen(rows, seed):
                    cap = lib.generate_bigdata_bound(rows, Rr) + lib.device_padding()
                    b = torch.empty(cap, dtype=torch.uint8, device="cuda")
                    sz = C.c_size_t()
                    _check(lib, lib.generate_bigdata(b.data_ptr(), cap - lib.device_padding(), rows, seed, Rr, C.byref(sz)))
                    return Table.from_device(b.data_ptr(), sz.value, lib=lib, keep=b), sz.value

                lt, lbytes = gen(Lr, 11)
                rt, rbytes = gen(Rr, 12)
                UID, RIGHT = 5, 6
                jspecs = {"count": dict(aggs=[(pc.A.AGG_COUNT_STAR, -1)]),
                          "group_right_gender_sum_left_age": dict(group_by=[RIGHT + 3], out_cols=[RIGHT + 3],
                                                                  aggs=[(pc.A.AGG_COUNT_STAR, -1), (pc.A.AGG_SUM, 2)])}
                jl = {"left_rows": Lr, "right_rows": Rr, "bytes": lbytes + rbytes}
                for jn, js in jspecs.items():
                    pl = pc.build(js, join=(rt, UID, UID))
                    lt.execute_raw(pl)
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    r = lt.execute_raw(pl)
                    torch.cuda.synchronize()
                    dt = time.perf_counter() - t0
                    jl[jn] = {"query_gbs": (lbytes + rbytes) / dt / 1e9, "query_ms": dt * 1e3, "rows_per_s": (Lr + Rr) / dt,
                              "groups": r["n_groups"], "count0": int(r["count0"]), "kernel_ms": r["kernel_ms"]}
                extra["join_100m_x_10m"] = jl
                del lt, rt
                torch.cuda.empty_cache()
            except Exception as ex:
                extra["join_error"] = repr(ex)
            # ---- BASELINE configs[3]: quoted / escaped-comma CSV with string predicates (the corpus of tools/bench_quoted.py and
            # tests/test_gpu_large_parity.py: quoted names with embedded commas and doubled quotes; the general kernel) ----
            try:
                import random
                rnd = random.Random(4)
                first = ["Ada", "Brook", "Cyrus", "Dana", "Eli", "Fay", "Gus", "Hana", "Ivo", "Jude", "Max", "Xena", "Alex"]
                lastn = ["Smith", "Jones", "Lee", "Fox", "Marx", "Nguyen", "O'Neil", "Baxter"]
                block = []
                for i in range(100_000):
                    f, l = rnd.choice(first), rnd.choice(lastn)
                    nm = f'"{l}, {f}"' if i % 3 else (f'"say ""{f}"""' if i % 2 else f + l)
                    block.append(f"{nm},{rnd.choice(['admin', 'user', 'moderator'])},{rnd.randint(10, 80)},{rnd.randint(100, 200) / 100}")
                qdata = b"name,role,age,height\n" + ("\n".join(block) + "\n").encode() * max(1, int(200 * args.bytes / 1e10))
                ql = {"bytes": len(qdata)}
                with Table.from_bytes(qdata, lib=lib) as qt:
                    for qn, qs in {"role_eq_admin": dict(where=("=", ("col", 1), ("const", "admin")), aggs=[(pc.A.AGG_COUNT_STAR, -1)]),
                                   "name_like_x": dict(where=("like", ("col", 0), ("const", "%x%")), aggs=[(pc.A.AGG_COUNT_STAR, -1)])}.items():
                        pl = pc.build(qs)
                        qt.execute_raw(pl)
                        torch.cuda.synchronize()
                        t0 = time.perf_counter()
                        r = qt.execute_raw(pl)
                        torch.cuda.synchronize()
                        dt = time.perf_counter() - t0
                        ql[qn] = {"scan_kernel_gbs": len(qdata) / (r["kernel_ms"] / 1e3) / 1e9, "query_gbs": len(qdata) / dt / 1e9,
                                  "count0": int(r["count0"])}
                extra["quoted_csv"] = ql
                del qdata
            except Exception as ex:
                extra["quoted_error"] = repr(ex)
            for name in ["scalar_aggs", "lean_group_abort_many", "count_height_gt_1_5"]:
                pl = pc.build(pc.plans()[name])
                table.execute_raw(pl)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                r = table.execute_raw(pl)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                extra[name] = {"scan_kernel_gbs": nbytes / (r["kernel_ms"] / 1e3) / 1e9, "query_gbs": nbytes / dt / 1e9,
                               "groups": r["n_groups"], "kernel_ms": r["kernel_ms"], "query_ms": dt * 1e3}
        try:
            cpu = cpu_baseline()
        except Exception as ex:  # the baseline must never sink the bench line
            cpu = {"value": None, "unit": UNIT, "cores": 1, "kind": "unavailable", "sample": repr(ex)}
        line = {
            "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "bytes_per_gpu": nbytes, "total_bytes": nbytes,
                       "l2": "no flush: the 10 GB input is far larger than the 126 MB L2",
                       "parallelism": "one GPU", "groups": head["groups"], "first_group_count": head["first_group_count"]},
            "rows_per_s": head["rows_per_s"], "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "roofline": head["roofline"], "cpu_baseline": cpu,
            "configs": {"configs[0]-shape groupby (headline)": legs["groupby"], "configs[1] count_where": legs["count_where"],
                        "configs[2] high_card": legs["high_card"]},
            "extra": extra,
        }
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
        return

    # =========================================================================================
    # N > 1: strong scaling of ONE file; weak scaling as a second figure
    # =========================================================================================
    def no_out(spec):  # merged groups' bare columns live on whichever rank holds their first row: not fetched here
        s = dict(spec)
        s["out_cols"] = []
        return s

    def truth_counts(table, spec):
        """this rank's own single-scan answer, for the cross-check: {first-row key columns: (count, sums)}"""
        r = table.execute(pc.build(spec))
        return {tuple(g["out"]): (g["count"], g["first_offset"], list(g["sum"])) for g in r["groups"]}

    def merged_truth(parts):
        m = {}
        for d in parts:
            for k, (c, fo, sums) in d.items():
                if k in m:
                    m[k] = (m[k][0] + c, min(m[k][1], fo), [a + b for a, b in zip(m[k][2], sums)])
                else:
                    m[k] = (c, fo, sums)
        return sorted(m.values(), key=lambda v: v[1])

    def run_gather(table, spec, nbytes_total, label):
        plan = pc.build(no_out(spec))
        ex = GatherExchange(lib, dist)
        for _ in range(args.warmup):
            last = ex.step(table, plan)
        sampler = ClockSampler(local)
        barrier()
        sampler.start()
        l0 = lib.total_kernel_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        kms = []
        e0.record()
        for _ in range(args.steps):
            last = ex.step(table, plan)
            kms.append(last["kernel_ms"])
        e1.record()
        barrier()
        clocks = sampler.stop()
        launches = lib.total_kernel_launches() - l0
        ms = allmax(e0.elapsed_time(e1)) / args.steps
        kernel_ms = allmax(sum(kms) / len(kms))
        # results: the device-merged groups against the ranks' own answers merged on the host
        parts = [None] * world
        dist.all_gather_object(parts, truth_counts(table, spec))
        want = merged_truth(parts)
        assert last["n_groups"] == len(want), (last["n_groups"], len(want))
        for g, (c, fo, sums) in enumerate(want):
            assert last["count"][g] == c and last["first_offset"][g] == fo, (label, g, last["count"][g], c)
            for a, s in enumerate(sums):
                assert abs(last["sum"][a][g] - s) <= 1e-12 * max(abs(s), 1.0), (label, g, a, last["sum"][a][g], s)
        rows = allsum(int(last["rows_scanned"]))
        return {"value": nbytes_total / (ms / 1e3) / 1e9, "ms_per_step": ms, "kernel_ms_max_rank": kernel_ms, "groups": len(want),
                "rows_per_s": rows / (ms / 1e3), "result_checked": "groups, first offsets, counts exact and sums within 1e-12 "
                "against the ranks' own single-scan answers merged on the host", "clocks": clocks, "launches": int(launches)}

    # ---- strong: rank r holds rows [r*R/N, (r+1)*R/N) of the one 10 GB file ----
    table, buf, nbytes = make_slice(total_rows, world, rank)
    sizes = [None] * world
    dist.all_gather_object(sizes, nbytes)
    table.set_global_offset(sum(sizes[:rank]))
    total_bytes = sum(sizes)
    strong = {}
    for leg in ("groupby", "count_where"):
        strong[leg] = run_gather(table, pc.plans()[LEGS[leg][0]], total_bytes, leg)
        strong[leg]["query"] = LEGS[leg][1]
        strong[leg]["roofline_per_gpu"] = roofline_of(nbytes, strong[leg]["kernel_ms_max_rank"], LEGS[leg][2])
    # config 3: all-to-all of the partial records by owner
    hc = None
    try:
        plan = pc.build(no_out(pc.plans()["group_high_card"]))
        ox = OwnerExchange(lib, dist)
        for _ in range(max(1, args.warmup - 1)):
            last = ox.step(table, plan)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        exch, kms, sent = [], [], 0
        e0.record()
        for _ in range(args.steps):
            last = ox.step(table, plan)
            exch.append(last["exchange_ms"])
            kms.append(last["kernel_ms"])
        e1.record()
        barrier()
        ms = allmax(e0.elapsed_time(e1)) / args.steps
        groups = allsum(int(last["n_groups"]))
        rows_in = allsum(int(last["rows_in_groups"]))
        rows = allsum(int(last["rows_scanned"]))
        assert rows_in == rows, (rows_in, rows)  # every row sits in exactly one owner's group
        if abs(args.bytes - 1e10) < 1:
            assert groups == 1835776, groups
        hc = {"query": LEGS["high_card"][1], "value": total_bytes / (ms / 1e3) / 1e9, "ms_per_step": ms, "groups": groups,
              "exchange_ms_max_rank": allmax(sum(exch) / len(exch)), "scan_kernel_ms_max_rank": allmax(sum(kms) / len(kms)),
              "records_sent_per_rank": int(last["records_sent"]), "record_bytes": int(last["record_bytes"]),
              "exchange": "all_to_all_single of the partial group records by owner = hash % world, then device merge + finish of "
                          "the owned groups on every rank",
              "result_checked": "sum of the owners' group counts == rows scanned over all ranks; distinct groups over all ranks"}
    except Exception as ex:
        hc = {"error": repr(ex)}
    del table, buf
    torch.cuda.empty_cache()

    # ---- weak: 10 GB per rank (one N x 10 GB file) ----
    weak = None
    try:
        wt, wbuf, wbytes = make_slice(total_rows * world, world, rank)
        wsizes = [None] * world
        dist.all_gather_object(wsizes, wbytes)
        wt.set_global_offset(sum(wsizes[:rank]))
        weak = run_gather(wt, pc.plans()["group_name"], sum(wsizes), "weak groupby")
        weak["total_bytes"] = sum(wsizes)
        weak["note"] = "weak scaling: every rank scans its own 10 GB slice of one N x 10 GB file"
        del wt, wbuf
    except Exception as ex:
        weak = {"error": repr(ex)}

    # ---- end to end at N GPUs: every rank stages its slice from page-locked host memory and scans it ----
    e2e = None
    try:
        table, buf, nbytes = make_slice(total_rows, world, rank)
        table.set_global_offset(sum(sizes[:rank]))
        pinned = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        pinned.copy_(buf[:nbytes])
        del table, buf
        torch.cuda.empty_cache()
        from cq_b200.engine import csv_config
        plan = pc.build(no_out(pc.plans()["group_name"]))
        ex = GatherExchange(lib, dist)
        n_e2e = max(2, min(args.steps, 3))

        def e2e_step():
            with Table.from_bytes((pinned.data_ptr(), nbytes), cfg=csv_config(has_header=(rank == 0)), lib=lib, pinned=True) as t:
                t.set_global_offset(sum(sizes[:rank]))
                return ex.step(t, plan)

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            r = e2e_step()
        barrier()
        dt = allmax((time.perf_counter() - t0) / n_e2e)
        assert r["n_groups"] == strong["groupby"]["groups"], (r["n_groups"], strong["groupby"]["groups"])
        e2e = {"value": total_bytes / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": 24 * 16 * 4 + 64,
               "ms_per_step": dt * 1e3, "steps": n_e2e,
               "note": "per rank and step: cqg_table_open_buffer(page-locked host slice) + cqg_execute_partial + NCCL all_gather + "
                       "merge + result on host + close"}
        del pinned
    except Exception as ex:
        e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "error": repr(ex)}

    if rank == 0:
        head = strong["groupby"]
        line = {
            "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "bytes_per_gpu": nbytes, "total_bytes": total_bytes,
                       "l2": "no flush: every rank's slice is far larger than the 126 MB L2",
                       "parallelism": f"byte-range slices x{world} of one file, NCCL all_gather of fixed-capacity partial "
                                      "group records, device merge", "groups": head["groups"],
                       "result_checked": head["result_checked"]},
            "rows_per_s": head["rows_per_s"], "e2e": e2e, "gpu_launches": head["launches"], "clocks": head["clocks"],
            "roofline": strong["groupby"]["roofline_per_gpu"], "cpu_baseline": None,
            "configs": {"configs[0]-shape groupby (headline)": strong["groupby"], "configs[1] count_where": strong["count_where"],
                        "configs[2] high_card": hc},
            "weak": weak,
        }
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
