#!/usr/bin/env python
"""bench.py — CSV GB/s of cq's scan + filter + aggregate hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--bytes B]

One step = one pass of the hot path over one batch of synthetic CSV (the seeded restatement of
the reference's utils/generate_big_dataset.py). Workload at N=1: BASELINE.json configs[1],
`SELECT COUNT(*) FROM f WHERE age > 40` over a 10^10-byte file (≫ L2, so no flush needed).
N>1 (torchrun, one rank per GPU): weak scaling — every rank scans its own 10^10-byte slice of
one N×10 GB file, the partial aggregates are exchanged with NCCL (all_gather of fixed-size
records) and merged on the device; value = all bytes / max-over-ranks time.

Prints ONE JSON line (rank 0). Keys beyond the base contract: roofline, cpu_baseline, extra.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "csv_scan_filter_groupby_throughput"
UNIT = "GB/s"
QUERY = "SELECT COUNT(*) FROM f WHERE age > 40"
WORKLOAD = f"BASELINE configs[1]: 10 GB synthetic CSV (seeded restatement of utils/generate_big_dataset.py): {QUERY}"


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
    sql = f"SELECT COUNT(*) FROM '{path}' WHERE age > 40"
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
        r = t.execute(pc.build(pc.plans()["count_age_gt_40"]))
    dt = time.perf_counter() - t0
    return dt, "port", str(r["groups"][0]["count"])


def cpu_baseline(target_seconds=12.0):
    path, n, rows = cpu_sample_file(20e6)
    dt, kind, _ = cpu_run_once(path)
    os.unlink(path)
    rate = n / dt
    nbytes = min(max(rate * target_seconds, 20e6), 400e6)
    path, n, rows = cpu_sample_file(nbytes)
    dt, kind, _ = cpu_run_once(path)
    os.unlink(path)
    return {"value": n / dt / 1e9, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": f"first {rows} rows ({n / 1e6:.1f} MB) of the seeded file, {QUERY}, wall clock {dt:.2f} s, "
                      f"single thread (the reference has no threads)"}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    calib_path, n, rows = cpu_sample_file(20e6)
    dt, kind, _ = cpu_run_once(calib_path)
    os.unlink(calib_path)
    total = args.steps + args.warmup
    per_step = max(2.0, min(20.0, 150.0 / max(total, 1)))
    nbytes = min(max(n / dt * per_step, 20e6), 400e6)
    path, n, rows = cpu_sample_file(nbytes)
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
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
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
    from cq_b200.engine import Table, _check, gpu

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

    # ---- synthetic input, resident in HBM ----
    rows = int(args.bytes / 29.89)
    cap = lib.generate_bigdata_bound(rows, 0) + lib.device_padding()
    buf = torch.empty(cap, dtype=torch.uint8, device="cuda")
    size = C.c_size_t()
    _check(lib, lib.generate_bigdata(buf.data_ptr(), cap - lib.device_padding(), rows, 1 + rank, 0, C.byref(size)))
    nbytes = size.value
    table = Table.from_device(buf.data_ptr(), nbytes, lib=lib, keep=buf)
    if world > 1:
        sizes = [None] * world
        dist.all_gather_object(sizes, nbytes)
        table.set_global_offset(sum(sizes[:rank]))
        total_bytes = sum(sizes)
    else:
        total_bytes = nbytes
    plan = pc.build(pc.plans()["count_age_gt_40"])

    def step_single():
        return table.execute_raw(plan)

    def step_multi():
        """scan own slice -> export the partial records -> NCCL all_gather -> merge -> finish"""
        p = C.c_void_p()
        _check(lib, lib.execute_partial(table.handle, C.byref(plan.q), C.byref(p)))
        rec = lib.partial_record_size(p)
        n = lib.partial_count(p)
        counts = torch.tensor([n], device="cuda", dtype=torch.int64)
        allc = torch.empty(world, device="cuda", dtype=torch.int64)
        dist.all_gather_into_tensor(allc, counts)
        ac = allc.tolist()
        nmax = max(ac)
        send = torch.zeros(max(nmax, 1) * rec, dtype=torch.uint8, device="cuda")
        got = C.c_int64()
        _check(lib, lib.partial_export(p, 0, 1, send.data_ptr(), nmax, C.byref(got)))
        recv = torch.empty(world * max(nmax, 1) * rec, dtype=torch.uint8, device="cuda")
        dist.all_gather_into_tensor(recv, send)
        m = C.c_void_p()
        _check(lib, lib.partial_new_like(p, C.byref(m)))
        if nmax and all(a == nmax for a in ac):
            # every rank sent the same number of records: the gathered buffer is dense, one merge launch
            _check(lib, lib.partial_merge(m, recv.data_ptr(), nmax * world))
        else:
            for r in range(world):
                if ac[r]:
                    _check(lib, lib.partial_merge(m, recv.data_ptr() + r * max(nmax, 1) * rec, ac[r]))
        res = C.POINTER(A.Result)()
        _check(lib, lib.partial_finish(m, table.handle, C.byref(res)))
        out = {"count0": res.contents.count[0] if res.contents.n_groups else 0, "kernel_ms": lib.partial_kernel_ms(p),
               "rows_scanned": lib.partial_rows_scanned(p)}
        lib.result_free(res)
        lib.partial_free(m)
        lib.partial_free(p)
        return out

    step = step_multi if world > 1 else step_single

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        last = step()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    launches0 = lib.total_kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    e0.record()
    for _ in range(args.steps):
        last = step()
        kernel_ms.append(last["kernel_ms"])
    e1.record()
    barrier()
    clocks = sampler.stop()
    launches = lib.total_kernel_launches() - launches0
    ms_total = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        k = torch.tensor([sum(kernel_ms) / len(kernel_ms)], device="cuda", dtype=torch.float64)
        dist.all_reduce(k, op=dist.ReduceOp.MAX)
        kernel_avg = float(k.item())
    else:
        kernel_avg = sum(kernel_ms) / len(kernel_ms)
    ms_per_step = ms_total / args.steps
    value = total_bytes / (ms_per_step / 1e3) / 1e9

    # ---- end to end: host buffer in, host result out, through the C-ABI ----
    e2e = None
    extra = {}
    cpu = None
    if True:
        pinned = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        pinned.copy_(buf[:nbytes])
        torch.cuda.synchronize()
        n_e2e = max(2, min(args.steps, 3))

        def e2e_step():
            with Table.from_bytes((pinned.data_ptr(), nbytes), lib=lib, pinned=True) as t:
                return t.execute_raw(plan)

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            r = e2e_step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n_e2e
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": total_bytes / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": 64,
               "ms_per_step": dt * 1e3, "steps": n_e2e,
               "note": "cqg_table_open_buffer(pinned host CSV) + cqg_execute + result on host + cqg_table_close per step"}
        assert r["count0"] == last["count0"] or world > 1
        del pinned

    # ---- the other BASELINE configs, device-resident, for the record (N=1 only) ----
    if world == 1 and not args.no_extra:
        for name in ["group_name", "scalar_aggs", "group_high_card"]:
            pl = pc.build(pc.plans()[name])
            r = table.execute_raw(pl)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = table.execute_raw(pl)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            extra[name] = {"scan_kernel_gbs": nbytes / (r["kernel_ms"] / 1e3) / 1e9, "query_gbs": nbytes / dt / 1e9,
                           "groups": r["n_groups"], "kernel_ms": r["kernel_ms"], "query_ms": dt * 1e3}

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak = json.load(open(peaks_path))["hbm_gbs"]
            peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        achieved = nbytes / (kernel_avg / 1e3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "scan_kernel_traffic.json")
        if os.path.exists(tpath):
            try:
                tj = json.load(open(tpath))
                traffic = tj["dram_bytes_per_input_byte"] * nbytes
            except Exception:
                traffic = None
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "kernel": "cqg::lean2_kernel (scalar lean plans; cqg::scan_kernel for the tiles and rows it hands over)", "kernel_ms": kernel_avg,
                    "algorithmic_bytes_per_launch": nbytes, "peak_source": peak_src,
                    "frac_of_nominal_8TBs": achieved / 8000.0}
        if world == 1:
            try:
                cpu = cpu_baseline()
            except Exception as ex:  # the baseline must never sink the bench line
                cpu = {"value": None, "unit": UNIT, "cores": 1, "kind": "unavailable", "sample": repr(ex)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "bytes_per_gpu": nbytes, "rows_per_gpu": int(last["rows_scanned"]), "total_bytes": total_bytes,
                       "l2": "no flush: the 10 GB input is far larger than the 126 MB L2",
                       "parallelism": f"byte-range shards x{world}" + (", NCCL all_gather of partial aggregates" if world > 1 else ""),
                       "result_count": int(last["count0"])},
            "rows_per_s": int(last["rows_scanned"]) * world / (ms_per_step / 1e3),
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "extra": extra,
        }
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
