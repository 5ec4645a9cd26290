#!/usr/bin/env python
"""bench.py — headline benchmark of the DO hot path (BASELINE.json metric, config 2).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU aligner, same metric

Workload (BASELINE.json configs[1], SURVEY 8d "Config 2"): 2^20 pairs of 1 kb uniform-random DNA strings
(PCG64 seed 20260101, pair k = strings 2k, 2k+1), L1-norm TCM, the reference's `align2d` band
("C-linear" dialect).  One "step" = one pass of the hot path (fill + traceback + medians) over the batch.

metric = GCUPS = DP cells the REFERENCE algorithm fills for these pairs (its band; 189 299 cells per
1 kb pair) / time.  `value`: kernel throughput with inputs resident in HBM (CUDA events on the launching
stream).  `e2e`: the same through the C-ABI host entry point with pinned HOST buffers, H2D/D2H inside the
timed region.  Under torchrun every rank runs its own 2^20 pairs (weak scaling, no data-path collective).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

L = 1000
SEED = 20260101
ALG_OPS_PER_CELL = 8          # 3 adds + 2 min + 2 compare/select + 1 table fetch


def l1_tcm(a=5):
    i = np.arange(a)
    return np.abs(np.subtract.outer(i, i)).astype(np.uint32)       # bench/strings/dna/dna-L1-norm.tcm


def make_strings(n_pairs, rank=0):
    rng = np.random.Generator(np.random.PCG64(SEED + rank))
    x = rng.integers(0, 4, size=(2 * n_pairs, L), dtype=np.uint8)
    return np.left_shift(np.uint8(1), x)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power)}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cells_per_pair():
    from pcg_b200.pairwise import PackedPairs  # noqa: F401  (geometry is the library's, checked vs oracle in tests)
    # closed form of the reference's band for lg = sh = 1001 (verified against the oracle in tests)
    lg = sh = L + 1
    lim = int(.1 * lg); delta = lim // 2
    W, H = 50 + delta, 50 + delta
    c = 0
    for i in range(lg):
        lo = max(0, i - (H - 1)); hi = min(sh - 1, i + W - 1)
        c += hi - lo + 1
    return c


def cpu_arm(n_sample_hint, threads, kind_pref="reference", seconds_target=12.0):
    """The reference's own CPU aligner (oracle/_ref) on all host threads over a bounded sample."""
    from oracle import pyoracle as po
    from pcg_b200.pairwise import PackedPairs
    kind = kind_pref if (kind_pref == "port" or os.path.exists(po.REF_SO)) else "port"
    sig = l1_tcm()
    probe = PackedPairs.uniform(make_strings(max(threads, 8)))
    sec, _, cells = po.cpu_bench_align2d(sig, probe, threads, kind)
    rate = cells.sum() / max(sec, 1e-6)
    n = int(max(threads * 4, min(n_sample_hint, seconds_target * rate / cells[0])))
    pp = PackedPairs.uniform(make_strings(n))
    sec, cost, cells = po.cpu_bench_align2d(sig, pp, threads, kind)
    return {"gcups": float(cells.sum() / sec / 1e9), "seconds": sec, "pairs": n, "kind": kind,
            "cores": threads, "cost_checksum": int(cost.astype(np.int64).sum())}


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = host_threads()
    vals = []
    info = None
    for _ in range(args.warmup and 1):
        cpu_arm(64, threads, seconds_target=1.0)
    t0 = time.time()
    for _ in range(args.steps):
        info = cpu_arm(1 << 20, threads, seconds_target=max(2.0, min(20.0, 60.0 / args.steps)))
        vals.append(info["gcups"])
    v = float(np.mean(vals))
    ms = 1e3 * (time.time() - t0) / args.steps
    line = {"impl": "reference", "metric": "2D DO alignment GCUPS", "value": v, "unit": "GCUPS",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic",
            "config": {"workload": "config2: 1 kb uniform DNA pairs, L1 TCM, align2d band (C-linear)",
                       "pairs_per_step": info["pairs"], "len": L},
            "cpu_baseline": {"value": v, "unit": "GCUPS", "cores": info["cores"], "kind": info["kind"],
                             "sample": f"{info['pairs']} pairs per step (first pairs of the seeded workload), "
                                       f"all {info['cores']} host threads, per-call malloc/free as the Haskell binding"},
            "e2e": {"value": v, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=1 << 20, help="pairs per GPU per step")
    ap.add_argument("--warps", type=int, default=32, help="warps per CTA of the narrow kernel")
    ap.add_argument("--ppw", type=int, default=8, help="pairs a warp fills before tracing them")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import pcg_b200 as P
    from pcg_b200.backend import Batch
    import ctypes as C

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — this path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n = args.pairs
    ctx = P.Context(local_rank)
    ctx.configure(pairs_per_warp=args.ppw, warps_per_cta=args.warps, max_len=1024, max_b=8)
    tcm = ctx.tcm(l1_tcm())
    strings = make_strings(n, rank)
    pp = P.PackedPairs.uniform(strings)
    cpp = cells_per_pair()
    total_cells = float(cpp) * n

    dev = torch.device("cuda", local_rank)
    t = lambda a: torch.from_numpy(a.view(np.uint8) if a.dtype == np.uint8 else a).to(dev)  # noqa: E731
    d_elems = t(pp.elems)
    d_off1, d_off2 = t(pp.off1.view(np.int64)), t(pp.off2.view(np.int64))
    d_len1, d_len2 = t(pp.len1.view(np.int32)), t(pp.len2.view(np.int32))
    d_out_off = t(pp.out_off.view(np.int64))
    d_g = torch.empty(pp.out_bytes, dtype=torch.uint8, device=dev)
    d_s1, d_s2 = torch.empty_like(d_g), torch.empty_like(d_g)
    d_olen = torch.empty(n, dtype=torch.int32, device=dev)
    d_cost = torch.empty(n, dtype=torch.int32, device=dev)
    d_cells = torch.empty(n, dtype=torch.int64, device=dev)
    b = Batch(n, d_elems.data_ptr(), pp.elems.nbytes, d_off1.data_ptr(), d_off2.data_ptr(), d_len1.data_ptr(),
              d_len2.data_ptr(), d_out_off.data_ptr(), pp.out_bytes, d_g.data_ptr(), d_s1.data_ptr(),
              d_s2.data_ptr(), d_olen.data_ptr(), d_cost.data_ptr(), d_cells.data_ptr())
    stream = torch.cuda.current_stream()

    def step():
        ctx.check(ctx.lib.pcgdo_align2d_batch_device(ctx.h, tcm.h, C.byref(b), C.c_void_p(stream.cuda_stream)),
                  "pcgdo_align2d_batch_device")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    assert int(d_cells[0].item()) == cpp, (int(d_cells[0].item()), cpp)
    cost_checksum = int(d_cost.to(torch.int64).sum().item())

    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    e0.record(stream)
    for _ in range(args.steps):
        step()
        kernel_ms.append(ctx.last_kernel_ms()[0])
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    if world > 1:
        tt = torch.tensor([ms_total], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt.item())
    ms_step = ms_total / args.steps
    value = total_cells * world / (ms_step * 1e-3) / 1e9
    k_ms = float(np.mean(kernel_ms))

    # ---- end to end through the host entry point: pinned host buffers, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        try:
            pin = lambda a: torch.from_numpy(a).pin_memory().numpy()   # noqa: E731
            h_elems = pin(pp.elems)
            h_g = torch.empty(pp.out_bytes, dtype=torch.uint8).pin_memory().numpy()
            h_s1 = torch.empty(pp.out_bytes, dtype=torch.uint8).pin_memory().numpy()
            h_s2 = torch.empty(pp.out_bytes, dtype=torch.uint8).pin_memory().numpy()
            h_cost = torch.empty(n, dtype=torch.int32).pin_memory().numpy()
            h_olen = torch.empty(n, dtype=torch.int32).pin_memory().numpy()
            hb = Batch(n, h_elems.ctypes.data, pp.elems.nbytes, pp.off1.ctypes.data, pp.off2.ctypes.data,
                       pp.len1.ctypes.data, pp.len2.ctypes.data, pp.out_off.ctypes.data, pp.out_bytes,
                       h_g.ctypes.data, h_s1.ctypes.data, h_s2.ctypes.data, h_olen.ctypes.data, h_cost.ctypes.data, None)

            def estep():
                ctx.check(ctx.lib.pcgdo_align2d_batch_host(ctx.h, tcm.h, C.byref(hb)), "pcgdo_align2d_batch_host")
            estep()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                estep()
            barrier()
            dt = (time.perf_counter() - t0) / args.steps
            if world > 1:
                tt = torch.tensor([dt], device=dev)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                dt = float(tt.item())
            assert int(h_cost.astype(np.int64).sum()) == cost_checksum
            h2d = pp.elems.nbytes + 2 * 8 * n + 2 * 4 * n + 8 * n
            d2h = 3 * pp.out_bytes + 4 * n + 4 * n
            e2e = {"value": total_cells * world / dt / 1e9, "unit": "GCUPS", "h2d_bytes_per_step": int(h2d),
                   "d2h_bytes_per_step": int(d2h), "ms_per_step": dt * 1e3}
            del h_elems, h_g, h_s1, h_s2
        except Exception as ex:   # pinned allocation can fail on small hosts: report, do not fake
            e2e = {"value": None, "unit": "GCUPS", "error": str(ex)[:200]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant (only) kernel: integer-issue bound, see DESIGN.md
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    int_peak = max(ctx.int32_peak(0), ctx.int32_peak(1)) / 1e12          # T lane-instr/s, measured now
    achieved_tops = total_cells * ALG_OPS_PER_CELL / (k_ms * 1e-3) / 1e12
    alg_bytes = n * (2 * L + 3 * 2 * L + 8) + total_cells / 4.0           # in + out + 2-bit directions spilled
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("do_linear_kernel_bytes_per_launch_at_bench_size")
        except Exception:
            pass
    roofline = {"bound": "int32-issue", "achieved": achieved_tops, "peak": int_peak, "unit": "Tinstr/s",
                "frac": achieved_tops / int_peak if int_peak else None, "traffic": traffic,
                "peak_source": "measured live (pcgdo_int32_peak: independent VIADDMNMX+IMAD chains on all SMs)",
                "alg_ops_per_cell": ALG_OPS_PER_CELL, "kernel_ms": k_ms,
                "hbm": {"alg_bytes_per_launch": alg_bytes, "achieved_gbs": alg_bytes / (k_ms * 1e-3) / 1e9,
                        "peak_gbs": hbm_peak, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                        "frac": alg_bytes / (k_ms * 1e-3) / 1e9 / hbm_peak}}

    cpu = None
    if not args.no_cpu and world == 1:
        info = cpu_arm(1 << 20, host_threads(), seconds_target=12.0)
        cpu = {"value": info["gcups"], "unit": "GCUPS", "cores": info["cores"], "kind": info["kind"],
               "sample": f"first {info['pairs']} pairs of the same seeded workload, {info['seconds']:.1f} s on "
                         f"{info['cores']} host threads (oracle/_ref align2d, per-call malloc/free as FFI.hsc)"}

    line = {"metric": "2D DO alignment GCUPS", "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": "config2: batched 2D DO, 1 kb uniform DNA pairs, L1 TCM, align2d band (C-linear)",
                       "pairs_per_gpu": n, "len": L, "cells_per_pair": cpp, "full_equiv_gcups": value * L * L / cpp,
                       "l2": "inputs (2.1 GB) and direction scratch (>4 GB) exceed the 126 MB L2",
                       "seed": SEED},
            "clocks": clocks, "e2e": e2e, "gpu_launches": args.steps, "roofline": roofline, "cpu_baseline": cpu,
            "cost_checksum": cost_checksum}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
