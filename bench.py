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

L = int(os.environ.get("BENCH_L", "1000"))      # (BENCH_L: experiments on other string lengths; the bench line is L = 1000)
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


DUAL_PIPE_PEAK_T = 36.5      # profiles/r01_int_pipes_ubench.txt: VIADDMNMX + IMAD together, 125.6 lane-instr/clk/SM at 1965 MHz


def int_issue_peak(ctx):
    """Roofline denominator of the DP fills: integer lane-instructions/s with BOTH integer pipes issuing (one VIADDMNMX on
    the ALU pipe + one IMAD on the FMA pipe per pair) — measured live after a warm-up, and never taken below the figure the
    dedicated micro-benchmark reached on this pool (a low live number would flatter the fraction)."""
    live = max(ctx.int32_peak(0), ctx.int32_peak(1), ctx.int32_peak(2)) / 1e12
    return {"peak": max(live, DUAL_PIPE_PEAK_T), "live": live,
            "source": "max(live pcgdo_int32_peak after warm-up, 36.5 T of profiles/r01_int_pipes_ubench.txt); "
                      "nominal 148 SMs x 128 lanes x 1.965 GHz = 37.2 T"}


def source_hash():
    """Hash of the sources of the config-2 path (the dense C-linear kernels, their launch code and shared headers): ties
    profiles/traffic.json (an ncu capture of that kernel) to the build being timed."""
    import hashlib
    h = hashlib.sha1()
    d = os.path.join(ROOT, "pcg_b200", "csrc")
    for f in ("do_linear.cu", "do_pack16.cuh", "do_device.cuh", "do_api.cu", "do_ctx.h", "do_host.h"):
        h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def bind_to_gpu_numa_node(local_rank):
    """Pin this process to the CPUs of its GPU's NUMA node BEFORE anything is allocated, so that its pinned host
    buffers (first touch) and the copy engine's DMA stay on the GPU's side of the host.  With one process per GPU this
    is what keeps N ranks from funnelling their 5.6 GB per step through one socket.  Returns what it did (for the line)."""
    try:
        import subprocess as sp
        bus = sp.run(["nvidia-smi", "-i", str(local_rank), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                     capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return {"gpu": local_rank, "pci": bus, "numa_node": node, "bound": False}
        txt = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
        cpus = set()
        for part in txt.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        use = (cpus & allowed) or allowed
        os.sched_setaffinity(0, use)
        return {"gpu": local_rank, "pci": bus, "numa_node": node, "cpus": len(use), "bound": bool(cpus & allowed)}
    except Exception as ex:
        return {"gpu": local_rank, "bound": False, "error": str(ex)[:120]}


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
    probe = PackedPairs.uniform(make_strings(max(64 * threads, 512)))      # ~0.3 s: warm threads, steady rate
    sec, _, cells = po.cpu_bench_align2d(sig, probe, threads, kind)
    rate = cells.sum() / max(sec, 1e-6)
    n = int(max(threads * 4, min(n_sample_hint, seconds_target * rate / cells[0])))
    pp = PackedPairs.uniform(make_strings(n))
    sec, cost, cells = po.cpu_bench_align2d(sig, pp, threads, kind)
    return {"gcups": float(cells.sum() / sec / 1e9), "seconds": sec, "pairs": n, "kind": kind,
            "cores": threads, "cost_checksum": int(cost.astype(np.int64).sum()), "cost": cost}



# ------------------------------------------------------------------------------------------------
# Config 5 (BASELINE.json configs[4]): post-order scoring of 1024 random 512-taxon binary trees x 64
# dynamic DNA characters (length 256, discrete TCM, C-linear dialect).  One "evaluation" = the post-order
# cost of one tree over all 64 characters (511 x 64 alignments).  Under torchrun the characters are
# split across ranks (strong scaling) and the per-tree sums are all-reduced over NCCL.
# ------------------------------------------------------------------------------------------------
PO = {"n_trees": 1024, "n_taxa": 512, "n_chars": 64, "leaf_len": 256, "node_cap": 1024,
      "tree_seed": 20260104, "leaf_seed": 20260105}


def discrete_tcm(a=5):
    return (1 - np.eye(a, dtype=np.uint32)).astype(np.uint32)      # bench/strings/dna/dna-discrete.tcm


def po_inputs(n_trees, n_taxa, n_chars, leaf_len):
    from pcg_b200.postorder import random_binary_trees
    children = random_binary_trees(n_trees, n_taxa, PO["tree_seed"])
    rng = np.random.Generator(np.random.PCG64(PO["leaf_seed"]))
    x = rng.integers(0, 4, size=(n_taxa, n_chars, leaf_len), dtype=np.uint8)
    return children, np.left_shift(np.uint8(1), x)


def po_cpu_arm(children, leaves, threads, seconds_target=10.0, kind_pref="reference"):
    """Reference align2d inside a CPU post-order loop, threads over (tree, character) columns, bounded sample."""
    from oracle import pyoracle as po
    kind = kind_pref if (kind_pref == "port" or os.path.exists(po.REF_SO)) else "port"
    n_taxa, n_chars, L_ = leaves.shape
    flat = np.ascontiguousarray(leaves.reshape(-1))
    off = np.arange(n_taxa * n_chars, dtype=np.uint64) * np.uint64(L_)
    ln = np.full(n_taxa * n_chars, L_, np.uint32)
    sig = discrete_tcm()
    sec, _, _ = po.cpu_bench_postorder(sig, children, n_taxa, n_chars, flat, off, ln, 0, threads, threads, kind)
    n_cols = int(max(threads, min(children.shape[0] * n_chars, threads * seconds_target / max(sec, 1e-3))))
    sec, cost, cells = po.cpu_bench_postorder(sig, children, n_taxa, n_chars, flat, off, ln, 0, n_cols, threads, kind)
    return {"tree_evals_per_s": n_cols / n_chars / sec, "gcups": float(cells.sum()) / sec / 1e9, "seconds": sec,
            "columns": n_cols, "kind": kind, "cores": threads, "col_cost": cost}


def run_postorder(args, rank, world, local_rank, ctx, torch, dist, full=True):
    """Returns the dict reported under "postorder" (rank 0) — value = whole-job tree evaluations / s."""
    import pcg_b200 as P
    n_trees, n_taxa, n_chars, leaf_len = (PO["n_trees"], PO["n_taxa"], PO["n_chars"], PO["leaf_len"])
    if args.po_trees:
        n_trees = args.po_trees
    children, leaves = po_inputs(n_trees, n_taxa, n_chars, leaf_len)
    from pcg_b200.sharding import character_block
    c_lo, c_hi = character_block(n_chars, rank, world)
    cpr = c_hi - c_lo
    my = leaves[:, c_lo:c_hi, :]
    ctx.configure(pairs_per_warp=args.ppw, warps_per_cta=args.warps, max_len=PO["node_cap"], max_b=32)
    tcm = ctx.tcm(discrete_tcm())
    forest = P.Forest(ctx, children, n_taxa, [[my[v, c] for c in range(cpr)] for v in range(n_taxa)], PO["node_cap"])
    dev = torch.device("cuda", local_rank)
    costs = torch.zeros(n_trees, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream()

    if args.reroot:
        forest.enable_rerooting()

    def step():
        if args.reroot:
            forest.score_rerooted_into(tcm, costs.data_ptr(), stream.cuda_stream)
        else:
            forest.score_into(tcm, costs.data_ptr(), stream.cuda_stream)
        if world > 1:
            dist.all_reduce(costs)          # the path's only collective: 8 KB of per-tree sums

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(1, min(args.warmup, 3))):
        step()
    barrier()
    checksum = int(costs.sum().item())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = max(1, args.po_steps)
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        tt = torch.tensor([ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    n_align = n_trees * n_chars * ((n_taxa - 1) + ((4 * n_taxa - 8) if args.reroot else 0))
    launches = ctx.last_kernel_ms()[1]
    # ---- roofline of the sweep: the same integer-issue bound as config 2, over the band cells of all alignments
    roofline = None
    if not args.reroot:
        cells_rank = float(forest.cells())
        cells_all = cells_rank
        if world > 1:
            tt = torch.tensor([cells_rank], device=dev, dtype=torch.float64)
            dist.all_reduce(tt)
            cells_all = float(tt.item())
        pk = int_issue_peak(ctx)
        ach = cells_rank * ALG_OPS_PER_CELL / (ms * 1e-3) / 1e12
        roofline = {"bound": "int32-issue", "achieved": ach, "peak": pk["peak"], "unit": "Tinstr/s", "frac": ach / pk["peak"],
                    "peak_live": pk["live"], "peak_source": pk["source"], "alg_ops_per_cell": ALG_OPS_PER_CELL,
                    "cells_per_sweep": cells_all, "gcups": cells_all / (ms * 1e-3) / 1e9, "traffic": None,
                    "note": "whole sweep (all levels: descriptor prep + DO kernels + totals) timed as one region; per-level "
                            "shares in profiles/"}
    # ---- candidate-tree turnover: a FRESH set of topologies is uploaded and levelled (on the device) before every
    # sweep, as a search does; alternates between two sets so that no sweep re-scores resident trees
    fresh = None
    if not args.reroot and not args.no_fresh_trees:
        from pcg_b200.postorder import random_binary_trees
        alt = [children, random_binary_trees(n_trees, n_taxa, PO["tree_seed"] + 1)]

        def fstep(i):
            forest.set_trees(alt[i & 1])
            step()
        fstep(1); fstep(0)
        barrier()
        t0 = time.perf_counter()
        for i in range(steps):
            fstep(i + 1)
        barrier()
        fms = (time.perf_counter() - t0) / steps * 1e3
        if world > 1:
            tt = torch.tensor([fms], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            fms = float(tt.item())
        forest.set_trees(children)
        step(); barrier()
        assert int(costs.sum().item()) == checksum, "scores changed after the trees were swapped out and back in"
        fresh = {"value": n_trees / (fms * 1e-3), "unit": "tree-evals/s", "ms_per_sweep": fms,
                 "how": "pcgdo_forest_set_trees (upload 4 MB of child ids, validate + level on the device) + sweep, host wall "
                        "clock, every sweep on topologies the forest has not scored in the previous sweep"}
    out = {"metric": "postorder tree cost evals/sec" + (" (with re-rooting)" if args.reroot else ""),
           "value": n_trees / (ms * 1e-3), "unit": "tree-evals/s",
           "ms_per_sweep": ms, "steps": steps, "scaling": "strong", "levels": forest.levels,
           "roofline": roofline, "fresh_trees": fresh,
           "alignments_per_sweep": n_align,
           "config": {"workload": "config5: post-order of random binary trees x dynamic DNA characters, discrete TCM",
                      "n_trees": n_trees, "n_taxa": n_taxa, "n_chars": n_chars, "leaf_len": leaf_len,
                      "chars_per_gpu": cpr, "collective": "1 all_reduce(int64[n_trees]) per sweep" if world > 1 else "none"},
           "cost_checksum": checksum, "gpu_launches_per_sweep": launches}
    if rank == 0 and world == 1 and not args.no_cpu and not args.reroot:
        info = po_cpu_arm(children, leaves, host_threads(), seconds_target=8.0)
        # parity spot check of the same columns against the device totals
        tr, ch = np.divmod(np.arange(info["columns"]), n_chars)
        dev_tot = np.zeros(info["columns"], np.int64)
        roots = {}
        for k in range(min(info["columns"], 64)):
            t, c = int(tr[k]), int(ch[k])
            if t not in roots:
                is_child = set(int(x) for x in children[t].reshape(-1))
                roots[t] = [n_taxa + x for x in range(n_taxa - 1) if n_taxa + x not in is_child][0]
            dev_tot[k] = forest.node(t, c, roots[t])[2]
            assert dev_tot[k] == int(info["col_cost"][k]), (t, c)
        out["cpu_baseline"] = {"value": info["tree_evals_per_s"], "unit": "tree-evals/s", "cores": info["cores"],
                               "kind": info["kind"], "gcups": info["gcups"],
                               "sample": f"first {info['columns']} (tree, character) columns, {info['seconds']:.1f} s on "
                                         f"{info['cores']} host threads; root totals of the first 64 columns checked "
                                         f"equal to the device's"}
    forest.close()
    del forest
    torch.cuda.empty_cache()
    # ---- the same sweep through the library's own multi-GPU path (pcgdo_multi: one process, all N devices, forests sharded
    # by character blocks, the three reductions).  Under torchrun rank 0 alone drives it while the other ranks — whose
    # forests are closed — wait on a CPU barrier (an NCCL barrier would spin on their GPUs).
    if world > 1 and not args.reroot and not args.no_library_multi:
        g = getattr(run_postorder, "_gloo", None)
        if g is None:
            g = dist.new_group(backend="gloo")
            run_postorder._gloo = g
        torch.cuda.synchronize()
        dist.barrier(group=g)
        if rank == 0:
            try:
                out["library_multi_gpu"] = library_multi_sweep(world, children, leaves, n_trees, n_taxa, args, checksum)
            except Exception as ex:
                out["library_multi_gpu"] = {"value": None, "error": str(ex)[:300]}
        dist.barrier(group=g)
    return out


def library_multi_sweep(n_dev, children, leaves, n_trees, n_taxa, args, checksum):
    import pcg_b200 as P
    from pcg_b200.backend import REDUCE_HOST, REDUCE_PEER, REDUCE_NCCL
    m = P.Multi(list(range(n_dev)))
    m.configure(args.ppw, args.warps, PO["node_cap"], 32)
    mt = m.tcm(discrete_tcm())
    mf = m.forest(children, n_taxa, leaves, PO["node_cap"])
    res = {}
    for name, mode in (("host_add", REDUCE_HOST), ("peer_kernel", REDUCE_PEER), ("nccl_allreduce", REDUCE_NCCL)):
        try:
            m.score(mt, mf, reduce_mode=mode)
            walls, sweeps, reds = [], [], []
            for _ in range(max(2, args.po_steps)):
                cost, wall, sweep, red = m.score(mt, mf, reduce_mode=mode)
                walls.append(wall); sweeps.append(sweep); reds.append(red)
            assert int(cost.sum()) == checksum, "library multi-GPU costs differ from the torch.distributed run"
            res[name] = {"ms_per_sweep_wall": float(np.mean(walls)), "slowest_device_sweep_ms": float(np.mean(sweeps)),
                         "reduce_ms": float(np.mean(reds)), "tree_evals_per_s": n_trees / (float(np.mean(walls)) * 1e-3)}
        except Exception as ex:
            res[name] = {"error": str(ex)[:200]}
    m.destroy_forest(mf)
    m.close()
    ok = [v for v in res.values() if "tree_evals_per_s" in v]
    return {"value": max(v["tree_evals_per_s"] for v in ok) if ok else None, "unit": "tree-evals/s", "devices": n_dev,
            "how": "ONE process: pcgdo_mforest_score over all devices (one host thread per device inside the library, "
                   "characters sharded), host wall clock per sweep including the reduction of int64[n_trees]",
            "reductions": res, "cost_checksum": checksum}


# ------------------------------------------------------------------------------------------------
# Config 3 (BASELINE.json configs[2]): affine-gap DO on 100 000 pairs of ~5 kb strings with IUPAC ambiguity
# codes (lengths U[4750, 5250], 10 % of the elements a uniform non-empty subset of {A,C,G,T}), discrete TCM,
# gap-open 3, C-affine dialect.  The reference's affine substitution lookup is undefined behaviour
# (alignCharacters.c:3636); the bench runs the PATCHED reading (`<< alphSize`) on both arms — the CPU arm is
# oracle/_ref/libpcgdo_ref_patched.so, the same sources with that one expression corrected.
# ------------------------------------------------------------------------------------------------
def config3_pairs(n_pairs, rank=0):
    from pcg_b200.pairwise import PackedPairs
    rng = np.random.Generator(np.random.PCG64(20260102 + rank))
    lens = rng.integers(4750, 5251, size=2 * n_pairs).astype(np.uint32)
    total = int(lens.astype(np.int64).sum())
    base = np.left_shift(np.uint8(1), rng.integers(0, 4, size=total, dtype=np.uint8))
    amb = rng.integers(0, 10, size=total, dtype=np.uint8) == 0
    vals = rng.integers(1, 16, size=total, dtype=np.uint8)
    elems = np.where(amb, vals, base).astype(np.uint8)
    del base, amb, vals
    starts = np.concatenate([[0], np.cumsum(lens.astype(np.uint64))]).astype(np.uint64)
    off1, off2 = starts[0:-1:2].copy(), starts[1:-1:2].copy()
    len1, len2 = lens[0::2].copy(), lens[1::2].copy()
    cap = (len1.astype(np.int64) + len2.astype(np.int64) + 3) & ~3
    ostarts = np.concatenate([[0], np.cumsum(cap)]).astype(np.uint64)
    elems = np.concatenate([elems, np.zeros(16, np.uint8)])
    return PackedPairs(n_pairs, elems, off1, off2, len1, len2, ostarts[:-1].copy(), int(ostarts[-1]) + 16)


def run_config3(args, rank, world, local_rank, ctx, torch, dist, steps=None, n=None):
    import ctypes as C
    from pcg_b200.backend import Batch
    if n is None:
        n = args.pairs if args.pairs != (1 << 20) else 100000
    steps = steps or args.steps
    warm = max(args.warmup, 3)
    GO, VARIANT = 3, 1
    pp = config3_pairs(n, rank)
    ctx.configure(pairs_per_warp=min(args.ppw, 8), warps_per_cta=16, max_len=5250, max_b=32)
    dev = torch.device("cuda", local_rank)
    t = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(tcm, k_steps):
        def step():
            ctx.check(ctx.lib.pcgdo_align2d_affine_batch_device(ctx.h, tcm.h, C.byref(b), C.c_void_p(stream.cuda_stream)),
                      "pcgdo_align2d_affine_batch_device")
        for _ in range(warm):
            step()
        cells = float(d_cells.sum().item())
        checksum = int(d_cost.to(torch.int64).sum().item())
        sampler = ClockSampler(local_rank)
        barrier()
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        kms = []
        e0.record(stream)
        for _ in range(k_steps):
            step()
            kms.append(ctx.last_kernel_ms()[0])
        e1.record(stream)
        barrier()
        clocks = sampler.stop()
        ms = e0.elapsed_time(e1) / k_steps
        cells_all = cells
        if world > 1:
            tt = torch.tensor([ms, cells], device=dev, dtype=torch.float64)
            mx = tt.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = tt.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            ms, cells_all = float(mx[0].item()), float(sm[1].item())
        return {"ms": ms, "kernel_ms": float(np.mean(kms)), "cells_rank": cells, "cells_all": cells_all,
                "gcups": cells_all / (ms * 1e-3) / 1e9, "checksum": checksum, "clocks": clocks,
                "launches": ctx.last_kernel_ms()[1]}

    tcm = ctx.tcm(discrete_tcm(), gap_open=GO, affine_variant=VARIANT)
    rd = timed(tcm, steps)
    cost_d = d_cost.cpu().numpy().copy()
    olen_d = d_olen.cpu().numpy().copy()
    # BASELINE.md §2 runs gap-open 3 with the L1-norm matrix as well (app/affine-safety.hs:71-87): second arm, fewer steps
    tcm_l1 = ctx.tcm(l1_tcm(), gap_open=GO, affine_variant=VARIANT)
    rl = timed(tcm_l1, max(3, steps // 3))

    # ---- end to end: pinned host buffers through pcgdo_align2d_affine_batch_host (chunked copy / align / copy pipeline)
    e2e = None
    if not args.no_e2e:
        try:
            pin = lambda a: torch.from_numpy(a).pin_memory().numpy()   # noqa: E731
            h_elems = pin(pp.elems)
            h_g, h_s1, h_s2 = (torch.empty(pp.out_bytes, dtype=torch.uint8).pin_memory().numpy() for _ in range(3))
            h_cost, h_olen = (torch.empty(n, dtype=torch.int32).pin_memory().numpy() for _ in range(2))
            h_off1, h_off2, h_len1, h_len2, h_ooff = (pin(a) for a in (pp.off1, pp.off2, pp.len1, pp.len2, pp.out_off))
            hb = Batch(n, h_elems.ctypes.data, pp.elems.nbytes, h_off1.ctypes.data, h_off2.ctypes.data, h_len1.ctypes.data,
                       h_len2.ctypes.data, h_ooff.ctypes.data, pp.out_bytes, h_g.ctypes.data, h_s1.ctypes.data,
                       h_s2.ctypes.data, h_olen.ctypes.data, h_cost.ctypes.data, None, None)

            def estep():
                ctx.check(ctx.lib.pcgdo_align2d_affine_batch_host(ctx.h, tcm.h, C.byref(hb)), "pcgdo_align2d_affine_batch_host")
            estep()
            barrier()
            esteps = max(3, steps // 2)
            t0 = time.perf_counter()
            for _ in range(esteps):
                estep()
            barrier()
            dt = (time.perf_counter() - t0) / esteps
            if world > 1:
                tt = torch.tensor([dt], device=dev)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                dt = float(tt.item())
            assert np.array_equal(h_cost, cost_d) and np.array_equal(h_olen, olen_d), "host path differs from the device-resident run"
            step_dev = lambda: ctx.check(ctx.lib.pcgdo_align2d_affine_batch_device(ctx.h, tcm.h, C.byref(b), C.c_void_p(stream.cuda_stream)), "affine")  # noqa: E731
            step_dev(); torch.cuda.synchronize()
            for k in np.linspace(0, n - 1, 16).astype(np.int64):
                o, ln = int(pp.out_off[k]), int(h_olen[k])
                for hbuf, dbuf in ((h_g, d_g), (h_s1, d_s1), (h_s2, d_s2)):
                    assert np.array_equal(hbuf[o:o + ln], dbuf[o:o + ln].cpu().numpy()), int(k)
            e2e = {"value": rd["cells_all"] / dt / 1e9, "unit": "GCUPS", "ms_per_step": dt * 1e3, "steps": esteps,
                   "h2d_bytes_per_step": int(pp.elems.nbytes + 40 * n),
                   "d2h_bytes_per_step": int(3 * int(pp.out_off[-1] + ((int(pp.len1[-1]) + int(pp.len2[-1]) + 3) & ~3)) + 8 * n)}
            del h_elems, h_g, h_s1, h_s2
        except Exception as ex:
            e2e = {"value": None, "unit": "GCUPS", "error": str(ex)[:200]}
    if rank != 0:
        return None
    pk = int_issue_peak(ctx)
    ach = rd["cells_rank"] * 24 / (rd["kernel_ms"] * 1e-3) / 1e12      # 24 integer ops per affine cell
    cpu = None
    if not args.no_cpu and world == 1:
        from oracle import pyoracle as po
        import types
        threads = host_threads()
        ns = max(threads, min(n, 4 * threads))
        sub = types.SimpleNamespace(n=ns, elems=pp.elems, off1=pp.off1[:ns].copy(), off2=pp.off2[:ns].copy(),
                                    len1=pp.len1[:ns].copy(), len2=pp.len2[:ns].copy())
        kind = "reference" if os.path.exists(po.REF_PATCHED_SO) else "port"
        sec, cost, _ = po.cpu_bench_align2d(discrete_tcm(), sub, threads, kind, want_cells=False, gap_open=GO,
                                            affine_variant=VARIANT)
        cells_s = float(d_cells[:ns].sum().item())
        assert np.array_equal(cost, cost_d[:ns]), "device costs differ from the CPU arm on the sample"
        cpu = {"value": cells_s / sec / 1e9, "unit": "GCUPS", "cores": threads, "kind": kind,
               "sample": f"first {ns} pairs, {sec:.1f} s on {threads} host threads "
                         f"(align2dAffine of the one-expression-patched reference build); costs equal to the device's"}
    return {"metric": "2D DO alignment GCUPS (affine)", "value": rd["gcups"], "unit": "GCUPS", "n_gpus": world,
            "steps": steps, "warmup": warm, "ms_per_step": rd["ms"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": "config3: affine-gap DO, ~5 kb IUPAC-ambiguous DNA pairs, discrete TCM, gap-open 3, "
                                   "C-affine dialect (patched substitution lookup)",
                       "pairs_per_gpu": n, "len": "U[4750,5250]", "cells_per_pair_avg": rd["cells_rank"] / n,
                       "l2": "inputs (1 GB) and direction scratch exceed the 126 MB L2"},
            "clocks": rd["clocks"], "e2e": e2e, "gpu_launches": rd["launches"] * steps,
            "roofline": {"bound": "int32-issue", "achieved": ach, "peak": pk["peak"], "unit": "Tinstr/s",
                         "frac": ach / pk["peak"], "peak_live": pk["live"], "peak_source": pk["source"],
                         "alg_ops_per_cell": 24, "kernel_ms": rd["kernel_ms"], "traffic": None},
            "cpu_baseline": cpu, "cost_checksum": rd["checksum"],
            "l1_tcm": {"workload": "same pairs, L1-norm matrix, gap-open 3", "value": rl["gcups"], "unit": "GCUPS",
                       "ms_per_step": rl["ms"], "kernel_ms": rl["kernel_ms"], "steps": max(3, steps // 3),
                       "cost_checksum": rl["checksum"], "clocks": rl["clocks"]}}

# ------------------------------------------------------------------------------------------------
# Config 4 (BASELINE.json configs[3]): amino-acid DO, 21-state alphabet, 500 000 pairs of 400 residues
# (single-symbol elements uniform over the 20 amino acids, seed 20260103), dialect S2 Ukkonen
# (unboxedUkkonenFullSpaceDO, PCG's production path for alphabets > 8).  Primary: the discrete metric
# (bench/strings/protein/protein-discrete.tcm -> discreteMetricPairwiseLogic, closed form in registers);
# secondary: the explicit matrix protein-pref-gap.tcm (substitution 2, indel 1) through the per-pair memo
# table that replaces tcm-memo.  cells = cells of the FINAL Ukkonen band of each pair (the refills of the
# doubling schedule are not counted, on either arm).  The reference runs this path in Haskell; no GHC here,
# so the CPU arm is the restatement ("port"), for the explicit matrix with every overlap answered by the
# compiled reference tcm-memo.
# ------------------------------------------------------------------------------------------------
def config4_strings(n_pairs, rank=0, L4=400):
    rng = np.random.Generator(np.random.PCG64(20260103 + rank))
    x = rng.integers(0, 20, size=(2 * n_pairs, L4), dtype=np.uint8)
    return np.left_shift(np.uint32(1), x.astype(np.uint32))


def pref_gap_tcm(a=21):
    m = 2 * (1 - np.eye(a, dtype=np.uint32))
    m[-1, :-1] = 1
    m[:-1, -1] = 1
    return m.astype(np.uint32)                                     # bench/strings/protein/protein-pref-gap.tcm


def run_config4(args, rank, world, local_rank, ctx, torch, dist, steps=None, n=None):
    import ctypes as C
    from pcg_b200.backend import Batch32
    A, L4, DIALECT = 21, 400, 2
    if n is None:
        n = args.pairs if args.pairs != (1 << 20) else 500000
    steps = steps or args.steps
    strings = config4_strings(n, rank, L4)
    elems = np.ascontiguousarray(strings.reshape(-1))
    k = np.arange(n, dtype=np.uint64)
    off1 = k * np.uint64(2 * L4); off2 = off1 + np.uint64(L4)
    ln = np.full(n, L4, np.uint32)
    out_off = off1.copy()
    cnt = elems.size + 4
    ctx.configure(pairs_per_warp=min(args.ppw, int(os.environ.get("BENCH_C4_PPW", "16"))), warps_per_cta=16, max_len=L4, max_b=32)
    memo = ctx.memo(pref_gap_tcm(A))
    dev = torch.device("cuda", local_rank)
    t = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
    d_elems = t(elems.view(np.int32))
    d_off1, d_off2, d_oo = t(off1.view(np.int64)), t(off2.view(np.int64)), t(out_off.view(np.int64))
    d_len = t(ln.view(np.int32))
    d_m = torch.empty(cnt, dtype=torch.int32, device=dev)
    d_a1, d_a2 = torch.empty_like(d_m), torch.empty_like(d_m)
    d_olen = torch.empty(n, dtype=torch.int32, device=dev)
    d_cost = torch.empty(n, dtype=torch.int32, device=dev)
    d_cells = torch.empty(n, dtype=torch.int64, device=dev)
    d_fo = torch.empty(n, dtype=torch.int32, device=dev)
    b = Batch32(n, d_elems.data_ptr(), elems.size, d_off1.data_ptr(), d_off2.data_ptr(), d_len.data_ptr(),
                d_len.data_ptr(), d_oo.data_ptr(), cnt, d_m.data_ptr(), d_a1.data_ptr(), d_a2.data_ptr(),
                d_olen.data_ptr(), d_cost.data_ptr(), d_cells.data_ptr(), d_fo.data_ptr())
    stream = torch.cuda.current_stream()
    sp = C.c_void_p(stream.cuda_stream)

    def step_discrete():
        ctx.check(ctx.lib.pcgdo_metric_batch_device(ctx.h, A, DIALECT, C.byref(b), sp), "pcgdo_metric_batch_device")

    def step_explicit():
        ctx.check(ctx.lib.pcgdo_explicit_batch_device(ctx.h, memo.h, DIALECT, C.byref(b), sp), "pcgdo_explicit_batch_device")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step):
        for _ in range(max(args.warmup, 3)):
            step()
        cells = float(d_cells.sum().item())
        checksum = int(d_cost.to(torch.int64).sum().item())
        sampler = ClockSampler(local_rank)
        barrier()
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        kms = []
        e0.record(stream)
        for _ in range(steps):
            step()
            kms.append(ctx.last_kernel_ms()[0])
        e1.record(stream)
        barrier()
        clocks = sampler.stop()
        ms = e0.elapsed_time(e1) / steps
        if world > 1:
            tt = torch.tensor([ms, cells], device=dev, dtype=torch.float64)
            mx = tt.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = tt.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            ms, cells_all = float(mx[0].item()), float(sm[1].item())
        else:
            cells_all = cells
        return {"ms": ms, "kernel_ms": float(np.mean(kms)), "cells_rank": cells, "cells_all": cells_all,
                "gcups": cells_all / (ms * 1e-3) / 1e9, "checksum": checksum, "clocks": clocks}

    rd = timed(step_discrete)
    cost_d = d_cost[:4096].cpu().numpy().copy()
    cells_d = d_cells[:4096].cpu().numpy().copy()
    rx = timed(step_explicit)
    cost_x = d_cost[:4096].cpu().numpy().copy()
    cells_x = d_cells[:4096].cpu().numpy().copy()

    # ---- end to end (discrete): pinned host buffers through pcgdo_metric_batch_host
    e2e = None
    if not args.no_e2e:
        try:
            pin = lambda a: torch.from_numpy(a).pin_memory().numpy()   # noqa: E731
            h_elems = pin(elems)
            h_m, h_a1, h_a2 = (torch.empty(cnt, dtype=torch.int32).pin_memory().numpy() for _ in range(3))
            h_cost, h_olen = (torch.empty(n, dtype=torch.int32).pin_memory().numpy() for _ in range(2))
            hb = Batch32(n, h_elems.ctypes.data, elems.size, off1.ctypes.data, off2.ctypes.data, ln.ctypes.data,
                         ln.ctypes.data, out_off.ctypes.data, cnt, h_m.ctypes.data, h_a1.ctypes.data, h_a2.ctypes.data,
                         h_olen.ctypes.data, h_cost.ctypes.data, None, None)

            def estep():
                ctx.check(ctx.lib.pcgdo_metric_batch_host(ctx.h, A, DIALECT, C.byref(hb)), "pcgdo_metric_batch_host")
            estep()
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                estep()
            barrier()
            dt = (time.perf_counter() - t0) / steps
            if world > 1:
                tt = torch.tensor([dt], device=dev)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                dt = float(tt.item())
            assert int(h_cost.astype(np.int64).sum()) == rd["checksum"]
            e2e = {"value": rd["cells_all"] / dt / 1e9, "unit": "GCUPS", "ms_per_step": dt * 1e3,
                   "h2d_bytes_per_step": int(elems.nbytes + 3 * 8 * n + 2 * 4 * n),
                   "d2h_bytes_per_step": int(3 * 4 * cnt + 2 * 4 * n)}
        except Exception as ex:
            e2e = {"value": None, "unit": "GCUPS", "error": str(ex)[:200]}
    if rank != 0:
        return None
    pk = int_issue_peak(ctx)
    int_peak = pk["peak"]
    ach = rd["cells_rank"] * ALG_OPS_PER_CELL / (rd["kernel_ms"] * 1e-3) / 1e12
    ach_x = rx["cells_rank"] * ALG_OPS_PER_CELL / (rx["kernel_ms"] * 1e-3) / 1e12
    cpu = cpu_x = None
    if not args.no_cpu and world == 1:
        from oracle import pyoracle as po
        threads = host_threads()
        ns = max(threads, min(4096, 40 * threads))
        sec, cost, _, _ = po.cpu_bench_haskell(DIALECT, A, elems, off1[:ns].copy(), ln[:ns].copy(), off2[:ns].copy(),
                                               ln[:ns].copy(), threads)
        assert np.array_equal(cost, cost_d[:ns]), "device costs differ from the CPU arm on the sample (discrete)"
        cpu = {"value": float(cells_d[:ns].sum()) / sec / 1e9, "unit": "GCUPS", "cores": threads, "kind": "port",
               "sample": f"first {ns} pairs, {sec:.1f} s on {threads} host threads (restated unboxedUkkonenFullSpaceDO, "
                         f"discrete metric; the reference's Haskell cannot be built here); costs equal to the device's"}
        nx = max(threads, min(4096, 8 * threads))
        sec, cost, _, lk = po.cpu_bench_haskell(DIALECT, A, elems, off1[:nx].copy(), ln[:nx].copy(), off2[:nx].copy(),
                                                ln[:nx].copy(), threads, sigma=pref_gap_tcm(A),
                                                through_reference_memo=os.path.exists(po.REF_MEMO_SO))
        assert np.array_equal(cost, cost_x[:nx]), "device costs differ from the CPU arm on the sample (explicit)"
        cpu_x = {"value": float(cells_x[:nx].sum()) / sec / 1e9, "unit": "GCUPS", "cores": threads, "kind": "port",
                 "tcm_memo_lookups_per_s": lk / sec if lk else None,
                 "sample": f"first {nx} pairs, {sec:.1f} s on {threads} host threads (restated dialect; every overlap "
                           f"answered by the compiled reference tcm-memo getCostAndMedian2D, one memo per thread)"}
    return {"metric": "2D DO alignment GCUPS (amino acids, a=21)", "value": rd["gcups"], "unit": "GCUPS", "n_gpus": world,
            "steps": steps, "warmup": max(args.warmup, 3), "ms_per_step": rd["ms"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": "config4: amino-acid DO, a=21, 400-residue pairs, discrete metric, S2 Ukkonen dialect "
                                   "(unboxedUkkonenFullSpaceDO)",
                       "pairs_per_gpu": n, "len": L4, "cells_per_pair_avg": rd["cells_rank"] / n,
                       "cells": "final Ukkonen band per pair (doubling refills not counted)",
                       "l2": "inputs (1.6 GB) and direction scratch exceed the 126 MB L2"},
            "clocks": rd["clocks"], "e2e": e2e, "gpu_launches": 2 * steps,
            "roofline": {"bound": "int32-issue", "achieved": ach, "peak": int_peak, "unit": "Tinstr/s",
                         "frac": ach / int_peak if int_peak else None, "peak_live": pk["live"], "peak_source": pk["source"],
                         "alg_ops_per_cell": ALG_OPS_PER_CELL, "kernel_ms": rd["kernel_ms"], "traffic": None,
                         "cells": "final Ukkonen band only: the refills of the doubling schedule are work the reference "
                                  "does too but are not counted on either arm"},
            "cpu_baseline": cpu, "cost_checksum": rd["checksum"],
            "explicit": {"workload": "same pairs, explicit matrix protein-pref-gap.tcm (sub 2 / indel 1) through the "
                                     "per-pair memo table (tcm-memo replacement)",
                         "value": rx["gcups"], "unit": "GCUPS", "ms_per_step": rx["ms"], "kernel_ms": rx["kernel_ms"],
                         "cells_per_pair_avg": rx["cells_rank"] / n, "cost_checksum": rx["checksum"],
                         "clocks": rx["clocks"],
                         "roofline": {"bound": "int32-issue", "achieved": ach_x, "peak": int_peak, "unit": "Tinstr/s",
                                      "frac": ach_x / int_peak, "alg_ops_per_cell": ALG_OPS_PER_CELL},
                         "cpu_baseline": cpu_x}}

# ------------------------------------------------------------------------------------------------
# Config 2, secondary dialect: the same 1 kb DNA pairs and L1 matrix through `ukkonenDO`
# (DO/Pairwise/Ukkonen/Internal.hs:79-202; S1 recurrence on a ribbon that doubles until the threshold exceeds the
# cost).  Device: dialect 1 of pcgdo_explicit_batch_* over the matrix as an explicit overlap.  CPU: north_star asks for
# the reference's ukkonenDO on the host cores; it is Haskell (no GHC here), so the timed code is the restatement
# ("port").  cells = cells of each pair's FINAL ribbon, on both arms.
# ------------------------------------------------------------------------------------------------
def run_config2_ukkonen(args, rank, world, local_rank, ctx, torch, dist, steps=10, n=1 << 17):
    import ctypes as C
    from pcg_b200.backend import Batch32
    A, DIALECT = 5, 1
    # Unrelated 1 kb strings drive ukkonenDO's ribbon to the whole 1001 x 1001 matrix (2001 diagonals); the warp kernels
    # hold bands of up to 1023 diagonals, so the DEVICE runs the 500-element prefixes of the same strings (1001
    # diagonals) — stated in the line — while the CPU port is timed on both lengths.
    LU = 500
    full = make_strings(n, rank).astype(np.uint32)
    strings = np.ascontiguousarray(full[:, :LU])
    elems = np.ascontiguousarray(strings.reshape(-1))
    k = np.arange(n, dtype=np.uint64)
    off1 = k * np.uint64(2 * LU); off2 = off1 + np.uint64(LU)
    ln = np.full(n, LU, np.uint32)
    cnt = elems.size + 4
    ctx.configure(pairs_per_warp=min(args.ppw, 8), warps_per_cta=16, max_len=LU, max_b=32)
    memo = ctx.memo(l1_tcm(A), structure="explicit")
    dev = torch.device("cuda", local_rank)
    t = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
    d_elems = t(elems.view(np.int32))
    d_off1, d_off2 = t(off1.view(np.int64)), t(off2.view(np.int64))
    d_len = t(ln.view(np.int32))
    d_m = torch.empty(cnt, dtype=torch.int32, device=dev)
    d_a1, d_a2 = torch.empty_like(d_m), torch.empty_like(d_m)
    d_olen, d_cost, d_fo = (torch.empty(n, dtype=torch.int32, device=dev) for _ in range(3))
    d_cells = torch.empty(n, dtype=torch.int64, device=dev)
    b = Batch32(n, d_elems.data_ptr(), elems.size, d_off1.data_ptr(), d_off2.data_ptr(), d_len.data_ptr(), d_len.data_ptr(),
                d_off1.data_ptr(), cnt, d_m.data_ptr(), d_a1.data_ptr(), d_a2.data_ptr(), d_olen.data_ptr(),
                d_cost.data_ptr(), d_cells.data_ptr(), d_fo.data_ptr())
    stream = torch.cuda.current_stream()

    def step():
        ctx.check(ctx.lib.pcgdo_explicit_batch_device(ctx.h, memo.h, DIALECT, C.byref(b), C.c_void_p(stream.cuda_stream)),
                  "pcgdo_explicit_batch_device")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(3):
        step()
    cells = float(d_cells.sum().item())
    checksum = int(d_cost.to(torch.int64).sum().item())
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kms = []
    e0.record(stream)
    for _ in range(steps):
        step()
        kms.append(ctx.last_kernel_ms()[0])
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1) / steps
    cells_all = cells
    if world > 1:
        tt = torch.tensor([ms, cells], device=dev, dtype=torch.float64)
        mx = tt.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tt.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, cells_all = float(mx[0].item()), float(sm[1].item())
    if rank != 0:
        return None
    cpu = None
    if not args.no_cpu and world == 1:
        from oracle import pyoracle as po
        threads = host_threads()
        ns = max(threads, min(n, 8 * threads))
        sec, cost, _, _ = po.cpu_bench_haskell(DIALECT, A, elems, off1[:ns].copy(), ln[:ns].copy(), off2[:ns].copy(),
                                               ln[:ns].copy(), threads, sigma=l1_tcm(A))
        assert np.array_equal(cost, d_cost[:ns].cpu().numpy()), "device costs differ from the restated ukkonenDO on the sample"
        cpu = {"value": float(d_cells[:ns].sum().item()) / sec / 1e9, "unit": "GCUPS", "cores": threads, "kind": "port",
               "sample": f"first {ns} pairs (500-element prefixes), {sec:.1f} s on {threads} host threads: ukkonenDO RESTATED in C "
                         f"(oracle/do_oracle.c) — the reference's is Haskell and cannot be built here; costs equal to the device's"}
        # the full 1 kb pairs of config 2 through the same port (CPU only: see above)
        n1k = max(threads, 2 * threads)
        e1k = np.ascontiguousarray(full[:2 * n1k].reshape(-1))
        k1 = np.arange(n1k, dtype=np.uint64)
        o1k = k1 * np.uint64(2 * L)
        l1k = np.full(n1k, L, np.uint32)
        sec1k, _, cells1k, _ = po.cpu_bench_haskell(DIALECT, A, e1k, o1k, l1k, o1k + np.uint64(L), l1k, threads, sigma=l1_tcm(A))
        # the port counts the refills of the doubling schedule; the final ribbon of unrelated strings is the whole matrix
        cpu["config2_1kb_pairs"] = {"value": float(n1k * (L + 1) * (L + 1)) / sec1k / 1e9, "unit": "GCUPS", "pairs": n1k,
                                    "seconds": sec1k, "cells": "(L+1)^2 per pair (final ribbon = full matrix)",
                                    "cells_incl_refills_gcups": float(cells1k.sum()) / sec1k / 1e9}
    pk = int_issue_peak(ctx)
    ach = cells * ALG_OPS_PER_CELL / (float(np.mean(kms)) * 1e-3) / 1e12
    return {"metric": "2D DO alignment GCUPS (ukkonenDO, S1 ribbon)", "value": cells_all / (ms * 1e-3) / 1e9, "unit": "GCUPS",
            "steps": steps, "ms_per_step": ms, "kernel_ms": float(np.mean(kms)), "pairs_per_gpu": n, "len": LU,
            "workload": "500-element prefixes of the config-2 strings, L1 matrix as an explicit overlap, dialect 1 (ukkonenDO)",
            "cells_per_pair_avg": cells / n, "cells": "final ribbon per pair", "clocks": clocks,
            "roofline": {"bound": "int32-issue", "achieved": ach, "peak": pk["peak"], "unit": "Tinstr/s",
                         "frac": ach / pk["peak"], "alg_ops_per_cell": ALG_OPS_PER_CELL},
            "cpu_baseline": cpu, "cost_checksum": checksum}


# ------------------------------------------------------------------------------------------------
# Config 1 (BASELINE.json configs[0], "bench-string-alignment-small"): the 190 unordered pairs of the 20 strings of
# the reference's bench/strings/dna/dna-sequences.fasta (lengths 8 ... 4608, half of them IUPAC-ambiguous), discrete
# TCM.  The strings travel as the committed fixture tests/golden/bench_strings_dna.json together with the costs the
# compiled reference produced for them.  It is the reference's own CPU-runnable case: a parity line more than a
# throughput line (190 pairs cannot fill a B200).
# ------------------------------------------------------------------------------------------------
def run_config1(args, ctx):
    import pcg_b200 as P
    from pcg_b200 import formats
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "bench_strings_dna.json")))
    codes = formats.iupac_dna_codes()
    enc = []
    for sq in g["text"]:
        e = formats.encode_fasta(sq, codes)
        enc.append(e[e != 16])
    tf = "dna-discrete.tcm"
    sigma, recs = np.array(g["tcms"][tf], np.uint32), g["pairs"][tf]
    firsts, seconds = [], []
    for i, j, *_ in recs:
        a, b = (enc[i], enc[j]) if len(enc[i]) >= len(enc[j]) else (enc[j], enc[i])
        firsts.append(a); seconds.append(b)
    pp = P.PackedPairs.from_lists(firsts, seconds)
    ctx.configure(pairs_per_warp=8, warps_per_cta=32, max_len=4700, max_b=32)
    tcm = ctx.tcm(sigma)
    for _ in range(max(args.warmup, 3)):
        r = P.align2d_batch(ctx, tcm, pp, want_cells=True)
    assert [int(x) for x in r.cost] == [rec[2] for rec in recs], "device costs differ from the reference's golden costs"
    t0 = time.perf_counter()
    kms = []
    for _ in range(args.steps):
        r = P.align2d_batch(ctx, tcm, pp, want_cells=True)
        kms.append(ctx.last_kernel_ms()[0])
    dt = (time.perf_counter() - t0) / args.steps
    cells = float(r.cells.sum())
    cpu = None
    if not args.no_cpu:
        from oracle import pyoracle as po
        kind = "reference" if os.path.exists(po.REF_SO) else "port"
        sec1, cost1, _ = po.cpu_bench_align2d(sigma, pp, 1, kind)
        secn, _, _ = po.cpu_bench_align2d(sigma, pp, host_threads(), kind)
        assert [int(x) for x in cost1] == [rec[2] for rec in recs]
        cpu = {"value": cells / sec1 / 1e9, "unit": "GCUPS", "cores": 1, "kind": kind,
               "us_per_pair": 1e6 * sec1 / len(recs), "all_threads_gcups": cells / secn / 1e9, "threads": host_threads(),
               "sample": "all 190 pairs, one pass, single thread (and all host threads); costs equal to the golden's"}
    return {"metric": "2D DO alignment GCUPS", "value": cells / (float(np.mean(kms)) * 1e-3) / 1e9, "unit": "GCUPS",
            "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": float(np.mean(kms)),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
            "data": "reference bench/strings/dna (committed fixture)",
            "config": {"workload": "config1: 190 pairs of the reference's DNA bench strings (8 ... 4608 elements, IUPAC "
                                   "ambiguity), discrete TCM, align2d", "pairs": len(recs), "cells": cells,
                       "us_per_pair_kernel": 1e3 * float(np.mean(kms)) / len(recs)},
            "e2e": {"value": cells / dt / 1e9, "unit": "GCUPS", "ms_per_step": dt * 1e3,
                    "h2d_bytes_per_step": int(pp.elems.nbytes + 40 * pp.n), "d2h_bytes_per_step": int(3 * pp.out_bytes + 16 * pp.n)},
            "gpu_launches": ctx.last_kernel_ms()[1] * args.steps, "cpu_baseline": cpu,
            "cost_checksum": int(np.asarray(r.cost, np.int64).sum())}

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
    ap.add_argument("--ppw", type=int, default=32, help="pairs a warp fills before tracing them")
    ap.add_argument("--workload", default="config2", choices=["config1", "config2", "config3", "config4", "config5"])
    ap.add_argument("--no-postorder", action="store_true", help="skip the config-5 post-order section")
    ap.add_argument("--po-trees", type=int, default=0, help="override the number of candidate trees")
    ap.add_argument("--po-steps", type=int, default=2)
    ap.add_argument("--reroot", action="store_true", help="config5: add the re-rooting pass (3 slots per node: use "
                                                           "--po-trees 256 to stay within HBM)")
    ap.add_argument("--no-fresh-trees", action="store_true", help="config5: skip the candidate-tree turnover variant")
    ap.add_argument("--no-library-multi", action="store_true", help="config5, N > 1: skip the one-process pcgdo_multi sweep")
    ap.add_argument("--no-extra", action="store_true", help="default workload: skip the config-3 / config-4 / Ukkonen sections")
    ap.add_argument("--section-steps", type=int, default=0, help="timed steps of the config-3 / config-4 sections (0 = 10)")
    ap.add_argument("--no-numa", action="store_true", help="N > 1: do not bind each rank to its GPU's NUMA node")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-compact", action="store_true", help="e2e: copy capacity-sized output slots back (no packing)")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    numa = bind_to_gpu_numa_node(local_rank) if world > 1 and not args.no_numa else None
    import torch
    import torch.distributed as dist
    import pcg_b200 as P
    from pcg_b200.backend import Batch
    import ctypes as C

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — this path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    # Everything below runs on one explicit stream: its handle is what the library launches on (a NULL handle would mean
    # the context's own non-blocking stream, which neither torch's events nor NCCL's stream ordering can see), the timing
    # events are recorded on it, and the all-reduce of config 5 is ordered after the sweep through it.
    torch.cuda.set_stream(torch.cuda.Stream(device=torch.device("cuda", local_rank)))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n = args.pairs
    ctx = P.Context(local_rank)
    if args.workload == "config1":
        if rank == 0:
            print(json.dumps(run_config1(args, ctx)), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return
    if args.workload == "config3":
        line = run_config3(args, rank, world, local_rank, ctx, torch, dist)
        if rank == 0:
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return
    if args.workload == "config4":
        line = run_config4(args, rank, world, local_rank, ctx, torch, dist)
        if rank == 0:
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return
    if args.workload == "config5":
        po = run_postorder(args, rank, world, local_rank, ctx, torch, dist)
        if rank == 0:
            line = {"metric": po["metric"], "value": po["value"], "unit": po["unit"], "n_gpus": world,
                    "steps": args.po_steps, "warmup": args.warmup, "ms_per_step": po["ms_per_sweep"],
                    "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32",
                    "data": "synthetic", "config": po["config"], "postorder": po}
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return
    ctx.configure(pairs_per_warp=args.ppw, warps_per_cta=args.warps, max_len=1024, max_b=32)
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
    cost_host = d_cost.cpu().numpy().copy()

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
    ctx_launches = ctx.last_kernel_ms()[1]

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
            # compact transfer (pcgdo_batch.out_off_actual): only the bytes the alignments produced come back
            h_actual = None if args.no_compact else torch.empty(n, dtype=torch.int64).pin_memory().numpy()
            # descriptors pinned too: a pageable source makes cudaMemcpyAsync stage 40 MB synchronously (~3 ms)
            h_off1, h_off2, h_len1, h_len2, h_ooff = (pin(a) for a in (pp.off1, pp.off2, pp.len1, pp.len2, pp.out_off))
            hb = Batch(n, h_elems.ctypes.data, pp.elems.nbytes, h_off1.ctypes.data, h_off2.ctypes.data,
                       h_len1.ctypes.data, h_len2.ctypes.data, h_ooff.ctypes.data, pp.out_bytes,
                       h_g.ctypes.data, h_s1.ctypes.data, h_s2.ctypes.data, h_olen.ctypes.data, h_cost.ctypes.data, None,
                       None if h_actual is None else h_actual.ctypes.data)

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
            # the strings that came back are the ones the device-resident run produced (spot check, 64 pairs)
            assert np.array_equal(h_olen, d_olen.cpu().numpy())
            where = pp.out_off if h_actual is None else h_actual.view(np.uint64)
            for k in np.linspace(0, n - 1, 64).astype(np.int64):
                o, oa, ln = int(pp.out_off[k]), int(where[k]), int(h_olen[k])
                for hbuf, dbuf in ((h_g, d_g), (h_s1, d_s1), (h_s2, d_s2)):
                    assert np.array_equal(hbuf[oa:oa + ln], dbuf[o:o + ln].cpu().numpy()), int(k)
            h2d = pp.elems.nbytes + 2 * 8 * n + 2 * 4 * n + 8 * n
            if h_actual is None:
                d2h = 3 * pp.out_bytes + 4 * n + 4 * n
            else:
                d2h = 3 * int(((h_olen.astype(np.int64) + 15) & ~15).sum()) + 4 * n + 4 * n + 8 * n   # packed in 16-byte units
            e2e = {"value": total_cells * world / dt / 1e9, "unit": "GCUPS", "h2d_bytes_per_step": int(h2d),
                   "d2h_bytes_per_step": int(d2h), "ms_per_step": dt * 1e3}
            del h_elems, h_g, h_s1, h_s2
        except Exception as ex:   # pinned allocation can fail on small hosts: report, do not fake
            e2e = {"value": None, "unit": "GCUPS", "error": str(ex)[:200]}

    # ---- second half of the metric: post-order tree scoring (config 5), same process, after the pair bench
    postorder = None
    del d_elems, d_g, d_s1, d_s2
    torch.cuda.empty_cache()
    if not args.no_postorder:
        try:
            postorder = run_postorder(args, rank, world, local_rank, ctx, torch, dist)
        except Exception as ex:
            postorder = {"value": None, "error": str(ex)[:300]}

    # ---- the other BASELINE configs as sections of the same line (each with roofline, cpu_baseline, e2e, clocks)
    sec_steps = args.section_steps or 10
    config3 = config4 = ukkonen = None
    if not args.no_extra:
        for name in ("config3", "config4", "ukkonen"):
            try:
                if name == "config3":
                    config3 = run_config3(args, rank, world, local_rank, ctx, torch, dist, steps=sec_steps, n=100000)
                elif name == "config4":
                    config4 = run_config4(args, rank, world, local_rank, ctx, torch, dist, steps=sec_steps, n=500000)
                else:
                    ukkonen = run_config2_ukkonen(args, rank, world, local_rank, ctx, torch, dist, steps=sec_steps)
            except Exception as ex:
                err = {"value": None, "error": str(ex)[:300]}
                if name == "config3":
                    config3 = err
                elif name == "config4":
                    config4 = err
                else:
                    ukkonen = err
            torch.cuda.empty_cache()

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
    pk = int_issue_peak(ctx)
    achieved_tops = total_cells * ALG_OPS_PER_CELL / (k_ms * 1e-3) / 1e12
    alg_bytes = n * (2 * L + 3 * 2 * L + 8) + total_cells / 4.0           # in + out + 2-bit directions spilled
    traffic, traffic_note = None, "no ncu capture committed for this build"
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            if tj.get("source_hash") == source_hash():
                traffic = tj.get("do_linear_kernel_bytes_per_launch_at_bench_size")
                traffic_note = tj.get("how")
            else:
                traffic_note = ("profiles/traffic.json was captured for other kernel sources (hash %s, now %s): not reported"
                                % (tj.get("source_hash"), source_hash()))
        except Exception:
            pass
    roofline = {"bound": "int32-issue", "achieved": achieved_tops, "peak": pk["peak"], "unit": "Tinstr/s",
                "frac": achieved_tops / pk["peak"], "traffic": traffic, "traffic_note": traffic_note,
                "peak_live": pk["live"], "peak_source": pk["source"],
                "alg_ops_per_cell": ALG_OPS_PER_CELL, "kernel_ms": k_ms,
                "hbm": {"alg_bytes_per_launch": alg_bytes, "achieved_gbs": alg_bytes / (k_ms * 1e-3) / 1e9,
                        "peak_gbs": hbm_peak, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                        "frac": alg_bytes / (k_ms * 1e-3) / 1e9 / hbm_peak}}

    cpu = None
    if not args.no_cpu and world == 1:
        info = cpu_arm(1 << 20, host_threads(), seconds_target=12.0)
        # the CPU arm ran the first pairs of the timed workload: its costs must be the device's
        assert np.array_equal(info["cost"], cost_host[:info["pairs"]]), "device costs differ from the reference's on the CPU sample"
        cpu = {"value": info["gcups"], "unit": "GCUPS", "cores": info["cores"], "kind": info["kind"],
               "sample": f"first {info['pairs']} pairs of the same seeded workload, {info['seconds']:.1f} s on "
                         f"{info['cores']} host threads (oracle/_ref align2d, per-call malloc/free as FFI.hsc); "
                         f"its {info['pairs']} costs equal the device's"}

    line = {"metric": "2D DO alignment GCUPS", "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": "config2: batched 2D DO, 1 kb uniform DNA pairs, L1 TCM, align2d band (C-linear)",
                       "pairs_per_gpu": n, "len": L, "cells_per_pair": cpp, "full_equiv_gcups": value * L * L / cpp,
                       "l2": "inputs (2.1 GB) and direction scratch (>4 GB) exceed the 126 MB L2",
                       "seed": SEED, "source_hash": source_hash(), "numa_binding_rank0": numa},
            "clocks": clocks, "e2e": e2e, "gpu_launches": ctx_launches * args.steps, "roofline": roofline, "cpu_baseline": cpu, "postorder": postorder,
            "config3": config3, "config4": config4, "config2_ukkonen": ukkonen,
            "cost_checksum": cost_checksum}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
