#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --set full` capture of the config-2 kernel at bench size: DRAM bytes per launch,
tied to the kernel sources by bench.source_hash() (bench.py reports `roofline.traffic` only when the hash matches)."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(rep, note):
    import bench
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, unit = rows[0], rows[1]
    d = dict(zip(hdr, rows[2]))
    u = dict(zip(hdr, unit))

    def to_bytes(k):
        v = float(d[k])
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u[k]]
    rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
    j = {"do_linear_kernel_bytes_per_launch_at_bench_size": int(rd + wr), "source_hash": bench.source_hash(),
         "how": f"ncu --set full --clock-control none, {os.path.basename(rep)} ({d['Kernel Name'][:60]}, {note}): "
                f"dram__bytes_read.sum {rd / 1e9:.3f} GB + dram__bytes_write.sum {wr / 1e9:.3f} GB"}
    json.dump(j, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"))
    print(j)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
