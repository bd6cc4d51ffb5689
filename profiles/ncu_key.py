#!/usr/bin/env python
"""Extract the handful of per-kernel numbers bench.py quotes (DRAM traffic per launch, pipe utilisation)
from an ncu report into a small committed JSON.   usage: ncu_key.py report.ncu-rep out.json"""
import csv
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slots_busy_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu_pipe_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "registers",
}
UNIT = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}


def main():
    raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, body = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    out = {"source": sys.argv[1].split("/")[-1], "kernels": {}}
    for r in body:
        name = r[idx["Kernel Name"]]
        short = name.split("(")[0].split("<")[0].split("::")[-1].split(" ")[-1]  # bare function name
        if short in out["kernels"]:
            continue  # first launch of each kernel
        k = {"full_name": name}
        for key, label in KEYS.items():
            if key in idx and r[idx[key]] != "":
                v = float(r[idx[key]].replace(",", ""))
                u = units[idx[key]]
                k[label] = v * UNIT[u] if u in UNIT else v
        if "dram_read" in k and "dram_write" in k:
            k["dram_traffic_bytes"] = k["dram_read"] + k["dram_write"]
        out["kernels"][short] = k
    json.dump(out, open(sys.argv[2], "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
