#!/usr/bin/env python
"""Summarises an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv` launch list of
bench.py into the per-class DRAM traffic file bench.py reads for `roofline.traffic`:

    python tools/ncu_dram_summary.py gpurun_out/r2_ncu_dram_2048.csv profiles/r2_ncu_dram_2048.json [--skip-after N]

Classes are bench.py's timing classes (one kernel family each). Only the launches of the FIRST IC are used
(--ic-launches: how many launches of each class one IC makes; the rest of the list is the checksum IC and the cross-check).
"""
import csv
import json
import re
import sys

CLASSES = [("fft_cols_x", re.compile(r"fft_cols_kernel<.*PassX")), ("fft_cols_y", re.compile(r"fft_cols_kernel<.*PassY")),
           ("fft_rows_c2r", re.compile(r"fft_rows_c2r")), ("rng_gen", re.compile(r"rng_gen_kernel")),
           ("rng_deliver", re.compile(r"deliver_kernel")), ("mirror", re.compile(r"mirror_kernel")),
           ("transfer_table", re.compile(r"transfer_table_kernel"))]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
        "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}


def main():
    src, dst = sys.argv[1], sys.argv[2]
    rows = []
    with open(src, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.DictReader(lines)
    launches = {}
    order = []
    for r in rd:
        i = int(r["ID"])
        if i not in launches:
            launches[i] = {"name": r["Kernel Name"], "m": {}}
            order.append(i)
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        launches[i]["m"][r["Metric Name"]] = v * UNIT.get(r["Metric Unit"], 1.0)
    out = {}
    total_t = sum(l["m"].get("gpu__time_duration.sum", 0.0) for l in launches.values())
    for cname, rx in CLASSES:
        ls = [launches[i] for i in order if rx.search(launches[i]["name"])]
        if not ls:
            continue
        half = ls[:max(1, len(ls) // 2)]       # bench.py --steps 1 --warmup 0 runs the IC twice (timed step + checksum IC)
        rd_b = sum(l["m"].get("dram__bytes_read.sum", 0.0) for l in half) / len(half)
        wr_b = sum(l["m"].get("dram__bytes_write.sum", 0.0) for l in half) / len(half)
        t = sum(l["m"].get("gpu__time_duration.sum", 0.0) for l in ls)
        out[cname] = {"launches_in_list": len(ls), "dram_bytes_read_per_launch": rd_b, "dram_bytes_write_per_launch": wr_b,
                      "dram_bytes_per_launch": rd_b + wr_b, "share_of_listed_gpu_time": t / total_t if total_t else None,
                      # the first IC's launches only: the tail of the list holds the small cross-check launches
                      "avg_ms_per_launch_under_ncu": 1e3 * sum(l["m"].get("gpu__time_duration.sum", 0.0) for l in half) / len(half),
                      "note": "ncu dram__bytes_read.sum + dram__bytes_write.sum, mean over the class's launches of one IC "
                              "(%s); cold-cache serialised replays" % src}
    json.dump(out, open(dst, "w"), indent=1)
    for k, v in out.items():
        print("%-16s %3d launches  dram r %.3f GB  w %.3f GB per launch  share %.3f" % (
            k, v["launches_in_list"], v["dram_bytes_read_per_launch"] / 1e9, v["dram_bytes_write_per_launch"] / 1e9,
            v["share_of_listed_gpu_time"] or 0))


if __name__ == "__main__":
    main()
