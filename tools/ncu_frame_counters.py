#!/usr/bin/env python
"""Counters of ONE frame out of an `ncu --set full` report of the frame's kernels (tools/profile_r2.sh captures
exactly the launches of one serial frame): warp instructions executed and DRAM bytes moved, per kernel and
summed.  bench.py reads the JSON for `roofline.traffic` and `roofline.frac_executed`.

    python tools/ncu_frame_counters.py gpurun_out/frame_serial.ncu-rep > profiles/r02/frame_counters.json
"""
import csv
import json
import subprocess
import sys


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {n: hdr.index(n) for n in ("Kernel Name", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                     "gpu__time_duration.sum", "smsp__thread_inst_executed.sum") if n in hdr}

    def value(row, name):
        v = float(row[col[name]].replace(",", ""))
        unit = units[col[name]].lower()
        scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "inst": 1, "": 1}.get(unit, 1)
        return v * scale

    kernels = []
    for d in data:
        name = d[col["Kernel Name"]].split("(")[0].split("::")[-1]
        kernels.append({"kernel": name, "us": value(d, "gpu__time_duration.sum"),
                        "warp_instructions": value(d, "smsp__inst_executed.sum"),
                        "thread_instructions": value(d, "smsp__thread_inst_executed.sum") if "smsp__thread_inst_executed.sum" in col else None,
                        "dram_bytes_read": value(d, "dram__bytes_read.sum"), "dram_bytes_written": value(d, "dram__bytes_write.sum")})
    # the capture may start mid-frame: keep the launches from the first primary kernel to the one before the next
    starts = [i for i, k in enumerate(kernels) if "primary" in k["kernel"]]
    if starts:
        end = starts[1] if len(starts) > 1 else len(kernels)
        kernels = kernels[starts[0]:end]
    doc = {
        "what": "one serial frame of the headline workload (MCSKIN_FRAME_LANES=1 MCSKIN_GRAPHS=0), ncu --set full --clock-control none",
        "report": rep,
        "warp_instructions_per_frame": sum(k["warp_instructions"] for k in kernels),
        "dram_bytes_per_frame": sum(k["dram_bytes_read"] + k["dram_bytes_written"] for k in kernels),
        "us_per_frame_under_ncu": sum(k["us"] for k in kernels),
        "kernels": kernels,
    }
    print(json.dumps(doc, indent=1))


if __name__ == "__main__":
    main()
