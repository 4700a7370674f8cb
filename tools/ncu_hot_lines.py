#!/usr/bin/env python
"""Top source lines of an ncu report by executed warp instructions / stall samples.

    ncu -i report.ncu-rep --page source --csv --print-source cuda,sass > src.csv
    python tools/ncu_hot_lines.py src.csv [N]
"""
import csv
import sys


def main():
    path = sys.argv[1]
    topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    kernel = None
    fname = "?"
    per_kernel = {}
    hdr = None
    for row in csv.reader(open(path, newline="")):
        if not row:
            continue
        if row[0] == "File Path":
            fname = row[1].rsplit("/", 1)[-1]
            continue
        if row[0] == "Function Name":
            kernel = row[1].split("(")[0][-40:]
            per_kernel.setdefault(kernel, [])
            continue
        if row[0] == "Line No":
            hdr = row
            continue
        if hdr is None or kernel is None or row[0] in ("", "File Name"):
            continue
        try:
            line_no = int(row[0])
        except ValueError:
            continue
        rec = dict(zip(hdr[4:], row[4:]))
        try:
            inst = int(rec.get("Instructions Executed", "0") or 0)
            samples = int(rec.get("# Samples", "0") or 0)
            tinst = int(rec.get("Thread Instructions Executed", "0") or 0)
        except ValueError:
            continue
        per_kernel[kernel].append((inst, samples, tinst, line_no, fname + ": " + row[1].strip()[:100]))
    for k, rows in per_kernel.items():
        total = sum(r[0] for r in rows) or 1
        tsamp = sum(r[1] for r in rows) or 1
        print(f"\n=== {k}\n    total warp-inst {total:,}  samples {tsamp:,}")
        for inst, samples, tinst, line_no, src in sorted(rows, reverse=True)[:topn]:
            eff = tinst / inst / 32 if inst else 0
            print(f"{100 * inst / total:6.2f}% inst {100 * samples / tsamp:6.2f}% stall  lanes {eff:4.2f}  L{line_no:<4d} {src}")


if __name__ == "__main__":
    main()
