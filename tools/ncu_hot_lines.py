#!/usr/bin/env python
"""Top source lines of one kernel by warp-stall samples, from an `ncu --set full --import-source on` capture read
offline (`ncu -i X.ncu-rep --page source --print-source cuda,sass --csv`).
Usage: python tools/ncu_hot_lines.py X.ncu-rep [N] > profiles/<name>.md"""
import csv
import io
import subprocess
import sys


def main():
    path, top_n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True, check=True).stdout
    cur_file, hdr, lines = None, None, []
    for r in csv.reader(io.StringIO(out)):
        if len(r) == 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr and len(r) == len(hdr) and r[0].strip().isdigit():        # a source-line row (SASS rows have an empty Line No)
            d = {}
            for k, v in zip(hdr, r):
                d.setdefault(k, v)                                        # "Source" appears twice: keep the CUDA one
            lines.append((cur_file, int(r[0]), d))
    num = lambda d, k: float(d.get(k, "0").replace(",", "") or 0) if d.get(k, "0").replace(",", "").replace(".", "").isdigit() else 0.0  # noqa: E731
    tot = sum(num(d, "# Samples") for _, _, d in lines) or 1.0
    tot_i = sum(num(d, "Instructions Executed") for _, _, d in lines) or 1.0
    stall_cols = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    print(f"Capture `{path.split('/')[-1]}`: {int(tot):,} warp samples, {int(tot_i):,} warp instructions attributed to source lines.\n")
    print("| % samples | % instructions | avg active threads | file:line | top stall reasons | source |")
    print("|---|---|---|---|---|---|")
    for f, ln, d in sorted(lines, key=lambda x: -num(x[2], "# Samples"))[:top_n]:
        s = num(d, "# Samples")
        st = sorted(((num(d, k), k[6:]) for k in stall_cols), reverse=True)[:3]
        reasons = ", ".join(f"{k} {100 * v / max(s, 1):.0f} %" for v, k in st if v > 0)
        thr = num(d, "Thread Instructions Executed") / max(1.0, num(d, "Instructions Executed"))
        src = d.get("Source", "").strip().replace("|", "\\|")[:120]
        print(f"| {100 * s / tot:.1f} | {100 * num(d, 'Instructions Executed') / tot_i:.1f} | {thr:.1f} | {f}:{ln} | {reasons} | `{src}` |")


if __name__ == "__main__":
    main()
