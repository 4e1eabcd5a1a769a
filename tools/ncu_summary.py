#!/usr/bin/env python
"""Compact per-kernel summary of `ncu --set full` captures (.ncu-rep), read offline with `ncu -i ... --page raw --csv`:
duration, DRAM traffic, issue utilisation, the warp-stall breakdown (PC sampling) and the shared-memory atomic /
load / store traffic.  Usage: python tools/ncu_summary.py a.ncu-rep [b.ncu-rep ...] > profiles/<name>.md"""
import csv
import io
import subprocess
import sys


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, r)) for r in rows[2:]], dict(zip(hdr, units))


def f(d, k, default=0.0):
    try:
        return float(str(d.get(k, "")).replace(",", ""))
    except ValueError:
        return default


def to_bytes(v, unit):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def to_us(v, unit):
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit, 1)


def main():
    print("| capture | kernel | grid x block | µs | DRAM read + write (MB) | DRAM % of peak | issue slots busy % | regs / thread | "
          "achieved occupancy % | shared atomics (inst; wavefronts; bank conflicts; G wavefronts/s) | shared ld / st wavefronts |")
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    stalls = []
    for path in sys.argv[1:]:
        rows, units = raw(path)
        for d in rows:
            name = d.get("Kernel Name", "?").split("(")[0].replace("void ", "").replace("<unnamed>::", "").replace("unnamed>::", "").strip()
            us = to_us(f(d, "gpu__time_duration.sum"), units.get("gpu__time_duration.sum", "ns"))
            rd = to_bytes(f(d, "dram__bytes_read.sum"), units.get("dram__bytes_read.sum", "byte"))
            wr = to_bytes(f(d, "dram__bytes_write.sum"), units.get("dram__bytes_write.sum", "byte"))
            at_i, at_w = f(d, "smsp__inst_executed_op_shared_atom.sum"), f(d, "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum")
            at_c = f(d, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum")
            cap = path.split("/")[-1].replace(".ncu-rep", "")
            print(f"| `{cap}` | `{name}` | {d.get('Grid Size', '?')} x {d.get('Block Size', '?')} | {us:.1f} | {rd / 1e6:.1f} + {wr / 1e6:.1f} | "
                  f"{f(d, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | {f(d, 'sm__inst_issued.avg.pct_of_peak_sustained_active') or f(d, 'smsp__issue_active.avg.pct'):.1f} | "
                  f"{int(f(d, 'launch__registers_per_thread'))} | {f(d, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | "
                  f"{int(at_i):,}; {int(at_w):,}; {int(at_c):,}; {at_w / max(us, 1e-9) / 1e3:.3f} | "
                  f"{int(f(d, 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum')):,} / {int(f(d, 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum')):,} |")
            samples = {k[len("smsp__pcsamp_warps_issue_stalled_"):]: f(d, k) for k in d
                       if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued")}
            tot = sum(samples.values()) or 1.0
            top = sorted(samples.items(), key=lambda kv: -kv[1])
            stalls.append((cap, name, int(tot), [(k, 100.0 * v / tot) for k, v in top if v / tot >= 0.02]))
    print("\nWarp-stall breakdown (PC sampling, share of all warp samples; `selected` = the warp issued):\n")
    print("| capture | kernel | samples | reasons >= 2 % |")
    print("|---|---|---|---|")
    for cap, name, tot, top in stalls:
        print(f"| `{cap}` | `{name}` | {tot:,} | " + ", ".join(f"{k} {p:.1f} %" for k, p in top) + " |")


if __name__ == "__main__":
    main()
