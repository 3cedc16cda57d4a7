#!/usr/bin/env python
"""Runs here (no GPU): exports every gpurun_out/<tag>_*.ncu-rep to profiles/<tag>_<name>_ncu_full_raw.csv
(`ncu -i ... --page raw --csv`) and prints / writes a one-line-per-kernel summary of the metrics the roofline
arguments use (duration, DRAM bytes, fp64 / DMMA / tensor pipe utilisation, L2 hit rate, registers, grid).
    python tools/ncu_extract.py r02"""
import csv
import glob
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
        "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}
COLS = [("gpu__time_duration.sum", "ms"), ("dram__bytes_read.sum", "B"), ("dram__bytes_write.sum", "B"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "%"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "%"),
        ("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "%"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "%"),
        ("lts__t_sector_hit_rate.pct", "%"), ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "%"),
        ("launch__registers_per_thread", ""), ("launch__grid_size", ""), ("launch__block_size", ""),
        ("smsp__inst_executed.sum", "")]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    lines = []
    for rep in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", tag + "_*.ncu-rep"))):
        name = os.path.basename(rep)[len(tag) + 1:-len(".ncu-rep")]
        res = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        if res.returncode != 0 or not res.stdout.strip():
            lines.append("%s: ncu export failed: %s" % (name, res.stderr.strip()[:200]))
            continue
        out = os.path.join(ROOT, "profiles", "%s_%s_ncu_full_raw.csv" % (tag, name))
        with open(out, "w") as f:
            f.write(res.stdout)
        rows = list(csv.reader(io.StringIO(res.stdout)))
        names, units = rows[0], rows[1]
        for r in rows[2:]:
            d = dict(zip(names, r))
            u = dict(zip(names, units))
            parts = ["%s | %s" % (name, d.get("Kernel Name", "?")[:60])]
            for col, want in COLS:
                if col not in d or d[col] == "":
                    continue
                try:
                    v = float(d[col].replace(",", ""))
                except ValueError:
                    continue
                if want in ("ms", "B"):
                    v *= UNIT.get(u[col], 1.0)
                parts.append("%s=%.6g%s" % (col.split(".")[0].replace("__", ":"), v, want))
            lines.append("  ".join(parts))
    text = "\n".join(lines)
    print(text)
    with open(os.path.join(ROOT, "profiles", "%s_ncu_summary.txt" % tag), "w") as f:
        f.write(text + "\n")


if __name__ == "__main__":
    main()
