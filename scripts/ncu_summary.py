#!/usr/bin/env python
"""Condense `ncu --set full` reports into the few numbers the roofline arguments use.

    python scripts/ncu_summary.py gpurun_out/a.ncu-rep [b.ncu-rep ...] > profiles/R2_ncu_summary.txt

Per profiled launch: kernel, duration, SM clock, tensor-pipe activity (sm__pipe_tensor_cycles_active,
sm__inst_executed_pipe_tc / _tmem / _tma), issue activity, XU (MUFU) and ALU/FMA pipes, DRAM bytes read + written
(= `roofline.traffic`), L2 throughput, achieved occupancy, registers.  Reads the report with `ncu -i ... --page raw --csv`
(no GPU needed).
"""
import csv
import io
import subprocess
import sys

COLS = [
    ("gpu__time_duration.sum", "dur_us", 1e-3),
    ("sm__cycles_active.avg", "sm_cycles", 1),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_%", 1),
    ("sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active", "inst_tc_%", 1),
    ("sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "inst_tmem_%", 1),
    ("sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active", "inst_tma_%", 1),
    ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue_%", 1),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu_%", 1),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu_%", 1),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_%", 1),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu_%", 1),
    ("dram__bytes_read.sum", "dram_rd", 1),
    ("dram__bytes_write.sum", "dram_wr", 1),
    ("lts__t_bytes.sum", "l2_bytes", 1),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_%", 1),
    ("sm__warps_active.avg.per_cycle_active", "warps_active", 1),
    ("launch__registers_per_thread", "regs", 1),
    ("launch__grid_size", "grid", 1),
]


def main():
    print("# " + " | ".join(["kernel"] + [c[1] for c in COLS]) + "   (units as ncu prints them: see the `unit` line per file)")
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            print(f"{rep}: no data")
            continue
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        print(f"## {rep}")
        print("unit | " + " | ".join(units[idx[c[0]]] if c[0] in idx else "-" for c in COLS))
        for r in rows[2:]:
            name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("ezr::", "")
            vals = []
            for key, _, _ in COLS:
                v = r[idx[key]] if key in idx else "-"
                try:
                    v = f"{float(v):.4g}"
                except ValueError:
                    pass
                vals.append(v)
            print(name + " | " + " | ".join(vals))


if __name__ == "__main__":
    main()
