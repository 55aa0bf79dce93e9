"""Summarise an `ncu --set full` report: one line per launch with the numbers DESIGN.md / bench.py quote.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_name.txt
"""
import csv
import io
import subprocess
import sys

COLS = [
    ("gpu__time_duration.sum", "us"),
    ("dram__bytes_read.sum", "rdMB"),
    ("dram__bytes_write.sum", "wrMB"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tc%"),
    ("sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active", "tcinst%"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]


def to_unit(v, unit, want):
    v = float(v.replace(",", "")) if v not in ("", "n/a") else float("nan")
    if want == "us":
        return v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
    if want in ("rdMB", "wrMB"):
        return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1.0)
    return v


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = [(hdr.index(m), lab) for m, lab in COLS if m in hdr]
    kn = hdr.index("Kernel Name")
    print(f"# {rep}: ncu --set full --clock-control none (per-launch, cold-cache, serialised)")
    print("  ".join(f"{lab:>8s}" for _, lab in idx) + "  kernel")
    for r in rows[2:]:
        vals = [to_unit(r[i], units[i], lab) for i, lab in idx]
        name = r[kn].replace("lasr::", "").replace("void ", "")
        name = name.split("(CUtensorMap")[0].split("(const")[0][:90]
        print("  ".join(f"{v:8.1f}" for v in vals) + "  " + name)


if __name__ == "__main__":
    main()
