"""profiles/traffic.json from `ncu --set full` reports: mean DRAM bytes (read + write) and mean duration per launch of
each kernel family bench.py reports (roofline.traffic is 'per launch like achieved').

    python tools/ncu_traffic.py gpurun_out/prof_gemm_r1e.ncu-rep [more.ncu-rep ...] > profiles/traffic.json
"""
import csv
import io
import json
import subprocess
import sys

FAMILY = [("gemm_", "pwconv_gemm"), ("dwconv", "dwconv"), ("bn_", "bn_pass"), ("ctc_lattice", "ctc_fwd"),
          ("ctc_grad", "ctc_bwd")]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TUNIT = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def main():
    acc = {}
    for rep in sys.argv[1:]:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        kn = hdr.index("Kernel Name")
        ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
        for r in rows[2:]:
            fam = next((f for key, f in FAMILY if key in r[kn]), None)
            if fam is None:
                continue
            rd = float(r[ir].replace(",", "")) * UNIT.get(units[ir], 1.0)
            wr = float(r[iw].replace(",", "")) * UNIT.get(units[iw], 1.0)
            us = float(r[it].replace(",", "")) * TUNIT.get(units[it], 1.0)
            a = acc.setdefault(fam, {"launches": 0, "bytes": 0.0, "us": 0.0, "source": []})
            a["launches"] += 1
            a["bytes"] += rd + wr
            a["us"] += us
            if rep not in a["source"]:
                a["source"].append(rep)
    out = {f: {"dram_bytes_per_launch": a["bytes"] / a["launches"], "us_per_launch_under_ncu": a["us"] / a["launches"],
               "launches": a["launches"], "source": a["source"],
               "how": "ncu --set full --clock-control none, one eager training step of asr13x1_b32_16s_bf16"}
           for f, a in acc.items()}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
