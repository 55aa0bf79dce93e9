"""Summarise an ncu launch list taken with --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum:
per kernel name: launches, total / mean duration, share, mean DRAM bytes per launch; and (--traffic) the
profiles/traffic.json that bench.py's roofline.traffic reads (mean DRAM bytes per launch of each kernel family).

    python tools/launch_summary2.py gpurun_out/launches_r1e.csv [--traffic profiles/traffic.json] > profiles/r1e_launches_summary.txt
"""
import collections
import csv
import json
import sys

FAMILY = [("gemm_", "pwconv_gemm"), ("dwconv", "dwconv"), ("bn_", "bn_pass"), ("ctc_lattice", "ctc_fwd"),
          ("ctc_grad", "ctc_bwd"), ("novograd", "novograd")]


def main():
    path = sys.argv[1]
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[start]
    kn, mn, mv, idc = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
    per = collections.OrderedDict()
    for r in rows[start + 1:]:
        if len(r) <= mv:
            continue
        d = per.setdefault(int(r[idc]), {"name": r[kn]})
        d[r[mn]] = float(r[mv].replace(",", ""))
    agg = collections.OrderedDict()
    fam = {}
    for d in per.values():
        name = d["name"].replace("void ", "").split("(")[0][:100]
        a = agg.setdefault(name, [0, 0.0, 0.0])
        us = d.get("gpu__time_duration.sum", 0.0) / 1e3
        by = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        a[0] += 1
        a[1] += us
        a[2] += by
        f = next((v for k, v in FAMILY if k in d["name"]), None)
        if f:
            fa = fam.setdefault(f, [0, 0.0, 0.0])
            fa[0] += 1
            fa[1] += us
            fa[2] += by
    total = sum(a[1] for a in agg.values())
    print(f"# {path}: {len(per)} launches, {total:.1f} us of kernel time (cold-cache, serialised under ncu)")
    print(f"# {'us':>9s} {'share':>6s} {'n':>5s} {'avg us':>8s} {'DRAM MB/launch':>15s}  kernel")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{a[1]:10.1f} {100 * a[1] / total:5.1f}% {a[0]:5d} {a[1] / a[0]:8.1f} {a[2] / a[0] / 1e6:15.2f}  {name}")
    print("# families (bench.py roofline.families):")
    for f, a in fam.items():
        print(f"#   {f:12s} launches {a[0]:4d}  share {100 * a[1] / total:5.1f}%  mean {a[1] / a[0]:6.1f} us  "
              f"mean DRAM {a[2] / a[0] / 1e6:7.2f} MB/launch")
    if "--traffic" in sys.argv:
        out = {f: {"dram_bytes_per_launch": a[2] / a[0], "us_per_launch_under_ncu": a[1] / a[0], "launches": a[0],
                   "share_of_kernel_time": a[1] / total, "source": "profiles/" + path.split("/")[-1],
                   "how": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                          "--clock-control none over `python bench.py --steps 2 --warmup 1 --no-cpu-baseline`"}
               for f, a in fam.items()}
        json.dump(out, open(sys.argv[sys.argv.index("--traffic") + 1], "w"), indent=1)


if __name__ == "__main__":
    main()
