"""Top instructions by warp-stall samples from an `ncu --set full --import-source on` report.
   python tools/ncu_source_top.py report.ncu-rep <kernel regex> [top_n]"""
import csv
import io
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{rx}"], capture_output=True, text=True).stdout
lines = raw.splitlines()
# the first kernel only
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))
rd = list(csv.DictReader(io.StringIO("\n".join(lines[start:end]))))
stalls = [k for k in rd[0] if k.startswith("stall_") and "Not Issued" not in k]
tot = {k: 0 for k in stalls}
rows = []
for i, r in enumerate(rd):
    n = int(r["# Samples"] or 0)
    st = {k: int(r[k] or 0) for k in stalls}
    for k in stalls:
        tot[k] += st[k]
    rows.append((n, i, r["Source"].strip(), int(r["Instructions Executed"] or 0), st))
allsamp = sum(r[0] for r in rows)
print(f"# {lines[0][:160]}")
print(f"# {len(rows)} SASS instructions, {allsamp} samples; stall totals: " + ", ".join(f"{k[6:]} {v}" for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v * 50 > allsamp))
print(f"# executed warp instructions: {sum(r[3] for r in rows)}")
for n, i, src, ex, st in sorted(rows, key=lambda r: -r[0])[:top]:
    ts = " ".join(f"{k[6:]}:{v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3] if v)
    print(f"{n:6d} {100 * n / allsamp:5.1f}%  #{i:4d} x{ex:8d}  {src[:70]:70s} {ts}")
