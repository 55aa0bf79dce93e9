"""Instruction mix + stall samples per opcode from `ncu -i X.ncu-rep --page source --csv` output (first kernel only)."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
print(rows[0][1][:150])
hdr = rows[1]
isrc, iex, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
cnt, smp = collections.Counter(), collections.Counter()
for r in rows[2:]:
    if len(r) <= max(iex, ismp):
        continue
    if r[0] == "Kernel Name" or r[iex] == "Instructions Executed":
        break
    s = re.sub(r"^@!?U?P\d+\s+", "", r[isrc].strip())
    op = ".".join(s.split()[0].split(".")[:2]) if s else ""
    cnt[op] += int(r[iex])
    smp[op] += int(r[ismp])
tot, ts = sum(cnt.values()), max(1, sum(smp.values()))
print("total warp-instructions", tot)
for op, c in cnt.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
    print(f"{op:22s} {c:11d} {100 * c / tot:5.1f}%   stall samples {100 * smp[op] / ts:5.1f}%")
