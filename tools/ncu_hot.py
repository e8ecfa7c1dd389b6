#!/usr/bin/env python
"""Top stall sites of the FIRST kernel in an `ncu --page source --csv --print-source sass` export."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:
    if len(r) != len(hdr) or r[0] == "Address":
        break
    data.append(r)
tot = sum(int(r[ix["# Samples"]]) for r in data)
print("total samples", tot, "instructions", len(data))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
top = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]]))[:n]
for i in sorted(top):
    r = data[i]
    s = {h: int(r[ix[h]]) for h in stalls}
    main = sorted(s.items(), key=lambda kv: -kv[1])[:2]
    print(i, r[ix["Source"]].strip()[:72].ljust(72), r[ix["# Samples"]].rjust(6), r[ix["Instructions Executed"]].rjust(8), main)
agg = {h: sum(int(r[ix[h]]) for r in data) for h in stalls}
print(sorted(agg.items(), key=lambda kv: -kv[1]))
