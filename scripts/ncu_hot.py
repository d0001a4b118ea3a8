"""Top SASS instructions by warp-stall samples from `ncu --page source --csv` output."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, isamp, iexec = hdr.index("Address"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
data = []
for i, r in enumerate(rows[2:]):
    try:
        data.append((int(r[isamp]), i, r[isrc].strip(), int(r[iexec])))
    except Exception:
        pass
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
for s, i, src, ex in sorted(data, reverse=True)[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"{100*s/tot:5.1f}% line {i:5d} exec {ex:9d}  {src[:110]}")
