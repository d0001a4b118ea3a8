"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one steady-state step
(delimited by the adamw_kernel launches) aggregated by kernel.  Usage:
    python scripts/summarize_launches.py gpurun_out/launches.csv [step_index] > profiles/<name>.txt"""
import collections
import csv
import re
import sys


def main(path, step=3):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [x["Kernel Name"] for x in rows]
    ends = [i for i, n in enumerate(names) if "adamw_kernel" in n]
    step = min(step, len(ends) - 1)
    a, b = ends[step - 1] + 1, ends[step] + 1
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for x in rows[a:b]:
        t = float(x["Metric Value"].replace(",", "")) / 1e3
        n = re.sub(r"<.*", "", x["Kernel Name"])[:70]
        agg[n][0] += 1
        agg[n][1] += t
        tot += t
    print(f"# {path}: step {step} of {len(ends)}, {b - a} launches, {tot / 1e3:.3f} ms of kernel time "
          f"(ncu per-launch times are cold-cache and serialised: compare SHARES)")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        print(f"{t:10.1f} us {100 * t / tot:5.1f}%  x{c:<4d} {n}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 3)
