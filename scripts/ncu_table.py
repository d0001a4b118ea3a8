"""Markdown summary of an `ncu --set full` report: one table per kernel, one row per profiled launch.
    python scripts/ncu_table.py gpurun_out/prof.ncu-rep > profiles/<name>.md      (runs without a GPU)"""
import collections
import csv
import io
import re
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "duration"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
        ("sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed", "UTCHMMA bf16 % of peak"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
        ("launch__registers_per_thread", "regs/thread"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__occupancy_limit_shared_mem", "CTAs/SM (smem limit)"),
        ("sm__cycles_elapsed.max", "SM cycles"), ("gpc__cycles_elapsed.avg.per_second", "SM clock")]


def main(path):
    if path.endswith(".csv"):                        # already exported on the GPU box (ncu -i rep --page raw --csv)
        raw = open(path).read()
    else:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {n: i for i, n in enumerate(hdr)}
    by = collections.OrderedDict()
    for r in data:
        by.setdefault(re.sub(r"\(.*", "", r[idx["Kernel Name"]]), []).append(r)
    print(f"# {path}: ncu --set full --clock-control none (one B200; per-launch, serialised, cold-ish caches)\n")
    for name, rs in by.items():
        print(f"## {name}  ({len(rs)} launches profiled)")
        cols = [(k, t) for k, t in COLS if k in idx]
        print("| launch | " + " | ".join(t for _, t in cols) + " |")
        print("|---" * (len(cols) + 1) + "|")
        for i, r in enumerate(rs):
            print(f"| {i} | " + " | ".join(f"{r[idx[k]]} {units[idx[k]]}".strip() for k, _ in cols) + " |")
        print()


if __name__ == "__main__":
    main(sys.argv[1])
