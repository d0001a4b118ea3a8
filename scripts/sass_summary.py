"""SASS evidence for profiles/: per kernel of libmimic_b200's objects, the count of the Blackwell-native mnemonics
(UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG / UTMAREDG = TMA load / store / reduce, UTCBAR =
tcgen05.commit, SYNCS = mbarrier, USETMAXREG = setmaxnreg, UCGABAR = cluster barrier) next to the legacy ones (HMMA =
mma.sync, which must be absent), registers and spills from the ptxas log.  Runs without a GPU:
    python scripts/sass_summary.py > profiles/r02_sass_summary.md"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKRED", "UTCBAR", "SYNCS", "USETMAXREG",
        "UCGABAR", "MUFU.EX2", "HMMA", "LDGSTS", "STL", "LDL"]


def main():
    print("# SASS of libmimic_b200.so's objects (cuobjdump -sass), round 2 final tree\n")
    print("| object | kernel | instrs | " + " | ".join(WANT) + " |")
    print("|---|---|---|" + "---|" * len(WANT))
    for obj in sorted(glob.glob(os.path.join(ROOT, "mimic_b200", "_build", "*.o"))):
        sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        cur, counts, n = None, collections.OrderedDict(), {}
        for line in sass.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                cur = subprocess.run(["cu++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
                cur = cur[: cur.rfind("(")] if "(" in cur else cur
                cur = cur.replace("void ", "").replace("mimic::", "")
                counts[cur] = collections.Counter()
                n[cur] = 0
                continue
            if cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
                n[cur] += 1
                for w in WANT:
                    if re.search(r"(?<![A-Z])" + re.escape(w), line):
                        counts[cur][w] += 1
        for kname, c in counts.items():
            if n[kname] < 50:
                continue
            print(f"| {os.path.basename(obj)} | `{kname[:90]}` | {n[kname]} | " + " | ".join(str(c.get(w, 0)) for w in WANT) + " |")
    print("\n## ptxas -v (registers, spills) per object\n")
    for log in sorted(glob.glob(os.path.join(ROOT, "mimic_b200", "_build", "*.o.log"))):
        txt = open(log).read()
        for m in re.finditer(r"Compiling entry function '(\S+)'.*?\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", txt):
            name = subprocess.run(["cu++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = name[: name.rfind("(")] if "(" in name else name
            print(f"- `{os.path.basename(log)[:-6]}` `{name[-80:]}`: {m.group(5)} registers, stack {m.group(2)} B, spill stores {m.group(3)} B, loads {m.group(4)} B")


if __name__ == "__main__":
    main()
