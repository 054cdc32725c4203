#!/usr/bin/env python
"""Per-source-line instruction counts of one kernel of an .ncu-rep captured with --import-source on.
usage: ncu_lines.py report.ncu-rep <kernel-id> [top]"""
import csv
import subprocess
import sys

rep, kid = sys.argv[1], sys.argv[2]


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda",
                      "--kernel-id", f":::{kid}"] if False else
                     ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# the report prints one block per (kernel launch, file); keep blocks in order and number launches
blocks, cur = [], None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = {"file": r[1], "func": None, "hdr": None, "rows": []}
        blocks.append(cur)
    elif r[0] == "Function Name":
        cur["func"] = r[1]
    elif r[0] == "Line No":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None and r[0].isdigit():
        cur["rows"].append(r)
launch, last_func, seen_files = -1, None, set()
for b in blocks:
    if b["func"] != last_func or b["file"] in seen_files:
        launch += 1
        seen_files = set()
        last_func = b["func"]
    seen_files.add(b["file"])
    b["launch"] = launch
sel = [b for b in blocks if str(b["launch"]) == kid]
tot = 0
lines = []
for b in sel:
    h = b["hdr"]
    ii, si = h.index("Instructions Executed"), h.index("# Samples")
    for r in b["rows"]:
        n = num(r[ii])
        tot += n
        lines.append((n, num(r[si]), b["file"].split("/")[-1], r[0], r[1].strip()[:110]))
print(sel[0]["func"] if sel else "?", "total warp instructions", tot)
by_samples = len(sys.argv) > 4 and sys.argv[4] == "smp"
print("total samples", sum(l[1] for l in lines))
for n, s, f, ln, src in sorted(lines, key=(lambda l: -l[1]) if by_samples else (lambda l: -l[0]))[:top]:
    print(f"{n:>10} {100.0 * n / max(tot, 1):5.1f}% smp {s:>6}  {f}:{ln}  {src}")
