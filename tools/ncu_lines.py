"""Attribute ncu warp-stall samples to CUDA source lines (ncu's CSV source page only exports SASS).

    python tools/ncu_lines.py <report.ncu-rep> <kernel regex> <cubin> <mangled function substring> [top]

Joins `ncu --page source --csv` (one row per SASS instruction, in address order) with `nvdisasm -gi` line info of the
same function from the cubin (`cuobjdump -xelf all libadept_b200.so`) by instruction index.
"""
import csv
import io
import re
import subprocess
import sys
from collections import Counter

rep, kregex, cubin, fsub = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kregex}",
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
c_samp = hdr.index("# Samples")
c_src = hdr.index("Source")
sass = []
for r in rows[hi + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"):
        break
    if len(r) > c_samp:
        sass.append((r[c_src].strip(), int(r[c_samp] or 0)))

dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and fsub in l)
instr_lines = []
cur = ("?", 0)
chain = []
pat_file = re.compile(r'//## File "([^"]+)", line (\d+)(.*)')
pat_ins = re.compile(r"^\s+/\*[0-9a-f]{4,}\*/\s+(.*?);")
for l in dis[start + 1:]:
    if l.startswith(".text.") or l.startswith(".section"):
        break
    m = pat_file.search(l)
    if m:
        if "inlined at" in m.group(3):
            chain.append((m.group(1).split("/")[-1], int(m.group(2))))
        else:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            chain = []
        continue
    m = pat_ins.match(l)
    if m:
        instr_lines.append((cur, tuple(chain), m.group(1)))

print(f"# sass instrs (ncu) {len(sass)}, (nvdisasm) {len(instr_lines)}")
n = min(len(sass), len(instr_lines))
by_line, by_outer, by_op = Counter(), Counter(), Counter()
total = sum(s for _, s in sass)
for (src, samp), (cur, chain, txt) in zip(sass[:n], instr_lines[:n]):
    by_line[cur] += samp
    outer = chain[-1] if chain else cur
    by_outer[outer] += samp
    op = txt.split()[0] if not txt.startswith("@") else txt.split()[1]
    by_op[op.split(".")[0]] += samp
print(f"# total samples {total}")
print("## by innermost source line")
for (f, ln), s in by_line.most_common(top):
    print(f"{s:7d} {100*s/total:5.1f}%  {f}:{ln}")
print("## by outermost (kernel-level) line")
for (f, ln), s in by_outer.most_common(top):
    print(f"{s:7d} {100*s/total:5.1f}%  {f}:{ln}")
print("## by opcode")
for op, s in by_op.most_common(20):
    print(f"{s:7d} {100*s/total:5.1f}%  {op}")
