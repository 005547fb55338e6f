"""Per-SASS-line summary of an ncu --import-source report: where the stall samples and shared-memory wavefronts go.
    python tools/ncu_src.py gpurun_out/X.ncu-rep <kernel regex> [top]
"""
import csv, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"], capture_output=True, text=True).stdout
blocks = out.split('"Kernel Name",')
blk = blocks[1]
lines = blk.splitlines()
print("kernel:", lines[0][:120])
rows = list(csv.DictReader(lines[1:]))
def f(r, k):
    try: return float(r[k])
    except Exception: return 0.0
tot_s = sum(f(r, "# Samples") for r in rows); tot_w = sum(f(r, "L1 Wavefronts Shared") for r in rows); tot_x = sum(f(r, "L1 Wavefronts Shared Excessive") for r in rows)
print(f"instructions {len(rows)}, samples {tot_s:.0f}, smem wavefronts {tot_w:.0f}, excessive {tot_x:.0f}")
stall_keys = [k for k in rows[0] if k.startswith("stall_") and "Not Issued" not in k]
agg = {k: sum(f(r, k) for r in rows) for k in stall_keys}
print("stall totals:", {k: int(v) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0})
# opcode class totals
cls = {}
for r in rows:
    op = r["Source"].split()[0] if r["Source"].split() else "?"
    if op.startswith("@"): op = r["Source"].split()[1]
    op = op.split(".")[0]
    c = cls.setdefault(op, [0, 0.0, 0.0, 0.0])
    c[0] += 1; c[1] += f(r, "# Samples"); c[2] += f(r, "Instructions Executed"); c[3] += f(r, "L1 Wavefronts Shared")
print(f"{'op':10s} {'n':>5s} {'samples':>8s} {'share':>6s} {'inst_exec':>12s} {'smem_wf':>10s}")
for op, c in sorted(cls.items(), key=lambda kv: -kv[1][1])[:18]:
    print(f"{op:10s} {c[0]:5d} {c[1]:8.0f} {c[1]/tot_s*100:5.1f}% {c[2]:12.0f} {c[3]:10.0f}")
print("top excessive-wavefront instructions:")
for r in sorted(rows, key=lambda r: -f(r, "L1 Wavefronts Shared Excessive"))[:top]:
    if f(r, "L1 Wavefronts Shared Excessive") <= 0: break
    print(f"  {r['Source'][:60]:60s} wf {f(r,'L1 Wavefronts Shared'):9.0f} ideal {f(r,'L1 Wavefronts Shared Ideal'):9.0f} excess {f(r,'L1 Wavefronts Shared Excessive'):9.0f}")
