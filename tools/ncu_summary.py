"""Summarise ncu output brought back in gpurun_out/ into small text files for profiles/.

    python tools/ncu_summary.py launches gpurun_out/X_launches.csv  > profiles/rNN_launches.txt
    python tools/ncu_summary.py full     gpurun_out/X_prof.ncu-rep  > profiles/rNN_ncu_full.txt

`launches`: per-kernel count / total / mean device time and share of all profiled launches (cold-cache, serialised:
shares are meaningful, absolutes are not).  `full`: the metrics DESIGN.md and bench.py quote, per profiled launch.
"""

import csv
import subprocess
import sys
from collections import OrderedDict

FULL_METRICS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread",
    "launch__block_size",
    "launch__grid_size",
    "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem",
    "smsp__cycles_elapsed.avg.per_second",
    "sm__cycles_elapsed.max",
    "smsp__sass_inst_executed_op_local_ld.sum",
    "smsp__sass_inst_executed_op_local_st.sum",
]


def launches(path):
    rows = []
    with open(path) as fh:
        lines = [ln for ln in fh if not ln.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((r["Kernel Name"], float(r["Metric Value"])))
    agg = OrderedDict()
    for name, ns in rows:
        short = name.split("(")[0].replace("void ", "")
        c = agg.setdefault(short, [0, 0.0])
        c[0] += 1
        c[1] += ns
    total = sum(v[1] for v in agg.values())
    print(f"# {len(rows)} launches, {total / 1e3:.1f} us total device time (ncu: cold cache, serialised)")
    print(f"{'kernel':70s} {'n':>5s} {'total_us':>10s} {'mean_us':>9s} {'share':>7s}")
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:70]:70s} {n:5d} {ns / 1e3:10.1f} {ns / n / 1e3:9.2f} {ns / total * 100:6.1f}%")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("==== " + r[idx["Kernel Name"]])
        for m in FULL_METRICS:
            if m in idx:
                print(f"  {m:70s} {r[idx[m]]:>16s} {units[idx[m]]}")
        if "dram__bytes_read.sum" in idx:
            def val(m):
                v, u = float(r[idx[m]].replace(",", "")), units[idx[m]]
                return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
            print(f"  {'traffic = dram read + write (bytes)':70s} {val('dram__bytes_read.sum') + val('dram__bytes_write.sum'):16.0f}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
