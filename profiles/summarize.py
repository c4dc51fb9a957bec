"""Turns gpurun_out captures into the small text summaries committed under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_r01.csv  > profiles/r01_launches.txt
    python profiles/summarize.py ncu      gpurun_out/prof_r01b.ncu-rep > profiles/r01_ncu_fwd_bwd.txt
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed.sum",
]


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hdr]
    ik, iv = h.index("Kernel Name"), h.index("Metric Value")
    agg = OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) < len(h):
            continue
        name = r[ik].split("(")[0].replace("void ", "")
        agg.setdefault(name, []).append(float(r[iv].replace(",", "")) / 1000.0)
    total = sum(sum(v) for v in agg.values())
    print("kernel launches under `ncu --metrics gpu__time_duration.sum --clock-control none` (cold cache, serialised:")
    print("compare SHARES, not absolutes).  us = microseconds per launch.")
    print(f"{'kernel':70s} {'n':>4s} {'mean us':>9s} {'share':>7s}")
    for k, v in agg.items():
        print(f"{k[:70]:70s} {len(v):4d} {sum(v)/len(v):9.2f} {100*sum(v)/total:6.1f}%")


def ncu(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units = rows[0], rows[1]
    for r in rows[2:]:
        print("=" * 100)
        print(r[h.index("Kernel Name")][:100])
        for k in KEYS:
            if k in h:
                i = h.index(k)
                print(f"  {k:72s} {r[i]:>18s} {units[i]}")
        i_r, i_w = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
        print(f"  dram traffic per launch (read + write)                                    "
              f"{float(r[i_r].replace(',', '')) + float(r[i_w].replace(',', '')):18.3f} {units[i_r]}")


if __name__ == "__main__":
    {"launches": launches, "ncu": ncu}[sys.argv[1]](sys.argv[2])
