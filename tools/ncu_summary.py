#!/usr/bin/env python
"""Turns ncu output brought back in gpurun_out/ into the small text summaries committed under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv            > profiles/rNN_launches.txt
    python tools/ncu_summary.py full     gpurun_out/prof.ncu-rep            > profiles/rNN_full.txt
"""
import collections
import csv
import io
import subprocess
import sys

FULL = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum",
        "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_xu.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__maximum_warps_per_active_cycle_pct"]


def launches(path):
    rows = list(csv.reader(open(path)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[h]
    agg = collections.OrderedDict()
    for r in rows[h + 1:]:
        d = dict(zip(H, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(d["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(d["Metric Unit"], 1.0)
        agg.setdefault(d["Kernel Name"].split("(")[0][:48], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"{'kernel':48s} {'launches':>8s} {'mean us':>10s} {'total us':>11s} {'share':>7s}")
    for k, v in agg.items():
        print(f"{k:48s} {len(v):8d} {sum(v)/len(v):10.1f} {sum(v):11.1f} {100*sum(v)/tot:6.1f}%")


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    H, U = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(H, r))
        print(f"== {d['Kernel Name'].split('(')[0]}  (launch id {d['ID']})")
        for m in FULL:
            if m in d:
                print(f"   {m:70s} {d[m]:>18s} {U[H.index(m)]}")
        # the five largest warp-stall reasons (warps waiting on the reason per issued instruction)
        st = []
        for m in H:
            if m.startswith("smsp__average_warps_issue_stalled_") and m.endswith("_per_issue_active.ratio"):
                try:
                    st.append((float(d[m].replace(",", "")), m))
                except ValueError:
                    pass
        for v, m in sorted(st, reverse=True)[:5]:
            print(f"   stall {m[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:63s} {v:18.2f} warps per issue")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
