#!/usr/bin/env python3
"""Turn ncu output brought back from the GPU box (gpurun_out/) into the text summaries committed under profiles/.

  tools/ncu_summary.py full <report.ncu-rep> [<kernel-name substring> ...]   key metrics + top stalls per captured launch
  tools/ncu_summary.py launches <launches.csv>                                per-kernel totals and shares of a launch list
                                                                              (ncu --metrics gpu__time_duration.sum --csv)
"""
import csv
import subprocess
import sys
from collections import OrderedDict

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
]


def full(report, wanted):
    raw = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    unit = dict(zip(hdr, units))
    seen = set()
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d.get("Kernel Name", "")
        if wanted and not any(w in name for w in wanted):
            continue
        if name in seen:  # one launch per kernel is enough
            continue
        seen.add(name)
        print("== " + name[:110])
        for m in METRICS:
            if m in d and d[m] not in ("", "n/a"):
                print("  %s: %s %s" % (m, d[m], unit.get(m, "")))
        try:
            ld = float(d["l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"].replace(",", ""))
            rq = float(d["l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"].replace(",", ""))
            print("  => sectors per load request: %.2f" % (ld / rq))
        except (KeyError, ValueError, ZeroDivisionError):
            pass
        stalls = {}
        for k, v in d.items():
            if k.startswith("smsp__average_warp") and k.endswith("per_issue_active.ratio"):
                try:
                    stalls[k.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")] = float(v.replace(",", ""))
                except ValueError:
                    pass
        top = sorted(stalls.items(), key=lambda kv: -kv[1])[:6]
        print("  top stalls (warps per issue): " + ", ".join("%s=%.2f" % kv for kv in top))
        print()


def launches(path):
    tot = OrderedDict()
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"].split("(")[0]
        ns = float(r["Metric Value"].replace(",", ""))
        if r.get("Metric Unit") == "us":
            ns *= 1e3
        t = tot.setdefault(name, [0, 0.0])
        t[0] += 1
        t[1] += ns
    total = sum(v[1] for v in tot.values())
    for name, (n, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print("%-62s launches=%4d total_us=%10.1f share=%.3f" % (name[:62], n, ns / 1e3, ns / total))


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "full":
        full(sys.argv[2], sys.argv[3:])
    elif len(sys.argv) == 3 and sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        sys.exit(__doc__)
