#!/usr/bin/env python3
"""Summarise ncu output for profiles/ (development aid).
    ncu_summary.py launches <launches.csv>        per-kernel totals of a `--metrics gpu__time_duration.sum` launch list
    ncu_summary.py full <report.ncu-rep> [regex]  key metrics of every captured launch of a `--set full` report
"""
import collections, csv, re, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "sm__warps_active.avg.per_cycle_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "sm__cycles_elapsed.max"]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        name = re.sub(r"^void ", "", re.sub(r"[<(].*", "", r[ki])).replace("unnamed>::", "")
        v = float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    print(f"# {path}: {sum(cnt.values())} launches, {T:.1f} ms of kernel time (ncu: serialised, cold caches)")
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        if v / T < 0.0005:
            continue
        print(f"{v:10.2f} ms {100 * v / T:5.1f}%  n={cnt[k]:5d}  {k}")


def full(path, pat=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if pat and not re.search(pat, name):
            continue
        print("kernel:", re.sub(r"\(.*", "", name))
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:70s} {r[i]} {units[i]}")


if __name__ == "__main__":
    (launches if sys.argv[1] == "launches" else full)(*sys.argv[2:])
