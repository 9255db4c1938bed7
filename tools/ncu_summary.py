"""Summarise ncu output for profiles/ (read here, on the CPU box).

    python tools/ncu_summary.py launches gpurun_out/x_launches.csv  > profiles/x_launch_summary.txt
    python tools/ncu_summary.py full     gpurun_out/x.ncu-rep       > profiles/x_top_kernels.txt
"""
import collections
import csv
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
           "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]


def short(name):
    return name.split("(")[0]


def launches(path):
    rows = [l for l in open(path) if l.startswith('"')]
    agg = collections.OrderedDict()
    for r in csv.DictReader(rows):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        a = agg.setdefault(short(r["Kernel Name"]), [0, 0.0])
        a[0] += 1
        a[1] += float(r["Metric Value"].replace(",", "")) / 1e6
    total = sum(a[1] for a in agg.values())
    print("# per-launch times are cold-cache and serialised: compare SHARES; raw list:", path.split("/")[-1])
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:<48s} launches={a[0]:3d} total={a[1]:8.3f} ms avg={a[1] / a[0]:8.4f} ms share={100 * a[1] / total:5.1f}%")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    seen = set()
    for r in rows[2:]:
        name = short(r[hdr.index("Kernel Name")])
        if name in seen:
            continue
        seen.add(name)
        print(f"## {name}")
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(f"{m:<90s} {r[i]} {units[i]}")
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
