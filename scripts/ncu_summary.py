"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/.

    python scripts/ncu_summary.py launches gpurun_out/launches_r1.csv > profiles/r01_launches.txt
    python scripts/ncu_summary.py raw gpurun_out/prof_conv_r1.ncu-rep > profiles/r01_conv_full.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_active.avg", "sm__cycles_elapsed.max", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max.per_second")


def launches(path):
    with open(path) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}[row["Metric Unit"]]
        a = agg[row["Kernel Name"][:90]]
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print("# per-kernel device time from `ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare shares)")
    print("%-92s %6s %12s %10s %7s" % ("kernel", "n", "total_us", "avg_us", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:20]:
        print("%-92s %6d %12.1f %10.2f %6.1f%%" % (k, v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
    print("%-92s %6d %12.1f" % ("TOTAL", sum(v[0] for v in agg.values()), tot))


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print("# selected metrics of `ncu --set full --clock-control none` capture %s (one column per captured launch)" % path)
    for i, h in enumerate(hdr):
        if h in KEEP or h == "Kernel Name":
            vals = [r[i][:60] for r in data]
            print("%-75s %-14s %s" % (h, units[i], "  ".join(vals)))


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
