"""Turn ncu outputs brought back in gpurun_out/ into the small tracked summaries under profiles/.

  python profiles/summarize.py launches gpurun_out/launches_r01.csv profiles/r01_launches.csv
      (input: `ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ... python bench.py ...`)
  python profiles/summarize.py full gpurun_out/prof.ncu-rep profiles/r01_ncu_full_summary.csv
      (input: `ncu --set full --clock-control none --import-source on -o ... python bench.py ...`)
"""
import collections
import csv
import subprocess
import sys

KEEP = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_issued.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__inst_executed.sum", "l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum",
    "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
]
STALL = "smsp__average_warps_issue_stalled_"


def launches(src, dst):
    lines = [l for l in open(src) if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    for r in rows:
        k = (r["Kernel Name"], r["Grid Size"], r["Block Size"])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r["Metric Value"])
    total = sum(a[1] for a in agg.values())
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "grid", "block", "launches", "avg_us", "total_us", "share_of_all_launch_time"])
        for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            w.writerow([k[0], k[1], k[2], n, "%.2f" % (t / n / 1e3), "%.1f" % (t / 1e3), "%.4f" % (t / total)])


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [i for i, h in enumerate(hdr) if h in KEEP or (h.startswith(STALL) and h.endswith("_per_issue_active.ratio"))]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i].replace(STALL, "stall_").replace("_per_issue_active.ratio", "") for i in cols])
        w.writerow([units[i] for i in cols])
        for r in rows[2:]:
            w.writerow([r[i] for i in cols])


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
