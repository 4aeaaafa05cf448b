"""Prints the headline metrics of an .ncu-rep (one line per profiled launch)."""
import subprocess
import sys
import csv
import io

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "smsp__inst_executed.sum", "sm__inst_executed_pipe_lsu.sum", "l1tex__t_bytes.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d.get("Kernel Name", "")[:80], "grid", d.get("Grid Size"), "block", d.get("Block Size"))
    for k in KEYS:
        if k in d:
            print(f"   {k:75s} {d[k]:>16s} {units[hdr.index(k)]}")
    if "--stalls" in sys.argv:   # warp-state shares: where the resident warps spend their cycles
        st = [(k, float(d[k].replace(",", ""))) for k in hdr if k.startswith("smsp__average_warps_issue_stalled_") and
              k.endswith("_per_issue_active.ratio") and d.get(k) not in (None, "", "n/a")]
        for k, x in sorted(st, key=lambda kv: -kv[1])[:8]:
            print(f"   stall {k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:40s} {x:8.2f} warps per issue")
