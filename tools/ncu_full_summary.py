"""Selected columns of an `ncu --set full` report (read in the build container: ncu -i REP --page raw --csv) -> a small csv
for profiles/.  Usage: python tools/ncu_full_summary.py REP.ncu-rep OUT.csv "comment line"."""
import csv, io, subprocess, sys
rep, out, comment = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_barrier.ratio",
        "smsp__average_warp_latency_issue_stalled_wait.ratio", "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio",
        "smsp__average_warp_latency_issue_stalled_sleeping.ratio"]
cols = [i for i, h in enumerate(hdr) if any(h == w or h.endswith("." + w) or w in h for w in want)]
with open(out, "w", newline="") as f:
    if comment:
        f.write(f'"# {comment}"\n')
    w = csv.writer(f)
    w.writerow([hdr[i] for i in cols])
    w.writerow([units[i] for i in cols])
    for r in rows[2:]:
        w.writerow([r[i] if i < len(r) else "" for i in cols])
print(f"{out}: {len(rows) - 2} kernels, {len(cols)} columns")
