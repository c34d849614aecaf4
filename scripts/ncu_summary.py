"""Print the key metrics of an .ncu-rep (raw page) -- used to fill profiles/."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, u = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.sum", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio"]
for d in rows[2:]:
    print("==", d[h.index("Kernel Name")][:80])
    for k in keys:
        if k in h:
            i = h.index(k); print(f"  {k:78s} {u[i]:14s} {d[i]}")
    for i, n in enumerate(h):
        if "warp_issue_stalled" in n and n.endswith("_per_warp_active.pct"):
            try:
                v = float(d[i])
            except ValueError:
                continue
            if v > 3: print(f"  stall {n.replace('smsp__warp_issue_stalled_','').replace('_per_warp_active.pct',''):40s} {v:.1f}%")
