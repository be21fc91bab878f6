"""Print the handful of ncu metrics that matter from a .ncu-rep (uses `ncu -i ... --page raw --csv`)."""
import csv, subprocess, sys
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active", "sm__inst_executed_pipe_tmem", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__occupancy_limit", "lts__t_sector_hit_rate.pct",
        "sm__cycles_active.avg", "sm__cycles_elapsed.max", "launch__waves_per_multiprocessor", "smsp__cycles_active.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared",
        "smsp__average_warps_issue_stalled", "launch__shared_mem_per_block", "sm__ctas_launched", "lts__t_bytes.sum ", "sm__pipe_tensor_op",
        "smsp__warp_issue_stalled", "launch__cluster"]
for path in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    print("==", path, vals[hdr.index("Kernel Name")][:80])
    for h, u, v in zip(hdr, units, vals):
        if any(w in h for w in WANT) and v not in ("", "0", "n/a"):
            if "stalled" in h and "pct" not in h and "ratio" not in h: continue
            print(f"   {h:95s} {v:>16s} {u}")
