"""Selected metrics of an `ncu -i <rep> --page raw --csv` dump read from stdin (one kernel per row)."""
import csv, sys
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sectors.sum", "lts__t_sectors_op_read.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "local_load_requests", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]
rows = list(csv.reader(sys.stdin))
hdr = None
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr = r; units = rows[i + 1]; data = rows[i + 2:]; break
if hdr is None:
    print("no kernel rows"); sys.exit(0)
for d in data:
    if len(d) != len(hdr):
        continue
    m = dict(zip(hdr, d)); u = dict(zip(hdr, units))
    print("# kernel", m.get("Kernel Name", "?")[:90])
    for k in WANT:
        if k in m:
            print(f"{k:78s} {m[k]} {u.get(k, '')}")
