"""Key figures of an ncu report (first kernel): duration, issue utilisation, stall reasons per issued instruction, FP64 pipe,
instruction-cache hit rate, DRAM traffic.  usage: python scripts/ncu_summary.py <report.ncu-rep>"""
import csv, subprocess, sys
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(txt.split("\n")))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (v, u) for h, v, u in zip(hdr, vals, units)}
for k in ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.per_cycle_active",
          "sm__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__icc_request_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
          "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_active.avg",
          "smsp__inst_executed.sum"]:
    for h in hdr:
        if h == k:
            print(f"{h:75s} {d[h][0]:>16s} {d[h][1]}")
print("stall cycles per issued instruction:")
for h in hdr:
    if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and float(d[h][0]) >= 0.05:
        print(f"   {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:28s} {float(d[h][0]):6.2f}")
