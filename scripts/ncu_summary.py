"""Summarise an .ncu-rep (read here, no GPU needed): python scripts/ncu_summary.py file.ncu-rep [kernel-index]"""
import csv, subprocess, sys
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput", "sm__throughput.avg.pct", "sm__pipe_tensor", "sm__inst_executed_pipe_xu", "pipe_xu", "pipe_fma", "pipe_alu",
        "pipe_lsu", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit",
        "smsp__issue_active.avg.pct", "issue_stalled", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "launch__grid_size", "launch__block_size", "smsp__cycles_active.avg", "tmem", "sm__inst_executed_pipe_uniform"]
for r in rows[2:]:
    print("=== kernel:", r[hdr.index("Kernel Name")][:80], " grid", r[hdr.index("Grid Size")], " block", r[hdr.index("Block Size")])
    out = []
    for i, h in enumerate(hdr):
        if any(k in h for k in KEYS):
            v = r[i]
            if "issue_stalled" in h and "pct" not in h and "ratio" not in h:
                continue
            out.append((h, v, units[i]))
    for h, v, u in out:
        try:
            fv = float(v.replace(",", ""))
            if fv == 0:
                continue
        except ValueError:
            pass
        print(f"  {h[:100]:100s} {v:>16s} {u}")
