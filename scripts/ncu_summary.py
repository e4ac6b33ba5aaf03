"""Per-kernel summary of selected counters of an `ncu --set full` report (read here, no GPU needed):

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep "header line" > profiles/<name>.summary.txt
"""
import csv
import subprocess
import sys

WANT = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "launch__grid_size",
        "launch__registers_per_thread", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__cycles_elapsed.avg", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
if len(sys.argv) > 2:
    print(sys.argv[2])
tot_r = tot_w = 0.0
for r in rows[2:]:
    print("=====", r[hdr.index("Kernel Name")][:90])
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"  {w} [{units[i]}] = {r[i]}")
    tot_r += float(r[hdr.index("dram__bytes_read.sum")])
    tot_w += float(r[hdr.index("dram__bytes_write.sum")])
u = units[hdr.index("dram__bytes_read.sum")]
print(f"===== total over the {len(rows) - 2} launches: dram read {tot_r:.1f} {u}, write {tot_w:.1f} {u}, sum {tot_r + tot_w:.1f} {u}")
