"""Key metrics of an `ncu --set full` report: python scripts/ncu_summary.py profiles/x.ncu-rep  (needs ncu; no GPU)."""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__cycles_elapsed.max", "smsp__warps_active.avg.per_cycle_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sectors_srcunit_tex_op_read.sum"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[-1]
name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
print("kernel: `%s`\n" % name[:120])
for w in WANT:
    if w in hdr:
        i = hdr.index(w)
        print("* `%s` = %s %s" % (w, vals[i], units[i]))
