"""Selected raw metrics of the first kernel in an .ncu-rep (ncu -i X --page raw --csv) as `metric,unit,value` lines - the
format of the profiles/*_ncu_full_*.csv files.  Usage: python tools/ncu_summary.py report.ncu-rep "comment line" > out.csv"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__cycles_active.avg", "launch__grid_size", "launch__cluster_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.max.pct_of_peak_sustained_elapsed", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
        "l1tex__m_l1tex2xbar_write_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum.per_second",
        "smsp__mem_tensor_reads_op_ldt.sum", "smsp__mem_tensor_reads_op_utcmma_matrix_a_sp_sf.sum", "smsp__mem_tensor_reads_op_utcmma_matrix_c.sum",
        "smsp__mem_tensor_writes_op_stt.sum", "smsp__mem_tensor_writes_op_utcmma.sum", "smsp__sass_inst_executed_op_utcmma.sum",
        "smsp__sass_inst_executed_op_tmem_ldt.sum", "smsp__sass_inst_executed_op_tmem_stt.sum"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
print("# " + (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]))
print("# kernel: " + vals[hdr.index("Kernel Name")])
print("metric,unit,value")
for w in WANT:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w},{units[i]},{vals[i]}")
