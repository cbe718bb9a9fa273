"""Key metrics of an ncu report, one line per kernel launch: python tools/ncu_summary.py <report.ncu-rep>"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
K = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("Block Size", "block"), ("gpu__time_duration.sum", "dur_us"),
     ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "dyn_smem_KB"),
     ("dram__bytes_read.sum", "dram_rd_MB"), ("dram__bytes_write.sum", "dram_wr"),
     ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
     ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
     ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
     ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%_of_active"),
     ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
     ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
     ("sm__cycles_active.avg", "sm_cycles_active_avg"), ("smsp__inst_executed.sum", "warp_insts")]
units = rows[1]
for r in rows[2:]:
    parts = []
    for k, short in K:
        if k in hdr:
            v = r[hdr.index(k)]
            u = units[hdr.index(k)]
            if short == "kernel":
                v = v.split("(")[0].replace("void ", "")[:48]
            parts.append(f"{short}={v}{(' ' + u) if short in ('dram_wr',) else ''}")
    print("  ".join(parts))
