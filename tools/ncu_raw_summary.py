"""Print selected metrics per kernel from an `ncu --page raw --csv` dump.  Usage: python tools/ncu_raw_summary.py <raw.csv> [regex]"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else
                 r"^(Kernel Name|gpu__time_duration.sum|dram__bytes_read.sum$|dram__bytes_write.sum$|gpu__dram_throughput.avg.pct|"
                 r"sm__pipe_tensor.*cycles_active.avg.pct_of_peak_sustained_(active|elapsed)|sm__warps_active.avg.pct|launch__registers|"
                 r"lts__throughput.avg.pct|smsp__issue_active.avg.pct|sm__inst_executed_pipe_(xu|fma|alu|fmaheavy|fmalite|uniform)\S*pct_of_peak_sustained_active|"
                 r"sm__cycles_elapsed.avg.per_second|launch__grid_size|launch__block_size|launch__shared_mem_per_block_dynamic|"
                 r"smsp__inst_executed.sum$|l1tex__data_pipe_lsu_wavefronts_mem_shared.sum$|sm__throughput.avg.pct|launch__occupancy_limit|sm__ctas_launched)")
hdr, units = rows[0], rows[1]
idx = [i for i, n in enumerate(hdr) if pat.search(n)]
for r in rows[2:]:
    print("----")
    for i in idx:
        print(f"  {hdr[i]:75s} {r[i]:>16s} {units[i]}")
