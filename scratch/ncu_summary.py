"""Key metrics of an ncu report (one kernel) as text: python scratch/ncu_summary.py <rep> [accepted_transitions]"""
import csv, subprocess, sys
rep = sys.argv[1]
acc = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(f"ncu -i {rep} --page raw --csv", shell=True, capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'launch__block_size', 'launch__grid_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__sectors_read.sum', 'dram__sectors_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'lts__t_sectors_srcunit_tex_op_read.sum', 'l1tex__m_xbar2l1tex_read_sectors_mem_lg_op_ld.sum',
        'l1tex__t_sector_hit_rate.pct', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warp_latency_per_inst_issued.ratio', 'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed']
for val in rows[2:]:
    for k in keys:
        if k in hdr:
            i = hdr.index(k)
            extra = ''
            if acc and k in ('dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum', 'dram__sectors_read.sum',
                             'l1tex__m_xbar2l1tex_read_sectors_mem_lg_op_ld.sum', 'lts__t_sectors_srcunit_tex_op_read.sum'):
                try:
                    v = float(val[i].replace(',', ''))
                    mult = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}.get(units[i], 1)
                    extra = f"   ({v * mult / acc:.2f} per accepted transition)"
                except Exception:
                    pass
            print(f"{k:86s} {val[i]:>22s} {units[i]}{extra}")
