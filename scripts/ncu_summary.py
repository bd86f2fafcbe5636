"""Key metrics + top stall reasons of an .ncu-rep (run where ncu is installed; no GPU needed)."""
import csv, subprocess, sys
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fp64.sum', 'sm__inst_executed_pipe_uniform.sum', 'smsp__cycles_active.avg']
def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print('kernel:', vals[hdr.index('Kernel Name')][:60])
        for k in KEYS:
            if k in hdr:
                print(f'  {k} = {vals[hdr.index(k)]} {units[hdr.index(k)]}')
        st = []
        for i, k in enumerate(hdr):
            if 'issue_stalled' in k and k.endswith('per_warp_active.pct') and 'not_issued' not in k:
                try: st.append((float(vals[i]), k.replace('smsp__warp_issue_stalled_', '').replace('_per_warp_active.pct', '')))
                except ValueError: pass
        print('  stalls:', ', '.join(f'{k} {v:.0f}' for v, k in sorted(st, reverse=True)[:8]))
if __name__ == '__main__':
    for p in sys.argv[1:]: main(p)
