"""Turn the .ncu-rep captures in gpurun_out/ into the small tracked summaries under profiles/.
usage: summarize_profiles.py r1   (reads gpurun_out/r1_*.ncu-rep, writes profiles/r1_*_ncu.csv)"""
import csv, glob, os, re, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum']


def raw(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def source_hist(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    lines = out.splitlines()
    rows = list(csv.reader(lines[1:]))
    if not rows:
        return {}, []
    hdr = rows[0]
    iS, iE, iN = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    hist, tot = collections.Counter(), 0
    stalls = collections.Counter()
    for r in rows[1:]:
        if len(r) <= iE:
            continue
        try:
            n = int(r[iE])
        except ValueError:
            continue
        m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[iS].strip())
        hist[(m.group(2) if m else r[iS]).split('.')[0]] += n
        tot += n
        for i in stall_cols:
            try:
                stalls[hdr[i][6:]] += int(r[i] or 0)
            except ValueError:
                pass
    return {k: v for k, v in hist.most_common(14)}, stalls.most_common(8)


def main(tag):
    for path in sorted(glob.glob(os.path.join(ROOT, 'gpurun_out', tag + '_*.ncu-rep'))):
        name = os.path.basename(path)[:-8]
        hdr, units, launches = raw(path)
        dst = os.path.join(ROOT, 'profiles', name + '_ncu.csv')
        with open(dst, 'w', newline='') as f:
            w = csv.writer(f)
            w.writerow(['kernel', 'metric', 'value', 'unit'])
            for vals in launches:
                k = vals[hdr.index('Kernel Name')]
                for key in KEYS:
                    if key in hdr:
                        w.writerow([k, key, vals[hdr.index(key)], units[hdr.index(key)]])
            hist, stalls = source_hist(path)
            for k, v in hist.items():
                w.writerow(['(all launches)', 'sass_warp_instructions.' + k, v, 'inst'])
            for k, v in stalls:
                w.writerow(['(all launches)', 'warp_stall_samples.' + k, v, 'samples'])
        print('wrote', dst)


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else 'r1')
