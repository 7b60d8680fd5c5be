"""Prints the metrics we track from an ncu report:  python tools/ncu_summary.py <file.ncu-rep> [more...]"""
import csv
import io
import subprocess
import sys

WANT = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed_pipe_fp64.sum', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'launch__waves_per_multiprocessor']


def main():
    for path in sys.argv[1:]:
        out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr = rows[0]
        stall = [i for i, h in enumerate(hdr) if 'warp_issue_stalled' in h and h.endswith('per_warp_active.pct')]
        print('==', path)
        for r in rows[2:]:
            for w in WANT:
                if w in hdr:
                    print('  %-72s %s' % (w, r[hdr.index(w)][:70]))
            st = sorted(((float(r[i] or 0), hdr[i]) for i in stall), reverse=True)[:7]
            for v, h in st:
                print('  stall %-66s %.1f' % (h.replace('smsp__warp_issue_stalled_', '').replace('_per_warp_active.pct', ''), v))
            print()


if __name__ == '__main__':
    main()
