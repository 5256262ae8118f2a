"""Print the roofline-relevant counters of an `ncu --page raw --csv` export."""
import csv
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_lsu.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']


def main(path, every=1):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    name_i = hdr.index('Kernel Name')
    seen = {}
    for r in rows[2:]:
        seen.setdefault(r[name_i], r)          # first launch of each kernel
    for name, r in seen.items():
        print('==', name[:100])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print('   %-82s %s %s' % (w, r[i], units[i]))


if __name__ == '__main__':
    main(sys.argv[1])
