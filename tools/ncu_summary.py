"""Prints the key metrics of every kernel in an `ncu --page raw --csv` export."""
import csv
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.sum',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.sum', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum']
STALL = 'smsp__average_warp'


def main(path, name_filter=None):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    for d in data:
        name = d[idx['Kernel Name']]
        if name_filter and name_filter not in name:
            continue
        print('=' * 110)
        print(name[:110], ' id', d[idx['ID']])
        for w in WANT:
            if w in idx:
                print(f"  {w:75s} {d[idx[w]]:>18s} {units[idx[w]]}")
        stalls = [(h, d[i]) for h, i in idx.items() if 'issue_stalled' in h and h.endswith('_per_warp_active.pct')]
        stalls = sorted(((h, float(v.replace(',', ''))) for h, v in stalls if v not in ('', 'n/a')), key=lambda x: -x[1])[:6]
        for h, v in stalls:
            print(f"  stall {h.replace('smsp__warp_issue_stalled_', '').replace('_per_warp_active.pct', ''):40s} {v:8.1f}")


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
