"""Summarise an ncu report: python tools/ncu_summary.py raw.csv [src.csv [kernel_index]]"""
import collections
import csv
import sys


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
            'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
            'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
            'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
            'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
            'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
            'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
            'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum',
            'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum',
            'smsp__warps_eligible.avg.per_cycle_active', 'smsp__warps_active.avg.per_cycle_active',
            'sm__cycles_elapsed.avg', 'sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_elapsed',
            'lts__throughput.avg.pct_of_peak_sustained_elapsed',
            'l1tex__throughput.avg.pct_of_peak_sustained_elapsed']
    for r in rows[2:]:
        print('----', r[idx['Kernel Name']][:90])
        for w in want:
            if w in idx:
                print('  %-78s %s %s' % (w, r[idx[w]], units[idx[w]]))
        for h in hdr:
            if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio'):
                v = float(r[idx[h]])
                if v > 0.15:
                    print('     stall %-30s %.2f' % (h.replace('smsp__average_warps_issue_stalled_', '')
                                                     .replace('_per_issue_active.ratio', ''), v))


def src(path, which, topn=40):
    rows = list(csv.reader(open(path)))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == 'Kernel Name':
            cur = {'name': r[1], 'rows': []}
            blocks.append(cur)
        elif r and r[0] == 'Address':
            cur['hdr'] = r
        elif cur is not None and r:
            cur['rows'].append(r)
    b = blocks[which]
    h = b['hdr']
    ix = {n: i for i, n in enumerate(h)}
    cols = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
    tot = sum(int(r[ix['# Samples']]) for r in b['rows'])
    print('==== source page kernel', which, b['name'][:80], 'instrs', len(b['rows']), 'samples', tot)
    agg = collections.Counter()
    byop = collections.Counter()
    for r in b['rows']:
        for c in cols:
            agg[c] += int(r[ix[c]])
        srcs = r[ix['Source']].split()
        op = srcs[1] if srcs[0].startswith('@') else srcs[0]
        byop[op.split('.')[0]] += int(r[ix['# Samples']])
    print(agg.most_common(12))
    print(byop.most_common(14))
    top = sorted(enumerate(b['rows']), key=lambda t: -int(t[1][ix['# Samples']]))[:topn]
    for i, r in sorted(top):
        st = {c.replace('stall_', ''): int(r[ix[c]]) for c in cols if int(r[ix[c]]) > 0}
        print(i, r[ix['Source']].strip()[:62].ljust(62), r[ix['# Samples']].rjust(4),
              sorted(st.items(), key=lambda t: -t[1])[:2])


if __name__ == '__main__':
    raw(sys.argv[1])
    if len(sys.argv) > 2:
        src(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
