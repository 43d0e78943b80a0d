"""Refresh profiles/ from the artefacts a gpurun call left in gpurun_out/ (run in the build container).

Expects: bench.json, bench_ref.json, launches.csv, prof_final.ncu-rep, aux_timing.json, sweep_n1.json
"""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, 'gpurun_out')
P = os.path.join(ROOT, 'profiles')
R = sys.argv[1] if len(sys.argv) > 1 else 'r1'


def main():
    os.makedirs(P, exist_ok=True)
    for src, dst in (('bench.json', '%s_bench_n1.json'), ('bench_ref.json', '%s_bench_reference_arm.json'),
                     ('aux_timing.json', '%s_aux_kernels.json'), ('sweep_n1.json', '%s_sweep_n1.json')):
        if os.path.exists(os.path.join(G, src)):
            shutil.copy(os.path.join(G, src), os.path.join(P, dst % R))
    rep = os.path.join(G, 'prof_final.ncu-rep')
    have_csv = os.path.exists(os.path.join(G, 'raw_final.csv')) and os.path.exists(os.path.join(G, 'src_final.csv'))
    if os.path.exists(rep) or have_csv:
        if os.path.exists(rep):
            raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
            src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'],
                                 capture_output=True, text=True).stdout
            open(os.path.join(G, 'raw_final.csv'), 'w').write(raw)
            open(os.path.join(G, 'src_final.csv'), 'w').write(src)
        else:       # the .ncu-rep (> 40 MB) stayed on the GPU box; its two CSV pages were exported there
            raw = open(os.path.join(G, 'raw_final.csv')).read()
        out = ''
        for k in (0, 1):
            res = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'ncu_summary.py'),
                                  os.path.join(G, 'raw_final.csv'), os.path.join(G, 'src_final.csv'), str(k)],
                                 capture_output=True, text=True).stdout
            out += res if k == 0 else res[res.index('==== source'):]
        open(os.path.join(P, '%s_strip_kernel_ncu_full.txt' % R), 'w').write(
            '# ncu --set full --clock-control none --import-source on -k regex:dc_strip_pipev -s 6 -c 2 '
            'python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-recnet --no-refinement\n' + out)
        rows = list(csv.reader(raw.splitlines()))
        idx = {h: i for i, h in enumerate(rows[0])}
        tr = {}
        for name, r in zip(('forward', 'adjoint'), rows[2:4]):
            def val(key):
                v, u = float(r[idx[key]]), rows[1][idx[key]]
                return int(v * {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1}[u])
            tr[name] = {'read': val('dram__bytes_read.sum'), 'write': val('dram__bytes_write.sum')}
        tr['forward']['algorithmic'] = 24 * 256 * 256 * 256
        tr['adjoint']['algorithmic'] = 16 * 256 * 256 * 256
        tr['bytes_per_fwd_adj_pair'] = sum(tr[k][d] for k in ('forward', 'adjoint') for d in ('read', 'write'))
        tr['source'] = 'ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch'
        tr['note'] = ('reads equal the algorithmic bytes (no re-reads); writes are below them because the '
                      'tail of the output is still dirty in the 126 MB L2 when the kernel ends')
        json.dump(tr, open(os.path.join(P, '%s_dram_traffic.json' % R), 'w'), indent=1)
    lc = os.path.join(G, 'launches.csv')
    if os.path.exists(lc):
        rows, hdr, out = list(csv.reader(open(lc))), None, []
        for r in rows:
            if r and r[0] == 'ID':
                hdr = r
            elif hdr and len(r) == len(hdr):
                d = dict(zip(hdr, r))
                name = re.sub(r'void at::.*', 'at:: RNG / elementwise kernel (synthetic data setup)', d['Kernel Name'])
                out.append((d['ID'], name[:150], d['Block Size'], d['Grid Size'], d['Metric Value']))
        with open(os.path.join(P, '%s_launches_bench.csv' % R), 'w') as f:
            f.write('# ncu --metrics gpu__time_duration.sum --clock-control none -c 300 python bench.py '
                    '--steps 3 --warmup 3 --no-cpu-baseline --no-recnet --no-refinement\n# cold-cache, serialised launches: '
                    'compare SHARES, not absolutes\nid,kernel,block,grid,duration_ns\n')
            for o in out:
                f.write(','.join('"%s"' % x if ',' in x else x for x in o) + '\n')
        agg, cnt = collections.Counter(), collections.Counter()
        for o in out:
            k = o[1].split('<')[0].split('(')[0].replace('void ', '')
            agg[k] += int(o[4])
            cnt[k] += 1
        tot = sum(agg.values())
        with open(os.path.join(P, '%s_launches_summary.txt' % R), 'w') as f:
            f.write('share of summed kernel time over the first 300 launches of `bench.py --steps 3 --warmup 3` '
                    '(setup, prepare, warm-up, timed steps, per-kernel timing loops)\n')
            for k, v in agg.most_common():
                f.write('%-45s n=%3d %9.1f us %5.1f %%\n' % (k, cnt[k], v / 1e3, 100 * v / tot))
        print(open(os.path.join(P, '%s_launches_summary.txt' % R)).read())


if __name__ == '__main__':
    main()
