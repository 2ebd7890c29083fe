"""Summarise ncu reports brought back from the GPU box into small JSON files for profiles/ (runs without a GPU).
  python tests/ncu_summarise.py full <report.ncu-rep> <out.json> "<what>"
  python tests/ncu_summarise.py launches <launches.csv> <out.json>
  python tests/ncu_summarise.py dram <metrics.csv> <out.json>"""
import csv
import io
import json
import re
import subprocess
import sys
from collections import defaultdict

KEYS = ['gpu__time_duration.sum', 'sm__cycles_active.avg', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed.avg.per_cycle_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'lts__t_sector_hit_rate.pct', 'smsp__inst_executed.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum']


def raw_page(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units = rows[0], rows[1]
    return head, units, rows[2:]


def full(path, out, what):
    head, units, rows = raw_page(path)
    res = []
    for r in rows:
        d = dict(zip(head, r))
        u = dict(zip(head, units))
        rec = {'kernel': d.get('Kernel Name', '')[:120]}
        for k in KEYS:
            if k in d and d[k] != '':
                try:
                    rec[k] = float(d[k].replace(',', ''))
                except ValueError:
                    rec[k] = d[k]
                if u.get(k):
                    rec[k + ' [unit]'] = u[k]
        res.append(rec)
    first = res[0] if res else {}
    conv = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    dram = 0.0
    for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
        if k in first:
            dram += first[k] * conv.get(first.get(k + ' [unit]', 'byte'), 1.0)
    json.dump({'what': what, 'source': path.split('/')[-1], 'dram_bytes_per_launch': dram, 'launches': res},
              open(out, 'w'), indent=1)
    print(json.dumps(first, indent=1)[:1500], 'dram bytes', dram)


def launches(path, out):
    rows = [r for r in csv.reader(open(path, errors='ignore')) if len(r) > 10]
    head = rows[0]
    iname, ival = head.index('Kernel Name'), head.index('Metric Value')
    iunit = head.index('Metric Unit')
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        try:
            v = float(r[ival].replace(',', ''))
        except ValueError:
            continue
        scale = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'nsecond': 1e-6, 'usecond': 1e-3, 'msecond': 1.0, 'second': 1e3}.get(r[iunit], 1e-6)
        name = re.sub(r'\(.*', '', r[iname])[:90]
        agg[name][0] += 1
        agg[name][1] += v * scale
    tot = sum(v[1] for v in agg.values())
    tab = sorted(([k, v[0], round(v[1], 3), round(100 * v[1] / tot, 2)] for k, v in agg.items()), key=lambda t: -t[2])
    ours = sum(t[2] for t in tab if any(k in t[0] for k in ('adalog', 'linf::', 'fused::', 'sel::')))
    json.dump({'total_kernel_ms': round(tot, 3), 'launches': sum(t[1] for t in tab), 'our_kernels_share_pct': round(100 * ours / tot, 1),
               'kernels': [dict(name=t[0], launches=t[1], ms=t[2], share_pct=t[3]) for t in tab[:40]]}, open(out, 'w'), indent=1)
    for t in tab[:16]:
        print(t)
    print('ours', round(100 * ours / tot, 1), '% of', round(tot, 1), 'ms')


def dram(path, out):
    rows = [r for r in csv.reader(open(path, errors='ignore')) if len(r) > 10]
    head = rows[0]
    iname, imet, ival, iunit, iid = (head.index(k) for k in ('Kernel Name', 'Metric Name', 'Metric Value', 'Metric Unit', 'ID'))
    per = defaultdict(dict)
    for r in rows[1:]:
        try:
            v = float(r[ival].replace(',', ''))
        except ValueError:
            continue
        conv = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3,
                'nsecond': 1e-9, 'usecond': 1e-6, 'msecond': 1e-3, 'second': 1.0}.get(r[iunit], 1.0)
        per[(r[iid], re.sub(r'\(.*', '', r[iname])[:70])][r[imet]] = v * conv
    agg = defaultdict(list)
    for (i, name), m in per.items():
        if 'gpu__time_duration.sum' in m:
            b = m.get('dram__bytes_read.sum', 0) + m.get('dram__bytes_write.sum', 0)
            agg[name].append((b, m['gpu__time_duration.sum']))
    res = {}
    for name, v in agg.items():
        v = v[3:] if len(v) > 6 else v                 # skip warm-up launches
        b = sum(x[0] for x in v) / len(v)
        t = sum(x[1] for x in v) / len(v)
        res[name] = dict(launches=len(v), dram_bytes_per_launch=b, seconds_per_launch=t, dram_gb_per_s=b / t / 1e9)
        print(name, res[name])
    json.dump(res, open(out, 'w'), indent=1)


if __name__ == '__main__':
    mode = sys.argv[1]
    if mode == 'full':
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else '')
    elif mode == 'launches':
        launches(sys.argv[2], sys.argv[3])
    else:
        dram(sys.argv[2], sys.argv[3])
