"""Condense ncu outputs (run locally, no GPU needed) into the small CSVs committed under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/launches_rN.csv "<command line used>"
  python tools/ncu_summary.py raw gpurun_out/prof.ncu-rep profiles/ncu_conv_fused_rN.csv
"""
import collections, csv, subprocess, sys

KEEP = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'sm__cycles_elapsed.avg', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum']


def launches(src, dst, cmd):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = next(r for r in rows if 'Kernel Name' in r)
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg, tot = collections.OrderedDict(), 0.0
    for r in rows[rows.index(hdr) + 1:]:
        try:
            v = float(r[vi].replace(',', ''))
        except ValueError:
            continue
        v = v / 1e3 if r[ui] == 'ns' else (v * 1e3 if r[ui] == 'ms' else v)
        name = r[ki].split('(')[0].replace('void ', '')
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
    with open(dst, 'w') as f:
        f.write(f'# ncu launch list summary (cold-cache, serialised; compare SHARES)\n# command: {cmd}\n')
        f.write('kernel,launches,total_us,share\n')
        for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f'{k},{n},{t:.1f},{t / tot:.4f}\n')
    print(open(dst).read())


def raw(src, dst):
    out = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    idx = [h.index('Kernel Name')] + [h.index(k) for k in KEEP if k in h]
    with open(dst, 'w') as f:
        w = csv.writer(f)
        w.writerow([h[i] for i in idx]); w.writerow([units[i] for i in idx])
        for r in rows[2:]:
            w.writerow([r[i] for i in idx])
    print(open(dst).read())


if __name__ == '__main__':
    {'launches': launches, 'raw': raw}[sys.argv[1]](*sys.argv[2:])
