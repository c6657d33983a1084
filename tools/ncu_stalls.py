"""Digest of the source page of an `ncu --set full --import-source on` report for one kernel (run locally, no GPU needed):
stall-reason totals, samples by opcode, top SASS instructions by samples.

  ncu -i gpurun_out/prof.ncu-rep --page source --csv --print-source sass > /tmp/src.csv
  python tools/ncu_stalls.py /tmp/src.csv "conv_fused_kernel<TpL3" 0 profiles/ncu_conv_fused_lig3_stalls_rN.txt "<header note>"
(kernel-name substring, index among the matching launches in the report)"""
import collections
import csv
import sys


def main(src, name, index, dst, note):
    rows = list(csv.reader(open(src)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
    hits = [i for i in starts if name in rows[i][1]]
    i0 = hits[int(index)]
    i1 = min([s for s in starts if s > i0] + [len(rows)])
    hdr = rows[i0 + 1]
    col = {h: k for k, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    body = [r for r in rows[i0 + 2:i1] if len(r) == len(hdr)]
    tot = sum(int(r[col['# Samples']] or 0) for r in body)
    reasons, opc = collections.Counter(), collections.Counter()
    for r in body:
        n = int(r[col['# Samples']] or 0)
        for h in stall_cols:
            reasons[h[6:]] += int(r[col[h]] or 0)
        sass = r[col['Source']].split()
        op = next((t for t in sass if not t.startswith('@')), '?').split('.')[0]
        opc[op] += n
    with open(dst, 'w') as f:
        f.write(f'# {note}\n# kernel: {rows[i0][1]}; total samples {tot} over {len(body)} SASS instructions\n')
        f.write('# stall reasons (all warps): ' + ', '.join(f'{k} {100 * v / tot:.1f}%' for k, v in reasons.most_common(8)) + '\n')
        f.write('# samples by opcode: ' + ', '.join(f'{k} {100 * v / tot:.1f}%' for k, v in opc.most_common(12)) + '\n')
        f.write('# top instructions by samples: samples, share, SASS, dominant stall reasons\n')
        for r in sorted(body, key=lambda r: -int(r[col['# Samples']] or 0))[:30]:
            n = int(r[col['# Samples']] or 0)
            top = sorted(((int(r[col[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:2]
            f.write(f"{n:6d} {100 * n / tot:5.1f}%  {r[col['Source']][:70]:70s} {dict((k, v) for v, k in top)}\n")
    print(open(dst).read()[:1500])


if __name__ == '__main__':
    main(*sys.argv[1:6])
