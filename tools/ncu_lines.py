"""Warp-state samples of one kernel per SOURCE LINE (run locally, no GPU needed): joins the SASS page of an
`ncu --set full --import-source on` report with the line table of the cubin.

  cuobjdump -xelf all diffphore_b200/csrc/conv_fused2.o; nvdisasm -gi -c conv_fused2.sm_100a.cubin > /tmp/cf2.dis
  ncu -i prof.ncu-rep --page source --csv --print-source sass > /tmp/sass.csv
  python tools/ncu_lines.py /tmp/cf2.dis TpL3ELb0 /tmp/sass.csv "conv_fused2_kernel" 0 [min_share]

Every SASS instruction is attributed to its inline chain (innermost line first, the kernel-body line last); the table lists
the kernel-body lines in source order with their samples, executed warp instructions and the dominant stall reasons, and
under each the innermost lines that carry most of its samples."""
import collections
import csv
import re
import sys


def line_table(dis, fun):
    """[(offset, ((file, line), ...))] of the function whose mangled name contains `fun`."""
    out, chain, fresh, on = [], [], True, False
    mark = re.compile(r'//## File "([^"]+)", line (\d+)')
    ins = re.compile(r'^\s*/\*([0-9a-f]{4,})\*/\s+(\S.*);')
    for ln in open(dis, errors='ignore'):
        if ln.startswith('.text.'):
            on = fun in ln
            continue
        if not on:
            continue
        m = mark.search(ln)
        if m:
            if fresh:
                chain, fresh = [], False
            chain.append((m.group(1).rsplit('/', 1)[-1], int(m.group(2))))
            continue
        m = ins.match(ln)
        if m:
            out.append((int(m.group(1), 16), tuple(chain)))
            fresh = True
    return out


def main(dis, fun, src, name, index, min_share='0.4'):
    table = line_table(dis, fun)
    rows = list(csv.reader(open(src)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
    hits = [i for i in starts if name in rows[i][1]]
    i0 = hits[int(index)]
    i1 = min([s for s in starts if s > i0] + [len(rows)])
    hdr = rows[i0 + 1]
    col = {h: k for k, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    body = [r for r in rows[i0 + 2:i1] if len(r) == len(hdr)]
    assert len(body) == len(table), (len(body), len(table))
    base = int(body[0][col['Address']], 16)
    tot = sum(int(r[col['# Samples']] or 0) for r in body)
    outer = collections.defaultdict(lambda: [0, 0, collections.Counter(), collections.Counter()])
    for r, (off, chain) in zip(body, table):
        assert int(r[col['Address']], 16) - base == off
        n, ex = int(r[col['# Samples']] or 0), int(r[col['Instructions Executed']] or 0)
        key = chain[-1] if chain else ('?', 0)
        o = outer[key]
        o[0] += n
        o[1] += ex
        for h in stall_cols:
            o[2][h[6:]] += int(r[col[h]] or 0)
        o[3][chain[0] if chain else ('?', 0)] += n
    print(f'# {rows[i0][1]}: {tot} samples, {len(body)} SASS instructions')
    for key in sorted(outer, key=lambda k: (k[0], k[1])):
        n, ex, st, inner = outer[key]
        if 100 * n / tot < float(min_share):
            continue
        top = ', '.join(f'{k} {100 * v / max(n, 1):.0f}%' for k, v in st.most_common(3))
        print(f'{key[0]}:{key[1]:4d}  {100 * n / tot:5.1f}%  samples {n:7d}  warp-inst {ex:9d}  [{top}]')
        for ik, v in inner.most_common(10):
            if ik != key and v > 0.06 * n:
                print(f'        {100 * v / tot:5.1f}%  in {ik[0]}:{ik[1]}')


if __name__ == '__main__':
    main(*sys.argv[1:7])
