"""Irreps bookkeeping for the tensor-product convolutions (host side, numpy only).

Mirrors what the reference gets from e3nn 0.5.1 (`o3.Irreps`, `o3.FullyConnectedTensorProduct`,
`o3.FullTensorProduct`; call sites score_model_phore.py:123,211,276,586-591): instruction enumeration, per-edge
weight layout, path normalisation, and the Wigner-3j constants (shipped as data/w3j.npz, extracted from the
buffers e3nn serialised into the reference checkpoint).
"""
import math
import os
from collections import namedtuple

import numpy as np

_W3J = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data', 'w3j.npz')
Instr = namedtuple('Instr', 'i1 i2 io w_off w_len pw')


def IRREP_SEQ(ns, nv):
    return [f'{ns}x0e', f'{ns}x0e + {nv}x1o', f'{ns}x0e + {nv}x1o + {nv}x1e',
            f'{ns}x0e + {nv}x1o + {nv}x1e + {ns}x0o']


def parse_irreps(s):
    out = []
    for tok in s.replace(' ', '').split('+'):
        mul, ir = tok.split('x') if 'x' in tok else (1, tok)
        out.append((int(mul), int(ir[:-1]), 1 if ir[-1] == 'e' else -1))
    return out


def irreps_str(irreps):
    return '+'.join(f"{m}x{l}{'e' if p == 1 else 'o'}" for m, l, p in irreps)


def irreps_dim(irreps):
    return sum(m * (2 * l + 1) for m, l, _ in irreps)


def irreps_offsets(irreps):
    off, out = 0, []
    for m, l, _ in irreps:
        out.append(off)
        off += m * (2 * l + 1)
    return out


def sh_irreps(lmax=2):
    return [(1, l, (-1) ** l) for l in range(lmax + 1)]


def fctp_instructions(in1, in2, out):
    raw = [(a, b, c) for a, (_, l1, p1) in enumerate(in1) for b, (_, l2, p2) in enumerate(in2)
           for c, (_, lo, po) in enumerate(out) if po == p1 * p2 and abs(l1 - l2) <= lo <= l1 + l2]
    fan = {}
    for a, b, c in raw:
        fan[c] = fan.get(c, 0) + in1[a][0] * in2[b][0]
    instrs, off = [], 0
    for a, b, c in raw:
        n = in1[a][0] * in2[b][0] * out[c][0]
        instrs.append(Instr(a, b, c, off, n, math.sqrt((2 * out[c][1] + 1) / fan[c])))
        off += n
    return instrs, off


def full_tp_irreps_out(in1, in2):
    outs = [(m1 * m2, lo, p1 * p2, a, b) for a, (m1, l1, p1) in enumerate(in1) for b, (m2, l2, p2) in enumerate(in2)
            for lo in range(abs(l1 - l2), l1 + l2 + 1)]
    order = sorted(range(len(outs)), key=lambda i: (outs[i][1], outs[i][2]))      # Irreps.sort(): (l, p), stable
    inv = {o: n for n, o in enumerate(order)}
    return [outs[i][:3] for i in order], [(outs[i][3], outs[i][4], inv[i]) for i in range(len(outs))]


def w3j(l1, l2, l3):
    if l1 == 0 and l2 == l3:
        return np.eye(2 * l2 + 1).reshape(1, 2 * l2 + 1, 2 * l2 + 1) / math.sqrt(2 * l2 + 1)
    if l2 == 0 and l1 == l3:
        return np.eye(2 * l1 + 1).reshape(2 * l1 + 1, 1, 2 * l1 + 1) / math.sqrt(2 * l1 + 1)
    if l3 == 0 and l1 == l2:
        return np.eye(2 * l1 + 1).reshape(2 * l1 + 1, 2 * l1 + 1, 1) / math.sqrt(2 * l1 + 1)
    z = np.load(_W3J)
    c = z[f'w3j_{l1}_{l2}_{l3}'].astype(np.float64)
    return c / np.linalg.norm(c)


def w3j_buffers():
    """name -> fp32 array, exactly as serialised by e3nn in the reference checkpoint."""
    z = np.load(_W3J)
    return {k: z[k] for k in z.files}
