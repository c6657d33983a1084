"""ORACLE (test infrastructure, not product code).

CPU restatement of the pieces of e3nn 0.5.1 that DiffPhore's score model calls.  e3nn is a third-party
dependency pinned in the reference's `src/environment_diffphore.yml:117` (e3nn==0.5.1) and is NOT vendored
in /root/reference nor installable here, so its published algorithm is restated:

  * o3.spherical_harmonics(lmax<=2, normalize=True, normalization='component')
        call sites: src/models/score_model_phore.py:365,404,434,737,754,891,893
  * o3.FullyConnectedTensorProduct(in1, in2, out, shared_weights=False)   (mode 'uvw')
        call site : src/models/score_model_phore.py:123,137
  * o3.FullTensorProduct(sh_irreps, "2e")                                  (mode 'uvuv', no weights)
        call site : src/models/score_model_phore.py:276,366
  * e3nn.nn.BatchNorm (eval mode)                                          src/models/score_model_phore.py:132,148

The Wigner-3j tensors are the buffers e3nn itself serialised into the shipped checkpoint
(`*._compiled_main_left_right._w3j_*`, extracted by tools/extract_w3j.py); the l=0 special cases use
delta/sqrt(2l+1) exactly like e3nn's specialised code path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import math
import os
from collections import namedtuple

import numpy as np
import torch

_W3J_FILE = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'w3j.npz')
_w3j_cache = {}


def w3j(l1, l2, l3, dtype=torch.float32):
    """Real-basis Wigner-3j tensor C[i,j,k] (Frobenius norm 1) in e3nn's y-polar basis."""
    key = (l1, l2, l3, dtype)
    if key in _w3j_cache:
        return _w3j_cache[key]
    if l1 == 0 and l2 == 0 and l3 == 0:
        c = torch.ones(1, 1, 1, dtype=dtype)
    elif l1 == 0 and l2 == l3:
        c = torch.eye(2 * l2 + 1, dtype=dtype).reshape(1, 2 * l2 + 1, 2 * l2 + 1) / math.sqrt(2 * l2 + 1)
    elif l2 == 0 and l1 == l3:
        c = torch.eye(2 * l1 + 1, dtype=dtype).reshape(2 * l1 + 1, 1, 2 * l1 + 1) / math.sqrt(2 * l1 + 1)
    elif l3 == 0 and l1 == l2:
        c = torch.eye(2 * l1 + 1, dtype=dtype).reshape(2 * l1 + 1, 2 * l1 + 1, 1) / math.sqrt(2 * l1 + 1)
    else:
        z = np.load(_W3J_FILE)
        name = f'w3j_{l1}_{l2}_{l3}'
        if name in z:
            c = torch.from_numpy(z[name].astype(np.float64))
        else:
            # w3j(l2,l1,l3)[j,i,k] = (-1)^(l1+l2+l3) w3j(l1,l2,l3)[i,j,k]
            name = f'w3j_{l2}_{l1}_{l3}'
            if name not in z:
                raise KeyError(f'no Wigner-3j for {(l1, l2, l3)}')
            c = torch.from_numpy(z[name].astype(np.float64)).transpose(0, 1) * (-1) ** (l1 + l2 + l3)
        # shipped buffers are fp32 roundings of algebraic numbers; re-normalise in fp64 then cast
        c = (c / c.norm()).to(dtype)
    _w3j_cache[key] = c
    return c


# ---------------------------------------------------------------------------------------------------
# Irreps helpers.  An irreps is a list of (mul, l, p) with p = +1 (even, 'e') or -1 (odd, 'o').
# ---------------------------------------------------------------------------------------------------
def parse_irreps(s):
    out = []
    for tok in s.replace(' ', '').split('+'):
        if 'x' in tok:
            mul, ir = tok.split('x')
            mul = int(mul)
        else:
            mul, ir = 1, tok
        out.append((mul, int(ir[:-1]), 1 if ir[-1] == 'e' else -1))
    return out


def irreps_dim(irreps):
    return sum(m * (2 * l + 1) for m, l, _ in irreps)


def irreps_slices(irreps):
    off, out = 0, []
    for m, l, _ in irreps:
        out.append((off, off + m * (2 * l + 1)))
        off += m * (2 * l + 1)
    return out


def sh_irreps(lmax=2):
    """o3.Irreps.spherical_harmonics(lmax): 1x0e + 1x1o + 1x2e."""
    return [(1, l, (-1) ** l) for l in range(lmax + 1)]


def spherical_harmonics(vec, lmax=2, only_l=None):
    """e3nn `o3.spherical_harmonics(irreps, vec, normalize=True, normalization='component')`.

    vec is first normalised with F.normalize (eps 1e-12; zero vector -> zeros, so Y1=Y2=0, Y0=1).
    Component order within l: e3nn's real basis, m=-l..l, y polar.  `only_l=2` reproduces the "2e"-only call
    at score_model_phore.py:365.
    """
    u = torch.nn.functional.normalize(vec, dim=-1)
    x, y, z = u[..., 0], u[..., 1], u[..., 2]
    s3, s5 = math.sqrt(3.0), math.sqrt(5.0)
    y0 = torch.ones_like(x)
    y1 = [s3 * x, s3 * y, s3 * z]
    y2 = [s5 * s3 * x * z, s5 * s3 * x * y, s5 * (y * y - 0.5 * (x * x + z * z)), s5 * s3 * y * z,
          s5 * (s3 / 2.0) * (z * z - x * x)]
    if only_l == 2:
        return torch.stack(y2, -1)
    comps = [y0]
    if lmax >= 1:
        comps += y1
    if lmax >= 2:
        comps += y2
    return torch.stack(comps, -1)


Instr = namedtuple('Instr', 'i1 i2 io w_off w_len pw')


def fctp_instructions(irreps_in1, irreps_in2, irreps_out):
    """Instruction table of FullyConnectedTensorProduct (loop order i1, i2, io; 'uvw'; path weights
    irrep_normalization='component', path_normalization='element').  Returns (instrs, weight_numel)."""
    raw = []
    for i1, (m1, l1, p1) in enumerate(irreps_in1):
        for i2, (m2, l2, p2) in enumerate(irreps_in2):
            for io, (mo, lo, po) in enumerate(irreps_out):
                if po == p1 * p2 and abs(l1 - l2) <= lo <= l1 + l2:
                    raw.append((i1, i2, io))
    fan = {}
    for i1, i2, io in raw:
        fan[io] = fan.get(io, 0) + irreps_in1[i1][0] * irreps_in2[i2][0]
    instrs, off = [], 0
    for i1, i2, io in raw:
        n = irreps_in1[i1][0] * irreps_in2[i2][0] * irreps_out[io][0]
        pw = math.sqrt((2 * irreps_out[io][1] + 1) / fan[io])
        instrs.append(Instr(i1, i2, io, off, n, pw))
        off += n
    return instrs, off


def fctp_apply(irreps_in1, irreps_in2, irreps_out, instrs, x1, x2, w):
    """out[z, w, k] += pw * sum_u W[z,u,0,w] * sum_ij C_ijk x1[z,u,i] x2[z,j]   (all mul_in2 == 1)."""
    E = x1.shape[0]
    s1, s2, so = irreps_slices(irreps_in1), irreps_slices(irreps_in2), irreps_slices(irreps_out)
    out = x1.new_zeros(E, irreps_dim(irreps_out))
    for ins in instrs:
        m1, l1, _ = irreps_in1[ins.i1]
        m2, l2, _ = irreps_in2[ins.i2]
        mo, lo, _ = irreps_out[ins.io]
        assert m2 == 1
        a = x1[:, s1[ins.i1][0]:s1[ins.i1][1]].reshape(E, m1, 2 * l1 + 1)
        b = x2[:, s2[ins.i2][0]:s2[ins.i2][1]].reshape(E, 2 * l2 + 1)
        C = w3j(l1, l2, lo, x1.dtype)
        zk = torch.einsum('ijk,zui,zj->zuk', C, a, b)                       # [E, m1, 2lo+1]
        W = w[:, ins.w_off:ins.w_off + ins.w_len].reshape(E, m1, mo)
        res = torch.bmm(W.transpose(1, 2), zk) * ins.pw                     # [E, mo, 2lo+1]
        out[:, so[ins.io][0]:so[ins.io][1]] += res.reshape(E, -1)
    return out


def full_tp_irreps_out(irreps_in1, irreps_in2):
    """FullTensorProduct output irreps: one per (i1, i2, l_out), sorted by (l, p) with odd(-1) before even(+1)
    (e3nn Irreps.sort, stable).  Returns (irreps_out_sorted, instr list [(i1,i2,io_sorted)])."""
    outs = []
    for i1, (m1, l1, p1) in enumerate(irreps_in1):
        for i2, (m2, l2, p2) in enumerate(irreps_in2):
            for lo in range(abs(l1 - l2), l1 + l2 + 1):
                outs.append((m1 * m2, lo, p1 * p2, i1, i2))
    order = sorted(range(len(outs)), key=lambda i: (outs[i][1], outs[i][2]))
    inv = {o: n for n, o in enumerate(order)}
    irreps_out = [outs[i][:3] for i in order]
    instrs = [(outs[i][3], outs[i][4], inv[i]) for i in range(len(outs))]
    return irreps_out, instrs


def full_tp_apply(irreps_in1, irreps_in2, x1, x2):
    """FullTensorProduct ('uvuv', unweighted): out_k = sqrt(2l_out+1) * sum_ij C_ijk a_i b_j  (muls == 1)."""
    irreps_out, instrs = full_tp_irreps_out(irreps_in1, irreps_in2)
    s1, s2, so = irreps_slices(irreps_in1), irreps_slices(irreps_in2), irreps_slices(irreps_out)
    out = x1.new_zeros(x1.shape[0], irreps_dim(irreps_out))
    for i1, i2, io in instrs:
        m1, l1, _ = irreps_in1[i1]
        m2, l2, _ = irreps_in2[i2]
        _, lo, _ = irreps_out[io]
        assert m1 == 1 and m2 == 1
        C = w3j(l1, l2, lo, x1.dtype)
        a = x1[:, s1[i1][0]:s1[i1][1]]
        b = x2[:, s2[i2][0]:s2[i2][1]]
        out[:, so[io][0]:so[io][1]] = math.sqrt(2 * lo + 1) * torch.einsum('ijk,zi,zj->zk', C, a, b)
    return irreps_out, out


def batchnorm_eval(x, irreps, weight, bias, running_mean, running_var, eps=1e-5):
    """e3nn.nn.BatchNorm in eval mode (affine=True, normalization='component', reduce='mean'): scalars (0e)
    are shifted by running_mean and get a bias; every other irrep (incl. 0o) is only rescaled per channel."""
    out = torch.empty_like(x)
    ix = iw = ib = 0
    for mul, l, p in irreps:
        d = 2 * l + 1
        f = x[:, ix:ix + mul * d].reshape(-1, mul, d)
        if l == 0 and p == 1:
            f = f - running_mean[ib:ib + mul].reshape(1, mul, 1)
        scale = (running_var[iw:iw + mul] + eps).pow(-0.5) * weight[iw:iw + mul]
        f = f * scale.reshape(1, mul, 1)
        if l == 0 and p == 1:
            f = f + bias[ib:ib + mul].reshape(1, mul, 1)
            ib += mul
        out[:, ix:ix + mul * d] = f.reshape(-1, mul * d)
        ix += mul * d
        iw += mul
    return out
