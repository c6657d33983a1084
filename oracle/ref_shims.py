"""ORACLE tooling (build container only): import the UNMODIFIED reference score model
(/root/reference/src/models/score_model_phore.py) with its un-installable third-party dependencies replaced by thin
shims, so that the reference's OWN code (graph builders, gating, angle matching, encoder wiring, heads) can be
executed here and used to pin oracle/model.py.

What is real: every line of score_model_phore.py and models/e3phore.py, torch itself.
What is shimmed (restated in oracle/e3nn_lite.py / oracle/model.py, third-party, pinned versions in
src/environment_diffphore.yml): e3nn 0.5.1 (o3.Irreps, FullyConnectedTensorProduct, FullTensorProduct,
spherical_harmonics, nn.BatchNorm), torch_cluster 1.6.0 (radius, radius_graph), torch_scatter 2.0.9 (scatter*),
torch_geometric.utils.to_dense_batch, and the reference's import-time table modules utils.so3 / utils.torus (they
write cache files into the read-only mount) plus the RDKit-importing dataset modules (only two constants are used).
The shim modules register the same parameter / buffer names as the real ones, so the reference class loads the
shipped checkpoint with strict=True.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch
from torch import nn

from . import e3nn_lite as e3
from . import model as om

REF_SRC = '/root/reference/src'


class Irreps(list):
    def __init__(self, spec):
        if isinstance(spec, Irreps):
            super().__init__(spec)
        elif isinstance(spec, str):
            super().__init__(e3.parse_irreps(spec))
        else:
            super().__init__([tuple(x) for x in spec])

    @classmethod
    def spherical_harmonics(cls, lmax):
        return cls(e3.sh_irreps(lmax))

    @property
    def dim(self):
        return e3.irreps_dim(self)


class _Compiled(nn.Module):
    def __init__(self, triples):
        super().__init__()
        z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'w3j.npz'))
        for a, b, c in triples:
            self.register_buffer(f'_w3j_{a}_{b}_{c}', torch.from_numpy(z[f'w3j_{a}_{b}_{c}'].copy()))


class FullyConnectedTensorProduct(nn.Module):
    def __init__(self, irreps_in1, irreps_in2, irreps_out, shared_weights=True, **kw):
        super().__init__()
        assert shared_weights is False
        self.irreps_in1, self.irreps_in2, self.irreps_out = Irreps(irreps_in1), Irreps(irreps_in2), Irreps(irreps_out)
        self.instrs, self.weight_numel = e3.fctp_instructions(self.irreps_in1, self.irreps_in2, self.irreps_out)
        tri = sorted({(self.irreps_in1[i.i1][1], self.irreps_in2[i.i2][1], self.irreps_out[i.io][1]) for i in self.instrs})
        self.weight = nn.Parameter(torch.zeros(0))
        self.register_buffer('output_mask', torch.ones(self.irreps_out.dim))
        self._compiled_main_left_right = _Compiled([t for t in tri if min(t) > 0])

    def forward(self, x1, x2, w):
        return e3.fctp_apply(self.irreps_in1, self.irreps_in2, self.irreps_out, self.instrs, x1, x2, w)


class FullTensorProduct(nn.Module):
    def __init__(self, irreps_in1, irreps_in2, **kw):
        super().__init__()
        self.irreps_in1, self.irreps_in2 = Irreps(irreps_in1), Irreps(irreps_in2)
        out, _ = e3.full_tp_irreps_out(self.irreps_in1, self.irreps_in2)
        self.irreps_out = Irreps(out)
        tri = sorted({(l1, l2, lo) for (_, l1, _) in self.irreps_in1 for (_, l2, _) in self.irreps_in2
                      for lo in range(abs(l1 - l2), l1 + l2 + 1)})
        self.weight = nn.Parameter(torch.zeros(0))
        self.register_buffer('output_mask', torch.ones(self.irreps_out.dim))
        self._compiled_main_left_right = _Compiled(tri)

    def forward(self, x1, x2):
        return e3.full_tp_apply(self.irreps_in1, self.irreps_in2, x1, x2)[1]


def spherical_harmonics(irreps, x, normalize, normalization='integral'):
    assert normalize and normalization == 'component'
    if isinstance(irreps, str):
        assert irreps == '2e'
        return e3.spherical_harmonics(x, only_l=2)
    lmax = max(l for _, l, _ in irreps)
    return e3.spherical_harmonics(x, lmax=lmax)


class BatchNorm(nn.Module):
    def __init__(self, irreps, eps=1e-5, **kw):
        super().__init__()
        self.irreps, self.eps = Irreps(irreps), eps
        n = sum(m for m, _, _ in self.irreps)
        ns = sum(m for m, l, p in self.irreps if l == 0 and p == 1)
        self.weight, self.bias = nn.Parameter(torch.ones(n)), nn.Parameter(torch.zeros(ns))
        self.register_buffer('running_mean', torch.zeros(ns))
        self.register_buffer('running_var', torch.ones(n))

    def forward(self, x):
        assert not self.training
        return e3.batchnorm_eval(x, self.irreps, self.weight, self.bias, self.running_mean, self.running_var, self.eps)


def _radius(x, y, r, batch_x=None, batch_y=None, max_num_neighbors=32, **kw):
    return om.radius_pairs(x, y, r, batch_x, batch_y, max_num_neighbors)


def _radius_graph(x, r, batch=None, loop=False, max_num_neighbors=32, **kw):
    return om.radius_graph(x, r, batch, max_num_neighbors)


def _scatter(src, index, dim=0, out=None, dim_size=None, reduce='sum'):
    assert dim == 0
    n = int(dim_size) if dim_size is not None else int(index.max()) + 1
    return om.scatter(src, index.long(), n, 'mean' if reduce == 'mean' else 'sum')


def _to_dense_batch(x, batch, fill_value=0):
    B = int(batch.max()) + 1
    counts = torch.bincount(batch, minlength=B)
    nmax = int(counts.max())
    ptr = torch.cat([counts.new_zeros(1), counts.cumsum(0)])
    idx = torch.arange(batch.shape[0]) - ptr[batch]
    out = x.new_full((B, nmax) + tuple(x.shape[1:]), fill_value)
    out[batch, idx] = x
    mask = torch.zeros(B, nmax, dtype=torch.bool)
    mask[batch, idx] = True
    return out, mask


def install(so3_norm, torus_norm):
    """Put the shims into sys.modules and return the reference's score_model_phore module."""
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    o3 = mod('e3nn.o3', Irreps=Irreps, FullyConnectedTensorProduct=FullyConnectedTensorProduct,
             FullTensorProduct=FullTensorProduct, spherical_harmonics=spherical_harmonics)
    enn = mod('e3nn.nn', BatchNorm=BatchNorm)
    mod('e3nn', o3=o3, nn=enn)
    mod('torch_cluster', radius=_radius, radius_graph=_radius_graph)
    mod('torch_scatter', scatter=_scatter, scatter_mean=None, scatter_add=None, scatter_max=None)
    tgu = mod('torch_geometric.utils', to_dense_batch=_to_dense_batch)
    mod('torch_geometric', utils=tgu)
    so3 = mod('utils.so3', score_norm=lambda eps: torch.from_numpy(np.asarray(so3_norm(eps.numpy()))).float())
    torus = mod('utils.torus', score_norm=lambda s: np.asarray(torus_norm(np.asarray(s))))
    mod('utils', so3=so3, torus=torus)
    pm = mod('datasets.process_mols', lig_feature_dims=(om.LIG_FEATURE_DIMS, 0))
    pp = mod('datasets.process_pharmacophore', phore_feature_dims=(om.PHORE_FEATURE_DIMS, 2))
    mod('datasets', process_mols=pm, process_pharmacophore=pp)
    mod('models')

    def load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        spec.loader.exec_module(m)
        return m

    load('models.e3phore', os.path.join(REF_SRC, 'models/e3phore.py'))
    return load('models.score_model_phore', os.path.join(REF_SRC, 'models/score_model_phore.py'))
