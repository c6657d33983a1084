"""Minimal PyG-compatible containers for the ligand–pharmacophore hetero graph.

torch_geometric is not installable here, so the host side carries just enough of `HeteroData` / `Batch`
(reference usage: src/datasets/pdbbind_phore.py:1143-1188 builds it, src/utils/sampling.py:210,254 re-collates
it, src/models/score_model_phore.py reads it) for the reference-facing code to read like the reference:

    data['ligand'].pos / .x / .batch / .edge_mask / .mask_rotate / .phorefp / .norm / .norm_angle1/2
    data['ligand', 'ligand'].edge_index / .edge_attr     (alias of ('ligand','lig_bond','ligand'), smp:721)
    data['phore'].x / .pos / .norm / .phoretype / .batch
    data['phore', 'phore'].edge_index                     (alias of ('phore','phore_contact','phore'))
    data.num_graphs, data.complex_t, data.name, ...

A real PyG HeteroDataBatch is accepted everywhere by duck typing (see engine.pack_batch).
"""
import copy

import numpy as np
import torch


class Store:
    """Attribute bag for one node or edge type."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def keys(self):
        return list(self.__dict__.keys())

    def __contains__(self, k):
        return k in self.__dict__

    @property
    def num_nodes(self):
        return self.x.shape[0] if 'x' in self.__dict__ else self.pos.shape[0]

    def to(self, device):
        for k, v in self.__dict__.items():
            if torch.is_tensor(v):
                self.__dict__[k] = v.to(device)
            elif isinstance(v, dict):
                self.__dict__[k] = {a: (b.to(device) if torch.is_tensor(b) else b) for a, b in v.items()}
        return self


_EDGE_ALIASES = {('ligand', 'ligand'): ('ligand', 'lig_bond', 'ligand'),
                 ('phore', 'phore'): ('phore', 'phore_contact', 'phore')}


class HeteroGraph:
    """One ligand–pharmacophore pair (or, with `num_graphs`, a collated batch of them)."""

    def __init__(self):
        object.__setattr__(self, '_stores', {})
        object.__setattr__(self, '_attrs', {})

    def __getitem__(self, key):
        if isinstance(key, int):                         # PyG Batch[i] (get_example): graph i of a collated batch
            return uncollate(self)[key]
        key = _EDGE_ALIASES.get(key, key)
        if key not in self._stores:
            self._stores[key] = Store()
        return self._stores[key]

    def __getattr__(self, name):
        attrs = object.__getattribute__(self, '_attrs')
        if name in attrs:
            return attrs[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        self._attrs[name] = value

    def __contains__(self, name):
        return name in self._attrs

    @property
    def node_types(self):
        return [k for k in self._stores if isinstance(k, str)]

    @property
    def edge_types(self):
        return [k for k in self._stores if isinstance(k, tuple)]

    def to(self, device):
        for s in self._stores.values():
            s.to(device)
        for k, v in self._attrs.items():
            if torch.is_tensor(v):
                self._attrs[k] = v.to(device)
            elif isinstance(v, dict):
                self._attrs[k] = {a: (b.to(device) if torch.is_tensor(b) else b) for a, b in v.items()}
        return self

    def clone(self):
        return copy.deepcopy(self)

    # ---- Batch API (torch_geometric.data.Batch.from_data_list / to_data_list)
    def to_data_list(self):
        return uncollate(self)


def collate(data_list):
    """PyG `Batch.from_data_list` semantics: node tensors concatenated along dim 0,
    edge_index offset by the cumulative node count of its endpoint types, `.batch` vectors, python / numpy
    attributes kept as lists."""
    out = HeteroGraph()
    node_off = {}
    for nt in data_list[0].node_types:
        stores = [d[nt] for d in data_list]
        counts = [s.num_nodes for s in stores]
        node_off[nt] = np.concatenate([[0], np.cumsum(counts)])
        dst = out[nt]
        slices = {}
        for k in stores[0].keys():
            vals = [getattr(s, k) for s in stores]
            if torch.is_tensor(vals[0]) and vals[0].dim() > 0:
                setattr(dst, k, torch.cat(vals, 0))
                slices[k] = np.concatenate([[0], np.cumsum([v.shape[0] for v in vals])]).tolist()
            else:
                setattr(dst, k, vals)
        dst._slices = slices            # like PyG's slice_dict: node stores may carry per-edge tensors (edge_mask)
        dst.batch = torch.repeat_interleave(torch.arange(len(data_list)), torch.tensor(counts))
        dst.ptr = torch.from_numpy(node_off[nt]).long()
    for et in data_list[0].edge_types:
        stores = [d[et] for d in data_list]
        dst = out[et]
        for k in stores[0].keys():
            vals = [getattr(s, k) for s in stores]
            if k == 'edge_index':
                vals = [v + torch.tensor([[node_off[et[0]][i]], [node_off[et[-1]][i]]], dtype=v.dtype)
                        for i, v in enumerate(vals)]
                setattr(dst, k, torch.cat(vals, 1))
            elif torch.is_tensor(vals[0]):
                setattr(dst, k, torch.cat(vals, 0))
            else:
                setattr(dst, k, vals)
        dst.ptr = torch.tensor(np.concatenate([[0], np.cumsum([s.edge_index.shape[1] for s in stores])])).long()
    for k in data_list[0]._attrs:
        vals = [d._attrs[k] for d in data_list]
        out._attrs[k] = torch.cat(vals, 0) if torch.is_tensor(vals[0]) else vals
    out._attrs['num_graphs'] = len(data_list)
    return out


def uncollate(batch):
    """PyG `Batch.to_data_list`."""
    n = batch.num_graphs
    outs = [HeteroGraph() for _ in range(n)]
    for nt in batch.node_types:
        src = batch[nt]
        ptr = src.ptr.tolist()
        slices = getattr(src, '_slices', {})
        for k in src.keys():
            if k in ('batch', 'ptr', 'node_t', 'node_sigma_emb', '_slices'):
                continue
            v = getattr(src, k)
            sl = slices.get(k, ptr)
            for i, o in enumerate(outs):
                setattr(o[nt], k, v[sl[i]:sl[i + 1]].clone() if torch.is_tensor(v) else v[i])
    for et in batch.edge_types:
        src = batch[et]
        eptr = src.ptr.tolist()
        p0, p1 = batch[et[0]].ptr.tolist(), batch[et[-1]].ptr.tolist()
        for k in src.keys():
            if k == 'ptr':
                continue
            v = getattr(src, k)
            for i, o in enumerate(outs):
                if k == 'edge_index':
                    e = v[:, eptr[i]:eptr[i + 1]].clone()
                    e[0] -= p0[i]
                    e[1] -= p1[i]
                    setattr(o[et], k, e)
                else:
                    setattr(o[et], k, v[eptr[i]:eptr[i + 1]].clone() if torch.is_tensor(v) else v[i])
    for k, v in batch._attrs.items():
        if k in ('num_graphs', 'complex_t', 'graph_sigma_emb'):
            continue
        for i, o in enumerate(outs):
            o._attrs[k] = v[i:i + 1].clone() if torch.is_tensor(v) else v[i]
    return outs


class DataLoader:
    """torch_geometric.loader.DataLoader(data_list, batch_size) as used at sampling.py:210 (no shuffle)."""

    def __init__(self, data_list, batch_size=1, shuffle=False):
        assert not shuffle
        self.data_list, self.batch_size = list(data_list), batch_size

    def __iter__(self):
        for i in range(0, len(self.data_list), self.batch_size):
            yield collate(self.data_list[i:i + self.batch_size])

    def __len__(self):
        return (len(self.data_list) + self.batch_size - 1) // self.batch_size


# ---- flat (npz-friendly) serialisation of one pair, used for committed fixtures ------------------------------
_LIG_KEYS = ('x', 'pos', 'edge_mask', 'phorefp', 'norm', 'norm_angle1', 'norm_angle2')
_PH_KEYS = ('x', 'pos', 'norm', 'phoretype')


def graph_to_arrays(g, prefix=''):
    out = {}
    for k in _LIG_KEYS:
        out[f'{prefix}lig_{k}'] = getattr(g['ligand'], k).numpy()
    mr = g['ligand'].mask_rotate
    out[f'{prefix}lig_mask_rotate'] = np.asarray(mr if isinstance(mr, np.ndarray) else mr[0])
    out[f'{prefix}bond_index'] = g['ligand', 'ligand'].edge_index.numpy()
    out[f'{prefix}bond_attr'] = g['ligand', 'ligand'].edge_attr.numpy()
    for k in _PH_KEYS:
        out[f'{prefix}ph_{k}'] = getattr(g['phore'], k).numpy()
    out[f'{prefix}ph_edge_index'] = g['phore', 'phore'].edge_index.numpy()
    return out


def graph_from_arrays(a, prefix='', name='pair'):
    g = HeteroGraph()
    for k in _LIG_KEYS:
        setattr(g['ligand'], k, torch.from_numpy(np.asarray(a[f'{prefix}lig_{k}'])))
    g['ligand'].mask_rotate = np.asarray(a[f'{prefix}lig_mask_rotate']).astype(bool)
    g['ligand', 'ligand'].edge_index = torch.from_numpy(np.asarray(a[f'{prefix}bond_index'])).long()
    g['ligand', 'ligand'].edge_attr = torch.from_numpy(np.asarray(a[f'{prefix}bond_attr'])).float()
    for k in _PH_KEYS:
        setattr(g['phore'], k, torch.from_numpy(np.asarray(a[f'{prefix}ph_{k}'])))
    g['phore', 'phore'].edge_index = torch.from_numpy(np.asarray(a[f'{prefix}ph_edge_index'])).long()
    g.name = name
    return g
