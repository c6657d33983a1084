"""Synthetic ligand–pharmacophore pairs with the reference's tensor schema.

There is no RDKit (and no dataset) here, so the benchmark / parity workloads of BASELINE.json
("synthetic 256 ligand–pharmacophore pairs (32 atoms / 8 phore points)", ...) are generated directly as the
tensors that the reference's preprocessing would hand to the sampler:

    get_lig_graph            src/datasets/process_mols.py:255-284   (x, pos, bond edge_index/edge_attr)
    generate_ligand_phore_feat   src/datasets/process_mols.py:376-417   (phorefp, norm, norm_angle1/2)
    get_transformation_mask  src/utils/torsion.py:13-61             (edge_mask, mask_rotate)
    get_phore_graph          src/datasets/process_pharmacophore.py:634-714 (phore x/pos/norm/edges)
    generate_graph           src/datasets/pdbbind_phore.py:1143-1188 (centring on the phore centroid)

Ligand = random tree (acyclic, bond length 1.5 A, non-bonded contacts >= 2 A), so rotatable bonds are exactly
the bonds whose two sides both have >= 2 atoms.  Seeded per pair: numpy default_rng(1000 + pair_id).
"""
import math

import numpy as np
import torch

from .graph import HeteroGraph

LIG_FEATURE_DIMS = [119, 4, 12, 12, 8, 10, 6, 6, 2, 8, 2, 2, 2, 2, 2, 2]   # process_mols.py:162-179
# process_pharmacophore.py:56,74 (index = phore type, last = EX)
PHORE_PRE_WEIGHT = [1.5, 1.2, 1.0, 1.5, 1.2, 0.5, 1.5, 1.0, 1.0, 1.0, 1.0]
PHORE_PRE_ALPHA = [1.0, 1.0, 0.7, 1.0, 1.0, 0.7, 1.0, 1.0, 0.7, 1.0, 0.837]


def _unit(rng, n=None):
    v = rng.normal(size=(3,) if n is None else (n, 3))
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def _grow_tree(rng, n_atoms, bond=1.5, min_contact=2.0):
    pos = np.zeros((n_atoms, 3))
    parent = -np.ones(n_atoms, dtype=np.int64)
    deg = np.zeros(n_atoms, dtype=np.int64)
    for a in range(1, n_atoms):
        for _ in range(10000):
            p = int(rng.choice(np.nonzero(deg[:a] < 4)[0]))
            cand = pos[p] + bond * _unit(rng)
            d = np.linalg.norm(pos[:a] - cand, axis=1)
            d[p] = np.inf
            if d.min() >= min_contact:
                break
        else:
            raise RuntimeError('could not place atom')
        pos[a], parent[a] = cand, p
        deg[a] += 1
        deg[p] += 1
    return pos, parent


def _transformation_mask(n_atoms, bonds):
    """get_transformation_mask (torsion.py:13-61) for an acyclic molecule; bonds = list of (begin, end), the
    directed copies are listed consecutively [(b,e),(e,b)] as in get_lig_graph."""
    adj = [[] for _ in range(n_atoms)]
    for b, e in bonds:
        adj[b].append(e)
        adj[e].append(b)
    mask_edges, rows = [], []
    for b, e in bonds:
        # component containing b after removing the bond
        seen = {b}
        stack = [b]
        while stack:
            u = stack.pop()
            for v in adj[u]:
                if (u == b and v == e) or (u == e and v == b) or v in seen:
                    continue
                seen.add(v)
                stack.append(v)
        side_b = seen
        side_e = set(range(n_atoms)) - side_b
        small = side_b if len(side_b) <= len(side_e) else side_e          # sorted(components, key=len)[0]
        if len(small) > 1:
            row = np.zeros(n_atoms, dtype=bool)
            row[list(small)] = True
            if b in small:          # rotating side must contain the edge's 2nd atom -> reversed copy carries it
                mask_edges += [False, True]
            else:
                mask_edges += [True, False]
            rows.append(row)
        else:
            mask_edges += [False, False]
    mask_rotate = np.stack(rows, 0) if rows else np.zeros((0, n_atoms), dtype=bool)
    return np.asarray(mask_edges, dtype=bool), mask_rotate


def make_pair(pair_id, n_atoms=32, n_phore=8, seed_base=1000):
    """One ligand–pharmacophore HeteroGraph in the reference schema (un-noised pose, centred on the
    pharmacophore centroid like pdbbind_phore.py:1179-1184)."""
    rng = np.random.default_rng(seed_base + pair_id)
    pos, parent = _grow_tree(rng, n_atoms)
    bonds = [(int(parent[a]), a) for a in range(1, n_atoms)]
    g = HeteroGraph()
    lig = g['ligand']
    # chemically plausible categorical features of an acyclic sp3/sp2 molecule (indices into the lists of
    # process_mols.py:127-152): the shipped weights are only well-conditioned on in-distribution features.
    deg = np.bincount(np.asarray(bonds).reshape(-1), minlength=n_atoms)
    elem = np.where(deg >= 4, 5, rng.choice([5, 6, 7], size=n_atoms, p=[0.7, 0.15, 0.15]))   # C / N / O (Z-1)
    elem = np.where((deg == 3) & (elem == 7), 6, elem)
    valence = np.asarray({5: 4, 6: 3, 7: 2}[int(e)] for e in elem) if False else np.vectorize({5: 4, 6: 3, 7: 2}.get)(elem)
    num_h = np.clip(valence - deg, 0, 8)
    x = np.zeros((n_atoms, 16), dtype=np.int64)
    x[:, 0] = elem
    x[:, 1] = 0                                                       # CHI_UNSPECIFIED
    x[:, 2] = np.clip(deg + num_h, 0, 10)                             # total degree
    x[:, 3] = 5                                                       # formal charge 0
    x[:, 4] = num_h                                                   # implicit valence
    x[:, 5] = num_h                                                   # total Hs
    x[:, 6] = 0                                                       # radical electrons
    x[:, 7] = np.where(rng.random(n_atoms) < 0.8, 2, 1)               # SP3 / SP2
    lig.x = torch.from_numpy(x).long()
    ei = np.asarray([c for b, e in bonds for c in ((b, e), (e, b))]).T.copy()
    btype = np.repeat(rng.integers(0, 4, size=len(bonds)), 2)
    g['ligand', 'ligand'].edge_index = torch.from_numpy(ei).long()
    g['ligand', 'ligand'].edge_attr = torch.nn.functional.one_hot(torch.from_numpy(btype), 4).float()
    mask_edges, mask_rotate = _transformation_mask(n_atoms, bonds)
    lig.edge_mask = torch.from_numpy(mask_edges)
    lig.mask_rotate = mask_rotate
    fp = np.zeros((n_atoms, 11), dtype=np.float32)
    fp[:, :10] = rng.random((n_atoms, 10)) < 0.15
    norm = np.zeros((n_atoms, 11, 3), dtype=np.float32)
    on = fp > 0
    norm[on] = _unit(rng, int(on.sum())) if on.any() else 0
    angles = np.asarray([0.0, math.pi / 3, math.pi], dtype=np.float32)
    a1 = np.where(on, angles[rng.integers(0, 3, size=on.shape)], 0).astype(np.float32)
    a2 = np.where(on, angles[rng.integers(0, 3, size=on.shape)], 0).astype(np.float32)
    lig.phorefp = torch.from_numpy(fp)
    lig.norm = torch.from_numpy(norm.reshape(n_atoms, 33))
    lig.norm_angle1 = torch.from_numpy(a1)
    lig.norm_angle2 = torch.from_numpy(a2)
    lig.ph = torch.from_numpy(fp.max(0))

    # pharmacophore: ceil(0.6 P) features near distinct atoms, the rest exclusion spheres on a 3-5 A shell
    n_feat = min(int(math.ceil(0.6 * n_phore)), n_phore - 1)
    n_ex = n_phore - n_feat
    atoms = rng.choice(n_atoms, size=n_feat, replace=False)
    fpos = pos[atoms] + _unit(rng, n_feat) * rng.random((n_feat, 1)) * 0.7
    ftype = rng.integers(0, 10, size=n_feat)
    fhas = rng.random(n_feat) < 0.5
    fnorm = np.where(fhas[:, None], _unit(rng, n_feat), 0.0)
    epos = []
    while len(epos) < n_ex:
        c = pos[rng.integers(0, n_atoms)] + _unit(rng) * rng.uniform(3.0, 5.0)
        if np.linalg.norm(pos - c, axis=1).min() >= 3.0:
            epos.append(c)
    ppos = np.concatenate([fpos, np.asarray(epos).reshape(n_ex, 3)], 0)
    ptype = np.concatenate([ftype, np.full(n_ex, 10)])
    phas = np.concatenate([fhas, np.zeros(n_ex, dtype=bool)])
    pnorm = np.concatenate([fnorm, np.zeros((n_ex, 3))], 0)
    is_ex = ptype == 10
    # phore_featurizer (process_pharmacophore.py:717-748): [type idx, index([True,False], isEX), index([True,False], has_norm), alpha, weight]
    px = np.stack([ptype, np.where(is_ex, 0, 1), np.where(phas, 0, 1),
                   np.asarray(PHORE_PRE_ALPHA)[ptype], np.asarray(PHORE_PRE_WEIGHT)[ptype]], 1)
    ph = g['phore']
    ph.x = torch.from_numpy(px).float()
    ph.norm = torch.from_numpy(pnorm).float()
    ph.phoretype = torch.nn.functional.one_hot(torch.from_numpy(ptype), 11).float()
    # get_phore_graph edges (process_pharmacophore.py:660-680)
    dist = np.linalg.norm(ppos[:, None] - ppos[None], axis=-1)
    src, dst = [], []
    for i in range(n_phore):
        if i < n_feat:
            d = [j for j in range(n_feat) if j != i]
        else:
            d = [int(j) for j in np.where(dist[i] < 5.0)[0] if j != i]
        if not d:
            d = [int(np.argsort(dist[i])[1])]
        src += [i] * len(d)
        dst += d
    g['phore', 'phore'].edge_index = torch.tensor([src, dst]).long()
    center = ppos.mean(0, keepdims=True)
    ph.pos = torch.from_numpy(ppos - center).float()
    lig.pos = torch.from_numpy(pos - center).float()
    g.original_center = torch.from_numpy(center).float()
    g.name = f'syn{pair_id}'
    return g


def make_pairs(n_pairs, n_atoms=32, n_phore=8, first=0):
    return [make_pair(first + i, n_atoms, n_phore) for i in range(n_pairs)]


SHIPPED_MODEL_KW = dict(sigma_embed_dim=20, ns=20, nv=10, num_conv_layers=4, distance_embed_dim=20, cross_distance_embed_dim=20,
                        consider_norm=True, boarder=True, use_phore_match_feat=True, cross_distance_transition=True,
                        phore_direction_transition=True, phoretype_match_transition=True, atom_weight='phore',
                        auto_phorefp=False, scaler=100.0, dropout=0.1, clash_cutoff=[1.0, 2.0, 3.0, 4.0, 5.0])


def random_state_dict(seed=0):
    """Random-init weights of the shipped architecture (model_parameters.yml; nn defaults) with non-trivial BatchNorm statistics:
    what the synthetic benchmark configurations run (there is no dataset the shipped checkpoint is in distribution for).
    Needs the reference-facing mirror `src/models/score_model_phore.py` on sys.path (bench.py / tests put it there)."""
    import os
    import sys
    src = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'src')
    if src not in sys.path:
        sys.path.insert(0, src)
    from models.score_model_phore import TensorProductScoreModel
    torch.manual_seed(seed)
    m = TensorProductScoreModel(None, torch.device('cpu'), None, **SHIPPED_MODEL_KW)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(seed + 1)
    for k in sd:
        if k.endswith('batch_norm.running_var'):
            sd[k] = torch.rand(sd[k].shape, generator=g) * 1.5 + 0.5
        elif k.endswith('batch_norm.running_mean') or k.endswith('batch_norm.bias'):
            sd[k] = torch.randn(sd[k].shape, generator=g) * 0.3
        elif k.endswith('batch_norm.weight'):
            sd[k] = torch.rand(sd[k].shape, generator=g) * 0.1 + 0.05     # small updates keep the random net well-conditioned
    return sd


def real_example_pairs(n_pairs, path=None):
    """The reference's example ligands x its 79-node example pharmacophore as packed tensors (tests/golden/real_pairs.npz, made by
    tools/make_real_inputs.py: pharmacophore side from the reference's own parser, ligand side from the reduced featuriser)."""
    import os
    from .graph import graph_from_arrays
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    a = np.load(path or os.path.join(root, 'tests', 'golden', 'real_pairs.npz'))
    names = list(a['names'])
    return [graph_from_arrays(a, f'p{k}_', str(names[k])) for k in range(min(n_pairs, len(names)))]
