"""Ligand ingestion / pose output, host side.

Mirrors the parts of /root/reference/src/datasets/process_mols.py the inference path needs: the tensor schema of
`get_lig_graph` (:255-284) + `generate_ligand_phore_feat` (:376-417, norms per :782-858), and the SD writers
`write_mol_with_multi_coords` (:888-921).  The reference does all of this with RDKit, which stays the intended
production path (north_star: "RDKit/AncPhore preprocessing path stays host Python"); RDKit is not installable here,
so for already-3D SD files a REDUCED, RDKit-free featuriser is provided (heavy atoms, explicit-H counts, ring
perception by cycle basis, crude aromaticity / hybridisation / pharmacophore typing).  It yields the same tensor
schema; chemical fidelity of the categorical features is not guaranteed (documented deviation).
"""
import random

import networkx as nx
import numpy as np
import torch

lig_feature_dims = ([119, 4, 12, 12, 8, 10, 6, 6, 2, 8, 2, 2, 2, 2, 2, 2], 0)
PHORETYPES = ['MB', 'HD', 'AR', 'PO', 'HA', 'HY', 'NE', 'CV', 'CR', 'XB', 'EX']
PI = 3.1415926
_Z = {'H': 1, 'B': 5, 'C': 6, 'N': 7, 'O': 8, 'F': 9, 'Si': 14, 'P': 15, 'S': 16, 'Cl': 17, 'Se': 34, 'Br': 35, 'I': 53}


def read_sdf_block(text):
    """First V2000 mol block of an SD file -> (name, elements, coords [n,3], bonds [(a, b, order)], raw lines)."""
    L = text.split('\n')
    na, nb = int(L[3][0:3]), int(L[3][3:6])
    elem, xyz = [], []
    for l in L[4:4 + na]:
        xyz.append([float(l[0:10]), float(l[10:20]), float(l[20:30])])
        elem.append(l[31:34].strip())
    bonds = [(int(l[0:3]) - 1, int(l[3:6]) - 1, int(l[6:9])) for l in L[4 + na:4 + na + nb]]
    # keep the property block (M  CHG / M  ISO / M  RAD ...) up to M  END: most non-RDKit writers carry charges only there
    props = []
    for l in L[4 + na + nb:]:
        if l.startswith('M  END') or l.startswith('$$$$'):
            break
        if l.startswith('M  '):
            props.append(l)
    return L[0].strip(), elem, np.asarray(xyz, dtype=np.float64), bonds, L[:4 + na + nb] + props


def transformation_mask(n, bonds):
    """get_transformation_mask (utils/torsion.py:13-61): a bond is rotatable when removing it disconnects the graph
    and the smaller side has more than one atom; the mask row marks that side."""
    G = nx.Graph()
    G.add_nodes_from(range(n))
    G.add_edges_from(bonds)
    mask_edges, rows = [], []
    for b, e in bonds:
        G2 = G.copy()
        G2.remove_edge(b, e)
        if not nx.is_connected(G2):
            side = list(sorted(nx.connected_components(G2), key=len)[0])
            if len(side) > 1:
                row = np.zeros(n, dtype=bool)
                row[side] = True
                rows.append(row)
                mask_edges += [False, True] if b in side else [True, False]
                continue
        mask_edges += [False, False]
    return np.asarray(mask_edges, dtype=bool), (np.stack(rows) if rows else np.zeros((0, n), dtype=bool))


def ligand_graph_from_sdf(path, graph, rng=None):
    """Fill graph['ligand'] / graph['ligand','lig_bond','ligand'] from a 3-D SD file (heavy atoms only, remove_hs=True)."""
    rng = rng or random.Random(0)
    name, elem, xyz, bonds, raw = read_sdf_block(open(path).read())
    heavy = [i for i, e in enumerate(elem) if e != 'H']
    idx = {a: k for k, a in enumerate(heavy)}
    n = len(heavy)
    num_h = np.zeros(n, dtype=int)
    hb = []
    for a, b, o in bonds:
        if a in idx and b in idx:
            hb.append((idx[a], idx[b], o))
        elif a in idx:
            num_h[idx[a]] += 1
        elif b in idx:
            num_h[idx[b]] += 1
    G = nx.Graph()
    G.add_nodes_from(range(n))
    G.add_edges_from([(a, b) for a, b, _ in hb])
    rings = nx.minimum_cycle_basis(G)
    el = [elem[a] for a in heavy]
    pos = xyz[heavy]
    dbl, tpl = np.zeros(n, bool), np.zeros(n, bool)
    for a, b, o in hb:
        if o == 2:
            dbl[a] = dbl[b] = True
        if o == 3:
            tpl[a] = tpl[b] = True
    arom = np.zeros(n, bool)
    for r in rings:
        if len(r) in (5, 6) and all(dbl[a] or el[a] in ('N', 'O', 'S') for a in r) and sum(dbl[a] for a in r) >= len(r) - 2:
            arom[list(r)] = True
    deg = np.asarray([G.degree(a) for a in range(n)])
    x = np.zeros((n, 16), dtype=np.int64)
    for a in range(n):
        hyb = 0 if tpl[a] else (1 if (dbl[a] or arom[a]) else 2)
        x[a] = [_Z.get(el[a], 119) - 1, 0, min(deg[a] + num_h[a], 11), 5, min(num_h[a], 7), min(num_h[a], 9), 0, hyb,
                int(arom[a]), min(sum(a in r for r in rings), 7)] + \
               [int(any(a in r and len(r) == s for r in rings)) for s in (3, 4, 5, 6, 7, 8)]
    lig = graph['ligand']
    lig.x = torch.from_numpy(x)
    ei, et = [], []
    for a, b, o in hb:
        t = 3 if (arom[a] and arom[b] and any(a in r and b in r for r in rings)) else min(o, 3) - 1
        ei += [(a, b), (b, a)]
        et += [t, t]
    graph['ligand', 'lig_bond', 'ligand'].edge_index = torch.tensor(ei, dtype=torch.long).T.contiguous()
    graph['ligand', 'lig_bond', 'ligand'].edge_attr = torch.nn.functional.one_hot(torch.tensor(et), 4).float()
    me, mr = transformation_mask(n, [(a, b) for a, b, _ in hb])
    lig.edge_mask, lig.mask_rotate = torch.from_numpy(me), mr
    fp = np.zeros((n, 11), dtype=np.float32)
    for a in range(n):
        nb = list(G.neighbors(a))
        if el[a] in ('N', 'O') and num_h[a] > 0:
            fp[a, 1] = 1
        if el[a] == 'O' or (el[a] == 'N' and num_h[a] == 0 and deg[a] < 3):
            fp[a, 4] = fp[a, 0] = 1
        if arom[a]:
            fp[a, 2] = fp[a, 8] = 1
        if (el[a] == 'C' and all(el[b] == 'C' for b in nb)) or el[a] in ('Cl', 'Br', 'I', 'F'):
            fp[a, 5] = 1
        if el[a] in ('Cl', 'Br', 'I'):
            fp[a, 9] = 1
    norm = np.zeros((n, 11, 3), dtype=np.float32)
    a1, a2 = np.zeros((n, 11), np.float32), np.zeros((n, 11), np.float32)
    for a in range(n):
        nbc = [pos[b] for b in G.neighbors(a)]
        if not nbc:
            continue
        root = np.mean(nbc, axis=0)
        for t in range(11):
            if fp[a, t] == 0:
                continue
            if PHORETYPES[t] == 'AR':
                if len(nbc) < 2:
                    continue
                two = rng.sample(nbc, 2)
                c = np.cross(two[0] - pos[a], two[1] - pos[a])
                norm[a, t] = c / (np.linalg.norm(c) + 1e-12)
                a1[a, t], a2[a, t] = 0.0, PI
            else:
                c = pos[a] - root
                norm[a, t] = c / (np.linalg.norm(c) + 1e-12)
                if PHORETYPES[t] in ('MB', 'HA', 'HD') and len(nbc) == 1:
                    a1[a, t] = a2[a, t] = PI / 3.0
    lig.phorefp, lig.norm = torch.from_numpy(fp), torch.from_numpy(norm.reshape(n, 33))
    lig.norm_angle1, lig.norm_angle2 = torch.from_numpy(a1), torch.from_numpy(a2)
    lig.ph = torch.from_numpy(fp.sum(0))
    lig.pos = torch.from_numpy(pos).float()
    graph.sdf_template = dict(lines=raw, heavy=heavy, name=name)
    return graph


def _reindexed_property_lines(props, keep):
    """`M  CHG` / `M  ISO` / `M  RAD` lines of the input mol block (count + (atom, value) pairs in 4-column fields) re-indexed to
    the heavy-atom numbering; entries on removed hydrogens are dropped, other `M  ` lines are passed through unchanged."""
    out = []
    for l in props:
        tag = l[3:6]
        if tag in ('CHG', 'ISO', 'RAD'):
            f = l[6:].split()
            pairs = [(int(f[1 + 2 * k]) - 1, f[2 + 2 * k]) for k in range(int(f[0]))]
            pairs = [(keep[a] + 1, v) for a, v in pairs if a in keep]
            for k in range(0, len(pairs), 8):
                part = pairs[k:k + 8]
                out.append(f'M  {tag}{len(part):3d}' + ''.join(f'{a:4d}{int(v):4d}' for a, v in part))
        elif not l.startswith('M  END'):
            out.append(l)
    return out


def write_mol_with_multi_coords(template, multi_new_coords, path, name, marker='', properties=None):
    """SD file with one record per pose: the heavy-atom coordinates of the input mol block are substituted
    (process_mols.py:888-921 does the same through RDKit on the H-stripped molecule; hydrogens keep their input
    coordinates here and are dropped from the record to stay consistent)."""
    lines, heavy = template['lines'], template['heavy']
    na, nb = int(lines[3][0:3]), int(lines[3][3:6])
    keep = {a: k for k, a in enumerate(heavy)}
    bond_lines = []
    for l in lines[4 + na:4 + na + nb]:
        a, b = int(l[0:3]) - 1, int(l[3:6]) - 1
        if a in keep and b in keep:
            bond_lines.append(f'{keep[a] + 1:3d}{keep[b] + 1:3d}' + l[6:])
    prop_lines = _reindexed_property_lines(lines[4 + na + nb:], keep)
    with open(path, 'w') as fh:
        for i, coords in enumerate(multi_new_coords):
            fh.write(f'{name}_{marker}_{i}\n{lines[1]}\n{lines[2]}\n')
            fh.write(f'{len(heavy):3d}{len(bond_lines):3d}' + lines[3][6:] + '\n')
            for k, a in enumerate(heavy):
                x, y, z = (float(v) for v in coords[k])
                fh.write(f'{x:10.4f}{y:10.4f}{z:10.4f}' + lines[4 + a][30:] + '\n')
            for l in bond_lines:
                fh.write(l + '\n')
            for l in prop_lines:
                fh.write(l + '\n')
            fh.write('M  END\n')
            if properties:
                for key, vals in properties.items():
                    fh.write(f'>  <{key}>  ({i + 1}) \n{vals[i]}\n\n')
            fh.write('$$$$\n')
