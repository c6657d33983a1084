"""Build-container-only: turn the reference's shipped examples (examples/phore/*.phore, examples/ligands/*.sdf)
into the tensors of SURVEY §8a-0 WITHOUT RDKit, and commit them as tests/golden/real_pairs.npz.

* Pharmacophore side: EXACT — the reference's own pure-Python parser and graph builder
  (src/datasets/process_pharmacophore.py: parse_phore :78, get_phore_graph :634, phore_featurizer :717) are
  imported from /root/reference with `rdkit` / `datasets.process_mols` stubbed out (they are only needed by other
  functions of that module).
* Ligand side: REDUCED featuriser (heavy atoms of the SDF, explicit-H counts, ring perception with networkx, crude
  aromaticity / hybridisation / pharmacophore typing, norms per process_mols.py:782-858).  Parity between oracle and
  CUDA path only needs identical tensors; these inputs are "real-shaped and in-distribution", not RDKit-exact.
    python tools/make_real_inputs.py
"""
import os, sys, types, math, random
from unittest import mock
import numpy as np, torch, networkx as nx

REF = '/root/reference'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffphore_b200.graph import HeteroGraph, graph_to_arrays
from diffphore_b200.synthetic import _transformation_mask   # only for acyclic; general version below

PHORETYPES = ['MB', 'HD', 'AR', 'PO', 'HA', 'HY', 'NE', 'CV', 'CR', 'XB', 'EX']
PI = 3.1415926


def load_ref_phore_module():
    import importlib.util
    for m in ['rdkit', 'rdkit.Chem', 'datasets', 'datasets.process_mols']:
        sys.modules[m] = mock.MagicMock()
    spec = importlib.util.spec_from_file_location('ref_process_pharmacophore',
                                                  os.path.join(REF, 'src/datasets/process_pharmacophore.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def read_sdf(path):
    L = open(path).read().split('\n')
    na, nb = int(L[3][0:3]), int(L[3][3:6])
    elem, xyz = [], []
    for l in L[4:4 + na]:
        xyz.append([float(l[0:10]), float(l[10:20]), float(l[20:30])])
        elem.append(l[31:34].strip())
    bonds = [(int(l[0:3]) - 1, int(l[3:6]) - 1, int(l[6:9])) for l in L[4 + na:4 + na + nb]]
    return elem, np.asarray(xyz), bonds


Z = {'H': 1, 'C': 6, 'N': 7, 'O': 8, 'F': 9, 'P': 15, 'S': 16, 'Cl': 17, 'Br': 35, 'I': 53}


def general_transformation_mask(n, bonds):
    """get_transformation_mask (torsion.py:13-61) with networkx, cyclic molecules included."""
    G = nx.Graph()
    G.add_nodes_from(range(n))
    G.add_edges_from(bonds)
    mask_edges, rows = [], []
    for b, e in bonds:
        G2 = G.copy()
        G2.remove_edge(b, e)
        if not nx.is_connected(G2):
            l = list(sorted(nx.connected_components(G2), key=len)[0])
            if len(l) > 1:
                row = np.zeros(n, dtype=bool); row[l] = True
                rows.append(row)
                mask_edges += [False, True] if b in l else [True, False]
                continue
        mask_edges += [False, False]
    return np.asarray(mask_edges, dtype=bool), (np.stack(rows) if rows else np.zeros((0, n), dtype=bool))


def ligand_graph(path, rng):
    elem, xyz, bonds = read_sdf(path)
    heavy = [i for i, e in enumerate(elem) if e != 'H']
    idx = {a: k for k, a in enumerate(heavy)}
    n = len(heavy)
    numH = np.zeros(n, dtype=int)
    hb = []
    for a, b, o in bonds:
        if a in idx and b in idx:
            hb.append((idx[a], idx[b], o))
        elif a in idx:
            numH[idx[a]] += 1
        elif b in idx:
            numH[idx[b]] += 1
    G = nx.Graph(); G.add_nodes_from(range(n)); G.add_edges_from([(a, b) for a, b, _ in hb])
    rings = nx.minimum_cycle_basis(G)
    el = [elem[a] for a in heavy]
    pos = xyz[heavy]
    has_double = np.zeros(n, bool); has_triple = np.zeros(n, bool)
    for a, b, o in hb:
        if o == 2: has_double[a] = has_double[b] = True
        if o == 3: has_triple[a] = has_triple[b] = True
    arom = np.zeros(n, bool)
    for r in rings:
        if len(r) in (5, 6) and all(has_double[a] or el[a] in ('N', 'O', 'S') for a in r) \
                and sum(has_double[a] for a in r) >= len(r) - 2:
            arom[list(r)] = True
    deg = np.asarray([G.degree(a) for a in range(n)])
    x = np.zeros((n, 16), dtype=np.int64)
    for a in range(n):
        hyb = 0 if has_triple[a] else (1 if (has_double[a] or arom[a]) else 2)
        nring = sum(a in r for r in rings)
        x[a] = [Z.get(el[a], 119) - 1, 0, min(deg[a] + numH[a], 11), 5, min(numH[a], 7), min(numH[a], 9), 0, hyb,
                int(arom[a]), min(nring, 7)] + [int(any(a in r and len(r) == s for r in rings)) for s in (3, 4, 5, 6, 7, 8)]
    g = HeteroGraph()
    lig = g['ligand']
    lig.x = torch.from_numpy(x)
    ei, et = [], []
    for a, b, o in hb:
        t = 3 if (arom[a] and arom[b] and any(a in r and b in r for r in rings)) else min(o, 3) - 1
        ei += [(a, b), (b, a)]; et += [t, t]
    g['ligand', 'ligand'].edge_index = torch.tensor(ei).T.long()
    g['ligand', 'ligand'].edge_attr = torch.nn.functional.one_hot(torch.tensor(et), 4).float()
    me, mr = general_transformation_mask(n, [(a, b) for a, b, _ in hb])
    lig.edge_mask = torch.from_numpy(me); lig.mask_rotate = mr
    # crude pharmacophore typing
    fp = np.zeros((n, 11), dtype=np.float32)
    for a in range(n):
        nb = list(G.neighbors(a))
        if el[a] in ('N', 'O') and numH[a] > 0: fp[a, 1] = 1                                  # HD
        if el[a] == 'O' or (el[a] == 'N' and numH[a] == 0 and deg[a] < 3): fp[a, 4] = 1; fp[a, 0] = 1   # HA, MB
        if arom[a]: fp[a, 2] = 1; fp[a, 8] = 1                                               # AR, CR
        if (el[a] == 'C' and all(el[b] == 'C' for b in nb)) or el[a] in ('Cl', 'Br', 'I', 'F'): fp[a, 5] = 1   # HY
        if el[a] in ('Cl', 'Br', 'I'): fp[a, 9] = 1                                          # XB
    norm = np.zeros((n, 11, 3), dtype=np.float32); a1 = np.zeros((n, 11), np.float32); a2 = np.zeros((n, 11), np.float32)
    for a in range(n):
        nbc = [pos[b] for b in G.neighbors(a)]
        root = np.mean(nbc, axis=0)
        for t in range(11):
            if fp[a, t] == 0: continue
            if PHORETYPES[t] == 'AR':
                if len(nbc) < 2: continue
                two = rng.sample(nbc, 2)
                c = np.cross(two[0] - pos[a], two[1] - pos[a]); norm[a, t] = c / (np.linalg.norm(c) + 1e-12)
                a1[a, t], a2[a, t] = 0.0, PI
            else:
                c = pos[a] - root; norm[a, t] = c / (np.linalg.norm(c) + 1e-12)
                if PHORETYPES[t] in ('MB', 'HA', 'HD') and len(nbc) == 1: a1[a, t] = a2[a, t] = PI / 3.0
    lig.phorefp = torch.from_numpy(fp); lig.norm = torch.from_numpy(norm.reshape(n, 33))
    lig.norm_angle1 = torch.from_numpy(a1); lig.norm_angle2 = torch.from_numpy(a2)
    lig.pos = torch.from_numpy(pos).float()
    return g


def main():
    pp = load_ref_phore_module()
    phore = pp.parse_phore(os.path.join(REF, 'examples/phore/sQC_QFA_complex.phore'))[0]
    rng = random.Random(0)
    out, names = {}, []
    ligs = sorted(os.listdir(os.path.join(REF, 'examples/ligands')))
    for k, f in enumerate(ligs):
        g = ligand_graph(os.path.join(REF, 'examples/ligands', f), rng)
        pp.get_phore_graph(phore, g, consider_ex=True, neighbor_cutoff=5.0, ex_connected=True)
        ph = g['phore']
        ph.x = ph.x.float()
        # generate_graph (pdbbind_phore.py:1143-1188): phoretype one-hot and centring on the phore centroid;
        # the example ligands are not posed in the pharmacophore frame, so put the ligand centroid there
        # (randomize_position re-poses it anyway).
        ph.phoretype = torch.nn.functional.one_hot(ph.x[:, 0].long(), 11).float()
        c = ph.pos.mean(0, keepdim=True)
        ph.pos = ph.pos - c
        g['ligand'].pos = g['ligand'].pos - g['ligand'].pos.mean(0, keepdim=True)
        out.update(graph_to_arrays(g, prefix=f'p{k}_'))
        names.append(f[:-4])
        print(f, g['ligand'].x.shape[0], 'atoms', int(g['ligand'].edge_mask.sum()), 'rot bonds', ph.pos.shape[0], 'phore nodes',
              g['phore', 'phore'].edge_index.shape[1], 'phore edges')
    out['names'] = np.asarray(names)
    np.savez_compressed(os.path.join(ROOT, 'tests/golden/real_pairs.npz'), **out)


if __name__ == '__main__':
    main()
