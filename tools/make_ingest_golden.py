"""Build-container-only: golden vectors for the data formats either side of the denoising path,
produced by the reference's OWN functions and shipped example outputs, committed as tests/golden/ingest.npz.

    python tools/make_ingest_golden.py          (needs /root/reference; RDKit / PyG are stubbed, they are not used by
                                                 the functions called here)
Contents
* `.phore` ingestion: `parse_phore` + `get_phore_graph` (src/datasets/process_pharmacophore.py:78-152, 634-714) on the shipped
  example pharmacophore -> x, pos, norm, edge_index (the file text itself is stored so the test needs no /root/reference);
* rotatable-bond masks: `get_transformation_mask` (src/utils/torsion.py:13-61) on the heavy-atom bond graphs of the 18 shipped
  example ligands (ring systems included);
* AncPhore `.score` parsing: `parse_score_file` (process_pharmacophore.py:885-925) on the shipped example score file, every
  `fitness` column and `return_all`;
* pose output known-answer test: the input ligand SD file, the 40 poses of the reference's shipped docked SD file
  (examples/output/1/mapping_process/...) and the shipped AncPhore scores of exactly those poses; the reference's
  `ranked_results.csv` and the record names / fitscore tags / first-atom coordinates of its `ranked_poses/*_ranked.sdf`.
"""
import importlib.util
import os
import sys
import types
from unittest import mock

import networkx as nx
import numpy as np
import torch

REF = '/root/reference'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffphore_b200.graph import HeteroGraph        # noqa: E402  (attribute container standing in for HeteroData)


def _load(name, path, stubs):
    for m in stubs:
        sys.modules.setdefault(m, mock.MagicMock())
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def read_sdf(path):
    L = open(path).read().split('\n')
    na, nb = int(L[3][0:3]), int(L[3][3:6])
    elem = [l[31:34].strip() for l in L[4:4 + na]]
    xyz = np.asarray([[float(l[0:10]), float(l[10:20]), float(l[20:30])] for l in L[4:4 + na]])
    bonds = [(int(l[0:3]) - 1, int(l[3:6]) - 1) for l in L[4 + na:4 + na + nb]]
    return elem, xyz, bonds


def main():
    out = {}
    # ---- .phore ingestion -------------------------------------------------------------------------------------------
    pp = _load('ref_process_pharmacophore', os.path.join(REF, 'src/datasets/process_pharmacophore.py'),
               ['rdkit', 'rdkit.Chem', 'datasets', 'datasets.process_mols'])
    phore_file = os.path.join(REF, 'examples/phore/sQC_QFA_complex.phore')
    out['phore_text'] = np.asarray(open(phore_file).read())
    phore = pp.parse_phore(phore_file)[0]
    g = HeteroGraph()
    pp.get_phore_graph(phore, g, consider_ex=True, neighbor_cutoff=5.0, ex_connected=True)
    out['phore_id'] = np.asarray(phore.id)
    out['phore_x'] = g['phore'].x.numpy()
    out['phore_pos'] = g['phore'].pos.numpy()
    out['phore_norm'] = g['phore'].norm.numpy()
    out['phore_edge_index'] = g['phore', 'phore_contact', 'phore'].edge_index.numpy()

    # ---- get_transformation_mask ------------------------------------------------------------------------------------
    tg = types.ModuleType('torch_geometric'); tgu = types.ModuleType('torch_geometric.utils'); tgd = types.ModuleType('torch_geometric.data')

    def to_networkx(data, to_undirected=False):
        G = nx.DiGraph()
        G.add_nodes_from(range(data.num_nodes))
        G.add_edges_from(data.edge_index.T.tolist())
        return G
    tgu.to_networkx = to_networkx
    tgd.Data = object
    sys.modules.update({'torch_geometric': tg, 'torch_geometric.utils': tgu, 'torch_geometric.data': tgd})
    tor = _load('ref_torsion', os.path.join(REF, 'src/utils/torsion.py'), [])

    class _Pyg:                                       # the two accessors get_transformation_mask uses
        def __init__(self, n, ei):
            self.n, self.ei = n, ei

        def to_homogeneous(self):
            return types.SimpleNamespace(num_nodes=self.n, edge_index=self.ei)

        def __getitem__(self, key):
            return types.SimpleNamespace(edge_index=self.ei)

    lig_dir = os.path.join(REF, 'examples/ligands')
    names = sorted(f[:-4] for f in os.listdir(lig_dir) if f.endswith('.sdf'))
    out['lig_names'] = np.asarray(names)
    out['task_file_text'] = np.asarray(open(os.path.join(REF, 'examples/task_file.csv')).read())      # cfg3: 15 pairs
    for nm in names:
        elem, xyz, bonds = read_sdf(os.path.join(lig_dir, nm + '.sdf'))
        heavy = [i for i, e in enumerate(elem) if e != 'H']
        idx = {a: k for k, a in enumerate(heavy)}
        hb = [(idx[a], idx[b]) for a, b in bonds if a in idx and b in idx]
        ei = torch.tensor([p for a, b in hb for p in ((a, b), (b, a))], dtype=torch.long).T
        me, mr = tor.get_transformation_mask(_Pyg(len(heavy), ei))
        out[f'lig_text_{nm}'] = np.asarray(open(os.path.join(lig_dir, nm + '.sdf')).read())
        out[f'mask_edges_{nm}'] = me
        out[f'mask_rotate_{nm}'] = mr

    # ---- .score parsing + pose-output KAT ----------------------------------------------------------------------------
    d = os.path.join(REF, 'examples/output/1/mapping_process/sQC_Substrate__STK936575')
    score_file = os.path.join(d, 'sQC_Substrate__STK936575.score')
    out['score_text'] = np.asarray(open(score_file).read())
    for f in range(1, 7):
        out[f'score_fitness_{f}'] = np.asarray(pp.parse_score_file(score_file, fitness=f))
    out['score_all'] = np.asarray(pp.parse_score_file(score_file, return_all=True))
    poses, rec = [], []
    L = open(os.path.join(d, 'sQC_Substrate__STK936575.sdf')).read().split('\n')
    i = 0
    while i + 3 < len(L):
        na = int(L[i + 3][0:3])
        poses.append([[float(l[0:10]), float(l[10:20]), float(l[20:30])] for l in L[i + 4:i + 4 + na]])
        rec.append(L[i])
        while L[i] != '$$$$':
            i += 1
        i += 1
    out['kat_poses'] = np.asarray(poses, dtype=np.float64)
    out['kat_record_names'] = np.asarray(rec)
    # the reference's own summary of that run (analyze_results, inference.py:321-350) and the head of its ranked SD file
    out['ranked_results_text'] = np.asarray(open(os.path.join(REF, 'examples/output/1/ranked_results.csv')).read())
    R = open(os.path.join(REF, 'examples/output/1/ranked_poses/sQC_Substrate__STK936575_ranked.sdf')).read().split('\n')
    rnames, fits, first = [], [], []
    for i, l in enumerate(R):
        if l.startswith('sQC_Substrate__STK936575_rank_'):
            rnames.append(l)
            first.append([float(R[i + 4][0:10]), float(R[i + 4][10:20]), float(R[i + 4][20:30])])
        if l.startswith('>  <fitscore>'):
            fits.append(R[i + 1])
    out['ranked_names'], out['ranked_fitscore_text'], out['ranked_first_atom'] = np.asarray(rnames), np.asarray(fits), np.asarray(first)
    # ---- get_perfect_similarity (inference.py:273-312), extracted unmodified from the reference's inference.py
    import ast
    src = open(os.path.join(REF, 'src/inference.py')).read()
    fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == 'get_perfect_similarity'][0]
    ns = {'torch': torch}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), 'inference.py', 'exec'), ns)
    rng = np.random.RandomState(5)
    ptypes = rng.randint(0, 11, size=(6, 9))
    ligph = rng.randint(0, 4, size=(6, 11)).astype(np.float32)
    ptypes[5] = 10                                                         # only exclusion spheres: weighted volume 0 -> -1
    sims = []
    for pt, lp in zip(ptypes, ligph):
        gg = HeteroGraph()
        gg['phore'].phoretype = torch.nn.functional.one_hot(torch.from_numpy(pt), 11).float()
        gg['ligand'].ph = torch.from_numpy(lp)
        gg.name = 'x'
        sims.append(ns['get_perfect_similarity'](gg))
    out['sim_phore_types'], out['sim_lig_ph'], out['sim_values'] = ptypes, ligph, np.asarray(sims, dtype=np.float64)
    # ---- read_input (inference.py:99-137), extracted unmodified, on a small directory tree (rebuilt by the test)
    import json, tempfile, pandas
    fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == 'read_input'][0]
    ns = {'os': os, 'pd': pandas}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), 'inference.py', 'exec'), ns)
    root = tempfile.mkdtemp()
    os.makedirs(os.path.join(root, 'phores')); os.makedirs(os.path.join(root, 'ligs'))
    for f in ('a.phore', 'b.phore'):
        open(os.path.join(root, 'phores', f), 'w').write('x')
    for f in ('l1.sdf', 'l2.sdf', 'l3.sdf'):
        open(os.path.join(root, 'ligs', f), 'w').write('x')
    open(os.path.join(root, 'lig.smi'), 'w').write('CCO\nc1ccccc1\n')
    open(os.path.join(root, 'task.csv'), 'w').write('ligand_description,phore\nligs/l1.sdf,phores/a.phore\nligs/l2.sdf,phores/a.phore\n'
                                                    'ligs/l1.sdf,phores/a.phore\n')
    cases = {'csv': (os.path.join(root, 'task.csv'), None, None),
             'single': (None, os.path.join(root, 'phores/a.phore'), os.path.join(root, 'ligs/l1.sdf')),
             'dirs': (None, os.path.join(root, 'phores'), os.path.join(root, 'ligs')),
             'smi': (None, os.path.join(root, 'phores/b.phore'), os.path.join(root, 'lig.smi'))}
    res = {}
    for k, a in cases.items():
        recs = ns['read_input'](*a)
        res[k] = sorted([r['phore'].replace(root, '<ROOT>'), r['ligand_description'].replace(root, '<ROOT>')] for r in recs)
    out['read_input_json'] = np.asarray(json.dumps(res))
    # ---- get_model (utils/utils.py:113-168), extracted unmodified: constructor kwargs it derives from the shipped model_parameters.yml
    import yaml
    from argparse import Namespace
    usrc = open(os.path.join(REF, 'src/utils/utils.py')).read()
    fn = [n for n in ast.parse(usrc).body if isinstance(n, ast.FunctionDef) and n.name == 'get_model'][0]
    captured = {}

    class _Recorder:
        def __init__(self, **kw):
            captured.update(kw)

        def to(self, device):
            return self
    ns = {'PhoreModel': _Recorder, 'get_timestep_embedding': lambda **kw: ('emb', kw), 'torch': torch,
          'DataParallel': lambda m: m}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), 'utils.py', 'exec'), ns)
    margs = Namespace(**yaml.full_load(open(os.path.join(REF, 'weights/diffphore_calibrated_warmuped_ft/model_parameters.yml'))))
    margs.no_torsion = False
    ns['get_model'](margs, torch.device('cpu'), t_to_sigma=None, no_parallel=True)
    kw = {k: v for k, v in captured.items() if k not in ('t_to_sigma', 'device', 'timestep_emb_func')}
    out['get_model_kwargs_json'] = np.asarray(json.dumps(kw, sort_keys=True))
    out['get_model_emb_json'] = np.asarray(json.dumps(captured['timestep_emb_func'][1], sort_keys=True))
    out['model_parameters_yml'] = np.asarray(open(os.path.join(REF, 'weights/diffphore_calibrated_warmuped_ft/model_parameters.yml')).read())
    # ---- parse_args (inference.py:54-96), extracted unmodified: flag names and defaults of the CLI
    from argparse import ArgumentParser, FileType
    fns = {n.name: n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef)}
    ns = {'ArgumentParser': ArgumentParser, 'FileType': FileType, 'Namespace': Namespace, 'sys': sys, 'os': os}
    for nm in ('str2bool', 'parse_args'):
        exec(compile(ast.Module(body=[fns[nm]], type_ignores=[]), 'inference.py', 'exec'), ns)
    argv, sys.argv = sys.argv, ['inference.py']
    out['parse_args_defaults_json'] = np.asarray(json.dumps(vars(ns['parse_args']()), sort_keys=True))
    sys.argv = ['inference.py', '--target_fishing', 'true', '--no_random', '--ode', '--overwrite', 'yes', '--cutoff', '0.4']
    out['parse_args_flags_json'] = np.asarray(json.dumps(vars(ns['parse_args']()), sort_keys=True))
    sys.argv = argv
    path = os.path.join(ROOT, 'tests/golden/ingest.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path), 'bytes;', len(names), 'ligands,', out['kat_poses'].shape, 'poses')


if __name__ == '__main__':
    main()
