"""Build-container-only: pin the SURVEY 8f-4 drivers on the reference's OWN code, committed as tests/golden/ref_rank4.npz.

    python tools/make_rank4_golden.py         (needs /root/reference; fresh process: it replaces sys.modules entries)
The UNMODIFIED reference utils/sampling.py (with geometry.py, torsion.py, diffusion_utils.py) is loaded over the dependency shims
of tools/make_sampler_golden.py and drives the UNMODIFIED reference TensorProductScoreModel (shipped checkpoint):
* sample_step (sampling.py:501-559) on a batch of 3 samples of the real-shaped example pair, torch.manual_seed(99);
* sampling_phore_with_fitscore (sampling.py:283-444) with random_samples = 3, 3 steps, torch.manual_seed(77), and
  calculate_fitscore replaced by a deterministic surrogate (minus the distance of the pose's centroid from the origin) - AncPhore in
  the loop would make the fixture depend on a closed binary; the selection arithmetic around it is what is pinned;
* get_updates_from_0_to_n (sampling.py:566-597).
"""
import os
import sys
import types
from functools import partial
from types import SimpleNamespace
from unittest import mock

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tools'))
from diffphore_b200 import graph as G                                       # noqa: E402
from oracle import ref_shims                                               # noqa: E402
from oracle.tables import So3ScoreNorm, TorusScoreNorm                     # noqa: E402
from oracle.model import sinusoidal_embedding                              # noqa: E402
from tests.parity_util import load_pairs                                   # noqa: E402
from make_sampler_golden import load, CKPT                                 # noqa: E402

ARGS = SimpleNamespace(tr_sigma_min=0.1, tr_sigma_max=5.0, rot_sigma_min=0.1, rot_sigma_max=1.5, tor_sigma_min=0.0314,
                       tor_sigma_max=3.14, no_torsion=False, keep_update=False, random_samples=3)


def surrogate_fitscore(args, ligand_pos, name, mol, store_ranked_pose=True, phore_file=None):
    return [float(-np.linalg.norm(np.asarray(p).mean(0))) for p in ligand_pos]


def main():
    so3n, torn = So3ScoreNorm(), TorusScoreNorm(seed=0)
    smp = ref_shims.install(so3n, torn)
    import networkx as nx
    tgu = sys.modules['torch_geometric.utils']
    tgu.to_networkx = lambda data, to_undirected=False: nx.DiGraph()
    sys.modules['torch_geometric.data'] = types.ModuleType('torch_geometric.data')
    sys.modules['torch_geometric.data'].Data = object
    sys.modules['torch_geometric.loader'] = types.ModuleType('torch_geometric.loader')
    sys.modules['torch_geometric.loader'].DataLoader = G.DataLoader
    for m in ('rdkit', 'rdkit.Chem'):
        sys.modules[m] = mock.MagicMock()
    sys.modules['datasets.process_pharmacophore'].calc_phore_fitting = None
    sys.modules['datasets.process_mols'].write_mol_with_multi_coords = None
    load('utils.geometry', 'utils/geometry.py')
    load('utils.torsion', 'utils/torsion.py')
    du = load('utils.diffusion_utils', 'utils/diffusion_utils.py')
    sampling = load('utils.sampling', 'utils/sampling.py')
    t_to_sigma = partial(du.t_to_sigma, args=ARGS)
    emb_f = lambda x: sinusoidal_embedding(10000 * x, 20)
    model = smp.TensorProductScoreModel(t_to_sigma=t_to_sigma, device=torch.device('cpu'), no_torsion=False, timestep_emb_func=emb_f,
        num_conv_layers=4, lig_max_radius=5.0, scale_by_sigma=True, sigma_embed_dim=20, ns=20, nv=10, distance_embed_dim=20,
        cross_distance_embed_dim=20, batch_norm=True, dropout=0.1, use_second_order_repr=False, cross_max_distance=25.0,
        dynamic_max_cross=False, confidence_mode=False, consider_norm=True, use_phore_rule=True, auto_phorefp=False,
        angle_match=True, cross_distance_transition=True, phore_direction_transition=True, phoretype_match_transition=True,
        new=True, ex_factor=-2.0, boarder=True, by_radius=False, clash_tolerance=0.4, clash_cutoff=[1.0, 2.0, 3.0, 4.0, 5.0],
        use_att=False, use_phore_match_feat=True, num_confidence_outputs=1, atom_weight='phore', trioformer_layer=2,
        contrastive_model=None, contrastive_node=True, norm_by_ph=False, dist_for_fitscore=False, angle_for_fitscore=False,
        type_for_fitscore=False, sigmoid_for_fitscore=False, readout='mean', as_exp=False, scaler=100.0)
    model.load_state_dict(torch.load(CKPT, map_location='cpu', weights_only=False), strict=True)
    model.eval()
    gold = np.load(os.path.join(ROOT, 'tests/golden/ref_sampler.npz'))
    base = load_pairs('real', 1)[0]
    n = base['ligand'].pos.shape[0]
    start = []
    for k in range(3):
        g = base.clone()
        g['ligand'].pos = torch.from_numpy(gold['samp_start_pos'][k * n:(k + 1) * n]).clone()
        g['ligand'].norm = torch.from_numpy(gold['samp_start_norm'][k * n:(k + 1) * n]).clone()
        g['ligand'].mask_rotate = [g['ligand'].mask_rotate]
        g.name = ['pair']
        g.mol = ['mol']
        g.original_center = torch.tensor([[1.0, -2.0, 0.5]])
        start.append(g)
    out = {}
    # ---- sample_step
    t = 0.6
    batch = G.collate([g.clone() for g in start])
    du.set_time_phore(batch, t, t, t, 3, 'cpu')
    sig = t_to_sigma(t, t, t)
    torch.manual_seed(99)
    dl, tor_p, tr_p, rot_p = sampling.sample_step(batch, model, ARGS, *sig, delta_t=0.05)
    out['step_t'] = np.float64(t)
    out['step_pos'] = torch.cat([g['ligand'].pos for g in dl]).numpy()
    out['step_norm'] = torch.cat([g['ligand'].norm.reshape(n, -1) for g in dl]).numpy()
    out['step_tr'], out['step_rot'], out['step_tor'] = tr_p.numpy(), rot_p.numpy(), np.asarray(tor_p)
    # ---- sampling_phore_with_fitscore, random_samples = 3
    steps = 3
    sch = du.get_t_schedule(inference_steps=steps)
    sampling.calculate_fitscore = surrogate_fitscore
    torch.manual_seed(77)
    res, conf = sampling.sampling_phore_with_fitscore([g.clone() for g in start], model, steps, sch, sch, sch, torch.device('cpu'),
                                                      t_to_sigma, ARGS, batch_size=3)
    assert conf is None
    out['fit_pos'] = torch.cat([g['ligand'].pos for g in res]).numpy()
    out['fit_steps'] = np.int64(steps)
    # ... and with random_samples = 0 (plain loop through the same function)
    args1 = SimpleNamespace(**{**vars(ARGS), 'random_samples': 0})
    torch.manual_seed(78)
    res1, _ = sampling.sampling_phore_with_fitscore([g.clone() for g in start], model, steps, sch, sch, sch, torch.device('cpu'),
                                                    t_to_sigma, args1, batch_size=3)
    out['fit1_pos'] = torch.cat([g['ligand'].pos for g in res1]).numpy()
    # ---- get_updates_from_0_to_n
    g_a, g_b = start[0].clone(), dl[0]
    g_a['ligand'].mask_rotate = g_a['ligand'].mask_rotate[0]
    tor_up = np.asarray(tor_p)[:int(g_a['ligand'].edge_mask.sum())].astype(np.float64)
    t2, r1 = sampling.get_updates_from_0_to_n(g_a, g_b, tor_up)
    out['upd0n_t'], out['upd0n_rot'], out['upd0n_tor'] = t2.numpy(), np.asarray(r1), tor_up
    path = os.path.join(ROOT, 'tests/golden/ref_rank4.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path), 'bytes')
    print({k: np.asarray(v).shape for k, v in out.items()})


if __name__ == '__main__':
    main()
