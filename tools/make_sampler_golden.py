"""Build-container-only: pin the sampler half of the oracle (SURVEY a-1, a-2, a-3, a-14, a-15) on the reference's OWN code,
committed as tests/golden/ref_sampler.npz.

    python tools/make_sampler_golden.py         (needs /root/reference; fresh process: it replaces sys.modules entries)
The UNMODIFIED reference modules utils/geometry.py, utils/torsion.py, utils/diffusion_utils.py and utils/sampling.py are loaded
over dependency shims (oracle/ref_shims.py + a DataLoader / to_networkx shim; rdkit and the dataset modules are stubs, nothing of
them runs) and executed on seeded inputs:
* randomize_position (sampling.py:16-63) under fixed numpy / torch seeds; the draws it consumes are replayed in the same order and
  stored, so that oracle.sampler.randomize_position can be fed the same draws;
* modify_conformer (diffusion_utils.py:23-79, with torsion.py:64-109 and geometry.py:71-136) with and without torsion updates;
* get_t_schedule, t_to_sigma, get_timestep_embedding('sinusoidal', 20, 10000) (diffusion_utils.py:16-20,82-145);
* sampling_phore (sampling.py:174-280) end to end, 6 steps, no_random=True and ode=True variants, driving the UNMODIFIED reference
  TensorProductScoreModel (shipped checkpoint, over the same shims as tools/make_golden.py) on a real-shaped pair x 3 samples.
"""
import importlib.util
import os
import sys
import types
from types import SimpleNamespace
from unittest import mock

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = '/root/reference/src'
sys.path.insert(0, ROOT)
from diffphore_b200 import graph as G                                       # noqa: E402
from oracle import ref_shims                                               # noqa: E402
from oracle.tables import So3ScoreNorm, TorusScoreNorm                     # noqa: E402
from oracle.model import sinusoidal_embedding                              # noqa: E402
from tests.parity_util import load_pairs                                   # noqa: E402

CKPT = '/root/reference/weights/diffphore_calibrated_warmuped_ft/best_ema_inference_epoch_model.pt'
ARGS = SimpleNamespace(tr_sigma_min=0.1, tr_sigma_max=5.0, rot_sigma_min=0.1, rot_sigma_max=1.5, tor_sigma_min=0.0314,
                       tor_sigma_max=3.14, no_torsion=False, keep_update=False)


def load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF_SRC, rel))
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


def main():
    so3n, torn = So3ScoreNorm(), TorusScoreNorm(seed=0)
    smp = ref_shims.install(so3n, torn)                                    # e3nn / cluster / scatter / utils.so3 / utils.torus shims
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
    geometry = load('utils.geometry', 'utils/geometry.py')
    torsion = load('utils.torsion', 'utils/torsion.py')
    du = load('utils.diffusion_utils', 'utils/diffusion_utils.py')
    sampling = load('utils.sampling', 'utils/sampling.py')
    out = {}

    # ---- schedule, sigmas, embedding
    sched = du.get_t_schedule(inference_steps=20)
    out['t_schedule'] = np.asarray(sched)
    out['t_to_sigma'] = np.asarray([du.t_to_sigma(t, t, t, ARGS) for t in sched])
    emb = du.get_timestep_embedding('sinusoidal', 20, 10000)
    out['sigma_emb'] = torch.stack([emb(torch.tensor([float(t)])) for t in sched]).squeeze(1).numpy()

    # ---- randomize_position with replayed draws
    graphs = load_pairs('synthetic', 2, 14, 5) + load_pairs('real', 1)
    dl = [g.clone() for g in graphs for _ in range(2)]
    for g in dl:
        g['ligand'].mask_rotate = [g['ligand'].mask_rotate]                # PyG batches hand the array over as a one-element list
    np.random.seed(1234)
    torch.manual_seed(1234)
    sampling.randomize_position(dl, False, False, ARGS.tr_sigma_max)
    np.random.seed(1234)
    torch.manual_seed(1234)
    from scipy.spatial.transform import Rotation as R
    tor = [np.random.uniform(low=-np.pi, high=np.pi, size=int(g['ligand'].edge_mask.sum())) for g in dl]
    rot, tr = [], []
    for _ in dl:
        rot.append(R.random().as_matrix())
        tr.append(torch.normal(mean=0, std=ARGS.tr_sigma_max, size=(1, 3)).numpy())
    out['rand_tor'] = np.concatenate(tor)
    out['rand_rot'] = np.stack(rot)
    out['rand_tr'] = np.concatenate(tr)
    out['rand_pos'] = torch.cat([g['ligand'].pos for g in dl]).numpy()
    out['rand_norm'] = torch.cat([g['ligand'].norm.reshape(g['ligand'].pos.shape[0], -1) for g in dl]).numpy()

    # ---- modify_conformer (with torsions / rigid only), continuing from the randomised poses
    rng = np.random.RandomState(7)
    upd_tr, upd_rot, upd_tor, res_pos, res_norm, rigid_pos, rigid_norm = [], [], [], [], [], [], []
    for g in dl:
        n_rot = int(g['ligand'].edge_mask.sum())
        trp = torch.from_numpy(rng.randn(1, 3).astype(np.float32) * 0.7)
        rotp = torch.from_numpy(rng.randn(3).astype(np.float32) * 0.4)
        torp = (rng.randn(n_rot) * 0.5).astype(np.float32)
        torp[::4] = 0.0                                                    # skipped bonds (torsion.py:83-84)
        a = du.modify_conformer(g.clone(), trp, rotp, torp)
        b = du.modify_conformer(g.clone(), trp, rotp, None)
        upd_tr.append(trp.numpy()); upd_rot.append(rotp.numpy()); upd_tor.append(torp)
        res_pos.append(a['ligand'].pos.numpy()); res_norm.append(a['ligand'].norm.numpy())
        rigid_pos.append(b['ligand'].pos.numpy()); rigid_norm.append(b['ligand'].norm.numpy())
    out.update(upd_tr=np.concatenate(upd_tr), upd_rot=np.stack(upd_rot), upd_tor=np.concatenate(upd_tor),
               upd_pos=np.concatenate(res_pos), upd_norm=np.concatenate(res_norm),
               rigid_pos=np.concatenate(rigid_pos), rigid_norm=np.concatenate(rigid_norm))

    # ---- sampling_phore end to end with the reference model
    from functools import partial
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
    sd = torch.load(CKPT, map_location='cpu', weights_only=False)
    model.load_state_dict(sd, strict=True)
    model.eval()
    steps = 6
    sch = du.get_t_schedule(inference_steps=steps)
    start = [g.clone() for g in dl[4:6]] + [dl[4].clone()]               # the real-shaped pair: 3 samples (two distinct poses)
    out['samp_start_pos'] = torch.cat([g['ligand'].pos for g in start]).numpy()
    out['samp_start_norm'] = torch.cat([g['ligand'].norm for g in start]).numpy()
    for tag, kw in (('norandom', dict(no_random=True)), ('ode', dict(ode=True))):
        res, conf = sampling.sampling_phore([g.clone() for g in start], model, steps, sch, sch, sch, torch.device('cpu'), t_to_sigma,
                                            ARGS, batch_size=3, **kw)
        assert conf is None
        out[f'samp_{tag}_pos'] = torch.cat([g['ligand'].pos for g in res]).numpy()
    out['samp_steps'] = np.int64(steps)
    path = os.path.join(ROOT, 'tests/golden/ref_sampler.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path), 'bytes')
    print({k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()
