"""Build-container-only: run the UNMODIFIED reference TensorProductScoreModel (over oracle/ref_shims.py) with the
shipped checkpoint on seeded inputs and commit inputs + outputs as tests/golden/ref_forward.npz.

    python tools/make_golden.py           (needs /root/reference; runs in a fresh process because it replaces
                                           sys.modules entries such as `utils`, `models`, `datasets`)
The fixtures pin oracle/model.py against the reference's own model code (tests/test_oracle.py).
"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from functools import partial
from types import SimpleNamespace
from oracle import ref_shims
from oracle.tables import So3ScoreNorm, TorusScoreNorm
from oracle import sampler as osamp
from oracle.model import OracleScoreModel, sinusoidal_embedding
from diffphore_b200.graph import collate, graph_from_arrays
from diffphore_b200.synthetic import make_pairs
from tests.parity_util import make_draws, oracle_initial_graphs, rel

CKPT = '/root/reference/weights/diffphore_calibrated_warmuped_ft/best_ema_inference_epoch_model.pt'
so3n, torn = So3ScoreNorm(), TorusScoreNorm(seed=0)
smp = ref_shims.install(so3n, torn)
args = SimpleNamespace(tr_sigma_min=0.1, tr_sigma_max=5.0, rot_sigma_min=0.1, rot_sigma_max=1.5, tor_sigma_min=0.0314, tor_sigma_max=3.14)


def t_to_sigma(t_tr, t_rot, t_tor):            # diffusion_utils.t_to_sigma partial
    return (args.tr_sigma_min ** (1 - t_tr) * args.tr_sigma_max ** t_tr, args.rot_sigma_min ** (1 - t_rot) * args.rot_sigma_max ** t_rot,
            args.tor_sigma_min ** (1 - t_tor) * args.tor_sigma_max ** t_tor)


emb = lambda x: sinusoidal_embedding(10000 * x, 20)        # get_timestep_embedding('sinusoidal', 20, 10000)
# kwargs exactly as utils/utils.py:get_model maps model_parameters.yml
model = smp.TensorProductScoreModel(t_to_sigma=t_to_sigma, device=torch.device('cpu'), no_torsion=False, timestep_emb_func=emb,
    num_conv_layers=4, lig_max_radius=5.0, scale_by_sigma=True, sigma_embed_dim=20, ns=20, nv=10, distance_embed_dim=20,
    cross_distance_embed_dim=20, batch_norm=True, dropout=0.1, use_second_order_repr=False, cross_max_distance=25.0,
    dynamic_max_cross=False, confidence_mode=False, consider_norm=True, use_phore_rule=True, auto_phorefp=False,
    angle_match=True, cross_distance_transition=True, phore_direction_transition=True, phoretype_match_transition=True,
    new=True, ex_factor=-2.0, boarder=True, by_radius=False, clash_tolerance=0.4, clash_cutoff=[1.0, 2.0, 3.0, 4.0, 5.0],
    use_att=False, use_phore_match_feat=True, num_confidence_outputs=1, atom_weight='phore', trioformer_layer=2,
    contrastive_model=None, contrastive_node=True, norm_by_ph=False, dist_for_fitscore=False, angle_for_fitscore=False,
    type_for_fitscore=False, sigmoid_for_fitscore=False, readout='mean', as_exp=False, scaler=100.0)
sd = torch.load(CKPT, map_location='cpu', weights_only=False)
print('reference class strict load:', model.load_state_dict(sd, strict=True))
model.eval()
oracle = OracleScoreModel(sd, so3n, torn)

a = np.load(os.path.join(ROOT, 'tests/golden/real_pairs.npz'))
cases = {'real': ([graph_from_arrays(a, f'p{k}_') for k in (11, 0, 5)], 2, 0.55), 'syn': (make_pairs(2, 32, 8), 2, 0.25)}
out = {}
for name, (graphs, S, t) in cases.items():
    init, _, n_rot = make_draws(graphs, S, 7)
    init['tr'] = init['tr'] / 5.0 * (0.1 ** (1 - t) * 5.0 ** t)
    dl = oracle_initial_graphs(graphs, S, init, n_rot)
    b1 = collate([g.clone() for g in dl]); osamp.set_time(b1, t, len(dl))
    with torch.no_grad():
        r = model(b1)
    b2 = collate([g.clone() for g in dl]); osamp.set_time(b2, t, len(dl))
    o = oracle(b2)
    print(name, 'oracle vs reference-over-shims:', [rel(x, y) for x, y in zip(o, r)], 'max |tr|', float(r[0].abs().max()))
    out[f'{name}_t'] = np.float64(t)
    out[f'{name}_pos'] = torch.cat([g['ligand'].pos for g in dl]).numpy()
    out[f'{name}_norm'] = torch.cat([g['ligand'].norm for g in dl]).numpy()
    for k, v in zip(('tr', 'rot', 'tor'), r):
        out[f'{name}_{k}'] = v.numpy()
    out[f'{name}_S'] = np.int64(S)
out['torus_rows'] = np.asarray(sorted(torn.rows.items()), dtype=np.float64)
out['so3_rows'] = np.asarray(sorted(so3n.rows.items()), dtype=np.float64)
np.savez_compressed(os.path.join(ROOT, 'tests/golden/ref_forward.npz'), **out)
print('wrote tests/golden/ref_forward.npz')
