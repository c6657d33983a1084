"""Frozen oracle outputs (SURVEY §8c "golden vectors the new repo must create itself", items 1-3), committed as
tests/golden/oracle_frozen.npz so that an accidental edit of oracle/ (the checker of every GPU parity test) is caught on CPU.

    python tools/make_oracle_frozen.py            (no /root/reference needed; deterministic: seeded inputs, seeded weights)
* forward: tr / rot / tor of the fp32 oracle, the activations after every ligand convolution layer / final_conv / tor_bond_conv
  (act_*), and tr / rot / tor of the same forward in float64 (*_f64_*), for one seeded batch of every config shape - cfg1 (real-shaped example pair, P = 79),
  cfg2 (32 atoms / 8 points), cfg4 (64 / 12), cfg5 (128 / 16) - 2 samples each at t = 0.6, random-init weights (seed 0);
  for cfg1 additionally with the shipped checkpoint when oracle/_ref/weights holds it (skipped otherwise);
* trajectory: final ligand coordinates of a 20-step run with injected initial poses and noise (2 pairs of 14 atoms / 5 points x 2);
* tables: so3 / torus score norms at the 20 noise levels of the schedule (torus table seed 0).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffphore_b200.graph import collate                                   # noqa: E402
from oracle import sampler as osamp                                        # noqa: E402
from oracle.model import OracleScoreModel, default_config                  # noqa: E402
from oracle.tables import So3ScoreNorm, TorusScoreNorm                     # noqa: E402
from tests.parity_util import (random_state_dict, real_state_dict, have_checkpoint, load_pairs, make_draws,   # noqa: E402
                               oracle_initial_graphs)

SHAPES = {'cfg1': ('real', 1, 0, 0), 'cfg2': ('synthetic', 1, 32, 8), 'cfg4': ('synthetic', 1, 64, 12),
          'cfg5': ('synthetic', 1, 128, 16)}


LAYER_KEYS = ('lig_node_attr1', 'lig_node_attr2', 'lig_node_attr3', 'lig_node_attr4', 'final_conv.out', 'tor_bond_conv.out')


def forward_case(kind, n_pairs, n_atoms, n_phore, sd, so3n, torn, t=0.6, samples=2, seed=3, dtype=torch.float32, layers=False):
    graphs = load_pairs(kind, n_pairs, n_atoms, n_phore)
    init, _, n_rot = make_draws(graphs, samples, seed)
    dl = oracle_initial_graphs(graphs, samples, init, n_rot)
    batch = collate(dl)
    osamp.set_time(batch, t, len(dl))
    om = OracleScoreModel(sd, so3n, torn, dtype=dtype)
    om.trace = {} if layers else None
    with torch.no_grad():
        res = [o.double().numpy() if dtype == torch.float64 else o.float().numpy() for o in om(batch)]
    if layers:                                          # per-convolution-layer activations (SURVEY 8c golden list, item 1)
        res += [om.trace[k].float().numpy() for k in LAYER_KEYS]
    return res


def trajectory_case(sd, so3n, torn, steps=20, samples=2, seed=11):
    graphs = load_pairs('synthetic', 2, 14, 5)
    init, noise, n_rot = make_draws(graphs, samples, seed, steps=steps)
    dl = oracle_initial_graphs(graphs, samples, init, n_rot)
    res = osamp.sampling(dl, OracleScoreModel(sd, so3n, torn), steps, default_config(), collate, batch_size=samples, noise=noise)
    return torch.cat([g['ligand'].pos for g in res]).numpy()


def build():
    torch.manual_seed(0)
    so3n, torn = So3ScoreNorm(), TorusScoreNorm(seed=0)
    out = {}
    sd = random_state_dict(0)
    for name, (kind, n_pairs, n_atoms, n_phore) in SHAPES.items():
        keys = ('tr', 'rot', 'tor') + tuple('act_' + k.replace('.', '_') for k in LAYER_KEYS)
        for key, val in zip(keys, forward_case(kind, n_pairs, n_atoms, n_phore, sd, so3n, torn, layers=True)):
            out[f'{name}_{key}'] = val
        for key, val in zip(('tr', 'rot', 'tor'), forward_case(kind, n_pairs, n_atoms, n_phore, sd, so3n, torn, dtype=torch.float64)):
            out[f'{name}_f64_{key}'] = val              # the same forward evaluated in float64 (noise floor of the fp32 paths)
    if have_checkpoint():
        for key, val in zip(('tr', 'rot', 'tor'), forward_case('real', 1, 0, 0, real_state_dict(), so3n, torn)):
            out[f'cfg1_shipped_{key}'] = val
    out['traj_pos'] = trajectory_case(sd, so3n, torn)
    sched = osamp.get_t_schedule(20)
    out['so3_norm'] = np.asarray([so3n(np.asarray([0.1 ** (1 - t) * 1.5 ** t], dtype=np.float32))[0] for t in sched])
    out['torus_norm'] = np.asarray([torn(np.asarray([0.0314 ** (1 - t) * 3.14 ** t], dtype=np.float32))[0] for t in sched])
    return out


if __name__ == '__main__':
    res = build()
    path = os.path.join(ROOT, 'tests/golden/oracle_frozen.npz')
    np.savez_compressed(path, **res)
    print('wrote', path, os.path.getsize(path), 'bytes:', {k: v.shape for k, v in res.items()})
