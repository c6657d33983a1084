"""CPU tests that pin the oracle (oracle/) — the checker of every GPU parity claim.

Pins available: the reference ships NO tests or golden vectors.  The oracle is pinned by
 (1) tests/golden/ref_forward.npz: outputs of the UNMODIFIED reference model class executed over dependency shims
     (tools/make_golden.py) with the shipped checkpoint;
 (2) e3nn's Wigner-3j buffers serialized in the checkpoint + SO(3)-equivariance of the whole forward;
 (3) the checkpoint's BatchNorm running statistics, which the restated forward reproduces on real-shaped inputs.
"""
import math
import os

import numpy as np
import pytest
import torch

from tests.parity_util import (ROOT, have_checkpoint, real_state_dict, random_state_dict, load_pairs, make_draws,
                               oracle_initial_graphs, rel)
from diffphore_b200.graph import collate, graph_from_arrays
from oracle import e3nn_lite as e3
from oracle import sampler as osamp
from oracle.model import OracleScoreModel, radius_graph, radius_pairs, scatter
from oracle.tables import So3ScoreNorm, TorusScoreNorm

needs_ckpt = pytest.mark.skipif(not have_checkpoint(), reason='shipped checkpoint not present (oracle/_ref/weights)')


def test_w3j_norm_and_known_values():
    for key in [(1, 1, 1), (1, 2, 1), (2, 2, 1), (2, 2, 0), (0, 2, 2)]:
        c = e3.w3j(*key, dtype=torch.float64)
        assert abs(float(c.norm()) - 1) < 1e-6
    eps = e3.w3j(1, 1, 1, torch.float64)
    assert abs(float(eps[0, 1, 2]) - 1 / math.sqrt(6)) < 1e-7 and abs(float(eps[0, 2, 1]) + 1 / math.sqrt(6)) < 1e-7


@needs_ckpt
def test_w3j_fixture_equals_checkpoint_buffers():
    sd = real_state_dict()
    z = np.load(os.path.join(ROOT, 'oracle', 'w3j.npz'))
    n = 0
    for k, v in sd.items():
        if '_w3j_' in k:
            assert np.array_equal(z['w3j_' + k.split('_w3j_')[1]], v.numpy()), k
            n += 1
    assert n == 47          # 18 conv layers x 2 + final_conv x 2 + final_tp_tor x 9


def test_sh_consistent_with_w3j_121():
    x = torch.nn.functional.normalize(torch.randn(50, 3, dtype=torch.float64), dim=-1)
    sh = e3.spherical_harmonics(x)
    out = torch.einsum('ijk,zi,zj->zk', e3.w3j(1, 2, 1, torch.float64), x, sh[:, 4:9] / math.sqrt(5))
    assert torch.allclose(out, out[0, 0] / x[0, 0] * x, atol=1e-6)
    assert abs(float(out[0, 0] / x[0, 0]) - 0.365148) < 1e-5
    assert torch.allclose((sh[:, 1:4] ** 2).sum(-1), torch.full((50,), 3.0, dtype=torch.float64))       # 'component'
    assert torch.allclose((sh[:, 4:9] ** 2).sum(-1), torch.full((50,), 5.0, dtype=torch.float64))
    assert torch.equal(e3.spherical_harmonics(torch.zeros(1, 3))[0], torch.tensor([1.] + [0.] * 8))     # zero vector


def test_fctp_against_naive_einsum_and_instruction_tables():
    seq = [e3.parse_irreps(s) for s in ('20x0e', '20x0e + 10x1o', '20x0e + 10x1o + 10x1e', '20x0e + 10x1o + 10x1e + 20x0o')]
    sh = e3.sh_irreps(2)
    for (a, b, numel) in [(seq[0], seq[1], 600), (seq[1], seq[2], 1100), (seq[2], seq[3], 1600), (seq[3], seq[3], 2200),
                          (seq[3], e3.parse_irreps('2x1o + 2x1e'), 200)]:
        instrs, n = e3.fctp_instructions(a, sh, b)
        assert n == numel
        E = 5
        x1, x2, w = torch.randn(E, e3.irreps_dim(a)), torch.randn(E, 9), torch.randn(E, n)
        out = e3.fctp_apply(a, sh, b, instrs, x1, x2, w)
        ref = torch.zeros(E, e3.irreps_dim(b))
        s1, s2, so = e3.irreps_slices(a), e3.irreps_slices(sh), e3.irreps_slices(b)
        for ins in instrs:
            (m1, l1, _), (_, l2, _), (mo, lo, _) = a[ins.i1], sh[ins.i2], b[ins.io]
            X = x1[:, s1[ins.i1][0]:s1[ins.i1][1]].reshape(E, m1, 2 * l1 + 1)
            Y = x2[:, s2[ins.i2][0]:s2[ins.i2][1]]
            W = w[:, ins.w_off:ins.w_off + ins.w_len].reshape(E, m1, 1, mo)
            r = torch.einsum('zuvw,ijk,zui,zvj->zwk', W, e3.w3j(l1, l2, lo), X, Y[:, None, :]) * ins.pw
            ref[:, so[ins.io][0]:so[ins.io][1]] += r.reshape(E, -1)
        assert rel(out, ref) < 1e-5
    sh45, _ = e3.full_tp_irreps_out(sh, [(1, 2, 1)])
    assert [(l, p) for _, l, p in sh45] == [(0, 1), (1, -1), (1, 1), (2, -1), (2, 1), (2, 1), (3, -1), (3, 1), (4, 1)]
    assert e3.fctp_instructions(seq[3], sh45, e3.parse_irreps('20x0o + 20x0e'))[1] == 1600


def test_batchnorm_scatter_radius_semantics():
    irreps = e3.parse_irreps('2x0e + 1x1o + 1x0o')
    x = torch.randn(4, 2 + 3 + 1)
    w, b, rm, rv = torch.rand(4) + 0.5, torch.randn(2), torch.randn(2), torch.rand(4) + 0.5
    y = e3.batchnorm_eval(x, irreps, w, b, rm, rv)
    s = w / torch.sqrt(rv + 1e-5)
    assert torch.allclose(y[:, :2], (x[:, :2] - rm) * s[:2] + b, atol=1e-6)
    assert torch.allclose(y[:, 2:5], x[:, 2:5] * s[2], atol=1e-6) and torch.allclose(y[:, 5], x[:, 5] * s[3], atol=1e-6)
    out = scatter(torch.tensor([[1.], [3.], [5.]]), torch.tensor([0, 0, 2]), 4, 'mean')
    assert out.reshape(-1).tolist() == [2.0, 0.0, 5.0, 0.0]
    # radius graph: strict <, per-graph, self removed, cap keeps the lowest indices (incl. self in the count of cap+1)
    pos = torch.tensor([[0., 0, 0], [5., 0, 0], [4.99, 0, 0], [0., 0, 0.1]])
    e = radius_graph(pos, 5.0, torch.zeros(4, dtype=torch.long))
    pairs = set(map(tuple, e.T.tolist()))
    assert (1, 0) not in pairs and (2, 0) in pairs and (0, 0) not in pairs
    pos = torch.rand(60, 3) * 0.5
    e = radius_graph(pos, 5.0, torch.zeros(60, dtype=torch.long), max_num_neighbors=32)
    centre_counts = torch.bincount(e[1], minlength=60)
    assert int(centre_counts.max()) <= 33 and int(centre_counts[50]) == 33 and int(centre_counts[0]) == 32
    assert int(e[0][e[1] == 50].max()) == 32
    e2 = radius_pairs(pos, pos[:3] , 5.0, torch.zeros(60, dtype=torch.long), torch.zeros(3, dtype=torch.long), 32)
    assert e2.shape[1] == 96 and int(e2[1].max()) == 31


def _noised(graphs, S, t, seed):
    init, _, n_rot = make_draws(graphs, S, seed)
    init['tr'] = init['tr'] / 5.0 * (0.1 ** (1 - t) * 5.0 ** t)
    return oracle_initial_graphs(graphs, S, init, n_rot)


def test_full_forward_so3_equivariance_without_clamp_and_translation_invariance():
    """With the H10 clamp disabled the restated network is exactly SE(3)-equivariant: pins SH, every CG path, the
    gather/scatter directions and the FullTensorProduct ordering between 1o and 1e.  (With the clamp — the shipped
    behaviour — it is not)"""
    from scipy.spatial.transform import Rotation
    sd = random_state_dict(3)
    graphs = load_pairs('synthetic', 2, 10, 5)
    dl = _noised(graphs, 1, 0.5, 1)
    tabs = (So3ScoreNorm(), TorusScoreNorm())
    m = OracleScoreModel(sd, *tabs, config=dict(no_clamp=True), dtype=torch.float64)
    b = collate([g.clone() for g in dl]); osamp.set_time(b, 0.5, len(dl))
    tr, rot, tor = m(b)
    R = torch.from_numpy(Rotation.random(random_state=5).as_matrix())
    shift = torch.tensor([[1.5, -2.0, 0.7]], dtype=torch.float64)
    dl2 = [g.clone() for g in dl]
    for g in dl2:
        n = g['ligand'].pos.shape[0]
        g['ligand'].pos = (g['ligand'].pos.double() @ R.T + shift).float()
        g['phore'].pos = (g['phore'].pos.double() @ R.T + shift).float()
        g['ligand'].norm = (g['ligand'].norm.double().reshape(n, 11, 3) @ R.T).reshape(n, 33).float()
        g['phore'].norm = (g['phore'].norm.double() @ R.T).float()
    b2 = collate(dl2); osamp.set_time(b2, 0.5, len(dl2))
    tr2, rot2, tor2 = m(b2)
    assert rel(tr2, tr @ R.T) < 2e-5 and rel(rot2, rot @ R.T) < 2e-5 and rel(tor2, tor) < 2e-5   # fp32 inputs re-rounded
    m_clamp = OracleScoreModel(sd, *tabs, dtype=torch.float64)
    b3 = collate([g.clone() for g in dl2]); osamp.set_time(b3, 0.5, len(dl2))
    b4 = collate([g.clone() for g in dl]); osamp.set_time(b4, 0.5, len(dl))
    assert rel(m_clamp(b3)[2], m_clamp(b4)[2]) > 1e-4          # H10: the shipped clamp breaks equivariance


def test_batch_composition_invariance():
    sd = random_state_dict(1)
    dl = _noised(load_pairs('synthetic', 3, 12, 5), 1, 0.4, 2)
    m = OracleScoreModel(sd, So3ScoreNorm(), TorusScoreNorm())
    b = collate([g.clone() for g in dl]); osamp.set_time(b, 0.4, 3)
    tr, rot, tor = m(b)
    b1 = collate([dl[1].clone()]); osamp.set_time(b1, 0.4, 1)
    tr1, rot1, tor1 = m(b1)
    assert rel(tr1[0], tr[1]) < 1e-5 and rel(rot1[0], rot[1]) < 1e-5


@needs_ckpt
def test_oracle_matches_unmodified_reference_model_golden():
    """tests/golden/ref_forward.npz was produced by /root/reference/src/models/score_model_phore.py itself."""
    z = np.load(os.path.join(ROOT, 'tests', 'golden', 'ref_forward.npz'))
    sd = real_state_dict()
    a = np.load(os.path.join(ROOT, 'tests', 'golden', 'real_pairs.npz'))
    from diffphore_b200.synthetic import make_pairs
    cases = {'real': [graph_from_arrays(a, f'p{k}_') for k in (11, 0, 5)], 'syn': make_pairs(2, 32, 8)}
    so3n, torn = So3ScoreNorm(), TorusScoreNorm(seed=0)
    for name, graphs in cases.items():
        S, t = int(z[f'{name}_S']), float(z[f'{name}_t'])
        dl = [g.clone() for g in graphs for _ in range(S)]
        o = 0
        for g in dl:
            n = g['ligand'].pos.shape[0]
            g['ligand'].pos = torch.from_numpy(z[f'{name}_pos'][o:o + n].copy())
            g['ligand'].norm = torch.from_numpy(z[f'{name}_norm'][o:o + n].copy())
            o += n
        b = collate(dl); osamp.set_time(b, t, len(dl))
        out = OracleScoreModel(sd, so3n, torn)(b)
        tol = 1e-5 if name == 'real' else 2e-4          # 'syn' drives the shipped weights far out of distribution
        for k, v in zip(('tr', 'rot', 'tor'), out):
            assert rel(v, z[f'{name}_{k}']) < tol, (name, k, rel(v, z[f'{name}_{k}']))


@needs_ckpt
def test_forward_reproduces_checkpoint_batchnorm_statistics():
    """Pre-BN activations of the restated forward on the 18 real-shaped example pairs have the second moments the
    shipped checkpoint recorded during training: pins instruction tables, weight offsets, path
    weights, SH normalisation, gather/scatter direction and the embeddings far better than chance."""
    from scipy.spatial.transform import Rotation as R
    sd = real_state_dict()
    pairs = load_pairs('real', 18)
    m = OracleScoreModel(sd, So3ScoreNorm(), TorusScoreNorm())
    rng = np.random.RandomState(0)
    acc = {}
    for t in (0.1, 0.4, 0.7, 1.0):
        gl = [p.clone() for p in pairs]
        osamp.randomize_position(gl, False, False, 5.0,
                                 [rng.normal(size=int(g['ligand'].edge_mask.sum())) * (0.0314 ** (1 - t) * 3.14 ** t) for g in gl],
                                 [R.from_rotvec(R.random(random_state=rng).as_rotvec() * min(1.0, (0.1 ** (1 - t) * 1.5 ** t) / 1.5)).as_matrix() for g in gl],
                                 [rng.normal(size=(1, 3)) * (0.1 ** (1 - t) * 5 ** t) for g in gl])
        b = collate(gl); osamp.set_time(b, t, len(gl))
        m.trace = {}
        m(b)
        for k, v in m.trace.items():
            if k.endswith('.pre_bn'):
                acc.setdefault(k[:-7], []).append(v)
    ratios = []
    for pre, vs in acc.items():
        if not pre.startswith('encoder.'):
            continue
        v = torch.cat(vs, 0)
        rm, rv = sd[pre + '.batch_norm.running_mean'], sd[pre + '.batch_norm.running_var']
        ratios.append(float((((v[:, :20] - rm) ** 2).mean(0) / rv[:20]).median()))
        corr = np.corrcoef(v[:, :20].mean(0).numpy(), rm.numpy())[0, 1]
        assert corr > 0.3, (pre, corr)
    assert len(ratios) == 21 and 0.3 < min(ratios) and max(ratios) < 4.0, ratios


def test_tables_product_equals_oracle():
    from diffphore_b200 import tables as pt
    s = np.asarray([0.1, 0.37, 1.5], dtype=np.float32)
    assert np.allclose(pt.So3ScoreNorm()(s), So3ScoreNorm()(s), rtol=1e-10)
    s = np.asarray([0.0314, 0.4, 3.14], dtype=np.float32)
    assert np.allclose(pt.TorusScoreNorm(seed=0)(s), TorusScoreNorm(seed=0)(s), rtol=1e-10)
    assert So3ScoreNorm()(np.float32([0.001]))[0] == So3ScoreNorm()(np.float32([0.01]))[0]        # clipped index


def test_score_norm_tables_equal_the_reference_functions():
    """tests/golden/tables_ref.npz (tools/make_tables_golden.py): the table-building functions of the reference's so3.py /
    torus.py, extracted unmodified and evaluated for the rows of the 20-step schedule.  so3: _exp_score_norms[idx]; torus: the
    deterministic score_ = grad / p rows (the Monte-Carlo average on top is unseeded in the reference, H1: checked against the
    exact second moment instead)."""
    from oracle import tables as ot
    from diffphore_b200 import tables as pt
    gold = np.load(os.path.join(ROOT, 'tests', 'golden', 'tables_ref.npz'))
    assert np.array_equal(ot.so3_eps_index(gold['so3_eps']), gold['so3_idx'])
    assert np.array_equal(pt.So3ScoreNorm.index(gold['so3_eps']), gold['so3_idx'])
    assert np.allclose(So3ScoreNorm()(gold['so3_eps']), gold['so3_exp_score_norm'], rtol=1e-12)
    assert np.allclose(pt.So3ScoreNorm()(gold['so3_eps']), gold['so3_exp_score_norm'], rtol=1e-9)
    assert np.array_equal(ot.torus_sigma_index(gold['torus_sigma']), gold['torus_idx'])
    assert np.array_equal(pt.TorusScoreNorm.index(gold['torus_sigma']), gold['torus_idx'])
    prod = pt.TorusScoreNorm(seed=0)
    for i, ref_row in zip(gold['torus_idx'], gold['torus_score_rows']):
        assert np.allclose(ot.torus_score_row(int(i))[::25], ref_row, rtol=1e-12, equal_nan=True)
        assert np.allclose(prod.score_row(int(i))[::25], ref_row, rtol=1e-9, equal_nan=True)
    # the seeded Monte-Carlo rows agree with the exact E[score^2] of the nearest-grid score table under the wrapped normal
    for i in gold['torus_idx'][[0, 9, 19]]:
        sig = (10 ** np.linspace(np.log10(ot.SIGMA_MIN), np.log10(ot.SIGMA_MAX), ot.SIGMA_N + 1) * np.pi)[int(i)]
        rng = np.random.RandomState(12345)
        s = sig * rng.randn(400000)
        s = (s + np.pi) % (2 * np.pi) - np.pi
        xi = (np.log(np.abs(s) / np.pi) - np.log(ot.X_MIN)) / (0 - np.log(ot.X_MIN)) * ot.TX_N
        exact = float((ot.torus_score_row(int(i))[np.round(np.clip(xi, 0, ot.TX_N)).astype(int)] ** 2).mean())
        assert abs(ot.torus_score_norm_row(int(i), 0) / exact - 1) < 0.05, (i, exact)


def _sampler_gold_graphs():
    """The start graphs of tools/make_sampler_golden.py: 2 synthetic pairs + the first real-shaped pair, 2 copies each."""
    graphs = load_pairs('synthetic', 2, 14, 5) + load_pairs('real', 1)
    return [g.clone() for g in graphs for _ in range(2)]


def test_sampler_oracle_equals_the_reference_sampler_code():
    """tests/golden/ref_sampler.npz (tools/make_sampler_golden.py): outputs of the UNMODIFIED reference utils/sampling.py,
    diffusion_utils.py, torsion.py, geometry.py run over shims on seeded inputs.  Pins randomize_position (with its draws replayed),
    modify_conformer with / without torsion updates (Kabsch, H4 norm-point quirk, skipped bonds), the t schedule, t_to_sigma and
    the sinusoidal embedding of the oracle AND of the product's host code."""
    from diffphore_b200 import sampler as psamp
    from diffphore_b200.engine import ModelWeights
    from oracle.model import sinusoidal_embedding
    gold = np.load(os.path.join(ROOT, 'tests', 'golden', 'ref_sampler.npz'))
    sched = osamp.get_t_schedule(20)
    assert np.array_equal(sched, gold['t_schedule']) and np.array_equal(psamp.get_t_schedule(20), gold['t_schedule'])
    w = ModelWeights(random_state_dict(0), 'cpu')
    for t, ref_sig, ref_emb in zip(sched, gold['t_to_sigma'], gold['sigma_emb']):
        assert np.allclose(w.t_to_sigma(t), ref_sig, rtol=2e-7)          # (model side: fp32, like the tensors of set_time_phore)
        assert np.allclose(sinusoidal_embedding(10000 * torch.tensor([float(t)]), 20)[0].numpy(), ref_emb, atol=1e-6)
        assert np.allclose(w.step_consts(float(t), So3ScoreNorm(), TorusScoreNorm(seed=0))[0:20].numpy(), ref_emb, atol=1e-6)
    # the mirrored host helpers of src/utils/diffusion_utils.py
    from types import SimpleNamespace
    from utils import diffusion_utils as mdu
    margs = SimpleNamespace(tr_sigma_min=0.1, tr_sigma_max=5.0, rot_sigma_min=0.1, rot_sigma_max=1.5, tor_sigma_min=0.0314,
                            tor_sigma_max=3.14)
    emb = mdu.get_timestep_embedding('sinusoidal', 20, 10000)
    assert np.array_equal(mdu.get_t_schedule(20), gold['t_schedule'])
    assert np.array_equal(np.asarray([mdu.t_to_sigma(t, t, t, margs) for t in sched]), gold['t_to_sigma'])
    assert np.array_equal(torch.stack([emb(torch.tensor([float(t)])) for t in sched]).squeeze(1).numpy(), gold['sigma_emb'])
    # randomize_position with the reference's draws
    dl = _sampler_gold_graphs()
    n_rot = [int(g['ligand'].edge_mask.sum()) for g in dl]
    offs = np.concatenate([[0], np.cumsum(n_rot)])
    osamp.randomize_position(dl, False, False, 5.0, [gold['rand_tor'][offs[i]:offs[i + 1]] for i in range(len(dl))],
                             list(gold['rand_rot']), list(gold['rand_tr']))
    assert float((torch.cat([g['ligand'].pos for g in dl]) - torch.from_numpy(gold['rand_pos'])).abs().max()) <= 2e-5
    assert float((torch.cat([g['ligand'].norm for g in dl]) - torch.from_numpy(gold['rand_norm'])).abs().max()) <= 2e-5
    # modify_conformer, with torsions (some skipped) and rigid only, starting from the reference's randomised poses
    ptr = np.concatenate([[0], np.cumsum([g['ligand'].pos.shape[0] for g in dl])])
    for i, g in enumerate(dl):
        g['ligand'].pos = torch.from_numpy(gold['rand_pos'][ptr[i]:ptr[i + 1]]).clone()
        g['ligand'].norm = torch.from_numpy(gold['rand_norm'][ptr[i]:ptr[i + 1]]).clone()
        trp, rotp = torch.from_numpy(gold['upd_tr'][i:i + 1]), torch.from_numpy(gold['upd_rot'][i])
        a = osamp.modify_conformer(g.clone(), trp, rotp, gold['upd_tor'][offs[i]:offs[i + 1]])
        b = osamp.modify_conformer(g.clone(), trp, rotp, None)
        for got, key in ((a['ligand'].pos, 'upd_pos'), (a['ligand'].norm, 'upd_norm'), (b['ligand'].pos, 'rigid_pos'),
                         (b['ligand'].norm, 'rigid_norm')):
            assert float((got - torch.from_numpy(gold[key][ptr[i]:ptr[i + 1]])).abs().max()) <= 2e-5, (i, key)


@needs_ckpt
def test_oracle_sampling_loop_equals_the_reference_sampling_phore():
    """sampling_phore of the reference (sampling.py:174-280) driving the UNMODIFIED reference model with the shipped checkpoint for
    6 steps (no_random and ode variants, tests/golden/ref_sampler.npz) against oracle.sampler.sampling + OracleScoreModel."""
    from oracle.model import default_config
    gold = np.load(os.path.join(ROOT, 'tests', 'golden', 'ref_sampler.npz'))
    base = load_pairs('real', 1)[0]
    n = base['ligand'].pos.shape[0]
    start = []
    for k in range(3):
        g = base.clone()
        g['ligand'].pos = torch.from_numpy(gold['samp_start_pos'][k * n:(k + 1) * n]).clone()
        g['ligand'].norm = torch.from_numpy(gold['samp_start_norm'][k * n:(k + 1) * n]).clone()
        start.append(g)
    om = OracleScoreModel(real_state_dict(), So3ScoreNorm(), TorusScoreNorm(seed=0))
    steps = int(gold['samp_steps'])
    for tag, kw in (('norandom', {}), ('ode', dict(ode=True))):
        res = osamp.sampling([g.clone() for g in start], om, steps, default_config(), collate, batch_size=3, noise=None, **kw)
        pos = torch.cat([g['ligand'].pos for g in res])
        ref = torch.from_numpy(gold[f'samp_{tag}_pos'])
        rmsd = [float(((pos[k * n:(k + 1) * n] - ref[k * n:(k + 1) * n]) ** 2).sum(1).mean().sqrt()) for k in range(3)]
        assert max(rmsd) <= 1e-4, (tag, rmsd)
        # samples 0 and 2 started from the same pose: the reference's own batched run separates them by a few 1e-6 A
        assert float((ref[:n] - ref[2 * n:]).abs().max()) <= 2e-5
        print(tag, 'RMSD oracle vs reference sampling_phore:', rmsd)


def test_conformer_update_oracle_properties():
    """modify_conformer keeps bond lengths and (by the Kabsch re-alignment) the centroid displacement equal to tr."""
    g = _noised(load_pairs('synthetic', 1, 14, 5), 1, 0.5, 3)[0]
    lig = g['ligand']
    ei = g['ligand', 'ligand'].edge_index
    d0 = (lig.pos[ei[0]] - lig.pos[ei[1]]).norm(dim=1)
    c0 = lig.pos.mean(0)
    n_rot = int(lig.edge_mask.sum())
    tr = torch.tensor([[0.3, -0.2, 0.5]])
    g2 = osamp.modify_conformer(g.clone(), tr, torch.tensor([0.2, 0.1, -0.3]), np.linspace(-0.5, 0.5, n_rot).astype(np.float32))
    d1 = (g2['ligand'].pos[ei[0]] - g2['ligand'].pos[ei[1]]).norm(dim=1)
    assert torch.allclose(d0, d1, atol=1e-5)
    assert torch.allclose(g2['ligand'].pos.mean(0) - c0, tr[0], atol=1e-5)
    assert g2['ligand'].norm.shape == (14, 33)


def test_oracle_sampling_loop_runs_and_is_deterministic():
    """sampling_phore restatement end to end (2 steps, per-step re-collation through collate/to_data_list)."""
    from oracle.model import default_config
    sd = random_state_dict(0)
    graphs = load_pairs('synthetic', 2, 8, 4)
    init, noise, n_rot = make_draws(graphs, 2, 1, steps=2)
    tabs = (So3ScoreNorm(), TorusScoreNorm())
    outs = []
    for _ in range(2):
        dl = oracle_initial_graphs(graphs, 2, init, n_rot)
        res = osamp.sampling(dl, OracleScoreModel(sd, *tabs), 2, default_config(), collate, batch_size=2, noise=noise)
        outs.append(torch.cat([g['ligand'].pos for g in res]))
        assert len(res) == 4 and res[0]['ligand'].edge_mask.shape[0] == graphs[0]['ligand', 'ligand'].edge_index.shape[1]
    assert torch.equal(outs[0], outs[1]) and torch.isfinite(outs[0]).all()
    res0 = osamp.sampling(oracle_initial_graphs(graphs, 2, init, n_rot), OracleScoreModel(sd, *tabs), 2, default_config(),
                          collate, batch_size=3, noise=None)
    assert not torch.equal(torch.cat([g['ligand'].pos for g in res0]), outs[0])       # no_random differs from noisy run


def test_oracle_reproduces_its_frozen_outputs():
    """tests/golden/oracle_frozen.npz (tools/make_oracle_frozen.py): forward outputs for one seeded batch of every config shape
    (cfg1 real-shaped P = 79, cfg2 32/8, cfg4 64/12, cfg5 128/16), a 20-step trajectory with injected draws, and the so3 / torus
    score norms of the 20-step schedule.  Guards the checker itself: an edit of oracle/ that moves any of these is caught here."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('make_oracle_frozen', os.path.join(ROOT, 'tools', 'make_oracle_frozen.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    gold = np.load(os.path.join(ROOT, 'tests', 'golden', 'oracle_frozen.npz'))
    now = mod.build()
    for k in gold.files:
        if k not in now:                                   # cfg1_shipped_* without the checkpoint
            assert k.startswith('cfg1_shipped_') and not have_checkpoint()
            continue
        if k == 'traj_pos':
            assert float(np.abs(now[k] - gold[k]).max()) <= 1e-3, k          # 20 chained steps (fp32 summation order may differ per host)
        else:
            assert rel(now[k], gold[k]) <= 1e-5, (k, rel(now[k], gold[k]))


def test_so3_score_vec_and_torus_score_equal_the_reference_functions():
    """utils.so3.score_vec (so3.py:84-89) and utils.torus.score (torus.py:46-55) - the training targets of the calibrated sampler
     - against the reference's own functions evaluated over its own table-building code (tests/golden/tables_ref.npz)."""
    from utils import so3, torus
    z = np.load(os.path.join(ROOT, 'tests/golden/tables_ref.npz'))
    got = np.stack([so3.score_vec(float(e), v) for e, v in zip(z['so3_vec_eps'], z['so3_vec'])])
    assert np.abs(got - z['so3_score_vec']).max() <= 1e-9 * max(1.0, np.abs(z['so3_score_vec']).max())
    assert np.allclose(torus.score(z['torus_x'], z['torus_x_sigma']), z['torus_score'], rtol=1e-12, atol=0, equal_nan=True)
