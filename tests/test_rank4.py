"""SURVEY 8f-4 drivers (utils.sampling.sample_step / sampling_phore_with_fitscore / get_updates_from_0_to_n) against the outputs of the
reference's OWN functions (tests/golden/ref_rank4.npz, tools/make_rank4_golden.py: the unmodified reference sampling.py driving the
unmodified reference model with the shipped checkpoint).  Tolerance: coordinates after ONE step within 1e-4 A RMSD per sample and
perturbations rel-L2 <= 1e-4 (fp32 path, same Gaussian draws: both sides draw from torch's CPU generator in the same order); the
3-step fitscore-guided runs (dt = 1/3, ill-conditioned) within 5e-3 A, see there."""
import os
from functools import partial
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from tests.parity_util import have_checkpoint, real_state_dict, load_pairs, rel, SHIPPED_KW

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ARGS = SimpleNamespace(tr_sigma_min=0.1, tr_sigma_max=5.0, rot_sigma_min=0.1, rot_sigma_max=1.5, tor_sigma_min=0.0314,
                       tor_sigma_max=3.14, no_torsion=False, keep_update=False, random_samples=3)
needs_ckpt = pytest.mark.skipif(not have_checkpoint(), reason='shipped checkpoint not present (oracle/_ref/weights)')


def _start_graphs():
    gold = np.load(os.path.join(ROOT, 'tests/golden/ref_sampler.npz'))
    base = load_pairs('real', 1)[0]
    n = base['ligand'].pos.shape[0]
    out = []
    for k in range(3):
        g = base.clone()
        g['ligand'].pos = torch.from_numpy(gold['samp_start_pos'][k * n:(k + 1) * n]).clone()
        g['ligand'].norm = torch.from_numpy(gold['samp_start_norm'][k * n:(k + 1) * n]).clone()
        g.name = 'pair'
        g.original_center = torch.tensor([[1.0, -2.0, 0.5]])
        out.append(g)
    return out, n


def _rmsd(a, b, n):
    return [float(((a[k * n:(k + 1) * n] - b[k * n:(k + 1) * n]) ** 2).sum(1).mean().sqrt()) for k in range(a.shape[0] // n)]


def surrogate_fitscore(args, ligand_pos, name, mol, store_ranked_pose=True, phore_file=None):
    return [float(-np.linalg.norm(np.asarray(p).mean(0))) for p in ligand_pos]


def test_get_updates_from_0_to_n_equals_the_reference_function():
    """Host helper of the calibrated sampler (sampling.py:566-597): torsion update + Kabsch re-alignment + Kabsch onto the target."""
    from utils.sampling import get_updates_from_0_to_n
    z = np.load(os.path.join(ROOT, 'tests/golden/ref_rank4.npz'))
    start, n = _start_graphs()
    g_b = start[0].clone()
    g_b['ligand'].pos = torch.from_numpy(z['step_pos'][:n]).clone()
    t2, r1 = get_updates_from_0_to_n(start[0], g_b, z['upd0n_tor'])
    assert np.abs(t2.numpy() - z['upd0n_t']).max() <= 2e-5 and np.abs(np.asarray(r1) - z['upd0n_rot']).max() <= 2e-5


@pytest.mark.gpu
@needs_ckpt
def test_sample_step_and_fitscore_guided_sampling_match_the_reference_functions(built_lib):
    from models.score_model_phore import TensorProductScoreModel
    from utils.sampling import sample_step, sampling_phore_with_fitscore
    from utils.diffusion_utils import set_time_phore, get_t_schedule, t_to_sigma as tts, get_timestep_embedding
    from diffphore_b200.graph import collate
    z = np.load(os.path.join(ROOT, 'tests/golden/ref_rank4.npz'))
    dev = torch.device('cuda:0')
    t_to_sigma = partial(tts, args=ARGS)
    model = TensorProductScoreModel(t_to_sigma, dev, get_timestep_embedding('sinusoidal', 20, 10000), **SHIPPED_KW)
    model.load_state_dict(real_state_dict(), strict=True)
    model.eval()
    start, n = _start_graphs()
    # ---- sample_step: one Euler-Maruyama step at t = 0.6, draws from torch.manual_seed(99) like the reference run
    t = float(z['step_t'])
    batch = collate([g.clone() for g in start])
    set_time_phore(batch, t, t, t, 3, 'cpu')
    torch.manual_seed(99)
    dl, tor_p, tr_p, rot_p = sample_step(batch, model, ARGS, *t_to_sigma(t, t, t), delta_t=0.05)
    assert rel(tr_p, z['step_tr']) <= 1e-4 and rel(rot_p, z['step_rot']) <= 1e-4 and rel(tor_p, z['step_tor']) <= 1e-4
    pos = torch.cat([g['ligand'].pos for g in dl])
    assert max(_rmsd(pos, torch.from_numpy(z['step_pos']), n)) <= 1e-4
    assert float((torch.cat([g['ligand'].norm.reshape(n, -1) for g in dl]) - torch.from_numpy(z['step_norm'])).abs().max()) <= 1e-4
    # ---- fitscore-guided sampling: 3 candidates per graph and step, best one kept (surrogate score on both sides)
    steps = int(z['fit_steps'])
    sch = get_t_schedule(steps)
    torch.manual_seed(77)
    res, conf = sampling_phore_with_fitscore([g.clone() for g in start], model, steps, sch, sch, sch, dev, t_to_sigma, ARGS,
                                             batch_size=3, fitscore_fn=surrogate_fitscore)
    assert conf is None
    r = _rmsd(torch.cat([g['ligand'].pos for g in res]), torch.from_numpy(z['fit_pos']), n)
    # 3-step schedule (dt = 1/3): perturbations ~7x those of the 20-step loop the 1e-4 A bound is stated for, and this trajectory is
    # ill-conditioned - the CPU restatement of the same run differs from the reference's output by 2.4e-3 A in fp32 and 1.2e-3 A in
    # fp64 (one sample each; measured with oracle.sampler on the replayed draws).  Bound: 5e-3 A here (CUDA measured 4e-5 .. 1.2e-3),
    # 1e-4 A for the single step above.
    assert max(r) <= 5e-3, r
    # ... and the plain loop through the same function (random_samples = 0)
    args1 = SimpleNamespace(**{**vars(ARGS), 'random_samples': 0})
    torch.manual_seed(78)
    res1, _ = sampling_phore_with_fitscore([g.clone() for g in start], model, steps, sch, sch, sch, dev, t_to_sigma, args1, batch_size=3)
    r1 = _rmsd(torch.cat([g['ligand'].pos for g in res1]), torch.from_numpy(z['fit1_pos']), n)
    assert max(r1) <= 5e-3, r1
    # reference failure modes are kept: ode with torsions has no defined torsion noise there
    with pytest.raises(NotImplementedError):
        sampling_phore_with_fitscore([g.clone() for g in start], model, steps, sch, sch, sch, dev, t_to_sigma, args1, ode=True)


@pytest.mark.gpu
@needs_ckpt
def test_calibrated_sampler_step_carries_the_clean_pose_onto_the_model_step(built_lib):
    """NoiseTransformPhore.sample_from_infer (pdbbind_phore.py:286-359): x_(n+1) -> x_n by the model (sample_step on the GPU), then the
    (translation, rotation, torsion) update of the CLEAN pose x_0 that reproduces x_n, applied on the GPU.  The reference's own
    consistency measure (its debug `rmsd`: re-expressed pose vs the model step) must be small, and the training targets well formed."""
    from models.score_model_phore import TensorProductScoreModel
    from datasets.pdbbind_phore import NoiseTransformPhore
    from utils.diffusion_utils import set_time_phore, t_to_sigma as tts, get_timestep_embedding
    dev = torch.device('cuda:0')
    t_to_sigma = partial(tts, args=ARGS)
    model = TensorProductScoreModel(t_to_sigma, dev, get_timestep_embedding('sinusoidal', 20, 10000), **SHIPPED_KW)
    model.load_state_dict(real_state_dict(), strict=True)
    model.eval()
    start, n = _start_graphs()
    nt = NoiseTransformPhore(t_to_sigma, False, delta_t=0.05, rate_from_infer=0.1, model=model, args=ARGS)
    data = start[1].clone()
    # x_(n+1) is x_0 moved by known updates, as apply_noise builds it; here: the second start pose, reached from the first by a
    # rigid move (both are randomised copies of the same conformer up to torsions, so the test only checks self-consistency)
    t = 0.5
    set_time_phore(data, t, t, t, 1, 'cpu')
    torch.manual_seed(5)
    res, info = nt.sample_from_infer(data.clone(), data, t, *t_to_sigma(t, t, t), torsion_updates=None, debug=True)
    assert float(info['rmsd']) <= 5e-3, info['rmsd']                        # x_0 + (0 -> n updates) lands on the model's x_n
    assert res.tr_score.shape == (1, 3) and res.rot_score.shape == (1, 3) and torch.isfinite(res.tor_score).all()
    assert res.tor_score.shape[0] == int(res['ligand'].edge_mask.sum()) == res.tor_sigma_edge.shape[0]
