"""Data formats either side of the denoising path: `.phore` ingestion, rotatable-bond masks, AncPhore
`.score` parsing and the pose-output SD writer, pinned against tests/golden/ingest.npz — outputs of the reference's own functions
and its shipped example run (tools/make_ingest_golden.py).  CPU only."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'src'))
import _bootstrap  # noqa: E402,F401
from datasets import process_mols as pm              # noqa: E402
from datasets import process_pharmacophore as pp     # noqa: E402
from diffphore_b200.graph import HeteroGraph         # noqa: E402

ANCPHORE = os.path.join(ROOT, 'oracle', '_ref', 'programs', 'AncPhore')


@pytest.fixture(scope='module')
def gold():
    return np.load(os.path.join(ROOT, 'tests/golden/ingest.npz'))


def test_phore_parser_and_graph_equal_the_reference(gold, tmp_path):
    f = tmp_path / 'x.phore'
    f.write_text(str(gold['phore_text']))
    phores = pp.parse_phore(str(f))
    assert len(phores) == 1 and phores[0].id == str(gold['phore_id'])
    g = pp.get_phore_graph(phores[0], HeteroGraph(), consider_ex=True, neighbor_cutoff=5.0, ex_connected=True)
    assert torch.equal(g['phore'].x, torch.from_numpy(gold['phore_x']))
    assert torch.equal(g['phore'].pos, torch.from_numpy(gold['phore_pos']))
    assert np.allclose(g['phore'].norm.numpy(), gold['phore_norm'], atol=1e-7)
    assert np.array_equal(g['phore', 'phore_contact', 'phore'].edge_index.numpy(), gold['phore_edge_index'])
    # skip_ex drops the exclusion spheres only
    assert len(pp.parse_phore(str(f), skip_ex=True)[0].exclusion_volumes) == 0
    with pytest.raises(FileNotFoundError):
        pp.parse_phore(str(tmp_path / 'missing.phore'))
    bad = tmp_path / 'bad.phore'
    bad.write_text('id\nHA\t1.0\t2.0\n$$$$\n')
    with pytest.raises(SyntaxError):
        pp.parse_phore(str(bad))


def test_transformation_mask_equals_the_reference_on_all_example_ligands(gold, tmp_path):
    n_rot = 0
    for nm in gold['lig_names']:
        f = tmp_path / f'{nm}.sdf'
        f.write_text(str(gold[f'lig_text_{nm}']))
        g = pm.ligand_graph_from_sdf(str(f), HeteroGraph())
        me, mr = g['ligand'].edge_mask.numpy(), g['ligand'].mask_rotate
        assert np.array_equal(me, gold[f'mask_edges_{nm}']), nm
        assert np.array_equal(mr, gold[f'mask_rotate_{nm}']), nm
        ei = g['ligand', 'lig_bond', 'ligand'].edge_index.numpy()
        assert np.array_equal(ei[:, 0::2], ei[::-1, 1::2])           # both directions consecutive (get_lig_graph)
        for (u, v), row in zip(ei.T[me], mr):                         # the asserts of torsion.py:93-94
            assert not row[u] and row[v]
        n_rot += len(mr)
    assert n_rot > 30


def test_score_file_parser_equals_the_reference(gold, tmp_path, capsys):
    f = tmp_path / 'x.score'
    f.write_text(str(gold['score_text']))
    for fit in range(1, 7):
        assert np.array_equal(np.asarray(pp.parse_score_file(str(f), fitness=fit)), gold[f'score_fitness_{fit}'])
    assert np.array_equal(np.asarray(pp.parse_score_file(str(f), return_all=True)), gold['score_all'])
    assert pp.parse_score_file(str(tmp_path / 'missing.score')) is None          # prints, returns None (reference :921-923)
    g = tmp_path / 'garbage.score'
    g.write_text('a\tb\n')
    assert pp.parse_score_file(str(g)) is None
    assert '[E] Failed to parse' in capsys.readouterr().out


def _read_records(path):
    L = open(path).read().split('\n')
    recs, i = [], 0
    while i + 3 < len(L):
        na, nb = int(L[i + 3][0:3]), int(L[i + 3][3:6])
        xyz = np.asarray([[float(l[0:10]), float(l[10:20]), float(l[20:30])] for l in L[i + 4:i + 4 + na]])
        elem = [l[31:34].strip() for l in L[i + 4:i + 4 + na]]
        bonds = [(int(l[0:3]), int(l[3:6]), int(l[6:9])) for l in L[i + 4 + na:i + 4 + na + nb]]
        j = i
        while L[j] != '$$$$':
            j += 1
        recs.append(dict(name=L[i], xyz=xyz, elem=elem, bonds=bonds, tail=L[i + 4 + na + nb:j]))
        i = j + 1
    return recs


def _template(gold, tmp_path):
    f = tmp_path / 'STK936575.sdf'
    f.write_text(str(gold['lig_text_STK936575']))
    return pm.ligand_graph_from_sdf(str(f), HeteroGraph())


def test_pose_writer_record_layout(gold, tmp_path):
    g = _template(gold, tmp_path)
    poses = gold['kat_poses']
    out = tmp_path / 'docked.sdf'
    pm.write_mol_with_multi_coords(g.sdf_template, poses, str(out), 'sQC_Substrate__STK936575')
    recs = _read_records(str(out))
    assert [r['name'] for r in recs] == list(gold['kat_record_names'])           # f"{name}_{marker}_{idx}" with marker ''
    for r, p in zip(recs, poses):
        assert np.array_equal(r['xyz'], np.round(p, 4)) and 'H' not in r['elem'] and len(r['bonds']) == 22
    ranked = tmp_path / 'ranked.sdf'
    fs = np.linspace(1, 0, len(poses))
    pm.write_mol_with_multi_coords(g.sdf_template, poses, str(ranked), 'x', marker='rank', properties={'fitscore': fs})
    recs = _read_records(str(ranked))
    assert recs[3]['name'] == 'x_rank_3' and recs[3]['tail'][:3] == ['M  END', '>  <fitscore>  (4) ', f'{fs[3]}']


@pytest.mark.skipif(not os.access(ANCPHORE, os.X_OK), reason='oracle/_ref/programs/AncPhore not present (build() copies it when '
                    '/root/reference exists)')
def test_pose_output_known_answer_through_ancphore(gold, tmp_path):
    """The reference's shipped poses, written by OUR SD writer from the input ligand file, scored by the reference's AncPhore
    binary against the shipped pharmacophore give the reference's shipped scores, to the printed digit, for every fitness column."""
    g = _template(gold, tmp_path)
    docked, phore = tmp_path / 'docked.sdf', tmp_path / 'ref.phore'
    phore.write_text(str(gold['phore_text']))
    pm.write_mol_with_multi_coords(g.sdf_template, gold['kat_poses'], str(docked), 'sQC_Substrate__STK936575')
    args = (str(docked), str(phore), str(tmp_path / 'o.score'), str(tmp_path / 'o.dbphore'), str(tmp_path / 'o.log'))
    for fit in range(1, 7):
        scores = pp.calc_phore_fitting(*args, overwrite=(fit == 1), fitness=fit, ancphore_path=ANCPHORE)
        assert scores is not None and np.array_equal(np.asarray(scores), gold[f'score_fitness_{fit}']), fit
    assert np.array_equal(np.asarray(pp.calc_phore_fitting(*args, target_fishing=True, ancphore_path=ANCPHORE)),
                          gold['score_fitness_5'])
    # error behaviour: missing binary / missing inputs -> message + None (process_pharmacophore.py:960-970,995-998)
    assert pp.calc_phore_fitting(str(docked), str(phore), str(tmp_path / 'p.score'), 'x', str(tmp_path / 'p.log'),
                                 ancphore_path=str(tmp_path / 'nope')) is None


@pytest.mark.skipif(not os.access(ANCPHORE, os.X_OK), reason='oracle/_ref/programs/AncPhore not present')
def test_pose_sink_and_summary_reproduce_the_reference_example_run(gold, tmp_path, capsys):
    """PoseSink (calculate_fitscore + dock log) and analyze_results on the reference's shipped poses reproduce the files of the
    reference's shipped example run: per-pose fitscores, ranked SD file order / names / fitscore tags, ranked_results.csv."""
    import json
    from types import SimpleNamespace
    import inference
    g = _template(gold, tmp_path)
    phore = tmp_path / 'sQC.phore'
    phore.write_text(str(gold['phore_text']))
    g.name, g.phore_file = 'sQC_Substrate__STK936575', str(phore)
    run_dir = tmp_path / 'run'
    args = SimpleNamespace(run_dir=str(run_dir), out_dir=str(run_dir), fitness=1, ancphore_path=os.path.dirname(ANCPHORE),
                           cutoff=0.4, report_results=True)
    sink = inference.PoseSink(args, workers=2)
    sink.submit(g, gold['kat_poses'], 16.595691680908203)
    (name, scores, run_time), = sink.drain()
    sink.close()
    assert name == g.name and np.array_equal(np.asarray(scores), gold['score_fitness_1'])
    log = json.load(open(run_dir / 'mapping_process' / name / f'{name}_dock.log'))
    assert log == {'name': name, 'fitscore': list(gold['score_fitness_1']), 'run_time': run_time}
    recs = _read_records(str(run_dir / 'ranked_poses' / f'{name}_ranked.sdf'))
    assert [r['name'] for r in recs] == list(gold['ranked_names'])
    assert [r['tail'][2] for r in recs] == list(gold['ranked_fitscore_text'])
    assert np.array_equal(np.asarray([r['xyz'][0] for r in recs]), gold['ranked_first_atom'])
    inference.analyze_results(args, {'name': [name], 'fitscore': [scores], 'run_time': [run_time]})
    assert open(run_dir / 'ranked_results.csv').read() == str(gold['ranked_results_text'])
    assert open(run_dir / 'ranked_results_gt0.4.csv').read() == str(gold['ranked_results_text'])
    assert 'Max fitscore: 0.4781' in capsys.readouterr().out


def test_scoring_failure_sentinel_and_job_plan(gold, tmp_path, capsys):
    from types import SimpleNamespace
    import inference
    g = _template(gold, tmp_path)
    g.name, g.phore_file = 'a__b', str(tmp_path / 'missing.phore')
    args = SimpleNamespace(run_dir=str(tmp_path / 'run'), fitness=1, ancphore_path=str(tmp_path))
    sink = inference.PoseSink(args, workers=1)
    sink.submit(g, gold['kat_poses'][:3], 1.0)
    (name, scores, _), = sink.drain()
    sink.close()
    assert scores == [-2.0] * 3 and 'set as -2.0' in capsys.readouterr().out       # inference.py:235-237
    assert not os.path.exists(tmp_path / 'run' / 'ranked_poses' / 'a__b_ranked.sdf')
    # a pair whose output step raises is reported and skipped, the others still come back
    sink = inference.PoseSink(args, workers=1)
    bad = HeteroGraph()
    bad.name, bad.phore_file, bad.sdf_template = 'x__bad', g.phore_file, None
    sink.submit(bad, gold['kat_poses'][:2], 1.0)
    sink.submit(g, gold['kat_poses'][:2], 1.0)
    res = sink.drain()
    sink.close()
    assert [r[0] for r in res] == ['a__b'] and 'x__bad, skipped' in capsys.readouterr().out
    import torch

    def fake(n, P, tag):
        h = HeteroGraph()
        h['ligand'].pos, h['phore'].pos, h.name = torch.zeros(n, 3), torch.zeros(P, 3), tag
        return h
    same = [fake(20, 79, k) for k in range(1000)]
    jobs = inference.plan_jobs(same, 40, pairs_cap=10 ** 6)
    assert [len(j) for j in jobs][:2] == [103, 103] and sum(len(j) for j in jobs) == 1000 and jobs[-1][-1].name == 999   # stable: input order kept
    assert [len(j) for j in inference.plan_jobs(same[:7], 40, pairs_cap=3)] == [3, 3, 1]
    assert [len(j) for j in inference.plan_jobs(same[:3], 10 ** 5, pairs_cap=8)] == [1, 1, 1]
    # pairs are bucketed by (pharmacophore size, ligand size): a job holds graphs of similar shape
    mixed = [fake(n, P, k) for k, (n, P) in enumerate([(30, 79), (14, 8), (31, 79), (14, 79), (15, 8), (29, 79)])]
    jobs = inference.plan_jobs(mixed, 1, pairs_cap=2, graphs_in_flight=2)
    assert [[g.name for g in j] for j in jobs] == [[1, 4], [3, 5], [0, 2]]


def test_perfect_similarity_matches_the_reference_formula(gold):
    import inference
    for pt, lp, ref in zip(gold['sim_phore_types'], gold['sim_lig_ph'], gold['sim_values']):      # the reference's own function
        gg = HeteroGraph()
        gg['phore'].phoretype = torch.nn.functional.one_hot(torch.from_numpy(pt), 11).float()
        gg['ligand'].ph = torch.from_numpy(lp)
        assert abs(inference.get_perfect_similarity(gg) - ref) <= 1e-6 * max(1.0, abs(ref))
    g = HeteroGraph()
    g['phore'].phoretype = torch.nn.functional.one_hot(torch.tensor([0, 1, 1, 4, 10, 10]), 11).float()
    g['ligand'].ph = torch.tensor([1., 1, 0, 0, 3, 0, 0, 0, 0, 0, 0])
    # weights 1 (EX 0), alpha 1 except AR/HY/CR: every counted type here has the same coefficient -> overlap 3 of volume 4
    assert abs(inference.get_perfect_similarity(g) - 0.75) < 1e-6
    g['phore'].phoretype = torch.nn.functional.one_hot(torch.tensor([10, 10]), 11).float()
    assert inference.get_perfect_similarity(g) == -1.0


class _FakeSampler:
    """Stands in for DenoisingSampler in the host-logic test of inference.fit (no GPU): returns every input pose `samples` times,
    fails for jobs containing a ligand named in `poison` (like a kernel error would)."""
    poison, jobs = (), []

    def __init__(self, *a, **kw):
        self.kw = kw

    def graphs_per_chunk(self, graphs, samples):
        return 1000

    def run(self, graphs, samples, keep_update=False, **kw):
        _FakeSampler.jobs.append([g.name for g in graphs])
        if any(g.name.split('__')[1] in _FakeSampler.poison for g in graphs):
            raise RuntimeError('CUDA error: injected')
        pos = torch.cat([g['ligand'].pos for g in graphs for _ in range(samples)])
        ptr = np.concatenate([[0], np.cumsum([g['ligand'].pos.shape[0] for g in graphs for _ in range(samples)])])
        self.last_trajectory = torch.stack([pos, pos + 1.0]) if keep_update else None
        return pos, ptr


def test_fit_host_logic_jobs_error_isolation_resume_and_keep_update(gold, tmp_path, monkeypatch, capsys):
    """inference.fit around a stand-in sampler: cross-pair jobs, a failing job re-run pair by pair with the offending pair skipped
    (inference.py:211-222), results in input order, resume from the dock logs (:177-183,248-252), keep_update trajectories."""
    from types import SimpleNamespace
    import inference
    monkeypatch.setattr(inference, 'DenoisingSampler', _FakeSampler)
    (tmp_path / 'p.phore').write_text(str(gold['phore_text']))
    graphs = []
    for nm in ['STK936575', 'STK243239', 'STL432840', 'STK255897', 'STK324209']:
        f = tmp_path / f'{nm}.sdf'
        f.write_text(str(gold[f'lig_text_{nm}']))
        graphs.append(inference.build_graph({'ligand_description': str(f), 'phore': str(tmp_path / 'p.phore')}))
    assert graphs[0].name == 'sQC_Substrate__STK936575' and graphs[0]['phore'].phoretype.shape == (79, 11)
    assert torch.allclose(graphs[0]['phore'].pos.mean(0), torch.zeros(3), atol=1e-4)          # centred on the phore centroid
    model = SimpleNamespace(score_norm_tables=lambda: (None, None), kernel_weights=lambda dev: None)
    args = SimpleNamespace(run_dir=str(tmp_path / 'run'), sample_per_complex=3, inference_steps=2, no_final_step_noise=False,
                           ode=False, seed=None, keep_update=True, min_similarity=-1.0, overwrite=False, no_random=False,
                           no_torsion=False, num_workers=2, pairs_per_job=2, fitness=1,
                           ancphore_path=os.path.dirname(ANCPHORE) if os.access(ANCPHORE, os.X_OK) else str(tmp_path))
    _FakeSampler.poison, _FakeSampler.jobs = ('STL432840',), []
    m = inference.fit(args, model, graphs, 'cpu', None)
    names = [g.name for g in graphs]
    by_size = [names[i] for i in sorted(range(5), key=lambda i: graphs[i]['ligand'].pos.shape[0])]      # jobs are bucketed by ligand size
    planned = [by_size[0:2], by_size[2:4], by_size[4:5]]
    expect = []
    for j in planned:                                                                          # the job with the bad pair fails -> pair by pair
        expect.append(j)
        if any('STL432840' in n for n in j) and len(j) > 1:
            expect += [[n] for n in j]
    assert _FakeSampler.jobs == expect
    assert m['name'] == [n for n in names if 'STL432840' not in n]                             # input order, bad pair skipped
    assert 'STL432840 to the reference pharamcophore, skipped' in capsys.readouterr().out
    assert all(len(f) == 3 for f in m['fitscore']) and len(m['dock_poses'][0]) == 3 and len(m['dock_poses'][0][0]) == 1
    center = graphs[0].original_center.numpy()
    # (the reference keeps both trajectories in the centred frame: inference.py:191-192, diffusion_utils.py:75-77)
    assert np.allclose(m['initial_poses'][0][0], graphs[0]['ligand'].pos.numpy(), atol=1e-5)
    assert np.allclose(m['dock_poses'][0][2][0], graphs[0]['ligand'].pos.numpy() + 1.0, atol=1e-5)
    sdf = _read_records(str(tmp_path / 'run' / 'mapping_process' / names[0] / f'{names[0]}.sdf'))
    assert len(sdf) == 3 and np.allclose(sdf[0]['xyz'], graphs[0]['ligand'].pos.numpy() + center, atol=1e-4)
    if os.access(ANCPHORE, os.X_OK):
        # resume: finished pairs come back from their dock logs, only the skipped pair is attempted again
        _FakeSampler.poison, _FakeSampler.jobs = (), []
        m2 = inference.fit(args, model, graphs, 'cpu', None)
        assert _FakeSampler.jobs == [[names[2]]] and m2['name'] == names                       # names[2] = the STL432840 pair
        assert [m2['fitscore'][i] for i in (0, 1, 3, 4)] == m['fitscore']


def test_read_input_equals_the_reference_function(gold, tmp_path):
    """inference.read_input against the reference's own function (inference.py:99-137, run by tools/make_ingest_golden.py on the
    same directory tree): csv with a duplicated row, single pair, directories of pharmacophores / ligands, a .smi file."""
    import json
    import inference
    root = str(tmp_path)
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
    ref = json.loads(str(gold['read_input_json']))
    for k, a in cases.items():
        recs = inference.read_input(*a)
        got = sorted([r['phore'].replace(root, '<ROOT>'), r['ligand_description'].replace(root, '<ROOT>')] for r in recs)
        assert got == ref[k], k
    assert len(ref['csv']) == 2 and len(ref['dirs']) == 6 and len(ref['smi']) == 2
    with pytest.raises(ValueError, match='Invalid input'):
        inference.read_input(None, os.path.join(root, 'missing.phore'), os.path.join(root, 'ligs'))


def test_get_model_kwarg_mapping_equals_the_reference_function(gold, monkeypatch):
    """utils.utils.get_model against the reference's own get_model (utils/utils.py:113-168, run by tools/make_ingest_golden.py with
    a recording constructor on the shipped model_parameters.yml): every constructor kwarg the mirror derives from the Namespace has
    the reference's value, the reference-only kwargs are accepted by the mirrored model class, and the embedding spec is the same."""
    import json
    import yaml
    from argparse import Namespace
    from utils import utils as uu
    ref = json.loads(str(gold['get_model_kwargs_json']))
    args = Namespace(**yaml.full_load(str(gold['model_parameters_yml'])))
    args.no_torsion = False
    captured, emb = {}, {}

    class Recorder:
        def __init__(self, **kw):
            captured.update(kw)

        def to(self, device):
            return self
    monkeypatch.setattr(uu, 'PhoreModel', Recorder)
    monkeypatch.setattr(uu, 'get_timestep_embedding', lambda *a, **kw: emb.update(args=a, kw=kw) or 'emb')
    uu.get_model(args, torch.device('cpu'), t_to_sigma=None, no_parallel=True)
    mine = {k: v for k, v in captured.items() if k not in ('t_to_sigma', 'device', 'timestep_emb_func')}
    assert set(mine) <= set(ref), set(mine) - set(ref)
    assert all(mine[k] == ref[k] for k in mine), {k: (mine[k], ref[k]) for k in mine if mine[k] != ref[k]}
    spec = json.loads(str(gold['get_model_emb_json']))
    got = list(emb['args']) + [emb['kw'][k] for k in ('embedding_type', 'embedding_dim', 'embedding_scale') if k in emb['kw']]
    assert got == [spec['embedding_type'], spec['embedding_dim'], spec['embedding_scale']]
    # the kwargs only the reference passes select features outside the shipped flag set; the mirrored class takes them all
    from models.score_model_phore import TensorProductScoreModel
    model = TensorProductScoreModel(t_to_sigma=None, device=torch.device('cpu'), timestep_emb_func=None, **ref)
    assert len(model.state_dict()) == 385


def test_cli_flags_and_defaults_equal_the_reference_parser(gold):
    """inference.parse_args against the reference's own parser (inference.py:54-96): same flags, same defaults (except --model_dir,
    which defaults to the shipped weights directory here), same str2bool / target_fishing handling; two new optional flags."""
    import json
    import inference
    ref = json.loads(str(gold['parse_args_defaults_json']))
    mine = vars(inference.parse_args([]))
    assert set(ref) <= set(mine) and set(mine) - set(ref) == {'seed', 'pairs_per_job'}
    assert {k for k in ref if ref[k] != mine[k]} == {'model_dir'}
    ref = json.loads(str(gold['parse_args_flags_json']))
    mine = vars(inference.parse_args(['--target_fishing', 'true', '--no_random', '--ode', '--overwrite', 'yes', '--cutoff', '0.4']))
    assert {k for k in ref if ref[k] != mine[k]} == {'model_dir'} and mine['fitness'] == ref['fitness']
