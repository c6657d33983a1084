"""Ligand–pharmacophore mapping with the B200 denoising path — same entry point and flags as the reference's
`src/inference.py` (parse_args :54-96, read_input :99-137, fit :139-271, analyze_results :321-350, main :382-468).

    python src/inference.py --phore examples/phore/X.phore --ligand examples/ligands/Y.sdf \
           --model_dir weights/diffphore_calibrated_warmuped_ft --out_dir results/run --sample_per_complex 40

Differences by design: `fit` batches ACROSS pairs: all pending pairs x samples are denoised in
HBM-sized chunks by one DenoisingSampler instead of one pair at a time; per-pair `run_time` is therefore the chunk
time divided evenly over its pairs.  Preprocessing: RDKit path when RDKit is importable (not in this image), else the
reduced RDKit-free featuriser for 3-D SD files (datasets/process_mols.py).  Scoring: AncPhore binary when available
(`--ancphore_path`), else fitscore = -2.0 like the reference's failure sentinel (inference.py:235-237).
"""
import _bootstrap  # noqa: F401  (repo root on sys.path)
import copy
import json
import os
import shutil
import time
import warnings
from argparse import ArgumentParser, FileType, Namespace
from functools import partial

import numpy as np
import torch
import yaml

from datasets.process_mols import ligand_graph_from_sdf, write_mol_with_multi_coords
from datasets.process_pharmacophore import parse_phore, get_phore_graph, calc_phore_fitting
from diffphore_b200.graph import HeteroGraph
from diffphore_b200.sampler import DenoisingSampler
from utils.diffusion_utils import t_to_sigma as t_to_sigma_compl
from utils.utils import get_model


def str2bool(v):
    return v if isinstance(v, bool) else str(v).lower() in ('yes', 'true', 't', 'y', '1')


def parse_args(argv=None):
    p = ArgumentParser()
    p.add_argument('--config', type=FileType(mode='r'), default=None)
    p.add_argument('--phore_ligand_csv', type=str, default=None)
    p.add_argument('--phore', type=str, default=None)
    p.add_argument('--ligand', type=str, default=None)
    p.add_argument('--out_dir', type=str, default='results/user_inference')
    p.add_argument('--cache_path', type=str, default='data/cache')
    p.add_argument('--split_file', type=str, default='data/splits/timesplit_no_lig_overlap_val')
    p.add_argument('--overwrite', type=str2bool, default=False)
    p.add_argument('--keep_local_structures', type=str2bool, default=False)
    p.add_argument('--sample_per_complex', type=int, default=40)
    p.add_argument('--save_visualisation', action='store_true', default=False)
    p.add_argument('--model_dir', type=str, default='../weights/diffphore_calibrated_warmuped_ft')
    p.add_argument('--ckpt', type=str, default='best_ema_inference_epoch_model.pt')
    p.add_argument('--batch_size', type=int, default=32)
    p.add_argument('--num_workers', type=int, default=40)
    p.add_argument('--inference_steps', type=int, default=20)
    p.add_argument('--actual_steps', type=int, default=None)
    p.add_argument('--no_random', action='store_true', default=False)
    p.add_argument('--ancphore_path', type=str, default='../programs/')
    p.add_argument('--no_final_step_noise', action='store_true', default=False)
    p.add_argument('--ode', action='store_true', default=False)
    p.add_argument('--no_torsion', action='store_true', default=False)
    p.add_argument('--cutoff', type=float, default=None)
    p.add_argument('--min_similarity', type=float, default=-1.0)
    p.add_argument('--report_results', type=str2bool, default=True)
    p.add_argument('--keep_update', type=str2bool, default=False)
    p.add_argument('--fitness', type=int, default=1)
    p.add_argument('--target_fishing', type=str2bool, default=False)
    p.add_argument('--seed', type=int, default=None, help='(new) seed of the device RNG; the reference is unseeded')
    args = p.parse_args(argv)
    if args.target_fishing:
        args.fitness = 5
    return args


def read_input(phore_ligand_csv=None, phore=None, ligand=None):
    """Records {phore, ligand_description} from a csv (columns ligand_description, phore) or a single / listed pair."""
    records = []
    if phore_ligand_csv is not None:
        import csv
        with open(phore_ligand_csv) as fh:
            for row in csv.DictReader(fh):
                records.append({'phore': row['phore'], 'ligand_description': row['ligand_description']})
    elif phore is not None and ligand is not None:
        def expand(x):
            if os.path.isfile(x) and not x.endswith(('.sdf', '.mol', '.mol2', '.phore')):
                return [l.strip() for l in open(x) if l.strip()]
            return [x]
        for ph in expand(phore):
            for lg in expand(ligand):
                records.append({'phore': ph, 'ligand_description': lg})
    if not records:
        raise ValueError('Invalid input. Either phore_ligand_csv or protein and ligand must be specified')
    return records


def build_graph(record):
    """generate_graph (datasets/pdbbind_phore.py:1143-1188) for one record: ligand + pharmacophore tensors, phoretype
    one-hot, everything centred on the pharmacophore centroid (kept in `original_center`)."""
    lig_file, phore_file = record['ligand_description'], record['phore']
    if not (os.path.exists(lig_file) and lig_file.endswith('.sdf')):
        raise NotImplementedError('without RDKit only 3-D .sdf ligands can be ingested (SMILES needs conformer generation)')
    g = HeteroGraph()
    phore = parse_phore(phore_file)[0]
    ligand_graph_from_sdf(lig_file, g)
    get_phore_graph(phore, g, consider_ex=True, neighbor_cutoff=5.0, ex_connected=True)
    ph = g['phore']
    ph.phoretype = torch.nn.functional.one_hot(ph.x[:, 0].long(), 11).float()
    center = ph.pos.mean(0, keepdim=True)
    ph.pos = ph.pos - center
    g['ligand'].pos = g['ligand'].pos - center
    g.original_center = center
    g.name = f"{phore.id}__{os.path.splitext(os.path.basename(lig_file))[0]}"
    g.phore_file = phore_file
    return g


def fit(args, model, complex_graphs, device, t_to_sigma, tmp_log='', n_report=1000):
    N = getattr(args, 'sample_per_complex', 1)
    so3n, torn = model.score_norm_tables()
    sampler = DenoisingSampler(model.kernel_weights(device), args.inference_steps, so3n, torn,
                               no_final_step_noise=args.no_final_step_noise)
    gen = None
    if getattr(args, 'seed', None) is not None:
        gen = torch.Generator(device=device).manual_seed(args.seed)
    names, fitscore, run_times = [], [], []
    todo = []
    for g in complex_graphs:
        docked = os.path.join(args.run_dir, f'ranked_poses/{g.name}_ranked.sdf')
        log_file = os.path.join(args.run_dir, f'mapping_process/{g.name}/{g.name}_dock.log')
        if os.path.exists(docked) and os.path.exists(log_file) and not args.overwrite:      # resume (inference.py:180-183)
            log = json.load(open(log_file))
            names.append(log['name']); fitscore.append(log['fitscore']); run_times.append(log['run_time'])
        elif g['ligand'].pos.shape[0] == 0:
            print(f'[W] Graph {g.name} with 0 atoms, skipped')
        else:
            todo.append(g)
    if todo:
        t0 = time.time()
        pos, ptr = sampler.run(todo, N, no_random=args.no_random, generator=gen, no_torsion=args.no_torsion)
        run_time = (time.time() - t0) / len(todo)
        for i, g in enumerate(todo):
            n = g['ligand'].pos.shape[0]
            poses = pos[ptr[i * N]:ptr[(i + 1) * N]].reshape(N, n, 3).numpy() + g.original_center.numpy()
            tmp = os.path.join(args.run_dir, f'mapping_process/{g.name}')
            os.makedirs(tmp, exist_ok=True)
            docked_file = os.path.join(tmp, f'{g.name}.sdf')
            write_mol_with_multi_coords(g.sdf_template, poses, docked_file, g.name)
            os.environ.setdefault('ANCPHORE', os.path.join(args.ancphore_path, 'AncPhore'))
            scores = calc_phore_fitting(docked_file, g.phore_file, os.path.join(tmp, f'{g.name}.score'),
                                        os.path.join(tmp, f'{g.name}.dbphore'), os.path.join(tmp, f'{g.name}.log'),
                                        overwrite=True, fitness=getattr(args, 'fitness', 1))
            if not scores:
                print(f'[W] fitscore calculated with error and set as -2.0 for `{g.name}`')
                scores = [-2.0] * N
            os.makedirs(os.path.join(args.run_dir, 'ranked_poses'), exist_ok=True)
            perm = np.argsort(np.asarray(scores))[::-1]
            write_mol_with_multi_coords(g.sdf_template, poses[perm], os.path.join(args.run_dir, f'ranked_poses/{g.name}_ranked.sdf'),
                                        g.name, marker='rank', properties={'fitscore': np.asarray(scores)[perm]})
            names.append(g.name); fitscore.append(list(map(float, scores))); run_times.append(run_time)
            json.dump({'name': g.name, 'fitscore': list(map(float, scores)), 'run_time': run_time},
                      open(os.path.join(tmp, f'{g.name}_dock.log'), 'w'), indent=4)
    return {'name': names, 'fitscore': fitscore, 'run_time': run_times}


def analyze_results(args, results):
    import pandas as pd
    df = pd.DataFrame(results)
    df['max_fitscore'] = df['fitscore'].map(lambda x: max(x) if len(x) else -2.0)
    df['top5_mean_fitscore'] = df['fitscore'].map(lambda x: float(np.sort(x)[-5:].mean()))
    df['target'] = df['name'].map(lambda x: x.split('__')[0])
    df['ligand'] = df['name'].map(lambda x: x.split('__')[1])
    df = df.sort_values(by=['max_fitscore', 'top5_mean_fitscore'], ascending=False)
    cols = ['target', 'ligand', 'max_fitscore', 'top5_mean_fitscore', 'run_time']
    df[cols].to_csv(os.path.join(args.out_dir, 'ranked_results.csv'), index=False)
    if args.report_results:
        print(df[cols].head(20).to_string(index=False))
    return df


def main(argv=None):
    warnings.filterwarnings('ignore', category=UserWarning)
    args = parse_args(argv)
    result_file = os.path.join(args.out_dir, 'inference_results.json')
    with open(f'{args.model_dir}/model_parameters.yml') as f:
        score_model_args = Namespace(**yaml.full_load(f))
    for k in ('sample_per_complex', 'inference_steps', 'actual_steps', 'ancphore_path', 'ode', 'no_torsion', 'no_random',
              'no_final_step_noise', 'overwrite', 'min_similarity', 'keep_update', 'fitness', 'seed'):
        setattr(score_model_args, k, getattr(args, k))
    score_model_args.run_dir = args.out_dir
    t_to_sigma = partial(t_to_sigma_compl, args=score_model_args)
    records = read_input(args.phore_ligand_csv, args.phore, args.ligand)
    graphs = []
    for r in records:
        try:
            graphs.append(build_graph(r))
        except Exception as e:                                              # the reference skips unreadable inputs too
            print(f"[W] Failed to process {r}: {e}")
    print('[I] Number of fitting samples:', len(graphs))
    if not graphs:
        print('[E] No valid fitting samples, please check your input. exit.')
        return None
    if not os.path.exists(result_file) or args.overwrite:
        os.makedirs(args.out_dir, exist_ok=True)
        if not torch.cuda.is_available():
            raise RuntimeError('the B200 denoising path needs a CUDA device (there is no CPU fallback)')
        device = torch.device('cuda')
        model = get_model(score_model_args, device, t_to_sigma=t_to_sigma, no_parallel=True)
        print(f'[I] Loading state dict from `{args.model_dir}/{args.ckpt}`')
        state_dict = torch.load(f'{args.model_dir}/{args.ckpt}', map_location=torch.device('cpu'), weights_only=False)
        model.load_state_dict(state_dict, strict=True)
        model.eval()
        print('\n>> Starting to fit <<')
        results = fit(score_model_args, model, graphs, device, t_to_sigma, tmp_log=result_file + '.tmp')
        json.dump(results, open(result_file, 'w'), indent=4)
    else:
        results = json.load(open(result_file))
    if results and results['name']:
        analyze_results(args, results)
    return results


if __name__ == '__main__':
    st = time.time()
    main()
    print(f'Job Finished! {time.time() - st:.3f} seconds cost.')
