"""Ligand–pharmacophore mapping with the B200 denoising path — same entry point and flags as the reference's
`src/inference.py` (parse_args :54-96, read_input :99-137, fit :139-271, analyze_results :321-350, main :382-468).

    python src/inference.py --phore examples/phore/X.phore --ligand examples/ligands/Y.sdf \
           --model_dir weights/diffphore_calibrated_warmuped_ft --out_dir results/run --sample_per_complex 40

Differences by design: `fit` batches ACROSS pairs: pending pairs x samples are denoised in jobs of
~4096 graphs by one DenoisingSampler instead of one pair at a time, and the SD writing + AncPhore scoring of a finished
job overlaps the next job (PoseSink); per-pair `run_time` is the job time divided evenly over its pairs.  Preprocessing: ALWAYS the
reduced RDKit-free featuriser for 3-D SD files (datasets/process_mols.py: crude aromaticity / hybridisation / charge / chirality
and phorefp / norm typing; RDKit is absent from this image and no RDKit branch exists) - a warning is printed once per run;
poses and fitscores of real ligands are therefore NOT comparable one-to-one with the reference's RDKit pipeline
(process_mols.py:255-417).  SMILES / .smi ligands pass read_input but are skipped (no conformer generation without RDKit).  Scoring: AncPhore binary when available
(`--ancphore_path`), else fitscore = -2.0 like the reference's failure sentinel (inference.py:235-237).
"""
import _bootstrap  # noqa: F401  (repo root on sys.path)
import json
import os
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
    p.add_argument('--pairs_per_job', type=int, default=None, help='(new) pairs denoised together per GPU job; default: '
                   'as many as give ~4096 graphs in flight and fit the resident HBM budget')
    args = p.parse_args(argv)
    if args.target_fishing:
        args.fitness = 5
    return args


def _expand(path, smi_lines=False):
    """A directory stands for every entry in it, a `.smi` file (ligand side only) for its lines, anything else for itself."""
    if os.path.isdir(path):
        return [os.path.join(path, name) for name in os.listdir(path)]
    if smi_lines and path.endswith('.smi'):
        with open(path) as fh:
            return [line.strip() for line in fh]
    return [path]


def read_input(phore_ligand_csv=None, phore=None, ligand=None):
    """Records {phore, ligand_description} with the input rules of the reference (inference.py:99-137; outputs pinned on its own
    function in tests/test_ingest.py): an existing csv with the columns `phore` / `ligand_description` wins (duplicate rows
    dropped); otherwise the product of the pharmacophores and ligands named on the command line, provided both paths exist."""
    if phore_ligand_csv is not None and os.path.exists(phore_ligand_csv):
        import pandas as pd
        records = pd.read_csv(phore_ligand_csv).drop_duplicates().to_dict('records')
    elif phore is not None and ligand is not None and os.path.exists(phore) and os.path.exists(ligand):
        records = [{'phore': ph, 'ligand_description': lg} for ph in _expand(phore) for lg in _expand(ligand, smi_lines=True)]
    else:
        records = []
    if not records:
        raise ValueError('Invalid input. Either phore_ligand_csv or protein and ligand must be specified')
    return records


def build_graph(record):
    """generate_graph (datasets/pdbbind_phore.py:1143-1188) for one record: ligand + pharmacophore tensors, phoretype
    one-hot, everything centred on the pharmacophore centroid (kept in `original_center`)."""
    lig_file, phore_file = record['ligand_description'], record['phore']
    if not build_graph.warned:
        build_graph.warned = True
        warnings.warn('RDKit-free REDUCED ligand featuriser active (no RDKit in this image): atom features, phorefp and norm typing '
                      'are approximations of process_mols.py:193-417; poses / fitscores are not comparable one-to-one with the '
                      "reference's RDKit pipeline", stacklevel=2)
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


build_graph.warned = False


def get_perfect_similarity(g, weights=(1.0, 1.0, 1.0, 1.0, 1.0, 0.0, 1.0, 1.0, 1.0, 1.0, 0.0),      # HY and EX do not count
                           alpha=(1.0, 1.0, 0.7, 1.0, 1.0, 0.7, 1.0, 1.0, 0.7, 1.0, 0.837)):
    """Type/count-only pharmacophore fingerprint similarity used by --min_similarity (inference.py:273-312)."""
    phore_volume = g['phore'].phoretype.sum(dim=0)
    overlap = torch.min(g['ligand'].ph, phore_volume)
    coeff = torch.tensor(weights).float()
    if alpha is not None:
        coeff = coeff * 7.999999999 * (torch.tensor(alpha) * torch.pi / 2) ** 1.5
    weighted_volume = (phore_volume * coeff).sum()
    if weighted_volume == 0:
        return -1.0
    return ((overlap * coeff).sum() / weighted_volume).item()


class PoseSink:
    """Pose output + scoring hand-off (calculate_fitscore, sampling.py:447-498; dock log, inference.py:242-246) off the
    GPU's critical path: a thread pool writes `mapping_process/{name}/{name}.sdf`, runs the AncPhore binary
    on it (a subprocess, so threads overlap), writes `ranked_poses/{name}_ranked.sdf` and `{name}_dock.log`, while the
    next job of pairs is being denoised."""

    def __init__(self, args, workers=8):
        from concurrent.futures import ThreadPoolExecutor
        self.args = args
        self.pool = ThreadPoolExecutor(max_workers=max(1, workers))
        self.pending = []

    def submit(self, g, poses, run_time):
        self.pending.append((g.name, self.pool.submit(self.process, g, poses, run_time)))

    def process(self, g, poses, run_time):
        args, name, N = self.args, g.name, len(poses)
        tmp = os.path.join(args.run_dir, f'mapping_process/{name}')
        os.makedirs(tmp, exist_ok=True)
        docked_file = os.path.join(tmp, f'{name}.sdf')
        write_mol_with_multi_coords(g.sdf_template, poses, docked_file, name)
        scores = calc_phore_fitting(docked_file, g.phore_file, os.path.join(tmp, f'{name}.score'),
                                    os.path.join(tmp, f'{name}.dbphore'), os.path.join(tmp, f'{name}.log'),
                                    overwrite=True, fitness=getattr(args, 'fitness', 1),
                                    ancphore_path=os.path.join(args.ancphore_path, 'AncPhore'))
        if scores is not None and len(scores) == N:
            ranked_dir = os.path.join(args.run_dir, 'ranked_poses')
            os.makedirs(ranked_dir, exist_ok=True)
            perm = np.argsort(np.array(scores))[::-1]
            write_mol_with_multi_coords(g.sdf_template, poses[perm], os.path.join(ranked_dir, f'{name}_ranked.sdf'), name,
                                        marker='rank', properties={'fitscore': np.array(scores)[perm]})
        json.dump({'name': name, 'fitscore': scores, 'run_time': run_time},
                  open(os.path.join(tmp, f'{name}_dock.log'), 'w'), indent=4)
        if scores is None or len(scores) == 0:
            print(f'[W] fitscore calculated with error and set as -2.0 for `{name}`')
            scores = [-2.0] * N
        return name, scores, run_time

    def drain(self):
        """Results of everything submitted so far, in submission order.  A pair whose output / scoring step raised (disk full,
        unreadable template, ...) is reported and skipped like any other failing pair; it does not take the job down."""
        done = []
        for name, f in self.pending:
            try:
                done.append(f.result())
            except Exception as e:
                print(f'[W] Error occured when writing / scoring the poses of {name}, skipped. {e}')
        self.pending = []
        return done

    def close(self):
        self.pool.shutdown(wait=True)


def plan_jobs(graphs, samples, pairs_cap, graphs_in_flight=4096):
    """Cross-pair batching: pairs are BUCKETED by (pharmacophore size, ligand size) - a stable sort, so that a job
    holds graphs of similar shape: the per-graph kernels (one CTA per graph, sized by the largest ligand of the chunk) and the
    256-edge pair tiles stay full, and a screening run over one pharmacophore keeps its ligands of equal size together - and then cut
    into jobs of about `graphs_in_flight` (pair, sample) graphs - enough to fill 148 SMs with 256-edge tile pairs - and never more
    than `pairs_cap` pairs (what fits the resident HBM budget).  Results are reported in input order (fit re-orders them).
    The reference batches only the samples of ONE pair (inference.py:184, sampling.py:210)."""
    per_job = max(1, min(pairs_cap, -(-graphs_in_flight // max(1, samples))))
    order = sorted(range(len(graphs)), key=lambda i: (graphs[i]['phore'].pos.shape[0], graphs[i]['ligand'].pos.shape[0]))
    ordered = [graphs[i] for i in order]
    return [ordered[k:k + per_job] for k in range(0, len(ordered), per_job)]


def fit(args, model, complex_graphs, device, t_to_sigma, tmp_log='', n_report=1000):
    """The reference's `fit` (inference.py:139-271) with its per-pair loop turned inside out: pairs that still need fitting are
    denoised job by job (plan_jobs) with all their samples resident on the GPU, and every finished job is handed to the
    PoseSink, which scores it while the next job runs.  Outputs, resume rule (:177-183, 248-252) and error rule (a failing
    pair is reported and skipped, :211-222) are the reference's; `run_time` of a pair = its job's time / pairs in the job."""
    N = getattr(args, 'sample_per_complex', 1)
    so3n, torn = model.score_norm_tables()
    sampler = DenoisingSampler(model.kernel_weights(device), args.inference_steps, so3n, torn,
                               no_final_step_noise=args.no_final_step_noise, ode=getattr(args, 'ode', False))
    gen = None
    if getattr(args, 'seed', None) is not None:
        # one stream per rank: under torchrun every rank denoises different pairs (round-robin), so equal seeds would hand them
        # identical initial-pose / noise streams
        gen = torch.Generator(device=device).manual_seed(args.seed + int(os.environ.get('RANK', 0)))
    keep = bool(getattr(args, 'keep_update', False))
    done, todo = {}, []
    for g in complex_graphs:
        if getattr(args, 'min_similarity', -1.0) > 0:
            max_sim = get_perfect_similarity(g)
            if max_sim < args.min_similarity:
                print(f'[I] `{g.name}` is excluded due to pharmacophore fingerprint similarity constraints '
                      f'({max_sim:.2f} < {args.min_similarity:.2f}).')
                continue
        docked = os.path.join(args.run_dir, f'ranked_poses/{g.name}_ranked.sdf')
        log_file = os.path.join(args.run_dir, f'mapping_process/{g.name}/{g.name}_dock.log')
        if os.path.exists(docked) and os.path.exists(log_file) and not args.overwrite:
            log = json.load(open(log_file))
            done[g.name] = (log['name'], log['fitscore'], log['run_time'])
        elif g['ligand'].pos.shape[0] == 0:
            print(f'[W] Graph {g.name} with 0 atoms, skipped')
        else:
            todo.append(g)
    seen = {}
    for g in todo:                                      # duplicate names would make two sink tasks write the same files
        k = seen.get(g.name, 0)
        seen[g.name] = k + 1
        if k:
            print(f'[W] duplicate pair name `{g.name}`: outputs of this copy go to `{g.name}_dup{k}`')
            g.name = f'{g.name}_dup{k}'
    sink = PoseSink(args, workers=min(getattr(args, 'num_workers', 8) or 1, os.cpu_count() or 1))
    initial_poses, dock_poses = {}, {}
    n_done, std_time = 0, time.time()

    def run_job(job):
        t0 = time.time()
        pos, ptr = sampler.run(job, N, no_random=args.no_random, generator=gen, no_torsion=args.no_torsion, keep_update=keep)
        run_time = (time.time() - t0) / len(job)
        out = []
        for i, g in enumerate(job):
            n, lo = g['ligand'].pos.shape[0], ptr[i * N]
            center = g.original_center.numpy()
            out.append((g, pos[lo:lo + N * n].reshape(N, n, 3).numpy() + center))
            if keep:                                    # inference.py:191-192,247-248 (initial pose, pose after every step)
                traj = sampler.last_trajectory[:, lo:lo + N * n].reshape(-1, N, n, 3).numpy()
                initial_poses[g.name] = [traj[0, s] for s in range(N)]
                dock_poses[g.name] = [[traj[k, s] for k in range(1, traj.shape[0])] for s in range(N)]
        for g, poses in out:                            # hand-off only once the WHOLE job succeeded: the pair-by-pair fallback
            sink.submit(g, poses, run_time)             # below must never find a pair of this job already submitted

    pairs_cap = sampler.graphs_per_chunk(todo, N) if todo else 1
    jobs = plan_jobs(todo, N, getattr(args, 'pairs_per_job', None) or pairs_cap)
    for job in jobs:
        try:
            run_job(job)
        except Exception as e:                          # isolate the offending pair: rerun the job one pair at a time
            if len(job) == 1:
                print(f'[W] Error occured when fitting {job[0].name} to the reference pharamcophore, skipped. {e}')
            for g in (job if len(job) > 1 else []):
                try:
                    run_job([g])
                except Exception as e1:
                    print(f'[W] Error occured when fitting {g.name} to the reference pharamcophore, skipped. {e1}')
        n_done += len(job)
        if tmp_log and n_done // n_report != (n_done - len(job)) // n_report:
            print(f'[I] {n_done}/{len(todo)} processed...')
            part = [done[k] for k in done] + [f.result() for _, f in sink.pending if f.done() and f.exception() is None]
            json.dump({'name': [r[0] for r in part], 'fitscore': [r[1] for r in part], 'run_time': [r[2] for r in part],
                       'batch': n_done, 'total_time': time.time() - std_time}, open(tmp_log, 'w'), indent=4)
    for r in sink.drain():
        done[r[0]] = r
    sink.close()
    order = [g.name for g in complex_graphs if g.name in done]          # the reference reports in input order
    metrics = {'name': order, 'fitscore': [done[k][1] for k in order], 'run_time': [done[k][2] for k in order]}
    if keep:
        metrics['initial_poses'] = [initial_poses.get(k) for k in order]
        metrics['dock_poses'] = [dock_poses.get(k) for k in order]
    return metrics


def analyze_results(args, results):
    """Summary table `ranked_results.csv` (tab separated; columns and ordering of inference.py:321-350)."""
    import pandas as pd
    df = pd.DataFrame({k: results[k] for k in ('name', 'fitscore', 'run_time')})
    df['max_fitscore'] = df['fitscore'].map(lambda x: max(x) if len(x) else -2.0)
    df['top5_mean_fitscore'] = df['fitscore'].map(lambda x: np.sort(x)[-5:].mean().item())
    df['target'] = df['name'].map(lambda x: x.split('__')[0])
    df['ligand'] = df['name'].map(lambda x: x.split('__')[1])
    df = df.sort_values(by=['max_fitscore', 'top5_mean_fitscore'], ascending=False)
    dump_file = os.path.join(args.out_dir, 'ranked_results.csv')
    print(f'[I] Dumping results to `{dump_file}`')
    df = df[['target', 'ligand', 'name', 'run_time', 'max_fitscore', 'top5_mean_fitscore', 'fitscore']]
    df.to_csv(dump_file, sep='\t', index=False)
    if args.cutoff is not None:
        df[df['max_fitscore'] >= args.cutoff].to_csv(os.path.join(args.out_dir, f'ranked_results_gt{args.cutoff}.csv'),
                                                     sep='\t', index=False)
    if args.report_results:
        print()
        print('#' * 25 + ' Pharmacophore Alignment Summary ' + '#' * 25)
        for thr in (0.7, 0.4):
            k = len(df[df['max_fitscore'] >= thr])
            print(f'Number of ligands with fitscore greater than {thr}: {k} ({100 * k / len(df):.2f}%)')
        print(f"Max fitscore: {df['max_fitscore'].max().item():.4f}")
        print(f"Average max fitscore: {df['max_fitscore'].mean().item():.4f}")
        print(f"Average runtime: {df['run_time'].mean().item():.4f}")
    return df


def _ranks():
    """(rank, world, local_rank) from the torchrun environment; (0, 1, 0) for a plain `python src/inference.py`."""
    return int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))


def merge_rank_results(parts, order):
    """Per-rank result dicts of `fit` -> one dict in the input order of the pairs (`order`: all pair names)."""
    rows = {}
    for part in parts:
        for i, name in enumerate(part['name']):
            rows[name] = {k: v[i] for k, v in part.items()}
    names = [n for n in order if n in rows]
    keys = [k for k in (parts[0].keys() if parts else [])]
    return {k: [rows[n][k] for n in names] for k in keys} if names else {'name': [], 'fitscore': [], 'run_time': []}


def gather_results(results, all_names, world):
    """The one exchange of a multi-rank run: every rank contributes the results table of its pairs, every rank gets the merged
    table (a few floats per pose; gloo over TCP on the host - the poses themselves are already on disk)."""
    import torch.distributed as dist
    own_group = not dist.is_initialized()
    if own_group:
        dist.init_process_group('gloo')
    parts = [None] * world
    dist.all_gather_object(parts, results)
    if own_group:
        dist.destroy_process_group()
    return merge_rank_results(parts, all_names)


def main(argv=None):
    """Single process: one GPU.  Under `torchrun --nproc-per-node N src/inference.py ...` the pairs are dealt round-robin to the N
    ranks (one GPU each, SURVEY §8e: the path shards embarrassingly), every rank writes the SD files / dock logs of its own
    pairs, and the only exchange is one gather of the per-pair results table at the end (gloo; poses never leave their rank)."""
    warnings.filterwarnings('ignore', category=UserWarning)
    args = parse_args(argv)
    rank, world, local_rank = _ranks()
    result_file = os.path.join(args.out_dir, 'inference_results.json')
    with open(f'{args.model_dir}/model_parameters.yml') as f:
        score_model_args = Namespace(**yaml.full_load(f))
    for k in ('sample_per_complex', 'inference_steps', 'actual_steps', 'ancphore_path', 'ode', 'no_torsion', 'no_random',
              'no_final_step_noise', 'overwrite', 'min_similarity', 'keep_update', 'fitness', 'seed', 'num_workers',
              'pairs_per_job'):
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
        device = torch.device('cuda', local_rank)
        torch.cuda.set_device(device)
        all_names = [g.name for g in graphs]
        graphs = graphs[rank::world]                                        # this rank's pairs
        model = get_model(score_model_args, device, t_to_sigma=t_to_sigma, no_parallel=True)
        print(f'[I] Loading state dict from `{args.model_dir}/{args.ckpt}`')
        state_dict = torch.load(f'{args.model_dir}/{args.ckpt}', map_location=torch.device('cpu'), weights_only=False)
        model.load_state_dict(state_dict, strict=True)
        model.eval()
        print('\n>> Starting to fit <<')
        print(f"[I] Please check the process files in `{os.path.join(args.out_dir, 'mapping_process/')}`")
        print(f"[I] Please check the ranked poses in `{os.path.join(args.out_dir, 'ranked_poses/')}`")
        results = fit(score_model_args, model, graphs, device, t_to_sigma,
                      tmp_log=result_file + ('.tmp' if world == 1 else f'.tmp{rank}'))
        for tmp in (result_file + '.tmp', result_file + f'.tmp{rank}'):
            if os.path.exists(tmp):
                os.remove(tmp)
        if world > 1:
            results = gather_results(results, all_names, world)
            if rank != 0:
                return results
        if args.keep_update:
            import pickle
            pickle.dump(results, open(result_file + '.pkl', 'wb'))
        else:
            json.dump(results, open(result_file, 'w'), indent=4)
    else:
        results = json.load(open(result_file))
    if rank == 0 and results and results['name']:
        analyze_results(args, results)
    return results


if __name__ == '__main__':
    st = time.time()
    main()
    print(f'Job Finished! {time.time() - st:.3f} seconds cost.')
