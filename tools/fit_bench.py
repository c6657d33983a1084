"""cfg3 end to end through the mirrored CLI (SURVEY §8f-1/2 measurement): the reference's examples/task_file.csv (15 ligands x the
79-node example pharmacophore, 40 samples each, shipped checkpoint) from text files to ranked SD files + AncPhore fitscores.

    python tools/fit_bench.py [--repeat 2]          (GPU box; inputs come from tests/golden/ingest.npz, the checkpoint and the
                                                     AncPhore binary from oracle/_ref/, which build() fills in the build container)
Prints one JSON line: wall time of `fit` (denoising + SD writing + scoring) for (a) the cross-pair job scheduler with the scoring
pool overlapped and (b) the reference's organisation (one pair per job, scoring in line), next to the sum of the `run_time`s the
reference ships for the same job (examples/output/2/inference_results.json: 199.4 s denoising only, GPU unstated).
"""
import argparse
import contextlib
import io
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'src'))
import inference                                     # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--repeat', type=int, default=2)
    ap.add_argument('--samples', type=int, default=40)
    a = ap.parse_args()
    gold = np.load(os.path.join(ROOT, 'tests/golden/ingest.npz'))
    tmp = tempfile.mkdtemp(prefix='fit_bench_')
    open(os.path.join(tmp, 'sQC_QFA_complex.phore'), 'w').write(str(gold['phore_text']))
    rows = ['ligand_description,phore']
    for line in str(gold['task_file_text']).strip().split('\n')[1:]:
        nm = os.path.basename(line.split(',')[0])[:-4]
        open(os.path.join(tmp, nm + '.sdf'), 'w').write(str(gold[f'lig_text_{nm}']))
        rows.append(f"{os.path.join(tmp, nm + '.sdf')},{os.path.join(tmp, 'sQC_QFA_complex.phore')}")
    open(os.path.join(tmp, 'task.csv'), 'w').write('\n'.join(rows) + '\n')
    n_pairs = len(rows) - 1
    base = ['--phore_ligand_csv', os.path.join(tmp, 'task.csv'), '--model_dir', os.path.join(ROOT, 'oracle/_ref/weights'),
            '--sample_per_complex', str(a.samples), '--ancphore_path', os.path.join(ROOT, 'oracle/_ref/programs'),
            '--overwrite', 'true', '--report_results', 'false', '--seed', '1']
    res = {}
    for tag, extra in (('overlapped_jobs', []), ('one_pair_per_job_inline_scoring', ['--pairs_per_job', '1', '--num_workers', '1'])):
        best, scored = None, None
        for r in range(a.repeat + 1):                       # first pass warms up (library load, CUDA context, graphs)
            out = os.path.join(tmp, f'out_{tag}_{r}')
            t_fit = {}
            orig_fit = inference.fit

            def timed_fit(*args, **kw):
                import torch
                torch.cuda.synchronize()
                t0 = time.time()
                m = orig_fit(*args, **kw)
                t_fit['s'] = time.time() - t0
                return m
            inference.fit = timed_fit
            try:
                with contextlib.redirect_stdout(io.StringIO()):
                    t0 = time.time()
                    m = inference.main(base + ['--out_dir', out] + extra)
                    total = time.time() - t0
            finally:
                inference.fit = orig_fit
            scored = sum(1 for f in m['fitscore'] if f and f[0] != -2.0)
            if r > 0 and (best is None or t_fit['s'] < best[0]):
                best = (t_fit['s'], total, float(np.mean([max(f) for f in m['fitscore']])))
        res[tag] = {'fit_s': round(best[0], 3), 'main_s': round(best[1], 3),
                    'samples_per_s_incl_output_and_scoring': round(n_pairs * a.samples / best[0], 1),
                    'pairs_scored_by_ancphore': scored, 'mean_max_fitscore': round(best[2], 4)}
    res.update(workload=f'{n_pairs} example ligands x sQC_QFA_complex.phore (P=79) x {a.samples} samples x 20 steps, shipped checkpoint',
               reference_shipped_run_time_s=199.4, reference_shipped_mean_max_fitscore=0.449,
               note='fit = denoising + SD files + AncPhore scoring + ranked SD files; reference_shipped_run_time_s = sum of run_time '
                    'in the reference\'s examples/output/2/inference_results.json (denoising only, its GPU unstated)')
    print(json.dumps(res))


if __name__ == '__main__':
    main()
