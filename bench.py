#!/usr/bin/env python
"""Benchmark of the DiffPhore denoising hot path on B200 (contract: see the task statement / DESIGN.md §Measurement).

One bench "step" = one full reverse-diffusion pass (20 denoising steps: score model + conformer update) over the
BASELINE.json configs[1] workload: 256 synthetic ligand–pharmacophore pairs (32 atoms / 8 pharmacophore points)
x 40 samples per pair = 10 240 poses per GPU.  metric = denoised samples/s.

  value : device-timed (CUDA events), graphs already packed and resident in HBM when the timed region starts
  e2e   : the same work through the public API (DenoisingSampler.run on host graphs): host packing, H2D of the
          packed batch, 20 steps, D2H of the final poses, all inside the timed region
  --impl reference : the oracle (CPU restatement of the reference's op graph) on the host cores, bounded sample
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'src'))

N_PAIRS, N_ATOMS, N_PHORE, SAMPLES, INF_STEPS = 256, 32, 8, 40, 20
WORKLOAD = 'synthetic 256 pairs (32 atoms / 8 phore points) x 40 samples x 20 denoising steps per GPU'


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        return float(p['hbm_gbs']), 'measured'
    except Exception:
        return 6650.0, 'fallback'


def ncu_traffic(tag):
    """dram read + write bytes of the first `tag` launch in the committed ncu --set full summary (None if absent)."""
    try:
        import csv
        rows = list(csv.reader(open(os.path.join(ROOT, 'profiles', 'ncu_conv_fused_r2.csv'))))
        h, units = rows[0], rows[1]
        ir, iw = h.index('dram__bytes_read.sum'), h.index('dram__bytes_write.sum')
        scale = {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1.0}
        for r in rows[2:]:
            if tag in r[0]:
                return float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]
    except Exception:
        pass
    return None


def tensor_peak():
    """Sustained dense bf16 cuBLAS throughput (the conv kernel is timed inside a seconds-long step)."""
    try:
        p = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        return float(p['bf16_tflops_sustained']), 'measured'
    except Exception:
        return 1400.0, 'fallback'


class ClockSampler:
    FIELDS = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,' \
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.FIELDS}', '--format=csv,noheader,nounits',
                                          '-lms', '200', '-i', str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace('.', '').isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace('.', '').isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[5:9]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def cpu_reference_rate(n_pairs, samples, threads, steps=INF_STEPS, first_pair=0, forward_only=False, real=False):
    """Oracle = reference op graph on CPU (materialised per-edge TP weights, index_add scatter, per-step re-collation,
    per-sample scipy/LAPACK conformer update); batch_size = min(samples, 40): the reference batches the copies of ONE pair,
    `sample_per_complex` = 40 at a time (sampling.py:210, inference.py:184).  forward_only: the model calls alone (same
    batches, same count) without noise / conformer update.  real: the cfg1 pair (STK936575 x the 79-node example pharmacophore).
    Returns (samples/s, seconds)."""
    from tests.parity_util import make_draws, oracle_initial_graphs
    from diffphore_b200.synthetic import random_state_dict, make_pairs, real_example_pairs
    from diffphore_b200.graph import collate
    from oracle.model import OracleScoreModel, default_config
    from oracle import sampler as osamp
    from oracle.tables import So3ScoreNorm, TorusScoreNorm
    torch.set_num_threads(threads)
    sd = random_state_dict(0)
    graphs = [real_example_pairs(12)[11]] if real else make_pairs(n_pairs, N_ATOMS, N_PHORE, first=first_pair)
    init, noise, n_rot = make_draws(graphs, samples, 0, steps=steps)
    dl = oracle_initial_graphs(graphs, samples, init, n_rot)
    so3n, torn = So3ScoreNorm(), TorusScoreNorm()
    om = OracleScoreModel(sd, so3n, torn)
    for t in osamp.get_t_schedule(steps):                       # table rows are a one-off import cost in the reference
        s = np.asarray([0.1 ** (1 - t) * 1.5 ** t], dtype=np.float32)
        so3n(s)
        torn(np.asarray([0.0314 ** (1 - t) * 3.14 ** t], dtype=np.float32))
    bs = min(samples, 40)
    t0 = time.time()
    if forward_only:
        for t in osamp.get_t_schedule(steps):
            for k in range(0, len(dl), bs):
                batch = collate(dl[k:k + bs])
                osamp.set_time(batch, float(t), batch.num_graphs)
                with torch.no_grad():
                    om(batch)
    else:
        osamp.sampling(dl, om, steps, default_config(), collate, batch_size=bs, noise=noise)
    dt = time.time() - t0
    return len(dl) / dt, dt


REF_PAIRS_PER_STEP = 1                                          # x 40 samples x 20 steps at batch 40: ~10-25 s of CPU work per bench step


def cpu_baseline_block(cores):
    """cpu_baseline of the main arm: the oracle port on `cores` threads, bounded samples of the cfg2 workload at the reference's
    batching (40 copies of one pair per batch), full step and forward only, plus the cfg1 job."""
    r, dt = cpu_reference_rate(REF_PAIRS_PER_STEP, SAMPLES, cores)
    rf, dtf = cpu_reference_rate(1, SAMPLES, cores, forward_only=True)
    r1, dt1 = cpu_reference_rate(1, 4, cores, real=True)
    return {'value': r, 'unit': 'samples/s', 'cores': cores, 'kind': 'port',
            'sample': f'{REF_PAIRS_PER_STEP} pairs x {SAMPLES} samples x {INF_STEPS} steps of the cfg2 shape at batch {SAMPLES} ({dt:.1f} s of CPU work, '
                      f'oracle port, torch.set_num_threads({cores})); extrapolates linearly in pairs',
            'forward_only': {'value': rf, 'unit': 'samples/s', 'sample': f'1 pair x {SAMPLES} samples x {INF_STEPS} model calls at batch {SAMPLES} ({dtf:.1f} s)'},
            'cfg1': {'value': r1, 'unit': 'samples/s', 'seconds': dt1,
                     'sample': 'STK936575 x sQC pharmacophore (N=20, P=79), 4 samples x 20 steps at batch 4, random-init weights'}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--pairs', type=int, default=N_PAIRS)
    ap.add_argument('--samples', type=int, default=SAMPLES)
    ap.add_argument('--atoms', type=int, default=N_ATOMS, help='ligand atoms per synthetic pair (default: cfg2 = 32)')
    ap.add_argument('--phore', type=int, default=N_PHORE, help='pharmacophore points per synthetic pair (default: cfg2 = 8)')
    ap.add_argument('--real', type=int, default=0, help='NOT the headline: the first N real-shaped pairs of tests/golden/real_pairs.npz '
                                                        '(reference example ligands x the 79-node example pharmacophore), shipped checkpoint if present')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-extra', action='store_true', help='skip the non-headline cfg3 / cfg4 / cfg5 lines')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    cores = os.cpu_count()

    if args.impl == 'reference':
        if rank != 0:
            return
        vals = []
        for _ in range(args.warmup if args.warmup < 2 else 1):
            cpu_reference_rate(1, 2, cores, steps=2)
        t0 = time.time()
        for k in range(args.steps):                              # every bench step denoises the NEXT pairs of the cfg2 workload
            r, dt = cpu_reference_rate(REF_PAIRS_PER_STEP, SAMPLES, cores, first_pair=(k * REF_PAIRS_PER_STEP) % N_PAIRS)
            vals.append(r)
        v = float(np.mean(vals))
        sample = (f'{REF_PAIRS_PER_STEP} pairs x {SAMPLES} samples x {INF_STEPS} steps of the cfg2 shape per bench step at batch {SAMPLES} '
                  f'(the reference batches the {SAMPLES} copies of one pair, sampling.py:210); {args.steps * REF_PAIRS_PER_STEP} distinct pairs over the run; '
                  'oracle port of the reference CPU path (e3nn / PyG / torch_cluster are not installable here)')
        print(json.dumps({'metric': 'denoised samples/sec (20-step)', 'value': v, 'unit': 'samples/s', 'n_gpus': args.gpus,
                          'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1000 * (time.time() - t0) / args.steps,
                          'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
                          'data': 'synthetic graphs, random-init weights of the shipped architecture',
                          'impl': 'reference', 'config': {'workload': WORKLOAD},
                          'cpu_baseline': {'value': v, 'unit': 'samples/s', 'cores': cores, 'kind': 'port', 'sample': sample},
                          'e2e': {'value': v, 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))
        return

    import torch.distributed as dist
    from diffphore_b200.engine import ModelWeights
    from diffphore_b200.sampler import DenoisingSampler
    from diffphore_b200.synthetic import make_pairs
    from diffphore_b200 import profiling
    from diffphore_b200.synthetic import random_state_dict, real_example_pairs
    import __graft_entry__ as ge
    ge.build()
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    sd = random_state_dict(0)
    graphs = make_pairs(args.pairs, args.atoms, args.phore, first=rank * args.pairs)
    ckpt = os.path.join(ROOT, 'oracle', '_ref', 'weights', 'best_ema_inference_epoch_model.pt')
    shipped_sd = torch.load(ckpt, map_location='cpu', weights_only=False) if os.path.exists(ckpt) else None
    if args.real:
        graphs = real_example_pairs(args.real)
        args.pairs = len(graphs)
        sd = shipped_sd if shipped_sd is not None else sd
    w = ModelWeights(sd, dev)
    sampler = DenoisingSampler(w, INF_STEPS)
    n_samples_local = args.pairs * args.samples
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def gather_poses(pos_dev):
        if world > 1:
            out = [torch.empty_like(pos_dev) for _ in range(world)]
            dist.all_gather(out, pos_dev)                       # the single NCCL collective of the path
            return out
        return [pos_dev]

    # ---- device-resident arm: pack once per bench step outside the timed region
    resident = sampler.prepare(graphs, args.samples)
    for _ in range(args.warmup):
        sampler.reset(resident, generator=gen)
        sampler.run_resident(resident, generator=gen)
        gather_poses(torch.cat([r[0].pos for r in resident]))
    barrier()
    prof = profiling.KernelTimer()
    clocks = ClockSampler(local)
    clocks.start()
    l0 = sampler.gpu_launches
    times = []
    for k_step in range(args.steps):
        sampler.reset(resident, generator=gen)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        # per-launch CUDA events (roofline, `kernels`) bracket every conv launch of the LAST timed step only: two events per
        # launch cost ~2.5 % of a step (799 vs 778 ms measured), which is measurement overhead, not work of the path
        sampler.run_resident(resident, generator=gen, timer=prof if k_step == args.steps - 1 else None)
        gather_poses(torch.cat([r[0].pos for r in resident]))
        e1.record()
        barrier()
        times.append(e0.elapsed_time(e1))
    clk = clocks.stop()
    launches = (sampler.gpu_launches - l0) // max(args.steps, 1)
    ms = float(np.mean(times))
    tmax = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item())
    value = n_samples_local * world / (ms / 1000.0)
    roof = prof.roofline_fused(*tensor_peak(), peaks()[0], traffic_bytes=ncu_traffic('TpL3')) or prof.roofline(*peaks())

    # ---- end-to-end arm through the public API with host inputs
    e2e = None
    if not args.no_e2e:
        # warm-up of the API path at full size (untimed): the device-resident arm's buffers are released first, so the
        # caching allocator serves the timed runs from blocks of the right sizes; its phases go to stderr for the record
        del resident
        barrier()
        t0 = time.perf_counter()
        res_w = sampler.prepare(graphs, args.samples)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        sampler.reset(res_w, generator=gen)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        sampler.run_resident(res_w, generator=gen)
        torch.cuda.synchronize(); t3 = time.perf_counter()
        if rank == 0:
            print(f'[bench] e2e phases (warm-up run): pack+H2D+setup {1e3 * (t1 - t0):.1f} ms, initial poses {1e3 * (t2 - t1):.1f} ms, '
                  f'20 steps {1e3 * (t3 - t2):.1f} ms', file=sys.stderr)
        del res_w
        # ... and two untimed calls of the API itself: run() uploads on its own copy stream, whose allocator pool is separate
        # (a first timed call that still grows that pool showed up as a +70 ms outlier in one of two processes)
        for _ in range(2):
            sampler.run(graphs, args.samples, generator=gen, pinned=True)
        barrier()
        t_e2e = []
        for _ in range(max(1, min(args.steps, 3))):
            barrier()
            t0 = time.perf_counter()
            pos, ptr = sampler.run(graphs, args.samples, generator=gen, pinned=True)
            gather_poses(pos.to(dev, non_blocking=True)) if world > 1 else None
            barrier()
            t_e2e.append(time.perf_counter() - t0)
        if rank == 0:
            print('[bench] e2e runs (s): ' + ', '.join(f'{t:.4f}' for t in t_e2e), file=sys.stderr)
        te = torch.tensor([float(np.mean(t_e2e))], device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {'value': n_samples_local * world / float(te.item()), 'unit': 'samples/s',
               'h2d_bytes_per_step': int(sampler.last_h2d_bytes), 'd2h_bytes_per_step': int(pos.numel() * 4)}

    # ---- the other BASELINE.json configurations, NOT the headline: same sampler API, one warm-up + two timed runs each
    def side_config(graphs_x, sd_x, samples_x, what):
        sm = sampler if sd_x is sd else DenoisingSampler(ModelWeights(sd_x, dev), INF_STEPS)
        res_x = sm.prepare(graphs_x, samples_x)
        n_x = len(graphs_x) * samples_x
        tt = []
        for k in range(3):
            sm.reset(res_x, generator=gen)
            barrier()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            sm.run_resident(res_x, generator=gen)
            gather_poses(torch.cat([r[0].pos for r in res_x]))
            a1.record()
            barrier()
            if k:
                tt.append(a0.elapsed_time(a1))
        del res_x
        te_x = []
        for k in range(3):
            barrier()
            t0 = time.perf_counter()
            pos_x, _ = sm.run(graphs_x, samples_x, generator=gen, pinned=True)
            gather_poses(pos_x.to(dev, non_blocking=True)) if world > 1 else None
            barrier()
            if k:
                te_x.append(time.perf_counter() - t0)
        tm = torch.tensor([float(np.mean(tt)), 1e3 * float(np.mean(te_x))], device=dev)
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        return {'workload': what, 'value': n_x * world / (float(tm[0]) / 1e3), 'e2e': n_x * world / (float(tm[1]) / 1e3),
                'unit': 'samples/s', 'ms_per_job': float(tm[0]), 'e2e_ms_per_job': float(tm[1]), 'samples_per_gpu': n_x}

    extra = None
    if not args.no_extra and not args.real and (args.pairs, args.samples, args.atoms, args.phore) == (N_PAIRS, SAMPLES, N_ATOMS, N_PHORE):
        extra = {}
        g3 = real_example_pairs(15)
        extra['cfg3_shape'] = side_config(g3, shipped_sd if shipped_sd is not None else sd, SAMPLES,
                                          f'{len(g3)} real-shaped pairs (reference example ligands x the 79-node example pharmacophore) x {SAMPLES} samples, '
                                          + ('shipped checkpoint' if shipped_sd is not None else 'random-init weights') + ', every rank runs the same job')
        extra['cfg4'] = side_config(make_pairs(512, 64, 12, first=rank * 512), sd, SAMPLES,
                                    f'synthetic 512 pairs per GPU (64 atoms / 12 phore points) x {SAMPLES} samples: 4096 pairs over 8 GPUs')
        extra['cfg5'] = side_config(make_pairs(1, 128, 16), sd, SAMPLES,
                                    f'1 pair (128 atoms / 16 phore points) x {SAMPLES} samples: latency-bound; every rank runs the same job')

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_block(cores)
    if rank == 0:
        print(json.dumps({'metric': 'denoised samples/sec (20-step)', 'value': value, 'unit': 'samples/s', 'n_gpus': world,
                          'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True,
                          'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
                          'data': 'synthetic graphs (diffphore_b200/synthetic.py), random-init weights of the shipped architecture',
                          'config': {'workload': WORKLOAD if (args.pairs, args.samples, args.atoms, args.phore) == (N_PAIRS, SAMPLES, N_ATOMS, N_PHORE)
                                     else (f'{args.pairs} real-shaped pairs (reference example ligands x 79-node pharmacophore) x {args.samples} samples x '
                                           f'{INF_STEPS} steps (NOT the headline configuration)' if args.real else
                                           f'synthetic {args.pairs} pairs ({args.atoms} atoms / {args.phore} phore points) x {args.samples} samples x '
                                           f'{INF_STEPS} denoising steps per GPU (NOT the headline configuration)'),
                                     'pairs_per_gpu': args.pairs, 'samples_per_pair': args.samples,
                                     'denoising_steps': INF_STEPS, 'l2': 'inputs larger than L2 (per conv: >= 130 MB of node features + 0.1-0.7 GB of per-edge hidden activations)',
                                     'parallelism': f'pairs sharded over {world} rank(s), one all_gather of poses'},
                          'clocks': clk, 'e2e': e2e, 'gpu_launches': int(launches), 'roofline': roof,
                          'kernels': prof.summary(), 'cpu_baseline': cpu, 'other_configs_not_headline': extra}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
