"""Reverse-diffusion driver on the GPU: the 20-step loop of sampling_phore with all samples of all pairs in flight.

Reference: /root/reference/src/utils/sampling.py:174-255 (sampling_phore), :16-63 (randomize_position),
/root/reference/src/utils/diffusion_utils.py:135-145 (get_t_schedule).  Differences by design:
the reference re-collates the N copies of ONE pair every step and updates conformers one sample at a time on the
host; here every (pair, sample) graph of a chunk stays resident in HBM for all steps, one kernel sequence per step.
"""
import math

import numpy as np
import torch

from .engine import Engine, ModelWeights
from .tables import So3ScoreNorm, TorusScoreNorm


def get_t_schedule(inference_steps):
    return np.linspace(1, 0, inference_steps + 1)[:-1]


def random_rotations(n, generator=None, device='cpu'):
    """Uniform random rotation matrices [n,3,3] (same law as scipy Rotation.random(): normalised Gaussian quaternion)."""
    q = torch.randn(n, 4, generator=generator, device=device)
    q = q / q.norm(dim=1, keepdim=True)
    r, i, j, k = q.unbind(1)
    return torch.stack([1 - 2 * (j * j + k * k), 2 * (i * j - k * r), 2 * (i * k + j * r),
                        2 * (i * j + k * r), 1 - 2 * (i * i + k * k), 2 * (j * k - i * r),
                        2 * (i * k - j * r), 2 * (j * k + i * r), 1 - 2 * (i * i + j * j)], 1).reshape(n, 3, 3)


class DenoisingSampler:
    def __init__(self, weights: ModelWeights, inference_steps=20, so3_norm=None, torus_norm=None,
                 weight_buffer_bytes=24 << 30, no_final_step_noise=False, resident_bytes=48 << 30,
                 cuda_graphs=True, graph_max_graphs=2048, ode=False, pipeline_graphs=None, pipeline_head_graphs=640):
        self.w = weights
        self.engine = Engine(weights)
        self.steps = inference_steps
        self.so3 = so3_norm or So3ScoreNorm()
        self.torus = torus_norm or TorusScoreNorm()
        self.weight_buffer_bytes = weight_buffer_bytes
        self.resident_bytes = resident_bytes
        # run(): optional cap on the graphs (pair x sample) per chunk of a one-shot job.  Jobs that need several chunks anyway are
        # pipelined (chunk k+1 packed / uploaded under chunk k's kernels); cutting a job that fits one chunk just to pipeline it
        # does not pay: measured on cfg2 (10240 graphs), 3 chunks hide 19 of the 29 ms of packing but add 1720 launches of
        # ~15-20 us fixed cost each (0.759 s vs 0.756 s per job) - hence None.
        self.pipeline_graphs = pipeline_graphs
        # ... but a SHORT head chunk does: 640 graphs (16 cfg2 pairs, 2 ms of packing) keep the GPU busy for ~50 ms while the host
        # packs the bulk of the job; one extra set of 860 launches.  Measured on cfg2: 0.751 -> 0.744 s per job (+1 %); 320 / 1280
        # graphs: 0.745 / 0.754 s.  0 disables it.
        self.pipeline_head_graphs = pipeline_head_graphs
        # resident chunks that are denoised repeatedly replay their whole loop as one captured CUDA graph (see _loop_graph)
        self.cuda_graphs, self.graph_max_graphs = cuda_graphs, graph_max_graphs
        self.no_final_step_noise = no_final_step_noise
        self.ode = ode                                   # --ode: 0.5 g^2 dt score, no noise (sampling.py:226-228)
        sched = get_t_schedule(inference_steps)
        rows = []
        for k in range(inference_steps):
            dt = sched[k] - sched[k + 1] if k < inference_steps - 1 else sched[k]          # sampling.py:206-208
            rows.append(weights.step_consts(float(sched[k]), self.so3, self.torus, dt=float(dt), ode=ode))
        self.sched = sched
        self.consts = torch.stack(rows).to(weights.device)                                 # [steps, 256]
        self.gpu_launches = 0
        self.last_h2d_bytes = 0

    # ------------------------------------------------------------------
    def graphs_per_chunk(self, graphs, samples):
        """Pairs per resident chunk.  Unfused kernels: the worst-case per-edge weight scratch must fit `weight_buffer_bytes`.
        Fused kernels keep the weights on chip, so only node / edge features count (~0.2 MB per cfg2 graph): cap the chunk
        at `max_graphs_per_chunk` graphs (pair x sample) to bound the resident working set."""
        k = self.w.cfg['max_neighbors']
        if self.engine.use_fused:
            worst = 1
            for g in graphs:
                n, P = g['ligand'].pos.shape[0], g['phore'].pos.shape[0]
                eb = g['ligand', 'ligand'].edge_index.shape[1]
                e_all = eb + n * min(n - 1, k + 1) + 2 * n * P + g['phore', 'phore'].edge_index.shape[1]
                worst = max(worst, e_all * 4 * 60 + (n + P) * 4 * 600)      # edge embeddings / SH / indices + node features
                if n > 256 or P > 256:                                      # a cross node with > 256 edges: unfused scratch too
                    worst = max(worst, n * P * 2200 * 4)
            return max(1, int(self.resident_bytes // (worst * samples)))
        worst = 1
        for g in graphs:
            n, P = g['ligand'].pos.shape[0], g['phore'].pos.shape[0]
            eb = g['ligand', 'ligand'].edge_index.shape[1]
            worst = max(worst, (eb + n * min(n - 1, k + 1)) * 2200, n * P * 2200)
        per_pair = worst * 4 * samples
        return max(1, int(self.weight_buffer_bytes // per_pair))

    # ------------------------------------------------------------------ device-resident API
    def prepare(self, graphs, samples_per_graph=1):
        """Pack all chunks onto the device once.  Returns [(PackedBatch, Workspace, pos0, norm0), ...]."""
        chunk = self.graphs_per_chunk(graphs, samples_per_graph)
        res, self.last_h2d_bytes = [], 0
        wbuf = None
        for c0 in range(0, len(graphs), chunk):
            b, ws = self.engine.pack(graphs[c0:c0 + chunk], samples_per_graph, wbuf)
            wbuf = ws.wbuf if wbuf is None or ws.wbuf.numel() > wbuf.numel() else wbuf
            self.last_h2d_bytes += b.h2d_bytes
            res.append((b, ws, b.pos.clone(), b.norm.clone()))
        return res

    def reset(self, resident, generator=None, init=None, randomize=True, no_torsion=False):
        """Restore the input poses and draw new initial poses (randomize_position, sampling.py:16-63)."""
        dev = self.w.device
        g_off = r_off = 0
        for b, ws, pos0, norm0 in resident:
            b.pos.copy_(pos0)
            b.norm.copy_(norm0)
            if randomize:
                if init is None:
                    tor0 = (torch.rand(max(b.n_rot, 1), generator=generator, device=dev) * 2 - 1) * math.pi
                    rot0 = random_rotations(b.B, generator, dev)
                    tr0 = torch.randn(b.B, 3, generator=generator, device=dev) * self.w.cfg['tr_sigma_max']
                else:
                    tor0 = torch.as_tensor(init['tor'][r_off:r_off + b.n_rot], dtype=torch.float32).to(dev)
                    rot0 = torch.as_tensor(init['rot'][g_off:g_off + b.B], dtype=torch.float32).to(dev)
                    tr0 = torch.as_tensor(init['tr'][g_off:g_off + b.B], dtype=torch.float32).to(dev)
                self.engine.randomize(b, tor0.contiguous(), rot0.reshape(-1, 9).contiguous(), tr0.contiguous(), no_torsion)
            g_off += b.B
            r_off += b.n_rot

    def _noise_buffers(self, b, ws):
        """Static per-chunk noise buffers for ALL steps ([steps, B, 3], [steps, B, 3], [steps, n_rot]): filled once per job before the
        loop (sampling.py:230-244 draws them step by step on the host), so that the loop itself launches nothing but our kernels."""
        if not hasattr(ws, 'z_all'):
            dev = self.w.device
            ws.z_all = (torch.zeros(self.steps, b.B, 3, device=dev), torch.zeros(self.steps, b.B, 3, device=dev),
                        torch.zeros(self.steps, max(b.n_rot, 1), device=dev))
        return ws.z_all

    def _step_args(self, ws, k, noise_mode):
        """(constants, tr_z, rot_z, tor_z) of step k: views into static tensors (CUDA-graph capturable)."""
        last = k == self.steps - 1
        if noise_mode == 'none' or (noise_mode == 'all_but_last' and last):
            return self.consts[k], None, None, None
        return self.consts[k], ws.z_all[0][k], ws.z_all[1][k], ws.z_all[2][k]

    def _loop_graph(self, b, ws, no_torsion, noise_mode):
        """The WHOLE denoising loop of a chunk (steps x (score model + conformer update), 43 launches each) captured as ONE CUDA
        graph.  Every kernel argument is a pointer into the chunk's static buffers: step k reads its constants from consts[k] and
        its noise from z_all[:, k]; dynamic edge and tile counts are read on the device.  Launch-bound small jobs (cfg1 / cfg5
        shapes: 4-40 graphs) spend ~0.5 ms per step in Python + launch overhead otherwise; between two replays nothing else runs."""
        key = (bool(no_torsion), noise_mode)
        cache = ws.__dict__.setdefault('_graphs', {})
        if key in cache:
            return cache[key]
        self._noise_buffers(b, ws)
        n_before = ws.n_launches                       # the warm-up step and the capture pass are not part of the job
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                  # one eager step: lazy per-kernel initialisation happens outside the capture
            pos0, norm0 = b.pos.clone(), b.norm.clone()
            sc, trz, rotz, torz = self._step_args(ws, 0, noise_mode)
            self.engine.forward(b, ws, sc)
            self.engine.update(b, ws, sc, trz, rotz, torz, no_torsion=no_torsion)
            b.pos.copy_(pos0); b.norm.copy_(norm0)
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        n0 = ws.n_launches
        with torch.cuda.graph(g):
            for k in range(self.steps):
                sc, trz, rotz, torz = self._step_args(ws, k, noise_mode)
                self.engine.forward(b, ws, sc)
                self.engine.update(b, ws, sc, trz, rotz, torz, no_torsion=no_torsion)
        cache[key] = (g, ws.n_launches - n0)
        ws.n_launches = n_before
        b.pos.copy_(pos0); b.norm.copy_(norm0)         # (capture does not execute, but keep the pose exactly as it was)
        return cache[key]

    def run_resident(self, resident, noise=None, no_random=False, generator=None, trace=None, no_torsion=False, timer=None,
                     pose_trace=None, whole_loop=True):
        """The 20-step loop (sampling.py:204-255) over device-resident chunks; poses end up in each chunk's b.pos.
        pose_trace: optional list; receives per chunk a device tensor [steps + 1, n_lig, 3] = the initial pose and the pose
        after every step (`keep_update`: initial_poses / docked_poses of inference.py:191-192, diffusion_utils.py:71-77).
        whole_loop: chunks of <= graph_max_graphs graphs replay the whole loop as one CUDA graph (resident chunks that are
        denoised again and again); False: eager launches - for a one-shot job capturing costs 50-180 ms (measured, 43-860 kernel
        nodes) while the eager loop of a small job is GPU-bound anyway (cfg5 shape: 45.9 ms eager = 45.9 ms replayed)."""
        self.engine.timer = timer
        g_off = r_off = 0
        for b, ws, _, _ in resident:
            n0 = ws.n_launches
            sl_g, sl_r = slice(g_off, g_off + b.B), slice(r_off, r_off + b.n_rot)
            noise_mode = 'none' if (no_random or self.ode) else ('all_but_last' if self.no_final_step_noise else 'all')
            # ---- all Gaussian draws of the job up front (device RNG, or the injected draws): tr, rot, tor per step
            if noise_mode != 'none':
                z_tr, z_rot, z_tor = self._noise_buffers(b, ws)
                if noise is None:
                    z_tr.normal_(generator=generator)
                    z_rot.normal_(generator=generator)
                    z_tor.normal_(generator=generator)
                else:
                    for dst, key, sl in ((z_tr, 'tr', sl_g), (z_rot, 'rot', sl_g), (z_tor, 'tor', sl_r)):
                        host = np.stack([np.asarray(noise[k][key], dtype=np.float32)[sl] for k in range(self.steps)])
                        if host.size:
                            dst[:, :host.shape[1]].copy_(torch.from_numpy(host))      # (the torsion buffer keeps one slot when n_rot = 0)
            use_graph = (self.cuda_graphs and timer is None and trace is None and pose_trace is None
                         and b.B <= self.graph_max_graphs)
            self.engine.concurrent = b.B <= self.graph_max_graphs        # latency-bound chunks: two streams inside the score model
            if use_graph and whole_loop:
                graph, n_l = self._loop_graph(b, ws, no_torsion, noise_mode)
                graph.replay()
                ws.n_launches += n_l
            else:
                traj = [b.pos.clone()] if pose_trace is not None else None
                for k in range(self.steps):
                    sc, trz, rotz, torz = self._step_args(ws, k, noise_mode)
                    self.engine.forward(b, ws, sc)
                    if trace is not None:
                        trace.append((ws.tr.clone().cpu(), ws.rot.clone().cpu(), ws.tor[:b.n_rot].clone().cpu()))
                    self.engine.update(b, ws, sc, trz, rotz, torz, no_torsion=no_torsion)
                    if traj is not None:
                        traj.append(b.pos.clone())
                if traj is not None:
                    pose_trace.append(torch.stack(traj))
            self.gpu_launches += ws.n_launches - n0
            g_off += b.B
            r_off += b.n_rot
        self.engine.timer = None
        self.engine.concurrent = False

    def _run_pipelined(self, graphs, samples_per_graph, bounds, no_random, generator, randomize, no_torsion, pinned):
        """run() for jobs of several chunks with device-side draws: chunk k+1 is packed on the host and uploaded on a copy
        stream while the GPU denoises chunk k (the 860 launches of a chunk take ~10 ms of host time, its kernels ~250 ms), so
        only the first chunk's packing is exposed.  Same kernels and per-chunk results as the plain path; the random streams
        are consumed chunk by chunk (initial poses, then noise) instead of all initial poses first."""
        if getattr(self, '_copy_stream', None) is None:
            self._copy_stream = torch.cuda.Stream(device=self.w.device)
        compute = torch.cuda.current_stream()
        self.last_h2d_bytes, wbuf = 0, None

        def pack(i):
            nonlocal wbuf
            with torch.cuda.stream(self._copy_stream):
                b, ws = self.engine.pack(graphs[bounds[i]:bounds[i + 1]], samples_per_graph, wbuf)
                wbuf = ws.wbuf if wbuf is None or ws.wbuf.numel() > wbuf.numel() else wbuf
                item = (b, ws, b.pos.clone(), b.norm.clone())
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
            self.last_h2d_bytes += b.h2d_bytes
            return item, ev

        resident, nxt, n_chunks = [], pack(0), len(bounds) - 1
        for i in range(n_chunks):
            item, ev = nxt
            compute.wait_event(ev)
            self.reset([item], generator=generator, randomize=randomize, no_torsion=no_torsion)
            self.run_resident([item], no_random=no_random, generator=generator, no_torsion=no_torsion, whole_loop=False)
            resident.append(item)
            if i + 1 < n_chunks:
                nxt = pack(i + 1)                           # host work and H2D of the next chunk, under this chunk's kernels
        self.last_trajectory = None
        n_tot = sum(b.n_lig for b, _, _, _ in resident)
        out = torch.empty(n_tot, 3, dtype=torch.float32, pin_memory=pinned)
        ptr, o = [0], 0
        for b, ws, _, _ in resident:
            out[o:o + b.n_lig].copy_(b.pos, non_blocking=pinned)
            o += b.n_lig
            ptr += (np.cumsum(b.n_per) + ptr[-1]).tolist()
        torch.cuda.synchronize()
        return out, np.asarray(ptr)

    # ------------------------------------------------------------------ host-facing API
    def run(self, graphs, samples_per_graph=1, noise=None, init=None, no_random=False, generator=None,
            randomize=True, trace=None, no_torsion=False, pinned=False, keep_update=False, start_poses=None):
        """Denoise `samples_per_graph` poses for every pair in `graphs` (host graphs in, host poses out).

        init : None (device RNG) or dict(tor=[sum n_rot] , rot=[B,3,3], tr=[B,3]) in graph order (pair-major).
        noise: None (device RNG, or zeros when no_random) or list over steps of dict(tr=[B,3], rot=[B,3], tor=[n_rot]).
        start_poses: optional (pos, norm) per SAMPLE replacing the pairs' input poses (graph order, pair-major).
        keep_update: also keep the pose after every step; `self.last_trajectory` = CPU tensor [steps + 1, n_lig_total, 3].
        Returns (pos [n_lig_total,3] float32 CPU tensor, lig_ptr numpy [B+1])."""
        chunk = self.graphs_per_chunk(graphs, samples_per_graph)
        if start_poses is None and init is None and noise is None and trace is None and not keep_update:
            # one-shot job with device-side draws: when it takes several chunks, packing / upload of chunk k+1 overlaps the kernels
            # of chunk k
            if self.pipeline_graphs:
                n_chunks = max(-(-len(graphs) // chunk), -(-len(graphs) * samples_per_graph // self.pipeline_graphs))
                chunk = -(-len(graphs) // n_chunks)
            bounds = list(range(0, len(graphs), chunk)) + [len(graphs)]
            # a short head chunk gets the GPU going while the host packs the bulk of the job
            head = self.pipeline_head_graphs // max(1, samples_per_graph)
            if head and bounds[1] > 8 * head:
                bounds.insert(1, head)
            if len(bounds) > 2:
                return self._run_pipelined(graphs, samples_per_graph, bounds, no_random, generator, randomize, no_torsion, pinned)
        resident = self.prepare(graphs, samples_per_graph)
        if start_poses is not None:
            # per-SAMPLE start poses (pos [n_lig_total, 3], norm [n_lig_total, 33] in graph order): the callers of the reference
            # API hand over deep copies of a pair that may already carry different poses
            o = 0
            for b, ws, pos0, norm0 in resident:
                pos0.copy_(torch.as_tensor(start_poses[0][o:o + b.n_lig], dtype=torch.float32))
                norm0.copy_(torch.as_tensor(start_poses[1][o:o + b.n_lig], dtype=torch.float32).reshape(b.n_lig, 33))
                o += b.n_lig
        self.reset(resident, generator=generator, init=init, randomize=randomize, no_torsion=no_torsion)
        pose_trace = [] if keep_update else None
        self.run_resident(resident, noise=noise, no_random=no_random, generator=generator, trace=trace, no_torsion=no_torsion,
                          pose_trace=pose_trace, whole_loop=False)
        self.last_trajectory = torch.cat(pose_trace, dim=1).cpu() if keep_update else None
        n_tot = sum(b.n_lig for b, _, _, _ in resident)
        out = torch.empty(n_tot, 3, dtype=torch.float32, pin_memory=pinned)
        ptr, o = [0], 0
        for b, ws, _, _ in resident:
            out[o:o + b.n_lig].copy_(b.pos, non_blocking=pinned)
            o += b.n_lig
            ptr += (np.cumsum(b.n_per) + ptr[-1]).tolist()
        torch.cuda.synchronize()
        return out, np.asarray(ptr)
