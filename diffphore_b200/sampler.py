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
                 weight_buffer_bytes=24 << 30, no_final_step_noise=False):
        self.w = weights
        self.engine = Engine(weights)
        self.steps = inference_steps
        self.so3 = so3_norm or So3ScoreNorm()
        self.torus = torus_norm or TorusScoreNorm()
        self.weight_buffer_bytes = weight_buffer_bytes
        self.no_final_step_noise = no_final_step_noise
        sched = get_t_schedule(inference_steps)
        rows = []
        for k in range(inference_steps):
            dt = sched[k] - sched[k + 1] if k < inference_steps - 1 else sched[k]          # sampling.py:206-208
            rows.append(weights.step_consts(float(sched[k]), self.so3, self.torus, dt=float(dt)))
        self.sched = sched
        self.consts = torch.stack(rows).to(weights.device)                                 # [steps, 256]
        self.gpu_launches = 0

    # ------------------------------------------------------------------
    def graphs_per_chunk(self, graphs, samples):
        k = self.w.cfg['max_neighbors']
        worst = 1
        for g in graphs:
            n, P = g['ligand'].pos.shape[0], g['phore'].pos.shape[0]
            eb = g['ligand', 'ligand'].edge_index.shape[1]
            worst = max(worst, (eb + n * min(n - 1, k + 1)) * 2200, n * P * 2200)
        per_pair = worst * 4 * samples
        return max(1, int(self.weight_buffer_bytes // per_pair))

    def run(self, graphs, samples_per_graph=1, noise=None, init=None, no_random=False, generator=None,
            randomize=True, trace=None, no_torsion=False):
        """Denoise `samples_per_graph` poses for every pair in `graphs`.

        init : None (device RNG) or dict(tor=[sum n_rot] , rot=[B,3,3], tr=[B,3]) in graph order (pair-major).
        noise: None (device RNG, or zeros when no_random) or list over steps of dict(tr=[B,3], rot=[B,3], tor=[n_rot]).
        Returns (pos [n_lig_total,3] float32 CPU tensor, lig_ptr numpy [B+1])."""
        dev = self.w.device
        out_pos, out_ptr = [], [0]
        chunk = self.graphs_per_chunk(graphs, samples_per_graph)
        g_off = r_off = 0
        for c0 in range(0, len(graphs), chunk):
            sub = graphs[c0:c0 + chunk]
            b, ws = self.engine.pack(sub, samples_per_graph)
            sl_g, sl_r = slice(g_off, g_off + b.B), slice(r_off, r_off + b.n_rot)
            if randomize:
                if init is None:
                    tor0 = (torch.rand(b.n_rot, generator=generator, device=dev) * 2 - 1) * math.pi
                    rot0 = random_rotations(b.B, generator, dev)
                    tr0 = torch.randn(b.B, 3, generator=generator, device=dev) * self.w.cfg['tr_sigma_max']
                else:
                    tor0 = torch.as_tensor(init['tor'][sl_r], dtype=torch.float32).to(dev)
                    rot0 = torch.as_tensor(init['rot'][sl_g], dtype=torch.float32).to(dev)
                    tr0 = torch.as_tensor(init['tr'][sl_g], dtype=torch.float32).to(dev)
                self.engine.randomize(b, tor0.contiguous(), rot0.reshape(-1, 9).contiguous(), tr0.contiguous(), no_torsion)
            for k in range(self.steps):
                sc = self.consts[k]
                self.engine.forward(b, ws, sc)
                if trace is not None:
                    trace.append((ws.tr.clone().cpu(), ws.rot.clone().cpu(), ws.tor[:b.n_rot].clone().cpu()))
                last = k == self.steps - 1
                if no_random or (self.no_final_step_noise and last):
                    z = (None, None, None)
                elif noise is None:
                    z = (torch.randn(b.B, 3, generator=generator, device=dev), torch.randn(b.B, 3, generator=generator, device=dev),
                         torch.randn(max(b.n_rot, 1), generator=generator, device=dev))
                else:
                    z = tuple(torch.as_tensor(np.asarray(noise[k][key])[s], dtype=torch.float32).contiguous().to(dev)
                              for key, s in (('tr', sl_g), ('rot', sl_g), ('tor', sl_r)))
                self.engine.update(b, ws, sc, *z, no_torsion=no_torsion)
            self.gpu_launches += ws.n_launches
            out_pos.append(b.pos.cpu())
            out_ptr += (np.cumsum(b.n_per) + out_ptr[-1]).tolist()
            g_off += b.B
            r_off += b.n_rot
        return torch.cat(out_pos, 0), np.asarray(out_ptr)
