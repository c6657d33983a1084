"""utils.diffusion_utils of the reference (/root/reference/src/utils/diffusion_utils.py): schedule / embedding /
time helpers kept on the host; `modify_conformer` (:23-79) is executed by dp_conformer_update on the GPU for whole
batches (see utils.sampling), the per-graph Python version is not part of this path."""
import math

import numpy as np
import torch
import torch.nn.functional as F


def _geometric(lo, hi, t):
    """Noise level between lo (t = 0) and hi (t = 1) on a geometric scale."""
    return lo ** (1 - t) * hi ** t


def t_to_sigma(t_tr, t_rot, t_tor, args):
    """(tr_sigma, rot_sigma, tor_sigma) at diffusion times (t_tr, t_rot, t_tor); reference diffusion_utils.py:16-20."""
    return (_geometric(args.tr_sigma_min, args.tr_sigma_max, t_tr), _geometric(args.rot_sigma_min, args.rot_sigma_max, t_rot),
            _geometric(args.tor_sigma_min, args.tor_sigma_max, t_tor))


def sinusoidal_embedding(timesteps, embedding_dim, max_positions=10000):
    """[n] -> [n, embedding_dim]: sin | cos of t * max_positions^(-i / (half - 1)) (reference diffusion_utils.py:82-93)."""
    if timesteps.dim() != 1:
        raise ValueError('timesteps must be a vector')
    half = embedding_dim // 2
    decay = math.log(max_positions) / (half - 1)
    freqs = torch.exp(torch.arange(half, dtype=torch.float32, device=timesteps.device) * -decay)
    phase = timesteps.float().unsqueeze(1) * freqs.unsqueeze(0)
    out = torch.cat([phase.sin(), phase.cos()], dim=1)
    return F.pad(out, (0, 1)) if embedding_dim % 2 else out


def get_timestep_embedding(embedding_type, embedding_dim, embedding_scale=10000):
    if embedding_type != 'sinusoidal':
        raise NotImplementedError('only the shipped sinusoidal embedding is supported')
    return lambda x: sinusoidal_embedding(embedding_scale * x, embedding_dim)


def get_t_schedule(inference_steps):
    """t_k = 1 - k / steps, k = 0 .. steps - 1 (reference diffusion_utils.py:135-145)."""
    return np.linspace(1, 0, inference_steps + 1)[:-1]


def set_time_phore(graphs, t_tr, t_rot, t_tor, batchsize, device):
    """Every node and every graph of the batch carries the same diffusion time (reference diffusion_utils.py:181-207)."""
    times = {'tr': t_tr, 'rot': t_rot, 'tor': t_tor}

    def stamped(count):
        return {key: t * torch.ones(count).to(device) for key, t in times.items()}
    for node_type in ('ligand', 'phore'):
        graphs[node_type].node_t = stamped(graphs[node_type].pos.shape[0])
    graphs.complex_t = stamped(batchsize)
