"""utils.diffusion_utils of the reference (/root/reference/src/utils/diffusion_utils.py): schedule / embedding /
time helpers kept on the host; `modify_conformer` (:23-79) is executed by dp_conformer_update on the GPU for whole
batches (see utils.sampling), the per-graph Python version is not part of this path."""
import math

import numpy as np
import torch
import torch.nn.functional as F


def t_to_sigma(t_tr, t_rot, t_tor, args):
    tr_sigma = args.tr_sigma_min ** (1 - t_tr) * args.tr_sigma_max ** t_tr
    rot_sigma = args.rot_sigma_min ** (1 - t_rot) * args.rot_sigma_max ** t_rot
    tor_sigma = args.tor_sigma_min ** (1 - t_tor) * args.tor_sigma_max ** t_tor
    return tr_sigma, rot_sigma, tor_sigma


def sinusoidal_embedding(timesteps, embedding_dim, max_positions=10000):
    assert len(timesteps.shape) == 1
    half_dim = embedding_dim // 2
    emb = math.log(max_positions) / (half_dim - 1)
    emb = torch.exp(torch.arange(half_dim, dtype=torch.float32, device=timesteps.device) * -emb)
    emb = timesteps.float()[:, None] * emb[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=1)
    if embedding_dim % 2 == 1:
        emb = F.pad(emb, (0, 1), mode='constant')
    return emb


def get_timestep_embedding(embedding_type, embedding_dim, embedding_scale=10000):
    if embedding_type != 'sinusoidal':
        raise NotImplementedError('only the shipped sinusoidal embedding is supported')
    return lambda x: sinusoidal_embedding(embedding_scale * x, embedding_dim)


def get_t_schedule(inference_steps):
    return np.linspace(1, 0, inference_steps + 1)[:-1]


def set_time_phore(graphs, t_tr, t_rot, t_tor, batchsize, device):
    for nt in ('ligand', 'phore'):
        n = graphs[nt].pos.shape[0]
        graphs[nt].node_t = {'tr': t_tr * torch.ones(n).to(device), 'rot': t_rot * torch.ones(n).to(device),
                             'tor': t_tor * torch.ones(n).to(device)}
    graphs.complex_t = {'tr': t_tr * torch.ones(batchsize).to(device), 'rot': t_rot * torch.ones(batchsize).to(device),
                        'tor': t_tor * torch.ones(batchsize).to(device)}
