"""utils.sampling of the reference (/root/reference/src/utils/sampling.py): `randomize_position` (:16-63) and
`sampling_phore` (:174-255) with the reference's signatures, executed on the GPU by diffphore_b200.sampler.

`sampling_phore(data_list, model, ...)` keeps the contract "list of graphs in -> list of graphs with final
ligand.pos out"; the 20-step loop, score model and conformer updates all run as sm_100a kernels with every graph
resident in HBM (no per-step re-collation, no per-sample host update).  `batch_size` is accepted for signature
compatibility; the device batch is sized by HBM instead.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from diffphore_b200.sampler import DenoisingSampler, random_rotations   # noqa: E402


def randomize_position(data_list, no_torsion, no_random, tr_sigma_max, keep_update=False):
    """Marks the graphs for device-side randomisation (dp_randomize_position runs inside sampling_phore, where the
    batch lives); draws come from torch's global generator.  The input poses are left untouched here."""
    for g in data_list:
        g._attrs['_dp_randomize'] = dict(no_torsion=no_torsion, no_random=no_random, tr_sigma_max=tr_sigma_max)


def sampling_phore(data_list, model, inference_steps, tr_schedule, rot_schedule, tor_schedule, device, t_to_sigma,
                   model_args, no_random=False, ode=False, visualization_list=None, confidence_model=None,
                   confidence_data_list=None, confidence_model_args=None, batch_size=20, no_final_step_noise=False):
    if confidence_model is not None or visualization_list is not None:
        raise NotImplementedError('B200 path: sampling without confidence model / visualisation only')
    so3n, torn = model.score_norm_tables()
    sampler = DenoisingSampler(model.kernel_weights(device), inference_steps, so3n, torn,
                               no_final_step_noise=no_final_step_noise, ode=ode)
    keep = bool(getattr(model_args, 'keep_update', False))
    randomize = all('_dp_randomize' in g for g in data_list)
    pos, ptr = sampler.run(list(data_list), 1, no_random=no_random, randomize=randomize,
                           no_torsion=getattr(model_args, 'no_torsion', False), keep_update=keep)
    for i, g in enumerate(data_list):
        g['ligand'].pos = pos[ptr[i]:ptr[i + 1]].clone()
        if keep:                                            # diffusion_utils.py:71-77 (poses after every step)
            g.docked_poses = [p.numpy() for p in sampler.last_trajectory[1:, ptr[i]:ptr[i + 1]]]
        g._attrs.pop('_dp_randomize', None)
    model.last_gpu_launches = sampler.gpu_launches
    return data_list, None
