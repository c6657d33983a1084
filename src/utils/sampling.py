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


def _same_topology(a, b):
    """True when b is a copy of pair a (the reference builds its data_list with copy.deepcopy, inference.py:184): same atoms,
    bonds, rotatable-bond masks and pharmacophore; only the pose (ligand.pos / ligand.norm) may differ."""
    la, lb, pa, pb = a['ligand'], b['ligand'], a['phore'], b['phore']
    if la.pos.shape != lb.pos.shape or pa.pos.shape != pb.pos.shape:
        return False
    ea, eb = a['ligand', 'ligand'], b['ligand', 'ligand']
    qa, qb = a['phore', 'phore'], b['phore', 'phore']
    if ea.edge_index.shape != eb.edge_index.shape or qa.edge_index.shape != qb.edge_index.shape:
        return False
    same = (torch.equal(la.x, lb.x) and torch.equal(ea.edge_index, eb.edge_index) and torch.equal(ea.edge_attr, eb.edge_attr)
            and torch.equal(torch.as_tensor(la.edge_mask), torch.as_tensor(lb.edge_mask))
            and torch.equal(pa.x, pb.x) and torch.equal(pa.pos, pb.pos) and torch.equal(pa.norm, pb.norm)
            and torch.equal(qa.edge_index, qb.edge_index) and torch.equal(la.phorefp, lb.phorefp)
            and torch.equal(la.norm_angle1, lb.norm_angle1) and torch.equal(la.norm_angle2, lb.norm_angle2))
    if not same:
        return False
    ma, mb = la.mask_rotate, lb.mask_rotate
    ma, mb = (ma if isinstance(ma, (np.ndarray, torch.Tensor)) else ma[0]), (mb if isinstance(mb, (np.ndarray, torch.Tensor)) else mb[0])
    return np.array_equal(np.asarray(ma), np.asarray(mb))


def group_copies(data_list):
    """Runs of consecutive copies of one pair -> (unique pairs, samples per pair) when every run has the same length, else None.
    With it the N copies the reference deep-copies per pair are uploaded ONCE and expanded on the device (PackedBatch)."""
    uniq, counts = [], []
    for g in data_list:
        if uniq and _same_topology(uniq[-1], g):
            counts[-1] += 1
        else:
            uniq.append(g)
            counts.append(1)
    if len(set(counts)) != 1 or counts[0] == 1:
        return None
    return uniq, counts[0]


def _sampler_for(model, device, inference_steps, no_final_step_noise, ode):
    """One DenoisingSampler (20 x step constants, weight images, CUDA-graph caches) per model and flag set, kept on the model."""
    cache = model.__dict__.setdefault('_dp_samplers', {})
    key = (str(device), int(inference_steps), bool(no_final_step_noise), bool(ode), id(model.kernel_weights(device)))
    if key not in cache:
        so3n, torn = model.score_norm_tables()
        cache[key] = DenoisingSampler(model.kernel_weights(device), inference_steps, so3n, torn,
                                      no_final_step_noise=no_final_step_noise, ode=ode)
    return cache[key]


def sampling_phore(data_list, model, inference_steps, tr_schedule, rot_schedule, tor_schedule, device, t_to_sigma,
                   model_args, no_random=False, ode=False, visualization_list=None, confidence_model=None,
                   confidence_data_list=None, confidence_model_args=None, batch_size=20, no_final_step_noise=False):
    if confidence_model is not None or visualization_list is not None:
        raise NotImplementedError('B200 path: sampling without confidence model / visualisation only')
    sampler = _sampler_for(model, device, inference_steps, no_final_step_noise, ode)
    keep = bool(getattr(model_args, 'keep_update', False))
    randomize = all('_dp_randomize' in g for g in data_list)
    data_list = list(data_list)
    grouped = group_copies(data_list)
    l0 = sampler.gpu_launches
    if grouped is not None:
        # the copies of a pair travel once; their (possibly different) start poses are per-sample arrays
        uniq, S = grouped
        start = (torch.cat([g['ligand'].pos for g in data_list]).float(),
                 torch.cat([g['ligand'].norm.reshape(g['ligand'].pos.shape[0], -1) for g in data_list]).float())
        pos, ptr = sampler.run(uniq, S, no_random=no_random, randomize=randomize, start_poses=start,
                               no_torsion=getattr(model_args, 'no_torsion', False), keep_update=keep)
    else:
        pos, ptr = sampler.run(data_list, 1, no_random=no_random, randomize=randomize,
                               no_torsion=getattr(model_args, 'no_torsion', False), keep_update=keep)
    for i, g in enumerate(data_list):
        g['ligand'].pos = pos[ptr[i]:ptr[i + 1]].clone()
        if keep:                                            # diffusion_utils.py:71-77 (poses after every step)
            g.docked_poses = [p.numpy() for p in sampler.last_trajectory[1:, ptr[i]:ptr[i + 1]]]
        g._attrs.pop('_dp_randomize', None)
    model.last_gpu_launches = sampler.gpu_launches - l0
    return data_list, None
