"""ORACLE (test infrastructure, not product code).

CPU restatement of the reverse-diffusion driver and the per-sample conformer update, bug-for-bug; PINNED on the reference's own
code: tests/golden/ref_sampler.npz holds outputs of the unmodified reference modules run over shims (tools/make_sampler_golden.py),
tests/test_oracle.py compares every function below with them.

    randomize_position                 /root/reference/src/utils/sampling.py:16-63
    sampling_phore                     /root/reference/src/utils/sampling.py:174-255
    get_t_schedule / t_to_sigma        /root/reference/src/utils/diffusion_utils.py:135-145,16-20
    set_time_phore                     /root/reference/src/utils/diffusion_utils.py:181-207
    modify_conformer                   /root/reference/src/utils/diffusion_utils.py:23-79
    modify_conformer_torsion_angles    /root/reference/src/utils/torsion.py:64-109
    axis_angle_to_matrix (quaternion)  /root/reference/src/utils/geometry.py:6-85
    rigid_transform_Kabsch_3D_torch    /root/reference/src/utils/geometry.py:88-136

Deviations, all so that oracle and CUDA path consume identical randomness: the Gaussian draws of
every step (`tr_z`, `rot_z`, `tor_z`, sampling.py:230-244) and the initial torsion / rotation / translation
draws (sampling.py:35,49,58) are taken from caller-supplied arrays instead of the global numpy / torch RNGs.
`collate` / `to_data_list` are callables supplied by the caller (PyG is absent): the per-step re-collation of
sampling.py:210,254 is kept because it is part of the reference's cost.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import copy
import math

import numpy as np
import torch
from scipy.spatial.transform import Rotation as R


def get_t_schedule(inference_steps):
    return np.linspace(1, 0, inference_steps + 1)[:-1]


def quaternion_to_matrix(q):
    r, i, j, k = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                     two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                     two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return o.reshape(q.shape[:-1] + (3, 3))


def axis_angle_to_matrix(axis_angle):
    angles = torch.norm(axis_angle, p=2, dim=-1, keepdim=True)
    half = 0.5 * angles
    small = angles.abs() < 1e-6
    s = torch.empty_like(angles)
    s[~small] = torch.sin(half[~small]) / angles[~small]
    s[small] = 0.5 - (angles[small] * angles[small]) / 48
    return quaternion_to_matrix(torch.cat([torch.cos(half), axis_angle * s], dim=-1))


def rigid_transform_kabsch(A, B):
    """geometry.py:88-136 on 3xN inputs."""
    cA, cB = torch.mean(A, dim=1, keepdim=True), torch.mean(B, dim=1, keepdim=True)
    H = (A - cA) @ (B - cB).T
    U, S, Vt = torch.linalg.svd(H)
    Rm = Vt.T @ U.T
    if torch.linalg.det(Rm) < 0:
        SS = torch.diag(torch.tensor([1., 1., -1.], dtype=A.dtype))
        Rm = (Vt.T @ SS) @ U.T
    assert math.fabs(torch.linalg.det(Rm) - 1) < 3e-3
    return Rm, -Rm @ cA + cB


def modify_conformer_torsion_angles(pos, edge_index, mask_rotate, torsion_updates, norm=None):
    """torsion.py:64-109: numpy, fp32 positions, fp64 scipy rotation matrices (H9)."""
    pos = copy.deepcopy(pos)
    if type(pos) != np.ndarray:
        pos = pos.cpu().numpy()
    if norm is not None and type(norm) != np.ndarray:
        norm = norm.cpu().numpy()
    for idx_edge, e in enumerate(edge_index.cpu().numpy()):
        if torsion_updates[idx_edge] == 0:
            continue
        u, v = e[0], e[1]
        assert not mask_rotate[idx_edge, u]
        assert mask_rotate[idx_edge, v]
        rot_vec = pos[u] - pos[v]
        rot_vec = rot_vec * torsion_updates[idx_edge] / np.linalg.norm(rot_vec)
        rot_mat = R.from_rotvec(rot_vec).as_matrix()
        pos[mask_rotate[idx_edge]] = (pos[mask_rotate[idx_edge]] - pos[v]) @ rot_mat.T + pos[v]
        if norm is not None:
            norm[:, mask_rotate[idx_edge]] = (norm[:, mask_rotate[idx_edge]] - pos[v]) @ rot_mat.T + pos[v]
    pos = torch.from_numpy(pos.astype(np.float32))
    norm = torch.from_numpy(norm.astype(np.float32)) if norm is not None else None
    return pos, norm


def _mask_rotate(lig):
    return lig.mask_rotate if isinstance(lig.mask_rotate, np.ndarray) else lig.mask_rotate[0]


def modify_conformer(data, tr_update, rot_update, torsion_updates):
    """diffusion_utils.py:23-79 (keep_update=False)."""
    lig = data['ligand']
    n = lig.x.shape[0]
    lig_center = torch.mean(lig.pos, dim=0, keepdim=True)
    lig_norm = lig.norm.reshape(-1, n, 3) + lig.pos.unsqueeze(0)             # H4: raw reinterpretation
    rot_mat = axis_angle_to_matrix(rot_update.squeeze())
    rigid_new_pos = (lig.pos - lig_center) @ rot_mat.T + tr_update + lig_center
    rigid_new_norm = (lig_norm - lig_center) @ rot_mat.T + tr_update + lig_center
    if torsion_updates is not None:
        bonds = data['ligand', 'ligand'].edge_index.T[lig.edge_mask]
        flex_pos, flex_norm = modify_conformer_torsion_angles(rigid_new_pos, bonds, _mask_rotate(lig),
                                                              torsion_updates, norm=rigid_new_norm)
        Rm, t = rigid_transform_kabsch(flex_pos.T, rigid_new_pos.T)
        aligned_pos = flex_pos @ Rm.T + t.T
        aligned_norm = flex_norm @ Rm.T + t.T - aligned_pos
        lig.pos = aligned_pos
        lig.norm = aligned_norm.reshape(n, -1)
    else:
        lig.pos = rigid_new_pos
        lig.norm = (rigid_new_norm - rigid_new_pos).reshape(n, -1)
    return data


def randomize_position(data_list, no_torsion, no_random, tr_sigma_max, tor_init, rot_init, tr_init):
    """sampling.py:16-63 with injected draws: tor_init[i] ~ U(-pi,pi) [n_rot], rot_init[i] = 3x3 rotation
    (scipy R.random().as_matrix()), tr_init[i] ~ N(0, tr_sigma_max) [1,3]."""
    if not no_torsion:
        for i, g in enumerate(data_list):
            lig = g['ligand']
            n = lig.x.shape[0]
            norm = lig.norm.reshape(-1, n, 3) + lig.pos.unsqueeze(0)
            lig.pos, lig.norm = modify_conformer_torsion_angles(
                lig.pos, g['ligand', 'ligand'].edge_index.T[lig.edge_mask], _mask_rotate(lig),
                np.asarray(tor_init[i], dtype=np.float64), norm=norm)
    for i, g in enumerate(data_list):
        lig = g['ligand']
        center = torch.mean(lig.pos, dim=0, keepdim=True)
        rot = torch.from_numpy(np.asarray(rot_init[i])).float()
        lig.pos = (lig.pos - center) @ rot.T
        lig.norm = ((lig.norm - center) @ rot.T - lig.pos).reshape(lig.pos.shape[0], -1)
        if not no_random:
            lig.pos = lig.pos + torch.as_tensor(tr_init[i], dtype=torch.float32).reshape(1, 3)


def set_time(batch, t, b):
    """set_time_phore (diffusion_utils.py:181-207): every node / graph carries the same scalar t (fp32)."""
    for nt in ('ligand', 'phore'):
        n = batch[nt].pos.shape[0]
        batch[nt].node_t = {k: t * torch.ones(n) for k in ('tr', 'rot', 'tor')}
    batch.complex_t = {k: t * torch.ones(b) for k in ('tr', 'rot', 'tor')}


def sampling(data_list, model, inference_steps, cfg, collate, batch_size, noise=None, no_torsion=False,
             trace=None, ode=False):
    """sampling_phore (sampling.py:174-255): Euler–Maruyama, or the probability-flow ODE branch (:226-228,240-241).

    noise: None (-> no_random=True, zeros) or a list over steps of dicts
           {'tr': [n,3], 'rot': [n,3], 'tor': [sum n_rot]} in data_list order.
    Returns the final data_list (positions in each graph's ['ligand'].pos)."""
    sched = get_t_schedule(inference_steps)
    for t_idx in range(inference_steps):
        t = sched[t_idx]
        dt = sched[t_idx] - sched[t_idx + 1] if t_idx < inference_steps - 1 else sched[t_idx]
        new_list, g0, tor0 = [], 0, 0
        for s in range(0, len(data_list), batch_size):
            batch = collate(data_list[s:s + batch_size])
            b = batch.num_graphs
            tr_sigma = cfg['tr_sigma_min'] ** (1 - t) * cfg['tr_sigma_max'] ** t
            rot_sigma = cfg['rot_sigma_min'] ** (1 - t) * cfg['rot_sigma_max'] ** t
            tor_sigma = cfg['tor_sigma_min'] ** (1 - t) * cfg['tor_sigma_max'] ** t
            set_time(batch, t, b)
            with torch.no_grad():
                tr_score, rot_score, tor_score = model(batch)
            tr_score, rot_score, tor_score = tr_score.float(), rot_score.float(), tor_score.float()
            if trace is not None:
                trace.append((tr_score.clone(), rot_score.clone(), tor_score.clone()))
            tr_g = tr_sigma * torch.sqrt(torch.tensor(2 * np.log(cfg['tr_sigma_max'] / cfg['tr_sigma_min'])))
            rot_g = 2 * rot_sigma * torch.sqrt(torch.tensor(np.log(cfg['rot_sigma_max'] / cfg['rot_sigma_min'])))
            n_tor = tor_score.shape[0]
            if noise is None:
                tr_z, rot_z, tor_z = torch.zeros(b, 3), torch.zeros(b, 3), torch.zeros(n_tor)
            else:
                tr_z = torch.as_tensor(noise[t_idx]['tr'][g0:g0 + b], dtype=torch.float32)
                rot_z = torch.as_tensor(noise[t_idx]['rot'][g0:g0 + b], dtype=torch.float32)
                tor_z = torch.as_tensor(noise[t_idx]['tor'][tor0:tor0 + n_tor], dtype=torch.float32)
            if ode:
                tr_perturb = 0.5 * tr_g ** 2 * dt * tr_score
                rot_perturb = 0.5 * rot_score * dt * rot_g ** 2
            else:
                tr_perturb = (tr_g ** 2 * dt * tr_score + tr_g * np.sqrt(dt) * tr_z)
                rot_perturb = (rot_score * dt * rot_g ** 2 + rot_g * np.sqrt(dt) * rot_z)
            if not no_torsion:
                tor_g = tor_sigma * torch.sqrt(torch.tensor(2 * np.log(cfg['tor_sigma_max'] / cfg['tor_sigma_min'])))
                tor_perturb = (0.5 * tor_g ** 2 * dt * tor_score).numpy() if ode else \
                    (tor_g ** 2 * dt * tor_score + tor_g * np.sqrt(dt) * tor_z).numpy()
            graphs = batch.to_data_list()
            off = 0
            for i, g in enumerate(graphs):
                k = int(g['ligand'].edge_mask.sum())        # per-graph count (the reference assumes it uniform)
                tp = tor_perturb[off:off + k] if not no_torsion else None
                off += k
                new_list.append(modify_conformer(g, tr_perturb[i:i + 1], rot_perturb[i:i + 1].squeeze(0), tp))
            g0 += b
            tor0 += n_tor
        data_list = new_list
    return data_list
