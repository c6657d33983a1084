"""utils.sampling of the reference (/root/reference/src/utils/sampling.py): `randomize_position` (:16-63) and
`sampling_phore` (:174-255) with the reference's signatures, executed on the GPU by diffphore_b200.sampler.

`sampling_phore(data_list, model, ...)` keeps the contract "list of graphs in -> list of graphs with final
ligand.pos out"; the 20-step loop, score model and conformer updates all run as sm_100a kernels with every graph
resident in HBM (no per-step re-collation, no per-sample host update).  `batch_size` is accepted for signature
compatibility; the device batch is sized by HBM instead.
"""
import copy
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


# =====================================================================================================================
# the drivers that re-use the denoising kernels one step at a time (calibrated sampler / fitscore-guided sampling)
# =====================================================================================================================
def _unit_update_consts(device):
    """Constant block for dp_conformer_update that applies GIVEN perturbations: perturbation = 1 * "score" + 0 * noise."""
    from diffphore_b200 import lib as L
    sc = torch.zeros(L.SC['SIZE'], dtype=torch.float32)
    sc[L.SC['TR_A']] = sc[L.SC['ROT_A']] = sc[L.SC['TOR_A']] = 1.0
    return sc.to(device)


def apply_perturbations(model, batch, tr_perturb, rot_perturb, tor_perturb):
    """modify_conformer (diffusion_utils.py:23-79) for every graph of a collated batch on the GPU, with the perturbations given
    (tr [b,3], rot [b,3], tor [sum n_rot] or None): the packed batch (cached on the model by topology, e.g. by the preceding
    model(batch) call) is updated in place by dp_conformer_update and read back.  Returns the data_list with the new ligand.pos / ligand.norm."""
    w = model.kernel_weights()
    _, eng, b, ws = model._pack_only(batch)
    dev = w.device
    b.pos.copy_(batch['ligand'].pos.to(torch.float32).reshape(b.n_lig, 3))
    b.norm.copy_(batch['ligand'].norm.to(torch.float32).reshape(b.n_lig, 33))
    ws.tr.copy_(torch.as_tensor(tr_perturb, dtype=torch.float32).reshape(b.B, 3))
    ws.rot.copy_(torch.as_tensor(rot_perturb, dtype=torch.float32).reshape(b.B, 3))
    no_torsion = tor_perturb is None
    if not no_torsion and b.n_rot:
        ws.tor[:b.n_rot].copy_(torch.as_tensor(np.asarray(tor_perturb), dtype=torch.float32).reshape(b.n_rot))
    sc = model.__dict__.setdefault('_unit_sc', _unit_update_consts(dev))
    eng.update(b, ws, sc, None, None, None, no_torsion=no_torsion)
    pos, norm = b.pos.cpu(), b.norm.cpu()
    graphs = batch.to_data_list()
    for g, lo, hi in zip(graphs, b.lig_ptr.cpu().tolist()[:-1], b.lig_ptr.cpu().tolist()[1:]):
        g['ligand'].pos = pos[lo:hi].clone()
        g['ligand'].norm = norm[lo:hi].reshape(g['ligand'].norm.shape).clone()
    return graphs


def sample_step(complex_graph_batch, model, model_args, tr_sigma, rot_sigma, tor_sigma, delta_t=0.05, no_random=False, ode=False):
    """One Euler-Maruyama step of a collated batch at given noise levels (reference sampling.py:501-559): score model on the GPU,
    perturbations on the host exactly as the reference computes them (the Gaussian draws come from torch's CPU generator in the
    reference's order tr, rot, tor), conformer update on the GPU.  Returns (data_list, tor_perturb, tr_perturb, rot_perturb)."""
    if getattr(model_args, 'keep_update', False):
        raise NotImplementedError('B200 path: sample_step without keep_update')
    b = complex_graph_batch.num_graphs
    with torch.no_grad():
        tr_score, rot_score, tor_score = model(complex_graph_batch)
    tr_score, rot_score, tor_score = tr_score.cpu(), rot_score.cpu(), tor_score.cpu()
    tr_g = tr_sigma * torch.sqrt(torch.tensor(2 * np.log(model_args.tr_sigma_max / model_args.tr_sigma_min)))
    rot_g = 2 * rot_sigma * torch.sqrt(torch.tensor(np.log(model_args.rot_sigma_max / model_args.rot_sigma_min)))
    if ode:
        tr_perturb = 0.5 * tr_g ** 2 * delta_t * tr_score
        rot_perturb = 0.5 * rot_score * delta_t * rot_g ** 2
    else:
        tr_z = torch.zeros((b, 3)) if no_random else torch.normal(mean=0, std=1, size=(b, 3))
        tr_perturb = tr_g ** 2 * delta_t * tr_score + tr_g * np.sqrt(delta_t) * tr_z
        rot_z = torch.zeros((b, 3)) if no_random else torch.normal(mean=0, std=1, size=(b, 3))
        rot_perturb = rot_score * delta_t * rot_g ** 2 + rot_g * np.sqrt(delta_t) * rot_z
    tor_perturb = None
    if not model_args.no_torsion:
        tor_g = tor_sigma * torch.sqrt(torch.tensor(2 * np.log(model_args.tor_sigma_max / model_args.tor_sigma_min)))
        if ode:
            tor_perturb = (0.5 * tor_g ** 2 * delta_t * tor_score).numpy()
        else:
            tor_z = torch.zeros(tor_score.shape) if no_random else torch.normal(mean=0, std=1, size=tor_score.shape)
            tor_perturb = (tor_g ** 2 * delta_t * tor_score + tor_g * np.sqrt(delta_t) * tor_z).numpy()
    data_list = apply_perturbations(model, complex_graph_batch, tr_perturb.float(), rot_perturb.float(), tor_perturb)
    return data_list, tor_perturb, tr_perturb, rot_perturb


def calculate_fitscore(args, ligand_pos, name, mol, phore_file=None, store_ranked_pose=True):
    """Poses -> SD file -> AncPhore fitness (reference sampling.py:447-498); `mol` is the SD template of the ligand
    (datasets.process_mols.ligand_graph_from_sdf keeps it on the graph as `sdf_template`)."""
    from datasets.process_mols import write_mol_with_multi_coords
    from datasets.process_pharmacophore import calc_phore_fitting
    tmp_path = os.path.join(args.run_dir, f'mapping_process/{name}')
    os.makedirs(tmp_path, exist_ok=True)
    docked_file = os.path.join(tmp_path, f'{name}.sdf')
    write_mol_with_multi_coords(mol, ligand_pos, docked_file, name)
    if phore_file is None or not os.path.exists(phore_file):
        raise NotImplementedError('calculate_fitscore needs the pharmacophore file of the pair (dataset look-ups are not mirrored)')
    scores = calc_phore_fitting(docked_file, phore_file, os.path.join(tmp_path, f'{name}.score'), os.path.join(tmp_path, f'{name}.dbphore'),
                                os.path.join(tmp_path, f'{name}.log'), overwrite=True, fitness=getattr(args, 'fitness', 1),
                                ancphore_path=os.path.join(getattr(args, 'ancphore_path', 'programs'), 'AncPhore'))
    if store_ranked_pose and scores is not None:
        ranked = os.path.join(args.run_dir, 'ranked_poses')
        os.makedirs(ranked, exist_ok=True)
        perm = np.argsort(np.array(scores))[::-1]
        write_mol_with_multi_coords(mol, np.asarray(ligand_pos)[perm], os.path.join(ranked, f'{name}_ranked.sdf'), name, marker='rank',
                                    properties={'fitscore': np.array(scores)[perm]})
    return scores


def sampling_phore_with_fitscore(data_list, model, inference_steps, tr_schedule, rot_schedule, tor_schedule, device, t_to_sigma,
                                 model_args, no_random=False, ode=False, visualization_list=None, confidence_model=None,
                                 confidence_data_list=None, confidence_model_args=None, batch_size=32, no_final_step_noise=False,
                                 fitscore_fn=None):
    """Fitscore-guided sampling (reference sampling.py:283-444): per step and batch, `random_samples` noisy updates of every graph
    are generated, the poses scored (AncPhore through calculate_fitscore, or `fitscore_fn(args, poses, name, mol, phore_file=...)`)
    and the best one of every graph kept.  The score model and the conformer updates of all candidates run on the GPU; the host
    side (noise from torch's CPU generator, the reference's own - quirky - index arithmetic, scoring) follows the reference line by
    line: d_sigma = g sqrt(dt); ode: 0.5 d_sigma tr (sic) and 0.5 d_sigma^2 rot; scores viewed as [random_samples, b] although
    the candidates are graph-major."""
    from utils.diffusion_utils import set_time_phore
    from diffphore_b200.graph import DataLoader
    if confidence_model is not None or visualization_list is not None:
        raise NotImplementedError('B200 path: sampling without confidence model / visualisation only')
    if getattr(model_args, 'keep_update', False):
        raise NotImplementedError('B200 path: sampling_phore_with_fitscore without keep_update')
    score_fn = fitscore_fn or calculate_fitscore
    R = getattr(model_args, 'random_samples', 0)
    for t_idx in range(inference_steps):
        t_tr, t_rot, t_tor = tr_schedule[t_idx], rot_schedule[t_idx], tor_schedule[t_idx]
        last = t_idx == inference_steps - 1
        dt_tr = tr_schedule[t_idx] - tr_schedule[t_idx + 1] if not last else tr_schedule[t_idx]
        dt_rot = rot_schedule[t_idx] - rot_schedule[t_idx + 1] if not last else rot_schedule[t_idx]
        dt_tor = tor_schedule[t_idx] - tor_schedule[t_idx + 1] if not last else tor_schedule[t_idx]
        new_data_list = []
        for batch in DataLoader(data_list, batch_size=batch_size):
            b = batch.num_graphs
            tr_sigma, rot_sigma, tor_sigma = t_to_sigma(t_tr, t_rot, t_tor)
            set_time_phore(batch, t_tr, t_rot, t_tor, b, 'cpu')
            with torch.no_grad():
                tr_score, rot_score, tor_score = model(batch)
            tr_score, rot_score, tor_score = tr_score.cpu(), rot_score.cpu(), tor_score.cpu()
            tr_g = tr_sigma * torch.sqrt(torch.tensor(2 * np.log(model_args.tr_sigma_max / model_args.tr_sigma_min)))
            rot_g = 2 * rot_sigma * torch.sqrt(torch.tensor(np.log(model_args.rot_sigma_max / model_args.rot_sigma_min)))
            d_sigma_tr, d_sigma_rot = tr_g * np.sqrt(dt_tr), rot_g * np.sqrt(dt_rot)
            last_step = no_final_step_noise and last
            multi = (not no_random) and R > 1 and not last_step
            if multi:
                d_sigma_tr, d_sigma_rot = d_sigma_tr.unsqueeze(0), d_sigma_rot.unsqueeze(0)
                tr_score, rot_score = tr_score.unsqueeze(0), rot_score.unsqueeze(0)
            if ode:
                tr_perturb = 0.5 * d_sigma_tr * tr_score
                rot_perturb = 0.5 * d_sigma_rot ** 2 * rot_score
            else:
                shape = (R, b, 3) if R > 1 else (b, 3)
                tr_z = torch.zeros((b, 3)) if (no_random or last_step) else torch.normal(mean=0, std=1, size=shape)
                tr_perturb = (d_sigma_tr ** 2 * tr_score + d_sigma_tr * tr_z).float()
                rot_z = torch.zeros((b, 3)) if (no_random or last_step) else torch.normal(mean=0, std=1, size=shape)
                rot_perturb = (d_sigma_rot ** 2 * rot_score + d_sigma_rot * rot_z).float()
            tor_perturb = None
            if not model_args.no_torsion:
                tor_g = tor_sigma * torch.sqrt(torch.tensor(2 * np.log(model_args.tor_sigma_max / model_args.tor_sigma_min)))
                d_sigma_tor = tor_g * np.sqrt(dt_tor)
                if (not no_random) and R > 1:
                    d_sigma_tor = d_sigma_tor.unsqueeze(0)
                if ode:
                    raise NotImplementedError('sampling_phore_with_fitscore(ode=True) with torsions: the reference raises NameError here '
                                              '(tor_z is undefined in its ode branch, sampling.py:371-380)')
                elif no_random or last_step:
                    tor_z = torch.zeros(tor_score.shape)
                else:
                    tor_z = torch.normal(mean=0, std=1, size=((R,) + tuple(tor_score.shape)) if R > 1 else tuple(tor_score.shape))
                tor_perturb = (d_sigma_tor ** 2 * tor_score + d_sigma_tor * tor_z).float().numpy()
            if (not no_random) and R > 1:
                if not multi:                                               # last step without noise: the reference indexes [j, i] too
                    raise NotImplementedError('random_samples > 1 with --no_final_step_noise: the reference indexes a 2-D array with 3 indices here')
                # candidates: graph-major copies (graph i, sample j) -> index i * R + j, all updated in ONE GPU batch
                graphs = batch.to_data_list()
                cand = [copy.deepcopy(graphs[i]) for i in range(b) for _ in range(R)]
                n_rot = [int(g['ligand'].edge_mask.sum()) for g in graphs]
                tpm = (tor_perturb.shape[1] // b) if tor_perturb is not None else 0
                trc = torch.stack([tr_perturb[j, i] for i in range(b) for j in range(R)])
                rotc = torch.stack([rot_perturb[j, i] for i in range(b) for j in range(R)])
                torc = None if tor_perturb is None else np.concatenate(
                    [tor_perturb[j, i * tpm:(i + 1) * tpm][:n_rot[i]] for i in range(b) for j in range(R)]) if b * R else None
                from diffphore_b200.graph import collate
                cb = collate(cand)
                set_time_phore(cb, t_tr, t_rot, t_tor, b * R, 'cpu')
                tmp = apply_perturbations(model, cb, trc, rotc, torc)
                g0 = graphs[0]
                filterHs = torch.not_equal(g0['ligand'].x[:, 0], 0).cpu().numpy()
                ligand_pos = np.asarray([g['ligand'].pos.cpu().numpy()[filterHs] for g in tmp])
                dock_pose = ligand_pos + g0.original_center.cpu().numpy()
                name = g0.name if isinstance(g0.name, str) else g0.name[0]
                phore_file = getattr(g0, 'phore_file', None) if 'phore_file' in g0 else None
                mol = getattr(g0, 'sdf_template', None) if 'sdf_template' in g0 else None
                scores = score_fn(model_args, dock_pose, name, mol, store_ranked_pose=False, phore_file=phore_file)
                idx = (torch.tensor(scores).view(R, -1)).argmax(dim=0) + torch.arange(b) * R
                new_data_list.extend([tmp[int(k)] for k in idx])
            else:
                new_data_list.extend(apply_perturbations(model, batch, tr_perturb, rot_perturb, tor_perturb))
        data_list = new_data_list
    return data_list, None


def t_centered_A(A, _R, t):
    return A.mean(axis=0) @ _R.T - A.mean(axis=0) + t


def get_updates_from_0_to_n(g, g_n, torsion_updates):
    """(translation, rotation vector) that carry graph g - after `torsion_updates` and the Kabsch re-alignment of the conformer update
    - onto g_n (reference sampling.py:566-597); host code, one graph."""
    from scipy.spatial.transform import Rotation
    from utils.geometry import rigid_transform_Kabsch_3D_torch
    from utils.torsion import modify_conformer_torsion_angles
    g_0 = copy.deepcopy(g)
    if torsion_updates is not None:
        mr = g_0['ligand'].mask_rotate
        flex, _ = modify_conformer_torsion_angles(g_0['ligand'].pos, g_0['ligand', 'ligand'].edge_index.T[g_0['ligand'].edge_mask],
                                                  mr if isinstance(mr, np.ndarray) else mr[0], torsion_updates, norm=None)
        _R, t = rigid_transform_Kabsch_3D_torch(flex.T, g_0['ligand'].pos.T)
        g_0['ligand'].pos = flex @ _R.T + t.T
    R1, t1 = rigid_transform_Kabsch_3D_torch(g_0['ligand'].pos.T, g_n['ligand'].pos.T)
    t2 = t_centered_A(g_0['ligand'].pos, R1, t1.T)
    return t2, Rotation.from_matrix(R1.numpy()).as_rotvec()
