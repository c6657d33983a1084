"""utils.torsion of the reference (/root/reference/src/utils/torsion.py:64-109): sequential torsion updates of ONE conformer on
the host (numpy + scipy), used by the single-graph training-time drivers; the denoising loop does this on the GPU."""
import numpy as np
import torch
from scipy.spatial.transform import Rotation as R


def modify_conformer_torsion_angles(pos, edge_index, mask_rotate, torsion_updates, norm=None):
    """Rotate, bond after bond, the atoms on the `mask_rotate` side of every rotatable bond (u, v) about the bond axis by
    torsion_updates[k] (skipped when 0); bonds act on the positions already modified by the earlier ones.  Returns (pos, norm)
    - norm (the 11 N "norm points", [K, N, 3]) follows the same rotations when given."""
    is_tensor = torch.is_tensor(pos)
    p = pos.detach().cpu().numpy().copy() if is_tensor else np.array(pos).copy()
    nr = None if norm is None else (norm.detach().cpu().numpy().copy() if torch.is_tensor(norm) else np.array(norm).copy())
    edges = edge_index.detach().cpu().numpy() if torch.is_tensor(edge_index) else np.asarray(edge_index)
    for k, (u, v) in enumerate(edges):
        theta = torsion_updates[k]
        if theta == 0:
            continue
        assert not mask_rotate[k, u] and mask_rotate[k, v]
        axis = p[u] - p[v]
        rot = R.from_rotvec(axis * theta / np.linalg.norm(axis)).as_matrix()
        side = mask_rotate[k]
        p[side] = (p[side] - p[v]) @ rot.T + p[v]
        if nr is not None:
            nr[:, side] = (nr[:, side] - p[v]) @ rot.T + p[v]
    if is_tensor:
        return torch.from_numpy(p.astype(np.float32)), (None if nr is None else torch.from_numpy(nr.astype(np.float32)))
    return p, nr
