"""utils.geometry of the reference (/root/reference/src/utils/geometry.py): host-side helpers for SINGLE graphs.  The batched
rigid / torsion update of the denoising loop runs on the GPU (dp_conformer_update); these numpy / torch versions serve the
training-time drivers that work on one graph at a time (get_updates_from_0_to_n, sampling.py:566-597)."""
import numpy as np
import torch


def rigid_transform_Kabsch_3D_torch(A, B):
    """Least-squares rigid transform (R, t) with R @ A + t ~ B for 3 x N point sets (reference geometry.py:88-136): covariance of
    the centred sets, SVD, last singular vector flipped when the determinant is negative."""
    if A.shape[0] != 3 or B.shape[0] != 3:
        raise Exception(f'matrix A / B is not 3xN, it is {tuple(A.shape)} / {tuple(B.shape)}')
    ca, cb = A.mean(dim=1, keepdim=True), B.mean(dim=1, keepdim=True)
    H = (A - ca) @ (B - cb).T
    U, S, Vt = torch.linalg.svd(H)
    R = Vt.T @ U.T
    if torch.linalg.det(R) < 0:
        SS = torch.diag(torch.tensor([1.0, 1.0, -1.0], device=A.device, dtype=A.dtype))
        R = (Vt.T @ SS) @ U.T
    assert abs(float(torch.linalg.det(R)) - 1) < 3e-3
    return R, -R @ ca + cb


def axis_angle_to_matrix(axis_angle):
    """Rodrigues formula through the unit quaternion like the reference (geometry.py:38-85; small-angle series below 1e-6)."""
    angles = torch.norm(axis_angle, p=2, dim=-1, keepdim=True)
    half = 0.5 * angles
    small = angles.abs() < 1e-6
    s = torch.where(small, 0.5 - angles * angles / 48, torch.sin(half) / torch.where(small, torch.ones_like(angles), angles))
    q = torch.cat([torch.cos(half), axis_angle * s], dim=-1)
    r, i, j, k = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                     two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                     two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return o.reshape(q.shape[:-1] + (3, 3))
