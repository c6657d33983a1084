"""Known-answer tests of the restated e3nn 0.5.1 operators (oracle/e3nn_lite.py) that do NOT go through e3nn_lite's own tables:
every expected value below is either a closed form typed in by hand (spherical harmonics, delta / epsilon Wigner-3j symbols,
instruction order and path weights derived on paper) or a number read off the shipped checkpoint's serialised e3nn buffers
(`_w3j_1_2_1`).  Closes SURVEY 8c A1 (FullTensorProduct output order) and A2 (Y2 sign convention) as far as
they can be closed without e3nn itself: the oracle and the CUDA kernels are both generated from e3nn_lite, so a wrong recollection
there would otherwise be self-consistent.

Reference call sites: score_model_phore.py:123 (FullyConnectedTensorProduct), :276,:366 (FullTensorProduct), :365,:737 (spherical
harmonics)."""
import math
import os

import numpy as np
import torch

from oracle import e3nn_lite as e3

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# nonzeros of e3nn's _w3j_1_2_1 as serialised in the shipped checkpoint (read by the surveyor from the .pt file)
W3J_121 = {(0, 0, 2): 0.316228, (0, 1, 1): 0.316228, (1, 1, 0): 0.316228, (1, 3, 2): 0.316228, (2, 0, 0): 0.316228,
           (2, 3, 1): 0.316228, (2, 4, 2): 0.316228, (0, 4, 0): -0.316228, (0, 2, 0): -0.182574, (2, 2, 2): -0.182574,
           (1, 2, 1): 0.365148}


def c121():
    c = np.zeros((3, 5, 3))
    for k, v in W3J_121.items():
        c[k] = v
    return c


def test_w3j_file_equals_the_checkpoint_numbers_of_the_survey():
    z = np.load(os.path.join(ROOT, 'oracle', 'w3j.npz'))
    assert np.abs(z['w3j_1_2_1'] - c121()).max() < 1e-6
    eps = np.zeros((3, 3, 3))
    for i, j, k, s in ((0, 1, 2, 1), (1, 2, 0, 1), (2, 0, 1, 1), (0, 2, 1, -1), (2, 1, 0, -1), (1, 0, 2, -1)):
        eps[i, j, k] = s / math.sqrt(6)
    assert np.abs(z['w3j_1_1_1'] - eps).max() < 1e-6                       # Levi-Civita / sqrt(6)
    assert np.abs(z['w3j_2_2_0'][:, :, 0] - np.eye(5) / math.sqrt(5)).max() < 1e-6
    assert np.abs(z['w3j_0_2_2'][0] - np.eye(5) / math.sqrt(5)).max() < 1e-6


def test_spherical_harmonics_hand_values_and_y2_sign_convention():
    """Y(l<=2), 'component' normalisation, of v = (1, 2, 2) / 3 and of the axes, typed in from the closed forms of SURVEY A.1;
    the Y2 ORDER and SIGNS are tied to e3nn's own serialised w3j by the identity sum_ij C121_ijk x_i Y2_j(x) = sqrt(2/3) x_k."""
    v = torch.tensor([[1.0, 2.0, 2.0], [3.0, 0.0, 0.0], [0.0, -2.0, 0.0], [0.0, 0.0, 0.5], [0.0, 0.0, 0.0]], dtype=torch.float64)
    y = e3.spherical_harmonics(v).numpy()
    expect = np.array([
        [1, 0.5773502692, 1.1547005384, 1.1547005384, 0.8606629658, 0.8606629658, 0.3726779962, 1.7213259317, 0.6454972244],
        [1, 1.7320508076, 0, 0, 0, 0, -1.1180339887, 0, -1.9364916731],
        [1, 0, -1.7320508076, 0, 0, 0, 2.2360679775, 0, 0],
        [1, 0, 0, 1.7320508076, 0, 0, -1.1180339887, 0, 1.9364916731],
        [1, 0, 0, 0, 0, 0, 0, 0, 0]])                                       # zero vector: F.normalize -> zeros, Y0 = 1
    assert np.abs(y - expect).max() < 1e-9
    assert abs((y[0, 4:] ** 2).sum() - 5.0) < 1e-9 and abs((y[0, 1:4] ** 2).sum() - 3.0) < 1e-9
    rng = np.random.default_rng(0)
    x = rng.normal(size=(50, 3))
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    y2 = e3.spherical_harmonics(torch.from_numpy(x), only_l=2).numpy()
    lhs = np.einsum('ijk,zi,zj->zk', c121(), x, y2)
    assert np.abs(lhs - math.sqrt(2.0 / 3.0) * x).max() < 2e-6              # wrong order or sign of any Y2 component breaks this


def test_fctp_instruction_order_path_weights_and_closed_forms():
    """FullyConnectedTensorProduct(1x0e+1x1o, 1x0e+1x1o+1x2e -> 1x0e+1x1o+1x1e), 'uvw', per-edge weights.  Derived on paper:
    instructions in loop order (i1, i2, io): 0e0e->0e, 0e1o->1o, 1o0e->1o, 1o1o->0e, 1o1o->1e, 1o2e->1o; path weights
    sqrt((2 lo + 1) / fan_in(out)) = sqrt(1/2), 1, 1, sqrt(1/2), sqrt(3), 1; Wigner symbols delta/sqrt(2l+1), epsilon/sqrt(6), C121."""
    in1, in2, out = e3.parse_irreps('1x0e + 1x1o'), e3.sh_irreps(2), e3.parse_irreps('1x0e + 1x1o + 1x1e')
    instrs, numel = e3.fctp_instructions(in1, in2, out)
    assert numel == 6 and [(i.i1, i.i2, i.io) for i in instrs] == [(0, 0, 0), (0, 1, 1), (1, 0, 1), (1, 1, 0), (1, 1, 2), (1, 2, 1)]
    assert np.allclose([i.pw for i in instrs], [math.sqrt(0.5), 1, 1, math.sqrt(0.5), math.sqrt(3), 1])
    rng = np.random.default_rng(1)
    E = 7
    x1, sh, w = rng.normal(size=(E, 4)), rng.normal(size=(E, 9)), rng.normal(size=(E, 6))
    got = e3.fctp_apply(in1, in2, out, instrs, torch.from_numpy(x1), torch.from_numpy(sh), torch.from_numpy(w)).numpy()
    a0, a = x1[:, 0], x1[:, 1:4]
    s0, s, S2 = sh[:, 0], sh[:, 1:4], sh[:, 4:9]
    r3 = math.sqrt(3)
    o0 = math.sqrt(0.5) * (w[:, 0] * a0 * s0 + w[:, 3] * (a * s).sum(1) / r3)
    o1 = (w[:, 1] * a0 / r3)[:, None] * s + (w[:, 2] * s0 / r3)[:, None] * a + w[:, 5][:, None] * np.einsum('ijk,zi,zj->zk', c121(), a, S2)
    o2 = (r3 * w[:, 4] / math.sqrt(6))[:, None] * np.cross(a, s)
    assert np.abs(got - np.concatenate([o0[:, None], o1, o2], 1)).max() < 5e-6


def test_fctp_layer_tables_equal_the_survey_numbers():
    """Weight offsets and path weights of the shipped layer-1 / layer-3 / tor_bond_conv convolutions as SURVEY A.3 lists them (the
    offsets are pinned by the checkpoint's fc.3 row counts 1100 / 2200 / 1600)."""
    seq = [e3.parse_irreps(s) for s in ('20x0e', '20x0e + 10x1o', '20x0e + 10x1o + 10x1e', '20x0e + 10x1o + 10x1e + 20x0o')]
    ins, n = e3.fctp_instructions(seq[1], e3.sh_irreps(2), seq[2])
    assert n == 1100 and [i.w_off for i in ins] == [0, 400, 600, 700, 900, 1000]
    assert np.allclose([i.pw for i in ins], [.182574, .273861, .273861, .182574, .547723, .273861], atol=1e-6)
    ins, n = e3.fctp_instructions(seq[3], e3.sh_irreps(2), seq[3])
    assert n == 2200 and [i.w_off for i in ins] == [0, 400, 600, 700, 900, 1000, 1100, 1200, 1300, 1500, 1600, 2000]
    assert np.allclose([i.pw for i in ins][:3] + [i.pw for i in ins][-2:], [.182574, .244949, .244949, .182574, .244949], atol=1e-6)
    sh45, _ = e3.full_tp_irreps_out(e3.sh_irreps(2), [(1, 2, 1)])
    ins, n = e3.fctp_instructions(seq[3], sh45, e3.parse_irreps('20x0o + 20x0e'))
    assert n == 1600 and [i.w_off for i in ins] == [0, 400, 600, 800, 1000, 1200] and np.allclose([i.pw for i in ins], .158114, atol=1e-6)


def test_full_tensor_product_output_order_and_values():
    """FullTensorProduct(1x0e+1x1o+1x2e, 1x2e): nine output irreps sorted by (l, parity) with odd before even at equal l and the
    0e x 2e -> 2e block BEFORE the 2e x 2e -> 2e block (stable sort), 45 components; out = sqrt(2 lo + 1) sum_ij C_ijk a_i b_j.
    Closed forms: 0e = (A2 . B2) / sqrt(5); the first 2e block = a0 * B2; 1o through C121."""
    irreps, out = e3.full_tp_apply(e3.sh_irreps(2), [(1, 2, 1)], torch.zeros(1, 9, dtype=torch.float64), torch.zeros(1, 5, dtype=torch.float64))
    assert irreps == [(1, 0, 1), (1, 1, -1), (1, 1, 1), (1, 2, -1), (1, 2, 1), (1, 2, 1), (1, 3, -1), (1, 3, 1), (1, 4, 1)] and out.shape[1] == 45
    rng = np.random.default_rng(2)
    a, b = rng.normal(size=(6, 9)), rng.normal(size=(6, 5))
    _, got = e3.full_tp_apply(e3.sh_irreps(2), [(1, 2, 1)], torch.from_numpy(a), torch.from_numpy(b))
    got = got.numpy()
    assert np.abs(got[:, 0] - (a[:, 4:] * b).sum(1) / math.sqrt(5)).max() < 1e-9
    assert np.abs(got[:, 1:4] - math.sqrt(3) * np.einsum('ijk,zi,zj->zk', c121(), a[:, 1:4], b)).max() < 5e-6
    assert np.abs(got[:, 12:17] - a[:, :1] * b).max() < 1e-9                  # 2e of 0e x 2e sits at components 12..16
    # 1e of 2e x 2e: antisymmetric in its arguments (swapping A2 and B2 flips the sign), and orthogonal to nothing else to check here
    a2 = a.copy(); a2[:, 4:] = b
    _, sw = e3.full_tp_apply(e3.sh_irreps(2), [(1, 2, 1)], torch.from_numpy(a2), torch.from_numpy(a[:, 4:].copy()))
    assert np.abs(sw.numpy()[:, 4:7] + got[:, 4:7]).max() < 1e-9
